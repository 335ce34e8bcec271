"""BCQP loops of the oracle against scipy, in the style of SimToolbox/Constraint/BCQPSolver_verify.py:19-57
(random SPD B^T D B problem with random bounds; the reference only prints the error norms -- here they are
asserted).  Also pins the structural quirks listed in SURVEY.md appendix A."""
import numpy as np
import pytest
import scipy.optimize as so
import scipy.sparse as sp

from scenarios import random_rods, thermal_velocity


def random_problem(n, seed, diag=0.0):
    rng = np.random.default_rng(seed)
    B = rng.uniform(-1, 1, size=(n, n))
    D = np.diag(10 ** rng.uniform(-1, 1, size=n))
    A = B.T @ D @ B + diag * np.eye(n)
    b = rng.uniform(-1, 1, size=n)
    v1, v2 = rng.uniform(-1, 1, size=n), rng.uniform(-1, 1, size=n)
    return A, b, np.minimum(v1, v2), np.maximum(v1, v2)


@pytest.mark.parametrize("choice", [0, 1])
def test_bcqp_against_lbfgsb(oracle, choice):
    A, b, lb, ub = random_problem(60, 7, diag=0.5)
    # 1e-6 is reachable; far below that both loops end in their "Stagnate" branch (BCQPSolver.cpp:229,332)
    rc, x, hist = oracle.bcqp_csr(sp.csr_matrix(A), b, lb, ub, np.zeros(60), 1e-6, 20000, choice)
    assert rc == 0 and hist[-1, 4] < 1e-6
    ref = so.minimize(lambda z: 0.5 * z @ A @ z + b @ z, np.zeros(60), jac=lambda z: A @ z + b, method="L-BFGS-B",
                      bounds=list(zip(lb, ub)), options=dict(maxiter=20000, ftol=1e-16, gtol=1e-12))
    assert np.abs(x - ref.x).max() < 1e-5
    assert np.all(x >= lb) and np.all(x <= ub)
    # history rows: {ite, 0, 0, step, resPhi, mvCount} (BCQPSolver.hpp:23)
    assert hist[0, 0] == 0 and np.all(np.diff(hist[:, 0]) == 1) and np.all(hist[:, 1:3] == 0)
    if choice == 0:
        assert np.array_equal(hist[:, 5], hist[:, 0] + 1)  # one mat-vec per BBPGD iteration


def test_default_bounds_nnls(oracle):
    # the constraint problem is min 1/2 x^T A x + b^T x, x >= 0 (unilateral rows): compare with NNLS on a
    # factorised version  A = C^T C,  b = -C^T d
    rng = np.random.default_rng(2)
    C = rng.normal(size=(80, 40))
    d = rng.normal(size=80)
    A, b = C.T @ C, -C.T @ d
    lb = np.full(40, -0.0)
    ub = np.full(40, np.finfo(float).max / 10)
    rc, x, hist = oracle.bcqp_csr(sp.csr_matrix(A), b, lb, ub, np.zeros(40), 1e-7, 50000, 0)
    xn, _ = so.nnls(C, d)
    assert rc == 0 and np.abs(x - xn).max() < 1e-6


def test_solve_constraints_structure(oracle):
    rods = random_rods(500, 1.0, seed=1, frac_immovable=0.1)
    lo, hi, pbc = [0] * 3, [1.0] * 3, (1, 1, 1)
    orods = oracle.make_rods(rods["gid"], rods["radius"], rods["length"], rods["pos"], rods["quat"], colBuf=0.025)
    blocks = oracle.collect_pairs(orods, lo, hi, pbc)
    assert len(blocks) > 300
    mu, dt = 1.0, 1e-4
    vnc = thermal_velocity(rods, mu, dt, seed=2)
    DT, d0, ik, bi, g0 = oracle.build_dtrans_dense(blocks, len(orods))
    M = oracle.build_mobility(orods, rods["immovable"], mu)
    # D^T rows: 12 entries [n, r x n] per block (ConstraintCollector.cpp:298-341)
    assert np.all(np.diff(DT.indptr) == 12)
    k = 5
    row = DT.getrow(k).toarray().ravel()
    I, J = blocks[k]["globalIndexI"], blocks[k]["globalIndexJ"]
    assert np.allclose(row[6 * I:6 * I + 3], blocks[k]["normI"])
    assert np.allclose(row[6 * I + 3:6 * I + 6], np.cross(blocks[k]["posI"], blocks[k]["normI"]))
    assert np.allclose(row[6 * J + 3:6 * J + 6], np.cross(blocks[k]["posJ"], blocks[k]["normJ"]))
    # mobility blocks are SPD for movable rods and zero for immovable ones (SylinderSystem.cpp:660-665)
    Md = M.toarray()
    for i in range(0, 20):
        blk = Md[6 * i:6 * i + 6, 6 * i:6 * i + 6]
        if rods["immovable"][i]:
            assert np.all(blk == 0)
        else:
            assert np.linalg.eigvalsh(blk).min() > 0
    # convergence check with all rods movable (two overlapping immovable rods make the QP unbounded)
    imm0 = np.zeros_like(rods["immovable"])
    M = oracle.build_mobility(orods, imm0, mu)
    res = 1e-6
    sol = oracle.solve_constraints(blocks, orods, imm0, mu, vnc, dt, res, 50000, 0)
    assert sol["rc"] == 0
    A = (DT @ M @ DT.T).toarray()
    q = d0 / dt + DT @ vnc
    x = sol["gamma"]
    grad = A @ x + q
    assert x.min() >= 0 and np.abs(np.where(x > 0, grad, np.minimum(grad, 0))).max() < 1.01 * res / dt
    # velocities/forces are consistent with gamma (ConstraintSolver.cpp:95-106, no bilateral rows here)
    assert np.allclose(sol["forceU"], DT.T @ x, rtol=0, atol=1e-9 * np.abs(sol["forceU"]).max())
    assert np.allclose(sol["velU"], M @ (DT.T @ x), rtol=0, atol=1e-9 * np.abs(sol["velU"]).max())
    assert np.all(sol["forceB"] == 0) and np.all(sol["velB"] == 0)
    # APGD reaches the same velocities
    sol2 = oracle.solve_constraints(blocks, orods, imm0, mu, vnc, dt, res, 50000, 1)
    assert sol2["rc"] == 0 and np.abs(sol2["velU"] - sol["velU"]).max() < 100 * res / dt


def test_itemax_quirk(oracle):
    # BCQPSolver.cpp:237-241: on the iteMax exit the returned iterate is the older one
    A, b, lb, ub = random_problem(30, 3, diag=0.1)
    As = sp.csr_matrix(A)
    _, x3, h3 = oracle.bcqp_csr(As, b, lb, ub, np.zeros(30), 1e-30, 3, 0)
    _, x2, h2 = oracle.bcqp_csr(As, b, lb, ub, np.zeros(30), 1e-30, 2, 0)
    _, x4, h4 = oracle.bcqp_csr(As, b, lb, ub, np.zeros(30), 1e-30, 4, 0)
    assert len(h3) == 4 and len(h2) == 3
    # x returned with iteMax=3 is iterate #2; with iteMax=4 iterate #3; iterates are distinct
    assert not np.array_equal(x3, x4)
    # verify by direct re-computation of iterate 2 from the recorded step sizes
    x = np.zeros(30)
    g = A @ x + b
    xs = [x]
    for it in range(1, 4):
        alpha = h4[it, 3]
        x = np.clip(xs[-1] - alpha * g, lb, ub)
        g = A @ x + b
        xs.append(x)
    assert np.allclose(x3, xs[2], atol=1e-14) and np.allclose(x4, xs[3], atol=1e-14)
