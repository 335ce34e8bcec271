"""BCQPSolver as ANY caller of the reference sees it (SimToolbox/Constraint/BCQPSolver.hpp:37-111), on the device:
a CSR matrix or the matrix-free constraint operator, the caller's b, the caller's bounds (setLowerBound / setUpperBound),
BBPGD and APGD -- against the reference's own BCQPSolver.cpp (oracle/_ref/libalens_refsys.so) on the same problems,
starting with the reference's own self-test problem BCQPSolver(int, double) (BCQPSolver.cpp:38-132, BCQPSolver_test.cpp)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import alens_b200
from scenarios import random_rods, thermal_velocity
from test_gpu_collect import gpu_collect

from oracle import pyrefsys as pr

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not pr.available(), reason="oracle/_ref/libalens_refsys.so missing (built where /root/reference exists)")]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def relerr(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("choice", [0, 1])
def test_reference_self_test_problem_on_the_device(ctx, choice):
    """the reference's own random SPD problem with random bounds (dumped by its selfTest), solved by the device BCQP:
    same iteration count, same history rows and solution to 1e-8"""
    import scipy.sparse as sp

    with tempfile.TemporaryDirectory() as d:
        p = pr.bcqp_selftest(d, 80, 0.5, 1e-7, 3000, choice)
    A = sp.csr_matrix(p["A"])
    q = alens_b200.Bcqp(ctx, b=p["b"], csr=(A.indptr, A.indices, A.data))
    lb0, ub0 = q.get_bounds()
    assert np.all(lb0 == -np.finfo(float).max / 10) and np.all(ub0 == np.finfo(float).max / 10)  # setDefaultBounds
    q.set_bounds(p["lb"], p["ub"])
    x, rep, hist = q.solve(np.zeros(80), 1e-7, 3000, choice)
    xr, hr, rcr = pr.bcqp_solve_csr(A.indptr, A.indices, A.data, p["b"], p["lb"], p["ub"], np.zeros(80), 1e-7, 3000, choice)
    assert np.array_equal(xr, p["x"])  # what the reference's selfTest itself wrote
    young = min(len(hist), len(hr), 40)  # BB steps amplify rounding differences: compare the young rows tightly
    np.testing.assert_allclose(hist[:young, 3:5], hr[:young, 3:5], rtol=1e-7, atol=1e-12)
    assert np.array_equal(hist[:young, 5], hr[:young, 5])
    # the problem is drawn from std::random_device (BCQPSolver.cpp:46-47): now and then one of the two runs ends on its
    # stagnation test (step size below 10 eps) a few iterations before the other would converge, which is a matter of rounding
    assert rep.status in (0, 1) and rcr in (0, 1)
    if rcr == 0 and rep.status == 0:
        assert abs(len(hist) - len(hr)) <= max(3, 0.1 * len(hr))
        assert np.abs(x - xr).max() < 1e-5  # both satisfy the same KKT tolerance (A is well conditioned: diag 0.5)
        assert np.all(x >= p["lb"]) and np.all(x <= p["ub"])
    q.close()


def test_fixed_iteration_parity_on_a_csr_problem(ctx):
    """equal iteration counts: iterates agree to 1e-8 (north_star tolerance)"""
    import scipy.sparse as sp

    rng = np.random.default_rng(3)
    n = 400
    B = sp.random(n, n, density=0.02, random_state=5, format="csr")
    A = (B.T @ B + sp.identity(n) * 0.3).tocsr()
    A.sort_indices()
    b = rng.uniform(-1, 1, n)
    lb, ub = np.minimum(*rng.uniform(-1, 1, (2, n))), None
    lb = np.minimum(rng.uniform(-1, 1, n), 0.0)
    ub = np.maximum(rng.uniform(-1, 1, n), 0.2)
    for choice, ite in ((0, 30), (1, 15)):
        q = alens_b200.Bcqp(ctx, b=b, csr=(A.indptr, A.indices, A.data))
        q.set_bounds(lb, ub)
        x, rep, hist = q.solve(np.zeros(n), 1e-30, ite, choice)
        xr, hr, _ = pr.bcqp_solve_csr(A.indptr, A.indices, A.data, b, lb, ub, np.zeros(n), 1e-30, ite, choice)
        assert rep.iterations == ite and len(hist) == len(hr)
        assert relerr(x, xr) < 1e-8
        np.testing.assert_allclose(hist[:, 3:5], hr[:, 3:5], rtol=1e-7)
        q.close()


def test_constraint_operator_with_caller_bounds(ctx, oracle):
    """BCQPSolver(ConstraintOperator, q) with bounds the CALLER sets -- an upper bound that is active, and a relaxed lower
    bound -- against the reference's BCQPSolver on the explicit matrix D^T M D of the same list"""
    import scipy.sparse as sp

    MU, DT = 1.0, 1e-4
    rods = random_rods(900, 1.2, seed=14, frac_sphere=0.1)
    lo, hi, pbc = [0, 0, 0], [1.2] * 3, (1, 1, 1)
    blocks = gpu_collect(ctx, rods, lo, hi, pbc, 0.025).copy()
    nc = len(blocks)
    ctx.calc_mobility(MU)
    vnc = thermal_velocity(rods, MU, DT, seed=2)
    ctx.setup_constraints(vnc, DT)
    orods = oracle.make_rods(rods["gid"], rods["radius"], rods["length"], oracle.wrap_positions(rods["pos"], lo, hi, pbc),
                             rods["quat"], 1.0, 1.0, 0.025)
    DTm, d0, ik, bi, g0 = oracle.build_dtrans_dense(blocks, len(orods))
    M = oracle.build_mobility(orods, rods["immovable"], MU)
    A = (DTm @ M @ DTm.T).tocsr()
    A.sort_indices()
    qvec = d0 / DT + DTm @ vnc
    # bounds on the scale of the multipliers themselves (a free solve gives it): a cap that is active and a slightly negative floor
    big = np.finfo(float).max / 10
    free, _, _ = pr.bcqp_solve_csr(A.indptr, A.indices, A.data, qvec, np.zeros(nc), np.full(nc, big), g0, 1e-30, 60, 0)
    rng = np.random.default_rng(0)
    lb = -0.02 * rng.uniform(size=nc) * free.max()
    ub = np.full(nc, 0.3 * free.max())
    x0 = np.clip(g0, lb, ub)
    q = alens_b200.Bcqp(ctx)  # the constraint operator of the setup, b = its q
    q.set_bounds(lb, ub)
    for choice, ite in ((0, 25), (1, 12)):
        x, rep, hist = q.solve(x0, 1e-30, ite, choice)
        xr, hr, _ = pr.bcqp_solve_csr(A.indptr, A.indices, A.data, qvec, lb, ub, x0, 1e-30, ite, choice)
        assert rep.iterations == len(hr) - 1 == ite
        assert np.all(x >= lb) and np.all(x <= ub) and (x == ub).sum() > 10 and (x < 0).sum() > 100  # the caller's bounds
        assert relerr(x, xr) < 1e-7
        np.testing.assert_allclose(hist[:, 4], hr[:, 4], rtol=1e-6)
    # a projection error is reported, not ignored (BCQPSolver.cpp:484-494).  With finite numbers the three branches of the
    # projected gradient cover every case; the error branch is what a NaN iterate falls into
    bad = x0.copy()
    bad[3] = np.nan
    with pytest.raises(alens_b200.AlensError) as ei:
        q.solve(bad, 1e-6, 5, 0)
    assert ei.value.code == -5
    q.close()


def test_cpp_bcqpsolver_mirror_selftest(tmp_path):
    """include/alens_b200/BCQPSolver.hpp: BCQPSolver(localSize, diagonal, ctx, seed) + selfTest through the C++ mirror;
    its dumped problem solved again by the reference's BCQPSolver"""
    exe = os.path.join(ROOT, "tests", "cpp", "test_bcqp")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")])
    import scipy.sparse as sp

    for choice in (0, 1):
        out = tmp_path / f"bcqp{choice}.bin"
        r = subprocess.run([exe, "60", "0.5", "77", str(choice), str(out)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        tag = "APGD_HISTORY," if choice else "BBPGD_HISTORY,"
        rows = [ln for ln in r.stdout.splitlines() if ln.startswith(tag)]
        raw = np.fromfile(out, dtype=np.float64)
        n = 60
        A, b, lb, ub, x = (raw[:n * n].reshape(n, n), raw[n * n:n * n + n], raw[n * n + n:n * n + 2 * n],
                           raw[n * n + 2 * n:n * n + 3 * n], raw[n * n + 3 * n:n * n + 4 * n])
        assert np.abs(A - A.T).max() < 1e-12 and np.all(np.linalg.eigvalsh(A) > 0) and np.all(lb <= ub)
        S = sp.csr_matrix(np.where(np.abs(A) > 1e-7, A, 0.0))
        xr, hr, rcr = pr.bcqp_solve_csr(S.indptr, S.indices, S.data, b, lb, ub, np.zeros(n), 1e-7, 3000, choice)
        assert abs(len(rows) - len(hr)) <= max(3, 0.1 * len(hr))
        if rcr == 0:
            assert np.abs(x - xr).max() < 1e-5
        assert np.all(x >= lb) and np.all(x <= ub)


def test_handle_outliving_its_context_is_retired_not_dangling(alens_lib):
    """alens_destroy retires the BCQP handles still open on the context: later calls answer ALENS_ERR_STATE (-3,
    no crash), alens_bcqp_destroy frees the shell"""
    import scipy.sparse as sp

    c = alens_b200.Context(device=0)
    A = sp.identity(8, format="csr")
    q = alens_b200.Bcqp(c, b=-np.ones(8), csr=(A.indptr, A.indices, A.data))
    x, rep, _ = q.solve(np.zeros(8), 1e-12, 50, 0)
    assert np.allclose(x, 1.0)
    c.close()
    rep = alens_b200.capi.SolveReport()
    import ctypes as C
    xx = np.zeros(8)
    rc = q.dll.alens_bcqp_run(q.h, xx.ctypes.data_as(C.POINTER(C.c_double)), C.c_double(1e-6), C.c_int(5), C.c_int(0), C.byref(rep))
    assert rc == -3
    q.close()
