"""The oracle (oracle/alens_oracle.c, numpy helpers) pinned against the reference's OWN code: SylinderSystem.cpp,
SylinderNear.hpp, Constraint/{ConstraintCollector,ConstraintSolver,ConstraintOperator,BCQPSolver}.cpp, Boundary.cpp,
Sylinder.cpp compiled unmodified into oracle/_ref/libalens_refsys.so against the stand-in headers of oracle/stubs
(recipe: oracle/Makefile `refsys`, driver: oracle/ref_system_driver.cpp).  CPU only.

Bit for bit: pair blocks (all fields + stress) on P_ref, boundary and link blocks, D^T/operator, mobility, gamma, the
uni/bi force and velocity split, every IteHistory row (BBPGD and APGD).  The committed fixture
tests/golden/refsolver.npz holds the reference's outputs for the three example configurations, so the solver half stays
pinned where the library is absent."""
import os
import tempfile

import numpy as np
import pytest

from scenarios import canonical_order, quat_from_z_to, random_rods, thermal_velocity

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")

from oracle import pyrefsys as pr  # noqa: E402

needs_ref = pytest.mark.skipif(not pr.available(), reason="oracle/_ref/libalens_refsys.so not built (needs /root/reference)")

BLOCK_FIELDS = ("delta0", "gamma", "gammaLB", "gidI", "gidJ", "globalIndexI", "globalIndexJ", "oneSide", "bilateral", "kappa",
                "normI", "normJ", "posI", "posJ", "labI", "labJ", "stress")


def _system(rods, lo, hi, pbc, colbuf=0.025, mu=1.0, dt=1e-4, nthreads=1, boundaries=(), **cfg):
    c = dict(simBoxLow=list(map(float, lo)), simBoxHigh=list(map(float, hi)), simBoxPBC=[bool(x) for x in pbc],
             sylinderColBuf=colbuf, viscosity=mu, dt=dt)
    c.update(cfg)
    s = pr.RefSystem(c, nthreads=nthreads, boundaries=boundaries)
    s.set_rods(rods)
    s.prepare_step()
    return s


def _orods(oracle, rods, lo, hi, pbc, colbuf, dratio=1.0, lratio=1.0):
    return oracle.make_rods(rods["gid"], rods["radius"], rods["length"], oracle.wrap_positions(rods["pos"], lo, hi, pbc),
                            rods["quat"], dratio, lratio, colbuf)


@needs_ref
@pytest.mark.parametrize("pbc,frac_sphere,nthreads", [((1, 1, 1), 0.0, 1), ((0, 0, 0), 0.3, 2), ((1, 0, 1), 0.15, 3)])
def test_pair_collection_is_the_reference_functor_bit_for_bit(oracle, pbc, frac_sphere, nthreads):
    """prepareStep + collectPairCollision of the reference (FDPS tree, CalcSylinderNearForce, collideStress): every block
    it finds is in the oracle's geometric list with identical fields; what it misses lies beyond its search radius"""
    n, box, colbuf = 3000, 1.8, 0.025
    rods = random_rods(n, box, seed=42 + sum(pbc), frac_sphere=frac_sphere)
    rods["pos"] = rods["pos"] * 1.3 - 0.2  # some rods outside the box: applyBoxBC wraps periodic axes only
    lo, hi = [0.0] * 3, [box] * 3
    s = _system(rods, lo, hi, pbc, colbuf, nthreads=nthreads, sylinderDiameterColRatio=1.1, sylinderLengthColRatio=0.95)
    sy = s.sylinders()
    assert np.array_equal(sy["gid"], rods["gid"])
    assert np.array_equal(sy["pos"], oracle.wrap_positions(rods["pos"], lo, hi, pbc))
    assert np.array_equal(sy["globalIndex"], np.arange(n))
    nref = s.collect_pair_collision()
    ref = s.constraints()
    geo = oracle.collect_pairs(_orods(oracle, rods, lo, hi, pbc, colbuf, 1.1, 0.95), lo, hi, pbc, with_stress=True)
    assert 0 < nref <= len(geo)
    # a pair may appear once per periodic image: key on (gidI, gidJ, labJ)
    key = lambda b: list(zip(b["gidI"].tolist(), b["gidJ"].tolist(), map(bytes, np.ascontiguousarray(b["labJ"]))))
    kg = {k: i for i, k in enumerate(key(geo))}
    assert len(set(key(ref))) == nref
    sel = geo[[kg[k] for k in key(ref)]]
    for f in BLOCK_FIELDS:
        assert np.array_equal(sel[f], ref[f]), f
    assert len(geo) - nref < 0.1 * len(geo)
    s.close()


@needs_ref
def test_functor_single_pairs_against_reference(oracle):
    """CalcSylinderNearForce::operator() itself (no tree) on pairs incl. spheres and coincident centres"""
    rng = np.random.default_rng(5)
    n = 400
    rods = random_rods(2 * n, 1.0, seed=3, frac_sphere=0.3, length=0.5)
    rods["pos"][1::2] = rods["pos"][0::2] + rng.normal(size=(n, 3)) * 0.15
    rods["gid"] = np.arange(2 * n, dtype=np.int32)
    o = oracle.make_rods(rods["gid"], rods["radius"], rods["length"], rods["pos"], rods["quat"], 1.0, 1.0, 0.05)
    o["pos"][1] = o["pos"][0]
    o["lengthCollision"][:2] = 0.0  # two coincident spheres
    hits = 0
    for k in range(n):
        want = pr.pair_functor(o[2 * k], o[2 * k + 1])
        got = oracle.pair_functor(o[2 * k], o[2 * k + 1], with_stress=True)
        assert (want is None) == (got is None), k
        if want is None:
            continue
        hits += 1
        for f in BLOCK_FIELDS:
            assert np.array_equal(want[f], got[f], equal_nan=True), (k, f)
    assert hits > 100
    assert np.all(pr.pair_functor(o[0], o[1])["normI"] == 0)  # Eigen's normalized() leaves a zero vector alone


@needs_ref
def test_boundary_and_link_blocks_against_reference(oracle):
    n, box, colbuf = 1500, 3.0, 0.025
    rods = random_rods(n, box, seed=23, frac_sphere=0.15, length=0.4)
    lo, hi, pbc = [0.0] * 3, [box] * 3, (1, 1, 0)
    bnd = [dict(type="wall", center=[0.0, 0.0, 0.3], norm=[0.0, 0.0, 1.0]),
           dict(type="sphere", center=[1.5, 1.5, 1.5], radius=1.4, inside=True),
           dict(type="tube", center=[1.5, 1.5, 0.0], axis=[0.0, 0.1, 1.0], radius=1.3, inside=False)]
    s = _system(rods, lo, hi, pbc, colbuf, boundaries=bnd, linkKappa=250.0, linkGap=0.02)
    s2 = None
    prev = np.arange(0, 600, 2)
    nxt = prev + 1
    s.add_links(rods["gid"][prev], rods["gid"][nxt])
    s.prepare_step()
    assert s.collect_boundary_collision() > 500
    B = s.constraints().copy()
    s.clear_constraints()
    assert s.collect_link_bilateral() == len(prev)
    Lk = s.constraints().copy()
    orods = _orods(oracle, rods, lo, hi, pbc, colbuf)
    ob = oracle.make_boundaries([dict(type="wall", center=[0.0, 0.0, 0.3], axis=[0.0, 0.0, 1.0]),
                                 dict(type="sphere", center=[1.5, 1.5, 1.5], radius=1.4, inside=True),
                                 dict(type="tube", center=[1.5, 1.5, 0.0], axis=[0.0, 0.1, 1.0], radius=1.3, inside=False)])
    OB = oracle.collect_boundary(orods, ob, colbuf)
    OL = oracle.collect_links(orods, rods["gid"][prev], rods["gid"][nxt], lo, hi, pbc, 250.0, 0.02)

    def same(a, b):
        assert len(a) == len(b)
        ka = np.lexsort((a["labJ"][:, 2], a["labJ"][:, 1], a["labJ"][:, 0], a["labI"][:, 0], a["gidJ"], a["gidI"]))
        kb = np.lexsort((b["labJ"][:, 2], b["labJ"][:, 1], b["labJ"][:, 0], b["labI"][:, 0], b["gidJ"], b["gidI"]))
        for f in BLOCK_FIELDS:
            assert np.array_equal(a[ka][f], b[kb][f]), f

    same(B, OB)
    same(Lk, OL)
    # Boundary::project itself
    rng = np.random.default_rng(0)
    for spec, ospec in zip(bnd, ob):
        for q in rng.normal(size=(50, 3)) * 2:
            p1, d1 = pr.boundary_project(spec["type"], spec["center"], spec.get("norm", spec.get("axis", [0, 0, 1])),
                                         spec.get("radius", 0.0), spec.get("inside", True), q)
            p2, d2 = oracle.boundary_project(ospec, q)
            assert np.array_equal(p1, p2) and np.array_equal(d1, d2)
    s.close()


@needs_ref
def test_mobility_drag_and_operator_against_reference(oracle):
    n, box, colbuf, mu, dt = 1200, 1.4, 0.03, 0.8, 1e-4
    rods = random_rods(n, box, seed=9, frac_sphere=0.2, frac_immovable=0.1, length_sigma=0.3)
    lo, hi, pbc = [0.0] * 3, [box] * 3, (1, 1, 1)
    s = _system(rods, lo, hi, pbc, colbuf, mu=mu, dt=dt)
    orods = _orods(oracle, rods, lo, hi, pbc, colbuf)
    for L, R in ((0.25, 0.0125), (0.01, 0.02), (3.0, 0.5)):
        assert pr.drag_coeff(L, R, mu) == oracle.drag_coeff(R, L, mu)  # Sylinder::calcDragCoeff
    x = np.random.default_rng(0).normal(size=6 * n)
    M = oracle.build_mobility(orods, rods["immovable"], mu)
    y = np.array([sum(M.data[p] * x[M.indices[p]] for p in range(M.indptr[r], M.indptr[r + 1])) for r in range(6 * n)])
    assert np.array_equal(s.mobility_apply(x), y)  # calcMobMatrix
    geo = oracle.collect_pairs(orods, lo, hi, pbc, with_stress=True)
    blocks = geo[canonical_order(geo)]
    g = np.abs(np.random.default_rng(1).normal(size=len(blocks)))
    y1, f1, v1 = s.operator_apply(blocks, dt, g)  # buildConstraintMatrixVector + ConstraintOperator::apply
    y2, f2, v2 = oracle.operator_apply(blocks, orods, rods["immovable"], mu, dt, g)
    assert np.array_equal(y1, y2) and np.array_equal(f1, f2) and np.array_equal(v1, v2)
    s.close()


def _solve_both(oracle, s, blocks, orods, imm, mu, vnc, dt, res, max_ite, choice):
    r = s.solve_blocks(blocks, vnc, dt, res, max_ite, choice)
    o = oracle.solve_constraints(blocks, orods, imm, mu, vnc, dt, res, max_ite, choice)
    return r, o


def _assert_same_solve(r, o):
    assert r["nIte"] == o["nIte"]
    assert np.array_equal(r["history"], o["history"])  # every row {ite, 0, 0, alpha, resPhi, mvCount}
    assert np.array_equal(r["gamma"], o["gamma"])
    for k in ("forceU", "velU", "forceB", "velB"):
        assert np.array_equal(r[k], o[k]), k


@needs_ref
@pytest.mark.parametrize("choice", [0, 1])
@pytest.mark.parametrize("max_ite,res", [(0, 1e-6), (1, 1e-30), (2, 1e-30), (37, 1e-30), (800, 1e-6)])
def test_solver_half_bit_for_bit_against_reference(oracle, choice, max_ite, res):
    """ConstraintSolver::setup/solveConstraints/writebackGamma + BCQPSolver::solveBBPGD/solveAPGD of the reference on the
    oracle's list (collisions + links + one-sided boundary blocks, immovable rods): gamma, split, history rows"""
    n, box, colbuf, mu, dt = 1500, 1.5, 0.025, 1.0, 1e-4
    rods = random_rods(n, box, seed=11, frac_sphere=0.1, frac_immovable=0.05)
    lo, hi, pbc = [0.0] * 3, [box] * 3, (1, 1, 0)
    s = _system(rods, lo, hi, pbc, colbuf, mu=mu, dt=dt)
    orods = _orods(oracle, rods, lo, hi, pbc, colbuf)
    coll = oracle.collect_pairs(orods, lo, hi, pbc, with_stress=True)
    coll = coll[canonical_order(coll)]
    prev = np.arange(0, 200, 2)
    links = oracle.collect_links(orods, rods["gid"][prev], rods["gid"][prev + 1], lo, hi, pbc, 300.0, 0.01)
    walls = oracle.collect_boundary(orods, oracle.make_boundaries([dict(type="wall", center=[0, 0, 0.2], axis=[0, 0, 1.0])]), colbuf)
    blocks = np.concatenate([coll, walls, links])
    assert blocks["bilateral"].sum() == 100 and blocks["oneSide"].sum() > 50
    vnc = thermal_velocity(rods, mu, dt, seed=1)
    r, o = _solve_both(oracle, s, blocks, orods, rods["immovable"], mu, vnc, dt, res, max_ite, choice)
    _assert_same_solve(r, o)
    assert len(r["history"]) >= min(max_ite, 1) + 1 and r["status"] == 0  # (convergence: the example configs below)
    # writebackGamma: gamma into the blocks, stress scaled (ConstraintCollector.cpp:439-461)
    wb = oracle.writeback_gamma(blocks, o["gamma"])
    assert np.array_equal(wb["gamma"], r["blocks"]["gamma"]) and np.array_equal(wb["stress"], r["blocks"]["stress"])
    s.close()


@needs_ref
def test_reference_self_test_problem(oracle):
    """BCQPSolver(int, double) + selfTest of the reference (BCQPSolver.cpp:38-132,391-429, BCQPSolver_test.cpp): its random
    SPD problem with random bounds, dumped as MatrixMarket files, solved again by the oracle: same iterate bit for bit"""
    import scipy.sparse as sp

    for choice in (0, 1):
        with tempfile.TemporaryDirectory() as d:
            p = pr.bcqp_selftest(d, 60, 0.5, 1e-7, 4000, choice)
        A = sp.csr_matrix(p["A"])
        assert np.all(p["lb"] <= p["ub"]) and np.abs(p["A"] - p["A"].T).max() < 1e-12
        rc, x, hist = oracle.bcqp_csr(A, p["b"], p["lb"], p["ub"], np.zeros(60), 1e-7, 4000, choice)
        assert np.array_equal(x, p["x"])
        xr, hr, rcr = pr.bcqp_solve_csr(A.indptr, A.indices, A.data, p["b"], p["lb"], p["ub"], np.zeros(60), 1e-7, 4000, choice)
        assert np.array_equal(xr, p["x"]) and np.array_equal(hr, hist) and rc == rcr
        if rc == 0 and len(hist) <= 4000:  # not stagnated (the random problem differs from run to run: std::random_device)
            assert hist[-1][4] < 1e-7
            g = A @ x + p["b"]  # KKT: the projected gradient vanishes
            q = np.where(x <= p["lb"] + 1e-12, np.minimum(g, 0), np.where(x >= p["ub"] - 1e-12, np.maximum(g, 0), g))
            assert np.abs(q).max() < 1e-6


@needs_ref
def test_velocities_and_euler_step_against_reference(oracle):
    n, box, mu, dt, kbt = 600, 1.5, 0.7, 1e-4, 0.00411
    rods = random_rods(n, box, seed=4, frac_sphere=0.2, frac_immovable=0.1)
    lo, hi, pbc = [0.0] * 3, [box] * 3, (1, 0, 1)
    s = _system(rods, lo, hi, pbc, mu=mu, dt=dt, KBT=kbt, rngSeed=77)
    rng = np.random.default_rng(3)
    f, v = rng.normal(size=6 * n), rng.normal(size=6 * n) * 0.1
    s.set_force_nonbrown(f)
    s.set_velocity_nonbrown(v)
    s.calc_velocity_brown()
    s.calc_velocity_noncon()
    V = s.velocities()
    W = pr.brown_normals(77, n)  # the deviates the reference's pool handed out
    vb = oracle.velocity_brown(rods["quat"], rods["radius"], rods["length"], rods["immovable"], mu, kbt, dt, W)
    assert np.abs(vb - V["velBrown"]).max() < 1e-13 * np.abs(vb).max()
    assert np.all(V["velBrown"].reshape(-1, 6)[rods["immovable"] == 1] == 0)
    orods = _orods(oracle, rods, lo, hi, pbc, 0.025)
    M = oracle.build_mobility(orods, rods["immovable"], mu)
    sy = s.sylinders()
    vnb = np.concatenate([sy["velNonB"], sy["omegaNonB"]], axis=1).reshape(-1)
    want = np.array([sum(M.data[p] * f[M.indices[p]] for p in range(M.indptr[r], M.indptr[r + 1])) for r in range(6 * n)]) + v
    assert np.array_equal(vnb, want)
    assert np.array_equal(V["velNonCon"], want + V["velBrown"])
    s.close()


# ---- committed reference outputs for the example configurations (generated by tests/golden/make_golden_solver.py)
def _example_cases(oracle):
    from oracle.pyoracle import BLOCK_DTYPE

    z = np.load(os.path.join(GOLD, "mixmotorsliding.npz"))
    rods = {k: z[k] for k in ("gid", "pos", "quat", "length", "radius", "immovable")}
    yield ("mixmotorsliding", rods, z["lo"], z["hi"], z["pbc"], float(z["colbuf"]), float(z["mu"]), float(z["dt"]),
           float(z["res"]), z["blocks"].view(BLOCK_DTYPE).copy(), 0, 10000)
    z = np.load(os.path.join(GOLD, "densemonolayer.npz"))
    n = len(z["gid"])
    rods = dict(gid=z["gid"], pos=z["pos"], quat=z["quat"], length=z["length"], radius=z["radius"],
                immovable=np.zeros(n, dtype=np.uint8))
    yield ("densemonolayer", rods, z["lo"], z["hi"], z["pbc"], float(z["colbuf"]), 1.0, 1e-5, 1e-6, None, 0, 10000)
    rng = np.random.default_rng(1234)  # Active3DNematics: 500 aligned rods
    n, box = 500, 0.7
    lo, hi = np.full(3, -0.35), np.full(3, 0.35)
    d = np.zeros((n, 3))
    d[:, 2] = np.where(rng.uniform(size=n) < 0.5, -1.0, 1.0)
    rods = dict(gid=np.arange(n, dtype=np.int32), pos=lo + rng.uniform(size=(n, 3)) * box, quat=quat_from_z_to(d),
                length=np.full(n, 0.25), radius=np.full(n, 0.0125), immovable=np.zeros(n, dtype=np.uint8))
    for choice in (0, 1):
        yield ("active3dnematics" + ("_apgd" if choice else ""), rods, lo, hi, np.array([1, 1, 1]), 0.025, 0.01, 1e-4, 1e-5,
               None, choice, 400)


def test_example_configs_against_committed_reference_outputs(oracle):
    """gamma / velocities / history of the REFERENCE's solver on the three example configurations (committed in
    tests/golden/refsolver.npz by make_golden_solver.py): the oracle reproduces them bit for bit; with the library present
    the fixture is also re-derived"""
    gold = np.load(os.path.join(GOLD, "refsolver.npz"))
    for name, rods, lo, hi, pbc, colbuf, mu, dt, res, extra, choice, max_ite in _example_cases(oracle):
        orods = _orods(oracle, rods, lo, hi, pbc, colbuf)
        coll = oracle.collect_pairs(orods, lo, hi, pbc, with_stress=True)
        blocks = coll[canonical_order(coll)]
        if extra is not None:
            blocks = np.concatenate([blocks, extra])
        vnc = np.zeros(6 * len(rods["gid"]))
        o = oracle.solve_constraints(blocks, orods, rods["immovable"], mu, vnc, dt, res, max_ite, choice)
        assert np.array_equal(o["gamma"], gold[name + "_gamma"]), name
        assert np.array_equal(o["history"], gold[name + "_history"]), name
        assert np.array_equal(o["velU"], gold[name + "_velU"]) and np.array_equal(o["velB"], gold[name + "_velB"]), name
        assert o["history"][-1][4] < res / dt, name
        if pr.available():
            s = _system(rods, lo, hi, pbc, colbuf, mu=mu, dt=dt)
            r = s.solve_blocks(blocks, vnc, dt, res, max_ite, choice)
            _assert_same_solve(r, o)
            s.close()


@pytest.mark.parametrize("pbc", [(1, 1, 1), (0, 0, 0), (1, 0, 1)])
def test_reference_two_species_search_is_the_all_images_ball_query(pbc):
    """SimToolbox/MPI/MixPairInteraction.hpp driven like MPI/MixPairInteraction_test.cpp: what FDPS hands the functor,
    filtered by distance <= max(rsTrg, rsSrc), is exactly the brute-force set over periodic images -- the statement
    tests/test_gpu_mix.py holds the device search to"""
    from mixsearch import brute_mix_pairs

    rng = np.random.default_rng(1)
    lo, hi = [0.0, 0.0, 0.0], [10.0, 8.0, 6.0]
    nt, ns = 300, 500
    trg = rng.uniform(-1, 11, size=(nt, 3)) if all(pbc) else rng.uniform(0.01, 0.99, size=(nt, 3)) * np.array(hi)
    src = rng.uniform(0, 1, size=(ns, 3)) * np.array(hi)
    trs, srs = rng.uniform(0.1, 1.5, size=nt), rng.uniform(0.1, 0.9, size=ns)
    pairs, dist = pr.mix_search(trg, trs, src, srs, lo, hi, pbc)
    want = brute_mix_pairs(trg, trs, src, srs, lo, hi, pbc)
    assert len(want) > 500 and np.array_equal(pairs, want)
    assert np.all(dist <= np.maximum(trs[pairs[:, 0]], srs[pairs[:, 1]]))
