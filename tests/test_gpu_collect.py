"""GPU pair collection vs the CPU oracle: bit-exact blocks (SURVEY.md 8c contract: P_gpu == P_geo)."""
import numpy as np
import pytest

from scenarios import canonical_order, random_rods

pytestmark = pytest.mark.gpu

FIELDS = ("delta0", "gamma", "gammaLB", "gidI", "gidJ", "globalIndexI", "globalIndexJ", "oneSide", "bilateral",
          "kappa", "normI", "normJ", "posI", "posJ", "labI", "labJ")


def gpu_collect(ctx, rods, lo, hi, pbc, colbuf, dratio=1.0, lratio=1.0):
    ctx.set_domain(lo, hi, pbc)
    ctx.set_collision_params(dratio, lratio, colbuf)
    ctx.set_rods(rods["gid"], rods["pos"], rods["quat"], rods["length"], rods["radius"], rods["immovable"], wrap=True)
    n = ctx.collect_pair_collision()
    blocks = ctx.get_constraints(with_stress=True)
    assert len(blocks) == n
    return blocks


def oracle_collect(oracle, rods, lo, hi, pbc, colbuf, dratio=1.0, lratio=1.0, method="cells"):
    pos = oracle.wrap_positions(rods["pos"], lo, hi, pbc)
    orods = oracle.make_rods(rods["gid"], rods["radius"], rods["length"], pos, rods["quat"], dratio, lratio, colbuf)
    return oracle.collect_pairs(orods, lo, hi, pbc, with_stress=True, method=method), orods


def assert_blocks_equal(g, o, stress_rtol=0.0):
    assert len(g) == len(o), (len(g), len(o))
    g = g[canonical_order(g)]
    o = o[canonical_order(o)]
    for f in FIELDS:
        assert np.array_equal(g[f], o[f]), f"field {f} differs"
    if stress_rtol == 0.0:
        assert np.array_equal(g["stress"], o["stress"]), "stress differs"
    else:
        np.testing.assert_allclose(g["stress"], o["stress"], rtol=stress_rtol, atol=1e-300)


@pytest.mark.parametrize("pbc", [(0, 0, 0), (1, 1, 1), (1, 0, 1)])
@pytest.mark.parametrize("n,box,L", [(3000, 2.0, 0.25), (1500, 3.0, 1.0)])
def test_collect_matches_oracle(ctx, oracle, pbc, n, box, L):
    rods = random_rods(n, box, length=L, radius=0.0125, seed=n + sum(pbc))
    lo, hi = [0, 0, 0], [box] * 3
    g = gpu_collect(ctx, rods, lo, hi, pbc, 0.025)
    o, _ = oracle_collect(oracle, rods, lo, hi, pbc, 0.025)
    assert len(o) > n // 4
    assert_blocks_equal(g, o)


def test_collect_spheres_and_mixed(ctx, oracle):
    rods = random_rods(2500, 1.2, length=0.25, radius=0.0125, seed=5, frac_sphere=0.4)
    lo, hi, pbc = [0, 0, 0], [1.2] * 3, (1, 1, 1)
    g = gpu_collect(ctx, rods, lo, hi, pbc, 0.02, dratio=1.1, lratio=0.9)
    o, _ = oracle_collect(oracle, rods, lo, hi, pbc, 0.02, dratio=1.1, lratio=0.9)
    assert len(o) > 500
    assert_blocks_equal(g, o)


def test_collect_tiny_periodic_box(ctx, oracle):
    # box barely larger than the interaction range: 2 cells per axis / 1 cell per axis with wrap
    for box, n in ((0.7, 500), (0.45, 120)):
        rods = random_rods(n, box, length=0.25, radius=0.0125, seed=11, lo=-box / 2)
        lo, hi, pbc = [-box / 2] * 3, [box / 2] * 3, (1, 1, 1)
        g = gpu_collect(ctx, rods, lo, hi, pbc, 0.025)
        o, _ = oracle_collect(oracle, rods, lo, hi, pbc, 0.025, method="brute")
        assert_blocks_equal(g, o)


def test_collect_rods_outside_box_are_wrapped(ctx, oracle):
    rods = random_rods(1000, 2.0, seed=3)
    rods["pos"] = rods["pos"] * 3.0 - 2.0  # many rods outside [0,2]^3
    lo, hi, pbc = [0, 0, 0], [2.0] * 3, (1, 1, 0)
    g = gpu_collect(ctx, rods, lo, hi, pbc, 0.025)
    o, _ = oracle_collect(oracle, rods, lo, hi, pbc, 0.025)
    np.testing.assert_array_equal(ctx.get_positions(), oracle.wrap_positions(rods["pos"], lo, hi, pbc))
    assert_blocks_equal(g, o)


def test_collect_edge_cases(ctx, oracle):
    lo, hi, pbc = [0, 0, 0], [1.0] * 3, (0, 0, 0)
    # empty
    rods = random_rods(0, 1.0)
    g = gpu_collect(ctx, rods, lo, hi, pbc, 0.025)
    assert len(g) == 0
    # single rod
    rods = random_rods(1, 1.0)
    assert len(gpu_collect(ctx, rods, lo, hi, pbc, 0.025)) == 0
    # the reference's fixed pair (SylinderNear_test.cpp:43-111): one block, known stress
    P0, P1 = np.array([1, 0, 0.0]), np.array([0, np.sqrt(3), 0.0])
    Q0, Q1 = np.array([0, 0, 1.0]), np.array([2, 2 * np.sqrt(3), 1.0])
    from scenarios import quat_from_z_to

    rods = dict(gid=np.array([0, 1], dtype=np.int32), pos=np.array([(P0 + P1) / 2, (Q0 + Q1) / 2]),
                quat=quat_from_z_to(np.array([P1 - P0, Q1 - Q0])),
                length=np.array([np.linalg.norm(P1 - P0), np.linalg.norm(Q1 - Q0)]), radius=np.array([0.4, 0.5]),
                immovable=np.zeros(2, dtype=np.uint8))
    g = gpu_collect(ctx, rods, [-5] * 3, [5] * 3, pbc, 0.5)
    assert len(g) == 1
    want = np.array([0, 0, 0.0160681, 0, 0, 0.0278307, 0.0160681, 0.0278307, 1.0])
    assert np.abs(g[0]["stress"] - want).max() < 1e-6
    assert abs(g[0]["delta0"] - 0.1) < 1e-12


def test_collect_is_deterministic(ctx):
    rods = random_rods(4000, 2.0, seed=9)
    lo, hi, pbc = [0, 0, 0], [2.0] * 3, (1, 1, 1)
    a = gpu_collect(ctx, rods, lo, hi, pbc, 0.025).copy()
    b = gpu_collect(ctx, rods, lo, hi, pbc, 0.025)
    assert a.tobytes() == b.tobytes()  # same order, same bits, run to run


@pytest.mark.parametrize("colbuf,n", [(0.025, 6000), (0.3, 2500)])
def test_split_search_gives_the_same_list_row_for_row(ctx, oracle, colbuf, n):
    """The default search (stage 1-2 kernel -> staged candidates -> dense exact query -> ordered emission through a
    per-cell bitmap) against the single-kernel search: same constraints in the SAME ORDER, bit for bit.  colbuf 0.3 makes
    cells with more candidates than the initial bitmap holds: that step falls back, the next one has grown its bitmap."""
    rods = random_rods(n, 2.2, seed=17, frac_sphere=0.1, length_sigma=0.2)
    lo, hi, pbc = [0, 0, 0], [2.2] * 3, (1, 1, 0)
    ctx.set_option("find_split", 0)
    ref = gpu_collect(ctx, rods, lo, hi, pbc, colbuf).copy()
    assert len(ref) > 5000
    ctx.set_option("find_split", 1)
    for _ in range(3):  # first pass may overflow and fall back; later passes run the split path with a wider bitmap
        got = gpu_collect(ctx, rods, lo, hi, pbc, colbuf)
        assert got.tobytes() == ref.tobytes()
    want, _ = oracle_collect(oracle, rods, lo, hi, pbc, colbuf)
    assert_blocks_equal(ref, want)


# ---- the narrow phase by itself (alens_dcp_query / alens_pair_functor) -------------------------------------------
def _segments(rng, n):
    """random, degenerate, parallel, touching and axis-aligned segment pairs (exact zeros make the clamped-root and
    0.5 fall-back branches of DCPQuery.hpp:311-472 reachable)"""
    P0 = rng.normal(size=(n, 3)); P1 = P0 + rng.normal(size=(n, 3)) * rng.choice([1.0, 1e-3], size=(n, 1))
    Q0 = rng.normal(size=(n, 3)); Q1 = Q0 + rng.normal(size=(n, 3))
    kind = np.arange(n) % 8
    Q1[kind == 1] = (Q0 + (P1 - P0))[kind == 1]                       # parallel, equal length
    Q1[kind == 2] = (Q0 - 2.5 * (P1 - P0))[kind == 2]                 # antiparallel
    Q0[kind == 3] = P0[kind == 3]                                      # shared end point
    P1[kind == 4] = P0[kind == 4]                                      # P degenerates to a point
    Q1[kind == 5] = Q0[kind == 5]; P1[kind == 5] = P0[kind == 5]      # both points
    g = np.round(rng.normal(size=(n, 12)) * 2) / 2                     # half-integer lattice: exact ties and zeros
    for a, c in ((P0, 0), (P1, 3), (Q0, 6), (Q1, 9)):
        a[kind == 6] = g[kind == 6, c:c + 3]
    Q1[kind == 7] = (Q0 + (P1 - P0) * (1 + 1e-12))[kind == 7]         # nearly parallel
    return P0, P1, Q0, Q1


def test_dcp_query_bit_exact_against_oracle_and_reference(ctx, oracle):
    from test_oracle_properties import HAVE_REF, QUIRK

    rng = np.random.default_rng(7)
    P0, P1, Q0, Q1 = _segments(rng, 6000)
    # the reference's reversal quirk (non-minimal distance), both orientations
    P0[0], P1[0], Q0[0], Q1[0] = QUIRK
    P0[1], P1[1], Q0[1], Q1[1] = QUIRK[1], QUIRK[0], QUIRK[2], QUIRK[3]
    d, P, Q = ctx.dcp_query(P0, P1, Q0, Q1)
    assert d[0] == 0.7071067811865476 and d[1] == 0.9265026063892198
    for k in range(len(d)):
        o = oracle.dcp_segseg(P0[k], P1[k], Q0[k], Q1[k])
        assert d[k] == o[0] and np.array_equal(P[k], o[1]) and np.array_equal(Q[k], o[2]), k
        if HAVE_REF and k < 2000:
            r = oracle.dcp_segseg(P0[k], P1[k], Q0[k], Q1[k], which="ref")
            assert d[k] == r[0] and np.array_equal(P[k], r[1]) and np.array_equal(Q[k], r[2]), k


def test_pair_functor_bit_exact_against_oracle(ctx, oracle):
    """CalcSylinderNearForce's body on independent pairs: spheres, rods, mixed, coincident centres (normalized() of a
    zero vector stays zero, Eigen >= 3.3), against the oracle's functor incl. the stress"""
    rng = np.random.default_rng(11)
    n = 3000
    rods = random_rods(2 * n, 1.2, seed=13, frac_sphere=0.3, length=0.5)
    rods["pos"][1::2] = rods["pos"][0::2] + rng.normal(size=(n, 3)) * 0.15  # partners close by: about half of them touch
    o = oracle.make_rods(rods["gid"], rods["radius"], rods["length"], rods["pos"], rods["quat"], 1.0, 1.0, 0.05)
    o["pos"][1] = o["pos"][0]  # coincident centres: pair 0
    if o["lengthCollision"][0] >= 2 * o["radiusCollision"][0] or o["lengthCollision"][1] >= 2 * o["radiusCollision"][1]:
        o["lengthCollision"][:2] = 0.0  # make both spheres
    geom = np.concatenate([o["pos"], o["direction"], o["lengthCollision"][:, None], o["radiusCollision"][:, None],
                           o["colBuf"][:, None]], axis=1)
    hit, blocks = ctx.pair_functor(geom[0::2], geom[1::2], with_stress=True)
    nh = 0
    for k in range(n):
        want = oracle.pair_functor(o[2 * k], o[2 * k + 1], with_stress=True)
        assert bool(hit[k]) == (want is not None), k
        if want is None:
            continue
        nh += 1
        for f in ("delta0", "gamma", "normI", "normJ", "posI", "posJ", "labI", "labJ", "stress"):
            assert np.array_equal(blocks[k][f], want[f], equal_nan=True), (k, f)
    assert hit[0] and np.all(blocks[0]["normI"] == 0)  # coincident spheres: zero normal, not NaN
    assert 500 < nh < n - 500


def test_equal_gids_never_collide(ctx, oracle):
    """the reference skips gidI >= gidJ (SylinderNear.hpp:210,225): two overlapping rods with the same gid give no block"""
    rods = random_rods(200, 0.8, seed=2)
    rods["gid"][:] = np.arange(200) // 2  # every gid twice
    lo, hi, pbc = [0.0] * 3, [0.8] * 3, (1, 1, 1)
    g = gpu_collect(ctx, rods, lo, hi, pbc, 0.05)
    o, _ = oracle_collect(oracle, rods, lo, hi, pbc, 0.05, method="brute")
    assert len(o) > 50 and np.all(o["gidI"] < o["gidJ"])
    assert_blocks_equal(g, o)


def test_set_rod_state_is_set_rods_without_the_static_fields(ctx, oracle):
    """alens_set_rod_state: new positions / orientations for the resident rod set -> the same constraint list as a full
    alens_set_rods of the moved rods"""
    rng = np.random.default_rng(3)
    rods = random_rods(3000, 2.0, seed=5, frac_sphere=0.1, frac_immovable=0.05)
    lo, hi, pbc = [0, 0, 0], [2.0] * 3, (1, 0, 1)
    moved = dict(rods)
    moved["pos"] = rods["pos"] + rng.normal(0, 0.05, size=rods["pos"].shape)
    q = rods["quat"] + rng.normal(0, 0.05, size=rods["quat"].shape)
    moved["quat"] = q / np.linalg.norm(q, axis=1)[:, None]
    want = gpu_collect(ctx, moved, lo, hi, pbc, 0.025).copy()
    gpu_collect(ctx, rods, lo, hi, pbc, 0.025)
    ctx.set_rod_state(moved["pos"], moved["quat"], wrap=True)
    n = ctx.collect_pair_collision()
    got = ctx.get_constraints(with_stress=True)
    assert n == len(want) > 3000
    assert_blocks_equal(got, want)
    import alens_b200
    c2 = alens_b200.Context(device=0)
    c2.set_domain(lo, hi, pbc)
    with pytest.raises(alens_b200.AlensError):
        c2.set_rod_state(np.zeros((0, 3)), np.zeros((0, 4)))  # no alens_set_rods yet: ALENS_ERR_STATE
    c2.close()


@pytest.mark.parametrize("pbc,sigma,n,box", [((1, 1, 1), 0.2, 9000, 2.4), ((0, 1, 0), 0.4, 7000, 2.4), ((1, 1, 1), 0.7, 6000, 2.0)])
def test_polydisperse_rods_long_rod_pass(ctx, oracle, pbc, sigma, n, box):
    """log-normal lengths: the cell grid is sized for twice the mean bounding radius, the few longer rods are paired with
    far partners by the long-rod pass -- the list must still be P_geo (all block fields bit for bit), and equal to the list
    of the plain search with cells sized for the longest rod (long_rods = 0)"""
    rods = random_rods(n, box, seed=41, length=0.1, radius=0.02, length_sigma=sigma, frac_sphere=0.05)
    rods["length"][::60] = 0.44 * box  # a few rods that cross many cells: most of their contacts are far from their centre
    rods["length"] = np.minimum(rods["length"], 0.45 * box)  # (one image per pair: shorter than half the box)
    lo, hi = [0, 0, 0], [box] * 3
    got = gpu_collect(ctx, rods, lo, hi, pbc, 0.05).copy()
    st = ctx.get_long_rod_stats()
    want, _ = oracle_collect(oracle, rods, lo, hi, pbc, 0.05)
    assert_blocks_equal(got, want)
    assert st["long_rods"] > 10 and st["short_radius"] < st["max_radius"]
    assert st["long_rows"] > (20 if sigma < 0.5 else 0), st  # contacts the 27-cell stencil cannot see
    ctx.set_option("long_rods", 0)
    plain = gpu_collect(ctx, rods, lo, hi, pbc, 0.05).copy()
    assert ctx.get_long_rod_stats()["long_rods"] == 0
    ctx.set_option("long_rods", 200)
    assert_blocks_equal(got, plain)
    # and the solve runs on the extended list
    gpu_collect(ctx, rods, lo, hi, pbc, 0.05)
    ctx.calc_mobility(1.0)
    rep = ctx.solve_constraints(np.zeros(6 * n), 1e-4, 1e-5, 50, 0)
    assert rep.iterations > 0 and np.isfinite(rep.residual)
