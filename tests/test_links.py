"""Bilateral links: SylinderSystem::collectLinkBilateral (SimToolbox/Sylinder/SylinderSystem.cpp:1386-1482).
CPU: known answers of the oracle restatement.  GPU: alens_collect_link_bilateral equals the oracle bit for bit (all block
fields and the stress), across a periodic face, and the constraint solve with the spring blocks matches the oracle."""
import numpy as np
import pytest

from scenarios import quat_from_z_to

FIELDS = ("delta0", "gamma", "gidI", "gidJ", "globalIndexI", "globalIndexJ", "oneSide", "bilateral", "kappa", "normI",
          "normJ", "posI", "posJ", "labI", "labJ", "stress")


def filaments(n_fil=40, n_seg=12, box=3.0, seg=0.2, radius=0.0125, gap=0.01, seed=3):
    """random worm-like chains of n_seg rods; consecutive rods linked plus end -> minus end"""
    rng = np.random.default_rng(seed)
    pos, dirs, prev, nxt = [], [], [], []
    gid = 0
    for _ in range(n_fil):
        p = rng.uniform(0, box, 3)
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        for s in range(n_seg):
            d = d + 0.25 * rng.normal(size=3)
            d /= np.linalg.norm(d)
            c = p + 0.5 * seg * d
            pos.append(c)
            dirs.append(d)
            if s > 0:
                prev.append(gid - 1)
                nxt.append(gid)
            p = p + (seg + 2 * radius + gap * rng.uniform(0.5, 1.5)) * d
            gid += 1
    n = len(pos)
    perm = rng.permutation(n)  # storage order is unrelated to the gid order
    rods = dict(gid=np.arange(n, dtype=np.int32)[perm], pos=np.array(pos)[perm], quat=quat_from_z_to(np.array(dirs))[perm],
                length=np.full(n, seg), radius=np.full(n, radius), immovable=np.zeros(n, dtype=np.uint8))
    return rods, np.array(prev, dtype=np.int32), np.array(nxt, dtype=np.int32)


def test_link_known_answer(oracle):
    # two collinear rods along x, plus end of rod 0 at x = 1, minus end of rod 1 at x = 1.3
    q = quat_from_z_to(np.array([[1.0, 0.0, 0.0], [1.0, 0.0, 0.0]]))
    rods = oracle.make_rods(np.array([5, 9], dtype=np.int32), np.array([0.1, 0.1]), np.array([1.0, 1.0]),
                            np.array([[0.5, 2.0, 2.0], [1.8, 2.0, 2.0]]), q, 1.0, 1.0, 0.025)
    lo, hi, pbc = [0.0] * 3, [4.0] * 3, [0, 0, 0]
    b = oracle.collect_links(rods, [5], [9], lo, hi, pbc, 100.0, 0.05)[0]
    np.testing.assert_allclose(b["delta0"], 0.3 - 0.1 - 0.1 - 0.05, atol=1e-14)
    assert b["gamma"] == 0 and b["bilateral"] == 1 and b["oneSide"] == 0 and b["kappa"] == 100.0
    np.testing.assert_allclose(b["normI"], [-1, 0, 0], atol=1e-14)
    np.testing.assert_allclose(b["posI"], [0.5, 0, 0], atol=1e-14)
    np.testing.assert_allclose(b["posJ"], [-0.5, 0, 0], atol=1e-14)
    # across the periodic face: rod 1 sits at x = 0.1 of a box of length 2 and is seen at x = 2.1 by rod 0 at x = 1.4
    rods["pos"][0] = [1.4, 1.0, 1.0]
    rods["pos"][1] = [0.1, 1.0, 1.0]
    b = oracle.collect_links(rods, [5], [9], [0.0] * 3, [2.0] * 3, [1, 0, 0], 100.0, 0.0)[0]
    np.testing.assert_allclose(b["labJ"], [1.6, 1.0, 1.0], atol=1e-14)   # minus end of the image at 2.1 - 0.5
    np.testing.assert_allclose(b["delta0"], 0.3 - 0.2, atol=1e-14)       # |Q - P| = 0.3 (the ends have passed each other)
    np.testing.assert_allclose(b["normI"], [1, 0, 0], atol=1e-14)        # (P - Q)/|P - Q| with P = 1.9, Q = 1.6
    assert b["gamma"] == 0
    with pytest.raises(ValueError):
        oracle.collect_links(rods, [5], [77], lo, hi, pbc, 1.0, 0.0)


@pytest.mark.gpu
def test_gpu_links_equal_the_oracle_and_solve(ctx, oracle):
    import alens_b200
    from test_gpu_solver import relerr

    rods, prev, nxt = filaments()
    lo, hi, pbc = [0.0] * 3, [3.0] * 3, (1, 1, 1)
    kappa, gap = 100.0, 0.01
    pos = oracle.wrap_positions(rods["pos"], lo, hi, pbc)
    orods = oracle.make_rods(rods["gid"], rods["radius"], rods["length"], pos, rods["quat"], 1.0, 1.0, 0.025)
    want = oracle.collect_links(orods, prev, nxt, lo, hi, pbc, kappa, gap)
    assert len(want) == len(prev) == 40 * 11
    ctx.set_domain(lo, hi, pbc)
    ctx.set_collision_params(1.0, 1.0, 0.025)
    ctx.set_rods(rods["gid"], rods["pos"], rods["quat"], rods["length"], rods["radius"], rods["immovable"])
    nc = ctx.collect_pair_collision()
    assert ctx.collect_link_bilateral(prev, nxt, kappa, gap) == len(want)
    blocks = ctx.get_constraints(with_stress=True).copy()
    got = blocks[nc:]
    for f in FIELDS:
        assert np.array_equal(got[f], want[f]), f  # link order, every bit
    wrapped = (np.abs(want["labJ"] - want["labI"]).max(axis=1) < 0.5).all() and \
              (np.abs(pos[want["globalIndexJ"]] - pos[want["globalIndexI"]]).max(axis=1) > 1.5).any()
    assert wrapped, "no link crosses a periodic face in this system"
    with pytest.raises(alens_b200.capi.AlensError):
        ctx.collect_link_bilateral([0], [10 ** 6], kappa, gap)  # unknown gid
    # solve: collisions + springs
    ctx.calc_mobility(1.0)
    dt, vnc = 1e-4, np.zeros(6 * len(rods["gid"]))
    rep = ctx.solve_constraints(vnc, dt, 1e-30, 25, 0)
    ref = oracle.solve_constraints(blocks, orods, rods["immovable"], 1.0, vnc, dt, 1e-30, 25, 0)
    assert rep.iterations == ref["nIte"] == 25
    assert relerr(ctx.get_gamma(), ref["gamma"]) < 1e-8
    out = ctx.get_force_velocity()
    assert np.abs(ref["velB"]).max() > 0
    for k in ("velU", "velB", "forceB"):
        assert relerr(out[k], ref[k]) < 1e-7 or np.abs(ref[k]).max() == 0, k
