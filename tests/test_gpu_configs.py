"""The reference's own example configurations (BASELINE.json configs 0-2) through the C ABI against the CPU oracle.

* MixMotorSliding as shipped: 2 rods of length 20 (one immovable) + the 97 doubly bound motors of ProteinInitial.dat as
  bilateral blocks (fixture tests/golden/mixmotorsliding.npz, generated from /root/reference by make_golden.py).
* DenseMonoLayer: the 9700 rods of TubuleInitial.dat (fixture densemonolayer.npz, which also holds the pair list the
  reference's own FDPS search produced and the geometric list).
* Active3DNematics: no rod file is shipped; 500 rods L = 0.25 in a periodic 0.7^3 box with directions exactly +-z
  (`initOrient [0,0,2]`, SylinderSystem.cpp:190-216) from a seeded generator: every pair is parallel, the degenerate
  branch of the closest-point query (DCPQuery.hpp:325-327,361-363,434-438).
Tolerances as in test_gpu_solver.py: pair lists bit for bit, gamma / velocities 1e-8 relative at equal iteration count.
"""
import os

import numpy as np
import pytest

from scenarios import canonical_order, quat_from_z_to
from test_gpu_solver import relerr

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LIST_FIELDS = ("gidI", "gidJ", "delta0", "normI", "posI", "posJ", "labI", "labJ")


def _load(ctx, rods, lo, hi, pbc, colbuf):
    ctx.set_domain(lo, hi, pbc)
    ctx.set_collision_params(1.0, 1.0, colbuf)
    ctx.set_rods(rods["gid"], rods["pos"], rods["quat"], rods["length"], rods["radius"], rods["immovable"])


def _oracle_rods(oracle, rods, lo, hi, colbuf, pbc):
    return oracle.make_rods(rods["gid"], rods["radius"], rods["length"], oracle.wrap_positions(rods["pos"], lo, hi, pbc),
                            rods["quat"], 1.0, 1.0, colbuf)


def _compare_lists(got, want):
    assert len(got) == len(want)
    g, w = got[canonical_order(got)], want[canonical_order(want)]
    for f in LIST_FIELDS:
        assert np.array_equal(g[f], w[f]), f


def test_mixmotorsliding_as_shipped(ctx, oracle):
    from alens_b200.capi import BLOCK_DTYPE

    z = np.load(os.path.join(GOLD, "mixmotorsliding.npz"))
    rods = {k: z[k] for k in ("gid", "pos", "quat", "length", "radius", "immovable")}
    lo, hi, pbc, colbuf = z["lo"], z["hi"], z["pbc"], float(z["colbuf"])
    dt, res, mu = float(z["dt"]), float(z["res"]), float(z["mu"])
    motors = z["blocks"].view(BLOCK_DTYPE).copy()
    assert len(motors) == 97 and rods["immovable"].tolist() == [1, 0] and np.allclose(rods["length"], 20.0)
    _load(ctx, rods, lo, hi, pbc, colbuf)
    # the two rods are 0.07 apart centre to centre: 0.045 between the surfaces, outside colBuf -> no collision block
    assert ctx.collect_pair_collision() == 0
    orods = _oracle_rods(oracle, rods, lo, hi, colbuf, pbc)
    assert len(oracle.collect_pairs(orods, lo, hi, pbc)) == 0
    ctx.append_constraints(motors)
    ctx.calc_mobility(mu)
    vnc = np.zeros(12)
    for max_ite in (15, 10000):
        rep = ctx.solve_constraints(vnc, dt, res, max_ite, 0)
        ref = oracle.solve_constraints(motors, orods, rods["immovable"], mu, vnc, dt, res, max_ite, 0)
        assert rep.iterations == ref["nIte"]
        out = ctx.get_force_velocity()
        assert relerr(ctx.get_gamma(), ref["gamma"]) < 1e-8
        for k in ("forceB", "velB"):
            assert relerr(out[k], ref[k]) < 1e-8, k
        assert np.all(out["velB"][:6] == 0) and np.abs(out["velB"][6:]).max() > 0  # rod 0 is immovable ('S')
        assert np.all(out["velU"] == 0) and np.all(out["forceU"] == 0)          # no unilateral row
    assert rep.status == 0 and rep.residual < res / dt
    # the motors are stretched beyond their rest length: they pull the free rod along -x / towards the fixed one
    assert ref["nIte"] > 0


def test_densemonolayer_initial_state(ctx, oracle):
    z = np.load(os.path.join(GOLD, "densemonolayer.npz"))
    n = len(z["gid"])
    rods = dict(gid=z["gid"], pos=z["pos"], quat=z["quat"], length=z["length"], radius=z["radius"],
                immovable=np.zeros(n, dtype=np.uint8))
    lo, hi, pbc, colbuf = z["lo"], z["hi"], z["pbc"], float(z["colbuf"])
    _load(ctx, rods, lo, hi, pbc, colbuf)
    nc = ctx.collect_pair_collision()
    blocks = ctx.get_constraints(with_stress=True).copy()
    # the geometric list stored with the fixture (oracle at generation time) and the reference's FDPS list
    order = canonical_order(blocks)
    key = np.stack([blocks["gidI"][order], blocks["gidJ"][order]], axis=1)
    geo = np.stack([z["geo_gidI"], z["geo_gidJ"]], axis=1)
    go = np.lexsort((geo[:, 1], geo[:, 0]))
    assert nc == len(geo) == 20567
    ko = np.lexsort((key[:, 1], key[:, 0]))
    assert np.array_equal(key[ko], geo[go])
    assert np.array_equal(blocks["delta0"][order][ko], z["geo_delta0"][go])
    have = set(map(tuple, key.tolist()))
    ref_pairs = list(zip(z["ref_gidI"].tolist(), z["ref_gidJ"].tolist()))
    assert len(ref_pairs) == 19120 and all(p in have for p in ref_pairs)  # P_ref(FDPS) is a subset of P_gpu
    d0 = dict(zip(map(tuple, key.tolist()), blocks["delta0"][order].tolist()))
    assert all(d0[p] == v for p, v in zip(ref_pairs, z["ref_delta0"].tolist()))
    # and against the oracle of today, every field
    orods = _oracle_rods(oracle, rods, lo, hi, colbuf, pbc)
    _compare_lists(blocks, oracle.collect_pairs(orods, lo, hi, pbc, with_stress=True))
    # one constraint solve with the example's parameters (dt 1e-5, conResTol 1e-6, mu 1)
    ctx.calc_mobility(1.0)
    vnc = np.zeros(6 * n)
    rep = ctx.solve_constraints(vnc, 1e-5, 1e-30, 30, 0)
    ref = oracle.solve_constraints(blocks, orods, rods["immovable"], 1.0, vnc, 1e-5, 1e-30, 30, 0)
    assert rep.iterations == ref["nIte"] == 30
    assert relerr(ctx.get_gamma(), ref["gamma"]) < 1e-8
    assert relerr(ctx.get_force_velocity()["velU"], ref["velU"]) < 1e-8
    rep = ctx.solve_constraints(vnc, 1e-5, 1e-6, 10000, 0)
    assert rep.status == 0 and rep.residual < 1e-6 / 1e-5


def test_active3dnematics_aligned(ctx, oracle):
    rng = np.random.default_rng(1234)
    n, box, colbuf, mu, dt, res = 500, 0.7, 0.025, 0.01, 1e-4, 1e-5
    lo, hi, pbc = np.full(3, -0.35), np.full(3, 0.35), np.array([1, 1, 1], dtype=np.int32)
    sign = np.where(rng.uniform(size=n) < 0.5, -1.0, 1.0)
    d = np.zeros((n, 3))
    d[:, 2] = sign
    rods = dict(gid=np.arange(n, dtype=np.int32), pos=lo + rng.uniform(size=(n, 3)) * box, quat=quat_from_z_to(d),
                length=np.full(n, 0.25), radius=np.full(n, 0.0125), immovable=np.zeros(n, dtype=np.uint8))
    _load(ctx, rods, lo, hi, pbc, colbuf)
    nc = ctx.collect_pair_collision()
    blocks = ctx.get_constraints(with_stress=True).copy()
    orods = _oracle_rods(oracle, rods, lo, hi, colbuf, pbc)
    want = oracle.collect_pairs(orods, lo, hi, pbc, with_stress=True)
    assert nc == len(want) > 100
    _compare_lists(blocks, want)
    g, w = blocks[canonical_order(blocks)], want[canonical_order(want)]
    assert np.array_equal(g["stress"], w["stress"])
    # parallel axes: the contact normal has no z component unless the rods touch end to end
    side = np.abs(g["normI"][:, 2]) < 1e-12
    assert side.sum() > 0.5 * nc
    ctx.calc_mobility(mu)
    vnc = np.zeros(6 * n)
    rep = ctx.solve_constraints(vnc, dt, 1e-30, 25, 0)
    ref = oracle.solve_constraints(blocks, orods, rods["immovable"], mu, vnc, dt, 1e-30, 25, 0)
    assert rep.iterations == ref["nIte"] == 25
    assert relerr(ctx.get_gamma(), ref["gamma"]) < 1e-8
    assert relerr(ctx.get_force_velocity()["velU"], ref["velU"]) < 1e-8
    rep = ctx.solve_constraints(vnc, dt, res, 10000, 0)
    assert rep.status == 0 and rep.residual < res / dt
