"""Boundary collisions: Boundary::project + SylinderSystem::collectBoundaryCollision
(SimToolbox/Boundary/Boundary.cpp, Sylinder/SylinderSystem.cpp:1093-1150).

CPU part: the oracle restatement against the properties the reference's own test asserts (Boundary_test.cpp with
Boundary::check, Boundary.cpp:43-75, :126-150, :211-245): the projection lies on the surface, |query - projection| =
|delta|, delta points into the allowed region.  GPU part: the device blocks equal the oracle's bit for bit, and a
constraint solve with walls pushes the rods back inside."""
import numpy as np
import pytest

from scenarios import random_rods

EPS = np.finfo(float).eps * 1e4  # Boundary.cpp:5

SPECS = [
    dict(type="sphere", center=[2.0, 2.0, 2.0], radius=2.0, inside=True),
    dict(type="sphere", center=[2.0, 2.0, 2.0], radius=2.0, inside=False),
    dict(type="wall", center=[2.0, 2.0, 2.0], axis=[1.0, 2.0, 3.0]),
    dict(type="tube", center=[2.0, 2.0, 2.0], axis=[1.0, 2.0, 3.0], radius=2.0, inside=True),
    dict(type="tube", center=[2.0, 2.0, 2.0], axis=[1.0, 2.0, 3.0], radius=2.0, inside=False),
]


@pytest.mark.parametrize("spec", SPECS)
def test_projection_properties_of_the_reference_test(oracle, spec):
    b = oracle.make_boundaries([spec])[0]
    ctr, ax, R = np.array(spec["center"]), b["axis"], spec.get("radius", 0.0)
    rng = np.random.default_rng(3)
    for _ in range(1000):  # Boundary_test.cpp: 1000 random queries in center + [-5, 5]^3
        q = ctr + rng.uniform(-5, 5, 3)
        proj, delta = oracle.boundary_project(b, q)
        assert abs(np.linalg.norm(q - proj) - np.linalg.norm(delta)) < EPS
        dn = delta / np.linalg.norm(delta)
        if spec["type"] == "sphere":
            assert abs(np.linalg.norm(proj - ctr) - R) < EPS
            n = (proj - ctr) / np.linalg.norm(proj - ctr)
            n = -n if spec["inside"] else n
        elif spec["type"] == "wall":
            assert abs((proj - ctr) @ ax) < EPS
            n = ax
        else:
            pa = ctr + ((proj - ctr) @ ax) * ax
            assert abs(R - np.linalg.norm(proj - pa)) < EPS
            n = (pa - proj) if spec["inside"] else (proj - pa)
            n = n / np.linalg.norm(n)
        assert abs(1 - n @ dn) < EPS


def _system(oracle):
    rods = random_rods(4000, 4.0, seed=23, frac_sphere=0.15, length=0.4)
    lo, hi = [0.0] * 3, [4.0] * 3
    pos = oracle.wrap_positions(rods["pos"], lo, hi)
    orods = oracle.make_rods(rods["gid"], rods["radius"], rods["length"], pos, rods["quat"], 1.0, 1.0, 0.025)
    return rods, lo, hi, orods


def test_collect_boundary_known_answers(oracle):
    """a rod poking through the wall z = 0 (allowed side z > 0) and a sphere near a spherical shell"""
    walls = oracle.make_boundaries([dict(type="wall", center=[0.0, 0.0, 0.0], axis=[0.0, 0.0, 2.0])])
    q = np.array([[0.0, 0.0, 0.0, 1.0]])  # direction +z
    r = oracle.make_rods(np.array([7], dtype=np.int32), np.array([0.1]), np.array([1.0]), np.array([[1.0, 1.0, 0.3]]), q,
                         1.0, 1.0, 0.025)
    blk = oracle.collect_boundary(r, walls, 0.025)
    assert len(blk) == 1  # minus end at z = -0.2 is outside; plus end at z = 0.8 is far inside
    b = blk[0]
    assert b["oneSide"] == 1 and b["bilateral"] == 0 and b["gidI"] == 7 == b["gidJ"]
    np.testing.assert_allclose(b["delta0"], -0.2 - 0.1, atol=1e-15)
    np.testing.assert_allclose(b["normI"], [0, 0, 1], atol=1e-15)
    np.testing.assert_allclose(b["posI"], [0, 0, -0.5], atol=1e-15)
    np.testing.assert_allclose(b["labJ"], [1.0, 1.0, 0.0], atol=1e-15)
    # inside but within (1 + 2 colBuf) r of the wall: a block with positive separation
    r["pos"][0, 2] = 0.5 + 0.1 * 1.04
    b = oracle.collect_boundary(r, walls, 0.025)[0]
    np.testing.assert_allclose(b["delta0"], 0.1 * 1.04 - 0.1, atol=1e-15)


@pytest.mark.gpu
def test_gpu_boundary_blocks_equal_the_oracle(ctx, oracle):
    rods, lo, hi, orods = _system(oracle)
    specs = [dict(type="sphere", center=[2.0, 2.0, 2.0], radius=1.9, inside=True),
             dict(type="wall", center=[0.0, 0.0, 0.3], axis=[0.3, -0.2, 2.0]),
             dict(type="tube", center=[2.0, 2.0, 2.0], axis=[1.0, 2.0, 3.0], radius=0.6, inside=False)]
    want = oracle.collect_boundary(orods, oracle.make_boundaries(specs), 0.025)
    assert len(want) > 300
    ctx.set_domain(lo, hi, (0, 0, 0))
    ctx.set_collision_params(1.0, 1.0, 0.025)
    ctx.set_rods(rods["gid"], rods["pos"], rods["quat"], rods["length"], rods["radius"], rods["immovable"])
    nc = ctx.collect_pair_collision()
    import alens_b200.capi as capi

    raw = np.zeros(len(specs), dtype=capi.BOUNDARY_DTYPE)  # un-normalised axes: the library normalises like the reference
    for o, s in zip(raw, specs):
        o["type"] = {"sphere": 0, "wall": 1, "tube": 2}[s["type"]]
        o["inside"] = 1 if s.get("inside", True) else 0
        o["center"], o["axis"], o["radius"] = s["center"], s.get("axis", [0, 0, 1]), s.get("radius", 0.0)
    assert ctx.collect_boundary_collision(raw[:0]) == 0  # no boundary: nothing happens
    far = raw[1:2].copy()
    far["center"] = [0.0, 0.0, -50.0]  # every rod far inside the allowed side: no block
    assert ctx.collect_boundary_collision(far) == 0
    added = ctx.collect_boundary_collision(raw)
    assert added == len(want)
    got = ctx.get_constraints(with_stress=False)[nc:]
    for f in ("delta0", "gamma", "gidI", "gidJ", "globalIndexI", "globalIndexJ", "oneSide", "bilateral", "kappa", "normI",
              "normJ", "posI", "posJ", "labI", "labJ"):
        assert np.array_equal(got[f], want[f]), f  # same order (boundary, rod, end), same bits


@pytest.mark.gpu
def test_gpu_solve_with_walls_matches_oracle(ctx, oracle):
    from test_gpu_solver import relerr

    rods, lo, hi, orods = _system(oracle)
    specs = [dict(type="wall", center=[0.0, 0.0, 0.4], axis=[0.0, 0.0, 1.0]),
             dict(type="wall", center=[0.0, 0.0, 3.6], axis=[0.0, 0.0, -1.0])]
    bnd = oracle.make_boundaries(specs)
    ctx.set_domain(lo, hi, (0, 0, 0))
    ctx.set_collision_params(1.0, 1.0, 0.025)
    ctx.set_rods(rods["gid"], rods["pos"], rods["quat"], rods["length"], rods["radius"], rods["immovable"])
    ctx.collect_pair_collision()
    ctx.collect_boundary_collision(bnd)
    blocks = ctx.get_constraints(with_stress=False).copy()
    ctx.calc_mobility(1.0)
    dt, vnc = 1e-4, np.zeros(6 * len(rods["gid"]))
    rep = ctx.solve_constraints(vnc, dt, 1e-30, 30, 0)
    ref = oracle.solve_constraints(blocks, orods, rods["immovable"], 1.0, vnc, dt, 1e-30, 30, 0)
    assert rep.iterations == ref["nIte"] == 30
    assert relerr(ctx.get_gamma(), ref["gamma"]) < 1e-8
    out = ctx.get_force_velocity()
    assert relerr(out["velU"], ref["velU"]) < 1e-8
    # rods that stick out below the lower wall are pushed up
    one = blocks[blocks["oneSide"] == 1]
    low = one[(one["labJ"][:, 2] < 1.0) & (one["delta0"] < -0.05)]
    assert len(low) > 0
    idx = np.unique(low["globalIndexI"])
    assert (out["velU"].reshape(-1, 6)[idx, 2] > 0).mean() > 0.9
