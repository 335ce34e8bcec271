"""The CUDA path against the REFERENCE'S OWN CODE (oracle/_ref/libalens_refsys.so: SylinderSystem.cpp, SylinderNear.hpp,
Constraint/*.cpp, Boundary.cpp, Sylinder.cpp compiled unmodified on the stand-in headers of oracle/stubs; the library is
prebuilt and travels to the GPU box).  No restatement in between:

  * pair list: every block the reference's FDPS search + functor finds is in the GPU list with bit-identical fields
    (incl. the stress); the GPU list is the geometric superset (SURVEY 8c contract)
  * boundary and link blocks: identical lists, bit for bit
  * solve: both solvers fed the SAME list -> same iteration count, gamma / forces / velocities / history to 1e-8
    (BASELINE.json north_star tolerance, fp64), through BBPGD and APGD
  * whole time steps (prepareStep -> velocities -> resolveConstraints -> stepEuler), the device solving the list the
    reference's search found at every step: positions and orientations after every Euler step
"""
import numpy as np
import pytest

from scenarios import canonical_order, random_rods, thermal_velocity
from test_reference_pin import BLOCK_FIELDS, _system

from oracle import pyrefsys as pr

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not pr.available(), reason="oracle/_ref/libalens_refsys.so missing (built where /root/reference exists)")]

TOL = 1e-8  # north_star: gamma and rod velocities within 1e-8 relative in fp64


def relerr(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)


def _gpu_load(ctx, rods, lo, hi, pbc, colbuf, dratio=1.0, lratio=1.0):
    ctx.set_domain(lo, hi, pbc)
    ctx.set_collision_params(dratio, lratio, colbuf)
    ctx.set_rods(rods["gid"], rods["pos"], rods["quat"], rods["length"], rods["radius"], rods["immovable"], wrap=True)


def _key(b):
    return list(zip(b["gidI"].tolist(), b["gidJ"].tolist(), map(bytes, np.ascontiguousarray(b["labJ"]))))


@pytest.mark.parametrize("pbc,frac_sphere", [((1, 1, 1), 0.0), ((0, 1, 0), 0.3)])
def test_pair_list_contains_the_reference_list_bit_for_bit(ctx, pbc, frac_sphere):
    n, box, colbuf = 4000, 2.0, 0.025
    rods = random_rods(n, box, seed=60 + sum(pbc), frac_sphere=frac_sphere)
    rods["pos"] = rods["pos"] * 1.2 - 0.1
    lo, hi = [0.0] * 3, [box] * 3
    s = _system(rods, lo, hi, pbc, colbuf, nthreads=2, sylinderDiameterColRatio=1.05, sylinderLengthColRatio=0.97)
    nref = s.collect_pair_collision()
    ref = s.constraints()
    _gpu_load(ctx, rods, lo, hi, pbc, colbuf, 1.05, 0.97)
    assert np.array_equal(ctx.get_positions(), s.sylinders()["pos"])  # applyBoxBC
    ngpu = ctx.collect_pair_collision()
    gpu = ctx.get_constraints(with_stress=True)
    assert 0 < nref <= ngpu and ngpu - nref < 0.1 * ngpu
    kg = {k: i for i, k in enumerate(_key(gpu))}
    assert len(kg) == ngpu
    sel = gpu[[kg[k] for k in _key(ref)]]  # KeyError = a reference pair the GPU missed
    for f in BLOCK_FIELDS:
        assert np.array_equal(sel[f], ref[f]), f
    s.close()


def test_boundary_and_link_collectors_equal_the_reference(ctx):
    from alens_b200.capi import BOUNDARY_DTYPE

    n, box, colbuf = 1500, 3.0, 0.025
    rods = random_rods(n, box, seed=23, frac_sphere=0.15, length=0.4)
    lo, hi, pbc = [0.0] * 3, [box] * 3, (1, 1, 0)
    bnd = [dict(type="wall", center=[0.0, 0.0, 0.3], norm=[0.0, 0.0, 1.0]),
           dict(type="sphere", center=[1.5, 1.5, 1.5], radius=1.4, inside=True),
           dict(type="tube", center=[1.5, 1.5, 0.0], axis=[0.0, 0.1, 1.0], radius=1.3, inside=False)]
    s = _system(rods, lo, hi, pbc, colbuf, boundaries=bnd, linkKappa=250.0, linkGap=0.02)
    prev = np.arange(0, 600, 2)
    s.add_links(rods["gid"][prev], rods["gid"][prev + 1])
    s.prepare_step()
    s.collect_boundary_collision()
    s.collect_link_bilateral()
    ref = s.constraints()
    gb = np.zeros(3, dtype=BOUNDARY_DTYPE)
    for o, b in zip(gb, bnd):
        o["type"] = {"sphere": 0, "wall": 1, "tube": 2}[b["type"]]
        o["inside"] = 1 if b.get("inside", True) else 0
        o["center"] = b["center"]
        o["axis"] = b.get("norm", b.get("axis", [0, 0, 1.0]))
        o["radius"] = b.get("radius", 0.0)
    _gpu_load(ctx, rods, lo, hi, pbc, colbuf)
    ctx.collect_pair_collision()
    ncoll = ctx.num_constraints()
    ctx.collect_boundary_collision(gb)
    ctx.collect_link_bilateral(rods["gid"][prev], rods["gid"][prev + 1], 250.0, 0.02)
    gpu = ctx.get_constraints(with_stress=True)[ncoll:]
    assert len(gpu) == len(ref) and ref["bilateral"].sum() == len(prev) and ref["oneSide"].sum() > 500

    def order(b):
        return np.lexsort((b["labJ"][:, 2], b["labJ"][:, 1], b["labJ"][:, 0], b["labI"][:, 0], b["gidJ"], b["gidI"], b["bilateral"]))

    g, r = gpu[order(gpu)], ref[order(ref)]
    for f in BLOCK_FIELDS:
        assert np.array_equal(g[f], r[f]), f
    s.close()


@pytest.mark.parametrize("choice,max_ite,res", [(0, 25, 1e-30), (0, 5000, 1e-6), (1, 12, 1e-30), (1, 5000, 1e-6)])
def test_solve_on_the_same_list_matches_the_reference_solver(ctx, choice, max_ite, res):
    """ConstraintSolver + BCQPSolver of the reference and the device solver on the GPU's own (geometric) list + link and
    wall blocks, immovable rods included"""
    n, box, colbuf, mu, dt = 3000, 1.9, 0.025, 1.0, 1e-4
    # (overlapping immovable rods and links stretched across the box make the problem infeasible: fixed-count runs only)
    rods = random_rods(n, box, seed=12, frac_sphere=0.1, frac_immovable=0.03 if max_ite < 1000 else 0.0)
    lo, hi, pbc = [0.0] * 3, [box] * 3, (1, 1, 1)
    s = _system(rods, lo, hi, pbc, colbuf, mu=mu, dt=dt, linkKappa=300.0, linkGap=0.01)
    _gpu_load(ctx, rods, lo, hi, pbc, colbuf)
    ctx.collect_pair_collision()
    prev = np.arange(0, 300, 2)
    if max_ite < 1000:  # (links between random rods are stretched across the box: that problem does not converge)
        ctx.collect_link_bilateral(rods["gid"][prev], rods["gid"][prev + 1], 300.0, 0.01)
    blocks = ctx.get_constraints(with_stress=True).copy()
    assert blocks["bilateral"].sum() == (len(prev) if max_ite < 1000 else 0) and len(blocks) > 5000
    vnc = thermal_velocity(rods, mu, dt, seed=2)
    ctx.calc_mobility(mu)
    rep = ctx.solve_constraints(vnc, dt, res, max_ite, choice)
    r = s.solve_blocks(blocks, vnc, dt, res, max_ite, choice, hist_cap=200000)
    hist = ctx.get_history()
    if max_ite < 1000:  # equal iteration count: iterates agree to rounding
        assert rep.iterations == r["nIte"] == max_ite
        assert relerr(ctx.get_gamma(), r["gamma"]) < TOL
        assert hist.shape == r["history"].shape
        np.testing.assert_allclose(hist[:, 3:5], r["history"][:, 3:5], rtol=1e-7)
        assert np.array_equal(hist[:, 5], r["history"][:, 5])  # mvCount
    else:  # converged: same residual bound; the velocities are unique, gamma is not (D^T M D is only PSD)
        assert rep.status == 0 and rep.residual < res / dt and r["history"][-1][4] < res / dt
        assert abs(rep.iterations - r["nIte"]) <= max(3, 0.05 * r["nIte"])
    out = ctx.get_force_velocity()
    vtol = TOL if max_ite < 1000 else 50 * res / dt / np.abs(r["velU"]).max()
    for k in ("velU", "velB", "forceU", "forceB"):
        scale = max(np.abs(r["velU" if k[0] == "v" else "forceU"]).max(), 1e-300)
        assert np.abs(out[k] - r[k]).max() < max(vtol, TOL) * scale * (1 if k[0] == "v" else 50), k
    s.close()


def test_time_steps_follow_the_reference_system(ctx):
    """4 steps of the reference's own loop (prepareStep, calcVelocityNonCon with a force, resolveConstraints = FDPS
    collection + ConstraintSolver, sumForceVelocity, stepEuler: SylinderSystem_main.cpp / runStep) against the resident
    device loop.  The reference's search misses some geometric contacts, so every step the device solves the
    list the reference found (checked to be a bit-identical subset of the device's own list) with the same fixed number
    of BBPGD iterations; positions and orientations must then agree to rounding after every Euler step."""
    n, box, colbuf, mu, dt, ite, steps = 3000, 1.7, 0.025, 1.0, 1e-4, 40, 4
    rods = random_rods(n, box, seed=8, frac_sphere=0.2, frac_immovable=0.02)
    lo, hi, pbc = [0.0] * 3, [box] * 3, (1, 1, 0)
    force = np.random.default_rng(2).normal(size=(n, 6)) * 0.05
    force[:, 3:] *= 0.01
    s = _system(rods, lo, hi, pbc, colbuf, mu=mu, dt=dt, conResTol=1e-30, conMaxIte=ite, nthreads=2)
    _gpu_load(ctx, rods, lo, hi, pbc, colbuf)
    path = 0.0  # sum over the steps of dt * max |velocity|: the scale position differences are measured against
    for k in range(steps):
        if k > 0:
            s.prepare_step()
            ctx.prepare_step(True)
        assert np.abs(ctx.get_positions() - s.sylinders()["pos"]).max() <= 1e-6 * path  # incl. the wrap across the faces
        s.collect_pair_collision()
        ref = s.constraints().copy()  # before the solve: gamma = initial guess, unit stress
        s.clear_constraints()
        ngpu = ctx.collect_pair_collision()
        gpu = ctx.get_constraints(with_stress=False)
        kg = {kk: i for i, kk in enumerate(_key(gpu))}
        if k == 0:  # same inputs bit for bit: the reference's list is a sub-list of the device's, fields identical
            sel = gpu[[kg[kk] for kk in _key(ref)]]
            for f in ("delta0", "gamma", "normI", "posI", "posJ", "labI", "labJ"):
                assert np.array_equal(sel[f], ref[f]), f
        assert len(ref) <= ngpu and len(ref) > 1000
        s.set_force_nonbrown(force.reshape(-1))
        s.calc_velocity_noncon()
        s.resolve_constraints()  # collects the same list again, solves, writes back
        s.sum_force_velocity()
        s.step_euler()
        ctx.clear_constraints()
        ctx.append_constraints(ref)
        ctx.calc_mobility(mu)
        vnb = ctx.calc_velocity_noncon(force_nonbrown=force.reshape(-1))
        sy = s.sylinders()
        assert relerr(vnb.reshape(-1, 6), np.concatenate([sy["velNonB"], sy["omegaNonB"]], axis=1)) < 1e-13
        rep = ctx.solve_constraints(None, dt, 1e-30, ite, 0)
        assert rep.iterations == ite
        out, want = ctx.get_force_velocity(), s.force_velocity()
        assert relerr(out["velU"], want["velU"]) < TOL and relerr(out["forceU"], want["forceU"]) < TOL
        ctx.step_euler(dt)
        pos, quat = ctx.get_rod_state()
        sy = s.sylinders()
        path += dt * np.abs(sy["vel"]).max()
        # a velocity difference of 1e-8 relative moves a rod by 1e-8 * dt * |v|; later steps start from slightly different
        # positions, and the overlapping start is stiff: 1e-6 of the path travelled is the bar
        assert np.abs(pos - sy["pos"]).max() < 1e-6 * path
        assert np.abs(quat - sy["orientation"]).max() < 1e-6
    assert path > 1e-2
    imm = rods["immovable"] == 1
    assert imm.sum() > 10 and np.array_equal(pos[imm], ctx_wrap(rods["pos"][imm], lo, hi, pbc))
    s.close()


def ctx_wrap(p, lo, hi, pbc):
    p = p.copy()
    for k in range(3):
        if pbc[k]:
            p[:, k] = lo[k] + np.mod(p[:, k] - lo[k], hi[k] - lo[k])
    return p


def test_brownian_velocity_with_the_reference_deviates(ctx):
    n, box, mu, dt, kbt = 1500, 1.5, 0.7, 1e-4, 0.00411
    rods = random_rods(n, box, seed=4, frac_sphere=0.2, frac_immovable=0.1)
    lo, hi, pbc = [0.0] * 3, [box] * 3, (1, 0, 1)
    s = _system(rods, lo, hi, pbc, mu=mu, dt=dt, KBT=kbt, rngSeed=5)
    s.calc_velocity_brown()
    s.calc_velocity_noncon()
    want = s.velocities()
    _gpu_load(ctx, rods, lo, hi, pbc, 0.025)
    ctx.calc_mobility(mu)
    vb = ctx.calc_velocity_brown(kbt, dt, normals12=pr.brown_normals(5, n))
    assert relerr(vb, want["velBrown"]) < 1e-11
    ctx.calc_velocity_noncon(vel_brown=vb)
    s.close()


def test_euler_step_equals_sylinder_step_euler(ctx):
    """Sylinder::stepEuler (Sylinder.cpp:91-99) with EquatnHelper::rotateEquatn per rod against k_step_euler: the rotated
    quaternion itself, component by component"""
    n, box, mu, dt = 400, 2.0, 1.0, 1e-3
    rods = random_rods(n, box, seed=6)
    lo, hi, pbc = [0.0] * 3, [box] * 3, (0, 0, 0)
    _gpu_load(ctx, rods, lo, hi, pbc, 0.025)
    ctx.collect_pair_collision()
    ctx.calc_mobility(mu)
    v = np.random.default_rng(3).normal(size=(n, 6))
    v[:5, 3:] = 0.0           # |omega| = 0: rotateEquatn returns early
    v[5:10, 3:] *= 1e-9       # below float epsilon: also untouched
    ctx.solve_constraints(v.reshape(-1), dt, 1e-6, 0, 0)  # maxIte 0: velUni/velBi of the initial guess
    out = ctx.get_force_velocity()
    total = v + (out["velU"] + out["velB"]).reshape(-1, 6)  # sumForceVelocity
    ctx.step_euler(dt)
    pos, quat = ctx.get_rod_state()
    p0 = ctx_pos0 = rods["pos"]
    for i in range(n):
        p, q = pr.sylinder_step_euler(p0[i], rods["quat"][i], total[i, :3], total[i, 3:], dt)
        assert np.abs(pos[i] - p).max() < 1e-14 and np.abs(quat[i] - q).max() < 1e-14, i


def test_protein_bilateral_blocks(ctx, oracle):
    """TubuleSystem::setProteinConstraints (SRC/TubuleSystem.cpp:694-745) on the device: block fields are the function's
    own expressions (restated here line by line), the stress is the reference's CalcSylinderNearForce::collideStress"""
    from alens_b200.capi import PROTEIN_DTYPE

    rng = np.random.default_rng(21)
    n, box, D = 600, 3.0, 0.025
    rods = random_rods(n, box, seed=2, length=0.6)
    lo, hi, pbc = [0.0] * 3, [box] * 3, (0, 0, 0)
    _gpu_load(ctx, rods, lo, hi, pbc, 0.025)
    ctx.collect_pair_collision()
    ncoll = ctx.num_constraints()
    m = 400
    pr_ = np.zeros(m, dtype=PROTEIN_DTYPE)
    iI, iJ = rng.integers(0, n, m), rng.integers(0, n, m)
    iJ = np.where(iJ == iI, (iJ + 1) % n, iJ)
    dirs = np.array([oracle.quat_to_dir(q) for q in rods["quat"]])
    for e, idx in enumerate((iI, iJ)):
        pr_["idBind"][:, e] = rods["gid"][idx]
        pr_["indexBind"][:, e] = idx
        pr_["centerBind"][:, e] = rods["pos"][idx]
        pr_["directionBind"][:, e] = dirs[idx]
        pr_["lenBind"][:, e] = rods["length"][idx]
        pr_["posEndBind"][:, e] = rods["pos"][idx] + dirs[idx] * (rng.uniform(-0.5, 0.5, m) * rods["length"][idx])[:, None]
    ell = np.linalg.norm(pr_["posEndBind"][:, 0] - pr_["posEndBind"][:, 1], axis=1)
    pr_["forceLength"] = ell - D  # ProteinData::getProteinForceLength, lookupType 0
    pr_["freeLength"] = 0.05
    pr_["kappa"] = rng.uniform(50.0, 200.0, m)
    single = rng.uniform(size=m) < 0.25
    pr_["idBind"][single, 1] = -1  # ID_UB: singly bound, not a constraint
    added = ctx.collect_protein_bilateral(pr_, D)
    assert added == (~single).sum()
    got = ctx.get_constraints(with_stress=True)[ncoll:]
    keep = pr_[~single]
    P, Q = keep["posEndBind"][:, 0], keep["posEndBind"][:, 1]
    d0 = keep["forceLength"] - keep["freeLength"]
    assert np.array_equal(got["delta0"], d0) and np.array_equal(got["gamma"], -d0 * keep["kappa"])
    pq = P - Q
    nrm = np.sqrt((pq[:, 0] * pq[:, 0] + pq[:, 1] * pq[:, 1]) + pq[:, 2] * pq[:, 2])
    assert np.array_equal(got["normI"], pq / nrm[:, None]) and np.array_equal(got["normJ"], -got["normI"])
    assert np.array_equal(got["posI"], P - keep["centerBind"][:, 0]) and np.array_equal(got["posJ"], Q - keep["centerBind"][:, 1])
    assert np.array_equal(got["labI"], P) and np.array_equal(got["labJ"], Q)
    assert np.all(got["bilateral"] == 1) and np.all(got["oneSide"] == 0) and np.array_equal(got["kappa"], keep["kappa"])
    assert np.array_equal(got["gidI"], keep["idBind"][:, 0]) and np.array_equal(got["globalIndexJ"], keep["indexBind"][:, 1])
    for k in range(0, len(keep), 7):
        p = keep[k]
        want = pr.collide_stress(p["directionBind"][0], p["directionBind"][1], p["centerBind"][0], p["centerBind"][1],
                                 p["lenBind"][0], p["lenBind"][1], D / 2, D / 2, 1.0, P[k], Q[k])
        assert np.array_equal(got["stress"][k], want), k
    # ... and the blocks take part in the solve
    ctx.calc_mobility(1.0)
    rep = ctx.solve_constraints(np.zeros(6 * n), 1e-4, 1e-5, 200, 0)
    assert rep.iterations > 0 and np.abs(ctx.get_force_velocity()["velB"]).max() > 0
