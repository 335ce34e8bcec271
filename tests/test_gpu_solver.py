"""GPU mobility / constraint operator / BCQP loops vs the CPU oracle, through the C ABI.

Tolerances (fp64): the elementwise arithmetic of the GPU kernels is compiled without FMA contraction and
mirrors the oracle's expression order; only reduction orders differ.  Operator parity is asserted at
1e-12 of the vector's max norm, BBPGD iterates at 1e-8 relative at equal iteration count while the
iteration is young, final rod velocities at the level the stopping tolerance allows.
"""
import numpy as np
import pytest

from scenarios import random_rods, thermal_velocity
from test_gpu_collect import gpu_collect

pytestmark = pytest.mark.gpu

MU, DT = 1.0, 1e-4


def setup_case(ctx, oracle, n=2000, box=1.6, seed=0, pbc=(1, 1, 1), frac_sphere=0.0, frac_immovable=0.0, L=0.25):
    rods = random_rods(n, box, length=L, radius=0.0125, seed=seed, frac_sphere=frac_sphere,
                       frac_immovable=frac_immovable)
    lo, hi = [0, 0, 0], [box] * 3
    blocks = gpu_collect(ctx, rods, lo, hi, pbc, 0.025).copy()
    pos = oracle.wrap_positions(rods["pos"], lo, hi, pbc)
    orods = oracle.make_rods(rods["gid"], rods["radius"], rods["length"], pos, rods["quat"], 1.0, 1.0, 0.025)
    ctx.calc_mobility(MU)
    return rods, orods, blocks


def relerr(a, b):
    s = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() / s


def test_mobility_apply(ctx, oracle):
    rods, orods, _ = setup_case(ctx, oracle, n=500, frac_sphere=0.3, frac_immovable=0.2, seed=4)
    x = np.random.default_rng(0).normal(size=6 * 500)
    y = ctx.mobility_apply(x)
    M = oracle.build_mobility(orods, rods["immovable"], MU)
    assert relerr(y, M @ x) < 1e-13
    imm = np.repeat(rods["immovable"] != 0, 6)
    assert np.all(y[imm] == 0)


@pytest.mark.parametrize("frac_sphere,frac_imm", [(0.0, 0.0), (0.3, 0.1)])
def test_operator_apply(ctx, oracle, frac_sphere, frac_imm):
    rods, orods, blocks = setup_case(ctx, oracle, seed=1, frac_sphere=frac_sphere, frac_immovable=frac_imm)
    nc = len(blocks)
    assert nc > 1000
    vnc = thermal_velocity(rods, MU, DT, seed=2)
    ctx.setup_constraints(vnc, DT)
    x = np.abs(np.random.default_rng(1).normal(size=nc))
    y, f, v = ctx.operator_apply(x, want_force_vel=True)
    yo, fo, vo = oracle.operator_apply(blocks, orods, rods["immovable"], MU, DT, x)
    assert relerr(f, fo) < 1e-12
    assert relerr(v, vo) < 1e-12
    assert relerr(y, yo) < 1e-12
    # symmetry / positive semi-definiteness of D^T M D (size-independent property)
    z = np.random.default_rng(2).normal(size=nc)
    yz = ctx.operator_apply(z)
    assert abs(x @ yz - z @ y) < 1e-10 * max(abs(x @ yz), 1.0)
    assert z @ yz > -1e-9 * np.abs(yz).max() * np.abs(z).max()


def add_bilateral(oracle, rods, orods, nb, seed=0):
    """links between random rod pairs, built like SylinderSystem::collectLinkBilateral (:1386-1482)"""
    from oracle.pyoracle import BLOCK_DTYPE

    rng = np.random.default_rng(seed)
    n = len(orods)
    b = np.zeros(nb, dtype=BLOCK_DTYPE)
    for k in range(nb):
        i, j = rng.choice(n, size=2, replace=False)
        I, J = orods[i], orods[j]
        Pp = I["pos"] + I["direction"] * (0.5 * I["length"])
        Qm = J["pos"] - J["direction"] * (0.5 * J["length"])
        rvec = Qm - Pp
        delta0 = np.linalg.norm(rvec) - I["radius"] - J["radius"] - 0.0
        nI = (Pp - Qm) / np.linalg.norm(Pp - Qm)
        b[k]["delta0"] = delta0 * 1e-3
        b[k]["gamma"] = 0.0
        b[k]["gidI"], b[k]["gidJ"] = I["gid"], J["gid"]
        b[k]["globalIndexI"], b[k]["globalIndexJ"] = I["globalIndex"], J["globalIndex"]
        b[k]["bilateral"] = 1
        b[k]["kappa"] = 1000.0 if k % 3 else 0.0
        b[k]["normI"], b[k]["normJ"] = nI, -nI
        b[k]["posI"], b[k]["posJ"] = Pp - I["pos"], Qm - J["pos"]
        b[k]["labI"], b[k]["labJ"] = Pp, Qm
        b[k]["stress"] = rng.normal(size=9)
    return b


def add_one_sided(orods, nb, seed=0):
    """wall-type blocks as SylinderSystem::collectBoundaryCollision builds them (:1093-1150)"""
    from oracle.pyoracle import BLOCK_DTYPE

    rng = np.random.default_rng(seed)
    b = np.zeros(nb, dtype=BLOCK_DTYPE)
    for k in range(nb):
        I = orods[rng.integers(len(orods))]
        end = I["pos"] + I["direction"] * (0.5 * I["length"])
        nrm = np.array([0.0, 0.0, 1.0])
        b[k]["delta0"] = -0.001 * rng.uniform()
        b[k]["gidI"] = b[k]["gidJ"] = I["gid"]
        b[k]["globalIndexI"] = b[k]["globalIndexJ"] = I["globalIndex"]
        b[k]["oneSide"] = 1
        b[k]["normI"] = b[k]["normJ"] = nrm
        b[k]["posI"] = b[k]["posJ"] = end - I["pos"]
        b[k]["labI"], b[k]["labJ"] = end, end - nrm * 0.01
    return b


def test_operator_with_bilateral_and_one_sided(ctx, oracle):
    rods, orods, blocks = setup_case(ctx, oracle, n=1500, seed=6)
    extra = np.concatenate([add_bilateral(oracle, rods, orods, 200, 1), add_one_sided(orods, 100, 2)])
    ctx.append_constraints(extra)
    allb = np.concatenate([blocks, extra])
    assert ctx.num_constraints() == len(allb)
    got = ctx.get_constraints(with_stress=False)
    for f in ("gidI", "gidJ", "globalIndexI", "globalIndexJ", "delta0", "oneSide", "bilateral", "kappa", "normI",
              "normJ", "posI", "posJ"):
        assert np.array_equal(got[f], allb[f]), f
    assert np.array_equal(got["stress"][len(blocks):], extra["stress"])
    ctx.setup_constraints(None, DT)
    x = np.random.default_rng(3).normal(size=len(allb))
    y, f, v = ctx.operator_apply(x, want_force_vel=True)
    yo, fo, vo = oracle.operator_apply(allb, orods, rods["immovable"], MU, DT, x)
    assert relerr(f, fo) < 1e-12 and relerr(v, vo) < 1e-12 and relerr(y, yo) < 1e-12


def check_solution(ctx, oracle, rods, orods, blocks, vnc, res, max_ite, choice, young=30):
    rep = ctx.solve_constraints(vnc, DT, res, max_ite, choice)
    ref = oracle.solve_constraints(blocks, orods, rods["immovable"], MU, vnc if vnc is not None else
                                   np.zeros(6 * len(orods)), DT, res, max_ite, choice)
    hist = ctx.get_history()
    ho = ref["history"]
    # iterate-by-iterate parity while the iteration is young (equal iteration count)
    m = min(young, len(hist), len(ho))
    assert m >= 2
    np.testing.assert_allclose(hist[:m, 4], ho[:m, 4], rtol=1e-8)  # residuals
    np.testing.assert_allclose(hist[:m, 3], ho[:m, 3], rtol=1e-8)  # step sizes
    assert np.array_equal(hist[:m, 0], ho[:m, 0]) and np.array_equal(hist[:m, 5], ho[:m, 5])
    return rep, ref, hist


@pytest.mark.parametrize("colbuf,zero_frac", [(0.025, 0.0), (0.025, 0.85), (0.15, 0.5), (0.15, 1.0)])
def test_force_kernels_agree_bit_for_bit(ctx, oracle, colbuf, zero_frac):
    """k_force_vel_rec (slot records + slot bitmap kept by the tail kernel: the default), k_force_vel_act (rod-major
    slots, x == 0 skipped) and k_slot_x + k_rod_sum against k_force_vel_lm (level-major, dense) and the oracle.
    Skipping a zero multiplier must not change a single bit; colbuf 0.15 gives 32-rod groups with more than 256
    slots (several batches per warp) and rods with more than 32 slots (several bitmap words per rod)."""
    rods = random_rods(2000, 1.6, seed=21, frac_sphere=0.1)
    lo, hi = [0, 0, 0], [1.6] * 3
    blocks = gpu_collect(ctx, rods, lo, hi, (1, 1, 1), colbuf).copy()
    orods = oracle.make_rods(rods["gid"], rods["radius"], rods["length"], oracle.wrap_positions(rods["pos"], lo, hi),
                             rods["quat"], 1.0, 1.0, colbuf)
    nc = len(blocks)
    assert nc > (20000 if colbuf > 0.1 else 1000)
    ctx.calc_mobility(MU)
    rng = np.random.default_rng(5)
    x = np.abs(rng.normal(size=nc))
    x[rng.uniform(size=nc) < zero_frac] = 0.0
    x[::7] *= -1.0  # negative multipliers and -0.0 occur in plain operator applies
    res = {}
    def select(kern):  # 3: k_force_vel_rec gathering {x, g} by row id (default), 30: with {x, g} copied into the records,
        # 32: records hold M * column (the velocity is summed directly: same numbers up to rounding, not bit for bit)
        ctx.set_option("force_kernel", 3 if kern in (30, 32) else kern)
        ctx.set_option("rec_mode", {30: 0, 32: 2}.get(kern, 1))  # (3 = rec_mode 1 here; the library's default is 2)

    for kern in (3, 30, 32, 1, 2, 0):  # 1 k_force_vel_act, 2 k_slot_x + k_rod_sum, 0 dense level-major
        select(kern)
        ctx.setup_constraints(None, DT)
        res[kern] = ctx.operator_apply(x, want_force_vel=True)
        if kern == 3:  # a second apply to another vector on the same setup: stale records / bits must not leak
            x2 = np.where(rng.uniform(size=nc) < 0.5, 0.0, rng.normal(size=nc))
            res["x2"] = (x2, ctx.operator_apply(x2, want_force_vel=True))
            res[3] = ctx.operator_apply(x, want_force_vel=True)
    for kern in (3, 30, 1, 2):
        for a, b in zip(res[0], res[kern]):
            assert np.array_equal(a, b)
    y32, f32, v32 = res[32]
    assert np.array_equal(f32, res[0][1])  # the force comes from the row geometry: bit for bit
    assert relerr(y32, res[0][0]) < 1e-13 and relerr(v32, res[0][2]) < 1e-13
    ctx.set_option("force_kernel", 0)
    ctx.setup_constraints(None, DT)
    for a, b in zip(ctx.operator_apply(res["x2"][0], want_force_vel=True), res["x2"][1]):
        assert np.array_equal(a, b)
    yo, fo, vo = oracle.operator_apply(blocks, orods, rods["immovable"], MU, DT, x)
    for a, b in zip(res[1], (yo, fo, vo)):
        assert relerr(a, b) < 1e-12 or np.abs(b).max() == 0
    # the BBPGD loop (x recomputed on the fly from {x_prev, g_prev}) gives the same iterates with both kernels
    vnc = thermal_velocity(rods, MU, DT, seed=9)
    gam = {}
    for kern in (3, 30, 32, 1, 2, 0):
        select(kern)
        rep = ctx.solve_constraints(vnc, DT, 1e-30, 15, 0)
        assert rep.iterations == 15
        gam[kern] = (ctx.get_gamma(), ctx.get_force_velocity()["velU"], ctx.get_history())
    for kern in (3, 30, 1, 2):
        for a, b in zip(gam[0], gam[kern]):
            assert np.array_equal(a, b)
    assert relerr(gam[32][0], gam[0][0]) < 1e-9 and relerr(gam[32][1], gam[0][1]) < 1e-9
    np.testing.assert_allclose(gam[32][2][:, 3:5], gam[0][2][:, 3:5], rtol=1e-8)
    # ... and a long run on the record kernels: rows leave and re-enter the live set many times
    for kern in (3, 30, 0):
        select(kern)
        rep = ctx.solve_constraints(vnc, DT, 1e-30, 150, 0)
        gam[kern] = (ctx.get_gamma(), ctx.get_force_velocity()["velU"], ctx.get_history())
    for kern in (3, 30):
        for a, b in zip(gam[0], gam[kern]):
            assert np.array_equal(a, b)
    select(32)  # back to the defaults


def test_bbpgd_matches_oracle_iterates(ctx, oracle):
    rods, orods, blocks = setup_case(ctx, oracle, n=2000, seed=7)
    vnc = thermal_velocity(rods, MU, DT, seed=3)
    # fixed small iteration budget: both sides run exactly 25 iterations, gamma must agree to 1e-8
    rep, ref, hist = check_solution(ctx, oracle, rods, orods, blocks, vnc, 1e-30, 25, 0, young=26)
    assert rep.iterations == 25 == ref["nIte"] and rep.matvecs == ref["mvCount"]
    g = ctx.get_gamma()
    assert relerr(g, ref["gamma"]) < 1e-8
    out = ctx.get_force_velocity()
    for k in ("forceU", "velU", "forceB", "velB"):
        assert relerr(out[k], ref[k]) < 1e-8 or np.abs(ref[k]).max() == 0, k


def test_bbpgd_converges_and_velocities_match(ctx, oracle):
    rods, orods, blocks = setup_case(ctx, oracle, n=2000, seed=8)
    vnc = thermal_velocity(rods, MU, DT, seed=4)
    res = 1e-7  # tol = res/dt = 1e-3 velocity units
    rep, ref, hist = check_solution(ctx, oracle, rods, orods, blocks, vnc, res, 20000, 0)
    assert rep.status == 0 and rep.residual < res / DT
    assert ref["resFinal"] < res / DT
    out = ctx.get_force_velocity()
    scale = np.abs(ref["velU"]).max()
    # two converged solves agree to O(tol): the velocity u = M D gamma is unique, gamma is not
    assert np.abs(out["velU"] - ref["velU"]).max() < 50 * res / DT
    assert scale > 100 * res / DT
    # KKT check of the GPU solution against the oracle operator: min(gamma, g) ~ 0
    g = ctx.get_gamma()
    assert g.min() >= 0
    y, _, _ = oracle.operator_apply(blocks, orods, rods["immovable"], MU, DT, g)
    # q = delta0/dt + D^T vnc
    DTm, d0, _, _, _ = oracle.build_dtrans_dense(blocks, len(orods))
    grad = y + d0 / DT + DTm @ vnc
    kkt = np.where(g > 0, np.abs(grad), np.minimum(grad, 0))
    assert np.abs(kkt).max() < 1.01 * res / DT


def test_bbpgd_with_bilateral(ctx, oracle):
    rods, orods, blocks = setup_case(ctx, oracle, n=1500, seed=9, frac_immovable=0.05)
    extra = np.concatenate([add_bilateral(oracle, rods, orods, 150, 5), add_one_sided(orods, 60, 6)])
    ctx.append_constraints(extra)
    allb = np.concatenate([blocks, extra])
    vnc = thermal_velocity(rods, MU, DT, seed=5)
    rep, ref, hist = check_solution(ctx, oracle, rods, orods, allb, vnc, 1e-30, 20, 0, young=21)
    g = ctx.get_gamma()
    assert relerr(g, ref["gamma"]) < 1e-8
    out = ctx.get_force_velocity()
    assert np.abs(ref["velB"]).max() > 0
    for k in ("forceU", "velU", "forceB", "velB"):
        assert relerr(out[k], ref[k]) < 1e-7, k
    # write-back: gamma into the blocks, stress scaled (ConstraintCollector.cpp:439-461)
    wb = ctx.get_constraints(with_stress=True, write_back=True)
    assert np.array_equal(wb["gamma"], g)
    want = oracle.writeback_gamma(allb, g)
    np.testing.assert_allclose(wb["stress"][len(blocks):], want["stress"][len(blocks):], rtol=1e-15)


def test_bbpgd_itemax_returns_older_iterate(ctx, oracle):
    # quirk 3: on the iteMax exit xk is the OLDER iterate, force/vel the newer one
    rods, orods, blocks = setup_case(ctx, oracle, n=800, seed=10)
    vnc = thermal_velocity(rods, MU, DT, seed=6)
    for ite in (0, 1, 2, 7):
        rep = ctx.solve_constraints(vnc, DT, 1e-30, ite, 0)
        ref = oracle.solve_constraints(blocks, orods, rods["immovable"], MU, vnc, DT, 1e-30, ite, 0)
        assert rep.iterations == ite
        assert relerr(ctx.get_gamma(), ref["gamma"]) < 1e-9
        assert relerr(ctx.get_force_velocity()["velU"], ref["velU"]) < 1e-9


def test_apgd_matches_oracle(ctx, oracle):
    rods, orods, blocks = setup_case(ctx, oracle, n=1200, seed=11)
    vnc = thermal_velocity(rods, MU, DT, seed=7)
    rep, ref, hist = check_solution(ctx, oracle, rods, orods, blocks, vnc, 1e-30, 12, 1, young=13)
    assert rep.matvecs == ref["mvCount"]
    assert relerr(ctx.get_gamma(), ref["gamma"]) < 1e-7
    rep = ctx.solve_constraints(vnc, DT, 1e-6, 5000, 1)
    assert rep.status == 0 and rep.residual < 1e-6 / DT


def test_empty_constraint_set(ctx, oracle):
    rods = random_rods(50, 50.0, seed=1)  # far apart: no contacts
    blocks = gpu_collect(ctx, rods, [0, 0, 0], [50.0] * 3, (0, 0, 0), 0.025)
    assert len(blocks) == 0
    ctx.calc_mobility(MU)
    rep = ctx.solve_constraints(None, DT, 1e-5, 100, 0)
    assert rep.iterations == 0 and rep.n_constraints == 0
    out = ctx.get_force_velocity()
    assert all(np.all(v == 0) for v in out.values())


def test_step_euler(ctx, oracle):
    rods, orods, blocks = setup_case(ctx, oracle, n=600, seed=12)
    vnc = thermal_velocity(rods, MU, DT, seed=8)
    ctx.solve_constraints(vnc, DT, 1e-6, 2000, 0)
    out = ctx.get_force_velocity()
    p0, q0 = ctx.get_rod_state()
    ctx.step_euler(DT)
    p1, q1 = ctx.get_rod_state()
    v = (vnc + out["velU"] + out["velB"]).reshape(-1, 6)
    np.testing.assert_allclose(p1, p0 + v[:, :3] * DT, rtol=0, atol=1e-15)
    # quaternion rotated by omega*dt (EquatnHelper.hpp:74-90), stays normalised
    np.testing.assert_allclose(np.linalg.norm(q1, axis=1), 1.0, atol=1e-14)
    w = np.linalg.norm(v[:, 3:], axis=1)
    ang = 2 * np.arccos(np.clip(np.abs((q0 * q1).sum(axis=1)), -1, 1))
    np.testing.assert_allclose(ang, w * DT, atol=1e-7)


def test_velocity_noncon_async_upload(ctx, oracle):
    """alens_set_velocity_noncon_async (side-stream H2D from pinned memory, overlapping the pair search) gives the
    same solve as passing the vector to alens_solve_constraints."""
    import torch

    rods = random_rods(3000, 1.8, seed=31)
    lo, hi = [0.0] * 3, [1.8] * 3
    vnc = thermal_velocity(rods, MU, DT, seed=12)
    pinned = torch.from_numpy(vnc.copy()).pin_memory()
    ctx.set_domain(lo, hi, (1, 1, 1))
    ctx.set_collision_params(1.0, 1.0, 0.025)
    out = []
    for mode in ("sync", "async"):
        ctx.set_rods(rods["gid"], rods["pos"], rods["quat"], rods["length"], rods["radius"], rods["immovable"])
        if mode == "async":
            ctx.set_velocity_noncon_async_raw(pinned.data_ptr())
        ctx.collect_pair_collision()
        ctx.calc_mobility(MU)
        rep = ctx.solve_constraints(vnc if mode == "sync" else None, DT, 1e-6, 200, 0)
        out.append((rep.iterations, ctx.get_gamma(), ctx.get_force_velocity()["velU"]))
    assert out[0][0] == out[1][0] > 0
    assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])


@pytest.mark.parametrize("monolayer", [False, True])
def test_calc_velocity_noncon(ctx, oracle, monolayer):
    """alens_calc_velocity_noncon = SylinderSystem::calcVelocityNonCon (SylinderSystem.cpp:724-800): M f + vNB + vB with the
    monolayer mask, resident for the solve that follows (compared with a solve that is handed the host vector)."""
    rods, orods, blocks = setup_case(ctx, oracle, n=1500, seed=41, frac_sphere=0.2, frac_immovable=0.1)
    n = len(rods["gid"])
    rng = np.random.default_rng(7)
    f, vnb, vb = rng.normal(size=6 * n), rng.normal(size=6 * n), rng.normal(size=6 * n)
    M = oracle.build_mobility(orods, rods["immovable"], MU)
    mask = np.ones(6 * n)
    if monolayer:
        mask.reshape(-1, 6)[:, [2, 3, 4]] = 0
    want_nb = (M @ f) * mask + vnb * mask
    want = want_nb + vb * mask
    got_nb = ctx.calc_velocity_noncon(f, vnb, vb, monolayer)
    assert relerr(got_nb, want_nb) < 1e-13
    rep = ctx.solve_constraints(None, DT, 1e-30, 10, 0)          # the resident vector
    g_res = ctx.get_gamma()
    rep2 = ctx.solve_constraints(want, DT, 1e-30, 10, 0)         # the same vector from the host
    assert rep.iterations == rep2.iterations == 10
    assert relerr(g_res, ctx.get_gamma()) < 1e-10
    only_v = ctx.calc_velocity_noncon(None, vnb, None, monolayer)  # absent terms
    assert np.array_equal(only_v, vnb * mask + 0.0)


def test_calc_velocity_brown(ctx, oracle):
    """alens_calc_velocity_brown = SylinderSystem::calcVelocityBrown (SylinderSystem.cpp:1020-1091): with the caller's
    normal deviates it reproduces the restated formula (1e-11: Cholesky / trig rounding); with the built-in counter-based
    generator it is reproducible, keyed by gid (independent of rod order) and has the right second moments."""
    n, mu, kbt, dt = 3000, 1.0, 0.00411, 1e-4
    rods = random_rods(n, 3.0, seed=51, frac_sphere=0.2, frac_immovable=0.1)
    def load(r):
        ctx.set_domain([0.0] * 3, [3.0] * 3, (1, 1, 1))
        ctx.set_collision_params(1.0, 1.0, 0.025)
        ctx.set_rods(r["gid"], r["pos"], r["quat"], r["length"], r["radius"], r["immovable"])
        ctx.calc_mobility(mu)
    load(rods)
    W = np.random.default_rng(5).normal(size=(n, 12))
    got = ctx.calc_velocity_brown(kbt, dt, W)
    want = oracle.velocity_brown(rods["quat"], rods["radius"], rods["length"], rods["immovable"], mu, kbt, dt, W)
    assert relerr(got, want) < 1e-11
    imm = np.repeat(rods["immovable"] != 0, 6)
    assert np.all(got[imm] == 0) and np.abs(got[~imm]).min() > 0
    # device generator: same (seed, step) -> same numbers; another step -> different; keyed by gid, not by storage order
    a = ctx.calc_velocity_brown(kbt, dt, None, seed=1234, step=7)
    b = ctx.calc_velocity_brown(kbt, dt, None, seed=1234, step=7)
    c = ctx.calc_velocity_brown(kbt, dt, None, seed=1234, step=8)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    perm = np.random.default_rng(1).permutation(n)
    load({k: v[perm] for k, v in rods.items()})
    ap = ctx.calc_velocity_brown(kbt, dt, None, seed=1234, step=7)
    assert np.array_equal(ap.reshape(n, 6), a.reshape(n, 6)[perm])
    # second moments: <omega omega^T> = 2 kBT / (zRot dt) I for every movable rod (SylinderSystem.cpp:1066)
    om = np.concatenate([ctx.calc_velocity_brown(kbt, dt, None, seed=99, step=s).reshape(n, 6)[:, 3:] for s in range(40)], axis=1)
    from oracle import pyoracle as po
    zr = np.array([po.drag_coeff(float(r), float(l), mu)[2] for r, l in zip(rods["radius"][perm], rods["length"][perm])])
    mov = rods["immovable"][perm] == 0
    ratio = (om[mov] ** 2).mean(axis=1) / (2 * kbt / (zr[mov] * dt))
    assert abs(ratio.mean() - 1) < 0.02 and abs(om[mov].mean()) < 0.02 * np.sqrt((om[mov] ** 2).mean())
