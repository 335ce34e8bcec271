"""The C++ drop-in classes (include/alens_b200/*.hpp: SylinderSystem / ConstraintSolver / ConstraintCollector /
BCQPSolver with the reference's method names) driven like the reference's own main loop, compared with the
ctypes path (bit-identical: same library, deterministic kernels) and with the CPU oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from scenarios import random_rods, thermal_velocity
from test_gpu_solver import add_bilateral, add_one_sided

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_dropin")


def oracle_directions(quat):
    """orientation * ez for (x, y, z, w) quaternions"""
    x, y, z, w = quat.T
    return np.stack([2 * (x * z + w * y), 2 * (y * z - w * x), 1 - 2 * (x * x + y * y)], axis=1)


def build_exe():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")])


def run_dropin(tmp_path, rods, lo, hi, pbc, colbuf, mu, dt, res, max_ite, vnb, host_blocks):
    n = len(rods["gid"])
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<i", n))
        f.write(np.asarray(lo, dtype="<f8").tobytes())
        f.write(np.asarray(hi, dtype="<f8").tobytes())
        f.write(np.asarray(pbc, dtype="<i4").tobytes())
        f.write(struct.pack("<dddd", colbuf, mu, dt, res))
        f.write(struct.pack("<i", max_ite))
        for i in range(n):
            f.write(struct.pack("<i", int(rods["gid"][i])))
            f.write(struct.pack("<dd", rods["radius"][i], rods["length"][i]))
            f.write(rods["pos"][i].astype("<f8").tobytes())
            f.write(rods["quat"][i].astype("<f8").tobytes())
        f.write(np.asarray(vnb, dtype="<f8").tobytes())
        f.write(struct.pack("<i", len(host_blocks)))
        f.write(host_blocks.tobytes())
    env = dict(os.environ, OMP_NUM_THREADS="3")
    r = subprocess.run([EXE, fin, fout], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stderr[-2000:])
    from alens_b200 import BLOCK_DTYPE

    buf = open(fout, "rb").read()
    nc, ite, resid = struct.unpack_from("<qid", buf, 0)
    off = 8 + 4 + 8
    out = {}
    for k in ("forceU", "velU", "forceB", "velB"):
        out[k] = np.frombuffer(buf, dtype="<f8", count=6 * n, offset=off)
        off += 48 * n
    st = np.frombuffer(buf, dtype="<f8", count=7 * n, offset=off).reshape(n, 7)
    off += 56 * n
    out["pos"], out["quat"] = st[:, :3], st[:, 3:]
    out["gamma"] = np.frombuffer(buf, dtype="<f8", count=nc, offset=off)
    off += 8 * nc
    out["blocks"] = np.frombuffer(buf, dtype=BLOCK_DTYPE, count=nc, offset=off)
    off += BLOCK_DTYPE.itemsize * nc
    diag = np.frombuffer(buf, dtype="<f8", count=31, offset=off)
    out["stress_uni"], out["stress_bi"] = diag[:9].reshape(3, 3), diag[9:18].reshape(3, 3)
    out["order_p"], out["order_Q"], out["vol_frac"] = diag[18:21], diag[21:30].reshape(3, 3), diag[30]
    out["max_gid"] = struct.unpack_from("<ii", buf, off + 31 * 8)
    out["record"] = r.stderr
    return nc, ite, resid, out


def test_dropin_step_equals_capi_path(tmp_path, ctx, oracle):
    build_exe()
    n, box, colbuf, mu, dt, res, max_ite = 1500, 1.4, 0.025, 1.0, 1e-4, 1e-6, 5000
    rods = random_rods(n, box, seed=21)
    lo, hi, pbc = [0.0] * 3, [box] * 3, (1, 1, 0)
    vnb = thermal_velocity(rods, mu, dt, seed=5)
    pos_w = oracle.wrap_positions(rods["pos"], lo, hi, pbc)
    orods = oracle.make_rods(rods["gid"], rods["radius"], rods["length"], pos_w, rods["quat"], 1.0, 1.0, colbuf)
    one = add_one_sided(orods, 30, 2)
    one["stress"] = np.random.default_rng(8).normal(size=(30, 9))  # counted only by sumLocalConstraintStress(withOneSide)
    host = np.concatenate([add_bilateral(oracle, rods, orods, 40, 1), one])
    nc, ite, resid, out = run_dropin(tmp_path, rods, lo, hi, pbc, colbuf, mu, dt, res, max_ite, vnb, host)
    assert "RECORD: BCQP residue" in out["record"]

    # the same step through the ctypes binding; host blocks in the order the pool is flattened
    # (3 OpenMP queues, round-robin push -> queue-major order)
    order = np.concatenate([np.arange(q, len(host), 3) for q in range(3)])
    ctx.set_domain(lo, hi, pbc)
    ctx.set_collision_params(1.0, 1.0, colbuf)
    ctx.set_rods(rods["gid"], rods["pos"], rods["quat"], rods["length"], rods["radius"], None, wrap=True)
    ctx.calc_mobility(mu)
    ncoll = ctx.collect_pair_collision()
    ctx.append_constraints(host[order])
    rep = ctx.solve_constraints(vnb, dt, res, max_ite, 0)
    assert (nc, ite) == (ncoll + len(host), rep.iterations) and resid == rep.residual
    fv = ctx.get_force_velocity()
    for k in ("forceU", "velU", "forceB", "velB"):
        assert np.array_equal(out[k], fv[k]), k
    assert np.array_equal(out["gamma"], ctx.get_gamma())
    wb = ctx.get_constraints(with_stress=True, write_back=True)
    assert out["blocks"].tobytes() == wb.tobytes()
    # per-step diagnostics of the mirror (calcConStress / calcOrderParameter / calcVolFrac / getMaxGid): the stress sums
    # come from the device-side reduction and equal the sums over the written-back blocks (KBT = 0.5 in test_dropin.cpp)
    uni, bi = ctx.sum_constraint_stress()
    two = wb["oneSide"] == 0
    ref_uni = wb["stress"][two & (wb["bilateral"] == 0)].sum(axis=0).reshape(3, 3)
    ref_bi = wb["stress"][two & (wb["bilateral"] != 0)].sum(axis=0).reshape(3, 3)
    assert np.abs(ref_uni).max() > 0 and np.abs(ref_bi).max() > 0
    np.testing.assert_allclose(uni, ref_uni, rtol=1e-12, atol=1e-12 * np.abs(ref_uni).max())
    np.testing.assert_allclose(bi, ref_bi, rtol=1e-12, atol=1e-12 * np.abs(ref_bi).max())
    uni1, _ = ctx.sum_constraint_stress(with_one_side=True)
    ref_uni1 = wb["stress"][wb["bilateral"] == 0].sum(axis=0).reshape(3, 3)
    np.testing.assert_allclose(uni1, ref_uni1, rtol=1e-12, atol=1e-12 * np.abs(ref_uni1).max())
    assert np.abs(uni1 - uni).max() > 0
    np.testing.assert_allclose(out["stress_uni"], uni / (n * 0.5), rtol=1e-14, atol=0)
    np.testing.assert_allclose(out["stress_bi"], bi / (n * 0.5), rtol=1e-14, atol=0)
    d = oracle_directions(out["quat"])  # the diagnostics are taken after runStep: the moved rods
    np.testing.assert_allclose(out["order_p"], d.mean(axis=0), rtol=0, atol=1e-13)
    np.testing.assert_allclose(out["order_Q"], (d[:, :, None] * d[:, None, :]).mean(axis=0) - np.eye(3) / 3, rtol=0, atol=1e-13)
    vol = np.pi * (0.25 * rods["length"] * (2 * rods["radius"]) ** 2 + (2 * rods["radius"]) ** 3 / 6)
    assert abs(out["vol_frac"] - vol.sum() / box**3) < 1e-9 * out["vol_frac"]  # the reference's pi has 11 digits
    assert out["max_gid"] == (int(rods["gid"].max()),) * 2
    ctx.step_euler(dt)
    p, q = ctx.get_rod_state()
    assert np.array_equal(out["pos"], p) and np.array_equal(out["quat"], q)

    # and against the oracle (converged: velocities agree to O(tol))
    allb = np.concatenate([ctx.get_constraints(with_stress=False)[:ncoll], host[order]])
    ref = oracle.solve_constraints(allb, orods, np.zeros(n, dtype=np.int32), mu, vnb, dt, res, max_ite, 0)
    assert rep.residual < res / dt and ref["resFinal"] < res / dt
    assert np.abs(out["velU"] - ref["velU"]).max() < 50 * res / dt
    assert np.abs(out["velB"] - ref["velB"]).max() < 50 * res / dt


def test_mirror_main_program_follows_the_reference_main_program(tmp_path):
    """the C++ mirror driven like SimToolbox/Sylinder/SylinderSystem_main.cpp -- SylinderSystem(configFile, posFile, argc,
    argv), then prepareStep / runStep, then writeResult -- from the SAME RunConfig.yaml and SylinderInitial.dat as the
    reference's own SylinderSystem (oracle/_ref/libalens_refsys.so).  A dilute suspension of two-rod filaments whose links
    are stretched (linkGap below the actual gap): bilateral constraints only, so both sides solve the same list."""
    from oracle import pyrefsys as pr

    if not pr.available():
        pytest.skip("oracle/_ref/libalens_refsys.so missing")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")])
    rng = np.random.default_rng(5)
    nf, L, R = 60, 0.5, 0.0125
    box = 12.0
    centers = rng.uniform(1.5, box - 1.5, size=(nf, 3))
    d = rng.normal(size=(nf, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    gap = 0.08  # > colBuf: the two rods of a filament do not collide
    rods = dict(gid=np.arange(2 * nf, dtype=np.int32), pos=np.zeros((2 * nf, 3)), quat=np.zeros((2 * nf, 4)),
                length=np.full(2 * nf, L), radius=np.full(2 * nf, R), immovable=np.zeros(2 * nf, dtype=np.uint8))
    from scenarios import quat_from_z_to
    rods["pos"][0::2] = centers - d * (0.5 * L + R + 0.5 * gap)
    rods["pos"][1::2] = centers + d * (0.5 * L + R + 0.5 * gap)
    rods["quat"][0::2] = rods["quat"][1::2] = quat_from_z_to(d)
    rods["immovable"][0] = 1
    links = [(2 * k, 2 * k + 1) for k in range(nf)]
    cfg = dict(pr.DEFAULTS)
    cfg.update(simBoxLow=[0.0] * 3, simBoxHigh=[box] * 3, simBoxPBC=[True, False, True], sylinderColBuf=0.025, viscosity=0.9,
               dt=1e-4, conResTol=1e-9, conMaxIte=100000, linkKappa=150.0, linkGap=0.02, initPreSteps=0, timeSnap=1.0,
               logLevel=5, timerLevel=5)
    work = tmp_path / "mirror"
    (work / "result" / "result0-399").mkdir(parents=True)
    pr.write_yaml(str(work / "RunConfig.yaml"), cfg)
    pr.write_dat(str(work / "SylinderInitial.dat"), rods, links)
    exe = os.path.join(ROOT, "tests", "cpp", "test_system_main")
    steps = 3
    # "restart": afterwards a second system resumes from a snapshot (reinitialize, SylinderSystem.cpp:106-175) and follows the first
    r = subprocess.run([exe, str(steps), str(work / "out.bin"), "restart"], cwd=str(work), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "restart ok" in r.stdout, r.stderr[-2000:]
    assert (work / "TimeStepInfo.txt").read_text().split() == [str(cfg["rngSeed"]), str(steps), "1", "Sylinder_1.pvtp"]
    raw = (work / "out.bin").read_bytes()
    n = int(np.frombuffer(raw[:4], dtype=np.int32)[0])
    got = np.frombuffer(raw[4:4 + 568 * n], dtype=pr.SYLINDER_DTYPE)
    nl = int(np.frombuffer(raw[4 + 568 * n:8 + 568 * n], dtype=np.int32)[0])
    glinks = np.frombuffer(raw[8 + 568 * n:], dtype=np.int32).reshape(nl, 2)
    assert n == 2 * nf and sorted(map(tuple, glinks.tolist())) == links
    # the reference, from the same two files
    s = pr.RefSystem(yaml_file=str(work / "RunConfig.yaml"), pos_file=str(work / "SylinderInitial.dat"), nthreads=1)
    init = s.sylinders().copy()
    for _ in range(steps):
        s.prepare_step()
        s.run_step()
        assert s.constraints()["bilateral"].all() and len(s.constraints()) == nf  # links only
    s.prepare_step()
    want = s.sylinders()
    assert np.array_equal(got["gid"], want["gid"]) and np.array_equal(got["isImmovable"], want["isImmovable"])
    assert np.array_equal(got["length"], want["length"]) and np.array_equal(got["radius"], want["radius"])
    moved = np.abs(want["pos"] - init["pos"]).max()
    assert moved > 1e-4
    assert np.abs(got["pos"] - want["pos"]).max() < 1e-6 * moved
    assert np.abs(got["orientation"] - want["orientation"]).max() < 1e-8
    assert np.all(got["pos"][0] == init["pos"][0])  # the immovable rod
    for name in ("Sylinder_r0_0.vtp", "Sylinder_0.pvtp", "ConBlock_r0_0.vtp", "ConBlock_0.pvtp", "SylinderAscii_0.dat"):
        assert (work / "result" / "result0-399" / name).stat().st_size > 100, name
    s.close()


@pytest.mark.parametrize("nranks", [2, 3])
def test_mirror_on_several_ranks_with_device_side_migration(nranks):
    """include/alens_b200/SylinderSystem.hpp on `nranks` ranks (one host thread + one context each; one GPU each when the
    box has them): Brownian steps with rods crossing slab faces, against one SylinderSystem holding everything
    (tests/cpp/test_multirank.cpp: each gid exactly once, group travels with the rod, contiguous globalIndex, positions)"""
    import torch

    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")])
    exe = os.path.join(ROOT, "tests", "cpp", "test_multirank")
    env = dict(os.environ)
    if torch.cuda.device_count() >= nranks:
        env["ALENS_TEST_DEVICES"] = ",".join(str(i) for i in range(nranks))
    # Ranks that SHARE one GPU (the 1-GPU test box) wait for each other inside kernels; anything that makes the driver
    # synchronise the whole device on behalf of a lagging rank stalls the waiting rank until its 4-s limit and the step
    # fails with ALENS_ERR_COMM.  Seen once in ~10 runs on a shared device, never with one GPU per rank (the configuration
    # bench.py times): one retry, with the first attempt's output kept in the report.
    first = None
    for attempt in range(2):
        r = subprocess.run([exe, str(nranks), "5"], capture_output=True, text=True, timeout=600, env=env)
        if r.returncode == 0 and "PASS" in r.stdout:
            break
        if first is None:
            first = (r.stdout[-1500:], r.stderr[-1500:])
            if "ALENS_TEST_DEVICES" in env:
                break  # one GPU per rank: no excuse
    if first is not None:
        import warnings

        warnings.warn(f"test_multirank {nranks}: first attempt failed: {first}")
    assert r.returncode == 0 and "PASS" in r.stdout, (r.stdout[-1500:], r.stderr[-1500:], first)
