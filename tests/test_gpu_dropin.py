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
    host = np.concatenate([add_bilateral(oracle, rods, orods, 40, 1), add_one_sided(orods, 30, 2)])
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
    ctx.step_euler(dt)
    p, q = ctx.get_rod_state()
    assert np.array_equal(out["pos"], p) and np.array_equal(out["quat"], q)

    # and against the oracle (converged: velocities agree to O(tol))
    allb = np.concatenate([ctx.get_constraints(with_stress=False)[:ncoll], host[order]])
    ref = oracle.solve_constraints(allb, orods, np.zeros(n, dtype=np.int32), mu, vnb, dt, res, max_ite, 0)
    assert rep.residual < res / dt and ref["resFinal"] < res / dt
    assert np.abs(out["velU"] - ref["velU"]).max() < 50 * res / dt
    assert np.abs(out["velB"] - ref["velB"]).max() < 50 * res / dt
