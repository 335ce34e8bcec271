#!/usr/bin/env python
"""bench.py -- constraint-solve steps/sec on synthetic random rod suspensions (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            our arm   (libalens_b200.so on cuda)
    python bench.py --impl reference --gpus N --steps K ...  reference CPU algorithm on the host cores

One "step" = one pass of the collision-constraint hot path over the whole suspension:
prepareStep (box wrap, cell list, sorted SoA) -> collectPairCollision -> calcMobOperator ->
ConstraintSolver::setup -> BBPGD to conResTol -> uni/bi split.  Workload at N=1:
BASELINE.json configs[3] "synthetic 1M random spherocylinders at high volume fraction, 1 GPU":
1e6 rods L=0.25 D=0.025 in a periodic cube at phi=0.10, colBuf=0.025, mu=1, dt=1e-5, conResTol=1e-5,
relaxed by a few untimed steps, driven by a seeded Brownian-scale velNonCon.

`value`  : steps/s with rods + velNonCon resident in HBM (CUDA events on the library's stream).
`e2e`    : steps/s through the C ABI with pinned HOST buffers: H2D of the rod state and velNonCon and
           D2H of the 4x6n force/velocity result inside the timed region.
`roofline`: dominant BCQP kernel, algorithmic bytes / CUDA-event duration vs MEASURED_PEAKS.json.
`cpu_baseline`: the reference algorithm on the host cores (oracle port; FDPS reference build for the
           pair search when oracle/_ref exists) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from scenarios import box_for_volume_fraction, random_rods, thermal_velocity  # noqa: E402

L_ROD, R_ROD, COLBUF, MU, DT, RES, MAXITE = 0.25, 0.0125, 0.025, 1.0, 1e-5, 1e-5, 10000
SLAB_AXIS = 0  # multi-GPU runs: slabs along x (2: along z, the slowest cell axis -- see profiles/README.md, late_halo)
SEED = 1234


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--rods", type=float, default=1e6, help="rods per GPU (weak scaling)")
    p.add_argument("--phi", type=float, default=0.10)
    p.add_argument("--relax", type=int, default=4, help="untimed relaxation steps before measuring")
    p.add_argument("--cpu-iters", type=int, default=12, help="BBPGD iterations in the bounded CPU sample")
    p.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    return p.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for ln in self.f.read().splitlines():
            t = [x.strip() for x in ln.split(",")]
            if len(t) < 9:
                continue
            try:
                sm.append(float(t[1]))
                mx.append(float(t[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), t[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def make_workload(n, phi, seed):
    box = box_for_volume_fraction(n, L_ROD, R_ROD, phi)
    rods = random_rods(n, box, L_ROD, R_ROD, seed=seed)
    return rods, box


def relax_on_gpu(ctx, rods, box, steps, configured=False):
    """untimed: the reference's own initPreSteps loop (SylinderSystem.cpp:88-101) run on the device"""
    if not configured:
        ctx.set_domain([0.0] * 3, [box] * 3, [1, 1, 1])
        ctx.set_collision_params(1.0, 1.0, COLBUF)
    ctx.set_rods(rods["gid"], rods["pos"], rods["quat"], rods["length"], rods["radius"], rods["immovable"], wrap=True)
    ctx.set_velocity_noncon(None)
    info = []
    for it in range(steps):
        if it > 0:
            ctx.prepare_step(True)
        nc = ctx.collect_pair_collision()
        ctx.calc_mobility(MU)
        rep = ctx.solve_constraints(None, DT, RES, MAXITE, 0)
        info.append((nc, rep.iterations))
        ctx.step_euler(DT)
    ctx.prepare_step(True)
    pos, quat = ctx.get_rod_state()
    out = dict(rods)
    out["pos"], out["quat"] = pos, quat
    return out, info


def relax_on_cpu(po, rods, box, steps, nthreads):
    """same loop with the oracle (only used by --impl reference when no GPU is visible)"""
    lo, hi, pbc = [0.0] * 3, [box] * 3, [1, 1, 1]
    pos, quat = rods["pos"].copy(), rods["quat"].copy()
    n = len(rods["gid"])
    for _ in range(steps):
        pos = po.wrap_positions(pos, lo, hi)
        orods = po.make_rods(rods["gid"], rods["radius"], rods["length"], pos, quat, 1.0, 1.0, COLBUF)
        blocks = po.collect_pairs(orods, lo, hi, pbc, method="cells", nthreads=nthreads)
        sol = po.solve_constraints(blocks, orods, rods["immovable"], MU, np.zeros(6 * n), DT, RES, MAXITE, 0,
                                   nthreads=nthreads, hist_cap=4)
        v = (sol["velU"] + sol["velB"]).reshape(n, 6)
        pos = pos + v[:, :3] * DT
        w = np.linalg.norm(v[:, 3:], axis=1)
        ok = w > np.finfo(np.float32).eps
        winv = np.where(ok, 1 / np.where(ok, w, 1), 0)
        sw, cw = np.sin(w * DT / 2), np.cos(w * DT / 2)
        s, p = quat[:, 3], quat[:, :3]
        om = v[:, 3:]
        xyz = (s * sw * winv)[:, None] * om + cw[:, None] * p + (sw * winv)[:, None] * np.cross(om, p)
        qw = s * cw - (p * om).sum(axis=1) * sw * winv
        qn = np.concatenate([xyz, qw[:, None]], axis=1)
        qn /= np.linalg.norm(qn, axis=1)[:, None]
        quat = np.where(ok[:, None], qn, quat)
    out = dict(rods)
    out["pos"], out["quat"] = po.wrap_positions(pos, lo, hi), quat
    return out


def cpu_step_sample(po, rods, box, vnc, cpu_iters, nthreads, gpu_iters=None):
    """Bounded CPU sample of the same workload.  Pair search + assembly run in full, the BBPGD loop runs
    `cpu_iters` iterations (or to convergence if that comes first); the per-iteration cost is extrapolated
    to the iteration count the tolerance needs (gpu_iters if known, else the CPU loop is run to the end)."""
    lo, hi, pbc = [0.0] * 3, [box] * 3, [1, 1, 1]
    n = len(rods["gid"])
    t0 = time.perf_counter()
    pos = po.wrap_positions(rods["pos"], lo, hi)
    orods = po.make_rods(rods["gid"], rods["radius"], rods["length"], pos, rods["quat"], 1.0, 1.0, COLBUF)
    t1 = time.perf_counter()
    kind = "port"
    if po.have_ref():
        # the reference's own FDPS tree + functor (oracle/_ref); list differs from P_geo by the known
        # FDPS search-radius quirk, so the solve below is fed the port's full list.
        po.fdps_collect(orods, lo, hi, pbc, nthreads=nthreads, rebuild=True)
        t_collect = po.fdps_last_seconds()
        blocks = po.collect_pairs(orods, lo, hi, pbc, method="cells", nthreads=nthreads)
        kind = "reference"
    else:
        tc = time.perf_counter()
        blocks = po.collect_pairs(orods, lo, hi, pbc, method="cells", nthreads=nthreads)
        t_collect = time.perf_counter() - tc
    full = gpu_iters is None
    sol = po.solve_constraints(blocks, orods, rods["immovable"], MU, vnc, DT, RES, MAXITE if full else cpu_iters, 0,
                               nthreads=nthreads, hist_cap=4)
    done_it = max(sol["nIte"], 1)
    t_iter = sol["tSolve"] / (done_it + 1)
    need = sol["nIte"] if full else gpu_iters
    t_step = (t1 - t0) + t_collect + sol["tAssemble"] + t_iter * (need + 1)
    return dict(t_step=t_step, t_prep=t1 - t0, t_collect=t_collect, t_assemble=sol["tAssemble"], t_iter=t_iter,
                iters_run=sol["nIte"], iters_needed=need, nc=len(blocks), kind=kind)


def main():
    a = parse()
    n = int(a.rods)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = os.cpu_count() or 1
    workload = (f"synthetic {n} random spherocylinders per GPU (L={L_ROD} D={2 * R_ROD} phi={a.phi} periodic cube, "
                f"colBuf={COLBUF} mu={MU} dt={DT} conResTol={RES} BBPGD, {a.relax} untimed relaxation steps, "
                f"Brownian-scale velNonCon seed {SEED})")

    if a.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(a, n, workload, ncores)

    import torch
    import torch.distributed as dist

    import alens_b200

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    stream = torch.cuda.Stream()
    ctx = alens_b200.Context(device=local, rank=rank, nranks=world)
    ctx.set_stream(stream.cuda_stream)

    rods, box = make_workload(n, a.phi, SEED + rank)
    if world > 1:
        # one global suspension: periodic box, `world` boxes long along SLAB_AXIS, slab r owned by rank r (SURVEY.md 8d
        # config 5); ghost rods within cutoff + skin of the slab faces are mirrored between neighbours every step
        rods["pos"][:, SLAB_AXIS] += rank * box
        rods["gid"] = (rods["gid"] + rank * n).astype(np.int32)
        max_r = 0.5 * L_ROD + R_ROD
        skin = 0.5 * (2 * max_r + COLBUF)  # rods drift during the untimed relaxation steps
        ctx.set_domain([0.0] * 3, [world * box if k == SLAB_AXIS else box for k in range(3)], [1, 1, 1])
        ctx.set_collision_params(1.0, 1.0, COLBUF)
        ctx.set_decomposition(SLAB_AXIS, rank * box, (rank + 1) * box, skin, max_r, rank * n)
        ctx.comm_create(int(1.25 * n))
        from alens_b200 import slabs

        ctx.comm_connect(slabs.exchange_blobs(ctx.comm_export()))  # the only host-side collective of the data path
    rods, relax_info = relax_on_gpu(ctx, rods, box, a.relax, configured=world > 1)
    vnc = thermal_velocity(rods, MU, DT, seed=SEED + 17 + rank)

    # pinned host buffers for the e2e leg
    def pin(x):
        t = torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
        return t

    h_gid, h_pos, h_quat = pin(rods["gid"]), pin(rods["pos"]), pin(rods["quat"])
    h_len, h_rad, h_imm, h_vnc = pin(rods["length"]), pin(rods["radius"]), pin(rods["immovable"]), pin(vnc)
    h_out = [torch.empty(6 * n, dtype=torch.float64).pin_memory() for _ in range(4)]

    def upload():
        ctx.set_rods_raw(n, h_gid.data_ptr(), h_pos.data_ptr(), h_quat.data_ptr(), h_len.data_ptr(),
                         h_rad.data_ptr(), h_imm.data_ptr(), wrap=True)

    def step_resident():
        ctx.prepare_step(True)
        nc = ctx.collect_pair_collision()
        ctx.calc_mobility(MU)
        rep = ctx.solve_constraints(None, DT, RES, MAXITE, 0)
        return nc, rep

    def step_e2e():
        upload()
        ctx.set_velocity_noncon_async_raw(h_vnc.data_ptr())  # this step's velNonCon: its H2D overlaps the pair search
        nc = ctx.collect_pair_collision()
        ctx.calc_mobility(MU)
        rep = ctx.solve_constraints_raw(None, DT, RES, MAXITE, 0)
        ctx.get_force_velocity_raw(*[t.data_ptr() for t in h_out])
        return nc, rep

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident leg ----
    upload()
    ctx.set_velocity_noncon(vnc)
    for _ in range(max(a.warmup, 3)):
        nc, rep = step_resident()
    ctx.set_profiling(True)
    ctx.reset_timers()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    phase = dict(upload_ms=0.0, collect_ms=0.0, setup_ms=0.0, solve_ms=0.0, split_ms=0.0)
    for _ in range(a.steps):
        nc, rep = step_resident()
        tm = ctx.get_timers()
        for k in phase:
            phase[k] += tm[k]
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    tm = ctx.get_timers()
    launches = tm["total_launches"]
    clocks = sampler.stop() if sampler else None
    ctx.set_profiling(False)

    # ---- e2e leg ----
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_e2e()
    barrier()
    ms_e2e = (time.perf_counter() - t0) * 1e3

    if world > 1:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant BCQP kernel (algorithmic bytes: DESIGN.md "Kernels") ----
    # Event pairs sit around every 8th BBPGD iteration of every timed step (they break the launch overlap, so not all).
    peak, peak_src = peaks()
    ninc = 2 * nc  # two-sided constraints only in this workload
    # k_force_vel_act reads the id of every incidence slot (4 B) and the 1-bit row mask, and only for the rows that can be
    # non-zero ("live", counted by k_bb_tail: 2 slots per live row) the {x, g} pair (16 B) and the 48 B column record;
    # per rod q + 1/drag (48 B) and the U row (48 B).  k_bb_tail (collision-only pool, no K^-1 term): 2 ids + 9 geometry
    # doubles + {x, g} in + b + flag = 113 B read, {x, g} out = 16 B written per constraint row (+ 1 mask bit).
    live_rows = tm["op_rows_live"] / max(tm["op_applies"], 1)
    dense_force = 52.0 * ninc + 16.0 * nc + 96.0 * n  # what the dense level-major kernel (force_kernel=0) moves
    kern = {
        "k_force_vel_act": (tm["op_force_vel_ms"], tm["op_force_vel_n"],
                            4.0 * ninc + nc / 8.0 + 2.0 * live_rows * (16.0 + 48.0) + 96.0 * n),
        "k_bb_tail": (tm["op_dtrans_ms"], tm["op_dtrans_n"], 129.0 * nc + nc / 8.0),
    }
    dom = max(kern, key=lambda k: kern[k][0])
    t_ms, cnt, bytes_ = kern[dom]
    avg_ms = t_ms / max(cnt, 1)
    achieved = bytes_ / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    roof = {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
            "avg_launch_us": round(avg_ms * 1e3, 2), "algorithmic_bytes_per_launch": int(bytes_),
            "all_kernels": {k: {"avg_us": round(1e3 * v[0] / max(v[1], 1), 2), "launches_timed": int(v[1]),
                                "algorithmic_bytes": int(v[2]),
                                "GBps": round(v[2] / max(1e-12, v[0] / max(v[1], 1) * 1e-3) / 1e9, 1)}
                            for k, v in kern.items()},
            "force_kernel_note": {"live_rows_per_apply": int(live_rows), "live_fraction": round(live_rows / max(nc, 1), 4),
                                  "dense_equivalent_bytes": int(dense_force),
                                  "dense_equivalent_GBps": round(dense_force / max(1e-12, kern["k_force_vel_act"][0] /
                                                                 max(kern["k_force_vel_act"][1], 1) * 1e-3) / 1e9, 1)}}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            with open(prof) as f:
                roof["traffic"] = json.load(f).get(dom)
        except Exception:
            pass

    # ---- cpu baseline (bounded sample, N=1 only) ----
    cpu = None
    if not a.no_cpu and world == 1:
        try:
            from oracle import pyoracle as po

            po.lib()
            s = cpu_step_sample(po, rods, box, vnc, a.cpu_iters, ncores, gpu_iters=rep.iterations)
            cpu = {"value": round(1.0 / s["t_step"], 5), "unit": "steps/s", "cores": ncores, "kind": s["kind"],
                   "sample": (f"pair search ({'FDPS reference build' if s['kind'] == 'reference' else 'oracle cell list'}"
                              f" {s['t_collect']:.2f}s) + CSR assembly ({s['t_assemble']:.2f}s) in full, "
                              f"{s['iters_run']} BBPGD iterations timed ({s['t_iter'] * 1e3:.1f} ms/iter) and "
                              f"extrapolated to the {s['iters_needed']} iterations the tolerance needs; "
                              f"{s['nc']} constraints, OpenMP {ncores} threads")}
        except Exception as ex:  # the baseline must never take the bench line down
            cpu = {"value": None, "unit": "steps/s", "cores": ncores, "kind": "port", "sample": f"failed: {ex!r}"}

    steps_total = a.steps * world
    line = {
        "metric": "constraint-solve steps/sec at 1M rods (1/2/4/8 B200); kernel HBM GB/s vs peak",
        "value": round(steps_total / (ms * 1e-3), 3), "unit": "steps/s", "n_gpus": world, "steps": a.steps,
        "warmup": max(a.warmup, 3), "ms_per_step": round(ms / a.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "rods_per_gpu": n, "constraints": int(nc),
                   "bbpgd_iterations": int(rep.iterations), "residual": float(rep.residual),
                   "parallelism": "1 process per GPU" + ("" if world == 1 else (
                       f", one periodic suspension of {world * n} rods in {world} {'xyz'[SLAB_AXIS]}-slabs: ghost-rod exchange per step, "
                       "U halo + 4-double allreduce per BBPGD iteration over NVLink peer memory; value = slab-steps/s "
                       "(global steps/s x GPUs)")),
                   "ghosts_rank0": ctx.num_ghosts() if world > 1 else None,
                   "l2": "inputs larger than L2 (constraint + incidence arrays > 600 MB)",
                   "relaxation": [list(map(int, x)) for x in relax_info],
                   "phase_ms_per_step": {k: round(v / a.steps, 3) for k, v in phase.items()}},
        "e2e": {"value": round(steps_total / (ms_e2e * 1e-3), 3), "unit": "steps/s",
                "h2d_bytes_per_step": int(n * (4 + 24 + 32 + 8 + 8 + 1 + 48)), "d2h_bytes_per_step": int(n * 4 * 48),
                "ms_per_step": round(ms_e2e / a.steps, 3)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def reference_arm(a, n, workload, ncores):
    """The reference's CPU algorithm for the same path on the host cores: FDPS pair search from the
    reference's own sources when oracle/_ref was built, CSR D/M + BBPGD as the reference structures them."""
    from oracle import pyoracle as po

    po.lib()
    rods, box = make_workload(n, a.phi, SEED)
    prepared = "cpu relaxation (oracle)"
    try:
        import alens_b200

        ctx = alens_b200.Context(0)
        rods, _ = relax_on_gpu(ctx, rods, box, a.relax)  # untimed input preparation only
        ctx.close()
        prepared = "gpu relaxation (untimed input preparation)"
    except Exception:
        rods = relax_on_cpu(po, rods, box, a.relax, ncores)
    vnc = thermal_velocity(rods, MU, DT, seed=SEED + 17)
    steps = max(1, min(a.steps, 3))
    for _ in range(min(a.warmup, 1)):
        cpu_step_sample(po, rods, box, vnc, a.cpu_iters, ncores, gpu_iters=None)
    ts, last = [], None
    for _ in range(steps):
        last = cpu_step_sample(po, rods, box, vnc, a.cpu_iters, ncores, gpu_iters=None)
        ts.append(last["t_step"])
    t = float(np.mean(ts))
    val = round(1.0 / t, 5)
    sample = (f"{steps} full steps: pair search ({'FDPS reference build' if last['kind'] == 'reference' else 'oracle cell list'} "
              f"{last['t_collect']:.2f}s) + CSR assembly {last['t_assemble']:.2f}s + {last['iters_run']} BBPGD "
              f"iterations to conResTol ({last['t_iter'] * 1e3:.1f} ms/iter); {last['nc']} constraints")
    line = {
        "impl": "reference",
        "metric": "constraint-solve steps/sec at 1M rods (1/2/4/8 B200); kernel HBM GB/s vs peak",
        "value": val, "unit": "steps/s", "n_gpus": a.gpus, "steps": steps, "warmup": min(a.warmup, 1),
        "ms_per_step": round(t * 1e3, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "rods_per_gpu": n, "input": prepared,
                   "note": "Trilinos/Tpetra cannot be built here: D/M as hand CSR, FDPS + DCPQuery from the reference"},
        "cpu_baseline": {"value": val, "unit": "steps/s", "cores": ncores, "kind": last["kind"], "sample": sample},
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
