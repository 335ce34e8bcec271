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
           D2H of the force/velocity result inside the timed region.
`roofline`: dominant BCQP kernel, algorithmic bytes / CUDA-event duration vs MEASURED_PEAKS.json;
           `traffic` = DRAM bytes per launch from the newest profiles/r*_ncu_full*.csv.
`cpu_baseline` / `--impl reference`: the reference's OWN pipeline on the host cores -- SylinderSystem::prepareStep,
           calcVelocityNonCon, resolveConstraints (FDPS search, functor, D^T assembly, ConstraintSolver, BCQPSolver)
           compiled unmodified into oracle/_ref/libalens_refsys.so on a stand-in for the Tpetra/Eigen/MPI containers
           (oracle/stubs); falls back to the C port (oracle/liboracle.so) where that library is absent.
`parity` : N = 1: the reference's pair list is a bit-identical sub-list of ours, and gamma / velocities of a fixed number
           of BBPGD iterations on OUR list agree with the reference's solver to 1e-8; N > 1: the device-side digests of
           the ranks (pair list exact, gamma sums 1e-8) add up to the digest of the same suspension run on one GPU, and
           the fused protocol (what is timed) equals the unfused one bit for bit.
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
# multi-GPU runs: slabs along z, the slowest axis of the cell order -- the rods a neighbour mirrors and the rows that read
# ghost velocities are then contiguous at the two ends of the sorted arrays, so only ~10 % of the force kernel's CTAs take
# part in the halo release and the tail kernel keeps ~10 % of its tiles for after the halo wait (with x-slabs every tile
# contains boundary rods; profiles/README.md).  ALENS_SLAB_AXIS=0/1/2 overrides.
SLAB_AXIS = int(os.environ.get("ALENS_SLAB_AXIS", "2"))
SEED = 1234


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--rods", type=float, default=1e6, help="rods per GPU (weak scaling) / in total (strong scaling)")
    p.add_argument("--phi", type=float, default=None, help="volume fraction (default 0.10 for S1, 0.40 for S2)")
    p.add_argument("--workload", default="S1", choices=["S1", "S2"],
                   help="S1: uniform positions, isotropic orientations (SURVEY 8d config 4 'reference-style'); "
                        "S2: jittered hexagonal lattice of z-aligned rods ('dense nematic', single GPU)")
    p.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                   help="weak: --rods per GPU, one box per GPU; strong: --rods in total, cut into N slabs")
    p.add_argument("--relax", type=int, default=4, help="untimed relaxation steps before measuring")
    p.add_argument("--cpu-iters", type=int, default=12, help="BBPGD iterations of the fixed-count parity solve")
    p.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (and the N = 1 parity check in it)")
    p.add_argument("--no-parity", action="store_true", help="skip the N > 1 digest check")
    p.add_argument("--stamps", action="store_true", default=True,
                   help="one extra (untimed) solve with per-iteration nanosecond stamps (default: on)")
    p.add_argument("--no-stamps", dest="stamps", action="store_false")
    a = p.parse_args()
    if a.phi is None:
        a.phi = 0.40 if a.workload == "S2" else 0.10
    return a


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


LENGTH_SIGMA = float(os.environ.get("ALENS_LENGTH_SIGMA", "0"))  # polydisperse variant of S1 (not the headline workload)


def make_workload(n, phi, seed, kind="S1"):
    """returns rods and the box edge lengths (3-vector, low corner at the origin)"""
    if kind == "S2":
        return dense_nematic(n, phi, seed)
    box = box_for_volume_fraction(n, L_ROD, R_ROD, phi)
    rods = random_rods(n, box, L_ROD, R_ROD, seed=seed)
    if LENGTH_SIGMA > 0:  # log-normal lengths with the same mean (setInitialFromConfig draws them like this, :228-232)
        rng = np.random.default_rng(seed + 99)
        L = L_ROD * np.exp(LENGTH_SIGMA * rng.normal(size=n) - 0.5 * LENGTH_SIGMA**2)
        rods["length"] = np.minimum(L, 0.4 * box)
    return rods, np.full(3, box)


def dense_nematic(n, phi, seed, sigma_theta=0.1, jitter=0.2):
    """SURVEY 8d config 4 'S2': jittered hexagonal lattice of z-aligned rods (sigma_theta rad, positional jitter 0.2 D) at
    volume fraction phi.  Lattice constant 1.3 D in the plane, layer height from phi; the box is commensurate with the
    lattice (periodic), surplus sites are removed at random so that exactly n rods remain."""
    from scenarios import quat_from_z_to

    rng = np.random.default_rng(seed)
    D = 2 * R_ROD
    vol = np.pi * R_ROD**2 * L_ROD + 4.0 / 3.0 * np.pi * R_ROD**3
    a = 1.3 * D
    h = vol / phi / (np.sqrt(3.0) / 2.0 * a * a)
    edge = (n * vol / phi) ** (1.0 / 3.0)
    nz = max(1, int(round(edge / h)))
    per = n / nz
    nx = max(1, int(round(np.sqrt(per * (np.sqrt(3.0) / 2.0)))))
    ny = int(np.ceil(per / nx))
    ny += ny & 1  # even number of rows: the staggered lattice closes periodically
    while nx * ny * nz < n:
        nz += 1
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    x = (ix + 0.5 * (iy & 1)) * a
    y = iy * (np.sqrt(3.0) / 2.0 * a)
    z = (iz + 0.5) * h
    pos = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
    keep = rng.permutation(len(pos))[:n]
    pos = pos[np.sort(keep)] + rng.normal(0.0, jitter * D, size=(n, 3))
    d = np.zeros((n, 3))
    d[:, 2] = 1.0
    d[:, :2] = rng.normal(0.0, sigma_theta, size=(n, 2))
    rods = dict(gid=rng.permutation(n).astype(np.int32), pos=pos, quat=quat_from_z_to(d), length=np.full(n, L_ROD),
                radius=np.full(n, R_ROD), immovable=np.zeros(n, dtype=np.uint8))
    return rods, np.array([nx * a, ny * np.sqrt(3.0) / 2.0 * a, nz * h])


def relax_on_gpu(ctx, rods, box, steps, configured=False):
    """untimed: the reference's own initPreSteps loop (SylinderSystem.cpp:88-101) run on the device"""
    if not configured:
        ctx.set_domain([0.0] * 3, list(np.broadcast_to(box, (3,))), [1, 1, 1])
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
    """same loop with the oracle port (only used by --impl reference when oracle/_ref is absent)"""
    box = np.broadcast_to(box, (3,))
    lo, hi, pbc = [0.0] * 3, list(box), [1, 1, 1]
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


def port_step(po, rods, box, vnc, nthreads):
    """one full step of the C port (oracle/liboracle.so): cell-list pair search, CSR D^T / D / M, unfused BBPGD"""
    box = np.broadcast_to(box, (3,))
    lo, hi, pbc = [0.0] * 3, list(box), [1, 1, 1]
    t0 = time.perf_counter()
    pos = po.wrap_positions(rods["pos"], lo, hi)
    orods = po.make_rods(rods["gid"], rods["radius"], rods["length"], pos, rods["quat"], 1.0, 1.0, COLBUF)
    blocks = po.collect_pairs(orods, lo, hi, pbc, method="cells", nthreads=nthreads)
    sol = po.solve_constraints(blocks, orods, rods["immovable"], MU, vnc, DT, RES, MAXITE, 0, nthreads=nthreads, hist_cap=4)
    return dict(t_step=time.perf_counter() - t0, nc=len(blocks), iters=sol["nIte"])


def reference_system(pr, box, nthreads, relax=0, n_init=0):
    """SylinderSystem of the reference with the bench's RunConfig (one rank, all host threads)"""
    box = np.broadcast_to(box, (3,))
    cfg = dict(simBoxLow=[0.0] * 3, simBoxHigh=[float(x) for x in box], simBoxPBC=[True] * 3, initBoxLow=[0.0] * 3,
               initBoxHigh=[float(x) for x in box], sylinderColBuf=COLBUF, sylinderDiameterColRatio=1.0,
               sylinderLengthColRatio=1.0, viscosity=MU, KBT=-1.0, dt=DT, conResTol=RES, conMaxIte=MAXITE,
               conSolverChoice=0, initPreSteps=relax, sylinderNumber=n_init, sylinderLength=L_ROD,
               sylinderDiameter=2 * R_ROD, rngSeed=SEED, logLevel=5, timerLevel=5)
    return pr.RefSystem(cfg, nthreads=nthreads)


def reference_step(sysr, vnc):
    """the reference's timestep up to the solve: prepareStep (applyBoxBC, decomposition, mobility matrix), the given
    non-Brownian velocity through calcVelocityNonCon, resolveConstraints (collect + setup + BCQP + split + write-back)"""
    t0 = time.perf_counter()
    sysr.prepare_step()
    t1 = time.perf_counter()
    sysr.set_velocity_nonbrown(vnc)
    sysr.calc_velocity_noncon()
    t2 = time.perf_counter()
    sysr.resolve_constraints()
    t3 = time.perf_counter()
    return dict(t_step=t3 - t0, t_prepare=t1 - t0, t_velocity=t2 - t1, t_resolve=t3 - t2)


def key_of(blocks):
    """sortable identity of a block: (gidI, gidJ, labJ bits) -- a pair may appear once per periodic image"""
    k = np.empty((len(blocks), 5), dtype=np.int64)
    k[:, 0], k[:, 1] = blocks["gidI"], blocks["gidJ"]
    k[:, 2:] = np.ascontiguousarray(blocks["labJ"]).view(np.int64).reshape(-1, 3)
    return k


def check_parity_n1(ctx, pr, sysr, vnc, iters):
    """N = 1, untimed: (1) every block of the reference's own collection is in our list with bit-identical delta0 /
    normal / contact points; (2) `iters` BBPGD iterations on OUR list by our solver and by the reference's
    ConstraintSolver + BCQPSolver: gamma and rod velocities to 1e-8 relative (BASELINE.json north_star)."""
    ref = sysr.constraints()
    ours = ctx.get_constraints(with_stress=False)
    ko, kr = key_of(ours), key_of(ref)
    order = np.lexsort(ko.T[::-1])
    ks = ko[order]
    # position of every reference key in our sorted key list
    view = lambda k: np.ascontiguousarray(k).view([("", k.dtype)] * k.shape[1]).ravel()
    pos = np.searchsorted(view(ks), view(kr))
    pos = np.minimum(pos, len(ks) - 1)
    found = (ks[pos] == kr).all(axis=1)
    sel = ours[order[pos]]
    same = bool(found.all())
    for f in ("delta0", "normI", "posI", "posJ", "labI"):
        same = same and bool(np.array_equal(sel[f][found], ref[f][found]))
    res = {"pair_list": ("ok" if same else "MISMATCH"), "ref_pairs": int(len(ref)), "our_pairs": int(len(ours)),
           "ref_pairs_missing_from_ours": int((~found).sum())}
    rep = ctx.solve_constraints(None, DT, 1e-30, iters, 0)
    g = ctx.get_gamma()
    out = ctx.get_force_velocity()
    r = sysr.solve_blocks(ours, vnc, DT, 1e-30, iters, 0, hist_cap=0)
    eg = float(np.abs(g - r["gamma"]).max() / max(np.abs(r["gamma"]).max(), 1e-300))
    ev = float(np.abs(out["velU"] - r["velU"]).max() / max(np.abs(r["velU"]).max(), 1e-300))
    res.update({"bbpgd_iterations_compared": int(rep.iterations), "gamma_rel_err": eg, "velocity_rel_err": ev,
                "solver": "ok" if (eg < 1e-8 and ev < 1e-8 and rep.iterations == iters) else "MISMATCH"})
    res["status"] = "ok" if res["pair_list"] == "ok" and res["solver"] == "ok" else "MISMATCH"
    return res


def bind_near_gpu(local):
    """N > 1: keep this rank's host threads -- and with them the pinned buffers they first touch -- on the CPUs next to its
    GPU (NVML's affinity mask), so that the per-step host<->device copies of 8 ranks do not all cross the socket link.
    Returns a description for the bench line, or None when the box gives no usable mask."""
    try:
        import pynvml

        pynvml.nvmlInit()
        idx = local
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        if vis and all(t.strip().isdigit() for t in vis.split(",")) and local < len(vis.split(",")):
            idx = int(vis.split(",")[local])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        near = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        use = near & allowed
        if len(use) < 4 or len(use) == len(allowed):
            return None
        os.sched_setaffinity(0, use)
        return f"{len(use)} of {len(allowed)} cpus (NVML affinity of GPU {idx})"
    except Exception:  # noqa: BLE001 -- placement is an optimisation, never a reason to fail
        return None


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = os.cpu_count() or 1
    strong = a.scaling == "strong" and world > 1
    n_total = int(a.rods) if (strong or world == 1) else int(a.rods) * world
    dist_name = ("uniform positions + isotropic orientations in a periodic cube" if a.workload == "S1" else
                 "jittered hexagonal lattice of z-aligned rods (sigma_theta 0.1 rad, jitter 0.2 D), periodic box")
    workload = (f"synthetic {int(a.rods)} random spherocylinders {'in total' if strong else 'per GPU'} ({a.workload}: "
                f"{dist_name}, L={L_ROD} D={2 * R_ROD} phi={a.phi}, colBuf={COLBUF} mu={MU} dt={DT} conResTol={RES} BBPGD, "
                f"{a.relax} untimed relaxation steps, Brownian-scale velNonCon seed {SEED})")

    if a.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(a, int(a.rods), workload, ncores)

    # (N = 1 stays unbound: its cpu_baseline leg gives the reference every host core)
    host_affinity = bind_near_gpu(local) if world > 1 and os.environ.get("ALENS_NO_BIND") != "1" else None

    import torch
    import torch.distributed as dist

    import alens_b200

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    stream = torch.cuda.Stream()
    ctx = alens_b200.Context(device=local, rank=rank, nranks=world)
    ctx.set_stream(stream.cuda_stream)
    # tuning knobs for A/B measurements: ALENS_OPTIONS="force_kernel=1,rec_mode=0" (only rec_mode 2 against the others changes
    # result bits: it sums M * column instead of applying M to the summed force)
    opts = dict(kv.split("=") for kv in os.environ.get("ALENS_OPTIONS", "").split(",") if "=" in kv)
    for k, v in opts.items():
        ctx.set_option(k, int(v))
    fk, rec_mode = int(opts.get("force_kernel", 3)), int(opts.get("rec_mode", 2))

    max_r = 0.5 * L_ROD + R_ROD
    skin = 0.5 * (2 * max_r + COLBUF)  # rods drift during the untimed relaxation steps
    if strong:
        # ONE suspension of --rods rods cut into `world` slabs: every rank generates the same rods and keeps its slab
        allrods, box = make_workload(n_total, a.phi, SEED, a.workload)
        gbox = box.copy()
        w = gbox[SLAB_AXIS] / world
        own = np.minimum(np.floor(allrods["pos"][:, SLAB_AXIS] / w).astype(int), world - 1) == rank
        rods = {k: v[own] for k, v in allrods.items()}
        slab = (rank * w, (rank + 1) * w)
        counts = [int((np.minimum(np.floor(allrods["pos"][:, SLAB_AXIS] / w).astype(int), world - 1) == r).sum())
                  for r in range(world)]
        base = int(sum(counts[:rank]))
        del allrods
    else:
        rods, box = make_workload(int(a.rods), a.phi, SEED + rank, a.workload)
        gbox = box.copy()
        gbox[SLAB_AXIS] *= world
        rods["pos"][:, SLAB_AXIS] += rank * box[SLAB_AXIS]
        rods["gid"] = (rods["gid"] + rank * int(a.rods)).astype(np.int32)
        slab = (rank * box[SLAB_AXIS], (rank + 1) * box[SLAB_AXIS])
        base = rank * int(a.rods)
        counts = [int(a.rods)] * world
    n = len(rods["gid"])
    if world > 1:
        # one global periodic suspension in `world` slabs along SLAB_AXIS; ghost rods within
        # cutoff + skin of the slab faces are mirrored between neighbours every step
        ctx.set_domain([0.0] * 3, list(gbox), [1, 1, 1])
        ctx.set_collision_params(1.0, 1.0, COLBUF)
        ctx.set_decomposition(SLAB_AXIS, slab[0], slab[1], skin, max_r, base)
        ctx.comm_create(int(1.25 * max(counts)) + 4096)
        from alens_b200 import slabs

        ctx.comm_connect(slabs.exchange_blobs(ctx.comm_export()))  # the only host-side collective of the data path
    rods, relax_info = relax_on_gpu(ctx, rods, gbox, a.relax, configured=world > 1)
    vnc = thermal_velocity(rods, MU, DT, seed=SEED + 17 + rank)
    # timing experiments that break results (never set for a reported number): applied after the relaxation steps
    for kv in os.environ.get("ALENS_LATE_OPTIONS", "").split(","):
        if "=" in kv:
            ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))

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

    # what saveForceVelocityConstraints needs back: force and velocity of the unilateral part; the bilateral arrays of a
    # collision-only pool are identically zero and are not transferred (the call reports that to the caller)
    e2e_out = [h_out[0].data_ptr(), h_out[1].data_ptr(), 0, 0]

    def step_e2e():
        # this step's inputs: positions, orientations and velNonCon.  gids, lengths, radii and immovable flags do not change
        # from step to step and stay resident (alens_set_rod_state is alens_set_rods without them)
        ctx.set_rod_state_raw(h_pos.data_ptr(), h_quat.data_ptr(), wrap=True)
        ctx.set_velocity_noncon_async_raw(h_vnc.data_ptr())  # this step's velNonCon: its H2D overlaps the pair search
        nc = ctx.collect_pair_collision()
        ctx.calc_mobility(MU)
        rep = ctx.solve_constraints_raw(None, DT, RES, MAXITE, 0)
        ctx.get_force_velocity_raw(*e2e_out)
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
    under_ncu = bool(os.environ.get("ALENS_NCU"))  # `ncu --profile-from-start off`: the launch list of exactly the timed region
    if under_ncu:
        torch.cuda.profiler.start()
    e0.record(stream)
    phase = dict(upload_ms=0.0, collect_ms=0.0, setup_ms=0.0, solve_ms=0.0, split_ms=0.0)
    for _ in range(a.steps):
        nc, rep = step_resident()
        tm = ctx.get_timers()
        for k in phase:
            phase[k] += tm[k]
    e1.record(stream)
    barrier()
    if under_ncu:
        torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    tm = ctx.get_timers()
    launches = tm["total_launches"]
    clocks = sampler.stop() if sampler else None
    ctx.set_profiling(False)
    live = ctx.get_live_stats()
    stats = ctx.get_collect_stats() if hasattr(ctx, "get_collect_stats") else None

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

    # ---- where an iteration's time goes (untimed extra solve with %globaltimer stamps inside the two kernels) ----
    breakdown = None
    if a.stamps:
        ctx.set_option("stamps", 1)
        step_resident()
        ctx.set_option("stamps", 0)
        st = ctx.get_stamps().astype(np.float64)
        if len(st) > 4:
            st = st[2:]  # steady state
            iv = {"force_us": st[:, 2] - st[:, 0], "force_to_halo_release_us": st[:, 1] - st[:, 0],
                  "gap_force_to_tail_us": st[:, 3] - st[:, 2], "tail_rows_us": st[:, 6] - st[:, 3],
                  "tail_halo_wait_us": np.maximum(st[:, 5] - st[:, 4], 0), "tail_start_to_halo_wait_us": st[:, 4] - st[:, 3],
                  "allreduce_us": st[:, 7] - st[:, 6], "gap_tail_to_next_force_us": np.append(st[1:, 0] - st[:-1, 7], np.nan),
                  "iteration_us": np.append(st[1:, 0] - st[:-1, 0], np.nan)}
            if world == 1:
                iv.pop("force_to_halo_release_us"); iv.pop("tail_halo_wait_us"); iv.pop("tail_start_to_halo_wait_us")
            mine = np.array([float(np.nanmean(v)) * 1e-3 for v in iv.values()])
            if world > 1:
                tt = torch.from_numpy(mine).cuda()
                allr = [torch.zeros_like(tt) for _ in range(world)]
                dist.all_gather(allr, tt)
                per_rank = np.stack([x.cpu().numpy() for x in allr])
            else:
                per_rank = mine[None]
            breakdown = {k: {"rank0": round(float(per_rank[0, i]), 2), "max_over_ranks": round(float(per_rank[:, i].max()), 2),
                             "min_over_ranks": round(float(per_rank[:, i].min()), 2)} for i, k in enumerate(iv)}

    # ---- N > 1 parity (untimed): digests of the ranks against the same suspension on ONE GPU, fused against unfused ----
    parity = None
    if world > 1 and not a.no_parity:
        parity = check_parity_multi(ctx, dist, torch, alens_b200, rods, gbox, vnc, rank, world, local, a.cpu_iters)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant BCQP kernel (algorithmic bytes: DESIGN.md "Kernels") ----
    # Event pairs sit around every 8th BBPGD iteration of every timed step (they break the launch overlap, so not all).
    peak, peak_src = peaks()
    ninc = 2 * nc  # two-sided constraints only in this workload
    # k_force_vel_rec: slot range per rod (4 B), the live slot bitmap (1 bit per slot), ONE 64-byte record per live slot
    # (row id + column block) and the {x, g} pair of its row (16 B), q + 1/drag (48 B) of the rods that have a live slot,
    # ghost flag (1 B) and the U row (48 B) of every rod.  k_bb_tail (collision-only pool, no K^-1 term): 2 ids + 9
    # geometry doubles + {x, g} in + b + flag = 113 B read and {x, g} out = 16 B written per constraint row, mask word in
    # and out (2 bits); the slot-bitmap bits of the few rows whose liveness flips are negligible.
    live_rows = tm["op_rows_live"] / max(tm["op_applies"], 1)
    dense_force = 52.0 * ninc + 16.0 * nc + 96.0 * n  # what the dense level-major kernel (force_kernel=0) moves
    if fk == 3 and rec_mode == 2:
        # records hold M * column: no mobility line; rod header (8 B) + end of the slot range (4 B) per rod instead of the
        # slot bitmap; 64-byte record + 16-byte {x, g} pair per live slot; ghost flag (1 B) and U row (48 B) per rod
        kern = {"k_force_vel_rec": (tm["op_force_vel_ms"], tm["op_force_vel_n"],
                                    12.0 * n + 80.0 * live["live_slots"] + 49.0 * n),
                "k_bb_tail": (tm["op_dtrans_ms"], tm["op_dtrans_n"], 129.0 * nc + nc / 8.0)}
    elif fk == 3 and rec_mode == 1:
        # records hold the row id: + one 16-byte {x, g} gather per live slot; the tail only flips bitmap bits
        kern = {"k_force_vel_rec": (tm["op_force_vel_ms"], tm["op_force_vel_n"],
                                    4.0 * (n + 1) + ninc / 8.0 + 80.0 * live["live_slots"] + 48.0 * live["live_rods"] + 49.0 * n),
                "k_bb_tail": (tm["op_dtrans_ms"], tm["op_dtrans_n"], 129.0 * nc + nc / 4.0)}
    elif fk == 3:
        kern = {"k_force_vel_rec": (tm["op_force_vel_ms"], tm["op_force_vel_n"],
                                    4.0 * (n + 1) + 2.0 * ninc / 8.0 + 64.0 * live["live_slots"] + 48.0 * live["live_rods"] + 49.0 * n),
                "k_bb_tail": (tm["op_dtrans_ms"], tm["op_dtrans_n"], 137.0 * nc + nc / 4.0 + 64.0 * live_rows)}
    else:  # k_force_vel_act: every slot id + row mask bit, {x, g} + 48-byte column record per live slot, rod data, U
        kern = {"k_force_vel_act": (tm["op_force_vel_ms"], tm["op_force_vel_n"],
                                    4.0 * ninc + nc / 8.0 + 2.0 * live_rows * (16.0 + 48.0) + 96.0 * n),
                "k_bb_tail": (tm["op_dtrans_ms"], tm["op_dtrans_n"], 129.0 * nc + nc / 8.0)}
    dom = max(kern, key=lambda k: kern[k][0])
    t_ms, cnt, bytes_ = kern[dom]
    avg_ms = t_ms / max(cnt, 1)
    achieved = bytes_ / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    traffic, traffic_src = ncu_traffic(dom)
    roof = {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
            "avg_launch_us": round(avg_ms * 1e3, 2), "algorithmic_bytes_per_launch": int(bytes_),
            "all_kernels": {k: {"avg_us": round(1e3 * v[0] / max(v[1], 1), 2), "launches_timed": int(v[1]),
                                "algorithmic_bytes": int(v[2]),
                                "GBps": round(v[2] / max(1e-12, v[0] / max(v[1], 1) * 1e-3) / 1e9, 1),
                                "frac": round(v[2] / max(1e-12, v[0] / max(v[1], 1) * 1e-3) / 1e9 / peak, 4)}
                            for k, v in kern.items()},
            "bbpgd_iteration": {"avg_us": round(1e3 * sum(v[0] / max(v[1], 1) for v in kern.values()), 2),
                                "algorithmic_bytes": int(sum(v[2] for v in kern.values())),
                                "frac": round(sum(v[2] for v in kern.values()) / max(1e-12, sum(
                                    v[0] / max(v[1], 1) for v in kern.values()) * 1e-3) / 1e9 / peak, 4)},
            "force_kernel_note": {"live_rows_per_apply": int(live_rows), "live_fraction": round(live_rows / max(nc, 1), 4),
                                  "live_slots": int(live["live_slots"]), "live_rods": int(live["live_rods"]),
                                  "dense_equivalent_bytes": int(dense_force)}}
    if stats:  # pair search: fp64 work of the narrow phase (SURVEY 8d asks for GFLOP/s next to GB/s)
        # exact closest-point query + contact assembly: ~230 fp64 operations per candidate that reaches the narrow phase
        # (counted from the SASS of k_cand_narrow: DADD/DMUL/DFMA-free build, 2 divisions, 1 square root)
        roof["pair_search"] = {"cells": stats["cells"], "narrow_phase_candidates": stats["candidates"], "contacts": stats["hits"],
                               "collect_ms": round(phase["collect_ms"] / a.steps, 3),
                               "fp64_gflops": round(230.0 * stats["candidates"] / max(1e-9, phase["collect_ms"] / a.steps * 1e-3) / 1e9, 1)}

    # ---- cpu baseline = the reference's own pipeline on this box's host cores, one full step (N = 1 only) ----
    cpu = None
    if not a.no_cpu and world == 1:
        cpu, par = cpu_baseline_and_parity(ctx, rods, gbox, vnc, ncores, a.cpu_iters)
        parity = par

    steps_total = a.steps * (1 if strong else world)
    line = {
        "metric": "constraint-solve steps/sec at 1M rods (1/2/4/8 B200); kernel HBM GB/s vs peak",
        "value": round(steps_total / (ms * 1e-3), 3), "unit": "steps/s", "n_gpus": world, "steps": a.steps,
        "warmup": max(a.warmup, 3), "ms_per_step": round(ms / a.steps, 3), "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "rods_per_gpu": n if not strong else counts, "rods_total": n_total,
                   "constraints": int(nc), "bbpgd_iterations": int(rep.iterations), "residual": float(rep.residual),
                   "parallelism": "1 process per GPU" + ("" if world == 1 else (
                       f", one periodic suspension of {n_total} rods in {world} {'xyz'[SLAB_AXIS]}-slabs: ghost-rod exchange per step, "
                       "U halo + 4-double allreduce per BBPGD iteration over NVLink peer memory; value = " +
                       ("global steps/s" if strong else "slab-steps/s (global steps/s x GPUs)"))),
                   "ghosts_rank0": ctx.num_ghosts() if world > 1 else None,
                   "host_affinity_rank0": host_affinity,
                   "l2": "inputs larger than L2 (constraint + incidence arrays > 600 MB)",
                   "relaxation": [list(map(int, x)) for x in relax_info],
                   "phase_ms_per_step": {k: round(v / a.steps, 3) for k, v in phase.items()}},
        "e2e": {"value": round(steps_total / (ms_e2e * 1e-3), 3), "unit": "steps/s",
                "h2d_bytes_per_step": int(n * (24 + 32 + 48)), "d2h_bytes_per_step": int(n * 2 * 48),
                "ms_per_step": round(ms_e2e / a.steps, 3),
                "note": "up: positions, orientations, velNonCon of the step (gid / length / radius / immovable flag stay resident); down: forceUni + velUni (the bilateral arrays are identically zero for a collision-only pool)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": cpu,
        "parity": parity,
    }
    if breakdown:
        line["iteration_breakdown_us"] = breakdown
        # the same algorithmic bytes over the kernels' durations INSIDE the running loop (first to last %globaltimer stamp of a
        # launch, programmatic dependent launch and warm L2 as in production) next to the event-timed isolated launches above
        inl = {}
        for kname, key in (("k_force_vel_rec", "force_us"), ("k_force_vel_act", "force_us"), ("k_bb_tail", "tail_rows_us")):
            if kname in kern and key in breakdown and breakdown[key]["max_over_ranks"] > 0:
                us = breakdown[key]["max_over_ranks"]
                gbps = kern[kname][2] / (us * 1e-6) / 1e9
                inl[kname] = {"us": us, "GBps": round(gbps, 1), "frac": round(gbps / peak, 4)}
        line["roofline"]["in_loop"] = inl
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the newest profiles/r*_ncu_full*.csv (condensed ncu --set full capture,
    tools/ncu_summary.py): mean of dram__bytes_read.sum + dram__bytes_write.sum over its launches; None if absent"""
    import csv
    import glob

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_full*.csv")), key=os.path.getmtime, reverse=True)
    files.sort(key=lambda f: os.path.basename(f).split("_")[0], reverse=True)  # newest round first
    for f in files:
        try:
            with open(f, newline="") as fh:
                rows = list(csv.reader(fh))
            H, units = rows[0], rows[1]
            ir, iw = H.index("dram__bytes_read.sum"), H.index("dram__bytes_write.sum")
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            vals = [float(r[ir]) * scale.get(units[ir], 1.0) + float(r[iw]) * scale.get(units[iw], 1.0)
                    for r in rows[2:] if kernel + "<" in r[0] or r[0].startswith(kernel + "(") or ("::" + kernel) in r[0] or
                    r[0].split("(")[0].split("<")[0].strip().endswith(kernel)]
            if vals:
                return int(sum(vals) / len(vals)), os.path.relpath(f, ROOT) + f" ({len(vals)} launches, cold caches)"
        except Exception:
            continue
    return None, None


def cpu_baseline_and_parity(ctx, rods, box, vnc, ncores, iters):
    """one full step of the reference's own pipeline on the host (all cores) on the GPU arm's relaxed state, plus the
    N = 1 parity check against it; the C port stands in when oracle/_ref is not there"""
    try:
        from oracle import pyrefsys as pr

        if pr.available():
            sysr = reference_system(pr, box, ncores)
            sysr.set_rods(rods)
            t = reference_step(sysr, vnc)
            nref = len(sysr.constraints())
            cpu = {"value": round(1.0 / t["t_step"], 5), "unit": "steps/s", "cores": ncores, "kind": "reference",
                   "sample": (f"1 full step of the reference's own SylinderSystem (prepareStep {t['t_prepare']:.2f}s, "
                              f"calcVelocityNonCon {t['t_velocity']:.2f}s, resolveConstraints {t['t_resolve']:.2f}s: FDPS "
                              f"collection of {nref} constraints + ConstraintSolver/BCQPSolver to conResTol), sources "
                              f"compiled unmodified on a stand-in for Tpetra/Eigen/MPI (oracle/stubs), OpenMP {ncores} threads")}
            try:
                par = check_parity_n1(ctx, pr, sysr, vnc, iters)
            except Exception as ex:  # noqa: BLE001
                par = {"status": f"failed: {ex!r}"}
            sysr.close()
            return cpu, par
        from oracle import pyoracle as po

        po.lib()
        s = port_step(po, rods, box, vnc, ncores)
        return ({"value": round(1.0 / s["t_step"], 5), "unit": "steps/s", "cores": ncores, "kind": "port",
                 "sample": f"1 full step of the C port (oracle/liboracle.so): {s['nc']} constraints, {s['iters']} BBPGD iterations"},
                {"status": "unchecked (oracle/_ref absent)"})
    except Exception as ex:  # the baseline must never take the bench line down
        return {"value": None, "unit": "steps/s", "cores": ncores, "kind": "port", "sample": f"failed: {ex!r}"}, None


def check_parity_multi(ctx, dist, torch, alens_b200, rods, gbox, vnc, rank, world, local, iters):
    """untimed.  (1) `iters` BBPGD iterations with the fused kernels (what was timed) and with the unfused protocol:
    list and gamma digests must be bit-identical.  (2) the same global suspension on ONE GPU (rank 0, no decomposition):
    the ranks' digests must add up to it -- rows and list hash exactly, gamma sums to 1e-8."""
    def digest():
        ctx.prepare_step(True)
        ctx.collect_pair_collision()
        ctx.calc_mobility(MU)
        rep = ctx.solve_constraints(None, DT, 1e-30, iters, 0)
        d = ctx.constraint_digest()
        d["iterations"] = rep.iterations
        return d

    fused = ctx.comm_mode()["fused"]
    d_f = digest()
    ctx.set_option("comm_fused", 0)
    d_u = digest()
    ctx.set_option("comm_fused", 1 if fused else 0)
    M = 1 << 64
    ints = torch.tensor([[d["rows"], d["list_hash"] >> 32, d["list_hash"] & 0xffffffff, d["gamma_hash"] >> 32,
                          d["gamma_hash"] & 0xffffffff] for d in (d_f, d_u)], dtype=torch.int64, device="cuda")
    flts = torch.tensor([[d["sum_gamma"], d["sum_gamma2"], d["sum_wgamma"]] for d in (d_f, d_u)], dtype=torch.float64,
                        device="cuda")
    dist.all_reduce(ints)
    dist.all_reduce(flts)
    # gather the relaxed state on rank 0
    n = len(rods["gid"])
    cnt = torch.tensor([n], dtype=torch.int64, device="cuda")
    cnts = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(cnts, cnt)
    cnts = [int(c.item()) for c in cnts]
    nmax = max(cnts)
    pos, quat = ctx.get_rod_state()
    pack = np.zeros((nmax, 3 + 4 + 1 + 6))
    pack[:n, :3], pack[:n, 3:7], pack[:n, 7], pack[:n, 8:] = pos, quat, rods["gid"], vnc.reshape(-1, 6)
    tp = torch.from_numpy(pack).cuda()
    alls = [torch.zeros_like(tp) for _ in range(world)] if rank == 0 else None
    dist.gather(tp, alls, dst=0)
    if rank != 0:
        return None

    def comb(hi, lo):
        return ((int(hi) << 32) + int(lo)) % M

    got = [dict(rows=int(ints[i, 0]), list_hash=comb(ints[i, 1], ints[i, 2]), gamma_hash=comb(ints[i, 3], ints[i, 4]),
                sums=[float(x) for x in flts[i]]) for i in range(2)]
    full = np.concatenate([alls[r][:cnts[r]].cpu().numpy() for r in range(world)])
    ntot = len(full)
    one = alens_b200.Context(device=local)
    one.set_domain([0.0] * 3, list(gbox), [1, 1, 1])
    one.set_collision_params(1.0, 1.0, COLBUF)
    one.set_rods(full[:, 7].astype(np.int32), full[:, :3], full[:, 3:7], np.full(ntot, L_ROD), np.full(ntot, R_ROD),
                 np.zeros(ntot, dtype=np.uint8), wrap=True)
    one.collect_pair_collision()
    one.calc_mobility(MU)
    one.solve_constraints(np.ascontiguousarray(full[:, 8:]).reshape(-1), DT, 1e-30, iters, 0)
    ref = one.constraint_digest()
    one.close()
    rel = max(abs(g - r) / max(abs(r), 1e-300) for g, r in
              zip(got[0]["sums"], (ref["sum_gamma"], ref["sum_gamma2"], ref["sum_wgamma"])))
    list_ok = got[0]["rows"] == ref["rows"] and got[0]["list_hash"] == ref["list_hash"]
    # fused and unfused protocols sum the BB dot products in a different order: same list bit for bit, gamma to rounding
    rel_fu = max(abs(a - b) / max(abs(b), 1e-300) for a, b in zip(got[0]["sums"], got[1]["sums"]))
    fused_ok = got[0]["rows"] == got[1]["rows"] and got[0]["list_hash"] == got[1]["list_hash"] and rel_fu < 1e-8
    return {"status": "ok" if (list_ok and fused_ok and rel < 1e-8) else "MISMATCH",
            "pair_list_vs_single_gpu": "ok" if list_ok else "MISMATCH", "rows": got[0]["rows"], "rows_single_gpu": ref["rows"],
            "gamma_sums_rel_err_vs_single_gpu": rel, "bbpgd_iterations_compared": iters,
            "fused_vs_unfused": "ok" if fused_ok else "MISMATCH", "gamma_sums_rel_err_fused_vs_unfused": rel_fu,
            "gamma_bits_identical_fused_vs_unfused": bool(got[0]["gamma_hash"] == got[1]["gamma_hash"]),
            "fused_protocol_timed": bool(fused)}


def reference_arm(a, n, workload, ncores):
    """The reference's own CPU implementation of the path on the host cores (oracle/_ref/libalens_refsys.so: the
    reference's SylinderSystem / Constraint sources, unmodified, on the stand-in containers of oracle/stubs).  It prepares
    its own input -- the same seeded suspension, relaxed by its own initPreSteps loop (SylinderSystem.cpp:88-101) -- and
    nothing of the product is loaded.  Falls back to the C port when the library is absent."""
    rods, box = make_workload(n, a.phi, SEED, a.workload)
    steps = max(1, min(a.steps, 3))
    from oracle import pyrefsys as pr

    if pr.available():
        t0 = time.perf_counter()
        sysr = reference_system(pr, box, ncores)
        sysr.set_rods(rods)
        relax = []
        for _ in range(a.relax):  # initPreSteps: prepareStep, calcVelocityNonCon, resolveConstraints, ..., stepEuler
            sysr.prepare_step()
            sysr.calc_velocity_noncon()
            sysr.resolve_constraints()
            relax.append(int(len(sysr.constraints())))
            sysr.sum_force_velocity()
            sysr.step_euler()
        t_relax = time.perf_counter() - t0
        sy = sysr.sylinders()
        state = dict(rods)
        state["pos"], state["quat"] = sy["pos"].copy(), sy["orientation"].copy()
        vnc = thermal_velocity(state, MU, DT, seed=SEED + 17)
        for _ in range(min(a.warmup, 1)):
            reference_step(sysr, vnc)
            sysr.set_rods(state)
        ts, last = [], None
        for _ in range(steps):
            last = reference_step(sysr, vnc)
            ts.append(last["t_step"])
            nc = len(sysr.constraints())
            sysr.set_rods(state)  # every timed step starts from the same state (the rods do not move: no stepEuler)
        sysr.close()
        kind = "reference"
        sample = (f"{steps} full steps of the reference's own SylinderSystem (prepareStep {last['t_prepare']:.2f}s, "
                  f"calcVelocityNonCon {last['t_velocity']:.2f}s, resolveConstraints {last['t_resolve']:.2f}s; {nc} constraints "
                  f"found by its FDPS search), input relaxed by {a.relax} of its own steps in {t_relax:.0f}s (untimed); sources "
                  f"compiled unmodified on a stand-in for Tpetra/Eigen/MPI (oracle/stubs), OpenMP {ncores} threads")
        prepared = "reference relaxation (its own initPreSteps loop, untimed)"
    else:
        from oracle import pyoracle as po

        po.lib()
        state = relax_on_cpu(po, rods, box, a.relax, ncores)
        vnc = thermal_velocity(state, MU, DT, seed=SEED + 17)
        ts, last = [], None
        for _ in range(steps):
            last = port_step(po, state, box, vnc, ncores)
            ts.append(last["t_step"])
        kind = "port"
        sample = f"{steps} full steps of the C port (oracle/liboracle.so): {last['nc']} constraints, {last['iters']} BBPGD iterations"
        prepared = "cpu relaxation (C port)"
    t = float(np.mean(ts))
    val = round(1.0 / t, 5)
    line = {
        "impl": "reference",
        "metric": "constraint-solve steps/sec at 1M rods (1/2/4/8 B200); kernel HBM GB/s vs peak",
        "value": val, "unit": "steps/s", "n_gpus": a.gpus, "steps": steps, "warmup": min(a.warmup, 1),
        "ms_per_step": round(t * 1e3, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "rods_per_gpu": n, "rods_total": n, "input": prepared,
                   "note": "Trilinos/Eigen/MPI are absent from this image: the reference's sources run on header stand-ins, one rank"},
        "cpu_baseline": {"value": val, "unit": "steps/s", "cores": ncores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
