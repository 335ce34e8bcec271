"""Per-kernel timing of the BCQP loop under the library's tuning knobs (alens_set_option), 1M-rod bench workload.
Run on the GPU box: python tools/kernel_sweep.py [n_rods] -> gpurun_out/sweep.json"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import alens_b200
import bench
from scenarios import thermal_velocity

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
rods, box = bench.make_workload(n, 0.10, bench.SEED)
ctx = alens_b200.Context(0)
rods, info = bench.relax_on_gpu(ctx, rods, box, 4)
vnc = thermal_velocity(rods, bench.MU, bench.DT, seed=bench.SEED + 17)
ctx.set_rods(rods["gid"], rods["pos"], rods["quat"], rods["length"], rods["radius"], rods["immovable"])
ctx.set_velocity_noncon(vnc)

def step():
    ctx.prepare_step(True)
    nc = ctx.collect_pair_collision()
    ctx.calc_mobility(bench.MU)
    return nc, ctx.solve_constraints(None, bench.DT, bench.RES, bench.MAXITE, 0)

out = []
ref_gamma = None
SWEEP = json.loads(os.environ.get("ALENS_SWEEP", "null")) or [
    {"force_kernel": 1}, {"force_kernel": 0, "force_block": 64, "force_chunk": 2}]
for opts in SWEEP:
    for k, v in opts.items():
        ctx.set_option(k, v)
    step()
    ctx.set_profiling(True)
    ctx.reset_timers()
    acc = {}
    for _ in range(3):
        nc, rep = step()
        tm = ctx.get_timers()
        for k in ("upload_ms", "collect_ms", "setup_ms", "solve_ms", "split_ms"):
            acc[k] = acc.get(k, 0.0) + tm[k] / 3
    tm = ctx.get_timers()
    ctx.set_profiling(False)
    g = ctx.get_gamma()
    if ref_gamma is None:
        ref_gamma = g
    row = dict(opts=opts, nc=nc, iters=rep.iterations, phases=acc,
               force_vel_us=1e3 * tm["op_force_vel_ms"] / max(tm["op_force_vel_n"], 1),
               tail_us=1e3 * tm["op_dtrans_ms"] / max(tm["op_dtrans_n"], 1),
               gamma_identical=bool(g.shape == ref_gamma.shape and (g == ref_gamma).all()), cand=ctx.get_collect_stats())
    print(json.dumps(row), flush=True)
    out.append(row)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w") as f:
    json.dump(out, f, indent=1)
