"""Timing experiments on k_force_vel_act: the kernel on the state a real BBPGD solve left behind, with parts of its
memory traffic switched off (alens_set_option force_dbg bits: 1 columns, 2 multipliers, 4 U store, 8 rod data,
16 nothing survives the mask, 32 no ids).  Results are garbage when dbg != 0: timing only."""
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
ctx.prepare_step(True)
nc = ctx.collect_pair_collision()
ctx.calc_mobility(bench.MU)
sets = json.loads(os.environ.get("ALENS_DBG", "null")) or [{"force_dbg": d} for d in (0, 1, 2, 3, 4, 8, 12, 16, 48, 63)]
for opts in sets:
    for k, v in opts.items():
        ctx.set_option(k, v)
    dbg = opts.get("force_dbg", 0)
    ctx.set_option("force_dbg", 0)
    rep = ctx.solve_constraints(None, bench.DT, bench.RES, 30, 0)  # a real state: 30 iterations in
    ctx.set_option("force_dbg", dbg)
    print(json.dumps({"opts": opts, "nc": nc, "force_vel_last_us": round(ctx.time_kernel("force_vel_last", 30), 2)}), flush=True)
    ctx.set_option("force_dbg", 0)
