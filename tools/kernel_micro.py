"""Micro-benchmark of the BBPGD kernels on the 1M-rod bench workload (alens_time_kernel) under option sets.
ALENS_MICRO='[{"force_chunk":4},...]' python tools/kernel_micro.py [n_rods]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import alens_b200
import bench
from scenarios import thermal_velocity

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
rods, box = bench.make_workload(n, 0.10, bench.SEED)
ctx = alens_b200.Context(0)
rods, info = bench.relax_on_gpu(ctx, rods, box, 2)
vnc = thermal_velocity(rods, bench.MU, bench.DT, seed=bench.SEED + 17)
ctx.set_rods(rods["gid"], rods["pos"], rods["quat"], rods["length"], rods["radius"], rods["immovable"])
nc = ctx.collect_pair_collision()
ctx.calc_mobility(bench.MU)
sets = json.loads(os.environ.get("ALENS_MICRO", "null")) or [{"force_chunk": 2}, {"force_chunk": 4}]
for opts in sets:
    for k, v in opts.items():
        ctx.set_option(k, v)
    row = {"opts": opts, "nc": nc}
    for which in ("force_vel", "tail", "force_vel_plain"):
        ctx.setup_constraints(vnc, bench.DT)
        row[which + "_us"] = round(ctx.time_kernel(which, 30), 2)
    print(json.dumps(row), flush=True)
