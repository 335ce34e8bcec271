"""One resident step at the bench size with the CUDA profiler API bracketing exactly that step
(run under `ncu --profile-from-start off`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
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
step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
nc, rep = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled step:", nc, rep.iterations, ctx.get_timers())
