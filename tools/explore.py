"""Exploration script (dev only): time the phases of one step at a given size on the GPU."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import alens_b200
from scenarios import random_rods, box_for_volume_fraction, thermal_velocity

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
phi = float(sys.argv[2]) if len(sys.argv) > 2 else 0.10
relax = int(sys.argv[3]) if len(sys.argv) > 3 else 5
L, R, colbuf, mu, dt, res = 0.25, 0.0125, 0.025, 1.0, 1e-5, 1e-5
box = box_for_volume_fraction(n, L, R, phi)
print("n", n, "box", box, flush=True)
t0 = time.time()
rods = random_rods(n, box, L, R, seed=1234)
print("gen %.2fs" % (time.time() - t0), flush=True)
ctx = alens_b200.Context(0)
ctx.set_domain([0] * 3, [box] * 3, [1, 1, 1])
ctx.set_collision_params(1.0, 1.0, colbuf)
t0 = time.time()
ctx.set_rods(rods["gid"], rods["pos"], rods["quat"], rods["length"], rods["radius"], rods["immovable"])
print("set_rods wall %.3fs" % (time.time() - t0), ctx.get_timers()["upload_ms"], flush=True)
for it in range(relax):
    t0 = time.time()
    if it > 0:
        ctx.prepare_step(True)
    nc = ctx.collect_pair_collision()
    t1 = time.time()
    ctx.calc_mobility(mu)
    rep = ctx.solve_constraints(None, dt, res, 3000, 0)
    t2 = time.time()
    tm = ctx.get_timers()
    blocks_overlap = None
    print("relax %d: nc=%d (%.2f/rod) collect %.1f ms cand %d setup %.1f solve %.1f ms ite %d res %.3g split %.2f" % (
        it, nc, nc / n, tm["collect_ms"], ctx.get_collect_stats()["candidates"], tm["setup_ms"], tm["solve_ms"],
        rep.iterations, rep.residual, tm["split_ms"]), flush=True)
    ctx.step_euler(dt)
# now a thermal step
vnc = thermal_velocity(rods, mu, dt, seed=7)
ctx.set_profiling(True)
for rep_i in range(3):
    ctx.prepare_step(True)
    nc = ctx.collect_pair_collision()
    ctx.calc_mobility(mu)
    ctx.reset_timers()
    rep = ctx.solve_constraints(vnc, dt, res, 3000, 0)
    tm = ctx.get_timers()
    print("thermal step: nc=%d solve %.2f ms ite %d res %.3g" % (nc, tm["solve_ms"], rep.iterations, rep.residual))
    for k in ("force_vel", "dtrans", "update"):
        nn = max(tm["op_%s_n" % k], 1)
        print("   %s: %.1f us avg over %d" % (k, 1e3 * tm["op_%s_ms" % k] / nn, nn))
    ninc = 2 * nc
    print("   bytes/iter est: fv %.1f MB, tail %.1f MB, upd %.1f MB" % ((ninc * 52 + nc * 0 + n * (48 + 48 + 8)) / 1e6, nc * (80 + 48) / 1e6, nc * 32 / 1e6))
