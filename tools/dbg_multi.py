import sys, threading, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, alens_b200
from scenarios import random_rods, thermal_velocity
from multirank import split_slabs, take
R = 2
n, box, colbuf, mu, dt, res = 6000, (4.8, 1.6, 1.6), 0.025, 1.0, 1e-4, 1e-6
lo, hi, pbc = np.zeros(3), np.array(box), (1, 1, 1)
rods = random_rods(n, box, seed=13)
parts = split_slabs(rods, lo, hi, R)
vnc = thermal_velocity(rods, mu, dt, seed=3)
max_r = float(np.max(0.5 * rods["length"] + rods["radius"]))
ctxs = []; base = 0; w = box[0] / R
for r in range(R):
    c = alens_b200.Context(0, r, R); c.set_domain(lo, hi, pbc); c.set_collision_params(1, 1, colbuf)
    c.set_decomposition(0, r * w, (r + 1) * w, 0.07, max_r, base); c.comm_create(8192); base += len(parts[r]); ctxs.append(c)
alens_b200.comm_connect_local(ctxs)
def work(r):
    c = ctxs[r]; loc = take(rods, parts[r])
    def say(*a): print(f"[{r}] {time.time():.2f}", *a, flush=True)
    try:
        c.set_rods(loc["gid"], loc["pos"], loc["quat"], loc["length"], loc["radius"], loc["immovable"]); say("set_rods", c.num_ghosts())
        nc = c.collect_pair_collision(); say("collect", nc)
        c.calc_mobility(mu); say("mob")
        v = np.ascontiguousarray(vnc.reshape(-1, 6)[parts[r]]).reshape(-1)
        c.setup_constraints(v, dt); say("setup")
        rep = c.solve_constraints(v, dt, res, 50, 0); say("solve", rep.iterations, rep.residual)
    except Exception as e:
        say("ERR", e)
th = [threading.Thread(target=work, args=(r,)) for r in range(R)]
[t.start() for t in th]; [t.join() for t in th]
