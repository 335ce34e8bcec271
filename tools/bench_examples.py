"""Latency of one constraint-solve step on the reference's small example configurations (BASELINE.json configs 1-3):
MixMotorSliding as shipped (2 rods, 97 motor blocks), DenseMonoLayer's initial state (9 700 rods, 20 567 contacts),
Active3DNematics (500 aligned rods).  Device time per resident step (CUDA events around the C-ABI calls) next to the
reference's own SylinderSystem / ConstraintSolver on the host (oracle/_ref/libalens_refsys.so) for the same step.
usage: python tools/bench_examples.py [reps]   -> one JSON line per configuration"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import alens_b200  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from oracle import pyrefsys as pr  # noqa: E402
from scenarios import canonical_order  # noqa: E402
from test_reference_pin import _example_cases, _orods, _system  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    for name, rods, lo, hi, pbc, colbuf, mu, dt, res, extra, choice, max_ite in _example_cases(po):
        n = len(rods["gid"])
        ctx = alens_b200.Context(0)
        ctx.set_domain(lo, hi, pbc)
        ctx.set_collision_params(1.0, 1.0, colbuf)
        ctx.set_rods(rods["gid"], rods["pos"], rods["quat"], rods["length"], rods["radius"], rods["immovable"], wrap=True)
        ctx.set_velocity_noncon(np.zeros(6 * n))

        def step():
            ctx.prepare_step(True)
            nc = ctx.collect_pair_collision()
            if extra is not None:
                ctx.append_constraints(extra)
            ctx.calc_mobility(mu)
            return nc, ctx.solve_constraints(None, dt, res, max_ite, choice)

        for _ in range(5):
            nc, rep = step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            nc, rep = step()
        torch.cuda.synchronize()
        t_gpu = (time.perf_counter() - t0) / reps
        tm = ctx.get_timers()
        line = {"config": name, "rods": n, "constraints": int(ctx.num_constraints()), "solver": "APGD" if choice else "BBPGD",
                "iterations": int(rep.iterations), "gpu_ms_per_step_wall": round(t_gpu * 1e3, 4),
                "gpu_phase_ms": {k: round(tm[k], 4) for k in ("upload_ms", "collect_ms", "setup_ms", "solve_ms", "split_ms")},
                "gpu_us_per_iteration": round(1e3 * tm["solve_ms"] / max(rep.iterations, 1), 2),
                "launches_per_step": int(tm["total_launches"] / (reps + 5))}
        ctx.close()
        if pr.available():
            s = _system(rods, lo, hi, pbc, colbuf, mu=mu, dt=dt, nthreads=os.cpu_count(), conResTol=res, conMaxIte=max_ite,
                        conSolverChoice=choice)
            ts = []
            for _ in range(5):
                t0 = time.perf_counter()
                s.prepare_step()
                if extra is not None:
                    s.append_constraints(extra)
                s.calc_velocity_noncon()
                s.resolve_constraints()
                ts.append(time.perf_counter() - t0)
            line["reference_cpu_ms_per_step"] = round(1e3 * float(np.median(ts)), 3)
            line["reference_constraints"] = int(len(s.constraints()))
            line["cpu_cores"] = os.cpu_count()
            s.close()
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
