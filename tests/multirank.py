"""Harness for the slab decomposition: run R ranks (one alens_ctx each, one host thread each) inside one process.
devices="shared": every rank on device 0 (the unfused protocol: one-thread wait / signal / reduce kernels);
devices="spread": one rank per GPU when the box has at least R of them -- the FUSED kernels that bench.py times at N > 1
(halo wait and mailbox allreduce inside k_bb_tail, remote U rows from k_force_vel_act); same code path as the
multi-process bench, only the bootstrap differs (alens_comm_connect_local instead of cudaIpc blobs).
devices=None: "spread" if possible, else "shared" (ALENS_TEST_DEVICES=0,1,... overrides)."""
import os
import threading

import numpy as np

import alens_b200


def split_slabs(rods, box_lo, box_hi, nranks, axis=0):
    """owner rank of every rod by its wrapped centre; returns list of index arrays (global order preserved)"""
    L = box_hi[axis] - box_lo[axis]
    x = rods["pos"][:, axis]
    xw = box_lo[axis] + np.mod(x - box_lo[axis], L)
    w = L / nranks
    owner = np.minimum((np.floor((xw - box_lo[axis]) / w)).astype(int), nranks - 1)
    return [np.nonzero(owner == r)[0] for r in range(nranks)]


def gpu_count():
    import torch

    return torch.cuda.device_count()


def pick_devices(nranks, devices=None):
    if devices is None and os.environ.get("ALENS_TEST_DEVICES"):
        env = [int(x) for x in os.environ["ALENS_TEST_DEVICES"].split(",")]
        return [env[r % len(env)] for r in range(nranks)]
    if devices is None:
        devices = "spread" if gpu_count() >= nranks else "shared"
    if devices == "shared":
        return [0] * nranks
    if devices == "spread":
        if gpu_count() < nranks:
            raise RuntimeError(f"devices='spread' needs {nranks} GPUs")
        return list(range(nranks))
    return list(devices)


def take(rods, idx):
    return {k: v[idx] for k, v in rods.items()}


def run_ranks(rods, lo, hi, pbc, nranks, colbuf, mu, dt, res, max_ite, vnc=None, skin=None, axis=0, devices=None,
              steps=1, want_blocks=True, migrate=False, brown=None, links=None, options=None):
    """returns per-rank dicts (idx = global rod indices owned, blocks, gamma, forceU/velU/..., report, history).
    migrate: alens_migrate_rods between stepEuler and prepareStep (the rank's rod set changes: `gid` tells which it holds
    at the end, `migrated` = (sent, received) totals); brown = (kBT, seed): velNonCon = the device's Brownian velocity,
    keyed by (seed, step, gid) and therefore the same whatever the decomposition; links = (prevGid, nextGid, kappa, gap):
    EVERY rank is handed the whole link map, as every rank of the reference reads it"""
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    parts = split_slabs(rods, lo, hi, nranks, axis)
    max_r = float(np.max(0.5 * rods["length"] + rods["radius"]))
    cutoff = 2 * max_r + colbuf
    skin = 0.25 * cutoff if skin is None else skin
    w = (hi[axis] - lo[axis]) / nranks
    devices = pick_devices(nranks, devices)
    ctxs = []
    base = 0
    for r in range(nranks):
        c = alens_b200.Context(device=devices[r], rank=r, nranks=nranks)
        for k, v in (options or {}).items():
            c.set_option(k, v)
        c.set_domain(lo, hi, pbc)
        c.set_collision_params(1.0, 1.0, colbuf)
        c.set_decomposition(axis, lo[axis] + r * w, lo[axis] + (r + 1) * w, skin, max_r, base)
        c.comm_create(max(4096, 2 * max(len(p) for p in parts)))
        base += len(parts[r])
        ctxs.append(c)
    alens_b200.comm_connect_local(ctxs)
    out = [None] * nranks
    errs = [None] * nranks

    def work(r):
        try:
            c, idx = ctxs[r], parts[r]
            loc = take(rods, idx)
            c.set_rods(loc["gid"], loc["pos"], loc["quat"], loc["length"], loc["radius"], loc["immovable"], wrap=True)
            v = None if vnc is None else np.ascontiguousarray(vnc.reshape(-1, 6)[idx]).reshape(-1)
            res_r = dict(idx=idx)
            moved = [0, 0]
            for s in range(steps):
                if s > 0:
                    c.step_euler(dt)
                    if migrate:
                        a, b = c.migrate_rods()
                        moved[0] += a
                        moved[1] += b
                    c.prepare_step(True)
                nc = c.collect_pair_collision()
                if links is not None:
                    res_r["links_added"] = c.collect_link_bilateral(links[0], links[1], links[2], links[3])
                c.calc_mobility(mu)
                if brown is not None:
                    vb = c.calc_velocity_brown(brown[0], dt, None, brown[1], s)
                    c.calc_velocity_noncon(vel_brown=vb)
                    v = None
                rep = c.solve_constraints(v, dt, res, max_ite, 0)
            res_r["long"] = c.get_long_rod_stats()
            res_r["migrated"] = tuple(moved)
            res_r["identity"] = c.get_rod_identity()
            res_r.update(nc=nc, report=rep, gamma=c.get_gamma(), history=c.get_history(), ghosts=c.num_ghosts(),
                         mode=c.comm_mode(), digest=c.constraint_digest(), stress=c.sum_constraint_stress())
            res_r.update(c.get_force_velocity())
            if want_blocks:
                res_r["blocks"] = c.get_constraints(with_stress=True, write_back=True)
            res_r["state"] = c.get_rod_state()
            out[r] = res_r
        except Exception as e:  # noqa: BLE001
            errs[r] = e

    th = [threading.Thread(target=work, args=(r,)) for r in range(nranks)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for c in ctxs:
        c.close()
    for e in errs:
        if e is not None:
            raise e
    return out
