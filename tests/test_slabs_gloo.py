"""Host side of the slab decomposition under torch.distributed (gloo, world_size 2, CPU only): ownership, the
global numbering (updateSylinderMap, SylinderSystem.cpp:868-880), the handle all-gather used for the peer-memory
bootstrap, and the ghost selection rule -- checked with the CPU oracle: the union of the slab-local pair lists
(owned + ghost rods) must be the global pair list, bit for bit, and every cross-slab pair must be seen by both owners.
The device side of the same protocol is covered by tests/test_gpu_multirank.py."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

FIELDS = ("gidI", "gidJ", "delta0", "normI", "posI", "posJ", "labI", "labJ")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _system(pbc_x):
    from scenarios import random_rods

    box = (3.2, 1.2, 1.2)
    rods = random_rods(3000, box, seed=5, length_sigma=0.2)
    return rods, [0.0, 0.0, 0.0], list(box), (pbc_x, 1, 0)


def _worker(rank, world, port, pbc_x, outdir):
    import torch.distributed as dist

    from alens_b200 import slabs
    from oracle import pyoracle as po
    from scenarios import canonical_order

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rods, lo, hi, pbc = _system(pbc_x)
        colbuf, skin, axis = 0.025, 0.05, 0
        rods["pos"] = po.wrap_positions(rods["pos"], lo, hi)
        owner = slabs.owner_of(rods["pos"], lo, hi, axis, world)
        mine = np.nonzero(owner == rank)[0]

        # global numbering: contiguous per rank, exclusive scan of the counts
        base, total = slabs.global_index_base(len(mine))
        assert total == len(rods["gid"])
        assert base == int((owner < rank).sum())

        # bootstrap all-gather: every rank ends up with every rank's handle, in rank order
        blobs = slabs.exchange_blobs(bytes([rank + 1]) * 64)
        assert blobs == [bytes([r + 1]) * 64 for r in range(world)]

        # ghost selection -> neighbours (gid lists travel; the geometry is looked up in the shared global system)
        max_r = float(np.max(0.5 * rods["length"] + rods["radius"]))
        width = slabs.ghost_width(max_r, colbuf, skin)
        local = {k: v[mine] for k, v in rods.items()}
        left, right, img_l, img_r = slabs.ghost_selection(local["pos"], lo, hi, pbc, axis, rank, world, width)
        nl, nr = slabs.neighbours(rank, world, bool(pbc[axis]))
        sends = [None] * world
        for nb, idx in ((nl, left), (nr, right)):
            if nb >= 0:
                prev = sends[nb] if sends[nb] is not None else np.zeros(0, dtype=np.int64)
                sends[nb] = np.concatenate([prev, mine[idx]])
        gathered = [None] * world
        dist.all_gather_object(gathered, sends)
        ghosts = [np.asarray(g[rank], dtype=np.int64) for g in gathered if g[rank] is not None]
        ghosts = np.unique(np.concatenate(ghosts)) if ghosts else np.zeros(0, dtype=np.int64)
        assert not np.intersect1d(ghosts, mine).size
        if world > 1:
            assert ghosts.size > 0

        # slab-local pair list with the oracle: owned + ghost rods, pairs of two ghosts are the neighbour's business
        sel = np.concatenate([mine, ghosts])
        sub = {k: v[sel] for k, v in rods.items()}
        orods = po.make_rods(sub["gid"], sub["radius"], sub["length"], sub["pos"], sub["quat"], 1.0, 1.0, colbuf)
        blocks = po.collect_pairs(orods, lo, hi, pbc, method="cells")
        own_gid = set(rods["gid"][mine].tolist())
        keep = np.array([(int(a) in own_gid) or (int(b) in own_gid) for a, b in zip(blocks["gidI"], blocks["gidJ"])],
                        dtype=bool)
        blocks = blocks[keep]
        np.save(os.path.join(outdir, f"blocks{rank}.npy"), blocks)
        dist.barrier()
        if rank == 0:
            allb = np.concatenate([np.load(os.path.join(outdir, f"blocks{r}.npy")) for r in range(world)])
            allb = allb[canonical_order(allb)]
            same = np.zeros(len(allb), bool)
            same[1:] = ((allb["gidI"][1:] == allb["gidI"][:-1]) & (allb["gidJ"][1:] == allb["gidJ"][:-1]) &
                        (allb["labJ"][1:] == allb["labJ"][:-1]).all(axis=1))
            uniq = allb[~same]
            gr = po.make_rods(rods["gid"], rods["radius"], rods["length"], rods["pos"], rods["quat"], 1.0, 1.0, colbuf)
            want = po.collect_pairs(gr, lo, hi, pbc, method="cells")
            want = want[canonical_order(want)]
            assert len(uniq) == len(want) > 1000, (len(uniq), len(want))
            for f in FIELDS:
                assert np.array_equal(uniq[f], want[f]), f
            # a pair whose rods have different owners is held by both of them
            gid2owner = dict(zip(rods["gid"].tolist(), owner.tolist()))
            cross = np.array([gid2owner[int(a)] != gid2owner[int(b)] for a, b in zip(want["gidI"], want["gidJ"])])
            assert cross.sum() > 0
            assert same.sum() == cross.sum(), (int(same.sum()), int(cross.sum()))
            for f in FIELDS:  # and the two copies are bit-identical
                assert np.array_equal(allb[f][np.nonzero(same)[0]], allb[f][np.nonzero(same)[0] - 1]), f
            with open(os.path.join(outdir, "ok"), "w") as fh:
                fh.write(f"{len(want)} pairs, {int(cross.sum())} cross-slab")
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("pbc_x", [1, 0])
def test_slab_host_logic_world_size_2(tmp_path, oracle, pbc_x):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(2, port, pbc_x, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


def test_single_process_defaults():
    from alens_b200 import slabs

    assert slabs.global_index_base(17) == (0, 17)
    assert slabs.neighbours(0, 1, True) == (-1, -1)
    assert slabs.neighbours(0, 4, True) == (3, 1) and slabs.neighbours(3, 4, False) == (2, -1)
    lo, hi = [0.0, 0.0, 0.0], [4.0, 1.0, 1.0]
    assert slabs.slab_bounds(lo, hi, 0, 1, 4) == (1.0, 2.0)
    pos = np.array([[0.1, 0, 0], [3.99, 0, 0], [4.2, 0, 0], [-0.1, 0, 0]])
    assert slabs.owner_of(pos, lo, hi, 0, 4).tolist() == [0, 3, 0, 3]
