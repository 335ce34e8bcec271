"""Host side of the slab decomposition: who owns a rod, which rods a neighbour must see as ghosts,
where a rank's rods start in the global numbering, and the bootstrap all-gather of the peer-memory handles.

Mirrors, on the host, what the reference does with FDPS domain decomposition + `updateSylinderMap`
(SimToolbox/Sylinder/SylinderSystem.cpp:868-880: globalIndex = exclusive scan of the local counts over the ranks)
and what `alens_b200/csrc/comm.cu` does on the device (`k_ghost_flags`: an owned rod whose coordinate along the slab
axis lies within `ghost_width` of a slab face is mirrored on that neighbour).  `torch.distributed` is plumbing only:
any backend works (NCCL on the GPU box, gloo in the CPU tests).
"""
import numpy as np


def slab_bounds(lo, hi, axis, rank, nranks):
    """[slab_lo, slab_hi) of `rank`: equal-width slabs along `axis`"""
    w = (float(hi[axis]) - float(lo[axis])) / nranks
    return float(lo[axis]) + rank * w, float(lo[axis]) + (rank + 1) * w


def owner_of(pos, lo, hi, axis, nranks):
    """owner rank of every rod = the slab holding its centre wrapped into [lo, hi) along `axis`"""
    L = float(hi[axis]) - float(lo[axis])
    x = np.asarray(pos, dtype=np.float64)[:, axis]
    xw = float(lo[axis]) + np.mod(x - float(lo[axis]), L)
    w = L / nranks
    return np.minimum(np.floor((xw - float(lo[axis])) / w).astype(np.int64), nranks - 1)


def ghost_width(max_bounding_radius, colbuf, skin):
    """cutoff + skin: two rods can touch up to 2 (L/2 + r) + colBuf apart (SylinderNear.hpp:367-414), and an owned
    rod may stray `skin` outside its slab before the host has to redistribute"""
    return 2.0 * float(max_bounding_radius) + float(colbuf) + float(skin)


def ghost_selection(pos, lo, hi, pbc, axis, rank, nranks, width):
    """(to_left, to_right, image_left, image_right): indices of this rank's rods the left / right neighbour needs and
    the periodic image (-1/0/+1 along `axis`) under which that neighbour sees them.  Same rule as k_ghost_flags."""
    slo, shi = slab_bounds(lo, hi, axis, rank, nranks)
    x = np.asarray(pos, dtype=np.float64)[:, axis]
    have_left = rank > 0 or (bool(pbc[axis]) and nranks > 1)
    have_right = rank < nranks - 1 or (bool(pbc[axis]) and nranks > 1)
    left = np.nonzero(x < slo + width)[0] if have_left else np.zeros(0, dtype=np.int64)
    right = np.nonzero(x >= shi - width)[0] if have_right else np.zeros(0, dtype=np.int64)
    img_left = 1 if rank == 0 else 0            # crossing the low box face: the receiver sees the rod one box up
    img_right = -1 if rank == nranks - 1 else 0
    return left, right, img_left, img_right


def neighbours(rank, nranks, periodic):
    left = rank - 1 if rank > 0 else (nranks - 1 if periodic and nranks > 1 else -1)
    right = rank + 1 if rank < nranks - 1 else (0 if periodic and nranks > 1 else -1)
    return left, right


def global_index_base(n_local, group=None):
    """exclusive scan of the local rod counts over the ranks (updateSylinderMap, SylinderSystem.cpp:868-880);
    returns (base of this rank, total)"""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return 0, int(n_local)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.tensor([int(n_local)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(counts, mine, group=group)
    counts = [int(c.item()) for c in counts]
    return sum(counts[:rank]), sum(counts)


def exchange_blobs(blob, group=None):
    """all-gather of the fixed-size peer-window handles (alens_comm_export -> alens_comm_connect), rank order"""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.from_numpy(np.frombuffer(bytes(blob), dtype=np.uint8).copy()).to(dev)
    allb = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allb, mine, group=group)
    return [bytes(t.cpu().numpy().tobytes()) for t in allb]
