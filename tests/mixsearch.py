"""Brute-force statement of the two-species neighbour search (SimToolbox/MPI/MixPairInteraction.hpp as driven by
MPI/MixPairInteraction_test.cpp): targets and sources are wrapped into the box on the periodic axes, and every
(target, source image) with distance <= max(rsTrg, rsSrc) is a pair.  Test helper only."""
import numpy as np


def wrap(x, lo, hi, pbc):
    x = np.array(x, dtype=np.float64, copy=True)
    for k in range(3):
        if pbc[k]:
            L = hi[k] - lo[k]
            x[:, k] = lo[k] + np.mod(x[:, k] - lo[k], L)
            x[x[:, k] >= hi[k], k] -= L
    return x


def brute_mix_pairs(trg_pos, trg_rs, src_pos, src_rs, lo, hi, pbc, nimg=1):
    """sorted array of (target, source) rows, one per image within reach (nimg images each side on periodic axes)"""
    lo, hi = np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)
    tp, sp = wrap(trg_pos, lo, hi, pbc), wrap(src_pos, lo, hi, pbc)
    rr = np.maximum(np.asarray(trg_rs)[:, None], np.asarray(src_rs)[None, :])
    rows = []
    span = [range(-nimg, nimg + 1) if pbc[k] else (0,) for k in range(3)]
    for a in span[0]:
        for b in span[1]:
            for c in span[2]:
                sh = np.array([a, b, c]) * (hi - lo)
                d = tp[:, None, :] - (sp[None, :, :] + sh)
                i, j = np.nonzero((d * d).sum(axis=2) <= rr * rr)
                rows.append(np.stack([i, j], axis=1))
    rows = np.concatenate(rows) if rows else np.zeros((0, 2), dtype=np.int64)
    return rows[np.lexsort((rows[:, 1], rows[:, 0]))]
