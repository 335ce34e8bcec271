"""Two-species neighbour search on the device (alens_mix_pair_search) against the brute-force statement and, when the
reference's own MixPairInteraction is built (oracle/_ref), against that -- SURVEY.md 8f.4."""
import numpy as np
import pytest

from mixsearch import brute_mix_pairs
from scenarios import random_rods

pytestmark = pytest.mark.gpu


def device_pairs(ctx, trg_pos, trg_rs, src_rs=None):
    row, idx = ctx.mix_pair_search(trg_pos, trg_rs, src_rs)
    assert row[0] == 0 and row[-1] == len(idx) and np.all(np.diff(row) >= 0)
    t = np.repeat(np.arange(len(trg_rs)), np.diff(row))
    rows = np.stack([t, idx.astype(np.int64)], axis=1)
    return rows[np.lexsort((rows[:, 1], rows[:, 0]))]


@pytest.mark.parametrize("pbc", [(1, 1, 1), (0, 0, 0), (1, 0, 1)])
@pytest.mark.parametrize("own_radii", [False, True])
def test_mix_search_equals_brute_force_and_the_reference(ctx, pbc, own_radii):
    rng = np.random.default_rng(11)
    box = 3.0
    rods = random_rods(2500, box, seed=21, frac_sphere=0.2)
    lo, hi = [0.0] * 3, [box, box * 0.8, box * 0.6]
    rods["pos"] *= np.array(hi) / box
    colbuf = 0.025
    ctx.set_domain(lo, hi, pbc)
    ctx.set_collision_params(1.0, 1.0, colbuf)
    ctx.set_rods(rods["gid"], rods["pos"], rods["quat"], rods["length"], rods["radius"], rods["immovable"], wrap=True)
    nt = 1800
    trg = rng.uniform(-0.5, box + 0.5, size=(nt, 3)) if all(pbc) else rng.uniform(0.02, 0.98, size=(nt, 3)) * np.array(hi)
    trs = rng.uniform(0.02, 0.45, size=nt)
    trs[::7] = 0.0  # a target that only sees sources through THEIR radius
    srs = rng.uniform(0.01, 0.3, size=len(rods["gid"])) if own_radii else 0.5 * (rods["length"] + 2 * rods["radius"]) + colbuf
    got = device_pairs(ctx, trg, trs, srs if own_radii else None)
    src_pos = ctx.get_positions()
    want = brute_mix_pairs(trg, trs, src_pos, srs, lo, hi, pbc)
    assert len(want) > 2000
    assert np.array_equal(got, want)
    from oracle import pyrefsys as pr
    if pr.available():
        ref, _ = pr.mix_search(trg, trs, src_pos, srs, lo, hi, pbc)
        assert np.array_equal(got, ref)


def test_mix_search_edge_cases(ctx):
    rods = random_rods(300, 1.0, seed=3)
    ctx.set_domain([0, 0, 0], [1, 1, 1], (1, 1, 1))
    ctx.set_collision_params(1.0, 1.0, 0.025)
    ctx.set_rods(rods["gid"], rods["pos"], rods["quat"], rods["length"], rods["radius"], rods["immovable"], wrap=True)
    row, idx = ctx.mix_pair_search(np.zeros((0, 3)), np.zeros(0))  # no targets
    assert len(row) == 1 and row[0] == 0 and len(idx) == 0
    # a search radius as large as the box: each source is seen through several images
    trg = np.array([[0.5, 0.5, 0.5]])
    got = device_pairs(ctx, trg, np.array([1.0]), np.full(300, 0.01))
    want = brute_mix_pairs(trg, [1.0], ctx.get_positions(), np.full(300, 0.01), [0, 0, 0], [1, 1, 1], (1, 1, 1), nimg=2)
    assert np.array_equal(got, want) and len(got) > 300
