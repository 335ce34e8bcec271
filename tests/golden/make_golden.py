"""Generates tests/golden/*.npz.  Run in the container that has /root/reference (it cannot travel to the
GPU box).  densemonolayer.npz: the rods of Examples/DenseMonoLayer/TubuleInitial.dat (parsed with the
reference's rules, SylinderSystem.cpp:317-344) together with the pair list found by the reference's own
FDPS search (oracle/_ref) and the oracle's geometric list at generation time."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import pyoracle as po  # noqa: E402
from scenarios import read_rod_file  # noqa: E402


def densemonolayer():
    rods = read_rod_file("/root/reference/Examples/DenseMonoLayer/TubuleInitial.dat")
    lo, hi, pbc, colbuf = np.zeros(3), np.full(3, 11.0), np.array([1, 1, 0], dtype=np.int32), 0.025
    pos = po.wrap_positions(rods["pos"], lo, hi)
    orods = po.make_rods(rods["gid"], rods["radius"], rods["length"], pos, rods["quat"], colBuf=colbuf)
    ref, _ = po.fdps_collect(orods, lo, hi, pbc, nthreads=1)
    geo = po.collect_pairs(orods, lo, hi, pbc, method="cells")
    print("DenseMonoLayer: rods", len(orods), "FDPS pairs", len(ref), "geometric pairs", len(geo))
    np.savez_compressed(os.path.join(HERE, "densemonolayer.npz"), gid=rods["gid"], pos=pos.astype(np.float64),
                        quat=rods["quat"], length=rods["length"], radius=rods["radius"], lo=lo, hi=hi, pbc=pbc,
                        colbuf=colbuf, ref_gidI=ref["gidI"], ref_gidJ=ref["gidJ"], ref_delta0=ref["delta0"],
                        geo_gidI=geo["gidI"], geo_gidJ=geo["gidJ"], geo_delta0=geo["delta0"])


if __name__ == "__main__":
    po.build(ref=True)
    densemonolayer()
