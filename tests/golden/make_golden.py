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


def mixmotorsliding():
    """Examples/MixMotorSliding as shipped: 2 rods (gid 0 immovable) and 100 doubly bound motors.  The motors become
    bilateral constraint blocks exactly as TubuleSystem::setProteinConstraints builds them (SRC/TubuleSystem.cpp:694-745):
    delta0 = |Q - P| - freeLength, gamma0 = -delta0 kappa, normI = (P - Q)/|P - Q|, posI = P - centerI, posJ = Q - centerJ,
    kappa = 100 pN/um, freeLength = 0.05 um (Examples/MixMotorSliding/ProteinConfig.yaml)."""
    base = "/root/reference/Examples/MixMotorSliding/"
    rods = read_rod_file(base + "TubuleInitial.dat")
    kappa, free_len = 100.0, 0.05
    P, Q, bI, bJ = [], [], [], []
    with open(base + "ProteinInitial.dat") as f:
        for ln in f:
            t = ln.split()
            if not t or t[0] != "P":
                continue
            P.append([float(x) for x in t[3:6]])
            Q.append([float(x) for x in t[6:9]])
            bI.append(int(t[9]))
            bJ.append(int(t[10]))
    P, Q, bI, bJ = np.array(P), np.array(Q), np.array(bI), np.array(bJ)
    both = (bI >= 0) & (bJ >= 0)  # singly bound motors are not constraints (TubuleSystem.cpp:708-711)
    P, Q, bI, bJ = P[both], Q[both], bI[both], bJ[both]
    gid2idx = {int(g): i for i, g in enumerate(rods["gid"])}
    iI = np.array([gid2idx[g] for g in bI])
    iJ = np.array([gid2idx[g] for g in bJ])
    from alens_b200.capi import BLOCK_DTYPE

    blk = np.zeros(len(P), dtype=BLOCK_DTYPE)
    pq = P - Q
    dist = np.sqrt((pq**2).sum(axis=1))
    blk["delta0"] = dist - free_len
    blk["gamma"] = -blk["delta0"] * kappa
    blk["gidI"], blk["gidJ"] = bI, bJ
    blk["globalIndexI"], blk["globalIndexJ"] = iI, iJ
    blk["oneSide"], blk["bilateral"], blk["kappa"] = 0, 1, kappa
    blk["normI"] = pq / dist[:, None]
    blk["normJ"] = -blk["normI"]
    blk["posI"] = P - rods["pos"][iI]
    blk["posJ"] = Q - rods["pos"][iJ]
    blk["labI"], blk["labJ"] = P, Q
    print("MixMotorSliding: rods", len(rods["gid"]), "bilateral blocks", len(blk))
    np.savez_compressed(os.path.join(HERE, "mixmotorsliding.npz"), gid=rods["gid"], pos=rods["pos"], quat=rods["quat"],
                        length=rods["length"], radius=rods["radius"], immovable=rods["immovable"],
                        lo=np.zeros(3), hi=np.array([20.0, 1.0, 1.0]), pbc=np.zeros(3, dtype=np.int32), colbuf=0.025,
                        dt=1e-5, res=1e-5, mu=1.0, blocks=blk.view(np.uint8))


if __name__ == "__main__":
    po.build(ref=True)
    densemonolayer()
    mixmotorsliding()
