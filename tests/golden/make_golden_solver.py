"""Generates tests/golden/refsolver.npz: gamma, uni/bi velocities and every IteHistory row that the REFERENCE's own
ConstraintCollector + ConstraintSolver + BCQPSolver (oracle/_ref/libalens_refsys.so = the reference sources compiled
unmodified against oracle/stubs, see oracle/ref_system_driver.cpp) produce for the three example configurations of
BASELINE.json -- MixMotorSliding as shipped (97 motor blocks), DenseMonoLayer's initial state (20 567 collision blocks),
Active3DNematics (500 aligned rods; BBPGD and APGD) -- on the full geometric constraint list in canonical order.

Run in the build container (needs /root/reference for `make -C oracle refsys`):
    python tests/golden/make_golden_solver.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import pyoracle as po  # noqa: E402
from oracle import pyrefsys as pr  # noqa: E402
from scenarios import canonical_order  # noqa: E402
from test_reference_pin import _example_cases, _orods, _system  # noqa: E402


def main():
    out = {}
    for name, rods, lo, hi, pbc, colbuf, mu, dt, res, extra, choice, max_ite in _example_cases(po):
        orods = _orods(po, rods, lo, hi, pbc, colbuf)
        coll = po.collect_pairs(orods, lo, hi, pbc, with_stress=True)
        blocks = coll[canonical_order(coll)]
        if extra is not None:
            blocks = np.concatenate([blocks, extra])
        s = _system(rods, lo, hi, pbc, colbuf, mu=mu, dt=dt)
        r = s.solve_blocks(blocks, np.zeros(6 * len(rods["gid"])), dt, res, max_ite, choice)
        s.close()
        print(name, "constraints", len(blocks), "iterations", r["nIte"], "residual", r["history"][-1][4], "status", r["status"])
        for k in ("gamma", "history", "velU", "velB"):
            out[f"{name}_{k}"] = r[k]
    np.savez_compressed(os.path.join(HERE, "refsolver.npz"), **out)


if __name__ == "__main__":
    main()
