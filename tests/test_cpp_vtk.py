"""Sylinder_r0_*.vtp / ConBlock_r0_*.vtp / *.pvtp written by the C++ mirror (include/alens_b200/VtkPolyWriter.hpp, Sylinder.hpp,
ConstraintCollector.hpp) against the files the REFERENCE's own writers produce for the same records (SylinderSystem::writeResult
of oracle/_ref/libalens_refsys.so: Sylinder.hpp:185-453, ConstraintCollector.cpp:76-224, Util/IOHelper.hpp, Util/Base64.hpp):
byte for byte.  Host-only: compiles a small program against the headers, no GPU."""
import os
import subprocess

import numpy as np
import pytest

from scenarios import random_rods, thermal_velocity
from test_reference_pin import _system

from oracle import pyrefsys as pr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROGRAM = r'''
#include "alens_b200/ConstraintCollector.hpp"
#include "alens_b200/Sylinder.hpp"
#include <cstdio>
#include <vector>
int main(int argc, char **argv) {
    FILE *in = std::fopen(argv[1], "rb");
    int n = 0, m = 0;
    if (std::fread(&n, 4, 1, in) != 1 || std::fread(&m, 4, 1, in) != 1) return 1;
    std::vector<Sylinder> rods(n);
    std::vector<ConstraintBlock> blocks(m);
    if (n && std::fread((void *)rods.data(), sizeof(Sylinder), n, in) != (size_t)n) return 2;
    if (m && std::fread((void *)blocks.data(), sizeof(ConstraintBlock), m, in) != (size_t)m) return 3;
    const std::string folder = argv[2];
    Sylinder::writeVTP(rods, n, folder + "/", "0", 0);
    Sylinder::writePVTP(folder + "/", "0", 1);
    ConstraintCollector col;
    auto &pool = *col.constraintPoolPtr;
    for (int i = 0; i < m; i++) pool[0].push_back(blocks[i]);
    col.writeVTP(folder + "/", "", "0", 0);
    col.writePVTP(folder + "/", "", "0", 1);
    return 0;
}
'''


@pytest.mark.skipif(not pr.available(), reason="oracle/_ref/libalens_refsys.so not built (needs /root/reference)")
def test_vtp_files_equal_the_reference_writers(tmp_path):
    n, box, mu, dt = 700, 1.1, 1.0, 1e-4
    rods = random_rods(n, box, seed=3, frac_sphere=0.1, frac_immovable=0.05)
    s = _system(rods, [0.0] * 3, [box] * 3, (1, 1, 0), 0.025, mu=mu, dt=dt, conResTol=1e-5, conMaxIte=50)
    s.set_velocity_nonbrown(thermal_velocity(rods, mu, dt, seed=4))
    s.calc_velocity_noncon()
    s.resolve_constraints()   # fills velCol / forceCol, writes gamma back into the blocks
    s.sum_force_velocity()
    folder = s.write_result()
    sy, blocks = s.sylinders(), s.constraints()
    assert len(blocks) > 500 and np.abs(sy["velCol"]).max() > 0
    src, exe, data = tmp_path / "prog.cpp", tmp_path / "prog", tmp_path / "in.bin"
    src.write_text(PROGRAM)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-fopenmp", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    with open(data, "wb") as f:
        f.write(np.array([len(sy), len(blocks)], dtype=np.int32).tobytes())
        f.write(sy.tobytes())
        f.write(blocks.tobytes())
    out = tmp_path / "out"
    out.mkdir()
    subprocess.check_call([str(exe), str(data), str(out)], env=dict(os.environ, OMP_NUM_THREADS="1"))
    for name in ("Sylinder_r0_0.vtp", "Sylinder_0.pvtp", "ConBlock_r0_0.vtp", "ConBlock_0.pvtp"):
        ours, ref = (out / name).read_bytes(), open(os.path.join(folder, name), "rb").read()
        assert len(ref) > 200 and ours == ref, name
    s.close()
