"""SylinderAscii_*.dat written by the C++ mirror (Sylinder::writeAscii, SimToolbox/Sylinder/Sylinder.cpp:101-109 and the
header of Sylinder.hpp:459-464) is what the reference's reader (SylinderSystem.cpp:317-344, restated in
scenarios.read_rod_file) parses back.  Host-only: compiles a small program against include/alens_b200/Sylinder.hpp."""
import os
import subprocess

import numpy as np

from scenarios import quat_from_z_to, read_rod_file

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROGRAM = r'''
#include "alens_b200/Sylinder.hpp"
#include <cstdlib>
int main(int argc, char **argv) {
    FILE *in = std::fopen(argv[1], "r");
    FILE *out = std::fopen(argv[2], "w");
    int n;
    if (std::fscanf(in, "%d", &n) != 1) return 1;
    SylinderAsciiHeader h;
    h.nparticle = n;
    h.time = 0.125;
    h.writeAscii(out);
    for (int i = 0; i < n; i++) {
        Sylinder sy;
        int imm;
        if (std::fscanf(in, "%d %d %lf %lf %lf %lf %lf %lf %lf %lf %lf", &sy.gid, &imm, &sy.radius, &sy.length, &sy.pos[0],
                        &sy.pos[1], &sy.pos[2], &sy.orientation[0], &sy.orientation[1], &sy.orientation[2],
                        &sy.orientation[3]) != 11)
            return 2;
        sy.isImmovable = imm != 0;
        sy.group = 3;
        sy.writeAscii(out);
    }
    std::fprintf(out, "L 0 1\n");
    std::fclose(out);
    return 0;
}
'''


def test_ascii_round_trip(tmp_path, oracle):
    rng = np.random.default_rng(2)
    n = 40
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    quat = quat_from_z_to(d)
    pos = rng.uniform(-5, 5, size=(n, 3))
    length = rng.uniform(0.1, 2.0, size=n)
    radius = rng.uniform(0.01, 0.05, size=n)
    imm = (rng.uniform(size=n) < 0.3).astype(int)
    src = tmp_path / "prog.cpp"
    src.write_text(PROGRAM)
    exe = tmp_path / "prog"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    inp = tmp_path / "rods.txt"
    with open(inp, "w") as f:
        f.write(f"{n}\n")
        for i in range(n):
            f.write(" ".join(repr(float(x)) for x in (i, imm[i], radius[i], length[i], *pos[i], *quat[i])).replace(".0 ", " ", 2) + "\n")
    out = tmp_path / "SylinderAscii_0.dat"
    subprocess.check_call([str(exe), str(inp), str(out)])
    lines = out.read_text().splitlines()
    assert lines[0].split() == [str(n)] and abs(float(lines[1]) - 0.125) < 1e-6  # "%d \n %lf\n"
    assert lines[-1] == "L 0 1"
    assert lines[2].split()[0] in ("C", "S") and lines[2].split()[-1] == "3"
    back = read_rod_file(str(out))
    assert np.array_equal(back["gid"], np.arange(n)) and np.array_equal(back["immovable"], imm)
    np.testing.assert_allclose(back["radius"], radius, rtol=1e-7)
    np.testing.assert_allclose(back["length"], length, rtol=1e-6)
    np.testing.assert_allclose(back["pos"], pos, atol=1e-6)
    # the direction survives (the quaternion itself is not unique)
    dirs = np.array([oracle.quat_to_dir(q) for q in back["quat"]])
    np.testing.assert_allclose(dirs, d, atol=1e-6)
