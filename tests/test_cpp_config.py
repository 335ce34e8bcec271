"""The mirror's readers of the reference's input files -- SylinderConfig(RunConfig.yaml) and the SylinderInitial.dat reader of
SylinderSystem -- against the reference's own SylinderConfig.cpp / setInitialFromFile (oracle/_ref/libalens_refsys.so).
Host-only: compiles a small program against the headers, no GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from scenarios import random_rods

from oracle import pyrefsys as pr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROGRAM = r'''
#include "alens_b200/SylinderSystem.hpp"
#include <cstdio>
int main(int argc, char **argv) {
    try {
        SylinderConfig c(argv[1]);
        std::printf("%.17g %d %d", (double)c.rngSeed, c.logLevel, c.timerLevel);
        for (int d = 0; d < 3; d++) std::printf(" %.17g", c.simBoxLow[d]);
        for (int d = 0; d < 3; d++) std::printf(" %.17g", c.simBoxHigh[d]);
        for (int d = 0; d < 3; d++) std::printf(" %d", c.simBoxPBC[d] ? 1 : 0);
        std::printf(" %d", c.monolayer ? 1 : 0);
        for (int d = 0; d < 3; d++) std::printf(" %.17g", c.initBoxLow[d]);
        for (int d = 0; d < 3; d++) std::printf(" %.17g", c.initBoxHigh[d]);
        for (int d = 0; d < 3; d++) std::printf(" %.17g", c.initOrient[d]);
        std::printf(" %d %d %.17g %.17g %.17g %.17g %d %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %d %d\n",
                    c.initCircularX ? 1 : 0, c.initPreSteps, c.viscosity, c.KBT, c.linkKappa, c.linkGap, c.sylinderFixed ? 1 : 0,
                    c.sylinderNumber, c.sylinderLength, c.sylinderLengthSigma, c.sylinderDiameter, c.sylinderDiameterColRatio,
                    c.sylinderLengthColRatio, c.sylinderColBuf, c.dt, c.timeTotal, c.timeSnap, c.conResTol, c.conMaxIte,
                    c.conSolverChoice);
        std::printf("%zu\n", c.boundaries.size());
        if (argc > 3) {
            const std::vector<Sylinder> rods = SylinderSystem::readSylinderFile(argv[2]);
            FILE *f = std::fopen(argv[3], "wb");
            std::fwrite(rods.data(), sizeof(Sylinder), rods.size(), f);
            std::fclose(f);
        }
    } catch (const std::exception &e) {
        std::printf("ERROR %s\n", e.what());
        return 3;
    }
    return 0;
}
'''


@pytest.fixture(scope="module")
def prog(tmp_path_factory):
    d = tmp_path_factory.mktemp("cfg")
    src, exe = d / "prog.cpp", d / "prog"
    src.write_text(PROGRAM)
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", os.path.join(ROOT, "alens_b200"), "-lalens_b200", "-Wl,-rpath," + os.path.join(ROOT, "alens_b200")])
    return str(exe)


def reference_config(path):
    L = pr.lib()
    out = np.zeros(64)
    with pr._Quiet():
        nb = L.refsys_parse_config(path.encode(), out.ctypes.data_as(C.POINTER(C.c_double)))
    k = 42
    assert out[k] == k, "layout of refsys_parse_config changed"
    return out[:k], nb


@pytest.mark.skipif(not pr.available(), reason="oracle/_ref/libalens_refsys.so not built (needs /root/reference)")
@pytest.mark.parametrize("case", ["defaults", "full"])
def test_run_config_yaml_is_read_like_the_reference_reads_it(prog, tmp_path, case):
    cfg = dict(pr.DEFAULTS)
    bnd = ()
    if case == "defaults":  # only the required keys: every optional one takes the reference's default (SylinderConfig.cpp:30-73)
        for k in ("logLevel", "timerLevel", "monolayer", "initPreSteps", "sylinderLengthSigma", "sylinderFixed", "sylinderColBuf",
                  "sylinderDiameterColRatio", "sylinderLengthColRatio", "linkKappa", "linkGap"):
            cfg.pop(k, None)
        cfg.update(simBoxLow=[-1.5, 0.0, 2.0], simBoxHigh=[3.5, 4.0, 9.0], simBoxPBC=[True, False, True], viscosity=0.75)
    else:
        cfg.update(rngSeed=77, logLevel=3, timerLevel=5, simBoxLow=[0.0, -2.0, 1.0], simBoxHigh=[11.0, 2.0, 6.5],
                   simBoxPBC=[False, True, False], monolayer=True, initBoxLow=[1.0, -1.0, 2.0], initBoxHigh=[9.0, 1.0, 5.0],
                   initOrient=[0.0, 0.5, 2.0], initCircularX=True, initPreSteps=17, thermEquilTime=0.0, viscosity=0.9, KBT=0.00411,
                   linkKappa=123.5, linkGap=0.07, sylinderFixed=True, sylinderNumber=321, sylinderLength=0.75,
                   sylinderLengthSigma=0.25, sylinderDiameter=0.03, sylinderDiameterColRatio=1.1, sylinderLengthColRatio=0.95,
                   sylinderColBuf=0.04, dt=2e-5, timeTotal=3.0, timeSnap=0.01, conResTol=3e-6, conMaxIte=4321, conSolverChoice=1)
        bnd = (dict(type="wall", center=[0.0, 0.0, 1.0], norm=[0.0, 0.0, 1.0]),
               dict(type="tube", center=[0.0, 0.0, 0.0], axis=[1.0, 0.0, 0.0], radius=2.0, inside=True),
               dict(type="sphere", center=[5.0, 0.0, 3.0], radius=9.0, inside=True))
    path = str(tmp_path / "RunConfig.yaml")
    pr.write_yaml(path, cfg, bnd)
    want, nb = reference_config(path)
    r = subprocess.run([prog, path], capture_output=True, text=True, check=True)
    lines = r.stdout.strip().splitlines()
    got = np.array([float(x) for x in lines[0].split()])
    assert len(got) == len(want)
    assert np.array_equal(got, want), (got, want)
    assert int(lines[1]) == nb == len(bnd)


@pytest.mark.skipif(not pr.available(), reason="oracle/_ref/libalens_refsys.so not built (needs /root/reference)")
def test_required_key_missing_is_an_error(prog, tmp_path):
    cfg = dict(pr.DEFAULTS)
    cfg.pop("conResTol")
    path = str(tmp_path / "RunConfig.yaml")
    pr.write_yaml(path, cfg)
    r = subprocess.run([prog, path], capture_output=True, text=True)
    assert r.returncode == 3 and "conResTol" in r.stdout  # the reference logs "critical" and exits (SylinderConfig.cpp:11-28)


@pytest.mark.skipif(not pr.available(), reason="oracle/_ref/libalens_refsys.so not built (needs /root/reference)")
def test_initial_dat_file_is_read_like_the_reference_reads_it(prog, tmp_path):
    """SylinderInitial.dat: `C|S gid radius mx my mz px py pz [group]` -> Sylinder records (setInitialFromFile :317-375)"""
    rods = random_rods(400, 3.0, seed=12, frac_sphere=0.1, frac_immovable=0.1, length_sigma=0.3)
    cfg = dict(pr.DEFAULTS)
    cfg.update(simBoxHigh=[3.0] * 3, sylinderNumber=400)
    ypath, dpath, out = str(tmp_path / "RunConfig.yaml"), str(tmp_path / "SylinderInitial.dat"), str(tmp_path / "rods.bin")
    pr.write_yaml(ypath, cfg)
    pr.write_dat(dpath, rods, [(int(rods["gid"][0]), int(rods["gid"][1]))])
    subprocess.run([prog, ypath, dpath, out], check=True, capture_output=True)
    got = np.fromfile(out, dtype=pr.SYLINDER_DTYPE)
    s = pr.RefSystem(yaml_file=ypath, pos_file=dpath, nthreads=1)
    want = s.sylinders().copy()
    s.close()
    assert len(got) == len(want) == 400
    order_g, order_w = np.argsort(got["gid"]), np.argsort(want["gid"])
    for f in ("gid", "isImmovable", "radius", "length", "pos", "orientation", "group"):
        assert np.array_equal(got[f][order_g], want[f][order_w]), f
