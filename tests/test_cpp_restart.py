"""Restart path of the C++ mirror (SylinderSystem::reinitialize / readSylinderVTK / readRestartFile / Sylinder::stepEuler,
reference: SylinderSystem.cpp:106-175, :406-476, Sylinder.cpp:91-99) on the host: tests/cpp/test_restart.cpp writes a
two-piece snapshot with the mirror's writers and reads it back; with oracle/_ref built, a snapshot written by the REFERENCE's
own SylinderSystem::writeResult is read through the same code and compared with the reference's records."""
import os
import subprocess

import numpy as np
import pytest

from oracle import pyrefsys as pr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_restart")
LIB = os.path.join(ROOT, "alens_b200", "libalens_b200.so")


def _build():
    if not os.path.exists(LIB):
        pytest.skip("libalens_b200.so not built")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp"), "test_restart"])


def test_snapshot_round_trip(tmp_path):
    _build()
    (tmp_path / "result" / "result0-399").mkdir(parents=True)
    r = subprocess.run([EXE, str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "restart ok" in r.stdout, r.stderr[-2000:]


@pytest.mark.skipif(not pr.available(), reason="oracle/_ref/libalens_refsys.so not built (needs /root/reference)")
def test_reads_the_snapshot_the_reference_writes(tmp_path):
    from scenarios import random_rods, thermal_velocity
    from test_reference_pin import _system

    _build()
    n, box, mu, dt = 600, 1.2, 1.0, 1e-4
    rods = random_rods(n, box, seed=9, frac_sphere=0.1, frac_immovable=0.05)
    s = _system(rods, [0.0] * 3, [box] * 3, (1, 0, 1), 0.025, mu=mu, dt=dt, conResTol=1e-5, conMaxIte=50)
    s.set_velocity_nonbrown(thermal_velocity(rods, mu, dt, seed=2))
    s.calc_velocity_noncon()
    s.resolve_constraints()
    s.sum_force_velocity()  # vel / omega of the records: what a restart steps with
    folder = s.write_result()
    ref = s.sylinders().copy()
    out = tmp_path / "rods.bin"
    r = subprocess.run([EXE, "read", os.path.join(folder, "Sylinder_0.pvtp"), str(out)], capture_output=True, text=True, timeout=120)
    s.close()  # removes the reference's working folder
    assert r.returncode == 0, r.stderr[-2000:]
    got = np.fromfile(out, dtype=pr.SYLINDER_DTYPE)
    assert len(got) == n and np.abs(ref["vel"]).max() > 0
    for k in ("gid", "group", "isImmovable"):
        assert np.array_equal(got[k], ref[k]), k
    for k in ("radius", "radiusCollision", "length", "lengthCollision", "vel", "omega"):  # Float32 in the file
        assert np.array_equal(got[k], ref[k].astype(np.float32).astype(np.float64)), k
    assert np.abs(got["pos"] - ref["pos"]).max() < 1e-14 * (1 + np.abs(ref["pos"]).max())

    def direction(q):
        x, y, z, w = q.T
        return np.stack([2 * (x * z + w * y), 2 * (y * z - w * x), 1 - 2 * (x * x + y * y)], axis=1)

    assert np.abs(direction(got["orientation"]) - direction(ref["orientation"])).max() < 3e-7  # znorm is Float32


@pytest.mark.skipif(not pr.available(), reason="oracle/_ref/libalens_refsys.so not built (needs /root/reference)")
def test_host_euler_step_equals_the_reference(tmp_path):
    """Sylinder::stepEuler of the mirror (the step a restart takes on the host) against the reference's own
    Sylinder::stepEuler + EquatnHelper::rotateEquatn (Sylinder.cpp:91-99, Util/EquatnHelper.hpp:74-90)"""
    _build()
    rng = np.random.default_rng(4)
    n, dt = 400, 2e-3
    rods = np.zeros(n, dtype=pr.SYLINDER_DTYPE)
    rods["pos"] = rng.uniform(-5, 5, size=(n, 3))
    q = rng.normal(size=(n, 4))
    rods["orientation"] = q / np.linalg.norm(q, axis=1)[:, None]
    rods["vel"] = rng.normal(size=(n, 3))
    rods["omega"] = rng.normal(size=(n, 3)) * rng.choice([0.0, 1e-9, 1.0, 40.0], size=(n, 1))  # below and above the float-epsilon cut
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    rods.tofile(fin)
    r = subprocess.run([EXE, "euler", str(fin), repr(dt), str(fout)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    got = np.fromfile(fout, dtype=pr.SYLINDER_DTYPE)
    for i in range(n):
        p, o = pr.sylinder_step_euler(rods["pos"][i], rods["orientation"][i], rods["vel"][i], rods["omega"][i], dt)
        assert np.array_equal(got["pos"][i], p), i
        assert np.abs(got["orientation"][i] - o).max() < 4e-16, i  # same formula; Eigen normalises with a vectorised norm
