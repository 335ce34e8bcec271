"""ctypes access to oracle/_ref/libalens_refsys.so: the reference's OWN SylinderSystem / ConstraintSolver / BCQPSolver
sources, compiled unmodified against the stand-in headers of oracle/stubs (see oracle/ref_system_driver.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, bench.py's CPU legs and tests/golden/make_golden.py, never by the
product.  The library travels to the GPU box prebuilt; /root/reference is only needed to (re)build it."""
import ctypes as C
import os
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "_ref", "libalens_refsys.so")

BLOCK_DTYPE = np.dtype([
    ("delta0", "<f8"), ("gamma", "<f8"), ("gammaLB", "<f8"),
    ("gidI", "<i4"), ("gidJ", "<i4"), ("globalIndexI", "<i4"), ("globalIndexJ", "<i4"),
    ("oneSide", "u1"), ("bilateral", "u1"), ("pad_", "u1", 6), ("kappa", "<f8"),
    ("normI", "<f8", 3), ("normJ", "<f8", 3), ("posI", "<f8", 3), ("posJ", "<f8", 3),
    ("labI", "<f8", 3), ("labJ", "<f8", 3), ("stress", "<f8", 9)])
assert BLOCK_DTYPE.itemsize == 272

# SimToolbox/Sylinder/Sylinder.hpp:38-84 (568 bytes)
SYLINDER_DTYPE = np.dtype([
    ("gid", "<i4"), ("globalIndex", "<i4"), ("rank", "<i4"), ("group", "<i4"), ("isImmovable", "u1"), ("pad_", "u1", 7),
    ("radius", "<f8"), ("radiusCollision", "<f8"), ("length", "<f8"), ("lengthCollision", "<f8"),
    ("radiusSearch", "<f8"), ("sepmin", "<f8"), ("colBuf", "<f8"), ("pos", "<f8", 3), ("orientation", "<f8", 4),
    ("vel", "<f8", 3), ("omega", "<f8", 3), ("velCol", "<f8", 3), ("omegaCol", "<f8", 3), ("velBi", "<f8", 3),
    ("omegaBi", "<f8", 3), ("velNonB", "<f8", 3), ("omegaNonB", "<f8", 3), ("force", "<f8", 3), ("torque", "<f8", 3),
    ("forceCol", "<f8", 3), ("torqueCol", "<f8", 3), ("forceBi", "<f8", 3), ("torqueBi", "<f8", 3),
    ("forceNonB", "<f8", 3), ("torqueNonB", "<f8", 3), ("velBrown", "<f8", 3), ("omegaBrown", "<f8", 3)])
assert SYLINDER_DTYPE.itemsize == 568

_lib = None


def available():
    return os.path.exists(PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{PATH} missing: `make -C oracle refsys` (needs /root/reference)")
        _lib = C.CDLL(PATH)
        _lib.refsys_create.restype = C.c_void_p
        _lib.refsys_num_constraints.restype = C.c_longlong
        for f in ("refsys_collect_pair_collision", "refsys_collect_boundary_collision", "refsys_collect_link_bilateral"):
            getattr(_lib, f).restype = C.c_longlong
        assert _lib.refsys_sizeof_block() == 272 and _lib.refsys_sizeof_sylinder() == 568
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class _Quiet:
    """FDPS prints a banner and the reference logs through spdlog/stdout: keep test logs readable"""

    def __init__(self, on=True):
        self.on = on

    def __enter__(self):
        if self.on:
            self.null = os.open(os.devnull, os.O_WRONLY)
            self.saved = (os.dup(1), os.dup(2))
            os.dup2(self.null, 1)
            os.dup2(self.null, 2)

    def __exit__(self, *a):
        if self.on:
            os.dup2(self.saved[0], 1)
            os.dup2(self.saved[1], 2)
            for fd in (self.null,) + self.saved:
                os.close(fd)


DEFAULTS = dict(rngSeed=1234, logLevel=4, timerLevel=4, simBoxLow=[0.0, 0.0, 0.0], simBoxHigh=[1.0, 1.0, 1.0],
                simBoxPBC=[False, False, False], monolayer=False, initPreSteps=0, viscosity=1.0, KBT=-1.0,
                sylinderNumber=0, sylinderLength=0.25, sylinderLengthSigma=0, sylinderDiameter=0.025, sylinderFixed=False,
                sylinderColBuf=0.025, sylinderDiameterColRatio=1.0, sylinderLengthColRatio=1.0, dt=1e-5, timeTotal=1.0,
                timeSnap=1e9, conResTol=1e-5, conMaxIte=10000, conSolverChoice=0, linkKappa=100.0, linkGap=0.01)


def _yaml_value(v):
    if isinstance(v, (bool, np.bool_)):
        return "true" if v else "false"
    if isinstance(v, (list, tuple, np.ndarray)):
        return "[" + ", ".join(_yaml_value(x) for x in v) + "]"
    if isinstance(v, (float, np.floating)):
        return repr(float(v))
    return str(v)


def write_yaml(path, cfg, boundaries=()):
    with open(path, "w") as f:
        for k, v in cfg.items():
            f.write(f"{k}: {_yaml_value(v)}\n")
        if boundaries:
            f.write("boundaries:\n")
            for b in boundaries:
                first = True
                for k, v in b.items():
                    f.write(("  - " if first else "    ") + f"{k}: {_yaml_value(v)}\n")
                    first = False


def write_dat(path, rods, links=()):
    """SylinderAscii format the reference reads (SylinderSystem.cpp:317-344, Sylinder.cpp:101-109)"""
    from .pyoracle import quat_to_dir
    with open(path, "w") as f:
        f.write(f"{len(rods['gid'])}\n0\n")
        for i in range(len(rods["gid"])):
            d = quat_to_dir(rods["quat"][i])
            m = rods["pos"][i] - 0.5 * rods["length"][i] * d
            p = rods["pos"][i] + 0.5 * rods["length"][i] * d
            t = "S" if rods["immovable"][i] else "C"
            f.write(f"{t} {int(rods['gid'][i])} {float(rods['radius'][i])!r} " + " ".join(repr(float(x)) for x in (*m, *p)) + " -1\n")
        for a, b in links:
            f.write(f"L {int(a)} {int(b)}\n")


class RefSystem:
    """The reference's SylinderSystem on one rank.  cfg overrides DEFAULTS (RunConfig.yaml keys)."""

    def __init__(self, cfg=None, pos_file=None, yaml_file=None, boundaries=(), nthreads=1, workdir=None, quiet=True):
        self.L = lib()
        self.quiet = quiet
        self.tmp = tempfile.TemporaryDirectory(prefix="alens_refsys_") if workdir is None else None
        self.workdir = workdir or self.tmp.name
        if yaml_file is None:
            full = dict(DEFAULTS)
            full.update(cfg or {})
            yaml_file = os.path.join(self.workdir, "RunConfig.yaml")
            write_yaml(yaml_file, full, boundaries)
        cwd = os.getcwd()
        try:
            with _Quiet(quiet):
                self.h = self.L.refsys_create(self.workdir.encode(), os.path.abspath(yaml_file).encode(),
                                              (os.path.abspath(pos_file) if pos_file else "").encode(), int(nthreads))
        finally:
            os.chdir(cwd)
        if not self.h:
            raise RuntimeError("refsys_create failed")
        self.h = C.c_void_p(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.L.refsys_destroy(self.h)
            self.h = None
        if self.tmp is not None:
            self.tmp.cleanup()
            self.tmp = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _q(self, fn, *a):
        cwd = os.getcwd()
        os.chdir(self.workdir)
        try:
            with _Quiet(self.quiet):
                return fn(self.h, *a)
        finally:
            os.chdir(cwd)

    # ---- rods
    @property
    def n(self):
        return self.L.refsys_num_rods(self.h)

    def sylinders(self):
        out = np.zeros(max(self.n, 1), dtype=SYLINDER_DTYPE)
        self.L.refsys_get_sylinders(self.h, C.c_void_p(out.ctypes.data))
        return out[:self.n]

    def set_rods(self, rods):
        g = np.ascontiguousarray(rods["gid"], dtype=np.int32)
        p = np.ascontiguousarray(rods["pos"], dtype=np.float64).reshape(-1)
        q = np.ascontiguousarray(rods["quat"], dtype=np.float64).reshape(-1)
        le = np.ascontiguousarray(rods["length"], dtype=np.float64)
        ra = np.ascontiguousarray(rods["radius"], dtype=np.float64)
        im = np.ascontiguousarray(rods.get("immovable", np.zeros(len(g))), dtype=np.uint8)
        self.L.refsys_set_rods(self.h, len(g), _ip(g), _dp(p), _dp(q), _dp(le), _dp(ra), im.ctypes.data_as(C.POINTER(C.c_ubyte)))

    def set_config(self, dt, res, max_ite, choice, mu, kbt=-1.0, colbuf=0.025, dratio=1.0, lratio=1.0, link_kappa=100.0,
                   link_gap=0.01, monolayer=False):
        self.L.refsys_set_config(self.h, C.c_double(dt), C.c_double(res), int(max_ite), int(choice), C.c_double(mu),
                                 C.c_double(kbt), C.c_double(colbuf), C.c_double(dratio), C.c_double(lratio),
                                 C.c_double(link_kappa), C.c_double(link_gap), int(bool(monolayer)))

    def prepare_step(self):
        self._q(self.L.refsys_prepare_step)

    def set_force_nonbrown(self, f):
        f = np.ascontiguousarray(f, dtype=np.float64)
        self.L.refsys_set_force_nonbrown(self.h, _dp(f), int(f.size))

    def set_velocity_nonbrown(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        self.L.refsys_set_velocity_nonbrown(self.h, _dp(v), int(v.size))

    def calc_velocity_brown(self):
        self._q(self.L.refsys_calc_velocity_brown)

    def calc_velocity_noncon(self):
        self._q(self.L.refsys_calc_velocity_noncon)

    def velocities(self):
        n6 = 6 * self.n
        a, b, c = np.zeros(n6), np.zeros(n6), np.zeros(n6)
        self.L.refsys_get_velocity(self.h, _dp(a), _dp(b), _dp(c))
        return dict(velNonCon=a, velBrown=b, velNonBrown=c)

    def mobility_apply(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros_like(x)
        self.L.refsys_mobility_apply(self.h, _dp(x), _dp(y))
        return y

    # ---- constraints
    def constraints(self):
        n = self.L.refsys_num_constraints(self.h)
        out = np.zeros(max(n, 1), dtype=BLOCK_DTYPE)
        self.L.refsys_get_constraints(self.h, C.c_void_p(out.ctypes.data))
        return out[:n]

    def clear_constraints(self):
        self.L.refsys_clear_constraints(self.h)

    def append_constraints(self, blocks):
        b = np.ascontiguousarray(blocks, dtype=BLOCK_DTYPE)
        self.L.refsys_append_constraints(self.h, C.c_void_p(b.ctypes.data), C.c_longlong(len(b)))

    def collect_pair_collision(self):
        return self._q(self.L.refsys_collect_pair_collision)

    def collect_boundary_collision(self):
        return self._q(self.L.refsys_collect_boundary_collision)

    def collect_link_bilateral(self):
        return self._q(self.L.refsys_collect_link_bilateral)

    def add_links(self, prev, nxt):
        p = np.ascontiguousarray(prev, dtype=np.int32)
        q = np.ascontiguousarray(nxt, dtype=np.int32)
        self.L.refsys_add_links(self.h, _ip(p), _ip(q), len(p))

    def resolve_constraints(self):
        self._q(self.L.refsys_resolve_constraints)

    def force_velocity(self):
        n6 = 6 * self.n
        out = {k: np.zeros(n6) for k in ("forceU", "velU", "forceB", "velB")}
        self.L.refsys_get_force_velocity(self.h, _dp(out["forceU"]), _dp(out["velU"]), _dp(out["forceB"]), _dp(out["velB"]))
        return out

    def write_result(self):
        """SylinderSystem::writeResult into <workdir>/result/result0-399/ (returns that folder)"""
        self._q(self.L.refsys_write_result)
        return os.path.join(self.workdir, "result", "result0-399")

    def sum_force_velocity(self):
        self._q(self.L.refsys_sum_force_velocity)

    def step_euler(self):
        self._q(self.L.refsys_step_euler)

    def run_step(self):
        self._q(self.L.refsys_run_step)

    def operator_apply(self, blocks, dt, x):
        """reference ConstraintOperator::apply on the D^T built from `blocks`: returns y, force, vel"""
        b = np.ascontiguousarray(blocks, dtype=BLOCK_DTYPE)
        x = np.ascontiguousarray(x, dtype=np.float64)
        y, f, v = np.zeros(len(b)), np.zeros(6 * self.n), np.zeros(6 * self.n)
        self._q(self.L.refsys_operator_apply, C.c_void_p(b.ctypes.data), C.c_longlong(len(b)), C.c_double(dt), _dp(x), _dp(y),
                _dp(f), _dp(v))
        return y, f, v

    def solve_blocks(self, blocks, vel_nc, dt, res, max_ite, choice, hist_cap=4096):
        """reference ConstraintCollector + ConstraintSolver + BCQPSolver on a given list (needs prepare_step)"""
        b = np.ascontiguousarray(blocks, dtype=BLOCK_DTYPE)
        v = np.ascontiguousarray(vel_nc, dtype=np.float64)
        n, n6 = len(b), 6 * self.n
        assert v.size == n6
        out = np.zeros(max(n, 1), dtype=BLOCK_DTYPE)
        gamma = np.zeros(max(n, 1))
        res_ = {k: np.zeros(n6) for k in ("forceU", "velU", "forceB", "velB")}
        hist = np.zeros((max(hist_cap, 1), 6))
        nh = C.c_int(0)
        rc = self._q(self.L.refsys_solve_blocks, C.c_void_p(b.ctypes.data), C.c_longlong(n), _dp(v), C.c_double(dt),
                     C.c_double(res), int(max_ite), int(choice), C.c_void_p(out.ctypes.data), _dp(gamma),
                     _dp(res_["forceU"]), _dp(res_["velU"]), _dp(res_["forceB"]), _dp(res_["velB"]), _dp(hist),
                     int(hist_cap), C.byref(nh))
        if rc <= -1000:
            raise RuntimeError("refsys_solve_blocks: ConstraintSolver and the direct BCQPSolver run disagree")
        res_.update(gamma=gamma[:n], blocks=out[:n], history=hist[:min(nh.value, hist_cap)].copy(),
                    nIte=(nh.value - 1) if hist_cap > 0 else None, status=rc)
        return res_


def bcqp_solve_csr(rowptr, colind, values, b, lb, ub, x0, tol, max_ite, choice, hist_cap=100000, nthreads=1):
    """reference BCQPSolver (solveBBPGD / solveAPGD) on a CSR matrix, returns (x, history rows, status)"""
    L = lib()
    rp = np.ascontiguousarray(rowptr, dtype=np.int64)
    ci = np.ascontiguousarray(colind, dtype=np.int32)
    va = np.ascontiguousarray(values, dtype=np.float64)
    bb = np.ascontiguousarray(b, dtype=np.float64)
    n = len(bb)
    l_ = None if lb is None else np.ascontiguousarray(lb, dtype=np.float64)
    u_ = None if ub is None else np.ascontiguousarray(ub, dtype=np.float64)
    x = np.array(x0, dtype=np.float64)
    hist = np.zeros((hist_cap, 6))
    nh = C.c_int(0)
    with _Quiet():
        rc = L.refbcqp_solve_csr(n, rp.ctypes.data_as(C.POINTER(C.c_longlong)), _ip(ci), _dp(va), _dp(bb), _dp(l_), _dp(u_),
                                 _dp(x), C.c_double(tol), int(max_ite), int(choice), _dp(hist), int(hist_cap), C.byref(nh),
                                 int(nthreads))
    return x, hist[:min(nh.value, hist_cap)].copy(), rc


def bcqp_selftest(workdir, local_size, diagonal, tol, max_ite, choice):
    """the reference's BCQPSolver(int, double)::selfTest; returns the dumped problem and solution (MatrixMarket files)"""
    L = lib()
    cwd = os.getcwd()
    try:
        with _Quiet():
            rc = L.refbcqp_selftest(workdir.encode(), int(local_size), C.c_double(diagonal), C.c_double(tol), int(max_ite),
                                    int(choice), 1)
    finally:
        os.chdir(cwd)
    assert rc == 0

    def dense(name):
        with open(os.path.join(workdir, name)) as f:
            rows = [ln for ln in f if not ln.startswith("%")]
        return np.array([float(x) for x in rows[1:]])

    with open(os.path.join(workdir, "Amat_TCMAT.mtx")) as f:
        rows = [ln.split() for ln in f if not ln.startswith("%")]
    n = int(rows[0][0])
    A = np.zeros((n, n))
    for r in rows[1:]:
        A[int(r[0]) - 1, int(r[1]) - 1] = float(r[2])
    sol = "xsolAPGD_TV.mtx" if choice == 1 else "xsolBBPGD_TV.mtx"
    return dict(A=A, b=dense("bvec_TV.mtx"), lb=dense("lbvec_TV.mtx"), ub=dense("ubvec_TV.mtx"), x=dense(sol))


def pair_functor(a, b):
    """CalcSylinderNearForce::operator() on one (target, source) pair of oracle ROD_DTYPE records"""
    out = np.zeros(1, dtype=BLOCK_DTYPE)
    pa, da = np.ascontiguousarray(a["pos"], dtype=np.float64), np.ascontiguousarray(a["direction"], dtype=np.float64)
    pb, db = np.ascontiguousarray(b["pos"], dtype=np.float64), np.ascontiguousarray(b["direction"], dtype=np.float64)
    with _Quiet():
        hit = lib().refsys_pair_functor(_dp(pa), _dp(da), C.c_double(a["lengthCollision"]), C.c_double(a["radiusCollision"]),
                                        C.c_double(a["colBuf"]), int(a["gid"]), _dp(pb), _dp(db),
                                        C.c_double(b["lengthCollision"]), C.c_double(b["radiusCollision"]),
                                        C.c_double(b["colBuf"]), int(b["gid"]), C.c_void_p(out.ctypes.data))
    return out[0] if hit else None


def brown_normals(seed, n_rods):
    """the deviates a fresh single-thread TRngPool(seed) hands to calcVelocityBrown, 12 per rod (Wrot, Wpos, Wrfdrot, Wrfdpos)"""
    out = np.zeros(12 * n_rods)
    with _Quiet():
        lib().refsys_brown_normals(int(seed), int(n_rods), _dp(out))
    return out


def collide_stress(dirI, dirJ, cI, cJ, lenI, lenJ, radI, radJ, rho, P, Q):
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (dirI, dirJ, cI, cJ, P, Q)]
    out = np.zeros(9)
    lib().refsys_collide_stress(_dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(a[3]), C.c_double(lenI), C.c_double(lenJ),
                                C.c_double(radI), C.c_double(radJ), C.c_double(rho), _dp(a[4]), _dp(a[5]), _dp(out))
    return out


def drag_coeff(length, radius, mu):
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    lib().refsys_drag_coeff(C.c_double(length), C.c_double(radius), C.c_double(mu), C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


def sylinder_step_euler(pos, quat, vel, omega, dt):
    p, q = np.array(pos, dtype=np.float64), np.array(quat, dtype=np.float64)
    v, w = np.ascontiguousarray(vel, dtype=np.float64), np.ascontiguousarray(omega, dtype=np.float64)
    lib().refsys_sylinder_step_euler(_dp(p), _dp(q), _dp(v), _dp(w), C.c_double(dt))
    return p, q


def boundary_project(kind, center, axis, radius, inside, query):
    c, a, q = (np.ascontiguousarray(x, dtype=np.float64) for x in (center, axis, query))
    proj, delta = np.zeros(3), np.zeros(3)
    lib().refsys_boundary_project({"sphere": 0, "wall": 1, "tube": 2}[kind], _dp(c), _dp(a), C.c_double(radius),
                                  int(bool(inside)), _dp(q), _dp(proj), _dp(delta))
    return proj, delta


def mix_search(trg_pos, trg_rs, src_pos, src_rs, box_low, box_high, pbc, nthreads=1):
    """the reference's MixPairInteraction (SimToolbox/MPI/MixPairInteraction.hpp) driven like MPI/MixPairInteraction_test.cpp:
    every (target, source image) FDPS gives the functor with distance <= max(rsTrg, rsSrc).  Returns (pairs[k, 2], dist[k])
    sorted by (target, source, dist)."""
    L = lib()
    L.refmix_search.restype = C.c_longlong
    tp = np.ascontiguousarray(trg_pos, dtype=np.float64)
    sp = np.ascontiguousarray(src_pos, dtype=np.float64)
    tr = np.ascontiguousarray(trg_rs, dtype=np.float64)
    sr = np.ascontiguousarray(src_rs, dtype=np.float64)
    lo = np.ascontiguousarray(box_low, dtype=np.float64)
    hi = np.ascontiguousarray(box_high, dtype=np.float64)
    pb = np.ascontiguousarray(pbc, dtype=np.int32)
    cap = 64 * (len(tp) + len(sp)) + 1024
    while True:
        pairs = np.zeros((cap, 2), dtype=np.int64)
        dist = np.zeros(cap)
        with _Quiet():
            n = L.refmix_search(len(tp), _dp(tp), _dp(tr), len(sp), _dp(sp), _dp(sr), _dp(lo), _dp(hi), _ip(pb), C.c_longlong(cap),
                                C.c_void_p(pairs.ctypes.data), _dp(dist), int(nthreads))
        if n <= cap:
            break
        cap = int(n)
    pairs, dist = pairs[:n], dist[:n]
    order = np.lexsort((dist, pairs[:, 1], pairs[:, 0]))
    return pairs[order], dist[order]
