"""ctypes loader for the CPU oracle (oracle/liboracle.so) and, when present, the reference build
(oracle/_ref/libalens_ref.so = the reference's own FDPS + DCPQuery compiled in place).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (alens_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

ROD_DTYPE = np.dtype(
    [
        ("gid", "<i4"), ("globalIndex", "<i4"), ("rank", "<i4"), ("pad_", "<i4"),
        ("radius", "<f8"), ("length", "<f8"), ("radiusCollision", "<f8"), ("lengthCollision", "<f8"),
        ("colBuf", "<f8"), ("pos", "<f8", 3), ("direction", "<f8", 3),
    ],
    align=True,
)
BLOCK_DTYPE = np.dtype(
    [
        ("delta0", "<f8"), ("gamma", "<f8"), ("gammaLB", "<f8"),
        ("gidI", "<i4"), ("gidJ", "<i4"), ("globalIndexI", "<i4"), ("globalIndexJ", "<i4"),
        ("oneSide", "u1"), ("bilateral", "u1"), ("pad_", "u1", 6),
        ("kappa", "<f8"),
        ("normI", "<f8", 3), ("normJ", "<f8", 3), ("posI", "<f8", 3), ("posJ", "<f8", 3),
        ("labI", "<f8", 3), ("labJ", "<f8", 3), ("stress", "<f8", 9),
    ],
    align=True,
)
REFPAIR_DTYPE = np.dtype(
    [
        ("gidI", "<i4"), ("gidJ", "<i4"), ("delta0", "<f8"), ("normI", "<f8", 3),
        ("posI", "<f8", 3), ("posJ", "<f8", 3), ("labI", "<f8", 3), ("labJ", "<f8", 3),
    ],
    align=True,
)
HIST_DTYPE = np.dtype([("v", "<f8", 6)])
assert ROD_DTYPE.itemsize == 104 and BLOCK_DTYPE.itemsize == 272


class SolveInfo(C.Structure):
    _fields_ = [
        ("nRods", C.c_int), ("nc", C.c_longlong), ("dt", C.c_double), ("res", C.c_double),
        ("maxIte", C.c_int), ("solverChoice", C.c_int), ("nthreads", C.c_int),
        ("nIte", C.c_int), ("mvCount", C.c_int), ("status", C.c_int),
        ("resFinal", C.c_double), ("tAssemble", C.c_double), ("tSolve", C.c_double),
    ]


class Csr(C.Structure):
    _fields_ = [("n", C.c_int), ("nnz", C.c_longlong), ("rowptr", C.POINTER(C.c_longlong)),
                ("col", C.POINTER(C.c_int)), ("val", C.POINTER(C.c_double))]


def build(ref=True):
    """make liboracle.so (always) and _ref/libalens_ref.so (only when /root/reference exists)."""
    subprocess.check_call(["make", "-s", "-C", HERE])
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


_lib = None
_ref = None


def _p(a, ct=C.c_double):
    return a.ctypes.data_as(C.POINTER(ct))


def _vp(a):
    return C.c_void_p(a.ctypes.data)


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        _lib = C.CDLL(path)
        L = _lib
        L.orc_dcp_segseg.restype = C.c_double
        L.orc_dist_point_seg.restype = C.c_double
        L.orc_collect_pairs_brute.restype = C.c_longlong
        L.orc_collect_pairs_cells.restype = C.c_longlong
        assert L.orc_sizeof_rod() == ROD_DTYPE.itemsize
        assert L.orc_sizeof_block() == BLOCK_DTYPE.itemsize
    return _lib


def have_ref():
    return os.path.exists(os.path.join(HERE, "_ref", "libalens_ref.so"))


def ref():
    global _ref
    if _ref is None:
        _ref = C.CDLL(os.path.join(HERE, "_ref", "libalens_ref.so"))
        _ref.ref_dcp_segseg.restype = C.c_double
        _ref.ref_dist_point_seg.restype = C.c_double
        _ref.ref_fdps_collect.restype = C.c_longlong
        _ref.ref_fdps_last_seconds.restype = C.c_double
        _ref.ref_fdps_last_candidates.restype = C.c_longlong
        assert _ref.ref_sizeof_rod() == ROD_DTYPE.itemsize
        assert _ref.ref_sizeof_pair() == REFPAIR_DTYPE.itemsize
    return _ref


# ----------------------------------------------------------------------------- geometry
def dcp_segseg(P0, P1, Q0, Q1, which="oracle"):
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (P0, P1, Q0, Q1)]
    Pl, Ql = np.zeros(3), np.zeros(3)
    s, t = C.c_double(0), C.c_double(0)
    f = lib().orc_dcp_segseg if which == "oracle" else ref().ref_dcp_segseg
    d = f(_p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(Pl), _p(Ql), C.byref(s), C.byref(t))
    return d, Pl, Ql, s.value, t.value


def dist_point_seg(pt, m, p, which="oracle"):
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (pt, m, p)]
    out = np.zeros(3)
    f = lib().orc_dist_point_seg if which == "oracle" else ref().ref_dist_point_seg
    d = f(_p(a[0]), _p(a[1]), _p(a[2]), _p(out))
    return d, out


def quat_to_dir(q):
    q = np.ascontiguousarray(q, dtype=np.float64)
    d = np.zeros(3)
    lib().orc_quat_to_dir(_p(q), _p(d))
    return d


def make_rods(gid, radius, length, pos, quat, dRatio=1.0, lRatio=1.0, colBuf=0.3, base=0):
    n = len(gid)
    gid = np.ascontiguousarray(gid, dtype=np.int32)
    radius = np.ascontiguousarray(radius, dtype=np.float64)
    length = np.ascontiguousarray(length, dtype=np.float64)
    pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(n, 3)
    quat = np.ascontiguousarray(quat, dtype=np.float64).reshape(n, 4)
    out = np.zeros(n, dtype=ROD_DTYPE)
    lib().orc_make_rods(n, _p(gid, C.c_int), _p(radius), _p(length), _p(pos), _p(quat), C.c_double(dRatio),
                        C.c_double(lRatio), C.c_double(colBuf), C.c_int(base), _vp(out))
    return out


def wrap_positions(pos, lo, hi, pbc=None):
    """applyBoxBC: periodic axes only (pbc=None: all three periodic)."""
    pos = np.array(pos, dtype=np.float64, order="C").reshape(-1, 3)
    lo = np.ascontiguousarray(lo, dtype=np.float64)
    hi = np.ascontiguousarray(hi, dtype=np.float64)
    pb = None if pbc is None else np.ascontiguousarray(pbc, dtype=np.int32)
    lib().orc_wrap_positions(len(pos), _p(pos), _p(lo), _p(hi), None if pb is None else _p(pb, C.c_int))
    return pos


def pair_functor(a, b, with_stress=True, which="oracle"):
    """a, b: 1-element ROD_DTYPE arrays (target, source).  Returns a BLOCK (or REFPAIR) record or None."""
    a = np.ascontiguousarray(a, dtype=ROD_DTYPE).reshape(1)
    b = np.ascontiguousarray(b, dtype=ROD_DTYPE).reshape(1)
    if which == "oracle":
        out = np.zeros(1, dtype=BLOCK_DTYPE)
        hit = lib().orc_pair_functor(_vp(a), _vp(b), int(with_stress), _vp(out))
    else:
        out = np.zeros(1, dtype=REFPAIR_DTYPE)
        hit = ref().ref_pair_block(_vp(a), _vp(b), _vp(out))
    return out[0] if hit else None


def collect_pairs(rods, lo, hi, pbc, with_stress=False, method="cells", nthreads=0, cap=None):
    rods = np.ascontiguousarray(rods, dtype=ROD_DTYPE)
    lo = np.ascontiguousarray(lo, dtype=np.float64)
    hi = np.ascontiguousarray(hi, dtype=np.float64)
    pbc = np.ascontiguousarray(pbc, dtype=np.int32)
    n = len(rods)
    cap = cap or max(1024, 8 * n)
    while True:
        out = np.zeros(cap, dtype=BLOCK_DTYPE)
        if method == "brute":
            cnt = lib().orc_collect_pairs_brute(n, _vp(rods), _p(lo), _p(hi), _p(pbc, C.c_int), int(with_stress),
                                                _vp(out), C.c_longlong(cap))
        else:
            cnt = lib().orc_collect_pairs_cells(n, _vp(rods), _p(lo), _p(hi), _p(pbc, C.c_int), int(with_stress),
                                                _vp(out), C.c_longlong(cap), int(nthreads))
        if cnt <= cap:
            return out[:cnt].copy()
        cap = int(cnt)


BOUNDARY_DTYPE = np.dtype([("type", "<i4"), ("inside", "<i4"), ("center", "<f8", 3), ("axis", "<f8", 3), ("radius", "<f8")],
                          align=True)


def make_boundaries(specs):
    """specs: list of dicts {type: 'sphere'|'wall'|'tube', center, axis (wall normal / tube axis), radius, inside};
    axes are normalised as the reference's constructors do (Boundary.cpp:96-100, :166-172)"""
    out = np.zeros(len(specs), dtype=BOUNDARY_DTYPE)
    for o, s in zip(out, specs):
        o["type"] = {"sphere": 0, "wall": 1, "tube": 2}[s["type"]]
        o["inside"] = 1 if s.get("inside", True) else 0
        o["center"] = s["center"]
        a = np.asarray(s.get("axis", [0.0, 0.0, 1.0]), dtype=np.float64)
        o["axis"] = a / np.sqrt((a * a).sum())
        o["radius"] = s.get("radius", 0.0)
    return out


def boundary_project(bnd, query):
    b = np.ascontiguousarray(bnd, dtype=BOUNDARY_DTYPE).reshape(1)
    q = np.ascontiguousarray(query, dtype=np.float64)
    proj, delta = np.zeros(3), np.zeros(3)
    lib().orc_boundary_project(_vp(b), _p(q), _p(proj), _p(delta))
    return proj, delta


def collect_boundary(rods, boundaries, col_buf):
    rods = np.ascontiguousarray(rods, dtype=ROD_DTYPE)
    bnd = np.ascontiguousarray(boundaries, dtype=BOUNDARY_DTYPE)
    cap = max(16, 2 * len(rods) * max(len(bnd), 1))
    out = np.zeros(cap, dtype=BLOCK_DTYPE)
    L = lib()
    L.orc_collect_boundary.restype = C.c_longlong
    cnt = L.orc_collect_boundary(len(rods), _vp(rods), len(bnd), _vp(bnd), C.c_double(col_buf), _vp(out), C.c_longlong(cap))
    return out[:cnt].copy()


def collect_links(rods, prev_gid, next_gid, lo, hi, pbc, link_kappa, link_gap):
    rods = np.ascontiguousarray(rods, dtype=ROD_DTYPE)
    prev_gid = np.ascontiguousarray(prev_gid, dtype=np.int32)
    next_gid = np.ascontiguousarray(next_gid, dtype=np.int32)
    lo = np.ascontiguousarray(lo, dtype=np.float64)
    hi = np.ascontiguousarray(hi, dtype=np.float64)
    pbc = np.ascontiguousarray(pbc, dtype=np.int32)
    out = np.zeros(max(len(prev_gid), 1), dtype=BLOCK_DTYPE)
    L = lib()
    L.orc_collect_links.restype = C.c_longlong
    cnt = L.orc_collect_links(len(rods), _vp(rods), C.c_longlong(len(prev_gid)), _p(prev_gid, C.c_int),
                              _p(next_gid, C.c_int), _p(lo), _p(hi), _p(pbc, C.c_int), C.c_double(link_kappa),
                              C.c_double(link_gap), _vp(out))
    if cnt < 0:
        raise ValueError("collect_links: unknown gid")
    return out[:cnt].copy()


def fdps_collect(rods, lo, hi, pbc, nthreads=1, rebuild=True):
    """The reference's own FDPS neighbour search + functor (oracle/_ref).  Returns (pairs, rods_wrapped)."""
    rods = np.array(rods, dtype=ROD_DTYPE, order="C")
    lo = np.ascontiguousarray(lo, dtype=np.float64)
    hi = np.ascontiguousarray(hi, dtype=np.float64)
    pbc = np.ascontiguousarray(pbc, dtype=np.int32)
    R = ref()
    # FDPS prints its banner on stderr at first Initialize(); keep test logs quiet
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(2)
    os.dup2(devnull, 2)
    try:
        cnt = R.ref_fdps_collect(len(rods), _vp(rods), _p(lo), _p(hi), _p(pbc, C.c_int), int(nthreads), int(rebuild))
    finally:
        os.dup2(saved, 2)
        os.close(devnull)
        os.close(saved)
    out = np.zeros(max(cnt, 1), dtype=REFPAIR_DTYPE)
    R.ref_fdps_get(_vp(out))
    out = out[:cnt]
    order = np.lexsort((out["gidJ"], out["gidI"]))
    return out[order].copy(), rods


def fdps_last_seconds():
    return ref().ref_fdps_last_seconds()


# ----------------------------------------------------------------------------- assembly / solve
def drag_coeff(radius, length, mu):
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    lib().orc_drag_coeff(C.c_double(radius), C.c_double(length), C.c_double(mu), C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


def velocity_brown(quat, radius, length, immovable, mu, kbt, dt, normals12):
    """SylinderSystem::calcVelocityBrown (SylinderSystem.cpp:1020-1091) for given N(0,1) draws (12 per rod, in the
    reference's draw order Wrot, Wpos, Wrfdrot, Wrfdpos), plain numpy per rod.  Returns 6n (velBrown, omegaBrown)."""
    quat = np.asarray(quat, dtype=np.float64)
    W = np.asarray(normals12, dtype=np.float64).reshape(-1, 12)
    n = len(quat)
    out = np.zeros((n, 6))
    delta, fac = dt * 0.1, np.sqrt(2 * kbt / dt)

    def rot_z(q):  # Eigen: q * (0,0,1), q = (x, y, z, w)
        x, y, z, w = q
        return np.array([2 * (w * y + x * z), 2 * (y * z - w * x), 1 - 2 * (x * x + y * y)])

    def rotate(q, omega, t):  # EquatnHelper::rotateEquatn (Util/EquatnHelper.hpp:74-90)
        w = np.linalg.norm(omega)
        if w < np.finfo(np.float32).eps:
            return q
        sw, cw, p, s = np.sin(w * t / 2), np.cos(w * t / 2), q[:3], q[3]
        xyz = s * sw * omega / w + cw * p + sw / w * np.cross(omega, p)
        qw = s * cw - p.dot(omega) * sw / w
        qn = np.concatenate([xyz, [qw]])
        return qn / np.linalg.norm(qn)

    for i in range(n):
        zp, zq, zr = drag_coeff(float(radius[i]), float(length[i]), mu)
        a, b, c = (0.0, 0.0, 0.0) if immovable[i] else (1 / zp, 1 / zq, 1 / zr)
        d = rot_z(quat[i])
        N = (a - b) * np.outer(d, d) + b * np.eye(3)
        L = np.linalg.cholesky(N) if not immovable[i] else np.zeros((3, 3))  # Eigen's LLT leaves a zero matrix untouched
        Wrot, Wpos, Wrr, Wrp = W[i, 0:3], W[i, 3:6], W[i, 6:9], W[i, 9:12]
        dr = rot_z(rotate(quat[i].copy(), Wrr, delta))
        Nr = (a - b) * np.outer(dr, dr) + b * np.eye(3)
        out[i, :3] = fac * (L @ Wpos) + (kbt / delta) * ((Nr - N) @ Wrp)
        out[i, 3:] = np.sqrt(c) * fac * Wrot
    return out.reshape(-1)


def build_dtrans_dense(blocks, n_rods):
    """D^T as a scipy CSR (tests only)."""
    import scipy.sparse as sp

    blocks = np.ascontiguousarray(blocks, dtype=BLOCK_DTYPE)
    nc = len(blocks)
    csr = Csr()
    d0, ik, bi, g0 = (np.zeros(nc + 1) for _ in range(4))
    lib().orc_build_dtrans(C.c_longlong(nc), _vp(blocks), n_rods, C.byref(csr), _p(d0), _p(ik), _p(bi), _p(g0))
    rowptr = np.ctypeslib.as_array(csr.rowptr, (nc + 1,)).copy()
    nnz = int(csr.nnz)
    col = np.ctypeslib.as_array(csr.col, (max(nnz, 1),))[:nnz].copy()
    val = np.ctypeslib.as_array(csr.val, (max(nnz, 1),))[:nnz].copy()
    lib().orc_csr_free(C.byref(csr))
    return sp.csr_matrix((val, col, rowptr), shape=(nc, 6 * n_rods)), d0[:nc], ik[:nc], bi[:nc], g0[:nc]


def build_mobility(rods, immovable, mu):
    import scipy.sparse as sp

    rods = np.ascontiguousarray(rods, dtype=ROD_DTYPE)
    n = len(rods)
    imm = np.ascontiguousarray(immovable, dtype=np.int32)
    csr = Csr()
    lib().orc_build_mobility(n, _vp(rods), _p(imm, C.c_int), C.c_double(mu), C.byref(csr))
    rowptr = np.ctypeslib.as_array(csr.rowptr, (6 * n + 1,)).copy()
    col = np.ctypeslib.as_array(csr.col, (18 * n,)).copy()
    val = np.ctypeslib.as_array(csr.val, (18 * n,)).copy()
    lib().orc_csr_free(C.byref(csr))
    return sp.csr_matrix((val, col, rowptr), shape=(6 * n, 6 * n))


def operator_apply(blocks, rods, immovable, mu, dt, x):
    blocks = np.ascontiguousarray(blocks, dtype=BLOCK_DTYPE)
    rods = np.ascontiguousarray(rods, dtype=ROD_DTYPE)
    imm = np.ascontiguousarray(immovable, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float64)
    nc, n = len(blocks), len(rods)
    y = np.zeros(nc + 1)
    f = np.zeros(6 * n + 1)
    v = np.zeros(6 * n + 1)
    lib().orc_operator_apply(_vp(blocks), C.c_longlong(nc), _vp(rods), _p(imm, C.c_int), n, C.c_double(mu),
                             C.c_double(dt), _p(x), _p(y), _p(f), _p(v))
    return y[:nc], f[:6 * n], v[:6 * n]


def solve_constraints(blocks, rods, immovable, mu, vel_nc, dt, res, max_ite, solver_choice=0, nthreads=0,
                      hist_cap=None):
    blocks = np.ascontiguousarray(blocks, dtype=BLOCK_DTYPE)
    rods = np.ascontiguousarray(rods, dtype=ROD_DTYPE)
    imm = np.ascontiguousarray(immovable, dtype=np.int32)
    vel_nc = np.ascontiguousarray(vel_nc, dtype=np.float64)
    nc, n = len(blocks), len(rods)
    info = SolveInfo(nRods=n, nc=nc, dt=dt, res=res, maxIte=max_ite, solverChoice=solver_choice, nthreads=nthreads)
    gamma = np.zeros(nc + 1)
    fu, vu, fb, vb = (np.zeros(6 * n + 1) for _ in range(4))
    hist_cap = hist_cap or (max_ite + 2)
    hist = np.zeros(hist_cap, dtype=HIST_DTYPE)
    nh = C.c_int(0)
    rc = lib().orc_solve_constraints(_vp(blocks), _vp(rods), _p(imm, C.c_int), C.c_double(mu), _p(vel_nc),
                                     C.byref(info), _p(gamma), _p(fu), _p(vu), _p(fb), _p(vb), _vp(hist), hist_cap,
                                     C.byref(nh))
    return dict(rc=rc, gamma=gamma[:nc], forceU=fu[:6 * n], velU=vu[:6 * n], forceB=fb[:6 * n], velB=vb[:6 * n],
                history=hist["v"][:min(nh.value, hist_cap)].copy(), nIte=info.nIte, mvCount=info.mvCount,
                resFinal=info.resFinal, tAssemble=info.tAssemble, tSolve=info.tSolve)


def bcqp_csr(A, b, lb, ub, x0, tol, max_ite, solver_choice=0):
    """Generic BCQP on an explicit scipy CSR (BCQPSolver_verify.py-style cross-check)."""
    A = A.tocsr()
    n = A.shape[0]
    rowptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
    col = np.ascontiguousarray(A.indices, dtype=np.int32)
    val = np.ascontiguousarray(A.data, dtype=np.float64)
    csr = Csr(n=n, nnz=len(val), rowptr=_p(rowptr, C.c_longlong), col=_p(col, C.c_int), val=_p(val))
    b, lb, ub = (np.ascontiguousarray(v, dtype=np.float64) for v in (b, lb, ub))
    x = np.array(x0, dtype=np.float64)
    hist = np.zeros(max_ite + 2, dtype=HIST_DTYPE)
    nh = C.c_int(0)
    rc = lib().orc_bcqp_csr(C.byref(csr), _p(b), _p(lb), _p(ub), _p(x), C.c_double(tol), int(max_ite),
                            int(solver_choice), _vp(hist), len(hist), C.byref(nh))
    return rc, x, hist["v"][:nh.value].copy()


def writeback_gamma(blocks, gamma):
    blocks = np.array(blocks, dtype=BLOCK_DTYPE, order="C")
    gamma = np.ascontiguousarray(gamma, dtype=np.float64)
    lib().orc_writeback_gamma(C.c_longlong(len(blocks)), _vp(blocks), _p(gamma))
    return blocks
