"""ctypes binding of include/alens_b200.h (one method per entry point, same names minus the prefix)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

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
assert BLOCK_DTYPE.itemsize == 272

BOUNDARY_DTYPE = np.dtype([("type", "<i4"), ("inside", "<i4"), ("center", "<f8", 3), ("axis", "<f8", 3), ("radius", "<f8")],
                          align=True)

# every symbol include/alens_b200.h declares (tests check that the library exports all of them)
EXPORTS = [
    "alens_create", "alens_destroy", "alens_last_error", "alens_version", "alens_set_stream",
    "alens_set_domain", "alens_set_collision_params", "alens_set_rods", "alens_set_rods_aos",
    "alens_get_positions", "alens_collect_pair_collision", "alens_append_constraints",
    "alens_clear_constraints", "alens_num_constraints", "alens_get_constraints", "alens_calc_mobility",
    "alens_mobility_apply", "alens_solve_constraints", "alens_setup_constraints", "alens_operator_apply",
    "alens_get_history", "alens_get_gamma", "alens_get_force_velocity", "alens_step_euler",
    "alens_get_rod_state", "alens_get_timers", "alens_reset_timers", "alens_get_collect_stats",
    "alens_set_decomposition", "alens_comm_create", "alens_comm_blob_size", "alens_comm_export",
    "alens_comm_connect", "alens_comm_connect_local", "alens_num_ghosts", "alens_prepare_step", "alens_set_velocity_noncon",
    "alens_set_velocity_noncon_async", "alens_collect_boundary_collision", "alens_collect_link_bilateral", "alens_calc_velocity_noncon", "alens_calc_velocity_brown",
    "alens_set_profiling", "alens_bcqp_solve", "alens_set_option", "alens_time_kernel",
    "alens_dcp_query", "alens_pair_functor", "alens_comm_mode", "alens_constraint_digest", "alens_sum_constraint_stress",
    "alens_get_live_stats", "alens_bcqp_create_csr", "alens_bcqp_create_constraint", "alens_bcqp_set_lower_bound",
    "alens_bcqp_set_upper_bound", "alens_bcqp_get_bounds", "alens_bcqp_run", "alens_bcqp_history", "alens_bcqp_size",
    "alens_bcqp_destroy", "alens_collect_protein_bilateral", "alens_get_pool_stats", "alens_get_stamps", "alens_mix_pair_search", "alens_migrate_rods", "alens_get_rod_identity", "alens_set_rod_state", "alens_get_long_rod_stats", "alens_set_rod_tags", "alens_get_rod_tags",
]


def comm_connect_local(contexts):
    """single-process bootstrap: contexts in rank order, each driven by its own host thread afterwards"""
    lib = contexts[0].lib
    arr = (C.c_void_p * len(contexts))(*[c.h for c in contexts])
    rc = lib.dll.alens_comm_connect_local(arr, C.c_int(len(contexts)))
    if rc != 0:
        raise AlensError(rc, lib.dll.alens_last_error(contexts[0].h).decode())


class SolveReport(C.Structure):
    _fields_ = [("status", C.c_int), ("iterations", C.c_int), ("matvecs", C.c_int), ("history_rows", C.c_int),
                ("residual", C.c_double), ("step", C.c_double), ("n_constraints", C.c_longlong),
                ("n_rods", C.c_int)]


class Timers(C.Structure):
    _fields_ = [("upload_ms", C.c_double), ("collect_ms", C.c_double), ("setup_ms", C.c_double),
                ("solve_ms", C.c_double), ("split_ms", C.c_double), ("download_ms", C.c_double),
                ("op_force_vel_ms", C.c_double), ("op_dtrans_ms", C.c_double), ("op_update_ms", C.c_double),
                ("op_force_vel_n", C.c_longlong), ("op_dtrans_n", C.c_longlong), ("op_update_n", C.c_longlong),
                ("op_launches", C.c_longlong), ("total_launches", C.c_longlong),
                ("op_rows_live", C.c_longlong), ("op_applies", C.c_longlong)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class AlensError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"alens_b200 error {code}: {msg}")
        self.code = code


def lib_path():
    return os.path.join(HERE, "libalens_b200.so")


def build(verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... (alens_b200/csrc/Makefile); cross-compiles without a GPU."""
    cmd = ["make", "-C", os.path.join(HERE, "csrc"), "-j4"]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)
    return lib_path()


_EMPTY = np.zeros(1)


def _dp(a):
    if a is None:
        return None
    if a.size == 0:  # a zero-length array still means "given" (non-NULL) across the C ABI
        a = _EMPTY
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Library:
    """The loaded shared library.  Fails loudly if the CUDA extension has not been built."""

    _inst = None

    def __init__(self, path=None):
        path = path or lib_path()
        if not os.path.exists(path):
            raise AlensError(-100, f"{path} not found: build it with alens_b200.build() / __graft_entry__.build(); "
                             "there is no CPU fallback")
        self.path = path
        self.dll = C.CDLL(path)
        d = self.dll
        d.alens_last_error.restype = C.c_char_p
        d.alens_last_error.argtypes = [C.c_void_p]
        d.alens_version.restype = C.c_char_p
        d.alens_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        d.alens_destroy.argtypes = [C.c_void_p]
        d.alens_destroy.restype = None

    @classmethod
    def get(cls):
        if cls._inst is None:
            cls._inst = Library()
        return cls._inst

    def version(self):
        return self.dll.alens_version().decode()


PROTEIN_DTYPE = np.dtype([("idBind", "<i4", 2), ("indexBind", "<i4", 2), ("centerBind", "<f8", (2, 3)),
                          ("directionBind", "<f8", (2, 3)), ("posEndBind", "<f8", (2, 3)), ("lenBind", "<f8", 2),
                          ("forceLength", "<f8"), ("freeLength", "<f8"), ("kappa", "<f8")])
assert PROTEIN_DTYPE.itemsize == 200


class Bcqp:
    """alens_bcqp: BCQPSolver for any caller (a CSR matrix, or the constraint operator of `ctx`'s last setup), with
    caller-set bounds.  Mirrors SimToolbox/Constraint/BCQPSolver.hpp:37-111."""

    def __init__(self, ctx, b=None, csr=None):
        self.ctx, self.dll = ctx, ctx.lib.dll
        self.h = C.c_void_p()
        bb = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
        if csr is None:
            rc = self.dll.alens_bcqp_create_constraint(ctx.h, _dp(bb), C.byref(self.h))
        else:
            rowptr, col, val = csr
            rp = np.ascontiguousarray(rowptr, dtype=np.int64)
            ci = np.ascontiguousarray(col, dtype=np.int32)
            va = np.ascontiguousarray(val, dtype=np.float64)
            rc = self.dll.alens_bcqp_create_csr(ctx.h, C.c_int(len(rp) - 1), rp.ctypes.data_as(C.POINTER(C.c_longlong)),
                                                ci.ctypes.data_as(C.POINTER(C.c_int)), _dp(va), _dp(bb), C.byref(self.h))
        ctx._ck(rc)
        self.dll.alens_bcqp_size.argtypes = [C.c_void_p]
        self.n = self.dll.alens_bcqp_size(self.h)

    def set_bounds(self, lb=None, ub=None):
        for name, v in (("alens_bcqp_set_lower_bound", lb), ("alens_bcqp_set_upper_bound", ub)):
            a = None if v is None else np.ascontiguousarray(v, dtype=np.float64)
            self.ctx._ck(getattr(self.dll, name)(self.h, _dp(a)))

    def get_bounds(self):
        lb, ub = np.zeros(max(self.n, 1)), np.zeros(max(self.n, 1))
        self.ctx._ck(self.dll.alens_bcqp_get_bounds(self.h, _dp(lb), _dp(ub)))
        return lb[:self.n], ub[:self.n]

    def solve(self, x0, tol, max_ite, solver_choice=0):
        x = np.array(x0, dtype=np.float64)
        if len(x) == 0:
            x = np.zeros(1)
        rep = SolveReport()
        rc = self.dll.alens_bcqp_run(self.h, _dp(x), C.c_double(tol), C.c_int(max_ite), C.c_int(solver_choice), C.byref(rep))
        rows = np.zeros((max(rep.history_rows, 1), 6))
        n = C.c_int(0)
        self.dll.alens_bcqp_history(self.h, _dp(rows), C.c_int(len(rows)), C.byref(n))
        self.ctx._ck(rc)
        return x[:self.n], rep, rows[:min(n.value, len(rows))]

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.dll.alens_bcqp_destroy.argtypes = [C.c_void_p]
            self.dll.alens_bcqp_destroy.restype = None
            self.dll.alens_bcqp_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One alens_ctx (one GPU).  Method names follow the C entry points."""

    def __init__(self, device=0, rank=0, nranks=1, library=None):
        self.lib = library or Library.get()
        self.h = C.c_void_p()
        rc = self.lib.dll.alens_create(device, rank, nranks, C.byref(self.h))
        if rc != 0:
            raise AlensError(rc, self.lib.dll.alens_last_error(None).decode())
        self.n_rods = 0

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.dll.alens_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise AlensError(rc, self.lib.dll.alens_last_error(self.h).decode())

    def _call(self, name, *args):
        self._ck(getattr(self.lib.dll, name)(self.h, *args))

    # ---- configuration
    def set_stream(self, cuda_stream):
        self._call("alens_set_stream", C.c_void_p(cuda_stream))

    def set_domain(self, lo, hi, pbc):
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        pbc = np.ascontiguousarray(pbc, dtype=np.int32)
        self._call("alens_set_domain", _dp(lo), _dp(hi), pbc.ctypes.data_as(C.POINTER(C.c_int)))

    def set_collision_params(self, diameter_col_ratio=1.0, length_col_ratio=1.0, col_buf=0.0):
        self._call("alens_set_collision_params", C.c_double(diameter_col_ratio), C.c_double(length_col_ratio),
                   C.c_double(col_buf))

    # ---- rods
    def set_rods(self, gid, pos, orientation, length, radius, immovable=None, wrap=True):
        n = len(gid)
        self._keep = [np.ascontiguousarray(gid, dtype=np.int32),
                      np.ascontiguousarray(pos, dtype=np.float64).reshape(-1),
                      np.ascontiguousarray(orientation, dtype=np.float64).reshape(-1),
                      np.ascontiguousarray(length, dtype=np.float64),
                      np.ascontiguousarray(radius, dtype=np.float64),
                      None if immovable is None else np.ascontiguousarray(immovable, dtype=np.uint8)]
        g, p, q, le, ra, im = self._keep
        assert p.size == 3 * n and q.size == 4 * n and le.size == n and ra.size == n
        self._call("alens_set_rods", C.c_int(n), g.ctypes.data_as(C.POINTER(C.c_int)), _dp(p), _dp(q), _dp(le),
                   _dp(ra), None if im is None else im.ctypes.data_as(C.POINTER(C.c_ubyte)), C.c_int(1 if wrap else 0))
        self.n_rods = n

    def set_rods_raw(self, n, gid_p, pos_p, quat_p, len_p, rad_p, imm_p, wrap=True):
        """pointer version (pinned host buffers owned by the caller, e.g. torch pinned tensors)"""
        self._call("alens_set_rods", C.c_int(n), C.c_void_p(gid_p), C.c_void_p(pos_p), C.c_void_p(quat_p),
                   C.c_void_p(len_p), C.c_void_p(rad_p), C.c_void_p(imm_p), C.c_int(1 if wrap else 0))
        self.n_rods = n

    def set_rods_aos(self, records, stride=568, wrap=True):
        buf = np.ascontiguousarray(records, dtype=np.uint8)
        n = buf.size // stride
        self._call("alens_set_rods_aos", C.c_int(n), C.c_void_p(buf.ctypes.data), C.c_size_t(stride),
                   C.c_int(1 if wrap else 0))
        self.n_rods = n

    def set_rod_state(self, pos, quat, wrap=True):
        """positions + orientations of the resident rod set (everything else unchanged since set_rods)"""
        p = np.ascontiguousarray(pos, dtype=np.float64)
        q = np.ascontiguousarray(quat, dtype=np.float64)
        assert len(p.reshape(-1, 3)) == self.n_rods and len(q.reshape(-1, 4)) == self.n_rods
        self._call("alens_set_rod_state", _dp(p), _dp(q), C.c_int(1 if wrap else 0))

    def set_rod_state_raw(self, pos_p, quat_p, wrap=True):
        self._call("alens_set_rod_state", C.c_void_p(pos_p), C.c_void_p(quat_p), C.c_int(1 if wrap else 0))

    def migrate_rods(self):
        """collective: rods that left the slab move to the neighbour rank; returns (sent, received)"""
        a, b = C.c_longlong(0), C.c_longlong(0)
        self._call("alens_migrate_rods", C.byref(a), C.byref(b))
        n, base = C.c_int(0), C.c_int(0)
        self._call("alens_get_rod_identity", C.byref(n), C.byref(base), None, None, None, None)
        self.n_rods = n.value
        return a.value, b.value

    def set_rod_tags(self, tags):
        t = None if tags is None else np.ascontiguousarray(tags, dtype=np.int64)
        self._call("alens_set_rod_tags", None if t is None else t.ctypes.data_as(C.POINTER(C.c_longlong)))

    def get_rod_tags(self):
        t = np.zeros(max(self.n_rods, 1), dtype=np.int64)
        self._call("alens_get_rod_tags", t.ctypes.data_as(C.POINTER(C.c_longlong)))
        return t[:self.n_rods]

    def get_rod_identity(self):
        """(globalIndexBase, gid, length, radius, immovable) of the resident owned rods"""
        n, base = C.c_int(0), C.c_int(0)
        self._call("alens_get_rod_identity", C.byref(n), C.byref(base), None, None, None, None)
        self.n_rods = n.value
        m = max(n.value, 1)
        gid, le, ra, im = np.zeros(m, dtype=np.int32), np.zeros(m), np.zeros(m), np.zeros(m, dtype=np.uint8)
        self._call("alens_get_rod_identity", C.byref(n), C.byref(base), gid.ctypes.data_as(C.POINTER(C.c_int)), _dp(le), _dp(ra),
                   im.ctypes.data_as(C.POINTER(C.c_ubyte)))
        return base.value, gid[:n.value], le[:n.value], ra[:n.value], im[:n.value]

    def mix_pair_search(self, trg_pos, trg_rs, src_rs=None):
        """targets x resident rods within max(rs_t, rs_j): returns (rowPtr[n+1], local rod indices)"""
        tp = np.ascontiguousarray(trg_pos, dtype=np.float64).reshape(-1, 3)
        tr = np.ascontiguousarray(trg_rs, dtype=np.float64)
        sr = None if src_rs is None else np.ascontiguousarray(src_rs, dtype=np.float64)
        n = len(tp)
        row = np.zeros(n + 1, dtype=np.int64)
        tot = C.c_longlong(0)
        self._call("alens_mix_pair_search", C.c_longlong(n), _dp(tp), _dp(tr), _dp(sr), row.ctypes.data_as(C.POINTER(C.c_longlong)),
                   None, C.c_longlong(0), C.byref(tot))
        idx = np.zeros(max(tot.value, 1), dtype=np.int32)
        self._call("alens_mix_pair_search", C.c_longlong(n), _dp(tp), _dp(tr), _dp(sr), row.ctypes.data_as(C.POINTER(C.c_longlong)),
                   idx.ctypes.data_as(C.POINTER(C.c_int)), C.c_longlong(len(idx)), C.byref(tot))
        return row, idx[:tot.value]

    # ---- the narrow phase by itself
    def dcp_query(self, P0, P1, Q0, Q1):
        """DCPQuery on n segment pairs (n x 3 arrays): returns dist[n], Ploc[n,3], Qloc[n,3]"""
        a = [np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 3) for x in (P0, P1, Q0, Q1)]
        n = len(a[0])
        dist, P, Q = np.zeros(n), np.zeros((n, 3)), np.zeros((n, 3))
        self._call("alens_dcp_query", C.c_longlong(n), _dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(a[3]), _dp(dist), _dp(P), _dp(Q))
        return dist, P, Q

    def pair_functor(self, geomI, geomJ, with_stress=True):
        """the functor body on n (I, J) pairs; geom = n x 9 {pos, direction, lengthCollision, radiusCollision, colBuf}"""
        gi = np.ascontiguousarray(geomI, dtype=np.float64).reshape(-1, 9)
        gj = np.ascontiguousarray(geomJ, dtype=np.float64).reshape(-1, 9)
        n = len(gi)
        hit = np.zeros(max(n, 1), dtype=np.uint8)
        blocks = np.zeros(max(n, 1), dtype=BLOCK_DTYPE)
        self._call("alens_pair_functor", C.c_longlong(n), _dp(gi), _dp(gj), C.c_int(1 if with_stress else 0),
                   hit.ctypes.data_as(C.POINTER(C.c_ubyte)), C.c_void_p(blocks.ctypes.data))
        return hit[:n].astype(bool), blocks[:n]

    def prepare_step(self, wrap=True):
        self._call("alens_prepare_step", C.c_int(1 if wrap else 0))

    def set_velocity_noncon(self, v):
        v = None if v is None else np.ascontiguousarray(v, dtype=np.float64)
        self._call("alens_set_velocity_noncon", _dp(v))

    def collect_boundary_collision(self, boundaries):
        """boundaries: structured array with the fields of alens_boundary (type, inside, center[3], axis[3], radius)"""
        b = np.ascontiguousarray(boundaries, dtype=BOUNDARY_DTYPE)
        n = C.c_longlong(0)
        self._call("alens_collect_boundary_collision", C.c_void_p(b.ctypes.data), C.c_int(len(b)), C.byref(n))
        return n.value

    def calc_velocity_brown(self, kbt, dt, normals12=None, seed=0, step=0):
        w = None if normals12 is None else np.ascontiguousarray(normals12, dtype=np.float64)
        out = np.zeros(6 * self.n_rods)
        self._call("alens_calc_velocity_brown", C.c_double(kbt), C.c_double(dt), _dp(w), C.c_ulonglong(seed),
                   C.c_ulonglong(step), _dp(out))
        return out

    def calc_velocity_noncon(self, force_nonbrown=None, vel_nonbrown=None, vel_brown=None, monolayer=False):
        """velNonCon = M f + vNB + vB on the device, kept resident; returns M f + vNB (Sylinder::velNonB / omegaNonB)"""
        a = [None if x is None else np.ascontiguousarray(x, dtype=np.float64) for x in (force_nonbrown, vel_nonbrown, vel_brown)]
        out = np.zeros(6 * self.n_rods)
        self._call("alens_calc_velocity_noncon", _dp(a[0]), _dp(a[1]), _dp(a[2]), C.c_int(1 if monolayer else 0), _dp(out))
        return out

    def collect_protein_bilateral(self, proteins, tubule_diameter):
        """proteins: structured array PROTEIN_DTYPE (alens_protein_bind); returns the number of blocks added"""
        p = np.ascontiguousarray(proteins, dtype=PROTEIN_DTYPE)
        n = C.c_longlong(0)
        self._call("alens_collect_protein_bilateral", C.c_void_p(p.ctypes.data), C.c_longlong(len(p)),
                   C.c_double(tubule_diameter), C.byref(n))
        return n.value

    def collect_link_bilateral(self, prev_gid, next_gid, link_kappa, link_gap):
        p = np.ascontiguousarray(prev_gid, dtype=np.int32)
        q = np.ascontiguousarray(next_gid, dtype=np.int32)
        n = C.c_longlong(0)
        self._call("alens_collect_link_bilateral", p.ctypes.data_as(C.POINTER(C.c_int)), q.ctypes.data_as(C.POINTER(C.c_int)),
                   C.c_longlong(len(p)), C.c_double(link_kappa), C.c_double(link_gap), C.byref(n))
        return n.value

    def set_velocity_noncon_async_raw(self, v_p):
        """pinned host pointer; the copy overlaps the calls that follow (see alens_b200.h)"""
        self._call("alens_set_velocity_noncon_async", C.c_void_p(v_p))

    def set_profiling(self, on):
        self._call("alens_set_profiling", C.c_int(1 if on else 0))

    def set_option(self, name, value):
        self._call("alens_set_option", C.c_char_p(name.encode()), C.c_longlong(int(value)))

    # ---- multi-GPU
    def set_decomposition(self, axis, slab_lo, slab_hi, skin, max_bounding_radius, global_base=0):
        self._call("alens_set_decomposition", C.c_int(axis), C.c_double(slab_lo), C.c_double(slab_hi),
                   C.c_double(skin), C.c_double(max_bounding_radius), C.c_int(global_base))

    def comm_create(self, max_local_rods):
        self._call("alens_comm_create", C.c_longlong(int(max_local_rods)))

    def comm_export(self):
        n = self.lib.dll.alens_comm_blob_size()
        buf = C.create_string_buffer(n)
        self._call("alens_comm_export", buf)
        return buf.raw

    def comm_connect(self, blobs):
        """blobs: list of bytes in rank order (e.g. from torch.distributed.all_gather_object)"""
        self._call("alens_comm_connect", C.c_char_p(b"".join(blobs)))

    def num_ghosts(self):
        a, b, c = C.c_int(0), C.c_int(0), C.c_int(0)
        self._call("alens_num_ghosts", C.byref(a), C.byref(b), C.byref(c))
        return dict(ghosts=a.value, sent_left=b.value, sent_right=c.value)

    def get_stamps(self, cap=4096):
        buf = np.zeros((cap, 8), dtype=np.uint64)
        n = C.c_int(0)
        self._call("alens_get_stamps", buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), C.c_int(cap), C.byref(n))
        return buf[:min(n.value, cap)]

    def get_pool_stats(self):
        a, b, c = C.c_longlong(0), C.c_longlong(0), C.c_longlong(0)
        self._call("alens_get_pool_stats", C.byref(a), C.byref(b), C.byref(c))
        return dict(collision=a.value, one_side=b.value, bilateral=c.value)

    def get_live_stats(self):
        a, b = C.c_longlong(0), C.c_longlong(0)
        self._call("alens_get_live_stats", C.byref(a), C.byref(b))
        return dict(live_slots=a.value, live_rods=b.value)

    def constraint_digest(self):
        """order-independent digest of the pool: dict(rows, list_hash, gamma_hash, sum_gamma, sum_gamma2, sum_wgamma)"""
        u = (C.c_ulonglong * 3)()
        f = (C.c_double * 3)()
        self._call("alens_constraint_digest", u, f)
        return dict(rows=int(u[0]), list_hash=int(u[1]), gamma_hash=int(u[2]), sum_gamma=f[0], sum_gamma2=f[1], sum_wgamma=f[2])

    def sum_constraint_stress(self, with_one_side=False):
        """(uni, bi): 3x3 sums of gamma * unit stress over the unilateral / bilateral blocks of the last solve
        (ConstraintCollector::sumLocalConstraintStress after writeBackGamma), reduced on the device"""
        u = np.zeros(9)
        b = np.zeros(9)
        self._call("alens_sum_constraint_stress", C.c_int(1 if with_one_side else 0), _dp(u), _dp(b))
        return u.reshape(3, 3), b.reshape(3, 3)

    def comm_mode(self):
        a, b = C.c_int(0), C.c_int(0)
        self._call("alens_comm_mode", C.byref(a), C.byref(b))
        return dict(connected=bool(a.value), fused=bool(b.value))

    def time_kernel(self, which, reps=20):
        us = C.c_double(0)
        self._call("alens_time_kernel", C.c_char_p(which.encode()), C.c_int(reps), C.byref(us))
        return us.value

    def get_positions(self):
        out = np.zeros((self.n_rods, 3))
        self._call("alens_get_positions", _dp(out))
        return out

    def get_rod_state(self):
        pos = np.zeros((self.n_rods, 3))
        q = np.zeros((self.n_rods, 4))
        self._call("alens_get_rod_state", _dp(pos), _dp(q))
        return pos, q

    # ---- constraints
    def collect_pair_collision(self):
        n = C.c_longlong(0)
        self._call("alens_collect_pair_collision", C.byref(n))
        return n.value

    def append_constraints(self, blocks):
        blocks = np.ascontiguousarray(blocks, dtype=BLOCK_DTYPE)
        self._call("alens_append_constraints", C.c_void_p(blocks.ctypes.data), C.c_longlong(len(blocks)))

    def clear_constraints(self):
        self._call("alens_clear_constraints")

    def num_constraints(self):
        n = C.c_longlong(0)
        self._call("alens_num_constraints", C.byref(n))
        return n.value

    def get_constraints(self, with_stress=True, write_back=False):
        n = self.num_constraints()
        out = np.zeros(max(n, 1), dtype=BLOCK_DTYPE)
        self._call("alens_get_constraints", C.c_void_p(out.ctypes.data), C.c_longlong(len(out)),
                   C.c_int(1 if with_stress else 0), C.c_int(1 if write_back else 0))
        return out[:n]

    # ---- mobility / solve
    def calc_mobility(self, viscosity):
        self._call("alens_calc_mobility", C.c_double(viscosity))

    def mobility_apply(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros_like(x)
        self._call("alens_mobility_apply", _dp(x), _dp(y))
        return y

    def setup_constraints(self, vel_nc, dt):
        v = None if vel_nc is None else np.ascontiguousarray(vel_nc, dtype=np.float64)
        self._call("alens_setup_constraints", _dp(v), C.c_double(dt))

    def solve_constraints(self, vel_nc, dt, res, max_ite, solver_choice=0):
        v = None if vel_nc is None else np.ascontiguousarray(vel_nc, dtype=np.float64)
        rep = SolveReport()
        self._call("alens_solve_constraints", _dp(v), C.c_double(dt), C.c_double(res), C.c_int(max_ite),
                   C.c_int(solver_choice), C.byref(rep))
        return rep

    def solve_constraints_raw(self, vel_nc_p, dt, res, max_ite, solver_choice=0):
        rep = SolveReport()
        self._call("alens_solve_constraints", C.c_void_p(vel_nc_p), C.c_double(dt), C.c_double(res),
                   C.c_int(max_ite), C.c_int(solver_choice), C.byref(rep))
        return rep

    def bcqp_solve(self, b, x0, tol, max_ite, solver_choice=0):
        b = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
        x = np.array(x0, dtype=np.float64)
        if len(x) == 0:
            x = np.zeros(1)
        rep = SolveReport()
        self._call("alens_bcqp_solve", _dp(b), _dp(x), C.c_double(tol), C.c_int(max_ite), C.c_int(solver_choice),
                   C.byref(rep))
        return x[:len(x0)], rep

    def operator_apply(self, x, want_force_vel=False):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(max(len(x), 1))
        if want_force_vel:
            f = np.zeros(6 * self.n_rods)
            v = np.zeros(6 * self.n_rods)
            self._call("alens_operator_apply", _dp(x), _dp(y), _dp(f), _dp(v))
            return y[:len(x)], f, v
        self._call("alens_operator_apply", _dp(x), _dp(y), None, None)
        return y[:len(x)]

    def get_history(self):
        n = C.c_int(0)
        self._call("alens_get_history", None, C.c_int(0), C.byref(n))
        rows = np.zeros((max(n.value, 1), 6))
        self._call("alens_get_history", _dp(rows), C.c_int(len(rows)), C.byref(n))
        return rows[:n.value]

    def get_gamma(self):
        n = self.num_constraints()
        g = np.zeros(max(n, 1))
        self._call("alens_get_gamma", _dp(g), C.c_longlong(len(g)))
        return g[:n]

    def get_force_velocity(self):
        out = [np.zeros(6 * self.n_rods) for _ in range(4)]
        self._call("alens_get_force_velocity", *[_dp(o) for o in out])
        return dict(forceU=out[0], velU=out[1], forceB=out[2], velB=out[3])

    def get_force_velocity_raw(self, fu_p, vu_p, fb_p, vb_p):
        self._call("alens_get_force_velocity", C.c_void_p(fu_p), C.c_void_p(vu_p), C.c_void_p(fb_p), C.c_void_p(vb_p))

    def step_euler(self, dt):
        self._call("alens_step_euler", C.c_double(dt))

    # ---- instrumentation
    def get_timers(self):
        t = Timers()
        self._call("alens_get_timers", C.byref(t))
        return t.as_dict()

    def reset_timers(self):
        self._call("alens_reset_timers")

    def get_long_rod_stats(self):
        a, b, r0, r1 = C.c_longlong(0), C.c_longlong(0), C.c_double(0), C.c_double(0)
        self._call("alens_get_long_rod_stats", C.byref(a), C.byref(b), C.byref(r0), C.byref(r1))
        return dict(long_rods=a.value, long_rows=b.value, short_radius=r0.value, max_radius=r1.value)

    def get_collect_stats(self):
        a, b, c = C.c_longlong(0), C.c_longlong(0), C.c_longlong(0)
        self._call("alens_get_collect_stats", C.byref(a), C.byref(b), C.byref(c))
        return dict(cells=a.value, candidates=b.value, hits=c.value)
