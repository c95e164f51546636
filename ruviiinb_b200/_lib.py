"""ctypes binding of libruvnb_b200.so (include/ruvnb.h).

The library is the product: there is no Python/NumPy fallback.  Importing this module never touches
the GPU; `load()` raises if the shared object was not built (`python -c 'import __graft_entry__ as g;
g.build()'`), and `Context()` raises if no CUDA device is usable.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libruvnb_b200.so")
MAX_K = 8

# every symbol include/ruvnb.h declares (tests/test_abi.py checks the header against this list)
SYMBOLS = [
    "ruvnb_abi_version", "ruvnb_opts_default", "ruvnb_create", "ruvnb_destroy", "ruvnb_last_error",
    "ruvnb_kernel_launches", "ruvnb_kernel_time", "ruvnb_kernel_time_reset", "ruvnb_stream", "ruvnb_last_exact_rows",
    "ruvnb_last_exact_cells", "ruvnb_selftest_math", "ruvnb_measure_fp64", "ruvnb_cert_diag",
    "ruvnb_comm_unique_id", "ruvnb_comm_init", "ruvnb_set_design",
    "ruvnb_upload_counts_dense", "ruvnb_upload_counts_csr", "ruvnb_winsorize", "ruvnb_get_counts", "ruvnb_set_state",
    "ruvnb_get_state", "ruvnb_init_W", "ruvnb_init_coef", "ruvnb_loglik", "ruvnb_irls_iteration", "ruvnb_zinb_pi0", "ruvnb_snapshot_best",
    "ruvnb_restore_best", "ruvnb_halve_steps", "ruvnb_accept", "ruvnb_disp_apl_grid", "ruvnb_estimate_disp", "ruvnb_estimate_disp_ex",
    "ruvnb_disp_newton", "ruvnb_fit_nb", "ruvnb_get_res", "ruvnb_fast_W_all", "ruvnb_fast_Mb_all", "ruvnb_fast_cap_psi",
    "ruvnb_fast_init", "ruvnb_fast_begin_outer", "ruvnb_fast_iteration", "ruvnb_fast_halve_steps", "ruvnb_fast_accept",
    "ruvnb_set_cell_shard", "ruvnb_set_row_mask", "ruvnb_upload_counts_rows", "ruvnb_upload_counts_done", "ruvnb_get_cells",
    "ruvnb_upload_counts_csc", "ruvnb_get_res_csc", "ruvnb_set_batches", "ruvnb_nbatch", "ruvnb_set_psi_callback",
    "ruvnb_host_alloc", "ruvnb_host_free", "ruvnb_get_res_ex", "ruvnb_fast_res", "ruvnb_selftest_squeeze_var", "ruvnb_estimate_disp_r", "ruvnb_set_res_zinb",
]


class Opts(C.Structure):
    """`ruvnb_opts` -- the arguments/constants of ruvIII.nb (R/ruvIIInb.R:27-28,155,524,533,597)."""
    _fields_ = [("k", C.c_int), ("robust", C.c_int), ("lambda_a", C.c_double), ("lambda_b", C.c_double),
                ("step_fac", C.c_double), ("inner_maxit", C.c_int), ("outer_maxit", C.c_int),
                ("inner_tol", C.c_double), ("outer_tol", C.c_double), ("psi_tol", C.c_double),
                ("zinb", C.c_int), ("huber_fastpath", C.c_int), ("verbose", C.c_int)]


class TraceRec(C.Structure):
    _fields_ = [("outer", C.c_int), ("iter", C.c_int), ("halving", C.c_int), ("degener", C.c_int),
                ("accepted", C.c_int), ("logl_old", C.c_double), ("logl_new", C.c_double), ("logl", C.c_double),
                ("bad_alpha", C.c_int), ("bad_W", C.c_int), ("bad_Mb", C.c_int)]


class FitStats(C.Structure):
    _fields_ = [("n_outer", C.c_int), ("n_inner", C.c_int), ("n_trace", C.c_int), ("inner_seconds", C.c_double),
                ("disp_seconds", C.c_double), ("total_seconds", C.c_double), ("kernel_launches", C.c_longlong),
                ("n_exact_mad_rows", C.c_int), ("n_exact_mad_cells", C.c_int)]


class ResStats(C.Structure):
    _fields_ = [("checksum", C.c_double), ("device_seconds", C.c_double), ("caps_ms", C.c_double), ("elem_ms", C.c_double),
                ("select_ms", C.c_double), ("mb_ms", C.c_double), ("elements", C.c_longlong), ("launches", C.c_longlong),
                ("clipped_lo", C.c_longlong), ("clipped_hi", C.c_longlong)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if not n.startswith("clipped")}


OUT_F64, OUT_I32, OUT_LOG1P, OUT_NONE = 0, 1, 2, 3
RES_KINDS = {"pearson": 1, "logcounts": 2, "quantile": 3}

_lib = None


def load():
    """dlopen the in-tree library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with __graft_entry__.build(); "
                           "ruviiinb_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, cp, ip, dp = C.c_void_p, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_double)
    lib.ruvnb_abi_version.restype = C.c_int
    lib.ruvnb_opts_default.argtypes = [C.POINTER(Opts)]
    lib.ruvnb_opts_default.restype = None
    lib.ruvnb_create.argtypes = [C.POINTER(vp), C.c_int]
    lib.ruvnb_destroy.argtypes = [vp]
    lib.ruvnb_destroy.restype = None
    lib.ruvnb_last_error.argtypes = [vp]
    lib.ruvnb_last_error.restype = cp
    lib.ruvnb_kernel_launches.argtypes = [vp]
    lib.ruvnb_kernel_launches.restype = C.c_longlong
    lib.ruvnb_kernel_time.argtypes = [vp, C.c_int, dp, C.POINTER(C.c_longlong)]
    lib.ruvnb_kernel_time_reset.argtypes = [vp]
    lib.ruvnb_kernel_time_reset.restype = None
    lib.ruvnb_stream.argtypes = [vp]
    lib.ruvnb_stream.restype = vp
    lib.ruvnb_last_exact_rows.argtypes = [vp]
    lib.ruvnb_last_exact_cells.argtypes = [vp]
    lib.ruvnb_measure_fp64.argtypes = [vp, dp]
    lib.ruvnb_cert_diag.argtypes = [vp, ip]
    lib.ruvnb_selftest_math.argtypes = [C.c_int, vp, vp, C.c_int64]
    lib.ruvnb_selftest_squeeze_var.argtypes = [vp, C.c_int64, C.c_double, vp, C.c_int, vp]
    lib.ruvnb_comm_unique_id.argtypes = [vp]
    lib.ruvnb_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.ruvnb_set_design.argtypes = [vp, C.c_int64, C.c_int64, C.c_int, vp, vp, vp, C.c_int64, C.c_int64]
    lib.ruvnb_upload_counts_dense.argtypes = [vp, vp, C.c_int, vp]
    lib.ruvnb_upload_counts_csr.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.ruvnb_winsorize.argtypes = [vp, vp, vp]
    lib.ruvnb_get_counts.argtypes = [vp, vp]
    lib.ruvnb_set_state.argtypes = [vp, cp, vp, C.c_int]
    lib.ruvnb_get_state.argtypes = [vp, cp, vp]
    lib.ruvnb_init_W.argtypes = [vp, C.c_int, C.c_int, C.c_double, vp, vp, ip]
    lib.ruvnb_init_coef.argtypes = [vp, C.POINTER(Opts)]
    lib.ruvnb_loglik.argtypes = [vp, C.POINTER(Opts), dp]
    lib.ruvnb_irls_iteration.argtypes = [vp, C.POINTER(Opts), dp, ip]
    lib.ruvnb_zinb_pi0.argtypes = [vp, C.POINTER(Opts), ip]
    lib.ruvnb_snapshot_best.argtypes = [vp]
    lib.ruvnb_restore_best.argtypes = [vp]
    lib.ruvnb_halve_steps.argtypes = [vp, C.POINTER(Opts)]
    lib.ruvnb_accept.argtypes = [vp]
    lib.ruvnb_disp_apl_grid.argtypes = [vp, C.POINTER(Opts), vp, vp]
    lib.ruvnb_estimate_disp.argtypes = [vp, C.POINTER(Opts), dp]
    lib.ruvnb_estimate_disp_ex.argtypes = [vp, C.POINTER(Opts), C.c_double, dp, vp, dp]
    lib.ruvnb_estimate_disp_r.argtypes = [vp, C.POINTER(Opts), C.c_int, C.c_double, dp, vp, vp]
    lib.ruvnb_disp_newton.argtypes = [vp, C.POINTER(Opts), vp, C.c_int]
    lib.ruvnb_fit_nb.argtypes = [vp, C.POINTER(Opts), vp, vp, C.c_int, vp, C.c_int, vp, C.POINTER(FitStats)]
    lib.ruvnb_get_res.argtypes = [vp, C.c_int, C.c_int64, vp, vp]
    lib.ruvnb_fast_W_all.argtypes = [vp, C.c_int, C.c_double, vp, vp, vp, vp, C.c_int, C.c_double, vp, ip, dp]
    lib.ruvnb_fast_Mb_all.argtypes = [vp, C.c_double, vp, C.c_int64, vp]
    lib.ruvnb_fast_cap_psi.argtypes = [vp, C.c_int64]
    lib.ruvnb_fast_init.argtypes = [vp, C.POINTER(Opts), vp]
    lib.ruvnb_fast_begin_outer.argtypes = [vp, C.POINTER(Opts), dp]
    lib.ruvnb_fast_iteration.argtypes = [vp, C.POINTER(Opts), dp, ip]
    lib.ruvnb_fast_halve_steps.argtypes = [vp, C.POINTER(Opts)]
    lib.ruvnb_fast_accept.argtypes = [vp]
    lib.ruvnb_set_res_zinb.argtypes = [vp, vp, C.c_double]
    lib.ruvnb_set_cell_shard.argtypes = [vp, C.c_int64, C.c_int64]
    lib.ruvnb_set_row_mask.argtypes = [vp, vp]
    lib.ruvnb_upload_counts_rows.argtypes = [vp, vp, C.c_int, C.c_int64, C.c_int64]
    lib.ruvnb_upload_counts_done.argtypes = [vp]
    lib.ruvnb_get_cells.argtypes = [vp, vp, C.c_int64, vp]
    lib.ruvnb_host_alloc.argtypes = [C.c_size_t]
    lib.ruvnb_host_alloc.restype = vp
    lib.ruvnb_host_free.argtypes = [vp]
    lib.ruvnb_host_free.restype = None
    lib.ruvnb_get_res_ex.argtypes = [vp, C.c_int, C.c_int64, vp, C.c_int, vp, C.c_int64, C.c_int64, C.POINTER(ResStats)]
    lib.ruvnb_set_batches.argtypes = [vp, vp, C.c_int64, C.c_int]
    lib.ruvnb_nbatch.argtypes = [vp]
    lib.ruvnb_set_psi_callback.argtypes = [vp, PSI_FN, vp]
    lib.ruvnb_upload_counts_csc.argtypes = [vp, vp, vp, vp, C.c_int64, C.c_int64]
    lib.ruvnb_get_res_csc.argtypes = [vp, C.c_int, C.c_int64, vp, C.c_int, C.c_int64, C.c_int64, vp, vp, vp, C.c_int64, C.POINTER(C.c_int64),
                                      C.POINTER(ResStats)]
    lib.ruvnb_fast_res.argtypes = [vp, C.c_int, C.c_double, vp, C.c_int64, C.c_int, vp, vp, C.c_int64, C.c_int64, C.POINTER(ResStats)]
    for s in SYMBOLS:
        f = getattr(lib, s)
        if f.restype is C.c_int and s not in ("ruvnb_abi_version",):
            pass
    _lib = lib
    return lib


PSI_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p)   # ruvnb_psi_fn(user, ctx)


class RuvnbError(RuntimeError):
    pass


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """One GPU, one problem (`ruvnb_ctx`).  Arrays cross the boundary as host NumPy arrays shaped like
    the R objects (column-major doubles; counts int32)."""

    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.ruvnb_create(C.byref(h), int(device))
        if rc != 0:
            raise RuvnbError(f"ruvnb_create(device={device}) failed with {rc}: no usable CUDA device "
                             "(ruviiinb_b200 has no CPU fallback)")
        self.h = h
        self.G = self.N = self.m = self.k = 0
        self.g0 = self.g1 = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.ruvnb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise RuvnbError(f"libruvnb_b200 error {rc}: {self.lib.ruvnb_last_error(self.h).decode()}")

    # ---- set-up -----------------------------------------------------------------------------
    def comm_init(self, uid, rank, nranks):
        buf = (C.c_ubyte * 128).from_buffer_copy(bytes(uid))
        self._ck(self.lib.ruvnb_comm_init(self.h, C.cast(buf, C.c_void_p), rank, nranks))

    def set_batches(self, batch, nbatch=None):
        """Batch label of every cell for a batch-specific dispersion; before set_design.  Without `nbatch` any integer labels are
        renumbered 0.. in sorted order (the reference's 1..max(batch)); with it the labels are taken as they are, 0-based, and some
        batches may be empty (a cell shard).  Returns the 0-based labels."""
        if nbatch is None:
            lab, b0 = np.unique(np.asarray(batch), return_inverse=True)
            nbatch = int(lab.size)
        else:
            b0 = np.asarray(batch)
        b0 = np.ascontiguousarray(b0, dtype=np.int32)
        self._ck(self.lib.ruvnb_set_batches(self.h, _ptr(b0), b0.size, int(nbatch)))
        self.nbatch = int(nbatch)
        return b0

    def set_psi_callback(self, fn):
        """fn(ctx) is called at every dispersion refresh of fit_nb (it reads the state and calls set_state("psi", ...)); None removes it."""
        if fn is None:
            self._psi_cb = PSI_FN(0)
        else:
            def tramp(_user, _ctx):
                try:
                    fn(self)
                    return 0
                except Exception as e:      # an exception must not cross the C frame
                    self._psi_cb_error = e
                    return -1
            self._psi_cb = PSI_FN(tramp)
        self._ck(self.lib.ruvnb_set_psi_callback(self.h, self._psi_cb, None))

    def set_design(self, G, N, grp, ctl, zeroinf=None, rows=None, m=None):
        grp = np.ascontiguousarray(grp, dtype=np.int32)
        ctl = np.ascontiguousarray(np.asarray(ctl).astype(np.uint8))
        zi = None if zeroinf is None else np.ascontiguousarray(np.asarray(zeroinf).astype(np.uint8))
        if m is None:   # (a cell shard passes the group count of the whole problem)
            m = int(grp.max()) + 1 if grp.size else 0
        g0, g1 = (0, G) if rows is None else rows
        self._ck(self.lib.ruvnb_set_design(self.h, G, N, m, _ptr(grp), _ptr(ctl), _ptr(zi), g0, g1))
        self.G, self.N, self.m, self.g0, self.g1 = G, N, m, g0, g1
        self.Gl = g1 - g0
        self.n_ctl = int(np.asarray(ctl).astype(bool).sum())

    def upload_counts(self, Y, Yctl=None, layout="gene_major"):
        """Y: local rows, int32.  layout 'gene_major' = C-order (Gl, N); 'cell_major' = R's column-major."""
        lay = 0 if layout == "gene_major" else 1
        Y = np.ascontiguousarray(Y, dtype=np.int32) if lay == 0 else np.asfortranarray(Y, dtype=np.int32)
        if Yctl is not None:
            Yctl = np.ascontiguousarray(Yctl, dtype=np.int32) if lay == 0 else np.asfortranarray(Yctl, dtype=np.int32)
        self._ck(self.lib.ruvnb_upload_counts_dense(self.h, _ptr(Y), lay, _ptr(Yctl)))

    def upload_counts_csr(self, indptr, indices, values, ctl_csr=None):
        ip = np.ascontiguousarray(indptr, dtype=np.int64)
        ix = np.ascontiguousarray(indices, dtype=np.int32)
        vl = np.ascontiguousarray(values, dtype=np.int32)
        c = [None, None, None]
        if ctl_csr is not None:
            c = [np.ascontiguousarray(ctl_csr[0], dtype=np.int64), np.ascontiguousarray(ctl_csr[1], dtype=np.int32),
                 np.ascontiguousarray(ctl_csr[2], dtype=np.int32)]
        self._ck(self.lib.ruvnb_upload_counts_csr(self.h, _ptr(ip), _ptr(ix), _ptr(vl), _ptr(c[0]), _ptr(c[1]), _ptr(c[2])))

    def winsorize(self):
        cap = np.empty(self.Gl, dtype=np.float64)
        nz = np.empty(self.Gl, dtype=np.int64)
        self._ck(self.lib.ruvnb_winsorize(self.h, _ptr(cap), _ptr(nz)))
        return cap, nz

    def get_counts(self):
        out = np.empty((self.Gl, self.N), dtype=np.int32)
        self._ck(self.lib.ruvnb_get_counts(self.h, _ptr(out)))
        return out

    # ---- state ------------------------------------------------------------------------------
    def _shape(self, name):
        nb = getattr(self, "nbatch", 1)
        if name in ("psi", "psi_w") and nb > 1:
            return (self.Gl, nb)
        return {"W": (self.N, self.k), "alpha": (self.Gl, self.k), "alpha_dev": (self.Gl, self.k),
                "Mb": (self.Gl, self.m), "cell_loglik": (self.N,), "cell_step": (self.N,),
                "wt_ctl": (self.n_ctl,)}.get(name, (self.Gl,))

    def set_state(self, name, a, k=0):
        if name == "W":
            self.k = int(np.asarray(a).shape[1]) if np.asarray(a).ndim == 2 else 1
            k = self.k
        a = np.asfortranarray(np.asarray(a, dtype=np.float64).reshape(self._shape(name), order="F"))
        self._ck(self.lib.ruvnb_set_state(self.h, name.encode(), _ptr(a), k))

    def get_state(self, name):
        out = np.empty(self._shape(name), dtype=np.float64, order="F")
        self._ck(self.lib.ruvnb_get_state(self.h, name.encode(), _ptr(out)))
        return out

    # ---- operators ----------------------------------------------------------------------------
    def init_W(self, k, max_iter=0, tol=0.0):
        """H1 (R/ruvIIInb.R:160-168) on the device; returns (W0 N x k, singular values, iterations) and sets "W"."""
        self.k = int(k)
        W = np.empty((self.N, self.k), dtype=np.float64, order="F")
        sv = np.empty(self.k, dtype=np.float64)
        it = C.c_int()
        self._ck(self.lib.ruvnb_init_W(self.h, self.k, int(max_iter), float(tol), _ptr(W), _ptr(sv), C.byref(it)))
        return W, sv, it.value

    def init_coef(self, opts):
        self._ck(self.lib.ruvnb_init_coef(self.h, C.byref(opts)))

    def loglik(self, opts):
        t = C.c_double()
        self._ck(self.lib.ruvnb_loglik(self.h, C.byref(opts), C.byref(t)))
        return t.value

    def irls_iteration(self, opts):
        t = C.c_double()
        bad = (C.c_int * 3)()
        self._ck(self.lib.ruvnb_irls_iteration(self.h, C.byref(opts), C.byref(t), bad))
        return t.value, list(bad)

    def zinb_pi0(self, opts):
        bad = C.c_int()
        self._ck(self.lib.ruvnb_zinb_pi0(self.h, C.byref(opts), C.byref(bad)))
        return bad.value

    def snapshot_best(self):
        self._ck(self.lib.ruvnb_snapshot_best(self.h))

    def restore_best(self):
        self._ck(self.lib.ruvnb_restore_best(self.h))

    def halve_steps(self, opts):
        self._ck(self.lib.ruvnb_halve_steps(self.h, C.byref(opts)))

    def accept(self):
        self._ck(self.lib.ruvnb_accept(self.h))

    def disp_apl_grid(self, opts):
        l0 = np.empty((self.Gl, 21), dtype=np.float64, order="F")
        rs = np.empty(self.Gl, dtype=np.float64)
        self._ck(self.lib.ruvnb_disp_apl_grid(self.h, C.byref(opts), _ptr(l0), _ptr(rs)))
        return l0, rs

    def estimate_disp(self, opts):
        c = C.c_double()
        self._ck(self.lib.ruvnb_estimate_disp(self.h, C.byref(opts), C.byref(c)))
        return c.value

    def estimate_disp_ex(self, opts, prior_df=None):
        """estimateDisp.par on the current parameters -> dict(common, trended, prior_df); sets "psi" (tagwise)."""
        c, pdf = C.c_double(), C.c_double()
        tr = np.empty(self.Gl, dtype=np.float64)
        self._ck(self.lib.ruvnb_estimate_disp_ex(self.h, C.byref(opts), -1.0 if prior_df is None else float(prior_df), C.byref(c), _ptr(tr),
                                                 C.byref(pdf)))
        return dict(common=c.value, trended=tr, prior_df=pdf.value)

    def estimate_disp_r(self, opts, robust=True, prior_df=None):
        """estimateDisp.par(..., robust = robust) on the current parameters -> dict(common, trended, prior_df (per local gene)); sets "psi"."""
        c = C.c_double()
        tr = np.empty(self.Gl, dtype=np.float64)
        pv = np.empty(self.Gl, dtype=np.float64)
        self._ck(self.lib.ruvnb_estimate_disp_r(self.h, C.byref(opts), int(bool(robust)), -1.0 if prior_df is None else float(prior_df),
                                                C.byref(c), _ptr(tr), _ptr(pv)))
        return dict(common=c.value, trended=tr, prior_df=pv)

    def disp_newton(self, opts, maxit=30):
        out = np.empty(self.Gl, dtype=np.float64)
        self._ck(self.lib.ruvnb_disp_newton(self.h, C.byref(opts), _ptr(out), maxit))
        return out

    def fit_nb(self, opts, W0=None, psi_inject=None, trace_cap=4096):
        """`ruvnb_fit_nb`.  psi_inject: (n_psi, Gl) array or None; with batches (n_psi, Gl, nbatch)."""
        if W0 is not None:
            self.k = int(np.asarray(W0).shape[1])
            W0 = np.asfortranarray(W0, dtype=np.float64)
        n_psi = 0
        if psi_inject is not None:
            nb = getattr(self, "nbatch", 1)
            if nb > 1:   # every schedule entry as R's G x nbatch matrix, column-major
                p3 = np.asarray(psi_inject, dtype=np.float64).reshape(-1, self.Gl, nb)
                psi_inject = np.ascontiguousarray(np.transpose(p3, (0, 2, 1)))
            else:
                psi_inject = np.ascontiguousarray(np.atleast_2d(psi_inject), dtype=np.float64)
            n_psi = psi_inject.shape[0]
        trace = (TraceRec * trace_cap)()
        logl = np.full(max(1, opts.outer_maxit), np.nan)
        st = FitStats()
        self._ck(self.lib.ruvnb_fit_nb(self.h, C.byref(opts), _ptr(W0), _ptr(psi_inject), n_psi,
                                       C.cast(trace, C.c_void_p), trace_cap, _ptr(logl), C.byref(st)))
        recs = [{f[0]: getattr(trace[i], f[0]) for f in TraceRec._fields_} for i in range(st.n_trace)]
        stats = {f[0]: getattr(st, f[0]) for f in FitStats._fields_}
        return logl[: st.n_outer].copy(), recs, stats

    def get_res(self, kind, block_size=5000, Mb_cells=None):
        out = np.empty((self.Gl, self.N), dtype=np.float64, order="F")
        mb = None if Mb_cells is None else np.asfortranarray(Mb_cells, dtype=np.float64)
        self._ck(self.lib.ruvnb_get_res(self.h, int(kind), int(block_size), _ptr(mb), _ptr(out)))
        return out

    def set_res_zinb(self, zflag_of_cell, mean_zeroinf):
        z = None if zflag_of_cell is None else np.ascontiguousarray(np.asarray(zflag_of_cell).astype(np.uint8))
        self._ck(self.lib.ruvnb_set_res_zinb(self.h, _ptr(z), float(mean_zeroinf)))

    # ---- cell shards, row mask, slab upload, sinks (include/ruvnb.h) -----------------------------------------------
    def set_cell_shard(self, N_total, c0):
        self._ck(self.lib.ruvnb_set_cell_shard(self.h, int(N_total), int(c0)))

    def set_row_mask(self, alive):
        """Rows reported by get_res_ex / fast_res; None restores all rows."""
        a = None if alive is None else np.ascontiguousarray(np.asarray(alive).astype(np.uint8))
        self._ck(self.lib.ruvnb_set_row_mask(self.h, _ptr(a)))
        self.Gout = self.Gl if a is None else int(a.sum())

    def upload_counts_rows(self, Y_rows, row0, device_ptr=None, nrows=None):
        """One gene-major slab of rows [row0, row0 + nrows): a host int32 array, or a device pointer (`device_ptr`, `nrows`)."""
        if device_ptr is not None:
            self._ck(self.lib.ruvnb_upload_counts_rows(self.h, C.c_void_p(int(device_ptr)), 1, int(row0), int(nrows)))
        else:
            Y = np.ascontiguousarray(Y_rows, dtype=np.int32)
            self._ck(self.lib.ruvnb_upload_counts_rows(self.h, _ptr(Y), 0, int(row0), Y.shape[0]))

    def upload_counts_csc(self, p, i, x, block_cols=None):
        """The slots of a dgCMatrix / scipy.sparse.csc_matrix (G x N, column = cell): p (N + 1), i (0-based rows), x (doubles).
        Uploaded in blocks of `block_cols` columns (default: everything at once) and transposed on the device."""
        p = np.ascontiguousarray(p, dtype=np.int32)
        i = np.ascontiguousarray(i, dtype=np.int32)
        x = np.ascontiguousarray(x, dtype=np.float64)
        N = p.size - 1
        if N != self.N:
            raise ValueError("the column pointer must have one entry per cell plus one")
        step = N if not block_cols else int(block_cols)
        for c0 in range(0, max(N, 1), max(step, 1)):
            nc = min(step, N - c0)
            self._ck(self.lib.ruvnb_upload_counts_csc(self.h, _ptr(p[c0:c0 + nc + 1]), _ptr(i), _ptr(x), int(c0), int(nc)))
        self.upload_counts_done()

    def upload_counts_done(self):
        self._ck(self.lib.ruvnb_upload_counts_done(self.h))

    def get_cells(self, cells):
        """Resident (winsorized) counts of the listed local cells: (Gl, n) int32."""
        cells = np.ascontiguousarray(cells, dtype=np.int64)
        out = np.empty((self.Gl, cells.size), dtype=np.int32, order="F")
        self._ck(self.lib.ruvnb_get_cells(self.h, _ptr(cells), cells.size, _ptr(out)))
        return out

    def _res_out(self, out_kind, ncells, out):
        rows = getattr(self, "Gout", None) or self.Gl
        if out_kind == OUT_NONE:
            return None, rows
        dt = np.int32 if out_kind == OUT_I32 else np.float64
        if out is None:
            out = np.empty((rows, ncells), dtype=dt, order="F")
        assert out.dtype == dt and out.shape == (rows, ncells) and out.flags.f_contiguous
        return out, rows

    def get_res_ex(self, kind, block_size=5000, Mb_cells=None, out_kind=OUT_F64, cells=None, out=None):
        """`ruvnb_get_res_ex`: returns (out or None, stats dict).  cells = (begin, end) restricts the pass to whole blocks."""
        cb, ce = (0, self.N) if cells is None else cells
        out, _ = self._res_out(out_kind, ce - cb, out)
        mb = None if Mb_cells is None else np.asfortranarray(Mb_cells, dtype=np.float64)
        st = ResStats()
        self._ck(self.lib.ruvnb_get_res_ex(self.h, int(kind), int(block_size), _ptr(mb), int(out_kind), _ptr(out), int(cb), int(ce), C.byref(st)))
        return out, st.as_dict()

    def get_res_csc(self, kind, block_size=5000, Mb_cells=None, log1p=False, cells=None, max_dense_bytes=1 << 31):
        """`ruvnb_get_res_csc`: (colptr int64, rowidx int32, values f64, stats) -- the slots of the dgCMatrix makeSCE stores.  The pass
        runs in ranges of whole blocks whose worst case (dense) fits `max_dense_bytes` of host memory, so the sink can never overflow."""
        cb, ce = (0, self.N) if cells is None else cells
        rows = getattr(self, "Gout", None) or self.Gl
        bs = min(int(block_size), max(self.N, 1))
        blocks_per_call = max(1, int(max_dense_bytes // (12 * rows * bs)))
        mb = None if Mb_cells is None else np.asfortranarray(Mb_cells, dtype=np.float64)
        ptrs, idx, val, stats = [np.zeros(1, dtype=np.int64)], [], [], None
        for b0 in range(cb, ce, blocks_per_call * bs):
            b1 = min(ce, b0 + blocks_per_call * bs)
            cap = rows * (b1 - b0)
            cp = np.empty(b1 - b0 + 1, dtype=np.int64); ri = np.empty(cap, dtype=np.int32); vv = np.empty(cap, dtype=np.float64)
            nnz = C.c_int64(0)
            st = ResStats()
            self._ck(self.lib.ruvnb_get_res_csc(self.h, int(kind), bs, _ptr(mb), int(bool(log1p)), int(b0), int(b1), _ptr(cp), _ptr(ri), _ptr(vv),
                                                int(cap), C.byref(nnz), C.byref(st)))
            n = int(nnz.value)
            ptrs.append(cp[1:] + ptrs[-1][-1]); idx.append(ri[:n].copy()); val.append(vv[:n].copy())
            d = st.as_dict()
            stats = d if stats is None else {k: (stats[k] + v if isinstance(v, (int, float)) else v) for k, v in d.items()}
        return np.concatenate(ptrs), (np.concatenate(idx) if idx else np.empty(0, np.int32)), (np.concatenate(val) if val else np.empty(0)), stats

    def fast_res(self, kind, annot_grp, lambda_b=16.0, block_size=5000, out_kind=OUT_F64, cells=None, out=None, want_Mb_all=False):
        """`ruvnb_fast_res`: Mb.all block -> caps -> residual of the same block on the device.  Returns (out, Mb_all or None, stats)."""
        cb, ce = (0, self.N) if cells is None else cells
        out, _ = self._res_out(out_kind, ce - cb, out)
        ag = np.ascontiguousarray(annot_grp, dtype=np.int32)
        mball = np.empty((self.Gl, ce - cb), dtype=np.float64, order="F") if want_Mb_all else None
        st = ResStats()
        self._ck(self.lib.ruvnb_fast_res(self.h, int(kind), float(lambda_b), _ptr(ag), int(block_size), int(out_kind), _ptr(out), _ptr(mball),
                                         int(cb), int(ce), C.byref(st)))
        return out, mball, st.as_dict()

    # ---- fastruvIII.nb subsample fit (F3/F4) ------------------------------------------------------
    def fast_init(self, opts, U0):
        U0 = np.asfortranarray(U0, dtype=np.float64)
        self.k = U0.shape[1]
        self._ck(self.lib.ruvnb_fast_init(self.h, C.byref(opts), _ptr(U0)))

    def fast_begin_outer(self, opts):
        t = C.c_double()
        self._ck(self.lib.ruvnb_fast_begin_outer(self.h, C.byref(opts), C.byref(t)))
        return t.value

    def fast_iteration(self, opts):
        t = C.c_double()
        bad = (C.c_int * 3)()
        self._ck(self.lib.ruvnb_fast_iteration(self.h, C.byref(opts), C.byref(t), bad))
        return t.value, list(bad)

    def fast_halve_steps(self, opts):
        self._ck(self.lib.ruvnb_fast_halve_steps(self.h, C.byref(opts)))

    def fast_accept(self):
        self._ck(self.lib.ruvnb_fast_accept(self.h))

    def fast_W_all(self, W_init, wt_ctl, W_med, W_mad, lambda_a=0.01, max_sweeps=8, tol=1e-5):
        """F5 (R/fastruvIIInb.R:632-716): returns (W, sweeps, crit)."""
        W_init = np.asfortranarray(W_init, dtype=np.float64)
        self.k = W_init.shape[1]
        out = np.empty_like(W_init, order="F")
        ns, cr = C.c_int(), C.c_double()
        wt = np.ascontiguousarray(wt_ctl, dtype=np.float64)
        med = np.ascontiguousarray(W_med, dtype=np.float64)
        mad = np.ascontiguousarray(W_mad, dtype=np.float64)
        self._ck(self.lib.ruvnb_fast_W_all(self.h, self.k, lambda_a, _ptr(wt), _ptr(W_init), _ptr(med), _ptr(mad), max_sweeps, tol,
                                           _ptr(out), C.byref(ns), C.byref(cr)))
        return out, ns.value, cr.value

    def fast_Mb_all(self, annot_grp, lambda_b=16.0, block_size=5000):
        """F6 (R/fastruvIIInb.R:719-748): G x N per-cell Mb."""
        ag = np.ascontiguousarray(annot_grp, dtype=np.int32)
        out = np.empty((self.Gl, self.N), dtype=np.float64, order="F")
        self._ck(self.lib.ruvnb_fast_Mb_all(self.h, lambda_b, _ptr(ag), int(block_size), _ptr(out)))
        return out

    KERNELS = ("devstats", "pass1", "w_update", "pass2", "loglik", "exact_mad", "gram", "seg_alpha", "seg_w", "seg_mb", "seg_tail")

    def kernel_times(self):
        """{name: (total_ms, launches)} of the sweep kernels since the last reset (CUDA events, lib stream)."""
        out = {}
        for i, n in enumerate(self.KERNELS):
            ms, cnt = C.c_double(), C.c_longlong()
            self._ck(self.lib.ruvnb_kernel_time(self.h, i, C.byref(ms), C.byref(cnt)))
            out[n] = (ms.value, cnt.value)
        return out

    def kernel_time_reset(self):
        self.lib.ruvnb_kernel_time_reset(self.h)

    @property
    def last_exact_rows(self):
        return int(self.lib.ruvnb_last_exact_rows(self.h))

    def cert_diag(self):
        out = (C.c_int * 8)()
        self._ck(self.lib.ruvnb_cert_diag(self.h, out))
        return list(out)

    def measure_fp64(self):
        t = C.c_double()
        self._ck(self.lib.ruvnb_measure_fp64(self.h, C.byref(t)))
        return t.value

    @property
    def last_exact_cells(self):
        return int(self.lib.ruvnb_last_exact_cells(self.h))

    @property
    def stream(self):
        return self.lib.ruvnb_stream(self.h)

    @property
    def kernel_launches(self):
        return int(self.lib.ruvnb_kernel_launches(self.h))


def fast_cap_psi(psi):
    """F7 (R/fastruvIIInb.R:765-770): psi capped at exp(median(log psi) + 3 mad(log psi))."""
    p = np.ascontiguousarray(psi, dtype=np.float64).copy()
    rc = load().ruvnb_fast_cap_psi(_ptr(p), p.size)
    if rc != 0:
        raise RuvnbError(f"ruvnb_fast_cap_psi failed with {rc}")
    return p


def squeeze_var_df_prior(s2, df1, covariate=None, robust=False):
    """csrc/limma_host.h on the host: df.prior of limma::squeezeVar as estimateDisp.par calls it (one value per gene)."""
    s2 = np.ascontiguousarray(s2, dtype=np.float64)
    cov = None if covariate is None else np.ascontiguousarray(covariate, dtype=np.float64)
    out = np.empty_like(s2)
    rc = load().ruvnb_selftest_squeeze_var(_ptr(s2), s2.size, float(df1), _ptr(cov), int(bool(robust)), _ptr(out))
    if rc != 0:
        raise RuvnbError(f"ruvnb_selftest_squeeze_var failed with {rc}")
    return out


def selftest_math(which, x):
    """Host evaluation of the kernels' table-driven exp ('exp') / log ('log') / unchecked log on the interleaved
    table ('log_nc2', normal positive arguments) -- csrc/fastmath.cuh."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    rc = load().ruvnb_selftest_math({"exp": 0, "log": 1, "log_nc2": 2}[which], _ptr(x), _ptr(out), x.size)
    if rc != 0:
        raise RuvnbError(f"ruvnb_selftest_math failed with {rc}")
    return out


def default_opts(**kw):
    o = Opts()
    load().ruvnb_opts_default(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o
