"""Ingest / egress formats around the path: counts as the slots of a dgCMatrix (compressed sparse COLUMN,
what ruvIII.nb returns and makeSCE / get.res read, R/ruvIIInb.R:611-612, R/get.res.R:14) transposed on the device, results as a
sparse matrix compacted on the device (`as(..., "sparseMatrix")`, R/makeSCE.R:24-38) and as a file-backed block sink
(R/get.res.R:21-23,125)."""
import numpy as np
import pytest
import scipy.sparse as sp


def _fit_inputs(G=70, N=640, k=2, m=3, n_ctl=20, seed=5):
    from ruviiinb_b200 import synth
    Y, M, ctl, truth = synth.make_counts(G, N, k, m, n_ctl, seed=seed, gmean_mu=-0.6, gmean_sd=1.0)   # mostly zeros
    return Y, M, ctl, truth


def _state_out(Y, M, ctl, truth, k):
    rng = np.random.default_rng(3)
    G, N = Y.shape
    W = np.linalg.qr(rng.standard_normal((N, k)))[0]
    return dict(counts=Y, W=W, M=M, ctl=ctl, a=0.3 * rng.standard_normal((G, k)), gmean=truth["gmean"], Mb=0.2 * rng.standard_normal((G, M.shape[1])),
                psi=truth["psi"], pi0=None, batch=np.ones(N, dtype=np.int64))


@pytest.mark.gpu
def test_csc_upload_equals_dense_upload():
    from ruviiinb_b200 import _lib
    Y, M, ctl, _ = _fit_inputs()
    G, N = Y.shape
    grp = np.argmax(M, axis=1).astype(np.int32)
    Yc = sp.csc_matrix(Y.astype(np.float64))
    with _lib.Context(0) as ctx:
        ctx.set_design(G, N, grp, ctl)
        ctx.upload_counts_csc(Yc.indptr, Yc.indices, Yc.data, block_cols=97)     # ragged column blocks, windows of the caller's slots
        got = ctx.get_counts()
        assert np.array_equal(got, Y)
        capd, nzd = ctx.winsorize()
    with _lib.Context(0) as ctx:
        ctx.set_design(G, N, grp, ctl)
        ctx.upload_counts(Y)
        cap, nz = ctx.winsorize()
    assert np.array_equal(cap, capd) and np.array_equal(nz, nzd)
    # non-integer doubles are rounded up like ceiling() (R/ruvIIInb.R:53); a row outside the matrix is an argument error
    Yf = Yc.copy(); Yf.data = Yf.data - 0.25
    with _lib.Context(0) as ctx:
        ctx.set_design(G, N, grp, ctl)
        ctx.upload_counts_csc(Yf.indptr, Yf.indices, Yf.data)
        assert np.array_equal(ctx.get_counts(), Y)
        bad = Yc.indices.copy(); bad[0] = G
        with pytest.raises(_lib.RuvnbError, match="row outside"):
            ctx.upload_counts_csc(Yc.indptr, bad, Yc.data)


@pytest.mark.gpu
def test_sparse_sink_equals_dense_get_res(tmp_path):
    from ruviiinb_b200 import api
    Y, M, ctl, truth = _fit_inputs()
    out = _state_out(Y, M, ctl, truth, 2)
    pac = api.get_res(out, type="quantile", block_size=200)
    s = api.get_res(out, type="quantile", block_size=200, sink="csc")
    assert sp.issparse(s) and s.format == "csc" and s.shape == pac.shape and s.has_sorted_indices
    assert np.array_equal(s.toarray(), pac) and s.nnz == np.count_nonzero(pac) < pac.size
    lg = api.get_res(out, type="quantile", block_size=200, sink="csc_log1p")
    assert np.allclose(lg.toarray(), np.log(pac + 1.0), rtol=0, atol=1e-12) and lg.nnz == s.nnz
    pear = api.get_res(out, type="pearson", block_size=200)
    sparse_pear = api.get_res(out, type="pearson", block_size=200, sink="csc")
    assert np.array_equal(sparse_pear.toarray(), pear)
    # a sparse `counts` (the dgCMatrix of the reference's output list) gives the same results
    out_sp = dict(out, counts=sp.csr_matrix(Y.astype(np.float64)))
    assert np.array_equal(api.get_res(out_sp, type="quantile", block_size=200), pac)
    # file-backed block sink: the same values, realised into a column-major .npy
    f = tmp_path / "pac.npy"
    mm = api.get_res(out, type="quantile", block_size=200, sink_file=str(f))
    assert isinstance(mm, np.memmap)
    del mm
    back = np.load(str(f), mmap_mode="r")
    assert back.shape == pac.shape and np.isfortran(back) and np.array_equal(np.asarray(back), pac)


@pytest.mark.gpu
def test_sparse_sink_capacity_error():
    from ruviiinb_b200 import _lib
    import ctypes as C
    Y, M, ctl, truth = _fit_inputs(G=30, N=256)
    out = _state_out(Y, M, ctl, truth, 2)
    G, N = Y.shape
    with _lib.Context(0) as ctx:
        ctx.set_design(G, N, np.argmax(M, axis=1).astype(np.int32), ctl)
        ctx.upload_counts(Y)
        for n in ("W", "gmean", "psi"):
            ctx.set_state(n, out[n])
        ctx.set_state("alpha", out["a"]); ctx.set_state("Mb", out["Mb"])
        cp = np.empty(N + 1, dtype=np.int64); ri = np.empty(4, dtype=np.int32); vv = np.empty(4)
        nnz = C.c_int64(0)
        rc = ctx.lib.ruvnb_get_res_csc(ctx.h, 2, 128, None, 0, 0, N, cp.ctypes.data, ri.ctypes.data, vv.ctypes.data, 4, C.byref(nnz), None)
        assert rc == -5 and nnz.value > 4      # RUVNB_ENOMEM, entries needed so far


@pytest.mark.gpu
def test_ruvIII_nb_sparse_in_sparse_out_and_makeSCE():
    from ruviiinb_b200 import api
    Y, M, ctl, truth = _fit_inputs(G=60, N=500, seed=8)
    alive = Y.sum(axis=1) > 0
    W0 = api.init_W(Y[alive] + 0.0, ctl[alive], 2)
    dense = api.ruvIII_nb(Y, M, ctl, k=2, outer_maxit=2, inner_maxit=4, W0=W0)
    sparse = api.ruvIII_nb(sp.csc_matrix(Y.astype(np.float64)), M, ctl, k=2, outer_maxit=2, inner_maxit=4, W0=W0)
    assert sp.issparse(sparse["counts"]) and np.array_equal(sparse["counts"].toarray(), dense["counts"])
    for n in ("W", "a", "gmean", "Mb", "psi", "logl"):
        assert np.array_equal(np.asarray(sparse[n]), np.asarray(dense[n])), n
    with pytest.raises(ValueError, match="cData"):
        api.makeSCE(sparse)
    with pytest.raises(ValueError, match="Wrong names"):
        api.makeSCE(sparse, cData={"cell": np.arange(Y.shape[1])}, assays=("foo",))
    sce = api.makeSCE(sparse, cData={"cell": np.arange(Y.shape[1])}, assays=("logcounts", "pearson", "logPAC"))
    assert set(sce["assays"]) == {"counts", "logcounts", "pearson", "logPAC"} and all(sp.issparse(v) for v in sce["assays"].values())
    assert np.allclose(sce["assays"]["logPAC"].toarray(), np.log(api.get_res(dense, type="quantile") + 1.0), atol=1e-12, rtol=0)
    assert np.array_equal(sce["assays"]["logcounts"].toarray(), api.get_res(dense, type="logcounts"))
    assert list(sce["rowData"]) == ["alpha1", "alpha2", "obj.psi"]
