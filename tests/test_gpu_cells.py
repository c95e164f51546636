"""The cell-block machinery behind BASELINE configs 4 and 5 (get.res / fastruvIII.nb at 10^6 cells): sinks of ruvnb_get_res_ex,
cell ranges, the row mask, slab upload from device memory, cell gather, the histogram Winsorize of cell shards, and the
fused Mb.all -> PAC pass (ruvnb_fast_res) -- against the oracle and against the plain entry points."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ctx_with_fit(per_cell=False):
    from ruviiinb_b200 import _lib
    from tests.parity import make_ctx
    from tests.test_gpu_getres import _fitted
    Y, ctl, truth, W, a, gmean = _fitted()
    ctx = make_ctx(Y, truth["grp"], ctl)
    ctx.set_state("W", W); ctx.set_state("alpha", a); ctx.set_state("gmean", gmean)
    ctx.set_state("Mb", truth["Mb"]); ctx.set_state("psi", truth["psi"])
    return ctx, Y, truth, dict(W=W, a=a, gmean=gmean)


def test_sinks_agree_and_checksum():
    from ruviiinb_b200 import _lib
    ctx, Y, truth, _ = _ctx_with_fit()
    with ctx:
        ref = ctx.get_res(3, block_size=300)
        i32, st = ctx.get_res_ex(3, block_size=300, out_kind=_lib.OUT_I32)
        assert i32.dtype == np.int32 and np.array_equal(i32, ref.astype(np.int32))
        assert st["elements"] == ref.size and st["checksum"] == float(ref.sum()) and st["launches"] > 0 and st["device_seconds"] > 0
        lg, _ = ctx.get_res_ex(3, block_size=300, out_kind=_lib.OUT_LOG1P)
        assert np.allclose(lg, np.log(ref + 1.0), rtol=1e-15, atol=0)
        none, st0 = ctx.get_res_ex(3, block_size=300, out_kind=_lib.OUT_NONE)
        assert none is None and st0["checksum"] == st["checksum"]
        pear = ctx.get_res(1, block_size=256)
        _, stp = ctx.get_res_ex(1, block_size=256, out_kind=_lib.OUT_NONE)
        assert abs(stp["checksum"] - pear.sum()) <= 1e-9 * np.abs(pear).sum()
        with pytest.raises(_lib.RuvnbError, match="int32 / log1p"):
            ctx.get_res_ex(1, out_kind=_lib.OUT_I32)


def test_cell_range_and_row_mask():
    from ruviiinb_b200 import _lib
    ctx, Y, truth, _ = _ctx_with_fit()
    with ctx:
        full = {k: ctx.get_res(k, block_size=256) for k in (1, 3)}
        for k in (1, 3):
            part, _ = ctx.get_res_ex(k, block_size=256, cells=(512, 1100))
            assert np.array_equal(part, full[k][:, 512:1100])
        with pytest.raises(_lib.RuvnbError, match="block boundary"):
            ctx.get_res_ex(3, block_size=256, cells=(100, 400))
        alive = np.ones(Y.shape[0], bool); alive[[3, 40, 95]] = False
        ctx.set_row_mask(alive)
        pac, st = ctx.get_res_ex(3, block_size=256, out_kind=_lib.OUT_I32)
        assert pac.shape == (int(alive.sum()), Y.shape[1]) and np.array_equal(pac, full[3][alive].astype(np.int32))
        assert st["elements"] == pac.size
        # Pearson: the clip quantiles are taken over the reported rows only
        from oracle import getres
        ctx.set_row_mask(None)
    ctx2, Y, truth, p = _ctx_with_fit()
    with ctx2:
        ctx2.set_row_mask(alive)
        got, _ = ctx2.get_res_ex(1, block_size=256)
    ref = getres.get_res(dict(counts=Y[alive], a=p["a"][alive], W=p["W"], gmean=p["gmean"][alive], Mb=truth["Mb"][alive][:, truth["grp"]],
                              psi=truth["psi"][alive]), getres.PEARSON, block_size=256)
    assert np.abs(got - ref).max() <= 1e-10 * np.abs(ref).max()


def test_slab_upload_from_device_and_cell_gather():
    import torch
    from ruviiinb_b200 import _lib, synth
    Y, M, ctl, truth = synth.make_counts(70, 900, 2, 3, 12, seed=5)
    with _lib.Context(0) as ctx:
        ctx.set_design(70, 900, truth["grp"], ctl)
        ctx.upload_counts_rows(Y[:30], 0)                                   # a host slab
        t = torch.as_tensor(Y[30:], device="cuda:0").contiguous()           # a device slab
        ctx.upload_counts_rows(None, 30, device_ptr=t.data_ptr(), nrows=40)
        ctx.upload_counts_done()
        assert np.array_equal(ctx.get_counts(), Y)
        cells = np.array([899, 0, 17, 450, 17])
        assert np.array_equal(ctx.get_cells(cells), Y[:, cells])


def test_winsorize_histogram_path_matches(monkeypatch):
    """The Winsorize of cell shards (per-gene histograms summed over the ranks, radix descent for quantiles beyond the 4096
    direct bins) on one rank against the per-row kernel."""
    from ruviiinb_b200 import _lib, synth
    Y, M, ctl, truth = synth.make_counts(60, 2000, 2, 3, 10, seed=6)
    rng = np.random.default_rng(1)
    Y[4] = rng.poisson(9000.0, 2000)                       # quantile far beyond the direct bins
    Y[9, rng.choice(2000, 30, replace=False)] = 700000     # 1.5 % of the cells huge: both order statistics large, cap = 700000
    Y[11, :5] = 5000                                       # large values, but the quantile is small
    Y[13] = 0
    with _lib.Context(0) as ctx:
        ctx.set_design(60, 2000, truth["grp"], ctl); ctx.upload_counts(Y)
        cap0, nz0 = ctx.winsorize(); Y0 = ctx.get_counts()
    monkeypatch.setenv("RUVNB_WZ_HIST", "1")
    with _lib.Context(0) as ctx:
        ctx.set_design(60, 2000, truth["grp"], ctl); ctx.set_cell_shard(2000, 0); ctx.upload_counts(Y)
        cap1, nz1 = ctx.winsorize(); Y1 = ctx.get_counts()
    assert np.array_equal(cap0, cap1) and np.array_equal(nz0, nz1) and np.array_equal(Y0, Y1)
    assert cap0[9] == 700000 and cap0[4] > 9000 and nz0[13] == 0


def test_fast_res_fused_matches_two_step():
    """ruvnb_fast_res (Mb.all block -> caps -> PAC on the device) against ruvnb_fast_Mb_all followed by ruvnb_get_res with the
    per-cell Mb from the host, and against the oracle chain."""
    from oracle import fastruv, getres
    from ruviiinb_b200 import _lib
    ctx, Y, truth, p = _ctx_with_fit()
    N = Y.shape[1]
    annot = np.where(np.arange(N) % 3 == 0, truth["grp"], -1).astype(np.int32)
    with ctx:
        Mb_all = ctx.fast_Mb_all(annot, lambda_b=16.0, block_size=256)
        two = ctx.get_res(3, block_size=256, Mb_cells=Mb_all)
        pac, mb2, st = ctx.fast_res(3, annot, lambda_b=16.0, block_size=256, out_kind=_lib.OUT_I32, want_Mb_all=True)
        assert np.array_equal(mb2, Mb_all) and np.array_equal(pac, two.astype(np.int32)) and st["mb_ms"] > 0
        pear, _, _ = ctx.fast_res(1, annot, lambda_b=16.0, block_size=256)
        assert np.array_equal(pear, ctx.get_res(1, block_size=256, Mb_cells=Mb_all))
    ctl = np.zeros(Y.shape[0], bool); ctl[:24] = True
    refMb = fastruv.fast_Mb_all(Y.astype(np.float64), ctl, p["gmean"], p["a"], p["W"], truth["psi"], truth["Mb"], annot)
    ref = getres.get_res(dict(counts=Y, a=p["a"], W=p["W"], gmean=p["gmean"], Mb=refMb, psi=truth["psi"]), getres.QUANTILE, block_size=256)
    assert (pac != ref).mean() <= 1e-4


def test_two_rank_cell_shards_match_single_process():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29519", os.path.join(ROOT, "tests", "mp_cells_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
