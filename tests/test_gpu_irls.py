"""GPU parity: the CUDA path (through the C ABI) against the NumPy oracle on the same seeded inputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-9  # scaled error on per-gene / per-cell parameters after ONE iteration (FP64 everywhere)


def _check(res, tol=TOL):
    bad = {k: v for k, v in res["err"].items() if not v < tol}
    assert not bad, (bad, res["err"])


@pytest.mark.parametrize("k,m", [(1, 1), (2, 3), (5, 7)])
def test_one_iteration_matches_oracle(k, m):
    from tests.parity import one_iteration_case
    res = one_iteration_case(G=192, N=1400, k=k, m=m, n_ctl=48, seed=100 + k)
    _check(res)
    assert res["bad"] == res["oracle_bad"]


def test_one_iteration_exact_mad_path():
    """huber_fastpath = 0 forces the exact radix-select MAD for every gene; same answer."""
    from tests.parity import one_iteration_case
    res = one_iteration_case(G=96, N=900, k=2, m=3, n_ctl=32, seed=5, fastpath=0)
    _check(res)
    s, so = res["sigma"], res["oracle_sigma"]
    # exact order statistics of deviances evaluated with the table-driven log (abs error ~1e-15 on d^2/2)
    assert np.abs(s / so - 1.0).max() < 1e-10, np.abs(s / so - 1.0).max()


def test_one_iteration_robust_huber():
    """robust = TRUE (k.huber = 1.345): weights are active, exact sigma needed."""
    from tests.parity import one_iteration_case
    res = one_iteration_case(G=96, N=900, k=2, m=3, n_ctl=32, seed=6, robust=True)
    _check(res)


def test_one_iteration_sparse_genes_fail_certificate():
    """Very low means: deviances concentrate, some rows fail the certificate and take the exact path."""
    from tests.parity import one_iteration_case
    res = one_iteration_case(G=128, N=1100, k=2, m=3, n_ctl=32, seed=8, gmean_mu=-3.5)
    _check(res)


def test_one_iteration_high_counts_beyond_the_tables():
    """Means of 50-150 with dispersion ~0.3: most counts are >= 128, where the per-gene count tables end and the Stirling
    series / the general logarithm take over (kernels.cuh: nb_big_lgterm)."""
    from tests.parity import one_iteration_case
    res = one_iteration_case(G=128, N=1100, k=2, m=3, n_ctl=32, seed=11, gmean_mu=4.6)
    _check(res)
    assert res["bad"] == res["oracle_bad"]


def test_cell_major_upload_and_halved_steps():
    from tests.parity import one_iteration_case
    step = np.where(np.arange(160) % 3 == 0, 0.5, 1.0)
    res = one_iteration_case(G=160, N=700, k=3, m=4, n_ctl=40, seed=9, layout="cell_major", step=step)
    _check(res)


def test_full_fit_trajectory_matches_oracle():
    """Whole ruvIII.nb NB loop with injected W0 / psi: same accept/halve decisions, 1e-6 on the outputs."""
    from tests.parity import fit_case
    res = fit_case()
    got = [(r["outer"], r["iter"], r["degener"], r["accepted"]) for r in res["recs"]]
    ref = [(r["outer"], r["iter"], int(r["degener"]), int(r["accepted"])) for r in res["ref_trace"]]
    assert got == ref
    for a, b in zip(res["recs"], res["ref_trace"]):
        assert abs(a["logl_new"] - b["logl_new"]) <= 1e-9 * abs(b["logl_new"])
    bad = {k: v for k, v in res["err"].items() if not v < 1e-6}
    assert not bad, res["err"]
    assert res["stats"]["n_inner"] == res["n_inner_ref"]
    assert res["stats"]["kernel_launches"] > 0


def test_winsorize_matches_oracle():
    """H0 (R/ruvIIInb.R:53-57): per-gene type-7 99.5 % cap, ceiling, rowSums(Y > 0); incl. a row with counts >= 4096."""
    from oracle import rstats
    from ruviiinb_b200 import synth
    from tests.parity import make_ctx
    Y, M, ctl, truth = synth.make_counts(64, 1003, 2, 3, 16, seed=21, gmean_mu=1.0)
    rng = np.random.default_rng(5)
    Y[3] = rng.poisson(9000.0, size=Y.shape[1])   # radix-descent branch (max >= 4096)
    Y[4, :20] = 50000 + np.arange(20)             # a heavy upper tail: the cap interpolates between order statistics
    Y[5] = 0                                      # an all-zero gene (dropped by the caller at :57-62)
    Yw, cap = rstats.winsorize_counts(Y.astype(np.float64))
    with make_ctx(Y, truth["grp"], ctl) as ctx:
        assert np.array_equal(ctx.get_counts(), Y)
        got_cap, nz = ctx.winsorize()
        assert np.array_equal(got_cap, cap)
        assert np.array_equal(nz, (Yw > 0).sum(axis=1))
        assert np.array_equal(ctx.get_counts(), Yw.astype(np.int32))


def test_csr_upload_equals_dense():
    """BASELINE config 3 feeds CSR rows (ruvnb_upload_counts_csr): same resident counts as the dense upload."""
    import scipy.sparse as sp
    from ruviiinb_b200 import _lib, synth
    Y, M, ctl, truth = synth.make_counts(90, 777, 2, 4, 20, seed=71, zinb=True, gmean_mu=-1.0)
    csr = sp.csr_matrix(Y)
    with _lib.Context(0) as ctx:
        ctx.set_design(Y.shape[0], Y.shape[1], truth["grp"], ctl)
        ctx.upload_counts_csr(csr.indptr, csr.indices, csr.data)
        assert np.array_equal(ctx.get_counts(), Y)


@pytest.mark.parametrize("G,N,k,m,n_ctl", [(40, 130, 8, 2, 9), (33, 257, 1, 5, 3), (24, 4100, 4, 33, 8)])
def test_one_iteration_edge_shapes(G, N, k, m, n_ctl):
    """k at its maximum (8), N just past a chunk boundary, very few control genes, more groups than a warp has lanes."""
    from tests.parity import one_iteration_case
    res = one_iteration_case(G=G, N=N, k=k, m=m, n_ctl=n_ctl, seed=300 + k, gmean_mu=0.5)
    _check(res, tol=1e-8)


def test_argument_errors_are_reported_not_crashed():
    from ruviiinb_b200 import _lib, synth
    Y, M, ctl, truth = synth.make_counts(20, 200, 2, 3, 5, seed=1)
    with _lib.Context(0) as ctx:
        with pytest.raises(_lib.RuvnbError, match="no counts|set_design|W not set|state"):
            ctx.loglik(_lib.default_opts(k=2))
    with _lib.Context(0) as ctx:
        grp_bad = truth["grp"].copy(); grp_bad[grp_bad == 1] = 0      # replicate group 1 has no cell: legal for a cell shard
        ctx.set_design(20, 200, np.where(grp_bad == 2, 2, grp_bad), ctl)   # (get.res / fastruvIII.nb passes), refused by the fit
        ctx.upload_counts(Y)
        ctx.set_state("W", np.zeros((200, 2)))
        with pytest.raises(_lib.RuvnbError, match="has no cell"):
            ctx.loglik(_lib.default_opts(k=2))
    with _lib.Context(0) as ctx:
        ctx.set_design(20, 200, truth["grp"], ctl)
        ctx.upload_counts(Y)
        with pytest.raises(_lib.RuvnbError, match="outside 1..8"):
            ctx.set_state("W", np.zeros((200, 9)))
        with pytest.raises(_lib.RuvnbError, match="W not set"):
            ctx.irls_iteration(_lib.default_opts(k=2))
