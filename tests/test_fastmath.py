"""csrc/fastmath.cuh is __host__ __device__: the table-driven exp / log the sweeps use, evaluated on the host
through ruvnb_selftest_math, against 50-digit mpmath (sample) and libm (bulk)."""
import numpy as np

from ruviiinb_b200 import _lib


def test_fexp_bulk_and_edges():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-30, 30, 200000), rng.uniform(-699, 699, 200000), [0.0, -0.0, 1.0, -705.5, 709.7, 1e-300]])
    got = _lib.selftest_math("exp", x)
    ref = np.exp(x)
    rel = np.abs(got - ref) / ref
    assert rel.max() < 4.5e-16  # libm itself is within 1 ulp (2.2e-16)
    assert _lib.selftest_math("exp", np.array([0.0]))[0] == 1.0
    with np.errstate(over="ignore"):
        assert np.array_equal(_lib.selftest_math("exp", np.array([-745.5, 709.9])), np.exp(np.array([-745.5, 709.9])))
    sp = _lib.selftest_math("exp", np.array([np.inf, -np.inf, np.nan, 800.0, -800.0]))
    assert sp[0] == np.inf and sp[1] == 0.0 and np.isnan(sp[2]) and sp[3] == np.inf and sp[4] == 0.0


def test_fexp_vs_mpmath():
    import mpmath as mp
    mp.mp.dps = 40
    rng = np.random.default_rng(1)
    x = rng.uniform(-40, 40, 3000)
    got = _lib.selftest_math("exp", x)
    ref = np.array([float(mp.exp(mp.mpf(float(v)))) for v in x])
    assert (np.abs(got - ref) / ref).max() < 3e-16


def test_flog_bulk_and_edges():
    rng = np.random.default_rng(2)
    v = np.concatenate([np.exp(rng.uniform(-50, 50, 300000)), 1.0 + rng.uniform(-1e-3, 1e-3, 100000), [1.0, 2.0, 0.5, 1e-299, 1e299]])
    got = _lib.selftest_math("log", v)
    ref = np.log(v)
    assert (np.abs(got - ref) / np.maximum(1.0, np.abs(ref))).max() < 3e-16
    assert np.abs(got - ref).max() < 1.5e-14
    sp = _lib.selftest_math("log", np.array([0.0, -1.0, np.inf, np.nan, 5e-324]))
    assert sp[0] == -np.inf and np.isnan(sp[1]) and sp[2] == np.inf and np.isnan(sp[3]) and np.isclose(sp[4], np.log(5e-324))


def test_flog_nc2_is_flog_on_normal_arguments():
    """The cached APL sweep's logarithm (one 16-byte table load, no range check) returns flog's bits wherever flog takes
    its fast path."""
    from ruviiinb_b200 import _lib
    rng = np.random.default_rng(3)
    v = np.concatenate([np.exp(rng.uniform(-600, 600, 200000)), 1.0 + rng.uniform(-1e-3, 1e-3, 50000), [1.0, 2.0, 1e-290, 1e290]])
    assert np.array_equal(_lib.selftest_math("log_nc2", v), _lib.selftest_math("log", v))
