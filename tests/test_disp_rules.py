"""The dispersion step's scoring sweeps differ from edgeR's glm_one_group in three documented ways:
the score is accumulated as S1 - e^beta S2, the start value is a neighbouring grid point's (or an earlier call's)
intercept, and a column stops one sweep early when the next step is predicted below the tolerance.  This file restates
those rules in NumPy (what disp.cuh's OpDispNR does per row and column) and checks the claim they rest on: the intercept
they return is the oracle's -- edgeR's -- to within its own tolerance.  CPU only."""
import numpy as np

from oracle import edger


def device_rule_fit(y, offset, phi, start, maxit=50, tol=1e-10):
    """One row, one dispersion: Fisher scoring as the kernel accumulates it, with the kernel's stopping rule.
    Returns (beta, sweeps)."""
    E = np.exp(offset)
    beta, prev = float(start), 0.0
    for it in range(maxit):
        eb = np.exp(beta)
        w = 1.0 / (1.0 + (eb * phi) * E)
        s1, s2 = float((y * w).sum()), eb * float((E * w).sum())
        step = (s1 - s2) / s2
        beta += step
        conv = abs(step) < tol
        if not conv and prev != 0.0 and abs(step) < 1e-8:
            rho = abs(step / prev)
            conv = rho < 0.5 and rho * abs(step) < tol
        prev = step
        if conv:
            return beta, it + 1
    return beta, maxit


def _rows(seed, n=4000):
    rng = np.random.default_rng(seed)
    cases = []
    for mean, phi_true in ((0.05, 0.5), (0.8, 0.3), (6.0, 0.1), (150.0, 0.05), (2.0, 3.0)):
        offset = 0.7 * rng.standard_normal(n)
        mu = mean * np.exp(offset)
        y = rng.negative_binomial(1.0 / phi_true, 1.0 / (1.0 + phi_true * mu)).astype(np.float64)
        cases.append((y, offset))
    return cases


def test_interleaved_warm_started_grid_reaches_edgers_intercepts():
    for y, offset in _rows(11):
        ref = np.array([edger.glm_one_group(y[None, :], offset[None, :], d)[0] for d in edger.GRID_DISP])
        beta0 = np.log((y / np.exp(offset)).sum() / y.size)            # edgeR's start (glm_one_group)
        got = np.full(21, np.nan)
        sweeps = np.zeros(21, dtype=int)
        for c in range(3):                                              # {0,3,..}, {1,4,..}, {2,5,..}
            for d in range(c, 21, 3):
                start = beta0 if c == 0 else got[d - 1]
                got[d], sweeps[d] = device_rule_fit(y, offset, edger.GRID_DISP[d], start)
        assert np.abs(got - ref).max() < 3e-10, np.abs(got - ref).max()
        assert sweeps.max() < 50
        # the warm-started sets need fewer sweeps than the cold one
        assert sweeps[1::3].mean() <= sweeps[0::3].mean() and sweeps[2::3].mean() <= sweeps[0::3].mean()


def test_start_value_does_not_move_the_root():
    y, offset = _rows(12)[2]
    phi = 0.4
    ref = edger.glm_one_group(y[None, :], offset[None, :], phi)[0]
    for start in (ref - 2.0, ref - 0.3, ref + 1e-6, ref + 0.3, ref + 2.0):
        beta, _ = device_rule_fit(y, offset, phi, start)
        assert abs(beta - ref) < 3e-10


def test_from_far_above_the_root_scoring_walks_down_by_one_per_sweep():
    """Why an intercept left over from very different parameters is not reused (k_start_cols' 2.0 guard)."""
    y, offset = _rows(13)[1]
    phi = 0.3
    ref = edger.glm_one_group(y[None, :], offset[None, :], phi)[0]
    E = np.exp(offset)
    beta = ref + 60.0
    for _ in range(5):
        eb = np.exp(beta)
        w = 1.0 / (1.0 + (eb * phi) * E)
        s1, s2 = (y * w).sum(), eb * (E * w).sum()
        step = (s1 - s2) / s2
        assert -1.0 <= step < -0.999
        beta += step


def test_apl_from_histogram_row_sums_and_scoring_information():
    """OpDispAPL's decomposition: sum lgamma terms = histogram . table, sum y log mu = beta sum y + sum y offset, and the
    Cox-Reid information taken from the LAST scoring sweep (evaluated before its step is applied)."""
    from scipy import special
    for y, offset in _rows(21, n=3000):
        E = np.exp(offset)
        hist = np.bincount(y.astype(np.int64), minlength=int(y.max()) + 1).astype(np.float64)
        sy, syo = y.sum(), (y * offset).sum()
        for phi in edger.GRID_DISP[::4]:
            r = 1.0 / phi
            # the scoring sweeps, keeping the information of the last one
            beta, prev, info = np.log((y / E).sum() / y.size), 0.0, np.nan
            for _ in range(50):
                eb = np.exp(beta)
                w = 1.0 / (1.0 + (eb * phi) * E)
                s1, s2 = (y * w).sum(), eb * (E * w).sum()
                step = (s1 - s2) / s2
                info = s2
                beta += step
                conv = abs(step) < 1e-10
                if not conv and prev != 0.0 and abs(step) < 1e-8:
                    rho = abs(step / prev)
                    conv = rho < 0.5 and rho * abs(step) < 1e-10
                prev = step
                if conv:
                    break
            yy = np.arange(hist.size, dtype=np.float64)
            table = special.gammaln(yy + r) - special.gammaln(r) - special.gammaln(yy + 1.0)   # table[0] = 0
            lg = (hist * table).sum()
            ll = lg + (beta * sy + syo) + y.size * r * np.log(r) - ((y + r) * np.log(np.exp(beta) * E + r)).sum()
            got = ll - 0.5 * np.log(max(info, 1e-10))
            ref = edger.adjusted_profile_lik(phi, y[None, :], offset[None, :])[0]
            assert abs(got - ref) <= 1e-9 * abs(ref), (phi, got, ref)
