"""CPU oracle for the RUV-III-NB hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, in NumPy float64, the algorithm of the reference R package
limfuxing/ruvIIInb 0.9.2.1 (pure R; `R/ruvIIInb.R`, `R/fastruvIIInb.R`, `R/get.res.R`,
`R/estimateDisp.par.R`).  It is the checker the CUDA path is compared against.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it.  Nothing under `ruviiinb_b200/` imports it, and the
product path raises when the CUDA library is missing rather than falling back to this code.

PARITY UNPINNED: the reference ships no tests, no golden vectors and no example data
(SURVEY.md section 4), R is not installed in the build container, and the arithmetic of
edgeR / limma / locfit / irlba / R's nmath lives in third-party packages that are not under
/root/reference.  The oracle is pinned only by (i) the reference source text, cited
file:line in every function, (ii) known-answer tests of the R `stats` semantics it restates
(`tests/test_oracle_rstats.py`) and (iii) cross-checks of the nmath restatements against
scipy.stats.nbinom (`tests/test_oracle_nmath.py`).
"""
