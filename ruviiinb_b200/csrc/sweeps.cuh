// sweeps.cuh -- launchers of the Y-streaming IRLS sweeps.  Each lives in its own translation unit (sweep_*.cu): the eight
// instantiations K = 1..8 of every sweep kernel dominate the build time, so they compile in parallel.  All launch on
// ctx->stream with ctx->launch_parts row parts (set by the caller) and count their launches in ctx->launches.
#pragma once
#include "launch.cuh"

// k_loglik: per-gene log-likelihood of `s` (+ certificate counters) into out[part][Gl]      (sweep_loglik.cu)
int launch_loglik(ruvnb_ctx* ctx, const RowGeom& g, const StateSet& s, double* out, int fastpath);
// k_ll_w: the same fused with the element-wise half of the next iteration's get.coef2: w -> ctx->Wc, w Z -> ctx->WZc, sum Z per
// group -> ctx->ZS                                                                            (sweep_llp1.cu)
int launch_ll_w(ruvnb_ctx* ctx, const RowGeom& g, const StateSet& s, double* out, int fastpath);
// k_gram: Gram / score / group sums from the two slabs (ctx->RMW must hold RM_W of the state's W) into ctx->A, u, sw, VS, S0c, SWZ
int launch_gram(ruvnb_ctx* ctx, const RowGeom& g);
// k_signdev: signed deviances of the listed rows into ctx->D (exact-MAD path)                (sweep_loglik.cu)
int launch_signdev(ruvnb_ctx* ctx, const RowGeom& g, const StateSet& s);
// k_pass1<HUBER>: Gram / score / group sums of `c` into ctx->A, u, sw, ZS, VS (+ ctx->Wc, S0c, SWZ)   (sweep_pass1.cu)
int launch_pass1(ruvnb_ctx* ctx, bool huber, const RowGeom& g, const StateSet& c, double kh);
// k_pass2<HUBER>: group sums S0, S1 of `c` against the new alpha / W of `n`                  (sweep_pass2.cu)
int launch_pass2(ruvnb_ctx* ctx, bool huber, const RowGeom& g, const StateSet& c, const StateSet& n, double kh);
// k_pass2w: T = sum w W_new per group from the weight cache ctx->Wc (into ctx->VS)           (sweep_pass2.cu)
int launch_pass2w(ruvnb_ctx* ctx, const RowGeom& g, const StateSet& n);
// per-cell W update over cells [pb, pe) (fast certificate kernel + exact kernel)             (sweep_w.cu)
int launch_w_update(ruvnb_ctx* ctx, const ruvnb_opts* o, const CtlPack& cp, const double* p_psi, const StateSet& c, StateSet& n,
                    double kh, long long pb, long long pe);
