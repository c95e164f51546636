// sweeps.cuh -- launchers of the Y-streaming IRLS sweeps.  Each lives in its own translation unit (sweep_*.cu): the eight
// instantiations K = 1..8 of every sweep kernel dominate the build time, so they compile in parallel.  All launch on
// ctx->stream with ctx->launch_parts row parts (set by the caller) and count their launches in ctx->launches.
#pragma once
#include "launch.cuh"

// k_loglik: per-gene log-likelihood of `s` (+ certificate counters) into out[part][Gl]      (sweep_loglik.cu)
int launch_loglik(ruvnb_ctx* ctx, const RowGeom& g, const StateSet& s, double* out, int fastpath);
// k_ll_pass1: the same fused with the next iteration's Gram / score sums (ctx->RMW must hold RM_W of s.W)  (sweep_llp1.cu)
int launch_ll_pass1(ruvnb_ctx* ctx, const RowGeom& g, const StateSet& s, double* out, int fastpath);
// k_signdev: signed deviances of the listed rows into ctx->D (exact-MAD path)                (sweep_loglik.cu)
int launch_signdev(ruvnb_ctx* ctx, const RowGeom& g, const StateSet& s);
// k_pass1<HUBER>: Gram / score / group sums of `c` into ctx->A, u, sw, ZS, VS                (sweep_pass1.cu)
int launch_pass1(ruvnb_ctx* ctx, bool huber, const RowGeom& g, const StateSet& c, double kh);
// k_pass2<HUBER>: group sums S0, S1 of `c` against the new alpha / W of `n`                  (sweep_pass2.cu)
int launch_pass2(ruvnb_ctx* ctx, bool huber, const RowGeom& g, const StateSet& c, const StateSet& n, double kh);
// k_pass2w: S0 and T = sum w W_new per group from the weight cache ctx->Wc (into ctx->S0 / ctx->VS)     (sweep_pass2.cu)
int launch_pass2w(ruvnb_ctx* ctx, const RowGeom& g, const StateSet& n);
// per-cell W update over cells [pb, pe) (fast certificate kernel + exact kernel)             (sweep_w.cu)
int launch_w_update(ruvnb_ctx* ctx, const ruvnb_opts* o, const CtlPack& cp, const double* p_psi, const StateSet& c, StateSet& n,
                    double kh, long long pb, long long pe);
