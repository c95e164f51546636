// svd.cu -- H1: initial W (R/ruvIIInb.R:160-168) by subspace iteration on the resident control rows
#include "ctx.cuh"
#include "launch.cuh"
#include "svd.cuh"

// ---- H1: initial W (R/ruvIIInb.R:160-168) by subspace iteration on the resident control rows (svd.cuh) -----------------
extern "C" int ruvnb_init_W(ruvnb_ctx* ctx, int k, int max_iter, double tol, double* W_out, double* sv_out, int* iters_out) {
  if (!ctx) return RUVNB_EARG;
  if (!ctx->have_counts) return ctx->fail(RUVNB_ESTATE, "init_W: no counts uploaded");
  RET(ensure_state(ctx, k));
  CK(cudaSetDevice(ctx->device));
  const int K = ctx->K, nc = ctx->n_ctl, KK = K * (K + 1) / 2; const long long Np = ctx->Np;
  if (nc < K) return ctx->fail(RUVNB_EARG, "init_W: %d control genes cannot carry %d factors", nc, K);
  if (max_iter <= 0) max_iter = 100;
  if (!(tol > 0)) tol = 1e-9;   // on the Ritz values, i.e. ~3e-5 on the vectors: irlba's own tolerance is 1e-5
  double *rm = nullptr, *U = nullptr, *G = nullptr, *T = nullptr, *A = nullptr, *Upart = nullptr;
  // cell parts of the A V product: enough CTAs for ~4 per SM, whole 128-cell chunks per part
  const int row_blocks = (nc + 8 * SVD_RPW - 1) / (8 * SVD_RPW);
  int n_sm = 148;
  CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->device));
  int parts = std::max(1, std::min((int)(Np / CH), (4 * n_sm + row_blocks - 1) / row_blocks));
  const long long seg = ((Np / CH + parts - 1) / parts) * CH;
  parts = (int)((Np + seg - 1) / seg);
  ScopedDev scratch;
  CK(scratch.alloc(&rm, (size_t)nc * 8)); CK(scratch.alloc(&U, (size_t)nc * K * 8)); CK(scratch.alloc(&G, KK * 8)); CK(scratch.alloc(&T, K * K * 8));
  CK(scratch.alloc(&Upart, (size_t)parts * nc * K * 8));
  if (scratch.alloc(&A, (size_t)nc * Np * 8) != cudaSuccess) {
    cudaGetLastError();
    return ctx->fail(RUVNB_ENOMEM, "init_W: no room for the %d x %lld FP64 slab of log-ratios (%.1f GB)", nc, Np, (double)nc * Np * 8 / 1e9);
  }
  double* V = ctx->cur.W; double* V2 = ctx->nw.W;
  k_svd_rowmean<<<nblk(nc, 8), 256, 0, ctx->stream>>>(ctx->d_ctl_rows, nc, Np, (double)ctx->N, rm); CKL();
  k_svd_build<<<nblk((long long)nc * Np, 256), 256, 0, ctx->stream>>>(ctx->d_ctl_rows, nc, Np, rm, A); CKL();
  auto av = [&](const double* Vin) -> int {
    DISPATCH_K(K, { k_svd_av<KT><<<dim3(row_blocks, parts), 256, 0, ctx->stream>>>(A, nc, Np, seg, Vin, Upart); CKL(); });
    k_svd_sum_parts<<<nblk((long long)nc * K, 256), 256, 0, ctx->stream>>>(Upart, (long long)nc * K, parts, U); CKL();
    return RUVNB_OK;
  };
  k_svd_start<<<nblk(Np * K, 256), 256, 0, ctx->stream>>>(V, Np, K, ctx->d_pos2cell); CKL();
  std::vector<double> Gh(KK), Th(K * K), prev(K, 0.0), diag(K, 0.0);
  // orthonormalise the K vectors in `X` in place: Gram -> Cholesky G = R'R on the host -> X <- X R^-1
  auto orth = [&](double* X) -> int {
    k_svd_gram<<<KK, 1024, 0, ctx->stream>>>(X, Np, K, G); CKL();
    CK(cudaMemcpyAsync(Gh.data(), G, KK * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<double> R(K * K, 0.0);   // upper triangular, row-major
    for (int j = 0; j < K; ++j) {
      for (int i = 0; i <= j; ++i) {
        double s = Gh[j * (j + 1) / 2 + i];
        for (int q = 0; q < i; ++q) s -= R[q * K + i] * R[q * K + j];
        if (i == j) { if (!(s > 0.0)) return ctx->fail(RUVNB_EARG, "init_W: the control-gene matrix has rank < k"); R[j * K + j] = sqrt(s); diag[j] = s; }
        else R[i * K + j] = s / R[i * K + i];
      }
    }
    // Rinv (upper triangular)
    for (auto& x : Th) x = 0.0;
    for (int j = 0; j < K; ++j) {
      Th[j * K + j] = 1.0 / R[j * K + j];
      for (int i = j - 1; i >= 0; --i) {
        double s = 0.0;
        for (int q = i + 1; q <= j; ++q) s -= R[i * K + q] * Th[q * K + j];
        Th[i * K + j] = s / R[i * K + i];
      }
    }
    CK(cudaMemcpyAsync(T, Th.data(), K * K * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_svd_apply<<<nblk(Np, 256), 256, 0, ctx->stream>>>(X, Np, K, T); CKL();
    return RUVNB_OK;
  };
  RET(orth(V));
  int it = 0;
  for (; it < max_iter; ++it) {
    RET(av(V));
    DISPATCH_K(K, { k_svd_atu<KT><<<(unsigned)(Np / 32), 256, 0, ctx->stream>>>(A, nc, Np, U, V2); CKL(); });
    std::swap(V, V2);
    RET(orth(V));   // diag = squared norms of A'A v before normalisation ~ sigma^4 along converged directions
    double ch = 0.0;
    for (int l = 0; l < K; ++l) { ch = std::max(ch, fabs(diag[l] - prev[l]) / std::max(diag[l], 1e-300)); prev[l] = diag[l]; }
    if (it >= 2 && ch < tol) { ++it; break; }
  }
  // Rayleigh-Ritz: B = A V, eigen-decomposition of B'B -> singular values and the rotation of V
  RET(av(V));
  std::vector<double> Uh((size_t)nc * K), BtB(K * K, 0.0), Q;
  CK(cudaMemcpyAsync(Uh.data(), U, Uh.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (int a = 0; a < K; ++a) for (int b = 0; b <= a; ++b) {
    double s = 0.0;
    for (int g = 0; g < nc; ++g) s += Uh[(size_t)g * K + a] * Uh[(size_t)g * K + b];
    BtB[a * K + b] = BtB[b * K + a] = s;
  }
  jacobi_eig_h(K, BtB, Q);
  CK(cudaMemcpyAsync(T, Q.data(), K * K * 8, cudaMemcpyHostToDevice, ctx->stream));
  k_svd_apply<<<nblk(Np, 256), 256, 0, ctx->stream>>>(V, Np, K, T); CKL();
  if (V != ctx->cur.W) { std::swap(ctx->cur.W, ctx->nw.W); }   // the result lives in cur.W
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->have_W = true; ctx->cur.sig_valid = false;
  // sign convention on the host copy, then written back so that "W" and W_out agree
  const long long N = ctx->N;
  std::vector<double> Wh((size_t)N * K);
  RET(ruvnb_get_state(ctx, "W", Wh.data()));
  for (int l = 0; l < K; ++l) {
    long long am = 0;
    for (long long c = 1; c < N; ++c) if (fabs(Wh[(size_t)l * N + c]) > fabs(Wh[(size_t)l * N + am])) am = c;
    if (Wh[(size_t)l * N + am] < 0) for (long long c = 0; c < N; ++c) Wh[(size_t)l * N + c] = -Wh[(size_t)l * N + c];
  }
  RET(ruvnb_set_state(ctx, "W", Wh.data(), K));
  if (W_out) memcpy(W_out, Wh.data(), Wh.size() * 8);
  if (sv_out) for (int l = 0; l < K; ++l) sv_out[l] = sqrt(std::max(BtB[l * K + l], 0.0));
  if (iters_out) *iters_out = it;
  return RUVNB_OK;   // scratch released by its scope
}
