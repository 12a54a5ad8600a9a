// C entry point of the tcgen05 GEMM core: argument checks, tile / cluster selection, TMA
// descriptors.  Device code lives in gemm_impl.cuh (instantiated in gemm_bf16.cu / gemm_f16.cu).
#include <stdlib.h>

#include "common.h"

#include "gemm_params.h"

namespace ub {
int gemm_dispatch_bf16(int bn, int cluster, int a_major, int b_major, const GemmParams& p,
                       const CUtensorMap& tmA, const CUtensorMap& tmB, int grid, cudaStream_t stream);
int gemm_dispatch_f16(int bn, int cluster, int a_major, int b_major, const GemmParams& p,
                      const CUtensorMap& tmA, const CUtensorMap& tmB, int grid, cudaStream_t stream);
int gemm_group_dispatch_bf16(const void* tm, const GroupedParams& g, int grid, cudaStream_t stream);
int gemm_group_dispatch_f16(const void* tm, const GroupedParams& g, int grid, cudaStream_t stream);
int launch_gemm_ln(int dtype, const GemmParams& p, const void* gamma, const void* beta, void* y,
                   long long ldy, const CUtensorMap& tmA, const CUtensorMap& tmB, cudaStream_t stream);

// Pick (N tile, CTAs per tile) minimising  waves x k-blocks x cycles-per-k-block + exposed tail.
// Cycles per k-block are MEASURED on B200 (K = 12288 sweep, mainloop only): they are far from
// proportional to the tile width — 1-SM: ~421 + 1.27*BN (583 @128, 745 @256); 2-SM pair:
// ~552 + 0.59*BN per 256-row pair tile (627 @128, 702 @256) — so fewer, wider tiles win until
// wave quantisation bites.  The epilogue of the last tile (~25 cycles per column of width) and
// a fixed launch / prologue cost are exposed once.
static void pick_config(int M, int N, int K, int sms, int* bn_out, int* cluster_out) {
  const int tiles_m = (M + BM - 1) / BM;
  const int num_kb = (K + BK - 1) / BK;
  const int cand[6][2] = {{256, 2}, {128, 2}, {256, 1}, {192, 1}, {128, 1}, {64, 1}};
  double best = 1e30;
  *bn_out = 128; *cluster_out = 1;
  for (int i = 0; i < 6; ++i) {
    const int bn = cand[i][0], c = cand[i][1];
    if (c == 2 && tiles_m < 2) continue;
    const int units = ((tiles_m + c - 1) / c) * ((N + bn - 1) / bn);
    const int slots = sms / c;
    const int waves = (units + slots - 1) / slots;
    const double per_kb = (c == 1) ? 421.0 + 1.27 * bn : 552.0 + 0.59 * bn;
    const double cost = static_cast<double>(waves) * num_kb * per_kb + 25.0 * bn + 3000.0;
    if (cost < best - 1e-9) { best = cost; *bn_out = bn; *cluster_out = c; }
  }
}

}  // namespace ub

extern "C" int ub200_gemm(const ub200_gemm_args* args, ub200_stream_t stream_) {
  using namespace ub;
  UB_CHECK_ARG(args != nullptr, "gemm: args is NULL");
  const ub200_gemm_args& a = *args;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  UB_CHECK_ARG(a.M > 0 && a.N > 0 && a.K > 0, "gemm: M, N, K must be positive (%d, %d, %d)", a.M,
               a.N, a.K);
  UB_CHECK_ARG(a.a != nullptr && a.b != nullptr && a.out != nullptr, "gemm: null operand");
  UB_CHECK_ARG(a.dtype == UB200_F16 || a.dtype == UB200_BF16, "gemm: bad dtype %d", a.dtype);
  UB_CHECK_ARG(a.N % 8 == 0 && a.ldo % 8 == 0, "gemm: N and ldo must be multiples of 8");
  const int epi = a.epilogue;
  UB_CHECK_ARG(!(epi & UB200_EPI_BIAS) || a.bias, "gemm: EPI_BIAS without bias");
  UB_CHECK_ARG(!(epi & UB200_EPI_RESIDUAL) || (a.residual && a.ldr % 8 == 0),
               "gemm: EPI_RESIDUAL needs residual with ldr %% 8 == 0");
  UB_CHECK_ARG(!(epi & UB200_EPI_GELU) || a.out2, "gemm: EPI_GELU without out2");
  UB_CHECK_ARG(!(epi & UB200_EPI_DGELU) || (a.aux && a.ldaux % 8 == 0),
               "gemm: EPI_DGELU needs aux with ldaux %% 8 == 0");
  UB_CHECK_ARG(!(epi & UB200_EPI_COLSUM) || (a.colsum && (reinterpret_cast<uintptr_t>(a.colsum) & 15) == 0),
               "gemm: EPI_COLSUM needs a 16-byte aligned colsum (vector atomics)");
  UB_CHECK_ARG(!((epi & UB200_EPI_GELU) && (epi & UB200_EPI_OUT_F32)),
               "gemm: EPI_GELU with fp32 output is not supported");
  UB_CHECK_ARG(a.dropout_p >= 0.f && a.dropout_p < 1.f, "gemm: dropout_p out of range");
  UB_CHECK_ARG(!(epi & UB200_EPI_ATOMIC) || epi == (UB200_EPI_ATOMIC | UB200_EPI_OUT_F32),
               "gemm: EPI_ATOMIC combines with EPI_OUT_F32 only");
  UB_CHECK_ARG(a.k_splits >= -1, "gemm: k_splits must be >= -1");
  UB_CHECK_ARG(a.k_splits == 0 || a.k_splits == 1 || (epi & UB200_EPI_ATOMIC),
               "gemm: k_splits > 1 needs EPI_ATOMIC | EPI_OUT_F32 and a zeroed output");
  UB_CHECK_ARG(a.n_valid >= 0 && a.n_valid <= a.N, "gemm: n_valid must be in [0, N]");
  const int n_valid = a.n_valid ? a.n_valid : a.N;

  if (epi & UB200_EPI_LN) {
    // fused residual + LayerNorm epilogue: its own kernel (4-CTA cluster over N), see gemm_ln.cu
    const int allowed = UB200_EPI_LN | UB200_EPI_BIAS | UB200_EPI_RESIDUAL | UB200_EPI_DROPOUT;
    UB_CHECK_ARG((epi & ~allowed) == 0 && (epi & UB200_EPI_BIAS) && (epi & UB200_EPI_RESIDUAL),
                 "gemm: EPI_LN combines with BIAS | RESIDUAL [| DROPOUT] only");
    UB_CHECK_ARG(a.a_major == 0 && a.b_major == 0, "gemm: EPI_LN needs K-major operands");
    UB_CHECK_ARG(a.ln_gamma && a.ln_beta && a.ln_out && a.ldln % 8 == 0, "gemm: EPI_LN needs ln_gamma / ln_beta / ln_out");
    if (a.N != 768 && a.N != 1024)
      return set_error(UB200_EUNSUPPORTED, "gemm: EPI_LN needs N = 768 or 1024 (got %d)", a.N);
    const int bnl = a.N / 4;
    CUtensorMap tmA, tmB;
    int rc = make_tma_2d(&tmA, a.a, a.dtype, a.M, a.K, a.lda, BM, BK);
    if (rc) return rc;
    rc = make_tma_2d(&tmB, a.b, a.dtype, a.N, a.K, a.ldb, bnl, BK);
    if (rc) return rc;
    GemmParams p{};
    p.M = a.M; p.N = a.N; p.K = a.K;
    p.epilogue = epi;
    p.bias = a.bias; p.residual = a.residual; p.out = a.out;
    p.ldr = a.ldr; p.ldo = a.ldo;
    if ((epi & UB200_EPI_DROPOUT) && a.dropout_p > 0.f) {
      uint32_t thr = static_cast<uint32_t>(a.dropout_p * 65536.0f + 0.5f);
      if (thr > 65535u) thr = 65535u;
      p.drop_thr16 = thr;
      p.drop_inv_keep = 65536.0f / static_cast<float>(65536u - thr);
    } else {
      p.epilogue &= ~UB200_EPI_DROPOUT;
      p.drop_thr16 = 0;
      p.drop_inv_keep = 1.f;
    }
    p.seed_lo = static_cast<uint32_t>(a.rng_seed);
    p.seed_hi = static_cast<uint32_t>(a.rng_seed >> 32);
    p.stream_lo = static_cast<uint32_t>(a.rng_stream);
    p.stream_hi = static_cast<uint32_t>(a.rng_stream >> 32);
    p.rng_dev = reinterpret_cast<const unsigned long long*>(a.rng_offset_dev);
    p.tiles_m = (a.M + BM - 1) / BM;
    p.tiles_n = 4;
    p.ksplit = 1;
    p.kb_per_split = (a.K + BK - 1) / BK;
    return launch_gemm_ln(a.dtype, p, a.ln_gamma, a.ln_beta, a.ln_out, a.ldln, tmA, tmB, stream);
  }

  const int sms = num_sms();
  int bn = 0, cluster = 0;
  pick_config(a.M, a.N, a.K, sms, &bn, &cluster);
  if (a.tile_n) {
    bn = a.tile_n;
    if (!a.cluster) cluster = ((bn == 128 || bn == 256) && a.M > BM) ? 2 : 1;
  }
  if (a.cluster) cluster = a.cluster;
  const bool splitk = a.k_splits > 1 || a.k_splits == -1;
  if (splitk) {              // split-K units live in the 1-SM kernel
    cluster = 1;
    if (!a.tile_n) bn = a.N >= 128 ? 128 : 64;
  }
  UB_CHECK_ARG(cluster == 1 || cluster == 2, "gemm: cluster must be 0, 1 or 2 (got %d)", cluster);
  UB_CHECK_ARG(!(cluster == 2 && (bn == 64 || bn == 192)), "gemm: cluster 2 needs tile_n 128 or 256");

  CUtensorMap tmA, tmB;
  int rc;
  if (a.a_major == 0)
    rc = make_tma_2d(&tmA, a.a, a.dtype, a.M, a.K, a.lda, BM, BK);
  else
    rc = make_tma_2d(&tmA, a.a, a.dtype, a.K, a.M, a.lda, BK, 64);
  if (rc) return rc;
  if (a.b_major == 0)  // in 2-SM mode each CTA stages half of the B rows
    rc = make_tma_2d(&tmB, a.b, a.dtype, n_valid, a.K, a.ldb, bn / cluster, BK);
  else
    rc = make_tma_2d(&tmB, a.b, a.dtype, a.K, n_valid, a.ldb, BK, 64);
  if (rc) return rc;

  GemmParams p;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.epilogue = epi;
  p.bias = a.bias; p.residual = a.residual; p.aux = a.aux;
  p.out = a.out; p.out2 = a.out2; p.colsum = a.colsum;
  p.ldr = a.ldr; p.ldaux = a.ldaux; p.ldo = a.ldo;
  if ((epi & UB200_EPI_DROPOUT) && a.dropout_p > 0.f) {
    uint32_t thr = static_cast<uint32_t>(a.dropout_p * 65536.0f + 0.5f);
    if (thr > 65535u) thr = 65535u;
    p.drop_thr16 = thr;
    p.drop_inv_keep = 65536.0f / static_cast<float>(65536u - thr);
  } else {
    p.epilogue &= ~UB200_EPI_DROPOUT;
    p.drop_thr16 = 0;
    p.drop_inv_keep = 1.f;
  }
  p.seed_lo = static_cast<uint32_t>(a.rng_seed);
  p.seed_hi = static_cast<uint32_t>(a.rng_seed >> 32);
  p.stream_lo = static_cast<uint32_t>(a.rng_stream);
  p.stream_hi = static_cast<uint32_t>(a.rng_stream >> 32);
  p.rng_dev = reinterpret_cast<const unsigned long long*>(a.rng_offset_dev);
  p.tiles_m = (a.M + BM - 1) / BM;
  p.tiles_n = (a.N + bn - 1) / bn;
  const int num_kb = (a.K + BK - 1) / BK;
  p.ksplit = 1;
  p.kb_per_split = num_kb;
  if (splitk) {
    const int tiles = p.tiles_m * p.tiles_n;
    int want = a.k_splits == -1 ? (sms + tiles - 1) / tiles : a.k_splits;
    if (want > num_kb) want = num_kb;
    if (want < 1) want = 1;
    p.kb_per_split = (num_kb + want - 1) / want;
    p.ksplit = (num_kb + p.kb_per_split - 1) / p.kb_per_split;   // every slice has >= 1 k-block
  }
  const int units = ((p.tiles_m + cluster - 1) / cluster) * p.tiles_n * p.ksplit;
  int slots = (a.max_ctas > 0 ? a.max_ctas : sms) / cluster;
  if (slots < 1) slots = 1;
  if (slots > units) slots = units;
  const int grid = slots * cluster;

  if (a.dtype == UB200_BF16)
    return gemm_dispatch_bf16(bn, cluster, a.a_major, a.b_major, p, tmA, tmB, grid, stream);
  return gemm_dispatch_f16(bn, cluster, a.a_major, a.b_major, p, tmA, tmB, grid, stream);
}

extern "C" int ub200_gemm_grouped(const ub200_gemm_args* args, int32_t count, ub200_stream_t stream_) {
  using namespace ub;
  UB_CHECK_ARG(args != nullptr && count >= 1 && count <= GEMM_MAX_GROUP,
               "gemm_grouped: need 1..%d problems", GEMM_MAX_GROUP);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  struct { CUtensorMap a[GEMM_MAX_GROUP]; CUtensorMap b[GEMM_MAX_GROUP]; } tm;
  GroupedParams g{};
  g.nprob = count; g.K = args[0].K; g.epilogue = args[0].epilogue;
  for (int i = 0; i < count; ++i) {
    const ub200_gemm_args& a = args[i];
    UB_CHECK_ARG(a.a && a.b && a.out, "gemm_grouped[%d]: null operand", i);
    UB_CHECK_ARG(a.a_major == 1 && a.b_major == 1, "gemm_grouped[%d]: operands must be MN-major (wgrad form)", i);
    UB_CHECK_ARG(a.K == g.K && a.dtype == args[0].dtype && a.epilogue == g.epilogue,
                 "gemm_grouped[%d]: K / dtype / epilogue must match problem 0", i);
    UB_CHECK_ARG(a.M > 0 && a.N > 0 && a.K > 0 && a.N % 8 == 0 && a.ldo % 8 == 0,
                 "gemm_grouped[%d]: bad shape", i);
  }
  // N tile shared by the group: fewest (rounds over the SMs) x (measured cycles per k-block of a
  // 128 x bn tile, see pick_config) — 4 base-layer wgrads: 432 tiles of 128 = 3 rounds x 583,
  // 288 tiles of 192 = 2 rounds x 665, 216 tiles of 256 = 2 rounds x 745.
  const int sms = num_sms();
  int bn = args[0].tile_n;
  if (bn == 0) {
    static const int env_bn = [] { const char* e = getenv("UB200_GROUP_BN"); return e ? atoi(e) : 0; }();
    bn = env_bn;
  }
  if (bn == 0) {
    double best = 1e30;
    const int cand[3] = {128, 192, 256};
    for (int c = 0; c < 3; ++c) {
      int t = 0;
      for (int i = 0; i < count; ++i) t += ((args[i].M + BM - 1) / BM) * ((args[i].N + cand[c] - 1) / cand[c]);
      const double cost = static_cast<double>((t + sms - 1) / sms) * (421.0 + 1.27 * cand[c]);
      if (cost < best - 1e-9) { best = cost; bn = cand[c]; }
    }
  }
  UB_CHECK_ARG(bn == 128 || bn == 192 || bn == 256, "gemm_grouped: tile_n must be 0, 128, 192 or 256 (got %d)", bn);
  g.bn = bn;
  int tiles = 0;
  for (int i = 0; i < count; ++i) {
    const ub200_gemm_args& a = args[i];
    int rc = make_tma_2d(&tm.a[i], a.a, a.dtype, a.K, a.M, a.lda, BK, 64);
    if (rc) return rc;
    rc = make_tma_2d(&tm.b[i], a.b, a.dtype, a.K, a.N, a.ldb, BK, 64);
    if (rc) return rc;
    g.M[i] = a.M; g.N[i] = a.N; g.out[i] = a.out; g.ldo[i] = a.ldo;
    g.tiles_n[i] = (a.N + bn - 1) / bn;
    g.tile_start[i] = tiles;
    tiles += ((a.M + BM - 1) / BM) * g.tiles_n[i];
  }
  for (int i = count; i <= GEMM_MAX_GROUP; ++i) g.tile_start[i] = tiles;
  for (int i = count; i < GEMM_MAX_GROUP; ++i) { tm.a[i] = tm.a[0]; tm.b[i] = tm.b[0]; }
  int grid = sms;
  if (grid > tiles) grid = tiles;
  if (args[0].dtype == UB200_BF16) return gemm_group_dispatch_bf16(&tm, g, grid, stream);
  return gemm_group_dispatch_f16(&tm, g, grid, stream);
}
