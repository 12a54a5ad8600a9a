// GEMM with the residual + LayerNorm epilogue fused in (north star: "residual+LayerNorm epilogues"):
//
//   s = dropout(A W^T + bias) + residual          (model/layer.py:112-113, :153-154)
//   y = LayerNorm(s) * gamma + beta               (:114, :155; eps 1e-12, biased variance, fp32 stats)
//
// One kernel replaces GEMM(+bias+dropout+residual) -> s to HBM -> ln_fwd_kernel (read s, write y):
// the row statistics need the whole row of N = H columns, which is wider than one CTA's accumulator
// (128 x 768 fp32 = 768 TMEM columns > 512), so the row is split over a CLUSTER of 4 CTAs along N
// (H = 768: 4 x 192, H = 1024: 4 x 256) that exchange per-row (mean, M2) through distributed shared
// memory:
//
//   pass 1  (epilogue warps, coalesced 4-lanes-per-row mapping of gemm_impl.cuh)
//           v = acc + bias -> dropout -> + residual -> round to 16 bit -> store s (saved for the
//           backward) and accumulate (count, mean, M2) of the ROUNDED values per row (Chan's
//           parallel update: no E[x^2] - E[x]^2 cancellation);
//   merge   4 lanes of a row (shuffles) -> the 2-3 epilogue warps sharing the row (smem) ->
//           barrier.cluster -> the 4 CTAs of the row (ld.shared::cluster) -> mean, rstd;
//   pass 2  re-read the CTA's own s (L2-hot, written by the same thread), normalise, scale, shift,
//           store y.
//
// The mainloop is the 1-SM pipeline of gemm_kernel (TMA producer warp, one MMA-issuing lane, TMEM
// accumulator); one 128 x BN tile per CTA (27 row tiles x 4 = 108 CTAs at C2 = one wave).
#include "common.h"
#include "gemm_impl.cuh"

namespace ub {

struct LnEpiParams {
  const void* gamma;   // [N] 16-bit
  const void* beta;    // [N] 16-bit
  void* y;             // [M, N] 16-bit, pitch ldy
  long long ldy;
  float inv_n;         // 1 / N
};

constexpr int LN_CLUSTER = 4;
constexpr float LN_FUSED_EPS = 1e-12f;

template <int BN>
struct GemmLnCfg {
  static constexpr int B_TILE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int STAGES = BN == 192 ? 4 : 3;
  static constexpr int TMEM_COLS = 256;
  static constexpr int EPI_SPLIT = epi_split(BN);          // warps sharing a TMEM lane quarter
  static constexpr int EPI_WARPS = 4 * EPI_SPLIT;
  static constexpr int THREADS = 128 + 32 * EPI_WARPS;
  static constexpr int CHUNKS = BN / 32 / EPI_SPLIT;        // 32-column blocks per epilogue warp
  static constexpr int BAR_BYTES = 256;
  static constexpr int EPI_STAGE_BYTES = EPI_WARPS * 32 * 33 * 4;
  static constexpr int STAT_BYTES = (EPI_SPLIT + 2) * BM * 2 * 4;   // wstat[SPLIT] + cstat + fin
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + EPI_STAGE_BYTES + STAT_BYTES + 1024;
};

// (na, ma, M2a) <- merge with (nb, mb, M2b)
__device__ __forceinline__ void chan_merge(float& na, float& ma, float& M2a, float nb, float mb, float M2b) {
  const float n = na + nb;
  const float d = mb - ma;
  const float f = nb / n;
  ma = fmaf(d, f, ma);
  M2a = M2a + M2b + d * d * na * f;
  na = n;
}

__device__ __forceinline__ float2 ld_dsmem_f2(uint32_t cluster_addr) {
  float2 v;
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(cluster_addr) : "memory");
  return v;
}

__device__ __forceinline__ void epi_bar_sync(int nthreads) {
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

template <int BN, bool kBF16, bool kDrop>
__global__ void __launch_bounds__(GemmLnCfg<BN>::THREADS, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmParams p, const LnEpiParams q) {
  using Cfg = GemmLnCfg<BN>;
  using T16 = typename Elem<kBF16>::T;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full_bar = empty_bar + Cfg::STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* epi_stage_base = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES);
  float* wstat = epi_stage_base + Cfg::EPI_WARPS * 32 * 33;      // [EPI_SPLIT][128][2]
  float* cstat = wstat + Cfg::EPI_SPLIT * BM * 2;                 // [128][2]  this CTA's (mean, M2) over BN columns
  float* fin = cstat + BM * 2;                                    // [128][2]  (mean, rstd) of the whole row

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int num_kb = (p.K + BK - 1) / BK;
  const int m0 = static_cast<int>(blockIdx.x / LN_CLUSTER) * BM;
  const int n0 = static_cast<int>(rank) * BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
        uint8_t* sB = sA + A_TILE_BYTES;
        mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
        tma_load_2d(sA, &tmA, &full_bar[stage], kb * BK, m0);
        tma_load_2d(sB, &tmB, &full_bar[stage], kb * BK, n0);
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(kBF16 ? 1 : 0, 0, 0, BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sA = smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const uint32_t sB = sA + A_TILE_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_ss(tmem_base, umma_smem_desc(sA + k * 32, 16, 1024), umma_smem_desc(sB + k * 32, 16, 1024),
                  idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tmem_full_bar);
    }
  }

  // ======================================================================= epilogue
  const bool is_epi = warp >= 4;
  const int quarter = warp & 3;                 // TMEM lanes [32*quarter, +32) == tile rows
  const int chalf = is_epi ? (warp - 4) >> 2 : 0;
  const int sub_r = lane >> 2;
  const int cg = (lane & 3) * 8;
  float* stage = epi_stage_base + (is_epi ? (warp - 4) : 0) * (32 * 33);
  const int row_base = m0 + quarter * 32;
  if (is_epi) {
    const DropoutRng rng = make_rng(p);
    float cnt[4], mean[4], M2[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) { cnt[it] = 0.f; mean[it] = 0.f; M2[it] = 0.f; }
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
    for (int cw = 0; cw < Cfg::CHUNKS; ++cw) {
      const int c = chalf * Cfg::CHUNKS + cw;
      uint32_t r[32];
      tmem_ld32(t_acc + c * 32, r);
      const int col = n0 + c * 32 + cg;
      // side inputs in the coalesced mapping, requested while the TMEM load is in flight
      const uint4 bias4 = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T16*>(p.bias) + col));
      uint4 side[4];
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int row = row_base + it * 8 + sub_r;
        side[it] = make_uint4(0, 0, 0, 0);
        if (row < p.M)
          side[it] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T16*>(p.residual) +
                                                          static_cast<long long>(row) * p.ldr + col));
      }
      tmem_ld_wait();
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 32; ++j) stage[lane * 33 + j] = __uint_as_float(r[j]);
      __syncwarp();
      float bias8[8];
      unpack8_<kBF16>(bias4, bias8);
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int rr = it * 8 + sub_r;
        const int row = row_base + rr;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = stage[rr * 33 + cg + i] + bias8[i];
        if (kDrop) {
          const uint64_t e = static_cast<uint64_t>(row) * static_cast<uint64_t>(p.N) + col;
          const uint4 rnd = rng.draw8(e >> 3);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = (rand16_of(rnd, i) < rng.thr16) ? 0.f : v[i] * rng.inv_keep;
        }
        float t[8];
        unpack8_<kBF16>(side[it], t);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += t[i];
        // the LayerNorm of the reference acts on the 16-bit sum: round first, then take statistics
        uint4 s16;
        s16.x = Elem<kBF16>::pack(v[0], v[1]); s16.y = Elem<kBF16>::pack(v[2], v[3]);
        s16.z = Elem<kBF16>::pack(v[4], v[5]); s16.w = Elem<kBF16>::pack(v[6], v[7]);
        if (row < p.M)
          *reinterpret_cast<uint4*>(reinterpret_cast<T16*>(p.out) + static_cast<long long>(row) * p.ldo + col) = s16;
        unpack8_<kBF16>(s16, v);
        float m8 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) m8 += v[i];
        m8 *= 0.125f;
        float q8 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = v[i] - m8; q8 = fmaf(d, d, q8); }
        if (cw == 0) { cnt[it] = 8.f; mean[it] = m8; M2[it] = q8; }
        else chan_merge(cnt[it], mean[it], M2[it], 8.f, m8, q8);
      }
    }
    // the 4 lanes of a row (equal counts)
#pragma unroll
    for (int it = 0; it < 4; ++it) {
#pragma unroll
      for (int off = 1; off <= 2; off <<= 1) {
        const float mb = __shfl_xor_sync(0xffffffffu, mean[it], off);
        const float qb = __shfl_xor_sync(0xffffffffu, M2[it], off);
        chan_merge(cnt[it], mean[it], M2[it], cnt[it], mb, qb);
      }
      if ((lane & 3) == 0) {
        float* w = wstat + (chalf * BM + quarter * 32 + it * 8 + sub_r) * 2;
        w[0] = mean[it];
        w[1] = M2[it];
      }
    }
    epi_bar_sync(Cfg::EPI_WARPS * 32);
    if (chalf == 0) {                 // one thread per tile row: merge the warps that share it
      const int rr = quarter * 32 + lane;
      float n = static_cast<float>(32 * Cfg::CHUNKS), m = wstat[rr * 2], s2 = wstat[rr * 2 + 1];
#pragma unroll
      for (int h = 1; h < Cfg::EPI_SPLIT; ++h)
        chan_merge(n, m, s2, static_cast<float>(32 * Cfg::CHUNKS), wstat[(h * BM + rr) * 2], wstat[(h * BM + rr) * 2 + 1]);
      cstat[rr * 2] = m;
      cstat[rr * 2 + 1] = s2;
    }
  }
  // ---- every CTA of the cluster has published its per-row partial statistics
  __syncwarp();
  cluster_sync_all();
  if (is_epi) {
    if (chalf == 0) {
      const int rr = quarter * 32 + lane;
      const uint32_t local = smem_u32(cstat + rr * 2);
      float n = 0.f, m = 0.f, s2 = 0.f;
#pragma unroll
      for (int r4 = 0; r4 < LN_CLUSTER; ++r4) {
        const float2 pr = ld_dsmem_f2(mapa_shared(local, static_cast<uint32_t>(r4)));
        if (r4 == 0) { n = static_cast<float>(BN); m = pr.x; s2 = pr.y; }
        else chan_merge(n, m, s2, static_cast<float>(BN), pr.x, pr.y);
      }
      fin[rr * 2] = m;
      fin[rr * 2 + 1] = rsqrtf(s2 * q.inv_n + LN_FUSED_EPS);
    }
    epi_bar_sync(Cfg::EPI_WARPS * 32);
    // ---- pass 2: y = (s - mean) * rstd * gamma + beta over this warp's columns
#pragma unroll 1
    for (int cw = 0; cw < Cfg::CHUNKS; ++cw) {
      const int c = chalf * Cfg::CHUNKS + cw;
      const int col = n0 + c * 32 + cg;
      float g8[8], b8[8];
      unpack8_<kBF16>(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T16*>(q.gamma) + col)), g8);
      unpack8_<kBF16>(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T16*>(q.beta) + col)), b8);
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int rr = quarter * 32 + it * 8 + sub_r;
        const int row = m0 + rr;
        if (row >= p.M) continue;
        // own store of pass 1 (same thread, same address): a plain (coherent) load
        const uint4 s16 = *reinterpret_cast<const uint4*>(reinterpret_cast<const T16*>(p.out) +
                                                          static_cast<long long>(row) * p.ldo + col);
        float v[8];
        unpack8_<kBF16>(s16, v);
        const float mu = fin[rr * 2], rs = fin[rr * 2 + 1];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (v[i] - mu) * rs * g8[i] + b8[i];
        store8<kBF16>(q.y, static_cast<long long>(row) * q.ldy + col, v);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // nobody exits while a peer may still read its statistics
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, bool kBF16, bool kDrop>
static int launch_gemm_ln_t(const GemmParams& p, const LnEpiParams& q, const CUtensorMap& tmA,
                            const CUtensorMap& tmB, cudaStream_t stream) {
  using Cfg = GemmLnCfg<BN>;
  auto kern = gemm_ln_kernel<BN, kBF16, kDrop>;
  static unsigned long long configured = 0;   // per instantiation, one bit per device
  if (first_use_on_device(configured))
    UB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  const int grid = p.tiles_m * LN_CLUSTER;
  ProfScope ps(stream);
  UB_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, stream, LN_CLUSTER, tmA,
                           tmB, p, q));
  return 0;
}

// N must be 768 (4 x 192) or 1024 (4 x 256); both operands K-major; epilogue = BIAS | RESIDUAL [| DROPOUT].
int launch_gemm_ln(int dtype, const GemmParams& p, const void* gamma, const void* beta, void* y,
                   long long ldy, const CUtensorMap& tmA, const CUtensorMap& tmB, cudaStream_t stream) {
  LnEpiParams q;
  q.gamma = gamma; q.beta = beta; q.y = y; q.ldy = ldy; q.inv_n = 1.0f / static_cast<float>(p.N);
  const bool drop = (p.epilogue & UB200_EPI_DROPOUT) != 0;
  const bool bf = dtype == UB200_BF16;
#define UB_LN_CASE(BNV)                                                                       \
  if (bf) return drop ? launch_gemm_ln_t<BNV, true, true>(p, q, tmA, tmB, stream)             \
                      : launch_gemm_ln_t<BNV, true, false>(p, q, tmA, tmB, stream);           \
  return drop ? launch_gemm_ln_t<BNV, false, true>(p, q, tmA, tmB, stream)                    \
              : launch_gemm_ln_t<BNV, false, false>(p, q, tmA, tmB, stream)
  if (p.N == 768) { UB_LN_CASE(192); }
  if (p.N == 1024) { UB_LN_CASE(256); }
#undef UB_LN_CASE
  return set_error(UB200_EUNSUPPORTED, "gemm+LayerNorm epilogue needs N = 768 or 1024 (got %d)", p.N);
}

}  // namespace ub
