// tcgen05 / TMA GEMM core for sm_100a.
//
//   D[M,N] = epilogue( sum_k A[m,k] * B[n,k] ),  16-bit operands, fp32 accumulation in TMEM.
//
// One persistent CTA per SM, 8 warps with fixed roles:
//   warp 0     TMA producer   — one lane streams 128xBK A tiles and BNxBK B tiles into a
//                               128B-swizzled smem ring (mbarrier full/empty pairs)
//   warp 1     MMA issuer     — one lane issues tcgen05.mma (M=128, N=BN, K=16) x4 per stage,
//                               tcgen05.commit releases smem slots and publishes accumulators
//   warp 2     TMEM allocator — 2 x BN fp32 columns (double-buffered accumulator)
//   warps 4-7  epilogue       — tcgen05.ld 32 lanes x 32 columns, bias / dropout / residual /
//                               GELU / dGELU / accumulate / column-sum, 16-byte global stores
// The mainloop of tile i+1 overlaps the epilogue of tile i through the two TMEM buffers.
//
// Operands may be K-major (contraction dim contiguous; nn.Linear forward) or MN-major
// (contraction dim strided; dgrad reads the weight un-transposed, wgrad reads both activation
// matrices un-transposed) — the UMMA descriptors encode the difference, no transposes are
// ever materialised.  Reference call sites replaced: model/layer.py:76-78,112,140,153 and
// their autograd mirrors.
#include "common.h"
#include "ptx.cuh"

namespace ub {

constexpr int BM = 128;
constexpr int BK = 64;               // 64 x 16-bit = one 128-byte swizzle row
constexpr int A_TILE_BYTES = BM * BK * 2;
constexpr int GEMM_THREADS = 256;

template <int BN>
struct GemmCfg {
  static constexpr int B_TILE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // + align slack
};

struct GemmParams {
  int M, N, K;
  int epilogue;
  const void* bias;
  const void* residual;
  const void* aux;
  void* out;
  void* out2;
  float* colsum;
  long long ldr, ldaux, ldo;
  uint32_t drop_thr16;
  float drop_inv_keep;
  uint32_t seed_lo, seed_hi, stream_lo, stream_hi;
  int tiles_m, tiles_n;
};

template <bool kBF16>
__device__ __forceinline__ void load8(const void* base, long long idx, float (&f)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(
      reinterpret_cast<const typename Elem<kBF16>::T*>(base) + idx));
  float2 t;
  t = Elem<kBF16>::unpack(u.x); f[0] = t.x; f[1] = t.y;
  t = Elem<kBF16>::unpack(u.y); f[2] = t.x; f[3] = t.y;
  t = Elem<kBF16>::unpack(u.z); f[4] = t.x; f[5] = t.y;
  t = Elem<kBF16>::unpack(u.w); f[6] = t.x; f[7] = t.y;
}
template <bool kBF16>
__device__ __forceinline__ void store8(void* base, long long idx, const float (&f)[8]) {
  uint4 u;
  u.x = Elem<kBF16>::pack(f[0], f[1]);
  u.y = Elem<kBF16>::pack(f[2], f[3]);
  u.z = Elem<kBF16>::pack(f[4], f[5]);
  u.w = Elem<kBF16>::pack(f[6], f[7]);
  *reinterpret_cast<uint4*>(reinterpret_cast<typename Elem<kBF16>::T*>(base) + idx) = u;
}

template <int BN, bool A_MN, bool B_MN, bool kBF16>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full_bar = empty_bar + Cfg::STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;        // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = (p.K + BK - 1) / BK;
  const int num_tiles = p.tiles_m * p.tiles_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.tiles_n) * BM;
        const int n0 = (tile % p.tiles_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sB = sA + A_TILE_BYTES;
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          if (!A_MN) {
            tma_load_2d(sA, &tmA, &full_bar[stage], kb * BK, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_2d(sA + j * (64 * BK * 2), &tmA, &full_bar[stage], m0 + j * 64, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d(sB, &tmB, &full_bar[stage], kb * BK, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sB + j * (64 * BK * 2), &tmB, &full_bar[stage], n0 + j * 64, kb * BK);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(kBF16 ? 1 : 0, A_MN ? 1 : 0, B_MN ? 1 : 0, BM, BN);
      // K-major: advance 16 elements = 32 B inside the swizzle row; 8-row groups 1024 B apart.
      // MN-major: advance 16 K-rows = 2048 B; 64-wide M/N groups one 8 KB box apart.
      constexpr uint32_t A_KSTEP = A_MN ? 2048 : 32, A_LBO = A_MN ? 8192 : 16;
      constexpr uint32_t B_KSTEP = B_MN ? 2048 : 32, B_LBO = B_MN ? 8192 : 16;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sB = sA + A_TILE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = umma_smem_desc(sA + k * A_KSTEP, A_LBO, 1024);
            const uint64_t db = umma_smem_desc(sB + k * B_KSTEP, B_LBO, 1024);
            umma_ss(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // smem slot free once these MMAs retire
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================================================================== epilogue
    const int quarter = warp - 4;  // == warp % 4 -> TMEM lanes [32*quarter, 32*quarter+32)
    const int epi = p.epilogue;
    DropoutRng rng;
    rng.k0 = p.seed_lo; rng.k1 = p.seed_hi; rng.s0 = p.stream_lo; rng.s1 = p.stream_hi;
    rng.thr16 = p.drop_thr16; rng.inv_keep = p.drop_inv_keep;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / p.tiles_n) * BM;
      const int n0 = (tile % p.tiles_n) * BN;
      const int row = m0 + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + static_cast<uint32_t>(acc * BN) +
                             (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(t_acc + c * 32, r);
        tmem_ld_wait();
        if (c == BN / 32 - 1) {
          // all of this warp's TMEM reads for the tile are done: hand the buffer back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
        const int col0 = n0 + c * 32;
        if (col0 >= p.N) continue;  // warp-uniform
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);

#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int col = col0 + g * 8;
          if (col >= p.N) {  // N % 8 == 0 is enforced on the host
#pragma unroll
            for (int i = 0; i < 8; ++i) v[g * 8 + i] = 0.f;
            continue;
          }
          float t[8];
          if (epi & UB200_EPI_BIAS) {
            load8<kBF16>(p.bias, col, t);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[g * 8 + i] += t[i];
          }
          if (epi & UB200_EPI_DROPOUT) {
            const uint64_t e = static_cast<uint64_t>(row) * static_cast<uint64_t>(p.N) + col;
            const uint4 rnd = rng.draw8(e >> 3);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              v[g * 8 + i] = (rand16_of(rnd, i) < rng.thr16) ? 0.f : v[g * 8 + i] * rng.inv_keep;
          }
          if (row_ok) {
            if (epi & UB200_EPI_RESIDUAL) {
              load8<kBF16>(p.residual, static_cast<long long>(row) * p.ldr + col, t);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[g * 8 + i] += t[i];
            }
            if (epi & UB200_EPI_GELU) {
              float pre[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                // the reference rounds the Linear output to 16 bits before GELU
                pre[i] = Elem<kBF16>::to_f(Elem<kBF16>::from_f(v[g * 8 + i]));
                v[g * 8 + i] = gelu_erf(pre[i]);
              }
              store8<kBF16>(p.out2, static_cast<long long>(row) * p.ldo + col, pre);
            }
            if (epi & UB200_EPI_DGELU) {
              load8<kBF16>(p.aux, static_cast<long long>(row) * p.ldaux + col, t);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[g * 8 + i] *= dgelu_erf(t[i]);
            }
            if (epi & UB200_EPI_OUT_F32) {
              float* o = reinterpret_cast<float*>(p.out) + static_cast<long long>(row) * p.ldo + col;
              if (epi & UB200_EPI_ACCUM) {
                const float4 o0 = *reinterpret_cast<const float4*>(o);
                const float4 o1 = *reinterpret_cast<const float4*>(o + 4);
                v[g * 8 + 0] += o0.x; v[g * 8 + 1] += o0.y; v[g * 8 + 2] += o0.z; v[g * 8 + 3] += o0.w;
                v[g * 8 + 4] += o1.x; v[g * 8 + 5] += o1.y; v[g * 8 + 6] += o1.z; v[g * 8 + 7] += o1.w;
              }
              *reinterpret_cast<float4*>(o) =
                  make_float4(v[g * 8 + 0], v[g * 8 + 1], v[g * 8 + 2], v[g * 8 + 3]);
              *reinterpret_cast<float4*>(o + 4) =
                  make_float4(v[g * 8 + 4], v[g * 8 + 5], v[g * 8 + 6], v[g * 8 + 7]);
            } else {
              if (epi & UB200_EPI_ACCUM) {
                load8<kBF16>(p.out, static_cast<long long>(row) * p.ldo + col, t);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[g * 8 + i] += t[i];
              }
              float o8[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) o8[i] = v[g * 8 + i];
              store8<kBF16>(p.out, static_cast<long long>(row) * p.ldo + col, o8);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[g * 8 + i] = 0.f;
          }
        }
        if (epi & UB200_EPI_COLSUM) {
          // transpose-reduce: after 5 exchange steps lane L holds sum over the warp's 32 rows
          // of column (col0 + L).  31 shuffles per 32x32 block.
#pragma unroll
          for (int s = 16; s >= 1; s >>= 1) {
#pragma unroll
            for (int i = 0; i < s; ++i) {
              const bool up = (lane & s) != 0;
              const float send = up ? v[i] : v[i + s];
              const float keep = up ? v[i + s] : v[i];
              v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
            }
          }
          if (col0 + lane < p.N) atomicAdd(p.colsum + col0 + lane, v[0]);
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------- host side
template <int BN, bool A_MN, bool B_MN, bool kBF16>
static int launch_gemm(const ub200_gemm_args& a, const GemmParams& p, const CUtensorMap& tmA,
                       const CUtensorMap& tmB, int grid, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_kernel<BN, A_MN, B_MN, kBF16>;
  static bool configured = false;  // per instantiation
  if (!configured) {
    UB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::SMEM_BYTES));
    configured = true;
  }
  {
    ProfScope ps(stream);
    kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
  }
  UB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int BN, bool kBF16>
static int dispatch_major(const ub200_gemm_args& a, const GemmParams& p, const CUtensorMap& tmA,
                          const CUtensorMap& tmB, int grid, cudaStream_t stream) {
  if (a.a_major == 0 && a.b_major == 0)
    return launch_gemm<BN, false, false, kBF16>(a, p, tmA, tmB, grid, stream);
  if (a.a_major == 0 && a.b_major == 1)
    return launch_gemm<BN, false, true, kBF16>(a, p, tmA, tmB, grid, stream);
  if (a.a_major == 1 && a.b_major == 1)
    return launch_gemm<BN, true, true, kBF16>(a, p, tmA, tmB, grid, stream);
  return set_error(UB200_EUNSUPPORTED, "gemm: a_major=1 with b_major=0 is not instantiated");
}

template <bool kBF16>
static int dispatch_bn(int bn, const ub200_gemm_args& a, const GemmParams& p,
                       const CUtensorMap& tmA, const CUtensorMap& tmB, int grid,
                       cudaStream_t stream) {
  switch (bn) {
    case 64: return dispatch_major<64, kBF16>(a, p, tmA, tmB, grid, stream);
    case 128: return dispatch_major<128, kBF16>(a, p, tmA, tmB, grid, stream);
    case 256: return dispatch_major<256, kBF16>(a, p, tmA, tmB, grid, stream);
  }
  return set_error(UB200_EINVAL, "gemm: tile_n must be 0, 64, 128 or 256 (got %d)", bn);
}

// Pick the N tile that minimises (waves x per-tile cost) on `sms` SMs.
static int pick_tile_n(int M, int N, int sms) {
  const int tiles_m = (M + BM - 1) / BM;
  int best = 128;
  double best_cost = 1e30;
  const int cands[3] = {256, 128, 64};
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    const int tiles = tiles_m * ((N + bn - 1) / bn);
    const int waves = (tiles + sms - 1) / sms;
    // per-tile time ~ BN (MMA) with a fixed overhead; narrow tiles are smem-bandwidth bound
    const double per_tile = (bn == 64 ? 80.0 : (double)bn) + 24.0;
    const double cost = waves * per_tile;
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  return best;
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
  UB_CHECK_ARG(!(epi & UB200_EPI_COLSUM) || a.colsum, "gemm: EPI_COLSUM without colsum");
  UB_CHECK_ARG(!((epi & UB200_EPI_GELU) && (epi & UB200_EPI_OUT_F32)),
               "gemm: EPI_GELU with fp32 output is not supported");
  UB_CHECK_ARG(a.dropout_p >= 0.f && a.dropout_p < 1.f, "gemm: dropout_p out of range");

  const int sms = num_sms();
  const int bn = a.tile_n ? a.tile_n : pick_tile_n(a.M, a.N, sms);

  CUtensorMap tmA, tmB;
  int rc;
  if (a.a_major == 0)
    rc = make_tma_2d(&tmA, a.a, a.dtype, a.M, a.K, a.lda, BM, BK);
  else
    rc = make_tma_2d(&tmA, a.a, a.dtype, a.K, a.M, a.lda, BK, 64);
  if (rc) return rc;
  if (a.b_major == 0)
    rc = make_tma_2d(&tmB, a.b, a.dtype, a.N, a.K, a.ldb, bn, BK);
  else
    rc = make_tma_2d(&tmB, a.b, a.dtype, a.K, a.N, a.ldb, BK, 64);
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
  p.tiles_m = (a.M + BM - 1) / BM;
  p.tiles_n = (a.N + bn - 1) / bn;
  const int tiles = p.tiles_m * p.tiles_n;
  int grid = a.max_ctas > 0 ? a.max_ctas : sms;
  if (grid > tiles) grid = tiles;

  if (a.dtype == UB200_BF16) return dispatch_bn<true>(bn, a, p, tmA, tmB, grid, stream);
  return dispatch_bn<false>(bn, a, p, tmA, tmB, grid, stream);
}
