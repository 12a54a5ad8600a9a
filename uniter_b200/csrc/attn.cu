// Fused variable-length multi-head self-attention for sm_100a (head_dim = 64).
//
// Replaces model/layer.py:80-100 of the reference (transpose_for_scores, QK^T, /sqrt(d), +mask,
// softmax, dropout, PV, permute+contiguous: ~10 launches over the padded [B,h,L,L] rectangle)
// with ONE kernel over the packed [T, 3H] QKV matrix:
//   * one CTA per (128-query tile, head, sequence); Q/K/V tiles are TMA'd straight out of the
//     packed QKV buffer (column offsets 0 / H / 2H select q / k / v, +64*head selects the head);
//   * S = Q K^T and O = P V run on tcgen05 (M=128, fp32 accumulators in TMEM);
//   * mask-by-omission: only the S_b valid keys of the sequence take part (the reference's
//     additive -10000 underflows to exactly 0 probability, so this is exact — SURVEY.md §8a E4);
//   * softmax in registers, one query row per thread (TMEM lane == row), exp2 with 1/sqrt(d)
//     folded in; probabilities are normalised and rounded to 16 bit BEFORE P.V, as the
//     reference's fp16 softmax output is;
//   * Philox dropout on P regenerated (not stored) by the backward kernel;
//   * ctx is written directly in [T, H] layout; the row-wise log-sum-exp is saved for backward.
//
// Backward (autograd mirror of the same lines) recomputes P from Q, K and the saved LSE:
//   dV = Pd^T dO,  dPd = dO V^T,  dS = P o (mask o dPd / keep - delta),  delta = rowsum(dO o O)
//   dQ = scale * dS K,  dK = scale * dS^T Q
// with all five contractions on tcgen05 from the same four TMA tiles (Q, K, V, dO); the
// transposed operands (P^T, dS^T, V as [keys x d], ...) are expressed through MN-major UMMA
// descriptors, nothing is transposed in memory.
#include "common.h"
#include "ptx.cuh"

namespace ub {

constexpr int ATT_D = 64;          // head dim (both UNITER configs)
constexpr int ATT_BM = 128;        // query rows per CTA == TMEM lanes
constexpr int ATT_BN = 128;        // keys per KV block
constexpr int ATT_TILE = ATT_BM * ATT_D * 2;   // 16 KB: one 128x64 16-bit tile
constexpr int ATT_MAXSEQ = 512;    // dropout index pitch == max_position_embeddings

struct AttnParams {
  const int* cu_seqlens;   // [B+1]
  int H, nheads, T;
  void* ctx;               // [T, H] 16-bit
  float* lse;              // [nheads, T]
  float scale;             // 1/sqrt(d)
  uint32_t drop_thr16;
  float drop_inv_keep;
  uint32_t seed_lo, seed_hi, stream_lo, stream_hi;
  // backward only
  const void* dctx;        // [T, H]
  void* dqkv;              // [T, 3H]
  float* dbias;            // [3H] fp32 column sums of dqkv (QKV bias gradient), accumulated; or NULL
  const unsigned long long* rng_dev;   // optional device-side dropout stream offset (graph replay)
};

// Column sums over the 32 rows (= lanes) of a warp of a 32-column register block: a butterfly in
// which every step halves the number of live columns per lane (16+8+4+2+1 = 31 shuffles); lane l
// ends up with the total of column l and adds it to dst[l].  All 32 lanes must call it.
__device__ __forceinline__ void warp_colsum32_atomic(const uint32_t (&r)[32], bool valid, float* dst,
                                                     int lane) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = valid ? __uint_as_float(r[i]) : 0.f;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  atomicAdd(dst + lane, v[0]);
}

// element index used to key the attention-probability dropout mask
__device__ __forceinline__ uint64_t attn_drop_group(int bh, int q, int key8) {
  return ((static_cast<uint64_t>(bh) * ATT_MAXSEQ + q) * ATT_MAXSEQ + key8) >> 3;
}

// write 8 consecutive 16-bit values (one 16-byte chunk) of row `r`, columns [col8, col8+8)
// into a K-major SWIZZLE_128B operand made of 64-column slabs of 128 rows x 128 B.
__device__ __forceinline__ void st_swz128(uint8_t* base, int r, int col8, uint4 v) {
  const int slab = col8 >> 6;
  const int chunk = (col8 & 63) >> 3;
  uint8_t* p = base + slab * ATT_TILE + r * 128 + ((chunk ^ (r & 7)) << 4);
  *reinterpret_cast<uint4*>(p) = v;
}

// kSingle: every sequence of the launch fits one 128-key block (max_seqlen <= 128: all of C2 / C4 / C5).
// The kernel is latency bound (TMA -> MMA -> TMEM -> exp -> smem -> MMA -> TMEM -> store, one chain
// per CTA), so what matters is how many CTAs an SM can hold.  With one key block Q and K are dead
// once S = Q K^T has completed and S is dead once P has been extracted, so P overwrites the Q|K
// tiles and O overwrites the S columns: 48 KB smem + 128 TMEM columns per CTA -> 4 CTAs / SM
// instead of 2 (80 KB, 256 columns).
template <bool kBF16, bool kSingle>
__global__ void __launch_bounds__(128, kSingle ? 4 : 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmQKV64,
                const AttnParams p) {
  pdl_launch_dependents();
  pdl_wait();   // cu_seqlens / qkv come from preceding kernels
  const int b = blockIdx.z, qt = blockIdx.x;
  const int seq0 = p.cu_seqlens[b];
  const int S = p.cu_seqlens[b + 1] - seq0;
  if (qt * ATT_BM >= S) return;  // whole CTA exits together, before any barrier / TMEM use
  // Head-pair packing: a sequence of <= 64 tokens fills only half of the 128-row MMA tile, so
  // one CTA carries TWO heads: rows 0-63 = head h0, rows 64-127 = head h0+1, keys likewise.
  // S = Q K^T over the 128 stacked keys is block diagonal in what matters; the off-diagonal
  // blocks of P are written as zeros so that O = P V stays one 128-key contraction.
  const bool pair = (S <= 64) && ((p.nheads & 1) == 0);
  if (pair && static_cast<int>(blockIdx.y) >= p.nheads / 2) return;
  const int h0 = pair ? 2 * blockIdx.y : blockIdx.y;
  const int nkv = (S + ATT_BN - 1) / ATT_BN;   // 1 in pair mode
  uint32_t rs0 = p.stream_lo, rs1 = p.stream_hi;
  if (p.drop_thr16) rng_add_dev_offset(p.rng_dev, rs0, rs1);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // kSingle requests no alignment slack (4 CTAs must fit an SM): the dynamic window starts 1024-aligned
  uint8_t* smem = smem_raw + (kSingle ? 0u : ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  if (kSingle && (smem_u32(smem_raw) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;
  uint8_t* sK = smem + ATT_TILE;
  uint8_t* sV = smem + 2 * ATT_TILE;
  uint8_t* sP = kSingle ? smem : smem + 3 * ATT_TILE;  // 2 slabs (kSingle: over the dead Q | K tiles)
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + (kSingle ? 3 : 5) * ATT_TILE);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmQKV64);
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, kSingle ? 128 : 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem;                        // 128 fp32 columns
  const uint32_t tO = kSingle ? tmem : tmem + 128;   // 64 fp32 columns (kSingle: over the dead S columns)
  const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;

  // this thread's query row: (head, position in the sequence, first key column of its block)
  const int half = pair ? (tid >> 6) : 0;
  const int head = h0 + half;
  const int qrow = pair ? (tid & 63) : qt * ATT_BM + tid;
  const int kcol0 = half * 64;
  const bool q_ok = qrow < S;
  const int bh = b * p.nheads + head;
  const float c = p.scale * 1.4426950408889634f;  // scale * log2(e)
  uint32_t ph_load = 0, ph_mma = 0;

  // ---------------------------------------------------------------- sweep 1: row max
  float m = -INFINITY, l = 0.f;
  for (int j = 0; j < nkv; ++j) {
    const int kv_len = pair ? S : min(ATT_BN, S - j * ATT_BN);
    const int n_pad = pair ? ATT_BN : ((kv_len + 15) & ~15);
    if (tid == 0) {
      if (pair) {
        mbar_expect_tx(bar_load, 3 * ATT_TILE);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          tma_load_2d(sQ + hh * (ATT_TILE / 2), &tmQKV64, bar_load, (h0 + hh) * ATT_D, seq0);
          tma_load_2d(sK + hh * (ATT_TILE / 2), &tmQKV64, bar_load, p.H + (h0 + hh) * ATT_D, seq0);
          tma_load_2d(sV + hh * (ATT_TILE / 2), &tmQKV64, bar_load, 2 * p.H + (h0 + hh) * ATT_D, seq0);
        }
      } else {
        const bool with_v = (nkv == 1);
        mbar_expect_tx(bar_load, (j == 0 ? ATT_TILE : 0) + ATT_TILE + (with_v ? ATT_TILE : 0));
        if (j == 0) tma_load_2d(sQ, &tmQKV, bar_load, head * ATT_D, seq0 + qt * ATT_BM);
        tma_load_2d(sK, &tmQKV, bar_load, p.H + head * ATT_D, seq0 + j * ATT_BN);
        if (with_v) tma_load_2d(sV, &tmQKV, bar_load, 2 * p.H + head * ATT_D, seq0 + j * ATT_BN);
      }
      mbar_wait(bar_load, ph_load);
      tc_fence_after();
      const uint32_t idesc = umma_idesc(kBF16 ? 1 : 0, 0, 0, ATT_BM, n_pad);
#pragma unroll
      for (int k = 0; k < ATT_D / 16; ++k)
        umma_ss(tS, umma_smem_desc(smem_u32(sQ) + k * 32, 16, 1024),
                umma_smem_desc(smem_u32(sK) + k * 32, 16, 1024), idesc, k != 0);
      umma_commit(bar_mma);
    }
    ph_load ^= 1;
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    tc_fence_after();
    for (int cc = 0; cc * 32 < kv_len; ++cc) {
      uint32_t r[32];
      tmem_ld32(tS + lane_off + kcol0 + cc * 32, r);
      tmem_ld_wait();
      if ((cc + 1) * 32 <= kv_len) {
#pragma unroll
        for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(r[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (cc * 32 + i < kv_len) m = fmaxf(m, __uint_as_float(r[i]));
      }
    }
    if (nkv > 1) {  // S (TMEM) and sK are about to be overwritten
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
  }
  const float mc = m * c;

  // ---------------------------------------------------------------- sweep 2: P and O = P V
  for (int j = 0; j < nkv; ++j) {
    const int kv_len = pair ? S : min(ATT_BN, S - j * ATT_BN);
    const int n_pad = pair ? ATT_BN : ((kv_len + 15) & ~15);
    if (nkv > 1) {
      if (tid == 0) {
        mbar_expect_tx(bar_load, 2 * ATT_TILE);
        tma_load_2d(sK, &tmQKV, bar_load, p.H + head * ATT_D, seq0 + j * ATT_BN);
        tma_load_2d(sV, &tmQKV, bar_load, 2 * p.H + head * ATT_D, seq0 + j * ATT_BN);
        mbar_wait(bar_load, ph_load);
        tc_fence_after();
        const uint32_t idesc = umma_idesc(kBF16 ? 1 : 0, 0, 0, ATT_BM, n_pad);
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k)
          umma_ss(tS, umma_smem_desc(smem_u32(sQ) + k * 32, 16, 1024),
                  umma_smem_desc(smem_u32(sK) + k * 32, 16, 1024), idesc, k != 0);
        umma_commit(bar_mma);
      }
      ph_load ^= 1;
      mbar_wait(bar_mma, ph_mma);
      ph_mma ^= 1;
      tc_fence_after();
    }
    // probabilities -> 16-bit -> swizzled smem (A operand of P.V)
    const int own_pad = pair ? 64 : n_pad;     // columns of this row's own block
    for (int cc = 0; cc * 32 < own_pad; ++cc) {
      uint32_t r[32];
      tmem_ld32(tS + lane_off + kcol0 + cc * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int key0 = cc * 32 + g * 8;  // within this row's key block
        if (key0 >= own_pad) break;
        float pv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float sc = __uint_as_float(r[g * 8 + i]);
          // unnormalised probability exp(scale*(s - max)) in (0, 1]; O is divided by the row sum
          // at the end (the 1/l factor commutes with dropout and with P.V)
          pv[i] = (key0 + i < kv_len) ? ex2_approx(fmaf(sc, c, -mc)) : 0.f;
          l += pv[i];
        }
        if (p.drop_thr16) {
          DropoutRng rng;
          rng.k0 = p.seed_lo; rng.k1 = p.seed_hi; rng.s0 = rs0; rng.s1 = rs1;
          const uint4 rnd = rng.draw8(attn_drop_group(bh, qrow, j * ATT_BN + key0));
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            // round P to 16 bit first (reference: softmax output is fp16, then dropout)
            const float pr = Elem<kBF16>::to_f(Elem<kBF16>::from_f(pv[i]));
            pv[i] = (rand16_of(rnd, i) < p.drop_thr16) ? 0.f : pr * p.drop_inv_keep;
          }
        }
        uint4 u;
        u.x = Elem<kBF16>::pack(pv[0], pv[1]);
        u.y = Elem<kBF16>::pack(pv[2], pv[3]);
        u.z = Elem<kBF16>::pack(pv[4], pv[5]);
        u.w = Elem<kBF16>::pack(pv[6], pv[7]);
        st_swz128(sP, tid, kcol0 + key0, u);
      }
    }
    if (pair) {   // the other head's 64 key columns of this row: exact zeros
#pragma unroll
      for (int g = 0; g < 8; ++g) st_swz128(sP, tid, (64 - kcol0) + g * 8, make_uint4(0, 0, 0, 0));
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t idesc = umma_idesc(kBF16 ? 1 : 0, 0, 1, ATT_BM, ATT_D);
      const int nk16 = n_pad >> 4;
      for (int kk = 0; kk < nk16; ++kk) {
        const uint32_t a = smem_u32(sP) + (kk >> 2) * ATT_TILE + (kk & 3) * 32;
        const uint32_t bb = smem_u32(sV) + kk * 2048;
        umma_ss(tO, umma_smem_desc(a, 16, 1024), umma_smem_desc(bb, 8192, 1024), idesc,
                (j | kk) != 0);
      }
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    tc_fence_after();
  }

  // ---------------------------------------------------------------- epilogue: O / l -> ctx[T, H]
  const float inv_l = 1.f / l;
  if (q_ok) p.lse[static_cast<size_t>(head) * p.T + seq0 + qrow] = m * p.scale + logf(l);
  {
    typename Elem<kBF16>::T* out = reinterpret_cast<typename Elem<kBF16>::T*>(p.ctx) +
                                   static_cast<size_t>(seq0 + qrow) * p.H + head * ATT_D;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      uint32_t r[32];
      tmem_ld32(tO + lane_off + cc * 32, r);
      tmem_ld_wait();
      if (q_ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = Elem<kBF16>::pack(__uint_as_float(r[g * 8 + 0]) * inv_l, __uint_as_float(r[g * 8 + 1]) * inv_l);
          u.y = Elem<kBF16>::pack(__uint_as_float(r[g * 8 + 2]) * inv_l, __uint_as_float(r[g * 8 + 3]) * inv_l);
          u.z = Elem<kBF16>::pack(__uint_as_float(r[g * 8 + 4]) * inv_l, __uint_as_float(r[g * 8 + 5]) * inv_l);
          u.w = Elem<kBF16>::pack(__uint_as_float(r[g * 8 + 6]) * inv_l, __uint_as_float(r[g * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(out + cc * 32 + g * 8) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, kSingle ? 128 : 256);
  }
}

constexpr int ATT_FWD_SMEM = 5 * ATT_TILE + 64 + 1024;
constexpr int ATT_FWD_SMEM_SINGLE = 3 * ATT_TILE + 64;

// =====================================================================================
// Backward.  One CTA per (128-key block j, head, sequence); loops over 128-query blocks i.
//   TMEM (512 cols): S[128] | dP[128] | dV[64] | dK[64] | dQ[64]
//   smem: Q_i, dO_i, K_j, V_j (TMA, 128B swizzle) + Pd and dS written by the softmax threads
//   (row = query, 64-key slabs) and consumed both K-major (dQ = dS K) and MN-major
//   (dV = Pd^T dO, dK = dS^T Q) by tcgen05.
// dQ of a sequence longer than one key block is accumulated with fp32 atomics in `dq_accum`.
// =====================================================================================
// kSingle (max_seqlen <= 128, one query block and one key block per sequence): V is dead after
// dP = dO V^T and S / dP are dead once P / dS have been extracted, so the first P slab overwrites the
// V tile and dV | dK | dQ overwrite the S | dP columns: 112 KB smem + 256 TMEM columns -> 2 CTAs / SM
// (128 registers x 256 threads x 2 = the whole register file) instead of 1.
template <bool kBF16, bool kSingle>
__global__ void __launch_bounds__(256, kSingle ? 2 : 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                const __grid_constant__ CUtensorMap tmQKV64, const __grid_constant__ CUtensorMap tmDO64,
                const AttnParams p, float* dq_accum) {
  using T16 = typename Elem<kBF16>::T;
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.z, j = blockIdx.x;
  const int seq0 = p.cu_seqlens[b];
  const int S = p.cu_seqlens[b + 1] - seq0;
  if (j * ATT_BN >= S) return;
  // head-pair packing for short sequences (see attn_fwd_kernel): rows / keys 0-63 belong to head
  // h0, 64-127 to head h0+1; P and dS are block diagonal with exact zeros off the diagonal.
  const bool pair = (S <= 64) && ((p.nheads & 1) == 0);
  if (pair && static_cast<int>(blockIdx.y) >= p.nheads / 2) return;
  const int h0 = pair ? 2 * blockIdx.y : blockIdx.y;
  const int nq = (S + ATT_BM - 1) / ATT_BM;
  const int nkv = (S + ATT_BN - 1) / ATT_BN;
  uint32_t rs0 = p.stream_lo, rs1 = p.stream_hi;
  if (p.drop_thr16) rng_add_dev_offset(p.rng_dev, rs0, rs1);
  const int kv_len = pair ? S : min(ATT_BN, S - j * ATT_BN);
  const int n_pad = pair ? ATT_BN : ((kv_len + 15) & ~15);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + (kSingle ? 0u : ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  if (kSingle && (smem_u32(smem_raw) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;
  uint8_t* sdO = smem + ATT_TILE;
  uint8_t* sK = smem + 2 * ATT_TILE;
  uint8_t* sV = smem + 3 * ATT_TILE;
  uint8_t* sP = smem + (kSingle ? 3 : 4) * ATT_TILE;   // 2 slabs (kSingle: slab 0 over the dead V tile)
  uint8_t* sDS = smem + (kSingle ? 5 : 6) * ATT_TILE;  // 2 slabs
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + (kSingle ? 7 : 8) * ATT_TILE);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);

  // 256 threads: thread pair (row, chalf) — both own TMEM lane `row`, each handles half of the
  // 32-column blocks (twice the warps per SM and half the serial softmax work per thread)
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row_t = tid & 127;
  const int chalf = tid >> 7;
  const int half = pair ? (row_t >> 6) : 0;
  const int head = h0 + half;
  const int kcol0 = half * 64;
  if (tid == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, kSingle ? 256 : 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tP = tmem + 128;
  const uint32_t tdV = tmem + (kSingle ? 0 : 256), tdK = tmem + (kSingle ? 64 : 320),
                 tdQ = tmem + (kSingle ? 128 : 384);
  const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const int bh = b * p.nheads + head;
  const float c = p.scale * 1.4426950408889634f;
  const uint32_t fmt = kBF16 ? 1 : 0;
  uint32_t ph_load = 0, ph_mma = 0;

  for (int i = 0; i < nq; ++i) {
    const int q_len = min(ATT_BM, S - i * ATT_BM);
    const int q_pad = pair ? ATT_BM : ((q_len + 15) & ~15);
    const int qrow = pair ? (row_t & 63) : i * ATT_BM + row_t;
    const bool q_ok = qrow < S;
    if (tid == 0) {
      if (pair) {
        mbar_expect_tx(bar_load, 4 * ATT_TILE);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int o = hh * (ATT_TILE / 2);
          tma_load_2d(sQ + o, &tmQKV64, bar_load, (h0 + hh) * ATT_D, seq0);
          tma_load_2d(sdO + o, &tmDO64, bar_load, (h0 + hh) * ATT_D, seq0);
          tma_load_2d(sK + o, &tmQKV64, bar_load, p.H + (h0 + hh) * ATT_D, seq0);
          tma_load_2d(sV + o, &tmQKV64, bar_load, 2 * p.H + (h0 + hh) * ATT_D, seq0);
        }
      } else {
        mbar_expect_tx(bar_load, (i == 0 ? 4 : 2) * ATT_TILE);
        tma_load_2d(sQ, &tmQKV, bar_load, head * ATT_D, seq0 + i * ATT_BM);
        tma_load_2d(sdO, &tmDO, bar_load, head * ATT_D, seq0 + i * ATT_BM);
        if (i == 0) {
          tma_load_2d(sK, &tmQKV, bar_load, p.H + head * ATT_D, seq0 + j * ATT_BN);
          tma_load_2d(sV, &tmQKV, bar_load, 2 * p.H + head * ATT_D, seq0 + j * ATT_BN);
        }
      }
      mbar_wait(bar_load, ph_load);
      tc_fence_after();
      const uint32_t idesc = umma_idesc(fmt, 0, 0, ATT_BM, n_pad);
#pragma unroll
      for (int k = 0; k < ATT_D / 16; ++k)   // S = Q K^T
        umma_ss(tS, umma_smem_desc(smem_u32(sQ) + k * 32, 16, 1024),
                umma_smem_desc(smem_u32(sK) + k * 32, 16, 1024), idesc, k != 0);
#pragma unroll
      for (int k = 0; k < ATT_D / 16; ++k)   // dPd = dO V^T
        umma_ss(tP, umma_smem_desc(smem_u32(sdO) + k * 32, 16, 1024),
                umma_smem_desc(smem_u32(sV) + k * 32, 16, 1024), idesc, k != 0);
      umma_commit(bar_mma);
    }
    ph_load ^= 1;

    // delta = rowsum(dO o O) and the saved log-sum-exp, straight from global (overlaps the MMAs)
    float delta = 0.f, lse2 = 0.f;
    if (q_ok) {
      const size_t off = static_cast<size_t>(seq0 + qrow) * p.H + head * ATT_D;
      const uint4* g_do = reinterpret_cast<const uint4*>(reinterpret_cast<const T16*>(p.dctx) + off);
      const uint4* g_o = reinterpret_cast<const uint4*>(reinterpret_cast<const T16*>(p.ctx) + off);
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        const uint4 a = __ldg(g_do + v), o = __ldg(g_o + v);
        float2 x, y;
        x = Elem<kBF16>::unpack(a.x); y = Elem<kBF16>::unpack(o.x); delta += x.x * y.x + x.y * y.y;
        x = Elem<kBF16>::unpack(a.y); y = Elem<kBF16>::unpack(o.y); delta += x.x * y.x + x.y * y.y;
        x = Elem<kBF16>::unpack(a.z); y = Elem<kBF16>::unpack(o.z); delta += x.x * y.x + x.y * y.y;
        x = Elem<kBF16>::unpack(a.w); y = Elem<kBF16>::unpack(o.w); delta += x.x * y.x + x.y * y.y;
      }
      lse2 = p.lse[static_cast<size_t>(head) * p.T + seq0 + qrow] * 1.4426950408889634f;
    }

    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    tc_fence_after();

    // 32-column blocks of this thread: pair mode -> one block of the row's own 64 columns (and
    // zero-fill of the matching block of the other head); else two of the up to four blocks.
    const int cc_begin = pair ? (half * 2 + chalf) : chalf * 2;
    const int cc_end = pair ? cc_begin + 1 : chalf * 2 + 2;
    for (int cc = cc_begin; cc < cc_end && cc * 32 < n_pad; ++cc) {
      uint32_t rs[32], rp[32];
      tmem_ld32(tS + lane_off + cc * 32, rs);
      tmem_ld32(tP + lane_off + cc * 32, rp);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int col0 = cc * 32 + g * 8;          // column in the 128-wide tile
        if (col0 >= n_pad) break;
        const int key0 = col0 - kcol0;             // key index inside this KV block
        float pd[8], ds[8];
        uint4 rnd = make_uint4(0, 0, 0, 0);
        if (p.drop_thr16) {
          DropoutRng rng;
          rng.k0 = p.seed_lo; rng.k1 = p.seed_hi; rng.s0 = rs0; rng.s1 = rs1;
          rnd = rng.draw8(attn_drop_group(bh, qrow, j * ATT_BN + key0));
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const bool ok = q_ok && (key0 + e < kv_len);
          float pr = ok ? ex2_approx(fmaf(__uint_as_float(rs[g * 8 + e]), c, -lse2)) : 0.f;
          pr = Elem<kBF16>::to_f(Elem<kBF16>::from_f(pr));   // P as the forward rounded it
          float dp = __uint_as_float(rp[g * 8 + e]);
          float pdv = pr;
          if (p.drop_thr16) {
            const bool drop = rand16_of(rnd, e) < p.drop_thr16;
            pdv = drop ? 0.f : pr * p.drop_inv_keep;
            dp = drop ? 0.f : dp * p.drop_inv_keep;
          }
          pd[e] = pdv;
          ds[e] = ok ? pr * (dp - delta) * p.scale : 0.f;
        }
        uint4 u;
        u.x = Elem<kBF16>::pack(pd[0], pd[1]); u.y = Elem<kBF16>::pack(pd[2], pd[3]);
        u.z = Elem<kBF16>::pack(pd[4], pd[5]); u.w = Elem<kBF16>::pack(pd[6], pd[7]);
        st_swz128(sP, row_t, col0, u);
        u.x = Elem<kBF16>::pack(ds[0], ds[1]); u.y = Elem<kBF16>::pack(ds[2], ds[3]);
        u.z = Elem<kBF16>::pack(ds[4], ds[5]); u.w = Elem<kBF16>::pack(ds[6], ds[7]);
        st_swz128(sDS, row_t, col0, u);
      }
    }
    if (pair) {   // exact zeros in the other head's key columns of this row
      const int zc = ((1 - half) * 2 + chalf) * 32;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        st_swz128(sP, row_t, zc + g * 8, make_uint4(0, 0, 0, 0));
        st_swz128(sDS, row_t, zc + g * 8, make_uint4(0, 0, 0, 0));
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      // dV += Pd^T dO ; dK += dS^T Q : A = [queries x keys] read MN-major (M = keys),
      // B = [queries x d] read MN-major (N = d); contraction over the q_pad query rows.
      const uint32_t idesc_t = umma_idesc(fmt, 1, 1, ATT_BM, ATT_D);
      const int nq16 = q_pad >> 4;
      for (int kk = 0; kk < nq16; ++kk) {
        umma_ss(tdV, umma_smem_desc(smem_u32(sP) + kk * 2048, ATT_TILE, 1024),
                umma_smem_desc(smem_u32(sdO) + kk * 2048, 8192, 1024), idesc_t, (i | kk) != 0);
      }
      for (int kk = 0; kk < nq16; ++kk) {
        umma_ss(tdK, umma_smem_desc(smem_u32(sDS) + kk * 2048, ATT_TILE, 1024),
                umma_smem_desc(smem_u32(sQ) + kk * 2048, 8192, 1024), idesc_t, (i | kk) != 0);
      }
      // dQ = dS K : A K-major over keys, B = K_j [keys x d] MN-major
      const uint32_t idesc_q = umma_idesc(fmt, 0, 1, ATT_BM, ATT_D);
      const int nk16 = n_pad >> 4;
      for (int kk = 0; kk < nk16; ++kk) {
        umma_ss(tdQ, umma_smem_desc(smem_u32(sDS) + (kk >> 2) * ATT_TILE + (kk & 3) * 32, 16, 1024),
                umma_smem_desc(smem_u32(sK) + kk * 2048, 8192, 1024), idesc_q, kk != 0);
      }
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    tc_fence_after();
    // dQ_i out (rows = queries); each thread of the pair writes one 32-column half
    {
      const int cc = chalf;
      uint32_t r[32];
      tmem_ld32(tdQ + lane_off + cc * 32, r);
      tmem_ld_wait();
      if (p.dbias)   // query-bias gradient: this key block's share of colsum(dQ) for `head`
        warp_colsum32_atomic(r, q_ok, p.dbias + head * ATT_D + cc * 32, tid & 31);
      if (q_ok) {
        if (nkv == 1) {
          T16* out = reinterpret_cast<T16*>(p.dqkv) + static_cast<size_t>(seq0 + qrow) * (3 * p.H) +
                     head * ATT_D + cc * 32;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 u;
            u.x = Elem<kBF16>::pack(__uint_as_float(r[g * 8 + 0]), __uint_as_float(r[g * 8 + 1]));
            u.y = Elem<kBF16>::pack(__uint_as_float(r[g * 8 + 2]), __uint_as_float(r[g * 8 + 3]));
            u.z = Elem<kBF16>::pack(__uint_as_float(r[g * 8 + 4]), __uint_as_float(r[g * 8 + 5]));
            u.w = Elem<kBF16>::pack(__uint_as_float(r[g * 8 + 6]), __uint_as_float(r[g * 8 + 7]));
            *reinterpret_cast<uint4*>(out + g * 8) = u;
          }
        } else {
          float* acc = dq_accum + static_cast<size_t>(seq0 + qrow) * p.H + head * ATT_D + cc * 32;
#pragma unroll
          for (int e = 0; e < 32; ++e) atomicAdd(acc + e, __uint_as_float(r[e]));
        }
      }
    }
    if (i + 1 < nq) {  // next iteration overwrites S / dP / dQ (TMEM) and sQ / sdO / sP / sDS
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
  }

  // dK_j, dV_j out (rows = keys)
  {
    const int key = pair ? (row_t & 63) : j * ATT_BN + row_t;
    const bool k_ok = key < S;
    T16* outk = reinterpret_cast<T16*>(p.dqkv) + static_cast<size_t>(seq0 + key) * (3 * p.H) +
                p.H + head * ATT_D;
    T16* outv = outk + p.H;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      {
        const int cc = chalf;
        uint32_t r[32];
        tmem_ld32((which ? tdV : tdK) + lane_off + cc * 32, r);
        tmem_ld_wait();
        if (p.dbias)   // key / value bias gradients
          warp_colsum32_atomic(r, k_ok, p.dbias + (which ? 2 : 1) * p.H + head * ATT_D + cc * 32, tid & 31);
        if (k_ok) {
          T16* out = (which ? outv : outk) + cc * 32;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 u;
            u.x = Elem<kBF16>::pack(__uint_as_float(r[g * 8 + 0]), __uint_as_float(r[g * 8 + 1]));
            u.y = Elem<kBF16>::pack(__uint_as_float(r[g * 8 + 2]), __uint_as_float(r[g * 8 + 3]));
            u.z = Elem<kBF16>::pack(__uint_as_float(r[g * 8 + 4]), __uint_as_float(r[g * 8 + 5]));
            u.w = Elem<kBF16>::pack(__uint_as_float(r[g * 8 + 6]), __uint_as_float(r[g * 8 + 7]));
            *reinterpret_cast<uint4*>(out + g * 8) = u;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, kSingle ? 256 : 512);
  }
}

// dqkv[:, 0:H] (16-bit) = dq_accum (fp32) for sequences spanning several key blocks (others
// wrote dQ directly).  One CTA per sequence.
template <bool kBF16>
__global__ void attn_dq_convert_kernel(const float* __restrict__ acc, void* dqkv,
                                       const int* __restrict__ cu_seqlens, int H) {
  pdl_launch_dependents();
  pdl_wait();
  const int seq0 = cu_seqlens[blockIdx.x];
  const int S = cu_seqlens[blockIdx.x + 1] - seq0;
  if (S <= ATT_BN) return;
  const int nvec = S * H / 8;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    const int row = (v * 8) / H, col = (v * 8) % H;
    const float* src = acc + static_cast<size_t>(seq0 + row) * H + col;
    const float4 a = *reinterpret_cast<const float4*>(src);
    const float4 b = *reinterpret_cast<const float4*>(src + 4);
    uint4 u;
    u.x = Elem<kBF16>::pack(a.x, a.y); u.y = Elem<kBF16>::pack(a.z, a.w);
    u.z = Elem<kBF16>::pack(b.x, b.y); u.w = Elem<kBF16>::pack(b.z, b.w);
    *reinterpret_cast<uint4*>(reinterpret_cast<typename Elem<kBF16>::T*>(dqkv) +
                              static_cast<size_t>(seq0 + row) * 3 * H + col) = u;
  }
}

constexpr int ATT_BWD_SMEM = 8 * ATT_TILE + 64 + 1024;
constexpr int ATT_BWD_SMEM_SINGLE = 7 * ATT_TILE + 64;

}  // namespace ub

extern "C" int ub200_attn_fwd(const ub200_attn_args* args, ub200_stream_t stream_) {
  using namespace ub;
  UB_CHECK_ARG(args != nullptr, "attn_fwd: args is NULL");
  const ub200_attn_args& a = *args;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  UB_CHECK_ARG(a.qkv && a.ctx && a.lse && a.cu_seqlens, "attn_fwd: null pointer");
  UB_CHECK_ARG(a.batch > 0 && a.total_tokens > 0, "attn_fwd: empty batch");
  UB_CHECK_ARG(a.num_heads > 0 && a.hidden == a.num_heads * ATT_D,
               "attn_fwd: head_dim must be 64 (hidden=%d heads=%d)", a.hidden, a.num_heads);
  UB_CHECK_ARG(a.max_seqlen > 0 && a.max_seqlen <= ATT_MAXSEQ,
               "attn_fwd: max_seqlen %d outside (0, %d]", a.max_seqlen, ATT_MAXSEQ);
  UB_CHECK_ARG(a.dtype == UB200_F16 || a.dtype == UB200_BF16, "attn_fwd: bad dtype");
  UB_CHECK_ARG(a.dropout_p >= 0.f && a.dropout_p < 1.f, "attn_fwd: dropout_p out of range");

  CUtensorMap tm, tm64;
  int rc = make_tma_2d(&tm, a.qkv, a.dtype, a.total_tokens, 3 * a.hidden, 3 * a.hidden, ATT_BM,
                       ATT_D);
  if (rc) return rc;
  rc = make_tma_2d(&tm64, a.qkv, a.dtype, a.total_tokens, 3 * a.hidden, 3 * a.hidden, 64, ATT_D);
  if (rc) return rc;
  AttnParams p{};
  p.cu_seqlens = a.cu_seqlens;
  p.H = a.hidden; p.nheads = a.num_heads; p.T = a.total_tokens;
  p.ctx = a.ctx; p.lse = a.lse;
  p.scale = 0.125f;
  if (a.dropout_p > 0.f) {
    uint32_t thr = static_cast<uint32_t>(a.dropout_p * 65536.0f + 0.5f);
    if (thr > 65535u) thr = 65535u;
    if (thr == 0u) thr = 1u;
    p.drop_thr16 = thr;
    p.drop_inv_keep = 65536.0f / static_cast<float>(65536u - thr);
  } else {
    p.drop_thr16 = 0; p.drop_inv_keep = 1.f;
  }
  p.seed_lo = static_cast<uint32_t>(a.rng_seed); p.seed_hi = static_cast<uint32_t>(a.rng_seed >> 32);
  p.stream_lo = static_cast<uint32_t>(a.rng_stream);
  p.stream_hi = static_cast<uint32_t>(a.rng_stream >> 32);
  p.rng_dev = reinterpret_cast<const unsigned long long*>(a.rng_offset_dev);

  dim3 grid((a.max_seqlen + ATT_BM - 1) / ATT_BM, a.num_heads, a.batch);
  const bool single = a.max_seqlen <= ATT_BN;
  const bool bf = a.dtype == UB200_BF16;
  void (*kern)(const CUtensorMap, const CUtensorMap, const AttnParams) =
      single ? (bf ? attn_fwd_kernel<true, true> : attn_fwd_kernel<false, true>)
             : (bf ? attn_fwd_kernel<true, false> : attn_fwd_kernel<false, false>);
  const int smem_bytes = single ? ATT_FWD_SMEM_SINGLE : ATT_FWD_SMEM;
  static unsigned long long configured[4] = {0, 0, 0, 0};   // one bit per device
  const int ci = (single ? 2 : 0) + (bf ? 1 : 0);
  if (first_use_on_device(configured[ci])) {
    UB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    if (single)   // 4 CTAs x 49 KB per SM: ask for the full shared-memory carve-out
      UB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         cudaSharedmemCarveoutMaxShared));
  }
  {
    ProfScope ps(stream);
    UB_CHECK_CUDA(launch_pdl(kern, grid, dim3(128), smem_bytes, stream, 1, tm, tm64, p));
  }
  UB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int64_t ub200_attn_bwd_workspace_bytes(int32_t total_tokens, int32_t hidden,
                                                  int32_t max_seqlen) {
  // fp32 dQ accumulator, only touched when some sequence is longer than one 128-key block
  if (max_seqlen <= ub::ATT_BN) return 0;
  return static_cast<int64_t>(total_tokens) * hidden * 4;
}

extern "C" int ub200_attn_bwd(const ub200_attn_args* args, ub200_stream_t stream_) {
  using namespace ub;
  UB_CHECK_ARG(args != nullptr, "attn_bwd: args is NULL");
  const ub200_attn_args& a = *args;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  UB_CHECK_ARG(a.qkv && a.ctx && a.lse && a.cu_seqlens && a.dctx && a.dqkv, "attn_bwd: null pointer");
  UB_CHECK_ARG(a.batch > 0 && a.total_tokens > 0, "attn_bwd: empty batch");
  UB_CHECK_ARG(a.num_heads > 0 && a.hidden == a.num_heads * ATT_D, "attn_bwd: head_dim must be 64");
  UB_CHECK_ARG(a.max_seqlen > 0 && a.max_seqlen <= ATT_MAXSEQ, "attn_bwd: max_seqlen %d outside (0, %d]",
               a.max_seqlen, ATT_MAXSEQ);
  UB_CHECK_ARG(a.dtype == UB200_F16 || a.dtype == UB200_BF16, "attn_bwd: bad dtype");
  const bool multi = a.max_seqlen > ATT_BN;
  UB_CHECK_ARG(!multi || a.workspace, "attn_bwd: max_seqlen > 128 needs the dQ workspace");

  CUtensorMap tmQ, tmD, tmQ64, tmD64;
  int rc = make_tma_2d(&tmQ, a.qkv, a.dtype, a.total_tokens, 3 * a.hidden, 3 * a.hidden, ATT_BM, ATT_D);
  if (rc) return rc;
  rc = make_tma_2d(&tmD, a.dctx, a.dtype, a.total_tokens, a.hidden, a.hidden, ATT_BM, ATT_D);
  if (rc) return rc;
  rc = make_tma_2d(&tmQ64, a.qkv, a.dtype, a.total_tokens, 3 * a.hidden, 3 * a.hidden, 64, ATT_D);
  if (rc) return rc;
  rc = make_tma_2d(&tmD64, a.dctx, a.dtype, a.total_tokens, a.hidden, a.hidden, 64, ATT_D);
  if (rc) return rc;
  AttnParams p{};
  p.cu_seqlens = a.cu_seqlens;
  p.H = a.hidden; p.nheads = a.num_heads; p.T = a.total_tokens;
  p.ctx = a.ctx; p.lse = a.lse; p.dctx = a.dctx; p.dqkv = a.dqkv; p.dbias = a.dbias;
  p.scale = 0.125f;
  if (a.dropout_p > 0.f) {
    uint32_t thr = static_cast<uint32_t>(a.dropout_p * 65536.0f + 0.5f);
    if (thr > 65535u) thr = 65535u;
    if (thr == 0u) thr = 1u;
    p.drop_thr16 = thr;
    p.drop_inv_keep = 65536.0f / static_cast<float>(65536u - thr);
  } else {
    p.drop_thr16 = 0; p.drop_inv_keep = 1.f;
  }
  p.seed_lo = static_cast<uint32_t>(a.rng_seed); p.seed_hi = static_cast<uint32_t>(a.rng_seed >> 32);
  p.stream_lo = static_cast<uint32_t>(a.rng_stream);
  p.stream_hi = static_cast<uint32_t>(a.rng_stream >> 32);
  p.rng_dev = reinterpret_cast<const unsigned long long*>(a.rng_offset_dev);
  float* acc = reinterpret_cast<float*>(a.workspace);
  if (multi)
    UB_CHECK_CUDA(cudaMemsetAsync(acc, 0, static_cast<size_t>(a.total_tokens) * a.hidden * 4, stream));

  dim3 grid((a.max_seqlen + ATT_BN - 1) / ATT_BN, a.num_heads, a.batch);
  const bool single = !multi;
  const int di = a.dtype == UB200_BF16 ? 1 : 0;
  void (*kern)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const AttnParams,
               float*) =
      single ? (di ? attn_bwd_kernel<true, true> : attn_bwd_kernel<false, true>)
             : (di ? attn_bwd_kernel<true, false> : attn_bwd_kernel<false, false>);
  const int smem_bytes = single ? ATT_BWD_SMEM_SINGLE : ATT_BWD_SMEM;
  static unsigned long long configured[4] = {0, 0, 0, 0};   // one bit per device
  const int ci = (single ? 2 : 0) + di;
  if (first_use_on_device(configured[ci])) {
    UB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    if (single)
      UB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         cudaSharedmemCarveoutMaxShared));
  }
  {
    ProfScope ps(stream);
    UB_CHECK_CUDA(launch_pdl(kern, grid, dim3(256), smem_bytes, stream, 1, tmQ, tmD, tmQ64, tmD64, p, acc));
  }
  UB_CHECK_CUDA(cudaGetLastError());
  if (multi) {
    ProfScope ps(stream);
    if (di) UB_CHECK_CUDA(launch_pdl(attn_dq_convert_kernel<true>, dim3(a.batch), dim3(256), 0, stream, 1, static_cast<const float*>(acc), a.dqkv, a.cu_seqlens, a.hidden));
    else UB_CHECK_CUDA(launch_pdl(attn_dq_convert_kernel<false>, dim3(a.batch), dim3(256), 0, stream, 1, static_cast<const float*>(acc), a.dqkv, a.cu_seqlens, a.hidden));
    UB_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}
