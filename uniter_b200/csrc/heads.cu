// Callers either side of the encoder stack:
//
//   ce_fwd / ce_bwd   fused softmax cross-entropy over the tied MLM decoder's [n, V] scores
//                     (model/pretrain.py:122-127: F.cross_entropy(prediction_scores, labels,
//                     reduction='none')): per-row log-sum-exp + loss in fp32 from the 16-bit
//                     scores; backward writes (softmax - onehot) * dloss IN PLACE over the scores,
//                     zeroing the padding columns [V, ld) that keep the row pitch a multiple of 8.
//   dgelu_mul         dpre = dy * gelu_erf'(pre): BertPredictionHeadTransform backward
//                     (model/layer.py:188-203) between its LayerNorm backward and dense dgrad.
//   sumsq / adamw     multi-tensor gradient norm and AdamW step with the reference's exact update
//                     (optim/adamw.py:77-101: bias-corrected step size, decoupled decay applied
//                     AFTER the Adam update) on fp32 master weights, fused with gradient
//                     unscaling, global-norm clipping (train_vqa.py:223-226) and the 16-bit
//                     model-weight refresh that apex O2 does as separate passes.
#include "common.h"
#include "ptx.cuh"

namespace ub {

template <bool kBF16>
__device__ __forceinline__ void h_unpack8(const uint4& u, float* f) {
  float2 t;
  t = Elem<kBF16>::unpack(u.x); f[0] = t.x; f[1] = t.y;
  t = Elem<kBF16>::unpack(u.y); f[2] = t.x; f[3] = t.y;
  t = Elem<kBF16>::unpack(u.z); f[4] = t.x; f[5] = t.y;
  t = Elem<kBF16>::unpack(u.w); f[6] = t.x; f[7] = t.y;
}
template <bool kBF16>
__device__ __forceinline__ uint4 h_pack8(const float* f) {
  uint4 u;
  u.x = Elem<kBF16>::pack(f[0], f[1]); u.y = Elem<kBF16>::pack(f[2], f[3]);
  u.z = Elem<kBF16>::pack(f[4], f[5]); u.w = Elem<kBF16>::pack(f[6], f[7]);
  return u;
}

// CTA-wide reductions (256 threads), result broadcast to every thread
__device__ __forceinline__ float block_max(float v, float* red) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) r = fmaxf(r, red[w]);
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) r += red[w];
  return r;
}

// ------------------------------------------------------------------------------ cross-entropy fwd
// One CTA per row.  The row (V x 2 bytes, 58 KB for the BERT vocabulary) is read twice (max, then
// sum of exp) and stays in L1 / L2 between the passes.
template <bool kBF16>
__global__ void __launch_bounds__(256)
ce_fwd_kernel(const void* __restrict__ logits_, long long ld, const long long* __restrict__ targets,
              float* __restrict__ loss, float* __restrict__ lse_out, int V) {
  pdl_launch_dependents();
  pdl_wait();
  using T16 = typename Elem<kBF16>::T;
  __shared__ float red[8];
  const int row = blockIdx.x;
  const T16* x = reinterpret_cast<const T16*>(logits_) + static_cast<long long>(row) * ld;
  const uint4* xv = reinterpret_cast<const uint4*>(x);
  const int nvec = (V + 7) >> 3;
  float m = -INFINITY;
  for (int v = threadIdx.x; v < nvec; v += 256) {
    float f[8];
    h_unpack8<kBF16>(__ldg(xv + v), f);
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (v * 8 + e < V) m = fmaxf(m, f[e]);
  }
  m = block_max(m, red);
  float s = 0.f;
  for (int v = threadIdx.x; v < nvec; v += 256) {
    float f[8];
    h_unpack8<kBF16>(__ldg(xv + v), f);
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (v * 8 + e < V) s += __expf(f[e] - m);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    const float lse = m + logf(s);
    lse_out[row] = lse;
    const long long t = targets[row];
    loss[row] = (t >= 0 && t < V) ? lse - Elem<kBF16>::to_f(x[t]) : 0.f;
  }
}

// ------------------------------------------------------------------------------ cross-entropy bwd
// d[r, c] = (exp(x[r, c] - lse[r]) - [c == target[r]]) * dloss[r]; columns [V, ncols) -> 0.
// `out` may alias `logits` (each 16-byte vector is read and written by the same thread).
template <bool kBF16>
__global__ void __launch_bounds__(256)
ce_bwd_kernel(const void* logits_, void* out_, long long ld, const long long* __restrict__ targets,
              const float* __restrict__ lse, const float* __restrict__ dloss, int V, int ncols) {
  pdl_launch_dependents();
  pdl_wait();
  using T16 = typename Elem<kBF16>::T;
  const int row = blockIdx.x;
  const uint4* xv = reinterpret_cast<const uint4*>(reinterpret_cast<const T16*>(logits_) +
                                                   static_cast<long long>(row) * ld);
  uint4* ov = reinterpret_cast<uint4*>(reinterpret_cast<T16*>(out_) + static_cast<long long>(row) * ld);
  const long long t = targets[row];
  const bool live = t >= 0 && t < V;
  const float g = live ? dloss[row] : 0.f;
  const float l = lse[row];
  const int nvec = ncols >> 3;
  for (int v = threadIdx.x; v < nvec; v += 256) {
    float f[8];
    h_unpack8<kBF16>(xv[v], f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = v * 8 + e;
      float d = 0.f;
      if (live && c < V) d = (__expf(f[e] - l) - (c == t ? 1.f : 0.f)) * g;
      f[e] = d;
    }
    ov[v] = h_pack8<kBF16>(f);
  }
}

// ------------------------------------------------------------------------------ dy * gelu'(pre)
// kTanh: out = dy * (1 - y^2) with y = tanh(pre) saved by the forward (BertPooler backward)
template <bool kBF16, bool kTanh = false>
__global__ void __launch_bounds__(256)
dgelu_mul_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ pre, uint4* __restrict__ out,
                 long long nvec) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float a[8], b[8];
    h_unpack8<kBF16>(__ldg(dy + i), a);
    h_unpack8<kBF16>(__ldg(pre + i), b);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] *= kTanh ? (1.0f - b[e] * b[e]) : dgelu_erf(b[e]);
    out[i] = h_pack8<kBF16>(a);
  }
}

// ------------------------------------------------------------------------------ multi-tensor optimizer
// A segment is one parameter tensor.  CTA `b` owns elements [ (b - blk_start[s]) * ADAM_CHUNK, +ADAM_CHUNK )
// of the segment s with blk_start[s] <= b < blk_start[s + 1] (binary search over <= a few hundred
// segments).
constexpr int ADAM_CHUNK = 4096;   // elements per CTA: 256 threads x 16

__device__ __forceinline__ int find_segment(const int* __restrict__ blk_start, int nseg, int b) {
  int lo = 0, hi = nseg - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(blk_start + mid) <= b) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__device__ __forceinline__ float grad_of(const ub200_adam_segment& sg, long long i) {
  if (sg.grad_dtype == UB200_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(sg.grad)[i]);
  if (sg.grad_dtype == UB200_F16) return __half2float(reinterpret_cast<const __half*>(sg.grad)[i]);
  return reinterpret_cast<const float*>(sg.grad)[i];
}

// sum of squares of all gradients (before unscaling) -> out[0] (fp32, atomically accumulated)
__global__ void __launch_bounds__(256)
sumsq_kernel(const ub200_adam_segment* __restrict__ segs, const int* __restrict__ blk_start, int nseg,
             float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8];
  const int s = find_segment(blk_start, nseg, blockIdx.x);
  const ub200_adam_segment sg = segs[s];
  const long long base = static_cast<long long>(blockIdx.x - blk_start[s]) * ADAM_CHUNK;
  float acc = 0.f;
  if (sg.grad_dtype != UB200_F32 && sg.n % 8 == 0 && (reinterpret_cast<uintptr_t>(sg.grad) & 15) == 0) {
    const bool bf = sg.grad_dtype == UB200_BF16;
    for (int j = threadIdx.x * 8; j < ADAM_CHUNK; j += 2048) {
      const long long i = base + j;
      if (i >= sg.n) break;
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(sg.grad) + i));
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = bf ? Elem<true>::unpack(w[q]) : Elem<false>::unpack(w[q]);
        acc = fmaf(f.x, f.x, acc);
        acc = fmaf(f.y, f.y, acc);
      }
    }
  } else {
    for (int j = threadIdx.x; j < ADAM_CHUNK; j += 256) {
      const long long i = base + j;
      if (i < sg.n) { const float g = grad_of(sg, i); acc = fmaf(g, g, acc); }
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0 && acc != 0.f) atomicAdd(out, acc);
}

struct AdamHyper {
  float beta1, beta2, eps;
  float inv_scale;        // 1 / loss_scale (gradient unscaling)
  float max_norm;         // <= 0: no clipping
  const float* sumsq;     // device scalar from sumsq_kernel (of the SCALED gradients) or NULL
  const ub200_adam_state* state;   // device-resident step counter / overflow flag, or NULL (legacy)
  const float* lr;                 // per-group learning rates (device), with `state`
};

// found_inf / step bookkeeping on the device (one thread), between sumsq_kernel and adamw_kernel
__global__ void adam_prep_kernel(const float* __restrict__ sumsq, ub200_adam_state* __restrict__ st) {
  pdl_launch_dependents();
  pdl_wait();
  const float s = *sumsq;
  const bool finite = (s == s) && (fabsf(s) <= 3.0e38f);
  st->found_inf = finite ? 0 : 1;
  if (finite) st->step += 1; else st->skipped += 1;
}

__global__ void __launch_bounds__(256)
adamw_kernel(const ub200_adam_segment* __restrict__ segs, const int* __restrict__ blk_start, int nseg,
             const AdamHyper h) {
  pdl_launch_dependents();
  pdl_wait();
  const int s = find_segment(blk_start, nseg, blockIdx.x);
  const ub200_adam_segment sg = segs[s];
  const long long base = static_cast<long long>(blockIdx.x - blk_start[s]) * ADAM_CHUNK;
  float step_size = sg.step_size, lr_wd = sg.lr_wd;
  if (h.state != nullptr) {
    // an overflowed gradient (fp16 loss scaling) must not touch masters / moments / weights:
    // g * 0 would be NaN for g = inf (apex's dynamic scaler skips such a step)
    if (h.state->found_inf) return;
    const float lr = __ldg(h.lr + sg.group);
    step_size = lr;
    if (sg.flags & 1) {
      // once per CTA, in double: 1 - beta^t loses all its digits in fp32 for beta = 0.999, t = 1
      const double t = static_cast<double>(h.state->step - sg.step_offset);
      step_size = static_cast<float>(static_cast<double>(lr) * sqrt(1.0 - pow(static_cast<double>(h.beta2), t)) /
                                     (1.0 - pow(static_cast<double>(h.beta1), t)));
    }
    lr_wd = lr * sg.weight_decay;
  }
  float gmul = h.inv_scale;
  if (h.max_norm > 0.f && h.sumsq != nullptr) {
    // torch.nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), applied iff < 1
    const float total = sqrtf(__ldg(h.sumsq)) * h.inv_scale;
    const float coef = h.max_norm / (total + 1e-6f);
    if (coef < 1.f) gmul *= coef;
  }
  auto update = [&](float g, float& m, float& v, float& p) {
    g *= gmul;
    m = m * h.beta1 + (1.0f - h.beta1) * g;              // optim/adamw.py:77
    v = v * h.beta2 + (1.0f - h.beta2) * g * g;          // :78
    const float denom = sqrtf(v) + h.eps;                // :79
    p = p - step_size * (m / denom);                     // :81-88 (bias-corrected step size)
    if (lr_wd > 0.f) p = p - lr_wd * p;                  // :99-100 decoupled decay AFTER the update
  };
  // vector path: 4 elements per thread (16-byte fp32 accesses, 8-byte 16-bit accesses) when the
  // segment is a 16-bit parameter with a 16-bit gradient of the same type and everything is aligned
  const bool vec = (sg.n % 4 == 0) && sg.model != nullptr && sg.grad_dtype == sg.model_dtype &&
                   sg.grad_dtype != UB200_F32 &&
                   ((reinterpret_cast<uintptr_t>(sg.master) | reinterpret_cast<uintptr_t>(sg.exp_avg) |
                     reinterpret_cast<uintptr_t>(sg.exp_avg_sq)) & 15) == 0 &&
                   ((reinterpret_cast<uintptr_t>(sg.grad) | reinterpret_cast<uintptr_t>(sg.model)) & 7) == 0;
  if (vec) {
    const bool bf = sg.grad_dtype == UB200_BF16;
    for (int j = threadIdx.x * 4; j < ADAM_CHUNK; j += 1024) {
      const long long i = base + j;
      if (i >= sg.n) break;
      const uint2 graw = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(sg.grad) + i);
      float4 m4 = *reinterpret_cast<const float4*>(sg.exp_avg + i);
      float4 v4 = *reinterpret_cast<const float4*>(sg.exp_avg_sq + i);
      float4 p4 = *reinterpret_cast<const float4*>(sg.master + i);
      const float2 g01 = bf ? Elem<true>::unpack(graw.x) : Elem<false>::unpack(graw.x);
      const float2 g23 = bf ? Elem<true>::unpack(graw.y) : Elem<false>::unpack(graw.y);
      update(g01.x, m4.x, v4.x, p4.x);
      update(g01.y, m4.y, v4.y, p4.y);
      update(g23.x, m4.z, v4.z, p4.z);
      update(g23.y, m4.w, v4.w, p4.w);
      *reinterpret_cast<float4*>(sg.exp_avg + i) = m4;
      *reinterpret_cast<float4*>(sg.exp_avg_sq + i) = v4;
      *reinterpret_cast<float4*>(sg.master + i) = p4;
      uint2 o;
      o.x = bf ? Elem<true>::pack(p4.x, p4.y) : Elem<false>::pack(p4.x, p4.y);
      o.y = bf ? Elem<true>::pack(p4.z, p4.w) : Elem<false>::pack(p4.z, p4.w);
      *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(sg.model) + i) = o;
    }
    return;
  }
  for (int j = threadIdx.x; j < ADAM_CHUNK; j += 256) {
    const long long i = base + j;
    if (i >= sg.n) break;
    float m = sg.exp_avg[i], v = sg.exp_avg_sq[i], p = sg.master[i];
    update(grad_of(sg, i), m, v, p);
    sg.exp_avg[i] = m; sg.exp_avg_sq[i] = v; sg.master[i] = p;
    if (sg.model) {
      if (sg.model_dtype == UB200_BF16) reinterpret_cast<__nv_bfloat16*>(sg.model)[i] = __float2bfloat16_rn(p);
      else if (sg.model_dtype == UB200_F16) reinterpret_cast<__half*>(sg.model)[i] = __float2half_rn(p);
      else reinterpret_cast<float*>(sg.model)[i] = p;
    }
  }
}

}  // namespace ub

// ------------------------------------------------------------------------------ C ABI
extern "C" int ub200_ce_fwd(const void* logits, int64_t ld, const int64_t* targets, float* loss,
                            float* lse, int32_t rows, int32_t vocab, int32_t dtype,
                            ub200_stream_t stream_) {
  using namespace ub;
  UB_CHECK_ARG(logits && targets && loss && lse, "ce_fwd: null pointer");
  UB_CHECK_ARG(rows > 0 && vocab > 0 && ld % 8 == 0 && ld >= (vocab + 7) / 8 * 8,
               "ce_fwd: need rows > 0, vocab > 0, ld %% 8 == 0 and ld >= vocab rounded up to 8");
  UB_CHECK_ARG((reinterpret_cast<uintptr_t>(logits) & 15) == 0, "ce_fwd: logits must be 16-byte aligned");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ProfScope ps(stream);
  const long long* tg = reinterpret_cast<const long long*>(targets);
  if (dtype == UB200_BF16)
    UB_CHECK_CUDA(launch_pdl(ce_fwd_kernel<true>, dim3(rows), dim3(256), 0, stream, 1, logits,
                             static_cast<long long>(ld), tg, loss, lse, vocab));
  else
    UB_CHECK_CUDA(launch_pdl(ce_fwd_kernel<false>, dim3(rows), dim3(256), 0, stream, 1, logits,
                             static_cast<long long>(ld), tg, loss, lse, vocab));
  return 0;
}

extern "C" int ub200_ce_bwd(const void* logits, void* dlogits, int64_t ld, const int64_t* targets,
                            const float* lse, const float* dloss, int32_t rows, int32_t vocab,
                            int32_t ncols, int32_t dtype, ub200_stream_t stream_) {
  using namespace ub;
  UB_CHECK_ARG(logits && dlogits && targets && lse && dloss, "ce_bwd: null pointer");
  UB_CHECK_ARG(rows > 0 && vocab > 0 && ncols % 8 == 0 && ncols >= vocab && ld % 8 == 0 && ld >= ncols,
               "ce_bwd: need vocab <= ncols <= ld, ncols %% 8 == 0, ld %% 8 == 0");
  UB_CHECK_ARG((reinterpret_cast<uintptr_t>(logits) & 15) == 0 && (reinterpret_cast<uintptr_t>(dlogits) & 15) == 0,
               "ce_bwd: buffers must be 16-byte aligned");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ProfScope ps(stream);
  const long long* tg = reinterpret_cast<const long long*>(targets);
  if (dtype == UB200_BF16)
    UB_CHECK_CUDA(launch_pdl(ce_bwd_kernel<true>, dim3(rows), dim3(256), 0, stream, 1, logits, dlogits,
                             static_cast<long long>(ld), tg, lse, dloss, vocab, ncols));
  else
    UB_CHECK_CUDA(launch_pdl(ce_bwd_kernel<false>, dim3(rows), dim3(256), 0, stream, 1, logits, dlogits,
                             static_cast<long long>(ld), tg, lse, dloss, vocab, ncols));
  return 0;
}

extern "C" int ub200_dgelu_mul(const void* dy, const void* pre, void* out, int64_t n, int32_t dtype,
                               ub200_stream_t stream_) {
  using namespace ub;
  UB_CHECK_ARG(dy && pre && out, "dgelu_mul: null pointer");
  UB_CHECK_ARG(n > 0 && n % 8 == 0, "dgelu_mul: n must be a positive multiple of 8");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const long long nvec = n / 8;
  long long blocks = (nvec + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  ProfScope ps(stream);
  if (dtype == UB200_BF16)
    UB_CHECK_CUDA(launch_pdl(dgelu_mul_kernel<true>, dim3(static_cast<int>(blocks)), dim3(256), 0, stream, 1,
                             reinterpret_cast<const uint4*>(dy), reinterpret_cast<const uint4*>(pre),
                             reinterpret_cast<uint4*>(out), nvec));
  else
    UB_CHECK_CUDA(launch_pdl(dgelu_mul_kernel<false>, dim3(static_cast<int>(blocks)), dim3(256), 0, stream, 1,
                             reinterpret_cast<const uint4*>(dy), reinterpret_cast<const uint4*>(pre),
                             reinterpret_cast<uint4*>(out), nvec));
  return 0;
}

extern "C" int ub200_dtanh_mul(const void* dy, const void* y, void* out, int64_t n, int32_t dtype,
                               ub200_stream_t stream_) {
  using namespace ub;
  UB_CHECK_ARG(dy && y && out, "dtanh_mul: null pointer");
  UB_CHECK_ARG(n > 0 && n % 8 == 0, "dtanh_mul: n must be a positive multiple of 8");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const long long nvec = n / 8;
  long long blocks = (nvec + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  ProfScope ps(stream);
  if (dtype == UB200_BF16)
    UB_CHECK_CUDA(launch_pdl(dgelu_mul_kernel<true, true>, dim3(static_cast<int>(blocks)), dim3(256), 0, stream, 1,
                             reinterpret_cast<const uint4*>(dy), reinterpret_cast<const uint4*>(y),
                             reinterpret_cast<uint4*>(out), nvec));
  else
    UB_CHECK_CUDA(launch_pdl(dgelu_mul_kernel<false, true>, dim3(static_cast<int>(blocks)), dim3(256), 0, stream, 1,
                             reinterpret_cast<const uint4*>(dy), reinterpret_cast<const uint4*>(y),
                             reinterpret_cast<uint4*>(out), nvec));
  return 0;
}

extern "C" int32_t ub200_adam_chunk(void) { return ub::ADAM_CHUNK; }

extern "C" int ub200_grad_sumsq(const ub200_adam_segment* segs_dev, const int32_t* blk_start_dev,
                                int32_t nseg, int32_t nblocks, float* out, ub200_stream_t stream_) {
  using namespace ub;
  UB_CHECK_ARG(segs_dev && blk_start_dev && out && nseg > 0 && nblocks > 0, "grad_sumsq: bad argument");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ProfScope ps(stream);
  UB_CHECK_CUDA(launch_pdl(sumsq_kernel, dim3(nblocks), dim3(256), 0, stream, 1, segs_dev, blk_start_dev,
                           nseg, out));
  return 0;
}

extern "C" int ub200_adam_prep(const float* sumsq, ub200_adam_state* state_dev, ub200_stream_t stream_) {
  using namespace ub;
  UB_CHECK_ARG(sumsq && state_dev, "adam_prep: null pointer");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ProfScope ps(stream);
  UB_CHECK_CUDA(launch_pdl(adam_prep_kernel, dim3(1), dim3(1), 0, stream, 1, sumsq, state_dev));
  return 0;
}

extern "C" int ub200_adamw_step(const ub200_adam_segment* segs_dev, const int32_t* blk_start_dev,
                                int32_t nseg, int32_t nblocks, float beta1, float beta2, float eps,
                                float inv_scale, float max_norm, const float* sumsq,
                                const ub200_adam_state* state_dev, const float* lr_dev,
                                ub200_stream_t stream_) {
  using namespace ub;
  UB_CHECK_ARG(segs_dev && blk_start_dev && nseg > 0 && nblocks > 0, "adamw_step: bad argument");
  UB_CHECK_ARG((state_dev == nullptr) == (lr_dev == nullptr), "adamw_step: state_dev and lr_dev go together");
  UB_CHECK_ARG(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f,
               "adamw_step: invalid hyper-parameters");
  UB_CHECK_ARG(max_norm <= 0.f || sumsq, "adamw_step: clipping needs the sumsq scalar");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  AdamHyper h{beta1, beta2, eps, inv_scale, max_norm, sumsq, state_dev, lr_dev};
  ProfScope ps(stream);
  UB_CHECK_CUDA(launch_pdl(adamw_kernel, dim3(nblocks), dim3(256), 0, stream, 1, segs_dev, blk_start_dev,
                           nseg, h));
  return 0;
}
