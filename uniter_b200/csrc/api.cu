// Library identity, error channel and TMA descriptor encoding for libub200.so.
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>

#include "common.h"

namespace ub {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static int g_sm_reserve = 0;

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  int n = cached[dev] - g_sm_reserve;
  n &= ~1;                 // CTA pairs (cta_group::2) need an even count
  return n < 2 ? 2 : n;
}

int set_sm_reserve(int n) {
  const int prev = g_sm_reserve;
  g_sm_reserve = n < 0 ? 0 : n;
  return prev;
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

int make_tma_2d(CUtensorMap* out, const void* base, int dtype, uint64_t rows, uint64_t cols,
                uint64_t ld, uint32_t box_rows, uint32_t box_cols) {
  auto enc = get_encode();
  if (enc == nullptr)
    return set_error(UB200_ECUDA, "cuTensorMapEncodeTiled driver entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0)
    return set_error(UB200_EINVAL, "TMA operand must be 16-byte aligned with pitch %% 8 == 0 "
                                   "(base=%p ld=%llu)", base, (unsigned long long)ld);
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out,
                   dtype == UB200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                       : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                   2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(UB200_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d "
                                  "(rows=%llu cols=%llu ld=%llu box=%ux%u)", (int)r,
                     (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld,
                     box_rows, box_cols);
  return 0;
}


thread_local int g_prof_tag = 0;
static bool g_prof_on = false;
static unsigned long long g_launches = 0;
struct ProfRec { int tag; cudaEvent_t e0, e1; };
static ProfRec g_recs[8192];
static int g_nrec = 0;

ProfScope::ProfScope(cudaStream_t s) : stream(s), slot(-1) {
  ++g_launches;
  if (g_prof_on && g_nrec < 8192) {
    slot = g_nrec++;
    ProfRec& r = g_recs[slot];
    r.tag = g_prof_tag;
    if (!r.e0) { cudaEventCreate(&r.e0); cudaEventCreate(&r.e1); }
    cudaEventRecord(r.e0, stream);
  }
}
ProfScope::~ProfScope() {
  if (slot >= 0) cudaEventRecord(g_recs[slot].e1, stream);
}

}  // namespace ub

extern "C" {

unsigned long long ub200_launch_count(void) { return ub::g_launches; }

int ub200_set_sm_reserve(int n) { return ub::set_sm_reserve(n); }

int ub200_profile_enable(int on) {
  ub::g_prof_on = on != 0;
  if (on) ub::g_nrec = 0;
  return 0;
}

// Synchronises the device, sums the recorded launch durations per tag (ms) and launch counts
// per tag into ms_out[ntags] / count_out[ntags]; clears the record list.
int ub200_profile_collect(float* ms_out, int* count_out, int ntags) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return ub::set_error(UB200_ECUDA, "profile_collect: %s", cudaGetErrorString(e));
  for (int i = 0; i < ntags; ++i) { ms_out[i] = 0.f; count_out[i] = 0; }
  for (int i = 0; i < ub::g_nrec; ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ub::g_recs[i].e0, ub::g_recs[i].e1);
    const int t = ub::g_recs[i].tag;
    if (t >= 0 && t < ntags) { ms_out[t] += ms; count_out[t] += 1; }
  }
  ub::g_nrec = 0;
  return 0;
}


int ub200_version(void) { return 100; /* 0.1.0 */ }

const char* ub200_last_error_string(void) { return ub::g_err; }

int ub200_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return ub::set_error(UB200_ECUDA, "no CUDA device: %s", cudaGetErrorString(e));
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10)
    return ub::set_error(UB200_EUNSUPPORTED, "libub200 is built for sm_100a only; device is sm_%d%d",
                         major, minor);
  return 0;
}

}  // extern "C"
