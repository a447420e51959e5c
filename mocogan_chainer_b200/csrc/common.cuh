// common.cuh — shared host/device helpers for libmcg.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/mcg.h"

namespace mcg {

// ------------------------------------------------------------------ host-side error + launch accounting
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

#define MCG_FAIL(code, ...)          \
  do {                               \
    ::mcg::set_error(__VA_ARGS__);   \
    return (code);                   \
  } while (0)

#define MCG_CHECK_LAUNCH(name)                                                          \
  do {                                                                                  \
    ::mcg::g_launches.fetch_add(1, std::memory_order_relaxed);                          \
    cudaError_t e_ = cudaGetLastError();                                                \
    if (e_ != cudaSuccess) MCG_FAIL((int)e_, "%s: launch failed: %s", name, cudaGetErrorString(e_)); \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ------------------------------------------------------------------ programmatic dependent launch (PDL)
// A step is ~270 short kernels in dependency chains, so the gap between a kernel's last CTA and the first CTA of the
// next one is paid ~150 times on the critical path.  With MCG_PDL=1 every kernel of this library is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: the next kernel of the stream may be scheduled while its
// predecessor drains, runs its prologue (barrier init, TMEM allocation, descriptor prefetch, index set-up), and blocks in
// pdl_wait() — EVERY kernel's first action before touching global memory — until the predecessor has completed and its
// writes are visible.  Because every kernel waits, completion stays transitive along a stream; kernels of other
// libraries (torch, NCCL) in between are launched normally and act as full barriers.  The persistent tcgen05 kernels
// additionally call pdl_launch_dependents() on entry (they own their SM for their whole life, so an early-scheduled
// successor only occupies spare thread slots); without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#ifndef MCG_PDL_DEFAULT
#define MCG_PDL_DEFAULT 0
#endif
// Streaming kernels: wait on entry.  MCG_PDL_EW_TRIGGER=1 (build flag) also lets THEIR successor be scheduled early; it
// is off by default because an early-scheduled persistent convolution camps on the SM's shared memory while it waits
// and keeps the other streams' convolutions out (branch concurrency).
#ifndef MCG_PDL_EW_TRIGGER
#define MCG_PDL_EW_TRIGGER 0
#endif
__device__ __forceinline__ void pdl_enter() {
#if MCG_PDL_EW_TRIGGER
  pdl_launch_dependents();
#endif
  pdl_wait();
}
int pdl_level();   // MCG_PDL (default MCG_PDL_DEFAULT), read once (elementwise.cu)

template <typename... KArgs>
struct PdlLaunch {
  void (*kern)(KArgs...);
  dim3 grid, block;
  size_t smem;
  cudaStream_t st;
  template <typename... Args>
  void operator()(Args&&... args) const {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_level() > 0 ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, static_cast<Args&&>(args)...);   // errors surface in MCG_CHECK_LAUNCH
  }
};
template <typename... KArgs>
static inline PdlLaunch<KArgs...> pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  return PdlLaunch<KArgs...>{kern, grid, block, smem, st};
}
static inline int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ------------------------------------------------------------------ element access by runtime/compile-time dtype
template <typename T> __device__ __forceinline__ float ld(const T* p, long long i);
template <> __device__ __forceinline__ float ld<float>(const float* p, long long i) { return p[i]; }
template <> __device__ __forceinline__ float ld<__nv_bfloat16>(const __nv_bfloat16* p, long long i) {
  return __bfloat162float(p[i]);
}
// uint8 pixels are read already normalised: (v - 128) / 128, the reference's datasets.py:91 (input pipeline only)
template <> __device__ __forceinline__ float ld<unsigned char>(const unsigned char* p, long long i) {
  return ((float)p[i] - 128.f) / 128.f;
}
template <typename T> __device__ __forceinline__ void st(T* p, long long i, float v);
template <> __device__ __forceinline__ void st<float>(float* p, long long i, float v) { p[i] = v; }
template <> __device__ __forceinline__ void st<__nv_bfloat16>(__nv_bfloat16* p, long long i, float v) {
  p[i] = __float2bfloat16_rn(v);
}

// 8 consecutive elements <-> 8 floats (16 B for bf16, 32 B for fp32); pointers must be 16-byte aligned.
template <typename T> __device__ __forceinline__ void ld8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void ld8<float>(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void ld8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
// The same 8 elements held as loaded (4 registers for bf16, 8 for fp32): a streaming kernel issues the loads of several
// rows back to back and converts afterwards, so every thread keeps 64-128 B in flight instead of 16-32 B.  With 2-4
// resident blocks per SM that is the ~64 KB per SM HBM3e needs; one row per iteration left these kernels at 40-60 % of
// the measured copy bandwidth.
template <typename T> struct Raw8;
template <> struct Raw8<float> { float4 a, b; };
template <> struct Raw8<__nv_bfloat16> { uint4 u; };
template <typename T> __device__ __forceinline__ Raw8<T> ldraw8(const T* p);
template <> __device__ __forceinline__ Raw8<float> ldraw8<float>(const float* p) {
  Raw8<float> r;
  r.a = *reinterpret_cast<const float4*>(p);
  r.b = *reinterpret_cast<const float4*>(p + 4);
  return r;
}
template <> __device__ __forceinline__ Raw8<__nv_bfloat16> ldraw8<__nv_bfloat16>(const __nv_bfloat16* p) {
  Raw8<__nv_bfloat16> r;
  r.u = *reinterpret_cast<const uint4*>(p);
  return r;
}
__device__ __forceinline__ void unpack8(const Raw8<float>& r, float (&v)[8]) {
  v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
}
__device__ __forceinline__ void unpack8(const Raw8<__nv_bfloat16>& r, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r.u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
// Rows in flight per thread / minimum resident blocks per SM (= register cap) of the three hot streaming kernels; the
// defaults are the best whole-step combination of profiles/r01_stream_kernels_ab.txt.  A kernel that is fastest ALONE
// (many rows in flight, 128 registers) is not the fastest inside the 4-stream step: its blocks must fit beside a
// persistent convolution CTA (38 K of the SM's 64 K registers), so the register cap matters more than the unroll.
#ifndef MCG_RED_U
#define MCG_RED_U 4      // one-tensor reductions (statistics, bias gradients)
#endif
#ifndef MCG_RED_U2
#define MCG_RED_U2 2     // two-tensor reductions (BatchNorm backward sums)
#endif
#ifndef MCG_RED_MB
#define MCG_RED_MB 3     // ncu: at 85-89 registers the BatchNorm-backward reduction fitted only 2 blocks per SM
#endif
#ifndef MCG_AFF_U
#define MCG_AFF_U 1
#endif
#ifndef MCG_AFF_MB
#define MCG_AFF_MB 3
#endif
#ifndef MCG_APP_U
#define MCG_APP_U 1
#endif
#ifndef MCG_APP_MB
#define MCG_APP_MB 3
#endif

template <typename T> __device__ __forceinline__ void st8(T* p, const float (&v)[8]);
template <> __device__ __forceinline__ void st8<float>(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void st8<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

// ------------------------------------------------------------------ activations
__device__ __forceinline__ float act_fwd(int act, float slope, float x) {
  switch (act) {
    case MCG_ACT_RELU: return x > 0.f ? x : 0.f;
    case MCG_ACT_LRELU: return x >= 0.f ? x : slope * x;
    case MCG_ACT_TANH: return tanhf(x);
    default: return x;
  }
}
// derivative given the pre-activation (or, for tanh/relu with use_output, the output)
__device__ __forceinline__ float act_grad(int act, float slope, float v, int is_output) {
  switch (act) {
    case MCG_ACT_RELU: return v > 0.f ? 1.f : 0.f;
    case MCG_ACT_LRELU: return v >= 0.f ? 1.f : slope;
    case MCG_ACT_TANH: {
      float o = is_output ? v : tanhf(v);
      return 1.f - o * o;
    }
    default: return 1.f;
  }
}

// ------------------------------------------------------------------ Philox4x32-10 (counter-based RNG)
struct StepState {  // mirrors the 8 x uint32 layout documented in mcg.h
  uint32_t seed_lo, seed_hi, step, frame_t, adam_t, r0, r1, r2;
};

__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float u32_to_unit(uint32_t x) {  // (0,1]
  return (float)(x >> 8) * (1.0f / 16777216.0f) + (0.5f / 16777216.0f);
}
// four independent N(0,1) draws for block `idx4` of stream (call_id, step)
__device__ __forceinline__ void philox_normal4(const StepState* s, int call_id, unsigned long long idx4,
                                               float (&n)[4]) {
  uint4 c = make_uint4((uint32_t)idx4, (uint32_t)(idx4 >> 32), (uint32_t)call_id, s->step);
  uint4 r = philox4x32(c, make_uint2(s->seed_lo, s->seed_hi));
  float u0 = u32_to_unit(r.x), u1 = u32_to_unit(r.y), u2 = u32_to_unit(r.z), u3 = u32_to_unit(r.w);
  float ra = sqrtf(-2.f * __logf(u0)), rb = sqrtf(-2.f * __logf(u2));
  float s0, c0, s1, c1;
  __sincosf(6.28318530718f * u1, &s0, &c0);
  __sincosf(6.28318530718f * u3, &s1, &c1);
  n[0] = ra * c0; n[1] = ra * s0; n[2] = rb * c1; n[3] = rb * s1;
}

}  // namespace mcg
