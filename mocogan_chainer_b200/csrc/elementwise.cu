// elementwise.cu — the HBM-bound passes of the MoCoGAN step: per-channel reductions (BatchNorm statistics, bias
// gradients, BN backward sums), the fused affine+activation+noise pass, video packing, tanh backward, Adam.
// All tensors are channels-last matrices [M][C]; vector kernels move 8 channels (16 B bf16 / 32 B fp32) per thread.
#include "common.cuh"
#include <stdarg.h>

namespace mcg {

int pdl_level() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MCG_PDL");
    v = e ? atoi(e) : MCG_PDL_DEFAULT;
  }
  return v;
}
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
std::atomic<long long> g_launches{0};

// =====================================================================================================
// Column reduction skeleton: two sums per channel, deterministic (per-block partials + ordered finalize).
// =====================================================================================================
constexpr int kRedThreads = 256;
constexpr int kRedMaxBlocks = 592;  // 4 x 148 SMs

static int pow2_ge(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

// Reduction functors.  `prep` loads the per-channel constants of the thread's channel group ONCE (a thread keeps the
// same channels for every row it visits); `operator()` then touches only the streamed tensors.
struct StatsF {  // sum y, sum y^2
  template <int VEC> struct Regs {};
  template <int VEC> __device__ __forceinline__ void prep(int, Regs<VEC>&) const {}
  static constexpr bool kTwo = false, kHasB = true;
  template <int VEC>
  __device__ __forceinline__ void acc(const float (&v)[VEC], const float (&)[VEC], const Regs<VEC>&, float (&a)[VEC],
                                      float (&b)[VEC]) const {
#pragma unroll
    for (int i = 0; i < VEC; ++i) { a[i] += v[i]; b[i] += v[i] * v[i]; }
  }
};
struct SumF {  // sum g
  template <int VEC> struct Regs {};
  template <int VEC> __device__ __forceinline__ void prep(int, Regs<VEC>&) const {}
  static constexpr bool kTwo = false, kHasB = false;
  template <int VEC>
  __device__ __forceinline__ void acc(const float (&v)[VEC], const float (&)[VEC], const Regs<VEC>&, float (&a)[VEC],
                                      float (&)[VEC]) const {
#pragma unroll
    for (int i = 0; i < VEC; ++i) a[i] += v[i];
  }
};
// ACT >= 0 fixes the activation at compile time (the hot layers: the runtime switch per element made these passes
// instruction-issue-bound, ~40 instructions per element); ACT = -1 keeps the runtime code.
template <int ACT>
struct BnBwdF {  // sum g', sum g'*xhat with g' = g * act'(scale*y+shift)
  const float *mean, *invstd, *scale, *shift;
  int act;
  float slope;
  template <int VEC> struct Regs { float mean[VEC], invstd[VEC], scale[VEC], shift[VEC]; };
  template <int VEC> __device__ __forceinline__ void prep(int c0, Regs<VEC>& r) const {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      r.mean[i] = mean[c0 + i]; r.invstd[i] = invstd[c0 + i]; r.scale[i] = scale[c0 + i]; r.shift[i] = shift[c0 + i];
    }
  }
  static constexpr bool kTwo = true, kHasB = true;
  template <int VEC>
  __device__ __forceinline__ void acc(const float (&gv)[VEC], const float (&yv)[VEC], const Regs<VEC>& r, float (&a)[VEC],
                                      float (&b)[VEC]) const {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float pre = r.scale[i] * yv[i] + r.shift[i];
      float gp = gv[i] * act_grad(ACT >= 0 ? ACT : act, slope, pre, 0);
      a[i] += gp;
      b[i] += gp * (yv[i] - r.mean[i]) * r.invstd[i];
    }
  }
};

// How a column reduction ends.  The separate finalize kernels (one warp per channel over <= 592 per-block partial rows)
// were ~50 launches of 5-8 us per step, each a link in a stats -> finalize -> apply dependency chain.  Now every block adds
// its column sums into one of kRedSlots slot rows with red.add (<= 37 adds per address — a single row serialised ~600) and
// takes a ticket; the block that draws the last ticket sums the 16 slot rows per channel in double, runs the finaliser
// (statistics -> mean / invstd / scale / shift / running averages, or plain sums -> outputs) and hands the slot rows and
// the ticket counter back ZEROED, which is the state every call expects to find them in (mcg.h: the workspace of these
// entry points must be zero-filled once, at allocation).
constexpr int kRedSlots = 16;
struct StatsFinal {
  double M;
  const float *gamma, *beta;
  float eps, decay;
  float *mean, *invstd, *scale, *shift, *avg_mean, *avg_var;
  __device__ __forceinline__ void operator()(int c, double s, double q) const {
    double mu = s / M;
    double var = q / M - mu * mu;
    if (var < 0) var = 0;
    double inv = 1.0 / sqrt(var + (double)eps);
    mean[c] = (float)mu;
    invstd[c] = (float)inv;
    float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
    float sc = ga * (float)inv;
    if (scale) scale[c] = sc;
    if (shift) shift[c] = be - (float)mu * sc;
    if (avg_mean) {
      double adjust = M / (M - 1.0 > 1.0 ? M - 1.0 : 1.0);
      avg_mean[c] = decay * avg_mean[c] + (1.f - decay) * (float)mu;
      avg_var[c] = decay * avg_var[c] + (1.f - decay) * (float)(adjust * (var + (double)eps));
    }
  }
};
struct Sum2Final {
  float *out_a, *out_b;
  int accumulate;
  float *acc_a, *acc_b;
  __device__ __forceinline__ void operator()(int c, double s, double q) const {
    // accumulating targets are parameter gradients: the real-clip and fake-clip branches of a pass may finish on two
    // streams at once, so they are added atomically
    if (out_a) { if (accumulate) atomicAdd(out_a + c, (float)s); else out_a[c] = (float)s; }
    if (out_b) { if (accumulate) atomicAdd(out_b + c, (float)q); else out_b[c] = (float)q; }
    if (acc_a) atomicAdd(acc_a + c, (float)s);
    if (acc_b) atomicAdd(acc_b + c, (float)q);
  }
};

template <typename T, int VEC, typename F, typename FIN>
__global__ void __launch_bounds__(kRedThreads, F::kTwo ? MCG_RED_MB : 4) colreduce_kernel(F f, const T* p0, const T* p1, long long M, int C,
                                                               int tpr, float* __restrict__ slots, FIN fin) {
  pdl_enter();
  extern __shared__ float red[];  // [rpb][tpr][2*VEC]
  const int CG = C / VEC;
  const int rpb = kRedThreads / tpr;
  const int tx = threadIdx.x % tpr, ty = threadIdx.x / tpr;
  float a[VEC], b[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) a[i] = b[i] = 0.f;
  if (tx < CG) {
    typename F::template Regs<VEC> regs;
    f.template prep<VEC>(tx * VEC, regs);
    const long long col = (long long)tx * VEC;
    // A block walks the matrix in TILES of U x rpb consecutive rows (one contiguous span of memory per tensor), U rows per
    // thread with all loads issued before the first use.  Rows of one thread that are a whole grid apart (the first
    // version) kept as many bytes in flight but opened U DRAM pages at once and ran slower than U = 1.
    constexpr int U = (VEC == 8) ? (F::kTwo ? (sizeof(T) == 4 ? 1 : MCG_RED_U2) : MCG_RED_U) : 1;
    const long long tile = (long long)rpb * U;
    long long base = (long long)blockIdx.x * tile;
    if constexpr (VEC == 8) {
      for (; base + tile <= M; base += (long long)gridDim.x * tile) {
        Raw8<T> r0[U], r1[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const long long m = base + (long long)u * rpb + ty;
          r0[u] = ldraw8<T>(p0 + m * C + col);
          if (F::kTwo) r1[u] = ldraw8<T>(p1 + m * C + col);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          float v0[8], v1[8];
          unpack8(r0[u], v0);
          if (F::kTwo) unpack8(r1[u], v1);
          f.template acc<VEC>(reinterpret_cast<const float(&)[VEC]>(v0), reinterpret_cast<const float(&)[VEC]>(v1), regs, a, b);
        }
      }
    }
    // the last, partial tile of the matrix (and everything when VEC == 1): row by row
    for (; base < M; base += (long long)gridDim.x * tile) {
      for (int u = 0; u < U; ++u) {
        const long long m = base + (long long)u * rpb + ty;
        if (m >= M) break;
        float v0[VEC], v1[VEC];
        if (VEC == 8) {
          ld8<T>(p0 + m * C + col, reinterpret_cast<float(&)[8]>(v0));
          if (F::kTwo) ld8<T>(p1 + m * C + col, reinterpret_cast<float(&)[8]>(v1));
        } else {
          v0[0] = ld<T>(p0, m * C + col);
          if (F::kTwo) v1[0] = ld<T>(p1, m * C + col);
        }
        f.template acc<VEC>(v0, v1, regs, a, b);
      }
    }
  }
  float* mine = red + ((size_t)ty * tpr + tx) * 2 * VEC;
#pragma unroll
  for (int i = 0; i < VEC; ++i) { mine[i] = a[i]; mine[VEC + i] = b[i]; }
  __syncthreads();
  if (ty == 0 && tx < CG) {
    for (int r = 1; r < rpb; ++r) {
      const float* o = red + ((size_t)r * tpr + tx) * 2 * VEC;
#pragma unroll
      for (int i = 0; i < VEC; ++i) { a[i] += o[i]; b[i] += o[VEC + i]; }
    }
    float* dst = slots + (size_t)(blockIdx.x % kRedSlots) * 2 * C;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      atomicAdd(dst + tx * VEC + i, a[i]);
      if (F::kHasB) atomicAdd(dst + C + tx * VEC + i, b[i]);
    }
  }
  // ticket: the last block to arrive sees every block's adds (fence before the ticket, fence after it)
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  int* counter = reinterpret_cast<int*>(slots + (size_t)kRedSlots * 2 * C);
  if (threadIdx.x == 0) s_last = atomicAdd(counter, 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int c = threadIdx.x; c < C; c += kRedThreads) {
    float sv[kRedSlots], qv[kRedSlots];
#pragma unroll
    for (int k = 0; k < kRedSlots; ++k) {
      sv[k] = __ldcg(slots + (size_t)k * 2 * C + c);
      qv[k] = F::kHasB ? __ldcg(slots + (size_t)k * 2 * C + C + c) : 0.f;
    }
    double sd = 0, qd = 0;
#pragma unroll
    for (int k = 0; k < kRedSlots; ++k) { sd += (double)sv[k]; qd += (double)qv[k]; }
#pragma unroll
    for (int k = 0; k < kRedSlots; ++k) {
      slots[(size_t)k * 2 * C + c] = 0.f;
      if (F::kHasB) slots[(size_t)k * 2 * C + C + c] = 0.f;
    }
    fin(c, sd, qd);
  }
  if (threadIdx.x == 0) *counter = 0;
}

template <typename F, typename FIN>
static int launch_colreduce(F f, FIN fin, const void* p0, const void* p1, long long M, int C, int dtype, void* ws,
                            size_t ws_bytes, cudaStream_t st, const char* name) {
  if (M <= 0 || C <= 0) MCG_FAIL(MCG_ERR_SHAPE, "%s: empty matrix M=%lld C=%d", name, M, C);
  const int VEC = (C % 8 == 0) ? 8 : 1;
  const int CG = C / VEC;
  if (CG > kRedThreads) MCG_FAIL(MCG_ERR_UNSUPPORTED, "%s: C=%d has too many channel groups", name, C);
  const int tpr = pow2_ge(CG);
  const int rpb = kRedThreads / tpr;
  long long want = (M + (long long)rpb * 8 - 1) / ((long long)rpb * 8);  // >= 8 rows per thread: few partials to finalize
  size_t smem = (size_t)kRedThreads * 2 * VEC * sizeof(float);
  // one wave: as many blocks as are resident at once for THIS instantiation (3 per SM for the BatchNorm-backward sums at
  // 85 registers, 4 for the statistics) — 592 blocks of a 3-per-SM kernel ran as a full wave plus a third of one
  int cap = kRedMaxBlocks;
  {
    static std::atomic<int> cached[2][2];      // [dtype][VEC == 8] -> resident blocks per SM (+1; 0 = not asked yet)
    std::atomic<int>& slot = cached[dtype == MCG_F32 ? 0 : 1][VEC == 8 ? 1 : 0];
    int per_sm = slot.load(std::memory_order_relaxed) - 1;
    if (per_sm < 0) {
      per_sm = 0;
      cudaError_t e = cudaSuccess;
      if (dtype == MCG_F32) e = VEC == 8 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, colreduce_kernel<float, 8, F, FIN>, kRedThreads, smem)
                                         : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, colreduce_kernel<float, 1, F, FIN>, kRedThreads, smem);
      else e = VEC == 8 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, colreduce_kernel<__nv_bfloat16, 8, F, FIN>, kRedThreads, smem)
                        : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, colreduce_kernel<__nv_bfloat16, 1, F, FIN>, kRedThreads, smem);
      if (e != cudaSuccess) { per_sm = 0; (void)cudaGetLastError(); }
      slot.store(per_sm + 1, std::memory_order_relaxed);
    }
    if (per_sm > 0 && per_sm * num_sms() < cap) cap = per_sm * num_sms();
  }
  int nblk = (int)(want < cap ? want : cap);
  if (nblk < 1) nblk = 1;
  const size_t need = (size_t)kRedSlots * 2 * C * sizeof(float) + 16;
  if (ws_bytes < need || !ws) MCG_FAIL(MCG_ERR_WORKSPACE, "%s: workspace %zu < %zu", name, ws_bytes, need);
  float* part = reinterpret_cast<float*>(ws);   // kRedSlots x 2C slot sums + the ticket counter, zero on entry and on exit
#define LAUNCH(T, V) pdl(colreduce_kernel<T, V, F, FIN>, nblk, kRedThreads, smem, st)(f, (const T*)p0, (const T*)p1, M, C, tpr, part, fin)
  if (dtype == MCG_F32) { if (VEC == 8) LAUNCH(float, 8); else LAUNCH(float, 1); }
  else if (dtype == MCG_BF16) { if (VEC == 8) LAUNCH(__nv_bfloat16, 8); else LAUNCH(__nv_bfloat16, 1); }
  else MCG_FAIL(MCG_ERR_UNSUPPORTED, "%s: dtype %d", name, dtype);
#undef LAUNCH
  MCG_CHECK_LAUNCH(name);
  return 0;
}

// =====================================================================================================
// affine + activation + noise
// =====================================================================================================
// VEC == 8 with FIXED: the launcher made (gridDim.x * 256) a multiple of C/8, so a thread keeps one channel group for all
// of its rows: scale/shift live in registers and the row index advances without a division.
template <typename TI, typename TO, int VEC, bool FIXED, int ACT = -1, bool NOISE = true>
__global__ void __launch_bounds__(256, (FIXED && VEC == 8) ? MCG_AFF_MB : 1) affine_act_noise_kernel(
    const TI* __restrict__ y, long long M, int C, long long P, const float* __restrict__ scale,
    const float* __restrict__ shift, int act, float slope, float sigma, const float* __restrict__ noise,
    long long ns_n, long long ns_c, long long ns_p, const StepState* __restrict__ rng, int call_id,
    TO* __restrict__ out) {
  pdl_enter();
  const int CG = C / VEC;
  const long long total = M * CG;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float sc[VEC], sh[VEC];
  int c0 = (int)(idx % CG) * VEC;
  long long m = idx / CG;
  const long long mstep = stride / CG;
  if (FIXED && scale) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) { sc[i] = scale[c0 + i]; sh[i] = shift[c0 + i]; }
  }
  // one row: affine + activation + noise on 8 (or 1) loaded values, then the store
  auto finish_row = [&](float (&v)[VEC], long long m, int c0, long long idx) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float pre = v[i];
      if (scale) pre = FIXED ? sc[i] * v[i] + sh[i] : scale[c0 + i] * v[i] + shift[c0 + i];
      v[i] = act_fwd(ACT >= 0 ? ACT : act, slope, pre);
    }
    if (!NOISE) {
      // the generator's layers: no add_noise, and none of the Philox code's registers
    } else if (noise) {
      long long n = m / P, p = m % P;
#pragma unroll
      for (int i = 0; i < VEC; ++i) v[i] += sigma * noise[n * ns_n + (long long)(c0 + i) * ns_c + p * ns_p];
    } else if (rng && sigma != 0.f) {
      if (VEC == 8) {
        float z[4];
        philox_normal4(rng, call_id, (unsigned long long)idx * 2, z);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] += sigma * z[i];
        philox_normal4(rng, call_id, (unsigned long long)idx * 2 + 1, z);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[4 + (VEC == 8 ? i : 0)] += sigma * z[i];
      } else {
        float z[4];
        philox_normal4(rng, call_id, (unsigned long long)idx, z);
        v[0] += sigma * z[0];
      }
    }
    if (VEC == 8) st8<TO>(out + m * C + c0, reinterpret_cast<const float(&)[8]>(v));
    else st<TO>(out, m * C + c0, v[0]);
  };
  if constexpr (FIXED && VEC == 8) {
    // MCG_AFF_U rows of the thread's channel group per iteration, all loads issued before the first use
    constexpr int U = MCG_AFF_U;
    for (; idx + (U - 1) * stride < total; idx += U * stride, m += U * mstep) {
      Raw8<TI> raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) raw[u] = ldraw8<TI>(y + (m + u * mstep) * C + c0);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float v[8];
        unpack8(raw[u], v);
        finish_row(v, m + u * mstep, c0, idx + u * stride);
      }
    }
  }
  for (; idx < total; idx += stride) {
    if (!FIXED) { m = idx / CG; c0 = (int)(idx % CG) * VEC; }
    float v[VEC];
    if (VEC == 8) ld8<TI>(y + m * C + c0, reinterpret_cast<float(&)[8]>(v));
    else v[0] = ld<TI>(y, m * C + c0);
    finish_row(v, m, c0, idx);
    if (FIXED) m += mstep;
  }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) pack_video_kernel(
    const TI* __restrict__ src, int N, int C, int T, int H, int W, long long s_n, long long s_c, long long s_t,
    long long s_h, long long s_w, const int* __restrict__ frame_ptr, float sigma, const float* __restrict__ noise,
    long long ns_n, long long ns_c, long long ns_p, const StepState* __restrict__ rng, int call_id,
    TO* __restrict__ out) {
  pdl_enter();
  // one thread per pixel: for a channels-first source the reads of each channel are coalesced along w and the C
  // outputs of a pixel are adjacent, so a warp writes one contiguous span
  const int Tout = frame_ptr ? 1 : T;
  const int t0 = frame_ptr ? *frame_ptr : 0;
  const long long pixels = (long long)N * Tout * H * W;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < pixels;
       pix += (long long)gridDim.x * blockDim.x) {
    long long r = pix;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H); r /= H;
    const int t = (int)(r % Tout);
    const long long n = r / Tout;
    const long long sbase = n * s_n + (long long)(t + t0) * s_t + (long long)h * s_h + (long long)w * s_w;
    const long long p = ((long long)t * H + h) * W + w;  // position inside the (T',H,W) block the noise tensor covers
    for (int c0 = 0; c0 < C; c0 += 4) {
      float z[4] = {0.f, 0.f, 0.f, 0.f};
      if (!noise && rng && sigma != 0.f) philox_normal4(rng, call_id, (unsigned long long)pix * ((C + 3) / 4) + c0 / 4, z);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c0 + i;
        if (c >= C) break;
        float v = ld<TI>(src, sbase + (long long)c * s_c);
        if (noise) v += sigma * noise[n * ns_n + (long long)c * ns_c + p * ns_p];
        else v += sigma * z[i];
        st<TO>(out, pix * C + c, v);
      }
    }
  }
}

// Same operation, one thread per ELEMENT: used when the source is channel-contiguous with many channels (dtype / layout
// conversions of activations), where consecutive threads then read consecutive addresses.
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) pack_elem_kernel(const TI* __restrict__ src, int N, int C, int T, int H, int W,
                                                        long long s_n, long long s_c, long long s_t, long long s_h,
                                                        long long s_w, TO* __restrict__ out) {
  pdl_enter();
  const long long total = (long long)N * T * H * W * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    long long r = idx / C;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H); r /= H;
    const int t = (int)(r % T);
    const long long n = r / T;
    st<TO>(out, idx, ld<TI>(src, n * s_n + (long long)c * s_c + (long long)t * s_t + (long long)h * s_h + (long long)w * s_w));
  }
}

// =====================================================================================================
// backward of (BN ->) activation, apply half
// =====================================================================================================
// With BN:  gy = gamma*invstd*(g' - (xhat*dgamma + dbeta)/M) = A*(g' - ((y - mean)*B + D))  with per-channel A = gamma*invstd,
// B = invstd*dgamma/M, D = dbeta/M.  FIXED (see affine_act_noise_kernel) keeps A, B, D, mean and the activation's
// scale/shift in registers.
template <typename TI, typename TO, int VEC, bool FIXED, int ACT = -1>
__global__ void __launch_bounds__(256, (FIXED && VEC == 8 && ACT >= 0 && sizeof(TI) == 2) ? MCG_APP_MB : 1) act_bn_bwd_apply_kernel(
    const TI* __restrict__ g, const TI* __restrict__ y, long long M, int C, const float* __restrict__ mean,
    const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ scale,
    const float* __restrict__ shift, int act, float slope, int use_output, const float* __restrict__ dgamma,
    const float* __restrict__ dbeta, float inv_m, TO* __restrict__ gy) {
  pdl_enter();
  const int CG = C / VEC;
  const long long total = M * CG;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int c0 = (int)(idx % CG) * VEC;
  long long m = idx / CG;
  const long long mstep = stride / CG;
  float cA[VEC], cB[VEC], cD[VEC], cM[VEC], sc[VEC], sh[VEC];
  if (FIXED && mean) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int c = c0 + i;
      cA[i] = gamma[c] * invstd[c];
      cB[i] = invstd[c] * dgamma[c] * inv_m;
      cD[i] = dbeta[c] * inv_m;
      cM[i] = mean[c];
      sc[i] = scale[c]; sh[i] = shift[c];
    }
  }
  auto finish_row = [&](float (&gv)[VEC], const float (&yv)[VEC], long long m, int c0) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      int c = c0 + i;
      if (mean) {
        if (FIXED) {
          float gp = gv[i] * act_grad(ACT >= 0 ? ACT : act, slope, sc[i] * yv[i] + sh[i], 0);
          gv[i] = cA[i] * (gp - ((yv[i] - cM[i]) * cB[i] + cD[i]));
        } else {
          float pre = scale[c] * yv[i] + shift[c];
          float gp = gv[i] * act_grad(ACT >= 0 ? ACT : act, slope, pre, 0);
          float xh = (yv[i] - mean[c]) * invstd[c];
          gv[i] = gamma[c] * invstd[c] * (gp - (xh * dgamma[c] + dbeta[c]) * inv_m);
        }
      } else {
        gv[i] = gv[i] * act_grad(ACT >= 0 ? ACT : act, slope, yv[i], use_output);
      }
    }
    if (VEC == 8) st8<TO>(gy + m * C + c0, reinterpret_cast<const float(&)[8]>(gv));
    else st<TO>(gy, m * C + c0, gv[0]);
  };
  if constexpr (FIXED && VEC == 8) {
    // MCG_APP_U rows per iteration: 2 x MCG_APP_U 16-byte (bf16) loads in flight per thread (the runtime-activation
    // variant carries the switch's code and would spill with more than 2)
    constexpr int U = ACT >= 0 ? MCG_APP_U : (MCG_APP_U < 2 ? MCG_APP_U : 2);
    for (; idx + (U - 1) * stride < total; idx += U * stride, m += U * mstep) {
      Raw8<TI> rg[U], ry[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        rg[u] = ldraw8<TI>(g + (m + u * mstep) * C + c0);
        ry[u] = ldraw8<TI>(y + (m + u * mstep) * C + c0);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float gv[8], yv[8];
        unpack8(rg[u], gv);
        unpack8(ry[u], yv);
        finish_row(gv, yv, m + u * mstep, c0);
      }
    }
  }
  for (; idx < total; idx += stride) {
    if (!FIXED) { m = idx / CG; c0 = (int)(idx % CG) * VEC; }
    float gv[VEC], yv[VEC];
    if (VEC == 8) {
      ld8<TI>(g + m * C + c0, reinterpret_cast<float(&)[8]>(gv));
      ld8<TI>(y + m * C + c0, reinterpret_cast<float(&)[8]>(yv));
    } else {
      gv[0] = ld<TI>(g, m * C + c0);
      yv[0] = ld<TI>(y, m * C + c0);
    }
    finish_row(gv, yv, m, c0);
    if (FIXED) m += mstep;
  }
}

template <typename TG, typename TO, typename TD>
__global__ void __launch_bounds__(256) tanh_bwd_video_kernel(const TG* __restrict__ gv, const TG* __restrict__ gi,
                                                             const TO* __restrict__ out_tn, int N, int T, int HW,
                                                             int C, const int* __restrict__ frame_ptr,
                                                             TD* __restrict__ g_tn) {
  pdl_enter();
  const int ft = frame_ptr ? *frame_ptr : -1;
  const long long per = (long long)HW * C;
  const long long total = (long long)T * N * per;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long e = idx % per;
    long long r = idx / per;
    int n = (int)(r % N);
    int t = (int)(r / N);
    float g = gv ? ld<TG>(gv, ((long long)n * T + t) * per + e) : 0.f;
    if (gi && t == ft) g += ld<TG>(gi, (long long)n * per + e);
    float o = ld<TO>(out_tn, idx);
    st<TD>(g_tn, idx, g * (1.f - o * o));
  }
}

// Generated clip -> uint8 pictures: videos (T*N, H, W, C) channels-last in tanh range (the generator's output storage)
// -> u8 (T, N, C, H, W) = ((v / 2 + 0.5) * 255) truncated (generate_samples.py:39) and, when grid != NULL, the
// size x size tiling of util.py:30-51 `to_grid` (T, C, size*H, size*W) in the same pass.
template <typename T>
__global__ void __launch_bounds__(256) video_to_uint8_kernel(const T* __restrict__ v, int Tn, int N, int C, int H, int W,
                                                             unsigned char* __restrict__ u8, unsigned char* __restrict__ grid,
                                                             int size) {
  pdl_enter();
  // one thread per 4 consecutive w of one (t, n, c, h) row: four strided reads, one 32-bit store per output
  const int W4 = W / 4;   // launcher guarantees W % 4 == 0
  const long long total = (long long)Tn * N * C * H * W4;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    long long r = o;
    const int w = (int)(r % W4) * 4; r /= W4;
    const int h = (int)(r % H); r /= H;
    const int c = (int)(r % C); r /= C;
    const int n = (int)(r % N);
    const int t = (int)(r / N);
    const long long src = ((((long long)t * N + n) * H + h) * W + w) * C + c;
    uint32_t packed = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float y = (ld<T>(v, src + (long long)i * C) / 2.f + 0.5f) * 255.f;
      const uint32_t q = (uint32_t)(y < 0.f ? 0.f : (y > 255.f ? 255.f : y));   // truncation, as astype(uint8)
      packed |= q << (8 * i);
    }
    if (u8) *reinterpret_cast<uint32_t*>(u8 + ((((long long)t * N + n) * C + c) * H + h) * W + w) = packed;
    if (grid && n < size * size) {
      const int gi = n / size, gj = n % size;
      *reinterpret_cast<uint32_t*>(grid + (((long long)t * C + c) * (size * H) + gi * H + h) * (long long)(size * W) + gj * W + w) = packed;
    }
  }
}

// =====================================================================================================
// Adam + WeightDecay, casts, RNG utilities, step state
// =====================================================================================================
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v,
                                                   __nv_bfloat16* __restrict__ pb, long long n, float alpha,
                                                   float beta1, float beta2, float eps, float wd, float gscale,
                                                   const int* __restrict__ t_ptr) {
  pdl_enter();
  __shared__ float lr_s;
  if (threadIdx.x == 0) {
    double t = (double)(*t_ptr);
    double fix1 = 1.0 - pow((double)beta1, t), fix2 = 1.0 - pow((double)beta2, t);
    lr_s = (float)((double)alpha * sqrt(fix2) / fix1);
  }
  __syncthreads();
  const float lr = lr_s, omb1 = 1.f - beta1, omb2 = 1.f - beta2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float pi = p[i];
    float gi = gscale * g[i] + wd * pi;
    float mi = m[i], vi = v[i];
    mi += omb1 * (gi - mi);
    vi += omb2 * (gi * gi - vi);
    pi -= lr * mi / (sqrtf(vi) + eps);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi;
    if (pb) pb[i] = __float2bfloat16_rn(pi);
  }
}
// dst[row][0..Cp) = src[row][0..C) followed by zeros: the generator's 60-channel latent padded to the 64 the tcgen05 path wants
template <typename T>
__global__ void __launch_bounds__(256) pad_channels_kernel(const T* __restrict__ src, T* __restrict__ dst, long long rows, int C, int Cp) {
  pdl_enter();
  const long long total = rows * Cp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / Cp;
    const int c = (int)(i - r * Cp);
    dst[i] = c < C ? src[r * C + c] : T(0.f);
  }
}
__global__ void cast_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ d, long long n) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    d[i] = __float2bfloat16_rn(s[i]);
}
// 8 elements per thread (n % 8 == 0, 16-byte aligned): the gradient arenas on their way into / out of the bf16 all-reduce
__global__ void __launch_bounds__(256) cast8_f32_bf16_kernel(const float4* __restrict__ s, uint4* __restrict__ d, long long n8) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = s[2 * i], b = s[2 * i + 1];
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
    h[0] = __floats2bfloat162_rn(a.x, a.y); h[1] = __floats2bfloat162_rn(a.z, a.w);
    h[2] = __floats2bfloat162_rn(b.x, b.y); h[3] = __floats2bfloat162_rn(b.z, b.w);
    d[i] = u;
  }
}
__global__ void __launch_bounds__(256) cast8_bf16_f32_kernel(const uint4* __restrict__ s, float4* __restrict__ d, long long n8) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const uint4 u = s[i];
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]), c = __bfloat1622float2(h[2]), e = __bfloat1622float2(h[3]);
    d[2 * i] = make_float4(a.x, a.y, b.x, b.y);
    d[2 * i + 1] = make_float4(c.x, c.y, e.x, e.y);
  }
}
__global__ void state_init_kernel(StepState* s, unsigned long long seed) {
  pdl_enter();
  s->seed_lo = (uint32_t)seed; s->seed_hi = (uint32_t)(seed >> 32);
  s->step = 0; s->frame_t = 0; s->adam_t = 0; s->r0 = s->r1 = s->r2 = 0;
}
__global__ void state_advance_kernel(StepState* s, int T) {
  pdl_enter();
  s->step += 1;
  s->adam_t += 1;
  uint4 r = philox4x32(make_uint4(0, 0, 0x7fffffff, s->step), make_uint2(s->seed_lo, s->seed_hi));
  s->frame_t = T > 0 ? r.x % (uint32_t)T : 0;
}
__global__ void randn_kernel(float* out, long long n, float sigma, const StepState* rng, int call_id) {
  pdl_enter();
  long long n4 = (n + 3) / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float z[4];
    philox_normal4(rng, call_id, (unsigned long long)i, z);
    for (int j = 0; j < 4; ++j)
      if (i * 4 + j < n) out[i * 4 + j] = sigma * z[j];
  }
}
__global__ void randint_kernel(int* out, long long n, int high, const StepState* rng, int call_id) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    uint4 r = philox4x32(make_uint4((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)call_id, rng->step),
                         make_uint2(rng->seed_lo, rng->seed_hi));
    out[i] = (int)(r.x % (uint32_t)high);
  }
}

static int grid_for(long long work, int threads = 256) {
  long long b = (work + threads - 1) / threads;
  long long cap = (long long)num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace mcg

using namespace mcg;

static bool dtype_ok(int d) { return d == MCG_F32 || d == MCG_BF16; }

// runtime dtype -> compile-time type tags
template <typename F> static void dispatch1(int d, F&& f) {
  if (d == MCG_F32) f(float{});
  else f(__nv_bfloat16{});
}
template <typename F> static void dispatch2(int d0, int d1, F&& f) {
  dispatch1(d0, [&](auto a) { dispatch1(d1, [&](auto b) { f(a, b); }); });
}

extern "C" {

int mcg_version(void) { return MCG_VERSION; }
const char* mcg_last_error(void) { return g_err; }
long long mcg_launch_count(void) { return g_launches.load(); }

size_t mcg_colreduce_workspace_bytes(long long M, int C) {
  (void)M;
  return (size_t)kRedSlots * 2 * (size_t)(C > 0 ? C : 1) * sizeof(float) + 256;
}

int mcg_bn_stats(const void* y, long long M, int C, int dtype, const float* gamma, const float* beta, float eps,
                 float decay, float* mean, float* invstd, float* scale, float* shift, float* avg_mean, float* avg_var,
                 void* workspace, size_t workspace_bytes, void* stream) {
  if (!y || !mean || !invstd) MCG_FAIL(MCG_ERR_SHAPE, "mcg_bn_stats: null pointer");
  return launch_colreduce(StatsF{}, StatsFinal{(double)M, gamma, beta, eps, decay, mean, invstd, scale, shift, avg_mean, avg_var},
                          y, nullptr, M, C, dtype, workspace, workspace_bytes, as_stream(stream), "mcg_bn_stats");
}

int mcg_colsum(const void* g, long long M, int C, int dtype, float* out, int accumulate, void* workspace,
               size_t workspace_bytes, void* stream) {
  if (!g || !out) MCG_FAIL(MCG_ERR_SHAPE, "mcg_colsum: null pointer");
  return launch_colreduce(SumF{}, Sum2Final{out, nullptr, accumulate, nullptr, nullptr}, g, nullptr, M, C, dtype, workspace,
                          workspace_bytes, as_stream(stream), "mcg_colsum");
}

int mcg_act_bn_bwd_reduce(const void* g, const void* y, long long M, int C, int dtype, const float* mean,
                          const float* invstd, const float* scale, const float* shift, int act, float slope,
                          float* dgamma, float* dbeta, float* acc_dgamma, float* acc_dbeta, void* workspace,
                          size_t workspace_bytes, void* stream) {
  if (!g || !y || !mean || !invstd || !scale || !shift || !dgamma || !dbeta)
    MCG_FAIL(MCG_ERR_SHAPE, "mcg_act_bn_bwd_reduce: null pointer");
  // first sum = sum g' -> dbeta ; second = sum g' xhat -> dgamma; acc_* are the parameter gradients, which accumulate
  // across the real and fake calls of one pass as in Chainer.
  const Sum2Final fin{dbeta, dgamma, 0, acc_dbeta, acc_dgamma};
#define MCG_RED(A) return launch_colreduce(BnBwdF<A>{mean, invstd, scale, shift, act, slope}, fin, g, y, M, C, dtype, workspace, \
                                           workspace_bytes, as_stream(stream), "mcg_act_bn_bwd_reduce")
  if (C % 8) MCG_RED(-1);
  else if (act == MCG_ACT_RELU) MCG_RED(MCG_ACT_RELU);
  else if (act == MCG_ACT_LRELU) MCG_RED(MCG_ACT_LRELU);
  else MCG_RED(-1);
#undef MCG_RED
}

int mcg_affine_act_noise(const void* y, long long M, int C, long long P, int dtype, const float* scale,
                         const float* shift, int act, float slope, float sigma, const float* noise, long long ns_n,
                         long long ns_c, long long ns_p, const void* rng_state, int call_id, void* out, int out_dtype,
                         void* stream) {
  if (!y || !out || M <= 0 || C <= 0 || P <= 0) MCG_FAIL(MCG_ERR_SHAPE, "mcg_affine_act_noise: bad arguments");
  if (!dtype_ok(dtype) || !dtype_ok(out_dtype)) MCG_FAIL(MCG_ERR_UNSUPPORTED, "mcg_affine_act_noise: dtype");
  if ((scale == nullptr) != (shift == nullptr)) MCG_FAIL(MCG_ERR_SHAPE, "mcg_affine_act_noise: scale/shift mismatch");
  cudaStream_t st = as_stream(stream);
  const StepState* rng = (const StepState*)rng_state;
  if (!scale && !noise && (!rng || sigma == 0.f) && C % 8 && (M * C) % 8 == 0) {
    // a bare activation (the generator's tanh on 3 channels): the channel structure is irrelevant, treat the tensor as
    // rows of 8 values so the 16-byte vector path is used
    M = M * C / 8;
    C = 8;
    P = 1;
  }
  dispatch2(dtype, out_dtype, [&](auto ti, auto to) {
    using TI = decltype(ti);
    using TO = decltype(to);
    if (C % 8 == 0 && 256 % (C / 8) == 0) {  // a thread keeps its channel group: per-channel constants in registers
      const bool has_noise = noise || (rng && sigma != 0.f);
#define MCG_AFF(A)                                                                                                  \
  do {                                                                                                              \
    if (has_noise)                                                                                                  \
      pdl(affine_act_noise_kernel<TI, TO, 8, true, A, true>, grid_for(M * (C / 8)), 256, 0, st)(                   \
          (const TI*)y, M, C, P, scale, shift, act, slope, sigma, noise, ns_n, ns_c, ns_p, rng, call_id, (TO*)out); \
    else                                                                                                            \
      pdl(affine_act_noise_kernel<TI, TO, 8, true, A, false>, grid_for(M * (C / 8)), 256, 0, st)(                  \
          (const TI*)y, M, C, P, scale, shift, act, slope, sigma, noise, ns_n, ns_c, ns_p, rng, call_id, (TO*)out); \
  } while (0)
      if (act == MCG_ACT_RELU) MCG_AFF(MCG_ACT_RELU);
      else if (act == MCG_ACT_LRELU) MCG_AFF(MCG_ACT_LRELU);
      else if (act == MCG_ACT_TANH) MCG_AFF(MCG_ACT_TANH);
      else MCG_AFF(MCG_ACT_NONE);
#undef MCG_AFF
    } else if (C % 8 == 0)
      pdl(affine_act_noise_kernel<TI, TO, 8, false>, grid_for(M * (C / 8)), 256, 0, st)(
          (const TI*)y, M, C, P, scale, shift, act, slope, sigma, noise, ns_n, ns_c, ns_p, rng, call_id, (TO*)out);
    else
      pdl(affine_act_noise_kernel<TI, TO, 1, false>, grid_for(M * C), 256, 0, st)(
          (const TI*)y, M, C, P, scale, shift, act, slope, sigma, noise, ns_n, ns_c, ns_p, rng, call_id, (TO*)out);
  });
  MCG_CHECK_LAUNCH("mcg_affine_act_noise");
  return 0;
}

int mcg_pack_video(const void* src, int src_dtype, int N, int C, int T, int H, int W, long long s_n, long long s_c,
                   long long s_t, long long s_h, long long s_w, const int* frame_ptr, float sigma, const float* noise,
                   long long ns_n, long long ns_c, long long ns_p, const void* rng_state, int call_id, void* out,
                   int out_dtype, void* stream) {
  if (!src || !out || N <= 0 || C <= 0 || T <= 0 || H <= 0 || W <= 0)
    MCG_FAIL(MCG_ERR_SHAPE, "mcg_pack_video: bad arguments");
  if ((!dtype_ok(src_dtype) && src_dtype != MCG_U8) || !dtype_ok(out_dtype)) MCG_FAIL(MCG_ERR_UNSUPPORTED, "mcg_pack_video: dtype");
  cudaStream_t st = as_stream(stream);
  long long total = (long long)N * (frame_ptr ? 1 : T) * H * W;
  if (src_dtype == MCG_U8) {   // pre-decoded pixels: normalised (v - 128) / 128 as they are read (datasets.py:91)
    dispatch1(out_dtype, [&](auto to) {
      using TO = decltype(to);
      pdl(pack_video_kernel<unsigned char, TO>, grid_for(total), 256, 0, st)((const unsigned char*)src, N, C, T, H, W, s_n, s_c, s_t,
                                                                          s_h, s_w, frame_ptr, sigma, noise, ns_n, ns_c, ns_p,
                                                                          (const StepState*)rng_state, call_id, (TO*)out);
    });
    MCG_CHECK_LAUNCH("mcg_pack_video(u8)");
    return 0;
  }
  const bool plain = !frame_ptr && !noise && (sigma == 0.f || !rng_state);
  if (plain && C > 4) {   // pure layout / dtype conversion of a wide tensor
    dispatch2(src_dtype, out_dtype, [&](auto ti, auto to) {
      using TI = decltype(ti);
      using TO = decltype(to);
      pdl(pack_elem_kernel<TI, TO>, grid_for(total * C), 256, 0, st)((const TI*)src, N, C, T, H, W, s_n, s_c, s_t, s_h, s_w,
                                                                     (TO*)out);
    });
    MCG_CHECK_LAUNCH("mcg_pack_video(elem)");
    return 0;
  }
  dispatch2(src_dtype, out_dtype, [&](auto ti, auto to) {
    using TI = decltype(ti);
    using TO = decltype(to);
    pdl(pack_video_kernel<TI, TO>, grid_for(total), 256, 0, st)((const TI*)src, N, C, T, H, W, s_n, s_c, s_t, s_h, s_w,
                                                               frame_ptr, sigma, noise, ns_n, ns_c, ns_p,
                                                               (const StepState*)rng_state, call_id, (TO*)out);
  });
  MCG_CHECK_LAUNCH("mcg_pack_video");
  return 0;
}

int mcg_act_bn_bwd_apply(const void* g, const void* y, long long M, int C, int dtype, const float* mean,
                         const float* invstd, const float* gamma, const float* scale, const float* shift, int act,
                         float slope, int use_output, const float* dgamma, const float* dbeta, void* gy, int out_dtype,
                         void* stream) {
  if (!g || !y || !gy || M <= 0 || C <= 0) MCG_FAIL(MCG_ERR_SHAPE, "mcg_act_bn_bwd_apply: bad arguments");
  if (mean && (!invstd || !gamma || !scale || !shift || !dgamma || !dbeta))
    MCG_FAIL(MCG_ERR_SHAPE, "mcg_act_bn_bwd_apply: BN mode needs invstd/gamma/scale/shift/dgamma/dbeta");
  if (!dtype_ok(dtype) || !dtype_ok(out_dtype)) MCG_FAIL(MCG_ERR_UNSUPPORTED, "mcg_act_bn_bwd_apply: dtype");
  cudaStream_t st = as_stream(stream);
  float inv_m = 1.0f / (float)M;
  dispatch2(dtype, out_dtype, [&](auto ti, auto to) {
    using TI = decltype(ti);
    using TO = decltype(to);
    if (C % 8 == 0 && 256 % (C / 8) == 0) {
#define MCG_APP(A)                                                                                                \
  pdl(act_bn_bwd_apply_kernel<TI, TO, 8, true, A>, grid_for(M * (C / 8)), 256, 0, st)(                           \
      (const TI*)g, (const TI*)y, M, C, mean, invstd, gamma, scale, shift, act, slope, use_output, dgamma, dbeta, \
      inv_m, (TO*)gy)
      if (act == MCG_ACT_RELU) MCG_APP(MCG_ACT_RELU);
      else if (act == MCG_ACT_LRELU) MCG_APP(MCG_ACT_LRELU);
      else MCG_APP(-1);
#undef MCG_APP
    } else if (C % 8 == 0)
      pdl(act_bn_bwd_apply_kernel<TI, TO, 8, false>, grid_for(M * (C / 8)), 256, 0, st)(
          (const TI*)g, (const TI*)y, M, C, mean, invstd, gamma, scale, shift, act, slope, use_output, dgamma, dbeta,
          inv_m, (TO*)gy);
    else
      pdl(act_bn_bwd_apply_kernel<TI, TO, 1, false>, grid_for(M * C), 256, 0, st)(
          (const TI*)g, (const TI*)y, M, C, mean, invstd, gamma, scale, shift, act, slope, use_output, dgamma, dbeta,
          inv_m, (TO*)gy);
  });
  MCG_CHECK_LAUNCH("mcg_act_bn_bwd_apply");
  return 0;
}

int mcg_tanh_bwd_video(const void* gv, const void* gi, int g_dtype, const void* out_tn, int out_dtype, int N, int T,
                       int HW, int C, const int* frame_ptr, void* g_tn, int gout_dtype, void* stream) {
  if (!out_tn || !g_tn || (!gv && !gi) || N <= 0 || T <= 0 || HW <= 0 || C <= 0)
    MCG_FAIL(MCG_ERR_SHAPE, "mcg_tanh_bwd_video: bad arguments");
  if (!dtype_ok(g_dtype) || !dtype_ok(out_dtype) || !dtype_ok(gout_dtype))
    MCG_FAIL(MCG_ERR_UNSUPPORTED, "mcg_tanh_bwd_video: dtype");
  cudaStream_t st = as_stream(stream);
  long long total = (long long)T * N * HW * C;
  dispatch2(g_dtype, out_dtype, [&](auto tg, auto to) {
    using TG = decltype(tg);
    using TO = decltype(to);
    dispatch1(gout_dtype, [&](auto td) {
      using TD = decltype(td);
      pdl(tanh_bwd_video_kernel<TG, TO, TD>, grid_for(total), 256, 0, st)((const TG*)gv, (const TG*)gi, (const TO*)out_tn,
                                                                         N, T, HW, C, frame_ptr, (TD*)g_tn);
    });
  });
  MCG_CHECK_LAUNCH("mcg_tanh_bwd_video");
  return 0;
}

int mcg_video_to_uint8(const void* videos, int dtype, int T, int N, int C, int H, int W, unsigned char* u8,
                       unsigned char* grid, int size, void* stream) {
  if (!videos || (!u8 && !grid) || T <= 0 || N <= 0 || C <= 0 || H <= 0 || W <= 0 || (grid && size <= 0))
    MCG_FAIL(MCG_ERR_SHAPE, "mcg_video_to_uint8: bad arguments");
  if (W % 4) MCG_FAIL(MCG_ERR_UNSUPPORTED, "mcg_video_to_uint8: W must be a multiple of 4");
  if (!dtype_ok(dtype)) MCG_FAIL(MCG_ERR_UNSUPPORTED, "mcg_video_to_uint8: dtype");
  cudaStream_t st = as_stream(stream);
  const long long total = (long long)T * N * C * H * (W / 4);
  if (grid) {   // cells beyond N stay black
    cudaError_t e = cudaMemsetAsync(grid, 0, (size_t)T * C * size * H * size * W, st);
    if (e != cudaSuccess) MCG_FAIL((int)e, "mcg_video_to_uint8: memset: %s", cudaGetErrorString(e));
  }
  dispatch1(dtype, [&](auto ti) {
    using TI = decltype(ti);
    pdl(video_to_uint8_kernel<TI>, grid_for(total), 256, 0, st)((const TI*)videos, T, N, C, H, W, u8, grid, size);
  });
  MCG_CHECK_LAUNCH("mcg_video_to_uint8");
  return 0;
}

int mcg_adam_step(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, float alpha, float beta1,
                  float beta2, float eps, float wd, float grad_scale, const int* t_ptr, void* stream) {
  if (!p || !g || !m || !v || !t_ptr || n <= 0) MCG_FAIL(MCG_ERR_SHAPE, "mcg_adam_step: bad arguments");
  pdl(adam_kernel, grid_for(n), 256, 0, as_stream(stream))(p, g, m, v, (__nv_bfloat16*)p_bf16, n, alpha, beta1, beta2,
                                                          eps, wd, grad_scale, t_ptr);
  MCG_CHECK_LAUNCH("mcg_adam_step");
  return 0;
}

int mcg_fill_zero(void* p, size_t bytes, void* stream) {
  if (!p && bytes) MCG_FAIL(MCG_ERR_SHAPE, "mcg_fill_zero: null pointer");
  cudaError_t e = cudaMemsetAsync(p, 0, bytes, as_stream(stream));     // a memset node when captured: no kernel, no SM
  if (e != cudaSuccess) MCG_FAIL((int)e, "mcg_fill_zero: %s", cudaGetErrorString(e));
  return 0;
}

int mcg_pad_channels(const void* src, void* dst, long long rows, int C, int Cp, int dtype, void* stream) {
  if (!src || !dst || rows <= 0 || C <= 0 || Cp < C) MCG_FAIL(MCG_ERR_SHAPE, "mcg_pad_channels: bad arguments");
  const int grid = grid_for(rows * Cp);
  if (dtype == MCG_BF16) pdl(pad_channels_kernel<__nv_bfloat16>, grid, 256, 0, as_stream(stream))((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, rows, C, Cp);
  else if (dtype == MCG_F32) pdl(pad_channels_kernel<float>, grid, 256, 0, as_stream(stream))((const float*)src, (float*)dst, rows, C, Cp);
  else MCG_FAIL(MCG_ERR_UNSUPPORTED, "mcg_pad_channels: dtype %d", dtype);
  MCG_CHECK_LAUNCH("mcg_pad_channels");
  return 0;
}

int mcg_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream) {
  if (!src || !dst || n <= 0) MCG_FAIL(MCG_ERR_SHAPE, "mcg_cast_f32_to_bf16: bad arguments");
  if (n % 8 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0)
    pdl(cast8_f32_bf16_kernel, grid_for(n / 8), 256, 0, as_stream(stream))((const float4*)src, (uint4*)dst, n / 8);
  else
    pdl(cast_kernel, grid_for(n), 256, 0, as_stream(stream))(src, (__nv_bfloat16*)dst, n);
  MCG_CHECK_LAUNCH("mcg_cast_f32_to_bf16");
  return 0;
}

int mcg_cast_bf16_to_f32(const void* src, float* dst, long long n, void* stream) {
  if (!src || !dst || n <= 0 || n % 8 || (reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15))
    MCG_FAIL(MCG_ERR_SHAPE, "mcg_cast_bf16_to_f32: needs n %% 8 == 0 and 16-byte aligned buffers");
  pdl(cast8_bf16_f32_kernel, grid_for(n / 8), 256, 0, as_stream(stream))((const uint4*)src, (float4*)dst, n / 8);
  MCG_CHECK_LAUNCH("mcg_cast_bf16_to_f32");
  return 0;
}

int mcg_step_state_init(void* state, unsigned long long seed, void* stream) {
  if (!state) MCG_FAIL(MCG_ERR_SHAPE, "mcg_step_state_init: null state");
  pdl(state_init_kernel, 1, 1, 0, as_stream(stream))((StepState*)state, seed);
  MCG_CHECK_LAUNCH("mcg_step_state_init");
  return 0;
}
int mcg_step_advance(void* state, int T, void* stream) {
  if (!state) MCG_FAIL(MCG_ERR_SHAPE, "mcg_step_advance: null state");
  pdl(state_advance_kernel, 1, 1, 0, as_stream(stream))((StepState*)state, T);
  MCG_CHECK_LAUNCH("mcg_step_advance");
  return 0;
}
int mcg_randn(float* out, long long n, float sigma, const void* rng_state, int call_id, void* stream) {
  if (!out || !rng_state || n <= 0) MCG_FAIL(MCG_ERR_SHAPE, "mcg_randn: bad arguments");
  pdl(randn_kernel, grid_for((n + 3) / 4), 256, 0, as_stream(stream))(out, n, sigma, (const StepState*)rng_state, call_id);
  MCG_CHECK_LAUNCH("mcg_randn");
  return 0;
}
int mcg_randint(int* out, long long n, int high, const void* rng_state, int call_id, void* stream) {
  if (!out || !rng_state || n <= 0 || high <= 0) MCG_FAIL(MCG_ERR_SHAPE, "mcg_randint: bad arguments");
  pdl(randint_kernel, grid_for(n), 256, 0, as_stream(stream))(out, n, high, (const StepState*)rng_state, call_id);
  MCG_CHECK_LAUNCH("mcg_randint");
  return 0;
}

}  // extern "C"
