// simt_conv.cu — fp32-accumulate implicit-GEMM convolution on the CUDA cores.
// This is the strict-tolerance (<= 1e-5) path and the path for layers the tcgen05 kernel does not take
// (3-channel image layers, the 1x1 -> 4x4 first deconvolution, the final dot-product layers).
// One kernel, three index maps (fprop / dgrad / wgrad) over channels-last activations and (Cout,kT,kH,kW,Cin) weights.
#include "common.cuh"
#include <stdlib.h>

namespace mcg {

enum { kFprop = 0, kDgrad = 1, kWgrad = 2 };
constexpr int TM = 64, TN = 64, TK = 16, kPad = 4;

struct Geo {
  int N, Cin, Cout, Ti, Hi, Wi, To, Ho, Wo, kT, kH, kW, sT, sH, sW, pT, pH, pW;
  int taps, K;  // taps = kT*kH*kW
};

// ---- operand fetchers: value of the implicit matrices at (row, k) -----------------------------------------------
// fprop  : A[m][k]  = x at output pixel m, k = (tap, ci)            B[k][col] = w[col][k]
// dgrad  : A[m][k]  = dy feeding input pixel m, k = (tap, co)        B[k][col] = w[co][tap][col]
// wgrad  : A[row][r]= dy[r][row] (row = co, r = output pixel)         B[r][col] = x at pixel r, col = (tap, ci)
template <typename T>
__device__ __forceinline__ float fetch_x(const Geo& g, const T* x, long long m, int k) {
  int tap = k / g.Cin, ci = k - tap * g.Cin;
  int kw = tap % g.kW, kh = (tap / g.kW) % g.kH, kt = tap / (g.kW * g.kH);
  int wo = (int)(m % g.Wo); long long r = m / g.Wo;
  int ho = (int)(r % g.Ho); r /= g.Ho;
  int to = (int)(r % g.To); int n = (int)(r / g.To);
  int ti = to * g.sT - g.pT + kt, hi = ho * g.sH - g.pH + kh, wi = wo * g.sW - g.pW + kw;
  if ((unsigned)ti >= (unsigned)g.Ti || (unsigned)hi >= (unsigned)g.Hi || (unsigned)wi >= (unsigned)g.Wi) return 0.f;
  return ld<T>(x, ((((long long)n * g.Ti + ti) * g.Hi + hi) * g.Wi + wi) * g.Cin + ci);
}
template <typename T>
__device__ __forceinline__ float fetch_dy_for_input(const Geo& g, const T* dy, long long m, int k) {
  int tap = k / g.Cout, co = k - tap * g.Cout;
  int kw = tap % g.kW, kh = (tap / g.kW) % g.kH, kt = tap / (g.kW * g.kH);
  int wi = (int)(m % g.Wi); long long r = m / g.Wi;
  int hi = (int)(r % g.Hi); r /= g.Hi;
  int ti = (int)(r % g.Ti); int n = (int)(r / g.Ti);
  int tt = ti + g.pT - kt, hh = hi + g.pH - kh, ww = wi + g.pW - kw;
  if (tt < 0 || hh < 0 || ww < 0) return 0.f;
  if (tt % g.sT || hh % g.sH || ww % g.sW) return 0.f;
  int to = tt / g.sT, ho = hh / g.sH, wo = ww / g.sW;
  if (to >= g.To || ho >= g.Ho || wo >= g.Wo) return 0.f;
  return ld<T>(dy, ((((long long)n * g.To + to) * g.Ho + ho) * g.Wo + wo) * g.Cout + co);
}

template <typename T, int MODE>
__global__ void __launch_bounds__(256) conv_simt_kernel(Geo g, const T* __restrict__ a_src, const T* __restrict__ b_act,
                                                        const float* __restrict__ w, const float* __restrict__ bias,
                                                        void* __restrict__ out, int out_bf16, int accumulate,
                                                        long long Mrows, int Ncols, long long Kred, long long k_per_split) {
  __shared__ float As[TK][TM + kPad];
  __shared__ float Bs[TK][TN + kPad];
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;
  const long long kbeg = (long long)blockIdx.z * k_per_split;
  long long kend = kbeg + k_per_split;
  if (kend > Kred) kend = Kred;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int tx = tid % 16, ty = tid / 16;

  for (long long k0 = kbeg; k0 < kend; k0 += TK) {
    // ---- A tile (TM x TK)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int i, kk;
      if (MODE == kWgrad) { i = tid % 64; kk = tid / 64 + 4 * j; }       // rows (co) contiguous in memory
      else { kk = tid % 16; i = tid / 16 + 16 * j; }                      // k (channels) contiguous in memory
      long long m = m0 + i, k = k0 + kk;
      float v = 0.f;
      if (m < Mrows && k < kend) {
        if (MODE == kFprop) v = fetch_x<T>(g, a_src, m, (int)k);
        else if (MODE == kDgrad) v = fetch_dy_for_input<T>(g, a_src, m, (int)k);
        else v = ld<T>(a_src, k * g.Cout + m);  // dy[r][co]
      }
      As[kk][i] = v;
    }
    // ---- B tile (TK x TN)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c, kk;
      if (MODE == kFprop) { kk = tid % 16; c = tid / 16 + 16 * j; }       // w[col][k]: k contiguous
      else { c = tid % 64; kk = tid / 64 + 4 * j; }                        // columns contiguous
      long long k = k0 + kk;
      int col = n0 + c;
      float v = 0.f;
      if (col < Ncols && k < kend) {
        if (MODE == kFprop) v = w[(long long)col * g.K + k];
        else if (MODE == kDgrad) {
          int tap = (int)(k / g.Cout), co = (int)(k - (long long)tap * g.Cout);
          v = w[((long long)co * g.taps + tap) * g.Cin + col];
        } else v = fetch_x<T>(g, b_act, k, col);
      }
      Bs[kk][c] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + ty * 4 + i;
    if (m >= Mrows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int col = n0 + tx * 4 + j;
      if (col >= Ncols) continue;
      float v = acc[i][j];
      long long o = m * Ncols + col;
      if (MODE == kWgrad) {
        atomicAdd(reinterpret_cast<float*>(out) + o, v);
      } else {
        if (bias) v += bias[col];
        if (out_bf16) {
          __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(out);
          if (accumulate) v += __bfloat162float(p[o]);
          p[o] = __float2bfloat16_rn(v);
        } else {
          float* p = reinterpret_cast<float*>(out);
          if (accumulate) v += p[o];
          p[o] = v;
        }
      }
    }
  }
}


// ---- "full-window" layers: the kernel covers the whole input and the output is 1x1(x1) — Di.dc5, Dv.dc5 and the
// generator's first deconvolution (net.py:44,137,178).  They are plain GEMMs over contiguous rows:
//   fprop : y[m][col] = bias[col] + sum_k x[m][k] * w[col][k]          (M = batch, K = taps*Cin)
//   dgrad : dx[m][k]  = bias[k % Cin] + sum_co dy[m][co] * w[co][k]
template <typename T>
__global__ void __launch_bounds__(256) fullwin_fprop_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, void* __restrict__ y,
                                                            int out_bf16, int M, int Ncols, int K, int cols_per_block) {
  __shared__ float red[8];
  const int m = blockIdx.x;
  const T* xr = x + (long long)m * K;
  for (int cc = 0; cc < cols_per_block; ++cc) {
    const int col = blockIdx.y * cols_per_block + cc;
    if (col >= Ncols) break;
    const float* wr = w + (long long)col * K;
    float acc = 0.f;
    if ((K & 7) == 0) {
      for (int k = threadIdx.x * 8; k < K; k += 256 * 8) {
        float xv[8], wv[8];
        ld8<T>(xr + k, xv);
        ld8<float>(wr + k, wv);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc = fmaf(xv[i], wv[i], acc);
      }
    } else {
      for (int k = threadIdx.x; k < K; k += 256) acc = fmaf(ld<T>(xr, k), wr[k], acc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int i = 0; i < 8; ++i) s += red[i];
      if (bias) s += bias[col];
      const long long o = (long long)m * Ncols + col;
      if (out_bf16) reinterpret_cast<__nv_bfloat16*>(y)[o] = __float2bfloat16_rn(s);
      else reinterpret_cast<float*>(y)[o] = s;
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void __launch_bounds__(256) fullwin_dgrad_kernel(const T* __restrict__ dy, const float* __restrict__ w,
                                                            const float* __restrict__ bias, void* __restrict__ dx,
                                                            int out_bf16, int accumulate, int M, int Cout, int K, int Cin) {
  // four consecutive k per thread (K % 4 == 0 is checked by the launcher; Cin % 4 == 0 keeps a quad inside one pixel)
  const long long quads = (long long)M * (K / 4);
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(q / (K / 4)), k = (int)(q % (K / 4)) * 4;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (bias) { a0 = bias[k % Cin]; a1 = bias[(k + 1) % Cin]; a2 = bias[(k + 2) % Cin]; a3 = bias[(k + 3) % Cin]; }
    const T* dr = dy + (long long)m * Cout;
    for (int co = 0; co < Cout; ++co) {
      const float d = ld<T>(dr, co);
      const float4 wv = *reinterpret_cast<const float4*>(w + (long long)co * K + k);
      a0 = fmaf(d, wv.x, a0); a1 = fmaf(d, wv.y, a1); a2 = fmaf(d, wv.z, a2); a3 = fmaf(d, wv.w, a3);
    }
    const long long o = (long long)m * K + k;
    if (out_bf16) {
      __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(dx) + o;
      if (accumulate) { a0 += __bfloat162float(p[0]); a1 += __bfloat162float(p[1]); a2 += __bfloat162float(p[2]); a3 += __bfloat162float(p[3]); }
      __nv_bfloat162 lo = __floats2bfloat162_rn(a0, a1), hi = __floats2bfloat162_rn(a2, a3);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&lo);
      u.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(p) = u;
    } else {
      float* p = reinterpret_cast<float*>(dx) + o;
      if (accumulate) { a0 += p[0]; a1 += p[1]; a2 += p[2]; a3 += p[3]; }
      *reinterpret_cast<float4*>(p) = make_float4(a0, a1, a2, a3);
    }
  }
}

static bool is_full_window(const Geo& g) {
  return g.To == 1 && g.Ho == 1 && g.Wo == 1 && g.pT == 0 && g.pH == 0 && g.pW == 0 && g.kT == g.Ti && g.kH == g.Hi &&
         g.kW == g.Wi;
}

static int make_geo(const mcg_conv_geom* c, Geo* g, const char* who) {
  if (!c) MCG_FAIL(MCG_ERR_SHAPE, "%s: null geometry", who);
  *g = Geo{c->N, c->Cin, c->Cout, c->Ti, c->Hi, c->Wi, c->To, c->Ho, c->Wo, c->kT, c->kH, c->kW,
           c->sT, c->sH, c->sW, c->pT, c->pH, c->pW, 0, 0};
  if (c->N <= 0 || c->Cin <= 0 || c->Cout <= 0 || c->kT <= 0 || c->kH <= 0 || c->kW <= 0 || c->sT <= 0 || c->sH <= 0 ||
      c->sW <= 0 || c->pT < 0 || c->pH < 0 || c->pW < 0)
    MCG_FAIL(MCG_ERR_SHAPE, "%s: non-positive dimension", who);
  auto osz = [](int i, int k, int s, int p) { return (i + 2 * p - k) / s + 1; };
  if (c->Ti + 2 * c->pT < c->kT || c->Hi + 2 * c->pH < c->kH || c->Wi + 2 * c->pW < c->kW)
    MCG_FAIL(MCG_ERR_SHAPE, "%s: kernel larger than padded input", who);
  if (osz(c->Ti, c->kT, c->sT, c->pT) != c->To || osz(c->Hi, c->kH, c->sH, c->pH) != c->Ho ||
      osz(c->Wi, c->kW, c->sW, c->pW) != c->Wo)
    MCG_FAIL(MCG_ERR_SHAPE, "%s: output extent (%d,%d,%d) inconsistent with input (%d,%d,%d)", who, c->To, c->Ho, c->Wo,
             c->Ti, c->Hi, c->Wi);
  g->taps = c->kT * c->kH * c->kW;
  g->K = g->taps * c->Cin;
  return 0;
}

int simt_conv(int mode, const mcg_conv_geom* c, const void* a, const void* b_act, const float* w, const float* bias,
              void* out, int dtype, int out_dtype, int accumulate, cudaStream_t st) {
  Geo g;
  const char* who = mode == kFprop ? "mcg_conv_fprop(simt)" : mode == kDgrad ? "mcg_conv_dgrad(simt)" : "mcg_conv_wgrad(simt)";
  if (int rc = make_geo(c, &g, who)) return rc;
  if (dtype != MCG_F32 && dtype != MCG_BF16) MCG_FAIL(MCG_ERR_UNSUPPORTED, "%s: dtype %d", who, dtype);
  static const bool no_fullwin = getenv("MCG_NO_FULLWIN") != nullptr;  // debugging aid
  if (!no_fullwin && is_full_window(g) && mode != kWgrad && (mode == kFprop || g.K % 4 == 0)) {
    const int ob = out_dtype == MCG_BF16;
    if (mode == kFprop) {
      int cpb = g.Cout >= 16 ? 4 : 1;
      dim3 grid((unsigned)g.N, (unsigned)((g.Cout + cpb - 1) / cpb));
      if (dtype == MCG_F32) fullwin_fprop_kernel<float><<<grid, 256, 0, st>>>((const float*)a, w, bias, out, ob, g.N, g.Cout, g.K, cpb);
      else fullwin_fprop_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)a, w, bias, out, ob, g.N, g.Cout, g.K, cpb);
    } else {
      long long total = (long long)g.N * (g.K / 4);
      long long nb = (total + 255) / 256;
      int blocks = (int)(nb < (long long)num_sms() * 16 ? nb : (long long)num_sms() * 16);
      if (dtype == MCG_F32) fullwin_dgrad_kernel<float><<<blocks, 256, 0, st>>>((const float*)a, w, bias, out, ob, accumulate, g.N, g.Cout, g.K, g.Cin);
      else fullwin_dgrad_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)a, w, bias, out, ob, accumulate, g.N, g.Cout, g.K, g.Cin);
    }
    MCG_CHECK_LAUNCH(who);
    return 0;
  }
  long long Mrows, Kred;
  int Ncols;
  long long Mo = (long long)g.N * g.To * g.Ho * g.Wo, Mi = (long long)g.N * g.Ti * g.Hi * g.Wi;
  if (mode == kFprop) { Mrows = Mo; Ncols = g.Cout; Kred = g.K; }
  else if (mode == kDgrad) { Mrows = Mi; Ncols = g.Cin; Kred = (long long)g.taps * g.Cout; }
  else { Mrows = g.Cout; Ncols = g.K; Kred = Mo; }
  long long gx = (Mrows + TM - 1) / TM;
  int gy = (Ncols + TN - 1) / TN;
  int splits = 1;
  if (mode == kWgrad) {  // split the pixel reduction so the grid fills the machine
    long long tiles = gx * gy;
    long long want = (4LL * num_sms() + tiles - 1) / tiles;
    long long maxs = (Kred + 255) / 256;
    splits = (int)(want < maxs ? want : maxs);
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
  }
  long long kps = (Kred + splits - 1) / splits;
  kps = (kps + TK - 1) / TK * TK;
  splits = (int)((Kred + kps - 1) / kps);
  if (gx > 2147483647LL || gy > 65535) MCG_FAIL(MCG_ERR_UNSUPPORTED, "%s: grid too large", who);
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)splits);
#define GO(T, MODE)                                                                                             \
  conv_simt_kernel<T, MODE><<<grid, 256, 0, st>>>(g, (const T*)a, (const T*)b_act, w, bias, out, out_dtype == MCG_BF16, \
                                                  accumulate, Mrows, Ncols, Kred, kps)
  if (dtype == MCG_F32) {
    if (mode == kFprop) GO(float, kFprop); else if (mode == kDgrad) GO(float, kDgrad); else GO(float, kWgrad);
  } else {
    if (mode == kFprop) GO(__nv_bfloat16, kFprop); else if (mode == kDgrad) GO(__nv_bfloat16, kDgrad); else GO(__nv_bfloat16, kWgrad);
  }
#undef GO
  MCG_CHECK_LAUNCH(who);
  return 0;
}

}  // namespace mcg
