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
  pdl_enter();
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
// y[m][col] = bias[col] + sum_k x[m][k] * w[col][k]: a block owns RM rows x CPB columns, so every weight row is read
// once per RM rows and every x row once per CPB columns (the one-row version re-read all of w for each row: G.dc1's
// backward-data moved 1.1 GB through L2 for 0.55 GFLOP).
template <typename T, int RM, int CPB>
__global__ void __launch_bounds__(256) fullwin_fprop_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, void* __restrict__ y,
                                                            int out_bf16, int M, int Ncols, int K) {
  pdl_enter();
  __shared__ float red[8][RM * CPB];
  const int m0 = blockIdx.x * RM, c0 = blockIdx.y * CPB;
  float acc[RM][CPB];
#pragma unroll
  for (int r = 0; r < RM; ++r)
#pragma unroll
    for (int c = 0; c < CPB; ++c) acc[r][c] = 0.f;
  if ((K & 7) == 0) {
    for (int k = threadIdx.x * 8; k < K; k += 256 * 8) {
      float wv[CPB][8];
#pragma unroll
      for (int c = 0; c < CPB; ++c) {
        if (c0 + c < Ncols) ld8<float>(w + (long long)(c0 + c) * K + k, wv[c]);
        else
#pragma unroll
          for (int i = 0; i < 8; ++i) wv[c][i] = 0.f;
      }
#pragma unroll
      for (int r = 0; r < RM; ++r) {
        if (m0 + r >= M) break;
        float xv[8];
        ld8<T>(x + (long long)(m0 + r) * K + k, xv);
#pragma unroll
        for (int c = 0; c < CPB; ++c)
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[r][c] = fmaf(xv[i], wv[c][i], acc[r][c]);
      }
    }
  } else {
    for (int k = threadIdx.x; k < K; k += 256)
#pragma unroll
      for (int r = 0; r < RM; ++r) {
        if (m0 + r >= M) break;
        const float xv = ld<T>(x, (long long)(m0 + r) * K + k);
#pragma unroll
        for (int c = 0; c < CPB; ++c)
          if (c0 + c < Ncols) acc[r][c] = fmaf(xv, w[(long long)(c0 + c) * K + k], acc[r][c]);
      }
  }
#pragma unroll
  for (int r = 0; r < RM; ++r)
#pragma unroll
    for (int c = 0; c < CPB; ++c) {
      float v = acc[r][c];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][r * CPB + c] = v;
    }
  __syncthreads();
  if (threadIdx.x < RM * CPB) {
    const int r = threadIdx.x / CPB, c = threadIdx.x % CPB;
    if (m0 + r < M && c0 + c < Ncols) {
      float sres = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) sres += red[i][threadIdx.x];
      if (bias) sres += bias[c0 + c];
      const long long o = (long long)(m0 + r) * Ncols + c0 + c;
      if (out_bf16) reinterpret_cast<__nv_bfloat16*>(y)[o] = __float2bfloat16_rn(sres);
      else reinterpret_cast<float*>(y)[o] = sres;
    }
  }
}

// dx[m][k] = bias[k % Cin] + sum_co dy[m][co] * w[co][k]: a thread owns four consecutive k of RM rows, so each 16-byte
// weight vector is loaded once per RM rows (K % 4 == 0 is checked by the launcher; Cin % 4 == 0 keeps a quad inside one
// pixel).  The dy values are warp-uniform broadcasts.
template <typename T, int RM>
__global__ void __launch_bounds__(256) fullwin_dgrad_kernel(const T* __restrict__ dy, const float* __restrict__ w,
                                                            const float* __restrict__ bias, void* __restrict__ dx,
                                                            int out_bf16, int accumulate, int M, int Cout, int K, int Cin) {
  pdl_enter();
  const int kq = K / 4;
  const long long quads = (long long)((M + RM - 1) / RM) * kq;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (long long)gridDim.x * blockDim.x) {
    const int m0 = (int)(q / kq) * RM, k = (int)(q % kq) * 4;
    float a[RM][4];
    const float b0 = bias ? bias[k % Cin] : 0.f, b1 = bias ? bias[(k + 1) % Cin] : 0.f, b2 = bias ? bias[(k + 2) % Cin] : 0.f,
                b3 = bias ? bias[(k + 3) % Cin] : 0.f;
#pragma unroll
    for (int r = 0; r < RM; ++r) { a[r][0] = b0; a[r][1] = b1; a[r][2] = b2; a[r][3] = b3; }
    for (int co = 0; co < Cout; ++co) {
      const float4 wv = *reinterpret_cast<const float4*>(w + (long long)co * K + k);
#pragma unroll
      for (int r = 0; r < RM; ++r) {
        const float d = (m0 + r < M) ? ld<T>(dy, (long long)(m0 + r) * Cout + co) : 0.f;
        a[r][0] = fmaf(d, wv.x, a[r][0]); a[r][1] = fmaf(d, wv.y, a[r][1]);
        a[r][2] = fmaf(d, wv.z, a[r][2]); a[r][3] = fmaf(d, wv.w, a[r][3]);
      }
    }
#pragma unroll
    for (int r = 0; r < RM; ++r) {
      if (m0 + r >= M) break;
      const long long o = (long long)(m0 + r) * K + k;
      float a0 = a[r][0], a1 = a[r][1], a2 = a[r][2], a3 = a[r][3];
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
}

// dw[co][k] += sum_m dy[m][co] * x[m][k] for a full-window layer (x[m] is one contiguous row of K values): a thread owns
// four consecutive k for CB output channels, rows are split over blockIdx.z and meet in fp32 red.add (wgrad accumulates).
template <typename T, int CB>
__global__ void __launch_bounds__(256) fullwin_wgrad_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                            float* __restrict__ dw, int M, int Cout, int K, int rows_per_split) {
  pdl_enter();
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= K / 4) return;
  const int k = q * 4, co0 = blockIdx.y * CB;
  const int m_begin = blockIdx.z * rows_per_split;
  int m_end = m_begin + rows_per_split;
  if (m_end > M) m_end = M;
  float acc[CB][4];
#pragma unroll
  for (int c = 0; c < CB; ++c) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f;
  for (int m = m_begin; m < m_end; ++m) {
    const T* xr = x + (long long)m * K + k;
    const float x0 = ld<T>(xr, 0), x1 = ld<T>(xr, 1), x2 = ld<T>(xr, 2), x3 = ld<T>(xr, 3);
#pragma unroll
    for (int c = 0; c < CB; ++c) {
      const float d = (co0 + c < Cout) ? ld<T>(dy, (long long)m * Cout + co0 + c) : 0.f;
      acc[c][0] = fmaf(d, x0, acc[c][0]); acc[c][1] = fmaf(d, x1, acc[c][1]);
      acc[c][2] = fmaf(d, x2, acc[c][2]); acc[c][3] = fmaf(d, x3, acc[c][3]);
    }
  }
#pragma unroll
  for (int c = 0; c < CB; ++c) {
    if (co0 + c >= Cout) break;
    float* o = dw + (long long)(co0 + c) * K + k;
    atomicAdd(o, acc[c][0]); atomicAdd(o + 1, acc[c][1]); atomicAdd(o + 2, acc[c][2]); atomicAdd(o + 3, acc[c][3]);
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
      if (g.Cout >= 16) {   // wide output (G.dc1 backward-data): 8 rows x 4 columns per block
        dim3 grid((unsigned)((g.N + 7) / 8), (unsigned)((g.Cout + 3) / 4));
        if (dtype == MCG_F32) pdl(fullwin_fprop_kernel<float, 8, 4>, grid, 256, 0, st)((const float*)a, w, bias, out, ob, g.N, g.Cout, g.K);
        else pdl(fullwin_fprop_kernel<__nv_bfloat16, 8, 4>, grid, 256, 0, st)((const __nv_bfloat16*)a, w, bias, out, ob, g.N, g.Cout, g.K);
      } else {              // a handful of columns (the discriminators' last layer): one row per block keeps the grid full
        dim3 grid((unsigned)g.N, (unsigned)g.Cout);
        if (dtype == MCG_F32) pdl(fullwin_fprop_kernel<float, 1, 1>, grid, 256, 0, st)((const float*)a, w, bias, out, ob, g.N, g.Cout, g.K);
        else pdl(fullwin_fprop_kernel<__nv_bfloat16, 1, 1>, grid, 256, 0, st)((const __nv_bfloat16*)a, w, bias, out, ob, g.N, g.Cout, g.K);
      }
    } else {
      const int rm = g.N >= 256 ? 8 : 1;   // many rows (G.dc1 forward): share each weight vector between 8 of them
      long long total = (long long)((g.N + rm - 1) / rm) * (g.K / 4);
      long long nb = (total + 255) / 256;
      int blocks = (int)(nb < (long long)num_sms() * 16 ? nb : (long long)num_sms() * 16);
#define GO_FW(T, RM_) pdl(fullwin_dgrad_kernel<T, RM_>, blocks, 256, 0, st)((const T*)a, w, bias, out, ob, accumulate, g.N, g.Cout, g.K, g.Cin)
      if (dtype == MCG_F32) { if (rm == 8) GO_FW(float, 8); else GO_FW(float, 1); }
      else { if (rm == 8) GO_FW(__nv_bfloat16, 8); else GO_FW(__nv_bfloat16, 1); }
#undef GO_FW
    }
    MCG_CHECK_LAUNCH(who);
    return 0;
  }
  if (!no_fullwin && is_full_window(g) && mode == kWgrad && g.K % 4 == 0) {
    // a = dy (N, Cout), b_act = x (N, K)
    const int quads = g.K / 4;
    const int cb = g.Cout >= 8 ? 8 : 1;
    const int gy = (g.Cout + cb - 1) / cb;
    long long ctas = (long long)((quads + 255) / 256) * gy;
    int splits = (int)((4LL * num_sms() + ctas - 1) / ctas);
    if (splits > (g.N + 7) / 8) splits = (g.N + 7) / 8;   // at least 8 rows per split
    if (splits < 1) splits = 1;
    const int rps = (g.N + splits - 1) / splits;
    splits = (g.N + rps - 1) / rps;
    dim3 grid((unsigned)((quads + 255) / 256), (unsigned)gy, (unsigned)splits);
#define GO_WG(T, CB_) pdl(fullwin_wgrad_kernel<T, CB_>, grid, 256, 0, st)((const T*)a, (const T*)b_act, (float*)out, g.N, g.Cout, g.K, rps)
    if (dtype == MCG_F32) { if (cb == 8) GO_WG(float, 8); else GO_WG(float, 1); }
    else { if (cb == 8) GO_WG(__nv_bfloat16, 8); else GO_WG(__nv_bfloat16, 1); }
#undef GO_WG
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
  pdl(conv_simt_kernel<T, MODE>, grid, 256, 0, st)(g, (const T*)a, (const T*)b_act, w, bias, out, out_dtype == MCG_BF16, \
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
