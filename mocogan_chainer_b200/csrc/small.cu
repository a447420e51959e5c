// small.cu — latency-bound pieces of the step: the motion-code GRU recurrence (one persistent warp per sample,
// hidden state in registers/shuffles and weights in shared memory across all T steps) and the GAN losses.
#include "common.cuh"

namespace mcg {

struct GruPtrs { const float* p[12]; };
struct GruGrads { float* p[12]; };
// order: 0 W_r.W 1 W_r.b 2 U_r.W 3 U_r.b 4 W_z.W 5 W_z.b 6 U_z.W 7 U_z.b 8 W.W 9 W.b 10 U.W 11 U.b
constexpr int kGruWarps = 4;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ void gru_load_weights(const GruPtrs& P, int H, int I, float* sm, float** W) {
  // layout in smem: the 12 arrays back to back
  int sizes[12] = {H * I, H, H * H, H, H * I, H, H * H, H, H * I, H, H * H, H};
  int off = 0;
  for (int a = 0; a < 12; ++a) {
    W[a] = sm + off;
    for (int i = threadIdx.x; i < sizes[a]; i += blockDim.x) sm[off + i] = P.p[a][i];
    off += sizes[a];
  }
}

__global__ void __launch_bounds__(kGruWarps * 32) gru_forward_kernel(GruPtrs P, const int* __restrict__ labels, int L,
                                                                     const float* __restrict__ h0,
                                                                     const float* __restrict__ eps,
                                                                     const float* __restrict__ zc, int T, int N, int H,
                                                                     int Zc, float* __restrict__ Z,
                                                                     float* __restrict__ cache) {
  pdl_enter();
  extern __shared__ float sm[];
  float* W[12];
  const int I = L + H;
  gru_load_weights(P, H, I, sm, W);
  __syncthreads();
  const int warp = threadIdx.x / 32, j = threadIdx.x % 32;
  const int n = blockIdx.x * kGruWarps + warp;
  if (n >= N) return;
  const bool on = j < H;
  const int jj = on ? j : 0;
  const int label = (labels && L > 0) ? labels[n] : -1;
  float h = on ? h0[(long long)n * H + j] : 0.f;
  const int ZW = Zc + H;
  for (int t = 0; t < T; ++t) {
    float e = on ? eps[((long long)t * N + n) * H + j] : 0.f;
    float ar = W[1][jj] + W[3][jj], az = W[5][jj] + W[7][jj], ah = W[9][jj] + W[11][jj];
    if (label >= 0) {
      ar += W[0][jj * I + label];
      az += W[4][jj * I + label];
      ah += W[8][jj * I + label];
    }
    for (int k = 0; k < H; ++k) {
      float ek = __shfl_sync(0xffffffffu, e, k), hk = __shfl_sync(0xffffffffu, h, k);
      ar += W[0][jj * I + L + k] * ek + W[2][jj * H + k] * hk;
      az += W[4][jj * I + L + k] * ek + W[6][jj * H + k] * hk;
      ah += W[8][jj * I + L + k] * ek;
    }
    float r = sigmoidf_(ar), z = sigmoidf_(az);
    float rh = r * h;
    for (int k = 0; k < H; ++k) ah += W[10][jj * H + k] * __shfl_sync(0xffffffffu, rh, k);
    float hb = tanhf(ah);
    float hn = z * hb + (1.f - z) * h;
    long long row = (long long)t * N + n;
    if (on) {
      float* c = cache + (row * 4) * H;
      c[j] = r;
      c[H + j] = z;
      c[2 * H + j] = hb;
      c[3 * H + j] = h;
      Z[row * ZW + Zc + j] = hn;
    }
    for (int c = j; c < Zc; c += 32) Z[row * ZW + c] = zc[(long long)n * Zc + c];
    h = hn;
  }
}

// Backward through the T steps.  Lane j owns row j of every weight matrix; its parameter gradients are accumulated in
// REGISTERS over all T steps (H <= kGruMaxH) and reach global memory with one red.add per element at the end — the
// per-step shared-memory float atomics of the first version (a CAS loop each, four warps on the same addresses) were
// 90 % of this kernel's time.
constexpr int kGruMaxH = 16;
__global__ void __launch_bounds__(kGruWarps * 32) gru_backward_kernel(GruPtrs P, GruGrads G, const int* __restrict__ labels,
                                                                      int L, const float* __restrict__ eps,
                                                                      const float* __restrict__ cache,
                                                                      const float* __restrict__ gz, int T, int N, int H,
                                                                      int Zc) {
  pdl_enter();
  extern __shared__ float sm[];
  float* W[12];
  const int I = L + H;
  gru_load_weights(P, H, I, sm, W);
  __syncthreads();
  const int warp = threadIdx.x / 32, j = threadIdx.x % 32;
  const int n = blockIdx.x * kGruWarps + warp;
  if (n >= N) return;
  const bool on = j < H;
  const int jj = on ? j : 0;
  const int label = (labels && L > 0) ? labels[n] : -1;
  const int ZW = Zc + H;
  float aWr[kGruMaxH], aWz[kGruMaxH], aW[kGruMaxH], aUr[kGruMaxH], aUz[kGruMaxH], aU[kGruMaxH];
#pragma unroll
  for (int k = 0; k < kGruMaxH; ++k) aWr[k] = aWz[k] = aW[k] = aUr[k] = aUz[k] = aU[k] = 0.f;
  float sr = 0.f, sz = 0.f, sa = 0.f;   // sums over t of grp, gzp, ga: the bias and label-column gradients
  float gh = 0.f;
  for (int t = T - 1; t >= 0; --t) {
    long long row = (long long)t * N + n;
    const float* c = cache + (row * 4) * H;
    float r = on ? c[j] : 0.f, z = on ? c[H + j] : 0.f, hb = on ? c[2 * H + j] : 0.f, hp = on ? c[3 * H + j] : 0.f;
    float e = on ? eps[row * H + j] : 0.f;
    if (on) gh += gz[row * ZW + Zc + j];
    float gzg = gh * (hb - hp), ghb = gh * z, ghp = gh * (1.f - z);
    float ga = ghb * (1.f - hb * hb);
    float grh = 0.f;
    for (int q = 0; q < H; ++q) grh += __shfl_sync(0xffffffffu, ga, q) * W[10][q * H + jj];
    float gr = grh * hp;
    ghp += grh * r;
    float gzp = gzg * z * (1.f - z), grp = gr * r * (1.f - r);
    float back = 0.f;
    for (int q = 0; q < H; ++q)
      back += __shfl_sync(0xffffffffu, gzp, q) * W[6][q * H + jj] + __shfl_sync(0xffffffffu, grp, q) * W[2][q * H + jj];
    ghp += back;
    float rh = r * hp;
#pragma unroll
    for (int k = 0; k < kGruMaxH; ++k) {
      if (k < H) {   // warp-uniform
        float ek = __shfl_sync(0xffffffffu, e, k), hk = __shfl_sync(0xffffffffu, hp, k), rhk = __shfl_sync(0xffffffffu, rh, k);
        aWr[k] = fmaf(grp, ek, aWr[k]); aWz[k] = fmaf(gzp, ek, aWz[k]); aW[k] = fmaf(ga, ek, aW[k]);
        aUr[k] = fmaf(grp, hk, aUr[k]); aUz[k] = fmaf(gzp, hk, aUz[k]); aU[k] = fmaf(ga, rhk, aU[k]);
      }
    }
    sr += grp; sz += gzp; sa += ga;
    gh = ghp;
  }
  if (on) {
#pragma unroll
    for (int k = 0; k < kGruMaxH; ++k) {
      if (k < H) {
        atomicAdd(&G.p[0][j * I + L + k], aWr[k]);
        atomicAdd(&G.p[4][j * I + L + k], aWz[k]);
        atomicAdd(&G.p[8][j * I + L + k], aW[k]);
        atomicAdd(&G.p[2][j * H + k], aUr[k]);
        atomicAdd(&G.p[6][j * H + k], aUz[k]);
        atomicAdd(&G.p[10][j * H + k], aU[k]);
      }
    }
    if (label >= 0) {
      atomicAdd(&G.p[0][j * I + label], sr);
      atomicAdd(&G.p[4][j * I + label], sz);
      atomicAdd(&G.p[8][j * I + label], sa);
    }
    atomicAdd(&G.p[1][j], sr); atomicAdd(&G.p[3][j], sr);
    atomicAdd(&G.p[5][j], sz); atomicAdd(&G.p[7][j], sz);
    atomicAdd(&G.p[9][j], sa); atomicAdd(&G.p[11][j], sa);
  }
}

// ------------------------------------------------------------------------------------------ losses
__device__ __forceinline__ float softplusf_(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }

__device__ float block_sum(float v, float* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w];
  __syncthreads();
  return s;
}

// CE over columns 1..C-1 of row n with target t; adds gradient/N into g row; returns -log p[t]
__device__ float ce_row(const float* y, int C, int t, float inv_n, float* g) {
  float mx = -INFINITY;
  for (int c = 1; c < C; ++c) mx = fmaxf(mx, y[c]);
  float s = 0.f;
  for (int c = 1; c < C; ++c) s += expf(y[c] - mx);
  float lse = mx + logf(s);
  for (int c = 1; c < C; ++c) g[c] += (expf(y[c] - lse) - ((c - 1) == t ? 1.f : 0.f)) * inv_n;
  return lse - y[1 + t];
}

__global__ void __launch_bounds__(128) loss_dis_kernel(const float* y_real, const float* y_fake, const int* t_real,
                                                       const int* t_fake, int N, int C, int use_ce, float* loss,
                                                       float* gy_real, float* gy_fake) {
  pdl_enter();
  __shared__ float sh[4];
  const float inv_n = 1.f / (float)N;
  for (int i = threadIdx.x; i < N * C; i += blockDim.x) { gy_real[i] = 0.f; gy_fake[i] = 0.f; }
  __syncthreads();
  float part = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {  // `[:1]`: sample 0 only, all C columns (updater.py:25-26)
    part += softplusf_(-y_real[c]) * inv_n + softplusf_(y_fake[c]) * inv_n;
    gy_real[c] = -sigmoidf_(-y_real[c]) * inv_n;
    gy_fake[c] = sigmoidf_(y_fake[c]) * inv_n;
  }
  __syncthreads();
  if (use_ce) {
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      part += ce_row(y_real + (long long)n * C, C, t_real[n], inv_n, gy_real + (long long)n * C) * inv_n;
      part += ce_row(y_fake + (long long)n * C, C, t_fake[n], inv_n, gy_fake + (long long)n * C) * inv_n;
    }
  }
  float tot = block_sum(part, sh);
  if (threadIdx.x == 0) *loss = tot;
}

__global__ void __launch_bounds__(128) loss_gen_kernel(const float* y_i, const float* y_v, const int* t_fake, int N, int C,
                                                       int use_ce, float* loss, float* gy_i, float* gy_v) {
  pdl_enter();
  __shared__ float sh[4];
  const float inv_n = 1.f / (float)N;
  for (int i = threadIdx.x; i < N * C; i += blockDim.x) { gy_i[i] = 0.f; gy_v[i] = 0.f; }
  __syncthreads();
  float part = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float a = y_i[(long long)n * C], b = y_v[(long long)n * C];
    part += (softplusf_(-a) + softplusf_(-b)) * inv_n;
    gy_i[(long long)n * C] = -sigmoidf_(-a) * inv_n;
    gy_v[(long long)n * C] = -sigmoidf_(-b) * inv_n;
    if (use_ce) {
      part += ce_row(y_i + (long long)n * C, C, t_fake[n], inv_n, gy_i + (long long)n * C) * inv_n;
      part += ce_row(y_v + (long long)n * C, C, t_fake[n], inv_n, gy_v + (long long)n * C) * inv_n;
    }
  }
  float tot = block_sum(part, sh);
  if (threadIdx.x == 0) *loss = tot;
}

__global__ void int_add_kernel(int* p, int d) {
  pdl_enter(); *p += d; }

}  // namespace mcg

using namespace mcg;

extern "C" {

static int gru_check(int T, int N, int H, int L, int Zc, const char* who) {
  if (T <= 0 || N <= 0 || H <= 0 || H > 32 || L < 0 || L > 64 || Zc < 0)
    MCG_FAIL(MCG_ERR_UNSUPPORTED, "%s: need 0<H<=32, 0<=L<=64 (got T=%d N=%d H=%d L=%d Zc=%d)", who, T, N, H, L, Zc);
  return 0;
}

int mcg_gru_forward(const float* const* params_host, const int* labels, int L, const float* h0, const float* eps,
                    const float* zc, int T, int N, int H, int Zc, float* z, float* cache, void* stream) {
  if (int rc = gru_check(T, N, H, L, Zc, "mcg_gru_forward")) return rc;
  if (!params_host || !h0 || !eps || !z || !cache || (Zc > 0 && !zc) || (L > 0 && !labels))
    MCG_FAIL(MCG_ERR_SHAPE, "mcg_gru_forward: null pointer");
  GruPtrs P;
  for (int i = 0; i < 12; ++i) P.p[i] = params_host[i];
  size_t smem = (size_t)(3 * (H * (L + H) + H * H) + 6 * H) * sizeof(float);
  pdl(gru_forward_kernel, (N + kGruWarps - 1) / kGruWarps, kGruWarps * 32, smem, as_stream(stream))(
      P, labels, L, h0, eps, zc, T, N, H, Zc, z, cache);
  MCG_CHECK_LAUNCH("mcg_gru_forward");
  return 0;
}

int mcg_gru_backward(const float* const* params_host, float* const* grads_host, const int* labels, int L,
                     const float* eps, const float* cache, const float* gz, int T, int N, int H, int Zc, void* stream) {
  if (int rc = gru_check(T, N, H, L, Zc, "mcg_gru_backward")) return rc;
  if (!params_host || !grads_host || !eps || !cache || !gz || (L > 0 && !labels))
    MCG_FAIL(MCG_ERR_SHAPE, "mcg_gru_backward: null pointer");
  GruPtrs P;
  GruGrads G;
  for (int i = 0; i < 12; ++i) { P.p[i] = params_host[i]; G.p[i] = grads_host[i]; }
  if (H > kGruMaxH) MCG_FAIL(MCG_ERR_UNSUPPORTED, "mcg_gru_backward: H = %d > %d", H, kGruMaxH);
  size_t smem = (size_t)(3 * (H * (L + H) + H * H) + 6 * H) * sizeof(float);
  pdl(gru_backward_kernel, (N + kGruWarps - 1) / kGruWarps, kGruWarps * 32, smem, as_stream(stream))(
      P, G, labels, L, eps, cache, gz, T, N, H, Zc);
  MCG_CHECK_LAUNCH("mcg_gru_backward");
  return 0;
}

int mcg_loss_dis(const float* y_real, const float* y_fake, const int* t_real, const int* t_fake, int N, int C, int use_ce,
                 float* loss, float* gy_real, float* gy_fake, void* stream) {
  if (!y_real || !y_fake || !loss || !gy_real || !gy_fake || N <= 0 || C <= 0)
    MCG_FAIL(MCG_ERR_SHAPE, "mcg_loss_dis: bad arguments");
  if (use_ce && (!t_real || !t_fake || C < 2)) MCG_FAIL(MCG_ERR_SHAPE, "mcg_loss_dis: CE needs labels and C >= 2");
  pdl(loss_dis_kernel, 1, 128, 0, as_stream(stream))(y_real, y_fake, t_real, t_fake, N, C, use_ce, loss, gy_real, gy_fake);
  MCG_CHECK_LAUNCH("mcg_loss_dis");
  return 0;
}

int mcg_loss_gen(const float* y_i, const float* y_v, const int* t_fake, int N, int C, int use_ce, float* loss,
                 float* gy_i, float* gy_v, void* stream) {
  if (!y_i || !y_v || !loss || !gy_i || !gy_v || N <= 0 || C <= 0) MCG_FAIL(MCG_ERR_SHAPE, "mcg_loss_gen: bad arguments");
  if (use_ce && (!t_fake || C < 2)) MCG_FAIL(MCG_ERR_SHAPE, "mcg_loss_gen: CE needs labels and C >= 2");
  pdl(loss_gen_kernel, 1, 128, 0, as_stream(stream))(y_i, y_v, t_fake, N, C, use_ce, loss, gy_i, gy_v);
  MCG_CHECK_LAUNCH("mcg_loss_gen");
  return 0;
}

int mcg_int_add(int* p, int delta, void* stream) {
  if (!p) MCG_FAIL(MCG_ERR_SHAPE, "mcg_int_add: null pointer");
  pdl(int_add_kernel, 1, 1, 0, as_stream(stream))(p, delta);
  MCG_CHECK_LAUNCH("mcg_int_add");
  return 0;
}

}  // extern "C"
