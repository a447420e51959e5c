// tc_conv.cu — tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 in, fp32 accumulate).
//
// One warp-specialised kernel, three operand views over channels-last activations (C, W, H, T, N) and
// (Cout, taps, Cin) weights:
//   fprop : D[128 output pixels][BN cout]   += A(x box of tap j, 64 ci)  * B(w rows = cout, K-major)
//   dgrad : D[128 input pixels of one stride class][BN cin] += A(dy box shifted by tap j, 64 co) * B(w, MN-major)
//           (stride-2 layers are decomposed into sH*sW parity classes, each a stride-1 conv over dy: no zero taps)
//   wgrad : D[2 x 64 (tap,ci)][BN cout]     += A(x box, MN-major: K = 64 pixels) * B(dy box, MN-major), split over
//           pixel boxes across CTAs, fp32 red.add into dw.
// A tiles are 5-D TMA boxes (elementStrides carry the conv stride, out-of-bounds coordinates give the zero padding),
// so no im2col buffer ever exists in HBM.  Roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM owner),
// warps 2-5 = epilogue (TMEM -> registers -> global).  Smem ring of STAGES x (16 KB A + BN*128 B B).
#include "common.cuh"
#include "tc_prims.cuh"
#include <map>
#include <mutex>
#include <vector>
#include <string.h>

namespace mcg {

enum { kFprop = 0, kDgrad = 1, kWgrad = 2 };

struct TcTap {
  int16_t dw, dh, dt, kidx;  // A-box coordinate offsets; kidx = linear tap index (kt,kh,kw)
};
struct TcParams {
  int BW, BH, BT, BB;      // pixel-box extents
  int nbw, nbh, nbt, nbb;  // boxes per dimension
  int EW, EH, ET, EN;      // extents the boxes tile (class sub-grid for dgrad) — used for masking
  int a_mul_w, a_mul_h, a_mul_t, a_add_w, a_add_h, a_add_t;  // A-box start = tile_start*mul + add + tap offset
  int o_mul_w, o_mul_h, o_mul_t;                              // output pixel = sub-grid coord*mul + class phase
  long long os_w, os_h, os_t, os_n;                           // output strides (elements)
  int cls_w, cls_h, cls_t;                                    // stride classes (1,1,1 unless dgrad)
  int full_w, full_h, full_t;                                 // full output extents (dgrad class masking)
  int chunks;                                                 // 64-channel K chunks per tap (fprop/dgrad)
  int Cin, Cout, Ktot;                                        // Ktot = taps*Cin (row length of w / dw)
  int tap_begin[8], tap_count[8];
  int total_boxes, boxes_per_split;                           // wgrad
  int total_slabs, kreal;                                     // wgrad: valid 64-row slabs; real row length of dw
  int out_f32;
  int planar_chunk, planar_cols;   // fprop: write column c to plane c/chunk as [plane][pixel][chunk] (0 = row-major)
  long long planar_stride;
  TcTap taps[64];
};

__device__ int g_tc_error = 0;

constexpr int kTcThreads = 192;
constexpr int A_BYTES = 128 * 128;  // 128 rows x 64 bf16

template <int MODE, int BN, int STAGES>
__global__ void __launch_bounds__(kTcThreads) tc_conv_kernel(const __grid_constant__ CUtensorMap mapA,
                                                             const __grid_constant__ CUtensorMap mapB,
                                                             const __grid_constant__ TcParams P, void* __restrict__ out,
                                                             const float* __restrict__ bias) {
  constexpr int B_BYTES = BN * 128;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* err = &g_tc_error;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&done_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
  }
  constexpr int TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));  // power of two >= BN
  if (warp == 1) { tmem_alloc(&tmem_slot, TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  // ---- tile decode ---------------------------------------------------------------------------------------
  int cls = 0, w0 = 0, h0 = 0, t0 = 0, n0 = 0;
  int nk = 0;
  int pb_begin = 0;
  if (MODE != kWgrad) {
    int b = blockIdx.x;
    int bw = b % P.nbw; b /= P.nbw;
    int bh = b % P.nbh; b /= P.nbh;
    int bt = b % P.nbt; b /= P.nbt;
    int bb = b % P.nbb; b /= P.nbb;
    cls = b;
    w0 = bw * P.BW; h0 = bh * P.BH; t0 = bt * P.BT; n0 = bb * P.BB;
    nk = P.tap_count[cls] * P.chunks;
  } else {
    pb_begin = blockIdx.z * P.boxes_per_split;
    int pe = pb_begin + P.boxes_per_split;
    if (pe > P.total_boxes) pe = P.total_boxes;
    nk = pe - pb_begin;
    if (nk < 0) nk = 0;
  }
  const int ncol0 = blockIdx.y * BN;

  if (warp == 0) {
    // ================================================= TMA producer =========================================
    if (lane == 0) {
      int u0 = blockIdx.x * 2;  // wgrad: the two 64-row slabs = (tap, chunk) pairs u0, u0+1
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        if (!mbar_wait(&empty_bar[s], ph ^ 1, err)) break;
        uint8_t* a_dst = smem + s * STAGE_BYTES;
        uint8_t* b_dst = a_dst + A_BYTES;
        mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
        if (MODE != kWgrad) {
          const int j = P.tap_begin[cls] + kb / P.chunks, c = kb % P.chunks;
          const TcTap tp = P.taps[j];
          tma_load_5d(a_dst, &mapA, &full_bar[s], c * 64, w0 * P.a_mul_w + P.a_add_w + tp.dw,
                      h0 * P.a_mul_h + P.a_add_h + tp.dh, t0 * P.a_mul_t + P.a_add_t + tp.dt, n0);
          if (MODE == kFprop) {
            tma_load_2d(b_dst, &mapB, &full_bar[s], tp.kidx * P.Cin + c * 64, ncol0);
          } else {
#pragma unroll
            for (int sl = 0; sl < BN / 64; ++sl)
              tma_load_2d(b_dst + sl * 8192, &mapB, &full_bar[s], tp.kidx * P.Cin + ncol0 + sl * 64, c * 64);
          }
        } else {
          int pb = pb_begin + kb;
          int bw = pb % P.nbw; pb /= P.nbw;
          int bh = pb % P.nbh; pb /= P.nbh;
          int bt = pb % P.nbt; pb /= P.nbt;
          int bb = pb;
          const int pw0 = bw * P.BW, ph0 = bh * P.BH, pt0 = bt * P.BT, pn0 = bb * P.BB;
#pragma unroll
          for (int sl = 0; sl < 2; ++sl) {
            const int u = u0 + sl;
            const bool live = u < P.total_slabs;  // an odd slab count leaves one dummy slab: channel coordinate out of
            const TcTap tp = P.taps[live ? u / P.chunks : 0];  // bounds -> TMA zero-fills it
            tma_load_5d(a_dst + sl * 8192, &mapA, &full_bar[s], live ? (u % P.chunks) * 64 : P.Cin, pw0 * P.a_mul_w + P.a_add_w + tp.dw,
                        ph0 * P.a_mul_h + P.a_add_h + tp.dh, pt0 * P.a_mul_t + P.a_add_t + tp.dt, pn0);
          }
#pragma unroll
          for (int sl = 0; sl < BN / 64; ++sl)
            tma_load_5d(b_dst + sl * 8192, &mapB, &full_bar[s], ncol0 + sl * 64, pw0, ph0, pt0, pn0);
        }
      }
    }
  } else if (warp == 1) {
    // ================================================= MMA issuer ===========================================
    if (lane == 0) {
      constexpr int A_MN = (MODE == kWgrad), B_MN = (MODE != kFprop);
      const uint32_t idesc = make_idesc_bf16(128, BN, A_MN, B_MN);
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        if (!mbar_wait(&full_bar[s], ph, err)) break;
        tc_fence_after();
        const uint32_t a0 = smem_u32(smem + s * STAGE_BYTES), b0 = a0 + A_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = A_MN ? make_smem_desc(a0 + k * 2048, 8192, 1024) : make_smem_desc(a0 + k * 32, 16, 1024);
          const uint64_t bd = B_MN ? make_smem_desc(b0 + k * 2048, 8192, 1024) : make_smem_desc(b0 + k * 32, 16, 1024);
          umma_bf16(tmem, ad, bd, idesc, (kb | k) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&done_bar);
    }
  } else {
    // ================================================= epilogue ============================================
    const int q = warp & 3;
    const int r = q * 32 + lane;  // accumulator row == TMEM lane
    const bool ok = (nk > 0) && mbar_wait(&done_bar, 0, err);
    tc_fence_after();
    if (ok) {
      if (MODE != kWgrad) {
        int rr = r;
        const int iw = rr % P.BW; rr /= P.BW;
        const int ih = rr % P.BH; rr /= P.BH;
        const int it = rr % P.BT; rr /= P.BT;
        const int ib = rr;
        const int pw = cls % P.cls_w, phh = (cls / P.cls_w) % P.cls_h, pt = cls / (P.cls_w * P.cls_h);
        const int ow = (w0 + iw) * P.o_mul_w + pw, oh = (h0 + ih) * P.o_mul_h + phh, ot = (t0 + it) * P.o_mul_t + pt;
        const int on = n0 + ib;
        const bool valid = ow < P.full_w && oh < P.full_h && ot < P.full_t && on < P.EN;
        const long long base = (long long)on * P.os_n + (long long)ot * P.os_t + (long long)oh * P.os_h + (long long)ow * P.os_w + ncol0;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem + (uint32_t(q * 32) << 16) + c0, v);
          tmem_ld_wait();
          if (valid) {
            float f[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) + (bias ? bias[ncol0 + c0 + i] : 0.f);
            if (MODE == kFprop && P.planar_chunk) {
              // planar bf16 output for the narrow-Cin dgrad GEMM: plane = (kt,kh) run, so col2im reads contiguous lines
              __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const int c = ncol0 + c0 + i;
                if (c < P.planar_cols) {
                  const int plane = c / P.planar_chunk, within = c - plane * P.planar_chunk;
                  uint2 u;
                  __nv_bfloat162 lo = __floats2bfloat162_rn(f[i], f[i + 1]), hi = __floats2bfloat162_rn(f[i + 2], f[i + 3]);
                  u.x = *reinterpret_cast<uint32_t*>(&lo);
                  u.y = *reinterpret_cast<uint32_t*>(&hi);
                  *reinterpret_cast<uint2*>(o + plane * P.planar_stride + (long long)ow * P.planar_chunk + within) = u;
                }
              }
            } else if (P.out_f32) {
              float* o = reinterpret_cast<float*>(out) + base + c0;
#pragma unroll
              for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
            } else {
              __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + base + c0;
#pragma unroll
              for (int i = 0; i < 32; i += 8) {
                uint4 u;
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
                h[0] = __floats2bfloat162_rn(f[i], f[i + 1]);
                h[1] = __floats2bfloat162_rn(f[i + 2], f[i + 3]);
                h[2] = __floats2bfloat162_rn(f[i + 4], f[i + 5]);
                h[3] = __floats2bfloat162_rn(f[i + 6], f[i + 7]);
                *reinterpret_cast<uint4*>(o + i) = u;
              }
            }
          }
        }
      } else {
        const int u = blockIdx.x * 2 + (r >> 6);
        const bool live = u < P.total_slabs;
        const TcTap tp = P.taps[live ? u / P.chunks : 0];
        const long long kidx = (long long)tp.kidx * P.Cin + (u % P.chunks) * 64 + (r & 63);
        const bool row_ok = live && kidx < P.kreal;
        float* dw = reinterpret_cast<float*>(out);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem + (uint32_t(q * 32) << 16) + c0, v);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < 32; ++i) atomicAdd(dw + (long long)(ncol0 + c0 + i) * P.kreal + kidx, __uint_as_float(v[i]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, TMEM_COLS);
}

// =============================================================================================================
// host side
// =============================================================================================================
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess) fn = (EncodeTiledFn)p;
  }
  return fn;
}

struct MapKey {
  const void* base;
  int rank;
  uint64_t dims[5], strides[4];
  uint32_t box[5], es[5];
  bool operator<(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) < 0; }
};
static std::map<MapKey, CUtensorMap> g_maps;
static std::mutex g_maps_mu;

static int get_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, const uint32_t* es) {
  MapKey k;
  memset(&k, 0, sizeof(k));
  k.base = base;
  k.rank = rank;
  for (int i = 0; i < rank; ++i) { k.dims[i] = dims[i]; k.box[i] = box[i]; k.es[i] = es[i]; }
  for (int i = 0; i < rank - 1; ++i) k.strides[i] = strides_bytes[i];
  std::lock_guard<std::mutex> lk(g_maps_mu);
  auto it = g_maps.find(k);
  if (it != g_maps.end()) { *out = it->second; return 0; }
  EncodeTiledFn enc = get_encode();
  if (!enc) MCG_FAIL(MCG_ERR_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], e5[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; e5[i] = es[i]; }
  for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gd, gs, bx, e5,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    MCG_FAIL(MCG_ERR_DRIVER, "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu,%llu box %u,%u,%u", (int)r, rank,
             (unsigned long long)gd[0], (unsigned long long)gd[1], (unsigned long long)(rank > 2 ? gd[2] : 0), bx[0], bx[1],
             rank > 2 ? bx[2] : 0);
  if (g_maps.size() > 8192) g_maps.clear();
  g_maps[k] = *out;
  return 0;
}

// activation tensor (C, W, H, T, N) bf16 channels-last; box of (64 ch, bw, bh, bt, bb) pixels with element strides
static int act_map(CUtensorMap* m, const void* base, int C, int W, int H, int T, int N, int bw, int bh, int bt, int bb,
                   int sw, int sh, int st) {
  uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)T, (uint64_t)N};
  uint64_t str[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2, (uint64_t)T * H * W * C * 2};
  uint32_t box[5] = {64, (uint32_t)(bw * sw), (uint32_t)(bh * sh), (uint32_t)(bt * st), (uint32_t)bb};
  uint32_t es[5] = {1, (uint32_t)sw, (uint32_t)sh, (uint32_t)st, 1};
  return get_map(m, base, 5, dims, str, box, es);
}

struct Box { int w, h, t, b; };
static int ceil_div(int a, int b) { return (a + b - 1) / b; }
// split `target` (a power of two) pixels into a (w,h,t,b) box minimising padded volume; prefer wide boxes
static Box choose_box(int target, int W, int H, int T, int B) {
  Box best{1, 1, 1, target};
  double best_cost = 1e300;
  for (int bw = 1; bw <= target; bw *= 2)
    for (int bh = 1; bw * bh <= target; bh *= 2)
      for (int bt = 1; bw * bh * bt <= target; bt *= 2) {
        int bb = target / (bw * bh * bt);
        if (bw > 128 || bh > 128 || bt > 128 || bb > 256) continue;
        double vol = (double)ceil_div(W, bw) * bw * ceil_div(H, bh) * bh * (double)ceil_div(T, bt) * bt * ceil_div(B, bb) * bb;
        double cost = vol - 1e-3 * bw - 1e-5 * bh;  // tie-break: contiguous rows first
        if (cost < best_cost) { best_cost = cost; best = Box{bw, bh, bt, bb}; }
      }
  return best;
}

static int pick_bn(int cols, long long mtiles) {
  // widest tile that still gives every SM work; columns must divide
  int best = 64;
  const int sms = num_sms();
  if (cols == 192) return 192;  // the 3-channel layers' (tap, ci) axis: one N tile, A read once
  for (int bn = 256; bn >= 64; bn /= 2) {
    if (cols % bn) continue;
    long long ctas = mtiles * (cols / bn);
    if (ctas >= sms || bn == 64) { best = bn; break; }
  }
  return best;
}

template <int MODE, int BN>
static int launch_tc(const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& P, dim3 grid, void* out, const float* bias,
                     cudaStream_t st, const char* who) {
  constexpr int STAGE = A_BYTES + BN * 128;
  constexpr int STAGES = (BN == 256) ? 4 : (BN == 192 ? 4 : (BN == 128 ? 3 : 4));  // BN<=128: ~96 KB so two CTAs share an SM
  size_t smem = (size_t)STAGES * STAGE + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_conv_kernel<MODE, BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) MCG_FAIL((int)e, "%s: cudaFuncSetAttribute: %s", who, cudaGetErrorString(e));
    configured = true;
  }
  tc_conv_kernel<MODE, BN, STAGES><<<grid, kTcThreads, smem, st>>>(ma, mb, P, out, bias);
  MCG_CHECK_LAUNCH(who);
  return 0;
}
template <int MODE>
static int launch_tc_bn(int bn, const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& P, dim3 grid, void* out,
                        const float* bias, cudaStream_t st, const char* who) {
  switch (bn) {
    case 64: return launch_tc<MODE, 64>(ma, mb, P, grid, out, bias, st, who);
    case 128: return launch_tc<MODE, 128>(ma, mb, P, grid, out, bias, st, who);
    case 192: return launch_tc<MODE, 192>(ma, mb, P, grid, out, bias, st, who);
    case 256: return launch_tc<MODE, 256>(ma, mb, P, grid, out, bias, st, who);
  }
  MCG_FAIL(MCG_ERR_UNSUPPORTED, "%s: BN=%d", who, bn);
}

bool tc_supported(const mcg_conv_geom* g) {
  if (g->Cin % 64 || g->Cout % 64) return false;
  if (g->sT > 2 || g->sH > 2 || g->sW > 2) return false;
  if (g->kT * g->kH * g->kW > 64) return false;
  if (g->sT * g->sH * g->sW > 8) return false;
  return true;
}

int tc_conv(int mode, const mcg_conv_geom* g, const void* a, const void* b, void* out, const float* bias, int out_dtype,
            cudaStream_t st, int kreal = 0, int planar_chunk = 0, int planar_cols = 0) {
  const char* who = mode == kFprop ? "mcg_conv_fprop(tc)" : mode == kDgrad ? "mcg_conv_dgrad(tc)" : "mcg_conv_wgrad(tc)";
  if (!tc_supported(g)) MCG_FAIL(MCG_ERR_UNSUPPORTED, "%s: needs Cin,Cout %% 64 == 0, stride <= 2, <= 64 taps", who);
  TcParams P;
  memset(&P, 0, sizeof(P));
  const int taps = g->kT * g->kH * g->kW;
  P.Cin = g->Cin; P.Cout = g->Cout; P.Ktot = taps * g->Cin;
  P.out_f32 = (out_dtype == MCG_F32);
  P.planar_chunk = planar_chunk;
  P.planar_cols = planar_cols;
  P.planar_stride = (long long)g->Wo * planar_chunk;   // planar mode is only used on 1-D line geometries (M = Wo)
  CUtensorMap ma, mb;
  int rc;
  if (mode == kFprop) {
    // a = x (N,Ti,Hi,Wi,Cin), b = w bf16 (Cout, taps*Cin), out = y (N,To,Ho,Wo,Cout)
    Box bx = choose_box(128, g->Wo, g->Ho, g->To, g->N);
    P.BW = bx.w; P.BH = bx.h; P.BT = bx.t; P.BB = bx.b;
    P.nbw = ceil_div(g->Wo, bx.w); P.nbh = ceil_div(g->Ho, bx.h); P.nbt = ceil_div(g->To, bx.t); P.nbb = ceil_div(g->N, bx.b);
    P.EW = g->Wo; P.EH = g->Ho; P.ET = g->To; P.EN = g->N;
    P.full_w = g->Wo; P.full_h = g->Ho; P.full_t = g->To;
    P.a_mul_w = g->sW; P.a_mul_h = g->sH; P.a_mul_t = g->sT; P.a_add_w = -g->pW; P.a_add_h = -g->pH; P.a_add_t = -g->pT;
    P.o_mul_w = P.o_mul_h = P.o_mul_t = 1;
    P.cls_w = P.cls_h = P.cls_t = 1;
    P.os_w = g->Cout; P.os_h = (long long)g->Wo * g->Cout; P.os_t = (long long)g->Ho * P.os_h; P.os_n = (long long)g->To * P.os_t;
    P.chunks = g->Cin / 64;
    P.tap_begin[0] = 0; P.tap_count[0] = taps;
    for (int kt = 0, j = 0; kt < g->kT; ++kt)
      for (int kh = 0; kh < g->kH; ++kh)
        for (int kw = 0; kw < g->kW; ++kw, ++j) P.taps[j] = TcTap{(int16_t)kw, (int16_t)kh, (int16_t)kt, (int16_t)j};
    if ((rc = act_map(&ma, a, g->Cin, g->Wi, g->Hi, g->Ti, g->N, bx.w, bx.h, bx.t, bx.b, g->sW, g->sH, g->sT))) return rc;
    long long mt = (long long)P.nbw * P.nbh * P.nbt * P.nbb;
    int bn = pick_bn(g->Cout, mt);
    uint64_t d2[2] = {(uint64_t)P.Ktot, (uint64_t)g->Cout}, s2[1] = {(uint64_t)P.Ktot * 2};
    uint32_t b2[2] = {64, (uint32_t)bn}, e2[2] = {1, 1};
    if ((rc = get_map(&mb, b, 2, d2, s2, b2, e2))) return rc;
    dim3 grid((unsigned)mt, (unsigned)(g->Cout / bn), 1);
    return launch_tc_bn<kFprop>(bn, ma, mb, P, grid, out, bias, st, who);
  }
  if (mode == kDgrad) {
    // a = dy (N,To,Ho,Wo,Cout), b = w bf16, out = dx (N,Ti,Hi,Wi,Cin); one class per residue of the input coordinate
    const int cw = g->sW, ch = g->sH, ct = g->sT;
    const int EW = ceil_div(g->Wi, cw), EH = ceil_div(g->Hi, ch), ET = ceil_div(g->Ti, ct);
    Box bx = choose_box(128, EW, EH, ET, g->N);
    P.BW = bx.w; P.BH = bx.h; P.BT = bx.t; P.BB = bx.b;
    P.nbw = ceil_div(EW, bx.w); P.nbh = ceil_div(EH, bx.h); P.nbt = ceil_div(ET, bx.t); P.nbb = ceil_div(g->N, bx.b);
    P.EW = EW; P.EH = EH; P.ET = ET; P.EN = g->N;
    P.full_w = g->Wi; P.full_h = g->Hi; P.full_t = g->Ti;
    P.a_mul_w = P.a_mul_h = P.a_mul_t = 1;
    P.o_mul_w = cw; P.o_mul_h = ch; P.o_mul_t = ct;
    P.cls_w = cw; P.cls_h = ch; P.cls_t = ct;
    P.os_w = g->Cin; P.os_h = (long long)g->Wi * g->Cin; P.os_t = (long long)g->Hi * P.os_h; P.os_n = (long long)g->Ti * P.os_t;
    P.chunks = g->Cout / 64;
    int ncls = cw * ch * ct, j = 0;
    for (int c = 0; c < ncls; ++c) {
      const int pw = c % cw, ph = (c / cw) % ch, pt = c / (cw * ch);
      P.tap_begin[c] = j;
      for (int kt = 0; kt < g->kT; ++kt) {
        if ((pt + g->pT - kt) % ct) continue;  // C++ % keeps sign; divisibility test is still correct
        for (int kh = 0; kh < g->kH; ++kh) {
          if ((ph + g->pH - kh) % ch) continue;
          for (int kw = 0; kw < g->kW; ++kw) {
            if ((pw + g->pW - kw) % cw) continue;
            if (j >= 64) MCG_FAIL(MCG_ERR_UNSUPPORTED, "%s: tap table overflow", who);
            // floor division is exact here (numerator divisible)
            P.taps[j++] = TcTap{(int16_t)((pw + g->pW - kw) / cw), (int16_t)((ph + g->pH - kh) / ch),
                                (int16_t)((pt + g->pT - kt) / ct), (int16_t)((kt * g->kH + kh) * g->kW + kw)};
          }
        }
      }
      P.tap_count[c] = j - P.tap_begin[c];
    }
    if ((rc = act_map(&ma, a, g->Cout, g->Wo, g->Ho, g->To, g->N, bx.w, bx.h, bx.t, bx.b, 1, 1, 1))) return rc;
    long long mt = (long long)P.nbw * P.nbh * P.nbt * P.nbb * ncls;
    int bn = pick_bn(g->Cin, mt);
    uint64_t d2[2] = {(uint64_t)P.Ktot, (uint64_t)g->Cout}, s2[1] = {(uint64_t)P.Ktot * 2};
    uint32_t b2[2] = {64, 64}, e2[2] = {1, 1};
    if ((rc = get_map(&mb, b, 2, d2, s2, b2, e2))) return rc;
    dim3 grid((unsigned)mt, (unsigned)(g->Cin / bn), 1);
    return launch_tc_bn<kDgrad>(bn, ma, mb, P, grid, out, bias, st, who);
  }
  // ---- wgrad: a = x (N,Ti,Hi,Wi,Cin), b = dy (N,To,Ho,Wo,Cout), out = dw fp32 (Cout, taps*Cin), accumulated
  {
    Box bx = choose_box(64, g->Wo, g->Ho, g->To, g->N);
    P.BW = bx.w; P.BH = bx.h; P.BT = bx.t; P.BB = bx.b;
    P.nbw = ceil_div(g->Wo, bx.w); P.nbh = ceil_div(g->Ho, bx.h); P.nbt = ceil_div(g->To, bx.t); P.nbb = ceil_div(g->N, bx.b);
    P.a_mul_w = g->sW; P.a_mul_h = g->sH; P.a_mul_t = g->sT; P.a_add_w = -g->pW; P.a_add_h = -g->pH; P.a_add_t = -g->pT;
    P.chunks = g->Cin / 64;
    for (int kt = 0, j = 0; kt < g->kT; ++kt)
      for (int kh = 0; kh < g->kH; ++kh)
        for (int kw = 0; kw < g->kW; ++kw, ++j) P.taps[j] = TcTap{(int16_t)kw, (int16_t)kh, (int16_t)kt, (int16_t)j};
    const int slabs = taps * P.chunks;
    const int mtiles = (slabs + 1) / 2;   // an odd count leaves one zero-filled dummy slab in the last tile
    P.total_slabs = slabs;
    P.kreal = kreal > 0 ? kreal : P.Ktot;
    int bn = 64;
    for (int c = 256; c >= 64; c /= 2)
      if (g->Cout % c == 0) { bn = c; break; }
    P.total_boxes = P.nbw * P.nbh * P.nbt * P.nbb;
    long long tiles = (long long)mtiles * (g->Cout / bn);
    // one full wave: resident CTAs per SM is 2 for BN <= 128 (96 KB of stages), 1 for BN = 256; rounding the split
    // count UP spills a few CTAs into a second wave that doubles the kernel time (ncu: Dv.dc2, 320 CTAs on 296 slots)
    const long long slots = (long long)num_sms() * (bn <= 128 ? 2 : 1);
    long long want = slots / tiles;
    long long maxs = ceil_div(P.total_boxes, 8);
    int splits = (int)(want < maxs ? want : maxs);
    if (splits < 1) splits = 1;
    P.boxes_per_split = ceil_div(P.total_boxes, splits);
    splits = ceil_div(P.total_boxes, P.boxes_per_split);
    if ((rc = act_map(&ma, a, g->Cin, g->Wi, g->Hi, g->Ti, g->N, bx.w, bx.h, bx.t, bx.b, g->sW, g->sH, g->sT))) return rc;
    if ((rc = act_map(&mb, b, g->Cout, g->Wo, g->Ho, g->To, g->N, bx.w, bx.h, bx.t, bx.b, 1, 1, 1))) return rc;
    dim3 grid((unsigned)mtiles, (unsigned)(g->Cout / bn), (unsigned)splits);
    return launch_tc_bn<kWgrad>(bn, ma, mb, P, grid, out, nullptr, st, who);
  }
}


// =============================================================================================================
// 3-channel image layers (Di.dc1, Dv.dc1, G.dc5): Cin < 64 cannot feed a 16-byte-aligned TMA box, so
//   fprop / wgrad : x is expanded once into an explicit im2col matrix cols[M][Kp] (Kp = taps*Cin rounded up to 64) in
//                   caller workspace and the layer becomes a plain GEMM = a 1x1 "convolution" over a 1-D line of M
//                   pixels through the same tcgen05 kernel;
//   dgrad         : Z[M][Kp] = dy . w^T as the same kind of GEMM (N = taps*Cin), then a line-staged col2im gather.
// These layers are HBM/L2-bound (Dv.dc1 writes 29.8 M outputs, G.dc5 reads 36.7 M inputs), not tensor-bound.
// =============================================================================================================
__global__ void __launch_bounds__(256) im2col_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ cols,
                                                     long long M, int Kp, int Cin, int Ti, int Hi, int Wi, int To, int Ho,
                                                     int Wo, int kT, int kH, int kW, int sT, int sH, int sW, int pT, int pH,
                                                     int pW) {
  const int groups = Kp / 8;
  const int K = kT * kH * kW * Cin;
  const long long total = M * groups;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long m = idx / groups;
    const int k0 = (int)(idx % groups) * 8;
    int wo = (int)(m % Wo); long long r = m / Wo;
    int ho = (int)(r % Ho); r /= Ho;
    int to = (int)(r % To); const long long n = r / To;
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = k0 + i;
      float val = 0.f;
      if (k < K) {
        const int tap = k / Cin, ci = k - tap * Cin;
        const int kw = tap % kW, kh = (tap / kW) % kH, kt = tap / (kW * kH);
        const int ti = to * sT - pT + kt, hi = ho * sH - pH + kh, wi = wo * sW - pW + kw;
        if ((unsigned)ti < (unsigned)Ti && (unsigned)hi < (unsigned)Hi && (unsigned)wi < (unsigned)Wi)
          val = __bfloat162float(x[((((long long)n * Ti + ti) * Hi + hi) * Wi + wi) * Cin + ci]);
      }
      v[i] = __float2bfloat16_rn(val);
    }
    *reinterpret_cast<uint4*>(cols + m * Kp + k0) = *reinterpret_cast<const uint4*>(v);
  }
}
// Line-staged im2col: one CTA per output line (n, to, ho).  The kT*kH source rows it needs are read coalesced into shared
// memory (zero-padded at the borders), then the line's Wo x Kp block of cols — one contiguous span — is written with
// 16-byte stores.  Source bytes are read ~kT*kH/(sT*sH) times from L2, cols bytes are written exactly once.
template <int CIN, int KW>
__global__ void __launch_bounds__(128) im2col_line_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ cols,
                                                          int Kp, int Ti, int Hi, int Wi, int To, int Ho, int Wo, int kT, int kH,
                                                          int sT, int sH, int sW, int pT, int pH, int pW) {
  extern __shared__ unsigned short rows[];  // [kT*kH][L]
  const int runs = kT * kH;
  const int L = (Wi + pW + KW) * CIN;
  const int K = runs * KW * CIN;
  int line = blockIdx.x;
  const int ho = line % Ho; line /= Ho;
  const int to = line % To;
  const long long n = line / To;
  const unsigned short* xs = reinterpret_cast<const unsigned short*>(x);
  for (int run = 0; run < runs; ++run) {
    const int kh = run % kH, kt = run / kH;
    const int ti = to * sT - pT + kt, hi = ho * sH - pH + kh;
    const bool row_ok = (unsigned)ti < (unsigned)Ti && (unsigned)hi < (unsigned)Hi;
    const long long src = (((n * Ti + ti) * Hi + hi) * (long long)Wi) * CIN;
    for (int e = threadIdx.x; e < L; e += blockDim.x) {
      const int w = e / CIN - pW;
      rows[run * L + e] = (row_ok && (unsigned)w < (unsigned)Wi) ? __ldg(xs + src + (e - pW * CIN)) : (unsigned short)0;
    }
  }
  __syncthreads();
  const int vec_per_row = Kp / 8;
  const long long m0 = ((n * To + to) * Ho + ho) * (long long)Wo;
  uint4* dst = reinterpret_cast<uint4*>(cols + m0 * Kp);
  for (int v = threadIdx.x; v < Wo * vec_per_row; v += blockDim.x) {
    const int wo = v / vec_per_row, k0 = (v % vec_per_row) * 8;
    __align__(16) unsigned short o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = k0 + i;
      unsigned short val = 0;
      if (k < K) {
        const int run = k / (KW * CIN), rem = k % (KW * CIN);
        val = rows[run * L + wo * sW * CIN + rem];
      }
      o[i] = val;
    }
    dst[v] = *reinterpret_cast<const uint4*>(o);
  }
}
// wp[co][Kp] = w[co][k] (k < K) else 0
__global__ void pad_rows_kernel(const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ wp, int rows, int K, int Kp) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows * Kp; i += gridDim.x * blockDim.x) {
    const int r = i / Kp, k = i % Kp;
    wp[i] = k < K ? w[(long long)r * K + k] : __float2bfloat16_rn(0.f);
  }
}
// wt[j = (tap, ci), zero-padded to Jp rows][co] = w[co][tap][ci]   (K-major B operand of the dgrad GEMM)
__global__ void transpose_jk_kernel(const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ wt, int Cout, int J,
                                    int Jp) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Jp * Cout; i += gridDim.x * blockDim.x) {
    const int j = i / Cout, co = i % Cout;
    wt[i] = j < J ? w[(long long)co * J + j] : __float2bfloat16_rn(0.f);
  }
}
// col2im for the narrow-Cin dgrad: dx[n,ti,hi,wi,ci] = bias[ci] + sum over taps of Z[pixel(to,ho,wo)][tap*Cin+ci].
// One CTA per input line (n, ti, hi): the (kt,kh) runs that reach this line are gathered from Z into shared memory
// (each a run of kW*Cin contiguous elements per output pixel), then the line's Wi*Cin outputs are summed in fp32.
__global__ void __launch_bounds__(128) col2im_line_kernel(const __nv_bfloat16* __restrict__ Z, const float* __restrict__ bias,
                                                          void* __restrict__ dx, int out_f32, long long Mpix, int Cin, int Ti, int Hi,
                                                          int Wi, int To, int Ho, int Wo, int kT, int kH, int kW, int sT, int sH,
                                                          int sW, int pT, int pH, int pW) {
  extern __shared__ unsigned short zs[];  // [valid runs][Wo][kW*Cin]
  __shared__ int run_row[64], run_col[64];
  __shared__ int nvalid;
  const int RUN = kW * Cin;
  int line = blockIdx.x;
  const int hi = line % Hi; line /= Hi;
  const int ti = line % Ti;
  const long long n = line / Ti;
  if (threadIdx.x == 0) {
    int v = 0;
    for (int kt = 0; kt < kT; ++kt) {
      const int tt = ti + pT - kt;
      if (tt < 0 || tt % sT || tt / sT >= To) continue;
      for (int kh = 0; kh < kH; ++kh) {
        const int hh = hi + pH - kh;
        if (hh < 0 || hh % sH || hh / sH >= Ho) continue;
        run_row[v] = (int)(((n * To + tt / sT) * Ho + hh / sH));   // Z row block (times Wo)
        run_col[v] = kt * kH + kh;                                  // plane index
        ++v;
      }
    }
    nvalid = v;
  }
  __syncthreads();
  const int nv = nvalid;
  // Z is planar: Z[run][pixel][RUN]; a line needs, per valid run, Wo*RUN contiguous elements
  const uint32_t* zsrc = reinterpret_cast<const uint32_t*>(Z);
  uint32_t* zs32 = reinterpret_cast<uint32_t*>(zs);
  const int words = Wo * RUN / 2;
  for (int v = 0; v < nv; ++v) {
    const long long base = ((long long)run_col[v] * Mpix + (long long)run_row[v] * Wo) * RUN / 2;
    for (int e = threadIdx.x; e < words; e += blockDim.x) zs32[v * words + e] = __ldg(zsrc + base + e);
  }
  __syncthreads();
  const long long obase = ((n * Ti + ti) * Hi + hi) * (long long)Wi * Cin;
  for (int o = threadIdx.x; o < Wi * Cin; o += blockDim.x) {
    const int wi = o / Cin, ci = o % Cin;
    float acc = bias ? bias[ci] : 0.f;
    for (int kw = 0; kw < kW; ++kw) {
      const int ww = wi + pW - kw;
      if (ww < 0 || ww % sW) continue;
      const int wo = ww / sW;
      if (wo >= Wo) continue;
      for (int v = 0; v < nv; ++v) {
        const __nv_bfloat16_raw raw = {zs[(v * Wo + wo) * RUN + kw * Cin + ci]};
        acc += __bfloat162float(__nv_bfloat16(raw));
      }
    }
    if (out_f32) reinterpret_cast<float*>(dx)[obase + o] = acc;
    else reinterpret_cast<__nv_bfloat16*>(dx)[obase + o] = __float2bfloat16_rn(acc);
  }
}

static long long round_up(long long a, long long b) { return (a + b - 1) / b * b; }

bool tc_small_supported(const mcg_conv_geom* g) {
  return g->Cin <= 16 && g->Cout % 64 == 0 && g->sT <= 2 && g->sH <= 2 && g->sW <= 2 && g->kT * g->kH * g->kW <= 64 &&
         g->kW % 4 == 0;
}
size_t tc_small_workspace(const mcg_conv_geom* g) {
  const long long M = (long long)g->N * g->To * g->Ho * g->Wo;
  const long long K = (long long)g->kT * g->kH * g->kW * g->Cin, Kp = round_up(K, 64);
  return (size_t)(round_up(M * Kp * 2, 1024) + round_up((long long)g->Cout * Kp * 2, 1024) + 4096);
}

int tc_conv_small(int mode, const mcg_conv_geom* g, const void* a, const void* b, void* out, const float* bias, int out_dtype,
                  void* ws, size_t ws_bytes, cudaStream_t st, bool cols_valid) {
  const char* who = mode == kFprop ? "mcg_conv_fprop(tc,small-C)" : mode == kDgrad ? "mcg_conv_dgrad(tc,small-C)" : "mcg_conv_wgrad(tc,small-C)";
  if (!tc_small_supported(g)) MCG_FAIL(MCG_ERR_UNSUPPORTED, "%s: needs Cin <= 16, Cout %% 64 == 0", who);
  const int taps = g->kT * g->kH * g->kW;
  const long long M = (long long)g->N * g->To * g->Ho * g->Wo;
  const int K = taps * g->Cin, Kp = (int)round_up(K, 64);
  if (!ws || ws_bytes < tc_small_workspace(g)) MCG_FAIL(MCG_ERR_WORKSPACE, "%s: workspace %zu < %zu", who, ws_bytes, tc_small_workspace(g));
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~uintptr_t(1023));
  __nv_bfloat16* cols = reinterpret_cast<__nv_bfloat16*>(base);
  __nv_bfloat16* wpad = reinterpret_cast<__nv_bfloat16*>(base + round_up(M * Kp * 2, 1024));
  int rc;
  if (mode == kDgrad) {
    // a = dy (N,To,Ho,Wo,Cout), b = w bf16 (Cout,taps,Cin), out = dx (N,Ti,Hi,Wi,Cin):
    //   Z[M][Kp] = dy[M][Cout] . wt[Kp][Cout]^T  (tcgen05 GEMM over the line of M output pixels), then col2im.
    if (M > 0x7fffffffLL) MCG_FAIL(MCG_ERR_UNSUPPORTED, "%s: too many pixels", who);
    transpose_jk_kernel<<<64, 256, 0, st>>>((const __nv_bfloat16*)b, wpad, g->Cout, K, Kp);
    MCG_CHECK_LAUNCH(who);
    mcg_conv_geom g2 = {1, g->Cout, Kp, 1, 1, (int)M, 1, 1, (int)M, 1, 1, 1, 1, 1, 1, 0, 0, 0};
    const int RUN = g->kW * g->Cin;
    if ((rc = tc_conv(kFprop, &g2, a, wpad, cols, nullptr, MCG_BF16, st, 0, RUN, K))) return rc;
    const long long lines = (long long)g->N * g->Ti * g->Hi;
    const int runs_max = ceil_div(g->kT, g->sT) * ceil_div(g->kH, g->sH);
    const size_t smem = (size_t)runs_max * g->Wo * g->kW * g->Cin * 2;
    if (lines > 0x7fffffffLL || smem > 48 * 1024 || g->kT * g->kH > 64) MCG_FAIL(MCG_ERR_UNSUPPORTED, "%s: line too large", who);
    col2im_line_kernel<<<(unsigned)lines, 128, smem, st>>>(cols, bias, out, out_dtype == MCG_F32, M, g->Cin, g->Ti, g->Hi, g->Wi,
                                                           g->To, g->Ho, g->Wo, g->kT, g->kH, g->kW, g->sT, g->sH, g->sW, g->pT,
                                                           g->pH, g->pW);
    MCG_CHECK_LAUNCH(who);
    return 0;
  }
  // fprop / wgrad: im2col, then a GEMM over a 1-D line of M pixels with Kp channels
  const void* x = a;
  if (!cols_valid) {
    const long long lines = (long long)g->N * g->To * g->Ho;
    const size_t smem = (size_t)g->kT * g->kH * (g->Wi + g->pW + g->kW) * g->Cin * 2;
    if (g->kW == 4 && g->Cin == 3 && lines < 0x7fffffffLL && smem <= 48 * 1024)
      im2col_line_kernel<3, 4><<<(unsigned)lines, 128, smem, st>>>((const __nv_bfloat16*)x, cols, Kp, g->Ti, g->Hi, g->Wi, g->To, g->Ho,
                                                                   g->Wo, g->kT, g->kH, g->sT, g->sH, g->sW, g->pT, g->pH, g->pW);
    else if (g->kW == 4 && g->Cin == 1 && lines < 0x7fffffffLL && smem <= 48 * 1024)
      im2col_line_kernel<1, 4><<<(unsigned)lines, 128, smem, st>>>((const __nv_bfloat16*)x, cols, Kp, g->Ti, g->Hi, g->Wi, g->To, g->Ho,
                                                                   g->Wo, g->kT, g->kH, g->sT, g->sH, g->sW, g->pT, g->pH, g->pW);
    else
      im2col_kernel<<<num_sms() * 16, 256, 0, st>>>((const __nv_bfloat16*)x, cols, M, Kp, g->Cin, g->Ti, g->Hi, g->Wi, g->To, g->Ho,
                                                     g->Wo, g->kT, g->kH, g->kW, g->sT, g->sH, g->sW, g->pT, g->pH, g->pW);
  }
  MCG_CHECK_LAUNCH(who);
  if (M > 0x7fffffffLL) MCG_FAIL(MCG_ERR_UNSUPPORTED, "%s: too many pixels", who);
  mcg_conv_geom g2 = {1, Kp, g->Cout, 1, 1, (int)M, 1, 1, (int)M, 1, 1, 1, 1, 1, 1, 0, 0, 0};
  if (mode == kFprop) {
    const void* wk = b;
    if (Kp != K) {
      pad_rows_kernel<<<32, 256, 0, st>>>((const __nv_bfloat16*)b, wpad, g->Cout, K, Kp);
      MCG_CHECK_LAUNCH(who);
      wk = wpad;
    }
    return tc_conv(kFprop, &g2, cols, wk, out, bias, out_dtype, st);
  }
  return tc_conv(kWgrad, &g2, cols, b, out, nullptr, MCG_F32, st, K);
}

}  // namespace mcg

using namespace mcg;
namespace mcg {
int simt_conv(int mode, const mcg_conv_geom* c, const void* a, const void* b_act, const float* w, const float* bias,
              void* out, int dtype, int out_dtype, int accumulate, cudaStream_t st);
}

extern "C" {

size_t mcg_conv_workspace_bytes(const mcg_conv_geom* g, int impl) {
  if (g && (impl & 0xff) == MCG_IMPL_TC && !tc_supported(g) && tc_small_supported(g)) return tc_small_workspace(g);
  return 0;
}

int mcg_conv_fprop(const mcg_conv_geom* g, const void* x, const void* w, const float* bias, void* y, int dtype,
                   int out_dtype, int impl, void* workspace, size_t workspace_bytes, void* stream) {
  if (!g || !x || !w || !y) MCG_FAIL(MCG_ERR_SHAPE, "mcg_conv_fprop: null pointer");
  const bool cols_valid = (impl & MCG_FLAG_COLS_VALID) != 0;
  impl &= 0xff;
  if (impl == MCG_IMPL_TC) {
    if (dtype != MCG_BF16) MCG_FAIL(MCG_ERR_UNSUPPORTED, "mcg_conv_fprop(tc): activations must be bf16");
    if (!tc_supported(g) && tc_small_supported(g))
      return tc_conv_small(0, g, x, w, y, bias, out_dtype, workspace, workspace_bytes, as_stream(stream), cols_valid);
    return tc_conv(0, g, x, w, y, bias, out_dtype, as_stream(stream));
  }
  return simt_conv(0, g, x, nullptr, (const float*)w, bias, y, dtype, out_dtype, 0, as_stream(stream));
}

int mcg_conv_dgrad(const mcg_conv_geom* g, const void* dy, const void* w, const float* bias, void* dx, int dtype,
                   int out_dtype, int accumulate, int impl, void* workspace, size_t workspace_bytes, void* stream) {
  if (!g || !dy || !w || !dx) MCG_FAIL(MCG_ERR_SHAPE, "mcg_conv_dgrad: null pointer");
  impl &= 0xff;
  if (impl == MCG_IMPL_TC) {
    if (dtype != MCG_BF16) MCG_FAIL(MCG_ERR_UNSUPPORTED, "mcg_conv_dgrad(tc): activations must be bf16");
    if (accumulate) MCG_FAIL(MCG_ERR_UNSUPPORTED, "mcg_conv_dgrad(tc): accumulate not supported");
    if (!tc_supported(g) && tc_small_supported(g))
      return tc_conv_small(1, g, dy, w, dx, bias, out_dtype, workspace, workspace_bytes, as_stream(stream), false);
    return tc_conv(1, g, dy, w, dx, bias, out_dtype, as_stream(stream));
  }
  return simt_conv(1, g, dy, nullptr, (const float*)w, bias, dx, dtype, out_dtype, accumulate, as_stream(stream));
}

int mcg_conv_wgrad(const mcg_conv_geom* g, const void* x, const void* dy, float* dw, int dtype, int impl, void* workspace,
                   size_t workspace_bytes, void* stream) {
  if (!g || !x || !dy || !dw) MCG_FAIL(MCG_ERR_SHAPE, "mcg_conv_wgrad: null pointer");
  const bool cols_valid = (impl & MCG_FLAG_COLS_VALID) != 0;
  impl &= 0xff;
  if (impl == MCG_IMPL_TC) {
    if (dtype != MCG_BF16) MCG_FAIL(MCG_ERR_UNSUPPORTED, "mcg_conv_wgrad(tc): activations must be bf16");
    if (!tc_supported(g) && tc_small_supported(g))
      return tc_conv_small(2, g, x, dy, dw, nullptr, MCG_F32, workspace, workspace_bytes, as_stream(stream), cols_valid);
    return tc_conv(2, g, x, dy, dw, nullptr, MCG_F32, as_stream(stream));
  }
  return simt_conv(2, g, dy, x, nullptr, nullptr, dw, dtype, MCG_F32, 1, as_stream(stream));
}

int mcg_tc_error_flag(int reset) {
  int v = 0;
  cudaMemcpyFromSymbol(&v, g_tc_error, sizeof(int));
  if (reset) {
    int z = 0;
    cudaMemcpyToSymbol(g_tc_error, &z, sizeof(int));
  }
  return v;
}

}  // extern "C"
