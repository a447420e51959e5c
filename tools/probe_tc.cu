// probe_tc.cu — one-shot hardware probe for the descriptor conventions the implicit-GEMM kernels rely on.
//   T1 K-major A x K-major B            T2 K-major A x MN-major B      T3 MN-major A x K-major B
//   T4 MN-major A x MN-major B          T5 K-major B with N = 16
//   T6 5-D TMA box with elementStrides = 2, negative start coordinates and out-of-bounds zero fill
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_tc probe_tc.cu
// Every mbarrier wait is bounded, so a wrong guess prints FAIL instead of hanging the box.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "../mocogan_chainer_b200/csrc/tc_prims.cuh"

using namespace mcg;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

#define CK(x)                                                                       \
  do {                                                                              \
    cudaError_t e_ = (x);                                                           \
    if (e_ != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                      \
    }                                                                               \
  } while (0)

static void make_map(CUtensorMap* m, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, const uint32_t* estr) {
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = estr[i]; }
  for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, base, gd, gs, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(3); }
}

constexpr int KB = 4;  // k-blocks of 64

template <int A_MN, int B_MN, int BN>
__global__ void __launch_bounds__(192, 1)
gemm_probe(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float* C,
           int* err) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int A_BYTES = 128 * 64 * 2;
  constexpr int B_BYTES = (BN < 64 ? 64 : BN) * 64 * 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + KB * A_BYTES;
  __shared__ uint64_t full[KB];
  __shared__ uint64_t done;
  __shared__ uint32_t tmem_base;
  int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x == 0) {
    for (int i = 0; i < KB; ++i) mbar_init(&full[i], 1);
    mbar_init(&done, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(&tmem_base, BN < 32 ? 32 : BN); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tbase = tmem_base;
  if (warp == 0 && lane == 0) {
    for (int kb = 0; kb < KB; ++kb) {
      uint32_t bytes = 128 * 64 * 2 + BN * 64 * 2;
      mbar_arrive_expect_tx(&full[kb], bytes);
      if (A_MN) {  // A stored [K][M]: two 64-wide M slabs, each 64 k-rows of 128 B
        tma_load_2d(sA + kb * A_BYTES, &mapA, &full[kb], 0, kb * 64);
        tma_load_2d(sA + kb * A_BYTES + 8192, &mapA, &full[kb], 64, kb * 64);
      } else {
        tma_load_2d(sA + kb * A_BYTES, &mapA, &full[kb], kb * 64, 0);
      }
      if (B_MN) {
        for (int s = 0; s < BN / 64; ++s)
          tma_load_2d(sB + kb * B_BYTES + s * 8192, &mapB, &full[kb], s * 64, kb * 64);
      } else {
        tma_load_2d(sB + kb * B_BYTES, &mapB, &full[kb], kb * 64, 0);
      }
    }
  } else if (warp == 1 && lane == 0) {
    uint32_t idesc = make_idesc_bf16(128, BN, A_MN, B_MN);
    bool ok = true;
    for (int kb = 0; kb < KB && ok; ++kb) {
      ok = mbar_wait(&full[kb], 0, err);
      tc_fence_after();
      if (!ok) break;
      uint32_t a0 = smem_u32(sA + kb * A_BYTES), b0 = smem_u32(sB + kb * B_BYTES);
      for (int k = 0; k < 4; ++k) {
        uint64_t ad = A_MN ? make_smem_desc(a0 + k * 2048, 8192, 1024) : make_smem_desc(a0 + k * 32, 16, 1024);
        uint64_t bd = B_MN ? make_smem_desc(b0 + k * 2048, 8192, 1024) : make_smem_desc(b0 + k * 32, 16, 1024);
        umma_bf16(tbase, ad, bd, idesc, (kb | k) ? 1u : 0u);
      }
    }
    umma_commit(&done);
  } else if (warp >= 2) {
    int q = warp % 4;  // TMEM lane quadrant this warp may read
    bool ok = mbar_wait(&done, 0, err);
    tc_fence_after();
    if (ok) {
      for (int c0 = 0; c0 < BN; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tbase + (uint32_t(q * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) C[(q * 32 + lane) * BN + c0 + j] = __uint_as_float(v[j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tbase, BN < 32 ? 32 : BN);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

template <int A_MN, int B_MN, int BN>
static bool run_gemm(const char* name) {
  const int M = 128, K = KB * 64;
  std::vector<float> A(M * K), B(BN * K);
  srand(7);
  for (auto& v : A) v = bf((rand() % 17 - 8) / 8.0f);
  for (auto& v : B) v = bf((rand() % 13 - 6) / 4.0f);
  std::vector<__nv_bfloat16> hA(M * K), hB(BN * K);
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) hA[A_MN ? k * M + m : m * K + k] = __float2bfloat16(A[m * K + k]);
  for (int n = 0; n < BN; ++n)
    for (int k = 0; k < K; ++k) hB[B_MN ? k * BN + n : n * K + k] = __float2bfloat16(B[n * K + k]);
  __nv_bfloat16 *dA, *dB;
  float* dC;
  int* derr;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dC, M * BN * 4));
  CK(cudaMalloc(&derr, 4));
  CK(cudaMemset(derr, 0, 4));
  CK(cudaMemset(dC, 0xff, M * BN * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap mA, mB;
  uint32_t one[2] = {1, 1};
  if (A_MN) {
    uint64_t d[2] = {(uint64_t)M, (uint64_t)K}, s[1] = {(uint64_t)M * 2};
    uint32_t b[2] = {64, 64};
    make_map(&mA, dA, 2, d, s, b, one);
  } else {
    uint64_t d[2] = {(uint64_t)K, (uint64_t)M}, s[1] = {(uint64_t)K * 2};
    uint32_t b[2] = {64, 128};
    make_map(&mA, dA, 2, d, s, b, one);
  }
  if (B_MN) {
    uint64_t d[2] = {(uint64_t)BN, (uint64_t)K}, s[1] = {(uint64_t)BN * 2};
    uint32_t b[2] = {64, 64};
    make_map(&mB, dB, 2, d, s, b, one);
  } else {
    uint64_t d[2] = {(uint64_t)K, (uint64_t)BN}, s[1] = {(uint64_t)K * 2};
    uint32_t b[2] = {64, (uint32_t)BN};
    make_map(&mB, dB, 2, d, s, b, one);
  }
  size_t smem = KB * (128 * 64 * 2 + (BN < 64 ? 64 : BN) * 64 * 2) + 1024;
  CK(cudaFuncSetAttribute(gemm_probe<A_MN, B_MN, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gemm_probe<A_MN, B_MN, BN><<<1, 192, smem>>>(mA, mB, dC, derr);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: FAIL launch/sync error %s\n", name, cudaGetErrorString(e)); exit(4); }
  int herr = 0;
  CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
  std::vector<float> C(M * BN);
  CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  int bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < BN; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * B[n * K + k];
      double d = fabs(ref - C[m * BN + n]);
      if (!(d <= 1e-3)) { if (bad < 4) printf("  mismatch m=%d n=%d got %f want %f\n", m, n, C[m * BN + n], ref); ++bad; }
      if (d > maxerr) maxerr = d;
    }
  bool pass = (bad == 0 && herr == 0);
  printf("%s: %s (timeout=%d, bad=%d, maxerr=%g)\n", name, pass ? "PASS" : "FAIL", herr, bad, maxerr);
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(derr);
  return pass;
}

// ---------------------------------------------------------------- T6: strided 5-D box
__global__ void tma5d_probe(const __grid_constant__ CUtensorMap map, uint8_t* out, int bytes, int c1, int c2, int c3,
                            int c4, int* err) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) smem[i] = 0xAB;
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, bytes);
    tma_load_5d(smem, &map, &bar, 0, c1, c2, c3, c4);
  }
  bool ok = mbar_wait(&bar, 0, err);
  __syncthreads();
  if (ok)
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}

static bool run_tma5d() {
  const int C = 64, W = 8, H = 8, T = 3, N = 3;
  std::vector<__nv_bfloat16> h((size_t)N * T * H * W * C);
  auto val = [&](int n, int t, int y, int x, int c) { return (float)(((n * T + t) * H + y) * W + x) + c / 64.0f; };
  for (int n = 0; n < N; ++n) for (int t = 0; t < T; ++t) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x)
    for (int c = 0; c < C; ++c) h[((((size_t)n * T + t) * H + y) * W + x) * C + c] = __float2bfloat16(val(n, t, y, x, c));
  __nv_bfloat16* d;
  uint8_t* dout;
  int* derr;
  const int rows = 4 * 4 * 2 * 2, bytes = rows * 128;
  CK(cudaMalloc(&d, h.size() * 2));
  CK(cudaMalloc(&dout, bytes));
  CK(cudaMalloc(&derr, 4));
  CK(cudaMemset(derr, 0, 4));
  CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  uint64_t dims[5] = {C, W, H, T, N};
  uint64_t str[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2, (uint64_t)T * H * W * C * 2};
  uint32_t box[5] = {64, 8, 8, 2, 2};
  uint32_t es[5] = {1, 2, 2, 1, 1};
  CUtensorMap m;
  make_map(&m, d, 5, dims, str, box, es);
  CK(cudaFuncSetAttribute(tma5d_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes + 1024));
  const int c1 = -1, c2 = -1, c3 = 1, c4 = 2;
  tma5d_probe<<<1, 128, bytes + 1024>>>(m, dout, bytes, c1, c2, c3, c4, derr);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("T6: FAIL launch/sync error %s\n", cudaGetErrorString(e)); exit(4); }
  int herr = 0;
  CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
  std::vector<uint8_t> out(bytes);
  CK(cudaMemcpy(out.data(), dout, bytes, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int r = 0; r < rows && !herr; ++r) {
    int iw = r % 4, ih = (r / 4) % 4, it = (r / 16) % 2, in = r / 32;
    int x = c1 + 2 * iw, y = c2 + 2 * ih, t = c3 + it, n = c4 + in;
    bool oob = x < 0 || x >= W || y < 0 || y >= H || t < 0 || t >= T || n < 0 || n >= N;
    for (int c = 0; c < C; ++c) {
      int chunk = c / 8, within = c % 8;
      int phys = r * 128 + ((chunk ^ (r % 8)) * 16) + within * 2;
      __nv_bfloat16 got;
      memcpy(&got, &out[phys], 2);
      float want = oob ? 0.0f : bf(val(n, t, y, x, c));
      if (__bfloat162float(got) != want) {
        if (bad < 6) printf("  T6 mismatch row %d (n%d t%d y%d x%d) c%d got %f want %f\n", r, n, t, y, x, c, __bfloat162float(got), want);
        ++bad;
      }
    }
  }
  bool pass = bad == 0 && herr == 0;
  printf("T6 tma5d stride2/neg/oob: %s (timeout=%d, bad=%d)\n", pass ? "PASS" : "FAIL", herr, bad);
  return pass;
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  g_encode = (EncodeTiledFn)fn;
  if (!g_encode) { printf("no cuTensorMapEncodeTiled\n"); return 5; }
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s sm_%d%d, %d SMs\n", p.name, p.major, p.minor, p.multiProcessorCount);
  int fails = 0;
  fails += !run_tma5d();
  fails += !run_gemm<0, 0, 128>("T1 K-major A, K-major B, N=128");
  fails += !run_gemm<0, 1, 128>("T2 K-major A, MN-major B, N=128");
  fails += !run_gemm<1, 0, 128>("T3 MN-major A, K-major B, N=128");
  fails += !run_gemm<1, 1, 256>("T4 MN-major A, MN-major B, N=256");
  fails += !run_gemm<0, 0, 16>("T5 K-major A, K-major B, N=16");
  fails += !run_gemm<0, 0, 256>("T7 K-major A, K-major B, N=256");
  printf("probe done, %d failing\n", fails);
  return fails ? 1 : 0;
}
