/* mcg.h — C ABI of libmcg.so: the B200 (sm_100a) kernels behind the MoCoGAN training step.
 *
 * The reference (raahii/mocogan-chainer) has no native code and no FFI: every operator on its hot path is a
 * chainer==3.1.0 FunctionNode reached from model/net.py and model/updater.py.  Each entry point below therefore
 * cites the reference call site whose Chainer operator it stands in for; the Python FunctionNode subclasses in
 * mocogan_chainer_b200/functions.py bind these symbols with ctypes (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; the caller owns all buffers;
 *   - kernels never allocate, free or synchronise; scratch comes through (workspace, workspace_bytes);
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*), so a whole step is graph-capturable;
 *   - return value 0 = success, <0 = framework error (bad shape / unsupported), >0 = cudaError_t;
 *     mcg_last_error() returns a thread-local message;
 *   - activations are channels-last: (N, T, H, W, C) with T = 1 for 2-D layers; element type `dtype`
 *     (MCG_F32 or MCG_BF16); weights are (Cout, kT, kH, kW, Cin): fp32 master (+ a bf16 copy for tcgen05);
 *   - there is no CPU fallback: a shape a kernel cannot take is an error.
 */
#ifndef MCG_H_
#define MCG_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCG_VERSION 100
enum { MCG_F32 = 0, MCG_BF16 = 1, MCG_U8 = 2 };   /* MCG_U8: only as the SOURCE of mcg_pack_video (pre-decoded pixels) */
enum { MCG_ACT_NONE = 0, MCG_ACT_RELU = 1, MCG_ACT_LRELU = 2, MCG_ACT_TANH = 3 };
enum { MCG_IMPL_SIMT = 0, MCG_IMPL_TC = 1 };          /* fp32 CUDA-core path | tcgen05 bf16 path */
/* OR-ed into `impl` for fprop/wgrad of a layer with Cin <= 16: the workspace already holds this x's im2col matrix
 * (written by an earlier fprop/wgrad call on the same x and geometry), so it is not rebuilt. */
#define MCG_FLAG_COLS_VALID 0x100
/* OR-ed into `impl` (tcgen05 path): the weight tensor w — and dw — has only n < g->Cout rows, because the tensor that
 * carries g->Cout channels (dy, or y for fprop) was zero-padded to a multiple of 64 channels by the caller
 * (the generator's first layer has 60 = dim_zc + dim_zm input channels, net.py:31,44).  Rows n..Cout-1 read as zero,
 * the matching columns of dw are not written, and fprop must not be given a bias. */
#define MCG_W_ROWS(n) (((n) & 0xffff) << 16)
enum { MCG_ERR_SHAPE = -1, MCG_ERR_UNSUPPORTED = -2, MCG_ERR_WORKSPACE = -3, MCG_ERR_DRIVER = -4 };

int mcg_version(void);
const char* mcg_last_error(void);

/* Geometry of one convolution (chainer L.Convolution2D / L.ConvolutionND: net.py:133-137,174-178; the same
 * struct read "backwards" describes L.DeconvolutionND: net.py:44-48, whose forward is this conv's dgrad). */
typedef struct {
  int N, Cin, Cout;
  int Ti, Hi, Wi;   /* conv input spatial extent  (T = 1 for 2-D) */
  int To, Ho, Wo;   /* conv output spatial extent */
  int kT, kH, kW;
  int sT, sH, sW;
  int pT, pH, pW;
} mcg_conv_geom;

/* ---- convolutions --------------------------------------------------------------------------------------
 * fprop : y[n,to,ho,wo,co] = bias[co] + sum_{kt,kh,kw,ci} x[n, to*sT-pT+kt, ..., ci] * w[co,kt,kh,kw,ci]
 *         (F.convolution_2d / convolution_nd forward; also deconvolution backward-data)
 * dgrad : dx[n,ti,hi,wi,ci] (+)= bias[ci] + sum dy[n,to,ho,wo,co] * w[co,kt,kh,kw,ci]
 *         (convolution backward-data; also F.deconvolution_nd FORWARD, where bias is the deconv bias)
 * wgrad : dw[co,kt,kh,kw,ci] += sum_{n,to,ho,wo} dy[..co] * x[..ci]      (always accumulates, fp32)
 * impl = MCG_IMPL_SIMT: x/dy/dx/y of type `dtype`, w fp32.  impl = MCG_IMPL_TC: activations bf16, w bf16,
 *         requires Cin % 64 == 0 and Cout % 64 == 0 and stride in {1,2}.  out_dtype selects y / dx type.
 *         Layers with Cin <= 16 (the 3-channel image layers) are also taken: fprop/wgrad through an im2col
 *         matrix in `workspace`, dgrad as Z = dy . w^T (a tcgen05 GEMM into `workspace`) + a line-staged col2im.
 * mcg_conv_workspace_bytes: scratch the three calls need for this geometry (0 for the implicit-GEMM shapes). */
size_t mcg_conv_workspace_bytes(const mcg_conv_geom* g, int impl);
int mcg_conv_fprop(const mcg_conv_geom* g, const void* x, const void* w, const float* bias, void* y, int dtype,
                   int out_dtype, int impl, void* workspace, size_t workspace_bytes, void* stream);
int mcg_conv_dgrad(const mcg_conv_geom* g, const void* dy, const void* w, const float* bias, void* dx, int dtype,
                   int out_dtype, int accumulate, int impl, void* workspace, size_t workspace_bytes, void* stream);
int mcg_conv_wgrad(const mcg_conv_geom* g, const void* x, const void* dy, float* dw, int dtype, int impl,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ---- per-channel reductions over a channels-last matrix [M][C] -----------------------------------------
 * mcg_bn_stats: training-mode L.BatchNormalization statistics (net.py:50-53,139-141,180-182; Chainer CPU
 *   path): mean, biased var; writes mean[C], invstd[C] = 1/sqrt(var+eps), and the fused affine
 *   scale[C] = gamma*invstd, shift[C] = beta - mean*scale; updates running stats in place when non-NULL:
 *   avg_mean = decay*avg_mean + (1-decay)*mean; avg_var = decay*avg_var + (1-decay)*m/max(m-1,1)*(var+eps).
 * workspace: mcg_colreduce_workspace_bytes(M, C) bytes that the caller ZERO-FILLS ONCE, when it allocates them, and
 *   reuses for every call of this group on the same stream: each reduction is ONE kernel — the blocks add their column
 *   sums into 16 slot rows, and the block that draws the last ticket finalises (statistics / sums -> outputs) and
 *   leaves slots and ticket counter zeroed again.  Summation order across blocks is not fixed (fp32 red.add).     */
size_t mcg_colreduce_workspace_bytes(long long M, int C);
int mcg_bn_stats(const void* y, long long M, int C, int dtype, const float* gamma, const float* beta, float eps,
                 float decay, float* mean, float* invstd, float* scale, float* shift, float* avg_mean,
                 float* avg_var, void* workspace, size_t workspace_bytes, void* stream);
/* mcg_colsum: out[c] (+)= sum_m g[m][c]  — bias gradients.                                                */
int mcg_colsum(const void* g, long long M, int C, int dtype, float* out, int accumulate, void* workspace,
               size_t workspace_bytes, void* stream);

/* ---- fused elementwise passes ---------------------------------------------------------------------------
 * mcg_affine_act_noise: out = act(scale[c]*y + shift[c]) + sigma*noise   (scale/shift NULL -> identity)
 *   = L.BatchNormalization apply + F.relu / F.leaky_relu(0.2) / F.tanh + add_noise (net.py:10-15,110-114,
 *   148-155,189-196) in one pass.  Noise source: `noise` (fp32, injected, indexed n*ns_n + c*ns_c + p*ns_p with
 *   p the flattened (t,h,w) position and n = m / P) when non-NULL; else Philox N(0,1) keyed by
 *   (rng_state, call_id) when sigma != 0 and rng_state != NULL; else none.                                */
int mcg_affine_act_noise(const void* y, long long M, int C, long long P, int dtype, const float* scale,
                         const float* shift, int act, float slope, float sigma, const float* noise,
                         long long ns_n, long long ns_c, long long ns_p, const void* rng_state, int call_id,
                         void* out, int out_dtype, void* stream);
/* mcg_pack_video: gathers a video given by arbitrary element strides into channels-last (N,T',H,W,C) and adds
 *   noise as above: Variable x[:, :, t] / transpose / add_noise of updater.py:97-108 and net.py:148,189.
 *   If frame_ptr != NULL only frame *frame_ptr is taken (T' = 1), read on the device so graphs replay.
 *   src_dtype MCG_U8: src holds pre-decoded pixels, read as (v - 128) / 128 (datasets.py:91) — the input pipeline's
 *   normalisation fused into this pass. */
int mcg_pack_video(const void* src, int src_dtype, int N, int C, int T, int H, int W, long long s_n,
                   long long s_c, long long s_t, long long s_h, long long s_w, const int* frame_ptr, float sigma,
                   const float* noise, long long ns_n, long long ns_c, long long ns_p, const void* rng_state,
                   int call_id, void* out, int out_dtype, void* stream);
/* Backward of (BN ->) activation:  g' = g * act'(pre), pre = scale*y+shift.
 * mcg_act_bn_bwd_reduce : dbeta[c] = sum g', dgamma[c] = sum g' * xhat  (xhat = (y-mean)*invstd); the optional
 *                         acc_dgamma/acc_dbeta (the parameters' .grad) are incremented by the same sums
 * mcg_act_bn_bwd_apply  : gy = gamma*invstd*(g' - (xhat*dgamma + dbeta)/M)     (Chainer BN backward, A.4)
 * With mean == NULL (no BN) apply degenerates to gy = g * act'(y) and reduce must not be called.
 * For MCG_ACT_TANH / MCG_ACT_RELU `y` may be the saved OUTPUT when use_output != 0 (tanh' = 1 - out^2).  */
int mcg_act_bn_bwd_reduce(const void* g, const void* y, long long M, int C, int dtype, const float* mean,
                          const float* invstd, const float* scale, const float* shift, int act, float slope,
                          float* dgamma, float* dbeta, float* acc_dgamma, float* acc_dbeta, void* workspace,
                          size_t workspace_bytes, void* stream);
int mcg_act_bn_bwd_apply(const void* g, const void* y, long long M, int C, int dtype, const float* mean,
                         const float* invstd, const float* gamma, const float* scale, const float* shift, int act,
                         float slope, int use_output, const float* dgamma, const float* dbeta, void* gy,
                         int out_dtype, void* stream);
/* mcg_tanh_bwd_video: g = (gv[n,t,h,w,c] + (t == *frame_ptr ? gi[n,h,w,c] : 0)) * (1 - out^2), re-ordered from
 *   the discriminators' (N,T,H,W,C) to the generator's (T,N,H,W,C) row order: the transpose + get_item backward
 *   of updater.py:102,107 fused with F.tanh backward (net.py:114).                                        */
int mcg_tanh_bwd_video(const void* gv, const void* gi, int g_dtype, const void* out_tn, int out_dtype, int N,
                       int T, int HW, int C, const int* frame_ptr, void* g_tn, int gout_dtype, void* stream);

/* mcg_video_to_uint8: the sample post-processing of generate_samples.py:39 and util.py:30-51,100-101 in one pass over the
 *   generator's output storage videos (T*N, H, W, C) (channels-last, tanh range): u8 (T, N, C, H, W) =
 *   ((v / 2 + 0.5) * 255) truncated to uint8 and/or grid (T, C, size*H, size*W) = to_grid(u8, size) (cells beyond N
 *   black).  Either output may be NULL.                                                                     */
int mcg_video_to_uint8(const void* videos, int dtype, int T, int N, int C, int H, int W, unsigned char* u8,
                       unsigned char* grid, int size, void* stream);

/* ---- motion-code GRU (L.StatelessGRU, net.py:39-41,61-81) ----------------------------------------------
 * params: 12 device pointers in the order W_r.W,W_r.b,U_r.W,U_r.b,W_z.W,W_z.b,U_z.W,U_z.b,W.W,W.b,U.W,U.b
 *   (W_*: (H, L+H) row-major, U_*: (H,H)); labels int32[N] or NULL (L = 0); h0 (N,H); eps (T,N,H); zc (N,Zc).
 * forward writes z (T*N, Zc+H) = [zc tiled | h_1..h_T] (net.py:99-107) and cache (T,N,4,H) = r,z,hbar,h_prev.
 * backward takes gz (T*N, Zc+H) (only the last H columns are read) and ACCUMULATES into the 12 grads.   */
int mcg_gru_forward(const float* const* params_host, const int* labels, int L, const float* h0, const float* eps,
                    const float* zc, int T, int N, int H, int Zc, float* z, float* cache, void* stream);
int mcg_gru_backward(const float* const* params_host, float* const* grads_host, const int* labels, int L,
                     const float* eps, const float* cache, const float* gz, int T, int N, int H, int Zc,
                     void* stream);

/* ---- losses (updater.py:21-63) ---------------------------------------------------------------------------
 * y_*: (N, C) fp32 discriminator outputs (C = 1, or 1+L for infogan).  t_*: int32 labels or NULL.
 * loss_dis: sum softplus(-y_real[:1])/N + sum softplus(y_fake)[:1]/N  (+ CE(y[:,1:], t) twice if use_ce)
 * loss_gen: sum softplus(-yi[:,0])/N + sum softplus(-yv[:,0])/N        (+ CE twice if use_ce)
 * Each writes the scalar loss and the gradients w.r.t. both inputs.                                      */
int mcg_loss_dis(const float* y_real, const float* y_fake, const int* t_real, const int* t_fake, int N, int C,
                 int use_ce, float* loss, float* gy_real, float* gy_fake, void* stream);
int mcg_loss_gen(const float* y_i, const float* y_v, const int* t_fake, int N, int C, int use_ce, float* loss,
                 float* gy_i, float* gy_v, void* stream);

/* ---- optimiser: chainer.optimizers.Adam + WeightDecay hook (train.py:93-101; Chainer v3.1.0 rule) -------
 * g = grad_scale*g + wd*p; m += (1-b1)(g-m); v += (1-b2)(g*g-v); p -= alpha*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps)
 * t is read from *t_ptr (device int, already incremented for this step).  p_bf16 (optional) receives the
 * rounded copy the tcgen05 kernels read.  One launch covers a whole model's flat parameter buffer.       */
int mcg_adam_step(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, float alpha, float beta1,
                  float beta2, float eps, float wd, float grad_scale, const int* t_ptr, void* stream);
int mcg_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream);
/* The inverse (n % 8 == 0, 16-byte aligned): the data-parallel layer sends a model's flat gradient through the
 * all-reduce as bf16 (half the NVLink bytes) and widens the sum back into the fp32 gradient buffer Adam reads.    */
int mcg_cast_bf16_to_f32(const void* src, float* dst, long long n, void* stream);
/* Link.cleargrads (Chainer: optimizer.update -> target.cleargrads, updater.py:111-113): the flat gradient buffer of a
 * model is zero-filled on the stream (cudaMemsetAsync: a memset node in a captured step, no kernel).           */
int mcg_fill_zero(void* p, size_t bytes, void* stream);
/* dst[row][0..Cp) = src[row][0..C), zeros beyond: the generator's latent z (net.py:106-107, 60 = dim_zc + dim_zm
 * channels) padded to the 64 channels the tcgen05 path needs (see MCG_W_ROWS).                                  */
int mcg_pad_channels(const void* src, void* dst, long long rows, int C, int Cp, int dtype, void* stream);

/* ---- device-side step state (RNG key, step counter, frame index) so a captured step replays fresh -------
 * state layout (8 x uint32): seed_lo, seed_hi, step, frame_t, adam_t, reserved[3].
 * mcg_step_advance: step += 1; adam_t += 1; frame_t = philox(seed, step) % T  (updater.py:96).           */
int mcg_step_state_init(void* state, unsigned long long seed, void* stream);
int mcg_step_advance(void* state, int T, void* stream);
/* N(0, sigma^2) fill: make_hidden (net.py:55-56) moved to the device.                                     */
int mcg_randn(float* out, long long n, float sigma, const void* rng_state, int call_id, void* stream);
int mcg_randint(int* out, long long n, int high, const void* rng_state, int call_id, void* stream);

int mcg_int_add(int* p, int delta, void* stream);   /* *p += delta on the stream (per-optimizer Adam step counter) */
/* Reads (and optionally clears) the device flag the tcgen05 kernels raise when a bounded mbarrier wait expires.
 * Synchronises the device; for tests and smoke(), never on the hot path.                                  */
int mcg_tc_error_flag(int reset);

/* SMs the persistent tcgen05 kernels may occupy (0 = all; default from MCG_TC_SMS).  Additive, no reference equivalent:
 * the data-parallel layer reserves a few SMs for the CTAs of the all-reduce that overlaps backward, so a
 * one-CTA-per-SM grid never runs a second wave behind them.  Not a per-launch argument: set once before the first step. */
int mcg_set_tc_sm_limit(int sms);
int mcg_get_tc_sm_limit(void);

/* Number of kernels launched through this library since load (bench.py's gpu_launches).                   */
long long mcg_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MCG_H_ */
