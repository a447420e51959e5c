// tc_conv.cu — tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 in, fp32 accumulate).
//
// One PERSISTENT warp-specialised kernel (one CTA per SM), three operand views over channels-last activations
// (C, W, H, T, N) and (Cout, taps, Cin) weights.  A CTA tile is MT blocks of 128 accumulator rows x BN columns:
//   fprop : D[MT x 128 output pixels][BN cout] += A(x box of tap j, 64 ci)  * B(w rows = cout, K-major)
//   dgrad : D[MT x 128 input pixels of one stride class][BN cin] += A(dy box shifted by tap j, 64 co) * B(w, MN-major)
//           (stride-2 layers are decomposed into sH*sW parity classes, each a stride-1 conv over dy: no zero taps)
//   wgrad : D[MT x 2 x 64 (tap,ci)][BN cout] += A(x box, MN-major: K = 64 pixels) * B(dy box, MN-major); the K loop
//           (pixel boxes) of all tiles is one range cut evenly over the CTAs (stream-K), fp32 red.add into dw.
// A tiles are 5-D TMA boxes (elementStrides carry the conv stride, out-of-bounds coordinates give the zero padding),
// so no im2col buffer ever exists in HBM.  Roles: warp 0 = TMA producer (one elected thread), warp 1 = MMA issuer
// (+ TMEM owner), warps 2-9 = epilogue (TMEM -> registers -> global; two warps per TMEM lane quarter).  Shared-memory
// ring of STAGES x (MT*16 KB A + BN*128 B B); accumulators double-buffered in TMEM when 2*MT*BN <= 512 columns, so the
// epilogue of one tile overlaps the MMAs of the next.  Every CTA owns one contiguous range of work units (128-pixel
// boxes for fprop/dgrad, K steps for wgrad), so the load imbalance is at most one unit.
#include "common.cuh"
#include "tc_prims.cuh"
#include <atomic>
#include <map>
#include <mutex>
#include <vector>
#include <string.h>
#include <stdlib.h>

namespace mcg {

enum { kFprop = 0, kDgrad = 1, kWgrad = 2 };
#ifndef MCG_TC_DYN_DEFAULT
#define MCG_TC_DYN_DEFAULT 0
#endif

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
  // Schedule: every CTA owns one contiguous range of `units` (persistent, one CTA per SM).
  // fprop/dgrad: unit = one 128-pixel box of one row = (cls, nt); a CTA walks its range MT boxes at a time (the last
  // step of a row or range may hold fewer), so the work per CTA differs by at most one box.
  // wgrad: unit = one K step (pixel box) of one tile = (mgroup of MT blocks of 2 slabs of 64 (tap,ci) rows, nt); a tile
  // cut between CTAs (stream-K) meets again in the fp32 red.add of the epilogue.
  int nboxes, ntn, ntiles;
  long long total_units, units_per_cta;
  int dyn, sched_slot;   // dynamic work distribution (fprop/dgrad): counter slot in g_tc_sched_ctr; 0 = static ranges
  int total_boxes;                                            // wgrad: K steps per tile
  int total_slabs, kreal;                                     // wgrad: valid 64-row slabs; real row length of dw
  int wrows;                                                  // real rows of w / dw (< Cout when the channels were padded)
  // temporal tap re-use (tc_conv_tr_kernel): a tap GROUP = the tr_kT temporal taps that share (kh,kw); taps[] then
  // lists groups (dt = first frame of the halo box, kidx = tap index of kt = 0)
  int tr_kT, tr_kstep, tr_rev;       // taps per group; tap-index step between kt and kt+1; 1 = tap kt reads frame offset kT-1-kt
  int tr_frame_bytes, tr_ablock_bytes;   // bytes of one frame of a block (R rows) / of one block's halo box ((BT+kT-1)*R rows)
  int tr_aslots, tr_bslots;
  // Temporal tap skipping (fprop / dgrad, unit temporal stride): a pixel box whose A frames for a temporal tap all lie
  // outside the A tensor multiplies pure zero fill — the video discriminator's data gradients (pT = 0, To = Ti - 3) spend
  // 23 % (dc2) to 43 % (dc4) of their K steps that way.  Taps are kt-major, so the taps worth issuing for a segment are
  // one contiguous run: tap-group index i (skip_per_kt taps each) reads frame t*a_mul_t + a_add_t + skip_dt0 + i*skip_dts.
  int skip_t, skip_per_kt[8], skip_nkt, skip_dt0, skip_dts, skip_aT;   // skip_per_kt: per stride class
  int str_rounds, str_rem_per; long long str_rem0;   // strided: full rounds of groups, then total - str_rem0 left-over units dealt str_rem_per per CTA
  int win, win_T;   // window-view A operand of the 3-channel layers (see TcWin): TMA coordinates (0, w, h, n*win_T + t, 0)
  // split-K (fprop of layers with too few pixel boxes to fill the machine, e.g. Dv.dc4: 18 boxes): unit = (k split, tile, box);
  // split ks multiplies taps [ks*ks_taps, (ks+1)*ks_taps) and red.adds its fp32 partial tile into a zeroed scratch tensor
  int ksplit, ks_taps;
  int strided;  // > 0: CTA c takes the unit groups c, c + grid, c + 2*grid, ... of `strided` units each (boxes cost
                // different numbers of K steps once taps are skipped; contiguous ranges would be frame-coherent and uneven)
  int box_tn;   // 1: boxes are numbered (w, h, n, t) — frames slowest — so the MT boxes of a step share their frame and skip alike
  int debug_skip_epi;
  int out_f32, ocols;                                        // ocols = channels of one output pixel row
  // depth-to-space epilogue (merged-class data gradient of the 3-channel layers): column ((e*sH + a)*sW + b)*C + ci of cell
  // (l, i, j) is input pixel (l*sT + e, i*sH + a, j*sW + b), channel ci; d2s_C = C (0 = off), d2s_* = the input tensor's extents
  int d2s_C, d2s_sT, d2s_sH, d2s_sW, d2s_Ti, d2s_Hi, d2s_Wi, d2s_cols;
  uint32_t d2s_tab[32];   // per column: e | a << 4 | b << 8 | ci << 12 (read with compile-time indices: the accumulators stay in registers)
  int store_cols;   // > 0 (bf16 row-major output): only the first store_cols columns (a multiple of 8) of a row are written, rows are store_cols wide
  int planar_chunk, planar_cols;   // fprop: write column c to plane c/chunk as [plane][pixel][chunk] (0 = row-major)
  long long planar_stride;
  uint8_t planar_plane[64], planar_within[64];   // per 4-column group c/4: plane index and offset inside the plane's chunk
  TcTap taps[64];
};

__device__ int g_tc_error = 0;

constexpr int kTcThreads = 320;      // warp 0 producer, warp 1 MMA issuer, warps 2-9 epilogue (two per TMEM lane quarter)
// The plain kernel can run a SECOND producer thread (warp 10; -DMCG_TC_PROD2=1): ncu's source page shows the MMA thread
// waiting for data 44 % of its samples while the producer is issuing, not waiting, 68 % of its own, which looked like an
// issue-bound producer.  Measured (A/B of two builds, every config-2 layer and the whole step): no difference — the
// kernels sit on the L2 -> SM ceiling (~13.5-14.4 TB/s delivered), not on the producer.  Kept, off.
#ifndef MCG_TC_PROD2
#define MCG_TC_PROD2 0
#endif
constexpr int kTcConvThreads = MCG_TC_PROD2 ? 352 : 320;
constexpr int kEpiThreads = 256;
constexpr int A_BYTES = 128 * 128;  // 128 rows x 64 bf16

struct TcSeg {
  int tile, k0, nk;    // K steps [k0, k0 + nk) of `tile` (fprop/dgrad: tile = row (cls, nt), all of its K steps)
  int box0, nlive;     // fprop/dgrad: first pixel box and number of 128-row blocks (<= MT) of this step
  int j0;              // fprop/dgrad: first tap (offset into the class's tap list) of this step — see TcParams::skip_t
};
// The same deterministic segment sequence is walked by the producer, the MMA issuer and the epilogue warps.
template <int MODE, int MT>
struct TcSegIter {
  long long u, u_end;
  __device__ __forceinline__ TcSegIter() : u(0), u_end(0) {}
  __device__ __forceinline__ explicit TcSegIter(const TcParams& P) {
    u = (long long)blockIdx.x * P.units_per_cta;
    u_end = u + P.units_per_cta;
    if (u_end > P.total_units) u_end = P.total_units;
  }
  __device__ __forceinline__ void set(long long a, long long b) { u = a; u_end = b; }
  __device__ __forceinline__ bool next(const TcParams& P, TcSeg& s) {
    if (u >= u_end) return false;
    if (MODE == kWgrad) {
      s.tile = (int)(u / P.total_boxes);
      s.k0 = (int)(u - (long long)s.tile * P.total_boxes);
      long long n = P.total_boxes - s.k0;
      if (n > u_end - u) n = u_end - u;
      s.nk = (int)n;
      s.box0 = 0; s.nlive = MT;
      u += n;
      return true;
    }
    long long ur = u;
    int ks = 0;
    if (P.ksplit > 1) {
      const long long per = (long long)P.ntiles * P.nboxes;
      ks = (int)(u / per);
      ur = u - ks * per;
    }
    s.tile = (int)(ur / P.nboxes);
    s.box0 = (int)(ur - (long long)s.tile * P.nboxes);
    long long n = P.nboxes - s.box0;
    if (n > u_end - u) n = u_end - u;
    if (n > MT) n = MT;
    s.nlive = (int)n;
    s.k0 = 0;
    s.j0 = 0;
    s.nk = P.tap_count[s.tile / P.ntn] * P.chunks;
    if (P.ksplit > 1) { s.j0 = ks * P.ks_taps; s.nk = P.ks_taps * P.chunks; }
    if (P.skip_t) {
      // frames covered by the step's boxes.  box index = ((n*nbt + t)*nbh + h)*nbw + w: a step that crosses into the next
      // sample keeps every tap; with box_tn = ((t*nbb + n)*nbh + h)*nbw + w: frames never wrap inside a step
      const int per_t = P.nbw * P.nbh, b_last = s.box0 + (int)n - 1;
      const int q0 = s.box0 / per_t, q1 = b_last / per_t;
      if (P.box_tn || q0 / P.nbt == q1 / P.nbt) {
        const int t_first = P.box_tn ? q0 / P.nbb : q0 % P.nbt, t_last = P.box_tn ? q1 / P.nbb : q1 % P.nbt;
        const int a_lo = t_first * P.BT * P.a_mul_t + P.a_add_t + P.skip_dt0;
        const int a_hi = (t_last * P.BT + P.BT - 1) * P.a_mul_t + P.a_add_t + P.skip_dt0;
        int i_lo, i_hi;   // tap groups i with [a_lo, a_hi] + i*dts meeting [0, aT)
        if (P.skip_dts > 0) { i_lo = -a_hi; i_hi = P.skip_aT - 1 - a_lo; }
        else { i_lo = a_lo - (P.skip_aT - 1); i_hi = a_hi; }
        if (i_lo < 0) i_lo = 0;
        if (i_hi > P.skip_nkt - 1) i_hi = P.skip_nkt - 1;
        const int groups = i_hi >= i_lo ? i_hi - i_lo + 1 : 0;
        const int per_kt = P.skip_per_kt[s.tile / P.ntn];
        s.j0 = groups ? i_lo * per_kt : 0;
        s.nk = groups * per_kt * P.chunks;
      }
    }
    u += n;
    return true;
  }
};

// ---- dynamic work distribution -------------------------------------------------------------------------------
// A static one-range-per-CTA split assumes all CTAs of the grid start together.  They do not when another stream's
// kernel (a small-grid convolution, a collective) holds some SMs: the CTAs that start late finish late and the whole
// kernel takes up to twice as long.  With P.dyn the CTAs of a fprop/dgrad launch instead DRAW ranges of units from a
// global counter (guided self-scheduling: remaining / (2 x grid), never less than one tile step), so a late CTA simply
// draws less.  The producer thread draws (one range ahead, so the atomic's latency hides under the loads of the current
// range) and publishes each range to the MMA and epilogue roles through a small shared-memory ring.
constexpr int kSchedDepth = 8;
constexpr int kSchedSlots = 1024;
__device__ int g_tc_sched_ctr[kSchedSlots];    // units handed out so far; the last CTA to leave resets its slot
__device__ int g_tc_sched_done[kSchedSlots];
struct TcSchedSmem {
  int u0[kSchedDepth], n[kSchedDepth];
  int first_u0;                                // drawn by thread 0 on kernel entry, under the prologue
};
__device__ __forceinline__ int tc_sched_chunk(int total, int handed_out, int mt) {
  int rem = total - handed_out;
  if (rem < 0) rem = 0;
  int c = rem / (2 * (int)gridDim.x);
  c -= c % mt;
  return c < mt ? mt : c;
}
// Consumer side (MMA issuer, epilogue threads); in static mode: the CTA's one range, no shared-memory traffic.
struct TcRanges {
  uint32_t full_a, empty_a;
  const TcSchedSmem* sm;
  int slot, grp;
  uint32_t ph;
  bool dyn, done;
  __device__ __forceinline__ TcRanges(const TcParams& P, const TcSchedSmem* sm_, uint32_t full, uint32_t empty)
      : full_a(full), empty_a(empty), sm(sm_), slot(0), grp(0), ph(0), dyn(P.dyn != 0), done(false) {}
  __device__ __forceinline__ bool next(const TcParams& P, long long& u, long long& u_end, int* err) {
    if (!dyn) {
      if (P.strided) {
        if (grp < P.str_rounds) {
          u = ((long long)blockIdx.x + (long long)grp * gridDim.x) * P.strided;
          u_end = u + P.strided;
        } else if (grp == P.str_rounds) {   // what the full rounds left over, cut evenly (unit, not group, granularity)
          u = P.str_rem0 + (long long)blockIdx.x * P.str_rem_per;
          u_end = u + P.str_rem_per;
          if (u_end > P.total_units) u_end = P.total_units;
        } else {
          return false;
        }
        ++grp;
        return u < u_end;
      }
      if (done) return false;
      done = true;
      u = (long long)blockIdx.x * P.units_per_cta;
      u_end = u + P.units_per_cta;
      if (u_end > P.total_units) u_end = P.total_units;
      return u < u_end;
    }
    if (!mbar_wait_a(full_a + slot * 8, ph, err)) return false;
    const int a = *(volatile const int*)&sm->u0[slot], c = *(volatile const int*)&sm->n[slot];
    mbar_arrive_a(empty_a + slot * 8);
    if (++slot == kSchedDepth) { slot = 0; ph ^= 1; }
    if (c <= 0) return false;
    u = a; u_end = (long long)a + c;
    return true;
  }
};

struct TcBox { int w0, h0, t0, n0; };
__device__ __forceinline__ TcBox tc_decode_box(const TcParams& P, int bi) {
  TcBox b;
  b.w0 = (bi % P.nbw) * P.BW; bi /= P.nbw;
  b.h0 = (bi % P.nbh) * P.BH; bi /= P.nbh;
  if (P.box_tn) {
    b.n0 = (bi % P.nbb) * P.BB; bi /= P.nbb;
    b.t0 = bi * P.BT;
  } else {
    b.t0 = (bi % P.nbt) * P.BT; bi /= P.nbt;
    b.n0 = bi * P.BB;
  }
  return b;
}

// The boxes of one step are consecutive: the first is decoded with divisions, the others follow by carrying.  (The
// epilogue of the short-K launches — the 3-channel layers' GEMMs, the generator's 2-D layers — decoded every box from
// scratch, with the divisors re-read from the parameter bank each time: ~3,500 cycles per 128-row block against ~500 of
// actual work, ncu source page of Dv.dc1's fprop.)
struct TcBoxDims {
  int BW, BH, BT, BB, wl, hl, tl, nl, tn;
  __device__ __forceinline__ explicit TcBoxDims(const TcParams& P)
      : BW(P.BW), BH(P.BH), BT(P.BT), BB(P.BB), wl(P.nbw * P.BW), hl(P.nbh * P.BH), tl(P.nbt * P.BT), nl(P.nbb * P.BB), tn(P.box_tn) {}
  __device__ __forceinline__ void next(TcBox& b) const {
    b.w0 += BW;
    if (b.w0 < wl) return;
    b.w0 = 0; b.h0 += BH;
    if (b.h0 < hl) return;
    b.h0 = 0;
    if (tn) {
      b.n0 += BB;
      if (b.n0 >= nl) { b.n0 = 0; b.t0 += BT; }
    } else {
      b.t0 += BT;
      if (b.t0 >= tl) { b.t0 = 0; b.n0 += BB; }
    }
  }
};

// Epilogue role (warps 2-9), shared by the kernels below: walks the same segment sequence as the producer / MMA roles,
// waits for an accumulator buffer, moves TMEM -> registers -> global, and hands the buffer back.
// A lone warp per scheduler runs the epilogue's dependent instruction chains at a fraction of the issue rate, so two
// warps share each TMEM lane quarter and take alternate 32-column chunks.
template <int MODE, int BN, int MT, int NBUF>
__device__ __forceinline__ void tc_epilogue_role(const TcParams& P, void* __restrict__ out, const float* __restrict__ bias,
                                                 uint32_t tmem, uint32_t tfull_a, uint32_t tempty_a, int warp, int lane, int* err,
                                                 TcRanges rng) {
  constexpr int ACC_COLS = MT * BN;
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int r = q * 32 + lane;  // accumulator row == TMEM lane
    TcSegIter<MODE, MT> iter;
    TcSeg sg;
    uint32_t seg = 0;
    long long ru, ru_end;
    bool alive = true;
    // per-thread constants of the fprop / dgrad epilogue: this row's pixel inside a box, the box walker, output geometry
    const TcBoxDims bd(P);
    int e_iw = 0, e_ih = 0, e_it = 0, e_ib = 0;
    if (MODE != kWgrad) {
      int rr = r;
      e_iw = rr % P.BW; rr /= P.BW;
      e_ih = rr % P.BH; rr /= P.BH;
      e_it = rr % P.BT; rr /= P.BT;
      e_ib = rr;
    }
    const int e_mw = P.o_mul_w, e_mh = P.o_mul_h, e_mt = P.o_mul_t, e_fw = P.full_w, e_fh = P.full_h, e_ft = P.full_t, e_en = P.EN;
    const long long e_sn = P.os_n, e_st = P.os_t, e_sh = P.os_h, e_sw = P.os_w;
    while (alive && rng.next(P, ru, ru_end, err)) {
    iter.set(ru, ru_end);
    while (iter.next(P, sg)) {
      const int nt = sg.tile % P.ntn, mg = sg.tile / P.ntn;
      const int ncol0 = nt * BN;
      const bool have_acc = sg.nk > 0;
      const uint32_t buf = seg % NBUF;
      if (have_acc) {
        if (!mbar_wait_a(tfull_a + buf * 8, (seg / NBUF) & 1, err)) { alive = false; break; }
        tc_fence_after();
      }
      const uint32_t acc = tmem + (uint32_t(q * 32) << 16) + buf * ACC_COLS;
      if (MODE != kWgrad) {
        const int cls = mg;
        const int pw = cls % P.cls_w, phh = (cls / P.cls_w) % P.cls_h, pt = cls / (P.cls_w * P.cls_h);
        const int iw = e_iw, ih = e_ih, itt = e_it, ib = e_ib;
        TcBox bx = tc_decode_box(P, sg.box0);
#pragma unroll 1
        for (int m = 0; m < sg.nlive; ++m) {
          if (m) bd.next(bx);
          const int ow = (bx.w0 + iw) * e_mw + pw, oh = (bx.h0 + ih) * e_mh + phh, ot = (bx.t0 + itt) * e_mt + pt;
          const int on = bx.n0 + ib;
          const bool valid = ow < e_fw && oh < e_fh && ot < e_ft && on < e_en;
          const long long base = (long long)on * e_sn + (long long)ot * e_st + (long long)oh * e_sh + (long long)ow * e_sw + ncol0;
#pragma unroll 1
          for (int c0 = half * 32; c0 < BN; c0 += 64) {
            uint32_t v[32];
            if (have_acc) {
              tmem_ld32(acc + m * BN + c0, v);
              tmem_ld_wait();
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = 0u;
            }
            float f[32];
            const bool colbias = bias && !(MODE == kFprop && P.d2s_C);
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) + (colbias ? bias[ncol0 + c0 + i] : 0.f);
            if (valid) {
              if (MODE == kFprop && P.d2s_C) {
                // the cell's sT*sH*sW*C values go straight to their pixels of dx (no Zm matrix, no depth-to-space pass);
                // all of them sit in the first 32 columns, so only the half-0 warps store
                if (c0 == 0) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) {
                    if (i < P.d2s_cols) {
                      const uint32_t tb = P.d2s_tab[i];
                      const int ti = ot * P.d2s_sT + (int)(tb & 15), hi = oh * P.d2s_sH + (int)((tb >> 4) & 15);
                      const int wi = ow * P.d2s_sW + (int)((tb >> 8) & 15), ci = (int)(tb >> 12);
                      if (ti < P.d2s_Ti && hi < P.d2s_Hi && wi < P.d2s_Wi) {
                        const long long dst = ((((long long)on * P.d2s_Ti + ti) * P.d2s_Hi + hi) * P.d2s_Wi + wi) * P.d2s_C + ci;
                        const float val = f[i] + (bias ? bias[ci] : 0.f);
                        if (P.out_f32) reinterpret_cast<float*>(out)[dst] = val;
                        else reinterpret_cast<__nv_bfloat16*>(out)[dst] = __float2bfloat16_rn(val);
                      }
                    }
                  }
                }
              } else if (MODE == kFprop && P.planar_chunk) {
                // planar bf16 output for the narrow-Cin dgrad GEMM: plane = (kt,kh) run, so col2im reads contiguous lines
                __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                  const int c = ncol0 + c0 + i;
                  if (c < P.planar_cols) {
                    const int plane = P.planar_plane[c >> 2], within = P.planar_within[c >> 2];
                    uint2 u;
                    __nv_bfloat162 lo = __floats2bfloat162_rn(f[i], f[i + 1]), hi = __floats2bfloat162_rn(f[i + 2], f[i + 3]);
                    u.x = *reinterpret_cast<uint32_t*>(&lo);
                    u.y = *reinterpret_cast<uint32_t*>(&hi);
                    *reinterpret_cast<uint2*>(o + plane * P.planar_stride + (long long)ow * P.planar_chunk + within) = u;
                  }
                }
              } else if (MODE == kFprop && P.ksplit > 1) {
                // split-K partial tile: fp32 red.add into the scratch tensor (same strides as y); bias and the cast to
                // the output type follow in splitk_finish_kernel
                float* o = reinterpret_cast<float*>(out) + base + c0;
#pragma unroll
                for (int i = 0; i < 32; ++i) atomicAdd(o + i, f[i]);
              } else if (P.out_f32) {
                float* o = reinterpret_cast<float*>(out) + base + c0;
#pragma unroll
                for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
              } else {
                __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + base + c0;
                const int lim = P.store_cols ? P.store_cols - (ncol0 + c0) : 32;   // columns beyond store_cols are not written
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                  if (i >= lim) break;
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
        }
      } else {
        float* dw = reinterpret_cast<float*>(out);
#pragma unroll 1
        for (int m = 0; m < MT; ++m) {
          const int u = (mg * MT + m) * 2 + (r >> 6);
          const bool live = u < P.total_slabs;
          const TcTap tp = P.taps[live ? u / P.chunks : 0];
          const long long kidx = (long long)tp.kidx * P.Cin + (u % P.chunks) * 64 + (r & 63);
          const bool row_ok = live && kidx < P.kreal && have_acc && !P.debug_skip_epi;
#pragma unroll 1
          for (int c0 = half * 32; c0 < BN; c0 += 64) {
            uint32_t v[32];
            if (have_acc) {
              tmem_ld32(acc + m * BN + c0, v);
              tmem_ld_wait();
            }
            if (row_ok) {
              float* dst = dw + (long long)(ncol0 + c0) * P.kreal + kidx;
              if (ncol0 + c0 + 32 <= P.wrows) {   // the common case, warp-uniform: no per-column test
#pragma unroll
                for (int i = 0; i < 32; ++i) atomicAdd(dst + (long long)i * P.kreal, __uint_as_float(v[i]));
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (ncol0 + c0 + i < P.wrows) atomicAdd(dst + (long long)i * P.kreal, __uint_as_float(v[i]));
              }
            }
          }
        }
      }
      if (have_acc) {
        tc_fence_before();
        mbar_arrive_a(tempty_a + buf * 8);   // 128 arrivals release the accumulator buffer to the MMA issuer
        ++seg;
      }
    }
    }
}

template <int MODE, int BN, int MT, int STAGES>
__global__ void __launch_bounds__(kTcConvThreads, 1) tc_conv_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                const __grid_constant__ CUtensorMap mapB,
                                                                const __grid_constant__ TcParams P, void* __restrict__ out,
                                                                const float* __restrict__ bias) {
  pdl_launch_dependents();
  constexpr int B_BYTES = BN * 128;
  constexpr int STAGE_BYTES = MT * A_BYTES + B_BYTES;
  constexpr int ACC_COLS = MT * BN;                       // fp32 accumulator columns of one tile
  constexpr int NBUF = (2 * ACC_COLS <= 512) ? 2 : 1;     // double-buffered accumulators when they fit in the 512 TMEM columns
  constexpr int TMEM_NEED = NBUF * ACC_COLS;
  constexpr int TMEM_COLS = TMEM_NEED <= 32 ? 32 : (TMEM_NEED <= 64 ? 64 : (TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512)));
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tfull_bar[2], tempty_bar[2];
  __shared__ uint64_t sfull_bar[kSchedDepth], sempty_bar[kSchedDepth];
  __shared__ TcSchedSmem sched;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* err = &g_tc_error;

  if (threadIdx.x == 0) {
    // the CTA's first range is drawn here, so the atomic's round trip hides under the rest of the prologue
    if (P.dyn) sched.first_u0 = atomicAdd(&g_tc_sched_ctr[P.sched_slot], tc_sched_chunk((int)P.total_units, 0, MT));
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], kEpiThreads); }
    for (int i = 0; i < kSchedDepth; ++i) { mbar_init(&sfull_bar[i], 1); mbar_init(&sempty_bar[i], 1 + kEpiThreads); }
    fence_barrier_init();
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 1) { tmem_alloc(&tmem_slot, TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  pdl_wait();   // everything above overlapped the predecessor's tail; no global / TMA access before this line

  const uint32_t smem_a = smem_u32(smem);
  const uint32_t full_a = smem_u32(&full_bar[0]), empty_a = smem_u32(&empty_bar[0]);
  const uint32_t tfull_a = smem_u32(&tfull_bar[0]), tempty_a = smem_u32(&tempty_bar[0]);
  const uint32_t sfull_a = smem_u32(&sfull_bar[0]), sempty_a = smem_u32(&sempty_bar[0]);

  if (warp == 0 || warp == 10) {
    // ================================================= TMA producer =========================================
    // One elected thread per producer warp; the loop is instruction-bound, so everything per K step is strength-reduced:
    // no divisions, tap / slab decode hoisted, barrier and stage addresses advanced incrementally.  With two producers
    // (static schedules only) both walk the whole step sequence and each issues the steps of its own parity.
    const uint32_t nprod = (MCG_TC_PROD2 && !P.dyn) ? 2u : 1u, prod = warp == 0 ? 0u : 1u;
    if (prod < nprod && elect_one()) {
      const uint64_t map_a = reinterpret_cast<uint64_t>(&mapA), map_b = reinterpret_cast<uint64_t>(&mapB);
      TcSegIter<MODE, MT> iter;
      TcSeg sg;
      uint32_t kstep = 0;   // ring steps issued by either producer so far
      const TcBoxDims pbd(P);
      int s = 0;
      uint32_t ph = 1;   // ring slot and the parity its `empty` barrier is waited with
      uint32_t stage_a = smem_a, full_s = full_a, empty_s = empty_a;
      bool alive = true;
      // work ranges: static = the CTA's one range; dynamic = drawn from the launch's counter, one range ahead
      const int total = (int)P.total_units;
      int cur_u0 = 0, cur_c = 0, nxt_u0 = 0, nxt_c = 0, pslot = 0;
      uint32_t pph = 1;
      bool static_done = false;
      if (P.dyn) { nxt_c = tc_sched_chunk(total, 0, MT); nxt_u0 = *(volatile int*)&sched.first_u0; }
      for (;;) {
        if (!alive) break;
        if (P.dyn) {
          cur_u0 = nxt_u0; cur_c = nxt_c;
          int n = total - cur_u0;
          if (n > cur_c) n = cur_c;
          if (n < 0) n = 0;
          if (!mbar_wait_a(sempty_a + pslot * 8, pph, err)) break;
          sched.u0[pslot] = cur_u0; sched.n[pslot] = n;
          mbar_arrive_a(sfull_a + pslot * 8);          // release: the two stores above are visible to the waiters
          if (++pslot == kSchedDepth) { pslot = 0; pph ^= 1; }
          if (n == 0) break;
          nxt_c = tc_sched_chunk(total, cur_u0 + cur_c, MT);
          nxt_u0 = atomicAdd(&g_tc_sched_ctr[P.sched_slot], nxt_c);   // consumed when this range's loads are issued
          iter.set(cur_u0, (long long)cur_u0 + n);
        } else if (P.strided) {
          long long gu, ge;                           // cur_c counts groups here (same sequence as TcRanges::next)
          if (cur_c < P.str_rounds) {
            gu = ((long long)blockIdx.x + (long long)cur_c * gridDim.x) * P.strided;
            ge = gu + P.strided;
          } else if (cur_c == P.str_rounds) {
            gu = P.str_rem0 + (long long)blockIdx.x * P.str_rem_per;
            ge = gu + P.str_rem_per < P.total_units ? gu + P.str_rem_per : P.total_units;
          } else {
            break;
          }
          ++cur_c;
          if (gu >= ge) break;
          iter.set(gu, ge);
        } else {
          if (static_done) break;
          static_done = true;
          iter = TcSegIter<MODE, MT>(P);
        }
      while (alive && iter.next(P, sg)) {
        const int nt = sg.tile % P.ntn, mg = sg.tile / P.ntn;
        const int ncol0 = nt * BN;
        if (MODE != kWgrad) {
          const int cls = mg;
          TcBox bx[MT];
          {
            TcBox cur = tc_decode_box(P, sg.box0);
#pragma unroll
            for (int m = 0; m < MT; ++m) {
              if (m && m < sg.nlive) pbd.next(cur);
              bx[m].w0 = cur.w0 * P.a_mul_w + P.a_add_w;
              bx[m].h0 = cur.h0 * P.a_mul_h + P.a_add_h;
              bx[m].t0 = cur.t0 * P.a_mul_t + P.a_add_t;
              bx[m].n0 = cur.n0;
            }
          }
          const uint32_t tx_bytes = sg.nlive * A_BYTES + B_BYTES;
          const int chunks = P.chunks;
          const int j_beg = P.tap_begin[cls] + sg.j0, j_end = j_beg + sg.nk / chunks;
          for (int j = j_beg; alive && j < j_end; ++j) {
            const TcTap tp = P.taps[j];
            const int kcol = tp.kidx * P.Cin;
            for (int c = 0; c < chunks; ++c) {
              if ((kstep++ & (nprod - 1)) == prod) {
              if (!mbar_wait_a(empty_s, ph, err)) { alive = false; break; }
              mbar_expect_tx_a(full_s, tx_bytes);
#pragma unroll
              for (int m = 0; m < MT; ++m)
                if (m < sg.nlive) {
                  if (MODE == kFprop && P.win)
                    tma_load_5d_a(stage_a + m * A_BYTES, map_a, full_s, 0, bx[m].w0, bx[m].h0,
                                  bx[m].n0 * P.win_T + bx[m].t0 + tp.dt, 0);
                  else
                    tma_load_5d_a(stage_a + m * A_BYTES, map_a, full_s, c * 64, bx[m].w0 + tp.dw, bx[m].h0 + tp.dh,
                                  bx[m].t0 + tp.dt, bx[m].n0);
                }
              if (MODE == kFprop) {
                tma_load_2d_a(stage_a + MT * A_BYTES, map_b, full_s, kcol + c * 64, ncol0);
              } else {
#pragma unroll
                for (int sl = 0; sl < BN / 64; ++sl)
                  tma_load_2d_a(stage_a + MT * A_BYTES + sl * 8192, map_b, full_s, kcol + ncol0 + sl * 64, c * 64);
              }
              }
              ++s; stage_a += STAGE_BYTES; full_s += 8; empty_s += 8;
              if (s == STAGES) { s = 0; ph ^= 1; stage_a = smem_a; full_s = full_a; empty_s = empty_a; }
            }
          }
        } else {
          const int u0 = mg * MT * 2;  // first of the tile's MT*2 slabs = (tap, chunk) pairs
          int sc[MT * 2], sw[MT * 2], sh[MT * 2], st_[MT * 2];
#pragma unroll
          for (int sl = 0; sl < MT * 2; ++sl) {
            const int u = u0 + sl;
            const bool live = u < P.total_slabs;  // a dummy slab's channel coordinate is out of bounds -> zero fill
            const TcTap tp = P.taps[live ? u / P.chunks : 0];
            sc[sl] = live ? (u % P.chunks) * 64 : P.Cin;
            sw[sl] = P.a_add_w + tp.dw; sh[sl] = P.a_add_h + tp.dh; st_[sl] = P.a_add_t + tp.dt;
          }
          TcBox pb = tc_decode_box(P, sg.k0);
          const int wlim = P.nbw * P.BW, hlim = P.nbh * P.BH, tlim = P.nbt * P.BT;
          const int BWs = P.BW, BHs = P.BH, BTs = P.BT, BBs = P.BB, mw = P.a_mul_w, mh = P.a_mul_h, mt_ = P.a_mul_t;
          for (int kb = 0; kb < sg.nk; ++kb) {
            if ((kstep++ & (nprod - 1)) == prod) {
            if (!mbar_wait_a(empty_s, ph, err)) { alive = false; break; }
            mbar_expect_tx_a(full_s, STAGE_BYTES);
            const int aw = pb.w0 * mw, ah = pb.h0 * mh, at = pb.t0 * mt_;
#pragma unroll
            for (int sl = 0; sl < MT * 2; ++sl) {
              if (P.win)      // a dummy slab (sc = Cin) asks for a frame past the end of the tensor: zero fill
                tma_load_5d_a(stage_a + sl * 8192, map_a, full_s, 0, pb.w0, pb.h0,
                              sc[sl] < P.Cin ? pb.n0 * P.win_T + pb.t0 + st_[sl] : 0x3fffffff, 0);
              else
                tma_load_5d_a(stage_a + sl * 8192, map_a, full_s, sc[sl], aw + sw[sl], ah + sh[sl], at + st_[sl], pb.n0);
            }
#pragma unroll
            for (int sl = 0; sl < BN / 64; ++sl)
              tma_load_5d_a(stage_a + MT * A_BYTES + sl * 8192, map_b, full_s, ncol0 + sl * 64, pb.w0, pb.h0, pb.t0, pb.n0);
            }
            ++s; stage_a += STAGE_BYTES; full_s += 8; empty_s += 8;
            if (s == STAGES) { s = 0; ph ^= 1; stage_a = smem_a; full_s = full_a; empty_s = empty_a; }
            pb.w0 += BWs;   // next pixel box, carried by hand
            if (pb.w0 >= wlim) {
              pb.w0 = 0; pb.h0 += BHs;
              if (pb.h0 >= hlim) {
                pb.h0 = 0; pb.t0 += BTs;
                if (pb.t0 >= tlim) { pb.t0 = 0; pb.n0 += BBs; }
              }
            }
          }
        }
      }
      }
    }
  } else if (warp == 1) {
    // ================================================= MMA issuer ===========================================
    if (elect_one()) {
      constexpr int A_MN = (MODE == kWgrad), B_MN = (MODE != kFprop);
      constexpr uint32_t A_KSTEP = (A_MN ? 2048 : 32) >> 4, B_KSTEP = (B_MN ? 2048 : 32) >> 4;   // per K=16 slice, in 16-B units
      const uint32_t idesc = make_idesc_bf16(128, BN, A_MN, B_MN);
      const uint32_t desc_hi = smem_desc_hi(1024);
      const uint32_t a_lo0 = smem_desc_lo(smem_a, A_MN ? 8192 : 16), b_lo0 = smem_desc_lo(smem_a + MT * A_BYTES, B_MN ? 8192 : 16);
      TcSegIter<MODE, MT> iter;
      TcRanges rng(P, &sched, sfull_a, sempty_a);
      TcSeg sg;
      uint32_t seg = 0, ph = 0;
      int s = 0;
      uint32_t a_lo = a_lo0, b_lo = b_lo0, full_s = full_a, empty_s = empty_a;
      bool alive = true;
      long long ru, ru_end;
      while (alive && rng.next(P, ru, ru_end, err)) {
      iter.set(ru, ru_end);
      while (alive && iter.next(P, sg)) {
        if (sg.nk == 0) continue;
        const uint32_t buf = seg % NBUF;
        if (!mbar_wait_a(tempty_a + buf * 8, ((seg / NBUF) & 1) ^ 1, err)) { alive = false; break; }
        tc_fence_after();
        const uint32_t acc = tmem + buf * ACC_COLS;
        const int nlive = sg.nlive;
        for (int kb = 0; kb < sg.nk; ++kb) {
          if (!mbar_wait_a(full_s, ph, err)) { alive = false; break; }
          tc_fence_after();
#pragma unroll
          for (int m = 0; m < MT; ++m) {
            if (m >= nlive) break;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_lh(acc + m * BN, a_lo + m * (A_BYTES >> 4) + k * A_KSTEP, desc_hi, b_lo + k * B_KSTEP, desc_hi, idesc,
                           (kb | k) ? 1u : 0u);
          }
          umma_commit_a(empty_s);
          ++s; a_lo += STAGE_BYTES >> 4; b_lo += STAGE_BYTES >> 4; full_s += 8; empty_s += 8;
          if (s == STAGES) { s = 0; ph ^= 1; a_lo = a_lo0; b_lo = b_lo0; full_s = full_a; empty_s = empty_a; }
        }
        if (alive) umma_commit_a(tfull_a + buf * 8);
        ++seg;
      }
      }
    }
  } else {
    // ================================================= epilogue ============================================
    tc_epilogue_role<MODE, BN, MT, NBUF>(P, out, bias, tmem, tfull_a, tempty_a, warp, lane, err,
                                         TcRanges(P, &sched, sfull_a, sempty_a));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, TMEM_COLS);
  if (P.dyn && threadIdx.x == 0) {
    // every draw of this CTA has returned; the last CTA to leave hands the counter slot back zeroed (the next launch
    // that uses the slot — the same graph node one replay later — is ordered after this kernel)
    __threadfence();
    if (atomicAdd(&g_tc_sched_done[P.sched_slot], 1) == (int)gridDim.x - 1) {
      g_tc_sched_ctr[P.sched_slot] = 0;
      g_tc_sched_done[P.sched_slot] = 0;
      __threadfence();
    }
  }
}

// =============================================================================================================
// Temporal tap re-use (video discriminator, kT = 4, sT = 1).  The kT taps that share (kh, kw) read the SAME pixels of kT
// consecutive frame windows, so ONE TMA box of BT + kT - 1 frames (pixel box (BW, BH, BT, 1), R = BW*BH rows per frame,
// R % 8 == 0) serves all of them: tap kt is the same shared-memory tile entered R*128*offset(kt) bytes further in — a
// whole number of 1024-byte swizzle atoms, so only the descriptor start address moves.  A traffic from L2 drops by
// kT*BT / (BT + kT - 1) (1.6x for BT = 2, 2.3x for BT = 4, 2.9x for BT = 8), which is what bounds the N = 64 / 128
// layers.  Two rings: halo boxes (one per group x 64-channel chunk, MT blocks) and weight tiles (one per K step).
// =============================================================================================================
constexpr int kTrMaxASlots = 4, kTrMaxBSlots = 8;
template <int MODE, int BN, int MT>
__global__ void __launch_bounds__(kTcThreads, 1) tc_conv_tr_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                   const __grid_constant__ CUtensorMap mapB,
                                                                   const __grid_constant__ TcParams P, void* __restrict__ out,
                                                                   const float* __restrict__ bias) {
  pdl_launch_dependents();
  static_assert(MODE == kFprop || MODE == kDgrad, "temporal re-use is for fprop / dgrad");
  constexpr int B_BYTES = BN * 128;
  constexpr int ACC_COLS = MT * BN;
  constexpr int NBUF = (2 * ACC_COLS <= 512) ? 2 : 1;
  constexpr int TMEM_NEED = NBUF * ACC_COLS;
  constexpr int TMEM_COLS = TMEM_NEED <= 32 ? 32 : (TMEM_NEED <= 64 ? 64 : (TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512)));
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t fullA[kTrMaxASlots], emptyA[kTrMaxASlots], fullB[kTrMaxBSlots], emptyB[kTrMaxBSlots], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* err = &g_tc_error;
  const int ASLOTS = P.tr_aslots, BSLOTS = P.tr_bslots;
  const int aslot_bytes = MT * P.tr_ablock_bytes;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kTrMaxASlots; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
    for (int i = 0; i < kTrMaxBSlots; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], kEpiThreads); }
    fence_barrier_init();
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 1) { tmem_alloc(&tmem_slot, TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  pdl_wait();   // everything above overlapped the predecessor's tail; no global / TMA access before this line
  const uint32_t ringA = smem_u32(smem), ringB = ringA + ASLOTS * aslot_bytes;
  const uint32_t fullA_a = smem_u32(&fullA[0]), emptyA_a = smem_u32(&emptyA[0]);
  const uint32_t fullB_a = smem_u32(&fullB[0]), emptyB_a = smem_u32(&emptyB[0]);
  const uint32_t tfull_a = smem_u32(&tfull_bar[0]), tempty_a = smem_u32(&tempty_bar[0]);
  const int KT = P.tr_kT;

  if (warp == 0) {
    // ================================================= TMA producer =========================================
    if (elect_one()) {
      const uint64_t map_a = reinterpret_cast<uint64_t>(&mapA), map_b = reinterpret_cast<uint64_t>(&mapB);
      TcSegIter<MODE, MT> iter(P);
      TcSeg sg;
      int sa = 0, sb = 0;
      uint32_t pha = 1, phb = 1;
      bool alive = true;
      while (alive && iter.next(P, sg)) {
        const int nt = sg.tile % P.ntn, cls = sg.tile / P.ntn;
        const int ncol0 = nt * BN;
        TcBox bx[MT];
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          bx[m] = tc_decode_box(P, sg.box0 + (m < sg.nlive ? m : 0));
          bx[m].w0 = bx[m].w0 * P.a_mul_w + P.a_add_w;
          bx[m].h0 = bx[m].h0 * P.a_mul_h + P.a_add_h;
          bx[m].t0 = bx[m].t0 * P.a_mul_t + P.a_add_t;
        }
        const uint32_t txa = sg.nlive * P.tr_ablock_bytes;
        const int j_end = P.tap_begin[cls] + P.tap_count[cls], chunks = P.chunks;
        for (int j = P.tap_begin[cls]; alive && j < j_end; ++j) {
          const TcTap tp = P.taps[j];
          for (int c = 0; alive && c < chunks; ++c) {
            if (!mbar_wait_a(emptyA_a + sa * 8, pha, err)) { alive = false; break; }
            mbar_expect_tx_a(fullA_a + sa * 8, txa);
#pragma unroll
            for (int m = 0; m < MT; ++m)
              if (m < sg.nlive)
                tma_load_5d_a(ringA + sa * aslot_bytes + m * P.tr_ablock_bytes, map_a, fullA_a + sa * 8, c * 64, bx[m].w0 + tp.dw,
                              bx[m].h0 + tp.dh, bx[m].t0 + tp.dt, bx[m].n0);
            if (++sa == ASLOTS) { sa = 0; pha ^= 1; }
            for (int kt = 0; kt < KT; ++kt) {
              if (!mbar_wait_a(emptyB_a + sb * 8, phb, err)) { alive = false; break; }
              mbar_expect_tx_a(fullB_a + sb * 8, B_BYTES);
              const int kcol = (tp.kidx + kt * P.tr_kstep) * P.Cin;
              if (MODE == kFprop) {
                tma_load_2d_a(ringB + sb * B_BYTES, map_b, fullB_a + sb * 8, kcol + c * 64, ncol0);
              } else {
#pragma unroll
                for (int sl = 0; sl < BN / 64; ++sl)
                  tma_load_2d_a(ringB + sb * B_BYTES + sl * 8192, map_b, fullB_a + sb * 8, kcol + ncol0 + sl * 64, c * 64);
              }
              if (++sb == BSLOTS) { sb = 0; phb ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================= MMA issuer ===========================================
    if (elect_one()) {
      constexpr int B_MN = (MODE != kFprop);
      constexpr uint32_t B_KSTEP = (B_MN ? 2048 : 32) >> 4;
      const uint32_t idesc = make_idesc_bf16(128, BN, 0, B_MN);
      const uint32_t desc_hi = smem_desc_hi(1024);
      const uint32_t a_lo0 = smem_desc_lo(ringA, 16), b_lo0 = smem_desc_lo(ringB, B_MN ? 8192 : 16);
      const uint32_t ablock16 = P.tr_ablock_bytes >> 4, aslot16 = aslot_bytes >> 4, frame16 = P.tr_frame_bytes >> 4;
      TcSegIter<MODE, MT> iter(P);
      TcSeg sg;
      uint32_t seg = 0, pha = 0, phb = 0;
      int sa = 0, sb = 0;
      bool alive = true;
      while (alive && iter.next(P, sg)) {
        const int cls = sg.tile / P.ntn;
        const int ngc = P.tap_count[cls] * P.chunks;   // group x chunk steps
        if (ngc == 0) continue;
        const uint32_t buf = seg % NBUF;
        if (!mbar_wait_a(tempty_a + buf * 8, ((seg / NBUF) & 1) ^ 1, err)) break;
        tc_fence_after();
        const uint32_t acc = tmem + buf * ACC_COLS;
        const int nlive = sg.nlive;
        uint32_t first = 1;
        for (int gc = 0; alive && gc < ngc; ++gc) {
          if (!mbar_wait_a(fullA_a + sa * 8, pha, err)) { alive = false; break; }
          tc_fence_after();
          const uint32_t a_slot = a_lo0 + sa * aslot16;
          for (int kt = 0; kt < KT; ++kt) {
            if (!mbar_wait_a(fullB_a + sb * 8, phb, err)) { alive = false; break; }
            tc_fence_after();
            const uint32_t a_tap = a_slot + (P.tr_rev ? (KT - 1 - kt) : kt) * frame16;
            const uint32_t b_lo = b_lo0 + sb * (B_BYTES >> 4);
#pragma unroll
            for (int m = 0; m < MT; ++m) {
              if (m >= nlive) break;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_lh(acc + m * BN, a_tap + m * ablock16 + k * 2, desc_hi, b_lo + k * B_KSTEP, desc_hi, idesc,
                             (first && k == 0) ? 0u : 1u);
            }
            first = 0;
            umma_commit_a(emptyB_a + sb * 8);
            if (++sb == BSLOTS) { sb = 0; phb ^= 1; }
          }
          umma_commit_a(emptyA_a + sa * 8);
          if (++sa == ASLOTS) { sa = 0; pha ^= 1; }
        }
        if (alive) umma_commit_a(tfull_a + buf * 8);
        ++seg;
      }
    }
  } else {
    // ================================================= epilogue ============================================
    tc_epilogue_role<MODE, BN, MT, NBUF>(P, out, bias, tmem, tfull_a, tempty_a, warp, lane, err,
                                         TcRanges(P, nullptr, 0, 0));   // P.dyn == 0 here: static ranges
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

// Window view of the 3-channel image layers (kH = kW = 4, stride 2).  x is re-laid once per call as
//   xk[frame][ho][wq][kh = 4][c4]   = x[frame][ho*sH + kh - pH][wq - pW][c]   (zero border, channels padded to 4)
// i.e. the kH input rows an output row reads are interleaved per pixel (2x the image, 38 MB for Dv.dc1).  The 4 px x 4 kh
// x 4 ch = 64 values an output pixel (ho, wo) reads are then 128 CONTIGUOUS bytes starting at pixel wo*sW, so the map
//   dims (64, Wo, Ho, frames)   strides (-, sW px = 64 B, row, frame)        [windows overlap along wo: loads only]
// with box (64, BW, BH, BF) lands [pixel][(kw, kh, c4) = 64] rows of 128 B in shared memory — exactly the K-major SW128
// A tile of a 64-channel layer.  The layer then runs through the implicit-GEMM kernels as a "1 x 1 x kT-tap convolution
// over 64 channels", and the M x 192 im2col matrix (179 MB for Dv.dc1, written and re-read on every call) is gone.
struct TcWin {
  const void* xk;
  int Wq, Ho, Wo, frames, T;    // T = frames per sample (merged coordinate = n*T + t)
  int sW;
};
static int win_map(CUtensorMap* m, const TcWin& w, int bw, int bh, int bf) {
  uint64_t dims[5] = {64, (uint64_t)w.Wo, (uint64_t)w.Ho, (uint64_t)w.frames, 1};
  const uint64_t row = (uint64_t)w.Wq * 32;
  uint64_t str[4] = {(uint64_t)w.sW * 32, row, (uint64_t)w.Ho * row, (uint64_t)w.frames * w.Ho * row};
  uint32_t box[5] = {64, (uint32_t)bw, (uint32_t)bh, (uint32_t)bf, 1};
  uint32_t es[5] = {1, 1, 1, 1, 1};
  return get_map(m, w.xk, 5, dims, str, box, es);
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

// ---- tile shape selection -------------------------------------------------------------------------------------
// A CTA tile is MT blocks of 128 accumulator rows x BN columns; one K step moves MT*16 KB of A and BN*128 B of B from L2
// into shared memory and issues MT*4 MMAs (2*MT*BN tensor cycles).  The kernel is persistent (one CTA per SM), so the
// cost of a shape is waves x (K steps x max(tensor cycles, bytes / L2 share)) — wide tiles raise FLOP/byte, small
// tiles fill the machine when a layer has few pixels.
struct TileCfg { int mt, bn; };
static int tc_env_int(const char* name) {
  const char* v = getenv(name);
  return v ? atoi(v) : 0;
}
// SMs the persistent kernels may occupy.  A data-parallel job leaves a few SMs to the collective's CTAs: a persistent
// grid of one CTA per SM launched while k SMs are held by an all-reduce runs its last k CTAs as a second wave (up to 2x
// the kernel time), a grid of sms - k CTAs does not (mcg_set_tc_sm_limit; 0 = all, MCG_TC_SMS overrides the default).
static std::atomic<int> g_tc_sm_limit{-1};
static int tc_sms() {
  int lim = g_tc_sm_limit.load(std::memory_order_relaxed);
  if (lim < 0) {
    lim = tc_env_int("MCG_TC_SMS");
    g_tc_sm_limit.store(lim, std::memory_order_relaxed);
  }
  const int n = num_sms();
  return (lim > 0 && lim < n) ? lim : n;
}
static double step_cycles(int mt, int bn, long long active_ctas) {
  const double chip_bw = 7400.0, sm_cap = 110.0;   // L2 -> SM bytes per cycle: whole chip (measured on these kernels), one SM
  const int sms = tc_sms();
  double share = chip_bw / (double)(active_ctas < sms ? (active_ctas > 0 ? active_ctas : 1) : sms);
  if (share > sm_cap) share = sm_cap;
  const double bytes = mt * 16384.0 + bn * 128.0, mma = 2.0 * mt * bn;
  const double ld = bytes / share;
  return ld > mma ? ld : mma;
}
static TileCfg pick_tile(int mode, long long nboxes, int ncls, int cols, int nk, double* cost_out = nullptr) {
  const int fm = tc_env_int("MCG_TC_MT"), fb = tc_env_int("MCG_TC_BN");
  const int sms = tc_sms();
  TileCfg best{1, 64};
  double best_cost = 1e300;
  const int bns[4] = {256, 192, 128, 64};
  for (int bi = 0; bi < 4; ++bi) {
    const int bn = bns[bi];
    if (cols % bn) continue;
    if (bn == 192 && mode != kFprop) continue;
    if (fb && bn != fb && cols % fb == 0) continue;
    for (int mt = 1; mt <= 4; ++mt) {
      if (mt == 3 || (mt == 4 && bn != 64)) continue;
      if (fm && mt != fm && !(fm == 4 && bn != 64 && mt == 2)) continue;
      const long long units = (long long)ncls * (cols / bn) * nboxes;
      const long long ctas = units < sms ? units : sms;
      const long long upc = (units + ctas - 1) / ctas;
      const long long full = upc / mt, rem = upc % mt;
      const double cost = (double)nk * ((double)full * step_cycles(mt, bn, ctas) + (rem ? step_cycles((int)rem, bn, ctas) : 0.0)) +
                          (double)(full + (rem ? 1 : 0)) * 300.0 + mt * bn * 6.0;
      if (cost < best_cost) { best_cost = cost; best = TileCfg{mt, bn}; }
    }
  }
  if (cost_out) *cost_out = best_cost;
  return best;
}
// cut `units` into equal contiguous ranges, one per CTA
static int split_units(TcParams& P, long long units, long long min_per_cta) {
  const int sms = tc_sms();
  long long ctas = units / (min_per_cta > 0 ? min_per_cta : 1);
  if (ctas > sms) ctas = sms;
  if (ctas < 1) ctas = 1;
  P.total_units = units;
  P.units_per_cta = (units + ctas - 1) / ctas;
  return (int)((units + P.units_per_cta - 1) / P.units_per_cta);
}

// Dynamic work distribution is worth its two atomics per range only when a CTA has several tile steps to draw from.
static std::atomic<int> g_tc_dyn{-1};
static std::atomic<unsigned> g_tc_sched_seq{0};
static void tc_set_schedule(TcParams& P, int mode, int grid, int mt) {
  int dyn = g_tc_dyn.load(std::memory_order_relaxed);
  if (dyn < 0) {
    const char* e = getenv("MCG_TC_DYN");
    dyn = e ? atoi(e) : MCG_TC_DYN_DEFAULT;
    g_tc_dyn.store(dyn, std::memory_order_relaxed);
  }
  P.dyn = 0;
  P.sched_slot = 0;
  // Strided groups make the skipped taps balance across CTAs, but CTAs then walk neighbouring boxes of one class in
  // lock-step; measured (B200, Dv layers): a win for the 64-column data gradients (weights are 11 % of a step's bytes), a
  // loss for the 128/256-column ones (33-50 %), which keep contiguous ranges and skip what their mixed-frame steps allow.
  const int fs = tc_env_int("MCG_TC_STRIDED");     // 1 / -1 force on / off
  const bool strided = fs > 0 || (fs == 0 && P.box_tn);
  P.strided = (mode != kWgrad && P.skip_t && strided) ? mt : 0;
  if (P.strided) {
    P.str_rounds = (int)((P.total_units / mt) / grid);
    P.str_rem0 = (long long)P.str_rounds * grid * mt;
    P.str_rem_per = (int)((P.total_units - P.str_rem0 + grid - 1) / grid);
    return;
  }
  if (dyn && mode != kWgrad && P.total_units < (1LL << 30) && P.total_units >= 4LL * grid * mt) {
    P.dyn = 1;
    P.sched_slot = (int)(g_tc_sched_seq.fetch_add(1, std::memory_order_relaxed) % kSchedSlots);
  }
}

template <int MODE, int BN, int MT>
static int launch_tc(const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& P, int grid, void* out, const float* bias,
                     cudaStream_t st, const char* who) {
  constexpr int STAGE = MT * A_BYTES + BN * 128;
  constexpr int BUDGET = 220 * 1024;
  constexpr int STAGES = (BUDGET / STAGE) > 8 ? 8 : (BUDGET / STAGE);
  static_assert(STAGES >= 2, "ring too shallow");
  size_t smem = (size_t)STAGES * STAGE + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_conv_kernel<MODE, BN, MT, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) MCG_FAIL((int)e, "%s: cudaFuncSetAttribute: %s", who, cudaGetErrorString(e));
    configured = true;
  }
  pdl(tc_conv_kernel<MODE, BN, MT, STAGES>, grid, kTcConvThreads, smem, st)(ma, mb, P, out, bias);
  MCG_CHECK_LAUNCH(who);
  return 0;
}
template <int MODE>
static int launch_tc_cfg(TileCfg c, const CUtensorMap& ma, const CUtensorMap& mb, TcParams& P, int grid, void* out,
                         const float* bias, cudaStream_t st, const char* who) {
  tc_set_schedule(P, MODE, grid, c.mt);
#define MCG_TC_CASE(bn_, mt_) \
  if (c.bn == bn_ && c.mt == mt_) return launch_tc<MODE, bn_, mt_>(ma, mb, P, grid, out, bias, st, who)
  MCG_TC_CASE(64, 1); MCG_TC_CASE(64, 2);
  if (MODE != kWgrad) { MCG_TC_CASE(64, 4); }
  MCG_TC_CASE(128, 1); MCG_TC_CASE(128, 2);
  MCG_TC_CASE(256, 1); MCG_TC_CASE(256, 2);
  if (MODE == kFprop) { MCG_TC_CASE(192, 1); MCG_TC_CASE(192, 2); }
#undef MCG_TC_CASE
  MCG_FAIL(MCG_ERR_UNSUPPORTED, "%s: tile %dx%d", who, c.mt * 128, c.bn);
}

// ---- temporal tap re-use: plan and launch ---------------------------------------------------------------------
struct TrPlan {
  bool ok;
  Box bx;                 // (BW, BH, BT, 1)
  int mt, bn, aslots, bslots;
  double cost;
};
// E* = extents the 128-pixel boxes tile, groups = (kh,kw) tap groups per class (max), KT temporal taps per group
static TrPlan plan_tr(int EW, int EH, int ET, int N, int cols, int ncls, int groups, int chunks, int KT) {
  TrPlan best;
  best.ok = false;
  best.cost = 1e300;
  const int sms = tc_sms();
  const double chip_bw = 7400.0, sm_cap = 110.0;
  const int fm = tc_env_int("MCG_TC_MT"), fb = tc_env_int("MCG_TC_BN"), fr = tc_env_int("MCG_TC_TR_R");
  for (int bw = 1; bw <= 64; bw *= 2)
    for (int bh = 1; bw * bh <= 64; bh *= 2) {
      const int R = bw * bh;
      if (R < 16 || (fr && R != fr)) continue;                 // R % 8 == 0 and BT = 128 / R <= 8
      const int bt = 128 / R;
      const long long nboxes = (long long)ceil_div(EW, bw) * ceil_div(EH, bh) * ceil_div(ET, bt) * N;
      const int ablock = (bt + KT - 1) * R * 128;
      const int bns[3] = {256, 128, 64};
      for (int bi = 0; bi < 3; ++bi) {
        const int bn = bns[bi];
        if (cols % bn || (fb && bn != fb && cols % fb == 0)) continue;
        for (int mt = 4; mt >= 1; mt /= 2) {
          if (mt * bn > 256 || (fm && mt != fm)) continue;     // keep the accumulators double-buffered
          const int bbytes = bn * 128;
          int aslots = 2;
          long long left = 220 * 1024 - (long long)aslots * mt * ablock;
          if (left < 3LL * bbytes) continue;
          int bslots = (int)(left / bbytes);
          if (bslots > kTrMaxBSlots) bslots = kTrMaxBSlots;
          left -= (long long)bslots * bbytes;
          while (aslots < kTrMaxASlots && left >= (long long)mt * ablock) { ++aslots; left -= (long long)mt * ablock; }
          const long long units = (long long)ncls * (cols / bn) * nboxes;
          const long long ctas = units < sms ? units : sms;
          const long long upc = (units + ctas - 1) / ctas;
          const long long steps = (upc + mt - 1) / mt;
          double share = chip_bw / (double)ctas;
          if (share > sm_cap) share = sm_cap;
          const double bytes = (double)mt * ablock + (double)KT * bbytes, mma = (double)KT * 2.0 * mt * bn;
          const double stepc = bytes / share > mma ? bytes / share : mma;
          const double cost = (double)steps * ((double)groups * chunks * stepc + 300.0) + mt * bn * 6.0;
          if (cost < best.cost) {
            best.ok = true; best.bx = Box{bw, bh, bt, 1}; best.mt = mt; best.bn = bn; best.aslots = aslots; best.bslots = bslots;
            best.cost = cost;
          }
        }
      }
    }
  return best;
}
template <int MODE, int BN, int MT>
static int launch_tr(const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& P, int grid, void* out, const float* bias,
                     cudaStream_t st, const char* who) {
  const size_t smem = (size_t)P.tr_aslots * MT * P.tr_ablock_bytes + (size_t)P.tr_bslots * BN * 128 + 1024;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_conv_tr_kernel<MODE, BN, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) MCG_FAIL((int)e, "%s: cudaFuncSetAttribute: %s", who, cudaGetErrorString(e));
    configured = smem;
  }
  pdl(tc_conv_tr_kernel<MODE, BN, MT>, grid, kTcThreads, smem, st)(ma, mb, P, out, bias);
  MCG_CHECK_LAUNCH(who);
  return 0;
}
template <int MODE>
static int launch_tr_cfg(int bn, int mt, const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& P, int grid, void* out,
                         const float* bias, cudaStream_t st, const char* who) {
#define MCG_TR_CASE(bn_, mt_) \
  if (bn == bn_ && mt == mt_) return launch_tr<MODE, bn_, mt_>(ma, mb, P, grid, out, bias, st, who)
  MCG_TR_CASE(64, 1); MCG_TR_CASE(64, 2); MCG_TR_CASE(64, 4);
  MCG_TR_CASE(128, 1); MCG_TR_CASE(128, 2);
  MCG_TR_CASE(256, 1);
#undef MCG_TR_CASE
  MCG_FAIL(MCG_ERR_UNSUPPORTED, "%s: temporal re-use tile %dx%d", who, mt * 128, bn);
}
// MCG_TC_TR=1 switches the temporal re-use kernel on (whenever the geometry allows).  It is OFF by default: measured on
// B200 it halves the L2 -> SM bytes of Dv.dc2 dgrad (2.18 -> 1.11 GB, ncu) yet runs 0.132 -> 0.17 ms, because the
// N = 64 / 128 layers are not L2-bound but shared-memory-operand-bound (a 128x64x16 MMA reads 6 KB of smem in its 32
// tensor cycles = 192 B/cycle against the SM's 128 B/cycle) and the temporal boxes pad T = 13 to 16 (+23 % MMAs).
static bool tr_wanted(const TrPlan& tr, double plain_cost) {
  (void)plain_cost;
  const char* v = getenv("MCG_TC_TR");
  return v && atoi(v) != 0 && tr.ok;
}

// =============================================================================================================
// Class-fused data gradient for the 64-channel stride-2 layers (Dv.dc2, G.dc4: k = 4, s = 2, p = 1 in H and W).
//
// The plain dgrad treats the sH*sW = 4 parity classes of the input grid as four independent stride-1 convolutions
// over dy with N = Cin = 64 accumulator columns each: every class loads its own 2 x 2 (kh, kw) boxes of dy per temporal
// tap, 16 box loads per (kt, 64-channel chunk) and 128 class pixels, and these kernels are bound by L2 -> SM bytes
// (the tensor pipe of Dv.dc2's dgrad is busy 45 % of the time while the SMs pull 2.2 GB through a ~7 KB/clk fabric).
// But class (ph, pw) reads dy at row offsets {-1, 0} (ph = 0) or {0, +1} (ph = 1), likewise in w: the 16 loads are only
// 9 distinct boxes.  Here ONE tile is 128 pixels of the class sub-grid x 256 columns = the four classes x 64 ci, laid
// out in TMEM in the cyclic order (0,0) (0,1) (1,1) (1,0) so that the classes sharing an offset are adjacent columns:
//   offset ( 0, 0)            -> one MMA, N = 256          offsets (-1,0) (0,+1) (+1,0) -> one MMA each, N = 128
//   offset ( 0,-1)            -> two MMAs, N = 64           the four corners            -> one MMA each, N = 64
// Two pixel blocks (MT = 2) share every weight slab, so the accumulators fill all 512 TMEM columns (single-buffered;
// the TMA ring keeps loading the next tile under the epilogue).  Per (kt, chunk) and 256 pixels the SM pulls
// 2 x 9 x 16 KB of dy + 16 x 8 KB of weights = 416 KB instead of 576 KB, and issues 40 wide MMAs instead of 64 narrow
// ones.  Temporal taps whose dy frame lies outside the clip are skipped like in the plain kernel.
// Work split: boxes are numbered (w, h, n, t) — frames slowest — and CTA c takes the block pairs c, c + grid, c + 2*grid, ...:
// the two blocks of a pair share their frame (so they skip the same temporal taps), every CTA samples all frames (so the
// skipped taps balance), and the pairs left over after the last full round lie in the last frame, the cheapest one.
// =============================================================================================================
struct Tc4Stage {
  int8_t dh, dw;            // dy box offset
  uint8_t nslab;            // weight slabs (64 co x 64 ci, one per class that uses this offset), in TMEM column order
  uint8_t nmma;             // MMAs over runs of adjacent classes
  uint8_t khkw[4];          // kh*kW + kw of each slab's tap
  uint8_t mma_col[2], mma_n[2], mma_slab[2];   // per MMA: first TMEM class slot, classes covered, first slab
  uint8_t big;              // uses the 4-slab ring slot
};
struct Tc4Params {
  int BW, BH, BT, BB, nbw, nbh, nbt, nbb;
  int EN, full_w, full_h, full_t;
  long long os_w, os_h, os_t, os_n;
  int chunks, kT, kHW, pT, aT, Cin;
  int nblocks, rounds, rem0, rem_per;
  int out_f32;
  int nst;                  // stages per (kt, chunk): the 9 offsets, widest first
  uint8_t cls_at_slot[4], slot_of_cls[4];
  Tc4Stage st[9];
};
// The stage table of the one geometry this kernel takes (k = 4, s = 2, p = 1), as compile-time constants: the producer
// and MMA loops are single threads whose issue rate bounds the kernel, and a table read from the parameter bank with a
// run-time index (LDC c[0x0][R + ...]) stalled the MMA thread on its long scoreboard for two thirds of its samples
// (profiles/r02_ncu_dv_dc2.txt).  The host still derives the table from the tap geometry and refuses to launch unless it
// equals this one.  Classes c = ph*2 + pw sit in TMEM slots (0,0) (0,1) (1,1) (1,0) = c 0, 1, 3, 2.
struct Tc4StageC { int dh, dw, nslab, nmma, khkw[4], col[2], n[2], slab[2]; };
__device__ constexpr Tc4StageC kTc4[9] = {
    {0, 0, 4, 1, {5, 6, 10, 9}, {0, 0}, {4, 0}, {0, 0}},      // centre: all four classes, one N = 256 MMA
    {-1, 0, 2, 1, {13, 14, 0, 0}, {0, 0}, {2, 0}, {0, 0}},    // ph = 0 row above: classes (0,0) (0,1)
    {0, -1, 2, 2, {7, 11, 0, 0}, {0, 3}, {1, 1}, {0, 1}},     // pw = 0 column left: classes (0,0) and (1,0), not adjacent
    {0, 1, 2, 1, {4, 8, 0, 0}, {1, 0}, {2, 0}, {0, 0}},       // pw = 1 column right: classes (0,1) (1,1)
    {1, 0, 2, 1, {2, 1, 0, 0}, {2, 0}, {2, 0}, {0, 0}},       // ph = 1 row below: classes (1,1) (1,0)
    {-1, -1, 1, 1, {15, 0, 0, 0}, {0, 0}, {1, 0}, {0, 0}},    // corners: one class each
    {-1, 1, 1, 1, {12, 0, 0, 0}, {1, 0}, {1, 0}, {0, 0}},
    {1, -1, 1, 1, {3, 0, 0, 0}, {3, 0}, {1, 0}, {0, 0}},
    {1, 1, 1, 1, {0, 0, 0, 0}, {2, 0}, {1, 0}, {0, 0}},
};
// Shared-memory ring in 8 KB granules.  A stage's pieces (a 16 KB dy tile per live block, one run of 1 / 2 / 4 adjacent
// weight slabs per MMA) are placed one after another; a piece that would cross the end of the ring starts over at 0.
// Stage sizes differ (40 - 64 KB), so fixed slots would leave a third of the ring idle — and what bounds these kernels
// is bytes in flight per SM (Little's law against the L2 round trip), not the number of stages.
constexpr int kT4Gran = 8192, kT4RingGr = 27, kT4Bars = 8;
constexpr int kT4Smem = kT4RingGr * kT4Gran + 1024;
__device__ __forceinline__ int tc4_take(int& head, int n, int& foot) {
  if (head + n > kT4RingGr) { foot += kT4RingGr - head; head = 0; }
  const int a = head;
  head += n;
  foot += n;
  return a;
}

struct Tc4Seg { int box0, nlive, kt_lo, kt_n; };
struct Tc4Iter {
  int k;
  __device__ __forceinline__ explicit Tc4Iter(const Tc4Params&) : k(0) {}
  __device__ __forceinline__ bool next(const Tc4Params& P, Tc4Seg& s) {
    if (k < P.rounds) {
      s.box0 = 2 * ((int)blockIdx.x + k * (int)gridDim.x);
      s.nlive = 2;
    } else if (k == P.rounds) {     // the blocks the full rounds left over: rem_per (1 or 2) per CTA
      s.box0 = P.rem0 + (int)blockIdx.x * P.rem_per;
      s.nlive = P.nblocks - s.box0 < P.rem_per ? P.nblocks - s.box0 : P.rem_per;
      if (s.nlive <= 0) return false;
    } else {
      return false;
    }
    ++k;
    // temporal taps worth issuing: dy frame = t + pT - kt must meet [0, aT) for some frame t of the step's boxes
    // (boxes are numbered (w, h, n, t): the two blocks of a pair share their frame except once per frame change)
    const int per_t = P.nbw * P.nbh * P.nbb;
    const int t_lo = (s.box0 / per_t) * P.BT, t_hi = ((s.box0 + s.nlive - 1) / per_t) * P.BT + P.BT - 1;
    int lo = t_lo + P.pT - (P.aT - 1), hi = t_hi + P.pT;
    if (lo < 0) lo = 0;
    if (hi > P.kT - 1) hi = P.kT - 1;
    s.kt_lo = lo; s.kt_n = hi >= lo ? hi - lo + 1 : 0;
    return true;
  }
};
__device__ __forceinline__ TcBox tc4_decode_box(const Tc4Params& P, int bi) {
  TcBox b;
  b.w0 = (bi % P.nbw) * P.BW; bi /= P.nbw;
  b.h0 = (bi % P.nbh) * P.BH; bi /= P.nbh;
  b.n0 = (bi % P.nbb) * P.BB; bi /= P.nbb;
  b.t0 = bi * P.BT;
  return b;
}

// 352 threads: warp 0 producer, warps 1 and 10 MMA issuers (one per pixel block of the pair), warps 2-9 epilogue.  With one
// issuer the kernel was bound by that thread's instruction stream even after the stage table became compile-time (ncu
// source page: 399 of its 444 samples issuing, 45 waiting for data, ~115 cycles per MMA against ~50 of tensor time): the two
// blocks of a pair accumulate into disjoint TMEM columns from the same weight slabs, so each gets its own issuing thread.
constexpr int kT4Threads = 352;
__global__ void __launch_bounds__(kT4Threads, 1) tc_dgrad4_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                  const __grid_constant__ CUtensorMap mapB,
                                                                  const __grid_constant__ Tc4Params P, void* __restrict__ out,
                                                                  const float* __restrict__ bias) {
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[kT4Bars], empty_bar[kT4Bars], tfull_bar, tempty_bar;
  __shared__ int ring_fp[kT4Bars];      // granules (waste included) each in-flight stage holds; producer thread only
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* err = &g_tc_error;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kT4Bars; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 2); }   // both issuers release a stage
    mbar_init(&tfull_bar, 2);
    mbar_init(&tempty_bar, kEpiThreads);
    fence_barrier_init();
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 1) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  pdl_wait();

  const uint32_t smem_a = smem_u32(smem);
  const uint32_t full_a = smem_u32(&full_bar[0]), empty_a = smem_u32(&empty_bar[0]);
  const uint32_t tfull_a = smem_u32(&tfull_bar), tempty_a = smem_u32(&tempty_bar);

  if (warp == 0) {
    if (elect_one()) {
      const uint64_t map_a = reinterpret_cast<uint64_t>(&mapA), map_b = reinterpret_cast<uint64_t>(&mapB);
      Tc4Iter iter(P);
      Tc4Seg sg;
      int head = 0, freeg = kT4RingGr, istage = 0, tail = 0;
      bool alive = true;
      while (alive && iter.next(P, sg)) {
        TcBox bx[2];
        bx[0] = tc4_decode_box(P, sg.box0);
        bx[1] = tc4_decode_box(P, sg.box0 + (sg.nlive > 1 ? 1 : 0));
        for (int kt = sg.kt_lo; alive && kt < sg.kt_lo + sg.kt_n; ++kt) {
          const int dt = P.pT - kt, kbase = kt * P.kHW;
          for (int c = 0; alive && c < P.chunks; ++c) {
#pragma unroll
            for (int s = 0; s < 9; ++s) {
              constexpr const Tc4StageC* T = kTc4;
              int h = head, foot = 0;
              const int a0 = tc4_take(h, 2, foot), a1 = sg.nlive > 1 ? tc4_take(h, 2, foot) : 0;
              const int b0 = tc4_take(h, T[s].n[0], foot), b1 = T[s].nmma > 1 ? tc4_take(h, T[s].n[1], foot) : 0;
              // room: stages retire in order; wait for as many of the oldest as this one's footprint (or a barrier pair) needs
              while (freeg < foot || istage - tail >= kT4Bars) {
                if (!mbar_wait_a(empty_a + (tail % kT4Bars) * 8, (tail / kT4Bars) & 1, err)) { alive = false; break; }
                freeg += ring_fp[tail % kT4Bars];
                ++tail;
              }
              if (!alive) break;
              head = h; freeg -= foot; ring_fp[istage % kT4Bars] = foot;
              const uint32_t fb = full_a + (istage % kT4Bars) * 8;
              ++istage;
              mbar_expect_tx_a(fb, sg.nlive * A_BYTES + T[s].nslab * 8192);
              tma_load_5d_a(smem_a + a0 * kT4Gran, map_a, fb, c * 64, bx[0].w0 + T[s].dw, bx[0].h0 + T[s].dh, bx[0].t0 + dt, bx[0].n0);
              if (sg.nlive > 1)
                tma_load_5d_a(smem_a + a1 * kT4Gran, map_a, fb, c * 64, bx[1].w0 + T[s].dw, bx[1].h0 + T[s].dh, bx[1].t0 + dt, bx[1].n0);
#pragma unroll
              for (int sl = 0; sl < T[s].n[0]; ++sl)
                tma_load_2d_a(smem_a + (b0 + sl) * kT4Gran, map_b, fb, (kbase + T[s].khkw[sl]) * 64, c * 64);
              if (T[s].nmma > 1) {
#pragma unroll
                for (int sl = 0; sl < T[s].n[1]; ++sl)
                  tma_load_2d_a(smem_a + (b1 + sl) * kT4Gran, map_b, fb, (kbase + T[s].khkw[T[s].slab[1] + sl]) * 64, c * 64);
              }
            }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 10) {
    if (elect_one()) {
      // issuer `mine` owns block `mine` of every pair.  Both walk every stage (and wait for its data, which keeps them
      // within one ring of each other); a stage is released by two arrivals — a tcgen05.commit from an issuer that
      // multiplied from it, a plain arrive from one that had no block in it (single-block steps).
      const int mine = warp == 1 ? 0 : 1;
      const uint32_t desc_hi = smem_desc_hi(1024);
      const uint32_t id64 = make_idesc_bf16(128, 64, 0, 1), id128 = make_idesc_bf16(128, 128, 0, 1), id256 = make_idesc_bf16(128, 256, 0, 1);
      Tc4Iter iter(P);
      Tc4Seg sg;
      uint32_t nseg = 0;
      int head = 0, istage = 0;
      bool alive = true;
      while (alive && iter.next(P, sg)) {
        if (sg.kt_n == 0) continue;
        if (!mbar_wait_a(tempty_a, (nseg & 1) ^ 1, err)) break;
        tc_fence_after();
        const int steps = sg.kt_n * P.chunks;
        const bool active = mine < sg.nlive;
        bool first = true;
        for (int g = 0; alive && g < steps; ++g) {
#pragma unroll
          for (int s = 0; s < 9; ++s) {
            constexpr const Tc4StageC* T = kTc4;
            int foot = 0;        // the same placement the producer computed
            const int a0 = tc4_take(head, 2, foot), a1 = sg.nlive > 1 ? tc4_take(head, 2, foot) : 0;
            const int b0 = tc4_take(head, T[s].n[0], foot), b1 = T[s].nmma > 1 ? tc4_take(head, T[s].n[1], foot) : 0;
            const int bar = istage % kT4Bars;
            if (!mbar_wait_a(full_a + bar * 8, (istage / kT4Bars) & 1, err)) { alive = false; break; }
            ++istage;
            if (active) {
              tc_fence_after();
              const uint32_t a_lo = smem_desc_lo(smem_a + (mine ? a1 : a0) * kT4Gran, 16);
#pragma unroll
              for (int j = 0; j < T[s].nmma; ++j) {
                const uint32_t d = tmem + mine * 256 + T[s].col[j] * 64;
                const uint32_t id = T[s].n[j] == 4 ? id256 : (T[s].n[j] == 2 ? id128 : id64);
                const uint32_t bj = smem_desc_lo(smem_a + (j ? b1 : b0) * kT4Gran, 8192);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16_lh(d, a_lo + k * 2, desc_hi, bj + k * (2048 >> 4), desc_hi, id, (first && k == 0) ? 0u : 1u);
              }
              umma_commit_a(empty_a + bar * 8);
              first = false;
            } else {
              mbar_arrive_a(empty_a + bar * 8);
            }
          }
        }
        if (alive) {
          if (active) umma_commit_a(tfull_a);
          else mbar_arrive_a(tfull_a);
        }
        ++nseg;
      }
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    int rr = r;
    const int iw = rr % P.BW; rr /= P.BW;
    const int ih = rr % P.BH; rr /= P.BH;
    const int itt = rr % P.BT; rr /= P.BT;
    const int ib = rr;
    const int c0 = half * 32;
    float bv[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) bv[i] = bias ? bias[c0 + i] : 0.f;
    Tc4Iter iter(P);
    Tc4Seg sg;
    uint32_t nseg = 0;
    while (iter.next(P, sg)) {
      const bool have_acc = sg.kt_n > 0;
      if (have_acc) {
        if (!mbar_wait_a(tfull_a, nseg & 1, err)) break;
        tc_fence_after();
      }
      const uint32_t acc = tmem + (uint32_t(q * 32) << 16);
      const int s_lo = 0, s_hi = 4;
#pragma unroll 1
      for (int m = 0; m < sg.nlive; ++m) {
        const TcBox bx = tc4_decode_box(P, sg.box0 + m);
        const int on = bx.n0 + ib, ot = bx.t0 + itt;
#pragma unroll 1
        for (int sl = s_lo; sl < s_hi; ++sl) {
          const int cls = sl ^ (sl >> 1);      // TMEM slot -> class: 0, 1, 3, 2 (kTc4's column order)
          const int ow = (bx.w0 + iw) * 2 + (cls & 1), oh = (bx.h0 + ih) * 2 + (cls >> 1);
          const bool valid = ow < P.full_w && oh < P.full_h && ot < P.full_t && on < P.EN;
          uint32_t v[32];
          if (have_acc) {
            tmem_ld32(acc + m * 256 + sl * 64 + c0, v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0u;
          }
          if (valid) {
            const long long base = (long long)on * P.os_n + (long long)ot * P.os_t + (long long)oh * P.os_h + (long long)ow * P.os_w + c0;
            if (P.out_f32) {
              float* o = reinterpret_cast<float*>(out) + base;
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<float4*>(o + i) = make_float4(__uint_as_float(v[i]) + bv[i], __uint_as_float(v[i + 1]) + bv[i + 1],
                                                                __uint_as_float(v[i + 2]) + bv[i + 2], __uint_as_float(v[i + 3]) + bv[i + 3]);
            } else {
              __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + base;
#pragma unroll
              for (int i = 0; i < 32; i += 8) {
                uint4 u;
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  h[e] = __floats2bfloat162_rn(__uint_as_float(v[i + 2 * e]) + bv[i + 2 * e], __uint_as_float(v[i + 2 * e + 1]) + bv[i + 2 * e + 1]);
                *reinterpret_cast<uint4*>(o + i) = u;
              }
            }
          }
        }
      }
      if (have_acc) {
        tc_fence_before();
        mbar_arrive_a(tempty_a);
        ++nseg;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// the fused kernel takes: Cin = 64, 4 x 4 spatial taps, stride 2, pad 1 in H and W, unit temporal stride, full weight rows,
// and enough pixel blocks that every CTA walks at least one pair
static bool dgrad4_geom_ok(const mcg_conv_geom* g, int wrows) {
  return g->Cin == 64 && g->Cout % 64 == 0 && g->kH == 4 && g->kW == 4 && g->sH == 2 && g->sW == 2 && g->pH == 1 && g->pW == 1 &&
         g->sT == 1 && g->kT <= 4 && wrows == g->Cout && g->Hi % 2 == 0 && g->Wi % 2 == 0;
}
static int tc_dgrad4(const mcg_conv_geom* g, const void* dy, const void* w, void* dx, const float* bias, int out_dtype,
                     cudaStream_t st, bool* taken) {
  const char* who = "mcg_conv_dgrad(tc,class-fused)";
  *taken = false;
  const int EW = g->Wi / 2, EH = g->Hi / 2, ET = g->Ti;
  const Box bx = choose_box(128, EW, EH, ET, g->N);
  Tc4Params P;
  memset(&P, 0, sizeof(P));
  P.BW = bx.w; P.BH = bx.h; P.BT = bx.t; P.BB = bx.b;
  P.nbw = ceil_div(EW, bx.w); P.nbh = ceil_div(EH, bx.h); P.nbt = ceil_div(ET, bx.t); P.nbb = ceil_div(g->N, bx.b);
  P.nblocks = P.nbw * P.nbh * P.nbt * P.nbb;
  const int sms = tc_sms();
  if (P.nblocks < 2 * sms) return 0;          // small layers (Di.dc2: 70 blocks) keep the per-class kernel
  *taken = true;
  P.EN = g->N; P.full_w = g->Wi; P.full_h = g->Hi; P.full_t = g->Ti;
  P.os_w = g->Cin; P.os_h = (long long)g->Wi * g->Cin; P.os_t = (long long)g->Hi * P.os_h; P.os_n = (long long)g->Ti * P.os_t;
  P.chunks = g->Cout / 64; P.kT = g->kT; P.kHW = g->kH * g->kW; P.pT = g->pT; P.aT = g->To; P.Cin = g->Cin;
  P.out_f32 = (out_dtype == MCG_F32);
  const int grid = sms;
  P.rounds = (P.nblocks / 2) / grid;
  P.rem0 = P.rounds * grid * 2;
  P.rem_per = (P.nblocks - P.rem0 + grid - 1) / grid;     // 0, 1 or 2
  // classes c = ph*2 + pw in the cyclic TMEM order (0,0) (0,1) (1,1) (1,0)
  const uint8_t order[4] = {0, 1, 3, 2};
  for (int s = 0; s < 4; ++s) { P.cls_at_slot[s] = order[s]; P.slot_of_cls[order[s]] = (uint8_t)s; }
  // which classes read dy at offset (dh, dw), and through which tap: dh = (ph + pH - kh) / sH where that division is exact
  struct Use { int cls, khkw; };
  std::vector<Use> uses[3][3];
  for (int c = 0; c < 4; ++c) {
    const int ph = c >> 1, pw = c & 1;
    for (int kh = 0; kh < 4; ++kh) {
      if ((ph + 1 - kh) % 2) continue;
      for (int kw = 0; kw < 4; ++kw) {
        if ((pw + 1 - kw) % 2) continue;
        const int dh = (ph + 1 - kh) / 2, dw = (pw + 1 - kw) / 2;
        if (dh < -1 || dh > 1 || dw < -1 || dw > 1) MCG_FAIL(MCG_ERR_UNSUPPORTED, "%s: offset out of the 3x3 window", who);
        uses[dh + 1][dw + 1].push_back(Use{c, kh * 4 + kw});
      }
    }
  }
  auto build = [&]() -> int {
    const int only_cls = -1;
    int n = 0;
    for (int pass = 4; pass >= 1; --pass)            // widest stages first: the first stage of a tile must cover every column
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
          std::vector<Use> u;
          for (const Use& x : uses[a][b])
            if (only_cls < 0 || x.cls == only_cls) u.push_back(x);
          if ((int)u.size() != pass) continue;
          if (only_cls >= 0 && pass != 1) continue;
          Tc4Stage& s = P.st[n++];
          memset(&s, 0, sizeof(s));
          s.dh = (int8_t)(a - 1); s.dw = (int8_t)(b - 1);
          // slabs in TMEM slot order
          for (size_t i = 0; i < u.size(); ++i)
            for (size_t j = i + 1; j < u.size(); ++j)
              if (P.slot_of_cls[u[j].cls] < P.slot_of_cls[u[i].cls]) std::swap(u[i], u[j]);
          s.nslab = (uint8_t)u.size();
          s.big = s.nslab > 2;
          for (size_t i = 0; i < u.size(); ++i) s.khkw[i] = (uint8_t)u[i].khkw;
          // MMAs over runs of adjacent slots
          size_t i = 0;
          while (i < u.size()) {
            size_t j = i + 1;
            while (j < u.size() && P.slot_of_cls[u[j].cls] == P.slot_of_cls[u[j - 1].cls] + 1) ++j;
            if (s.nmma >= 2 || (j - i) == 3) return -1;
            s.mma_col[s.nmma] = P.slot_of_cls[u[i].cls]; s.mma_n[s.nmma] = (uint8_t)(j - i); s.mma_slab[s.nmma] = (uint8_t)i;
            ++s.nmma;
            i = j;
          }
        }
    return n;
  };
  P.nst = build();
  if (P.nst != 9) MCG_FAIL(MCG_ERR_UNSUPPORTED, "%s: unexpected stage table (%d stages)", who, P.nst);
  {
    static const Tc4StageC kHost[9] = {
        {0, 0, 4, 1, {5, 6, 10, 9}, {0, 0}, {4, 0}, {0, 0}},   {-1, 0, 2, 1, {13, 14, 0, 0}, {0, 0}, {2, 0}, {0, 0}},
        {0, -1, 2, 2, {7, 11, 0, 0}, {0, 3}, {1, 1}, {0, 1}},  {0, 1, 2, 1, {4, 8, 0, 0}, {1, 0}, {2, 0}, {0, 0}},
        {1, 0, 2, 1, {2, 1, 0, 0}, {2, 0}, {2, 0}, {0, 0}},    {-1, -1, 1, 1, {15, 0, 0, 0}, {0, 0}, {1, 0}, {0, 0}},
        {-1, 1, 1, 1, {12, 0, 0, 0}, {1, 0}, {1, 0}, {0, 0}},  {1, -1, 1, 1, {3, 0, 0, 0}, {3, 0}, {1, 0}, {0, 0}},
        {1, 1, 1, 1, {0, 0, 0, 0}, {2, 0}, {1, 0}, {0, 0}}};
    for (int i = 0; i < 9; ++i) {      // the kernel's compile-time table must be what the tap geometry gives
      const Tc4Stage& a = P.st[i];
      const Tc4StageC& b = kHost[i];
      bool same = a.dh == b.dh && a.dw == b.dw && a.nslab == b.nslab && a.nmma == b.nmma;
      for (int k = 0; same && k < a.nslab; ++k) same = a.khkw[k] == b.khkw[k];
      for (int k = 0; same && k < a.nmma; ++k) same = a.mma_col[k] == b.col[k] && a.mma_n[k] == b.n[k] && a.mma_slab[k] == b.slab[k];
      if (!same) MCG_FAIL(MCG_ERR_UNSUPPORTED, "%s: stage %d of the derived table differs from the kernel's", who, i);
    }
  }
  CUtensorMap ma, mb;
  int rc;
  if ((rc = act_map(&ma, dy, g->Cout, g->Wo, g->Ho, g->To, g->N, bx.w, bx.h, bx.t, bx.b, 1, 1, 1))) return rc;
  const int Ktot = g->kT * g->kH * g->kW * g->Cin;
  uint64_t d2[2] = {(uint64_t)Ktot, (uint64_t)g->Cout}, s2[1] = {(uint64_t)Ktot * 2};
  uint32_t b2[2] = {64, 64}, e2[2] = {1, 1};
  if ((rc = get_map(&mb, w, 2, d2, s2, b2, e2))) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_dgrad4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kT4Smem);
    if (e != cudaSuccess) MCG_FAIL((int)e, "%s: cudaFuncSetAttribute: %s", who, cudaGetErrorString(e));
    configured = true;
  }
  pdl(tc_dgrad4_kernel, grid, kT4Threads, kT4Smem, st)(ma, mb, P, dx, bias);
  MCG_CHECK_LAUNCH(who);
  return 0;
}

// ---- split-K for fprop ------------------------------------------------------------------------------------------
// y[m][c] = bias[c] + scratch[m][c], cast to the output type (8 elements per thread; Cout % 8 == 0)
__global__ void __launch_bounds__(256) splitk_finish_kernel(const float4* __restrict__ acc, const float* __restrict__ bias,
                                                            void* __restrict__ y, int out_f32, long long n8, int C) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = acc[2 * i], b = acc[2 * i + 1];
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (bias) {
      const int c0 = (int)((i * 8) % C);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += bias[c0 + k];
    }
    if (out_f32) {
      float4* o = reinterpret_cast<float4*>(y) + 2 * i;
      o[0] = make_float4(v[0], v[1], v[2], v[3]);
      o[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      uint4 u;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
      for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
      reinterpret_cast<uint4*>(y)[i] = u;
    }
  }
}
// A layer whose widest tiles give fewer than half as many work units as there are SMs (Dv.dc4: 2,240 pixels = 18 boxes x 2
// column tiles of 256) either runs narrow tiles — every pixel box re-read once per column tile, every weight tile once
// per box: 0.88 GB through L2 for 37.6 GF — or leaves most SMs idle.  Split-K keeps the 128 x 256 tile and cuts the
// tap loop over S CTAs.  Returns S (1 = no split) for an fprop geometry; the scratch tensor is M x Cout fp32.
static int splitk_plan(const mcg_conv_geom* g, int wrows) {
  // OFF unless MCG_TC_SPLITK=1: measured on Dv.dc4 fprop (S = 4, 144 CTAs of 128 x 256 tiles), 0.061 -> 0.078 ms — the 4.6 M
  // scalar fp32 red.adds of the partial tiles plus the memset and finish launches cost more than the idle SMs did
  if (!tc_env_int("MCG_TC_SPLITK") || g->Cout % 256 || g->Cin % 64 || (wrows && wrows != g->Cout) || g->pT > 0) return 1;
  const int taps = g->kT * g->kH * g->kW;
  const Box bx = choose_box(128, g->Wo, g->Ho, g->To, g->N);
  const long long nboxes = (long long)ceil_div(g->Wo, bx.w) * ceil_div(g->Ho, bx.h) * ceil_div(g->To, bx.t) * ceil_div(g->N, bx.b);
  const long long units = nboxes * (g->Cout / 256);
  const int sms = tc_sms();
  if (units * 2 > sms) return 1;
  // the extra memset + finish launches only pay for a layer with real work (Dv.dc4: 37.6 GF; not Di.dc3/dc4: 2.3 GF)
  if (2.0 * g->N * g->To * g->Ho * g->Wo * (double)g->Cout * g->Cin * taps < 1e10) return 1;
  int S = 1;
  while (S < 8 && units * (S * 2) <= sms && taps % (S * 2) == 0 && (taps / (S * 2)) * (g->Cin / 64) >= 16) S *= 2;
  return S;
}
size_t tc_splitk_workspace(const mcg_conv_geom* g) {
  if (splitk_plan(g, 0) <= 1) return 0;
  return (size_t)g->N * g->To * g->Ho * g->Wo * g->Cout * sizeof(float) + 256;
}

bool tc_supported(const mcg_conv_geom* g) {
  if (g->Cin % 64 || g->Cout % 64) return false;
  if (g->sT > 2 || g->sH > 2 || g->sW > 2) return false;
  if (g->kT * g->kH * g->kW > 64) return false;
  if (g->sT * g->sH * g->sW > 8) return false;
  return true;
}

int tc_conv(int mode, const mcg_conv_geom* g, const void* a, const void* b, void* out, const float* bias, int out_dtype,
            cudaStream_t st, int kreal = 0, int planar_chunk = 0, int planar_cols = 0, int wrows = 0, const TcWin* win = nullptr,
            const int* d2s = nullptr, int store_cols = 0, float* splitk_scratch = nullptr) {   // d2s: {C, sT, sH, sW, Ti, Hi, Wi} -> depth-to-space epilogue (fprop)
  const char* who = mode == kFprop ? "mcg_conv_fprop(tc)" : mode == kDgrad ? "mcg_conv_dgrad(tc)" : "mcg_conv_wgrad(tc)";
  if (!tc_supported(g)) MCG_FAIL(MCG_ERR_UNSUPPORTED, "%s: needs Cin,Cout %% 64 == 0, stride <= 2, <= 64 taps", who);
  TcParams P;
  memset(&P, 0, sizeof(P));
  const int taps = g->kT * g->kH * g->kW;
  P.Cin = g->Cin; P.Cout = g->Cout; P.Ktot = taps * g->Cin;
  P.out_f32 = (out_dtype == MCG_F32);
  P.planar_chunk = planar_chunk;
  P.planar_cols = planar_cols;
  if (d2s) {
    if (mode != kFprop || d2s[0] * d2s[1] * d2s[2] * d2s[3] > 32) MCG_FAIL(MCG_ERR_UNSUPPORTED, "%s: depth-to-space epilogue", who);
    P.d2s_C = d2s[0]; P.d2s_sT = d2s[1]; P.d2s_sH = d2s[2]; P.d2s_sW = d2s[3]; P.d2s_Ti = d2s[4]; P.d2s_Hi = d2s[5]; P.d2s_Wi = d2s[6];
    P.d2s_cols = d2s[0] * d2s[1] * d2s[2] * d2s[3];
    for (int e = 0, c = 0; e < d2s[1]; ++e)
      for (int a2 = 0; a2 < d2s[2]; ++a2)
        for (int b2 = 0; b2 < d2s[3]; ++b2)
          for (int ci = 0; ci < d2s[0]; ++ci, ++c) P.d2s_tab[c] = (uint32_t)(e | (a2 << 4) | (b2 << 8) | (ci << 12));
  }
  P.planar_stride = (long long)g->Wo * planar_chunk;   // planar mode is only used on 1-D line geometries (M = Wo)
  // wrows < Cout: the weight tensor has fewer rows than the (zero-padded) channel count of the activations it meets —
  // rows beyond are TMA out-of-bounds zero fill, columns of dw beyond are not written
  if (wrows <= 0 || wrows > g->Cout) wrows = g->Cout;
  if (wrows < g->Cout && mode == kFprop && bias) MCG_FAIL(MCG_ERR_UNSUPPORTED, "%s: bias with padded weight rows", who);
  P.wrows = wrows;
  if (planar_chunk) {
    if (planar_chunk % 4 || planar_cols > 256) MCG_FAIL(MCG_ERR_UNSUPPORTED, "%s: planar chunk %d / cols %d", who, planar_chunk, planar_cols);
    for (int c = 0; c < planar_cols; c += 4) {
      P.planar_plane[c >> 2] = (uint8_t)(c / planar_chunk);
      P.planar_within[c >> 2] = (uint8_t)(c % planar_chunk);
    }
  }
  CUtensorMap ma, mb;
  int rc;
  if (mode == kFprop) {
    // a = x (N,Ti,Hi,Wi,Cin), b = w bf16 (Cout, taps*Cin), out = y (N,To,Ho,Wo,Cout)
    // (window view: frames of different samples are not a box of the merged frame axis, so 3-D layers take one sample per box)
    Box bx = (win && win->T > 1) ? choose_box(128, g->Wo, g->Ho, g->To, 1) : choose_box(128, g->Wo, g->Ho, g->To, g->N);
    if (win) { P.win = 1; P.win_T = win->T; }
    P.BW = bx.w; P.BH = bx.h; P.BT = bx.t; P.BB = bx.b;
    P.nbw = ceil_div(g->Wo, bx.w); P.nbh = ceil_div(g->Ho, bx.h); P.nbt = ceil_div(g->To, bx.t); P.nbb = ceil_div(g->N, bx.b);
    P.EW = g->Wo; P.EH = g->Ho; P.ET = g->To; P.EN = g->N;
    P.full_w = g->Wo; P.full_h = g->Ho; P.full_t = g->To;
    P.a_mul_w = g->sW; P.a_mul_h = g->sH; P.a_mul_t = g->sT; P.a_add_w = -g->pW; P.a_add_h = -g->pH; P.a_add_t = -g->pT;
    P.o_mul_w = P.o_mul_h = P.o_mul_t = 1;
    P.cls_w = P.cls_h = P.cls_t = 1;
    P.os_w = g->Cout; P.os_h = (long long)g->Wo * g->Cout; P.os_t = (long long)g->Ho * P.os_h; P.os_n = (long long)g->To * P.os_t;
    P.ocols = g->Cout;
    if (store_cols > 0) {       // narrow row-major output: rows of store_cols (< Cout) columns
      if (store_cols % 8 || store_cols > g->Cout || out_dtype != MCG_BF16) MCG_FAIL(MCG_ERR_UNSUPPORTED, "%s: store_cols %d", who, store_cols);
      P.store_cols = store_cols;
      P.os_w = store_cols; P.os_h = (long long)g->Wo * store_cols; P.os_t = (long long)g->Ho * P.os_h; P.os_n = (long long)g->To * P.os_t;
    }
    P.chunks = g->Cin / 64;
    P.tap_begin[0] = 0; P.tap_count[0] = taps;
    for (int kt = 0, j = 0; kt < g->kT; ++kt)
      for (int kh = 0; kh < g->kH; ++kh)
        for (int kw = 0; kw < g->kW; ++kw, ++j) P.taps[j] = TcTap{(int16_t)kw, (int16_t)kh, (int16_t)kt, (int16_t)j};
    P.nboxes = P.nbw * P.nbh * P.nbt * P.nbb;
    double plain_cost = 0;
    TileCfg cfg = pick_tile(kFprop, P.nboxes, 1, g->Cout, taps * P.chunks, &plain_cost);
    const int S = (splitk_scratch && !planar_chunk && !win && !d2s && !store_cols) ? splitk_plan(g, wrows == g->Cout ? 0 : wrows) : 1;
    if (S > 1) {
      // widest tile, tap loop cut S ways; partial tiles meet in the fp32 scratch (zeroed here), then bias + cast
      cfg = TileCfg{1, 256};
      P.ksplit = S; P.ks_taps = taps / S;
      if ((rc = act_map(&ma, a, g->Cin, g->Wi, g->Hi, g->Ti, g->N, bx.w, bx.h, bx.t, bx.b, g->sW, g->sH, g->sT))) return rc;
      P.ntn = g->Cout / cfg.bn;
      P.ntiles = P.ntn;
      const int grid = split_units(P, (long long)S * P.ntiles * P.nboxes, 1);
      uint64_t d2[2] = {(uint64_t)P.Ktot, (uint64_t)wrows}, s2[1] = {(uint64_t)P.Ktot * 2};
      uint32_t b2[2] = {64, (uint32_t)cfg.bn}, e2[2] = {1, 1};
      if ((rc = get_map(&mb, b, 2, d2, s2, b2, e2))) return rc;
      const long long n = (long long)g->N * g->To * g->Ho * g->Wo * g->Cout;
      cudaError_t e = cudaMemsetAsync(splitk_scratch, 0, (size_t)n * sizeof(float), st);
      if (e != cudaSuccess) MCG_FAIL((int)e, "%s: cudaMemsetAsync: %s", who, cudaGetErrorString(e));
      if ((rc = launch_tc_cfg<kFprop>(cfg, ma, mb, P, grid, splitk_scratch, nullptr, st, who))) return rc;
      long long nb = (n / 8 + 255) / 256;
      if (nb > (long long)num_sms() * 8) nb = (long long)num_sms() * 8;
      pdl(splitk_finish_kernel, (unsigned)nb, 256, 0, st)(reinterpret_cast<const float4*>(splitk_scratch), bias, out, out_dtype == MCG_F32,
                                                         n / 8, g->Cout);
      MCG_CHECK_LAUNCH(who);
      return 0;
    }
    if (g->kT >= 2 && g->sT == 1 && !planar_chunk && !win) {
      const TrPlan tr = plan_tr(g->Wo, g->Ho, g->To, g->N, g->Cout, 1, g->kH * g->kW, P.chunks, g->kT);
      if (tr_wanted(tr, plain_cost)) {
        const Box tb = tr.bx;
        P.BW = tb.w; P.BH = tb.h; P.BT = tb.t; P.BB = 1;
        P.nbw = ceil_div(g->Wo, tb.w); P.nbh = ceil_div(g->Ho, tb.h); P.nbt = ceil_div(g->To, tb.t); P.nbb = g->N;
        P.nboxes = P.nbw * P.nbh * P.nbt * P.nbb;
        P.tr_kT = g->kT; P.tr_kstep = g->kH * g->kW; P.tr_rev = 0;
        P.tr_frame_bytes = tb.w * tb.h * 128; P.tr_ablock_bytes = (tb.t + g->kT - 1) * P.tr_frame_bytes;
        P.tr_aslots = tr.aslots; P.tr_bslots = tr.bslots;
        P.tap_count[0] = g->kH * g->kW;
        for (int kh = 0, j = 0; kh < g->kH; ++kh)
          for (int kw = 0; kw < g->kW; ++kw, ++j) P.taps[j] = TcTap{(int16_t)kw, (int16_t)kh, 0, (int16_t)j};
        if ((rc = act_map(&ma, a, g->Cin, g->Wi, g->Hi, g->Ti, g->N, tb.w, tb.h, tb.t + g->kT - 1, 1, g->sW, g->sH, 1))) return rc;
        P.ntn = g->Cout / tr.bn;
        P.ntiles = P.ntn;
        const int grid = split_units(P, (long long)P.ntiles * P.nboxes, 1);
        uint64_t d2[2] = {(uint64_t)P.Ktot, (uint64_t)wrows}, s2[1] = {(uint64_t)P.Ktot * 2};
        uint32_t b2[2] = {64, (uint32_t)tr.bn}, e2[2] = {1, 1};
        if ((rc = get_map(&mb, b, 2, d2, s2, b2, e2))) return rc;
        return launch_tr_cfg<kFprop>(tr.bn, tr.mt, ma, mb, P, grid, out, bias, st, who);
      }
    }
    if (win) {
      if ((rc = win_map(&ma, *win, bx.w, bx.h, win->T > 1 ? bx.t : bx.b))) return rc;
    } else if ((rc = act_map(&ma, a, g->Cin, g->Wi, g->Hi, g->Ti, g->N, bx.w, bx.h, bx.t, bx.b, g->sW, g->sH, g->sT))) {
      return rc;
    }
    if (g->kT > 1 && g->pT > 0 && !tc_env_int("MCG_TC_NOSKIP")) {   // temporal zero padding: edge boxes skip the taps that read only padding
      P.skip_t = 1; P.skip_per_kt[0] = g->kH * g->kW; P.skip_nkt = g->kT; P.skip_dt0 = 0; P.skip_dts = 1; P.skip_aT = g->Ti;
      P.box_tn = tc_env_int("MCG_TC_STRIDED") > 0 || (tc_env_int("MCG_TC_STRIDED") == 0 && cfg.bn == 64);
    }
    P.ntn = g->Cout / cfg.bn;
    P.ntiles = P.ntn;
    const int grid = split_units(P, (long long)P.ntiles * P.nboxes, 1);
    uint64_t d2[2] = {(uint64_t)P.Ktot, (uint64_t)wrows}, s2[1] = {(uint64_t)P.Ktot * 2};
    uint32_t b2[2] = {64, (uint32_t)cfg.bn}, e2[2] = {1, 1};
    if ((rc = get_map(&mb, b, 2, d2, s2, b2, e2))) return rc;
    return launch_tc_cfg<kFprop>(cfg, ma, mb, P, grid, out, bias, st, who);
  }
  if (mode == kDgrad) {
    // a = dy (N,To,Ho,Wo,Cout), b = w bf16, out = dx (N,Ti,Hi,Wi,Cin); one class per residue of the input coordinate
    if (dgrad4_geom_ok(g, wrows) && !tc_env_int("MCG_TC_NOFUSE") && !tc_env_int("MCG_TC_TR") && !tc_env_int("MCG_TC_MT") &&
        !tc_env_int("MCG_TC_BN")) {
      bool taken = false;
      if ((rc = tc_dgrad4(g, a, b, out, bias, out_dtype, st, &taken))) return rc;
      if (taken) return 0;
    }
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
    P.ocols = g->Cin;
    P.chunks = g->Cout / 64;
    int ncls = cw * ch * ct, j = 0, max_taps = 0;
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
      if (P.tap_count[c] > max_taps) max_taps = P.tap_count[c];
    }
    P.nboxes = P.nbw * P.nbh * P.nbt * P.nbb;
    double plain_cost = 0;
    const TileCfg cfg = pick_tile(kDgrad, P.nboxes, ncls, g->Cin, max_taps * P.chunks, &plain_cost);
    if (g->kT >= 2 && g->sT == 1) {
      const TrPlan tr = plan_tr(EW, EH, ET, g->N, g->Cin, ncls, max_taps / g->kT, P.chunks, g->kT);
      if (tr_wanted(tr, plain_cost)) {
        const Box tb = tr.bx;
        P.BW = tb.w; P.BH = tb.h; P.BT = tb.t; P.BB = 1;
        P.nbw = ceil_div(EW, tb.w); P.nbh = ceil_div(EH, tb.h); P.nbt = ceil_div(ET, tb.t); P.nbb = g->N;
        P.nboxes = P.nbw * P.nbh * P.nbt * P.nbb;
        P.tr_kT = g->kT; P.tr_kstep = g->kH * g->kW; P.tr_rev = 1;
        P.tr_frame_bytes = tb.w * tb.h * 128; P.tr_ablock_bytes = (tb.t + g->kT - 1) * P.tr_frame_bytes;
        P.tr_aslots = tr.aslots; P.tr_bslots = tr.bslots;
        // groups: the (kh, kw) taps of each class; every kt belongs to it (ct = 1).  dy frame of tap kt = t + pT - kt.
        int jg = 0;
        for (int c = 0; c < ncls; ++c) {
          const int pw = c % cw, ph = (c / cw) % ch;
          P.tap_begin[c] = jg;
          for (int kh = 0; kh < g->kH; ++kh) {
            if ((ph + g->pH - kh) % ch) continue;
            for (int kw = 0; kw < g->kW; ++kw) {
              if ((pw + g->pW - kw) % cw) continue;
              P.taps[jg++] = TcTap{(int16_t)((pw + g->pW - kw) / cw), (int16_t)((ph + g->pH - kh) / ch),
                                   (int16_t)(g->pT - (g->kT - 1)), (int16_t)(kh * g->kW + kw)};
            }
          }
          P.tap_count[c] = jg - P.tap_begin[c];
        }
        if ((rc = act_map(&ma, a, g->Cout, g->Wo, g->Ho, g->To, g->N, tb.w, tb.h, tb.t + g->kT - 1, 1, 1, 1, 1))) return rc;
        P.ntn = g->Cin / tr.bn;
        P.ntiles = ncls * P.ntn;
        const int grid = split_units(P, (long long)P.ntiles * P.nboxes, 1);
        uint64_t d2[2] = {(uint64_t)P.Ktot, (uint64_t)wrows}, s2[1] = {(uint64_t)P.Ktot * 2};
        uint32_t b2[2] = {64, 64}, e2[2] = {1, 1};
        if ((rc = get_map(&mb, b, 2, d2, s2, b2, e2))) return rc;
        return launch_tr_cfg<kDgrad>(tr.bn, tr.mt, ma, mb, P, grid, out, bias, st, who);
      }
    }
    if ((rc = act_map(&ma, a, g->Cout, g->Wo, g->Ho, g->To, g->N, bx.w, bx.h, bx.t, bx.b, 1, 1, 1))) return rc;
    if (g->kT > 1 && ct == 1 && !tc_env_int("MCG_TC_NOSKIP")) {
      // input frame t reads dy frame t + pT - kt: near both ends of the clip most temporal taps fall outside dy (To < Ti)
      P.skip_t = 1; P.skip_nkt = g->kT; P.skip_dt0 = g->pT; P.skip_dts = -1; P.skip_aT = g->To;
      P.box_tn = tc_env_int("MCG_TC_STRIDED") > 0 || (tc_env_int("MCG_TC_STRIDED") == 0 && cfg.bn == 64);
      for (int c = 0; c < ncls; ++c) P.skip_per_kt[c] = P.tap_count[c] / g->kT;
    }
    P.ntn = g->Cin / cfg.bn;
    P.ntiles = ncls * P.ntn;
    const int grid = split_units(P, (long long)P.ntiles * P.nboxes, 1);
    uint64_t d2[2] = {(uint64_t)P.Ktot, (uint64_t)wrows}, s2[1] = {(uint64_t)P.Ktot * 2};
    uint32_t b2[2] = {64, 64}, e2[2] = {1, 1};
    if ((rc = get_map(&mb, b, 2, d2, s2, b2, e2))) return rc;
    return launch_tc_cfg<kDgrad>(cfg, ma, mb, P, grid, out, bias, st, who);
  }
  // ---- wgrad: a = x (N,Ti,Hi,Wi,Cin), b = dy (N,To,Ho,Wo,Cout), out = dw fp32 (Cout, taps*Cin), accumulated
  {
    Box bx = (win && win->T > 1) ? choose_box(64, g->Wo, g->Ho, g->To, 1) : choose_box(64, g->Wo, g->Ho, g->To, g->N);
    if (win) { P.win = 1; P.win_T = win->T; }
    P.BW = bx.w; P.BH = bx.h; P.BT = bx.t; P.BB = bx.b;
    P.nbw = ceil_div(g->Wo, bx.w); P.nbh = ceil_div(g->Ho, bx.h); P.nbt = ceil_div(g->To, bx.t); P.nbb = ceil_div(g->N, bx.b);
    P.a_mul_w = g->sW; P.a_mul_h = g->sH; P.a_mul_t = g->sT; P.a_add_w = -g->pW; P.a_add_h = -g->pH; P.a_add_t = -g->pT;
    P.chunks = g->Cin / 64;
    for (int kt = 0, j = 0; kt < g->kT; ++kt)
      for (int kh = 0; kh < g->kH; ++kh)
        for (int kw = 0; kw < g->kW; ++kw, ++j) P.taps[j] = TcTap{(int16_t)kw, (int16_t)kh, (int16_t)kt, (int16_t)j};
    const int slabs = taps * P.chunks;
    P.total_slabs = slabs;
    P.kreal = kreal > 0 ? kreal : P.Ktot;
    P.debug_skip_epi = tc_env_int("MCG_TC_DEBUG_NOEPI");
    P.total_boxes = P.nbw * P.nbh * P.nbt * P.nbb;
    // widest tile the channel counts allow: B (dy) is re-read once per row group, A (x) once per column tile, and the
    // fp32 red.add volume grows with the number of K splits, so few large tiles win; stream-K keeps every SM busy.
    TileCfg cfg{slabs >= 4 ? 2 : 1, 64};
    for (int c = 256; c >= 64; c /= 2)
      if (g->Cout % c == 0) { cfg.bn = c; break; }
    if (tc_env_int("MCG_TC_WMT")) cfg.mt = tc_env_int("MCG_TC_WMT");
    if (tc_env_int("MCG_TC_WBN") && g->Cout % tc_env_int("MCG_TC_WBN") == 0) cfg.bn = tc_env_int("MCG_TC_WBN");
    const int mgroups = ceil_div(slabs, 2 * cfg.mt);
    P.ntn = g->Cout / cfg.bn;
    P.ntiles = mgroups * P.ntn;
    const int ctas = split_units(P, (long long)P.ntiles * P.total_boxes, 8);   // at least 8 K steps per CTA
    if (win) {
      if ((rc = win_map(&ma, *win, bx.w, bx.h, win->T > 1 ? bx.t : bx.b))) return rc;
    } else if ((rc = act_map(&ma, a, g->Cin, g->Wi, g->Hi, g->Ti, g->N, bx.w, bx.h, bx.t, bx.b, g->sW, g->sH, g->sT))) {
      return rc;
    }
    if ((rc = act_map(&mb, b, g->Cout, g->Wo, g->Ho, g->To, g->N, bx.w, bx.h, bx.t, bx.b, 1, 1, 1))) return rc;
    return launch_tc_cfg<kWgrad>(cfg, ma, mb, P, ctas, out, nullptr, st, who);
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
  pdl_enter();
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
  pdl_enter();
  extern __shared__ __align__(16) unsigned short rows[];  // [kT*kH][L]: 8 zeros | Wi*CIN source values | 8 zeros
  const int runs = kT * kH;
  const int RW = Wi * CIN;                  // launcher guarantees RW % 8 == 0, pW*CIN <= 8, (KW-pW-1)*CIN <= 8
  const int L = RW + 16;
  const int K = runs * KW * CIN;
  int line = blockIdx.x;
  const int ho = line % Ho; line /= Ho;
  const int to = line % To;
  const long long n = line / To;
  {
    // source rows are copied with 16-byte loads (a row of the channels-last clip is contiguous and 16-byte aligned)
    const int vpr_src = RW / 8 + 2;         // vectors per staged row, pads included
    uint4* rows128 = reinterpret_cast<uint4*>(rows);
    for (int e = threadIdx.x; e < runs * vpr_src; e += blockDim.x) {
      const int run = e / vpr_src, j = e - run * vpr_src;
      const int kh = run % kH, kt = run / kH;
      const int ti = to * sT - pT + kt, hi = ho * sH - pH + kh;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (j > 0 && j <= RW / 8 && (unsigned)ti < (unsigned)Ti && (unsigned)hi < (unsigned)Hi)
        v = __ldg(reinterpret_cast<const uint4*>(x + (((n * Ti + ti) * Hi + hi) * (long long)Wi) * CIN) + (j - 1));
      rows128[e] = v;
    }
  }
  __syncthreads();
  // Each thread owns ONE 16-byte column group (8 consecutive k) and walks the output pixels of the line, so the
  // k -> (run, tap, channel) decode is done once; consecutive threads write consecutive 16-byte vectors.
  const int vpr = Kp / 8;                 // vectors per cols row
  const int wq_n = blockDim.x / vpr;      // pixels in flight per pass
  const int kv = threadIdx.x % vpr, wq = threadIdx.x / vpr;
  if (wq >= wq_n) return;
  int off[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = kv * 8 + i;
    off[i] = k < K ? (k / (KW * CIN)) * L + (8 - pW * CIN) + k % (KW * CIN) : -1;
  }
  const long long m0 = ((n * To + to) * Ho + ho) * (long long)Wo;
  uint4* dst = reinterpret_cast<uint4*>(cols + m0 * Kp);
  const int step = sW * CIN;
  for (int wo = wq; wo < Wo; wo += wq_n) {
    const int base = wo * step;
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t lo = off[2 * i] >= 0 ? rows[off[2 * i] + base] : 0u;
      const uint32_t hi = off[2 * i + 1] >= 0 ? rows[off[2 * i + 1] + base] : 0u;
      o[i] = lo | (hi << 16);
    }
    dst[wo * vpr + kv] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
// wp[co][Kp] = w[co][k] (k < K) else 0
__global__ void pad_rows_kernel(const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ wp, int rows, int K, int Kp) {
  pdl_enter();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows * Kp; i += gridDim.x * blockDim.x) {
    const int r = i / Kp, k = i % Kp;
    wp[i] = k < K ? w[(long long)r * K + k] : __float2bfloat16_rn(0.f);
  }
}
// wt[j = (tap, ci), zero-padded to Jp rows][co] = w[co][tap][ci]   (K-major B operand of the dgrad GEMM)
__global__ void transpose_jk_kernel(const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ wt, int Cout, int J,
                                    int Jp) {
  pdl_enter();
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
  pdl_enter();
  // Only the taps whose parity matches this input line can reach it: kt = (ti+pT) % sT + a*sT, kh = (hi+pH) % sH + b*sH.
  // Slot (a, b) of the staging buffer holds that run's Wo*kW*Cin values (zeros when the run falls outside the output);
  // every thread decides that for the vectors it copies, so there is no serial set-up phase.
  extern __shared__ __align__(16) unsigned short zs[];  // [slots][Wo][kW*Cin]
  const int RUN = kW * Cin;
  const int nsh = (kH + sH - 1) / sH, nst = (kT + sT - 1) / sT;
  const int slots = nst * nsh;
  int line = blockIdx.x;
  const int hi = line % Hi; line /= Hi;
  const int ti = line % Ti;
  const long long n = line / Ti;
  const int ktp = (ti + pT) % sT, khp = (hi + pH) % sH;
  const int vecs = Wo * RUN / 8;   // launcher guarantees (Wo * RUN) % 8 == 0
  {
    const uint4* zsrc = reinterpret_cast<const uint4*>(Z);
    uint4* zs128 = reinterpret_cast<uint4*>(zs);
    for (int e = threadIdx.x; e < slots * vecs; e += blockDim.x) {
      const int sl = e / vecs, i = e - sl * vecs;
      const int kt = ktp + (sl / nsh) * sT, kh = khp + (sl % nsh) * sH;
      const int tt = (ti + pT - kt) / sT, hh = (hi + pH - kh) / sH;   // exact when the numerators are >= 0
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (kt < kT && kh < kH && ti + pT - kt >= 0 && hi + pH - kh >= 0 && tt < To && hh < Ho) {
        const long long row = (n * To + tt) * Ho + hh;   // Z row block (times Wo)
        v = __ldg(zsrc + ((long long)(kt * kH + kh) * Mpix + row * Wo) * RUN / 8 + i);
      }
      zs128[e] = v;
    }
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
      for (int v = 0; v < slots; ++v) {
        const __nv_bfloat16_raw raw = {zs[(v * Wo + wo) * RUN + kw * Cin + ci]};
        acc += __bfloat162float(__nv_bfloat16(raw));
      }
    }
    if (out_f32) reinterpret_cast<float*>(dx)[obase + o] = acc;
    else reinterpret_cast<__nv_bfloat16*>(dx)[obase + o] = __float2bfloat16_rn(acc);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Narrow-Cin data gradient as ONE stride-1 convolution over dy ("merged classes").  With ti = sT*l + e, hi = sH*i + a,
// wi = sW*j + b the taps that reach class (e,a,b) read dy at (l+dt, i+dh, j+dw) for a small window of offsets that is
// the same for every class, so
//   Zm[n,l,i,j][(e,a,b,ci)] = sum_{dt,dh,dw,co} dy[n,l+dt,i+dh,j+dw,co] * Wm[(e,a,b,ci)][(dt,dh,dw)][co]
// is an ordinary fprop with 64 (padded) output columns, where Wm holds w[co,kt,kh,kw,ci] at the offsets that class uses
// and zeros elsewhere; dx is Zm with its column blocks spread back over the stride grid (depth-to-space).  dy is read
// window-size times from L2 and nothing of size M x taps*Cin ever exists (the Z = dy.w^T + col2im form wrote and re-read
// 179 MB for Dv.dc1).
struct MergedGeom {
  int Cin, Cout, kT, kH, kW, sT, sH, sW, pT, pH, pW;
  int nT, nH, nW;            // window extents (offsets dmin .. dmin+n-1)
  int dt_min, dh_min, dw_min;
};
// Wm[row = ((e*sH + a)*sW + b)*Cin + ci (zero rows up to 64)][tap' = (u*nH + v)*nW + q][co]
__global__ void merged_weights_kernel(const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ wm, MergedGeom G) {
  pdl_enter();
  const int taps2 = G.nT * G.nH * G.nW;
  const long long total = 64LL * taps2 * G.Cout;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(idx % G.Cout);
    long long r = idx / G.Cout;
    const int tap2 = (int)(r % taps2);
    const int row = (int)(r / taps2);
    __nv_bfloat16 val = __float2bfloat16_rn(0.f);
    if (row < G.sT * G.sH * G.sW * G.Cin) {
      const int ci = row % G.Cin;
      int cls = row / G.Cin;
      const int b = cls % G.sW; cls /= G.sW;
      const int a = cls % G.sH;
      const int e = cls / G.sH;
      const int q = tap2 % G.nW, v = (tap2 / G.nW) % G.nH, u = tap2 / (G.nW * G.nH);
      const int kt = e + G.pT - G.sT * (u + G.dt_min), kh = a + G.pH - G.sH * (v + G.dh_min), kw = b + G.pW - G.sW * (q + G.dw_min);
      if (kt >= 0 && kt < G.kT && kh >= 0 && kh < G.kH && kw >= 0 && kw < G.kW)
        val = w[((((long long)co * G.kT + kt) * G.kH + kh) * G.kW + kw) * G.Cin + ci];
    }
    wm[idx] = val;
  }
}
// dx[n,ti,hi,wi,ci] = bias[ci] + Zm[n, ti/sT, hi/sH, wi/sW][((ti%sT*sH + hi%sH)*sW + wi%sW)*Cin + ci].
// One thread per (cell, e, a): the sW*Cin values of classes (e, a, 0..sW-1) are contiguous in the Zm row AND in dx
// (sW neighbouring pixels of one line), so the thread moves one short contiguous run.
__global__ void __launch_bounds__(256) depth_to_space_kernel(const __nv_bfloat16* __restrict__ zm, const float* __restrict__ bias,
                                                             void* __restrict__ dx, int out_f32, long long cells, int Cin, int Ti,
                                                             int Hi, int Wi, int L, int I, int J, int sT, int sH, int sW, int zcols) {
  pdl_enter();
  const int RUN = sW * Cin, sub = sT * sH;
  const long long total = cells * sub;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int ea = (int)(idx % sub);
    long long r = idx / sub;
    const int j = (int)(r % J); r /= J;
    const int i = (int)(r % I); r /= I;
    const int l = (int)(r % L);
    const long long n = r / L;
    const int e = ea / sH, a = ea % sH;
    const int ti = l * sT + e, hi = i * sH + a, wi0 = j * sW;
    if (ti >= Ti || hi >= Hi) continue;
    const __nv_bfloat16* src = zm + (idx / sub) * zcols + ea * RUN;
    const long long dst = (((n * Ti + ti) * Hi + hi) * (long long)Wi + wi0) * Cin;
    for (int q = 0; q < RUN; ++q) {
      if (wi0 + q / Cin >= Wi) break;
      const float v = __bfloat162float(src[q]) + (bias ? bias[q % Cin] : 0.f);
      if (out_f32) reinterpret_cast<float*>(dx)[dst + q] = v;
      else reinterpret_cast<__nv_bfloat16*>(dx)[dst + q] = __float2bfloat16_rn(v);
    }
  }
}
static int floor_div(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }
static bool merged_geom(const mcg_conv_geom* g, MergedGeom* G) {
  if (g->sT * g->sH * g->sW * g->Cin > 64) return false;
  *G = MergedGeom{g->Cin, g->Cout, g->kT, g->kH, g->kW, g->sT, g->sH, g->sW, g->pT, g->pH, g->pW, 0, 0, 0, 0, 0, 0};
  // offsets d = (phase + p - k) / s over phases 0..s-1 and taps 0..k-1 (only exact divisions are real taps, the
  // window just has to cover them)
  G->dt_min = floor_div(g->pT - (g->kT - 1), g->sT); G->nT = floor_div(g->sT - 1 + g->pT, g->sT) - G->dt_min + 1;
  G->dh_min = floor_div(g->pH - (g->kH - 1), g->sH); G->nH = floor_div(g->sH - 1 + g->pH, g->sH) - G->dh_min + 1;
  G->dw_min = floor_div(g->pW - (g->kW - 1), g->sW); G->nW = floor_div(g->sW - 1 + g->pW, g->sW) - G->dw_min + 1;
  return G->nT * G->nH * G->nW <= 64;
}

static long long round_up(long long a, long long b) { return (a + b - 1) / b * b; }

// ---- window-view path of the 3-channel layers (see TcWin): layout helpers -----------------------------------------
// xk[f][ho][wq][kh][0..4) = x[f][ho*sH + kh - pH][wq - pW][c] inside the image and for c < C, 0 elsewhere
__global__ void __launch_bounds__(256) interleave_rows_kernel(const __nv_bfloat16* __restrict__ x, uint4* __restrict__ xk,
                                                              long long frames, int H, int W, int C, int Ho, int Wq, int sH,
                                                              int pH, int pW) {
  pdl_enter();
  // one thread per (frame, ho, wq): reads the pixel from the kH = 4 rows (consecutive threads read consecutive pixels of a
  // row) and writes its 32 contiguous bytes [kh][c4]
  const long long total = frames * Ho * Wq;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int wq = (int)(i % Wq);
    const long long r = i / Wq;
    const int ho = (int)(r % Ho);
    const long long f = r / Ho;
    const int w = wq - pW;
    __align__(16) __nv_bfloat16 v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = __float2bfloat16_rn(0.f);
    if ((unsigned)w < (unsigned)W) {
#pragma unroll
      for (int kh = 0; kh < 4; ++kh) {
        const int h = ho * sH + kh - pH;
        if ((unsigned)h < (unsigned)H) {
          const __nv_bfloat16* src = x + ((f * H + h) * (long long)W + w) * C;
          for (int c = 0; c < C; ++c) v[kh * 4 + c] = src[c];
        }
      }
    }
    xk[2 * i] = *reinterpret_cast<const uint4*>(v);
    xk[2 * i + 1] = *reinterpret_cast<const uint4*>(v + 8);
  }
}
// The same, one block per (frame, ho): the kH = 4 source rows are staged in shared memory with 16-byte loads (a row of
// the channels-last image is contiguous), then every thread assembles whole 32-byte output pixels.  W*C % 8 == 0.
__global__ void __launch_bounds__(128) interleave_rows_line_kernel(const __nv_bfloat16* __restrict__ x, uint4* __restrict__ xk, int H,
                                                                   int W, int C, int Ho, int Wq, int sH, int pH, int pW) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned short srow[];     // [4][W*C]
  const int RW = W * C, vpr = RW / 8;
  const int ho = blockIdx.x % Ho;
  const long long f = blockIdx.x / Ho;
  uint4* s128 = reinterpret_cast<uint4*>(srow);
  for (int e = threadIdx.x; e < 4 * vpr; e += blockDim.x) {
    const int kh = e / vpr, j = e - kh * vpr;
    const int h = ho * sH + kh - pH;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if ((unsigned)h < (unsigned)H) v = __ldg(reinterpret_cast<const uint4*>(x + (f * H + h) * (long long)RW) + j);
    s128[e] = v;
  }
  __syncthreads();
  uint4* dst = xk + ((f * Ho + ho) * (long long)Wq) * 2;
  for (int wq = threadIdx.x; wq < Wq; wq += blockDim.x) {
    const int w = wq - pW;
    __align__(16) unsigned short v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = 0;
    if ((unsigned)w < (unsigned)W) {
#pragma unroll
      for (int kh = 0; kh < 4; ++kh)
        for (int c = 0; c < C; ++c) v[kh * 4 + c] = srow[kh * RW + w * C + c];
    }
    dst[2 * wq] = *reinterpret_cast<const uint4*>(v);
    dst[2 * wq + 1] = *reinterpret_cast<const uint4*>(v + 8);
  }
}
// Patch rows: xp[f][ho][wo][kw][kh][c4] = x[f][ho*sH + kh - pH][wo*sW + kw - pW][c] — the 4 x 4 spatial window of every
// OUTPUT pixel as one 128-byte row (a 2-D im2col per frame: 73 MB for Dv.dc1 against 179 MB for the full 3-D im2col; the
// temporal taps stay implicit).  The rows are 128-byte aligned, which the overlapping window view's are not (its rows
// start every 64 B) — an alternative to the window view, opt-in (MCG_TC_PATCHROWS=1): measured slower.
// One block per (frame, ho); W*C % 8 == 0.
__global__ void __launch_bounds__(128) patch_rows_line_kernel(const __nv_bfloat16* __restrict__ x, uint4* __restrict__ xp, int H, int W,
                                                              int C, int Ho, int Wo, int sH, int sW, int pH, int pW) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned short srow[];     // [4][W*C]
  const int RW = W * C, vpr = RW / 8;
  const int ho = blockIdx.x % Ho;
  const long long f = blockIdx.x / Ho;
  uint4* s128 = reinterpret_cast<uint4*>(srow);
  for (int e = threadIdx.x; e < 4 * vpr; e += blockDim.x) {
    const int kh = e / vpr, j = e - kh * vpr;
    const int h = ho * sH + kh - pH;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if ((unsigned)h < (unsigned)H) v = __ldg(reinterpret_cast<const uint4*>(x + (f * H + h) * (long long)RW) + j);
    s128[e] = v;
  }
  __syncthreads();
  // a thread writes one (wo, kw) quarter row = 32 B = [kh][c4]; consecutive threads write consecutive 32-byte pieces
  uint4* dst = xp + ((f * Ho + ho) * (long long)Wo) * 8;
  for (int i = threadIdx.x; i < Wo * 4; i += blockDim.x) {
    const int wo = i >> 2, kw = i & 3;
    const int w = wo * sW + kw - pW;
    __align__(16) unsigned short v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = 0;
    if ((unsigned)w < (unsigned)W) {
#pragma unroll
      for (int kh = 0; kh < 4; ++kh)
        for (int c = 0; c < C; ++c) v[kh * 4 + c] = srow[kh * RW + w * C + c];
    }
    dst[2 * i] = *reinterpret_cast<const uint4*>(v);
    dst[2 * i + 1] = *reinterpret_cast<const uint4*>(v + 8);
  }
}
// w4[co][kt][kw][kh][0..4) = w[co][kt][kh][kw][c < C], 0 beyond: the K order of the window rows
__global__ void pad_w4_kernel(const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ w4, int rows_kt, int C) {
  pdl_enter();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows_kt * 64; i += gridDim.x * blockDim.x) {
    const int c = i & 3, kh = (i >> 2) & 3, kw = (i >> 4) & 3, rk = i >> 6;
    w4[i] = c < C ? w[((long long)rk * 16 + kh * 4 + kw) * C + c] : __float2bfloat16_rn(0.f);
  }
}
// dw[co][kt][kh][kw][c] += dw4[co][kt][kw][kh][c]  (the padded channel is dropped); atomic: real and fake branches share dw
__global__ void unpad_dw4_kernel(const float* __restrict__ dw4, float* __restrict__ dw, int rows_kt, int C) {
  pdl_enter();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows_kt * 16 * C; i += gridDim.x * blockDim.x) {
    const int c = i % C, tap = (i / C) & 15, rk = i / (16 * C);
    const int kh = tap >> 2, kw = tap & 3;
    atomicAdd(dw + i, dw4[(rk * 16 + kw * 4 + kh) * 4 + c]);
  }
}
static bool win_geom_ok(const mcg_conv_geom* g) {
  return g->Cin <= 4 && g->Cout % 64 == 0 && g->kH == 4 && g->kW == 4 && g->sW == 2 && g->sH >= 1 && g->sH <= 2 && g->sT == 1 &&
         g->pT == 0 && g->kT <= 16 && !tc_env_int("MCG_TC_NOWIN");
}
struct WinLayout { int Wq; size_t xq_bytes, w4_bytes, dw4_bytes; };
static WinLayout win_layout(const mcg_conv_geom* g) {
  WinLayout L;
  L.Wq = g->Wi + g->pW > (g->Wo - 1) * g->sW + g->kW ? g->Wi + g->pW : (g->Wo - 1) * g->sW + g->kW;
  const long long win_b = (long long)g->N * g->Ti * g->Ho * L.Wq * 32, patch_b = (long long)g->N * g->Ti * g->Ho * g->Wo * 128;
  L.xq_bytes = (size_t)round_up(win_b > patch_b ? win_b : patch_b, 1024);
  L.w4_bytes = (size_t)round_up((long long)g->Cout * g->kT * 64 * 2, 1024);
  L.dw4_bytes = (size_t)round_up((long long)g->Cout * g->kT * 64 * 4, 1024);
  return L;
}

bool tc_small_supported(const mcg_conv_geom* g) {
  return g->Cin <= 16 && g->Cout % 64 == 0 && g->sT <= 2 && g->sH <= 2 && g->sW <= 2 && g->kT * g->kH * g->kW <= 64 &&
         g->kW % 4 == 0;
}
size_t tc_small_workspace(const mcg_conv_geom* g) {
  // the largest of what the paths this geometry can take need (tc_conv_small picks by the same predicates)
  const long long M = (long long)g->N * g->To * g->Ho * g->Wo;
  const long long K = (long long)g->kT * g->kH * g->kW * g->Cin, Kp = round_up(K, 64);
  const size_t cols_need = (size_t)(round_up(M * Kp * 2, 1024) + round_up((long long)g->Cout * Kp * 2, 1024) + 4096);
  MergedGeom G;
  const bool merged = merged_geom(g, &G) && !getenv("MCG_NO_MERGED_DGRAD"), win = win_geom_ok(g);
  size_t need = 0;
  if (win) {                       // fprop / wgrad read a re-laid copy of x through a window view: no im2col matrix
    const WinLayout L = win_layout(g);
    need = L.xq_bytes + L.w4_bytes + L.dw4_bytes + 4096;
  }
  if (!win || !merged) need = cols_need > need ? cols_need : need;   // im2col GEMM (fprop / wgrad) or Z = dy.w^T + col2im (dgrad)
  if (merged) {                    // dgrad: Zm (one 64-column row per stride cell) + the merged weights
    const long long cells = (long long)g->N * ceil_div(g->Ti, g->sT) * ceil_div(g->Hi, g->sH) * ceil_div(g->Wi, g->sW);
    const size_t m = (size_t)(round_up(cells * 128, 1024) + round_up(64LL * G.nT * G.nH * G.nW * g->Cout * 2, 1024) + 4096);
    if (m > need) need = m;
  }
  return need;
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
    MergedGeom G;
    if (merged_geom(g, &G) && !getenv("MCG_NO_MERGED_DGRAD")) {   // (same predicate as tc_small_workspace)
      const int L = ceil_div(g->Ti, g->sT), I = ceil_div(g->Hi, g->sH), J = ceil_div(g->Wi, g->sW);
      const long long cells = (long long)g->N * L * I * J;
      __nv_bfloat16* zm = reinterpret_cast<__nv_bfloat16*>(base);
      __nv_bfloat16* wm = reinterpret_cast<__nv_bfloat16*>(base + round_up(cells * 128, 1024));
      pdl(merged_weights_kernel, 128, 256, 0, st)((const __nv_bfloat16*)b, wm, G);
      MCG_CHECK_LAUNCH(who);
      // stride-1 fprop over dy: "input" = dy, window (nT,nH,nW), zero padding -dmin (the far side is TMA out-of-bounds fill)
      mcg_conv_geom g2 = {g->N, g->Cout, 64, g->To, g->Ho, g->Wo, L, I, J, G.nT, G.nH, G.nW, 1, 1, 1, -G.dt_min, -G.dh_min, -G.dw_min};
      if (g->sT * g->sH * g->sW * g->Cin <= 32 && G.nT * G.nH * G.nW >= 18 && !tc_env_int("MCG_TC_NOD2S")) {
        // the convolution's epilogue writes every cell's values straight to their pixels of dx (+ bias): no Zm, no extra
        // pass.  Only where the K loop is long enough to hide the scattered 2-byte stores (Dv.dc1: 36 taps, 0.189 ->
        // 0.171 ms); the 2-D layers' 9-tap loop is epilogue-bound already (G.dc5: 0.101 -> 0.119 ms with it)
        const int d2s[7] = {g->Cin, g->sT, g->sH, g->sW, g->Ti, g->Hi, g->Wi};
        return tc_conv(kFprop, &g2, a, wm, out, bias, out_dtype, st, 0, 0, 0, 0, nullptr, d2s);
      }
      // Zm keeps only the columns that carry classes (rounded up to 8): 32 B instead of 128 B per cell for the 2-D layers
      const int zcols = (int)round_up((long long)g->sT * g->sH * g->sW * g->Cin, 8);
      if ((rc = tc_conv(kFprop, &g2, a, wm, zm, nullptr, MCG_BF16, st, 0, 0, 0, 0, nullptr, nullptr, zcols))) return rc;
      long long nb = (cells * g->sT * g->sH + 255) / 256;
      if (nb > (long long)num_sms() * 16) nb = (long long)num_sms() * 16;
      pdl(depth_to_space_kernel, (unsigned)nb, 256, 0, st)(zm, bias, out, out_dtype == MCG_F32, cells, g->Cin, g->Ti, g->Hi, g->Wi,
                                                         L, I, J, g->sT, g->sH, g->sW, zcols);
      MCG_CHECK_LAUNCH(who);
      return 0;
    }
    if (M > 0x7fffffffLL) MCG_FAIL(MCG_ERR_UNSUPPORTED, "%s: too many pixels", who);
    pdl(transpose_jk_kernel, 64, 256, 0, st)((const __nv_bfloat16*)b, wpad, g->Cout, K, Kp);
    MCG_CHECK_LAUNCH(who);
    mcg_conv_geom g2 = {1, g->Cout, Kp, 1, 1, (int)M, 1, 1, (int)M, 1, 1, 1, 1, 1, 1, 0, 0, 0};
    const int RUN = g->kW * g->Cin;
    if ((rc = tc_conv(kFprop, &g2, a, wpad, cols, nullptr, MCG_BF16, st, 0, RUN, K))) return rc;
    const long long lines = (long long)g->N * g->Ti * g->Hi;
    const size_t smem = (size_t)ceil_div(g->kT, g->sT) * ceil_div(g->kH, g->sH) * g->Wo * g->kW * g->Cin * 2;
    if (lines > 0x7fffffffLL || smem > 48 * 1024 || (g->Wo * g->kW * g->Cin) % 8) MCG_FAIL(MCG_ERR_UNSUPPORTED, "%s: line too large", who);
    pdl(col2im_line_kernel, (unsigned)lines, 128, smem, st)(cols, bias, out, out_dtype == MCG_F32, M, g->Cin, g->Ti, g->Hi, g->Wi,
                                                           g->To, g->Ho, g->Wo, g->kT, g->kH, g->kW, g->sT, g->sH, g->sW, g->pT,
                                                           g->pH, g->pW);
    MCG_CHECK_LAUNCH(who);
    return 0;
  }
  if (win_geom_ok(g)) {
    // fprop / wgrad through the window view (TcWin): pad x once (reused by wgrad: cols_valid), then the implicit-GEMM
    // kernels with kT taps of 64 "channels" = (kh, kw, c4)
    const WinLayout L = win_layout(g);
    uint8_t* xq = base;
    __nv_bfloat16* w4 = reinterpret_cast<__nv_bfloat16*>(base + L.xq_bytes);
    float* dw4 = reinterpret_cast<float*>(base + L.xq_bytes + L.w4_bytes);
    const long long frames = (long long)g->N * g->Ti;
    if (frames * g->Ho > 0x7fffffffLL) MCG_FAIL(MCG_ERR_UNSUPPORTED, "%s: too many rows", who);
    const int rows_kt = g->Cout * g->kT, rows_taps = rows_kt * 16;
    // default: the row-interleaved copy read through the overlapping window view.  MCG_TC_PATCHROWS=1: patch rows (one
    // 128-byte aligned row per output pixel, twice the bytes) through an ordinary activation map — measured slower
    // (Dv.dc1 fprop 0.089 against 0.073 ms), kept as the A/B that showed the window rows' 64-byte alignment is not a cost
    const bool patches = tc_env_int("MCG_TC_PATCHROWS") && (g->Wi * g->Cin) % 8 == 0 && (size_t)g->Wi * g->Cin * 8 <= 48 * 1024;
    if (patches) {
      if (!cols_valid) {
        pdl(patch_rows_line_kernel, (unsigned)(frames * g->Ho), 128, (size_t)g->Wi * g->Cin * 8, st)(
            (const __nv_bfloat16*)a, reinterpret_cast<uint4*>(xq), g->Hi, g->Wi, g->Cin, g->Ho, g->Wo, g->sH, g->sW, g->pH, g->pW);
        MCG_CHECK_LAUNCH(who);
      }
      // xp is a plain channels-last tensor (N, Ti, Ho, Wo, 64): a kT x 1 x 1 convolution over it
      mcg_conv_geom gp = {g->N, 64, g->Cout, g->Ti, g->Ho, g->Wo, g->To, g->Ho, g->Wo, g->kT, 1, 1, 1, 1, 1, 0, 0, 0};
      if (mode == kFprop) {
        pdl(pad_w4_kernel, 32, 256, 0, st)((const __nv_bfloat16*)b, w4, rows_kt, g->Cin);
        MCG_CHECK_LAUNCH(who);
        return tc_conv(kFprop, &gp, xq, w4, out, bias, out_dtype, st);
      }
      cudaError_t e = cudaMemsetAsync(dw4, 0, (size_t)rows_taps * 4 * sizeof(float), st);
      if (e != cudaSuccess) MCG_FAIL((int)e, "%s: cudaMemsetAsync: %s", who, cudaGetErrorString(e));
      if ((rc = tc_conv(kWgrad, &gp, xq, b, dw4, nullptr, MCG_F32, st))) return rc;
      pdl(unpad_dw4_kernel, 32, 256, 0, st)(dw4, reinterpret_cast<float*>(out), rows_kt, g->Cin);
      MCG_CHECK_LAUNCH(who);
      return 0;
    }
    if (!cols_valid) {
      if ((g->Wi * g->Cin) % 8 == 0 && frames * g->Ho < 0x7fffffffLL && (size_t)g->Wi * g->Cin * 8 <= 48 * 1024) {
        pdl(interleave_rows_line_kernel, (unsigned)(frames * g->Ho), 128, (size_t)g->Wi * g->Cin * 8, st)(
            (const __nv_bfloat16*)a, reinterpret_cast<uint4*>(xq), g->Hi, g->Wi, g->Cin, g->Ho, L.Wq, g->sH, g->pH, g->pW);
      } else {
      long long nb = (frames * g->Ho * L.Wq + 255) / 256;
      if (nb > (long long)num_sms() * 16) nb = (long long)num_sms() * 16;
      pdl(interleave_rows_kernel, (unsigned)nb, 256, 0, st)((const __nv_bfloat16*)a, reinterpret_cast<uint4*>(xq), frames, g->Hi,
                                                            g->Wi, g->Cin, g->Ho, L.Wq, g->sH, g->pH, g->pW);
      }
      MCG_CHECK_LAUNCH(who);
    }
    const TcWin win{xq, L.Wq, g->Ho, g->Wo, (int)frames, g->Ti, g->sW};
    mcg_conv_geom g2 = {g->N, 64, g->Cout, g->Ti, g->Ho, g->Wo, g->To, g->Ho, g->Wo, g->kT, 1, 1, 1, 1, 1, 0, 0, 0};
    if (mode == kFprop) {
      pdl(pad_w4_kernel, 32, 256, 0, st)((const __nv_bfloat16*)b, w4, rows_kt, g->Cin);
      MCG_CHECK_LAUNCH(who);
      return tc_conv(kFprop, &g2, xq, w4, out, bias, out_dtype, st, 0, 0, 0, 0, &win);
    }
    cudaError_t e = cudaMemsetAsync(dw4, 0, (size_t)rows_taps * 4 * sizeof(float), st);
    if (e != cudaSuccess) MCG_FAIL((int)e, "%s: cudaMemsetAsync: %s", who, cudaGetErrorString(e));
    if ((rc = tc_conv(kWgrad, &g2, xq, b, dw4, nullptr, MCG_F32, st, 0, 0, 0, 0, &win))) return rc;
    pdl(unpad_dw4_kernel, 32, 256, 0, st)(dw4, reinterpret_cast<float*>(out), rows_kt, g->Cin);
    MCG_CHECK_LAUNCH(who);
    return 0;
  }
  // fprop / wgrad: im2col, then a GEMM over a 1-D line of M pixels with Kp channels
  const void* x = a;
  if (!cols_valid) {
    const long long lines = (long long)g->N * g->To * g->Ho;
    const size_t smem = (size_t)g->kT * g->kH * (g->Wi * g->Cin + 16) * 2;
    const bool line_ok = lines < 0x7fffffffLL && smem <= 48 * 1024 && (g->Wi * g->Cin) % 8 == 0 && g->pW * g->Cin <= 8 &&
                         (g->kW - g->pW - 1) * g->Cin <= 8 && (g->Wo - 1) * g->sW - g->pW + g->kW - 1 <= g->Wi + (8 / g->Cin) - 1;
    if (g->kW == 4 && g->Cin == 3 && line_ok)
      pdl(im2col_line_kernel<3, 4>, (unsigned)lines, 128, smem, st)((const __nv_bfloat16*)x, cols, Kp, g->Ti, g->Hi, g->Wi, g->To, g->Ho,
                                                                   g->Wo, g->kT, g->kH, g->sT, g->sH, g->sW, g->pT, g->pH, g->pW);
    else if (g->kW == 4 && g->Cin == 1 && line_ok)
      pdl(im2col_line_kernel<1, 4>, (unsigned)lines, 128, smem, st)((const __nv_bfloat16*)x, cols, Kp, g->Ti, g->Hi, g->Wi, g->To, g->Ho,
                                                                   g->Wo, g->kT, g->kH, g->sT, g->sH, g->sW, g->pT, g->pH, g->pW);
    else
      pdl(im2col_kernel, num_sms() * 16, 256, 0, st)((const __nv_bfloat16*)x, cols, M, Kp, g->Cin, g->Ti, g->Hi, g->Wi, g->To, g->Ho,
                                                     g->Wo, g->kT, g->kH, g->kW, g->sT, g->sH, g->sW, g->pT, g->pH, g->pW);
  }
  MCG_CHECK_LAUNCH(who);
  if (M > 0x7fffffffLL) MCG_FAIL(MCG_ERR_UNSUPPORTED, "%s: too many pixels", who);
  mcg_conv_geom g2 = {1, Kp, g->Cout, 1, 1, (int)M, 1, 1, (int)M, 1, 1, 1, 1, 1, 1, 0, 0, 0};
  if (mode == kFprop) {
    const void* wk = b;
    if (Kp != K) {
      pdl(pad_rows_kernel, 32, 256, 0, st)((const __nv_bfloat16*)b, wpad, g->Cout, K, Kp);
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
  if (g && (impl & 0xff) == MCG_IMPL_TC && tc_supported(g)) return tc_splitk_workspace(g);   // fprop split-K scratch (else 0)
  return 0;
}

int mcg_conv_fprop(const mcg_conv_geom* g, const void* x, const void* w, const float* bias, void* y, int dtype,
                   int out_dtype, int impl, void* workspace, size_t workspace_bytes, void* stream) {
  if (!g || !x || !w || !y) MCG_FAIL(MCG_ERR_SHAPE, "mcg_conv_fprop: null pointer");
  const bool cols_valid = (impl & MCG_FLAG_COLS_VALID) != 0;
  const int wrows = (impl >> 16) & 0xffff;
  impl &= 0xff;
  if (impl == MCG_IMPL_TC) {
    if (dtype != MCG_BF16) MCG_FAIL(MCG_ERR_UNSUPPORTED, "mcg_conv_fprop(tc): activations must be bf16");
    if (!tc_supported(g) && tc_small_supported(g))
      return tc_conv_small(0, g, x, w, y, bias, out_dtype, workspace, workspace_bytes, as_stream(stream), cols_valid);
    // split-K needs its scratch tensor; a caller that passes none gets the unsplit kernel
    float* scratch = nullptr;
    if (workspace && workspace_bytes >= tc_splitk_workspace(g) && tc_splitk_workspace(g) > 0)
      scratch = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    return tc_conv(0, g, x, w, y, bias, out_dtype, as_stream(stream), 0, 0, 0, wrows, nullptr, nullptr, 0, scratch);
  }
  if (wrows) MCG_FAIL(MCG_ERR_UNSUPPORTED, "mcg_conv_fprop: MCG_W_ROWS needs MCG_IMPL_TC");
  return simt_conv(0, g, x, nullptr, (const float*)w, bias, y, dtype, out_dtype, 0, as_stream(stream));
}

int mcg_conv_dgrad(const mcg_conv_geom* g, const void* dy, const void* w, const float* bias, void* dx, int dtype,
                   int out_dtype, int accumulate, int impl, void* workspace, size_t workspace_bytes, void* stream) {
  if (!g || !dy || !w || !dx) MCG_FAIL(MCG_ERR_SHAPE, "mcg_conv_dgrad: null pointer");
  const int wrows = (impl >> 16) & 0xffff;
  impl &= 0xff;
  if (impl == MCG_IMPL_TC) {
    if (dtype != MCG_BF16) MCG_FAIL(MCG_ERR_UNSUPPORTED, "mcg_conv_dgrad(tc): activations must be bf16");
    if (accumulate) MCG_FAIL(MCG_ERR_UNSUPPORTED, "mcg_conv_dgrad(tc): accumulate not supported");
    if (!tc_supported(g) && tc_small_supported(g))
      return tc_conv_small(1, g, dy, w, dx, bias, out_dtype, workspace, workspace_bytes, as_stream(stream), false);
    return tc_conv(1, g, dy, w, dx, bias, out_dtype, as_stream(stream), 0, 0, 0, wrows);
  }
  if (wrows) MCG_FAIL(MCG_ERR_UNSUPPORTED, "mcg_conv_dgrad: MCG_W_ROWS needs MCG_IMPL_TC");
  return simt_conv(1, g, dy, nullptr, (const float*)w, bias, dx, dtype, out_dtype, accumulate, as_stream(stream));
}

int mcg_conv_wgrad(const mcg_conv_geom* g, const void* x, const void* dy, float* dw, int dtype, int impl, void* workspace,
                   size_t workspace_bytes, void* stream) {
  if (!g || !x || !dy || !dw) MCG_FAIL(MCG_ERR_SHAPE, "mcg_conv_wgrad: null pointer");
  const bool cols_valid = (impl & MCG_FLAG_COLS_VALID) != 0;
  const int wrows = (impl >> 16) & 0xffff;
  impl &= 0xff;
  if (wrows && impl != MCG_IMPL_TC) MCG_FAIL(MCG_ERR_UNSUPPORTED, "mcg_conv_wgrad: MCG_W_ROWS needs MCG_IMPL_TC");
  if (impl == MCG_IMPL_TC) {
    if (dtype != MCG_BF16) MCG_FAIL(MCG_ERR_UNSUPPORTED, "mcg_conv_wgrad(tc): activations must be bf16");
    if (!tc_supported(g) && tc_small_supported(g))
      return tc_conv_small(2, g, x, dy, dw, nullptr, MCG_F32, workspace, workspace_bytes, as_stream(stream), cols_valid);
    return tc_conv(2, g, x, dy, dw, nullptr, MCG_F32, as_stream(stream), 0, 0, 0, wrows);
  }
  return simt_conv(2, g, dy, x, nullptr, nullptr, dw, dtype, MCG_F32, 1, as_stream(stream));
}

int mcg_set_tc_sm_limit(int sms) {
  if (sms < 0) MCG_FAIL(MCG_ERR_SHAPE, "mcg_set_tc_sm_limit: %d < 0", sms);
  g_tc_sm_limit.store(sms, std::memory_order_relaxed);
  return 0;
}
int mcg_get_tc_sm_limit(void) { return tc_sms(); }

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
