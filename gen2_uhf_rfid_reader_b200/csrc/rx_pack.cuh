// rx_pack.cuh -- the roofline kernel of the reference configuration (decimation 5, 25 taps, rings inside one tile):
// same arithmetic as rx_fused_split.cuh, re-scheduled so that the exact replay of the gate's float running sums
// (avg_ampl, gate_impl.cc:131; dc_est, gate_impl.cc:141) costs one warp per *CTA* instead of one warp per segment.
//
// A CTA owns G <= kPMaxSeg capture segments and runs them in lockstep, one 128-sample tile per step:
//   tile warp g (G of them)  everything of segment g that is lane-parallel or sparse: TMA bulk copy of the raw tile,
//                            block-sum matched filter, exact |y|, amplitude / DC ring differences (P1);
//                            thresholds by ballot, the edge / pulse state machine on 128-bit masks, the DC-ring
//                            differences around gate activity (P3); window emission y - dc_est (E).
//   chain warp (one)         the order-dependent part of ALL segments of the CTA at once: lane 8*c + g replays running
//                            sum c (0 avg_ampl, 1 dc_est.re, 2 dc_est.im) of segment g -- up to 24 dependent FADD
//                            chains per instruction instead of 3, at the same FADD latency per step.
//   decoder warp g           decodes segment g's windows while they are still being gated (streaming, as before).
// Two named barriers per step (X: tiles ready -> chain, Y: sums ready -> tile warps; a parked warp costs no issue
// slots) replace every polled hand-off of the split kernel; tile warps and the chain warp overlap one step apart:
//   step i, tile warp:  P1(i) | arrive X(i) | sync Y(i-1) | E(i-3) | P3(i-1)
//   step i, chain warp: sync X(i) | avg_ampl over tile i, dc_est over the closed samples of tile i-2 | arrive Y(i)
// Shared memory per segment: 3 raw stages (15 KB), a 4-tile time ring of y and |y| (6 KB), ring snapshot, decoder stage;
// per CTA: the running-sum buffers (2 + 4x2 per segment, 560 B each, skewed so the chain warp's 128-bit accesses are
// bank-conflict free).  HBM traffic is unchanged: every raw sample is read once, 64 B are written per window.
#pragma once

#include "rx_fused_split.cuh"

namespace rfid_b200 {

constexpr int kPS = 4;                       // tiles in the time ring (tile i, lookback i-1, fix-ups i-2, emission i-3)
constexpr int kPRing = kPS * kTT;
constexpr int kPRawStages = 3;
constexpr int kPMaxSeg = 7;                  // segments per CTA (chain lanes 8*c + g, g < 8)
constexpr int kPChainBuf = kTT + 12;         // floats per running-sum buffer: read-ahead pad; 560 B = 48 mod 128
constexpr int kPackMaxThreads = 32 * (2 * kPMaxSeg + 1);

struct PackArgs {
  const float2* iq;
  unsigned long long n_raw;
  const rfid_b200_segment* segs;
  int nseg;
  int max_windows;
  rfid_b200_window_result* results;
  int32_t* counts;
  float2* window_tap;
  float2* win_scratch;
  int win_stride, rn16_pad;
  int dstage_samples;
  RxConfig cfg;
  int G;                                     // segments per CTA of this launch
  int raw_stage_samples;
  int seg_bytes;                             // per-segment shared-memory region
  int o_raw, o_ring_y, o_ring_a, o_snap, o_dstage;  // offsets inside a segment region
  int off_dA, off_dD, off_seg;               // offsets from the dynamic shared-memory base
  int smem_bytes;
};

struct PackSegCtl {
  uint64_t raw_full[kPRawStages];
  uint64_t win_ready[2], win_free[2];
  uint64_t bell[2];          // doorbell of the streaming decode: one arrival per progress publication
  int meta_kind[2], meta_open[2], meta_ordinal[2], meta_len[2];
  int progress[2];
  int aborted[2];
  int n_e[kPS];
  int n_ev[kPS];
  TileEvent ev[kPS][kMaxTileEvents];
};

#ifdef RFID_B200_PHASE_PROFILE
// developer aid (tools/pack_profile.py): per-phase clock64() sums of every tile warp / chain warp, written to the window tap
#define PP_DECL long long pp_t0 = clock64(), pp_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long* pp_step_log = nullptr; int pp_step = 0;
#define PP_MARK(i) { const long long pp_t1 = clock64(); pp_acc[i] += pp_t1 - pp_t0; if (pp_step_log && lane == 0) pp_step_log[pp_step * 8 + (i)] = pp_t1 - pp_t0; pp_t0 = pp_t1; }
#define PP_SUB(i) { const long long pp_t2 = clock64(); if (pp_step_log && lane == 0) pp_step_log[128 * 8 + pp_step * 8 + (i)] = pp_t2 - pp_t0; }
#define PP_DUMP(row) if (lane == 0 && A.window_tap) { long long* o = reinterpret_cast<long long*>(A.window_tap) + (size_t)(row) * 8; for (int q_ = 0; q_ < 8; q_++) o[q_] = pp_acc[q_]; }
#else
#define PP_DECL
#define PP_MARK(i)
#define PP_SUB(i)
#define PP_DUMP(row)
#endif

enum : int { PBAR_X = 1, PBAR_Y = 3 };
template <int BASE>
__device__ __forceinline__ void pbar_sync(int parity, int count)
{
  if (parity == 0) asm volatile("bar.sync %0, %1;" ::"n"((int)BASE), "r"(count) : "memory");
  else asm volatile("bar.sync %0, %1;" ::"n"((int)(BASE + 1)), "r"(count) : "memory");
}
template <int BASE>
__device__ __forceinline__ void pbar_arrive(int parity, int count)
{
  __threadfence_block();
  if (parity == 0) asm volatile("bar.arrive %0, %1;" ::"n"((int)BASE), "r"(count) : "memory");
  else asm volatile("bar.arrive %0, %1;" ::"n"((int)(BASE + 1)), "r"(count) : "memory");
}

template <int DECIM, int MFQ>
__global__ void __launch_bounds__(kPackMaxThreads, 1) rx_pack_kernel(const PackArgs A)
{
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ PackSegCtl ctl_all[kPMaxSeg];

  static_assert(MFQ >= 2, "pack kernel: block-sum matched filter");
  constexpr int Q = kTT / 32;
  static_assert(MFQ - 1 <= Q, "block-sum halo comes from the neighbouring lane only");
  static_assert(Q == 4, "four outputs per lane");

  const int G = A.G;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int seg0 = blockIdx.x * G;
  const int g_act = min(G, A.nseg - seg0);   // segments this CTA really has
  const RxConfig& C = A.cfg;
  const int bar_count = 32 * (G + 1);

  float* const dA = reinterpret_cast<float*>(smem + A.off_dA);   // [2][G][kPChainBuf]
  float* const dD = reinterpret_cast<float*>(smem + A.off_dD);   // [kPS][2][G][kPChainBuf]

  // ---- init: zero the time rings (win_samples / dc_samples start at 0, gate_impl.cc:55-56), barriers
  for (int g = 0; g < G; g++) {
    unsigned char* sb = smem + A.off_seg + (size_t)g * A.seg_bytes;
    float2* ry = reinterpret_cast<float2*>(sb + A.o_ring_y);
    float* ra = reinterpret_cast<float*>(sb + A.o_ring_a);
    for (int i = threadIdx.x; i < kPRing; i += blockDim.x) { ry[i] = make_float2(0.f, 0.f); ra[i] = 0.f; }
  }
  if (threadIdx.x < G) {
    PackSegCtl& c = ctl_all[threadIdx.x];
    for (int s = 0; s < kPRawStages; s++) mbar_init(&c.raw_full[s], 1);
    for (int s = 0; s < 2; s++) { mbar_init(&c.win_ready[s], 1); mbar_init(&c.win_free[s], 1); mbar_init(&c.bell[s], 1); }
    for (int s = 0; s < kPS; s++) { c.n_e[s] = 0; c.n_ev[s] = 0; }
    mbar_fence_init();
  }
  // lockstep length: the longest segment of the CTA
  int max_tiles = 0;
  for (int g = 0; g < g_act; g++) {
    const int n_out_g = (int)(A.segs[seg0 + g].length / DECIM);
    max_tiles = max(max_tiles, (n_out_g + kTT - 1) / kTT);
  }
  const int nsteps = max_tiles + 3;
  __syncthreads();

  if (warp < G) {
    // ======================================================================================= tile warp
    const int g = warp;
    const bool have = g < g_act;
    const int seg = seg0 + g;
    rfid_b200_segment sg;
    sg.offset = 0; sg.length = 0; sg.reserved = 0;
    if (have) sg = A.segs[seg];
    const int n_out = (int)(sg.length / DECIM);
    const int ntiles = (n_out + kTT - 1) / kTT;
    PackSegCtl& B = ctl_all[g];
    unsigned char* sb = smem + A.off_seg + (size_t)g * A.seg_bytes;
    float2* raw = reinterpret_cast<float2*>(sb + A.o_raw);
    float2* ring_y = reinterpret_cast<float2*>(sb + A.o_ring_y);
    float* ring_a = reinterpret_cast<float*>(sb + A.o_ring_a);
    float2* snap = reinterpret_cast<float2*>(sb + A.o_snap);
    float2* const win_base = A.win_scratch + (size_t)seg * A.win_stride;
    auto bufA = [&](int tile) { return dA + (size_t)((tile & 1) * G + g) * kPChainBuf; };
    auto bufD = [&](int tile, int comp) { return dD + (size_t)(((tile & (kPS - 1)) * 2 + comp) * G + g) * kPChainBuf; };

    // ---- P1 state: raw tile geometry (as rx_fused_split.cuh)
    const int odd = (int)(sg.offset & 1ull);
    const uint32_t fast_bytes = (uint32_t)((DECIM * kTT + 2 * odd) * 8);
    int fast_tiles = 0;
    {
      const long long by_len = ((long long)sg.length - 1 - (long long)DECIM * (kTT - 1)) / ((long long)DECIM * kTT);
      const long long room = (long long)A.n_raw - (long long)sg.offset + (DECIM - 1) + odd - (DECIM * kTT + 2 * odd);
      const long long by_buf = room >= 0 ? room / ((long long)DECIM * kTT) : -1;
      long long f = (by_len < by_buf ? by_len : by_buf) + 1;
      if (sg.length < (unsigned)(DECIM * kTT)) f = 0;
      fast_tiles = f < 0 ? 0 : (f > ntiles ? ntiles : (int)f);
    }
    const float2* const fast_src = A.iq + sg.offset - (DECIM - 1) - odd;
    FusedArgs FA;  // issue_tile_load only reads iq / n_raw
    FA.iq = A.iq; FA.n_raw = A.n_raw;
    auto load_tile = [&](int k, int rs_) {
      float2* dst = raw + (size_t)rs_ * A.raw_stage_samples;
      if (k >= 1 && k < fast_tiles) {
        mbar_arrive_expect_tx(&B.raw_full[rs_], fast_bytes);
        tma_load_1d(dst, fast_src + (size_t)k * (DECIM * kTT), fast_bytes, &B.raw_full[rs_]);
      } else {
        issue_tile_load<DECIM>(FA, sg, k, dst, &B.raw_full[rs_]);
      }
    };
    if (lane == 0)
      for (int k = 0; k < kPRawStages && k < ntiles; k++) load_tile(k, k);
    const float winlen_f = (float)C.win_length, dclen_f = (float)C.dc_length;
    float2 b_keep[MFQ - 1];
#pragma unroll
    for (int m = 0; m < MFQ - 1; m++) b_keep[m] = make_float2(0.f, 0.f);
    int rs = 0;
    uint32_t raw_par = 0;

    // ---- P3 state: the gate (gate_impl.cc:45, global_vars.cc:47, reader_impl.cc:259,262)
    bool sig_pos = false;
    int n_samples = 0, num_pulses = 0;
    bool gate_open = false;
    int to_ungate = C.len_rn16;
    int wcount = 0, open_idx = 0;
    bool cur_store = false;
    int nq = 1;
    bool terminated = false;
    int closed_since = C.dc_length;
    const int half_pw = C.n_PW / 2;

    // ---- E state: window emission, three tiles behind P1
    bool f_open = false, f_store = false;
    int f_wpos = 0, n_signalled = 0, n_freed = 0, f_slot = 0, wsig_ordinal = 0;
    float2 dc_open = make_float2(0.f, 0.f);
    float2* win = win_base;

    PP_DECL
#ifdef RFID_B200_PHASE_PROFILE
    if (seg == 0 && A.window_tap) pp_step_log = reinterpret_cast<long long*>(A.window_tap) + (size_t)(A.nseg + gridDim.x) * 8;
#endif
    for (int i = 0; i < nsteps; i++) {
#ifdef RFID_B200_PHASE_PROFILE
      pp_step = i;
#endif
      PP_MARK(5)
      // ================================================================= P1(i): matched filter, |y|, ring differences
      if (i < ntiles) {
        const int k = i, ts = k & (kPS - 1);
        const float2* stage = raw + (size_t)rs * A.raw_stage_samples;
        const int delta = -odd - (k > 0 ? DECIM - 1 : 0);
        const int nvalid = min(kTT, n_out - k * kTT);
        mbar_wait(&B.raw_full[rs], raw_par);
        PP_MARK(0)
        const int t0 = Q * lane;
        const int base = DECIM * t0 - (DECIM - 1) - delta;
        float2 w[MFQ - 1 + Q];
#pragma unroll
        for (int h = 0; h < Q; h += 2) {
          float2 x[2 * DECIM];
          if (odd == 0 && k > 0) {
            const float4* p4 = reinterpret_cast<const float4*>(stage + base + DECIM * h);
#pragma unroll
            for (int j = 0; j < DECIM; j++) {
              const float4 v = p4[j];
              x[2 * j] = make_float2(v.x, v.y);
              x[2 * j + 1] = make_float2(v.z, v.w);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 2 * DECIM; j++) {
              const bool before = (k == 0) && (DECIM * (t0 + h) - (DECIM - 1) + j < 0);  // before sample 0 of the segment: +0
              x[j] = before ? make_float2(0.f, 0.f) : stage[base + DECIM * h + j];
            }
          }
#pragma unroll
          for (int q = 0; q < 2; q++) {
            float2 b = x[DECIM * q];
#pragma unroll
            for (int j = 1; j < DECIM; j++) b = c_add2(b, x[DECIM * q + j]);
            w[MFQ - 1 + h + q] = b;
          }
        }
        PP_SUB(0)
        __syncwarp();  // raw stage consumed
        if (lane == 0 && k + kPRawStages < ntiles) load_tile(k + kPRawStages, rs);
        __syncwarp();
        PP_SUB(1)
        if (++rs == kPRawStages) { rs = 0; raw_par ^= 1u; }
#pragma unroll
        for (int m = 0; m < MFQ - 1; m++) {
          const float2 mine = w[Q + m];
          const float ux = __shfl_up_sync(0xffffffffu, mine.x, 1), uy = __shfl_up_sync(0xffffffffu, mine.y, 1);
          w[m] = lane ? make_float2(ux, uy) : b_keep[m];
          b_keep[m] = make_float2(__shfl_sync(0xffffffffu, mine.x, 31), __shfl_sync(0xffffffffu, mine.y, 31));
        }
        PP_SUB(2)
        float2 y[Q];
        float a[Q];
        bool risky = false;
#pragma unroll
        for (int q = 0; q < Q; q++) {
          y[q] = w[q];
#pragma unroll
          for (int m = 1; m < MFQ; m++) y[q] = c_add2(y[q], w[q + m]);
          bool rq;
          a[q] = cabsf_quick(y[q].x, y[q].y, rq);  // gate_impl.cc:130
          risky = risky || rq;
        }
        if (__any_sync(0xffffffffu, risky)) {  // rare (about one tile in 60): the exact evaluation for the whole warp
#pragma unroll
          for (int q = 0; q < Q; q++) a[q] = cabsf_ref(y[q].x, y[q].y);
        }
        {
          float4* py = reinterpret_cast<float4*>(ring_y + ts * kTT + t0);
          py[0] = make_float4(y[0].x, y[0].y, y[1].x, y[1].y);
          py[1] = make_float4(y[2].x, y[2].y, y[3].x, y[3].y);
          *reinterpret_cast<float4*>(ring_a + ts * kTT + t0) = make_float4(a[0], a[1], a[2], a[3]);
        }
        PP_SUB(3)
        __syncwarp();  // this tile's |y| and y visible to the lookbacks below
        float xd[Q], xr[Q], xi[Q];
        int ia = ts * kTT + t0 - C.win_length, iy = ts * kTT + t0 - C.dc_length;
        if (ia < 0) ia += kPRing;
        if (iy < 0) iy += kPRing;
        if (((C.win_length | C.dc_length) & 3) == 0) {  // lookback groups are aligned and never straddle the ring's end
          const float4 oa = *reinterpret_cast<const float4*>(ring_a + ia);
          xd[0] = f_sub(a[0], oa.x); xd[1] = f_sub(a[1], oa.y); xd[2] = f_sub(a[2], oa.z); xd[3] = f_sub(a[3], oa.w);
#pragma unroll
          for (int q = 0; q < Q; q += 2) {
            const float4 oy = *reinterpret_cast<const float4*>(ring_y + iy + q);
            xr[q] = f_sub(y[q].x, oy.x); xi[q] = f_sub(y[q].y, oy.y);
            xr[q + 1] = f_sub(y[q + 1].x, oy.z); xi[q + 1] = f_sub(y[q + 1].y, oy.w);
          }
        } else {
#pragma unroll
          for (int q = 0; q < Q; q++) {
            int ja = ia + q, jy = iy + q;
            if (ja >= kPRing) ja -= kPRing;
            if (jy >= kPRing) jy -= kPRing;
            const float2 old = ring_y[jy];
            xd[q] = f_sub(a[q], ring_a[ja]);
            xr[q] = f_sub(y[q].x, old.x);
            xi[q] = f_sub(y[q].y, old.y);
          }
        }
        float mx = fabsf(xd[0]), mn = mx;
#pragma unroll
        for (int q = 0; q < Q; q++) {
          mx = fmaxf(fmaxf(mx, fabsf(xd[q])), fmaxf(fabsf(xr[q]), fabsf(xi[q])));
          mn = fminf(fminf(mn, fabsf(xd[q])), fminf(fabsf(xr[q]), fabsf(xi[q])));
        }
        PP_SUB(4)
        const bool all_ok = C.win_div_fast && C.dc_div_fast && mn >= kDivFastMin && mx <= kDivFastMax;
        float qd[Q], qr[Q], qi[Q];
        if (__all_sync(0xffffffffu, all_ok)) {
#pragma unroll
          for (int q = 0; q < Q; q++) {
            qd[q] = f_div_fast(xd[q], winlen_f, C.win_recip);
            qr[q] = f_div_fast(xr[q], dclen_f, C.dc_recip);
            qi[q] = f_div_fast(xi[q], dclen_f, C.dc_recip);
          }
        } else {  // an exact zero, a denormal, or an unverified divisor somewhere in the warp: IEEE division
#pragma unroll
          for (int q = 0; q < Q; q++) {
            qd[q] = f_div_const(xd[q], winlen_f, C.win_recip, C.win_div_fast);
            qr[q] = f_div_const(xr[q], dclen_f, C.dc_recip, C.dc_div_fast);
            qi[q] = f_div_const(xi[q], dclen_f, C.dc_recip, C.dc_div_fast);
          }
        }
        if (nvalid < kTT) {
          // the segment's last, partial tile: the chain warp runs whole groups of 16 steps, so the slots past the end
          // hold -0.0f (x + -0.0f == x for every x, including both zeros: the running sum is carried unchanged)
#pragma unroll
          for (int q = 0; q < Q; q++)
            if (t0 + q >= nvalid) { qd[q] = -0.0f; qr[q] = -0.0f; qi[q] = -0.0f; }
        }
        *reinterpret_cast<float4*>(bufA(k) + t0) = make_float4(qd[0], qd[1], qd[2], qd[3]);
        *reinterpret_cast<float4*>(bufD(k, 0) + t0) = make_float4(qr[0], qr[1], qr[2], qr[3]);
        *reinterpret_cast<float4*>(bufD(k, 1) + t0) = make_float4(qi[0], qi[1], qi[2], qi[3]);
        __syncwarp();
      }
      PP_MARK(1)
      pbar_arrive<PBAR_X>(i & 1, bar_count);                 // tile i is ready for the chain warp
      if (i >= 1) pbar_sync<PBAR_Y>((i - 1) & 1, bar_count);  // avg_ampl of tile i-1 and dc_est of tile i-3 are final
      PP_MARK(2)

      // ================================================================= E(i-3): window emission (gate_impl.cc:173,187)
      if (i >= 3 && i - 3 < ntiles && (f_open || B.n_ev[(i - 3) & (kPS - 1)] != 0)) {  // nothing to emit on most tiles
        const int t = i - 3, ps = t & (kPS - 1);
        const float2* py = ring_y + ps * kTT;
        const float* pe_re = bufD(t, 0);
        const float* pe_im = bufD(t, 1);
        const int pvalid = min(kTT, n_out - t * kTT);
        const int pnev = B.n_ev[ps];
        int pos = 0;
        for (int e = 0; e <= pnev; e++) {
          const bool last = e == pnev;
          const int etype = last ? 0 : B.ev[ps][e].type;
          const int epos = last ? pvalid : B.ev[ps][e].pos;
          if (f_open) {
            const int take = epos - pos;
            if (f_store && take > 0) {
              for (int j = lane; j < take; j += 32) win[f_wpos + j] = c_sub(py[pos + j], dc_open);
              __threadfence_block();  // samples first, then the counter the decoder (same CTA) polls
              __syncwarp();
              if (lane == 0) { *(volatile int*)&B.progress[f_slot] = f_wpos + take; mbar_arrive(&B.bell[f_slot]); }
            }
            f_wpos += take;
            pos = epos;
          }
          if (etype == 2) {
            f_open = false;  // the decoder already has the window: it saw progress reach its length
            pos = epos;
          } else if (etype == 1) {
            const int j = B.ev[ps][e].a;
            dc_open = make_float2(pe_re[j], pe_im[j]);  // dc_est right after the trigger sample
            f_store = B.ev[ps][e].c != 0;
            f_open = true;
            win = win_base + (B.ev[ps][e].d ? A.rn16_pad : 0);
            if (f_store) {
              // the scratch area and the meta slot are reused two hand-offs later
              while (n_signalled - n_freed >= 2) { mbar_wait(&B.win_free[n_freed & 1], (n_freed >> 1) & 1); n_freed++; }
              f_slot = n_signalled & 1;
              if (lane == 0) {
                win[0] = c_sub(py[epos], dc_open);
                const int knd = B.ev[ps][e].d;
                B.meta_kind[f_slot] = knd; B.meta_ordinal[f_slot] = wsig_ordinal; B.meta_open[f_slot] = B.ev[ps][e].b;
                B.meta_len[f_slot] = knd ? C.len_epc : C.len_rn16;
                *(volatile int*)&B.progress[f_slot] = 0;
                B.aborted[f_slot] = 0;
              }
              __threadfence_block();
              __syncwarp();
              if (lane == 0) {
                *(volatile int*)&B.progress[f_slot] = 1;
                __threadfence_block();
                mbar_arrive(&B.win_ready[f_slot]);  // hand the window to the decoder NOW (streaming decode)
              }
              n_signalled++;
            }
            wsig_ordinal++;
            f_wpos = 1;
            pos = epos + 1;
          }
        }
        __syncwarp();
      }

      PP_MARK(3)
      // ================================================================= P3(i-1): thresholds, state machine, DC list
      if (i >= 1 && i - 1 < ntiles) {
        const int t = i - 1, s = t & (kPS - 1);
        int nev = 0, n_e = 0;
        const int nvalid = min(kTT, n_out - t * kTT);
        const float* davg = bufA(t);
        const float* ta = ring_a + s * kTT;
        const float2* ty = ring_y + s * kTT;
        float* er = bufD(t, 0);
        float* ei = bufD(t, 1);
        bool list_rebuilt = false;
        if (!terminated && gate_open && to_ungate - n_samples > nvalid) {
          // the whole tile lies inside an open window (gate_impl.cc:182-195): nothing to detect, no DC update
          n_samples += nvalid;
          list_rebuilt = true;
        } else if (!terminated) {
          // thresholds (gate_impl.cc:136,148,154).  First a one-vote test in the lanes' natural 4-sample groups: while the
          // signal is high and no sample of the tile falls below its threshold, no edge can occur (carrier only).
          unsigned lt[4] = {0u, 0u, 0u, 0u}, gt[4] = {0u, 0u, 0u, 0u};
          bool quiet = false;
          if (sig_pos && !gate_open && nvalid == kTT) {
            const float4 av = *reinterpret_cast<const float4*>(davg + 4 * lane);
            const float4 aa = *reinterpret_cast<const float4*>(ta + 4 * lane);
            const bool below = aa.x < f_mul(av.x, kThreshFraction) || aa.y < f_mul(av.y, kThreshFraction) ||
                               aa.z < f_mul(av.z, kThreshFraction) || aa.w < f_mul(av.w, kThreshFraction);
            quiet = !__any_sync(0xffffffffu, below);
          }
          bool have_masks = false;
          auto make_masks = [&]() {
#pragma unroll
            for (int r = 0; r < 4; r++) {
              const int p = r * 32 + lane;
              const float thr = f_mul(davg[p], kThreshFraction);
              const float a = ta[p];
              lt[r] = __ballot_sync(0xffffffffu, a < thr);
              gt[r] = __ballot_sync(0xffffffffu, a > thr);
            }
            if (nvalid < kTT) {
#pragma unroll
              for (int r = 0; r < 4; r++) {
                const int left = nvalid - r * 32;
                const unsigned vm = left >= 32 ? 0xffffffffu : (left > 0 ? (1u << left) - 1u : 0u);
                lt[r] &= vm;
                gt[r] &= vm;
              }
            }
            have_masks = true;
          };
          int pos = 0;
          while (pos < nvalid) {
            if (!gate_open) {
              const int run_start = pos;
              int p_open = -1;
              // (the one-vote result only covers a run that starts the tile with the signal high)
              if (!have_masks && !(quiet && run_start == 0)) make_masks();
              if (sig_pos && (lt[0] | lt[1] | lt[2] | lt[3]) == 0u) {
                // carrier only (the common case): no falling edge can occur, only the open test remains
                if (num_pulses > kNumPulsesCommand) {
                  const int cand = run_start + max(0, C.n_T1 - n_samples);
                  if (cand < nvalid) { p_open = cand; num_pulses = 0; n_samples = 1; }
                }
                if (p_open < 0) n_samples += nvalid - run_start;
              } else {
                GateFsm fs = {sig_pos, n_samples, num_pulses};
                const Mask128 ltm = {{lt[0], lt[1], lt[2], lt[3]}}, gtm = {{gt[0], gt[1], gt[2], gt[3]}};
                p_open = fsm_closed_run(ltm, gtm, run_start, nvalid, C.n_T1, half_pw, fs);
                sig_pos = fs.sig_pos; n_samples = fs.n_samples; num_pulses = fs.num_pulses;
              }
              const bool opened = p_open >= 0;
              pos = opened ? p_open + 1 : nvalid;
              // ---- DC tracker inputs of the closed run [run_start, pos) (gate_impl.cc:141-143; includes the trigger)
              const int len = pos - run_start;
              if (run_start == 0 && pos == nvalid && !opened && closed_since >= C.dc_length) {
                // no gate activity and the ring lookback is time-contiguous: P1's differences are exact
              } else {
                list_rebuilt = true;
#pragma unroll 1
                for (int j = lane; j < len; j += 32) {
                  const int p = run_start + j, m = closed_since + j;
                  const float2 yv = ty[p];
                  float2 old;
                  if (m < C.dc_length) {
                    old = snap[m];  // ring contents from before the window
                  } else {
                    int iy = s * kTT + p - C.dc_length;
                    if (iy < 0) iy += kPRing;
                    old = ring_y[iy];
                  }
                  er[n_e + j] = f_div_const(f_sub(yv.x, old.x), dclen_f, C.dc_recip, C.dc_div_fast);
                  ei[n_e + j] = f_div_const(f_sub(yv.y, old.y), dclen_f, C.dc_recip, C.dc_div_fast);
                }
              }
              closed_since = min(closed_since + len, 1 << 24);
              n_e += len;
              if (opened) {
                // READER COMMAND DETECTED (gate_impl.cc:164-180): keep the dc ring as it stands now
#pragma unroll 1
                for (int j = lane; j < C.dc_length; j += 32) {
                  int iy = s * kTT + (pos - 1) - C.dc_length + 1 + j;
                  if (iy < 0) iy += kPRing;
                  snap[j] = ring_y[iy];
                }
                gate_open = true;
                open_idx = t * kTT + pos - 1;
                cur_store = wcount < A.max_windows;
                if (lane == 0 && nev < kMaxTileEvents) {
                  TileEvent& ev = B.ev[s][nev];
                  ev.type = 1; ev.pos = pos - 1; ev.a = n_e - 1; ev.b = open_idx; ev.c = cur_store ? 1 : 0; ev.d = wcount & 1;
                }
                nev++;
              }
            } else {
              // ---- open: samples pass through (gate_impl.cc:182-195); emitted two steps later
              list_rebuilt = true;
              const int take = min(to_ungate - n_samples, nvalid - pos);
              n_samples += take; pos += take;
              if (n_samples >= to_ungate) {
                gate_open = false;
                const int kind = wcount & 1;  // windows alternate RN16, EPC
                if (lane == 0 && nev < kMaxTileEvents) {
                  TileEvent& ev = B.ev[s][nev];
                  ev.type = 2; ev.pos = pos; ev.a = kind; ev.b = wcount; ev.c = to_ungate; ev.d = open_idx;
                }
                nev++;
                wcount++;
                closed_since = 0;
                // ACK after RN16 -> GATE_SEEK_EPC, Query/QueryRep after EPC -> GATE_SEEK_RN16 (gate_impl.cc:112-123)
                to_ungate = kind ? C.len_rn16 : C.len_epc;
                n_samples = 0;
                if (kind) {
                  nq++;
                  if (nq > C.max_queries) { terminated = true; break; }  // gate_impl.cc:101-109
                }
              }
            }
          }
        } else {
          list_rebuilt = true;
        }
        if (list_rebuilt || terminated) {
          // the closed-sample list is shorter than the tile: pad its last group of 16 with -0.0f (see P1)
          const int n16 = (n_e + 15) & ~15;
          if (lane < 16 && n_e + lane < n16) { er[n_e + lane] = -0.0f; ei[n_e + lane] = -0.0f; }
        }
        if (lane == 0) { B.n_e[s] = n_e; B.n_ev[s] = min(nev, kMaxTileEvents); }
        __syncwarp();
      }
      PP_MARK(4)
    }
    if (have) { PP_DUMP(seg) }
    // ---- end of the segment
    if (have) {
      if (f_open && f_store && lane == 0) {
        // the segment ended inside a window the decoder is already working on: let it run to the end
        *(volatile int*)&B.aborted[f_slot] = 1;
        __threadfence_block();
        *(volatile int*)&B.progress[f_slot] = 1 << 30;
        mbar_arrive(&B.bell[f_slot]);
      }
      __syncwarp();
      if (lane == 0) A.counts[seg] = wcount;
    }
    while (n_signalled - n_freed >= 2) { mbar_wait(&B.win_free[n_freed & 1], (n_freed >> 1) & 1); n_freed++; }
    if (lane == 0) {
      B.meta_kind[n_signalled & 1] = -1;
      __threadfence_block();
      mbar_arrive(&B.win_ready[n_signalled & 1]);
    }
  } else if (warp == G) {
    // ======================================================================================= chain warp
    // lane 8*c + g: running sum c (0 avg_ampl, 1 dc_est.re, 2 dc_est.im) of segment g
    const int comp = lane >> 3, g = lane & 7;
    const bool active = comp < 3 && g < g_act;
    int n_out = 0;
    if (active) n_out = (int)(A.segs[seg0 + g].length / DECIM);
    float acc = 0.f;
    PP_DECL
#ifdef RFID_B200_PHASE_PROFILE
    if (blockIdx.x == 0 && A.window_tap) pp_step_log = reinterpret_cast<long long*>(A.window_tap) + (size_t)(A.nseg + gridDim.x) * 8 + 64 * 8;
#endif
    for (int i = 0; i < nsteps; i++) {
#ifdef RFID_B200_PHASE_PROFILE
      pp_step = i;
#endif
      pbar_sync<PBAR_X>(i & 1, bar_count);
      PP_MARK(0)
      {
        // branch-free selection of this lane's buffer and length, then ONE convergent loop for all 24 chains
        const int t = i - 2;
        const int n_avg = min(kTT, max(0, n_out - i * kTT));
        const int n_dc = (active && comp > 0 && i >= 2) ? ctl_all[g].n_e[t & (kPS - 1)] : 0;
        const int n = active ? (comp == 0 ? n_avg : n_dc) : 0;
        const int ofsA = ((i & 1) * G + g) * kPChainBuf;
        const int ofsD = (((t & (kPS - 1)) * 2 + (comp - 1)) * G + g) * kPChainBuf;
        float* buf = comp == 0 ? dA + ofsA : dD + (active && comp > 0 ? ofsD : 0);
        const int n16 = (n + 15) & ~15;
        __syncwarp();
        chain_inplace(buf, n16, acc);
      }
      __syncwarp();
      PP_MARK(1)
      if (i + 1 < nsteps) pbar_arrive<PBAR_Y>(i & 1, bar_count);
      PP_MARK(2)
    }
    PP_DUMP(A.nseg + blockIdx.x)
  } else {
    // ======================================================================================= decoder warp
    const int g = warp - G - 1;
    if (g < g_act) {
      const int seg = seg0 + g;
      PackSegCtl& B = ctl_all[g];
      unsigned char* sb = smem + A.off_seg + (size_t)g * A.seg_bytes;
      float2* dstage = reinterpret_cast<float2*>(sb + A.o_dstage);
      float2* const win_base = A.win_scratch + (size_t)seg * A.win_stride;
      for (int j = 0;; j++) {
        mbar_wait_relaxed(&B.win_ready[j & 1], (j >> 1) & 1, 20000);  // parked by the hardware until a window arrives
        const int kind = B.meta_kind[j & 1];
        if (kind < 0) break;
        const int ordinal = B.meta_ordinal[j & 1], open_idx = B.meta_open[j & 1], len = B.meta_len[j & 1];
        const float2* win = win_base + (kind ? A.rn16_pad : 0);
        WindowDecode wd;
        decode_window_staged(C, kind, win, len, dstage, A.dstage_samples, wd, (const volatile int*)&B.progress[j & 1], &B.bell[j & 1]);
        rfid_b200_window_result* dst = A.results + (size_t)seg * A.max_windows + ordinal;
        const bool aborted = *(volatile int*)&B.aborted[j & 1] != 0;
        if (lane == 0 && !aborted) store_result(dst, wd, seg, ordinal, open_idx, len, kind);
#ifndef RFID_B200_PHASE_PROFILE
        if (A.window_tap) {
          float2* tap = A.window_tap + ((size_t)seg * A.max_windows + ordinal) * C.len_epc;
          for (int p = lane; p < len; p += 32) tap[p] = __ldcg(win + p);
        }
#endif
        __syncwarp();
        if (lane == 0) mbar_arrive(&B.win_free[j & 1]);
      }
    }
  }
}

}  // namespace rfid_b200
