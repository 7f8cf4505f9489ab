// rx_fused.cuh -- the fused capture-mode kernel: matched filter + decimation,
// gate (reader-command detection, windowing) and tag_decoder for one capture
// segment per CTA, streaming the segment through shared memory in tiles so the
// only HBM traffic is one read of every raw sample (8 B/sample) plus 64-byte
// result records.
//
// Replaces (reference gr-rfid/): filter.fir_filter_ccc(5,[1]*25) (apps/reader.py:65,75),
// gate_impl::general_work (lib/gate_impl.cc:85-200) and
// tag_decoder_impl::general_work (lib/tag_decoder_impl.cc:196-397).
//
// CTA = 4 warps with fixed roles, connected by mbarrier pipelines:
//   warp 0      sequencer: everything that is order-dependent in the reference -- the two float
//               running means (avg_ampl: one dependent FADD per sample on lane 0; dc_est: lanes 1,2),
//               the edge/pulse state machine (bit-mask hopping, a few steps per command), window
//               bookkeeping.  Lane-parallel inside a tile for everything else (thresholds, DC ring
//               differences, window emission).
//   warps 1..2  workers: wait for the TMA bulk copy of the next raw tile, block-sum matched filter,
//               |y| (exact cabsf), amplitude-ring difference /win_length  -> tile stage.
//   warp 3      decoder: when the sequencer closes a window, decodes it from shared memory
//               (preamble correlation, channel estimate, FM0 decisions, period search, CRC-16)
//               and writes the result record.
#pragma once

#include "rx_common.cuh"
#include "rx_decode.cuh"

namespace rfid_b200 {

constexpr int kTT = 128;          // decimated samples per tile
constexpr int kRawStages = 3;
constexpr int kTileStages = 3;
constexpr int kWorkerWarps = 2;
constexpr int kWorkerThreads = kWorkerWarps * 32;
constexpr int kFusedThreads = 32 * (2 + kWorkerWarps);

struct FusedArgs {
  const float2* iq;              // raw capture (device), 16-byte aligned
  unsigned long long n_raw;      // total samples in the capture buffer
  const rfid_b200_segment* segs;
  int nseg;
  int max_windows;               // record slots per segment
  rfid_b200_window_result* results;
  int32_t* counts;
  float2* window_tap;            // optional: ungated samples of every stored window (stride len_epc)
  RxConfig cfg;
  // shared-memory carve-up (bytes from the dynamic smem base), computed on the host
  int off_raw, raw_stage_samples;
  int off_bhist, bhist_size;     // float2[bhist_size] (+ partial-block ring right after it when mf_rem > 0)
  int off_ahist, ahist_size;     // float[ahist_size]
  int off_tile_y, off_tile_a, off_tile_d;
  int off_ycl, ycl_size;         // float2[ycl_size]
  int off_e;                     // float[2][kTT]
  int off_win, off_M;
  int smem_bytes;
};

struct FusedBars {
  uint64_t raw_full[kRawStages];
  uint64_t tile_full[kTileStages];
  uint64_t tile_empty[kTileStages];
  uint64_t win_ready, win_free;
  int meta_kind, meta_open, meta_ordinal, meta_len;
};

__device__ __forceinline__ int next_set128(unsigned long long lo, unsigned long long hi, int pos)
{
  if (pos < 64) {
    unsigned long long m = lo & (~0ull << pos);
    if (m) return __ffsll((long long)m) - 1;
    pos = 64;
  }
  if (pos < 128) {
    unsigned long long m = hi & (~0ull << (pos - 64));
    if (m) return 64 + __ffsll((long long)m) - 1;
  }
  return 128;
}

// sequential in-place running sum: acc = acc + buf[i]; buf[i] = acc   (one rounding per step)
__device__ __forceinline__ void chain_inplace(float* buf, int n, float& acc)
{
  int i = 0;
  for (; i + 4 <= n; i += 4) {
    float4 v = *reinterpret_cast<float4*>(buf + i);
    acc = f_add(acc, v.x); v.x = acc;
    acc = f_add(acc, v.y); v.y = acc;
    acc = f_add(acc, v.z); v.z = acc;
    acc = f_add(acc, v.w); v.w = acc;
    *reinterpret_cast<float4*>(buf + i) = v;
  }
  for (; i < n; i++) {
    acc = f_add(acc, buf[i]);
    buf[i] = acc;
  }
}

// raw tile geometry: segment-relative index of the first sample held in the stage buffer
template <int DECIM>
__device__ __forceinline__ long long tile_load_start(unsigned long long seg_off, int k)
{
  long long lo = (long long)DECIM * k * kTT - (DECIM - 1);
  if (lo < 0) lo = 0;
  long long abs_lo = (long long)seg_off + lo;
  abs_lo &= ~1ll;  // 16-byte aligned source
  return abs_lo - (long long)seg_off;
}

template <int DECIM>
__device__ __forceinline__ void issue_tile_load(const FusedArgs& A, const rfid_b200_segment& sg, int k, float2* stage,
                                                uint64_t* bar)
{
  const long long start = tile_load_start<DECIM>(sg.offset, k);  // may be -1
  long long hi = (long long)DECIM * ((long long)k * kTT + kTT - 1);
  if (hi > (long long)sg.length - 1) hi = (long long)sg.length - 1;
  long long count = hi - start + 1;  // samples start..hi
  long long count_al = (count + 1) & ~1ll;
  const long long abs_start = (long long)sg.offset + start;
  float2 tail = make_float2(0.f, 0.f);
  bool patch = false;
  if ((unsigned long long)(abs_start + count_al) > A.n_raw) {
    // the rounded-up copy would run one sample past the capture buffer: copy an even count and
    // fetch the last sample with a plain load
    count_al -= 2;
    tail = A.iq[abs_start + count - 1];
    patch = true;
  }
  const uint32_t bytes = (uint32_t)(count_al * 8);
  if (patch) stage[count - 1] = tail;
  mbar_arrive_expect_tx(bar, bytes);
  if (bytes) tma_load_1d(stage, A.iq + abs_start, bytes, bar);
}

template <int DECIM>
__global__ void __launch_bounds__(kFusedThreads) rx_fused_kernel(const FusedArgs A)
{
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ FusedBars B;

  const int seg = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const RxConfig& C = A.cfg;
  const rfid_b200_segment sg = A.segs[seg];
  const int n_out = (int)(sg.length / DECIM);
  const int ntiles = (n_out + kTT - 1) / kTT;

  float2* raw = reinterpret_cast<float2*>(smem + A.off_raw);
  float2* bhist = reinterpret_cast<float2*>(smem + A.off_bhist);
  float2* phist = bhist + A.bhist_size;
  float* ahist = reinterpret_cast<float*>(smem + A.off_ahist);
  float2* tile_y = reinterpret_cast<float2*>(smem + A.off_tile_y);
  float* tile_a = reinterpret_cast<float*>(smem + A.off_tile_a);
  float* tile_d = reinterpret_cast<float*>(smem + A.off_tile_d);
  float2* ycl = reinterpret_cast<float2*>(smem + A.off_ycl);
  float* e_re = reinterpret_cast<float*>(smem + A.off_e);
  float* e_im = e_re + kTT;
  float2* win = reinterpret_cast<float2*>(smem + A.off_win);
  float* Msq = reinterpret_cast<float*>(smem + A.off_M);

  // ---- init: zero the history rings (win_samples / dc_samples start at 0, gate_impl.cc:55-56;
  //      x[<0] = +0 for the matched filter), set up the barriers
  for (int i = threadIdx.x; i < A.bhist_size * (C.mf_rem ? 2 : 1); i += kFusedThreads) bhist[i] = make_float2(0.f, 0.f);
  for (int i = threadIdx.x; i < A.ahist_size; i += kFusedThreads) ahist[i] = 0.f;
  for (int i = threadIdx.x; i < A.ycl_size; i += kFusedThreads) ycl[i] = make_float2(0.f, 0.f);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kRawStages; s++) mbar_init(&B.raw_full[s], 1);
    for (int s = 0; s < kTileStages; s++) {
      mbar_init(&B.tile_full[s], kWorkerWarps);
      mbar_init(&B.tile_empty[s], 1);
    }
    mbar_init(&B.win_ready, 1);
    mbar_init(&B.win_free, 1);
    mbar_fence_init();
  }
  __syncthreads();

  if (warp >= 1 && warp <= kWorkerWarps) {
    // =========================================================== workers
    const int wt = threadIdx.x - 32;
    if (wt == 0) {
      for (int k = 0; k < kRawStages && k < ntiles; k++)
        issue_tile_load<DECIM>(A, sg, k, raw + (size_t)k * A.raw_stage_samples, &B.raw_full[k]);
    }
    const int bmask = A.bhist_size - 1, amask = A.ahist_size - 1;
    const float winlen_f = (float)C.win_length;
    for (int k = 0; k < ntiles; k++) {
      const int rs = k % kRawStages, ts = k % kTileStages;
      const float2* stage = raw + (size_t)rs * A.raw_stage_samples;
      const long long start = tile_load_start<DECIM>(sg.offset, k);
      mbar_wait(&B.raw_full[rs], (k / kRawStages) & 1);
      // ---- block sums B(n) = x[D*n-D+1 .. D*n], ascending (and the partial block when ntaps % D != 0)
#pragma unroll
      for (int r = 0; r < kTT / kWorkerThreads; r++) {
        const int t = wt + r * kWorkerThreads;
        const int n = k * kTT + t;
        if (n < n_out) {
          const long long base = (long long)DECIM * n - (DECIM - 1) - start;  // stage index of x[D*n-D+1]
          float2 x[DECIM];
#pragma unroll
          for (int j = 0; j < DECIM; j++) {
            long long si = base + j;
            // only the very first block of a segment reaches before sample 0 (reads as +0)
            x[j] = ((long long)DECIM * n - (DECIM - 1) + j >= 0) ? stage[si] : make_float2(0.f, 0.f);
          }
          float2 b = x[0];
#pragma unroll
          for (int j = 1; j < DECIM; j++) b = c_add(b, x[j]);
          bhist[n & bmask] = b;
          if (C.mf_rem) {  // P(n): newest mf_rem samples of the block, ascending (static indexing only)
            float2 p = make_float2(0.f, 0.f);
            bool started = false;
#pragma unroll
            for (int j = 0; j < DECIM; j++) {
              if (j >= DECIM - C.mf_rem) {
                p = started ? c_add(p, x[j]) : x[j];
                started = true;
              }
            }
            phist[n & bmask] = p;
          }
        }
      }
      named_bar_sync(1, kWorkerThreads);  // raw stage rs fully consumed, block sums visible
      if (wt == 0 && k + kRawStages < ntiles)
        issue_tile_load<DECIM>(A, sg, k + kRawStages, raw + (size_t)rs * A.raw_stage_samples, &B.raw_full[rs]);
      mbar_wait(&B.tile_empty[ts], ((k / kTileStages) & 1) ^ 1);
      // ---- y[n] = ((P(n-q) + B(n-q+1)) + ...) + B(n);  a = |y|
      float a_reg[kTT / kWorkerThreads];
#pragma unroll
      for (int r = 0; r < kTT / kWorkerThreads; r++) {
        const int t = wt + r * kWorkerThreads;
        const int n = k * kTT + t;
        a_reg[r] = 0.f;
        if (n < n_out) {
          float2 y;
          int m = n - C.mf_q + 1;
          if (C.mf_rem) {
            y = phist[(n - C.mf_q) & bmask];
          } else {
            y = bhist[m & bmask];
            m++;
          }
          for (; m <= n; m++) y = c_add(y, bhist[m & bmask]);
          const float a = cabsf_ref(y.x, y.y);  // gate_impl.cc:130
          tile_y[ts * kTT + t] = y;
          tile_a[ts * kTT + t] = a;
          ahist[n & amask] = a;
          a_reg[r] = a;
        }
      }
      named_bar_sync(1, kWorkerThreads);  // amplitude ring visible
      // ---- (a - win_samples[win_index]) / win_length   (gate_impl.cc:131)
#pragma unroll
      for (int r = 0; r < kTT / kWorkerThreads; r++) {
        const int t = wt + r * kWorkerThreads;
        const int n = k * kTT + t;
        if (n < n_out) tile_d[ts * kTT + t] = f_div(f_sub(a_reg[r], ahist[(n - C.win_length) & amask]), winlen_f);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&B.tile_full[ts]);
    }
  } else if (warp == 0) {
    // =========================================================== sequencer
    float acc = 0.f;  // lane 0: avg_ampl, lane 1: dc_est.re, lane 2: dc_est.im
    bool sig_pos = false;          // signal_state, starts NEG_EDGE (gate_impl.cc:45)
    int n_samples = 0, num_pulses = 0;
    bool gate_open = false;
    int to_ungate = C.len_rn16;    // first SEEK is for an RN16 (global_vars.cc:47, reader_impl.cc:262)
    int wcount = 0, wsignalled = 0;
    int n_closed = 0;              // closed-sample ordinal (index into the DC ring stream)
    int wpos = 0, open_idx = 0;
    int nq = 1;                    // n_queries_sent after START -> SEND_QUERY (reader_impl.cc:259)
    bool terminated = false, store_this = false;
    float2 dc_open = make_float2(0.f, 0.f);
    const int ymask = A.ycl_size - 1;
    const float dclen_f = (float)C.dc_length;
    const int half_pw = C.n_PW / 2;

    for (int k = 0; k < ntiles; k++) {
      const int ts = k % kTileStages;
      mbar_wait(&B.tile_full[ts], (k / kTileStages) & 1);
      const int nvalid = min(kTT, n_out - k * kTT);
      float* davg = tile_d + ts * kTT;
      const float* ta = tile_a + ts * kTT;
      const float2* ty = tile_y + ts * kTT;
      if (!terminated) {
        // ---- avg_ampl recurrence (gate_impl.cc:131): one dependent add per sample, lane 0
        if (lane == 0) chain_inplace(davg, nvalid, acc);
        __syncwarp();
        // ---- threshold flags (gate_impl.cc:136,148,154)
        unsigned lt[4], gt[4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const int i = r * 32 + lane;
          bool v = i < nvalid;
          float thr = v ? f_mul(davg[i], kThreshFraction) : 0.f;
          float a = v ? ta[i] : 0.f;
          lt[r] = __ballot_sync(0xffffffffu, v && a < thr);
          gt[r] = __ballot_sync(0xffffffffu, v && a > thr);
        }
        const unsigned long long LT_lo = lt[0] | ((unsigned long long)lt[1] << 32), LT_hi = lt[2] | ((unsigned long long)lt[3] << 32);
        const unsigned long long GT_lo = gt[0] | ((unsigned long long)gt[1] << 32), GT_hi = gt[2] | ((unsigned long long)gt[3] << 32);

        int pos = 0;
        while (pos < nvalid) {
          if (!gate_open) {
            // ---- closed: hop from edge to edge (gate_impl.cc:145-162) until the tile ends or the gate opens
            const int run_start = pos;
            bool opened = false;
            while (pos < nvalid) {
              if (sig_pos) {
                const int p_fall = next_set128(LT_lo, LT_hi, pos);
                int p_open = 1 << 30;
                if (num_pulses > kNumPulsesCommand) p_open = pos + max(0, C.n_T1 - n_samples);
                if (p_fall >= nvalid && p_open >= nvalid) { n_samples += nvalid - pos; pos = nvalid; break; }
                if (p_fall <= p_open) { n_samples = 0; sig_pos = false; pos = p_fall + 1; }
                else { pos = p_open + 1; opened = true; break; }
              } else {
                const int p_rise = next_set128(GT_lo, GT_hi, pos);
                if (p_rise >= nvalid) { n_samples += nvalid - pos; pos = nvalid; break; }
                const int n_at = n_samples + (p_rise - pos + 1);
                num_pulses = (n_at > half_pw) ? num_pulses + 1 : 0;
                n_samples = 0; sig_pos = true; pos = p_rise + 1;
              }
            }
            // ---- DC tracker over the closed run [run_start, pos) (gate_impl.cc:141-143); the run
            //      includes the trigger sample, as in the reference (update precedes the open test)
            const int len = pos - run_start;
            for (int j = lane; j < len; j += 32) ycl[(n_closed + j) & ymask] = ty[run_start + j];
            __syncwarp();
            for (int j = lane; j < len; j += 32) {
              const float2 yv = ty[run_start + j];
              const float2 old = ycl[(n_closed + j - C.dc_length) & ymask];
              e_re[j] = f_div(f_sub(yv.x, old.x), dclen_f);
              e_im[j] = f_div(f_sub(yv.y, old.y), dclen_f);
            }
            __syncwarp();
            if (lane == 1) chain_inplace(e_re, len, acc);
            if (lane == 2) chain_inplace(e_im, len, acc);
            __syncwarp();
            n_closed += len;
            if (opened) {
              // READER COMMAND DETECTED (gate_impl.cc:164-180)
              dc_open = make_float2(__shfl_sync(0xffffffffu, acc, 1), __shfl_sync(0xffffffffu, acc, 2));
              gate_open = true;
              open_idx = k * kTT + pos - 1;
              store_this = wcount < A.max_windows;
              if (store_this && wsignalled > 0) mbar_wait(&B.win_free, (wsignalled - 1) & 1);  // window buffer free
              if (store_this && lane == 0) win[0] = c_sub(ty[pos - 1], dc_open);
              wpos = 1;
              num_pulses = 0;
              n_samples = 1;
            }
          } else {
            // ---- open: pass samples through with the frozen DC estimate (gate_impl.cc:182-195)
            const int take = min(to_ungate - n_samples, nvalid - pos);
            if (store_this)
              for (int j = lane; j < take; j += 32) win[wpos + j] = c_sub(ty[pos + j], dc_open);
            wpos += take; n_samples += take; pos += take;
            if (n_samples >= to_ungate) {
              gate_open = false;
              const int kind = wcount & 1;  // windows alternate RN16, EPC
              if (store_this) {
                __syncwarp();
                if (lane == 0) { B.meta_kind = kind; B.meta_open = open_idx; B.meta_ordinal = wcount; B.meta_len = to_ungate; }
                __syncwarp();
                if (lane == 0) mbar_arrive(&B.win_ready);
                wsignalled++;
              }
              wcount++;
              // the Gen2 logic answers (ACK after RN16 -> GATE_SEEK_EPC, Query/QueryRep after EPC ->
              // GATE_SEEK_RN16) and the next gate call applies it (gate_impl.cc:112-123)
              to_ungate = kind ? C.len_rn16 : C.len_epc;
              n_samples = 0;
              if (kind) {
                nq++;
                if (nq > C.max_queries) { terminated = true; break; }  // gate_impl.cc:101-109
              }
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&B.tile_empty[ts]);
    }
    // ---- shut the decoder down, publish the window count
    if (wsignalled > 0) mbar_wait(&B.win_free, (wsignalled - 1) & 1);
    if (lane == 0) {
      B.meta_kind = -1;
      A.counts[seg] = wcount;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&B.win_ready);
  } else {
    // =========================================================== decoder
    for (int j = 0;; j++) {
      mbar_wait(&B.win_ready, j & 1);
      const int kind = B.meta_kind;
      if (kind < 0) break;
      const int ordinal = B.meta_ordinal, open_idx = B.meta_open, len = B.meta_len;
      WindowDecode wd;
      decode_window_warp(C, kind, win, len, Msq, wd);
      rfid_b200_window_result* dst = A.results + (size_t)seg * A.max_windows + ordinal;
      if (lane == 0) store_result(dst, wd, seg, ordinal, open_idx, len, kind);
      if (A.window_tap) {
        float2* tap = A.window_tap + ((size_t)seg * A.max_windows + ordinal) * C.len_epc;
        for (int p = lane; p < len; p += 32) tap[p] = win[p];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&B.win_free);
    }
  }
}

}  // namespace rfid_b200
