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
#ifndef RFID_B200_CHAIN_WARP_OFS
#define RFID_B200_CHAIN_WARP_OFS 1
#endif
constexpr int kChainWarpOfs = RFID_B200_CHAIN_WARP_OFS;   // chain warp = 2G + this, loader = 2G + (1 - this): picks the chain's SM sub-partition
constexpr int kPChainBuf = kTT + 12;         // floats per running-sum buffer: read-ahead pad; 560 B = 48 mod 128
constexpr int kPackMaxThreads = 32 * (3 * kPMaxSeg + 2);

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
  uint64_t raw_empty[kPRawStages];  // warp A (1 arrival) -> loader warp
  uint64_t go, done;         // warp B -> warp C: "emit tile go_tile" / warp C -> warp B: "copied, the ring slot may go"
  int go_tile;               // tile to emit; -1: the segment is over
  int emit[kPS];             // does tile t (slot t & 3) contain window samples or gate events?  (written by P3)
  int n_e[kPS];
  int n_ev[kPS];
  TileEvent ev[kPS][kMaxTileEvents];
};

#ifdef RFID_B200_PHASE_PROFILE
// developer aid (tools/pack_profile.py): per-phase clock64() sums of every tile warp / chain warp, written to the window tap
#define PP_DECL long long pp_t0 = clock64(), pp_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long* pp_step_log = nullptr; long long* pp_abs_log = nullptr; int pp_step = 0;
#define PP_MARK(i) { const long long pp_t1 = clock64(); pp_acc[i] += pp_t1 - pp_t0; if (pp_step_log && lane == 0) { pp_step_log[pp_step * 8 + (i)] = pp_t1 - pp_t0; if (pp_abs_log) pp_abs_log[pp_step * 8 + (i)] = pp_t1 - pp_cta_t0; } pp_t0 = pp_t1; }
#define PP_SUB(i) { const long long pp_t2 = clock64(); if (pp_step_log && lane == 0) pp_step_log[128 * 8 + pp_step * 8 + (i)] = pp_t2 - pp_t0; }
#define PP_SUBB(i) { const long long pp_t2 = clock64(); if (pp_step_log && lane == 0) pp_step_log[64 * 8 + pp_step * 8 + (i)] = pp_t2 - pp_t0; }
#define PP_DUMP(row) if (lane == 0 && A.window_tap) { long long* o = reinterpret_cast<long long*>(A.window_tap) + (size_t)(row) * 8; for (int q_ = 0; q_ < 8; q_++) o[q_] = pp_acc[q_]; }
#else
#define PP_DECL
#define PP_MARK(i)
#define PP_SUB(i)
#define PP_SUBB(i)
#define PP_DUMP(row)
#endif

// release / acquire fence at CTA scope (fence.sc, which __threadfence_block() emits, is far more expensive and not needed:
// the window hand-off is a plain producer -> consumer edge)
__device__ __forceinline__ void fence_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }

#ifdef RFID_B200_EXP_RELAXED_ARRIVE
#define MBAR_ARRIVE_EXP mbar_arrive_relaxed
#else
#define MBAR_ARRIVE_EXP mbar_arrive
#endif
enum : int { PBAR_X = 1, PBAR_Y = 3, PBAR_PAIR = 5 };  // + segment slot: warps A and B of one segment
__device__ __forceinline__ void pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
template <int BASE>
__device__ __forceinline__ void pbar_sync(int parity, int count)
{
  if (parity == 0) asm volatile("bar.sync %0, %1;" ::"n"((int)BASE), "r"(count) : "memory");
  else asm volatile("bar.sync %0, %1;" ::"n"((int)(BASE + 1)), "r"(count) : "memory");
}
// (bar.arrive / bar.sync order the executing thread's prior shared-memory accesses for the threads that complete the
// barrier -- the PTX producer/consumer pattern; no separate fence, which would cost a MEMBAR.SC on the critical path)
template <int BASE>
__device__ __forceinline__ void pbar_arrive(int parity, int count)
{
  if (parity == 0) asm volatile("bar.arrive %0, %1;" ::"n"((int)BASE), "r"(count) : "memory");
  else asm volatile("bar.arrive %0, %1;" ::"n"((int)(BASE + 1)), "r"(count) : "memory");
}


// ---- the edge / pulse state machine of one closed run as pure 128-bit mask arithmetic -------------------------------
// Same decisions as fsm_closed_run (rx_fused_split.cuh) and therefore as the reference's sample loop
// (gate_impl.cc:145-180), but without per-edge work: every quantity below is warp-uniform, computed redundantly by all
// lanes from the two 128-bit threshold masks -- no shuffles, no ballots, no find-nth-set-bit.
//   states     carry chain (state' = rise | keep & state), as before
//   pulses     a rise at p is a valid pulse when the fall before it is more than half_pw back: no fall bit at p-1 .. p-half_pw
//   num_pulses valid rises since the last invalid one (plus the carried count while no invalid rise has occurred)
//   opening    the carried state reaches T1 before the first edge, or a rise r with num_pulses > 5 is followed by
//              n_T1 + 1 edge-free samples inside the tile (gate opens at r + 1 + n_T1; a fall there wins)
__device__ __forceinline__ unsigned m_below(int n) { return n <= 0 ? 0u : (n >= 32 ? 0xffffffffu : ((1u << n) - 1u)); }
__device__ __forceinline__ int m_ffs128(const unsigned* a)
{
  const unsigned long long lo = a[0] | ((unsigned long long)a[1] << 32), hi = a[2] | ((unsigned long long)a[3] << 32);
  return lo ? __ffsll((long long)lo) - 1 : (hi ? 63 + __ffsll((long long)hi) : 128);
}
__device__ __forceinline__ int m_fls128(const unsigned* a)
{
  const unsigned long long lo = a[0] | ((unsigned long long)a[1] << 32), hi = a[2] | ((unsigned long long)a[3] << 32);
  return hi ? 127 - __clzll((long long)hi) : (lo ? 63 - __clzll((long long)lo) : -1);
}
// number of set bits of a[] at positions [lo, hi)
__device__ __forceinline__ int m_popc_range(const unsigned* a, int lo, int hi)
{
  int n = 0;
#pragma unroll
  for (int w = 0; w < 4; w++) n += __popc(a[w] & m_below(hi - 32 * w) & ~m_below(lo - 32 * w));
  return n;
}
__device__ __forceinline__ bool m_any_range(const unsigned* a, int lo, int hi)
{
  unsigned v = 0;
#pragma unroll
  for (int w = 0; w < 4; w++) v |= a[w] & m_below(hi - 32 * w) & ~m_below(lo - 32 * w);
  return v != 0u;
}
// single-bit helpers with static word indices (a dynamically indexed register array would be demoted to local memory)
__device__ __forceinline__ unsigned m_word_bit(int p, int w) { return (p >> 5) == w ? (1u << (p & 31)) : 0u; }
__device__ __forceinline__ bool m_bit(const unsigned* a, int p)
{
  return ((a[0] & m_word_bit(p, 0)) | (a[1] & m_word_bit(p, 1)) | (a[2] & m_word_bit(p, 2)) | (a[3] & m_word_bit(p, 3))) != 0u;
}
__device__ __forceinline__ void m_clear(unsigned* a, int p)
{
#pragma unroll
  for (int w = 0; w < 4; w++) a[w] &= ~m_word_bit(p, w);
}

__device__ __forceinline__ int fsm_closed_run_masks(const Mask128& lt, const Mask128& gt, int from, int nvalid, int n_T1,
                                                    int half_pw, GateFsm& st)
{
  unsigned F[4], R[4], X[4], RS[4], FE[4], E[4];
  unsigned carry = st.sig_pos ? 1u : 0u;
#pragma unroll
  for (int w = 0; w < 4; w++) {
    const unsigned live = ~m_below(from - 32 * w);
    F[w] = lt.w[w] & live;
    R[w] = gt.w[w] & live;
    const unsigned Pk = ~(F[w] | R[w]);
    const unsigned long long sum = (unsigned long long)(R[w] | Pk) + R[w] + carry;
    X[w] = Pk ^ (unsigned)sum;  // carry INTO each bit = state before that position
    carry = (unsigned)(sum >> 32);
    RS[w] = ~X[w] & R[w];
    FE[w] = X[w] & F[w];
    E[w] = RS[w] | FE[w];
  }
  const int e_first = m_ffs128(E);
  const int first_edge = e_first < nvalid ? e_first : nvalid;
  // ---- the carried state reaches T1 before anything happens
  if (st.sig_pos && st.num_pulses > kNumPulsesCommand) {
    const int p_open = from + max(0, n_T1 - st.n_samples);
    if (p_open < first_edge && p_open < nvalid) {
      st.sig_pos = true; st.num_pulses = 0; st.n_samples = 1;
      return p_open;
    }
  }
  if (e_first >= nvalid) { st.n_samples += nvalid - from; return -1; }
  // ---- valid / invalid pulses
  unsigned VR[4], IR[4];
#pragma unroll
  for (int w = 0; w < 4; w++) VR[w] = RS[w];
  {
    unsigned S[4] = {FE[0], FE[1], FE[2], FE[3]};
    for (int k = 0; k < half_pw; k++) {  // fall bits moved up by k+1 positions knock the rises they reach out
      const unsigned s3 = (S[3] << 1) | (S[2] >> 31), s2 = (S[2] << 1) | (S[1] >> 31), s1 = (S[1] << 1) | (S[0] >> 31), s0 = S[0] << 1;
      S[0] = s0; S[1] = s1; S[2] = s2; S[3] = s3;
#pragma unroll
      for (int w = 0; w < 4; w++) VR[w] &= ~S[w];
    }
  }
  if (m_bit(RS, e_first)) {
    // the tile's first edge is a rise: its pulse began before the run (n_samples carried in)
    const bool valid = st.n_samples + (e_first - from + 1) > half_pw;
    if (!valid) m_clear(VR, e_first);
  }
#pragma unroll
  for (int w = 0; w < 4; w++) IR[w] = RS[w] & ~VR[w];
  const int np_in = st.num_pulses;
  auto np_at = [&](int r) {  // num_pulses right after the rise at r
    unsigned res[4];
#pragma unroll
    for (int w = 0; w < 4; w++) res[w] = IR[w] & m_below(r + 1 - 32 * w);
    const int last = m_fls128(res);
    return last >= 0 ? m_popc_range(VR, last + 1, r + 1) : np_in + m_popc_range(VR, 0, r + 1);
  };
  // ---- a rise followed by n_T1 + 1 quiet samples inside the tile
  const int lim = nvalid - 1 - n_T1;  // rises at or above lim cannot open the gate within this tile
  if (lim > 0) {
    unsigned cand[4];
#pragma unroll
    for (int w = 0; w < 4; w++) cand[w] = RS[w] & m_below(lim - 32 * w);
    if (lim <= n_T1 + 1) {
      // (any earlier rise below lim is followed by a fall within its T1: only the last one can qualify)
      const int r = m_fls128(cand);
#pragma unroll
      for (int w = 0; w < 4; w++) cand[w] = r >= 0 ? m_word_bit(r, w) : 0u;
    }
    while ((cand[0] | cand[1] | cand[2] | cand[3]) != 0u) {
      const int r = m_ffs128(cand);
      m_clear(cand, r);
      if (!m_any_range(E, r + 1, r + 2 + n_T1) && np_at(r) > kNumPulsesCommand) {
        st.sig_pos = true; st.num_pulses = 0; st.n_samples = 1;
        return r + 1 + n_T1;
      }
    }
  }
  // ---- no opening: state after the last edge
  const int e_last = m_fls128(E);
  st.sig_pos = m_bit(RS, e_last);
  st.n_samples = nvalid - 1 - e_last;
  const int r_last = m_fls128(RS);
  if (r_last >= 0) st.num_pulses = np_at(r_last);
  return -1;
}


// ---- the same state machine with the four mask words spread over lanes 0..3 ---------------------------------------
// fsm_closed_run_masks walks every 128-bit quantity as four dependent 32-bit words in every lane: ~400 dependent scalar
// instructions, i.e. a couple of thousand cycles of latency for one warp.  Here lane w (w < 4) owns word w; positions and
// counts are combined with redux.sync / vote (one instruction each), the carry chain across the words by carry look-ahead
// (both sums per word, a 3-step select on two ballots).  Identical decisions; same arguments and result.
__device__ __forceinline__ int fsm_closed_run_lanes(const Mask128& lt, const Mask128& gt, int from, int nvalid, int n_T1,
                                                    int half_pw, GateFsm& st)
{
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int w = lane & 3;
  const bool own = lane < 4;
  const int base = 32 * w;
  const unsigned ltw = w == 0 ? lt.w[0] : (w == 1 ? lt.w[1] : (w == 2 ? lt.w[2] : lt.w[3]));
  const unsigned gtw = w == 0 ? gt.w[0] : (w == 1 ? gt.w[1] : (w == 2 ? gt.w[2] : gt.w[3]));
  const unsigned live = ~m_below(from - base);
  const unsigned F = ltw & live, R = gtw & live;
  const unsigned Pk = ~(F | R);
  // carry look-ahead over the four words
  const unsigned long long sum0 = (unsigned long long)(R | Pk) + R, sum1 = sum0 + 1ull;
  const unsigned C0 = __ballot_sync(FULL, own && (sum0 >> 32) != 0ull), C1 = __ballot_sync(FULL, own && (sum1 >> 32) != 0ull);
  const unsigned cin0 = st.sig_pos ? 1u : 0u;
  const unsigned cin1 = cin0 ? (C1 & 1u) : (C0 & 1u);
  const unsigned cin2 = cin1 ? ((C1 >> 1) & 1u) : ((C0 >> 1) & 1u);
  const unsigned cin3 = cin2 ? ((C1 >> 2) & 1u) : ((C0 >> 2) & 1u);
  const unsigned cin = w == 0 ? cin0 : (w == 1 ? cin1 : (w == 2 ? cin2 : cin3));
  const unsigned X = Pk ^ (unsigned)(cin ? sum1 : sum0);   // state before each position
  const unsigned RS = ~X & R, FE = X & F, E = RS | FE;
  auto first_of = [&](unsigned m) { return __reduce_min_sync(FULL, (own && m) ? base + __ffs(m) - 1 : 128); };
  auto last_of = [&](unsigned m) { return __reduce_max_sync(FULL, (own && m) ? base + 31 - __clz(m) : -1); };
  auto bit_at = [&](unsigned m, int p) { return __any_sync(FULL, own && (p >> 5) == w && ((m >> (p & 31)) & 1u)); };
  const int e_first = first_of(E);
  const int first_edge = e_first < nvalid ? e_first : nvalid;
  if (st.sig_pos && st.num_pulses > kNumPulsesCommand) {
    const int p_open = from + max(0, n_T1 - st.n_samples);
    if (p_open < first_edge && p_open < nvalid) {
      st.sig_pos = true; st.num_pulses = 0; st.n_samples = 1;
      return p_open;
    }
  }
  if (e_first >= nvalid) { st.n_samples += nvalid - from; return -1; }
  // valid pulses: no fall within the half_pw positions before the rise
  const unsigned fe_prev = __shfl_up_sync(FULL, FE, 1);
  const unsigned fe_lo = w == 0 ? 0u : fe_prev;
  unsigned knock = 0u;
  for (int k = 1; k <= half_pw; k++) knock |= (FE << k) | (fe_lo >> (32 - k));
  unsigned VR = RS & ~knock;
  if (bit_at(RS, e_first)) {
    const bool valid = st.n_samples + (e_first - from + 1) > half_pw;   // the pulse began before the run
    if (!valid && (e_first >> 5) == w) VR &= ~(1u << (e_first & 31));
  }
  const unsigned IR = RS & ~VR;
  const int np_in = st.num_pulses;
  auto np_at = [&](int r) {
    const unsigned upto = m_below(r + 1 - base);
    const int last = last_of(IR & upto);
    const unsigned rng = upto & ~m_below(last + 1 - base);
    const int cnt = __reduce_add_sync(FULL, own ? __popc(VR & rng) : 0);
    return last >= 0 ? cnt : np_in + cnt;
  };
  const int lim = nvalid - 1 - n_T1;
  if (lim > 0) {
    unsigned cand = RS & m_below(lim - base);
    if (lim <= n_T1 + 1) {
      const int r = last_of(cand);
      cand = (r >= 0 && (r >> 5) == w) ? (1u << (r & 31)) : 0u;
    }
    while (true) {
      const int r = first_of(cand);
      if (r >= 128) break;
      if ((r >> 5) == w) cand &= ~(1u << (r & 31));
      const unsigned quiet_rng = m_below(r + 2 + n_T1 - base) & ~m_below(r + 1 - base);
      if (!__any_sync(FULL, own && (E & quiet_rng) != 0u) && np_at(r) > kNumPulsesCommand) {
        st.sig_pos = true; st.num_pulses = 0; st.n_samples = 1;
        return r + 1 + n_T1;
      }
    }
  }
  const int e_last = last_of(E);
  st.sig_pos = bit_at(RS, e_last);
  st.n_samples = nvalid - 1 - e_last;
  const int r_last = last_of(RS);
  if (r_last >= 0) st.num_pulses = np_at(r_last);
  return -1;
}

template <int DECIM, int MFQ>
__global__ void __launch_bounds__(kPackMaxThreads, 1) rx_pack_kernel(const PackArgs A)
{
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ PackSegCtl ctl_all[kPMaxSeg];
#ifdef RFID_B200_PHASE_PROFILE
  __shared__ long long pp_cta_t0;
  if (threadIdx.x == 0) pp_cta_t0 = clock64();
#endif

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
    for (int s = 0; s < kPRawStages; s++) { mbar_init(&c.raw_full[s], 1); mbar_init(&c.raw_empty[s], 1); }
    mbar_init(&c.go, 1); mbar_init(&c.done, 1);
    for (int s = 0; s < kPS; s++) { c.n_e[s] = 0; c.n_ev[s] = 0; c.emit[s] = 0; }
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

  if (warp < 2 * G) {
    // ======================================================================================= tile warps A (P1) and B (E, P3)
    const bool is_a = warp < G;
    const int g = is_a ? warp : warp - G;
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

    const int odd = (int)(sg.offset & 1ull);
    const float dclen_f = (float)C.dc_length;
    const int pair_bar = PBAR_PAIR + g;

    if (is_a) {
      // ------------------------------------------------------------------------------------- warp A: P1
    const float winlen_f = (float)C.win_length;
    float2 b_keep[MFQ - 1];
#pragma unroll
    for (int m = 0; m < MFQ - 1; m++) b_keep[m] = make_float2(0.f, 0.f);
    int rs = 0;
    uint32_t raw_par = 0;

    PP_DECL
#ifdef RFID_B200_PHASE_PROFILE
    if (seg == 0 && A.window_tap) { pp_step_log = reinterpret_cast<long long*>(A.window_tap) + (size_t)(A.nseg + gridDim.x) * 8; pp_abs_log = pp_step_log + 320 * 8; }
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
        __syncwarp();  // raw stage consumed
        if (lane == 0) MBAR_ARRIVE_EXP(&B.raw_empty[rs]);
        PP_SUB(0)
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
      pbar_arrive<PBAR_X>(i & 1, bar_count);   // tile i is ready for the chain warp
      pair_sync(pair_bar);                      // warp B is done with step i: ring slot, sum buffers of tile i+1 are free
    }
    if (have) { PP_DUMP(seg) }
    } else {
      // ------------------------------------------------------------------------------------- warp B: TMA issue, E, P3
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

    uint32_t go_count = 0;      // emission requests sent to warp C so far (phase parity of the done barrier)

    PP_DECL
#ifdef RFID_B200_PHASE_PROFILE
    if (seg == 0 && A.window_tap) { pp_step_log = reinterpret_cast<long long*>(A.window_tap) + (size_t)(A.nseg + gridDim.x) * 8 + 192 * 8; pp_abs_log = pp_step_log + (384 - 192) * 8; }
#endif
    for (int i = 0; i < nsteps; i++) {
#ifdef RFID_B200_PHASE_PROFILE
      pp_step = i;
#endif
      PP_MARK(5)
      PP_MARK(0)
      if (i >= 1) pbar_sync<PBAR_Y>((i - 1) & 1, bar_count);  // avg_ampl of tile i-1 and dc_est of tile i-3 are final
      PP_MARK(2)
      // ================================================================= E(i-3) is warp C's: ring the bell, carry on with P3
      const bool need_emit = i >= 3 && i - 3 < ntiles && B.emit[(i - 3) & (kPS - 1)] != 0;
      if (need_emit && lane == 0) { B.go_tile = i - 3; MBAR_ARRIVE_EXP(&B.go); }
      PP_MARK(3)
      // ================================================================= P3(i-1): thresholds, state machine, DC list
      if (i >= 1 && i - 1 < ntiles) {
        const int t = i - 1, s = t & (kPS - 1);
        int nev = 0, n_e = 0;
        const bool open_at_start = gate_open;
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
              PP_SUBB(3)
              if (!have_masks && !(quiet && run_start == 0)) make_masks();
              PP_SUBB(4)
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
#ifdef RFID_B200_FSM_PARALLEL_EDGES
                p_open = fsm_closed_run(ltm, gtm, run_start, nvalid, C.n_T1, half_pw, fs);
#elif defined(RFID_B200_FSM_SCALAR_MASKS)
                p_open = fsm_closed_run_masks(ltm, gtm, run_start, nvalid, C.n_T1, half_pw, fs);
#else
                p_open = fsm_closed_run_lanes(ltm, gtm, run_start, nvalid, C.n_T1, half_pw, fs);
#endif
                sig_pos = fs.sig_pos; n_samples = fs.n_samples; num_pulses = fs.num_pulses;
              }
              PP_SUBB(5)
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
        if (lane == 0) { B.n_e[s] = n_e; B.n_ev[s] = min(nev, kMaxTileEvents); B.emit[s] = (open_at_start || nev > 0) ? 1 : 0; }
        __syncwarp();
      }
      PP_MARK(4)
      if (need_emit) {  // warp C has copied tile i-3 out of the ring (it ran alongside P3)
        mbar_wait(&B.done, go_count & 1u);
        go_count++;
      }
      pair_sync(pair_bar);
    }
    // ---- end of the segment
    if (have && lane == 0) A.counts[seg] = wcount;
    if (lane == 0) { B.go_tile = -1; MBAR_ARRIVE_EXP(&B.go); }
    }
  } else if (warp == kChainWarpOfs + 2 * G) {
    // ======================================================================================= chain warp
    // lane 8*c + g: running sum c (0 avg_ampl, 1 dc_est.re, 2 dc_est.im) of segment g
    const int comp = lane >> 3, g = lane & 7;
    const bool active = comp < 3 && g < g_act;
    int n_out = 0;
    if (active) n_out = (int)(A.segs[seg0 + g].length / DECIM);
    float acc = 0.f;
    PP_DECL
#ifdef RFID_B200_PHASE_PROFILE
    if (blockIdx.x == 0 && A.window_tap) { pp_step_log = reinterpret_cast<long long*>(A.window_tap) + (size_t)(A.nseg + gridDim.x) * 8 + 64 * 8; pp_abs_log = pp_step_log + (448 - 64) * 8; }
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
  } else if (warp == (1 - kChainWarpOfs) + 2 * G) {
    // ======================================================================================= loader warp
    // lane g streams the raw tiles of segment g through its three stages (TMA bulk copies), as far ahead as warp A frees them
    const int g = lane;
    if (g < g_act) {
      const rfid_b200_segment sg = A.segs[seg0 + g];
      const int n_out = (int)(sg.length / DECIM);
      const int ntiles = (n_out + kTT - 1) / kTT;
      PackSegCtl& B = ctl_all[g];
      float2* raw = reinterpret_cast<float2*>(smem + A.off_seg + (size_t)g * A.seg_bytes + A.o_raw);
      // raw tile geometry (as rx_fused_split.cuh)
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
for (int k = 0; k < ntiles; k++) {
        const int rs_ = k % kPRawStages;
        if (k >= kPRawStages) mbar_wait_relaxed(&B.raw_empty[rs_], ((k / kPRawStages) - 1) & 1, 4000);
        load_tile(k, rs_);
      }
    }
  } else {
    // ======================================================================================= warp C: emission + decode
    // Copies the window samples of a tile (y - dc_est, gate_impl.cc:173,187) from the time ring into this segment's
    // scratch when warp B asks for it, confirms (the ring slot is recycled one step later), and decodes every window
    // that closed.  Writer and reader of the scratch are this one warp: no fence, no progress counter, no flow control.
    const int g = warp - 2 * G - 2;
    if (g < g_act) {
      const int seg = seg0 + g;
      const int n_out = (int)(A.segs[seg].length / DECIM);
      PackSegCtl& B = ctl_all[g];
      unsigned char* sb = smem + A.off_seg + (size_t)g * A.seg_bytes;
      const float2* ring_y = reinterpret_cast<const float2*>(sb + A.o_ring_y);
      float2* dstage = reinterpret_cast<float2*>(sb + A.o_dstage);
      float2* const win_base = A.win_scratch + (size_t)seg * A.win_stride;
      bool f_open = false, f_store = false;
      int f_wpos = 0, wsig_ordinal = 0, cur_ordinal = 0, cur_kind = 0, cur_open_idx = 0;
      float2 dc_open = make_float2(0.f, 0.f);
      float2* win = win_base;
      for (uint32_t rq = 0;; rq++) {
        mbar_wait_relaxed(&B.go, rq & 1u, 20000);  // parked by the hardware until warp B rings
        const int t = B.go_tile;
        if (t < 0) break;
        const int ps = t & (kPS - 1);
        const float2* py = ring_y + ps * kTT;
        const float* pe_re = dD + (size_t)((ps * 2 + 0) * G + g) * kPChainBuf;
        const float* pe_im = dD + (size_t)((ps * 2 + 1) * G + g) * kPChainBuf;
        const int pvalid = min(kTT, n_out - t * kTT);
        const int pnev = B.n_ev[ps];
        int n_closed = 0, c_kind[2] = {0, 0}, c_ord[2] = {0, 0}, c_open[2] = {0, 0};
        int pos = 0;
        for (int e = 0; e <= pnev; e++) {
          const bool last = e == pnev;
          const int etype = last ? 0 : B.ev[ps][e].type;
          const int epos = last ? pvalid : B.ev[ps][e].pos;
          if (f_open) {
            const int take = epos - pos;
            if (f_store && take > 0)
              for (int j = lane; j < take; j += 32) win[f_wpos + j] = c_sub(py[pos + j], dc_open);
            f_wpos += take;
            pos = epos;
          }
          if (etype == 2) {
            if (f_open && f_store && n_closed < 2) { c_kind[n_closed] = cur_kind; c_ord[n_closed] = cur_ordinal; c_open[n_closed] = cur_open_idx; n_closed++; }
            f_open = false;
            pos = epos;
          } else if (etype == 1) {
            const int j = B.ev[ps][e].a;
            dc_open = make_float2(pe_re[j], pe_im[j]);  // dc_est right after the trigger sample
            f_store = B.ev[ps][e].c != 0;
            f_open = true;
            cur_kind = B.ev[ps][e].d;
            cur_ordinal = wsig_ordinal++;
            cur_open_idx = B.ev[ps][e].b;
            win = win_base + (cur_kind ? A.rn16_pad : 0);
            if (f_store && lane == 0) win[0] = c_sub(py[epos], dc_open);
            f_wpos = 1;
            pos = epos + 1;
          }
        }
        __syncwarp();
        if (lane == 0) MBAR_ARRIVE_EXP(&B.done);  // (release: the ring reads above are complete)
        for (int k = 0; k < n_closed; k++) {
          const int kind = c_kind[k], len = kind ? C.len_epc : C.len_rn16;
          const float2* wv = win_base + (kind ? A.rn16_pad : 0);
          WindowDecode wd;
          decode_window_staged(C, kind, wv, len, dstage, A.dstage_samples, wd);
          rfid_b200_window_result* dst = A.results + (size_t)seg * A.max_windows + c_ord[k];
          if (lane == 0) store_result(dst, wd, seg, c_ord[k], c_open[k], len, kind);
#ifndef RFID_B200_PHASE_PROFILE
          if (A.window_tap) {
            float2* tap = A.window_tap + ((size_t)seg * A.max_windows + c_ord[k]) * C.len_epc;
            for (int p2 = lane; p2 < len; p2 += 32) tap[p2] = __ldcg(wv + p2);
          }
#endif
          __syncwarp();
        }
      }
    }
  }
}

}  // namespace rfid_b200
