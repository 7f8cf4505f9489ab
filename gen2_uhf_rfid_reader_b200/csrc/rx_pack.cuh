// rx_pack.cuh -- the roofline kernel of the reference configuration (decimation 5, 25 taps, rings inside one tile):
// same arithmetic as rx_fused_split.cuh, re-scheduled around two measurements (tools/pack_profile.py):
//   * the exact replay of the gate's float running sums (avg_ampl, gate_impl.cc:131; dc_est, gate_impl.cc:141) is a
//     dependent FADD chain -- it costs the same whether 3 or 24 lanes of the warp carry a chain;
//   * everything else of a segment is one warp's worth of *latency*, not of issue slots: a warp that walks a tile through
//     matched filter, |y|, ring differences (or thresholds + the edge / pulse state machine) needs 2-3 thousand cycles per
//     pass almost independently of how many samples the pass covers.
// So a CTA owns G <= kPMaxSeg capture segments and runs them in lockstep, one 256-sample tile per step, with one warp per
// role and segment, and ONE chain warp for all of them:
//   warp A[g]   P1: waits for the two raw half-tiles (TMA bulk copies issued by the loader warp), block-sum matched
//               filter, exact |y|, amplitude / DC ring differences of 8 outputs per lane (two groups of 4 consecutive
//               ones, 128 apart: twice the independent work per dependency chain of the 128-sample version).
//   chain warp  lane 8*c + g replays running sum c (0 avg_ampl, 1 dc_est.re, 2 dc_est.im) of segment g.
//   warp B[g]   P3: thresholds by ballot, the edge / pulse state machine on 256-bit masks spread over 8 lanes, the DC-ring
//               differences around gate activity; posts dc_est of every opening window when the chain has produced it.
//   warp C[g]   copies the window samples of a tile from the time ring to the segment's L2-resident scratch one step
//               after P3 (as y; dc_est is subtracted when the decoder reads them -- the same exact float subtraction,
//               gate_impl.cc:173,187), then decodes every window that closed (rx_decode.cuh).  It is the only writer and
//               the only reader of the scratch: no fence, no progress counter.
//   loader      lane g streams segment g's raw samples through its two half-tile stages as warp A frees them.
// Hand-offs per step: two named barriers (X: tile ready -> chain, Y: sums ready -> B; a parked warp costs no issue slots),
// one pair barrier A[g] <-> B[g] (ring slot / sum buffers of the next tile are free), mbarriers for loader and warp C.
//   step i:  A: P1(i) | arrive X(i) | pair(i)        chain: sync X(i) | avg_ampl(i), dc_est(i-2) | arrive Y(i)
//            B: sync Y(i-1) | dc of windows opened in tile i-3 | P3(i-1) | C done(i-2)? | ring C for tile i-1 | pair(i)
// Shared memory per segment: 2 raw half-tile stages (10 KB), a 3-tile time ring of y (6 KB) and a 2-tile ring of |y|
// (2 KB), ring snapshot; per CTA the running-sum buffers (2 + 4x2 per segment, 1072 B each, skewed so the chain warp's
// 128-bit accesses are bank-conflict free).  HBM traffic is unchanged: every raw sample is read once, 64 B are written per
// window (plus the window scratch, which lives in L2).
#pragma once

#include "rx_fused_split.cuh"

namespace rfid_b200 {

constexpr int kT2 = 2 * kTT;                 // decimated samples per step (two half-tiles of kTT)
constexpr int kRingY = 3 * kT2;              // time ring of y: tile i, i-1 (P3, emission), the tail of i-2 (DC lookback)
constexpr int kRingA = 2 * kT2;              // time ring of |y|: tile i, i-1
constexpr int kPDS = 4;                      // DC-list slots: P1 (i), P3 fix-ups (i-1), chain (i-2), dc hand-over (i-3)
constexpr int kPMaxSeg = 7;                  // segments per CTA (chain lanes 8*c + g, g < 8)
constexpr int kPChainBuf = kT2 + 12;         // floats per running-sum buffer: read-ahead pad; 1072 B = 48 mod 128
constexpr int kPackMaxThreads = 32 * (4 * kPMaxSeg + 2);   // A0, A1, B, C per segment + loader + chain

struct PackArgs {
  const float2* iq;
  unsigned long long n_raw;
  const rfid_b200_segment* segs;
  int nseg;
  int seg_base;                              // added to the segment index stored in the records
  int max_windows;
  rfid_b200_window_result* results;
  int32_t* counts;
  float2* window_tap;
  float2* win_scratch;
  int win_stride, rn16_pad;
  RxConfig cfg;
  int G;                                     // segments per CTA of this launch
  int raw_stage_samples;
  int seg_bytes;                             // per-segment shared-memory region
  int o_raw, o_ring_y, o_ring_a, o_snap, o_dstage;  // offsets inside a segment region
  int dstage_samples;
  int off_dA, off_dD, off_seg;               // offsets from the dynamic shared-memory base
  int smem_bytes;
};

struct PackSegCtl {
  uint64_t raw_full[2];      // loader (TMA transaction count) -> warp A
  uint64_t raw_empty[2];     // warps A (stage 0: A0 and A1's halo read, stage 1: A1) -> loader
  uint64_t half_rdy;         // warp A0 -> warp A1: y and |y| of the first half-tile are in the ring
  float2 keep[2][4];         // warp A1 -> warp A0 (next step): the tile's last MFQ-1 block sums, by tile parity
  uint64_t go, done;         // warp B -> warp C: "emit tile go_tile" / warp C -> warp B: "copied, the ring slot may go"
  uint64_t dc_rdy[2];        // warp B -> warp C: dc_est of the window in slot (kind) is in dc_val
  float2 dc_val[2];
  int go_tile;               // tile to emit; -1: the segment is over
  int emit[kPDS];            // does tile t (slot t & 3) contain window samples or gate events?  (written by P3)
  int n_e[kPDS];
  int n_ev[kPDS];
  TileEvent ev[kPDS][kMaxTileEvents];
};

#ifdef RFID_B200_PHASE_PROFILE
// developer aid (tools/pack_profile.py): absolute clock64() stamps (since CTA start) of warps A, B, C, chain of CTA 0
#define PP_DECL long long* pp_log = nullptr;
#define PP_AT(i) { if (pp_log && lane == 0) pp_log[pp_step * 8 + (i)] = clock64() - pp_cta_t0; }
#else
#define PP_DECL
#define PP_AT(i)
#endif

enum : int { PBAR_X = 1, PBAR_Y = 3, PBAR_PAIR = 5 };  // + segment slot: warps A and B of one segment
__device__ __forceinline__ void pair_sync(int id) { asm volatile("bar.sync %0, 96;" ::"r"(id) : "memory"); }  // A0, A1, B
// (bar.arrive / bar.sync order the executing thread's prior shared-memory accesses for the threads that complete the
// barrier -- the PTX producer/consumer pattern; no separate fence)
template <int BASE>
__device__ __forceinline__ void pbar_sync(int parity, int count)
{
  if (parity == 0) asm volatile("bar.sync %0, %1;" ::"n"((int)BASE), "r"(count) : "memory");
  else asm volatile("bar.sync %0, %1;" ::"n"((int)(BASE + 1)), "r"(count) : "memory");
}
template <int BASE>
__device__ __forceinline__ void pbar_arrive(int parity, int count)
{
  if (parity == 0) asm volatile("bar.arrive %0, %1;" ::"n"((int)BASE), "r"(count) : "memory");
  else asm volatile("bar.arrive %0, %1;" ::"n"((int)(BASE + 1)), "r"(count) : "memory");
}

// ---- the edge / pulse state machine of one closed run on 256-bit masks, the eight mask words spread over lanes 0..7 ----
// Same decisions as fsm_closed_run (rx_fused_split.cuh) and therefore as the reference's sample loop
// (gate_impl.cc:145-180):
//   states     state' = rise | keep & state is the carry recurrence of a binary addition; across the words by carry
//              look-ahead (both sums per word, a select chain over two ballots)
//   pulses     a rise at p is a valid pulse when the fall before it is more than half_pw back: no fall bit at p-1 .. p-half_pw
//   num_pulses valid rises since the last invalid one (plus the carried count while no invalid rise has occurred)
//   opening    the carried state reaches T1 before the first edge, or a rise r with num_pulses > 5 is followed by
//              n_T1 + 1 edge-free samples inside the tile (gate opens at r + 1 + n_T1; a fall there wins)
// Positions and counts are combined with redux.sync / vote (one instruction each).
struct Mask256 { unsigned w[8]; };
struct GateFsm2 {
  bool sig_pos;
  int n_samples, num_pulses;
};
__device__ __forceinline__ unsigned m_below(int n) { return n <= 0 ? 0u : (n >= 32 ? 0xffffffffu : ((1u << n) - 1u)); }

__device__ __forceinline__ int fsm_closed_run_lanes(const Mask256& lt, const Mask256& gt, int from, int nvalid, int n_T1,
                                                    int half_pw, GateFsm2& st)
{
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int w = lane & 7;
  const bool own = lane < 8;
  const int base = 32 * w;
  unsigned ltw = lt.w[0], gtw = gt.w[0];
#pragma unroll
  for (int k = 1; k < 8; k++) { ltw = w == k ? lt.w[k] : ltw; gtw = w == k ? gt.w[k] : gtw; }
  const unsigned live = ~m_below(from - base);
  const unsigned F = ltw & live, R = gtw & live;
  const unsigned Pk = ~(F | R);
  const unsigned long long sum0 = (unsigned long long)(R | Pk) + R, sum1 = sum0 + 1ull;
  const unsigned C0 = __ballot_sync(FULL, own && (sum0 >> 32) != 0ull), C1 = __ballot_sync(FULL, own && (sum1 >> 32) != 0ull);
  unsigned cin = st.sig_pos ? 1u : 0u, cw = cin;
#pragma unroll
  for (int k = 0; k < 7; k++) {            // carry into word k+1
    cin = cin ? ((C1 >> k) & 1u) : ((C0 >> k) & 1u);
    cw = w == k + 1 ? cin : cw;
  }
  const unsigned X = Pk ^ (unsigned)(cw ? sum1 : sum0);   // state before each position
  const unsigned RS = ~X & R, FE = X & F, E = RS | FE;
  auto first_of = [&](unsigned m) { return __reduce_min_sync(FULL, (own && m) ? base + __ffs(m) - 1 : 1024); };
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
  auto np_at = [&](int r) {  // num_pulses right after the rise at r
    const unsigned upto = m_below(r + 1 - base);
    const int last = last_of(IR & upto);
    const unsigned rng = upto & ~m_below(last + 1 - base);
    const int cnt = __reduce_add_sync(FULL, own ? __popc(VR & rng) : 0);
    return last >= 0 ? cnt : np_in + cnt;
  };
  const int lim = nvalid - 1 - n_T1;  // rises at or above lim cannot open the gate within this tile
  if (lim > 0) {
    unsigned cand;
    if (n_T1 + 1 >= 32) {
      // A rise that opens the gate is followed by n_T1 + 1 >= 32 edge-free positions, so inside its own word it is the
      // highest edge; the first edge of the words above it (suffix minimum over the lanes) decides.  At most
      // nvalid / (n_T1 + 1) rises qualify, so the loop below runs once or twice.
      const int hi = E ? 31 - __clz(E) : -1;
      const int fe = (own && E) ? base + __ffs(E) - 1 : 1024;
      int nx = __shfl_down_sync(FULL, fe, 1);
      if (w == 7) nx = 1024;
      {
        int t1 = __shfl_down_sync(FULL, nx, 1); if (w >= 6) t1 = 1024; nx = min(nx, t1);
        int t2 = __shfl_down_sync(FULL, nx, 2); if (w >= 5) t2 = 1024; nx = min(nx, t2);
        int t4 = __shfl_down_sync(FULL, nx, 4); if (w >= 3) t4 = 1024; nx = min(nx, t4);
      }
      const int r = base + hi;
      const bool q = own && hi >= 0 && ((RS >> hi) & 1u) && r < lim && nx > r + 1 + n_T1;
      cand = q ? (1u << hi) : 0u;
    } else {
      cand = RS & m_below(lim - base);
    }
    while (true) {
      const int r = first_of(cand);
      if (r >= 1024) break;
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
  int pp_step = 0;
#endif

  static_assert(MFQ >= 2, "pack kernel: block-sum matched filter");
  constexpr int Q = kTT / 32;                // outputs per lane and half-tile
  static_assert(MFQ - 1 <= Q, "block-sum halo comes from the neighbouring lane only");
  static_assert(Q == 4, "four consecutive outputs per lane and half-tile");

  const int G = A.G;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int seg0 = blockIdx.x * G;
  const int g_act = min(G, A.nseg - seg0);   // segments this CTA really has
  const RxConfig& C = A.cfg;
  const int bar_x_count = 32 * (2 * G + 1);   // X: both P1 warps of every segment arrive, the chain warp waits
  const int bar_count = 32 * (G + 1);         // Y: the chain warp arrives, every B warp waits

  float* const dA = reinterpret_cast<float*>(smem + A.off_dA);   // [2][G][kPChainBuf]
  float* const dD = reinterpret_cast<float*>(smem + A.off_dD);   // [kPDS][2][G][kPChainBuf]

  // ---- init: barriers first, so that the loader warp can start the first raw half-tiles on their way while the other
  // warps zero the time rings (win_samples / dc_samples start at 0, gate_impl.cc:55-56)
  if (threadIdx.x < G) {
    PackSegCtl& c = ctl_all[threadIdx.x];
    for (int s = 0; s < 2; s++) { mbar_init(&c.raw_full[s], 1); mbar_init(&c.raw_empty[s], s == 0 ? 2 : 1); mbar_init(&c.dc_rdy[s], 1); }
    mbar_init(&c.go, 1); mbar_init(&c.done, 1); mbar_init(&c.half_rdy, 1);
    for (int m = 0; m < 4; m++) { c.keep[0][m] = make_float2(0.f, 0.f); c.keep[1][m] = make_float2(0.f, 0.f); }
    for (int s = 0; s < kPDS; s++) { c.n_e[s] = 0; c.n_ev[s] = 0; c.emit[s] = 0; }
    mbar_fence_init();
  }
  __syncthreads();
  int max_tiles = 0;
  if (warp != 3 * G) {
    for (int g = 0; g < G; g++) {
      unsigned char* sb = smem + A.off_seg + (size_t)g * A.seg_bytes;
      float4* ry = reinterpret_cast<float4*>(sb + A.o_ring_y);
      float4* ra = reinterpret_cast<float4*>(sb + A.o_ring_a);
      const int tid = threadIdx.x - (warp > 3 * G ? 32 : 0), nth = blockDim.x - 32;
      for (int i = tid; i < kRingY / 2; i += nth) ry[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int i = tid; i < kRingA / 4; i += nth) ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // lockstep length: the longest segment of the CTA
    for (int g = 0; g < g_act; g++) {
      const int n_out_g = (int)(A.segs[seg0 + g].length / DECIM);
      max_tiles = max(max_tiles, (n_out_g + kT2 - 1) / kT2);
    }
    asm volatile("bar.sync 15, %0;" ::"r"((int)blockDim.x - 32) : "memory");   // everybody but the loader warp
  }
  const int nsteps = max_tiles + 3;

  if (warp < 3 * G) {
    // ======================================================================================= warps A0 / A1 (P1) and B (P3)
    const bool is_a = warp < 2 * G;
    const int ahalf = warp < G ? 0 : 1;       // which half-tile a P1 warp owns
    const int g = warp < G ? warp : (warp < 2 * G ? warp - G : warp - 2 * G);
    const bool have = g < g_act;
    const int seg = seg0 + g;
    rfid_b200_segment sg;
    sg.offset = 0; sg.length = 0; sg.reserved = 0;
    if (have) sg = A.segs[seg];
    const int n_out = (int)(sg.length / DECIM);
    const int ntiles = (n_out + kT2 - 1) / kT2;
    PackSegCtl& B = ctl_all[g];
    unsigned char* sb = smem + A.off_seg + (size_t)g * A.seg_bytes;
    float2* raw = reinterpret_cast<float2*>(sb + A.o_raw);
    float2* ring_y = reinterpret_cast<float2*>(sb + A.o_ring_y);
    float* ring_a = reinterpret_cast<float*>(sb + A.o_ring_a);
    float2* snap = reinterpret_cast<float2*>(sb + A.o_snap);
    auto bufA = [&](int tile) { return dA + (size_t)((tile & 1) * G + g) * kPChainBuf; };
    auto bufD = [&](int tile, int comp) { return dD + (size_t)(((tile & (kPDS - 1)) * 2 + comp) * G + g) * kPChainBuf; };
    const float dclen_f = (float)C.dc_length;
    const int pair_bar = PBAR_PAIR + g;
    PP_DECL

    if (is_a) {
      // ------------------------------------------------------------------------------------- warps A0 / A1: P1 of one half-tile
      const int odd = (int)(sg.offset & 1ull);
      const float winlen_f = (float)C.win_length;
      const int h2 = ahalf;
#ifdef RFID_B200_PHASE_PROFILE
      if (seg == 0 && A.window_tap) pp_log = reinterpret_cast<long long*>(A.window_tap) + (h2 ? 256 * 8 : 0);
#endif
      for (int i = 0; i < nsteps; i++) {
#ifdef RFID_B200_PHASE_PROFILE
        pp_step = i;
#endif
        PP_AT(0)
        const int k = 2 * i + h2;                      // half-tile index = tile index of the 128-sample kernels
        if (k * kTT < n_out) {
          const int nvalid = min(kTT, n_out - k * kTT);  // outputs of this half-tile
          const int t0 = Q * lane;
          const float2* stage = raw + (size_t)h2 * A.raw_stage_samples;
          const int delta = -odd - (k > 0 ? DECIM - 1 : 0);
          const int base = DECIM * t0 - (DECIM - 1) - delta;
          // ---- block sums B(n) = x[D*n-D+1 .. D*n], ascending
          float2 w[MFQ - 1 + Q];
          mbar_wait(&B.raw_full[h2], (uint32_t)(i & 1));
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
          // ---- the MFQ-1 block sums before this half-tile (lane 0's halo)
          float2 halo[MFQ - 1];
          if (h2 == 0) {
            // ... are the last ones of the previous tile: warp A1 left them in the keep slot one step ago
#pragma unroll
            for (int m = 0; m < MFQ - 1; m++) halo[m] = B.keep[(i + 1) & 1][m];   // written while tile i-1 was processed
          } else {
            // ... are the last ones of the first half-tile: lanes 0..MFQ-2 recompute them from the tail of stage 0
            mbar_wait(&B.raw_full[0], (uint32_t)(i & 1));
            const float2* stage0 = raw;
            const int delta0 = -odd - (k - 1 > 0 ? DECIM - 1 : 0);
            const int tl = kTT - (MFQ - 1) + min(lane, MFQ - 2);          // output kTT-4 .. kTT-1 of half-tile 0
            const int b0 = DECIM * tl - (DECIM - 1) - delta0;
            float2 hb = stage0[b0];
#pragma unroll
            for (int j = 1; j < DECIM; j++) hb = c_add2(hb, stage0[b0 + j]);
#pragma unroll
            for (int m = 0; m < MFQ - 1; m++) halo[m] = make_float2(__shfl_sync(0xffffffffu, hb.x, m), __shfl_sync(0xffffffffu, hb.y, m));
          }
          __syncwarp();  // raw stages consumed: the loader may refill them
          if (lane == 0) { mbar_arrive(&B.raw_empty[h2]); if (h2 == 1) mbar_arrive(&B.raw_empty[0]); }
          PP_AT(1)
#pragma unroll
          for (int m = 0; m < MFQ - 1; m++) {
            const float2 mine = w[Q + m];
            const float ux = __shfl_up_sync(0xffffffffu, mine.x, 1), uy = __shfl_up_sync(0xffffffffu, mine.y, 1);
            w[m] = lane ? make_float2(ux, uy) : halo[m];
            if (h2 == 1) {  // hand the tile's last block sums to warp A0 (read after the pair barrier)
              const float ex = __shfl_sync(0xffffffffu, mine.x, 31), ey = __shfl_sync(0xffffffffu, mine.y, 31);
              if (lane == 0) B.keep[i & 1][m] = make_float2(ex, ey);
            }
          }
          // ---- y = sum of MFQ block sums (canonical order), a = |y| (gate_impl.cc:130)
          float2 y[Q];
          float a[Q];
          bool risky = false;
#pragma unroll
          for (int q = 0; q < Q; q++) {
            y[q] = w[q];
#pragma unroll
            for (int m = 1; m < MFQ; m++) y[q] = c_add2(y[q], w[q + m]);
            // outputs past the end of the segment (last, partial tile) are computed from stale shared memory: give them
            // a harmless value so that they cannot push the whole warp onto the rare exact-evaluation paths below
            if (nvalid < kTT && t0 + q >= nvalid) y[q] = make_float2(1.0f, 0.0f);
            bool rq;
            a[q] = cabsf_quick(y[q].x, y[q].y, rq);
            risky = risky || rq;
          }
          if (__any_sync(0xffffffffu, risky)) {  // rare (about one tile in 60): the exact evaluation for the whole warp
#pragma unroll
            for (int q = 0; q < Q; q++) a[q] = cabsf_ref(y[q].x, y[q].y);
          }
          const int by = (i % 3) * kT2 + h2 * kTT, ba = (i & 1) * kT2 + h2 * kTT;   // this half-tile's position in the time rings
          {
            float4* py = reinterpret_cast<float4*>(ring_y + by + t0);
            py[0] = make_float4(y[0].x, y[0].y, y[1].x, y[1].y);
            py[1] = make_float4(y[2].x, y[2].y, y[3].x, y[3].y);
            *reinterpret_cast<float4*>(ring_a + ba + t0) = make_float4(a[0], a[1], a[2], a[3]);
          }
          PP_AT(2)
          __syncwarp();  // this half-tile's |y| and y visible to the lookbacks below
          if (h2 == 0) {
            if (lane == 0 && (k + 1) * kTT < n_out) mbar_arrive(&B.half_rdy);   // A1's lookbacks reach into this half
          } else {
            mbar_wait(&B.half_rdy, (uint32_t)(i & 1));
          }
          // ---- ring differences (gate_impl.cc:131,141)
          float xd[Q], xr[Q], xi[Q];
          int ia = ba + t0 - C.win_length, iy = by + t0 - C.dc_length;
          if (ia < 0) ia += kRingA;
          if (iy < 0) iy += kRingY;
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
              if (ja >= kRingA) ja -= kRingA;
              if (jy >= kRingY) jy -= kRingY;
              const float2 old = ring_y[jy];
              xd[q] = f_sub(a[q], ring_a[ja]);
              xr[q] = f_sub(y[q].x, old.x);
              xi[q] = f_sub(y[q].y, old.y);
            }
          }
          if (nvalid < kTT) {
#pragma unroll
            for (int q = 0; q < Q; q++)
              if (t0 + q >= nvalid) { xd[q] = 1.0f; xr[q] = 1.0f; xi[q] = 1.0f; }
          }
          // all twelve divisions in flight together; the multiply-correct quotients need every dividend in the verified range
          float mx = fabsf(xd[0]), mn = mx;
#pragma unroll
          for (int q = 0; q < Q; q++) {
            mx = fmaxf(fmaxf(mx, fabsf(xd[q])), fmaxf(fabsf(xr[q]), fabsf(xi[q])));
            mn = fminf(fminf(mn, fabsf(xd[q])), fminf(fabsf(xr[q]), fabsf(xi[q])));
          }
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
            // the segment's last, partial half-tile: the chain warp runs whole groups of 16 steps, so the slots past the
            // end hold -0.0f (x + -0.0f == x for every x, including both zeros: the running sum is carried unchanged)
#pragma unroll
            for (int q = 0; q < Q; q++)
              if (t0 + q >= nvalid) { qd[q] = -0.0f; qr[q] = -0.0f; qi[q] = -0.0f; }
          }
          *reinterpret_cast<float4*>(bufA(i) + h2 * kTT + t0) = make_float4(qd[0], qd[1], qd[2], qd[3]);
          *reinterpret_cast<float4*>(bufD(i, 0) + h2 * kTT + t0) = make_float4(qr[0], qr[1], qr[2], qr[3]);
          *reinterpret_cast<float4*>(bufD(i, 1) + h2 * kTT + t0) = make_float4(qi[0], qi[1], qi[2], qi[3]);
          __syncwarp();
        }
        PP_AT(3)
        pbar_arrive<PBAR_X>(i & 1, bar_x_count);   // this half of tile i is ready for the chain warp
        pair_sync(pair_bar);                        // warp B is done with step i: ring slot and sum buffers of tile i+1 are free
      }
    } else {
      // ------------------------------------------------------------------------------------- warp B: P3
      // the gate (gate_impl.cc:45, global_vars.cc:47, reader_impl.cc:259,262)
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
      uint32_t go_count = 0;      // emission requests answered by warp C so far (phase parity of the done barrier)
      bool done_pending = false;
#ifdef RFID_B200_PHASE_PROFILE
      if (seg == 0 && A.window_tap) pp_log = reinterpret_cast<long long*>(A.window_tap) + 64 * 8;
#endif
      for (int i = 0; i < nsteps; i++) {
#ifdef RFID_B200_PHASE_PROFILE
        pp_step = i;
#endif
        PP_AT(0)
        if (i >= 1) pbar_sync<PBAR_Y>((i - 1) & 1, bar_count);  // avg_ampl of tile i-1 and dc_est of tile i-3 are final
        PP_AT(1)
        // ---- dc_est right after the trigger sample of every window that opened in tile i-3: hand it to warp C
        if (i >= 3 && i - 3 < ntiles) {
          const int t = i - 3, ps = t & (kPDS - 1);
          const int pnev = B.n_ev[ps];
          for (int e = 0; e < pnev; e++) {
            if (B.ev[ps][e].type == 1 && B.ev[ps][e].c != 0 && lane == 0) {  // (windows beyond max_windows are not decoded)
              const int j = B.ev[ps][e].a, slot = B.ev[ps][e].d;
              B.dc_val[slot] = make_float2(bufD(t, 0)[j], bufD(t, 1)[j]);
              mbar_arrive(&B.dc_rdy[slot]);
            }
          }
          __syncwarp();
        }
        // ================================================================= P3(i-1): thresholds, state machine, DC list
        if (i >= 1 && i - 1 < ntiles) {
          const int t = i - 1, s = t & (kPDS - 1);
          int nev = 0, n_e = 0;
          const bool open_at_start = gate_open;
          const int nvalid = min(kT2, n_out - t * kT2);
          const int by = (t % 3) * kT2, ba = (t & 1) * kT2;
          const float* davg = bufA(t);
          const float* ta = ring_a + ba;
          const float2* ty = ring_y + by;
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
            Mask256 lt, gt;
#pragma unroll
            for (int r = 0; r < 8; r++) { lt.w[r] = 0u; gt.w[r] = 0u; }
            bool quiet = false;
            if (sig_pos && !gate_open && nvalid == kT2) {
              bool below = false;
#pragma unroll
              for (int h2 = 0; h2 < 2; h2++) {
                const float4 av = *reinterpret_cast<const float4*>(davg + h2 * kTT + 4 * lane);
                const float4 aa = *reinterpret_cast<const float4*>(ta + h2 * kTT + 4 * lane);
                below = below || aa.x < f_mul(av.x, kThreshFraction) || aa.y < f_mul(av.y, kThreshFraction) ||
                        aa.z < f_mul(av.z, kThreshFraction) || aa.w < f_mul(av.w, kThreshFraction);
              }
              quiet = !__any_sync(0xffffffffu, below);
            }
            bool have_masks = false;
            auto make_masks = [&]() {
#pragma unroll
              for (int r = 0; r < 8; r++) {
                const int p = r * 32 + lane;
                const float thr = f_mul(davg[p], kThreshFraction);
                const float av = ta[p];
                lt.w[r] = __ballot_sync(0xffffffffu, av < thr);
                gt.w[r] = __ballot_sync(0xffffffffu, av > thr);
              }
              if (nvalid < kT2) {
#pragma unroll
                for (int r = 0; r < 8; r++) {
                  const unsigned vm = m_below(nvalid - r * 32);
                  lt.w[r] &= vm;
                  gt.w[r] &= vm;
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
                PP_AT(4)
                if (!have_masks && !(quiet && run_start == 0)) make_masks();
                PP_AT(5)
                unsigned any_lt = 0u;
#pragma unroll
                for (int r = 0; r < 8; r++) any_lt |= lt.w[r];
                if (sig_pos && any_lt == 0u) {
                  // carrier only (the common case): no falling edge can occur, only the open test remains
                  if (num_pulses > kNumPulsesCommand) {
                    const int cand = run_start + max(0, C.n_T1 - n_samples);
                    if (cand < nvalid) { p_open = cand; num_pulses = 0; n_samples = 1; }
                  }
                  if (p_open < 0) n_samples += nvalid - run_start;
                } else {
                  GateFsm2 fs = {sig_pos, n_samples, num_pulses};
                  p_open = fsm_closed_run_lanes(lt, gt, run_start, nvalid, C.n_T1, half_pw, fs);
                  sig_pos = fs.sig_pos; n_samples = fs.n_samples; num_pulses = fs.num_pulses;
                }
                PP_AT(6)
                const bool opened = p_open >= 0;
                pos = opened ? p_open + 1 : nvalid;
                // ---- DC tracker inputs of the closed run [run_start, pos) (gate_impl.cc:141-143; includes the trigger)
                const int len = pos - run_start;
                if (run_start == 0 && pos == nvalid && !opened && closed_since >= C.dc_length) {
                  // no gate activity and the ring lookback is time-contiguous: P1's differences are exact
                } else {
                  list_rebuilt = true;
#pragma unroll 1
                  for (int j0 = 0; j0 < len; j0 += 32) {
                    const int j = j0 + lane;
                    const bool valid = j < len;
                    const int p = run_start + (valid ? j : 0), m = closed_since + j;
                    const float2 yv = ty[p];
                    float2 old;
                    if (m < C.dc_length) {
                      old = snap[valid ? m : 0];  // ring contents from before the window
                    } else {
                      int iy = by + p - C.dc_length;
                      if (iy < 0) iy += kRingY;
                      old = ring_y[iy];
                    }
                    const float xr = valid ? f_sub(yv.x, old.x) : 1.0f, xi = valid ? f_sub(yv.y, old.y) : 1.0f;
                    float qr, qi;
                    if (__all_sync(0xffffffffu, C.dc_div_fast && f_div_fast_ok(xr) && f_div_fast_ok(xi))) {
                      qr = f_div_fast(xr, dclen_f, C.dc_recip);
                      qi = f_div_fast(xi, dclen_f, C.dc_recip);
                    } else {
                      qr = f_div_const(xr, dclen_f, C.dc_recip, C.dc_div_fast);
                      qi = f_div_const(xi, dclen_f, C.dc_recip, C.dc_div_fast);
                    }
                    if (valid) { er[n_e + j] = qr; ei[n_e + j] = qi; }
                  }
                }
                PP_AT(7)
                closed_since = min(closed_since + len, 1 << 24);
                n_e += len;
                if (opened) {
                  // READER COMMAND DETECTED (gate_impl.cc:164-180): keep the dc ring as it stands now
#pragma unroll 1
                  for (int j = lane; j < C.dc_length; j += 32) {
                    int iy = by + (pos - 1) - C.dc_length + 1 + j;
                    if (iy < 0) iy += kRingY;
                    snap[j] = ring_y[iy];
                  }
                  gate_open = true;
                  open_idx = t * kT2 + pos - 1;
                  cur_store = wcount < A.max_windows;
                  if (lane == 0 && nev < kMaxTileEvents) {
                    TileEvent& ev = B.ev[s][nev];
                    ev.type = 1; ev.pos = pos - 1; ev.a = n_e - 1; ev.b = open_idx; ev.c = cur_store ? 1 : 0; ev.d = wcount & 1;
                  }
                  nev++;
                }
              } else {
                // ---- open: samples pass through (gate_impl.cc:182-195); warp C copies them out one step later
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
        PP_AT(2)
        // ---- emission by warp C: tile i-2's copy (requested last step) must be out of the ring before tile i+1 is written
        if (done_pending) { mbar_wait(&B.done, go_count & 1u); go_count++; done_pending = false; }
        if (i >= 1 && i - 1 < ntiles && B.emit[(i - 1) & (kPDS - 1)] != 0) {
          if (lane == 0) { B.go_tile = i - 1; mbar_arrive(&B.go); }
          done_pending = true;
        }
        PP_AT(3)
        pair_sync(pair_bar);
      }
      // ---- end of the segment
      if (done_pending) { mbar_wait(&B.done, go_count & 1u); go_count++; }
      if (have && lane == 0) A.counts[seg] = wcount;
      if (lane == 0) { B.go_tile = -1; mbar_arrive(&B.go); }
    }
  } else if (warp == 3 * G + 1) {
    // ======================================================================================= chain warp
    // lane 8*c + g: running sum c (0 avg_ampl, 1 dc_est.re, 2 dc_est.im) of segment g
    const int comp = lane >> 3, g = lane & 7;
    const bool active = comp < 3 && g < g_act;
    int n_out = 0;
    if (active) n_out = (int)(A.segs[seg0 + g].length / DECIM);
    float acc = 0.f;
    PP_DECL
#ifdef RFID_B200_PHASE_PROFILE
    if (blockIdx.x == 0 && A.window_tap) pp_log = reinterpret_cast<long long*>(A.window_tap) + 128 * 8;
#endif
    for (int i = 0; i < nsteps; i++) {
#ifdef RFID_B200_PHASE_PROFILE
      pp_step = i;
#endif
      PP_AT(0)
      pbar_sync<PBAR_X>(i & 1, bar_x_count);
      PP_AT(1)
      {
        // branch-free selection of this lane's buffer and length, then ONE convergent loop for all 24 chains
        const int t = i - 2;
        const int n_avg = min(kT2, max(0, n_out - i * kT2));
        const int n_dc = (active && comp > 0 && i >= 2) ? ctl_all[g].n_e[t & (kPDS - 1)] : 0;
        const int n = active ? (comp == 0 ? n_avg : n_dc) : 0;
        const int ofsA = ((i & 1) * G + g) * kPChainBuf;
        const int ofsD = (((t & (kPDS - 1)) * 2 + (comp - 1)) * G + g) * kPChainBuf;
        float* buf = comp == 0 ? dA + ofsA : dD + (active && comp > 0 ? ofsD : 0);
        const int n16 = (n + 15) & ~15;
        __syncwarp();
        chain_inplace(buf, n16, acc);
      }
      __syncwarp();
      PP_AT(2)
      if (i + 1 < nsteps) pbar_arrive<PBAR_Y>(i & 1, bar_count);
    }
  } else if (warp == 3 * G) {
    // ======================================================================================= loader warp
    // lane g streams the raw half-tiles of segment g through its two stages (TMA bulk copies) as warp A frees them
    const int g = lane;
    if (g < g_act) {
      const rfid_b200_segment sg = A.segs[seg0 + g];
      const int n_out = (int)(sg.length / DECIM);
      const int nhalf = (n_out + kTT - 1) / kTT;
      PackSegCtl& B = ctl_all[g];
      float2* raw = reinterpret_cast<float2*>(smem + A.off_seg + (size_t)g * A.seg_bytes + A.o_raw);
      // raw half-tile geometry (as rx_fused_split.cuh)
      const int odd = (int)(sg.offset & 1ull);
      const uint32_t fast_bytes = (uint32_t)((DECIM * kTT + 2 * odd) * 8);
      int fast_tiles = 0;
      {
        const long long by_len = ((long long)sg.length - 1 - (long long)DECIM * (kTT - 1)) / ((long long)DECIM * kTT);
        const long long room = (long long)A.n_raw - (long long)sg.offset + (DECIM - 1) + odd - (DECIM * kTT + 2 * odd);
        const long long by_buf = room >= 0 ? room / ((long long)DECIM * kTT) : -1;
        long long f = (by_len < by_buf ? by_len : by_buf) + 1;
        if (sg.length < (unsigned)(DECIM * kTT)) f = 0;
        fast_tiles = f < 0 ? 0 : (f > nhalf ? nhalf : (int)f);
      }
      const float2* const fast_src = A.iq + sg.offset - (DECIM - 1) - odd;
      FusedArgs FA;  // issue_tile_load only reads iq / n_raw
      FA.iq = A.iq; FA.n_raw = A.n_raw;
      for (int k = 0; k < nhalf; k++) {
        const int rs_ = k & 1;
        if (k >= 2) mbar_wait_idle(&B.raw_empty[rs_], (uint32_t)(((k >> 1) - 1) & 1), 100);
        float2* dst = raw + (size_t)rs_ * A.raw_stage_samples;
        if (k >= 1 && k < fast_tiles) {
          mbar_arrive_expect_tx(&B.raw_full[rs_], fast_bytes);
          tma_load_1d(dst, fast_src + (size_t)k * (DECIM * kTT), fast_bytes, &B.raw_full[rs_]);
        } else {
          issue_tile_load<DECIM>(FA, sg, k, dst, &B.raw_full[rs_]);
        }
      }
    }
  } else {
    // ======================================================================================= warp C: emission + decode
    const int g = warp - 3 * G - 2;
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
      uint32_t n_dc[2] = {0u, 0u};   // dc_est values taken per slot (phase parity of dc_rdy)
      float2* win = win_base;
      PP_DECL
#ifdef RFID_B200_PHASE_PROFILE
      if (seg == 0 && A.window_tap) pp_log = reinterpret_cast<long long*>(A.window_tap) + 192 * 8;
#endif
      for (uint32_t rq = 0;; rq++) {
        mbar_wait_idle(&B.go, rq & 1u, 400);  // nothing to do until warp B rings (its deadline is a whole step away)
        const int t = B.go_tile;
        if (t < 0) break;
#ifdef RFID_B200_PHASE_PROFILE
        pp_step = t;
#endif
        PP_AT(0)
        const int ps = t & (kPDS - 1);
        const float2* py = ring_y + (t % 3) * kT2;
        const int pvalid = min(kT2, n_out - t * kT2);
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
              for (int j = lane; j < take; j += 32) win[f_wpos + j] = py[pos + j];
            f_wpos += take;
            pos = epos;
          }
          if (etype == 2) {
            if (f_open && f_store && n_closed < 2) { c_kind[n_closed] = cur_kind; c_ord[n_closed] = cur_ordinal; c_open[n_closed] = cur_open_idx; n_closed++; }
            f_open = false;
            pos = epos;
          } else if (etype == 1) {
            f_store = B.ev[ps][e].c != 0;
            f_open = true;
            cur_kind = B.ev[ps][e].d;
            cur_ordinal = wsig_ordinal++;
            cur_open_idx = B.ev[ps][e].b;
            win = win_base + (cur_kind ? A.rn16_pad : 0);
            if (f_store && lane == 0) win[0] = py[epos];
            f_wpos = 1;
            pos = epos + 1;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&B.done);  // the ring reads above are complete
        PP_AT(1)
        for (int k = 0; k < n_closed; k++) {
          const int kind = c_kind[k], len = kind ? C.len_epc : C.len_rn16;
          const float2* wv = win_base + (kind ? A.rn16_pad : 0);
          // dc_est of this window: posted by warp B three steps after the window opened
          mbar_wait_idle(&B.dc_rdy[kind], (kind ? n_dc[1] : n_dc[0]) & 1u, 200);
          if (kind) n_dc[1]++; else n_dc[0]++;
          const float2 dc = B.dc_val[kind];
          WindowDecode wd;
#ifdef RFID_B200_PHASE_PROFILE
          decode_window_staged(C, kind, wv, len, dstage, A.dstage_samples, wd, nullptr, nullptr, dc, pp_log ? pp_log + (kind ? 60 : 61) * 8 : nullptr);
#else
          decode_window_staged(C, kind, wv, len, dstage, A.dstage_samples, wd, nullptr, nullptr, dc);
#endif
          rfid_b200_window_result* dst = A.results + (size_t)seg * A.max_windows + c_ord[k];
          if (lane == 0) store_result(dst, wd, seg + A.seg_base, c_ord[k], c_open[k], len, kind);
#ifndef RFID_B200_PHASE_PROFILE
          if (A.window_tap) {
            float2* tap = A.window_tap + ((size_t)(seg + A.seg_base) * A.max_windows + c_ord[k]) * C.len_epc;
            for (int p2 = lane; p2 < len; p2 += 32) tap[p2] = c_sub(__ldcg(wv + p2), dc);
          }
#endif
          __syncwarp();
        }
        PP_AT(2)
      }
    }
  }
}

}  // namespace rfid_b200
