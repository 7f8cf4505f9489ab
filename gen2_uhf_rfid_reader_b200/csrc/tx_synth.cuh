// tx_synth.cuh -- the reader's TX side on the GPU:
//   * tx_synth_kernel: the PIE command waveforms of reader_impl (reader_impl.cc:51-125 waveform tables,
//     :251-372 what each Gen2 logic state emits, :383-443 CRC-5), sample for sample, at the DAC rate;
//   * sim_slot_kernel: a closed-loop inventory-slot simulator built on the same waveforms -- reader TX
//     -> zero-order hold to the ADC rate -> TX/RX edge response -> leakage + tag backscatter (FM0 at BLF
//     40 kHz, T1 after the command, per-tag gain / clock offset) + white noise.  The ACK carries the RN16
//     that the receive chain actually decoded from the first half of the slot, and only a tag whose RN16
//     matches answers with its EPC -- Query -> tag model -> decode -> ACK -> tag model -> decode.
// Everything is counter-based (Philox-4x32-10 keyed by seed and segment id), so a rank that generates only
// its shard of segments produces exactly the samples the single-process run would.
#pragma once

#include <cstdint>

#include "../../include/rfid_b200.h"
#include "rx_common.cuh"

namespace rfid_b200 {

// ------------------------------------------------------------------ reader waveform tables
// Sample counts exactly as the reference derives them: float sample period, float products, truncation
// when the vectors are sized (reader_impl.cc:51-71).
struct TxTiming {
  int n_data0, n_data1, n_pw, n_delim, n_rtcal, n_trcal, n_cw, n_cwquery, n_cwack, n_pdown;
};

struct TxSymbol {
  int len;   // samples
  int high;  // the first `high` samples are 1.0, the rest 0.0
};

constexpr int kTxMaxSymbols = 40;

__host__ __device__ inline int tx_crc5(const int* bits17)
{
  unsigned reg = 0x09;  // preset 01001; x^5 + x^3 + 1 (reader_impl.cc:383-443)
  for (int i = 0; i < 17; i++) {
    const unsigned fb = ((reg >> 4) & 1u) ^ (unsigned)(bits17[i] & 1);
    reg = (reg << 1) & 0x1Fu;
    if (fb) reg ^= 0x09;
  }
  return (int)reg;
}

// Symbols of one Gen2-logic emission (one reader_impl::general_work call).  Returns the symbol count.
__host__ __device__ inline int tx_build(const TxTiming& T, int kind, int arg, int fixed_q, TxSymbol* sym)
{
  const TxSymbol data0 = {T.n_data0, T.n_data0 / 2};          // half on, half off (:92)
  const TxSymbol data1 = {T.n_data1, 3 * T.n_data1 / 4};      // three quarters on (:93)
  const TxSymbol delim = {T.n_delim, 0};
  const TxSymbol rtcal = {T.n_rtcal, T.n_rtcal - T.n_pw};
  const TxSymbol trcal = {T.n_trcal, T.n_trcal - T.n_pw};
  int n = 0;
  auto frame_sync = [&]() { sym[n++] = delim; sym[n++] = data0; sym[n++] = rtcal; };
  auto bit = [&](int b) { sym[n++] = b ? data1 : data0; };
  switch (kind) {
    case RFID_B200_TX_START:  // power the tag before the first Query (:237-243)
    case RFID_B200_TX_CW:     // carrier during the EPC reply (:322-328)
      sym[n++] = TxSymbol{T.n_cwack, T.n_cwack};
      break;
    case RFID_B200_TX_POWER_DOWN:
      sym[n++] = TxSymbol{T.n_pdown, 0};
      break;
    case RFID_B200_TX_NAK: {  // frame-sync + 11000000 + cw (:245-263)
      frame_sync();
      const int code[8] = {1, 1, 0, 0, 0, 0, 0, 0};
      for (int i = 0; i < 8; i++) bit(code[i]);
      sym[n++] = TxSymbol{T.n_cw, T.n_cw};
      break;
    }
    case RFID_B200_TX_QUERY: {  // preamble + 1000|DR|M|TRext|Sel|Session|Target|Q|CRC-5 + cw (:265-285)
      frame_sync();
      sym[n++] = trcal;
      int b[22] = {1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int i = 0; i < 4; i++) b[13 + i] = (fixed_q >> (3 - i)) & 1;
      const int crc = tx_crc5(b);
      for (int k = 0; k < 5; k++) b[17 + k] = (crc >> (4 - k)) & 1;
      for (int i = 0; i < 22; i++) bit(b[i]);
      sym[n++] = TxSymbol{T.n_cwquery, T.n_cwquery};
      break;
    }
    case RFID_B200_TX_QUERY_REP:  // frame-sync + 00 + session 00 + cw (:330-344)
      frame_sync();
      for (int i = 0; i < 4; i++) bit(0);
      sym[n++] = TxSymbol{T.n_cwquery, T.n_cwquery};
      break;
    case RFID_B200_TX_QUERY_ADJUST: {  // frame-sync + 1001 + session + UpDn(unchanged) + cw (:346-366)
      frame_sync();
      const int code[9] = {1, 0, 0, 1, 0, 0, 0, 0, 0};
      for (int i = 0; i < 9; i++) bit(code[i]);
      sym[n++] = TxSymbol{T.n_cwquery, T.n_cwquery};
      break;
    }
    case RFID_B200_TX_ACK:  // frame-sync + 01 + RN16, no carrier of its own (:290-320)
      frame_sync();
      bit(0);
      bit(1);
      for (int i = 15; i >= 0; i--) bit((arg >> i) & 1);
      break;
    default:
      break;
  }
  return n;
}

__host__ __device__ inline long long tx_length(const TxTiming& T, int kind, int arg, int fixed_q)
{
  TxSymbol sym[kTxMaxSymbols];
  const int n = tx_build(T, kind, arg, fixed_q, sym);
  long long t = 0;
  for (int i = 0; i < n; i++) t += sym[i].len;
  return t;
}

// One CTA per emission; offsets[c] = first output sample of emission c.
__global__ void __launch_bounds__(256) tx_synth_kernel(TxTiming T, int fixed_q, const rfid_b200_tx_command* __restrict__ script,
                                                       const unsigned long long* __restrict__ offsets, int n_cmd,
                                                       float* __restrict__ out)
{
  __shared__ TxSymbol s_sym[kTxMaxSymbols];
  __shared__ int s_n;
  const int c = blockIdx.x;
  if (c >= n_cmd) return;
  if (threadIdx.x == 0) s_n = tx_build(T, script[c].kind, script[c].arg, fixed_q, s_sym);
  __syncthreads();
  float* o = out + offsets[c];
  for (int s = 0; s < s_n; s++) {
    const TxSymbol y = s_sym[s];
    for (int i = threadIdx.x; i < y.len; i += blockDim.x) o[i] = i < y.high ? 1.0f : 0.0f;
    o += y.len;
  }
}

// ------------------------------------------------------------------ counter-based randomness
struct Philox4 {
  unsigned int x, y, z, w;
};

__device__ __forceinline__ Philox4 philox4x32_10(unsigned int c0, unsigned int c1, unsigned int c2, unsigned int c3,
                                                 unsigned int k0, unsigned int k1)
{
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const unsigned int hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned int hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const unsigned int n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}

__device__ __forceinline__ float u01(unsigned int v) { return ((float)(v >> 8) + 0.5f) * (1.0f / 16777216.0f); }  // (0,1)

enum { kStreamRound = 1, kStreamSegment = 2, kStreamEpc = 3, kStreamNoise = 4 };

// ------------------------------------------------------------------ slot simulator
constexpr int kSimMaxTags = 16;
constexpr int kSimMaxFir = 64;
constexpr int kSimThreads = 256;
constexpr int kSimMaxSymbols = 2 * kTxMaxSymbols + 4;
constexpr int kRn16Halves = 12 + 2 * 17;   // preamble + 16 bits + dummy
constexpr int kEpcHalves = 12 + 2 * 129;   // preamble + 128 bits + dummy

struct SimArgs {
  rfid_b200_sim_params p;
  TxTiming T;             // DAC-rate waveform tables
  int fixed_q;
  int hold;               // ADC samples per DAC sample (zero-order hold)
  int seg_len;            // ADC samples per segment
  int lead_dac;           // DAC samples of carrier before command 1
  double sps;             // ADC samples per microsecond
  double dac_us;          // microseconds per DAC sample
  int n_fir;
  float fir[kSimMaxFir];  // edge response at the ADC rate
  long long first_segment;
  int nseg;
  int phase;              // 0: samples before the ACK; 1: from the ACK on (closed loop); 2: whole slot, open loop
  float2* iq;             // nseg * seg_len
  rfid_b200_segment* segs;
  const rfid_b200_window_result* rn16_records;  // phase 1: one record slot per segment (the decoded RN16 window)
  const int32_t* rn16_counts;
  rfid_b200_sim_truth* truth;
  int mask_words;
};

struct SimShared {
  TxSymbol sym[kSimMaxSymbols];
  int off[kSimMaxSymbols];    // first DAC sample of each symbol
  int n_sym;
  int ack_start_dac;          // first DAC sample of the ACK
  int cmd1_end_dac;           // end of command 1's last PIE symbol (= its last rising edge)
  int cmd2_end_dac;
  int n_present, strongest, replier, ack_rn16;
  float g_re[kSimMaxTags], g_im[kSimMaxTags];
  float hk[kSimMaxTags];      // ADC samples per half symbol of tag k
  float jit[kSimMaxTags];     // extra reply delay of tag k, us
  float t0_rn16[kSimMaxTags]; // ADC sample at which tag k's RN16 reply starts
  float t0_epc;
  int present[kSimMaxTags];
  int rn16[kSimMaxTags];
  unsigned char epc[16];
  unsigned char lv_rn16[kSimMaxTags][kRn16Halves];
  unsigned char lv_epc[kEpcHalves];
};

__device__ inline void fm0_levels(const unsigned char* bits, int nbits, unsigned char* lv)
{
  // 12 preamble halves (global_vars.h:136), then data + dummy '1' (global_vars.h:104-107): a symbol starts
  // with the inverse of the level before it; bit 1 holds that level, bit 0 inverts again mid-symbol
  const unsigned char pre[12] = {1, 1, 0, 1, 0, 0, 1, 0, 0, 0, 1, 1};
  for (int i = 0; i < 12; i++) lv[i] = pre[i];
  int last = 1;
  for (int j = 0; j <= nbits; j++) {
    const int b = j < nbits ? bits[j] : 1;
    const int first = 1 - last;
    const int second = b ? first : 1 - first;
    lv[12 + 2 * j] = (unsigned char)first;
    lv[12 + 2 * j + 1] = (unsigned char)second;
    last = second;
  }
}

__device__ inline unsigned int crc16_gen2_bytes(const unsigned char* d, int n)
{
  unsigned int crc = 0xFFFF;
  for (int i = 0; i < n; i++) {
    crc ^= (unsigned int)d[i] << 8;
    for (int k = 0; k < 8; k++) crc = (crc & 0x8000u) ? ((crc << 1) ^ 0x1021u) & 0xFFFFu : (crc << 1) & 0xFFFFu;
  }
  return (~crc) & 0xFFFFu;
}

// PC 0x3000 + 96-bit EPC of tag k (last byte 0x27 + k: the recording's tag is 0x27) + CRC-16
__device__ inline void sim_epc_frame(unsigned long long seed, int k, unsigned char* f)
{
  const Philox4 a = philox4x32_10((unsigned)k, 0u, 0u, kStreamEpc, (unsigned)seed, (unsigned)(seed >> 32));
  const Philox4 b = philox4x32_10((unsigned)k, 0u, 1u, kStreamEpc, (unsigned)seed, (unsigned)(seed >> 32));
  const unsigned int w[3] = {a.x, a.y, b.x};
  f[0] = 0x30; f[1] = 0x00;
  for (int i = 0; i < 12; i++) f[2 + i] = (unsigned char)(w[i >> 2] >> (8 * (i & 3)));
  f[13] = (unsigned char)(0x27 + k);
  const unsigned int c = crc16_gen2_bytes(f, 14);
  f[14] = (unsigned char)(c >> 8);
  f[15] = (unsigned char)(c & 0xFF);
}

__global__ void __launch_bounds__(kSimThreads) sim_slot_kernel(SimArgs A)
{
  extern __shared__ __align__(16) unsigned char sim_smem[];
  SimShared& S = *reinterpret_cast<SimShared*>(sim_smem);
  unsigned int* low = reinterpret_cast<unsigned int*>(sim_smem + ((sizeof(SimShared) + 15) & ~(size_t)15));  // 1 bit / ADC sample
  const int tid = threadIdx.x;
  const int s = blockIdx.x;
  if (s >= A.nseg) return;
  const long long id = A.first_segment + s;
  const unsigned int k0 = (unsigned int)A.p.seed, k1 = (unsigned int)(A.p.seed >> 32);
  const int slots = 1 << A.fixed_q;
  const long long round = id / slots;
  const int slot = (int)(id % slots);
  const int K = A.p.n_tags < kSimMaxTags ? (A.p.n_tags > 0 ? A.p.n_tags : 0) : kSimMaxTags;

  for (int w = tid; w < A.mask_words; w += kSimThreads) low[w] = 0u;
  if (tid == 0) {
    // ---- who answers in this slot, with which RN16, gain and clock ----
    int n_present = 0, strongest = -1;
    float best = -1.f;
    for (int k = 0; k < K; k++) {
      const Philox4 r = philox4x32_10((unsigned)round, (unsigned)(round >> 32), (unsigned)k, kStreamRound, k0, k1);
      const Philox4 q = philox4x32_10((unsigned)id, (unsigned)(id >> 32), (unsigned)k, kStreamSegment, k0, k1);
      const float amp = A.p.tag_gain * (0.6f + 0.8f * u01(r.y));
      const float ph = A.p.tag_phase + (u01(r.z) - 0.5f) * 1.2f;
      S.present[k] = (int)(r.x % (unsigned)slots) == slot;
      S.rn16[k] = (int)(q.x & 0xFFFFu);
      S.g_re[k] = amp * cosf(ph);
      S.g_im[k] = amp * sinf(ph);
      S.hk[k] = (float)(12.5 * A.sps) * (1.f + (2.f * u01(q.y) - 1.f) * A.p.clock_pct * 0.01f);
      S.jit[k] = 2.f * u01(q.z);
      if (S.present[k]) {
        n_present++;
        if (amp > best) { best = amp; strongest = k; }
      }
    }
    S.n_present = n_present;
    S.strongest = strongest;
    // ---- reader timeline at the DAC rate: carrier, command 1 + its carrier, ACK (then carrier to the end) ----
    int n = 0, t = 0;
    S.sym[n] = TxSymbol{A.lead_dac, A.lead_dac}; S.off[n++] = 0;
    t = A.lead_dac;
    const int n1 = tx_build(A.T, slot == 0 ? RFID_B200_TX_QUERY : RFID_B200_TX_QUERY_REP, 0, A.fixed_q, S.sym + n);
    for (int i = 0; i < n1; i++) {
      if (i == n1 - 1) S.cmd1_end_dac = t;  // the last symbol is the carrier after the command
      S.off[n + i] = t;
      t += S.sym[n + i].len;
    }
    n += n1;
    S.ack_start_dac = t;
    // which RN16 does the reader acknowledge?
    int ack = 0;
    if (A.phase == 1) {
      const rfid_b200_window_result& r = A.rn16_records[s];
      ack = (A.rn16_counts[s] >= 1 && r.kind == RFID_B200_RN16) ? (r.tag_id & 0xFFFF) : -1;
    } else if (A.phase == 2) {
      ack = strongest >= 0 ? S.rn16[strongest] : 0;
    }
    S.ack_rn16 = ack;
    int replier = -1;
    if (A.phase == 2) {
      replier = strongest;
    } else if (A.phase == 1 && ack >= 0) {
      float bp = -1.f;
      for (int k = 0; k < K; k++) {
        const float pw = S.g_re[k] * S.g_re[k] + S.g_im[k] * S.g_im[k];
        if (S.present[k] && S.rn16[k] == ack && pw > bp) { bp = pw; replier = k; }
      }
    }
    S.replier = replier;
    if (ack >= 0) {  // a reader that saw no RN16 window sends no ACK (reader_impl.cc:292)
      const int n2 = tx_build(A.T, RFID_B200_TX_ACK, ack, A.fixed_q, S.sym + n);
      for (int i = 0; i < n2; i++) { S.off[n + i] = t; t += S.sym[n + i].len; }
      n += n2;
    }
    S.cmd2_end_dac = t;
    S.n_sym = n;
    // tag reply start = T1 (248 us measured on the recording) + the tag's own delay after the command's last edge
    for (int k = 0; k < K; k++) S.t0_rn16[k] = (float)(((double)S.cmd1_end_dac * A.dac_us + 248.0 + (double)S.jit[k]) * A.sps);
    S.t0_epc = replier >= 0 ? (float)(((double)S.cmd2_end_dac * A.dac_us + 248.0 + (double)S.jit[replier]) * A.sps) : 0.f;
    for (int i = 0; i < 16; i++) S.epc[i] = 0;
    if (replier >= 0) {
      sim_epc_frame(A.p.seed, replier, S.epc);
      unsigned char bits[128];
      for (int i = 0; i < 128; i++) bits[i] = (S.epc[i >> 3] >> (7 - (i & 7))) & 1;
      fm0_levels(bits, 128, S.lv_epc);
    }
  }
  __syncthreads();
  // ---- low-pulse mask at the ADC rate (zero-order hold of the DAC waveform) ----
  for (int y = tid; y < S.n_sym; y += kSimThreads) {
    const long long a = (long long)(S.off[y] + S.sym[y].high) * A.hold;
    long long b = (long long)(S.off[y] + S.sym[y].len) * A.hold;
    if (b > A.seg_len) b = A.seg_len;
    for (long long m = a; m < b; m++) atomicOr(&low[m >> 5], 1u << (m & 31));
  }
  if (tid >= 32 && tid < 32 + K && S.present[tid - 32]) {
    unsigned char bits[16];
    const int k = tid - 32;
    for (int i = 0; i < 16; i++) bits[i] = (S.rn16[k] >> (15 - i)) & 1;
    fm0_levels(bits, 16, S.lv_rn16[k]);
  }
  __syncthreads();

  const int ack_adc = min(A.seg_len, S.ack_start_dac * A.hold);
  const int n_lo = A.phase == 1 ? ack_adc : 0;
  const int n_hi = A.phase == 0 ? ack_adc : A.seg_len;
  float2* out = A.iq + (size_t)s * A.seg_len;
  const float floor_level = A.p.floor_level;
  for (int n = n_lo + tid; n < n_hi; n += kSimThreads) {
    // TX/RX edge response: env[n] = sum_k fir[k] * ideal[n + 1 - k]
    float env = 0.f;
    for (int k = 0; k < A.n_fir; k++) {
      const int m = n + 1 - k;
      const bool lo = m >= 0 && m < A.seg_len && ((low[m >> 5] >> (m & 31)) & 1u);
      env += A.fir[k] * (lo ? floor_level : 1.0f);
    }
    float rr = A.p.leak_re, ri = A.p.leak_im;
    for (int k = 0; k < K; k++) {
      if (!S.present[k]) continue;
      const float h = floorf(((float)n - S.t0_rn16[k]) / S.hk[k]);
      if (h >= 0.f && h < (float)kRn16Halves && S.lv_rn16[k][(int)h]) { rr += S.g_re[k]; ri += S.g_im[k]; }
    }
    if (S.replier >= 0) {
      const int k = S.replier;
      const float h = floorf(((float)n - S.t0_epc) / S.hk[k]);
      if (h >= 0.f && h < (float)kEpcHalves && S.lv_epc[(int)h]) { rr += S.g_re[k]; ri += S.g_im[k]; }
    }
    const Philox4 z = philox4x32_10((unsigned)id, (unsigned)(id >> 32), (unsigned)n, kStreamNoise, k0, k1);
    const float mag = A.p.noise_sigma * sqrtf(-2.0f * logf(u01(z.x)));
    float sn, cs;
    sincosf(6.283185307179586f * u01(z.y), &sn, &cs);
    out[n] = make_float2(rr * env + mag * cs, ri * env + mag * sn);
  }
  if (tid == 0) {
    A.segs[s].offset = (unsigned long long)s * (unsigned long long)A.seg_len;
    A.segs[s].length = (unsigned int)(A.phase == 0 ? ack_adc : A.seg_len);
    A.segs[s].reserved = 0;
    if (A.truth && A.phase != 0) {
      rfid_b200_sim_truth& t = A.truth[s];
      t.is_query = slot == 0;
      t.n_replies = S.n_present;
      t.strongest_rn16 = S.strongest >= 0 ? S.rn16[S.strongest] : -1;
      t.acked_rn16 = S.ack_rn16;
      t.replier = S.replier;
      t.reserved[0] = t.reserved[1] = t.reserved[2] = 0;
      for (int i = 0; i < 16; i++) t.epc[i] = S.epc[i];
    }
  }
}

}  // namespace rfid_b200
