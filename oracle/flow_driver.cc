/* TEST INFRASTRUCTURE -- deterministic flowgraph driver for the gr::rfid blocks.
 *
 * This file is the scheduler half of the oracle.  It is compiled twice:
 *
 *   -DDRIVE_REFERENCE : together with the reference's own, UNMODIFIED block
 *       sources taken where they lie under /root/reference/gr-rfid/lib
 *       (global_vars.cc gate_impl.cc tag_decoder_impl.cc reader_impl.cc) against
 *       oracle/shim  ->  oracle/_ref/libgen2ref[_q4].so           (build_ref.sh)
 *   (default)         : together with THIS repo's thin host blocks
 *       (gen2_uhf_rfid_reader_b200/blocks/*.cc, which marshal into the CUDA
 *       library through the C-ABI)  ->  libgen2flow_b200.so
 *
 * so the same scheduler drives either implementation ("A/B under one
 * scheduler").  It wires the offline graph of
 * apps/reader.py:101-112:
 *     file_source -> fir_filter_ccc(5,[1]*25) -> gate -> tag_decoder:0 -> reader -> (TX samples)
 * with a deterministic round-robin schedule: run `reader`
 * until idle, then repeatedly: one gate.general_work() call on the next chunk;
 * hand its output to tag_decoder; whenever the decoder consumes a window, route
 * its port-0 floats to `reader` and run the reader until its state stops
 * changing.  The output is independent of the chunk size (tests check that).
 *
 * Nothing in the product path links or calls this file.
 */
#include <chrono>
#include <cstdint>
#include <cstdio>

#ifdef DRIVE_REFERENCE
/* white-box access to the reference's private members (h_est, T_global,
 * char_bits, tag_sync) -- std headers first so libstdc++ is not affected */
#include <gnuradio/attributes.h>
#include <gnuradio/block.h>
#define private public
#include "gate_impl.h"
#include "reader_impl.h"
#include "tag_decoder_impl.h"
#undef private
#else
#include "gate_impl.h"
#include "matched_filter_impl.h"
#include "reader_impl.h"
#include "tag_decoder_impl.h"
#endif

#include "../include/rfid_b200.h"
#include "mf_canonical.h"

gr::shim_logger gr::block::s_info = {false, "[info]"};
gr::shim_logger gr::block::s_debug = {false, "[debug]"};

using namespace gr::rfid;

namespace {

struct flow {
  gate::sptr G;
  tag_decoder::sptr D;
  reader::sptr R;
  std::vector<float> tx;  /* reader output (TX envelope, reader_impl.cc:44-46) */
  bool keep_tx;
  std::vector<float> rbuf;

  flow(int fs_dec, int dac_rate, bool keep) : keep_tx(keep)
  {
    /* construction order of apps/reader.py:76-78; gate's ctor allocates reader_state */
    G = gate::make(fs_dec);
    D = tag_decoder::make(fs_dec);
    R = reader::make(fs_dec, dac_rate);
    rbuf.resize(1 << 16);
  }

  /* run the Gen2 logic until it has nothing more to say */
  void run_reader(const float* in, int n_in)
  {
    for (int guard = 0; guard < 16; guard++) {
      GEN2_LOGIC_STATUS before = reader_state->gen2_logic_status;
      if (before == IDLE) break;
      gr_vector_int ni(1, n_in);
      gr_vector_const_void_star iv(1, (const void*)in);
      gr_vector_void_star ov(1, (void*)rbuf.data());
      R->shim_reset_counts();
      int w = R->general_work((int)rbuf.size(), ni, iv, ov);
      if (keep_tx && w > 0) tx.insert(tx.end(), rbuf.begin(), rbuf.begin() + w);
      n_in = 0; /* reader consumed its input (reader_impl.cc:214,378) */
      if (reader_state->gen2_logic_status == before) break;
    }
  }
};

inline void pack_bits(const float* b, int n, uint8_t* out)
{
  memset(out, 0, 16);
  for (int i = 0; i < n; i++)
    if (b[i] != 0.0f) out[i >> 3] |= (uint8_t)(0x80 >> (i & 7));
}

/* Decode one decimated stream y[0..n) with freshly constructed blocks. */
int run_stream(const gr_complex* y, size_t n, int fs_dec, int dac_rate, int chunk, int segment,
               rfid_b200_window_result* recs, int max_recs, rfid_b200_stats* stats, std::vector<float>* tx_out,
               std::string* results_text)
{
  flow f(fs_dec, dac_rate, tx_out != nullptr);
  std::vector<gr_complex> gout((size_t)chunk + 16);
  std::vector<gr_complex> dq; /* gate -> decoder queue */
  std::vector<float> dout0(4096);
  std::vector<gr_complex> dout1(4096);
  int nrec = 0;

  f.run_reader(nullptr, 0); /* START -> SEND_QUERY -> IDLE */

  size_t pos = 0;
  while (pos < n) {
    int nin = (int)std::min((size_t)chunk, n - pos);
    gr_vector_int ni(1, nin);
    gr_vector_const_void_star iv(1, (const void*)(y + pos));
    gr_vector_void_star ov(1, (void*)gout.data());
    f.G->shim_reset_counts();
    int written = f.G->general_work(nin, ni, iv, ov);
    int consumed = f.G->shim_consumed();
    pos += (size_t)consumed;
    if (written > 0) dq.insert(dq.end(), gout.begin(), gout.begin() + written);
    if (consumed == 0 && written == 0) break; /* cannot happen; guards against a stuck block */

    /* decoder fires once a complete window is queued (tag_decoder_impl.cc:223,291) */
    while (!dq.empty() && (int)dq.size() >= reader_state->n_samples_to_ungate) {
      int need = reader_state->n_samples_to_ungate;
      int kind = reader_state->decoder_status == DECODER_DECODE_RN16 ? RFID_B200_RN16 : RFID_B200_EPC;
      int epc_before = reader_state->reader_stats.n_epc_correct;

      rfid_b200_window_result r;
      memset(&r, 0, sizeof(r));
      r.segment = segment;
      r.window = nrec;
      r.open_index = (int32_t)(pos - (size_t)need);
      r.length = need;
      r.kind = kind;
#ifdef DRIVE_REFERENCE
      /* tag_sync is idempotent (it only writes h_est): call it once ourselves to
       * learn the index the reference keeps in a local (tag_decoder_impl.cc:225,298).
       * Timing runs (no record buffer) skip this harness work: only the blocks' own calls are timed. */
      tag_decoder_impl* Di = dynamic_cast<tag_decoder_impl*>(f.D.get());
      const bool lean = recs == nullptr;
      int shifted = lean ? 0 : Di->tag_sync(dq.data(), (int)dq.size());
      int half = (int)(Di->n_samples_TAG_BIT / 2);
      int m = shifted - (int)(TAG_PREAMBLE_BITS * Di->n_samples_TAG_BIT + Di->n_samples_TAG_BIT / 2);
      r.sync_index = m;
      if (!lean) {
        /* score = |sum of the six preamble-high taps|^2 at the chosen offset: the
         * value of the reference's local `max` (tag_decoder_impl.cc:89-98) -- zero
         * weights contribute exact zeros, so the running sum equals this one */
        const int J[6] = {0, 1, 3, 6, 10, 11};
        float cr = 0.0f, ci = 0.0f;
        for (int k = 0; k < 6; k++) {
          gr_complex s = dq[(size_t)(int)(m + J[k] * Di->n_samples_TAG_BIT / 2)];
          cr = cr + s.real();
          ci = ci + s.imag();
        }
        (void)half;
        r.score = cr * cr + ci * ci;
      }
#endif
      gr_vector_int dni(1, (int)dq.size());
      gr_vector_const_void_star div(1, (const void*)dq.data());
      gr_vector_void_star dov(2);
      dov[0] = dout0.data();
      dov[1] = dout1.data();
      f.D->shim_reset_counts();
      f.D->general_work((int)dout0.size(), dni, div, dov);
      int dcons = f.D->shim_consumed();
      int dprod = f.D->shim_produced(0);
      if (dcons <= 0) break;

#ifdef DRIVE_REFERENCE
      r.h_re = Di->h_est.real();
      r.h_im = Di->h_est.imag();
      if (lean) {
        /* nothing to record */
      } else if (kind == RFID_B200_RN16) {
        r.T = 0.0f;
        r.crc_ok = -1;
        pack_bits(dout0.data(), dprod < 16 ? dprod : 16, r.bits);
        r.tag_id = ((int)r.bits[0] << 8) | r.bits[1];
      } else {
        r.T = Di->T_global;
        float fb[128];
        for (int i = 0; i < 128; i++) fb[i] = Di->char_bits[i] == '1' ? 1.0f : 0.0f;
        pack_bits(fb, 128, r.bits);
        r.crc_ok = reader_state->reader_stats.n_epc_correct > epc_before ? 1 : 0;
        r.tag_id = r.bits[13];
      }
#else
      {
        const rfid_b200_window_result* lr = dynamic_cast<tag_decoder_impl*>(f.D.get())->last_result();
        int32_t seg = r.segment, win = r.window, oi = r.open_index;
        r = *lr;
        r.segment = seg; r.window = win; r.open_index = oi;
        (void)epc_before;
      }
#endif
      if (nrec < max_recs) recs[nrec] = r;
      nrec++;
      dq.erase(dq.begin(), dq.begin() + dcons);
      f.run_reader(dout0.data(), dprod);
    }
  }

  if (stats) {
    memset(stats, 0, sizeof(*stats));
    const READER_STATS& s = reader_state->reader_stats;
    stats->n_queries_sent = s.n_queries_sent;
    stats->cur_inventory_round = s.cur_inventory_round;
    stats->cur_slot_number = s.cur_slot_number;
    stats->max_slot_number = s.max_slot_number;
    stats->n_epc_correct = s.n_epc_correct;
    stats->n_windows = nrec;
    stats->terminated = reader_state->status == TERMINATED ? 1 : 0;
    int k = 0;
    for (std::map<int, int>::const_iterator it = s.tag_reads.begin(); it != s.tag_reads.end(); ++it) {
      if (k < RFID_B200_MAX_TAGS) { stats->tag_id[k] = it->first; stats->tag_reads[k] = it->second; }
      k++;
    }
    stats->n_unique_tags = k;
  }
  if (results_text) {
    /* reader::print_results() writes to std::cout (reader_impl.cc:173-192): capture it */
    std::ostringstream cap;
    std::streambuf* old = std::cout.rdbuf(cap.rdbuf());
    f.R->print_results();
    std::cout.rdbuf(old);
    *results_text = cap.str();
  }
  if (tx_out) tx_out->swap(f.tx);
  delete reader_state; /* the reference leaks it on re-init (global_vars.cc:36) */
  reader_state = nullptr;
  return nrec;
}

}  // namespace

extern "C" {

/* Continuous stream: raw I/Q -> canonical MF -> blocks.  Returns number of
 * windows (may exceed max_recs; only max_recs are stored) or <0 on error.
 * results_text: caller buffer receiving print_results() output (may be NULL).
 * tx: caller buffer for the reader's TX envelope (may be NULL). */
__attribute__((visibility("default"))) int gen2flow_run_stream(
    const float* iq_raw, size_t n_raw, int adc_rate, int decim, int ntaps, int dac_rate, int chunk,
    rfid_b200_window_result* recs, int max_recs, rfid_b200_stats* stats, char* results_text, size_t results_cap,
    float* tx, size_t tx_cap, size_t* tx_n, float* y_out)
{
  if (!iq_raw || decim <= 0 || ntaps <= 0 || chunk <= 0) return -1;
  std::vector<gr_complex> y(n_raw / (size_t)decim + 2);
  size_t ny = 0;
#ifdef DRIVE_REFERENCE
  ny = oracle_mf_boxcar(iq_raw, n_raw, ntaps, decim, (float*)y.data());
#else
  {
    /* this repo's matched-filter host block (GPU) stands where fir_filter_ccc stands in apps/reader.py:75;
     * it is fed in scheduler-sized chunks like the other blocks */
    matched_filter::sptr M = matched_filter::make(decim, ntaps);
    size_t pos_in = 0;
    const int in_chunk = chunk * decim;
    while (pos_in < n_raw) {
      int nin = (int)std::min((size_t)in_chunk, n_raw - pos_in);
      if (nin < decim) break; /* an incomplete decimation group at the very end produces nothing */
      gr_vector_int ni(1, nin);
      gr_vector_const_void_star iv(1, (const void*)(iq_raw + 2 * pos_in));
      gr_vector_void_star ov(1, (void*)(y.data() + ny));
      M->shim_reset_counts();
      int w = M->general_work(chunk + 1, ni, iv, ov);
      pos_in += (size_t)M->shim_consumed();
      ny += (size_t)w;
    }
  }
#endif
  if (y_out) memcpy(y_out, y.data(), ny * sizeof(gr_complex));
  std::vector<float> txv;
  std::string text;
  int n = run_stream(y.data(), ny, adc_rate / decim, dac_rate, chunk, 0, recs, max_recs, stats,
                     tx ? &txv : nullptr, results_text ? &text : nullptr);
  if (results_text && results_cap) {
    size_t c = std::min(results_cap - 1, text.size());
    memcpy(results_text, text.data(), c);
    results_text[c] = 0;
  }
  if (tx && tx_n) {
    size_t c = std::min(tx_cap, txv.size());
    memcpy(tx, txv.data(), c * sizeof(float));
    *tx_n = txv.size();
  }
  return n;
}

/* Already-decimated stream (block-mode parity: the flowgraph's own MF ran upstream). */
__attribute__((visibility("default"))) int gen2flow_run_decimated(
    const float* y, size_t ny, int fs_dec, int dac_rate, int chunk, rfid_b200_window_result* recs, int max_recs,
    rfid_b200_stats* stats)
{
  return run_stream((const gr_complex*)y, ny, fs_dec, dac_rate, chunk, 0, recs, max_recs, stats, nullptr, nullptr);
}

/* Independent segments, fresh blocks per segment.  Record k of
 * segment s -> recs[s*max_per_seg + k]; counts[s] = windows found.  *seconds
 * (may be NULL) = steady_clock time of MF + gate + decoder + reader over all
 * segments (no I/O, no allocation of the input). */
__attribute__((visibility("default"))) int gen2flow_run_segments(
    const float* iq_raw, const rfid_b200_segment* segs, int nseg, int adc_rate, int decim, int ntaps, int dac_rate,
    int chunk, rfid_b200_window_result* recs, int max_per_seg, int32_t* counts, double* seconds)
{
  std::vector<gr_complex> y;
  std::vector<float> scratch;
  auto t0 = std::chrono::steady_clock::now();
  for (int s = 0; s < nseg; s++) {
    size_t n_raw = segs[s].length;
    y.resize(n_raw / (size_t)decim + 1);
    scratch.resize(2 * (n_raw / (size_t)decim + (size_t)(ntaps / decim) + 2));
    /* canonical order with every block sum formed once (bit-identical to oracle_mf_boxcar, tests/test_oracle.py) */
    size_t ny = oracle_mf_boxcar_blocked(iq_raw + 2 * segs[s].offset, n_raw, ntaps, decim, (float*)y.data(), scratch.data());
    int n = run_stream(y.data(), ny, adc_rate / decim, dac_rate, chunk, s,
                       recs ? recs + (size_t)s * max_per_seg : nullptr, recs ? max_per_seg : 0, nullptr, nullptr,
                       nullptr);
    if (counts) counts[s] = n;
  }
  auto t1 = std::chrono::steady_clock::now();
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  return 0;
}

/* The reader block alone (no gate / decoder, hence no GPU): drive its state machine through a scripted
 * inventory -- START, then per round: Query or QueryRep (alternating), ACK with the given 16 RN16 bits,
 * CW -- and return the TX envelope.  Lets the CPU suite compare this repo's reader block with the
 * reference's sample for sample. */
__attribute__((visibility("default"))) int gen2flow_reader_script(const float* rn16_bits, int n_rounds, int dac_rate,
                                                                   float* tx, size_t tx_cap, size_t* tx_n)
{
  initialize_reader_state();
  reader::sptr R = reader::make(400000, dac_rate);
  std::vector<float> buf(1 << 16), all;
  auto step = [&](const float* in, int n_in) {
    gr_vector_int ni(1, n_in);
    gr_vector_const_void_star iv(1, (const void*)in);
    gr_vector_void_star ov(1, (void*)buf.data());
    int w = R->general_work((int)buf.size(), ni, iv, ov);
    if (w > 0) all.insert(all.end(), buf.begin(), buf.begin() + w);
  };
  step(nullptr, 0); /* START */
  for (int r = 0; r < n_rounds; r++) {
    if (r > 0) reader_state->gen2_logic_status = (r & 1) ? SEND_QUERY_REP : SEND_QUERY;
    step(nullptr, 0); /* Query / QueryRep */
    step(nullptr, 0); /* IDLE: nothing */
    reader_state->gen2_logic_status = SEND_ACK;
    step(rn16_bits + 16 * r, 7);  /* incomplete RN16: must not answer */
    step(rn16_bits + 16 * r, 16); /* ACK */
    step(nullptr, 0);             /* CW */
  }
  int nq = reader_state->reader_stats.n_queries_sent;
  size_t c = std::min(tx_cap, all.size());
  memcpy(tx, all.data(), c * sizeof(float));
  *tx_n = all.size();
  delete reader_state;
  reader_state = nullptr;
  return nq;
}

__attribute__((visibility("default"))) void gen2flow_set_logging(int info, int debug)
{
  gr::block::s_info.enabled = info != 0;
  gr::block::s_debug.enabled = debug != 0;
}

/* which build is this: 1 = reference blocks, 0 = this repo's host blocks */
__attribute__((visibility("default"))) int gen2flow_is_reference(void)
{
#ifdef DRIVE_REFERENCE
  return 1;
#else
  return 0;
#endif
}

__attribute__((visibility("default"))) int gen2flow_fixed_q(void) { return FIXED_Q; }

}  /* extern "C" */
