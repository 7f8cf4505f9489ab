/* reader_impl.cc -- Gen2 reader logic and command waveforms, written for this build.
 * Behavioural contract (checked against the reference's committed TX capture misc/data/file_sink
 * and against the compiled reference): gr-rfid/lib/reader_impl.cc:43-129 (sample counts and
 * frames), :131-162 (command bit fields), :173-192 (print_results), :200-380 (state machine),
 * :383-443 (CRC-5). */
#include "reader_impl.h"

#include <gnuradio/io_signature.h>

#include <cstring>

namespace gr {
namespace rfid {

reader::sptr reader::make(int sample_rate, int dac_rate)
{
  return gnuradio::get_initial_sptr(new reader_impl(sample_rate, dac_rate));
}

/* a PIE symbol: carrier on, then a low pulse of `low` samples at its end */
reader_impl::wave reader_impl::high_then_low(int total, int low)
{
  wave w((size_t)total, 0.0f);
  for (int i = 0; i < total - low; i++) w[(size_t)i] = 1.0f;
  return w;
}

reader_impl::reader_impl(int sample_rate, int dac_rate)
    : gr::block("reader", gr::io_signature::make(1, 1, sizeof(float)), gr::io_signature::make(1, 1, sizeof(float)))
{
  (void)sample_rate;
  GR_LOG_INFO(d_logger, "Block initialized");
  /* durations are given in us; the reference keeps the sample counts in floats derived from a
   * float sample period (reader_impl.cc:51-60) and truncates when sizing the vectors */
  const float sample_d = 1.0 / dac_rate * pow(10, 6);
  const float n_data0 = 2 * PW_D / sample_d, n_data1 = 4 * PW_D / sample_d, n_pw = PW_D / sample_d;
  const float n_cw = CW_D / sample_d, n_delim = DELIM_D / sample_d, n_trcal = TRCAL_D / sample_d;
  const int pw = (int)n_pw;

  /* data-0: half on / half off; data-1: three quarters on (reader_impl.cc:92-93) */
  d_data0 = wave((size_t)n_data0, 0.0f);
  for (size_t i = 0; i < d_data0.size() / 2; i++) d_data0[i] = 1.0f;
  d_data1 = wave((size_t)n_data1, 0.0f);
  for (size_t i = 0; i < 3 * d_data1.size() / 4; i++) d_data1[i] = 1.0f;
  d_delim = wave((size_t)n_delim, 0.0f);
  d_rtcal = high_then_low((int)(n_data0 + n_data1), pw);
  d_trcal = high_then_low((int)n_trcal, pw);
  d_cw = wave((size_t)n_cw, 1.0f);

  /* carrier that powers the tag while it answers (reader_impl.cc:69-71) */
  d_cw_after_query = wave((size_t)((T1_D + T2_D + RN16_D) / sample_d), 1.0f);
  d_cw_after_ack = wave((size_t)((3 * T1_D + T2_D + EPC_D) / sample_d), 1.0f);
  d_power_down = wave((size_t)(P_DOWN_D / sample_d), 0.0f);

  GR_LOG_INFO(d_logger, "Number of samples data 0 : " << d_data0.size());
  GR_LOG_INFO(d_logger, "Number of samples data 1 : " << d_data1.size());
  GR_LOG_INFO(d_logger, "Number of slots : " << std::pow(2, FIXED_Q));
  GR_LOG_INFO(d_logger, "Carrier wave after a query transmission in samples : " << d_cw_after_query.size());
  GR_LOG_INFO(d_logger, "Carrier wave after ACK transmission in samples : " << d_cw_after_ack.size());

  /* frames: preamble = delimiter, data-0, RTcal, TRcal; frame-sync = the same without TRcal */
  append(d_frame_sync, d_delim);
  append(d_frame_sync, d_data0);
  append(d_frame_sync, d_rtcal);
  d_preamble = d_frame_sync;
  append(d_preamble, d_trcal);
  /* QueryRep = frame-sync + 00 + session 00; NAK = frame-sync + 11000000 */
  d_query_rep = d_frame_sync;
  for (int i = 0; i < 4; i++) append(d_query_rep, d_data0);
  d_nak = d_frame_sync;
  for (int i = 0; i < 8; i++) append(d_nak, NAK_CODE[i] ? d_data1 : d_data0);

  build_query_bits();
  d_query_adjust_bits.assign(QADJ_CODE, QADJ_CODE + 4);
  d_query_adjust_bits.insert(d_query_adjust_bits.end(), SESSION, SESSION + 2);
  d_query_adjust_bits.insert(d_query_adjust_bits.end(), Q_UPDN[1], Q_UPDN[1] + 3);
}

reader_impl::~reader_impl() {}

/* Query = 1000 | DR | M | TRext | Sel | Session | Target | Q | CRC-5 */
void reader_impl::build_query_bits()
{
  std::vector<int>& b = d_query_bits;
  b.assign(QUERY_CODE, QUERY_CODE + 4);
  b.push_back(DR);
  b.insert(b.end(), M, M + 2);
  b.push_back(TREXT);
  b.insert(b.end(), SEL, SEL + 2);
  b.insert(b.end(), SESSION, SESSION + 2);
  b.push_back(TARGET);
  b.insert(b.end(), Q_VALUE[FIXED_Q], Q_VALUE[FIXED_Q] + 4);
  crc5_append(b);
}

/* CRC-5 of EPC Gen2 (x^5 + x^3 + 1, preset 01001) over the 17 Query bits, appended MSB first.
 * The reference runs the same shift register one bit per loop iteration (reader_impl.cc:383-443). */
void reader_impl::crc5_append(std::vector<int>& bits)
{
  unsigned reg = 0x09;
  for (size_t i = 0; i < 17; i++) {
    const unsigned fb = ((reg >> 4) & 1u) ^ (unsigned)(bits[i] & 1);
    reg = (reg << 1) & 0x1Fu;
    if (fb) reg ^= 0x09;
  }
  for (int k = 4; k >= 0; k--) bits.push_back((int)((reg >> k) & 1u));
}

int reader_impl::emit(float* out, int at, const wave& w) const
{
  if (!w.empty()) std::memcpy(out + at, &w[0], w.size() * sizeof(float));
  return at + (int)w.size();
}

int reader_impl::emit_bits(float* out, int at, const std::vector<int>& bits) const
{
  for (size_t i = 0; i < bits.size(); i++) at = emit(out, at, bits[i] == 1 ? d_data1 : d_data0);
  return at;
}

void reader_impl::print_results()
{
  const READER_STATS& rs = reader_state->reader_stats;
  std::cout << "\n --------------------------" << std::endl;
  std::cout << "| Number of queries/queryreps sent : " << rs.n_queries_sent - 1 << std::endl;
  std::cout << "| Current Inventory round : " << rs.cur_inventory_round << std::endl;
  std::cout << " --------------------------" << std::endl;
  std::cout << "| Correctly decoded EPC : " << rs.n_epc_correct << std::endl;
  std::cout << "| Number of unique tags : " << rs.tag_reads.size() << std::endl;
  for (std::map<int, int>::const_iterator it = rs.tag_reads.begin(); it != rs.tag_reads.end(); ++it) {
    std::cout << std::hex << "| Tag ID : " << it->first << "  ";
    std::cout << "Num of reads : " << std::dec << it->second << std::endl;
  }
  std::cout << " --------------------------" << std::endl;
}

void reader_impl::forecast(int noutput_items, gr_vector_int& ninput_items_required)
{
  (void)noutput_items;
  ninput_items_required[0] = 0; /* the logic runs without input; RN16 bits arrive when they arrive */
}

int reader_impl::general_work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
                              gr_vector_void_star& output_items)
{
  (void)noutput_items;
  const float* in = static_cast<const float*>(input_items[0]);
  float* out = static_cast<float*>(output_items[0]);
  READER_STATE* st = reader_state;
  const int n_in = ninput_items[0];
  int w = 0;

  switch (st->gen2_logic_status) {
    case START: /* power the tag before the first Query */
      GR_LOG_INFO(d_debug_logger, "START");
      w = emit(out, w, d_cw_after_ack);
      st->gen2_logic_status = SEND_QUERY;
      break;

    case POWER_DOWN:
      GR_LOG_INFO(d_debug_logger, "POWER DOWN");
      w = emit(out, w, d_power_down);
      st->gen2_logic_status = START;
      break;

    case SEND_NAK_QR:
    case SEND_NAK_Q:
      GR_LOG_INFO(d_debug_logger, "SEND NAK");
      w = emit(out, w, d_nak);
      w = emit(out, w, d_cw);
      st->gen2_logic_status = st->gen2_logic_status == SEND_NAK_QR ? SEND_QUERY_REP : SEND_QUERY;
      break;

    case SEND_QUERY:
    case SEND_QUERY_REP:
    case SEND_QUERY_ADJUST: {
      /* a new slot: the next window is an RN16 (reader_impl.cc:259-262, 333-336, 350-353) */
      const GEN2_LOGIC_STATUS cmd = st->gen2_logic_status;
      GR_LOG_INFO(d_debug_logger, (cmd == SEND_QUERY ? "QUERY" : cmd == SEND_QUERY_REP ? "SEND QUERY_REP" : "SEND QUERY_ADJUST"));
      GR_LOG_INFO(d_debug_logger, "INVENTORY ROUND : " << st->reader_stats.cur_inventory_round
                                      << " SLOT NUMBER : " << st->reader_stats.cur_slot_number);
      st->reader_stats.n_queries_sent += 1;
      st->decoder_status = DECODER_DECODE_RN16;
      st->gate_status = GATE_SEEK_RN16;
      if (cmd == SEND_QUERY) {
        w = emit(out, w, d_preamble);
        w = emit_bits(out, w, d_query_bits);
      } else if (cmd == SEND_QUERY_REP) {
        w = emit(out, w, d_query_rep);
      } else {
        w = emit(out, w, d_frame_sync);
        w = emit_bits(out, w, d_query_adjust_bits);
      }
      w = emit(out, w, d_cw_after_query);
      st->gen2_logic_status = IDLE;
      break;
    }

    case SEND_ACK:
      GR_LOG_INFO(d_debug_logger, "SEND ACK");
      if (n_in == RN16_BITS - 1) { /* waits until the decoder has delivered all 16 bits (:292) */
        st->decoder_status = DECODER_DECODE_EPC;
        st->gate_status = GATE_SEEK_EPC;
        std::vector<int> ack(ACK_CODE, ACK_CODE + 2);
        for (int i = 0; i < RN16_BITS - 1; i++) ack.push_back(in[i] == 1.0f ? 1 : 0);
        w = emit(out, w, d_frame_sync);
        w = emit_bits(out, w, ack);
        st->gen2_logic_status = SEND_CW;
      }
      break;

    case SEND_CW: /* carrier for the EPC reply */
      GR_LOG_INFO(d_debug_logger, "SEND CW");
      w = emit(out, w, d_cw_after_ack);
      st->gen2_logic_status = IDLE;
      break;

    default: /* IDLE */
      break;
  }
  consume_each(n_in); /* input is dropped in every state (reader_impl.cc:214,378) */
  return w;
}

}  // namespace rfid
}  // namespace gr
