/* tag_decoder_impl.cc -- the tag decoder block as a thin host.  Frame sync, channel estimate,
 * FM0 decisions, symbol-period search and CRC-16 run in the CUDA kernel behind
 * rfid_b200_decoder_work(); this file keeps the GNU Radio contract, the inventory bookkeeping
 * and the hand-over to the Gen2 logic (gr-rfid/lib/tag_decoder_impl.cc:196-397). */
#include "tag_decoder_impl.h"

#include <gnuradio/io_signature.h>

#include <cstring>

namespace gr {
namespace rfid {

tag_decoder::sptr tag_decoder::make(int sample_rate)
{
  std::vector<int> sizes;
  sizes.push_back(sizeof(float));       /* port 0: RN16 bits for the reader block */
  sizes.push_back(sizeof(gr_complex));  /* port 1: debug stream (connected, never produced) */
  return gnuradio::get_initial_sptr(new tag_decoder_impl(sample_rate, sizes));
}

tag_decoder_impl::tag_decoder_impl(int sample_rate, std::vector<int> output_sizes)
    : gr::block("tag_decoder", gr::io_signature::make(1, 1, sizeof(gr_complex)), gr::io_signature::makev(2, 2, output_sizes)),
      d_ctx(0)
{
  std::memset(&d_last, 0, sizeof(d_last));
  d_ctx = b200_make_context(sample_rate, "tag_decoder");
}

tag_decoder_impl::~tag_decoder_impl() { rfid_b200_destroy(d_ctx); }

void tag_decoder_impl::forecast(int noutput_items, gr_vector_int& ninput_items_required)
{
  ninput_items_required[0] = noutput_items;
}

/* end of a slot: move to the next slot or, after the last one, to the next inventory round
 * (tag_decoder_impl.cc:271-287, 331-343, 369-383) */
void tag_decoder_impl::next_slot(bool count_round_tags)
{
  READER_STATS& rs = reader_state->reader_stats;
  if (rs.cur_slot_number > rs.max_slot_number) {
    rs.cur_slot_number = 1;
    if (count_round_tags) rs.unique_tags_round.push_back((int)rs.tag_reads.size());
    rs.cur_inventory_round += 1;
    reader_state->gen2_logic_status = SEND_QUERY;
  } else {
    reader_state->gen2_logic_status = SEND_QUERY_REP;
  }
}

int tag_decoder_impl::general_work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
                                   gr_vector_void_star& output_items)
{
  (void)noutput_items;
  const float* in = static_cast<const float*>(input_items[0]);
  float* out = static_cast<float*>(output_items[0]);
  READER_STATE* st = reader_state;
  READER_STATS& rs = st->reader_stats;
  const int have = ninput_items[0];
  const int need = st->n_samples_to_ungate;
  int consumed = 0;

  /* the reference only acts once the gate has delivered a complete window (:223, :291) */
  if (need > 0 && have >= need) {
    float bits[128];
    if (st->decoder_status == DECODER_DECODE_RN16) {
      b200_check(rfid_b200_decoder_work(d_ctx, RFID_B200_RN16, in, have, &d_last, bits), d_ctx, "rfid_b200_decoder_work");
      if (d_last.crc_ok != -2) {
        GR_LOG_INFO(d_debug_logger, "RN16 DECODED");
        std::memcpy(out, bits, 16 * sizeof(float)); /* for the ACK (:261-266) */
        produce(0, 16);
        st->gen2_logic_status = SEND_ACK;
      } else {
        rs.cur_slot_number++;
        next_slot(true);
      }
    } else {
      rs.cur_slot_number++; /* an EPC window always ends the slot (:295) */
      b200_check(rfid_b200_decoder_work(d_ctx, RFID_B200_EPC, in, have, &d_last, bits), d_ctx, "rfid_b200_decoder_work");
      if (d_last.crc_ok == 1) {
        next_slot(true);
        rs.n_epc_correct += 1;
        const int tag = d_last.tag_id; /* EPC bits 104..111 (:348-352) */
        GR_LOG_INFO(d_debug_logger, "EPC CORRECTLY DECODED, TAG ID : " << tag);
        std::map<int, int>::iterator it = rs.tag_reads.find(tag);
        if (it == rs.tag_reads.end()) rs.tag_reads[tag] = 1;
        else it->second++;
      } else {
        next_slot(false); /* a CRC failure does not record the round's tag count (:369-378) */
        GR_LOG_INFO(d_debug_logger, "EPC FAIL TO DECODE");
      }
    }
    consumed = need;
  }
  consume_each(consumed);
  return WORK_CALLED_PRODUCE;
}

}  // namespace rfid
}  // namespace gr
