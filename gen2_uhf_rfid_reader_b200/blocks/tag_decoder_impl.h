/* tag_decoder_impl.h -- thin GNU Radio host for the GPU tag decoder (replaces
 * gr-rfid/lib/tag_decoder_impl.h). */
#ifndef INCLUDED_RFID_TAG_DECODER_IMPL_H
#define INCLUDED_RFID_TAG_DECODER_IMPL_H

#include <rfid/tag_decoder.h>

#include <vector>

#include "b200_block_common.h"
#include "rfid/global_vars.h"

namespace gr {
namespace rfid {

class tag_decoder_impl : public tag_decoder
{
  rfid_b200_ctx* d_ctx;
  rfid_b200_window_result d_last;

  void next_slot(bool count_round_tags);

public:
  tag_decoder_impl(int sample_rate, std::vector<int> output_sizes);
  ~tag_decoder_impl();

  /*! everything the GPU derived from the most recent window (score, channel estimate, T, bits, CRC) */
  const rfid_b200_window_result* last_result() const { return &d_last; }

  void forecast(int noutput_items, gr_vector_int& ninput_items_required);
  int general_work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
                   gr_vector_void_star& output_items);
};

}  // namespace rfid
}  // namespace gr
#endif
