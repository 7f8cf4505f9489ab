/* reader_impl.h -- Gen2 logic + PIE command generator (host side; replaces
 * gr-rfid/lib/reader_impl.h).  Not on the GPU path: it emits a few thousand envelope samples per
 * command and closes the protocol loop by steering the gate and the decoder. */
#ifndef INCLUDED_RFID_READER_IMPL_H
#define INCLUDED_RFID_READER_IMPL_H

#include <rfid/reader.h>

#include <vector>

#include "rfid/global_vars.h"

namespace gr {
namespace rfid {

class reader_impl : public reader
{
  typedef std::vector<float> wave;

  /* PIE symbols and fixed frames as envelope samples (1 = carrier on) */
  wave d_data0, d_data1, d_delim, d_rtcal, d_trcal, d_cw;
  wave d_preamble, d_frame_sync, d_query_rep, d_nak;
  wave d_cw_after_query, d_cw_after_ack, d_power_down;
  std::vector<int> d_query_bits, d_query_adjust_bits;

  static wave high_then_low(int total, int low);
  static void append(wave& dst, const wave& src) { dst.insert(dst.end(), src.begin(), src.end()); }
  int emit(float* out, int at, const wave& w) const;
  int emit_bits(float* out, int at, const std::vector<int>& bits) const;

  void build_query_bits();
  static void crc5_append(std::vector<int>& bits);

public:
  reader_impl(int sample_rate, int dac_rate);
  ~reader_impl();

  void print_results();
  void forecast(int noutput_items, gr_vector_int& ninput_items_required);
  int general_work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
                   gr_vector_void_star& output_items);

  /* for tests */
  const std::vector<int>& query_bits() const { return d_query_bits; }
};

}  // namespace rfid
}  // namespace gr
#endif
