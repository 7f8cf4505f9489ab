/* matched_filter_impl.h -- GPU replacement for the flowgraph's own matched filter
 * `filter.fir_filter_ccc(self.decim, [1]*25)` (gr-rfid/apps/reader.py:65,75; GNU Radio block, external to
 * the reference).  SURVEY.md section 8(f) rank 3: with this block in the graph the decimating boxcar also runs on
 * the GPU in block mode.  Usage in reader.py:   self.matched_filter = rfid.matched_filter(self.decim, 25)   */
#ifndef INCLUDED_RFID_MATCHED_FILTER_IMPL_H
#define INCLUDED_RFID_MATCHED_FILTER_IMPL_H

#include <gnuradio/block.h>
#include <rfid/api.h>

#include "b200_block_common.h"

namespace gr {
namespace rfid {

class RFID_API matched_filter : virtual public gr::block
{
public:
  typedef boost::shared_ptr<matched_filter> sptr;
  /*! boxcar of `ntaps` ones, decimation `decim` (fir_filter_ccc(decim, [1]*ntaps)) */
  static sptr make(int decim, int ntaps);
};

class matched_filter_impl : public matched_filter
{
  rfid_b200_ctx* d_ctx;
  int d_decim;

public:
  matched_filter_impl(int decim, int ntaps);
  ~matched_filter_impl();
  void forecast(int noutput_items, gr_vector_int& ninput_items_required);
  int general_work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
                   gr_vector_void_star& output_items);
};

}  // namespace rfid
}  // namespace gr
#endif
