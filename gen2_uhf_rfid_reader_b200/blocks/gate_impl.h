/* gate_impl.h -- thin GNU Radio host for the GPU gate (replaces gr-rfid/lib/gate_impl.h). */
#ifndef INCLUDED_RFID_GATE_IMPL_H
#define INCLUDED_RFID_GATE_IMPL_H

#include <rfid/gate.h>

#include <vector>

#include "b200_block_common.h"
#include "rfid/global_vars.h"

namespace gr {
namespace rfid {

class gate_impl : public gate
{
  rfid_b200_ctx* d_ctx;
  bool d_window_open;            /* a window is being forwarded (spans work calls) */
  std::vector<float> d_magn;     /* scratch for |out|^2 of one call */

public:
  gate_impl(int sample_rate);
  ~gate_impl();

  void forecast(int noutput_items, gr_vector_int& ninput_items_required);
  int general_work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
                   gr_vector_void_star& output_items);
};

}  // namespace rfid
}  // namespace gr
#endif
