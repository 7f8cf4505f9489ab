/* gate_impl.cc -- the gate block as a thin host: every sample decision is made by the CUDA kernel
 * behind rfid_b200_gate_work(); this file only keeps the GNU Radio contract and the shared
 * reader_state flags of the reference (gr-rfid/lib/gate_impl.cc:32-200). */
#include "gate_impl.h"

#include <gnuradio/io_signature.h>
#include <sys/time.h>

namespace gr {
namespace rfid {

gate::sptr gate::make(int sample_rate) { return gnuradio::get_initial_sptr(new gate_impl(sample_rate)); }

gate_impl::gate_impl(int sample_rate)
    : gr::block("gate", gr::io_signature::make(1, 1, sizeof(gr_complex)), gr::io_signature::make(1, 1, sizeof(gr_complex))),
      d_ctx(0), d_window_open(false)
{
  d_ctx = b200_make_context(sample_rate, "gate");
  GR_LOG_INFO(d_logger, "gate on CUDA device; RN16 window " << rfid_b200_window_length(d_ctx, RFID_B200_RN16)
                            << " samples, EPC window " << rfid_b200_window_length(d_ctx, RFID_B200_EPC) << " samples");
  /* first block to be constructed by apps/reader.py:76: it owns the shared state (gate_impl.cc:69) */
  initialize_reader_state();
}

gate_impl::~gate_impl() { rfid_b200_destroy(d_ctx); }

void gate_impl::forecast(int noutput_items, gr_vector_int& ninput_items_required)
{
  ninput_items_required[0] = noutput_items;
}

int gate_impl::general_work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
                            gr_vector_void_star& output_items)
{
  (void)noutput_items;
  const float* in = static_cast<const float*>(input_items[0]);
  float* out = static_cast<float*>(output_items[0]);
  const int n_items = ninput_items[0];
  READER_STATE* st = reader_state;

  /* stop rule, evaluated once per work call like the reference (gate_impl.cc:101-109) */
  const bool budget_spent = st->reader_stats.n_queries_sent > MAX_NUM_QUERIES ||
                            st->reader_stats.tag_reads.size() > (size_t)NUMBER_UNIQUE_TAGS;
  if (budget_spent && st->status != TERMINATED) {
    st->status = TERMINATED;
    gettimeofday(&st->reader_stats.end, 0);
    std::cout << "| Execution time : " << st->reader_stats.end.tv_sec - st->reader_stats.start.tv_sec << " seconds"
              << std::endl;
    GR_LOG_INFO(d_logger, "Termination");
  }

  /* the Gen2 logic block asks for the next window through gate_status (gate_impl.cc:112-123) */
  int seek = 0;
  if (st->gate_status == GATE_SEEK_EPC) seek = 2;
  else if (st->gate_status == GATE_SEEK_RN16) seek = 1;
  if (seek) {
    st->gate_status = GATE_CLOSED;
    st->n_samples_to_ungate = rfid_b200_window_length(d_ctx, seek == 2 ? RFID_B200_EPC : RFID_B200_RN16);
    d_window_open = false;
  }

  int consumed = n_items, written = 0, closed = 0;
  if (st->status == RUNNING) {
    if ((int)d_magn.size() < n_items) d_magn.resize(n_items);
    b200_check(rfid_b200_gate_work(d_ctx, seek, in, n_items, out, n_items, &consumed, &written, &closed,
                                   n_items ? &d_magn[0] : 0),
               d_ctx, "rfid_b200_gate_work");
    if (written > 0) {
      if (!d_window_open) {
        GR_LOG_INFO(d_debug_logger, "READER COMMAND DETECTED");
        d_window_open = true;
        st->gate_status = GATE_OPEN;
        st->magn_squared_samples.resize(0);
      }
      st->magn_squared_samples.insert(st->magn_squared_samples.end(), d_magn.begin(), d_magn.begin() + written);
    }
    if (closed) {
      d_window_open = false;
      st->gate_status = GATE_CLOSED;
    }
  }
  consume_each(consumed);
  return written;
}

}  // namespace rfid
}  // namespace gr
