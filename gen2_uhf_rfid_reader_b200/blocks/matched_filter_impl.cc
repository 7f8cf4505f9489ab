/* matched_filter_impl.cc -- thin host over rfid_b200_mf_work (canonical block-sum boxcar + decimation on the
 * GPU, ntaps-1 samples of history kept in the context, like gr::filter::fir_filter_ccc's history). */
#include "matched_filter_impl.h"

#include <gnuradio/io_signature.h>

namespace gr {
namespace rfid {

matched_filter::sptr matched_filter::make(int decim, int ntaps)
{
  return gnuradio::get_initial_sptr(new matched_filter_impl(decim, ntaps));
}

matched_filter_impl::matched_filter_impl(int decim, int ntaps)
    : gr::block("matched_filter", gr::io_signature::make(1, 1, sizeof(gr_complex)), gr::io_signature::make(1, 1, sizeof(gr_complex))),
      d_ctx(0), d_decim(decim)
{
  rfid_b200_params p;
  rfid_b200_default_params(&p);
  p.decim = decim;
  p.ntaps = ntaps;
  const char* dev = std::getenv("RFID_B200_DEVICE");
  if (dev) p.device = std::atoi(dev);
  b200_check(rfid_b200_create(&p, &d_ctx), 0, "rfid::matched_filter: rfid_b200_create (no CPU fallback)");
}

matched_filter_impl::~matched_filter_impl() { rfid_b200_destroy(d_ctx); }

void matched_filter_impl::forecast(int noutput_items, gr_vector_int& ninput_items_required)
{
  ninput_items_required[0] = noutput_items * d_decim;  /* fixed-rate decimator */
}

int matched_filter_impl::general_work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
                                      gr_vector_void_star& output_items)
{
  const float* in = static_cast<const float*>(input_items[0]);
  float* out = static_cast<float*>(output_items[0]);
  /* consume whole decimation groups only, and no more than the output buffer can take */
  int n_in = ninput_items[0] / d_decim * d_decim;
  if (n_in / d_decim > noutput_items) n_in = noutput_items * d_decim;
  int written = 0;
  if (n_in > 0) b200_check(rfid_b200_mf_work(d_ctx, in, n_in, out, noutput_items + 1, &written), d_ctx, "rfid_b200_mf_work");
  consume_each(n_in);
  return written;
}

}  // namespace rfid
}  // namespace gr
