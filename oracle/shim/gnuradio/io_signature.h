/* TEST INFRASTRUCTURE (oracle shim) -- minimal gr::io_signature. */
#ifndef ORACLE_SHIM_GNURADIO_IO_SIGNATURE_H
#define ORACLE_SHIM_GNURADIO_IO_SIGNATURE_H
#include <gnuradio/attributes.h>
namespace gr {
class io_signature {
 public:
  typedef std::shared_ptr<io_signature> sptr;
  int min_streams, max_streams;
  std::vector<int> sizeof_stream_items;
  static sptr make(int min_s, int max_s, int item_size) {
    sptr p(new io_signature);
    p->min_streams = min_s; p->max_streams = max_s;
    p->sizeof_stream_items.assign(max_s > 0 ? max_s : 1, item_size);
    return p;
  }
  static sptr makev(int min_s, int max_s, const std::vector<int>& sizes) {
    sptr p(new io_signature);
    p->min_streams = min_s; p->max_streams = max_s; p->sizeof_stream_items = sizes;
    return p;
  }
};
}  // namespace gr
#endif
