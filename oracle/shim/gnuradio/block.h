/* TEST INFRASTRUCTURE (oracle shim) -- minimal gr::block.
 * Provides exactly the gr::block services the gr-rfid blocks use
 * (reference: gate_impl.cc:42-44,198; tag_decoder_impl.cc:51-53,266,395-396;
 * reader_impl.cc:44-46,378): constructor with name + signatures, consume_each,
 * produce, WORK_CALLED_PRODUCE, loggers.  The driver reads back what a block
 * consumed/produced through the shim_* accessors. */
#ifndef ORACLE_SHIM_GNURADIO_BLOCK_H
#define ORACLE_SHIM_GNURADIO_BLOCK_H
#include <gnuradio/attributes.h>
#include <gnuradio/io_signature.h>

typedef std::complex<float> gr_complex;
typedef std::vector<int> gr_vector_int;
typedef std::vector<const void*> gr_vector_const_void_star;
typedef std::vector<void*> gr_vector_void_star;

namespace boost {
using std::shared_ptr;
}
namespace gnuradio {
template <class T>
boost::shared_ptr<T> get_initial_sptr(T* p) { return boost::shared_ptr<T>(p); }
}  // namespace gnuradio

namespace gr {
struct shim_logger {
  bool enabled;
  const char* tag;
};
typedef shim_logger* logger_ptr;

class block {
 public:
  enum { WORK_CALLED_PRODUCE = -2, WORK_DONE = -1 };
  block() : d_logger(&s_info), d_debug_logger(&s_debug), d_consumed(0) {}
  block(const std::string& name, io_signature::sptr in, io_signature::sptr out)
      : d_logger(&s_info), d_debug_logger(&s_debug), d_name(name), d_in(in), d_out(out), d_consumed(0) {
    d_produced.assign(8, 0);
  }
  virtual ~block() {}
  virtual void forecast(int, gr_vector_int&) {}
  virtual int general_work(int, gr_vector_int&, gr_vector_const_void_star&, gr_vector_void_star&) { return 0; }
  void consume_each(int n) { d_consumed += n; }
  void consume(int, int n) { d_consumed += n; }
  void produce(int port, int n) {
    if (d_produced.size() < 8) d_produced.assign(8, 0);
    d_produced[port] += n;
  }
  /* driver side */
  void shim_reset_counts() { d_consumed = 0; d_produced.assign(8, 0); }
  int shim_consumed() const { return d_consumed; }
  int shim_produced(int port) const { return port < (int)d_produced.size() ? d_produced[port] : 0; }
  static shim_logger s_info, s_debug;

 protected:
  logger_ptr d_logger, d_debug_logger;
  std::string d_name;
  io_signature::sptr d_in, d_out;

 private:
  int d_consumed;
  std::vector<int> d_produced;
};
}  // namespace gr

#define ORACLE_SHIM_LOG(lg, msg)                                     \
  do {                                                               \
    if ((lg) && (lg)->enabled) {                                     \
      std::ostringstream _oss; _oss << msg;                          \
      std::cerr << (lg)->tag << " " << _oss.str() << std::endl;      \
    }                                                                \
  } while (0)
#define GR_LOG_INFO(lg, msg) ORACLE_SHIM_LOG(lg, msg)
#define GR_LOG_DEBUG(lg, msg) ORACLE_SHIM_LOG(lg, msg)
#define GR_LOG_WARN(lg, msg) ORACLE_SHIM_LOG(lg, msg)
#define GR_LOG_EMERG(lg, msg) ORACLE_SHIM_LOG(lg, msg)
#endif
