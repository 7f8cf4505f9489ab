/* rfid/reader.h -- public interface of the Gen2 logic / command generator block (drop-in for
 * gr-rfid/include/rfid/reader.h:38-53).  Host-side only: it is not on the GPU path. */
#ifndef INCLUDED_RFID_READER_H
#define INCLUDED_RFID_READER_H

#include <gnuradio/block.h>
#include <rfid/api.h>

namespace gr {
namespace rfid {

class RFID_API reader : virtual public gr::block
{
public:
  typedef boost::shared_ptr<reader> sptr;

  /*! prints the inventory statistics block the README documents (README.md:46-53) */
  virtual void print_results() = 0;

  /*! \param sample_rate rate of the RX chain after decimation, Hz
   *  \param dac_rate    rate of the generated TX envelope, Hz */
  static sptr make(int sample_rate, int dac_rate);
};

}  // namespace rfid
}  // namespace gr

#endif /* INCLUDED_RFID_READER_H */
