/* Interface compatibility header: the class / function / constant names, signatures and values declared here are those of
 * gr-rfid's installed public header of the same name (nkargas/Gen2-UHF-RFID-Reader, Copyright 2014 Nikos Kargas
 * <nkargas@isc.tuc.gr>, GNU General Public License version 3 or later), because blocks built against it must be
 * drop-in replacements.  The implementation behind the interface is this repository's own.  This file is distributed
 * under the GNU General Public License, version 3 or (at your option) any later version; see LICENSE. */
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
