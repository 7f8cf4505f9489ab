/* Interface compatibility header: the class / function / constant names, signatures and values declared here are those of
 * gr-rfid's installed public header of the same name (nkargas/Gen2-UHF-RFID-Reader, Copyright 2014 Nikos Kargas
 * <nkargas@isc.tuc.gr>, GNU General Public License version 3 or later), because blocks built against it must be
 * drop-in replacements.  The implementation behind the interface is this repository's own.  This file is distributed
 * under the GNU General Public License, version 3 or (at your option) any later version; see LICENSE. */
/* rfid/tag_decoder.h -- public interface of the tag decoder block (drop-in for
 * gr-rfid/include/rfid/tag_decoder.h:35-49).
 *
 * Input: the gate's windows (complex).  Output port 0: the 16 RN16 bits as floats 0./1. for the
 * reader block; output port 1: a complex debug stream that is connected by apps/reader.py:116 but
 * never produced.  Decoding runs on the GPU through rfid_b200_decoder_work(). */
#ifndef INCLUDED_RFID_TAG_DECODER_H
#define INCLUDED_RFID_TAG_DECODER_H

#include <gnuradio/block.h>
#include <rfid/api.h>

namespace gr {
namespace rfid {

class RFID_API tag_decoder : virtual public gr::block
{
public:
  typedef boost::shared_ptr<tag_decoder> sptr;

  /*! \param sample_rate rate of the (decimated) input stream in Hz */
  static sptr make(int sample_rate);
};

}  // namespace rfid
}  // namespace gr

#endif /* INCLUDED_RFID_TAG_DECODER_H */
