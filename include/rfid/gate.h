/* Interface compatibility header: the class / function / constant names, signatures and values declared here are those of
 * gr-rfid's installed public header of the same name (nkargas/Gen2-UHF-RFID-Reader, Copyright 2014 Nikos Kargas
 * <nkargas@isc.tuc.gr>, GNU General Public License version 3 or later), because blocks built against it must be
 * drop-in replacements.  The implementation behind the interface is this repository's own.  This file is distributed
 * under the GNU General Public License, version 3 or (at your option) any later version; see LICENSE. */
/* rfid/gate.h -- public interface of the gate block (drop-in for gr-rfid/include/rfid/gate.h:38-53).
 *
 * The gate watches the matched-filtered stream for reader commands (amplitude dips), swallows
 * them, and forwards exactly one RN16- or EPC-sized window of DC-removed tag samples after each
 * command.  In this build the work is done on the GPU: general_work() marshals the scheduler's
 * buffers through rfid_b200_gate_work() (include/rfid_b200.h). */
#ifndef INCLUDED_RFID_GATE_H
#define INCLUDED_RFID_GATE_H

#include <gnuradio/block.h>
#include <rfid/api.h>

namespace gr {
namespace rfid {

class RFID_API gate : virtual public gr::block
{
public:
  typedef boost::shared_ptr<gate> sptr;

  /*! \param sample_rate rate of the (decimated) input stream in Hz, e.g. 400000 */
  static sptr make(int sample_rate);
};

}  // namespace rfid
}  // namespace gr

#endif /* INCLUDED_RFID_GATE_H */
