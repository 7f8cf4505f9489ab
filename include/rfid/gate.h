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
