/* global_vars.cc -- definition of the state shared by the gr::rfid blocks
 * (replaces gr-rfid/lib/global_vars.cc:32-54). */
#include "rfid/global_vars.h"

namespace gr {
namespace rfid {

READER_STATE* reader_state = 0;

void initialize_reader_state()
{
  /* the reference allocates a fresh record each time a gate block is constructed */
  READER_STATE* st = new READER_STATE;
  READER_STATS& rs = st->reader_stats;
  rs.n_queries_sent = 0;
  rs.n_epc_correct = 0;
  rs.cur_inventory_round = 1;
  rs.cur_slot_number = 1;
  rs.max_slot_number = (int)std::pow(2, FIXED_Q);
  rs.max_inventory_round = 0;
  st->status = RUNNING;
  st->gen2_logic_status = START;
  st->gate_status = GATE_SEEK_RN16;
  st->decoder_status = DECODER_DECODE_RN16;
  st->n_samples_to_ungate = 0;
  gettimeofday(&rs.start, 0);
  rs.end = rs.start;
  reader_state = st;
}

}  // namespace rfid
}  // namespace gr
