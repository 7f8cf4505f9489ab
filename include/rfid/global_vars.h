/* Interface compatibility header: the class / function / constant names, signatures and values declared here are those of
 * gr-rfid's installed public header of the same name (nkargas/Gen2-UHF-RFID-Reader, Copyright 2014 Nikos Kargas
 * <nkargas@isc.tuc.gr>, GNU General Public License version 3 or later), because blocks built against it must be
 * drop-in replacements.  The implementation behind the interface is this repository's own.  This file is distributed
 * under the GNU General Public License, version 3 or (at your option) any later version; see LICENSE. */
/* rfid/global_vars.h -- protocol constants and the state shared by the three gr::rfid blocks.
 *
 * This header is part of the reference's installed API (gr-rfid/include/rfid/global_vars.h,
 * installed by include/rfid/CMakeLists.txt:23-29): applications and the blocks themselves refer
 * to these names, so names, types and values are kept; the file itself is written for this build.
 * The GPU chain does not depend on the process-global `reader_state` (each C-ABI context carries
 * its own state); the thin host blocks keep it up to date so that code written against the
 * reference (e.g. reader::print_results, apps/reader.py) keeps working.
 */
#ifndef INCLUDED_RFID_GLOBAL_VARS_H
#define INCLUDED_RFID_GLOBAL_VARS_H

#include <rfid/api.h>
#include <sys/time.h>

#include <cmath>
#include <map>
#include <vector>

namespace gr {
namespace rfid {

/* ---- enumerations (global_vars.h:31-34 of the reference) ---- */
enum STATUS { RUNNING, TERMINATED };
enum GEN2_LOGIC_STATUS {
  SEND_QUERY, SEND_ACK, SEND_QUERY_REP, IDLE, SEND_CW, START, SEND_QUERY_ADJUST, SEND_NAK_QR, SEND_NAK_Q, POWER_DOWN
};
enum GATE_STATUS { GATE_OPEN, GATE_CLOSED, GATE_SEEK_RN16, GATE_SEEK_EPC };
enum DECODER_STATUS { DECODER_DECODE_RN16, DECODER_DECODE_EPC };

/* ---- inventory statistics (:36-53) ---- */
struct READER_STATS {
  int n_queries_sent;
  int cur_inventory_round;
  int cur_slot_number;
  int max_slot_number;
  int max_inventory_round;
  int n_epc_correct;
  std::vector<int> unique_tags_round;
  std::map<int, int> tag_reads;
  struct timeval start, end;
};

/* ---- cross-block state (:55-67) ---- */
struct READER_STATE {
  STATUS status;
  GEN2_LOGIC_STATUS gen2_logic_status;
  GATE_STATUS gate_status;
  DECODER_STATUS decoder_status;
  READER_STATS reader_stats;
  std::vector<float> magn_squared_samples; /* |gate output|^2 of the current window */
  int n_samples_to_ungate;                 /* current window length, shared by gate and decoder */
};

/* ---- reader configuration (:72-76) ---- */
const int FIXED_Q = 0;            /* 2^FIXED_Q slots per inventory round */
const int MAX_NUM_QUERIES = 1000; /* stop after this many Query/QueryRep */

/* Q as four bits, MSB first (:79-85) */
const int Q_VALUE[16][4] = {{0, 0, 0, 0}, {0, 0, 0, 1}, {0, 0, 1, 0}, {0, 0, 1, 1}, {0, 1, 0, 0}, {0, 1, 0, 1},
                            {0, 1, 1, 0}, {0, 1, 1, 1}, {1, 0, 0, 0}, {1, 0, 0, 1}, {1, 0, 1, 0}, {1, 0, 1, 1},
                            {1, 1, 0, 0}, {1, 1, 0, 1}, {1, 1, 1, 0}, {1, 1, 1, 1}};

const bool P_DOWN = false;

/* ---- durations in microseconds (:90-97) ---- */
const int CW_D = 250;       /* carrier wave after a command */
const int P_DOWN_D = 2000;  /* power down */
const int T1_D = 240;       /* reader command -> tag reply */
const int T2_D = 480;       /* tag reply -> next reader command */
const int PW_D = 12;        /* half Tari */
const int DELIM_D = 12;     /* frame delimiter */
const int TRCAL_D = 200;    /* BLF = DR / TRcal = 8 / 200 us = 40 kHz */
const int RTCAL_D = 72;     /* 6 * PW */

const int NUM_PULSES_COMMAND = 5;   /* pulses that identify a reader command (:99) */
const int NUMBER_UNIQUE_TAGS = 100; /* stop after more than this many distinct tags (:100) */

/* ---- bit counts (:103-108) ---- */
const int PILOT_TONE = 12;
const int TAG_PREAMBLE_BITS = 6;
const int RN16_BITS = 17;  /* 16 + dummy */
const int EPC_BITS = 129;  /* PC 16 + EPC 96 + CRC 16 + dummy */
const int QUERY_LENGTH = 22;

/* ---- tag link timing (:110-113) ---- */
const int T_READER_FREQ = 40e3; /* backscatter link frequency */
const float TAG_BIT_D = 1.0 / T_READER_FREQ * pow(10, 6); /* us per tag bit */
const int RN16_D = (RN16_BITS + TAG_PREAMBLE_BITS) * TAG_BIT_D;
const int EPC_D = (EPC_BITS + TAG_PREAMBLE_BITS) * TAG_BIT_D;

/* ---- command fields (:115-133) ---- */
const int QUERY_CODE[4] = {1, 0, 0, 0};
const int M[2] = {0, 0};
const int SEL[2] = {0, 0};
const int SESSION[2] = {0, 0};
const int TARGET = 0;
const int TREXT = 0;
const int DR = 0;
const int NAK_CODE[8] = {1, 1, 0, 0, 0, 0, 0, 0};
const int ACK_CODE[2] = {0, 1};
const int QADJ_CODE[4] = {1, 0, 0, 1};
const int Q_UPDN[3][3] = {{1, 1, 0}, {0, 0, 0}, {0, 1, 1}}; /* +1, unchanged, -1 */

/* FM0 preamble as half-symbol levels (:136) */
const int TAG_PREAMBLE[] = {1, 1, 0, 1, 0, 0, 1, 0, 0, 0, 1, 1};

/* ---- gate parameters (:139-143) ---- */
const float THRESH_FRACTION = 0.75;
const int WIN_SIZE_D = 250; /* us, amplitude averaging window */
const int DC_SIZE_D = 120;  /* us, DC estimation window */

/* ---- the shared state (:146-147) ---- */
extern RFID_API READER_STATE* reader_state;
extern RFID_API void initialize_reader_state();

}  // namespace rfid
}  // namespace gr

#endif /* INCLUDED_RFID_GLOBAL_VARS_H */
