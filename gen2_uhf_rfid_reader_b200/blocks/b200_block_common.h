/* b200_block_common.h -- what the three thin host blocks share: creation of the C-ABI context
 * for a block that only knows its (decimated) stream rate. */
#ifndef RFID_B200_BLOCK_COMMON_H
#define RFID_B200_BLOCK_COMMON_H

#include <cstdlib>
#include <stdexcept>
#include <string>

#include "rfid_b200.h"

namespace gr {
namespace rfid {

inline rfid_b200_ctx* b200_make_context(int stream_rate, const char* who)
{
  rfid_b200_params p;
  rfid_b200_default_params(&p);
  /* the flowgraph's own fir_filter_ccc runs upstream (apps/reader.py:75), so the block sees the
   * already decimated stream: describe it as "rate = stream_rate, no further decimation" */
  p.adc_rate = stream_rate;
  p.decim = 1;
  p.ntaps = 1;
  const char* dev = std::getenv("RFID_B200_DEVICE");
  if (dev) p.device = std::atoi(dev);
  rfid_b200_ctx* ctx = 0;
  int rc = rfid_b200_create(&p, &ctx);
  if (rc != RFID_B200_OK)
    throw std::runtime_error(std::string("rfid::") + who + ": rfid_b200_create failed: " + rfid_b200_strerror(rc) +
                             " (this build has no CPU fallback)");
  return ctx;
}

inline void b200_check(int rc, rfid_b200_ctx* ctx, const char* what)
{
  if (rc != RFID_B200_OK)
    throw std::runtime_error(std::string(what) + ": " + rfid_b200_strerror(rc) + " " + rfid_b200_last_cuda_error(ctx));
}

}  // namespace rfid
}  // namespace gr
#endif
