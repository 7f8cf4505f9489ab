/* TEST INFRASTRUCTURE (oracle shim) -- stands in for <gnuradio/attributes.h>.
 * GNU Radio is not installed in this image; this shim exists only so that the
 * reference's block sources can be compiled *unchanged, where they lie* under
 * /root/reference by oracle/build_ref.sh, and so that this repo's own thin host
 * blocks (gen2_uhf_rfid_reader_b200/blocks) can be exercised by the same
 * deterministic scheduler.  Never linked into the product library. */
#ifndef ORACLE_SHIM_GNURADIO_ATTRIBUTES_H
#define ORACLE_SHIM_GNURADIO_ATTRIBUTES_H
/* the reference relies on transitive includes from the real GNU Radio headers */
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
#define __GR_ATTR_EXPORT __attribute__((visibility("default")))
#define __GR_ATTR_IMPORT __attribute__((visibility("default")))
#endif
