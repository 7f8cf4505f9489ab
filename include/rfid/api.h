/* rfid/api.h -- symbol visibility for the gr::rfid block library (B200 host blocks).
 * Same macro name the reference installs (gr-rfid/include/rfid/api.h:27-31) so that
 * application code including <rfid/...> compiles unchanged. */
#ifndef INCLUDED_RFID_API_H
#define INCLUDED_RFID_API_H

#include <gnuradio/attributes.h>

#if defined(gnuradio_rfid_EXPORTS)
#define RFID_API __GR_ATTR_EXPORT
#else
#define RFID_API __GR_ATTR_IMPORT
#endif

#endif /* INCLUDED_RFID_API_H */
