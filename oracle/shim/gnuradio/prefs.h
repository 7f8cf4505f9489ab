/* TEST INFRASTRUCTURE (oracle shim) -- empty <gnuradio/prefs.h>. */
#include <gnuradio/attributes.h>
