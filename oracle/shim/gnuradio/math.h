/* TEST INFRASTRUCTURE (oracle shim) -- empty <gnuradio/math.h>. */
#include <gnuradio/attributes.h>
