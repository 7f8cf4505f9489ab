#!/usr/bin/env bash
# TEST INFRASTRUCTURE -- builds oracle/_ref/: the reference's OWN block sources,
# compiled unchanged from where they lie under /root/reference, against the
# GNU Radio header shim (oracle/shim) and linked with the deterministic
# scheduler oracle/flow_driver.cc.  Outputs only into oracle/_ref/ (git-ignored,
# but shipped to the GPU box).  Does not run the reference's build system
# (it needs GNU Radio/Boost/CppUnit, absent here).
#
#   oracle/_ref/libgen2ref.so     FIXED_Q = 0 (reference default, global_vars.h:72)
#   oracle/_ref/libgen2ref_q4.so  FIXED_Q = 4: the header constant is shadowed by
#                                 a sed-patched copy placed earlier on the -I path
#                                 (the reference has no -D hook)
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${GEN2_REFERENCE_ROOT:-/root/reference}/gr-rfid"
OUT="$HERE/_ref"
if [ ! -d "$REF/lib" ]; then
  echo "build_ref.sh: $REF not present (GPU box?) -- keeping prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$OUT"
Q4="$(mktemp -d)"; trap 'rm -rf "$Q4"' EXIT; mkdir -p "$Q4/rfid"
SRCS="$REF/lib/global_vars.cc $REF/lib/gate_impl.cc $REF/lib/tag_decoder_impl.cc $REF/lib/reader_impl.cc"
# -O3 -DNDEBUG = the reference's CMake Release default (gr-rfid/CMakeLists.txt:29-31);
# gnu++14 because of `register` (tag_decoder_impl.cc:403-404); plain x86-64 => no FMA contraction.
CXXFLAGS="-std=gnu++14 -O3 -DNDEBUG -w -fPIC -shared -fvisibility=hidden -DDRIVE_REFERENCE"
g++ $CXXFLAGS -I"$HERE/shim" -I"$REF/include" -I"$REF/lib" \
    $SRCS "$HERE/flow_driver.cc" -o "$OUT/libgen2ref.so"
sed 's/const int FIXED_Q *= *0;/const int FIXED_Q              = 4;/' \
    "$REF/include/rfid/global_vars.h" > "$Q4/rfid/global_vars.h"
grep -q 'FIXED_Q              = 4;' "$Q4/rfid/global_vars.h"
g++ $CXXFLAGS -I"$HERE/shim" -I"$Q4" -I"$REF/include" -I"$REF/lib" \
    $SRCS "$HERE/flow_driver.cc" -o "$OUT/libgen2ref_q4.so"
echo "built $OUT/libgen2ref.so $OUT/libgen2ref_q4.so"
