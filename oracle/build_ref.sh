#!/bin/bash
# TEST INFRASTRUCTURE. Compiles the reference's self-contained TPC-H dbgen (velox/tpch/gen/dbgen/*.cpp)
# from the sources where they lie under /root/reference into oracle/_ref/libtpchref.so, together
# with oracle/dbgen_wrap.cpp. No reference source is copied into this repository; the reference's own
# build system (cmake, folly ...) is not used. oracle/_ref/ is git-ignored but travels with gpurun.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REFERENCE_ROOT:-/root/reference}"
GEN="$REF/velox/tpch/gen/dbgen"
[ -d "$GEN" ] || { echo "reference dbgen not found at $GEN (nothing to build)"; exit 0; }
mkdir -p "$HERE/_ref"
OUT="$HERE/_ref/libtpchref.so"
if [ -f "$OUT" ] && [ "$OUT" -nt "$HERE/dbgen_wrap.cpp" ]; then exit 0; fi
g++ -O2 -std=c++20 -fPIC -shared -w -I"$HERE/ref_shim" -I"$REF" -I"$GEN/include" \
    "$GEN"/bm_utils.cpp "$GEN"/build.cpp "$GEN"/dbgen.cpp "$GEN"/dbgen_gunk.cpp "$GEN"/permute.cpp "$GEN"/rnd.cpp \
    "$GEN"/rng64.cpp "$GEN"/speed_seed.cpp "$GEN"/text.cpp "$HERE/dbgen_wrap.cpp" -o "$OUT"
echo "built $OUT"
