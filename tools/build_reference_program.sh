#!/bin/sh
# Builds the reference's OWN main program (test.cpp: batch_input_test, ct_pt_matrix_mul_test, ct_ct_matrix_mul_test,
# all_layer_test) UNMODIFIED against the B200 backend: the header-only seal:: facade (include/facade), the fused module
# headers (include/facade_fused) and libmoai_b200.so.  No Bootstrapper.cpp / NTL needed.
#   usage: tools/build_reference_program.sh [/path/to/MOAI checkout] [output binary]
#   MOAI_FUSED=0 keeps the reference's own module headers (one kernel launch per Evaluator call).
set -e
REPO="$(cd "$(dirname "$0")/.." && pwd)"
MOAI="${1:-/root/reference}"
OUT="${2:-$REPO/moai_b200_reference_test}"
PKG="$REPO/moai-fhe-transformerinference-public_b200"
FUSED="-I$REPO/include/facade_fused"
[ "${MOAI_FUSED:-1}" = "0" ] && FUSED=""
[ -f "$PKG/libmoai_b200.so" ] || python "$REPO/__graft_entry__.py"
g++ -std=c++17 -O2 -fopenmp -w $FUSED -I"$REPO/include/facade" -I"$REPO/include" -I"$MOAI/include" "$MOAI/test.cpp" \
    -L"$PKG" -lmoai_b200 -Wl,-rpath,"$PKG" -o "$OUT"
echo "$OUT  (run it from the MOAI checkout: all_layer_test reads data/ by relative path)"
