#!/usr/bin/env bash
# oracle/run_mirror_unit_tests.sh — TEST INFRASTRUCTURE.
# Compiles the reference's own unit-test sources for the state classes (read in place from $AZ_REFERENCE/tests, UNMODIFIED) against the B200
# host mirror (alphazero-multi-game_b200/host/alphazero_host.{hpp,cpp}: same class names, CPU only — the state classes need no GPU) with
# the stand-in gtest of oracle/mini_gtest, and writes oracle/_ref/mirror_state_tests.  tests/test_ref_unit_tests.py runs it beside
# oracle/_ref/ref_state_tests (the same sources on the reference itself, oracle/build_ref.sh) and compares the verdicts test by test.
set -euo pipefail
REF="${AZ_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
HOST="$ROOT/alphazero-multi-game_b200/host"
OUT="$HERE/_ref"
[ -d "$REF/tests" ] || { echo "run_mirror_unit_tests: reference not present at $REF" >&2; exit 0; }
mkdir -p "$OUT"
TMP="$(mktemp -d /tmp/az_mirror_ut.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
CXX="${CXX:-g++}"
JSONINC="$(python3 -c 'import os, sysconfig; print(os.path.join(sysconfig.get_paths()["purelib"], "include", "cudnn_frontend", "thirdparty"))')"
FLAGS="-std=c++17 -O1 -w -fPIC -ffp-contract=off -I/usr/local/cuda/include -I$HERE/mini_gtest -I$HERE/mirror_shim_include -I$HOST -I$ROOT/include -I$JSONINC -pthread"
UT="games/gomoku/gomoku_state_test.cpp games/go/go_state_test.cpp games/chess/chess_state_test.cpp core/igamestate_test.cpp core/game_factory_test.cpp"
OBJ=""
for t in $UT; do
  o="$TMP/ut_$(echo "$t" | sed 's|/|_|g').o"
  $CXX $FLAGS -c "$REF/tests/$t" -o "$o" &
  OBJ="$OBJ $o"
done
$CXX $FLAGS -c "$HERE/mini_gtest/main.cpp" -o "$TMP/ut_main.o" &
wait
# the mirror itself: alphazero_host.cpp links against the C ABI (libaz_b200.so) for the search classes; the state classes do not call it
$CXX -o "$OUT/mirror_state_tests" $OBJ "$TMP/ut_main.o" "$HOST/alphazero_host.cpp" $FLAGS -L"$ROOT/alphazero-multi-game_b200" -laz_b200 -Wl,-rpath,"$ROOT/alphazero-multi-game_b200" -ldl
echo "run_mirror_unit_tests: wrote $OUT/mirror_state_tests"
