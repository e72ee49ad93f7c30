#!/usr/bin/env bash
# oracle/build_ref.sh — TEST INFRASTRUCTURE.
# Builds oracle/_ref/libaz_ref.so from the reference's own sources (read in place from
# $AZ_REFERENCE, default /root/reference) + oracle/ref_harness.cpp.
#
# The reference at HEAD neither compiles nor terminates, so the sources are
# copied to a throw-away temp dir OUTSIDE the repo, the §8c patch shim is applied there with
# sed, the library is linked into oracle/_ref/ and the temp dir is deleted.  No reference
# source ever lands in the repo; oracle/_ref/ is git-ignored (it still travels to the GPU box).
#
# Patch shim (each line is one sed below):
#  1. zobrist_hash.h: delegating ctor ZobristHash(GameType,int boardSize,int numPieceTypes,unsigned seed=0)
#     (callers: gomoku_state.cpp:32, go_state.cpp:21 pass (GameType, boardSize, numPieceTypes))
#  2. parallel_mcts.cpp:1557 lambda: capture numThreads
#  3. parallel_mcts.cpp:683 expandNodeWithPolicy: drop the inner lock_guard (callers hold the same
#     non-recursive mutex → self-deadlock on the first search())
#  4. parallel_mcts.h: public read-only accessor getRootNode() (harness needs root child stats)
#  5. chess sources are compiled with -include alphazero/games/chess/chess_state.h (chess_rules.h:165 uses
#     PieceColor::WHITE on an enum it only forward-declares at :19)
#  6. chess_state.cpp: ChessState::cloneWithMove (:1232-1236) applies the move WITHOUT the legality re-check.  Unpatched,
#     makeMove -> isLegalMove -> ChessRules::isLegalMove -> moveExposesKing (chess_rules.cpp:745-751) -> cloneWithMove ->
#     makeMove recurses without end on the very first getLegalMoves().
#  7. registry.cpp: #include <mutex> (game_factory.cpp's createGameState needs GameRegistry)
# Also built: src/selfplay/{dataset,game_record}.cpp + src/core/game_factory.cpp (createGameState) — unpatched; nlohmann json
# comes from site-packages (cudnn_frontend/thirdparty).
# Float environment: plain x86-64 (no -march=native, so no FMA contraction), -ffp-contract=off.
set -euo pipefail
REF="${AZ_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/src/mcts" ]; then
  echo "build_ref: reference not present at $REF — keeping any prebuilt $OUT/libaz_ref.so" >&2
  exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d /tmp/az_ref_build.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$TMP/src" "$TMP/include"
cp -r "$REF/include/alphazero" "$TMP/include/"
mkdir -p "$TMP/src/core" "$TMP/src/games" "$TMP/src/mcts" "$TMP/src/nn" "$TMP/src/selfplay"
cp "$REF/src/core/zobrist_hash.cpp" "$REF/src/core/game_factory.cpp" "$REF/src/core/registry.cpp" "$TMP/src/core/"
cp -r "$REF/src/games/gomoku" "$REF/src/games/go" "$REF/src/games/chess" "$TMP/src/games/"
cp "$REF/src/selfplay/dataset.cpp" "$REF/src/selfplay/game_record.cpp" "$TMP/src/selfplay/"
cp "$REF/src/mcts/parallel_mcts.cpp" "$REF/src/mcts/mcts_node.cpp" \
   "$REF/src/mcts/transposition_table.cpp" "$REF/src/mcts/thread_pool.cpp" "$TMP/src/mcts/"
cp "$REF/src/nn/neural_network.cpp" "$REF/src/nn/batch_queue.cpp" \
   "$REF/src/nn/random_policy_network.cpp" "$TMP/src/nn/"

# --- patch shim -------------------------------------------------------------------------------
# 1
sed -i 's|^    ZobristHash(int boardSize, int numPieceTypes, int numPlayers, unsigned seed = 0);|&\n    ZobristHash(GameType, int boardSize, int numPieceTypes, unsigned seed = 0) : ZobristHash(boardSize, numPieceTypes, 2, seed) {}|' \
  "$TMP/include/alphazero/core/zobrist_hash.h"
grep -q "ZobristHash(GameType" "$TMP/include/alphazero/core/zobrist_hash.h"
# 2
sed -i '1557s|\[this, i, &completedSimulations\]|[this, i, numThreads, \&completedSimulations]|' "$TMP/src/mcts/parallel_mcts.cpp"
# 3 (the lock_guard directly after the expandNodeWithPolicy signature)
python3 - "$TMP/src/mcts/parallel_mcts.cpp" <<'EOF'
import sys, re
p = sys.argv[1]; s = open(p).read()
sig = "void ParallelMCTS::expandNodeWithPolicy("
i = s.index(sig)
j = s.index("std::lock_guard<std::mutex> lock(node->expansionMutex);", i)
assert j - i < 400, "unexpected layout"
s = s[:j] + "/* shim: callers already hold expansionMutex */" + s[j + len("std::lock_guard<std::mutex> lock(node->expansionMutex);"):]
open(p, "w").write(s)
EOF
# 4
sed -i 's|^    std::unique_ptr<MCTSNode> rootNode_;|&\n  public: const MCTSNode* getRootNode() const { return rootNode_.get(); }\n  private:|' \
  "$TMP/include/alphazero/mcts/parallel_mcts.h"
grep -q "getRootNode" "$TMP/include/alphazero/mcts/parallel_mcts.h"

# 7 (src/core/registry.cpp:8 uses std::unique_lock without <mutex>)
sed -i '1i #include <mutex>' "$TMP/src/core/registry.cpp"
# 6
python3 - "$TMP/src/games/chess/chess_state.cpp" <<'EOF'
import sys
p = sys.argv[1]; s = open(p).read()
a = """void ChessState::makeMove(const ChessMove& move) {
    if (!isLegalMove(move)) {"""
b = """    ChessState newState(*this);
    newState.makeMove(move);
    return newState;"""
assert s.count(a) == 1 and s.count(b) == 1, "unexpected layout"
s = s.replace(a, """static thread_local int az_shim_unchecked = 0;   /* shim 6 */
void ChessState::makeMove(const ChessMove& move) {
    if (!az_shim_unchecked && !isLegalMove(move)) {""")
s = s.replace(b, """    ChessState newState(*this);
    ++az_shim_unchecked; newState.makeMove(move); --az_shim_unchecked;   /* shim 6: no legality re-check inside moveExposesKing */
    return newState;""")
open(p, "w").write(s)
EOF

JSONINC="$(python3 -c 'import os, sysconfig; print(os.path.join(sysconfig.get_paths()["purelib"], "include", "cudnn_frontend", "thirdparty"))')"
[ -f "$JSONINC/nlohmann/json.hpp" ] || { echo "build_ref: nlohmann/json.hpp not found under $JSONINC" >&2; exit 1; }
CXX="${CXX:-g++}"
FLAGS="-std=c++17 -O2 -fPIC -ffp-contract=off -DLIBTORCH_OFF -w -I$TMP/include -I$JSONINC -pthread"
SRCS="$(find "$TMP/src" -name '*.cpp' | sort)"
OBJS=""
for f in $SRCS; do
  o="$TMP/$(echo "$f" | sed "s|$TMP/||; s|/|_|g").o"
  extra=""; case "$f" in */games/chess/*) extra="-include alphazero/games/chess/chess_state.h";; esac   # 5
  $CXX $FLAGS $extra -c "$f" -o "$o" &
  OBJS="$OBJS $o"
done
$CXX $FLAGS -c "$HERE/ref_harness.cpp" -o "$TMP/ref_harness.o" &
wait
$CXX -shared -Wl,-z,defs -o "$OUT/libaz_ref.so" $OBJS "$TMP/ref_harness.o" -pthread
echo "build_ref: wrote $OUT/libaz_ref.so"

# The reference's own unit-test sources for the state classes (tests/{games,core}/*_test.cpp), read in place and compiled UNMODIFIED with the
# stand-in gtest of oracle/mini_gtest against the objects above: oracle/_ref/ref_state_tests prints one PASS / FAIL / CRASH line per test.
# tests/test_ref_unit_tests.py runs it next to the same sources compiled against the B200 host mirror.
UT="games/gomoku/gomoku_state_test.cpp games/go/go_state_test.cpp games/chess/chess_state_test.cpp core/igamestate_test.cpp core/game_factory_test.cpp"
UTOBJ=""
for t in $UT; do
  o="$TMP/ut_$(echo "$t" | sed 's|/|_|g').o"
  extra=""; case "$t" in games/chess/*) extra="-include alphazero/games/chess/chess_state.h";;
                         core/igamestate_test.cpp) extra="-include alphazero/core/game_factory.h";; esac   # the test calls createGameState, which igamestate.h does not declare
  $CXX $FLAGS -I"$HERE/mini_gtest" $extra -c "$REF/tests/$t" -o "$o" &
  UTOBJ="$UTOBJ $o"
done
$CXX $FLAGS -I"$HERE/mini_gtest" -c "$HERE/mini_gtest/main.cpp" -o "$TMP/ut_main.o" &
wait
$CXX -o "$OUT/ref_state_tests" $UTOBJ "$TMP/ut_main.o" $OBJS -pthread
echo "build_ref: wrote $OUT/ref_state_tests"
