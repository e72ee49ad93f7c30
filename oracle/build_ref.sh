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
mkdir -p "$TMP/src/core" "$TMP/src/games" "$TMP/src/mcts" "$TMP/src/nn"
cp "$REF/src/core/zobrist_hash.cpp" "$TMP/src/core/"
cp -r "$REF/src/games/gomoku" "$REF/src/games/go" "$TMP/src/games/"
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

CXX="${CXX:-g++}"
FLAGS="-std=c++17 -O2 -fPIC -ffp-contract=off -DLIBTORCH_OFF -w -I$TMP/include -pthread"
SRCS="$(find "$TMP/src" -name '*.cpp' | sort)"
OBJS=""
for f in $SRCS; do
  o="$TMP/$(echo "$f" | sed "s|$TMP/||; s|/|_|g").o"
  $CXX $FLAGS -c "$f" -o "$o" &
  OBJS="$OBJS $o"
done
$CXX $FLAGS -c "$HERE/ref_harness.cpp" -o "$TMP/ref_harness.o" &
wait
$CXX -shared -o "$OUT/libaz_ref.so" $OBJS "$TMP/ref_harness.o" -pthread
echo "build_ref: wrote $OUT/libaz_ref.so"
