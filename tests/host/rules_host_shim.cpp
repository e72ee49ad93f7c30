// tests/host/rules_host_shim.cpp — TEST INFRASTRUCTURE: compiles the product's host+device rules header
// (alphazero-multi-game_b200/csrc/gomoku.cuh) with plain g++ so its bit-board logic can be checked against
// the oracle on the CPU box, before any GPU time is spent.  Not part of the product path.
#include "gomoku.cuh"
#include <cstring>

template <class G>
static int replay(const int* moves, int n, int* result, int* player, unsigned long long* key, float* planes, int* winner) {
    typename G::State s; G::init(s);
    for (int i = 0; i < n; ++i) { if (moves[i] < 0 || moves[i] >= G::CELLS || G::occupied(s, moves[i])) return -1; G::apply(s, moves[i]); }
    *result = G::result(s); *player = s.player; *key = G::key(s); *winner = G::winner(s);
    if (planes) for (int c = 0; c < G::PLANES; ++c) for (int x = 0; x < G::N; ++x) for (int y = 0; y < G::N; ++y)
        planes[(c * G::N + x) * G::N + y] = G::feature(s, c, x, y);
    return 0;
}
extern "C" int host_gomoku_replay(int n_board, const int* moves, int n, int* result, int* player, unsigned long long* key, float* planes, int* winner) {
    if (n_board == 15) return replay<az::Gomoku<15>>(moves, n, result, player, key, planes, winner);
    if (n_board == 9) return replay<az::Gomoku<9>>(moves, n, result, player, key, planes, winner);
    return -2;
}
