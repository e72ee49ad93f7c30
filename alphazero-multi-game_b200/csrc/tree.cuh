// tree.cuh — device data layout of the batched search: one independent game tree per slot, all trees
// in structure-of-arrays node pools in HBM, searched one simulation per tree per wave.
//
// Replaces the reference's heap MCTSNode graph (include/alphazero/mcts/mcts_node.h:54-75: atomics, a mutex,
// two std::vectors per node) with flat arrays.  A node is identified by its index inside its tree's pool;
// all children of an expanded node are materialised at expansion (reference expandNodeWithPolicy,
// src/mcts/parallel_mcts.cpp:727-740) as one contiguous block [first, first+nchild) in legal-move order,
// so "child i" of the reference is node first+i here.
#pragma once
#include "common.cuh"

namespace az {

enum : uint8_t { NF_TERMINAL = 1, NF_RESULT_SHIFT = 1 /* bits 1-2: GameResult */ };
enum : int8_t { LEAF_EVAL = 0, LEAF_TERMINAL = 1, LEAF_NONE = 2 };
enum : uint8_t { TF_ACTIVE = 1, TF_FIRST_FILL = 2, TF_OVERFLOW = 4, TF_GAME_OVER = 8, TF_NEED_NOISE = 16 };

constexpr int MAX_DEPTH = 256;   // path buffer length (reference maxSearchDepth is 1000; Gomoku <= 225)

// Node storage: ONE pool of `pool_nodes` nodes shared by all trees.  Tree t owns the region [base[t], base[t] + limit[t]); node indices
// are relative to base[t].  The reference's trees are unbounded heap graphs (parallel_mcts.cpp:727-740, 1065-1108): a tree keeps the
// chosen child's subtree across moves, so its size depends on how peaked the policy is.  Regions are therefore re-cut at every move
// commit (k_region_plan): region = kept subtree + (simulations + 1) x MAX_CHILDREN nodes of growth + an equal share of the slack, and
// the kept subtrees are copied breadth-first into the other of two pool buffers (k_reroot_copy) — a tree with a large kept subtree
// borrows room from the others, and an expansion can only fail when the WHOLE pool is exhausted (reported as an error, never silent).
struct TreePools {
    // node arrays, [pool_nodes]
    int32_t* N;        // visitCount (for the root this includes the leaked virtual loss, QUIRK M7)
    float* W;          // valueSum
    float* P;          // prior
    int32_t* first;    // index of child 0, -1 = not expanded
    int32_t* sub;      // number of descendants (nodes below this one); maintained at expansion, gives the kept-subtree size at re-root
    int16_t* act;      // action leading to this node (reference action index; Go pass = -1)
    int16_t* nchild;
    uint8_t* flags;    // NF_TERMINAL | result << 1
    // per tree, [T]
    int64_t* base;     // first node of the tree's region
    int32_t* limit;    // nodes in the region
    int32_t* root;
    int32_t* alloc;    // bump pointer = number of nodes in use
    int32_t* root_vl;  // virtualLoss currently parked on the root (reference leaks 3 per simulation)
    uint8_t* tflags;
    int32_t* move_num; // moves played in the current game (SelfPlayManager moveNum)
    uint32_t* game_id; // generation counter per slot
};

struct WaveBuffers {
    int32_t* path;       // [T][MAX_DEPTH] node indices root..leaf
    int32_t* path_len;   // [T] number of entries in path (0 = nothing to back up: root-expansion wave)
    int32_t* leaf_node;  // [T]
    int8_t* leaf_kind;   // [T]
    float* leaf_value;   // [T] terminal value (LEAF_TERMINAL)
    float* policy;       // [T][A] evaluator output (probabilities over the action space)
    float* value;        // [T]
    int32_t* eval_slot;  // [T] compacted NN batch index of this tree's leaf, -1 = none
    int32_t* n_eval;     // [1] number of leaves that need an evaluation this wave
    uint64_t* eval_key;  // [T] hash evaluators + EvalTT: the key the leaf is evaluated under (its own, or the first-seen position's); else nullptr
};

// Model of the reference's per-game TranspositionTable (src/mcts/transposition_table.cpp:44-84, 128-176; lookups at
// parallel_mcts.cpp:320-336, 851) for games whose table key is coarser than the evaluator's input.  Chess: ChessState::getHash() covers
// the piece placement only (QUIRK C8), so a leaf whose placement was evaluated earlier in the game receives THAT position's policy /
// value, whatever the side to move, castling rights or e.p. square.  With a hash evaluator the cached result is a function of the
// first-seen position's evaluator key, so the table maps placement key -> evaluator key: [T][cap] open addressing, key 0 = empty,
// first store wins, one table per game (cleared when the slot starts a new game).  Gomoku / Go keys cover the whole evaluator
// input: their tables are result-transparent and not materialised.
struct EvalTT { uint64_t* keys; uint64_t* vals; int32_t* count; int32_t cap; };

struct SearchParams {
    float c_puct;        // MCTSConfig::cPuct (1.5)
    int virtual_loss;    // MCTSConfig::virtualLoss (3)
    int max_depth;
};

struct Stats {           // mirrors mcts::MCTSStats (parallel_mcts.h:77-99) + engine counters
    unsigned long long simulations, evaluations, terminal_leaves, nodes_created, nodes_expanded,
        pool_overflows, moves, games, samples_dropped;
};

}  // namespace az
