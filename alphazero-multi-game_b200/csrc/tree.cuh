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
    int16_t* legal;      // [T][MAX_CHILDREN] wide policy heads (chess): the leaf's legal actions in child order, by TREE; else nullptr
    int32_t* n_legal;    // [T]
    int32_t* slot_tree;  // [T] evaluation slot → the tree whose leaf it holds
    // in-wave evaluation dedup (M16, the TranspositionTable's role inside one wave): leaves of different trees with the SAME network input
    // share one evaluation.  dd_keys / dd_owner: open-addressing set of the wave's input keys [dd_mask + 1] (cleared every wave), owner =
    // the smallest tree index with that key; dd_idx[t] = the tree's entry.  nullptr = off.
    unsigned long long* dd_keys; int32_t* dd_owner; int32_t* dd_idx; unsigned int dd_mask;
    uint64_t* eval_key;  // [T] hash evaluators + EvalTT: the key the leaf is evaluated under (its own, or the first-seen position's); else nullptr
    int32_t* cache_entry; // [T] EvalCache entry that serves this tree's leaf (-1: none); nullptr = no cache
};

// Evaluation cache across waves (M16: what the reference's TranspositionTable is — 64-bit key -> (policy, value), transposition_table.cpp:44-84,
// 128-176; lookups parallel_mcts.cpp:320-336, 851).  Key = G::w_input_key, the WHOLE network input, so a hit returns exactly what the
// network would compute: the search is bit-identical with the cache on or off (policies are kept in fp32).  CACHE_WAYS-way buckets;
// `stamp` = wave number of the entry's last store or hit.  Probes happen in k_select, stores in k_expand_backup of the same wave: a
// store claims a way with atomicCAS(stamp, seen, wave) and only ways whose stamp is not the current wave can be claimed — so an entry
// that was hit (read by k_expand_backup) or stored in this wave is never overwritten in it; oldest stamp is the victim.  Which
// evaluations get stored depends on warp timing; what a hit returns does not.
constexpr int CACHE_WAYS = 4;
struct EvalCache {
    unsigned long long* keys;   // [cap], 0 = empty
    uint32_t* stamp;            // [cap]
    float* value;               // [cap]
    float* policy;              // [cap][pw]: action-indexed (pw = ACTIONS), or the children's raw priors in child order for LEGAL_POLICY games (pw = MAX_CHILDREN)
    unsigned int mask;          // cap - 1
    int pw;
    uint32_t* wave;             // device counters (the wave sequence is replayed as a CUDA graph: nothing per wave may be a kernel argument): [0] = number of the wave
                                // k_select is running (>= 1), [1] = number of the wave k_expand_backup is running; k_dedup_encode, between the two, copies [0] to [1] and
                                // advances [0]
};
__device__ __forceinline__ unsigned int cache_bucket(const EvalCache& ec, unsigned long long k) { return ((unsigned int)(mix64(k) >> 24) & ec.mask) & ~(unsigned int)(CACHE_WAYS - 1); }

// Model of the reference's per-game TranspositionTable (src/mcts/transposition_table.cpp:44-84, 128-176; lookups at
// parallel_mcts.cpp:320-336, 851) for games whose table key is coarser than the evaluator's input.  Chess: ChessState::getHash() covers
// the piece placement only (QUIRK C8), so a leaf whose placement was evaluated earlier in the game receives THAT position's policy /
// value, whatever the side to move, castling rights or e.p. square.  With a hash evaluator the cached result is a function of the
// first-seen position's evaluator key, so the table maps placement key -> evaluator key: [T][cap] open addressing, key 0 = empty,
// first store wins, one table per game (cleared when the slot starts a new game).  Gomoku / Go keys cover the whole evaluator
// input: their tables are result-transparent and not materialised.
struct EvalTT { uint64_t* keys; uint64_t* vals; int32_t* count; int32_t cap; };

// AZ_EVAL_DUP_STATS (profiling only): how many leaf evaluations an evaluation cache could have skipped.  Two open-addressing key sets per
// key flavour — one cleared every wave ("another tree evaluates the same input in this wave"), one kept for the whole run ("this input was
// evaluated before, in any tree, any game") — for (a) the exact network input and (b) the reference TranspositionTable's key.
struct DupStats {
    unsigned long long* wave_keys; unsigned long long* run_keys; unsigned long long* wave_keys_ref; unsigned long long* run_keys_ref;
    unsigned int wave_mask, run_mask;
    unsigned long long* counters;      // [0] evaluations, [1] duplicate of the wave (input), [2] seen before in the run (input), [3] / [4] the same for the reference key, [5] run-table inserts refused (full)
};
__device__ __forceinline__ bool dup_probe_insert(unsigned long long* keys, unsigned int mask, unsigned long long k, bool* full) {
    if (k == 0) k = 1;
    unsigned int i = (unsigned int)mix64(k) & mask;
    for (int probe = 0; probe < 64; ++probe, i = (i + 1) & mask) {
        const unsigned long long old = atomicCAS(&keys[i], 0ULL, k);
        if (old == 0ULL) return false;          // inserted: not seen before
        if (old == k) return true;
    }
    if (full) *full = true;
    return false;
}

struct SearchParams {
    float c_puct;        // MCTSConfig::cPuct (1.5)
    int virtual_loss;    // MCTSConfig::virtualLoss (3)
    int max_depth;
};

struct Stats {           // mirrors mcts::MCTSStats (parallel_mcts.h:77-99) + engine counters
    unsigned long long simulations, evaluations, terminal_leaves, nodes_created, nodes_expanded,
        pool_overflows, moves, games, samples_dropped, eval_shared,   // eval_shared: leaf evaluations served by another tree's evaluation of the same input in the same wave
        eval_cached;                                                  // ... served by the evaluation cache (an earlier wave's evaluation of the same input)
};

}  // namespace az
