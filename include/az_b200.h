/* az_b200.h — C ABI of the B200-native batched self-play engine (libaz_b200.so).
 *
 * This is the drop-in boundary #3 of SURVEY.md §8(b): plain C, pointers + sizes, no torch / STL types.
 * Every entry point names the reference interface it stands in for (paths relative to the reference repo).
 * The closest precedent in the reference is its plugin ABI (include/alphazero/core/plugin_api.h:19-37).
 *
 * Conventions: every function returns 0 on success and a negative status on failure (az_last_error() gives
 * the message); no exception crosses the boundary; the caller owns all host buffers, the engine owns all
 * device memory; calls on one engine must be serialised by the caller (one host thread per engine / GPU).
 * There is NO CPU fallback: az_engine_create fails when no sm_100-class CUDA device is present.
 */
#ifndef AZ_B200_H
#define AZ_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define AZ_API __attribute__((visibility("default")))
#else
#define AZ_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* core::GameType, include/alphazero/core/igamestate.h:16-20 */
enum { AZ_GAME_GOMOKU = 0, AZ_GAME_CHESS = 1, AZ_GAME_GO = 2 };
/* core::GameResult, include/alphazero/core/igamestate.h:25-30 */
enum { AZ_ONGOING = 0, AZ_DRAW = 1, AZ_WIN_PLAYER1 = 2, AZ_WIN_PLAYER2 = 3 };
/* evaluator behind nn::NeuralNetwork::predict (include/alphazero/nn/neural_network.h:29) */
enum { AZ_EVAL_HASH = 0,   /* stateless integer-mix evaluator (parity runs) */
       AZ_EVAL_RESNET = 1, /* policy/value ResNet, 16-bit operands on tcgen05 tensor cores (kind::f16), fp32 accumulation */
       AZ_EVAL_HASH_PEAKED = 2, /* AZ_EVAL_HASH with one action's raw prior x 4096 (a peaked policy: deep, narrow trees; node-pool stress / parity) */
       AZ_EVAL_EXTERNAL = 3 /* the caller evaluates the leaves on the host (az_engine_set_external_evaluator): any nn::NeuralNetwork implementation
                               behind ParallelMCTS; one host round trip per wave — the compatibility path, not the fast one */ };
/* 16-bit storage type of the network's activations and conv weights.  nn::TorchNeuralNetworkConfig::useFp16
 * (include/alphazero/nn/torch_neural_network.h:29: the reference's own reduced-precision mode is fp16, model_.to(torch::kHalf),
 * src/nn/torch_neural_network.cpp:183,267,409).  fp16 and bf16 run the same tcgen05.mma.kind::f16 instruction at the same rate; fp16
 * carries 11 significand bits instead of 8, which is what the policy KL <= 1e-3 tolerance needs on the BASELINE network
 * (profiles/r2_kl_rounding_experiment.md).  Stores saturate at +-65504 instead of overflowing. */
enum { AZ_NET_FP16 = 0, AZ_NET_BF16 = 1 };

/* mcts::MCTSConfig (include/alphazero/mcts/parallel_mcts.h:41-74) + SelfPlayManager exploration params
 * (src/selfplay/self_play_manager.cpp:16-37) + engine sizing. */
typedef struct az_config {
    int32_t game;               /* AZ_GAME_* */
    int32_t board_size;         /* 15 (Gomoku) */
    int32_t n_slots;            /* concurrent games = trees searched per wave */
    int32_t num_simulations;    /* MCTSConfig::numSimulations (800) */
    float   c_puct;             /* MCTSConfig::cPuct (1.5) */
    int32_t virtual_loss;       /* MCTSConfig::virtualLoss (3) */
    int32_t evaluator;          /* AZ_EVAL_* */
    int32_t net_blocks;         /* residual blocks (10) */
    int32_t net_channels;       /* trunk width (128) */
    int32_t max_nodes_per_tree; /* node-pool capacity per slot; 0 = default */
    int32_t deterministic;      /* 1 = noise off, first-max-visit move (the reference's forced useBatchInference
                                   branch, parallel_mcts.cpp:1018-1021,1037-1039); 0 = Dirichlet noise + temperature */
    float   dirichlet_alpha;    /* 0.03 */
    float   dirichlet_epsilon;  /* 0.25 */
    float   init_temperature;   /* 1.0 */
    float   final_temperature;  /* 0.0 */
    int32_t temperature_drop_move; /* 30 */
    int32_t auto_restart;       /* 1 = a finished game's slot starts a new game (self-play); 0 = slot goes idle */
    int32_t sample_ring_capacity; /* finished-game sample records held on device until drained; 0 = default */
    int32_t device;             /* CUDA device ordinal */
    uint64_t seed;              /* Philox key for noise / temperature sampling */
    int32_t n_streams;          /* stream groups the slots are split into (tree kernels of one group overlap the
                                   network pass of another); 0 = default (1: measured no gain under the 1 kW power cap) */
    int32_t net_precision;      /* AZ_NET_FP16 (default) | AZ_NET_BF16 */
    int32_t tt_entries;         /* chess + hash evaluators only: per-tree capacity of the model of the reference's TranspositionTable
                                   (keyed by the piece placement, QUIRK C8: a leaf whose placement was evaluated before in the game gets that
                                   position's evaluation, src/mcts/parallel_mcts.cpp:320-336).  0 = default (on, sized from num_simulations),
                                   -1 = off (every leaf evaluated on its own input) */
    int32_t eval_dedup;         /* ResNet evaluator: 0 = default (on): leaves of different trees that present the SAME network input in the same wave
                                   share one evaluation (the role of the reference's TranspositionTable — an evaluation cache, M16 — inside one wave;
                                   result-transparent: bit-identical searches with it on or off); -1 = off */
    int32_t eval_cache_entries; /* evaluation cache ACROSS waves (the reference's TranspositionTable, src/mcts/transposition_table.cpp:44-84, 128-176: 64-bit
                                   key -> (policy, value)): key = the whole network input, fp32 policies, 4-way buckets, oldest entry replaced — a hit
                                   returns exactly what the network would compute, so searches are bit-identical with it on or off.  0 = default (ResNet
                                   evaluator: four moves' worth of evaluations of all slots, 64 K ... 4 M entries; hash evaluators: off), > 0 = entries (rounded down to a power of two), -1 = off.  Needs
                                   eval_dedup >= 0.  Cleared by az_engine_load_weights */
    int32_t dense_policy;       /* wide policy heads (chess, 20480 actions): 0 = inside the waves the network computes the logits of the leaf's
                                   legal moves only (their softmax equals the full softmax renormalised over the legal moves, which is what the
                                   expansion computes, parallel_mcts.cpp:705-724); 1 = all A logits + the full softmax, as az_engine_nn_forward does */
} az_config;

/* mcts::MCTSStats (include/alphazero/mcts/parallel_mcts.h:77-99) + engine counters, cumulative */
typedef struct az_stats {
    uint64_t simulations, evaluations, terminal_leaves, nodes_created, nodes_expanded, pool_overflows,
             moves, games, samples_dropped;
    uint64_t kernel_launches;   /* CUDA kernels this engine has launched */
    uint64_t waves;
    uint64_t eval_shared;       /* leaf evaluations served by another tree's evaluation of the same network input in the same wave (MCTSStats::cacheHits);
                                   network evaluations actually run = evaluations - eval_shared - eval_cached */
    uint64_t eval_cached;       /* leaf evaluations served by the evaluation cache (an earlier wave's evaluation of the same network input) */
} az_stats;

typedef struct az_engine az_engine;

AZ_API void az_config_default(az_config* cfg);
AZ_API const char* az_last_error(void);

/* ParallelMCTS ctor + SelfPlayManager ctor (src/mcts/parallel_mcts.cpp:44-96, src/selfplay/self_play_manager.cpp:16-37) */
AZ_API int az_engine_create(const az_config* cfg, az_engine** out);
AZ_API int az_engine_destroy(az_engine* e);

/* TorchNeuralNetwork model load (src/nn/torch_neural_network.cpp:90): `blob` is the AZW1 fp32 weight file
 * written by alphazero-multi-game_b200/net.py:export_weights; BatchNorm is folded and bf16 images built here. */
AZ_API int az_engine_load_weights(az_engine* e, const void* blob, size_t bytes);

/* createGameState + ParallelMCTS(rootState,...) for every slot: empty boards, fresh trees
 * (src/selfplay/self_play_manager.cpp:157-175) */
AZ_API int az_engine_reset_games(az_engine* e);
/* ParallelMCTS::initialize(rootState) for one slot: root = empty board + `moves`.  `first_fill_order`
 * (may be NULL → descending order) is the legal-move order of the root's first expansion (QUIRK G2). */
AZ_API int az_engine_set_root(az_engine* e, int slot, const int32_t* moves, int n_moves,
                       const int32_t* first_fill_order, int n_order);
/* ParallelMCTS::search() on every active slot: root expansion if needed, then `sims` simulations per tree
 * as `sims` waves (src/mcts/parallel_mcts.cpp:142-274, serial semantics per tree). sims<=0 → config value. */
AZ_API int az_engine_search(az_engine* e, int sims);
/* root children in child order (rootNode_->children[i]): action, visitCount, valueSum, prior;
 * *n_children in: capacity, out: count.  root_visits/root_value_sum are the root node's own fields. */
AZ_API int az_engine_root_stats(az_engine* e, int slot, int32_t* actions, int32_t* visits, float* value_sums,
                         float* priors, int32_t* n_children, int32_t* root_visits, float* root_value_sum);
/* ParallelMCTS::updateWithMove(action) per slot (src/mcts/parallel_mcts.cpp:1065-1108); actions[slot] = -2 skips a slot */
AZ_API int az_engine_advance(az_engine* e, const int32_t* actions, int n);
/* ParallelMCTS::addDirichletNoise(alpha, epsilon) on every active slot's root (expanding it first if needed),
 * src/mcts/parallel_mcts.cpp:1110-1171 */
AZ_API int az_engine_add_dirichlet_noise(az_engine* e, float alpha, float epsilon);
/* SelfPlayManager::playSingleGame loop body for all slots, `n_moves` times: search → getActionProbabilities /
 * selectAction / getRootValue → record → makeMove → updateWithMove → noise (src/selfplay/self_play_manager.cpp:187-217) */
AZ_API int az_engine_play(az_engine* e, int n_moves);
/* action chosen by the last az_engine_play / az_engine_advance per slot (-2 = slot did not move) */
AZ_API int az_engine_last_actions(az_engine* e, int32_t* actions, int n);
/* per-slot state: result (AZ_*), ply, player to move */
AZ_API int az_engine_slot_state(az_engine* e, int slot, int32_t* result, int32_t* ply, int32_t* player);

/* GameRecord/MoveData stream (include/alphazero/selfplay/game_record.h:18-70): finished-game samples.
 * Record layout: az_engine_sample_layout.  drain copies up to cap records to host and empties the ring. */
typedef struct az_sample_layout {
    int32_t record_bytes, off_game_id, off_slot, off_ply, off_action, off_player, off_z, off_result,
            off_root_value, off_root_visits, off_state, state_bytes, off_visits, n_visits;
} az_sample_layout;
AZ_API int az_engine_sample_layout(az_engine* e, az_sample_layout* out);
AZ_API int az_engine_drain_samples(az_engine* e, void* host_buf, size_t cap_records, size_t* n_records);
/* device-side variant for the NCCL all-gather: packs the ring into caller-provided DEVICE memory */
AZ_API int az_engine_drain_samples_device(az_engine* e, void* dev_buf, size_t cap_records, size_t* n_records);

/* Dataset::extractExamples + Dataset::augmentExample (src/selfplay/dataset.cpp:64-114, 245-436) on the device: turns drained
 * sample records into training tensors.  Per record k = (augment && game != chess) ? 8 : 1 examples in the reference's order
 * (original, rot90, rot180, rot270, flipH, flipH(rot90), flipH(rot180), flipH(rot270)); all planes and the first N*N policy
 * entries are moved by the same map, entries past N*N (Go's pass) stay.  planes fp32 [n*k][C][N][N], policy fp32 [n*k][A]
 * = visit counts / their sum by ACTION (getVisitCountDistribution at temperature 1, mcts_node.cpp:289-322; the reference
 * stores it by child index), value fp32 [n*k] = game result seen from the player to move.  Host buffers. */
AZ_API int az_engine_make_examples(az_engine* e, const void* samples, size_t n_records, int augment, float* planes, float* policy, float* value);

/* Dataset::extractExamples (src/selfplay/dataset.cpp:64-114) + augmentExample (:245-436) for GameRecords that come through the host
 * API as move lists (GameRecord::getMoves, include/alphazero/selfplay/game_record.h): moves [n_games][max_moves] (reference action
 * codes; Go pass = -1), n_moves[g], results[g] = GameResult code (0 ONGOING, 1 DRAW, 2 WIN_PLAYER1, 3 WIN_PLAYER2).  One example per
 * recorded move i of game g, in game order: planes of the state after moves[g][0..i) (getEnhancedTensorRepresentation), the
 * caller's policy vector for that move (policy_in [sum n_moves][policy_len], copied / permuted exactly as the reference does: an
 * entry moves only when its old and new index are both < policy_len), value = result seen from the player to move (:84-96).
 * augment != 0 and the game is not chess: the 8 images per example (original, rot90, rot180, rot270, flipH, flipH of the rotations).
 * Outputs (host): planes fp32 [n*k][C][N][N], policy fp32 [n*k][policy_len], value fp32 [n*k].  An illegal recorded move is an error
 * (the reference's makeMove throws). */
AZ_API int az_engine_examples_from_games(az_engine* e, const int32_t* moves, const int32_t* n_moves, const int8_t* results, int n_games, int max_moves,
                                         const float* policy_in, int policy_len, int augment, float* planes, float* policy, float* value);

/* nn::NeuralNetwork::predictBatch (include/alphazero/nn/neural_network.h:38-42) supplied by the CALLER, for engines created with
 * AZ_EVAL_EXTERNAL: once per wave the engine hands over the n leaves that need an evaluation as move sequences from their slots' current
 * roots — slot[i], path_len[i] actions at path_actions + i * max_len (reference action codes; 0 actions = the root itself) — and expects
 * policy[i * actions .. ] (probabilities over the action space, what predict returns) and value[i].  Called on the thread that calls
 * az_engine_search / az_engine_play; a non-zero return aborts the search with an error. */
typedef int (*az_eval_fn)(int n, const int32_t* slot, const int32_t* path_actions, const int32_t* path_len, int max_len, int actions,
                          float* policy, float* value, void* user);
AZ_API int az_engine_set_external_evaluator(az_engine* e, az_eval_fn fn, void* user);
AZ_API int az_engine_get_stats(az_engine* e, az_stats* out);
AZ_API int az_engine_sync(az_engine* e);
/* Where the time of a step goes (no reference counterpart; MCTSStats has only counters): every 64th wave of a ResNet engine is bracketed
 * kernel by kernel with CUDA events on its stream, every move commit of az_engine_play as a whole.  Sums in ms over the sampled waves /
 * moves: a step of S simulations costs about (S + 1) x (select + dedup_encode + evaluator + expand_backup) / waves_sampled + commit / moves_sampled.
 * evaluator = stem + trunk + head_conv + conv1x1_gemm + policy_fc + value_fc + policy_value. */
typedef struct az_timing {
    uint64_t waves_sampled, moves_sampled;
    double select_ms, dedup_encode_ms, evaluator_ms, expand_backup_ms, commit_ms;
    double stem_ms, trunk_ms, head_conv_ms, conv1x1_gemm_ms, policy_fc_ms, value_fc_ms, policy_value_ms;
} az_timing;
AZ_API int az_engine_get_timing(az_engine* e, az_timing* out);
/* ParallelMCTS::setCPuct / setVirtualLoss / setConfig (src/mcts/parallel_mcts.cpp:1173-1261): take effect from the next search */
AZ_API int az_engine_set_search_params(az_engine* e, float c_puct, int virtual_loss);
/* ParallelMCTS::setNumSimulations (src/mcts/parallel_mcts.cpp:1183-1185) / SelfPlayManager's numSimulations: simulations per move of az_engine_play and the
 * default of az_engine_search, from the next search on.  The node pool is sized at creation: counts above az_config.num_simulations need a
 * max_nodes_per_tree that holds (num_simulations + 1) x max-children nodes per tree, else the call fails */
AZ_API int az_engine_set_num_simulations(az_engine* e, int num_simulations);
/* MCTSNode::children / actions / visitCount / valueSum / prior of ANY node (include/alphazero/mcts/mcts_node.h:54-75; what
 * ParallelMCTS::printSearchPath walks, parallel_mcts.cpp:1390-1450): the node reached from the slot's root by the action sequence
 * `path` (n_path = 0: the root).  Children in child order; *n_children in: capacity, out: count; node_* = the node's own fields,
 * node_flags bit 0 = terminal, bits 1-2 = GameResult. */
AZ_API int az_engine_node_stats(az_engine* e, int slot, const int32_t* path, int n_path, int32_t* actions, int32_t* visits, float* value_sums,
                                float* priors, int32_t* n_children, int32_t* node_visits, float* node_value_sum, float* node_prior, int32_t* node_flags);

/* Device memory for the host layer's multi-GPU exchange (SelfPlayManager over several GPUs: finished-game samples are drained on each
 * device with az_engine_drain_samples_device, all-gathered with ncclAllGather and read from one device).  Nothing in the
 * reference corresponds to these (it has no GPU memory of its own); they keep the host layer free of a CUDA runtime dependency. */
AZ_API int az_device_count(int* n);
AZ_API int az_device_alloc(int device, size_t bytes, void** out);
AZ_API int az_device_free(int device, void* p);
AZ_API int az_device_memcpy(int device, void* dst, const void* src, size_t bytes, int to_host);
AZ_API int az_device_sync(int device);

/* NeuralNetwork::predictBatch (src/nn/torch_neural_network.cpp:224-363) on caller-supplied feature planes:
 * planes fp32 [n][C][H][W] (host) → policy fp32 [n][A] (softmax over the A logits, :298-316), value fp32 [n]. */
AZ_API int az_engine_nn_forward(az_engine* e, const float* planes, int n, float* policy, float* value, float* logits_or_null);
/* same forward on already-resident device inputs in the trunk's layout, `reps` times (kernel timing) */
AZ_API int az_engine_nn_bench(az_engine* e, int n_boards, int reps, float* ms_per_rep);

/* one 128→128-channel 3x3 conv layer (the dominant kernel) alone, `reps` launches, CUDA-event timed on the
 * engine's stream: ms per launch (roofline numerator for bench.py) */
AZ_API int az_engine_conv_bench(az_engine* e, int n_boards, int reps, float* ms_per_launch);
/* the same kernel timed LIVE inside production waves: every 64th network pass of az_engine_search / az_engine_play brackets its
 * 128->128 conv launches with CUDA events on the launching stream; returns the accumulated ms and the number of launches covered */
AZ_API int az_engine_conv_sampled(az_engine* e, double* ms_sum, unsigned long long* n_launches);
/* CUDA events on the engine's own stream (torch.cuda.Event only sees torch's stream): record slot idx (0..7),
 * elapsed(i → j) in ms after synchronising on j */
AZ_API int az_engine_event_record(az_engine* e, int idx);
AZ_API int az_engine_event_elapsed(az_engine* e, int i, int j, float* ms);

/* IGameState on the device rules kernels (include/alphazero/core/igamestate.h:60-223), batched: replays
 * `n_moves[g]` moves of game g from the empty board, then reports per game: legal moves in reference order
 * (legal[g*A .. ], n_legal[g]), isTerminal, getGameResult, getCurrentPlayer and the enhanced tensor
 * (planes [g][C][H][W] fp32, may be NULL). */
AZ_API int az_rules_replay(az_engine* e, const int32_t* moves, const int32_t* n_moves, int n_games, int max_moves,
                    int32_t* legal, int32_t* n_legal, int32_t* terminal, int32_t* result, int32_t* player,
                    float* planes);

#ifdef __cplusplus
}
#endif
#endif /* AZ_B200_H */
