// tree_kernels.cuh — the per-wave tree kernels: warp-per-tree PUCT selection, leaf replay + terminal
// detection + feature encoding, expansion with legal-order prior normalisation, backup; plus the
// per-move kernels (choose / record sample, re-root with compaction, game turnover, Dirichlet noise).
//
// One warp owns one tree and runs ONE simulation per wave, so inside a tree the arithmetic is exactly the
// reference's serial ParallelMCTS (numThreads = 1): every float operation below is a single correctly
// rounded fp32 op in the reference's order (no FMA contraction) and every tie-break follows the reference
// scan order — visit counts are bit-exact, see tests/test_engine_gpu.py.
#pragma once
#include <cuda_bf16.h>
#include <cfloat>
#include "tree.cuh"

namespace az {

constexpr int HASH_EVAL_CHUNK = 1024;

AZ_D int warp_bcast(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// Each warp owns one tree; its game state lives in the warp's slice of dynamic shared memory (G::Warp, see the
// "Warp API" of gomoku.cuh / go.cuh).  Blocks are 128 threads = 4 warps.
template <class G>
AZ_D typename G::Warp& warp_ws(unsigned char* smem, int extra_bytes_per_warp = 0) {
    constexpr int WS = (int)((sizeof(typename G::Warp) + 15) / 16 * 16);
    return *reinterpret_cast<typename G::Warp*>(smem + (size_t)(threadIdx.x >> 5) * (WS + (extra_bytes_per_warp + 15) / 16 * 16));
}
template <class G>
constexpr size_t warp_ws_bytes(int extra_bytes_per_warp = 0) { return 4 * ((sizeof(typename G::Warp) + 15) / 16 * 16 + (size_t)(extra_bytes_per_warp + 15) / 16 * 16); }

// ------------------------------------------------------------------------------------------------
// Selection (M3/M4) + leaf replay + terminal test (M10) + feature encoding (G5 / Go7).
// mode 0: one simulation; mode 1: root-expansion wave (search() preamble, parallel_mcts.cpp:153-174).
template <class G>
__global__ void __launch_bounds__(128) k_select(TreePools tp, const typename G::State* __restrict__ root_state,
                                               typename G::Leaf* __restrict__ leaf_state, WaveBuffers wb,
                                               SearchParams sp, typename G::EncTarget enc, EvalTT tt, int T, int mode, DupStats ds, EvalCache ec) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= T) return;
    typename G::Warp& w = warp_ws<G>(smem);
    const size_t base = (size_t)tp.base[t];
    const uint8_t tf = tp.tflags[t];
    int8_t kind = LEAF_NONE;
    int node = tp.root[t];
    int depth = 0;
    float tvalue = 0.0f;

    const bool live = (tf & TF_ACTIVE) && !(tf & TF_GAME_OVER);
    if (live) {
        const int rf = tp.first[base + node];
        const uint8_t rflags = tp.flags[base + node];
        const int rootN = tp.N[base + node], rootNc = tp.nchild[base + node];      // with first / flags: the root's header in one round trip
        const bool rterm = rflags & NF_TERMINAL;
        if (mode == 1) {
            if (rf < 0 && !rterm) { kind = LEAF_EVAL; G::w_load_root(w, root_state + t, lane); }   // root needs its first evaluation
        } else if (rf >= 0 && !rterm) {
            G::w_load_root(w, root_state + t, lane);
            // --- selectLeafWithPath (parallel_mcts.cpp:456-535).  The root carries one fresh virtual
            // loss while its children are scored (:461), so parentVisits = N_root + virtualLoss (:539).
            int* path = wb.path + (size_t)t * MAX_DEPTH;
            if (lane == 0) path[0] = node;
            int parentN = rootN + sp.virtual_loss;
            int f = rf;
            int nc = rootNc;
            uint8_t lfl = rflags;          // flags of the node the descent stands on
            while (true) {
                // QUIRK M4: Q is negated only for children of depth-1 nodes (mcts_node.cpp:88-93)
                const bool negate = (depth == 1);
                const float sq = fsqrt((float)parentN);
                float best = -FLT_MAX; int bi = 0x7fffffff;
                // children in batches of 4 x 32: all N / W / P loads of a batch are issued before the first is used (with W and P loaded only
                // behind the test of N, every 32 children cost two dependent memory round trips: 14 per level on a full Gomoku node)
                for (int i0 = lane; i0 < nc; i0 += 128) {
                    int cn[4]; float cw[4], cp[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + 32 * u;
                        const size_t c = base + f + (i < nc ? i : 0);
                        cn[u] = tp.N[c]; cw[u] = tp.W[c]; cp[u] = tp.P[c];
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + 32 * u;
                        if (i >= nc) break;
                        const int n = cn[u];
                        float sc;
                        if (n == 0) sc = FLT_MAX;                   // mcts_node.cpp:63-66
                        else {
                            float q = fdiv(cw[u], (float)n);        // children carry no virtual loss when scored
                            if (negate) q = -q;
                            const float uu = fdiv(fmul(fmul(sp.c_puct, cp[u]), sq), fadd(1.0f, (float)n));
                            const float d = n < 5 ? fmul(0.05f, (float)(5 - n)) : 0.0f;   // :113-116
                            sc = fadd(fadd(q, uu), d);
                        }
                        if (sc > best) { best = sc; bi = i; }       // strict >: first child wins ties (:556)
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                }
                if (bi == 0x7fffffff) break;                    // no selectable child (bestChild == nullptr)
                const int child = f + bi;
                // the child's action and header in ONE round trip, issued before the move is applied (act / first / flags / nchild / N are
                // independent loads; behind w_apply's shared-memory stores they would be a second dependent round trip per level)
                const int cact = (int)tp.act[base + child];
                const int cfirst = tp.first[base + child];
                const uint8_t cfl = tp.flags[base + child];
                const int cnc = tp.nchild[base + child];
                const int cN = tp.N[base + child];
                G::w_apply(w, cact, lane);
                ++depth;
                if (lane == 0) path[depth] = child;
                node = child;
                f = cfirst; nc = cnc; parentN = cN; lfl = cfl;
                if (f < 0 || (cfl & NF_TERMINAL) || depth >= sp.max_depth - 1) break;
            }
            // --- leaf classification (parallel_mcts.cpp:300-313)
            const uint8_t fl = lfl;
            if (fl & NF_TERMINAL) { kind = LEAF_TERMINAL; tvalue = result_to_value((fl >> NF_RESULT_SHIFT) & 3, G::w_player(w)); }
            else {
                const int res = G::w_result(w, lane);
                if (res != RES_ONGOING) {
                    kind = LEAF_TERMINAL; tvalue = result_to_value(res, G::w_player(w));
                    if (lane == 0) tp.flags[base + node] = (uint8_t)(NF_TERMINAL | (res << NF_RESULT_SHIFT));
                } else kind = LEAF_EVAL;
            }
            ++depth;   // path length
        }
    }
    int slot = -1;
    const bool dedup = wb.dd_keys != nullptr;
    if (kind == LEAF_EVAL) {
        if (dedup) {
            // register the leaf's network input in the wave's key set; slots are handed out, and the planes written, by k_dedup_encode
            const uint64_t ki = G::w_input_key(w, lane);
            if (lane == 0) {
                unsigned long long k = ki ? ki : 1ULL;
                // evaluation cache first (EvalCache, tree.cuh): an input evaluated in an earlier wave needs neither a slot nor an owner
                int ce = -1;
                if (ec.keys != nullptr) {
                    const unsigned int b = cache_bucket(ec, k);
                    const ulonglong2 k01 = *reinterpret_cast<const ulonglong2*>(ec.keys + b), k23 = *reinterpret_cast<const ulonglong2*>(ec.keys + b + 2);   // the bucket: one 32-byte sector
                    static_assert(CACHE_WAYS == 4, "the probe reads four ways");
                    ce = k01.x == k ? (int)b : k01.y == k ? (int)b + 1 : k23.x == k ? (int)b + 2 : k23.y == k ? (int)b + 3 : -1;
                    if (ce >= 0) ec.stamp[ce] = ec.wave[0];         // hit: refreshed, and protected from this wave's stores
                    wb.cache_entry[t] = ce;
                }
                if (ce < 0) {
                    unsigned int i = (unsigned int)mix64(k) & wb.dd_mask;
                    while (true) {
                        const unsigned long long old = atomicCAS(&wb.dd_keys[i], 0ULL, k);
                        if (old == 0ULL || old == k) break;
                        i = (i + 1) & wb.dd_mask;                   // the set has 4 x T entries: always terminates
                    }
                    atomicMin(&wb.dd_owner[i], t);
                    wb.dd_idx[t] = (int32_t)i;
                }
            }
        } else {
            if (lane == 0) { slot = atomicAdd(wb.n_eval, 1); if (wb.slot_tree) wb.slot_tree[slot] = t; }
            slot = warp_bcast(slot, 0);
        }
        if constexpr (G::TT_COARSE) {
            // TranspositionTable lookup / store of the reference (parallel_mcts.cpp:153-163 for the root, :316-358 for a leaf): the table
            // key is the placement; a hit evaluates under the first-seen position's key.  One simulation per tree per wave keeps the
            // reference's serial lookup-then-store order.
            if (tt.keys != nullptr && wb.eval_key != nullptr) {
                const uint64_t own = G::w_key(w, lane);
                if (lane == 0) {
                    uint64_t ek = own, pk = G::w_tt_key(w);
                    if (pk == 0) pk = 1;
                    const size_t tb = (size_t)t * tt.cap;
                    const uint32_t mask = (uint32_t)tt.cap - 1;
                    uint32_t i = (uint32_t)mix64(pk) & mask;
                    for (int probe = 0; probe < tt.cap; ++probe, i = (i + 1) & mask) {
                        const uint64_t k = tt.keys[tb + i];
                        if (k == pk) { ek = tt.vals[tb + i]; break; }
                        if (k == 0) { if (tt.count[t] < tt.cap / 2) { tt.keys[tb + i] = pk; tt.vals[tb + i] = own; tt.count[t] += 1; } break; }
                    }
                    wb.eval_key[t] = ek;
                }
            }
        }
    }
    if (lane == 0) {
        wb.leaf_kind[t] = kind; wb.leaf_node[t] = node; wb.path_len[t] = (mode == 1 || kind == LEAF_NONE) ? 0 : depth;
        wb.leaf_value[t] = tvalue; wb.eval_slot[t] = slot;
    }
    if (kind == LEAF_EVAL) {
        G::w_store_leaf(w, leaf_state + t, lane);
        if (enc.ptr != nullptr && !dedup) G::w_encode(w, lane, enc, slot);
        if (ds.counters != nullptr) {
            const uint64_t ki = G::w_input_key(w, lane), kr = G::w_ref_tt_key(w);
            if (lane == 0) {
                bool full = false;
                atomicAdd(&ds.counters[0], 1ULL);
                if (dup_probe_insert(ds.wave_keys, ds.wave_mask, ki, nullptr)) atomicAdd(&ds.counters[1], 1ULL);
                if (dup_probe_insert(ds.run_keys, ds.run_mask, ki, &full)) atomicAdd(&ds.counters[2], 1ULL);
                if (dup_probe_insert(ds.wave_keys_ref, ds.wave_mask, kr, nullptr)) atomicAdd(&ds.counters[3], 1ULL);
                if (dup_probe_insert(ds.run_keys_ref, ds.run_mask, kr, &full)) atomicAdd(&ds.counters[4], 1ULL);
                if (full) atomicAdd(&ds.counters[5], 1ULL);
            }
        }
        if constexpr (G::LEGAL_POLICY) {
            // wide policy head: the network computes the logits of the leaf's legal moves only (heads.cu k_policy_legal_value).  The leaf's
            // terminal test (w_result) has just enumerated them; the root-expansion wave does it here
            if (wb.legal != nullptr) {
                const int n = G::w_store_legal(w, lane, wb.legal + (size_t)t * G::MAX_CHILDREN, mode == 1);
                if (lane == 0) wb.n_legal[t] = n;
            }
        }
    }
}

// In-wave evaluation dedup, second half: the owner of every distinct network input (the smallest tree index that registered it) takes an
// evaluation slot and writes its feature planes; every other tree with that input points at the owner.  Result-transparent: the
// network output of a board does not depend on its position in the batch (tests/test_nn_gpu.py), so sharing the owner's policy / value
// is bit-identical to evaluating the duplicate.
template <class G>
__global__ void __launch_bounds__(128) k_dedup_encode(const typename G::Leaf* __restrict__ leaf_state, const typename G::State* __restrict__ root_state,
                                                     WaveBuffers wb, typename G::EncTarget enc, int T, Stats* stats, uint32_t* cache_wave) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (cache_wave != nullptr && blockIdx.x == 0 && threadIdx.x == 0) { const uint32_t wv = cache_wave[0]; cache_wave[1] = wv; cache_wave[0] = wv + 1; }   // EvalCache::wave (tree.cuh)
    if (t >= T) return;
    if (wb.leaf_kind[t] != LEAF_EVAL) return;
    if (wb.cache_entry != nullptr && wb.cache_entry[t] >= 0) {          // served by the evaluation cache
        if (lane == 0) atomicAdd(&stats->eval_cached, 1ULL);
        return;
    }
    const int owner = wb.dd_owner[wb.dd_idx[t]];
    if (owner != t) {
        if (lane == 0) { wb.eval_slot[t] = -2 - owner; atomicAdd(&stats->eval_shared, 1ULL); }
        return;
    }
    typename G::Warp& w = warp_ws<G>(smem);
    int slot = 0;
    if (lane == 0) { slot = atomicAdd(wb.n_eval, 1); wb.eval_slot[t] = slot; if (wb.slot_tree) wb.slot_tree[slot] = t; }
    slot = warp_bcast(slot, 0);
    if (enc.ptr == nullptr) return;                                       // hash evaluators read the leaf state themselves
    G::w_load_leaf(w, leaf_state + t, root_state + t, lane);
    G::w_encode(w, lane, enc, slot);
}

// AZ_EVAL_EXTERNAL: the leaves of the wave as move sequences from their roots, by evaluation slot (the host replays them on its own states)
__global__ void k_leaf_paths(TreePools tp, WaveBuffers wb, int32_t* slot_tree, int32_t* path_actions /*[T][MAX_DEPTH]*/, int32_t* path_len, int T) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T || wb.leaf_kind[t] != LEAF_EVAL) return;
    const int slot = wb.eval_slot[t];
    const size_t base = (size_t)tp.base[t];
    const int n = max(wb.path_len[t] - 1, 0);               // path[0] is the root; the root-expansion wave has no path
    const int* path = wb.path + (size_t)t * MAX_DEPTH;
    for (int i = 0; i < n; ++i) path_actions[(size_t)slot * MAX_DEPTH + i] = (int32_t)tp.act[base + path[i + 1]];
    path_len[slot] = n; slot_tree[slot] = t;
}

// ------------------------------------------------------------------------------------------------
// Stateless HashEvaluator — the deterministic evaluator used for bit-exact parity
// against the reference's serial search.  One warp per leaf; the policy sum is accumulated in ascending
// action order by one lane so it is the reference's fp32 sum bit for bit.
template <class G>
__global__ void __launch_bounds__(128) k_hash_eval(const typename G::Leaf* __restrict__ leaf_state, const typename G::State* __restrict__ root_state,
                                                  WaveBuffers wb, int T, int peaked) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= T) return;
    if (wb.leaf_kind[t] != LEAF_EVAL) return;
    constexpr int A = G::ACTIONS;
    constexpr int CH = A < HASH_EVAL_CHUNK ? A : HASH_EVAL_CHUNK;      // raw values staged per chunk (chess: A = 20480)
    typename G::Warp& w = warp_ws<G>(smem, CH * 4);
    float* raw = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(&w) + (sizeof(typename G::Warp) + 15) / 16 * 16);
    const int slot = wb.eval_slot[t];
    if (slot < 0) return;                                                 // shared with another tree of the wave, or served by the evaluation cache
    G::w_load_leaf(w, leaf_state + t, root_state + t, lane);
    const uint64_t h = (G::TT_COARSE && wb.eval_key != nullptr) ? wb.eval_key[t] : G::w_key(w, lane);
    // AZ_EVAL_HASH_PEAKED: the raw prior of action mix(h ^ 0x5EED) % A is multiplied by 4096 (exact) before the normalisation
    const int peak = peaked ? (int)(mix64(h ^ 0x5EEDULL) % (uint64_t)A) : -1;
    float sum = 0.0f;
    for (int c0 = 0; c0 < A; c0 += CH) {
        __syncwarp();
        for (int i = lane; i < CH && c0 + i < A; i += 32) {
            const uint64_t r = mix64(h + (uint64_t)(c0 + i) * 0x9E3779B97F4A7C15ULL) >> 40;
            raw[i] = fdiv((float)(r + 1), 16777216.0f);
            if (c0 + i == peak) raw[i] = fmul(raw[i], 4096.0f);
        }
        __syncwarp();
        if (lane == 0) for (int i = 0; i < CH && c0 + i < A; ++i) sum = fadd(sum, raw[i]);      // ascending action order, one lane
    }
    sum = __shfl_sync(0xffffffffu, sum, 0);
    for (int i = lane; i < A; i += 32) {
        const uint64_t r = mix64(h + (uint64_t)i * 0x9E3779B97F4A7C15ULL) >> 40;
        float rw = fdiv((float)(r + 1), 16777216.0f);
        if (i == peak) rw = fmul(rw, 4096.0f);
        wb.policy[(size_t)slot * A + i] = fdiv(rw, sum);
    }
    if (lane == 0) {
        float v = fdiv((float)(mix64(h ^ 0xABCDEFULL) >> 40), 16777216.0f);
        v = fsub(fmul(v, 2.0f), 1.0f);
        wb.value[slot] = fmul(v, 0.5f);
    }
}

// ------------------------------------------------------------------------------------------------
// Expansion (M5) + backup (M6/M7).
template <class G>
__global__ void __launch_bounds__(128) k_expand_backup(TreePools tp, const typename G::Leaf* __restrict__ leaf_state,
                                                      const typename G::State* __restrict__ root_state, WaveBuffers wb, const int16_t* __restrict__ root_order,
                                                      const int32_t* __restrict__ root_order_n, SearchParams sp,
                                                      int T, Stats* stats, EvalCache ec) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= T) return;
    // The kernel is a chain of dependent memory round trips (one warp per tree): every per-tree scalar it will need is loaded here, in ONE
    // round trip, before the first of them is used (they were four: kind -> base / leaf -> slot / cache entry -> tree flags / alloc, and
    // further down limit, the dedup entry and the path length one at a time).
    const int8_t kind = wb.leaf_kind[t];
    const long long base_ = tp.base[t];
    float v = wb.leaf_value[t];
    const int leaf = wb.leaf_node[t];
    const int slot_ = wb.eval_slot[t];
    const int ce_ = ec.keys != nullptr ? wb.cache_entry[t] : -1;
    const int ddi_ = (ec.keys != nullptr && wb.dd_idx != nullptr) ? wb.dd_idx[t] : -1;
    const uint32_t cwave = ec.keys != nullptr ? ec.wave[1] : 0u;
    const uint8_t tf = tp.tflags[t];
    const int troot = tp.root[t];
    const int alloc = tp.alloc[t];
    const int tlimit = tp.limit[t];
    const int plen = wb.path_len[t];
    if (kind == LEAF_NONE) return;
    constexpr int A = G::ACTIONS, MC = G::MAX_CHILDREN;
    constexpr int EXTRA = MC * 4 + (MC * 2 + 15) / 16 * 16;
    typename G::Warp& w = warp_ws<G>(smem, EXTRA);
    float* raw = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(&w) + (sizeof(typename G::Warp) + 15) / 16 * 16);
    int16_t* acts = reinterpret_cast<int16_t*>(raw + MC);
    const size_t base = (size_t)base_;
    int n_new = 0;                                      // children created by this simulation: added to every path node's descendant count
    // the path nodes' statistics (one node per lane, backup below): nothing in the expansion writes W / N / sub of a path node, so they are
    // fetched now and are in flight behind the expansion
    const int* path = wb.path + (size_t)t * MAX_DEPTH;
    const int pj = lane < plen ? path[lane] : 0;
    float pW = 0.0f; int pN = 0, pS = 0;
    if (lane < plen) { pW = tp.W[base + pj]; pN = tp.N[base + pj]; pS = tp.sub[base + pj]; }

    if (kind == LEAF_EVAL) {
        int slot = slot_;
        const bool own_eval = slot >= 0;                     // this tree's leaf went through the evaluator itself
        const int ce = ce_;
        // the key this evaluation is stored under (the wave's key set entry of this tree), fetched ahead of the store
        const unsigned long long store_key = (ec.keys != nullptr && own_eval && ddi_ >= 0) ? wb.dd_keys[ddi_] : 0ULL;
        const float* pol; const float* cpol = nullptr;
        if (ce >= 0) {                                       // evaluation cache: an earlier wave's evaluation of the same network input
            cpol = ec.policy + (size_t)ce * ec.pw;
            pol = G::LEGAL_POLICY ? nullptr : cpol;
            v = ec.value[ce];
        } else {
            if (slot <= -2) slot = wb.eval_slot[-2 - slot];  // in-wave dedup: the owner's evaluation of the same network input
            pol = wb.policy + (size_t)slot * A;
            v = wb.value[slot];
        }
        G::w_load_leaf(w, leaf_state + t, root_state + t, lane);
        const bool first_fill = (tf & TF_FIRST_FILL) && leaf == troot;
        // children = the legal moves in the reference's order, priors = policy[action] (0 for out-of-range actions, i.e.
        // Go's pass at -1), expandNodeWithPolicy parallel_mcts.cpp:690-711
        int n;
        if constexpr (G::LEGAL_POLICY) {
            // the selection kernel already enumerated this leaf's legal moves for the policy head: no second move generation
            if (wb.legal != nullptr) n = G::w_enumerate_pre(lane, wb.legal + (size_t)t * MC, wb.n_legal[t], acts, raw, pol);
            else n = G::w_enumerate(w, lane, nullptr, 0, acts, raw, pol);
        } else n = G::w_enumerate(w, lane, first_fill ? root_order + (size_t)t * MC : nullptr, first_fill ? root_order_n[t] : 0, acts, raw, pol);
        if constexpr (G::LEGAL_POLICY) {
            if (cpol != nullptr) { for (int i = lane; i < n; i += 32) raw[i] = cpol[i]; __syncwarp(); }      // cached: the children's raw priors in child order
        }
        if (ec.keys != nullptr && own_eval) {
            // store this evaluation: claim the oldest way of the bucket that was neither hit nor stored in this wave (tree.cuh)
            int claimed = -1;
            if (lane == 0) {
                const unsigned long long k = store_key;
                const unsigned int b = cache_bucket(ec, k);
                const uint32_t wave = cwave;
                for (int attempt = 0; attempt < CACHE_WAYS && claimed < 0; ++attempt) {
                    int best = -1; uint32_t best_age = 0, best_s = 0;
                    // the bucket's four stamps and four keys as eight independent loads (one round trip), then the choice
                    uint32_t st[CACHE_WAYS]; unsigned long long ky[CACHE_WAYS];
#pragma unroll
                    for (int wy = 0; wy < CACHE_WAYS; ++wy) { st[wy] = *(volatile uint32_t*)&ec.stamp[b + wy]; ky[wy] = *(volatile unsigned long long*)&ec.keys[b + wy]; }
#pragma unroll
                    for (int wy = 0; wy < CACHE_WAYS; ++wy) {
                        const uint32_t s = st[wy];
                        if (s == wave) continue;
                        const uint32_t age = ky[wy] == 0ULL ? 0xffffffffu : wave - s;
                        if (best < 0 || age > best_age) { best = wy; best_age = age; best_s = s; }
                    }
                    if (best < 0) break;
                    if (atomicCAS(&ec.stamp[b + best], best_s, wave) == best_s) claimed = (int)(b + best);
                }
                if (claimed >= 0) { ec.keys[claimed] = k; ec.value[claimed] = v; }
            }
            claimed = warp_bcast(claimed, 0);
            if (claimed >= 0) {
                float* dst = ec.policy + (size_t)claimed * ec.pw;
                if constexpr (G::LEGAL_POLICY) { for (int i = lane; i < n; i += 32) dst[i] = raw[i]; }
                else { for (int i = lane; i < A; i += 32) dst[i] = pol[i]; }
            }
        }
        if (alloc + n > tlimit) {
            if (lane == 0) { tp.tflags[t] = tf | TF_OVERFLOW; atomicAdd(&stats->pool_overflows, 1ULL); }
        } else if (n > 0) {
            // policySum accumulated in child order (parallel_mcts.cpp:705-711) — serial on purpose
            float sum = 0.0f;
            if (lane == 0) for (int i = 0; i < n; ++i) if (acts[i] >= 0 && acts[i] < A) sum = fadd(sum, raw[i]);
            sum = __shfl_sync(0xffffffffu, sum, 0);
            const float uniform = fdiv(1.0f, (float)n);
            for (int i = lane; i < n; i += 32) {
                const size_t c = base + alloc + i;
                tp.N[c] = 0; tp.W[c] = 0.0f; tp.first[c] = -1; tp.sub[c] = 0; tp.nchild[c] = 0; tp.flags[c] = 0; tp.act[c] = acts[i];
                tp.P[c] = sum > 0.0f ? fdiv(raw[i], sum) : uniform;     // :714-724
            }
            n_new = n;
            if (lane == 0) {
                tp.first[base + leaf] = alloc; tp.nchild[base + leaf] = (int16_t)n; tp.alloc[t] = alloc + n;
                if (first_fill) tp.tflags[t] = tf & ~TF_FIRST_FILL;
                atomicAdd(&stats->nodes_created, (unsigned long long)n);
                atomicAdd(&stats->nodes_expanded, 1ULL);
            }
        }
        if (lane == 0) atomicAdd(&stats->evaluations, 1ULL);
    } else if (lane == 0) atomicAdd(&stats->terminal_leaves, 1ULL);

    // --- virtual loss add/remove + backpropagate, folded (mcts_node.cpp:168-196, parallel_mcts.cpp:782-833).
    // Per path node the reference does W -= 3 (added after selection), then in reverse path order
    // W += 3, N += 1, W += v.  The root received one extra virtual loss at selection start that is never
    // removed (QUIRK M7): N_root += 4, VL_root += 3, W_root goes through -3,-3,+3,+v.
    if (plen == 0 && n_new > 0 && lane == 0) tp.sub[base + leaf] += n_new;      // root-expansion wave: no path, the leaf is the root
    if (plen > 0) {
        // every path node's update depends only on its own W / N and on the sign the value has at its depth (v at the leaf, negated once per
        // level up): one node per lane instead of a serial walk — the same fp32 operations per node, in the same order
        const float vl = (float)sp.virtual_loss;
        for (int j = lane; j < plen; j += 32) {
            const size_t c = base + (j == lane ? pj : path[j]);
            const float cv = ((plen - 1 - j) & 1) ? -v : v;
            float wv; int nv, sv;
            if (j == lane) { wv = pW; nv = pN; sv = pS; }                                // fetched ahead of the expansion
            else { wv = tp.W[c]; nv = tp.N[c]; sv = tp.sub[c]; }                         // paths deeper than 32 nodes: the three loads together, then the stores
            if (j == 0) { wv = fsub(wv, vl); wv = fsub(wv, vl); wv = fadd(wv, vl); wv = fadd(wv, cv);
                          tp.N[c] = nv + sp.virtual_loss + 1; tp.root_vl[t] += sp.virtual_loss; }
            else { wv = fsub(wv, vl); wv = fadd(wv, vl); wv = fadd(wv, cv); tp.N[c] = nv + 1; }
            tp.W[c] = wv;
            if (n_new) tp.sub[c] = sv + n_new;
        }
        if (lane == 0) atomicAdd(&stats->simulations, 1ULL);
    }
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (counter-based RNG; per-slot streams: key = seed, counter = (slot, game, move, draw)).
struct Philox {
    uint32_t k0, k1, c0, c1, c2, c3; uint32_t out[4]; int have;
    AZ_D Philox(uint64_t seed, uint32_t a, uint32_t b, uint32_t c) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), c0(0), c1(a), c2(b), c3(c), have(0) {}
    AZ_D void round4() {
        uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = c3, a = k0, b = k1;
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
            x0 = hi1 ^ x1 ^ a; x1 = lo1; x2 = hi0 ^ x3 ^ b; x3 = lo0;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        out[0] = x0; out[1] = x1; out[2] = x2; out[3] = x3; ++c0; have = 4;
    }
    AZ_D uint32_t next() { if (!have) round4(); return out[--have]; }
    AZ_D float uniform() { return ((float)(next() >> 8) + 0.5f) * (1.0f / 16777216.0f); }   // (0,1)
    AZ_D float normal() { const float u1 = uniform(), u2 = uniform(); return sqrtf(-2.0f * __logf(u1)) * __cosf(6.28318530718f * u2); }
    AZ_D float gamma(float alpha) {   // Marsaglia–Tsang, with the alpha<1 boost
        float boost = 1.0f;
        if (alpha < 1.0f) { boost = __powf(uniform(), 1.0f / alpha); alpha += 1.0f; }
        const float d = alpha - 1.0f / 3.0f, c = rsqrtf(9.0f * d);
        for (int it = 0; it < 64; ++it) {
            const float x = normal(); float v = 1.0f + c * x;
            if (v <= 0.0f) continue;
            v = v * v * v; const float u = uniform();
            if (__logf(u) < 0.5f * x * x + d - d * v + d * __logf(v)) return d * v * boost;
        }
        return d * boost;
    }
};

// Dirichlet noise on the root priors (M11, parallel_mcts.cpp:1110-1171): P = (1-eps) P + eps * noise_i by
// child index.  The reference draws from libstdc++'s gamma_distribution with a random_device seed, which is
// unpinned; here the draws come from Philox(seed; slot, game, move).
__global__ void __launch_bounds__(128) k_dirichlet(TreePools tp, int T, int slot_base, float alpha, float eps, uint64_t seed, float* scratch /*[T][maxA]*/, int maxA) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= T) return;
    const uint8_t tf = tp.tflags[t];
    if (!(tf & TF_ACTIVE) || (tf & TF_GAME_OVER) || !(tf & TF_NEED_NOISE)) return;
    const size_t base = (size_t)tp.base[t];
    const int root = tp.root[t];
    const int f = tp.first[base + root];
    if (f < 0) return;                                   // terminal root: nothing to perturb
    const int nc = tp.nchild[base + root];
    float* nz = scratch + (size_t)t * maxA;
    float part = 0.0f;
    for (int i = lane; i < nc; i += 32) {
        Philox rng(seed, (uint32_t)(slot_base + t), tp.game_id[t], ((uint32_t)tp.move_num[t] << 12) | (uint32_t)i);
        const float g = fmaxf(1e-10f, rng.gamma(alpha));
        nz[i] = g; part += g;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    const float inv = part > 0.0f ? 1.0f / part : 0.0f;
    for (int i = lane; i < nc; i += 32) {
        const size_t c = base + f + i;
        const float noise = part > 0.0f ? nz[i] * inv : 1.0f / (float)nc;
        tp.P[c] = (1.0f - eps) * tp.P[c] + eps * noise;
    }
    if (lane == 0) tp.tflags[t] = tf & ~TF_NEED_NOISE;
}

__global__ void k_flag_noise(TreePools tp, int T) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < T && (tp.tflags[t] & TF_ACTIVE) && !(tp.tflags[t] & TF_GAME_OVER)) tp.tflags[t] |= TF_NEED_NOISE;
}

// ------------------------------------------------------------------------------------------------
// Sample record: one per move played (the device-side equivalent of selfplay::MoveData,
// include/alphazero/selfplay/game_record.h:18-70, with action-indexed visit counts instead of the
// reference's child-ordered policy — SURVEY.md §8f.1).
template <class G>
struct alignas(16) Sample {
    uint32_t game_id; int32_t slot;
    int16_t ply, action; int8_t player, z, result, pad_;
    float root_value; int32_t root_visits;
    typename G::Snapshot state;                    // position the move was chosen from
    uint16_t visits[(G::SAMPLE_VISITS + 7) / 8 * 8];   // root child visit counts: by action for Gomoku / Go (pass last), (action, count) pairs for chess
};

struct MoveParams {
    int deterministic;          // 1: first-max-visit child (what SelfPlayManager's forced useBatchInference does, M13)
    float init_temperature, final_temperature; int temperature_drop_move;   // self_play_manager.cpp:236-240
    uint64_t seed;
};

// Choose the move (M13), record the sample (S1/S2), advance the root state.  Leaves the chosen child's
// old index in chosen_child[t] for k_reroot.  forced_action: nullptr, or per-slot action to play
// (-2 = leave slot alone) for the manual ParallelMCTS::updateWithMove path (M14).
template <class G>
__global__ void __launch_bounds__(128) k_choose_move(TreePools tp, typename G::State* root_state, MoveParams mp,
                                                    Sample<G>* game_buf /*[T][max_moves]*/, int max_moves,
                                                    const int32_t* forced_action, int32_t* chosen_child,
                                                    int32_t* chosen_action, int T, Stats* stats) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= T) return;
    typename G::Warp& w = warp_ws<G>(smem);
    if (lane == 0) { chosen_child[t] = -2; chosen_action[t] = -2; }
    const uint8_t tf = tp.tflags[t];
    if (!(tf & TF_ACTIVE) || (tf & TF_GAME_OVER)) return;
    if (forced_action && forced_action[t] == -2) return;
    const size_t base = (size_t)tp.base[t];
    const int root = tp.root[t];
    const int f = tp.first[base + root];
    const int nc = f >= 0 ? tp.nchild[base + root] : 0;
    int bi = -1, action = -1;
    if (forced_action) {
        action = forced_action[t];
        int found = 0x7fffffff;
        for (int i = lane; i < nc; i += 32) if (tp.act[base + f + i] == action) found = min(found, i);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) found = min(found, __shfl_xor_sync(0xffffffffu, found, o));
        bi = found == 0x7fffffff ? -1 : found;
    } else {
        if (nc == 0) return;     // unexpanded / terminal root: nothing to choose (host keeps roots expanded)
        const int mv = tp.move_num[t];
        const float T_ = mv >= mp.temperature_drop_move ? mp.final_temperature : mp.init_temperature;
        if (mp.deterministic || T_ <= 0.0f) {
            // getBestActions()[0] / max_element: first child with the maximum visit count
            int bn = -1; bi = 0x7fffffff;
            for (int i = lane; i < nc; i += 32) { const int n = tp.N[base + f + i]; if (n > bn) { bn = n; bi = i; } }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const int on = __shfl_xor_sync(0xffffffffu, bn, o), oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (on > bn || (on == bn && oi < bi)) { bn = on; bi = oi; }
            }
        } else {
            // sample child ∝ N^(1/T)  (getVisitCountDistribution + discrete_distribution, M12/M13)
            const float e = 1.0f / fmaxf(0.01f, T_);
            float part = 0.0f;
            for (int i = lane; i < nc; i += 32) part += powf((float)tp.N[base + f + i], e);
            float tot = part;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
            if (lane == 0) {
                Philox rng(mp.seed ^ 0x5eedULL, (uint32_t)t, tp.game_id[t], (uint32_t)mv);
                const float target = rng.uniform() * tot;
                float acc = 0.0f; bi = nc - 1;
                for (int i = 0; i < nc; ++i) { acc += powf((float)tp.N[base + f + i], e); if (acc > target) { bi = i; break; } }
            }
            bi = warp_bcast(bi, 0);
        }
        action = tp.act[base + f + bi];
    }
    G::w_load_root(w, root_state + t, lane);
    // --- record the sample
    const int mv = tp.move_num[t];
    if (game_buf && mv < max_moves) {
        Sample<G>* sp_ = game_buf + (size_t)t * max_moves + mv;
        for (int i = lane; i < (int)(sizeof(sp_->visits) / 2); i += 32) sp_->visits[i] = 0;
        __syncwarp();
        for (int i = lane; i < nc; i += 32) G::record_visit(sp_->visits, i, tp.act[base + f + i], tp.N[base + f + i]);
        if (lane == 0) {
            const int rn = tp.N[base + root];
            sp_->game_id = tp.game_id[t]; sp_->slot = t; sp_->ply = (int16_t)G::w_ply(w); sp_->action = (int16_t)action;
            sp_->player = (int8_t)G::w_player(w); sp_->z = 0; sp_->result = 0; sp_->pad_ = 0;
            sp_->root_value = (nc == 0 || rn == 0) ? 0.0f : fdiv(tp.W[base + root], (float)rn);   // getRootValue (M15)
            sp_->root_visits = rn;
        }
        G::w_snapshot(w, &sp_->state, lane);
    }
    // --- advance the root state (updateWithMove, M14).  A forced action the rules reject leaves the slot untouched
    // (the reference's makeMove throws): chosen_action = -3 tells the host.
    const bool ok = G::w_apply(w, action, lane, forced_action != nullptr);
    if (!ok) { if (lane == 0) chosen_action[t] = -3; return; }
    G::w_store_root(w, root_state + t, lane);
    if (lane == 0) {
        chosen_child[t] = bi >= 0 ? f + bi : -1;
        chosen_action[t] = action;
        tp.move_num[t] = mv + 1;
        atomicAdd(&stats->moves, 1ULL);
    }
}

// Re-root (M14, updateWithMove parallel_mcts.cpp:1065-1108) + region re-cut.  Three steps per move commit:
//   k_region_need  — per tree: nodes kept = 1 + descendants of the chosen child (the whole tree if the slot did not move, 1 for a fresh root)
//   k_region_plan  — one block: prefix sum of kept + growth budget over the trees → new base / limit in the OTHER pool buffer; when the
//                    pool cannot hold every tree's worst-case growth the budgets shrink evenly and `short_nodes` reports by how much
//   k_reroot_copy  — one warp per tree: breadth-first copy of the kept subtree old buffer → new region (children stay contiguous and in
//                    order); the reference frees the siblings recursively, here they are simply not copied.
__global__ void k_region_need(TreePools tp, const int32_t* __restrict__ chosen_child, int32_t* kept, int T) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int cc = chosen_child[t];
    const size_t base = (size_t)tp.base[t];
    kept[t] = cc == -2 ? tp.alloc[t] : (cc < 0 ? 1 : tp.sub[base + cc] + 1);
}
struct RegionPlan { unsigned long long need_total, kept_total, short_nodes; };
__global__ void __launch_bounds__(1024) k_region_plan(const int32_t* __restrict__ kept, int64_t* new_base, int32_t* new_limit, int T, int budget,
                                                     long long pool_nodes, RegionPlan* plan) {
    __shared__ long long s_part[1024];
    __shared__ long long s_kept_total;
    const int tid = threadIdx.x, per = (T + 1023) / 1024;
    const int t0 = min(tid * per, T), t1 = min(t0 + per, T);
    long long k = 0;
    for (int t = t0; t < t1; ++t) k += kept[t];
    s_part[tid] = k;
    __syncthreads();
    if (tid == 0) { long long acc = 0; for (int i = 0; i < 1024; ++i) { const long long v = s_part[i]; s_part[i] = acc; acc += v; } s_kept_total = acc; }
    __syncthreads();
    const long long kept_total = s_kept_total;
    // growth per tree: the full budget plus an equal share of the slack, or — pool too small — an equal share of what is left
    const long long room = pool_nodes - kept_total;
    long long grow = room > 0 ? room / T : 0;
    if (tid == 0) { plan->need_total = (unsigned long long)(kept_total + (long long)T * budget); plan->kept_total = (unsigned long long)kept_total;
                    plan->short_nodes = grow < budget ? (unsigned long long)((long long)T * (budget - grow)) : 0ULL; }
    grow = min(grow, (long long)0x7fffffff - 0x1000000);
    long long b = s_part[tid] + (long long)t0 * grow;
    for (int t = t0; t < t1; ++t) {
        new_base[t] = b;
        const long long lim = min((long long)kept[t] + grow, pool_nodes - b);       // (kept_total > pool_nodes: truncated, flagged by the copy)
        new_limit[t] = (int32_t)max(lim, 0LL);
        b += kept[t] + grow;
    }
}

template <class G>
__global__ void __launch_bounds__(128) k_reroot_copy(TreePools tp, TreePools np, const typename G::State* __restrict__ root_state,
                                                    const int32_t* __restrict__ chosen_child, int T, Stats* stats) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= T) return;
    const int cc = chosen_child[t];
    typename G::Warp& w = warp_ws<G>(smem);
    const size_t ob = (size_t)tp.base[t], nb = (size_t)np.base[t];
    const int lim = np.limit[t];
    int count = 1;
    bool truncated = lim < 1;
    if (cc == -1) {
        // child did not exist: fresh root (parallel_mcts.cpp:1097-1101)
        if (lane == 0 && !truncated) { np.N[nb] = 0; np.W[nb] = 0.0f; np.P[nb] = 0.0f; np.first[nb] = -1; np.sub[nb] = 0; np.act[nb] = -1; np.nchild[nb] = 0; np.flags[nb] = 0; }
    } else if (!truncated) {
        const int r = cc == -2 ? tp.root[t] : cc;               // slot did not move: the whole tree goes over as it is
        // new node j: all fields copied; np.first[j] temporarily holds the OLD index of j's first child until j is processed
        if (lane == 0) {
            np.N[nb] = tp.N[ob + r]; np.W[nb] = tp.W[ob + r]; np.P[nb] = tp.P[ob + r]; np.act[nb] = tp.act[ob + r]; np.flags[nb] = tp.flags[ob + r];
            np.sub[nb] = tp.sub[ob + r]; np.nchild[nb] = tp.nchild[ob + r]; np.first[nb] = tp.first[ob + r];
        }
        __syncwarp();
        for (int j0 = 0; j0 < count && !truncated;) {
            // only nodes that exist at the start of the chunk; nodes appended meanwhile wait for a later chunk
            const int end = min(j0 + 32, count);
            const int j = j0 + lane;
            int fo = -1, no = 0;
            if (j < end) { fo = np.first[nb + j]; no = np.nchild[nb + j]; }
            unsigned m = __ballot_sync(0xffffffffu, fo >= 0);
            while (m) {
                const int src = __ffs(m) - 1; m &= m - 1;
                const int f = __shfl_sync(0xffffffffu, fo, src), n = __shfl_sync(0xffffffffu, no, src);
                if (count + n > lim) { truncated = true; break; }
                // 4 x 32 children per pass, all eight fields of all four loaded before the first store: the copy moves ~10 GB per Gomoku move
                // (profiles/r2_commit_kernels_ncu.csv) and with one child per lane in flight it ran at 8 % of the HBM roof
                for (int i0 = lane; i0 < n; i0 += 128) {
                    int32_t vN[4], vF[4], vS[4]; float vW[4], vP[4]; int16_t vA[4], vC[4]; uint8_t vL[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + 32 * u;
                        if (i < n) {
                            const size_t o = ob + f + i;
                            vN[u] = tp.N[o]; vW[u] = tp.W[o]; vP[u] = tp.P[o]; vA[u] = tp.act[o]; vL[u] = tp.flags[o];
                            vS[u] = tp.sub[o]; vC[u] = tp.nchild[o]; vF[u] = tp.first[o];
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + 32 * u;
                        if (i < n) {
                            const size_t d = nb + count + i;
                            np.N[d] = vN[u]; np.W[d] = vW[u]; np.P[d] = vP[u]; np.act[d] = vA[u]; np.flags[d] = vL[u];
                            np.sub[d] = vS[u]; np.nchild[d] = vC[u]; np.first[d] = vF[u];
                        }
                    }
                }
                if (lane == 0) np.first[nb + j0 + src] = count;
                count += n;
            }
            __syncwarp();
            j0 = end;
        }
        // pool exhausted: the slot continues from an unexpanded root (flagged TF_OVERFLOW and counted — an error, see az_engine_search)
        if (truncated && lane == 0) { np.first[nb] = -1; np.nchild[nb] = 0; np.sub[nb] = 0; }
    }
    // the new root's terminal status comes from the state, as in the root MCTSNode ctor (mcts_node.cpp:24-25)
    G::w_load_root(w, root_state + t, lane);
    const int res = cc == -2 ? RES_ONGOING : G::w_root_result(w, lane);
    if (lane == 0) {
        uint8_t tf = tp.tflags[t];
        if (cc != -2) {
            if (res != RES_ONGOING && !truncated) { np.flags[nb] = (uint8_t)(NF_TERMINAL | (res << NF_RESULT_SHIFT)); }
            if (res != RES_ONGOING) tf |= TF_GAME_OVER;
            tf &= ~TF_FIRST_FILL;
            tp.root_vl[t] = 0;
        }
        if (truncated) { tf |= TF_OVERFLOW; atomicAdd(&stats->pool_overflows, 1ULL); }
        tp.tflags[t] = tf;
        tp.root[t] = 0; tp.alloc[t] = truncated ? 1 : count;
    }
}

// Game turnover (S1, self_play_manager.cpp:187-234): stamp z / result into the finished game's samples,
// append them to the output ring, then (auto_restart) start a new game in the slot.
template <class G>
__global__ void __launch_bounds__(128) k_finish_games(TreePools tp, typename G::State* root_state, Sample<G>* game_buf, int max_moves,
                                                     Sample<G>* out, int out_cap, int* out_count, const int16_t* __restrict__ default_order,
                                                     int default_order_n, int16_t* root_order, int32_t* root_order_n,
                                                     int auto_restart, int noise_every_even_move, int T, Stats* stats, EvalTT tt) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= T) return;
    typename G::Warp& w = warp_ws<G>(smem);
    uint8_t tf = tp.tflags[t];
    if (!(tf & TF_ACTIVE)) return;
    if (!(tf & TF_GAME_OVER)) {
        // addDirichletNoise after every even move (self_play_manager.cpp:209-211): move_num was already incremented
        if (noise_every_even_move && lane == 0 && tp.move_num[t] > 0 && ((tp.move_num[t] - 1) % 2 == 0)) tp.tflags[t] = tf | TF_NEED_NOISE;
        return;
    }
    const size_t base = (size_t)tp.base[t];
    G::w_load_root(w, root_state + t, lane);
    const int res = G::w_root_result(w, lane);
    const int n = min(tp.move_num[t], max_moves);
    if (game_buf && out) {
        int dst = -1;
        if (lane == 0) {     // reserve n records, or nothing at all if the ring is full
            int old = *(volatile int*)out_count;
            while (old + n <= out_cap) {
                const int seen = atomicCAS(out_count, old, old + n);
                if (seen == old) { dst = old; break; }
                old = seen;
            }
        }
        dst = warp_bcast(dst, 0);
        if (dst >= 0) {
            for (int j = 0; j < n; ++j) {
                Sample<G>* s = game_buf + (size_t)t * max_moves + j;
                if (lane == 0) { s->result = (int8_t)res; s->z = (int8_t)result_to_value(res, s->player); }
                __syncwarp();
                const uint4* src = reinterpret_cast<const uint4*>(s);
                uint4* d = reinterpret_cast<uint4*>(out + dst + j);
                for (int i = lane; i < (int)(sizeof(Sample<G>) / 16); i += 32) d[i] = src[i];
            }
        } else if (lane == 0) atomicAdd(&stats->samples_dropped, (unsigned long long)n);
    }
    if (auto_restart) { G::w_init(w, lane); G::w_store_root(w, root_state + t, lane); }
    if (lane == 0) {
        atomicAdd(&stats->games, 1ULL);
        if (auto_restart) {
            tp.N[base] = 0; tp.W[base] = 0.0f; tp.P[base] = 0.0f; tp.first[base] = -1; tp.sub[base] = 0; tp.act[base] = -1; tp.nchild[base] = 0; tp.flags[base] = 0;
            tp.root[t] = 0; tp.alloc[t] = 1; tp.root_vl[t] = 0; tp.move_num[t] = 0; tp.game_id[t] += 1;
            tp.tflags[t] = TF_ACTIVE | (G::FIRST_FILL ? TF_FIRST_FILL : 0) | (noise_every_even_move ? TF_NEED_NOISE : 0);
            root_order_n[t] = default_order_n;
        } else tp.tflags[t] = tf & ~TF_ACTIVE;   // keeps TF_GAME_OVER for the host to see
    }
    if (auto_restart) for (int i = lane; i < default_order_n; i += 32) root_order[(size_t)t * G::MAX_CHILDREN + i] = default_order[i];
    if (auto_restart && tt.keys) { for (int i = lane; i < tt.cap; i += 32) tt.keys[(size_t)t * tt.cap + i] = 0; if (lane == 0) tt.count[t] = 0; }     // a new game gets a new table
}

// ------------------------------------------------------------------------------------------------
// Training examples from sample records: Dataset::extractExamples + augmentExample (src/selfplay/dataset.cpp:64-114, 245-436).
// One warp per (record, augmentation).  The 8 maps, in the reference's order, on tensor index (i, j) of an N x N plane:
// identity; rot90 (j, N-1-i); rot180 (N-1-i, N-1-j); rot270 (N-1-j, i); flipH (i, N-1-j); flipH after rot90 (j, i);
// flipH after rot180 (N-1-i, j); flipH after rot270 (N-1-j, N-1-i).  Every plane — the coordinate planes included — and the
// first N*N policy entries move by the same map.
AZ_D void dihedral(int aug, int n, int i, int j, int& i2, int& j2) {
    switch (aug) {
        case 1: i2 = j; j2 = n - 1 - i; break;
        case 2: i2 = n - 1 - i; j2 = n - 1 - j; break;
        case 3: i2 = n - 1 - j; j2 = i; break;
        case 4: i2 = i; j2 = n - 1 - j; break;
        case 5: i2 = j; j2 = i; break;
        case 6: i2 = n - 1 - i; j2 = j; break;
        case 7: i2 = n - 1 - j; j2 = n - 1 - i; break;
        default: i2 = i; j2 = j;
    }
}
template <class G>
__global__ void __launch_bounds__(128) k_make_examples(const Sample<G>* __restrict__ samples, int n, int k_aug, float* planes, float* policy, float* value) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= n * k_aug) return;
    const int rec = wid / k_aug, aug = wid % k_aug;
    typename G::Warp& w = warp_ws<G>(smem);
    const Sample<G>* sm = samples + rec;
    G::w_from_snapshot(w, &sm->state, lane);
    constexpr int N = G::N, CELLS = N * N, A = G::ACTIONS;
    float* pl = planes + (size_t)wid * G::PLANES * CELLS;
    for (int idx = lane; idx < G::PLANES * CELLS; idx += 32) {
        const int c = idx / CELLS, cell = idx % CELLS, i = cell / N, j = cell % N;
        int i2, j2; dihedral(aug, N, i, j, i2, j2);
        pl[(c * N + i2) * N + j2] = G::tensor_value(w, c, i, j);
    }
    float* po = policy + (size_t)wid * A;
    for (int a = lane; a < A; a += 32) po[a] = 0.0f;
    __syncwarp();
    const int total = G::policy_total(sm->visits, lane);
    G::policy_for_each(sm->visits, lane, [&](int a, int cnt) {
        int a2 = a;
        if (a < CELLS) { int i2, j2; dihedral(aug, N, a / N, a % N, i2, j2); a2 = i2 * N + j2; }
        if (a2 < A) po[a2] = total > 0 ? fdiv((float)cnt, (float)total) : 0.0f;
    });
    if (lane == 0) value[wid] = (float)sm->z;
}

// Dataset::extractExamples (src/selfplay/dataset.cpp:64-114) + augmentExample (:245-436) from GAME RECORDS (move lists), one warp per
// position: position (g, i) = the state after moves[g][0..i) → getEnhancedTensorRepresentation, the caller's policy vector for that
// position (any length P: the reference never looks at its meaning), value = the game's result seen from the player to move
// (:84-96).  The 7 extra images are produced the way the reference produces them: rot90 / rot180 / rot270 / flipH from the original,
// then flipH of the three rotations; a policy entry moves only when both its old and its new index are < P.
// kind: 0 rot90, 1 rot180, 2 rot270, 3 flipH (the oracle's orc_move)
__device__ __forceinline__ void example_move(int kind, int n, int i, int j, int& i2, int& j2) {
    if (kind == 0) { i2 = j; j2 = n - 1 - i; }
    else if (kind == 1) { i2 = n - 1 - i; j2 = n - 1 - j; }
    else if (kind == 2) { i2 = n - 1 - j; j2 = i; }
    else { i2 = i; j2 = n - 1 - j; }
}
template <class G>
__global__ void __launch_bounds__(128) k_examples_from_games(const int32_t* __restrict__ moves, const int32_t* __restrict__ pos_game, const int32_t* __restrict__ pos_ply,
                                                            int n_pos, int max_moves, const int8_t* __restrict__ results, const float* __restrict__ policy_in, int P, int k_aug,
                                                            float* planes, float* policy, float* value, uint64_t* hist_scratch, int32_t* error_flag) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= n_pos) return;
    typename G::Warp& w = warp_ws<G>(smem);
    const int g = pos_game[wid], ply = pos_ply[wid];
    G::w_init(w, lane);
    G::w_attach_history(w, hist_scratch + (size_t)wid * (max_moves + 1), lane);
    for (int i = 0; i < ply; ++i)
        if (!G::w_apply(w, moves[(size_t)g * max_moves + i], lane, true)) { if (lane == 0) atomicExch(error_flag, 1 + g); return; }     // makeMove throws (dataset.cpp:76)
    constexpr int N = G::N, CELLS = N * N;
    const size_t pe = (size_t)G::PLANES * CELLS;
    float* pl0 = planes + (size_t)wid * k_aug * pe;
    float* po0 = policy + (size_t)wid * k_aug * P;
    G::w_planes(w, lane, pl0);
    const float* pin = policy_in + (size_t)wid * P;
    for (int a = lane; a < P; a += 32) po0[a] = pin[a];
    const int res = results[g], player = G::w_player(w);
    float gv = res == RES_WIN_P1 ? 1.0f : (res == RES_WIN_P2 ? -1.0f : 0.0f);
    if (player == 2) gv = -gv;
    if (lane == 0) for (int k = 0; k < k_aug; ++k) value[(size_t)wid * k_aug + k] = gv;
    __syncwarp();
    for (int k = 1; k < k_aug; ++k) {
        const int src = k <= 4 ? 0 : k - 4, kind = k <= 3 ? k - 1 : 3;
        const float* spl = pl0 + (size_t)src * pe; const float* spo = po0 + (size_t)src * P;
        float* dpl = pl0 + (size_t)k * pe; float* dpo = po0 + (size_t)k * P;
        for (int a = lane; a < P; a += 32) dpo[a] = spo[a];
        __syncwarp();
        for (int idx = lane; idx < (int)pe; idx += 32) {
            const int c = idx / CELLS, cell = idx % CELLS;
            int i2, j2; example_move(kind, N, cell / N, cell % N, i2, j2);
            dpl[(c * N + i2) * N + j2] = spl[idx];
        }
        for (int oi = lane; oi < CELLS; oi += 32) {
            int i2, j2; example_move(kind, N, oi / N, oi % N, i2, j2);
            const int ni = i2 * N + j2;
            if (oi < P && ni < P) dpo[ni] = spo[oi];
        }
        __syncwarp();
    }
}

}  // namespace az
