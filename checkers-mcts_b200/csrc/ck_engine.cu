// ck_engine.cu -- K2/K5/K6: batched PUCT self-play / arena engine, one warp per game.
//
// Replaces the reference's sequential Python loop
//   generate_Checkers_data._generate_data (training_pipeline.py:334-419)
//   tournament_Checkers._start_tournament (:505-559)
//   MCTS.begin_tree_search / tree_policy / select_child / backpropagation / best_child /
//   new_root_node (MCTS.py:59-295), MCTS_Node (:345-430)
// with a lock-step device loop: every round runs ONE tree kernel (consume the previous
// evaluation -> expand + backup -> play moves / re-root when a search is complete -> PUCT
// descent to the next leaf -> stage its network input) and ONE batched evaluation of all
// staged leaves.  Simulations that end in a terminal child need no evaluation and are
// finished inside the tree kernel.
//
// HBM layout (per engine):
//   pos [slot][3 buffers][cap]  uint4  p1,p2,k,meta                      16 B / node
//   stat[slot][3 buffers][cap]  uint4  N (u32), W (f32), P (f32), link   16 B / node
//     link = first_child[0:22) | n_children[22:28) | status[28:30) | parent_player[30]
//   Each game owns two trees (one per colour, training_pipeline.py:353,372) living in two
//   of its three buffers; the third is the to-space of the re-root compaction (K5).
//   Children of a node are contiguous and stored in node.children order (legal list
//   reversed, MCTS.py:72-75), so a PUCT step is one coalesced 16 B x b load.
//   hist[slot][max_plies+1] ck_pos, path[slot][128] u32, leaves/policy/value per net.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <thread>
#include <vector>
#include "ck_common.cuh"
#include "ck_device_fn.cuh"
#include "ck_net.cuh"
#include "ck_rules.cuh"

namespace ck {

constexpr int kMaxDepth = 128;
constexpr int kWarpsPerBlock = 4;
constexpr uint32_t kFcMask = 0x3FFFFFu;
constexpr int kPowTable = 1 << 17;
constexpr int kNoIdx = 1 << 30;

enum { PH_NEED_ROOT = 0, PH_SEARCH = 1, PH_HALT = 2 };

struct Slot {
    int32_t game;            // local game index, -1 idle
    int32_t hist_len, move_count, phase, cur, sims_done;
    int32_t pend_leaf, pend_row, pend_depth, pend_net;
    int32_t buf[2], scratch;
    int32_t root[2], alloc[2], best[2], exists[2];
    int32_t nrec, misses, search_id, manual, manual_target;
    int32_t stagger_until;   // plies of this game that are searched with cfg.stagger_budget (warm start, first game of a slot only)
    double tau;
    unsigned long long tot_sims, tot_evals;      // monotonic per slot; per-game totals are deltas
    unsigned long long g_sims0, g_evals0;
    unsigned long long tot_hits;                 // expansions served by the evaluation cache (a subset of tot_evals)
};

struct Counters {
    unsigned long long games_finished, moves, nodes, compactions;
    int32_t next_game, error, halted, pad0;
    // per evaluator batch (index = network in an arena, slot group in overlapped self-play, else 0):
    int32_t active[2];       // slots that still have work after this round
    int32_t batch_count[2];  // leaves staged this round
    int32_t deferred[2];     // leaves that asked for a batch row beyond leaf_cap this round (they ask again next round)
    int32_t leaf_cap[2];     // rows the evaluator batch of this round may hold (round_begin_kernel)
};

struct EngineDev {
    ck_engine_cfg cfg;
    int32_t n_slots, cap, max_plies, max_rec, n_games, compact_need, max_term, arena_half;
    float one_minus_eps;
    uint4 *pos, *stat;
    ck_pos *hist;
    uint32_t *path;
    Slot *slots;
    Counters *ctr;
    ck_leaf *leaves[2];
    float *policy[2], *value[2];
    ck_record *rec;
    ck_game_result *results;
    const double *pow_half;      // host libm pow(n, 0.5) table (MCTS.py:110 uses n ** 0.5)
    const double *log_tab;       // host libm log(n) table (np.log(node.n), MCTS.py:114); rollout evaluators only
    int32_t uct;                 // 1: NEURAL_NET=False tree policy (UCT, one child per visit, playouts)
    uint32_t round;              // lock-step round counter (keys the random playouts)
    uint4 *cache;                // evaluation cache [slot][cache_entries][8 x uint4]; nullptr: off
    int32_t cache_entries;       // per slot, a power of two
    int32_t max_chain;           // simulations a slot may complete inside one round without a network evaluation (terminal + cached)
    int32_t cache_game_tag;      // 1: the evaluator depends on the game (salted stubs), so entries are tagged with it
    int32_t wave;                // positions one full wave of the tower evaluates (tiles x positions per tile x SMs); 0: no batch shaping
    int32_t wave_slack10;        // the batch is cut back to whole waves while the excess is below wave_slack10 / 10 of a wave
    // overlapped self-play: the slots form two groups; while the tower evaluates the leaves of one group the tree
    // kernel of the other runs next to it on the same SMs.  Per launch: which group, and its slot range.
    int32_t groups;              // 1 or 2
    int32_t grp, slot0, slot_n;
};

// ---- small device helpers --------------------------------------------------------------
__device__ __forceinline__ ck_pos to_pos(const uint4 v) { ck_pos p; p.p1 = v.x; p.p2 = v.y; p.k = v.z; p.meta = v.w; return p; }
__device__ __forceinline__ uint4 from_pos(const ck_pos &p) { return make_uint4(p.p1, p.p2, p.k, p.meta); }
__device__ __forceinline__ int link_nchild(uint32_t l) { return (int)((l >> 22) & 63u); }
__device__ __forceinline__ int link_status(uint32_t l) { return (int)((l >> 28) & 3u); }
__device__ __forceinline__ int link_pp(uint32_t l) { return (int)((l >> 30) & 1u); }
__device__ __forceinline__ bool same_state(const ck_pos &a, const ck_pos &b) {
    return a.p1 == b.p1 && a.p2 == b.p2 && a.k == b.k && ((a.meta ^ b.meta) & 1u) == 0;
}
__device__ __forceinline__ uint4 shfl4(uint4 v, int src) {
    v.x = __shfl_sync(CK_FULL, v.x, src); v.y = __shfl_sync(CK_FULL, v.y, src);
    v.z = __shfl_sync(CK_FULL, v.z, src); v.w = __shfl_sync(CK_FULL, v.w, src);
    return v;
}
__device__ __forceinline__ double shfl_xor_d(double v, int m) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(CK_FULL, lo, m); hi = __shfl_xor_sync(CK_FULL, hi, m);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_up_d(double v, int d) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_up_sync(CK_FULL, lo, d); hi = __shfl_up_sync(CK_FULL, hi, d);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_d(double v, int src) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(CK_FULL, lo, src); hi = __shfl_sync(CK_FULL, hi, src);
    return __hiloint2double(hi, lo);
}

struct WarpCtx {
    const EngineDev &E;
    Slot &S;                 // shared-memory copy of the slot
    uint32_t *path;          // global, kMaxDepth entries
    ck_pos *hist;            // global, max_plies + 1 entries
    int slot, lane;
    __device__ uint4 *pos_buf(int buf) const { return E.pos + ((int64_t)slot * 3 + buf) * E.cap; }
    __device__ uint4 *stat_buf(int buf) const { return E.stat + ((int64_t)slot * 3 + buf) * E.cap; }
    __device__ uint4 *pos_of(int t) const { return pos_buf(S.buf[t]); }
    __device__ uint4 *stat_of(int t) const { return stat_buf(S.buf[t]); }
};

__device__ __forceinline__ int global_game(const EngineDev &E, int local) {
    return E.cfg.game_id_base + local * (E.cfg.game_id_stride ? E.cfg.game_id_stride : 1);
}
__device__ __forceinline__ uint64_t game_key(const EngineDev &E, int local) {
    return mix64(E.cfg.seed ^ mix64((uint64_t)(uint32_t)global_game(E, local) + 0x51ED270B1ull));
}
__device__ __forceinline__ int net_of(const EngineDev &E, int local_game, int player) {
    if (!E.cfg.arena) return E.grp;                 // evaluator batch index: the slot group (0 unless overlapped)
    const int p1_net = global_game(E, local_game) < E.arena_half ? 0 : 1;   // training_pipeline.py:523-528
    return player == 0 ? p1_net : 1 - p1_net;
}

__device__ void dev_error(const EngineDev &E, int code) { atomicCAS(&E.ctr->error, 0, code); }

// MCTS_Node.backpropagation + MCTS.determine_reward (MCTS.py:148-186, 419-430): every node
// on the path gets N += 1 and W += reward seen from the player to move in its PARENT.
__device__ void backup(const WarpCtx &c, int depth, bool is_outcome, int outcome, float value, int leaf_player) {
    uint4 *stat = c.stat_of(c.S.cur);
    for (int d = c.lane; d <= depth; d += 32) {
        const uint32_t x = c.path[d];
        uint4 st = stat[x];
        const int pp = link_pp(st.w);
        float reward;
        if (is_outcome) {
            if (outcome == CK_P1_WINS) reward = pp == 0 ? 1.f : -1.f;
            else if (outcome == CK_P2_WINS) reward = pp == 1 ? 1.f : -1.f;
            else reward = 0.f;
        } else {
            reward = leaf_player != pp ? -value : value;
        }
        st.x += 1u;
        st.y = __float_as_uint(__fadd_rn(__uint_as_float(st.y), reward));
        *reinterpret_cast<uint2 *>(&stat[x]) = make_uint2(st.x, st.y);
    }
    __syncwarp();
}

// Marsaglia-Tsang gamma(alpha) from Philox words (np.random.dirichlet, MCTS.py:108);
// statistical parity only (SURVEY 8b RNG row)
// Exp(1) from one random word.  A 24-bit uniform and a single-precision logarithm are ample for exploration
// noise (the sum that enters PUCT is formed in float64 as in the reference); the float64 log was a third of
// the instructions of a descent.
__device__ __forceinline__ float exp1_from(uint32_t w) {
    return -__logf((float)((w >> 8) + 1u) * (1.0f / 16777216.0f));
}

__device__ double gamma_sample(const Philox &rng, uint32_t c0, uint32_t c1, uint32_t c2, double alpha) {
    uint32_t r[4];
    if (alpha == 1.0) { rng(c0, c1, c2, 0x44495231u, r); return (double)exp1_from(r[0]); }
    const double a = alpha < 1.0 ? alpha + 1.0 : alpha;
    const double d = a - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * d);
    double out = d;
    for (uint32_t it = 0; it < 64; ++it) {
        rng(c0, c1, c2, 0x47414D00u + it, r);
        const double u1 = u01(r[0], r[1]), u2 = u01(r[2], r[3]);
        const double x = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
        double v = 1.0 + cc * x;
        if (v <= 0) continue;
        v = v * v * v;
        uint32_t q[4];
        rng(c0, c1, c2, 0x47414E00u + it, q);
        if (log(u01(q[0], q[1])) < 0.5 * x * x + d - d * v + d * log(v)) { out = d * v; break; }
    }
    if (alpha < 1.0) { rng(c0, c1, c2, 0x47414F00u, r); out *= pow(u01(r[0], r[1]), 1.0 / alpha); }
    return out;
}

// PUCT score of one child (MCTS.select_child, MCTS.py:101-116) with the dtypes numpy >= 2
// produces: q = f32(W/N), (1-eps)*P in f32, everything else left-to-right in f64.
__device__ __forceinline__ double puct(const EngineDev &E, const uint4 cst, double noise, double sqrt_n) {
    const uint32_t n = cst.x;
    const float q = n ? __fdiv_rn(__uint_as_float(cst.y), (float)n) : 0.f;
    const float scaled = __fmul_rn(E.one_minus_eps, __uint_as_float(cst.z));
    const double psa = __dadd_rn((double)scaled, __dmul_rn(E.cfg.epsilon, noise));
    const double t = __ddiv_rn(__dmul_rn(__dmul_rn(E.cfg.uct_c, psa), sqrt_n), (double)(1u + n));
    return __dadd_rn((double)q, t);
}

// One descent from the root (MCTS.tree_policy, MCTS.py:59-99).  Returns the leaf to evaluate
// (depth in *out_depth), -1 when the descent ended in a terminal child (already backed up),
// -2 on error.
__device__ int select_leaf(const WarpCtx &c, int *out_depth) {
    const EngineDev &E = c.E;
    const int t = c.S.cur;
    const uint4 *stat = c.stat_of(t);
    const Philox rng(game_key(E, c.S.game));
    int node = c.S.root[t], depth = 0;
    uint32_t noise_a[4] = {0u, 0u, 0u, 0u};
    if (c.lane == 0) c.path[0] = (uint32_t)node;
    uint4 st = stat[node];
    for (;;) {
        const int b = link_nchild(st.w);
        if (b == 0) { *out_depth = depth; __syncwarp(); return node; }
        const int fc = (int)(st.w & kFcMask);
        const uint32_t pn = st.x;
        const double sqrt_n = pn < (uint32_t)kPowTable ? E.pow_half[pn] : sqrt((double)pn);
        uint4 c0 = make_uint4(0, 0, 0, 0), c1 = c0;
        if (c.lane < b) c0 = stat[fc + c.lane];
        if (c.lane + 32 < b) c1 = stat[fc + 32 + c.lane];
        double n0 = 0.0, n1 = 0.0;
        if (E.cfg.epsilon != 0.0) {                      // fresh Dirichlet(alpha) at every node, every visit (:107-108)
            // the noise is exploration only (statistical parity, SURVEY 8b RNG row): gamma variates and their
            // normalisation in single precision; the sum that enters PUCT is formed in float64 as in the reference
            const uint32_t cc0 = (uint32_t)c.S.search_id, cc1 = (uint32_t)c.S.sims_done;
            float g0, g1 = 0.f;
            if (E.cfg.alpha == 1.0) {
                // Gamma(1) = Exp(1): one Philox block (four words) serves a child slot for four consecutive
                // levels of the descent -- word depth & 3 of the block keyed by (search, simulation, depth / 4, child)
                if ((depth & 3) == 0) {
                    rng(cc0, cc1, (uint32_t)((depth >> 2) << 8 | c.lane), 0x44495231u, noise_a);
                }
                const uint32_t w = (depth & 2) ? ((depth & 1) ? noise_a[3] : noise_a[2]) : ((depth & 1) ? noise_a[1] : noise_a[0]);
                g0 = c.lane < b ? exp1_from(w) : 0.f;
                // more than 32 children is rare: those slots draw their own block per level
                if (c.lane + 32 < b) g1 = (float)gamma_sample(rng, cc0, cc1, (uint32_t)(depth << 8 | (c.lane + 32)), 1.0);
            } else {
                g0 = c.lane < b ? (float)gamma_sample(rng, cc0, cc1, (uint32_t)(depth << 8 | c.lane), E.cfg.alpha) : 0.f;
                g1 = c.lane + 32 < b ? (float)gamma_sample(rng, cc0, cc1, (uint32_t)(depth << 8 | (c.lane + 32)), E.cfg.alpha) : 0.f;
            }
            float tot = g0 + g1;
#pragma unroll
            for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(CK_FULL, tot, o);
            const float inv = __fdividef(1.0f, tot);
            n0 = (double)(g0 * inv); n1 = (double)(g1 * inv);
        }
        // arg-max with np.argmax's tie rule (first maximum, MCTS.py:116): butterfly maximum of the scores -- only the
        // rounds the child count needs, lanes beyond it hold -inf --, then the first lane that holds it
        const double kNegInf = __longlong_as_double(0xFFF0000000000000ll);
        const double u0 = c.lane < b ? puct(E, c0, n0, sqrt_n) : kNegInf;
        const double u1 = c.lane + 32 < b ? puct(E, c1, n1, sqrt_n) : kNegInf;
        double m = fmax(u0, u1);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
            if (o < b) m = fmax(m, shfl_xor_d(m, o));                               // b is warp-uniform
        const uint32_t lt = b >= 32 ? 0xFFFFFFFFu : ((1u << b) - 1u);
        const uint32_t first0 = __ballot_sync(CK_FULL, u0 == m) & lt;
        const uint32_t first1 = b > 32 ? __ballot_sync(CK_FULL, u1 == m) & ((1u << (b - 32)) - 1u) : 0u;
        // scores are never NaN unless the network left its range (reported as CK_ERR_NET_RANGE): child 0 then
        int best_i = first0 ? __ffs((int)first0) - 1 : first1 ? 32 + __ffs((int)first1) - 1 : 0;
        const uint4 pick = best_i < 32 ? shfl4(c0, best_i) : shfl4(c1, best_i - 32);
        node = fc + best_i;
        ++depth;
        if (depth >= kMaxDepth) { dev_error(E, CK_ERR_DEPTH); *out_depth = depth - 1; return -2; }
        if (c.lane == 0) c.path[depth] = (uint32_t)node;
        const int cs = link_status(pick.w);
        if (cs != CK_ONGOING) {                         // terminal child: no evaluation (:93-94,145-146)
            __syncwarp();
            backup(c, depth, true, cs, 0.f, 0);
            return -1;
        }
        st = pick;
    }
}

// stage the network input of a leaf (Checkers.predict, Checkers.py:431-432).  The batch of a round is shaped to whole
// waves of the tower kernel (round_begin_kernel): a leaf that asks for a row beyond the round's cap is not staged and
// the slot repeats the same descent in the next round -- deterministic (the tree did not change and the exploration
// noise is keyed by search and simulation number), so results do not depend on which slots had to wait.
__device__ void stage_leaf(const WarpCtx &c, int leaf, int depth) {
    const EngineDev &E = c.E;
    Slot &S = c.S;
    __syncwarp();
    if (c.lane == 0) {
        const int net = net_of(E, S.game, S.cur);
        if (E.wave > 0) {
            const int row = atomicAdd(&E.ctr->batch_count[net], 1);
            if (row >= *(volatile int32_t *)&E.ctr->leaf_cap[net]) {
                atomicSub(&E.ctr->batch_count[net], 1);
                atomicAdd(&E.ctr->deferred[net], 1);
                leaf = -1;
            } else {
                S.pend_row = row;
            }
        } else {
            S.pend_row = atomicAdd(&E.ctr->batch_count[net], 1);
        }
    }
    leaf = __shfl_sync(CK_FULL, leaf, 0);
    if (leaf < 0) { __syncwarp(); return; }
    if (c.lane == 0) {
        const int net = net_of(E, S.game, S.cur);
        const ck_pos p = to_pos(c.pos_of(S.cur)[leaf]);
        ck_leaf L;
        int p5;
        const int cnt = gen_moves(p, NullSink{}, L.mask);
        outcome_of(p, cnt > 0, &p5);
        L.p1 = p.p1; L.p2 = p.p2; L.k = p.k;
        L.info = (p.meta & 1u) | ((uint32_t)p5 << 8) | (((uint32_t)global_game(E, S.game) & 0xFFFFu) << 16);
        E.leaves[net][S.pend_row] = L;
        S.pend_leaf = leaf; S.pend_depth = depth; S.pend_net = net;
    }
    __syncwarp();
}

// ---- evaluation cache ---------------------------------------------------------------------------
// Half of a game's leaf evaluations repeat an earlier evaluation of the same network input: each colour
// searches its own tree (training_pipeline.py:353,372) over nearly the same positions, and a tree holds
// transpositions as separate nodes (scripts/dup_rate.py: 46-50 % of the evaluations of a cfg2 game, median
// distance 150-230 evaluations).  The network is a pure function of (position, side to move, plane 5) --
// the legal-action planes follow from the position -- so what an expansion needs from it, the children's
// priors (Checkers.predict's masked, renormalised policy, Checkers.py:434-452) and the value, is kept in a
// direct-mapped table per slot: 128 B per entry = key (p1, p2, k, tag) | value | up to 24 priors in
// node.children order.  A hit expands the leaf inside the same round with bit-identical numbers; results do
// not depend on the cache (tests run with and without it).  One warp owns a slot, so there are no races.
constexpr int kCachePriors = 24;
constexpr int kCacheU4 = 8;                          // uint4 per entry
__device__ __forceinline__ int p5_ongoing(uint32_t meta) {      // plane-5 numerator of a position that is not a draw (outcome_of)
    return meta_ply(meta) + 1 >= 80 ? meta_rev(meta) + 1 : 0;
}
__device__ __forceinline__ uint32_t cache_tag(const EngineDev &E, const Slot &S, uint32_t meta, int net) {
    return 0x80000000u | (meta & 1u) | ((uint32_t)p5_ongoing(meta) << 1) | ((uint32_t)net << 8) |
           (E.cache_game_tag ? (((uint32_t)S.game & 0x3FFFFFu) << 9) : 0u);
}
__device__ __forceinline__ uint4 *cache_entry(const WarpCtx &c, const ck_pos &p, uint32_t tag) {
    const uint32_t h = mix32(p.p1 * 0x9E3779B1u ^ mix32(p.p2 ^ mix32(p.k + tag * 0x85EBCA6Bu)));
    return c.E.cache + ((int64_t)c.slot * c.E.cache_entries + (int64_t)(h & (uint32_t)(c.E.cache_entries - 1))) * kCacheU4;
}

// Expansion of a leaf (MCTS.py:71-77): all children at once in reversed legal order, each with its
// own terminal test (MCTS_Node.__init__, :374-375), priors from the masked + renormalised policy
// (Checkers.py:434-452), then the signed value backup.  kHit: priors and value come from the slot's
// evaluation cache (`entry`), otherwise from the evaluator's output row, and the entry is (re)written.
template <bool kHit>
__device__ void expand_leaf(const WarpCtx &c, int leaf, int depth, int net, int row, uint4 *entry, uint32_t tag) {
    const EngineDev &E = c.E;
    Slot &S = c.S;
    const int t = S.cur;
    uint4 *pos = c.pos_of(t), *stat = c.stat_of(t);
    const ck_pos lp = to_pos(pos[leaf]);
    // every lane derives the legal-action planes itself (a few dozen bit operations); the successors are
    // then built one per lane (kth_move + make_child_fast) instead of serially by lane 0
    uint32_t mask[8];
    const int b = gen_moves(lp, NullSink{}, mask);
    const bool jump = (mask[4] | mask[5] | mask[6] | mask[7]) != 0;
    const uint32_t use[4] = {jump ? mask[4] : mask[0], jump ? mask[5] : mask[1], jump ? mask[6] : mask[2], jump ? mask[7] : mask[3]};
    const Side lsd = side_of(lp);
    uint32_t hop[4];
    hop_sets(lsd, hop);
    const float *prow = nullptr;
    float psum = 1.f, value;
    if (kHit) {
        value = __uint_as_float(entry[1].x);
    } else {
        prow = E.policy[net] + (int64_t)row * CK_POLICY_SIZE;
        float masked[16];
        psum = masked_policy_sum(prow, mask, c.lane, masked);
        value = E.value[net][row];
    }
    const int fc = S.alloc[t];
    if (fc + b > E.cap) {
        dev_error(E, CK_ERR_POOL_OVERFLOW);
        if (c.lane == 0) { S.pend_leaf = -1; S.phase = PH_HALT; }
        __syncwarp();
        return;
    }
    const bool fill = !kHit && entry != nullptr && b <= kCachePriors;
    float *eprior = entry ? reinterpret_cast<float *>(entry + 2) : nullptr;
    const int lplayer = meta_player(lp.meta);
    for (int i = c.lane; i < b; i += 32) {
        int ms, md;
        kth_move(lsd, use, jump, b - 1 - i, &ms, &md);                     // node.children = legal list reversed (:72-75)
        const ck_pos ch = make_child_fast(lp, lsd, hop, ms, md, jump);
        int p5;
        const int stc = status_of(ch, &p5);
        const float prior = kHit ? eprior[i] : __fdiv_rn(prow[meta_action(ch.meta)], psum);
        if (fill) eprior[i] = prior;
        pos[fc + i] = from_pos(ch);
        stat[fc + i] = make_uint4(0u, __float_as_uint(0.f), __float_as_uint(prior),
                                  ((uint32_t)stc << 28) | ((uint32_t)lplayer << 30));
    }
    __syncwarp();
    if (c.lane == 0) {
        const uint32_t old = stat[leaf].w;
        stat[leaf].w = (old & 0xF0000000u) | ((uint32_t)b << 22) | (uint32_t)fc;
        S.alloc[t] = fc + b;
        S.pend_leaf = -1;
        S.tot_evals += 1; S.tot_sims += 1; S.sims_done += 1;
        if (kHit) S.tot_hits += 1;
        if (fill) {
            entry[1] = make_uint4(__float_as_uint(value), (uint32_t)b, 0u, 0u);
            entry[0] = make_uint4(lp.p1, lp.p2, lp.k, tag);
        }
        atomicAdd(&E.ctr->nodes, (unsigned long long)b);
    }
    __syncwarp();
    backup(c, depth, false, 0, value, lplayer);
}

// the evaluation the previous round staged has arrived
__device__ void expand_pending(const WarpCtx &c) {
    Slot &S = c.S;
    uint4 *entry = nullptr;
    uint32_t tag = 0u;
    if (c.E.cache) {
        const ck_pos lp = to_pos(c.pos_of(S.cur)[S.pend_leaf]);
        tag = cache_tag(c.E, S, lp.meta, S.pend_net);
        entry = cache_entry(c, lp, tag);
    }
    expand_leaf<false>(c, S.pend_leaf, S.pend_depth, S.pend_net, S.pend_row, entry, tag);
}

// leaf found by the descent: expand it from the cache if its input was evaluated before
__device__ bool expand_cached(const WarpCtx &c, int leaf, int depth) {
    const Slot &S = c.S;
    const ck_pos lp = to_pos(c.pos_of(S.cur)[leaf]);
    const int net = net_of(c.E, S.game, S.cur);
    const uint32_t tag = cache_tag(c.E, S, lp.meta, net);
    uint4 *entry = cache_entry(c, lp, tag);
    const uint4 key = entry[0];
    if (key.x != lp.p1 || key.y != lp.p2 || key.z != lp.k || key.w != tag) return false;
    expand_leaf<true>(c, leaf, depth, net, 0, entry, tag);
    return true;
}

// ---- NEURAL_NET=False (the reference's iteration-0 self-play, MCTS.py:78-89,113-115,132-146) ---------
// Plain UCT: a node with unvisited successors adds ONE child per visit and plays the game out from it;
// a fully expanded node picks argmax q + 2c * (2 ln N / n) ** 0.5 in float64 (rewards are integers there,
// so q is an int / int quotient).  The block of all successors is laid out at the node's first visit
// (creating a node has no side effect in the reference beyond its own move generation) and the node's
// otherwise unused prior field counts how many of them exist for the search: children [0, count).
__device__ __forceinline__ int uct_expanded(const uint4 st) { return min((int)st.z, link_nchild(st.w)); }

__device__ int select_leaf_uct(const WarpCtx &c, int *out_depth) {
    const EngineDev &E = c.E;
    Slot &S = c.S;
    const int t = S.cur;
    uint4 *pos = c.pos_of(t), *stat = c.stat_of(t);
    int node = S.root[t], depth = 0;
    if (c.lane == 0) c.path[0] = (uint32_t)node;
    for (;;) {
        const uint4 st = stat[node];
        int b = link_nchild(st.w), fc = (int)(st.w & kFcMask), k = (int)st.z;
        if (b == 0) {                                    // first visit: lay out the successor block
            const ck_pos lp = to_pos(pos[node]);
            uint32_t mask[8];
            b = gen_moves(lp, NullSink{}, mask);
            const bool jump = (mask[4] | mask[5] | mask[6] | mask[7]) != 0;
            const uint32_t use[4] = {jump ? mask[4] : mask[0], jump ? mask[5] : mask[1], jump ? mask[6] : mask[2], jump ? mask[7] : mask[3]};
            const Side lsd = side_of(lp);
            uint32_t hop[4];
            hop_sets(lsd, hop);
            fc = S.alloc[t];
            if (b == 0 || fc + b > E.cap) {
                dev_error(E, b == 0 ? CK_ERR_STATE : CK_ERR_POOL_OVERFLOW);
                *out_depth = depth;
                return -2;
            }
            const int lplayer = meta_player(lp.meta);
            for (int i = c.lane; i < b; i += 32) {
                int ms, md;
                kth_move(lsd, use, jump, b - 1 - i, &ms, &md);             // unvisited_child_states.pop(): last first
                const ck_pos ch = make_child_fast(lp, lsd, hop, ms, md, jump);
                int p5;
                const int stc = status_of(ch, &p5);
                pos[fc + i] = from_pos(ch);
                stat[fc + i] = make_uint4(0u, __float_as_uint(0.f), 0u, ((uint32_t)stc << 28) | ((uint32_t)lplayer << 30));
            }
            __syncwarp();
            if (c.lane == 0) {
                stat[node].w = (st.w & 0xF0000000u) | ((uint32_t)b << 22) | (uint32_t)fc;
                S.alloc[t] = fc + b;
            }
            k = 0;
        }
        if (k < b) {                                     // add the next child and simulate from it (:78-89)
            const int child = fc + k;
            ++depth;
            if (depth >= kMaxDepth) { dev_error(E, CK_ERR_DEPTH); *out_depth = depth - 1; return -2; }
            if (c.lane == 0) {
                stat[node].z = (uint32_t)(k + 1);
                c.path[depth] = (uint32_t)child;
                atomicAdd(&E.ctr->nodes, 1ull);
            }
            __syncwarp();
            const int cs = link_status(stat[child].w);
            if (cs != CK_ONGOING) { backup(c, depth, true, cs, 0.f, 0); return -1; }   // a terminal node is its own outcome
            *out_depth = depth;
            return child;
        }
        const uint32_t pn = st.x;
        const double ln_n = pn < (uint32_t)kPowTable ? E.log_tab[pn] : log((double)pn);
        const double two_ln = __dmul_rn(2.0, ln_n), two_c = __dmul_rn(2.0, E.cfg.uct_c);
        uint4 c0 = make_uint4(0, 0, 0, 0), c1 = c0;
        if (c.lane < b) c0 = stat[fc + c.lane];
        if (c.lane + 32 < b) c1 = stat[fc + 32 + c.lane];
        double best_u = 0.0;
        int best_i = kNoIdx;
        if (c.lane < b) {
            const double q = c0.x ? __ddiv_rn((double)__uint_as_float(c0.y), (double)c0.x) : 0.0;
            best_u = __dadd_rn(q, __dmul_rn(two_c, sqrt(__ddiv_rn(two_ln, (double)c0.x))));
            best_i = c.lane;
        }
        if (c.lane + 32 < b) {
            const double q = c1.x ? __ddiv_rn((double)__uint_as_float(c1.y), (double)c1.x) : 0.0;
            const double u1 = __dadd_rn(q, __dmul_rn(two_c, sqrt(__ddiv_rn(two_ln, (double)c1.x))));
            if (u1 > best_u) { best_u = u1; best_i = c.lane + 32; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double ou = shfl_xor_d(best_u, o);
            const int oi = __shfl_xor_sync(CK_FULL, best_i, o);
            const bool take = (oi != kNoIdx) && (best_i == kNoIdx || ou > best_u || (ou == best_u && oi < best_i));
            if (take) { best_u = ou; best_i = oi; }
        }
        const uint32_t plink = best_i < 32 ? __shfl_sync(CK_FULL, c0.w, best_i) : __shfl_sync(CK_FULL, c1.w, best_i - 32);
        node = fc + best_i;
        ++depth;
        if (depth >= kMaxDepth) { dev_error(E, CK_ERR_DEPTH); *out_depth = depth - 1; return -2; }
        if (c.lane == 0) c.path[depth] = (uint32_t)node;
        const int cs = link_status(plink);
        if (cs != CK_ONGOING) {
            __syncwarp();
            backup(c, depth, true, cs, 0.f, 0);
            return -1;
        }
    }
}

// stage the start of a playout: position with its draw-rule counters (meta rides in mask[0])
__device__ void stage_playout(const WarpCtx &c, int leaf, int depth) {
    const EngineDev &E = c.E;
    Slot &S = c.S;
    __syncwarp();
    if (c.lane == 0) {
        const ck_pos p = to_pos(c.pos_of(S.cur)[leaf]);
        ck_leaf L;
        L.p1 = p.p1; L.p2 = p.p2; L.k = p.k;
        L.info = (p.meta & 1u) | (((uint32_t)global_game(E, S.game) & 0xFFFFu) << 16);
        L.mask[0] = p.meta;
        L.mask[1] = (uint32_t)global_game(E, S.game);    // the full game id keys the playout's random stream
#pragma unroll
        for (int i = 2; i < 8; ++i) L.mask[i] = 0u;
        const int row = atomicAdd(&E.ctr->batch_count[0], 1);
        E.leaves[0][row] = L;
        S.pend_leaf = leaf; S.pend_row = row; S.pend_depth = depth; S.pend_net = 0;
    }
    __syncwarp();
}

// the playout's outcome string is backed up from the node it started at (MCTS_Node.simulation, :412-417)
__device__ void finish_playout(const WarpCtx &c) {
    Slot &S = c.S;
    const int outcome = (int)c.E.value[0][S.pend_row];
    const int depth = S.pend_depth;
    __syncwarp();
    if (c.lane == 0) { S.pend_leaf = -1; S.tot_evals += 1; S.tot_sims += 1; S.sims_done += 1; }
    __syncwarp();
    backup(c, depth, true, outcome, 0.f, 0);
}

// K5: copy the subtree under `root` of tree t into the scratch buffer in BFS order and make
// the scratch buffer the tree's buffer.  32 nodes per iteration: each lane relocates the
// child block of one node; block positions come from a warp prefix sum.
__device__ void compact_tree(const WarpCtx &c, int t, int root) {
    Slot &S = c.S;
    const uint4 *spos = c.pos_of(t), *sstat = c.stat_of(t);
    uint4 *dpos = c.pos_buf(S.scratch), *dstat = c.stat_buf(S.scratch);
    if (c.lane == 0) { dpos[0] = spos[root]; dstat[0] = sstat[root]; }
    __syncwarp();
    int count = 1;
    for (int head = 0; head < count;) {
        const int end = min(head + 32, count);           // nodes appended below are handled by later rounds
        const int i = head + c.lane;
        const bool valid = i < end;
        const uint32_t link = valid ? dstat[i].w : 0u;
        const int nch = link_nchild(link), ofc = (int)(link & kFcMask);
        int incl = nch;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(CK_FULL, incl, o);
            if (c.lane >= o) incl += v;
        }
        const int total = __shfl_sync(CK_FULL, incl, 31);
        const int excl = incl - nch;
        if (nch > 0) dstat[i].w = (link & ~kFcMask) | (uint32_t)(count + excl);
        // The children of this round's (up to 32) parents go to [count, count + total) in parent order.
        // One lane per destination node: the parent is found by a shuffle binary search over the
        // exclusive prefix sums, and four steps of loads are issued before their stores so that the
        // round costs about two memory round trips instead of one per child of the widest parent
        // (the serial per-lane copy made a re-rooting slot the tail of every tree_step launch).
        for (int j0 = 0; j0 < total; j0 += 128) {
            uint4 vp[4], vs[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + 32 * u + c.lane;
                int lo = 0;
#pragma unroll
                for (int step = 16; step; step >>= 1) {
                    const int e = __shfl_sync(CK_FULL, excl, (lo + step) & 31);
                    if (lo + step < 32 && e <= j) lo += step;                // largest parent with excl <= j
                }
                const int src = __shfl_sync(CK_FULL, ofc, lo) + (j - __shfl_sync(CK_FULL, excl, lo));
                if (j < total) { vp[u] = spos[src]; vs[u] = sstat[src]; }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + 32 * u + c.lane;
                if (j < total) { dpos[count + j] = vp[u]; dstat[count + j] = vs[u]; }
            }
        }
        count += total;
        head = end;
        __syncwarp();
    }
    __syncwarp();
    if (c.lane == 0) {
        const int old = S.buf[t];
        S.buf[t] = S.scratch; S.scratch = old;
        S.root[t] = 0; S.alloc[t] = count;
        atomicAdd(&c.E.ctr->compactions, 1ull);
    }
    __syncwarp();
}

// Root of the coming search: first search of a colour builds a fresh root
// (training_pipeline.py:353,371-373), later ones re-root the colour's own tree through the
// states played since its last move (MCTS.new_root_node, MCTS.py:250-295).
__device__ void setup_root(const WarpCtx &c) {
    const EngineDev &E = c.E;
    Slot &S = c.S;
    const int L = S.hist_len;
    const ck_pos cur = c.hist[L - 1];
    const int t = meta_player(cur.meta);
    const int parent_player = L >= 2 ? meta_player(c.hist[L - 2].meta) : 1 - t;   // MCTS.py:167-173
    bool fresh = !S.exists[t];
    if (!fresh) {
        const uint4 *pos = c.pos_of(t), *stat = c.stat_of(t);
        int counter = 1;
        for (int idx = L - 3; idx >= 0 && meta_player(c.hist[L - 2].meta) == meta_player(c.hist[idx].meta); --idx) ++counter;
        int nr = S.best[t];
        for (int i = L - counter; i < L; ++i) {
            const ck_pos want = c.hist[i];
            const uint4 nst = stat[nr];
            const uint32_t link = nst.w;
            const int b = E.uct ? uct_expanded(nst) : link_nchild(link), fc = (int)(link & kFcMask);
            int found = kNoIdx;
            for (int j = c.lane; j < b; j += 32)
                if (found == kNoIdx && same_state(to_pos(pos[fc + j]), want)) found = j;
#pragma unroll
            for (int o = 16; o; o >>= 1) found = min(found, __shfl_xor_sync(CK_FULL, found, o));
            if (found != kNoIdx) nr = fc + found;
        }
        if (same_state(to_pos(pos[nr]), cur)) {
            __syncwarp();
            if (c.lane == 0) {
                S.root[t] = nr;
                uint4 *wstat = c.stat_of(t);
                wstat[nr].w = (wstat[nr].w & ~(1u << 30)) | ((uint32_t)parent_player << 30);
            }
            __syncwarp();
            if (E.cap - S.alloc[t] < E.compact_need || E.cfg.compact_always) compact_tree(c, t, nr);
        } else {
            // the reference raises here (MCTS.py:292); the dead code after the raise shows the
            // intent -- a fresh root.  Counted, never silent.
            fresh = true;
            __syncwarp();
            if (c.lane == 0) S.misses += 1;
        }
    }
    if (fresh) {
        __syncwarp();
        if (c.lane == 0) {
            uint4 *pos = c.pos_of(t), *stat = c.stat_of(t);
            pos[0] = from_pos(cur);
            stat[0] = make_uint4(0u, __float_as_uint(0.f), __float_as_uint(0.f), (uint32_t)parent_player << 30);
            S.root[t] = 0; S.alloc[t] = 1; S.exists[t] = 1;
            atomicAdd(&E.ctr->nodes, 1ull);
        }
    }
    __syncwarp();
    if (c.lane == 0) { S.cur = t; S.phase = PH_SEARCH; S.sims_done = 0; S.search_id += 1; }
    __syncwarp();
}

__device__ void init_game(const WarpCtx &c, int local_game) {
    Slot &S = c.S;
    __syncwarp();
    if (c.lane == 0) {
        S.game = local_game;
        S.hist_len = 1; S.move_count = 0; S.phase = PH_NEED_ROOT; S.cur = 0; S.sims_done = 0;
        S.pend_leaf = -1;
        S.buf[0] = 0; S.buf[1] = 1; S.scratch = 2;
        S.exists[0] = S.exists[1] = 0;
        S.nrec = 0; S.misses = 0; S.search_id = 0; S.g_sims0 = S.tot_sims; S.g_evals0 = S.tot_evals;
        // warm start (benchmarks): the first game of a slot plays a number of opening plies at a small budget, so that the
        // slots reach full-budget play at different stages of their games instead of in lock step.  The number is a
        // function of the GAME (which slot claims which game depends on warp scheduling; records must not)
        S.stagger_until = (c.E.cfg.stagger_plies > 0 && S.tot_sims == 0)
                              ? (int32_t)(mix32((uint32_t)local_game * 0x9E3779B1u ^ (uint32_t)c.E.cfg.seed) % (uint32_t)c.E.cfg.stagger_plies) : 0;
        if (!c.E.cfg.reference_tau_quirk || S.tau < -1e300) S.tau = c.E.cfg.tau;
        c.hist[0] = start_position();
    }
    __syncwarp();
}

// take the next staged game or go idle
__device__ bool refill(const WarpCtx &c) {
    int g = -1;
    if (c.lane == 0 && *(volatile int32_t *)&c.E.ctr->next_game < c.E.n_games) {
        g = atomicAdd(&c.E.ctr->next_game, 1);
        if (g >= c.E.n_games) g = -1;
    }
    g = __shfl_sync(CK_FULL, g, 0);
    __syncwarp();
    if (g < 0) { if (c.lane == 0) c.S.game = -1; __syncwarp(); return false; }
    init_game(c, g);
    return true;
}

// K6: MCTS.best_child (MCTS.py:226-248) + the game loop body after a search
// (training_pipeline.py:362-411): record, step, TERMINATE_CNT adjudication, terminal record,
// rewards.  Returns false when the slot went idle.
__device__ bool play_move(const WarpCtx &c) {
    const EngineDev &E = c.E;
    Slot &S = c.S;
    const int t = S.cur;
    uint4 *pos = c.pos_of(t), *stat = c.stat_of(t);
    const int root = S.root[t];
    const uint4 rst = stat[root];
    const ck_pos rpos = to_pos(pos[root]);
    const int b = E.uct ? uct_expanded(rst) : link_nchild(rst.w), fc = (int)(rst.w & kFcMask);
    const int parent_player = link_pp(rst.w);
    uint32_t n0 = 0, n1 = 0;
    ck_pos k0 = rpos, k1 = rpos;
    uint32_t l0 = 0, l1 = 0;
    if (c.lane < b) { const uint4 s = stat[fc + c.lane]; n0 = s.x; l0 = s.w; k0 = to_pos(pos[fc + c.lane]); }
    if (c.lane + 32 < b) { const uint4 s = stat[fc + 32 + c.lane]; n1 = s.x; l1 = s.w; k1 = to_pos(pos[fc + 32 + c.lane]); }
    int best;
    if (!E.cfg.training || S.tau <= 0.0) {              // argmax visits, first maximum (:236-238)
        long long key = -1;
        if (c.lane < b) key = ((long long)n0 << 8) | (long long)(255 - c.lane);
        if (c.lane + 32 < b) { const long long k2 = ((long long)n1 << 8) | (long long)(255 - (c.lane + 32)); if (k2 > key) key = k2; }
#pragma unroll
        for (int o = 16; o; o >>= 1) { const long long ok = __shfl_xor_sync(CK_FULL, key, o); if (ok > key) key = ok; }
        best = 255 - (int)(key & 255);
    } else {                                             // sample ~ n^(1/tau) (:239-246)
        const double inv = 1.0 / S.tau;
        const double e0 = c.lane < b ? pow((double)n0, inv) : 0.0;
        const double e1 = c.lane + 32 < b ? pow((double)n1, inv) : 0.0;
        double s0 = e0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const double v = shfl_up_d(s0, o); if (c.lane >= o) s0 += v; }
        const double tot0 = shfl_d(s0, 31);
        double s1 = e1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const double v = shfl_up_d(s1, o); if (c.lane >= o) s1 += v; }
        const double total = tot0 + shfl_d(s1, 31);
        s1 += tot0;
        uint32_t r[4];
        Philox(game_key(E, S.game))((uint32_t)S.search_id, 0xFFFFFFFFu, 0u, 0x54415521u, r);
        const double u = u01(r[0], r[1]) * total;
        int pick = kNoIdx;
        if (c.lane < b && u <= s0) pick = c.lane;
        if (pick == kNoIdx && c.lane + 32 < b && u <= s1) pick = c.lane + 32;
#pragma unroll
        for (int o = 16; o; o >>= 1) pick = min(pick, __shfl_xor_sync(CK_FULL, pick, o));
        best = pick == kNoIdx ? b - 1 : pick;
        __syncwarp();
        if (c.lane == 0 && S.move_count > E.cfg.tau_decay_delay) {
            double tau = S.tau - E.cfg.tau_decay;
            if (fabs(tau) <= 1e-8) tau = 0.0;           // np.isclose(tau, 0)
            S.tau = tau;
        }
    }
    const ck_pos bpos = best < 32 ? to_pos(shfl4(from_pos(k0), best)) : to_pos(shfl4(from_pos(k1), best - 32));
    const uint32_t blink = best < 32 ? __shfl_sync(CK_FULL, l0, best) : __shfl_sync(CK_FULL, l1, best - 32);
    const int bstatus = link_status(blink);
    const int gg = global_game(E, S.game);
    // record [root.state, visit planes, q] (:364-369)
    ck_record *rec = nullptr;
    if (E.cfg.keep_records && S.nrec < E.max_rec) rec = E.rec + (int64_t)S.game * E.max_rec + S.nrec;
    if (rec) {
        if (c.lane < b) { rec->action[c.lane] = (uint16_t)meta_action(k0.meta); rec->visits[c.lane] = n0; }
        if (c.lane + 32 < b) { rec->action[c.lane + 32] = (uint16_t)meta_action(k1.meta); rec->visits[c.lane + 32] = n1; }
        if (c.lane == 0) {
            int p5;
            const int cnt = gen_moves(rpos, NullSink{}, rec->mask);
            outcome_of(rpos, cnt > 0, &p5);
            rec->pos = rpos; rec->plane5 = p5; rec->n_children = b;
            const float q = rst.x ? __fdiv_rn(__uint_as_float(rst.y), (float)rst.x) : 0.f;
            rec->q = parent_player != meta_player(rpos.meta) ? -q : q;
            rec->z = 0; rec->root_n = rst.x; rec->root_w = __uint_as_float(rst.y);
            rec->chosen = meta_action(bpos.meta); rec->game = gg; rec->ply = S.nrec;
        }
    }
    bool done = false, terminated = false;
    int outcome = CK_ONGOING;
    __syncwarp();
    if (c.lane == 0) {
        if (rec) S.nrec += 1;
        S.best[t] = fc + best;
        if (S.hist_len <= E.max_plies) { c.hist[S.hist_len] = bpos; S.hist_len += 1; }
        else dev_error(E, CK_ERR_STATE);
        S.move_count += 1;
        atomicAdd(&E.ctr->moves, 1ull);
    }
    __syncwarp();
    if (bstatus != CK_ONGOING) { done = true; outcome = bstatus; }
    if (!done && E.cfg.terminate_cnt > 0 && S.move_count >= E.cfg.terminate_cnt) {   // (:387-405)
        done = true; terminated = true;
        const int p1 = popc32(bpos.p1), p2 = popc32(bpos.p2);
        const int q1 = popc32(bpos.p1 & bpos.k), q2 = popc32(bpos.p2 & bpos.k);
        outcome = p1 > p2 ? CK_P1_WINS : p1 < p2 ? CK_P2_WINS : q1 > q2 ? CK_P1_WINS : q1 < q2 ? CK_P2_WINS : CK_DRAW;
    }
    if (!done) {
        if (c.lane == 0) S.phase = PH_NEED_ROOT;
        __syncwarp();
        return true;
    }
    __syncwarp();
    if (!terminated && E.cfg.keep_records && S.nrec < E.max_rec && c.lane == 0) {    // terminal record (:406-409)
        ck_record *tr = E.rec + (int64_t)S.game * E.max_rec + S.nrec;
        int p5;
        const int cnt = gen_moves(bpos, NullSink{}, tr->mask);
        outcome_of(bpos, cnt > 0, &p5);
        tr->pos = bpos; tr->plane5 = p5; tr->n_children = 0;
        tr->q = outcome == CK_DRAW ? 0.f : -1.f;
        tr->z = 0; tr->root_n = 0; tr->root_w = 0.f; tr->chosen = -1; tr->game = gg; tr->ply = S.nrec;
        S.nrec += 1;
    }
    __syncwarp();
    if (E.cfg.keep_records) {                            // _add_rewards (:439-455)
        ck_record *base = E.rec + (int64_t)S.game * E.max_rec;
        for (int i = c.lane; i < S.nrec; i += 32) {
            const int pl = meta_player(base[i].pos.meta);
            base[i].z = outcome == CK_P1_WINS ? (pl == 0 ? 1 : -1) : outcome == CK_P2_WINS ? (pl == 1 ? 1 : -1) : 0;
        }
    }
    if (c.lane == 0) {
        ck_game_result &r = E.results[S.game];
        r.game = gg; r.outcome = outcome; r.move_count = S.move_count; r.terminated = terminated ? 1 : 0;
        r.n_records = S.nrec; r.reroot_misses = S.misses;
        r.p1_net = E.cfg.arena ? net_of(E, S.game, 0) : 0; r.reserved = 0;
        r.sims = S.tot_sims - S.g_sims0; r.nn_evals = S.tot_evals - S.g_evals0;
        __threadfence();
        atomicAdd(&E.ctr->games_finished, 1ull);
    }
    __syncwarp();
    return refill(c);
}

// ---- the per-round tree kernel ------------------------------------------------------------
// CTAs per SM the register allocation is held to (95 registers, no spills).  Raising it to 6 or 7 fits the
// whole cfg2 grid (1024 CTAs) in one wave at the price of spills; measured neutral to slightly negative in
// steady state (3.74 / 3.73 / 3.72 / 3.70 M sims/s at 4 / 5 / 6 / 7), so the spill-free setting stays.
#ifndef CK_TREE_OCC
#define CK_TREE_OCC 5
#endif
#ifndef CK_OVERLAP_DEFAULT
#define CK_OVERLAP_DEFAULT 0
#endif
// kWarps = 4: the stand-alone launch.  kWarps = 1: the overlapped launch -- one-warp blocks (3 K registers) fit next to a
// resident tower CTA (576 threads x 96 registers leave 10 K of the SM's 64 K), three per SM.
template <bool kUct, int kWarps>
__global__ void __launch_bounds__(kWarps * 32, kWarps == 4 ? CK_TREE_OCC : 20)
tree_step_kernel(const EngineDev E) {
    __shared__ Slot s_slot[kWarps];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = E.slot0 + blockIdx.x * kWarps + w;
    if (slot >= E.slot0 + E.slot_n) return;
    Slot &S = s_slot[w];
    {
        const int *src = reinterpret_cast<const int *>(E.slots + slot);
        int *dst = reinterpret_cast<int *>(&S);
        for (int i = lane; i < (int)(sizeof(Slot) / 4); i += 32) dst[i] = src[i];
    }
    __syncwarp();
    WarpCtx c{E, S, E.path + (int64_t)slot * kMaxDepth, E.hist + (int64_t)slot * (E.max_plies + 1), slot, lane};
    bool live = true;
    if (E.ctr->error != 0) live = false;
    if (live && S.game < 0) live = S.manual ? false : refill(c);
    if (live && S.phase == PH_HALT) live = false;
    if (live && S.pend_leaf >= 0) { if (kUct) finish_playout(c); else expand_pending(c); }
    int term_iters = 0, chain = 0;
    const int max_term = E.max_term, max_chain = E.max_chain;
    while (live && S.phase != PH_HALT) {
        if (S.phase == PH_NEED_ROOT) setup_root(c);
        const int target = S.manual ? S.manual_target : (S.move_count < S.stagger_until ? E.cfg.stagger_budget : E.cfg.budget);
        if (S.sims_done >= target) {                     // MCTS.computational_budget (MCTS.py:196-198)
            if (S.manual) { __syncwarp(); if (lane == 0) S.phase = PH_HALT; __syncwarp(); break; }
            if (!play_move(c)) { live = false; break; }
            continue;
        }
        int depth = 0;
        const int leaf = kUct ? select_leaf_uct(c, &depth) : select_leaf(c, &depth);
        if (leaf >= 0) {
            if (kUct) { stage_playout(c, leaf, depth); break; }
            if (E.cache != nullptr && expand_cached(c, leaf, depth)) {       // evaluated before: no network call
                if (S.phase == PH_HALT || ++chain >= max_chain) break;
                continue;
            }
            stage_leaf(c, leaf, depth);
            break;
        }
        __syncwarp();
        if (leaf == -2) { if (lane == 0) S.phase = PH_HALT; __syncwarp(); break; }
        if (lane == 0) { S.tot_sims += 1; S.sims_done += 1; }
        __syncwarp();
        if (++term_iters >= max_term || ++chain >= max_chain) break;
    }
    if (lane == 0 && S.game >= 0 && S.phase != PH_HALT) atomicAdd(&E.ctr->active[E.grp], 1);
    __syncwarp();
    {
        int *dst = reinterpret_cast<int *>(E.slots + slot);
        const int *src = reinterpret_cast<const int *>(&S);
        for (int i = lane; i < (int)(sizeof(Slot) / 4); i += 32) dst[i] = src[i];
    }
}

// single-search API: renumber slot 0's current tree around its root (ck_tree_reroot when the pool runs short)
__global__ void __launch_bounds__(32) manual_compact_kernel(const EngineDev E) {
    __shared__ Slot S;
    const int lane = threadIdx.x;
    {
        const int *src = reinterpret_cast<const int *>(E.slots);
        int *dst = reinterpret_cast<int *>(&S);
        for (int i = lane; i < (int)(sizeof(Slot) / 4); i += 32) dst[i] = src[i];
    }
    __syncwarp();
    WarpCtx c{E, S, E.path, E.hist, 0, lane};
    compact_tree(c, S.cur, S.root[S.cur]);
    {
        int *dst = reinterpret_cast<int *>(E.slots);
        const int *src = reinterpret_cast<const int *>(&S);
        for (int i = lane; i < (int)(sizeof(Slot) / 4); i += 32) dst[i] = src[i];
    }
}

// Start of a round: shape the coming evaluator batch and zero the round's counters (one tiny launch instead of a
// memset).  The tower evaluates 4 positions per CTA iteration on every SM, so a batch of k full waves plus a few
// positions costs k + 1 iterations; at cfg2 with the evaluation cache ~3700 leaves ask for a row per round
// (6.24 waves -> 7 iterations, the tensor pipe idle 9 % of the launch).  The cap is the request count of the previous
// round rounded down to whole waves; the leaves beyond it wait one round.  Measured: see DESIGN.md.
__global__ void round_begin_kernel(const EngineDev E) {
    Counters *c = E.ctr;
    if (threadIdx.x == 0) {
        const int nb = E.cfg.arena ? 2 : 1;                  // arena: both networks' batches; otherwise this launch's group
        for (int i = 0; i < nb; ++i) {
            const int k = E.cfg.arena ? i : E.grp;
            int cap = 0x7FFFFFFF;
            if (E.wave > 0) {
                // a leaf that waits costs its slot a whole evaluate-then-chain cycle (~2.4 simulations at cfg2), a wave iteration
                // saved is worth ~590 of them: cutting the batch back pays while the excess is below ~0.5 wave (measured
                // break-even, profiles/r2e_steady_sweep.jsonl); applied below 0.4 wave (the long-run cfg2 state asks for
                // 6.25-6.33 waves every round, profiles/r2y_batch_hist.jsonl; 0.3 -> 0.4: +4 %, profiles/r2z_slack.jsonl)
                const int asked = c->batch_count[k] + c->deferred[k];       // previous round of this group
                const int full = asked / E.wave * E.wave;
                if (full >= E.wave && (asked - full) * 10 < E.wave_slack10 * E.wave) cap = full;
            }
            c->leaf_cap[k] = cap;
            c->active[k] = 0; c->batch_count[k] = 0; c->deferred[k] = 0;
        }
    }
}

// simulation / evaluation totals live per slot (no hot-path atomics); summed on demand
__global__ void sum_slots_kernel(const EngineDev E, unsigned long long *out /* sims, evals, cache hits */) {
    unsigned long long s = 0, e = 0, h = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < E.n_slots; i += gridDim.x * blockDim.x) {
        s += E.slots[i].tot_sims; e += E.slots[i].tot_evals; h += E.slots[i].tot_hits;
    }
    atomicAdd(out, s); atomicAdd(out + 1, e); atomicAdd(out + 2, h);
}

// ---- stub evaluators (deterministic parity tests; twins of the oracle's cko_eval_*) ----------
__global__ void __launch_bounds__(128)
stub_eval_kernel(const ck_leaf *__restrict__ leaves, const int32_t *__restrict__ n_dev, int kind,
                 float *__restrict__ policy, float *__restrict__ value) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= *n_dev) return;
    const ck_leaf L = leaves[row];
    float *dst = policy + (int64_t)row * CK_POLICY_SIZE;
    if (kind == CK_EVAL_HASH || kind == CK_EVAL_HASH_SALTED) {
        const uint32_t p5 = (L.info >> 8) & 0xFFu;
        const uint32_t salt = kind == CK_EVAL_HASH_SALTED ? (L.info >> 16) : 0u;
        uint32_t h = mix32(L.p1 ^ 0x9e3779b9u ^ (salt * 0x9E3779B1u));
        h = mix32(h ^ L.p2);
        h = mix32(h ^ L.k);
        h = mix32(h ^ (L.info & 1u) ^ (p5 << 8));
        for (int i = lane; i < CK_POLICY_SIZE; i += 32) {
            const uint32_t g = mix32(h + (uint32_t)i * 0x85ebca6bu);
            const float v = __fadd_rn((float)((g >> 8) & 0xFFFFu), 1.0f);
            dst[i] = __fmul_rn(__fmul_rn(__fmul_rn(v, v), v), 0x1p-58f);
        }
        if (lane == 0) {
            const uint32_t g = mix32(h ^ 0xdeadbeefu);
            value[row] = __fsub_rn(__fmul_rn((float)(g & 0xFFFFFFu), 0x1p-23f), 1.0f);
        }
    } else {
        for (int i = lane; i < CK_POLICY_SIZE; i += 32) dst[i] = 1.0f / 512.0f;
        if (lane == 0) {
            float v = 0.f;
            if (kind == CK_EVAL_UNIFORM_MATERIAL) {
                const int a = __popc(L.p1) + __popc(L.p1 & L.k), b = __popc(L.p2) + __popc(L.p2 & L.k);
                v = __fdiv_rn((float)((L.info & 1u) == 0 ? a - b : b - a), 32.0f);
            }
            value[row] = v;
        }
    }
}

// Playouts of the staged positions (MCTS.default_policy without a net, MCTS.py:132-143): thread per game,
// uniformly random legal successors until determine_outcome ends the game; value[row] = CK_* outcome.
// CK_EVAL_ROLLOUT_HASH replaces the random index by a hash of the position the move is chosen from
// (twin of the oracle's cko_hash_choice) so that whole searches can be compared bit for bit.
__global__ void __launch_bounds__(128)
playout_eval_kernel(const ck_leaf *__restrict__ leaves, const int32_t *__restrict__ n_dev, int kind, uint64_t seed,
                    uint32_t round, float *__restrict__ value, int lanes_log2) {
    // playouts diverge from the first ply on, so a warp costs the sum of its lanes' paths: small batches use
    // only the first 2^lanes_log2 lanes of every warp and spread over more warps (and SMs) instead
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = tid & 31;
    if (lane >= (1 << lanes_log2)) return;
    const int row = ((tid >> 5) << lanes_log2) + lane;
    if (row >= *n_dev) return;
    const ck_leaf L = leaves[row];
    ck_pos cur;
    cur.p1 = L.p1; cur.p2 = L.p2; cur.k = L.k; cur.meta = L.mask[0];
    int st;
    if (kind == CK_EVAL_ROLLOUT_HASH) {
        st = play_out(cur, 0, HashChoice{L.mask[1]}, nullptr);               // salted with the game id
    } else {
        st = play_out(cur, 0, PhiloxChoice{Philox(mix64(seed ^ mix64(((uint64_t)L.mask[1] << 32) | round))), 0x504C4159u}, nullptr);
    }
    value[row] = (float)st;
}

// ---- packed records (iteration-end gather) ------------------------------------------------------
// A ck_record is 372 bytes, most of it empty child slots.  Packed form: a 40-byte header per record plus one word per
// child (action << 23 | visits) -- about 75 bytes per record -- written game by game (local game order, plies in
// order) straight from the engine's record store, so that the gather moves a fifth of the bytes and needs no host
// bounce.  The legal-action planes of a searched position are exactly its children's actions and are rebuilt by the
// receiver; a terminal record whose position still has legal moves (a draw by the 80-ply rule) appends its 8 mask words.
__device__ __forceinline__ int packed_words(const ck_record &r, bool *mask_follows) {
    const bool mf = r.n_children == 0 && (r.mask[0] | r.mask[1] | r.mask[2] | r.mask[3] | r.mask[4] | r.mask[5] | r.mask[6] | r.mask[7]) != 0u;
    *mask_follows = mf;
    return r.n_children + (mf ? 8 : 0);
}

// thread per game: records and child words of a finished game (0 / 0 otherwise)
__global__ void pack_count_kernel(const EngineDev E, int64_t *__restrict__ counts) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= E.n_games) return;
    const ck_game_result r = E.results[g];
    int64_t nrec = 0, nw = 0;
    if (r.game >= 0 && r.outcome >= 0) {
        nrec = r.n_records;
        const ck_record *rec = E.rec + (int64_t)g * E.max_rec;
        for (int i = 0; i < r.n_records; ++i) { bool mf; nw += packed_words(rec[i], &mf); }
    }
    counts[2 * g] = nrec; counts[2 * g + 1] = nw;
}

// warp per game; offs[2g], offs[2g+1] = first header / first child word of game g
__global__ void __launch_bounds__(128)
pack_write_kernel(const EngineDev E, const int64_t *__restrict__ counts, const int64_t *__restrict__ offs,
                  ck_record_hdr *__restrict__ hdr, uint32_t *__restrict__ words) {
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (g >= E.n_games) return;
    const int nrec = (int)counts[2 * g];
    if (nrec == 0) return;
    const ck_record *rec = E.rec + (int64_t)g * E.max_rec;
    ck_record_hdr *h = hdr + offs[2 * g];
    int64_t wbase = offs[2 * g + 1];
    for (int r0 = 0; r0 < nrec; r0 += 32) {
        const int i = r0 + lane;
        bool mf = false;
        const int nw = i < nrec ? packed_words(rec[i], &mf) : 0;
        int incl = nw;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(CK_FULL, incl, o); if (lane >= o) incl += v; }
        const int total = __shfl_sync(CK_FULL, incl, 31);
        if (i < nrec) {
            const ck_record &R = rec[i];
            ck_record_hdr o;
            o.pos = R.pos; o.q = R.q; o.root_w = R.root_w; o.root_n = R.root_n; o.game = R.game;
            o.ply = (uint16_t)R.ply; o.chosen = (int16_t)R.chosen; o.n_children = (uint8_t)R.n_children;
            o.plane5 = (uint8_t)R.plane5; o.z = (int8_t)R.z; o.flags = mf ? 1u : 0u;
            h[i] = o;
            uint32_t *w = words + wbase + (incl - nw);
            for (int j = 0; j < R.n_children; ++j) w[j] = ((uint32_t)R.action[j] << 23) | (R.visits[j] & 0x7FFFFFu);
            if (mf) for (int j = 0; j < 8; ++j) w[j] = R.mask[j];
        }
        wbase += total;
    }
}

}  // namespace ck

// =================================================================================================
// host side
// =================================================================================================
using namespace ck;

struct ck_engine {
    EngineDev dev;           // device pointers + config, passed by value to the kernels
    ck_net *net[2] = {nullptr, nullptr};
    uint64_t net_gen[2] = {0, 0};       // weights generation the evaluation cache was filled under
    cudaStream_t stream = nullptr;
    cudaStream_t stream_b = nullptr;    // arena: the second network evaluates next to the first one; overlapped self-play: the tree stream
    cudaEvent_t ev_tree[2] = {nullptr, nullptr}, ev_eval[2] = {nullptr, nullptr};   // overlapped self-play: per slot group
    cudaEvent_t ev_fork_run = nullptr;
    bool eval_pending[2] = {false, false};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_fork = nullptr, ev_join = nullptr;
    std::vector<cudaEvent_t> prof_ev;   // 3 per round in profile mode: before eval, after tower, after eval
    std::vector<char> fetched;          // per local game: records already handed out by ck_records_fetch_new
    Counters *h_ctr = nullptr;        // pinned
    unsigned long long *d_tot = nullptr;   // ck_engine_run: simulation / evaluation totals before and after
    int64_t n_games = 0;
    size_t rec_cap = 0, res_cap = 0;
    bool profile = false;
    bool begun = false;
    uint64_t total_steps = 0;
    int64_t tree_epoch = 0;             // single-search API: times the node ids were renumbered
};

static void engine_free(ck_engine *e) {
    if (!e) return;
    EngineDev &d = e->dev;
    cudaFree(d.pos); cudaFree(d.stat); cudaFree(d.hist); cudaFree(d.path); cudaFree(d.slots); cudaFree(d.ctr);
    for (int k = 0; k < 2; ++k) { cudaFree(d.leaves[k]); cudaFree(d.policy[k]); cudaFree(d.value[k]); }
    cudaFree(d.rec); cudaFree(d.results); cudaFree((void *)d.pow_half); cudaFree((void *)d.log_tab); cudaFree(e->d_tot);
    cudaFree(d.cache);
    if (e->h_ctr) cudaFreeHost(e->h_ctr);
    if (e->stream_b) cudaStreamDestroy(e->stream_b);
    for (int k = 0; k < 2; ++k) { if (e->ev_tree[k]) cudaEventDestroy(e->ev_tree[k]); if (e->ev_eval[k]) cudaEventDestroy(e->ev_eval[k]); }
    if (e->ev_fork_run) cudaEventDestroy(e->ev_fork_run);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    for (cudaEvent_t ev : e->prof_ev) cudaEventDestroy(ev);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

#define CK_E(expr)                                                                           \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            ck::fail(CK_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));       \
            engine_free(e);                                                                  \
            return nullptr;                                                                  \
        }                                                                                    \
    } while (0)

extern "C" {

ck_engine *ck_engine_create(const ck_engine_cfg *cfg) {
    if (!cfg || cfg->n_slots < 1 || cfg->budget < 1) { fail(CK_ERR_ARG, "ck_engine_create: n_slots and budget must be >= 1"); return nullptr; }
    if (cfg->stagger_plies > 0 && cfg->stagger_budget < 1) { fail(CK_ERR_ARG, "ck_engine_create: stagger_plies needs stagger_budget >= 1"); return nullptr; }
    const bool uct = cfg->evaluator == CK_EVAL_ROLLOUT || cfg->evaluator == CK_EVAL_ROLLOUT_HASH;
    if (cfg->evaluator < CK_EVAL_NET || cfg->evaluator > CK_EVAL_ROLLOUT_HASH) { fail(CK_ERR_ARG, "ck_engine_create: unknown evaluator"); return nullptr; }
    if (cfg->arena && (uct || cfg->evaluator_p2 == CK_EVAL_ROLLOUT || cfg->evaluator_p2 == CK_EVAL_ROLLOUT_HASH)) {
        fail(CK_ERR_ARG, "ck_engine_create: the playout evaluators (NEURAL_NET=False) are a self-play mode, not an arena evaluator");
        return nullptr;
    }
    DeviceGuard g(cfg->device);
    if (!g.ok) { fail(CK_ERR_CUDA, "ck_engine_create: cannot select CUDA device " + std::to_string(cfg->device)); return nullptr; }
    ck_engine *e = new ck_engine();
    EngineDev &d = e->dev;
    memset(&d, 0, sizeof(d));
    d.cfg = *cfg;
    d.n_slots = cfg->n_slots;
    d.max_plies = cfg->max_plies > 0 ? cfg->max_plies : 2048;
    int cap = cfg->pool_cap;
    if (cap <= 0) {
        // default: as many nodes per tree buffer as 40 % of the free HBM allows (up to 128 Ki), so that a
        // tree is compacted once or twice per game rather than every few moves; never below what one
        // search can allocate
        cap = 32768;
        while (cap < 3 * cfg->budget * 24) cap *= 2;
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            while (cap < 131072 && (size_t)cfg->n_slots * 3 * (size_t)(2 * cap) * 2 * sizeof(uint4) <= free_b / 5 * 2) cap *= 2;
        }
    }
    if (cap > (int)kFcMask) cap = (int)kFcMask;
    d.cap = cap;
    int need = cfg->budget * CK_MAX_CHILDREN;            // room for a whole search in the worst case
    if (need > cap / 2) need = cap / 2;
    d.compact_need = need;
    // Simulations that end in a terminal child need no evaluation and are finished inside the round; a
    // slot near the end of its game would chain dozens of them and become the tail of the launch (at 64
    // per round the steady-state tree_step took 0.35 ms instead of 0.1 ms: scripts/long_run.py), so the
    // default lets a slot finish four and carries the rest into the next rounds.
    d.max_term = cfg->max_terminal_sims_per_step > 0 ? cfg->max_terminal_sims_per_step : 4;
    d.max_chain = cfg->max_chain_per_step > 0 ? cfg->max_chain_per_step : 6;      // 4 / 5 / 6 / 7 / 8 / 10 measured, profiles/r3c, r3d
    if (d.max_chain < d.max_term) d.max_chain = d.max_term;
    // evaluation cache: per slot a power of two of 128-byte entries (default 16384 = 2 MB per slot, 8.6 GB at cfg2: hit
    // rate 41.8 % / 44.1 % at 4096 / 16384 entries in the warm-started bench, +4 % simulations/s); the playout modes
    // have no network, the uniform stubs nothing worth caching (but they exercise the path in the tests)
    if (!uct && cfg->eval_cache_entries >= 0) {
        int want = cfg->eval_cache_entries > 0 ? cfg->eval_cache_entries : 16384;
        int ent = 16;
        while (ent < want && ent < (1 << 20)) ent *= 2;
        if (cfg->eval_cache_entries == 0) {
            // the default yields to the trees: at most a quarter of the free HBM (16384 entries x 65536 slots would be 137 GB)
            size_t free_b = 0, total_b = 0;
            if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
                while (ent > 256 && (size_t)cfg->n_slots * ent * kCacheU4 * sizeof(uint4) > free_b / 4) ent /= 2;
        }
        d.cache_entries = ent;
        d.cache_game_tag = (cfg->evaluator == CK_EVAL_HASH_SALTED || cfg->evaluator_p2 == CK_EVAL_HASH_SALTED) ? 1 : 0;
    }
    d.one_minus_eps = (float)(1.0 - cfg->epsilon);       // (1 - eps) * float32 array stays float32 (numpy >= 2)
    if (cfg->keep_records) d.max_rec = (cfg->terminate_cnt > 0 ? cfg->terminate_cnt : d.max_plies) + 1;
    // batch shaping only where the tower runs whole waves: self-play with the network, two tiles per CTA
    {
        // CK_BATCH_WAVES=0 switches it off; CK_BATCH_WAVES=n (> 1) forces a wave of n rows for every evaluator (tests)
        static const int shape_env = getenv("CK_BATCH_WAVES") ? atoi(getenv("CK_BATCH_WAVES")) : -1;
        const int wave = 4 * num_sms(cfg->device);
        static const int slack_env = getenv("CK_BATCH_SLACK10") ? atoi(getenv("CK_BATCH_SLACK10")) : -1;   // measurement knob
        d.wave_slack10 = shape_env > 1 ? 10 : (slack_env >= 0 ? slack_env : 4);     // forced mode always cuts back (every round defers leaves)
        if (shape_env > 1) d.wave = cfg->arena ? 0 : shape_env;
        else d.wave = (shape_env != 0 && cfg->evaluator == CK_EVAL_NET && !cfg->arena && cfg->n_slots >= 2 * wave) ? wave : 0;
    }
    // Overlapped self-play: two slot groups, the tree kernel of one runs next to the tower of the other (CK_OVERLAP=0|1
    // overrides).  Needs two full-size batches of tower waves to pay: see DESIGN.md section 5.
    {
        static const int ov_env = getenv("CK_OVERLAP") ? atoi(getenv("CK_OVERLAP")) : -1;
        const int wave = 4 * num_sms(cfg->device);
        const bool can = cfg->evaluator == CK_EVAL_NET && !cfg->arena && cfg->n_slots >= 4 * wave;
        d.groups = (can && (ov_env > 0 || (ov_env < 0 && CK_OVERLAP_DEFAULT))) ? 2 : 1;
    }
    d.grp = 0; d.slot0 = 0; d.slot_n = cfg->n_slots;
    CK_E(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CK_E(cudaEventCreate(&e->ev0)); CK_E(cudaEventCreate(&e->ev1));
    if (d.groups == 2) {
        CK_E(cudaStreamCreateWithFlags(&e->stream_b, cudaStreamNonBlocking));
        CK_E(cudaEventCreateWithFlags(&e->ev_fork_run, cudaEventDisableTiming));
        for (int k = 0; k < 2; ++k) {
            CK_E(cudaEventCreateWithFlags(&e->ev_tree[k], cudaEventDisableTiming));
            CK_E(cudaEventCreateWithFlags(&e->ev_eval[k], cudaEventDisableTiming));
        }
    }
    if (cfg->arena) {
        CK_E(cudaStreamCreateWithFlags(&e->stream_b, cudaStreamNonBlocking));
        CK_E(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
        CK_E(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
    }
    const size_t nodes = (size_t)d.n_slots * 3 * d.cap;
    CK_E(cudaMalloc(&d.pos, nodes * sizeof(uint4)));
    CK_E(cudaMalloc(&d.stat, nodes * sizeof(uint4)));
    CK_E(cudaMalloc(&d.hist, (size_t)d.n_slots * (d.max_plies + 1) * sizeof(ck_pos)));
    CK_E(cudaMalloc(&d.path, (size_t)d.n_slots * kMaxDepth * sizeof(uint32_t)));
    CK_E(cudaMalloc(&d.slots, (size_t)d.n_slots * sizeof(Slot)));
    CK_E(cudaMalloc(&d.ctr, sizeof(Counters)));
    CK_E(cudaMemset(d.ctr, 0, sizeof(Counters)));
    for (int k = 0; k < ((cfg->arena || d.groups == 2) ? 2 : 1); ++k) {
        CK_E(cudaMalloc(&d.leaves[k], (size_t)d.n_slots * sizeof(ck_leaf)));
        CK_E(cudaMalloc(&d.policy[k], (size_t)d.n_slots * CK_POLICY_SIZE * sizeof(float)));
        CK_E(cudaMalloc(&d.value[k], (size_t)d.n_slots * sizeof(float)));
    }
    if (d.cache_entries > 0) {
        const size_t bytes = (size_t)d.n_slots * d.cache_entries * kCacheU4 * sizeof(uint4);
        CK_E(cudaMalloc(&d.cache, bytes));
        CK_E(cudaMemset(d.cache, 0, bytes));
    }
    CK_E(cudaMallocHost(&e->h_ctr, sizeof(Counters) + 2 * sizeof(int32_t)));   // + the two networks' range flags
    CK_E(cudaMalloc(&e->d_tot, 6 * sizeof(unsigned long long)));
    {
        // node.n ** 0.5 is libm pow in the reference (MCTS.py:110) and differs from sqrt for
        // some integers; the table is built with the host's libm so the two agree.
        std::vector<double> tab(kPowTable);
        for (int i = 0; i < kPowTable; ++i) tab[i] = pow((double)i, 0.5);
        double *p = nullptr;
        CK_E(cudaMalloc(&p, kPowTable * sizeof(double)));
        d.pow_half = p;
        CK_E(cudaMemcpy(p, tab.data(), kPowTable * sizeof(double), cudaMemcpyHostToDevice));
    }
    d.uct = uct ? 1 : 0;
    if (uct) {
        // np.log(node.n) for every parent visit count a search can reach, from the host's libm like pow_half
        std::vector<double> tab(kPowTable);
        tab[0] = 0.0;
        for (int i = 1; i < kPowTable; ++i) tab[i] = log((double)i);
        double *p = nullptr;
        CK_E(cudaMalloc(&p, kPowTable * sizeof(double)));
        d.log_tab = p;
        CK_E(cudaMemcpy(p, tab.data(), kPowTable * sizeof(double), cudaMemcpyHostToDevice));
    }
    {
        std::vector<Slot> init(d.n_slots);
        memset(init.data(), 0, init.size() * sizeof(Slot));
        for (auto &s : init) { s.game = -1; s.pend_leaf = -1; s.tau = -1e301; s.phase = PH_NEED_ROOT; }
        CK_E(cudaMemcpy(d.slots, init.data(), init.size() * sizeof(Slot), cudaMemcpyHostToDevice));
    }
    return e;
}

void ck_engine_destroy(ck_engine *e) {
    if (!e) return;
    DeviceGuard g(e->dev.cfg.device);
    cudaDeviceSynchronize();
    engine_free(e);
}

int ck_engine_set_net(ck_engine *e, int which, ck_net *net) {
    if (!e || which < 0 || which > 1 || !net) return fail(CK_ERR_ARG, "ck_engine_set_net: bad arguments");
    if (net->device != e->dev.cfg.device) return fail(CK_ERR_ARG, "ck_engine_set_net: net lives on another device");
    e->net[which] = net;
    e->net_gen[which] = 0;              // (generations start at 1) the next run clears the evaluation cache
    return CK_OK;
}

// BUDGET of the searches that start (or are under way) from now on; the trees already built are kept.  bench.py
// uses it to desynchronise the slots with a cheap pre-roll before the timed region.
int ck_engine_set_budget(ck_engine *e, int32_t budget) {
    if (!e || budget < 1) return fail(CK_ERR_ARG, "ck_engine_set_budget: budget must be >= 1");
    EngineDev &d = e->dev;
    int64_t need = (int64_t)budget * CK_MAX_CHILDREN;       // room for a whole search in the worst case (see ck_engine_create)
    d.compact_need = (int32_t)(need > d.cap / 2 ? d.cap / 2 : need);
    d.cfg.budget = budget;
    return CK_OK;
}

int ck_engine_set_profile(ck_engine *e, int on) {
    if (!e) return fail(CK_ERR_ARG, "ck_engine_set_profile: null engine");
    e->profile = on != 0;
    return CK_OK;
}

// cached evaluations are only valid for the weights they were computed with
static int engine_check_cache(ck_engine *e) {
    EngineDev &d = e->dev;
    if (!d.cache || d.cfg.evaluator != CK_EVAL_NET) return CK_OK;
    bool stale = false;
    for (int k = 0; k < 2; ++k)
        if (e->net[k] && e->net[k]->weights_gen != e->net_gen[k]) { stale = true; e->net_gen[k] = e->net[k]->weights_gen; }
    if (stale) CK_CUDA(cudaMemsetAsync(d.cache, 0, (size_t)d.n_slots * d.cache_entries * kCacheU4 * sizeof(uint4), e->stream));
    return CK_OK;
}

static int engine_reset_slots(ck_engine *e, int manual) {
    EngineDev &d = e->dev;
    std::vector<Slot> init(d.n_slots);
    memset(init.data(), 0, init.size() * sizeof(Slot));
    for (auto &s : init) { s.game = -1; s.pend_leaf = -1; s.tau = -1e301; s.phase = PH_NEED_ROOT; s.manual = manual; }
    CK_CUDA(cudaMemcpyAsync(d.slots, init.data(), init.size() * sizeof(Slot), cudaMemcpyHostToDevice, e->stream));
    CK_CUDA(cudaMemsetAsync(d.ctr, 0, sizeof(Counters), e->stream));
    CK_CUDA(cudaStreamSynchronize(e->stream));
    return CK_OK;
}

int ck_engine_begin(ck_engine *e, int64_t n_games) {
    if (!e || n_games < 1 || n_games > (1 << 30)) return fail(CK_ERR_ARG, "ck_engine_begin: bad n_games");
    EngineDev &d = e->dev;
    DeviceGuard g(d.cfg.device);
    if (d.cfg.evaluator == CK_EVAL_NET) {
        for (int k = 0; k < (d.cfg.arena ? 2 : 1); ++k)
            if (!e->net[k] || !e->net[k]->have_weights) return fail(CK_ERR_NO_NET, "ck_engine_begin: evaluator is the network but no weights are attached");
    }
    if ((size_t)n_games > e->res_cap) {
        cudaFree(d.results); d.results = nullptr; e->res_cap = 0;
        CK_CUDA(cudaMalloc(&d.results, (size_t)n_games * sizeof(ck_game_result)));
        e->res_cap = (size_t)n_games;
    }
    if (d.cfg.keep_records && (size_t)n_games * d.max_rec > e->rec_cap) {
        cudaFree(d.rec); d.rec = nullptr; e->rec_cap = 0;
        CK_CUDA(cudaMalloc(&d.rec, (size_t)n_games * d.max_rec * sizeof(ck_record)));
        e->rec_cap = (size_t)n_games * d.max_rec;
    }
    CK_CUDA(cudaMemsetAsync(d.results, 0xFF, (size_t)n_games * sizeof(ck_game_result), e->stream));
    // records are handed out whole: entries of action[] / visits[] beyond n_children must read as zero
    if (d.cfg.keep_records) CK_CUDA(cudaMemsetAsync(d.rec, 0, (size_t)n_games * d.max_rec * sizeof(ck_record), e->stream));
    d.n_games = (int32_t)n_games;
    const int stride = d.cfg.game_id_stride ? d.cfg.game_id_stride : 1;
    d.arena_half = (int32_t)((n_games * stride) / 2);
    e->n_games = n_games;
    e->fetched.assign((size_t)n_games, 0);
    int rc = engine_reset_slots(e, 0);
    if (rc != CK_OK) return rc;
    e->eval_pending[0] = e->eval_pending[1] = false;
    e->begun = true;
    return CK_OK;
}

static void launch_tree_step(ck_engine *e) {
    EngineDev &d = e->dev;
    const int grid = (d.n_slots + kWarpsPerBlock - 1) / kWarpsPerBlock;
    d.round += 1;
    d.grp = 0; d.slot0 = 0; d.slot_n = d.n_slots;
    if (d.uct) tree_step_kernel<true, kWarpsPerBlock><<<grid, kWarpsPerBlock * 32, 0, e->stream>>>(d);
    else tree_step_kernel<false, kWarpsPerBlock><<<grid, kWarpsPerBlock * 32, 0, e->stream>>>(d);
}

// Overlapped self-play (two slot groups): one round of group g = [tree stream] wait for the group's previous
// evaluation -> round_begin -> tree_step of its slots (one-warp blocks: they fit next to the resident tower CTAs of the
// OTHER group's evaluation) -> [main stream] wait for that -> tower + heads of the group's leaves.  Enqueued g = 0, 1, 0, 1,
// ...: the tree work of one group runs while the tensor pipe evaluates the other.
static int engine_round_group(ck_engine *e, int g, int *launches, cudaEvent_t *prof = nullptr) {
    EngineDev &d = e->dev;
    const int half = d.n_slots / 2;
    EngineDev k = d;
    k.grp = g; k.slot0 = g ? half : 0; k.slot_n = g ? d.n_slots - half : half;
    cudaStream_t ts = e->stream_b;
    if (e->eval_pending[g]) CK_CUDA(cudaStreamWaitEvent(ts, e->ev_eval[g], 0));
    round_begin_kernel<<<1, 32, 0, ts>>>(k);
    tree_step_kernel<false, 1><<<k.slot_n, 32, 0, ts>>>(k);
    CK_CUDA(cudaGetLastError());
    CK_CUDA(cudaEventRecord(e->ev_tree[g], ts));
    CK_CUDA(cudaStreamWaitEvent(e->stream, e->ev_tree[g], 0));
    if (prof) { CK_CUDA(cudaEventRecord(prof[0], e->stream)); e->net[0]->ev_after_tower = prof[1]; }
    int rc = net_forward_rows(e->net[0], d.leaves[g], k.slot_n, &d.ctr->batch_count[g], d.policy[g], d.value[g], e->stream, launches, g);
    if (prof) e->net[0]->ev_after_tower = nullptr;
    if (rc != CK_OK) return rc;
    if (prof) CK_CUDA(cudaEventRecord(prof[2], e->stream));
    CK_CUDA(cudaEventRecord(e->ev_eval[g], e->stream));
    e->eval_pending[g] = true;
    if (launches) *launches += 2;
    return CK_OK;
}

static int engine_eval(ck_engine *e, int *launches) {
    EngineDev &d = e->dev;
    // arena: the two networks' batches are independent (each at most half of the slots), so the second one
    // is evaluated on its own stream and its CTAs run on the SMs the first launch leaves idle
    const bool fork = d.cfg.arena && e->stream_b != nullptr && e->net[0] != e->net[1];     // one net object = one scratch area
    if (fork) {
        CK_CUDA(cudaEventRecord(e->ev_fork, e->stream));
        CK_CUDA(cudaStreamWaitEvent(e->stream_b, e->ev_fork, 0));
    }
    for (int k = 0; k < (d.cfg.arena ? 2 : 1); ++k) {
        cudaStream_t st = (fork && k == 1) ? e->stream_b : e->stream;
        int kind = d.cfg.evaluator;
        if (k == 1 && d.cfg.evaluator_p2 >= 0) kind = d.cfg.evaluator_p2;
        if (kind == CK_EVAL_ROLLOUT || kind == CK_EVAL_ROLLOUT_HASH) {
            // about 2048 warps keep every SM busy; fewer lanes per warp below that (measured at 4 Ki / 16 Ki / 64 Ki
            // games: 2 / 8 / 32 lanes are the fastest, profiles/r1o_uct_playouts.jsonl)
            static const int lanes_env = getenv("CK_PLAYOUT_LANES_LOG2") ? atoi(getenv("CK_PLAYOUT_LANES_LOG2")) : -1;
            int lanes_log2 = lanes_env;
            if (lanes_log2 < 0 || lanes_log2 > 5) {
                lanes_log2 = 0;
                while (lanes_log2 < 5 && (d.n_slots >> (lanes_log2 + 1)) >= 2048) ++lanes_log2;
            }
            const int64_t threads = (((int64_t)d.n_slots + (1 << lanes_log2) - 1) >> lanes_log2) * 32;
            playout_eval_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(d.leaves[k], &d.ctr->batch_count[k], kind, d.cfg.seed,
                                                                                  d.round, d.value[k], lanes_log2);
            if (launches) *launches += 1;
        } else if (kind == CK_EVAL_NET) {
            int rc = net_forward_rows(e->net[k], d.leaves[k], d.n_slots, &d.ctr->batch_count[k], d.policy[k], d.value[k], st, launches);
            if (rc != CK_OK) return rc;
        } else {
            stub_eval_kernel<<<(d.n_slots * 32 + 127) / 128, 128, 0, st>>>(d.leaves[k], &d.ctr->batch_count[k], kind, d.policy[k], d.value[k]);
            if (launches) *launches += 1;
        }
    }
    if (fork) {
        CK_CUDA(cudaEventRecord(e->ev_join, e->stream_b));
        CK_CUDA(cudaStreamWaitEvent(e->stream, e->ev_join, 0));
    }
    return CK_OK;
}

// One lock-step round = tree kernel + evaluation of the staged leaves.
static int engine_round(ck_engine *e, int *launches) {
    EngineDev &d = e->dev;
    if (d.groups == 2) {
        int rc = engine_round_group(e, 0, launches);
        return rc != CK_OK ? rc : engine_round_group(e, 1, launches);
    }
    round_begin_kernel<<<1, 32, 0, e->stream>>>(d);      // batch cap of the round, counters to zero
    launch_tree_step(e);
    if (launches) *launches += 2;
    CK_CUDA(cudaGetLastError());
    return engine_eval(e, launches);
}

static int engine_poll(ck_engine *e) {
    if (e->dev.groups == 2) CK_CUDA(cudaStreamSynchronize(e->stream_b));       // the tree stream runs ahead of the main stream's waits
    CK_CUDA(cudaMemcpyAsync(e->h_ctr, e->dev.ctr, sizeof(Counters), cudaMemcpyDeviceToHost, e->stream));
    int32_t *h_range = reinterpret_cast<int32_t *>(e->h_ctr + 1);
    h_range[0] = h_range[1] = 0;
    if (e->dev.cfg.evaluator == CK_EVAL_NET)
        for (int k = 0; k < 2; ++k)
            if (e->net[k]) CK_CUDA(cudaMemcpyAsync(h_range + k, e->net[k]->d_range_flag, sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    CK_CUDA(cudaStreamSynchronize(e->stream));
    if (h_range[0] | h_range[1]) {
        for (int k = 0; k < 2; ++k) if (e->net[k]) cudaMemsetAsync(e->net[k]->d_range_flag, 0, sizeof(int32_t), e->stream);
        return fail(CK_ERR_NET_RANGE, "engine: a network activation left the split-fp16 range of the tensor-core evaluator "
                                      "(searches since the last check used invalid priors); use CK_NET_IMPL_SIMT for these weights");
    }
    if (e->h_ctr->error != 0) {
        const int code = e->h_ctr->error;
        return fail(code, code == CK_ERR_POOL_OVERFLOW ? "engine: node pool overflow (raise pool_cap)" :
                          code == CK_ERR_DEPTH ? "engine: tree deeper than 128 plies" : "engine: device-side state error (history capacity?)");
    }
    return CK_OK;
}

int ck_engine_run(ck_engine *e, int64_t n_steps, ck_run_stats *stats) {
    if (!e || !e->begun) return fail(CK_ERR_STATE, "ck_engine_run: call ck_engine_begin first");
    EngineDev &d = e->dev;
    DeviceGuard g(d.cfg.device);
    int rc = engine_check_cache(e);
    if (rc != CK_OK) return rc;
    rc = engine_poll(e);
    if (rc != CK_OK) return rc;
    const Counters before = *e->h_ctr;
    unsigned long long *d_tot = e->d_tot;
    CK_CUDA(cudaMemsetAsync(d_tot, 0, 6 * sizeof(unsigned long long), e->stream));
    sum_slots_kernel<<<32, 128, 0, e->stream>>>(d, d_tot);
    int launches = 0;
    double eval_ms = 0.0, tower_ms = 0.0;
    CK_CUDA(cudaEventRecord(e->ev0, e->stream));
    int64_t steps = 0;
    const int check = 16;
    const int per_round_ev = 3 * d.groups;               // before the evaluation, after the tower, after the heads -- per slot group
    if (e->profile && e->prof_ev.size() < (size_t)per_round_ev * check) {
        while (e->prof_ev.size() < (size_t)per_round_ev * check) { cudaEvent_t ev; CK_CUDA(cudaEventCreate(&ev)); e->prof_ev.push_back(ev); }
    }
    if (d.groups == 2) {                                 // the tree stream starts after whatever the main stream did before this call
        CK_CUDA(cudaEventRecord(e->ev_fork_run, e->stream));
        CK_CUDA(cudaStreamWaitEvent(e->stream_b, e->ev_fork_run, 0));
    }
    for (;;) {
        const int64_t chunk = n_steps > 0 ? std::min<int64_t>(check, n_steps - steps) : check;
        for (int64_t i = 0; i < chunk; ++i) {
            if (e->profile && d.groups == 2) {
                for (int g = 0; g < 2; ++g) {
                    rc = engine_round_group(e, g, &launches, &e->prof_ev[6 * i + 3 * g]);
                    if (rc != CK_OK) return rc;
                }
            } else if (e->profile) {
                round_begin_kernel<<<1, 32, 0, e->stream>>>(d);
                launch_tree_step(e);
                launches += 2;
                CK_CUDA(cudaEventRecord(e->prof_ev[3 * i], e->stream));
                if (e->net[0]) e->net[0]->ev_after_tower = e->prof_ev[3 * i + 1];
                rc = engine_eval(e, &launches);
                if (e->net[0]) e->net[0]->ev_after_tower = nullptr;
                if (rc != CK_OK) return rc;
                CK_CUDA(cudaEventRecord(e->prof_ev[3 * i + 2], e->stream));
            } else {
                rc = engine_round(e, &launches);
                if (rc != CK_OK) return rc;
            }
        }
        steps += chunk;
        rc = engine_poll(e);
        if (rc != CK_OK) return rc;
        if (e->profile) {
            for (int64_t i = 0; i < chunk * d.groups; ++i) {
                float a = 0.f, b = 0.f;
                cudaEventElapsedTime(&a, e->prof_ev[3 * i], e->prof_ev[3 * i + 2]);
                if (d.cfg.evaluator == CK_EVAL_NET && cudaEventElapsedTime(&b, e->prof_ev[3 * i], e->prof_ev[3 * i + 1]) == cudaSuccess) tower_ms += b;
                eval_ms += a;
            }
            cudaGetLastError();
        }
        if (n_steps > 0 && steps >= n_steps) break;
        if (n_steps <= 0 && e->h_ctr->active[0] + e->h_ctr->active[1] == 0 && e->h_ctr->batch_count[0] == 0 && e->h_ctr->batch_count[1] == 0) break;
        if (n_steps <= 0 && e->h_ctr->games_finished >= (unsigned long long)e->n_games) break;
    }
    CK_CUDA(cudaEventRecord(e->ev1, e->stream));
    sum_slots_kernel<<<32, 128, 0, e->stream>>>(d, d_tot + 3);
    unsigned long long tot[6];
    CK_CUDA(cudaMemcpyAsync(tot, d_tot, sizeof(tot), cudaMemcpyDeviceToHost, e->stream));
    CK_CUDA(cudaStreamSynchronize(e->stream));
    e->total_steps += (uint64_t)steps;
    if (stats) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e->ev0, e->ev1);
        memset(stats, 0, sizeof(*stats));
        const Counters &now = *e->h_ctr;
        stats->steps = (uint64_t)steps;
        stats->games_finished = now.games_finished - before.games_finished;
        stats->moves = now.moves - before.moves;
        stats->nodes_created = now.nodes - before.nodes;
        stats->compactions = now.compactions - before.compactions;
        stats->gpu_ms = ms;
        stats->eval_ms = eval_ms;
        stats->tower_ms = tower_ms;
        stats->kernel_launches = (uint64_t)launches;
        stats->sims = tot[3] - tot[0]; stats->nn_evals = tot[4] - tot[1]; stats->cache_hits = tot[5] - tot[2];
    }
    return CK_OK;
}


int ck_selfplay_run(ck_engine *e, int64_t n_games, ck_run_stats *stats) {
    if (!e) return fail(CK_ERR_ARG, "ck_selfplay_run: null engine");
    if (e->dev.cfg.arena) return fail(CK_ERR_ARG, "ck_selfplay_run: engine was created for arena play");
    int rc = ck_engine_begin(e, n_games);
    if (rc != CK_OK) return rc;
    return ck_engine_run(e, 0, stats);
}

int ck_arena_run(ck_engine *e, int64_t n_games, ck_run_stats *stats) {
    if (!e) return fail(CK_ERR_ARG, "ck_arena_run: null engine");
    if (!e->dev.cfg.arena) return fail(CK_ERR_ARG, "ck_arena_run: engine was created for self-play");
    int rc = ck_engine_begin(e, n_games);
    if (rc != CK_OK) return rc;
    return ck_engine_run(e, 0, stats);
}

int64_t ck_games_finished(ck_engine *e) {
    if (!e || !e->begun) return 0;
    DeviceGuard g(e->dev.cfg.device);
    if (engine_poll(e) != CK_OK) return -1;
    return (int64_t)e->h_ctr->games_finished;
}

// finished games in local game order (unfinished ones are skipped)
int ck_games_fetch(ck_engine *e, ck_game_result *out, int64_t cap) {
    if (!e || !e->begun || !out) return fail(CK_ERR_ARG, "ck_games_fetch: bad arguments");
    DeviceGuard g(e->dev.cfg.device);
    std::vector<ck_game_result> all((size_t)e->n_games);
    CK_CUDA(cudaMemcpy(all.data(), e->dev.results, all.size() * sizeof(ck_game_result), cudaMemcpyDeviceToHost));
    int64_t k = 0;
    for (const ck_game_result &r : all) {
        if (r.game < 0 || r.outcome < 0) continue;
        if (k < cap) out[k] = r;
        ++k;
    }
    return k <= cap ? CK_OK : fail(CK_ERR_ARG, "ck_games_fetch: buffer too small");
}

int64_t ck_records_count(ck_engine *e) {
    if (!e || !e->begun || !e->dev.cfg.keep_records) return 0;
    DeviceGuard g(e->dev.cfg.device);
    std::vector<ck_game_result> all((size_t)e->n_games);
    if (cudaMemcpy(all.data(), e->dev.results, all.size() * sizeof(ck_game_result), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    int64_t k = 0;
    for (const ck_game_result &r : all) if (r.game >= 0 && r.outcome >= 0) k += r.n_records;
    return k;
}

// records of all finished games, game by game in local game order, plies in order
int ck_records_fetch(ck_engine *e, ck_record *out, int64_t cap) {
    if (!e || !e->begun || !out) return fail(CK_ERR_ARG, "ck_records_fetch: bad arguments");
    if (!e->dev.cfg.keep_records) return fail(CK_ERR_STATE, "ck_records_fetch: engine was created with keep_records = 0");
    DeviceGuard g(e->dev.cfg.device);
    std::vector<ck_game_result> all((size_t)e->n_games);
    CK_CUDA(cudaMemcpy(all.data(), e->dev.results, all.size() * sizeof(ck_game_result), cudaMemcpyDeviceToHost));
    int64_t k = 0;
    for (size_t gi = 0; gi < all.size(); ++gi) {
        const ck_game_result &r = all[gi];
        if (r.game < 0 || r.outcome < 0) continue;
        if (k + r.n_records > cap) return fail(CK_ERR_ARG, "ck_records_fetch: buffer too small");
        CK_CUDA(cudaMemcpy(out + k, e->dev.rec + gi * e->dev.max_rec, (size_t)r.n_records * sizeof(ck_record), cudaMemcpyDeviceToHost));
        k += r.n_records;
    }
    return CK_OK;
}

// records of the games that finished since the previous call (bench e2e / streaming consumers)
int ck_records_fetch_new(ck_engine *e, ck_record *out, int64_t cap, int64_t *n_out, int64_t *n_games_out) {
    if (!e || !e->begun || !n_out) return fail(CK_ERR_ARG, "ck_records_fetch_new: bad arguments");
    DeviceGuard g(e->dev.cfg.device);
    std::vector<ck_game_result> all((size_t)e->n_games);
    CK_CUDA(cudaMemcpy(all.data(), e->dev.results, all.size() * sizeof(ck_game_result), cudaMemcpyDeviceToHost));
    int64_t k = 0, ng = 0;
    for (size_t gi = 0; gi < all.size(); ++gi) {
        const ck_game_result &r = all[gi];
        if (r.game < 0 || r.outcome < 0 || e->fetched[gi]) continue;
        if (e->dev.cfg.keep_records && out) {
            if (k + r.n_records > cap) break;
            CK_CUDA(cudaMemcpy(out + k, e->dev.rec + gi * e->dev.max_rec, (size_t)r.n_records * sizeof(ck_record), cudaMemcpyDeviceToHost));
            k += r.n_records;
        }
        e->fetched[gi] = 1;
        ++ng;
    }
    *n_out = k;
    if (n_games_out) *n_games_out = ng;
    return CK_OK;
}

// packed records of all finished games into caller-provided DEVICE buffers (NULL buffers: sizes only)
int ck_records_pack_device(ck_engine *e, ck_record_hdr *d_hdr, int64_t hdr_cap, uint32_t *d_words, int64_t word_cap,
                           int64_t *n_records, int64_t *n_words) {
    if (!e || !e->begun || !n_records || !n_words) return fail(CK_ERR_ARG, "ck_records_pack_device: bad arguments");
    if (!e->dev.cfg.keep_records) return fail(CK_ERR_STATE, "ck_records_pack_device: engine was created with keep_records = 0");
    EngineDev &d = e->dev;
    // a child word holds 23 bits of visits; visit counts are bounded by the simulations of all searches of a game
    if ((int64_t)d.cfg.budget * d.max_rec >= (1 << 23)) return fail(CK_ERR_ARG, "ck_records_pack_device: BUDGET x plies exceeds the packed format's 23-bit visit counts; use ck_records_fetch");
    DeviceGuard g(d.cfg.device);
    const int64_t ng = e->n_games;
    int64_t *d_counts = nullptr;
    CK_CUDA(cudaMalloc(&d_counts, (size_t)ng * 4 * sizeof(int64_t)));           // counts, then offsets
    int64_t *d_offs = d_counts + 2 * ng;
    std::vector<int64_t> counts((size_t)2 * ng), offs((size_t)2 * ng);
    pack_count_kernel<<<(unsigned)((ng + 127) / 128), 128, 0, e->stream>>>(d, d_counts);
    cudaError_t ce = cudaMemcpyAsync(counts.data(), d_counts, (size_t)2 * ng * sizeof(int64_t), cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    int64_t nr = 0, nw = 0;
    for (int64_t i = 0; i < ng; ++i) { offs[2 * i] = nr; offs[2 * i + 1] = nw; nr += counts[2 * i]; nw += counts[2 * i + 1]; }
    *n_records = nr; *n_words = nw;
    int rc = CK_OK;
    if (ce == cudaSuccess && d_hdr && d_words) {
        if (nr > hdr_cap || nw > word_cap) rc = fail(CK_ERR_ARG, "ck_records_pack_device: buffers too small");
        else if (nr > 0) {
            ce = cudaMemcpyAsync(d_offs, offs.data(), (size_t)2 * ng * sizeof(int64_t), cudaMemcpyHostToDevice, e->stream);
            pack_write_kernel<<<(unsigned)((ng * 32 + 127) / 128), 128, 0, e->stream>>>(d, d_counts, d_offs, d_hdr, d_words);
            if (ce == cudaSuccess) ce = cudaGetLastError();
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
        }
    }
    cudaFree(d_counts);
    if (ce != cudaSuccess) return fail(CK_ERR_CUDA, std::string("ck_records_pack_device: ") + cudaGetErrorString(ce));
    return rc;
}

// the same into HOST buffers (a fifth of the bytes of ck_records_fetch over PCIe)
int ck_records_fetch_packed(ck_engine *e, ck_record_hdr *hdr, int64_t hdr_cap, uint32_t *words, int64_t word_cap,
                            int64_t *n_records, int64_t *n_words) {
    int rc = ck_records_pack_device(e, nullptr, 0, nullptr, 0, n_records, n_words);
    if (rc != CK_OK || !hdr || !words) return rc;
    if (*n_records > hdr_cap || *n_words > word_cap) return fail(CK_ERR_ARG, "ck_records_fetch_packed: buffers too small");
    if (*n_records == 0) return CK_OK;
    DeviceGuard g(e->dev.cfg.device);
    ck_record_hdr *d_hdr = nullptr;
    uint32_t *d_words = nullptr;
    CK_CUDA(cudaMalloc(&d_hdr, (size_t)*n_records * sizeof(ck_record_hdr)));
    cudaError_t ce = cudaMalloc(&d_words, (size_t)(*n_words + 1) * sizeof(uint32_t));
    if (ce == cudaSuccess) {
        rc = ck_records_pack_device(e, d_hdr, *n_records, d_words, *n_words + 1, n_records, n_words);
        if (rc == CK_OK) ce = cudaMemcpy(hdr, d_hdr, (size_t)*n_records * sizeof(ck_record_hdr), cudaMemcpyDeviceToHost);
        if (rc == CK_OK && ce == cudaSuccess) ce = cudaMemcpy(words, d_words, (size_t)*n_words * sizeof(uint32_t), cudaMemcpyDeviceToHost);
    }
    cudaFree(d_hdr); cudaFree(d_words);
    if (ce != cudaSuccess) return fail(CK_ERR_CUDA, std::string("ck_records_fetch_packed: ") + cudaGetErrorString(ce));
    return rc;
}

// Packed records -> ck_record structs on the host (no device work): the inverse of the pack kernels, for consumers of
// ck_records_fetch_packed / the pooled gather that want the full structs.  The legal-action planes are rebuilt from the
// children's action ids (plane = a >> 6, square of (x, y) = ((a >> 3) & 7, a & 7) is 4x + (y >> 1)).
int ck_records_unpack(const ck_record_hdr *hdr, int64_t n, const uint32_t *words, int64_t n_words, ck_record *out) {
    if (n < 0 || n_words < 0 || (n > 0 && (!hdr || !out)) || (n_words > 0 && !words))
        return fail(CK_ERR_ARG, "ck_records_unpack: bad arguments");
    std::vector<int64_t> offs((size_t)n + 1);
    offs[0] = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int nch = hdr[i].n_children;
        if (nch > CK_MAX_CHILDREN) return fail(CK_ERR_ARG, "ck_records_unpack: a header claims more than CK_MAX_CHILDREN children");
        offs[i + 1] = offs[i] + nch + (((hdr[i].flags & 1u) && nch == 0) ? 8 : 0);
    }
    if (offs[n] != n_words)
        return fail(CK_ERR_ARG, "packed records: " + std::to_string((long long)offs[n]) + " child words expected, " +
                                std::to_string((long long)n_words) + " present");
    auto work = [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            const ck_record_hdr &h = hdr[i];
            ck_record &r = out[i];
            memset(&r, 0, sizeof(r));
            r.pos = h.pos; r.q = h.q; r.root_w = h.root_w; r.root_n = h.root_n; r.game = h.game;
            r.ply = h.ply; r.chosen = h.chosen; r.n_children = h.n_children; r.plane5 = h.plane5; r.z = h.z;
            const uint32_t *w = words + offs[i];
            for (int c = 0; c < (int)h.n_children; ++c) {
                const uint32_t a = w[c] >> 23;
                r.action[c] = (uint16_t)a;
                r.visits[c] = w[c] & 0x7FFFFFu;
                r.mask[a >> 6] |= 1u << (4u * ((a >> 3) & 7u) + ((a & 7u) >> 1));
            }
            if ((h.flags & 1u) && h.n_children == 0) memcpy(r.mask, w, 8 * sizeof(uint32_t));
        }
    };
    const int64_t per = 1 << 16;
    int nt = (int)std::min<int64_t>((n + per - 1) / per, std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u));
    if (nt <= 1) { work(0, n); return CK_OK; }
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t) pool.emplace_back(work, n * t / nt, n * (t + 1) / nt);
    for (auto &th : pool) th.join();
    return CK_OK;
}

// ---- single-search API (slot 0) -----------------------------------------------------------
static int manual_slot(ck_engine *e, Slot *s) {
    CK_CUDA(cudaMemcpy(s, e->dev.slots, sizeof(Slot), cudaMemcpyDeviceToHost));
    if (!s->manual || s->game < 0) return fail(CK_ERR_STATE, "ck_tree_*: call ck_tree_set_root first");
    return CK_OK;
}

int ck_tree_set_root(ck_engine *e, const ck_pos *root, int32_t parent_player) {
    if (!e || !root) return fail(CK_ERR_ARG, "ck_tree_set_root: bad arguments");
    EngineDev &d = e->dev;
    DeviceGuard g(d.cfg.device);
    if (d.cfg.evaluator == CK_EVAL_NET && (!e->net[0] || !e->net[0]->have_weights))
        return fail(CK_ERR_NO_NET, "ck_tree_set_root: evaluator is the network but no weights are attached");
    if (e->res_cap < 1) {
        CK_CUDA(cudaMalloc(&d.results, sizeof(ck_game_result)));
        e->res_cap = 1;
    }
    d.n_games = 1; e->n_games = 1; d.arena_half = 1;
    int rc = engine_reset_slots(e, 1);
    if (rc != CK_OK) return rc;
    Slot s;
    memset(&s, 0, sizeof(s));
    const int player = (int)(root->meta & 1u);
    s.game = 0; s.manual = 1; s.manual_target = 0; s.pend_leaf = -1;
    s.phase = PH_NEED_ROOT; s.cur = player; s.tau = d.cfg.tau;
    s.buf[0] = 0; s.buf[1] = 1; s.scratch = 2;
    // history = [dummy previous state with the wanted parent player, root]
    ck_pos hist[2];
    hist[0] = *root;
    const int pp = parent_player >= 0 ? parent_player : 1 - player;
    hist[0].meta = (root->meta & ~1u) | (uint32_t)pp;
    hist[0].p1 = hist[0].p2 = hist[0].k = 0;
    hist[1] = *root;
    s.hist_len = 2;
    CK_CUDA(cudaMemcpy(d.hist, hist, sizeof(hist), cudaMemcpyHostToDevice));
    CK_CUDA(cudaMemcpy(d.slots, &s, sizeof(Slot), cudaMemcpyHostToDevice));
    e->begun = true;
    return CK_OK;
}

int ck_tree_search(ck_engine *e, int32_t sims) {
    if (!e || sims < 0) return fail(CK_ERR_ARG, "ck_tree_search: bad arguments");
    EngineDev &d = e->dev;
    DeviceGuard g(d.cfg.device);
    Slot s;
    int rc = manual_slot(e, &s);
    if (rc != CK_OK) return rc;
    rc = engine_check_cache(e);
    if (rc != CK_OK) return rc;
    s.manual_target = sims; s.sims_done = 0;       // BUDGET new simulations on top of inherited statistics (MCTS.py:217)
    s.search_id += 1;                              // a new search draws new exploration noise
    if (s.phase == PH_HALT) s.phase = PH_SEARCH;
    CK_CUDA(cudaMemcpy(d.slots, &s, sizeof(Slot), cudaMemcpyHostToDevice));
    for (;;) {
        for (int i = 0; i < 16; ++i) { rc = engine_round(e, nullptr); if (rc != CK_OK) return rc; }
        rc = engine_poll(e);
        if (rc != CK_OK) return rc;
        if (e->h_ctr->active[0] + e->h_ctr->active[1] == 0) break;
    }
    return CK_OK;
}

int ck_tree_root(ck_engine *e, uint32_t *n, float *w, int32_t *n_children) {
    if (!e) return fail(CK_ERR_ARG, "ck_tree_root: null engine");
    EngineDev &d = e->dev;
    DeviceGuard g(d.cfg.device);
    Slot s;
    int rc = manual_slot(e, &s);
    if (rc != CK_OK) return rc;
    uint4 st;
    CK_CUDA(cudaMemcpy(&st, d.stat + (size_t)s.buf[s.cur] * d.cap + s.root[s.cur], sizeof(st), cudaMemcpyDeviceToHost));
    if (n) *n = st.x;
    if (w) memcpy(w, &st.y, 4);
    if (n_children) *n_children = d.uct ? std::min<int32_t>((int32_t)st.z, (int32_t)((st.w >> 22) & 63u)) : (int32_t)((st.w >> 22) & 63u);
    return CK_OK;
}

// children of `node` (-1: the current root) in node.children order; idx receives their node ids
int ck_tree_children(ck_engine *e, int32_t node, int32_t *idx, ck_pos *pos, uint32_t *n, float *w, float *p,
                     int32_t *status, int32_t *count) {
    if (!e || !count) return fail(CK_ERR_ARG, "ck_tree_children: bad arguments");
    EngineDev &d = e->dev;
    DeviceGuard g(d.cfg.device);
    Slot s;
    int rc = manual_slot(e, &s);
    if (rc != CK_OK) return rc;
    const size_t base = (size_t)s.buf[s.cur] * d.cap;
    if (node < 0) node = s.root[s.cur];
    if (node >= s.alloc[s.cur]) return fail(CK_ERR_ARG, "ck_tree_children: node id out of range");
    uint4 st;
    CK_CUDA(cudaMemcpy(&st, d.stat + base + node, sizeof(st), cudaMemcpyDeviceToHost));
    int b = (int)((st.w >> 22) & 63u);
    const int fc = (int)(st.w & kFcMask);
    if (d.uct) b = std::min(b, (int)st.z);                // NEURAL_NET=False: only the children added so far exist
    *count = b;
    if (b == 0) return CK_OK;
    uint4 cs[CK_MAX_CHILDREN], cp[CK_MAX_CHILDREN];
    CK_CUDA(cudaMemcpy(cs, d.stat + base + fc, b * sizeof(uint4), cudaMemcpyDeviceToHost));
    CK_CUDA(cudaMemcpy(cp, d.pos + base + fc, b * sizeof(uint4), cudaMemcpyDeviceToHost));
    for (int i = 0; i < b; ++i) {
        if (idx) idx[i] = fc + i;
        if (pos) { pos[i].p1 = cp[i].x; pos[i].p2 = cp[i].y; pos[i].k = cp[i].z; pos[i].meta = cp[i].w; }
        if (n) n[i] = cs[i].x;
        if (w) memcpy(w + i, &cs[i].y, 4);
        if (p) memcpy(p + i, &cs[i].z, 4);
        if (status) status[i] = (int32_t)((cs[i].w >> 28) & 3u);
    }
    return CK_OK;
}

int ck_tree_root_children(ck_engine *e, ck_pos *pos, uint32_t *n, float *w, float *p, int32_t *status) {
    int32_t count = 0;
    return ck_tree_children(e, -1, nullptr, pos, n, w, p, status, &count);
}

int ck_tree_best_child(ck_engine *e, int32_t move_count, int32_t *index) {
    if (!e || !index) return fail(CK_ERR_ARG, "ck_tree_best_child: bad arguments");
    (void)move_count;                      // temperature sampling is the engine's business; the shim asks for the robust child
    uint32_t n[CK_MAX_CHILDREN];
    int32_t b = 0;
    int rc = ck_tree_children(e, -1, nullptr, nullptr, n, nullptr, nullptr, nullptr, &b);
    if (rc != CK_OK) return rc;
    if (b == 0) return fail(CK_ERR_STATE, "ck_tree_best_child: root has no children");
    int best = 0;
    for (int i = 1; i < b; ++i) if (n[i] > n[best]) best = i;     // np.argmax: first maximum (MCTS.py:236-238)
    *index = best;
    return CK_OK;
}

// MCTS.new_root_node (MCTS.py:281-288): `node` (a node id from ck_tree_children) becomes the root
// of the same tree and keeps its statistics; its link already names its parent's player.
int ck_tree_reroot(ck_engine *e, int32_t node) {
    if (!e) return fail(CK_ERR_ARG, "ck_tree_reroot: null engine");
    EngineDev &d = e->dev;
    DeviceGuard g(d.cfg.device);
    Slot s;
    int rc = manual_slot(e, &s);
    if (rc != CK_OK) return rc;
    if (node < 0 || node >= s.alloc[s.cur]) return fail(CK_ERR_ARG, "ck_tree_reroot: node id out of range");
    s.root[s.cur] = node;
    s.phase = PH_HALT; s.sims_done = 0;
    CK_CUDA(cudaMemcpy(d.slots, &s, sizeof(Slot), cudaMemcpyHostToDevice));
    if (d.cap - s.alloc[s.cur] < d.compact_need || d.cfg.compact_always) {
        // not enough room for another search: keep the subtree under the new root only.  Node ids change
        // (the root becomes node 0); ck_tree_epoch tells the caller that ids it holds are stale.
        manual_compact_kernel<<<1, 32, 0, e->stream>>>(d);
        CK_CUDA(cudaGetLastError());
        CK_CUDA(cudaStreamSynchronize(e->stream));
        e->tree_epoch += 1;
    }
    return CK_OK;
}

int64_t ck_tree_epoch(ck_engine *e) { return e ? e->tree_epoch : -1; }

int64_t ck_engine_pool_cap(ck_engine *e) { return e ? e->dev.cap : -1; }

int ck_tree_advance(ck_engine *e, int32_t child_index) {
    int32_t idx[CK_MAX_CHILDREN], b = 0;
    int rc = ck_tree_children(e, -1, idx, nullptr, nullptr, nullptr, nullptr, nullptr, &b);
    if (rc != CK_OK) return rc;
    if (child_index < 0 || child_index >= b) return fail(CK_ERR_ARG, "ck_tree_advance: child index out of range");
    return ck_tree_reroot(e, idx[child_index]);
}

int64_t ck_tree_node_count(ck_engine *e) {
    if (!e) return -1;
    DeviceGuard g(e->dev.cfg.device);
    Slot s;
    if (manual_slot(e, &s) != CK_OK) return -1;
    return s.alloc[s.cur];
}

}  // extern "C"
