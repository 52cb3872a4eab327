/* ckb200.h -- C ABI of libckb200.so: B200-native batched Checkers self-play engine.
 *
 * The reference (AlexMGitHub/Checkers-MCTS) has no FFI layer; its seams are duck-typed
 * Python objects.  Each entry point below names the reference code it
 * replaces (file:line into the reference tree).  All pointers are plain host pointers
 * unless the name ends in _device; no torch / CUDA types appear in any signature
 * (streams are passed as void*).  Every function returns CK_OK (0) or a CK_ERR_* code;
 * ck_last_error() gives the message of the last failure on the calling thread.
 *
 * There is no CPU fallback behind this ABI: every compute entry point runs CUDA
 * kernels compiled for sm_100a and fails with CK_ERR_CUDA when no device is usable.
 */
#ifndef CKB200_H
#define CKB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CK_ABI_VERSION 2
#define CK_MAX_CHILDREN 48          /* 12 kings x 4 directions */
#define CK_POLICY_SIZE 512          /* 8 action planes x 8 x 8 (Checkers.py:434) */
#define CK_NET_PARAM_COUNT 1321774  /* create_nn, training_pipeline.py:44-120 */

enum { CK_ONGOING = 0, CK_P1_WINS = 1, CK_P2_WINS = 2, CK_DRAW = 3 };
enum {
    CK_OK = 0, CK_ERR_CUDA = 1, CK_ERR_ARG = 2, CK_ERR_POOL_OVERFLOW = 3,
    CK_ERR_DEPTH = 4, CK_ERR_STATE = 5, CK_ERR_NOMEM = 6, CK_ERR_NO_NET = 7,
    CK_ERR_NET_RANGE = 8      /* a network activation left the range of the split-fp16 tensor-core path; results invalid */
};

/* Compact position.  Square s = 4*x + (y>>1) over the playable squares x%2 != y%2
 * (x = row, y = column, Checkers.py:415-423).  p1/p2: all pieces of each side; k: kings
 * of either side.  meta: bit0 player to move (0 = player1), bits1-7 `rev` (consecutive
 * reversible plies, saturating at 127), bits8-16 action id (plane-6)*64 + x*8 + y of the
 * move that produced the position (Checkers.py:138-143), bit17 action valid,
 * bits18-31 ply index (len(history)-1, saturating). */
typedef struct { uint32_t p1, p2, k, meta; } ck_pos;

/* Compact network input for one position: planes 0-4 from pos, plane 5 = plane5/80,
 * planes 6-13 = mask (Checkers.py:37-48, 431-432). */
typedef struct { uint32_t p1, p2, k, info /* bit0 player, bits8-15 plane5 numerator, bits16-31 game tag (stub evaluators only) */; uint32_t mask[8]; } ck_leaf;

const char *ck_last_error(void);
int ck_abi_version(void);
int ck_device_count(void);

/* ---- K1: legal successors + outcome -------------------------------------------------
 * Replaces Checkers._check_moves/_check_jumps/_check_king_jumps (Checkers.py:94-304) and
 * determine_outcome (:306-364) for a batch of n positions.  children[i*max_children + j]
 * is the j-th entry of the reference's raw legal list (generation order); counts[i] the
 * list length; masks[i*8 + p] plane 6+p as a 32-bit square set; status[i] CK_*;
 * plane5[i] the numerator of state[5] (= n/80).  Any output pointer may be NULL. */
int ck_movegen(int device, const ck_pos *pos, int64_t n, int32_t max_children,
               ck_pos *children, int32_t *counts, uint32_t *masks, uint8_t *status, uint8_t *plane5);
/* same on device buffers; launches on `stream` (cudaStream_t) and does not synchronise.
 * elapsed_ms (optional, host) makes the call time itself with CUDA events on `stream`. */
int ck_movegen_device(const ck_pos *d_pos, int64_t n, int32_t max_children, ck_pos *d_children,
                      int32_t *d_counts, uint32_t *d_masks, uint8_t *d_status, uint8_t *d_plane5,
                      void *stream);

/* Packed (CSR) variant on device buffers: the successors of position i are
 * d_children[d_offsets[i] .. d_offsets[i+1]) in the same list order, d_offsets has n+1 entries
 * (d_offsets[n] = total).  Successors that would land at or beyond child_cap are not written
 * (d_offsets still reports the full sizes, so a caller can size the buffer with d_children = NULL
 * first).  n <= 2^26.  This is the layout the cfg4 sweep is measured on: it moves only the
 * algorithmic bytes of SURVEY.md 8(d).  Calls on one device must not overlap (shared workspace). */
int ck_movegen_csr(int device, const ck_pos *pos, int64_t n, ck_pos *children, int64_t child_cap,
                   uint32_t *offsets, uint32_t *masks, uint8_t *status, uint8_t *plane5);   /* host buffers */
int ck_movegen_csr_device(const ck_pos *d_pos, int64_t n, ck_pos *d_children, int64_t child_cap,
                          uint32_t *d_offsets, uint32_t *d_masks, uint8_t *d_status, uint8_t *d_plane5,
                          void *stream);

/* ---- K4: random playouts -----------------------------------------------------------
 * MCTS.default_policy, non-NN branch (MCTS.py:132-143): uniformly random legal moves to
 * the end of the game.  outcome[i] CK_*; plies[i] the playout length.  max_plies <= 0:
 * unlimited. */
int ck_rollout(int device, const ck_pos *pos, int64_t n, uint64_t seed, int32_t max_plies,
               uint8_t *outcome, int32_t *plies);
/* the same on device-resident buffers, asynchronously on `stream` (the device the buffers live on must be current) */
int ck_rollout_device(const ck_pos *d_pos, int64_t n, uint64_t seed, int32_t max_plies,
                      uint8_t *d_outcome, int32_t *d_plies, void *stream);

/* ---- K3: policy/value network ------------------------------------------------------
 * Replaces neural_net.predict behind Checkers.predict (Checkers.py:425-438); the
 * architecture is create_nn (training_pipeline.py:44-120).  Weights are one float32 blob
 * of CK_NET_PARAM_COUNT values in Keras layer order and Keras layouts (conv kernels
 * [kh,kw,Cin,Cout], dense [in,out], BN gamma,beta,moving_mean,moving_variance):
 *   for L in conv0..conv6:      kernel, bias, gamma, beta, mean, var
 *   policy conv3x3 (128->128):  kernel, bias, gamma, beta, mean, var
 *   policy conv1x1 (128->8):    kernel, bias, gamma, beta, mean, var
 *   policy dense (512->512):    kernel, bias
 *   value conv1x1 (128->1):     kernel, bias, gamma, beta, mean, var
 *   value dense (64->64):       kernel, bias, gamma, beta, mean, var
 *   value dense (64->1):        kernel, bias
 */
typedef struct ck_net ck_net;
enum { CK_NET_IMPL_TC = 0 /* tcgen05 split-fp16 tower */, CK_NET_IMPL_SIMT = 1 /* fp32 CUDA-core cross-check */ };
ck_net *ck_net_create(int device);
void ck_net_destroy(ck_net *);
int ck_net_set_impl(ck_net *, int impl);
int ck_net_set_weights(ck_net *, const float *blob, int64_t count);          /* host blob */
int ck_net_set_weights_device(ck_net *, const float *d_blob, int64_t count); /* device blob (e.g. a torch tensor) */
/* raw softmax policy [n,512] and tanh value [n] (what Keras predict returns) */
int ck_net_forward(ck_net *, const ck_leaf *leaves, int64_t n, float *policy, float *value);
/* the same plus what north_star's 1e-5 contract is stated on: the policy logits [n,512] (input of the Dense softmax,
 * training_pipeline.py:97-100) and the value head's pre-tanh output [n] (:111-112).  Any output pointer may be NULL. */
int ck_net_forward_logits(ck_net *, const ck_leaf *leaves, int64_t n, float *policy, float *value, float *logits, float *value_pre);
/* debug / parity: the tower's outputs of the most recent tensor-core forward call of n positions (its batch must have
 * been exactly n): pflat [n,512] = policy conv1x1 + ReLU + BN flattened in (x,y,c) order (training_pipeline.py:89-97),
 * vconv [n,64] = value conv1x1 + ReLU + BN (:102-105).  Not valid with CK_TOWER=ss. */
int ck_net_last_features(ck_net *, int64_t n, float *pflat, float *vconv);
/* The tensor-core path keeps activations as split fp16 scaled by 2^4: a BatchNorm output with |a| >= 4094 does not
 * fit.  The kernels flag it; the host entry points above return CK_ERR_NET_RANGE, the engine reports it from
 * ck_engine_run, and after ck_net_forward_device the caller asks here (synchronises the device). */
int ck_net_range_status(ck_net *);
/* Keras signature: x float32 [n,8,8,14] channels-last (Checkers.py:431-433) */
int ck_net_forward_planes(ck_net *, const float *x, int64_t n, float *policy, float *value);
int ck_net_forward_device(ck_net *, const ck_leaf *d_leaves, int64_t n, float *d_policy, float *d_value,
                          void *stream);
/* Checkers.predict glue (Checkers.py:434-437): prior = policy*mask / sum(policy*mask) with
 * numpy's float32 pairwise summation order; masks[i*8+p]; bit-exact. */
int ck_mask_renorm(int device, const float *policy, const uint32_t *masks, int64_t n, float *prior);

/* ---- engine: self-play / arena (K2, K5, K6) -----------------------------------------
 * Replaces MCTS.tree_policy/select_child/backpropagation/begin_tree_search/best_child/
 * new_root_node (MCTS.py:59-295, 350-430) and the game loops
 * generate_Checkers_data._generate_data (training_pipeline.py:334-419) and
 * tournament_Checkers._start_tournament (:505-559). */
enum { CK_EVAL_NET = 0, CK_EVAL_UNIFORM_ZERO = 1, CK_EVAL_UNIFORM_MATERIAL = 2, CK_EVAL_HASH = 3,
       CK_EVAL_HASH_SALTED = 4 /* hash stub salted with the global game id: concurrent games differ */,
       /* NEURAL_NET=False (MCTS.py:78-89,113-115,132-146): plain UCT, one child added per visit, one playout to the
        * end of the game per simulation -- the reference's iteration-0 self-play.  Self-play engines only. */
       CK_EVAL_ROLLOUT = 5      /* uniformly random playouts */,
       CK_EVAL_ROLLOUT_HASH = 6 /* playout moves picked by a hash of the position (deterministic parity tests) */ };

typedef struct {
    int32_t device;
    int32_t n_slots;          /* concurrent games resident on the GPU */
    int32_t pool_cap;         /* nodes per tree buffer (0: default) */
    int32_t max_plies;        /* history capacity per game (0: default 2048) */
    int32_t budget;           /* BUDGET, CONSTRAINT='rollout' (MCTS.py:188-201) */
    int32_t training;         /* TRAINING */
    int32_t tau_decay_delay;  /* TEMP_DECAY_DELAY */
    int32_t terminate_cnt;    /* TERMINATE_CNT; <= 0: no ply cap (tournament) */
    double uct_c;             /* UCT_C */
    double alpha, epsilon;    /* DIRICHLET_ALPHA / DIRICHLET_EPSILON */
    double tau, tau_decay;    /* TEMPERATURE_TAU / TEMPERATURE_DECAY */
    uint64_t seed;
    int32_t evaluator;        /* CK_EVAL_* ; stubs exist for deterministic parity tests */
    int32_t evaluator_p2;     /* arena with stub evaluators: evaluator of the second net; -1: same */
    int32_t arena;            /* 0 self-play (one net); 1 arena: net 0 is player1 in games < n/2 (:523-528) */
    int32_t keep_records;     /* store training records (self-play) */
    int32_t reference_tau_quirk; /* 1: tau is never reset between games: a game inherits the tau its SLOT's
                                  * previous game ended with (0 once a game has outlasted the decay, so every later game is arg-max
                                  * play as in the reference); which slot claims a later game is not fixed when several end in one round */
    int32_t game_id_base;     /* global id of local game i = base + i*stride (multi-GPU sharding) */
    int32_t game_id_stride;   /* 0 is treated as 1 */
    int32_t max_terminal_sims_per_step; /* simulations ending in a terminal child that a slot may finish inside one round; 0: default 4 */
    int32_t compact_always;   /* 1: compact the kept subtree at every re-root (default: only when the pool runs low) */
    int32_t eval_cache_entries; /* evaluation cache entries per slot (128 B each, rounded up to a power of two); 0: default 16384;
                                 * < 0: no cache.  A leaf whose network input (position, side to move, plane 5) was evaluated
                                 * before in the same slot is expanded from the cached priors / value: same numbers, no network call */
    int32_t max_chain_per_step; /* simulations a slot may complete inside one round without a network evaluation
                                 * (terminal children + cache hits); 0: default 6 */
    int32_t stagger_budget;   /* warm start for throughput measurements: the first game a slot plays (staged game g) plays its first */
    int32_t stagger_plies;    /* hash(g) mod stagger_plies plies at stagger_budget sims/move, every later move (and game) at
                               * BUDGET, so the slots reach full-budget play at scattered stages of their games.  0: off */
    int32_t reserved0;
} ck_engine_cfg;

typedef struct {
    ck_pos   pos;             /* root state (training_pipeline.py:369) */
    uint32_t mask[8];         /* planes 6..13 */
    int32_t  plane5;          /* numerator of plane 5 */
    int32_t  n_children;      /* 0 for the terminal record (:406-409) */
    uint16_t action[CK_MAX_CHILDREN];  /* child action ids in node.children order */
    uint32_t visits[CK_MAX_CHILDREN];  /* child.n (prob = n / sum n, :433-434) */
    float    q;               /* root q from the root player's view (:365-368) */
    int32_t  z;               /* _add_rewards (:439-455) */
    uint32_t root_n;
    float    root_w;
    int32_t  chosen;          /* chosen action id, -1 for the terminal record */
    int32_t  game;            /* global game id */
    int32_t  ply;             /* index of the record inside its game */
} ck_record;

/* Packed record (the wire format of the iteration-end gather, training_pipeline.py:277-284,326-329): a fixed header
 * plus one 32-bit word per child, action id << 23 | visits, in node.children order.  The legal-action planes of a
 * searched position are its children's actions; flags bit 0: a terminal record (n_children = 0) whose position still has
 * legal moves (draw by the 80-ply rule) is followed by its 8 mask words instead. */
typedef struct {
    ck_pos   pos;
    float    q, root_w;
    uint32_t root_n;
    int32_t  game;
    uint16_t ply;
    int16_t  chosen;
    uint8_t  n_children, plane5;
    int8_t   z;
    uint8_t  flags;
} ck_record_hdr;             /* 40 bytes */

typedef struct {
    int32_t game;             /* global game id */
    int32_t outcome;          /* CK_* (adjudicated when terminated) */
    int32_t move_count;
    int32_t terminated;       /* hit TERMINATE_CNT */
    int32_t n_records;
    int32_t reroot_misses;    /* MCTS.py:292 would have raised */
    int32_t p1_net;           /* arena: which net played player1 */
    int32_t reserved;
    uint64_t sims, nn_evals;
} ck_game_result;

typedef struct {
    uint64_t sims;            /* root.selection() calls completed (MCTS.py:220,430) */
    uint64_t nn_evals;        /* leaf evaluations (expansions), including those served by the evaluation cache */
    uint64_t steps;           /* tree-step + eval rounds executed */
    uint64_t games_finished;
    uint64_t moves;
    uint64_t nodes_created;
    uint64_t compactions;
    double   gpu_ms;          /* CUDA-event time of the call on the engine's stream */
    double   eval_ms;         /* of which inside the evaluator kernels (profile mode; else 0) */
    double   tower_ms;        /* of which inside the tcgen05 tower kernel (profile mode; else 0) */
    uint64_t kernel_launches;
    uint64_t cache_hits;      /* expansions served by the evaluation cache: network evaluations run = nn_evals - cache_hits */
} ck_run_stats;

typedef struct ck_engine ck_engine;
ck_engine *ck_engine_create(const ck_engine_cfg *cfg);
void ck_engine_destroy(ck_engine *);
int ck_engine_set_net(ck_engine *, int which /*0|1*/, ck_net *net);
/* stage n_games new games (all from the start position); clears finished-game storage */
int ck_engine_begin(ck_engine *, int64_t n_games);
/* run at most n_steps lock-step rounds (n_steps <= 0: until every staged game finished) */
int ck_engine_run(ck_engine *, int64_t n_steps, ck_run_stats *stats);
/* convenience: begin + run to completion == _generate_data / _start_tournament */
int ck_selfplay_run(ck_engine *, int64_t n_games, ck_run_stats *stats);
int ck_arena_run(ck_engine *, int64_t n_games, ck_run_stats *stats);
int64_t ck_games_finished(ck_engine *);
int ck_games_fetch(ck_engine *, ck_game_result *out, int64_t cap);
int64_t ck_records_count(ck_engine *);
int ck_records_fetch(ck_engine *, ck_record *out, int64_t cap);
/* records of the games that finished since the previous call; *n_out records, *n_games_out games */
int ck_records_fetch_new(ck_engine *, ck_record *out, int64_t cap, int64_t *n_out, int64_t *n_games_out);
/* change BUDGET for the searches that start or are under way from the next ck_engine_run on (MCTS.computational_budget,
 * MCTS.py:188-201, reads the class attribute at every check, so the reference can be re-budgeted between moves too) */
int ck_engine_set_budget(ck_engine *, int32_t budget);
/* records of all finished games, packed (ck_record_hdr + child words), game by game in local game order, written into
 * caller-provided DEVICE buffers (e.g. torch tensors that a NCCL send reads directly).  With d_hdr = d_words = NULL only
 * the sizes are returned.  ck_records_fetch_packed: the same into host buffers. */
int ck_records_pack_device(ck_engine *, ck_record_hdr *d_hdr, int64_t hdr_cap, uint32_t *d_words, int64_t word_cap,
                           int64_t *n_records, int64_t *n_words);
int ck_records_fetch_packed(ck_engine *, ck_record_hdr *hdr, int64_t hdr_cap, uint32_t *words, int64_t word_cap,
                            int64_t *n_records, int64_t *n_words);
/* packed records -> ck_record structs, on the host (no engine, no device): what a consumer of ck_records_fetch_packed or of
 * the pooled gather (ckb200/dist.py) calls when it wants the full structs; fails if the child words do not add up */
int ck_records_unpack(const ck_record_hdr *hdr, int64_t n_records, const uint32_t *words, int64_t n_words, ck_record *out);
/* time the evaluator separately inside ck_engine_run (adds two events per step) */
int ck_engine_set_profile(ck_engine *, int on);

/* ---- single-search API behind the MCTS / MCTS_Node shim (MCTS.py:210-295) ------------
 * The engine must have been created with n_slots >= 1; these calls use slot 0 only. */
int ck_tree_set_root(ck_engine *, const ck_pos *root, int32_t parent_player /* -1: opposite of root's */);
int ck_tree_search(ck_engine *, int32_t sims);
int ck_tree_root(ck_engine *, uint32_t *n, float *w, int32_t *n_children);
int ck_tree_root_children(ck_engine *, ck_pos *pos, uint32_t *n, float *w, float *p, int32_t *status);
/* children of `node` (-1: the root) in node.children order; idx receives their node ids.  p: the prior
 * (with the playout evaluators there are no priors and the field holds engine bookkeeping) */
int ck_tree_children(ck_engine *, int32_t node, int32_t *idx, ck_pos *pos, uint32_t *n, float *w, float *p,
                     int32_t *status, int32_t *count);
int ck_tree_best_child(ck_engine *, int32_t move_count, int32_t *index);
/* make `node` (an id from ck_tree_children) the root of the same tree (MCTS.new_root_node).  When the
 * node pool cannot hold another search the subtree under the new root is compacted: the root becomes
 * node 0, every other id changes and ck_tree_epoch() advances by one. */
int ck_tree_reroot(ck_engine *, int32_t node);
/* number of times ck_tree_reroot has renumbered the tree; ids from ck_tree_children are valid only
 * while this value is unchanged */
int64_t ck_tree_epoch(ck_engine *);
/* re-root on the child with this index (MCTS.new_root_node for a one-ply advance) */
int ck_tree_advance(ck_engine *, int32_t child_index);
int64_t ck_tree_node_count(ck_engine *);
/* nodes one tree buffer of this engine can hold (pool_cap after defaults); a search that would exceed it fails with
 * CK_ERR_POOL_OVERFLOW, so open-ended callers (CONSTRAINT='time') compare it with ck_tree_node_count */
int64_t ck_engine_pool_cap(ck_engine *);

#ifdef __cplusplus
}
#endif
#endif
