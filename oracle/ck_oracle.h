/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the Checkers MCTS + NN-eval hot path.
 *
 * A plain-C, coordinate-based (8x8 array, NOT bitboard) restatement of the reference's
 * algorithms, deliberately written in a different style from the CUDA product code so
 * that the two are independent.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 *
 * Parity status: PINNED for move generation, outcome/draw rule, masked-prior glue, the
 * PUCT tree and the self-play loop -- checked against the unmodified reference Python
 * (Checkers.py / MCTS.py / training_pipeline.py run here through oracle/ref_harness.py)
 * and against the golden vectors in tests/golden/ that were generated from it.
 * The network arithmetic (TensorFlow/Keras, absent from /root/reference and from this
 * image) is NOT restated here; see oracle/net_oracle.py ("parity unpinned" for the NN).
 */
#ifndef CK_ORACLE_H
#define CK_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Compact position shared by oracle, C-ABI and tests (see include/ckb200.h).
 * Square s = 4*x + (y>>1) for the playable squares x%2 != y%2 (x=row, y=col,
 * Checkers.py:415-423).  meta: bit0 player to move (0=P1), bits1-7 rev (consecutive
 * reversible plies, saturating), bits8-16 action id (plane-6)*64+x*8+y of the move that
 * produced the position, bit17 action-valid, bits18-31 ply index (len(history)-1). */
typedef struct { uint32_t p1, p2, k, meta; } cko_pos;

#define CKO_MAX_CHILDREN 48
enum { CKO_ONGOING = 0, CKO_P1_WINS = 1, CKO_P2_WINS = 2, CKO_DRAW = 3 };

/* Checkers._check_moves + determine_outcome (Checkers.py:94-364).
 * children: raw _check_moves list in generation order; mask: planes 6..13 as 32-bit
 * square sets; status: determine_outcome; plane5: numerator n of state[5] = n/80.
 * returns the number of children. */
int cko_movegen(const cko_pos *pos, cko_pos *children, uint32_t mask[8], int *status, int *plane5);

void cko_start_position(cko_pos *pos);
uint64_t cko_perft(const cko_pos *pos, int depth, uint64_t *hops, uint64_t *continuations);

/* Checkers.predict glue (Checkers.py:434-437): mask, numpy pairwise float32 sum, divide.
 * policy_in/out are float32[512] in (plane-6, x, y) order. */
void cko_mask_renorm(const float *policy_in, const uint32_t mask[8], float *policy_out);

/* Evaluator callback == Checkers.predict's call into neural_net.predict
 * (Checkers.py:431-434): raw 512 softmax outputs + scalar value for one position. */
typedef void (*cko_eval_fn)(const cko_pos *pos, const uint32_t mask[8], int plane5,
                            float *policy512, float *value, void *ctx);
void cko_eval_uniform_zero(const cko_pos *, const uint32_t *, int, float *, float *, void *);
void cko_eval_uniform_material(const cko_pos *, const uint32_t *, int, float *, float *, void *);
/* deterministic pseudo-random peaky policy + value from an integer hash of the position;
 * exact float ops only so that the CUDA stub evaluator can reproduce it bit for bit */
void cko_eval_hash(const cko_pos *, const uint32_t *, int, float *, float *, void *);
/* same with the hash salted by *(uint32_t*)ctx (0 == cko_eval_hash) */
void cko_eval_hash_salted(const cko_pos *, const uint32_t *, int, float *, float *, void *);

typedef struct {
    double uct_c;        /* UCT_C */
    int32_t budget;      /* BUDGET (CONSTRAINT='rollout') */
    int32_t training;    /* TRAINING */
    double alpha, epsilon;            /* DIRICHLET_* */
    double tau, tau_decay;            /* TEMPERATURE_* */
    int32_t tau_decay_delay;          /* TEMP_DECAY_DELAY */
    int32_t terminate_cnt;            /* TERMINATE_CNT, <=0: no cap (tournament) */
    uint64_t seed;
    int32_t rollout;     /* NEURAL_NET=False: 0 = off (PUCT + evaluator); CKO_ROLLOUT_RANDOM / CKO_ROLLOUT_HASH =
                          * plain UCT with one playout per simulation (MCTS.py:78-89,113-115,132-143) */
    int32_t reserved;
} cko_cfg;
enum { CKO_ROLLOUT_OFF = 0, CKO_ROLLOUT_RANDOM = 1, CKO_ROLLOUT_HASH = 2 };

typedef struct cko_tree cko_tree;   /* one MCTS tree (MCTS_Node graph + root) */
typedef struct cko_game cko_game;   /* one self-play / arena game (two trees) */

/* --- single-tree API (MCTS.py) --------------------------------------------------- */
/* parent_player: player to move in game_env.history[-2] (MCTS.py:167-173); pass -1 for
 * "no previous state" (opposite of the root's player). */
cko_tree *cko_tree_new(const cko_pos *root, int parent_player, const cko_cfg *cfg,
                       cko_eval_fn eval, void *ctx);
void cko_tree_free(cko_tree *);
void cko_tree_search(cko_tree *, int sims);              /* MCTS.begin_tree_search */
/* root children in node.children order (legal list reversed); returns count */
int  cko_tree_root_children(const cko_tree *, cko_pos *pos, uint32_t *n, float *w, float *p,
                            int32_t *terminal);
void cko_tree_root_stats(const cko_tree *, uint32_t *n, float *w);
int  cko_tree_best_child(cko_tree *, int move_count);   /* MCTS.best_child; child index */
uint64_t cko_tree_node_count(const cko_tree *);
uint64_t cko_tree_nn_evals(const cko_tree *);

/* --- game API (training_pipeline._generate_data / _start_tournament) ------------- */
typedef struct {
    cko_pos  pos;          /* root state (training_pipeline.py:369) */
    uint32_t mask[8];      /* planes 6..13 of the root state */
    int32_t  plane5;       /* numerator of plane 5 */
    int32_t  n_children;   /* 0 for the terminal record (:406-409) */
    uint16_t action[CKO_MAX_CHILDREN];   /* child action ids, node.children order */
    uint32_t visits[CKO_MAX_CHILDREN];   /* child.n  (prob = n / sum n, :433-434) */
    float    q;            /* root q from the root player's view (:365-368) */
    int32_t  z;            /* _add_rewards (:439-455) */
    uint32_t root_n;       /* diagnostics: root N and W after the search */
    float    root_w;
    int32_t  chosen;       /* chosen child's action id, -1 for the terminal record */
} cko_record;

cko_game *cko_game_new(const cko_cfg *cfg, cko_eval_fn eval_p1, void *ctx1,
                       cko_eval_fn eval_p2, void *ctx2);
void cko_game_free(cko_game *);
/* plays one ply (search + move). returns 1 while the game continues, 0 when done. */
int  cko_game_play_ply(cko_game *);
void cko_game_state(const cko_game *, cko_pos *pos);
int  cko_game_outcome(const cko_game *);      /* CKO_* */
int  cko_game_move_count(const cko_game *);
int  cko_game_terminated(const cko_game *);   /* hit TERMINATE_CNT */
int  cko_game_num_records(const cko_game *);  /* incl. terminal record once done */
void cko_game_get_records(const cko_game *, cko_record *out);
uint64_t cko_game_total_sims(const cko_game *);
uint64_t cko_game_nn_evals(const cko_game *);
uint64_t cko_game_reroot_misses(const cko_game *);

/* random playout from pos (MCTS.default_policy non-NN branch, MCTS.py:132-143);
 * returns CKO_*; max_plies<=0: unlimited */
int cko_random_playout(const cko_pos *pos, uint64_t *rng_state, int *plies, int max_plies);
/* the same playout with np.random.randint replaced by a hash of the current position
 * (index = cko_hash_choice(position, n_legal), optionally salted with a game tag -- the evaluator ctx of
 * cko_game_new / cko_tree_new); deterministic twin for parity tests */
int cko_hash_playout(const cko_pos *pos, uint32_t salt, int *plies);
uint32_t cko_hash_choice(const cko_pos *pos, uint32_t n);

#ifdef __cplusplus
}
#endif
#endif
