/* TEST INFRASTRUCTURE ONLY -- see ck_oracle.h.  CPU restatement of the reference's
 * Checkers rules (Checkers.py), PUCT tree (MCTS.py) and self-play / arena game loops
 * (training_pipeline.py).  Coordinate-based on purpose: the CUDA product code is
 * bitboard-based, so agreement between the two is meaningful.
 */
#include "ck_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------
 * board helpers
 * ---------------------------------------------------------------------------------- */
enum { EMPTY = 0, P1_MAN = 1, P1_KING = 2, P2_MAN = 3, P2_KING = 4 };

typedef struct { int8_t sq[8][8]; } board_t;

#define META_PLAYER(m)   ((int)((m) & 1u))
#define META_REV(m)      ((int)(((m) >> 1) & 0x7Fu))
#define META_ACTION(m)   ((int)(((m) >> 8) & 0x1FFu))
#define META_HAS_ACT(m)  ((int)(((m) >> 17) & 1u))
#define META_PLY(m)      ((int)(((m) >> 18) & 0x3FFFu))

static uint32_t make_meta(int player, int rev, int action, int has_action, int ply) {
    if (rev > 127) rev = 127;
    if (ply > 0x3FFF) ply = 0x3FFF;
    return (uint32_t)player | ((uint32_t)rev << 1) | ((uint32_t)action << 8) |
           ((uint32_t)has_action << 17) | ((uint32_t)ply << 18);
}

static int sq_index(int x, int y) { return 4 * x + (y >> 1); }
static int playable(int x, int y) { return (x & 1) != (y & 1); }
static int inb(int v) { return v > -1 && v < 8; }

static void unpack(const cko_pos *p, board_t *b) {
    memset(b, 0, sizeof(*b));
    for (int x = 0; x < 8; ++x)
        for (int y = 0; y < 8; ++y) {
            if (!playable(x, y)) continue;
            uint32_t bit = 1u << sq_index(x, y);
            if (p->p1 & bit) b->sq[x][y] = (p->k & bit) ? P1_KING : P1_MAN;
            else if (p->p2 & bit) b->sq[x][y] = (p->k & bit) ? P2_KING : P2_MAN;
        }
}

static void pack(const board_t *b, cko_pos *p) {
    p->p1 = p->p2 = p->k = 0;
    for (int x = 0; x < 8; ++x)
        for (int y = 0; y < 8; ++y) {
            int v = b->sq[x][y];
            if (!v) continue;
            uint32_t bit = 1u << sq_index(x, y);
            if (v == P1_MAN || v == P1_KING) p->p1 |= bit; else p->p2 |= bit;
            if (v == P1_KING || v == P2_KING) p->k |= bit;
        }
}

void cko_start_position(cko_pos *pos) {
    board_t b;
    memset(&b, 0, sizeof(b));
    for (int row = 0; row < 8; ++row)          /* Checkers.init_board, Checkers.py:415-423 */
        for (int col = 0; col < 8; ++col)
            if (row % 2 != col % 2) {
                if (row < 3) b.sq[row][col] = P1_MAN;
                else if (row > 4) b.sq[row][col] = P2_MAN;
            }
    pack(&b, pos);
    pos->meta = make_meta(0, 0, 0, 0, 0);
}

static int is_opp(int v, int player) {
    return player == 0 ? (v == P2_MAN || v == P2_KING) : (v == P1_MAN || v == P1_KING);
}

/* ------------------------------------------------------------------------------------
 * move generation: Checkers._check_moves / _check_jumps / _check_king_jumps
 * ---------------------------------------------------------------------------------- */
typedef struct {
    board_t b;
    int next_player;
    int plane, x, y;
    int reversible;     /* king, non-capture */
} gen_t;

/* Does the piece that just landed on (x,y) have another hop?  Mirrors the recursive
 * call at Checkers.py:230-232 / 279-281: opponent pieces are read from the child's
 * planes (temp_state) but emptiness of the landing square from the PARENT's stale
 * `board` sum (Checkers.py:117,219,272). */
static int man_can_hop(const board_t *child, const board_t *stale, int x, int y, int fwd, int player) {
    for (int ydir = -1; ydir < 2; ydir += 2)
        if (inb(y + ydir) && inb(x + fwd))
            if (is_opp(child->sq[x + fwd][y + ydir], player))
                if (inb(y + 2 * ydir) && inb(x + 2 * fwd))
                    if (stale->sq[x + 2 * fwd][y + 2 * ydir] == EMPTY) return 1;
    return 0;
}

static int king_can_hop(const board_t *child, const board_t *stale, int x, int y, int player) {
    for (int ydir = -1; ydir < 2; ydir += 2)
        for (int fwd = -1; fwd < 2; fwd += 2)
            if (inb(x + fwd) && inb(y + ydir))
                if (is_opp(child->sq[x + fwd][y + ydir], player))
                    if (inb(x + 2 * fwd) && inb(y + 2 * ydir))
                        if (stale->sq[x + 2 * fwd][y + 2 * ydir] == EMPTY) return 1;
    return 0;
}

static int jump_plane(int fwd, int ydir) {      /* Checkers.py:238-253 */
    if (fwd == 1 && ydir == 1) return 13;
    if (fwd == 1 && ydir == -1) return 12;
    if (fwd == -1 && ydir == 1) return 11;
    return 10;
}

static int check_moves(const board_t *b, int player, gen_t *out, uint32_t mask[8]) {
    gen_t moves[CKO_MAX_CHILDREN + 8], jumps[CKO_MAX_CHILDREN + 8];
    int nm = 0, nj = 0;
    const int man = player == 0 ? P1_MAN : P2_MAN, king = man + 1;
    const int fwd = player == 0 ? 1 : -1;
    memset(mask, 0, 8 * sizeof(uint32_t));
    /* men, row-major (np.where order, Checkers.py:111-124) */
    for (int x = 0; x < 8; ++x)
        for (int y = 0; y < 8; ++y) {
            if (b->sq[x][y] != man) continue;
            for (int side = 1; side >= -1; side -= 2) {       /* y+1 first, then y-1 (:125,145) */
                int ny = y + side, nx = x + fwd;
                if (!inb(ny) || !inb(nx)) continue;
                if (b->sq[nx][ny] != EMPTY) continue;
                gen_t *g = &moves[nm++];
                g->b = *b;
                g->b.sq[x][y] = EMPTY;
                int kinged = (fwd == 1 && nx == 7) || (fwd == -1 && nx == 0);
                g->b.sq[nx][ny] = (int8_t)(kinged ? king : man);
                g->next_player = 1 - player;
                g->plane = fwd == 1 ? (side == 1 ? 9 : 8) : (side == 1 ? 7 : 6);
                g->x = x; g->y = y; g->reversible = 0;
                mask[g->plane - 6] |= 1u << sq_index(x, y);
            }
            for (int ydir = -1; ydir < 2; ydir += 2) {        /* _check_jumps (:214) */
                if (!(inb(y + ydir) && inb(x + fwd))) continue;
                if (!is_opp(b->sq[x + fwd][y + ydir], player)) continue;
                if (!(inb(y + 2 * ydir) && inb(x + 2 * fwd))) continue;
                if (b->sq[x + 2 * fwd][y + 2 * ydir] != EMPTY) continue;
                gen_t *g = &jumps[nj++];
                g->b = *b;
                g->b.sq[x][y] = EMPTY;
                g->b.sq[x + fwd][y + ydir] = EMPTY;
                int lx = x + 2 * fwd, ly = y + 2 * ydir, more = 0;
                if ((fwd == 1 && lx == 7) || (fwd == -1 && lx == 0)) {
                    g->b.sq[lx][ly] = (int8_t)king;           /* kinged: turn over (:225-227) */
                } else {
                    g->b.sq[lx][ly] = (int8_t)man;
                    more = man_can_hop(&g->b, b, lx, ly, fwd, player);
                }
                g->next_player = more ? player : 1 - player;
                g->plane = jump_plane(fwd, ydir);
                g->x = x; g->y = y; g->reversible = 0;
                mask[g->plane - 6] |= 1u << sq_index(x, y);
            }
        }
    /* kings, row-major (Checkers.py:168-196) */
    for (int x = 0; x < 8; ++x)
        for (int y = 0; y < 8; ++y) {
            if (b->sq[x][y] != king) continue;
            for (int xmove = -1; xmove < 2; xmove += 2)
                for (int ymove = -1; ymove < 2; ymove += 2) {
                    if (!(inb(x + xmove) && inb(y + ymove))) continue;
                    if (b->sq[x + xmove][y + ymove] != EMPTY) continue;
                    gen_t *g = &moves[nm++];
                    g->b = *b;
                    g->b.sq[x][y] = EMPTY;
                    g->b.sq[x + xmove][y + ymove] = (int8_t)king;
                    g->next_player = 1 - player;
                    g->plane = xmove == 1 ? (ymove == 1 ? 9 : 8) : (ymove == 1 ? 7 : 6);
                    g->x = x; g->y = y; g->reversible = 1;
                    mask[g->plane - 6] |= 1u << sq_index(x, y);
                }
            for (int ydir = -1; ydir < 2; ydir += 2)          /* _check_king_jumps (:266-267) */
                for (int f = -1; f < 2; f += 2) {
                    if (!(inb(x + f) && inb(y + ydir))) continue;
                    if (!is_opp(b->sq[x + f][y + ydir], player)) continue;
                    if (!(inb(x + 2 * f) && inb(y + 2 * ydir))) continue;
                    if (b->sq[x + 2 * f][y + 2 * ydir] != EMPTY) continue;
                    gen_t *g = &jumps[nj++];
                    g->b = *b;
                    g->b.sq[x][y] = EMPTY;
                    g->b.sq[x + f][y + ydir] = EMPTY;
                    g->b.sq[x + 2 * f][y + 2 * ydir] = (int8_t)king;
                    int more = king_can_hop(&g->b, b, x + 2 * f, y + 2 * ydir, player);
                    g->next_player = more ? player : 1 - player;
                    g->plane = jump_plane(f, ydir);
                    g->x = x; g->y = y; g->reversible = 0;
                    mask[g->plane - 6] |= 1u << sq_index(x, y);
                }
        }
    if (nj) {                                                 /* jumps are mandatory (:197-199) */
        mask[0] = mask[1] = mask[2] = mask[3] = 0;
        if (nj > CKO_MAX_CHILDREN) { fprintf(stderr, "cko: too many jumps\n"); abort(); }
        memcpy(out, jumps, (size_t)nj * sizeof(gen_t));
        return nj;
    }
    if (nm > CKO_MAX_CHILDREN) { fprintf(stderr, "cko: too many moves\n"); abort(); }
    memcpy(out, moves, (size_t)nm * sizeof(gen_t));
    return nm;
}

int cko_movegen(const cko_pos *pos, cko_pos *children, uint32_t mask[8], int *status, int *plane5) {
    board_t b;
    gen_t gen[CKO_MAX_CHILDREN];
    unpack(pos, &b);
    const int player = META_PLAYER(pos->meta), rev = META_REV(pos->meta), ply = META_PLY(pos->meta);
    int n = check_moves(&b, player, gen, mask);
    for (int i = 0; i < n; ++i) {
        pack(&gen[i].b, &children[i]);
        int action = (gen[i].plane - 6) * 64 + gen[i].x * 8 + gen[i].y;
        children[i].meta = make_meta(gen[i].next_player, gen[i].reversible ? rev + 1 : 0, action, 1, ply + 1);
    }
    /* determine_outcome (Checkers.py:306-364).  len(history) == ply+1.  The 80-state scan
     * (:332-343) breaks at the first earlier state whose piece count or men planes differ,
     * i.e. after rev+1 states; no break <=> rev >= 79. */
    int p1cnt = __builtin_popcount(pos->p1), p2cnt = __builtin_popcount(pos->p2);
    int draw_cond = 0, p5 = 0;
    if (ply + 1 >= 80) {
        if (rev >= 79) draw_cond = 1; else p5 = rev + 1;
    }
    int st;
    if (p2cnt == 0) st = CKO_P1_WINS;
    else if (p1cnt == 0) st = CKO_P2_WINS;
    else if (n == 0) st = (1 - player) == 0 ? CKO_P1_WINS : CKO_P2_WINS;
    else if (draw_cond) { st = CKO_DRAW; p5 = 80; }
    else st = CKO_ONGOING;
    if (status) *status = st;
    if (plane5) *plane5 = p5;
    return n;
}

uint64_t cko_perft(const cko_pos *pos, int depth, uint64_t *hops, uint64_t *conts) {
    cko_pos ch[CKO_MAX_CHILDREN];
    uint32_t mask[8];
    int st, p5;
    int n = cko_movegen(pos, ch, mask, &st, &p5);
    if (st != CKO_ONGOING) return 0;            /* get_legal_next_states returns [] (:91) */
    if (depth == 1) {
        for (int i = 0; i < n; ++i) {
            if (hops && META_ACTION(ch[i].meta) >= 4 * 64) ++*hops;
            if (conts && META_PLAYER(ch[i].meta) == META_PLAYER(pos->meta)) ++*conts;
        }
        return (uint64_t)n;
    }
    uint64_t t = 0;
    for (int i = 0; i < n; ++i) t += cko_perft(&ch[i], depth - 1, hops, conts);
    return t;
}

/* ------------------------------------------------------------------------------------
 * Checkers.predict glue: numpy float32 pairwise sum (numpy/_core/src/umath/loops_utils.h.src
 * @TYPE@_pairwise_sum, numpy 2.3.5 as installed here) over the 512 masked priors.
 * ---------------------------------------------------------------------------------- */
static float np_pairwise_sum_f32(const float *a, int n) {
    if (n < 8) {
        float res = 0.f;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        float r[8];
        int i;
        for (i = 0; i < 8; ++i) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum_f32(a, n2) + np_pairwise_sum_f32(a + n2, n - n2);
    }
}

static int action_legal(const uint32_t mask[8], int a) {
    int plane = a >> 6, x = (a >> 3) & 7, y = a & 7;
    if (!playable(x, y)) return 0;
    return (int)((mask[plane] >> sq_index(x, y)) & 1u);
}

void cko_mask_renorm(const float *policy_in, const uint32_t mask[8], float *policy_out) {
    float tmp[512];
    for (int a = 0; a < 512; ++a)                        /* prob_planes *= action_mask (:436) */
        tmp[a] = action_legal(mask, a) ? policy_in[a] : policy_in[a] * 0.0f;
    volatile float s = 0.0f + np_pairwise_sum_f32(tmp, 512);   /* np.sum (:437) */
    for (int a = 0; a < 512; ++a) policy_out[a] = tmp[a] / s;
}

/* ------------------------------------------------------------------------------------
 * built-in evaluators
 * ---------------------------------------------------------------------------------- */
void cko_eval_uniform_zero(const cko_pos *pos, const uint32_t *mask, int plane5, float *policy, float *value, void *ctx) {
    (void)pos; (void)mask; (void)plane5; (void)ctx;
    for (int i = 0; i < 512; ++i) policy[i] = 1.0f / 512.0f;
    *value = 0.0f;
}

static int material(const cko_pos *pos) {
    int p1 = __builtin_popcount(pos->p1) + __builtin_popcount(pos->p1 & pos->k);
    int p2 = __builtin_popcount(pos->p2) + __builtin_popcount(pos->p2 & pos->k);
    return META_PLAYER(pos->meta) == 0 ? p1 - p2 : p2 - p1;
}

void cko_eval_uniform_material(const cko_pos *pos, const uint32_t *mask, int plane5, float *policy, float *value, void *ctx) {
    (void)mask; (void)plane5; (void)ctx;
    for (int i = 0; i < 512; ++i) policy[i] = 1.0f / 512.0f;
    *value = (float)material(pos) / 32.0f;
}

static uint32_t mix32(uint32_t h) {
    h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
    return h;
}

static void eval_hash_impl(const cko_pos *pos, int plane5, float *policy, float *value, uint32_t salt);
void cko_eval_hash(const cko_pos *pos, const uint32_t *mask, int plane5, float *policy, float *value, void *ctx) {
    (void)mask; (void)ctx;
    eval_hash_impl(pos, plane5, policy, value, 0u);
}
/* ctx points at a uint32 salt (tests use the game id) so that concurrent games differ */
void cko_eval_hash_salted(const cko_pos *pos, const uint32_t *mask, int plane5, float *policy, float *value, void *ctx) {
    (void)mask;
    eval_hash_impl(pos, plane5, policy, value, *(const uint32_t *)ctx);
}
static void eval_hash_impl(const cko_pos *pos, int plane5, float *policy, float *value, uint32_t salt) {
    uint32_t h = mix32(pos->p1 ^ 0x9e3779b9u ^ (salt * 0x9E3779B1u));
    h = mix32(h ^ pos->p2);
    h = mix32(h ^ pos->k);
    h = mix32(h ^ (pos->meta & 1u) ^ ((uint32_t)plane5 << 8));
    for (int i = 0; i < 512; ++i) {
        uint32_t g = mix32(h + (uint32_t)i * 0x85ebca6bu);
        float v = (float)((g >> 8) & 0xFFFFu) + 1.0f;          /* 1..65536 exact */
        volatile float v2 = v * v;                              /* no FMA: products only */
        volatile float v3 = v2 * v;
        policy[i] = v3 * 0x1p-58f;
    }
    uint32_t g = mix32(h ^ 0xdeadbeefu);
    *value = (float)(g & 0xFFFFFFu) * 0x1p-23f - 1.0f;          /* 24-bit dyadic in [-1,1) */
}

/* ------------------------------------------------------------------------------------
 * RNG (xoshiro256**); not bit-compatible with numpy by design (SURVEY 8b RNG row)
 * ---------------------------------------------------------------------------------- */
typedef struct { uint64_t s[4]; } rng_t;
static uint64_t splitmix64(uint64_t *x) {
    uint64_t z = (*x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static void rng_seed(rng_t *r, uint64_t seed) { for (int i = 0; i < 4; ++i) r->s[i] = splitmix64(&seed); }
static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t rng_next(rng_t *r) {
    uint64_t *s = r->s, result = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return result;
}
static double rng_uniform(rng_t *r) { return (double)((rng_next(r) >> 11) + 1) * 0x1p-53; } /* (0,1] */
static double rng_normal(rng_t *r) {
    double u1 = rng_uniform(r), u2 = rng_uniform(r);
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}
static double rng_gamma(rng_t *r, double a) {       /* Marsaglia-Tsang */
    if (a == 1.0) return -log(rng_uniform(r));
    if (a < 1.0) return rng_gamma(r, a + 1.0) * pow(rng_uniform(r), 1.0 / a);
    double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (;;) {
        double x = rng_normal(r), v = 1.0 + c * x;
        if (v <= 0) continue;
        v = v * v * v;
        double u = rng_uniform(r);
        if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) return d * v;
    }
}

/* ------------------------------------------------------------------------------------
 * MCTS tree: MCTS_Node (MCTS.py:345-430) and MCTS class methods (:59-295)
 * ---------------------------------------------------------------------------------- */
static uint64_t splitmix64(uint64_t *x);

typedef struct node {
    cko_pos pos;
    uint32_t mask[8];
    int plane5, status, terminal, player;
    struct node *parent;
    struct node **children;      /* node.children: legal list reversed (MCTS.py:72-75) */
    int nchildren;
    cko_pos *unvisited;          /* unvisited_child_states, generation order */
    int n_unvisited;
    uint32_t n;                  /* _number_of_visits */
    float w;                     /* _total_reward; float32 under numpy>=2 */
    float p;                     /* _prior_prob */
} node_t;

struct cko_tree {
    node_t *root;
    int root_parent_player;      /* player to move in game_env.history[-2] */
    cko_cfg cfg;
    double tau;                  /* shared with the owning game when there is one */
    double *tau_ref;
    cko_eval_fn eval; void *ctx;
    rng_t rng;
    uint64_t playout_rng;        /* random playouts (NEURAL_NET=False) */
    uint64_t nodes, evals, rollout_count;
};

static node_t *node_new(cko_tree *t, const cko_pos *pos, node_t *parent) {
    node_t *nd = (node_t *)calloc(1, sizeof(node_t));
    cko_pos ch[CKO_MAX_CHILDREN];
    nd->pos = *pos;
    nd->player = META_PLAYER(pos->meta);
    nd->parent = parent;
    int n = cko_movegen(pos, ch, nd->mask, &nd->status, &nd->plane5);  /* MCTS.py:374 */
    if (nd->status != CKO_ONGOING) n = 0;                              /* Checkers.py:91 */
    nd->terminal = n == 0;
    if (n) {
        nd->unvisited = (cko_pos *)malloc((size_t)n * sizeof(cko_pos));
        memcpy(nd->unvisited, ch, (size_t)n * sizeof(cko_pos));
    }
    nd->n_unvisited = n;
    ++t->nodes;
    return nd;
}

static void node_free(cko_tree *t, node_t *nd, const node_t *keep) {
    if (!nd || nd == keep) return;
    for (int i = 0; i < nd->nchildren; ++i) node_free(t, nd->children[i], keep);
    free(nd->children);
    free(nd->unvisited);
    free(nd);
    --t->nodes;
}

static float node_q(const node_t *nd) { return nd->n ? nd->w / (float)nd->n : 0.0f; }

/* MCTS.determine_reward + MCTS_Node.backpropagation (MCTS.py:148-186, 419-430) */
static void backprop(cko_tree *t, node_t *leaf, int is_outcome, int outcome, float value, int sim_player) {
    for (node_t *x = leaf; x; x = x->parent) {
        int parent_player = x->parent ? x->parent->player : t->root_parent_player;
        float reward;
        if (is_outcome) {
            if (outcome == CKO_P1_WINS) reward = parent_player == 0 ? 1.f : -1.f;
            else if (outcome == CKO_P2_WINS) reward = parent_player == 1 ? 1.f : -1.f;
            else reward = 0.f;
        } else {
            reward = sim_player != parent_player ? -value : value;
        }
        x->n += 1;
        x->w += reward;
    }
    ++t->rollout_count;
}

static node_t *select_child(cko_tree *t, node_t *nd) {          /* MCTS.py:101-116 */
    const int b = nd->nchildren;
    double dir[CKO_MAX_CHILDREN], tot = 0;
    for (int j = 0; j < b; ++j) { dir[j] = rng_gamma(&t->rng, t->cfg.alpha); tot += dir[j]; }
    const double sqrt_n = pow((double)nd->n, 0.5);               /* node.n ** 0.5 */
    const float one_minus_eps = (float)(1.0 - t->cfg.epsilon);
    int best = 0; double best_u = 0;
    for (int j = 0; j < b; ++j) {
        node_t *c = nd->children[j];
        volatile float scaled = one_minus_eps * c->p;            /* float32 array op */
        double psa = (double)scaled + t->cfg.epsilon * (dir[j] / tot);
        double u = (double)node_q(c) + t->cfg.uct_c * psa * sqrt_n / (double)(1 + c->n);
        if (j == 0 || u > best_u) { best = j; best_u = u; }     /* np.argmax: first max */
    }
    return nd->children[best];
}

/* NEURAL_NET=False: child.q + 2c * (2 ln(N) / n) ** 0.5, all in float64 -- rewards are Python ints
 * there, so q is an int / int true division (MCTS.py:113-115, 389-394) */
static node_t *select_child_uct(cko_tree *t, node_t *nd) {
    const int b = nd->nchildren;
    const double ln_n = log((double)nd->n);
    int best = 0; double best_u = 0;
    for (int j = 0; j < b; ++j) {
        const node_t *c = nd->children[j];
        const double q = c->n ? (double)c->w / (double)c->n : 0.0;
        const double u = q + (2.0 * t->cfg.uct_c) * pow(2.0 * ln_n / (double)c->n, 0.5);
        if (j == 0 || u > best_u) { best = j; best_u = u; }
    }
    return nd->children[best];
}

/* MCTS_Node.simulation in that mode: default_policy plays from the node to the end of the game
 * (a terminal node is its own outcome) and the outcome string is backed up (MCTS.py:132-146, 412-417) */
static void simulate(cko_tree *t, node_t *nd) {
    int outcome = nd->status;
    if (!nd->terminal)
        outcome = t->cfg.rollout == CKO_ROLLOUT_HASH ? cko_hash_playout(&nd->pos, t->ctx ? *(const uint32_t *)t->ctx : 0u, NULL)
                                                     : cko_random_playout(&nd->pos, &t->playout_rng, NULL, 0);
    backprop(t, nd, 1, outcome, 0.f, nd->player);
}

static void tree_policy(cko_tree *t, node_t *nd) {              /* MCTS.py:59-99 */
    for (;;) {
        if (nd->n_unvisited && t->cfg.rollout) {                 /* one child per visit (:78-89) */
            if (!nd->children) nd->children = (node_t **)malloc((size_t)nd->n_unvisited * sizeof(node_t *));
            node_t *c = node_new(t, &nd->unvisited[--nd->n_unvisited], nd);     /* pop() from the end */
            nd->children[nd->nchildren++] = c;
            simulate(t, c);
            return;
        }
        if (nd->n_unvisited) {
            float policy[512], prior[512], value;
            t->eval(&nd->pos, nd->mask, nd->plane5, policy, &value, t->ctx);
            ++t->evals;
            cko_mask_renorm(policy, nd->mask, prior);
            int b = nd->n_unvisited;
            nd->children = (node_t **)malloc((size_t)b * sizeof(node_t *));
            for (int i = 0; i < b; ++i)                          /* pop() from the end */
                nd->children[i] = node_new(t, &nd->unvisited[b - 1 - i], nd);
            nd->nchildren = b;
            nd->n_unvisited = 0;
            for (int i = 0; i < b; ++i)                          /* set_prior_probs */
                nd->children[i]->p = prior[META_ACTION(nd->children[i]->pos.meta)];
            backprop(t, nd, 0, 0, value, nd->player);
            return;
        }
        if (nd->terminal) {   /* only reachable for a terminal root; the reference would fail */
            fprintf(stderr, "cko: search from a terminal root\n");
            abort();
        }
        node_t *c = t->cfg.rollout ? select_child_uct(t, nd) : select_child(t, nd);
        if (c->terminal) {                                       /* MCTS.py:93-94,145-146 */
            backprop(t, c, 1, c->status, 0.f, c->player);
            return;
        }
        nd = c;
    }
}

cko_tree *cko_tree_new(const cko_pos *root, int parent_player, const cko_cfg *cfg, cko_eval_fn eval, void *ctx) {
    cko_tree *t = (cko_tree *)calloc(1, sizeof(cko_tree));
    t->cfg = *cfg;
    t->eval = eval; t->ctx = ctx;
    t->tau = cfg->tau; t->tau_ref = &t->tau;
    rng_seed(&t->rng, cfg->seed);
    t->playout_rng = cfg->seed * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    t->root = node_new(t, root, NULL);
    t->root_parent_player = parent_player >= 0 ? parent_player : 1 - t->root->player;
    return t;
}

void cko_tree_free(cko_tree *t) {
    if (!t) return;
    node_t *r = t->root;
    while (r && r->parent) r = r->parent;
    node_free(t, r, NULL);
    free(t);
}

void cko_tree_search(cko_tree *t, int sims) {                    /* MCTS.py:210-224 */
    t->rollout_count = 0;
    while (t->rollout_count < (uint64_t)sims) tree_policy(t, t->root);
}

int cko_tree_root_children(const cko_tree *t, cko_pos *pos, uint32_t *n, float *w, float *p, int32_t *terminal) {
    const node_t *r = t->root;
    for (int i = 0; i < r->nchildren; ++i) {
        if (pos) pos[i] = r->children[i]->pos;
        if (n) n[i] = r->children[i]->n;
        if (w) w[i] = r->children[i]->w;
        if (p) p[i] = r->children[i]->p;
        if (terminal) terminal[i] = r->children[i]->terminal ? r->children[i]->status : 0;
    }
    return r->nchildren;
}

void cko_tree_root_stats(const cko_tree *t, uint32_t *n, float *w) { *n = t->root->n; *w = t->root->w; }
uint64_t cko_tree_node_count(const cko_tree *t) { return t->nodes; }
uint64_t cko_tree_nn_evals(const cko_tree *t) { return t->evals; }

int cko_tree_best_child(cko_tree *t, int move_count) {           /* MCTS.py:226-248 */
    const node_t *r = t->root;
    double tau = *t->tau_ref;
    if (!t->cfg.training || tau <= 0) {
        int best = 0;
        for (int i = 1; i < r->nchildren; ++i)
            if (r->children[i]->n > r->children[best]->n) best = i;
        return best;
    }
    double ex[CKO_MAX_CHILDREN], total = 0;
    for (int i = 0; i < r->nchildren; ++i) { ex[i] = pow((double)r->children[i]->n, 1.0 / tau); total += ex[i]; }
    if (move_count > t->cfg.tau_decay_delay) {
        tau -= t->cfg.tau_decay;
        if (fabs(tau) <= 1e-8) tau = 0;                          /* np.isclose(tau, 0) */
        *t->tau_ref = tau;
    }
    double u = rng_uniform(&t->rng) * total, acc = 0;
    for (int i = 0; i < r->nchildren; ++i) { acc += ex[i]; if (u <= acc) return i; }
    return r->nchildren - 1;
}

/* ------------------------------------------------------------------------------------
 * game loop: training_pipeline._generate_data (:334-419) / _start_tournament (:505-559)
 * ---------------------------------------------------------------------------------- */
struct cko_game {
    cko_cfg cfg;
    cko_tree *tree[2];           /* root_node1 / root_node2 trees */
    node_t *best_child[2];       /* best_child1 / best_child2 */
    cko_eval_fn eval[2]; void *ctx[2];
    cko_pos *history; int hist_len, hist_cap;
    int done, outcome, move_count, terminated;
    double tau;
    cko_record *rec; int nrec, rec_cap;
    uint64_t sims, evals, misses;
};

static void hist_push(cko_game *g, const cko_pos *p) {
    if (g->hist_len == g->hist_cap) {
        g->hist_cap = g->hist_cap ? 2 * g->hist_cap : 256;
        g->history = (cko_pos *)realloc(g->history, (size_t)g->hist_cap * sizeof(cko_pos));
    }
    g->history[g->hist_len++] = *p;
}

static cko_record *rec_push(cko_game *g) {
    if (g->nrec == g->rec_cap) {
        g->rec_cap = g->rec_cap ? 2 * g->rec_cap : 256;
        g->rec = (cko_record *)realloc(g->rec, (size_t)g->rec_cap * sizeof(cko_record));
    }
    cko_record *r = &g->rec[g->nrec++];
    memset(r, 0, sizeof(*r));
    return r;
}

cko_game *cko_game_new(const cko_cfg *cfg, cko_eval_fn e1, void *c1, cko_eval_fn e2, void *c2) {
    cko_game *g = (cko_game *)calloc(1, sizeof(cko_game));
    g->cfg = *cfg;
    g->eval[0] = e1; g->ctx[0] = c1;
    g->eval[1] = e2 ? e2 : e1; g->ctx[1] = e2 ? c2 : c1;
    g->tau = cfg->tau;
    cko_pos start;
    cko_start_position(&start);
    hist_push(g, &start);
    return g;
}

void cko_game_free(cko_game *g) {
    if (!g) return;
    cko_tree_free(g->tree[0]);
    cko_tree_free(g->tree[1]);
    free(g->history); free(g->rec); free(g);
}

static int same_state(const cko_pos *a, const cko_pos *b) {
    return a->p1 == b->p1 && a->p2 == b->p2 && a->k == b->k && ((a->meta ^ b->meta) & 1u) == 0;
}

/* MCTS.new_root_node (MCTS.py:250-295) on tree `t`, starting from old_root = best_child */
static void new_root(cko_game *g, cko_tree *t, node_t *old_root) {
    const cko_pos *h = g->history; const int L = g->hist_len;
    int counter = 1, idx = L - 3;
    while (idx >= 0 && META_PLAYER(h[L - 2].meta) == META_PLAYER(h[idx].meta)) { ++counter; --idx; }
    node_t *nr = old_root;
    for (int i = L - counter; i < L; ++i)
        for (int c = 0; c < nr->nchildren; ++c)
            if (same_state(&nr->children[c]->pos, &h[i])) { nr = nr->children[c]; break; }
    node_t *top = t->root;
    if (same_state(&nr->pos, &h[L - 1])) {
        /* detach: new_root.parent = None; everything else becomes garbage */
        if (nr->parent) {
            node_t *par = nr->parent;
            for (int c = 0; c < par->nchildren; ++c) if (par->children[c] == nr) par->children[c] = NULL;
            nr->parent = NULL;
            node_free(t, top, NULL);
        }
        t->root = nr;
    } else {
        /* the reference raises here (MCTS.py:292); the dead code after it shows the
         * intent: a fresh root.  Count the event. */
        ++g->misses;
        node_free(t, top, NULL);
        t->root = node_new(t, &h[L - 1], NULL);
    }
}

int cko_game_play_ply(cko_game *g) {
    if (g->done) return 0;
    const cko_pos *cur = &g->history[g->hist_len - 1];
    const int pl = META_PLAYER(cur->meta);
    int parent_player;
    if (g->hist_len >= 2) parent_player = META_PLAYER(g->history[g->hist_len - 2].meta);
    else parent_player = 1 - pl;
    cko_cfg tcfg = g->cfg;
    tcfg.seed = g->cfg.seed * 2 + (uint64_t)pl + 1;
    if (!g->tree[pl]) {
        /* root_node1 = MCTS_Node(initial_state) / root_node2 = MCTS_Node(state, initial_state=..) */
        g->tree[pl] = cko_tree_new(cur, parent_player, &tcfg, g->eval[pl], g->ctx[pl]);
        g->tree[pl]->tau_ref = &g->tau;
    } else {
        new_root(g, g->tree[pl], g->best_child[pl]);
    }
    cko_tree *t = g->tree[pl];
    t->root_parent_player = parent_player;
    uint64_t ev0 = t->evals;
    cko_tree_search(t, g->cfg.budget);
    g->sims += t->rollout_count;
    g->evals += t->evals - ev0;
    node_t *root = t->root;
    int bi = cko_tree_best_child(t, g->move_count);
    node_t *best = root->children[bi];
    g->best_child[pl] = best;
    /* game_env.step(best_child.state) */
    hist_push(g, &best->pos);
    g->move_count += 1;
    if (best->status != CKO_ONGOING) { g->done = 1; g->outcome = best->status; }
    /* record (training_pipeline.py:364-369) */
    cko_record *r = rec_push(g);
    r->pos = root->pos;
    memcpy(r->mask, root->mask, sizeof(r->mask));
    r->plane5 = root->plane5;
    r->n_children = root->nchildren;
    for (int i = 0; i < root->nchildren; ++i) {
        r->action[i] = (uint16_t)META_ACTION(root->children[i]->pos.meta);
        r->visits[i] = root->children[i]->n;
    }
    float q = node_q(root);
    r->q = parent_player != root->player ? -q : q;
    r->root_n = root->n; r->root_w = root->w;
    r->chosen = META_ACTION(best->pos.meta);
    /* TERMINATE_CNT adjudication (:387-405) */
    if (!g->done && g->cfg.terminate_cnt > 0 && g->move_count >= g->cfg.terminate_cnt) {
        g->terminated = 1; g->done = 1;
        const cko_pos *s = &best->pos;
        int p1 = __builtin_popcount(s->p1), p2 = __builtin_popcount(s->p2);
        int k1 = __builtin_popcount(s->p1 & s->k), k2 = __builtin_popcount(s->p2 & s->k);
        if (p1 > p2) g->outcome = CKO_P1_WINS;
        else if (p1 < p2) g->outcome = CKO_P2_WINS;
        else if (k1 > k2) g->outcome = CKO_P1_WINS;
        else if (k1 < k2) g->outcome = CKO_P2_WINS;
        else g->outcome = CKO_DRAW;
    }
    if (g->done && !g->terminated) {                              /* terminal record (:406-409) */
        cko_record *tr = rec_push(g);
        tr->pos = best->pos;
        memcpy(tr->mask, best->mask, sizeof(tr->mask));
        tr->plane5 = best->plane5;
        tr->q = g->outcome == CKO_DRAW ? 0.f : -1.f;
        tr->chosen = -1;
    }
    if (g->done)                                                  /* _add_rewards (:439-455) */
        for (int i = 0; i < g->nrec; ++i) {
            int player = META_PLAYER(g->rec[i].pos.meta);
            if (g->outcome == CKO_P1_WINS) g->rec[i].z = player == 0 ? 1 : -1;
            else if (g->outcome == CKO_P2_WINS) g->rec[i].z = player == 1 ? 1 : -1;
            else g->rec[i].z = 0;
        }
    return !g->done;
}

void cko_game_state(const cko_game *g, cko_pos *pos) { *pos = g->history[g->hist_len - 1]; }
int cko_game_outcome(const cko_game *g) { return g->outcome; }
int cko_game_move_count(const cko_game *g) { return g->move_count; }
int cko_game_terminated(const cko_game *g) { return g->terminated; }
int cko_game_num_records(const cko_game *g) { return g->nrec; }
void cko_game_get_records(const cko_game *g, cko_record *out) { memcpy(out, g->rec, (size_t)g->nrec * sizeof(cko_record)); }
uint64_t cko_game_total_sims(const cko_game *g) { return g->sims; }
uint64_t cko_game_nn_evals(const cko_game *g) { return g->evals; }
uint64_t cko_game_reroot_misses(const cko_game *g) { return g->misses; }

/* ------------------------------------------------------------------------------------
 * random playout: MCTS.default_policy, non-NN branch (MCTS.py:132-143)
 * ---------------------------------------------------------------------------------- */
int cko_random_playout(const cko_pos *pos, uint64_t *rng_state, int *plies, int max_plies) {
    cko_pos cur = *pos, ch[CKO_MAX_CHILDREN];
    uint32_t mask[8];
    int st, p5, n, k = 0;
    for (;;) {
        n = cko_movegen(&cur, ch, mask, &st, &p5);
        if (st != CKO_ONGOING) break;
        if (max_plies > 0 && k >= max_plies) break;
        uint64_t r = splitmix64(rng_state);
        cur = ch[(int)(r % (uint64_t)n)];
        ++k;
    }
    if (plies) *plies = k;
    return st;
}

/* deterministic stand-in for np.random.randint(0, n) inside a playout: a hash of the position the
 * move is chosen from (same mixing as the hash evaluator); salt = game tag, 0 for the golden vectors */
static uint32_t hash_choice(const cko_pos *pos, uint32_t n, uint32_t salt) {
    uint32_t h = mix32(pos->p1 ^ 0x9e3779b9u ^ (salt * 0x9E3779B1u));
    h = mix32(h ^ pos->p2);
    h = mix32(h ^ pos->k);
    h = mix32(h ^ (pos->meta & 1u));
    return h % n;
}
uint32_t cko_hash_choice(const cko_pos *pos, uint32_t n) { return hash_choice(pos, n, 0u); }

int cko_hash_playout(const cko_pos *pos, uint32_t salt, int *plies) {
    cko_pos cur = *pos, ch[CKO_MAX_CHILDREN];
    uint32_t mask[8];
    int st, p5, n, k = 0;
    for (;;) {
        n = cko_movegen(&cur, ch, mask, &st, &p5);
        if (st != CKO_ONGOING) break;
        cur = ch[hash_choice(&cur, (uint32_t)n, salt)];
        ++k;
    }
    if (plies) *plies = k;
    return st;
}
