// ck_rules.cuh -- 32-bit bitboard Checkers rules, thread-level, host/device.
//
// B200-native restatement of Checkers._check_moves / _check_jumps / _check_king_jumps
// (reference Checkers.py:94-304) and determine_outcome (:306-364).  One thread handles one
// position; every kernel in this library (movegen sweep, playouts, tree expansion) calls
// these functions.  The header also compiles for the host so that the rule logic can be
// checked on a machine without a GPU (tests/host_rules/); the shipped library only
// instantiates the device side.
#pragma once
#include <stdint.h>
#include "../../include/ckb200.h"

#if defined(__CUDACC__)
#define CK_HD __host__ __device__ __forceinline__
#else
#define CK_HD inline
#endif

namespace ck {

CK_HD int popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
CK_HD int ffs32(uint32_t v) {      // index of lowest set bit, v != 0
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}

// ---- meta word ---------------------------------------------------------------------
CK_HD int meta_player(uint32_t m) { return (int)(m & 1u); }
CK_HD int meta_rev(uint32_t m) { return (int)((m >> 1) & 0x7Fu); }
CK_HD int meta_action(uint32_t m) { return (int)((m >> 8) & 0x1FFu); }
CK_HD int meta_has_action(uint32_t m) { return (int)((m >> 17) & 1u); }
CK_HD int meta_ply(uint32_t m) { return (int)((m >> 18) & 0x3FFFu); }
CK_HD uint32_t make_meta(int player, int rev, int action, int has_action, int ply) {
    if (rev > 127) rev = 127;
    if (ply > 0x3FFF) ply = 0x3FFF;
    return (uint32_t)player | ((uint32_t)rev << 1) | ((uint32_t)action << 8) |
           ((uint32_t)has_action << 17) | ((uint32_t)ply << 18);
}

// ---- geometry ------------------------------------------------------------------------
// square s = 4*x + c, c = y>>1; rows with even x use odd columns (y = 2c+1).
CK_HD int sq_x(int s) { return s >> 2; }
CK_HD int sq_y(int s) { return ((s & 3) << 1) | (((s >> 2) & 1) ^ 1); }
CK_HD int sq_of(int x, int y) { return (x << 2) | (y >> 1); }
CK_HD bool inb(int v) { return (unsigned)v < 8u; }

// direction index d: bit1 = (dx > 0), bit0 = (dy > 0)  ->  planes 6+d (move) / 10+d (jump)
// 0 = UL(-1,-1)  1 = UR(-1,+1)  2 = BL(+1,-1)  3 = BR(+1,+1)      (Checkers.py:45-46)
CK_HD int dir_dx(int d) { return (d & 2) ? 1 : -1; }
CK_HD int dir_dy(int d) { return (d & 1) ? 1 : -1; }

// ---- whole-board direction sets by shifts ---------------------------------------------------
// Rows with even x hold columns 1,3,5,7, rows with odd x columns 0,2,4,6, so the diagonal
// neighbours of square s are s-4 / s-3 (even row, up), s-5 / s-4 (odd row, up), s+4 / s+5 (even
// row, down), s+3 / s+4 (odd row, down); two steps along a diagonal are always s -+ 9 / -+ 7.
constexpr uint32_t kEvenRows = 0x0F0F0F0Fu, kOddRows = 0xF0F0F0F0u;
constexpr uint32_t kNotCol0 = ~0x11111111u, kNotCol3 = ~0x88888888u;   // c = s & 3

// squares whose neighbour in direction d belongs to T
CK_HD uint32_t nb_in(int d, uint32_t T) {
    switch (d) {
        case 0: return ((T << 4) & kEvenRows) | ((T << 5) & kOddRows & kNotCol0);
        case 1: return ((T << 3) & kEvenRows & kNotCol3) | ((T << 4) & kOddRows);
        case 2: return ((T >> 4) & kEvenRows) | ((T >> 3) & kOddRows & kNotCol0);
        default: return ((T >> 5) & kEvenRows & kNotCol3) | ((T >> 4) & kOddRows);
    }
}
// squares whose landing square two steps away in direction d belongs to T
CK_HD uint32_t land_in(int d, uint32_t T) {
    switch (d) {
        case 0: return (T << 9) & kNotCol0;
        case 1: return (T << 7) & kNotCol3;
        case 2: return (T >> 7) & kNotCol0;
        default: return (T >> 9) & kNotCol3;
    }
}

// generation order of the directions inside one piece:
//   man moves  : y+1 then y-1 (Checkers.py:125,145);  man jumps : ydir=-1 then +1 (:214)
//   king moves : UL,UR,BL,BR (:169-170);               king jumps: UL,BL,UR,BR (:266-267)
CK_HD int order_dir(bool king, bool jump, int player, int i) {
    if (king) return jump ? ((i & 1) << 1 | (i >> 1)) : i;
    const int base = player == 0 ? 2 : 0;
    return jump ? base + i : base + (1 - i);
}

struct Side { uint32_t own, opp, kings, empty; int player; };

CK_HD Side side_of(const ck_pos &p) {
    Side s;
    s.player = meta_player(p.meta);
    s.own = s.player == 0 ? p.p1 : p.p2;
    s.opp = s.player == 0 ? p.p2 : p.p1;
    s.kings = p.k;
    s.empty = ~(p.p1 | p.p2);
    return s;
}

// pieces of `own` that may use direction d: kings any, men only forward (player1 moves +x,
// Checkers.py:120,131-132)
CK_HD uint32_t movers(uint32_t own, uint32_t kings, int player, int d) {
    const bool forward = player == 0 ? (d >= 2) : (d < 2);
    return forward ? own : (own & kings);
}

// source squares with a legal plain move (mv) / single hop (jp) per direction.  `empty_land` is
// the emptiness used for landing squares (the parent's stale board in continuation tests,
// Checkers.py:219,272).
struct DirSets { uint32_t mv[4], jp[4]; };
CK_HD DirSets dir_sets(uint32_t own, uint32_t opp, uint32_t kings, uint32_t empty, uint32_t empty_land, int player) {
    DirSets D;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        const uint32_t who = movers(own, kings, player, d);
        D.mv[d] = who & nb_in(d, empty);
        D.jp[d] = who & nb_in(d, opp) & land_in(d, empty_land);
    }
    return D;
}

// successor for moving the piece on s in direction d (jump: a single hop)
CK_HD ck_pos make_child(const ck_pos &par, const Side &sd, int s, int d, bool jump) {
    const int x = sq_x(s), y = sq_y(s);
    const int dx = dir_dx(d), dy = dir_dy(d);
    const bool king = (sd.kings >> s) & 1u;
    const int step = jump ? 2 : 1;
    const int tx = x + step * dx, ty = y + step * dy;
    const uint32_t sb = 1u << s, tb = 1u << sq_of(tx, ty);
    uint32_t own = (sd.own & ~sb) | tb, opp = sd.opp, kings = sd.kings & ~sb;
    if (jump) {
        const uint32_t mb = 1u << sq_of(x + dx, y + dy);
        opp &= ~mb;
        kings &= ~mb;
    }
    const bool kinged = !king && (sd.player == 0 ? tx == 7 : tx == 0);
    if (king || kinged) kings |= tb;
    int next_player = 1 - sd.player;
    if (jump && !kinged) {
        // the hopping piece keeps the move iff it can hop again from its landing square:
        // opponent pieces from the child, landing emptiness from the PARENT's stale board
        // (Checkers.py:225-237, 279-281)
        uint32_t again = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) again |= movers(tb, king ? tb : 0u, sd.player, e) & nb_in(e, opp) & land_in(e, sd.empty);
        if (again) next_player = sd.player;
    }
    ck_pos c;
    c.p1 = sd.player == 0 ? own : opp;
    c.p2 = sd.player == 0 ? opp : own;
    c.k = kings;
    const int action = ((jump ? 4 : 0) + d) * 64 + x * 8 + y;
    const int rev = (king && !jump) ? meta_rev(par.meta) + 1 : 0;
    c.meta = make_meta(next_player, rev, action, 1, meta_ply(par.meta) + 1);
    return c;
}

// squares from which a hop in direction e is geometrically possible on this board: an opponent piece next
// door and an empty landing square behind it (whoever stands on the square)
CK_HD void hop_sets(const Side &sd, uint32_t J[4]) {
#pragma unroll
    for (int e = 0; e < 4; ++e) J[e] = nb_in(e, sd.opp) & land_in(e, sd.empty);
}

// make_child with the per-position work hoisted out: J = hop_sets(parent).  Same result as make_child;
// index arithmetic on the 32-square board instead of (x, y) coordinates, and the "can the hopping piece hop
// again" test (stale parent emptiness, child opponents, Checkers.py:225-237, 279-281) reduced to bit tests:
// removing the captured piece only changes the hop set of the direction that points back at it.
CK_HD ck_pos make_child_fast(const ck_pos &par, const Side &sd, const uint32_t J[4], int s, int d, bool jump) {
    const int x = s >> 2, odd = x & 1;
    const int md = (d == 0 ? -4 : d == 1 ? -3 : d == 2 ? 4 : 5) - odd;      // one diagonal step from an even / odd row
    const int jd = d == 0 ? -9 : d == 1 ? -7 : d == 2 ? 7 : 9;              // two steps
    const int t = s + (jump ? jd : md);
    const uint32_t sb = 1u << s, tb = 1u << t;
    const bool king = (sd.kings >> s) & 1u;
    uint32_t own = (sd.own ^ sb) | tb, opp = sd.opp, kings = sd.kings & ~sb;
    if (jump) {
        const uint32_t mb = 1u << (s + md);
        opp &= ~mb;
        kings &= ~mb;
    }
    const int tx = t >> 2;
    const bool kinged = !king && (sd.player == 0 ? tx == 7 : tx == 0);
    if (king || kinged) kings |= tb;
    int next_player = 1 - sd.player;
    if (jump && !kinged) {
        uint32_t again = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const bool allowed = king || (sd.player == 0 ? e >= 2 : e < 2);
            if (allowed && e != 3 - d) again |= (J[e] >> t) & 1u;
        }
        if (again) next_player = sd.player;
    }
    ck_pos c;
    c.p1 = sd.player == 0 ? own : opp;
    c.p2 = sd.player == 0 ? opp : own;
    c.k = kings;
    const int action = ((jump ? 4 : 0) + d) * 64 + 2 * s + (odd ^ 1);       // x*8 + y = 2s + (row even)
    const int rev = (king && !jump) ? meta_rev(par.meta) + 1 : 0;
    c.meta = make_meta(next_player, rev, action, 1, meta_ply(par.meta) + 1);
    return c;
}

CK_HD bool any_legal(const Side &sd) {
    const DirSets D = dir_sets(sd.own, sd.opp, sd.kings, sd.empty, sd.empty, sd.player);
    return (D.mv[0] | D.mv[1] | D.mv[2] | D.mv[3] | D.jp[0] | D.jp[1] | D.jp[2] | D.jp[3]) != 0;
}

// determine_outcome (Checkers.py:306-364) given whether the side to move has a legal move.
// plane5 receives the numerator of state[5]: 0 before ply 79, rev+1 afterwards, 80 on a draw.
CK_HD int outcome_of(const ck_pos &p, bool has_move, int *plane5) {
    const int player = meta_player(p.meta), rev = meta_rev(p.meta), len = meta_ply(p.meta) + 1;
    int p5 = 0;
    bool draw = false;
    if (len >= 80) { if (rev >= 79) draw = true; else p5 = rev + 1; }
    int st;
    if (p.p2 == 0) st = CK_P1_WINS;
    else if (p.p1 == 0) st = CK_P2_WINS;
    else if (!has_move) st = player == 0 ? CK_P2_WINS : CK_P1_WINS;
    else if (draw) { st = CK_DRAW; p5 = 80; }
    else st = CK_ONGOING;
    if (plane5) *plane5 = p5;
    return st;
}

// status of a position without materialising its children (MCTS_Node.__init__ only needs
// `terminal`, MCTS.py:374-375)
CK_HD int status_of(const ck_pos &p, int *plane5) { return outcome_of(p, any_legal(side_of(p)), plane5); }

// child sinks for gen_moves: where (and whether) successor n is materialised
struct NullSink {
    CK_HD bool want_any() const { return false; }
    CK_HD bool want(int) const { return false; }
    CK_HD void put(int, const ck_pos &) const {}
};
struct ArraySink {
    ck_pos *dst; int cap;
    CK_HD bool want_any() const { return true; }
    CK_HD bool want(int n) const { return n < cap; }
    CK_HD void put(int n, const ck_pos &c) const { dst[n] = c; }
};
struct PickSink {          // keep only successor `target` (random playouts)
    ck_pos *dst; int target;
    CK_HD bool want_any() const { return true; }
    CK_HD bool want(int n) const { return n == target; }
    CK_HD void put(int, const ck_pos &c) const { *dst = c; }
};

// Full generation in the reference's list order.  Returns the raw _check_moves count (also
// for finished games, as the reference does).  The legal-action planes are exactly the
// per-direction source sets; jumps are mandatory and clear the move planes (:197-199).
template <typename Sink>
CK_HD int gen_moves(const ck_pos &p, const Sink &sink, uint32_t mask[8]) {
    const Side sd = side_of(p);
    const DirSets D = dir_sets(sd.own, sd.opp, sd.kings, sd.empty, sd.empty, sd.player);
    const bool jump = (D.jp[0] | D.jp[1] | D.jp[2] | D.jp[3]) != 0;
    uint32_t use[4];
    int total = 0;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        use[d] = jump ? D.jp[d] : D.mv[d];
        mask[d] = jump ? 0u : use[d];
        mask[4 + d] = jump ? use[d] : 0u;
        total += popc32(use[d]);
    }
    if (!sink.want_any()) return total;
    const uint32_t any = use[0] | use[1] | use[2] | use[3];
    int n = 0;
    for (int pass = 0; pass < 2; ++pass) {               // men first, then kings (:111-124,168-196)
        const bool king = pass == 1;
        for (uint32_t rem = any & (king ? sd.kings : ~sd.kings); rem; rem &= rem - 1) {
            const int s = ffs32(rem);
            const int nd = king ? 4 : 2;
            for (int i = 0; i < nd; ++i) {
                const int d = order_dir(king, jump, sd.player, i);
                if (!((use[d] >> s) & 1u)) continue;
                if (sink.want(n)) sink.put(n, make_child(p, sd, s, d, jump));
                ++n;
            }
        }
    }
    return total;
}

// Random access into the list gen_moves produces: the k-th successor (0 <= k < count) without
// materialising the others.  `use` are the per-direction source sets (= the legal-action planes:
// mask[0..3] for plain moves, mask[4..7] for jumps).  The list order is men in ascending square order
// with their two forward directions in order_dir order, then kings likewise with four directions, so
// the square is found by a five-step search over prefix popcounts.  Used by the packed movegen kernel,
// where one lane per SUCCESSOR (instead of per position) keeps warps converged and stores coalesced.
// (source square, direction) of the k-th entry of that list
CK_HD void kth_move(const Side &sd, const uint32_t use[4], bool jump, int k, int *s_out, int *d_out) {
    // One converged path for men and kings (a warp of playouts or of successors mixes both all the time) and
    // constant indices into use[] only: the ordered direction slots of the phase the k-th entry falls in are
    // selected up front -- men: their two forward directions (order_dir), kings: UL,UR,BL,BR or UL,BL,UR,BR.
    const uint32_t kings = sd.kings;
    const bool p0 = sd.player == 0;
    const uint32_t f_lo = p0 ? use[2] : use[0], f_hi = p0 ? use[3] : use[1];    // forward-left / forward-right
    const int base = p0 ? 2 : 0;
    const uint32_t ua = (jump ? f_lo : f_hi) & ~kings, ub = (jump ? f_hi : f_lo) & ~kings;
    const int men_total = popc32(ua) + popc32(ub);
    const bool kp = k >= men_total;                                             // the entry is a king's
    k -= kp ? men_total : 0;
    const uint32_t m0 = kp ? (use[0] & kings) : ua;
    const uint32_t m1 = kp ? ((jump ? use[2] : use[1]) & kings) : ub;
    const uint32_t m2 = kp ? ((jump ? use[1] : use[2]) & kings) : 0u;
    const uint32_t m3 = kp ? (use[3] & kings) : 0u;
    const int e0 = kp ? 0 : (jump ? base : base + 1), e1 = kp ? (jump ? 2 : 1) : (jump ? base + 1 : base);
    const int e2 = jump ? 1 : 2;
    int s = 0;
#pragma unroll
    for (int step = 16; step; step >>= 1) {
        const uint32_t below = (1u << (s + step)) - 1u;
        if (popc32(m0 & below) + popc32(m1 & below) + popc32(m2 & below) + popc32(m3 & below) <= k) s += step;
    }
    const uint32_t below = (1u << s) - 1u;
    const int r = k - popc32(m0 & below) - popc32(m1 & below) - popc32(m2 & below) - popc32(m3 & below);
    // the r-th of this piece's directions that are set, in slot order
    const int b0 = (int)((m0 >> s) & 1u), b1 = (int)((m1 >> s) & 1u), b2 = (int)((m2 >> s) & 1u);
    int d = 3;
    if (b2 && b0 + b1 == r) d = e2;
    if (b1 && b0 == r) d = e1;
    if (b0 && r == 0) d = e0;
    *s_out = s; *d_out = d;
}
CK_HD ck_pos kth_successor(const ck_pos &p, const uint32_t use[4], bool jump, int k) {
    const Side sd = side_of(p);
    int s, d;
    kth_move(sd, use, jump, k, &s, &d);
    return make_child(p, sd, s, d, jump);
}

// ---- playouts (MCTS.default_policy without a net, MCTS.py:132-143) ------------------------------------
// One generation pass per ply: `choose(position, ply, n_legal)` names the successor in generation order
// and it is built from the legal-action planes.  Returns the CK_* outcome; *plies = plies played.
constexpr int kPlayMoved = -1;
// one ply: kPlayMoved after moving `cur` on, otherwise the playout is over and the CK_* outcome is returned
// (CK_ONGOING when it was cut off at max_plies)
template <typename Choose>
CK_HD int play_step(ck_pos &cur, int k, int max_plies, const Choose &choose) {
    uint32_t mask[8];
    const int cnt = gen_moves(cur, NullSink{}, mask);
    const int st = outcome_of(cur, cnt > 0, nullptr);
    if (st != CK_ONGOING || (max_plies > 0 && k >= max_plies)) return st;
    const int pick = choose(cur, k, cnt);
    const bool jump = (mask[4] | mask[5] | mask[6] | mask[7]) != 0;
    const Side sd = side_of(cur);
    uint32_t hop[4];
    hop_sets(sd, hop);
    int ms, md;
    const uint32_t use[4] = {jump ? mask[4] : mask[0], jump ? mask[5] : mask[1], jump ? mask[6] : mask[2], jump ? mask[7] : mask[3]};
    kth_move(sd, use, jump, pick, &ms, &md);
    cur = make_child_fast(cur, sd, hop, ms, md, jump);
    return kPlayMoved;
}
template <typename Choose>
CK_HD int play_out(ck_pos cur, int max_plies, const Choose &choose, int *plies) {
    int k = 0, st;
    while ((st = play_step(cur, k, max_plies, choose)) == kPlayMoved) ++k;
    if (plies) *plies = k;
    return st;
}

// deterministic stand-in for the uniform choice (parity tests; twin of the oracle's cko_hash_choice)
CK_HD uint32_t mix32(uint32_t h) {
    h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
    return h;
}
struct HashChoice {
    uint32_t salt;
    CK_HD int operator()(const ck_pos &p, int, int cnt) const {
        uint32_t h = mix32(p.p1 ^ 0x9e3779b9u ^ (salt * 0x9E3779B1u));
        h = mix32(h ^ p.p2);
        h = mix32(h ^ p.k);
        h = mix32(h ^ (p.meta & 1u));
        return (int)(h % (uint32_t)cnt);
    }
};

CK_HD ck_pos start_position() {
    ck_pos p;
    p.p1 = 0x00000FFFu; p.p2 = 0xFFF00000u; p.k = 0; p.meta = 0;
    return p;
}

}  // namespace ck
