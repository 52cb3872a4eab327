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

// Which of the four diagonal steps from s are a legal plain move / a legal single hop,
// ignoring the piece kind.  `empty_land` is the occupancy complement used for landing
// squares (callers pass the parent's stale board for continuation tests, Checkers.py:219,272).
CK_HD void sq_dirs(uint32_t opp, uint32_t empty, uint32_t empty_land, int s, uint32_t &mv, uint32_t &jp) {
    const int x = sq_x(s), y = sq_y(s);
    mv = 0; jp = 0;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        const int nx = x + dir_dx(d), ny = y + dir_dy(d);
        if (!inb(nx) || !inb(ny)) continue;
        const uint32_t nb = 1u << sq_of(nx, ny);
        if (empty & nb) { mv |= 1u << d; continue; }
        if (!(opp & nb)) continue;
        const int lx = nx + dir_dx(d), ly = ny + dir_dy(d);
        if (!inb(lx) || !inb(ly)) continue;
        if (empty_land & (1u << sq_of(lx, ly))) jp |= 1u << d;
    }
}

// directions a piece may use: men only forward (player1 moves +x, Checkers.py:120,131-132)
CK_HD uint32_t dir_allow(bool king, int player) { return king ? 0xFu : (player == 0 ? 0xCu : 0x3u); }

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
        uint32_t mv, jp;
        sq_dirs(opp, ~(own | opp), sd.empty, sq_of(tx, ty), mv, jp);
        if (jp & dir_allow(king, sd.player)) next_player = sd.player;
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

// per-square legal direction sets of the side to move
CK_HD void piece_dirs(const Side &sd, int s, uint32_t &mv, uint32_t &jp) {
    sq_dirs(sd.opp, sd.empty, sd.empty, s, mv, jp);
    const uint32_t allow = dir_allow((sd.kings >> s) & 1u, sd.player);
    mv &= allow; jp &= allow;
}

CK_HD bool any_jump(const Side &sd) {
    for (uint32_t rem = sd.own; rem; rem &= rem - 1) {
        uint32_t mv, jp;
        piece_dirs(sd, ffs32(rem), mv, jp);
        if (jp) return true;
    }
    return false;
}
CK_HD bool any_legal(const Side &sd) {
    for (uint32_t rem = sd.own; rem; rem &= rem - 1) {
        uint32_t mv, jp;
        piece_dirs(sd, ffs32(rem), mv, jp);
        if (mv | jp) return true;
    }
    return false;
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
    CK_HD bool want(int) const { return false; }
    CK_HD void put(int, const ck_pos &) const {}
};
struct ArraySink {
    ck_pos *dst; int cap;
    CK_HD bool want(int n) const { return n < cap; }
    CK_HD void put(int n, const ck_pos &c) const { dst[n] = c; }
};
struct PickSink {          // keep only successor `target` (random playouts)
    ck_pos *dst; int target;
    CK_HD bool want(int n) const { return n == target; }
    CK_HD void put(int, const ck_pos &c) const { *dst = c; }
};

// Full generation in the reference's list order.  Returns the raw _check_moves count (also
// for finished games, as the reference does).
template <typename Sink>
CK_HD int gen_moves(const ck_pos &p, const Sink &sink, uint32_t mask[8]) {
    const Side sd = side_of(p);
    const bool jump = any_jump(sd);
#pragma unroll
    for (int i = 0; i < 8; ++i) mask[i] = 0;
    int n = 0;
    for (int pass = 0; pass < 2; ++pass) {               // men first, then kings (:111-124,168-196,197-199)
        const bool king = pass == 1;
        for (uint32_t rem = king ? (sd.own & sd.kings) : (sd.own & ~sd.kings); rem; rem &= rem - 1) {
            const int s = ffs32(rem);
            uint32_t mv, jp;
            piece_dirs(sd, s, mv, jp);
            const uint32_t use = jump ? jp : mv;
            const int nd = king ? 4 : 2;
            for (int i = 0; i < nd; ++i) {
                const int d = order_dir(king, jump, sd.player, i);
                if (!((use >> d) & 1u)) continue;
                mask[(jump ? 4 : 0) + d] |= 1u << s;
                if (sink.want(n)) sink.put(n, make_child(p, sd, s, d, jump));
                ++n;
            }
        }
    }
    return n;
}

CK_HD ck_pos start_position() {
    ck_pos p;
    p.p1 = 0x00000FFFu; p.p2 = 0xFFF00000u; p.k = 0; p.meta = 0;
    return p;
}

}  // namespace ck
