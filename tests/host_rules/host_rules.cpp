// TEST INFRASTRUCTURE ONLY.  Compiles the product's thread-level rule functions
// (checkers-mcts_b200/csrc/ck_rules.cuh, host/device header) for the HOST so that the
// bitboard logic can be checked against the oracle on a machine without a GPU.  Nothing in
// the product links or loads this library; the shipped libckb200.so has no CPU compute path.
#include "../../checkers-mcts_b200/csrc/ck_rules.cuh"

extern "C" int ckh_movegen(const ck_pos *pos, ck_pos *children, uint32_t *mask, int *status, int *plane5) {
    int n = ck::gen_moves(*pos, ck::ArraySink{children, CK_MAX_CHILDREN}, mask);
    int p5;
    int st = ck::outcome_of(*pos, n > 0, &p5);
    *status = st; *plane5 = p5;
    return n;
}
// random access: successor k through kth_successor (the packed movegen kernel's per-lane path)
extern "C" void ckh_kth(const ck_pos *pos, int k, ck_pos *out) {
    uint32_t mask[8];
    ck::gen_moves(*pos, ck::NullSink{}, mask);
    const bool jump = (mask[4] | mask[5] | mask[6] | mask[7]) != 0;
    *out = ck::kth_successor(*pos, jump ? mask + 4 : mask, jump, k);
}
// every successor once more through make_child_fast, in gen_moves' order (men, then kings; order_dir)
extern "C" int ckh_movegen_fast(const ck_pos *pos, ck_pos *children) {
    uint32_t mask[8];
    ck::gen_moves(*pos, ck::NullSink{}, mask);
    const bool jump = (mask[4] | mask[5] | mask[6] | mask[7]) != 0;
    const uint32_t *use = jump ? mask + 4 : mask;
    const ck::Side sd = ck::side_of(*pos);
    uint32_t J[4];
    ck::hop_sets(sd, J);
    const uint32_t any = use[0] | use[1] | use[2] | use[3];
    int n = 0;
    for (int pass = 0; pass < 2; ++pass)
        for (uint32_t rem = any & (pass ? sd.kings : ~sd.kings); rem; rem &= rem - 1) {
            const int s = ck::ffs32(rem);
            for (int i = 0; i < (pass ? 4 : 2); ++i) {
                const int d = ck::order_dir(pass == 1, jump, sd.player, i);
                if ((use[d] >> s) & 1u) children[n++] = ck::make_child_fast(*pos, sd, J, s, d, jump);
            }
        }
    return n;
}
// the playout loop of K4 / playout_eval_kernel with the hashed successor choice (oracle: cko_hash_playout)
extern "C" int ckh_hash_playout(const ck_pos *pos, uint32_t salt, int *plies) {
    return ck::play_out(*pos, 0, ck::HashChoice{salt}, plies);
}
extern "C" int ckh_status(const ck_pos *pos, int *plane5) { return ck::status_of(*pos, plane5); }
extern "C" void ckh_start(ck_pos *p) { *p = ck::start_position(); }
