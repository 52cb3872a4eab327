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
extern "C" int ckh_status(const ck_pos *pos, int *plane5) { return ck::status_of(*pos, plane5); }
extern "C" void ckh_start(ck_pos *p) { *p = ck::start_position(); }
