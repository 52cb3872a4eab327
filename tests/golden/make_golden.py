#!/usr/bin/env python3
"""Generates the golden vectors in tests/golden/ by running the UNMODIFIED reference
(/root/reference, via oracle/ref_harness.py).  Run in the build container only:

    python tests/golden/make_golden.py

Outputs (all small, committed):
  movegen_cases.npz   positions + reference children / masks / outcome / plane 5
  perft.json          perft under reference semantics (hop == ply), depth 1..7
  predict_glue.npz    Checkers.predict mask+renormalise on random policy vectors
  mcts_kat.json       KAT-A / KAT-B / hash-stub first searches + per-move root statistics
  selfplay_*.npz      full records of _generate_data run verbatim with stub nets
  tournament.json     _start_tournament run verbatim with two different stub nets
  net_model10.npz, net_model5.npz
                      the reference's own trained networks data/model/Checkers_Model{10,5}_*.h5 (read with
                      ckb200.h5lite): the float32 weight blob, 256 positions from uniformly random play, and the
                      float64 restatement's policy logits / pre-tanh value / softmax / tanh on them
                      (python tests/golden/make_golden.py net  regenerates only these)
  tournament_results.json
                      W/L/D of the reference's own evaluation tournaments (data/tournament_results/Tournament*.txt)
                      and its final round-robin table (data/final_eval/*.txt), parsed
  uct_kat.json        NEURAL_NET=False (UCT + one playout per simulation, the iteration-0 mode): first search
                      and per-move root statistics of a _generate_data game, with np.random.randint inside
                      MCTS.default_policy replaced by a hash of the position the move is chosen from
                      (python tests/golden/make_golden.py uct  regenerates only this file)
"""
import json
import os
import pickle
import shutil
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from helpers import KerasLikeStub, codec, hashed_playouts  # noqa: E402
from oracle import ref_harness as H  # noqa: E402


# ---------------------------------------------------------------------------------------
def ref_case(env, history, rev, ply):
    """Run the reference rules on history[-1]; returns a dict in compact form."""
    st = history[-1]
    raw = env._check_moves(history)
    done, outcome = env.determine_outcome(history, legal_moves=raw)
    pos = codec.encode_state(st, rev, ply)
    kids = []
    for r in raw:
        e = codec.encode_state(r, 0, 0)
        kids.append((e[0], e[1], e[2], codec.meta_player(e[3]), codec.meta_action(e[3])))
    mask = [codec.plane_to_bits(st[6 + i]) for i in range(8)]
    p5 = st[5, 0, 0] * 80
    assert (st[5] == st[5, 0, 0]).all() and abs(p5 - round(p5)) < 1e-9
    return dict(pos=pos, kids=kids, mask=mask, status=codec.OUTCOME_CODES[outcome], plane5=int(round(p5)))


def child_rev(env_state_before, child_state):
    """rev update rule cross-checked here from the planes: a ply is reversible iff the men
    planes and the piece count are unchanged (Checkers.py:336-343)."""
    same_men = (env_state_before[0] == child_state[0]).all() and (env_state_before[2] == child_state[2]).all()
    same_cnt = env_state_before[0:4].sum() == child_state[0:4].sum()
    return same_men and same_cnt


def random_game_cases(ref, rng, n_games, start_state=None, max_plies=400):
    env = ref.Checkers.Checkers()
    out = []
    for _ in range(n_games):
        env.reset()
        if start_state is not None:
            env.state = start_state.copy()
            env.history = [env.state]
            env.legal_next_states = env.get_legal_next_states(env.history)
            env.done, env.outcome = env.determine_outcome(env.history)
        rev, ply = 0, 0
        while True:
            out.append(ref_case(env, env.history, rev, ply))
            if env.done or ply >= max_plies:
                break
            before = env.state
            nxt = env.legal_next_states[rng.randint(len(env.legal_next_states))]
            env.step(nxt)
            rev = rev + 1 if child_rev(before, nxt) else 0
            ply += 1
    return out


def synthetic_state(rng):
    """random men/kings placement, either side to move (history of length 1)."""
    st = np.zeros((15, 8, 8))
    squares = [(x, y) for x in range(8) for y in range(8) if x % 2 != y % 2]
    n = rng.randint(2, 25)
    idx = rng.permutation(32)[:n]
    for j, i in enumerate(idx):
        x, y = squares[i]
        side = j % 2 if rng.rand() < 0.8 else rng.randint(2)
        king = rng.rand() < 0.35
        if not king and ((side == 0 and x == 7) or (side == 1 and x == 0)):
            king = True           # a man cannot stand on its king row
        st[side * 2 + (1 if king else 0), x, y] = 1
    st[4] = rng.randint(2)
    return st


def make_movegen(ref):
    rng = np.random.RandomState(20261017)
    cases = random_game_cases(ref, rng, 25)
    # king endgames: long reversible sequences -> draw rule + plane 5 (Checkers.py:332-343,357-360)
    for _ in range(6):
        st = np.zeros((15, 8, 8))
        sq = [(x, y) for x in range(8) for y in range(8) if x % 2 != y % 2]
        pick = rng.permutation(32)[:4]
        for j, i in enumerate(pick):
            st[1 if j < 2 else 3, sq[i][0], sq[i][1]] = 1
        st[4] = rng.randint(2)
        cases += random_game_cases(ref, rng, 1, start_state=st, max_plies=200)
    env = ref.Checkers.Checkers()
    for _ in range(1500):
        st = synthetic_state(rng)
        cases.append(ref_case(env, [st], 0, 0))
    pos = np.array([c["pos"] for c in cases], dtype=np.uint32)
    counts = np.array([len(c["kids"]) for c in cases], dtype=np.int32)
    kids = np.array([k for c in cases for k in c["kids"]], dtype=np.uint32).reshape(-1, 5)
    mask = np.array([c["mask"] for c in cases], dtype=np.uint32)
    status = np.array([c["status"] for c in cases], dtype=np.int32)
    plane5 = np.array([c["plane5"] for c in cases], dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "movegen_cases.npz"), pos=pos, counts=counts, kids=kids,
                        mask=mask, status=status, plane5=plane5)
    print("movegen cases:", len(cases), "children:", len(kids), "max b:", counts.max(),
          "draws:", int((status == 3).sum()), "plane5>0:", int((plane5 > 0).sum()))


def make_perft(ref):
    env = ref.Checkers.Checkers()

    def perft(history, d):
        nxt = env.get_legal_next_states(history)
        if d == 1:
            hops = sum(1 for s in nxt if s[14, 0, 0] >= 10)
            cont = sum(1 for s in nxt if s[4, 0, 0] == history[-1][4, 0, 0])
            return len(nxt), hops, cont
        t = [0, 0, 0]
        for s in nxt:
            r = perft(history + [s], d - 1)
            t = [a + b for a, b in zip(t, r)]
        return tuple(t)

    res = {}
    for d in range(1, 8):
        env.reset()
        res[str(d)] = list(perft([env.state], d))
        print("perft", d, res[str(d)])
    res["8_survey_reported"] = [838275, 112891, 17769]
    json.dump(res, open(os.path.join(HERE, "perft.json"), "w"), indent=1)


def make_predict_glue(ref):
    rng = np.random.RandomState(7)

    class Net(object):
        def predict(self, x):
            return [self.pol.copy().reshape(1, 512), np.array([[np.float32(0.25)]], dtype=np.float32)]

    net = Net()
    env = ref.Checkers.Checkers(net)
    pols, masks, outs = [], [], []
    for i in range(64):
        env.reset()
        for _ in range(rng.randint(0, 40)):
            if env.done:
                break
            env.step(env.legal_next_states[rng.randint(len(env.legal_next_states))])
        if env.done:
            continue
        logits = rng.randn(512).astype(np.float32) * (3 if i % 2 else 1)
        e = np.exp(logits - logits.max()).astype(np.float32)
        net.pol = (e / e.sum(dtype=np.float32)).astype(np.float32)
        prob_planes, q = env.predict(env.state)
        assert prob_planes.dtype == np.float32
        pols.append(net.pol)
        masks.append([codec.plane_to_bits(env.state[6 + j]) for j in range(8)])
        outs.append(prob_planes.reshape(512))
    np.savez_compressed(os.path.join(HERE, "predict_glue.npz"), policy=np.array(pols, dtype=np.float32),
                        mask=np.array(masks, dtype=np.uint32), prior=np.array(outs, dtype=np.float32))
    print("predict glue cases:", len(pols))


MCTS_KW = dict(GAME_ENV=None, UCT_C=4, CONSTRAINT='rollout', BUDGET=400, MULTIPROC=False, NEURAL_NET=True,
               VERBOSE=False, TRAINING=False, DIRICHLET_ALPHA=1.0, DIRICHLET_EPSILON=0.0, TEMPERATURE_TAU=0,
               TEMPERATURE_DECAY=0, TEMP_DECAY_DELAY=0)


def node_children(node):
    return [dict(action=codec.action_id(c.state[14, 0, 0], c.state[14, 0, 1], c.state[14, 0, 2]),
                 n=int(c.n), w=float(c.w), p=float(c.p), terminal=bool(c.terminal)) for c in node.children]


def run_pipeline_game(kind, budget, terminate_cnt, workdir, log, **mcts_overrides):
    """generate_Checkers_data._generate_data run verbatim; MCTS.best_child wrapped to log the
    root statistics at every move (the wrapper only observes)."""
    H.set_load_model(lambda path: KerasLikeStub(kind))
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        with H.reference_modules(with_pipeline=True) as ref:
            orig = ref.MCTS.MCTS.best_child.__func__

            def spy(cls, node, criterion='robust'):
                log.append(dict(root_n=int(node.n), root_w=float(node.w), children=node_children(node)))
                return orig(cls, node, criterion)

            ref.MCTS.MCTS.best_child = classmethod(spy)
            sp = dict(NUM_SELFPLAY_GAMES=1, TRAINING_ITERATION=0, TERMINATE_CNT=terminate_cnt, NUM_CPUS=1, NN_FN='stub')
            mk = dict(MCTS_KW, BUDGET=budget, TRAINING=True, **mcts_overrides)
            fn = ref.training_pipeline.generate_Checkers_data(sp, mk).generate_data()
            data = pickle.load(open(fn, 'rb'))
    finally:
        os.chdir(cwd)
    return data


def save_records(name, data, log, meta):
    states = np.array([e[0] for e in data], dtype=np.float64)
    # store compactly: bit planes + scalars
    pl = np.array([[codec.plane_to_bits(s[i]) for i in (0, 1, 2, 3, 6, 7, 8, 9, 10, 11, 12, 13)] for s in states],
                  dtype=np.uint32)
    player = states[:, 4, 0, 0].astype(np.int8)
    plane5 = np.round(states[:, 5, 0, 0] * 80).astype(np.int16)
    assert np.allclose(states[:, 5, 0, 0], plane5 / 80, atol=0, rtol=0)
    act = states[:, 14, 0, 0:3].astype(np.int16)
    probs = np.array([e[1] for e in data], dtype=np.float64).reshape(len(data), 512)
    q = np.array([float(e[2]) for e in data], dtype=np.float64)
    z = np.array([int(e[3]) for e in data], dtype=np.int8)
    np.savez_compressed(os.path.join(HERE, name), planes=pl, player=player, plane5=plane5, action=act,
                        probs=probs, q=q, z=z,
                        root_n=np.array([m["root_n"] for m in log], dtype=np.int64),
                        root_w=np.array([m["root_w"] for m in log], dtype=np.float64),
                        meta=json.dumps(meta))


def make_mcts(workdir):
    kat = {}
    for kind in ("uniform_zero", "uniform_material", "hash"):
        with H.reference_modules() as ref:
            env = ref.Checkers.Checkers(KerasLikeStub(kind))
            ref.MCTS.MCTS(**dict(MCTS_KW, GAME_ENV=env))
            root = ref.MCTS.MCTS_Node(env.state)
            ref.MCTS.MCTS.begin_tree_search(root)
            kat[kind + "_first_search"] = dict(budget=400, root_n=int(root.n), root_w=float(root.w),
                                               children=node_children(root))
    # per-move statistics over the opening of a self-play game (two trees, re-rooting)
    for kind, budget, plies in (("uniform_zero", 400, 12), ("uniform_material", 400, 15), ("hash", 200, 40)):
        log = []
        t = time.time()
        data = run_pipeline_game(kind, budget, plies, workdir, log)
        kat[kind + "_game"] = dict(budget=budget, terminate_cnt=plies, moves=log,
                                   chosen=[[int(v) for v in e[0][14, 0, 0:3]] for e in data[1:]])
        print("kat game", kind, len(log), "plies", round(time.time() - t, 1), "s")
    json.dump(kat, open(os.path.join(HERE, "mcts_kat.json"), "w"))
    # full games with records
    for kind, budget, cap in (("hash", 100, 120), ("uniform_material", 60, 200)):
        log = []
        t = time.time()
        data = run_pipeline_game(kind, budget, cap, workdir, log)
        save_records("selfplay_%s.npz" % kind, data, log, dict(kind=kind, budget=budget, terminate_cnt=cap))
        print("selfplay", kind, len(data), "records", round(time.time() - t, 1), "s")


def make_uct(workdir):
    kat = {}
    with hashed_playouts():
        with H.reference_modules() as ref:
            env = ref.Checkers.Checkers(None)
            ref.MCTS.MCTS(**dict(MCTS_KW, GAME_ENV=env, NEURAL_NET=False, BUDGET=300))
            root = ref.MCTS.MCTS_Node(env.state)
            t = time.time()
            ref.MCTS.MCTS.begin_tree_search(root)
            kat["first_search"] = dict(budget=300, root_n=int(root.n), root_w=float(root.w), children=node_children(root))
            print("uct first search", round(time.time() - t, 1), "s")
        for name, budget, plies in (("game", 150, 40),):
            log = []
            t = time.time()
            data = run_pipeline_game("uniform_zero", budget, plies, workdir, log, NEURAL_NET=False)
            kat[name] = dict(budget=budget, terminate_cnt=plies, moves=log,
                             chosen=[[int(v) for v in e[0][14, 0, 0:3]] for e in data[1:]],
                             q=[float(e[2]) for e in data], z=[int(e[3]) for e in data])
            print("uct game", len(log), "plies", round(time.time() - t, 1), "s")
    json.dump(kat, open(os.path.join(HERE, "uct_kat.json"), "w"))


def make_tournament(workdir):
    nets = {"data/model/A": "hash", "data/model/B": "uniform_material"}
    H.set_load_model(lambda path: KerasLikeStub(nets[path]))
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        with H.reference_modules(with_pipeline=True) as ref:
            tk = dict(NEW_NN_FN="data/model/A", OLD_NN_FN="data/model/B", TOURNEY_GAMES=2, NUM_CPUS=1)
            mk = dict(MCTS_KW, BUDGET=60)
            t = ref.training_pipeline.tournament_Checkers(tk, mk)
            outcomes = t._start_tournament()
    finally:
        os.chdir(cwd)
    json.dump(dict(budget=60, nets=nets, outcomes=outcomes), open(os.path.join(HERE, "tournament.json"), "w"), indent=1)
    print("tournament:", outcomes)


def random_play_leaves(n, seed):
    """positions reached by uniformly random legal play (oracle rules, pinned to the reference by the movegen
    goldens above) -> (leaves as LEAF_DTYPE-compatible uint32 [n,12], NN input planes float32 [n,8,8,14])"""
    from oracle import oracle as O
    rng = np.random.RandomState(seed)
    leaves = np.zeros((n, 12), dtype=np.uint32)
    planes = np.zeros((n, 8, 8, 14), dtype=np.float32)
    i = 0
    while i < n:
        pos = O.start_position()
        for _ply in range(rng.randint(1, 140)):
            kids, mask, st, p5 = O.movegen(pos)
            if st != codec.ONGOING:
                break
            pos = kids[rng.randint(len(kids))]
        kids, mask, st, p5 = O.movegen(pos)
        if st != codec.ONGOING:
            continue
        leaves[i] = [pos[0], pos[1], pos[2], (pos[3] & 1) | (p5 << 8)] + list(mask)
        planes[i] = codec.nn_input_planes(pos, mask, p5)
        i += 1
    return leaves, planes


def make_net():
    """trained-weight fixtures for K3 (the reference ships 11 networks, training_pipeline.py:345,515-516)"""
    import glob
    from ckb200 import h5lite
    from ckb200 import net as N
    from oracle import net_oracle as NO
    leaves, planes = random_play_leaves(256, 20261018)
    for it in (10, 5):
        fn = glob.glob(os.path.join(H.REFERENCE_DIR, "data", "model", "Checkers_Model%d_*.h5" % it))[0]
        blob = h5lite.keras_h5_to_blob(fn)
        pol, val, logits, vpre = NO.forward(N.unpack(blob), planes, pre_activation=True)
        np.savez_compressed(os.path.join(HERE, "net_model%d.npz" % it), blob=blob.astype(np.float32), leaves=leaves,
                            policy=pol, value=val, logits=logits, value_pre=vpre, source=os.path.basename(fn))
        print("net golden", os.path.basename(fn), "max|logit|", float(np.abs(logits).max()), "max|vpre|", float(np.abs(vpre).max()))


def make_tournament_results():
    """the reference's recorded evaluation results, parsed (which iteration beat which): the statistical pin of the
    network semantics -- a replay of a pairing on the engine has to land in the same place"""
    import glob
    import re
    out = {"tournaments": [], "final_eval": None}
    for fn in sorted(glob.glob(os.path.join(H.REFERENCE_DIR, "data", "tournament_results", "Tournament*.txt"))):
        rows = re.findall(r"Checkers_Model(\d+)_[^ ]+\.h5\s*\u2502\s*(\d+)/(\d+)/(\d+)", open(fn, encoding="utf-8").read())
        (a, w, l, d), (b, _, _, _) = rows[0], rows[1]
        out["tournaments"].append(dict(file=os.path.basename(fn), new=int(a), old=int(b), wins=int(w), losses=int(l), draws=int(d)))
    fe = glob.glob(os.path.join(H.REFERENCE_DIR, "data", "final_eval", "Checkers_Final_Evaluation_*.txt"))
    fe = [f for f in fe if "Params" not in f]
    if fe:
        table = []
        for line in open(fe[0], encoding="utf-8"):
            cells = [c.strip() for c in line.split("\u2502")[1:-1]]
            if len(cells) == 13 and cells[0].isdigit():
                table.append([int(c) for c in cells[1:]])
        out["final_eval"] = dict(file=os.path.basename(fe[0]), budget=400, table=table)
    json.dump(out, open(os.path.join(HERE, "tournament_results.json"), "w"), indent=1)
    print("tournament results:", [(t["new"], t["old"], t["wins"], t["losses"], t["draws"]) for t in out["tournaments"]])


def main():
    assert H.reference_available(), "reference not mounted"
    if sys.argv[1:] == ["net"]:
        make_net()
        make_tournament_results()
        return
    workdir = tempfile.mkdtemp(prefix="ckref_")
    os.makedirs(os.path.join(workdir, "data/training_data"))
    os.makedirs(os.path.join(workdir, "data/tournament_results"))
    try:
        if sys.argv[1:] == ["uct"]:
            make_uct(workdir)
            return
        with H.reference_modules() as ref:
            make_movegen(ref)
            make_predict_glue(ref)
            make_perft(ref)
        make_mcts(workdir)
        make_uct(workdir)
        make_tournament(workdir)
    finally:
        shutil.rmtree(workdir, ignore_errors=True)


if __name__ == "__main__":
    main()
