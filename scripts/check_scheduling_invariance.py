"""Diagnostic: do scheduling knobs change any record?  Plays 4096 warm-started games with the network evaluator under different
batch-shaping / chain-cap / cache settings (one subprocess each: the shaping knobs are read once per process) and
compares the record arrays field by field.  Everything must print IDENTICAL.  Usage: python scripts/debug_shaping.py [slots]"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "checkers-mcts_b200")
code = (
    "import sys, json; sys.path[:0] = %r\n"
    "import numpy as np\n"
    "from ckb200 import lib, net as N\n"
    "a = json.loads(sys.argv[2])\n"
    "net = lib.Net(0); net.set_weights(N.random_init_blob(3))\n"
    "eng = lib.Engine(lib.make_cfg(n_slots=a['slots'], budget=a['budget'], training=True, terminate_cnt=a['term'], evaluator=a['ev'], seed=5,\n"
    "                              stagger_budget=a['sb'], stagger_plies=a['sp'], max_chain_per_step=a['chain'], eval_cache_entries=a['cache']))\n"
    "if a['ev'] == 'net': eng.set_net(0, net)\n"
    "st = eng.selfplay(a['slots'])\n"
    "r = eng.records(); r = r[np.lexsort((r['ply'], r['game']))]\n"
    "np.save(sys.argv[1], r); print(json.dumps(dict(steps=st['steps'], sims=st['sims'])))\n") % ([ROOT, PKG],)


def run(tmp, name, env_extra, **a):
    fn = os.path.join(tmp, name + ".npy")
    env = dict(os.environ)
    env.update(env_extra)
    r = subprocess.run([sys.executable, "-c", code, fn, json.dumps(a)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return np.load(fn), json.loads(r.stdout.strip().splitlines()[-1])


def diff(a, b, label):
    if a.tobytes() == b.tobytes():
        print(label, "IDENTICAL", len(a))
        return
    bad = {}
    for f in a.dtype.names:
        x, y = a[f], b[f]
        ne = (np.ascontiguousarray(x).view(np.uint8).reshape(len(a), -1) != np.ascontiguousarray(y).view(np.uint8).reshape(len(b), -1)).any(axis=1)
        if ne.any():
            bad[f] = (int(ne.sum()), int(np.argmax(ne)))
    print(label, "DIFFERENT", len(a), bad)
    any_ne = (a.view(np.uint8).reshape(len(a), -1) != b.view(np.uint8).reshape(len(b), -1)).any(axis=1)
    idx = np.nonzero(any_ne)[0]
    games = np.unique(a["game"][idx])
    print("  differing records", len(idx), "games", len(games), "first plies per game", sorted({int(a["ply"][i]) for i in idx})[:10])
    i = idx[0]
    for f in ("game", "ply", "pos", "n_children", "root_n", "root_w", "q", "chosen"):
        print("   ", f, a[f][i], b[f][i])
    n = int(a["n_children"][i])
    print("    visits", a["visits"][i][:n], b["visits"][i][:n])
    firsts = {}
    for j in idx:
        g = int(a["game"][j])
        firsts.setdefault(g, int(a["ply"][j]))
    print("  first differing ply per game (up to 12):", list(firsts.items())[:12])


base = dict(slots=int(sys.argv[1]) if len(sys.argv) > 1 else 4096, budget=96, term=24, ev="net", sb=8, sp=16, chain=0, cache=0)
with tempfile.TemporaryDirectory() as tmp:
    off, s0 = run(tmp, "off", {"CK_BATCH_WAVES": "0"}, **base)
    off2, s0b = run(tmp, "off2", {"CK_BATCH_WAVES": "0"}, **base)
    diff(off, off2, "off vs off (repeat)")
    dflt, s1 = run(tmp, "dflt", {}, **base)
    diff(off, dflt, "off vs default shaping")
    ch2, s2 = run(tmp, "ch2", {"CK_BATCH_WAVES": "0"}, **dict(base, chain=2))
    diff(off, ch2, "off vs off+chain2")
    nc, s3 = run(tmp, "nocache", {"CK_BATCH_WAVES": "0"}, **dict(base, cache=-1))
    diff(off, nc, "off vs off+nocache")
    ns, s4 = run(tmp, "nostagger", {"CK_BATCH_WAVES": "0"}, **dict(base, sb=0, sp=0))
    ns2, s5 = run(tmp, "nostagger_shaped", {}, **dict(base, sb=0, sp=0))
    diff(ns, ns2, "no stagger: off vs default shaping")
    print(s0, s0b, s1, s2, s3, s4, s5)
