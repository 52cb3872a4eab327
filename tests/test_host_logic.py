"""CPU tests of the host-side Python that sits above the C ABI: state codec, weight-blob layout,
record conversion to the reference's pickle format, the torch restatement of the network."""
import os

import numpy as np

from helpers import GOLDEN, codec
from ckb200 import net as N
from ckb200 import records as R
from oracle import oracle as O


def test_net_layout_matches_reference_parameter_count():
    lay = N.layout()
    assert sum(int(np.prod(s)) for _o, s in lay.values()) == N.NET_PARAM_COUNT == 1321774
    offs = [o for o, _s in lay.values()]
    assert offs == sorted(offs) and offs[0] == 0
    blob = N.random_init_blob(0)
    p = N.unpack(blob)
    assert p["conv0/kernel"].shape == (3, 3, 14, 128) and p["policy_head/kernel"].shape == (512, 512)
    assert (p["conv3/bn_gamma"] == 1).all() and (p["conv3/bias"] == 0).all() and (p["value_dense1/bn_var"] == 1).all()
    lim = np.sqrt(6.0 / (9 * 128 + 9 * 128))
    assert np.abs(p["conv1/kernel"]).max() <= lim and np.abs(p["conv1/kernel"]).max() > 0.9 * lim
    assert (N.random_init_blob(0) == blob).all() and (N.random_init_blob(1) != blob).any()


def test_torch_restatement_shapes_and_normalisation():
    from oracle import net_oracle as NO
    blob = N.random_init_blob(2, 0.1)
    x = np.zeros((3, 8, 8, 14), dtype=np.float32)
    x[0] = codec.nn_input_planes(O.start_position(), O.movegen(O.start_position())[1], 0)
    x[1, ..., 4] = 1
    p, v = NO.forward(N.unpack(blob), x)
    assert p.shape == (3, 512) and v.shape == (3,) and np.allclose(p.sum(1), 1) and (np.abs(v) < 1).all()
    m = NO.TorchKerasLike(blob)
    kp, kv = m.predict(x)
    assert kp.dtype == np.float32 and kv.shape == (3, 1) and np.abs(kp - p).max() < 1e-5


def test_record_conversion_matches_reference_pickle_layout():
    """oracle self-play records -> [state, probs, q, z] must equal the golden reference stream"""
    f = np.load(os.path.join(GOLDEN, "selfplay_hash.npz"))
    import json
    meta = json.loads(str(f["meta"]))
    g = O.Game(O.make_cfg(budget=meta["budget"], training=True, terminate_cnt=meta["terminate_cnt"]), "hash")
    g.play()
    from ckb200 import lib_types as T
    recs = T.records_from_dicts(g.records())
    out = R.to_reference_list(recs)
    assert len(out) == len(f["q"])
    for i, (state, probs, q, z) in enumerate(out):
        assert state.shape == (15, 8, 8) and state.dtype == np.float64
        assert [codec.plane_to_bits(state[j]) for j in (0, 1, 2, 3, 6, 7, 8, 9, 10, 11, 12, 13)] == [int(v) for v in f["planes"][i]]
        assert probs.reshape(512).tobytes() == f["probs"][i].tobytes()
        assert float(q) == f["q"][i] and z == f["z"][i]
    assert isinstance(out[-1][2], int) and out[-1][2] in (0, -1)          # terminal record carries a Python int
    x, (p, v) = R.training_batch(recs[:5])
    assert x.shape == (5, 8, 8, 14) and p.shape == (5, 512) and v.shape == (5,)
    assert v[0] == (float(recs[0]["q"]) + int(recs[0]["z"])) / 2


def test_codec_roundtrip_random_positions():
    rng = np.random.RandomState(0)
    pos = O.start_position()
    for _ in range(200):
        kids, mask, st, p5 = O.movegen(pos)
        state = codec.decode_state(pos, mask, p5)
        back = codec.encode_state(state, codec.meta_rev(pos[3]), codec.meta_ply(pos[3]))
        assert back[:3] == pos[:3] and codec.meta_player(back[3]) == codec.meta_player(pos[3])
        assert codec.meta_action(back[3]) == codec.meta_action(pos[3])
        if st != codec.ONGOING:
            pos = O.start_position()
        else:
            pos = kids[rng.randint(len(kids))]


def _oracle_reference_records():
    g = O.Game(O.make_cfg(budget=24, training=True, terminate_cnt=40), "hash")
    g.play()
    from ckb200 import lib_types as T
    return R.to_reference_list(T.records_from_dicts(g.records()))


def test_keras_generator_and_merge_data(tmp_path, monkeypatch):
    """loader side of the self-play wire format: the batcher reproduces
    Keras_Generator.__getitem__ (reference training_pipeline.py:297-307), merge_data concatenates the
    per-worker pickles (:277-284)."""
    import pickle
    import training_pipeline as TP
    data = _oracle_reference_records()
    gen = TP.Keras_Generator(data, 7)
    assert len(gen) == -(-len(data) // 7)
    seen = 0
    for b, (x, (probs, v)) in enumerate(gen):
        chunk = data[b * 7:(b + 1) * 7]
        assert x.shape == (len(chunk), 8, 8, 14) and x.dtype == np.float64
        assert probs.shape == (len(chunk), 512) and v.shape == (len(chunk),)
        for j, e in enumerate(chunk):
            assert (x[j] == np.moveaxis(e[0][:14], 0, -1)).all()
            assert (probs[j] == e[1].reshape(512)).all()
            assert v[j] == (np.float64(e[2]) + e[3]) / 2          # q is float32 (or a Python int), the batch float64
        seen += len(chunk)
    assert seen == len(data)
    # against the reference class itself where it is mounted (the build container)
    ref_dir = "/root/reference"
    if os.path.isdir(ref_dir):
        from oracle import ref_harness as RH
        ref_tp = RH.load_reference(with_pipeline=True).training_pipeline
        rg = ref_tp.Keras_Generator(data, 7)
        for b in range(len(gen)):
            x, (probs, v) = gen[b]
            rx, (rp, rv) = rg[b]                       # reference __len__ uses np.int (gone in numpy 2); __getitem__ runs
            assert x.tobytes() == rx.tobytes() and probs.tobytes() == rp.tobytes() and v.tobytes() == rv.tobytes()
    # merge_data
    monkeypatch.chdir(tmp_path)
    os.makedirs("data/training_data")
    for k in range(3):
        with open("data/training_data/part%d.pkl" % k, "wb") as f:
            pickle.dump(data[k::3], f)
    merged = TP.merge_data(["part0.pkl", "part1.pkl", "part2.pkl"], 5)
    assert len(merged) == len(data)
    files = [f for f in os.listdir("data/training_data") if f.startswith("Checkers_Data5_")]
    assert len(files) == 1
    back = TP.load_training_data(os.path.join("data/training_data", files[0]))
    assert len(back) == len(data) and (back[0][0] == data[0][0]).all()


def test_keras_h5_import_of_reference_models():
    """the reference's shipped .h5 networks are read without h5py/TensorFlow and mapped onto
    the weight blob by layer topology.  Runs where the reference tree is mounted (the build container)."""
    import pytest
    d = "/root/reference/data/model"
    if not os.path.isdir(d):
        pytest.skip("reference models not mounted")
    from ckb200 import h5lite
    from oracle import net_oracle as NO
    fns = sorted(f for f in os.listdir(d) if f.endswith(".h5"))
    assert len(fns) == 11
    init = [f for f in fns if "Model0_" in f][0]
    dsets, raw = h5lite.read_datasets(os.path.join(d, init))
    assert len(dsets) == 70 and sum(a.size for a in dsets.values()) == N.NET_PARAM_COUNT
    roles = h5lite.layer_roles(h5lite.model_config(raw))
    assert roles["policy_head"] == "policy_head" and roles["value_head"] == "value_head" and roles["conv0"] == "conv2d"
    p0 = N.unpack(h5lite.keras_h5_to_blob(os.path.join(d, init)))
    # iteration 0 is the freshly initialised network: Glorot-uniform kernels, unit BN statistics
    for name, fan in (("conv0/kernel", 9 * 14 + 9 * 128), ("conv3/kernel", 18 * 128), ("policy_head/kernel", 1024)):
        lim = np.sqrt(6.0 / fan)
        assert 0.95 * lim < np.abs(p0[name]).max() <= lim * (1 + 1e-6)
    assert (p0["conv5/bn_var"] == 1).all() and (p0["conv5/bn_gamma"] == 1).all() and (p0["value_dense1/bn_mean"] == 0).all()
    # a trained iteration: the imported network must put its policy mass on the legal moves of the start
    # position (it does only if conv / BN / Flatten(x,y,c) / Dense semantics and the layer mapping are right)
    sp = O.start_position()
    kids, mask, st, p5 = O.movegen(sp)
    x = codec.nn_input_planes(sp, mask, p5).reshape(1, 8, 8, 14)
    legal = np.zeros(512, dtype=bool)
    for k in kids:
        legal[codec.meta_action(k[3])] = True
    last = [f for f in fns if "Model10_" in f][0]
    import training_pipeline as TP
    blob = TP.load_blob(os.path.join(d, last))
    p, v = NO.forward(N.unpack(blob), x)
    assert p[0][legal].sum() > 0.98 and legal[p[0].argmax()] and abs(v[0]) < 0.5
    pi, _ = NO.forward(p0, x)
    assert pi[0][legal].sum() < 0.1                                  # the untrained one does not


def test_training_step_matches_reference_recipe(tmp_path, monkeypatch):
    """the PyTorch training network equals the float64 restatement on the same blob, blobs
    round-trip, the CLR schedule is the reference's triangular formula, and a short train_nn run lowers
    the loss and writes the best epoch's weights."""
    import torch
    import training_pipeline as TP
    from ckb200 import train as T
    from oracle import net_oracle as NO
    blob = N.random_init_blob(3, 0.2)
    model = T.CheckersNet(blob)
    assert (model.blob() == blob).all()                                      # Keras layouts <-> torch layouts
    data = _oracle_reference_records()
    gen = TP.Keras_Generator(data, 16)
    x, (probs, v) = gen[0]
    kp, kv = model.predict(x)
    rp, rv = NO.forward(N.unpack(blob), np.asarray(x, dtype=np.float32))
    assert np.abs(kp - rp).max() < 1e-5 and np.abs(kv.reshape(-1) - rv).max() < 1e-5
    # CLR: triangular, base at 0 and 2*step, max at step (CLR/clr_callback.py:105-111)
    assert T.clr_triangular(0, 5e-5, 1e-2, 40) == 5e-5 and abs(T.clr_triangular(40, 5e-5, 1e-2, 40) - 1e-2) < 1e-12
    assert abs(T.clr_triangular(20, 5e-5, 1e-2, 40) - (5e-5 + (1e-2 - 5e-5) * 0.5)) < 1e-12
    assert abs(T.clr_triangular(80, 5e-5, 1e-2, 40) - 5e-5) < 1e-12 and abs(T.clr_triangular(100, 5e-5, 1e-2, 40) - (5e-5 + (1e-2 - 5e-5) * 0.5)) < 1e-12
    # loss = CE + MSE + L2 over kernels and biases
    xt = torch.as_tensor(np.asarray(x, dtype=np.float32))
    loss, ce, mse = T.loss_terms(model.eval(), xt, torch.as_tensor(np.asarray(probs, dtype=np.float32)), torch.as_tensor(np.asarray(v, dtype=np.float32)))
    p = N.unpack(blob)
    l2 = sum(float((p[k].astype(np.float64) ** 2).sum()) for k in p if k.endswith("/kernel") or k.endswith("/bias")) * 1e-3
    ref_ce = float(-(np.asarray(probs) * np.log(np.clip(rp, 1e-7, None))).sum(1).mean())
    ref_mse = float(((rv - np.asarray(v)) ** 2).mean())
    assert abs(ce.item() - ref_ce) < 1e-4 and abs(mse.item() - ref_mse) < 1e-5 and abs(loss.item() - (ref_ce + ref_mse + l2)) < 1e-3
    # a short run
    monkeypatch.chdir(tmp_path)
    os.makedirs("data/model")
    np.random.seed(0); torch.manual_seed(0)
    kw = dict(TRAINING_ITERATION=0, NN_BASE_LR=5e-5, NN_MAX_LR=2e-3, CLR_SS_COEFF=4, BATCH_SIZE=16, EPOCHS=3, CONV_REG=1e-3,
              DENSE_REG=1e-3, NUM_KERNELS=128, VAL_SPLIT=0.2, MIN_DELTA=0.01, PATIENCE=20, POLICY_LOSS_WEIGHT=1.0, VALUE_LOSS_WEIGHT=1.0)
    nn0 = TP.create_nn(**kw)
    hist, fn = TP.train_nn(list(data), nn0, **dict(kw, device="cpu", verbose=False))
    assert len(hist["loss"]) == 3 and hist["loss"][-1] < hist["loss"][0] and len(hist["val_loss"]) == 3
    assert fn.startswith("data/model/Checkers_Model1_") and os.path.exists(fn)
    out = TP.load_blob(fn)
    assert out.shape == (N.NET_PARAM_COUNT,) and (out != N.random_init_blob(0)).any()
    assert os.path.exists(TP.plot_history(hist, nn0, 0))
    # the reference's own CyclicLR where it is mounted (last: its import leaves TensorFlow stubs in sys.modules)
    if os.path.isdir("/root/reference"):
        from oracle import ref_harness as RH
        ref = RH.load_reference(with_pipeline=True).training_pipeline
        clr = ref.CyclicLR(base_lr=5e-5, max_lr=1e-2, step_size=40, mode='triangular')
        for it in (0, 7, 40, 63, 80, 131):
            clr.clr_iterations = float(it)
            assert abs(float(clr.clr()) - T.clr_triangular(it, 5e-5, 1e-2, 40)) < 1e-15


def test_play_checkers_move_listing_and_record_q():
    """host-side pieces of the console game and of the playout-mode records: the (from, to) listing of legal
    moves for plain moves, jumps and promotions of either side (play_Checkers.py:62-84), and the exact
    int / int quotient the reference records as q when NEURAL_NET=False"""
    import play_Checkers as P
    rng = np.random.RandomState(4)
    seen_jump = seen_promo = False
    for _game in range(30):
        pos = O.start_position()
        for _ply in range(90):
            kids, _mask, status, _p5 = O.movegen(pos)
            if status != 0:
                break
            state = codec.decode_state(pos)
            listing = P.states_to_piece_positions(state, [codec.decode_state(k) for k in kids])
            assert len(listing) == len(kids)
            mover = (0, 1) if state[4, 0, 0] == 0 else (2, 3)
            for (src, dst), kid in zip(listing, kids):
                ks = codec.decode_state(kid)
                (sx, sy), (dx, dy) = (src[0] - 1, src[1] - 1), (dst[0] - 1, dst[1] - 1)
                assert state[mover[0], sx, sy] + state[mover[1], sx, sy] == 1       # a piece of the mover stood there
                assert ks[mover[0], dx, dy] + ks[mover[1], dx, dy] == 1             # and stands here now
                assert ks[mover[0], sx, sy] + ks[mover[1], sx, sy] == 0
                assert abs(dx - sx) == abs(dy - sy) and abs(dx - sx) in (1, 2)
                seen_jump |= abs(dx - sx) == 2
                seen_promo |= bool(state[mover[0], sx, sy] == 1 and ks[mover[1], dx, dy] == 1)
            pos = kids[rng.randint(len(kids))]
    assert seen_jump and seen_promo
    rec = np.zeros(1, dtype=[("pos", np.uint32, 4), ("mask", np.uint32, 8), ("plane5", np.int32), ("n_children", np.int32),
                             ("action", np.uint16, 48), ("visits", np.uint32, 48), ("q", np.float32), ("z", np.int32),
                             ("root_n", np.uint32), ("root_w", np.float32)])[0]
    rec["pos"] = O.start_position()
    rec["n_children"], rec["action"][:2], rec["visits"][:2] = 2, (149, 151), (70, 80)
    rec["root_n"], rec["root_w"], rec["q"] = 150, -7.0, np.float32(7.0) / np.float32(150.0)     # flipped to the root player's view
    assert R.to_reference(rec, playouts=True)[2] == 7 / 150 and type(R.to_reference(rec, playouts=True)[2]) is float
    assert R.to_reference(rec)[2] == np.float32(7.0) / np.float32(150.0)


def test_multiproc_playouts_fail_like_the_reference():
    """MULTIPROC=True with NEURAL_NET=False: the reference's first simulation raises a TypeError (MCTS.py:83-87 feeds
    pool.map's tuples to backpropagation; observed with the unmodified reference), the shim raises the same before
    it touches the device"""
    import pytest
    import MCTS as M
    kw = dict(GAME_ENV=None, UCT_C=4, CONSTRAINT='rollout', BUDGET=4, MULTIPROC=True, NEURAL_NET=False, VERBOSE=False,
              TRAINING=False, DIRICHLET_ALPHA=1.0, DIRICHLET_EPSILON=0.25, TEMPERATURE_TAU=0, TEMPERATURE_DECAY=0,
              TEMP_DECAY_DELAY=0)
    M.MCTS(**kw)
    try:
        with pytest.raises(TypeError, match="'int' and 'tuple'"):
            M.MCTS.begin_tree_search(object())
    finally:
        M.MCTS(**dict(kw, MULTIPROC=False))


def test_packed_records_round_trip():
    """ckb200.records.pack / unpack (numpy twins of the device-side pack kernel and of what rank 0 does after the
    gather): bit-identical round trip on real game records, including a terminal record that carries its legal-action
    planes (a draw by the 80-ply rule still has legal moves, Checkers.py:344-360)"""
    from ckb200 import lib_types as T
    from ckb200 import records as R
    recs = []
    for g in range(4):
        gm = O.Game(O.make_cfg(budget=24, training=True, terminate_cnt=0 if g % 2 else 60, epsilon=0.25, tau=1.0, tau_decay=0.1,
                               tau_decay_delay=10, seed=g), "hash_salted", None, salt=g)
        gm.play()
        recs.append(T.records_from_dicts(gm.records(), game=g))
        gm.close()
    recs = np.concatenate(recs)
    draw = np.zeros(1, dtype=T.RECORD_DTYPE)                       # synthetic draw-terminal record: n_children 0, mask set
    draw["pos"] = (0x00000001, 0x80000000, 0x80000001, 1 | (79 << 1) | (200 << 18))
    draw["mask"][0] = (1, 0, 0, 0, 0, 0, 0, 0)
    draw["plane5"], draw["chosen"], draw["game"], draw["ply"], draw["q"] = 80, -1, 9, 120, 0.0
    recs = np.concatenate([recs[:7], draw, recs[7:]])
    hdr, words = R.pack(recs)
    assert hdr.dtype.itemsize == 40 and int(hdr["flags"].sum()) == 1
    assert len(words) == int(recs["n_children"].sum()) + 8
    back = R.unpack(hdr, words)
    assert back.tobytes() == recs.tobytes()
    assert hdr.nbytes + words.nbytes < recs.nbytes / 4
    try:
        R.unpack(hdr, words[:-1])
        raise AssertionError("truncated word stream accepted")
    except ValueError:
        pass
    # the library's host-side unpack (ck_records_unpack, what generate_data() and the pooled gather call): same bytes,
    # same error; also on a batch large enough for its worker threads, and on nothing at all
    from ckb200 import lib as L
    assert L.records_unpack(hdr, words).tobytes() == recs.tobytes()
    big = np.concatenate([recs] * 2500)                             # > 2 x 65536 records
    bh, bw = R.pack(big)
    assert L.records_unpack(bh, bw).tobytes() == big.tobytes()
    assert len(L.records_unpack(hdr[:0], words[:0])) == 0
    for bad in (words[:-1], np.concatenate([words, words[:1]])):
        try:
            L.records_unpack(hdr, bad)
            raise AssertionError("wrong word count accepted")
        except ValueError as e:
            assert "child words expected" in str(e)


def test_keras_h5_export_through_a_reference_template(tmp_path):
    """weights written back as a Keras .h5 (reference training_pipeline.py:186-191 saves .h5): the HDF5 structure,
    model_config and training configuration of a file the reference itself saved are kept byte for byte, the weight
    payloads are replaced, the stored optimizer state is zeroed.  Read back with the importer, blob for blob."""
    import pytest
    d = "/root/reference/data/model"
    if not os.path.isdir(d):
        pytest.skip("reference models not mounted")
    from ckb200 import h5lite
    fns = sorted(f for f in os.listdir(d) if f.endswith(".h5"))
    template = os.path.join(d, [f for f in fns if "Model5_" in f][0])
    other = h5lite.keras_h5_to_blob(os.path.join(d, [f for f in fns if "Model10_" in f][0]))
    blob = (other * np.float32(1.25) + np.float32(0.001)).astype(np.float32)       # weights that are in no shipped file
    out = h5lite.blob_to_keras_h5(blob, template, str(tmp_path / "Checkers_Model11_test.h5"))
    assert os.path.getsize(out) == os.path.getsize(template)
    assert h5lite.keras_h5_to_blob(out).tobytes() == blob.tobytes()
    a, b = open(template, "rb").read(), open(out, "rb").read()
    assert h5lite.model_config(a) == h5lite.model_config(b)
    # nothing but dataset payloads changed: the bytes that differ lie inside float datasets
    da, _ = h5lite.read_datasets(out, prefix="optimizer_weights")
    assert da and all((v == 0).all() for v in da.values())
    diff = np.flatnonzero(np.frombuffer(a, np.uint8) != np.frombuffer(b, np.uint8))
    assert 0 < len(diff) <= 3 * 4 * N.NET_PARAM_COUNT
    # save_nn_to_disk writes it next to the .npy when a template is named
    import training_pipeline as TP
    from ckb200 import train as T
    os.makedirs(tmp_path / "data" / "model")
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        fn = TP.save_nn_to_disk(T.CheckersNet(blob), 12, "01-Jan-2026(00:00:00)", h5_template=template)
        assert fn.endswith(".npy") and h5lite.keras_h5_to_blob(fn[:-4] + ".h5").tobytes() == blob.tobytes()
    finally:
        os.chdir(cwd)


def test_bench_reference_arm_line():
    """`bench.py --impl reference` (CPU only): the reference's own self-play code timed on the host cores prints the
    contract's JSON line -- same metric / unit / config as the CUDA arm, `impl`, `cpu_baseline` and a zero-copy `e2e`."""
    import json
    import subprocess
    import sys
    from oracle import build_ref
    from oracle import ref_harness as H
    if not (H.reference_available() or build_ref.available()):
        import pytest
        pytest.skip("neither the reference tree nor its compiled copy (oracle/_ref) is present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mcts_sims_per_sec" and d["unit"] == "sims/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "4096 concurrent self-play games" in d["config"]["workload"] and d["gpu_launches"] == 0
