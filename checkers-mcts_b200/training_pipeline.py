"""Drop-in for the self-play / tournament entry points of the reference's ``training_pipeline``
(reference training_pipeline.py:310-600) on the B200 engine.

``generate_Checkers_data(selfplay_kwargs, mcts_kwargs).generate_data()`` and
``tournament_Checkers(tourney_kwargs, mcts_kwargs).start_tournament()`` keep the reference's
constructor arguments, file names and file formats (pickled ``[state, probs, q, z]`` lists,
``tabulate`` fancy_grid tournament tables), so ``train_Checkers.py`` can call them unchanged.
What differs is where the work happens: all games of all ``NUM_CPUS`` "workers" run concurrently
on the GPU (one warp per game, batched network evaluation) instead of one game per process.

Networks: ``NN_FN`` / ``NEW_NN_FN`` / ``OLD_NN_FN`` name a weight blob saved by ``save_blob``
(``.npy``, 1,321,774 float32 in the layout of include/ckb200.h).  Keras ``.h5`` import is a
"next" row; asking for one raises with that explanation rather than guessing.
"""
import os
import pickle
from datetime import datetime

import numpy as np
from tabulate import tabulate

from ckb200 import dist as _D
from ckb200 import lib as _L
from ckb200 import net as _N
from ckb200 import records as _R


def create_timestamp():
    return datetime.now(tz=None).strftime("%d-%b-%Y(%H:%M:%S)")


def save_blob(blob, filename):
    np.save(filename, np.asarray(blob, dtype=np.float32))
    return filename if filename.endswith(".npy") else filename + ".npy"


def load_blob(filename):
    """-> float32 weight blob, or an evaluator name for 'stub:<kind>' (tests / demos)"""
    if isinstance(filename, str) and filename.startswith("stub:"):
        return filename[5:]
    if isinstance(filename, str) and filename.endswith(".h5"):
        from ckb200 import h5lite                       # the reference's own model files (model.save(...h5), :186-191)
        return h5lite.keras_h5_to_blob(filename)
    blob = np.load(filename)
    if blob.size != _N.NET_PARAM_COUNT:
        raise ValueError("weight blob %s has %d values, expected %d" % (filename, blob.size, _N.NET_PARAM_COUNT))
    return blob.astype(np.float32).reshape(-1)


def load_model(filename):
    """Stand-in for ``tensorflow.keras.models.load_model`` in the training phase (train_Checkers.py:163): the saved
    network (.npy blob or the reference's .h5) as a trainable ``ckb200.train.CheckersNet``."""
    from ckb200 import train as T
    return T.CheckersNet(load_blob(filename))


def load_training_data(filename):
    with open(filename, 'rb') as file:
        return pickle.load(file)


def create_nn(**kwargs):
    """Randomly initialised policy/value network with the reference's architecture and Keras default
    initialisers (reference :44-120).  Returns a ``ckb200.train.CheckersNet`` (PyTorch); regularisation and
    loss weights are taken from the kwargs again by ``train_nn``."""
    from ckb200 import train as T
    if kwargs.get('NUM_KERNELS', 128) != 128:
        raise ValueError('the device evaluator is built for NUM_KERNELS = 128')
    return T.CheckersNet(seed=kwargs.get('SEED', 0))


def train_nn(training_data, neural_network, **kwargs):
    """Train with the reference's recipe (reference :123-179): Adam + triangular CLR per batch, early stopping
    and best-epoch checkpoint on val_loss.  ``neural_network`` is a CheckersNet, a weight blob or a weight file
    (.npy / the reference's .h5).  Returns (history, filename of the best epoch's weights)."""
    from ckb200 import train as T
    if isinstance(neural_network, str):
        neural_network = load_blob(neural_network)
    return T.train_nn(training_data, neural_network, Keras_Generator, create_timestamp(), **kwargs)


def set_nn_lrate(neural_network, lrate):
    """Kept for API compatibility (reference :181-184): the learning rate is set per batch by the CLR schedule."""
    neural_network.base_lr = lrate


def save_nn_to_disk(neural_network, iteration, timestamp, h5_template=None):
    """Save the network with timestamp and iteration in the file name (reference :186-191).  The weights are written as
    a .npy blob, which every NN_FN consumer of this package accepts.  With ``h5_template`` (or the environment variable
    CK_H5_TEMPLATE) naming any model file the reference saved (data/model/Checkers_Model*.h5), a Keras ``.h5`` the
    reference's own ``load_model`` reads is written next to it (``ckb200.h5lite.blob_to_keras_h5``: the template's HDF5
    structure with the weight payloads replaced; no h5py / TensorFlow needed)."""
    filename = 'data/model/Checkers_Model' + str(iteration) + '_' + timestamp + '.npy'
    out = neural_network.save(filename)
    h5_template = h5_template or os.environ.get('CK_H5_TEMPLATE')
    if h5_template:
        from ckb200 import h5lite
        h5lite.blob_to_keras_h5(neural_network.blob(), h5_template, filename[:-4] + '.h5')
    return out


def plot_history(history, nn, TRAINING_ITERATION):
    """The reference plots loss per epoch with matplotlib (:199-217); without it the curve is written as text."""
    filename = 'data/plots/Checkers_Model' + str(TRAINING_ITERATION + 1) + '_TrainingLoss_' + create_timestamp() + '.txt'
    import os
    os.makedirs(os.path.dirname(filename), exist_ok=True)
    with open(filename, 'w') as file:
        file.write('epoch loss val_loss lr\n')
        for i, loss in enumerate(history['loss']):
            val = history['val_loss'][i] if i < len(history.get('val_loss', [])) else float('nan')
            file.write('%d %.6f %.6f %.3e\n' % (i + 1, loss, val, history['lr'][i]))
    return filename


def save_merged_files(memory, iteration, timestamp):
    """Save merged training data as one pickle (reference :269-275)."""
    filename = 'data/training_data/Checkers_Data' + str(iteration) + '_' + timestamp + '.pkl'
    with open(filename, 'wb') as file:
        pickle.dump(memory, file)
    return filename


def merge_data(data_fns, iteration):
    """Merge the per-worker self-play files of one iteration into a single file (reference :277-284);
    like the reference it returns the merged list, the file name only goes to disk."""
    training_data = []
    for fn in data_fns:
        training_data.extend(load_training_data('data/training_data/' + fn))
    save_merged_files(training_data, iteration, create_timestamp())
    return training_data


class Keras_Generator(object):
    """Batch view of a list of self-play records ``[state[15,8,8], probs[8,8,8], q, z]`` with the layout
    the reference feeds to ``fit`` (:288-307): features = planes 0..13 channels-last ``[B,8,8,14]``,
    labels = ``[visit probabilities flattened to 512, (q + z) / 2]``.  Indexable and iterable; it does
    not subclass ``keras.utils.Sequence`` (TensorFlow is not a dependency) but has the same ``len`` /
    ``__getitem__`` contract, so it can be handed to a Keras or PyTorch training loop unchanged."""

    def __init__(self, data, batch_size):
        self.data = data
        self.batch_size = batch_size

    def __len__(self):
        return int(np.ceil(len(self.data) / float(self.batch_size)))

    def __getitem__(self, idx):
        if idx < 0 or idx >= len(self):
            raise IndexError(idx)
        data = self.data[idx * self.batch_size: (idx + 1) * self.batch_size]
        states = np.array([e[0][:14] for e in data])
        states = np.moveaxis(states, 1, -1)
        probs = np.array([np.array(e[1]).flatten() for e in data])
        qvals = np.array([e[2] for e in data])
        zvals = np.array([e[3] for e in data])
        return (states, [probs, (qvals + zvals) / 2])

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


def record_params(phase, **kwargs):
    """Document the parameters used in the training pipeline (reference :225-244)."""
    folders = {'selfplay': 'data/training_data/Checkers_SelfPlay_Params_', 'training': 'data/model/Checkers_Training_Params_',
               'evaluation': 'data/tournament_results/Checkers_Evaluation_Params_',
               'final': 'data/final_eval/Checkers_Final_Evaluation_Params_'}
    if phase not in folders:
        raise ValueError('Invalid phase!')
    filename = folders[phase] + create_timestamp() + '.txt'
    with open(filename, 'w') as file:
        for key, val in kwargs.items():
            file.write('{} = {}\n'.format(key, val))
    return filename


def _engine_cfg(mcts_kwargs, n_slots, terminate_cnt, evaluator, evaluator_p2=None, arena=False, keep_records=True,
                device=0, seed=None, game_id_base=0, game_id_stride=1):
    if mcts_kwargs.get('CONSTRAINT', 'rollout') != 'rollout':
        raise ValueError("the device engine only supports CONSTRAINT='rollout'")
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little")          # np.random.seed() from OS entropy (:341)
    return _L.make_cfg(n_slots=n_slots, budget=mcts_kwargs['BUDGET'], device=device, uct_c=mcts_kwargs['UCT_C'],
                       training=mcts_kwargs['TRAINING'], alpha=mcts_kwargs['DIRICHLET_ALPHA'],
                       epsilon=mcts_kwargs['DIRICHLET_EPSILON'], tau=mcts_kwargs['TEMPERATURE_TAU'],
                       tau_decay=mcts_kwargs['TEMPERATURE_DECAY'], tau_decay_delay=mcts_kwargs['TEMP_DECAY_DELAY'],
                       terminate_cnt=terminate_cnt, seed=seed, evaluator=evaluator, evaluator_p2=evaluator_p2,
                       arena=arena, keep_records=keep_records, game_id_base=game_id_base, game_id_stride=game_id_stride)


def selfplay_workers(num_cpus):
    """host processes that write the per-worker pickle files: NUM_CPUS of them, bounded by the host's cores"""
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    return max(1, min(int(num_cpus), cores))


def _np_empty(dtype):
    return np.zeros(0, dtype=dtype)


def _dist_setup(device, seed):
    """(rank, world, CUDA device of this process, seed shared by all ranks)"""
    rank, world, local_rank = _D.rank_world()
    if world > 1:
        device = local_rank
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little") >> 1      # np.random.seed() from OS entropy (:341)
    if world > 1:
        seed = _D.broadcast_int(seed, rank, world, device="cuda:%d" % device)
    return rank, world, device, seed


def _attach(engine, which, spec, device):
    if isinstance(spec, str):
        return None
    net = _L.Net(device)
    net.set_weights(spec)
    engine.set_net(which, net)
    return net


class generate_Checkers_data(object):
    """Self-play data generation (reference :310-469)."""

    def __init__(self, selfplay_kwargs, mcts_kwargs):
        self.NUM_SELFPLAY_GAMES = selfplay_kwargs['NUM_SELFPLAY_GAMES']
        self.TRAINING_ITERATION = selfplay_kwargs['TRAINING_ITERATION']
        self.TERMINATE_CNT = selfplay_kwargs['TERMINATE_CNT']
        self.num_cpus = selfplay_kwargs['NUM_CPUS']       # number of "workers": total games = games x workers
        self.nn_fn = selfplay_kwargs['NN_FN']
        self.device = selfplay_kwargs.get('DEVICE', 0)
        self.max_slots = selfplay_kwargs.get('MAX_CONCURRENT_GAMES', 4096)
        self.seed = selfplay_kwargs.get('SEED')
        # under torchrun the games normally shard over the ranks; SINGLE_PROCESS keeps this call local to its process
        self.single_process = bool(selfplay_kwargs.get('SINGLE_PROCESS', False))
        self.mcts_kwargs = mcts_kwargs
        self.stats = None
        self.n_records, self.host_seconds = 0, 0.0

    def generate_data(self):
        """plays NUM_SELFPLAY_GAMES x NUM_CPUS games on the GPU; returns the list of pickle files,
        one per worker as in the reference (a single name when NUM_CPUS == 1)"""
        # NEURAL_NET=False is the reference's iteration-0 mode (train_Checkers.py:78): plain UCT with random
        # playouts, no network involved (the reference still loads NN_FN there and never calls it)
        playouts = not self.mcts_kwargs.get('NEURAL_NET', True)
        total = self.NUM_SELFPLAY_GAMES * self.num_cpus
        # under torchrun (one process per GPU) the games shard by index over the ranks, game g on rank g mod world,
        # every rank plays its share with the same seed (a game's random stream depends on its global index only)
        # and rank 0 pools the records: the files are the same as from a single process
        if self.single_process:
            rank, world, device = 0, 1, self.device
            seed = self.seed if self.seed is not None else int.from_bytes(os.urandom(8), "little") >> 1
        else:
            rank, world, device, seed = _dist_setup(self.device, self.seed)
        spec = self.mcts_kwargs.get('PLAYOUT_EVALUATOR', 'rollout') if playouts else load_blob(self.nn_fn)
        base, stride, n_local = _D.shard(total, rank, world)
        recs, games = _np_empty(_L.RECORD_DTYPE), _np_empty(_L.GAME_DTYPE)
        import time as _time
        _t_host = _time.time()
        if n_local > 0:
            cfg = _engine_cfg(self.mcts_kwargs, min(n_local, self.max_slots), self.TERMINATE_CNT,
                              spec if isinstance(spec, str) else "net", device=device, seed=seed,
                              game_id_base=base, game_id_stride=stride)
            eng, net = _L.Engine(cfg), None
            try:
                net = _attach(eng, 0, spec, device)
                self.stats = eng.selfplay(n_local)
                _t_host = _time.time()                    # everything from here on is host-side work on finished games
                games = eng.games()
                if world > 1:                             # packed on the device, exact-size NCCL transfers to rank 0
                    recs, _ms = _D.gather_engine_records(eng, rank, world, "cuda:%d" % device)
                else:
                    recs = _L.records_unpack(*eng.records_packed())     # a fifth of the bytes of the full records over PCIe
            finally:                                      # device memory goes back also when a run fails
                eng.close()
                if net is not None:
                    net.close()
        elif world > 1:
            raise ValueError('NUM_SELFPLAY_GAMES x NUM_CPUS must be at least the number of ranks')
        if world > 1:
            games = _D.gather_records(games, rank, world, device="cuda:%d" % device)
            if rank != 0:
                return []
        recs = recs[np.argsort(recs["game"], kind="stable")]       # game by game, plies in order
        games = games[np.argsort(games["game"], kind="stable")]
        names = {1: 'player1_wins', 2: 'player2_wins', 3: 'draw'}
        for g in games:
            print('{} after {} moves!'.format(names[int(g["outcome"])], int(g["move_count"])))
        timestamp = create_timestamp()
        # worker p played games [p*N, (p+1)*N) (:326-329); its file is converted to the reference's list format and
        # pickled by a pool process (the conversion is numpy over whole record arrays, see ckb200.records)
        bounds = np.searchsorted(recs["game"], np.arange(self.num_cpus + 1) * self.NUM_SELFPLAY_GAMES)
        jobs = [(recs[bounds[p]:bounds[p + 1]], self._filename(self.TRAINING_ITERATION, timestamp, p)) for p in range(self.num_cpus)]
        filenames = _R.save_reference_pickles(jobs, playouts, workers=selfplay_workers(self.num_cpus))
        self.n_records = int(len(recs))
        self.host_seconds = _time.time() - _t_host        # sorting, conversion to the reference's lists, pickling
        return filenames if self.num_cpus > 1 else filenames[0]

    @staticmethod
    def _filename(iteration, timestamp, process_num):
        return 'data/training_data/Checkers_Data' + str(iteration) + '_' + timestamp + '_P' + str(process_num) + '.pkl'

    def _save_memory(self, memory, iteration, timestamp, process_num):
        filename = self._filename(iteration, timestamp, process_num)
        with open(filename, 'wb') as file:
            pickle.dump(memory, file)
        return filename


class tournament_Checkers(object):
    """New-net versus old-net arena (reference :472-600)."""

    def __init__(self, tourney_kwargs, mcts_kwargs):
        self.nn1_fn = tourney_kwargs['NEW_NN_FN']
        self.nn2_fn = tourney_kwargs['OLD_NN_FN']
        self.NUM_GAMES = tourney_kwargs['TOURNEY_GAMES']
        self.num_cpus = tourney_kwargs['NUM_CPUS']
        self.device = tourney_kwargs.get('DEVICE', 0)
        self.seed = tourney_kwargs.get('SEED')
        self.mcts_kwargs = mcts_kwargs
        self.stats = None

    def _start_tournament(self, process_num=0):
        """TOURNEY_GAMES games, the new net is player 1 in the first half (:523-528); rows
        [game#, p1_file, p2_file, outcome, move_count] (:552-553)"""
        s1, s2 = load_blob(self.nn1_fn), load_blob(self.nn2_fn)
        ev1 = s1 if isinstance(s1, str) else "net"
        ev2 = s2 if isinstance(s2, str) else "net"
        rank, world, device, seed = _dist_setup(self.device, None if self.seed is None else self.seed + process_num)
        if self.NUM_GAMES % world:
            raise ValueError('TOURNEY_GAMES must be a multiple of the number of ranks')
        base, stride, n_local = _D.shard(self.NUM_GAMES, rank, world)       # games shard over the ranks as in self-play
        cfg = _engine_cfg(self.mcts_kwargs, min(n_local, 4096), 0, ev1, ev2, arena=True, keep_records=False,
                          device=device, seed=seed, game_id_base=base, game_id_stride=stride)
        eng, nets = _L.Engine(cfg), []
        try:
            nets.append(_attach(eng, 0, s1, device))
            nets.append(_attach(eng, 1, s2, device))
            self.stats = eng.arena(n_local)
            games = eng.games()
        finally:
            eng.close()
            for n in nets:
                if n is not None:
                    n.close()
        if world > 1:
            games = _D.gather_records(games, rank, world, device="cuda:%d" % device)
            if rank != 0:
                return []
        games = games[np.argsort(games["game"], kind="stable")]
        names = {1: 'player1_wins', 2: 'player2_wins', 3: 'draw'}
        rows = []
        for g in games:
            p1, p2 = (self.nn1_fn, self.nn2_fn) if int(g["p1_net"]) == 0 else (self.nn2_fn, self.nn1_fn)
            # the reference strips the model folder from the names it records (:550-553)
            rows.append([int(g["game"]) + 1, p1.replace('data/model/', ''), p2.replace('data/model/', ''),
                         names[int(g["outcome"])], int(g["move_count"])])
        return rows

    def start_tournament(self):
        game_outcomes = []
        for proc in range(self.num_cpus):                 # each worker plays its own TOURNEY_GAMES (:491-494)
            game_outcomes.extend(self._start_tournament(proc))
        if not game_outcomes:                             # a rank other than 0 of a multi-GPU run
            return None
        filename = self._save_tourney_results(game_outcomes)
        print('Tournament over!  View results in tournament folder!')
        return filename

    def _save_tourney_results(self, game_outcomes):
        fn1, fn2 = game_outcomes[0][1], game_outcomes[0][2]
        wins = {fn1: 0, fn2: 0}
        draws = 0
        for idx, row in enumerate(game_outcomes):
            row[0] = idx + 1
            if row[3] == 'player1_wins':
                wins[row[1]] += 1
            elif row[3] == 'player2_wins':
                wins[row[2]] += 1
            else:
                draws += 1
        summary = [[fn1, '{}/{}/{}'.format(wins[fn1], wins[fn2], draws)], [fn2, '{}/{}/{}'.format(wins[fn2], wins[fn1], draws)]]
        filename = 'data/tournament_results/Tournament_' + create_timestamp() + '.txt'
        with open(filename, 'w') as file:
            file.write(tabulate(summary, tablefmt='fancy_grid', headers=['Neural Network', 'Wins/Losses/Draws']))
            file.write('\n\n')
            file.write(tabulate(game_outcomes, tablefmt='fancy_grid',
                                headers=['Game Number', 'Player 1', 'Player 2', 'Outcome', 'Turn Count']))
        return filename


class final_evaluation(object):
    """Round-robin among trained models (reference :603-719): every model plays every other model twice (once
    as each colour); a win scores +1 for the winner and -1 for the loser in the pairing table, the row sums are
    the models' points.  Results go to data/final_eval/ as a fancy_grid table plus the points per model
    iteration as text (the reference plots them with matplotlib)."""

    def __init__(self, model_iter_list, tourney_kwargs, mcts_kwargs):
        self.model_iter_list = list(model_iter_list)
        self.model_fn_list = []
        self.tourney_kwargs = dict(tourney_kwargs)
        self.mcts_kwargs = mcts_kwargs
        self.num_cpus = tourney_kwargs['NUM_CPUS']
        self.tourney_kwargs['TOURNEY_GAMES'] = 2
        fns = sorted(os.listdir('data/model'))
        for iter_num in self.model_iter_list:
            for fn in fns:
                if 'Model' + str(iter_num) + '_' in fn and (fn.endswith('.h5') or fn.endswith('.npy')):
                    self.model_fn_list.append(fn)
                    break
        if len(self.model_fn_list) != len(self.model_iter_list):
            raise ValueError('Model(s) not found!')
        self.table = np.zeros((len(self.model_iter_list), len(self.model_iter_list)))
        self.game_outcomes = []

    def start_evaluation(self, num_cpus=1):
        """Every pairing is one two-game arena (training_pipeline.py:632-658).  The reference fans the pairings of one
        "new" model out over ``num_cpus`` processes; here up to ``max(num_cpus, 16)`` pairings run CONCURRENTLY on the
        GPU, each in its own engine on its own stream (the library calls release the GIL), because a two-game arena by
        itself is pure launch latency.  Under torchrun the pairings run one after the other (their games shard over
        the ranks and the collectives must stay ordered)."""
        model_fn_list = self.model_fn_list.copy()
        pairings = []                                     # (group index, new, old) in the reference's order
        for grp in range(len(self.model_fn_list) - 1):
            new_nn_fn = model_fn_list.pop()
            pairings += [(grp, new_nn_fn, old_nn_fn) for old_nn_fn in model_fn_list]
        world = _D.rank_world()[1]
        if world == 1 and len(pairings) > 1:
            import concurrent.futures as cf
            with cf.ThreadPoolExecutor(max_workers=min(len(pairings), max(int(num_cpus), 16))) as pool:
                results = list(pool.map(lambda p: self._wrapper_func(p[1], p[2]), pairings))
        else:
            results = [self._wrapper_func(p[1], p[2]) for p in pairings]
        groups = [[] for _ in range(len(self.model_fn_list) - 1)]
        for (grp, _new, _old), rows in zip(pairings, results):
            groups[grp].extend(rows)
        self.game_outcomes.extend(groups)
        if _D.rank_world()[0] != 0:                       # rank 0 holds the pooled results of a multi-GPU run
            return None
        filename = self._parse_tourney_results()
        print('Final evaluation over!  View results in final_eval folder!')
        return filename

    def _wrapper_func(self, new_nn_fn, old_nn_fn):
        tourney_kwargs = dict(self.tourney_kwargs, NUM_CPUS=1, NEW_NN_FN='data/model/' + new_nn_fn,
                              OLD_NN_FN='data/model/' + old_nn_fn)
        mcts_kwargs = dict(self.mcts_kwargs, NN_FN=tourney_kwargs['NEW_NN_FN'])
        print('Beginning tournament between {} and {}!'.format(tourney_kwargs['NEW_NN_FN'], tourney_kwargs['OLD_NN_FN']))
        return tournament_Checkers(tourney_kwargs, mcts_kwargs)._start_tournament()

    def _parse_tourney_results(self):
        for game_outcomes in self.game_outcomes:
            for _game_num, p1_fn, p2_fn, outcome, _move_count in game_outcomes:
                p1_idx = self.model_fn_list.index(p1_fn)
                p2_idx = self.model_fn_list.index(p2_fn)
                if outcome == 'player1_wins':
                    self.table[p1_idx, p2_idx] += 1
                    self.table[p2_idx, p1_idx] -= 1
                elif outcome == 'player2_wins':
                    self.table[p1_idx, p2_idx] -= 1
                    self.table[p2_idx, p1_idx] += 1
        model_scores = np.sum(self.table, axis=1)
        self.model_scores = model_scores
        stamp = create_timestamp()
        os.makedirs('data/final_eval', exist_ok=True)
        with open('data/final_eval/Checkers_Final_Evaluation_' + stamp + '_points.txt', 'w') as file:
            for it, score in zip(self.model_iter_list, model_scores):
                file.write('{} {}\n'.format(it, score))
        table = np.hstack((self.table, np.transpose(model_scores[np.newaxis])))
        filename = 'data/final_eval/Checkers_Final_Evaluation_' + stamp + '.txt'
        with open(filename, 'w') as file:
            file.write(tabulate(table, headers=self.model_iter_list + ['Total'], showindex=self.model_iter_list,
                                tablefmt='fancy_grid'))
        return filename
