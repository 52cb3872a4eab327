#!/usr/bin/env python3
"""One iteration of the training pipeline on the device engine -- the counterpart of the reference's
``train_Checkers.py`` (train_Checkers.py:33-214): self-play -> training -> evaluation tournament, plus the
final round-robin among the trained models.  The phase switches, the kwargs dictionaries and their default
values are the reference's; they are arguments of ``run_iteration`` (and command-line flags) instead of
constants edited in the file.

    python train_Checkers.py --iteration 0 --selfplay --training            # UCT + random playouts, first net
    python train_Checkers.py --iteration 1 --nn-fn data/model/Checkers_Model1_....npy --selfplay --training --evaluation
    python train_Checkers.py --final-evaluation 0 1 2 3

Differences forced by the platform: self-play and tournaments run on the GPU (``NUM_CPUS`` keeps its meaning
"number of workers": total games = games x workers); the learning-rate finder and the matplotlib plots of the
reference are not part of this package (``plot_history`` writes the curve as text); networks are saved as
``.npy`` weight blobs, and every ``*_NN_FN`` also accepts the reference's Keras ``.h5`` files.
"""
import argparse
import os

from training_pipeline import (create_nn, create_timestamp, final_evaluation, generate_Checkers_data, load_model,
                               load_training_data, merge_data, plot_history, record_params, save_nn_to_disk,
                               tournament_Checkers, train_nn)


def default_selfplay_kwargs(iteration, nn_fn):          # train_Checkers.py:80-86
    return {'TRAINING_ITERATION': iteration, 'NN_FN': nn_fn, 'NUM_SELFPLAY_GAMES': 100, 'TERMINATE_CNT': 200,
            'NUM_CPUS': 1}


def default_mcts_kwargs(iteration):                     # train_Checkers.py:78,88-102
    return {'GAME_ENV': None, 'UCT_C': 4, 'CONSTRAINT': 'rollout', 'BUDGET': 200, 'MULTIPROC': False,
            'NEURAL_NET': iteration != 0,               # random rollouts generate the first dataset
            'VERBOSE': False, 'TRAINING': True, 'DIRICHLET_ALPHA': 1.0, 'DIRICHLET_EPSILON': 0.25,
            'TEMPERATURE_TAU': 1.0, 'TEMPERATURE_DECAY': 0.1, 'TEMP_DECAY_DELAY': 10}


def default_training_kwargs(iteration):                 # train_Checkers.py:111-127
    return {'TRAINING_ITERATION': iteration, 'NN_BASE_LR': 5e-5, 'NN_MAX_LR': 1e-2, 'CLR_SS_COEFF': 4,
            'BATCH_SIZE': 128, 'EPOCHS': 100, 'CONV_REG': 0.001, 'DENSE_REG': 0.001, 'NUM_KERNELS': 128,
            'VAL_SPLIT': 0.20, 'MIN_DELTA': 0.01, 'PATIENCE': 20, 'POLICY_LOSS_WEIGHT': 1.0,
            'VALUE_LOSS_WEIGHT': 1.0, 'SLIDING_WINDOW': 1}


def default_tourney_kwargs(iteration, old_nn_fn, new_nn_fn):    # train_Checkers.py:179-185
    return {'TRAINING_ITERATION': iteration, 'OLD_NN_FN': old_nn_fn, 'NEW_NN_FN': new_nn_fn, 'TOURNEY_GAMES': 2,
            'NUM_CPUS': 5}


def default_tourney_mcts_kwargs(new_nn_fn):             # train_Checkers.py:187-201
    return {'NN_FN': new_nn_fn, 'UCT_C': 4, 'CONSTRAINT': 'rollout', 'BUDGET': 200, 'MULTIPROC': False,
            'NEURAL_NET': True, 'VERBOSE': False, 'TRAINING': False, 'DIRICHLET_ALPHA': 1.0,
            'DIRICHLET_EPSILON': 0.25, 'TEMPERATURE_TAU': 0, 'TEMPERATURE_DECAY': 0, 'TEMP_DECAY_DELAY': 0}


def _window_data(iteration, sliding_window):
    """training data of the last SLIDING_WINDOW iterations found in data/training_data (:146-157)"""
    start = max(0, iteration + 1 - sliding_window)
    fns = sorted(os.listdir('data/training_data'))
    data = []
    for it in range(start, iteration + 1):
        for fn in fns:
            if 'Data' + str(it) + '_' in fn:
                data.extend(load_training_data('data/training_data/' + fn))
    return data


def run_iteration(TRAINING_ITERATION, NN_FN=None, NEW_NN_FN=None, SELFPLAY=True, TRAINING=False, EVALUATION=False,
                  selfplay_kwargs=None, mcts_kwargs=None, training_kwargs=None, tourney_kwargs=None,
                  tourney_mcts_kwargs=None):
    """The body of train_Checkers.py (:75-208).  The ``*_kwargs`` arguments override single entries of the
    reference's defaults.  Returns a dict with the files the phases wrote."""
    for d in ('data/training_data', 'data/model', 'data/tournament_results', 'data/plots'):
        os.makedirs(d, exist_ok=True)
    out = {}
    sp = dict(default_selfplay_kwargs(TRAINING_ITERATION, NN_FN), **(selfplay_kwargs or {}))
    mk = dict(default_mcts_kwargs(TRAINING_ITERATION), **(mcts_kwargs or {}))
    data_fns = None
    # under torchrun every rank plays its share of the games; files, training and parameter records are rank 0's
    # business (the other ranks wait at the barriers and take the new network's file name from rank 0)
    from ckb200 import dist as _D
    rank, world, _local = _D.rank_world()
    if SELFPLAY:
        data_fns = generate_Checkers_data(sp, mk).generate_data()
        if rank == 0:
            record_params('selfplay', **{**sp, **mk})
        out['data_fns'] = data_fns
    _barrier(world)
    tk = dict(default_training_kwargs(TRAINING_ITERATION), **(training_kwargs or {}))
    if TRAINING and rank != 0:
        TRAINING = False                                  # rank 0 trains; the result is broadcast below
        _names = _D.broadcast_strings([None, None], rank, world)
        NN_FN, NEW_NN_FN = _names
        out.update(OLD_NN_FN=NN_FN, NEW_NN_FN=NEW_NN_FN)
    elif TRAINING:
        if data_fns is not None and not isinstance(data_fns, str) and len(data_fns) > 1:
            merge_data([os.path.basename(fn) for fn in data_fns], TRAINING_ITERATION)   # one file per iteration
            for fn in data_fns:
                os.remove(fn)
        training_data = _window_data(TRAINING_ITERATION, tk['SLIDING_WINDOW'])
        if TRAINING_ITERATION == 0 and NN_FN is None:
            nn = create_nn(**tk)
            NN_FN = save_nn_to_disk(nn, 0, create_timestamp())
        else:
            nn = load_model(NN_FN)
        history, NEW_NN_FN = train_nn(training_data, nn, **tk)
        out['plot'] = plot_history(history, nn, TRAINING_ITERATION)
        tk['OLD_NN_FN'], tk['NEW_NN_FN'] = NN_FN, NEW_NN_FN
        record_params('training', **tk)
        out.update(history=history, OLD_NN_FN=NN_FN, NEW_NN_FN=NEW_NN_FN)
        if world > 1:
            _D.broadcast_strings([NN_FN, NEW_NN_FN], rank, world)
    _barrier(world)
    if EVALUATION:
        tyk = dict(default_tourney_kwargs(TRAINING_ITERATION, NN_FN, NEW_NN_FN), **(tourney_kwargs or {}))
        tmk = dict(default_tourney_mcts_kwargs(NEW_NN_FN), **(tourney_mcts_kwargs or {}))
        print('Beginning tournament between {} and {}!'.format(NEW_NN_FN, NN_FN))
        out['tourney_fn'] = tournament_Checkers(tyk, tmk).start_tournament()
        if rank == 0:
            record_params('evaluation', **{**tyk, **tmk})
    _barrier(world)
    return out


def _barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def run_final_evaluation(fe_model_nums, tourney_kwargs=None, tourney_mcts_kwargs=None, num_cpus=4):
    """FINAL_EVALUATION (:47-48, 211-214): round-robin among the models of the listed iterations"""
    tyk = dict(default_tourney_kwargs(max(fe_model_nums), None, None), **(tourney_kwargs or {}))
    tmk = dict(default_tourney_mcts_kwargs(None), **(tourney_mcts_kwargs or {}))
    fe = final_evaluation(list(fe_model_nums), tyk, tmk)
    return fe.start_evaluation(num_cpus=num_cpus)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('--iteration', type=int, default=0, help='TRAINING_ITERATION')
    ap.add_argument('--nn-fn', default=None, help='NN_FN: current network (not needed for iteration 0)')
    ap.add_argument('--new-nn-fn', default=None, help='NEW_NN_FN: required for --evaluation without --training')
    ap.add_argument('--selfplay', action='store_true')
    ap.add_argument('--training', action='store_true')
    ap.add_argument('--evaluation', action='store_true')
    ap.add_argument('--final-evaluation', type=int, nargs='+', default=None, metavar='ITER',
                    help='iteration numbers of the models to compare')
    ap.add_argument('--games', type=int, default=None, help='NUM_SELFPLAY_GAMES')
    ap.add_argument('--budget', type=int, default=None, help='BUDGET for self-play and tournament searches')
    ap.add_argument('--epochs', type=int, default=None, help='EPOCHS')
    ap.add_argument('--tourney-games', type=int, default=None, help='TOURNEY_GAMES per worker')
    args = ap.parse_args(argv)
    sp = {} if args.games is None else {'NUM_SELFPLAY_GAMES': args.games}
    mk = {} if args.budget is None else {'BUDGET': args.budget}
    tk = {} if args.epochs is None else {'EPOCHS': args.epochs}
    tyk = {} if args.tourney_games is None else {'TOURNEY_GAMES': args.tourney_games}
    if args.selfplay or args.training or args.evaluation:
        out = run_iteration(args.iteration, args.nn_fn, args.new_nn_fn, args.selfplay, args.training, args.evaluation,
                            sp, mk, tk, tyk, dict(mk))
        for key in ('data_fns', 'OLD_NN_FN', 'NEW_NN_FN', 'plot', 'tourney_fn'):
            if key in out:
                print('{} = {}'.format(key, out[key]))
    if args.final_evaluation:
        run_final_evaluation(args.final_evaluation, tyk, dict(mk))


if __name__ == '__main__':
    main()
