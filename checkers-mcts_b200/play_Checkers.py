#!/usr/bin/env python3
"""Console demonstration game on the device engine -- the counterpart of the reference's
``play_Checkers.py`` (play_Checkers.py:42-168): human versus search or search versus search, one
tree per searching player that is re-rooted through the opponent's reply, optional tree print-outs
after every move.  The Pygame board of the reference (``Checkers_GUI``) is not part of this
package; everything else (option list for the human, board print-out, ``MCTS.print_tree``) is the
same text interface.

    python play_Checkers.py --model data/model/Checkers_Model10_12-Feb-2021(14:50:36).h5
    python play_Checkers.py --model stub:hash --budget 100 --max-plies 40 --quiet

``--model`` takes a Keras ``.h5`` file of the reference, a weight blob saved by this package, or
``stub:<kind>`` for a device stub evaluator (hash, uniform_zero, uniform_material).
"""
import argparse

import numpy as np

from Checkers import Checkers
from MCTS import MCTS, MCTS_Node

DEFAULT_MCTS_KWARGS = {       # competitive-play settings of the reference (play_Checkers.py:88-103)
    'UCT_C': 4, 'CONSTRAINT': 'rollout', 'BUDGET': 400, 'MULTIPROC': False, 'NEURAL_NET': True,
    'VERBOSE': False, 'TRAINING': False, 'DIRICHLET_ALPHA': 1.0, 'DIRICHLET_EPSILON': 0.25,
    'TEMPERATURE_TAU': 0, 'TEMPERATURE_DECAY': 0, 'TEMP_DECAY_DELAY': 0,
}


def states_to_piece_positions(state, next_states):
    """``[(row, col) from, (row, col) to]`` (1-based, as the reference prints them) for every legal
    next state: the mover's square that emptied and the one that filled (play_Checkers.py:62-84)."""
    mover = (0, 1) if state[4, 0, 0] == 0 else (2, 3)
    before = (state[mover[0]] + state[mover[1]]) > 0
    moves = []
    for nxt in next_states:
        after = (nxt[mover[0]] + nxt[mover[1]]) > 0
        src = np.argwhere(before & ~after)[0]
        dst = np.argwhere(after & ~before)[0]
        moves.append([(int(src[0]) + 1, int(src[1]) + 1), (int(dst[0]) + 1, int(dst[1]) + 1)])
    return moves


def get_human_input(game_env, input_fn=input, echo=print):
    """list the legal moves, read an option number until it is valid, play it; returns the chosen
    state (play_Checkers.py:42-60)"""
    while True:
        options = game_env.legal_next_states
        for idx, (src, dst) in enumerate(states_to_piece_positions(game_env.state, options)):
            echo('Option #{}: {} to {}'.format(idx + 1, src, dst))
        choice = int(input_fn('Enter option number: ')) - 1
        if 0 <= choice < len(options):
            game_env.step(options[choice])
            game_env.print_board()
            return options[choice]
        echo('Invalid selection!  Try again!')


def load_net(model, device=0):
    from ckb200.net import KerasLikeNet, StubNet
    from training_pipeline import load_blob
    spec = load_blob(model)
    return StubNet(spec) if isinstance(spec, str) else KerasLikeNet(spec, device)


def play(neural_net, mcts_kwargs=None, human_player1=False, human_player2=False, print_trees=True, tree_depth=1,
         max_plies=None, quiet=False, input_fn=input, on_search=None):
    """The reference's game loop (play_Checkers.py:108-157).  ``on_search(root, best_child)`` is called after
    every search (the GUI hook of the reference); returns ``(outcome, move_count)``."""
    kwargs = dict(DEFAULT_MCTS_KWARGS)
    kwargs.update(mcts_kwargs or {})
    game_env = Checkers(neural_net)
    kwargs['GAME_ENV'] = game_env
    MCTS(**kwargs)
    initial_state = game_env.state
    if not quiet:
        game_env.print_board()
    if human_player1 and human_player2:
        human_player2 = False
    roots = {'player1': None, 'player2': None}
    best = {'player1': None, 'player2': None}
    if not human_player1:
        roots['player1'] = MCTS_Node(initial_state, parent=None)
    while not game_env.done and (max_plies is None or game_env.move_count < max_plies):
        who = game_env.current_player(game_env.state)
        if (who == 'player1' and human_player1) or (who == 'player2' and human_player2):
            get_human_input(game_env, input_fn)
            continue
        if best[who] is not None:                       # re-root this player's tree through the moves played since
            roots[who] = MCTS.new_root_node(best[who])
        elif roots[who] is None:                        # the second player's first search (:145-147)
            roots[who] = MCTS_Node(game_env.state, parent=None, initial_state=initial_state)
        MCTS.begin_tree_search(roots[who])
        best[who] = MCTS.best_child(roots[who])
        if on_search is not None:
            on_search(roots[who], best[who])
        game_env.step(best[who].state)
        if not quiet:
            if print_trees:
                MCTS.print_tree(roots[who], tree_depth)
            game_env.print_board()
    return game_env.outcome, game_env.move_count


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('--model', default='data/model/Checkers_Model10_12-Feb-2021(14:50:36).h5')
    ap.add_argument('--budget', type=int, default=DEFAULT_MCTS_KWARGS['BUDGET'])
    ap.add_argument('--human', choices=('none', 'player1', 'player2'), default='none')
    ap.add_argument('--tree-depth', type=int, default=1)
    ap.add_argument('--no-trees', action='store_true')
    ap.add_argument('--max-plies', type=int, default=None)
    ap.add_argument('--quiet', action='store_true')
    ap.add_argument('--device', type=int, default=0)
    args = ap.parse_args(argv)
    outcome, plies = play(load_net(args.model, args.device), {'BUDGET': args.budget},
                          human_player1=args.human == 'player1', human_player2=args.human == 'player2',
                          print_trees=not args.no_trees, tree_depth=args.tree_depth, max_plies=args.max_plies,
                          quiet=args.quiet)
    print('{} after {} moves!'.format(outcome if outcome else 'unfinished', plies))


if __name__ == '__main__':
    main()
