"""Drop-in for the reference's ``Checkers`` game environment (reference Checkers.py:28-452) on top
of libckb200: the rules run on the GPU (K1 bitboard move generation, ``ck_movegen``), this class
only converts between the reference's 15x8x8 float64 state tensors and the compact positions.

Same attributes and methods, same argument meaning, same errors as the reference:
``state, history, legal_next_states, move_count, done, outcome, neural_net``;
``step, get_legal_next_states, determine_outcome, current_player, reset, init_board, predict,
set_prior_probs, print_board``.  There is no CPU rules fallback: without a CUDA device the calls
raise ``ckb200.lib.CkError``.
"""
import numpy as np

from ckb200 import codec
from ckb200 import lib as _L


def _reversible(prev, cur):
    """a ply is reversible iff it changed neither the men planes nor the piece count
    (the reference's draw-window scan, Checkers.py:335-343)"""
    return bool((prev[0] == cur[0]).all() and (prev[2] == cur[2]).all()
                and prev[0:4].sum() == cur[0:4].sum())


def history_counters(history):
    """-> (rev, ply) of history[-1]: consecutive trailing reversible plies, ply index."""
    rev = 0
    for i in range(len(history) - 1, 0, -1):
        if not _reversible(history[i - 1], history[i]):
            break
        rev += 1
    return rev, len(history) - 1


class Checkers(object):
    """Class to represent a game of Checkers (API of reference Checkers.py:28-452)."""

    def __init__(self, neural_net=None, device=0):
        self.device = device
        self.player1_man = 'x'
        self.player1_king = u'Ж'
        self.player2_man = 'o'
        self.player2_king = u'Ǒ'
        self.neural_net = neural_net
        self.reset()

    # ---- state management ---------------------------------------------------------------
    def init_board(self):
        """pieces on their starting squares (Checkers.py:415-423)"""
        self.state[:] = codec.decode_state(codec.start_position())

    def reset(self):
        self.state = np.zeros((15, 8, 8), dtype=float)
        self.init_board()
        self.history = [self.state]
        self.legal_next_states = self.get_legal_next_states(self.history)
        self.move_count = 0
        self.done = False
        self.outcome = None

    def current_player(self, state):
        return 'player1' if int(state[4, 0, 0]) == 0 else 'player2'

    # ---- rules (GPU) ---------------------------------------------------------------------
    def _movegen(self, history):
        state = history[-1]
        rev, ply = history_counters(history)
        pos = codec.encode_state(state, rev, ply)
        out = _L.movegen(np.array([pos], dtype=_L.POS_DTYPE), device=self.device)
        n = int(out["counts"][0])
        kids = [tuple(int(v) for v in out["children"][0, j]) for j in range(n)]
        return kids, [int(v) for v in out["masks"][0]], int(out["status"][0]), int(out["plane5"][0])

    def _check_moves(self, history):
        """raw legal successor list; also writes the legal-action planes 6-13 of history[-1] in
        place, as the reference does (Checkers.py:94-200)"""
        kids, mask, _status, _p5 = self._movegen(history)
        state = history[-1]
        for i in range(8):
            state[6 + i] = 0
            codec.bits_to_plane(mask[i], state[6 + i])
        return [codec.decode_state(k) for k in kids]

    def determine_outcome(self, history, legal_moves=[]):
        """(done, outcome) with outcome in 'player1_wins' / 'player2_wins' / 'draw' / None; writes the
        draw-counter plane 5 of history[-1] (Checkers.py:306-364)"""
        _kids, _mask, status, p5 = self._movegen(history)
        history[-1][5] = p5 / 80
        return status != codec.ONGOING, codec.OUTCOME_NAMES[status]

    def get_legal_next_states(self, history):
        legal = self._check_moves(history)
        done, _outcome = self.determine_outcome(history, legal_moves=legal)
        return [] if done else legal

    def step(self, next_state):
        if any((next_state[:5] == x[:5]).all() for x in self.legal_next_states):
            self.state = next_state
            self.history.append(self.state)
            self.legal_next_states = self._check_moves(self.history)
            self.done, self.outcome = self.determine_outcome(self.history, legal_moves=self.legal_next_states)
            self.move_count += 1
            return self.state, self.outcome, self.done
        raise ValueError('Illegal next state (invalid move)!')

    # ---- network glue ----------------------------------------------------------------------
    def predict(self, state):
        """masked + renormalised prior planes [8,8,8] and the value (Checkers.py:425-438); the mask /
        renormalisation runs in the ck_mask_renorm kernel (numpy's float32 summation order)."""
        nn_inp = np.moveaxis(state[:14], 0, -1).reshape(1, 8, 8, 14)
        prob_planes, q_value = self.neural_net.predict(nn_inp)
        mask = [codec.plane_to_bits(state[6 + i]) for i in range(8)]
        prior = _L.mask_renorm(np.asarray(prob_planes, dtype=np.float32).reshape(1, 512),
                               np.array([mask], dtype=np.uint32), device=self.device)
        return prior.reshape(8, 8, 8), np.asarray(q_value).reshape(-1)[0]

    def set_prior_probs(self, child_nodes, prob_planes):
        for child in child_nodes:
            layer = int(child.state[14, 0, 0]) - 6
            x = int(child.state[14, 0, 1])
            y = int(child.state[14, 0, 2])
            if x % 2 == y % 2:
                raise ValueError('Invalid (x,y) locations for probabilities!')
            if not (0 <= layer <= 7):
                raise ValueError('Invalid layer for probabilities!')
            child._prior_prob = prob_planes[layer, x, y]

    # ---- console view ----------------------------------------------------------------------
    def print_board(self):
        sym = {0: self.player1_man, 1: self.player1_king, 2: self.player2_man, 3: self.player2_king}
        rows = []
        for x in range(8):
            row = []
            for y in range(8):
                c = ' '
                for plane, s in sym.items():
                    if self.state[plane, x, y]:
                        c = s
                row.append(c)
            rows.append(row)
        try:
            from tabulate import tabulate
            print(tabulate(rows, tablefmt="fancy_grid"))
        except ImportError:
            print("\n".join(" ".join(r) for r in rows))
