"""Drop-in for the reference's ``MCTS`` / ``MCTS_Node`` (reference MCTS.py:36-430) on top of the
device engine.  The whole tree lives in GPU memory (structure-of-arrays node pool, libckb200
``ck_tree_*`` entry points); ``MCTS_Node`` objects are thin views that fetch ``n / w / p`` and
their children lazily.  Each root built with ``MCTS_Node(state)`` starts its own device tree, so the
reference's one-tree-per-player game loops (play_Checkers.py:128-157) run unchanged.  Class-level configuration, method names, argument meaning and raised
errors follow the reference:

    MCTS(GAME_ENV=..., UCT_C=..., CONSTRAINT='rollout', BUDGET=..., MULTIPROC=False, NEURAL_NET=True,
         VERBOSE=..., TRAINING=..., DIRICHLET_ALPHA=..., DIRICHLET_EPSILON=..., TEMPERATURE_TAU=...,
         TEMPERATURE_DECAY=..., TEMP_DECAY_DELAY=...)
    root = MCTS_Node(state); MCTS.begin_tree_search(root); best = MCTS.best_child(root)
    root = MCTS.new_root_node(best)

Differences that are forced by the device design and stated here rather than hidden:
``game_env.neural_net`` must be a device network (``ckb200.net.KerasLikeNet``) or a stub
(``ckb200.net.StubNet``) -- an arbitrary host ``predict`` object cannot be called from a CUDA
kernel and there is no CPU fallback; with ``CONSTRAINT='time'`` the clock is read every
``MCTS.time_check_sims`` simulations instead of before each one.  ``NEURAL_NET=False``
(MCTS.py:78-89,113-115: plain UCT, one child per visit, one random playout per simulation) runs on
the device as well and needs no network object.
"""
from datetime import datetime

import numpy as np

from ckb200 import codec
from ckb200 import lib as _L
from Checkers import history_counters


class _Tree(object):
    """One search tree in GPU memory: a single-slot engine (``ck_tree_*`` entry points use slot 0)."""

    def __init__(self, engine, net, kind):
        self.engine, self.net, self.kind = engine, net, kind
        self.epoch = 0               # engine.tree_epoch() the node ids of this tree's views belong to
        self.root_idx = None
        self.closed = False

    def close(self):
        if not self.closed:
            self.closed = True
            self.engine.close()


class MCTS(object):
    # The reference keeps one Python tree per player (play_Checkers.py:128-157, training_pipeline.py:353-377)
    # and a tournament game two more; the least recently created device tree is dropped beyond this many.
    max_trees = 4
    time_check_sims = 32     # CONSTRAINT='time': simulations between two looks at the clock
    _trees = []
    _tree_serial = 0         # every tree draws from its own random stream
    engine_options = {}      # extra ckb200.lib.make_cfg arguments (pool_cap, compact_always, ...)

    @classmethod
    def __init__(cls, **kwargs):
        cls.game_env = kwargs['GAME_ENV']
        cls.uct_c = kwargs['UCT_C']
        cls.constraint = kwargs['CONSTRAINT']
        cls.budget = kwargs['BUDGET']
        cls.multiproc = kwargs['MULTIPROC']
        cls.neural_net = kwargs['NEURAL_NET']
        cls.verbose = kwargs['VERBOSE']
        cls.training = kwargs['TRAINING']
        cls.alpha = kwargs['DIRICHLET_ALPHA']
        cls.epsilon = kwargs['DIRICHLET_EPSILON']
        cls.tau = kwargs['TEMPERATURE_TAU']
        cls.tau_decay = kwargs['TEMPERATURE_DECAY']
        cls.tau_decay_delay = kwargs['TEMP_DECAY_DELAY']
        cls.seed = kwargs.get('SEED', 1)
        cls.engine_options = dict(kwargs.get('ENGINE_OPTIONS', {}))
        if cls.constraint not in ('rollout', 'time'):
            raise ValueError('Invalid MCTS computational constraint!')
        cls._close_trees()

    # ---- engine plumbing ---------------------------------------------------------------------
    @classmethod
    def _close_trees(cls):
        for tree in cls._trees:
            tree.close()
        cls._trees = []

    @classmethod
    def _evaluator_kind(cls, net):
        kind = getattr(net, 'ck_evaluator', None)
        if not cls.neural_net:       # NEURAL_NET=False: UCT with playouts (MCTS.py:78-89); a StubNet('rollout_hash') picks the deterministic twin
            return kind if kind in ('rollout', 'rollout_hash') else 'rollout'
        return kind

    @classmethod
    def _new_tree(cls):
        """a fresh device tree configured from the class-level search parameters"""
        net = getattr(cls.game_env, 'neural_net', None)
        kind = cls._evaluator_kind(net)
        device = getattr(cls.game_env, 'device', 0)
        if kind is None:
            raise TypeError('game_env.neural_net must be a ckb200 device network (ckb200.net.KerasLikeNet) or a '
                            'ckb200.net.StubNet; the search runs on the GPU and cannot call a host predict()')
        # the engine sizes its node pool from the per-search budget; a time-limited search reserves room for
        # 2048 simulations' worth of expansions between compactions and reports CK_ERR_POOL_OVERFLOW beyond the pool
        budget = int(cls.budget) if cls.constraint == 'rollout' else 2048
        cfg = _L.make_cfg(n_slots=1, budget=budget, device=device, uct_c=cls.uct_c, training=cls.training,
                          alpha=cls.alpha, epsilon=cls.epsilon, tau=cls.tau, tau_decay=cls.tau_decay,
                          tau_decay_delay=cls.tau_decay_delay, evaluator=kind, keep_records=False,
                          seed=(cls.seed + 0x9E3779B97F4A7C15 * cls._tree_serial) % (1 << 63), **cls.engine_options)
        cls._tree_serial += 1
        tree = _Tree(_L.Engine(cfg), net, kind)
        if kind == "net":
            tree.engine.set_net(0, net.net)
        while len(cls._trees) >= cls.max_trees:
            cls._trees.pop(0).close()
        cls._trees.append(tree)
        return tree

    # ---- reference API -----------------------------------------------------------------------
    @classmethod
    def get_legal_next_states(cls, history):
        return cls.game_env.get_legal_next_states(history)

    @classmethod
    def determine_outcome(cls, node):
        return cls.game_env.determine_outcome(node.history)

    @classmethod
    def current_player(cls, state):
        return cls.game_env.current_player(state)

    @classmethod
    def begin_tree_search(cls, root_node):
        """BUDGET new simulations from ``root_node`` on top of any inherited statistics
        (MCTS.py:210-224)."""
        start = datetime.now()
        if cls.multiproc and not cls.neural_net:
            # MCTS.py:83-87 hands pool.map's (outcome, player) tuples to backpropagation, and the first simulation ends in
            # this TypeError (observed with the unmodified reference); the batched playout evaluator is the parallel path here
            raise TypeError("unsupported operand type(s) for +=: 'int' and 'tuple' (MULTIPROC=True with NEURAL_NET=False "
                            "fails like this in the reference; playouts already run in parallel on the device, pass MULTIPROC=False)")
        net = getattr(cls.game_env, 'neural_net', None)
        tree = root_node._tree
        if tree is not None and not tree.closed and tree.kind == "net" and tree.net is not net \
                and cls._evaluator_kind(net) == "net":
            tree.engine.set_net(0, net.net)          # the caller swapped game_env.neural_net (tournaments do)
            tree.net = net
        if tree is None or tree.closed or root_node._idx is None or root_node._epoch != tree.epoch \
                or tree.kind != cls._evaluator_kind(net):
            # a node that is not part of a live device tree: start a fresh tree at its state
            tree = cls._new_tree()
            rev, ply = history_counters(root_node.history)
            pos = codec.encode_state(root_node.state, rev, ply)
            hist = cls.game_env.history
            parent_player = int(hist[-2][4, 0, 0]) if len(hist) >= 2 else -1    # MCTS.py:167-173
            tree.engine.tree_set_root(pos, parent_player)
            root_node._tree, root_node._idx, root_node._epoch = tree, 0, tree.epoch   # a fresh root is node 0
        elif root_node._idx != tree.root_idx:
            tree.engine.tree_reroot(root_node._idx)
            epoch = tree.engine.tree_epoch()
            if epoch != tree.epoch:
                # the engine compacted the tree around the new root: it is node 0 now and every
                # other view of this tree is stale
                tree.epoch = epoch
                root_node._idx, root_node._epoch = 0, epoch
        tree.root_idx = root_node._idx
        eng = tree.engine
        if cls.verbose:
            print('Starting search!')
        if cls.constraint == 'rollout':
            eng.tree_search(cls.budget)
            cls.rollout_count = cls.budget
        else:
            # CONSTRAINT='time' (MCTS.py:193-195): BUDGET seconds of wall clock, checked between batches of
            # simulations (the reference checks before every single one)
            cls.rollout_count = 0
            cap = eng.pool_cap()
            while (datetime.now() - start).total_seconds() < cls.budget:
                # the device tree lives in a fixed node pool: a search that could outgrow it stops early with the
                # simulations it has (the reference's Python tree has no such bound) instead of failing the move
                if eng.tree_node_count() + cls.time_check_sims * _L.MAX_CHILDREN > cap:
                    print('Node pool exhausted after {} rollouts: search stopped before its time budget'.format(cls.rollout_count))
                    break
                eng.tree_search(cls.time_check_sims)
                cls.rollout_count += cls.time_check_sims
        root_node._children = None
        root_node._number_of_visits, root_node._total_reward, _b = eng.tree_root()
        if cls.verbose:
            print('Stopped  search after {} rollouts and {} duration!'.format(
                cls.rollout_count, str(datetime.now() - start)[2:-4]))

    @classmethod
    def best_child(cls, node, criterion='robust'):
        """most visited child, or a sample ~ n^(1/tau) while training with tau > 0 (MCTS.py:226-248)"""
        if cls.neural_net:
            criterion = 'robust'
        if criterion == 'max':                    # highest total reward (NEURAL_NET=False only, MCTS.py:232-234)
            return node.children[int(np.argmax([child.w for child in node.children]))]
        if criterion != 'robust':
            raise ValueError('Invalid winner selection criterion!')
        children = node.children
        visits = [child.n for child in children]
        if not cls.training or cls.tau <= 0:
            return children[int(np.argmax(visits))]
        expon = [float(n) ** (1 / cls.tau) for n in visits]
        total = np.sum(expon)
        probs = [n / total for n in expon]
        if cls.game_env.move_count > cls.tau_decay_delay:
            cls.tau -= cls.tau_decay
            if np.isclose(cls.tau, 0):
                cls.tau = 0
        return children[int(np.random.choice(len(children), p=probs))]

    @classmethod
    def new_root_node(cls, old_root):
        """walk from the previously chosen node through the states played since (multi-hop aware)
        to the node of the current game state (MCTS.py:250-295)"""
        hist = cls.game_env.history
        new_state = cls.game_env.state
        counter, state_idx = 1, -3
        while len(hist) + state_idx >= 0 and \
                cls.game_env.current_player(hist[-2]) == cls.game_env.current_player(hist[state_idx]):
            counter += 1
            state_idx -= 1
        new_root = old_root
        for idx in range(-counter, 0, 1):
            for child in new_root.children:
                if (child.state[:5] == hist[idx][:5]).all():
                    new_root = child
                    break
        if (new_root.state[:5] == new_state[:5]).all():
            new_root.parent = None
            return new_root
        raise ValueError('All child nodes should be visited!  Consider '
                         'increasing number of rollouts or comment out this'
                         'error.')

    @classmethod
    def print_tree(cls, root_node, max_tree_depth=10):
        root_depth = root_node.depth

        def walk(node):
            w = node.w
            w_str = str(round(float(w), 1) if float(w) % 1 else int(w))
            print('\t' * (node.depth - root_depth) + '|- ({}/{}) ({:.1f}%)'.format(w_str, node.n, node.pwin))
            if node.depth - root_depth < max_tree_depth:
                for child in reversed(node.children):
                    walk(child)
        walk(root_node)


class MCTS_Node(object):
    """View of one node of the device tree (API of reference MCTS.py:345-430)."""

    def __init__(self, state, parent=None, initial_state=None):
        self.state = state
        self.player = MCTS.current_player(self.state)
        self.parent = parent
        if parent:
            self.history = parent.history.copy()
            self.history.append(state)
        else:
            self.history = [state]
            if initial_state is not None:
                self.history.insert(0, initial_state)
        self.depth = len(self.history)
        self._children = None
        self._number_of_visits = 0
        self._total_reward = 0
        self._prior_prob = 0
        self._tree = None                # device tree this view belongs to
        self._idx = None                 # node id inside that tree
        self._epoch = -1                 # tree epoch the id is valid for
        self._status = None
        self.printed = False

    # -- device-backed fields ------------------------------------------------------------------
    @property
    def children(self):
        if self._children is None:
            self._children = []
            tree = self._tree
            if tree is not None and not tree.closed and self._idx is not None and self._epoch == tree.epoch:
                for c in tree.engine.tree_children(self._idx):
                    node = MCTS_Node.__new__(MCTS_Node)
                    node.state = codec.decode_state(c["pos"])
                    node.player = MCTS.current_player(node.state)
                    node.parent = self
                    node.history = self.history + [node.state]
                    node.depth = self.depth + 1
                    node._children = None
                    node._number_of_visits, node._total_reward, node._prior_prob = c["n"], c["w"], c["p"]
                    node._tree, node._idx, node._epoch, node._status = tree, c["idx"], self._epoch, c["terminal"]
                    node.printed = False
                    self._children.append(node)
        return self._children

    @property
    def terminal(self):
        if self._status is None:
            self._status = 0 if MCTS.get_legal_next_states(self.history) else 1
        return self._status != 0

    @property
    def unvisited_child_states(self):
        """legal successors that are not in the tree yet: all of them before the first visit, none afterwards with
        a network (all children are added at once), the not yet visited ones with NEURAL_NET=False"""
        if self.terminal:
            return []
        legal = MCTS.get_legal_next_states(self.history)
        kids = self.children
        return [s for s in legal if not any((s[:5] == c.state[:5]).all() for c in kids)]

    @property
    def w(self):
        return self._total_reward

    @property
    def n(self):
        return self._number_of_visits

    @property
    def q(self):
        return self.w / self.n if self.n else 0

    @property
    def p(self):
        return self._prior_prob

    @property
    def pwin(self):
        return np.round((self.q + 1) / 2 * 100, 1)
