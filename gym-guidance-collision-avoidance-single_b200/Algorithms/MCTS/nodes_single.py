"""Drop-in for the reference's Algorithms/MCTS/nodes_single.py: same class names, constructor
arguments, attributes and methods; `move` and `rollout` run on the GPU (libgca, gca_mcts.cu).

    state = SingleAircraftState(state=last_observation)      # raw obs of Simulators/SingleAircraftMCTSEnv
    root = SingleAircraftNode(state=state)
"""
import itertools

import numpy as np

from gca_b200 import abi, mcts as _dev
from .common import MCTSNode, MCTSState
from .config_single import Config

_counter = [0]


def _next_id():
    _counter[0] = (_counter[0] + 1) & 0x7fffffff
    return _counter[0]


class SingleAircraftState(MCTSState):
    RANDOM_INTRUDERS = False     # True in nodes_single_randintru.py: 6 values per intruder, intruders that turn

    @classmethod
    def model_config(cls):
        return abi.make_mcts_config(Config, random_intruders=cls.RANDOM_INTRUDERS)

    def __init__(self, state, hit_wall=False, conflict=False, reach_goal=False, prev_action=None, depth=0):
        MCTSState.__init__(self, np.asarray(state, dtype=np.float64))
        self.hit_wall = hit_wall
        self.conflict = conflict
        self.reach_goal = reach_goal
        self.prev_action = prev_action
        self.depth = depth
        self.config = Config()
        self.G = self.config.G
        self.scale = self.config.scale
        self.nearest_x = -1
        self.nearest_y = -1

    # nodes_single.py:25-32
    def reward(self):
        if self.hit_wall or self.conflict:
            return 0
        if self.reach_goal:
            return 1
        return 1 - self.dist_goal() / 1200.0

    # nodes_single.py:34-37
    def is_terminal_state(self, search_depth):
        return bool(self.reach_goal or self.conflict or self.hit_wall or self.depth == search_depth)

    # nodes_single.py:39-100, on the device
    def move(self, action):
        import torch
        st = torch.as_tensor(self.state[None].copy(), device="cuda")
        code = torch.tensor([int(action[0]) * 3 + int(action[1])], dtype=torch.int32, device="cuda")
        seed = int(np.random.randint(2 ** 31))          # the reference draws from the global numpy stream
        flags = int(_dev.move(st, code, self.model_config(), seed=seed, id0=_next_id())[0].item())
        return type(self)(st[0].cpu().numpy(), bool(flags & abi.MCTS_WALL), bool(flags & abi.MCTS_CONFLICT),
                                   bool(flags & abi.MCTS_GOAL), tuple(action), self.depth + 1)

    def get_legal_actions(self):
        return list(itertools.product(range(3), repeat=2))

    def dist_goal(self):
        dx = self.ownx - self.goalx
        dy = self.owny - self.goaly
        return np.sqrt(dx ** 2 + dy ** 2)

    def dist_intruder(self):
        distance = 5000
        # (len - 9) // 4 intruders of 4 values (:112), (len - 8) // 6 of 6 values in nodes_single_randintru.py:126
        k, near = (6, (len(self.state) - 8) // 6) if self.RANDOM_INTRUDERS else (4, (len(self.state) - 9) // 4)
        for i in range(near):
            d = self.metric(self.state[k * i], self.state[k * i + 1], self.ownx, self.owny)
            if d < distance:
                distance = d
                self.nearest_x, self.nearest_y = self.state[k * i], self.state[k * i + 1]
        return distance

    def metric(self, x1, y1, x2, y2):
        return np.sqrt((x1 - x2) ** 2 + (y1 - y2) ** 2)

    ownx = property(lambda self: self.state[-8])
    owny = property(lambda self: self.state[-7])
    own_vx = property(lambda self: self.state[-6])
    own_vy = property(lambda self: self.state[-5])
    own_speed = property(lambda self: self.state[-4])
    own_heading = property(lambda self: self.state[-3])
    goalx = property(lambda self: self.state[-2])
    goaly = property(lambda self: self.state[-1])


class SingleAircraftNode(MCTSNode):
    def __init__(self, state, parent=None):
        MCTSNode.__init__(self, parent)
        self.state = state

    @property
    def untried_actions(self):
        if not hasattr(self, "_untried_actions"):
            self._untried_actions = self.state.get_legal_actions()
        return self._untried_actions

    def expand(self):
        action = self.untried_actions.pop()              # (2, 2) first (Q27)
        child = type(self)(self.state.move(action), parent=self)
        self.children.append(child)
        return child

    def is_terminal_node(self, search_depth):
        return self.state.is_terminal_state(search_depth)

    # nodes_single.py:198-204: one random playout from this node, as a single device launch
    def rollout(self, search_depth):
        import torch
        s = self.state
        if s.is_terminal_state(search_depth):
            return s.reward()
        root = torch.as_tensor(s.state[None].copy(), device="cuda")
        r, _, _ = _dev.playouts(root, 1, depth=search_depth - s.depth, cfg=s.model_config(),
                                seed=int(np.random.randint(2 ** 31)), root_id0=_next_id())
        return float(r[0, 0].item())

    def backpropagate(self, result):
        self.n += 1.
        self.q += result
        if self.parent:
            self.parent.backpropagate(result)
