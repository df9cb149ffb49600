"""Base classes of the reference's Algorithms/MCTS/common.py (same names and semantics)."""
import numpy as np


class MCTSState:
    def __init__(self, state):
        self.state = state

    def reward(self):
        raise NotImplementedError

    def is_terminal_state(self, search_depth):
        raise NotImplementedError

    def move(self, action):
        raise NotImplementedError

    def get_legal_actions(self):
        raise NotImplementedError


class MCTSNode:
    def __init__(self, parent=None):
        self.parent = parent
        self.children = []
        self.q = 0.
        self.n = 0

    def is_fully_expanded(self):
        return len(self.untried_actions) == 0

    def best_child(self, c_param=1.4):
        # UCT: q/n + c * sqrt(2 ln N / n); c = 1.4 in the tree, 0 for the final pick (common.py:47-52, Q27)
        weights = [(c.q / c.n) + c_param * np.sqrt((2 * np.log(self.n) / c.n)) for c in self.children]
        return self.children[int(np.argmax(weights))]

    def rollout_policy(self, possible_moves):
        return possible_moves[np.random.randint(len(possible_moves))]
