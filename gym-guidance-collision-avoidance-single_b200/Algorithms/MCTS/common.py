"""Abstract state / node of the search tree, with the names Algorithms/MCTS/common.py of the reference exports
(`MCTSState`, `MCTSNode`) and the two pieces of behaviour its subclasses inherit: the UCT child selection
(common.py:47-52) and the uniform rollout policy (:54-55)."""
import math

import numpy as np


def _abstract(name):
    def method(self, *args, **kwargs):
        raise NotImplementedError("%s.%s" % (type(self).__name__, name))
    method.__name__ = name
    return method


class MCTSState(object):
    """Holds `.state`; subclasses provide reward / is_terminal_state / move / get_legal_actions."""

    def __init__(self, state):
        self.state = state

    reward = _abstract("reward")
    is_terminal_state = _abstract("is_terminal_state")
    move = _abstract("move")
    get_legal_actions = _abstract("get_legal_actions")


class MCTSNode(object):
    """Tree node: `.parent`, `.children`, value sum `.q`, visit count `.n`."""

    def __init__(self, parent=None):
        self.parent, self.children = parent, []
        self.q, self.n = 0., 0

    expand = _abstract("expand")
    is_terminal_node = _abstract("is_terminal_node")
    rollout = _abstract("rollout")
    backpropagate = _abstract("backpropagate")

    def is_fully_expanded(self):
        return not self.untried_actions

    def uct(self, child, c_param):
        """q / n + c * sqrt(2 ln N / n) of one child."""
        return (child.q / child.n) + c_param * math.sqrt(2 * math.log(self.n) / child.n)

    def best_child(self, c_param=1.4):
        """c = 1.4 inside the tree, 0 for the final pick (Q27); the first maximum wins, like np.argmax."""
        best, best_w = None, -math.inf
        for child in self.children:
            w = self.uct(child, c_param)
            if w > best_w:
                best, best_w = child, w
        return best

    def rollout_policy(self, possible_moves):
        return possible_moves[np.random.randint(len(possible_moves))]
