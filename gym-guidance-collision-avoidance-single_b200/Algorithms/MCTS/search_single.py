"""Drop-in for the reference's Algorithms/MCTS/search_single.py:5-22."""


class MCTS:
    def __init__(self, node):
        self.root = node

    def best_action(self, simulations, search_depth):
        for _ in range(simulations):
            v = self.tree_policy(search_depth)
            reward = v.rollout(search_depth)
            v.backpropagate(reward)
        return self.root.best_child(c_param=0.)

    def tree_policy(self, search_depth):
        current = self.root
        while not current.is_terminal_node(search_depth):
            if not current.is_fully_expanded():
                return current.expand()
            current = current.best_child()
        return current
