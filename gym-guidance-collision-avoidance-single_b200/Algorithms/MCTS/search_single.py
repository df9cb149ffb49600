"""`MCTS(root).best_action(simulations, search_depth)` - the entry point Algorithms/MCTS/Agent.py:37-41 calls
(reference: Algorithms/MCTS/search_single.py:5-22).

Two ways to the same answer type (a child node of `root` whose `.state.prev_action` is the chosen (a0, a1)):
  * when the model has no position noise (the reference's configuration) the whole search runs on the device
    (gca_mcts_search: tree resident in HBM) and only the chosen child is materialised on the host;
  * otherwise the tree lives in Python objects (Algorithms/MCTS/common.py, nodes_single.py) whose move / rollout
    calls go to the device one at a time - the structure of the reference's loop.
"""


class MCTS(object):
    def __init__(self, node):
        self.root = node

    # ---- device-resident search -------------------------------------------------------------------------------
    def _device_search(self, simulations, search_depth):
        import numpy as np
        import torch
        from gca_b200 import abi, mcts as dev
        from .config_single import Config
        from .nodes_single import SingleAircraftNode
        if Config.position_sigma != 0 or simulations <= 0 or self.root.children:
            return None
        if getattr(self.root.state, "RANDOM_INTRUDERS", False):      # nodes_single_randintru.py: every playout moves its own intruders
            return None
        cfg = abi.make_mcts_config(Config)
        root_state = torch.as_tensor(np.asarray(self.root.state.state, np.float64)[None], device="cuda")
        action = dev.search(root_state, simulations, search_depth, cfg=cfg, seed=int(np.random.randint(2 ** 31)))
        chosen = tuple(int(x) for x in action[0].tolist())
        if chosen[0] < 0:
            return None
        child = SingleAircraftNode(self.root.state.move(chosen), parent=self.root)
        self.root.children.append(child)
        return child

    # ---- host tree ----------------------------------------------------------------------------------------------
    def tree_policy(self, search_depth):
        """Descend by UCT until a node that still has an untried action (expanded once) or is terminal."""
        node = self.root
        while True:
            if node.is_terminal_node(search_depth):
                return node
            if node.is_fully_expanded():
                node = node.best_child()
                continue
            return node.expand()

    def best_action(self, simulations, search_depth, device=True):
        if device:
            picked = self._device_search(simulations, search_depth)
            if picked is not None:
                return picked
        done = 0
        while done < simulations:
            leaf = self.tree_policy(search_depth)
            leaf.backpropagate(leaf.rollout(search_depth))
            done += 1
        return self.root.best_child(c_param=0.)
