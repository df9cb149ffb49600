"""Pins the MCTS part of the oracle: every recorded move(), rollout() and whole best_action()
search of the reference (Algorithms/MCTS) is replayed from the recorded numpy draws."""
import os

import numpy as np
import pytest

from helpers import GOLDEN
from gca_b200 import abi
from oracle import oracle as orc


def mcts_cfg():
    from Algorithms.MCTS.config_single import Config
    return abi.make_mcts_config(Config)


@pytest.mark.parametrize("n", [3, 80])
def test_move_matches_reference(n):
    g = np.load(os.path.join(GOLDEN, "mcts_n%d.npz" % n))
    cfg = mcts_cfg()
    seen = set()
    for k in range(len(g["mv_root"])):
        root = g["roots"][g["mv_root"][k]]
        a = g["mv_action"][k]
        tape = np.nan_to_num(g["mv_tape"][k], nan=0.0)
        st, flags, reward, used = orc.mcts_move(cfg, n, root, int(a[0]) * 3 + int(a[1]), tape=tape)
        assert used == g["mv_tape_len"][k]
        assert np.array_equal(st, g["mv_out_state"][k]), k
        assert bool(flags & abi.MCTS_WALL) == bool(g["mv_hit_wall"][k])
        assert bool(flags & abi.MCTS_CONFLICT) == bool(g["mv_conflict"][k])
        assert bool(flags & abi.MCTS_GOAL) == bool(g["mv_reach_goal"][k])
        assert reward == g["mv_reward"][k]
        seen.add(flags)
    assert {0, abi.MCTS_WALL, abi.MCTS_CONFLICT, abi.MCTS_GOAL} <= seen      # every branch exercised


@pytest.mark.parametrize("n", [3, 80])
def test_rollout_matches_reference(n):
    g = np.load(os.path.join(GOLDEN, "mcts_n%d.npz" % n))
    cfg = mcts_cfg()
    for k in range(len(g["ro_root"])):
        root = g["roots"][g["ro_root"][k]]
        tape = np.nan_to_num(g["ro_tape"][k], nan=0.0)
        reward, first, flags, used = orc.mcts_rollout(cfg, n, root, int(g["ro_depth"][k]), tape=tape)
        assert used == g["ro_tape_len"][k], k
        assert reward == g["ro_reward"][k], k


@pytest.mark.parametrize("n", [3, 80])
def test_search_matches_reference(n):
    g = np.load(os.path.join(GOLDEN, "mcts_n%d.npz" % n))
    cfg = mcts_cfg()
    for k in range(len(g["bs_root"])):
        root = g["roots"][g["bs_root"][k]]
        tape = np.nan_to_num(g["bs_tape"][k], nan=0.0)
        best, cn, cq, ca, used = orc.mcts_search(cfg, n, root, int(g["bs_sims"][k]), int(g["bs_depth"][k]), tape)
        assert used == g["bs_tape_len"][k], k
        want = g["bs_action"][k]
        assert best == int(want[0]) * 3 + int(want[1]), k
        assert np.array_equal(cn, g["bs_child_n"][k]) and np.array_equal(cq, g["bs_child_q"][k])
        wa = g["bs_child_action"][k]
        assert np.array_equal(ca, np.where(wa[:, 0] >= 0, wa[:, 0] * 3 + wa[:, 1], -1))
