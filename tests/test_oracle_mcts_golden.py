"""Pins the MCTS part of the oracle: every recorded move(), rollout() and whole best_action()
search of the reference (Algorithms/MCTS) is replayed from the recorded numpy draws."""
import os

import numpy as np
import pytest

from helpers import GOLDEN
from gca_b200 import abi
from oracle import oracle as orc


def mcts_cfg(rnd=False):
    from Algorithms.MCTS.config_single import Config
    return abi.make_mcts_config(Config, random_intruders=rnd)


# (file stem, N): nodes_single.py on 4 N + 8 states, nodes_single_randintru.py on 6 N + 8 states
MODEL_CASES = [("mcts", 3), ("mcts", 80), ("mctsrnd_model", 3), ("mctsrnd_model", 20)]


@pytest.mark.parametrize("stem,n", MODEL_CASES)
def test_move_matches_reference(stem, n):
    g = np.load(os.path.join(GOLDEN, "%s_n%d.npz" % (stem, n)))
    cfg = mcts_cfg(stem != "mcts")
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


@pytest.mark.parametrize("stem,n", MODEL_CASES)
def test_rollout_matches_reference(stem, n):
    g = np.load(os.path.join(GOLDEN, "%s_n%d.npz" % (stem, n)))
    cfg = mcts_cfg(stem != "mcts")
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


def test_philox_search_invariants():
    """The Philox variant of the oracle's best_action() (the contract of gca_mcts_search): same tree code as the
    tape-pinned search; structural invariants and determinism."""
    from gca_b200 import abi
    from Algorithms.MCTS.config_single import Config
    cfg = abi.make_mcts_config(Config)
    rng = np.random.RandomState(4)
    n, R = 80, 6
    roots = np.zeros((R, 4 * n + 8))
    for r in range(R):
        ip = rng.uniform(0, 800, (n, 2)); sp = rng.uniform(5 / 3, 8 / 3, n); hd = rng.uniform(0, 2 * np.pi, n)
        roots[r, :4 * n] = np.stack([ip[:, 0], ip[:, 1], sp * np.cos(hd), sp * np.sin(hd)], -1).ravel()
        roots[r, 4 * n:] = [400, 400, 1.2, 1.2, 1.7, 0.78, rng.uniform(100, 700), rng.uniform(100, 700)]
    best, cn, cq, ca = orc.mcts_search_philox(cfg, n, roots, 100, 3, seed=5, root_id0=3)
    best2, cn2, cq2, ca2 = orc.mcts_search_philox(cfg, n, roots, 100, 3, seed=5, root_id0=3)
    assert np.array_equal(best, best2) and np.array_equal(cn, cn2) and np.array_equal(cq, cq2)
    assert np.array_equal(ca, np.tile(np.arange(8, -1, -1), (R, 1)))          # expand() pops actions from the end (Q27)
    assert np.all(cn.sum(1) == 100) and np.all(cn >= 1)                       # every simulation descends through one root child
    assert np.all((cq >= 0) & (cq <= cn))                                     # rewards lie in [0, 1]
    pick = np.array([ca[r, np.argmax(cq[r] / cn[r])] for r in range(R)])      # best_child(c = 0): first maximum of q / n
    assert np.array_equal(best, pick)
    other, _, _, _ = orc.mcts_search_philox(cfg, n, roots, 100, 3, seed=6, root_id0=3)
    cnb = orc.mcts_search_philox(cfg, n, roots, 100, 3, seed=6, root_id0=3)[1]
    assert not np.array_equal(cn, cnb)                                        # another key, another search
    # root ids are global: a batch split in two gives the same answers
    a = orc.mcts_search_philox(cfg, n, roots[:3], 100, 3, seed=5, root_id0=3)
    b = orc.mcts_search_philox(cfg, n, roots[3:], 100, 3, seed=5, root_id0=6)
    assert np.array_equal(np.concatenate([a[0], b[0]]), best) and np.array_equal(np.concatenate([a[2], b[2]]), cq)
