"""Algorithms/MCTS/Agent.py:12-63 (run_experiment) batched on the device: episodes of Simulators/SingleAircraftMCTSEnv
driven by device-resident UCT searches (re-plan every 5 steps), printing the summary the reference prints (:55-62).
Usage on a GPU box: python tools/mcts_experiment.py [--envs 1024] [--episodes 4096] [-s 100] [-d 3]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.build()
from gca_b200 import mcts  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1024)
    ap.add_argument("--episodes", "-e", type=int, default=4096)
    ap.add_argument("--no_simulations", "-s", type=int, default=100)
    ap.add_argument("--search_depth", "-d", type=int, default=3)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--random-intruders", action="store_true",
                    help="Agent_RandInt.py: SingleAircraftMCTSRandIntruderEnv + the nodes_single_randintru.py model")
    args = ap.parse_args()
    out = mcts.run_experiment(args.envs, args.episodes, args.no_simulations, args.search_depth, seed=args.seed,
                              random_intruders=args.random_intruders)
    print("----------------------------------------")
    print("intruders: ", 80)
    print("search depth: ", args.search_depth)
    print("simulation: ", args.no_simulations)
    print("episodes: ", out["episodes"], " env steps: ", out["env_steps"], " searches: ", out["searches"])
    print("time per decision (ms, device): ", out["search_ms_total"] / max(out["searches"], 1))
    print("NMAC prob: ", out["nmac_prob"])
    print("goal prob: ", out["goal_prob"])
    print("average conflicts: ", out["average_conflicts"])
    print("average episode length: ", out["average_length"], " average return: ", out["average_return"])
    print("wall time (s): ", out["wall_s"], " decisions/s: ", out["searches_per_sec"])


if __name__ == "__main__":
    main()
