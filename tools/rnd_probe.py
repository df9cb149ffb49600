"""Terminal fraction / per-root survivors of the random-intruder planner model on the bench's roots."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gym-guidance-collision-avoidance-single_b200")]
import torch
from gca_b200.batched import BatchedAircraftEnv
from gca_b200 import abi, mcts
from Simulators.config import Config as SimConfig
from Algorithms.MCTS.config_single import Config as MctsConfig
B, N = 65536, 80
env = BatchedAircraftEnv("SingleAircraftMCTSRandIntruderEnv", B, SimConfig, n_intruders=N, mode="fast", draws="philox", seed=8)
env.reset()
for i in range(155):
    env.step(torch.randint(0, 9, (B,), device="cuda", dtype=torch.int32))
roots = env.obs[:2048].double().contiguous()
mcfg = abi.make_mcts_config(MctsConfig, random_intruders=True)
r, f, fl = mcts.playouts(roots, 100, depth=3, cfg=mcfg, seed=1)
fl = fl.cpu().numpy()
print("terminal fraction", (fl != 0).mean(), "wall", (fl == abi.MCTS_WALL).mean(), "conflict", (fl == abi.MCTS_CONFLICT).mean(), "goal", (fl == abi.MCTS_GOAL).mean())
print("roots with every playout terminal", ((fl != 0).mean(1) == 1).mean(), "roots with none", ((fl != 0).mean(1) == 0).mean())
own = roots[:, 6 * N:6 * N + 2].cpu().numpy()
print("ownship outside the map at the root:", ((own[:, 0] < 0) | (own[:, 0] > 800) | (own[:, 1] < 0) | (own[:, 1] > 800)).mean())
