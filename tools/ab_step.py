"""A/B of two builds of libgca on the headline step: GCA_LIB=<so> python tools/ab_step.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gym-guidance-collision-avoidance-single_b200")]
import torch  # noqa: E402
import bench  # noqa: E402
from gca_b200.batched import BatchedAircraftEnv  # noqa: E402
from gym_guidance_collision_avoidance_single.envs.config import Config  # noqa: E402

B, N = 65536, 80
env = BatchedAircraftEnv("SingleAircraft2Env", B, Config, n_intruders=N, mode="fast", draws="philox", seed=0)
env.reset()
acts = [torch.rand((B, 2), device="cuda") * 2 - 1 for _ in range(50)]
for rep in range(3):
    ms = bench.graph_step_ms(lambda i: env.step(acts[i]), 50, reps=20)
    print(os.environ.get("GCA_LIB", "default"), "us per step: %.2f" % (ms * 1e3))
