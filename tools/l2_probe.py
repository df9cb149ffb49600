"""Step time against batch size (position plane = 640 B x envs): shows what part of the step the L2 carries.
Usage on the GPU box: python tools/l2_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gym-guidance-collision-avoidance-single_b200")]
import torch  # noqa: E402
import bench  # noqa: E402
from gca_b200.batched import BatchedAircraftEnv  # noqa: E402
from gym_guidance_collision_avoidance_single.envs.config import Config  # noqa: E402

N = 80
for B in [int(x) for x in os.environ.get("GCA_PROBE_B", "8192,16384,32768,49152,65536,98304,131072,262144").split(",")]:
    env = BatchedAircraftEnv("SingleAircraft2Env", B, Config, n_intruders=N, mode="fast", seed=1)
    env.reset()
    acts = [torch.rand((B, 2), device="cuda") * 2 - 1 for _ in range(20)]
    ms = bench.graph_step_ms(lambda i: env.step(acts[i]), 20, reps=20)
    env.check()
    env.close()
    print("B %7d  plane %6.1f MB  step %7.2f us  %.3f ns per env-step  algorithmic %.2f TB/s" % (
        B, B * N * 8 / 1e6, ms * 1e3, ms * 1e6 / B, B * 3350 / (ms * 1e-3) / 1e12), flush=True)
