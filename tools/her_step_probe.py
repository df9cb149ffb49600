"""HER-variant step alone and with the relabel reward (bench.py her leg split up): python tools/her_step_probe.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.build()
import torch  # noqa: E402
import bench  # noqa: E402
from gca_b200 import abi  # noqa: E402
from gca_b200.batched import BatchedAircraftEnv, compute_reward  # noqa: E402
from gym_guidance_collision_avoidance_single.envs.config import Config  # noqa: E402

B, N, k = 65536, 80, 4
out = {}
for variant in ("SingleAircraftHEREnv", "SingleAircraft2Env", "SingleAircraftDiscreteHEREnv"):
    env = BatchedAircraftEnv(variant, B, Config, n_intruders=N, mode="fast", draws="philox", seed=4)
    env.reset()
    if env.continuous:
        acts = [torch.rand((B, 2), device="cuda") * 2 - 1 for _ in range(8)]
    else:
        acts = [torch.randint(0, 3, (B,), device="cuda", dtype=torch.int32) for _ in range(8)]
    out[variant] = {"step_only_ms": bench.graph_step_ms(lambda i: env.step(acts[i % 8]), 8)}
    if env.is_goal_env:
        goals = torch.rand((k, B, 2), device="cuda")
        kind = abi.OBS_HER if variant == "SingleAircraftHEREnv" else abi.OBS_DHER

        def one(i):
            env.step(acts[i % 8])
            ag = env.achieved.unsqueeze(0).expand(k, B, 2).reshape(k * B, 2)
            return compute_reward(ag, goals.reshape(k * B, 2), Config.goal_radius, kind)
        out[variant]["step_plus_relabel_ms"] = bench.graph_step_ms(one, 8)
    env.close()
print(json.dumps(out))
