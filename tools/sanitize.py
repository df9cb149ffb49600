"""Small-batch exercise of every kernel for compute-sanitizer (memcheck / racecheck / initcheck).
Usage on the GPU box: compute-sanitizer --tool memcheck python tools/sanitize.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gym-guidance-collision-avoidance-single_b200")]
import torch  # noqa: E402
from gca_b200.batched import BatchedAircraftEnv  # noqa: E402
from gca_b200.stack import ImageBatch  # noqa: E402
from gym_guidance_collision_avoidance_single.envs.config import Config  # noqa: E402

for variant, B, N, mode in [("SingleAircraft2Env", 1000, 80, "fast"), ("SingleAircraftEnv", 77, 33, "fast"),
                            ("SingleAircraftHEREnv", 300, 80, "fast"), ("SingleAircraftEnv", 130, 7, "faithful"),
                            ("SingleAircraftEnv", 64, 0, "fast"), ("SingleAircraftDiscreteHEREnv", 90, 200, "fast")]:
    env = BatchedAircraftEnv(variant, B, Config, n_intruders=N, mode=mode, seed=3)
    env.reset()
    for t in range(60):
        if env.continuous:
            a = torch.rand((B, 2), device="cuda", dtype=env.real) * 2 - 1
        else:
            a = torch.randint(0, 3, (B,), device="cuda", dtype=torch.int32)
        o, r, d, i = env.step(a, auto_reset=(t % 2 == 0))
        if t % 7 == 0 and bool(d.any()):
            env.reset(mask=d)
    env.observe()
    st = env.get_state()
    env.set_state(st)
    torch.cuda.synchronize()
    env.close()
    print("ok", variant, B, N, mode, flush=True)
img = ImageBatch(16, Config, n_intruders=20, frame_stack=4, seed=1)
img.reset()
for t in range(5):
    img.step(torch.randint(0, 9, (16,), device="cuda", dtype=torch.int32))
torch.cuda.synchronize()
print("ok stack")
