"""One rasteriser launch on a small batch (for ncu): python tools/raster_probe.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.build()
import torch  # noqa: E402
from gca_b200.stack import ImageBatch  # noqa: E402
from gym_guidance_collision_avoidance_single.envs.config import Config  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
env = ImageBatch(B, Config, n_intruders=80, frame_stack=4, seed=3)
env.reset()
a = torch.randint(0, 9, (B,), device="cuda", dtype=torch.int32)
for _ in range(3):
    env.step(a)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    env._raster(None)
e.record()
torch.cuda.synchronize()
print("raster ms per launch (B=%d): %.3f" % (B, s.elapsed_time(e) / 5))
