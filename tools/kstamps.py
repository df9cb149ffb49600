"""Kernel-level timeline of one step in steady state (debug build with -DGCA_PHASE_TIMING).
Usage on the GPU box: GCA_LIB=.../libgca_timing.so python tools/kstamps.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gym-guidance-collision-avoidance-single_b200")]
import torch  # noqa: E402
from gca_b200 import abi  # noqa: E402
from gca_b200.batched import BatchedAircraftEnv  # noqa: E402
from gym_guidance_collision_avoidance_single.envs.config import Config  # noqa: E402

B, N = 65536, 80
env = BatchedAircraftEnv("SingleAircraft2Env", B, Config, n_intruders=N, mode="fast", seed=1)
env.reset()
acts = [torch.rand((B, 2), device="cuda") * 2 - 1 for _ in range(8)]
lib = abi.load()
lib.gca_debug_kstamps.argtypes = [C.c_void_p, C.c_int]
buf = np.zeros(8, np.uint64)
g = torch.cuda.CUDAGraph()
for i in range(300):
    env.step(acts[i % 8])
torch.cuda.synchronize()
with torch.cuda.graph(g):
    for i in range(8):
        env.step(acts[i])
for rep in range(3):
    g.replay()
torch.cuda.synchronize()
lib.gca_debug_kstamps(buf.ctypes.data, 1)
g.replay()
torch.cuda.synchronize()
lib.gca_debug_kstamps(buf.ctypes.data, 1)
st = buf.astype(np.int64).reshape(4, 2)
t0 = st[0, 0]
print("8 graph steps: own first-start .. last-end etc (us, relative):")
for k, name in enumerate(["own", "intruders", "finish"]):
    if st[k, 1] > 0:
        print("  %-10s first start %8.2f   last end %8.2f" % (name, (st[k, 0] - t0) / 1e3, (st[k, 1] - t0) / 1e3))
# single steps
for rep in range(3):
    lib.gca_debug_kstamps(buf.ctypes.data, 1)
    env.step(acts[rep])
    torch.cuda.synchronize()
    lib.gca_debug_kstamps(buf.ctypes.data, 1)
    st = buf.astype(np.int64).reshape(4, 2)
    t0 = st[0, 0]
    print("single step:", "  ".join("%s %.2f..%.2f" % (n, (st[k, 0] - t0) / 1e3, (st[k, 1] - t0) / 1e3)
                                    for k, n in enumerate(["own", "intr", "fin", "rspawn"]) if st[k, 1] > 0))
fb = np.zeros(2048 * 8, np.uint64)
lib.gca_debug_fin.argtypes = [C.c_void_p]
lib.gca_debug_fin(fb.ctypes.data)
f = fb.astype(np.int64).reshape(2048, 8)
dur = (f[:, 1] - f[:, 0]) / 1e3
print("finish per tile (us): mean %.2f p50 %.2f p90 %.2f p99 %.2f max %.2f" % (dur.mean(), np.median(dur), np.percentile(dur, 90), np.percentile(dur, 99), dur.max()))
for r in range(0, 4):
    m = f[:, 3] == r
    if m.any():
        print("  resets=%d: %4d tiles, mean %.2f us" % (r, m.sum(), dur[m].mean()))
for r in range(0, 8):
    m = (f[:, 2] == r) & (f[:, 3] == 0)
    if m.any():
        print("  no reset, max respawns per lane=%d: %4d tiles, mean %.2f us" % (r, m.sum(), dur[m].mean()))
print("finish start (rel to first) p10 %.1f p50 %.1f p90 %.1f max %.1f ; end max %.1f" % tuple(
    [(np.percentile(f[:, 0], q) - f[:, 0].min()) / 1e3 for q in (10, 50, 90, 100)] + [(f[:, 1].max() - f[:, 0].min()) / 1e3]))
print("finish phases (us, mean): loads %.2f  replay+respawn %.2f  reward+obs %.2f  reset+counters %.2f" % (
    ((f[:, 4] - f[:, 0]) / 1e3).mean(), ((f[:, 5] - f[:, 4]) / 1e3).mean(), ((f[:, 6] - f[:, 5]) / 1e3).mean(), ((f[:, 1] - f[:, 6]) / 1e3).mean()))
