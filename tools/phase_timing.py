"""Per-warp phase timing of the step kernel (debug build with -DGCA_PHASE_TIMING).
Usage on the GPU box: GCA_LIB=.../libgca_timing.so python tools/phase_timing.py > gpurun_out/phases.txt"""
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
for i in range(30):
    env.step(acts[i % 8])
torch.cuda.synchronize()
lib = abi.load()
tiles = 32
nt = B // tiles
buf = np.zeros(nt * 8, np.uint64)
lib.gca_debug_phase_stamps.argtypes = [C.c_void_p, C.c_int]
assert lib.gca_debug_phase_stamps(buf.ctypes.data, nt * 8) == 0
st = buf.reshape(nt, 8).astype(np.int64)
t0 = st[:, 0].min()
rel = (st[:, :5] - t0) / 1e3          # us
sm = st[:, 7]
print("tiles", nt, "kernel span us", rel[:, 4].max())
names = ["start", "A done", "B done", "C done", "end"]
for j, n in enumerate(names):
    v = rel[:, j]
    print("%-8s min %7.2f  p10 %7.2f  median %7.2f  p90 %7.2f  max %7.2f" % (n, v.min(), np.percentile(v, 10), np.median(v),
                                                                          np.percentile(v, 90), v.max()))
d = np.diff(rel, axis=1)
for j, n in enumerate(["A", "B", "C", "D"]):
    v = d[:, j]
    print("phase %s  mean %7.2f  p10 %7.2f  median %7.2f  p90 %7.2f  max %7.2f" % (n, v.mean(), np.percentile(v, 10), np.median(v),
                                                                               np.percentile(v, 90), v.max()))
# per-SM finish time
fin = {}
for s_, e in zip(sm, rel[:, 4]):
    fin[s_] = max(fin.get(s_, 0), e)
f = np.array(sorted(fin.values()))
print("per-SM finish: min %.2f p10 %.2f median %.2f p90 %.2f max %.2f (n_sm=%d)" % (f.min(), np.percentile(f, 10), np.median(f),
                                                                                    np.percentile(f, 90), f.max(), len(f)))
cnt = np.bincount(sm.astype(int))
print("tiles per SM: min %d max %d" % (cnt[cnt > 0].min(), cnt.max()))
# finish time by SM id (looks for die / GPC structure in the imbalance)
order = sorted(fin.items())
print("finish by smid:", " ".join("%d:%.0f" % (k, v_) for k, v_ in order))
# B duration vs tile index (address structure)
bd = d[:, 1]
print("phase B by tile octile:", " ".join("%.1f" % bd[i * nt // 8:(i + 1) * nt // 8].mean() for i in range(8)))
# per-stage decomposition (every 32nd tile): wait for the TMA load | compute + stage | write-out
if hasattr(lib, "gca_debug_stage_stamps"):
    sb = np.zeros(64 * 16 * 4, np.uint64)
    lib.gca_debug_stage_stamps.argtypes = [C.c_void_p, C.c_int]
    assert lib.gca_debug_stage_stamps(sb.ctypes.data, sb.size) == 0
    ss = sb.reshape(64, 16, 4).astype(np.int64)
    nst = int(os.environ.get("GCA_NSTAGES", "10"))
    ss = ss[:, :nst]
    wait = (ss[:, :, 1] - ss[:, :, 0]) / 1e3
    comp = (ss[:, :, 2] - ss[:, :, 1]) / 1e3
    wout = (ss[:, :, 3] - ss[:, :, 2]) / 1e3
    print("per stage (us)   wait %.2f   compute %.2f   write-out %.2f   total %.2f" % (wait.mean(), comp.mean(), wout.mean(),
                                                                                    ((ss[:, -1, 3] - ss[:, 0, 0]) / 1e3 / nst).mean()))
    print("wait by stage:", " ".join("%.2f" % x for x in wait.mean(0)))
    print("compute by stage:", " ".join("%.2f" % x for x in comp.mean(0)))
    print("write-out by stage:", " ".join("%.2f" % x for x in wout.mean(0)))
