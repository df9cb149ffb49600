"""Kernel-level timeline of one step in steady state (debug build with -DGCA_PHASE_TIMING).
Usage on the GPU box: python tools/kstamps.py   (builds lib/libgca_timing.so first)"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gym-guidance-collision-avoidance-single_b200")
sys.path[:0] = [ROOT, PKG]
lib_path = os.path.join(PKG, "lib", "libgca_timing.so")
if not os.path.exists(lib_path) or os.environ.get("GCA_REBUILD"):
    csrc = os.path.join(PKG, "csrc")
    cu = sorted(os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cu"))
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
                           "-DGCA_PHASE_TIMING", "-Xcompiler", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-I" + csrc]
                          + cu + ["-o", lib_path])
os.environ["GCA_LIB"] = lib_path
import torch  # noqa: E402
from gca_b200 import abi  # noqa: E402
from gca_b200.batched import BatchedAircraftEnv  # noqa: E402
from gym_guidance_collision_avoidance_single.envs.config import Config  # noqa: E402

B, N = 65536, int(os.environ.get("GCA_N", "80"))
env = BatchedAircraftEnv("SingleAircraft2Env", B, Config, n_intruders=N, mode="fast", seed=1)
env.reset()
acts = [torch.rand((B, 2), device="cuda") * 2 - 1 for _ in range(8)]
lib = abi.load()
lib.gca_debug_kstamps.argtypes = [C.c_void_p, C.c_int]
buf = np.zeros(8, np.uint64)
g1 = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    for i in range(300):
        env.step(acts[i % 8])
torch.cuda.synchronize()
with torch.cuda.graph(g1):
    env.step(acts[0])
g8 = torch.cuda.CUDAGraph()
with torch.cuda.graph(g8):
    for i in range(8):
        env.step(acts[i])
names = ["own role", "main kernel (in: first block, out: last streaming block)", "finish kernel"]


def show(st, label):
    t0 = int(min(st[0], st[2]))
    rel = lambda v: (int(v) - t0) / 1e3
    print(label)
    print("  own role          first in %7.2f   last out %7.2f" % (rel(st[0]), rel(st[1])))
    print("  main kernel       first in %7.2f   last streaming block out %7.2f" % (rel(st[2]), rel(st[3])))
    print("  streaming role    first block in %7.2f   first record seen %7.2f" % (rel(st[6]), rel(st[7])))
    print("  finish kernel     first in %7.2f   last out %7.2f" % (rel(st[4]), rel(st[5])))


for rep in range(3):
    g8.replay()
torch.cuda.synchronize()
for rep in range(3):
    # steady state: 8 steps of a graph, then ONE more step as its own graph, stamped
    g8.replay()
    lib.gca_debug_kstamps(buf.ctypes.data, 1)          # (syncs through the symbol copy)
    g8.replay()
    torch.cuda.synchronize()
    lib.gca_debug_kstamps(buf.ctypes.data, 1)
    st = buf.astype(np.int64)
    print("8-step graph: first in .. last out = %.2f us per step" % ((st[5] - min(st[0], st[2])) / 8e3))
    g1.replay()
    torch.cuda.synchronize()
    lib.gca_debug_kstamps(buf.ctypes.data, 1)
    show(buf.astype(np.int64), "single-step graph (us, relative to the first block of the step):")
fb = np.zeros(2048 * 8, np.uint64)
lib.gca_debug_fin.argtypes = [C.c_void_p]
lib.gca_debug_fin(fb.ctypes.data)
f = fb.astype(np.int64).reshape(2048, 8)
dur = (f[:, 1] - f[:, 0]) / 1e3
print("finish_tile per tile (us): mean %.2f p50 %.2f p90 %.2f p99 %.2f max %.2f" % (dur.mean(), np.median(dur), np.percentile(dur, 90), np.percentile(dur, 99), dur.max()))
sp = (f[:, 7] - f[:, 1]) / 1e3
print("spawn phase (i) per tile (us): mean %.2f p50 %.2f p99 %.2f max %.2f" % (sp.mean(), np.median(sp), np.percentile(sp, 99), sp.max()))
print("finish start (rel to first) p10 %.1f p50 %.1f p90 %.1f max %.1f ; finish_tile end max %.1f ; spawn (i) end max %.1f" % tuple(
    [(np.percentile(f[:, 0], q) - f[:, 0].min()) / 1e3 for q in (10, 50, 90, 100)] + [(f[:, 1].max() - f[:, 0].min()) / 1e3, (f[:, 7].max() - f[:, 0].min()) / 1e3]))
print("finish phases (us, mean): loads %.2f  replay %.2f  reward+stores %.2f  reset+counters %.2f" % (
    ((f[:, 4] - f[:, 0]) / 1e3).mean(), ((f[:, 5] - f[:, 4]) / 1e3).mean(), ((f[:, 6] - f[:, 5]) / 1e3).mean(), ((f[:, 1] - f[:, 6]) / 1e3).mean()))
env.check()
