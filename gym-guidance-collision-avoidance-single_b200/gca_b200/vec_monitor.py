"""Episode statistics for the batched VecEnv: baselines' VecMonitor / Monitor without a host sync per step.

Reference: Algorithms/baselines-master/baselines/common/vec_env/vec_monitor.py (VecMonitor: eprets += rews, eplens += 1,
epinfo {'r', 'l', 't'} on done) and baselines/bench/monitor.py:98-124 (ResultsWriter: the `# {json}` header line and the
r,l,t csv rows that `load_results` reads).  The accumulators and a ring of finished-episode records live on the device
(csrc/gca_reward.cu: monitor_kernel behind gca_monitor_update); `drain()` copies the new records to the host, stamps
them with the wall-clock time of the step they finished in and writes the csv rows.
"""
import csv
import ctypes as C
import json
import os.path as osp
import time

import numpy as np

from . import abi

EXT = "monitor.csv"
_REC = np.dtype([("env", np.int32), ("length", np.int32), ("ep_return", np.float32), ("step", np.uint32)])


class ResultsWriter(object):
    """bench/monitor.py:98-124, same file format."""

    def __init__(self, filename=None, header="", extra_keys=()):
        self.extra_keys = extra_keys
        if filename is None:
            self.f = None
            self.logger = None
        else:
            if not filename.endswith(EXT):
                if osp.isdir(filename):
                    filename = osp.join(filename, EXT)
                else:
                    filename = filename + "." + EXT
            self.f = open(filename, "wt")
            if isinstance(header, dict):
                header = "# {} \n".format(json.dumps(header))
            self.f.write(header)
            self.logger = csv.DictWriter(self.f, fieldnames=("r", "l", "t") + tuple(extra_keys))
            self.logger.writeheader()
            self.f.flush()

    def write_row(self, epinfo):
        if self.logger:
            self.logger.writerow(epinfo)
            self.f.flush()

    def close(self):
        if self.f is not None:
            self.f.close()
            self.f = None


def reduce_stats(counters, reduce=True):
    """Sum a rank's counter tensor over the process group (if one is initialised) and name the entries."""
    import torch
    t = counters.clone()
    if reduce and torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)
    v = t.cpu().tolist()
    return dict(zip(abi.STAT_NAMES, v[: len(abi.STAT_NAMES)]))


class AircraftVecMonitor(object):
    """VecMonitor(venv, filename) for an AircraftVecEnv that keeps its outputs on the device."""

    def __init__(self, venv, filename=None, ring_capacity=1 << 20):
        import torch
        self.venv = venv
        self.num_envs = venv.num_envs
        self.observation_space, self.action_space = venv.observation_space, venv.action_space
        self.tstart = time.time()
        self.results_writer = ResultsWriter(filename, header={"t_start": self.tstart})
        dev = venv.batch.device
        self.eprets = torch.zeros(self.num_envs, dtype=torch.float32, device=dev)
        self.eplens = torch.zeros(self.num_envs, dtype=torch.int32, device=dev)
        self.cap = int(ring_capacity)
        self._ring = torch.zeros((self.cap, 4), dtype=torch.int32, device=dev)          # gca_episode_record x cap
        self._count = torch.zeros(1, dtype=torch.int64, device=dev)
        self._stats = torch.zeros(8, dtype=torch.int64, device=dev)                     # gca_stats_update counters
        self._drained = 0
        self._step = 0
        self._times = {}
        self.episode_rewards, self.episode_lengths, self.episode_times = [], [], []

    def reset(self):
        obs = self.venv.reset()
        self.eprets.zero_()                                   # vec_monitor.py:16-19
        self.eplens.zero_()
        return obs

    def step_async(self, actions):
        self.venv.step_async(actions)

    def step_wait(self):
        import torch
        obs, rews, dones, infos = self.venv.step_wait()
        b = self.venv.batch
        ret = (obs, rews, dones, infos)
        if getattr(self.venv, "host", False):
            # host=True hands out numpy views of pinned buffers: the accumulators live on the device, so the 6 bytes per
            # env of (reward, done, info) go back up (never a host pointer into a device kernel)
            rews = torch.as_tensor(np.ascontiguousarray(rews), device=b.device)
            dones = torch.as_tensor(np.ascontiguousarray(dones).astype(np.uint8), device=b.device)
            infos = torch.as_tensor(np.ascontiguousarray(infos, np.uint8), device=b.device)
        abi.check(b.lib.gca_monitor_update(rews.data_ptr(), 1 if rews.dtype == torch.float64 else 0, dones.data_ptr(),
                                           self.num_envs, self.eprets.data_ptr(), self.eplens.data_ptr(),
                                           self._ring.data_ptr(), self.cap, self._count.data_ptr(),
                                           self._step & 0xffffffff, b.device.index or 0,
                                           C.c_void_p(torch.cuda.current_stream(b.device).cuda_stream)))
        abi.check(b.lib.gca_stats_update(dones.data_ptr(), infos.data_ptr(), self.num_envs, self._stats.data_ptr(),
                                         b.device.index or 0, C.c_void_p(torch.cuda.current_stream(b.device).cuda_stream)))
        self._times[self._step & 0xffffffff] = round(time.time() - self.tstart, 6)
        self._step += 1
        return ret

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def drain(self):
        """Finished episodes since the last call as epinfo dicts {'r', 'l', 't', 'env'} (one host sync)."""
        total = int(self._count.item())
        new = total - self._drained
        if new <= 0:
            return []
        if new > self.cap:
            raise abi.GcaError("episode ring overflow: %d records since the last drain, capacity %d" % (new, self.cap))
        idx = (np.arange(self._drained, total) % self.cap).astype(np.int64)
        import torch
        raw = self._ring[torch.as_tensor(idx, device=self._ring.device)].cpu().numpy()
        recs = np.ascontiguousarray(raw).view(_REC).reshape(-1)
        self._drained = total
        order = np.lexsort((recs["env"], recs["step"]))        # step by step, envs in index order like the reference loop
        out = []
        for r in recs[order]:
            epinfo = {"r": float(r["ep_return"]), "l": int(r["length"]), "t": self._times.get(int(r["step"]), 0.0)}
            self.results_writer.write_row(epinfo)
            self.episode_rewards.append(epinfo["r"]); self.episode_lengths.append(epinfo["l"])
            self.episode_times.append(epinfo["t"])
            out.append(dict(epinfo, env=int(r["env"])))
        self._times = {}                                      # every step so far has been drained
        return out

    def stats(self, reduce=True):
        """Counters since construction {steps, episodes, nmac, conflict_steps, goal, wall, maxsteps}; with `reduce` and an
        initialised torch.distributed group they are summed over all ranks (the one collective of the path: a 64-byte
        all-reduce per call, NCCL over NVLink when the group's backend is nccl)."""
        return reduce_stats(self._stats, reduce)

    def close(self):
        self.drain()
        self.results_writer.close()
        self.venv.close()
