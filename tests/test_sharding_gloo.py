"""Multi-GPU plumbing on CPU: world_size-2 gloo processes compute their env shards, the rank-0
aggregation of bench.py's timing (MAX over ranks) and the per-rank Philox key ranges.

The data path has no collective (envs are independent); what needs checking on the
host is that shards tile the global env-id space without overlap and that the oracle - standing
in for a rank's device - gives results independent of how the batch is split."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, per_rank, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "gym-guidance-collision-avoidance-single_b200")]
    import torch
    import torch.distributed as dist
    from gca_b200 import variants
    from gym_guidance_collision_avoidance_single.envs.config import Config
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = variants.make_config("SingleAircraftEnv", Config)
    env = orc.OracleEnv(cfg, per_rank, 12, draws=1, trig=orc.TRIG_SHARED, seed=9, env_id0=rank * per_rank,
                        f32_positions=True, auto_reset=True)
    env.reset()
    rng = np.random.RandomState(3)
    acts = rng.randint(0, 9, (6, world * per_rank))
    for t in range(6):
        env.step(acts[t, rank * per_rank:(rank + 1) * per_rank])
    np.save(os.path.join(out_dir, "obs_%d.npy" % rank), env.obs)
    # bench.py's reduction: elapsed time = MAX over ranks, throughput = all ranks' units / that
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ids = torch.tensor([rank * per_rank, (rank + 1) * per_rank - 1])
    gathered = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, ids)
    # the one optional collective of the path: the statistics reduce of AircraftVecMonitor.stats()
    from gca_b200.vec_monitor import reduce_stats
    codes, dones = env.info, env.done
    mine = torch.tensor([per_rank, int(dones.sum()), int((codes == 1).sum()), int((codes == 2).sum()),
                         int((codes == 3).sum()), int((codes == 4).sum()), int((codes == 5).sum()), 0])
    total = reduce_stats(mine)
    if rank == 0:
        np.save(os.path.join(out_dir, "meta.npy"), np.array([t.item()] + [int(x) for g in gathered for x in g]))
        np.save(os.path.join(out_dir, "stats.npy"), np.array([total["steps"], total["episodes"], total["conflict_steps"]]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding(tmp_path):
    world, per_rank = 2, 40
    mp.spawn(_worker, args=(world, _free_port(), per_rank, str(tmp_path)), nprocs=world, join=True)
    meta = np.load(tmp_path / "meta.npy")
    assert meta[0] == 11.0                                   # MAX over ranks
    assert meta[1:].tolist() == [0, 39, 40, 79]              # contiguous, non-overlapping global env ids
    sys.path[:0] = [ROOT, os.path.join(ROOT, "gym-guidance-collision-avoidance-single_b200")]
    from gca_b200 import variants
    from gym_guidance_collision_avoidance_single.envs.config import Config
    from oracle import oracle as orc
    whole = orc.OracleEnv(variants.make_config("SingleAircraftEnv", Config), world * per_rank, 12, draws=1,
                          trig=orc.TRIG_SHARED, seed=9, f32_positions=True, auto_reset=True)
    whole.reset()
    acts = np.random.RandomState(3).randint(0, 9, (6, world * per_rank))
    for t in range(6):
        whole.step(acts[t])
    sharded = np.concatenate([np.load(tmp_path / ("obs_%d.npy" % r)) for r in range(world)])
    assert np.array_equal(sharded, whole.obs)                # results do not depend on the number of ranks
    stats = np.load(tmp_path / "stats.npy")                  # summed over the two ranks = the unsharded batch's last step
    assert stats.tolist() == [world * per_rank, int(whole.done.sum()), int((whole.info == 2).sum())]
