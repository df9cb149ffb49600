"""Host side of the MCTS forward model: batched device moves / playouts and a batched planner.

Reference: Algorithms/MCTS/nodes_single.py (move, reward, rollout), search_single.py, common.py.
The drop-in node classes live in Algorithms/MCTS/ (same module names as the reference) and call
into this module; `plan_actions` is the batched entry point (B roots at once).
"""
import ctypes as C

import numpy as np

from . import abi


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise abi.GcaError("no CUDA device: the MCTS kernels have no CPU fallback")
    return torch


def default_config():
    from Algorithms.MCTS.config_single import Config
    return Config


def _stream(dev):
    return C.c_void_p(_torch().cuda.current_stream(dev).cuda_stream)


def move(states, actions, cfg=None, tape=None, cursor=None, seed=0, id0=0, first_frame=0):
    """SingleAircraftState.move for m states at once (nodes_single.py:39-100).

    states: CUDA float64 [m, 4N+8], advanced IN PLACE; actions: int32 [m] codes a0*3+a1.
    tape/cursor: optional recorded numpy draws [m, L] and int64 cursors [m] (replay parity tests).
    Returns uint8 flags [m] (abi.MCTS_WALL / MCTS_CONFLICT / MCTS_GOAL).
    """
    torch = _torch()
    lib = abi.load()
    cfg = cfg or abi.make_mcts_config(default_config())
    assert states.dtype == torch.float64 and states.is_contiguous() and states.is_cuda
    m, L = states.shape
    n = (L - 8) // (6 if cfg.random_intruders else 4)
    actions = actions.to(torch.int32).contiguous()
    flags = torch.zeros(m, dtype=torch.uint8, device=states.device)
    tp = None
    if tape is not None:
        tp = abi.GcaTape(tape.data_ptr(), tape.shape[1], cursor.data_ptr())
    abi.check(lib.gca_mcts_move(C.byref(cfg), n, states.data_ptr(), actions.data_ptr(), flags.data_ptr(), m,
                                C.byref(tp) if tp is not None else None, int(seed), int(id0), int(first_frame),
                                states.device.index or 0, _stream(states.device)))
    return flags


def playouts(roots, n_playouts, depth=None, cfg=None, first_action=None, seed=0, root_id0=0):
    """Random playouts (Node.rollout, nodes_single.py:198-204) from each root state.

    roots: CUDA float64 [R, 4N+8]; first_action: optional int8 [R, n_playouts] (-1 = random first move).
    Returns (rewards float64 [R, P], first int8 [R, P], flags uint8 [R, P]).
    """
    torch = _torch()
    lib = abi.load()
    cfg = cfg or abi.make_mcts_config(default_config())
    depth = cfg.search_depth if depth is None else int(depth)
    assert roots.dtype == torch.float64 and roots.is_contiguous() and roots.is_cuda
    R, L = roots.shape
    n = (L - 8) // (6 if cfg.random_intruders else 4)
    dev = roots.device
    rewards = torch.empty((R, n_playouts), dtype=torch.float64, device=dev)
    first = torch.empty((R, n_playouts), dtype=torch.int8, device=dev)
    flags = torch.empty((R, n_playouts), dtype=torch.uint8, device=dev)
    fa = None
    if first_action is not None:
        fa = first_action.to(torch.int8).contiguous()
    abi.check(lib.gca_mcts_playouts(C.byref(cfg), n, roots.data_ptr(), R, int(n_playouts), depth,
                                    fa.data_ptr() if fa is not None else None, int(seed), int(root_id0),
                                    rewards.data_ptr(), first.data_ptr(), flags.data_ptr(), dev.index or 0, _stream(dev)))
    return rewards, first, flags


_workspaces = {}


def search(roots, n_simulations=None, depth=None, cfg=None, seed=0, root_id0=0, return_children=False):
    """MCTS(root).best_action(no_simulations, search_depth) (search_single.py:8-22) for R roots at once, the UCT
    trees resident on the device (csrc/gca_mcts.cu: mcts_search_kernel, one lane per root).

    roots: CUDA float64 [R, 4N+8] raw observations of Simulators/SingleAircraftMCTSEnv.  Needs position_sigma == 0
    (the reference's config); otherwise use the node classes of Algorithms/MCTS.
    Returns int64 [R, 2] = best_node.state.prev_action per root (Algorithms/MCTS/Agent.py:41), and with
    return_children also (child_n f64 [R, 9], child_q f64 [R, 9], child_action int32 [R, 9])."""
    torch = _torch()
    lib = abi.load()
    cfg_cls = default_config()
    cfg = cfg or abi.make_mcts_config(cfg_cls)
    sims = int(cfg_cls.no_simulation if n_simulations is None else n_simulations)
    depth = cfg.search_depth if depth is None else int(depth)
    roots = roots.to(torch.float64).contiguous()
    assert roots.is_cuda
    R, L = roots.shape
    n = (L - 8) // (6 if cfg.random_intruders else 4)
    dev = roots.device
    need = int(lib.gca_mcts_search_workspace(C.byref(cfg), n, R, sims, depth))
    key = (dev.index or 0)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(max(need, 16), dtype=torch.uint8, device=dev)    # scratch, reused by later calls
        _workspaces[key] = ws
    best = torch.empty(R, dtype=torch.int32, device=dev)
    cn = torch.empty((R, 9), dtype=torch.float64, device=dev)
    cq = torch.empty((R, 9), dtype=torch.float64, device=dev)
    ca = torch.empty((R, 9), dtype=torch.int32, device=dev)
    abi.check(lib.gca_mcts_search(C.byref(cfg), n, roots.data_ptr(), R, sims, depth, int(seed), int(root_id0),
                                  ws.data_ptr(), ws.numel(), best.data_ptr(), cn.data_ptr(), cq.data_ptr(),
                                  ca.data_ptr(), dev.index or 0, _stream(dev)))
    b = best.to(torch.int64)
    act = torch.stack([b // 3, b % 3], -1)
    return (act, cn, cq, ca) if return_children else act


def run_experiment(num_envs, no_episodes, no_simulations=None, search_depth=None, seed=0, device=0, replan_every=5,
                   max_steps=None, sim_config=None, random_intruders=False):
    """Algorithms/MCTS/Agent.py:12-63 run_experiment, batched: `num_envs` SingleAircraftMCTSEnv instances advance
    together; every `replan_every` steps (Agent.py:34) each env's raw observation becomes the root of a device-resident
    UCT search whose best first action is repeated until the next re-plan.  Runs until `no_episodes` episodes have
    finished (in order of completion) and returns the statistics the reference prints (:55-62) plus throughput.

    random_intruders: Algorithms/MCTS/Agent_RandInt.py instead - Simulators/SingleAircraftMCTSRandIntruderEnv driven by
    the nodes_single_randintru.py model.  Every playout of that model moves its own intruders, so there is no
    device-resident tree: each decision spends its simulations as root-parallel playouts (`plan_actions`)."""
    import time
    torch = _torch()
    from .batched import BatchedAircraftEnv
    if sim_config is None:
        from Simulators.config import Config as sim_config
    cfg_cls = default_config()
    cfg = abi.make_mcts_config(cfg_cls, random_intruders=random_intruders)
    variant = "SingleAircraftMCTSRandIntruderEnv" if random_intruders else "SingleAircraftMCTSEnv"
    env = BatchedAircraftEnv(variant, num_envs, sim_config, mode="faithful", device=device, seed=seed)
    obs = env.reset()
    B = num_envs
    dev = obs.device
    t_in_ep = torch.zeros(B, dtype=torch.int64, device=dev)          # episode_time_step + 1 of each env
    action = torch.zeros((B, 2), dtype=torch.int64, device=dev)
    ret = torch.zeros(B, dtype=torch.float64, device=dev)
    results = {"n": 0, "g": 0, "other": 0}
    conflicts, returns, lengths = [], [], []
    episodes = steps = searches = 0
    search_ms = 0.0
    t0 = time.perf_counter()
    while episodes < no_episodes and (max_steps is None or steps < max_steps):
        replan = (t_in_ep % replan_every) == 0                        # Agent.py:34
        if bool(replan.any()):
            idx = replan.nonzero().squeeze(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if random_intruders:
                action[idx] = plan_actions(obs[idx], no_simulations, search_depth, cfg=cfg, seed=seed + 1 + steps)
            else:
                action[idx] = search(obs[idx], no_simulations, search_depth, cfg=cfg, seed=seed + 1 + steps)
            e1.record()
            torch.cuda.synchronize()
            search_ms += e0.elapsed_time(e1)
            searches += int(idx.numel())
        # no auto-reset: a finished env keeps its final state until its statistics are read (Agent.py:50-52)
        obs, rew, done, info = env.step((action[:, 0] * 3 + action[:, 1]).to(torch.int32), auto_reset=False)
        steps += 1
        ret += rew.to(torch.float64)
        t_in_ep += 1
        d = done.bool()
        if bool(d.any()):
            for code in info[d].tolist():
                name = abi.INFO_STR[code]
                results[name if name in ("n", "g") else "other"] += 1
            conflicts += env.counters()[d, 0].tolist()                  # env.no_conflict
            returns += ret[d].tolist()
            lengths += t_in_ep[d].tolist()
            episodes += int(d.sum())
            ret[d] = 0.0
            t_in_ep[d] = 0
            obs = env.reset(mask=done)                                  # last_observation = env.reset()
    wall = time.perf_counter() - t0
    env.close()
    n_done = max(episodes, 1)
    return {"episodes": episodes, "env_steps": steps * B, "searches": searches,
            "nmac_prob": results["n"] / n_done, "goal_prob": results["g"] / n_done,
            "average_conflicts": float(np.mean(conflicts)) if conflicts else 0.0,
            "average_return": float(np.mean(returns)) if returns else 0.0,
            "average_length": float(np.mean(lengths)) if lengths else 0.0,
            "search_ms_total": search_ms, "searches_per_sec": searches / (search_ms * 1e-3) if search_ms else 0.0,
            "wall_s": wall}


def plan_actions(obs, n_simulations=None, depth=None, cfg=None, seed=0, root_id0=0):
    """Batched planner: for each of the B raw observations pick the first action with the best mean
    playout reward, spending `n_simulations` playouts per root spread evenly over the 9 actions
    (root-parallel Monte-Carlo; the UCT tree of search_single.py is the drop-in single-root path).
    Returns int64 [B, 2] (a0, a1) like SingleAircraftState.prev_action."""
    torch = _torch()
    cfg_cls = default_config()
    cfg = cfg or abi.make_mcts_config(cfg_cls)
    sims = int(n_simulations or cfg_cls.no_simulation)
    per = max(1, (sims + 8) // 9)
    roots = obs.to(torch.float64).contiguous()
    B = roots.shape[0]
    fa = torch.arange(9, device=roots.device, dtype=torch.int8).repeat_interleave(per).repeat(B, 1)
    rewards, _, _ = playouts(roots, 9 * per, depth=depth, cfg=cfg, first_action=fa, seed=seed, root_id0=root_id0)
    mean = rewards.view(B, 9, per).mean(-1)
    best = mean.argmax(-1)
    return torch.stack([best // 3, best % 3], -1)
