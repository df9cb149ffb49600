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
    n = (L - 8) // 4
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
    n = (L - 8) // 4
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
