"""CPU restatement of baselines' HER "future" sampler - TEST INFRASTRUCTURE, NOT PRODUCT.

Follows Algorithms/baselines-master/baselines/her/her_sampler.py:19-61 (_sample_her_transitions) with the four
np.random calls replaced by explicit draw arrays (the tape recorded from the reference, or the Philox stream of
csrc/gca_her.cu), and the reward functions PKG/SingleAircraftHEREnv.py:194-196 / PKG/SingleAircraftDiscreteHEREnv.py:
184-186.  Pinned against tests/golden/her_sampler.npz (outputs of the unmodified reference function)."""
import ctypes as C

import numpy as np

OBS_HER, OBS_DHER = 1, 2


def compute_reward(ag, g, radius, kind):
    d = np.linalg.norm(ag - g, axis=-1)
    if kind == OBS_HER:
        return -(d > radius).astype(np.float32)          # PKG/SingleAircraftHEREnv.py:194-196
    return (d < radius).astype(np.float32)               # PKG/SingleAircraftDiscreteHEREnv.py:184-186


def sample_her_transitions(episode_batch, batch_size, replay_k, radius, kind, draws):
    """draws: dict(episode_idxs, t_samples int64 [batch]; u_her, u_offset float64 [batch])."""
    future_p = 1 - (1. / (1 + replay_k))                                                 # :14-17
    T = episode_batch["u"].shape[1]                                                      # :22
    episode_idxs = np.asarray(draws["episode_idxs"], np.int64)                           # :27
    t_samples = np.asarray(draws["t_samples"], np.int64)                                 # :28
    full = dict(episode_batch)
    full["o_2"] = episode_batch["o"][:, 1:, :]                                           # replay_buffer.py:46-47
    full["ag_2"] = episode_batch["ag"][:, 1:, :]
    transitions = {k: full[k][episode_idxs, t_samples].copy() for k in full}            # :29-30
    her_indexes = np.where(np.asarray(draws["u_her"]) < future_p)                        # :34
    future_offset = (np.asarray(draws["u_offset"]) * (T - t_samples)).astype(int)        # :35-36
    future_t = (t_samples + 1 + future_offset)[her_indexes]                              # :37
    transitions["g"][her_indexes] = episode_batch["ag"][episode_idxs[her_indexes], future_t]   # :42-43
    transitions["r"] = compute_reward(transitions["ag_2"], transitions["g"], radius, kind)     # :51-54
    ft = np.full(batch_size, -1, np.int32)
    ft[her_indexes] = future_t
    return transitions, ft


def philox_draws(batch, E, T, seed, call):
    """The Philox stream of csrc/gca_her.cu: counter (b lo, b hi, call, block), block 0 -> (episode, t), 1 -> (her, offset)."""
    from . import oracle as orc
    L = orc.lib()
    u = (C.c_double * 2)()
    e = np.zeros(batch, np.int64); t = np.zeros(batch, np.int64); uh = np.zeros(batch); uo = np.zeros(batch)
    for b in range(batch):
        L.gca_oracle_philox_uniform2(seed, b & 0xffffffff, b >> 32, call, 0, u)
        e[b] = min(int(u[0] * E), E - 1); t[b] = min(int(u[1] * T), T - 1)
        L.gca_oracle_philox_uniform2(seed, b & 0xffffffff, b >> 32, call, 1, u)
        uh[b], uo[b] = u[0], u[1]
    return {"episode_idxs": e, "t_samples": t, "u_her": uh, "u_offset": uo}
