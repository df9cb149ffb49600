"""Image-observation environments: SingleAircraftStackEnv (PKG/SingleAircraftStackEnv.py) and
the 4-frame stack of baselines' VecFrameStack (common/vec_env/vec_frame_stack.py:6-30).

The frame comes from the device rasteriser (csrc/gca_raster.cu); the stack is a ring of planes
[B, k, H/4, W/4] so that a step writes one 40 KB plane per env instead of rolling 160 KB.
"""
import ctypes as C

import numpy as np

from . import abi, sprites as _sprites
from .batched import BatchedAircraftEnv
from .single import _SingleBase
from .spaces import Box, Discrete


def _torch():
    import torch
    return torch


class ImageBatch(object):
    """B StackEnv instances: fused step kernel + rasteriser, frames kept in a k-plane ring on the GPU."""

    def __init__(self, num_envs, config=None, n_intruders=None, frame_stack=1, device=0, seed=0, env_id0=0,
                 mode="fast", sprites=None, sprite_dir=None):
        torch = _torch()
        if config is None:
            from gym_guidance_collision_avoidance_single.envs.config import Config as config
        self.batch = BatchedAircraftEnv("SingleAircraftStackEnv", num_envs, config, n_intruders=n_intruders, mode=mode,
                                        device=device, seed=seed, env_id0=env_id0)
        b = self.batch
        self.num_envs, self.k = b.num_envs, int(frame_stack)
        self.H, self.W = int(config.window_height) // 4, int(config.window_width) // 4
        sp = _sprites.resolve_sprites(sprites, sprite_dir)     # the reference's PNGs when a checkout is known
        self.sprites = torch.as_tensor(sp, device=b.device)
        self.ring = torch.zeros((self.num_envs, self.k, self.H, self.W), dtype=torch.uint8, device=b.device)
        self.slot = 0
        self.launches = 0

    def _raster(self, clear_mask):
        b = self.batch
        torch = _torch()
        plane = self.H * self.W
        abi.check(b.lib.gca_raster(b._h, self.sprites.data_ptr(), self.ring.data_ptr(), self.k * plane, plane, self.k,
                                   self.slot, clear_mask.data_ptr() if clear_mask is not None else None,
                                   C.c_void_p(torch.cuda.current_stream(b.device).cuda_stream)))
        self.launches += 1

    def reset(self):
        self.batch.reset()
        self.ring.zero_()                                  # VecFrameStack.reset: stackedobs[...] = 0
        self.slot = 0
        self._raster(None)
        return self.frame()

    def step(self, actions, auto_reset=True):
        obs, rew, done, info = self.batch.step(actions, auto_reset=auto_reset)
        self.slot = (self.slot + 1) % self.k
        self._raster(done if (auto_reset and self.k > 1) else None)
        return self.frame(), rew, done, info

    def frame(self):
        """Newest frame, uint8 [B, H/4, W/4, 1] (a view of the ring)."""
        return self.ring[:, self.slot].unsqueeze(-1)

    def plane_order(self):
        """Ring planes from oldest to newest."""
        return [(self.slot + 1 + j) % self.k for j in range(self.k)]

    def stacked(self):
        """Materialised VecFrameStack observation uint8 [B, H/4, W/4, k], newest in the last channel."""
        return self.ring[:, self.plane_order()].permute(0, 2, 3, 1).contiguous()

    def close(self):
        self.batch.close()


class SingleAircraftStackEnv(_SingleBase):
    """Image observation uint8 [200, 200, 1] (PKG/SingleAircraftStackEnv.py:13-214): max_steps rule,
    non-terminal wall penalty, goal reward 10000 (Q18)."""
    VARIANT = "SingleAircraftStackEnv"

    def __init__(self, device=0, seed=None, mode="faithful", time_limit=0, sprites=None, sprite_dir=None):
        self._sprites_arg = _sprites.resolve_sprites(sprites, sprite_dir)
        self._img = None
        _SingleBase.__init__(self, device=device, seed=seed, mode=mode, time_limit=time_limit)

    def load_config(self):
        _SingleBase.load_config(self)
        self.max_steps = self.Config.max_steps
        self.steps = 0

    def _build_spaces(self):
        torch = _torch()
        dim = (self.window_width // 4, self.window_height // 4, 1)
        self.observation_space = Box(low=0, high=255, shape=dim, dtype=np.uint8)
        self.action_space = Discrete(9)
        self._sp = torch.as_tensor(self._sprites_arg, device=self._batch.device)
        self._img = torch.zeros((1, 1, dim[1], dim[0]), dtype=torch.uint8, device=self._batch.device)

    def _format_obs(self):
        torch = _torch()
        b = self._batch
        plane = self._img.shape[2] * self._img.shape[3]
        abi.check(b.lib.gca_raster(b._h, self._sp.data_ptr(), self._img.data_ptr(), plane, plane, 1, 0, None,
                                   C.c_void_p(torch.cuda.current_stream(b.device).cuda_stream)))
        self.steps = int(b.get_state()["ep_steps"][0])
        return self._img[0, 0].cpu().numpy()[:, :, None]
