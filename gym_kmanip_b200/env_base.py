"""KManipEnv: the Gymnasium-facing env class of the drop-in boundary, plus ``make`` for the registered ids.

Mirrors the interface of reference gym_kmanip/env_base.py:16-267: same constructor keywords, the same Dict
observation / action spaces (float64 observations in [-1, 1], float32 actions in [-1, 1]), the same ``info`` keys,
``reset`` -> ``(obs, info)``, ``step`` -> ``(obs, reward, terminated, False, info)`` with ``is_success = reward > 2``.
The backend is selected through the same seam (``self.env = new(self)``), bound to the CUDA backend of
env_sim.py.  ``make(id)`` applies the 64-step TimeLimit that ``gym.make`` adds from the registration
(reference __init__.py:28,247).  ``log_h5py=True`` writes per-episode files in the reference's layout (log_episode.py);
camera observations (the Vision ids, ``render()``) come from the ray-casting kernels of csrc/km_render.cuh; the rerun logger
and the real-robot backend are out of scope and raise NotImplementedError.
"""
from __future__ import annotations

import os
import time
from collections import OrderedDict
from typing import Any, Dict, List, Optional

import numpy as np

from . import constants as K
from .spaces import Box, Dict as DictSpace, GymEnv


class KManipEnv(GymEnv):
    metadata = {"render_modes": ["rgb_array"], "render_fps": K.FPS}

    def __init__(self, seed: int = 0, render_mode: str = "rgb_array", obs_list: Optional[List[str]] = None,
                 act_list: Optional[List[str]] = None, sim: bool = True, mjcf_filename: str = K.SOLO_ARM_MJCF,
                 urdf_filename: str = K.SOLO_ARM_URDF, q_pos_home=None, q_dict=None, q_keys=None, q_id_r_mask=None,
                 q_id_l_mask=None, ctrl_id_r_grip=None, ctrl_id_l_grip=None, log_prefix: str = "test",
                 log_rerun: bool = False, log_h5py: bool = False, device: int = 0, dtype: str = "float64", ik_mode: int = 1):
        super().__init__()
        obs_list = list(obs_list) if obs_list is not None else ["q_pos", "q_vel", "cube_pos", "cube_orn"]
        act_list = list(act_list) if act_list is not None else ["eer_pos", "eer_orn", "grip_r"]
        if log_rerun:
            raise NotImplementedError("the rerun visualisation logger is out of scope (SURVEY.md 2.1 #8)")
        if not sim:
            raise NotImplementedError("the real-robot backend is out of scope (SURVEY.md 2.1 #6)")
        self.render_mode = render_mode
        self.seed = seed
        self.step_idx = 0
        self.episode_idx = 0
        self.q_pos_home = q_pos_home
        self.q_len = len(q_pos_home)
        self.q_dict = q_dict
        self.q_keys = q_keys
        assert len(q_keys) == self.q_len, "q parameters do not match"
        self.q_id_r_mask, self.q_id_l_mask = q_id_r_mask, q_id_l_mask
        self.ctrl_id_r_grip, self.ctrl_id_l_grip = ctrl_id_r_grip, ctrl_id_l_grip
        self.cameras: list = [K.CAMERAS[o.split("/")[-1]] for o in obs_list if "camera" in o]   # env_base.py:75-80
        self.log_rerun, self.log_h5py = False, bool(log_h5py)
        self._log = None
        if self.log_h5py:   # per-episode files in the reference's layout (env_base.py:82-95, log_h5py.py)
            import uuid
            from datetime import datetime
            name = "{}.{}.{}".format(log_prefix, str(uuid.uuid4())[:6], datetime.now().strftime(K.DATE_FORMAT))
            self.log_dir = os.path.join(K.DATA_DIR, name)
            os.makedirs(self.log_dir, exist_ok=True)
        self.mjcf_filename, self.urdf_filename = mjcf_filename, urdf_filename
        # observation space (env_base.py:116-147)
        self.obs_list = obs_list
        od = OrderedDict()
        for key, shape in (("q_pos", (self.q_len,)), ("q_vel", (self.q_len,)), ("cube_pos", (3,)), ("cube_orn", (4,))):
            if key in obs_list:
                od[key] = Box(low=-1, high=1, shape=shape, dtype=K.OBS_DTYPE)
        for cam in self.cameras:            # env_base.py:140-147
            od[cam.log_name] = Box(low=cam.low, high=cam.high, shape=(cam.h, cam.w, 3), dtype=cam.dtype)
        self.observation_space = DictSpace(od)
        # action space (env_base.py:149-190)
        self.act_list = act_list
        ad = OrderedDict()
        sizes = dict(eel_pos=3, eel_orn=3, eer_pos=3, eer_orn=3, grip_l=1, grip_r=1,
                     q_pos_r=0 if q_id_r_mask is None else len(q_id_r_mask),
                     q_pos_l=0 if q_id_l_mask is None else len(q_id_l_mask))
        for key in K.ACTION_KEY_ORDER:
            if key in act_list:
                ad[key] = Box(low=-1, high=1, shape=(sizes[key],), dtype=K.ACT_DTYPE)
        self.action_space = DictSpace(ad)
        self.action_len = len(self.action_space.spaces)
        # backend seam (env_base.py:192-200)
        self.sim = sim
        from .env_sim import new
        # the single-env class is about fidelity, not throughput: fp64 and the exact-parity IK (restated scipy TRF)
        self.env = new(self, device=device, dtype=dtype, ik_mode=ik_mode)
        self.info: Dict[str, Any] = {
            "step": self.step_idx, "episode": self.episode_idx, "is_success": False, "q_keys": self.q_keys,
            "q_len": self.q_len, "a_len": self.action_len, "obs_list": self.obs_list, "act_list": self.act_list,
            "cameras": self.cameras, "sim": self.sim,
        }

    def render(self):
        return self.env.k_render(K.CAMERAS["top"])

    def _stamp(self, sim_time, reward, terminated, success):
        self.info["step"] = self.step_idx
        self.info["episode"] = self.episode_idx
        self.info["sim_time"] = sim_time
        self.info["cpu_time"] = time.time()
        self.info["reward"] = reward
        self.info["is_success"] = success
        self.info["terminated"] = terminated

    def reset(self, seed=None, options=None):
        super().reset(seed=seed)
        cube_xyz = (options or {}).get("cube_xyz") if isinstance(options, dict) else None
        terminated, reward, _, observation, sim_time = self.env.k_reset(cube_xyz=cube_xyz)
        self.step_idx = 0
        self.episode_idx += 1
        self._stamp(sim_time, reward, terminated, False)
        if self.log_h5py:
            from . import log_episode
            log_episode.end(self._log)
            self._log = log_episode.new(self.log_dir, self.info)
        return observation, self.info

    def step(self, action):
        terminated, reward, _, observation, sim_time = self.env.k_step(action)
        self.step_idx += 1
        self._stamp(sim_time, reward, terminated, reward > K.REWARD_SUCCESS_THRESHOLD)
        if self._log is not None and self.step_idx <= K.MAX_EPISODE_STEPS:
            self._log.step(action, observation, self.info)
        return observation, reward, terminated, False, self.info

    def close(self):
        if self._log is not None:
            self._log.end()
            self._log = None
        self.env.k_close()
        super().close()


class TimeLimit:
    """The truncation ``gym.make`` adds for ``max_episode_steps`` (reference __init__.py:28,247)."""

    def __init__(self, env: KManipEnv, max_episode_steps: int = K.MAX_EPISODE_STEPS):
        self.env = env
        self._max, self._elapsed = max_episode_steps, 0

    @property
    def unwrapped(self):
        return self.env

    def __getattr__(self, name):
        return getattr(self.env, name)

    def reset(self, **kw):
        self._elapsed = 0
        return self.env.reset(**kw)

    def step(self, action):
        obs, rew, term, trunc, info = self.env.step(action)
        self._elapsed += 1
        return obs, rew, term, trunc or self._elapsed >= self._max, info

    def close(self):
        self.env.close()


def make(env_id: str, **kwargs):
    """``gym.make(id, **kwargs)`` for the ids registered at reference __init__.py:244-483."""
    if env_id not in K.ENV_REGISTRY:
        raise KeyError(f"unknown env id {env_id!r}; registered: {sorted(K.ENV_REGISTRY)}")
    kw = dict(K.ENV_REGISTRY[env_id])
    kw.update(kwargs)
    return TimeLimit(KManipEnv(**kw), K.MAX_EPISODE_STEPS)
