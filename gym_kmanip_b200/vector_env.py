"""KManipVectorEnv: the batched env -- N independent envs advanced by one fused CUDA launch per step.

The single-env semantics are those of KManipEnv (reference env_base.py:219-259) under the 64-step TimeLimit of
``gym.make`` (reference __init__.py:28,247), applied per env with same-step autoreset: on the step an env reaches
``max_episode_steps`` its ``truncated`` flag is 1, ``info["final_obs"]`` keeps the last observation of the
finished episode and ``obs`` already holds the first observation of the next one (cube re-spawned from the
counter-based generator keyed by (seed, global env id, episode)).  All arrays are torch CUDA tensors; observation
and action dicts use the reference's keys.  Envs shard across GPUs / ranks by contiguous global-id ranges with no
per-step communication (sharding.py).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional

from . import constants as K
from .batch_sim import BatchSim
from .flatmodel import action_layout, obs_layout
from .spaces import Box, Dict as DictSpace


class KManipVectorEnv:
    def __init__(self, env_id: str, num_envs: int, device: int = 0, dtype: str = "float32", seed: int = 0, env0: int = 0,
                 max_episode_steps: int = K.MAX_EPISODE_STEPS, log_dir: Optional[str] = None, log_env_ids=None,
                 copy: bool = True, **sim_kwargs):
        """copy=True (default, what gymnasium's VectorEnv does): reset() / step() return fresh tensors, so an observation
        kept from one step is not overwritten by the next.  copy=False returns views into the simulator's output buffers
        (zero extra kernels; the next step overwrites them -- only for loops that consume an observation before stepping)."""
        kw = K.ENV_REGISTRY[env_id]
        self.env_id, self.num_envs = env_id, int(num_envs)
        self.copy = bool(copy)
        self.sim = BatchSim(env_id, num_envs, device=device, dtype=dtype, seed=seed, env0=env0,
                            max_episode_steps=max_episode_steps, **sim_kwargs)
        self.device = self.sim.device
        self.obs_list, self.act_list = list(kw["obs_list"]), list(kw["act_list"])
        n_r = len(kw["q_id_r_mask"]) if kw.get("q_id_r_mask") is not None else 0
        n_l = len(kw["q_id_l_mask"]) if kw.get("q_id_l_mask") is not None else 0
        self.action_layout = action_layout(self.act_list, n_r, n_l)
        self.obs_layout = {k: v for k, v in obs_layout(self.sim.q_len).items() if k in self.obs_list}
        # camera observations of the Vision ids (reference env_base.py:140-147): one uint8 image batch per camera, rendered
        # from the post-step state by the ray-casting kernels (csrc/km_render.cuh)
        self.cameras = [K.CAMERAS[o.split("/")[-1]] for o in self.obs_list if "camera" in o]
        spaces_ = [(k, Box(-1, 1, (sl.stop - sl.start,), K.OBS_DTYPE)) for k, sl in self.obs_layout.items()]
        spaces_ += [(c.log_name, Box(c.low, c.high, (c.h, c.w, 3), c.dtype)) for c in self.cameras]
        self.single_observation_space = DictSpace(OrderedDict(spaces_))
        self._images = {c.log_name: self.sim.torch.empty(self.num_envs, c.h, c.w, 3, dtype=self.sim.torch.uint8, device=self.sim.device)
                        for c in self.cameras}
        self.single_action_space = DictSpace(OrderedDict(
            (k, Box(-1, 1, (sl.stop - sl.start,), K.ACT_DTYPE)) for k, sl in self.action_layout.items()))
        # batched spaces as gymnasium.vector.VectorEnv exposes them (leading num_envs axis)
        self.observation_space = DictSpace(OrderedDict(
            (k, Box(float(sp.low.min()), float(sp.high.max()), (self.num_envs,) + tuple(sp.shape), sp.dtype))
            for k, sp in self.single_observation_space.spaces.items()))
        self.action_space = DictSpace(OrderedDict(
            (k, Box(-1, 1, (self.num_envs,) + tuple(sp.shape), sp.dtype)) for k, sp in self.single_action_space.spaces.items()))
        t = self.sim.torch
        self._act = t.zeros(self.num_envs, self.sim.act_dim, dtype=t.float32, device=self.device)
        # episode logger (reference log_h5py.py) fed from device ring buffers; log_env_ids: local env indices (default: env 0)
        self.log = None
        if log_dir is not None:
            from .log_episode import BatchEpisodeLog
            ids = [0] if log_env_ids is None else list(log_env_ids)
            self.log = BatchEpisodeLog(log_dir, ids, self.sim.q_len, len(self.action_layout), self.action_layout["grip_r"].start,
                                       t, self.device, env0=env0,
                                       info={"env_id": env_id, "obs_list": self.obs_list, "act_list": self.act_list})

    # -------------------------------------------------------------------------------- helpers
    def _obs_dict(self, flat, images: bool = True) -> "OrderedDict[str, object]":
        if self.copy:
            flat = flat.clone()
        obs = OrderedDict((k, flat[:, sl]) for k, sl in self.obs_layout.items())
        if images:
            for c in self.cameras:          # images of the stored (post-step, post-autoreset) state
                obs[c.log_name] = self.sim.render(c, out=None if self.copy else self._images[c.log_name])
        return obs

    def _out(self, x):
        return x.clone() if self.copy else x

    @property
    def episode_return(self):
        """Running return of every env (accumulated by the step kernel)."""
        return self.sim.episode_return

    @property
    def totals(self):
        """Rollout totals [sum of rewards, env steps, finished episodes, success steps] accumulated by the step kernel
        (km_episode_stats): a float64 CUDA tensor."""
        return self.sim.episode_stats()

    def flatten_action(self, action):
        """Dict of [n, k] tensors (reference keys) -> the flat [n, act_dim] float32 record; flat tensors pass through."""
        t = self.sim.torch
        if isinstance(action, dict):
            for k, sl in self.action_layout.items():
                self._act[:, sl] = t.as_tensor(action[k], device=self.device).to(t.float32).view(self.num_envs, -1)
            return self._act
        return action

    def sample_actions(self, generator=None):
        t = self.sim.torch
        return t.rand(self.num_envs, self.sim.act_dim, device=self.device, generator=generator) * 2 - 1

    # -------------------------------------------------------------------------------- episode API
    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        mask = cube = None
        if options:
            mask, cube = options.get("reset_mask"), options.get("cube_xyz")
        flat = self.sim.reset(mask=mask, cube_xyz=cube)     # (the reset kernel also zeroes the running returns)
        return self._obs_dict(flat), {}

    def step(self, action):
        """One env step of every env: ONE kernel launch.  Returns (obs, reward, terminated, truncated, info); info holds
        the batched form of the reference's per-step info (env_base.py:243-250): is_success, episode, step, sim_time, plus
        final_obs / final_return of the episodes that ended on this step, episode_return, con_flags, ncon."""
        act = self.flatten_action(action)
        sim = self.sim
        flat, rew, term, trunc = sim.step(act, autoreset=True)
        if self.log is not None:
            self.log.step(act, flat, sim.final_obs, trunc)
        o = self._out
        # flags: bool tensors in copy mode (a conversion is a fresh tensor already); the simulator's uint8 buffers otherwise
        f = (lambda x: x.bool()) if self.copy else (lambda x: x)
        info: Dict[str, object] = {
            "is_success": f(sim.is_success), "episode": o(sim.episode), "step": o(sim.step_count), "sim_time": o(sim.sim_time),
            "final_obs": self._obs_dict(sim.final_obs, images=False), "final_return": o(sim.final_return),
            "episode_return": o(sim.episode_return), "con_flags": o(sim.con_flags), "ncon": o(sim.ncon),
        }
        return self._obs_dict(flat), o(rew), f(term), f(trunc), info

    @property
    def sim_step_counts(self):
        _, step, _ = self.sim.get_state()
        return step

    def close(self):
        self.sim.close()
