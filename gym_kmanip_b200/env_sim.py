"""Simulation backend behind the reference's seam: ``new(gym_env)`` -> object with k_reset / k_step / k_render / k_close.

Mirrors reference gym_kmanip/env_sim.py:182-211 (KManipEnvSim + new).  Where the reference drives dm_control's
episode loop around MuJoCo, this backend owns one km_handle with a single env on the GPU and makes one fused
kernel launch per env step; every number comes from the CUDA library (gym_kmanip_b200/csrc) -- there is no CPU
path.  Each call returns the reference's 5-tuple ``(terminated, reward, discount, observation, sim_time)``.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional

import numpy as np

from . import constants as K
from .batch_sim import BatchSim
from .flatmodel import action_layout, obs_layout


class _SiteView:
    def __init__(self, xpos, xmat):
        self.xpos, self.xmat = xpos, xmat


class _Data:
    """Read-only stand-in for ``physics.data`` with the fields the reference's callers touch
    (examples/1_control.py:26, examples/2_synthetic_data.py:33-34, examples/4_teleop.py:31,103)."""

    def __init__(self, backend: "KManipEnvSim"):
        self._b = backend

    def _state(self):
        st, _, _ = self._b.sim.get_state()
        return st[0].double().cpu().numpy()

    def _field(self, name):
        return self._state()[self._b.sim.state_slices()[name]]

    @property
    def qpos(self):
        return self._field("qpos")

    @property
    def qvel(self):
        return self._field("qvel")

    @property
    def ctrl(self):
        return self._field("ctrl")

    @property
    def time(self):
        return float(self._field("time")[0])

    @property
    def mocap_pos(self):
        return self._field("mocap").reshape(-1, 7)[:, :3]

    @property
    def mocap_quat(self):
        return self._field("mocap").reshape(-1, 7)[:, 3:]

    @property
    def ncon(self):
        ncon, _ = self._b.sim.contacts()
        return int(ncon[0].item())

    def site(self, name: str) -> _SiteView:
        arm = {"eer_site_pos": 0, "eel_site_pos": 1}[name]
        pos, mat = self._b.sim.site_poses()
        if arm >= pos.shape[1]:
            raise KeyError(f"site {name} is not tracked for this action list")
        return _SiteView(pos[0, arm].double().cpu().numpy(), mat[0, arm].double().cpu().numpy().reshape(9))


class _Model:
    """Read-only stand-in for ``physics.model`` (flat mjModel-named arrays of the completed scene)."""

    def __init__(self, flat: dict):
        self._flat = flat
        for k in ("nq", "nv", "nu", "nbody", "njnt", "ngeom", "nsite", "nmocap"):
            setattr(self, k, flat[k])
        self.jnt_range = np.array(flat["jnt_range"], dtype=np.float64)
        self.opt_timestep = flat["opt"]["timestep"]

    def name2id(self, name: str, kind: str) -> int:
        return self._flat[f"{kind}_name"].index(name)

    def id2name(self, idx: int, kind: str) -> str:
        return self._flat[f"{kind}_name"][idx]


class _Physics:
    def __init__(self, backend: "KManipEnvSim"):
        self.data = _Data(backend)
        self.model = _Model(backend.sim.flat)


class KManipEnvSim:
    """One simulated env (n = 1) on the GPU with the reference's ``k_*`` protocol."""

    def __init__(self, gym_env, device: int = 0, dtype: str = "float64", ik_mode: int = 1):
        self.gym_env = gym_env
        kwargs = dict(
            mjcf_filename=gym_env.mjcf_filename, q_pos_home=gym_env.q_pos_home, q_id_r_mask=gym_env.q_id_r_mask,
            q_id_l_mask=gym_env.q_id_l_mask, ctrl_id_r_grip=gym_env.ctrl_id_r_grip, ctrl_id_l_grip=gym_env.ctrl_id_l_grip,
            obs_list=gym_env.obs_list, act_list=gym_env.act_list)
        # truncation is the TimeLimit wrapper's job for the single-env class (reference __init__.py:28,247), so the
        # backend itself never truncates
        self.sim = BatchSim("custom", 1, device=device, dtype=dtype, seed=gym_env.seed, env_kwargs=kwargs,
                            max_episode_steps=2 ** 30, ik_mode=ik_mode)
        n_r = 0 if gym_env.q_id_r_mask is None else len(gym_env.q_id_r_mask)
        n_l = 0 if gym_env.q_id_l_mask is None else len(gym_env.q_id_l_mask)
        self._act_layout = action_layout(gym_env.act_list, n_r, n_l)
        self._obs_layout = obs_layout(self.sim.q_len)
        self._act = np.zeros((1, self.sim.act_dim), dtype=np.float32)
        self._h_obs = np.zeros((1, self.sim.obs_dim), dtype=np.float64 if dtype == "float64" else np.float32)
        self._h_rew = np.zeros(1, dtype=self._h_obs.dtype)
        self._h_trunc = np.zeros(1, dtype=np.uint8)
        self.physics = _Physics(self)
        self._physics = self.physics

    # -- reference env_sim.py:110-146: obs dict in obs_list order, float64, fresh copies
    def _obs_dict(self) -> "OrderedDict[str, np.ndarray]":
        obs = OrderedDict()
        for k in self.gym_env.obs_list:
            if k in self._obs_layout:
                obs[k] = self._h_obs[0, self._obs_layout[k]].astype(K.OBS_DTYPE)
        for cam in self.gym_env.cameras:      # env_sim.py:140-145
            obs[cam.log_name] = self.k_render(cam)
        return obs

    def k_render(self, cam):
        """reference env_sim.py:187-188: uint8 [h, w, 3] image of camera `cam` (a constants.Cam)."""
        img = np.empty((1, cam.h, cam.w, 3), dtype=np.uint8)
        self.sim.render_host(cam, img)
        return img[0]

    def k_reset(self, cube_xyz: Optional[np.ndarray] = None):
        xyz = None if cube_xyz is None else np.ascontiguousarray(cube_xyz, dtype=self._h_obs.dtype).reshape(1, 3)
        self.sim.reset_host(None, xyz, self._h_obs)
        return False, None, None, self._obs_dict(), 0.0

    def k_step(self, action):
        self._act[:] = 0
        for k, sl in self._act_layout.items():
            if k in action:
                self._act[0, sl] = np.asarray(action[k], dtype=np.float32).reshape(-1)
        self.sim.step_host(self._act, self._h_obs, self._h_rew, self._h_trunc, autoreset=False)
        sim_time = self.physics.data.time
        return False, float(self._h_rew[0]), 1.0, self._obs_dict(), sim_time

    def k_close(self):
        self.sim.close()


def new(gym_env, device: int = 0, dtype: str = "float64", ik_mode: int = 1) -> KManipEnvSim:
    """Factory the env class calls through the backend seam (reference env_base.py:192-196, env_sim.py:206-211)."""
    return KManipEnvSim(gym_env, device=device, dtype=dtype, ik_mode=ik_mode)
