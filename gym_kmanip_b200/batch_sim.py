"""BatchSim: thin Python owner of one km_handle (one device, N envs) with torch tensors for all buffers.

PyTorch is plumbing here (device memory, streams); every number is produced by the CUDA library.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import _lib, constants as K, flatmodel, mjcf

SCENE_ID = {"solo_arm": 0, "dual_arm": 1, "torso": 2}


class BatchSim:
    def __init__(self, env_id: str, num_envs: int, device: int = 0, dtype: str = "float32", seed: int = 0, env0: int = 0,
                 ik_iters: int = K.DEVICE_IK_ITERS, ik_teleport: bool = True, max_episode_steps: int = K.MAX_EPISODE_STEPS,
                 env_kwargs: Optional[dict] = None, ik_mode: int = 0, n_sub_steps: int = 0):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("gym_kmanip_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.torch = torch
        self.L = _lib.load()
        self.env_id = env_id
        self.kw = dict(K.ENV_REGISTRY[env_id]) if env_kwargs is None else dict(env_kwargs)
        self.scene = mjcf.scene_of_mjcf(self.kw["mjcf_filename"])
        self.flat = mjcf.load_flat(self.scene)
        self.pm = flatmodel.PackedModel(self.flat)
        self.task = flatmodel.make_task(self.flat, self.kw, ik_iters=ik_iters, ik_teleport=ik_teleport,
                                        max_episode_steps=max_episode_steps, ik_mode=ik_mode, n_sub_steps=n_sub_steps)
        self.n = int(num_envs)
        self.device_index = int(device)
        self.device = torch.device("cuda", self.device_index)
        if dtype in ("float32", 32, torch.float32):
            self.tdtype = torch.float32
        elif dtype in ("float64", 64, torch.float64):
            self.tdtype = torch.float64
        else:
            raise ValueError(f"dtype must be 'float32' or 'float64', got {dtype!r}")
        h = C.c_void_p()
        _lib.check(self.L.km_create(self.pm.ref(), C.byref(self.task), SCENE_ID[self.scene], self.n, self.device_index,
                                    32 if self.tdtype == torch.float32 else 64, seed, env0, C.byref(h)))
        self.h = h
        L = self.L
        self.nq, self.nv, self.nu, self.nmocap = L.km_nq(h), L.km_nv(h), L.km_nu(h), L.km_nmocap(h)
        self.obs_dim, self.act_dim, self.state_dim, self.max_contacts = (L.km_obs_dim(h), L.km_act_dim(h), L.km_state_dim(h),
                                                                         L.km_max_contacts(h))
        self.q_len = self.task.q_len
        kw = dict(device=self.device)
        self.obs = torch.zeros(self.n, self.obs_dim, dtype=self.tdtype, **kw)
        self.final_obs = torch.zeros(self.n, self.obs_dim, dtype=self.tdtype, **kw)
        self.reward = torch.zeros(self.n, dtype=self.tdtype, **kw)
        self.truncated = torch.zeros(self.n, dtype=torch.uint8, **kw)
        self.terminated = torch.zeros(self.n, dtype=torch.uint8, **kw)
        self.con_flags = torch.zeros(self.n, dtype=torch.int32, **kw)
        self.ncon = torch.zeros(self.n, dtype=torch.int32, **kw)
        self.con_geoms = torch.full((self.n, 2 * self.max_contacts), -1, dtype=torch.int32, **kw)
        # per-step `info` arrays (reference env_base.py:243-250), written by the step kernel itself
        self.is_success = torch.zeros(self.n, dtype=torch.uint8, **kw)
        self.episode_return = torch.zeros(self.n, dtype=self.tdtype, **kw)
        self.final_return = torch.zeros(self.n, dtype=self.tdtype, **kw)
        self.sim_time = torch.zeros(self.n, dtype=self.tdtype, **kw)
        self.step_count = torch.zeros(self.n, dtype=torch.int32, **kw)
        self.episode = torch.zeros(self.n, dtype=torch.int32, **kw)
        self._out_min = _lib.StepOut(self.obs.data_ptr(), self.final_obs.data_ptr(), self.reward.data_ptr(),
                                     self.truncated.data_ptr(), None, None, None, None, None, None, None, None, None, None)
        self._bind_outputs()

    def _bind_outputs(self):
        self._out = _lib.StepOut(self.obs.data_ptr(), self.final_obs.data_ptr(), self.reward.data_ptr(),
                                 self.truncated.data_ptr(), self.terminated.data_ptr(), self.con_flags.data_ptr(),
                                 self.ncon.data_ptr(), self.con_geoms.data_ptr(), self.is_success.data_ptr(),
                                 self.episode_return.data_ptr(), self.final_return.data_ptr(), self.sim_time.data_ptr(),
                                 self.step_count.data_ptr(), self.episode.data_ptr())

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.km_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def configure(self, lanes_per_env: int = 0, envs_per_block: int = 0):
        _lib.check(self.L.km_configure(self.h, lanes_per_env, envs_per_block))
        return self.launch_config()

    def set_env_ordering(self, mode: int = -1):
        """Cost-ordered walk of the step kernel (km_set_env_ordering): -1 automatic, 0 off, 1 on.  Results do not depend on it."""
        _lib.check(self.L.km_set_env_ordering(self.h, int(mode)))

    def launch_config(self) -> Dict[str, int]:
        v = [C.c_int(0) for _ in range(5)]
        _lib.check(self.L.km_launch_config(self.h, *[C.byref(x) for x in v]))
        return dict(zip(["lanes_per_env", "envs_per_block", "grid", "ctas_per_sm", "smem_bytes"], [x.value for x in v]))

    @property
    def launches(self) -> int:
        return int(self.L.km_launch_count(self.h))

    # ------------------------------------------------------------------ episode API
    def reset(self, mask=None, cube_xyz=None):
        """Reset all envs (or those with mask != 0). Returns the observation tensor [n, obs_dim] (a view, overwritten by step)."""
        t = self.torch
        mp = None
        if mask is not None:
            self._mask = t.as_tensor(mask).to(device=self.device, dtype=t.uint8).contiguous()
            if self._mask.numel() != self.n:
                raise ValueError(f"reset mask must have num_envs = {self.n} entries, got {self._mask.numel()}")
            mp = self._mask.data_ptr()
        xp = None
        if cube_xyz is not None:
            self._xyz = t.as_tensor(cube_xyz, device=self.device).to(self.tdtype).contiguous()
            if self._xyz.numel() != 3 * self.n:
                raise ValueError(f"cube_xyz must be [num_envs = {self.n}, 3]")
            xp = self._xyz.data_ptr()
        _lib.check(self.L.km_reset(self.h, mp, xp, self.obs.data_ptr(), self._stream()))
        return self.obs

    def step(self, action, autoreset: bool = True, contacts: bool = True):
        """One env step of every env.  action: float32 CUDA tensor [n, act_dim]."""
        t = self.torch
        if action.dtype != t.float32 or action.device != self.device or not action.is_contiguous():
            action = action.to(device=self.device, dtype=t.float32).contiguous()   # (also moves tensors of another GPU)
        if action.numel() != self.n * self.act_dim:
            raise ValueError(f"action must be [num_envs = {self.n}, act_dim = {self.act_dim}]")
        self._act = action
        _lib.check(self.L.km_step(self.h, action.data_ptr(), C.byref(self._out if contacts else self._out_min), int(autoreset),
                                  self._stream()))
        return self.obs, self.reward, self.terminated, self.truncated

    def step_host(self, action_host, obs_host, reward_host, truncated_host, autoreset: bool = True):
        """The reference-facing call with HOST buffers (numpy or pinned torch tensors): copies inside."""
        def ptr(x):
            return x.ctypes.data if isinstance(x, np.ndarray) else x.data_ptr()
        self.L.km_set_host_stream(self.h, self._stream())   # same stream as the device-pointer calls: ordered with them
        _lib.check(self.L.km_step_host(self.h, ptr(action_host), ptr(obs_host), ptr(reward_host), ptr(truncated_host),
                                       int(autoreset)))

    def reset_host(self, mask_host=None, cube_xyz_host=None, obs_host=None):
        """km_reset_host: reset with HOST buffers (numpy or pinned torch tensors; any may be None)."""
        def ptr(x):
            return None if x is None else (x.ctypes.data if isinstance(x, np.ndarray) else x.data_ptr())
        self.L.km_set_host_stream(self.h, self._stream())
        _lib.check(self.L.km_reset_host(self.h, ptr(mask_host), ptr(cube_xyz_host), ptr(obs_host)))

    # ------------------------------------------------------------------ state access (teacher forcing, physics shim)
    def get_state(self):
        t = self.torch
        st = t.empty(self.n, self.state_dim, dtype=self.tdtype, device=self.device)
        step = t.empty(self.n, dtype=t.int32, device=self.device)
        ep = t.empty(self.n, dtype=t.int32, device=self.device)
        _lib.check(self.L.km_get_state(self.h, st.data_ptr(), step.data_ptr(), ep.data_ptr(), self._stream()))
        return st, step, ep

    def set_state(self, state, step=None, episode=None):
        t = self.torch
        st = t.as_tensor(state, device=self.device).to(self.tdtype).contiguous().view(self.n, self.state_dim)
        sp = ep = None
        keep = [st]
        if step is not None:
            s = t.as_tensor(step, device=self.device).to(t.int32).contiguous()
            keep.append(s)
            sp = s.data_ptr()
        if episode is not None:
            e = t.as_tensor(episode, device=self.device).to(t.int32).contiguous()
            keep.append(e)
            ep = e.data_ptr()
        _lib.check(self.L.km_set_state(self.h, st.data_ptr(), sp, ep, self._stream()))
        self._keep = keep

    def state_slices(self) -> Dict[str, slice]:
        o, out = 0, {}
        for k, n in (("qpos", self.nq), ("qvel", self.nv), ("ctrl", self.nu), ("warm", self.nv), ("mocap", 7 * self.nmocap),
                     ("time", 1), ("cube_lo", 3)):
            out[k] = slice(o, o + n)
            o += n
        return out

    def contacts(self):
        _lib.check(self.L.km_contacts(self.h, self.ncon.data_ptr(), self.con_geoms.data_ptr(), self._stream()))
        return self.ncon, self.con_geoms

    def site_poses(self):
        """End-effector site positions [n, n_arm, 3] and rotation matrices [n, n_arm, 3, 3] of the stored state."""
        t = self.torch
        na = int(self.L.km_n_arm(self.h))
        pos = t.empty(self.n, na, 3, dtype=self.tdtype, device=self.device)
        mat = t.empty(self.n, na, 3, 3, dtype=self.tdtype, device=self.device)
        _lib.check(self.L.km_site_poses(self.h, pos.data_ptr(), mat.data_ptr(), self._stream()))
        return pos, mat

    # ------------------------------------------------------------------ camera observations (Vision ids)
    def render(self, cam, out=None):
        """Image of every env's stored state from camera `cam` (a constants.Cam or its name): uint8 CUDA tensor [n, h, w, 3]."""
        from . import render as R
        t = self.torch
        cam = K.CAMERAS[cam] if isinstance(cam, str) else cam
        key = (cam.name, cam.w, cam.h)
        if not hasattr(self, "_cams"):
            self._cams, self._visual = {}, R.visual_struct(self.flat)
        if key not in self._cams:
            self._cams[key] = R.camera_struct(self.flat, cam.name, cam.w, cam.h)
        if out is None:
            out = t.empty(self.n, cam.h, cam.w, 3, dtype=t.uint8, device=self.device)
        assert out.is_cuda and out.dtype == t.uint8 and out.is_contiguous() and out.numel() == self.n * cam.h * cam.w * 3
        _lib.check(self.L.km_render(self.h, C.byref(self._cams[key]), C.byref(self._visual), out.data_ptr(), self._stream()))
        return out

    def render_host(self, cam, out: np.ndarray):
        """km_render_host: the same with a HOST image buffer [n, h, w, 3] uint8 (copy inside)."""
        from . import render as R
        cam = K.CAMERAS[cam] if isinstance(cam, str) else cam
        self.L.km_set_host_stream(self.h, self._stream())
        _lib.check(self.L.km_render_host(self.h, C.byref(R.camera_struct(self.flat, cam.name, cam.w, cam.h)),
                                         C.byref(R.visual_struct(self.flat)), out.ctypes.data))
        return out

    def render_records(self):
        """The per-env render records of the last render() (float32 [n, rec_floats]); for tests of the pixel stage."""
        t = self.torch
        nf = int(self.L.km_render_record_floats(self.h))
        out = t.empty(self.n, nf, dtype=t.float32, device=self.device)
        _lib.check(self.L.km_get_render_records(self.h, out.data_ptr(), self._stream()))
        return out

    def episode_stats(self, reset: bool = False):
        """Rollout totals accumulated by the step kernel: float64 CUDA tensor [sum of rewards, env steps, finished episodes,
        success steps] since the last reset of the totals (km_episode_stats); reset=True zeroes them afterwards."""
        t = self.torch
        tot = t.empty(4, dtype=t.float64, device=self.device)
        _lib.check(self.L.km_episode_stats(self.h, tot.data_ptr(), None, int(reset), self._stream()))
        return tot

    def solver_stats(self):
        t = self.torch
        it = t.empty(self.n, dtype=t.int32, device=self.device)
        ls = t.empty(self.n, dtype=t.int32, device=self.device)
        _lib.check(self.L.km_solver_stats(self.h, it.data_ptr(), ls.data_ptr(), self._stream()))
        return it, ls
