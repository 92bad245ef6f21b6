"""Per-episode logger in the reference's ACT / LeRobot layout (reference gym_kmanip/log_h5py.py:13-61).

One file per episode with the datasets ``observations/qpos`` and ``observations/qvel`` (MAX_EPISODE_STEPS x q_len),
``action`` (MAX_EPISODE_STEPS x a_len), the attribute ``sim`` and a ``metadata`` group holding the serialisable entries
of ``info``.  The reference's quirks are kept on purpose: the logged "qpos" / "qvel" are the *normalised observations*
``q_pos`` / ``q_vel`` (log_h5py.py:56-57), and every row of ``action`` is the ``grip_r`` command broadcast over the
a_len columns (log_h5py.py:55; a_len is the number of action keys, env_base.py:191).

h5py is not part of the build image: when it imports, ``episode_<k>.hdf5`` is written exactly as the reference does;
otherwise the same tree is written as ``episode_<k>.hdf5`` by the minimal pure-Python writer ``hdf5_min.py``
(fixed-length string / numeric attributes; bools as uint8).  The Vision ids also log their camera observations
(``observations/images/<cam.name>``, ``metadata/<cam.log_name>``: resolution, focal_length, principal_point) -- single-env
class only; the batched logger keeps to the joint-space datasets.  Files are written when the episode ends (the reference
flushes every step).
"""
from __future__ import annotations

import json
import os
from typing import Any, Dict

import numpy as np

from . import constants as K

try:                                  # pragma: no cover - not installed in the build image
    import h5py
    HAVE_H5PY = True
except Exception:                     # noqa: BLE001
    h5py = None
    HAVE_H5PY = False

H5PY_CHUNK_SIZE_BYTES = 1024 ** 2 * 2   # reference __init__.py:44


class EpisodeLog:
    def __init__(self, log_dir: str, info: Dict[str, Any]):
        assert os.path.exists(log_dir), f"Directory {log_dir} does not exist"
        self.q_len, self.a_len = int(info["q_len"]), int(info["a_len"])
        self.stem = os.path.join(log_dir, f"episode_{info['episode']}")
        self.attrs = {"sim": bool(info["sim"])}
        self.metadata = {}
        for key, value in info.items():
            try:
                json.dumps(value)
                self.metadata[key] = value
            except TypeError:
                pass                   # the reference prints "Could not save ..." and moves on (log_h5py.py:21-24)
        self.qpos = np.zeros((K.MAX_EPISODE_STEPS, self.q_len), dtype=np.float32)
        self.qvel = np.zeros((K.MAX_EPISODE_STEPS, self.q_len), dtype=np.float32)
        self.action = np.zeros((K.MAX_EPISODE_STEPS, self.a_len), dtype=np.float32)
        # camera observations of the Vision ids (reference log_h5py.py:36-46 `cam`, env_base.py:231-234): one uint8 dataset
        # per camera under observations/images/<cam.name> and its intrinsics as attributes of metadata/<cam.log_name>
        self.cams = list(info.get("cameras") or [])
        self.images = {c.name: np.zeros((K.MAX_EPISODE_STEPS, c.h, c.w, c.c), dtype=c.dtype) for c in self.cams}
        self.path = None

    def step(self, action: Dict[str, np.ndarray], observation: Dict[str, np.ndarray], info: Dict[str, Any]) -> None:
        i = int(info["step"]) - 1
        self.action[i] = action["grip_r"]                 # broadcast over a_len columns, as in the reference
        self.qpos[i] = observation["q_pos"]
        self.qvel[i] = observation["q_vel"]
        for c in self.cams:                                # log_h5py.py:59-60
            self.images[c.name][i] = observation[c.log_name]

    def _cam_attrs(self, c) -> Dict[str, Any]:
        return {"resolution": [c.w, c.h], "focal_length": c.fl, "principal_point": list(c.pp)}

    def end(self) -> str:
        if HAVE_H5PY:                                      # pragma: no cover
            self.path = self.stem + ".hdf5"
            with h5py.File(self.path, "w", rdcc_nbytes=H5PY_CHUNK_SIZE_BYTES) as f:
                f.attrs["sim"] = self.attrs["sim"]
                g = f.create_group("metadata")
                for key, value in self.metadata.items():
                    try:
                        g.attrs[key] = value
                    except TypeError:
                        pass
                f.create_group("observations/images")
                for c in self.cams:
                    gc = f.create_group(f"metadata/{c.log_name}")
                    for key, value in self._cam_attrs(c).items():
                        gc.attrs[key] = value
                    f.create_dataset(f"/observations/images/{c.name}", data=self.images[c.name], chunks=(1, c.h, c.w, c.c))
                f.create_dataset("observations/qpos", data=self.qpos)
                f.create_dataset("observations/qvel", data=self.qvel)
                f.create_dataset("action", data=self.action)
        else:
            # no h5py in this image: the same tree through the minimal HDF5 writer of this package (version 0 superblock,
            # symbol-table groups, contiguous datasets -- hdf5_min.py), so consumers of the ACT / LeRobot layout get .hdf5
            from . import hdf5_min
            self.path = self.stem + ".hdf5"
            meta = {}
            for key, value in self.metadata.items():
                try:
                    hdf5_min._attr_value(value)
                    meta[key] = value
                except (TypeError, ValueError):
                    pass               # "Could not save ..." in the reference (log_h5py.py:21-24)
            attrs = {"": {"sim": self.attrs["sim"]}, "metadata": meta}
            mtree: Dict[str, Any] = {}
            for c in self.cams:                            # metadata/<cam.log_name>: "camera/head" nests a group per component
                node, path = mtree, "metadata"
                for part in c.log_name.split("/"):
                    node = node.setdefault(part, {})
                    path += "/" + part
                attrs[path] = self._cam_attrs(c)
            hdf5_min.write(self.path, {"action": self.action, "metadata": mtree,
                                       "observations": {"images": {c.name: self.images[c.name] for c in self.cams},
                                                        "qpos": self.qpos, "qvel": self.qvel}}, attrs=attrs)
        return self.path


def read_episode(path: str, with_images: bool = False):
    """Reads an episode file written by this module: (qpos, qvel, action, root attrs, metadata attrs) and, with
    ``with_images``, two more items: {camera name: images}, {camera log name: intrinsics}.  Uses h5py when it imports, else
    the reader of hdf5_min (which parses exactly the subset of HDF5 that hdf5_min writes)."""
    if HAVE_H5PY:                                          # pragma: no cover
        with h5py.File(path, "r") as f:
            out = (f["observations/qpos"][:], f["observations/qvel"][:], f["action"][:], dict(f.attrs), dict(f["metadata"].attrs))
            if with_images:
                cams = {}
                f["metadata"].visititems(lambda nm, o: cams.__setitem__(nm, dict(o.attrs)) if len(o.attrs) else None)
                out += ({k: v[:] for k, v in f["observations/images"].items()}, cams)
            return out
    from . import hdf5_min
    tree, attrs = hdf5_min.read(path)
    out = (tree["observations"]["qpos"], tree["observations"]["qvel"], tree["action"], attrs.get("", {}), attrs.get("metadata", {}))
    if with_images:
        out += (tree["observations"]["images"], {k[len("metadata/"):]: v for k, v in attrs.items() if k.startswith("metadata/")})
    return out


def new(log_dir: str, info: Dict[str, Any]) -> EpisodeLog:
    return EpisodeLog(log_dir, info)


def step(f: EpisodeLog, action, observation, info) -> None:
    f.step(action, observation, info)


def end(f: EpisodeLog):
    return None if f is None else f.end()


class BatchEpisodeLog:
    """Episode logger of the batched env: device ring buffers, one file per finished episode of each logged env.

    ``KManipVectorEnv(..., log_dir=..., log_env_ids=[...])`` feeds it every step.  Each logged env owns one column of
    three ring buffers on the device (``q_pos`` / ``q_vel`` observations and the ``grip_r`` command, MAX_EPISODE_STEPS
    rows, float32 -- the arrays of reference log_h5py.py:27-43); a step appends one row per env with a single indexed
    store and nothing leaves the device until an env's episode ends (truncation), when that env's rows are copied to the
    host and written with the same dataset names, quirks and zero padding as the single-env logger above.  Files are
    named ``env<global id>_episode_<k>`` (the reference has one env, hence ``episode_<k>``).
    """

    def __init__(self, log_dir: str, env_ids, q_len: int, a_len: int, grip_col: int, torch, device, env0: int = 0,
                 info: Dict[str, Any] = None, max_steps: int = K.MAX_EPISODE_STEPS):
        assert os.path.exists(log_dir), f"Directory {log_dir} does not exist"
        t = self.torch = torch
        self.log_dir, self.q_len, self.a_len, self.grip_col, self.env0 = log_dir, int(q_len), int(a_len), int(grip_col), int(env0)
        self.max_steps = int(max_steps)
        self.ids = t.as_tensor(list(env_ids), dtype=t.long, device=device)
        n = self.ids.numel()
        self.cols = t.arange(n, device=device)
        self.qpos = t.zeros(self.max_steps, n, self.q_len, dtype=t.float32, device=device)
        self.qvel = t.zeros(self.max_steps, n, self.q_len, dtype=t.float32, device=device)
        self.grip = t.zeros(self.max_steps, n, dtype=t.float32, device=device)
        self.row = t.zeros(n, dtype=t.long, device=device)          # next free row of each column
        self.episode = [1] * n
        self.info = dict(info or {})
        self.paths = []

    def step(self, act_flat, obs_flat, final_obs_flat, done) -> int:
        """Append this step (the finished episode's last observation comes from ``final_obs_flat``); returns files written."""
        t = self.torch
        d = done[self.ids].bool()
        obs = t.where(d[:, None], final_obs_flat[self.ids], obs_flat[self.ids]).to(t.float32)
        r = self.row.clamp(max=self.max_steps - 1)
        keep = self.row < self.max_steps                             # the reference stops logging past MAX_EPISODE_STEPS
        rk, ck = r[keep], self.cols[keep]
        self.qpos[rk, ck] = obs[keep, :self.q_len]
        self.qvel[rk, ck] = obs[keep, self.q_len:2 * self.q_len]
        self.grip[rk, ck] = act_flat[self.ids, self.grip_col].to(t.float32)[keep]
        self.row += 1
        if not bool(d.any()):                                        # one flag read per step; rows stay on the device
            return 0
        cols = t.nonzero(d).flatten()
        qpos, qvel, grip = self.qpos[:, cols].cpu().numpy(), self.qvel[:, cols].cpu().numpy(), self.grip[:, cols].cpu().numpy()
        rows = self.row[cols].cpu().numpy()
        for k, c in enumerate(cols.tolist()):
            gid = self.env0 + int(self.ids[c])
            meta = dict(self.info, env=gid, episode=self.episode[c], step=int(rows[k]), q_len=self.q_len, a_len=self.a_len, sim=True)
            f = EpisodeLog.__new__(EpisodeLog)
            f.q_len, f.a_len, f.attrs, f.path = self.q_len, self.a_len, {"sim": True}, None
            f.cams, f.images = [], {}
            f.stem = os.path.join(self.log_dir, f"env{gid:06d}_episode_{self.episode[c]}")
            f.metadata = {key: v for key, v in meta.items() if _jsonable(v)}
            f.qpos, f.qvel = qpos[:, k], qvel[:, k]
            f.action = np.repeat(grip[:, k, None], self.a_len, axis=1)   # grip_r broadcast over a_len columns (log_h5py.py:55)
            self.paths.append(f.end())
            self.episode[c] += 1
        self.qpos[:, cols] = 0
        self.qvel[:, cols] = 0
        self.grip[:, cols] = 0
        self.row[cols] = 0
        return len(cols)


def _jsonable(v) -> bool:
    try:
        json.dumps(v)
        return True
    except TypeError:
        return False
