"""Per-episode logger in the reference's ACT / LeRobot layout (reference gym_kmanip/log_h5py.py:13-61).

One file per episode with the datasets ``observations/qpos`` and ``observations/qvel`` (MAX_EPISODE_STEPS x q_len),
``action`` (MAX_EPISODE_STEPS x a_len), the attribute ``sim`` and a ``metadata`` group holding the serialisable entries
of ``info``.  The reference's quirks are kept on purpose: the logged "qpos" / "qvel" are the *normalised observations*
``q_pos`` / ``q_vel`` (log_h5py.py:56-57), and every row of ``action`` is the ``grip_r`` command broadcast over the
a_len columns (log_h5py.py:55; a_len is the number of action keys, env_base.py:191).

h5py is not part of the build image: when it imports, ``episode_<k>.hdf5`` is written exactly as the reference does;
otherwise the same arrays go to ``episode_<k>.npz`` under the same dataset names (``np.load(path)["observations/qpos"]``)
with the attributes in a JSON string ``__attrs__``.  Camera datasets belong to the Vision ids, which are out of scope.
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
        self.path = None

    def step(self, action: Dict[str, np.ndarray], observation: Dict[str, np.ndarray], info: Dict[str, Any]) -> None:
        i = int(info["step"]) - 1
        self.action[i] = action["grip_r"]                 # broadcast over a_len columns, as in the reference
        self.qpos[i] = observation["q_pos"]
        self.qvel[i] = observation["q_vel"]

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
                f.create_dataset("observations/qpos", data=self.qpos)
                f.create_dataset("observations/qvel", data=self.qvel)
                f.create_dataset("action", data=self.action)
        else:
            self.path = self.stem + ".npz"
            np.savez(self.path, **{"observations/qpos": self.qpos, "observations/qvel": self.qvel, "action": self.action,
                                   "__attrs__": np.array(json.dumps({"sim": self.attrs["sim"], "metadata": self.metadata}))})
        return self.path


def new(log_dir: str, info: Dict[str, Any]) -> EpisodeLog:
    return EpisodeLog(log_dir, info)


def step(f: EpisodeLog, action, observation, info) -> None:
    f.step(action, observation, info)


def end(f: EpisodeLog):
    return None if f is None else f.end()
