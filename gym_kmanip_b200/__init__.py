"""gym_kmanip_b200 -- B200-native batched simulator for the gym-kmanip env-step hot path.

Import surface mirrors the reference package (constants at module level, reference
gym_kmanip/__init__.py:11-222; env ids of its register() calls via ``make``).  The CUDA extension and
torch are only imported when an env is constructed.
"""
from .constants import *          # noqa: F401,F403
from .constants import ENV_REGISTRY, ACTION_KEY_ORDER   # noqa: F401

__all__ = [n for n in dir() if not n.startswith("_")]


def make(env_id: str, **kwargs):
    """Counterpart of ``gym.make(id)`` for the ids registered at reference __init__.py:244-483."""
    from .env_base import make as _make
    return _make(env_id, **kwargs)


def make_vec(env_id: str, num_envs: int, **kwargs):
    """Batched env (all envs advanced by one fused CUDA launch per step)."""
    from .vector_env import KManipVectorEnv
    return KManipVectorEnv(env_id, num_envs, **kwargs)
