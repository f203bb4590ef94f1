"""jobshoplab_b200: B200-native batched JobShopLab env-step path.

    from jobshoplab_b200 import JobShopLabEnv, BatchedJobShopLabEnv, load_config

mirrors ``from jobshoplab import JobShopLabEnv, load_config`` (jobshoplab/__init__.py:1-3)."""
from .config import Config, load_config  # noqa: F401


def __getattr__(name):
    if name in ("JobShopLabEnv", "BatchedJobShopLabEnv"):
        from . import env

        return getattr(env, name)
    if name == "SB3VecEnv":
        from . import vec_env

        return vec_env.SB3VecEnv
    raise AttributeError(name)
