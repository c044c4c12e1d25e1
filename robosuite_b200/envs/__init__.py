from .base import REGISTERED_ENVS, BatchedMujocoEnv, make  # noqa: F401
from .lift import BatchedLift  # noqa: F401
from .stack import BatchedStack  # noqa: F401
