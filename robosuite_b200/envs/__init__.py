from .base import REGISTERED_ENVS, BatchedMujocoEnv, make  # noqa: F401
from .lift import BatchedLift  # noqa: F401
from .stack import BatchedStack  # noqa: F401
from .door import BatchedDoor  # noqa: F401
from .nut_assembly import BatchedNutAssembly, BatchedNutAssemblyRound, BatchedNutAssemblySquare  # noqa: F401
from .pick_place import BatchedPickPlace  # noqa: F401
