"""Exception classes with the reference's names (robosuite/utils/errors.py), so `except XMLError` etc. written against the
reference keeps working.  B2SError (the C library's error codes, robosuite_b200/engine.py) derives from SimulationError."""


class robosuiteError(Exception):
    """Base class for exceptions in robosuite_b200 (same name as the reference's base class)."""


class XMLError(robosuiteError):
    """Raised when the MJCF handed to the model compiler is malformed or uses an element this engine does not implement."""


class SimulationError(robosuiteError):
    """Raised when the simulation itself fails (device errors, invalid state)."""


class RandomizationError(robosuiteError):
    """Raised when a placement sampler cannot find a valid placement."""
