"""B200-native batched manipulation simulator keeping robosuite's make / reset / step / controller_config surface."""
__version__ = "0.1.0"


def make(env_name, *args, **kwargs):
    from .envs import make as _make

    return _make(env_name, *args, **kwargs)
