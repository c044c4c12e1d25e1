"""Fixture generator: run the reference's own Python MJCF composer (robosuite @ /root/reference) with a stub
`mujoco` module and capture the composed MJCF that the reference would hand to MjSim.from_xml_string
(environments/base.py:255-275).  Runs only in the build container (needs /root/reference); the outputs are
committed under tests/golden/mjcf/.

Usage: python tools/gen_mjcf_fixtures.py
"""
import os, sys, types
from unittest.mock import MagicMock

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "mjcf")


def install_stubs():
    mj = types.ModuleType("mujoco")
    mj.__version__ = "3.3.0"

    class MjModel:  # noqa
        @classmethod
        def from_xml_string(cls, *a, **k):
            return MagicMock()

        @classmethod
        def from_xml_path(cls, *a, **k):
            return MagicMock()

    class MjData:  # noqa
        pass

    mj.MjModel = MjModel
    mj.MjData = MjData
    mj.__getattr__ = lambda name: MagicMock()
    sys.modules["mujoco"] = mj
    tc = types.ModuleType("termcolor")
    tc.colored = lambda s, *a, **k: s
    sys.modules["termcolor"] = tc
    for m in ["mujoco.viewer", "qpsolvers", "pynput", "pynput.keyboard", "hid", "mink", "h5py", "gymnasium",
              "gymnasium.spaces", "gymnasium.core"]:
        sys.modules[m] = MagicMock()
    os.environ["NUMBA_DISABLE_JIT"] = "1"
    sys.path.insert(0, REF)


class _Captured(Exception):
    pass


def compose(task, robot, seed=0):
    import robosuite as suite
    from robosuite.environments.base import MujocoEnv

    box = {}

    def fake_init_sim(self, xml_string=None):
        xml = xml_string if xml_string else self.model.get_xml()
        for p in getattr(self, "_xml_processors", []) or []:
            xml = p(xml)
        box["xml"] = xml
        raise _Captured()

    orig = MujocoEnv._initialize_sim
    MujocoEnv._initialize_sim = fake_init_sim
    try:
        suite.make(task, robots=robot, has_renderer=False, has_offscreen_renderer=False, use_camera_obs=False,
                   seed=seed)
    except _Captured:
        pass
    finally:
        MujocoEnv._initialize_sim = orig
    return box["xml"]


if __name__ == "__main__":
    install_stubs()
    os.makedirs(OUT, exist_ok=True)
    combos = [("Lift", "Panda"), ("Lift", "Sawyer"), ("Stack", "Sawyer"), ("Stack", "Panda"),
              ("NutAssemblyRound", "Panda"), ("Door", "Panda"), ("PickPlace", "Panda")]
    for task, robot in combos:
        xml = compose(task, robot)
        fn = os.path.join(OUT, f"{task}_{robot}.xml")
        with open(fn, "w") as f:
            f.write(xml)
        print(fn, len(xml))
