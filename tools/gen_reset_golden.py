"""Reset-state statistics of the reference stack (running on oracle/mujoco_shim): for each task, the qpos after env.reset() over
many resets -> per-coordinate min / max / mean / std in tests/golden/reset_golden.npz.  tests compare the batched samplers'
output distributions with them.  Build container only.  Usage: python tools/gen_reset_golden.py"""
import os, sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_env_golden as g  # noqa: E402


def collect(task, n=250, seed=0):
    import robosuite as suite

    env = suite.make(task, robots="Panda", has_renderer=False, has_offscreen_renderer=False, use_camera_obs=False,
                     hard_reset=False, control_freq=20, seed=seed)
    qs = []
    for _ in range(n):
        env.reset()
        qs.append(np.array(env.sim.data.qpos))
    return np.array(qs)


if __name__ == "__main__":
    g.install()
    out = {}
    for task in ("Lift", "Stack", "NutAssemblyRound", "PickPlace", "Door"):
        q = collect(task)
        out[task + "/min"], out[task + "/max"], out[task + "/mean"], out[task + "/std"] = q.min(0), q.max(0), q.mean(0), q.std(0)
        out[task + "/n"] = np.array(len(q))
        print(task, q.shape, "std>0 coords:", int((q.std(0) > 1e-9).sum()))
    np.savez_compressed(os.path.join(g.ROOT, "tests", "golden", "reset_golden.npz"), **out)
