"""Wall-clock of the unmodified reference Python stack stepping on the CPU oracle through oracle/mujoco_shim (BASELINE config 1:
one Lift/Panda environment, OSC_POSE, random actions).  'reference Python + restated CPU engine (not Google MuJoCo)'.
Build container only (needs /root/reference)."""
import sys, time
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 1)[0])
import gen_env_golden as g

g.install()
import robosuite as suite

env = suite.make("Lift", robots="Panda", has_renderer=False, has_offscreen_renderer=False, use_camera_obs=False,
                 hard_reset=False, horizon=500, control_freq=20, seed=0)
env.reset()
low, high = env.action_spec
rng = np.random.default_rng(0)
for _ in range(20):
    env.step(rng.uniform(low, high))
n = 200
t0 = time.perf_counter()
for _ in range(n):
    env.step(rng.uniform(low, high))
dt = time.perf_counter() - t0
print("reference stack on oracle shim: %.1f env-steps/s (1 process, %d steps, %.2f s)" % (n / dt, n, dt))
