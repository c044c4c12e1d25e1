"""Smoke of the unit-queue mode with the ring counters printed after every control step (B2S_UNIT_DEBUG=1)."""
import os
import sys

os.environ["B2S_UNIT_DEBUG"] = "1"
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import robosuite_b200 as suite  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
env = suite.make("Lift", robots="Panda", num_envs=n, seed=1, horizon=10 ** 9)
env.sim.set_mode(2)
gen = torch.Generator(device=env.device)
gen.manual_seed(3)
for i in range(steps):
    env.sim.env_step(torch.rand((n, env.action_dim), generator=gen, device=env.device, dtype=env.dtype) * 2 - 1, 25)
torch.cuda.synchronize()
print("ok", float(env.sim.qpos.abs().max()), int(env.sim.warn.abs().max()))
