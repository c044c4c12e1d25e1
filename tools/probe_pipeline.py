"""ncu target: steady-state pipelined control steps of one task (default Lift/Panda, 4096 envs).  The pre-roll runs in FUSED mode
(one launch per control step, so ncu's -s skip count only has to skip the pipelined warm-up steps that follow), then N pipelined
steps run.  usage: python tools/probe_pipeline.py [task] [robot] [n_env] [controller] [pipelined_steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import robosuite_b200 as suite  # noqa: E402
from robosuite_b200 import controller_config as cc  # noqa: E402

task = sys.argv[1] if len(sys.argv) > 1 else "Lift"
robot = sys.argv[2] if len(sys.argv) > 2 else "Panda"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
ctrl = sys.argv[4] if len(sys.argv) > 4 else "OSC_POSE"
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
kw = {}
if ctrl != "OSC_POSE":
    kw["controller_configs"] = cc.refactor_composite_controller_config(cc.load_part_controller_config(ctrl), robot, ["right"])
env = suite.make(task, robots=robot, num_envs=n, seed=1, horizon=10 ** 9, **kw)
sim = env.sim
g = torch.Generator(device="cuda"); g.manual_seed(0)
sim.set_mode(0)
for t in range(int(os.environ.get("PREROLL", "100"))):
    sim.env_step(torch.rand((n, env.action_dim), generator=g, device="cuda") * 2 - 1, 25)
torch.cuda.synchronize()
sim.set_mode(1)
a = torch.rand((n, env.action_dim), generator=g, device="cuda") * 2 - 1
for _ in range(steps):
    sim.env_step(a, 25)
torch.cuda.synchronize()
print("done, warn", int(sim.warn.abs().max()))
