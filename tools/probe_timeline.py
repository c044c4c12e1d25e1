"""Event-timed launch timeline of one pipelined control step (B2S_TIMELINE=1): per-kernel mean time incl. launch gaps."""
import os, sys, time, torch
sys.path.insert(0, ".")
os.environ["B2S_TIMELINE"] = "1"
import robosuite_b200 as suite
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
env = suite.make("Lift", robots="Panda", num_envs=n, seed=1, horizon=10**9, kernel_mode="fused")
sim = env.sim
g = torch.Generator(device="cuda"); g.manual_seed(0)
for t in range(100):
    sim.env_step(torch.rand((n, 7), generator=g, device="cuda") * 2 - 1, 25)
torch.cuda.synchronize()
a = torch.rand((n, 7), generator=g, device="cuda") * 2 - 1
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record(); sim.env_step(a, 25); t1.record(); torch.cuda.synchronize()
print("fused step: %.3f ms" % t0.elapsed_time(t1))
sim.set_mode(1)
for _ in range(2):
    sim.env_step(a, 25)
    sys.stderr.flush()
