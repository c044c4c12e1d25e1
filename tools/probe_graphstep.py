"""Steady-state time of one pipelined control step in CUDA-graph mode (events), for timing experiments with env switches."""
import sys, torch
sys.path.insert(0, ".")
import robosuite_b200 as suite
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = suite.make("Lift", robots="Panda", num_envs=n, seed=1, horizon=10**9, kernel_mode="fused")
sim = env.sim
g = torch.Generator(device="cuda"); g.manual_seed(0)
for t in range(100):
    sim.env_step(torch.rand((n, 7), generator=g, device="cuda") * 2 - 1, 25)
torch.cuda.synchronize()
sim.set_mode(1)
a = torch.rand((n, 7), generator=g, device="cuda") * 2 - 1
sim.env_step(a, 25); torch.cuda.synchronize()
ts = []
for _ in range(4):
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); sim.env_step(a, 25); t1.record(); torch.cuda.synchronize()
    ts.append(t0.elapsed_time(t1))
print("n=%d step ms:" % n, " ".join("%.2f" % t for t in ts))
