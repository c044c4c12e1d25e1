import sys, torch
sys.path.insert(0, ".")
import robosuite_b200 as suite
env = suite.make("Lift", robots="Panda", num_envs=4096, seed=1, horizon=10**9)
sim = env.sim
g = torch.Generator(device="cuda"); g.manual_seed(0)
# pre-roll in fused mode (1 launch per step keeps ncu's skip count small), then one pipelined step
for t in range(100):
    sim.env_step(torch.rand((4096, 7), generator=g, device="cuda") * 2 - 1, 25)
torch.cuda.synchronize()
sim.set_mode(1)
a = torch.rand((4096, 7), generator=g, device="cuda") * 2 - 1
for _ in range(3):
    sim.env_step(a, 25)
torch.cuda.synchronize()
print("done")
