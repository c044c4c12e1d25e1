import sys, torch
sys.path.insert(0, ".")
import robosuite_b200 as suite
n = 1024
env = suite.make("Lift", robots="Panda", num_envs=n, seed=1, horizon=10**9, kernel_mode="fused")
sim = env.sim
g = torch.Generator(device="cuda"); g.manual_seed(0)
for t in range(100):
    sim.env_step(torch.rand((n, 7), generator=g, device="cuda") * 2 - 1, 25)
torch.cuda.synchronize()
sim.set_mode(1)
a = torch.rand((n, 7), generator=g, device="cuda") * 2 - 1
print("=== pipeline step", flush=True)
sim.env_step(a, 2)
torch.cuda.synchronize()
print(model_names := [(i, nm) for i, nm in enumerate(env.model.names["geom"]) if nm and env.model.geom_contype[i] | env.model.geom_conaffinity[i]])
