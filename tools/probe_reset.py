"""Where does the time of an in-step reset go?  (BatchedGymWrapper.step with horizon-500 episodes at staggered phases)"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import robosuite_b200 as suite
from robosuite_b200.wrappers import BatchedGymWrapper

task = sys.argv[1] if len(sys.argv) > 1 else "Lift"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
env = suite.make(task, robots="Panda", num_envs=n, seed=1, horizon=500)
sim = env.sim
g = torch.Generator(device="cuda"); g.manual_seed(0)
def act():
    return torch.rand((n, env.action_dim), generator=g, device="cuda") * 2 - 1
for _ in range(60):
    sim.env_step(act(), 25)
torch.cuda.synchronize()
def timeit(name, fn, reps=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    print("%-40s %8.2f ms" % (name, dt * 1e3), flush=True)
a = act()
timeit("sim.env_step (pipeline graph)", lambda: sim.env_step(a, 25))
timeit("sim.forward (fused, export)", lambda: sim.forward())
timeit("sim.env_step after forward", lambda: sim.env_step(a, 25))
mask = torch.zeros(n, dtype=torch.bool, device="cuda"); mask[::512] = True
timeit("env._sample_reset_state(8)", lambda: env._sample_reset_state(8))
timeit("env.reset(mask of 8)", lambda: env.reset(mask=mask))
timeit("sim.ctrl_reset(mask)", lambda: sim.ctrl_reset(mask.to(torch.uint8)))
timeit("env.step (no reset)", lambda: env.step(a))
w = BatchedGymWrapper(env)
env.timestep[:] = torch.randint(0, env.horizon, (n,), generator=g, device="cuda")
env._max_steps_since_reset = env.horizon
timeit("wrapper.step with staggered resets", lambda: w.step(a), reps=10)
