"""Two handles stepping concurrently: after every control step compare every state array of the two (and of a third handle stepped
alone) and report the first divergence (array, environments, magnitude)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import test_gpu_boundary as T  # noqa: E402
from tests.util import lift_states, load  # noqa: E402
from robosuite_b200.mjcf.compiler import pack_model  # noqa: E402

L = T._lib()
model = load("Lift_Panda")
blob = pack_model(model)
n, steps = 64, 12
q, _ = lift_states(model, n, seed=31)
rng = np.random.default_rng(4)
acts = rng.uniform(-1, 1, size=(steps, n, 7))
acts[:, : n // 2, 2] = -1.0
acts[:, :, 6] = 1.0
acts_d = torch.as_tensor(acts, dtype=torch.float32, device="cuda")
os.environ["B2S_NO_GJK_CACHE"] = "1"
NAMES = ["qpos", "qvel", "qacc_warmstart", "ctrl", "ctrl_goal_pos", "ctrl_goal_ori", "ctrl_torque", "warn", "time"]


def setup(stream):
    h = T._create(L, blob, n, 0)
    if stream is not None:
        assert L.b2s_set_stream(h, C.c_void_p(stream.cuda_stream)) == 0
    c = T._lift_osc_cfg(L, h)
    assert L.b2s_ctrl_config(h, C.byref(c)) == 0
    assert L.b2s_set_export(h, 0) == 0 and L.b2s_set_mode(h, 1) == 0
    with torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(torch.cuda.current_stream()):
        T._arr(L, h, "qpos").copy_(torch.as_tensor(q, dtype=torch.float32))
        assert L.b2s_forward(h) == 0 and L.b2s_ctrl_reset(h, None) == 0
    return h


sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
ha, hb, hc = setup(sa), setup(sb), setup(None)
torch.cuda.synchronize()
nsub = int(os.environ.get("NSUB", "25"))
found = False
for t in range(steps):
    for k in range(25 // nsub):
        with torch.cuda.stream(sa):
            assert L.b2s_env_step(ha, C.c_void_p(acts_d[t].data_ptr()), nsub) == 0
        with torch.cuda.stream(sb):
            assert L.b2s_env_step(hb, C.c_void_p(acts_d[t].data_ptr()), nsub) == 0
        torch.cuda.synchronize()
        assert L.b2s_env_step(hc, C.c_void_p(acts_d[t].data_ptr()), nsub) == 0
        torch.cuda.synchronize()
        for nm in NAMES:
            try:
                a, b, c = (T._arr(L, h, nm) for h in (ha, hb, hc))
            except Exception:
                continue
            for tag, x, y in (("A vs B", a, b), ("A vs alone", a, c), ("B vs alone", b, c)):
                if not torch.equal(x, y):
                    d = (x.double() - y.double()).abs().reshape(n, -1)
                    envs = torch.nonzero(d.amax(1) > 0).flatten().tolist()
                    print(f"step {t}.{k} {nm}: {tag} differ in envs {envs[:16]} max |d| {float(d.max()):.3g}")
                    found = True
        if found:
            break
    if found:
        break
print("diverged" if found else "identical through all steps")
