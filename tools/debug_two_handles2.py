"""Bisect the two-handle divergence: same scenario as tests/test_gpu_boundary.py::test_two_handles..., one variant per process.
usage: python tools/debug_two_handles2.py <mode> <variant>   variant: conc | seq | swap | onlyB"""
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

mode = int(sys.argv[1]); variant = sys.argv[2]
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


def setup(stream):
    h = T._create(L, blob, n, 0)
    if stream is not None:
        assert L.b2s_set_stream(h, C.c_void_p(stream.cuda_stream)) == 0
    c = T._lift_osc_cfg(L, h)
    assert L.b2s_ctrl_config(h, C.byref(c)) == 0
    assert L.b2s_set_export(h, 0) == 0 and L.b2s_set_mode(h, mode) == 0
    with torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(torch.cuda.current_stream()):
        T._arr(L, h, "qpos").copy_(torch.as_tensor(q, dtype=torch.float32))
        assert L.b2s_forward(h) == 0 and L.b2s_ctrl_reset(h, None) == 0
    return h


def run(h, st):
    for t in range(steps):
        with torch.cuda.stream(st):
            assert L.b2s_env_step(h, C.c_void_p(acts_d[t].data_ptr()), 25) == 0


sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
ha, hb = setup(sa), setup(sb)
torch.cuda.synchronize()
if variant == "conc":
    for t in range(steps):
        with torch.cuda.stream(sa):
            assert L.b2s_env_step(ha, C.c_void_p(acts_d[t].data_ptr()), 25) == 0
        with torch.cuda.stream(sb):
            assert L.b2s_env_step(hb, C.c_void_p(acts_d[t].data_ptr()), 25) == 0
elif variant == "seq":
    run(ha, sa); torch.cuda.synchronize(); run(hb, sb)
elif variant == "swap":
    for t in range(steps):
        with torch.cuda.stream(sb):
            assert L.b2s_env_step(hb, C.c_void_p(acts_d[t].data_ptr()), 25) == 0
        with torch.cuda.stream(sa):
            assert L.b2s_env_step(ha, C.c_void_p(acts_d[t].data_ptr()), 25) == 0
elif variant == "onlyB":
    run(hb, sb); torch.cuda.synchronize(); run(ha, sa)
torch.cuda.synchronize()
qa, qb = T._arr(L, ha, "qpos").clone(), T._arr(L, hb, "qpos").clone()
L.b2s_destroy(ha); L.b2s_destroy(hb)
hc = setup(None)
run(hc, torch.cuda.current_stream())
torch.cuda.synchronize()
qc = T._arr(L, hc, "qpos").clone()


def where(x, y):
    return torch.nonzero((x != y).any(1)).flatten().tolist()


print(f"mode {mode} {variant}: A!=B {where(qa, qb)[:4]} A!=alone {where(qa, qc)[:4]} B!=alone {where(qb, qc)[:4]}")
