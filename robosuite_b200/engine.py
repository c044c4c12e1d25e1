"""ctypes binding of libb2s.so (the C ABI in include/b2s.h) with device arrays exposed as torch.cuda tensors.

Host-side mirror of the reference's engine shim `robosuite/utils/binding_utils.py` (MjSim: from_xml_string, reset,
forward, step, step1, step2, get_state/set_state), batched over `n_env` environments.  There is no CPU fallback:
construction fails loudly when the CUDA library or a GPU is missing.
"""
import ctypes as C
import os

import numpy as np

from .errors import SimulationError
from .mjcf.compiler import Model, compile_mjcf, pack_model

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

B2S_F32, B2S_F64, B2S_I32, B2S_I64 = 0, 1, 2, 3


class B2SError(SimulationError, RuntimeError):
    """non-zero return code of the C library (message from b2s_last_error)"""


class CtrlCfg(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("action_dim", C.c_int), ("n_arm", C.c_int), ("arm_dof", C.c_int * 8),
        ("arm_qpos", C.c_int * 8), ("arm_act", C.c_int * 8), ("eef_site", C.c_int), ("base_site", C.c_int),
        ("n_grip", C.c_int), ("grip_act", C.c_int * 4), ("grip_sign", C.c_double * 4), ("grip_speed", C.c_double),
        ("kp", C.c_double * 6), ("damping_ratio", C.c_double * 6), ("input_max", C.c_double * 6),
        ("input_min", C.c_double * 6), ("output_max", C.c_double * 6), ("output_min", C.c_double * 6),
        ("null_kp", C.c_double), ("uncouple_pos_ori", C.c_int), ("n_obs_site", C.c_int),
        ("jv_kp", C.c_double * 8), ("jv_ki", C.c_double * 8), ("jv_kd", C.c_double * 8), ("jv_in_max", C.c_double * 8),
        ("jv_in_min", C.c_double * 8), ("jv_out_max", C.c_double * 8), ("jv_out_min", C.c_double * 8),
        ("jv_vel_lo", C.c_double), ("jv_vel_hi", C.c_double), ("jv_use_vel_limits", C.c_int), ("jv_torque_comp", C.c_int),
    ]


def lib():
    global _LIB
    if _LIB is None:
        so = os.environ.get("B2S_LIB", os.path.join(_HERE, "libb2s.so"))
        if not os.path.exists(so):
            raise B2SError(f"{so} is missing: run `python -m robosuite_b200.build` (no CPU fallback exists)")
        L = C.CDLL(so)
        L.b2s_last_error.restype = C.c_char_p
        L.b2s_create.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.b2s_destroy.argtypes = [C.c_void_p]
        L.b2s_destroy.restype = None
        L.b2s_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.b2s_reset.argtypes = [C.c_void_p, C.c_void_p]
        for fn in ("b2s_forward", "b2s_step1", "b2s_step2"):
            getattr(L, fn).argtypes = [C.c_void_p]
        L.b2s_step.argtypes = [C.c_void_p, C.c_int]
        L.b2s_array.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                C.POINTER(C.c_int64)]
        L.b2s_jac_site.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.b2s_get_state.argtypes = [C.c_void_p, C.c_void_p]
        L.b2s_set_state.argtypes = [C.c_void_p, C.c_void_p]
        L.b2s_name2id.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.b2s_id2name.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.b2s_id2name.restype = C.c_char_p
        L.b2s_full_m.argtypes = [C.c_void_p, C.c_void_p]
        L.b2s_jac_body.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.b2s_jac_geom.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.b2s_ctrl_config.argtypes = [C.c_void_p, C.POINTER(CtrlCfg)]
        L.b2s_ctrl_reset.argtypes = [C.c_void_p, C.c_void_p]
        L.b2s_env_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.b2s_reset_envs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.b2s_body_pose_override.argtypes = [C.c_void_p, C.c_int]
        L.b2s_obs_config.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.b2s_task_config.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.b2s_task_config2.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.b2s_task_objects.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.b2s_timeline.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.b2s_task_table.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.b2s_set_export.argtypes = [C.c_void_p, C.c_int]
        L.b2s_set_profile.argtypes = [C.c_void_p, C.c_int]
        L.b2s_set_mode.argtypes = [C.c_void_p, C.c_int]
        L.b2s_launch_count.argtypes = [C.c_void_p]
        L.b2s_launch_count.restype = C.c_int64
        _LIB = L
    return _LIB


class _DevArray:
    """Minimal __cuda_array_interface__ carrier so torch can alias library-owned device memory without a copy."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}
        self._owner = owner


_TYPESTR = {B2S_F32: "<f4", B2S_F64: "<f8", B2S_I32: "<i4", B2S_I64: "<i8"}


_LIVE = None  # weak set of open BatchedSim objects (a device holds at most B2S_NSLOT = 8 live handles: descriptor slots)


def close_all():
    """Destroy every live handle of this process (test teardown, interpreter exit)."""
    for sim in list(_LIVE or ()):
        sim.close()


class BatchedSim:
    """n_env independent copies of one compiled model, stepped by the per-warp CUDA engine."""

    def __init__(self, model, n_env, device=0, precision="f32", maxcon=None, maxefc=None, tier_small=None):
        import copy

        import torch

        if isinstance(model, str):
            model = compile_mjcf(model)
        assert isinstance(model, Model)
        if maxcon is not None or maxefc is not None or tier_small is not None:
            model = copy.copy(model)  # capacities travel inside the model blob: never write them into the caller's (shared) Model
            if maxcon is not None:
                model.opt_maxcon = int(maxcon)
            if maxefc is not None:
                model.opt_maxefc = int(maxefc)
            if tier_small is not None:
                # small tier of the tail kernel: capacities (contacts, constraint rows) almost every environment stays within; the
                # rest is re-run with (maxcon, maxefc) - results are the same, shared memory per warp is 2-3x smaller
                model.opt_maxcon_small, model.opt_maxefc_small = int(tier_small[0]), int(tier_small[1])
        self.model = model
        self.n_env = int(n_env)
        self.device = int(device)
        self.torch_device = torch.device("cuda", self.device)
        self.precision = B2S_F32 if precision in ("f32", "float32") else B2S_F64
        self.dtype = torch.float32 if self.precision == B2S_F32 else torch.float64
        blob = pack_model(model)
        self._h = C.c_void_p()
        self._L = lib()
        self._check(self._L.b2s_create(blob, len(blob), self.n_env, self.device, self.precision, C.byref(self._h)))
        self._cache = {}
        global _LIVE
        if _LIVE is None:
            import weakref

            _LIVE = weakref.WeakSet()
        _LIVE.add(self)

    def _check(self, rc):
        if rc != 0:
            raise B2SError(self._L.b2s_last_error().decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.b2s_destroy(self._h)
            self._h = None

    free = close

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def array(self, name):
        """torch tensor aliasing the named device array (leading dim n_env)."""
        import torch

        if name not in self._cache:
            ptr, dt, nd = C.c_void_p(), C.c_int(), C.c_int()
            shape = (C.c_int64 * 4)()
            self._check(self._L.b2s_array(self._h, name.encode(), C.byref(ptr), C.byref(dt), C.byref(nd), shape))
            shp = [int(shape[i]) for i in range(nd.value)]
            if any(s == 0 for s in shp):
                t = torch.zeros(shp, device=self.torch_device)
            else:
                t = torch.as_tensor(_DevArray(ptr.value, shp, _TYPESTR[dt.value], self), device=self.torch_device)
            self._cache[name] = t
        return self._cache[name]

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        try:
            return self.array(name)
        except B2SError:
            raise AttributeError(name)

    def set_stream(self, stream):
        self._check(self._L.b2s_set_stream(self._h, C.c_void_p(stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream))))

    def reset(self, mask=None):
        self._check(self._L.b2s_reset(self._h, None if mask is None else C.c_void_p(mask.data_ptr())))

    def forward(self):
        self._check(self._L.b2s_forward(self._h))

    def step1(self):
        self._check(self._L.b2s_step1(self._h))

    def step2(self):
        self._check(self._L.b2s_step2(self._h))

    def step(self, n_substeps=1):
        self._check(self._L.b2s_step(self._h, int(n_substeps)))

    def jac_site(self, site_id):
        import torch

        jp = torch.empty((self.n_env, 3, self.model.nv), dtype=self.dtype, device=self.torch_device)
        jr = torch.empty_like(jp)
        self._check(self._L.b2s_jac_site(self._h, int(site_id), C.c_void_p(jp.data_ptr()), C.c_void_p(jr.data_ptr())))
        return jp, jr

    def ctrl_config(self, cfg: CtrlCfg):
        self._check(self._L.b2s_ctrl_config(self._h, C.byref(cfg)))

    def reset_envs(self, mask=None, qpos=None):
        """masked episode reset entirely on the device (b2s_reset_envs): mask uint8 [n_env] or None, qpos [n_env, nq] or None (qpos0)"""
        if qpos is not None:
            assert qpos.is_cuda and qpos.dtype == self.dtype and qpos.is_contiguous() and qpos.shape == (self.n_env, self.model.nq)
        self._check(self._L.b2s_reset_envs(self._h, None if mask is None else C.c_void_p(mask.data_ptr()),
                                           None if qpos is None else C.c_void_p(qpos.data_ptr())))

    def body_pose_override(self, body_id):
        """(pos [n_env, 3], quat [n_env, 4]) tensors that replace the constant world pose of a world-welded body per environment
        (b2s_body_pose_override; the reference writes model.body_pos / body_quat per reset, door.py:417-427)"""
        self._check(self._L.b2s_body_pose_override(self._h, int(body_id)))
        return self.array("body_xpos_ov:%d" % body_id), self.array("body_xquat_ov:%d" % body_id)

    def ctrl_reset(self, mask=None):
        self._check(self._L.b2s_ctrl_reset(self._h, None if mask is None else C.c_void_p(mask.data_ptr())))

    def env_step(self, action, n_substeps):
        assert action.is_cuda and action.dtype == self.dtype and action.is_contiguous()
        self._check(self._L.b2s_env_step(self._h, C.c_void_p(action.data_ptr()), int(n_substeps)))

    def obs_config(self, ops, a, b):
        ops, a, b = (np.ascontiguousarray(x, dtype=np.int32) for x in (ops, a, b))
        self._check(self._L.b2s_obs_config(self._h, len(ops), ops.ctypes.data, a.ctypes.data, b.ctypes.data))

    def task_config(self, body, site, left, right, obj):
        left, right, obj = (np.ascontiguousarray(x, dtype=np.int32) for x in (left, right, obj))
        self._check(self._L.b2s_task_config(self._h, int(body), int(site), left.ctypes.data, len(left), right.ctypes.data,
                                            len(right), obj.ctypes.data, len(obj)))

    def task_config2(self, body2, obj2):
        obj2 = np.ascontiguousarray(obj2, dtype=np.int32)
        self._check(self._L.b2s_task_config2(self._h, int(body2), obj2.ctypes.data, len(obj2)))

    def task_table(self, rows):
        """rows: [(op, a, b), ...] in the observation-table encoding -> array `task_vec` [n_env, len(rows)]"""
        arr = np.ascontiguousarray(rows, dtype=np.int32).reshape(-1, 3)
        op, a, b = (np.ascontiguousarray(arr[:, k]) for k in range(3))
        self._check(self._L.b2s_task_table(self._h, len(op), op.ctypes.data, a.ctypes.data, b.ctypes.data))

    def task_objects(self, geom_lists):
        """per-object grasp flags (task_out[:, 5] = bit i set when object i is grasped); at most 4 objects"""
        flat = np.ascontiguousarray([g for l in geom_lists for g in l], dtype=np.int32)
        cnt = np.ascontiguousarray([len(l) for l in geom_lists], dtype=np.int32)
        self._check(self._L.b2s_task_objects(self._h, len(geom_lists), flat.ctypes.data, cnt.ctypes.data))

    def set_export(self, flag):
        """whether b2s_env_step also writes the derived arrays (xpos, contacts, ...) of its last substep to HBM"""
        self._check(self._L.b2s_set_export(self._h, int(bool(flag))))

    def set_mode(self, mode):
        """0 = fused single kernel, 1 = pipelined phase kernels, 2 = unit queue (one persistent kernel per control step);
        identical results"""
        self._check(self._L.b2s_set_mode(self._h, int(mode)))

    def timeline(self, enable=-1):
        """(mean_us[8], count[8]) of the last timed call; enable=1/0 switches event-timed eager launches on/off"""
        m = (C.c_double * 8)()
        n = (C.c_int * 8)()
        self._check(self._L.b2s_timeline(self._h, int(enable), m, n))
        return list(m), list(n)

    def set_profile(self, flag):
        self._check(self._L.b2s_set_profile(self._h, int(bool(flag))))

    @property
    def launch_count(self):
        return int(self._L.b2s_launch_count(self._h))

    # ---- MjSim-style state I/O (binding_utils.py:1155-1184): flattened [time, qpos, qvel] per env
    def get_state(self):
        import torch

        out = torch.empty((self.n_env, 1 + self.model.nq + self.model.nv), dtype=self.dtype, device=self.torch_device)
        self._check(self._L.b2s_get_state(self._h, C.c_void_p(out.data_ptr())))
        return out

    def set_state(self, flat):
        flat = flat.to(device=self.torch_device, dtype=self.dtype).contiguous()
        assert flat.shape == (self.n_env, 1 + self.model.nq + self.model.nv)
        self._check(self._L.b2s_set_state(self._h, C.c_void_p(flat.data_ptr())))

    # ---- MjModel name tables / mj_fullM / body and geom Jacobians (binding_utils.py:362-492, 853-878; controller.py:226-229)
    def name2id(self, objtype, name):
        return int(self._L.b2s_name2id(self._h, objtype.encode(), name.encode()))

    def id2name(self, objtype, idx):
        r = self._L.b2s_id2name(self._h, objtype.encode(), int(idx))
        return None if r is None else r.decode()

    def full_m(self):
        import torch

        out = torch.empty((self.n_env, self.model.nv, self.model.nv), dtype=self.dtype, device=self.torch_device)
        self._check(self._L.b2s_full_m(self._h, C.c_void_p(out.data_ptr())))
        return out

    def _jac(self, fn, idx):
        import torch

        jp = torch.empty((self.n_env, 3, self.model.nv), dtype=self.dtype, device=self.torch_device)
        jr = torch.empty_like(jp)
        self._check(fn(self._h, int(idx), C.c_void_p(jp.data_ptr()), C.c_void_p(jr.data_ptr())))
        return jp, jr

    def jac_body(self, body_id):
        return self._jac(self._L.b2s_jac_body, body_id)

    def jac_geom(self, geom_id):
        return self._jac(self._L.b2s_jac_geom, geom_id)
