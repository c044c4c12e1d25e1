"""TEST INFRASTRUCTURE - ctypes binding of the CPU oracle (oracle/libb2s_oracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libb2s_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.o_model_load.restype = C.c_void_p
        L.o_model_load.argtypes = [C.c_char_p, C.c_size_t]
        L.o_model_free.argtypes = [C.c_void_p]
        L.o_model_set_body_pose.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.o_data_new.restype = C.c_void_p
        L.o_data_new.argtypes = [C.c_void_p]
        L.o_data_free.argtypes = [C.c_void_p]
        L.o_data_field.restype = C.POINTER(C.c_double)
        L.o_data_field.argtypes = [C.c_void_p, C.c_char_p]
        L.o_get_int.argtypes = [C.c_void_p, C.c_char_p]
        L.o_model_int.argtypes = [C.c_void_p, C.c_char_p]
        L.o_efc_int.restype = C.POINTER(C.c_int)
        L.o_efc_int.argtypes = [C.c_void_p, C.c_char_p]
        L.o_get_contact.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        for fn in ("o_reset_data", "o_forward", "o_step1", "o_step2", "o_step", "o_kinematics", "o_crb", "o_factor_m",
                   "o_collision", "o_make_constraint", "o_com_vel", "o_passive", "o_rne_bias", "o_fwd_actuation",
                   "o_fwd_acceleration", "o_fwd_constraint", "o_euler"):
            getattr(L, fn).argtypes = [C.c_void_p, C.c_void_p]
            getattr(L, fn).restype = None
        L.o_jac.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double),
                            C.POINTER(C.c_double), C.c_int]
        L.o_full_m.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.o_collide_pair.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.o_set_time.argtypes = [C.c_void_p, C.c_double]
        L.o_ctrl_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.o_ctrl_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.o_env_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_int]
        for fn in ("o_ctrl_reset", "o_ctrl_run", "o_env_step"):
            getattr(L, fn).restype = None
        _LIB = L
    return _LIB


class CtrlCfg(C.Structure):
    """Mirror of OCtrlCfg (oracle/o_ctrl.c); same field order as b2s_ctrl_cfg in include/b2s.h."""
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


class CtrlState(C.Structure):
    _fields_ = [("goal_pos", C.c_double * 3), ("goal_ori", C.c_double * 9), ("initial_joint", C.c_double * 8),
                ("grip_action", C.c_double * 4), ("torques", C.c_double * 8), ("jv_goal", C.c_double * 8),
                ("jv_last_err", C.c_double * 8), ("jv_summed", C.c_double * 8), ("jv_derr", (C.c_double * 8) * 5),
                ("jv_ptr", C.c_int), ("jv_size", C.c_int), ("jv_saturated", C.c_int)]


_SHAPES = {
    "qpos": ("nq",), "qvel": ("nv",), "qacc": ("nv",), "qacc_warmstart": ("nv",), "ctrl": ("nu",),
    "qfrc_applied": ("nv",), "mocap_pos": ("nmocap", 3), "mocap_quat": ("nmocap", 4), "xpos": ("nbody", 3),
    "xquat": ("nbody", 4), "xmat": ("nbody", 9), "xipos": ("nbody", 3), "ximat": ("nbody", 9), "xanchor": ("njnt", 3),
    "xaxis": ("njnt", 3), "geom_xpos": ("ngeom", 3), "geom_xmat": ("ngeom", 9), "site_xpos": ("nsite", 3),
    "site_xmat": ("nsite", 9), "cdof": ("nv", 6), "cinert": ("nbody", 10), "qM": ("nM",),
    "M": ("nv", "nv"), "L": ("nv", "nv"), "cvel": ("nbody", 6), "cdof_dot": ("nv", 6), "qfrc_bias": ("nv",),
    "qfrc_passive": ("nv",), "qfrc_actuator": ("nv",), "actuator_force": ("nu",), "qfrc_smooth": ("nv",),
    "qacc_smooth": ("nv",), "qfrc_constraint": ("nv",), "sensordata": ("nsensordata",),
}
_MAXEFC = 512


class Oracle:
    """One model + one data instance of the oracle; numpy views alias the C memory."""

    def __init__(self, blob: bytes):
        L = lib()
        self._L = L
        self._blob = blob
        self.m = L.o_model_load(blob, len(blob))
        if not self.m:
            raise ValueError("bad model blob")
        self.d = L.o_data_new(self.m)
        self.n = {k: L.o_model_int(self.m, k.encode()) for k in
                  ("nq", "nv", "nu", "nbody", "njnt", "ngeom", "nsite", "nM", "nmocap", "nsensordata")}
        for name, shp in _SHAPES.items():
            shape = tuple(self.n[s] if isinstance(s, str) else s for s in shp)
            size = int(np.prod(shape))
            if size == 0:
                setattr(self, name, np.zeros(shape))
                continue
            p = L.o_data_field(self.d, name.encode())
            setattr(self, name, np.ctypeslib.as_array(p, shape=(size,)).reshape(shape))
        nv = self.n["nv"]
        self._efcJ = np.ctypeslib.as_array(L.o_data_field(self.d, b"efc_J"), shape=(_MAXEFC * max(nv, 1),))
        self._time = L.o_data_field(self.d, b"time")

    def __del__(self):
        try:
            self._L.o_data_free(self.d)
            self._L.o_model_free(self.m)
        except Exception:
            pass

    def set_body_pose(self, body, pos, quat):
        """model.body_pos[body] = pos; model.body_quat[body] = quat (what the reference does per reset for the Door)"""
        p = (C.c_double * 3)(*[float(x) for x in pos]); q = (C.c_double * 4)(*[float(x) for x in quat])
        self._L.o_model_set_body_pose(self.m, int(body), p, q)

    # scalars
    @property
    def time(self):
        return self._time[0]

    @time.setter
    def time(self, v):
        self._time[0] = v

    def geti(self, name):
        return self._L.o_get_int(self.d, name.encode())

    ncon = property(lambda s: s.geti("ncon"))
    nefc = property(lambda s: s.geti("nefc"))

    def efc(self, name):
        n = self.nefc
        if name == "J":
            nv = self.n["nv"]
            return self._efcJ[: n * nv].reshape(n, nv)
        if name in ("type", "id", "state"):
            return np.ctypeslib.as_array(self._L.o_efc_int(self.d, name.encode()), shape=(_MAXEFC,))[:n]
        return np.ctypeslib.as_array(self._L.o_data_field(self.d, ("efc_" + name).encode()), shape=(_MAXEFC,))[:n]

    def contacts(self):
        out = []
        buf = (C.c_double * 26)()
        ib = (C.c_int * 4)()
        for i in range(self.ncon):
            self._L.o_get_contact(self.d, i, buf, ib)
            b = np.array(buf)
            out.append(dict(dist=b[0], pos=b[1:4], frame=b[4:13].reshape(3, 3), friction=b[13:18], solref=b[18:20],
                            solimp=b[20:25], mu=b[25], dim=ib[0], geom1=ib[1], geom2=ib[2], efc_address=ib[3]))
        return out

    def jac(self, point, body):
        nv = self.n["nv"]
        jp = np.zeros((3, nv))
        jr = np.zeros((3, nv))
        pt = np.ascontiguousarray(point, dtype=np.float64)
        self._L.o_jac(self.m, self.d, jp.ctypes.data_as(C.POINTER(C.c_double)), jr.ctypes.data_as(C.POINTER(C.c_double)),
                      pt.ctypes.data_as(C.POINTER(C.c_double)), int(body))
        return jp, jr

    # ---- controller half (oracle/o_ctrl.c)
    def ctrl_setup(self, cfg):
        self.ctrl_cfg = cfg
        self.ctrl_state = CtrlState()

    def ctrl_reset(self):
        self._L.o_ctrl_reset(self.m, self.d, C.byref(self.ctrl_cfg), C.byref(self.ctrl_state))

    def ctrl_run(self, action=None):
        a = None
        if action is not None:
            arr = np.ascontiguousarray(action, dtype=np.float64)
            a = arr.ctypes.data_as(C.POINTER(C.c_double))
        self._L.o_ctrl_run(self.m, self.d, C.byref(self.ctrl_cfg), C.byref(self.ctrl_state), a)

    def env_step(self, action, nsub=25):
        arr = np.ascontiguousarray(action, dtype=np.float64)
        self._L.o_env_step(self.m, self.d, C.byref(self.ctrl_cfg), C.byref(self.ctrl_state),
                           arr.ctypes.data_as(C.POINTER(C.c_double)), int(nsub))

    def __getattr__(self, name):
        if name in ("reset_data", "forward", "step1", "step2", "step", "kinematics", "crb", "factor_m", "collision",
                    "make_constraint", "com_vel", "passive", "rne_bias", "fwd_actuation", "fwd_acceleration",
                    "fwd_constraint", "euler"):
            fn = getattr(self._L, "o_" + name)
            return lambda: fn(self.m, self.d)
        raise AttributeError(name)
