// Fused controller (runs between step1 and step2 of every substep inside the step kernel).
#pragma once
#include "b2s_solver.cuh"

template <typename R> struct CtrlState { R goal_pos[3]; R goal_ori[9]; };

template <typename R> DEV void ctrl_load(Eng<R>& e, const DState<R>& s, const CtrlCfgDev& cc, CtrlState<R>& cs, int env) {}
template <typename R> DEV void ctrl_run(Eng<R>& e, const DState<R>& s, const CtrlCfgDev& cc, CtrlState<R>& cs, int env, bool policy_step) {}
template <typename R> DEV void ctrl_store(Eng<R>& e, const DState<R>& s, const CtrlCfgDev& cc, CtrlState<R>& cs, int env) {}
template <typename R>
__global__ void ctrl_reset_kernel(const __grid_constant__ DModel<R> m, const __grid_constant__ DState<R> s,
                                  const __grid_constant__ CtrlCfgDev cc, const uint8_t* mask) {}
