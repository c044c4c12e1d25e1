/* libb2s - C ABI of the B200-native batched rigid-body engine (one environment per warp, sm_100a).
 *
 * Every entry point replaces one call the reference makes into its third-party engine through
 * `robosuite/utils/binding_utils.py` (the `MjSim` shim, SURVEY.md section 8b "Seam 1"), batched over n_env
 * independent environments.  All array arguments are DEVICE pointers unless the name ends in `_host`.
 * Return value: 0 on success, negative error code otherwise; `b2s_last_error()` returns a thread-local message
 * (the Python layer maps codes to the exception types of `robosuite/utils/errors.py`).
 * A handle is not thread-safe (matches the reference: one MjSim per env per process, binding_utils.py:1059).
 */
#ifndef B2S_H
#define B2S_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2s_sim b2s_sim;

enum { B2S_OK = 0, B2S_ERR_ARG = -1, B2S_ERR_CUDA = -2, B2S_ERR_MODEL = -3, B2S_ERR_UNSUPPORTED = -4 };
enum { B2S_F32 = 0, B2S_F64 = 1, B2S_I32 = 2, B2S_I64 = 3 };
/* controller kinds for the fused control step (controller_config "type", controllers/parts/controller_factory.py:145) */
/* arm part controllers: controllers/parts/arm/osc.py, parts/generic/joint_vel.py, joint_pos.py, joint_tor.py.  The two joint-space
 * kinds 3 / 4 keep their per-joint scaling in jv_in/out_*, their gains in jv_kp / jv_kd (kd = 2 sqrt(kp) damping_ratio). */
enum { B2S_CTRL_NONE = 0, B2S_CTRL_OSC_POSE = 1, B2S_CTRL_JOINT_VELOCITY = 2, B2S_CTRL_JOINT_POSITION = 3, B2S_CTRL_JOINT_TORQUE = 4,
       B2S_CTRL_OSC_POSITION = 5 /* controllers/parts/arm/osc.py:152-166: 3-dim arm action, orientation held */ };

/* MjSim.from_xml_string (binding_utils.py:1074-1087): `model_blob` is the flat compiled model produced by
 * robosuite_b200.mjcf.compiler.pack_model (host memory).  precision: B2S_F32 (production) or B2S_F64 (debug). */
int b2s_create(const void* model_blob_host, size_t nbytes, int n_env, int device, int precision, b2s_sim** out);
/* MjSim.free (binding_utils.py:1186-1192) */
void b2s_destroy(b2s_sim* sim);
const char* b2s_last_error(void);
/* all device work of this handle is enqueued on `cuda_stream` (a cudaStream_t); default: the legacy default stream */
int b2s_set_stream(b2s_sim* sim, void* cuda_stream);

/* MjSim.reset -> mj_resetData (binding_utils.py:1089-1091); env_mask: n_env bytes on device (non-zero = reset) or NULL */
int b2s_reset(b2s_sim* sim, const uint8_t* env_mask);
/* MjSim.forward -> mj_forward (binding_utils.py:1093-1095) */
int b2s_forward(b2s_sim* sim);
/* MjSim.step1 / step2 -> mj_step1 / mj_step2 (binding_utils.py:1101-1107); ctrl is read between them */
int b2s_step1(b2s_sim* sim);
int b2s_step2(b2s_sim* sim);
/* MjSim.step -> mj_step (binding_utils.py:1097-1099), repeated n_substeps times inside one kernel with ctrl held */
int b2s_step(b2s_sim* sim, int n_substeps);

/* Named device arrays, leading dimension n_env: qpos qvel qacc qacc_warmstart ctrl time xpos xquat xmat
 * site_xpos site_xmat geom_xpos geom_xmat qM(dense nv x nv) qfrc_bias qfrc_passive qfrc_actuator qfrc_constraint
 * actuator_force ncon contact_geom contact_dist contact_pos contact_frame nefc efc_force warn ...
 * (the attributes the reference touches).  dtype is B2S_F32/F64/I32. */
int b2s_array(b2s_sim* sim, const char* name, void** dev_ptr, int* dtype, int* ndim, int64_t shape[4]);

/* MjSim.get_state().flatten() / set_state_from_flattened (binding_utils.py:1155-1184, MjSimState.flatten :56-70): device buffers
 * [n_env, 1 + nq + nv] holding time, qpos, qvel per environment.  b2s_set_state does what the reference does after it: nothing else
 * (call b2s_forward next, as callers of set_state_from_flattened do). */
int b2s_get_state(b2s_sim* sim, void* out_dev);
int b2s_set_state(b2s_sim* sim, const void* in_dev);
/* MjModel.{body,joint,geom,site,actuator,...}_name2id / id2name (binding_utils.py:362-492).  type: "body" "joint" "geom" "site"
 * "actuator" "mesh" "camera" "light".  name2id returns the id or -1 (unknown name / type); id2name returns a pointer that stays valid
 * for the life of the handle, NULL when out of range ("" for unnamed objects). */
int b2s_name2id(const b2s_sim* sim, const char* type, const char* name);
const char* b2s_id2name(const b2s_sim* sim, const char* type, int id);
/* mj_fullM (controllers/parts/controller.py:226-229 builds the dense mass matrix from qM): out_dev [n_env, nv, nv], valid after
 * b2s_forward / b2s_step1 */
int b2s_full_m(b2s_sim* sim, void* out_dev);
/* MjData.get_body_jacp/jacr, get_geom_jacp/jacr (binding_utils.py:853-878 and the geom variants): Jacobian of the body frame origin /
 * geom centre, [n_env, 3, nv] device buffers (either may be NULL), valid after b2s_forward / b2s_step1 */
int b2s_jac_body(b2s_sim* sim, int body_id, void* jacp, void* jacr);
int b2s_jac_geom(b2s_sim* sim, int geom_id, void* jacp, void* jacr);

/* MjData.get_site_jacp/jacr (binding_utils.py:826-852): jacp/jacr are [n_env,3,nv] device buffers (either may be NULL);
 * valid after b2s_forward/b2s_step1. */
int b2s_jac_site(b2s_sim* sim, int site_id, void* jacp, void* jacr);

/* Fused control step = MujocoEnv.step's substep loop (environments/base.py:494-505):
 * n_substeps x { step1 ; controller(action) -> ctrl ; step2 } in ONE kernel with state resident on chip.
 * Configure once with b2s_ctrl_config (fields of controllers/config/default/parts/osc_pose.json + robot indices),
 * then call b2s_env_step(action[n_env, action_dim]) per control step.  obs_out[n_env, obs_dim] may be NULL. */
typedef struct {
  int kind;            /* B2S_CTRL_* */
  int action_dim;      /* arm dims + gripper dims */
  int n_arm;           /* number of arm joints (7) */
  int arm_dof[8];      /* dof (= qvel) index of each arm joint */
  int arm_qpos[8];
  int arm_act[8];      /* actuator index of each arm joint */
  int eef_site;        /* ref_name site id */
  int base_site;       /* "{prefix}{part}_center" site id (controller origin) */
  int n_grip;          /* gripper actuators (2) */
  int grip_act[4];
  double grip_sign[4]; /* format_action signs (models/grippers/panda_gripper.py:55-57) */
  double grip_speed;   /* 0.2 per policy step */
  double kp[6], damping_ratio[6];
  double input_max[6], input_min[6], output_max[6], output_min[6];
  double null_kp;      /* nullspace_torques joint_kp (control_utils.py:7-40), 10 */
  int uncouple_pos_ori;
  int n_obs_site;      /* sites appended to obs (task layer) - reserved */
  /* JOINT_VELOCITY (controllers/parts/generic/joint_vel.py:60-209): per-joint PID gains and action scaling */
  double jv_kp[8], jv_ki[8], jv_kd[8], jv_in_max[8], jv_in_min[8], jv_out_max[8], jv_out_min[8];
  double jv_vel_lo, jv_vel_hi;
  int jv_use_vel_limits, jv_torque_comp;
} b2s_ctrl_cfg;
int b2s_ctrl_config(b2s_sim* sim, const b2s_ctrl_cfg* cfg);
/* controller.reset_goal + update_initial_joints (osc.py:520-544) for masked envs (NULL = all); needs a prior forward */
int b2s_ctrl_reset(b2s_sim* sim, const uint8_t* env_mask);
int b2s_env_step(b2s_sim* sim, const void* action, int n_substeps);
/* MujocoEnv.reset for a device-resident subset (environments/base.py:277-347: _reset_internal -> sim.forward -> controller reset -> observation
 * cache emptied): masked envs (device bytes, NULL = all) take qpos from `qpos_new` ([n_env, nq] device array of the handle's precision holding a
 * sampled initial state for every environment; NULL = qpos0), velocities / accelerations / warm start / ctrl / time / warn cleared, forward pass,
 * controller goals rebuilt, GJK warm starts dropped.  Every launch is asynchronous on the handle's stream: no host round trip, so a horizon-based
 * auto-reset can be enqueued after every step. */
int b2s_reset_envs(b2s_sim* sim, const uint8_t* env_mask, const void* qpos_new);

/* Per-environment world pose of a body welded to the world.  The reference writes sampled placements into the MODEL per reset
 * (Door: `sim.model.body_pos[door_body_id] = door_pos; body_quat = door_quat`, environments/manipulation/door.py:417-427); a batch shares
 * its model constants, so such poses are per-environment DATA here.  After the call the arrays "body_xpos_ov:<id>" [n_env, 3] and
 * "body_xquat_ov:<id>" [n_env, 4] (initialised to the model's pose; fetch them with b2s_array and write them like any state array)
 * replace the constant world pose of that body in every kinematics pass.  Bodies welded to an overridden body keep THEIR constant world
 * pose: override them too (pose = parent pose * local pose).  At most 4 bodies per handle. */
int b2s_body_pose_override(b2s_sim* sim, int body_id);

/* Observation program = MujocoEnv._get_observations flattened (environments/base.py:429-465): one (op, a, b) entry
 * per output scalar (ops: enum OB_* in csrc/b2s_types.cuh; OB_REL_*_LAG entries read the previous sample, as the reference's
 * sensor ordering does: manipulation_env.py:268-329).  Creates the device arrays "obs" [n_env, obs_dim] and "obs_fresh" [n_env]
 * (1 = observation cache empty; set it when an environment is reset).  "obs" is written by b2s_env_step after the LAST
 * substep - reset()'s forced update advances the observables' period timer by one model step, so every later sample falls on
 * the last substep of a control step (utils/observables.py:214-259, environments/base.py:418-427) - and by b2s_forward for
 * environments whose obs_fresh flag is set. obs_dim <= 128. */
int b2s_obs_config(b2s_sim* sim, int obs_dim, const int* op_host, const int* a_host, const int* b_host);
/* Task outputs "task_out" [n_env,8] = (target body height, |site - body|, grasp flag, horizontal |body - body2|,
 * obj-obj2 contact flag, 0, 0, 0) from the poses/contacts of the
 * last step1 (what the reference's reward()/_check_grasp read: manipulation/lift.py:224-273, manipulation_env.py:331-376).
 * Geom id lists: left / right finger(pad) groups and object geoms. */
int b2s_task_config(b2s_sim* sim, int body, int site, const int* left, int nleft, const int* right, int nright,
                    const int* obj, int nobj);
/* optional second object (Stack's cubeB: staged_rewards, manipulation/stack.py:266-312); call after b2s_task_config */
int b2s_task_config2(b2s_sim* sim, int body2, const int* obj2, int nobj2);
/* optional per-object grasp flags for multi-object tasks (NutAssembly / PickPlace staged_rewards check the grasp against
 * the geoms of the still-active objects: manipulation/nut_assembly.py:318-327, pick_place.py:352-361): up to 4 objects,
 * geoms = concatenated geom id lists, counts[i] = length of list i.  task_out[5] = sum_i 2^i * grasped_i. */
int b2s_task_objects(b2s_sim* sim, int nobjects, const int* geoms, const int* counts);
/* task table: n (<= 64) scalars in the observation-table encoding, evaluated after the LAST substep of b2s_env_step into
 * the array "task_vec" [n_env, n] - the poses reward()/_check_success() read from sim.data after the step
 * (manipulation/door.py:219-266, nut_assembly.py:247-400, pick_place.py:275-425). Requires b2s_obs_config first. */
int b2s_task_table(b2s_sim* sim, int n, const int* op, const int* a, const int* b);

/* b2s_env_step also exports the derived arrays of its last substep (xpos, contacts, efc ...) when flag != 0 (default 1);
 * the throughput path switches it off so that per-step HBM traffic is state + action + obs only */
int b2s_set_export(b2s_sim* sim, int flag);

/* Scheduling of b2s_env_step / b2s_step: 0 = fused (one kernel per call, state resident in shared memory for all
 * substeps), 1 = pipeline (five phase kernels per substep exchanging a workspace row through L2).  Results are
 * identical; see DESIGN.md section 5 for when each wins. */
int b2s_set_mode(b2s_sim* sim, int mode);

/* debugging aid: b2s_env_step accumulates per-phase clock cycles per environment into the array "prof" [n_env,12]
 * (0 kinematics, 1 velocity+crb, 2 collision, 3 constraint rows, 4 controller, 5 actuation+smooth acc, 6 solver,
 * 7 integrate, 11 time spent waiting at block barriers) and collision candidate counts into "dbg" [n_env,4] */
int b2s_set_profile(b2s_sim* sim, int flag);

/* number of kernels this handle has launched since creation (bench.py "gpu_launches") */
int64_t b2s_launch_count(const b2s_sim* sim);
/* Measurement aid (pipeline mode): enable = 1 switches b2s_env_step / b2s_step to eager launches bracketed by timing events,
 * enable = 0 back to CUDA-graph replay, enable < 0 only reads.  mean_us / count (may be NULL) receive, for the LAST call made
 * while enabled, the mean event-to-event interval per launch kind: [1] counter memset, [2] phase 0, [3] analytic narrow phase,
 * [4] convex narrow phase, [5] merged tail phase (or phase 2), [6] phase 3, [7] phase 4. */
int b2s_timeline(b2s_sim* sim, int enable, double mean_us[8], int count[8]);

/* Host-only diagnostic (no device needed): words per warp of the shared-memory layouts and of the global workspace row that a model
 * of these dimensions gets - out_words[5] = fused kernel, phase 0, tail small tier, tail large tier, row.  out_layouts / out_pio may
 * be NULL; see csrc/b2s_capi.cu for their packing (tests/test_cpu_layouts.py checks the layout invariants with them). */
int b2s_debug_layouts(int nq, int nv, int nu, int nbody, int ncg, int nsite, int hc_stride, int maxcon, int maxefc, int mc_small,
                      int me_small, int osc_in_tail, int* out_words, int* out_layouts, int* out_pio);

#ifdef __cplusplus
}
#endif
#endif
