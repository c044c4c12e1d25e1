/* TEST INFRASTRUCTURE - CPU oracle, pipeline order of the engine calls at
 * `robosuite/utils/binding_utils.py:1093-1107` (forward / step / step1 / step2).  Stage order per SURVEY.md
 * Appendix C: step1 = position stage (kinematics, CRB, factor, collision, constraint rows) + velocity stage
 * (body velocities, passive, bias); the controller runs between step1 and step2 (`environments/base.py:496-501`);
 * step2 = actuation, smooth acceleration, constraint solve, Euler integration. */
#include "b2s_oracle.h"

void o_step1(const OModel* m, OData* d) {
  o_kinematics(m, d);
  o_crb(m, d);
  o_factor_m(m, d);
  o_collision(m, d);
  o_com_vel(m, d);
  o_make_constraint(m, d);
  o_passive(m, d);
  o_rne_bias(m, d);
}

static void fwd_rest(const OModel* m, OData* d) {
  o_fwd_actuation(m, d);
  o_fwd_acceleration(m, d);
  o_fwd_constraint(m, d);
}

void o_step2(const OModel* m, OData* d) {
  fwd_rest(m, d);
  o_euler(m, d);
}

void o_forward(const OModel* m, OData* d) {
  o_step1(m, d);
  fwd_rest(m, d);
}

void o_step(const OModel* m, OData* d) {
  o_step1(m, d);
  o_step2(m, d);
}
