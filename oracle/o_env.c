/* TEST INFRASTRUCTURE - CPU oracle, small accessor API for the Python test harness (ctypes). */
#include "b2s_oracle.h"
#include <string.h>

int o_get_int(const OData* d, const char* name) {
  if (!strcmp(name, "ncon")) return d->ncon;
  if (!strcmp(name, "nefc")) return d->nefc;
  if (!strcmp(name, "nf")) return d->nf;
  if (!strcmp(name, "nl")) return d->nl;
  if (!strcmp(name, "solver_niter")) return d->solver_niter;
  if (!strcmp(name, "warn_flags")) return d->warn_flags;
  return -1;
}
int o_model_int(const OModel* m, const char* name) {
  if (!strcmp(name, "nq")) return m->nq;
  if (!strcmp(name, "nv")) return m->nv;
  if (!strcmp(name, "nu")) return m->nu;
  if (!strcmp(name, "nbody")) return m->nbody;
  if (!strcmp(name, "njnt")) return m->njnt;
  if (!strcmp(name, "ngeom")) return m->ngeom;
  if (!strcmp(name, "nsite")) return m->nsite;
  if (!strcmp(name, "nM")) return m->nM;
  if (!strcmp(name, "nmocap")) return m->nmocap;
  if (!strcmp(name, "nsensordata")) return m->nsensordata;
  return -1;
}
void o_get_contact(const OData* d, int i, double* buf, int* ibuf) {
  const OContact* c = d->contact + i;
  buf[0] = c->dist;
  memcpy(buf + 1, c->pos, 3 * sizeof(double));
  memcpy(buf + 4, c->frame, 9 * sizeof(double));
  memcpy(buf + 13, c->friction, 5 * sizeof(double));
  memcpy(buf + 18, c->solref, 2 * sizeof(double));
  memcpy(buf + 20, c->solimp, 5 * sizeof(double));
  buf[25] = c->mu;
  ibuf[0] = c->dim; ibuf[1] = c->geom1; ibuf[2] = c->geom2; ibuf[3] = c->efc_address;
}
const int* o_efc_int(const OData* d, const char* name) {
  if (!strcmp(name, "type")) return d->efc_type;
  if (!strcmp(name, "id")) return d->efc_id;
  if (!strcmp(name, "state")) return d->efc_state;
  return 0;
}
void o_set_time(OData* d, double t) { d->time = t; }

/* what the reference does per reset for the Door: sim.model.body_pos[id] = pos; body_quat[id] = quat (door.py:417-427).  The model owns a
 * private copy of its blob, so the constants are writable. */
void o_model_set_body_pose(OModel* m, int body, const double* pos, const double* quat) {
  double* bp = (double*)m->body_pos + 3 * body; double* bq = (double*)m->body_quat + 4 * body;
  for (int k = 0; k < 3; k++) bp[k] = pos[k];
  for (int k = 0; k < 4; k++) bq[k] = quat[k];
}
