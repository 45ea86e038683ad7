/* ORACLE (test infrastructure, NOT product code).  See engine.h for the status banner:
 * physics parity is UNPINNED (MuJoCo absent), controller / state machine / reward are pinned by
 * tests/golden (reference python run on this engine).
 *
 * Every function names the upstream routine it restates.  Reference call sites:
 *   physics.step()    simulation/environment/robot_env.py:100,119,142,157
 *   physics.reset()   simulation/environment/robot_env.py:62
 *   mj_jacBody        simulation/controller/actuator.py:89
 *   data.contact      simulation/controller/actuator.py:157-176
 */
#include "engine.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define MINVAL 1e-15

/* ------------------------------------------------------------------ small math */
static inline double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(double *r, const double *a, const double *b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void copy3(double *r, const double *a) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
static inline void sub3(double *r, const double *a, const double *b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
static inline void add3(double *r, const double *a, const double *b) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
static inline void addscl3(double *r, const double *a, const double *b, double s) { r[0] = a[0] + s * b[0]; r[1] = a[1] + s * b[1]; r[2] = a[2] + s * b[2]; }
static inline void scl3(double *r, const double *a, double s) { r[0] = a[0] * s; r[1] = a[1] * s; r[2] = a[2] * s; }
static inline double norm3(const double *a) { return sqrt(dot3(a, a)); }
static inline double normalize3(double *a) {
  double n = norm3(a);
  if (n < MINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; return n; }
  a[0] /= n; a[1] /= n; a[2] /= n;
  return n;
}
static void quat_mul(double *r, const double *a, const double *b) {
  double t[4] = {a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                 a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]};
  memcpy(r, t, sizeof t);
}
static void quat_normalize(double *q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  for (int i = 0; i < 4; i++) q[i] /= n;
}
static void quat2mat(double *R, const double *q) { /* row-major 3x3 */
  double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = w * w + x * x - y * y - z * z; R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = w * w - x * x + y * y - z * z; R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = w * w - x * x - y * y + z * z;
}
static void rot_vec_quat(double *r, const double *v, const double *q) {
  double R[9];
  quat2mat(R, q);
  double t[3] = {R[0] * v[0] + R[1] * v[1] + R[2] * v[2], R[3] * v[0] + R[4] * v[1] + R[5] * v[2], R[6] * v[0] + R[7] * v[1] + R[8] * v[2]};
  copy3(r, t);
}
static void mulmat3vec(double *r, const double *R, const double *v) {
  double t[3] = {R[0] * v[0] + R[1] * v[1] + R[2] * v[2], R[3] * v[0] + R[4] * v[1] + R[5] * v[2], R[6] * v[0] + R[7] * v[1] + R[8] * v[2]};
  copy3(r, t);
}
static void mulmat3Tvec(double *r, const double *R, const double *v) {
  double t[3] = {R[0] * v[0] + R[3] * v[1] + R[6] * v[2], R[1] * v[0] + R[4] * v[1] + R[7] * v[2], R[2] * v[0] + R[5] * v[1] + R[8] * v[2]};
  copy3(r, t);
}
static void axisangle2quat(double *q, const double *axis, double angle) {
  double s = sin(angle * 0.5);
  q[0] = cos(angle * 0.5); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}

/* dense Cholesky A = L L^T (lower, in place, stride n); returns 0 on success */
static int chol_factor(double *A, int n) {
  for (int j = 0; j < n; j++) {
    double s = A[j * n + j];
    for (int k = 0; k < j; k++) s -= A[j * n + k] * A[j * n + k];
    if (s < MINVAL) s = MINVAL;
    s = sqrt(s);
    A[j * n + j] = s;
    for (int i = j + 1; i < n; i++) {
      double t = A[i * n + j];
      for (int k = 0; k < j; k++) t -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = t / s;
    }
  }
  return 0;
}
static void chol_solve(const double *L, int n, double *x) {
  for (int i = 0; i < n; i++) {
    double s = x[i];
    for (int k = 0; k < i; k++) s -= L[i * n + k] * x[k];
    x[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = x[i];
    for (int k = i + 1; k < n; k++) s -= L[k * n + i] * x[k];
    x[i] = s / L[i * n + i];
  }
}

/* ------------------------------------------------------------------ model construction */
OModel *orc_model_new(void) {
  OModel *m = (OModel *)calloc(1, sizeof(OModel));
  m->mpr_tolerance = 1e-6;
  m->mpr_iterations = 50;
  return m;
}
void orc_model_free(OModel *m) {
  if (!m) return;
  for (int i = 0; i < O_MAXM; i++) { free(m->mesh[i].vert); free(m->mesh[i].adjadr); free(m->mesh[i].adj); }
  free(m);
}
#define FIELD_D(nm, cap) if (!strcmp(name, #nm)) { if (n > (cap)) return -2; memcpy(m->nm, v, n * sizeof(double)); return 0; }
#define SCALAR_D(nm) if (!strcmp(name, #nm)) { m->nm = v[0]; return 0; }
int orc_set_d(OModel *m, const char *name, const double *v, int n) {
  SCALAR_D(timestep) SCALAR_D(impratio) SCALAR_D(tolerance) SCALAR_D(mpr_tolerance)
  FIELD_D(gravity, 3) FIELD_D(body_pos, O_MAXB * 3) FIELD_D(body_quat, O_MAXB * 4) FIELD_D(body_ipos, O_MAXB * 3)
  FIELD_D(body_iquat, O_MAXB * 4) FIELD_D(body_mass, O_MAXB) FIELD_D(body_inertia, O_MAXB * 3)
  FIELD_D(jnt_pos, O_MAXJ * 3) FIELD_D(jnt_axis, O_MAXJ * 3) FIELD_D(jnt_range, O_MAXJ * 2) FIELD_D(jnt_solref, 2)
  FIELD_D(jnt_solimp, 5) FIELD_D(qpos0, O_MAXQ) FIELD_D(dof_armature, O_MAXV) FIELD_D(dof_damping, O_MAXV)
  FIELD_D(geom_pos, O_MAXG * 3) FIELD_D(geom_quat, O_MAXG * 4) FIELD_D(geom_friction, O_MAXG * 3)
  FIELD_D(geom_margin, O_MAXG) FIELD_D(geom_gap, O_MAXG) FIELD_D(geom_solref, O_MAXG * 2) FIELD_D(geom_solimp, O_MAXG * 5)
  FIELD_D(geom_rbound, O_MAXG) FIELD_D(geom_size, O_MAXG * 3) FIELD_D(act_gear, O_MAXU) FIELD_D(act_ctrlrange, O_MAXU * 2)
  FIELD_D(body_invweight0, O_MAXB * 2) FIELD_D(dof_invweight0, O_MAXV) SCALAR_D(meaninertia)
  return -1;
}
/* read back what orc_model_finalize derived (tests/test_model_compiler.py compares it with the product compiler) */
int orc_get_d(const OModel *m, const char *name, double *out, int n) {
  const double *src = NULL; int cap = 0;
  if (!strcmp(name, "body_invweight0")) { src = m->body_invweight0; cap = 2 * m->nbody; }
  else if (!strcmp(name, "dof_invweight0")) { src = m->dof_invweight0; cap = m->nv; }
  else if (!strcmp(name, "meaninertia")) { src = &m->meaninertia; cap = 1; }
  if (!src || n < cap) return -1;
  memcpy(out, src, cap * sizeof(double));
  return cap;
}
#define FIELD_I(nm, cap) if (!strcmp(name, #nm)) { if (n > (cap)) return -2; memcpy(m->nm, v, n * sizeof(int)); return 0; }
#define SCALAR_I(nm) if (!strcmp(name, #nm)) { m->nm = v[0]; return 0; }
int orc_set_i(OModel *m, const char *name, const int *v, int n) {
  SCALAR_I(nbody) SCALAR_I(njnt) SCALAR_I(nq) SCALAR_I(nv) SCALAR_I(nu) SCALAR_I(ngeom) SCALAR_I(nmesh) SCALAR_I(npair)
  SCALAR_I(iterations) SCALAR_I(cone_elliptic) SCALAR_I(mpr_iterations)
  FIELD_I(body_parentid, O_MAXB) FIELD_I(body_weldid, O_MAXB) FIELD_I(body_jntadr, O_MAXB) FIELD_I(body_jntnum, O_MAXB)
  FIELD_I(body_dofadr, O_MAXB) FIELD_I(body_dofnum, O_MAXB) FIELD_I(jnt_type, O_MAXJ) FIELD_I(jnt_bodyid, O_MAXJ)
  FIELD_I(jnt_qposadr, O_MAXJ) FIELD_I(jnt_dofadr, O_MAXJ) FIELD_I(jnt_limited, O_MAXJ) FIELD_I(dof_bodyid, O_MAXV)
  FIELD_I(dof_jntid, O_MAXV) FIELD_I(dof_parentid, O_MAXV) FIELD_I(geom_type, O_MAXG) FIELD_I(geom_bodyid, O_MAXG)
  FIELD_I(geom_meshid, O_MAXG) FIELD_I(geom_condim, O_MAXG) FIELD_I(act_dofid, O_MAXU) FIELD_I(pair_geom1, O_MAXPAIR)
  FIELD_I(pair_geom2, O_MAXPAIR)
  return -1;
}
int orc_set_mesh(OModel *m, int id, int nvert, const double *vert, const int *adjadr, const int *adj) {
  if (id < 0 || id >= O_MAXM) return -1;
  OMesh *ms = &m->mesh[id];
  free(ms->vert); free(ms->adjadr); free(ms->adj);
  ms->nvert = nvert;
  ms->vert = (double *)malloc(sizeof(double) * 3 * nvert);
  memcpy(ms->vert, vert, sizeof(double) * 3 * nvert);
  ms->adjadr = (int *)malloc(sizeof(int) * (nvert + 1));
  memcpy(ms->adjadr, adjadr, sizeof(int) * (nvert + 1));
  ms->adj = (int *)malloc(sizeof(int) * (adjadr[nvert] > 0 ? adjadr[nvert] : 1));
  memcpy(ms->adj, adj, sizeof(int) * adjadr[nvert]);
  return 0;
}

OData *orc_data_new(const OModel *m) { (void)m; return (OData *)calloc(1, sizeof(OData)); }
void orc_data_free(OData *d) { free(d); }

/* ------------------------------------------------------------------ position stage */
/* engine_core_smooth.c : mj_kinematics */
static void mj_kinematics(const OModel *m, OData *d) {
  for (int j = 0; j < m->njnt; j++)
    if (m->jnt_type[j] == O_JNT_FREE) quat_normalize(d->qpos + m->jnt_qposadr[j] + 3);
  memset(d->xpos, 0, 3 * sizeof(double));
  d->xquat[0] = 1; d->xquat[1] = d->xquat[2] = d->xquat[3] = 0;
  quat2mat(d->xmat, d->xquat);
  copy3(d->xipos, d->xpos);
  quat2mat(d->ximat, d->xquat);
  for (int b = 1; b < m->nbody; b++) {
    int pid = m->body_parentid[b];
    double *xpos = d->xpos + 3 * b, *xquat = d->xquat + 4 * b;
    int jadr = m->body_jntadr[b], jnum = m->body_jntnum[b];
    if (jnum == 1 && m->jnt_type[jadr] == O_JNT_FREE) {
      int qa = m->jnt_qposadr[jadr];
      copy3(xpos, d->qpos + qa);
      memcpy(xquat, d->qpos + qa + 3, 4 * sizeof(double));
      copy3(d->xanchor + 3 * jadr, xpos);
      rot_vec_quat(d->xaxis + 3 * jadr, m->jnt_axis + 3 * jadr, xquat);
    } else {
      double t[3];
      mulmat3vec(t, d->xmat + 9 * pid, m->body_pos + 3 * b);
      add3(xpos, d->xpos + 3 * pid, t);
      quat_mul(xquat, d->xquat + 4 * pid, m->body_quat + 4 * b);
      for (int k = 0; k < jnum; k++) {
        int j = jadr + k;
        double *anchor = d->xanchor + 3 * j, *axis = d->xaxis + 3 * j;
        rot_vec_quat(t, m->jnt_pos + 3 * j, xquat);
        add3(anchor, xpos, t);
        rot_vec_quat(axis, m->jnt_axis + 3 * j, xquat);
        double q = d->qpos[m->jnt_qposadr[j]] - m->qpos0[m->jnt_qposadr[j]];
        if (m->jnt_type[j] == O_JNT_SLIDE) {
          addscl3(xpos, xpos, axis, q);
        } else { /* hinge */
          double ql[4];
          axisangle2quat(ql, m->jnt_axis + 3 * j, q);
          quat_mul(xquat, xquat, ql);
          rot_vec_quat(t, m->jnt_pos + 3 * j, xquat);
          sub3(xpos, anchor, t);
        }
      }
    }
    quat_normalize(xquat);
    quat2mat(d->xmat + 9 * b, xquat);
    double t[3], qi[4];
    mulmat3vec(t, d->xmat + 9 * b, m->body_ipos + 3 * b);
    add3(d->xipos + 3 * b, xpos, t);
    quat_mul(qi, xquat, m->body_iquat + 4 * b);
    quat2mat(d->ximat + 9 * b, qi);
  }
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_bodyid[g];
    double t[3], q[4];
    mulmat3vec(t, d->xmat + 9 * b, m->geom_pos + 3 * g);
    add3(d->geom_xpos + 3 * g, d->xpos + 3 * b, t);
    quat_mul(q, d->xquat + 4 * b, m->geom_quat + 4 * g);
    quat_normalize(q);
    quat2mat(d->geom_xmat + 9 * g, q);
  }
}

/* engine_core_smooth.c : mj_comPos  (subtree CoM, cinert, cdof about the tree root's subtree CoM) */
static void mj_comPos(const OModel *m, OData *d) {
  double mass[O_MAXB];
  for (int b = 0; b < m->nbody; b++) {
    mass[b] = m->body_mass[b];
    scl3(d->subtree_com + 3 * b, d->xipos + 3 * b, m->body_mass[b]);
  }
  for (int b = m->nbody - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    add3(d->subtree_com + 3 * p, d->subtree_com + 3 * p, d->subtree_com + 3 * b);
    mass[p] += mass[b];
  }
  for (int b = 0; b < m->nbody; b++) {
    if (mass[b] < MINVAL) copy3(d->subtree_com + 3 * b, d->xipos + 3 * b);
    else scl3(d->subtree_com + 3 * b, d->subtree_com + 3 * b, 1.0 / mass[b]);
  }
  for (int b = 1; b < m->nbody; b++) {
    const double *com = d->subtree_com + 3 * m->body_rootid[b];
    const double *R = d->ximat + 9 * b, *I = m->body_inertia + 3 * b;
    double mb = m->body_mass[b], off[3];
    sub3(off, d->xipos + 3 * b, com);
    /* R diag(I) R^T */
    double M[9];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) M[3 * i + j] = R[3 * i] * I[0] * R[3 * j] + R[3 * i + 1] * I[1] * R[3 * j + 1] + R[3 * i + 2] * I[2] * R[3 * j + 2];
    double dd = dot3(off, off);
    double *c = d->cinert + 10 * b;
    c[0] = M[0] + mb * (dd - off[0] * off[0]);
    c[1] = M[4] + mb * (dd - off[1] * off[1]);
    c[2] = M[8] + mb * (dd - off[2] * off[2]);
    c[3] = M[1] - mb * off[0] * off[1];
    c[4] = M[2] - mb * off[0] * off[2];
    c[5] = M[5] - mb * off[1] * off[2];
    c[6] = mb * off[0]; c[7] = mb * off[1]; c[8] = mb * off[2];
    c[9] = mb;
  }
  memset(d->cinert, 0, 10 * sizeof(double));
  for (int j = 0; j < m->njnt; j++) {
    int b = m->jnt_bodyid[j], da = m->jnt_dofadr[j];
    const double *com = d->subtree_com + 3 * m->body_rootid[b];
    double off[3];
    if (m->jnt_type[j] == O_JNT_FREE) {
      for (int k = 0; k < 3; k++) {
        double *c = d->cdof + 6 * (da + k);
        memset(c, 0, 6 * sizeof(double));
        c[3 + k] = 1;
      }
      sub3(off, com, d->xpos + 3 * b);
      for (int k = 0; k < 3; k++) {
        double *c = d->cdof + 6 * (da + 3 + k);
        double ax[3] = {d->xmat[9 * b + k], d->xmat[9 * b + 3 + k], d->xmat[9 * b + 6 + k]};
        copy3(c, ax);
        cross3(c + 3, ax, off);
      }
    } else if (m->jnt_type[j] == O_JNT_SLIDE) {
      double *c = d->cdof + 6 * da;
      c[0] = c[1] = c[2] = 0;
      copy3(c + 3, d->xaxis + 3 * j);
    } else {
      double *c = d->cdof + 6 * da;
      sub3(off, com, d->xanchor + 3 * j);
      copy3(c, d->xaxis + 3 * j);
      cross3(c + 3, d->xaxis + 3 * j, off);
    }
  }
}

/* spatial inertia (10 numbers) times motion vector [ang;lin] -> force vector [torque;force] */
static void mul_inert_vec(double *r, const double *c, const double *v) {
  const double *w = v, *l = v + 3, *md = c + 6;
  double t[3];
  r[0] = c[0] * w[0] + c[3] * w[1] + c[4] * w[2];
  r[1] = c[3] * w[0] + c[1] * w[1] + c[5] * w[2];
  r[2] = c[4] * w[0] + c[5] * w[1] + c[2] * w[2];
  cross3(t, md, l);
  add3(r, r, t);
  cross3(t, md, w);
  r[3] = c[9] * l[0] - t[0]; r[4] = c[9] * l[1] - t[1]; r[5] = c[9] * l[2] - t[2];
}
static double dot6(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5]; }

/* engine_core_smooth.c : mj_crb + armature, then dense Cholesky (stands in for mj_factorM's L^T D L) */
static void mj_crb(const OModel *m, OData *d) {
  int nv = m->nv;
  double crb[O_MAXB * 10];
  memcpy(crb, d->cinert, sizeof(double) * 10 * m->nbody);
  for (int b = m->nbody - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    if (p > 0) for (int k = 0; k < 10; k++) crb[10 * p + k] += crb[10 * b + k];
  }
  memset(d->qM, 0, sizeof(double) * nv * nv);
  for (int i = 0; i < nv; i++) {
    double buf[6];
    mul_inert_vec(buf, crb + 10 * m->dof_bodyid[i], d->cdof + 6 * i);
    d->qM[i * nv + i] = dot6(d->cdof + 6 * i, buf) + m->dof_armature[i];
    for (int j = m->dof_parentid[i]; j >= 0; j = m->dof_parentid[j]) {
      double v = dot6(d->cdof + 6 * j, buf);
      d->qM[i * nv + j] = v;
      d->qM[j * nv + i] = v;
    }
  }
  memcpy(d->qLD, d->qM, sizeof(double) * nv * nv);
  chol_factor(d->qLD, nv);
}

/* engine_support.c : mj_jac (point on body) */
static void mj_jac(const OModel *m, const OData *d, double *jacp, double *jacr, const double *point, int body) {
  int nv = m->nv;
  if (jacp) memset(jacp, 0, sizeof(double) * 3 * nv);
  if (jacr) memset(jacr, 0, sizeof(double) * 3 * nv);
  while (body > 0 && m->body_dofnum[body] == 0) body = m->body_parentid[body];
  if (body <= 0) return;
  double off[3];
  sub3(off, point, d->subtree_com + 3 * m->body_rootid[body]);
  int i = m->body_dofadr[body] + m->body_dofnum[body] - 1;
  for (; i >= 0; i = m->dof_parentid[i]) {
    const double *c = d->cdof + 6 * i;
    if (jacr) { jacr[i] = c[0]; jacr[nv + i] = c[1]; jacr[2 * nv + i] = c[2]; }
    if (jacp) {
      double t[3];
      cross3(t, c, off);
      jacp[i] = c[3] + t[0]; jacp[nv + i] = c[4] + t[1]; jacp[2 * nv + i] = c[5] + t[2];
    }
  }
}
void orc_jac_body(const OModel *m, const OData *d, int body, double *jacp, double *jacr) { mj_jac(m, d, jacp, jacr, d->xpos + 3 * body, body); }

/* ------------------------------------------------------------------ collision */
typedef struct { double v[3], v1[3], v2[3]; } Sup;

/* engine_collision_convex.c : mjccd_support for a mesh geom — exhaustive arg-max over hull vertices */
static int support_geom(const OModel *m, const OData *d, int g, const double *dir, double *out) {
  const OMesh *ms = &m->mesh[m->geom_meshid[g]];
  const double *R = d->geom_xmat + 9 * g;
  double l[3];
  mulmat3Tvec(l, R, dir);
  int best = 0;
  double bv = -1e300;
  for (int i = 0; i < ms->nvert; i++) {
    double v = dot3(ms->vert + 3 * i, l);
    if (v > bv) { bv = v; best = i; }
  }
  double t[3];
  mulmat3vec(t, R, ms->vert + 3 * best);
  add3(out, d->geom_xpos + 3 * g, t);
  return best;
}
/* libccd support.c : __ccdSupport on the Minkowski difference obj1 - obj2, each inflated by margin/2 */
static void support_md(const OModel *m, const OData *d, int g1, int g2, double margin, const double *dir, Sup *s) {
  double n[3], nd[3];
  copy3(n, dir);
  normalize3(n);
  scl3(nd, n, -1);
  support_geom(m, d, g1, n, s->v1);
  support_geom(m, d, g2, nd, s->v2);
  addscl3(s->v1, s->v1, n, 0.5 * margin);
  addscl3(s->v2, s->v2, nd, 0.5 * margin);
  sub3(s->v, s->v1, s->v2);
}
static void portal_dir(const Sup *p, double *dir) { /* libccd mpr.c : portalDir */
  double a[3], b[3];
  sub3(a, p[2].v, p[1].v);
  sub3(b, p[3].v, p[1].v);
  cross3(dir, a, b);
  normalize3(dir);
}
static int portal_reach_tol(const Sup *p, const Sup *v4, const double *dir, double tol) {
  double dv1 = dot3(p[1].v, dir), dv2 = dot3(p[2].v, dir), dv3 = dot3(p[3].v, dir), dv4 = dot3(v4->v, dir);
  double d1 = dv4 - dv1, d2 = dv4 - dv2, d3 = dv4 - dv3;
  double mn = d1 < d2 ? d1 : d2;
  mn = mn < d3 ? mn : d3;
  return mn <= tol;
}
static void expand_portal(Sup *p, const Sup *v4) { /* libccd mpr.c : expandPortal */
  double v4v0[3];
  cross3(v4v0, v4->v, p[0].v);
  double dt = dot3(p[1].v, v4v0);
  if (dt > 0) {
    dt = dot3(p[2].v, v4v0);
    if (dt > 0) p[1] = *v4; else p[3] = *v4;
  } else {
    dt = dot3(p[3].v, v4v0);
    if (dt > 0) p[2] = *v4; else p[1] = *v4;
  }
}
/* closest point of triangle (a,b,c) to the origin */
static double origin_tri_closest(const double *a, const double *b, const double *c, double *w) {
  double ab[3], ac[3], ap[3];
  sub3(ab, b, a); sub3(ac, c, a); scl3(ap, a, -1);
  double d1 = dot3(ab, ap), d2 = dot3(ac, ap);
  if (d1 <= 0 && d2 <= 0) { copy3(w, a); return norm3(w); }
  double bp[3]; scl3(bp, b, -1);
  double d3 = dot3(ab, bp), d4 = dot3(ac, bp);
  if (d3 >= 0 && d4 <= d3) { copy3(w, b); return norm3(w); }
  double vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) { double v = d1 / (d1 - d3); addscl3(w, a, ab, v); return norm3(w); }
  double cp[3]; scl3(cp, c, -1);
  double d5 = dot3(ab, cp), d6 = dot3(ac, cp);
  if (d6 >= 0 && d5 <= d6) { copy3(w, c); return norm3(w); }
  double vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) { double v = d2 / (d2 - d6); addscl3(w, a, ac, v); return norm3(w); }
  double va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
    double v = (d4 - d3) / ((d4 - d3) + (d5 - d6)), bc[3];
    sub3(bc, c, b); addscl3(w, b, bc, v); return norm3(w);
  }
  double den = 1.0 / (va + vb + vc), v = vb * den, u = vc * den;
  addscl3(w, a, ab, v); addscl3(w, w, ac, u);
  return norm3(w);
}
static void find_pos(const Sup *p, double *pos) { /* libccd mpr.c : findPos */
  double dir[3], b[4], t[3], sum;
  portal_dir(p, dir);
  cross3(t, p[1].v, p[2].v); b[0] = dot3(t, p[3].v);
  cross3(t, p[3].v, p[2].v); b[1] = dot3(t, p[0].v);
  cross3(t, p[0].v, p[1].v); b[2] = dot3(t, p[3].v);
  cross3(t, p[2].v, p[1].v); b[3] = dot3(t, p[0].v);
  sum = b[0] + b[1] + b[2] + b[3];
  if (sum <= 0) {
    b[0] = 0;
    cross3(t, p[2].v, p[3].v); b[1] = dot3(t, dir);
    cross3(t, p[3].v, p[1].v); b[2] = dot3(t, dir);
    cross3(t, p[1].v, p[2].v); b[3] = dot3(t, dir);
    sum = b[1] + b[2] + b[3];
  }
  double inv = 1.0 / sum, p1[3] = {0, 0, 0}, p2[3] = {0, 0, 0};
  for (int i = 0; i < 4; i++) { addscl3(p1, p1, p[i].v1, b[i]); addscl3(p2, p2, p[i].v2, b[i]); }
  for (int k = 0; k < 3; k++) pos[k] = 0.5 * (p1[k] + p2[k]) * inv;
}
/* libccd mpr.c : ccdMPRPenetration.  Returns 1 and (depth, dir, pos) when the inflated shapes intersect. */
static int mpr_penetration(const OModel *m, const OData *d, int g1, int g2, double margin, double *depth, double *dir_out, double *pos) {
  Sup p[4], v4;
  double dir[3], va[3], vb[3];
  /* discoverPortal */
  copy3(p[0].v1, d->geom_xpos + 3 * g1);
  copy3(p[0].v2, d->geom_xpos + 3 * g2);
  sub3(p[0].v, p[0].v1, p[0].v2);
  if (norm3(p[0].v) < 1e-10) p[0].v[0] += 1e-10 * 10;
  scl3(dir, p[0].v, -1);
  normalize3(dir);
  support_md(m, d, g1, g2, margin, dir, &p[1]);
  if (dot3(p[1].v, dir) < 0) return 0;
  cross3(dir, p[0].v, p[1].v);
  int special = 0;
  if (norm3(dir) < 1e-10) {
    special = (norm3(p[1].v) < 1e-10) ? 1 : 2;
  } else {
    normalize3(dir);
    support_md(m, d, g1, g2, margin, dir, &p[2]);
    if (dot3(p[2].v, dir) < 0) return 0;
    sub3(va, p[1].v, p[0].v);
    sub3(vb, p[2].v, p[0].v);
    cross3(dir, va, vb);
    normalize3(dir);
    if (dot3(dir, p[0].v) > 0) { Sup t = p[1]; p[1] = p[2]; p[2] = t; scl3(dir, dir, -1); }
    for (int it = 0; it < 100; it++) {
      support_md(m, d, g1, g2, margin, dir, &p[3]);
      if (dot3(p[3].v, dir) < 0) return 0;
      int cont = 0;
      cross3(va, p[1].v, p[3].v);
      if (dot3(va, p[0].v) < -1e-14) { p[2] = p[3]; cont = 1; }
      if (!cont) {
        cross3(va, p[3].v, p[2].v);
        if (dot3(va, p[0].v) < -1e-14) { p[1] = p[3]; cont = 1; }
      }
      if (!cont) break;
      sub3(va, p[1].v, p[0].v);
      sub3(vb, p[2].v, p[0].v);
      cross3(dir, va, vb);
      normalize3(dir);
    }
  }
  if (special == 1) { /* findPenetrTouch */
    *depth = 0; dir_out[0] = dir_out[1] = dir_out[2] = 0;
    for (int k = 0; k < 3; k++) pos[k] = 0.5 * (p[1].v1[k] + p[1].v2[k]);
    return 1;
  }
  if (special == 2) { /* findPenetrSegment */
    for (int k = 0; k < 3; k++) pos[k] = 0.5 * (p[1].v1[k] + p[1].v2[k]);
    copy3(dir_out, p[1].v);
    *depth = normalize3(dir_out);
    return 1;
  }
  /* refinePortal */
  for (;;) {
    portal_dir(p, dir);
    if (dot3(dir, p[1].v) >= 0) break; /* portal encapsulates the origin */
    support_md(m, d, g1, g2, margin, dir, &v4);
    if (dot3(v4.v, dir) < 0 || portal_reach_tol(p, &v4, dir, m->mpr_tolerance)) return 0;
    expand_portal(p, &v4);
  }
  /* findPenetr */
  for (int it = 0;; it++) {
    portal_dir(p, dir);
    support_md(m, d, g1, g2, margin, dir, &v4);
    if (portal_reach_tol(p, &v4, dir, m->mpr_tolerance) || it > m->mpr_iterations) {
      double w[3];
      *depth = origin_tri_closest(p[1].v, p[2].v, p[3].v, w);
      if (*depth < 1e-14) copy3(dir_out, dir); else { copy3(dir_out, w); normalize3(dir_out); }
      find_pos(p, pos);
      return 1;
    }
    expand_portal(p, &v4);
  }
}

/* engine_util_misc / engine_collision_driver : mju_makeFrame */
static void make_frame(double *f) {
  normalize3(f);
  if (norm3(f + 3) < 0.5) {
    f[3] = f[4] = f[5] = 0;
    if (f[1] < 0.5 && f[1] > -0.5) f[4] = 1; else f[5] = 1;
  }
  double t = dot3(f, f + 3);
  addscl3(f + 3, f + 3, f, -t);
  normalize3(f + 3);
  cross3(f + 6, f, f + 3);
}

/* engine_collision_driver.c : mj_contactParam + contact bookkeeping in mj_collideGeoms */
static void add_contact(const OModel *m, OData *d, int g1, int g2, double dist, const double *pos, const double *normal, double margin) {
  if (d->ncon >= O_MAXCON) return;
  OContact *c = &d->contact[d->ncon++];
  memset(c, 0, sizeof(*c));
  c->dist = dist;
  copy3(c->pos, pos);
  copy3(c->frame, normal);
  make_frame(c->frame);
  c->geom1 = g1; c->geom2 = g2;
  double gap = fmax(m->geom_gap[g1], m->geom_gap[g2]);
  c->includemargin = margin - gap;
  c->dim = m->geom_condim[g1] > m->geom_condim[g2] ? m->geom_condim[g1] : m->geom_condim[g2];
  double f[3];
  for (int k = 0; k < 3; k++) f[k] = fmax(m->geom_friction[3 * g1 + k], m->geom_friction[3 * g2 + k]);
  c->friction[0] = f[0]; c->friction[1] = f[0]; c->friction[2] = f[1]; c->friction[3] = f[2]; c->friction[4] = f[2];
  /* solmix equal: plain average (positive solref) */
  for (int k = 0; k < 2; k++) c->solref[k] = 0.5 * (m->geom_solref[2 * g1 + k] + m->geom_solref[2 * g2 + k]);
  for (int k = 0; k < 5; k++) c->solimp[k] = 0.5 * (m->geom_solimp[5 * g1 + k] + m->geom_solimp[5 * g2 + k]);
}

/* engine_collision_convex.c : mjc_PlaneConvex for a mesh: support vertex, then hull-graph neighbours of it
 * that are also within the margin, up to 3 contacts in total (neighbour order = ascending hull-local id;
 * the upstream order is qhull's and cannot be reproduced without it — DESIGN.md "unpinned"). */
static void collide_plane_mesh(const OModel *m, OData *d, int g1, int g2, double margin) {
  const double *R1 = d->geom_xmat + 9 * g1, *p1 = d->geom_xpos + 3 * g1;
  double n[3] = {R1[2], R1[5], R1[8]}, nn[3], v[3], dif[3], pos[3];
  scl3(nn, n, -1);
  int vi = support_geom(m, d, g2, nn, v);
  sub3(dif, v, p1);
  double dist = dot3(dif, n);
  if (dist > margin) return;
  addscl3(pos, v, n, -0.5 * dist);
  add_contact(m, d, g1, g2, dist, pos, n, margin);
  const OMesh *ms = &m->mesh[m->geom_meshid[g2]];
  int cnt = 1;
  for (int e = ms->adjadr[vi]; e < ms->adjadr[vi + 1] && cnt < 3; e++) {
    double t[3];
    mulmat3vec(t, d->geom_xmat + 9 * g2, ms->vert + 3 * ms->adj[e]);
    add3(v, d->geom_xpos + 3 * g2, t);
    sub3(dif, v, p1);
    dist = dot3(dif, n);
    if (dist > margin) continue;
    addscl3(pos, v, n, -0.5 * dist);
    add_contact(m, d, g1, g2, dist, pos, n, margin);
    cnt++;
  }
}
/* engine_collision_primitive.c : mjc_PlaneBox — the corners in index order (bit 0 = x, 1 = y, 2 = z sign), keeping those
 * within the margin that do not point away from the plane, at most 4 (gripper_two_fingers.xml:133 is the only box). */
static void collide_plane_box(const OModel *m, OData *d, int g1, int g2, double margin) {
  const double *R1 = d->geom_xmat + 9 * g1, *R2 = d->geom_xmat + 9 * g2, *sz = m->geom_size + 3 * g2;
  double n[3] = {R1[2], R1[5], R1[8]}, dif[3], vec[3], corner[3], pos[3];
  sub3(dif, d->geom_xpos + 3 * g2, d->geom_xpos + 3 * g1);
  double dist = dot3(dif, n);
  int cnt = 0;
  for (int i = 0; i < 8 && cnt < 4; i++) {
    vec[0] = (i & 1) ? sz[0] : -sz[0]; vec[1] = (i & 2) ? sz[1] : -sz[1]; vec[2] = (i & 4) ? sz[2] : -sz[2];
    mulmat3vec(corner, R2, vec);
    double ldist = dot3(n, corner);
    if (dist + ldist > margin || ldist > 0) continue;
    addscl3(pos, corner, n, -0.5 * (dist + ldist));
    add3(pos, pos, d->geom_xpos + 3 * g2);
    add_contact(m, d, g1, g2, dist + ldist, pos, n, margin);
    cnt++;
  }
}
/* engine_collision_convex.c : mjc_Convex (libccd MPR, one contact) */
static void collide_convex(const OModel *m, OData *d, int g1, int g2, double margin) {
  double depth, dir[3], pos[3];
  if (!mpr_penetration(m, d, g1, g2, margin, &depth, dir, pos)) return;
  if (norm3(dir) < 0.5) return;
  add_contact(m, d, g1, g2, margin - depth, pos, dir, margin);
}
/* engine_collision_driver.c : mj_collision over the pre-filtered pair list + bounding-sphere test of mj_collideGeoms */
static void mj_collision(const OModel *m, OData *d) {
  d->ncon = 0;
  for (int p = 0; p < m->npair; p++) {
    int g1 = m->pair_geom1[p], g2 = m->pair_geom2[p];
    double margin = fmax(m->geom_margin[g1], m->geom_margin[g2]);
    if (m->geom_type[g1] == O_GEOM_PLANE) {
      const double *R1 = d->geom_xmat + 9 * g1;
      double n[3] = {R1[2], R1[5], R1[8]}, dif[3];
      sub3(dif, d->geom_xpos + 3 * g2, d->geom_xpos + 3 * g1);
      if (dot3(dif, n) > margin + m->geom_rbound[g2]) continue;
      if (m->geom_type[g2] == O_GEOM_BOX) collide_plane_box(m, d, g1, g2, margin);
      else collide_plane_mesh(m, d, g1, g2, margin);
    } else {
      double dif[3];
      sub3(dif, d->geom_xpos + 3 * g2, d->geom_xpos + 3 * g1);
      double bound = m->geom_rbound[g1] + m->geom_rbound[g2] + margin;
      if (dot3(dif, dif) > bound * bound) continue;
      collide_convex(m, d, g1, g2, margin);
    }
  }
}

void orc_forward_position(const OModel *m, OData *d) {
  mj_kinematics(m, d);
  mj_comPos(m, d);
  mj_crb(m, d);
  mj_collision(m, d);
}

/* ------------------------------------------------------------------ constraints */
/* engine_core_constraint.c : getimpedance */
static void get_impedance(const double *solimp, double pos, double margin, double *imp) {
  if (solimp[0] == solimp[1] || solimp[2] <= MINVAL) { *imp = 0.5 * (solimp[0] + solimp[1]); return; }
  double x = (pos - margin) / solimp[2];
  if (x < 0) x = -x;
  if (x >= 1) { *imp = solimp[1]; return; }
  if (x <= 0) { *imp = solimp[0]; return; }
  double y, p = solimp[4], mid = solimp[3];
  if (p == 1) y = x;
  else if (x <= mid) y = pow(x, p) / pow(mid, p - 1);
  else y = 1 - pow(1 - x, p) / pow(1 - mid, p - 1);
  *imp = solimp[0] + y * (solimp[1] - solimp[0]);
}

/* engine_core_constraint.c : mj_makeConstraint (limits, contacts), mj_diagApprox, mj_makeImpedance, mj_referenceConstraint */
static void mj_makeConstraint(const OModel *m, OData *d) {
  int nv = m->nv, ne = 0;
  double solref_row[O_MAXEFC][2], solimp_row[O_MAXEFC][5];
  /* joint limits */
  for (int j = 0; j < m->njnt; j++) {
    if (!m->jnt_limited[j] || m->jnt_type[j] == O_JNT_FREE) continue;
    double value = d->qpos[m->jnt_qposadr[j]];
    for (int side = -1; side <= 1; side += 2) {
      double dist = side * (m->jnt_range[2 * j + (side + 1) / 2] - value);
      if (dist < 0 /* jnt_margin = 0 */) {
        double *J = d->efc_J + ne * nv;
        memset(J, 0, sizeof(double) * nv);
        J[m->jnt_dofadr[j]] = -(double)side;
        d->efc_type[ne] = 0; d->efc_id[ne] = j;
        d->efc_pos[ne] = dist; d->efc_margin[ne] = 0;
        d->efc_diagApprox[ne] = m->dof_invweight0[m->jnt_dofadr[j]];
        memcpy(solref_row[ne], m->jnt_solref, sizeof(double) * 2);
        memcpy(solimp_row[ne], m->jnt_solimp, sizeof(double) * 5);
        ne++;
      }
    }
  }
  /* contacts, elliptic cones: dim rows each */
  double jp1[3 * O_MAXV], jr1[3 * O_MAXV], jp2[3 * O_MAXV], jr2[3 * O_MAXV];
  for (int ci = 0; ci < d->ncon; ci++) {
    OContact *c = &d->contact[ci];
    if (ne + c->dim > O_MAXEFC) { c->efc_address = -1; continue; }
    int b1 = m->geom_bodyid[c->geom1], b2 = m->geom_bodyid[c->geom2];
    mj_jac(m, d, jp1, jr1, c->pos, b1);
    mj_jac(m, d, jp2, jr2, c->pos, b2);
    c->efc_address = ne;
    double tran = m->body_invweight0[2 * b1] + m->body_invweight0[2 * b2];
    double rot = m->body_invweight0[2 * b1 + 1] + m->body_invweight0[2 * b2 + 1];
    for (int r = 0; r < c->dim; r++) {
      double *J = d->efc_J + (ne + r) * nv;
      const double *ax = c->frame + 3 * (r < 3 ? r : r - 3);
      for (int i = 0; i < nv; i++) {
        if (r < 3) J[i] = ax[0] * (jp2[i] - jp1[i]) + ax[1] * (jp2[nv + i] - jp1[nv + i]) + ax[2] * (jp2[2 * nv + i] - jp1[2 * nv + i]);
        else J[i] = ax[0] * (jr2[i] - jr1[i]) + ax[1] * (jr2[nv + i] - jr1[nv + i]) + ax[2] * (jr2[2 * nv + i] - jr1[2 * nv + i]);
      }
      d->efc_type[ne + r] = r == 0 ? 1 : 2;
      d->efc_id[ne + r] = ci;
      d->efc_pos[ne + r] = r == 0 ? c->dist : 0;
      d->efc_margin[ne + r] = r == 0 ? c->includemargin : 0;
      d->efc_diagApprox[ne + r] = r < 3 ? tran : rot;
      memcpy(solref_row[ne + r], c->solref, sizeof(double) * 2);
      memcpy(solimp_row[ne + r], c->solimp, sizeof(double) * 5);
    }
    ne += c->dim;
  }
  d->nefc = ne;
  /* mj_makeImpedance */
  for (int i = 0; i < ne; i++) {
    double imp, tc = solref_row[i][0], dr = solref_row[i][1], dmax = solimp_row[i][1];
    if (tc < 2 * m->timestep) tc = 2 * m->timestep; /* refsafe */
    get_impedance(solimp_row[i], d->efc_pos[i], d->efc_margin[i], &imp);
    double R = (1 - imp) * d->efc_diagApprox[i] / imp;
    d->efc_R[i] = R > MINVAL ? R : MINVAL;
    double K = 1.0 / (dmax * dmax * tc * tc * dr * dr), B = 2.0 / (dmax * tc);
    d->efc_KBIP[4 * i] = K; d->efc_KBIP[4 * i + 1] = B; d->efc_KBIP[4 * i + 2] = imp; d->efc_KBIP[4 * i + 3] = 0;
  }
  for (int ci = 0; ci < d->ncon; ci++) { /* friction rows of elliptic cones */
    OContact *c = &d->contact[ci];
    int i = c->efc_address;
    if (i < 0 || c->dim < 2) continue;
    d->efc_R[i + 1] = d->efc_R[i] / fmax(MINVAL, m->impratio);
    for (int j = 2; j < c->dim; j++) d->efc_R[i + j] = d->efc_R[i + 1] * c->friction[0] * c->friction[0] / (c->friction[j - 1] * c->friction[j - 1]);
    c->mu = c->friction[0] * sqrt(d->efc_R[i + 1] / d->efc_R[i]);
  }
  /* mj_referenceConstraint */
  for (int i = 0; i < ne; i++) {
    d->efc_D[i] = 1.0 / d->efc_R[i];
    double v = 0;
    for (int k = 0; k < nv; k++) v += d->efc_J[i * nv + k] * d->qvel[k];
    d->efc_vel[i] = v;
    d->efc_aref[i] = -d->efc_KBIP[4 * i + 1] * v - d->efc_KBIP[4 * i] * d->efc_KBIP[4 * i + 2] * (d->efc_pos[i] - d->efc_margin[i]);
  }
}

/* ------------------------------------------------------------------ velocity / force stages */
static void cross_motion(double *r, const double *v, const double *mv) {
  double t[3];
  cross3(r, v, mv);
  cross3(r + 3, v, mv + 3);
  cross3(t, v + 3, mv);
  add3(r + 3, r + 3, t);
}
static void cross_force(double *r, const double *v, const double *f) {
  double t[3];
  cross3(r, v, f);
  cross3(t, v + 3, f + 3);
  add3(r, r, t);
  cross3(r + 3, v, f + 3);
}
/* engine_core_smooth.c : mj_comVel */
static void mj_comVel(const OModel *m, OData *d) {
  memset(d->cvel, 0, 6 * sizeof(double));
  for (int b = 1; b < m->nbody; b++) {
    double cvel[6];
    memcpy(cvel, d->cvel + 6 * m->body_parentid[b], sizeof cvel);
    int da = m->body_dofadr[b];
    for (int k = 0; k < m->body_jntnum[b]; k++) {
      int j = m->body_jntadr[b] + k, a = m->jnt_dofadr[j];
      if (m->jnt_type[j] == O_JNT_FREE) {
        memset(d->cdof_dot + 6 * a, 0, 18 * sizeof(double));
        for (int i = 0; i < 3; i++) for (int c = 0; c < 6; c++) cvel[c] += d->cdof[6 * (a + i) + c] * d->qvel[a + i];
        for (int i = 3; i < 6; i++) cross_motion(d->cdof_dot + 6 * (a + i), cvel, d->cdof + 6 * (a + i));
        for (int i = 3; i < 6; i++) for (int c = 0; c < 6; c++) cvel[c] += d->cdof[6 * (a + i) + c] * d->qvel[a + i];
      } else {
        cross_motion(d->cdof_dot + 6 * a, cvel, d->cdof + 6 * a);
        for (int c = 0; c < 6; c++) cvel[c] += d->cdof[6 * a + c] * d->qvel[a];
      }
    }
    (void)da;
    memcpy(d->cvel + 6 * b, cvel, sizeof cvel);
  }
}
/* engine_core_smooth.c : mj_rne (flg_acc = 0) -> qfrc_bias */
static void mj_rne(const OModel *m, OData *d) {
  double cacc[O_MAXB * 6], cfrc[O_MAXB * 6];
  memset(cacc, 0, sizeof cacc);
  memset(cfrc, 0, sizeof cfrc);
  cacc[3] = -m->gravity[0]; cacc[4] = -m->gravity[1]; cacc[5] = -m->gravity[2];
  for (int b = 1; b < m->nbody; b++) {
    memcpy(cacc + 6 * b, cacc + 6 * m->body_parentid[b], 6 * sizeof(double));
    int da = m->body_dofadr[b];
    for (int k = 0; k < m->body_dofnum[b]; k++)
      for (int c = 0; c < 6; c++) cacc[6 * b + c] += d->cdof_dot[6 * (da + k) + c] * d->qvel[da + k];
    double t1[6], t2[6];
    mul_inert_vec(t1, d->cinert + 10 * b, cacc + 6 * b);
    mul_inert_vec(t2, d->cinert + 10 * b, d->cvel + 6 * b);
    cross_force(cfrc + 6 * b, d->cvel + 6 * b, t2);
    for (int c = 0; c < 6; c++) cfrc[6 * b + c] += t1[c];
  }
  for (int b = m->nbody - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    if (p > 0) for (int c = 0; c < 6; c++) cfrc[6 * p + c] += cfrc[6 * b + c];
  }
  for (int i = 0; i < m->nv; i++) d->qfrc_bias[i] = dot6(d->cdof + 6 * i, cfrc + 6 * m->dof_bodyid[i]);
}
/* engine_forward.c : mj_fwdVelocity (passive) + mj_fwdActuation + mj_fwdAcceleration */
static void forward_smooth(const OModel *m, OData *d, int actuation) {
  int nv = m->nv;
  mj_comVel(m, d);
  for (int i = 0; i < nv; i++) d->qfrc_passive[i] = -m->dof_damping[i] * d->qvel[i];
  mj_rne(m, d);
  memset(d->qfrc_actuator, 0, sizeof(double) * nv);
  if (actuation)
    for (int u = 0; u < m->nu; u++) {
      double c = d->ctrl[u];
      if (c < m->act_ctrlrange[2 * u]) c = m->act_ctrlrange[2 * u];
      if (c > m->act_ctrlrange[2 * u + 1]) c = m->act_ctrlrange[2 * u + 1];
      d->qfrc_actuator[m->act_dofid[u]] += m->act_gear[u] * c;
    }
  /* mj_xfrcAccumulate: Cartesian force/torque applied at xipos */
  memset(d->qfrc_applied, 0, sizeof(double) * nv);
  double jp[3 * O_MAXV], jr[3 * O_MAXV];
  for (int b = 1; b < m->nbody; b++) {
    const double *x = d->xfrc_applied + 6 * b;
    if (x[0] == 0 && x[1] == 0 && x[2] == 0 && x[3] == 0 && x[4] == 0 && x[5] == 0) continue;
    mj_jac(m, d, jp, jr, d->xipos + 3 * b, b);
    for (int i = 0; i < nv; i++)
      d->qfrc_applied[i] += jp[i] * x[0] + jp[nv + i] * x[1] + jp[2 * nv + i] * x[2] + jr[i] * x[3] + jr[nv + i] * x[4] + jr[2 * nv + i] * x[5];
  }
  for (int i = 0; i < nv; i++) {
    d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_applied[i] + d->qfrc_actuator[i];
    d->qacc_smooth[i] = d->qfrc_smooth[i];
  }
  chol_solve(d->qLD, nv, d->qacc_smooth);
}

/* ------------------------------------------------------------------ Newton solver (engine_solver.c : mj_solNewton, primal) */
typedef struct {
  int nv, nefc;
  double qacc[O_MAXV], Ma[O_MAXV], jar[O_MAXEFC], grad[O_MAXV], search[O_MAXV], Mv[O_MAXV], Jv[O_MAXEFC];
  double cost, gauss;
  int state[O_MAXEFC]; /* 0 satisfied, 1 quadratic, 2 cone (first row of contact) */
} Newton;

/* mj_constraintUpdate: forces, states and cost at jar; optionally the cone Hessian blocks */
static double constraint_update(const OModel *m, OData *d, const double *jar, double *force, int *state) {
  double cost = 0;
  int i = 0;
  while (i < d->nefc) {
    if (d->efc_type[i] == 0) {
      if (jar[i] < 0) { force[i] = -d->efc_D[i] * jar[i]; cost += 0.5 * d->efc_D[i] * jar[i] * jar[i]; state[i] = 1; }
      else { force[i] = 0; state[i] = 0; }
      i++;
      continue;
    }
    const OContact *c = &d->contact[d->efc_id[i]];
    int dim = c->dim;
    double mu = c->mu, U[6], N, T = 0;
    U[0] = jar[i] * mu;
    for (int j = 1; j < dim; j++) { U[j] = jar[i + j] * c->friction[j - 1]; T += U[j] * U[j]; }
    N = U[0];
    T = sqrt(T);
    if (N >= mu * T || (T <= 0 && N >= 0)) {
      for (int j = 0; j < dim; j++) { force[i + j] = 0; state[i + j] = 0; }
    } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
      for (int j = 0; j < dim; j++) {
        force[i + j] = -d->efc_D[i + j] * jar[i + j];
        cost += 0.5 * d->efc_D[i + j] * jar[i + j] * jar[i + j];
        state[i + j] = 1;
      }
    } else {
      double Dm = d->efc_D[i] / (mu * mu * (1 + mu * mu)), NmT = N - mu * T;
      cost += 0.5 * Dm * NmT * NmT;
      force[i] = -Dm * NmT * mu;
      for (int j = 1; j < dim; j++) force[i + j] = -force[i] / T * U[j] * c->friction[j - 1];
      for (int j = 0; j < dim; j++) state[i + j] = 2;
    }
    i += dim;
  }
  return cost;
}

/* Hessian H = M + J^T (D or cone block) J, dense */
static void make_hessian(const OModel *m, const OData *d, const double *jar, const int *state, double *H) {
  int nv = m->nv;
  memcpy(H, d->qM, sizeof(double) * nv * nv);
  int i = 0;
  while (i < d->nefc) {
    if (d->efc_type[i] == 0 || state[i] != 2) {
      int dim = d->efc_type[i] == 0 ? 1 : d->contact[d->efc_id[i]].dim;
      for (int j = 0; j < dim; j++) {
        if (state[i + j] != 1) continue;
        const double *J = d->efc_J + (i + j) * nv;
        double D = d->efc_D[i + j];
        for (int a = 0; a < nv; a++) {
          if (J[a] == 0) continue;
          for (int b = 0; b < nv; b++) H[a * nv + b] += D * J[a] * J[b];
        }
      }
      i += dim;
      continue;
    }
    /* engine_solver.c : HessianCone */
    const OContact *c = &d->contact[d->efc_id[i]];
    int dim = c->dim;
    double mu = c->mu, U[6], T = 0, N, scale[6], hc[36];
    U[0] = jar[i] * mu; scale[0] = mu;
    for (int j = 1; j < dim; j++) { scale[j] = c->friction[j - 1]; U[j] = jar[i + j] * scale[j]; T += U[j] * U[j]; }
    T = sqrt(T); N = U[0];
    double Dm = d->efc_D[i] / (mu * mu * (1 + mu * mu));
    memset(hc, 0, sizeof hc);
    hc[0] = 1;
    for (int j = 1; j < dim; j++) { hc[j] = -mu / T * U[j]; hc[j * dim] = hc[j]; }
    double scl = mu * N / (T * T * T);
    for (int j = 1; j < dim; j++)
      for (int k = 1; k < dim; k++) hc[j * dim + k] = scl * U[j] * U[k];
    for (int j = 1; j < dim; j++) hc[j * dim + j] += mu * mu - mu * N / T;
    for (int j = 0; j < dim; j++)
      for (int k = 0; k < dim; k++) hc[j * dim + k] *= Dm * scale[j] * scale[k];
    for (int j = 0; j < dim; j++)
      for (int k = 0; k < dim; k++) {
        double h = hc[j * dim + k];
        if (h == 0) continue;
        const double *Jj = d->efc_J + (i + j) * nv, *Jk = d->efc_J + (i + k) * nv;
        for (int a = 0; a < nv; a++) {
          if (Jj[a] == 0) continue;
          for (int b = 0; b < nv; b++) H[a * nv + b] += h * Jj[a] * Jk[b];
        }
      }
    i += dim;
  }
}

/* cost and its first two derivatives along qacc + alpha*search (engine_solver.c : PrimalEval) */
static void line_eval(const OModel *m, const OData *d, const Newton *s, double alpha, double q0, double q1, double q2, double *f, double *df, double *ddf) {
  (void)m;
  double F = q0 + alpha * q1 + 0.5 * alpha * alpha * q2, D1 = q1 + alpha * q2, D2 = q2;
  int i = 0;
  while (i < d->nefc) {
    if (d->efc_type[i] == 0) {
      double x = s->jar[i] + alpha * s->Jv[i];
      if (x < 0) { F += 0.5 * d->efc_D[i] * x * x; D1 += d->efc_D[i] * x * s->Jv[i]; D2 += d->efc_D[i] * s->Jv[i] * s->Jv[i]; }
      i++;
      continue;
    }
    const OContact *c = &d->contact[d->efc_id[i]];
    int dim = c->dim;
    double mu = c->mu, U[6], V[6], T = 0, UV = 0, VV = 0;
    U[0] = (s->jar[i] + alpha * s->Jv[i]) * mu; V[0] = s->Jv[i] * mu;
    for (int j = 1; j < dim; j++) {
      U[j] = (s->jar[i + j] + alpha * s->Jv[i + j]) * c->friction[j - 1];
      V[j] = s->Jv[i + j] * c->friction[j - 1];
      T += U[j] * U[j]; UV += U[j] * V[j]; VV += V[j] * V[j];
    }
    T = sqrt(T);
    double N = U[0];
    if (N >= mu * T || (T <= 0 && N >= 0)) {
      /* top zone: nothing */
    } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
      for (int j = 0; j < dim; j++) {
        double x = s->jar[i + j] + alpha * s->Jv[i + j], D = d->efc_D[i + j];
        F += 0.5 * D * x * x; D1 += D * x * s->Jv[i + j]; D2 += D * s->Jv[i + j] * s->Jv[i + j];
      }
    } else {
      double Dm = d->efc_D[i] / (mu * mu * (1 + mu * mu)), NmT = N - mu * T;
      double N1 = V[0], T1 = UV / T, T2 = (VV - T1 * T1) / T;
      F += 0.5 * Dm * NmT * NmT;
      D1 += Dm * NmT * (N1 - mu * T1);
      D2 += Dm * ((N1 - mu * T1) * (N1 - mu * T1) + NmT * (-mu * T2));
    }
    i += dim;
  }
  *f = F; *df = D1; *ddf = D2;
}

/* exact 1-D minimisation (engine_solver.c : PrimalSearch restated as a safeguarded Newton on f') */
static double line_search(const OModel *m, const OData *d, const Newton *s, double q0, double q1, double q2, double gtol) {
  double f, df, ddf, lo = 0, hi = -1, dlo, a;
  line_eval(m, d, s, 0, q0, q1, q2, &f, &df, &ddf);
  if (df >= 0 || ddf <= 0) return 0;
  dlo = df;
  a = -df / ddf;
  for (int it = 0; it < 60; it++) {
    line_eval(m, d, s, a, q0, q1, q2, &f, &df, &ddf);
    if (fabs(df) < gtol) return a;
    if (df < 0) { lo = a; dlo = df; } else { hi = a; }
    double an = a - df / ddf;
    if (hi < 0) { /* no upper bracket yet: keep stepping forward */
      if (an <= lo) an = 2 * lo + 1e-12;
    } else if (!(an > lo && an < hi)) {
      an = 0.5 * (lo + hi);
    }
    if (hi > 0 && (hi - lo) < 1e-15 * fmax(1.0, hi)) return 0.5 * (lo + hi);
    a = an;
  }
  (void)dlo;
  return a;
}

static void mj_fwdConstraint(const OModel *m, OData *d) {
  int nv = m->nv, ne = d->nefc;
  if (ne == 0) {
    memcpy(d->qacc, d->qacc_smooth, sizeof(double) * nv);
    memset(d->qfrc_constraint, 0, sizeof(double) * nv);
    d->solver_iter = 0;
    return;
  }
  static __thread Newton S;
  Newton *s = &S;
  double H[O_MAXV * O_MAXV], scale = 1.0 / (m->meaninertia * (nv > 1 ? nv : 1));
  /* warm start: pick the better of qacc_warmstart and qacc_smooth */
  double best_cost = 0;
  for (int trial = 0; trial < 2; trial++) {
    const double *q = trial == 0 ? d->qacc_warmstart : d->qacc_smooth;
    double jar[O_MAXEFC], force[O_MAXEFC], Ma[O_MAXV];
    int st[O_MAXEFC];
    for (int i = 0; i < ne; i++) {
      double v = 0;
      for (int k = 0; k < nv; k++) v += d->efc_J[i * nv + k] * q[k];
      jar[i] = v - d->efc_aref[i];
    }
    double g = 0;
    for (int i = 0; i < nv; i++) {
      double v = 0;
      for (int k = 0; k < nv; k++) v += d->qM[i * nv + k] * q[k];
      Ma[i] = v;
      g += 0.5 * (v - d->qfrc_smooth[i]) * (q[i] - d->qacc_smooth[i]);
    }
    double c = g + constraint_update(m, d, jar, force, st);
    if (trial == 0 || c < best_cost) {
      best_cost = c;
      memcpy(s->qacc, q, sizeof(double) * nv);
      memcpy(s->Ma, Ma, sizeof(double) * nv);
      memcpy(s->jar, jar, sizeof(double) * ne);
    }
  }
  double cost = constraint_update(m, d, s->jar, d->efc_force, s->state);
  double gauss = 0;
  for (int i = 0; i < nv; i++) gauss += 0.5 * (s->Ma[i] - d->qfrc_smooth[i]) * (s->qacc[i] - d->qacc_smooth[i]);
  cost += gauss;
  int iter = 0;
  double improvement = 0, gradient = 0;
  for (; iter < m->iterations; iter++) {
    /* gradient and Newton direction */
    for (int i = 0; i < nv; i++) {
      double v = 0;
      for (int r = 0; r < ne; r++) v += d->efc_J[r * nv + i] * d->efc_force[r];
      d->qfrc_constraint[i] = v;
      s->grad[i] = s->Ma[i] - d->qfrc_smooth[i] - v;
    }
    gradient = 0;
    for (int i = 0; i < nv; i++) gradient += s->grad[i] * s->grad[i];
    gradient = scale * sqrt(gradient);
    if (iter > 0 && (improvement < m->tolerance || gradient < m->tolerance)) break;
    if (iter == 0 && gradient < m->tolerance) break;
    make_hessian(m, d, s->jar, s->state, H);
    chol_factor(H, nv);
    for (int i = 0; i < nv; i++) s->search[i] = -s->grad[i];
    chol_solve(H, nv, s->search);
    /* line search set-up */
    double q1 = 0, q2 = 0, snorm = 0;
    for (int i = 0; i < nv; i++) {
      double v = 0;
      for (int k = 0; k < nv; k++) v += d->qM[i * nv + k] * s->search[k];
      s->Mv[i] = v;
      q1 += s->search[i] * (s->Ma[i] - d->qfrc_smooth[i]);
      q2 += s->search[i] * v;
      snorm += s->search[i] * s->search[i];
    }
    snorm = sqrt(snorm);
    for (int r = 0; r < ne; r++) {
      double v = 0;
      for (int k = 0; k < nv; k++) v += d->efc_J[r * nv + k] * s->search[k];
      s->Jv[r] = v;
    }
    if (snorm < MINVAL) break;
    double gtol = m->tolerance * 0.01 * snorm / scale;
    double alpha = line_search(m, d, s, gauss, q1, q2, gtol);
    if (alpha == 0) break;
    for (int i = 0; i < nv; i++) { s->qacc[i] += alpha * s->search[i]; s->Ma[i] += alpha * s->Mv[i]; }
    for (int r = 0; r < ne; r++) s->jar[r] += alpha * s->Jv[r];
    double old = cost;
    gauss = 0;
    for (int i = 0; i < nv; i++) gauss += 0.5 * (s->Ma[i] - d->qfrc_smooth[i]) * (s->qacc[i] - d->qacc_smooth[i]);
    cost = gauss + constraint_update(m, d, s->jar, d->efc_force, s->state);
    improvement = scale * (old - cost);
  }
  for (int i = 0; i < nv; i++) {
    double v = 0;
    for (int r = 0; r < ne; r++) v += d->efc_J[r * nv + i] * d->efc_force[r];
    d->qfrc_constraint[i] = v;
  }
  memcpy(d->qacc, s->qacc, sizeof(double) * nv);
  d->solver_iter = iter;
  d->solver_improvement = improvement;
  d->solver_gradient = gradient;
}

/* engine_forward.c : mj_Euler with implicit joint damping, mj_advance / mj_integratePos */
static void mj_Euler(const OModel *m, OData *d) {
  int nv = m->nv;
  double h = m->timestep, A[O_MAXV * O_MAXV], qacc[O_MAXV];
  int damped = 0;
  for (int i = 0; i < nv; i++) if (m->dof_damping[i] > 0) damped = 1;
  if (!damped) memcpy(qacc, d->qacc, sizeof(double) * nv);
  else {
    memcpy(A, d->qM, sizeof(double) * nv * nv);
    for (int i = 0; i < nv; i++) { A[i * nv + i] += h * m->dof_damping[i]; qacc[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i]; }
    chol_factor(A, nv);
    chol_solve(A, nv, qacc);
  }
  for (int i = 0; i < nv; i++) d->qvel[i] += h * qacc[i];
  for (int j = 0; j < m->njnt; j++) {
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    if (m->jnt_type[j] == O_JNT_FREE) {
      for (int k = 0; k < 3; k++) d->qpos[qa + k] += h * d->qvel[da + k];
      double w[3] = {d->qvel[da + 3], d->qvel[da + 4], d->qvel[da + 5]};
      double ang = normalize3(w) * h, dq[4];
      if (ang > 0) {
        axisangle2quat(dq, w, ang);
        quat_mul(d->qpos + qa + 3, d->qpos + qa + 3, dq);
        quat_normalize(d->qpos + qa + 3);
      }
    } else {
      d->qpos[qa] += h * d->qvel[da];
    }
  }
  d->time += h;
}

/* dm_control Physics.step(), legacy mode: mj_step2 (uses the position stage already computed for the
 * current qpos) followed by mj_step1 (position + velocity stage of the NEW state). */
void orc_step(const OModel *m, OData *d) {
  forward_smooth(m, d, 1);
  mj_makeConstraint(m, d);
  mj_fwdConstraint(m, d);
  memcpy(d->qacc_warmstart, d->qacc, sizeof(double) * m->nv);
  mj_Euler(m, d);
  orc_forward_position(m, d);
}

/* dm_control Physics.reset(): mj_resetData then forward() with actuation disabled */
void orc_reset(const OModel *m, OData *d) {
  memset(d, 0, sizeof(*d));
  memcpy(d->qpos, m->qpos0, sizeof(double) * m->nq);
  orc_forward_position(m, d);
  forward_smooth(m, d, 0);
  mj_makeConstraint(m, d);
  mj_fwdConstraint(m, d);
  memcpy(d->qacc_warmstart, d->qacc, sizeof(double) * m->nv);
}

/* engine_setconst.c : mj_setConst — body_rootid, invweight0, meaninertia at qpos0 */
void orc_model_finalize(OModel *m) {
  for (int b = 0; b < m->nbody; b++) {
    int r = b;
    while (r > 0 && m->body_parentid[r] > 0) r = m->body_parentid[r];
    m->body_rootid[b] = r;
  }
  OData *d = orc_data_new(m);
  memcpy(d->qpos, m->qpos0, sizeof(double) * m->nq);
  mj_kinematics(m, d);
  mj_comPos(m, d);
  mj_crb(m, d);
  int nv = m->nv;
  double tr = 0;
  for (int i = 0; i < nv; i++) tr += d->qM[i * nv + i];
  m->meaninertia = nv ? tr / nv : 1;
  double Minv[O_MAXV * O_MAXV];
  for (int i = 0; i < nv; i++) {
    double e[O_MAXV];
    memset(e, 0, sizeof e);
    e[i] = 1;
    chol_solve(d->qLD, nv, e);
    for (int k = 0; k < nv; k++) Minv[k * nv + i] = e[k];
  }
  for (int b = 0; b < m->nbody; b++) {
    m->body_invweight0[2 * b] = m->body_invweight0[2 * b + 1] = 0;
    if (b == 0 || m->body_weldid[b] == 0) continue;
    double J[6 * O_MAXV];
    mj_jac(m, d, J, J + 3 * nv, d->xipos + 3 * b, b);
    double A[6];
    for (int r = 0; r < 6; r++) {
      double v = 0;
      for (int a = 0; a < nv; a++)
        for (int c = 0; c < nv; c++) v += J[r * nv + a] * Minv[a * nv + c] * J[r * nv + c];
      A[r] = v;
    }
    m->body_invweight0[2 * b] = fmax(MINVAL, (A[0] + A[1] + A[2]) / 3);
    m->body_invweight0[2 * b + 1] = fmax(MINVAL, (A[3] + A[4] + A[5]) / 3);
  }
  for (int j = 0; j < m->njnt; j++) {
    int da = m->jnt_dofadr[j];
    if (m->jnt_type[j] == O_JNT_FREE) {
      double t = (Minv[da * nv + da] + Minv[(da + 1) * nv + da + 1] + Minv[(da + 2) * nv + da + 2]) / 3;
      double r = (Minv[(da + 3) * nv + da + 3] + Minv[(da + 4) * nv + da + 4] + Minv[(da + 5) * nv + da + 5]) / 3;
      for (int k = 0; k < 3; k++) { m->dof_invweight0[da + k] = t; m->dof_invweight0[da + 3 + k] = r; }
    } else {
      m->dof_invweight0[da] = Minv[da * nv + da];
    }
  }
  orc_data_free(d);
}

/* ================================================================== environment layer */
/* simulation/utils/transformations.py:1179 quaternion_matrix + :1035 euler_from_matrix(axes 'sxyz')
 * == euler_from_quaternion(q, axes=(0,0,0,1)) read back as [roll, pitch, yaw]  (actuator.py:50-56) */
static void euler_sxyz_from_mat(const double *M, double *e) {
  const double EPS = 4 * 2.220446049250313e-16;
  double cy = sqrt(M[0] * M[0] + M[3] * M[3]);
  if (cy > EPS) { e[0] = atan2(M[7], M[8]); e[1] = atan2(-M[6], cy); e[2] = atan2(M[3], M[0]); }
  else { e[0] = atan2(-M[5], M[4]); e[1] = atan2(-M[6], cy); e[2] = 0; }
}
/* transformations.py:972 euler_matrix(ai,aj,ak,'sxyz') 3x3 part */
static void euler_matrix_sxyz(double ai, double aj, double ak, double *M) {
  double si = sin(ai), sj = sin(aj), sk = sin(ak), ci = cos(ai), cj = cos(aj), ck = cos(ak);
  double cc = ci * ck, cs = ci * sk, sc = si * ck, ss = si * sk;
  M[0] = cj * ck; M[1] = sj * sc - cs; M[2] = sj * cc + ss;
  M[3] = cj * sk; M[4] = sj * ss + cc; M[5] = sj * cs - sc;
  M[6] = -sj; M[7] = cj * si; M[8] = cj * ci;
}

OEnv *orc_env_new(const OModel *m, const OEnvCfg *cfg, int body_ee, int body_object, const int *finger1, const int *finger2) {
  OEnv *e = (OEnv *)calloc(1, sizeof(OEnv));
  e->m = m;
  e->d = orc_data_new(m);
  e->cfg = *cfg;
  e->body_ee = body_ee; e->body_object = body_object;
  e->finger1_body[0] = finger1[0]; e->finger1_body[1] = finger1[1];
  e->finger2_body[0] = finger2[0]; e->finger2_body[1] = finger2[1];
  /* robot_env.py:30-33 — the direction vector is NOT normalised */
  e->target_dir[0] = 1; e->target_dir[1] = cfg->direction == 45 ? 1 : 0;
  if (cfg->direction != 0 && cfg->direction != 45) {
    /* robot_env.py:46-54 `_get_direction` (disabled upstream, :29): a unit vector at an arbitrary angle, theta rounded to 2 decimals */
    double theta = round(cfg->direction * (M_PI / 180.0) * 100.0) / 100.0;
    e->target_dir[0] = cos(theta); e->target_dir[1] = sin(theta);
  }
  return e;
}
void orc_env_free(OEnv *e) { if (e) { orc_data_free(e->d); free(e); } }

/* actuator.py:134-184 */
int orc_check_grasp(const OEnv *e) {
  const OModel *m = e->m;
  const OData *d = e->d;
  int t1 = 0, t2 = 0;
  for (int i = 0; i < d->ncon; i++) {
    int b1 = m->geom_bodyid[d->contact[i].geom1], b2 = m->geom_bodyid[d->contact[i].geom2];
    int other = -1;
    if (b1 == e->body_object) other = b2; else if (b2 == e->body_object) other = b1;
    if (other < 0) continue;
    if (other == e->finger1_body[0] || other == e->finger1_body[1]) t1 = 1;
    if (other == e->finger2_body[0] || other == e->finger2_body[1]) t2 = 1;
  }
  return t1 + 2 * t2;
}
/* utils.py:30-31 */
static double project(const double *p, const double *dir) { return (p[0] * dir[0] + p[1] * dir[1]) / (dir[0] * dir[0] + dir[1] * dir[1]); }
/* actuator.py:198-215 */
int orc_pheromone_level(const OEnv *e) {
  const double *ee = e->d->xpos + 3 * e->body_ee;
  double pr = project(ee, e->target_dir), dx = pr * e->target_dir[0] - ee[0], dy = pr * e->target_dir[1] - ee[1];
  double c = 1.0 / exp(sqrt(dx * dx + dy * dy));
  return c > 0.82 ? 3 : c > 0.6 ? 2 : c > 0.37 ? 1 : 0;
}
/* reward.py:18-41 */
double orc_agent_reward(const double *init_obj, const double *final_obj, const double *dir, int gripper_open, const double *controls, int grasped) {
  double reward = 0, ip = project(init_obj, dir), fp = project(final_obj, dir);
  double dx = fp * dir[0] - final_obj[0], dy = fp * dir[1] - final_obj[1];
  double lat = sqrt(dx * dx + dy * dy), trav = fp - ip;
  if (trav > 0 && trav < 0.1 && lat < 0.1) {
    reward = trav;
    /* `np.all(controls) != 0` : all elements non-zero */
    if (!gripper_open && (controls[0] != 0 && controls[1] != 0) && grasped == 3) {
      reward *= 2;
      if (final_obj[2] > 0) reward *= 1.5;
    }
  }
  return reward * 30;
}

/* actuator.py:21-44,58-102,249-293 */
void orc_get_target_pose(OEnv *e, const double *action, double *target_qpos) {
  const OModel *m = e->m;
  OData *d = e->d;
  const OEnvCfg *c = &e->cfg;
  double a[6];
  if (c->include_roll) memcpy(a, action, sizeof a);
  else { a[0] = action[0]; a[1] = action[1]; a[2] = action[2]; a[3] = 0; a[4] = action[3]; a[5] = action[4]; }
  double t[3] = {a[0] * c->max_translation, a[1] * c->max_translation, a[2] * c->max_translation};
  double rot[2] = {a[3] * c->max_rotation, a[4] * c->max_rotation};
  double len = norm3(t);
  if (len > c->max_translation) scl3(t, t, c->max_translation / len);
  for (int k = 0; k < 2; k++) rot[k] = rot[k] < -c->max_rotation ? -c->max_rotation : rot[k] > c->max_rotation ? c->max_rotation : rot[k];
  /* current pose */
  const double *cur_pos = d->xpos + 3 * e->body_ee, *q = d->xquat + 4 * e->body_ee;
  /* quaternion_matrix normalises by sqrt(2/|q|^2) */
  double nq = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3], qn[4], Rq[9], cur_ori[3];
  for (int k = 0; k < 4; k++) qn[k] = q[k] / sqrt(nq);
  quat2mat(Rq, qn);
  euler_sxyz_from_mat(Rq, cur_ori);
  double Rold[9], Rrel[9], Rnew[9], pos[3], ori[3];
  euler_matrix_sxyz(cur_ori[0], cur_ori[1], cur_ori[2], Rold);
  euler_matrix_sxyz(rot[0], 0.0, rot[1], Rrel);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Rnew[3 * i + j] = Rold[3 * i] * Rrel[j] + Rold[3 * i + 1] * Rrel[3 + j] + Rold[3 * i + 2] * Rrel[6 + j];
  mulmat3vec(pos, Rold, t);
  add3(pos, pos, cur_pos);
  euler_sxyz_from_mat(Rnew, ori);
  /* _enforce_constraints */
  if (!c->include_roll) ori[0] = 0;
  else { if (ori[0] > M_PI / 4) ori[0] = M_PI / 4; if (ori[0] < -M_PI / 4) ori[0] = -M_PI / 4; }
  ori[1] = 0;
  if (pos[2] < 0.1) pos[2] = 0.1;
  if (pos[2] > 0.5) pos[2] = 0.5;
  double err[6];
  sub3(err, pos, cur_pos);
  sub3(err + 3, ori, cur_ori);
  /* Jacobian of ee, first five dofs; pinv via normal equations (J has full column rank) */
  double jp[3 * O_MAXV], jr[3 * O_MAXV], J[6 * 5], A[25], b[5];
  int nv = m->nv;
  orc_jac_body(m, d, e->body_ee, jp, jr);
  for (int r = 0; r < 3; r++)
    for (int k = 0; k < 5; k++) { J[r * 5 + k] = jp[r * nv + k]; J[(3 + r) * 5 + k] = jr[r * nv + k]; }
  for (int i = 0; i < 5; i++) {
    b[i] = 0;
    for (int r = 0; r < 6; r++) b[i] += J[r * 5 + i] * err[r];
    for (int k = 0; k < 5; k++) {
      double v = 0;
      for (int r = 0; r < 6; r++) v += J[r * 5 + i] * J[r * 5 + k];
      A[i * 5 + k] = v;
    }
  }
  chol_factor(A, 5);
  chol_solve(A, 5, b);
  for (int k = 0; k < 5; k++) target_qpos[k] = d->qpos[k] + b[k];
}

static double max_abs_diff(const double *a, const double *b, int n) {
  double mx = 0;
  for (int i = 0; i < n; i++) { double v = fabs(a[i] - b[i]); if (v > mx) mx = v; }
  return mx;
}

/* robot_env.py:56-75 */
void orc_env_reset(OEnv *e, OStepOut *out) {
  orc_reset(e->m, e->d);
  e->d->xfrc_applied[6 * e->body_ee + 2] = -(0.438 * e->m->gravity[2]);
  e->episode_step = 0;
  e->status = 0;
  e->gripper_open = 1;
  if (out) {
    memset(out, 0, sizeof(*out));
    out->grasp = orc_check_grasp(e);
    out->pheromone = orc_pheromone_level(e);
    out->achieved_goal[0] = (float)e->d->xpos[3 * e->body_object];
    out->achieved_goal[1] = (float)e->d->xpos[3 * e->body_object + 1];
    out->desired_goal[0] = e->target_dir[0]; out->desired_goal[1] = e->target_dir[1];
    out->gripper_open = 1;
  }
}

/* robot_env.py:77-241 */
void orc_env_step(OEnv *e, const double *action, OStepOut *out) {
  const OModel *m = e->m;
  OData *d = e->d;
  const OEnvCfg *c = &e->cfg;
  static const double SCALE_T = 20.0; /* MinMaxScaler: 2 / (2*max_translation) with the default 0.05 */
  double scale[5];
  for (int k = 0; k < 3; k++) scale[k] = 1.0 / c->max_translation;
  for (int k = 3; k < 5; k++) scale[k] = 1.0 / c->max_rotation;
  (void)SCALE_T;
  memset(out, 0, sizeof(*out));
  double init_obj[3], init_qpos[5], target[5];
  copy3(init_obj, d->xpos + 3 * e->body_object);
  memcpy(init_qpos, d->qpos, sizeof init_qpos);
  double open_close = action[c->include_roll ? 5 : 4];
  orc_get_target_pose(e, action, target);
  memcpy(out->target_qpos, target, sizeof target);
  int step_limit = c->max_steps, reached_target = 0, reached_initial = 0;
  for (int i = 0; i < c->max_steps; i++) {
    for (int k = 0; k < 5; k++) d->ctrl[k] = (target[k] - d->qpos[k]) * scale[k];
    orc_step(m, d);
    out->nsub_a++;
    step_limit--;
    if (max_abs_diff(d->qpos, target, 5) < c->pos_tolerance) {
      reached_target = 1;
      for (int k = 0; k < 5; k++) d->ctrl[k] = 0;
      break;
    }
  }
  if (step_limit == 0) {
    memcpy(target, init_qpos, sizeof target);
    for (int i = 0; i < c->max_steps; i++) {
      for (int k = 0; k < 5; k++) d->ctrl[k] = (target[k] - d->qpos[k]) * scale[k];
      orc_step(m, d);
      out->nsub_b++;
      if (max_abs_diff(d->qpos, target, 5) < c->pos_tolerance) {
        reached_initial = 1;
        for (int k = 0; k < 5; k++) d->ctrl[k] = 0;
        break;
      }
    }
  }
  if (!reached_target && !reached_initial) { out->fail = 1; e->status = 1; }
  int object_grasped = 0;
  if (reached_target) {
    if (open_close > 0. && !e->gripper_open) {
      double tq[2] = {0.4, 0.4};
      d->ctrl[5] = d->ctrl[6] = 0.5;
      for (int i = 0; i < c->max_steps; i++) {
        double deltas = max_abs_diff(tq, d->qpos + 5, 2);
        orc_step(m, d);
        out->nsub_c++;
        if (deltas < c->grasp_tolerance || (d->qpos[5] > tq[0] && d->qpos[6] > tq[1])) {
          d->ctrl[5] = d->ctrl[6] = 0;
          e->gripper_open = 1;
          break;
        }
      }
      d->ctrl[5] = d->ctrl[6] = 0;
    } else if (open_close < 0. && e->gripper_open) {
      double tq[2] = {-0.4, -0.4};
      d->ctrl[5] = d->ctrl[6] = -1;
      for (int i = 0; i < c->max_steps; i++) {
        double deltas = max_abs_diff(tq, d->qpos + 5, 2);
        object_grasped = orc_check_grasp(e);
        orc_step(m, d);
        out->nsub_c++;
        if (deltas < c->grasp_tolerance) { d->ctrl[5] = d->ctrl[6] = 0; e->gripper_open = 0; break; }
        if (object_grasped == 3) { e->gripper_open = 0; break; }
      }
      d->ctrl[5] = d->ctrl[6] = 0;
    }
  }
  const double *fo = d->xpos + 3 * e->body_object, *fg = d->xpos + 3 * e->body_ee;
  if (sqrt((fo[0] - fg[0]) * (fo[0] - fg[0]) + (fo[1] - fg[1]) * (fo[1] - fg[1])) > 1.) e->status = 1;
  double pr = project(fo, e->target_dir);
  out->desired_goal[0] = (float)(pr * e->target_dir[0]);
  out->desired_goal[1] = (float)(pr * e->target_dir[1]);
  out->achieved_goal[0] = (float)fo[0];
  out->achieved_goal[1] = (float)fo[1];
  out->grasp = orc_check_grasp(e);
  out->pheromone = orc_pheromone_level(e);
  double reward = orc_agent_reward(init_obj, fo, e->target_dir, e->gripper_open, d->ctrl + 5, object_grasped);
  if (c->her_buffer) {
    double gx = out->desired_goal[0] - out->achieved_goal[0], gy = out->desired_goal[1] - out->achieved_goal[1];
    reward += 1.0 / exp(sqrt(gx * gx + gy * gy));
  }
  out->reward = reward;
  if (e->status != 0) out->done = 1;
  else if (e->episode_step == c->time_horizon - 1) { out->done = 1; e->status = 2; }
  else out->done = 0;
  out->total_distance = sqrt((fo[0] - init_obj[0]) * (fo[0] - init_obj[0]) + (fo[1] - init_obj[1]) * (fo[1] - init_obj[1]));
  double ip = project(init_obj, e->target_dir), dx = pr * e->target_dir[0] - fo[0], dy = pr * e->target_dir[1] - fo[1];
  double lat = sqrt(dx * dx + dy * dy), trav = pr - ip;
  out->line_distance = (trav > 0. && trav < 0.1 && lat < 0.1) ? trav : 0.;
  e->episode_step++;
  out->status = e->status;
  out->object_grasped = object_grasped;
  out->gripper_open = e->gripper_open;
  out->reached_target = reached_target;
  out->reached_initial = reached_initial;
  copy3(out->init_obj_pos, init_obj);
  copy3(out->final_obj_pos, fo);
  copy3(out->gripper_pos, fg);
  e->total_substeps += out->nsub_a + out->nsub_b + out->nsub_c;
}

/* ------------------------------------------------------------------ threaded rollout (CPU baseline) */
typedef struct {
  const OModel *m; const OEnvCfg *cfg; int body_ee, body_object; const int *f1, *f2;
  int env0, env1, nsteps; const double *actions; double reward; long substeps, transitions;
  int skip; double timed_seconds; long ncon_sum; /* skip: leading transitions per env that are rolled but not counted or timed */
} Job;
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
static void *rollout_worker(void *arg) {
  Job *j = (Job *)arg;
  for (int e = j->env0; e < j->env1; e++) {
    OEnv *env = orc_env_new(j->m, j->cfg, j->body_ee, j->body_object, j->f1, j->f2);
    OStepOut out;
    orc_env_reset(env, &out);
    long sub0 = 0;
    double t0 = now_s();
    for (int s = 0; s < j->nsteps; s++) {
      if (s == j->skip) { sub0 = env->total_substeps; t0 = now_s(); }
      orc_env_step(env, j->actions + ((size_t)e * j->nsteps + s) * 6, &out);
      if (s >= j->skip) { j->reward += out.reward; j->transitions++; j->ncon_sum += env->d->ncon; }
      if (out.done) orc_env_reset(env, NULL);
    }
    if (j->nsteps > j->skip) { j->timed_seconds += now_s() - t0; j->substeps += env->total_substeps - sub0; }
    orc_env_free(env);
  }
  return NULL;
}
/* the same with a window: every env rolls `skip` transitions from reset first (not counted, not timed), then nsteps - skip
 * counted ones.  *timed_seconds = max over threads of the time each spent inside its counted windows (all threads busy
 * throughout), *ncon_sum = sum over counted transitions of the contact count at the end of the transition. */
long orc_rollout_window(const OModel *m, const OEnvCfg *cfg, int body_ee, int body_object, const int *finger1, const int *finger2,
                        int nenv, int nsteps, int skip, const double *actions, int nthreads, double *reward_sum, long *transitions,
                        double *timed_seconds, long *ncon_sum) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > nenv) nthreads = nenv;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
  Job *jobs = (Job *)calloc(nthreads, sizeof(Job));
  for (int t = 0; t < nthreads; t++) {
    Job *j = &jobs[t];
    j->m = m; j->cfg = cfg; j->body_ee = body_ee; j->body_object = body_object; j->f1 = finger1; j->f2 = finger2;
    j->env0 = (int)((long)nenv * t / nthreads); j->env1 = (int)((long)nenv * (t + 1) / nthreads);
    j->nsteps = nsteps; j->skip = skip; j->actions = actions;
    pthread_create(&th[t], NULL, rollout_worker, j);
  }
  long sub = 0, tr = 0, nc = 0;
  double rs = 0, ts = 0;
  for (int t = 0; t < nthreads; t++) {
    pthread_join(th[t], NULL);
    sub += jobs[t].substeps; tr += jobs[t].transitions; rs += jobs[t].reward; nc += jobs[t].ncon_sum;
    if (jobs[t].timed_seconds > ts) ts = jobs[t].timed_seconds;
  }
  if (reward_sum) *reward_sum = rs;
  if (transitions) *transitions = tr;
  if (timed_seconds) *timed_seconds = ts;
  if (ncon_sum) *ncon_sum = nc;
  free(th); free(jobs);
  return sub;
}

long orc_rollout_threads(const OModel *m, const OEnvCfg *cfg, int body_ee, int body_object, const int *finger1, const int *finger2,
                         int nenv, int nsteps, const double *actions, int nthreads, double *reward_sum, long *transitions) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > nenv) nthreads = nenv;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
  Job *jobs = (Job *)calloc(nthreads, sizeof(Job));
  for (int t = 0; t < nthreads; t++) {
    Job *j = &jobs[t];
    j->m = m; j->cfg = cfg; j->body_ee = body_ee; j->body_object = body_object; j->f1 = finger1; j->f2 = finger2;
    j->env0 = (int)((long)nenv * t / nthreads); j->env1 = (int)((long)nenv * (t + 1) / nthreads);
    j->nsteps = nsteps; j->actions = actions;
    pthread_create(&th[t], NULL, rollout_worker, j);
  }
  long sub = 0, tr = 0;
  double rs = 0;
  for (int t = 0; t < nthreads; t++) { pthread_join(th[t], NULL); sub += jobs[t].substeps; tr += jobs[t].transitions; rs += jobs[t].reward; }
  if (reward_sum) *reward_sum = rs;
  if (transitions) *transitions = tr;
  free(th); free(jobs);
  return sub;
}
