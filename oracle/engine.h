/* ORACLE (test infrastructure, NOT product code) — fp64, single-environment CPU restatement of the
 * reference hot path: dm_control/MuJoCo `physics.step()` for the reference scenes plus the
 * RobotEnv / Actuator / Reward logic around it.
 *
 * PARITY UNPINNED for the physics: MuJoCo is not vendored under /root/reference and cannot be
 * installed in the build container, and the reference holds no golden vectors.
 * engine.c therefore restates MuJoCo 2.2.x's published algorithms stage by stage (names of the
 * upstream functions are kept) and is anchored on the reference's own call sites.  The controller,
 * state machine and reward ARE pinned: tools/gen_golden.py runs the reference's unmodified
 * RobotEnv / Actuator / Reward python on top of this engine (through oracle/fake_dm_control.py)
 * and the resulting fixtures live in tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library.  The product path (csrc/) never links or calls it.
 */
#ifndef ORACLE_ENGINE_H
#define ORACLE_ENGINE_H

#ifdef __cplusplus
extern "C" {
#endif

#define O_MAXB 12
#define O_MAXJ 12
#define O_MAXV 16
#define O_MAXQ 16
#define O_MAXG 8
#define O_MAXM 8
#define O_MAXU 8
#define O_MAXPAIR 32
#define O_MAXCON 64
#define O_MAXEFC (O_MAXCON * 4 + 2 * O_MAXJ)

enum { O_GEOM_PLANE = 0, O_GEOM_BOX = 6, O_GEOM_MESH = 7 };
enum { O_JNT_FREE = 0, O_JNT_SLIDE = 2, O_JNT_HINGE = 3 };

typedef struct {
  int nvert;
  double *vert;  /* nvert x 3 hull vertices, mesh frame */
  int *adjadr;   /* nvert + 1 */
  int *adj;      /* neighbour lists (hull-local ids, ascending) */
} OMesh;

typedef struct {
  int nbody, njnt, nq, nv, nu, ngeom, nmesh, npair;
  double timestep, gravity[3], impratio, tolerance;
  int iterations, cone_elliptic;
  double mpr_tolerance;
  int mpr_iterations;
  /* bodies */
  int body_parentid[O_MAXB], body_weldid[O_MAXB], body_jntadr[O_MAXB], body_jntnum[O_MAXB];
  int body_dofadr[O_MAXB], body_dofnum[O_MAXB], body_rootid[O_MAXB];
  double body_pos[O_MAXB * 3], body_quat[O_MAXB * 4], body_ipos[O_MAXB * 3], body_iquat[O_MAXB * 4];
  double body_mass[O_MAXB], body_inertia[O_MAXB * 3], body_invweight0[O_MAXB * 2];
  /* joints */
  int jnt_type[O_MAXJ], jnt_bodyid[O_MAXJ], jnt_qposadr[O_MAXJ], jnt_dofadr[O_MAXJ], jnt_limited[O_MAXJ];
  double jnt_pos[O_MAXJ * 3], jnt_axis[O_MAXJ * 3], jnt_range[O_MAXJ * 2];
  double jnt_solref[2], jnt_solimp[5];
  double qpos0[O_MAXQ];
  /* dofs */
  int dof_bodyid[O_MAXV], dof_jntid[O_MAXV], dof_parentid[O_MAXV];
  double dof_armature[O_MAXV], dof_damping[O_MAXV], dof_invweight0[O_MAXV];
  /* geoms */
  int geom_type[O_MAXG], geom_bodyid[O_MAXG], geom_meshid[O_MAXG], geom_condim[O_MAXG];
  double geom_pos[O_MAXG * 3], geom_quat[O_MAXG * 4], geom_friction[O_MAXG * 3], geom_margin[O_MAXG];
  double geom_gap[O_MAXG], geom_solref[O_MAXG * 2], geom_solimp[O_MAXG * 5], geom_rbound[O_MAXG], geom_size[O_MAXG * 3];
  OMesh mesh[O_MAXM];
  /* actuators (motors) */
  int act_dofid[O_MAXU];
  double act_gear[O_MAXU], act_ctrlrange[O_MAXU * 2];
  /* collision candidates, MuJoCo contact order */
  int pair_geom1[O_MAXPAIR], pair_geom2[O_MAXPAIR];
  double meaninertia;
} OModel;

typedef struct {
  double dist, pos[3], frame[9], includemargin, friction[5], solref[2], solimp[5], mu;
  int dim, geom1, geom2, efc_address;
} OContact;

typedef struct {
  double qpos[O_MAXQ], qvel[O_MAXV], ctrl[O_MAXU], qacc[O_MAXV], qacc_warmstart[O_MAXV];
  double xfrc_applied[O_MAXB * 6];
  /* position stage */
  double xpos[O_MAXB * 3], xquat[O_MAXB * 4], xmat[O_MAXB * 9], xipos[O_MAXB * 3], ximat[O_MAXB * 9];
  double xanchor[O_MAXJ * 3], xaxis[O_MAXJ * 3];
  double geom_xpos[O_MAXG * 3], geom_xmat[O_MAXG * 9];
  double subtree_com[O_MAXB * 3];
  double cinert[O_MAXB * 10]; /* Ixx Iyy Izz Ixy Ixz Iyz, m*d[3], m  about subtree_com[root] */
  double cdof[O_MAXV * 6], cdof_dot[O_MAXV * 6], cvel[O_MAXB * 6];
  double qM[O_MAXV * O_MAXV]; /* dense, row stride nv */
  double qLD[O_MAXV * O_MAXV]; /* Cholesky factor of qM (lower) */
  int ncon;
  OContact contact[O_MAXCON];
  /* velocity / force stage */
  double qfrc_bias[O_MAXV], qfrc_passive[O_MAXV], qfrc_actuator[O_MAXV], qfrc_applied[O_MAXV];
  double qfrc_smooth[O_MAXV], qacc_smooth[O_MAXV], qfrc_constraint[O_MAXV];
  /* constraints */
  int nefc;
  int efc_type[O_MAXEFC]; /* 0 limit, 1 contact (first row), 2 contact (other rows) */
  int efc_id[O_MAXEFC];
  double efc_J[O_MAXEFC * O_MAXV], efc_pos[O_MAXEFC], efc_margin[O_MAXEFC], efc_diagApprox[O_MAXEFC];
  double efc_R[O_MAXEFC], efc_D[O_MAXEFC], efc_KBIP[O_MAXEFC * 4], efc_vel[O_MAXEFC], efc_aref[O_MAXEFC];
  double efc_force[O_MAXEFC];
  int solver_iter;
  double solver_improvement, solver_gradient;
  double time;
} OData;

/* ---- model construction (called through ctypes with string keys) ---- */
OModel *orc_model_new(void);
void orc_model_free(OModel *m);
int orc_set_d(OModel *m, const char *name, const double *v, int n);
int orc_set_i(OModel *m, const char *name, const int *v, int n);
int orc_get_d(const OModel *m, const char *name, double *out, int n); /* body_invweight0 | dof_invweight0 | meaninertia */
int orc_set_mesh(OModel *m, int id, int nvert, const double *vert, const int *adjadr, const int *adj);
void orc_model_finalize(OModel *m); /* body_rootid, invweight0, meaninertia at qpos0 */

/* ---- physics ---- */
OData *orc_data_new(const OModel *m);
void orc_data_free(OData *d);
void orc_reset(const OModel *m, OData *d);           /* mj_resetData + forward(actuation disabled) */
void orc_forward_position(const OModel *m, OData *d); /* kinematics, comPos, crb, factor, collision */
void orc_step(const OModel *m, OData *d);            /* dm_control legacy step: mj_step2 then mj_step1 */
void orc_jac_body(const OModel *m, const OData *d, int body, double *jacp, double *jacr);

/* ---- environment (RobotEnv + Actuator + Reward, reference robot_env.py / actuator.py / reward.py) ---- */
typedef struct {
  int max_steps, time_horizon, include_roll, her_buffer, direction;
  double pos_tolerance, grasp_tolerance, max_translation, max_rotation;
} OEnvCfg;

typedef struct {
  const OModel *m;
  OData *d;
  OEnvCfg cfg;
  double target_dir[2];
  int body_ee, body_object, finger1_body[2], finger2_body[2];
  int gripper_open, episode_step, status;
  long total_substeps;
} OEnv;

typedef struct {
  double reward, achieved_goal[2], desired_goal[2];
  int done, status, grasp, pheromone, object_grasped, gripper_open;
  int reached_target, reached_initial, fail;
  int nsub_a, nsub_b, nsub_c;
  double total_distance, line_distance, init_obj_pos[3], final_obj_pos[3], gripper_pos[3];
  double target_qpos[5];
} OStepOut;

OEnv *orc_env_new(const OModel *m, const OEnvCfg *cfg, int body_ee, int body_object, const int *finger1, const int *finger2);
void orc_env_free(OEnv *e);
void orc_env_reset(OEnv *e, OStepOut *out);
void orc_env_step(OEnv *e, const double *action, OStepOut *out);
void orc_get_target_pose(OEnv *e, const double *action, double *target_qpos);
int orc_check_grasp(const OEnv *e);
int orc_pheromone_level(const OEnv *e);
double orc_agent_reward(const double *init_obj, const double *final_obj, const double *dir, int gripper_open,
                        const double *controls, int grasped);

/* windowed variant: `skip` leading transitions per env are rolled but neither counted nor timed (bench.py reference arm) */
long orc_rollout_window(const OModel *m, const OEnvCfg *cfg, int body_ee, int body_object, const int *finger1, const int *finger2,
                        int nenv, int nsteps, int skip, const double *actions, int nthreads, double *reward_sum, long *transitions,
                        double *timed_seconds, long *ncon_sum);
/* rollout of many independent envs on host threads: the CPU baseline of bench.py */
long orc_rollout_threads(const OModel *m, const OEnvCfg *cfg, int body_ee, int body_object, const int *finger1,
                         const int *finger2, int nenv, int nsteps, const double *actions /* nenv*nsteps*6 */,
                         int nthreads, double *reward_sum, long *transitions);

#ifdef __cplusplus
}
#endif
#endif
