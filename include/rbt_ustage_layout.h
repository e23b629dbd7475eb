/* rbt_ustage_layout.h -- records of the stage layer of the UNCONSTRAINED path (fixed base, no contacts; iiwa14).
 *
 * What the reference keeps per stage for this path, packed [batch][N+1][record], fp64, column-major:
 *   linearization record  (inputs; produced by the linearize* halves that call Pinocchio -- out of scope)
 *       UnconstrDynamics members ID_, dID_dq_, dID_dv_, dID_da_            include/robotoc/dynamics/unconstr_dynamics.hpp
 *       SplitKKTMatrix Qxx, Qaa, Quu ; SplitKKTResidual lx, la, lu, Fx     after quadratizeStageCost + linearizeConstraints +
 *       linearizeUnconstrForwardEuler + linearizeUnconstrDynamics          src/unconstr/unconstr_intermediate_stage.cpp:71-88
 *   expansion record      lu and Quu as they stand after Constraints::condenseSlackAndDual (read by expandDual,
 *                         src/dynamics/unconstr_dynamics.cpp:98-104)
 *   PDIPM record          slack, dual, residual, cmpl, cond, dslack, ddual of every inequality row (ConstraintComponentData)
 *   solution record       SplitSolution q, v, a, u, beta, lmd, gmm
 *   expanded direction    du, dbeta (SplitDirection)
 * The KKT / Riccati / direction records of the recursion itself are rbt_ulayout (rbt_layout.h).
 * The constraint table is the rbt_constraint_table of rbt_stage_layout.h with n_contacts = 0.
 */
#ifndef RBT_USTAGE_LAYOUT_H_
#define RBT_USTAGE_LAYOUT_H_
#include "rbt_layout.h"
#include "rbt_stage_layout.h"

typedef struct rbt_ustage_layout {
  int nv, nx, nbox, ncp;
  int l_dIDdq, l_dIDdv, l_dIDda, l_ID, l_Qxx, l_Qaa, l_Quu, l_lx, l_la, l_lu, l_Fx, l_stride;
  int e_lu, e_Quu, e_stride;
  int c_slack, c_dual, c_res, c_cmpl, c_cond, c_dslack, c_ddual, c_stride;
  int s_q, s_v, s_a, s_u, s_beta, s_lmd, s_gmm, s_stride;
  int x_du, x_dbeta, x_stride;
} rbt_ustage_layout;

static inline RBT_HD void rbt_make_ustage_layout(int nv, int n_box, rbt_ustage_layout* S) {
  const int nx = 2 * nv;
  int o = 0;
  S->nv = nv; S->nx = nx; S->nbox = n_box; S->ncp = rbt_up2(n_box);
  S->l_dIDdq = o; o += rbt_up2(nv * nv);
  S->l_dIDdv = o; o += rbt_up2(nv * nv);
  S->l_dIDda = o; o += rbt_up2(nv * nv);
  S->l_ID = o; o += rbt_up2(nv);
  S->l_Qxx = o; o += rbt_up2(nx * nx);
  S->l_Qaa = o; o += rbt_up2(nv * nv);
  S->l_Quu = o; o += rbt_up2(nv * nv);
  S->l_lx = o; o += rbt_up2(nx);
  S->l_la = o; o += rbt_up2(nv);
  S->l_lu = o; o += rbt_up2(nv);
  S->l_Fx = o; o += rbt_up2(nx);
  S->l_stride = rbt_up16(o);
  o = 0;
  S->e_lu = o; o += rbt_up2(nv);
  S->e_Quu = o; o += rbt_up2(nv * nv);
  S->e_stride = rbt_up16(o);
  o = 0;
  S->c_slack = o; o += S->ncp;
  S->c_dual = o; o += S->ncp;
  S->c_res = o; o += S->ncp;
  S->c_cmpl = o; o += S->ncp;
  S->c_cond = o; o += S->ncp;
  S->c_dslack = o; o += S->ncp;
  S->c_ddual = o; o += S->ncp;
  S->c_stride = rbt_up16(o);
  o = 0;
  S->s_q = o; o += rbt_up2(nv);
  S->s_v = o; o += rbt_up2(nv);
  S->s_a = o; o += rbt_up2(nv);
  S->s_u = o; o += rbt_up2(nv);
  S->s_beta = o; o += rbt_up2(nv);
  S->s_lmd = o; o += rbt_up2(nv);
  S->s_gmm = o; o += rbt_up2(nv);
  S->s_stride = rbt_up16(o);
  o = 0;
  S->x_du = o; o += rbt_up2(nv);
  S->x_dbeta = o; o += rbt_up2(nv);
  S->x_stride = rbt_up16(o);
}

#define RBT_USTAGE_FIELDS(X) \
  X(nv) X(nx) X(nbox) X(ncp) X(l_dIDdq) X(l_dIDdv) X(l_dIDda) X(l_ID) X(l_Qxx) X(l_Qaa) X(l_Quu) X(l_lx) X(l_la) X(l_lu) \
  X(l_Fx) X(l_stride) X(e_lu) X(e_Quu) X(e_stride) X(c_slack) X(c_dual) X(c_res) X(c_cmpl) X(c_cond) X(c_dslack) \
  X(c_ddual) X(c_stride) X(s_q) X(s_v) X(s_a) X(s_u) X(s_beta) X(s_lmd) X(s_gmm) X(s_stride) X(x_du) X(x_dbeta) X(x_stride)

static inline int rbt_ustage_layout_field(const rbt_ustage_layout* S, const char* name) {
#define X(f) if (rbt_streq_(name, #f)) return S->f;
  RBT_USTAGE_FIELDS(X)
#undef X
  return -1;
}
#endif
