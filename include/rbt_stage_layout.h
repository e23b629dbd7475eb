/*
 * rbt_stage_layout.h -- packed fp64 records of the per-stage "condense / expand / update" layer.
 *
 * These records carry what the reference keeps in
 *   ContactDynamicsData        include/robotoc/dynamics/contact_dynamics_data.hpp   (M=dIDda, J=dCda, dIDCdqv, IDC, MJtJinv, ...)
 *   StateEquationData          include/robotoc/dynamics/state_equation_data.hpp      (Fqq_prev, Fqq_inv, Fqq_prev_inv 6x6 blocks)
 *   ConstraintComponentData    include/robotoc/constraints/constraint_component_data.hpp (slack,dual,residual,cmpl,cond,dslack,ddual)
 *   SplitSolution              include/robotoc/core/split_solution.hpp
 *   SplitDirection (daf, dbetamu, dnu_passive)   src/core/split_direction.cpp:7-22
 * The linearization record is the INPUT boundary of the hot path: it is what the (out-of-scope, Pinocchio-based)
 * linearize* halves produce before condensing starts (intermediate_stage.cpp:94-132).
 *
 * All blocks are column-major with FIXED leading dimensions (nvf = nv + nf_max rows for the stacked [a;f] blocks,
 * nf_max for J/Qff, 5 for the friction-cone Jacobians), except the switching-constraint blocks which use ld = ns of
 * the stage like the KKT record.  Offsets are multiples of 2 doubles, strides multiples of 16.
 */
#ifndef RBT_STAGE_LAYOUT_H_
#define RBT_STAGE_LAYOUT_H_

#include "rbt_layout.h"

#ifdef __cplusplus
extern "C" {
#endif

#define RBT_MAX_BOX_ROWS 128
#define RBT_MAX_CONTACTS 8

/* variable a box limit acts on */
enum { RBT_VAR_Q = 0, RBT_VAR_V = 1, RBT_VAR_A = 2, RBT_VAR_U = 3 };

/* One inequality row  sign*(x[idx] - bound) + slack = 0  (sign = -1: lower limit, +1: upper limit), i.e. J = sign*e_idx.
 * Mirrors JointPosition/Velocity/Acceleration/TorquesLower/UpperLimit (src/constraints/joint_*_limit.cpp:68-90). */
typedef struct rbt_box_row {
  int var;
  int idx;  /* index inside dq / dv / da / du */
  int sign;
} rbt_box_row;

/* Constraint table shared by the batch (examples/anymal/trot.cpp:135-148: 6 box limits x 12 joints + friction cones). */
typedef struct rbt_constraint_table {
  int n_box;
  int n_contacts;           /* max point contacts; each active one contributes 5 friction-cone rows (friction_cone.cpp) */
  int impact_friction_cone; /* != 0: impact stages carry the same 5 rows per active impact (ImpactFrictionCone,
                               src/constraints/impact_friction_cone.cpp; examples/anymal/run.cpp:173-181); box limits never act
                               on impact stages (Constraints::condenseSlackAndDual(impact_status, ...), constraints.cpp:347-354) */
  int pad_;
  double barrier;           /* mu  (constraint_component_base.hpp:44) */
  double fraction_to_boundary;  /* tau (constraint_component_base.hpp:45) */
  rbt_box_row box[RBT_MAX_BOX_ROWS];
} rbt_constraint_table;

typedef struct rbt_stage_dims {
  int nv, nu, n_passive, nf_max, ns_max, n_contacts, n_box;
} rbt_stage_dims;

typedef struct rbt_stage_layout {
  int nv, nu, nx, np, nfm, nvf, nsm, ncon, nbox, nc, ncp, nq;
  /* ---- linearization record (input) */
  int l_M, l_J, l_D, l_IDC, l_Qaa, l_Qff, l_Qqf, l_Qxx, l_Quu, l_lx, l_la, l_lf, l_lu, l_Fx, l_lup, l_se3;
  int l_Phix, l_Phia, l_p, l_Phit, l_ha, l_hf, l_hx, l_hu, l_fx, l_sc, l_dgdq, l_dgdf, l_stride;
  /* ---- expansion record (written by condense, read by expand / update) */
  int e_Z, e_R, e_r, e_Qafqv, e_Qafu, e_laf, e_Qxup, e_Quup, e_lup, e_Phia, e_haf, e_Fqqpi, e_stride;
  /* ---- PDIPM record */
  int c_slack, c_dual, c_res, c_cmpl, c_cond, c_dslack, c_ddual, c_stride;
  /* ---- solution record */
  int s_q, s_v, s_a, s_dv, s_u, s_f, s_lmd, s_gmm, s_beta, s_mu, s_nup, s_xi, s_stride;
  /* ---- expanded direction record (beyond the Riccati direction record) */
  int x_daf, x_dbetamu, x_dnup, x_stride;
} rbt_stage_layout;

static inline RBT_HD void rbt_make_stage_layout(const rbt_stage_dims* d, rbt_stage_layout* L) {
  const int nv = d->nv, nu = d->nu, nx = 2 * nv, np = d->n_passive, nfm = d->nf_max, nvf = nv + nfm, nsm = d->ns_max;
  const int ncon = d->n_contacts, nbox = d->n_box, nc = nbox + 5 * ncon, ncp = rbt_up2(nc);
  int o = 0;
  L->nv = nv; L->nu = nu; L->nx = nx; L->np = np; L->nfm = nfm; L->nvf = nvf; L->nsm = nsm; L->ncon = ncon;
  L->nbox = nbox; L->nc = nc; L->ncp = ncp; L->nq = nv + (np == 6 ? 1 : 0);
  L->l_M = o; o += rbt_up2(nv * nv);
  L->l_J = o; o += rbt_up2(nfm * nv);
  L->l_D = o; o += rbt_up2(nvf * nx);
  L->l_IDC = o; o += rbt_up2(nvf);
  L->l_Qaa = o; o += rbt_up2(nv);
  L->l_Qff = o; o += rbt_up2(nfm * nfm);
  L->l_Qqf = o; o += rbt_up2(nv * nfm);
  L->l_Qxx = o; o += rbt_up2(nx * nx);
  L->l_Quu = o; o += rbt_up2(nu * nu);
  L->l_lx = o; o += rbt_up2(nx);
  L->l_la = o; o += rbt_up2(nv);
  L->l_lf = o; o += rbt_up2(nfm);
  L->l_lu = o; o += rbt_up2(nu);
  L->l_Fx = o; o += rbt_up2(nx);
  L->l_lup = o; o += rbt_up2(np);
  L->l_se3 = o; o += 3 * 36; /* Fqq top-left (dSub/dqf), Fqq_prev (dSub/dq0 at i-1), Fqq_cur (dSub/dq0 at i): 6x6 each */
  L->l_Phix = o; o += rbt_up2(nsm * nx);
  L->l_Phia = o; o += rbt_up2(nsm * nv);
  L->l_p = o; o += rbt_up2(nsm);
  L->l_Phit = o; o += rbt_up2(nsm);
  L->l_ha = o; o += rbt_up2(nv);
  L->l_hf = o; o += rbt_up2(nfm);
  L->l_hx = o; o += rbt_up2(nx);
  L->l_hu = o; o += rbt_up2(nu);
  L->l_fx = o; o += rbt_up2(nx);
  L->l_sc = o; o += 4; /* {h, Qtt, 0, 0} */
  L->l_dgdq = o; o += rbt_up2(ncon * 5 * nv);
  L->l_dgdf = o; o += rbt_up2(ncon * 15);
  L->l_stride = rbt_up16(o);

  o = 0;
  L->e_Z = o; o += rbt_up2(nvf * nvf);
  L->e_R = o; o += rbt_up2(nvf * nx);
  L->e_r = o; o += rbt_up2(nvf);
  L->e_Qafqv = o; o += rbt_up2(nvf * nx);
  L->e_Qafu = o; o += rbt_up2(nvf * nv);
  L->e_laf = o; o += rbt_up2(nvf);
  L->e_Qxup = o; o += rbt_up2(nx * np);
  L->e_Quup = o; o += rbt_up2(np * nu);
  L->e_lup = o; o += rbt_up2(np);
  L->e_Phia = o; o += rbt_up2(nsm * nv);
  L->e_haf = o; o += rbt_up2(nvf);
  L->e_Fqqpi = o; o += 36;
  L->e_stride = rbt_up16(o);

  o = 0;
  L->c_slack = o; o += ncp;
  L->c_dual = o; o += ncp;
  L->c_res = o; o += ncp;
  L->c_cmpl = o; o += ncp;
  L->c_cond = o; o += ncp;
  L->c_dslack = o; o += ncp;
  L->c_ddual = o; o += ncp;
  L->c_stride = rbt_up16(o);

  o = 0;
  L->s_q = o; o += rbt_up2(L->nq);
  L->s_v = o; o += rbt_up2(nv);
  L->s_a = o; o += rbt_up2(nv);
  L->s_dv = o; o += rbt_up2(nv);
  L->s_u = o; o += rbt_up2(nu);
  L->s_f = o; o += rbt_up2(nfm);
  L->s_lmd = o; o += rbt_up2(nv);
  L->s_gmm = o; o += rbt_up2(nv);
  L->s_beta = o; o += rbt_up2(nv);
  L->s_mu = o; o += rbt_up2(nfm);
  L->s_nup = o; o += rbt_up2(np);
  L->s_xi = o; o += rbt_up2(nsm);
  L->s_stride = rbt_up16(o);

  o = 0;
  L->x_daf = o; o += rbt_up2(nvf);
  L->x_dbetamu = o; o += rbt_up2(nvf);
  L->x_dnup = o; o += rbt_up2(np);
  L->x_stride = rbt_up16(o);
}


/* ---------------- host wire format of the linearization record (the PCIe-facing form used by rbt_iteration_host_wire)
 * Only what a robotoc linearisation can hold travels:
 *  - the symmetric blocks M, Qff, Qxx, Quu as packed upper triangles (column-major packed, element (i,j), i <= j, at
 *    j(j+1)/2 + i -- BLAS "UPLO=U" packed storage);
 *  - Qqf is NOT part of the wire record: no cost writes it, it is zero until the friction-cone condensing fills it
 *    (src/constraints/friction_cone.cpp:219 is the only writer; cost_function.cpp:194 only scales it) -- the device zero-fills it;
 *  - the STO section [ha | hf | hx | hu | fx | {h, Qtt}] only if the schedule has a switching-time stage (with_sto; the
 *    kernels read it on such stages only) -- otherwise the device zero-fills it;
 *  - neither padding nor the switching-constraint section (sent separately, only for stages that carry one).
 * A host adaptor fills it straight from the reference's Eigen members (SplitKKTMatrix::Qxx etc. are symmetric by
 * construction); the device expands it back into the rbt_stage_layout record. */
typedef struct rbt_wire_seg { int lin_off, wire_off, n, sym; } rbt_wire_seg; /* sym: n x n packed ; else n doubles copied */
typedef struct rbt_wire_zero { int lin_off, n; } rbt_wire_zero;             /* sections of the record the device zero-fills */
#define RBT_WIRE_MAX_SEGS 8
typedef struct rbt_wire_layout { int nseg, nzero, w_stride, with_sto; rbt_wire_seg seg[RBT_WIRE_MAX_SEGS]; rbt_wire_zero zero[2]; } rbt_wire_layout;

static inline RBT_HD void rbt_make_wire_layout(const rbt_stage_layout* L, int with_sto, rbt_wire_layout* W) {
  int o = 0, k = 0, n = 0;
  const int cone = L->l_dgdf + rbt_up2(15 * L->ncon) - L->l_dgdq;
  const int offs[RBT_WIRE_MAX_SEGS] = {L->l_M, L->l_J, L->l_Qff, L->l_Qxx, L->l_Quu, L->l_lx, L->l_dgdq, L->l_ha};
  const int ns[RBT_WIRE_MAX_SEGS] = {L->nv, L->l_Qff - L->l_J, L->nfm, L->nx, L->nu, L->l_Phix - L->l_lx, cone, L->l_dgdq - L->l_ha};
  const int sy[RBT_WIRE_MAX_SEGS] = {1, 0, 1, 1, 1, 0, 0, 0};
  n = with_sto ? RBT_WIRE_MAX_SEGS : RBT_WIRE_MAX_SEGS - 1;
  for (k = 0; k < RBT_WIRE_MAX_SEGS; ++k) {
    W->seg[k].lin_off = offs[k]; W->seg[k].wire_off = o; W->seg[k].n = ns[k]; W->seg[k].sym = sy[k];
    if (k < n) o += rbt_up2(sy[k] ? ns[k] * (ns[k] + 1) / 2 : ns[k]);
  }
  W->nseg = n;
  W->with_sto = with_sto ? 1 : 0;
  W->zero[0].lin_off = L->l_Qqf; W->zero[0].n = L->l_Qxx - L->l_Qqf;
  W->zero[1].lin_off = L->l_ha; W->zero[1].n = L->l_dgdq - L->l_ha;
  W->nzero = with_sto ? 1 : 2;
  W->w_stride = rbt_up2(o);
}

/* one record: rbt_stage_layout linearization record -> wire record (reads the upper triangles) */
static inline void rbt_pack_wire_record(const rbt_wire_layout* W, const double* lin, double* wire) {
  int k, i, j;
  for (k = 0; k < W->nseg; ++k) {
    const rbt_wire_seg* g = &W->seg[k];
    double* dst = wire + g->wire_off;
    const double* src = lin + g->lin_off;
    if (!g->sym) { for (i = 0; i < g->n; ++i) dst[i] = src[i]; continue; }
    for (j = 0; j < g->n; ++j)
      for (i = 0; i <= j; ++i) dst[j * (j + 1) / 2 + i] = src[i + j * g->n];
  }
}
/* wire record -> linearization record (what the device kernel does; used by the CPU tests) */
static inline void rbt_unpack_wire_record(const rbt_wire_layout* W, const double* wire, double* lin) {
  int k, i, j;
  for (k = 0; k < W->nseg; ++k) {
    const rbt_wire_seg* g = &W->seg[k];
    const double* src = wire + g->wire_off;
    double* dst = lin + g->lin_off;
    if (!g->sym) { for (i = 0; i < g->n; ++i) dst[i] = src[i]; continue; }
    for (j = 0; j < g->n; ++j)
      for (i = 0; i < g->n; ++i) dst[i + j * g->n] = (i <= j) ? src[j * (j + 1) / 2 + i] : src[i * (i + 1) / 2 + j];
  }
  for (k = 0; k < W->nzero; ++k)
    for (i = 0; i < W->zero[k].n; ++i) lin[W->zero[k].lin_off + i] = 0.0;
}

#define RBT_STAGE_LAYOUT_FIELDS(X) \
  X(nv) X(nu) X(nx) X(np) X(nfm) X(nvf) X(nsm) X(ncon) X(nbox) X(nc) X(ncp) X(nq) \
  X(l_M) X(l_J) X(l_D) X(l_IDC) X(l_Qaa) X(l_Qff) X(l_Qqf) X(l_Qxx) X(l_Quu) X(l_lx) X(l_la) X(l_lf) X(l_lu) X(l_Fx) \
  X(l_lup) X(l_se3) X(l_Phix) X(l_Phia) X(l_p) X(l_Phit) X(l_ha) X(l_hf) X(l_hx) X(l_hu) X(l_fx) X(l_sc) X(l_dgdq) \
  X(l_dgdf) X(l_stride) \
  X(e_Z) X(e_R) X(e_r) X(e_Qafqv) X(e_Qafu) X(e_laf) X(e_Qxup) X(e_Quup) X(e_lup) X(e_Phia) X(e_haf) X(e_Fqqpi) X(e_stride) \
  X(c_slack) X(c_dual) X(c_res) X(c_cmpl) X(c_cond) X(c_dslack) X(c_ddual) X(c_stride) \
  X(s_q) X(s_v) X(s_a) X(s_dv) X(s_u) X(s_f) X(s_lmd) X(s_gmm) X(s_beta) X(s_mu) X(s_nup) X(s_xi) X(s_stride) \
  X(x_daf) X(x_dbetamu) X(x_dnup) X(x_stride)

static inline int rbt_stage_layout_field(const rbt_stage_layout* L, const char* name) {
#define X(f) if (rbt_streq_(name, #f)) return L->f;
  RBT_STAGE_LAYOUT_FIELDS(X)
#undef X
  return -1;
}

#ifdef __cplusplus
}
#endif
#endif /* RBT_STAGE_LAYOUT_H_ */
