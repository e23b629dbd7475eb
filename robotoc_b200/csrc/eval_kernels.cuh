// eval_kernels.cuh -- the small stage-parallel pieces a device-resident SQP loop needs besides the linear algebra
// (SURVEY.md 8a rows a12 / a13 / a16):
//   perf_index_kernel + perf_reduce_kernel   PerformanceIndex of evalKKT, summed over the horizon
//       {Intermediate,Impact,Terminal}Stage::evalKKT summaries   src/ocp/intermediate_stage.cpp:128-132, impact_stage.cpp:109-113,
//       terminal_stage.cpp:97-100;  SplitKKTResidual::KKTError   include/robotoc/core/split_kkt_residual.hxx:90-104;
//       ConstraintComponentData::KKTError / logBarrier           include/robotoc/constraints/constraint_component_data.hpp:122,
//       pdipm.hxx:194-200;  sum over stages                      src/ocp/direct_multiple_shooting.cpp:155-158
//   slack_dual_positive_kernel               pdipm::setSlackAndDualPositive   include/robotoc/constraints/pdipm.hxx:13-24
//   initial_state_direction_kernel           computeInitialStateDirection     src/dynamics/state_equation.cpp:98-109
// All HBM-streaming, a few hundred doubles per stage.
#pragma once
#include "stage_kernels.cuh"

namespace rbt {

struct EvalParams {
  StageParams sp;
  double* stage_perf;  // [batch][n_grid][4]: {cost_barrier, primal_feasibility, dual_feasibility, kkt_error} of every stage
  double* perf;        // [batch][8]: {cost (0), cost_barrier, primal_feas, dual_feas, kkt_error, sqrt(kkt_error), 0, 0}
  const double* x0in;  // [batch][2 nv]: {q0 (-) s0.q, v0}
  double* dx0;         // [batch][nx]
};

// One warp per (OCP, stage).  Every lane accumulates a strided share of each vector; the four sums are combined by a
// fixed shuffle tree, so the result is reproducible run to run.
__global__ void __launch_bounds__(128) perf_index_kernel(const EvalParams q) {
  const StageParams& p = q.sp;
  const rbt_stage_layout& S = p.S;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= p.batch * p.n_grid) return;
  const int i = w % p.n_grid;
  const rbt_stage_ctrl c = p.ctrl[i];
  const double* lin = p.lin + size_t(w) * S.l_stride;
  const double* con = p.con + size_t(w) * S.c_stride;
  double kkt = 0.0, pf = 0.0, df = 0.0, lb = 0.0;
  auto acc_p = [&](const double* v, int n) {
    for (int e = lane; e < n; e += 32) {
      const double x = v[e];
      kkt = fma(x, x, kkt);
      pf += fabs(x);
    }
  };
  auto acc_d = [&](const double* v, int n) {
    for (int e = lane; e < n; e += 32) {
      const double x = v[e];
      kkt = fma(x, x, kkt);
      df += fabs(x);
    }
  };
  acc_d(lin + S.l_lx, S.nx);
  if (c.type != RBT_TERMINAL) {
    const bool impact = c.type == RBT_IMPACT;
    acc_p(lin + S.l_Fx, S.nx);
    acc_d(lin + S.l_la, S.nv);  // la | ldv
    acc_d(lin + S.l_lf, c.nf);
    acc_p(lin + S.l_IDC, S.nv + c.nf);
    if (!impact) {
      acc_d(lin + S.l_lu, S.nu);
      acc_d(lin + S.l_lup, S.np);
      acc_p(lin + S.l_p, c.ns);
    }
    if (!impact || p.tab.impact_friction_cone != 0) {  // impact stages: the ImpactFrictionCone rows only
      const double mu = p.tab.barrier;
      for (int r = lane + (impact ? S.nbox : 0); r < S.nc; r += 32) {
        const bool cone = r >= S.nbox;
        if (cone && !((c.contact_mask >> ((r - S.nbox) / 5)) & 1)) continue;
        if (!cone && __ldg(p.row_level + r) + c.ineq_gate > 2) continue;  // level not valid on this grid point
        const double sl = con[S.c_slack + r], du = con[S.c_dual + r], res = con[S.c_res + r];
        const double cm = sl * du - mu;
        kkt = fma(res, res, kkt);
        kkt = fma(cm, cm, kkt);
        pf += fabs(res);
        df += fabs(cm);
        lb -= mu * log(sl);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    kkt += __shfl_xor_sync(0xffffffffu, kkt, o);
    pf += __shfl_xor_sync(0xffffffffu, pf, o);
    df += __shfl_xor_sync(0xffffffffu, df, o);
    lb += __shfl_xor_sync(0xffffffffu, lb, o);
  }
  if (lane == 0) {
    double* out = q.stage_perf + size_t(w) * 4;
    out[0] = lb; out[1] = pf; out[2] = df; out[3] = kkt;
  }
}

// One thread per OCP: the horizon sum in stage order (deterministic), and KKTError() = sqrt(kkt_error).
__global__ void perf_reduce_kernel(const EvalParams q) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= q.sp.batch) return;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  const double* s = q.stage_perf + size_t(b) * q.sp.n_grid * 4;
  for (int i = 0; i < q.sp.n_grid; ++i) {
    a0 += s[4 * i + 0]; a1 += s[4 * i + 1]; a2 += s[4 * i + 2]; a3 += s[4 * i + 3];
  }
  double* o = q.perf + size_t(b) * 8;
  o[0] = 0.0; o[1] = a0; o[2] = a1; o[3] = a2; o[4] = a3; o[5] = sqrt(a3); o[6] = 0.0; o[7] = 0.0;
}

// slack <- max(slack, sqrt(barrier)); dual <- barrier / slack on every inequality row of every constrained stage.
__global__ void slack_dual_positive_kernel(const EvalParams q) {
  const StageParams& p = q.sp;
  const rbt_stage_layout& S = p.S;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)p.batch * p.n_grid * S.ncp;
  if (e >= total) return;
  const int r = int(e % S.ncp);
  const long long st = e / S.ncp;
  const int i = int(st % p.n_grid);
  const int type = p.ctrl[i].type;
  if (r >= S.nc || type == RBT_TERMINAL || (type == RBT_IMPACT && (p.tab.impact_friction_cone == 0 || r < S.nbox))) return;
  if (r < S.nbox && __ldg(p.row_level + r) + p.ctrl[i].ineq_gate > 2) return;  // level not valid on this grid point
  double* con = p.con + size_t(st) * S.c_stride;
  const double sb = sqrt(p.tab.barrier);
  const double sl = fmax(con[S.c_slack + r], sb);
  con[S.c_slack + r] = sl;
  con[S.c_dual + r] = p.tab.barrier / sl;
}

// SURVEY.md 8f-2, first slice: Constraints::linearizeConstraints of the joint-limit components (constraints.cpp:283-306,
// joint_*_limit.cpp:47-63) -- residual = sign (x - bound) + slack into the PDIPM record, l_x += sign dual into the gradient of
// the linearization record, for the rows whose level is valid on the grid point.  One thread per (grid point, target entry
// (var, idx)): the lower and the upper limit of an entry meet on one gradient entry and are added in table order, like the
// reference's component loop (bit-reproducible).
__global__ void linearize_joint_limits_kernel(const EvalParams q, const double* __restrict__ bound) {
  const StageParams& p = q.sp;
  const rbt_stage_layout& S = p.S;
  const int nt = 3 * S.nv + S.nu;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)p.batch * p.n_grid * nt) return;
  const int t = int(e % nt);
  const long long st = e / nt;
  const rbt_stage_ctrl c = p.ctrl[int(st % p.n_grid)];
  if (c.type == RBT_TERMINAL || c.type == RBT_IMPACT) return;
  const int4 e4 = __ldg(p.tgt + t);
  if (e4.x == 0) return;  // no limit acts on this entry
  const int var = t < 3 * S.nv ? t / S.nv : RBT_VAR_U, idx = t < 3 * S.nv ? t % S.nv : t - 3 * S.nv;
  const double* s = p.sol + size_t(st) * S.s_stride;
  double* l = const_cast<double*>(p.lin) + size_t(st) * S.l_stride;
  double* con = p.con + size_t(st) * S.c_stride;
  double x, *g;
  switch (var) {
    case RBT_VAR_Q: x = s[S.s_q + idx + (S.np == 6 ? 1 : 0)]; g = l + S.l_lx + idx; break;  // q has one more entry (quaternion)
    case RBT_VAR_V: x = s[S.s_v + idx]; g = l + S.l_lx + S.nv + idx; break;
    case RBT_VAR_A: x = s[S.s_a + idx]; g = l + S.l_la + idx; break;
    default: x = s[S.s_u + idx]; g = l + S.l_lu + idx; break;
  }
  const int ent[4] = {e4.x, e4.y, e4.z, e4.w};
  double grad = *g;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (ent[k] == 0) continue;
    const int r = (ent[k] < 0 ? -ent[k] : ent[k]) - 1;
    if (__ldg(p.row_level + r) + c.ineq_gate > 2) continue;
    const double sg = ent[k] < 0 ? -1.0 : 1.0;
    con[S.c_res + r] = sg * (x - bound[r]) + con[S.c_slack + r];
    grad += sg * con[S.c_dual + r];
  }
  *g = grad;
}

// dx0 = [ -Fqq_prev_inv (q0 (-) q)[0:6] | (q0 (-) q)[6:] | v0 - v ]   (Fqq_prev_inv of stage 0 is left by the condensing kernel)
__global__ void initial_state_direction_kernel(const EvalParams q) {
  const StageParams& p = q.sp;
  const rbt_stage_layout& S = p.S;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= p.batch * S.nx) return;
  const int b = e / S.nx, k = e % S.nx, nv = S.nv;
  const double* in = q.x0in + size_t(b) * 2 * nv;
  const double* ex0 = p.ex + size_t(b) * p.n_grid * S.e_stride;
  const double* sol0 = p.sol + size_t(b) * p.n_grid * S.s_stride;
  double v;
  if (k < nv) {
    if (S.np == 6 && k < 6) {
      double a = 0.0;
      for (int l = 0; l < 6; ++l) a = fma(ex0[S.e_Fqqpi + k + 6 * l], in[l], a);
      v = -a;
    } else {
      v = in[k];
    }
  } else {
    v = in[k] - sol0[S.s_v + (k - nv)];  // in[nv + j] = v0[j]
  }
  q.dx0[size_t(b) * S.nx + k] = v;
}

}  // namespace rbt
