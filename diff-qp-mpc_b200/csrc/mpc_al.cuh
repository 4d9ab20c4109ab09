// mpc_al.cuh -- the augmented-Lagrangian MPC solve as ONE kernel launch (sm_100a), one warp per
// MPC problem, everything between the call's inputs and outputs resident in shared memory.
//
// Reference behaviour being re-implemented (swami1995/diff-qp-mpc, /root/reference):
//   qpth/AL_mpc.py:254-321     MPC.al_solve: warm start, al_iter x NewtonAL, lambda / rho updates
//   qpth/al_utils.py:16-34     warm_start_al
//   qpth/al_utils.py:363-460   NewtonAL.forward: 4 Newton steps on the AL merit
//   qpth/al_utils.py:62-102    merit_grad_hessian        (dense J (B,M,N), H = diag(C) + rho Jc^T Jc)
//   qpth/al_utils.py:503-527   line_search_newton        (20 step sizes 2^-k, argmin of the merit)
//   qpth/al_utils.py:37-59     merit_function
//   qpth/al_utils.py:462-500   NewtonAL.backward         -> k_al_backward
//
// What is different from the reference (same numbers, different work):
//   * H = diag(C) + rho Jc^T Jc is block TRIDIAGONAL in the knot ordering [x0 u0 x1 u1 ...]:
//       D_t   = diag(C_t) + rho ([A_t B_t]^T [A_t B_t] + E_x + diag_u(active bounds))
//       S_t   = H_{t+1,t} = -rho [A_t B_t ; 0]            (only the top nx rows are non-zero)
//     so the reference's dense (N x N) Cholesky (N = T (nx+nu)) becomes T block steps of size
//     nx+nu; J (B,M,N) and H (B,N,N) are never materialised.
//   * dynamics Jacobians come from forward-mode dual numbers in registers (mpc_dynamics.cuh).
//   * the 20 line-search candidates are evaluated by 20 lanes of the warp, each rolling the merit
//     over the horizon in registers; the argmin is a warp shuffle reduction.
#pragma once
#include "mpc_dynamics.cuh"

namespace b200mpc {

template <typename R>
struct ALArgs {
  int B, T, al_iter, newton_steps, n_ls, warm, hist_len;
  DynParams P;
  const R *x_init, *u_init, *x0, *C, *c, *u_lower, *u_upper;
  R *lam, *rho;                                        // (B,M), (B): in/out
  const R *cost_hist_in, *lam_hist_in, *rho_hist_in;   // (K,B) (K,B,M) (K,B), oldest first
  R *cost_hist_out, *lam_hist_out, *rho_hist_out;      // (al_iter+1, ...)
  R *xu_out;                                           // (B,T,nx+nu) in the solver's precision
  float *x_out, *u_out;                                // the reference returns float32 (AL_mpc.py:319-320)
  R *status;                                           // (B) status of the last line search
  R *factor;                                           // (B, factor_elems) block Cholesky of the last Hessian
  R *scratch;                                          // global scratch when the problem does not fit smem
  long long scratch_stride;
  int use_smem;
};

__host__ __device__ inline int pad4(int v) { return (v + 3) & ~3; }
__host__ __device__ inline int al_factor_elems(int T, int nx, int nu) {
  const int nt = nx + nu;
  return T * nt * nt + (T > 1 ? (T - 1) * nx * nt : 0);
}
__host__ __device__ inline long long al_scratch_elems(int T, int nx, int nu) {
  const int nt = nx + nu, M = T * nx + 2 * T * nu;
  return (long long)4 * pad4(T * nt) + pad4(M) + pad4(T * nx) + 2 * pad4(T * nu) + pad4(al_factor_elems(T, nx, nu)) +
         pad4(nx) + 32;
}

template <typename R> __device__ __forceinline__ R r_nan();
template <> __device__ __forceinline__ double r_nan<double>() { return __longlong_as_double(0x7ff8000000000000LL); }
template <> __device__ __forceinline__ float r_nan<float>() { return __int_as_float(0x7fc00000); }
template <typename R> __device__ __forceinline__ R r_inf();
template <> __device__ __forceinline__ double r_inf<double>() { return __longlong_as_double(0x7ff0000000000000LL); }
template <> __device__ __forceinline__ float r_inf<float>() { return __int_as_float(0x7f800000); }

__device__ __forceinline__ double r_rsqrt(double x) { return rsqrt(x); }
__device__ __forceinline__ float r_rsqrt(float x) { return rsqrtf(x); }

template <typename R> __device__ __forceinline__ R warp_sum(R v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename R, int NX, int NU>
struct ALScratch {
  R *xu, *C, *c, *g, *lam, *w, *ul, *uu, *D, *E, *x0, *mer;
  __device__ __forceinline__ void carve(R* q, int T) {
    constexpr int NT = NX + NU;
    auto take = [&](int n) { R* r = q; q += pad4(n); return r; };
    xu = take(T * NT); C = take(T * NT); c = take(T * NT); g = take(T * NT);
    lam = take(T * NX + 2 * T * NU); w = take(T * NX); ul = take(T * NU); uu = take(T * NU);
    D = take(al_factor_elems(T, NX, NU)); E = D + T * NT * NT;
    x0 = take(NX); mer = take(32);
  }
};

// Merit of the candidate xu + s * g (al_utils.py:37-59) rolled over the horizon by ONE lane.
// force: first state overwritten by x0 (line_search_newton, al_utils.py:515).
template <class Dyn, typename R>
__device__ __forceinline__ R merit_eval(const ALScratch<R, Dyn::NX, Dyn::NU>& S, const DynParams& P, int T, R rho,
                                        bool use_upd, R s, bool force) {
  constexpr int NX = Dyn::NX, NU = Dyn::NU, NT = NX + NU;
  const int neq = T * NX;
  R cost = R(0), pen = R(0), lin = R(0);
  R fprev[NX];
  for (int t = 0; t < T; t++) {
    R z[NT];
#pragma unroll
    for (int j = 0; j < NT; j++) {
      z[j] = S.xu[t * NT + j];
      if (use_upd) z[j] += s * S.g[t * NT + j];
    }
    if (force && t == 0) {
#pragma unroll
      for (int j = 0; j < NX; j++) z[j] = S.x0[j];
    }
    R q1 = R(0), q2 = R(0);
#pragma unroll
    for (int j = 0; j < NT; j++) { q1 += z[j] * S.C[t * NT + j] * z[j]; q2 += S.c[t * NT + j] * z[j]; }
    cost += R(0.5) * q1 + q2;
#pragma unroll
    for (int j = 0; j < NX; j++) {
      const R r = (t == 0) ? z[j] - S.x0[j] : z[j] - fprev[j];
      const R l = (t == 0) ? S.lam[(T - 1) * NX + j] : S.lam[(t - 1) * NX + j];
      pen += r * r;
      lin += l * r;
    }
#pragma unroll
    for (int j = 0; j < NU; j++) {
      const R r1 = z[NX + j] - S.uu[t * NU + j], r2 = -z[NX + j] + S.ul[t * NU + j];
      const R c1 = r1 > R(0) ? r1 : (r1 != r1 ? r1 : R(0)), c2 = r2 > R(0) ? r2 : (r2 != r2 ? r2 : R(0));
      pen += c1 * c1 + c2 * c2;
      lin += S.lam[neq + t * 2 * NU + j] * r1 + S.lam[neq + t * 2 * NU + NU + j] * r2;
    }
    if (t < T - 1) Dyn::template step<R>(P, z, z + NX, fprev);
  }
  return cost + R(0.5) * rho * pen + lin;
}

// The same merit for problems whose vectors live in the GLOBAL slab (rex): executed by the WHOLE warp.  The per-knot data
// (xu, g, C, c, the multipliers and bounds of the knot: 4 NT + NX + 4 NU reals) is fetched cooperatively -- three coalesced
// loads per lane, issued one knot ahead -- into a double-buffered shared-memory stage and read from there by the candidate
// lanes (`active`); with every candidate lane walking the slab on its own each knot cost a chain of L2 round trips
// (ncu: 20 % of the samples of the rex solve, nearly all long-scoreboard).  Same operations in the same order.
template <class Dyn, typename R>
__device__ __forceinline__ R merit_eval_staged(const ALScratch<R, Dyn::NX, Dyn::NU>& S, const DynParams& P, int T, R rho,
                                               bool use_upd, R s, bool force, bool active, int lane, R* stage) {
  constexpr int NX = Dyn::NX, NU = Dyn::NU, NT = NX + NU;
  constexpr int O_G = NT, O_C = 2 * NT, O_c = 3 * NT, O_LX = 4 * NT, O_LU = O_LX + NX, O_UU = O_LU + 2 * NU, O_UL = O_UU + NU,
                CNT = O_UL + NU, STG = (CNT + 3) & ~3, NLD = (CNT + 31) / 32;
  const int neq = T * NX;
  auto fetch = [&](int t, R (&r)[NLD]) {
#pragma unroll
    for (int k = 0; k < NLD; k++) {
      const int idx = lane + 32 * k;
      R v = R(0);
      if (idx < O_G) v = S.xu[t * NT + idx];
      else if (idx < O_C) v = S.g[t * NT + idx - O_G];
      else if (idx < O_c) v = S.C[t * NT + idx - O_C];
      else if (idx < O_LX) v = S.c[t * NT + idx - O_c];
      else if (idx < O_LU) v = S.lam[(t == 0 ? T - 1 : t - 1) * NX + idx - O_LX];
      else if (idx < O_UU) v = S.lam[neq + t * 2 * NU + idx - O_LU];
      else if (idx < O_UL) v = S.uu[t * NU + idx - O_UU];
      else if (idx < CNT) v = S.ul[t * NU + idx - O_UL];
      r[k] = v;
    }
  };
  auto put = [&](R* buf, const R (&r)[NLD]) {
#pragma unroll
    for (int k = 0; k < NLD; k++) { const int idx = lane + 32 * k; if (idx < CNT) buf[idx] = r[k]; }
  };
  R x0v[NX];
#pragma unroll
  for (int j = 0; j < NX; j++) x0v[j] = S.x0[j];
  R pre[NLD];
  fetch(0, pre);
  put(stage, pre);
  __syncwarp();
  R cost = R(0), pen = R(0), lin = R(0);
  R fprev[NX];
  for (int t = 0; t < T; t++) {
    const R* b = stage + (t & 1) * STG;
    if (t + 1 < T) fetch(t + 1, pre);
    if (active) {
      R z[NT];
#pragma unroll
      for (int j = 0; j < NT; j++) {
        z[j] = b[j];
        if (use_upd) z[j] += s * b[O_G + j];
      }
      if (force && t == 0) {
#pragma unroll
        for (int j = 0; j < NX; j++) z[j] = x0v[j];
      }
      R q1 = R(0), q2 = R(0);
#pragma unroll
      for (int j = 0; j < NT; j++) { q1 += z[j] * b[O_C + j] * z[j]; q2 += b[O_c + j] * z[j]; }
      cost += R(0.5) * q1 + q2;
#pragma unroll
      for (int j = 0; j < NX; j++) {
        const R r = (t == 0) ? z[j] - x0v[j] : z[j] - fprev[j];
        const R l = b[O_LX + j];
        pen += r * r;
        lin += l * r;
      }
#pragma unroll
      for (int j = 0; j < NU; j++) {
        const R r1 = z[NX + j] - b[O_UU + j], r2 = -z[NX + j] + b[O_UL + j];
        const R c1 = r1 > R(0) ? r1 : (r1 != r1 ? r1 : R(0)), c2 = r2 > R(0) ? r2 : (r2 != r2 ? r2 : R(0));
        pen += c1 * c1 + c2 * c2;
        lin += b[O_LU + j] * r1 + b[O_LU + NU + j] * r2;
      }
      if (t < T - 1) Dyn::template step<R>(P, z, z + NX, fprev);
    }
    if (t + 1 < T) put(stage + ((t + 1) & 1) * STG, pre);
    __syncwarp();
  }
  return cost + R(0.5) * rho * pen + lin;
}

// Gradient (into S.g) and block-tridiagonal Hessian (S.D lower blocks, S.E = H_{t+1,t} top rows)
// of the AL merit at S.xu (al_utils.py:62-102).
// TILE16 (NT = 16, fp64: block_cholesky_tiles16 follows): the Jacobian [A_t B_t] is ALSO written transposed into D_t's slot
// (16 x 16, columns 12..15 zero); the Hessian blocks D_t = diag + rho J^T J are then formed as DMMA tile products inside the
// factorisation (J^T J = JT JT^T is of the A B^T form), E is scaled by -rho as it is loaded there, and the two scalar
// passes below are skipped.
template <class Dyn, typename R, bool TILE16 = false>
__device__ __forceinline__ void assemble(const ALScratch<R, Dyn::NX, Dyn::NU>& S, const DynParams& P, int T, R rho, int lane) {
  constexpr int NX = Dyn::NX, NU = Dyn::NU, NT = NX + NU;
  const int neq = T * NX;
  // Jacobians by forward-mode duals, DW directions at a time (bounds the register footprint)
  // (two at a time for the my_envs cart-pole and the quadrotor: measured 5 % faster than four -- fewer live
  // registers / less local memory outweigh the extra passes; four elsewhere)
  constexpr int DWMAX = (NX == 4 || NX >= 12) ? 2 : 4;
  constexpr int DW = NT <= DWMAX ? NT : DWMAX;
  typedef Dual<R, DW> DR;
  // one task = one knot and one group of DW directions, tasks dealt round-robin to the lanes: with "lane = knot" a
  // horizon of 40 knots kept 8 lanes busy for a second round of all NT / DW passes while 24 idled
  constexpr int NPASS = (NT + DW - 1) / DW;
  for (int task = lane; task < (T - 1) * NPASS; task += 32) {
    const int t = task / NPASS, c0 = (task - t * NPASS) * DW;
    {
      {
        DR z[NT], f[NX];
#pragma unroll
        for (int j = 0; j < NT; j++) {
          z[j] = DR(S.xu[t * NT + j]);
#pragma unroll
          for (int q = 0; q < DW; q++) z[j].d[q] = (j == c0 + q) ? R(1) : R(0);
        }
        Dyn::template step<DR>(P, z, z + NX, f);
#pragma unroll
        for (int i = 0; i < NX; i++) {
#pragma unroll
          for (int q = 0; q < DW; q++)
            if (c0 + q < NT) {
              S.E[(t * NX + i) * NT + c0 + q] = f[i].d[q];
              if (TILE16) S.D[t * NT * NT + (c0 + q) * NT + i] = f[i].d[q];
            }
          if (c0 == 0) S.w[t * NX + i] = S.lam[t * NX + i] + rho * (S.xu[(t + 1) * NT + i] - f[i].v);
        }
        if (TILE16) {
#pragma unroll
          for (int q = 0; q < DW; q++)
            if (c0 + q < NT) {
#pragma unroll
              for (int i = NX; i < NT; i++) S.D[t * NT * NT + (c0 + q) * NT + i] = R(0);
            }
        }
      }
    }
  }
  if (lane < NX) S.w[(T - 1) * NX + lane] = S.lam[(T - 1) * NX + lane] + rho * (S.xu[lane] - S.x0[lane]);
  __syncwarp();
  for (int idx = lane; idx < T * NT; idx += 32) {
    const int t = idx / NT, j = idx - t * NT;
    R gv = S.C[idx] * S.xu[idx] + S.c[idx];
    if (t < T - 1) {
      R acc = R(0);
#pragma unroll
      for (int i = 0; i < NX; i++) acc += S.E[(t * NX + i) * NT + j] * S.w[t * NX + i];
      gv -= acc;
    }
    if (j < NX) {
      gv += (t >= 1) ? S.w[(t - 1) * NX + j] : S.w[(T - 1) * NX + j];
    } else {
      const int ju = j - NX;
      const R u = S.xu[idx];
      const R r1 = u - S.uu[t * NU + ju], r2 = -u + S.ul[t * NU + ju];
      const R c1 = r1 > R(0) ? r1 : (r1 != r1 ? r1 : R(0)), c2 = r2 > R(0) ? r2 : (r2 != r2 ? r2 : R(0));
      gv += S.lam[neq + t * 2 * NU + ju] - S.lam[neq + t * 2 * NU + NU + ju] + rho * (c1 - c2);
    }
    S.g[idx] = gv;
  }
  if constexpr (!TILE16) {
  for (int idx = lane; idx < T * NT * NT; idx += 32) {
    const int t = idx / (NT * NT), rem = idx - t * NT * NT, j = rem / NT, k = rem - j * NT;
    if (k > j) continue;
    R acc = R(0);
    if (t < T - 1) {
#pragma unroll
      for (int i = 0; i < NX; i++) acc += S.E[(t * NX + i) * NT + j] * S.E[(t * NX + i) * NT + k];
    }
    if (j == k) {
      if (j < NX) acc += R(1);
      else {
        const int ju = j - NX;
        const R u = S.xu[t * NT + j];
        if (u - S.uu[t * NU + ju] > R(0)) acc += R(1);
        if (-u + S.ul[t * NU + ju] > R(0)) acc += R(1);
      }
    }
    S.D[idx] = (j == k ? S.C[t * NT + j] : R(0)) + rho * acc;
  }
  __syncwarp();
  for (int idx = lane; idx < (T - 1) * NX * NT; idx += 32) S.E[idx] = -rho * S.E[idx];
  }
  __syncwarp();
}

// In-place block Cholesky: D_t <- L_tt (lower), E_t <- L_{t+1,t} top rows = S_t L_tt^-T.
// A non-positive pivot poisons the factor with NaN (the reference's cholesky_ex would hand back
// an unusable factor and fall back to a dense LU solve for the WHOLE batch, al_utils.py:419-427).
// rd: NT reals of scratch (the reciprocal diagonal of the block being factored, for the row solves of E_t).
template <int NX, int NU, typename R>
__device__ __forceinline__ void block_cholesky(R* D, R* E, int T, int lane, R* rd) {
  constexpr int NT = NX + NU;
  for (int t = 0; t < T; t++) {
    R* Dt = D + t * NT * NT;
    if (t > 0) {
      const R* Ep = E + (t - 1) * NX * NT;
      for (int idx = lane; idx < NX * NX; idx += 32) {
        const int i = idx / NX, k = idx - i * NX;
        if (k <= i) {
          R acc = R(0);
#pragma unroll
          for (int j = 0; j < NT; j++) acc += Ep[i * NT + j] * Ep[k * NT + j];
          Dt[i * NT + k] -= acc;
        }
      }
      __syncwarp();
    }
#pragma unroll 1
    for (int j = 0; j < NT; j++) {
      const R djj = Dt[j * NT + j];
      // 1 / sqrt(d) first (MUFU seed + Newton: ~12 instructions), then l = d * (1 / sqrt(d)): a double sqrt followed by a
      // double division is ~55 instructions on the pivot chain.  d <= 0 gives NaN as before.
      const R rjj = djj > R(0) ? r_rsqrt(djj) : r_nan<R>();
      const R ljj = djj * rjj;
      __syncwarp();
      if (lane >= j && lane < NT) Dt[lane * NT + j] = (lane == j) ? ljj : Dt[lane * NT + j] * rjj;
      if (lane == 0) rd[j] = rjj;
      __syncwarp();
      for (int idx = lane; idx < NT * NT; idx += 32) {
        const int i = idx / NT, k = idx - i * NT;
        if (k > j && k <= i) Dt[idx] -= Dt[i * NT + j] * Dt[k * NT + j];
      }
      __syncwarp();
    }
    if (t < T - 1) {
      if (lane < NX) {
        R* row = E + (t * NX + lane) * NT;
#pragma unroll 1
        for (int j = 0; j < NT; j++) {
          R v = row[j];
          for (int k = 0; k < j; k++) v -= row[k] * Dt[j * NT + k];
          row[j] = v * rd[j];
        }
      }
      __syncwarp();
    }
  }
}

// The same factorisation for problems whose state lives in the GLOBAL scratch slab (rex quadrotor, T = 40: 175 KB per
// problem): the block being factored is staged in shared memory (5 KB per warp: D_t, the factored E_{t-1} and E_t, leading
// dimension NT + 1 so that row and column walks are bank-conflict free).  In the slab every step of the column loop was a
// dependent global-memory round trip (~10^3 cycles x 16 columns x T blocks x 8 Newton steps: ~20 of the 21 ms of a rex
// solve); the arithmetic and its order are unchanged.
template <int NX, int NU>
__host__ __device__ constexpr int al_stage_elems() {
  // the larger of: the staged block factorisation (D_t, E_{t-1}, E_t) and 32 + the double-buffered merit stage
  constexpr int a = ((NX + NU) * (NX + NU + 1) + 2 * NX * (NX + NU + 1) + 3) & ~3;
  constexpr int b = 32 + 2 * ((4 * (NX + NU) + NX + 4 * NU + 3) & ~3);
  return a > b ? a : b;
}

template <int NX, int NU, typename R>
__device__ __forceinline__ void block_cholesky_staged(R* D, R* E, int T, int lane, R* wk) {
  constexpr int NT = NX + NU, LD = NT + 1;
  R* sD = wk;
  R* sEp = wk + NT * LD;
  R* sEn = sEp + NX * LD;
  for (int t = 0; t < T; t++) {
    R* Dt = D + t * NT * NT;
    for (int idx = lane; idx < NT * NT; idx += 32) sD[(idx / NT) * LD + (idx % NT)] = Dt[idx];
    if (t < T - 1) {
      const R* Eg = E + t * NX * NT;
      for (int idx = lane; idx < NX * NT; idx += 32) sEn[(idx / NT) * LD + (idx % NT)] = Eg[idx];
    }
    __syncwarp();
    if (t > 0) {
      for (int idx = lane; idx < NX * NX; idx += 32) {
        const int i = idx / NX, k = idx - i * NX;
        if (k <= i) {
          R acc = R(0);
#pragma unroll
          for (int j = 0; j < NT; j++) acc += sEp[i * LD + j] * sEp[k * LD + j];
          sD[i * LD + k] -= acc;
        }
      }
      __syncwarp();
    }
#pragma unroll 1
    for (int j = 0; j < NT; j++) {
      const R djj = sD[j * LD + j];
      const R ljj = djj > R(0) ? sqrt(djj) : r_nan<R>();
      __syncwarp();
      if (lane >= j && lane < NT) sD[lane * LD + j] = (lane == j) ? ljj : sD[lane * LD + j] / ljj;
      __syncwarp();
      for (int idx = lane; idx < NT * NT; idx += 32) {
        const int i = idx / NT, k = idx - i * NT;
        if (k > j && k <= i) sD[i * LD + k] -= sD[i * LD + j] * sD[k * LD + j];
      }
      __syncwarp();
    }
    if (t < T - 1) {
      if (lane < NX) {
        R* row = sEn + lane * LD;
#pragma unroll 1
        for (int j = 0; j < NT; j++) {
          R v = row[j];
          for (int k = 0; k < j; k++) v -= row[k] * sD[j * LD + k];
          row[j] = v / sD[j * LD + j];
        }
      }
      __syncwarp();
      R* Eg = E + t * NX * NT;
      for (int idx = lane; idx < NX * NT; idx += 32) Eg[idx] = sEn[(idx / NT) * LD + (idx % NT)];
      R* tmp = sEp; sEp = sEn; sEn = tmp;
    }
    for (int idx = lane; idx < NT * NT; idx += 32) Dt[idx] = sD[(idx / NT) * LD + (idx % NT)];
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The block factorisation for NT = nx + nu = 16 (rex quadrotor: nx = 12, nu = 4) on the FP64 TENSOR CORES, one warp per
// problem, the blocks in registers as 8x8 DMMA accumulator tiles (lane (g, q) = (lane >> 2, lane & 3) holds row g, columns
// 2q, 2q+1 of a tile) -- the tile algebra of qp_wres.cuh:
//   * D_t (16x16) = [[A00, .], [A10, A11]] is factored as L Delta L^T: each 8x8 diagonal tile by quad shuffles, carrying
//     W = L_JJ^-1 along (td_diag), X10 = A10 W0^T and A11 -= X10 Delta0^-1 X10^T as DMMAs;
//   * E_t (12x16, two row tiles) becomes Y = E L^-T = [E0 W0^T, (E1 - Y0 L10^T) W1^T]: every product has the form A B^T,
//     which mma.sync.m8n8k4.f64 computes from two accumulator-layout operands without moving data;
//   * the Schur complement of the next block, D_{t+1}[0:12, 0:12] -= Y Delta^-1 Y^T, uses Y straight from registers.
// 28 DMMAs and two 8-step pivot chains per knot; the results are written back in the CHOLESKY convention the block solves
// and k_al_backward read (L_chol = L Delta^1/2, E_chol = Y Delta^-1/2), so nothing else changes.  A non-positive pivot
// gives a NaN factor, as before.  (ncu before: 37 % of the samples of k_al_solve<RexQuadrotor> in the scalar block factor.)
__device__ __forceinline__ void td_dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void td_abt(double& c0, double& c1, double a0, double a1, double b0, double b1) {  // C += A B^T
  td_dmma(c0, c1, a0, b0);
  td_dmma(c0, c1, a1, b1);
}
__device__ __forceinline__ double td_shfl(double v, int src, int width = 32) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(0xffffffffu, lo, src, width);
  hi = __shfl_sync(0xffffffffu, hi, src, width);
  return __hiloint2double(hi, lo);
}
// LDL^T of one 8x8 diagonal tile (lower triangle valid): on exit column k < g of row g holds L[g][k] Delta_k, (w0, w1) = L^-1,
// rinv[0..8) (shared memory) the reciprocal pivots
__device__ __forceinline__ void td_diag(double& c0, double& c1, double& w0, double& w1, double* rinv, int lane) {
  const int g = lane >> 2, q = lane & 3, q8 = 8 * q;
  w0 = (g == 2 * q) ? 1.0 : 0.0;
  w1 = (g == 2 * q + 1) ? 1.0 : 0.0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const int kq = k >> 1;
    const double ck = (k & 1) ? c1 : c0;
    const double dk = td_shfl(ck, 4 * k + kq);
    const double ckm = (g > k) ? ck : 0.0;
    const double cik = td_shfl(ckm, kq, 4);
    const double cj0 = td_shfl(ckm, q8 + kq);
    const double cj1 = td_shfl(ckm, q8 + 4 + kq);
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(dk));
    const double e = fma(-dk, r, 1.0);
    const double t = fma(e, e, e);
    r = fma(r, t, r);
    if (lane == 0) rinv[k] = r;
    const double nl = -(cik * r);
    c0 = fma(nl, cj0, c0);
    c1 = fma(nl, cj1, c1);
    if (k < 7) {
      const double wk0 = td_shfl(w0, q + 4 * k), wk1 = td_shfl(w1, q + 4 * k);
      w0 = fma(nl, wk0, w0);
      w1 = fma(nl, wk1, w1);
    }
  }
}

// On entry (assemble<.., TILE16 = true>): D_t's slot holds J_t^T = [A_t B_t]^T (16 x 16, columns 12..15 zero; nothing for the
// last knot, which has no dynamics), E_t the UNSCALED J_t (12 x 16).  Cd, xu (T x 16), uu, ul (T x 4): the diagonal terms of
// al_utils.py:62-102.  rinv: 16 doubles of shared memory.
__device__ __forceinline__ void block_cholesky_tiles16(double* D, double* E, int T, int lane, double* rinv, double rho,
                                                       const double* Cd, const double* xu, const double* uu, const double* ul) {
  constexpr int NT = 16, NX = 12, NU = 4;
  const int g = lane >> 2, q = lane & 3, c = 2 * q;
  double Y[2][2][2];  // Y of the previous knot: [row tile][column tile][slot]
  double yr[2][2];    // its reciprocal pivots of this lane's columns, per column tile
  for (int t = 0; t < T; t++) {
    double* Dt = D + t * NT * NT;
    // ---- D_t = diag(C_t) + rho (J^T J + E_x + active control bounds), J^T J = JT JT^T as tile products
    double a00[2] = {0.0, 0.0}, a10[2] = {0.0, 0.0}, a11[2] = {0.0, 0.0};
    double e[2][2][2];
    if (t < T - 1) {
      double jt[2][2][2];  // JT tiles: [row tile of tau][column tile of the next state]
#pragma unroll
      for (int r = 0; r < 2; r++)
#pragma unroll
        for (int k = 0; k < 2; k++) {
          const double2 v = *reinterpret_cast<const double2*>(Dt + (8 * r + g) * NT + 8 * k + c);
          jt[r][k][0] = v.x; jt[r][k][1] = v.y;
        }
#pragma unroll
      for (int k = 0; k < 2; k++) {
        td_abt(a00[0], a00[1], jt[0][k][0], jt[0][k][1], jt[0][k][0], jt[0][k][1]);
        td_abt(a10[0], a10[1], jt[1][k][0], jt[1][k][1], jt[0][k][0], jt[0][k][1]);
        td_abt(a11[0], a11[1], jt[1][k][0], jt[1][k][1], jt[1][k][0], jt[1][k][1]);
      }
      const double* Eg = E + t * NX * NT;
      const double nr = -rho;
#pragma unroll
      for (int r = 0; r < 2; r++)
#pragma unroll
        for (int k = 0; k < 2; k++) {
          double2 v = make_double2(0.0, 0.0);
          if (8 * r + g < NX) v = *reinterpret_cast<const double2*>(Eg + (8 * r + g) * NT + 8 * k + c);
          e[r][k][0] = nr * v.x; e[r][k][1] = nr * v.y;   // H_{t+1,t} = -rho [A_t B_t]
        }
    }
    {
      // diagonal terms, on the lanes that hold a diagonal entry (c == g or c + 1 == g)
      const bool on0 = (c == g), on1 = (c + 1 == g);
      double ex0 = 0.0, ex1 = 0.0, cd0 = 0.0, cd1 = 0.0;
      if (on0 || on1) {
        cd0 = Cd[t * NT + g]; cd1 = Cd[t * NT + 8 + g];
        ex0 = 1.0;                                     // rows 0..7 are states
        if (8 + g < NX) ex1 = 1.0;
        else {
          const int ju = 8 + g - NX;
          const double u = xu[t * NT + 8 + g];
          if (u - uu[t * NU + ju] > 0.0) ex1 += 1.0;
          if (-u + ul[t * NU + ju] > 0.0) ex1 += 1.0;
        }
      }
      a00[0] = rho * (a00[0] + (on0 ? ex0 : 0.0)) + (on0 ? cd0 : 0.0);
      a00[1] = rho * (a00[1] + (on1 ? ex0 : 0.0)) + (on1 ? cd0 : 0.0);
      a10[0] = rho * a10[0]; a10[1] = rho * a10[1];
      a11[0] = rho * (a11[0] + (on0 ? ex1 : 0.0)) + (on0 ? cd1 : 0.0);
      a11[1] = rho * (a11[1] + (on1 ? ex1 : 0.0)) + (on1 ? cd1 : 0.0);
    }
    // ---- Schur complement of the previous knot: D_t[0:12, 0:12] -= Y Delta^-1 Y^T (rows 12..15 of Y are zero)
    if (t > 0) {
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const double s0 = -yr[k][0], s1 = -yr[k][1];
        td_abt(a00[0], a00[1], Y[0][k][0] * s0, Y[0][k][1] * s1, Y[0][k][0], Y[0][k][1]);
        td_abt(a10[0], a10[1], Y[1][k][0] * s0, Y[1][k][1] * s1, Y[0][k][0], Y[0][k][1]);
        td_abt(a11[0], a11[1], Y[1][k][0] * s0, Y[1][k][1] * s1, Y[1][k][0], Y[1][k][1]);
      }
    }
    // ---- L Delta L^T of the 16x16 block
    double w0[2], w1[2];
    td_diag(a00[0], a00[1], w0[0], w0[1], rinv, lane);
    __syncwarp();
    const double2 r0 = *reinterpret_cast<const double2*>(rinv + c);
    double x10[2] = {0.0, 0.0};
    td_abt(x10[0], x10[1], a10[0], a10[1], w0[0], w0[1]);                     // X10 = A10 W0^T = L10 Delta0
    td_abt(a11[0], a11[1], x10[0] * -r0.x, x10[1] * -r0.y, x10[0], x10[1]);   // A11 -= X10 Delta0^-1 X10^T
    td_diag(a11[0], a11[1], w1[0], w1[1], rinv + 8, lane);
    __syncwarp();
    const double2 r1 = *reinterpret_cast<const double2*>(rinv + 8 + c);
    const double sr0x = sqrt(r0.x), sr0y = sqrt(r0.y), sr1x = sqrt(r1.x), sr1y = sqrt(r1.y);  // Delta^-1/2 of this lane's columns
    // ---- Y = E L^-T, written back as E_chol = Y Delta^-1/2
    if (t < T - 1) {
      double* Eg = E + t * NX * NT;
#pragma unroll
      for (int r = 0; r < 2; r++) {
        double y0[2] = {0.0, 0.0};
        td_abt(y0[0], y0[1], e[r][0][0], e[r][0][1], w0[0], w0[1]);           // Y0 = E0 W0^T
        double tt[2] = {0.0, 0.0};
        td_abt(tt[0], tt[1], y0[0], y0[1], x10[0] * r0.x, x10[1] * r0.y);     // Y0 L10^T, L10 = X10 Delta0^-1
        tt[0] = e[r][1][0] - tt[0]; tt[1] = e[r][1][1] - tt[1];
        double y1[2] = {0.0, 0.0};
        td_abt(y1[0], y1[1], tt[0], tt[1], w1[0], w1[1]);                     // Y1 = (E1 - Y0 L10^T) W1^T
        Y[r][0][0] = y0[0]; Y[r][0][1] = y0[1]; Y[r][1][0] = y1[0]; Y[r][1][1] = y1[1];
        if (8 * r + g < NX) {
          *reinterpret_cast<double2*>(Eg + (8 * r + g) * NT + c) = make_double2(y0[0] * sr0x, y0[1] * sr0y);
          *reinterpret_cast<double2*>(Eg + (8 * r + g) * NT + 8 + c) = make_double2(y1[0] * sr1x, y1[1] * sr1y);
        }
      }
      yr[0][0] = r0.x; yr[0][1] = r0.y; yr[1][0] = r1.x; yr[1][1] = r1.y;
    }
    // ---- L_chol = L Delta^1/2 (lower triangle; the diagonal is Delta^1/2 = 1 / sqrt(rinv))
    {
      double2 o0, o1, o2;
      o0.x = (c < g) ? a00[0] * sr0x : (c == g ? 1.0 / sr0x : 0.0);
      o0.y = (c + 1 < g) ? a00[1] * sr0y : (c + 1 == g ? 1.0 / sr0y : 0.0);
      o1.x = x10[0] * sr0x; o1.y = x10[1] * sr0y;
      o2.x = (c < g) ? a11[0] * sr1x : (c == g ? 1.0 / sr1x : 0.0);
      o2.y = (c + 1 < g) ? a11[1] * sr1y : (c + 1 == g ? 1.0 / sr1y : 0.0);
      *reinterpret_cast<double2*>(Dt + g * NT + c) = o0;
      *reinterpret_cast<double2*>(Dt + (8 + g) * NT + c) = o1;
      *reinterpret_cast<double2*>(Dt + (8 + g) * NT + 8 + c) = o2;
    }
    __syncwarp();
  }
}

// The same tile factorisation for SMALL knot blocks (NT = nx + nu <= 8: pendulum, integrator, the cart-poles), fp64: D_t
// is ONE 8x8 accumulator tile padded by the identity, E_t (nx x NT) one tile padded by zeros.  Per knot: one Schur product,
// one 8-step pivot chain, one product for Y = E W^T (6 DMMAs) instead of NT rounds of (sqrt, divide, warp barrier, rank-1
// update, warp barrier) on shared memory -- ncu had 36 % of the samples of k_al_solve<Cartpole1L> in the scalar block factor.
// D (assembled by `assemble`, lower triangle) and E (already scaled by -rho) are read and written in place, in the Cholesky
// convention (L_chol = L Delta^1/2, E_chol = Y Delta^-1/2).
template <int NX, int NU>
__device__ __forceinline__ void block_cholesky_tiles8(double* D, double* E, int T, int lane, double* rinv) {
  constexpr int NT = NX + NU;
  static_assert(NT <= 8, "one tile per knot block");
  const int g = lane >> 2, q = lane & 3, c = 2 * q;
  double Y[2] = {0.0, 0.0}, yr[2] = {1.0, 1.0};
  for (int t = 0; t < T; t++) {
    double* Dt = D + t * NT * NT;
    double a[2];
    a[0] = (g < NT && c <= g) ? Dt[g * NT + c] : ((g >= NT && c == g) ? 1.0 : 0.0);
    a[1] = (g < NT && c + 1 <= g) ? Dt[g * NT + c + 1] : ((g >= NT && c + 1 == g) ? 1.0 : 0.0);
    double e[2] = {0.0, 0.0};
    if (t < T - 1) {
      const double* Eg = E + t * NX * NT;
      if (g < NX && c < NT) e[0] = Eg[g * NT + c];
      if (g < NX && c + 1 < NT) e[1] = Eg[g * NT + c + 1];
    }
    if (t > 0) td_abt(a[0], a[1], Y[0] * -yr[0], Y[1] * -yr[1], Y[0], Y[1]);   // D_t[0:nx, 0:nx] -= Y Delta^-1 Y^T
    double w[2];
    td_diag(a[0], a[1], w[0], w[1], rinv, lane);
    __syncwarp();
    const double2 r = *reinterpret_cast<const double2*>(rinv + c);
    const double srx = sqrt(r.x), sry = sqrt(r.y);
    if (t < T - 1) {
      double y[2] = {0.0, 0.0};
      td_abt(y[0], y[1], e[0], e[1], w[0], w[1]);                               // Y = E W^T
      Y[0] = y[0]; Y[1] = y[1]; yr[0] = r.x; yr[1] = r.y;
      double* Eg = E + t * NX * NT;
      if (g < NX && c < NT) Eg[g * NT + c] = y[0] * srx;
      if (g < NX && c + 1 < NT) Eg[g * NT + c + 1] = y[1] * sry;
    }
    if (g < NT) {
      if (c <= g) Dt[g * NT + c] = (c < g) ? a[0] * srx : 1.0 / srx;
      if (c + 1 <= g) Dt[g * NT + c + 1] = (c + 1 < g) ? a[1] * sry : 1.0 / sry;
    }
    __syncwarp();
  }
}

// g <- H^-1 g with the block factor (two block-bidiagonal sweeps; in-block solves by shuffles).
template <int NX, int NU, typename R>
__device__ __forceinline__ void block_solve(const R* D, const R* E, R* g, int T, int lane) {
  constexpr int NT = NX + NU;
  static_assert(NT <= 32, "one lane per row of a knot block");
  // Software-pipelined over the knots: the factor row / column, the E row / column and the right-hand side of knot t+1 are
  // loaded while knot t is substituted, and the previous knot's solution is broadcast by shuffles instead of being
  // re-read from memory -- with the state in the global slab every knot was a chain of dependent L2 round trips
  // (ncu: 20 % of the samples of the rex solve).  Same operations in the same order.
  const bool act = lane < NT;
  // ---- forward: L y = g
  {
    R rowv[NT], erow[NT];
#pragma unroll
    for (int j = 0; j < NT; j++) { rowv[j] = (act && j <= lane) ? D[lane * NT + j] : R(0); erow[j] = R(0); }
    R gv = act ? g[lane] : R(0);
    R vprev = R(0);
    for (int t = 0; t < T; t++) {
      R rowv_n[NT], erow_n[NT], gv_n = R(0);
      if (t + 1 < T) {
        const R* Dn = D + (t + 1) * NT * NT;
        const R* En = E + (t * NX + lane) * NT;
#pragma unroll
        for (int j = 0; j < NT; j++) {
          rowv_n[j] = (act && j <= lane) ? Dn[lane * NT + j] : R(0);
          erow_n[j] = (lane < NX) ? En[j] : R(0);
        }
        if (act) gv_n = g[(t + 1) * NT + lane];
      } else {
#pragma unroll
        for (int j = 0; j < NT; j++) { rowv_n[j] = R(0); erow_n[j] = R(0); }
      }
      R v = gv;
      if (t > 0) {
        R acc = R(0);
#pragma unroll
        for (int j = 0; j < NT; j++) acc += erow[j] * __shfl_sync(0xffffffffu, vprev, j);
        if (lane < NX) v -= acc;
      }
      R dg = R(1);
#pragma unroll
      for (int j = 0; j < NT; j++) dg = (lane == j) ? rowv[j] : dg;
      dg = R(1) / dg;  // one division per lane and knot instead of NT (a double division is ~25 instructions)
#pragma unroll
      for (int j = 0; j < NT; j++) {
        const R vj = __shfl_sync(0xffffffffu, v, j) * __shfl_sync(0xffffffffu, dg, j);
        if (lane == j) v = vj;
        else if (lane > j && act) v -= rowv[j] * vj;
      }
      if (act) g[t * NT + lane] = v;
      vprev = v;
      gv = gv_n;
#pragma unroll
      for (int j = 0; j < NT; j++) { rowv[j] = rowv_n[j]; erow[j] = erow_n[j]; }
    }
  }
  // ---- backward: L^T x = y
  {
    R colv[NT], ecol[NX];
    {
      const R* Dl = D + (T - 1) * NT * NT;
#pragma unroll
      for (int j = 0; j < NT; j++) colv[j] = (act && j >= lane) ? Dl[j * NT + lane] : R(0);
#pragma unroll
      for (int i = 0; i < NX; i++) ecol[i] = R(0);
    }
    R gv = act ? g[(T - 1) * NT + lane] : R(0);
    R vnext = R(0);
    for (int t = T - 1; t >= 0; t--) {
      R colv_n[NT], ecol_n[NX], gv_n = R(0);
      if (t > 0) {
        const R* Dn = D + (t - 1) * NT * NT;
        const R* En = E + (t - 1) * NX * NT;
#pragma unroll
        for (int j = 0; j < NT; j++) colv_n[j] = (act && j >= lane) ? Dn[j * NT + lane] : R(0);
#pragma unroll
        for (int i = 0; i < NX; i++) ecol_n[i] = act ? En[i * NT + lane] : R(0);
        if (act) gv_n = g[(t - 1) * NT + lane];
      } else {
#pragma unroll
        for (int j = 0; j < NT; j++) colv_n[j] = R(0);
#pragma unroll
        for (int i = 0; i < NX; i++) ecol_n[i] = R(0);
      }
      R v = gv;
      if (t < T - 1) {
        R acc = R(0);
#pragma unroll
        for (int i = 0; i < NX; i++) acc += ecol[i] * __shfl_sync(0xffffffffu, vnext, i);
        if (act) v -= acc;
      }
      R dg = R(1);
#pragma unroll
      for (int j = 0; j < NT; j++) dg = (lane == j) ? colv[j] : dg;
      dg = R(1) / dg;
#pragma unroll
      for (int j = NT - 1; j >= 0; j--) {
        const R vj = __shfl_sync(0xffffffffu, v, j) * __shfl_sync(0xffffffffu, dg, j);
        if (lane == j) v = vj;
        else if (lane < j) v -= colv[j] * vj;
      }
      if (act) g[t * NT + lane] = v;
      vnext = v;
      gv = gv_n;
#pragma unroll
      for (int j = 0; j < NT; j++) colv[j] = colv_n[j];
#pragma unroll
      for (int i = 0; i < NX; i++) ecol[i] = ecol_n[i];
    }
  }
  __syncwarp();
}

template <class Dyn, typename R>
// Occupancy: the cart-pole of deqmpc/my_envs (NX = 4, RK4 under 4-wide duals) compiles to 252 registers = 2 CTAs/SM;
// capped at 128 (4 CTAs/SM, ~0.8 KB of spills) it runs 45 % faster (5.04 -> 3.48 ms at T=20, B=4096).  The two-link
// model (NX = 6) loses 5-7 % under the same cap and the other environments already sit at <= 166 registers.
#ifndef B200MPC_REX_MINB
#define B200MPC_REX_MINB 0
#endif
__global__ void __launch_bounds__(128, (Dyn::NX == 4 ? 4 : (Dyn::NX >= 12 ? B200MPC_REX_MINB : 0))) k_al_solve(const ALArgs<R> a) {
  constexpr int NX = Dyn::NX, NU = Dyn::NU, NT = NX + NU;
  extern __shared__ __align__(16) unsigned char al_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int prob = blockIdx.x * (blockDim.x >> 5) + warp;
  if (prob >= a.B) return;  // warps never meet at a CTA barrier
  const int T = a.T, M = T * NX + 2 * T * NU, neq = T * NX;
  ALScratch<R, NX, NU> S;
  S.carve(a.use_smem ? reinterpret_cast<R*>(al_smem) + (size_t)warp * a.scratch_stride
                     : a.scratch + (size_t)prob * a.scratch_stride, T);
  const int fe = al_factor_elems(T, NX, NU);
  if (!a.use_smem) {  // global mode: factor straight into the output buffer of the call (no copy at the end)
    S.D = a.factor + (size_t)prob * fe;
    S.E = S.D + T * NT * NT;
  }
  const size_t B = (size_t)a.B;

  // ---- load the problem
  for (int idx = lane; idx < T * NT; idx += 32) {
    const int t = idx / NT, j = idx - t * NT;
    S.xu[idx] = j < NX ? a.x_init[((size_t)prob * T + t) * NX + j] : a.u_init[((size_t)prob * T + t) * NU + (j - NX)];
    S.C[idx] = a.C[(size_t)prob * T * NT + idx];
    S.c[idx] = a.c[(size_t)prob * T * NT + idx];
  }
  for (int idx = lane; idx < T * NU; idx += 32) {
    S.ul[idx] = a.u_lower[(size_t)prob * T * NU + idx];
    S.uu[idx] = a.u_upper[(size_t)prob * T * NU + idx];
  }
  for (int idx = lane; idx < M; idx += 32) S.lam[idx] = a.lam[(size_t)prob * M + idx];
  if (lane < NX) S.x0[lane] = a.x0[(size_t)prob * NX + lane];
  R rho = a.rho[prob];
  __syncwarp();

  // ---- cost at the start, warm start from the previous call's history (AL_mpc.py:268-278)
  R cost_start;
  {
    R acc = R(0);
    for (int idx = lane; idx < T * NT; idx += 32) acc += R(0.5) * S.xu[idx] * S.C[idx] * S.xu[idx] + S.c[idx] * S.xu[idx];
    cost_start = warp_sum(acc);
  }
  if (a.warm && a.hist_len > 0) {
    int pick = a.hist_len - 1;  // torch.max over an all-False column returns index 0 = newest
    for (int k = a.hist_len - 1; k >= 0; k--) {
      if (a.cost_hist_in[(size_t)k * B + prob] < cost_start) { pick = k; break; }
    }
    const R* lh = a.lam_hist_in + ((size_t)pick * B + prob) * M;
    R nh = R(0), nl = R(0);
    for (int idx = lane; idx < M; idx += 32) { nh += lh[idx] * lh[idx]; nl += S.lam[idx] * S.lam[idx]; }
    nh = warp_sum(nh); nl = warp_sum(nl);
    const R scale = sqrt(nh) / sqrt(nl);
    for (int idx = lane; idx < M; idx += 32) S.lam[idx] *= scale;
    rho = a.rho_hist_in[(size_t)pick * B + prob];
    __syncwarp();
  }
  if (lane == 0) { a.cost_hist_out[prob] = cost_start; a.rho_hist_out[prob] = rho; }
  for (int idx = lane; idx < M; idx += 32) a.lam_hist_out[(size_t)prob * M + idx] = S.lam[idx];

  R status = R(0);
  for (int ai = 0; ai < a.al_iter; ai++) {
    // ---- NewtonAL.forward (al_utils.py:363-460)
    R* mstage = reinterpret_cast<R*>(al_smem) + (size_t)warp * al_stage_elems<NX, NU>() + 32;  // global mode only
    R merit = a.use_smem ? merit_eval<Dyn, R>(S, a.P, T, rho, false, R(0), false)
                         : merit_eval_staged<Dyn, R>(S, a.P, T, rho, false, R(0), false, true, lane, mstage);
    for (int ns = 0; ns < a.newton_steps; ns++) {
      constexpr bool kTile16 = NX == 12 && NU == 4 && sizeof(R) == 8;
      assemble<Dyn, R, kTile16>(S, a.P, T, rho, lane);
      if constexpr (kTile16) {
        // 16x16 knot blocks as DMMA tiles in registers (a.use_smem or not: the blocks are read and written once per knot)
        double* rinv = a.use_smem ? reinterpret_cast<double*>(S.mer) : reinterpret_cast<double*>(al_smem) + (size_t)warp * al_stage_elems<NX, NU>();
        block_cholesky_tiles16(reinterpret_cast<double*>(S.D), reinterpret_cast<double*>(S.E), T, lane, rinv, (double)rho,
                               reinterpret_cast<const double*>(S.C), reinterpret_cast<const double*>(S.xu),
                               reinterpret_cast<const double*>(S.uu), reinterpret_cast<const double*>(S.ul));
      } else if constexpr (NT <= 8 && sizeof(R) == 8) {
        // small knot blocks as one DMMA tile each (shared-memory or global-slab state alike)
        double* rinv = a.use_smem ? reinterpret_cast<double*>(S.mer) : reinterpret_cast<double*>(al_smem) + (size_t)warp * al_stage_elems<NX, NU>();
        block_cholesky_tiles8<NX, NU>(reinterpret_cast<double*>(S.D), reinterpret_cast<double*>(S.E), T, lane, rinv);
      } else if (a.use_smem) block_cholesky<NX, NU, R>(S.D, S.E, T, lane, S.mer);
      else block_cholesky_staged<NX, NU, R>(S.D, S.E, T, lane, reinterpret_cast<R*>(al_smem) + (size_t)warp * al_stage_elems<NX, NU>());
      block_solve<NX, NU, R>(S.D, S.E, S.g, T, lane);
      for (int idx = lane; idx < T * NT; idx += 32) S.g[idx] = -S.g[idx];
      __syncwarp();
      // line search: lane k evaluates step 2^-k (al_utils.py:503-527)
      R mv = r_inf<R>();
      R step = R(1);
      for (int k = 0; k < lane; k++) step *= R(0.5);
      if (a.use_smem) {
        if (lane < a.n_ls) mv = merit_eval<Dyn, R>(S, a.P, T, rho, true, step, true);
      } else {
        const R mm = merit_eval_staged<Dyn, R>(S, a.P, T, rho, true, step, true, lane < a.n_ls, lane, mstage);
        if (lane < a.n_ls) mv = mm;
      }
      // argmin, first index on ties, NaN wins (torch.min propagates NaN)
      R bv = mv;
      int bi = lane;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const R ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        const bool bnan = bv != bv, onan = ov != ov;
        bool take;
        if (bnan || onan) take = onan && (!bnan || oi < bi);
        else take = ov < bv || (ov == bv && oi < bi);
        if (take) { bv = ov; bi = oi; }
      }
      const R bstep = __shfl_sync(0xffffffffu, step, bi);
      status = (bv < merit) ? R(1) : R(0);
      for (int idx = lane; idx < T * NT; idx += 32) {
        const int t = idx / NT, j = idx - t * NT;
        R xn = S.xu[idx] + bstep * S.g[idx];
        if (t == 0 && j < NX) xn = S.x0[j];
        S.xu[idx] = status * xn + (R(1) - status) * S.xu[idx];
      }
      merit = bv;
      __syncwarp();
    }
    // ---- multiplier / penalty update (AL_mpc.py:298-310)
    for (int t = lane; t < T; t += 32) {
      if (t < T - 1) {
        R f[NX];
        Dyn::template step<R>(a.P, S.xu + t * NT, S.xu + t * NT + NX, f);
#pragma unroll
        for (int i = 0; i < NX; i++) S.lam[t * NX + i] += rho * (S.xu[(t + 1) * NT + i] - f[i]);
      }
      if (t == 0) {
#pragma unroll
        for (int i = 0; i < NX; i++) S.lam[(T - 1) * NX + i] += rho * (S.xu[i] - S.x0[i]);
      }
#pragma unroll
      for (int j = 0; j < NU; j++) {
        const R u = S.xu[t * NT + NX + j];
        const R l1 = S.lam[neq + t * 2 * NU + j] + rho * (u - S.uu[t * NU + j]);
        const R l2 = S.lam[neq + t * 2 * NU + NU + j] + rho * (-u + S.ul[t * NU + j]);
        S.lam[neq + t * 2 * NU + j] = l1 < R(0) ? R(0) : l1;       // torch.clamp(min=0): NaN stays NaN
        S.lam[neq + t * 2 * NU + NU + j] = l2 < R(0) ? R(0) : l2;
      }
    }
    __syncwarp();
    R acc = R(0);
    for (int idx = lane; idx < T * NT; idx += 32) acc += R(0.5) * S.xu[idx] * S.C[idx] * S.xu[idx] + S.c[idx] * S.xu[idx];
    acc = warp_sum(acc);
    rho = rho * R(10);
    if (lane == 0) { a.cost_hist_out[(size_t)(ai + 1) * B + prob] = acc; a.rho_hist_out[(size_t)(ai + 1) * B + prob] = rho; }
    for (int idx = lane; idx < M; idx += 32) a.lam_hist_out[((size_t)(ai + 1) * B + prob) * M + idx] = S.lam[idx];
  }

  // ---- outputs + the state the reference keeps on the module
  for (int idx = lane; idx < M; idx += 32) a.lam[(size_t)prob * M + idx] = S.lam[idx];
  if (lane == 0) { a.rho[prob] = rho; a.status[prob] = status; }
  for (int idx = lane; idx < T * NT; idx += 32) {
    const int t = idx / NT, j = idx - t * NT;
    const R v = S.xu[idx];
    a.xu_out[(size_t)prob * T * NT + idx] = v;
    if (j < NX) a.x_out[((size_t)prob * T + t) * NX + j] = (float)v;
    else a.u_out[((size_t)prob * T + t) * NU + (j - NX)] = (float)v;
  }
  if (a.use_smem)
    for (int idx = lane; idx < fe; idx += 32) a.factor[(size_t)prob * fe + idx] = S.D[idx];
}

// NewtonAL.backward (al_utils.py:462-500): ig = -H^-1 grad ; dC = ig * x_est ; dc = ig.
template <typename R>
struct ALBackArgs {
  int B, T;
  const R *factor, *xu, *grad;
  R *dC, *dc;
};

template <int NX, int NU, typename R>
__global__ void __launch_bounds__(128) k_al_backward(const ALBackArgs<R> a) {
  constexpr int NT = NX + NU;
  extern __shared__ __align__(16) unsigned char al_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int prob = blockIdx.x * (blockDim.x >> 5) + warp;
  if (prob >= a.B) return;
  const int T = a.T;
  R* g = reinterpret_cast<R*>(al_smem) + (size_t)warp * pad4(T * NT);
  const int fe = al_factor_elems(T, NX, NU);
  const R* D = a.factor + (size_t)prob * fe;
  const R* E = D + T * NT * NT;
  for (int idx = lane; idx < T * NT; idx += 32) g[idx] = a.grad[(size_t)prob * T * NT + idx];
  __syncwarp();
  block_solve<NX, NU, R>(D, E, g, T, lane);
  for (int idx = lane; idx < T * NT; idx += 32) {
    const R ig = -g[idx];
    a.dc[(size_t)prob * T * NT + idx] = ig;
    a.dC[(size_t)prob * T * NT + idx] = ig * a.xu[(size_t)prob * T * NT + idx];
  }
}

// Stand-alone batched dynamics: step (xn) and Jacobians (A (N,nx,nx), B (N,nx,nu)), one thread per
// knot (deqmpc/envs.py:16-31,68-82; qpth/env_dx/*.py forward).
template <class Dyn, typename R>
__global__ void k_dyn_step(DynParams P, const R* x, const R* u, R* xn, R* A, R* Bm, long long N) {
  constexpr int NX = Dyn::NX, NU = Dyn::NU, NT = NX + NU;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  if (A == nullptr) {
    R z[NT], f[NX];
#pragma unroll
    for (int j = 0; j < NX; j++) z[j] = x[i * NX + j];
#pragma unroll
    for (int j = 0; j < NU; j++) z[NX + j] = u[i * NU + j];
    Dyn::template step<R>(P, z, z + NX, f);
#pragma unroll
    for (int j = 0; j < NX; j++) xn[i * NX + j] = f[j];
  } else {
    constexpr int DW = NT <= 4 ? NT : 4;
    typedef Dual<R, DW> DR;
#pragma unroll 1
    for (int c0 = 0; c0 < NT; c0 += DW) {
      DR z[NT], f[NX];
#pragma unroll
      for (int j = 0; j < NT; j++) {
        z[j] = DR(j < NX ? x[i * NX + j] : u[i * NU + (j - NX)]);
#pragma unroll
        for (int q = 0; q < DW; q++) z[j].d[q] = (j == c0 + q) ? R(1) : R(0);
      }
      Dyn::template step<DR>(P, z, z + NX, f);
#pragma unroll
      for (int r = 0; r < NX; r++) {
        if (c0 == 0) xn[i * NX + r] = f[r].v;
#pragma unroll
        for (int q = 0; q < DW; q++) {
          const int j = c0 + q;
          if (j < NX) A[(i * NX + r) * NX + j] = f[r].d[q];
          else if (j < NT) Bm[(i * NX + r) * NU + (j - NX)] = f[r].d[q];
        }
      }
    }
  }
}

// Open-loop rollout x_{t+1} = f(x_t, u_t), t < T-1 (AL_mpc.py:398-411), one thread per trajectory.
template <class Dyn, typename R>
__global__ void k_dyn_rollout(DynParams P, const R* x0, const R* u, R* xs, long long B, int T) {
  constexpr int NX = Dyn::NX, NU = Dyn::NU;
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  R x[NX], uu[NU], xn[NX];
#pragma unroll
  for (int j = 0; j < NX; j++) { x[j] = x0[b * NX + j]; xs[(b * T) * NX + j] = x[j]; }
  for (int t = 0; t + 1 < T; t++) {
#pragma unroll
    for (int j = 0; j < NU; j++) uu[j] = u[(b * T + t) * NU + j];
    Dyn::template step<R>(P, x, uu, xn);
#pragma unroll
    for (int j = 0; j < NX; j++) { x[j] = xn[j]; xs[(b * T + t + 1) * NX + j] = xn[j]; }
  }
}

}  // namespace b200mpc
