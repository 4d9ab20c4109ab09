// b200mpc.cu -- host side of the MPC entry points of libb200qp.so (include/b200mpc.h): problem
// validation, shared-memory sizing, environment dispatch.  No torch types anywhere.
#include <cuda_runtime.h>
#include <string.h>

#include "../../include/b200mpc.h"
#include "mpc_al.cuh"

namespace b200qp { int cuda_fail(cudaError_t e, const char* what); }

namespace b200mpc {

#define CKM(call)                                                 \
  do {                                                            \
    cudaError_t e_ = (call);                                      \
    if (e_ != cudaSuccess) return b200qp::cuda_fail(e_, #call);   \
  } while (0)

constexpr int kWarpsPerCta = 4;
constexpr size_t kSmemLimit = 200 * 1024;

static bool env_dims(int env, int& nx, int& nu) {
  switch (env) {
    case ENV_PENDULUM: nx = Pendulum::NX; nu = Pendulum::NU; return true;
    case ENV_INTEGRATOR: nx = Integrator::NX; nu = Integrator::NU; return true;
    case ENV_PENDULUM_DX: nx = PendulumDx::NX; nu = PendulumDx::NU; return true;
    case ENV_CARTPOLE_DX: nx = CartpoleDx::NX; nu = CartpoleDx::NU; return true;
    case ENV_REX_QUADROTOR: nx = RexQuadrotor::NX; nu = RexQuadrotor::NU; return true;
    case ENV_PENDULUM1L: nx = Pendulum1L::NX; nu = Pendulum1L::NU; return true;
    case ENV_CARTPOLE1L: nx = Cartpole1L::NX; nu = Cartpole1L::NU; return true;
    case ENV_CARTPOLE2L: nx = Cartpole2L::NX; nu = Cartpole2L::NU; return true;
  }
  return false;
}

static int check(const b200mpc_problem_t* p, int& nx, int& nu) {
  if (!p || p->B < 1 || p->T < 2) return B200QP_EINVAL;
  if (p->dtype != B200QP_F64 && p->dtype != B200QP_F32) return B200QP_EINVAL;
  if (!env_dims(p->env, nx, nu)) return B200QP_EINVAL;
  if (p->al_iter < 1 || p->al_iter > 64 || p->newton_steps < 1 || p->n_ls < 1 || p->n_ls > 32) return B200QP_EINVAL;
  if (p->warm && p->hist_len < 1) return B200QP_EINVAL;
  return B200QP_OK;
}

template <class Dyn, typename R>
static int solve_t(const b200mpc_problem_t* p, const b200mpc_buffers_t* b, cudaStream_t st) {
  ALArgs<R> a;
  memset(&a, 0, sizeof(a));
  a.B = p->B; a.T = p->T; a.al_iter = p->al_iter; a.newton_steps = p->newton_steps; a.n_ls = p->n_ls;
  a.warm = p->warm; a.hist_len = p->hist_len;
  for (int i = 0; i < MAX_PARAMS; i++) a.P.v[i] = p->params[i];
  a.x_init = (const R*)b->x_init; a.u_init = (const R*)b->u_init; a.x0 = (const R*)b->x0;
  a.C = (const R*)b->C; a.c = (const R*)b->c; a.u_lower = (const R*)b->u_lower; a.u_upper = (const R*)b->u_upper;
  a.lam = (R*)b->lam; a.rho = (R*)b->rho;
  a.cost_hist_in = (const R*)b->cost_hist_in; a.lam_hist_in = (const R*)b->lam_hist_in; a.rho_hist_in = (const R*)b->rho_hist_in;
  a.cost_hist_out = (R*)b->cost_hist_out; a.lam_hist_out = (R*)b->lam_hist_out; a.rho_hist_out = (R*)b->rho_hist_out;
  a.xu_out = (R*)b->xu; a.x_out = b->x; a.u_out = b->u; a.status = (R*)b->status; a.factor = (R*)b->factor;
  a.scratch = (R*)b->scratch;
  a.scratch_stride = al_scratch_elems(p->T, Dyn::NX, Dyn::NU);
  const size_t smem = (size_t)kWarpsPerCta * a.scratch_stride * sizeof(R);
  a.use_smem = smem <= kSmemLimit;
  if (!a.use_smem && !a.scratch) return B200QP_EINVAL;
  auto k = k_al_solve<Dyn, R>;
  const size_t dyn_smem = a.use_smem ? smem : 0;
  if (dyn_smem > 48 * 1024) CKM(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
  const int grid = (p->B + kWarpsPerCta - 1) / kWarpsPerCta;
  k<<<grid, 32 * kWarpsPerCta, dyn_smem, st>>>(a);
  CKM(cudaGetLastError());
  return B200QP_OK;
}

template <int NX, int NU, typename R>
static int backward_t(const b200mpc_problem_t* p, const void* factor, const void* xu, const void* grad, void* dC, void* dc,
                      cudaStream_t st) {
  ALBackArgs<R> a;
  a.B = p->B; a.T = p->T; a.factor = (const R*)factor; a.xu = (const R*)xu; a.grad = (const R*)grad;
  a.dC = (R*)dC; a.dc = (R*)dc;
  const size_t smem = (size_t)kWarpsPerCta * pad4(p->T * (NX + NU)) * sizeof(R);
  if (smem > kSmemLimit) return B200QP_ETOOBIG;
  auto k = k_al_backward<NX, NU, R>;
  if (smem > 48 * 1024) CKM(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (p->B + kWarpsPerCta - 1) / kWarpsPerCta;
  k<<<grid, 32 * kWarpsPerCta, smem, st>>>(a);
  CKM(cudaGetLastError());
  return B200QP_OK;
}

template <class Dyn, typename R>
static int dyn_t(const double* params, const void* x, const void* u, void* xn, void* A, void* Bm, long long N, cudaStream_t st) {
  DynParams P;
  for (int i = 0; i < MAX_PARAMS; i++) P.v[i] = params[i];
  const int nt = 128;
  const long long grid = (N + nt - 1) / nt;
  k_dyn_step<Dyn, R><<<(unsigned)grid, nt, 0, st>>>(P, (const R*)x, (const R*)u, (R*)xn, (R*)A, (R*)Bm, N);
  CKM(cudaGetLastError());
  return B200QP_OK;
}

template <class Dyn, typename R>
static int rollout_t(const double* params, const void* x0, const void* u, void* xs, long long B, int T, cudaStream_t st) {
  DynParams P;
  for (int i = 0; i < MAX_PARAMS; i++) P.v[i] = params[i];
  const int nt = 128;
  k_dyn_rollout<Dyn, R><<<(unsigned)((B + nt - 1) / nt), nt, 0, st>>>(P, (const R*)x0, (const R*)u, (R*)xs, B, T);
  CKM(cudaGetLastError());
  return B200QP_OK;
}

#define ENV_DISPATCH(ENV, EXPR_MACRO)                         \
  switch (ENV) {                                              \
    case ENV_PENDULUM: EXPR_MACRO(Pendulum);                  \
    case ENV_INTEGRATOR: EXPR_MACRO(Integrator);              \
    case ENV_PENDULUM_DX: EXPR_MACRO(PendulumDx);             \
    case ENV_CARTPOLE_DX: EXPR_MACRO(CartpoleDx);             \
    case ENV_REX_QUADROTOR: EXPR_MACRO(RexQuadrotor);         \
    case ENV_PENDULUM1L: EXPR_MACRO(Pendulum1L);              \
    case ENV_CARTPOLE1L: EXPR_MACRO(Cartpole1L);              \
    case ENV_CARTPOLE2L: EXPR_MACRO(Cartpole2L);              \
  }                                                           \
  return B200QP_EINVAL

}  // namespace b200mpc

using namespace b200mpc;

extern "C" {

int b200mpc_env_dims(int env, int* nx, int* nu) {
  int a, b;
  if (!nx || !nu || !env_dims(env, a, b)) return B200QP_EINVAL;
  *nx = a; *nu = b;
  return B200QP_OK;
}

size_t b200mpc_factor_elems(const b200mpc_problem_t* prob) {
  int nx, nu;
  if (check(prob, nx, nu)) return 0;
  return (size_t)al_factor_elems(prob->T, nx, nu);
}

size_t b200mpc_scratch_bytes(const b200mpc_problem_t* prob) {
  int nx, nu;
  if (check(prob, nx, nu)) return 0;
  const size_t es = prob->dtype == B200QP_F64 ? 8 : 4;
  const size_t per = (size_t)al_scratch_elems(prob->T, nx, nu) * es;
  if ((size_t)kWarpsPerCta * per <= kSmemLimit) return 0;
  return per * (size_t)prob->B;
}

int b200mpc_al_solve(const b200mpc_problem_t* prob, const b200mpc_buffers_t* b, b200qp_stream_t stream) {
  int nx, nu;
  int rc = check(prob, nx, nu);
  if (rc) return rc;
  if (nx + nu > 32) return B200QP_ETOOBIG;
  if (!b || !b->x_init || !b->u_init || !b->x0 || !b->C || !b->c || !b->u_lower || !b->u_upper || !b->lam || !b->rho ||
      !b->cost_hist_out || !b->lam_hist_out || !b->rho_hist_out || !b->xu || !b->x || !b->u || !b->status || !b->factor)
    return B200QP_EINVAL;
  if (prob->warm && (!b->cost_hist_in || !b->lam_hist_in || !b->rho_hist_in)) return B200QP_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
#define SOLVE_CASE(DYN) return prob->dtype == B200QP_F64 ? solve_t<DYN, double>(prob, b, st) : solve_t<DYN, float>(prob, b, st)
  ENV_DISPATCH(prob->env, SOLVE_CASE);
}

int b200mpc_al_backward(const b200mpc_problem_t* prob, const void* factor, const void* xu, const void* grad, void* dC,
                        void* dc, b200qp_stream_t stream) {
  int nx, nu;
  int rc = check(prob, nx, nu);
  if (rc) return rc;
  if (!factor || !xu || !grad || !dC || !dc) return B200QP_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
#define BACK_CASE(DYN)                                                                                     \
  return prob->dtype == B200QP_F64 ? backward_t<DYN::NX, DYN::NU, double>(prob, factor, xu, grad, dC, dc, st) \
                                   : backward_t<DYN::NX, DYN::NU, float>(prob, factor, xu, grad, dC, dc, st)
  ENV_DISPATCH(prob->env, BACK_CASE);
}

int b200dyn_jac(int env, int dtype, const double* params, const void* x, const void* u, void* xn, void* A, void* Bm,
                int64_t N, b200qp_stream_t stream) {
  if (!params || !x || !u || !xn || N < 0 || (A == nullptr) != (Bm == nullptr)) return B200QP_EINVAL;
  if (dtype != B200QP_F64 && dtype != B200QP_F32) return B200QP_EINVAL;
  if (N == 0) return B200QP_OK;
  cudaStream_t st = (cudaStream_t)stream;
#define DYN_CASE(DYN) \
  return dtype == B200QP_F64 ? dyn_t<DYN, double>(params, x, u, xn, A, Bm, N, st) : dyn_t<DYN, float>(params, x, u, xn, A, Bm, N, st)
  ENV_DISPATCH(env, DYN_CASE);
}

int b200dyn_rollout(int env, int dtype, const double* params, const void* x0, const void* u, void* xs, int64_t B,
                    int32_t T, b200qp_stream_t stream) {
  if (!params || !x0 || !u || !xs || B < 0 || T < 1) return B200QP_EINVAL;
  if (dtype != B200QP_F64 && dtype != B200QP_F32) return B200QP_EINVAL;
  if (B == 0) return B200QP_OK;
  cudaStream_t st = (cudaStream_t)stream;
#define ROLL_CASE(DYN) \
  return dtype == B200QP_F64 ? rollout_t<DYN, double>(params, x0, u, xs, B, T, st) : rollout_t<DYN, float>(params, x0, u, xs, B, T, st)
  ENV_DISPATCH(env, ROLL_CASE);
}

int b200dyn_step(int env, int dtype, const double* params, const void* x, const void* u, void* xn, int64_t N,
                 b200qp_stream_t stream) {
  return b200dyn_jac(env, dtype, params, x, u, xn, nullptr, nullptr, N, stream);
}

}  // extern "C"
