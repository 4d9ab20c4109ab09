// b200mpc.cu -- host side of the MPC entry points of libb200qp.so (include/b200mpc.h): problem
// validation, shared-memory sizing, environment dispatch.  No torch types anywhere.
#include <cuda_runtime.h>
#include <stdlib.h>
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
// B200MPC_FORCE_GLOBAL=1 (read once): keep every problem's state in the global scratch slab, whatever its size -- the mode
// long horizons / the rex quadrotor take by necessity; lets the tests drive those code paths with the small goldens
static bool force_global() {
  static const bool v = [] { const char* e = getenv("B200MPC_FORCE_GLOBAL"); return e && e[0] && e[0] != '0'; }();
  return v;
}

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
    case ENV_CARTPOLE1L_V1: nx = Cartpole1LV1::NX; nu = Cartpole1LV1::NU; return true;
    case ENV_CARTPOLE2L_V1: nx = Cartpole2LV1::NX; nu = Cartpole2LV1::NU; return true;
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
  a.use_smem = smem <= kSmemLimit && !force_global();
  if (!a.use_smem && !a.scratch) return B200QP_EINVAL;
  auto k = k_al_solve<Dyn, R>;
  // global-scratch mode: 5 KB per warp to stage the block being factored (block_cholesky_staged)
  const size_t dyn_smem = a.use_smem ? smem : (size_t)kWarpsPerCta * al_stage_elems<Dyn::NX, Dyn::NU>() * sizeof(R);
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
    case ENV_CARTPOLE1L_V1: EXPR_MACRO(Cartpole1LV1);         \
    case ENV_CARTPOLE2L_V1: EXPR_MACRO(Cartpole2LV1);         \
  }                                                           \
  return B200QP_EINVAL


// ---------------------------------------------------------------------------------------------------------------
// Expert-data sampling on the device (deqmpc/datagen.py:358-408 sample_trajectory, deqmpc/utils.py:256-288
// unnormalize_states_*): the reference assembles every training batch with a Python loop over bsz windows, a torch.cat per
// window and a sequential un-normalisation over T on the host.  Here the concatenated expert data stays in HBM, the
// candidate start indices (drawn by the caller with the reference's own RNG call) are filtered and compacted by one CTA,
// and one thread per window gathers (state, action, mask), forms the running mask product and un-wraps the angles.
__global__ void __launch_bounds__(1024) k_data_select(const float* __restrict__ mask, const long long* __restrict__ idxs, int n_idx,
                                                      int bsz, long long* __restrict__ sel, int* __restrict__ status) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int start = 0; start < n_idx; start += 1024) {
    const int i = start + tid;
    const long long ix = i < n_idx ? idxs[i] : 0;
    const int v = (i < n_idx && mask[ix] != 0.0f) ? 1 : 0;   // datagen.py:374: windows may not start on an end-of-trajectory row
    const unsigned bal = __ballot_sync(0xffffffffu, v);
    const int wpre = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; w++) off += s_warp[w];
    const int rank = off + wpre;
    if (v && rank < bsz) sel[rank] = ix;
    __syncthreads();
    if (tid == 0) { int t = 0; for (int w = 0; w < 32; w++) t += s_warp[w]; s_base += t; }
    __syncthreads();
  }
  if (tid == 0) status[0] = s_base;
}

__global__ void __launch_bounds__(128) k_data_gather(const float* __restrict__ state, const float* __restrict__ action,
                                                     const float* __restrict__ mask, long long N, int nx, int nu,
                                                     const long long* __restrict__ sel, int bsz, int T, int unnorm,
                                                     float* __restrict__ os, float* __restrict__ oa, float* __restrict__ om) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= bsz) return;
  const long long i0 = sel[j];
  float* s = os + (size_t)j * T * nx;
  float* a = oa + (size_t)j * T * nu;
  float* m = om + (size_t)j * T;
  float run = 1.0f;
  for (int t = 0; t < T; t++) {
    const long long r = i0 + t;
    const bool in = r < N;                                   // datagen.py:381-399: zero padding past the end of the data
    for (int c = 0; c < nx; c++) s[t * nx + c] = in ? state[r * nx + c] : 0.0f;
    for (int c = 0; c < nu; c++) a[t * nu + c] = in ? action[r * nu + c] : 0.0f;
    run *= in ? mask[r] : 0.0f;                              // datagen.py:404-405: running product of the masks
    m[t] = run;
  }
  const float kHalfPi = 1.57079637050628662f, kTwoPi = 6.28318548202514648f;  // float32(pi / 2), 2 * float32(pi)
  if (unnorm == 1) {                                         // utils.py:256-270 (pendulum): sign of the ANGLE, column 0
    float prev = s[0];
    for (int t = 0; t < T; t++) {
      const float cur = s[t * nx];
      if (fabsf(cur - prev) > kHalfPi) s[t * nx] = cur - (cur > 0.0f ? 1.0f : (cur < 0.0f ? -1.0f : 0.0f)) * kTwoPi;
      prev = s[t * nx];
    }
  } else if (unnorm == 2) {                                  // utils.py:273-288 (n-link cart-pole): sign of the JUMP, columns
    const int nq = nx / 2 + 1;                               // 1 .. nx/2 (the reference's own column range, kept as is)
    for (int c = 1; c < nq; c++) {
      float prev = s[c];
      for (int t = 0; t < T; t++) {
        const float cur = s[t * nx + c];
        const float d = cur - prev;
        if (fabsf(d) > kHalfPi) s[t * nx + c] = cur - (d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f)) * kTwoPi;
        prev = s[t * nx + c];
      }
    }
  }
}

}  // namespace b200mpc

using namespace b200mpc;

extern "C" {

int b200data_sample_windows(const float* state, const float* action, const float* mask, long long N, int nx, int nu,
                            const long long* idxs, int n_idx, int bsz, int T, int unnormalize, long long* sel,
                            float* out_state, float* out_action, float* out_mask, int* status, b200qp_stream_t stream) {
  if (!state || !action || !mask || !idxs || !sel || !out_state || !out_action || !out_mask || !status) return B200QP_EINVAL;
  if (N < 1 || nx < 1 || nu < 1 || n_idx < 1 || bsz < 1 || T < 1 || unnormalize < 0 || unnormalize > 2) return B200QP_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  CKM(cudaMemsetAsync(sel, 0, sizeof(long long) * (size_t)bsz, st));
  k_data_select<<<1, 1024, 0, st>>>(mask, idxs, n_idx, bsz, sel, status);
  CKM(cudaGetLastError());
  k_data_gather<<<(unsigned)((bsz + 127) / 128), 128, 0, st>>>(state, action, mask, N, nx, nu, sel, bsz, T, unnormalize,
                                                              out_state, out_action, out_mask);
  CKM(cudaGetLastError());
  return B200QP_OK;
}

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
  if ((size_t)kWarpsPerCta * per <= kSmemLimit && !force_global()) return 0;
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
