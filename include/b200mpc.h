/* b200mpc.h -- C ABI of the B200-native augmented-Lagrangian MPC solve and batched dynamics
 * (part of libb200qp.so).
 *
 * There is no FFI on this path in the reference (swami1995/diff-qp-mpc): its boundary is the
 * Python call surface `qpth.AL_mpc.MPC(...)(x0, cost, dx, dx_jac)`.  Each entry point below
 * replaces the torch-op sequence of one reference function; INTEGRATION.md shows the ctypes stub.
 *
 *   b200mpc_al_solve     replaces qpth/AL_mpc.py:254-321 (MPC.al_solve) =
 *                        qpth/al_utils.py:16-34 (warm_start_al) + al_iter x
 *                        qpth/al_utils.py:363-460 (NewtonAL.forward: merit_grad_hessian :62-102,
 *                        cholesky_ex + cholesky_solve, line_search_newton :503-527,
 *                        merit_function :37-59) + the multiplier/penalty update :298-310
 *   b200mpc_al_backward  replaces qpth/al_utils.py:462-500 (NewtonAL.backward)
 *   b200dyn_step / _jac  replace the dynamics modules' forward and *_jac forward
 *                        (deqmpc/envs.py:16-31,68-82,199-233; qpth/env_dx/pendulum.py:49-84;
 *                        qpth/env_dx/cartpole.py:63-96)
 *
 * Conventions as in b200qp.h: plain pointers and sizes, device pointers borrowed and contiguous,
 * work enqueued on `stream`, 0 on success / negative B200QP_E* code on error.
 */
#ifndef B200MPC_H
#define B200MPC_H

#include <stddef.h>
#include <stdint.h>
#include "b200qp.h"

#ifdef __cplusplus
extern "C" {
#endif

#define B200MPC_ENV_PENDULUM 0      /* deqmpc/envs.py PendulumDynamics: params dt,g,m,l               */
#define B200MPC_ENV_INTEGRATOR 1    /* deqmpc/envs.py IntegratorDynamics (nx=2,nu=1): params dt       */
#define B200MPC_ENV_PENDULUM_DX 2   /* qpth/env_dx/pendulum.py PendulumDx: params dt,g,m,l,max_torque  */
#define B200MPC_ENV_CARTPOLE_DX 3   /* qpth/env_dx/cartpole.py CartpoleDx: params dt,gravity,masscart,
                                       masspole,length,total_mass,polemass_length,force_mag           */
#define B200MPC_ENV_REX_QUADROTOR 4 /* deqmpc/rex_quadrotor.py RexQuadrotor_dynamics (nx=12, nu=4, RK4): params dt,
                                       mass,act_scale,kf(forces),kf,km,bf,motor_dist,mass*g[3],Bf[3],J[9],Jinv[9],
                                       ss[12],cd[3],cross_A[3]  (float32-rounded where the reference is float32)  */
/* deqmpc/my_envs (CasADi-generated in the reference: <env>/src/generated_dynamics.c, generated_derivatives.c,
 * dynamics_gpu.cu; Python side deqmpc/my_envs/dynamics.py:27-108): state (q, qd), one control on the first joint,
 * classical RK4 with step dt */
#define B200MPC_ENV_PENDULUM1L 5    /* pendulum1l (nx=2): params dt, 1/I, mgl/I                          */
#define B200MPC_ENV_CARTPOLE1L 6    /* cartpole1l, cartpole1l_v2 (nx=4): params dt, mt, ml, I, g         */
#define B200MPC_ENV_CARTPOLE2L 7    /* cartpole2l (nx=6): params dt, mt, h1, h2, J1, J2, k, g            */
#define B200MPC_ENV_CARTPOLE1L_V1 8 /* deqmpc/envs_v1.py:28-82 OneLinkCartpoleDynamics (nx=4): params dt, M, m, l, g */
#define B200MPC_ENV_CARTPOLE2L_V1 9 /* deqmpc/envs_v1.py:226-310 TwoLinkCartpoleDynamics (nx=6): params dt        */
#define B200MPC_MAX_PARAMS 64

typedef struct {
  int32_t B, T;                  /* batch, horizon                                               */
  int32_t env;                   /* B200MPC_ENV_*  (fixes nx and nu)                             */
  int32_t dtype;                 /* B200QP_F64 | B200QP_F32: the solver's internal precision     */
  int32_t al_iter;               /* AL_mpc.MPC(al_iter=2)                                        */
  int32_t newton_steps;          /* 4   (al_utils.py:397)                                        */
  int32_t n_ls;                  /* 20  (al_utils.py:504), at most 32                            */
  int32_t warm;                  /* 0: just (re)initialised, 1: warm-start from the history      */
  int32_t hist_len;              /* entries in the *_hist_in arrays (al_iter+1 of the last call) */
  int32_t reserved;
  double params[B200MPC_MAX_PARAMS];
} b200mpc_problem_t;

typedef struct {
  /* inputs, solver precision */
  const void *x_init, *u_init;   /* (B,T,nx) (B,T,nu) initial trajectory                         */
  const void *x0;                /* (B,nx)                                                       */
  const void *C, *c;             /* (B,T,nx+nu) diagonal of the quadratic cost, linear cost      */
  const void *u_lower, *u_upper; /* (B,T,nu)                                                     */
  /* solver state carried between calls (AL_mpc.py:316-318): in/out */
  void *lam, *rho;               /* (B,M) multipliers, (B) penalty; M = T nx + 2 T nu            */
  const void *cost_hist_in, *lam_hist_in, *rho_hist_in;  /* (K,B) (K,B,M) (K,B), oldest first    */
  void *cost_hist_out, *lam_hist_out, *rho_hist_out;     /* (al_iter+1, ...)                     */
  /* outputs */
  void *xu;                      /* (B,T,nx+nu) solution in solver precision                     */
  float *x, *u;                  /* (B,T,nx) (B,T,nu) float32, what the reference returns        */
  void *status;                  /* (B) 1 if the last line search improved the merit             */
  void *factor;                  /* (B, b200mpc_factor_elems) block Cholesky of the last Hessian */
  void *scratch;                 /* b200mpc_scratch_bytes() bytes (may be NULL when that is 0)    */
} b200mpc_buffers_t;

/* state/control sizes of an environment; returns 0 on success */
int b200mpc_env_dims(int env, int* nx, int* nu);
/* elements (not bytes) per problem of the saved block factor */
size_t b200mpc_factor_elems(const b200mpc_problem_t* prob);
/* bytes of global scratch the solve needs (0 when the problem is shared-memory resident) */
size_t b200mpc_scratch_bytes(const b200mpc_problem_t* prob);

int b200mpc_al_solve(const b200mpc_problem_t* prob, const b200mpc_buffers_t* buf, b200qp_stream_t stream);

/* grad: (B,T,nx+nu) gradient w.r.t. the solution xu; dC, dc: (B,T,nx+nu) */
int b200mpc_al_backward(const b200mpc_problem_t* prob, const void* factor, const void* xu, const void* grad,
                        void* dC, void* dc, b200qp_stream_t stream);

/* xn (N,nx) = f(x (N,nx), u (N,nu)) */
int b200dyn_step(int env, int dtype, const double* params, const void* x, const void* u, void* xn, int64_t N,
                 b200qp_stream_t stream);
/* step + Jacobians A (N,nx,nx) = df/dx, Bm (N,nx,nu) = df/du */
int b200dyn_jac(int env, int dtype, const double* params, const void* x, const void* u, void* xn, void* A, void* Bm,
                int64_t N, b200qp_stream_t stream);

/* open-loop rollout xs (B,T,nx): xs[:,0] = x0, xs[:,t+1] = f(xs[:,t], u[:,t])  (qpth/AL_mpc.py:398-411) */
int b200dyn_rollout(int env, int dtype, const double* params, const void* x0, const void* u, void* xs, int64_t B,
                    int32_t T, b200qp_stream_t stream);

/* Expert-data sampling on the device: deqmpc/datagen.py:358-408 `sample_trajectory` (bsz windows of T consecutive rows of
 * the concatenated expert data, windows may not START on an end-of-trajectory row (mask == 0), rows past the end of the
 * data are zero, the returned mask is the running product of the row masks) followed by deqmpc/utils.py:256-288
 * `unnormalize_states_pendulum` (unnormalize = 1) / `unnormalize_states_cartpole_nlink` (= 2) or nothing (= 0).
 * state [N][nx], action [N][nu], mask [N]: float32 device arrays (the reference keeps them in float32).  idxs [n_idx]: the
 * candidate start rows, drawn by the caller exactly as the reference draws them (np.random.randint(0, N, 2 bsz)); they are
 * consumed in order, skipping masked rows.  sel [bsz] receives the chosen start rows; status[0] the number of admissible
 * candidates (the reference raises IndexError when it is < bsz: outputs are then incomplete). */
int b200data_sample_windows(const float* state, const float* action, const float* mask, long long N, int nx, int nu,
                            const long long* idxs, int n_idx, int bsz, int T, int unnormalize, long long* sel,
                            float* out_state, float* out_action, float* out_mask, int* status, b200qp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200MPC_H */
