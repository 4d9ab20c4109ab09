/* b200qp.h -- C ABI of the B200-native batched PDIPM QP solver (libb200qp.so).
 *
 * There is no FFI on this path in the reference (swami1995/diff-qp-mpc): its boundary is the
 * Python call surface.  Each entry point below replaces the torch-op sequence of one reference
 * function; a maintainer binds it with the ctypes stub shown in INTEGRATION.md.
 *
 *   b200qp_forward   replaces qpth/qp.py:74-126 (QPFunctionFn.forward) =
 *                    qpth/solvers/pdipm/batch.py:377-428 (pre_factor_kkt)
 *                  + qpth/solvers/pdipm/batch.py:46-208  (forward: Mehrotra loop, best-iterate
 *                    tracking, batch-global termination) + :434-469 factor_kkt + :351-374 solve_kkt
 *                  + :211-214 get_step
 *   b200qp_backward  replaces qpth/qp.py:129-183 (QPFunctionFn.backward: adjoint KKT solve that
 *                    reuses the forward pre-factorisation, outer-product gradients)
 *   b200qp_kkt_solve replaces pre_factor_kkt + factor_kkt + solve_kkt as a stand-alone call
 *                    (qpth/solvers/pdipm/batch.py:351-469; what test.py:222-247 exercises)
 *   b200qp_solve_host  forward(+backward) with HOST buffers: the H2D/D2H copies are inside.
 *
 * Conventions: plain pointers and sizes only; all device pointers are borrowed, contiguous,
 * on the current device; all work is enqueued on `stream` (no hidden synchronisation except in
 * b200qp_solve_host); return 0 on success, a negative B200QP_E* code on argument/launch errors.
 * Matrices are row-major.  Batch strides are in ELEMENTS; stride 0 = parameter shared by the
 * whole batch (qpth/util.py:69-75 expandParam).
 */
#ifndef B200QP_H
#define B200QP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200QP_F64 0
#define B200QP_F32 1

#define B200QP_OK 0
#define B200QP_EINVAL (-1)   /* bad sizes / null pointers / unsupported dtype            */
#define B200QP_ECUDA (-2)    /* a CUDA runtime call failed (see b200qp_last_cuda_error)   */
#define B200QP_ETOOBIG (-3)  /* problem does not fit the kernels' shared/global workspace */

/* DenseQPFunction semantics (qpth/qp.py:187-271, qpth/solvers/pdipm/batch_LU.py): every KKT solve is
 * a solve with the matrix regularised by kkt_reg * diag(+I_x, +I_s, -I_z, -I_y) followed by one step
 * of iterative refinement against the unregularised matrix; get_step maps dv == 0 to 1; backward uses
 * d = lams / slacks without the 1e-8 clamp. */
#define B200QP_FLAG_DENSE 1
/* Force the one-launch-per-iteration kernels.  Without it fp64 problems with neq == 0, nineq < 64, nz <= 32 take
 * the RESIDENT route (csrc/qp_resident.cuh): several iterations per launch with the batch-global step fill of
 * get_step (qpth/solvers/pdipm/batch.py:211-214) speculated and repaired, bit-identical results.  Two situations
 * are never speculated (a ratio test that is fill-only while the fill is still unknown; z and s ratios that turn
 * NaN at different iterations): forward then sets status[B200QP_ST_SPEC_FAIL] != 0, the outputs are NOT valid,
 * and the caller repeats the call with this flag (b200qp_solve_host* and the Python layer do so themselves). */
#define B200QP_FLAG_EXACT 2
/* Host-buffer entry points only (b200qp_solve_host*): return the gradients w.r.t. Q and G in FACTORED form.  The
 * reference forms them as outer products of four vectors (qpth/qp.py:158-174): dQ = (dx z^T + z dx^T) / 2,
 * dG = dlam z^T + lam dx^T with dx = dp, dlam = -dh, z = zhat, lam = lams -- all of which the call returns anyway.
 * With this flag dQ and dG are neither written nor copied to the host (they may be NULL): the device->host
 * traffic of a step drops from 23.5 KB to 2.2 KB per problem at nz = 30, nineq = 60, which is what bounds the
 * pipelined host path (PCIe duplex).  The dense form stays the default contract. */
#define B200QP_FLAG_FACTORED_GRAD 4

#define B200QP_MAX_ITER_CAP 64
#define B200QP_STATUS_DOUBLES 8

/* status[] layout (device buffer of B200QP_STATUS_DOUBLES doubles, written by forward) */
#define B200QP_ST_NITER 0        /* loop bodies entered (== reference's iteration count)      */
#define B200QP_ST_BEST_MAX 1     /* max over the batch of the best residual (NaN propagates)  */
#define B200QP_ST_Q_FAIL 2       /* number of problems whose Q factorisation failed           */
#define B200QP_ST_AQA_FAIL 3     /* number of problems whose A Q^-1 A^T factorisation failed  */
#define B200QP_ST_LAUNCHES 4     /* kernels launched by the call                              */
#define B200QP_ST_SPEC_FAIL 5    /* resident route only: != 0 -> repeat with B200QP_FLAG_EXACT  */
#define B200QP_ST_NAN_ONSET 6    /* resident route only: first iteration with a NaN step ratio (-1: none) */

typedef void* b200qp_stream_t;   /* a cudaStream_t */

typedef struct {
  int32_t nb, nz, nineq, neq;    /* batch, variables, inequality rows, equality rows          */
  int32_t dtype;                 /* B200QP_F64 | B200QP_F32                                   */
  int32_t max_iter;              /* maxIter        (qpth/qp.py:19-21 default 20)              */
  int32_t not_improved_lim;      /* notImprovedLim (default 3)                                */
  int32_t flags;                 /* B200QP_FLAG_* (0 for QPFunction)                           */
  double eps;                    /* eps            (default 1e-12)                            */
  int64_t sQ, sp, sG, sh, sA, sb; /* batch strides in elements, 0 = shared                    */
  double kkt_reg;                /* DenseQPFunction's KKT regularisation (1e-7), else 0       */
} b200qp_problem_t;

/* Bytes of device workspace forward needs; the same buffer must be handed to backward. */
size_t b200qp_workspace_bytes(const b200qp_problem_t* prob);

/* Solve the batch.  Outputs: zhat (nb,nz), lams (nb,nineq), nus (nb,neq), slacks (nb,nineq),
 * status (device, 8 doubles).  workspace: b200qp_workspace_bytes() bytes, 16-byte aligned. */
int b200qp_forward(const b200qp_problem_t* prob,
                   const void* Q, const void* p, const void* G, const void* h,
                   const void* A, const void* b,
                   void* zhat, void* lams, void* nus, void* slacks,
                   void* workspace, double* status, b200qp_stream_t stream);

/* The forward in phases, for callers that must couple several devices: the reference's termination
 * test and its get_step fill are reductions over the WHOLE batch (qpth/solvers/pdipm/batch.py:127-131,
 * 141,213); inside one call they live in one 64-byte slot per iteration at
 * workspace + b200qp_slot_offset(prob) + 64 * it, every field of which is an unsigned 64/32-bit MAX
 * reduction.  A batch sharded over R ranks is solved EXACTLY like the unsharded batch by running
 *   phase BEGIN; for it in 0..max_iter-1 { phase it; element-wise unsigned MAX of slot `it` across ranks }
 *   phase END
 * (b200qp/qp.py does this with torch.distributed when `process_group` is given). */
#define B200QP_PHASE_ALL (-1000)   /* what b200qp_forward does                                  */
#define B200QP_PHASE_BEGIN (-1001) /* zero the slots, pre-factorise, initial point              */
#define B200QP_PHASE_END (-1002)   /* iteration count / worst residual -> status                */
size_t b200qp_slot_offset(const b200qp_problem_t* prob);

/* Caller-evaluated residual callbacks.  This fork of qpth evaluates cost_grad(x) (in place of Q x + p) and dyn_res(x)
 * (in place of A x - b) at the top of every PDIPM iteration (qpth/solvers/pdipm/batch.py:93-102; its MPC callers pass
 * the NON-linear dynamics residual, qpth/qp_wrapper.py:303-316, sl1qp_mpc.py:312-320).  Protocol, one launch per
 * iteration like the phased mode above:
 *   b200qp_forward_phase(BEGIN)
 *   for it in 0..max_iter-1:
 *     b200qp_forward_cb_step(it, x)        applies the step of iteration it-1 (nothing once the loop has terminated)
 *                                          and writes the iterate x[nb][nz] the reference would hand to the callbacks
 *     caller evaluates cg = cost_grad(x) [nb][nz] and / or ry = dyn_res(x) [nb][neq] on the same stream
 *     b200qp_forward_phase_cb(it, ..., cg or NULL, ry or NULL)
 *   b200qp_forward_phase(END)
 * A NULL vector keeps the kernel's own Q x + p / A x - b.  The KKT solves, step rule, termination and backward are
 * those of the plain call (the reference's are, too). */
int b200qp_forward_cb_step(const b200qp_problem_t* prob, int it, void* x_out, void* workspace, b200qp_stream_t stream);
int b200qp_forward_phase_cb(const b200qp_problem_t* prob, int phase,
                            const void* Q, const void* p, const void* G, const void* h,
                            const void* A, const void* b,
                            void* zhat, void* lams, void* nus, void* slacks,
                            void* workspace, double* status,
                            const void* cost_grad_x, const void* dyn_res_x, b200qp_stream_t stream);
int b200qp_forward_phase(const b200qp_problem_t* prob, int phase,
                         const void* Q, const void* p, const void* G, const void* h,
                         const void* A, const void* b,
                         void* zhat, void* lams, void* nus, void* slacks,
                         void* workspace, double* status, b200qp_stream_t stream);

/* Adjoint solve.  Needs the forward's workspace (pre-factorisation) and outputs.
 * Writes PER-PROBLEM gradients: dQ (nb,nz,nz) dp (nb,nz) dG (nb,nineq,nz) dh (nb,nineq)
 * dA (nb,neq,nz) db (nb,neq); the caller averages them for shared parameters
 * (qpth/qp.py:160-178 uses .mean(0)). */
int b200qp_backward(const b200qp_problem_t* prob,
                    const void* zhat, const void* lams, const void* nus, const void* slacks,
                    const void* dl_dzhat,
                    void* dQ, void* dp, void* dG, void* dh, void* dA, void* db,
                    void* workspace, b200qp_stream_t stream);

/* The d-independent pre-factorisation alone (qpth/solvers/pdipm/batch.py:377-428 pre_factor_kkt):
 * Q^-1, [A;G]Q^-1, the Schur blocks and their factor go to `workspace` for later
 * b200qp_kkt_solve(prefactor = 0) calls.  status (device, 8 doubles, may be NULL) receives the
 * Q / A Q^-1 A^T failure counts in slots B200QP_ST_Q_FAIL / B200QP_ST_AQA_FAIL. */
int b200qp_prefactor(const b200qp_problem_t* prob, const void* Q, const void* G, const void* A,
                     void* workspace, double* status, b200qp_stream_t stream);

/* Stand-alone KKT solve: given d (nb,nineq) and right-hand sides, return the solution of
 *   [Q 0 G' A'; 0 D I 0; G I 0 0; A 0 0 0] [dx ds dz dy]' = -[rx rs rz ry]'.
 * prefactor != 0 recomputes the d-independent part from Q, G, A first. */
int b200qp_kkt_solve(const b200qp_problem_t* prob, int prefactor,
                     const void* Q, const void* G, const void* A, const void* d,
                     const void* rx, const void* rs, const void* rz, const void* ry,
                     void* dx, void* ds, void* dz, void* dy,
                     void* workspace, b200qp_stream_t stream);

/* Whole solve with HOST buffers (pageable or pinned): copies inputs to the device, runs
 * forward (and backward when dl_dzhat != NULL), copies results back, synchronises.
 * Gradient pointers may be NULL when dl_dzhat is NULL.  status: 8 host doubles. */
int b200qp_solve_host(const b200qp_problem_t* prob,
                      const void* Q, const void* p, const void* G, const void* h,
                      const void* A, const void* b, const void* dl_dzhat,
                      void* zhat, void* lams, void* nus, void* slacks,
                      void* dQ, void* dp, void* dG, void* dh, void* dA, void* db,
                      double* status);

/* Pipelined form of b200qp_solve_host for back-to-back batches (serving): B200QP_HOST_SLOTS slots
 * (0 .. B200QP_HOST_SLOTS-1), each with its own device arena.  submit(slot, ...) enqueues H2D copies,
 * the solve and the D2H copies and returns; wait(slot) blocks until that job's results are in the
 * caller's host buffers.  Inputs of job k+1 travel while job k computes and job k-1's gradients travel
 * back (three streams, PCIe full duplex): a job has three stages, so rotate over THREE slots to keep
 * every stage busy (two slots leave the bus idle a third of the time).  A slot's host buffers must stay
 * untouched between submit and wait; submitting to a busy slot waits for it first.  Host buffers should
 * be pinned (pageable memory makes the copies synchronous). */
#define B200QP_HOST_SLOTS 4
int b200qp_solve_host_submit(int slot, const b200qp_problem_t* prob,
                             const void* Q, const void* p, const void* G, const void* h,
                             const void* A, const void* b, const void* dl_dzhat,
                             void* zhat, void* lams, void* nus, void* slacks,
                             void* dQ, void* dp, void* dG, void* dh, void* dA, void* db,
                             double* status);
int b200qp_solve_host_wait(int slot);

/* Launch profiling for bench.py's roofline leg (no reference counterpart): when enabled, every
 * kernel launch of forward/backward is bracketed by CUDA events on the caller's stream.
 * b200qp_profile_read synchronises on the last event and returns the number of launches n,
 * filling ms[i] (duration) and kind[i] (0 prefactor, 1 initial point, 2 PDIPM iteration,
 * 3 finalize, 4 backward, 5 one launch of the resident route = several iterations of every problem,
 * 6 its history reduction + selection) for i < min(n, cap). */
void b200qp_profile_enable(int on);
int b200qp_profile_read(float* ms, int* kind, int cap);

/* Process-wide tuning knobs (defaults come from the environment variables of the same upper-case name with the
 * prefix B200QP_, read once): "res" (1) take the resident route when eligible, "res_ch" (4) its iterations per
 * launch, "res_spec" (1) its compile-time-size specialisation of the nz = 30 / nineq = 60 shape; the routing overrides
 * "mid_fast" (B200QP_MID=fast), "blk_nt" (256), "factor_tile" (B200QP_FACTOR=tile), "force_generic".  Returns 0, or
 * B200QP_EINVAL for an unknown name.  Not to be changed between the forward and the backward of one problem. */
int b200qp_set_option(const char* name, int value);

/* Text of the last CUDA error seen by this library on the calling thread ("" if none). */
const char* b200qp_last_cuda_error(void);

/* Library version / build tag, e.g. "b200qp 0.1 sm_100a". */
const char* b200qp_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200QP_H */
