// qp_kernels.cuh -- the batched PDIPM QP kernels (one CTA per QP, sm_100a).
//
// Reference behaviour being re-implemented (swami1995/diff-qp-mpc, paths under /root/reference):
//   qpth/solvers/pdipm/batch.py:377-428  pre_factor_kkt   -> k_prefactor
//   qpth/solvers/pdipm/batch.py:46-208   forward          -> k_pdipm_iter (one launch / iteration)
//   qpth/solvers/pdipm/batch.py:434-469  factor_kkt       -> ldlt_factor on T = R + diag(s/z)
//   qpth/solvers/pdipm/batch.py:351-374  solve_kkt        -> kkt_solve
//   qpth/solvers/pdipm/batch.py:211-214  get_step         -> step_pieces + deferred fill
//   qpth/qp.py:129-183                   backward         -> k_backward
//
// Elimination used here (mathematically the reference's block elimination, re-ordered so the
// only sequential work per iteration is one nineq x nineq LDL^T and two solves with it):
//   once per call:   Qi = Q^-1,  BQi = [A;G] Qi,  [Saa Sag; Sga Sgg] = BQi [A;G]^T,
//                    Saa = La Da La^T,  V = La^-1 Sag,  R = Sgg - V^T Da^-1 V
//   per iteration:   T = R + diag(s/z) = L D L^T
//   per solve:       t = Qi rx; [hy;hz] = BQi rx + [-ry; rs/d - rz]; u = La^-1 hy;
//                    qz = T^-1 (hz - V^T Da^-1 u); qy = La^-T Da^-1 (u - V qz);
//                    dz = -qz, dy = -qy, ds = (-rs - dz)/d, dx = -t + BQi^T [qy;qz]
// oracle/algo_model.py is the same algebra on the CPU and is what the golden-vector tests pin.
//
// Batch-global couplings of the reference loop are resolved at kernel boundaries through one
// `Slot` of device-side reductions per iteration (no host round trip, CUDA-graph capturable):
// iteration i's kernel reads slot i-1 to (a) decide whether the reference would have returned
// at iteration i-1, (b) obtain the get_step fill values of iteration i-1's final direction, then
// applies that step and carries on; work done after the reference's return point is never
// observable because best iterates are updated before it.
#pragma once
#include <type_traits>
#include "qp_common.cuh"
#include "qp_blocked.cuh"

namespace b200qp {

struct Slot {                       // zero-initialised once per forward call
  unsigned long long best_max;      // max_b best_resid[b]          (bits of a non-negative double)
  unsigned long long mu_min_inv;    // ~ord_key(min_b mu[b])        (so that atomicMax works)
  unsigned long long amax_z;        // ord_key(max_b,i  -z/dz)
  unsigned long long amax_s;        // ord_key(max_b,i  -s/ds)
  unsigned int improved;            // any problem improved its best residual
  unsigned int best_nan, mu_nan, az_nan, as_nan;  // NaN seen in the respective reduction
  unsigned int pad[3];
};
static_assert(sizeof(Slot) == 64, "Slot must be 64 bytes");

struct Control {                    // lives after the slots
  unsigned int q_fail;              // problems whose Q factorisation failed
  unsigned int aqa_fail;            // problems whose A Q^-1 A^T factorisation failed
  unsigned int kev_inv;             // resident route (qp_resident.cuh): kKevBase - first iteration with a NaN step ratio
  unsigned int need_exact;          // resident route: something that is never speculated occurred -> rerun exactly
  unsigned int pad[12];
};


// Position of R[i][k] (k <= i) in register-tile order for a (mpad, nt) tiling: block q (row-major
// over the (a,b) blocks that touch the lower triangle), thread t = (k%TC)*TR + i%TR, with a
// TR x TC thread grid (TR = 8 for one-warp CTAs, 16 otherwise).
__host__ __device__ inline int rtile_index(int i, int k, int mpad, int nt) {
  if (nt < 0) {  // DMMA accumulator-fragment order (qp_dmma.cuh: frag_index)
    const int I = i >> 3, K = k >> 3, r = i & 7, c = k & 7;
    return (I * (I + 1) / 2 + K) * 64 + (r * 4 + (c >> 1)) * 2 + (c & 1);
  }
  const int TR = nt == 32 ? 8 : 16, TC = nt / TR, NB = mpad / TC;
  const int a = i / TR, ti = i - a * TR, b = k / TC, tk = k - b * TC;
  int q = b;
  for (int aa = 0; aa < a; aa++) {
    int c = (TR * aa + TR - 1) / TC + 1;
    q += c < NB ? c : NB;
  }
  return q * nt + tk * TR + ti;
}
__host__ __device__ inline int rtile_elems(int mpad, int nt) {
  const int TR = nt == 32 ? 8 : 16, TC = nt / TR, NA = mpad / TR, NB = mpad / TC;
  int q = 0;
  for (int aa = 0; aa < NA; aa++) {
    int c = (TR * aa + TR - 1) / TC + 1;
    q += c < NB ? c : NB;
  }
  return q * nt;
}

constexpr int FLAG_POISON = 1, FLAG_FILL_Z = 2, FLAG_FILL_S = 4;

template <typename T>
struct KArgs {
  int nb, n, m, p;
  int ldn, ldm, ldp;
  // inputs (borrowed)
  const T *Q, *pv, *G, *h, *A, *b;
  long long sQ, sp, sG, sh, sA, sb;
  // pre-factorisation (workspace, per-problem strides in elements)
  T *Qi, *BQi, *R, *V, *UA, *pinvA, *F, *pinvF, *Tscr, *pinvTscr;
  long long sQi, sBQi, sR, sV, sUA, sF, sT;
  // iterate + direction
  T *x, *s, *z, *y, *dx, *ds, *dz, *dy, *rmu;
  int* flags;
  double* best_resid;
  // outputs (best iterate)
  T *bx, *bs, *bz, *by;
  Slot* slots;
  Control* ctl;
  double* status;
  int iter, max_iter, lim;
  double eps;
  int launches;
  int dense;                 // DenseQPFunction semantics (B200QP_FLAG_DENSE)
  double reg;                // KKT regularisation of the dense mode, else 0
  int prob0;                 // first problem of this launch (chunked pre-factorisation / backward)
  int fso[36];               // fast path: shared-memory carve-up offsets in elements (qp_fast.cuh:fast_offsets)
  int pre_smem;              // prefactor: F / Qi working copies live in dynamic shared memory
  int pre_blocked;           // prefactor: byte offset (+1) of the panel buffer in dynamic shared memory when the
                             // blocked tensor-core route is taken (fp64, large nz; qp_blocked.cuh), else 0
  int rtile_mpad, rtile_nt;  // != 0: R is stored in register-tile order (qp_fast.cuh); nt = -1: DMMA fragment order
  // caller-evaluated residual callbacks of this fork (qpth/solvers/pdipm/batch.py:93-102): when non-null, cb_cg[nb][n]
  // = cost_grad(x) replaces Q x + p in rx and cb_ry[nb][p] = dyn_res(x) replaces A x - b (b200qp_forward_phase_cb)
  const T *cb_cg, *cb_ry;
};

// ------------------------------------------------------------------------------------------
// Decide, from the slots of iterations < upto, whether the reference loop has returned, and at
// which iteration.  Executed by one full warp; every lane returns the same value.
// Returns the 0-based iteration at which the reference returned, or -1.
__device__ __forceinline__ int eval_termination(const Slot* slots, int upto, int lim, double eps, int lane) {
  unsigned long long mi = 0ULL, mc = 0ULL;
  for (int base = 0; base < upto; base += 32) {
    const int j = base + lane;
    bool improved = false, c23 = false;
    if (j < upto) {
      const Slot* sl = slots + j;
      improved = sl->improved != 0;
      const double bmax = sl->best_nan ? __longlong_as_double(0x7ff8000000000000LL)
                                       : __longlong_as_double((long long)sl->best_max);
      const double mmin = sl->mu_nan ? __longlong_as_double(0x7ff8000000000000LL) : ord_unkey(~sl->mu_min_inv);
      c23 = (bmax < eps) || (mmin > 1e32);
    }
    mi |= (unsigned long long)__ballot_sync(0xffffffffu, improved) << base;
    mc |= (unsigned long long)__ballot_sync(0xffffffffu, c23) << base;
  }
  int stall = 0;
  for (int j = 0; j < upto; j++) {
    if (j > 0) stall = ((mi >> j) & 1ULL) ? 0 : stall + 1;
    if (stall == lim || ((mc >> j) & 1ULL)) return j;
  }
  return -1;
}

// Shared-memory carve-up of one problem.
template <typename T>
struct Smem {
  T *Tm, *pinvT, *BQi, *V, *UA, *pinvA;
  T *x, *s, *z, *y, *d, *rx, *rz, *ry, *t, *hv, *u;
  T *dxa, *dsa, *dza, *dya, *rsc, *dxc, *dsc, *dzc, *dyc, *part, *red;
  T* panel;  // global-resident T only: panel buffer of the blocked factorisation (qp_blocked.cuh), else nullptr
  int* ctrl;
};

__host__ __device__ inline int round4(int v) { return (v + 3) & ~3; }

// Elements of T needed in shared memory; mats=true includes the matrices.
__host__ __device__ inline size_t smem_elems(int n, int m, int p, int ldn, int ldm, int ldp, int nt, bool mats) {
  size_t e = 0;
  if (mats) {
    e += round4(m * ldm) + round4(m) + round4((p + m) * ldn);
    if (p > 0) e += round4(p * ldm) + round4(p * ldp) + round4(p);
  }
  e += (size_t)5 * round4(n) + (size_t)10 * round4(m) + (size_t)6 * round4(p > 0 ? p : 1) + round4(p + m);
  e += round4(nt) + 4 * 32;
  if (!mats) e += round4(blk_panel_elems(m));
  e += 8;  // ctrl ints
  return e;
}

template <typename T, bool SMEM>
__device__ __forceinline__ void carve(Smem<T>& S, unsigned char* raw, const KArgs<T>& a, int prob, int nt) {
  T* q = reinterpret_cast<T*>(raw);
  auto take = [&](int cnt) { T* r = q; q += round4(cnt); return r; };
  const int n = a.n, m = a.m, p = a.p, pp = p > 0 ? p : 1;
  if (SMEM) {
    S.Tm = take(m * a.ldm);
    S.pinvT = take(m);
    S.BQi = take((p + m) * a.ldn);
    if (p > 0) {
      S.V = take(p * a.ldm);
      S.UA = take(p * a.ldp);
      S.pinvA = take(p);
    } else {
      S.V = S.UA = S.pinvA = nullptr;
    }
  } else {
    S.Tm = a.Tscr + (size_t)prob * a.sT;
    S.pinvT = a.pinvTscr + (size_t)prob * round4(m);
    S.BQi = a.BQi + (size_t)prob * a.sBQi;
    S.V = a.V + (size_t)prob * a.sV;
    S.UA = a.UA + (size_t)prob * a.sUA;
    S.pinvA = a.pinvA + (size_t)prob * round4(pp);
  }
  S.x = take(n); S.rx = take(n); S.t = take(n); S.dxa = take(n); S.dxc = take(n);
  S.s = take(m); S.z = take(m); S.d = take(m); S.rz = take(m); S.dsa = take(m); S.dza = take(m);
  S.rsc = take(m); S.dsc = take(m); S.dzc = take(m); T* spare = take(m); (void)spare;
  S.y = take(pp); S.ry = take(pp); S.u = take(pp); S.dya = take(pp); S.dyc = take(pp); T* sp2 = take(pp); (void)sp2;
  S.hv = take(p + m);
  S.part = take(nt);
  S.red = take(4 * 32);
  S.panel = SMEM ? nullptr : take(blk_panel_elems(m));
  S.ctrl = reinterpret_cast<int*>(q);
}

// Stage the d-independent matrices of problem `prob` into shared memory (async).
template <typename T, bool SMEM>
__device__ __forceinline__ void stage_mats(const Smem<T>& S, const KArgs<T>& a, int prob, int tid, int nt) {
  if (SMEM) {
    cp_async_block(S.Tm, a.R + (size_t)prob * a.sR, round4(a.m * a.ldm), tid, nt);
    cp_async_block(S.BQi, a.BQi + (size_t)prob * a.sBQi, round4((a.p + a.m) * a.ldn), tid, nt);
    if (a.p > 0) {
      cp_async_block(S.V, a.V + (size_t)prob * a.sV, round4(a.p * a.ldm), tid, nt);
      cp_async_block(S.UA, a.UA + (size_t)prob * a.sUA, round4(a.p * a.ldp), tid, nt);
      cp_async_block(S.pinvA, a.pinvA + (size_t)prob * round4(a.p), round4(a.p), tid, nt);
    }
    cp_async_commit();
  }
}

// T = R + diag(1/d) and its LDL^T.  Needs the staged copy of R (SMEM) or copies it (global).
template <typename T, bool SMEM, int NT>
__device__ __forceinline__ bool build_and_factor_T(const Smem<T>& S, const KArgs<T>& a, int prob, int tid, int nt) {
  const int m = a.m, ldm = a.ldm;
  if constexpr (!SMEM && std::is_same<T, double>::value) {
    // global-resident T: blocked tensor-core factorisation that reads R + diag(1/d) in place (qp_blocked.cuh)
    return ldlt_factor_blocked(S.Tm, ldm, m, S.pinvT, S.panel, tid, nt, a.R + (size_t)prob * a.sR, S.d);
  }
  if (SMEM) {
    cp_async_wait_all();
    __syncthreads();
  } else {
    const T* R = a.R + (size_t)prob * a.sR;
    for (int i = tid; i < m * ldm; i += nt) S.Tm[i] = R[i];
    __syncthreads();
  }
  for (int i = tid; i < m; i += nt) S.Tm[(size_t)i * ldm + i] += T(1) / S.d[i];
  __syncthreads();
  if (SMEM) {
    if (NT == 128) {
      if (m <= 32) return ldlt_factor_reg<T, 32, 128>(S.Tm, ldm, m, S.pinvT, S.part, tid);
      if (m <= 64) return ldlt_factor_reg<T, 64, 128>(S.Tm, ldm, m, S.pinvT, S.part, tid);
    } else {
      if (m <= 128) return ldlt_factor_reg<T, 128, 256>(S.Tm, ldm, m, S.pinvT, S.part, tid);
    }
  }
  return ldlt_factor(S.Tm, ldm, m, S.pinvT, tid, nt);
}

// Block-elimination KKT solve (see file header).  rx / rz / ry may be nullptr (= zero vector).
// All vector arguments are shared-memory arrays; outputs must not alias inputs.
// Entry: inputs visible (caller barrier).  Exit: outputs visible (ends with a barrier).
template <typename T>
__device__ __forceinline__ void kkt_solve(const Smem<T>& S, const KArgs<T>& a, int prob, const T* rx, const T* rs,
                                          const T* rz, const T* ry, T* dx, T* ds, T* dz, T* dy, int tid, int nt) {
  const int n = a.n, m = a.m, p = a.p, ldn = a.ldn, ldm = a.ldm, ldp = a.ldp;
  const int lane = tid & 31, warp = tid >> 5;
  if (rx) {
    if (S.panel != nullptr) gemv_rows_warp(S.BQi, ldn, p + m, n, rx, S.hv, tid, nt);  // BQi in global: coalesced rows
    else gemv_rows_thread(S.BQi, ldn, p + m, n, rx, S.hv, tid, nt);
    gemv_cols(a.Qi + (size_t)prob * a.sQi, ldn, n, n, rx, S.t, S.part, tid, nt);  // ends with barrier
  }
  for (int i = tid; i < m; i += nt) {
    T v = rs[i] / S.d[i];
    if (rz) v -= rz[i];
    S.hv[p + i] = rx ? S.hv[p + i] + v : v;
  }
  for (int j = tid; j < p; j += nt) {
    T v = rx ? S.hv[j] : T(0);
    if (ry) v -= ry[j];
    S.u[j] = v;
  }
  __syncthreads();
  if (p > 0) {
    if (warp == 0) {
      unit_fwd_warp(S.UA, ldp, p, S.u, lane);  // u = La^-1 hy
      for (int j = lane; j < p; j += 32) S.hv[j] = S.u[j] * S.pinvA[j];  // Da^-1 u
    }
    __syncthreads();
    // hz -= V^T (Da^-1 u)
    gemv_cols(S.V, ldm, p, m, S.hv, S.dsc /*scratch*/, S.part, tid, nt);
    for (int i = tid; i < m; i += nt) S.hv[p + i] -= S.dsc[i];
    __syncthreads();
  }
  bool blocked = false;
  if constexpr (std::is_same<T, double>::value) {
    if (S.panel != nullptr) {  // global-resident factor: blocked sweeps by the whole CTA
      ldlt_solve_blocked(S.Tm, ldm, m, S.pinvT, S.hv + p, tid, nt);  // qz
      blocked = true;
    }
  }
  if (!blocked) {
    if (warp == 0) ldlt_solve_warp(S.Tm, ldm, m, S.pinvT, S.hv + p, lane);  // qz
    __syncthreads();
  }
  if (p > 0) {
    gemv_rows_thread(S.V, ldm, p, m, S.hv + p, S.hv, tid, nt);  // V qz  -> hv[0:p]
    __syncthreads();
    if (warp == 0) {
      for (int j = lane; j < p; j += 32) S.hv[j] = (S.u[j] - S.hv[j]) * S.pinvA[j];
      __syncwarp();
      unit_bwd_warp(S.UA, ldp, p, S.hv, lane);  // qy
    }
    __syncthreads();
  }
  // dx = -t + BQi^T q ; dz = -qz ; ds = (-rs - dz)/d ; dy = -qy
  gemv_cols(S.BQi, ldn, p + m, n, S.hv, dx, S.part, tid, nt);
  for (int c = tid; c < n; c += nt) dx[c] = rx ? dx[c] - S.t[c] : dx[c];
  for (int i = tid; i < m; i += nt) {
    const T w = -S.hv[p + i];
    dz[i] = w;
    ds[i] = (-rs[i] - w) / S.d[i];
  }
  for (int j = tid; j < p; j += nt) dy[j] = -S.hv[j];
  __syncthreads();
}

// get_step pieces of one problem (batch.py:211-214): the min over entries the fill does NOT
// overwrite (NaN propagating, +inf if none), whether some entry is overwritten, and the
// NaN-propagating max of a = -v/dv over ALL entries (the batch-global fill candidate).
template <typename T>
__device__ __forceinline__ void step_pieces(const T* v, const T* dv, int m, int tid, int nt, T* red, T& rmu,
                                            bool& has_fill, T& amax) {
  T r[3] = {t_inf<T>(), T(0), -t_inf<T>()};
  for (int i = tid; i < m; i += nt) {
    const T dvi = dv[i];
    const T ai = -v[i] / dvi;
    if (dvi > T(0)) r[1] = T(1); else r[0] = nanmin(r[0], ai);
    r[2] = nanmax(r[2], ai);
  }
  // three different ops: do them one at a time (tiny)
  T a0[1] = {r[0]}, a1[1] = {r[1]}, a2[1] = {r[2]};
  block_reduce<1>(a0, OpNanMin(), red, tid, nt);
  block_reduce<1>(a1, OpSum(), red + 32, tid, nt);
  block_reduce<1>(a2, OpNanMax(), red + 64, tid, nt);
  rmu = a0[0];
  has_fill = a1[0] > T(0);
  amax = a2[0];
}

// ------------------------------------------------------------------------------------------
// One PDIPM iteration (a.iter >= 0) or the initial point (a.iter == -1).
template <typename T, bool SMEM, int NT>
__global__ void __launch_bounds__(NT, (!SMEM && sizeof(T) == 8) ? (NT == 256 ? 3 : 6) : 1) k_pdipm_iter(const KArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int prob = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = a.n, m = a.m, p = a.p, it = a.iter;
  Smem<T> S;
  carve<T, SMEM>(S, smem_raw, a, prob, NT);

  T fill_z = T(1), fill_s = T(1);
  if (it > 0) {
    // (a) has the reference already returned?  (b) fills of iteration it-1's final get_step
    if (warp == 0) {
      const int term = eval_termination(a.slots, it, a.lim, a.eps, lane);
      if (lane == 0) S.ctrl[0] = term;
    }
  }
  int flags = a.flags[prob];
  if (it < 0) flags = 0;
  if (!(flags & FLAG_POISON)) stage_mats<T, SMEM>(S, a, prob, tid, NT);
  if (it > 0) {
    __syncthreads();
    if (S.ctrl[0] >= 0) {
      if (SMEM) cp_async_wait_all();
      return;
    }
    const Slot* sl = a.slots + (it - 1);
    const double gz = sl->az_nan ? 0.0 : ord_unkey(sl->amax_z);
    const double gs = sl->as_nan ? 0.0 : ord_unkey(sl->amax_s);
    fill_z = gz > 1.0 ? (T)gz : T(1);
    fill_s = gs > 1.0 ? (T)gs : T(1);
  }
  Slot* slot = a.slots + (it >= 0 ? it : 0);

  if (flags & FLAG_POISON) {
    // NaN iterate: can never improve; still takes part in the batch-global reductions.
    if (it == 0) {  // poisoned by the initial point: the reference's outputs are all NaN
      T* bx = a.bx + (size_t)prob * n; T* bs = a.bs + (size_t)prob * m; T* bz = a.bz + (size_t)prob * m;
      for (int c = tid; c < n; c += NT) bx[c] = t_nan<T>();
      for (int i = tid; i < m; i += NT) { bs[i] = t_nan<T>(); bz[i] = t_nan<T>(); }
      if (p > 0) { T* by = a.by + (size_t)prob * p; for (int j = tid; j < p; j += NT) by[j] = t_nan<T>(); }
    }
    if (tid == 0) {
      double br = __longlong_as_double(0x7ff8000000000000LL);
      if (it == 0) a.best_resid[prob] = br; else br = a.best_resid[prob];
      if (br != br) slot->best_nan = 1; else atomic_max_key(&slot->best_max, (unsigned long long)__double_as_longlong(br));
      slot->mu_nan = 1;
      slot->az_nan = 1;
      slot->as_nan = 1;
    }
    return;
  }

  T* gx = a.x + (size_t)prob * round4(n);
  T* gs_ = a.s + (size_t)prob * round4(m);
  T* gz_ = a.z + (size_t)prob * round4(m);
  T* gy = a.y + (size_t)prob * round4(p > 0 ? p : 1);
  T* gdx = a.dx + (size_t)prob * round4(n);
  T* gds = a.ds + (size_t)prob * round4(m);
  T* gdz = a.dz + (size_t)prob * round4(m);
  T* gdy = a.dy + (size_t)prob * round4(p > 0 ? p : 1);
  const T* Qg = a.Q + (size_t)prob * a.sQ;
  const T* Gg = a.G + (size_t)prob * a.sG;
  const T* Ag = a.A + (size_t)prob * a.sA;
  const T* pg = a.pv + (size_t)prob * a.sp;
  const T* hg = a.h + (size_t)prob * a.sh;
  const T* bg = a.b + (size_t)prob * a.sb;

  if (it < 0) {
    // ---- initial point: d = 1, solve with (rx,rs,rz,ry) = (p, 0, -h, -b)   (batch.py:60-66)
    for (int i = tid; i < m; i += NT) { S.d[i] = T(1); S.rsc[i] = T(0); S.rz[i] = -hg[i]; }
    for (int c = tid; c < n; c += NT) S.rx[c] = pg[c];
    for (int j = tid; j < p; j += NT) S.ry[j] = -bg[j];
    __syncthreads();
    const bool ok = build_and_factor_T<T, SMEM, NT>(S, a, prob, tid, NT);
    if (!ok) {
      if (tid == 0) a.flags[prob] = FLAG_POISON;
      return;
    }
    kkt_solve(S, a, prob, S.rx, S.rsc, S.rz, p > 0 ? S.ry : (const T*)nullptr, S.x, S.s, S.z, S.y, tid, NT);
    // shift s and z so that their minimum is >= 1   (batch.py:76-86)
    T mn[2] = {t_inf<T>(), t_inf<T>()};
    for (int i = tid; i < m; i += NT) { mn[0] = nanmin(mn[0], S.s[i]); mn[1] = nanmin(mn[1], S.z[i]); }
    block_reduce<2>(mn, OpNanMin(), S.red, tid, NT);
    for (int i = tid; i < m; i += NT) {
      T sv = S.s[i], zv = S.z[i];
      if (mn[0] < T(0)) sv -= mn[0] - T(1);
      if (mn[1] < T(0)) zv -= mn[1] - T(1);
      gs_[i] = sv; gz_[i] = zv;
    }
    for (int c = tid; c < n; c += NT) gx[c] = S.x[c];
    for (int j = tid; j < p; j += NT) gy[j] = S.y[j];
    if (tid == 0) a.flags[prob] = 0;
    return;
  }

  // ---- load the iterate; apply the previous iteration's step (batch.py:190-204)
  {
    T alpha = T(0);
    if (it > 0) {
      const T rz_ = a.rmu[(size_t)prob * 2], rs_ = a.rmu[(size_t)prob * 2 + 1];
      const T stz = (flags & FLAG_FILL_Z) ? nanmin(rz_, fill_z) : rz_;
      const T sts = (flags & FLAG_FILL_S) ? nanmin(rs_, fill_s) : rs_;
      alpha = nanmin(T(0.999) * nanmin(stz, sts), T(1));
    }
    for (int c = tid; c < n; c += NT) { T v = gx[c]; if (it > 0) v += alpha * gdx[c]; S.x[c] = v; }
    for (int i = tid; i < m; i += NT) {
      T sv = gs_[i], zv = gz_[i];
      if (it > 0) { sv += alpha * gds[i]; zv += alpha * gdz[i]; }
      S.s[i] = sv; S.z[i] = zv;
    }
    for (int j = tid; j < p; j += NT) { T v = gy[j]; if (it > 0) v += alpha * gdy[j]; S.y[j] = v; }
  }
  __syncthreads();

  // ---- residuals (batch.py:93-108)
  // rx = Q x + p + G^T z + A^T y ; rz = G x + s - h ; ry = A x - b
  gemv_rows_warp(Qg, n, n, n, S.x, S.rx, tid, NT);
  gemv_rows_warp(Gg, n, m, n, S.x, S.rz, tid, NT);
  if (p > 0) gemv_rows_warp(Ag, n, p, n, S.x, S.ry, tid, NT);
  gemv_cols(Gg, n, m, n, S.z, S.t, S.part, tid, NT);  // G^T z -> t (barrier inside)
  if (p > 0) gemv_cols(Ag, n, p, n, S.y, S.dxc, S.part, tid, NT);
  T acc[4] = {T(0), T(0), T(0), T(0)};  // |rx|^2, |rz|^2, |ry|^2, s.z
  for (int c = tid; c < n; c += NT) {
    T v = (a.cb_cg ? a.cb_cg[(size_t)prob * n + c] : S.rx[c] + pg[c]) + S.t[c];
    if (p > 0) v += S.dxc[c];
    S.rx[c] = v;
    acc[0] += v * v;
  }
  for (int i = tid; i < m; i += NT) {
    const T sv = S.s[i], zv = S.z[i];
    const T v = S.rz[i] + sv - hg[i];
    S.rz[i] = v;
    acc[1] += v * v;
    acc[3] += sv * zv;
    S.d[i] = zv / sv;
  }
  for (int j = tid; j < p; j += NT) {
    const T v = a.cb_ry ? a.cb_ry[(size_t)prob * p + j] : S.ry[j] - bg[j];
    S.ry[j] = v;
    acc[2] += v * v;
  }
  block_reduce<4>(acc, OpSum(), S.red, tid, NT);
  const T mu = fabs(acc[3] / T(m));
  const T pri = (p > 0 ? sqrt(acc[2]) : T(0)) + sqrt(acc[1]);
  const T resid = pri + sqrt(acc[0]) + T(m) * mu;
  const T t4 = acc[3];

  // ---- factor T = R + diag(s/z)   (batch.py:110-114)
  const bool ok = build_and_factor_T<T, SMEM, NT>(S, a, prob, tid, NT);

  // ---- best-iterate bookkeeping + global reductions (batch.py:119-144)
  {
    const double rd = (double)resid;
    const double prev = a.best_resid[prob];
    const bool better = (it == 0) ? true : (rd < prev);
    __syncthreads();  // everyone has read best_resid before thread 0 may overwrite it
    if (better) {
      T* bx = a.bx + (size_t)prob * n; T* bs = a.bs + (size_t)prob * m; T* bz = a.bz + (size_t)prob * m;
      for (int c = tid; c < n; c += NT) bx[c] = S.x[c];
      for (int i = tid; i < m; i += NT) { bs[i] = S.s[i]; bz[i] = S.z[i]; }
      if (p > 0) { T* by = a.by + (size_t)prob * p; for (int j = tid; j < p; j += NT) by[j] = S.y[j]; }
    }
    if (tid == 0) {
      const double br = better ? rd : prev;
      if (better) a.best_resid[prob] = rd;
      if (better && it > 0 && !slot->improved) slot->improved = 1;
      if (br != br) slot->best_nan = 1; else atomic_max_key(&slot->best_max, (unsigned long long)__double_as_longlong(br));
      const double mud = (double)mu;
      if (mud != mud) slot->mu_nan = 1; else atomic_max_key(&slot->mu_min_inv, ~ord_key(mud));
    }
  }
  if (!ok) {
    if (tid == 0) {
      a.flags[prob] = FLAG_POISON;
      slot->az_nan = 1;
      slot->as_nan = 1;
    }
    return;
  }

  // ---- affine direction (batch.py:151-152):  rs = z
  kkt_solve(S, a, prob, S.rx, S.z, S.rz, p > 0 ? S.ry : (const T*)nullptr, S.dxa, S.dsa, S.dza, S.dya, tid, NT);

  // ---- centering (batch.py:161-169); the clamp at 1 makes alpha_aff independent of the fill
  T alpha_aff;
  {
    T rmz, rms, amz, ams; bool hz, hs;
    step_pieces(S.z, S.dza, m, tid, NT, S.red, rmz, hz, amz);
    step_pieces(S.s, S.dsa, m, tid, NT, S.red, rms, hs, ams);
    const T stz = hz ? nanmin(rmz, T(1)) : rmz;
    const T sts = hs ? nanmin(rms, T(1)) : rms;
    alpha_aff = nanmin(nanmin(stz, sts), T(1));
  }
  T t3v[1] = {T(0)};
  for (int i = tid; i < m; i += NT) t3v[0] += (S.s[i] + alpha_aff * S.dsa[i]) * (S.z[i] + alpha_aff * S.dza[i]);
  block_reduce<1>(t3v, OpSum(), S.red, tid, NT);
  const T ratio = t3v[0] / t4;
  const T sig = ratio * ratio * ratio;
  for (int i = tid; i < m; i += NT) S.rsc[i] = (-mu * sig + S.dsa[i] * S.dza[i]) / S.s[i];
  __syncthreads();

  // ---- corrector (batch.py:171-182): rx = rz = ry = 0
  kkt_solve(S, a, prob, (const T*)nullptr, S.rsc, (const T*)nullptr, (const T*)nullptr, S.dxc, S.dsc, S.dzc, S.dyc, tid, NT);

  // ---- combined direction; its get_step pieces; hand over to the next launch
  for (int c = tid; c < n; c += NT) { const T v = S.dxa[c] + S.dxc[c]; gdx[c] = v; }
  for (int i = tid; i < m; i += NT) {
    const T vs = S.dsa[i] + S.dsc[i], vz = S.dza[i] + S.dzc[i];
    S.dsa[i] = vs; S.dza[i] = vz;
    gds[i] = vs; gdz[i] = vz;
  }
  for (int j = tid; j < p; j += NT) gdy[j] = S.dya[j] + S.dyc[j];
  // the iterate itself is unchanged until the step is applied by the next launch
  if (it == 0) {
    // nothing: x,s,z,y already in global memory from the init kernel
  } else {
    for (int c = tid; c < n; c += NT) gx[c] = S.x[c];
    for (int i = tid; i < m; i += NT) { gs_[i] = S.s[i]; gz_[i] = S.z[i]; }
    for (int j = tid; j < p; j += NT) gy[j] = S.y[j];
  }
  __syncthreads();
  {
    T rmz, rms, amz, ams; bool hz, hs;
    step_pieces(S.z, S.dza, m, tid, NT, S.red, rmz, hz, amz);
    step_pieces(S.s, S.dsa, m, tid, NT, S.red, rms, hs, ams);
    if (tid == 0) {
      a.rmu[(size_t)prob * 2] = rmz;
      a.rmu[(size_t)prob * 2 + 1] = rms;
      a.flags[prob] = (hz ? FLAG_FILL_Z : 0) | (hs ? FLAG_FILL_S : 0);
      const double dz_ = (double)amz, ds_ = (double)ams;
      if (dz_ != dz_) slot->az_nan = 1; else atomic_max_key(&slot->amax_z, ord_key(dz_));
      if (ds_ != ds_) slot->as_nan = 1; else atomic_max_key(&slot->amax_s, ord_key(ds_));
    }
  }
}

// ------------------------------------------------------------------------------------------
// Caller-evaluated residual callbacks (this fork's dyn_res / cost_grad, batch.py:93-102): the callbacks need the
// iterate AFTER the step of iteration it-1, which the iteration kernels apply at their own start.  This kernel applies
// that step on its own (same termination test, same fill, same alpha as the head of k_pdipm_iter / k_fast_iter), zeroes
// the stored direction -- the iteration kernel that follows then adds alpha * 0 -- and hands x to the caller.
template <typename T>
__global__ void __launch_bounds__(128) k_cb_step(const KArgs<T> a, T* x_out) {
  __shared__ int s_term;
  const int prob = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NT = 128;
  const int n = a.n, m = a.m, p = a.p, it = a.iter;
  const int pp = p > 0 ? p : 1;
  T* gx = a.x + (size_t)prob * round4(n);
  if (tid == 0) s_term = -1;
  __syncthreads();
  if (it > 0 && warp == 0) {
    const int term = eval_termination(a.slots, it, a.lim, a.eps, lane);
    if (lane == 0) s_term = term;
  }
  __syncthreads();
  const int flags = a.flags[prob];
  if (it > 0 && s_term < 0 && !(flags & FLAG_POISON)) {
    const Slot* sl = a.slots + (it - 1);
    const double gz = sl->az_nan ? 0.0 : ord_unkey(sl->amax_z);
    const double gs = sl->as_nan ? 0.0 : ord_unkey(sl->amax_s);
    const T fill_z = gz > 1.0 ? (T)gz : T(1), fill_s = gs > 1.0 ? (T)gs : T(1);
    const T rz_ = a.rmu[(size_t)prob * 2], rs_ = a.rmu[(size_t)prob * 2 + 1];
    const T stz = (flags & FLAG_FILL_Z) ? nanmin(rz_, fill_z) : rz_;
    const T sts = (flags & FLAG_FILL_S) ? nanmin(rs_, fill_s) : rs_;
    const T alpha = nanmin(T(0.999) * nanmin(stz, sts), T(1));
    T* gs_ = a.s + (size_t)prob * round4(m);
    T* gz_ = a.z + (size_t)prob * round4(m);
    T* gy = a.y + (size_t)prob * round4(pp);
    T* gdx = a.dx + (size_t)prob * round4(n);
    T* gds = a.ds + (size_t)prob * round4(m);
    T* gdz = a.dz + (size_t)prob * round4(m);
    T* gdy = a.dy + (size_t)prob * round4(pp);
    for (int c = tid; c < n; c += NT) { gx[c] += alpha * gdx[c]; gdx[c] = T(0); }
    for (int i = tid; i < m; i += NT) { gs_[i] += alpha * gds[i]; gz_[i] += alpha * gdz[i]; gds[i] = T(0); gdz[i] = T(0); }
    for (int j = tid; j < p; j += NT) { gy[j] += alpha * gdy[j]; gdy[j] = T(0); }
  }
  __syncthreads();
  for (int c = tid; c < n; c += NT) x_out[(size_t)prob * n + c] = gx[c];
}

// ------------------------------------------------------------------------------------------
// Pre-factorisation (d-independent part).  Works in global memory (L1/L2 resident: one-off).
template <typename T, int NT>
__global__ void __launch_bounds__(NT) k_prefactor(const KArgs<T> a) {
  const int prob = blockIdx.x + a.prob0, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = a.n, m = a.m, p = a.p, ldn = a.ldn, ldm = a.ldm, ldp = a.ldp;
  const T* Qg = a.Q + (size_t)prob * a.sQ;
  const T* Gg = a.G + (size_t)prob * a.sG;
  const T* Ag = a.A + (size_t)prob * a.sA;
  T* F = a.F + (size_t)prob * a.sF;
  T* pinvF = a.pinvF + (size_t)prob * round4(n);
  T* Qi = a.Qi + (size_t)prob * a.sQi;
  T* BQi = a.BQi + (size_t)prob * a.sBQi;
  T* R = a.R + (size_t)prob * a.sR;
  T* V = a.V + (size_t)prob * a.sV;
  T* UA = a.UA + (size_t)prob * a.sUA;
  T* pinvA = a.pinvA + (size_t)prob * round4(p > 0 ? p : 1);

  // F = Q ; LDL^T ; Qi = Q^-1.  F and the working copy of Qi sit in shared memory when they fit
  // (a.pre_smem), else in the global workspace; the inversion is two CTA-parallel triangular
  // sweeps over the whole n x n identity (one barrier per column) instead of one thread per column.
  extern __shared__ __align__(16) unsigned char pre_smem_raw[];
  T* Fs = a.pre_smem ? reinterpret_cast<T*>(pre_smem_raw) : F;
  T* Qs = a.pre_smem ? Fs + round4(n * ldn) : Qi;
  T* pinvFs = a.pre_smem ? Qs + round4(n * ldn) : pinvF;
  for (int i = tid; i < n * n; i += NT) {
    const int r = i / n, c = i - r * n;
    Fs[(size_t)r * ldn + c] = Qg[i] + ((r == c) ? (T)a.reg : T(0));
    Qs[(size_t)r * ldn + c] = (r == c) ? T(1) : T(0);
  }
  __syncthreads();
  bool blocked = false;
  if constexpr (sizeof(T) == 8) blocked = a.pre_blocked != 0;
  bool okQ;
  if constexpr (sizeof(T) == 8) {
    if (blocked) {
      double* panel = reinterpret_cast<double*>(pre_smem_raw + (a.pre_blocked - 1));
      okQ = ldlt_factor_blocked(reinterpret_cast<double*>(Fs), ldn, n, reinterpret_cast<double*>(pinvFs), panel, tid, NT);
      ldlt_inverse_cols(reinterpret_cast<const double*>(Fs), ldn, n, reinterpret_cast<const double*>(pinvFs),
                        reinterpret_cast<double*>(Qs), tid, NT);
      __syncthreads();
    } else {
      okQ = ldlt_factor(Fs, ldn, n, pinvFs, tid, NT);
    }
  } else {
    okQ = ldlt_factor(Fs, ldn, n, pinvFs, tid, NT);
  }
  if (!okQ && tid == 0) atomicAdd(&a.ctl->q_fail, 1u);
  // L^-1: column j of L eliminates below row j; only columns c <= j of the identity are non-zero.
  // Lane = column, warp w takes rows j+1+w, j+1+w+nw, ...: no index divisions, conflict-free.
  if (!blocked) {
    const int nw = NT >> 5;
    for (int j = 0; j + 1 < n; j++) {
      for (int c = lane; c <= j; c += 32) {
        const T qjc = Qs[(size_t)j * ldn + c];
        for (int i = j + 1 + warp; i < n; i += nw) Qs[(size_t)i * ldn + c] -= Fs[(size_t)j * ldn + i] * qjc;
      }
      __syncthreads();
    }
    for (int k = warp; k < n; k += nw) {
      const T pk = pinvFs[k];
      for (int c = lane; c <= k; c += 32) Qs[(size_t)k * ldn + c] *= pk;
    }
    __syncthreads();
    // L^-T: row j eliminates above
    for (int j = n - 1; j > 0; j--) {
      for (int c = lane; c < n; c += 32) {
        const T qjc = Qs[(size_t)j * ldn + c];
        for (int i = warp; i < j; i += nw) Qs[(size_t)i * ldn + c] -= Fs[(size_t)i * ldn + j] * qjc;
      }
      __syncthreads();
    }
  }
  if (a.pre_smem) {
    for (int e = tid; e < n * n; e += NT) {
      const int r = e / n, c = e - r * n;
      Qi[(size_t)r * ldn + c] = Qs[(size_t)r * ldn + c];
    }
  }
  // BQi = [A;G] Qi  and  M = BQi [A;G]^T.  With a.pre_smem == 2 the stacked constraint matrix
  // B = [A;G] and BQi are staged in shared memory ((p+m) x ldn each, odd leading dimension =>
  // conflict-free row and column walks); otherwise operands come from global memory.
  const int pm = p + m;
  T* Bs = nullptr;
  T* BQs = BQi;
  if (a.pre_smem == 2) {
    Bs = pinvFs + round4(n);
    BQs = Bs + round4(pm * ldn);
    for (int e = tid; e < pm * n; e += NT) {
      const int r = e / n, c = e - r * n;
      Bs[(size_t)r * ldn + c] = r < p ? Ag[(size_t)r * n + c] : Gg[(size_t)(r - p) * n + c];
    }
    __syncthreads();
  }
  bool done = false;
  if constexpr (sizeof(T) == 8) {
    // fp64, no equalities, fragment-order R: both products on the FP64 tensor cores (m8n8k4), the
    // Schur block leaves the accumulators straight into R's fragment order (one 16-byte store per
    // lane per tile instead of an index computation per element).
    if (Bs != nullptr && p == 0 && a.rtile_nt < 0) {
      const int fr = lane >> 2, kc = lane & 3, fc = kc * 2, nw = NT >> 5;
      const int MT = (m + 7) >> 3, NTn = (n + 7) >> 3;
      const double* Bd = reinterpret_cast<const double*>(Bs);
      const double* Qd = reinterpret_cast<const double*>(Qs);
      double* BQd = reinterpret_cast<double*>(BQs);
      double* BQg = reinterpret_cast<double*>(BQi);
      double* Rd = reinterpret_cast<double*>(R);
      for (int t = warp; t < MT * NTn; t += nw) {
        const int I = t / NTn, J = t - I * NTn;
        const int row = 8 * I + fr, colb = 8 * J + fr;
        double c0 = 0.0, c1 = 0.0;
        for (int k0 = 0; k0 < n; k0 += 4) {
          const int k = k0 + kc;
          const double av = (row < m && k < n) ? Bd[(size_t)row * ldn + k] : 0.0;
          const double bv = (k < n && colb < n) ? Qd[(size_t)k * ldn + colb] : 0.0;
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                       : "+d"(c0), "+d"(c1) : "d"(av), "d"(bv));
        }
        const int col = 8 * J + fc;
        if (row < m) {
          if (col < n) { BQd[(size_t)row * ldn + col] = c0; BQg[(size_t)row * ldn + col] = c0; }
          if (col + 1 < n) { BQd[(size_t)row * ldn + col + 1] = c1; BQg[(size_t)row * ldn + col + 1] = c1; }
        }
      }
      __syncthreads();
      const int ntl = MT * (MT + 1) / 2;
      for (int t = warp; t < ntl; t += nw) {
        int I = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
        if ((I + 1) * (I + 2) / 2 <= t) I++;
        const int K = t - I * (I + 1) / 2;
        const int row = 8 * I + fr, rowk = 8 * K + fr;
        double c0 = 0.0, c1 = 0.0;
        for (int k0 = 0; k0 < n; k0 += 4) {
          const int k = k0 + kc;
          const double av = (row < m && k < n) ? BQd[(size_t)row * ldn + k] : 0.0;
          const double bv = (rowk < m && k < n) ? Bd[(size_t)rowk * ldn + k] : 0.0;
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                       : "+d"(c0), "+d"(c1) : "d"(av), "d"(bv));
        }
        *reinterpret_cast<double2*>(Rd + (size_t)t * 64 + lane * 2) = make_double2(c0, c1);
      }
      done = true;
    }
    if (!done && blocked) {
      // large nz: both products on the FP64 tensor cores with operands straight from global / shared
      // memory (16 x 16 output blocks per warp); the full M when there are equalities (V needs the
      // upper-right block), its lower triangle otherwise
      const double* Qd = reinterpret_cast<const double*>(Qs);
      double* BQg = reinterpret_cast<double*>(BQi);
      auto brow = [&](int r) -> const double* {
        return reinterpret_cast<const double*>(r < p ? Ag + (size_t)r * n : Gg + (size_t)(r - p) * n);
      };
      dmma_gemm(pm, n, n, false, brow, [&](int k, int c) { return Qd[(size_t)k * ldn + c]; },
                [&](int r, int c, double v) { BQg[(size_t)r * ldn + c] = v; }, tid, NT);
      __syncthreads();
      double* UAd = reinterpret_cast<double*>(UA);
      double* Vd = reinterpret_cast<double*>(V);
      double* Rd = reinterpret_cast<double*>(R);
      const double reg = a.reg;
      const int rt_mpad = a.rtile_mpad, rt_nt = a.rtile_nt;
      dmma_gemm(pm, pm, n, p == 0, [&](int r) -> const double* { return BQg + (size_t)r * ldn; },
                [&](int k, int c) { return brow(c)[k]; },
                [&](int r, int q, double v) {
                  if (r < p && q < p) { if (q <= r) UAd[(size_t)r * ldp + q] = v + ((r == q) ? reg : 0.0); }
                  else if (r < p) Vd[(size_t)r * ldm + (q - p)] = v;
                  else if (q >= p && q <= r)
                    Rd[rt_mpad ? (size_t)rtile_index(r - p, q - p, rt_mpad, rt_nt) : (size_t)(r - p) * ldm + (q - p)] = v;
                }, tid, NT);
      done = true;
    }
  }
  if (!done) {
    for (int e = tid; e < pm * n; e += NT) {
      const int r = e / n, c = e - r * n;
      const T* brow = Bs ? Bs + (size_t)r * ldn : (r < p ? Ag + (size_t)r * n : Gg + (size_t)(r - p) * n);
      T a0 = 0, a1 = 0;
      int k = 0;
      for (; k + 1 < n; k += 2) { a0 += brow[k] * Qs[(size_t)k * ldn + c]; a1 += brow[k + 1] * Qs[(size_t)(k + 1) * ldn + c]; }
      if (k < n) a0 += brow[k] * Qs[(size_t)k * ldn + c];
      const T v = a0 + a1;
      BQi[(size_t)r * ldn + c] = v;
      if (Bs) BQs[(size_t)r * ldn + c] = v;
    }
    __syncthreads();
    // M = BQi [A;G]^T  ->  Saa (UA, lower), Sag (V), Sgg (R, lower incl. diagonal)
    for (int e = tid; e < pm * pm; e += NT) {
      const int r = e / pm, q = e - r * pm;
      const bool need = (r < p && q < p && q <= r) || (r < p && q >= p) || (r >= p && q >= p && q <= r);
      if (!need) continue;
      const T* brow = Bs ? Bs + (size_t)q * ldn : (q < p ? Ag + (size_t)q * n : Gg + (size_t)(q - p) * n);
      const T* qrow = BQs + (size_t)r * ldn;
      T a0 = 0, a1 = 0;
      int k = 0;
      for (; k + 1 < n; k += 2) { a0 += qrow[k] * brow[k]; a1 += qrow[k + 1] * brow[k + 1]; }
      if (k < n) a0 += qrow[k] * brow[k];
      const T v = a0 + a1;
      if (r < p && q < p) UA[(size_t)r * ldp + q] = v + ((r == q) ? (T)a.reg : T(0));
      else if (r < p) V[(size_t)r * ldm + (q - p)] = v;
      else R[a.rtile_mpad ? (size_t)rtile_index(r - p, q - p, a.rtile_mpad, a.rtile_nt) : (size_t)(r - p) * ldm + (q - p)] = v;
    }
  }
  __syncthreads();
  if (p > 0) {
    const bool okA = ldlt_factor(UA, ldp, p, pinvA, tid, NT);
    if (!okA && tid == 0) atomicAdd(&a.ctl->aqa_fail, 1u);
    // V <- La^-1 Sag  (column i of V by thread i)
    for (int i = tid; i < m; i += NT) {
      for (int j = 0; j < p; j++) {
        const T vj = V[(size_t)j * ldm + i];
        const T* row = UA + (size_t)j * ldp;
        for (int k = j + 1; k < p; k++) V[(size_t)k * ldm + i] -= row[k] * vj;
      }
    }
    __syncthreads();
    // R -= V^T Da^-1 V  (lower triangle)
    for (int e = tid; e < m * m; e += NT) {
      const int r = e / m, q = e - r * m;
      if (q > r) continue;
      T acc = 0;
      for (int j = 0; j < p; j++) acc += V[(size_t)j * ldm + r] * pinvA[j] * V[(size_t)j * ldm + q];
      R[a.rtile_mpad ? (size_t)rtile_index(r, q, a.rtile_mpad, a.rtile_nt) : (size_t)r * ldm + q] -= acc;
    }
  }
  // mirror R's lower triangle so that the staged copy is symmetric (only the lower is used)
  (void)lane; (void)warp;
}

// ------------------------------------------------------------------------------------------
// Backward (qp.py:129-183): d = clamp(lam)/clamp(s); one factor + one solve; outer products.
template <typename T>
struct BArgs {
  const T *zhat, *lams, *nus, *slacks, *gz;  // best iterate + dl/dzhat
  T *dQ, *dp, *dG, *dh, *dA, *db;
};

template <typename T, bool SMEM, int NT>
__global__ void __launch_bounds__(NT) k_backward(const KArgs<T> a, const BArgs<T> g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int prob = blockIdx.x + a.prob0, tid = threadIdx.x;
  const int n = a.n, m = a.m, p = a.p;
  Smem<T> S;
  carve<T, SMEM>(S, smem_raw, a, prob, NT);
  stage_mats<T, SMEM>(S, a, prob, tid, NT);
  const T* zh = g.zhat + (size_t)prob * n;
  const T* lam = g.lams + (size_t)prob * m;
  const T* sl = g.slacks + (size_t)prob * m;
  const T* nu = g.nus + (size_t)prob * p;
  const T* gz = g.gz + (size_t)prob * n;
  for (int i = tid; i < m; i += NT) {
    const T lv = lam[i], sv = sl[i];
    S.z[i] = lv;  // lams
    // torch.clamp(x, min=c): NaN stays NaN
    const T lc = (lv < T(1e-8)) ? T(1e-8) : lv, sc = (sv < T(1e-8)) ? T(1e-8) : sv;
    S.d[i] = lc / sc;
    S.rsc[i] = T(0);
  }
  for (int c = tid; c < n; c += NT) { S.rx[c] = gz[c]; S.x[c] = zh[c]; }
  for (int j = tid; j < p; j += NT) S.y[j] = nu[j];
  __syncthreads();
  const bool ok = build_and_factor_T<T, SMEM, NT>(S, a, prob, tid, NT);
  if (!ok) {
    for (int i = tid; i < m * a.ldm; i += NT) S.Tm[i] = t_nan<T>();
    for (int i = tid; i < m; i += NT) S.pinvT[i] = t_nan<T>();
    __syncthreads();
  }
  kkt_solve(S, a, prob, S.rx, S.rsc, (const T*)nullptr, (const T*)nullptr, S.dxa, S.dsa, S.dza, S.dya, tid, NT);
  // dp = dx ; dh = -dlam ; db = -dnu
  T* dp = g.dp + (size_t)prob * n; T* dh = g.dh + (size_t)prob * m;
  for (int c = tid; c < n; c += NT) dp[c] = S.dxa[c];
  for (int i = tid; i < m; i += NT) dh[i] = -S.dza[i];
  if (p > 0) { T* db = g.db + (size_t)prob * p; for (int j = tid; j < p; j += NT) db[j] = -S.dya[j]; }
  // dQ = 0.5 (dx zhat^T + zhat dx^T)
  T* dQ = g.dQ + (size_t)prob * n * n;
  for (int e = tid; e < n * n; e += NT) {
    const int r = e / n, c = e - r * n;
    dQ[e] = T(0.5) * (S.dxa[r] * S.x[c] + S.x[r] * S.dxa[c]);
  }
  // dG = dlam zhat^T + lam dx^T
  T* dG = g.dG + (size_t)prob * m * n;
  for (int e = tid; e < m * n; e += NT) {
    const int r = e / n, c = e - r * n;
    dG[e] = S.dza[r] * S.x[c] + S.z[r] * S.dxa[c];
  }
  if (p > 0) {
    T* dA = g.dA + (size_t)prob * p * n;
    for (int e = tid; e < p * n; e += NT) {
      const int r = e / n, c = e - r * n;
      dA[e] = S.dya[r] * S.x[c] + S.y[r] * S.dxa[c];
    }
  }
}

// ------------------------------------------------------------------------------------------
// Stand-alone KKT solve with caller-provided d and right-hand sides (batch.py:351-374 + :434-469).
template <typename T>
struct SArgs {
  const T *d, *rx, *rs, *rz, *ry;
  T *dx, *ds, *dz, *dy;
};

template <typename T, bool SMEM, int NT>
__global__ void __launch_bounds__(NT) k_kkt_solve(const KArgs<T> a, const SArgs<T> g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int prob = blockIdx.x, tid = threadIdx.x;
  const int n = a.n, m = a.m, p = a.p;
  Smem<T> S;
  carve<T, SMEM>(S, smem_raw, a, prob, NT);
  stage_mats<T, SMEM>(S, a, prob, tid, NT);
  for (int i = tid; i < m; i += NT) {
    S.d[i] = g.d[(size_t)prob * m + i];
    S.rsc[i] = g.rs[(size_t)prob * m + i];
    S.rz[i] = g.rz[(size_t)prob * m + i];
  }
  for (int c = tid; c < n; c += NT) S.rx[c] = g.rx[(size_t)prob * n + c];
  for (int j = tid; j < p; j += NT) S.ry[j] = g.ry[(size_t)prob * p + j];
  __syncthreads();
  const bool ok = build_and_factor_T<T, SMEM, NT>(S, a, prob, tid, NT);
  if (!ok) {
    for (int i = tid; i < m; i += NT) S.pinvT[i] = t_nan<T>();
    __syncthreads();
  }
  kkt_solve(S, a, prob, S.rx, S.rsc, S.rz, p > 0 ? S.ry : (const T*)nullptr, S.dxa, S.dsa, S.dza, S.dya, tid, NT);
  for (int c = tid; c < n; c += NT) g.dx[(size_t)prob * n + c] = S.dxa[c];
  for (int i = tid; i < m; i += NT) { g.ds[(size_t)prob * m + i] = S.dsa[i]; g.dz[(size_t)prob * m + i] = S.dza[i]; }
  for (int j = tid; j < p; j += NT) g.dy[(size_t)prob * p + j] = S.dya[j];
}

}  // namespace b200qp
