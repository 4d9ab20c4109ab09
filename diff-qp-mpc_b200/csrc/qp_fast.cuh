// qp_fast.cuh -- the shared-memory/register resident fast path of the PDIPM kernels:
//   nineq <= 64   ONE WARP per QP (32-thread CTA): no CTA barriers at all, the whole trailing
//                 matrix of the LDL^T (72 doubles per lane at MPAD=64) lives in registers;
//   nineq <= 128  256-thread CTA per QP.
//
// Differences from the generic kernels in qp_kernels.cuh (same algebra, same outputs):
//   * R = G Q^-1 G^T (Schur-corrected) is stored by the pre-factorisation in REGISTER-TILE ORDER
//     (block q, thread t -> one element), so the iteration kernel loads T = R + diag(s/z)
//     straight into the registers of the 2-D cyclic LDL^T with perfectly coalesced loads and no
//     shared-memory staging;
//   * the factor is kept PACKED (strict upper triangle of U = L^T, row-major) => 14.6 KB instead
//     of 29 KB at nineq=60, which is what lets 5-6 CTAs share an SM;
//   * one call site for the KKT solve (predictor/corrector are two trips through the same
//     code: smaller instruction footprint, the kernel was instruction-fetch bound), step-length
//     pieces and centering are done by one warp with shuffles instead of CTA-wide reductions.
#pragma once
#include "qp_kernels.cuh"

#ifndef FAST_DMMA_MIN_CTAS
#define FAST_DMMA_MIN_CTAS 4
#endif

namespace b200qp {

template <int MPAD, int NT>
struct Tile {
  static constexpr int TR = NT == 32 ? 8 : 16, TC = NT / TR, NA = MPAD / TR, NB = MPAD / TC;
  __host__ __device__ static constexpr bool alive(int a, int b) { return TC * b <= TR * a + TR - 1; }
  __host__ __device__ static constexpr int nblk() {
    int c = 0;
    for (int a = 0; a < NA; a++)
      for (int b = 0; b < NB; b++)
        if (alive(a, b)) c++;
    return c;
  }
};

// 1/x for a positive pivot: MUFU seed + two Newton steps (5 instructions instead of the ~15 of an
// IEEE division; <= 1 ulp).  Falls back to the division outside the safe exponent range.
__device__ __forceinline__ double pivot_rcp(double x) {
  if (x > 1e-290 && x < 1e290) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
  }
  return 1.0 / x;
}
__device__ __forceinline__ float pivot_rcp(float x) { return 1.0f / x; }

}  // namespace b200qp
#include "qp_dmma.cuh"
namespace b200qp {

template <typename T>
struct FS {  // shared-memory carve-up of the fast path
  T *Up, *pinvT, *BQi, *V, *UA, *pinvA;
  T *x, *s, *z, *y, *d, *rx, *rz, *ry, *t, *hv, *u, *dx, *ds, *dz, *dy, *rsc, *scr, *scrn, *part, *red, *colbuf, *small;
  T *ax, *as, *az, *ay, *r2x, *r2s, *r2z, *r2y;  // dense mode: saved predictor solution, refinement right-hand side
};

// fk = 1: DMMA factorisation (qp_dmma.cuh): the packed factor carries one bordered column and the
// column buffer becomes the [mpad][10] panel buffer (+8 panel pivots).
__host__ __device__ inline int fast_up_elems(int m, int fk) { return fk ? (m + 1) * m / 2 + 1 : m * (m - 1) / 2 + 1; }
__host__ __device__ inline int fast_colbuf_elems(int mpad, int fk) { return fk ? mpad * 10 + 8 : 2 * mpad; }
__host__ __device__ inline size_t fast_smem_elems(int n, int m, int p, int ldn, int ldm, int ldp, int nt, int mpad, int fk = 0) {
  const int pp = p > 0 ? p : 1;
  size_t e = round4(fast_up_elems(m, fk)) + round4(m) + round4((p + m) * ldn);
  if (p > 0) e += round4(p * ldm) + round4(p * ldp) + round4(p);
  e += (size_t)5 * round4(n) + (size_t)8 * round4(m) + (size_t)4 * round4(pp) + round4(p + m);
  e += round4(nt) + 4 * 32 + round4(fast_colbuf_elems(mpad, fk)) + 16;
  e += (size_t)2 * round4(n) + (size_t)4 * round4(m) + (size_t)2 * round4(pp);  // dense-mode vectors
  return e;
}

// Element offsets of the carve-up, computed ONCE on the host (KArgs::fso) so that the kernels form
// each shared-memory pointer with one constant-bank load instead of re-deriving the whole chain
// of rounded sizes (ncu: ~1.9 k instructions per problem-iteration went into that).
constexpr int kFsoCount = 36;
static_assert(sizeof(((KArgs<double>*)nullptr)->fso) / sizeof(int) >= kFsoCount, "KArgs::fso is too small for the carve-up");
__host__ __device__ inline void fast_offsets(int n, int m, int p, int ldn, int ldm, int ldp, int nt, int mpad, int fk,
                                             int* o) {
  int q = 0, i = 0;
  auto take = [&](int cnt) { o[i++] = q; q += round4(cnt); };
  const int pp = p > 0 ? p : 1;
  take(fast_up_elems(m, fk));           // 0 Up
  take(m);                              // 1 pinvT
  take((p + m) * ldn);                  // 2 BQi
  if (p > 0) { take(p * ldm); take(p * ldp); take(p); } else { o[i++] = 0; o[i++] = 0; o[i++] = 0; }  // 3 V 4 UA 5 pinvA
  take(n); take(n); take(n); take(n); take(n);                                   // 6 x 7 rx 8 t 9 dx 10 scrn
  take(m); take(m); take(m); take(m); take(m); take(m); take(m); take(m);        // 11 s 12 z 13 d 14 rz 15 ds 16 dz 17 rsc 18 scr
  take(pp); take(pp); take(pp); take(pp);                                        // 19 y 20 ry 21 u 22 dy
  take(p + m);                          // 23 hv
  take(nt);                             // 24 part
  take(4 * 32);                         // 25 red
  take(fast_colbuf_elems(mpad, fk));    // 26 colbuf
  take(16);                             // 27 small
  take(n); take(m); take(m); take(pp);  // 28 ax 29 as 30 az 31 ay
  take(n); take(m); take(m); take(pp);  // 32 r2x 33 r2s 34 r2z 35 r2y
}

template <typename T>
__device__ __forceinline__ void fast_carve(FS<T>& S, unsigned char* raw, const KArgs<T>& a) {
  T* q = reinterpret_cast<T*>(raw);
  const int* o = a.fso;
  S.Up = q + o[0]; S.pinvT = q + o[1]; S.BQi = q + o[2];
  if (a.p > 0) { S.V = q + o[3]; S.UA = q + o[4]; S.pinvA = q + o[5]; }
  else { S.V = S.UA = S.pinvA = nullptr; }
  S.x = q + o[6]; S.rx = q + o[7]; S.t = q + o[8]; S.dx = q + o[9]; S.scrn = q + o[10];
  S.s = q + o[11]; S.z = q + o[12]; S.d = q + o[13]; S.rz = q + o[14]; S.ds = q + o[15]; S.dz = q + o[16];
  S.rsc = q + o[17]; S.scr = q + o[18];
  S.y = q + o[19]; S.ry = q + o[20]; S.u = q + o[21]; S.dy = q + o[22];
  S.hv = q + o[23]; S.part = q + o[24]; S.red = q + o[25]; S.colbuf = q + o[26]; S.small = q + o[27];
  S.ax = q + o[28]; S.as = q + o[29]; S.az = q + o[30]; S.ay = q + o[31];
  S.r2x = q + o[32]; S.r2s = q + o[33]; S.r2z = q + o[34]; S.r2y = q + o[35];
}

template <typename T>
__device__ __forceinline__ void fast_stage(const FS<T>& S, const KArgs<T>& a, int prob, int tid, int nt) {
  cp_async_block(S.BQi, a.BQi + (size_t)prob * a.sBQi, round4((a.p + a.m) * a.ldn), tid, nt);
  if (a.p > 0) {
    cp_async_block(S.V, a.V + (size_t)prob * a.sV, round4(a.p * a.ldm), tid, nt);
    cp_async_block(S.UA, a.UA + (size_t)prob * a.sUA, round4(a.p * a.ldp), tid, nt);
    cp_async_block(S.pinvA, a.pinvA + (size_t)prob * round4(a.p), round4(a.p), tid, nt);
  }
  cp_async_commit();
}

// T = R + diag(1/d) = L D L^T with the trailing matrix in registers (see ldlt_factor_reg in
// qp_common.cuh for the scheme); R comes from global memory in tile order, U goes to shared
// memory packed.  Returns false (uniformly) on a non-positive / NaN pivot.
template <typename T, int MPAD, int NT>
__device__ __forceinline__ bool fast_factor(const T* __restrict__ Rt, const T* d, T* Up, T* pinv, T* colbuf, int m,
                                            int tid) {
  using TL = Tile<MPAD, NT>;
  constexpr int TR = TL::TR, TC = TL::TC, NA = TL::NA, NB = TL::NB;
  const int ti = tid % TR, tk = tid / TR;
  T A[NA][NB];
  {
    int q = 0;
#pragma unroll
    for (int a = 0; a < NA; a++) {
#pragma unroll
      for (int b = 0; b < NB; b++) {
        if (TL::alive(a, b)) {
          const int i = ti + TR * a, k = tk + TC * b;
          T v = (i < m && k <= i) ? Rt[q * NT + tid] : T(0);
          if (i == k && i < m) v += T(1) / d[i];
          A[a][b] = v;
          q++;
        }
      }
    }
  }
  if (tk == 0) {
#pragma unroll
    for (int a = 0; a < NA; a++) {
      const int i = ti + TR * a;
      if (i < m) colbuf[i] = A[a][0];
    }
    if (ti == 0) pinv[0] = (A[0][0] > T(0)) ? pivot_rcp(A[0][0]) : t_nan<T>();
  }
  cta_sync<NT>();
  bool ok = true;
  int base = 0;  // packed offset of row j of U
#pragma unroll
  for (int jb = 0; jb < NB; jb++) {
    for (int tkk = 0; tkk < TC; tkk++) {
      const int j = jb * TC + tkk;
      if (j >= m || !ok) break;  // uniform
      const T* cb = colbuf + (j & 1) * MPAD;
      T* cbn = colbuf + ((j + 1) & 1) * MPAD;
      const T pj = pinv[j];  // published by the owner of the diagonal element one step earlier
      if (is_nan(pj)) { ok = false; break; }  // uniform: same shared word for every thread
      T li[NA], ck[NB];
      // no predicates: rows/cols that are already eliminated (or >= m) only feed dead registers
#pragma unroll
      for (int a = 0; a < NA; a++) li[a] = cb[ti + TR * a] * pj;
#pragma unroll
      for (int b = jb; b < NB; b++) ck[b] = cb[tk + TC * b];
      if (tk == tkk) {
#pragma unroll
        for (int a = 0; a < NA; a++) {
          const int i = ti + TR * a;
          if (i > j && i < m) Up[base + i - j - 1] = li[a];
        }
      }
#pragma unroll
      for (int a = 0; a < NA; a++) {
        if (TR * a + TR - 1 > j) {  // uniform: block row still alive
#pragma unroll
          for (int b = jb; b < NB; b++) {
            if (TL::alive(a, b)) A[a][b] -= li[a] * ck[b];
          }
        }
      }
      // publish column j+1 and the reciprocal of its pivot
      const int jn = j + 1;
      if (tkk < TC - 1) {
        if (tk == tkk + 1) {
#pragma unroll
          for (int a = 0; a < NA; a++) {
            if (TL::alive(a, jb)) {
              const int i = ti + TR * a;
              if (i >= jn && i < m) {
                const T v = A[a][jb];
                cbn[i] = v;
                if (i == jn) pinv[jn] = (v > T(0)) ? pivot_rcp(v) : t_nan<T>();
              }
            }
          }
        }
      } else if (jb + 1 < NB) {
        if (tk == 0) {
#pragma unroll
          for (int a = 0; a < NA; a++) {
            if (TL::alive(a, jb + 1 < NB ? jb + 1 : jb)) {
              const int i = ti + TR * a;
              if (i >= jn && i < m) {
                const T v = A[a][jb + 1 < NB ? jb + 1 : jb];
                cbn[i] = v;
                if (i == jn) pinv[jn] = (v > T(0)) ? pivot_rcp(v) : t_nan<T>();
              }
            }
          }
        }
      }
      base += m - 1 - j;
      cta_sync<NT>();
    }
  }
  if (!ok) {
    for (int i = tid; i < m; i += NT) pinv[i] = t_nan<T>();
    for (int i = tid; i < m * (m - 1) / 2; i += NT) Up[i] = t_nan<T>();
    cta_sync<NT>();
  }
  return ok;
}

// (L D L^T) x = v with packed unit-upper U; one warp, working vector in registers.
template <typename T, int RPL>
__device__ __forceinline__ void fast_ldlt_solve(const T* Up, int m, const T* pinv, T* v, int lane) {
  T r[RPL];
  int rb[RPL];  // packed base of row i for the backward sweep
#pragma unroll
  for (int s = 0; s < RPL; s++) {
    const int i = s * 32 + lane;
    r[s] = i < m ? v[i] : T(0);
    rb[s] = i * (m - 1) - (i * (i - 1)) / 2 - i - 1;  // + j  ->  index of U[i][j]
  }
  int base = 0;
#pragma unroll
  for (int s = 0; s < RPL; s++) {
    const int jend = min(32, m - s * 32);
    for (int jj = 0; jj < jend; jj++) {
      const int j = s * 32 + jj;
      const T yj = shfl_d(r[s], jj);
      const T* row = Up + base - j - 1;  // row[i] = U[j][i]
#pragma unroll
      for (int s2 = s; s2 < RPL; s2++) {
        const int i = s2 * 32 + lane;
        if (i > j && i < m) r[s2] -= row[i] * yj;
      }
      base += m - 1 - j;
    }
  }
#pragma unroll
  for (int s = 0; s < RPL; s++) {
    const int i = s * 32 + lane;
    if (i < m) r[s] *= pinv[i];
  }
#pragma unroll
  for (int s = RPL - 1; s >= 0; s--) {
    const int jend = min(32, m - s * 32);
    for (int jj = jend - 1; jj >= 0; jj--) {
      const int j = s * 32 + jj;
      const T xj = shfl_d(r[s], jj);
#pragma unroll
      for (int s2 = 0; s2 <= s; s2++) {
        const int i = s2 * 32 + lane;
        if (i < j) r[s2] -= Up[rb[s2] + j] * xj;
      }
    }
  }
#pragma unroll
  for (int s = 0; s < RPL; s++) {
    const int i = s * 32 + lane;
    if (i < m) v[i] = r[s];
  }
}

// Block-elimination KKT solve, fast path, in three stages so that the predictor's right-hand side
// can ride through the factorisation (qp_dmma.cuh):
//   fast_kkt_pre   hv[p..p+m) <- hz - V^T Da^-1 u  (ready for T^-1), u, t = Qi rx
//   (T^-1 hv by one warp)
//   fast_kkt_post  dz = -qz, dy = -qy, ds = (-rs - dz)/d, dx = -t + BQi^T [qy;qz]
// has_rx=false means rx = rz = ry = 0 (corrector).  accumulate=true adds the solution to
// dx/ds/dz/dy instead of overwriting.  gdx != nullptr additionally streams the final
// dx/ds/dz/dy to global memory.  Each stage ends with a barrier.
template <typename T, int NT>
__device__ __forceinline__ void fast_kkt_pre(const FS<T>& S, const KArgs<T>& a, int prob, bool has_rx, const T* rx,
                                             const T* rs, const T* rz, const T* ry, int tid,
                                             const RegMat<8, T>* qi_regs = nullptr) {
  const int n = a.n, m = a.m, p = a.p, ldn = a.ldn, ldm = a.ldm, ldp = a.ldp;
  const int lane = tid & 31, warp = tid >> 5;
  if (has_rx) {
    gemv_rows_thread(S.BQi, ldn, p + m, n, rx, S.hv, tid, NT);
    if (qi_regs != nullptr) {  // t = Qi^T rx from the register-resident copy (NT = 128, n <= 32)
      regmat_cols<8, T>(*qi_regs, rx, S.part, n, lane, warp);
      cta_sync<NT>();
      if (tid < n) S.t[tid] = regmat_colsum(S.part, tid);
      cta_sync<NT>();
    } else {
      gemv_cols_nt<T, NT>(a.Qi + (size_t)prob * a.sQi, ldn, n, n, rx, S.t, S.part, tid);  // t = Qi rx
    }
  }
  for (int i = tid; i < m; i += NT) {
    T v = rs[i] / S.d[i];
    if (has_rx) v += S.hv[p + i] - rz[i];
    S.hv[p + i] = v;
  }
  for (int j = tid; j < p; j += NT) S.u[j] = has_rx ? S.hv[j] - ry[j] : T(0);
  cta_sync<NT>();
  if (p > 0) {
    if (warp == 0) {
      unit_fwd_warp(S.UA, ldp, p, S.u, lane);
      for (int j = lane; j < p; j += 32) S.hv[j] = S.u[j] * S.pinvA[j];
    }
    cta_sync<NT>();
    for (int i = tid; i < m; i += NT) {  // hz -= V^T (Da^-1 u)
      T acc = T(0);
      for (int j = 0; j < p; j++) acc += S.V[(size_t)j * ldm + i] * S.hv[j];
      S.hv[p + i] -= acc;
    }
    cta_sync<NT>();
  }
}

template <typename T, int NT>
__device__ __forceinline__ void fast_kkt_post(const FS<T>& S, const KArgs<T>& a, int prob, bool has_rx, const T* rs,
                                              bool accumulate, T* gdx, T* gds, T* gdz, T* gdy, int tid) {
  const int n = a.n, m = a.m, p = a.p, ldn = a.ldn, ldm = a.ldm, ldp = a.ldp;
  const int lane = tid & 31, warp = tid >> 5;
  if (p > 0) {
    gemv_rows_thread(S.V, ldm, p, m, S.hv + p, S.hv, tid, NT);
    cta_sync<NT>();
    if (warp == 0) {
      for (int j = lane; j < p; j += 32) S.hv[j] = (S.u[j] - S.hv[j]) * S.pinvA[j];
      __syncwarp();
      unit_bwd_warp(S.UA, ldp, p, S.hv, lane);
    }
    cta_sync<NT>();
  }
  gemv_cols_nt<T, NT>(S.BQi, ldn, p + m, n, S.hv, S.scrn, S.part, tid);  // BQi^T q
  for (int c = tid; c < n; c += NT) {
    T v = S.scrn[c];
    if (has_rx) v -= S.t[c];
    if (accumulate) v += S.dx[c];
    S.dx[c] = v;
    if (gdx) gdx[c] = v;
  }
  for (int i = tid; i < m; i += NT) {
    const T w = -S.hv[p + i];
    const T dsv = (-rs[i] - w) / S.d[i];
    const T nz_ = accumulate ? S.dz[i] + w : w;
    const T ns_ = accumulate ? S.ds[i] + dsv : dsv;
    S.dz[i] = nz_; S.ds[i] = ns_;
    if (gdz) { gdz[i] = nz_; gds[i] = ns_; }
  }
  for (int j = tid; j < p; j += NT) {
    const T v = accumulate ? S.dy[j] - S.hv[j] : -S.hv[j];
    S.dy[j] = v;
    if (gdy) gdy[j] = v;
  }
  cta_sync<NT>();
}

// Factor-kind dispatch.  FK = 0: register-tile rank-1 LDL^T (fast_factor); FK = 1: DMMA blocked
// LDL^T (qp_dmma.cuh; double, NT = 128, m < MPAD).  hz (FK = 1 only): right-hand side carried as a
// bordered row, see dmma_factor.
template <typename T, int MPAD, int NT, int FK>
__device__ __forceinline__ bool factor_any(const T* __restrict__ Rt, const FS<T>& S, const T* hz, int m, int tid,
                                           const DmmaTiles<MPAD, NT / 32>* pre = nullptr) {
  if constexpr (FK == 1) {
    return dmma_factor<MPAD, NT / 32>(Rt, S.scr, hz, S.Up, S.pinvT, S.colbuf, m, tid, pre);  // S.scr = 1/d, filled by the caller
  } else {
    return fast_factor<T, MPAD, NT>(Rt, S.d, S.Up, S.pinvT, S.colbuf, m, tid);
  }
}
// v <- T^-1 v by ONE warp (from_border: v <- L^-T of the bordered row, FK = 1 only)
template <typename T, int MPAD, int FK>
__device__ __forceinline__ void tri_solve_any(const FS<T>& S, int m, T* v, bool from_border, int lane) {
  if constexpr (FK == 1) dmma_ldlt_solve<(MPAD + 31) / 32>(S.Up, m, m + 1, S.pinvT, v, from_border, lane);
  else fast_ldlt_solve<T, (MPAD + 31) / 32>(S.Up, m, S.pinvT, v, lane);
}

template <typename T, int MPAD, int NT, int FK = 0>
__device__ __forceinline__ void fast_kkt_solve(const FS<T>& S, const KArgs<T>& a, int prob, bool has_rx, const T* rx,
                                               const T* rs, const T* rz, const T* ry, bool accumulate, T* gdx, T* gds,
                                               T* gdz, T* gdy, int tid) {
  fast_kkt_pre<T, NT>(S, a, prob, has_rx, rx, rs, rz, ry, tid);
  if ((tid >> 5) == 0) tri_solve_any<T, MPAD, FK>(S, a.m, S.hv + a.p, false, tid & 31);
  cta_sync<NT>();
  fast_kkt_post<T, NT>(S, a, prob, has_rx, rs, accumulate, gdx, gds, gdz, gdy, tid);
}

// DenseQPFunction's iterative refinement (batch_LU.py:228-236): given the solution l0 =
// (S.dx, S.ds, S.dz, S.dy) of the REGULARISED system for the right-hand side (rx, rs, rz, ry)
// [rs in the kernels' scaled form rs_ref / s], form r2 = K l0 + (rx, rs, rz, ry) with the
// UNREGULARISED K (s-row z ds + s dz), solve the regularised system for it and add: l = l0 + d.
// has_rx=false: rx = rz = ry = 0.  Entry: l0 visible.  Exit: refined solution visible (barrier).
template <typename T, int MPAD, int NT, int FK>
__device__ __forceinline__ void dense_refine(const FS<T>& S, const KArgs<T>& a, int prob, bool has_rx, const T* rx,
                                             const T* rs, const T* rz, const T* ry, const T* zv, const T* sv, int tid) {
  const int n = a.n, m = a.m, p = a.p;
  const T* Qg = a.Q + (size_t)prob * a.sQ;
  const T* Gg = a.G + (size_t)prob * a.sG;
  const T* Ag = a.A + (size_t)prob * a.sA;
  gemv_rows_warp(Qg, n, n, n, S.dx, S.r2x, tid, NT);            // Q dx
  gemv_rows_warp(Gg, n, m, n, S.dx, S.r2z, tid, NT);            // G dx
  if (p > 0) gemv_rows_warp(Ag, n, p, n, S.dx, S.r2y, tid, NT);  // A dx
  gemv_cols_nt<T, NT>(Gg, n, m, n, S.dz, S.t, S.part, tid);     // G^T dz  (barrier inside)
  if (p > 0) gemv_cols_nt<T, NT>(Ag, n, p, n, S.dy, S.scrn, S.part, tid);
  for (int c = tid; c < n; c += NT) {
    T v = S.r2x[c] + S.t[c];
    if (p > 0) v += S.scrn[c];
    if (has_rx) v += rx[c];
    S.r2x[c] = v;
  }
  for (int i = tid; i < m; i += NT) {
    S.r2s[i] = rs[i] + (zv[i] / sv[i]) * S.ds[i] + S.dz[i];
    T v = S.r2z[i] + S.ds[i];
    if (has_rx) v += rz[i];
    S.r2z[i] = v;
  }
  for (int j = tid; j < p; j += NT) {
    T v = S.r2y[j];
    if (has_rx) v += ry[j];
    S.r2y[j] = v;
  }
  cta_sync<NT>();
  fast_kkt_pre<T, NT>(S, a, prob, true, S.r2x, S.r2s, S.r2z, S.r2y, tid);
  if ((tid >> 5) == 0) tri_solve_any<T, MPAD, FK>(S, m, S.hv + p, false, tid & 31);
  cta_sync<NT>();
  fast_kkt_post<T, NT>(S, a, prob, true, S.r2s, true, (T*)nullptr, (T*)nullptr, (T*)nullptr, (T*)nullptr, tid);
}

// get_step pieces for (z,dz) and (s,ds) at once, by ONE warp: out = {rmu_z, rmu_s, amax_z, amax_s},
// has = bit0 (some dz > 0) | bit1 (some ds > 0).  Every lane returns the same values.
template <typename T>
__device__ __forceinline__ void warp_step_pieces(const T* z, const T* dz, const T* s, const T* ds, int m, int lane,
                                                 T (&out)[4], int& has, bool zero_is_one = false) {
  T rz = t_inf<T>(), rs = t_inf<T>(), az = -t_inf<T>(), as = -t_inf<T>();
  bool hz = false, hs = false;
  for (int i = lane; i < m; i += 32) {
    const T dzi = dz[i], dsi = ds[i];
    T a1 = -z[i] / dzi, a2 = -s[i] / dsi;
    if (zero_is_one) {  // batch_LU.py:208: a[dv == 0] = 1.0
      if (dzi == T(0)) a1 = T(1);
      if (dsi == T(0)) a2 = T(1);
    }
    if (dzi > T(0)) hz = true; else rz = nanmin(rz, a1);
    if (dsi > T(0)) hs = true; else rs = nanmin(rs, a2);
    az = nanmax(az, a1);
    as = nanmax(as, a2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    rz = nanmin(rz, shfl_x(rz, o));
    rs = nanmin(rs, shfl_x(rs, o));
    az = nanmax(az, shfl_x(az, o));
    as = nanmax(as, shfl_x(as, o));
  }
  has = (__any_sync(0xffffffffu, hz) ? 1 : 0) | (__any_sync(0xffffffffu, hs) ? 2 : 0);
  out[0] = rz; out[1] = rs; out[2] = az; out[3] = as;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += shfl_x(v, o);
  return v;
}

// ------------------------------------------------------------------------------------------
// INIT=true: initial point (batch.py:60-86).  INIT=false: PDIPM iteration a.iter (batch.py:91-204).
template <typename T, int MPAD, int NT, bool INIT, int FK = 0>
__global__ void __launch_bounds__(NT, (NT == 128 ? (FK == 1 ? FAST_DMMA_MIN_CTAS : 5) : (NT == 64 && FK == 1 ? 5 : 1))) k_fast_iter(const KArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int prob = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = a.n, m = a.m, p = a.p, it = a.iter;
  FS<T> S;
  fast_carve(S, smem_raw, a);
  int* ictl = reinterpret_cast<int*>(S.small);

  T fill_z = T(1), fill_s = T(1);
  int flags = 0;
  if (!INIT) {
    if (it > 0 && warp == 0) {
      const int term = eval_termination(a.slots, it, a.lim, a.eps, lane);
      if (lane == 0) ictl[0] = term;
    }
    flags = a.flags[prob];
  }
  if (!(flags & FLAG_POISON)) fast_stage(S, a, prob, tid, NT);
  if (!INIT && it > 0) {
    cta_sync<NT>();
    if (ictl[0] >= 0) { cp_async_wait_all(); return; }
    const Slot* sl = a.slots + (it - 1);
    const double gz = sl->az_nan ? 0.0 : ord_unkey(sl->amax_z);
    const double gs = sl->as_nan ? 0.0 : ord_unkey(sl->amax_s);
    fill_z = gz > 1.0 ? (T)gz : T(1);
    fill_s = gs > 1.0 ? (T)gs : T(1);
  }
  Slot* slot = a.slots + (INIT ? 0 : it);
  const int pp = p > 0 ? p : 1;
  T* gx = a.x + (size_t)prob * round4(n);
  T* gs_ = a.s + (size_t)prob * round4(m);
  T* gz_ = a.z + (size_t)prob * round4(m);
  T* gy = a.y + (size_t)prob * round4(pp);
  T* gdx = a.dx + (size_t)prob * round4(n);
  T* gds = a.ds + (size_t)prob * round4(m);
  T* gdz = a.dz + (size_t)prob * round4(m);
  T* gdy = a.dy + (size_t)prob * round4(pp);
  const T* Rt = a.R + (size_t)prob * a.sR;
  const T* pg = a.pv + (size_t)prob * a.sp;
  const T* hg = a.h + (size_t)prob * a.sh;
  const T* bg = a.b + (size_t)prob * a.sb;

  if (INIT) {
    const T d_init = T(1) + (T)a.reg;  // (z + reg) / s at z = s = 1
    for (int i = tid; i < m; i += NT) {
      S.d[i] = d_init; S.scr[i] = T(1) / d_init + (T)a.reg; S.rsc[i] = T(0); S.rz[i] = -hg[i];
      S.z[i] = T(1); S.s[i] = T(1);
    }
    for (int c = tid; c < n; c += NT) S.rx[c] = pg[c];
    for (int j = tid; j < p; j += NT) S.ry[j] = -bg[j];
    cta_sync<NT>();
    const bool ok = factor_any<T, MPAD, NT, FK>(Rt, S, (const T*)nullptr, m, tid);
    cp_async_wait_all();
    cta_sync<NT>();
    if (!ok) {
      if (tid == 0) a.flags[prob] = FLAG_POISON;
      return;
    }
    fast_kkt_solve<T, MPAD, NT, FK>(S, a, prob, true, S.rx, S.rsc, S.rz, S.ry, false, (T*)nullptr, (T*)nullptr,
                                (T*)nullptr, (T*)nullptr, tid);
    if (a.dense) dense_refine<T, MPAD, NT, FK>(S, a, prob, true, S.rx, S.rsc, S.rz, S.ry, S.z, S.s, tid);
    T mn[2] = {t_inf<T>(), t_inf<T>()};
    for (int i = tid; i < m; i += NT) { mn[0] = nanmin(mn[0], S.ds[i]); mn[1] = nanmin(mn[1], S.dz[i]); }
    block_reduce<2>(mn, OpNanMin(), S.red, tid, NT);
    for (int i = tid; i < m; i += NT) {
      T sv = S.ds[i], zv = S.dz[i];
      if (mn[0] < T(0)) sv -= mn[0] - T(1);
      if (mn[1] < T(0)) zv -= mn[1] - T(1);
      gs_[i] = sv; gz_[i] = zv;
    }
    if (tid < n) gx[tid] = S.dx[tid];
    for (int j = tid; j < p; j += NT) gy[j] = S.dy[j];
    if (tid == 0) a.flags[prob] = 0;
    return;
  } else {
    if (flags & FLAG_POISON) {
      if (it == 0) {
        T* bx = a.bx + (size_t)prob * n; T* bs = a.bs + (size_t)prob * m; T* bz = a.bz + (size_t)prob * m;
        for (int c = tid; c < n; c += NT) bx[c] = t_nan<T>();
        for (int i = tid; i < m; i += NT) { bs[i] = t_nan<T>(); bz[i] = t_nan<T>(); }
        if (p > 0) { T* by = a.by + (size_t)prob * p; for (int j = tid; j < p; j += NT) by[j] = t_nan<T>(); }
      }
      if (tid == 0) {
        double br = __longlong_as_double(0x7ff8000000000000LL);
        if (it == 0) a.best_resid[prob] = br; else br = a.best_resid[prob];
        if (br != br) slot->best_nan = 1; else atomic_max_key(&slot->best_max, (unsigned long long)__double_as_longlong(br));
        slot->mu_nan = 1; slot->az_nan = 1; slot->as_nan = 1;
      }
      return;
    }
    const T* Qg = a.Q + (size_t)prob * a.sQ;
    const T* Gg = a.G + (size_t)prob * a.sG;
    const T* Ag = a.A + (size_t)prob * a.sA;
    DmmaTiles<MPAD, NT / 32> tiles;
    if constexpr (FK == 1) dmma_prefetch<MPAD, NT / 32>(Rt, tid, tiles);
    // register-resident Q and G for the residual mat-vecs (loads in flight while the iterate arrives)
    constexpr bool REGMV = NT == 128 && MPAD <= 64;
    const bool regmv = REGMV && n <= 32;
    RegMat<16, T> Gm;
    RegMat<8, T> Qm;
    if (regmv) {
      regmat_load<16, T>(Gm, Gg, n, m, n, lane, warp);
      regmat_load<8, T>(Qm, Qg, n, n, n, lane, warp);
    }
    // ---- iterate + previous step
    {
      T alpha = T(0);
      if (it > 0) {
        const T rz_ = a.rmu[(size_t)prob * 2], rs_ = a.rmu[(size_t)prob * 2 + 1];
        const T stz = (flags & FLAG_FILL_Z) ? nanmin(rz_, fill_z) : rz_;
        const T sts = (flags & FLAG_FILL_S) ? nanmin(rs_, fill_s) : rs_;
        alpha = nanmin(T(0.999) * nanmin(stz, sts), T(1));
      }
      for (int c = tid; c < n; c += NT) { T v = gx[c]; if (it > 0) { v += alpha * gdx[c]; gx[c] = v; } S.x[c] = v; }
      for (int i = tid; i < m; i += NT) {
        T sv = gs_[i], zv = gz_[i];
        if (it > 0) { sv += alpha * gds[i]; zv += alpha * gdz[i]; gs_[i] = sv; gz_[i] = zv; }
        S.s[i] = sv; S.z[i] = zv;
      }
      for (int j = tid; j < p; j += NT) { T v = gy[j]; if (it > 0) { v += alpha * gdy[j]; gy[j] = v; } S.y[j] = v; }
    }
    cta_sync<NT>();
    // ---- residuals
    RegMat<8, T> Qim;
    if (regmv) {
      regmat_rows<8, T>(Qm, S.x, S.rx, n, n, lane, warp);
      regmat_rows<16, T>(Gm, S.x, S.rz, m, n, lane, warp);
      regmat_cols<16, T>(Gm, S.z, S.part, m, lane, warp);  // G^T z, four partials per column
      if constexpr (FK == 1) regmat_load<8, T>(Qim, a.Qi + (size_t)prob * a.sQi, a.ldn, n, n, lane, warp);
      if (p > 0) gemv_rows_warp(Ag, n, p, n, S.x, S.ry, tid, NT);
      cta_sync<NT>();
      if (tid < n) S.t[tid] = regmat_colsum(S.part, tid);
      cta_sync<NT>();
    } else {
      gemv_rows_warp(Qg, n, n, n, S.x, S.rx, tid, NT);
      gemv_rows_warp(Gg, n, m, n, S.x, S.rz, tid, NT);
      if (p > 0) gemv_rows_warp(Ag, n, p, n, S.x, S.ry, tid, NT);
      gemv_cols_nt<T, NT>(Gg, n, m, n, S.z, S.t, S.part, tid);  // G^T z
    }
    if (p > 0) gemv_cols_nt<T, NT>(Ag, n, p, n, S.y, S.scrn, S.part, tid);  // A^T y
    T acc[4] = {T(0), T(0), T(0), T(0)};
    for (int c = tid; c < n; c += NT) {
      T v = (a.cb_cg ? a.cb_cg[(size_t)prob * n + c] : S.rx[c] + pg[c]) + S.t[c];
      if (p > 0) v += S.scrn[c];
      S.rx[c] = v;
      acc[0] += v * v;
    }
    for (int i = tid; i < m; i += NT) {
      const T sv = S.s[i], zv = S.z[i];
      const T v = S.rz[i] + sv - hg[i];
      S.rz[i] = v;
      acc[1] += v * v;
      acc[3] += sv * zv;
      const T dv = (zv + (T)a.reg) / sv;  // reg = 0 unless DenseQPFunction
      S.d[i] = dv;
      S.scr[i] = T(1) / dv + (T)a.reg;
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

    bool ok;
    if constexpr (FK == 1) {
      // the predictor's right-hand side rides through the factorisation as a bordered row: its
      // forward substitution is free (qp_dmma.cuh)
      cp_async_wait_all();
      cta_sync<NT>();
      fast_kkt_pre<T, NT>(S, a, prob, true, S.rx, S.z, S.rz, S.ry, tid, regmv ? &Qim : nullptr);
      ok = factor_any<T, MPAD, NT, FK>(Rt, S, S.hv + p, m, tid, &tiles);
    } else {
      ok = factor_any<T, MPAD, NT, FK>(Rt, S, (const T*)nullptr, m, tid);
    }

    {
      const double rd = (double)resid;
      const double prev = a.best_resid[prob];
      const bool better = (it == 0) ? true : (rd < prev);
      cta_sync<NT>();
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
    cp_async_wait_all();
    cta_sync<NT>();
    if (!ok) {
      if (tid == 0) { a.flags[prob] = FLAG_POISON; slot->az_nan = 1; slot->as_nan = 1; }
      return;
    }
    // ---- predictor (pass 0: rs = z) and corrector (pass 1: rs = rsc, zero rx/rz/ry)
    for (int pass = 0; pass < 2; pass++) {
      const bool aff = pass == 0;
      const T* rs_ = aff ? S.z : S.rsc;
      if (!(FK == 1 && aff)) fast_kkt_pre<T, NT>(S, a, prob, aff, S.rx, rs_, S.rz, S.ry, tid);
      if (warp == 0) tri_solve_any<T, MPAD, FK>(S, m, S.hv + p, FK == 1 && aff, lane);
      cta_sync<NT>();
      if (!a.dense) {
        fast_kkt_post<T, NT>(S, a, prob, aff, rs_, !aff, aff ? (T*)nullptr : gdx, aff ? (T*)nullptr : gds,
                             aff ? (T*)nullptr : gdz, aff ? (T*)nullptr : gdy, tid);
      } else {
        // DenseQPFunction: regularised solve + one refinement step per KKT solve (batch_LU.py:212-244)
        fast_kkt_post<T, NT>(S, a, prob, aff, rs_, false, (T*)nullptr, (T*)nullptr, (T*)nullptr, (T*)nullptr, tid);
        dense_refine<T, MPAD, NT, FK>(S, a, prob, aff, S.rx, rs_, S.rz, S.ry, S.z, S.s, tid);
        if (aff) {
          for (int c = tid; c < n; c += NT) S.ax[c] = S.dx[c];
          for (int i = tid; i < m; i += NT) { S.as[i] = S.ds[i]; S.az[i] = S.dz[i]; }
          for (int j = tid; j < p; j += NT) S.ay[j] = S.dy[j];
        } else {
          for (int c = tid; c < n; c += NT) { const T v = S.dx[c] + S.ax[c]; S.dx[c] = v; gdx[c] = v; }
          for (int i = tid; i < m; i += NT) {
            const T vs = S.ds[i] + S.as[i], vz = S.dz[i] + S.az[i];
            S.ds[i] = vs; S.dz[i] = vz; gds[i] = vs; gdz[i] = vz;
          }
          for (int j = tid; j < p; j += NT) { const T v = S.dy[j] + S.ay[j]; S.dy[j] = v; gdy[j] = v; }
        }
        cta_sync<NT>();
      }
      if (warp == 0) {
        T pc[4]; int has;
        warp_step_pieces(S.z, S.dz, S.s, S.ds, m, lane, pc, has, a.dense != 0);
        if (aff) {
          // the clamp at 1 makes alpha_aff independent of the batch-global fill
          const T stz = (has & 1) ? nanmin(pc[0], T(1)) : pc[0];
          const T sts = (has & 2) ? nanmin(pc[1], T(1)) : pc[1];
          const T alpha_aff = nanmin(nanmin(stz, sts), T(1));
          T t3 = T(0);
          for (int i = lane; i < m; i += 32) t3 += (S.s[i] + alpha_aff * S.ds[i]) * (S.z[i] + alpha_aff * S.dz[i]);
          t3 = warp_sum(t3);
          const T ratio = t3 / t4;
          const T sig = ratio * ratio * ratio;
          for (int i = lane; i < m; i += 32) S.rsc[i] = (-mu * sig + S.ds[i] * S.dz[i]) / S.s[i];
        } else if (lane == 0) {
          a.rmu[(size_t)prob * 2] = pc[0];
          a.rmu[(size_t)prob * 2 + 1] = pc[1];
          a.flags[prob] = ((has & 1) ? FLAG_FILL_Z : 0) | ((has & 2) ? FLAG_FILL_S : 0);
          const double dz_ = (double)pc[2], ds_ = (double)pc[3];
          if (dz_ != dz_) slot->az_nan = 1; else atomic_max_key(&slot->amax_z, ord_key(dz_));
          if (ds_ != ds_) slot->as_nan = 1; else atomic_max_key(&slot->amax_s, ord_key(ds_));
        }
      }
      if (aff) cta_sync<NT>();
    }
  }
}

// ------------------------------------------------------------------------------------------
template <typename T, int MPAD, int NT, int FK = 0>
__global__ void __launch_bounds__(NT, (NT == 128 ? (FK == 1 ? FAST_DMMA_MIN_CTAS : 5) : (NT == 64 && FK == 1 ? 5 : 1))) k_fast_backward(const KArgs<T> a, const BArgs<T> g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int prob = blockIdx.x + a.prob0, tid = threadIdx.x;
  const int n = a.n, m = a.m, p = a.p;
  FS<T> S;
  fast_carve(S, smem_raw, a);
  fast_stage(S, a, prob, tid, NT);
  const T* zh = g.zhat + (size_t)prob * n;
  const T* lam = g.lams + (size_t)prob * m;
  const T* sl = g.slacks + (size_t)prob * m;
  const T* nu = g.nus + (size_t)prob * p;
  const T* gz = g.gz + (size_t)prob * n;
  for (int i = tid; i < m; i += NT) {
    const T lv = lam[i], sv = sl[i];
    S.z[i] = lv;
    // QPFunction clamps (qp.py:146-149); DenseQPFunction solves with the best iterate's K as is
    const T lc = (!a.dense && lv < T(1e-8)) ? T(1e-8) : lv, sc = (!a.dense && sv < T(1e-8)) ? T(1e-8) : sv;
    S.d[i] = lc / sc;
    S.scr[i] = T(1) / (lc / sc);
    S.rsc[i] = T(0);
    S.rz[i] = T(0);
  }
  for (int c = tid; c < n; c += NT) { S.rx[c] = gz[c]; S.x[c] = zh[c]; }
  for (int j = tid; j < p; j += NT) { S.y[j] = nu[j]; S.ry[j] = T(0); }
  cta_sync<NT>();
  factor_any<T, MPAD, NT, FK>(a.R + (size_t)prob * a.sR, S, (const T*)nullptr, m, tid);  // NaN factor on failure
  cp_async_wait_all();
  cta_sync<NT>();
  fast_kkt_solve<T, MPAD, NT, FK>(S, a, prob, true, S.rx, S.rsc, S.rz, S.ry, false, (T*)nullptr, (T*)nullptr, (T*)nullptr,
                              (T*)nullptr, tid);
  cta_sync<NT>();
  T* dp = g.dp + (size_t)prob * n; T* dh = g.dh + (size_t)prob * m;
  for (int c = tid; c < n; c += NT) dp[c] = S.dx[c];
  for (int i = tid; i < m; i += NT) dh[i] = -S.dz[i];
  if (p > 0) { T* db = g.db + (size_t)prob * p; for (int j = tid; j < p; j += NT) db[j] = -S.dy[j]; }
  T* dQ = g.dQ + (size_t)prob * n * n;
  for (int e = tid; e < n * n; e += NT) {
    const int r = e / n, c = e - r * n;
    dQ[e] = T(0.5) * (S.dx[r] * S.x[c] + S.x[r] * S.dx[c]);
  }
  T* dG = g.dG + (size_t)prob * m * n;
  for (int e = tid; e < m * n; e += NT) {
    const int r = e / n, c = e - r * n;
    dG[e] = S.dz[r] * S.x[c] + S.z[r] * S.dx[c];
  }
  if (p > 0) {
    T* dA = g.dA + (size_t)prob * p * n;
    for (int e = tid; e < p * n; e += NT) {
      const int r = e / n, c = e - r * n;
      dA[e] = S.dy[r] * S.x[c] + S.y[r] * S.dx[c];
    }
  }
}

template <typename T, int MPAD, int NT, int FK = 0>
__global__ void __launch_bounds__(NT, (NT == 128 ? (FK == 1 ? FAST_DMMA_MIN_CTAS : 5) : (NT == 64 && FK == 1 ? 5 : 1))) k_fast_kkt(const KArgs<T> a, const SArgs<T> g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int prob = blockIdx.x, tid = threadIdx.x;
  const int n = a.n, m = a.m, p = a.p;
  FS<T> S;
  fast_carve(S, smem_raw, a);
  fast_stage(S, a, prob, tid, NT);
  for (int i = tid; i < m; i += NT) {
    S.d[i] = g.d[(size_t)prob * m + i];
    S.scr[i] = T(1) / S.d[i];
    S.rsc[i] = g.rs[(size_t)prob * m + i];
    S.rz[i] = g.rz[(size_t)prob * m + i];
  }
  for (int c = tid; c < n; c += NT) S.rx[c] = g.rx[(size_t)prob * n + c];
  for (int j = tid; j < p; j += NT) S.ry[j] = g.ry[(size_t)prob * p + j];
  cta_sync<NT>();
  factor_any<T, MPAD, NT, FK>(a.R + (size_t)prob * a.sR, S, (const T*)nullptr, m, tid);
  cp_async_wait_all();
  cta_sync<NT>();
  fast_kkt_solve<T, MPAD, NT, FK>(S, a, prob, true, S.rx, S.rsc, S.rz, S.ry, false, g.dx + (size_t)prob * n,
                              g.ds + (size_t)prob * m, g.dz + (size_t)prob * m, p > 0 ? g.dy + (size_t)prob * p : (T*)nullptr, tid);
}

}  // namespace b200qp
