// qp_blocked.cuh -- blocked LDL^T and blocked triangular sweeps for QPs whose nineq x nineq system does
// not fit in shared memory (BASELINE configs[4]: nz = 100 / 200, nineq = 200 / 400), fp64, sm_100a.
//
// The matrix T = R + diag(s/z) of such a problem lives in a global-memory slab (L2 resident while its
// CTA works on it).  The column-by-column factorisation of qp_common.cuh:ldlt_factor does a barrier
// and a global read-modify-write of the whole trailing matrix PER COLUMN (m barriers, m^3/3 global
// FMAs); measured 18 ms per launch for 1000 QPs at nineq = 200 (0.7 % of the FP64 peak).  Here:
//   * right-looking, 16-column panels: the 16 x 16 diagonal block is factored by one warp in registers
//     (shuffles broadcast the pivot row), the rows below it by ONE THREAD PER ROW (a 16-entry register
//     row, the factored diagonal block broadcast from shared memory), which also writes the unit-upper
//     rows U = L^T coalesced and leaves the unscaled columns W = L D in a shared panel buffer;
//   * the trailing matrix is updated tile by tile on the FP64 tensor cores: C(8x8, global) -=
//     (W D^-1)(W)^T, four DMMA m8n8k4 per tile and panel, operands from the shared panel buffer
//     (row stride 20 doubles: conflict-free fragment loads); 3 barriers per PANEL;
//   * the triangular sweeps run 32 columns at a time: one warp solves the 32 x 32 diagonal block in
//     registers, then every thread updates its remaining entries with 32 independent FMAs.
// Storage convention identical to ldlt_factor: on exit the strict upper triangle holds U (unit
// diagonal implied), pinv[j] = 1/D_j, the lower triangle is destroyed.
#pragma once
#include "qp_common.cuh"

namespace b200qp {

constexpr int kBlkPW = 16;  // panel width
constexpr int kBlkPS = 20;  // doubles per row of the panel buffer
__host__ __device__ inline int blk_panel_elems(int m) { return ((m + 7) & ~7) * kBlkPS + kBlkPW + 8; }

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// 1/x for a positive pivot: MUFU seed + two Newton steps (full double accuracy away from the range ends)
__device__ __forceinline__ double blk_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  if (!(x > 1e-290 && x < 1e290)) r = 1.0 / x;  // rare
  return r;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// All threads of the CTA; begins and ends with a barrier.  Wp: blk_panel_elems(m) doubles of shared memory.
// R0 != nullptr: the matrix to factor is R0 + diag(1 / dvec) (R0 read-only, same layout); it is read in place
// by the first panel step, whose trailing update writes the first version of S -- no separate copy pass.
__device__ __forceinline__ bool ldlt_factor_blocked(double* __restrict__ S, int ld, int m, double* pinv, double* Wp,
                                                    int tid, int nt, const double* __restrict__ R0 = nullptr,
                                                    const double* dvec = nullptr) {
  constexpr int PW = kBlkPW, PS = kBlkPS;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  const int fr = lane >> 2, fc = (lane & 3) * 2, kc = lane & 3;
  const int m8 = (m + 7) & ~7;
  double* ppan = Wp + m8 * PS;   // the panel's reciprocal pivots (0 for columns past m)
  double* flag = ppan + PW;      // NaN once a pivot was not positive
  if (tid == 0) flag[0] = 0.0;
  __syncthreads();
  bool ok = true;
#pragma unroll 1
  for (int c0 = 0; c0 < m; c0 += PW) {
    const int pw = min(PW, m - c0);
    // (1) diagonal block, one warp: lane owns row c0 + lane
    if (warp == 0) {
      double a[PW];
      {
        const double* src = (c0 == 0 && R0) ? R0 : S;
#pragma unroll
        for (int c = 0; c < PW; c++) a[c] = (lane < pw && c <= lane) ? src[(size_t)(c0 + lane) * ld + c0 + c] : 0.0;
        if (c0 == 0 && R0 && lane < pw) {
          const double dl = 1.0 / dvec[lane];
#pragma unroll
          for (int c = 0; c < PW; c++) if (c == lane) a[c] += dl;
        }
      }
      bool bad = false;
#pragma unroll
      for (int k = 0; k < PW; k++) {
        const double dk = shfl_d(a[k], k);
        double pkk = 0.0;
        if (k < pw) {
          pkk = (dk > 0.0) ? blk_rcp(dk) : t_nan<double>();
          if (!(dk > 0.0)) bad = true;
        }
        const double w = a[k];
        const double l = w * pkk;
        if (lane == k) { ppan[k] = pkk; if (k < pw) pinv[c0 + k] = pkk; }
        if (lane > k && lane < pw) S[(size_t)(c0 + k) * ld + c0 + lane] = l;  // U[c0+k][c0+lane]
#pragma unroll
        for (int c = k + 1; c < PW; c++) {
          const double wck = shfl_d(w, c);
          a[c] -= l * wck;
        }
      }
#pragma unroll
      for (int c = 0; c < PW; c++) if (lane < PW) Wp[lane * PS + c] = (c <= lane) ? a[c] : 0.0;
      if (bad && lane == 0) flag[0] = t_nan<double>();
    }
    __syncthreads();
    if (is_nan(flag[0])) { ok = false; break; }  // uniform
    // (2) rows below the diagonal block: W21 = A21 L11^-T (unscaled), U rows to global, W to the panel buffer
    for (int i = c0 + pw + tid; i < m8; i += nt) {
      double a[PW];
      if (i < m) {
        const double* row = ((c0 == 0 && R0) ? R0 : S) + (size_t)i * ld + c0;
#pragma unroll
        for (int c = 0; c < PW; c++) a[c] = c < pw ? row[c] : 0.0;
#pragma unroll
        for (int k = 0; k < PW; k++) {
          if (k < pw) {
            const double l = a[k] * ppan[k];
            S[(size_t)(c0 + k) * ld + i] = l;
#pragma unroll
            for (int c = k + 1; c < PW; c++) a[c] -= l * Wp[c * PS + k];
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < PW; c++) a[c] = 0.0;
      }
      double* wr = Wp + (size_t)(i - c0) * PS;
#pragma unroll
      for (int c = 0; c < PW; c += 2) *reinterpret_cast<double2*>(wr + c) = make_double2(a[c], a[c + 1]);
    }
    __syncthreads();
    // (3) trailing update, 8x8 tiles of the lower triangle (diagonal tiles included)
    const int r0 = c0 + pw;
    if (r0 < m) {
      const int nt8 = (m - r0 + 7) >> 3, ntiles = nt8 * (nt8 + 1) / 2;
      const double s0 = -ppan[kc], s1 = -ppan[4 + kc], s2 = -ppan[8 + kc], s3 = -ppan[12 + kc];
      // (d) the next panel's diagonal block and first rows are touched next by ONE warp: pull them into L1 now
      if (warp == nwarps - 1 && r0 + lane < m) {
        const double* nx = S + (size_t)(r0 + lane) * ld + r0;
        prefetch_l1(nx);
        prefetch_l1(nx + 15);
      }
      // four tiles in flight per warp: their loads are issued together, then the DMMAs, then the stores
      constexpr int TF = 4;
      int I = 0, K = warp;  // tile t = warp, warp + nwarps, ... of the row-major lower-triangle enumeration
      while (K > I) { K -= I + 1; I++; }
      const bool first = (c0 == 0 && R0 != nullptr);
      const double* src = first ? R0 : S;
      for (int t = warp; t < ntiles; t += TF * nwarps) {
        int tI[TF], tK[TF];
        double x0[TF], x1[TF];
        bool v0[TF], v1[TF];
#pragma unroll
        for (int f = 0; f < TF; f++) {
          tI[f] = I; tK[f] = K;
          const bool live = t + f * nwarps < ntiles;
          const int gi = r0 + 8 * I + fr, gk = r0 + 8 * K + fc;
          v0[f] = live && gi < m && gk < m;
          v1[f] = live && gi < m && gk + 1 < m;
          const double* rp = src + (size_t)gi * ld + gk;
          x0[f] = v0[f] ? rp[0] : 0.0;
          x1[f] = v1[f] ? rp[1] : 0.0;
          K += nwarps;
          while (K > I) { K -= I + 1; I++; }
        }
#pragma unroll
        for (int f = 0; f < TF; f++) {
          if (t + f * nwarps >= ntiles) break;  // warp-uniform
          const int gi = r0 + 8 * tI[f] + fr, gk = r0 + 8 * tK[f] + fc;
          if (first && tI[f] == tK[f] && gi < m) {
            if (gi == gk) x0[f] += 1.0 / dvec[gi];
            if (gi == gk + 1) x1[f] += 1.0 / dvec[gi];
          }
          const double* ra = Wp + (size_t)(pw + 8 * tI[f] + fr) * PS + kc;
          const double* rb = Wp + (size_t)(pw + 8 * tK[f] + fr) * PS + kc;
          dmma884(x0[f], x1[f], ra[0] * s0, rb[0]);
          dmma884(x0[f], x1[f], ra[4] * s1, rb[4]);
          dmma884(x0[f], x1[f], ra[8] * s2, rb[8]);
          dmma884(x0[f], x1[f], ra[12] * s3, rb[12]);
          double* cp = S + (size_t)gi * ld + gk;
          if (v0[f]) cp[0] = x0[f];
          if (v1[f]) cp[1] = x1[f];
        }
      }
    }
    __syncthreads();
  }
  if (!ok) {
    for (int i = tid; i < m; i += nt) pinv[i] = t_nan<double>();
    __syncthreads();
  }
  return ok;
}

// v <- (L D L^T)^-1 v with the factor above.  All threads of the CTA; v in shared memory and visible
// on entry (caller barrier); ends with a barrier.
__device__ __forceinline__ void ldlt_solve_blocked(const double* __restrict__ U, int ld, int m, const double* pinv,
                                                   double* v, int tid, int nt) {
  const int lane = tid & 31, warp = tid >> 5;
  // forward: L y = v, L[i][j] = U[j][i]
#pragma unroll 1
  for (int j0 = 0; j0 < m; j0 += 32) {
    const int bw = min(32, m - j0);
    if (warp == 0) {
      double u[32];
#pragma unroll
      for (int jj = 0; jj < 32; jj++) u[jj] = (jj < bw && lane > jj && lane < bw) ? U[(size_t)(j0 + jj) * ld + j0 + lane] : 0.0;
      double r = lane < bw ? v[j0 + lane] : 0.0;
#pragma unroll
      for (int jj = 0; jj < 32; jj++) {
        const double yj = shfl_d(r, jj);
        r -= u[jj] * yj;
      }
      if (lane < bw) v[j0 + lane] = r;
    }
    __syncthreads();
    if (warp == 1 && j0 + 32 + lane < m) {  // next diagonal block -> L1 while the update runs
      const double* nx = U + (size_t)(j0 + 32 + lane) * ld + j0 + 32;
      prefetch_l1(nx); prefetch_l1(nx + 16); prefetch_l1(nx + 31);
    }
    for (int i = j0 + bw + tid; i < m; i += nt) {
      const double* col = U + (size_t)j0 * ld + i;
      double acc0 = v[i], acc1 = 0.0;
#pragma unroll
      for (int j8 = 0; j8 < 32; j8 += 8) {
        if (j8 < bw) {  // uniform
          double u8[8];
#pragma unroll
          for (int q = 0; q < 8; q++) u8[q] = (j8 + q < bw) ? col[(size_t)(j8 + q) * ld] : 0.0;
#pragma unroll
          for (int q = 0; q < 8; q += 2) {
            if (j8 + q < bw) acc0 -= u8[q] * v[j0 + j8 + q];
            if (j8 + q + 1 < bw) acc1 -= u8[q + 1] * v[j0 + j8 + q + 1];
          }
        }
      }
      v[i] = acc0 + acc1;
    }
    __syncthreads();
  }
  for (int i = tid; i < m; i += nt) v[i] *= pinv[i];
  __syncthreads();
  // backward: L^T x = y, x_i = y_i - sum_{j>i} U[i][j] x_j
#pragma unroll 1
  for (int j0 = ((m - 1) >> 5) << 5; j0 >= 0; j0 -= 32) {
    const int bw = min(32, m - j0);
    if (warp == 0) {
      double u[32];
      const double* row = U + (size_t)(j0 + lane) * ld + j0;
#pragma unroll
      for (int jj = 0; jj < 32; jj++) u[jj] = (jj < bw && jj > lane && lane < bw) ? row[jj] : 0.0;
      double r = lane < bw ? v[j0 + lane] : 0.0;
#pragma unroll
      for (int jj = 31; jj >= 0; jj--) {
        const double xj = shfl_d(r, jj);
        r -= u[jj] * xj;
      }
      if (lane < bw) v[j0 + lane] = r;
    }
    __syncthreads();
    if (warp == 1 && j0 >= 32) {  // previous diagonal block -> L1 while the update runs
      const double* nx = U + (size_t)(j0 - 32 + lane) * ld + j0 - 32;
      prefetch_l1(nx); prefetch_l1(nx + 16); prefetch_l1(nx + 31);
    }
    for (int i = tid; i < j0; i += nt) {
      const double* row = U + (size_t)i * ld + j0;
      double acc0 = v[i], acc1 = 0.0;
#pragma unroll
      for (int j8 = 0; j8 < 32; j8 += 8) {
        if (j8 < bw) {  // uniform
          double u8[8];
#pragma unroll
          for (int q = 0; q < 8; q++) u8[q] = (j8 + q < bw) ? row[j8 + q] : 0.0;
#pragma unroll
          for (int q = 0; q < 8; q += 2) {
            if (j8 + q < bw) acc0 -= u8[q] * v[j0 + j8 + q];
            if (j8 + q + 1 < bw) acc1 -= u8[q + 1] * v[j0 + j8 + q + 1];
          }
        }
      }
      v[i] = acc0 + acc1;
    }
    __syncthreads();
  }
}

// X = (L D L^T)^-1 from the factor above, ONE THREAD PER COLUMN, no barriers inside: thread c solves
// L y = e_c, scales by D^-1 and solves L^T x = y, eight rows at a time in registers; its column of X is
// its only working storage (coalesced across threads), the factor is read through warp-uniform
// (broadcast) loads.  Loop bounds are warp-uniform (rows start at the warp's first column) so that the
// broadcasts stay broadcasts.  X may be global or shared; call with all threads, barrier before/after.
__device__ __forceinline__ void ldlt_inverse_cols(const double* __restrict__ U, int ld, int n, const double* pinv,
                                                  double* X, int tid, int nt) {
  constexpr int RB = 8;
  const int lane = tid & 31;
  for (int c = tid; c < n; c += nt) {
    const int cw = c - lane;               // first column of this warp (warp-uniform)
    const int ib = (cw / RB) * RB;
    for (int i = 0; i < ib; i++) X[(size_t)i * ld + c] = 0.0;
#pragma unroll 1
    for (int i0 = ib; i0 < n; i0 += RB) {
      double y[RB];
#pragma unroll
      for (int r = 0; r < RB; r++) y[r] = (i0 + r == c) ? 1.0 : 0.0;
      for (int j = cw; j < i0; j++) {
        const double yj = X[(size_t)j * ld + c];
        const double* urow = U + (size_t)j * ld + i0;
#pragma unroll
        for (int r = 0; r < RB; r++) if (i0 + r < n) y[r] -= urow[r] * yj;
      }
#pragma unroll
      for (int r = 1; r < RB; r++) {
#pragma unroll
        for (int r2 = 0; r2 < r; r2++) if (i0 + r < n) y[r] -= U[(size_t)(i0 + r2) * ld + i0 + r] * y[r2];
      }
#pragma unroll
      for (int r = 0; r < RB; r++) if (i0 + r < n) X[(size_t)(i0 + r) * ld + c] = y[r];
    }
#pragma unroll 1
    for (int i0 = ((n - 1) / RB) * RB; i0 >= 0; i0 -= RB) {
      double x[RB];
#pragma unroll
      for (int r = 0; r < RB; r++) x[r] = (i0 + r < n) ? X[(size_t)(i0 + r) * ld + c] * pinv[i0 + r] : 0.0;
      for (int j = i0 + RB; j < n; j++) {
        const double xj = X[(size_t)j * ld + c];
#pragma unroll
        for (int r = 0; r < RB; r++) if (i0 + r < n) x[r] -= U[(size_t)(i0 + r) * ld + j] * xj;
      }
#pragma unroll
      for (int r = RB - 2; r >= 0; r--) {
#pragma unroll
        for (int r2 = r + 1; r2 < RB; r2++) if (i0 + r2 < n) x[r] -= U[(size_t)(i0 + r) * ld + i0 + r2] * x[r2];
      }
#pragma unroll
      for (int r = 0; r < RB; r++) if (i0 + r < n) X[(size_t)(i0 + r) * ld + c] = x[r];
    }
  }
}

// C = A B on the FP64 tensor cores, operands straight from global / shared memory, one 16 x 16 output
// block (2 x 2 DMMA tiles) per warp and step.  arow(r) -> pointer to row r of A (K contiguous entries);
// bval(k, c) -> B[k][c]; store(r, c, v) consumes the result; `lower` restricts to blocks J <= I.
template <class ARow, class BVal, class Store>
__device__ __forceinline__ void dmma_gemm(int rows, int cols, int K, bool lower, ARow arow, BVal bval, Store store,
                                          int tid, int nt) {
  const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  const int fr = lane >> 2, kc = lane & 3, fc = kc * 2;
  const int TR = (rows + 15) >> 4, TC = (cols + 15) >> 4;
  for (int t = warp; t < TR * TC; t += nw) {
    const int I = t / TC, J = t - I * TC;
    if (lower && J > I) continue;
    const int r0 = 16 * I + fr, r1 = r0 + 8, c0 = 16 * J + fr, c1 = c0 + 8;
    const double* a0p = r0 < rows ? arow(r0) : nullptr;
    const double* a1p = r1 < rows ? arow(r1) : nullptr;
    double acc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
#pragma unroll 2
    for (int k0 = 0; k0 < K; k0 += 4) {
      const int k = k0 + kc;
      const bool kin = k < K;
      const double a0 = (a0p && kin) ? a0p[k] : 0.0;
      const double a1 = (a1p && kin) ? a1p[k] : 0.0;
      const double b0 = (kin && c0 < cols) ? bval(k, c0) : 0.0;
      const double b1 = (kin && c1 < cols) ? bval(k, c1) : 0.0;
      dmma884(acc[0][0][0], acc[0][0][1], a0, b0);
      dmma884(acc[0][1][0], acc[0][1][1], a0, b1);
      dmma884(acc[1][0][0], acc[1][0][1], a1, b0);
      dmma884(acc[1][1][0], acc[1][1][1], a1, b1);
    }
#pragma unroll
    for (int ii = 0; ii < 2; ii++) {
#pragma unroll
      for (int jj = 0; jj < 2; jj++) {
        const int r = 16 * I + 8 * ii + fr, c = 16 * J + 8 * jj + fc;
        if (r < rows && c < cols) store(r, c, acc[ii][jj][0]);
        if (r < rows && c + 1 < cols) store(r, c + 1, acc[ii][jj][1]);
      }
    }
  }
}

}  // namespace b200qp
