// qp_dmma.cuh -- blocked right-looking LDL^T of T = R + diag(1/d) on the FP64 tensor cores
// (DMMA m8n8k4, sm_100a), 128-thread CTA per QP, fp64 only.
//
// Why: ncu on the register-tile factorisation of qp_fast.cuh (profiles/r01/) shows 45 % of the
// iteration kernel's instructions and 33 % of its stall samples in the factor, 60 CTA barriers
// per factorisation and ~830 cycles per column step.  This version needs 16 barriers and ~5x
// fewer instructions (3 barriers per 8-column panel):
//   * the trailing matrix lives in registers as 8x8 DMMA accumulator tiles (36 lower tiles at
//     MPAD=64, 9 per warp, 2 doubles per lane per tile);
//   * per 8-column panel: the owners of the panel's tiles drop them into a [MPAD][10] shared
//     buffer; ONE warp factors the panel with each lane owning two rows (shuffles broadcast the
//     pivot row, no barrier inside the panel); it leaves the unscaled columns W = L D in the
//     buffer, the unit-lower columns packed in `Up` for the triangular sweeps, 1/D in `pinv`;
//   * every warp then updates its tiles with two DMMAs per tile: C -= (W D^-1)(W)^T.
//   * an optional extra ROW (index m) carries a right-hand side: after the factorisation
//     U[j][m] = (D^-1 L^-1 rhs)_j, i.e. the forward substitution of that solve is free.
// R is stored by the pre-factorisation in FRAGMENT ORDER (frag_index below) so that loading a
// tile is one coalesced 16-byte load per lane.
#pragma once
#include "qp_common.cuh"
// included by qp_fast.cuh after pivot_rcp() is defined

namespace b200qp {

constexpr int kPanelStride = 10;  // doubles per row of the panel buffer: conflict-free fragment loads

__host__ __device__ inline int frag_tiles(int mpad) { const int n = mpad / 8; return n * (n + 1) / 2; }
__host__ __device__ inline int frag_elems(int mpad) { return frag_tiles(mpad) * 64; }
// position of T[i][k] (k's tile column <= i's tile row) in fragment order
__host__ __device__ inline int frag_index(int i, int k) {
  const int I = i >> 3, K = k >> 3, r = i & 7, c = k & 7;
  return (I * (I + 1) / 2 + K) * 64 + (r * 4 + (c >> 1)) * 2 + (c & 1);
}
// packed offset of row j of the unit upper factor with logical size mm: U[j][i] = Up[urow(j,mm) + i]
__host__ __device__ inline int urow(int j, int mm) { return j * mm - (j * (j + 1)) / 2 - j - 1; }

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Issue the global loads of this thread's accumulator tiles early (whole 16-byte fragments, the
// buffer is fully allocated) so that their latency hides behind the residual computation.
template <int MPAD, int NW = 4>
struct DmmaTiles {
  static constexpr int SLOTS = (MPAD / 8 * (MPAD / 8 + 1) / 2 + NW - 1) / NW;
  double2 raw[SLOTS];
};
template <int MPAD, int NW = 4>
__device__ __forceinline__ void dmma_prefetch(const double* __restrict__ Rf, int tid, DmmaTiles<MPAD, NW>& tl) {
  constexpr int NTILES = MPAD / 8 * (MPAD / 8 + 1) / 2;
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int s = 0; s < DmmaTiles<MPAD, NW>::SLOTS; s++) {
    const int t = NW * s + warp;
    tl.raw[s] = make_double2(0.0, 0.0);
    if (t < NTILES) tl.raw[s] = *reinterpret_cast<const double2*>(Rf + (size_t)t * 64 + lane * 2);
  }
}

// T = R + diag(dinv) (m x m, dinv = 1/d = s/z) bordered by the optional row `hz` at index m; factor in place.
//   Rf    global, fragment order (frag_index), only entries i,k < m are read
//   Up    shared, packed unit upper factor of logical size mm = m + 1  ((m+1)m/2 doubles)
//   pinv  shared, m reciprocal pivots
//   Pb    shared panel buffer, MPAD * kPanelStride doubles (+8 for the panel's pivots)
// Returns false (uniformly) when a pivot of a real column is <= 0 or NaN; Up/pinv are then NaN.
// Must be called by all 128 threads; begins and ends with a CTA barrier.
template <int MPAD, int NW = 4>
__device__ __forceinline__ bool dmma_factor(const double* __restrict__ Rf, const double* dinv, const double* hz, double* Up,
                                            double* pinv, double* Pb, int m, int tid,
                                            const DmmaTiles<MPAD, NW>* pre = nullptr) {
  constexpr int NTI = MPAD / 8, NTILES = NTI * (NTI + 1) / 2, SLOTS = (NTILES + NW - 1) / NW, PS = kPanelStride, NTH = 32 * NW;
  const int lane = tid & 31, warp = tid >> 5;
  const int fr = lane >> 2, fc = (lane & 3) * 2;  // accumulator fragment: row fr, columns fc, fc+1
  const int mm = m + 1;
  double C[SLOTS][2];
  int tI[SLOTS], tK[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS; s++) {
    const int t = NW * s + warp;
    int I = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);  // exact for the <= 136 tiles used
    if ((I + 1) * (I + 2) / 2 <= t) I++;
    const int K = t - I * (I + 1) / 2;
    tI[s] = I; tK[s] = K;
    C[s][0] = 0.0; C[s][1] = 0.0;
    if (t < NTILES) {
      const double2 rr = pre ? pre->raw[s] : *reinterpret_cast<const double2*>(Rf + (size_t)t * 64 + lane * 2);
      if (I != K && 8 * I + 8 <= m) {  // uniform: interior tile, every entry is a real strictly-lower one
        C[s][0] = rr.x; C[s][1] = rr.y;
      } else {
        const int i = 8 * I + fr, k = 8 * K + fc;
        double v0 = 0.0, v1 = 0.0;
        if (i < m && k < m) {  // k even, m may be odd: guard the second element separately
          v0 = (k <= i) ? rr.x : 0.0;  // the strict upper part of a diagonal tile is never used
          v1 = (k + 1 <= i) ? rr.y : 0.0;
          if (I == K) {
            if (i == k) v0 += dinv[i];
            if (i == k + 1) v1 += dinv[i];
          }
        } else if (i == m) {
          if (hz != nullptr) {
            if (k < m) v0 = hz[k];
            if (k + 1 < m) v1 = hz[k + 1];
          }
        }
        if (i >= m) {  // bordered row / padding: unit diagonal
          if (i == k) v0 = 1.0;
          if (i == k + 1) v1 = 1.0;
        }
        C[s][0] = v0; C[s][1] = v1;
      }
    }
  }
  double* ppan = Pb + MPAD * PS;  // the current panel's 8 reciprocal pivots (1 for bordered/padding columns)
  bool ok = true;
  const int npan = (mm + 7) >> 3;
#pragma unroll 1
  for (int J = 0; J < npan; J++) {
    // (a) owners publish the panel's tiles
#pragma unroll
    for (int s = 0; s < SLOTS; s++) {
      if (NW * s + warp < NTILES && tK[s] == J) {
        *reinterpret_cast<double2*>(Pb + (8 * tI[s] + fr) * PS + fc) = make_double2(C[s][0], C[s][1]);
      }
    }
    __syncthreads();
    // (b) one warp factors the panel: lane owns rows i0, i1
    if (warp == 0) {
      const int rbase = 8 * J;
      const int i0 = rbase + lane, i1 = rbase + 32 + lane;
      double a0[8], a1[8];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        double2 v = make_double2(0.0, 0.0), w = make_double2(0.0, 0.0);
        if (i0 < MPAD) v = *reinterpret_cast<const double2*>(Pb + i0 * PS + 2 * q);
        if (i1 < MPAD) w = *reinterpret_cast<const double2*>(Pb + i1 * PS + 2 * q);
        a0[2 * q] = v.x; a0[2 * q + 1] = v.y;
        a1[2 * q] = w.x; a1[2 * q + 1] = w.y;
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const double dk = shfl_d(a0[k], k);
        // 1/dk: MUFU seed + two Newton steps, no branches on the critical path; a non-positive or
        // NaN pivot of a real column poisons the factor, bordered/padding columns use 1
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(dk));
        double e = fma(-dk, r, 1.0);
        r = fma(r, e, r);
        e = fma(-dk, r, 1.0);
        r = fma(r, e, r);
        if (!(dk > 1e-290 && dk < 1e290)) r = (dk > 0.0) ? 1.0 / dk : t_nan<double>();  // rare
        const double pkk = (rbase + k < m) ? r : 1.0;
        if (lane == 0) ppan[k] = pkk;  // re-read below: keeping eight pivots live costs 16 registers
        const double w0 = a0[k], w1 = a1[k];
        const double l0 = w0 * pkk, l1 = w1 * pkk;
#pragma unroll
        for (int c = k + 1; c < 8; c++) {
          const double wck = shfl_d(w0, c);
          a0[c] -= l0 * wck;
          a1[c] -= l1 * wck;
        }
      }
      // off the critical path: reciprocal pivots, packed unit-lower columns for the sweeps
      __syncwarp();
      if (lane < 8 && rbase + lane < m) pinv[rbase + lane] = ppan[lane];
      {
        int ub = urow(rbase, mm);
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const int j = rbase + k;
          if (j < m) {
            const double pkk = ppan[k];
            if (i0 > j && i0 < mm) Up[ub + i0] = a0[k] * pkk;
            if (i1 < mm) Up[ub + i1] = a1[k] * pkk;
          }
          ub += mm - j - 2;  // urow(j + 1) - urow(j)
        }
      }
      // unscaled columns W = L D of the rows below the diagonal block feed the trailing update
#pragma unroll
      for (int q = 0; q < 4; q++) {
        if (i0 >= rbase + 8 && i0 < MPAD) *reinterpret_cast<double2*>(Pb + i0 * PS + 2 * q) = make_double2(a0[2 * q], a0[2 * q + 1]);
        if (i1 < MPAD) *reinterpret_cast<double2*>(Pb + i1 * PS + 2 * q) = make_double2(a1[2 * q], a1[2 * q + 1]);
      }
    }
    __syncthreads();
    // (c) trailing update of every tile right of the panel:  C -= (W D^-1) W^T
    {
      const int kc = lane & 3;
      const double s0 = -ppan[kc], s1 = -ppan[4 + kc];
      if (is_nan(ppan[0] + ppan[1] + ppan[2] + ppan[3] + ppan[4] + ppan[5] + ppan[6] + ppan[7])) ok = false;  // uniform
#pragma unroll
      for (int s = 0; s < SLOTS; s++) {
        if (NW * s + warp < NTILES && tK[s] > J) {
          const double* ra = Pb + (8 * tI[s] + fr) * PS + kc;
          const double* rb = Pb + (8 * tK[s] + fr) * PS + kc;
          dmma_m8n8k4(C[s][0], C[s][1], ra[0] * s0, rb[0]);
          dmma_m8n8k4(C[s][0], C[s][1], ra[4] * s1, rb[4]);
        }
      }
    }
    __syncthreads();  // the next panel's tiles overwrite the buffer these fragments were read from
  }
  if (!ok) {
    for (int i = tid; i < m; i += NTH) pinv[i] = t_nan<double>();
    for (int i = tid; i < mm * (mm - 1) / 2; i += NTH) Up[i] = t_nan<double>();
    __syncthreads();
  }
  return ok;
}

// Triangular sweeps with the packed factor of logical size mm (m real rows), ONE warp, working
// vector in registers.  from_border: the right-hand side is the bordered row of the factor,
// r_j = U[j][m] = (D^-1 L^-1 rhs)_j, so only the backward sweep runs.
template <int RPL>
__device__ __forceinline__ void dmma_ldlt_solve(const double* Up, int m, int mm, const double* pinv, double* v,
                                                bool from_border, int lane) {
  double r[RPL];
  int rb[RPL];
#pragma unroll
  for (int s = 0; s < RPL; s++) {
    const int i = s * 32 + lane;
    rb[s] = urow(i, mm);
    if (from_border) r[s] = i < m ? Up[rb[s] + m] : 0.0;
    else r[s] = i < m ? v[i] : 0.0;
  }
  if (!from_border) {
#pragma unroll
    for (int s = 0; s < RPL; s++) {
      const int jend = min(32, m - s * 32);
#pragma unroll 4
      for (int jj = 0; jj < jend; jj++) {
        const int j = s * 32 + jj;
        const double yj = shfl_d(r[s], jj);
        const double* row = Up + urow(j, mm);
#pragma unroll
        for (int s2 = s; s2 < RPL; s2++) {
          const int i = s2 * 32 + lane;
          if (i > j && i < m) r[s2] -= row[i] * yj;
        }
      }
    }
#pragma unroll
    for (int s = 0; s < RPL; s++) {
      const int i = s * 32 + lane;
      if (i < m) r[s] *= pinv[i];
    }
  }
#pragma unroll
  for (int s = RPL - 1; s >= 0; s--) {
    const int jend = min(32, m - s * 32);
#pragma unroll 4
    for (int jj = jend - 1; jj >= 0; jj--) {
      const int j = s * 32 + jj;
      const double xj = shfl_d(r[s], jj);
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

}  // namespace b200qp
