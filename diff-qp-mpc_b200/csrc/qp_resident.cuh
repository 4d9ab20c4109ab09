// qp_resident.cuh -- the RESIDENT route of the PDIPM forward: several iterations of one QP per launch with the
// problem's matrices held on chip (fp64, nineq < 64, nz <= 32, neq == 0: the headline shape of BASELINE.json).
//
// Reference being re-implemented: qpth/solvers/pdipm/batch.py:46-214 (forward, get_step), :351-374 (solve_kkt),
// :434-469 (factor_kkt).  Same elimination as qp_fast.cuh, written the way the reference writes it:
//     t = Q^-1 rx;  hz = G t + rs/d - rz;  w = -T^-1 hz  (T = R + diag(1/d));  dx = Q^-1(-rx - G^T w);
//     ds = (-rs - w)/d;  dz = w          (batch.py:355-372; Q^-1 is the explicit inverse of the pre-factorisation)
// so the only matrices an iteration touches are Q, G, Q^-1 (shared memory, staged ONCE per launch) and R (read
// from L2 into the DMMA accumulator tiles).  The affine dx is never formed (only dz_aff / ds_aff enter Mehrotra's
// centering), which removes two of the nine mat-vecs of an iteration.
//
// Why a launch can run several iterations although the reference loop is coupled across the batch:
//   * termination (batch.py:127-144) only decides WHICH iterate is returned.  Every problem records its iterate
//     and residual of every iteration (history); k_res_finish reproduces the reference's stopping iteration from
//     the per-iteration batch reductions and picks each problem's best iterate before it.
//   * get_step's fill (batch.py:211-214) is max(1, a.max()) over the whole batch.  For a problem with at least
//     one non-positive direction entry the fill never binds unless a.max() is NaN, in which case it is exactly
//     1.0.  "Some ratio of the batch is NaN" is monotone in the iteration (a NaN iterate stays NaN), so the only
//     unknown is the first such iteration K*.  A launch speculates "not yet", publishes NaN events through one
//     atomic (Control::kev_inv), and records for every iteration whether the applied step length depended on the
//     guess.  The next launch restarts exactly the problems whose guess was wrong from the first such iteration
//     (their iterate of that iteration is in the history); from K* on the fill is known to be 1.0 and nothing is
//     speculated any more.  Fill-only rows while the fill is still a finite unknown, and z / s ratios whose NaN
//     onsets differ, are not speculated at all: they raise Control::need_exact and the call falls back to the
//     one-launch-per-iteration kernels of qp_fast.cuh (never observed on the benchmark distributions; forced in
//     tests/test_qp_resident_gpu.py).
//   The chunked run therefore returns bit-for-bit what the same kernels return with one iteration per launch
//   (tested), and the iteration count of the reference.
#pragma once
#include <limits.h>
#include "qp_fast.cuh"

namespace b200qp {

constexpr int kResMaxIter = 30;      // history masks are 32-bit
constexpr int kKevBase = 1 << 20;    // Control::kev_inv = kKevBase - (first iteration with a NaN ratio), 0 = none
constexpr int kPst = 8;              // ints of per-problem state

// per-problem state (ints, zero-initialised by the host at the start of a forward call)
//  [0] stage: 0 = initial point still to do, else 1 + next iteration      [1] 0, or 2 + iteration whose factor failed
//  [2] used  : bit it = the step of iteration it was taken with fill = 1.0 (NaN regime)
//  [3] sens  : bit it = the step length of iteration it differs between the two regimes
//  [4] fo    : bit it = a ratio test of iteration it was fill-only while the fill was speculated
//  [5] aznan / [6] asnan : bit it = this problem's z / s ratios of iteration it contain a NaN
struct ResOff {  // shared-memory carve-up of k_res_chunk, offsets in doubles (res_off)
  int G, Q, Qi, Up, pinv, Pb, x, rx, t, p, qx, s, z, d, di, rz, hz, dz, h, part, misc, total;
};

struct RArgs {
  ResOff off;    // for run-time sizes (the compile-time specialisations ignore it)
  double* hist;  // [nb][max_iter + 1][hs]   x | s | z at the START of each iteration
  double* rec;   // [nb][max_iter][2]        residual, mu of each iteration
  int* pst;      // [nb][kPst]
  int it_end;    // this launch runs iterations < it_end
  int hs;
};

__host__ __device__ inline int res_hs(int n, int m) { return round4(n) + 2 * round4(m); }


__device__ __forceinline__ void cp_async8(double* dst_smem, const double* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(src));
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += shfl_x(v, o);
  return v;
}


// ------------------------------------------------------------------------------------------------------------
// Shared-memory carve-up (offsets in doubles).  With compile-time sizes (the <NC, MC> specialisations of the
// kernel) every offset is an immediate; otherwise the host computes it once and it travels in the kernel
// arguments (constant bank) -- ncu on the first version of this kernel showed 31 % of all executed instructions
// re-deriving these sums of rounded sizes.
__host__ __device__ constexpr int res_r4(int v) { return (v + 3) & ~3; }
__host__ __device__ constexpr ResOff res_off(int n, int m, int mpad) {
  ResOff o{};
  const int ldn = n | 1;
  int q = 0;
  o.G = q; q += res_r4(m * ldn);
  o.Q = q; q += res_r4(n * ldn);
  o.Qi = q; q += res_r4(n * ldn);
  o.Up = q; q += res_r4((m + 1) * m / 2 + 1);
  o.pinv = q; q += res_r4(m);
  o.Pb = q; q += res_r4(mpad * kPanelStride + 16);
  o.x = q; q += res_r4(n); o.rx = q; q += res_r4(n); o.t = q; q += res_r4(n); o.p = q; q += res_r4(n); o.qx = q; q += res_r4(n);
  o.s = q; q += res_r4(m); o.z = q; q += res_r4(m); o.d = q; q += res_r4(m); o.di = q; q += res_r4(m);
  o.rz = q; q += res_r4(m); o.hz = q; q += res_r4(m); o.dz = q; q += res_r4(m); o.h = q; q += res_r4(m);
  o.part = q; q += 128;
  o.misc = q; q += 16;
  o.total = q;
  return o;
}

// ------------------------------------------------------------------------------------------------------------
// Blocked right-looking LDL^T of T = R + diag(dinv), bordered by the row hz at index m, on the FP64 tensor cores
// (same algorithm as dmma_factor, qp_dmma.cuh).  Differences that cut the executed instructions by ~3x:
//   * tiles are owned by ROW: warp w holds tile rows w and NTI-1-w (9 tiles each at MPAD = 64), slot index = tile
//     column, so every register-array index is static and a warp loads the scaled A fragment of a row once per
//     panel and re-uses it (and the B fragment of a column) for all its tiles;
//   * operands are 16-byte fragments: the contraction index of the two DMMAs of a tile is permuted so that a
//     lane's two elements are adjacent columns;
//   * with compile-time sizes the panel loop is unrolled, so offsets are immediates and predicates fold.
template <int MPAD>
struct RowTiles {
  static constexpr int NTI = MPAD / 8;
  static constexpr int NA = NTI == 8 ? 4 : NTI;  // slots of row a = warp        (tile column K <= warp)
  static constexpr int NB = NTI == 8 ? 8 : 1;    // slots of row b = NTI-1-warp  (MPAD = 64 only)
  static constexpr bool HAS_B = NTI == 8;
};

template <int MPAD>
__device__ __forceinline__ void res_prefetch(const double* __restrict__ Rf, int lane, int warp,
                                             double (&Ca)[RowTiles<MPAD>::NA][2], double (&Cb)[RowTiles<MPAD>::NB][2]) {
  using RT = RowTiles<MPAD>;
  const int Ia = warp, Ib = RT::NTI - 1 - warp;
  const double* ra = Rf + (size_t)(Ia * (Ia + 1) / 2) * 64 + lane * 2;
#pragma unroll
  for (int K = 0; K < RT::NA; K++) {
    double2 v = make_double2(0.0, 0.0);
    if (K <= Ia) v = *reinterpret_cast<const double2*>(ra + K * 64);
    Ca[K][0] = v.x; Ca[K][1] = v.y;
  }
  if constexpr (RT::HAS_B) {
    const double* rb = Rf + (size_t)(Ib * (Ib + 1) / 2) * 64 + lane * 2;
#pragma unroll
    for (int K = 0; K < RT::NB; K++) {
      double2 v = make_double2(0.0, 0.0);
      if (K <= Ib) v = *reinterpret_cast<const double2*>(rb + K * 64);
      Cb[K][0] = v.x; Cb[K][1] = v.y;
    }
  }
}

// diagonal, bordered row and padding of one tile that is not interior
__device__ __forceinline__ void res_fix_tile(double (&c)[2], int I, int K, int fr, int fc, int m, const double* dinv,
                                             const double* hz) {
  if (!(I != K && 8 * I + 8 <= m)) {
    const int i = 8 * I + fr, k = 8 * K + fc;
    double v0 = 0.0, v1 = 0.0;
    if (i < m && k < m) {
      v0 = (k <= i) ? c[0] : 0.0;
      v1 = (k + 1 <= i) ? c[1] : 0.0;
      if (I == K) {
        if (i == k) v0 += dinv[i];
        if (i == k + 1) v1 += dinv[i];
      }
    } else if (i == m) {
      if (k < m) v0 = hz[k];
      if (k + 1 < m) v1 = hz[k + 1];
    }
    if (i >= m) {
      if (i == k) v0 = 1.0;
      if (i == k + 1) v1 = 1.0;
    }
    c[0] = v0; c[1] = v1;
  }
}

template <int MPAD, int MC, class Idle>
__device__ __forceinline__ bool res_factor(double (&Ca)[RowTiles<MPAD>::NA][2], double (&Cb)[RowTiles<MPAD>::NB][2],
                                           const double* dinv, const double* hz, double* Up, double* pinv, double* Pb,
                                           int m_rt, int tid, Idle&& idle) {
  using RT = RowTiles<MPAD>;
  constexpr int PS = kPanelStride;
  const int m = MC > 0 ? MC : m_rt;
  const int lane = tid & 31, warp = tid >> 5;
  const int fr = lane >> 2, kc = lane & 3, fc = kc * 2;
  const int mm = m + 1;
  const int Ia = warp, Ib = RT::NTI - 1 - warp;
#pragma unroll
  for (int K = 0; K < RT::NA; K++)
    if (K <= Ia) res_fix_tile(Ca[K], Ia, K, fr, fc, m, dinv, hz);
  if constexpr (RT::HAS_B) {
#pragma unroll
    for (int K = 0; K < RT::NB; K++)
      if (K <= Ib) res_fix_tile(Cb[K], Ib, K, fr, fc, m, dinv, hz);
  }
  double* ppan = Pb + MPAD * PS;  // [0..8): -1/D of the panel's columns (trailing-update scale); [8..16): 1/D
  double* rowa = Pb + (8 * Ia + fr) * PS + fc;  // this lane's fragment of tile row a / b in the panel buffer
  double* rowb = Pb + (8 * Ib + fr) * PS + fc;
  const double* colk = Pb + fr * PS + fc;        // + 8 K PS: fragment of tile row K (the B operand)
  bool ok = true;
  const int npan = (mm + 7) >> 3;
#pragma unroll
  for (int J = 0; J < (MC > 0 ? (MC + 8) / 8 : RT::NTI); J++) {
    if (J < npan) {  // folds for compile-time sizes
      // (a) owners publish the panel's tiles
#pragma unroll
      for (int K = 0; K < RT::NA; K++)
        if (K == J && K <= Ia) *reinterpret_cast<double2*>(rowa) = make_double2(Ca[K][0], Ca[K][1]);
      if constexpr (RT::HAS_B) {
#pragma unroll
        for (int K = 0; K < RT::NB; K++)
          if (K == J && K <= Ib) *reinterpret_cast<double2*>(rowb) = make_double2(Cb[K][0], Cb[K][1]);
      }
      __syncthreads();
      // (b) one warp factors the panel: lane owns rows i0, i1
      if (warp == 0) {
        const int rbase = 8 * J;
        const int i0 = rbase + lane, i1 = rbase + 32 + lane;
        const bool two = rbase + 32 < MPAD;  // the second row of a lane exists
        double a0[8], a1[8];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          double2 v = make_double2(0.0, 0.0), w = make_double2(0.0, 0.0);
          if (i0 < MPAD) v = *reinterpret_cast<const double2*>(Pb + i0 * PS + 2 * q);
          if (two && i1 < MPAD) w = *reinterpret_cast<const double2*>(Pb + i1 * PS + 2 * q);
          a0[2 * q] = v.x; a0[2 * q + 1] = v.y;
          a1[2 * q] = w.x; a1[2 * q + 1] = w.y;
        }
        double pk[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const double dk = shfl_d(a0[k], k);
          double r;
          asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(dk));
          double e = fma(-dk, r, 1.0);
          r = fma(r, e, r);
          e = fma(-dk, r, 1.0);
          r = fma(r, e, r);
          if (!(dk > 1e-290 && dk < 1e290)) r = (dk > 0.0) ? 1.0 / dk : t_nan<double>();  // rare
          const double pkk = (rbase + k < m) ? r : 1.0;  // bordered / padding columns: unit pivot
          pk[k] = pkk;
          const double w0 = a0[k], w1 = a1[k];
          const double l0 = w0 * pkk, l1 = w1 * pkk;
#pragma unroll
          for (int c = k + 1; c < 8; c++) {
            const double wck = shfl_d(w0, c);
            a0[c] -= l0 * wck;
            if (two) a1[c] -= l1 * wck;
          }
        }
        // off the pivot chain: reciprocal pivots, packed unit-lower columns for the sweeps, W = L D for the update
        if (lane < 8) {
          double pl = pk[0];
#pragma unroll
          for (int k = 1; k < 8; k++) pl = lane == k ? pk[k] : pl;
          ppan[lane] = -pl; ppan[8 + lane] = pl;
          if (rbase + lane < m) pinv[rbase + lane] = pl;
        }
        {
          int ub = urow(rbase, mm);
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const int j = rbase + k;
            if (j < m) {
              if (i0 > j && i0 < mm) Up[ub + i0] = a0[k] * pk[k];
              if (two && i1 < mm) Up[ub + i1] = a1[k] * pk[k];
            }
            ub += mm - j - 2;
          }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
          if (i0 >= rbase + 8 && i0 < MPAD) *reinterpret_cast<double2*>(Pb + i0 * PS + 2 * q) = make_double2(a0[2 * q], a0[2 * q + 1]);
          if (two && i1 < MPAD) *reinterpret_cast<double2*>(Pb + i1 * PS + 2 * q) = make_double2(a1[2 * q], a1[2 * q + 1]);
        }
      } else {
        idle(J);
      }
      __syncthreads();
      // (c) trailing update  C_IK -= (W_I D^-1) W_K^T  of this warp's tiles right of the panel
      {
        const double2 sc = *reinterpret_cast<const double2*>(ppan + fc);
        if (is_nan(ppan[0] + ppan[1] + ppan[2] + ppan[3] + ppan[4] + ppan[5] + ppan[6] + ppan[7])) ok = false;  // uniform
        double2 aa = make_double2(0.0, 0.0), ab = make_double2(0.0, 0.0);
        if (J < Ia) { aa = *reinterpret_cast<const double2*>(rowa); aa.x *= sc.x; aa.y *= sc.y; }
        if (RT::HAS_B && J < Ib) { ab = *reinterpret_cast<const double2*>(rowb); ab.x *= sc.x; ab.y *= sc.y; }
#pragma unroll
        for (int K = 1; K < RT::NTI; K++) {
          if (K > J && (K <= Ia || (RT::HAS_B && K <= Ib))) {
            const double2 bv = *reinterpret_cast<const double2*>(colk + 8 * K * PS);
            if (K < RT::NA && K <= Ia) {
              dmma_m8n8k4(Ca[K < RT::NA ? K : 0][0], Ca[K < RT::NA ? K : 0][1], aa.x, bv.x);
              dmma_m8n8k4(Ca[K < RT::NA ? K : 0][0], Ca[K < RT::NA ? K : 0][1], aa.y, bv.y);
            }
            if (RT::HAS_B && K <= Ib) {
              dmma_m8n8k4(Cb[K < RT::NB ? K : 0][0], Cb[K < RT::NB ? K : 0][1], ab.x, bv.x);
              dmma_m8n8k4(Cb[K < RT::NB ? K : 0][0], Cb[K < RT::NB ? K : 0][1], ab.y, bv.y);
            }
          }
        }
      }
      __syncthreads();
    }
  }
  return ok;
}

// ------------------------------------------------------------------------------------------------------------
// Triangular sweeps of ONE warp with the working vector in registers (lane owns rows lane, lane + 32), packed
// unit-lower factor `Up` (urow, qp_dmma.cuh).  With a compile-time size the column loop is unrolled: the row
// offsets are immediates and most row predicates fold (6-7 instructions per column).
template <int RPL, int MC>
__device__ __forceinline__ void res_fwd(const double* Up, int m_rt, double (&r)[RPL], int lane) {
  const int m = MC > 0 ? MC : m_rt, mm = m + 1;
  if constexpr (MC > 0) {
#pragma unroll
    for (int j = 0; j < MC; j++) {
      const int s = j >> 5;
      const double yj = shfl_d(r[s], j & 31);
      const double* row = Up + urow(j, MC + 1);
#pragma unroll
      for (int s2 = s; s2 < RPL; s2++) {
        const int i = s2 * 32 + lane;
        if (i > j && i < MC) r[s2] -= row[i] * yj;
      }
    }
  } else {
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
  }
}

template <int RPL, int MC>
__device__ __forceinline__ void res_bwd(const double* Up, int m_rt, double (&r)[RPL], int lane) {
  const int m = MC > 0 ? MC : m_rt, mm = m + 1;
  const double* rb[RPL];
#pragma unroll
  for (int s = 0; s < RPL; s++) {
    const int i = s * 32 + lane;
    rb[s] = Up + urow(i < m ? i : 0, mm);  // rb[s][j] = U[i][j]
  }
  if constexpr (MC > 0) {
#pragma unroll
    for (int j = MC - 1; j > 0; j--) {
      const int s = j >> 5;
      const double xj = shfl_d(r[s], j & 31);
#pragma unroll
      for (int s2 = 0; s2 <= s; s2++) {
        const int i = s2 * 32 + lane;
        if (i < j) r[s2] -= rb[s2][j] * xj;
      }
    }
  } else {
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
          if (i < j) r[s2] -= rb[s2][j] * xj;
        }
      }
    }
  }
}

// get_step pieces (batch.py:211-214) of (z, dz) and (s, ds) from registers: the NaN-propagating minimum over the
// entries the fill does not overwrite (+inf if none), whether some entry is overwritten, and (AMAX) the
// NaN-propagating maximum of a = -v / dv over all entries.  Every lane returns the same values.
template <int RPL, bool AMAX>
__device__ __forceinline__ void res_pieces(const double (&z)[RPL], const double (&dz)[RPL], const double (&s)[RPL],
                                           const double (&ds)[RPL], int m, int lane, double (&out)[4], int& has) {
  double rz = t_inf<double>(), rs = t_inf<double>(), az = -t_inf<double>(), as = -t_inf<double>();
  bool hz = false, hs = false;
#pragma unroll
  for (int q = 0; q < RPL; q++) {
    if (q * 32 + lane < m) {
      const double a1 = -z[q] / dz[q], a2 = -s[q] / ds[q];
      if (dz[q] > 0.0) hz = true; else rz = nanmin(rz, a1);
      if (ds[q] > 0.0) hs = true; else rs = nanmin(rs, a2);
      if (AMAX) { az = nanmax(az, a1); as = nanmax(as, a2); }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    rz = nanmin(rz, shfl_x(rz, o));
    rs = nanmin(rs, shfl_x(rs, o));
    if (AMAX) { az = nanmax(az, shfl_x(az, o)); as = nanmax(as, shfl_x(as, o)); }
  }
  has = (__any_sync(0xffffffffu, hz) ? 1 : 0) | (__any_sync(0xffffffffu, hs) ? 2 : 0);
  out[0] = rz; out[1] = rs; out[2] = az; out[3] = as;
}

// out[row] = sum_c M[row * ld + c] v[c] for the rows of a (rows <= 64) x (cols <= 32) matrix in shared memory with an
// odd leading dimension: two threads per row (the halves of a warp take columns [0,16) / [16,32)), combined by
// one shuffle.  Returns the sum in every thread; row = warp * 16 + (lane & 15).
__device__ __forceinline__ double res_mv2(const double* M, int ld, int rows, int cols, const double* v, int lane, int warp) {
  const int row = warp * 16 + (lane & 15), c0 = (lane >> 4) * 16;
  double a0 = 0.0, a1 = 0.0;
  if (row < rows) {
    const double* mr = M + row * ld + c0;
    const double* vv = v + c0;
    const int cnt = cols - c0;  // may be <= 0
#pragma unroll
    for (int c = 0; c < 16; c += 2) {
      if (c < cnt) a0 = fma(mr[c], vv[c], a0);
      if (c + 1 < cnt) a1 = fma(mr[c + 1], vv[c + 1], a1);
    }
  }
  double a = a0 + a1;
  a += shfl_x(a, 16);
  return a;
}
// Same for a (rows <= 32) x (cols <= 32) matrix with four threads per row (quarter warps take eight columns
// each); row = warp * 8 + (lane & 7).
__device__ __forceinline__ double res_mv4(const double* M, int ld, int rows, int cols, const double* v, int lane, int warp) {
  const int row = warp * 8 + (lane & 7), c0 = (lane >> 3) * 8;
  double a0 = 0.0, a1 = 0.0;
  if (row < rows) {
    const double* mr = M + row * ld + c0;
    const double* vv = v + c0;
    const int cnt = cols - c0;
#pragma unroll
    for (int c = 0; c < 8; c += 2) {
      if (c < cnt) a0 = fma(mr[c], vv[c], a0);
      if (c + 1 < cnt) a1 = fma(mr[c + 1], vv[c + 1], a1);
    }
  }
  double a = a0 + a1;
  a += shfl_x(a, 8);
  a += shfl_x(a, 16);
  return a;
}
// part[warp * 32 + c] = sum over the rows r = warp, warp + 4, ... of M[r * ld + c] u[r]   (c = lane < cols)
template <int MC>
__device__ __forceinline__ void res_mvt(const double* M, int ld, int rows, int cols, const double* u, double* part, int lane,
                                        int warp) {
  double a0 = 0.0, a1 = 0.0;
  if (lane < cols) {
    if constexpr (MC > 0) {
      const double* mp = M + warp * ld + lane;
      const double* up = u + warp;
#pragma unroll
      for (int k = 0; k < (MC + 3) / 4; k += 2) {
        if (4 * k + 3 < MC || warp + 4 * k < MC) a0 = fma(mp[4 * k * ld], up[4 * k], a0);
        if (k + 1 < (MC + 3) / 4 && (4 * k + 7 < MC || warp + 4 * k + 4 < MC)) a1 = fma(mp[(4 * k + 4) * ld], up[4 * k + 4], a1);
      }
    } else {
      int r = warp;
      for (; r + 4 < rows; r += 8) {
        a0 = fma(M[r * ld + lane], u[r], a0);
        a1 = fma(M[(r + 4) * ld + lane], u[r + 4], a1);
      }
      if (r < rows) a0 = fma(M[r * ld + lane], u[r], a0);
    }
  }
  part[warp * 32 + lane] = a0 + a1;
}

// ------------------------------------------------------------------------------------------------------------
// One launch = the iterations [.., ra.it_end) of every problem.  <NC, MC> != 0: compile-time nz / nineq.
template <int MPAD, int NC, int MC>
__global__ void __launch_bounds__(128, 4) k_res_chunk(const KArgs<double> a, const RArgs ra) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using RT = RowTiles<MPAD>;
  constexpr int RPL = MPAD / 32;
  const int prob = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = NC > 0 ? NC : a.n, m = MC > 0 ? MC : a.m, mm = m + 1, ldn = n | 1;
  const int n4 = res_r4(n), m4 = res_r4(m), hs = n4 + 2 * m4;

  // ---- where does this problem (re)start?
  int* ps = ra.pst + (size_t)prob * kPst;
  int stage = ps[0], poison = ps[1];
  unsigned used = (unsigned)ps[2], sens = (unsigned)ps[3], fo = (unsigned)ps[4], azm = (unsigned)ps[5], asm_ = (unsigned)ps[6];
  const unsigned kevi = *(volatile unsigned*)&a.ctl->kev_inv;
  const int Kstart = kevi ? kKevBase - (int)kevi : INT_MAX;  // exact for every iteration of the completed launches
  int it = stage - 1;
  if (stage > 1 && Kstart < it) {
    const unsigned bad = sens & ~used & ~((1u << Kstart) - 1u) & ((1u << it) - 1u);
    if (bad) {  // the step of iteration j was taken with the wrong fill: redo from there
      it = __ffs(bad) - 1;
      const unsigned keep = (1u << it) - 1u;
      used &= keep; sens &= keep; fo &= keep; azm &= keep; asm_ &= keep;
      poison = 0;
    }
  }
  if (poison || it >= ra.it_end) return;  // uniform

  // ---- shared memory
  const ResOff o = (NC > 0 && MC > 0) ? res_off(NC > 0 ? NC : 1, MC > 0 ? MC : 1, MPAD) : ra.off;
  double* sm = reinterpret_cast<double*>(smem_raw);
  double* sG = sm + o.G; double* sQ = sm + o.Q; double* sQi = sm + o.Qi;
  double* Up = sm + o.Up; double* pinv = sm + o.pinv; double* Pb = sm + o.Pb;
  double* vx = sm + o.x; double* vrx = sm + o.rx; double* vt = sm + o.t; double* vp = sm + o.p; double* vqx = sm + o.qx;
  double* vs = sm + o.s; double* vz = sm + o.z; double* vd = sm + o.d; double* vdi = sm + o.di; double* vrz = sm + o.rz;
  double* vhz = sm + o.hz; double* vdz = sm + o.dz; double* vh = sm + o.h;
  double* part = sm + o.part; double* misc = sm + o.misc;
  double* hist = ra.hist + (size_t)prob * (a.max_iter + 1) * hs;

  // ---- stage Q, G (row stride n -> odd ldn), Q^-1, p, h and the iterate
  {
    const double* Qg = a.Q + (size_t)prob * a.sQ;
    const double* Gg = a.G + (size_t)prob * a.sG;
    for (int r = warp; r < m; r += 4)
      if (lane < n) cp_async8(sG + r * ldn + lane, Gg + r * n + lane);
    for (int r = warp; r < n; r += 4)
      if (lane < n) cp_async8(sQ + r * ldn + lane, Qg + r * n + lane);
    cp_async_block(sQi, a.Qi + (size_t)prob * a.sQi, res_r4(n * ldn), tid, 128);
    cp_async_commit();
    const double* pg = a.pv + (size_t)prob * a.sp;
    const double* hg = a.h + (size_t)prob * a.sh;
    if (tid < n) vp[tid] = pg[tid];
    if (tid < m) vh[tid] = hg[tid];
    if (it >= 0) {
      const double* hh = hist + (size_t)it * hs;
      if (tid < n) vx[tid] = hh[tid];
      if (tid < m) { vs[tid] = hh[n4 + tid]; vz[tid] = hh[n4 + m4 + tid]; }
    }
  }
  const double* Rf = a.R + (size_t)prob * a.sR;
  double Ca[RT::NA][2], Cb[RT::NB][2];
  res_prefetch<MPAD>(Rf, lane, warp, Ca, Cb);
  cp_async_wait_all();
  __syncthreads();

  bool alive = true;
  bool have_hist = true;  // the history entry of the first iteration of this launch already exists
#pragma unroll 1
  for (; it < ra.it_end; ++it) {
    const bool init = it < 0;
    // ---------------- residuals (batch.py:93-108)
    if (!init) {
      if (!have_hist) {
        double* hh = hist + (size_t)it * hs;
        if (tid < n) hh[tid] = vx[tid];
        if (tid < m) { hh[n4 + tid] = vs[tid]; hh[n4 + m4 + tid] = vz[tid]; }
      }
      if (tid < m) {
        const double dv = vz[tid] / vs[tid];
        vd[tid] = dv;
        vdi[tid] = 1.0 / dv;
      }
      const double gx = res_mv2(sG, ldn, m, n, vx, lane, warp);
      {
        const int row = warp * 16 + (lane & 15);
        if (lane < 16 && row < m) vrz[row] = gx + vs[row] - vh[row];
      }
      const double qx = res_mv4(sQ, ldn, n, n, vx, lane, warp);
      {
        const int row = warp * 8 + (lane & 7);
        if (lane < 8 && row < n) vqx[row] = qx;
      }
      res_mvt<MC>(sG, ldn, m, n, vz, part, lane, warp);
      __syncthreads();
      if (tid < n) vrx[tid] = ((part[tid] + part[32 + tid]) + (part[64 + tid] + part[96 + tid])) + (vqx[tid] + vp[tid]);
    } else {
      // initial point (batch.py:60-66): d = 1, (rx, rs, rz) = (p, 0, -h)
      if (tid < n) vrx[tid] = vp[tid];
      if (tid < m) { vd[tid] = 1.0; vdi[tid] = 1.0; vz[tid] = 0.0; vs[tid] = 0.0; vrz[tid] = -vh[tid]; }
    }
    have_hist = false;
    __syncthreads();
    // ---------------- right-hand side of the reduced system: hz = G Q^-1 rx + rs / d - rz   (rs = z)
    {
      const double t = res_mv4(sQi, ldn, n, n, vrx, lane, warp);
      const int row = warp * 8 + (lane & 7);
      if (lane < 8 && row < n) vt[row] = t;
    }
    __syncthreads();
    {
      const double gt = res_mv2(sG, ldn, m, n, vt, lane, warp);
      const int row = warp * 16 + (lane & 15);
      if (lane < 16 && row < m) vhz[row] = vz[row] / vd[row] + (gt - vrz[row]);
    }
    __syncthreads();
    // ---------------- T = R + diag(1/d) = L D L^T with hz riding as the bordered row
    const bool ok = res_factor<MPAD, MC>(Ca, Cb, vdi, vhz, Up, pinv, Pb, m, tid, [&](int J) {
      if (warp == 1 && J == 0 && !init) {
        // residual norms and mu by an otherwise idle warp (batch.py:103-108)
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        if (lane < n) s0 = vrx[lane] * vrx[lane];
#pragma unroll
        for (int qq = 0; qq < RPL; qq++) {
          const int i = qq * 32 + lane;
          if (i < m) { s1 = fma(vrz[i], vrz[i], s1); s2 = fma(vs[i], vz[i], s2); }
        }
        s0 = warp_sum_d(s0); s1 = warp_sum_d(s1); s2 = warp_sum_d(s2);
        const double mu = fabs(s2 / (double)m);
        const double resid = sqrt(s1) + sqrt(s0) + (double)m * mu;
        if (lane == 0) {
          misc[0] = mu; misc[1] = s2;
          double* rc = ra.rec + ((size_t)prob * a.max_iter + it) * 2;
          rc[0] = resid; rc[1] = mu;
        }
      }
    });
    // the tiles of the next iteration travel while one warp runs the sweeps
    res_prefetch<MPAD>(Rf, lane, warp, Ca, Cb);
    if (!ok) {
      // non-positive / NaN pivot: this problem can never improve again (qp_common.cuh header); its ratios count
      // as NaN from this iteration on
      if (tid == 0) {
        ps[1] = it + 2;
        atomicMax(&a.ctl->kev_inv, (unsigned)(kKevBase - (it < 0 ? 0 : it)));
      }
      alive = false;
      break;
    }
    // ---------------- one warp: predictor, centering, corrector, step length
    if (warp == 0) {
      double zr[RPL], sr[RPL], dr[RPL], qa[RPL];
#pragma unroll
      for (int s = 0; s < RPL; s++) {
        const int i = s * 32 + lane;
        zr[s] = i < m ? vz[i] : 1.0; sr[s] = i < m ? vs[i] : 1.0; dr[s] = i < m ? vd[i] : 1.0;
        qa[s] = i < m ? Up[urow(i, mm) + m] : 0.0;  // D^-1 L^-1 hz: the bordered row of the factor
      }
      res_bwd<RPL, MC>(Up, m, qa, lane);
      if (init) {
        // x, s, z of the initial point; shift s and z so that their minima are >= 1 (batch.py:76-86)
        double mn_s = t_inf<double>(), mn_z = t_inf<double>();
#pragma unroll
        for (int s = 0; s < RPL; s++) {
          if (s * 32 + lane < m) { mn_s = nanmin(mn_s, qa[s]); mn_z = nanmin(mn_z, -qa[s]); }
        }
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) { mn_s = nanmin(mn_s, shfl_x(mn_s, o2)); mn_z = nanmin(mn_z, shfl_x(mn_z, o2)); }
#pragma unroll
        for (int s = 0; s < RPL; s++) {
          const int i = s * 32 + lane;
          if (i < m) {
            double sv = qa[s], zv = -qa[s];
            if (mn_s < 0.0) sv -= mn_s - 1.0;
            if (mn_z < 0.0) zv -= mn_z - 1.0;
            vdz[i] = -qa[s];
            vs[i] = sv; vz[i] = zv;
          }
        }
        if (lane == 0) misc[2] = 1.0;
      } else {
        double dza[RPL], dsa[RPL];
#pragma unroll
        for (int s = 0; s < RPL; s++) { dza[s] = -qa[s]; dsa[s] = (-zr[s] - dza[s]) / dr[s]; }
        double pc[4]; int has;
        res_pieces<RPL, false>(zr, dza, sr, dsa, m, lane, pc, has);
        // the clamp at 1 makes alpha_aff independent of the batch-global fill (batch.py:161-163)
        const double stz = (has & 1) ? nanmin(pc[0], 1.0) : pc[0];
        const double sts = (has & 2) ? nanmin(pc[1], 1.0) : pc[1];
        const double alpha_aff = nanmin(nanmin(stz, sts), 1.0);
        double t3 = 0.0;
#pragma unroll
        for (int s = 0; s < RPL; s++)
          if (s * 32 + lane < m) t3 += (sr[s] + alpha_aff * dsa[s]) * (zr[s] + alpha_aff * dza[s]);
        t3 = warp_sum_d(t3);
        const double mu = misc[0], t4 = misc[1];
        const double ratio = t3 / t4;
        const double sig = ratio * ratio * ratio;
        double rsc[RPL], qc[RPL];
#pragma unroll
        for (int s = 0; s < RPL; s++) {
          rsc[s] = (-mu * sig + dsa[s] * dza[s]) / sr[s];
          qc[s] = (s * 32 + lane < m) ? rsc[s] / dr[s] : 0.0;
        }
        res_fwd<RPL, MC>(Up, m, qc, lane);
#pragma unroll
        for (int s = 0; s < RPL; s++) { const int i = s * 32 + lane; if (i < m) qc[s] *= pinv[i]; }
        res_bwd<RPL, MC>(Up, m, qc, lane);
        double dz[RPL], ds[RPL];
#pragma unroll
        for (int s = 0; s < RPL; s++) {
          const double dzc = -qc[s];
          const double dsc = (-rsc[s] - dzc) / dr[s];
          dz[s] = dza[s] + dzc; ds[s] = dsa[s] + dsc;
        }
        res_pieces<RPL, true>(zr, dz, sr, ds, m, lane, pc, has);
        const bool nz_ = is_nan(pc[2]), ns_ = is_nan(pc[3]);
        if (nz_ || ns_) {
          if (lane == 0) atomicMax(&a.ctl->kev_inv, (unsigned)(kKevBase - it));
          if (nz_) azm |= 1u << it;
          if (ns_) asm_ |= 1u << it;
        }
        // fill regime of this iteration: exact when it >= Kstart, else the freshest published event
        int Kdyn = Kstart;
        {
          const unsigned kv = *(volatile unsigned*)&a.ctl->kev_inv;
          const int Kn = kv ? kKevBase - (int)kv : INT_MAX;
          Kdyn = Kn < Kdyn ? Kn : Kdyn;
        }
        const bool F = it >= Kdyn;
        // F = 0: fill = max(1, a.max()) >= every unfilled ratio -> the unfilled minimum (+inf when fill-only:
        //        0.999 * fill >= 1 is validated by k_res_reduce);  F = 1: fill = 1.0 exactly
        const double a0_ = nanmin(0.999 * nanmin(pc[0], pc[1]), 1.0);
        const double z1 = (has & 1) ? nanmin(pc[0], 1.0) : pc[0], s1 = (has & 2) ? nanmin(pc[1], 1.0) : pc[1];
        const double a1_ = nanmin(0.999 * nanmin(z1, s1), 1.0);
        const bool same = (a0_ == a1_) || (is_nan(a0_) && is_nan(a1_));
        if (!same) sens |= 1u << it;
        if (F) used |= 1u << it;
        else if (((has & 1) && pc[0] == t_inf<double>()) || ((has & 2) && pc[1] == t_inf<double>())) fo |= 1u << it;
        const double alpha = F ? a1_ : a0_;
#pragma unroll
        for (int s = 0; s < RPL; s++) {
          const int i = s * 32 + lane;
          if (i < m) { vdz[i] = dz[s]; vs[i] = sr[s] + alpha * ds[s]; vz[i] = zr[s] + alpha * dz[s]; }
        }
        if (lane == 0) misc[2] = alpha;
      }
    }
    __syncthreads();
    // ---------------- dx = Q^-1 (-rx - G^T dz);  x += alpha dx   (initial point: x = dx)
    res_mvt<MC>(sG, ldn, m, n, vdz, part, lane, warp);
    __syncthreads();
    if (tid < n) vt[tid] = -vrx[tid] - ((part[tid] + part[32 + tid]) + (part[64 + tid] + part[96 + tid]));
    __syncthreads();
    {
      const double dx = res_mv4(sQi, ldn, n, n, vt, lane, warp);
      const int row = warp * 8 + (lane & 7);
      if (lane < 8 && row < n) vx[row] = init ? dx : vx[row] + misc[2] * dx;
    }
    __syncthreads();
  }
  // ---- hand the state to the next launch
  if (alive) {
    double* hh = hist + (size_t)it * hs;
    if (tid < n) hh[tid] = vx[tid];
    if (tid < m) { hh[n4 + tid] = vs[tid]; hh[n4 + m4 + tid] = vz[tid]; }
  }
  // the masks live in warp 0's registers (every lane holds the same values)
  if (tid == 0) {
    ps[0] = (alive ? it : it + 1) + 1;
    if (alive) ps[1] = 0;
    ps[2] = (int)used; ps[3] = (int)sens; ps[4] = (int)fo; ps[5] = (int)azm; ps[6] = (int)asm_;
  }
}

// ------------------------------------------------------------------------------------------------------------
// After the last launch: the reference's batch reductions per iteration (batch.py:119-144) from the per-problem
// records, the stopping iteration, validation of everything that was speculated, and each problem's best iterate
// before the stop.  One launch, thread per problem for the reductions; the last CTA to finish evaluates the
// termination, then a second kernel copies the winners.
template <int U>
__global__ void k_res_reduce(const KArgs<double> a, const RArgs ra) {
  const int prob = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool act = prob < a.nb;
  int poison_it = INT_MAX;
  unsigned azm = 0, asm_ = 0, fo = 0, used = 0;
  if (act) {
    const int* ps = ra.pst + (size_t)prob * kPst;
    if (ps[1]) poison_it = ps[1] - 2;
    used = (unsigned)ps[2]; fo = (unsigned)ps[4]; azm = (unsigned)ps[5]; asm_ = (unsigned)ps[6];
  }
  if (__any_sync(0xffffffffu, act && (fo & ~used) != 0u)) { if (lane == 0) a.ctl->need_exact = 1; }
  double best = __longlong_as_double(0x7ff8000000000000LL);
  for (int it = 0; it < a.max_iter; it++) {
    Slot* slot = a.slots + it;
    bool improved = false, mu_nan = false, az = false, as = false;
    double mu = t_inf<double>();
    if (act) {
      if (it <= poison_it) {
        const double* rc = ra.rec + ((size_t)prob * a.max_iter + it) * 2;
        const double rd = rc[0];
        mu = rc[1];
        const bool better = (it == 0) ? true : (rd < best);
        if (better) { best = rd; improved = it > 0; }
        mu_nan = mu != mu;
        az = (azm >> it) & 1u; as = (asm_ >> it) & 1u;
        if (it == poison_it) { az = true; as = true; }
      } else {
        mu_nan = true; az = true; as = true;  // poisoned: NaN iterate, never improves
      }
    }
    const bool bn = act && (best != best);
    unsigned long long bkey = (act && !bn) ? (unsigned long long)__double_as_longlong(best) : 0ULL;
    unsigned long long mkey = (act && !mu_nan) ? ~ord_key(mu) : 0ULL;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long b2 = __shfl_xor_sync(0xffffffffu, bkey, o), m2 = __shfl_xor_sync(0xffffffffu, mkey, o);
      bkey = b2 > bkey ? b2 : bkey;
      mkey = m2 > mkey ? m2 : mkey;
    }
    const bool any_imp = __any_sync(0xffffffffu, improved), any_bn = __any_sync(0xffffffffu, bn);
    const bool any_mn = __any_sync(0xffffffffu, act && mu_nan), any_az = __any_sync(0xffffffffu, az), any_as = __any_sync(0xffffffffu, as);
    if (lane == 0) {
      if (bkey) atomic_max_key(&slot->best_max, bkey);
      if (mkey) atomic_max_key(&slot->mu_min_inv, mkey);
      if (any_imp && !slot->improved) slot->improved = 1;
      if (any_bn && !slot->best_nan) slot->best_nan = 1;
      if (any_mn && !slot->mu_nan) slot->mu_nan = 1;
      if (any_az && !slot->az_nan) slot->az_nan = 1;
      if (any_as && !slot->as_nan) slot->as_nan = 1;
    }
  }
}

template <int U>
__global__ void k_res_select(const KArgs<double> a, const RArgs ra, double* status, int launches) {
  __shared__ int s_niter;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp == 0) {
    const int term = eval_termination(a.slots, a.max_iter, a.lim, a.eps, lane);
    const int n_iter = term >= 0 ? term + 1 : a.max_iter;
    if (lane == 0) s_niter = n_iter;
    if (blockIdx.x == 0) {
      // z / s ratios that turn NaN at different iterations are not speculated (file header)
      bool mixed = false;
      for (int j = lane; j < n_iter; j += 32) mixed |= (a.slots[j].az_nan != 0) != (a.slots[j].as_nan != 0);
      mixed = __any_sync(0xffffffffu, mixed);
      if (lane == 0) {
        if (mixed) a.ctl->need_exact = 1;
        const Slot* sl = a.slots + (n_iter - 1);
        status[0] = (double)n_iter;
        status[1] = sl->best_nan ? __longlong_as_double(0x7ff8000000000000LL) : __longlong_as_double((long long)sl->best_max);
        status[2] = (double)a.ctl->q_fail;
        status[3] = (double)a.ctl->aqa_fail;
        status[4] = (double)launches;
        status[5] = (double)(a.ctl->need_exact | (mixed ? 1u : 0u));
        status[6] = a.ctl->kev_inv ? (double)(kKevBase - (int)a.ctl->kev_inv) : -1.0;
        status[7] = 0.0;
      }
    }
  }
  __syncthreads();
  const int n_iter = s_niter;
  const int n = a.n, m = a.m;
  // one warp per problem
  const int prob = blockIdx.x * (blockDim.x >> 5) + warp;
  if (prob >= a.nb) return;
  const int* ps = ra.pst + (size_t)prob * kPst;
  const int poison_it = ps[1] ? ps[1] - 2 : INT_MAX;
  int bi = -1;
  if (poison_it >= 0) {
    double best = 0.0;
    const int last = n_iter - 1 < poison_it ? n_iter - 1 : poison_it;
    for (int it = 0; it <= last; it++) {
      const double rd = ra.rec[((size_t)prob * a.max_iter + it) * 2];
      if (it == 0 || rd < best) { best = rd; bi = it; }
    }
  }
  double* bx = a.bx + (size_t)prob * n; double* bs = a.bs + (size_t)prob * m; double* bz = a.bz + (size_t)prob * m;
  if (bi < 0) {  // poisoned by the initial point: the reference's outputs are all NaN
    for (int c = lane; c < n; c += 32) bx[c] = t_nan<double>();
    for (int i = lane; i < m; i += 32) { bs[i] = t_nan<double>(); bz[i] = t_nan<double>(); }
  } else {
    const double* hh = ra.hist + ((size_t)prob * (a.max_iter + 1) + bi) * ra.hs;
    for (int c = lane; c < n; c += 32) bx[c] = hh[c];
    for (int i = lane; i < m; i += 32) { bs[i] = hh[round4(n) + i]; bz[i] = hh[round4(n) + round4(m) + i]; }
  }
}

}  // namespace b200qp
