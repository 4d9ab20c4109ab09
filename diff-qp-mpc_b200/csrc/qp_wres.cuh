// qp_wres.cuh -- ONE WARP PER QP on the resident route (fp64, neq == 0, nineq < 64, nz <= 32: the headline shape).
//
// Same algorithm, same launch protocol and same per-problem records as k_res_chunk (qp_resident.cuh; reference:
// qpth/solvers/pdipm/batch.py:46-214 forward / get_step, :351-374 solve_kkt, :434-469 factor_kkt), different
// machine mapping.  ncu on the 128-thread-CTA kernels showed ~25 k executed warp instructions per
// problem-iteration (for 117 kflop), 4 problems in flight per SM and 40 % of the stall samples on CTA barriers
// behind one-warp phases.  Here a problem never leaves its warp:
//   * T = R + diag(s/z) lives in REGISTERS as the 36 lower 8x8 DMMA accumulator tiles of the 64-padded matrix
//     (2 doubles per lane per tile).  Right-looking LDL^T by 8-column panels: the diagonal tile is factored with
//     quad shuffles and carries W = L_JJ^-1 along; the tiles below are X_IJ = C_IJ W^T by two DMMAs each (the
//     accumulator registers ARE the A operand when the contraction index is split into even / odd columns, and W
//     in accumulator layout IS the B operand); the trailing update C_IK -= X_IJ D^-1 X_KJ^T is two DMMAs per tile
//     with the accumulator registers of X_IJ as A and the scaled accumulator registers of X_KJ as B.  No layout
//     conversion, no shared-memory round trip, no barrier.
//   * the predictor's right-hand side rides through the factorisation as the bordered row m (as in qp_dmma.cuh);
//   * the triangular sweeps are tile mat-vecs straight from the factor registers (forward: quad reductions,
//     backward: reductions across quads), 8x8 diagonal blocks by their explicit inverse W;
//   * G sits in shared memory (leading dimension 34: conflict-free 16-byte loads by row and by column pair), Q and
//     Q^-1 are streamed from L2 as quad-per-row fragments, vectors live in registers in two layouts
//     (compact: elements lane, lane + 32; pair: elements 2l, 2l + 1 in lane l < 16) and in 5 KB of shared memory
//     wherever a mat-vec or a sweep needs a broadcast.
// 8 problems per SM in flight (register bound), ~5 k executed instructions per problem-iteration.
#pragma once
#include "qp_resident.cuh"

namespace b200qp {

constexpr int kWLd = 34;  // leading dimension of G in shared memory (doubles)
#ifndef B200QP_WRES_MINB
#define B200QP_WRES_MINB 8  // resident warps (= CTAs) per SM the chunk kernel is compiled for
#endif

struct WOff { int G, x, rx, t, qx, p, z, dz, dinv, hz, u, q, r, rinv, h, total; };
__host__ __device__ constexpr WOff wres_off(int m) {
  WOff o{};
  int p = 0;
  o.G = p; p += ((m + 1) & ~1) * kWLd;
  o.x = p; p += 32; o.rx = p; p += 32; o.t = p; p += 32; o.qx = p; p += 32; o.p = p; p += 32;
  o.z = p; p += 64; o.dz = p; p += 64; o.dinv = p; p += 64; o.hz = p; p += 64;
  o.u = p; p += 64; o.q = p; p += 64; o.r = p; p += 64; o.rinv = p; p += 64; o.h = p; p += 64;
  o.total = p;
  return o;
}

__device__ __forceinline__ void w_dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// 64-bit shuffles as two 32-bit ones, low word first (the toolkit's double overloads go through volatile asm moves and
// shuffle the high word first, which costs a register-pair fix-up per shuffle under register pressure)
__device__ __forceinline__ double wshfl(double v, int src) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(0xffffffffu, lo, src);
  hi = __shfl_sync(0xffffffffu, hi, src);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double wshfl_x(double v, int mask) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_xor_sync(0xffffffffu, lo, mask);
  hi = __shfl_xor_sync(0xffffffffu, hi, mask);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double quad_sum(double v) { v += wshfl_x(v, 1); v += wshfl_x(v, 2); return v; }
__device__ __forceinline__ double oct_sum(double v) { v += wshfl_x(v, 4); v += wshfl_x(v, 8); v += wshfl_x(v, 16); return v; }
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += wshfl_x(v, o);
  return v;
}

__host__ __device__ constexpr int w_tile(int I, int K) { return I * (I + 1) / 2 + K; }

// ------------------------------------------------------------------------------------------------------------
// Quad-per-row fragment of an (n x n), n <= 32, row-major matrix in global memory: lane (g, q) holds
// e[8 j + c] = M[8 j + g][8 q + c].  Loaded early (the latency hides behind whatever follows), applied later.
struct WMat { double e[32]; };
__device__ __forceinline__ void w_gload(WMat& M, const double* __restrict__ src, int ld, int n, int g, int q) {
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int r = 8 * j + g;
#pragma unroll
    for (int c = 0; c < 8; c++) {
      const int col = 8 * q + c;
      M.e[8 * j + c] = (r < n && col < n) ? __ldg(src + (size_t)r * ld + col) : 0.0;
    }
  }
}
// out[r] = sum_c M[r][c] v[c] for r < 32 (rows >= n give 0); v, out in shared memory (32 entries, zero padded)
__device__ __forceinline__ void w_gapply(const WMat& M, const double* v, double* out, int g, int q) {
  double vv[8];
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const double2 t = *reinterpret_cast<const double2*>(v + 8 * q + 2 * c);
    vv[2 * c] = t.x; vv[2 * c + 1] = t.y;
  }
#pragma unroll
  for (int j = 0; j < 4; j++) {
    double a0 = 0.0, a1 = 0.0;
#pragma unroll
    for (int c = 0; c < 8; c += 2) {
      a0 = fma(M.e[8 * j + c], vv[c], a0);
      a1 = fma(M.e[8 * j + c + 1], vv[c + 1], a1);
    }
    const double s = quad_sum(a0 + a1);
    if (q == (j & 3)) out[8 * j + g] = s;
  }
}

// y[r] = sum_c G[r][c] v[c] for the rows r = lane and lane + 32 (compact layout); G, v in shared memory
template <int RPL>
__device__ __forceinline__ void w_gv(const double* sG, const double* v, int n, int mr, int lane, double (&y)[RPL]) {
  const int np = (n + 1) >> 1;
#pragma unroll
  for (int s = 0; s < RPL; s++) {
    const int r = s * 32 + lane;
    double a0 = 0.0, a1 = 0.0;
    if (r < mr) {
      const double* row = sG + r * kWLd;
#pragma unroll
      for (int c = 0; c < 16; c++) {
        if (c < np) {
          const double2 gv = *reinterpret_cast<const double2*>(row + 2 * c);
          const double2 vv = *reinterpret_cast<const double2*>(v + 2 * c);
          a0 = fma(gv.x, vv.x, a0);
          a1 = fma(gv.y, vv.y, a1);
        }
      }
    }
    y[s] = a0 + a1;
  }
}
// (o0, o1) = columns 2l, 2l + 1 (l = lane & 15, pair layout, both half-warps get the sum) of G^T u; G, u in shared memory
__device__ __forceinline__ void w_gtu(const double* sG, const double* u, int mr, int lane, double& o0, double& o1) {
  const int l = lane & 15, h = lane >> 4;
  const double* gp = sG + h * kWLd + 2 * l;
  const double* up = u + h;
  double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    if (2 * j < mr) {
      const double2 gv = *reinterpret_cast<const double2*>(gp + 2 * j * kWLd);
      const double uu = up[2 * j];
      a0 = fma(gv.x, uu, a0); a1 = fma(gv.y, uu, a1);
    }
    if (2 * j + 2 < mr) {
      const double2 gv = *reinterpret_cast<const double2*>(gp + (2 * j + 2) * kWLd);
      const double uu = up[2 * j + 2];
      b0 = fma(gv.x, uu, b0); b1 = fma(gv.y, uu, b1);
    }
  }
  o0 = a0 + b0; o1 = a1 + b1;
  o0 += wshfl_x(o0, 16); o1 += wshfl_x(o1, 16);
}

// ------------------------------------------------------------------------------------------------------------
// Lane coordinates, computed once and made opaque to the compiler (which otherwise re-derives them from
// SR_TID.X -- a ~20-cycle special-register read -- at every use under register pressure).
struct WLane {
  int lane, g, q, q8, lim0, lim1;
  __device__ __forceinline__ void init() {
    lane = threadIdx.x & 31;
    asm volatile("" : "+r"(lane));
    g = lane >> 2; q = lane & 3; q8 = 8 * q;
    lim0 = (2 * q <= g) ? 2 * q : 0;          // slot s of a diagonal tile is updated at steps k < lim_s
    lim1 = (2 * q + 1 <= g) ? 2 * q + 1 : 0;
    asm volatile("" : "+r"(g), "+r"(q), "+r"(q8), "+r"(lim0), "+r"(lim1));
  }
};
// c += a * b where k < lim holds (one compare + one predicated DFMA instead of DFMA + two selects)
__device__ __forceinline__ void pfma_lt(double& c, double a, double b, int k, int lim) {
  asm("{\n .reg .pred pp;\n setp.lt.s32 pp, %3, %4;\n @pp fma.rn.f64 %0, %1, %2, %0;\n}" : "+d"(c) : "d"(a), "d"(b), "r"(k), "r"(lim));
}
// quad broadcast: the value of lane (lane & ~3) + src, src an immediate
__device__ __forceinline__ double wshfl_quad(double v, int src) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(0xffffffffu, lo, src, 4);
  hi = __shfl_sync(0xffffffffu, hi, src, 4);
  return __hiloint2double(hi, lo);
}

// LDL^T of one 8x8 diagonal tile in accumulator layout (lane (g, q): row g, columns 2q, 2q+1; lower triangle
// valid, the rest finite), carrying W = L^-1 along.  Straight-line code, ~30 instructions per column:
//   * column k is broadcast with the rows <= k masked to zero, so the multiplier of an eliminated row and the
//     "pivot-row" entry of an eliminated column are exact zeros and every update is unconditional;
//   * 1 / d_k = MUFU seed + one cubically convergent step r0 (1 + e + e^2), e = 1 - d r0 (>= 20 -> 60 bits).  A
//     pivot <= 0, NaN, denormal or >= 2^1022 gives r <= 0, NaN, inf or 0: the caller checks the eight stored
//     reciprocals once per tile (such a T is the reference's poisoned iterate).
//   c0, c1 in/out: on exit column k < g of row g holds X[g][k] = L[g][k] D_k
//   w0, w1 out   : W = L^-1 (unit lower triangular), accumulator layout
//   rinv out     : shared memory, the 8 reciprocal pivots (1 for bordered / padding columns, base + k >= m)
template <int MC>
__device__ __forceinline__ void w_diag(double& c0, double& c1, double& w0, double& w1, double* rinv, int base, int m_rt,
                                       const WLane& L) {
  const int m = MC > 0 ? MC : m_rt;
  w0 = (L.g == 2 * L.q) ? 1.0 : 0.0;
  w1 = (L.g == 2 * L.q + 1) ? 1.0 : 0.0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const int kq = k >> 1;
    const double ck = (k & 1) ? c1 : c0;
    const double dk = wshfl(ck, 4 * k + kq);         // C[k][k]
    const double ckm = (L.g > k) ? ck : 0.0;         // rows that are still being eliminated
    const double cik = wshfl_quad(ckm, kq);          // C[g][k]     (0 for g <= k)
    const double cj0 = wshfl(ckm, L.q8 + kq);        // C[2q][k]    (0 for 2q <= k)
    const double cj1 = wshfl(ckm, L.q8 + 4 + kq);    // C[2q+1][k]
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(dk));
    const double e = fma(-dk, r, 1.0);
    const double t = fma(e, e, e);
    r = fma(r, t, r);
    if (base + k >= m) r = 1.0;  // folds for compile-time sizes
    if (L.lane == 0) rinv[k] = r;
    const double nl = -(cik * r);
    c0 = fma(nl, cj0, c0);   // entries (g, j) with j > g are never read unmasked
    c1 = fma(nl, cj1, c1);
    if (k < 7) {
      const double wk0 = wshfl(w0, L.q + 4 * k), wk1 = wshfl(w1, L.q + 4 * k);  // row k of W (final after step k-1)
      w0 = fma(nl, wk0, w0);
      w1 = fma(nl, wk1, w1);
    }
  }
}

// T = R + diag(dinv), bordered by the row hz at index m, padded by the identity: fix the tiles that are not interior
template <int NTI, int MC>
__device__ __forceinline__ void w_fix_tiles(double (&C)[NTI * (NTI + 1) / 2][2], const double* sdinv, const double* shz, int m_rt,
                                            int g, int q) {
  const int m = MC > 0 ? MC : m_rt;
#pragma unroll
  for (int I = 0; I < NTI; I++) {
#pragma unroll
    for (int K = 0; K <= I; K++) {
      if (!(I != K && 8 * I + 8 <= m)) {  // uniform; folds for compile-time sizes
        double& c0 = C[w_tile(I, K)][0];
        double& c1 = C[w_tile(I, K)][1];
        const int i = 8 * I + g, k = 8 * K + 2 * q;
        double v0, v1;
        if (i < m) {
          v0 = (k <= i && k < m) ? c0 : 0.0;
          v1 = (k + 1 <= i && k + 1 < m) ? c1 : 0.0;
          if (I == K) {
            const double di = sdinv[i];
            if (k == i) v0 += di;
            if (k + 1 == i) v1 += di;
          }
        } else if (i == m) {
          const double2 hv = *reinterpret_cast<const double2*>(shz + k);  // zero beyond m
          v0 = (k == m) ? 1.0 : hv.x;
          v1 = (k + 1 == m) ? 1.0 : hv.y;
        } else {
          v0 = (k == i) ? 1.0 : 0.0;
          v1 = (k + 1 == i) ? 1.0 : 0.0;
        }
        c0 = v0; c1 = v1;
      }
    }
  }
}

// Same for tiles written by the tensor-core path of k_prefactor (qp_kernels.cuh; fp64, no equalities, all
// (m + 7) / 8 == NTI tile rows present): padding entries are exact zeros and the strict upper triangle of a diagonal
// tile holds the mirrored values, which the factorisation never reads.  Left to do: the diagonal (1/d, or 1 for the
// bordered / padding rows) and the bordered row.
template <int NTI, int MC>
__device__ __forceinline__ void w_fix_tiles_clean(double (&C)[NTI * (NTI + 1) / 2][2], const double* sdinv, const double* shz,
                                                  int g, int q) {
  static_assert(MC > 0 && (MC + 8) / 8 == NTI, "compile-time nineq filling all tile rows");
  constexpr int Ib = MC >> 3, gb = MC & 7;
#pragma unroll
  for (int I = 0; I < NTI; I++) {
    const int i = 8 * I + g;
    const double di = (i < MC) ? sdinv[i] : 1.0;
    if (2 * q == g) C[w_tile(I, I)][0] += di;
    if (2 * q + 1 == g) C[w_tile(I, I)][1] += di;
  }
#pragma unroll
  for (int K = 0; K <= Ib; K++) {
    const double2 hv = *reinterpret_cast<const double2*>(shz + 8 * K + 2 * q);  // zero beyond m
    if (g == gb) {
      const int k = 8 * K + 2 * q;
      C[w_tile(Ib, K)][0] = (k == MC) ? 1.0 : hv.x;
      C[w_tile(Ib, K)][1] = (k + 1 == MC) ? 1.0 : hv.y;
    }
  }
}

// Blocked LDL^T of the register tiles (file header).  On exit: tiles (I, K < I) hold X = L D, diagonal tiles hold
// W = L_II^-1, srinv[0..MPAD) the reciprocal pivots, su[k] = (D^-1 L^-1 hz)_k for k < m (0 beyond): the forward
// substitution of the bordered right-hand side.  Returns false when some pivot was not a positive normal number
// below 2^1022 (see w_diag).
template <int NTI, int MC>
__device__ __forceinline__ bool w_factor(double (&C)[NTI * (NTI + 1) / 2][2], double* srinv, double* su, int m_rt, const WLane& L) {
  bool good = true;
  const int m = MC > 0 ? MC : m_rt;
  const int Ib = m >> 3, gb = m & 7;
  const int q = L.q;
#pragma unroll
  for (int J = 0; J < NTI; J++) {
    if (J <= Ib) {  // uniform; tile rows beyond the bordered one are the identity
      double w0, w1;
      double& d0 = C[w_tile(J, J)][0];
      double& d1 = C[w_tile(J, J)][1];
      w_diag<MC>(d0, d1, w0, w1, srinv + 8 * J, 8 * J, m, L);
      __syncwarp();
      const double2 sc = *reinterpret_cast<const double2*>(srinv + 8 * J + 2 * q);
      good = good && (sc.x > 0.0) && (sc.x < t_inf<double>()) && (sc.y > 0.0) && (sc.y < t_inf<double>());
      if (J == Ib && L.g == gb)
        *reinterpret_cast<double2*>(su + 8 * J + 2 * q) = make_double2(2 * q < gb ? d0 * sc.x : 0.0, 2 * q + 1 < gb ? d1 * sc.y : 0.0);
      d0 = w0; d1 = w1;
      double b0[NTI], b1[NTI];
#pragma unroll
      for (int I = J + 1; I < NTI; I++) {
        if (I <= Ib) {
          double x0 = 0.0, x1 = 0.0;
          w_dmma(x0, x1, C[w_tile(I, J)][0], w0);
          w_dmma(x0, x1, C[w_tile(I, J)][1], w1);
          C[w_tile(I, J)][0] = x0; C[w_tile(I, J)][1] = x1;
          if (I == Ib && L.g == gb) *reinterpret_cast<double2*>(su + 8 * J + 2 * q) = make_double2(x0 * sc.x, x1 * sc.y);
          b0[I] = -sc.x * x0; b1[I] = -sc.y * x1;
        }
      }
#pragma unroll
      for (int I = J + 1; I < NTI; I++) {
        if (I <= Ib) {
#pragma unroll
          for (int K = J + 1; K <= I; K++) {
            w_dmma(C[w_tile(I, K)][0], C[w_tile(I, K)][1], C[w_tile(I, J)][0], b0[K]);
            w_dmma(C[w_tile(I, K)][0], C[w_tile(I, K)][1], C[w_tile(I, J)][1], b1[K]);
          }
        }
      }
    }
  }
  return __all_sync(0xffffffffu, good);
}

// su <- D^-1 L^-1 sr (natural order in shared memory; entries >= m come out 0)
template <int NTI, int MC>
__device__ __forceinline__ void w_fwd(const double (&C)[NTI * (NTI + 1) / 2][2], const double* sr, const double* srinv, double* su,
                                      int m_rt, int lane, int g, int q) {
  const int m = MC > 0 ? MC : m_rt;
  const int Ib = m >> 3;
#pragma unroll
  for (int I = 0; I < NTI; I++) {
    if (I <= Ib) {
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int K = 0; K < I; K++) {
        const double2 uv = *reinterpret_cast<const double2*>(su + 8 * K + 2 * q);
        a0 = fma(C[w_tile(I, K)][0], uv.x, a0);
        a1 = fma(C[w_tile(I, K)][1], uv.y, a1);
      }
      const double v = sr[8 * I + g] - quad_sum(a0 + a1);  // row layout
      const double v0 = wshfl(v, 8 * q), v1 = wshfl(v, 8 * q + 4);  // v[2q], v[2q+1]
      const double y = quad_sum(fma(C[w_tile(I, I)][0], v0, C[w_tile(I, I)][1] * v1));
      const double u = (8 * I + g < m) ? y * srinv[8 * I + g] : 0.0;
      if (q == 0) su[8 * I + g] = u;
      __syncwarp();
    }
  }
}
// sq <- L^-T su
template <int NTI, int MC>
__device__ __forceinline__ void w_bwd(const double (&C)[NTI * (NTI + 1) / 2][2], const double* su, const double* srinv, double* sq,
                                      int m_rt, int lane, int g, int q) {
  const int m = MC > 0 ? MC : m_rt;
  const int Ib = m >> 3;
#pragma unroll
  for (int K = NTI - 1; K >= 0; K--) {
    if (K <= Ib) {
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int I = K + 1; I < NTI; I++) {
        if (I <= Ib) {
          const double qi = sq[8 * I + g];
          a0 = fma(C[w_tile(I, K)][0], qi, a0);
          a1 = fma(C[w_tile(I, K)][1], qi, a1);
        }
      }
      a0 = oct_sum(a0); a1 = oct_sum(a1);  // column layout, every lane
      const double2 rv = *reinterpret_cast<const double2*>(srinv + 8 * K + 2 * q);
      const double2 uv = *reinterpret_cast<const double2*>(su + 8 * K + 2 * q);
      const double v0 = fma(-rv.x, a0, uv.x), v1 = fma(-rv.y, a1, uv.y);
      const double s0 = wshfl(v0, g >> 1), s1 = wshfl(v1, g >> 1);
      const double vg = (g & 1) ? s1 : s0;  // v[g]
      const double p0 = oct_sum(C[w_tile(K, K)][0] * vg), p1 = oct_sum(C[w_tile(K, K)][1] * vg);
      if (lane < 4) *reinterpret_cast<double2*>(sq + 8 * K + 2 * q) = make_double2(p0, p1);
      __syncwarp();
    }
  }
}

// get_step pieces (batch.py:211-214) of (z, dz) and (s, ds) in the compact layout: the NaN-propagating minimum of
// a = -v / dv over the entries the fill does not overwrite (dv > 0 is overwritten; +inf if none is left), whether some
// entry is overwritten (has bit 0 / 1), and whether a holds a NaN anywhere (nanz / nans: the batch maximum that defines
// the fill is then NaN).  NaNs are tracked as flags and combined by votes: the reductions are plain minima.
template <int RPL>
__device__ __forceinline__ void w_pieces(const double (&z)[RPL], const double (&dz)[RPL], const double (&s)[RPL],
                                         const double (&ds)[RPL], int m, int lane, double& rz, double& rs, int& has, bool& nanz,
                                         bool& nans) {
  double mz = t_inf<double>(), ms = t_inf<double>();
  bool hz = false, hs = false, nrz = false, nrs = false, naz = false, nas = false;
#pragma unroll
  for (int q = 0; q < RPL; q++) {
    if (q * 32 + lane < m) {
      const double a1 = -z[q] / dz[q], a2 = -s[q] / ds[q];
      const bool n1 = is_nan(a1), n2 = is_nan(a2);
      naz |= n1; nas |= n2;
      if (dz[q] > 0.0) hz = true; else { nrz |= n1; mz = fmin(mz, a1); }
      if (ds[q] > 0.0) hs = true; else { nrs |= n2; ms = fmin(ms, a2); }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mz = fmin(mz, wshfl_x(mz, o));
    ms = fmin(ms, wshfl_x(ms, o));
  }
  has = (__any_sync(0xffffffffu, hz) ? 1 : 0) | (__any_sync(0xffffffffu, hs) ? 2 : 0);
  rz = __any_sync(0xffffffffu, nrz) ? t_nan<double>() : mz;
  rs = __any_sync(0xffffffffu, nrs) ? t_nan<double>() : ms;
  nanz = __any_sync(0xffffffffu, naz);
  nans = __any_sync(0xffffffffu, nas);
}

// ------------------------------------------------------------------------------------------------------------
// One launch = the iterations [.., ra.it_end) of every problem, one warp per problem, WPC independent warps
// (problems) per CTA.  With WPC > 1 the warps of a CTA re-align at the top of every iteration (one CTA barrier):
// they execute the same ~100 KB of straight-line code, and in step they share its instruction-cache lines.
// NTI = tiles per dimension (8 NTI > nineq); <NC, MC> != 0: compile-time nz / nineq.
template <int NTI, int NC, int MC, int WPC>
__global__ void __launch_bounds__(32 * WPC, WPC == 1 ? B200QP_WRES_MINB : 8 / WPC) k_wres_chunk(const KArgs<double> a, const RArgs ra) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NT = NTI * (NTI + 1) / 2, MPAD = 8 * NTI, RPL = (MPAD + 31) / 32;
  WLane L;
  L.init();
  const int lane = L.lane, g = L.g, q = L.q, l16 = lane & 15;
  const int prob = blockIdx.x * WPC + (WPC > 1 ? (int)(threadIdx.x >> 5) : 0);
  if (prob >= a.nb) return;
  const int n = NC > 0 ? NC : a.n, m = MC > 0 ? MC : a.m;
  const int n4 = res_r4(n), m4 = res_r4(m), hs = n4 + 2 * m4;
  const int mr = (m + 1) & ~1;

  // ---- where does this problem (re)start?  (protocol of qp_resident.cuh)
  int* ps = ra.pst + (size_t)prob * kPst;
  int stage = ps[0], poison = ps[1];
  unsigned used = (unsigned)ps[2], sens = (unsigned)ps[3], fo = (unsigned)ps[4], azm = (unsigned)ps[5], asm_ = (unsigned)ps[6];
  const unsigned kevi = *(volatile unsigned*)&a.ctl->kev_inv;
  const int Kstart = kevi ? kKevBase - (int)kevi : INT_MAX;
  int it = stage - 1;
  if (stage > 1 && Kstart < it) {
    const unsigned bad = sens & ~used & ~((1u << Kstart) - 1u) & ((1u << it) - 1u);
    if (bad) {
      it = __ffs(bad) - 1;
      const unsigned keep = (1u << it) - 1u;
      used &= keep; sens &= keep; fo &= keep; azm &= keep; asm_ &= keep;
      poison = 0;
    }
  }
  // Once the first NaN event K* of the batch is known (and behind us) nothing is speculated any more: the fill is 1.0
  // in every remaining iteration, so this launch runs the problem to the end and the later launches find it done.
  const int it_end = (Kstart <= it) ? a.max_iter : ra.it_end;
  if (poison || it >= it_end) return;  // uniform

  const WOff o = wres_off(m);
  double* sm = reinterpret_cast<double*>(smem_raw) + (WPC > 1 ? (threadIdx.x >> 5) * o.total : 0);
  double* sG = sm + o.G;
  double* sx = sm + o.x; double* srx = sm + o.rx; double* st = sm + o.t; double* sqx = sm + o.qx;
  double* sz = sm + o.z; double* sdz = sm + o.dz; double* sdinv = sm + o.dinv; double* shz = sm + o.hz;
  double* su = sm + o.u; double* sq = sm + o.q; double* sr = sm + o.r; double* srinv = sm + o.rinv;
  double* sp = sm + o.p; double* sh = sm + o.h;
  double* hist = ra.hist + (size_t)prob * (a.max_iter + 1) * hs;

  // ---- stage G (zero padded to 32 columns / an even number of rows), zero the vectors
  {
    const double* Gg = a.G + (size_t)prob * a.sG;
    for (int r = 0; r < mr; r++) {
      if (r < m && lane < n) cp_async8(sG + r * kWLd + lane, Gg + (size_t)r * n + lane);
      else sG[r * kWLd + lane] = 0.0;
    }
    cp_async_commit();
    for (int i = lane; i < o.total - o.x; i += 32) sm[o.x + i] = 0.0;
  }
  const double* Qg = a.Q + (size_t)prob * a.sQ;
  const double* Qig = a.Qi + (size_t)prob * a.sQi;
  const double* Rf = a.R + (size_t)prob * a.sR;
  const int ldqi = a.ldn;
  // p, h and x live in shared memory (they are constants or rarely touched: registers are the scarce resource), the
  // iterate s, z in registers (compact layout)
  double zr[RPL], sr_[RPL];
  {
    const double* pg = a.pv + (size_t)prob * a.sp;
    const double* hg = a.h + (size_t)prob * a.sh;
    const double* hh = hist + (size_t)(it < 0 ? 0 : it) * hs;
    __syncwarp();
    if (lane < n) { sp[lane] = pg[lane]; if (it >= 0) sx[lane] = hh[lane]; }
#pragma unroll
    for (int s = 0; s < RPL; s++) {
      const int i = s * 32 + lane;
      if (i < m) sh[i] = hg[i];
      zr[s] = (i < m && it >= 0) ? hh[n4 + m4 + i] : 0.0;
      sr_[s] = (i < m && it >= 0) ? hh[n4 + i] : 0.0;
    }
  }
  cp_async_wait_all();
  __syncwarp();

  bool alive = true;
  bool have_hist = true;  // the history entry of the first iteration of this launch already exists
  double C[NT][2];
  // Register-heavy loads (Q, Q^-1 as quad-per-row fragments: 64 registers each; the 36 tiles of R: 144) are issued one
  // phase ahead of their use, at points where the registers they land in are dead:
  //   Q^-1 #1 (for t = Q^-1 rx)  with Q at the top of the iteration          (the tiles are not loaded yet)
  //   R tiles                    after the Q product                        (Q's registers are free)
  //   Q^-1 #2 (for dx)           after the last sweep                       (the factor is dead)
#pragma unroll 1
  for (; it < it_end; ++it) {
    if (WPC > 1) __syncthreads();  // exited warps do not take part
    const bool init = it < 0;
    double dr[RPL], rz[RPL], mu = 0.0, t4 = 0.0;
    WMat MQi;
    w_gload(MQi, Qig, ldqi, n, g, q);
    if (!init) {
      WMat MQ;
      w_gload(MQ, Qg, n, n, g, q);
      if (!have_hist) {
        double* hh = hist + (size_t)it * hs;
        if (lane < n) hh[lane] = sx[lane];
#pragma unroll
        for (int s = 0; s < RPL; s++) {
          const int i = s * 32 + lane;
          if (i < m) { hh[n4 + i] = sr_[s]; hh[n4 + m4 + i] = zr[s]; }
        }
      }
      // ---------------- residuals (batch.py:93-108)
#pragma unroll
      for (int s = 0; s < RPL; s++) {
        const int i = s * 32 + lane;
        dr[s] = zr[s] / sr_[s];
        if (i < m) { sz[i] = zr[s]; sdinv[i] = 1.0 / dr[s]; } else dr[s] = 1.0;
      }
      __syncwarp();
      double gx[RPL];
      w_gv<RPL>(sG, sx, n, mr, lane, gx);
#pragma unroll
      for (int s = 0; s < RPL; s++) { const int i = s * 32 + lane; rz[s] = gx[s] + sr_[s] - (i < m ? sh[i] : 0.0); }
      w_gapply(MQ, sx, sqx, g, q);
    } else {
      // initial point (batch.py:60-66): d = 1, (rx, rs, rz) = (p, 0, -h)
#pragma unroll
      for (int s = 0; s < RPL; s++) {
        const int i = s * 32 + lane;
        dr[s] = 1.0; zr[s] = 0.0; sr_[s] = 0.0; rz[s] = i < m ? -sh[i] : 0.0;
        if (i < m) sdinv[i] = 1.0;
      }
    }
    // the tiles of T travel while the right-hand side is formed
#pragma unroll
    for (int t = 0; t < NT; t++) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(Rf + (size_t)t * 64 + lane * 2));
      C[t][0] = v.x; C[t][1] = v.y;
    }
    if (!init) {
      double gz0, gz1;
      w_gtu(sG, sz, mr, lane, gz0, gz1);
      __syncwarp();
      double rx0, rx1;
      {
        const double2 qx = *reinterpret_cast<const double2*>(sqx + 2 * l16);
        const double2 pv = *reinterpret_cast<const double2*>(sp + 2 * l16);
        rx0 = gz0 + (qx.x + pv.x);
        rx1 = gz1 + (qx.y + pv.y);
      }
      if (lane < 16) *reinterpret_cast<double2*>(srx + 2 * lane) = make_double2(rx0, rx1);
      // residual norms and mu (batch.py:103-108)
      {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        if (lane < 16) s0 = fma(rx0, rx0, rx1 * rx1);
#pragma unroll
        for (int s = 0; s < RPL; s++) {
          if (s * 32 + lane < m) { s1 = fma(rz[s], rz[s], s1); s2 = fma(sr_[s], zr[s], s2); }
        }
        s0 = wsum(s0); s1 = wsum(s1); s2 = wsum(s2);
        mu = fabs(s2 / (double)m);
        t4 = s2;
        const double resid = sqrt(s1) + sqrt(s0) + (double)m * mu;
        if (lane == 0) {
          double* rc = ra.rec + ((size_t)prob * a.max_iter + it) * 2;
          rc[0] = resid; rc[1] = mu;
        }
      }
    } else {
      if (lane < 16) *reinterpret_cast<double2*>(srx + 2 * lane) = *reinterpret_cast<const double2*>(sp + 2 * lane);
    }
    have_hist = false;
    // ---------------- right-hand side of the reduced system: hz = G Q^-1 rx + rs / d - rz   (rs = z)
    __syncwarp();
    w_gapply(MQi, srx, st, g, q);
    __syncwarp();
    {
      double gt[RPL];
      w_gv<RPL>(sG, st, n, mr, lane, gt);
#pragma unroll
      for (int s = 0; s < RPL; s++) {
        const int i = s * 32 + lane;
        if (i < m) shz[i] = zr[s] / dr[s] + (gt[s] - rz[s]);
      }
    }
    __syncwarp();
    // ---------------- T = R + diag(1/d) = L D L^T with hz riding as the bordered row
    if constexpr (MC > 0 && (MC + 8) / 8 == NTI) w_fix_tiles_clean<NTI, MC>(C, sdinv, shz, g, q);
    else w_fix_tiles<NTI, MC>(C, sdinv, shz, m, g, q);
    const bool bad = !w_factor<NTI, MC>(C, srinv, su, m, L);
    if (bad) {
      // non-positive / NaN pivot: this problem can never improve again; its ratios count as NaN from here on
      if (lane == 0) {
        ps[1] = it + 2;
        atomicMax(&a.ctl->kev_inv, (unsigned)(kKevBase - (it < 0 ? 0 : it)));
      }
      alive = false;
      break;
    }
    __syncwarp();
    // ---------------- predictor: qa = T^-1 hz
    w_bwd<NTI, MC>(C, su, srinv, sq, m, lane, g, q);
    double qa[RPL];
#pragma unroll
    for (int s = 0; s < RPL; s++) { const int i = s * 32 + lane; qa[s] = i < m ? sq[i] : 0.0; }
    double alpha = 1.0;
    if (init) {
      w_gload(MQi, Qig, ldqi, n, g, q);
      // x, s, z of the initial point; shift s and z so that their minima are >= 1 (batch.py:76-86)
      double mn_s = t_inf<double>(), mn_z = t_inf<double>();
#pragma unroll
      for (int s = 0; s < RPL; s++) {
        if (s * 32 + lane < m) { mn_s = nanmin(mn_s, qa[s]); mn_z = nanmin(mn_z, -qa[s]); }
      }
#pragma unroll
      for (int o2 = 16; o2 > 0; o2 >>= 1) { mn_s = nanmin(mn_s, wshfl_x(mn_s, o2)); mn_z = nanmin(mn_z, wshfl_x(mn_z, o2)); }
#pragma unroll
      for (int s = 0; s < RPL; s++) {
        const int i = s * 32 + lane;
        double sv = qa[s], zv = -qa[s];
        if (mn_s < 0.0) sv -= mn_s - 1.0;
        if (mn_z < 0.0) zv -= mn_z - 1.0;
        if (i < m) { sdz[i] = -qa[s]; sr_[s] = sv; zr[s] = zv; }
      }
    } else {
      double dza[RPL], dsa[RPL];
#pragma unroll
      for (int s = 0; s < RPL; s++) { dza[s] = -qa[s]; dsa[s] = (-zr[s] - dza[s]) / dr[s]; }
      double pz, psl; int has; bool nz_, ns_;
      double zq[RPL], sq_[RPL];
#pragma unroll
      for (int s = 0; s < RPL; s++) { const bool v = s * 32 + lane < m; zq[s] = v ? zr[s] : 1.0; sq_[s] = v ? sr_[s] : 1.0; }
      w_pieces<RPL>(zq, dza, sq_, dsa, m, lane, pz, psl, has, nz_, ns_);
      // the clamp at 1 makes alpha_aff independent of the batch-global fill (batch.py:161-163)
      const double stz = (has & 1) ? nanmin(pz, 1.0) : pz;
      const double sts = (has & 2) ? nanmin(psl, 1.0) : psl;
      const double alpha_aff = nanmin(nanmin(stz, sts), 1.0);
      double t3 = 0.0;
#pragma unroll
      for (int s = 0; s < RPL; s++)
        if (s * 32 + lane < m) t3 += (sr_[s] + alpha_aff * dsa[s]) * (zr[s] + alpha_aff * dza[s]);
      t3 = wsum(t3);
      const double ratio = t3 / t4;
      const double sig = ratio * ratio * ratio;
      double rsc[RPL];
#pragma unroll
      for (int s = 0; s < RPL; s++) {
        const int i = s * 32 + lane;
        rsc[s] = (-mu * sig + dsa[s] * dza[s]) / sq_[s];
        if (i < m) sr[i] = rsc[s] / dr[s];
      }
      __syncwarp();
      // ---------------- corrector: qc = T^-1 (rsc / d)
      w_fwd<NTI, MC>(C, sr, srinv, su, m, lane, g, q);
      w_bwd<NTI, MC>(C, su, srinv, sq, m, lane, g, q);
      w_gload(MQi, Qig, ldqi, n, g, q);  // the factor is dead: Q^-1 for dx travels behind the step-length logic
      double dz[RPL], ds[RPL];
#pragma unroll
      for (int s = 0; s < RPL; s++) {
        const int i = s * 32 + lane;
        const double qc = i < m ? sq[i] : 0.0;
        const double dzc = -qc;
        const double dsc = (-rsc[s] - dzc) / dr[s];
        dz[s] = dza[s] + dzc; ds[s] = dsa[s] + dsc;
      }
      w_pieces<RPL>(zq, dz, sq_, ds, m, lane, pz, psl, has, nz_, ns_);
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
      Kdyn = __shfl_sync(0xffffffffu, Kdyn, 0);  // one view per warp
      const bool F = it >= Kdyn;
      // F = 0: fill = max(1, a.max()) >= every unfilled ratio -> the unfilled minimum (+inf when fill-only:
      //        0.999 * fill >= 1 is validated by k_res_reduce);  F = 1: fill = 1.0 exactly
      const double a0_ = nanmin(0.999 * nanmin(pz, psl), 1.0);
      const double z1 = (has & 1) ? nanmin(pz, 1.0) : pz, s1 = (has & 2) ? nanmin(psl, 1.0) : psl;
      const double a1_ = nanmin(0.999 * nanmin(z1, s1), 1.0);
      const bool same = (a0_ == a1_) || (is_nan(a0_) && is_nan(a1_));
      if (!same) sens |= 1u << it;
      if (F) used |= 1u << it;
      else if (((has & 1) && pz == t_inf<double>()) || ((has & 2) && psl == t_inf<double>())) fo |= 1u << it;
      alpha = F ? a1_ : a0_;
#pragma unroll
      for (int s = 0; s < RPL; s++) {
        const int i = s * 32 + lane;
        if (i < m) { sdz[i] = dz[s]; sr_[s] = sr_[s] + alpha * ds[s]; zr[s] = zr[s] + alpha * dz[s]; }
      }
    }
    __syncwarp();
    // ---------------- dx = Q^-1 (-rx - G^T dz);  x += alpha dx   (initial point: x = dx)
    {
      double gd0, gd1;
      w_gtu(sG, sdz, mr, lane, gd0, gd1);
      if (lane < 16) {
        const double2 rx = *reinterpret_cast<const double2*>(srx + 2 * lane);
        *reinterpret_cast<double2*>(st + 2 * lane) = make_double2(-rx.x - gd0, -rx.y - gd1);
      }
      __syncwarp();
      w_gapply(MQi, st, sqx, g, q);
      __syncwarp();
      if (lane < 16) {
        const double2 dx = *reinterpret_cast<const double2*>(sqx + 2 * lane);
        double2 xv = *reinterpret_cast<const double2*>(sx + 2 * lane);
        xv.x = init ? dx.x : xv.x + alpha * dx.x;
        xv.y = init ? dx.y : xv.y + alpha * dx.y;
        *reinterpret_cast<double2*>(sx + 2 * lane) = xv;
      }
      __syncwarp();
    }
  }
  // ---- hand the state to the next launch
  if (alive) {
    double* hh = hist + (size_t)it * hs;
    if (lane < n) hh[lane] = sx[lane];
#pragma unroll
    for (int s = 0; s < RPL; s++) {
      const int i = s * 32 + lane;
      if (i < m) { hh[n4 + i] = sr_[s]; hh[n4 + m4 + i] = zr[s]; }
    }
  }
  if (lane == 0) {
    ps[0] = (alive ? it : it + 1) + 1;
    if (alive) ps[1] = 0;
    ps[2] = (int)used; ps[3] = (int)sens; ps[4] = (int)fo; ps[5] = (int)azm; ps[6] = (int)asm_;
  }
}


// ------------------------------------------------------------------------------------------------------------
// pre_factor_kkt (qpth/solvers/pdipm/batch.py:377-428) with one warp per QP, everything on the FP64 tensor cores and
// in registers (fp64, neq == 0, nz <= 32, nineq < 64).  The 128-thread-CTA k_prefactor spends 3.2 ms on 32768
// problems of the headline shape in ~90 CTA barriers (column-by-column LDL^T of Q and two triangular sweeps over the
// identity); its arithmetic is 0.25 ms of the FP64 pipe.  Here:
//   * Q (10 lower 8x8 tiles, accumulator layout) -> w_factor: X = L D below the diagonal, W = L_JJ^-1 on it;
//   * U = L^-T by block back-substitution written so that every tile product has the form A B^T, the one a DMMA
//     computes from two accumulator-layout operands without any data movement (w_factor's trick):
//         U_JJ = W_J^T,   U_JI = -(sum_{J <= K < I} U_JK L_IK^T) W_I^T   (J < I)
//   * Q^-1 = U D^-1 U^T:   Qi_IJ = sum_{K >= I} (U_IK D_K^-1) U_JK^T  (I >= J), the upper tiles by transposition;
//   * per tile row I of G (loaded from global memory straight into accumulator layout):
//         (G Q^-1)_IJ = sum_K G_IK Qi_JK^T,    R_IK = sum_J (G Q^-1)_IJ G_KJ^T   (K <= I)
//     R leaves the accumulators in fragment order (one 16-byte store per lane and tile).
// 648 DMMAs per problem at the headline shape, no shared-memory traffic beyond the 64 reciprocal pivots, no barrier.
__device__ __forceinline__ void w_transpose(double c0, double c1, double& t0, double& t1, const WLane& L) {
  const int s0 = L.q8 + (L.g >> 1);
  const double a0 = wshfl(c0, s0), a1 = wshfl(c1, s0), b0 = wshfl(c0, s0 + 4), b1 = wshfl(c1, s0 + 4);
  t0 = (L.g & 1) ? a1 : a0;
  t1 = (L.g & 1) ? b1 : b0;
}
// C = A B^T for accumulator-layout tiles (a0, a1), (b0, b1), accumulated into (c0, c1)
__device__ __forceinline__ void w_abt(double& c0, double& c1, double a0, double a1, double b0, double b1) {
  w_dmma(c0, c1, a0, b0);
  w_dmma(c0, c1, a1, b1);
}

template <int NTI, int NC, int MC>
__global__ void __launch_bounds__(32, 8) k_wres_prefactor(const KArgs<double> a) {
  __shared__ __align__(16) double s_rinv[32];
  __shared__ __align__(16) double s_u[32];
  WLane L;
  L.init();
  const int lane = L.lane, g = L.g, q = L.q;
  const int prob = blockIdx.x + a.prob0;
  const int n = NC > 0 ? NC : a.n, m = MC > 0 ? MC : a.m;
  const int MT = (m + 7) >> 3;  // tile rows of G / R
  const double* Qg = a.Q + (size_t)prob * a.sQ;
  const double* Gg = a.G + (size_t)prob * a.sG;
  const bool vec = ((n & 1) == 0) && ((a.sQ & 1) == 0) && ((a.sG & 1) == 0) && ((((size_t)a.Q | (size_t)a.G) & 15) == 0);
  // one accumulator-layout tile of a row-major matrix with `rows` x n valid entries; `diag1`: identity padding
  auto load_tile = [&](const double* M, int rows, int I, int K, bool diag1, double& c0, double& c1) {
    const int i = 8 * I + g, k = 8 * K + 2 * q;
    c0 = (diag1 && i == k && i >= n) ? 1.0 : 0.0;
    c1 = (diag1 && i == k + 1 && i >= n) ? 1.0 : 0.0;
    if (i < rows) {
      if (vec) {
        if (k < n) { const double2 v = __ldg(reinterpret_cast<const double2*>(M + (size_t)i * n + k)); c0 = v.x; c1 = v.y; }
      } else {
        if (k < n) c0 = __ldg(M + (size_t)i * n + k);
        if (k + 1 < n) c1 = __ldg(M + (size_t)i * n + k + 1);
      }
    }
  };
  // ---- G travels first (64 registers); Q is needed at once
  double Gt[NTI][4][2];
#pragma unroll
  for (int I = 0; I < NTI; I++)
#pragma unroll
    for (int K = 0; K < 4; K++) {
      Gt[I][K][0] = 0.0; Gt[I][K][1] = 0.0;
      if (I < MT) load_tile(Gg, m, I, K, false, Gt[I][K][0], Gt[I][K][1]);
    }
  double C[10][2];
#pragma unroll
  for (int I = 0; I < 4; I++)
#pragma unroll
    for (int K = 0; K <= I; K++) {
      load_tile(Qg, n, I, K, true, C[w_tile(I, K)][0], C[w_tile(I, K)][1]);
      if (I == K && a.reg != 0.0) {
        const int i = 8 * I + g;
        if (i < n && i == 8 * K + 2 * q) C[w_tile(I, K)][0] += a.reg;
        if (i < n && i == 8 * K + 2 * q + 1) C[w_tile(I, K)][1] += a.reg;
      }
    }
  // ---- Q = L D L^T  (the "bordered row" of w_factor is row n of the identity padding here, or absent when n == 32)
  s_rinv[lane] = 1.0;  // tile rows that are all padding are skipped by w_factor
  __syncwarp();
  const bool okQ = w_factor<4, NC>(C, s_rinv, s_u, n, L);
  if (!okQ && lane == 0) atomicAdd(&a.ctl->q_fail, 1u);
  __syncwarp();
  double rv[4][2];  // reciprocal pivots of this lane's columns
#pragma unroll
  for (int K = 0; K < 4; K++) {
    const double2 r = *reinterpret_cast<const double2*>(s_rinv + 8 * K + 2 * q);
    rv[K][0] = r.x; rv[K][1] = r.y;
  }
  // ---- U = L^-T (upper tiles U[J][I], J <= I, stored at w_tile(I, J))
  double U[10][2];
#pragma unroll
  for (int J = 0; J < 4; J++) w_transpose(C[w_tile(J, J)][0], C[w_tile(J, J)][1], U[w_tile(J, J)][0], U[w_tile(J, J)][1], L);
#pragma unroll
  for (int I = 1; I < 4; I++) {
#pragma unroll
    for (int J = 0; J < I; J++) {
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int K = J; K < I; K++)  // L_IK = X_IK D_K^-1
        w_abt(s0, s1, U[w_tile(K, J)][0], U[w_tile(K, J)][1], C[w_tile(I, K)][0] * rv[K][0], C[w_tile(I, K)][1] * rv[K][1]);
      double t0 = 0.0, t1 = 0.0;
      w_abt(t0, t1, s0, s1, C[w_tile(I, I)][0], C[w_tile(I, I)][1]);
      U[w_tile(I, J)][0] = -t0; U[w_tile(I, J)][1] = -t1;
    }
  }
  // ---- Q^-1 = U D^-1 U^T: full 4 x 4 tiles (Qi[I][J], I >= J computed, I < J transposed)
  double Qi[4][4][2];
#pragma unroll
  for (int I = 0; I < 4; I++)
#pragma unroll
    for (int J = 0; J <= I; J++) {
      double c0 = 0.0, c1 = 0.0;
#pragma unroll
      for (int K = I; K < 4; K++)
        w_abt(c0, c1, U[w_tile(K, I)][0] * rv[K][0], U[w_tile(K, I)][1] * rv[K][1], U[w_tile(K, J)][0], U[w_tile(K, J)][1]);
      Qi[I][J][0] = c0; Qi[I][J][1] = c1;
      if (J < I) w_transpose(c0, c1, Qi[J][I][0], Qi[J][I][1], L);
    }
  {
    double* Qo = a.Qi + (size_t)prob * a.sQi;
    const int ld = a.ldn;
#pragma unroll
    for (int I = 0; I < 4; I++)
#pragma unroll
      for (int J = 0; J < 4; J++) {
        const int i = 8 * I + g, k = 8 * J + 2 * q;
        if (i < n && k < n) Qo[(size_t)i * ld + k] = Qi[I][J][0];
        if (i < n && k + 1 < n) Qo[(size_t)i * ld + k + 1] = Qi[I][J][1];
      }
  }
  // ---- G Q^-1 and R = G Q^-1 G^T, one tile row at a time
  double* Bo = a.BQi + (size_t)prob * a.sBQi;
  double* Rf = a.R + (size_t)prob * a.sR;
  const int ld = a.ldn;
#pragma unroll
  for (int I = 0; I < NTI; I++) {
    if (I < MT) {  // uniform
      double B[4][2];
#pragma unroll
      for (int J = 0; J < 4; J++) {
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int K = 0; K < 4; K++) w_abt(c0, c1, Gt[I][K][0], Gt[I][K][1], Qi[J][K][0], Qi[J][K][1]);
        B[J][0] = c0; B[J][1] = c1;
        const int i = 8 * I + g, k = 8 * J + 2 * q;
        if (i < m && k < n) Bo[(size_t)i * ld + k] = c0;
        if (i < m && k + 1 < n) Bo[(size_t)i * ld + k + 1] = c1;
      }
#pragma unroll
      for (int K = 0; K <= I; K++) {
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int J = 0; J < 4; J++) w_abt(c0, c1, B[J][0], B[J][1], Gt[K][J][0], Gt[K][J][1]);
        *reinterpret_cast<double2*>(Rf + (size_t)w_tile(I, K) * 64 + lane * 2) = make_double2(c0, c1);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// QPFunctionFn.backward (qpth/qp.py:129-183) with one warp per QP and the same building blocks: d = clamp(lam) /
// clamp(slack), T = R + diag(1/d) factored in registers with the right-hand side G Q^-1 dl/dz riding as the
// bordered row, one backward sweep, dx = -Q^-1 dl/dz - (G Q^-1)^T dlam, then the outer-product gradients written with
// 16-byte stores (dQ = (dx z' + z dx') / 2, dG = dlam z' + lam dx', dp = dx, dh = -dlam).  A failed factorisation
// (some lam / slack is NaN or the clamped T is not positive definite) gives NaN gradients, as the reference's LU does.
template <int NTI, int NC, int MC>
__global__ void __launch_bounds__(32, 8) k_wres_backward(const KArgs<double> a, const BArgs<double> ga) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NT = NTI * (NTI + 1) / 2, MPAD = 8 * NTI, RPL = (MPAD + 31) / 32;
  WLane L;
  L.init();
  const int lane = L.lane, g = L.g, q = L.q;
  const int prob = blockIdx.x + a.prob0;
  const int n = NC > 0 ? NC : a.n, m = MC > 0 ? MC : a.m;
  const int mr = (m + 1) & ~1;
  const WOff o = wres_off(m);
  double* sm = reinterpret_cast<double*>(smem_raw);
  double* sG = sm + o.G;
  double* sx = sm + o.x; double* srx = sm + o.rx; double* st = sm + o.t; double* sqx = sm + o.qx;
  double* sz = sm + o.z; double* sdz = sm + o.dz; double* sdinv = sm + o.dinv; double* shz = sm + o.hz;
  double* su = sm + o.u; double* sq = sm + o.q; double* srinv = sm + o.rinv;
  {
    // the inputs Q, G are not part of the backward call (include/b200qp.h): B = G Q^-1 from the pre-factorisation
    const double* Bg = a.BQi + (size_t)prob * a.sBQi;
    for (int r = 0; r < mr; r++) {
      if (r < m && lane < n) cp_async8(sG + r * kWLd + lane, Bg + (size_t)r * a.ldn + lane);
      else sG[r * kWLd + lane] = 0.0;
    }
    cp_async_commit();
    for (int i = lane; i < o.total - o.x; i += 32) sm[o.x + i] = 0.0;
  }
  const double* Qig = a.Qi + (size_t)prob * a.sQi;
  const double* Rf = a.R + (size_t)prob * a.sR;
  WMat MQi;
  w_gload(MQi, Qig, a.ldn, n, g, q);
  double C[NT][2];
#pragma unroll
  for (int t = 0; t < NT; t++) {
    const double2 v = __ldg(reinterpret_cast<const double2*>(Rf + (size_t)t * 64 + lane * 2));
    C[t][0] = v.x; C[t][1] = v.y;
  }
  __syncwarp();
  {
    const double* zh = ga.zhat + (size_t)prob * n;
    const double* gz = ga.gz + (size_t)prob * n;
    const double* lam = ga.lams + (size_t)prob * m;
    const double* sl = ga.slacks + (size_t)prob * m;
    if (lane < n) { sx[lane] = zh[lane]; srx[lane] = gz[lane]; }
#pragma unroll
    for (int s = 0; s < RPL; s++) {
      const int i = s * 32 + lane;
      if (i < m) {
        const double lv = lam[i], sv = sl[i];
        const double lc = lv < 1e-8 ? 1e-8 : lv, sc = sv < 1e-8 ? 1e-8 : sv;  // qp.py:146-149 (NaN stays NaN)
        sz[i] = lv;
        sdinv[i] = 1.0 / (lc / sc);
      }
    }
  }
  cp_async_wait_all();
  __syncwarp();
  w_gapply(MQi, srx, st, g, q);  // t = Q^-1 dl/dz
  __syncwarp();
  {
    double gt[RPL];
    w_gv<RPL>(sG, srx, n, mr, lane, gt);  // hz = G Q^-1 dl/dz
#pragma unroll
    for (int s = 0; s < RPL; s++) { const int i = s * 32 + lane; if (i < m) shz[i] = gt[s]; }
  }
  __syncwarp();
  if constexpr (MC > 0 && (MC + 8) / 8 == NTI) w_fix_tiles_clean<NTI, MC>(C, sdinv, shz, g, q);
  else w_fix_tiles<NTI, MC>(C, sdinv, shz, m, g, q);
  const bool ok = w_factor<NTI, MC>(C, srinv, su, m, L);
  __syncwarp();
  w_bwd<NTI, MC>(C, su, srinv, sq, m, lane, g, q);
  // dlam = -T^-1 hz
#pragma unroll
  for (int s = 0; s < RPL; s++) { const int i = s * 32 + lane; if (i < m) sdz[i] = ok ? -sq[i] : t_nan<double>(); }
  __syncwarp();
  {
    double gd0, gd1;
    w_gtu(sG, sdz, mr, lane, gd0, gd1);
    if (lane < 16) {  // dx = -t - (G Q^-1)^T dlam
      const double2 tv = *reinterpret_cast<const double2*>(st + 2 * lane);
      *reinterpret_cast<double2*>(sqx + 2 * lane) = make_double2(-tv.x - gd0, -tv.y - gd1);
    }
    __syncwarp();
  }
  // ---- gradients
  double* dp = ga.dp + (size_t)prob * n; double* dh = ga.dh + (size_t)prob * m;
  if (lane < n) dp[lane] = sqx[lane];
#pragma unroll
  for (int s = 0; s < RPL; s++) { const int i = s * 32 + lane; if (i < m) dh[i] = -sdz[i]; }
  if (ga.dQ == nullptr) return;  // factored gradients (B200QP_FLAG_FACTORED_GRAD): dp, dh, zhat, lams are the factors
  double* dQ = ga.dQ + (size_t)prob * n * n;
  double* dG = ga.dG + (size_t)prob * m * n;
  if ((n & 1) == 0) {
    const int l = lane & 15, h = lane >> 4, c = 2 * l;
    const double2 zc = *reinterpret_cast<const double2*>(sx + c), dc = *reinterpret_cast<const double2*>(sqx + c);
    if (c < n) {
      for (int r = h; r < n; r += 2) {
        const double dr_ = 0.5 * sqx[r], zr_ = 0.5 * sx[r];
        *reinterpret_cast<double2*>(dQ + (size_t)r * n + c) = make_double2(fma(dr_, zc.x, zr_ * dc.x), fma(dr_, zc.y, zr_ * dc.y));
      }
      for (int r = h; r < m; r += 2) {
        const double dl = sdz[r], lm = sz[r];
        *reinterpret_cast<double2*>(dG + (size_t)r * n + c) = make_double2(fma(dl, zc.x, lm * dc.x), fma(dl, zc.y, lm * dc.y));
      }
    }
  } else {
    if (lane < n) {
      const double zc = sx[lane], dc = sqx[lane];
      for (int r = 0; r < n; r++) dQ[(size_t)r * n + lane] = 0.5 * (sqx[r] * zc + sx[r] * dc);
      for (int r = 0; r < m; r++) dG[(size_t)r * n + lane] = sdz[r] * zc + sz[r] * dc;
    }
  }
}

}  // namespace b200qp
