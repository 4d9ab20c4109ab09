// qp_common.cuh -- shared device helpers for the batched PDIPM kernels (sm_100a).
//
// Layout conventions
//   * one CTA per QP; every per-problem matrix is row-major with an ODD leading dimension
//     (ldn = nz|1, ldm = nineq|1, ldp = neq|1) so that both "thread walks a row" and "thread
//     walks a column" are shared-memory bank-conflict free for 8-byte and 4-byte words;
//   * symmetric-positive-definite blocks are factored as  S = L D L^T  (no square roots, unit
//     diagonal => the multiply by 1/l_jj leaves the triangular-solve dependency chain); the
//     factor is kept as the unit UPPER triangle U = L^T (U[j][i] = L[i][j], i > j) plus the
//     vector of reciprocal pivots pinv[j] = 1/D_j.  A non-positive or NaN pivot "poisons" the
//     problem: its iterates become NaN and it can never improve again, which is what the
//     reference's partial-pivot LU degenerates to once z/s has gone negative/NaN
//     ("Pivoting / Cholesky vs LU").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200qp {

// ----------------------------------------------------------------------------- numeric helpers
template <typename T> __device__ __forceinline__ T t_nan();
template <> __device__ __forceinline__ double t_nan<double>() { return __longlong_as_double(0x7ff8000000000000LL); }
template <> __device__ __forceinline__ float t_nan<float>() { return __int_as_float(0x7fc00000); }
template <typename T> __device__ __forceinline__ T t_inf();
template <> __device__ __forceinline__ double t_inf<double>() { return __longlong_as_double(0x7ff0000000000000LL); }
template <> __device__ __forceinline__ float t_inf<float>() { return __int_as_float(0x7f800000); }

template <typename T> __device__ __forceinline__ bool is_nan(T v) { return v != v; }

// torch.min / torch.max semantics: NaN propagates.
template <typename T> __device__ __forceinline__ T nanmin(T a, T b) {
  return is_nan(a) ? a : (is_nan(b) ? b : (b < a ? b : a));
}
template <typename T> __device__ __forceinline__ T nanmax(T a, T b) {
  return is_nan(a) ? a : (is_nan(b) ? b : (b > a ? b : a));
}

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ float shfl_d(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double shfl_x(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ float shfl_x(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// Total order key for doubles (non-NaN): a < b  <=>  key(a) < key(b) as unsigned.
__device__ __forceinline__ unsigned long long ord_key(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}
__device__ __forceinline__ double ord_unkey(unsigned long long k) {
  unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
  return __longlong_as_double((long long)b);
}
__device__ __forceinline__ void atomic_max_key(unsigned long long* addr, unsigned long long key) {
  if (*(volatile unsigned long long*)addr < key) atomicMax(addr, key);
}

// CTA-wide barrier; a one-warp CTA (warp-per-problem kernels) only needs a warp barrier.
template <int NT> __device__ __forceinline__ void cta_sync() {
  if (NT == 32) __syncwarp(); else __syncthreads();
}
__device__ __forceinline__ void cta_sync_rt(int nt) {
  if (nt == 32) __syncwarp(); else __syncthreads();
}

// ----------------------------------------------------------------------------- block reductions
struct OpSum { template <typename T> __device__ __forceinline__ T operator()(T a, T b) const { return a + b; } };
struct OpNanMin { template <typename T> __device__ __forceinline__ T operator()(T a, T b) const { return nanmin(a, b); } };
struct OpNanMax { template <typename T> __device__ __forceinline__ T operator()(T a, T b) const { return nanmax(a, b); } };
struct OpMin { template <typename T> __device__ __forceinline__ T operator()(T a, T b) const { return b < a ? b : a; } };

// Reduce K values per thread across the CTA; every thread gets the results. `scratch` holds
// K*32 elements.  Fixed tree order => bitwise reproducible, independent of the batch.
template <int K, typename T, typename Op>
__device__ __forceinline__ void block_reduce(T (&v)[K], Op op, T* scratch, int tid, int nt) {
  const int lane = tid & 31, warp = tid >> 5, nw = (nt + 31) >> 5;
#pragma unroll
  for (int k = 0; k < K; k++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] = op(v[k], shfl_x(v[k], o));
  }
  if (nt == 32) return;  // one warp: the shuffle tree already left the result in every lane
  __syncthreads();  // scratch may still be read by a previous reduction
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; k++) scratch[k * 32 + warp] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; k++) {
    T r = scratch[k * 32];
    for (int w = 1; w < nw; w++) r = op(r, scratch[k * 32 + w]);
    v[k] = r;
  }
}

// ----------------------------------------------------------------------------- small BLAS
// out[c] = sum_r M[r*ld + c] * v[r]   (c < cols).  Threads walk columns => coalesced / conflict
// free for row-major M.  `part` holds nt elements.  Ends with a __syncthreads().
template <typename T>
__device__ __forceinline__ void gemv_cols(const T* __restrict__ M, int ld, int rows, int cols,
                                          const T* v, T* out, T* part, int tid, int nt) {
  if (cols <= nt) {
    int groups = nt / cols;
    if (groups > rows) groups = rows > 0 ? rows : 1;
    const int c = tid % cols, g = tid / cols;
    if (g < groups) {
      T a0 = 0, a1 = 0;
      int r = g;
#pragma unroll 4
      for (; r + groups < rows; r += 2 * groups) {
        a0 += M[(size_t)r * ld + c] * v[r];
        a1 += M[(size_t)(r + groups) * ld + c] * v[r + groups];
      }
      if (r < rows) a0 += M[(size_t)r * ld + c] * v[r];
      part[tid] = a0 + a1;
    }
    __syncthreads();
    if (tid < cols) {
      T s = part[tid];
      for (int gg = 1; gg < groups; gg++) s += part[gg * cols + tid];
      out[tid] = s;
    }
  } else {
    for (int c = tid; c < cols; c += nt) {
      T a0 = 0, a1 = 0;
      int r = 0;
#pragma unroll 4
      for (; r + 1 < rows; r += 2) {
        a0 += M[(size_t)r * ld + c] * v[r];
        a1 += M[(size_t)(r + 1) * ld + c] * v[r + 1];
      }
      if (r < rows) a0 += M[(size_t)r * ld + c] * v[r];
      out[c] = a0 + a1;
    }
  }
  __syncthreads();
}

// Same contract as gemv_cols with the CTA size as a template parameter: column groups when the
// CTA has at least two threads per column, otherwise 4-way unrolled serial sums per column.
template <typename T, int NT>
__device__ __forceinline__ void gemv_cols_nt(const T* __restrict__ M, int ld, int rows, int cols, const T* v, T* out,
                                             T* part, int tid) {
  if (2 * cols <= NT) {
    const int groups = NT / cols, c = tid % cols, g = tid / cols;
    if (g < groups) {
      T a0 = 0, a1 = 0;
      int r = g;
#pragma unroll 4
      for (; r + groups < rows; r += 2 * groups) {
        a0 += M[(size_t)r * ld + c] * v[r];
        a1 += M[(size_t)(r + groups) * ld + c] * v[r + groups];
      }
      if (r < rows) a0 += M[(size_t)r * ld + c] * v[r];
      part[tid] = a0 + a1;
    }
    cta_sync<NT>();
    if (tid < cols) {
      T s = part[tid];
      for (int gg = 1; gg < groups; gg++) s += part[gg * cols + tid];
      out[tid] = s;
    }
  } else {
    for (int c = tid; c < cols; c += NT) {
      T a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      int r = 0;
      for (; r + 3 < rows; r += 4) {
        a0 += M[(size_t)r * ld + c] * v[r];
        a1 += M[(size_t)(r + 1) * ld + c] * v[r + 1];
        a2 += M[(size_t)(r + 2) * ld + c] * v[r + 2];
        a3 += M[(size_t)(r + 3) * ld + c] * v[r + 3];
      }
      for (; r < rows; r++) a0 += M[(size_t)r * ld + c] * v[r];
      out[c] = (a0 + a1) + (a2 + a3);
    }
  }
  cta_sync<NT>();
}

// ----------------------------------------------------------------------------- register-resident GEMV
// Reduce K (8 or 16) per-lane values across the warp with a halving butterfly (K + log2(32/K) - 1
// shuffles instead of 5 K).  On return v[0] of lane l is the warp-wide sum of value index
// (l >> 1) & 15 for K = 16, (l >> 2) & 7 for K = 8.
template <int K, typename T>
__device__ __forceinline__ void warp_reduce_multi(T (&v)[K], int lane) {
  static_assert(K == 8 || K == 16, "K must be 8 or 16");
  if (K == 16) {
    const bool up = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { const T snd = up ? v[i] : v[i + 8], kp = up ? v[i + 8] : v[i]; v[i] = kp + shfl_x(snd, 16); }
  }
  {
    const int off = K == 16 ? 8 : 16;
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < 4; i++) { const T snd = up ? v[i] : v[i + 4], kp = up ? v[i + 4] : v[i]; v[i] = kp + shfl_x(snd, off); }
  }
  {
    const int off = K == 16 ? 4 : 8;
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < 2; i++) { const T snd = up ? v[i] : v[i + 2], kp = up ? v[i + 2] : v[i]; v[i] = kp + shfl_x(snd, off); }
  }
  {
    const int off = K == 16 ? 2 : 4;
    const bool up = (lane & off) != 0;
    const T snd = up ? v[0] : v[1], kp = up ? v[1] : v[0];
    v[0] = kp + shfl_x(snd, off);
  }
  if (K == 8) v[0] += shfl_x(v[0], 2);
  v[0] += shfl_x(v[0], 1);
}

// A (rows x cols <= 4*RPW x 32) matrix held in the registers of a 4-warp CTA: lane = column,
// warp w owns rows w, w+4, ...  Loaded ONCE with coalesced loads (issue it early: the latency
// hides behind whatever follows) and then used for M v (regmat_rows) and M^T u (regmat_cols).
template <int RPW, typename T>
struct RegMat { T e[RPW]; };

template <int RPW, typename T>
__device__ __forceinline__ void regmat_load(RegMat<RPW, T>& M, const T* __restrict__ g, int ld, int rows, int cols, int lane,
                                            int warp) {
#pragma unroll
  for (int k = 0; k < RPW; k++) {
    const int r = warp + 4 * k;
    M.e[k] = (r < rows && lane < cols) ? g[(size_t)r * ld + lane] : T(0);
  }
}
// out[r] = sum_c M[r][c] v[c];  v, out in shared memory.  No barrier.
template <int RPW, typename T>
__device__ __forceinline__ void regmat_rows(const RegMat<RPW, T>& M, const T* v, T* out, int rows, int cols, int lane, int warp) {
  const T vc = lane < cols ? v[lane] : T(0);
  T pr[RPW];
#pragma unroll
  for (int k = 0; k < RPW; k++) pr[k] = M.e[k] * vc;
  warp_reduce_multi<RPW>(pr, lane);
  const int k = RPW == 16 ? (lane >> 1) & 15 : (lane >> 2) & 7;
  const bool writer = RPW == 16 ? (lane & 1) == 0 : (lane & 3) == 0;
  const int r = warp + 4 * k;
  if (writer && r < rows) out[r] = pr[0];
}
// part[warp*32 + c] = sum over this warp's rows of M[r][c] u[r];  the caller adds the four
// partials after a barrier (regmat_colsum).  u in shared memory.
template <int RPW, typename T>
__device__ __forceinline__ void regmat_cols(const RegMat<RPW, T>& M, const T* u, T* part, int rows, int lane, int warp) {
  T a0 = T(0), a1 = T(0);
#pragma unroll
  for (int k = 0; k < RPW; k += 2) {
    const int r0 = warp + 4 * k, r1 = r0 + 4;
    a0 += M.e[k] * (r0 < rows ? u[r0] : T(0));
    a1 += M.e[k + 1] * (r1 < rows ? u[r1] : T(0));
  }
  part[warp * 32 + lane] = a0 + a1;
}
template <typename T>
__device__ __forceinline__ T regmat_colsum(const T* part, int c) { return (part[c] + part[32 + c]) + (part[64 + c] + part[96 + c]); }

// out[r] = sum_c M[r*ld + c] * v[c]  (r < rows), one thread per row: conflict free in shared
// memory when ld is odd.  No trailing barrier.
template <typename T>
__device__ __forceinline__ void gemv_rows_thread(const T* M, int ld, int rows, int cols, const T* v,
                                                 T* out, int tid, int nt) {
  for (int r = tid; r < rows; r += nt) {
    const T* row = M + (size_t)r * ld;
    T a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int c = 0;
    for (; c + 3 < cols; c += 4) {
      a0 += row[c] * v[c];
      a1 += row[c + 1] * v[c + 1];
      a2 += row[c + 2] * v[c + 2];
      a3 += row[c + 3] * v[c + 3];
    }
    for (; c < cols; c++) a0 += row[c] * v[c];
    out[r] = (a0 + a1) + (a2 + a3);
  }
}

// out[r] = sum_c M[r*ld + c] * v[c], one WARP per row (coalesced for row-major M in global
// memory).  No trailing barrier.
template <typename T>
__device__ __forceinline__ void gemv_rows_warp(const T* __restrict__ M, int ld, int rows, int cols,
                                               const T* v, T* out, int tid, int nt) {
  const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  for (int r0 = warp; r0 < rows; r0 += 4 * nw) {  // 4 rows in flight per warp
    T a[4] = {T(0), T(0), T(0), T(0)};
    for (int c = lane; c < cols; c += 32) {
      const T vc = v[c];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int r = r0 + u * nw;
        if (r < rows) a[u] += M[(size_t)r * ld + c] * vc;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a[u] += shfl_x(a[u], o);
      const int r = r0 + u * nw;
      if (lane == 0 && r < rows) out[r] = a[u];
    }
  }
}

// ----------------------------------------------------------------------------- LDL^T
// In-place S = L D L^T on the lower triangle of S (m x m, leading dim ld).  On exit the strict
// UPPER triangle holds U = L^T (unit diagonal implied) and pinv[j] = 1/D_j.  The lower triangle
// is destroyed.  One barrier per column; the reciprocal of the next pivot is computed by the
// thread that updates it so it overlaps the rest of the trailing update.
// Returns false (for every thread) when a pivot is <= 0 or NaN; pinv[] then ends in NaN.
template <typename T>
__device__ __forceinline__ bool ldlt_factor(T* S, int ld, int m, T* pinv, int tid, int nt) {
  int tpr = 1;
  while (tpr * 2 * m <= nt) tpr *= 2;
  const int rg = tid / tpr, kk = tid % tpr, nrg = nt / tpr;
  if (tid == 0) {
    T d0 = S[0];
    pinv[0] = (d0 > T(0)) ? T(1) / d0 : t_nan<T>();
  }
  __syncthreads();
  bool ok = true;
  for (int j = 0; j < m; j++) {
    const T pj = pinv[j];
    if (is_nan(pj)) { ok = false; break; }  // uniform: written before the last barrier
    for (int i = j + 1 + rg; i < m; i += nrg) {
      T* rowi = S + (size_t)i * ld;
      const T lij = rowi[j] * pj;
      for (int k = j + 1 + kk; k <= i; k += tpr) {
        const T v = rowi[k] - lij * S[(size_t)k * ld + j];
        rowi[k] = v;
        if (k == j + 1 && i == j + 1) pinv[j + 1] = (v > T(0)) ? T(1) / v : t_nan<T>();
      }
      if (kk == 0) S[(size_t)j * ld + i] = lij;
    }
    __syncthreads();
  }
  return ok;
}

// Register-tiled variant for m <= MPAD (compile time): the trailing matrix lives in REGISTERS,
// distributed 2-D cyclically over a TR x TC thread grid (element (i,k) -> thread (i%TR, k%TC)), so a
// column step costs NA+NB shared-memory reads and up to NA*NB FMAs per thread instead of two
// loads + one store per FMA.  Only the current column travels through shared memory (`colbuf`,
// 2*MPAD elements, double buffered): one barrier per column, every thread forms 1/D_j itself.
// Same outputs as ldlt_factor (unit upper U in S, pinv[]); S's lower triangle is left intact.
template <typename T, int MPAD, int NT>
__device__ __forceinline__ bool ldlt_factor_reg(T* S, int ld, int m, T* pinv, T* colbuf, int tid) {
  constexpr int TR = 16, TC = NT / 16;
  constexpr int NA = MPAD / TR, NB = MPAD / TC;
  static_assert(NA >= 1 && NB >= 1, "tile too small for the thread grid");
  const int ti = tid % TR, tk = tid / TR;
  T A[NA][NB];
#pragma unroll
  for (int a = 0; a < NA; a++) {
#pragma unroll
    for (int b = 0; b < NB; b++) {
      if (TC * b <= TR * a + TR - 1) {  // block touches the lower triangle (compile time)
        const int i = ti + TR * a, k = tk + TC * b;
        A[a][b] = (i < m && k <= i) ? S[(size_t)i * ld + k] : T(0);
      }
    }
  }
  if (tk == 0) {
#pragma unroll
    for (int a = 0; a < NA; a++) {
      const int i = ti + TR * a;
      if (i < m) colbuf[i] = A[a][0];
    }
  }
  __syncthreads();
  bool ok = true;
  // j = jb*TC + tkk with jb unrolled: the register column index of column j (and of every block
  // that is still alive) is then a compile-time constant.
#pragma unroll
  for (int jb = 0; jb < NB; jb++) {
    for (int tkk = 0; tkk < TC; tkk++) {
      const int j = jb * TC + tkk;
      if (j >= m || !ok) break;  // uniform
      const T* cb = colbuf + (j & 1) * MPAD;
      T* cbn = colbuf + ((j + 1) & 1) * MPAD;
      const T dj = cb[j];
      if (!(dj > T(0))) { ok = false; break; }  // uniform (same shared word for every thread)
      const T pj = T(1) / dj;
      if (tid == 0) pinv[j] = pj;
      T li[NA], ck[NB];
#pragma unroll
      for (int a = 0; a < NA; a++) {
        const int i = ti + TR * a;
        li[a] = (i > j && i < m) ? cb[i] * pj : T(0);
      }
#pragma unroll
      for (int b = jb; b < NB; b++) {
        const int k = tk + TC * b;
        ck[b] = (k > j && k < m) ? cb[k] : T(0);
      }
      if (tk == tkk) {
#pragma unroll
        for (int a = 0; a < NA; a++) {
          const int i = ti + TR * a;
          if (i > j && i < m) S[(size_t)j * ld + i] = li[a];
        }
      }
#pragma unroll
      for (int a = 0; a < NA; a++) {
        if (TR * a + TR - 1 > j) {  // uniform: block row still alive
#pragma unroll
          for (int b = jb; b < NB; b++) {
            if (TC * b <= TR * a + TR - 1) A[a][b] -= li[a] * ck[b];
          }
        }
      }
      // publish column j+1 (its register column is jb, or jb+1 when j closes this block column)
      const int jn = j + 1;
      if (tkk < TC - 1) {
        if (tk == tkk + 1) {
#pragma unroll
          for (int a = 0; a < NA; a++) {
            if (TC * jb <= TR * a + TR - 1) {
              const int i = ti + TR * a;
              if (i >= jn && i < m) cbn[i] = A[a][jb];
            }
          }
        }
      } else if (jb + 1 < NB) {
        if (tk == 0) {
#pragma unroll
          for (int a = 0; a < NA; a++) {
            if (TC * (jb + 1) <= TR * a + TR - 1) {
              const int i = ti + TR * a;
              if (i >= jn && i < m) cbn[i] = A[a][jb + 1 < NB ? jb + 1 : jb];
            }
          }
        }
      }
      __syncthreads();
    }
  }
  if (!ok) {
    for (int i = tid; i < m; i += NT) pinv[i] = t_nan<T>();
    __syncthreads();
  }
  return ok;
}

// Solve (L D L^T) x = v in place with ONE warp; v lives in shared/global memory, the working
// vector in registers (RPL values per lane), broadcasts by shuffle.
template <typename T, int RPL>
__device__ __forceinline__ void ldlt_solve_regs(const T* U, int ld, int m, const T* pinv, T* v, int lane) {
  T r[RPL];
#pragma unroll
  for (int s = 0; s < RPL; s++) {
    int i = s * 32 + lane;
    r[s] = i < m ? v[i] : T(0);
  }
#pragma unroll
  for (int s = 0; s < RPL; s++) {
    const int jend = min(32, m - s * 32);
    for (int jj = 0; jj < jend; jj++) {
      const int j = s * 32 + jj;
      const T yj = shfl_d(r[s], jj);
      const T* row = U + (size_t)j * ld;
#pragma unroll
      for (int s2 = s; s2 < RPL; s2++) {
        int i = s2 * 32 + lane;
        if (i > j && i < m) r[s2] -= row[i] * yj;
      }
    }
  }
#pragma unroll
  for (int s = 0; s < RPL; s++) {
    int i = s * 32 + lane;
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
        int i = s2 * 32 + lane;
        if (i < j) r[s2] -= U[(size_t)i * ld + j] * xj;
      }
    }
  }
#pragma unroll
  for (int s = 0; s < RPL; s++) {
    int i = s * 32 + lane;
    if (i < m) v[i] = r[s];
  }
}

// Same, any m, working vector stays in memory (one warp, __syncwarp per column).
template <typename T>
__device__ __forceinline__ void ldlt_solve_mem(const T* U, int ld, int m, const T* pinv, T* v, int lane) {
  for (int j = 0; j < m; j++) {
    const T yj = v[j];
    const T* row = U + (size_t)j * ld;
    for (int i = j + 1 + lane; i < m; i += 32) v[i] -= row[i] * yj;
    __syncwarp();
  }
  for (int i = lane; i < m; i += 32) v[i] *= pinv[i];
  __syncwarp();
  for (int j = m - 1; j > 0; j--) {
    const T xj = v[j];
    for (int i = lane; i < j; i += 32) v[i] -= U[(size_t)i * ld + j] * xj;
    __syncwarp();
  }
}

// Dispatch; call from ALL threads of warp 0 only (v must be visible: barrier before and after).
template <typename T>
__device__ __forceinline__ void ldlt_solve_warp(const T* U, int ld, int m, const T* pinv, T* v, int lane) {
  if (m <= 32) ldlt_solve_regs<T, 1>(U, ld, m, pinv, v, lane);
  else if (m <= 64) ldlt_solve_regs<T, 2>(U, ld, m, pinv, v, lane);
  else if (m <= 96) ldlt_solve_regs<T, 3>(U, ld, m, pinv, v, lane);
  else if (m <= 128) ldlt_solve_regs<T, 4>(U, ld, m, pinv, v, lane);
  else ldlt_solve_mem<T>(U, ld, m, pinv, v, lane);
}

// Unit-lower forward solve only:  v <- L^-1 v  (one warp, memory resident; used for the small
// equality block).
template <typename T>
__device__ __forceinline__ void unit_fwd_warp(const T* U, int ld, int m, T* v, int lane) {
  for (int j = 0; j < m; j++) {
    const T yj = v[j];
    const T* row = U + (size_t)j * ld;
    for (int i = j + 1 + lane; i < m; i += 32) v[i] -= row[i] * yj;
    __syncwarp();
  }
}
// Unit-upper backward solve only:  v <- L^-T v.
template <typename T>
__device__ __forceinline__ void unit_bwd_warp(const T* U, int ld, int m, T* v, int lane) {
  for (int j = m - 1; j > 0; j--) {
    const T xj = v[j];
    for (int i = lane; i < j; i += 32) v[i] -= U[(size_t)i * ld + j] * xj;
    __syncwarp();
  }
}

// 16-byte cooperative copy global -> shared (count in elements, both 16B aligned, count*sizeof(T)
// a multiple of 16).  Uses cp.async (LDGSTS); wait with cp_async_wait_all().
template <typename T>
__device__ __forceinline__ void cp_async_block(T* dst_smem, const T* src, int count, int tid, int nt) {
  const int n16 = (count * (int)sizeof(T)) >> 4;
  const char* s = reinterpret_cast<const char*>(src);
  unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  for (int i = tid; i < n16; i += nt) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d + (unsigned)i * 16u), "l"(s + (size_t)i * 16));
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

}  // namespace b200qp
