// qp_host.cuh -- host-side workspace layout and launcher declarations shared by the translation
// units of libb200qp.so (kernels are instantiated per (dtype, residency) in qp_inst.cu).
#pragma once
#include <cuda_runtime.h>
#include "../../include/b200qp.h"
#include <stdlib.h>
#include "qp_kernels.cuh"
#include "qp_fast.cuh"
#include "qp_resident.cuh"
#include "qp_wres.cuh"

namespace b200qp {

int cuda_fail(cudaError_t e, const char* what);
#define CK(call)                                        \
  do {                                                  \
    cudaError_t e_ = (call);                            \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
  } while (0)

constexpr size_t kSmemResidentLimit = 200 * 1024;  // bytes/CTA we are willing to ask for
constexpr size_t kSmemMax = 227 * 1024;

struct Layout {
  int nb, n, m, p, ldn, ldm, ldp, nt, mpad;
  bool smem, fast;
  int fk;  // 1: DMMA factorisation + fragment-order R (qp_dmma.cuh)
  size_t es, smem_bytes;
  long long sQi, sBQi, sR, sV, sUA, sF, sT;
  size_t oQi, oBQi, oR, oV, oUA, opinvA, oF, opinvF, oT, opinvT;
  size_t ox, os, oz, oy, odx, ods, odz, ody, ormu, oflags, obest, oslots, octl, total;
  // resident route (qp_resident.cuh): several iterations per launch, history + records per problem
  bool res;
  int res_chunk, res_spec, res_warp;
  bool wpre;  // pre-factorisation by k_wres_prefactor (one warp per QP, qp_wres.cuh)
  size_t res_smem, ohist, orec, opst;
};

// Process-wide tuning knobs.  Read from the environment ONCE (first use) and changeable through
// b200qp_set_option; a layout never depends on anything else, so forward and backward of one problem
// descriptor always agree on it.
struct Options {
  int res = 1;        // B200QP_RES       0: never take the resident route
  int res_chunk = 10; // B200QP_RES_CH    iterations per launch of the resident route
  int res_spec = 1;   // B200QP_RES_SPEC  1: use the compile-time-size specialisation where one exists (nz=30, nineq=60)
  int res_warp = 1;   // B200QP_RES_WARP  1: one warp per QP with the factor in registers (qp_wres.cuh), 0: 128-thread CTAs
  int res_pre = 1;    // B200QP_RES_PRE   1: warp-per-QP pre-factorisation on the shapes of the resident route
  int mid_fast = 0;   // B200QP_MID=fast  64 < nineq <= 128 on the register-tile route
  int blk_nt = 0;     // B200QP_BLK_NT    256 forces the wide CTAs of the blocked route
  int factor_tile = 0;// B200QP_FACTOR=tile
  int force_generic = 0;  // B200QP_FORCE_GENERIC
};
Options& options();

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline int make_layout(const b200qp_problem_t* pr, Layout& L) {
  if (!pr || pr->nb < 1 || pr->nz < 1 || pr->nineq < 1 || pr->neq < 0) return B200QP_EINVAL;
  if (pr->dtype != B200QP_F64 && pr->dtype != B200QP_F32) return B200QP_EINVAL;
  if (pr->max_iter < 1 || pr->max_iter > B200QP_MAX_ITER_CAP) return B200QP_EINVAL;
  L.nb = pr->nb; L.n = pr->nz; L.m = pr->nineq; L.p = pr->neq;
  L.ldn = L.n | 1; L.ldm = L.m | 1; L.ldp = (L.p > 0 ? L.p : 1) | 1;
  L.es = pr->dtype == B200QP_F64 ? 8 : 4;
  const int widest = (L.n > L.p + L.m) ? L.n : (L.p + L.m);
  L.nt = (widest <= 128 && L.m <= 64) ? 128 : 256;
  const size_t full = smem_elems(L.n, L.m, L.p, L.ldn, L.ldm, L.ldp, L.nt, true) * L.es;
  const size_t vecs = smem_elems(L.n, L.m, L.p, L.ldn, L.ldm, L.ldp, L.nt, false) * L.es;
  L.smem = full <= kSmemResidentLimit;
  L.smem_bytes = L.smem ? full : vecs;
  // fast path (qp_fast.cuh): nineq <= 64 with 128 threads, <= 128 with 256 threads
  L.mpad = L.m <= 32 ? 32 : (L.m <= 64 ? 64 : 128);
  L.fk = 0;
  // fp64 problems with nineq > 64 take the global-resident blocked tensor-core route (qp_blocked.cuh): measured
  // 2-3x faster than the register-tile fast path at 64 < nineq <= 128 (B200QP_MID=fast restores the old routing)
  const Options& opt = options();
  const bool blocked_route = pr->dtype == B200QP_F64 && L.m > 64 && !(pr->flags & B200QP_FLAG_DENSE) && !opt.mid_fast;
  if (blocked_route) {
    L.smem = false;
    L.smem_bytes = vecs;
    // 128-thread CTAs (6 per SM instead of 3 x 256 threads) for the smaller blocked problems: the kernel is latency
    // bound, more problems in flight win -- nineq = 80 / 100 / 128: 173 k -> 239 k, 133 k -> 193 k, 111 k -> 136 k
    // solves/s (B200QP_BLK_NT=256 restores the wider CTAs)
    const size_t s128 = smem_elems(L.n, L.m, L.p, L.ldn, L.ldm, L.ldp, 128, false) * L.es;
    if (opt.blk_nt != 256 && kSmemMax / (s128 + 1024) >= 4) {  // at least four narrow CTAs fit on an SM
      L.nt = 128;
      L.smem_bytes = s128;
    }
  }
  L.fast = !blocked_route && L.m <= 128 && (L.m <= 64 || widest <= 256) && !opt.force_generic;
  if (L.fast) {
    const int generic_nt = L.nt;
    L.nt = L.m <= 64 ? 128 : 256;  // 32/64-thread CTA variants were measured slower and are not instantiated
    // DMMA factorisation: fp64, 128-thread CTAs, one spare row for the bordered right-hand side
    L.fk = (pr->dtype == B200QP_F64 && L.nt == 128 && L.m < 64 && !opt.factor_tile) ? 1 : 0;
    if (L.fk) L.mpad = L.m < 32 ? 32 : 64;
    const size_t fb = fast_smem_elems(L.n, L.m, L.p, L.ldn, L.ldm, L.ldp, L.nt, L.mpad, L.fk) * L.es;
    if ((L.nt == 128 && widest > 128) || fb > kSmemResidentLimit) { L.fast = false; L.fk = 0; L.nt = generic_nt; }
    else { L.smem = true; L.smem_bytes = fb; }
  }
  if (L.smem_bytes > kSmemMax) return B200QP_ETOOBIG;
  if ((pr->flags & B200QP_FLAG_DENSE) && !L.fast) return B200QP_ETOOBIG;  // dense mode lives in the fast kernels
  const int pp = L.p > 0 ? L.p : 1;
  L.sQi = round4(L.n * L.ldn);
  L.sBQi = round4((L.p + L.m) * L.ldn);
  L.sR = L.fk ? frag_elems(L.mpad) : (L.fast ? round4(rtile_elems(L.mpad, L.nt)) : round4(L.m * L.ldm));
  L.sV = round4(pp * L.ldm);
  L.sUA = round4(pp * L.ldp);
  L.sF = round4(L.n * L.ldn);
  L.sT = L.smem ? 0 : round4(L.m * L.ldm);
  size_t off = 0;
  const size_t nb = (size_t)L.nb;
  auto put = [&](size_t elems, size_t esz) { size_t o = off; off = align_up(off + elems * esz, 256); return o; };
  L.oQi = put(nb * L.sQi, L.es);
  L.oBQi = put(nb * L.sBQi, L.es);
  L.oR = put(nb * L.sR, L.es);
  L.oV = put(L.p > 0 ? nb * L.sV : 4, L.es);
  L.oUA = put(L.p > 0 ? nb * L.sUA : 4, L.es);
  L.opinvA = put(nb * round4(pp), L.es);
  L.oF = put(nb * L.sF, L.es);
  L.opinvF = put(nb * round4(L.n), L.es);
  L.oT = put(L.smem ? 4 : nb * L.sT, L.es);
  L.opinvT = put(L.smem ? 4 : nb * round4(L.m), L.es);
  L.ox = put(nb * round4(L.n), L.es);
  L.os = put(nb * round4(L.m), L.es);
  L.oz = put(nb * round4(L.m), L.es);
  L.oy = put(nb * round4(pp), L.es);
  L.odx = put(nb * round4(L.n), L.es);
  L.ods = put(nb * round4(L.m), L.es);
  L.odz = put(nb * round4(L.m), L.es);
  L.ody = put(nb * round4(pp), L.es);
  L.ormu = put(nb * 2, L.es);
  L.oflags = put(nb, sizeof(int));
  L.obest = put(nb, sizeof(double));
  L.oslots = put(B200QP_MAX_ITER_CAP, sizeof(Slot));
  L.octl = put(1, sizeof(Control));
  // resident route: fp64, no equalities, nineq < 64 (one spare row for the bordered right-hand side), nz <= 32
  L.res = opt.res && L.fast && L.fk && L.p == 0 && L.n <= 32 && pr->max_iter <= kResMaxIter &&
          !(pr->flags & (B200QP_FLAG_DENSE | B200QP_FLAG_EXACT));
  L.res_chunk = opt.res_chunk < 1 ? 1 : opt.res_chunk;
  L.res_spec = opt.res_spec ? 1 : 0;
  L.res_warp = opt.res_warp;  // 0: 128-thread CTAs per QP; 1 / 4 / 8: one warp per QP, that many QPs per CTA
  L.res_smem = (size_t)res_off(L.n, L.m, L.mpad).total * sizeof(double);
  L.ohist = L.orec = L.opst = 0;
  // same shapes, but independent of the EXACT flag: the exact re-run of a call starts from the same pre-factorisation
  L.wpre = opt.res_pre && opt.res_warp && L.fast && L.fk && L.p == 0 && L.n <= 32 && !(pr->flags & B200QP_FLAG_DENSE);
  if (L.res) {
    L.ohist = put(nb * (size_t)(pr->max_iter + 1) * res_hs(L.n, L.m), sizeof(double));
    L.orec = put(nb * (size_t)pr->max_iter * 2, sizeof(double));
    L.opst = put(nb * kPst, sizeof(int));
  }
  L.total = off;
  return B200QP_OK;
}


// Launchers, explicitly instantiated in qp_inst.cu for T in {double,float} x SMEM in {true,false}.
template <typename T, bool SMEM> int launch_iter(const KArgs<T>& a, const Layout& L, cudaStream_t st);
template <typename T, bool SMEM> int launch_backward(const KArgs<T>& a, const BArgs<T>& g, const Layout& L, cudaStream_t st);
template <typename T, bool SMEM> int launch_kkt_solve(const KArgs<T>& a, const SArgs<T>& g, const Layout& L, cudaStream_t st);
template <typename T> int launch_prefactor(const KArgs<T>& a, const Layout& L, cudaStream_t st);
// fast-path launchers, explicitly instantiated in qp_inst.cu -DINST_FAST=1 for T in {double,float}
template <typename T> int fast_iter(const KArgs<T>& a, const Layout& L, cudaStream_t st);
template <typename T> int fast_backward(const KArgs<T>& a, const BArgs<T>& g, const Layout& L, cudaStream_t st);
template <typename T> int fast_kkt(const KArgs<T>& a, const SArgs<T>& g, const Layout& L, cudaStream_t st);
// resident route (qp_inst.cu -DINST_RES=1): one launch = the iterations [.., ra.it_end) of every problem
int res_chunk(const KArgs<double>& a, const RArgs& ra, const Layout& L, cudaStream_t st);
int res_finish(const KArgs<double>& a, const RArgs& ra, double* status, int launches, cudaStream_t st);
int res_backward(const KArgs<double>& a, const BArgs<double>& g, const Layout& L, cudaStream_t st);  // warp per QP (qp_wres.cuh)
int res_prefactor(const KArgs<double>& a, const Layout& L, cudaStream_t st);                        // warp per QP (qp_wres.cuh)

}  // namespace b200qp
