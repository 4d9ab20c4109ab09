// qp_inst.cu -- kernel instantiations, compiled six times (parallel build):
//   -DINST_T=double|float -DINST_FAST=1                      the fast-path kernels (qp_fast.cuh, qp_dmma.cuh)
//   -DINST_T=double|float -DINST_FAST=0 -DINST_SMEM=1|0      the generic kernels (+ -DINST_PREFACTOR once per dtype)
#include <type_traits>
#include "qp_host.cuh"

namespace b200qp {

template <typename K>
static cudaError_t ensure_smem(K kernel, size_t bytes) {
  // The attribute is sticky per function; raising it every call costs ~1 us and keeps this
  // stateless (the reference's module-global cache, batch.py:431, is what we avoid).
  if (bytes > 48 * 1024) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
  return cudaSuccess;
}

#define LAUNCH_NT(KNAME, ...)                                                                  \
  do {                                                                                         \
    if (L.nt == 128) { auto k = KNAME<T, SMEM, 128>; CK(ensure_smem(k, L.smem_bytes)); k<<<(unsigned)a.nb, 128, L.smem_bytes, st>>>(__VA_ARGS__); } \
    else { auto k = KNAME<T, SMEM, 256>; CK(ensure_smem(k, L.smem_bytes)); k<<<(unsigned)a.nb, 256, L.smem_bytes, st>>>(__VA_ARGS__); } \
    CK(cudaGetLastError());                                                                    \
    return B200QP_OK;                                                                          \
  } while (0)

#if INST_RES
// resident route (qp_resident.cuh): MPAD in {32, 64} with run-time sizes, plus the compile-time-size specialisation of the
// headline shape of BASELINE.json (nz = 30, nineq = 60)
template <int MPAD, int NC, int MC>
static int res_launch(const KArgs<double>& a, const RArgs& ra, const Layout& L, cudaStream_t st) {
  auto k = k_res_chunk<MPAD, NC, MC>;
  CK(ensure_smem(k, L.res_smem));
  k<<<(unsigned)a.nb, 128, L.res_smem, st>>>(a, ra);
  CK(cudaGetLastError());
  return B200QP_OK;
}
template <int NTI, int NC, int MC, int WPC>
static int wres_launch(const KArgs<double>& a, const RArgs& ra, const Layout& L, cudaStream_t st) {
  auto k = k_wres_chunk<NTI, NC, MC, WPC>;
  const size_t smem = (size_t)wres_off(L.m).total * sizeof(double) * WPC;
  static bool once = false;  // per instantiation
  if (!once) {
    CK(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    once = true;
  }
  CK(ensure_smem(k, smem));
  k<<<(unsigned)((a.nb + WPC - 1) / WPC), 32 * WPC, smem, st>>>(a, ra);
  CK(cudaGetLastError());
  return B200QP_OK;
}
int res_chunk(const KArgs<double>& a, const RArgs& ra, const Layout& L, cudaStream_t st) {
  if (L.res_warp) {  // one warp per QP (qp_wres.cuh)
    if (L.res_spec && L.n == 30 && L.m == 60)
      return L.res_warp == 8 ? wres_launch<8, 30, 60, 8>(a, ra, L, st) : (L.res_warp == 4 ? wres_launch<8, 30, 60, 4>(a, ra, L, st) : wres_launch<8, 30, 60, 1>(a, ra, L, st));
    return L.m < 32 ? wres_launch<4, 0, 0, 1>(a, ra, L, st) : wres_launch<8, 0, 0, 1>(a, ra, L, st);
  }
  if (L.res_spec && L.n == 30 && L.m == 60) return res_launch<64, 30, 60>(a, ra, L, st);
  return L.mpad == 32 ? res_launch<32, 0, 0>(a, ra, L, st) : res_launch<64, 0, 0>(a, ra, L, st);
}
template <int NTI, int NC, int MC>
static int wres_bwd_launch(const KArgs<double>& a, const BArgs<double>& g, const Layout& L, cudaStream_t st) {
  auto k = k_wres_backward<NTI, NC, MC>;
  const size_t smem = (size_t)wres_off(L.m).total * sizeof(double);
  static bool once = false;
  if (!once) {
    CK(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    once = true;
  }
  CK(ensure_smem(k, smem));
  k<<<(unsigned)a.nb, 32, smem, st>>>(a, g);
  CK(cudaGetLastError());
  return B200QP_OK;
}
int res_backward(const KArgs<double>& a, const BArgs<double>& g, const Layout& L, cudaStream_t st) {
  if (L.res_spec && L.n == 30 && L.m == 60) return wres_bwd_launch<8, 30, 60>(a, g, L, st);
  return L.m < 32 ? wres_bwd_launch<4, 0, 0>(a, g, L, st) : wres_bwd_launch<8, 0, 0>(a, g, L, st);
}
template <int NTI, int NC, int MC>
static int wres_pre_launch(const KArgs<double>& a, cudaStream_t st) {
  k_wres_prefactor<NTI, NC, MC><<<(unsigned)a.nb, 32, 0, st>>>(a);
  CK(cudaGetLastError());
  return B200QP_OK;
}
int res_prefactor(const KArgs<double>& a, const Layout& L, cudaStream_t st) {
  if (L.res_spec && L.n == 30 && L.m == 60) return wres_pre_launch<8, 30, 60>(a, st);
  return L.m <= 32 ? wres_pre_launch<4, 0, 0>(a, st) : wres_pre_launch<8, 0, 0>(a, st);
}
int res_finish(const KArgs<double>& a, const RArgs& ra, double* status, int launches, cudaStream_t st) {
  k_res_reduce<0><<<(unsigned)((a.nb + 127) / 128), 128, 0, st>>>(a, ra);
  CK(cudaGetLastError());
  k_res_select<0><<<(unsigned)((a.nb + 3) / 4), 128, 0, st>>>(a, ra, status, launches);
  CK(cudaGetLastError());
  return B200QP_OK;
}
#elif INST_FAST
// fast path: (MPAD, NT) in {(32,128), (64,128), (128,256)}; FK = 1 (DMMA) for fp64 with nineq < 64
#define LAUNCH_ONE(KEXPR, NTV, ...)                                                                \
  do { auto k = KEXPR; CK(ensure_smem(k, L.smem_bytes)); k<<<(unsigned)a.nb, NTV, L.smem_bytes, st>>>(__VA_ARGS__); } while (0)
// (One-warp and two-warp CTA variants of these kernels were measured slower -- profiles/r01/ -- and
// are not instantiated; the kernel templates stay generic in NT.)
#define LAUNCH_FAST(KN, TAIL, ...)                                                                 \
  do {                                                                                             \
    if (L.fk) {                                                                                    \
      if constexpr (std::is_same<T, double>::value) {                                              \
        if (L.mpad == 32) LAUNCH_ONE((KN<T, 32, 128 TAIL COMMA 1>), 128, __VA_ARGS__);        \
        else LAUNCH_ONE((KN<T, 64, 128 TAIL COMMA 1>), 128, __VA_ARGS__);                          \
      }                                                                                            \
    } else if (L.mpad == 128) LAUNCH_ONE((KN<T, 128, 256 TAIL>), 256, __VA_ARGS__);                \
    else {                                                                                       \
      if (L.mpad == 32) LAUNCH_ONE((KN<T, 32, 128 TAIL>), 128, __VA_ARGS__);                       \
      else LAUNCH_ONE((KN<T, 64, 128 TAIL>), 128, __VA_ARGS__);                                    \
    }                                                                                              \
    CK(cudaGetLastError());                                                                        \
    return B200QP_OK;                                                                              \
  } while (0)
#define COMMA ,
template <typename T> int fast_iter(const KArgs<T>& a, const Layout& L, cudaStream_t st) {
  if (a.iter < 0) LAUNCH_FAST(k_fast_iter, COMMA true, a);
  LAUNCH_FAST(k_fast_iter, COMMA false, a);
}
template <typename T> int fast_backward(const KArgs<T>& a, const BArgs<T>& g, const Layout& L, cudaStream_t st) {
  LAUNCH_FAST(k_fast_backward, , a, g);
}
template <typename T> int fast_kkt(const KArgs<T>& a, const SArgs<T>& g, const Layout& L, cudaStream_t st) {
  LAUNCH_FAST(k_fast_kkt, , a, g);
}
template int fast_iter<INST_T>(const KArgs<INST_T>&, const Layout&, cudaStream_t);
template int fast_backward<INST_T>(const KArgs<INST_T>&, const BArgs<INST_T>&, const Layout&, cudaStream_t);
template int fast_kkt<INST_T>(const KArgs<INST_T>&, const SArgs<INST_T>&, const Layout&, cudaStream_t);
#else  // generic kernels (+ dispatch to the fast launchers, which live in their own translation unit)
#if INST_SMEM
#define FAST_OR(FN, ...) if (L.fast) return FN(__VA_ARGS__, L, st)
#else
#define FAST_OR(FN, ...)
#endif

template <typename T, bool SMEM> int launch_iter(const KArgs<T>& a, const Layout& L, cudaStream_t st) { FAST_OR(fast_iter<T>, a); LAUNCH_NT(k_pdipm_iter, a); }
template <typename T, bool SMEM> int launch_backward(const KArgs<T>& a, const BArgs<T>& g, const Layout& L, cudaStream_t st) { FAST_OR(fast_backward<T>, a, g); LAUNCH_NT(k_backward, a, g); }
template <typename T, bool SMEM> int launch_kkt_solve(const KArgs<T>& a, const SArgs<T>& g, const Layout& L, cudaStream_t st) { FAST_OR(fast_kkt<T>, a, g); LAUNCH_NT(k_kkt_solve, a, g); }

template int launch_iter<INST_T, INST_SMEM>(const KArgs<INST_T>&, const Layout&, cudaStream_t);
template int launch_backward<INST_T, INST_SMEM>(const KArgs<INST_T>&, const BArgs<INST_T>&, const Layout&, cudaStream_t);
template int launch_kkt_solve<INST_T, INST_SMEM>(const KArgs<INST_T>&, const SArgs<INST_T>&, const Layout&, cudaStream_t);

#ifdef INST_PREFACTOR
template <typename T> int launch_prefactor(const KArgs<T>& a_in, const Layout& L, cudaStream_t st) {
  KArgs<T> a = a_in;
  const size_t bytes = ((size_t)2 * round4(L.n * L.ldn) + round4(L.n)) * sizeof(T);
  const size_t bytes2 = bytes + (size_t)2 * round4((L.p + L.m) * L.ldn) * sizeof(T);
  a.pre_smem = bytes2 <= 100 * 1024 ? 2 : (bytes <= kSmemResidentLimit ? 1 : 0);
  size_t dyn = a.pre_smem == 2 ? bytes2 : (a.pre_smem ? bytes : 0);
  a.pre_blocked = 0;
  if (sizeof(T) == 8 && a.pre_smem != 2 && L.n >= 48) {  // blocked tensor-core route (qp_blocked.cuh)
    a.pre_blocked = (int)dyn + 1;
    dyn += (size_t)blk_panel_elems(L.n) * sizeof(double);
  }
  const bool narrow = L.nt == 128 && (L.fast || L.m <= 64);  // the blocked route's 128-thread choice is for the iteration
  if (narrow) { auto k = k_prefactor<T, 128>; CK(ensure_smem(k, dyn)); k<<<(unsigned)a.nb, 128, dyn, st>>>(a); }
  else { auto k = k_prefactor<T, 256>; CK(ensure_smem(k, dyn)); k<<<(unsigned)a.nb, 256, dyn, st>>>(a); }
  CK(cudaGetLastError());
  return B200QP_OK;
}
template int launch_prefactor<INST_T>(const KArgs<INST_T>&, const Layout&, cudaStream_t);
#endif
#endif  // INST_FAST

}  // namespace b200qp
