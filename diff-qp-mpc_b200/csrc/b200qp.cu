// b200qp.cu -- host side of libb200qp.so: workspace layout, launch sequencing, the C ABI
// declared in include/b200qp.h.  No torch types anywhere; device pointers + a cudaStream_t.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <type_traits>

#include "../../include/b200qp.h"
#include "qp_host.cuh"

namespace b200qp {

static thread_local char g_err[512] = "";

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && v[0]) ? atoi(v) : dflt;
}
Options& options() {
  static Options o = [] {
    Options x;
    x.res = env_int("B200QP_RES", x.res);
    x.res_chunk = env_int("B200QP_RES_CH", x.res_chunk);
    x.res_spec = env_int("B200QP_RES_SPEC", x.res_spec);
    x.res_warp = env_int("B200QP_RES_WARP", x.res_warp);
    x.res_pre = env_int("B200QP_RES_PRE", x.res_pre);
    const char* mid = getenv("B200QP_MID");
    x.mid_fast = (mid && mid[0] == 'f') ? 1 : 0;
    x.blk_nt = env_int("B200QP_BLK_NT", 0);
    const char* fke = getenv("B200QP_FACTOR");
    x.factor_tile = (fke && fke[0] == 't') ? 1 : 0;
    x.force_generic = getenv("B200QP_FORCE_GENERIC") != nullptr;
    return x;
  }();
  return o;
}

int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return B200QP_ECUDA;
}
// ------------------------------------------------------------------------------------------
// After the last iteration launch: iteration count and worst best-residual -> status.
static __global__ void k_finalize(const Slot* slots, const Control* ctl, double* status, int max_iter, int lim, double eps,
                           int launches) {
  const int lane = threadIdx.x;
  const int term = eval_termination(slots, max_iter, lim, eps, lane);
  if (lane == 0) {
    const int n_iter = term >= 0 ? term + 1 : max_iter;
    const Slot* sl = slots + (n_iter - 1);
    status[0] = (double)n_iter;
    status[1] = sl->best_nan ? __longlong_as_double(0x7ff8000000000000LL) : __longlong_as_double((long long)sl->best_max);
    status[2] = (double)ctl->q_fail;
    status[3] = (double)ctl->aqa_fail;
    status[4] = (double)launches;
    status[5] = 0.0; status[6] = -1.0; status[7] = 0.0;
  }
}


template <typename T>
static void fill_args(KArgs<T>& a, const b200qp_problem_t* pr, const Layout& L, void* ws) {
  memset(&a, 0, sizeof(a));
  char* w = static_cast<char*>(ws);
  a.nb = L.nb; a.n = L.n; a.m = L.m; a.p = L.p; a.ldn = L.ldn; a.ldm = L.ldm; a.ldp = L.ldp;
  a.sQ = pr->sQ; a.sp = pr->sp; a.sG = pr->sG; a.sh = pr->sh; a.sA = pr->sA; a.sb = pr->sb;
  a.Qi = (T*)(w + L.oQi); a.BQi = (T*)(w + L.oBQi); a.R = (T*)(w + L.oR); a.V = (T*)(w + L.oV);
  a.UA = (T*)(w + L.oUA); a.pinvA = (T*)(w + L.opinvA); a.F = (T*)(w + L.oF); a.pinvF = (T*)(w + L.opinvF);
  a.Tscr = (T*)(w + L.oT); a.pinvTscr = (T*)(w + L.opinvT);
  a.sQi = L.sQi; a.sBQi = L.sBQi; a.sR = L.sR; a.sV = L.sV; a.sUA = L.sUA; a.sF = L.sF; a.sT = L.sT;
  a.x = (T*)(w + L.ox); a.s = (T*)(w + L.os); a.z = (T*)(w + L.oz); a.y = (T*)(w + L.oy);
  a.dx = (T*)(w + L.odx); a.ds = (T*)(w + L.ods); a.dz = (T*)(w + L.odz); a.dy = (T*)(w + L.ody);
  a.rmu = (T*)(w + L.ormu);
  a.flags = (int*)(w + L.oflags);
  a.best_resid = (double*)(w + L.obest);
  a.slots = (Slot*)(w + L.oslots);
  a.ctl = (Control*)(w + L.octl);
  a.max_iter = pr->max_iter; a.lim = pr->not_improved_lim; a.eps = pr->eps;
  a.dense = (pr->flags & B200QP_FLAG_DENSE) ? 1 : 0; a.reg = a.dense ? pr->kkt_reg : 0.0;
  if (L.fast) fast_offsets(L.n, L.m, L.p, L.ldn, L.ldm, L.ldp, L.nt, L.mpad, L.fk, a.fso);
  a.rtile_mpad = L.fast ? L.mpad : 0; a.rtile_nt = L.fk ? -1 : (L.fast ? L.nt : 0);
}

// ---------------------------------------------------------------------------- launch profiling
// Optional CUDA-event bracketing of every kernel launch of forward/backward (bench.py's roofline
// leg).  Events are recorded on the caller's stream, so they see exactly the kernels timed.
struct Prof {
  bool on = false;
  int n = 0;
  static constexpr int kCap = 256;
  cudaEvent_t ev[kCap + 1];
  int kind[kCap];  // 0 prefactor, 1 init, 2 iteration, 3 finalize, 4 backward
  bool made = false;
};
static Prof g_prof;

static void prof_begin(cudaStream_t st) {
  if (!g_prof.on) return;
  if (!g_prof.made) {
    for (int i = 0; i <= Prof::kCap; i++) cudaEventCreate(&g_prof.ev[i]);
    g_prof.made = true;
  }
  g_prof.n = 0;
  cudaEventRecord(g_prof.ev[0], st);
}
static void prof_mark(int kind, cudaStream_t st) {
  if (!g_prof.on || g_prof.n >= Prof::kCap) return;
  g_prof.kind[g_prof.n] = kind;
  g_prof.n++;
  cudaEventRecord(g_prof.ev[g_prof.n], st);
}

template <typename T>
static int run_prefactor(KArgs<T>& a, const Layout& L, cudaStream_t st) {
  if constexpr (std::is_same<T, double>::value) {
    if (L.wpre) return res_prefactor(a, L, st);
  }
  return launch_prefactor<T>(a, L, st);
}

#define DISPATCH_KERNEL(FN, T, L, ...)                                   \
  do {                                                                   \
    int rc_ = (L).smem ? FN<T, true>(__VA_ARGS__, (L), st) : FN<T, false>(__VA_ARGS__, (L), st); \
    if (rc_) return rc_;                                                 \
  } while (0)

template <typename T>
static int forward_t(const b200qp_problem_t* pr, const Layout& L, const void* Q, const void* p, const void* G,
                     const void* h, const void* A, const void* b, void* zhat, void* lams, void* nus, void* slacks,
                     void* ws, double* status, cudaStream_t st, bool prefactored = false,
                     int phase = B200QP_PHASE_ALL, const void* cb_cg = nullptr, const void* cb_ry = nullptr) {
  KArgs<T> a;
  fill_args(a, pr, L, ws);
  a.cb_cg = (const T*)cb_cg; a.cb_ry = (const T*)cb_ry;
  a.Q = (const T*)Q; a.pv = (const T*)p; a.G = (const T*)G; a.h = (const T*)h;
  a.A = (const T*)(A ? A : G); a.b = (const T*)(b ? b : h);
  a.bx = (T*)zhat; a.bz = (T*)lams; a.bs = (T*)slacks; a.by = (T*)nus;
  a.status = status;
  int launches = 0;
  const bool all = phase == B200QP_PHASE_ALL;
  if constexpr (std::is_same<T, double>::value) {
    if (L.res && all) {
      // resident route (qp_resident.cuh): ceil(max_iter / chunk) launches + one repair launch + reduce/select
      char* w = static_cast<char*>(ws);
      RArgs ra;
      ra.hist = (double*)(w + L.ohist); ra.rec = (double*)(w + L.orec); ra.pst = (int*)(w + L.opst);
      ra.hs = res_hs(L.n, L.m);
      ra.off = res_off(L.n, L.m, L.mpad);
      prof_begin(st);
      if (!prefactored) {
        CK(cudaMemsetAsync(a.slots, 0, sizeof(Slot) * B200QP_MAX_ITER_CAP + sizeof(Control), st));
        int rc = run_prefactor(a, L, st);
        if (rc) return rc;
      }
      CK(cudaMemsetAsync(ra.pst, 0, sizeof(int) * kPst * (size_t)L.nb, st));
      prof_mark(0, st);
      launches = 2;
      for (int e = L.res_chunk;; e += L.res_chunk) {
        ra.it_end = e < pr->max_iter ? e : pr->max_iter;
        int rc = res_chunk(a, ra, L, st);
        if (rc) return rc;
        prof_mark(5, st);
        launches++;
        if (ra.it_end == pr->max_iter) break;
      }
      {  // repair launch: problems whose last guesses were wrong redo their tail, everything else exits at once
        int rc = res_chunk(a, ra, L, st);
        if (rc) return rc;
        prof_mark(5, st);
        launches++;
      }
      int rc = res_finish(a, ra, status, launches + 2, st);
      if (rc) return rc;
      prof_mark(6, st);
      return B200QP_OK;
    }
  }
  if (all || phase == B200QP_PHASE_BEGIN) {
    prof_begin(st);
    if (!prefactored) {  // the host-buffer path pre-factors chunk by chunk while the inputs arrive
      CK(cudaMemsetAsync(a.slots, 0, sizeof(Slot) * B200QP_MAX_ITER_CAP + sizeof(Control), st));
      int rc = run_prefactor(a, L, st);
      if (rc) return rc;
    }
    prof_mark(0, st);
    launches++;
    a.iter = -1;
    DISPATCH_KERNEL(launch_iter, T, L, a);
    prof_mark(1, st);
    launches++;
  }
  for (int it = 0; it < pr->max_iter; it++) {
    if (!all && phase != it) continue;
    a.iter = it;
    DISPATCH_KERNEL(launch_iter, T, L, a);
    prof_mark(2, st);
    launches++;
  }
  if (all || phase == B200QP_PHASE_END) {
    launches = pr->max_iter + 3;
    k_finalize<<<1, 32, 0, st>>>(a.slots, a.ctl, status, pr->max_iter, pr->not_improved_lim, pr->eps, launches);
    CK(cudaGetLastError());
    prof_mark(3, st);
  }
  return B200QP_OK;
}

template <typename T>
static int backward_t(const b200qp_problem_t* pr, const Layout& L, const void* zhat, const void* lams, const void* nus,
                      const void* slacks, const void* gz, void* dQ, void* dp, void* dG, void* dh, void* dA, void* db,
                      void* ws, cudaStream_t st, int lo = 0, int cnt = -1) {
  KArgs<T> a;
  fill_args(a, pr, L, ws);
  if (cnt >= 0) { a.prob0 = lo; a.nb = cnt; }  // a chunk of the batch (host-buffer path)
  BArgs<T> g;
  g.zhat = (const T*)zhat; g.lams = (const T*)lams; g.nus = (const T*)(nus ? nus : zhat); g.slacks = (const T*)slacks;
  g.gz = (const T*)gz;
  g.dQ = (T*)dQ; g.dp = (T*)dp; g.dG = (T*)dG; g.dh = (T*)dh; g.dA = (T*)dA; g.db = (T*)db;
  const bool chained = g_prof.on && g_prof.n > 0 && (g_prof.kind[g_prof.n - 1] == 3 || g_prof.kind[g_prof.n - 1] == 6);
  if (g_prof.on && !chained) prof_begin(st);
  if (chained) cudaEventRecord(g_prof.ev[g_prof.n], st);  // restart the bracket after host-side gaps
  bool done = false;
  if constexpr (std::is_same<T, double>::value) {
    if (L.res && L.res_warp) {  // the shapes of the resident route: one warp per QP (qp_wres.cuh)
      int rc = res_backward(a, g, L, st);
      if (rc) return rc;
      done = true;
    }
  }
  if (!done) DISPATCH_KERNEL(launch_backward, T, L, a, g);
  prof_mark(4, st);
  return B200QP_OK;
}

template <typename T>
static int kkt_solve_t(const b200qp_problem_t* pr, const Layout& L, int prefactor, const void* Q, const void* G,
                       const void* A, const void* d, const void* rx, const void* rs, const void* rz, const void* ry,
                       void* dx, void* ds, void* dz, void* dy, void* ws, cudaStream_t st) {
  KArgs<T> a;
  fill_args(a, pr, L, ws);
  a.Q = (const T*)Q; a.G = (const T*)G; a.A = (const T*)(A ? A : G);
  if (prefactor) {
    CK(cudaMemsetAsync(a.slots, 0, sizeof(Slot) * B200QP_MAX_ITER_CAP + sizeof(Control), st));
    int rc = run_prefactor(a, L, st);
    if (rc) return rc;
  }
  SArgs<T> g;
  g.d = (const T*)d; g.rx = (const T*)rx; g.rs = (const T*)rs; g.rz = (const T*)rz; g.ry = (const T*)ry;
  g.dx = (T*)dx; g.ds = (T*)ds; g.dz = (T*)dz; g.dy = (T*)dy;
  DISPATCH_KERNEL(launch_kkt_solve, T, L, a, g);
  return B200QP_OK;
}

static __global__ void k_prefactor_status(const Control* ctl, double* status) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < B200QP_STATUS_DOUBLES; i++) status[i] = 0.0;
    status[B200QP_ST_Q_FAIL] = (double)ctl->q_fail;
    status[B200QP_ST_AQA_FAIL] = (double)ctl->aqa_fail;
    status[B200QP_ST_LAUNCHES] = 2.0;
  }
}

template <typename T>
static int prefactor_t(const b200qp_problem_t* pr, const Layout& L, const void* Q, const void* G, const void* A, void* ws,
                       double* status, cudaStream_t st) {
  KArgs<T> a;
  fill_args(a, pr, L, ws);
  a.Q = (const T*)Q; a.G = (const T*)G; a.A = (const T*)(A ? A : G);
  CK(cudaMemsetAsync(a.slots, 0, sizeof(Slot) * B200QP_MAX_ITER_CAP + sizeof(Control), st));
  int rc = run_prefactor(a, L, st);
  if (rc) return rc;
  if (status) {
    k_prefactor_status<<<1, 32, 0, st>>>(a.ctl, status);
    CK(cudaGetLastError());
  }
  return B200QP_OK;
}

// Pre-factorisation of problems [lo, lo + cnt) only (host-buffer path: overlaps the H2D copies).
template <typename T>
static int prefactor_range_t(const b200qp_problem_t* pr, const Layout& L, const void* Q, const void* G, const void* A,
                             void* ws, int lo, int cnt, cudaStream_t st) {
  KArgs<T> a;
  fill_args(a, pr, L, ws);
  a.Q = (const T*)Q; a.G = (const T*)G; a.A = (const T*)(A ? A : G);
  a.prob0 = lo; a.nb = cnt;
  return run_prefactor(a, L, st);
}

template <typename T>
static int cb_step_t(const b200qp_problem_t* pr, const Layout& L, int it, void* x_out, void* ws, cudaStream_t st) {
  KArgs<T> a;
  fill_args(a, pr, L, ws);
  a.iter = it;
  k_cb_step<T><<<(unsigned)L.nb, 128, 0, st>>>(a, (T*)x_out);
  CK(cudaGetLastError());
  return B200QP_OK;
}

// ---------------------------------------------------------------------------- host-buffer path
// B200QP_HOST_SLOTS device arenas so that consecutive host-buffer solves overlap: the inputs of solve k+1
// cross the PCIe bus (stream g_cin) and the gradients of solve k-1 go back (stream g_cout) while the kernels
// of solve k run (stream g_st).  A job is three stages (H2D, kernels, D2H) of similar length at the headline
// shape, so the steady state needs THREE jobs in flight: with two slots the period is (H2D + kernels + D2H) / 2
// (measured 22.9 ms against 16.5 ms of kernels), with three it is the longest stage.  b200qp_solve_host uses
// slot 0 synchronously.
struct Arena {
  char* base = nullptr;
  size_t cap = 0;
  static constexpr int kMaxChunks = 8;
  cudaEvent_t ev_in[kMaxChunks], ev_out[kMaxChunks], ev_misc[2], ev_done;
  bool events = false;
  bool busy = false;
  // the job in flight, kept so that b200qp_solve_host_wait can repeat it on the exact route (B200QP_ST_SPEC_FAIL)
  b200qp_problem_t job_prob;
  const void* job_in[7];
  void* job_out[10];
  double* job_status = nullptr;
};
static Arena g_arenas[B200QP_HOST_SLOTS];
static cudaStream_t g_st = nullptr, g_cin = nullptr, g_cout = nullptr;
static std::mutex g_host_mu;

static int g_host_device = -1;  // the host-buffer pipeline (streams, arenas) belongs to the device of its first use
static int arena_reserve(Arena& A, size_t bytes) {
  int dev = -1;
  CK(cudaGetDevice(&dev));
  if (g_host_device < 0) g_host_device = dev;
  if (dev != g_host_device) {  // one process per GPU: refuse to mix devices instead of corrupting them
    snprintf(g_err, sizeof(g_err), "b200qp_solve_host*: first used on device %d, now called with device %d current", g_host_device, dev);
    return B200QP_EINVAL;
  }
  if (!g_st) CK(cudaStreamCreateWithFlags(&g_st, cudaStreamNonBlocking));
  if (!g_cin) CK(cudaStreamCreateWithFlags(&g_cin, cudaStreamNonBlocking));
  if (!g_cout) CK(cudaStreamCreateWithFlags(&g_cout, cudaStreamNonBlocking));
  if (!A.events) {
    for (int i = 0; i < Arena::kMaxChunks; i++) {
      CK(cudaEventCreateWithFlags(&A.ev_in[i], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&A.ev_out[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < 2; i++) CK(cudaEventCreateWithFlags(&A.ev_misc[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&A.ev_done, cudaEventDisableTiming));
    A.events = true;
  }
  if (bytes <= A.cap) return B200QP_OK;
  if (A.base) {
    CK(cudaDeviceSynchronize());  // the other slot may still be running: keep it simple and safe
    CK(cudaFree(A.base));
  }
  A.base = nullptr;
  A.cap = 0;
  CK(cudaMalloc(&A.base, bytes));
  A.cap = bytes;
  return B200QP_OK;
}

}  // namespace b200qp

using namespace b200qp;

extern "C" {

size_t b200qp_workspace_bytes(const b200qp_problem_t* prob) {
  Layout L;
  if (make_layout(prob, L) != B200QP_OK) return 0;
  return L.total;
}

size_t b200qp_slot_offset(const b200qp_problem_t* prob) {
  Layout L;
  if (make_layout(prob, L) != B200QP_OK) return 0;
  return L.oslots;
}

int b200qp_forward_phase(const b200qp_problem_t* prob, int phase, const void* Q, const void* p, const void* G,
                         const void* h, const void* A, const void* b, void* zhat, void* lams, void* nus, void* slacks,
                         void* workspace, double* status, b200qp_stream_t stream) {
  Layout L;
  int rc = make_layout(prob, L);
  if (rc) return rc;
  if (!Q || !p || !G || !h || !zhat || !lams || !slacks || !workspace || !status) return B200QP_EINVAL;
  if (prob->neq > 0 && (!A || !b || !nus)) return B200QP_EINVAL;
  if (phase != B200QP_PHASE_ALL && phase != B200QP_PHASE_BEGIN && phase != B200QP_PHASE_END &&
      (phase < 0 || phase >= prob->max_iter))
    return B200QP_EINVAL;
  if (phase != B200QP_PHASE_ALL) L.res = false;  // the phased (exact-sharded) mode is one launch per iteration by definition
  cudaStream_t st = (cudaStream_t)stream;
  if (prob->dtype == B200QP_F64)
    return forward_t<double>(prob, L, Q, p, G, h, A, b, zhat, lams, nus, slacks, workspace, status, st, false, phase);
  return forward_t<float>(prob, L, Q, p, G, h, A, b, zhat, lams, nus, slacks, workspace, status, st, false, phase);
}

int b200qp_forward_phase_cb(const b200qp_problem_t* prob, int phase, const void* Q, const void* p, const void* G,
                            const void* h, const void* A, const void* b, void* zhat, void* lams, void* nus, void* slacks,
                            void* workspace, double* status, const void* cost_grad_x, const void* dyn_res_x,
                            b200qp_stream_t stream) {
  Layout L;
  int rc = make_layout(prob, L);
  if (rc) return rc;
  if (!Q || !p || !G || !h || !zhat || !lams || !slacks || !workspace || !status) return B200QP_EINVAL;
  if (prob->neq > 0 && (!A || !b || !nus)) return B200QP_EINVAL;
  if (phase < 0 || phase >= prob->max_iter) return B200QP_EINVAL;  // callbacks only enter the iterations proper
  L.res = false;
  cudaStream_t st = (cudaStream_t)stream;
  if (prob->dtype == B200QP_F64)
    return forward_t<double>(prob, L, Q, p, G, h, A, b, zhat, lams, nus, slacks, workspace, status, st, false, phase,
                             cost_grad_x, dyn_res_x);
  return forward_t<float>(prob, L, Q, p, G, h, A, b, zhat, lams, nus, slacks, workspace, status, st, false, phase,
                          cost_grad_x, dyn_res_x);
}

int b200qp_forward_cb_step(const b200qp_problem_t* prob, int it, void* x_out, void* workspace, b200qp_stream_t stream) {
  Layout L;
  int rc = make_layout(prob, L);
  if (rc) return rc;
  if (!x_out || !workspace || it < 0 || it >= prob->max_iter) return B200QP_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (prob->dtype == B200QP_F64) return cb_step_t<double>(prob, L, it, x_out, workspace, st);
  return cb_step_t<float>(prob, L, it, x_out, workspace, st);
}

int b200qp_forward(const b200qp_problem_t* prob, const void* Q, const void* p, const void* G, const void* h,
                   const void* A, const void* b, void* zhat, void* lams, void* nus, void* slacks, void* workspace,
                   double* status, b200qp_stream_t stream) {
  Layout L;
  int rc = make_layout(prob, L);
  if (rc) return rc;
  if (!Q || !p || !G || !h || !zhat || !lams || !slacks || !workspace || !status) return B200QP_EINVAL;
  if (prob->neq > 0 && (!A || !b || !nus)) return B200QP_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (prob->dtype == B200QP_F64)
    return forward_t<double>(prob, L, Q, p, G, h, A, b, zhat, lams, nus, slacks, workspace, status, st);
  return forward_t<float>(prob, L, Q, p, G, h, A, b, zhat, lams, nus, slacks, workspace, status, st);
}

int b200qp_backward(const b200qp_problem_t* prob, const void* zhat, const void* lams, const void* nus,
                    const void* slacks, const void* dl_dzhat, void* dQ, void* dp, void* dG, void* dh, void* dA,
                    void* db, void* workspace, b200qp_stream_t stream) {
  Layout L;
  int rc = make_layout(prob, L);
  if (rc) return rc;
  if (!zhat || !lams || !slacks || !dl_dzhat || !dQ || !dp || !dG || !dh || !workspace) return B200QP_EINVAL;
  if (prob->neq > 0 && (!nus || !dA || !db)) return B200QP_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (prob->dtype == B200QP_F64)
    return backward_t<double>(prob, L, zhat, lams, nus, slacks, dl_dzhat, dQ, dp, dG, dh, dA, db, workspace, st);
  return backward_t<float>(prob, L, zhat, lams, nus, slacks, dl_dzhat, dQ, dp, dG, dh, dA, db, workspace, st);
}

int b200qp_prefactor(const b200qp_problem_t* prob, const void* Q, const void* G, const void* A, void* workspace,
                     double* status, b200qp_stream_t stream) {
  Layout L;
  int rc = make_layout(prob, L);
  if (rc) return rc;
  if (!Q || !G || !workspace || (prob->neq > 0 && !A)) return B200QP_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (prob->dtype == B200QP_F64) return prefactor_t<double>(prob, L, Q, G, A, workspace, status, st);
  return prefactor_t<float>(prob, L, Q, G, A, workspace, status, st);
}

int b200qp_kkt_solve(const b200qp_problem_t* prob, int prefactor, const void* Q, const void* G, const void* A,
                     const void* d, const void* rx, const void* rs, const void* rz, const void* ry, void* dx, void* ds,
                     void* dz, void* dy, void* workspace, b200qp_stream_t stream) {
  Layout L;
  int rc = make_layout(prob, L);
  if (rc) return rc;
  if (!d || !rx || !rs || !rz || !dx || !ds || !dz || !workspace) return B200QP_EINVAL;
  if (prefactor && (!Q || !G)) return B200QP_EINVAL;
  if (prob->neq > 0 && (!ry || !dy || (prefactor && !A))) return B200QP_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (prob->dtype == B200QP_F64)
    return kkt_solve_t<double>(prob, L, prefactor, Q, G, A, d, rx, rs, rz, ry, dx, ds, dz, dy, workspace, st);
  return kkt_solve_t<float>(prob, L, prefactor, Q, G, A, d, rx, rs, rz, ry, dx, ds, dz, dy, workspace, st);
}

int b200qp_solve_host_wait(int slot) {
  if (slot < 0 || slot >= B200QP_HOST_SLOTS) return B200QP_EINVAL;
  Arena& AR = g_arenas[slot];
  {
    std::lock_guard<std::mutex> lock(g_host_mu);
    if (!AR.busy) return B200QP_OK;
  }
  CK(cudaEventSynchronize(AR.ev_done));
  bool again = false;
  {
    std::lock_guard<std::mutex> lock(g_host_mu);
    AR.busy = false;
    again = AR.job_status && AR.job_status[B200QP_ST_SPEC_FAIL] != 0.0 && !(AR.job_prob.flags & B200QP_FLAG_EXACT);
  }
  if (again) {  // the resident route met something it never speculates: same job, one launch per iteration
    b200qp_problem_t pr = AR.job_prob;
    pr.flags |= B200QP_FLAG_EXACT;
    int rc = b200qp_solve_host_submit(slot, &pr, AR.job_in[0], AR.job_in[1], AR.job_in[2], AR.job_in[3], AR.job_in[4],
                                      AR.job_in[5], AR.job_in[6], AR.job_out[0], AR.job_out[1], AR.job_out[2], AR.job_out[3],
                                      AR.job_out[4], AR.job_out[5], AR.job_out[6], AR.job_out[7], AR.job_out[8], AR.job_out[9],
                                      AR.job_status);
    if (rc) return rc;
    return b200qp_solve_host_wait(slot);
  }
  return B200QP_OK;
}

int b200qp_solve_host_submit(int slot, const b200qp_problem_t* prob, const void* Q, const void* p, const void* G,
                             const void* h, const void* A, const void* b, const void* dl_dzhat, void* zhat, void* lams,
                             void* nus, void* slacks, void* dQ, void* dp, void* dG, void* dh, void* dA, void* db,
                             double* status) {
  if (slot < 0 || slot >= B200QP_HOST_SLOTS) return B200QP_EINVAL;
  Layout L;
  int rc = make_layout(prob, L);
  if (rc) return rc;
  if (!Q || !p || !G || !h || !zhat || !lams || !slacks || !status) return B200QP_EINVAL;
  const size_t es = L.es, nb = (size_t)L.nb, n = (size_t)L.n, m = (size_t)L.m, pe = (size_t)L.p;
  if (pe > 0 && (!A || !b || !nus)) return B200QP_EINVAL;
  const bool bwd = dl_dzhat != nullptr;
  const bool factored = (prob->flags & B200QP_FLAG_FACTORED_GRAD) != 0;
  if (bwd && ((!factored && (!dQ || !dG)) || !dp || !dh || (pe > 0 && (!dA || !db)))) return B200QP_EINVAL;
  Arena& AR = g_arenas[slot];
  {
    int rcw = b200qp_solve_host_wait(slot);  // the previous job of this slot still owns the arena (and its output buffers)
    if (rcw) return rcw;
  }
  std::lock_guard<std::mutex> lock(g_host_mu);
  AR.job_prob = *prob;
  {
    const void* in_[7] = {Q, p, G, h, A, b, dl_dzhat};
    void* out_[10] = {zhat, lams, nus, slacks, dQ, dp, dG, dh, dA, db};
    for (int i = 0; i < 7; i++) AR.job_in[i] = in_[i];
    for (int i = 0; i < 10; i++) AR.job_out[i] = out_[i];
    AR.job_status = status;
  }
  auto cnt = [&](int64_t stride, size_t per) { return (stride == 0 ? 1 : nb) * per; };
  const size_t bQ = cnt(prob->sQ, n * n) * es, bp = cnt(prob->sp, n) * es, bG = cnt(prob->sG, m * n) * es,
               bh = cnt(prob->sh, m) * es, bA = cnt(prob->sA, pe * n) * es, bb = cnt(prob->sb, pe) * es;
  const size_t bz = nb * n * es, bl = nb * m * es, bn = nb * pe * es;
  size_t off = 0;
  auto put = [&](size_t bytes) { size_t o = off; off = align_up(off + (bytes ? bytes : 16), 256); return o; };
  const size_t oQ = put(bQ), op = put(bp), oG = put(bG), oh = put(bh), oA = put(bA), ob = put(bb);
  const size_t oz = put(bz), ol = put(bl), on = put(bn), os = put(bl), ost = put(8 * sizeof(double));
  const size_t ogz = put(bwd ? bz : 0), odQ = put(bwd ? nb * n * n * es : 0), odp = put(bwd ? bz : 0),
               odG = put(bwd ? nb * m * n * es : 0), odh = put(bwd ? bl : 0), odA = put(bwd ? nb * pe * n * es : 0),
               odb = put(bwd ? bn : 0);
  const size_t ows = put(L.total);
  rc = arena_reserve(AR, off);
  if (rc) return rc;
  char* d = AR.base;
  cudaStream_t st = g_st, cs = g_cin, co = g_cout;
  // Pipeline: the batch is cut into chunks; the pre-factorisation of chunk c runs while chunk
  // c+1 is still on the PCIe bus, and the gradients of chunk c go back to the host while the
  // backward kernel works on chunk c+1.  The PDIPM loop itself needs the whole batch (its
  // termination and step fill are batch-global), so it sits between the two pipelines.
  const int nch = (int)nb >= 4096 ? Arena::kMaxChunks : 1;
  auto lo_of = [&](int c) { return (int)((size_t)c * nb / nch); };
  auto slice_h2d = [&](size_t off, const void* src, int64_t stride, size_t per, int lo, int cnt) -> cudaError_t {
    if (stride == 0 || per == 0) return cudaSuccess;  // shared parameters are copied once, below
    return cudaMemcpyAsync(d + off + (size_t)lo * per * es, (const char*)src + (size_t)lo * per * es, (size_t)cnt * per * es,
                           cudaMemcpyHostToDevice, cs);
  };
  auto shared_h2d = [&](size_t off, const void* src, int64_t stride, size_t bytes) -> cudaError_t {
    if (stride != 0 || bytes == 0) return cudaSuccess;
    return cudaMemcpyAsync(d + off, src, bytes, cudaMemcpyHostToDevice, cs);
  };
  CK(shared_h2d(oQ, Q, prob->sQ, bQ)); CK(shared_h2d(op, p, prob->sp, bp)); CK(shared_h2d(oG, G, prob->sG, bG));
  CK(shared_h2d(oh, h, prob->sh, bh));
  if (pe > 0) { CK(shared_h2d(oA, A, prob->sA, bA)); CK(shared_h2d(ob, b, prob->sb, bb)); }
  {
    KArgs<double> a0;  // only for the slot / control block address (same for both dtypes)
    fill_args(a0, prob, L, d + ows);
    CK(cudaMemsetAsync(a0.slots, 0, sizeof(Slot) * B200QP_MAX_ITER_CAP + sizeof(Control), st));
  }
  for (int c = 0; c < nch; c++) {
    const int lo = lo_of(c), cnt = lo_of(c + 1) - lo;
    CK(slice_h2d(oQ, Q, prob->sQ, n * n, lo, cnt));
    CK(slice_h2d(oG, G, prob->sG, m * n, lo, cnt));
    if (pe > 0) CK(slice_h2d(oA, A, prob->sA, pe * n, lo, cnt));
    CK(cudaEventRecord(AR.ev_in[c], cs));
    CK(cudaStreamWaitEvent(st, AR.ev_in[c], 0));
    rc = prob->dtype == B200QP_F64
             ? prefactor_range_t<double>(prob, L, d + oQ, d + oG, pe ? d + oA : nullptr, d + ows, lo, cnt, st)
             : prefactor_range_t<float>(prob, L, d + oQ, d + oG, pe ? d + oA : nullptr, d + ows, lo, cnt, st);
    if (rc) return rc;
  }
  // the vectors are not needed by the pre-factorisation: they travel behind the matrices
  CK(slice_h2d(op, p, prob->sp, n, 0, (int)nb));
  CK(slice_h2d(oh, h, prob->sh, m, 0, (int)nb));
  if (pe > 0) CK(slice_h2d(ob, b, prob->sb, pe, 0, (int)nb));
  if (bwd) CK(cudaMemcpyAsync(d + ogz, dl_dzhat, bz, cudaMemcpyHostToDevice, cs));
  CK(cudaEventRecord(AR.ev_misc[0], cs));
  CK(cudaStreamWaitEvent(st, AR.ev_misc[0], 0));
  void* dA_ = pe ? (void*)(d + oA) : nullptr;
  void* db_ = pe ? (void*)(d + ob) : nullptr;
  rc = prob->dtype == B200QP_F64
           ? forward_t<double>(prob, L, d + oQ, d + op, d + oG, d + oh, dA_, db_, d + oz, d + ol, d + on, d + os, d + ows,
                               (double*)(d + ost), st, true)
           : forward_t<float>(prob, L, d + oQ, d + op, d + oG, d + oh, dA_, db_, d + oz, d + ol, d + on, d + os, d + ows,
                              (double*)(d + ost), st, true);
  if (rc) return rc;
  CK(cudaEventRecord(AR.ev_misc[1], st));
  CK(cudaStreamWaitEvent(co, AR.ev_misc[1], 0));
  CK(cudaMemcpyAsync(zhat, d + oz, bz, cudaMemcpyDeviceToHost, co));
  CK(cudaMemcpyAsync(lams, d + ol, bl, cudaMemcpyDeviceToHost, co));
  CK(cudaMemcpyAsync(slacks, d + os, bl, cudaMemcpyDeviceToHost, co));
  if (pe > 0) CK(cudaMemcpyAsync(nus, d + on, bn, cudaMemcpyDeviceToHost, co));
  CK(cudaMemcpyAsync(status, d + ost, 8 * sizeof(double), cudaMemcpyDeviceToHost, co));
  if (bwd) {
    auto slice_d2h = [&](void* dst, size_t off, size_t per, int lo, int cnt) -> cudaError_t {
      if (per == 0) return cudaSuccess;
      return cudaMemcpyAsync((char*)dst + (size_t)lo * per * es, d + off + (size_t)lo * per * es, (size_t)cnt * per * es,
                             cudaMemcpyDeviceToHost, co);
    };
    // factored gradients: the warp-per-QP backward kernel skips the outer products altogether; the other backward
    // kernels still write them to the arena, they are just not copied
    const bool skipQG = factored && prob->dtype == B200QP_F64 && L.res && L.res_warp;
    for (int c = 0; c < nch; c++) {
      const int lo = lo_of(c), cnt = lo_of(c + 1) - lo;
      rc = prob->dtype == B200QP_F64
               ? backward_t<double>(prob, L, d + oz, d + ol, d + on, d + os, d + ogz, skipQG ? nullptr : d + odQ, d + odp,
                                    skipQG ? nullptr : d + odG, d + odh, d + odA, d + odb, d + ows, st, lo, cnt)
               : backward_t<float>(prob, L, d + oz, d + ol, d + on, d + os, d + ogz, d + odQ, d + odp, d + odG, d + odh,
                                   d + odA, d + odb, d + ows, st, lo, cnt);
      if (rc) return rc;
      CK(cudaEventRecord(AR.ev_out[c], st));
      CK(cudaStreamWaitEvent(co, AR.ev_out[c], 0));
      if (!factored) CK(slice_d2h(dQ, odQ, n * n, lo, cnt));
      CK(slice_d2h(dp, odp, n, lo, cnt));
      if (!factored) CK(slice_d2h(dG, odG, m * n, lo, cnt));
      CK(slice_d2h(dh, odh, m, lo, cnt));
      if (pe > 0) { CK(slice_d2h(dA, odA, pe * n, lo, cnt)); CK(slice_d2h(db, odb, pe, lo, cnt)); }
    }
  }
  CK(cudaEventRecord(AR.ev_done, co));
  AR.busy = true;
  return B200QP_OK;
}

int b200qp_solve_host(const b200qp_problem_t* prob, const void* Q, const void* p, const void* G, const void* h,
                      const void* A, const void* b, const void* dl_dzhat, void* zhat, void* lams, void* nus,
                      void* slacks, void* dQ, void* dp, void* dG, void* dh, void* dA, void* db, double* status) {
  int rc = b200qp_solve_host_submit(0, prob, Q, p, G, h, A, b, dl_dzhat, zhat, lams, nus, slacks, dQ, dp, dG, dh, dA, db,
                                    status);
  if (rc) return rc;
  return b200qp_solve_host_wait(0);
}

void b200qp_profile_enable(int on) { g_prof.on = on != 0; g_prof.n = 0; }

int b200qp_profile_read(float* ms, int* kind, int cap) {
  if (!g_prof.on || g_prof.n == 0) return 0;
  cudaEventSynchronize(g_prof.ev[g_prof.n]);
  int n = g_prof.n < cap ? g_prof.n : cap;
  for (int i = 0; i < n; i++) {
    float t = 0.f;
    cudaEventElapsedTime(&t, g_prof.ev[i], g_prof.ev[i + 1]);
    ms[i] = t;
    kind[i] = g_prof.kind[i];
  }
  return n;
}

int b200qp_set_option(const char* name, int value) {
  if (!name) return B200QP_EINVAL;
  Options& o = options();
  if (!strcmp(name, "res")) o.res = value;
  else if (!strcmp(name, "res_ch")) o.res_chunk = value;
  else if (!strcmp(name, "res_spec")) o.res_spec = value;
  else if (!strcmp(name, "res_warp")) o.res_warp = value;
  else if (!strcmp(name, "res_pre")) o.res_pre = value;
  else if (!strcmp(name, "mid_fast")) o.mid_fast = value;
  else if (!strcmp(name, "blk_nt")) o.blk_nt = value;
  else if (!strcmp(name, "factor_tile")) o.factor_tile = value;
  else if (!strcmp(name, "force_generic")) o.force_generic = value;
  else return B200QP_EINVAL;
  return B200QP_OK;
}

const char* b200qp_last_cuda_error(void) { return g_err; }

const char* b200qp_version(void) { return "b200qp 0.2 sm_100a"; }

}  // extern "C"
