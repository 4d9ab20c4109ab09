// mpc_dynamics.cuh -- batched dynamics of the MPC environments, written ONCE per environment as a
// function template over the scalar type: instantiated with a plain real it is the step
// x+ = f(x,u); instantiated with forward-mode dual numbers it yields the Jacobians [df/dx df/du]
// that the reference obtains by tiling the batch nx times and calling torch.autograd.grad
// (deqmpc/envs.py:68-82, deqmpc/rex_quadrotor.py:132-146).
//
//   Pendulum    deqmpc/envs.py:5-48        semi-implicit Euler, theta measured from upright
//   Integrator  deqmpc/envs.py:182-214     double integrator, semi-implicit Euler
//   PendulumDx  qpth/env_dx/pendulum.py:49-84   (cos, sin, thdot) state, clamped torque
//   CartpoleDx  qpth/env_dx/cartpole.py:63-96   (x, dx, cos, sin, dth) state, clamped force
//   RexQuadrotor deqmpc/rex_quadrotor.py:7-146  12-state MRP rigid body, RK4, act_scale = 100
//   Pendulum1L / Cartpole1L / Cartpole2L   deqmpc/my_envs/{pendulum1l,cartpole1l,cartpole1l_v2,cartpole2l}:
//       the reference ships these as CasADi-GENERATED straight-line code (src/generated_dynamics.c,
//       generated_derivatives.c, 300-10500 lines, one thread per batch row in dynamics_gpu.cu); here
//       the rigid-body model behind that code is written out (M(q) qdd = tau - bias) and integrated
//       by the same classical RK4, state (q, qd), control = force/torque on the first joint only
//       (deqmpc/my_envs/dynamics.py:27-108)
#pragma once
#include <cuda_runtime.h>

namespace b200mpc {

// ---------------------------------------------------------------------------- dual numbers
template <typename R, int N>
struct Dual {
  R v;
  R d[N];
  __device__ __forceinline__ Dual() {}
  __device__ __forceinline__ Dual(R c) : v(c) {
#pragma unroll
    for (int i = 0; i < N; i++) d[i] = R(0);
  }
};
template <typename R, int N> __device__ __forceinline__ Dual<R, N> operator+(const Dual<R, N>& a, const Dual<R, N>& b) {
  Dual<R, N> r; r.v = a.v + b.v;
#pragma unroll
  for (int i = 0; i < N; i++) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <typename R, int N> __device__ __forceinline__ Dual<R, N> operator-(const Dual<R, N>& a, const Dual<R, N>& b) {
  Dual<R, N> r; r.v = a.v - b.v;
#pragma unroll
  for (int i = 0; i < N; i++) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <typename R, int N> __device__ __forceinline__ Dual<R, N> operator-(const Dual<R, N>& a) {
  Dual<R, N> r; r.v = -a.v;
#pragma unroll
  for (int i = 0; i < N; i++) r.d[i] = -a.d[i];
  return r;
}
template <typename R, int N> __device__ __forceinline__ Dual<R, N> operator*(const Dual<R, N>& a, const Dual<R, N>& b) {
  Dual<R, N> r; r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < N; i++) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
template <typename R, int N> __device__ __forceinline__ Dual<R, N> operator/(const Dual<R, N>& a, const Dual<R, N>& b) {
  Dual<R, N> r; const R ib = R(1) / b.v; r.v = a.v * ib;
#pragma unroll
  for (int i = 0; i < N; i++) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
  return r;
}
template <typename R, int N> __device__ __forceinline__ Dual<R, N> operator+(const Dual<R, N>& a, R b) { Dual<R, N> r = a; r.v += b; return r; }
template <typename R, int N> __device__ __forceinline__ Dual<R, N> operator+(R b, const Dual<R, N>& a) { return a + b; }
template <typename R, int N> __device__ __forceinline__ Dual<R, N> operator-(const Dual<R, N>& a, R b) { Dual<R, N> r = a; r.v -= b; return r; }
template <typename R, int N> __device__ __forceinline__ Dual<R, N> operator-(R b, const Dual<R, N>& a) { return (-a) + b; }
template <typename R, int N> __device__ __forceinline__ Dual<R, N> operator*(const Dual<R, N>& a, R b) {
  Dual<R, N> r; r.v = a.v * b;
#pragma unroll
  for (int i = 0; i < N; i++) r.d[i] = a.d[i] * b;
  return r;
}
template <typename R, int N> __device__ __forceinline__ Dual<R, N> operator*(R b, const Dual<R, N>& a) { return a * b; }
template <typename R, int N> __device__ __forceinline__ Dual<R, N> operator/(const Dual<R, N>& a, R b) { return a * (R(1) / b); }
template <typename R, int N> __device__ __forceinline__ Dual<R, N> operator/(R a, const Dual<R, N>& b) { return Dual<R, N>(a) / b; }

__device__ __forceinline__ void m_sincos(double x, double& s, double& c) { sincos(x, &s, &c); }
__device__ __forceinline__ void m_sincos(float x, float& s, float& c) { sincosf(x, &s, &c); }
__device__ __forceinline__ double m_sin(double x) { return sin(x); }
__device__ __forceinline__ float m_sin(float x) { return sinf(x); }
__device__ __forceinline__ double m_cos(double x) { return cos(x); }
__device__ __forceinline__ float m_cos(float x) { return cosf(x); }
template <typename R, int N> __device__ __forceinline__ Dual<R, N> m_sin(const Dual<R, N>& a) {
  Dual<R, N> r; R c; m_sincos(a.v, r.v, c);
#pragma unroll
  for (int i = 0; i < N; i++) r.d[i] = c * a.d[i];
  return r;
}
template <typename R, int N> __device__ __forceinline__ Dual<R, N> m_cos(const Dual<R, N>& a) {
  Dual<R, N> r; R sp; m_sincos(a.v, sp, r.v); const R s = -sp;
#pragma unroll
  for (int i = 0; i < N; i++) r.d[i] = s * a.d[i];
  return r;
}
// sine and cosine of the same angle in one call: one argument reduction instead of two (four under duals, where each of
// m_sin / m_cos needs both).  ncu on k_al_solve<Cartpole1L>: 20 % of the samples were in sin / cos.
template <typename R, int N> __device__ __forceinline__ void m_sincos(const Dual<R, N>& a, Dual<R, N>& s, Dual<R, N>& c) {
  R sv, cv;
  m_sincos(a.v, sv, cv);
  s.v = sv; c.v = cv;
  const R ns = -sv;
#pragma unroll
  for (int i = 0; i < N; i++) { s.d[i] = cv * a.d[i]; c.d[i] = ns * a.d[i]; }
}
__device__ __forceinline__ double m_atan2(double y, double x) { return atan2(y, x); }
__device__ __forceinline__ float m_atan2(float y, float x) { return atan2f(y, x); }
template <typename R, int N> __device__ __forceinline__ Dual<R, N> m_atan2(const Dual<R, N>& y, const Dual<R, N>& x) {
  Dual<R, N> r; r.v = m_atan2(y.v, x.v);
  const R inv = R(1) / (x.v * x.v + y.v * y.v);
#pragma unroll
  for (int i = 0; i < N; i++) r.d[i] = (x.v * y.d[i] - y.v * x.d[i]) * inv;
  return r;
}
// clamp with the sub-gradient torch.clamp uses (1 inside and AT the bounds, 0 outside)
__device__ __forceinline__ double m_clamp(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }
__device__ __forceinline__ float m_clamp(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }
template <typename R, int N> __device__ __forceinline__ Dual<R, N> m_clamp(const Dual<R, N>& a, R lo, R hi) {
  if (a.v < lo) return Dual<R, N>(lo);
  if (a.v > hi) return Dual<R, N>(hi);
  return a;
}

// sign with zero derivative (torch.sign)
__device__ __forceinline__ double m_sign(double x) { return x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : 0.0); }
__device__ __forceinline__ float m_sign(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }
template <typename R, int N> __device__ __forceinline__ R m_sign(const Dual<R, N>& a) { return m_sign(a.v); }

template <typename S> struct real_of { typedef S type; };
template <typename R, int N> struct real_of<Dual<R, N>> { typedef R type; };

// ---------------------------------------------------------------------------- environments
constexpr int ENV_PENDULUM = 0, ENV_INTEGRATOR = 1, ENV_PENDULUM_DX = 2, ENV_CARTPOLE_DX = 3, ENV_REX_QUADROTOR = 4;
constexpr int ENV_PENDULUM1L = 5, ENV_CARTPOLE1L = 6, ENV_CARTPOLE2L = 7;
constexpr int ENV_CARTPOLE1L_V1 = 8, ENV_CARTPOLE2L_V1 = 9;  // deqmpc/envs_v1.py RK4 cart-poles
constexpr int MAX_PARAMS = 64;

struct DynParams { double v[MAX_PARAMS]; };

// params: dt, g, m, l
struct Pendulum {
  static constexpr int NX = 2, NU = 1;
  template <typename S>
  __device__ static __forceinline__ void step(const DynParams& P, const S* x, const S* u, S* xn) {
    typedef typename real_of<S>::type R;
    const R dt = (R)P.v[0], g = (R)P.v[1], m = (R)P.v[2], l = (R)P.v[3];
    const S acc = (u[0] + (m * g * l) * m_sin(x[0])) / (m * l * l);
    const S nthdot = x[1] + acc * dt;
    xn[0] = x[0] + nthdot * dt;
    xn[1] = nthdot;
  }
};

// params: dt ; nx = 2, nu = 1 (the reference's default IntegratorDynamics)
struct Integrator {
  static constexpr int NX = 2, NU = 1;
  template <typename S>
  __device__ static __forceinline__ void step(const DynParams& P, const S* x, const S* u, S* xn) {
    typedef typename real_of<S>::type R;
    const R dt = (R)P.v[0];
    const S vel = x[1] + u[0] * dt;
    xn[0] = x[0] + vel * dt;
    xn[1] = vel;
  }
};

// params: dt, g, m, l, max_torque   (qpth/env_dx/pendulum.py:49-84, the `simple` model)
struct PendulumDx {
  static constexpr int NX = 3, NU = 1;
  template <typename S>
  __device__ static __forceinline__ void step(const DynParams& P, const S* x, const S* u, S* xn) {
    typedef typename real_of<S>::type R;
    const R dt = (R)P.v[0], g = (R)P.v[1], m = (R)P.v[2], l = (R)P.v[3], mt = (R)P.v[4];
    const S uc = m_clamp(u[0], -mt, mt);
    const S cth = x[0], sth = x[1], dth = x[2];
    const S th = m_atan2(sth, cth);
    const S newdth = dth + dt * ((R(-3) * g / (R(2) * l)) * (-sth) + R(3) * uc / (m * l * l));
    const S newth = th + newdth * dt;
    m_sincos(newth, xn[1], xn[0]);
    xn[2] = newdth;
  }
};

// params: dt, gravity, masscart, masspole, length, total_mass, polemass_length, force_mag
// (qpth/env_dx/cartpole.py:63-96; the float32 model constants are rounded on the host the way
// the reference's float32 parameter tensor rounds them)
struct CartpoleDx {
  static constexpr int NX = 5, NU = 1;
  template <typename S>
  __device__ static __forceinline__ void step(const DynParams& P, const S* x, const S* u, S* xn) {
    typedef typename real_of<S>::type R;
    const R dt = (R)P.v[0], gravity = (R)P.v[1], masspole = (R)P.v[3], length = (R)P.v[4];
    const R total_mass = (R)P.v[5], pml = (R)P.v[6], fmag = (R)P.v[7];
    const S uc = m_clamp(u[0], -fmag, fmag);
    const S px = x[0], dx = x[1], cth = x[2], sth = x[3], dth = x[4];
    const S th = m_atan2(sth, cth);
    const R itm = R(1) / total_mass;   // one division instead of three per step
    const S cart_in = (uc + pml * (dth * dth) * sth) * itm;
    const S th_acc = (gravity * sth - cth * cart_in) / (length * (R(4.0 / 3.0) - masspole * (cth * cth) * itm));
    const S xacc = cart_in - pml * th_acc * cth * itm;
    const S nth = th + dt * dth;
    xn[0] = px + dt * dx;
    xn[1] = dx + dt * xacc;
    m_sincos(nth, xn[3], xn[2]);
    xn[4] = dth + dt * th_acc;
  }
};

// params (deqmpc/rex_quadrotor.py:9-49; values the reference keeps in float32 tensors are passed
// already rounded through float32 by the host):
//   0 dt, 1 mass, 2 act_scale, 3 kf used by forces() (literal 0.0244101), 4 kf, 5 km, 6 bf,
//   7 motor_dist, 8..10 mass*g, 11..13 Bf, 14..22 J, 23..31 J^-1, 32..43 ss (4 x 3), 44..46 cd,
//   47..49 cross_A
struct RexQuadrotor {
  static constexpr int NX = 12, NU = 4;

  template <typename S, typename R>
  __device__ static __forceinline__ void quatrot(const S* q, const R* r, S* out) {  // rexquad_utils.py:211-220, constant r
    const S qs = q[0];
    const S dotqq = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    const S dotqr = q[1] * r[0] + q[2] * r[1] + q[3] * r[2];
    const S c0 = q[2] * r[2] - q[3] * r[1], c1 = q[3] * r[0] - q[1] * r[2], c2 = q[1] * r[1] - q[2] * r[0];
    const S a = qs * qs - dotqq;
    out[0] = a * r[0] + R(2) * q[1] * dotqr + R(2) * qs * c0;
    out[1] = a * r[1] + R(2) * q[2] * dotqr + R(2) * qs * c1;
    out[2] = a * r[2] + R(2) * q[3] * dotqr + R(2) * qs * c2;
  }
  template <typename S, typename R>
  __device__ static __forceinline__ void quatrot_v(const S* q, const S* r, S* out) {  // same, variable r
    const S qs = q[0];
    const S dotqq = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    const S dotqr = q[1] * r[0] + q[2] * r[1] + q[3] * r[2];
    const S c0 = q[2] * r[2] - q[3] * r[1], c1 = q[3] * r[0] - q[1] * r[2], c2 = q[1] * r[1] - q[2] * r[0];
    const S a = qs * qs - dotqq;
    out[0] = a * r[0] + R(2) * q[1] * dotqr + R(2) * qs * c0;
    out[1] = a * r[1] + R(2) * q[2] * dotqr + R(2) * qs * c1;
    out[2] = a * r[2] + R(2) * q[3] * dotqr + R(2) * qs * c2;
  }
  template <typename S, typename R>
  __device__ static __forceinline__ void mrp2quat(const S* m, R sgn, S* q) {  // rexquad_utils.py:297-300 on sgn * m
    const S sq = m[0] * m[0] + m[1] * m[1] + m[2] * m[2];
    const S inv = R(1) / (R(1) + sq);
    q[0] = (R(1) - sq) * inv;
    q[1] = (R(2) * sgn) * m[0] * inv;
    q[2] = (R(2) * sgn) * m[1] * inv;
    q[3] = (R(2) * sgn) * m[2] * inv;
  }

  // continuous dynamics (rex_quadrotor.py:113-128); us = act_scale * u
  template <typename S>
  __device__ static __forceinline__ void deriv(const DynParams& P, const S* x, const S* us, S* dxdt) {
    typedef typename real_of<S>::type R;
    const S* pm = x + 3; const S* v = x + 6; const S* w = x + 9;
    S q[4], qn[4];
    mrp2quat<S, R>(pm, R(1), q);
    // the quaternion of -m is the conjugate: (2 * -1) m inv == -(2 m inv) exactly, so no second evaluation (one division)
    qn[0] = q[0]; qn[1] = -q[1]; qn[2] = -q[2]; qn[3] = -q[3];
    // forces (rex_quadrotor.py:51-68)
    const R kff = (R)P.v[3];
    const S Fz = kff * us[0] + kff * us[1] + kff * us[2] + kff * us[3];
    R mg[3] = {(R)P.v[8], (R)P.v[9], (R)P.v[10]};
    S grav[3];
    quatrot<S, R>(qn, mg, grav);
    S f[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const S df = (-m_sign(pm[i]) * R(0.5) * R(1.27)) * (pm[i] * pm[i]) * ((R)P.v[44 + i] * (R)P.v[47 + i]);
      f[i] = df + grav[i] + (R)P.v[11 + i];
    }
    f[2] = f[2] + Fz;
    // moments (rex_quadrotor.py:70-86)
    const R kf = (R)P.v[4], km = (R)P.v[5], bf = (R)P.v[6], L = (R)P.v[7];
    S tau[3];
    tau[0] = S(R(0)); tau[1] = S(R(0));
    tau[2] = km * us[0] - km * us[1] + km * us[2] - km * us[3];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const S bz = kf * us[i] + bf;
      const R ax = L * (R)P.v[32 + 3 * i], ay = L * (R)P.v[32 + 3 * i + 1];
      tau[0] = tau[0] + ay * bz;
      tau[1] = tau[1] - ax * bz;
    }
    // kinematics (rexquad_utils.py:393-403)
    const S p0 = pm[0], p1 = pm[1], p2 = pm[2];
    const S a00 = R(1) + p0 * p0 - p1 * p1 - p2 * p2, a01 = R(2) * (p0 * p1 - p2), a02 = R(2) * (p0 * p2 + p1);
    const S a10 = R(2) * (p1 * p0 + p2), a11 = R(1) - p0 * p0 + p1 * p1 - p2 * p2, a12 = R(2) * (p1 * p2 - p0);
    const S a20 = R(2) * (p2 * p0 - p1), a21 = R(2) * (p2 * p1 + p0), a22 = R(1) - p0 * p0 - p1 * p1 + p2 * p2;
    S pdot[3];
    quatrot_v<S, R>(q, v, pdot);
    dxdt[0] = pdot[0]; dxdt[1] = pdot[1]; dxdt[2] = pdot[2];
    dxdt[3] = R(0.25) * (a00 * w[0] + a01 * w[1] + a02 * w[2]);
    dxdt[4] = R(0.25) * (a10 * w[0] + a11 * w[1] + a12 * w[2]);
    dxdt[5] = R(0.25) * (a20 * w[0] + a21 * w[1] + a22 * w[2]);
    const R im = R(1) / (R)P.v[1];
    dxdt[6] = f[0] * im - (w[1] * v[2] - w[2] * v[1]);
    dxdt[7] = f[1] * im - (w[2] * v[0] - w[0] * v[2]);
    dxdt[8] = f[2] * im - (w[0] * v[1] - w[1] * v[0]);
    S Jw[3], tt[3];
#pragma unroll
    for (int i = 0; i < 3; i++) Jw[i] = (R)P.v[14 + 3 * i] * w[0] + (R)P.v[15 + 3 * i] * w[1] + (R)P.v[16 + 3 * i] * w[2];
    tt[0] = tau[0] - (w[1] * Jw[2] - w[2] * Jw[1]);
    tt[1] = tau[1] - (w[2] * Jw[0] - w[0] * Jw[2]);
    tt[2] = tau[2] - (w[0] * Jw[1] - w[1] * Jw[0]);
#pragma unroll
    for (int i = 0; i < 3; i++) dxdt[9 + i] = (R)P.v[23 + 3 * i] * tt[0] + (R)P.v[24 + 3 * i] * tt[1] + (R)P.v[25 + 3 * i] * tt[2];
  }

  template <typename S>
  __device__ static void step(const DynParams& P, const S* x, const S* u, S* xn) {  // RK4, rex_quadrotor.py:98-108
    typedef typename real_of<S>::type R;
    const R dt = (R)P.v[0], dt2 = dt / R(2), as = (R)P.v[2];
    // The four stages share ONE copy of the derivative code (a rolled loop; same operations in the same order as
    // ((k1 + 2 k2) + 2 k3) + k4): ncu showed 26 % of the stall samples of k_al_solve<RexQuadrotor> waiting for instructions,
    // with four inlined copies of `deriv` per step and four instantiations of `step` in the kernel.
    S us[NU], y[NX], k[NX], acc[NX];
#pragma unroll
    for (int i = 0; i < NU; i++) us[i] = as * u[i];
#pragma unroll
    for (int i = 0; i < NX; i++) y[i] = x[i];
#pragma unroll 1
    for (int st = 0; st < 4; st++) {
      deriv<S>(P, y, us, k);
      const R wgt = (st == 0 || st == 3) ? R(1) : R(2);
      const R h = (st == 2) ? dt : dt2;
#pragma unroll
      for (int i = 0; i < NX; i++) {
        if (st == 0) acc[i] = k[i]; else acc[i] = acc[i] + wgt * k[i];
        y[i] = x[i] + h * k[i];
      }
    }
#pragma unroll
    for (int i = 0; i < NX; i++) xn[i] = x[i] + (dt / R(6)) * acc[i];
  }
};

// ---- deqmpc/my_envs: second-order systems x = (q, qd), classical RK4 of qdd = Model::accel(q, qd, u)
template <class Model>
struct SecondOrderRK4 {
  static constexpr int NQ = Model::NQ, NX = 2 * Model::NQ, NU = 1;
  template <typename S>
  __device__ static __forceinline__ void step(const DynParams& P, const S* x, const S* u, S* xn) {
    typedef typename real_of<S>::type R;
    const R h = (R)P.v[0], h2 = h / R(2);
    S y[NX], a1[NQ], a2[NQ], a3[NQ], a4[NQ], v2[NQ], v3[NQ], v4[NQ];
    Model::template accel<S>(P, x, x + NQ, u[0], a1);
#pragma unroll
    for (int i = 0; i < NQ; i++) { y[i] = x[i] + h2 * x[NQ + i]; v2[i] = x[NQ + i] + h2 * a1[i]; y[NQ + i] = v2[i]; }
    Model::template accel<S>(P, y, y + NQ, u[0], a2);
#pragma unroll
    for (int i = 0; i < NQ; i++) { y[i] = x[i] + h2 * v2[i]; v3[i] = x[NQ + i] + h2 * a2[i]; y[NQ + i] = v3[i]; }
    Model::template accel<S>(P, y, y + NQ, u[0], a3);
#pragma unroll
    for (int i = 0; i < NQ; i++) { y[i] = x[i] + h * v3[i]; v4[i] = x[NQ + i] + h * a3[i]; y[NQ + i] = v4[i]; }
    Model::template accel<S>(P, y, y + NQ, u[0], a4);
#pragma unroll
    for (int i = 0; i < NQ; i++) {
      xn[i] = x[i] + (h / R(6)) * (x[NQ + i] + R(2) * v2[i] + R(2) * v3[i] + v4[i]);
      xn[NQ + i] = x[NQ + i] + (h / R(6)) * (a1[i] + R(2) * a2[i] + R(2) * a3[i] + a4[i]);
    }
  }
};

// params: dt, 1/I, m g l / I     (pendulum1l: thdd = 4 u - 2 g sin th, angle from the downward rest position)
struct Pendulum1LModel {
  static constexpr int NQ = 1;
  template <typename S>
  __device__ static __forceinline__ void accel(const DynParams& P, const S* q, const S* qd, const S& u, S* a) {
    typedef typename real_of<S>::type R;
    a[0] = (R)P.v[1] * u - (R)P.v[2] * m_sin(q[0]);
  }
};

// params: dt, total mass mt, pole first moment ml, pole inertia about the joint I, g
//   [mt  -ml c] [xdd ]   [u - ml s thd^2]
//   [-ml c   I] [thdd] = [    ml g s    ]        (theta = 0 upright, counter-clockwise positive)
// cartpole1l: (11, 1, 2, 9.81); cartpole1l_v2: (0.7, 0.1, 0.05, 9.81)
struct Cartpole1LModel {
  static constexpr int NQ = 2;
  template <typename S>
  __device__ static __forceinline__ void accel(const DynParams& P, const S* q, const S* qd, const S& u, S* a) {
    typedef typename real_of<S>::type R;
    const R mt = (R)P.v[1], ml = (R)P.v[2], I = (R)P.v[3], g = (R)P.v[4];
    S s, c;
    m_sincos(q[1], s, c);
    const S mc = ml * c;
    const S r0 = u - ml * s * (qd[1] * qd[1]);
    const S r1 = (ml * g) * s;
    const R imt = R(1) / mt;             // loop-invariant: one division per step instead of two per stage
    const S f = mc * imt;                // eliminate the cart row
    a[1] = (r1 + f * r0) / (I - f * mc);
    a[0] = (r0 + mc * a[1]) * imt;
  }
};

// params: dt, mt, h1, h2, J1, J2, k, g  (cart + two links; Lagrange's equations in absolute link angles
// phi1 = th1, phi2 = th1 + th2, joint torques on the relative angles; cartpole2l: 12, 2, 1, 3, 2, 1, 9.81)
struct Cartpole2LModel {
  static constexpr int NQ = 3;
  template <typename S>
  __device__ static __forceinline__ void accel(const DynParams& P, const S* q, const S* qd, const S& u, S* a) {
    typedef typename real_of<S>::type R;
    const R mt = (R)P.v[1], h1 = (R)P.v[2], h2 = (R)P.v[3], J1 = (R)P.v[4], J2 = (R)P.v[5], k = (R)P.v[6], g = (R)P.v[7];
    const S p2 = q[1] + q[2], w1 = qd[1], w2 = qd[1] + qd[2];
    S s1, c1, s2, c2;
    m_sincos(q[1], s1, c1);
    m_sincos(p2, s2, c2);
    const S s12 = s1 * c2 - c1 * s2, c12 = c1 * c2 + s1 * s2;
    const S m01 = -h1 * c1, m02 = -h2 * c2, m12 = k * c12;
    const S r0 = u - h1 * s1 * (w1 * w1) - h2 * s2 * (w2 * w2);
    const S r1 = (h1 * g) * s1 - k * s12 * (w2 * w2);
    const S r2 = (h2 * g) * s2 + k * s12 * (w1 * w1);
    // symmetric elimination, no pivoting (the mass matrix is SPD)
    const R imt = R(1) / mt;
    const S f1 = m01 * imt, f2 = m02 * imt;
    const S b11 = J1 - f1 * m01, b12 = m12 - f1 * m02, b22 = J2 - f2 * m02;
    const S t1 = r1 - f1 * r0, t2 = r2 - f2 * r0;
    const S f3 = b12 / b11;
    const S ph2 = (t2 - f3 * t1) / (b22 - f3 * b12);
    const S ph1 = (t1 - b12 * ph2) / b11;
    a[0] = (r0 - m01 * ph1 - m02 * ph2) * imt;
    a[1] = ph1;
    a[2] = ph2 - ph1;
  }
};

// ---- deqmpc/envs_v1.py: closed-form cart-pole accelerations under the same classical RK4
// OneLinkCartpoleDynamics (envs_v1.py:28-82): params dt, M, m, l, g (the reference sets g = -9.81); state (x, th, xd, thd)
struct Cartpole1LV1Model {
  static constexpr int NQ = 2;
  template <typename S>
  __device__ static __forceinline__ void accel(const DynParams& P, const S* q, const S* qd, const S& u, S* a) {
    typedef typename real_of<S>::type R;
    const R M = (R)P.v[1], m = (R)P.v[2], l = (R)P.v[3], g = (R)P.v[4];
    S s, c;
    m_sincos(q[1], s, c);
    const S w2 = qd[1] * qd[1];
    const S den = M + m * (s * s);
    a[0] = (u + (m * l) * w2 * s - (m * g) * s * c) / den;
    a[1] = (-(u * c) - (m * l) * w2 * s * c + ((M + m) * g) * s) / (l * den);
  }
};
// TwoLinkCartpoleDynamics (envs_v1.py:226-310): the OpenOCL double cart-pole with its numeric constants
// (M = 5, m1 = m2 = l1 = l2 = 1, g = 9.81) folded into the formulas, f = -u; params dt only
struct Cartpole2LV1Model {
  static constexpr int NQ = 3;
  template <typename S>
  __device__ static __forceinline__ void accel(const DynParams& P, const S* q, const S* qd, const S& u, S* a) {
    typedef typename real_of<S>::type R;
    const S f = -u;
    const S q1 = q[1], q2 = q[2], w1 = qd[1], w2 = qd[2];
    const S w11 = w1 * w1, w12 = w1 * w2, w22 = w2 * w2;
    S c1, s1, c2, s2, c_1m2, s_1m2, c_1p2, s_1p2, c_1p22, s_1p22, c_21, s_21, c_22, s_22, c_21p2, s_21p2, c_21p22, s_21p22;
    m_sincos(q1, s1, c1);
    m_sincos(q2, s2, c2);
    m_sincos(q1 - q2, s_1m2, c_1m2);
    m_sincos(q1 + q2, s_1p2, c_1p2);
    m_sincos(q1 + R(2) * q2, s_1p22, c_1p22);
    m_sincos(R(2) * q1, s_21, c_21);
    m_sincos(R(2) * q2, s_22, c_22);
    m_sincos(R(2) * q1 + q2, s_21p2, c_21p2);
    m_sincos(R(2) * q1 + R(2) * q2, s_21p22, c_21p22);
    const S den = R(3) * c_21 - R(22) * c_22 - c_21p22 + R(34);
    a[0] = (R(4) * f * c_22 - R(6) * f + R(4) * w11 * c1 + w11 * c_1m2 - w11 * c_1p22 + R(2) * w12 * c_1m2 + w22 * c_1m2
            - R(29.43) * s_21 + R(9.81) * s_21p22) / den;
    a[1] = (-(R(8) * f * s1) + R(4) * f * s_1p22 + R(3) * w11 * s_21 + R(23) * w11 * s2 + R(22) * w11 * s_22 + w11 * s_21p2
            + R(46) * w12 * s2 + R(2) * w12 * s_21p2 + R(23) * w22 * s2 + w22 * s_21p2 - R(490.5) * c1 + R(215.82) * c_1p22) / den;
    const S t = R(3) * s1 + s_1p2;
    a[2] = -((R(100) * w11 * s2 + R(981) * c_1p2) * (-(t * t) + R(28) * c2 + R(42))
             + R(0.5) * (R(200) * w12 * s2 + R(100) * w22 * s2 - R(2943) * c1 - R(981) * c_1p2)
                   * (R(25) * c2 + R(3) * c_21p2 + c_21p22 + R(13))
             + R(50) * (R(2) * s1 + R(3) * s_1m2 - R(2) * s_1p2 - s_1p22)
                   * (-(R(2) * f) + R(3) * w11 * c1 + w11 * c_1p2 + R(2) * w12 * c_1p2 + w22 * c_1p2))
           / (R(75) * c_21 - R(550) * c_22 - R(25) * c_21p22 + R(850));
  }
};

typedef SecondOrderRK4<Pendulum1LModel> Pendulum1L;
typedef SecondOrderRK4<Cartpole1LV1Model> Cartpole1LV1;
typedef SecondOrderRK4<Cartpole2LV1Model> Cartpole2LV1;
typedef SecondOrderRK4<Cartpole1LModel> Cartpole1L;
typedef SecondOrderRK4<Cartpole2LModel> Cartpole2L;

}  // namespace b200mpc
