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

__device__ __forceinline__ double m_sin(double x) { return sin(x); }
__device__ __forceinline__ float m_sin(float x) { return sinf(x); }
__device__ __forceinline__ double m_cos(double x) { return cos(x); }
__device__ __forceinline__ float m_cos(float x) { return cosf(x); }
template <typename R, int N> __device__ __forceinline__ Dual<R, N> m_sin(const Dual<R, N>& a) {
  Dual<R, N> r; r.v = m_sin(a.v); const R c = m_cos(a.v);
#pragma unroll
  for (int i = 0; i < N; i++) r.d[i] = c * a.d[i];
  return r;
}
template <typename R, int N> __device__ __forceinline__ Dual<R, N> m_cos(const Dual<R, N>& a) {
  Dual<R, N> r; r.v = m_cos(a.v); const R s = -m_sin(a.v);
#pragma unroll
  for (int i = 0; i < N; i++) r.d[i] = s * a.d[i];
  return r;
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

template <typename S> struct real_of { typedef S type; };
template <typename R, int N> struct real_of<Dual<R, N>> { typedef R type; };

// ---------------------------------------------------------------------------- environments
constexpr int ENV_PENDULUM = 0, ENV_INTEGRATOR = 1, ENV_PENDULUM_DX = 2, ENV_CARTPOLE_DX = 3;
constexpr int MAX_PARAMS = 16;

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
    xn[0] = m_cos(newth);
    xn[1] = m_sin(newth);
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
    const S cart_in = (uc + pml * (dth * dth) * sth) / total_mass;
    const S th_acc = (gravity * sth - cth * cart_in) / (length * (R(4.0 / 3.0) - masspole * (cth * cth) / total_mass));
    const S xacc = cart_in - pml * th_acc * cth / total_mass;
    const S nth = th + dt * dth;
    xn[0] = px + dt * dx;
    xn[1] = dx + dt * xacc;
    xn[2] = m_cos(nth);
    xn[3] = m_sin(nth);
    xn[4] = dth + dt * th_acc;
  }
};

}  // namespace b200mpc
