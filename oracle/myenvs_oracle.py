"""CPU oracle for the CasADi-generated `deqmpc/my_envs` dynamics.  TEST INFRASTRUCTURE, NOT
PRODUCT: imported only by tests/, oracle/gen_golden_myenvs.py and bench.py's CPU-baseline leg.

Two back-ends behind the same `package.dynamics / package.derivatives` interface the reference's torch
extensions export (deqmpc/my_envs/cartpole1l/src/dynamics.cpp:40-54):

* `RefPackage(name)`  -- the REFERENCE ITSELF: oracle/_ref/lib<name>_ref.so, built by oracle/Makefile from the
  reference's own generated_dynamics.c / generated_derivatives.c, driven exactly like
  cartpole1l/src/dynamics_cpu.cpp:8-27,30-57 (one eval_forward_dynamics / eval_forward_derivatives call per
  batch row; arg = {q, qdot, tau, h}).
* `PortPackage(name)` -- a restatement: the rigid-body model re-derived from the generated straight-line
  code (classical RK4 of  M(q) qdd = tau - bias(q, qd)), Jacobians by forward-mode duals.  PINNED against
  RefPackage by oracle/gen_golden_myenvs.py and by tests/test_myenvs_oracle_cpu.py on the committed goldens.

`Dynamics` restates deqmpc/my_envs/dynamics.py:27-108,249-258 (state split, tau on the first joint only,
Jacobian block assembly and transposes)."""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NQ = {"cartpole1l": 2, "cartpole1l_v2": 2, "cartpole2l": 3, "pendulum1l": 1}


class RefPackage:
    def __init__(self, name):
        path = os.path.join(HERE, "_ref", f"lib{name}_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: run `make -C oracle` where /root/reference is mounted")
        self.lib = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
        self.nq = NQ[name]
        for f in (self.lib.eval_forward_dynamics, self.lib.eval_forward_derivatives):
            f.restype = ctypes.c_int

    def _call(self, fn, ins, outs):
        dp = ctypes.POINTER(ctypes.c_double)
        arg = (dp * len(ins))(*[a.ctypes.data_as(dp) for a in ins])
        res = (dp * len(outs))(*[a.ctypes.data_as(dp) for a in outs])
        fn(arg, res, None, None, 0)

    def dynamics(self, q, qd, tau, h):
        q, qd, tau, h = (np.ascontiguousarray(a, dtype=np.float64) for a in (q, qd, tau, h))
        qo, qdo = np.zeros_like(q), np.zeros_like(q)
        for b in range(q.shape[0]):  # dynamics_cpu.cpp:36-42
            self._call(self.lib.eval_forward_dynamics, [q[b], qd[b], tau[b], h[b]], [qo[b], qdo[b]])
        return qo, qdo

    def derivatives(self, q, qd, tau, h):
        q, qd, tau, h = (np.ascontiguousarray(a, dtype=np.float64) for a in (q, qd, tau, h))
        n = self.nq
        J = [np.zeros((q.shape[0], n, n)) for _ in range(6)]
        for b in range(q.shape[0]):  # dynamics_cpu.cpp:50-57
            self._call(self.lib.eval_forward_derivatives, [q[b], qd[b], tau[b], h[b]], [j[b] for j in J])
        return tuple(J)


# ------------------------------------------------------------------------------------------ port
class _Dual:
    """value (N,) + K directional derivatives (N,K)"""
    __slots__ = ("v", "d")
    __array_ufunc__ = None  # ndarray * _Dual defers to _Dual.__rmul__

    def __init__(self, v, d):
        self.v, self.d = v, d

    @staticmethod
    def lift(x):
        return x if isinstance(x, _Dual) else _Dual(np.asarray(x, dtype=np.float64), 0.0)

    def __add__(self, o):
        o = _Dual.lift(o)
        return _Dual(self.v + o.v, self.d + o.d)
    __radd__ = __add__

    def __neg__(self):
        return _Dual(-self.v, -self.d)

    def __sub__(self, o):
        return self + (-_Dual.lift(o))

    def __rsub__(self, o):
        return (-self) + o

    def __mul__(self, o):
        o = _Dual.lift(o)
        sv, ov = np.asarray(self.v)[..., None], np.asarray(o.v)[..., None]
        return _Dual(self.v * o.v, self.d * ov + sv * o.d)
    __rmul__ = __mul__

    def __truediv__(self, o):
        o = _Dual.lift(o)
        r = self.v / o.v
        return _Dual(r, (self.d - np.asarray(r)[..., None] * o.d) / np.asarray(o.v)[..., None])

    def __rtruediv__(self, o):
        return _Dual.lift(o) / self


def _sin(a):
    return _Dual(np.sin(a.v), np.cos(a.v)[..., None] * a.d) if isinstance(a, _Dual) else np.sin(a)


def _cos(a):
    return _Dual(np.cos(a.v), -np.sin(a.v)[..., None] * a.d) if isinstance(a, _Dual) else np.cos(a)


def _solve_sym(M, r):
    """Gaussian elimination without pivoting on a small SPD system given as nested lists of scalars/duals."""
    n = len(r)
    M = [row[:] for row in M]
    r = r[:]
    for k in range(n):
        for i in range(k + 1, n):
            f = M[i][k] / M[k][k]
            for j in range(k, n):
                M[i][j] = M[i][j] - f * M[k][j]
            r[i] = r[i] - f * r[k]
    x = [None] * n
    for i in reversed(range(n)):
        acc = r[i]
        for j in range(i + 1, n):
            acc = acc - M[i][j] * x[j]
        x[i] = acc / M[i][i]
    return x


G = 9.81
# (total mass, pole mass * com distance, pole inertia about the joint)
CARTPOLE1 = {"cartpole1l": (11.0, 1.0, 2.0), "cartpole1l_v2": (0.7, 0.1, 0.05)}
# total mass, first moments h1 = m1 a1 + m2 L1, h2 = m2 a2, joint inertias J1 = I1 + m1 a1^2 + m2 L1^2,
# J2 = I2 + m2 a2^2, coupling k = m2 L1 a2 (identified from the generated code: least-squares residual 1e-5
# with accelerations taken by a 1e-7 step, then confirmed to rounding by gen_golden_myenvs.py)
CARTPOLE2 = (12.0, 2.0, 1.0, 3.0, 2.0, 1.0)


def accel(name, q, qd, tau):
    """Continuous-time joint accelerations of the models behind the generated code (lists of scalars)."""
    if name == "pendulum1l":
        # thdd = 4 tau - 2 g sin(th); angle measured from the downward rest position
        return [4.0 * tau[0] - (2.0 * G) * _sin(q[0])]
    if name in CARTPOLE1:
        mt, ml, I = CARTPOLE1[name]
        s, c = _sin(q[1]), _cos(q[1])
        M = [[mt, -ml * c], [-ml * c, I]]
        r = [tau[0] - ml * s * qd[1] * qd[1], tau[1] + (ml * G) * s]
        return _solve_sym(M, r)
    if name == "cartpole2l":
        # Lagrange's equations in (x, phi1, phi2) with ABSOLUTE link angles phi1 = th1, phi2 = th1 + th2 (zero =
        # upright, counter-clockwise positive); the joint torques act on the RELATIVE angles
        mt, h1, h2, J1, J2, k = CARTPOLE2
        p1, p2, w1, w2 = q[1], q[1] + q[2], qd[1], qd[1] + qd[2]
        s1, c1, s2, c2 = _sin(p1), _cos(p1), _sin(p2), _cos(p2)
        s12, c12 = _sin(p1 - p2), _cos(p1 - p2)
        M = [[mt, -h1 * c1, -h2 * c2], [-h1 * c1, J1, k * c12], [-h2 * c2, k * c12, J2]]
        r = [tau[0] - h1 * s1 * w1 * w1 - h2 * s2 * w2 * w2,
             tau[1] - tau[2] + (h1 * G) * s1 - k * s12 * w2 * w2,
             tau[2] + (h2 * G) * s2 + k * s12 * w1 * w1]
        a = _solve_sym(M, r)
        return [a[0], a[1], a[2] - a[1]]
    raise NotImplementedError(name)


def rk4(name, q, qd, tau, h):
    """Classical RK4 on (q, qd) with step h: the integrator CasADi unrolled into the generated code."""
    n = len(q)

    def f(q_, qd_):
        return qd_, accel(name, q_, qd_, tau)
    k1q, k1v = f(q, qd)
    k2q, k2v = f([q[i] + (h / 2.0) * k1q[i] for i in range(n)], [qd[i] + (h / 2.0) * k1v[i] for i in range(n)])
    k3q, k3v = f([q[i] + (h / 2.0) * k2q[i] for i in range(n)], [qd[i] + (h / 2.0) * k2v[i] for i in range(n)])
    k4q, k4v = f([q[i] + h * k3q[i] for i in range(n)], [qd[i] + h * k3v[i] for i in range(n)])
    qn = [q[i] + (h / 6.0) * (k1q[i] + 2.0 * k2q[i] + 2.0 * k3q[i] + k4q[i]) for i in range(n)]
    vn = [qd[i] + (h / 6.0) * (k1v[i] + 2.0 * k2v[i] + 2.0 * k3v[i] + k4v[i]) for i in range(n)]
    return qn, vn


class PortPackage:
    def __init__(self, name):
        self.name, self.nq = name, NQ[name]

    def dynamics(self, q, qd, tau, h):
        n = self.nq
        q, qd, tau, h = (np.asarray(a, dtype=np.float64) for a in (q, qd, tau, h))
        qn, vn = rk4(self.name, [q[:, i] for i in range(n)], [qd[:, i] for i in range(n)], [tau[:, i] for i in range(n)],
                     h[:, 0])
        return np.stack(qn, 1), np.stack(vn, 1)

    def derivatives(self, q, qd, tau, h):
        """Six (N,nq,nq) blocks in the generated code's layout: block[b, i, j] = d out_j / d in_i."""
        n = self.nq
        q, qd, tau, h = (np.asarray(a, dtype=np.float64) for a in (q, qd, tau, h))
        N, K = q.shape[0], 3 * n
        eye = np.eye(K)

        def seed(a, off):
            return [_Dual(a[:, i], np.broadcast_to(eye[off + i], (N, K)).copy()) for i in range(n)]
        qn, vn = rk4(self.name, seed(q, 0), seed(qd, n), seed(tau, 2 * n), h[:, 0])
        dq = np.stack([o.d for o in qn], 2)   # (N, K in, n out)
        dv = np.stack([o.d for o in vn], 2)
        return (dq[:, 0:n], dq[:, n:2 * n], dq[:, 2 * n:], dv[:, 0:n], dv[:, n:2 * n], dv[:, 2 * n:])


class Dynamics:
    """deqmpc/my_envs/dynamics.py:27-108,249-258 on numpy arrays."""

    def __init__(self, package, nx, dt):
        self.package, self.nx, self.nu, self.nq, self.dt = package, nx, 1, nx // 2, dt

    def _split(self, state, action):
        state, action = np.asarray(state, dtype=np.float64), np.asarray(action, dtype=np.float64)
        tau = np.zeros((state.shape[0], self.nq))
        tau[:, 0] = action[:, 0]                                   # dynamics.py:55-56
        h = np.full((state.shape[0], 1), self.dt)
        return state[:, :self.nq].copy(), state[:, self.nq:].copy(), tau, h

    def forward(self, state, action):
        return np.concatenate(self.package.dynamics(*self._split(state, action)), -1)   # dynamics.py:61-63

    def derivatives(self, state, action):
        qq, qv, qt, vq, vv, vt = self.package.derivatives(*self._split(state, action))
        q_jac_x = np.concatenate((qq, qv), -2)                     # dynamics.py:100-107
        v_jac_x = np.concatenate((vq, vv), -2)
        x_jac_x = np.concatenate((q_jac_x, v_jac_x), -1)
        x_jac_u = np.concatenate((qt, vt), -1)[:, :1, :]
        return np.swapaxes(x_jac_x, -1, -2), np.swapaxes(x_jac_u, -1, -2)

    def dynamics_derivatives(self, state, action):
        return self.forward(state, action), self.derivatives(state, action)
