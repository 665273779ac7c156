"""TEST INFRASTRUCTURE -- a second, independent restatement of Trixi.jl's 3D Euler flux-differencing DGSEM (numpy, uniform
periodic TreeMesh, polydeg 3), written without looking at oracle/*.hpp: whole-array operations over an (ez, ey, ex, k, j, i,
v) grid instead of the oracle's element/interface containers. It exists to separate "the oracle has a 3D-only bug" from
"the recalled Trixi digits are off" for tests/golden/trixi_regression_norms.json:euler_ec_3d.

Follows: flux differencing /root/reference/src/solvers/dg_3d_kernel.jl:188-257 (all l != i, which equals Trixi's symmetric
form), surface integral :1773-1799, Jacobian :1802-1818, max_dt /root/reference/src/callbacks_step/stepsize_dg_3d.jl:20-45,
error norms /root/reference/src/callbacks_step/analysis_dg_3d.jl:45-89; flux_ranocha / ln_mean / weak blast wave / CK2N54
as in SURVEY.md Appendix A.6-A.8.
"""
import numpy as np
from numpy.polynomial import legendre as L

GAMMA = 1.4
RK_A = [0, -567301805773 / 1357537059087, -2404267990393 / 2016746695238, -3550918686646 / 2091501179385,
        -1275806237668 / 842570457699]
RK_B = [1432997174477 / 9575080441755, 5161836677717 / 13612068292357, 1720146321549 / 2090206949498,
        3134564353537 / 4481467310338, 2277821191437 / 14882151754819]


def lgl(n):
    P = L.Legendre.basis(n - 1)
    x = np.concatenate([[-1.0], np.sort(P.deriv().roots().real), [1.0]])
    return x, 2 / (n * (n - 1) * P(x) ** 2)


def lagrange_matrix(xa, xn):
    V = np.ones((len(xa), len(xn)))
    for j in range(len(xn)):
        for m in range(len(xn)):
            if m != j:
                V[:, j] *= (xa - xn[m]) / (xn[j] - xn[m])
    return V


XN, WN = lgl(4)
XA, WA = lgl(7)
VDM = lagrange_matrix(XA, XN)


def _dsplit():
    bw = np.array([1 / np.prod([XN[j] - XN[m] for m in range(4) if m != j]) for j in range(4)])
    D = np.zeros((4, 4))
    for i in range(4):
        for j in range(4):
            if i != j:
                D[i, j] = bw[j] / bw[i] / (XN[i] - XN[j])
        D[i, i] = -D[i].sum()
    Ds = 2 * D
    Ds[0, 0] += 1 / WN[0]
    Ds[3, 3] -= 1 / WN[3]
    return Ds


DSPLIT = _dsplit()


def ln_mean(x, y):
    f2 = (x * (x - 2 * y) + y * y) / (x * (x + 2 * y) + y * y)
    small = f2 < 1e-4
    return np.where(small, (x + y) / (2 + f2 * (2 / 3 + f2 * (2 / 5 + f2 * 2 / 7))),
                    (y - x) / np.log(np.where(small, np.e, y / x)))


def inv_ln_mean(x, y):
    f2 = (x * (x - 2 * y) + y * y) / (x * (x + 2 * y) + y * y)
    small = f2 < 1e-4
    return np.where(small, (2 + f2 * (2 / 3 + f2 * (2 / 5 + f2 * 2 / 7))) / (x + y),
                    np.log(np.where(small, np.e, y / x)) / np.where(small, 1.0, y - x))


def cons2prim(u):
    rho = u[..., 0]
    v = u[..., 1:4] / rho[..., None]
    return rho, v, (GAMMA - 1) * (u[..., 4] - 0.5 * (u[..., 1:4] * v).sum(-1))


def flux_ranocha(ul, ur, o):
    rl, vl, pl = cons2prim(ul)
    rr, vr, pr = cons2prim(ur)
    rho_mean = ln_mean(rl, rr)
    inv_rho_p_mean = pl * pr * inv_ln_mean(rl * pr, rr * pl)
    va, pa, vsq = 0.5 * (vl + vr), 0.5 * (pl + pr), 0.5 * (vl * vr).sum(-1)
    f = np.empty_like(ul)
    f[..., 0] = rho_mean * va[..., o]
    for d in range(3):
        f[..., 1 + d] = f[..., 0] * va[..., d]
    f[..., 1 + o] += pa
    f[..., 4] = f[..., 0] * (vsq + inv_rho_p_mean / (GAMMA - 1)) + 0.5 * (pl * vr[..., o] + pr * vl[..., o])
    return f


def weak_blast_wave(x):
    r = np.sqrt((x ** 2).sum(-1))
    out = r > 0.5
    phi = np.arctan2(x[..., 1], x[..., 0])
    safe_r = np.where(r == 0, 1.0, r)
    theta = np.where(r == 0, 0.0, np.arccos(np.where(r == 0, 1.0, x[..., 2] / safe_r)))
    rho, p = np.where(out, 1.0, 1.1691), np.where(out, 1.0, 1.245)
    v = [np.where(out, 0.0, 0.1882 * c) for c in (np.cos(phi) * np.sin(theta), np.sin(phi) * np.sin(theta),
                                                  np.cos(theta))]
    return np.stack([rho, rho * v[0], rho * v[1], rho * v[2],
                     p / (GAMMA - 1) + 0.5 * rho * (v[0] ** 2 + v[1] ** 2 + v[2] ** 2)], -1)


class UniformPeriodic3D:
    """[cmin, cmax]^3, 2^level elements per direction; arrays are (ez, ey, ex, k, j, i, v)."""

    def __init__(self, level, cmin=-2.0, cmax=2.0):
        self.n = 2 ** level
        self.length = cmax - cmin
        self.dx = self.length / self.n
        self.inv_jacobian = 2 / self.dx
        c = cmin + self.dx * (np.arange(self.n) + 0.5)
        n = self.n
        self.x = np.zeros((n, n, n, 4, 4, 4, 3))
        self.x[..., 0] = c[None, None, :, None, None, None] + self.dx / 2 * XN[None, None, None, None, None, :]
        self.x[..., 1] = c[None, :, None, None, None, None] + self.dx / 2 * XN[None, None, None, None, :, None]
        self.x[..., 2] = c[:, None, None, None, None, None] + self.dx / 2 * XN[None, None, None, :, None, None]

    def rhs(self, u):
        du = np.zeros_like(u)
        for o, (eax, nax) in enumerate([(2, 5), (1, 4), (0, 3)]):
            um, dm = np.moveaxis(u, nax, 0), np.moveaxis(du, nax, 0)
            for i in range(4):
                for l in range(4):
                    if l != i:
                        dm[i] += DSPLIT[i, l] * flux_ranocha(um[i], um[l], o)
            f = flux_ranocha(um[3], np.roll(um[0], -1, axis=eax), o)     # +face of e with -face of e+1
            dm[3] += f / WN[3]
            dm[0] -= np.roll(f, 1, axis=eax) / WN[0]
        return -self.inv_jacobian * du

    def max_dt(self, u):
        rho, v, p = cons2prim(u)
        lam = np.abs(v) + np.sqrt(GAMMA * p / rho)[..., None]
        return 2 / (4 * (self.inv_jacobian * lam.max(axis=(3, 4, 5)).sum(-1)).max())

    def solve(self, u, tend, cfl):
        t, steps, tmp = 0.0, 0, np.zeros_like(u)
        while True:
            dt = cfl * self.max_dt(u)
            if t + dt >= tend - 1e-14:
                dt = tend - t
            for s in range(5):
                tmp = RK_A[s] * tmp + dt * self.rhs(u)
                u = u + RK_B[s] * tmp
            t, steps = t + dt, steps + 1
            if abs(t - tend) < 1e-14:
                return u, steps

    def error_norms(self, u, ic):
        ua = np.einsum('ai,bj,ck,zyxkjiv->zyxcbav', VDM, VDM, VDM, u)
        xa = np.einsum('ai,bj,ck,zyxkjid->zyxcbad', VDM, VDM, VDM, self.x)
        w = np.einsum('c,b,a->cba', WA, WA, WA) * (self.dx / 2) ** 3
        d = ic(xa) - ua
        return (np.sqrt(np.einsum('zyxcbav,cba->v', d ** 2, w) / self.length ** 3),
                np.abs(d).max(axis=(0, 1, 2, 3, 4, 5)))

    def morton_permutation(self, element_centers):
        """perm[ez, ey, ex] = index of that element in a list given by its cell centres (Trixi leaf order)."""
        idx = np.round((element_centers - (self.x[0, 0, 0, 0, 0, 0] - self.dx / 2 * (1 + XN[0]))) / self.dx - 0.5)
        idx = idx.astype(int)
        perm = np.zeros((self.n,) * 3, dtype=int)
        perm[idx[:, 2], idx[:, 1], idx[:, 0]] = np.arange(len(idx))
        return perm
