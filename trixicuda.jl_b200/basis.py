"""LobattoLegendreBasisGPU / MortarL2GPU / SolutionAnalyzer operators on the host (numpy).

Mirrors the fields of /root/reference/src/solvers/basis_lobatto_legendre.jl:25-47,60-100 (basis),
:155-173 (mortar operators) and :116-132 (analyzer). In the Julia drop-in these matrices come straight from
Trixi.jl and are passed to `trixib200_create`; here they are computed with numpy (barycentric formulas,
Newton-polished Gauss/Lobatto nodes). All matrices are returned as (row, col) numpy arrays; the C ABI takes
them column-major (Julia order), see `_lib.pack_matrix`.
"""
import numpy as np
from numpy.polynomial import legendre as npleg


def _legendre(n, x):
    """P_n(x) and P_n'(x) by the three-term recurrence (unnormalised)."""
    x = np.asarray(x, dtype=np.float64)
    p0, p1 = np.ones_like(x), x.copy()
    d0, d1 = np.zeros_like(x), np.ones_like(x)
    if n == 0:
        return p0, d0
    for k in range(2, n + 1):
        p2 = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k
        d2 = d0 + (2 * k - 1) * p1
        p0, p1, d0, d1 = p1, p2, d1, d2
    return p1, d1


def gauss_lobatto_nodes_weights(n_nodes):
    p = n_nodes - 1
    if p == 0:
        return np.array([0.0]), np.array([2.0])
    if p == 1:
        return np.array([-1.0, 1.0]), np.array([1.0, 1.0])
    # interior nodes: roots of P_p'; start from the companion-matrix roots, polish with Newton on
    # q = P_{p+1} - P_{p-1} (q' = (2p+1) P_p)
    c = np.zeros(p + 1)
    c[p] = 1.0
    x = np.sort(npleg.legroots(npleg.legder(c)))
    for _ in range(4):
        pp1, _ = _legendre(p + 1, x)
        pm1, _ = _legendre(p - 1, x)
        pp, _ = _legendre(p, x)
        x = x - (pp1 - pm1) / ((2 * p + 1) * pp)
    x = 0.5 * (x - x[::-1])                      # enforce symmetry
    nodes = np.concatenate([[-1.0], x, [1.0]])
    pp, _ = _legendre(p, nodes)
    weights = 2.0 / (p * (p + 1) * pp ** 2)
    return nodes, weights


def gauss_nodes_weights(n_nodes):
    x = np.sort(npleg.legroots(np.eye(n_nodes + 1)[n_nodes]))
    for _ in range(4):
        pn, dn = _legendre(n_nodes, x)
        x = x - pn / dn
    x = 0.5 * (x - x[::-1])
    _, dn = _legendre(n_nodes, x)
    return x, 2.0 / ((1 - x ** 2) * dn ** 2)


def barycentric_weights(nodes):
    diff = nodes[:, None] - nodes[None, :]
    np.fill_diagonal(diff, 1.0)
    return 1.0 / diff.prod(axis=1)


def lagrange_interpolating_polynomials(x, nodes, wbary):
    hit = np.isclose(x, nodes, rtol=np.sqrt(np.finfo(float).eps), atol=0.0)
    if hit.any():
        return hit.astype(np.float64)
    t = wbary / (x - nodes)
    return t / t.sum()


def polynomial_interpolation_matrix(nodes_in, nodes_out):
    wb = barycentric_weights(nodes_in)
    return np.stack([lagrange_interpolating_polynomials(x, nodes_in, wb) for x in nodes_out])


def polynomial_derivative_matrix(nodes):
    n = nodes.shape[0]
    wb = barycentric_weights(nodes)
    D = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            if i != j:
                D[i, j] = wb[j] / wb[i] / (nodes[i] - nodes[j])
        D[i, i] = -(D[i].sum() - D[i, i])
    return D


class LobattoLegendreBasisGPU:
    """`LobattoLegendreBasisGPU(polydeg)` -- operator set of the reference basis type."""

    def __init__(self, polydeg, RealT=np.float64):
        if RealT is not np.float64:
            raise NotImplementedError("libtrixib200 computes in Float64 (reference API default RealT=Float64)")
        self.polydeg = int(polydeg)
        n = self.polydeg + 1
        self.nodes, self.weights = gauss_lobatto_nodes_weights(n)
        self.inverse_weights = 1.0 / self.weights
        D = polynomial_derivative_matrix(self.nodes)
        self.derivative_matrix = D
        # calc_dhat: Dhat[j, n] = -D[n, j] * w_n / w_j
        self.derivative_dhat = -(D.T * (self.weights[None, :] / self.weights[:, None]))
        ds = 2 * D
        ds[0, 0] += 1 / self.weights[0]
        ds[-1, -1] -= 1 / self.weights[-1]
        self.derivative_split = ds
        self.derivative_split_transpose = ds.T.copy()
        wb = barycentric_weights(self.nodes)
        self.boundary_interpolation = np.stack(
            [lagrange_interpolating_polynomials(-1.0, self.nodes, wb) / self.weights,
             lagrange_interpolating_polynomials(1.0, self.nodes, wb) / self.weights], axis=1)
        V = np.stack([_legendre(m, self.nodes)[0] * np.sqrt(m + 0.5) for m in range(n)], axis=1)
        self.inverse_vandermonde_legendre = np.linalg.inv(V)

    @property
    def nnodes(self):
        return self.polydeg + 1


class MortarL2GPU:
    """`MortarL2GPU(basis)`: forward (interpolation) and reverse (Gauss L2 projection) operators."""

    def __init__(self, basis):
        nodes = basis.nodes
        n = nodes.shape[0]
        wb = barycentric_weights(nodes)
        self.forward_upper = np.stack([lagrange_interpolating_polynomials(0.5 * (x + 1), nodes, wb) for x in nodes])
        self.forward_lower = np.stack([lagrange_interpolating_polynomials(0.5 * (x - 1), nodes, wb) for x in nodes])
        g, gw = gauss_nodes_weights(n)
        gwb = barycentric_weights(g)
        g2l = polynomial_interpolation_matrix(g, nodes)
        l2g = polynomial_interpolation_matrix(nodes, g)
        pu = np.zeros((n, n))
        pl = np.zeros((n, n))
        for j in range(n):
            pu[:, j] = 0.5 * lagrange_interpolating_polynomials(0.5 * (g[j] + 1), g, gwb) * gw[j] / gw
            pl[:, j] = 0.5 * lagrange_interpolating_polynomials(0.5 * (g[j] - 1), g, gwb) * gw[j] / gw
        self.reverse_upper = g2l @ pu @ l2g
        self.reverse_lower = g2l @ pl @ l2g


class SolutionAnalyzer:
    """`SolutionAnalyzer(basis; analysis_polydeg = 2*polydeg)`"""

    def __init__(self, basis, analysis_polydeg=None):
        p = 2 * basis.polydeg if analysis_polydeg is None else analysis_polydeg
        self.nodes, self.weights = gauss_lobatto_nodes_weights(p + 1)
        self.vandermonde = polynomial_interpolation_matrix(basis.nodes, self.nodes)
