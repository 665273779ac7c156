// TEST INFRASTRUCTURE ONLY -- CPU oracle for the TreeMesh DGSEM rhs! hot path.
// Parity status: "parity unpinned" (Trixi.jl is not vendored under /root/reference and Julia is
// not installed; every formula here is a restatement of Trixi.jl <= 0.13 semantics, see
// SURVEY.md Appendix A and oracle/ORACLE_ASSUMPTIONS.md).
//
// Lobatto-Legendre basis + mortar operators, restating what the reference pulls out of Trixi at
//   /root/reference/src/solvers/basis_lobatto_legendre.jl:60-100  (nodes, weights, dhat, dsplit, ...)
//   /root/reference/src/solvers/basis_lobatto_legendre.jl:155-173 (MortarL2GPU: forward/reverse)
//   /root/reference/src/solvers/basis_lobatto_legendre.jl:116-132 (SolutionAnalyzer)
// All matrices are stored COLUMN-MAJOR (Julia): M(i,j) = m[i + n*j].
#pragma once
#include <cmath>
#include <vector>
#include <cassert>
#include <algorithm>

namespace orc {

using vec = std::vector<double>;

struct Mat {  // column-major dense matrix
  int r = 0, c = 0;
  vec a;
  Mat() {}
  Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
  double& operator()(int i, int j) { return a[i + (size_t)r * j]; }
  double operator()(int i, int j) const { return a[i + (size_t)r * j]; }
};

inline Mat matmul(const Mat& A, const Mat& B) {
  Mat C(A.r, B.c);
  for (int j = 0; j < B.c; ++j)
    for (int i = 0; i < A.r; ++i) {
      double s = 0;
      for (int k = 0; k < A.c; ++k) s += A(i, k) * B(k, j);
      C(i, j) = s;
    }
  return C;
}

// Kopriva, Algorithm 24: q = L_{N+1} - L_{N-1}, q', and L_N
inline void calc_q_and_l(int N, double x, double& q, double& qder, double& L) {
  double L_Nm2 = 1.0, L_Nm1 = x, Lder_Nm2 = 0.0, Lder_Nm1 = 1.0;
  double L_N = 0, Lder_N = 0;
  for (int i = 2; i <= N; ++i) {
    L_N = ((2 * i - 1) * x * L_Nm1 - (i - 1) * L_Nm2) / i;
    Lder_N = Lder_Nm2 + (2 * i - 1) * L_Nm1;
    L_Nm2 = L_Nm1; L_Nm1 = L_N;
    Lder_Nm2 = Lder_Nm1; Lder_Nm1 = Lder_N;
  }
  q = (2 * N + 1) / double(N + 1) * (x * L_N - L_Nm2);
  qder = (2 * N + 1) * L_N;
  L = L_N;
}

// Trixi `gauss_lobatto_nodes_weights(n_nodes)` (Kopriva Alg. 25)
inline void gauss_lobatto_nodes_weights(int n_nodes, vec& nodes, vec& weights) {
  const int n_iterations = 20;
  const double tolerance = 2 * 2.220446049250313e-16;
  nodes.assign(n_nodes, 0.0);
  weights.assign(n_nodes, 0.0);
  int N = n_nodes - 1;
  if (N == 0) { nodes[0] = 0; weights[0] = 2; return; }
  if (N == 1) { nodes = {-1.0, 1.0}; weights = {1.0, 1.0}; return; }
  nodes[0] = -1.0; weights[0] = 2.0 / (N * (N + 1));
  nodes[N] = 1.0; weights[N] = weights[0];
  for (int j = 1; j <= (N + 1) / 2 - 1; ++j) {
    double x = -std::cos(M_PI * ((j + 0.25) / N - 3.0 / (8 * N * M_PI * (j + 0.25))));
    double q, qder, L;
    for (int k = 0; k < n_iterations; ++k) {
      calc_q_and_l(N, x, q, qder, L);
      double dx = -q / qder;
      x += dx;
      if (std::fabs(dx) < tolerance * std::fabs(x)) break;
    }
    calc_q_and_l(N, x, q, qder, L);
    nodes[j] = x;
    weights[j] = weights[0] / (L * L);
    nodes[N - j] = -x;
    weights[N - j] = weights[j];
  }
  if (N % 2 == 0) {
    double q, qder, L;
    calc_q_and_l(N, 0.0, q, qder, L);
    nodes[N / 2] = 0.0;
    weights[N / 2] = weights[0] / (L * L);
  }
}

// Legendre polynomial (normalized, Kopriva Alg. 22 scaled by sqrt(N+1/2)) and derivative
inline void legendre_polynomial_and_derivative(int N, double x, double& poly, double& deriv) {
  if (N == 0) { poly = 1.0; deriv = 0.0; }
  else if (N == 1) { poly = x; deriv = 1.0; }
  else {
    double p2 = 1.0, p1 = x, d2 = 0.0, d1 = 1.0;
    poly = 0; deriv = 0;
    for (int i = 2; i <= N; ++i) {
      poly = ((2 * i - 1) * x * p1 - (i - 1) * p2) / i;
      deriv = d2 + (2 * i - 1) * p1;
      p2 = p1; p1 = poly; d2 = d1; d1 = deriv;
    }
  }
  double s = std::sqrt(N + 0.5);
  poly *= s; deriv *= s;
}

// Trixi `gauss_nodes_weights(n_nodes)` (Kopriva Alg. 23)
inline void gauss_nodes_weights(int n_nodes, vec& nodes, vec& weights) {
  const int n_iterations = 20;
  const double tolerance = 2 * 2.220446049250313e-16;
  nodes.assign(n_nodes, 0.0);
  weights.assign(n_nodes, 0.0);
  int N = n_nodes - 1;
  if (N == 0) { nodes[0] = 0; weights[0] = 2; return; }
  if (N == 1) {
    nodes = {-std::sqrt(1.0 / 3.0), std::sqrt(1.0 / 3.0)};
    weights = {1.0, 1.0};
    return;
  }
  for (int j = 0; j <= (N + 1) / 2 - 1; ++j) {
    double x = -std::cos(M_PI * (2 * j + 1) / (2 * N + 2));
    double poly, deriv;
    for (int k = 0; k < n_iterations; ++k) {
      legendre_polynomial_and_derivative(N + 1, x, poly, deriv);
      double dx = -poly / deriv;
      x += dx;
      if (std::fabs(dx) < tolerance * std::fabs(x)) break;
    }
    legendre_polynomial_and_derivative(N + 1, x, poly, deriv);
    nodes[j] = x;
    weights[j] = (2 * N + 3) / ((1 - x * x) * deriv * deriv);
    nodes[N - j] = -x;
    weights[N - j] = weights[j];
  }
  if (N % 2 == 0) {
    double poly, deriv;
    legendre_polynomial_and_derivative(N + 1, 0.0, poly, deriv);
    nodes[N / 2] = 0.0;
    weights[N / 2] = (2 * N + 3) / (deriv * deriv);
  }
}

inline vec barycentric_weights(const vec& nodes) {
  int n = (int)nodes.size();
  vec w(n, 1.0);
  for (int j = 1; j < n; ++j)
    for (int k = 0; k < j; ++k) {
      w[k] *= nodes[k] - nodes[j];
      w[j] *= nodes[j] - nodes[k];
    }
  for (int j = 0; j < n; ++j) w[j] = 1.0 / w[j];
  return w;
}

// Kopriva Alg. 37
inline Mat polynomial_derivative_matrix(const vec& nodes) {
  int n = (int)nodes.size();
  vec wb = barycentric_weights(nodes);
  Mat D(n, n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      if (j != i) {
        D(i, j) = wb[j] / wb[i] * 1.0 / (nodes[i] - nodes[j]);
        D(i, i) -= D(i, j);
      }
  return D;
}

inline bool isapprox(double a, double b) {  // Julia isapprox default rtol = sqrt(eps)
  return std::fabs(a - b) <= 1.4901161193847656e-08 * std::max(std::fabs(a), std::fabs(b));
}

// Kopriva Alg. 34
inline vec lagrange_interpolating_polynomials(double x, const vec& nodes, const vec& wbary) {
  int n = (int)nodes.size();
  vec poly(n, 0.0);
  for (int i = 0; i < n; ++i)
    if (isapprox(x, nodes[i])) { poly[i] = 1.0; return poly; }
  double total = 0;
  for (int i = 0; i < n; ++i) { poly[i] = wbary[i] / (x - nodes[i]); total += poly[i]; }
  for (int i = 0; i < n; ++i) poly[i] /= total;
  return poly;
}

// Kopriva Alg. 32: interpolation matrix nodes_in -> nodes_out, (n_out x n_in)
inline Mat polynomial_interpolation_matrix(const vec& nodes_in, const vec& nodes_out) {
  int ni = (int)nodes_in.size(), no = (int)nodes_out.size();
  vec wb = barycentric_weights(nodes_in);
  Mat V(no, ni);
  for (int k = 0; k < no; ++k) {
    bool match = false;
    for (int j = 0; j < ni; ++j)
      if (isapprox(nodes_out[k], nodes_in[j])) { match = true; V(k, j) = 1.0; }
    if (!match) {
      double s = 0;
      for (int j = 0; j < ni; ++j) {
        double t = wb[j] / (nodes_out[k] - nodes_in[j]);
        V(k, j) = t; s += t;
      }
      for (int j = 0; j < ni; ++j) V(k, j) /= s;
    }
  }
  return V;
}

inline Mat invert(const Mat& A) {  // Gauss-Jordan with partial pivoting
  int n = A.r;
  Mat M = A, I(n, n);
  for (int i = 0; i < n; ++i) I(i, i) = 1.0;
  for (int c = 0; c < n; ++c) {
    int p = c;
    for (int r = c + 1; r < n; ++r) if (std::fabs(M(r, c)) > std::fabs(M(p, c))) p = r;
    if (p != c) for (int j = 0; j < n; ++j) { std::swap(M(c, j), M(p, j)); std::swap(I(c, j), I(p, j)); }
    double d = M(c, c);
    for (int j = 0; j < n; ++j) { M(c, j) /= d; I(c, j) /= d; }
    for (int r = 0; r < n; ++r) if (r != c) {
      double f = M(r, c);
      if (f != 0.0) for (int j = 0; j < n; ++j) { M(r, j) -= f * M(c, j); I(r, j) -= f * I(c, j); }
    }
  }
  return I;
}

struct Basis {
  int N = 0;  // nnodes = polydeg + 1
  vec nodes, weights, inverse_weights;
  Mat D, Dhat, Dsplit, Dsplit_transpose, boundary_interpolation /* N x 2 */, inverse_vandermonde_legendre;
  Mat forward_upper, forward_lower, reverse_upper, reverse_lower;
  // analyzer (2*polydeg+1 LGL nodes)
  int NA = 0;
  vec analysis_nodes, analysis_weights;
  Mat analysis_vandermonde;  // NA x N

  explicit Basis(int polydeg) {
    N = polydeg + 1;
    gauss_lobatto_nodes_weights(N, nodes, weights);
    inverse_weights.resize(N);
    for (int i = 0; i < N; ++i) inverse_weights[i] = 1.0 / weights[i];
    D = polynomial_derivative_matrix(nodes);
    // calc_dhat: dhat = D^T; dhat[j,n] *= -w[n]/w[j]   => Dhat(j,n) = -D(n,j)*w_n/w_j
    Dhat = Mat(N, N);
    for (int n = 0; n < N; ++n)
      for (int j = 0; j < N; ++j) Dhat(j, n) = -(D(n, j) * (weights[n] / weights[j]));
    // calc_dsplit: 2D, corners corrected by -/+ 1/w
    Dsplit = Mat(N, N);
    for (int j = 0; j < N; ++j) for (int i = 0; i < N; ++i) Dsplit(i, j) = 2 * D(i, j);
    Dsplit(0, 0) += 1 / weights[0];
    Dsplit(N - 1, N - 1) -= 1 / weights[N - 1];
    Dsplit_transpose = Mat(N, N);
    for (int j = 0; j < N; ++j) for (int i = 0; i < N; ++i) Dsplit_transpose(i, j) = Dsplit(j, i);
    // boundary_interpolation[:,1] = calc_lhat(-1), [:,2] = calc_lhat(+1): l_i(x)/w_i
    vec wb = barycentric_weights(nodes);
    boundary_interpolation = Mat(N, 2);
    vec lm = lagrange_interpolating_polynomials(-1.0, nodes, wb);
    vec lp = lagrange_interpolating_polynomials(1.0, nodes, wb);
    for (int i = 0; i < N; ++i) {
      boundary_interpolation(i, 0) = lm[i] / weights[i];
      boundary_interpolation(i, 1) = lp[i] / weights[i];
    }
    // vandermonde_legendre + inverse
    Mat V(N, N);
    for (int i = 0; i < N; ++i)
      for (int m = 0; m < N; ++m) {
        double p, d;
        legendre_polynomial_and_derivative(m, nodes[i], p, d);
        V(i, m) = p;
      }
    inverse_vandermonde_legendre = invert(V);
    // mortar operators
    forward_upper = Mat(N, N); forward_lower = Mat(N, N);
    for (int j = 0; j < N; ++j) {
      vec pu = lagrange_interpolating_polynomials(0.5 * (nodes[j] + 1), nodes, wb);
      vec pl = lagrange_interpolating_polynomials(0.5 * (nodes[j] - 1), nodes, wb);
      for (int i = 0; i < N; ++i) { forward_upper(j, i) = pu[i]; forward_lower(j, i) = pl[i]; }
    }
    vec gn, gw;
    gauss_nodes_weights(N, gn, gw);
    vec gwb = barycentric_weights(gn);
    Mat Pu(N, N), Pl(N, N);
    for (int j = 0; j < N; ++j) {
      vec pu = lagrange_interpolating_polynomials(0.5 * (gn[j] + 1), gn, gwb);
      vec pl = lagrange_interpolating_polynomials(0.5 * (gn[j] - 1), gn, gwb);
      for (int i = 0; i < N; ++i) {
        Pu(i, j) = 0.5 * pu[i] * gw[j] / gw[i];
        Pl(i, j) = 0.5 * pl[i] * gw[j] / gw[i];
      }
    }
    Mat g2l = polynomial_interpolation_matrix(gn, nodes);
    Mat l2g = polynomial_interpolation_matrix(nodes, gn);
    reverse_upper = matmul(matmul(g2l, Pu), l2g);
    reverse_lower = matmul(matmul(g2l, Pl), l2g);
    // analyzer
    NA = 2 * polydeg + 1;
    gauss_lobatto_nodes_weights(NA, analysis_nodes, analysis_weights);
    analysis_vandermonde = polynomial_interpolation_matrix(nodes, analysis_nodes);
  }
};

}  // namespace orc
