// ahx_eig.cc -- top eigenvector of the sparse strandedness matrix O for flipAll.
//
// The reference factorises the dense N x N O with gonum mat64.EigenSym
// (orientation.go:33-44, O(N^3) in pure Go) and only uses the sign pattern of
// the eigenvector of the LARGEST eigenvalue.  O has E << N^2 nonzeros, so we
// run Lanczos with full reorthogonalisation on the sparse form (restarted from
// the current Ritz vector until the residual is small).  The global sign of an
// eigenvector is arbitrary in both implementations; flipWhole settles it.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "ahx_internal.h"

namespace ahx {
namespace {

void spmv(int32_t n, int64_t E, const int32_t *pa, const int32_t *pb, const double *w, const double *x, double *y) {
  std::fill(y, y + n, 0.0);
  for (int64_t e = 0; e < E; ++e) {
    y[pa[e]] += w[e] * x[pb[e]];
    y[pb[e]] += w[e] * x[pa[e]];
  }
}
double dot(int32_t n, const double *a, const double *b) { double s = 0; for (int32_t i = 0; i < n; ++i) s += a[i] * b[i]; return s; }

// Dense symmetric eigen-decomposition by cyclic Jacobi (k is at most a few hundred).
void jacobi(std::vector<double> &A, std::vector<double> &V, int k) {
  V.assign(static_cast<size_t>(k) * k, 0.0);
  for (int i = 0; i < k; ++i) V[static_cast<size_t>(i) * k + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0, diag = 0;
    for (int i = 0; i < k; ++i) { diag += A[static_cast<size_t>(i) * k + i] * A[static_cast<size_t>(i) * k + i]; for (int j = i + 1; j < k; ++j) off += A[static_cast<size_t>(i) * k + j] * A[static_cast<size_t>(i) * k + j]; }
    if (off <= 1e-30 * (diag + 1e-300)) break;
    for (int p = 0; p < k; ++p)
      for (int q = p + 1; q < k; ++q) {
        const double apq = A[static_cast<size_t>(p) * k + q];
        if (std::fabs(apq) < 1e-300) continue;
        const double theta = (A[static_cast<size_t>(q) * k + q] - A[static_cast<size_t>(p) * k + p]) / (2 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
        const double c = 1 / std::sqrt(t * t + 1), s = t * c;
        for (int r = 0; r < k; ++r) {
          const double arp = A[static_cast<size_t>(r) * k + p], arq = A[static_cast<size_t>(r) * k + q];
          A[static_cast<size_t>(r) * k + p] = c * arp - s * arq; A[static_cast<size_t>(r) * k + q] = s * arp + c * arq;
        }
        for (int r = 0; r < k; ++r) {
          const double apr = A[static_cast<size_t>(p) * k + r], aqr = A[static_cast<size_t>(q) * k + r];
          A[static_cast<size_t>(p) * k + r] = c * apr - s * aqr; A[static_cast<size_t>(q) * k + r] = s * apr + c * aqr;
        }
        for (int r = 0; r < k; ++r) {
          const double vrp = V[static_cast<size_t>(r) * k + p], vrq = V[static_cast<size_t>(r) * k + q];
          V[static_cast<size_t>(r) * k + p] = c * vrp - s * vrq; V[static_cast<size_t>(r) * k + q] = s * vrp + c * vrq;
        }
      }
  }
}

}  // namespace

double top_eigenvector(int32_t n, int64_t E, const int32_t *pa, const int32_t *pb, const double *w, double *v) {
  std::fill(v, v + n, 0.0);
  if (n <= 0 || E <= 0) return 0.0;
  const int m = std::min<int32_t>(n, 200);
  std::vector<double> Q(static_cast<size_t>(m + 1) * n), alpha(m), beta(m), r(n), T, S;
  // deterministic start vector with a component along (almost surely) every eigenvector
  uint64_t lcg = 0x2545F4914F6CDD1DULL;
  for (int32_t i = 0; i < n; ++i) { lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL; v[i] = 0.5 + static_cast<double>(lcg >> 11) * (1.0 / 9007199254740992.0); }
  double lambda = 0.0, anorm = 0.0;
  for (int64_t e = 0; e < E; ++e) anorm = std::max(anorm, std::fabs(w[e]));
  for (int restart = 0; restart < 50; ++restart) {
    double nv = std::sqrt(dot(n, v, v));
    if (nv == 0) { v[0] = 1; nv = 1; }
    for (int32_t i = 0; i < n; ++i) Q[i] = v[i] / nv;
    int k = 0;
    for (int j = 0; j < m; ++j) {
      double *qj = &Q[static_cast<size_t>(j) * n];
      spmv(n, E, pa, pb, w, qj, r.data());
      alpha[j] = dot(n, r.data(), qj);
      for (int pass = 0; pass < 2; ++pass)   // full reorthogonalisation, twice
        for (int i = 0; i <= j; ++i) {
          const double *qi = &Q[static_cast<size_t>(i) * n];
          const double c = dot(n, r.data(), qi);
          for (int32_t t = 0; t < n; ++t) r[t] -= c * qi[t];
        }
      beta[j] = std::sqrt(dot(n, r.data(), r.data()));
      k = j + 1;
      if (beta[j] <= 1e-12 * (anorm * n + 1e-300) || j + 1 == m) break;
      double *qn = &Q[static_cast<size_t>(j + 1) * n];
      for (int32_t t = 0; t < n; ++t) qn[t] = r[t] / beta[j];
    }
    T.assign(static_cast<size_t>(k) * k, 0.0);
    for (int i = 0; i < k; ++i) {
      T[static_cast<size_t>(i) * k + i] = alpha[i];
      if (i + 1 < k) { T[static_cast<size_t>(i) * k + i + 1] = beta[i]; T[static_cast<size_t>(i + 1) * k + i] = beta[i]; }
    }
    jacobi(T, S, k);
    int top = 0;
    for (int i = 1; i < k; ++i) if (T[static_cast<size_t>(i) * k + i] > T[static_cast<size_t>(top) * k + top]) top = i;
    lambda = T[static_cast<size_t>(top) * k + top];
    std::fill(v, v + n, 0.0);
    for (int i = 0; i < k; ++i) {
      const double s = S[static_cast<size_t>(i) * k + top];
      const double *qi = &Q[static_cast<size_t>(i) * n];
      for (int32_t t = 0; t < n; ++t) v[t] += s * qi[t];
    }
    const double resid = std::fabs(beta[k - 1] * S[static_cast<size_t>(k - 1) * k + top]);
    if (k == n || resid <= 1e-10 * (std::fabs(lambda) + 1e-300)) break;
  }
  return lambda;
}

}  // namespace ahx
