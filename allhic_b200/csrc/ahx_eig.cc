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

void spmv(int32_t n, int64_t E, const int32_t *pa, const int32_t *pb, const double *w, const double *diag, const double *x, double *y) {
  if (diag) for (int32_t i = 0; i < n; ++i) y[i] = diag[i] * x[i];
  else std::fill(y, y + n, 0.0);
  for (int64_t e = 0; e < E; ++e) {
    y[pa[e]] += w[e] * x[pb[e]];
    y[pb[e]] += w[e] * x[pa[e]];
  }
}
// four independent accumulators: the compiler may not reassociate a plain reduction loop
double dot(int32_t n, const double *a, const double *b) {
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  int32_t i = 0;
  for (; i + 4 <= n; i += 4) { s0 += a[i] * b[i]; s1 += a[i + 1] * b[i + 1]; s2 += a[i + 2] * b[i + 2]; s3 += a[i + 3] * b[i + 3]; }
  for (; i < n; ++i) s0 += a[i] * b[i];
  return (s0 + s1) + (s2 + s3);
}

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

// Largest eigenvalue of the symmetric tridiagonal matrix (alpha[0..k), beta[0..k-1)) by bisection on the
// Sturm count, and its eigenvector by inverse iteration (tridiagonal elimination with partial pivoting,
// as LAPACK's dgtsv): O(k) per step instead of a dense O(k^3) decomposition at every convergence test.
int sturm_below(int k, const double *alpha, const double *beta, double x) {   // eigenvalues < x
  int cnt = 0;
  double q = 1.0;
  for (int i = 0; i < k; ++i) {
    const double b2 = i ? beta[i - 1] * beta[i - 1] : 0.0;
    q = alpha[i] - x - (i ? b2 / q : 0.0);
    if (q == 0.0) q = 1e-300;
    if (q < 0) ++cnt;
  }
  return cnt;
}
void tri_solve(int k, const double *alpha, const double *beta, double mu, std::vector<double> &b) {
  std::vector<double> d(k), dl(std::max(k - 1, 1)), du(std::max(k - 1, 1)), du2(std::max(k - 2, 1), 0.0);
  for (int i = 0; i < k; ++i) d[i] = alpha[i] - mu;
  for (int i = 0; i + 1 < k; ++i) dl[i] = du[i] = beta[i];
  for (int i = 0; i + 1 < k; ++i) {
    if (std::fabs(d[i]) >= std::fabs(dl[i])) {
      if (d[i] == 0.0) d[i] = 1e-300;
      const double f = dl[i] / d[i];
      d[i + 1] -= f * du[i];
      b[i + 1] -= f * b[i];
      if (i + 2 < k) du2[i] = 0.0;
    } else {   // swap rows i and i + 1
      const double f = d[i] / dl[i];
      d[i] = dl[i];
      const double t = d[i + 1];
      d[i + 1] = du[i] - f * t;
      if (i + 2 < k) { du2[i] = du[i + 1]; du[i + 1] = -f * du2[i]; }
      du[i] = t;
      const double tb = b[i];
      b[i] = b[i + 1];
      b[i + 1] = tb - f * b[i + 1];
    }
  }
  if (d[k - 1] == 0.0) d[k - 1] = 1e-300;
  b[k - 1] /= d[k - 1];
  if (k > 1) b[k - 2] = (b[k - 2] - du[k - 2] * b[k - 1]) / d[k - 2];
  for (int i = k - 3; i >= 0; --i) b[i] = (b[i] - du[i] * b[i + 1] - du2[i] * b[i + 2]) / d[i];
}
double tri_top(int k, const double *alpha, const double *beta, std::vector<double> &s) {
  double lo = alpha[0], hi = alpha[0];
  for (int i = 0; i < k; ++i) {
    const double r = (i ? std::fabs(beta[i - 1]) : 0.0) + (i + 1 < k ? std::fabs(beta[i]) : 0.0);
    lo = std::min(lo, alpha[i] - r); hi = std::max(hi, alpha[i] + r);
  }
  for (int it = 0; it < 200 && hi - lo > 4e-16 * std::max(std::fabs(lo), std::fabs(hi)); ++it) {
    const double mid = 0.5 * (lo + hi);
    if (mid <= lo || mid >= hi) break;
    if (sturm_below(k, alpha, beta, mid) >= k) hi = mid; else lo = mid;   // all k eigenvalues below mid: the top one is too
  }
  const double lambda = 0.5 * (lo + hi);
  s.assign(k, 0.0);
  for (int i = 0; i < k; ++i) s[i] = 1.0 / std::sqrt(static_cast<double>(k)) * (1.0 + 0.01 * ((i * 37) % 11));
  const double mu = lambda + 1e-13 * (std::fabs(lambda) + 1e-300);   // just above the eigenvalue: the shifted matrix is definite and nearly singular
  for (int it = 0; it < 4; ++it) {
    tri_solve(k, alpha, beta, mu, s);
    double nrm = 0;
    for (double x : s) nrm += x * x;
    nrm = std::sqrt(nrm);
    if (!(nrm > 0) || !std::isfinite(nrm)) { s.assign(k, 0.0); s[0] = 1.0; break; }
    for (double &x : s) x /= nrm;
  }
  return lambda;
}

}  // namespace

// Dense path: the whole spectrum by Jacobi (what the reference's mat64.EigenSym does), for small groups
// and as the fallback when Lanczos does not converge.
static double top_eigenvector_dense(int32_t n, int64_t E, const int32_t *pa, const int32_t *pb, const double *w, const double *diag, double *v) {
  std::vector<double> A(static_cast<size_t>(n) * n, 0.0), V;
  for (int64_t e = 0; e < E; ++e) { A[static_cast<size_t>(pa[e]) * n + pb[e]] = w[e]; A[static_cast<size_t>(pb[e]) * n + pa[e]] = w[e]; }
  if (diag) for (int32_t i = 0; i < n; ++i) A[static_cast<size_t>(i) * n + i] = diag[i];
  jacobi(A, V, n);
  int top = 0;
  for (int i = 1; i < n; ++i) if (A[static_cast<size_t>(i) * n + i] > A[static_cast<size_t>(top) * n + top]) top = i;
  for (int32_t i = 0; i < n; ++i) v[i] = V[static_cast<size_t>(i) * n + top];
  return A[static_cast<size_t>(top) * n + top];
}

double top_eigenvector(int32_t n, int64_t E, const int32_t *pa, const int32_t *pb, const double *w, const double *diag, double *v, EigInfo *info) {
  std::fill(v, v + n, 0.0);
  EigInfo local;
  if (!info) info = &local;
  *info = EigInfo{};
  if (n <= 0 || (E <= 0 && !diag)) return 0.0;
  constexpr int kDense = 48, kDenseFallback = 1500;
  if (n <= kDense) { info->dense = true; info->converged = true; return top_eigenvector_dense(n, E, pa, pb, w, diag, v); }
  const int m = std::min<int32_t>(n, 200);
  std::vector<double> Q(static_cast<size_t>(m + 1) * n), alpha(m), beta(m), r(n), S;
  // deterministic start vector with a component along (almost surely) every eigenvector
  uint64_t lcg = 0x2545F4914F6CDD1DULL;
  for (int32_t i = 0; i < n; ++i) { lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL; v[i] = 0.5 + static_cast<double>(lcg >> 11) * (1.0 / 9007199254740992.0); }
  double lambda = 0.0, anorm = 0.0;
  for (int64_t e = 0; e < E; ++e) anorm = std::max(anorm, std::fabs(w[e]));
  if (diag) for (int32_t i = 0; i < n; ++i) anorm = std::max(anorm, std::fabs(diag[i]));
  // Ritz pair of the current k-step factorisation: top eigenvalue of T, residual |beta_k s_k|
  auto ritz = [&](int k) {
    lambda = tri_top(k, alpha.data(), beta.data(), S);
    return std::fabs(beta[k - 1] * S[k - 1]);
  };
  const double tol = 1e-10;
  for (int restart = 0; restart < 50; ++restart) {
    info->restarts = restart + 1;
    double nv = std::sqrt(dot(n, v, v));
    if (nv == 0) { v[0] = 1; nv = 1; }
    for (int32_t i = 0; i < n; ++i) Q[i] = v[i] / nv;
    int k = 0;
    double resid = 0.0;
    for (int j = 0; j < m; ++j) {
      double *qj = &Q[static_cast<size_t>(j) * n];
      spmv(n, E, pa, pb, w, diag, qj, r.data());
      info->matvecs++;
      alpha[j] = dot(n, r.data(), qj);
      // full reorthogonalisation; a second pass only when the first one cancelled most of r (Kahan / Parlett: "twice is enough")
      for (int pass = 0; pass < 2; ++pass) {
        const double before = std::sqrt(dot(n, r.data(), r.data()));
        for (int i = 0; i <= j; ++i) {
          const double *qi = &Q[static_cast<size_t>(i) * n];
          const double c = dot(n, r.data(), qi);
          for (int32_t t = 0; t < n; ++t) r[t] -= c * qi[t];
        }
        if (std::sqrt(dot(n, r.data(), r.data())) > 0.7 * before) break;
      }
      beta[j] = std::sqrt(dot(n, r.data(), r.data()));
      k = j + 1;
      const bool last = beta[j] <= 1e-12 * (anorm * n + 1e-300) || j + 1 == m;
      // the Ritz residual is cheap to test (a k x k problem): stop as soon as the top pair has converged
      if (last || (k >= 16 && k % 8 == 0)) {
        resid = ritz(k);
        if (last || resid <= tol * (std::fabs(lambda) + 1e-300)) break;
      }
      double *qn = &Q[static_cast<size_t>(j + 1) * n];
      for (int32_t t = 0; t < n; ++t) qn[t] = r[t] / beta[j];
    }
    std::fill(v, v + n, 0.0);
    for (int i = 0; i < k; ++i) {
      const double s = S[i];
      const double *qi = &Q[static_cast<size_t>(i) * n];
      for (int32_t t = 0; t < n; ++t) v[t] += s * qi[t];
    }
    info->residual = resid / (std::fabs(lambda) + 1e-300);
    if (k == n || resid <= tol * (std::fabs(lambda) + 1e-300)) { info->converged = true; break; }
  }
  // near-degenerate top eigenvalues / disconnected contig graphs: take the dense decomposition when it is affordable
  if (!info->converged && n <= kDenseFallback) { info->dense = true; info->converged = true; info->residual = 0.0; return top_eigenvector_dense(n, E, pa, pb, w, diag, v); }
  return lambda;
}

}  // namespace ahx
