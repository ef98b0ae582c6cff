// ahx_internal.h -- declarations shared by the translation units of libahx.
#ifndef AHX_INTERNAL_H_
#define AHX_INTERNAL_H_

#include <cstdint>
#include <string>

#include "ahx.h"

namespace ahx {

void set_global_error(const std::string &s);
const char *global_error();

double go_log(double x);
void golden_and_sumlog(const int64_t *d, int64_t n, int32_t *bins, double *sumlog);

struct ClmView {
  int32_t n; const int64_t *sizes; int64_t E;
  const int32_t *pa, *pb, *nl; const int8_t *strand; const int32_t *slots;
  int64_t H; const int32_t *hist; bool finalized;
  const double *diag;   // O[a][a] = strandedness * nlinks of self-pair CLM lines (orientation.go:124-132), or nullptr
};
ClmView view_of(const ahx_clm *clm);

// Largest-eigenvalue eigenvector of the sparse symmetric matrix
// O[a][b] = O[b][a] = w[e], O[a][a] = diag[a] (orientation.go:124-132; diag may be null): Lanczos with
// full reorthogonalisation, stopped as soon as the top Ritz pair has converged; dense Jacobi for small
// groups and as the fallback when Lanczos does not converge.  v: n doubles.  Returns the Ritz value.
struct EigInfo { bool converged = false, dense = false; double residual = 0.0; int restarts = 0, matvecs = 0; };
double top_eigenvector(int32_t n, int64_t E, const int32_t *pa, const int32_t *pb, const double *w, const double *diag, double *v,
                       EigInfo *info = nullptr);

}  // namespace ahx

#endif  // AHX_INTERNAL_H_
