// Test-only host instantiation of the product's cooperative eigen-solver (vins_mono_b200/csrc/sym_eig.h) so the
// exact code the marginalisation kernel runs can be checked against numpy on a box without a GPU.
#include "../../vins_mono_b200/csrc/sym_eig.h"
#include <vector>
extern "C" void host_sym_eig(const double* A, int n, double* evals, double* evecs) {
    const int ld = n | 1;
    std::vector<double> V(static_cast<size_t>(n) * ld), e(n), cs(4 * n), scal(16);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) V[i * ld + j] = A[i * n + j];
    vb::sym_eig(vb::HostCtx(), V.data(), n, ld, evals, e.data(), cs.data(), scal.data());
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) evecs[i * n + j] = V[i * ld + j];
}
