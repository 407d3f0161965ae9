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

#include "../../vins_mono_b200/csrc/prior_floor.h"
// The prior's eigenvalue floor (prior_floor.h), single-threaded.  In/out: A (n x n) -> A+, g (n) -> g0; c0, stats[2].
extern "C" void host_prior_floor(double* A, double* g, int n, double eps, int force_full, double* c0, int* stats) {
    const int ld = n | 1;
    std::vector<double> V(static_cast<size_t>(n) * ld), d(n), e(n), cs(4 * n), scal(16), tv(n), Ev(static_cast<size_t>(n) * ld);
    std::vector<double> work(vb::prior_floor_work(n, 1, 1));
    vb::prior_floor<vb::HostCtx, 3>(vb::HostCtx(), A, g, c0, n, eps, V.data(), ld, d.data(), e.data(), cs.data(), scal.data(), tv.data(),
                                    force_full ? nullptr : work.data(), Ev.data(), stats);
}
