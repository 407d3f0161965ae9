// The eigenvalue floor of the marginalisation prior (marginalization_factor.cpp:268-297 in the reference: A' = V S V^T,
// S+ = S where S > eps else 0, linearized_jacobians = sqrt(S+) V^T, linearized_residuals = sqrt(S+)^-1 V^T b').  The
// solve consumes the prior in information form: A+ = V S+ V^T, g0 = V 1+ V^T b' and c0 = b'^T V S+^-1 V^T b' (what
// J^T J, J^T r and r^T r of the reference's factor are).  Only the part of the spectrum at or below the floor has to be
// SEPARATED for that:
//     A+ = A' - sum_dropped w_k v_k v_k^T,      g0 = b' - sum_dropped v_k (v_k^T b'),
//     c0 = sum_explicit,kept (v_k^T b')^2 / w_k + g_hi^T M^-1 g_hi,
// where "explicit" are the eigenpairs below tau = max(1e-12 |T|, 4 eps) (se_small_eigs: bisection + inverse iteration on
// the tridiagonal form, no QL rotation chain), g_hi = b' minus its explicit components and M = A' with the explicit
// eigenvalues replaced by sigma = trace / n, which makes it positive definite with condition <= 1e12: one LDL^T pass with
// g_hi carried as an extra row yields g_hi^T M^-1 g_hi.  Falls back to the full decomposition (se_finish, then the sums
// over all eigenpairs) when more than SE_KMAX eigenvalues lie below tau, a vector fails its residual check or a pivot of
// M is not positive; both routes implement the same definition and agree to eps |A'|.
// Written for the same cooperative contexts as sym_eig.h (tests/test_sym_eig.py runs this code single-threaded).
#pragma once
#include "sym_eig.h"

namespace vb {

// Doubles of scratch behind `work` (partial route): the LDL^T copy M (n x (n|1)), X (SE_KMAX x n), se_small_eigs' work,
// projections / eigenvalues / flags, the LDL^T entry table (one int per entry of the lower triangle).
SE_HD inline int prior_floor_work(int n, int nt, int ws) {
    return n * (n | 1) + SE_KMAX * n + se_small_work(n, SE_KMAX, nt, ws) + 2 * SE_KMAX + 8 + (n * (n + 1) / 2 + 1) / 2;
}

// In:  Ap (n x n, row-major) = A' as assembled (not exactly symmetric), g (n) = b'.
// Out: Ap = A+, g = g0, *c0; stats[0] = explicit pairs (-1: full decomposition ran), stats[1] = dropped pairs.
// V: (n x ld) scratch for the decomposition; d, e (n), cs (4n), scal (16), tv (n): small arrays; work: prior_floor_work()
// doubles, or null to force the full decomposition (Ev: n x ld scratch for its recomposition, may alias work).
template <class Ctx, int SE_PER_LANE>
SE_HD void prior_floor(Ctx ctx, double* Ap, double* g, double* c0, int n, double eps, double* V, int ld, double* d,
                       double* e, double* cs, double* scal, double* tv, double* work, double* Ev, int* stats) {
    const int tid = ctx.tid(), nt = ctx.nt(), LD = ctx.lead();
    const int wid = ctx.wid(), nw = ctx.nw(), lane = ctx.lane(), ws = ctx.ws();
    int k = -1;
    const int ldm = n | 1;
    double* M = work;
    for (int idx = tid; idx < n * n; idx += nt) {  // the symmetrised A': V for the decomposition, M for the partial route
        const int i = idx / n, j = idx - i * n;
        const double v = 0.5 * (Ap[i * n + j] + Ap[j * n + i]);
        V[i * ld + j] = v;
        if (M) M[i * ldm + j] = v;
    }
    ctx.sync();
    double* X = work ? M + (size_t)n * ldm : nullptr;
    double* sw = work ? X + (size_t)SE_KMAX * n : nullptr;
    double* xb = work ? sw + se_small_work(n, SE_KMAX, nt, ws) : nullptr;  // x_k . b'
    double* lam = work ? xb + SE_KMAX : nullptr;
    double* fl = work ? lam + SE_KMAX : nullptr;  // [0] pivot failure, [1] corner, [2] sigma, [3..6] pivot | 1/pivot, double buffered
    if (n >= 2) se_tridiag<Ctx, SE_PER_LANE>(ctx, V, n, ld, d, e, cs, scal);
    if (work && n >= 2) k = se_small_eigs<Ctx, SE_PER_LANE>(ctx, V, n, ld, d, e, cs, 1e-12, 4.0 * eps, SE_KMAX, lam, X, sw);
    SE_T0();
    if (k >= 0) {
        // projections of b' on the explicit vectors; sigma
        for (int j = wid; j < k; j += nw) {
            double s = 0.0;
            for (int i = lane; i < n; i += ws) s += X[(size_t)j * n + i] * g[i];
            s = ctx.wsum(s);
            if (lane == 0) xb[j] = s;
        }
        if (tid == 0) {
            double tr = 0.0;
            for (int i = 0; i < n; i++) tr += d[i];
            fl[0] = 0.0;
            fl[1] = 0.0;
            fl[2] = tr > 0.0 ? tr / n : 1.0;
        }
        ctx.sync();
        const double sigma = fl[2];
        // A+ = sym(A') - sum_dropped w_k x_k x_k^T (final);  M = sym(A') + sum_explicit (sigma - w_k) x_k x_k^T;
        // g <- g0, tv = g_hi, b' kept in cs[0, n) for the fall-back
        for (int idx = tid; idx < n * n; idx += nt) {
            const int i = idx / n, j = idx - i * n;
            const double v = M[i * ldm + j];
            double ap = v, mm = v;
            for (int t = 0; t < k; t++) {
                const double xx = X[(size_t)t * n + i] * X[(size_t)t * n + j];
                mm += (sigma - lam[t]) * xx;
                if (!(lam[t] > eps)) ap -= lam[t] * xx;
            }
            Ap[idx] = ap;
            M[i * ldm + j] = mm;
        }
        for (int i = tid; i < n; i += nt) {
            const double b = g[i];
            double hi = b, g0 = b;
            for (int t = 0; t < k; t++) {
                const double c = X[(size_t)t * n + i] * xb[t];
                hi -= c;
                if (!(lam[t] > eps)) g0 -= c;
            }
            cs[i] = b;
            tv[i] = hi;
            g[i] = g0;
        }
        ctx.sync();
        // LDL^T of [M g_hi; g_hi^T 0] without pivoting: the corner ends as -g_hi^T M^-1 g_hi.  One barrier per column; the
        // thread that finishes the next pivot publishes it and its reciprocal (double buffered) so that the division is off
        // the other threads' path.
        if (tid == 0) {
            fl[3] = M[0];
            fl[4] = 1.0 / M[0];
        }
        ctx.sync();
        // Entries of the trailing triangle are dealt to the threads once, enumerated from the fixed corner (n-1, n-1)
        // outwards (reversed row/column r' <= c', index c'(c'+1)/2 + r'): entry idx is live at column j exactly when
        // idx < m(m+1)/2, so a thread's entries and their (row, column) never change and the last live one is the next pivot.
        const int E = n * (n + 1) / 2;
        int* tab = reinterpret_cast<int*>(fl + 8);  // packed row << 16 | column, one int per entry
        for (int idx = tid; idx < E; idx += nt) {
            int cp = (int)((sqrt(8.0 * (double)idx + 1.0) - 1.0) * 0.5);
            while (cp * (cp + 1) / 2 > idx) cp--;
            while ((cp + 1) * (cp + 2) / 2 <= idx) cp++;
            const int rp = idx - cp * (cp + 1) / 2;
            tab[idx] = ((n - 1 - rp) << 16) | (n - 1 - cp);
        }
        for (int j = 0; j < n; j++) {
            const int pb = 3 + 2 * (j & 1), nb = 3 + 2 * ((j + 1) & 1);
            const double p = fl[pb];
            if (!(p > 0.0) || !(p < 1e300)) {
                if (tid == 0) fl[0] = 1.0;
                break;  // every thread reads the same p
            }
            const double ip = fl[pb + 1];
            const int m = n - 1 - j, live = m * (m + 1) / 2;
            const double zj = tv[j];
            for (int idx = tid; idx < live; idx += nt) {
                const int rc = tab[idx], row = rc >> 16, col = rc & 0xffff;
                const double v = M[row * ldm + col] - M[row * ldm + j] * ip * M[col * ldm + j];
                M[row * ldm + col] = v;
                if (idx == live - 1) {  // (j+1, j+1): the next pivot is final
                    fl[nb] = v;
                    fl[nb + 1] = 1.0 / v;
                }
            }
            for (int r = tid; r < m; r += nt) tv[j + 1 + r] -= M[(j + 1 + r) * ldm + j] * ip * zj;
            if (tid == 0) fl[1] += zj * zj * ip;
            ctx.sync();
        }
        ctx.sync();
        SE_STAMP(14);  // A+, g0, M, LDL^T
        if (fl[0] != 0.0) {  // a pivot of M was not positive: restore b', the full decomposition rewrites A+ entirely
            for (int i = tid; i < n; i += nt) g[i] = cs[i];
            ctx.sync();
            k = -1;
        }
    }
    if (k >= 0) {
        if (tid == 0) {
            int dropped = 0;
            double cex = 0.0;
            for (int t = 0; t < k; t++) {
                if (lam[t] > eps)
                    cex += xb[t] * xb[t] / lam[t];
                else
                    dropped++;
            }
            c0[0] = cex + fl[1];
            if (stats) {
                stats[0] = k;
                stats[1] = dropped;
            }
        }
        ctx.sync();
        return;
    }
    // ---- full decomposition: A+ = V diag(w+) V^T, g0 = V 1+ V^T b', c0 = sum_kept (v_k^T b')^2 / w_k
    if (n >= 2)
        se_finish<Ctx, SE_PER_LANE>(ctx, V, n, ld, d, e, cs, scal);
    else {
        if (tid == 0) {
            d[0] = V[0];
            V[0] = 1.0;
        }
        ctx.sync();
    }
    for (int t = tid; t < n; t += nt) {
        double s = 0;
        for (int i = 0; i < n; i++) s += V[i * ld + t] * g[i];
        tv[t] = s;  // v_t^T b'
    }
    ctx.sync();
    if (tid < LD) {
        double cpart = 0.0;
        int dropped = 0;
        for (int t = tid; t < n; t += LD) {
            if (d[t] > eps)
                cpart += tv[t] * tv[t] / d[t];
            else
                dropped++;
        }
        cpart = ctx.lead_sum(cpart);
        const double dr = ctx.lead_sum((double)dropped);
        if (tid == 0) {
            c0[0] = cpart;
            if (stats) {
                stats[0] = -1;
                stats[1] = (int)dr;
            }
        }
    }
    for (int t = tid; t < n; t += nt) e[t] = d[t] > eps ? d[t] : 0.0;  // e is free after the decomposition
    ctx.sync();
    if (Ev) {  // room for a scaled copy of V: one multiply less in the n^3 product
        for (int idx = tid; idx < n * n; idx += nt) {
            const int i = idx / n, t = idx - i * n;
            Ev[i * ld + t] = V[i * ld + t] * e[t];
        }
        ctx.sync();
        for (int idx = tid; idx < n * n; idx += nt) {
            const int i = idx / n, j = idx - i * n;
            double s = 0;
            for (int t = 0; t < n; t++) s += Ev[i * ld + t] * V[j * ld + t];
            Ap[idx] = s;
        }
    } else {
        for (int idx = tid; idx < n * n; idx += nt) {
            const int i = idx / n, j = idx - i * n;
            double s = 0;
            for (int t = 0; t < n; t++) s += V[i * ld + t] * e[t] * V[j * ld + t];
            Ap[idx] = s;
        }
    }
    for (int i = tid; i < n; i += nt) {
        double s = 0;
        for (int t = 0; t < n; t++)
            if (d[t] > eps) s += V[i * ld + t] * tv[t];
        cs[i] = s;  // cs is free as well
    }
    ctx.sync();
    for (int i = tid; i < n; i += nt) g[i] = cs[i];
    ctx.sync();
}

}  // namespace vb
