// Symmetric eigen-decomposition by Householder tridiagonalisation + implicit-shift QL (the EISPACK tred2 / tql2
// pair, as in JAMA), written once for a cooperative "context": on the device one CTA executes it (strided loops +
// __syncthreads), on the host a single thread does (used by tests/test_sym_eig.py to check this exact code against
// numpy).  This is the algorithm family Eigen's SelfAdjointEigenSolver uses in the reference
// (marginalization_factor.cpp:268, :283): absolute accuracy eps*|A|.  ~100x fewer instructions than cyclic
// Jacobi at n = 75, which is what made marginalisation the slowest kernel of a frame.
//
// In:  V (n x n, leading dimension ld, row-major) holds the symmetric matrix (both triangles).
// Out: d[0..n) eigenvalues (unsorted), V columns = eigenvectors (V[i*ld + k] = component i of eigenvector k).
// Work: e[n], cs[2n] (rotation coefficients of one QL sweep), scal[16].
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define SE_HD __host__ __device__
#else
#define SE_HD
#endif

namespace vb {

// Execution contexts.  tid/nt: thread index / count for strided loops; sync: barrier over all threads.
// "lead" group: the threads that run the short reductions of the scalar phases (one warp on the device);
// "row groups": grp() adjacent threads share one dot product, grp_sum() adds over the group;
// wid/nw/lane/ws: warp coordinates for the (row, strided column) loops of the rank-k updates.
struct HostCtx {
    int tid() const { return 0; }
    int nt() const { return 1; }
    void sync() const {}
    int lead() const { return 1; }
    double lead_sum(double x) const { return x; }
    int lead_min(int x) const { return x; }
    void lead_sync() const {}
    int grp() const { return 1; }
    double grp_sum(double x) const { return x; }
    int wid() const { return 0; }
    int nw() const { return 1; }
    int lane() const { return 0; }
    int ws() const { return 1; }
};

#if defined(__CUDACC__)
struct CtaCtx {  // blockDim.x a multiple of 32
    __device__ int tid() const { return threadIdx.x; }
    __device__ int nt() const { return blockDim.x; }
    __device__ void sync() const { __syncthreads(); }
    __device__ int lead() const { return 32; }
    __device__ double lead_sum(double x) const {
#pragma unroll
        for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        return x;
    }
    __device__ int lead_min(int x) const {
#pragma unroll
        for (int o = 16; o; o >>= 1) x = min(x, __shfl_xor_sync(0xffffffffu, x, o));
        return x;
    }
    __device__ void lead_sync() const { __syncwarp(); }
    __device__ int grp() const { return 4; }
    __device__ double grp_sum(double x) const {
        x += __shfl_xor_sync(0xffffffffu, x, 1);
        x += __shfl_xor_sync(0xffffffffu, x, 2);
        return x;
    }
    __device__ int wid() const { return threadIdx.x >> 5; }
    __device__ int nw() const { return blockDim.x >> 5; }
    __device__ int lane() const { return threadIdx.x & 31; }
    __device__ int ws() const { return 32; }
};
#endif

SE_HD inline int max_i(int a, int b) { return a > b ? a : b; }

template <class Ctx>
SE_HD void sym_eig(Ctx ctx, double* V, int n, int ld, double* d, double* e, double* cs, double* scal) {
    const int tid = ctx.tid(), nt = ctx.nt();
    const int LD = ctx.lead(), G = ctx.grp();
    const int wid = ctx.wid(), nw = ctx.nw(), lane = ctx.lane(), ws = ctx.ws();
    const int gi = tid / G, gl = tid - gi * G, ng = nt / G;  // row-group coordinates
#define VV(i, j) V[(i) * ld + (j)]
    if (n == 1) {
        if (tid == 0) {
            d[0] = V[0];
            V[0] = 1.0;
        }
        ctx.sync();
        return;
    }
#if defined(__CUDA_ARCH__)
    const long long clk0 = clock64();
#endif
    // ---- tred2 part 1: Householder reduction, reflector i stored in row i / column i of V, sub-diagonal in cs[n + i]
    for (int j = tid; j < n; j += nt) d[j] = VV(n - 1, j);
    ctx.sync();
    for (int i = n - 1; i > 0; i--) {
        if (tid < LD) {  // scalar phase on the lead group
            double part = 0.0;
            for (int k = tid; k < i; k += LD) part += fabs(d[k]);
            const double scale = ctx.lead_sum(part);
            if (scale == 0.0) {
                if (tid == 0) {
                    cs[n + i] = d[i - 1];
                    scal[1] = 0.0;
                    scal[0] = 0.0;
                }
            } else {
                part = 0.0;
                for (int k = tid; k < i; k += LD) {
                    const double v = d[k] / scale;
                    d[k] = v;
                    part += v * v;
                }
                double h = ctx.lead_sum(part);
                ctx.lead_sync();
                if (tid == 0) {
                    const double f = d[i - 1];
                    double g = sqrt(h);
                    if (f > 0) g = -g;
                    cs[n + i] = scale * g;
                    h = h - f * g;
                    d[i - 1] = f - g;
                    scal[1] = h;
                    scal[0] = scale;
                }
            }
        }
        ctx.sync();
        if (scal[0] == 0.0) {
            for (int j = tid; j < i; j += nt) {
                d[j] = VV(i - 1, j);
                VV(i, j) = 0.0;
                VV(j, i) = 0.0;
            }
            if (tid == 0) d[i] = 0.0;
            ctx.sync();
            continue;
        }
        const double h = scal[1];
        // e[0..i) = A u: A symmetric, stored in the lower triangle of V (rows/cols < i), u = d; a group per row
        for (int j0 = 0; j0 < i; j0 += ng) {
            const int j = j0 + gi;
            double g = 0.0;
            if (j < i)
                for (int k = gl; k < i; k += G) g += (k <= j ? VV(j, k) : VV(k, j)) * d[k];
            g = ctx.grp_sum(g);
            if (j < i && gl == 0) {
                e[j] = g;
                VV(j, i) = d[j];
            }
        }
        ctx.sync();
        if (tid < LD) {
            double part = 0.0;
            for (int j = tid; j < i; j += LD) {
                const double v = e[j] / h;
                e[j] = v;
                part += v * d[j];
            }
            const double f = ctx.lead_sum(part);
            const double hh = f / (h + h);
            for (int j = tid; j < i; j += LD) e[j] -= hh * d[j];
        }
        ctx.sync();
        // rank-2 update of the lower triangle: V[k][j] -= d[j] e[k] + e[j] d[k],  j <= k < i
        for (int k = wid; k < i; k += nw) {
            const double ek = e[k], dk = d[k];
            for (int j = lane; j <= k; j += ws) VV(k, j) -= (d[j] * ek + e[j] * dk);
        }
        ctx.sync();
        for (int j = tid; j < i; j += nt) {
            d[j] = VV(i - 1, j);
            VV(i, j) = 0.0;
        }
        if (tid == 0) d[i] = h;
        ctx.sync();
    }
    // ---- tred2 part 2: accumulate the transformations (e is scratch here)
    for (int i = 0; i < n - 1; i++) {
        if (tid == 0) {
            VV(n - 1, i) = VV(i, i);
            VV(i, i) = 1.0;
        }
        const double h = d[i + 1];
        if (h != 0.0) {
            for (int k = tid; k <= i; k += nt) d[k] = VV(k, i + 1) / h;
            ctx.sync();
            for (int j0 = 0; j0 <= i; j0 += ng) {  // g_j = sum_k V[k][i+1] V[k][j]
                const int j = j0 + gi;
                double g = 0.0;
                if (j <= i)
                    for (int k = gl; k <= i; k += G) g += VV(k, i + 1) * VV(k, j);
                g = ctx.grp_sum(g);
                if (j <= i && gl == 0) e[j] = g;
            }
            ctx.sync();
            for (int k = wid; k <= i; k += nw) {
                const double dk = d[k];
                for (int j = lane; j <= i; j += ws) VV(k, j) -= e[j] * dk;
            }
        }
        for (int k = tid; k <= i; k += nt) VV(k, i + 1) = 0.0;
        ctx.sync();
    }
    for (int j = tid; j < n; j += nt) {
        d[j] = VV(n - 1, j);
        VV(n - 1, j) = 0.0;
    }
    ctx.sync();
    if (tid == 0) VV(n - 1, n - 1) = 1.0;
    // sub-diagonal for tql2: e[i-1] = e_tred2[i], e[n-1] = 0
    for (int i = tid; i < n; i += nt) e[i] = (i + 1 < n) ? cs[n + i + 1] : 0.0;
    if (tid == 0) {
        scal[2] = 0.0;  // f: accumulated shift
        scal[3] = 0.0;  // tst1
    }
    ctx.sync();
#if defined(__CUDA_ARCH__)
    const long long clk1 = clock64();
#endif
    // ---- tql2: implicit-shift QL on (d, e), rotations accumulated into V
    const double eps = 2.220446049250313e-16;
    for (int l = 0; l < n; l++) {
        if (tid < LD) {  // tst1 and the first negligible sub-diagonal element at or after l
            const double tst1 = fmax(scal[3], fabs(d[l]) + fabs(e[l]));
            int mm = n - 1;  // e[n-1] = 0 always qualifies
            for (int k = l + tid; k < n; k += LD)
                if (fabs(e[k]) <= eps * tst1) {
                    mm = k;
                    break;
                }
            mm = ctx.lead_min(mm);
            ctx.lead_sync();
            if (tid == 0) {
                scal[3] = tst1;
                scal[6] = (double)mm;
            }
        }
        ctx.sync();
        const double tst1 = scal[3];
        const int m = (int)scal[6];
        if (m > l) {
            for (int iter = 0; iter < 60; iter++) {
                // scalar recurrence of one QL sweep on thread 0; rotation coefficients go to cs[2i], cs[2i+1]
                if (tid == 0) {
                    double g = d[l];
                    double p = (d[l + 1] - g) / (2.0 * e[l]);
                    double r = sqrt(p * p + 1.0);
                    if (p < 0) r = -r;
                    d[l] = e[l] / (p + r);
                    d[l + 1] = e[l] * (p + r);
                    const double dl1 = d[l + 1];
                    double h = g - d[l];
                    const double hs = h;  // shift of this sweep: d[i] -= hs for i >= l+2, applied lazily
                    scal[8] = hs;
                    scal[2] += h;
                    p = m >= l + 2 ? d[m] - hs : d[m];
                    double c = 1.0, c2 = c, c3 = c;
                    const double el1 = e[l + 1];
                    double s = 0.0, s2 = 0.0;
                    for (int i = m - 1; i >= l; i--) {
                        c3 = c2;
                        c2 = c;
                        s2 = s;
                        const double ei = e[i], di = i >= l + 2 ? d[i] - hs : d[i];
                        g = c * ei;
                        h = c * p;
                        const double rr = p * p + ei * ei;
                        if (rr > 0.0) {
#if defined(__CUDA_ARCH__)
                            const double ri = rsqrt(rr);
#else
                            const double ri = 1.0 / sqrt(rr);
#endif
                            r = rr * ri;
                            s = ei * ri;
                            c = p * ri;
                        } else {
                            r = 0.0;
                            s = 0.0;
                            c = 1.0;
                        }
                        e[i + 1] = s2 * r;
                        p = c * di - s * g;
                        d[i + 1] = h + s * (c * g + s * di);
                        cs[2 * i] = c;
                        cs[2 * i + 1] = s;
                    }
                    p = -s * s2 * c3 * el1 * e[l] / dl1;
                    e[l] = s * p;
                    d[l] = c * p;
                    scal[7] = fabs(e[l]) > eps * tst1 ? 1.0 : 0.0;
                }
                ctx.sync();
                const bool again = scal[7] != 0.0;
                const double hs_all = scal[8];
                for (int i = max_i(m + 1, l + 2) + tid; i < n; i += nt) d[i] -= hs_all;  // entries the sweep did not touch
                for (int k = tid; k < n; k += nt) {
                    double vi1 = VV(k, m);
                    for (int i = m - 1; i >= l; i--) {
                        const double c = cs[2 * i], s = cs[2 * i + 1];
                        const double vi = VV(k, i);
                        VV(k, i + 1) = s * vi + c * vi1;
                        vi1 = c * vi - s * vi1;
                    }
                    VV(k, l) = vi1;
                }
                ctx.sync();
                if (!again) break;
            }
        }
        if (tid == 0) {
            d[l] = d[l] + scal[2];
            e[l] = 0.0;
        }
        ctx.sync();
    }
#if defined(__CUDA_ARCH__)
    if (tid == 0) {  // phase cycle counters for profiling: tridiagonalisation, QL
        scal[4] = (double)(clk1 - clk0);
        scal[5] = (double)(clock64() - clk1);
    }
    ctx.sync();
#endif
#undef VV
}

}  // namespace vb
