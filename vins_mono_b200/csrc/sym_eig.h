// Symmetric eigen-decomposition by Householder tridiagonalisation + implicit-shift QL (the EISPACK tred2 / tql2
// pair, as in JAMA), written once for a cooperative "context": on the device one CTA executes it (strided loops +
// __syncthreads), on the host a single thread does (used by tests/test_sym_eig.py to check this exact code against
// numpy).  This is the algorithm family Eigen's SelfAdjointEigenSolver uses in the reference
// (marginalization_factor.cpp:268, :283): absolute accuracy eps*|A|.  ~100x fewer instructions than cyclic
// Jacobi at n = 75, which is what made marginalisation the slowest kernel of a frame.
//
// In:  V (n x n, leading dimension ld, row-major) holds the symmetric matrix (both triangles).
// Out: d[0..n) eigenvalues (unsorted), V columns = eigenvectors (V[i*ld + k] = component i of eigenvector k).
// Work: e[n], cs[4n] (rotation coefficients, double buffered on the device), scal[16].
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define SE_HD __host__ __device__
#else
#define SE_HD
#endif

namespace vb {

constexpr int SE_PER_LANE_MAX = 5;  // columns per lane in the warp-wide phases of the device context: n <= 160

#if defined(SE_PROF) && defined(__CUDACC__)  // harness/micro/eig_bench.cu: per-phase cycle counters (thread 0)
__shared__ long long se_clk[16];
__host__ __device__ inline long long se_now() {
#if defined(__CUDA_ARCH__)
    return clock64();
#else
    return 0;
#endif
}
__host__ __device__ inline void se_stamp(int k, long long* t0) {
#if defined(__CUDA_ARCH__)
    if (threadIdx.x == 0) {
        const long long t1 = clock64();
        se_clk[k] += t1 - *t0;
        *t0 = t1;
    }
#endif
}
#define SE_T0() long long se_t0_ = se_now()
#define SE_STAMP(k) se_stamp(k, &se_t0_)
#else
#define SE_T0() do {} while (0)
#define SE_STAMP(k) do {} while (0)
#endif

// Execution contexts.  tid/nt: thread index / count for strided loops; sync: barrier over all threads.
// "lead" group: the threads that run the short reductions of the scalar phases (one warp on the device);
// "row groups": grp() adjacent threads share one dot product, grp_sum() adds over the group;
// wid/nw/lane/ws: warp coordinates for the (row, strided column) loops of the rank-k updates.
struct HostCtx {
    static constexpr bool kPipelinedQL = false;
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
    void warp_sync() const {}
    bool is_aux() const { return true; }
    int aux_lane() const { return 0; }
    int aux_size() const { return 1; }
    double aux_sum(double x) const { return x; }
};

#if defined(__CUDACC__)
struct CtaCtx {  // blockDim.x a multiple of 32 and >= 128; named barriers 1..4 are used by the pipelined QL
    static constexpr bool kPipelinedQL = true;
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
    __device__ int grp() const { return 8; }
    __device__ double grp_sum(double x) const {
        x += __shfl_xor_sync(0xffffffffu, x, 1);
        x += __shfl_xor_sync(0xffffffffu, x, 2);
        x += __shfl_xor_sync(0xffffffffu, x, 4);
        return x;
    }
    __device__ int wid() const { return threadIdx.x >> 5; }
    __device__ int nw() const { return blockDim.x >> 5; }
    __device__ int lane() const { return threadIdx.x & 31; }
    __device__ int ws() const { return 32; }
    __device__ void warp_sync() const { __syncwarp(); }
    // aux group: the last warp (idle in the row-group loops when there are more groups than rows)
    __device__ bool is_aux() const { return (threadIdx.x >> 5) == (blockDim.x >> 5) - 1; }
    __device__ int aux_lane() const { return threadIdx.x & 31; }
    __device__ int aux_size() const { return 32; }
    __device__ double aux_sum(double x) const { return lead_sum(x); }
};
#endif

SE_HD inline int max_i(int a, int b) { return a > b ? a : b; }

// One implicit-shift QL sweep on the unreduced block l..m of the tridiagonal (d, e), in two parts.
// ql_head: the shift from the 2x2 block at l; updates d[l], d[l+1], adds the shift to *fshift and returns it in *hs:
// the caller owes d[i] -= *hs for every i >= l+2 before ql_chase (on the device the producer warp does that in
// parallel).  ql_chase: the bulge chase; updates d and e and writes the m - l plane rotations to cs[2i], cs[2i+1]
// (i = m-1 .. l; to be applied to columns i, i+1 of the eigenvector matrix).  Returns whether e[l] is still
// significant.  Single thread.  The chase is the serial bottleneck of the whole decomposition (~n^2/2 dependent
// steps of rsqrt + 5 multiply-adds); its loop is kept to the minimum: walking pointers, operands of step i-1 loaded
// before the dependent chain of step i, no rotating copies (the two old cosines/sines the closing formula needs are
// read back from cs).
SE_HD inline void ql_head(double* d, const double* e, int l, double* fshift, double* hs, double* dl1) {
    const double g = d[l];
    const double p = (d[l + 1] - g) / (2.0 * e[l]);
    double r = sqrt(p * p + 1.0);
    if (p < 0) r = -r;
    d[l] = e[l] / (p + r);
    d[l + 1] = e[l] * (p + r);
    *dl1 = d[l + 1];
    const double h = g - d[l];
    *hs = h;
    *fshift += h;
}

SE_HD inline double se_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
    // MUFU.RSQ64H seed + one third-order Newton step: what rsqrt() does minus its special-case fix-up (x is a sum of
    // two squares, positive and far from the exponent limits here)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = fma(-x, y * y, 1.0);
    return fma(fma(t, 0.375, 0.5), y * t, y);
#else
    return 1.0 / sqrt(x);
#endif
}

SE_HD inline bool ql_chase(double* d, double* e, double* cs, int l, int m, double eps, double tst1, double dl1) {
    double p = d[m];
    double c = 1.0, s = 0.0;
    const double el1 = e[l + 1];
    double* __restrict__ ep = e + m;       // ep[-1] = e[i], ep[0] = e[i+1]
    double* __restrict__ dp = d + m;
    double* __restrict__ cp = cs + 2 * m;  // cp[-2], cp[-1] = cs[2i], cs[2i+1]
    double ei = ep[-1], di = dp[-1];
#pragma unroll 1
    for (int i = m - 1; i >= l; i--) {
        double en = 0.0, dn = 0.0;
        if (i > l) {
            en = ep[-2];
            dn = dp[-2];
        }
        const double s2 = s;
        const double g = c * ei;
        const double h = c * p;
        const double rr = p * p + ei * ei;
        // branch-free: rr = 0 only when p = e[i] = 0 (no rotation: c = 1, s = 0)
        const double ri = se_rsqrt(rr > 1e-290 ? rr : 1e-290);
        const double r = rr * ri;
        s = ei * ri;
        c = rr > 0.0 ? p * ri : 1.0;
        ep[0] = s2 * r;
        p = c * di - s * g;
        dp[0] = h + s * (c * g + s * di);
        cp[-2] = c;
        cp[-1] = s;
        ep--;
        dp--;
        cp -= 2;
        ei = en;
        di = dn;
    }
    const double c3 = (m - l >= 3) ? cs[2 * (l + 2)] : 1.0;
    const double s2 = (m - l >= 2) ? cs[2 * (l + 1) + 1] : 0.0;
    p = -s * s2 * c3 * el1 * e[l] / dl1;
    e[l] = s * p;
    d[l] = c * p;
    return fabs(e[l]) > eps * tst1;
}

#if defined(__CUDACC__)
__device__ __forceinline__ void se_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void se_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// Device form of the QL stage: the scalar recurrence (one thread, inherently serial: ~n^2 dependent steps) and the
// application of its rotations to the eigenvector matrix (n independent rows) run concurrently.  Warp 0 produces the
// rotation sequences sweep by sweep into a double-buffered coefficient array; warps 1..3 (one thread per row) consume
// them.  Hand-over through named barriers: full[b] = 1 + b (producer arrives, consumers wait), empty[b] = 3 + b
// (consumers arrive, producer waits before reusing buffer b).  d and e are touched by the producer only, V by the
// consumers only.  cs holds 4n doubles, scal[9..12] the (l, m) of the sweep in each buffer.  n <= 192 (two rows per
// consumer thread).
template <bool TWO_ROWS>
__device__ inline void ql_pipelined(double* V, int n, int ld, double* d, double* e, double* cs, double* scal) {
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    constexpr int PART = 128;  // producer warp + three consumer warps
    const double eps = 2.220446049250313e-16;
    double* meta = scal + 9;
    if (wid == 0) {
        double f = 0.0, tst1 = 0.0;
        int b = 0;
        bool pend[2] = {false, false};
        SE_T0();
        for (int l = 0; l < n; l++) {
            tst1 = fmax(tst1, fabs(d[l]) + fabs(e[l]));
            int m = n - 1;
            for (int k = l + lane; k < n; k += 32)
                if (fabs(e[k]) <= eps * tst1) {
                    m = k;
                    break;
                }
#pragma unroll
            for (int o = 16; o; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (m > l) {
                for (int iter = 0; iter < 60; iter++) {
                    SE_STAMP(8);   // search for m, bookkeeping
                    if (pend[b]) {
                        se_bar_sync(3 + b, PART);
                        pend[b] = false;
                    }
                    SE_STAMP(9);   // waiting for the consumers
                    double hs = 0.0, dl1 = 0.0;
                    if (lane == 0) ql_head(d, e, l, &f, &hs, &dl1);
                    hs = __shfl_sync(0xffffffffu, hs, 0);
                    for (int i = l + 2 + lane; i < n; i += 32) d[i] -= hs;
                    __syncwarp();
                    int again = 0;
                    if (lane == 0) {
                        again = ql_chase(d, e, cs + b * 2 * n, l, m, eps, tst1, dl1) ? 1 : 0;
                        meta[2 * b] = (double)l;
                        meta[2 * b + 1] = (double)m;
                    }
                    SE_STAMP(10);  // the sweep recurrence
                    again = __shfl_sync(0xffffffffu, again, 0);
                    __syncwarp();
                    __threadfence_block();
                    se_bar_arrive(1 + b, PART);
                    pend[b] = true;
                    b ^= 1;
                    if (!again) break;
                }
            }
            if (lane == 0) {
                d[l] += f;
                e[l] = 0.0;
            }
            __syncwarp();
        }
        if (pend[b]) {  // termination message, then drain the other buffer's release
            se_bar_sync(3 + b, PART);
            pend[b] = false;
        }
        if (lane == 0) meta[2 * b] = -1.0;
        __syncwarp();
        __threadfence_block();
        se_bar_arrive(1 + b, PART);
        b ^= 1;
        if (pend[b]) se_bar_sync(3 + b, PART);
    } else if (wid <= 3) {
        const int k = tid - 32;
        int b = 0;
        for (;;) {
            se_bar_sync(1 + b, PART);
            const int l = (int)meta[2 * b];
            if (l < 0) break;
            const int m = (int)meta[2 * b + 1];
            for (int kr = k; kr < n && kr < (TWO_ROWS ? 192 : 96); kr += 96) {
                const double* c2 = cs + b * 2 * n;
                double* row = V + kr * ld;
                double vi1 = row[m];
                for (int i = m - 1; i >= l; i--) {
                    const double c = c2[2 * i], s = c2[2 * i + 1];
                    const double vi = row[i];
                    row[i + 1] = s * vi + c * vi1;
                    vi1 = c * vi - s * vi1;
                }
                row[l] = vi1;
            }
            __threadfence_block();
            se_bar_arrive(3 + b, PART);
            b ^= 1;
        }
    }
    __syncthreads();
}
#endif

// SE_PER_LANE: columns per lane of the warp-wide phases on the device (n <= 32 * SE_PER_LANE); callers pick the smallest
// instantiation that covers their n (the unrolled per-lane loops cost instructions even when predicated off).
template <class Ctx, int SE_PER_LANE = 3>
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
    // ---- tred2 part 1: Householder reduction, reflector i stored in row i / column i of V.
    // Restructured for few, short barrier phases (3 per step; the textbook order needs 5 and two extra warp reductions):
    //   A  symmetric matvec p = A u with the OLD u (row groups)  ||  aux group: h = |u|^2 and the reflector scalars
    //   B  lead group: correct p for the changed last component of u, scale by 1/h, f = p.u, p -= (f / 2h) u
    //   C  rank-2 update of the lower triangle (a warp per row); the warp owning row i-1 then emits the next u
    // No overflow scaling (EISPACK's scale): the entries here are far from the exponent limits.
    // Scratch inside cs: u double buffer [0,n) with d, sub-diagonal [n,2n), h_i [2n,3n), 1/h_i [3n,4n).
    double* sub = cs + n;
    double* hv = cs + 2 * n;
    double* rhv = cs + 3 * n;
    double* u = d;
    double* un = cs;
    for (int j = tid; j < n; j += nt) u[j] = VV(n - 1, j);
    if (tid == 0) hv[0] = 0.0;
    ctx.sync();
    SE_T0();
    for (int i = n - 1; i > 0; i--) {
        const int npi = (i + 32) / 32;  // per-lane columns in use at this step (j <= i)
        // ---- A
        for (int j0 = 0; j0 < i; j0 += ng) {
            const int j = j0 + gi;
            double g = 0.0, g2 = 0.0;
            if (j < i) {
                // row part (k <= j: contiguous) and column part (k > j: stride ld), two accumulators each for ILP
                const double* rowj = &VV(j, 0);
                int k = gl;
                for (; k + G <= j; k += 2 * G) {
                    g += rowj[k] * u[k];
                    g2 += rowj[k + G] * u[k + G];
                }
                if (k <= j) {
                    g += rowj[k] * u[k];
                    k += G;
                }
                for (; k + G < i; k += 2 * G) {
                    g += VV(k, j) * u[k];
                    g2 += VV(k + G, j) * u[k + G];
                }
                if (k < i) g += VV(k, j) * u[k];
            }
            g = ctx.grp_sum(g + g2);
            if (j < i && gl == 0) e[j] = g;
        }
        if (ctx.is_aux()) {
            double part = 0.0;
            if (ctx.aux_size() == 1) {
                for (int k = 0; k < i; k++) part += u[k] * u[k];
            } else {  // n <= SE_PER_LANE * 32: the loads are issued together
#pragma unroll
                for (int t = 0; t < SE_PER_LANE; t++) {
                    if (t < npi) {
                        const int k = ctx.aux_lane() + 32 * t;
                        const double a = k < i ? u[k] : 0.0;
                        part += a * a;
                    }
                }
            }
            const double hsum = ctx.aux_sum(part);
            if (ctx.aux_lane() == 0) {
                if (hsum == 0.0) {
                    scal[0] = 1.0;  // nothing to annihilate
                    scal[1] = 0.0;
                    sub[i] = 0.0;
                    hv[i] = 0.0;
                } else {
                    const double f = u[i - 1];
                    double g = sqrt(hsum);
                    if (f > 0) g = -g;
                    const double h = hsum - f * g;
                    scal[0] = 0.0;
                    scal[1] = h;
                    scal[10] = -g;     // change of the last component of u
                    scal[11] = f - g;  // its new value
                    sub[i] = g;
                    hv[i] = h;
                }
            }
        }
        ctx.sync();
        SE_STAMP(0);
        if (scal[0] != 0.0) {
            for (int j = tid; j < i; j += nt) {
                un[j] = VV(i - 1, j);
                VV(i, j) = 0.0;
                VV(j, i) = 0.0;
            }
            ctx.sync();
            double* t = u;
            u = un;
            un = t;
            continue;
        }
        // ---- B
        if (tid < LD) {
            const double h = scal[1], rh = 1.0 / h, du = scal[10], ulast = scal[11];
            if (LD == 1) {
                double part = 0.0;
                for (int j = 0; j < i; j++) {
                    const double uj = (j == i - 1) ? ulast : u[j];
                    const double ej = (e[j] + VV(i - 1, j) * du) * rh;
                    e[j] = ej;
                    part += ej * uj;
                }
                const double hh = part * 0.5 * rh;
                for (int j = 0; j < i; j++) {
                    const double uj = (j == i - 1) ? ulast : u[j];
                    e[j] -= hh * uj;
                    u[j] = uj;
                    VV(j, i) = uj;
                }
            } else {  // one warp, n <= SE_PER_LANE * 32: the lane's elements stay in registers, loads issued together
                double uj[SE_PER_LANE], ej[SE_PER_LANE];
                double part = 0.0;
#pragma unroll
                for (int t = 0; t < SE_PER_LANE; t++) {
                    uj[t] = ej[t] = 0.0;
                    if (t < npi) {
                        const int j = tid + 32 * t;
                        uj[t] = j < i ? ((j == i - 1) ? ulast : u[j]) : 0.0;
                        ej[t] = j < i ? (e[j] + VV(i - 1, j) * du) * rh : 0.0;
                        part += ej[t] * uj[t];
                    }
                }
                const double hh = ctx.lead_sum(part) * 0.5 * rh;
#pragma unroll
                for (int t = 0; t < SE_PER_LANE; t++) {
                    const int j = tid + 32 * t;
                    if (t < npi && j < i) {
                        e[j] = ej[t] - hh * uj[t];
                        u[j] = uj[t];
                        VV(j, i) = uj[t];
                    }
                }
            }
        }
        ctx.sync();
        SE_STAMP(1);
        // ---- C
        {
            // each lane keeps its columns' u and e in registers for the whole phase (n <= SE_PER_LANE * 32 on the device)
            double uj[SE_PER_LANE], ej[SE_PER_LANE];
#pragma unroll
            for (int t = 0; t < SE_PER_LANE; t++) {
                const int j = lane + t * ws;
                uj[t] = (t < npi && j < i) ? u[j] : 0.0;
                ej[t] = (t < npi && j < i) ? e[j] : 0.0;
            }
            for (int k = wid; k < i; k += nw) {
                const double ek = e[k], uk = u[k];
                if (ws == 1) {  // host context: plain loop
                    for (int j = 0; j <= k; j++) {
                        const double v = VV(k, j) - (u[j] * ek + e[j] * uk);
                        VV(k, j) = v;
                        if (k == i - 1) un[j] = v;
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < SE_PER_LANE; t++) {
                        const int j = lane + t * ws;
                        if (t < npi && j <= k) {
                            const double v = VV(k, j) - (uj[t] * ek + ej[t] * uk);
                            VV(k, j) = v;
                            if (k == i - 1) un[j] = v;  // row i-1 is the next u
                        }
                    }
                }
            }
        }
        for (int j = tid; j < i; j += nt) VV(i, j) = 0.0;
        ctx.sync();
        SE_STAMP(2);
        double* t = u;
        u = un;
        un = t;
    }
    // ---- tred2 part 2: accumulate the transformations (2 barrier phases per step)
    for (int i = tid; i < n; i += nt) rhv[i] = hv[i] != 0.0 ? 1.0 / hv[i] : 0.0;
    if (tid == 0) {
        VV(n - 1, 0) = VV(0, 0);
        VV(0, 0) = 1.0;
    }
    ctx.sync();
    for (int i = 0; i < n - 1; i++) {
        const int npi = (i + 32) / 32;
        const bool active = hv[i + 1] != 0.0;
        if (active) {
            const double rh = rhv[i + 1];
            for (int j0 = 0; j0 <= i; j0 += ng) {  // g_j = (sum_k V[k][i+1] V[k][j]) / h
                const int j = j0 + gi;
                double g = 0.0;
                double g2 = 0.0;
                if (j <= i) {
                    int k = gl;
                    for (; k + G <= i; k += 2 * G) {
                        g += VV(k, i + 1) * VV(k, j);
                        g2 += VV(k + G, i + 1) * VV(k + G, j);
                    }
                    if (k <= i) g += VV(k, i + 1) * VV(k, j);
                }
                g = ctx.grp_sum(g + g2);
                if (j <= i && gl == 0) e[j] = g * rh;
            }
            ctx.sync();
            SE_STAMP(6);
        }
        double ej2[SE_PER_LANE];
#pragma unroll
        for (int t = 0; t < SE_PER_LANE; t++) {
            const int j = lane + t * ws;
            ej2[t] = (t < npi && active && j <= i) ? e[j] : 0.0;
        }
        for (int k = wid; k <= i; k += nw) {
            if (active) {
                const double c = VV(k, i + 1);
                if (ws == 1) {
                    for (int j = 0; j <= i; j++) VV(k, j) -= e[j] * c;
                } else {
#pragma unroll
                    for (int t = 0; t < SE_PER_LANE; t++) {
                        const int j = lane + t * ws;
                        if (t < npi && j <= i) VV(k, j) -= ej2[t] * c;
                    }
                }
            }
            ctx.warp_sync();
            if (lane == 0) VV(k, i + 1) = 0.0;
        }
        if (tid == 0 && i + 1 < n - 1) {  // prepare the next step: save the diagonal entry in the last row
            VV(n - 1, i + 1) = VV(i + 1, i + 1);
            VV(i + 1, i + 1) = 1.0;
        }
        ctx.sync();
        SE_STAMP(7);
    }
    for (int j = tid; j < n; j += nt) {
        d[j] = VV(n - 1, j);
        VV(n - 1, j) = 0.0;
    }
    ctx.sync();
    if (tid == 0) VV(n - 1, n - 1) = 1.0;
    // sub-diagonal for tql2: e[i-1] = e_tred2[i], e[n-1] = 0
    for (int i = tid; i < n; i += nt) e[i] = (i + 1 < n) ? sub[i + 1] : 0.0;
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
#if defined(__CUDA_ARCH__)
    if constexpr (Ctx::kPipelinedQL) {
        ql_pipelined<(SE_PER_LANE > 3)>(V, n, ld, d, e, cs, scal);
        if (tid == 0) {
            scal[4] = (double)(clk1 - clk0);
            scal[5] = (double)(clock64() - clk1);
        }
        ctx.sync();
        return;
    }
#endif
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
                    double hs, dl1;
                    ql_head(d, e, l, &scal[2], &hs, &dl1);
                    for (int i = l + 2; i < n; i++) d[i] -= hs;
                    scal[7] = ql_chase(d, e, cs, l, m, eps, tst1, dl1) ? 1.0 : 0.0;
                }
                ctx.sync();
                const bool again = scal[7] != 0.0;
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
