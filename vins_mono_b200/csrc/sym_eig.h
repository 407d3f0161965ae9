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
    double wsum(double x) const { return x; }
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
    __device__ double wsum(double x) const { return lead_sum(x); }  // any full warp
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
#define VV(i, j) V[(i) * ld + (j)]
// Stage 1 (n >= 2): Householder tridiagonalisation.  Leaves the diagonal of T in VV(i, i), the sub-diagonal T(i, i-1) in
// cs[n + i], the reflector scalars h_i in cs[2n + i] and reflector i (components 0 .. i-1) in column i of V; d, e and
// cs[0, n) are scratch afterwards.  se_finish or se_small_eigs continue from this state.
template <class Ctx, int SE_PER_LANE = 3>
SE_HD void se_tridiag(Ctx ctx, double* V, int n, int ld, double* d, double* e, double* cs, double* scal) {
    const int tid = ctx.tid(), nt = ctx.nt();
    const int LD = ctx.lead(), G = ctx.grp();
    const int wid = ctx.wid(), nw = ctx.nw(), lane = ctx.lane(), ws = ctx.ws();
    const int gi = tid / G, gl = tid - gi * G, ng = nt / G;  // row-group coordinates
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
#if defined(__CUDA_ARCH__)
    if (tid == 0) scal[4] = (double)(clock64() - clk0);
#endif
}

// Stage 2: accumulation of the reflectors into V (tred2 part 2) and implicit-shift QL with the rotations applied to V.
template <class Ctx, int SE_PER_LANE = 3>
SE_HD void se_finish(Ctx ctx, double* V, int n, int ld, double* d, double* e, double* cs, double* scal) {
    const int tid = ctx.tid(), nt = ctx.nt();
    const int LD = ctx.lead(), G = ctx.grp();
    const int wid = ctx.wid(), nw = ctx.nw(), lane = ctx.lane(), ws = ctx.ws();
    const int gi = tid / G, gl = tid - gi * G, ng = nt / G;  // row-group coordinates
    double* sub = cs + n;
    double* hv = cs + 2 * n;
    double* rhv = cs + 3 * n;
    (void)LD;
    SE_T0();
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
    if (tid == 0) scal[5] = (double)(clock64() - clk1);  // phase cycle counter for profiling
    ctx.sync();
#endif
}

// ------------------------------------------------------------------------------------------------
// The eigenpairs below a threshold only (the marginalisation prior needs nothing else of its decomposition: the kept part
// of the spectrum never has to be separated from the matrix, see prior_floor.h).  LAPACK's dstebz / dstein scheme on the
// tridiagonal form: Sturm-sequence counts + multisection for the eigenvalues (every thread evaluates one abscissa per
// round), inverse iteration with partial-pivoting LU of T - lambda I for the vectors, joint modified Gram-Schmidt (the
// eigenvalues at the noise floor eps |T| form one cluster), Rayleigh quotients, back-transformation through the stored
// reflectors.  None of it has the O(n^2)-step serial rotation chain of QL.
constexpr int SE_KMAX = 16;

// Eigenvalues of the tridiagonal (td, te2 = squared sub-diagonal) below x: negative pivots of the LDL^T of T - x I.
SE_HD inline int se_sturm(const double* td, const double* te2, int n, double x, double pivmin) {
    double q = td[0] - x;
    if (fabs(q) <= pivmin) q = -pivmin;
    int c = q < 0.0 ? 1 : 0;
    for (int i = 1; i < n; i++) {
        q = (td[i] - x) - te2[i - 1] * (1.0 / q);
        if (fabs(q) <= pivmin) q = -pivmin;
        c += q < 0.0 ? 1 : 0;
    }
    return c;
}

// doubles of scratch se_small_eigs needs in `work` for n, kmax and nt threads (ws = threads per warp of the context)
SE_HD inline int se_small_work(int n, int kmax, int nt, int ws) {
    const int c = nt > kmax * ws ? nt : kmax * ws;
    return 16 + 3 * kmax + c + kmax * 5 * n;
}

// Continues from se_tridiag.  tau = max(tau_rel * |T|, tau_min).  Returns k >= 0: lam[j] (Rayleigh quotients, ascending
// up to noise) and X[j * n .. +n) = orthonormal eigenvectors of the ORIGINAL matrix for all eigenvalues below tau; or -1 when
// there are more than kmax (<= SE_KMAX) of them or a vector failed its residual check.  V and cs[n, 3n) are left intact
// either way (se_finish can still run).  d, e hold T afterwards; work[0] = |T| (max row sum), work[1] = tau.
template <class Ctx, int SE_PER_LANE = 3>
SE_HD int se_small_eigs(Ctx ctx, const double* V, int n, int ld, double* d, double* e, double* cs, double tau_rel,
                        double tau_min, int kmax, double* lam, double* X, double* work) {
    const int tid = ctx.tid(), nt = ctx.nt();
    const int LD = ctx.lead();
    const int wid = ctx.wid(), nw = ctx.nw(), lane = ctx.lane(), ws = ctx.ws();
    const double ulp = 2.220446049250313e-16;
    const double* sub = cs + n;
    const double* hv = cs + 2 * n;
    double* te2 = cs;
    double* S = work;
    double* lohi = work + 16;
    double* lhat = lohi + 2 * kmax;
    double* cnt = lhat + kmax;
    double* fac = cnt + (nt > kmax * ws ? nt : kmax * ws);
    for (int i = tid; i < n; i += nt) {
        d[i] = V[i * ld + i];
        const double ei = i + 1 < n ? sub[i + 1] : 0.0;
        e[i] = ei;
        te2[i] = ei * ei;
    }
    ctx.sync();
    if (tid == 0) {
        double tn = 0.0, gl = d[0], emax = 0.0;
        for (int i = 0; i < n; i++) {
            const double r = fabs(e[i]) + (i > 0 ? fabs(e[i - 1]) : 0.0);
            tn = fmax(tn, fabs(d[i]) + r);
            gl = fmin(gl, d[i] - r);
            emax = fmax(emax, te2[i]);
        }
        S[0] = tn;
        S[1] = fmax(tau_rel * tn, tau_min);
        S[2] = gl - 2.0 * ulp * tn * n - 1e-300;
        S[6] = 2.2250738585072014e-308 * fmax(1.0, emax);
        S[5] = 0.0;  // failure flag
    }
    ctx.sync();
    const double tnorm = S[0], tau = S[1], pivmin = S[6];
    const double atol = 2.0 * ulp * tnorm;
    SE_T0();
    const int k = se_sturm(d, te2, n, tau, pivmin);  // every thread: the same arithmetic, the same answer
    if (k == 0) return 0;
    if (k > kmax) return -1;
    // ---- hull of the k wanted eigenvalues in ONE round: every thread evaluates one abscissa of a geometric ladder (ratio
    // 2^c per rung) from -|T| up to -atol/4 and from +atol/4 up to tau, so the hull comes out within one rung of the
    // extreme wanted eigenvalues whatever their magnitude (skipped by narrow contexts: bisection below does it all)
    double lo = S[2], hi = tau;
    if (nt >= 64) {
        const int half = nt / 2;
        const double xmin = 0.25 * atol;
        double x;
        if (tid < half) {
            const double c = log2(tnorm / xmin) / (double)(half - 1);
            x = -tnorm * exp2(-c * (double)tid);
        } else {
            const double c = tau > xmin ? log2(tau / xmin) / (double)(nt - half - 1) : 0.0;
            x = tau * exp2(-c * (double)(nt - 1 - tid));
        }
        const int c = se_sturm(d, te2, n, x, pivmin);
        cnt[tid] = (double)c;
        if (tid == 0) {
            S[3] = lo;
            S[4] = hi;
        }
        ctx.sync();
        if (c <= 0 && (tid == nt - 1 || cnt[tid + 1] > 0.0)) S[3] = x;
        if (c >= k && (tid == 0 || cnt[tid - 1] < (double)k)) S[4] = x;
        ctx.sync();
        lo = S[3];
        hi = S[4];
        ctx.sync();
    }
    // ---- eigenvalue j (0-based from below): (ws + 1)-section by warp j
    for (int j = wid; j < k; j += nw) {
        double lj = lo, hj = hi;
        double* wc = cnt + (j % nw) * ws;
        for (int round = 0; round < 128 && hj - lj > atol; round++) {
            const double x = lj + (hj - lj) * ((double)(lane + 1) / (double)(ws + 1));
            const int c = se_sturm(d, te2, n, x, pivmin);
            wc[lane] = (double)c;
            if (lane == 0) {
                lohi[2 * j] = lj;
                lohi[2 * j + 1] = hj;
            }
            ctx.warp_sync();
            if (c <= j && (lane == ws - 1 || wc[lane + 1] > (double)j)) lohi[2 * j] = x;
            if (c > j && (lane == 0 || wc[lane - 1] <= (double)j)) lohi[2 * j + 1] = x;
            ctx.warp_sync();
            lj = lohi[2 * j];
            hj = lohi[2 * j + 1];
            ctx.warp_sync();
        }
        if (lane == 0) lhat[j] = 0.5 * (lj + hj);
    }
    ctx.sync();
    SE_STAMP(11);  // bisection
    if (tid == 0) {  // keep the shifts of a cluster apart (dstein's perturbation)
        const double sep = 4.0 * ulp * tnorm;
        for (int j = 1; j < k; j++)
            if (lhat[j] - lhat[j - 1] < sep) lhat[j] = lhat[j - 1] + sep;
    }
    ctx.sync();
    // ---- inverse iteration: P L U = T - lhat[j] I (dlagtf), then U^-1 L^-1 P applied three times (dlagts with perturbed pivots)
    const double tiny = ulp * tnorm;
    for (int j = wid; j < k; j += nw) {
        if (lane != 0) continue;
        double* __restrict__ fa = fac + (size_t)j * 5 * n;  // reciprocal of the U diagonal
        double* __restrict__ fb = fa + n;                    // U first super-diagonal
        double* __restrict__ fd = fb + n;                    // U second super-diagonal
        double* __restrict__ fl = fd + n;                    // multipliers
        double* __restrict__ fi = fl + n;                    // 1: rows i, i+1 were interchanged
        const double sh = lhat[j];
        // row i of the partly eliminated matrix is (a, b) in columns (i, i+1), carried in registers; one reciprocal per row
        // serves the multiplier and the back substitutions; a vanishing pivot (the shift IS an eigenvalue) becomes eps |T|
        double a = d[0] - sh, b = e[0];
        for (int i = 0; i + 1 < n; i++) {
            const double ci = e[i];                  // sub-diagonal entry of row i + 1
            const double an = d[i + 1] - sh, bn = e[i + 1];  // row i + 1: (ci, an, bn) in columns (i, i+1, i+2)
            if (fabs(a) >= fabs(ci)) {
                const double piv = fabs(a) < tiny ? (a < 0.0 ? -tiny : tiny) : a;
                const double r = 1.0 / piv;
                const double m = a != 0.0 ? ci * r : 0.0;
                fa[i] = r;
                fb[i] = b;
                fd[i] = 0.0;
                fl[i] = m;
                fi[i] = 0.0;
                a = an - m * b;
                b = bn;
            } else {
                const double r = 1.0 / ci;
                const double m = a * r;
                fa[i] = r;
                fb[i] = an;
                fd[i] = bn;
                fl[i] = m;
                fi[i] = 1.0;
                a = b - m * an;
                b = -m * bn;
            }
        }
        {
            const double piv = fabs(a) < tiny ? (a < 0.0 ? -tiny : tiny) : a;
            fa[n - 1] = 1.0 / piv;
            fb[n - 1] = 0.0;
            fd[n - 1] = 0.0;
        }
    }
    ctx.sync();
    constexpr int INV_ITERS = 3;
    for (int it = 0; it < INV_ITERS; it++) {
        for (int j = wid; j < k; j += nw) {
            if (lane != 0) continue;
            const double* __restrict__ fa = fac + (size_t)j * 5 * n;
            const double* __restrict__ fb = fa + n;
            const double* __restrict__ fd = fb + n;
            const double* __restrict__ fl = fd + n;
            const double* __restrict__ fi = fl + n;
            double* __restrict__ y = X + (size_t)j * n;
            if (it == 0) {  // start vector: fixed pseudo-random numbers in (-1, 1)
                unsigned sd = 12345u + 7919u * (unsigned)j;
                for (int i = 0; i < n; i++) {
                    sd = sd * 1664525u + 1013904223u;
                    y[i] = (double)(sd >> 8) * (2.0 / 16777216.0) - 1.0;
                }
            }
            // forward: y <- L^-1 P y, the running entry in a register (one multiply-add per row on the dependent chain)
            double cur = y[0];
            for (int i = 0; i + 1 < n; i++) {
                const double nxt = y[i + 1], m = fl[i];
                if (fi[i] == 0.0) {
                    y[i] = cur;
                    cur = nxt - m * cur;
                } else {
                    y[i] = nxt;
                    cur = cur - m * nxt;
                }
            }
            // backward: y <- U^-1 y with the reciprocal pivots
            double y1 = cur * fa[n - 1], y2 = 0.0;
            y[n - 1] = y1;
            for (int i = n - 2; i >= 0; i--) {
                const double t = (y[i] - fb[i] * y1 - fd[i] * y2) * fa[i];
                y[i] = t;
                y2 = y1;
                y1 = t;
            }
        }
        ctx.sync();
        if (tid < LD) {  // joint modified Gram-Schmidt, one group of threads
            for (int j = 0; j < k; j++) {
                double* y = X + (size_t)j * n;
                double nb = 0.0;
                for (int i = tid; i < n; i += LD) nb += y[i] * y[i];
                nb = ctx.lead_sum(nb);
                for (int jp = 0; jp < j; jp++) {
                    const double* z = X + (size_t)jp * n;
                    double s = 0.0;
                    for (int i = tid; i < n; i += LD) s += z[i] * y[i];
                    s = ctx.lead_sum(s);
                    for (int i = tid; i < n; i += LD) y[i] -= s * z[i];
                    ctx.lead_sync();
                }
                double nn = 0.0;
                for (int i = tid; i < n; i += LD) nn += y[i] * y[i];
                nn = ctx.lead_sum(nn);
                if (it == INV_ITERS - 1 && !(nn > 1e-12 * nb) && tid == 0) S[5] = 1.0;  // no direction of its own left
                const double sc = nn > 0.0 ? 1.0 / sqrt(nn) : 0.0;
                for (int i = tid; i < n; i += LD) y[i] *= sc;
                ctx.lead_sync();
            }
        }
        ctx.sync();
    }
    SE_STAMP(12);  // inverse iteration
    double* rh = cs + 3 * n;  // free until se_finish recomputes it
    for (int i = tid; i < n; i += nt) rh[i] = hv[i] != 0.0 ? 1.0 / hv[i] : 0.0;
    ctx.sync();
    // ---- Rayleigh quotients, residual check, back-transformation x = H_(n-1) .. H_1 u
    for (int j = wid; j < k; j += nw) {
        double* y = X + (size_t)j * n;
        double rq = 0.0;
        for (int i = lane; i < n; i += ws) {
            double t = d[i] * y[i];
            if (i > 0) t += e[i - 1] * y[i - 1];
            if (i + 1 < n) t += e[i] * y[i + 1];
            rq += y[i] * t;
        }
        rq = ctx.wsum(rq);
        double rs = 0.0;
        for (int i = lane; i < n; i += ws) {
            double t = (d[i] - rq) * y[i];
            if (i > 0) t += e[i - 1] * y[i - 1];
            if (i + 1 < n) t += e[i] * y[i + 1];
            rs += t * t;
        }
        rs = ctx.wsum(rs);
        if (lane == 0) {
            lam[j] = rq;
            const double lim = 1e3 * ulp * tnorm;
            if (!(rs <= lim * lim)) S[5] = 1.0;
        }
        ctx.warp_sync();
        if (ws == 1) {
            for (int i = 1; i < n; i++) {
                const double r = rh[i];
                if (r == 0.0) continue;
                double s = 0.0;
                for (int t = 0; t < i; t++) s += V[t * ld + i] * y[t];
                s *= r;
                for (int t = 0; t < i; t++) y[t] -= s * V[t * ld + i];
            }
        } else {  // the lane's components stay in registers: the chain per reflector is multiply-adds + one warp sum
            double yy[SE_PER_LANE];
#pragma unroll
            for (int u = 0; u < SE_PER_LANE; u++) yy[u] = lane + u * ws < n ? y[lane + u * ws] : 0.0;
            for (int i = 1; i < n; i++) {
                const double r = rh[i];
                if (r == 0.0) continue;
                double vv[SE_PER_LANE];
                double s = 0.0;
#pragma unroll
                for (int u = 0; u < SE_PER_LANE; u++) {
                    const int t = lane + u * ws;
                    vv[u] = t < i ? V[t * ld + i] : 0.0;
                    s += vv[u] * yy[u];
                }
                s = ctx.wsum(s) * r;
#pragma unroll
                for (int u = 0; u < SE_PER_LANE; u++) yy[u] -= s * vv[u];
            }
#pragma unroll
            for (int u = 0; u < SE_PER_LANE; u++)
                if (lane + u * ws < n) y[lane + u * ws] = yy[u];
        }
    }
    ctx.sync();
    SE_STAMP(13);  // Rayleigh quotients + back-transformation
    return S[5] != 0.0 ? -1 : k;
}

template <class Ctx, int SE_PER_LANE = 3>
SE_HD void sym_eig(Ctx ctx, double* V, int n, int ld, double* d, double* e, double* cs, double* scal) {
    if (n == 1) {
        if (ctx.tid() == 0) {
            d[0] = V[0];
            V[0] = 1.0;
        }
        ctx.sync();
        return;
    }
    se_tridiag<Ctx, SE_PER_LANE>(ctx, V, n, ld, d, e, cs, scal);
    se_finish<Ctx, SE_PER_LANE>(ctx, V, n, ld, d, e, cs, scal);
}
#undef VV

}  // namespace vb
