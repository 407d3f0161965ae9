// Per-kernel device timing for bench.py's roofline object: when enabled, every launch group is bracketed by CUDA
// events on the launching stream and waited for (so it serialises the pipeline: profile passes are separate from
// timed passes).
#pragma once
#include <cuda_runtime.h>

namespace vb {

struct KernelProfile {
    static constexpr int MAX_K = 16;
    bool on = false;
    cudaEvent_t a = nullptr, b = nullptr;
    double ms[MAX_K] = {};
    int count[MAX_K] = {};
    void enable(bool v) {
        on = v;
        if (v && !a) {
            cudaEventCreate(&a);
            cudaEventCreate(&b);
        }
        for (int i = 0; i < MAX_K; i++) ms[i] = 0, count[i] = 0;
    }
    void begin(cudaStream_t s) {
        if (on) cudaEventRecord(a, s);
    }
    void end(int k, cudaStream_t s, int launches = 1) {
        if (!on) return;
        cudaEventRecord(b, s);
        cudaEventSynchronize(b);
        float t = 0;
        cudaEventElapsedTime(&t, a, b);
        ms[k] += t;
        count[k] += launches;
    }
    ~KernelProfile() {
        if (a) cudaEventDestroy(a);
        if (b) cudaEventDestroy(b);
    }
};

}  // namespace vb
