// Accuracy of the special-function-unit forms used by the product arithmetic (pdp_common.cuh), measured exhaustively
// against fp64 on the device:  nvcc -arch=sm_100a -o probe_math probe_math.cu && ./probe_math
#include <cstdio>
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>

__device__ float fast_log(float x) { float r; asm("lg2.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r * 0.693147182464599609375f; }
__device__ float fast_exp_stat(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.44269502162933349609375f)); return r; }

struct Acc { double max_abs, max_rel, max_ulp; unsigned long long n; };

__device__ void atomicMaxD(double* a, double v) {
    unsigned long long* p = (unsigned long long*)a; unsigned long long old = *p, assumed;
    do { assumed = old; if (__longlong_as_double(assumed) >= v) break; old = atomicCAS(p, assumed, __double_as_longlong(v)); } while (assumed != old);
}

// mode 0: fast_log, 1: logf, 2: fast_exp_stat, 3: expf
__global__ void k(uint32_t lo, uint32_t hi, int mode, Acc* acc) {
    double ma = 0, mr = 0;
    for (uint64_t b = lo + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; b <= hi; b += (uint64_t)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((uint32_t)b);
        float r; double t;
        if (mode == 0) { r = fast_log(x); t = log((double)x); }
        else if (mode == 1) { r = logf(x); t = log((double)x); }
        else if (mode == 2) { r = fast_exp_stat(x); t = exp((double)x); }
        else { r = expf(x); t = exp((double)x); }
        const double e = fabs((double)r - t);
        if (e > ma) ma = e;
        if (t != 0.0 && e / fabs(t) > mr) mr = e / fabs(t);
    }
    atomicMaxD(&acc->max_abs, ma); atomicMaxD(&acc->max_rel, mr);
}

static void run(const char* name, float a, float b, int mode) {
    Acc* d; cudaMalloc(&d, sizeof(Acc)); cudaMemset(d, 0, sizeof(Acc));
    uint32_t lo, hi; memcpy(&lo, &a, 4); memcpy(&hi, &b, 4);
    k<<<1184, 256>>>(lo, hi, mode, d);
    Acc h; cudaMemcpy(&h, d, sizeof(Acc), cudaMemcpyDeviceToHost);
    printf("%-34s x in [%g, %g]: max abs err %.3e  max rel err %.3e (%.2f ulp of 2^-24)\n", name, a, b, h.max_abs, h.max_rel, h.max_rel / 5.96e-8);
    cudaFree(d);
}

int main() {
    run("lg2.approx * ln2", 0.5f, 1.0f, 0);
    run("logf", 0.5f, 1.0f, 1);
    run("lg2.approx * ln2", 0.999f, 1.0f, 0);
    run("lg2.approx * ln2", 1e-6f, 0.5f, 0);
    run("logf", 1e-6f, 0.5f, 1);
    run("lg2.approx * ln2 (subnormal)", 1e-40f, 1.1e-38f, 0);
    run("ex2.approx.ftz(x log2e)", 0.0f, 30.0f, 2);
    run("expf", 0.0f, 30.0f, 3);
    return 0;
}
