// Exhaustive check of the product build's logarithms (pdp_common.cuh) against libdevice's logf, over all 2^32 arguments:
//   L40(x)     vs logf(max.NaN(x, 1e-40f))          L40_1m(x) vs logf(max.NaN(1 - x, 1e-40f))
// Counts the arguments whose results differ in any bit (two NaNs count as equal).
//   nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -o probe_log probe_log.cu && ./probe_log
#include <cstdio>
#include <cstdint>
#include "../../pdp_solver_b200/csrc/pdp_common.cuh"
void pdp_set_error(const char*, ...) {}

__global__ void k(unsigned long long* bad, uint32_t* first) {
    unsigned long long b0 = 0, b1 = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < (1ull << 32); i += (uint64_t)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((uint32_t)i);
        const float a0 = L40(x), r0 = logf(tmaxf(x, PDP_EPS40));
        const float a1 = L40_1m(x), r1 = logf(tmaxf(1.f - x, PDP_EPS40));
        if (!(a0 != a0 && r0 != r0) && __float_as_uint(a0) != __float_as_uint(r0)) { ++b0; atomicMin(&first[0], (uint32_t)i); }
        if (!(a1 != a1 && r1 != r1) && __float_as_uint(a1) != __float_as_uint(r1)) { ++b1; atomicMin(&first[1], (uint32_t)i); }
    }
    atomicAdd(&bad[0], b0); atomicAdd(&bad[1], b1);
}
int main() {
    unsigned long long* d; uint32_t* f;
    cudaMalloc(&d, 16); cudaMemset(d, 0, 16); cudaMalloc(&f, 8); cudaMemset(f, 0xff, 8);
    k<<<148 * 8, 256>>>(d, f);
    unsigned long long h[2]; uint32_t hf[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); cudaMemcpy(hf, f, 8, cudaMemcpyDeviceToHost);
    printf("L40: %llu of 2^32 arguments differ from logf (first 0x%08x);  L40_1m: %llu differ (first 0x%08x)  [%s]\n", h[0], hf[0], h[1], hf[1],
           cudaGetErrorString(cudaDeviceSynchronize()));
    return (h[0] || h[1]) ? 1 : 0;
}
