// Hardware probe (run on the B200): tcgen05.mma kind::tf32 rate of ONE SM for a 128 x N x 8 instruction stream whose A
// operand descriptor starts on the 1024-byte swizzle atom (shift 0) or is shifted by s pixel slots of 128 bytes (what
// the halo-reuse convolution does for its filter taps), N = 64 / 128 / 256.  Operands are whatever is in shared memory
// (values do not matter for the rate); `reps` instructions are issued by one thread, then committed to an mbarrier.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 --cudart static -o mma_rate_probe mma_rate_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../rten_b200/csrc/ptx.cuh"

using namespace rtb;

struct Out {
    long long clk[8][3];  // [shift][n]
};

__global__ void __launch_bounds__(128, 1) rate_kernel(Out* out, int reps) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sa = base;                 // 640 slots x 128 B
    uint8_t* sb = base + 640 * 128;     // 256 rows x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(sb + 256 * 128);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (640 + 256) * 32; i += 128) reinterpret_cast<float*>(base)[i] = 1.0f;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_ptr, 512);
        tmem_relinquish();
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    const int shifts[8] = {0, 8, 1, 3, 58, 59, 117, 118};
    const int ns[3] = {64, 128, 256};
    uint32_t phase = 0;
    if (threadIdx.x == 0) {
        for (int ni = 0; ni < 3; ni++) {
            const uint32_t idesc = make_idesc(1, 2, 2, 128, ns[ni]);
            for (int s = 0; s < 8; s++) {
                const uint64_t ad = make_kmajor_sw128_desc(smem_u32(sa) + shifts[s] * 128);
                const uint64_t bd = make_kmajor_sw128_desc(smem_u32(sb));
                for (int pass = 0; pass < 2; pass++) {  // pass 0 warms up
                    const long long t0 = clock64();
                    for (int i = 0; i < reps; i += 4) {
#pragma unroll
                        for (int k = 0; k < 4; k++) umma_tf32(tmem, ad + 2 * k, bd + 2 * k, idesc, 1u);
                    }
                    umma_commit(bar);
                    mbar_wait(bar, phase);
                    phase ^= 1;
                    tc_fence_after();
                    if (pass) out->clk[s][ni] = clock64() - t0;
                }
            }
        }
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

int main() {
    Out* d;
    cudaMalloc(&d, sizeof(Out));
    const int smem = (640 + 256) * 128 + 1024 + 64;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int reps = 4096;
    rate_kernel<<<1, 128, smem>>>(d, reps);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("kernel failed: %s\n", cudaGetErrorString(e));
        return 1;
    }
    Out h;
    cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const int shifts[8] = {0, 8, 1, 3, 58, 59, 117, 118};
    const int ns[3] = {64, 128, 256};
    for (int ni = 0; ni < 3; ni++)
        for (int s = 0; s < 8; s++)
            printf("N=%3d A start shifted by %3d slots: %7.1f clk per 128xNx8 tf32 MMA (ideal %d)  -> %5.0f flop/clk/SM\n", ns[ni], shifts[s],
                   (double)h.clk[s][ni] / reps, ns[ni] / 2, 2.0 * 128 * ns[ni] * 8 / ((double)h.clk[s][ni] / reps));
    return 0;
}
