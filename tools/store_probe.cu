// Hardware probe (run on the B200): how fast can ONE SM push an output tile out to L2 / HBM?
//   A  bulk tensor stores (cp.async.bulk.tensor.4d ... bulk_group) of 128 x 32-float chunks (16 KB, 128B-swizzled staging
//      buffer -- what the GEMM epilogue does), D stores in flight per CTA, into an [M, N] f32 matrix
//   B  the same bytes with plain st.global.v4 from registers, thread = row (each thread writes its own 128-byte row piece)
//   C  st.global.v4 with a warp writing 4 full 128-byte row pieces per instruction (coalesced)
// for N = 64 / 256 / 1024 columns, all 148 SMs or a single SM.  Reports bytes per clock per SM and GB/s.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 --cudart static -o store_probe store_probe.cu && ./store_probe
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../rten_b200/csrc/ptx.cuh"

using namespace rtb;

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            fprintf(stderr, "%s failed: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

struct Params {
    float* out;
    int M, N;          // matrix
    int chunks;        // 128 x 32 chunks per CTA
    int depth;         // (A) stores in flight per CTA
    long long* clk;    // per-CTA elapsed clocks
};

// chunk c of CTA b -> (row tile, column chunk): consecutive chunks walk the columns of one row tile first (like the epilogue)
__device__ __forceinline__ void chunk_coord(const Params& p, int b, int c, int& m0, int& n0) {
    const int per_row = p.N / 32;
    const long long g = (long long)b * p.chunks + c;
    n0 = (int)(g % per_row) * 32;
    m0 = (int)((g / per_row) % (p.M / 128)) * 128;
}

__global__ void __launch_bounds__(128, 1) tma_store_kernel(const __grid_constant__ CUtensorMap map, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int r = threadIdx.x;
    for (int d = 0; d < p.depth; d++)
        for (int j = 0; j < 8; j++) *reinterpret_cast<float4*>(base + d * 16384 + r * 128 + j * 16) = make_float4(r, j, d, 1.0f);
    fence_proxy_async();
    __syncthreads();
    const long long t0 = clock64();
    if (threadIdx.x == 0) {
        for (int c = 0; c < p.chunks; c++) {
            int m0, n0;
            chunk_coord(p, blockIdx.x, c, m0, n0);
            // the buffer about to be re-used must have been read: allow depth-1 stores in flight
            switch (p.depth) {
                case 1: asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); break;
                case 2: asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); break;
                case 3: asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); break;
                default: asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory"); break;
            }
            tma_store_4d(&map, base + (c % p.depth) * 16384, n0, m0, 0, 0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // writes performed
        p.clk[blockIdx.x] = clock64() - t0;
    }
}

// thread = row: each thread writes its own 128-byte piece of a row (8 x st.global.v4)
__global__ void __launch_bounds__(128, 1) row_store_kernel(const Params p) {
    const int r = threadIdx.x;
    const long long t0 = clock64();
    for (int c = 0; c < p.chunks; c++) {
        int m0, n0;
        chunk_coord(p, blockIdx.x, c, m0, n0);
        float4* dst = reinterpret_cast<float4*>(p.out + (long long)(m0 + r) * p.N + n0);
#pragma unroll
        for (int j = 0; j < 8; j++) dst[j] = make_float4(r, j, c, 1.0f);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) p.clk[blockIdx.x] = clock64() - t0;
}

// coalesced: a warp instruction writes 4 row pieces of 128 bytes (lane = 8 * row + 16-byte column)
__global__ void __launch_bounds__(128, 1) coalesced_store_kernel(const Params p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t0 = clock64();
    for (int c = 0; c < p.chunks; c++) {
        int m0, n0;
        chunk_coord(p, blockIdx.x, c, m0, n0);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int row = warp * 32 + j * 4 + (lane >> 3);
            *reinterpret_cast<float4*>(p.out + (long long)(m0 + row) * p.N + n0 + (lane & 7) * 4) = make_float4(row, j, c, 1.0f);
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) p.clk[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    CK(cudaSetDevice(0));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
    int clock_khz = 0;
    CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
    long long* dclk;
    CK(cudaMalloc(&dclk, 148 * 8));
    CK(cudaFuncSetAttribute(tma_store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384 + 1024));
    const int chunks = 64;  // 1 MB per CTA
    for (int N : {64, 256, 1024}) {
        const int M = 148 * chunks * 128 * 32 / N;  // every CTA owns distinct rows / columns: 148 MB in all
        float* out;
        CK(cudaMalloc(&out, (size_t)M * N * 4));
        CUtensorMap map;
        cuuint64_t dims[4] = {(cuuint64_t)N, (cuuint64_t)M, 1, 1}, strides[3] = {(cuuint64_t)N * 4, (cuuint64_t)M * N * 4, (cuuint64_t)M * N * 4};
        cuuint32_t box[4] = {32, 128, 1, 1}, es[4] = {1, 1, 1, 1};
        if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            fprintf(stderr, "tensor map encode failed\n");
            return 1;
        }
        for (int grid : {1, 148}) {
            auto report = [&](const char* name, int depth, float ms) {
                std::vector<long long> clk(grid);
                CK(cudaMemcpy(clk.data(), dclk, grid * 8, cudaMemcpyDeviceToHost));
                double mean = 0;
                for (long long c : clk) mean += (double)c / grid;
                const double bytes = (double)chunks * 16384;
                printf("N=%5d grid=%3d %-34s depth=%d: %6.1f B/clk/SM  (%8.0f clk per 16 KB chunk)   whole launch %7.1f GB/s\n", N, grid, name, depth,
                       bytes / mean, mean / chunks, bytes * grid / (ms * 1e-3) / 1e9);
            };
            cudaEvent_t e0, e1;
            CK(cudaEventCreate(&e0));
            CK(cudaEventCreate(&e1));
            for (int depth : {1, 2, 3, 4}) {
                Params p{out, M, N, chunks, depth, dclk};
                float best = 1e30f;
                for (int rep = 0; rep < 3; rep++) {
                    CK(cudaEventRecord(e0));
                    tma_store_kernel<<<grid, 128, 4 * 16384 + 1024>>>(map, p);
                    CK(cudaEventRecord(e1));
                    CK(cudaEventSynchronize(e1));
                    float ms;
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    best = ms < best ? ms : best;
                }
                CK(cudaGetLastError());
                report("bulk tensor store (128 x 32 box)", depth, best);
            }
            for (int which = 0; which < 2; which++) {
                Params p{out, M, N, chunks, 0, dclk};
                float best = 1e30f;
                for (int rep = 0; rep < 3; rep++) {
                    CK(cudaEventRecord(e0));
                    if (which == 0)
                        row_store_kernel<<<grid, 128>>>(p);
                    else
                        coalesced_store_kernel<<<grid, 128>>>(p);
                    CK(cudaEventRecord(e1));
                    CK(cudaEventSynchronize(e1));
                    float ms;
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    best = ms < best ? ms : best;
                }
                CK(cudaGetLastError());
                report(which == 0 ? "st.global.v4, thread = row" : "st.global.v4, coalesced 128 B pieces", 0, best);
            }
        }
        CK(cudaFree(out));
    }
    printf("SM clock attribute: %d kHz\n", clock_khz);
    return 0;
}
