// Hardware probe (run on the B200): does a K-major SWIZZLE_128B shared-memory matrix descriptor accept a start address
// that is NOT aligned to the 1024-byte swizzle atom (rows shifted by s * 128 B)?  The halo-reuse 3x3 convolution
// addresses the nine filter taps as shifted windows of ONE activation patch in shared memory, so it needs exactly this.
// Two interpretations are tested per shift: descriptor base_offset = 0 (swizzle phase taken from the absolute
// shared-memory address bits [7,10)) and base_offset = (start >> 7) & 7 (the PTX ISA's rule for unaligned starts).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 --cudart static -o desc_probe desc_probe.cu && ./desc_probe
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../rten_b200/csrc/ptx.cuh"

using namespace rtb;

constexpr int SLOTS = 320;     // activation patch: SLOTS rows of 32 floats (128 B)
constexpr int NB = 32;         // B rows (output columns)
constexpr int NSHIFT = 12;

struct Params {
    int shifts[NSHIFT];
    int variant;  // 0: base_offset = 0; 1: base_offset = (start >> 7) & 7
    float* out;   // [NSHIFT][128][NB]
};

__device__ __forceinline__ uint64_t desc_with(uint32_t addr, uint32_t base_offset) {
    uint64_t d = make_kmajor_sw128_desc(addr);
    d |= static_cast<uint64_t>(base_offset & 7) << 49;
    return d;
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap map_a,
                                                       const __grid_constant__ CUtensorMap map_b, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sa = base;                    // SLOTS * 128 B
    uint8_t* sb = base + SLOTS * 128;      // NB * 128 B (SLOTS * 128 is a multiple of 1024)
    uint64_t* bar = reinterpret_cast<uint64_t*>(sb + NB * 128);
    uint64_t* mma_bar = bar + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(mma_bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_ptr, 32);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, SLOTS * 128 + NB * 128);
        // the A patch is loaded in boxes of <= 256 rows
        tma_load_2d(sa, &map_a, bar, 0, 0);
        tma_load_2d(sa + 160 * 128, &map_a, bar, 0, 160);
        tma_load_2d(sb, &map_b, bar, 0, 0);
    }
    mbar_wait(bar, 0);
    const uint32_t idesc = make_idesc(1, 2, 2, 128, NB);
    uint32_t phase = 0;
    for (int s = 0; s < NSHIFT; s++) {
        if (threadIdx.x == 0) {
            tc_fence_after();
            const uint32_t a0 = smem_u32(sa) + p.shifts[s] * 128;
            const uint32_t bo = p.variant ? ((a0 >> 7) & 7) : 0;
            for (int k = 0; k < 4; k++)
                umma_tf32(tmem, desc_with(a0 + 32 * k, bo), make_kmajor_sw128_desc(smem_u32(sb) + 32 * k), idesc, k ? 1u : 0u);
            umma_commit(mma_bar);
        }
        mbar_wait(mma_bar, phase);
        phase ^= 1;
        tc_fence_after();
        uint32_t v[32];
        tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16), v);
        tmem_ld_wait();
        float* o = p.out + ((size_t)s * 128 + warp * 32 + lane) * NB;
        for (int j = 0; j < 32; j++) o[j] = __uint_as_float(v[j]);
        tc_fence_before();
        __syncthreads();
    }
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 32);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return 2;                                                                 \
        }                                                                             \
    } while (0)

int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
    std::vector<float> a((size_t)SLOTS * 32), b((size_t)NB * 32);
    uint64_t st = 88172645463325252ull;
    auto rnd = [&]() {
        st ^= st << 13;
        st ^= st >> 7;
        st ^= st << 17;
        return (float)((int)(st % 17) - 8);  // small integers: exact in TF32 and in the f32 accumulator
    };
    for (auto& v : a) v = rnd();
    for (auto& v : b) v = rnd();
    float *da, *db, *dout;
    CK(cudaMalloc(&da, a.size() * 4));
    CK(cudaMalloc(&db, b.size() * 4));
    CK(cudaMalloc(&dout, (size_t)NSHIFT * 128 * NB * 4));
    CK(cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
    CUtensorMap ma, mb;
    {
        cuuint64_t dims[2] = {32, SLOTS}, strides[1] = {128};
        cuuint32_t box[2] = {32, 160}, es[2] = {1, 1};
        if (enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, da, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            printf("encode A failed\n");
            return 2;
        }
        cuuint64_t dimsb[2] = {32, NB};
        cuuint32_t boxb[2] = {32, NB};
        if (enc(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, db, dimsb, strides, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            printf("encode B failed\n");
            return 2;
        }
    }
    const int smem = SLOTS * 128 + NB * 128 + 1024 + 256;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    Params p;
    const int shifts[NSHIFT] = {0, 8, 1, 2, 3, 7, 9, 16, 17, 31, 59, 118};
    for (int i = 0; i < NSHIFT; i++) p.shifts[i] = shifts[i];
    p.out = dout;
    std::vector<float> out((size_t)NSHIFT * 128 * NB);
    int verdict[2] = {1, 1};
    for (int variant = 0; variant < 2; variant++) {
        p.variant = variant;
        CK(cudaMemset(dout, 0xff, out.size() * 4));
        probe_kernel<<<1, 128, smem>>>(ma, mb, p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("variant %d: kernel failed: %s\n", variant, cudaGetErrorString(e));
            verdict[variant] = 0;
            cudaGetLastError();
            continue;
        }
        CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
        for (int s = 0; s < NSHIFT; s++) {
            int bad = 0;
            for (int r = 0; r < 128; r++)
                for (int n = 0; n < NB; n++) {
                    double acc = 0;
                    for (int k = 0; k < 32; k++) acc += (double)a[(size_t)(r + shifts[s]) * 32 + k] * b[(size_t)n * 32 + k];
                    if ((float)acc != out[((size_t)s * 128 + r) * NB + n]) bad++;
                }
            printf("variant %d (base_offset %s) shift %3d rows: %s (%d of %d wrong)\n", variant,
                   variant ? "= (start>>7)&7" : "= 0", shifts[s], bad ? "MISMATCH" : "exact", bad, 128 * NB);
            if (bad) verdict[variant] = 0;
        }
    }
    printf("RESULT base_offset0_all_shifts_ok=%d base_offset_rule_all_shifts_ok=%d\n", verdict[0], verdict[1]);
    return 0;
}
