// standalone check of the 3-D TMA box load used by k_fused_step
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#include <stdlib.h>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int NT>
__global__ void k(const __grid_constant__ CUtensorMap tmap, int x0, int y0, float* out, int bp) {
    extern __shared__ float sm_dyn[];
    float* smb = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(sm_dyn) + 127) & ~(uintptr_t)127);
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smb + 9 * NT);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bp * NT * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(smb)), "l"(&tmap), "r"(smem_u32(bar)), "r"(x0), "r"(y0), "r"(0) : "memory");
    }
    __syncthreads();
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(0) : "memory");
    for (int p = 0; p < bp; p++) out[p * NT + threadIdx.x] = smb[p * NT + threadIdx.x];
}

template <int NT>
__global__ void k2(const __grid_constant__ CUtensorMap tmap, int x0, int y0, float* out) {
    extern __shared__ float sm_dyn[];
    float* smb = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(sm_dyn) + 127) & ~(uintptr_t)127);
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smb + 9 * NT);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(NT * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(smb)), "l"(&tmap), "r"(smem_u32(bar)), "r"(x0), "r"(y0) : "memory");
    }
    __syncthreads();
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(0) : "memory");
    out[threadIdx.x] = smb[threadIdx.x];
}
int main(int argc, char** argv) {
    int rank = argc > 1 ? atoi(argv[1]) : 3, bp = argc > 2 ? atoi(argv[2]) : 9, l2 = argc > 3 ? atoi(argv[3]) : 1;
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    Enc enc = (Enc)fn;
    const int NT = 128;
    for (int W : {256, 16, 4096}) {
        int rows = 272; size_t pe = (size_t)rows * W;
        std::vector<float> h(18 * pe);
        for (size_t i = 0; i < h.size(); i++) h[i] = (float)(i % 100003);
        float* d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
        float* out; cudaMalloc(&out, 9 * NT * 4);
        for (int set = 0; set < 2; set++) for (int x0 : {-6, 110}) {
            alignas(64) CUtensorMap m;
            cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)rows, 9}; cuuint64_t strides[2] = {(cuuint64_t)W * 4, pe * 4};
            cuuint32_t box[3] = {NT, 1, (cuuint32_t)bp}, es[3] = {1, 1, 1};
            CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d + set * 9 * pe, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (l2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            cudaFuncSetAttribute(k<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 9 * NT * 4 + 256);
            if (rank == 3) k<NT><<<1, NT, 9 * NT * 4 + 256>>>(m, x0, 5, out, bp); else k2<NT><<<1, NT, 9 * NT * 4 + 256>>>(m, x0, 5, out);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<float> o(9 * NT); cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int p = 0; p < (rank == 3 ? bp : 1); p++) for (int t = 0; t < NT; t++) {
                int x = x0 + t; float want = (x >= 0 && x < W) ? h[(size_t)set * 9 * pe + p * pe + (size_t)5 * W + x] : 0.0f;
                if (o[p * NT + t] != want) bad++;
            }
            printf("W=%d set=%d x0=%d: encode %d, kernel %s, mismatches %d\n", W, set, x0, (int)r, cudaGetErrorString(e), bad);
            if (e != cudaSuccess) return 1;
        }
        cudaFree(d); cudaFree(out);
    }
    return 0;
}
