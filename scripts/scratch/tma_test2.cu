#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <vector>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// mode 0: 1-D bulk copy (no descriptor); mode 1: 2-D tensor, descriptor in global memory; mode 2: descriptor as __grid_constant__ param
__global__ void k(const __grid_constant__ CUtensorMap tmap, const CUtensorMap* gmap, const float* src, int mode, float* out, int cx, int cy) {
    __shared__ __align__(128) float smb[128];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(512) : "memory");
        if (mode == 0)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(smb)), "l"(src), "r"(512), "r"(smem_u32(&bar)) : "memory");
        else if (mode == 1)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(smem_u32(smb)), "l"(gmap), "r"(smem_u32(&bar)), "r"(cx), "r"(cy) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(smem_u32(smb)), "l"(&tmap), "r"(smem_u32(&bar)), "r"(cx), "r"(cy) : "memory");
    }
    __syncthreads();
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    out[threadIdx.x] = smb[threadIdx.x];
}
int main(int argc, char** argv) {
    int mode = argc > 1 ? atoi(argv[1]) : 0, how = 0, cx = argc > 2 ? atoi(argv[2]) : 0, cy = argc > 3 ? atoi(argv[3]) : 0;
    typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (how == 0) cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    else cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &q);
    Enc enc = (Enc)fn;
    int W = 256, rows = 64;
    std::vector<float> h((size_t)W * rows);
    for (size_t i = 0; i < h.size(); i++) h[i] = (float)i;
    float *d, *out; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice); cudaMalloc(&out, 512);
    alignas(64) CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)rows}; cuuint64_t strides[1] = {(cuuint64_t)W * 4};
    cuuint32_t box[2] = {128, 1}, es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUtensorMap* gm; cudaMalloc(&gm, 128); cudaMemcpy(gm, &m, 128, cudaMemcpyHostToDevice);
    unsigned long long* w = (unsigned long long*)&m;
    printf("encode %d q %d desc words: %016llx %016llx %016llx %016llx\n", (int)r, (int)q, w[0], w[1], w[2], w[3]);
    k<<<1, 128>>>(m, gm, d, mode, out, cx, cy);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> o(128); cudaMemcpy(o.data(), out, 512, cudaMemcpyDeviceToHost);
    int bad = 0; for (int t = 0; t < 128; t++) { int x = cx + t; float want = (x >= 0 && x < W && cy >= 0 && cy < rows) ? h[(size_t)cy * W + x] : 0.0f; if (o[t] != want) bad++; }
    printf("mode %d cx %d cy %d: kernel %s, mismatches %d\n", mode, cx, cy, cudaGetErrorString(e), bad);
    return 0;
}
