// Experiment: host->device + device->host at the e2e leg's size, from ordinary pinned and from write-combined pinned memory.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main() {
    const size_t B = 805306368;
    void *hn, *hw, *ho, *di, *dout;
    CK(cudaHostAlloc(&hn, B, cudaHostAllocPortable));
    CK(cudaHostAlloc(&hw, B, cudaHostAllocPortable | cudaHostAllocWriteCombined));
    CK(cudaHostAlloc(&ho, B, cudaHostAllocPortable));
    CK(cudaMalloc(&di, B)); CK(cudaMalloc(&dout, B));
    cudaStream_t up, down; CK(cudaStreamCreate(&up)); CK(cudaStreamCreate(&down));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int mode = 0; mode < 6; mode++) {
        void* src = (mode & 1) ? hw : hn;
        const bool bidir = mode >= 2 && mode < 4, chunked = mode >= 4;
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0, 0));
            CK(cudaStreamWaitEvent(up, e0, 0)); CK(cudaStreamWaitEvent(down, e0, 0));
            const int iters = 5;
            for (int k = 0; k < iters; k++) {
                if (chunked) {      // 3 fields as separate copies, both directions
                    for (int f = 0; f < 3; f++) {
                        CK(cudaMemcpyAsync((char*)di + f * (B / 3), (char*)src + f * (B / 3), B / 3, cudaMemcpyHostToDevice, up));
                        CK(cudaMemcpyAsync((char*)ho + f * (B / 3), (char*)dout + f * (B / 3), B / 3, cudaMemcpyDeviceToHost, down));
                    }
                } else {
                    CK(cudaMemcpyAsync(di, src, B, cudaMemcpyHostToDevice, up));
                    if (bidir) CK(cudaMemcpyAsync(ho, dout, B, cudaMemcpyDeviceToHost, down));
                }
            }
            CK(cudaEventRecord(e1, up)); CK(cudaStreamWaitEvent(down, e1, 0));
            cudaEvent_t e2; CK(cudaEventCreate(&e2)); CK(cudaEventRecord(e2, down));
            CK(cudaEventSynchronize(e2));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e2));
            if (rep) printf("%s source, %s: %.2f ms per 805 MB step (%.1f GB/s per direction)\n", (mode & 1) ? "write-combined" : "ordinary pinned",
                            chunked ? "H2D + D2H in 3 chunks" : bidir ? "H2D + D2H" : "H2D only", ms / iters, B / (ms / iters) / 1e6);
        }
    }
    return 0;
}
