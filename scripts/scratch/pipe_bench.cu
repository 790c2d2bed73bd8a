// Micro-benchmark: issue rate of scalar vs packed fp32 (FADD/FADD2, FMUL2, FFMA2), ALU-pipe ops
// (FMNMX, FSETP+FSEL) and mixes, per SM sub-partition.  nvcc -arch=sm_100a -o pipe_bench pipe_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float addf(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float mulf(float a, float b) { float r; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float maxf(float a, float b) { float r; asm volatile("max.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float self(float a, float b, float c) { float r; asm volatile("{.reg .pred p; setp.gt.f32 p, %1, %3; selp.f32 %0, %1, %2, p;}" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
constexpr int NCH = 8, ITERS = 4096;
template <int MODE> __global__ void k(float* out, long long* cyc, float seed) {
    float f[NCH]; u64 p[NCH];
    for (int j = 0; j < NCH; j++) { f[j] = seed + j + threadIdx.x; float2 t = make_float2(f[j], f[j] + 1); p[j] = *reinterpret_cast<u64*>(&t); }
    float2 ct = make_float2(seed, seed * 0.5f); u64 c2 = *reinterpret_cast<u64*>(&ct);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int j = 0; j < NCH; j++) {
            if (MODE == 0) f[j] = addf(f[j], seed);
            if (MODE == 1) p[j] = add2(p[j], c2);
            if (MODE == 2) p[j] = mul2(p[j], c2);
            if (MODE == 3) p[j] = fma2(p[j], c2, c2);
            if (MODE == 4) f[j] = maxf(f[j], seed);
            if (MODE == 5) f[j] = self(f[j], seed, 1.0f);
            if (MODE == 6) { p[j] = add2(p[j], c2); f[j] = maxf(f[j], seed); }                       // 1 packed + 1 alu
            if (MODE == 7) { f[j] = addf(f[j], seed); f[j] = maxf(f[j], seed); }                     // dependent scalar add+max
            if (MODE == 8) { p[j] = add2(p[j], c2); f[j] = addf(f[j], seed); }                       // packed + scalar fma pipe
            if (MODE == 9) { p[j] = add2(p[j], c2); f[j] = maxf(f[j], seed); f[(j + 1) % NCH] = maxf(f[(j + 1) % NCH], 3.0f); }  // 1 packed + 2 alu
            if (MODE == 10) f[j] = mulf(f[j], seed);
            if (MODE == 11) { f[j] = addf(f[j], seed); f[j] = mulf(f[j], seed); f[j] = maxf(f[j], seed); }
        }
    }
    long long t1 = clock64();
    float s = 0; for (int j = 0; j < NCH; j++) { float2 t = *reinterpret_cast<float2*>(&p[j]); s += f[j] + t.x + t.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int per_iter, int threads) {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    k<MODE><<<148, threads>>>(out, cyc, 1.0001f); cudaDeviceSynchronize();
    k<MODE><<<148, threads>>>(out, cyc, 1.0001f); cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; i++) c += h[i]; c /= 148;
    double warps_per_smsp = threads / 32 / 4.0;
    double inst = (double)ITERS * NCH * per_iter * warps_per_smsp;
    printf("%-34s threads %4d: %.3f warp-instr/clk/SMSP (%.2f clk per instr)\n", name, threads, inst / c, c / inst);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int th : {128, 256, 512, 1024}) {
        run<0>("FADD", 1, th); run<10>("FMUL", 1, th); run<1>("FADD2", 1, th); run<2>("FMUL2", 1, th); run<3>("FFMA2", 1, th);
        run<4>("FMNMX", 1, th); run<5>("FSETP+FSEL", 2, th); run<6>("FADD2+FMNMX", 2, th); run<7>("FADD->FMNMX dep", 2, th);
        run<8>("FADD2+FADD", 2, th); run<9>("FADD2+2xFMNMX", 3, th); run<11>("FADD->FMUL->FMNMX dep", 3, th);
    }
    return 0;
}
