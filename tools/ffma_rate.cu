// Issue-rate probe for the FP32 pipe on sm_100a: scalar FFMA (register / immediate multiplier) against the packed
// FFMA2 / FADD2 forms (fma.rn.f32x2). Prints warp-instructions per clock per SM sub-partition and the equivalent
// scalar FMA lanes per clock per SM. Build: nvcc -gencode arch=compute_100a,code=sm_100a -o build/ffma_rate tools/ffma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }

constexpr int kIters = 2048;

template <int MODE>
__global__ void __launch_bounds__(1024) probe(float* out, long long* cyc, float m, float c) {
    float s[16];
    u64 p[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { s[i] = threadIdx.x * 0.001f + i; p[i] = pk(s[i], s[i] + 0.5f); }
    u64 mm = pk(m, m), cc = pk(c, c);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) s[i] = fmaf(s[i], m, c);                        // FFMA reg, reg, reg
            if (MODE == 1) s[i] = fmaf(s[i], 0.4432097971f, c);            // FFMA with an immediate multiplier
            if (MODE == 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(mm), "l"(cc));
            if (MODE == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cc));
            if (MODE == 4) s[i] = s[i] + c;                                // FADD
            if (MODE == 5) { s[i] = fmaf(s[i], m, c); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(mm), "l"(cc)); }
            if (MODE == 6) { if (i & 3) s[i] = fmaf(s[i], m, c); else s[i] = __sinf(s[i]); }   // 3 FFMA : 1 MUFU
        }
    }
    long long t1 = clock64();
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += s[i] + lo(p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads, int per_iter_instr, int lanes_per_instr) {
    int dev_sms = 0;
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * dev_sms * threads);
    cudaMalloc(&cyc, sizeof(long long) * dev_sms);
    for (int rep = 0; rep < 2; ++rep) probe<MODE><<<dev_sms, threads>>>(out, cyc, 0.999f, 1e-3f);
    cudaDeviceSynchronize();
    long long h[512];
    cudaMemcpy(h, cyc, sizeof(long long) * dev_sms, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < dev_sms; ++i) avg += h[i]; avg /= dev_sms;
    double warps_per_smsp = threads / 32 / 4.0;
    double ipc = kIters * 16.0 * per_iter_instr * warps_per_smsp / avg;
    printf("%-34s warps/SMSP %4.1f  cycles %9.0f  warp-instr/clk/SMSP %.3f  fp32 lanes/clk/SM %.1f\n", name, warps_per_smsp, avg,
           ipc, ipc * 4 * 32 * lanes_per_instr / per_iter_instr);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int threads : {256, 512, 768}) {
        run<0>("FFMA r,r,r", threads, 1, 1);
        run<1>("FFMA r,imm,r", threads, 1, 1);
        run<2>("FFMA2 (fma.rn.f32x2)", threads, 1, 2);
        run<3>("FADD2 (add.rn.f32x2)", threads, 1, 2);
        run<4>("FADD", threads, 1, 1);
        run<5>("FFMA + FFMA2 interleaved", threads, 2, 3);
        run<6>("3 FFMA : 1 MUFU.SIN", threads, 1, 1);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
