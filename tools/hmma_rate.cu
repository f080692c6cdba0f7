// Issue-rate probe for the legacy warp-level tensor path on sm_100a: mma.sync.aligned.m16n8k16 (fp16 operands, fp32
// accumulate; SASS HMMA.16816.F32), alone and mixed with FFMA2 / MUFU work, for 1..8 warps per SM sub-partition.
// Prints cycles per HMMA per SM and the equivalent dense MAC/clk/SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o build/hmma_rate tools/hmma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096;

__device__ __forceinline__ void hmma(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// MODE 0: 8 independent accumulators, HMMA only.  MODE 1: plus 8 FFMA per HMMA on other registers.  MODE 2: chained
// (each HMMA's A operand depends on the previous result through a pack), measures dependent latency.
template <int MODE>
__global__ void __launch_bounds__(1024) probe(float* out, long long* cyc, unsigned seed) {
    unsigned a[4] = {seed + threadIdx.x, seed * 3u, seed * 5u, seed * 7u}, b[2] = {seed * 11u, seed * 13u};
    float d[8][4];
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f; s[i] = threadIdx.x * 1e-3f + i; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 2) {
                unsigned aa[4] = {a[0] ^ __float_as_uint(d[(i + 7) & 7][0]), a[1], a[2], a[3]};
                hmma(d[i], aa, b);
            } else {
                hmma(d[i], a, b);
            }
            if (MODE == 1) {
#pragma unroll
                for (int k = 0; k < 8; ++k) s[k] = fmaf(s[k], 0.999f, 1e-3f);
            }
        }
    }
    long long t1 = clock64();
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += d[i][0] + d[i][1] + d[i][2] + d[i][3] + s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * sms * threads);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    for (int rep = 0; rep < 2; ++rep) probe<MODE><<<sms, threads>>>(out, cyc, 12345u);
    cudaDeviceSynchronize();
    long long h[512];
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
    double warps = threads / 32.0;
    double hmma_per_sm = kIters * 8.0 * warps;
    printf("%-44s warps/SM %4.0f  cycles %9.0f  cycles/HMMA/SM %.3f  MAC/clk/SM %.0f\n", name, warps, avg, avg / hmma_per_sm,
           hmma_per_sm * 2048.0 / avg);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int threads : {128, 256, 512, 1024}) run<0>("HMMA.16816.F32, 8 independent accumulators", threads);
    for (int threads : {128, 256, 512, 1024}) run<1>("HMMA + 8 FFMA per HMMA", threads);
    for (int threads : {128, 256, 512}) run<2>("HMMA chained through A", threads);
    return 0;
}
