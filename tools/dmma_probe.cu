// FP64 throughput probe for sm_100a: plain DFMA against the FP64 tensor-core path (mma.sync.m8n8k4.f64).
// Answers the north star's question "FP64 DMMA tensor cores ... only if ncu shows gain" for B200: if the DMMA rate
// is not above the DFMA rate, mapping the Q2/Q3 basis contraction onto mma.sync cannot raise the FP64 ceiling of the
// assembly kernels (it could only save issue slots).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/dmma_probe.cu -o gpurun_out/dmma_probe && gpurun_out/dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int CHAINS = 8;

__global__ void dfma_kernel(double *out, double a, double b) {
    double acc[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = threadIdx.x * 1e-3 + c;
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// D(8x8) += A(8x4) * B(4x8): per warp 8*8*4 = 256 FMA = 512 flops per instruction
__global__ void dmma_kernel(double *out, double a, double b) {
    double acc[CHAINS][2];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c][0] = acc[c][1] = threadIdx.x * 1e-3 + c;
    double fa = a + threadIdx.x * 1e-9, fb = b;
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[c][0]), "+d"(acc[c][1])
                         : "d"(fa), "d"(fb));
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += acc[c][0] + acc[c][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int threads = 256, blocks = sms * 8;
    double *out;
    cudaMalloc(&out, sizeof(double) * threads * blocks);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int which = 0; which < 2; ++which) {
        float best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            if (which == 0)
                dfma_kernel<<<blocks, threads>>>(out, 0.999999, 1e-9);
            else
                dmma_kernel<<<blocks, threads>>>(out, 0.999999, 1e-9);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        const double nthreads = (double)threads * blocks;
        const double flops = which == 0 ? nthreads * ITERS * CHAINS * 2.0
                                        : (nthreads / 32.0) * ITERS * CHAINS * 512.0;
        printf("DMMA_PROBE %s: %.3f ms, %.2f TFLOP/s FP64 (%d SMs, %d threads/block, %d blocks)\n",
               which == 0 ? "DFMA (fma pipe)        " : "DMMA mma.sync.m8n8k4   ", best, flops / (best * 1e-3) / 1e12, sms,
               threads, blocks);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        printf("CUDA error: %s\n", cudaGetErrorString(e));
        return 1;
    }
    return 0;
}
