// FP64 pipe microbenchmark for B200: DFMA vs DMMA (mma.sync f64) vs cublasDgemm.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_peak tools/fp64_peak.cu -lcublas
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do { cudaError_t err_ = (x); if (err_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(err_), __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void dfma_kernel(double *out, double a, double b, int iters)
{
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dmma884_kernel(double *out, int iters)
{
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dmma16816_kernel(double *out, int iters)
{
    double c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + (threadIdx.x + i) * 1e-9;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = 1.0 - (threadIdx.x + i) * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                           "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F f, int reps = 5)
{
    cudaEvent_t s, e;
    cudaEventCreate(&s); cudaEventCreate(&e);
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(s); f(); cudaEventRecord(e);
        CK(cudaEventSynchronize(e));
        float ms; cudaEventElapsedTime(&ms, s, e);
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, p.multiProcessorCount, p.clockRate);
    double *out; CK(cudaMalloc(&out, sizeof(double) * 148 * 64 * 1024));
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        int threads = warps * 32, blocks = p.multiProcessorCount * 2;
        float ms = time_ms([&] { dfma_kernel<8><<<blocks, threads>>>(out, 1.000001, 1e-9, iters); });
        double tf = 2.0 * 8 * iters * (double)threads * blocks / (ms * 1e-3) / 1e12;
        printf(", \"dfma_tflops_w%d\": %.2f", warps * 2, tf);
    }
    for (int warps : {4, 8, 16}) {
        int threads = warps * 32, blocks = p.multiProcessorCount * 2;
        float ms = time_ms([&] { dmma884_kernel<8><<<blocks, threads>>>(out, iters); });
        double tf = 2.0 * 8 * 8 * 4 * 8 * iters * (double)warps * blocks / (ms * 1e-3) / 1e12;
        printf(", \"dmma884_tflops_w%d\": %.2f", warps * 2, tf);
        ms = time_ms([&] { dmma16816_kernel<4><<<blocks, threads>>>(out, iters); });
        tf = 2.0 * 16 * 8 * 16 * 4 * iters * (double)warps * blocks / (ms * 1e-3) / 1e12;
        printf(", \"dmma16816_tflops_w%d\": %.2f", warps * 2, tf);
    }
    cublasHandle_t h; cublasCreate(&h);
    for (int cfg = 0; cfg < 2; ++cfg) {
        int M = cfg == 0 ? 8192 : 512, N = cfg == 0 ? 8192 : 32768, K = cfg == 0 ? 8192 : 512;
        double *A, *B, *C;
        CK(cudaMalloc(&A, sizeof(double) * M * K)); CK(cudaMalloc(&B, sizeof(double) * K * N)); CK(cudaMalloc(&C, sizeof(double) * M * N));
        CK(cudaMemset(A, 0, sizeof(double) * M * K)); CK(cudaMemset(B, 0, sizeof(double) * K * N));
        double one = 1.0, zero = 0.0;
        float ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, M, N, K, &one, A, M, B, K, &zero, C, M); });
        printf(", \"cublas_dgemm_%dx%dx%d_tflops\": %.2f", M, N, K, 2.0 * M * N * K / (ms * 1e-3) / 1e12);
        cudaFree(A); cudaFree(B); cudaFree(C);
    }
    printf("}\n");
    return 0;
}
