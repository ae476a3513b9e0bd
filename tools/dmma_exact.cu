// Is mma.sync.m8n8k4.f64 bit-identical to an ascending-k chain of FMAs?  (decides whether the DMMA
// path can keep bit-exact parity with the CPU oracle)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cuda_runtime.h>
__global__ void k(const double *A, const double *B, const double *C, double *D)
{
    const int t = threadIdx.x;
    double a = A[(t / 4) * 4 + (t % 4)];          // A[row][k], row-major 8x4
    double b = B[(t % 4) * 8 + (t / 4)];          // B[k][col], 4x8
    double c0 = C[(t / 4) * 8 + 2 * (t % 4)], c1 = C[(t / 4) * 8 + 2 * (t % 4) + 1];
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    D[(t / 4) * 8 + 2 * (t % 4)] = c0;
    D[(t / 4) * 8 + 2 * (t % 4) + 1] = c1;
}
int main()
{
    double hA[32], hB[32], hC[64], hD[64];
    double *A, *B, *C, *D;
    cudaMalloc(&A, 256); cudaMalloc(&B, 256); cudaMalloc(&C, 512); cudaMalloc(&D, 512);
    long asc = 0, desc = 0, pair = 0, none = 0, total = 0;
    srand(1);
    for (int trial = 0; trial < 2000; ++trial) {
        for (int i = 0; i < 32; ++i) { hA[i] = (rand() / (double)RAND_MAX - 0.5) * pow(2.0, rand() % 40 - 20); hB[i] = (rand() / (double)RAND_MAX - 0.5) * pow(2.0, rand() % 40 - 20); }
        for (int i = 0; i < 64; ++i) hC[i] = (trial & 1) ? 0.0 : (rand() / (double)RAND_MAX - 0.5);
        cudaMemcpy(A, hA, 256, cudaMemcpyHostToDevice); cudaMemcpy(B, hB, 256, cudaMemcpyHostToDevice); cudaMemcpy(C, hC, 512, cudaMemcpyHostToDevice);
        k<<<1, 32>>>(A, B, C, D);
        cudaMemcpy(hD, D, 512, cudaMemcpyDeviceToHost);
        for (int r = 0; r < 8; ++r) for (int c = 0; c < 8; ++c) {
            double up = hC[r * 8 + c], dn = hC[r * 8 + c];
            for (int kk = 0; kk < 4; ++kk) up = fma(hA[r * 4 + kk], hB[kk * 8 + c], up);
            for (int kk = 3; kk >= 0; --kk) dn = fma(hA[r * 4 + kk], hB[kk * 8 + c], dn);
            double p = fma(hA[r*4+1], hB[8+c], hA[r*4+0]*hB[c]) + fma(hA[r*4+3], hB[24+c], hA[r*4+2]*hB[16+c]) + hC[r*8+c];
            const double d = hD[r * 8 + c];
            ++total;
            if (memcmp(&d, &up, 8) == 0) ++asc; else if (memcmp(&d, &dn, 8) == 0) ++desc; else if (d == p) ++pair; else ++none;
        }
    }
    printf("{\"total\": %ld, \"equal_ascending_fma_chain\": %ld, \"equal_descending\": %ld, \"pairwise\": %ld, \"other\": %ld}\n", total, asc, desc, pair, none);
    return 0;
}
