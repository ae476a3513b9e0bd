// int8 tensor-core peak of one B200 for the instruction the split GEMM issues: tcgen05.mma.cta_group::1.kind::i8,
// M = 128, K = 32 bytes, operands resident in shared memory (no loads in the loop), accumulators in TMEM.
// One CTA per SM, one thread issues; two shapes: N = 256 (the largest single instruction) and the split GEMM's own
// instruction mix per K step (N = 256 + 192, 256 + 128, 256 + 64, 256, 192, 128, 64 -> 28 x 64 columns).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/int8_peak tools/int8_peak.cu
// Prints JSON: {"n256_tops": .., "split_mix_tops": .., "sm_count": .., "mhz_note": ..}
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../analyticmesh_b200/csrc/split.cuh"

using namespace amb;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

// mode 0: 2 x N=256 per step; mode 1: the split GEMM's 10 instructions per K step (SD = 7)
__global__ void __launch_bounds__(128, 1) peak_kernel(int steps, int mode, unsigned long long *sink)
{
    extern __shared__ unsigned char smem_raw[];
    __shared__ unsigned long long bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem_raw)[i] = 0x01010101u * (i & 3);
    const uint32_t bar_a = smem_u32(&bar);
    if (threadIdx.x == 0) {
        mbar_init(bar_a, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t sa = base, sb = base + 7 * SP_BM * SP_BK;          // A digits [7][128][32], B digits [7][64][32]
        uint32_t phase = 0;
        for (int s = 0; s < steps; ++s) {
            if (mode == 0) {
                tc_mma_i8(tmem, tc_smem_desc(sa), tc_smem_desc(sb), tc_idesc_i8(256), s > 0);
                tc_mma_i8(tmem + 256, tc_smem_desc(sa + SP_BM * SP_BK), tc_smem_desc(sb), tc_idesc_i8(256), s > 0);
            } else {
#pragma unroll
                for (int i = 0; i < 7; ++i) {
                    const uint64_t adesc = tc_smem_desc(sa + i * (SP_BM * SP_BK));
                    const int n_total = (7 - i) * SP_BN;
#pragma unroll
                    for (int c0 = 0; c0 < n_total; c0 += 256) {
                        const int n = (n_total - c0 < 256) ? (n_total - c0) : 256;
                        tc_mma_i8(tmem + (uint32_t)(i * SP_BN + c0), adesc, tc_smem_desc(sb + c0 * SP_BK), tc_idesc_i8(n),
                                  (s > 0 || i > 0) ? 1u : 0u);
                    }
                }
            }
            if ((s & 63) == 63 || s == steps - 1) {      // bound the queue: wait for the batch before issuing more
                tc_commit(bar_a);
                mbar_wait(bar_a, phase);
                phase ^= 1u;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
    }
    if (threadIdx.x == 0 && sink) sink[blockIdx.x] = steps;
}

int main()
{
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int smem = 96 * 1024;
    CK(cudaFuncSetAttribute(peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    unsigned long long *sink;
    CK(cudaMalloc(&sink, sizeof(unsigned long long) * sms));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    double tops[2] = {0, 0};
    for (int mode = 0; mode < 2; ++mode) {
        const int steps = 200000;
        const double macs_per_step = mode == 0 ? 2.0 * 128 * 256 * 32 : 28.0 * 128 * 64 * 32;
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaEventRecord(a));
            peak_kernel<<<sms, 128, smem>>>(steps, mode, sink);
            CK(cudaEventRecord(b));
            CK(cudaEventSynchronize(b));
            CK(cudaGetLastError());
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, a, b));
            if (rep > 0 && ms < best) best = ms;
        }
        tops[mode] = 2.0 * macs_per_step * steps * sms / (best * 1e-3) / 1e12;
    }
    printf("{\"n256_tops\": %.1f, \"split_mix_tops\": %.1f, \"sm_count\": %d, "
           "\"what\": \"tcgen05.mma.cta_group::1.kind::i8 M=128 K=32B issued back to back on resident smem operands, 1 CTA/SM; "
           "split_mix = the 10 instructions per K step of split_gemm_kernel<7> (28 digit products of 128x64)\"}\n",
           tops[0], tops[1], sms);
    return 0;
}
