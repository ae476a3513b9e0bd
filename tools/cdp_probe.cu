// Probe: does device-side cudaMemcpyAsync (CDP2, needs -rdc=true -lcudadevrt) work on this box?
// nvcc -O2 -rdc=true -gencode arch=compute_100,code=sm_100 -o tools/cdp_probe tools/cdp_probe.cu -lcudadevrt
#include <cstdio>
#include <cuda_runtime.h>
#include <thrust/device_vector.h>
#include <thrust/fill.h>
__global__ void plain(int *a) { a[threadIdx.x] = threadIdx.x; }
__global__ void cdp(int *dst, const int *src) {
    if (threadIdx.x == 0) cudaMemcpyAsync(dst, src, 32 * sizeof(int), cudaMemcpyDeviceToDevice);
}
int main() {
    int n = 0; cudaError_t e = cudaGetDeviceCount(&n);
    printf("device count %d (%s)\n", n, cudaGetErrorString(e));
    int *a, *b; cudaMalloc(&a, 128); cudaMalloc(&b, 128);
    plain<<<1, 32>>>(a);
    e = cudaDeviceSynchronize(); printf("plain kernel in rdc module: %s / last %s\n", cudaGetErrorString(e), cudaGetErrorString(cudaGetLastError()));
    cdp<<<1, 32>>>(b, a);
    e = cudaDeviceSynchronize(); printf("cdp kernel: %s / last %s\n", cudaGetErrorString(e), cudaGetErrorString(cudaGetLastError()));
    int h[32]; cudaMemcpy(h, b, 128, cudaMemcpyDeviceToHost); printf("b[5]=%d\n", h[5]);
    try { thrust::device_vector<int> v(1000); thrust::fill(v.begin(), v.end(), 7); int x = v[3]; printf("thrust ok %d\n", x); }
    catch (std::exception &ex) { printf("thrust failed: %s\n", ex.what()); }
    return 0;
}
