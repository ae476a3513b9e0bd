// Link-time guard used ONLY when building the reference's CUDA extension as a reported baseline
// (baseline/build_ref.sh).  The reference launches kernels with zero threads whenever a work list
// is empty (e.g. backend/src/kernel.cu:43-45 with no extra constraints, inc/kernel.h:590-592 with
// num == 0) and never checks the result; with CUDA 12.9's thrust the stale
// cudaErrorInvalidConfiguration makes the next thrust call throw.  `-Xlinker --wrap=cudaLaunchKernel`
// routes the reference's launches through this function, which turns empty launches into no-ops.
// The reference SOURCES are not modified.
#include <cuda_runtime.h>
extern "C" cudaError_t __real_cudaLaunchKernel(const void *func, dim3 grid, dim3 block, void **args, size_t smem,
                                               cudaStream_t stream);
extern "C" cudaError_t __wrap_cudaLaunchKernel(const void *func, dim3 grid, dim3 block, void **args, size_t smem,
                                               cudaStream_t stream)
{
    if (grid.x == 0 || grid.y == 0 || grid.z == 0 || block.x == 0 || block.y == 0 || block.z == 0) return cudaSuccess;
    return __real_cudaLaunchKernel(func, grid, block, args, smem, stream);
}
