// scan.cuh -- exclusive prefix sum fused into the kernel that produces the items.
//
// A BFS level needs two prefix sums (polygon sizes -> CSR offsets, winners per parent -> new state ids).  As
// stand-alone three-kernel scans they cost more launch latency than work (levels hold 10^3..10^5 items and a
// sharded march repeats them 273 times), so the producer kernel scans its own tile instead:
//   * every block owns FS_TILE consecutive items (FS_ITEMS per thread), computes their block-local exclusive
//     prefix and publishes the tile total;
//   * the block that finishes last (atomic ticket) scans the tile totals -> block_off[], writes the grand total
//     and re-arms the ticket;
//   * a consumer in a LATER kernel adds  block_off[i / FS_TILE] + local[i].
#pragma once
#include "common.cuh"

namespace amb {

constexpr int FS_THREADS = 256;
constexpr int FS_ITEMS = 8;
constexpr int FS_TILE = FS_THREADS * FS_ITEMS;

template <typename T>
struct FusedScanT {
    T *block_sums;                   // [gridDim.x]
    T *block_off;                    // [gridDim.x]
    unsigned int *ticket;            // zero before the launch; left zero
    unsigned long long *total;       // grand total (may be null)
    // optional: after the totals are known, *bump_dst += *bump_src (the running corner count of the march)
    unsigned long long *bump_dst;
    const unsigned long long *bump_src;
};
using FusedScan = FusedScanT<uint32_t>;
using FusedScan64 = FusedScanT<unsigned long long>;     // two 32-bit counters scanned at once (low | high << 32)

// v: this thread's FS_ITEMS consecutive items (item index = blockIdx.x * FS_TILE + threadIdx.x * FS_ITEMS + i)
// excl: their exclusive prefix within the block.  Must be called by all FS_THREADS threads of every block.
// Returns true in every thread of the block that finished last (after it has written block_off[] and the total);
// *grand_total then holds the total.
template <typename T>
__device__ __forceinline__ bool fused_scan_block(const T (&v)[FS_ITEMS], T (&excl)[FS_ITEMS], const FusedScanT<T> &fs,
                                                 T *grand_total = nullptr)
{
    __shared__ T sh_w[FS_THREADS / 32];
    __shared__ T sh_run;
    __shared__ bool sh_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T sum = 0;
#pragma unroll
    for (int i = 0; i < FS_ITEMS; ++i) sum += v[i];
    T inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) sh_w[warp] = inc;
    __syncthreads();
    T woff = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < FS_THREADS / 32; ++w) {
        if (w < warp) woff += sh_w[w];
        tile_total += sh_w[w];
    }
    T run = woff + inc - sum;
#pragma unroll
    for (int i = 0; i < FS_ITEMS; ++i) {
        excl[i] = run;
        run += v[i];
    }
    if (threadIdx.x == 0) {
        fs.block_sums[blockIdx.x] = tile_total;
        __threadfence();
        sh_last = (atomicAdd(fs.ticket, 1u) == gridDim.x - 1);
        sh_run = 0;
    }
    __syncthreads();
    if (!sh_last) return false;
    __threadfence();
    const int nb = (int)gridDim.x;
    for (int base = 0; base < nb; base += FS_THREADS) {
        const int b = base + (int)threadIdx.x;
        const T x = (b < nb) ? __ldcg(fs.block_sums + b) : T(0);
        T in2 = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const T t = __shfl_up_sync(0xFFFFFFFFu, in2, o);
            if (lane >= o) in2 += t;
        }
        __syncthreads();                 // sh_w is reused
        if (lane == 31) sh_w[warp] = in2;
        __syncthreads();
        T wo = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < FS_THREADS / 32; ++w) {
            if (w < warp) wo += sh_w[w];
            tot += sh_w[w];
        }
        if (b < nb) fs.block_off[b] = sh_run + wo + in2 - x;
        __syncthreads();
        if (threadIdx.x == 0) sh_run += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (fs.total) *fs.total = (unsigned long long)(sh_run & T(0xFFFFFFFFu));    // low counter
        if (fs.bump_dst) *fs.bump_dst += *fs.bump_src;
        *fs.ticket = 0;
    }
    if (grand_total) *grand_total = sh_run;
    return true;
}

// stand-alone form: local[i] = block-local exclusive prefix of in[i]
__global__ void __launch_bounds__(FS_THREADS) scan_local_kernel(const uint32_t *in, int n, uint32_t *local, FusedScan fs)
{
    pdl_enter();
    const int base = blockIdx.x * FS_TILE + threadIdx.x * FS_ITEMS;
    uint32_t v[FS_ITEMS], e[FS_ITEMS];
#pragma unroll
    for (int i = 0; i < FS_ITEMS; ++i) v[i] = (base + i < n) ? in[base + i] : 0u;
    fused_scan_block(v, e, fs);
#pragma unroll
    for (int i = 0; i < FS_ITEMS; ++i)
        if (base + i < n) local[base + i] = e[i];
}

}  // namespace amb
