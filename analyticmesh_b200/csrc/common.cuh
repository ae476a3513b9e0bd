// common.cuh -- shared definitions for the B200 Analytic Marching engine.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

namespace amb {

constexpr int VSLOTS = 32;                  // working polygon capacity: one vertex per lane
constexpr int VERT_MAX_REF = 20;            // reference inc/macro.h:42 (we keep larger polygons, but count them)
constexpr double EPS_FEAS = 1e-20;          // reference inc/macro.h:18 (float64)
constexpr uint64_t SLOT_EMPTY = 0xFFFFFFFFFFFFFFFFull;
constexpr uint32_t CAND_TAG = 0x80000000u;  // low word of a slot: TAG | candidate index (this level), else state id
constexpr int NO_SLOT = -1;

// device-resident counters (index into an array of unsigned long long)
enum Counter {
    CNT_CORNERS = 0,      // running total of stored corners
    CNT_FACES,
    CNT_CANDIDATES,
    CNT_UNBOUNDED,
    CNT_OVERFLOW,
    CNT_OVER_VERTMAX,
    CNT_INCONSISTENT,
    CNT_STITCH_MISS,
    CNT_NEW,              // winners of the current level (scan total)
    CNT_CHUNK_CORNERS,    // corners of the current chunk (scan total)
    CNT_VERTS,            // unique vertices (combine)
    CNT_XCHG_ERROR,       // sharded march: barrier time-outs (low 16 bits), bad records, inbox overflows (xchg.cuh)
    CNT_NUM
};

__host__ __device__ inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// contribution of 32-bit word `v` at word index `w` to the additive key hash.
// H(key) = sum_w word_mix(w, key[w])  (mod 2^64): flipping one bit updates H in O(1).
__host__ __device__ inline uint64_t word_mix(uint32_t w, uint32_t v)
{
    return splitmix64((uint64_t(w + 1) << 32) | v);
}

__device__ __forceinline__ uint32_t slot_fp(uint64_t h) { return uint32_t(h >> 32); }

// Programmatic dependent launch: the kernels of a BFS level form one long dependency chain of small launches, so
// the launch latency between two of them is paid ~25 times per level.  A kernel launched with the
// programmatic-stream-serialization attribute may be scheduled while its predecessor still runs; it must call
// pdl_enter() before it touches global memory (griddepcontrol.wait returns once the predecessor grid has
// completed and its writes are visible).  launch_dependents is issued first so that the kernel after this one
// can be scheduled early as well.
__device__ __forceinline__ void pdl_wait()
{
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 900
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
// The successor may be scheduled once every CTA of this grid has executed this (or exited).  Short kernels issue
// it at once; persistent kernels issue it when their work is done -- a successor that becomes resident early
// only holds registers and shared memory that the running kernels (this one, or ITS successor's predecessor)
// still need: measured, digit kernel at 2 instead of 3 CTAs/SM because the next GEMM's CTAs sat next to it.
__device__ __forceinline__ void pdl_trigger()
{
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 900
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_enter()
{
    pdl_trigger();
    pdl_wait();
}

// Sharded march: where a polygon goes when its owner pushes it into every rank's exchange block (xchg.cuh); the
// clip kernel does the push itself, so the layout is a plain struct here.
constexpr int PEER_MAX = 16;
struct PeerPush {
    unsigned char *base[PEER_MAX];   // exchange block of every rank as mapped in this process
    int world, rank;
    unsigned long long region_off;   // this rank's polygon region inside a block: edges [cap_corners] ...
    unsigned long long xyz_off;      // ... | vertices [cap_corners][3], relative to region_off
    unsigned long long cnt_base;     // [level states] int: polygon size, written by the state's owner
    unsigned long long where_base;   // [level states] int2: (source rank, corner offset)
    int cap_corners;
    int *cursor;                     // device counter of this rank: corners pushed so far in the level
};

// first activation bit of every hidden layer (kernel parameter)
constexpr int MAX_LAYERS = 64;
struct LayerOffs {
    int D;
    int off[MAX_LAYERS + 2];      // off[h], h = 1..D+1
};

}  // namespace amb
