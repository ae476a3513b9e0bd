// xchg.cuh -- the per-level exchange of the sharded march over NVLink peer memory (no NCCL, no host in the loop).
//
// One march is spread over `world` GPUs, one process each (the reference has no multi-GPU path at
// all, its only scale-out is the serial voxel loop of backend/main.py:475-556).  Compose + clip of a state run
// on the rank that owns it; the key arena, the polygon CSR and the state numbering are replicated and stay
// bit-identical on every rank; the visited set is sharded by key hash.  Two things cross the links per level:
//
//   polygons    the owner of a state PUSHES its polygon (edge ids + vertices, 28 B per corner, compacted), its size
//               and its location {rank, offset} into every peer's block with plain stores over NVLink; after one
//               device-side barrier every rank prefix-sums the sizes (already in state order) and copies the
//               polygons from its own inbox into the CSR -- sizes and offsets never visit the host.
//               (Round 1: three ncclAllReduce over zero-padded buffers + a host sync per level.)
//   winners     neighbour candidates are de-duplicated by the rank that owns the candidate key's hash (a rank
//               probes / inserts only 1/world of the candidates into its shard of the visited set); the 32-bit
//               mask "which edge slots of state s discovered a new state" is pushed to every peer and OR-ed
//               after a second barrier.  4 B per state instead of the candidate keys: every rank can rebuild
//               the winners' keys itself because it holds the parents' keys.
//
// Every rank owns ONE exchange block (cudaMalloc, exported with cudaIpcGetMemHandle, opened by the peers):
//     [0, XCHG_CTRL_BYTES)         control: arrive[src] epoch flags (one 128-byte line per source rank)
//     world polygon regions        region r is written by rank r only: edges [cap_corners] | vertices [cap_corners][3]
//     world mask regions           [mask_cap] uint32, region r written by rank r only
//     sizes, locations             [mask_cap] int + [mask_cap] int2, entry s written by the owner of state s
// A region is reused every level: a rank writes level l+1 only after the barrier that follows level l's winner
// exchange, which every rank reaches only after it has consumed level l's polygons (stream order).
#pragma once
#include "frontier.cuh"

namespace amb {

constexpr int XCHG_MAX_WORLD = 16;
constexpr size_t XCHG_CTRL_BYTES = 4096;
constexpr size_t XCHG_HDR_BYTES = 64;

struct XchgPeers {
    unsigned char *base[XCHG_MAX_WORLD];   // exchange block of every rank as mapped in THIS process (base[rank] = own)
    int world, rank;
};

struct XchgLayout {
    size_t region_bytes;      // one polygon region
    size_t mask_base;         // offset of mask region 0
    size_t cnt_base;          // [mask_cap] int: polygon size per level-local state, written by the state's owner
    size_t where_base;        // [mask_cap] int2: (source rank, corner offset in that rank's region)
    int cap_corners;          // corners per polygon region
    int mask_cap;             // states per level the sharded march accepts
    __host__ __device__ size_t region(int r) const { return XCHG_CTRL_BYTES + (size_t)r * region_bytes; }
    __host__ __device__ size_t edge_off() const { return 0; }
    __host__ __device__ size_t xyz_off() const { return ((size_t)cap_corners * 4 + 15) & ~size_t(15); }
    __host__ __device__ size_t mask(int r) const { return mask_base + (size_t)r * mask_cap * 4; }
};

__device__ __forceinline__ unsigned long long xchg_now_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Device-side barrier over all ranks: thread q announces `epoch` in rank q's block (release: everything this
// rank wrote before -- earlier kernels of the stream included -- is visible to whoever observes the flag) and
// waits until rank q has announced it here.  A peer that never arrives (crashed process) ends the wait after
// timeout_ns and raises the error counter instead of hanging the GPU.
__global__ void xchg_barrier_kernel(XchgPeers p, uint32_t epoch, unsigned long long timeout_ns, unsigned long long *counters)
{
    pdl_enter();
    const int q = threadIdx.x;
    if (q < p.world) {
        __threadfence_system();
        uint32_t *remote = reinterpret_cast<uint32_t *>(p.base[q]) + p.rank * 32;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
        const uint32_t *mine = reinterpret_cast<const uint32_t *>(p.base[p.rank]) + q * 32;
        const unsigned long long t0 = xchg_now_ns();
        for (;;) {
            uint32_t v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
            if ((int32_t)(v - epoch) >= 0) break;
            if (xchg_now_ns() - t0 > timeout_ns) {
                atomicAdd(counters + CNT_XCHG_ERROR, 1ull);
                break;
            }
            __nanosleep(64);
        }
        __threadfence_system();
    }
}

// ---- polygons: owner -> every rank ------------------------------------------------------------------------
struct XchgPackArgs {
    const int *idx;           // the level-local indices of the states this rank owns (prefix of the permutation)
    int n;                    // ... and their number
    const int *cnt;           // clip output, stride VSLOTS per level-local state
    const int *edges;
    const double *verts;
    int *cursor;              // device counter, zero at launch: corners packed so far
    XchgPeers p;
    XchgLayout lay;
    unsigned long long *counters;
};

__global__ void xchg_pack_kernel(const XchgPackArgs a)
{
    pdl_enter();
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= a.n) return;
    const size_t reg = a.lay.region(a.p.rank);
    const int s = a.idx[i];
    int k = a.cnt[s];
    int off = 0;
    if (lane == 0 && k > 0) {
        off = atomicAdd(a.cursor, k);
        if (off + k > a.lay.cap_corners) {      // reported; the level is then incomplete and the march fails
            atomicAdd(a.counters + CNT_XCHG_ERROR, 1ull << 32);
            off = -1;
        }
    }
    off = __shfl_sync(0xFFFFFFFFu, off, 0);
    if (off < 0) k = 0, off = 0;
    int e = 0;
    double x = 0, y = 0, z = 0;
    if (lane < k) {
        e = a.edges[(size_t)s * VSLOTS + lane];
        const double *v = a.verts + ((size_t)s * VSLOTS + lane) * 3;
        x = v[0]; y = v[1]; z = v[2];
    }
    for (int q = 0; q < a.p.world; ++q) {
        unsigned char *b = a.p.base[q];
        if (lane == 0) {
            reinterpret_cast<int *>(b + a.lay.cnt_base)[s] = k;
            reinterpret_cast<int2 *>(b + a.lay.where_base)[s] = make_int2(a.p.rank, off);
        }
        if (lane < k) {
            reinterpret_cast<int *>(b + reg + a.lay.edge_off())[off + lane] = e;
            double *o = reinterpret_cast<double *>(b + reg + a.lay.xyz_off()) + (size_t)(off + lane) * 3;
            o[0] = x; o[1] = y; o[2] = z;
        }
    }
}

// own inbox -> global CSR, in state order (the sharded counterpart of compact_faces_kernel)
struct XchgCompactArgs {
    const unsigned char *own;
    XchgLayout lay;
    const int *cnt_all;
    const uint32_t *off;          // block-local exclusive scan of cnt_all (scan.cuh) ...
    const uint32_t *block_off;    // ... + offset of the FS_TILE block
    const int2 *where;
    int S, sid0;
    long long *face_off;
    int *face_edges;
    double *face_xyz;
    unsigned long long *counters;
};

__global__ void xchg_compact_kernel(const XchgCompactArgs a)
{
    pdl_enter();
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (s >= a.S) return;
    const long long base = (long long)a.counters[CNT_CORNERS] + a.off[s] + a.block_off[s / FS_TILE];
    const int k = a.cnt_all[s];
    if (lane == 0) {
        a.face_off[a.sid0 + s] = base;
        if (s == a.S - 1) a.face_off[a.sid0 + s + 1] = base + k;
        if (k > 0) atomicAdd(a.counters + CNT_FACES, 1ull);
    }
    if (lane < k) {
        const int2 w = a.where[s];
        const unsigned char *reg = a.own + a.lay.region(w.x);
        a.face_edges[base + lane] = reinterpret_cast<const int *>(reg + a.lay.edge_off())[w.y + lane];
        const double *v = reinterpret_cast<const double *>(reg + a.lay.xyz_off()) + (size_t)(w.y + lane) * 3;
        double *o = a.face_xyz + (size_t)(base + lane) * 3;
        o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
    }
}

// ---- winners: hash owner -> every rank ----------------------------------------------------------------------
// thread per parent state: bit j of the mask = candidate (s, j) was inserted by THIS rank and won its slot
__global__ void xchg_push_masks_kernel(const LevelArgs a, XchgPeers p, XchgLayout lay)
{
    pdl_enter();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.S) return;
    uint32_t m = 0;
    const int k = int(a.face_off[a.lb + s + 1] - a.face_off[a.lb + s]);
    for (int j = 0; j < k; ++j) {
        const int slot = a.cand_slot[(size_t)s * VSLOTS + j];
        if (slot != NO_SLOT && uint32_t(a.table.slots[slot]) == (CAND_TAG | cand_index(s, j))) m |= 1u << j;
    }
    const size_t off = lay.mask(p.rank);
    for (int q = 0; q < p.world; ++q) reinterpret_cast<uint32_t *>(p.base[q] + off)[s] = m;
}

// ---- winners of a level: one launch ------------------------------------------------------------------------
// Per parent state: the 32-bit winner mask (single GPU / replicated table: read from the visited set; sharded
// table: OR of the masks pushed by all ranks), the bucket histogram of the children this rank will compose, and
// the prefix sums that number the children (scan.cuh; two counters at once: winners | free winners << 32).
//
// Load balance of the sharded march.  A child normally inherits the owner of its parent, because the plane rows
// it re-uses live in the parent's level buffer.  A child whose flipped neuron lies in hidden layer 1 re-uses
// nothing (all layers >= 2 are recomputed, layer 1 is the shared table): it is "free" and may go to any rank.
// Every rank computes the same per-owner load of the non-free children (integer work units) and the block that
// finishes last cuts the sequence of free children into `world` runs that level the loads:  free child f goes to
// the rank r with cuts[r] <= f < cuts[r + 1] (finalize_kernel).  Deterministic, replicated, no communication.
struct WinArgs {
    LayerOffs lo;
    int *next_counts;                  // [D + 3] bucket histogram of the next level's states owned by this rank
    int rank, world;
    uint32_t *wmask;                   // [S] out
    unsigned long long *win_base;      // [S] out: block-local exclusive prefix (winners | free winners << 32)
    FusedScan64 fs;
    int *zero_a, n_zero_a, *zero_b;    // small cursors of the next level's kernels, cleared here
    const unsigned char *own;          // sharded table: this rank's exchange block
    XchgLayout lay;
    int balance;                       // deal the free children (sharded modes only)
    unsigned long long *loads;         // [world] zero at launch, cleared again by the last block
    int *cuts;                         // [world + 1] out
    int unit_clip;                     // work of a child = 2 * (recomputed layers) + unit_clip
};

template <bool SHARDED_TABLE>
__global__ void __launch_bounds__(FS_THREADS) winners_scan_kernel(const LevelArgs a, const WinArgs w)
{
    pdl_enter();
    if (blockIdx.x == 0 && (int)threadIdx.x < w.n_zero_a) w.zero_a[threadIdx.x] = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0 && w.zero_b) *w.zero_b = 0;
    const int s0 = blockIdx.x * FS_TILE + threadIdx.x * FS_ITEMS;
    const int D = w.lo.D;
    unsigned long long v[FS_ITEMS], e[FS_ITEMS];
#pragma unroll 1
    for (int i = 0; i < FS_ITEMS; ++i) {
        const int s = s0 + i;
        v[i] = 0;
        if (s >= a.S) continue;
        const long long fo = a.face_off[a.lb + s];
        uint32_t m = 0;
        if (SHARDED_TABLE) {
            for (int r = 0; r < w.world; ++r) m |= reinterpret_cast<const uint32_t *>(w.own + w.lay.mask(r))[s];
        } else {
            const int k = int(a.face_off[a.lb + s + 1] - fo);
            for (int j = 0; j < k; ++j) {
                const int slot = a.cand_slot[(size_t)s * VSLOTS + j];
                if (slot != NO_SLOT && uint32_t(a.table.slots[slot]) == (CAND_TAG | cand_index(s, j))) m |= 1u << j;
            }
        }
        w.wmask[s] = m;
        if (m == 0u) continue;
        const int own_s = a.owner ? int(a.owner[a.lb + s]) : w.rank;
        unsigned n_free = 0;
        unsigned long long units = 0;
        for (uint32_t t = m; t; t &= t - 1) {
            const int ed = a.face_edges[fo + (__ffs(t) - 1)];
            int b = 1;
            while (b < D && ed >= w.lo.off[b + 1]) ++b;
            if (w.balance && b == 1) { ++n_free; continue; }
            units += (unsigned long long)(2 * (D - b) + w.unit_clip);
            if (own_s == w.rank) atomicAdd(w.next_counts + b, 1);
        }
        if (w.balance && units) atomicAdd(w.loads + own_s, units);
        v[i] = (unsigned long long)__popc(m) | ((unsigned long long)n_free << 32);
    }
    unsigned long long total = 0;
    const bool last = fused_scan_block(v, e, w.fs, &total);
#pragma unroll
    for (int i = 0; i < FS_ITEMS; ++i)
        if (s0 + i < a.S) w.win_base[s0 + i] = e[i];
    if (last && w.balance && threadIdx.x == 0) {
        const long long F = (long long)(total >> 32), W = 2 * (D - 1) + w.unit_clip;
        long long load[XCHG_MAX_WORLD], take[XCHG_MAX_WORLD], sum = 0;
        for (int r = 0; r < w.world; ++r) {
            load[r] = (long long)__ldcg(w.loads + r);
            w.loads[r] = 0;
            sum += load[r];
        }
        const long long target = (sum + F * W + w.world - 1) / w.world;
        long long rem = F;
        for (int r = 0; r < w.world; ++r) {
            long long need = target > load[r] ? (target - load[r]) / W : 0;
            take[r] = need < rem ? need : rem;
            rem -= take[r];
        }
        int c = 0;
        for (int r = 0; r < w.world; ++r) {      // what rounding left over: dealt evenly
            take[r] += rem / w.world + (r < rem % w.world ? 1 : 0);
            w.cuts[r] = c;
            c += (int)take[r];
        }
        w.cuts[w.world] = c;
        atomicAdd(w.next_counts + 1, (int)take[w.rank]);
    }
}

}  // namespace amb
