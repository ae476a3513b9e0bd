// frontier.cuh -- device-resident BFS frontier, visited set and neighbour enumeration.
//
// Replaces reference backend/inc/states.h + src/states.cu + inc/hash.h (LIFO stack of states stored
// as one double per bit, 32-bit-fingerprint-only hash set) and inc/kernel.h:1390-1506
// (inferNewStates: per-thread serial copy of L doubles per edge + device-side cudaMemcpyAsync).
//
// Design
//   * keys are bit-packed, padded to whole uint4 (KW4 = ceil(L/128) 16-byte words), stored once in an
//     append-only arena; state id = arena index.  A BFS level is a contiguous id range.
//   * visited set = open addressing over 64-bit slots {fingerprint:32 | value:32}; a hit on the
//     fingerprint is always confirmed by a full-key compare (group-cooperative uint4 loads), so two
//     different patterns can never be merged (the reference compares fingerprints only).
//   * the key hash is additive over words (common.cuh word_mix), so the hash of "parent key with bit
//     e flipped" costs O(1) from the parent's stored hash; candidate keys are never materialised --
//     a candidate is the pair (parent state, edge slot); only winners write a key.
//   * two-phase, deterministic insertion: phase 1 claims a slot with CAS(EMPTY -> TAG|cand) or
//     lowers it with atomicMin when the slot already holds a candidate with the same key, so the
//     winner of every new key is its smallest candidate index regardless of thread timing; phase 2
//     counts winners per parent, a prefix sum assigns new state ids in (parent, edge-slot) order,
//     and the winners write key / hash / parent / edge / seed point and replace the tag by the id.
//     State numbering -- and therefore face and vertex numbering -- is reproducible run to run
//     (the reference's follows atomicAdd order).
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"
#include "scan.cuh"

namespace amb {
namespace cg = cooperative_groups;

struct TableRef {
    unsigned long long *slots;
    uint32_t mask;
};

// ---- key helpers ----------------------------------------------------------------------------

__device__ __forceinline__ void flip_bit(uint4 &v, int bit /*0..127*/)
{
    const uint32_t m = 1u << (bit & 31);
    switch (bit >> 5) {
        case 0: v.x ^= m; break;
        case 1: v.y ^= m; break;
        case 2: v.z ^= m; break;
        default: v.w ^= m; break;
    }
}

// (A with bit eA flipped) == (B with bit eB flipped) ?   e < 0: no flip
template <int G>
__device__ __forceinline__ bool keys_equal(const cg::thread_block_tile<G> &tile, const uint4 *A, int eA, const uint4 *B,
                                           int eB, int kw4)
{
    bool diff = false;
    for (int q = tile.thread_rank(); q < kw4; q += G) {
        uint4 x = A[q], y = B[q];
        if (eA >= 0 && (eA >> 7) == q) flip_bit(x, eA & 127);
        if (eB >= 0 && (eB >> 7) == q) flip_bit(y, eB & 127);
        diff |= ((x.x ^ y.x) | (x.y ^ y.y) | (x.z ^ y.z) | (x.w ^ y.w)) != 0u;
    }
    return !tile.any(diff);
}

__device__ __forceinline__ uint64_t hash_flip(uint64_t h, const uint32_t *key, int e)
{
    const uint32_t w = uint32_t(e) >> 5, old = key[w];
    return h - word_mix(w, old) + word_mix(w, old ^ (1u << (e & 31)));
}

// ---- level description shared by the three frontier kernels ------------------------------------

struct LevelArgs {
    const uint32_t *keys;                 // arena base
    const unsigned long long *hsum;       // arena base
    const long long *face_off;            // [n_states + 1]
    const int *face_edges;                // [corners]
    const double *face_xyz;               // [corners][3]
    int kw, kw4, L;
    int lb, S;                            // level = states [lb, lb + S)
    TableRef table;
    int *cand_slot;                       // [S][VSLOTS] slot claimed/matched by candidate, NO_SLOT otherwise
    uint32_t *nwin;                       // [S] winners per parent
    const uint32_t *win_base;             // [S] exclusive scan of nwin
    unsigned long long *counters;
    // finalize
    uint32_t *keys_w;
    unsigned long long *hsum_w;
    int *parent, *via_edge;
    double *seedpt;
    uint8_t *owner;                       // sharded mode: rank that composes/clips the state (children inherit)
    int n_states;                         // states before this level's children are appended
    // sharded visited set (xchg.cuh): a rank inserts only the candidates whose key hash it owns; the winners of
    // all ranks arrive as one 32-bit mask per parent (bit j = edge slot j discovered a new state)
    int world, rank;                      // world <= 1: single table, winners are read from the table
    const uint32_t *wmask;
    // fused winner kernel (xchg.cuh winners_scan_kernel): id offset = win_off64[s / FS_TILE] + win_base64[s], low half
    // = index among the level's children, high half = index among its free children (dealt by `cuts`)
    const unsigned long long *win_base64, *win_off64;
    int cap_states;                       // finalize: arena capacity (children beyond it are skipped; the host re-runs)
    const int *cuts;                      // [world + 1] or nullptr: free child f -> rank r with cuts[r] <= f < cuts[r + 1]
    int free_below_bit;                   // a child is free when its flipped neuron index is below this (first bit of layer 2)
    int n_ranks;
};

__device__ __forceinline__ int key_hash_owner(uint64_t h, int world) { return int(((h >> 32) * (uint64_t)world) >> 32); }

// candidate (parent sid, edge slot j) <-> 31-bit index within the level
__device__ __forceinline__ uint32_t cand_index(int s_local, int j) { return (uint32_t(s_local) << 5) | uint32_t(j); }

// phase 1: one group of G lanes per parent state.  Lane j owns candidate j (the state across polygon edge j): it
// hashes and probes on its own, so the k dependent probe chains of a state run side by side instead of one
// after the other; only a fingerprint match needs the whole group (cooperative full-key compare), and those are
// resolved one at a time.  cand_slot[s][j] is written for j < k only (nobody reads the rest).
template <int G>
__global__ void expand_insert_kernel(const LevelArgs a)
{
    pdl_enter();
    cg::thread_block_tile<G> tile = cg::tiled_partition<G>(cg::this_thread_block());
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    if (s >= a.S) return;
    const int sid = a.lb + s;
    const long long fo = a.face_off[sid];
    const int k = int(a.face_off[sid + 1] - fo);
    const uint32_t *key = a.keys + (size_t)sid * a.kw;
    const uint4 *key4 = reinterpret_cast<const uint4 *>(key);
    const uint64_t h0 = a.hsum[sid];
    int ncand = 0;
    for (int j0 = 0; j0 < k; j0 += G) {
        const int j = j0 + (int)tile.thread_rank();
        int result = NO_SLOT, e = -1;
        bool active = false;
        uint32_t fp = 0, slot = 0;
        uint64_t mine = 0;
        if (j < k) {
            e = a.face_edges[fo + j];
            if (e >= 0 && e < a.L) {
                const uint64_t h = hash_flip(h0, key, e);
                fp = slot_fp(h);
                mine = (uint64_t(fp) << 32) | CAND_TAG | cand_index(s, j);
                slot = uint32_t(h) & a.table.mask;
                active = (a.world <= 1) || key_hash_owner(h, a.world) == a.rank;
            } else {
                e = -1;
            }
        }
        ncand += __popc(tile.ballot(e >= 0));
        while (tile.any(active)) {
            unsigned long long v = 0;
            bool pend = false;
            if (active) {
                for (;;) {                                   // probe until the slot is claimed or a fingerprint matches
                    v = a.table.slots[slot];
                    if (v == SLOT_EMPTY) {
                        v = atomicCAS(a.table.slots + slot, SLOT_EMPTY, mine);
                        if (v == SLOT_EMPTY) { result = int(slot); active = false; break; }
                    }
                    if (uint32_t(v >> 32) == fp) { pend = true; break; }
                    slot = (slot + 1) & a.table.mask;
                }
            }
            unsigned pm = tile.ballot(pend);
            while (pm) {                                     // full-key compares, one candidate at a time
                const int src = __ffs(pm) - 1;
                pm &= pm - 1;
                const unsigned long long vs = tile.shfl(v, src);
                const int es = tile.shfl(e, src);
                const uint32_t low = uint32_t(vs);
                bool same;
                if (low & CAND_TAG) {                        // another candidate of this level
                    const uint32_t ci = low & ~CAND_TAG;
                    const int sid2 = a.lb + int(ci >> 5);
                    const int e2 = a.face_edges[a.face_off[sid2] + (ci & 31u)];
                    same = keys_equal<G>(tile, key4, es, reinterpret_cast<const uint4 *>(a.keys + (size_t)sid2 * a.kw), e2,
                                         a.kw4);
                } else {                                     // an already visited state
                    same = keys_equal<G>(tile, key4, es, reinterpret_cast<const uint4 *>(a.keys + (size_t)low * a.kw), -1,
                                         a.kw4);
                }
                if ((int)tile.thread_rank() == src) {
                    if (same) {
                        if (low & CAND_TAG) {                // same key: the smaller candidate index wins the slot
                            atomicMin(a.table.slots + slot, mine);
                            result = int(slot);
                        }
                        active = false;
                    } else {
                        slot = (slot + 1) & a.table.mask;
                    }
                }
            }
        }
        if (j < k) a.cand_slot[(size_t)s * VSLOTS + j] = result;
    }
    if (tile.thread_rank() == 0 && ncand) atomicAdd(a.counters + CNT_CANDIDATES, (unsigned long long)ncand);
}

// phase 2b: winners append their state (group of G lanes per parent)
template <int G>
__global__ void finalize_kernel(const LevelArgs a)
{
    pdl_enter();
    cg::thread_block_tile<G> tile = cg::tiled_partition<G>(cg::this_thread_block());
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    if (s >= a.S) return;
    if (a.wmask ? (a.wmask[s] == 0u) : (a.nwin[s] == 0)) return;
    const int sid = a.lb + s;
    const long long fo = a.face_off[sid];
    const int k = int(a.face_off[sid + 1] - fo);
    const uint4 *key4 = reinterpret_cast<const uint4 *>(a.keys + (size_t)sid * a.kw);
    int nid, fidx = 0;
    if (a.win_base64) {
        const unsigned long long pr = a.win_base64[s] + a.win_off64[s / FS_TILE];
        nid = a.n_states + int(uint32_t(pr));
        fidx = int(pr >> 32);
    } else {
        nid = a.n_states + int(a.win_base[s]);
    }
    const uint32_t wm = a.wmask ? a.wmask[s] : 0u;
    for (int j = 0; j < k; ++j) {
        const int slot = a.cand_slot[(size_t)s * VSLOTS + j];
        unsigned long long v = 0;
        if (a.wmask) {                      // winners of all ranks; `slot` is valid on the key's hash owner only
            if (!((wm >> j) & 1u)) continue;
            if (slot != NO_SLOT) v = a.table.slots[slot];
        } else {
            if (slot == NO_SLOT) continue;
            v = a.table.slots[slot];
            if (uint32_t(v) != (CAND_TAG | cand_index(s, j))) continue;
        }
        const int e = a.face_edges[fo + j];
        if (nid >= a.cap_states) {              // speculative launch with too small an arena: skipped, re-run by the host
            ++nid;
            if (a.cuts != nullptr && e < a.free_below_bit) ++fidx;
            continue;
        }
        uint4 *dst = reinterpret_cast<uint4 *>(a.keys_w + (size_t)nid * a.kw);
        for (int q = tile.thread_rank(); q < a.kw4; q += G) {
            uint4 x = key4[q];
            if ((e >> 7) == q) flip_bit(x, e & 127);
            dst[q] = x;
        }
        // seed point of the child = mid-point of the shared edge; extent of the parent's polygon around it = a hint for
        // the size of the child's polygon (clip.cuh starts from a square of a few times this size and retries if it was
        // too small).  The corners are spread over the lanes (max is exact in any order).
        const int j2 = (j + 1 == k) ? 0 : j + 1;
        const double *p = a.face_xyz + (size_t)(fo + j) * 3, *q2 = a.face_xyz + (size_t)(fo + j2) * 3;
        const double mx = 0.5 * (p[0] + q2[0]), my = 0.5 * (p[1] + q2[1]), mz = 0.5 * (p[2] + q2[2]);
        double ext = 0.0;
        for (int c = tile.thread_rank(); c < k; c += G) {
            const double *v = a.face_xyz + (size_t)(fo + c) * 3;
            ext = fmax(ext, fmax(fabs(v[0] - mx), fmax(fabs(v[1] - my), fabs(v[2] - mz))));
        }
        for (int o = G / 2; o > 0; o >>= 1) ext = fmax(ext, tile.shfl_xor(ext, o));
        if (tile.thread_rank() == 0) {
            a.hsum_w[nid] = hash_flip(a.hsum[sid], a.keys + (size_t)sid * a.kw, e);
            a.parent[nid] = sid;
            a.via_edge[nid] = e;
            if (a.owner) {
                int own = a.owner[sid];
                if (a.cuts != nullptr && e < a.free_below_bit) {      // free child: dealt to level the loads
                    own = 0;
                    while (own + 1 < a.n_ranks && fidx >= a.cuts[own + 1]) ++own;
                }
                a.owner[nid] = uint8_t(own);
            }
            a.seedpt[(size_t)nid * 4 + 0] = mx;
            a.seedpt[(size_t)nid * 4 + 1] = my;
            a.seedpt[(size_t)nid * 4 + 2] = mz;
            a.seedpt[(size_t)nid * 4 + 3] = ext;
            if (slot != NO_SLOT) a.table.slots[slot] = (v & 0xFFFFFFFF00000000ull) | uint32_t(nid);
        }
        ++nid;
        if (a.cuts != nullptr && e < a.free_below_bit) ++fidx;
    }
}

// ---- explicit-key insertion (seed states; later: records received from other GPUs) --------------

struct XArgs {
    const uint32_t *xkeys;                // [N][kw]
    const unsigned long long *xh;         // [N]
    const double *xpt;                    // [N][3]
    const int *xparent, *xvia;            // [N] or nullptr (seeds: -1)
    int N, kw, kw4;
    const uint32_t *keys;                 // arena (visited states)
    TableRef table;
    int *x_slot;                          // [N]
    uint32_t *xwin;                       // [N] 0/1
    const uint32_t *win_base;             // [N]
    uint32_t *keys_w;
    unsigned long long *hsum_w;
    int *parent, *via_edge;
    double *seedpt;
    int n_states;
};

template <int G>
__global__ void x_insert_kernel(const XArgs a)
{
    cg::thread_block_tile<G> tile = cg::tiled_partition<G>(cg::this_thread_block());
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    if (i >= a.N) return;
    const uint4 *key4 = reinterpret_cast<const uint4 *>(a.xkeys + (size_t)i * a.kw);
    const uint64_t h = a.xh[i];
    const uint32_t fp = slot_fp(h);
    const uint64_t mine = (uint64_t(fp) << 32) | CAND_TAG | uint32_t(i);
    uint32_t slot = uint32_t(h) & a.table.mask;
    int result = NO_SLOT;
    for (;;) {
        unsigned long long v = 0;
        if (tile.thread_rank() == 0) {
            v = a.table.slots[slot];
            if (v == SLOT_EMPTY) v = atomicCAS(a.table.slots + slot, SLOT_EMPTY, mine);
        }
        v = tile.shfl(v, 0);
        if (v == SLOT_EMPTY) { result = int(slot); break; }
        if (uint32_t(v >> 32) == fp) {
            const uint32_t low = uint32_t(v);
            if (low & CAND_TAG) {
                const uint32_t i2 = low & ~CAND_TAG;
                if (keys_equal<G>(tile, key4, -1, reinterpret_cast<const uint4 *>(a.xkeys + (size_t)i2 * a.kw), -1, a.kw4)) {
                    if (tile.thread_rank() == 0) atomicMin(a.table.slots + slot, mine);
                    result = int(slot);
                    break;
                }
            } else if (keys_equal<G>(tile, key4, -1, reinterpret_cast<const uint4 *>(a.keys + (size_t)low * a.kw), -1,
                                     a.kw4)) {
                break;
            }
        }
        slot = (slot + 1) & a.table.mask;
    }
    if (tile.thread_rank() == 0) a.x_slot[i] = result;
}

__global__ void x_count_kernel(const XArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.N) return;
    const int slot = a.x_slot[i];
    a.xwin[i] = (slot != NO_SLOT && uint32_t(a.table.slots[slot]) == (CAND_TAG | uint32_t(i))) ? 1u : 0u;
}

template <int G>
__global__ void x_finalize_kernel(const XArgs a)
{
    cg::thread_block_tile<G> tile = cg::tiled_partition<G>(cg::this_thread_block());
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    if (i >= a.N) return;
    if (!a.xwin[i]) return;
    const int nid = a.n_states + int(a.win_base[i]);
    const uint4 *src = reinterpret_cast<const uint4 *>(a.xkeys + (size_t)i * a.kw);
    uint4 *dst = reinterpret_cast<uint4 *>(a.keys_w + (size_t)nid * a.kw);
    for (int q = tile.thread_rank(); q < a.kw4; q += G) dst[q] = src[q];
    if (tile.thread_rank() == 0) {
        a.hsum_w[nid] = a.xh[i];
        a.parent[nid] = a.xparent ? a.xparent[i] : -1;
        a.via_edge[nid] = a.xvia ? a.xvia[i] : -1;
        a.seedpt[(size_t)nid * 4 + 0] = a.xpt[(size_t)i * 3 + 0];
        a.seedpt[(size_t)nid * 4 + 1] = a.xpt[(size_t)i * 3 + 1];
        a.seedpt[(size_t)nid * 4 + 2] = a.xpt[(size_t)i * 3 + 2];
        a.seedpt[(size_t)nid * 4 + 3] = 0.0;                  // no size hint for a seed state
        const int slot = a.x_slot[i];
        a.table.slots[slot] = (a.table.slots[slot] & 0xFFFFFFFF00000000ull) | uint32_t(nid);
    }
}

__global__ void seed_owner_kernel(uint8_t *owner, int n, int world)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) owner[i] = uint8_t(i % world);
}

// bool (N x L bytes) -> packed keys + additive hash; one thread per (seed, 32-bit word)
__global__ void pack_states_kernel(const uint8_t *states, int N, int L, int kw, uint32_t *xkeys)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)N * kw) return;
    const int i = int(t / kw), w = int(t % kw);
    uint32_t v = 0;
    const uint8_t *row = states + (size_t)i * L;
    for (int b = 0; b < 32; ++b) {
        const int j = w * 32 + b;
        if (j < L && row[j]) v |= 1u << b;
    }
    xkeys[t] = v;
}

__global__ void hash_keys_kernel(const uint32_t *xkeys, int N, int kw, unsigned long long *xh)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    uint64_t h = 0;
    for (int w = 0; w < kw; ++w) h += word_mix(uint32_t(w), xkeys[(size_t)i * kw + w]);
    xh[i] = h;
}

// rebuild the table from the stored hashes after it has grown (all stored keys are distinct)
__global__ void rehash_kernel(const unsigned long long *hsum, int n, TableRef table, int world, int rank)
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    const uint64_t h = hsum[id];
    if (world > 1 && key_hash_owner(h, world) != rank) return;      // sharded visited set: another rank's key
    const unsigned long long mine = (uint64_t(slot_fp(h)) << 32) | uint32_t(id);
    uint32_t slot = uint32_t(h) & table.mask;
    while (atomicCAS(table.slots + slot, SLOT_EMPTY, mine) != SLOT_EMPTY) slot = (slot + 1) & table.mask;
}

// lookup of (key of state sid with the neuron bits e1, e2 cleared/set as given) -- used by stitching.
// Returns the state id or -1.  `want1/want2`: desired value of bit e1/e2 (e < 0: ignored).
template <int G>
__device__ __forceinline__ int lookup_sibling(const cg::thread_block_tile<G> &tile, const uint32_t *keys,
                                              const unsigned long long *hsum, int kw, int kw4, TableRef table, int sid,
                                              int e1, int e2)
{
    const uint32_t *key = keys + (size_t)sid * kw;
    uint64_t h = hsum[sid];
    if (e1 >= 0 && e2 >= 0 && (e1 >> 5) == (e2 >> 5)) {
        const uint32_t w = uint32_t(e1) >> 5, old = key[w];
        h = h - word_mix(w, old) + word_mix(w, old ^ (1u << (e1 & 31)) ^ (1u << (e2 & 31)));
    } else {
        if (e1 >= 0) h = hash_flip(h, key, e1);
        if (e2 >= 0) h = hash_flip(h, key, e2);
    }
    const uint32_t fp = slot_fp(h);
    uint32_t slot = uint32_t(h) & table.mask;
    const uint4 *key4 = reinterpret_cast<const uint4 *>(key);
    for (;;) {
        const unsigned long long v = table.slots[slot];
        if (v == SLOT_EMPTY) return -1;
        if (uint32_t(v >> 32) == fp) {
            const int t = int(uint32_t(v));
            const uint4 *B = reinterpret_cast<const uint4 *>(keys + (size_t)t * kw);
            bool diff = false;
            for (int q = tile.thread_rank(); q < kw4; q += G) {
                uint4 x = key4[q];
                if (e1 >= 0 && (e1 >> 7) == q) flip_bit(x, e1 & 127);
                if (e2 >= 0 && (e2 >> 7) == q) flip_bit(x, e2 & 127);
                const uint4 y = B[q];
                diff |= ((x.x ^ y.x) | (x.y ^ y.y) | (x.z ^ y.z) | (x.w ^ y.w)) != 0u;
            }
            if (!tile.any(diff)) return t;
        }
        slot = (slot + 1) & table.mask;
    }
}

}  // namespace amb
