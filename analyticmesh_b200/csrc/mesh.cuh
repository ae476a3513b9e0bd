// mesh.cuh -- prefix sums, face compaction, shared-vertex stitching and indexing.
//
// Replaces reference backend/inc/polymesh.h:120-346 (append_verts / uniqueVertHash /
// indexingVertices / indexingFaces).  The reference merges vertices by hashing coordinates
// truncated to int(coord * 1e9) (fingerprint-only); here a vertex is identified
// topologically: the corner of state s between constraint edges (ea, eb) is the same vertex as the
// matching corner of the up to three sibling states that differ from s in the bits ea / eb.  The
// owner of a vertex is the sibling with the smallest (bit(e_hi), bit(e_lo)) pattern that exists in
// the visited set and has the corner; siblings are found with exact (full-key) visited-set lookups,
// so the stitching involves no coordinate comparison and no probabilistic fingerprint.
#pragma once
#include "frontier.cuh"

namespace amb {

// ---- exclusive prefix sum of uint32 (three-kernel, recursive on the block sums) -----------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void scan_reduce_kernel(const uint32_t *in, int n, uint32_t *block_sums)
{
    __shared__ uint32_t sh[SCAN_THREADS / 32];
    const int base = blockIdx.x * SCAN_TILE;
    uint32_t v = 0;
    for (int i = threadIdx.x; i < SCAN_TILE; i += SCAN_THREADS)
        if (base + i < n) v += in[base + i];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < SCAN_THREADS / 32; ++w) t += sh[w];
        block_sums[blockIdx.x] = t;
    }
}

// out[i] = block_off[block] + exclusive prefix within the block; optionally total -> *total
__global__ void scan_apply_kernel(const uint32_t *in, int n, const uint32_t *block_off, uint32_t *out,
                                  unsigned long long *total)
{
    __shared__ uint32_t sh[SCAN_THREADS / 32];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0u;
        sum += v[i];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = sum;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < warp; ++w) woff += sh[w];
    uint32_t run = (block_off ? block_off[blockIdx.x] : 0u) + woff + inc - sum;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
    if (total != nullptr && base <= n - 1 && n - 1 < base + SCAN_ITEMS) *total = run;
}

// ---- face compaction: per-chunk scratch (stride VSLOTS) -> global CSR ----------------------------
struct CompactArgs {
    const int *cnt;               // [S]
    const uint32_t *off;          // [S] exclusive scan of cnt (block-local when block_off is given, scan.cuh)
    const uint32_t *block_off;    // offsets of the FS_TILE blocks, or nullptr
    const int *edges;             // [S][VSLOTS]
    const double *verts;          // [S][VSLOTS][3]
    int S, sid0;
    long long *face_off;          // global [n_states + 1]
    int *face_edges;
    double *face_xyz;
    unsigned long long *counters; // CNT_CORNERS = running total before this chunk
    const uint8_t *owner;         // sharded mode: only the owner's polygons are copied (others arrive by all-reduce)
    int rank;
};

__global__ void compact_faces_kernel(const CompactArgs a)
{
    pdl_enter();
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (s >= a.S) return;
    const long long base = (long long)a.counters[CNT_CORNERS] + a.off[s] + (a.block_off ? a.block_off[s / FS_TILE] : 0u);
    const int k = a.cnt[s];
    if (lane == 0) {
        a.face_off[a.sid0 + s] = base;
        if (s == a.S - 1) a.face_off[a.sid0 + s + 1] = base + k;
        if (k > 0) atomicAdd(a.counters + CNT_FACES, 1ull);
    }
    if (lane < k && (a.owner == nullptr || a.owner[a.sid0 + s] == a.rank)) {
        a.face_edges[base + lane] = a.edges[(size_t)s * VSLOTS + lane];
        const double *v = a.verts + ((size_t)s * VSLOTS + lane) * 3;
        double *o = a.face_xyz + (size_t)(base + lane) * 3;
        o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
    }
}

// running corner total += corners of the chunk just compacted (after every block has read the old base)
__global__ void bump_counters_kernel(unsigned long long *counters)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) counters[CNT_CORNERS] += counters[CNT_CHUNK_CORNERS];
}

// ---- stitching ------------------------------------------------------------------------------
struct StitchArgs {
    const uint32_t *keys;
    const unsigned long long *hsum;
    const long long *face_off;
    const int *face_edges;
    int kw, kw4, L, n_states;
    TableRef table;
    long long *owner;             // [corners] owner corner index
    unsigned long long *counters;
};

__device__ __forceinline__ int key_bit(const uint32_t *key, int j) { return (key[j >> 5] >> (j & 31)) & 1u; }

// corner of state t whose two edges are {ea, eb}; -1 if none
__device__ __forceinline__ int find_corner(const long long *face_off, const int *face_edges, int t, int ea, int eb)
{
    const long long fo = face_off[t];
    const int k = int(face_off[t + 1] - fo);
    for (int i = 0; i < k; ++i) {
        const int ga = face_edges[fo + (i == 0 ? k - 1 : i - 1)], gb = face_edges[fo + i];
        if ((ga == ea && gb == eb) || (ga == eb && gb == ea)) return i;
    }
    return -1;
}

template <int G>
__global__ void stitch_owner_kernel(const StitchArgs a)
{
    cg::thread_block_tile<G> tile = cg::tiled_partition<G>(cg::this_thread_block());
    const int sid = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    if (sid >= a.n_states) return;
    const long long fo = a.face_off[sid];
    const int k = int(a.face_off[sid + 1] - fo);
    const uint32_t *key = a.keys + (size_t)sid * a.kw;
    for (int i = 0; i < k; ++i) {
        const int ea = a.face_edges[fo + (i == 0 ? k - 1 : i - 1)], eb = a.face_edges[fo + i];
        const int na = (ea < a.L) ? ea : -1, nb = (eb < a.L) ? eb : -1;
        const int e_lo = (na >= 0 && nb >= 0) ? min(na, nb) : max(na, nb);   // single neuron edge -> e_lo
        const int e_hi = (na >= 0 && nb >= 0) ? max(na, nb) : -1;
        const int b_lo = (e_lo >= 0) ? key_bit(key, e_lo) : 0;
        const int b_hi = (e_hi >= 0) ? key_bit(key, e_hi) : 0;
        const int self_pat = (b_hi << 1) | b_lo;
        long long own = fo + i;
        for (int pat = 0; pat < self_pat; ++pat) {
            const int f_lo = ((pat & 1) != b_lo), f_hi = (((pat >> 1) & 1) != b_hi);
            if ((f_lo && e_lo < 0) || (f_hi && e_hi < 0)) continue;
            const int t = lookup_sibling<G>(tile, a.keys, a.hsum, a.kw, a.kw4, a.table, sid, f_lo ? e_lo : -1,
                                            f_hi ? e_hi : -1);
            if (t < 0) continue;
            const int ci = find_corner(a.face_off, a.face_edges, t, ea, eb);
            if (ci >= 0) { own = a.face_off[t] + ci; break; }
            if (tile.thread_rank() == 0) atomicAdd(a.counters + CNT_STITCH_MISS, 1ull);
        }
        if (tile.thread_rank() == 0) a.owner[fo + i] = own;
    }
}

__global__ void owner_flags_kernel(const long long *owner, long long n, uint32_t *flag)
{
    const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (c < n) flag[c] = (owner[c] == c) ? 1u : 0u;
}

// vertex ids for every corner + unique vertex coordinates (scaled)
__global__ void index_corners_kernel(const long long *owner, const uint32_t *vid_of_corner, const uint32_t *flag,
                                     long long n, const double *xyz, double scale, double cx, double cy, double cz,
                                     int *corner_vid, double *vertices)
{
    const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (c >= n) return;
    corner_vid[c] = int(vid_of_corner[owner[c]]);
    if (flag[c]) {
        const uint32_t v = vid_of_corner[c];
        vertices[(size_t)v * 3 + 0] = xyz[(size_t)c * 3 + 0] * scale + cx;
        vertices[(size_t)v * 3 + 1] = xyz[(size_t)c * 3 + 1] * scale + cy;
        vertices[(size_t)v * 3 + 2] = xyz[(size_t)c * 3 + 2] * scale + cz;
    }
}

// ---- digests + topology self-check (bench.py prints them, so that two runs -- 1 GPU / N GPUs, two engine
// builds -- can be compared without moving the 11 GB of keys and polygons to the host) ---------------------
// Positional checksum of an array of 64-bit words:  sum_i splitmix64(word_i + splitmix64(i + salt))  mod 2^64.
// Any changed word and any two exchanged positions change the sum (up to a 2^-64 coincidence).
__device__ __forceinline__ void digest_accumulate(unsigned long long *acc, uint64_t v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(acc, (unsigned long long)v);
}

__global__ void digest_words_kernel(const void *data, long long n_words, int word_bytes, uint64_t salt,
                                    unsigned long long *acc)
{
    uint64_t sum = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_words; i += stride) {
        const uint64_t w = (word_bytes == 8) ? static_cast<const uint64_t *>(data)[i]
                                             : (uint64_t) static_cast<const uint32_t *>(data)[i];
        sum += splitmix64(w + splitmix64((uint64_t)i + salt));
    }
    digest_accumulate(acc, sum);
}

// Order-independent digest of the region set (one warp per state): the per-state value depends only on the
// state's key (and its edge loop / vertex bits), the states' values are mixed once more and summed, so the
// result does not depend on the order in which the states were visited.
//   acc[0]: keys + edge loops ("topology": must equal the CPU oracle's value)   acc[1]: + vertex bits
__global__ void digest_states_kernel(const uint32_t *keys, int kw, int n_words, const long long *face_off,
                                     const int *face_edges, const double *face_xyz, long long n_states,
                                     unsigned long long *acc)
{
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    uint64_t topo = 0, full = 0;
    for (long long s = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5); s < n_states; s += warps) {
        uint64_t h = 0, g = 0;
        for (int w = lane; w < n_words; w += 32) h += word_mix((uint32_t)w, keys[(size_t)s * kw + w]);
        const long long fo = face_off[s];
        const int k = int(face_off[s + 1] - fo);
        if (lane < k) {
            h += splitmix64((uint64_t)(uint32_t)face_edges[fo + lane] + splitmix64(0xE0000000ull + (uint64_t)lane));
            for (int c = 0; c < 3; ++c)
                g += splitmix64((uint64_t)__double_as_longlong(face_xyz[(size_t)(fo + lane) * 3 + c]) +
                                splitmix64(0xF0000000ull + (uint64_t)(3 * lane + c)));
        }
        for (int o = 16; o > 0; o >>= 1) {
            h += __shfl_xor_sync(0xFFFFFFFFu, h, o);
            g += __shfl_xor_sync(0xFFFFFFFFu, g, o);
        }
        h += splitmix64((uint64_t)k);
        if (lane == 0) {
            topo += splitmix64(h);
            full += splitmix64(h + splitmix64(g));
        }
    }
    if (lane == 0 && (topo | full)) {
        atomicAdd(acc + 0, (unsigned long long)topo);
        atomicAdd(acc + 1, (unsigned long long)full);
    }
}

// Edge incidence: the polygon edge of state s carried by neuron e is shared with the state whose key differs
// in bit e.  out[0] boundary edges (extra constraints), out[1] neuron edges whose sibling state was visited
// and has the same edge (each shared edge is counted from both sides), out[2] sibling state never visited,
// out[3] sibling visited but without that edge.
template <int G>
__global__ void edge_incidence_kernel(const StitchArgs a, unsigned long long *out)
{
    cg::thread_block_tile<G> tile = cg::tiled_partition<G>(cg::this_thread_block());
    const int sid = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    if (sid >= a.n_states) return;
    const long long fo = a.face_off[sid];
    const int k = int(a.face_off[sid + 1] - fo);
    unsigned n[4] = {0, 0, 0, 0};
    for (int i = 0; i < k; ++i) {
        const int e = a.face_edges[fo + i];
        if (e < 0 || e >= a.L) { ++n[0]; continue; }
        const int t = lookup_sibling<G>(tile, a.keys, a.hsum, a.kw, a.kw4, a.table, sid, e, -1);
        if (t < 0) { ++n[2]; continue; }
        const long long fo2 = a.face_off[t];
        const int k2 = int(a.face_off[t + 1] - fo2);
        bool has = false;
        for (int j = 0; j < k2; ++j) has |= (a.face_edges[fo2 + j] == e);
        ++n[has ? 1 : 3];
    }
    if (tile.thread_rank() == 0)
        for (int c = 0; c < 4; ++c)
            if (n[c]) atomicAdd(out + c, (unsigned long long)n[c]);
}

}  // namespace amb
