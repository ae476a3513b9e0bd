// clip.cuh -- warp-cooperative extraction of the face polygon of one linear region.
//
// Replaces the reference's ordering + start search + vertex pivoting + orientation stages
// (backend/inc/process.h:208-321, inc/kernel.h:799-1385): there a host loop launches ~10 kernels
// and does 4 blocking scalar copies per pivot step; here one warp owns one state and clips the
// zero-level plane of the region's affine function against all C half-spaces in ONE pass over the
// plane rows (each row is read exactly once, 64 rows per warp iteration, coalesced 1 KiB loads):
//
//   lane = plane:  d = sigma (a . v_j + c) for the <= 32 current vertices (broadcast from shared
//                  memory); a plane "cuts" if some vertex has d * rsqrt(|a|^2) > EPS -- the
//                  reference's feasibility predicate (inc/kernel.h:760-768, 1063-1071)
//   ballot         the cutting planes, applied one at a time in row order
//   lane = vertex: outside mask by ballot; the outside vertices of a convex polygon form one cyclic
//                  run [a..b]; edges a and b+1 are cut, the run is replaced by the new edge
//   new vertices   are the Cramer solution of (new plane, old edge plane, level plane) -- the
//                  reference's vertex formula (inc/kernel.h:555-573) -- never an interpolation, so
//                  final coordinates do not depend on the clipping order.
//
// The polygon starts as a square (artificial edges, ids < 0) in the level plane centred on the state's
// seed point and sized from the parent's polygon; a final polygon that still has an artificial edge is
// retried with a larger square and finally dropped as unbounded (counted).  The loop is kept
// counter-clockwise around +w_equ, which is the reference's output orientation (inc/kernel.h:1349-1353);
// flip_insideout reverses it.
//
// Streaming: the constraints of a state are four contiguous row segments -- the shared layer-1 table, the
// rows the state inherits from its parent (read from the PARENT's level buffer and stored into the state's
// own buffer on the way: this replaces a separate copy kernel), the state's own rows, the extra
// constraints.  Each segment is streamed in 64-row blocks through a per-warp cp.async ring; a copy
// instruction moves 32 consecutive 16-byte pieces (512 B, fully coalesced), every lane then tests two rows
// against the polygon's vertices (each vertex load serves both rows).
//
// Output convention (shared with the oracle wrapper): vertices v_0..v_{k-1}; edge id g_i is the
// constraint that carries the segment v_i -> v_{i+1}; the cycle is rotated so that g_0 is minimal.
#pragma once
#include "common.cuh"

namespace amb {

struct ClipArgs {
    const uint32_t *keys;     // key of state 0 of the chunk
    int kw;
    const double *P1;         // [n1][4] rows of hidden layer 1 (shared by all states)
    int n1;
    const double *P;          // [S][R][4] rows of hidden layers >= 2
    long long p_stride;       // doubles per state (= 4 R)
    const double *equ;        // [S][4]
    const double *extra;      // [E][4]
    int L, E, S, flip;
    const double *seedpt;     // [S][4] seed point (x, y, z) and size hint (0 = none) of state 0 of the chunk
    const int *idx;           // optional list of the states to process (sharded mode); S = its length
    // read-through: rows of layers 2..bucket are read from the parent's level buffer and written to the
    // state's own buffer on the way (they replace copy_parent_rows_kernel); nullptr = everything is own
    const double *P_prev;     // [S_prev][R][4]
    double *P_own;            // = P, writable
    const int *bucket;        // per state of the level
    const int *parent;        // indexed by global state id
    int lb, prev_lb;
    LayerOffs lo;
    int *out_cnt;             // [S]
    int *out_edges;           // [S][VSLOTS]
    double *out_verts;        // [S][VSLOTS][3]
    unsigned long long *counters;
    int tile_stride, tile_offset, tile;   // chained launches (compose.cuh chain_position); stride <= 1: the whole list
    int rows_by_slot;         // sharded march: own rows live at the list position (= permutation slot) ...
    const int *prev_slot_of;  // ... and the parent's rows at prev_slot_of[parent - prev_lb]
    int push_on;              // sharded march: the polygon goes straight into every rank's exchange block (xchg.cuh) ...
    PeerPush push;            // ... instead of the local scratch (out_cnt / out_edges / out_verts are then unused)
};

__device__ __forceinline__ double det3(double a, double b, double c, double d, double e, double f, double g, double h,
                                       double i)
{
    return a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
}

// rows are (a, c); solves a_i . x + c_i = 0 by Cramer's rule
__device__ __forceinline__ void solve3(const double *r0, const double *r1, const double *r2, double *x)
{
    const double d0 = det3(r0[0], r0[1], r0[2], r1[0], r1[1], r1[2], r2[0], r2[1], r2[2]);
    const double dx = det3(-r0[3], r0[1], r0[2], -r1[3], r1[1], r1[2], -r2[3], r2[1], r2[2]);
    const double dy = det3(r0[0], -r0[3], r0[2], r1[0], -r1[3], r1[2], r2[0], -r2[3], r2[2]);
    const double dz = det3(r0[0], r0[1], -r0[3], r1[0], r1[1], -r1[3], r2[0], r2[1], -r2[3]);
    x[0] = dx / d0;
    x[1] = dy / d0;
    x[2] = dz / d0;
}

constexpr int CLIP_WARPS = 8;
constexpr int CLIP_KEY_WORDS = 128;   // key words kept in shared memory (L <= 4096)
constexpr size_t clip_ring_bytes(int rpl, int depth) { return size_t(CLIP_WARPS) * depth * rpl * 32 * 4 * sizeof(double); }

__device__ __forceinline__ void clip_cp16(void *smem, const void *gmem)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}

// ---- bulk (TMA 1-D) streaming: one instruction of one lane moves a whole 64-row block (2 KiB) and signals an
// mbarrier; the inherited rows leave shared memory the same way (bulk shared -> global).  Replaces 4 predicated
// 16-byte cp.async + 4 16-byte stores per lane and block.
__device__ __forceinline__ void clip_mbar_init(uint32_t bar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar));
}
__device__ __forceinline__ void clip_mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    for (unsigned spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1u << 24)) asm volatile("trap;\n");       // a lost copy must fail loudly, never hang the GPU
    }
}
__device__ __forceinline__ void clip_bulk_load(void *smem_dst, const void *gmem_src, uint32_t bytes, uint32_t bar)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(dst), "l"(gmem_src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void clip_bulk_store(void *gmem_dst, const void *smem_src, uint32_t bytes)
{
    const unsigned src = (unsigned)__cvta_generic_to_shared(smem_src);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gmem_dst), "r"(src), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}

template <int MINB, int RPL, int DEPTH, bool BULK = false>
__global__ void __launch_bounds__(CLIP_WARPS * 32, MINB) clip_kernel(const ClipArgs a)
{
    pdl_enter();
    extern __shared__ __align__(16) double s_ring[];  // [warp][DEPTH][RPL][32 lanes][4]: plane rows in flight
    __shared__ double s_pl[CLIP_WARPS][VSLOTS][4];   // plane of every polygon edge
    __shared__ __align__(16) double s_vx[CLIP_WARPS][VSLOTS + 4][4];   // vertex j = edge j ^ edge j+1 (x, y, z, -);
                                                     // slots k..k+3 repeat vertex 0 (unguarded 4-way loop)
    __shared__ int s_ed[CLIP_WARPS][VSLOTS];
    __shared__ uint32_t s_key[CLIP_WARPS][CLIP_KEY_WORDS];
    __shared__ __align__(8) unsigned long long s_bar[BULK ? CLIP_WARPS : 1][DEPTH];   // one mbarrier per ring slot (BULK)

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int li = blockIdx.x * CLIP_WARPS + wib;
    const int slot = (a.tile_stride <= 1) ? li : ((li / a.tile) * a.tile_stride + a.tile_offset) * a.tile + (li % a.tile);
    if (slot >= a.S) return;
    const int s = a.idx ? a.idx[slot] : slot;
    double(*pl)[4] = s_pl[wib];
    double(*vx)[4] = s_vx[wib];
    int *ed = s_ed[wib];
    const unsigned FULL = 0xFFFFFFFFu;

    const uint32_t *key = a.keys + (size_t)s * a.kw;
    // the key goes to shared memory once (first CLIP_KEY_WORDS words; longer keys read the rest from global)
    const int kwords = (a.L + 31) >> 5;
    uint32_t *skey = s_key[wib];
    for (int w = lane; w < CLIP_KEY_WORDS; w += 32) skey[w] = (w < kwords) ? key[w] : 0u;
    __syncwarp();
    double eq[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) eq[j] = a.equ[(size_t)s * 4 + j];
    const double nn = eq[0] * eq[0] + eq[1] * eq[1] + eq[2] * eq[2];
    const bool dead = !(nn > 0.0) || !isfinite(nn) || !isfinite(eq[3]);
    const double sp[3] = {a.seedpt[(size_t)s * 4 + 0], a.seedpt[(size_t)s * 4 + 1], a.seedpt[(size_t)s * 4 + 2]};
    const double hint = a.seedpt[(size_t)s * 4 + 3];
    const double huge = 1e4 * fmax(1.0, fmax(fabs(sp[0]), fmax(fabs(sp[1]), fabs(sp[2]))));
    // The starting square is a few times the parent polygon's extent (children resemble their parents),
    // so only the planes near the face ever cut it.  If an artificial edge survives, the square was too
    // small: retry 64x larger, at the latest with the huge default on the third attempt.
    double big = (hint > 0.0 && isfinite(hint)) ? fmin(8.0 * hint, huge) : huge;
    int k = 0;
    int n_inconsistent = 0;
    bool overflow = false;
    const uint32_t bar0 = BULK ? (uint32_t)__cvta_generic_to_shared(&s_bar[BULK ? wib : 0][0]) : 0u;
    uint32_t bar_phase = 0;                          // bit d: parity the next wait on ring slot d must see
    if (BULK) {
        if (lane == 0) {
            for (int d = 0; d < DEPTH; ++d) clip_mbar_init(bar0 + 8u * d);
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        }
        __syncwarp();
    }
    for (int attempt = 0;; ++attempt) {
    k = 0;
    overflow = false;
    if (!dead) {
        // centre of the bounding square: seed point projected onto the level plane
        const double t = (eq[0] * sp[0] + eq[1] * sp[1] + eq[2] * sp[2] + eq[3]) / nn;
        const double x0[3] = {sp[0] - t * eq[0], sp[1] - t * eq[1], sp[2] - t * eq[2]};
        // in-plane orthonormal basis (u, v) with u x v along +n
        const double inv = rsqrt(nn);
        const double n[3] = {eq[0] * inv, eq[1] * inv, eq[2] * inv};
        int ax = 0;
        if (fabs(n[1]) < fabs(n[ax])) ax = 1;
        if (fabs(n[2]) < fabs(n[ax])) ax = 2;
        const double e[3] = {ax == 0 ? 1.0 : 0.0, ax == 1 ? 1.0 : 0.0, ax == 2 ? 1.0 : 0.0};
        double u[3] = {n[1] * e[2] - n[2] * e[1], n[2] * e[0] - n[0] * e[2], n[0] * e[1] - n[1] * e[0]};
        const double ul = rsqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        u[0] *= ul; u[1] *= ul; u[2] *= ul;
        const double v[3] = {n[1] * u[2] - n[2] * u[1], n[2] * u[0] - n[0] * u[2], n[0] * u[1] - n[1] * u[0]};
        if (lane < 4) {
            // edges in counter-clockwise order: +u, +v, -u, -v ; vertex j = edge j ^ edge j+1
            const double su = (lane == 0) ? 1.0 : (lane == 2) ? -1.0 : 0.0;
            const double sv = (lane == 1) ? 1.0 : (lane == 3) ? -1.0 : 0.0;
            const double d[3] = {su * u[0] + sv * v[0], su * u[1] + sv * v[1], su * u[2] + sv * v[2]};
            pl[lane][0] = d[0]; pl[lane][1] = d[1]; pl[lane][2] = d[2];
            pl[lane][3] = -(d[0] * x0[0] + d[1] * x0[1] + d[2] * x0[2]) - big;
            ed[lane] = -1 - lane;
            const double cu = (lane == 0 || lane == 3) ? big : -big;   // corners: (+,+) (-,+) (-,-) (+,-)
            const double cv = (lane == 0 || lane == 1) ? big : -big;
            vx[lane][0] = x0[0] + cu * u[0] + cv * v[0];
            vx[lane][1] = x0[1] + cu * u[1] + cv * v[1];
            vx[lane][2] = x0[2] + cu * u[2] + cv * v[2];
        }
        k = 4;
    }
    __syncwarp();
    // Bounding sphere of the current polygon (centre = vertex mean, radius inflated by 1e-6 + 1e-12): a plane whose
    // signed value at the centre is <= -|a| R has no vertex on its positive side, which rejects almost every row
    // with 8 FP64 operations instead of 4 per vertex.  Conservative: a row is only ever rejected when the exact
    // predicate below would reject it as well, so the result does not depend on the filter.
    double bcx = 0.0, bcy = 0.0, bcz = 0.0, bR2 = 0.0;
    auto update_bound = [&]() {
        double sx = 0.0, sy = 0.0, sz = 0.0;
        for (int j = 0; j < k; ++j) { sx += vx[j][0]; sy += vx[j][1]; sz += vx[j][2]; }
        const double inv_k = (k > 0) ? 1.0 / (double)k : 0.0;
        bcx = sx * inv_k; bcy = sy * inv_k; bcz = sz * inv_k;
        double r2 = 0.0;
        for (int j = 0; j < k; ++j) {
            const double dx = vx[j][0] - bcx, dy = vx[j][1] - bcy, dz = vx[j][2] - bcz;
            r2 = fmax(r2, dx * dx + dy * dy + dz * dz);
        }
        const double r = sqrt(r2) * (1.0 + 1e-6) + 1e-12 * (1.0 + fabs(bcx) + fabs(bcy) + fabs(bcz));
        bR2 = r * r;
    };
    update_bound();

    // per-warp ring of DEPTH blocks of RPL * 32 rows: the next DEPTH-1 blocks are in flight without holding registers
    double *ring = s_ring + (size_t)wib * DEPTH * (RPL * 32) * 4;

    // applies the cutting planes of one 32-row half block, one at a time in row order (p, rs: this lane's row)
    auto apply_cuts = [&](unsigned long long todo, const double (&p)[RPL][4], const double (&rs)[RPL], int ebase) {
        while (todo) {
            const int idx = __ffsll((long long)todo) - 1;          // bit 32 h + lane: row h of that lane, in row order
            todo &= todo - 1;
            const int src = idx & 31;
            const bool second = RPL > 1 && idx >= 32;
            double q[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) q[j] = __shfl_sync(FULL, second ? p[RPL - 1][j] : p[0][j], src);
            const double qrs = __shfl_sync(FULL, second ? rs[RPL - 1] : rs[0], src);
            bool out = false;
            if (lane < k) {
                const double d = q[0] * vx[lane][0] + q[1] * vx[lane][1] + q[2] * vx[lane][2] + q[3];
                out = (d * qrs > EPS_FEAS);
            }
            const unsigned om = __ballot_sync(FULL, out);
            if (om == 0u) continue;
            const unsigned full = (k == 32) ? FULL : ((1u << k) - 1u);
            if (om == full) { k = 0; break; }                      // region misses the level set
            const unsigned prev = ((om << 1) | (om >> (k - 1))) & full;
            const unsigned starts = om & ~prev;
            const int ra = __ffs(starts) - 1;                      // first outside vertex of the run
            const unsigned rot = ((om >> ra) | (ra ? (om << (k - ra)) : 0u)) & full;
            const int r = __ffs(~rot) - 1;                         // run length
            if (__popc(om) != r) ++n_inconsistent;
            const int rb1 = (ra + r) % k;                          // entry after the run
            const int knew = k - r + 2;
            if (knew > VSLOTS) { overflow = true; break; }
            // gather the new entry of this lane from the old arrays
            double npl[4] = {0, 0, 0, 0}, nvx[3] = {0, 0, 0};
            int ned = 0;
            if (lane < knew) {
                if (lane == 0) {          // (edge ra, vertex = edge ra ^ new plane)
#pragma unroll
                    for (int j = 0; j < 4; ++j) npl[j] = pl[ra][j];
                    ned = ed[ra];
                    solve3(q, npl, eq, nvx);
                } else if (lane == 1) {   // (new plane, vertex = new plane ^ edge rb1)
#pragma unroll
                    for (int j = 0; j < 4; ++j) npl[j] = q[j];
                    ned = ebase + idx;
                    solve3(q, pl[rb1], eq, nvx);
                } else {
                    const int o = (rb1 + lane - 2) % k;
#pragma unroll
                    for (int j = 0; j < 4; ++j) npl[j] = pl[o][j];
#pragma unroll
                    for (int j = 0; j < 3; ++j) nvx[j] = vx[o][j];
                    ned = ed[o];
                }
            }
            __syncwarp();
            if (lane < knew) {
#pragma unroll
                for (int j = 0; j < 4; ++j) pl[lane][j] = npl[j];
#pragma unroll
                for (int j = 0; j < 3; ++j) vx[lane][j] = nvx[j];
                ed[lane] = ned;
            }
            __syncwarp();
            k = knew;
        }
        update_bound();
    };
    // The constraints come in three contiguous row segments (layer-1 rows shared by all states, the
    // state's own rows, the extra constraints); each segment is streamed in 64-row blocks with a running
    // pointer, so the loop body carries no per-row address selection.
    static_assert(RPL <= 2, "apply_cuts selects between at most two rows per lane");
    // rows inherited from the parent (layers 2..bucket): read there, stored into the own buffer on the way
    int inh = 0;
    const double *inh_src = nullptr;
    if (a.P_prev != nullptr) {
        const int bkt = a.bucket[s];
        if (bkt >= 2 && bkt <= a.lo.D) {
            inh = a.lo.off[bkt + 1] - a.n1;
            const int pi = a.parent[a.lb + s] - a.prev_lb;
            inh_src = a.P_prev + (size_t)(a.prev_slot_of ? a.prev_slot_of[pi] : pi) * a.p_stride;
        }
    }
#pragma unroll 1
    for (int seg = 0; seg < 4; ++seg) {
        const int row = a.rows_by_slot ? slot : s;
        const double *own = a.P + (size_t)row * a.p_stride;
        const double *rows0 = (seg == 0) ? a.P1 : (seg == 1) ? inh_src : (seg == 2) ? own + (size_t)inh * 4 : a.extra;
        const int c0 = (seg == 0) ? 0 : (seg == 1) ? a.n1 : (seg == 2) ? a.n1 + inh : a.L;
        const int nrows = (seg == 0) ? a.n1 : (seg == 1) ? inh : (seg == 2) ? a.L - a.n1 - inh : a.E;
        const bool has_bits = seg < 3;
        double *wr = (seg == 1) ? a.P_own + (size_t)row * a.p_stride + (size_t)lane * 2 : nullptr;
        if (nrows <= 0 || k <= 0 || overflow) continue;
        constexpr int BR = RPL * 32;                 // rows per block
        const int nblk = (nrows + BR - 1) / BR;
        // a block is BR * 32 contiguous bytes: copy instruction i moves its 16-byte pieces 32 i .. 32 i + 31
        // (fully coalesced 512 B), whichever rows they belong to; the rows are read back after a warp sync
        const double *src = rows0 + (size_t)lane * 2;
        int n_issued = 0, n_waited = 0;      // BULK: blocks whose copy was issued / whose barrier was consumed
        auto fetch = [&](int b) {            // rows BR b .. BR b + BR - 1 -> ring slot b % DEPTH
            if (BULK) {
                if (b < nblk) {
                    if (lane == 0) {
                        if (wr != nullptr)   // the slot's previous block may still be read by its bulk store
                            asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");
                        const int rows = (nrows - b * BR < BR) ? (nrows - b * BR) : BR;
                        clip_bulk_load(ring + (size_t)(b % DEPTH) * (BR * 4), rows0 + (size_t)b * (BR * 4),
                                       (uint32_t)rows * 32u, bar0 + 8u * (uint32_t)(b % DEPTH));
                    }
                    ++n_issued;
                }
                return;
            }
            double *dst = ring + (size_t)(b % DEPTH) * (BR * 4) + lane * 2;
            const double *r = src + (size_t)b * (BR * 4);
            const int row0 = b * BR + (lane >> 1);
#pragma unroll
            for (int i = 0; i < 2 * RPL; ++i)
                if (row0 + 16 * i < nrows) clip_cp16(dst + i * 64, r + i * 64);
            asm volatile("cp.async.commit_group;\n" ::);
        };
        auto wait_block = [&](int b) {       // block b has landed in its ring slot
            if (BULK) {
                const uint32_t d = (uint32_t)(b % DEPTH);
                clip_mbar_wait(bar0 + 8u * d, (bar_phase >> d) & 1u);
                bar_phase ^= 1u << d;
                ++n_waited;
            } else {
                asm volatile("cp.async.wait_group %0;\n" ::"n"(DEPTH - 2));
            }
            __syncwarp();                    // every lane is past its reads of the slot that is refilled next
        };
#pragma unroll
        for (int b = 0; b < DEPTH - 1; ++b) fetch(b);
        for (int b = 0; b < nblk && k > 0 && !overflow; ++b) {
            wait_block(b);
            if (wr != nullptr) {                     // inherited rows: materialise them in the own buffer
                if (BULK) {
                    if (lane == 0) {
                        const int rows = (nrows - b * BR < BR) ? (nrows - b * BR) : BR;
                        clip_bulk_store(a.P_own + (size_t)row * a.p_stride + (size_t)b * (BR * 4),
                                        ring + (size_t)(b % DEPTH) * (BR * 4), (uint32_t)rows * 32u);
                    }
                } else {
                    const double *slot = ring + (size_t)(b % DEPTH) * (BR * 4) + lane * 2;
                    double *dst = wr + (size_t)b * (BR * 4);
                    const int row0 = b * BR + (lane >> 1);
#pragma unroll
                    for (int i = 0; i < 2 * RPL; ++i)
                        if (row0 + 16 * i < nrows)
                            *reinterpret_cast<double2 *>(dst + i * 64) = *reinterpret_cast<const double2 *>(slot + i * 64);
                }
            }
            const double *mine = ring + (size_t)(b % DEPTH) * (BR * 4) + lane * 4;
            double2 lo[RPL], hi[RPL];
#pragma unroll
            for (int h = 0; h < RPL; ++h) {
                lo[h] = *reinterpret_cast<const double2 *>(mine + h * 128);
                hi[h] = *reinterpret_cast<const double2 *>(mine + h * 128 + 2);
            }
            fetch(b + DEPTH - 1);
            double p[RPL][4];
            double rs[RPL];
            bool cuts[RPL], pos[RPL];
#pragma unroll
            for (int h = 0; h < RPL; ++h) {
                rs[h] = 0.0;
                cuts[h] = pos[h] = false;
                const int r = b * BR + h * 32 + lane;
                const int c = c0 + r;
                // sign 1 - 2 bit applied as an XOR on the IEEE sign bit (exact, and off the FP64 pipe); rows
                // past the end of the segment are zero and never cut
                int flipbit = 0;
                if (has_bits) {
                    const uint32_t w = ((c >> 5) < CLIP_KEY_WORDS) ? skey[c >> 5] : key[c >> 5];
                    flipbit = int((w >> (c & 31)) & 1u) << 31;
                }
                const bool live = r < nrows;
                p[h][0] = live ? __hiloint2double(__double2hiint(lo[h].x) ^ flipbit, __double2loint(lo[h].x)) : 0.0;
                p[h][1] = live ? __hiloint2double(__double2hiint(lo[h].y) ^ flipbit, __double2loint(lo[h].y)) : 0.0;
                p[h][2] = live ? __hiloint2double(__double2hiint(hi[h].x) ^ flipbit, __double2loint(hi[h].x)) : 0.0;
                p[h][3] = live ? __hiloint2double(__double2hiint(hi[h].y) ^ flipbit, __double2loint(hi[h].y)) : 0.0;
            }
            // Fast filter (see update_bound): d_j = t + a . (v_j - c) <= t + |a| R, and d_j * rs > EPS needs d_j > 0
            // (rs >= 0; NaN compares false either way), so t <= 0 and t^2 >= |a|^2 R^2 means the plane cannot cut.
#pragma unroll
            for (int h = 0; h < RPL; ++h) {
                const double t = p[h][0] * bcx + p[h][1] * bcy + p[h][2] * bcz + p[h][3];
                const double aa = p[h][0] * p[h][0] + p[h][1] * p[h][1] + p[h][2] * p[h][2];
                pos[h] = (t > 0.0) || (t * t < aa * bR2);
            }
            bool any_pos = false;
#pragma unroll
            for (int h = 0; h < RPL; ++h) any_pos |= pos[h];
            if (!__any_sync(FULL, any_pos)) continue;                   // the common case: nothing near the polygon
#pragma unroll
            for (int h = 0; h < RPL; ++h) {
                if (pos[h]) {   // rare: evaluate the reference's predicate exactly
                    rs[h] = rsqrt(p[h][0] * p[h][0] + p[h][1] * p[h][1] + p[h][2] * p[h][2]);
                    for (int j = 0; j < k; ++j) {
                        const double d = p[h][0] * vx[j][0] + p[h][1] * vx[j][1] + p[h][2] * vx[j][2] + p[h][3];
                        cuts[h] |= (d * rs[h] > EPS_FEAS);
                    }
                }
            }
            unsigned long long todo = 0ull;
#pragma unroll
            for (int h = 0; h < RPL; ++h) todo |= (unsigned long long)__ballot_sync(FULL, cuts[h]) << (32 * h);
            if (todo) apply_cuts(todo, p, rs, c0 + b * BR);
        }
        if (BULK) {                                 // drain: every issued copy is consumed before the ring is reused
            while (n_waited < n_issued) wait_block(n_waited);
            if (wr != nullptr && lane == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
        } else {
            asm volatile("cp.async.wait_all;\n" ::);   // drain before the next segment reuses the ring
        }
        __syncwarp();
    }

    asm volatile("cp.async.wait_all;\n" ::);   // the loop may leave early with copies still in flight
    // artificial edge left and the square was not yet the huge default -> too small a start: retry
    const bool art_left = (k > 0) && !overflow && __any_sync(FULL, (lane < k) && (ed[lane] < 0));
    if (!art_left || big >= huge) break;
    big = (attempt >= 1) ? huge : fmin(big * 64.0, huge);
    __syncwarp();
    }   // attempts

    // ---- finish: validity, orientation, canonical rotation --------------------------------------
    int unbounded = 0;
    bool finite = true;
    if (k > 0 && !overflow) {
        const bool art = (lane < k) && (ed[lane] < 0);
        const bool bad = (lane < k) && !(isfinite(vx[lane][0]) && isfinite(vx[lane][1]) && isfinite(vx[lane][2]));
        unbounded = __any_sync(FULL, art);
        finite = !__any_sync(FULL, bad);
    }
    const bool keep = (k >= 3) && !overflow && !unbounded && finite;
    if (lane == 0) {
        if (k > 0 && unbounded) atomicAdd(a.counters + CNT_UNBOUNDED, 1ull);
        if (overflow) atomicAdd(a.counters + CNT_OVERFLOW, 1ull);
        if (n_inconsistent) atomicAdd(a.counters + CNT_INCONSISTENT, (unsigned long long)n_inconsistent);
        if (keep && k > VERT_MAX_REF) atomicAdd(a.counters + CNT_OVER_VERTMAX, 1ull);
        if (!a.push_on) a.out_cnt[s] = keep ? k : 0;
    }
    int push_off = 0, kk = keep ? k : 0;
    if (a.push_on) {
        // the owner's push (was a separate pack kernel): compact slot from this rank's cursor, then size + location
        // into EVERY rank's block -- also for a state without polygon, whose entry would otherwise be stale
        if (lane == 0 && kk > 0) {
            push_off = atomicAdd(a.push.cursor, kk);
            if (push_off + kk > a.push.cap_corners) {
                atomicAdd(a.counters + CNT_XCHG_ERROR, 1ull << 32);
                push_off = -1;
            }
        }
        push_off = __shfl_sync(FULL, push_off, 0);
        if (push_off < 0) { kk = 0; push_off = 0; }
        if (lane < a.push.world) {
            unsigned char *b = a.push.base[lane];
            reinterpret_cast<int *>(b + a.push.cnt_base)[s] = kk;
            reinterpret_cast<int2 *>(b + a.push.where_base)[s] = make_int2(a.push.rank, push_off);
        }
        if (kk == 0) return;
    }
    if (!keep) return;
    // g_i = edge carrying v_i -> v_{i+1} = ed[(i+1) % k]; reversed loop: v'_i = v_{k-1-i}, g'_i = ed[k-1-i]
    int g = 0x7FFFFFFF;
    int vsrc = 0;
    if (lane < k) {
        if (!a.flip) { vsrc = lane; g = ed[(lane + 1) % k]; }
        else { vsrc = k - 1 - lane; g = ed[k - 1 - lane]; }
    }
    // argmin of g over the lanes
    int best = g, besti = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int ob = __shfl_xor_sync(FULL, best, o), oi = __shfl_xor_sync(FULL, besti, o);
        if (ob < best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    // output slot of this lane after rotation by besti
    if (lane < k) {
        const int dst = (lane - besti + k) % k;
        if (a.push_on) {
            const double x = vx[vsrc][0], y = vx[vsrc][1], z = vx[vsrc][2];
            for (int q = 0; q < a.push.world; ++q) {
                unsigned char *r = a.push.base[q] + a.push.region_off;
                reinterpret_cast<int *>(r)[push_off + dst] = g;
                double *ov = reinterpret_cast<double *>(r + a.push.xyz_off) + (size_t)(push_off + dst) * 3;
                ov[0] = x; ov[1] = y; ov[2] = z;
            }
        } else {
            a.out_edges[(size_t)s * VSLOTS + dst] = g;
            double *ov = a.out_verts + ((size_t)s * VSLOTS + dst) * 3;
            ov[0] = vx[vsrc][0]; ov[1] = vx[vsrc][1]; ov[2] = vx[vsrc][2];
        }
    }
}

}  // namespace amb
