// clip.cuh -- warp-cooperative extraction of the face polygon of one linear region.
//
// Replaces the reference's ordering + start search + vertex pivoting + orientation stages
// (backend/inc/process.h:208-321, inc/kernel.h:799-1385): there a host loop launches ~10 kernels
// and does 4 blocking scalar copies per pivot step; here one warp owns one state and clips the
// zero-level plane of the region's affine function against all C half-spaces in ONE pass over the
// plane rows (each row is read exactly once, 32 rows per warp iteration, coalesced 1 KiB loads):
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
// The polygon starts as a large square (artificial edges, ids < 0) in the level plane centred on
// the state's seed point; a final polygon that still has an artificial edge is unbounded and is
// dropped (counted).  The loop is kept counter-clockwise around +w_equ, which is the reference's
// output orientation (inc/kernel.h:1349-1353); flip_insideout reverses it.
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
    int *out_cnt;             // [S]
    int *out_edges;           // [S][VSLOTS]
    double *out_verts;        // [S][VSLOTS][3]
    unsigned long long *counters;
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
constexpr int CLIP_DEPTH = 4;     // cp.async ring: rows of the next CLIP_DEPTH-1 iterations are in flight per lane
constexpr size_t CLIP_RING_BYTES = size_t(CLIP_WARPS) * CLIP_DEPTH * 32 * 4 * sizeof(double);

__device__ __forceinline__ void clip_cp16(void *smem, const void *gmem)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}

__global__ void __launch_bounds__(CLIP_WARPS * 32, 3) clip_kernel(const ClipArgs a)
{
    extern __shared__ __align__(16) double s_ring[];  // [warp][CLIP_DEPTH][32 lanes][4]: plane rows in flight
    __shared__ double s_pl[CLIP_WARPS][VSLOTS][4];   // plane of every polygon edge
    __shared__ __align__(16) double s_vx[CLIP_WARPS][VSLOTS + 4][4];   // vertex j = edge j ^ edge j+1 (x, y, z, -);
                                                     // slots k..k+3 repeat vertex 0 (unguarded 4-way loop)
    __shared__ int s_ed[CLIP_WARPS][VSLOTS];

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int slot = blockIdx.x * CLIP_WARPS + wib;
    if (slot >= a.S) return;
    const int s = a.idx ? a.idx[slot] : slot;
    double(*pl)[4] = s_pl[wib];
    double(*vx)[4] = s_vx[wib];
    int *ed = s_ed[wib];
    const unsigned FULL = 0xFFFFFFFFu;

    const uint32_t *key = a.keys + (size_t)s * a.kw;
    // the key lives in registers: lane l holds words l, l+32, l+64, l+96 (L <= 4096); the word of a
    // 32-row block is fetched with one shuffle instead of a dependent global load per iteration
    const int kwords = (a.L + 31) >> 5;
    uint32_t kreg0 = (lane < kwords) ? key[lane] : 0u, kreg1 = (lane + 32 < kwords) ? key[lane + 32] : 0u;
    uint32_t kreg2 = (lane + 64 < kwords) ? key[lane + 64] : 0u, kreg3 = (lane + 96 < kwords) ? key[lane + 96] : 0u;
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
    const int C = a.L + a.E;
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
    if (lane < 4) {   // padding slots
        vx[k + lane][0] = vx[0][0]; vx[k + lane][1] = vx[0][1]; vx[k + lane][2] = vx[0][2];
    }
    __syncwarp();

    // row of constraint c for this state (layer 1 rows are shared, extra constraints follow the neurons)
    auto row_of = [&](int c) -> const double * {
        return (c < a.n1) ? a.P1 + (size_t)c * 4
             : (c < a.L)  ? a.P + (size_t)s * a.p_stride + (size_t)(c - a.n1) * 4
                          : a.extra + (size_t)(c - a.L) * 4;
    };
    // Every lane streams its own row of each 32-row block through a private shared-memory ring with
    // cp.async, so CLIP_DEPTH-1 loads per lane are in flight without holding registers.
    double *ring = s_ring + ((size_t)wib * CLIP_DEPTH * 32 + lane) * 4;
    auto fetch = [&](int blk) {          // rows of block `blk` -> ring slot blk % CLIP_DEPTH
        const int c = blk * 32 + lane;
        if (c < C) {
            const double *r = row_of(c);
            double *dst = ring + (size_t)(blk % CLIP_DEPTH) * 32 * 4;
            clip_cp16(dst, r);
            clip_cp16(dst + 2, r + 2);
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
#pragma unroll
    for (int b = 0; b < CLIP_DEPTH - 1; ++b) fetch(b);
    for (int base = 0; base < C && k > 0 && !overflow; base += 32) {
        const int c = base + lane;
        const int blk = base >> 5;
        asm volatile("cp.async.wait_group %0;\n" ::"n"(CLIP_DEPTH - 2));
        const double *mine = ring + (size_t)(blk % CLIP_DEPTH) * 32 * 4;
        const double2 lo = *reinterpret_cast<const double2 *>(mine);
        const double2 hi = *reinterpret_cast<const double2 *>(mine + 2);
        fetch(blk + CLIP_DEPTH - 1);
        uint32_t kword;   // activation bits of rows base .. base+31 (warp-uniform source lane)
        if (blk < 128) {
            const uint32_t sel = (blk < 32) ? kreg0 : (blk < 64) ? kreg1 : (blk < 96) ? kreg2 : kreg3;
            kword = __shfl_sync(FULL, sel, blk & 31);
        } else {
            kword = (blk < kwords) ? key[blk] : 0u;
        }
        double p[4] = {0, 0, 0, 0};
        double rs = 0.0;
        bool cuts = false;
        if (c < C) {
            // sign 1 - 2 bit applied as an XOR on the IEEE sign bit (exact, and off the FP64 pipe)
            const int flipbit = (c < a.L) ? int((kword >> lane) & 1u) << 31 : 0;
            p[0] = __hiloint2double(__double2hiint(lo.x) ^ flipbit, __double2loint(lo.x));
            p[1] = __hiloint2double(__double2hiint(lo.y) ^ flipbit, __double2loint(lo.y));
            p[2] = __hiloint2double(__double2hiint(hi.x) ^ flipbit, __double2loint(hi.x));
            p[3] = __hiloint2double(__double2hiint(hi.y) ^ flipbit, __double2loint(hi.y));
            // Fast filter: d_j * rs > EPS needs d_j > 0 (rs >= 0; NaN compares false either way), so a
            // plane with no vertex on its positive side cannot cut.  Slots k..k+3 repeat vertex 0, so the
            // loop runs unguarded in steps of four.
            bool pos = false;
            for (int j = 0; j < k; j += 4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const double2 xy = *reinterpret_cast<const double2 *>(&vx[j + u][0]);
                    const double z = vx[j + u][2];
                    const double d = p[0] * xy.x + p[1] * xy.y + p[2] * z + p[3];
                    pos |= (d > 0.0);
                }
            }
            if (pos) {   // rare: evaluate the reference's predicate exactly
                rs = rsqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
                for (int j = 0; j < k; ++j) {
                    const double d = p[0] * vx[j][0] + p[1] * vx[j][1] + p[2] * vx[j][2] + p[3];
                    cuts |= (d * rs > EPS_FEAS);
                }
            }
        }
        unsigned todo = __ballot_sync(FULL, cuts);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            double q[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) q[j] = __shfl_sync(FULL, p[j], src);
            const double qrs = __shfl_sync(FULL, rs, src);
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
                    ned = base + src;
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
            if (lane < 4) {   // refresh the padding slots
                vx[k + lane][0] = vx[0][0]; vx[k + lane][1] = vx[0][1]; vx[k + lane][2] = vx[0][2];
            }
            __syncwarp();
        }
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
        a.out_cnt[s] = keep ? k : 0;
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
        a.out_edges[(size_t)s * VSLOTS + dst] = g;
        double *ov = a.out_verts + ((size_t)s * VSLOTS + dst) * 3;
        ov[0] = vx[vsrc][0]; ov[1] = vx[vsrc][1]; ov[2] = vx[vsrc][2];
    }
}

}  // namespace amb
