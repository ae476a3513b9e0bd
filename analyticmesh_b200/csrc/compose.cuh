// compose.cuh -- batched affine composition (the dominant kernel).
//
// Replaces reference backend/inc/process.h:18-193 (fill_constraints): there, per hidden layer, two
// mask kernels + one m=3 strided-batched cuBLAS GEMM + one bias GEMM per batch of <= 1024 states,
// with the activation bits expanded to one double per bit.  Here one launch per layer computes, for
// a whole frontier chunk of S states,
//
//      P[s][m][0..3] = sum_k  W[m][k] * bit(s, k) * Pin[s][k][0..3]   (+ bias[m] on column 3)
//
// as ONE FP64 GEMM  (M = n_out) x (N = 4*S) x (K = n_in)  whose B operand is masked by the packed
// activation key while it is staged into shared memory (cp.async with src-size 0 zero-fills a
// masked row, so inactive neurons cost no global read), and whose weights are shared by all S
// states.  The math is FP64 tensor-core DMMA (mma.sync.m8n8k4.f64): tcgen05 has no FP64 kind, and
// on B200 DFMA and DMMA share the same 37.0 TFLOP/s peak (tools/fp64_peak.cu), but a register-tiled
// DFMA kernel is capped near 64 % of it by register-file bandwidth (three 64-bit operands per FMA;
// first version of this kernel, profiles/r01_compose_gemm_v1_ncu.md), while DMMA reads 4 operand
// pairs per 256 FMAs.  DMMA was measured to be bit-identical to the ascending-k FMA chain
// (tools/dmma_exact.cu: 128000/128000), so every output is still exactly the value the CPU
// restatement used by the parity tests computes.
//
// Tile: 128 (m) x 16 states (= 64 columns) x 16 (k); 256 threads = 8 warps as 4 x 2, warp tile
// 32 x 32 = 4 x 4 DMMA tiles; 4-stage cp.async pipeline, ~100 KiB shared memory, two CTAs per SM.
// The mask words of the tile's states are staged in shared memory once, so the producer never waits
// on a dependent global load.
#pragma once
#include "common.cuh"

namespace amb {

constexpr int GM_BM = 128;
constexpr int GM_AS_STRIDE = GM_BM + 4;   // +32 B per k-row: fragment loads and cp.async writes conflict-free

// tile configuration: WM x WN warps, BK rows per pipeline stage, ST stages, BS states per tile,
// MINB resident CTAs per SM
template <int WM_, int WN_, int BK_, int ST_, int BS_ = 32, int MINB_ = 1>
struct GemmCfg {
    static constexpr int WM = WM_, WN = WN_, BK = BK_, ST = ST_, BS = BS_, MINB = MINB_;
    static constexpr int BN = BS * 4;                   // columns of the tile
    static constexpr int BS_STRIDE = BN + 4;
    static constexpr int THREADS = WM * WN * 32;
    static constexpr int MI = GM_BM / WM / 8;           // DMMA tiles per warp along m
    static constexpr int NJ = BN / WN / 8;              // ... along n
    static constexpr int SMEM_A = BK * GM_AS_STRIDE;    // doubles per stage
    static constexpr int SMEM_B = BK * BS_STRIDE;
    static constexpr int CHUNKS_A = BK * 64 / THREADS;  // 16-byte cp.async per thread per stage
    static constexpr int CHUNKS_B = BK * BS * 2 / THREADS;
    static constexpr size_t TILE_BYTES = size_t(ST) * (SMEM_A + SMEM_B) * sizeof(double);
    // + mask words of the BS states of the tile: BS x nw uint32, nw = ceil((bit0 % 32 + K) / 32) + 1 <= K/32 + 4
    static size_t smem_bytes(int K) { return TILE_BYTES + size_t(BS) * (K / 32 + 4) * sizeof(uint32_t); }
};
// Measured on B200 at 8x512 (tools/run_case.py mlp8x512s, whole march, TFLOP/s of this kernel):
//   128 x 16 states, 2 CTAs/SM (default) 31.2 | 128 x 32 states, 1 CTA/SM 27.7 | 16 warps 27.3 |
//   BK 32 / 3 stages 28.5 | 128 x 8 states 28.5.   Two co-resident CTAs hide each other's prologue,
//   epilogue and barrier phases, which a single CTA per SM exposes.
using GemmDefault = GemmCfg<4, 2, 16, 4, 16, 2>;
using GemmWide = GemmCfg<4, 2, 16, 4, 32, 1>;
constexpr int GM_BK = 16;                 // k padding unit of the staged weights (multiple of every BK/2)
constexpr int GM_KPAD = 32;               // weights are zero padded to a multiple of this many k rows

struct GemmArgs {
    const double *Wt;           // [Kpad][Mpad], k-major, zero padded
    int Mpad, M, K;
    const double *Bsrc;         // rows of the input layer for state 0: [K][4]
    long long b_stride;         // doubles between consecutive states (0: shared table, layer 1)
    const uint32_t *keys;       // key of state 0 of the chunk
    int kw;                     // 32-bit words per key
    int bit0;                   // first activation bit of the input layer
    double *out;                // rows of the output layer for state 0: [M][4]
    long long out_stride;       // doubles between consecutive states
    const double *bias;         // [M] or nullptr
    int S;                      // states (permutation slots) of the launch
    int accumulate;             // out += result (hidden-layer skip with a transform)
    const int *perm;            // optional: tile slot i works on state perm[i] (states sorted by the layer
                                // of their flipped neuron, see classify_kernel); nullptr = identity
    int m_tiles;                // Mpad / GM_BM; blockIdx.x = s_tile * m_tiles + m_tile
    int tile_stride, tile_offset;   // this launch handles state tiles tile_offset, tile_offset + tile_stride, ...
    int rows_by_slot;               // sharded march: a state's rows live at its permutation slot, not at its level index
                                    // (independent chains of launches on several streams fill each other's tails)
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// D(8x8) += A(8x4) * B(4x8), FP64.  Measured on B200 (tools/dmma_exact.cu): bit-identical to the
// ascending-k chain c = fma(a_k, b_k, c), k = 0..3.
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <class C>
__global__ void __launch_bounds__(C::THREADS, C::MINB) compose_gemm_kernel(const GemmArgs a)
{
    extern __shared__ __align__(16) double smem[];
    double *As = smem;                               // [ST][BK][AS_STRIDE]
    double *Bs = smem + C::ST * C::SMEM_A;           // [ST][BK][BS_STRIDE]
    uint32_t *smask = reinterpret_cast<uint32_t *>(smem + C::ST * (C::SMEM_A + C::SMEM_B));

    const int tid = threadIdx.x;
    const int m0 = (blockIdx.x % a.m_tiles) * GM_BM;      // m fastest: CTAs that share a B tile are co-resident
    const int s0 = ((blockIdx.x / a.m_tiles) * a.tile_stride + a.tile_offset) * C::BS;
    const int S = a.S;
    const int KT = (a.K + C::BK - 1) / C::BK;
    const int lane = tid & 31, warp = tid >> 5;

    // ---- mask words of the tile's 32 states -> shared memory (one coalesced pass) ---------------
    // smask[sl][j] = key word (bit0/32 + j) of the state in tile slot sl; nw = words that cover the layer
    const int w0 = a.bit0 >> 5, sh = a.bit0 & 31;
    const int nw = (sh + a.K + 31) / 32 + 1;
    for (int i = tid; i < C::BS * nw; i += C::THREADS) {
        const int sl = i / nw, j = i - sl * nw;
        const int slot = s0 + sl;
        uint32_t v = 0;
        if (slot < S && w0 + j < a.kw) {
            const int st = a.perm ? a.perm[slot] : slot;
            v = a.keys[(size_t)st * a.kw + w0 + j];
        }
        smask[i] = v;
    }

    // ---- producer mapping ------------------------------------------------------------------
    // A: BK rows x 1 KiB = BK*64 chunks of 16 B (rows of 64 chunks, coalesced)
    // B: chunk c -> tile slot c / (2 BK), k row (c % (2 BK)) / 2, (xy | zc) half c % 2: a state's BK rows
    //    are BK * 32 B contiguous in global memory
    long long bbase[C::CHUNKS_B];                    // element offset of the state's rows, -1 = empty slot
#pragma unroll
    for (int i = 0; i < C::CHUNKS_B; ++i) {
        const int slot = s0 + (tid + i * C::THREADS) / (2 * C::BK);
        bbase[i] = (slot < S) ? (long long)((a.perm && !a.rows_by_slot) ? a.perm[slot] : slot) * a.b_stride : -1;
    }
    __syncthreads();                                 // smask visible

    auto load_stage = [&](int kt, int slot) {
        double *as = As + slot * C::SMEM_A;
        double *bs = Bs + slot * C::SMEM_B;
        const int k0 = kt * C::BK;
#pragma unroll
        for (int i = 0; i < C::CHUNKS_A; ++i) {
            const int c = tid + i * C::THREADS;
            const int row = c >> 6, col = (c & 63) * 2;
            cp_async16(as + row * GM_AS_STRIDE + col, a.Wt + (size_t)(k0 + row) * a.Mpad + m0 + col, 16);
        }
        const int pos = sh + k0;                     // bit position of the slab inside the staged words
#pragma unroll
        for (int i = 0; i < C::CHUNKS_B; ++i) {
            const int c = tid + i * C::THREADS;
            const int sl = c / (2 * C::BK), r = c % (2 * C::BK);
            const int bk = r >> 1, bhalf = r & 1;
            const int k = k0 + bk;
            const uint32_t *mw = smask + sl * nw + (pos >> 5);
            const uint32_t bits = __funnelshift_r(mw[0], mw[1], pos & 31);
            int bytes = 0;
            const double *src = a.Bsrc;
            if (bbase[i] >= 0 && k < a.K && ((bits >> bk) & 1u)) {
                bytes = 16;
                src = a.Bsrc + bbase[i] + (size_t)k * 4 + bhalf * 2;
            }
            cp_async16(bs + bk * C::BS_STRIDE + sl * 4 + bhalf * 2, src, bytes);
        }
    };

    // ---- consumer mapping: warp (wm, wn) owns rows [8 MI wm, +8 MI) x columns [8 NJ wn, +8 NJ) -------
    // DMMA fragments (PTX m8n8k4.f64): a = A[row = lane/4][k = lane%4], b = B[k = lane%4][col = lane/4],
    //                                   c0,c1 = C[row = lane/4][col = 2 (lane%4) + {0,1}]
    const int wm = warp % C::WM, wn = warp / C::WM;
    const int fr = lane >> 2, fk = lane & 3;
    double acc[C::MI][C::NJ][2];
#pragma unroll
    for (int i = 0; i < C::MI; ++i)
#pragma unroll
        for (int j = 0; j < C::NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int st = 0; st < C::ST - 1; ++st) {
        if (st < KT) load_stage(st, st);
        cp_async_commit();
    }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<C::ST - 2>();
        __syncthreads();
        if (kt + C::ST - 1 < KT) load_stage(kt + C::ST - 1, (kt + C::ST - 1) % C::ST);
        cp_async_commit();

        const double *as = As + (kt % C::ST) * C::SMEM_A + fk * GM_AS_STRIDE + wm * (8 * C::MI) + fr;
        const double *bs = Bs + (kt % C::ST) * C::SMEM_B + fk * C::BS_STRIDE + wn * (8 * C::NJ) + fr;
#pragma unroll
        for (int kk = 0; kk < C::BK / 4; ++kk) {
            double av[C::MI], bv[C::NJ];
#pragma unroll
            for (int i = 0; i < C::MI; ++i) av[i] = as[kk * 4 * GM_AS_STRIDE + i * 8];
#pragma unroll
            for (int j = 0; j < C::NJ; ++j) bv[j] = bs[kk * 4 * C::BS_STRIDE + j * 8];
#pragma unroll
            for (int i = 0; i < C::MI; ++i)
#pragma unroll
                for (int j = 0; j < C::NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], av[i], bv[j]);
        }
    }
    cp_async_wait<0>();

    // ---- epilogue: + bias on column 3, store 16 B per (row, state) ---------------------------
    // columns 2 fk, 2 fk + 1 of an 8-column block = components comp0, comp0+1 of state (block*2 + fk/2)
    const int comp0 = (fk & 1) * 2;
#pragma unroll
    for (int j = 0; j < C::NJ; ++j) {
        const int slot = s0 + wn * (2 * C::NJ) + j * 2 + (fk >> 1);
        if (slot >= S) continue;
        const int s = (a.perm && !a.rows_by_slot) ? a.perm[slot] : slot;
        double *dst = a.out + (size_t)s * a.out_stride + comp0;
#pragma unroll
        for (int i = 0; i < C::MI; ++i) {
            const int m = m0 + wm * (8 * C::MI) + i * 8 + fr;
            if (m >= a.M) continue;
            double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
            if (a.bias != nullptr && comp0 == 2) v.y += a.bias[m];
            double2 *p = reinterpret_cast<double2 *>(dst + (size_t)m * 4);
            if (a.accumulate) {
                const double2 o = *p;
                v.x = o.x + v.x;
                v.y = o.y + v.y;
            }
            *p = v;
        }
    }
}

// ---- skip connections that are not GEMMs -----------------------------------------------------

// from the raw input: P[s][m][0..2] += T[m][0..2]   (T == nullptr: += I3)   process.h:86-92,109-115
__global__ void skip_input_kernel(double *out, long long out_stride, int M, int S, const double *T, const int *perm,
                                  int tile_states, int tile_stride, int tile_offset, int rows_by_slot)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)S * M) return;
    const int slot = int(t / M), m = int(t % M);
    if ((slot / tile_states) % tile_stride != tile_offset) return;    // another chain's tile
    const int s = (perm && !rows_by_slot) ? perm[slot] : slot;
    double *p = out + (size_t)s * out_stride + (size_t)m * 4;
    if (T != nullptr) {
        p[0] += T[3 * m + 0];
        p[1] += T[3 * m + 1];
        p[2] += T[3 * m + 2];
    } else if (m < 3) {
        p[m] += 1.0;
    }
}

// identity skip from hidden layer `src`: out[s][m][:] += bit(s, src_bit0+m) * in[s][m][:]   process.h:93-105
__global__ void skip_hidden_identity_kernel(double *out, long long out_stride, const double *in, long long in_stride,
                                            const uint32_t *keys, int kw, int src_bit0, int M, int S, const int *perm,
                                            int tile_states, int tile_stride, int tile_offset, int rows_by_slot)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)S * M * 4) return;
    const int c = int(t & 3);
    const long long r = t >> 2;
    const int slot = int(r / M), m = int(r % M);
    if ((slot / tile_states) % tile_stride != tile_offset) return;
    const int s = perm ? perm[slot] : slot;
    const int row = rows_by_slot ? slot : s;
    const int bit = src_bit0 + m;
    if ((keys[(size_t)s * kw + (bit >> 5)] >> (bit & 31)) & 1u)
        out[(size_t)row * out_stride + (size_t)m * 4 + c] += in[(size_t)row * in_stride + (size_t)m * 4 + c];
}

// ---- incremental composition ----------------------------------------------------------------------
// A child differs from its parent in ONE bit, in hidden layer l_e.  The rows of hidden layers <= l_e
// depend only on bits of layers < l_e, so they are bit-identical to the parent's rows: they are
// copied from the previous level's plane buffer (kept resident) and only layers > l_e are recomputed.
// States are bucketed by b = l_e (1 for seeds and for children whose parent rows are gone); the
// launch of fc layer h works on the prefix of the bucket-sorted permutation with b <= h.

// Sharded mode: a state owned by another rank goes to bucket D+1, which no launch touches.
__global__ void classify_kernel(const int *via_edge, const int *parent, int lb, int S, int prev_lb, int prev_S,
                                LayerOffs lo, int *bucket, int *counts /* [D+2], zeroed */, const uint8_t *owner,
                                int rank)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const int e = via_edge[lb + s], p = parent[lb + s];
    int b = 1;
    if (owner != nullptr && owner[lb + s] != rank) {
        b = lo.D + 1;
    } else if (e >= 0 && p >= prev_lb && p < prev_lb + prev_S) {
        while (b < lo.D && e >= lo.off[b + 1]) ++b;
    }
    bucket[s] = b;
    atomicAdd(counts + b, 1);
}

// counts[b] -> cursor[b] = start of bucket b ; n_prefix[h] = #states with bucket <= h   (b = 0..D+1)
__global__ void bucket_offsets_kernel(int *counts, int *cursor, int *n_prefix, int D)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int run = 0;
    for (int b = 0; b <= D + 1; ++b) {
        cursor[b] = run;
        run += counts[b];
        n_prefix[b] = run;
        counts[b] = 0;
    }
}

__global__ void scatter_kernel(const int *bucket, int S, int *cursor, int *perm)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    perm[atomicAdd(cursor + bucket[s], 1)] = s;
}

// classify + scatter in one launch: the host already knows the bucket sizes of the level (histogram produced by
// the previous level's winner kernel), so the start of every bucket is a launch parameter and a state goes
// straight to  perm[base[b] + cursor[b]++].  cursor[] must be zero at launch (cleared by the winner kernel).
struct BucketBase { int base[MAX_LAYERS + 2]; };
__global__ void classify_scatter_kernel(const int *via_edge, const int *parent, int lb, int S, int prev_lb, int prev_S,
                                        LayerOffs lo, BucketBase bb, int *bucket, int *cursor, int *perm,
                                        const uint8_t *owner, int rank, int *slot_of)
{
    pdl_enter();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const int e = via_edge[lb + s], p = parent[lb + s];
    int b = 1;
    if (owner != nullptr && owner[lb + s] != rank) {
        b = lo.D + 1;
    } else if (e >= 0 && p >= prev_lb && p < prev_lb + prev_S) {
        while (b < lo.D && e >= lo.off[b + 1]) ++b;
    }
    bucket[s] = b;
    const int pos = bb.base[b] + atomicAdd(cursor + b, 1);
    perm[pos] = s;
    if (slot_of != nullptr) slot_of[s] = pos;       // inverse permutation: where the state's rows live (sharded march)
}

// rows of hidden layers 2..b of the parent -> own rows; one warp per state
__global__ void copy_parent_rows_kernel(const int *bucket, const int *parent, int lb, int S, int prev_lb,
                                        const double *prev, double *cur, long long stride, LayerOffs lo, int n1,
                                        const int *slot_of, const int *prev_slot_of)
{
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (s >= S) return;
    const int b = bucket[s];
    if (b < 2 || b > lo.D) return;
    const long long n16 = (long long)(lo.off[b + 1] - n1) * 2;      // 16-byte pieces to copy
    const int pi = parent[lb + s] - prev_lb;
    const uint4 *src = reinterpret_cast<const uint4 *>(prev + (size_t)(prev_slot_of ? prev_slot_of[pi] : pi) * stride);
    uint4 *dst = reinterpret_cast<uint4 *>(cur + (size_t)(slot_of ? slot_of[s] : s) * stride);
    long long i = lane;
    for (; i + 96 < n16; i += 128) {                              // 4 independent 16 B loads in flight per lane
        const uint4 v0 = src[i], v1 = src[i + 32], v2 = src[i + 64], v3 = src[i + 96];
        dst[i] = v0; dst[i + 32] = v1; dst[i + 64] = v2; dst[i + 96] = v3;
    }
    for (; i < n16; i += 32) dst[i] = src[i];
}

// ---- output layer: the level plane (w_equ, b_equ - iso) ----------------------------------------
// One thread per (state, component): a single ascending-k FMA chain, same order as the oracle.
struct EquSkip {
    int kind;                 // 0 none, 1 input identity, 2 input linear, 3 hidden identity, 4 hidden linear
    const double *T;          // kind 2: [1][3]; kind 4: [1][n_src]
    const double *src;        // kind 3/4: rows of the source layer for state 0
    long long src_stride;
    int src_bit0, src_n;
};
constexpr int EQU_MAX_SKIPS = 4;
struct EquArgs {
    const double *w;          // [K] last fc layer
    double bias, iso;
    const double *in;         // rows of the last hidden layer for state 0
    long long in_stride;
    const uint32_t *keys;
    int kw, bit0, K, S;
    double *equ;              // [S][4]
    const int *idx;           // optional list of the states to compute (sharded mode); nullptr = all
    int n_skips;
    EquSkip skips[EQU_MAX_SKIPS];
    // read-through (see SliceArgs): states whose flipped neuron lies in the last hidden layer read its rows
    // from the parent's level buffer
    const int *bucket;        // per state of the level, nullptr = nobody
    const int *parent;
    int lb, prev_lb, D;
    const double *alt_in;     // rows of the last hidden layer of the previous level's state 0
    // chained launches: this launch handles the list positions of tiles tile_offset, tile_offset + tile_stride, ...
    int tile_stride, tile_offset, tile;
    // sharded march: own rows live at the list position (= permutation slot), the parent's at prev_slot_of[parent]
    int rows_by_slot;
    const int *prev_slot_of;
};

// i-th position handled by a chained launch -> position in the whole list (tiles of `tile` positions dealt round-robin)
__device__ __forceinline__ int chain_position(int i, int tile, int stride, int offset)
{
    return (stride <= 1) ? i : ((i / tile) * stride + offset) * tile + (i % tile);
}

__device__ __forceinline__ double masked_chain(const double *w, const double *rows, const uint32_t *key, int bit0,
                                               int K, int c)
{
    // same ascending-k chain over the active rows; the loads of four rows are issued unconditionally and ahead of
    // the chain (memory-level parallelism), only the FMAs are predicated
    double acc = 0.0;
    int k = 0;
    for (; k + 4 <= K; k += 4) {
        const int bit = bit0 + k;
        const uint32_t lo = key[bit >> 5], hi = ((bit & 31) > 28) ? key[(bit >> 5) + 1] : 0u;
        const uint32_t m = __funnelshift_r(lo, hi, bit & 31);
        const double r0 = rows[(size_t)k * 4 + c], r1 = rows[(size_t)k * 4 + 4 + c];
        const double r2 = rows[(size_t)k * 4 + 8 + c], r3 = rows[(size_t)k * 4 + 12 + c];
        const double w0 = w[k], w1 = w[k + 1], w2 = w[k + 2], w3 = w[k + 3];
        if (m & 1u) acc = fma(w0, r0, acc);
        if (m & 2u) acc = fma(w1, r1, acc);
        if (m & 4u) acc = fma(w2, r2, acc);
        if (m & 8u) acc = fma(w3, r3, acc);
    }
    for (; k < K; ++k) {
        const int bit = bit0 + k;
        if ((key[bit >> 5] >> (bit & 31)) & 1u) acc = fma(w[k], rows[(size_t)k * 4 + c], acc);
    }
    return acc;
}

__global__ void equ_kernel(const EquArgs a)
{
    pdl_enter();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int pos = chain_position(t >> 2, a.tile, a.tile_stride, a.tile_offset);
    if (pos >= a.S) return;
    const int s = a.idx ? a.idx[pos] : pos, c = t & 3;
    const uint32_t *key = a.keys + (size_t)s * a.kw;
    const int row = a.rows_by_slot ? pos : s;
    const double *rows = a.in + (size_t)row * a.in_stride;
    if (a.bucket != nullptr && a.bucket[s] == a.D) {
        const int pi = a.parent[a.lb + s] - a.prev_lb;
        rows = a.alt_in + (size_t)(a.prev_slot_of ? a.prev_slot_of[pi] : pi) * a.in_stride;
    }
    double v = masked_chain(a.w, rows, key, a.bit0, a.K, c);
    if (c == 3) v += a.bias;
    for (int i = 0; i < a.n_skips; ++i) {
        const EquSkip &sk = a.skips[i];
        if (sk.kind == 1) {
            if (c == 0) v += 1.0;
        } else if (sk.kind == 2) {
            if (c < 3) v += sk.T[c];
        } else if (sk.kind == 3) {
            if ((key[sk.src_bit0 >> 5] >> (sk.src_bit0 & 31)) & 1u) v += sk.src[(size_t)row * sk.src_stride + c];
        } else if (sk.kind == 4) {
            v += masked_chain(sk.T, sk.src + (size_t)row * sk.src_stride, key, sk.src_bit0, sk.src_n, c);
        }
    }
    if (c == 3) v -= a.iso;
    a.equ[(size_t)s * 4 + c] = v;
}

// Warp-cooperative level plane (tcgen05 path, no skip into the output layer): one warp per state, lane l owns the
// neurons k = l, l + 32, ... (coalesced 1 KiB row reads instead of 512 dependent FMAs per thread), partial sums are
// combined by a fixed butterfly, so the result is deterministic (it differs from the sequential chain of equ_kernel
// in the last bits only; the DMMA path keeps equ_kernel, whose chain is the oracle's bit for bit).
__global__ void __launch_bounds__(256) equ_warp_kernel(const EquArgs a)
{
    pdl_enter();
    const int lane = threadIdx.x & 31;
    const int li = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int pos = chain_position(li, a.tile, a.tile_stride, a.tile_offset);
    if (pos >= a.S) return;
    const int s = a.idx ? a.idx[pos] : pos;
    const uint32_t *key = a.keys + (size_t)s * a.kw;
    const double *rows = a.in + (size_t)(a.rows_by_slot ? pos : s) * a.in_stride;
    if (a.bucket != nullptr && a.bucket[s] == a.D) {
        const int pi = a.parent[a.lb + s] - a.prev_lb;
        rows = a.alt_in + (size_t)(a.prev_slot_of ? a.prev_slot_of[pi] : pi) * a.in_stride;
    }
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k = lane; k < a.K; k += 32) {
        const int bit = a.bit0 + k;
        if ((key[bit >> 5] >> (bit & 31)) & 1u) {
            const double w = a.w[k];
            const double2 p = *reinterpret_cast<const double2 *>(rows + (size_t)k * 4);
            const double2 q = *reinterpret_cast<const double2 *>(rows + (size_t)k * 4 + 2);
            acc[0] = fma(w, p.x, acc[0]); acc[1] = fma(w, p.y, acc[1]);
            acc[2] = fma(w, q.x, acc[2]); acc[3] = fma(w, q.y, acc[3]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[c] += __shfl_xor_sync(0xFFFFFFFFu, acc[c], o);
    if (lane < 4) {
        double v = (lane == 0) ? acc[0] : (lane == 1) ? acc[1] : (lane == 2) ? acc[2] : acc[3];
        if (lane == 3) v = v + a.bias - a.iso;
        a.equ[(size_t)s * 4 + lane] = v;
    }
}

}  // namespace amb
