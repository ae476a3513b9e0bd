// split.cuh -- the affine-composition GEMM on the 5th-generation tensor cores (tcgen05, kind::i8).
// Bounded error, NOT bit-identical to the reference's FMA chain: tests/test_split_gpu.py checks the rows against
// the integer restatement (bit for bit) AND against the oracle's FP64 rows (within the bound stated below).
//
// tcgen05 has no FP64 kind, so the FP64 contraction  P_out = W . diag(mask) . P_in  (compose.cuh) is
// computed by a fixed-point integer split ("Ozaki scheme") whose only errors are the truncation of the operands
// and the dropped low-order digit products -- every product that IS formed and every accumulation is exact:
// every row of W and every column of the masked P_in is scaled by a power of two and cut into SD signed
// base-256 digits,
//
//        x  =  2^(e-6) * sum_t d_t 2^(-8t) + r,   d_t in [-128, 127],  |r| <= 2^(e - 8 SD + 1),  2^e > max |x|
//
// (e: one exponent per weight row / per plane column).  All digit products are exact in int32:
// |d d'| <= 2^14, K <= 2^13 terms and <= SD products per accumulator stay below 2^31.  The digit planes
// are multiplied pairwise on the tensor cores, products of equal weight 2^(-8 (i+j)) share one TMEM
// accumulator D_g (g = i + j < SD), and the epilogue evaluates
//
//        out[m][n] = 2^(eA_m - 6) 2^(eB_n - 6) * ( ... (D_{SD-1} 2^-8 + D_{SD-2}) 2^-8 + ... + D_0 )  (+ bias)
//
// in FP64 (every step is one correctly rounded operation, so tests/split_emul.py reproduces the result
// bit for bit on the CPU).  With SD = 7 the operands carry 54 bits below their row / column maximum,
// i.e. the rounding of the result is at the level of an FP64 dot product; the dropped products
// (i + j >= SD) are below 2^-50 of max|w| max|p| K.
//
// Kernel layout (persistent: one CTA of 64 + 128 EW threads per SM strides over tiles of 128 output neurons x 16 states):
//   warp 0    TMA producer: per 32-byte K step one 3-D box of the weight digits [SD][128][32] and one of
//             the plane digits [SD][64][32] (SWIZZLE_32B), multi-stage mbarrier ring; runs ahead into the
//             next tile while the epilogue drains TMEM
//   warp 1    TMEM allocation (512 columns, once) + single-thread MMA issue: for weight digit i ONE
//             instruction covers the plane digits j = 0 .. SD-1-i, because their accumulators
//             g = i + j are adjacent 64-column windows of TMEM (N = 64 (SD - i), cut at 256)
//   warps 2.. epilogue, EW warps per TMEM lane quarter (a warp may only read the lanes 32 (warp % 4) ..): tile constants
//             staged in shared memory behind the MMAs, tcgen05.ld of the SD accumulators, exact int -> double, FP64
//             Horner, scale, bias, input-skip addend, 16-byte stores.  One accumulator set fills TMEM (SD x 64 of the
//             512 columns), so the next tile's MMAs wait until the LAST tcgen05.ld of the tile has completed: with
//             EW = 1 that is after three of the four 16-column chunks have been loaded, converted and stored by the one
//             warp of the quarter; with EW warps per quarter every warp drains only 64 / EW columns (in chunks of CW)
//             and the serial part shrinks to (64 / EW - CW) columns of arithmetic.  Every output is computed by the
//             same instruction sequence whatever EW: results are bit-identical (tools/epi_check.py).  MEASURED on the
//             8x512 march (profiles/r02_split_epilogue_warps.md): EW = 2 and EW = 4 are 3 % SLOWER than EW = 1
//             (1403 vs 1364 ms of GEMM time), i.e. the serial epilogue is not what keeps the tensor pipe at 58 %:
//             the K loop itself is bound by the operand fill (43 008 B of TMA traffic per K step and SM against
//             ~46 B/clk of TMA service per SM = 935 clk, the MMAs of a K step take 896-953 clk).  EW = 1 stays the default.
// Plane digits come from slice_rows_reg_kernel (one pass, rows kept in registers) or slice_rows_kernel (any K).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace amb {

constexpr int SP_BM = 128;        // output neurons per tile (TMEM lanes)
constexpr int SP_BS = 16;         // states per tile
constexpr int SP_BN = SP_BS * 4;  // columns per tile
constexpr int SP_BK = 32;         // K bytes per pipeline stage = K of one kind::i8 instruction
constexpr int SP_THREADS = 192;    // EW = 1; in general 64 + 128 EW (split_threads)
__host__ __device__ constexpr int split_threads(int ew) { return 64 + 128 * ew; }
constexpr int SP_KPAD = 32;       // digit rows are padded to a multiple of this many K bytes

template <int SD, int BK = SP_BK>      // BK: K bytes per pipeline stage (32 = one kind::i8 instruction, 64 = two)
struct SplitCfg {
    static constexpr int STAGE_A = SD * SP_BM * BK;
    static constexpr int STAGE_B = SD * SP_BN * BK;
    static constexpr int STAGE = STAGE_A + STAGE_B;
    static constexpr int STAGES = (215 * 1024) / STAGE;
    static constexpr int TMEM_COLS = 512;
    static constexpr int FRAC_BITS = 8 * SD - 2;         // |X| = |x| 2^(FRAC_BITS - e) <= 2^FRAC_BITS
    static constexpr size_t SMEM = size_t(STAGES) * STAGE + 1024 /* alignment slack */ + 2048 /* barriers, tile constants */;
    static_assert(BK == 32 || BK == 64, "stage depth");
    static_assert(STAGES >= 2, "the ring needs two stages");
    static_assert(SD * SP_BN <= TMEM_COLS, "accumulators exceed TMEM");
    static_assert(STAGE % 1024 == 0 && STAGE_A % 1024 == 0, "stage bases must keep the swizzle alignment");
};

// ---- host + device: digits of one value ------------------------------------------------------------
// X = rint(x * 2^(FRAC - e)) as a signed integer; digit t (t = 0 most significant) is byte (SD-1-t)
// of  (X + 0x80..80) ^ 0x80..80  read as int8:  X = sum_t d_t 256^(SD-1-t).
__host__ __device__ inline unsigned long long split_pack(long long X, int SD)
{
    const unsigned long long C = (SD >= 8) ? 0x8080808080808080ull : (0x8080808080808080ull >> (8 * (8 - SD)));
    return ((unsigned long long)X + C) ^ C;
}

// exponent e with 2^(e-1) <= mx < 2^e for a normal, finite mx > 0; 0 otherwise (zero / subnormal / non-finite
// columns produce zero digits) -- frexp() without its slow paths
__device__ __forceinline__ int split_exponent(double mx)
{
    const int be = (__double2hiint(mx) >> 20) & 0x7FF;
    return (be == 0 || be == 0x7FF) ? 0 : be - 1022;
}
// 2^k for -1022 <= k <= 1023
__device__ __forceinline__ double split_pow2(int k) { return __hiloint2double((1023 + k) << 20, 0); }

// ---- plane digits (B operand), one warp per state slot ------------------------------------------------
struct SliceArgs {
    const double *src;          // rows of the input layer for state 0: [K][4]
    long long stride;           // doubles between states (0: the shared layer-1 table)
    const uint32_t *keys;
    int kw, bit0, K, Kpad;      // Kpad: K rounded up to SP_KPAD (the padding digits are written as zeros)
    const int *perm;            // slot -> state (nullptr = identity)
    int S;                      // slots
    signed char *dig;           // [SD][ncap][pitch], column n = slot * 4 + component
    long long pitch;            // bytes between columns (>= Kpad)
    long long slice_stride;     // bytes between digit planes (ncap * pitch)
    double *scale;              // [S * 4]  2^(e - 6)
    int tile_stride, tile_offset;   // this launch handles the slots of tiles tile_offset, tile_offset + tile_stride, ...
    // read-through: slots >= alt_from_slot take the rows of this layer from their PARENT's level buffer (they
    // are identical to the parent's and are not copied first; clip_kernel materialises them later)
    int alt_from_slot;              // S (or more) = no such slots
    const double *alt_src;          // rows of this layer of the previous level's state 0
    const int *parent;              // global state id of the parent, indexed by global state id
    int lb, prev_lb;                // first state id of this / the previous level
    // sharded march: own rows live at the permutation slot, the parent's at prev_slot_of[parent]
    int rows_by_slot;
    const int *prev_slot_of;
};

__device__ __forceinline__ const double *slice_rows_of(const SliceArgs &a, int slot, int s)
{
    if (slot >= a.alt_from_slot) {
        const int pi = a.parent[a.lb + s] - a.prev_lb;
        return a.alt_src + (size_t)(a.prev_slot_of ? a.prev_slot_of[pi] : pi) * a.stride;
    }
    return a.src + (size_t)(a.rows_by_slot ? slot : s) * a.stride;
}

template <int SD>
__global__ void __launch_bounds__(256) slice_rows_kernel(const SliceArgs a)
{
    pdl_enter();
    const int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (slot >= a.S || (slot / SP_BS) % a.tile_stride != a.tile_offset) return;
    const int s = a.perm ? a.perm[slot] : slot;
    const double *rows = slice_rows_of(a, slot, s);
    const uint32_t *key = a.keys + (size_t)s * a.kw;

    auto active = [&](int k) -> bool {
        const int bit = a.bit0 + k;
        return k < a.K && ((key[bit >> 5] >> (bit & 31)) & 1u);
    };
    // pass 1: per component, max |x| over the active rows
    double mx[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k0 = lane * 4; k0 < a.K; k0 += 128) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + j;
            if (active(k)) {
                const double2 p = *reinterpret_cast<const double2 *>(rows + (size_t)k * 4);
                const double2 q = *reinterpret_cast<const double2 *>(rows + (size_t)k * 4 + 2);
                mx[0] = fmax(mx[0], fabs(p.x)); mx[1] = fmax(mx[1], fabs(p.y));
                mx[2] = fmax(mx[2], fabs(q.x)); mx[3] = fmax(mx[3], fabs(q.y));
            }
        }
    }
    double mul[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx[c] = fmax(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
        const int e = split_exponent(mx[c]);                            // mx = f 2^e, f in [0.5, 1)
        mul[c] = split_pow2(max(-1022, min(1023, SplitCfg<SD>::FRAC_BITS - e)));
        if (lane == 0) a.scale[(size_t)slot * 4 + c] = split_pow2(max(-1022, e - 6));
    }
    // pass 2: digits; lane owns 4 consecutive k -> one 32-bit store per (digit, component)
    signed char *col0 = a.dig + (size_t)slot * 4 * a.pitch;
    for (int k0 = lane * 4; k0 < a.Kpad; k0 += 128) {
        unsigned long long Y[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + j;
            double v[4] = {0.0, 0.0, 0.0, 0.0};
            if (active(k)) {
                const double2 p = *reinterpret_cast<const double2 *>(rows + (size_t)k * 4);
                const double2 q = *reinterpret_cast<const double2 *>(rows + (size_t)k * 4 + 2);
                v[0] = p.x; v[1] = p.y; v[2] = q.x; v[3] = q.y;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) Y[j][c] = split_pack(__double2ll_rn(v[c] * mul[c]), SD);
        }
#pragma unroll
        for (int t = 0; t < SD; ++t) {
            const int p = SD - 1 - t;                                     // byte of Y holding digit t
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t w = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) w |= (uint32_t)((Y[j][c] >> (8 * p)) & 0xFFull) << (8 * j);
                *reinterpret_cast<uint32_t *>(col0 + (size_t)t * a.slice_stride + (size_t)c * a.pitch + k0) = w;
            }
        }
    }
}

// Same result, one pass: for K <= 128 NIT the lane keeps its rows in registers between the column
// maximum and the digit extraction (every load is issued up front, nothing is read twice).  Warps are
// persistent and software-pipelined over the slots: the dependent chain slot -> state -> (parent) -> key words
// of the NEXT slot is fetched while the rows of the current one are in flight, so a state costs one memory
// latency instead of four.
template <int NIT>
struct SliceMeta {
    const double *rows;
    uint32_t bits[NIT];         // activation bits of this lane's 4 rows per 128-row iteration (low 4 bits)
};

template <int NIT>
__device__ __forceinline__ SliceMeta<NIT> slice_meta(const SliceArgs &a, int slot, int lane)
{
    SliceMeta<NIT> m;
    const int s = a.perm ? a.perm[slot] : slot;
    m.rows = slice_rows_of(a, slot, s);
    const uint32_t *key = a.keys + (size_t)s * a.kw;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const int k0 = it * 128 + lane * 4;
        m.bits[it] = 0;
        if (k0 < a.K) {
            const int bit = a.bit0 + k0, w = bit >> 5;
            const uint32_t w0 = key[w], w1 = (w + 1 < a.kw) ? key[w + 1] : 0u;
            m.bits[it] = __funnelshift_r(w0, w1, bit & 31);
        }
    }
    return m;
}

template <int SD, int NIT>
__global__ void __launch_bounds__(128, 3) slice_rows_reg_kernel(const SliceArgs a)
{
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int n_warps = gridDim.x * (blockDim.x >> 5);
    int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    // slots of other chains' tiles are skipped (single-chain launches take every slot)
    auto mine = [&](int sl) { return (sl / SP_BS) % a.tile_stride == a.tile_offset; };
    while (slot < a.S && !mine(slot)) slot += n_warps;
    if (slot >= a.S) return;
    SliceMeta<NIT> cur = slice_meta<NIT>(a, slot, lane);
    while (slot < a.S) {
        if (slot + n_warps >= a.S) pdl_trigger();      // last state of this warp: the GEMM's CTAs may move in
        double v[NIT][4][4];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int k0 = it * 128 + lane * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double2 p = make_double2(0.0, 0.0), q = make_double2(0.0, 0.0);
                if (k0 + j < a.K && ((cur.bits[it] >> j) & 1u)) {
                    p = *reinterpret_cast<const double2 *>(cur.rows + (size_t)(k0 + j) * 4);
                    q = *reinterpret_cast<const double2 *>(cur.rows + (size_t)(k0 + j) * 4 + 2);
                }
                v[it][j][0] = p.x; v[it][j][1] = p.y; v[it][j][2] = q.x; v[it][j][3] = q.y;
            }
        }
        int next = slot + n_warps;
        while (next < a.S && !mine(next)) next += n_warps;
        SliceMeta<NIT> nxt = cur;
        if (next < a.S) nxt = slice_meta<NIT>(a, next, lane);        // in flight together with the rows above

        double mul[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            double mx = 0.0;
#pragma unroll
            for (int it = 0; it < NIT; ++it)
#pragma unroll
                for (int j = 0; j < 4; ++j) mx = fmax(mx, fabs(v[it][j][c]));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            const int e = split_exponent(mx);
            mul[c] = split_pow2(max(-1022, min(1023, SplitCfg<SD>::FRAC_BITS - e)));
            if (lane == 0) a.scale[(size_t)slot * 4 + c] = split_pow2(max(-1022, e - 6));
        }
        signed char *col0 = a.dig + (size_t)slot * 4 * a.pitch;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int k0 = it * 128 + lane * 4;
            if (k0 >= a.Kpad) continue;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t lo[4], hi[4];
                signed char *dst = col0 + (size_t)c * a.pitch + k0;          // digit plane 0; planes are slice_stride apart
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const unsigned long long Y = split_pack(__double2ll_rn(v[it][j][c] * mul[c]), SD);
                    lo[j] = (uint32_t)Y;
                    hi[j] = (uint32_t)(Y >> 32);
                }
#pragma unroll
                for (int t = 0; t < SD; ++t) {
                    // digit t = byte p of the four packed values -> one word, three byte permutes
                    const int p = SD - 1 - t, b = p & 3;
                    const uint32_t sel = (uint32_t)(b | ((4 + b) << 4));
                    const uint32_t t01 = __byte_perm(p < 4 ? lo[0] : hi[0], p < 4 ? lo[1] : hi[1], sel);
                    const uint32_t t23 = __byte_perm(p < 4 ? lo[2] : hi[2], p < 4 ? lo[3] : hi[3], sel);
                    *reinterpret_cast<uint32_t *>(dst) = __byte_perm(t01, t23, 0x5410);
                    dst += a.slice_stride;
                }
            }
        }
        cur = nxt;
        slot = next;
    }
}

// ---- PTX wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "SPLIT_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra SPLIT_DONE_%=;\n\t"
        "bra SPLIT_WAIT_%=;\n\t"
        "SPLIT_DONE_%=:\n\t"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, int8 x int8 -> int32
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
// K-major operand, rows of BK bytes, SWIZZLE_32B (BK = 32) / SWIZZLE_64B (BK = 64): 8-row groups are 8 BK bytes apart
// (SBO); version 1 (sm_100).  With BK = 64 the second 32-byte K slab of a row starts 32 B further: the swizzle is a
// function of the shared-memory address bits, so the descriptor's start address is simply advanced (the row-group bases
// stay aligned to the 512-byte swizzle period).
template <int BK = SP_BK>
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t addr)
{
    constexpr uint64_t layout = (BK == 32) ? 6 : 4;      // UMMA LayoutType: SWIZZLE_32B = 6, SWIZZLE_64B = 4
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | (uint64_t(1) << 16) | (uint64_t((8 * BK) >> 4) << 32) | (uint64_t(1) << 46) |
           (layout << 61);
}
// kind::i8 instruction descriptor: D = S32, A = B = signed 8 bit, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t tc_idesc_i8(int n)
{
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(SP_BM >> 4) << 24);
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, int (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, int (&v)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_cols(uint32_t taddr, int (&v)[16]) { tc_ld16(taddr, v); }
__device__ __forceinline__ void tc_ld_cols(uint32_t taddr, int (&v)[8]) { tc_ld8(taddr, v); }
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

struct SplitArgs {
    int k_steps;                // pipeline stages per tile: ceil(Kpad / BK)
    int M, m_tiles, S;          // output neurons, M tiles, state slots of the launch
    int n_tiles;                // tiles of this launch = m_tiles * (state tiles dealt to this chain)
    const int *perm;            // slot -> state
    double *out;                // rows of the output layer for state 0: [M][4]
    long long out_stride;       // doubles between states
    const double *bias;         // [M] or nullptr
    const double *scaleA;       // [Mpad]   2^(eA - 6)
    const double *scaleB;       // [S * 4]  2^(eB - 6)
    int accumulate;             // out += result
    int tile_stride, tile_offset;
    const double *add_in;       // fused skip connection from the raw input: out[m][0..2] += add_in[3 m + 0..2]
    int add_identity;           //   ... or += I3 (identity skip: rows m < 3 get +1 on their own column)
    int rows_by_slot;           // sharded march: the output rows of slot p live at row p (not at perm[p])
};

// exact int32 -> double on the FP64 pipe (LOP3 + DADD; I2F.F64 runs at a quarter of that rate)
__device__ __forceinline__ double i32_to_f64(int x)
{
    return __hiloint2double(0x43300000, (int)((unsigned)x ^ 0x80000000u)) - 4503601774854144.0;   // 2^52 + 2^31
}

// Persistent: gridDim.x CTAs (one per SM) stride over the tiles; the TMA producer runs ahead into the next
// tile while the epilogue drains TMEM, TMEM / barriers are set up once per CTA.
// EW: epilogue warps per TMEM lane quarter (1, 2 or 4); CW: accumulator columns per tcgen05.ld chunk (16 or 8)
template <int SD, int EW = 1, int CW = 16, int BK = SP_BK>
__global__ void __launch_bounds__(split_threads(EW), 1)
split_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const SplitArgs a)
{
    using C = SplitCfg<SD, BK>;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw0 = smem_u32(smem_raw);
    const uint32_t base = (raw0 + 1023u) & ~1023u;
    // after the stages: full[STAGES], empty[STAGES], tmem_full, tmem_empty, tmem_ptr | perm[2][16] | scaleB[2][64]
    const uint32_t bars = base + C::STAGES * C::STAGE;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * C::STAGES, bar_tfull = bars + 16 * C::STAGES;
    const uint32_t bar_tempty = bar_tfull + 8, tmem_slot = bar_tfull + 16;
    unsigned char *tail = smem_raw + (bars - raw0);
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(tail + 16 * C::STAGES + 16);
    int *s_perm = reinterpret_cast<int *>(tail + 16 * C::STAGES + 32);                 // [2][SP_BS]
    double *s_scale = reinterpret_cast<double *>(tail + 16 * C::STAGES + 32 + 128);    // [2][SP_BN]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KS = a.k_steps;

    if (threadIdx.x == 0) {
        for (int i = 0; i < C::STAGES; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, 1);
        }
        mbar_init(bar_tfull, 1);
        mbar_init(bar_tempty, 4 * EW);                               // one arrive per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(tmem_slot), "n"(C::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot_ptr;
    // barriers and TMEM are set up while the digit kernel before this launch is still running; nothing it wrote
    // is touched before this point
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 900
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif

    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmA) : "memory");
            asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmB) : "memory");
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
                const int m0 = (t % a.m_tiles) * SP_BM;              // m fastest: CTAs sharing plane digits run together
                const int n0 = ((t / a.m_tiles) * a.tile_stride + a.tile_offset) * SP_BN;
                for (int ks = 0; ks < KS; ++ks) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
                    const uint32_t sa = base + stage * C::STAGE, sb = sa + C::STAGE_A;
                    mbar_expect_tx(bar_full + 8 * stage, C::STAGE);
                    tma_load_3d(sa, &tmA, ks * BK, m0, 0, bar_full + 8 * stage);
                    tma_load_3d(sb, &tmB, ks * BK, n0, 0, bar_full + 8 * stage);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, tphase = 0;
            for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
                mbar_wait(bar_tempty, tphase ^ 1u);                  // the epilogue has drained the previous tile
                tc_fence_after();
                for (int ks = 0; ks < KS; ++ks) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = base + stage * C::STAGE, sb = sa + C::STAGE_A;
                    if constexpr (BK == SP_BK) {
#pragma unroll
                        for (int i = 0; i < SD; ++i) {
                            const uint64_t adesc = tc_smem_desc(sa + i * (SP_BM * SP_BK));
                            const int n_total = (SD - i) * SP_BN;
#pragma unroll
                            for (int c0 = 0; c0 < n_total; c0 += 256) {
                                const int n = (n_total - c0 < 256) ? (n_total - c0) : 256;
                                tc_mma_i8(tmem + (uint32_t)(i * SP_BN + c0), adesc, tc_smem_desc(sb + c0 * SP_BK), tc_idesc_i8(n),
                                          (ks > 0 || i > 0) ? 1u : 0u);
                            }
                        }
                    } else {
#pragma unroll
                        for (int kk = 0; kk < BK / SP_BK; ++kk) {    // 32-byte K slabs of the stage
#pragma unroll
                            for (int i = 0; i < SD; ++i) {
                                const uint64_t adesc = tc_smem_desc<BK>(sa + i * (SP_BM * BK) + kk * SP_BK);
                                const int n_total = (SD - i) * SP_BN;
#pragma unroll
                                for (int c0 = 0; c0 < n_total; c0 += 256) {
                                    const int n = (n_total - c0 < 256) ? (n_total - c0) : 256;
                                    tc_mma_i8(tmem + (uint32_t)(i * SP_BN + c0), adesc,
                                              tc_smem_desc<BK>(sb + c0 * BK + kk * SP_BK), tc_idesc_i8(n),
                                              (ks > 0 || kk > 0 || i > 0) ? 1u : 0u);
                                }
                            }
                        }
                    }
                    tc_commit(bar_empty + 8 * stage);                // frees the stage once its MMAs have read it
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                }
                tc_commit(bar_tfull);                                // accumulators of this tile complete
                tphase ^= 1u;
            }
        }
    } else {
        // ---- epilogue: this warp reads TMEM lanes [32 q, 32 q + 32), q = warp % 4, and of those the columns
        // [part * COLS_W, (part + 1) * COLS_W) of every accumulator, CW at a time -------------------------------
        static_assert(EW == 1 || EW == 2 || EW == 4, "epilogue warps per lane quarter");
        static_assert((CW == 16 || CW == 8) && (SP_BN / EW) % CW == 0, "chunk width");
        constexpr int COLS_W = SP_BN / EW;
        constexpr int NCH = COLS_W / CW;
        const int q = warp & 3;
        const int part = (EW == 1) ? 0 : (warp - 2) >> 2;            // 0 .. EW-1
        const int e = threadIdx.x - 64;                              // 0 .. 128 EW - 1
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t tphase = 0;
        int buf = 0;
        for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x, buf ^= 1) {
            const int m0 = (t % a.m_tiles) * SP_BM;
            const int s0 = ((t / a.m_tiles) * a.tile_stride + a.tile_offset) * SP_BS;
            // tile constants -> shared memory, hidden behind the tile's MMAs
            if (e < SP_BS) s_perm[buf * SP_BS + e] = (s0 + e < a.S) ? ((a.perm && !a.rows_by_slot) ? a.perm[s0 + e] : s0 + e) : -1;
            if (e < SP_BN) s_scale[buf * SP_BN + e] = (s0 + (e >> 2) < a.S) ? a.scaleB[(size_t)s0 * 4 + e] : 0.0;
            const int m = m0 + q * 32 + lane;
            const bool m_ok = m < a.M;
            const double sA = m_ok ? a.scaleA[m] : 0.0;
            const double bias = (m_ok && a.bias != nullptr) ? a.bias[m] : 0.0;
            double add[3] = {0.0, 0.0, 0.0};
            if (m_ok && a.add_in != nullptr) { add[0] = a.add_in[3 * m]; add[1] = a.add_in[3 * m + 1]; add[2] = a.add_in[3 * m + 2]; }
            if (m_ok && a.add_identity && m < 3) add[m] = 1.0;
            const bool has_add = a.add_in != nullptr || a.add_identity;
            asm volatile("bar.sync 1, %0;\n" ::"n"(128 * EW) : "memory");
            mbar_wait(bar_tfull, tphase);
            tphase ^= 1u;
            tc_fence_after();
#pragma unroll 1
            for (int cb = 0; cb < NCH; ++cb) {
                const int c0 = part * COLS_W + cb * CW;              // first accumulator column of the chunk
                int v[SD][CW];
#pragma unroll
                for (int g = 0; g < SD; ++g) tc_ld_cols(trow + (uint32_t)(g * SP_BN + c0), v[g]);
                tc_ld_wait();
                if (cb == NCH - 1) {                                 // this warp's share of TMEM is read: the next tile's
                    tc_fence_before();                               // MMAs start when every epilogue warp has arrived
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar_tempty) : "memory");
                }
#pragma unroll
                for (int st = 0; st < CW / 4; ++st) {
                    const int s = s_perm[buf * SP_BS + (c0 >> 2) + st];
                    if (s < 0 || !m_ok) continue;
                    const double *sb = s_scale + buf * SP_BN + c0 + st * 4;
                    double r[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        double acc = i32_to_f64(v[SD - 1][st * 4 + c]);
#pragma unroll
                        for (int g = SD - 2; g >= 0; --g) acc = fma(acc, 0.00390625, i32_to_f64(v[g][st * 4 + c]));
                        r[c] = acc * sA * sb[c];
                    }
                    r[3] += bias;
                    if (has_add) { r[0] += add[0]; r[1] += add[1]; r[2] += add[2]; }
                    double2 *dst = reinterpret_cast<double2 *>(a.out + (size_t)s * a.out_stride + (size_t)m * 4);
                    if (a.accumulate) {
                        const double2 o0 = dst[0], o1 = dst[1];
                        r[0] = o0.x + r[0]; r[1] = o0.y + r[1]; r[2] = o1.x + r[2]; r[3] = o1.y + r[3];
                    }
                    dst[0] = make_double2(r[0], r[1]);
                    dst[1] = make_double2(r[2], r[3]);
                }
            }
        }
    }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(C::TMEM_COLS) : "memory");
    }
}

}  // namespace amb
