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
// states.  Register-tiled DFMA: on B200 the DFMA pipe and DMMA (mma.sync f64) both peak at
// 37.0 TFLOP/s (tools/fp64_peak.cu, measured), and tcgen05 has no FP64 kind, so SIMT DFMA is the
// roofline-equivalent choice -- and it keeps every output a single ascending-k FMA chain, i.e.
// bit-identical to the CPU restatement used by the parity tests.
//
// Tile: 128 (m) x 32 states (= 128 columns) x 16 (k), 256 threads, 8x8 outputs per thread,
// 3-stage cp.async pipeline, 97.5 KiB shared memory, one CTA per SM.
#pragma once
#include "common.cuh"

namespace amb {

constexpr int GM_BM = 128;
constexpr int GM_BS = 32;                 // states per tile
constexpr int GM_BN = GM_BS * 4;          // 128 columns
constexpr int GM_BK = 16;
constexpr int GM_STAGES = 3;
constexpr int GM_THREADS = 256;
constexpr int GM_BS_STRIDE = GM_BN + 4;   // +32 B per k-row: conflict-free cp.async writes
constexpr int GM_SMEM_A = GM_BK * GM_BM;                 // doubles per stage
constexpr int GM_SMEM_B = GM_BK * GM_BS_STRIDE;
constexpr size_t GM_SMEM_BYTES = size_t(GM_STAGES) * (GM_SMEM_A + GM_SMEM_B) * sizeof(double);

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
    int S;                      // states in the chunk
    int accumulate;             // out += result (hidden-layer skip with a transform)
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// 16 consecutive activation bits starting at bit position `pos` of a key
__device__ __forceinline__ uint32_t key_bits16(const uint32_t *key, int kw, int pos)
{
    const int w = pos >> 5;
    const uint32_t lo = key[w];
    const uint32_t hi = (w + 1 < kw) ? key[w + 1] : 0u;
    return __funnelshift_r(lo, hi, pos & 31) & 0xFFFFu;
}

__global__ void __launch_bounds__(GM_THREADS, 1) compose_gemm_kernel(const GemmArgs a)
{
    extern __shared__ __align__(16) double smem[];
    double *As = smem;                               // [STAGES][BK][BM]
    double *Bs = smem + GM_STAGES * GM_SMEM_A;       // [STAGES][BK][BS_STRIDE]

    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * GM_BM;
    const int s0 = blockIdx.y * GM_BS;
    const int KT = (a.K + GM_BK - 1) / GM_BK;

    // ---- producer mapping ------------------------------------------------------------------
    // A: 16 rows x 1 KiB = 1024 chunks of 16 B, 4 per thread, rows of 64 chunks (coalesced)
    // B: per state 16 k x 32 B = 512 B contiguous in global; warp w stages states w, w+8, w+16, w+24
    const int lane = tid & 31, warp = tid >> 5;
    const int bk = lane >> 1, bhalf = lane & 1;      // k row and (xy | zc) half handled by this lane

    auto load_stage = [&](int kt, int slot) {
        double *as = As + slot * GM_SMEM_A;
        double *bs = Bs + slot * GM_SMEM_B;
        const int k0 = kt * GM_BK;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = tid + i * GM_THREADS;      // 0..1023
            const int row = c >> 6, col = (c & 63) * 2;
            cp_async16(as + row * GM_BM + col, a.Wt + (size_t)(k0 + row) * a.Mpad + m0 + col, 16);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int sl = warp + 8 * i;             // local state
            const int s = s0 + sl;
            const int k = k0 + bk;
            int bytes = 0;
            const double *src = a.Bsrc;
            if (s < a.S && k < a.K) {
                const uint32_t bits = key_bits16(a.keys + (size_t)s * a.kw, a.kw, a.bit0 + k0);
                if ((bits >> bk) & 1u) {
                    bytes = 16;
                    src = a.Bsrc + (size_t)s * a.b_stride + (size_t)k * 4 + bhalf * 2;
                }
            }
            cp_async16(bs + bk * GM_BS_STRIDE + sl * 4 + bhalf * 2, src, bytes);
        }
    };

    // ---- consumer mapping ------------------------------------------------------------------
    const int tx = tid & 15, ty = tid >> 4;          // 16 x 16 threads
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int st = 0; st < GM_STAGES - 1; ++st) {
        if (st < KT) load_stage(st, st);
        cp_async_commit();
    }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<GM_STAGES - 2>();
        __syncthreads();
        if (kt + GM_STAGES - 1 < KT) load_stage(kt + GM_STAGES - 1, (kt + GM_STAGES - 1) % GM_STAGES);
        cp_async_commit();

        const double *as = As + (kt % GM_STAGES) * GM_SMEM_A + ty * 8;
        const double *bs = Bs + (kt % GM_STAGES) * GM_SMEM_B + tx * 2;
#pragma unroll
        for (int k = 0; k < GM_BK; ++k) {
            double av[8], bv[4][2];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double2 t = *reinterpret_cast<const double2 *>(as + k * GM_BM + 2 * i);
                av[2 * i] = t.x;
                av[2 * i + 1] = t.y;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double2 t = *reinterpret_cast<const double2 *>(bs + k * GM_BS_STRIDE + 32 * j);
                bv[j][0] = t.x;
                bv[j][1] = t.y;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i][j][0] = fma(av[i], bv[j][0], acc[i][j][0]);
                    acc[i][j][1] = fma(av[i], bv[j][1], acc[i][j][1]);
                }
        }
    }
    cp_async_wait<0>();

    // ---- epilogue: + bias on column 3, store 16 B per (row, state) ---------------------------
    const int comp0 = (tx & 1) * 2;                  // this thread holds components comp0, comp0+1
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int s = s0 + 8 * j + (tx >> 1);
        if (s >= a.S) continue;
        double *dst = a.out + (size_t)s * a.out_stride + comp0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = m0 + ty * 8 + i;
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
__global__ void skip_input_kernel(double *out, long long out_stride, int M, int S, const double *T)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)S * M) return;
    const int s = int(t / M), m = int(t % M);
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
                                            const uint32_t *keys, int kw, int src_bit0, int M, int S)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)S * M * 4) return;
    const int c = int(t & 3);
    const long long r = t >> 2;
    const int s = int(r / M), m = int(r % M);
    const int bit = src_bit0 + m;
    if ((keys[(size_t)s * kw + (bit >> 5)] >> (bit & 31)) & 1u)
        out[(size_t)s * out_stride + (size_t)m * 4 + c] += in[(size_t)s * in_stride + (size_t)m * 4 + c];
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
    int n_skips;
    EquSkip skips[EQU_MAX_SKIPS];
};

__device__ __forceinline__ double masked_chain(const double *w, const double *rows, const uint32_t *key, int bit0,
                                               int K, int c)
{
    double acc = 0.0;
    for (int k = 0; k < K; ++k) {
        const int bit = bit0 + k;
        if ((key[bit >> 5] >> (bit & 31)) & 1u) acc = fma(w[k], rows[(size_t)k * 4 + c], acc);
    }
    return acc;
}

__global__ void equ_kernel(const EquArgs a)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.S * 4) return;
    const int s = t >> 2, c = t & 3;
    const uint32_t *key = a.keys + (size_t)s * a.kw;
    double v = masked_chain(a.w, a.in + (size_t)s * a.in_stride, key, a.bit0, a.K, c);
    if (c == 3) v += a.bias;
    for (int i = 0; i < a.n_skips; ++i) {
        const EquSkip &sk = a.skips[i];
        if (sk.kind == 1) {
            if (c == 0) v += 1.0;
        } else if (sk.kind == 2) {
            if (c < 3) v += sk.T[c];
        } else if (sk.kind == 3) {
            if ((key[sk.src_bit0 >> 5] >> (sk.src_bit0 & 31)) & 1u) v += sk.src[(size_t)s * sk.src_stride + c];
        } else if (sk.kind == 4) {
            v += masked_chain(sk.T, sk.src + (size_t)s * sk.src_stride, key, sk.src_bit0, sk.src_n, c);
        }
    }
    if (c == 3) v -= a.iso;
    a.equ[(size_t)s * 4 + c] = v;
}

}  // namespace amb
