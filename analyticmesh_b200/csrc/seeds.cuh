// seeds.cuh -- the surface-point initialiser on the device ("dichotomy").
//
// Replaces reference backend/main.py:252-326 (dichotomy), :83-91 (init_within_ball: CPU rejection sampling),
// :70-80 (constraints_filter), python `random.sample` for the pairing and :408-411 (a forward pass that
// materialises an (N, L) bool tensor for the states).  Here:
//   * trial points come from a counter-based generator (Philox4x32-10 keyed by the user's seed: point t of round r
//     is a pure function of (seed, r, t)), rejected against the ball and the extra constraints by the thread that
//     draws them;
//   * f is evaluated by a small FP64 tile GEMM per layer over all points at once (weights shared, like the march);
//   * positive / negative trial points are paired WITHOUT replacement by a keyed bijection of [0, n_pos * n_neg)
//     (cycle-walking Feistel network) instead of python's random.sample;
//   * the bisection runs on the device; the host only reads the mean |f - iso| per iteration (the reference's
//     stopping rule, main.py:318-322);
//   * the activation pattern of the final points is written straight as packed keys (bit j <-> word j / 32, bit
//     j % 32) -- no (N, L) byte tensor, no H2D of 4 MB of states.
#pragma once
#include "common.cuh"

namespace amb {

// ---- Philox4x32-10 ----------------------------------------------------------------------------------------
__host__ __device__ inline void philox4x32(uint32_t (&ctr)[4], uint32_t k0, uint32_t k1)
{
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * ctr[0], p1 = (uint64_t)0xCD9E8D57u * ctr[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ ctr[1] ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ ctr[3] ^ k1, n3 = (uint32_t)p0;
        ctr[0] = n0; ctr[1] = n1; ctr[2] = n2; ctr[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
__host__ __device__ inline double u32_to_unit(uint32_t x) { return ((double)x + 0.5) * (1.0 / 4294967296.0); }   // (0, 1)

// trial points: uniform in the ball of radius R, inside every extra constraint (w . x + b < 0)
__global__ void seed_sample_kernel(double *pts, int *valid, int T, double R, const double *extra /*[E][4]*/, int E,
                                   uint64_t seed, uint32_t round)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    double x = 0, y = 0, z = 0;
    int ok = 0;
    for (uint32_t attempt = 0; attempt < 256u && !ok; ++attempt) {
        uint32_t c[4] = {(uint32_t)t, round, attempt, 0x5EED0001u};
        philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        x = (2.0 * u32_to_unit(c[0]) - 1.0) * R;
        y = (2.0 * u32_to_unit(c[1]) - 1.0) * R;
        z = (2.0 * u32_to_unit(c[2]) - 1.0) * R;
        ok = (x * x + y * y + z * z < R * R);
        for (int e = 0; e < E && ok; ++e)
            ok = (extra[4 * e] * x + extra[4 * e + 1] * y + extra[4 * e + 2] * z + extra[4 * e + 3] < 0.0);
    }
    pts[3 * t] = x; pts[3 * t + 1] = y; pts[3 * t + 2] = z;
    valid[t] = ok;
}

// ---- forward pass: one tile GEMM per fully connected layer --------------------------------------------------
// out[p][m] (+)= sum_k in[p][k] * W[m][k] (+ bias[m]);  in: [P][ldi], W row-major [M][K], out: [P][ldo]
constexpr int SG_T = 64, SG_K = 16;
__global__ void __launch_bounds__(256) seed_gemm_kernel(const double *in, int ldi, const double *W, const double *bias,
                                                        double *out, int ldo, int P, int M, int K, int accumulate)
{
    __shared__ double sA[SG_K][SG_T + 1], sB[SG_K][SG_T + 1];
    const int p0 = blockIdx.y * SG_T, m0 = blockIdx.x * SG_T;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += SG_K) {
        for (int i = threadIdx.x; i < SG_T * SG_K; i += 256) {
            const int r = i / SG_K, c = i % SG_K;
            sA[c][r] = (p0 + r < P && k0 + c < K) ? in[(size_t)(p0 + r) * ldi + k0 + c] : 0.0;
            sB[c][r] = (m0 + r < M && k0 + c < K) ? W[(size_t)(m0 + r) * K + k0 + c] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SG_K; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = sA[kk][ty * 4 + i]; b[i] = sB[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            const int p = p0 + ty * 4 + i, m = m0 + tx * 4 + j;
            if (p < P && m < M) {
                double v = acc[i][j] + (bias ? bias[m] : 0.0);
                if (accumulate) v += out[(size_t)p * ldo + m];
                out[(size_t)p * ldo + m] = v;
            }
        }
}

// identity skip: out[p][m] += src[p][m] (m < min(M, src width))
__global__ void seed_add_kernel(double *out, int ldo, const double *src, int lds, int P, int M)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)P * M) return;
    const int p = int(i / M), m = int(i % M);
    out[(size_t)p * ldo + m] += src[(size_t)p * lds + m];
}

// ReLU in place + the layer's activation bits into the packed keys (thread per (point, 32-bit word of the layer))
__global__ void seed_relu_bits_kernel(double *act, int ld, int P, int M, int bit0, uint32_t *keys, int kw)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int words = (M + 31) / 32;
    if (i >= (long long)P * words) return;
    const int p = int(i / words), w = int(i % words);
    uint32_t bits = 0;
    for (int b = 0; b < 32; ++b) {
        const int m = w * 32 + b;
        if (m >= M) break;
        double &v = act[(size_t)p * ld + m];
        if (v > 0.0) bits |= 1u << b;
        else v = 0.0;
    }
    // the layer starts at bit0, which need not be word aligned: split over two key words
    const int g = bit0 + w * 32, gw = g >> 5, sh = g & 31;
    if (keys != nullptr && bits) {
        atomicOr(keys + (size_t)p * kw + gw, bits << sh);
        if (sh && (bits >> (32 - sh))) atomicOr(keys + (size_t)p * kw + gw + 1, bits >> (32 - sh));
    }
}

// ---- pairing: keyed bijection of [0, n) (cycle-walking Feistel over the next power of four) ------------------
__host__ __device__ inline uint64_t feistel_permute(uint64_t i, uint64_t n, uint64_t key)
{
    int half_bits = 1;
    while ((1ull << (2 * half_bits)) < n) ++half_bits;
    const uint64_t mask = (1ull << half_bits) - 1;
    uint64_t x = i;
    do {
        uint64_t l = x >> half_bits, r = x & mask;
        for (int round = 0; round < 4; ++round) {
            const uint64_t f = splitmix64(r ^ (key + 0x9E3779B97F4A7C15ull * (uint64_t)(round + 1))) & mask;
            const uint64_t nl = r, nr = l ^ f;
            l = nl; r = nr;
        }
        x = (l << half_bits) | r;
    } while (x >= n);
    return x;
}

// sign classes of the trial points: cls[t] = +1 / -1 / 0 (invalid or exactly on the surface)
__global__ void seed_classify_kernel(const double *val, const int *valid, int T, double iso, uint32_t *is_pos, uint32_t *is_neg)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const double v = val[t] - iso;
    is_pos[t] = (valid[t] && v > 0.0) ? 1u : 0u;
    is_neg[t] = (valid[t] && v < 0.0) ? 1u : 0u;
}
// index lists from the exclusive scans
__global__ void seed_lists_kernel(const uint32_t *is_pos, const uint32_t *off_pos, const uint32_t *is_neg,
                                  const uint32_t *off_neg, int T, int *pos_list, int *neg_list)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    if (is_pos[t]) pos_list[off_pos[t]] = t;
    if (is_neg[t]) neg_list[off_neg[t]] = t;
}
// pair i of this round: index = permute(i) in [0, n_pos * n_neg) -> (pos[index % n_pos], neg[index / n_pos])
// (the reference's index arithmetic, main.py:293-297)
__global__ void seed_pairs_kernel(const double *pts, const int *pos_list, const int *neg_list, int n_pos, int n_neg,
                                  int n_take, int dst0, uint64_t key, double *pt_pos, double *pt_neg)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_take) return;
    const uint64_t idx = feistel_permute((uint64_t)i, (uint64_t)n_pos * (uint64_t)n_neg, key);
    const int a = pos_list[idx % (uint64_t)n_pos], b = neg_list[idx / (uint64_t)n_pos];
    for (int c = 0; c < 3; ++c) {
        pt_pos[3 * (size_t)(dst0 + i) + c] = pts[3 * (size_t)a + c];
        pt_neg[3 * (size_t)(dst0 + i) + c] = pts[3 * (size_t)b + c];
    }
}

// ---- bisection ----------------------------------------------------------------------------------------------
__global__ void seed_mid_kernel(const double *pt_pos, const double *pt_neg, double *mid, int n3)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n3) mid[i] = (pt_neg[i] + pt_pos[i]) / 2;
}
// sum_i |f(mid_i) - iso| with a fixed summation order (one block): the stopping rule must not depend on timing
__global__ void __launch_bounds__(1024) seed_err_kernel(const double *val, double iso, int N, double *out)
{
    __shared__ double sh[1024];
    double e = 0.0;
    for (int i = threadIdx.x; i < N; i += 1024) e += fabs(val[i] - iso);
    sh[threadIdx.x] = e;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0];
}
// the bracket shrinks to the half that still changes sign (main.py:323-326)
__global__ void seed_bisect_kernel(const double *val, double iso, const double *mid, double *pt_pos, double *pt_neg, int N)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double *dst = (val[i] - iso < 0.0) ? pt_neg : pt_pos;
    for (int c = 0; c < 3; ++c) dst[3 * (size_t)i + c] = mid[3 * (size_t)i + c];
}

}  // namespace amb
