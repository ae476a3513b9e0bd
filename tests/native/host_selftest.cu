// Host-side checks of the pieces of the CUDA sources that are plain arithmetic (compiled with nvcc, run on the CPU;
// no GPU needed): the pairing bijection and the counter-based generator of the device initialiser (seeds.cuh), the
// digit split of the tcgen05 path (split.cuh), the exchange-block layout and the hash ownership of the sharded march
// (xchg.cuh / frontier.cuh), the chain deal of the stream chains (compose.cuh).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../analyticmesh_b200/csrc/seeds.cuh"
#include "../../analyticmesh_b200/csrc/split.cuh"
#include "../../analyticmesh_b200/csrc/xchg.cuh"

using namespace amb;
static int fails = 0;
#define CHECK(c) do { if (!(c)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); ++fails; } } while (0)

static int chain_pos_host(int i, int tile, int stride, int offset)
{
    return (stride <= 1) ? i : ((i / tile) * stride + offset) * tile + (i % tile);
}

int main()
{
    // 1. feistel_permute is a bijection of [0, n) for any n and key
    for (uint64_t n : {1ull, 2ull, 3ull, 17ull, 1000ull, 4096ull, 65537ull}) {
        std::vector<char> seen(n, 0);
        for (uint64_t i = 0; i < n; ++i) {
            const uint64_t j = feistel_permute(i, n, 0x1234567ull + n);
            CHECK(j < n);
            if (j < n) { CHECK(!seen[j]); seen[j] = 1; }
        }
    }
    // two keys give different permutations
    {
        int same = 0;
        for (uint64_t i = 0; i < 1000; ++i) same += feistel_permute(i, 1000, 1) == feistel_permute(i, 1000, 2);
        CHECK(same < 50);
    }
    // 2. Philox4x32-10: deterministic, counter sensitive, roughly uniform
    {
        uint32_t a[4] = {1, 2, 3, 4}, b[4] = {1, 2, 3, 4}, c[4] = {2, 2, 3, 4};
        philox4x32(a, 7, 9); philox4x32(b, 7, 9); philox4x32(c, 7, 9);
        CHECK(memcmp(a, b, sizeof a) == 0);
        CHECK(memcmp(a, c, sizeof a) != 0);
        double mean = 0;
        for (uint32_t t = 0; t < 20000; ++t) {
            uint32_t x[4] = {t, 0, 0, 0x5EED0001u};
            philox4x32(x, 5, 0);
            const double u = u32_to_unit(x[0]);
            CHECK(u > 0.0 && u < 1.0);
            mean += u;
        }
        mean /= 20000;
        CHECK(mean > 0.49 && mean < 0.51);
    }
    // 3. split_pack: X = sum_t d_t 256^(SD-1-t) with signed 8-bit digits, for every digit count of the tcgen05 path
    for (int sd = 6; sd <= 8; ++sd) {
        const long long lim = 1ll << (8 * sd - 2);
        for (long long X : {0ll, 1ll, -1ll, 127ll, 128ll, -128ll, -129ll, lim, -lim, lim - 12345, 987654321012ll % lim}) {
            const unsigned long long Y = split_pack(X, sd);
            long long back = 0;
            for (int t = 0; t < sd; ++t) back = back * 256 + (long long)(signed char)((Y >> (8 * (sd - 1 - t))) & 0xFF);
            CHECK(back == X);
        }
    }
    // 4. exchange block layout: regions do not overlap, everything 16-byte aligned
    {
        XchgLayout lay{};
        const int world = 8;
        lay.cap_corners = 3000000;
        lay.region_bytes = (lay.xyz_off() + (size_t)lay.cap_corners * 24 + 255) & ~size_t(255);
        lay.mask_cap = 1 << 21;
        lay.mask_base = XCHG_CTRL_BYTES + (size_t)world * lay.region_bytes;
        lay.cnt_base = lay.mask_base + (size_t)world * lay.mask_cap * 4;
        lay.where_base = lay.cnt_base + (size_t)lay.mask_cap * 4;
        CHECK(lay.xyz_off() >= (size_t)lay.cap_corners * 4 && lay.xyz_off() % 16 == 0);
        for (int r = 0; r + 1 < world; ++r) CHECK(lay.region(r) + lay.region_bytes <= lay.region(r + 1));
        CHECK(lay.region(world - 1) + lay.region_bytes <= lay.mask(0));
        CHECK(lay.mask(world - 1) + (size_t)lay.mask_cap * 4 <= lay.cnt_base);
        CHECK(lay.region(3) % 16 == 0 && lay.mask(3) % 16 == 0 && lay.cnt_base % 16 == 0 && lay.where_base % 8 == 0);
        CHECK(XCHG_MAX_WORLD * 32 * 4 <= (int)XCHG_CTRL_BYTES);     // one 128-byte flag line per source rank
        CHECK(PEER_MAX == XCHG_MAX_WORLD);
    }
    // 5. hash ownership: a partition of the key space, balanced
    for (int world : {2, 3, 8, 16}) {
        std::vector<int> cnt(world, 0);
        for (uint64_t i = 0; i < 200000; ++i) {
            const int o = int(((splitmix64(i) >> 32) * (uint64_t)world) >> 32);
            CHECK(o >= 0 && o < world);
            ++cnt[o];
        }
        for (int r = 0; r < world; ++r) CHECK(cnt[r] > 0.9 * 200000 / world && cnt[r] < 1.1 * 200000 / world);
    }
    // 6. the chain deal: every list position is handled by exactly one chain
    for (int nc : {1, 2, 4, 8})
        for (int n : {1, 15, 16, 17, 1000, 4099}) {
            const int tile = 16;
            std::vector<int> hit(n, 0);
            for (int c = 0; c < nc; ++c) {
                const int tiles = (n + tile - 1) / tile, mine = (tiles - c + nc - 1) / nc;
                for (int i = 0; i < (mine > 0 ? mine * tile : 0); ++i) {
                    const int pos = chain_pos_host(i, tile, nc, c);
                    if (pos < n) ++hit[pos];
                }
            }
            for (int i = 0; i < n; ++i) CHECK(hit[i] == 1);
        }
    printf(fails ? "host selftest: %d FAILED\n" : "host selftest ok\n", fails);
    return fails ? 1 : 0;
}
