// am_engine.cu -- host driver + C ABI of the B200-native Analytic Marching engine.
//
// Replaces reference backend/src/cuam_kernel.cu (global singleton, host-driven LIFO loop with
// O(100-1000) launches and blocking scalar copies per batch of 1024 states) by a handle-based,
// level-synchronous BFS that keeps the frontier, the visited set and the mesh on the device and
// synchronises with the host once per BFS level (one 16-byte read-back of the counters).
//
// Per level [lb, le):   for each chunk of states
//                           compose (one FP64 GEMM launch per hidden layer, compose.cuh)
//                           level plane (equ_kernel), clip (clip.cuh), scan + compact faces
//                       expand + insert (frontier.cuh phase 1), count winners, scan,
//                       -> read back n_new, grow arenas, finalize (phase 2)
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include <cuda.h>
#include <dlfcn.h>
#include <fcntl.h>
#include <unistd.h>

#include "../../include/am_b200.h"
#include "clip.cuh"
#include "common.cuh"
#include "compose.cuh"
#include "frontier.cuh"
#include "mesh.cuh"
#include "seeds.cuh"
#include "split.cuh"
#include "xchg.cuh"

using namespace amb;

namespace {

thread_local std::string g_create_error;

struct CudaFail {
    std::string msg;
    int code = AM_ERR_CUDA;
};
struct CapacityFail : CudaFail {
    explicit CapacityFail(std::string m) : CudaFail{std::move(m), AM_ERR_CAPACITY} {}
};

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess)                                                                           \
            throw CudaFail{std::string(#call) + " failed: " + cudaGetErrorString(e__) + " (" __FILE__ ":" + \
                           std::to_string(__LINE__) + ")"};                                               \
    } while (0)

// Launch with the programmatic-stream-serialization attribute (common.cuh pdl_enter): the kernel may be scheduled
// while its predecessor in the stream is still running.  AM_B200_PDL=0 falls back to plain launches.
// AM_B200_PDL: 0 never, 1 (default) on BFS levels that are small for this rank -- the launch-latency-bound ones;
// measured on one B200: wide levels run ~5 % slower with the attribute than without -- 2 always.
int pdl_mode()
{
    static const int mode = [] {
        const char *e = getenv("AM_B200_PDL");
        return e ? atoi(e) : 1;
    }();
    return mode;
}
thread_local bool g_pdl_now = false;      // set per BFS level by process_level

template <typename... KArgs, typename... Args>
void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = g_pdl_now ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...));
}

template <typename F>
void for_split_digits(F &&f)
{
    f(std::integral_constant<int, 6>{});
    f(std::integral_constant<int, 7>{});
    f(std::integral_constant<int, 8>{});
}

// ---- virtual-memory backed growth (CUDA VMM through the runtime's driver entry points: no libcuda link)
struct VmApi {
    bool ok = false;
    size_t gran = 0;
    int device = 0;
    decltype(&cuMemAddressReserve) reserve = nullptr;
    decltype(&cuMemAddressFree) free_va = nullptr;
    decltype(&cuMemCreate) create = nullptr;
    decltype(&cuMemRelease) release = nullptr;
    decltype(&cuMemMap) map = nullptr;
    decltype(&cuMemUnmap) unmap = nullptr;
    decltype(&cuMemSetAccess) set_access = nullptr;
    decltype(&cuMemGetAllocationGranularity) granularity = nullptr;
    CUmemAllocationProp prop{};
    size_t va_span = 0;          // virtual range reserved per buffer

    static VmApi &get()
    {
        static VmApi api = [] {
            VmApi a;
            if (const char *e = getenv("AM_B200_NO_VMM")) if (atoi(e)) return a;
            auto sym = [](const char *name, void **fn) {
                cudaDriverEntryPointQueryResult q;
                return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess &&
                       q == cudaDriverEntryPointSuccess && *fn != nullptr;
            };
            if (!sym("cuMemAddressReserve", (void **)&a.reserve) || !sym("cuMemAddressFree", (void **)&a.free_va) ||
                !sym("cuMemCreate", (void **)&a.create) || !sym("cuMemRelease", (void **)&a.release) ||
                !sym("cuMemMap", (void **)&a.map) || !sym("cuMemUnmap", (void **)&a.unmap) ||
                !sym("cuMemSetAccess", (void **)&a.set_access) ||
                !sym("cuMemGetAllocationGranularity", (void **)&a.granularity)) {
                cudaGetLastError();
                return a;
            }
            if (cudaGetDevice(&a.device) != cudaSuccess) return a;
            a.prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
            a.prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
            a.prop.location.id = a.device;
            if (a.granularity(&a.gran, &a.prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || a.gran == 0)
                return a;
            size_t free_b = 0, total_b = 0;
            if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return a;
            a.va_span = (total_b + a.gran - 1) / a.gran * a.gran;
            a.ok = true;
            return a;
        }();
        return api;
    }
};

// Device buffer that grows.  With `vm` set (the big append-only arenas) growth maps more physical memory
// behind a reserved virtual range: the pointer never moves, nothing is copied or freed, and no device
// synchronisation is needed.  Otherwise (or if VMM is unavailable) classic malloc + copy + free.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    bool vm = false;
    std::vector<std::pair<CUmemGenericAllocationHandle, size_t>> chunks;
    template <typename T>
    T *as() const { return static_cast<T *>(p); }

    bool grow_vm(size_t bytes)
    {
        VmApi &api = VmApi::get();
        if (!api.ok) return false;
        if (!p) {
            CUdeviceptr va = 0;
            if (api.reserve(&va, api.va_span, 0, 0, 0) != CUDA_SUCCESS) return false;
            p = reinterpret_cast<void *>(va);
        }
        if (bytes > api.va_span) return false;
        size_t add = std::max(bytes - cap, std::max(cap / 2, size_t(32) << 20));
        add = (add + api.gran - 1) / api.gran * api.gran;
        if (cap + add > api.va_span) add = api.va_span - cap;
        CUmemGenericAllocationHandle hnd;
        if (api.create(&hnd, add, &api.prop, 0) != CUDA_SUCCESS) {          // retry with the bare minimum
            add = (bytes - cap + api.gran - 1) / api.gran * api.gran;
            if (api.create(&hnd, add, &api.prop, 0) != CUDA_SUCCESS) throw CudaFail{"out of device memory (cuMemCreate)"};
        }
        const CUdeviceptr at = reinterpret_cast<CUdeviceptr>(p) + cap;
        if (api.map(at, add, 0, hnd, 0) != CUDA_SUCCESS) {
            api.release(hnd);
            throw CudaFail{"cuMemMap failed"};
        }
        CUmemAccessDesc acc{};
        acc.location = api.prop.location;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        if (api.set_access(at, add, &acc, 1) != CUDA_SUCCESS) throw CudaFail{"cuMemSetAccess failed"};
        chunks.emplace_back(hnd, add);
        cap += add;
        return true;
    }

    // grow to at least `bytes`; keeps the first `keep` bytes
    void reserve(size_t bytes, size_t keep = 0, bool geometric = true)
    {
        if (bytes <= cap) return;
        if (vm && (p == nullptr || !chunks.empty())) {
            if (grow_vm(bytes)) return;
            if (!chunks.empty()) throw CudaFail{"virtual range exhausted"};
        }
        size_t want = bytes;
        if (geometric && cap) want = std::max(bytes, cap * 2);
        void *np = nullptr;
        if (p) CK(cudaDeviceSynchronize());   // in-flight kernels on the engine stream may still use `p`
        CK(cudaMalloc(&np, want));
        if (p && keep) CK(cudaMemcpy(np, p, std::min(keep, cap), cudaMemcpyDeviceToDevice));
        if (p) CK(cudaFree(p));
        p = np;
        cap = want;
    }
    void release()
    {
        if (!chunks.empty() || (vm && p && VmApi::get().ok && cap == 0)) {
            VmApi &api = VmApi::get();
            cudaDeviceSynchronize();
            size_t off = 0;
            for (auto &c : chunks) {
                api.unmap(reinterpret_cast<CUdeviceptr>(p) + off, c.second);
                api.release(c.first);
                off += c.second;
            }
            chunks.clear();
            if (p) api.free_va(reinterpret_cast<CUdeviceptr>(p), api.va_span);
        } else if (p) {
            cudaFree(p);
        }
        p = nullptr;
        cap = 0;
    }
};

// ---- NCCL, bound at run time (the library is already loaded in a torch process; no link dependency)
struct NcclApi {
    struct UniqueId { char internal[128]; };
    typedef int (*get_id_t)(UniqueId *);
    typedef int (*init_rank_t)(void **comm, int nranks, UniqueId id, int rank);
    typedef int (*allreduce_t)(const void *send, void *recv, size_t count, int dtype, int op, void *comm,
                               cudaStream_t stream);
    typedef int (*destroy_t)(void *comm);
    get_id_t get_id = nullptr;
    init_rank_t init_rank = nullptr;
    allreduce_t allreduce = nullptr;
    destroy_t destroy = nullptr;
    bool ok = false;
    static NcclApi &get()
    {
        static NcclApi api = [] {
            NcclApi a;
            void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
            if (!lib) return a;
            a.get_id = (get_id_t)dlsym(lib, "ncclGetUniqueId");
            a.init_rank = (init_rank_t)dlsym(lib, "ncclCommInitRank");
            a.allreduce = (allreduce_t)dlsym(lib, "ncclAllReduce");
            a.destroy = (destroy_t)dlsym(lib, "ncclCommDestroy");
            a.ok = a.get_id && a.init_rank && a.allreduce && a.destroy;
            return a;
        }();
        return api;
    }
};

// Pinned host array that only grows: device->host copies of the mesh run at PCIe speed instead of being
// staged through pageable memory, and a repeated am_combine pays neither allocation nor page faults.
template <typename T>
struct HostBuf {
    T *p = nullptr;
    size_t cap = 0, n = 0;
    void resize(size_t want)
    {
        if (want > cap) {
            if (p) cudaFreeHost(p);
            p = nullptr;
            cap = 0;
            const size_t ncap = std::max<size_t>(want + want / 8, 1024);
            CK(cudaMallocHost((void **)&p, ncap * sizeof(T)));
            cap = ncap;
        }
        n = want;
    }
    void release()
    {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = n = 0;
    }
    T *data() { return p; }
    const T *data() const { return p; }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    T &operator[](size_t i) { return p[i]; }
    const T &operator[](size_t i) const { return p[i]; }
};

// ---- tcgen05 split-integer composition (split.cuh): tensor maps + weight digits --------------------------
struct TmaApi {
    decltype(&cuTensorMapEncodeTiled) encode = nullptr;
    static TmaApi &get()
    {
        static TmaApi api = [] {
            TmaApi a;
            cudaDriverEntryPointQueryResult q;
            void *fn = nullptr;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
                q == cudaDriverEntryPointSuccess)
                a.encode = (decltype(a.encode))fn;
            else
                cudaGetLastError();
            return a;
        }();
        return api;
    }
};

// 3-D map over int8 digit planes [SD][rows][pitch]: box = 32 K bytes x box_rows rows x SD planes, SWIZZLE_32B
CUtensorMap make_digit_map(const void *ptr, int k_bytes, size_t pitch, size_t rows, int sd, int box_rows, int bk = SP_BK)
{
    TmaApi &api = TmaApi::get();
    if (!api.encode) throw CudaFail{"cuTensorMapEncodeTiled is not available in this driver"};
    CUtensorMap m;
    const cuuint64_t dims[3] = {(cuuint64_t)k_bytes, (cuuint64_t)rows, (cuuint64_t)sd};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * rows};
    const cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)box_rows, (cuuint32_t)sd};      // rows of bk bytes, swizzled in bk-byte spans
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult rc = api.encode(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(ptr), dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) throw CudaFail{"cuTensorMapEncodeTiled failed (" + std::to_string((int)rc) + ")"};
    return m;
}

// digits of a row-major (M, K) weight matrix, one power-of-two scale per row (see split.cuh)
void split_weight_digits(const double *w, int M, int K, int Mpad, int Kpad, int sd, std::vector<signed char> &dig,
                         std::vector<double> &scale)
{
    dig.assign((size_t)sd * Mpad * Kpad, 0);
    scale.assign(Mpad, 0.0);
    const int frac = 8 * sd - 2;
    for (int m = 0; m < M; ++m) {
        double mx = 0.0;
        for (int k = 0; k < K; ++k) mx = std::max(mx, std::fabs(w[(size_t)m * K + k]));
        int e = 0;
        if (mx >= 2.2250738585072014e-308 && mx < 1.7e308) std::frexp(mx, &e);      // subnormal rows count as zero
        const double mul = std::ldexp(1.0, std::max(-1022, std::min(1023, frac - e)));
        scale[m] = std::ldexp(1.0, std::max(-1022, e - 6));
        for (int k = 0; k < K; ++k) {
            const unsigned long long Y = split_pack(std::llrint(w[(size_t)m * K + k] * mul), sd);
            for (int t = 0; t < sd; ++t)
                dig[((size_t)t * Mpad + m) * Kpad + k] = (signed char)((Y >> (8 * (sd - 1 - t))) & 0xFF);
        }
    }
}

struct SplitWeights {
    DevBuf dig, scale;
    CUtensorMap map;
    int Kpad = 0, Mpad = 0;
    bool ready = false;
};

struct Skip {
    int src;        // 0 = raw input, j >= 1 = hidden layer j
    int tm;         // transform index
};

int pow2_group(int kw4)
{
    int g = 1;
    while (g < kw4 && g < 32) g <<= 1;
    return g;
}

}  // namespace

struct am_handle {
    bool f64 = true;
    int D = 0;
    std::vector<int> n;          // n[0]=3 .. n[D+1]=1
    std::vector<int> off;        // off[h], h=1..D ; off[D+1] = L
    int L = 0, E = 0, kw = 0, kw4 = 0, n1 = 0, R = 0, G = 1;
    std::vector<std::vector<Skip>> skips;   // per fc layer h = 1..D (index h)
    int n_tm = 0;

    // ---- network on the device (double) ----
    std::vector<DevBuf> Wt, bias;           // index h = 1..D-1 (hidden -> hidden)
    std::vector<int> Mpad, Kpad;
    DevBuf P1, wout, extra;
    double bout = 0.0;
    std::vector<DevBuf> TM, TMt;            // row-major (out,in) and k-major padded copies
    std::vector<int> tm_h, tm_w, tm_Mpad;
    bool weights_loaded = false;
    std::vector<unsigned long long> new_sums, old_sums;   // device-side checksums of the tensors (this call / cached)
    std::vector<char> new_sum_valid, old_sum_valid;
    DevBuf sum_dev;
    std::vector<std::vector<unsigned char>> raw_cache;   // raw bytes of the tensors of the last load (see tensor_changed)
    int stats_layers_reloaded = 0;

    // ---- state arena ----
    DevBuf keys, hsum, parent, via, seedpt, face_off;
    size_t cap_states = 0;
    DevBuf face_edges, face_xyz;
    size_t cap_corners = 0;
    DevBuf table;
    uint32_t tcap = 0;
    long long n_states = 0;
    std::vector<long long> level_begin;

    // ---- scratch ----
    DevBuf planes, equ, f_cnt, f_off, f_edges, f_verts, cand_slot, nwin, wbase, scan_a, scan_b, counters;
    DevBuf xkeys, xh, xpt, xslot, xstates;
    DevBuf cmb_owner, cmb_flag, cmb_vid, cmb_cvid, cmb_verts;   // am_combine scratch, kept between calls
    DevBuf digest_acc;
    // incremental composition: plane rows of the previous and the current level stay resident
    DevBuf lvl_planes[2], bucket, perm, bcounts;
    bool prev_resident = false;
    long long prev_lb = 0, prev_S = 0;
    int prev_buf = 0;
    long long n_incremental_levels = 0;
    bool incremental = true;
    size_t resident_budget = 0;                 // bytes for the two resident level buffers
    // sharded mode (one march over several GPUs): compose + clip of a state run on its owner rank only,
    // the per-level polygons are combined with an all-reduce supplied by the host (NCCL through
    // torch.distributed), frontier / visited set / mesh are replicated and stay bit-identical on all ranks
    int gemm_variant = 0;                       // 0/1: FP64 DMMA tiles, 2: tcgen05 int8 split (split.cuh)
    int split_digits = 7;
    int split_bk = 64;            // K bytes per pipeline stage of split_gemm_kernel (AM_B200_SPLIT_BK = 32 | 64): 64-byte TMA rows
                                  // fill shared memory faster than 32-byte rows, GEMM time -9.8 % (profiles/r02_split_bk64.md)
    int split_epi = 1;            // epilogue warps per TMEM lane quarter of split_gemm_kernel (AM_B200_SPLIT_EPI = 1 | 2 | 4;
                                  // 2 and 4 measured 3 % slower, profiles/r02_split_epilogue_warps.md)
    // read-through of the parent's rows instead of copy_parent_rows_kernel (tcgen05 path, no hidden-source skips)
    bool lazy_ok = false;
    const double *lazy_prev = nullptr;          // previous level's rows while a level is processed lazily, else nullptr
    long long lazy_lb = 0, lazy_prev_lb = 0;
    const double *fused_add_in = nullptr;       // input skip of the layer being launched, applied in the GEMM epilogue
    int fused_add_identity = 0;
    // clip kernel variant (AM_B200_CLIP_MINB), measured in profiles/r02_bench_clip_depth.md: 7 (default) = 3 CTAs/SM, one
    // row per lane, ring of 4 blocks per warp (24 instead of 16 resident warps: the kernel is occupancy / latency bound);
    // 9 = same with a ring of 6; 5 / 2 / 6 = 2 CTAs/SM, two rows per lane, ring of 4 / 3 / 5; 3 = 3 CTAs/SM two rows
    // (spills); 8 = 4 CTAs/SM (spills); 4 = TMA 1-D bulk copies
    int clip_minb = 7;
    int num_sms = 148;
    std::vector<SplitWeights> splitW, splitTM;  // index h = 1..D-1 / transform index
    DevBuf bdig, bscale;                        // plane digits [SD][b_ncap][b_pitch] and column scales of the current launch
    size_t b_ncap = 0, b_pitch = 0;
    std::vector<std::pair<int, CUtensorMap>> b_maps;   // per K bytes; rebuilt when bdig moves
    const void *b_maps_ptr = nullptr;
    size_t b_maps_ncap = 0;
    static constexpr int MAX_CHAINS = 8;
    int last_n_chain = 1;
    long long pdl_below = 4096;                 // states per rank and level below which launches use PDL
    bool force_perm_order = false;
    int finalize_G = 8;
    // sharded march: the level buffers hold the owned states' rows only, at their permutation slot (AM_B200_COMPACT_ROWS)
    bool compact_rows = true, prev_compact = false;
    DevBuf slot_of[2];                          // inverse permutation of the previous / current level
    int slot_buf = 0;
    int cur_rows_by_slot = 0;                   // set while a level is processed
    const int *cur_slot_of = nullptr, *cur_prev_slot_of = nullptr;
    bool push_in_clip = true;                   // AM_B200_PUSH_IN_CLIP=0: separate xchg_pack_kernel
    bool equ_warp = true;                       // AM_B200_EQU_WARP=0: the sequential level-plane kernel on every path
    int n_chains = 0;   // 0 = automatic: 1 on a single GPU (launches fill the machine), 4 when sharded (measured +1.4 % at 8 GPUs)
    cudaStream_t chain_stream[MAX_CHAINS] = {};
    cudaEvent_t fork_event = nullptr, join_event[MAX_CHAINS] = {};
    double pending_gemm_flops = 0.0;
    int shard_rank = 0, shard_world = 1;
    am_allreduce_fn shard_cb = nullptr;
    void *shard_user = nullptr;
    void *nccl_comm = nullptr;                  // set by am_set_shard_nccl: collectives issued from C++
    DevBuf owner, xchg;
    long long shard_owned_states = 0;
    // device-initiated exchange over NVLink peer memory (xchg.cuh; am_set_shard_p2p)
    bool p2p = false;
    void *xblock = nullptr;                     // this rank's exchange block (cudaMalloc + cudaIpcGetMemHandle)
    XchgPeers xpeers{};
    XchgLayout xlay{};
    uint32_t xepoch = 0;
    unsigned long long xtimeout_ns = 20000000000ull;
    DevBuf xcursor, wmask;
    // native surface-point initialiser (seeds.cuh): row-major FP64 weights, activations, stored seeds
    std::vector<DevBuf> Wrow, Brow;             // index l = 0..D
    std::vector<DevBuf> sd_act;                 // index l = 1..D: [P][n_l]
    DevBuf sd_pts, sd_valid, sd_val, sd_flags, sd_offs, sd_lists, sd_pos, sd_neg, sd_mid, sd_err, sd_tot, sd_keys;
    long long n_stored_seeds = 0;
    DevBuf bal_loads, bal_cuts;                 // load balance of the sharded march (xchg.cuh winners_scan_kernel)
    bool balance = true;
    DevBuf level_cursor;                        // [D + 3] bucket cursors of classify_scatter_kernel, zero between levels
    DevBuf fs_sums, fs_off, fs_sums2, fs_off2, fs_ticket;   // fused scans (scan.cuh): CSR offsets / new state ids
    bool table_sharded = false;                 // the visited set holds only the keys whose hash this rank owns
    int *h_npre = nullptr;                      // pinned
    int *h_next = nullptr;                      // pinned: bucket histogram of the NEXT level (count_winners_kernel)
    DevBuf next_counts;
    bool next_valid = false;
    unsigned long long *h_counters = nullptr;   // pinned
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t win_event = nullptr;

    // ---- results ----
    bool has_march = false, has_mesh = false;
    am_stats stats{};
    HostBuf<double> h_vertices;
    HostBuf<int> h_corner_vid;
    HostBuf<long long> h_face_off;
    double gemm_ms = 0.0, gemm_flops = 0.0;
    long long gemm_launches = 0;
    // span kinds: 0 composition chain of a chunk, 1 compose phase, 2 clip, 3 frontier, 4 tensor GEMM kernel, 5 digit kernel
    // 6 exchange barriers (time waiting for the slowest rank), 7 polygon push, 8 unpack + scan + CSR, 9 neighbour
    // enumeration + visited-set insert, 10 winner masks / count, 11 finalize
    static constexpr int N_KINDS = 12;
    double host_wait_s = 0.0, host_level_s = 0.0;      // host wall time blocked in the per-level sync / spent per level
    double kind_ms[N_KINDS] = {}, kind_flops[N_KINDS] = {};
    long long kind_launches[N_KINDS] = {};
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    struct Span { size_t a, b; int kind; double flops; };
    std::vector<Span> spans;
    std::string err;

    // ------------------------------------------------------------------------------------------
    cudaEvent_t ev()
    {
        if (ev_used == ev_pool.size()) {
            cudaEvent_t e;
            CK(cudaEventCreate(&e));
            ev_pool.push_back(e);
        }
        cudaEvent_t e = ev_pool[ev_used++];
        CK(cudaEventRecord(e, stream));
        return e;
    }
    // CUDA events between two kernels serialise them (no programmatic overlap), so the per-phase / per-kernel spans
    // are recorded on every trace_every-th BFS level only and scaled up (AM_B200_TRACE_EVERY=1: every level)
    int trace_every = 4, span_scale = 1;
    bool trace_level = true;
    bool timing_on() const { return trace_level && ev_used < 60000; }
    size_t span_begin() { ev(); return ev_used - 1; }
    void span_end(size_t a, int kind, double flops = 0.0)
    {
        ev();
        spans.push_back(Span{a, ev_used - 1, kind, flops});
    }

    void free_all()
    {
        for (auto &b : Wt) b.release();
        for (auto &b : bias) b.release();
        for (auto &b : TM) b.release();
        for (auto &b : TMt) b.release();
        for (auto &w : splitW) { w.dig.release(); w.scale.release(); }
        for (auto &w : splitTM) { w.dig.release(); w.scale.release(); }
        bdig.release(); bscale.release();
        DevBuf *all[] = {&P1, &wout, &extra, &keys, &hsum, &parent, &via, &seedpt, &face_off, &face_edges, &face_xyz,
                         &table, &planes, &equ, &f_cnt, &f_off, &f_edges, &f_verts, &cand_slot, &nwin, &wbase, &scan_a,
                         &scan_b, &counters, &xkeys, &xh, &xpt, &xslot, &xstates, &lvl_planes[0], &lvl_planes[1],
                         &bucket, &perm, &bcounts, &owner, &xchg, &cmb_owner, &cmb_flag, &cmb_vid, &cmb_cvid, &cmb_verts,
                         &digest_acc, &xcursor, &wmask,
                         &level_cursor, &fs_sums, &fs_off, &fs_sums2, &fs_off2, &fs_ticket, &bal_loads, &bal_cuts,
                         &sd_pts, &sd_valid, &sd_val, &sd_flags, &sd_offs, &sd_lists, &sd_pos, &sd_neg, &sd_mid, &sd_err, &sd_tot,
                         &sd_keys, &slot_of[0], &slot_of[1], &sum_dev};
        for (auto &b : Wrow) b.release();
        for (auto &b : Brow) b.release();
        for (auto &b : sd_act) b.release();
        for (DevBuf *b : all) b->release();
        for (cudaEvent_t e : ev_pool) cudaEventDestroy(e);
        ev_pool.clear();
        if (fork_event) cudaEventDestroy(fork_event);
        fork_event = nullptr;
        if (win_event) cudaEventDestroy(win_event);
        win_event = nullptr;
        if (copy_stream) cudaStreamDestroy(copy_stream);
        copy_stream = nullptr;
        for (int c = 0; c < MAX_CHAINS; ++c) {
            if (join_event[c]) cudaEventDestroy(join_event[c]);
            if (chain_stream[c]) cudaStreamDestroy(chain_stream[c]);
            join_event[c] = nullptr;
            chain_stream[c] = nullptr;
        }
        h_vertices.release(); h_corner_vid.release(); h_face_off.release();
        if (h_counters) cudaFreeHost(h_counters);
        h_counters = nullptr;
        if (h_npre) cudaFreeHost(h_npre);
        h_npre = nullptr;
        if (h_next) cudaFreeHost(h_next);
        h_next = nullptr;
        next_counts.release();
    }

    // rows of hidden layer h for state 0 of a chunk + stride (doubles)
    const double *layer_rows(const double *base, int h, long long *stride) const
    {
        if (h == 1) {
            *stride = 0;
            return P1.as<double>();
        }
        *stride = 4LL * R;
        return base + 4LL * (off[h] - n1);
    }

    void ensure_states(size_t want)
    {
        if (want <= cap_states) return;
        size_t ncap = std::max<size_t>(want, std::max<size_t>(cap_states + cap_states / 2, 1 << 16));
        const size_t keep = (size_t)n_states;
        keys.reserve(ncap * kw * 4, keep * kw * 4, false);
        hsum.reserve(ncap * 8, keep * 8, false);
        parent.reserve(ncap * 4, keep * 4, false);
        via.reserve(ncap * 4, keep * 4, false);
        seedpt.reserve(ncap * 32, keep * 32, false);
        face_off.reserve((ncap + 1) * 8, (keep + 1) * 8, false);
        owner.reserve(ncap, keep, false);
        cap_states = ncap;
    }
    void ensure_corners(size_t want, size_t keep)
    {
        if (want <= cap_corners) return;
        size_t ncap = std::max<size_t>(want, std::max<size_t>(cap_corners + cap_corners / 2, 1 << 18));
        face_edges.reserve(ncap * 4, keep * 4, false);
        face_xyz.reserve(ncap * 24, keep * 24, false);
        cap_corners = ncap;
    }
    void ensure_table(size_t entries)
    {
        const bool shard_table = p2p && shard_world > 1 && table_sharded;
        if (shard_table) entries = entries / shard_world + entries / (4 * shard_world) + 4096;   // hash-owned share + slack
        uint64_t want = 1ull << 16;
        while (want < (uint64_t)entries * 2) want <<= 1;
        if (want <= tcap) return;
        if (want > (1ull << 31)) throw CapacityFail{"the visited set would need more than 2^31 slots"};
        table.reserve((size_t)want * 8, 0, false);
        tcap = (uint32_t)want;
        CK(cudaMemsetAsync(table.p, 0xFF, (size_t)tcap * 8, stream));
        if (n_states > 0) {
            TableRef t{table.as<unsigned long long>(), tcap - 1};
            rehash_kernel<<<(unsigned)((n_states + 255) / 256), 256, 0, stream>>>(
                hsum.as<unsigned long long>(), (int)n_states, t, shard_table ? shard_world : 1, shard_rank);
            ++stats.n_launches;
            CK(cudaGetLastError());
        }
    }

    // exclusive scan of n uint32 -> out ; total (64-bit) written to *total_dev if not null
    void scan(const uint32_t *in, uint32_t *out, int n, unsigned long long *total_dev)
    {
        if (n <= 0) {
            if (total_dev) CK(cudaMemsetAsync(total_dev, 0, 8, stream));
            return;
        }
        const int nb = (n + SCAN_TILE - 1) / SCAN_TILE;
        if (nb == 1) {
            scan_apply_kernel<<<1, SCAN_THREADS, 0, stream>>>(in, n, nullptr, out, total_dev);
            ++stats.n_launches;
            CK(cudaGetLastError());
            return;
        }
        scan_a.reserve((size_t)nb * 4);
        scan_b.reserve((size_t)nb * 4 + (size_t)((nb + SCAN_TILE - 1) / SCAN_TILE) * 8 + 64);
        scan_reduce_kernel<<<nb, SCAN_THREADS, 0, stream>>>(in, n, scan_a.as<uint32_t>());
        ++stats.n_launches;
        CK(cudaGetLastError());
        if (nb <= SCAN_TILE) {
            scan_apply_kernel<<<1, SCAN_THREADS, 0, stream>>>(scan_a.as<uint32_t>(), nb, nullptr, scan_b.as<uint32_t>(),
                                                            nullptr);
            ++stats.n_launches;
        } else {  // three levels: up to SCAN_TILE^3 items
            const int nb2 = (nb + SCAN_TILE - 1) / SCAN_TILE;
            uint32_t *l2 = scan_b.as<uint32_t>() + nb;
            scan_reduce_kernel<<<nb2, SCAN_THREADS, 0, stream>>>(scan_a.as<uint32_t>(), nb, l2);
            ++stats.n_launches;
            scan_apply_kernel<<<1, SCAN_THREADS, 0, stream>>>(l2, nb2, nullptr, l2 + nb2, nullptr);
            ++stats.n_launches;
            scan_apply_kernel<<<nb2, SCAN_THREADS, 0, stream>>>(scan_a.as<uint32_t>(), nb, l2 + nb2,
                                                              scan_b.as<uint32_t>(), nullptr);
            ++stats.n_launches;
        }
        CK(cudaGetLastError());
        scan_apply_kernel<<<nb, SCAN_THREADS, 0, stream>>>(in, n, scan_b.as<uint32_t>(), out, total_dev);
        ++stats.n_launches;
        CK(cudaGetLastError());
    }

    // fused scan descriptor (scan.cuh) for n items; which = 0: CSR offsets, 1: new state ids
    FusedScan fused_scan(int n, int which, unsigned long long *total)
    {
        const size_t nb = (size_t)(n + FS_TILE - 1) / FS_TILE + 1;
        DevBuf &sums = which ? fs_sums2 : fs_sums, &offs = which ? fs_off2 : fs_off;
        sums.reserve(nb * 4, 0, false);
        offs.reserve(nb * 4, 0, false);
        FusedScan fs{};
        fs.block_sums = sums.as<uint32_t>(); fs.block_off = offs.as<uint32_t>();
        fs.ticket = fs_ticket.as<unsigned int>() + which; fs.total = total;
        fs.bump_dst = nullptr; fs.bump_src = nullptr;
        return fs;
    }

    // ---------------- composition of one chunk: keys of states [sid0, sid0+Sc) -------------------
    // ---- tcgen05 path: plane digits of the launch + tensor maps ------------------------------------------
    CUtensorMap b_map(int kbytes)
    {
        if (b_maps_ptr != bdig.p || b_maps_ncap != b_ncap) {
            b_maps.clear();
            b_maps_ptr = bdig.p;
            b_maps_ncap = b_ncap;
        }
        for (auto &kv : b_maps)
            if (kv.first == kbytes) return kv.second;
        b_maps.emplace_back(kbytes, make_digit_map(bdig.p, kbytes, b_pitch, b_ncap, split_digits, SP_BN, split_bk));
        return b_maps.back().second;
    }
    void ensure_split_scratch(size_t slots)
    {
        const size_t unit = (size_t)std::max(SP_KPAD, split_bk);      // K is padded to whole pipeline stages
        size_t pitch = unit;
        for (int l = 1; l <= D; ++l) pitch = std::max<size_t>(pitch, ((size_t)n[l] + unit - 1) / unit * unit);
        size_t ncap = (std::max<size_t>(slots, SP_BS) * 4 + SP_BN - 1) / SP_BN * SP_BN;
        if (ncap <= b_ncap && pitch == b_pitch) return;
        ncap = std::max(ncap, (b_ncap + b_ncap / 2 + SP_BN - 1) / SP_BN * SP_BN);
        bdig.reserve((size_t)split_digits * ncap * pitch, 0, false);
        bscale.reserve(ncap * 8, 0, false);
        b_ncap = ncap;
        b_pitch = pitch;
    }
    template <int SD>
    bool launch_split(const SplitWeights &w, int M, int K, const double *Bsrc, long long bstride, int bit0, double *out,
                      const double *bias_, const uint32_t *keys0, int Sc, int accumulate, const int *perm_, int chain,
                      int n_chain, cudaStream_t cs, int alt_from_slot, const double *alt_src)
    {
        const int tiles = (Sc + SP_BS - 1) / SP_BS;
        const int mine = (tiles - chain + n_chain - 1) / n_chain;
        if (mine <= 0) return false;
        if (!w.ready) throw CudaFail{"weight digits were not prepared"};
        SliceArgs sa{};
        sa.src = Bsrc; sa.stride = bstride; sa.keys = keys0; sa.kw = kw; sa.bit0 = bit0; sa.K = K; sa.Kpad = w.Kpad;
        sa.perm = perm_; sa.S = Sc; sa.dig = bdig.as<signed char>(); sa.pitch = (long long)b_pitch;
        sa.slice_stride = (long long)(b_ncap * b_pitch); sa.scale = bscale.as<double>();
        sa.tile_stride = n_chain; sa.tile_offset = chain;
        sa.alt_from_slot = alt_from_slot; sa.alt_src = alt_src; sa.parent = parent.as<int>();
        sa.lb = (int)lazy_lb; sa.prev_lb = (int)lazy_prev_lb;
        sa.rows_by_slot = cur_rows_by_slot; sa.prev_slot_of = cur_prev_slot_of;
        const bool t = timing_on() && n_chain == 1;      // per-kernel events (the roofline of the dominant kernel)
        size_t e0 = 0;
        if (t) e0 = span_begin();
        const unsigned sgrid = (unsigned)std::min((Sc + 3) / 4, num_sms * 3);     // persistent warps, 3 CTAs per SM
        if (w.Kpad <= 256)
            launch_k(slice_rows_reg_kernel<SD, 2>, dim3(sgrid), dim3(128), 0, cs, sa);
        else if (w.Kpad <= 512)
            launch_k(slice_rows_reg_kernel<SD, 4>, dim3(sgrid), dim3(128), 0, cs, sa);
        else
            launch_k(slice_rows_kernel<SD>, dim3((unsigned)((Sc + 7) / 8)), dim3(256), 0, cs, sa);
        ++stats.n_launches;
        if (t) span_end(e0, 5, 0.0);
        SplitArgs g{};
        g.k_steps = (w.Kpad + split_bk - 1) / split_bk; g.M = M; g.m_tiles = w.Mpad / SP_BM; g.S = Sc; g.perm = perm_;
        g.out = out; g.out_stride = 4LL * R; g.bias = bias_; g.scaleA = w.scale.as<double>();
        g.scaleB = bscale.as<double>(); g.accumulate = accumulate; g.tile_stride = n_chain; g.tile_offset = chain;
        g.add_in = accumulate ? nullptr : fused_add_in;
        g.add_identity = accumulate ? 0 : fused_add_identity;
        g.n_tiles = g.m_tiles * mine;
        g.rows_by_slot = cur_rows_by_slot;
        if (t) e0 = span_begin();
        const dim3 ggrid((unsigned)std::min(g.n_tiles, num_sms));
        if (split_bk == 64)       // 64-byte K slabs per stage (two MMA K steps, SWIZZLE_64B boxes); same results
            launch_k(split_gemm_kernel<SD, 1, 16, 64>, ggrid, dim3(split_threads(1)), SplitCfg<SD, 64>::SMEM, cs, w.map, b_map(w.Kpad), g);
        else switch (split_epi) { // epilogue warps per TMEM lane quarter (split.cuh); results do not depend on it
            case 2: launch_k(split_gemm_kernel<SD, 2, 16>, ggrid, dim3(split_threads(2)), SplitCfg<SD>::SMEM, cs, w.map, b_map(w.Kpad), g); break;
            case 4: launch_k(split_gemm_kernel<SD, 4, 8>, ggrid, dim3(split_threads(4)), SplitCfg<SD>::SMEM, cs, w.map, b_map(w.Kpad), g); break;
            default: launch_k(split_gemm_kernel<SD, 1, 16>, ggrid, dim3(split_threads(1)), SplitCfg<SD>::SMEM, cs, w.map, b_map(w.Kpad), g); break;
        }
        if (t) span_end(e0, 4, 2.0 * M * (double)K * 4.0 * Sc);
        return true;
    }

    void launch_gemm(const double *Wt_, int Mpad_, int M, int K, const double *Bsrc, long long bstride, int bit0,
                     double *out, const double *bias_, const uint32_t *keys0, int Sc, int accumulate, const int *perm_,
                     int chain = 0, int n_chain = 1, const SplitWeights *sw = nullptr, int alt_from_slot = 0x7fffffff,
                     const double *alt_src = nullptr)
    {
        GemmArgs g{};
        g.Wt = Wt_; g.Mpad = Mpad_; g.M = M; g.K = K;
        g.Bsrc = Bsrc; g.b_stride = bstride;
        g.keys = keys0; g.kw = kw; g.bit0 = bit0;
        g.out = out; g.out_stride = 4LL * R;
        g.bias = bias_; g.S = Sc; g.accumulate = accumulate;
        g.perm = perm_; g.m_tiles = Mpad_ / GM_BM;
        g.tile_stride = n_chain; g.tile_offset = chain;
        g.rows_by_slot = cur_rows_by_slot;
        cudaStream_t cs = (n_chain > 1) ? chain_stream[chain] : stream;
        bool launched = false;
        auto go = [&](auto cfg) {
            using C = decltype(cfg);
            const int tiles = (Sc + C::BS - 1) / C::BS;
            const int mine = (tiles - chain + n_chain - 1) / n_chain;     // tiles chain, chain + n_chain, ...
            if (mine <= 0) return;
            dim3 grid((unsigned)(g.m_tiles * mine));
            compose_gemm_kernel<C><<<grid, C::THREADS, C::smem_bytes(K), cs>>>(g);
            launched = true;
        };
        switch (gemm_variant) {   // AM_B200_GEMM_VARIANT: see DESIGN.md section 4
            case 1: go(GemmWide{}); break;
            case 2:
                if (sw == nullptr) throw CudaFail{"split GEMM without weight digits"};
                switch (split_digits) {
                    case 6: launched = launch_split<6>(*sw, M, K, Bsrc, bstride, bit0, out, bias_, keys0, Sc, accumulate, perm_, chain, n_chain, cs, alt_from_slot, alt_src); break;
                    case 8: launched = launch_split<8>(*sw, M, K, Bsrc, bstride, bit0, out, bias_, keys0, Sc, accumulate, perm_, chain, n_chain, cs, alt_from_slot, alt_src); break;
                    default: launched = launch_split<7>(*sw, M, K, Bsrc, bstride, bit0, out, bias_, keys0, Sc, accumulate, perm_, chain, n_chain, cs, alt_from_slot, alt_src); break;
                }
                break;
            default: go(GemmDefault{}); break;
        }
        if (launched) ++stats.n_launches;
        CK(cudaGetLastError());
        if (chain == 0) {
            const double flops = 2.0 * M * (double)K * 4.0 * Sc;
            stats.compose_flops += flops;
            pending_gemm_flops += flops;
        }
    }

    // prm / npre: optional bucket-sorted permutation and, per fc layer h, the number of leading
    // permutation slots whose states need layer h+1 recomputed (incremental mode); null = all states
    // n_equ / equ_idx: states whose level plane is needed (all, or the owned ones in sharded mode)
    // tail: optional per-chain continuation (the clip launch of the chain's states), run on the chain's stream
    // right after the chain's level-plane kernel, before the chains join
    void compose_chunk(const uint32_t *keys0, int S_all, double iso, double *base, const int *prm, const int *npre,
                       int n_equ, const int *equ_idx,
                       const std::function<void(int, int, cudaStream_t)> *tail = nullptr)
    {
        // The layer launches of one chunk form a dependency chain (layer h+1 reads layer h of the same
        // states).  The state tiles are dealt round-robin to n_chain independent chains on separate
        // streams, so the tail of one chain's launch is filled by the others' CTAs.
        const int n_chain = planned_chains();
        last_n_chain = n_chain;
        if (gemm_variant == 2) ensure_split_scratch((size_t)S_all);
        const bool timed = timing_on();
        size_t span0 = 0;
        pending_gemm_flops = 0.0;
        if (timed) span0 = span_begin();
        if (n_chain > 1) {
            CK(cudaEventRecord(fork_event, stream));
            for (int c = 0; c < n_chain; ++c) CK(cudaStreamWaitEvent(chain_stream[c], fork_event, 0));
        }
        const int tile_states = (gemm_variant == 1) ? GemmWide::BS : GemmDefault::BS;
        for (int h = 1; h < D; ++h) {   // fc layer h: hidden h -> hidden h+1
            long long bstride;
            const double *Bsrc = layer_rows(base, h, &bstride);
            double *out = base + 4LL * (off[h + 1] - n1);
            const int Sc = npre ? npre[h] : S_all;
            if (Sc == 0) continue;                 // nobody needs this layer recomputed
            // tcgen05 path: the first skip from the raw input is an addend of the GEMM epilogue (no extra pass)
            const Skip *fused = nullptr;
            fused_add_in = nullptr;
            fused_add_identity = 0;
            if (gemm_variant == 2 && !skips[h].empty() && skips[h][0].src == 0) {   // first in the list: order kept
                const Skip &sk = skips[h][0];
                fused = &sk;
                if (tm_h[sk.tm] == 0 && tm_w[sk.tm] == 0) fused_add_identity = 1;
                else fused_add_in = TM[sk.tm].as<double>();
            }
            for (int c = 0; c < n_chain; ++c) {
                cudaStream_t cs = (n_chain > 1) ? chain_stream[c] : stream;
                // read-through: the permutation slots [npre[h-1], npre[h]) flipped a neuron of layer h itself, the
                // rows of layer h are their parent's
                int alt_from = 0x7fffffff;
                const double *alt_src = nullptr;
                if (lazy_prev != nullptr && npre != nullptr && h >= 2) {
                    alt_from = npre[h - 1];
                    alt_src = lazy_prev + 4LL * (off[h] - n1);
                }
                launch_gemm(Wt[h].as<double>(), Mpad[h], n[h + 1], n[h], Bsrc, bstride, off[h], out,
                            bias[h].as<double>(), keys0, Sc, 0, prm, c, n_chain, gemm_variant == 2 ? &splitW[h] : nullptr,
                            alt_from, alt_src);
                for (const Skip &sk : skips[h]) {
                    if (&sk == fused) continue;
                    const bool identity = (tm_h[sk.tm] == 0 && tm_w[sk.tm] == 0);
                    const int M = n[h + 1];
                    if (sk.src == 0) {
                        const long long tot = (long long)Sc * M;
                        skip_input_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, cs>>>(
                            out, 4LL * R, M, Sc, identity ? nullptr : TM[sk.tm].as<double>(), prm, tile_states, n_chain, c,
                            cur_rows_by_slot);
                        ++stats.n_launches;
                    } else {
                        long long sstride;
                        const double *src = layer_rows(base, sk.src, &sstride);
                        if (identity) {
                            const long long tot = (long long)Sc * M * 4;
                            skip_hidden_identity_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, cs>>>(
                                out, 4LL * R, src, sstride, keys0, kw, off[sk.src], M, Sc, prm, tile_states, n_chain, c,
                                cur_rows_by_slot);
                            ++stats.n_launches;
                        } else {
                            launch_gemm(TMt[sk.tm].as<double>(), tm_Mpad[sk.tm], M, n[sk.src], src, sstride, off[sk.src],
                                        out, nullptr, keys0, Sc, 1, prm, c, n_chain,
                                        gemm_variant == 2 ? &splitTM[sk.tm] : nullptr);
                        }
                    }
                    CK(cudaGetLastError());
                }
            }
        }
        auto launch_equ = [&](int c, int nc, cudaStream_t cs) {
            EquArgs e{};
            e.w = wout.as<double>();
            e.bias = bout;
            e.iso = iso;
            e.in = layer_rows(base, D, &e.in_stride);
            const int Sc = n_equ;
            if (Sc <= 0) return;
            e.idx = equ_idx;
            e.keys = keys0; e.kw = kw; e.bit0 = off[D]; e.K = n[D]; e.S = Sc;
            e.equ = equ.as<double>();
            e.bucket = nullptr;
            e.tile_stride = nc; e.tile_offset = c; e.tile = chain_tile();
            e.rows_by_slot = cur_rows_by_slot; e.prev_slot_of = cur_prev_slot_of;
            if (lazy_prev != nullptr && D >= 2) {
                e.bucket = bucket.as<int>(); e.parent = parent.as<int>();
                e.lb = (int)lazy_lb; e.prev_lb = (int)lazy_prev_lb; e.D = D;
                e.alt_in = lazy_prev + 4LL * (off[D] - n1);
            }
            e.n_skips = 0;
            for (const Skip &sk : skips[D]) {
                if (e.n_skips == EQU_MAX_SKIPS) throw CudaFail{"more than 4 skips into the output layer"};
                EquSkip &q = e.skips[e.n_skips++];
                const bool identity = (tm_h[sk.tm] == 0 && tm_w[sk.tm] == 0);
                q.T = identity ? nullptr : TM[sk.tm].as<double>();
                q.src = nullptr; q.src_stride = 0; q.src_bit0 = 0; q.src_n = 0;
                if (sk.src == 0) {
                    q.kind = identity ? 1 : 2;
                } else {
                    q.kind = identity ? 3 : 4;
                    q.src = layer_rows(base, sk.src, &q.src_stride);
                    q.src_bit0 = off[sk.src];
                    q.src_n = n[sk.src];
                }
            }
            const int mine = chain_slots(Sc, c, nc);
            if (mine <= 0) return;
            if (gemm_variant == 2 && e.n_skips == 0 && equ_warp)      // warp per state, coalesced rows (compose.cuh)
                launch_k(equ_warp_kernel, dim3((mine + 7) / 8), dim3(256), 0, cs, e);
            else
                launch_k(equ_kernel, dim3((mine * 4 + 127) / 128), dim3(128), 0, cs, e);
            ++stats.n_launches;
            CK(cudaGetLastError());
        };
        const bool chained_tail = (tail != nullptr) && n_chain > 1;
        if (chained_tail)
            for (int c = 0; c < n_chain; ++c) {
                launch_equ(c, n_chain, chain_stream[c]);
                (*tail)(c, n_chain, chain_stream[c]);
            }
        if (n_chain > 1) {
            for (int c = 0; c < n_chain; ++c) {
                CK(cudaEventRecord(join_event[c], chain_stream[c]));
                CK(cudaStreamWaitEvent(stream, join_event[c], 0));
            }
        }
        if (timed) span_end(span0, 0, pending_gemm_flops);
        if (!chained_tail) {
            launch_equ(0, 1, stream);
            if (tail != nullptr) (*tail)(0, 1, stream);
        }
    }

    // slots of a list of n handled by chain c of nc: whole tiles of SP_BS slots dealt round-robin (the same deal as
    // the composition launches, so a chain only depends on its own kernels)
    int planned_chains() const { return (D >= 3) ? (n_chains > 0 ? n_chains : (shard_world > 1 ? 4 : 1)) : 1; }
    int chain_tile() const { return gemm_variant == 2 ? SP_BS : (gemm_variant == 1 ? GemmWide::BS : GemmDefault::BS); }
    int chain_slots(int n, int c, int nc) const
    {
        if (nc <= 1) return n;
        const int ct = chain_tile();
        const int tiles = (n + ct - 1) / ct;
        const int mine = (tiles - c + nc - 1) / nc;
        return mine > 0 ? mine * ct : 0;
    }

    size_t chunk_states() const
    {
        double gib = 8.0;
        if (const char *s = getenv("AM_B200_PLANE_GIB")) gib = atof(s);
        const size_t per = std::max<size_t>((size_t)R * 32, 32);
        size_t c = (size_t)(gib * (1ull << 30)) / per;
        c = std::max<size_t>(c, 1024);
        c = std::min<size_t>(c, 1u << 22);
        return (c / 32) * 32;
    }

    void ensure_chunk_scratch(size_t Sc)
    {
        planes.reserve(std::max<size_t>((size_t)R * 32 * Sc, 64), 0, false);
        ensure_chunk_scratch_no_planes(Sc);
    }
    void ensure_chunk_scratch_no_planes(size_t Sc)
    {
        equ.reserve(Sc * 32, 0, false);
        f_cnt.reserve(Sc * 4, 0, false);
        f_off.reserve(Sc * 4, 0, false);
        f_edges.reserve(Sc * VSLOTS * 4, 0, false);
        f_verts.reserve(Sc * VSLOTS * 24, 0, false);
    }

    void read_counters()
    {
        CK(cudaMemcpyAsync(h_counters, counters.p, CNT_NUM * 8, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
    }

    // in-place integer sum over all ranks, ordered on the engine's stream
    void allreduce_i32(void *ptr, long long n32)
    {
        if (nccl_comm) {
            const int rc = NcclApi::get().allreduce(ptr, ptr, (size_t)n32, /*ncclInt32*/ 2, /*ncclSum*/ 0, nccl_comm, stream);
            if (rc != 0) throw CudaFail{"ncclAllReduce failed (" + std::to_string(rc) + ")"};
        } else if (shard_cb) {
            if (shard_cb(shard_user, ptr, n32, (void *)stream) != 0) throw CudaFail{"the host all-reduce callback failed"};
        } else {
            throw CudaFail{"sharded mode without a collective"};
        }
    }

    template <typename F>
    void dispatch_group(F &&f)
    {
        switch (G) {
            case 1: f(std::integral_constant<int, 1>{}); break;
            case 2: f(std::integral_constant<int, 2>{}); break;
            case 4: f(std::integral_constant<int, 4>{}); break;
            case 8: f(std::integral_constant<int, 8>{}); break;
            case 16: f(std::integral_constant<int, 16>{}); break;
            default: f(std::integral_constant<int, 32>{}); break;
        }
    }
};

// =================================================================================================
namespace {

enum PtrKind { PK_HOST, PK_DEVICE };
PtrKind classify(const void *p)
{
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return PK_HOST;
    }
    return (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) ? PK_DEVICE : PK_HOST;
}

// fetch `count` reals (float or double, host or device) as doubles on the host
std::vector<double> fetch_real(const void *p, size_t count, bool f64)
{
    std::vector<double> out(count);
    if (count == 0) return out;
    if (p == nullptr) throw CudaFail{"null data pointer"};
    const size_t es = f64 ? 8 : 4;
    std::vector<unsigned char> tmp;
    const void *src = p;
    if (classify(p) == PK_DEVICE) {
        tmp.resize(count * es);
        CK(cudaMemcpy(tmp.data(), p, count * es, cudaMemcpyDeviceToHost));
        src = tmp.data();
    }
    if (f64) memcpy(out.data(), src, count * 8);
    else for (size_t i = 0; i < count; ++i) out[i] = (double)static_cast<const float *>(src)[i];
    return out;
}

void upload(DevBuf &b, const void *src, size_t bytes, cudaStream_t st)
{
    b.reserve(std::max<size_t>(bytes, 16), 0, false);
    if (bytes) CK(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, st));
}

void prepare_split_weights(am_handle *h, SplitWeights &sw, const double *w, int M, int K)
{
    const int unit = std::max(SP_KPAD, h->split_bk);      // K is padded to whole pipeline stages
    sw.Kpad = (K + unit - 1) / unit * unit;
    sw.Mpad = (M + SP_BM - 1) / SP_BM * SP_BM;
    std::vector<signed char> dig;
    std::vector<double> scale;
    split_weight_digits(w, M, K, sw.Mpad, sw.Kpad, h->split_digits, dig, scale);
    upload(sw.dig, dig.data(), dig.size(), h->stream);
    upload(sw.scale, scale.data(), scale.size() * 8, h->stream);
    CK(cudaStreamSynchronize(h->stream));
    sw.map = make_digit_map(sw.dig.p, sw.Kpad, (size_t)sw.Kpad, (size_t)sw.Mpad, h->split_digits, SP_BM, h->split_bk);
    sw.ready = true;
}

// Raw bytes of one network tensor (host or device) compared with what the previous call saw: a layer whose
// bytes did not change keeps its device copies (k-major FP64 matrix, int8 digit planes, tensor map).  A batch
// of latent-conditioned shapes (BASELINE config 5) differs only in biases[0]; re-marching the same network
// uploads nothing.  (Reference: weights are re-bound on every call, backend/src/cuam.cpp:114-137.)
// Device-resident tensors are compared by a 64-bit positional checksum computed on the device (digest_words_kernel,
// all tensors of the call in one batch, one 8-byte-per-tensor read-back) instead of being copied to the host on every
// march: 14.7 MB of D2H + memcmp per call at 8x512, which is 1 % of a march that takes 0.45 s on 8 GPUs.
void checksum_device_tensors(am_handle *h, const std::vector<std::pair<const void *, size_t>> &tensors)
{
    const size_t n = tensors.size();
    h->new_sums.assign(n, 0);
    h->new_sum_valid.assign(n, 0);
    h->sum_dev.reserve(std::max<size_t>(n * 8, 64));
    CK(cudaMemsetAsync(h->sum_dev.p, 0, n * 8, h->stream));
    bool any = false;
    for (size_t i = 0; i < n; ++i) {
        const void *p = tensors[i].first;
        const size_t count = tensors[i].second;
        if (!p || count == 0 || classify(p) != PK_DEVICE) continue;
        digest_words_kernel<<<(unsigned)std::min<size_t>((count + 255) / 256, (size_t)h->num_sms * 4), 256, 0, h->stream>>>(
            p, (long long)count, h->f64 ? 8 : 4, 0x5EED0000ull + i, h->sum_dev.as<unsigned long long>() + i);
        h->new_sum_valid[i] = 1;
        any = true;
    }
    if (!any) return;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h->new_sums.data(), h->sum_dev.p, n * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
}

bool tensor_changed(am_handle *h, size_t slot, const void *p, size_t count)
{
    const size_t bytes = count * (h->f64 ? 8 : 4);
    if (h->raw_cache.size() <= slot) h->raw_cache.resize(slot + 1);
    if (h->old_sums.size() <= slot) { h->old_sums.resize(slot + 1, 0); h->old_sum_valid.resize(slot + 1, 0); }
    std::vector<unsigned char> &c = h->raw_cache[slot];
    if (bytes && p == nullptr) throw CudaFail{"null data pointer"};
    const bool have_sum = slot < h->new_sum_valid.size() && h->new_sum_valid[slot];
    if (have_sum && h->weights_loaded && h->old_sum_valid[slot] && h->old_sums[slot] == h->new_sums[slot] && c.size() == bytes)
        return false;                                      // same bytes as last time (device-side checksum)
    h->old_sum_valid[slot] = have_sum ? 1 : 0;
    if (have_sum) h->old_sums[slot] = h->new_sums[slot];
    const unsigned char *src = static_cast<const unsigned char *>(p);
    std::vector<unsigned char> tmp;
    if (bytes && classify(p) == PK_DEVICE) {
        tmp.resize(bytes);
        CK(cudaMemcpy(tmp.data(), p, bytes, cudaMemcpyDeviceToHost));
        src = tmp.data();
    }
    if (h->weights_loaded && c.size() == bytes && (bytes == 0 || memcmp(c.data(), src, bytes) == 0)) return false;
    if (!tmp.empty()) c.swap(tmp);
    else c.assign(src, src + bytes);
    return true;
}

std::vector<double> cached_real(const am_handle *h, size_t slot, size_t count)
{
    std::vector<double> out(count);
    const std::vector<unsigned char> &c = h->raw_cache[slot];
    if (h->f64) memcpy(out.data(), c.data(), count * 8);
    else for (size_t i = 0; i < count; ++i) out[i] = (double)reinterpret_cast<const float *>(c.data())[i];
    return out;
}

int load_weights(am_handle *h, const void *const *W, const void *const *B, const void *const *TMp, const int *tm_shapes,
                 int n_tm)
{
    const int D = h->D;
    if (n_tm < h->n_tm) throw CudaFail{"arc_table references transform " + std::to_string(h->n_tm - 1) +
                                       " but only " + std::to_string(n_tm) + " were given"};
    if (h->weights_loaded && (int)h->tm_h.size() != n_tm) h->weights_loaded = false;
    for (int t = 0; t < n_tm && h->weights_loaded; ++t)
        if (h->tm_h[t] != tm_shapes[2 * t] || h->tm_w[t] != tm_shapes[2 * t + 1]) h->weights_loaded = false;
    // cache slots: W[l] -> 2 l, B[l] -> 2 l + 1 (l = 0..D), transform t -> 2 (D + 1) + t
    h->stats_layers_reloaded = 0;
    {
        std::vector<std::pair<const void *, size_t>> tl((size_t)2 * (D + 1) + n_tm, {nullptr, 0});
        for (int l = 0; l <= D; ++l) {
            tl[2 * l] = {W[l], (size_t)h->n[l + 1] * h->n[l]};
            tl[2 * l + 1] = {B[l], (size_t)h->n[l + 1]};
        }
        for (int t = 0; t < n_tm; ++t) tl[2 * (size_t)(D + 1) + t] = {TMp[t], (size_t)tm_shapes[2 * t] * tm_shapes[2 * t + 1]};
        checksum_device_tensors(h, tl);
    }
    // layer 0 -> shared table of hidden layer 1 rows
    {
        const bool cw = tensor_changed(h, 0, W[0], (size_t)h->n1 * 3), cb = tensor_changed(h, 1, B[0], (size_t)h->n1);
        if (cw || cb) {
            auto w0 = cached_real(h, 0, (size_t)h->n1 * 3);
            auto b0 = cached_real(h, 1, (size_t)h->n1);
            std::vector<double> p1((size_t)h->n1 * 4);
            for (int r = 0; r < h->n1; ++r) {
                p1[4 * r + 0] = w0[3 * r + 0]; p1[4 * r + 1] = w0[3 * r + 1]; p1[4 * r + 2] = w0[3 * r + 2];
                p1[4 * r + 3] = b0[r];
            }
            upload(h->P1, p1.data(), p1.size() * 8, h->stream);
            upload(h->Wrow[0], w0.data(), w0.size() * 8, h->stream);
            upload(h->Brow[0], b0.data(), b0.size() * 8, h->stream);
            CK(cudaStreamSynchronize(h->stream));
            ++h->stats_layers_reloaded;
        }
    }
    for (int l = 1; l < D; ++l) {
        const int K = h->n[l], M = h->n[l + 1];
        const int Kp = (K + GM_KPAD - 1) / GM_KPAD * GM_KPAD, Mp = (M + GM_BM - 1) / GM_BM * GM_BM;
        h->Kpad[l] = Kp; h->Mpad[l] = Mp;
        const bool cw = tensor_changed(h, 2 * l, W[l], (size_t)M * K), cb = tensor_changed(h, 2 * l + 1, B[l], (size_t)M);
        if (cb) {
            auto b = cached_real(h, 2 * l + 1, (size_t)M);
            upload(h->bias[l], b.data(), b.size() * 8, h->stream);
            upload(h->Brow[l], b.data(), b.size() * 8, h->stream);
            CK(cudaStreamSynchronize(h->stream));
        }
        if (!cw) continue;
        ++h->stats_layers_reloaded;
        auto w = cached_real(h, 2 * l, (size_t)M * K);
        upload(h->Wrow[l], w.data(), w.size() * 8, h->stream);
        if (h->gemm_variant != 2) {      // k-major FP64 copy: the DMMA path only
            std::vector<double> wt((size_t)Kp * Mp, 0.0);
            for (int m = 0; m < M; ++m)
                for (int k = 0; k < K; ++k) wt[(size_t)k * Mp + m] = w[(size_t)m * K + k];
            upload(h->Wt[l], wt.data(), wt.size() * 8, h->stream);
            CK(cudaStreamSynchronize(h->stream));
        } else {
            prepare_split_weights(h, h->splitW[l], w.data(), M, K);
        }
    }
    {
        const bool cw = tensor_changed(h, 2 * D, W[D], (size_t)h->n[D]), cb = tensor_changed(h, 2 * D + 1, B[D], 1);
        if (cw) {
            auto w = cached_real(h, 2 * D, (size_t)h->n[D]);
            upload(h->wout, w.data(), w.size() * 8, h->stream);
            upload(h->Wrow[D], w.data(), w.size() * 8, h->stream);
            CK(cudaStreamSynchronize(h->stream));
        }
        if (cw || cb) {
            auto b = cached_real(h, 2 * D + 1, 1);
            h->bout = b[0];
            upload(h->Brow[D], b.data(), 8, h->stream);
            CK(cudaStreamSynchronize(h->stream));
        }
    }
    h->TM.resize(n_tm); h->TMt.resize(n_tm);
    h->splitTM.resize(n_tm);
    h->tm_h.resize(n_tm, 0); h->tm_w.resize(n_tm, 0); h->tm_Mpad.resize(n_tm, 0);
    for (int t = 0; t < n_tm; ++t) {
        const int th = tm_shapes[2 * t], tw = tm_shapes[2 * t + 1];
        h->tm_h[t] = th; h->tm_w[t] = tw;
        if (th == 0 && tw == 0) continue;
        if (!tensor_changed(h, 2 * (size_t)(D + 1) + t, TMp[t], (size_t)th * tw)) continue;
        ++h->stats_layers_reloaded;
        auto w = cached_real(h, 2 * (size_t)(D + 1) + t, (size_t)th * tw);
        upload(h->TM[t], w.data(), w.size() * 8, h->stream);
        const int Kp = (tw + GM_KPAD - 1) / GM_KPAD * GM_KPAD, Mp = (th + GM_BM - 1) / GM_BM * GM_BM;
        h->tm_Mpad[t] = Mp;
        if (h->gemm_variant != 2) {
            std::vector<double> wt((size_t)Kp * Mp, 0.0);
            for (int m = 0; m < th; ++m)
                for (int k = 0; k < tw; ++k) wt[(size_t)k * Mp + m] = w[(size_t)m * tw + k];
            upload(h->TMt[t], wt.data(), wt.size() * 8, h->stream);
        }
        CK(cudaStreamSynchronize(h->stream));
        if (h->gemm_variant == 2 && tw >= 1) {
            bool feeds_hidden = false;       // digit planes only for transforms read by a hidden-source skip GEMM
            for (int l = 1; l < D; ++l)
                for (const Skip &sk : h->skips[l]) feeds_hidden |= (sk.tm == t && sk.src >= 1);
            if (feeds_hidden) prepare_split_weights(h, h->splitTM[t], w.data(), th, tw);
        }
    }
    // shape checks of the skips (reference backend/src/cuam.cpp:153-173)
    for (int l = 1; l <= D; ++l)
        for (const Skip &sk : h->skips[l]) {
            const int th = h->tm_h[sk.tm], tw = h->tm_w[sk.tm];
            const int src_w = (sk.src == 0) ? 3 : h->n[sk.src];
            if (th == 0 && tw == 0) {
                if (src_w != h->n[l + 1] && !(sk.src == 0 && h->n[l + 1] <= 3))
                    throw CudaFail{"identity skip between layers of different width"};
            } else if (th != h->n[l + 1] || tw != src_w) {
                throw CudaFail{"transform " + std::to_string(sk.tm) + " has shape (" + std::to_string(th) + "," +
                               std::to_string(tw) + "), expected (" + std::to_string(h->n[l + 1]) + "," +
                               std::to_string(src_w) + ")"};
            }
        }
    h->weights_loaded = true;
    return AM_OK;
}

// states == nullptr: the seeds found by am_seed_dichotomy (packed keys + points, already on the device)
void insert_seeds(am_handle *h, const uint8_t *states, const double *points, long long N)
{
    cudaStream_t st = h->stream;
    h->xkeys.reserve((size_t)N * h->kw * 4, 0, false);
    h->xh.reserve((size_t)N * 8, 0, false);
    h->xslot.reserve((size_t)N * 4, 0, false);
    h->nwin.reserve((size_t)N * 4, 0, false);
    h->wbase.reserve((size_t)N * 4, 0, false);
    if (states == nullptr) {
        h->xpt.reserve((size_t)N * 24, 0, false);
        CK(cudaMemcpyAsync(h->xkeys.p, h->sd_keys.p, (size_t)N * h->kw * 4, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(h->xpt.p, h->sd_mid.p, (size_t)N * 24, cudaMemcpyDeviceToDevice, st));
    } else {
        h->xstates.reserve((size_t)N * h->L, 0, false);
        if (classify(states) == PK_DEVICE)
            CK(cudaMemcpyAsync(h->xstates.p, states, (size_t)N * h->L, cudaMemcpyDeviceToDevice, st));
        else
            CK(cudaMemcpyAsync(h->xstates.p, states, (size_t)N * h->L, cudaMemcpyHostToDevice, st));
        upload(h->xpt, points, (size_t)N * 24, st);
        const long long tot = N * h->kw;
        pack_states_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(h->xstates.as<uint8_t>(), (int)N, h->L, h->kw,
                                                                         h->xkeys.as<uint32_t>());
        ++h->stats.n_launches;
    }
    hash_keys_kernel<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(h->xkeys.as<uint32_t>(), (int)N, h->kw,
                                                                 h->xh.as<unsigned long long>());
    ++h->stats.n_launches;
    CK(cudaGetLastError());
    h->ensure_table((size_t)h->n_states + (size_t)N);
    h->ensure_states((size_t)h->n_states + (size_t)N);
    XArgs x{};
    x.xkeys = h->xkeys.as<uint32_t>(); x.xh = h->xh.as<unsigned long long>(); x.xpt = h->xpt.as<double>();
    x.xparent = nullptr; x.xvia = nullptr;
    x.N = (int)N; x.kw = h->kw; x.kw4 = h->kw4;
    x.keys = h->keys.as<uint32_t>();
    x.table = TableRef{h->table.as<unsigned long long>(), h->tcap - 1};
    x.x_slot = h->xslot.as<int>(); x.xwin = h->nwin.as<uint32_t>(); x.win_base = h->wbase.as<uint32_t>();
    x.keys_w = h->keys.as<uint32_t>(); x.hsum_w = h->hsum.as<unsigned long long>();
    x.parent = h->parent.as<int>(); x.via_edge = h->via.as<int>(); x.seedpt = h->seedpt.as<double>();
    x.n_states = (int)h->n_states;
    const int G = h->G;
    const unsigned gb = (unsigned)((N * G + 255) / 256);
    h->dispatch_group([&](auto g) { x_insert_kernel<decltype(g)::value><<<gb, 256, 0, st>>>(x); });
    ++h->stats.n_launches;
    x_count_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(x);
    ++h->stats.n_launches;
    CK(cudaGetLastError());
    h->scan(h->nwin.as<uint32_t>(), h->wbase.as<uint32_t>(), (int)N, h->counters.as<unsigned long long>() + CNT_NEW);
    h->dispatch_group([&](auto g) { x_finalize_kernel<decltype(g)::value><<<gb, 256, 0, st>>>(x); });
    ++h->stats.n_launches;
    CK(cudaGetLastError());
    h->read_counters();
    h->n_states += (long long)h->h_counters[CNT_NEW];
    if (h->shard_world > 1 && h->n_states > 0) {
        seed_owner_kernel<<<(unsigned)((h->n_states + 255) / 256), 256, 0, st>>>(h->owner.as<uint8_t>(), (int)h->n_states,
                                                                                  h->shard_world);
        ++h->stats.n_launches;
        CK(cudaGetLastError());
    }
}

// per-level polygon scratch (stride VSLOTS per state)
struct Scratch { int *cnt; int *edges; double *verts; };
Scratch scratch_of(am_handle *h, size_t S)
{
    (void)S;
    return Scratch{h->f_cnt.as<int>(), h->f_edges.as<int>(), h->f_verts.as<double>()};
}

// clip the states idx[0..n) (or all Sc states when idx == nullptr) of the range starting at sid0
void run_clip(am_handle *h, long long sid0, int n, const double *planes_base, int flip, const int *idx, Scratch sc,
              int chain = 0, int n_chain = 1, cudaStream_t cs = nullptr)
{
    if (n <= 0) return;
    cudaStream_t st = cs ? cs : h->stream;
    const int mine = h->chain_slots(n, chain, n_chain);
    if (mine <= 0) return;
    ClipArgs ca{};
    ca.keys = h->keys.as<uint32_t>() + (size_t)sid0 * h->kw; ca.kw = h->kw;
    ca.P1 = h->P1.as<double>(); ca.n1 = h->n1;
    ca.P = planes_base; ca.p_stride = 4LL * h->R;
    ca.equ = h->equ.as<double>();
    ca.extra = h->extra.as<double>();
    ca.L = h->L; ca.E = h->E; ca.S = n; ca.flip = flip;
    ca.seedpt = h->seedpt.as<double>() + (size_t)sid0 * 4;
    ca.idx = idx;
    ca.P_prev = h->lazy_prev;
    ca.P_own = const_cast<double *>(planes_base);
    ca.bucket = h->bucket.as<int>(); ca.parent = h->parent.as<int>();
    ca.lb = (int)h->lazy_lb; ca.prev_lb = (int)h->lazy_prev_lb;
    ca.lo.D = h->D;
    for (int l = 1; l <= h->D + 1; ++l) ca.lo.off[l] = h->off[l];
    ca.out_cnt = sc.cnt; ca.out_edges = sc.edges; ca.out_verts = sc.verts;
    ca.counters = h->counters.as<unsigned long long>();
    ca.tile_stride = n_chain; ca.tile_offset = chain; ca.tile = h->chain_tile();
    ca.rows_by_slot = h->cur_rows_by_slot; ca.prev_slot_of = h->cur_prev_slot_of;
    ca.push_on = 0;
    if (h->p2p && h->shard_world > 1 && h->push_in_clip) {      // the clip warp pushes its polygon to every rank itself
        ca.push_on = 1;
        for (int q = 0; q < h->shard_world; ++q) ca.push.base[q] = h->xpeers.base[q];
        ca.push.world = h->shard_world; ca.push.rank = h->shard_rank;
        ca.push.region_off = h->xlay.region(h->shard_rank) + h->xlay.edge_off();
        ca.push.xyz_off = h->xlay.xyz_off() - h->xlay.edge_off();
        ca.push.cnt_base = h->xlay.cnt_base; ca.push.where_base = h->xlay.where_base;
        ca.push.cap_corners = h->xlay.cap_corners; ca.push.cursor = h->xcursor.as<int>();
    }
    const unsigned cgrid = (unsigned)((mine + CLIP_WARPS - 1) / CLIP_WARPS);
    switch (h->clip_minb) {   // AM_B200_CLIP_MINB: 2 = two CTAs/SM, no spills (default); 3 = three CTAs/SM
        case 3: launch_k(clip_kernel<3, 2, 3>, dim3(cgrid), dim3(CLIP_WARPS * 32), clip_ring_bytes(2, 3), st, ca); break;
        case 4: launch_k(clip_kernel<2, 2, 3, true>, dim3(cgrid), dim3(CLIP_WARPS * 32), clip_ring_bytes(2, 3), st, ca); break;
        case 6: launch_k(clip_kernel<2, 2, 5>, dim3(cgrid), dim3(CLIP_WARPS * 32), clip_ring_bytes(2, 5), st, ca); break;
        case 8: launch_k(clip_kernel<4, 1, 4>, dim3(cgrid), dim3(CLIP_WARPS * 32), clip_ring_bytes(1, 4), st, ca); break;
        case 9: launch_k(clip_kernel<3, 1, 6>, dim3(cgrid), dim3(CLIP_WARPS * 32), clip_ring_bytes(1, 6), st, ca); break;
        case 2: launch_k(clip_kernel<2, 2, 3>, dim3(cgrid), dim3(CLIP_WARPS * 32), clip_ring_bytes(2, 3), st, ca); break;
        case 5: launch_k(clip_kernel<2, 2, 4>, dim3(cgrid), dim3(CLIP_WARPS * 32), clip_ring_bytes(2, 4), st, ca); break;
        default: launch_k(clip_kernel<3, 1, 4>, dim3(cgrid), dim3(CLIP_WARPS * 32), clip_ring_bytes(1, 4), st, ca); break;
    }
    ++h->stats.n_launches;
    CK(cudaGetLastError());
}

// scan + compact the polygons of the states [sid0, sid0+Sc) into the global CSR.
// Sharded mode: the polygon sizes are summed over the ranks first (a rank's count is zero for states it
// does not own), every rank then knows the level's CSR offsets, writes only its own polygons into the
// zeroed CSR range, and the range is summed over the ranks: 4 + 28 B per corner cross NVLink instead of
// the fixed-stride scratch.
void store_faces(am_handle *h, long long sid0, int Sc, Scratch sc, bool fused = false)
{
    cudaStream_t st = h->stream;
    unsigned long long *cnt = h->counters.as<unsigned long long>();
    const bool sharded = h->shard_world > 1;
    h->f_off.reserve((size_t)Sc * 4, 0, false);
    if (fused && !sharded) {
        // one launch scans the polygon sizes (block-local prefix + block offsets), one copies the polygons; the
        // running corner total is advanced by the level's winner kernel (process_level)
        FusedScan fs = h->fused_scan(Sc, 0, cnt + CNT_CHUNK_CORNERS);
        launch_k(scan_local_kernel, dim3((Sc + FS_TILE - 1) / FS_TILE), dim3(FS_THREADS), 0, st, reinterpret_cast<uint32_t *>(sc.cnt), Sc,
                                                                             h->f_off.as<uint32_t>(), fs);
        ++h->stats.n_launches;
        CompactArgs co{};
        co.owner = nullptr; co.rank = 0;
        co.cnt = sc.cnt; co.off = h->f_off.as<uint32_t>(); co.block_off = fs.block_off;
        co.edges = sc.edges; co.verts = sc.verts;
        co.S = Sc; co.sid0 = (int)sid0;
        co.face_off = h->face_off.as<long long>(); co.face_edges = h->face_edges.as<int>();
        co.face_xyz = h->face_xyz.as<double>(); co.counters = cnt;
        launch_k(compact_faces_kernel, dim3((Sc + 7) / 8), dim3(256), 0, st, co);
        ++h->stats.n_launches;
        CK(cudaGetLastError());
        return;
    }
    if (sharded) h->allreduce_i32(sc.cnt, Sc);
    h->scan(reinterpret_cast<uint32_t *>(sc.cnt), h->f_off.as<uint32_t>(), Sc, cnt + CNT_CHUNK_CORNERS);
    long long base = 0, n_lvl = 0;
    if (sharded) {
        CK(cudaMemcpyAsync(h->h_counters + CNT_CHUNK_CORNERS, cnt + CNT_CHUNK_CORNERS, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        base = (long long)h->h_counters[CNT_CORNERS];          // exact: read back at the end of the last level
        n_lvl = (long long)h->h_counters[CNT_CHUNK_CORNERS];
        if (n_lvl > 0) {
            CK(cudaMemsetAsync(h->face_edges.as<int>() + base, 0, (size_t)n_lvl * 4, st));
            CK(cudaMemsetAsync(h->face_xyz.as<double>() + base * 3, 0, (size_t)n_lvl * 24, st));
        }
    }
    CompactArgs co{};
    co.owner = sharded ? h->owner.as<uint8_t>() : nullptr;
    co.rank = h->shard_rank;
    co.cnt = sc.cnt; co.off = h->f_off.as<uint32_t>(); co.block_off = nullptr;
    co.edges = sc.edges; co.verts = sc.verts;
    co.S = Sc; co.sid0 = (int)sid0;
    co.face_off = h->face_off.as<long long>(); co.face_edges = h->face_edges.as<int>();
    co.face_xyz = h->face_xyz.as<double>(); co.counters = cnt;
    launch_k(compact_faces_kernel, dim3((Sc + 7) / 8), dim3(256), 0, st, co);
    ++h->stats.n_launches;
    if (sharded && n_lvl > 0) {
        h->allreduce_i32(h->face_edges.as<int>() + base, n_lvl);
        h->allreduce_i32(h->face_xyz.as<double>() + base * 3, n_lvl * 6);
    }
    bump_counters_kernel<<<1, 32, 0, st>>>(cnt);
    ++h->stats.n_launches;
    CK(cudaGetLastError());
}

// The sharded counterpart of store_faces: this rank's polygons are pushed into every rank's inbox over NVLink,
// one device-side barrier, then sizes -> prefix sum -> CSR from the own inbox (xchg.cuh).  No host
// synchronisation, no collective library.
void xchg_barrier(am_handle *h)
{
    ++h->xepoch;
    const bool t = h->timing_on();
    size_t e0 = 0;
    if (t) e0 = h->span_begin();
    launch_k(xchg_barrier_kernel, dim3(1), dim3(32), 0, h->stream, h->xpeers, h->xepoch, h->xtimeout_ns, h->counters.as<unsigned long long>());
    if (t) h->span_end(e0, 6);
    ++h->stats.n_launches;
    CK(cudaGetLastError());
}

void store_faces_p2p(am_handle *h, long long sid0, int Sc, Scratch sc, const int *idx, int n_mine)
{
    cudaStream_t st = h->stream;
    unsigned long long *cnt = h->counters.as<unsigned long long>();
    if (Sc > h->xlay.mask_cap)
        throw CapacityFail{"sharded march: a BFS level of " + std::to_string(Sc) + " states exceeds the exchange block "
                           "(raise AM_B200_XCHG_LEVEL_STATES)"};
    h->f_off.reserve((size_t)Sc * 4, 0, false);
    const bool tm = h->timing_on();
    size_t e0 = 0;
    if (n_mine > 0 && !h->push_in_clip) {
        XchgPackArgs pa{};      // the cursor is zero: cleared by the previous level's winner kernel
        pa.idx = idx; pa.n = n_mine; pa.cnt = sc.cnt; pa.edges = sc.edges; pa.verts = sc.verts;
        pa.cursor = h->xcursor.as<int>(); pa.p = h->xpeers; pa.lay = h->xlay; pa.counters = cnt;
        if (tm) e0 = h->span_begin();
        launch_k(xchg_pack_kernel, dim3((unsigned)((n_mine + 7) / 8)), dim3(256), 0, st, pa);
        if (tm) h->span_end(e0, 7);
        ++h->stats.n_launches;
    }
    xchg_barrier(h);
    if (tm) e0 = h->span_begin();
    const unsigned char *own = h->xpeers.base[h->shard_rank];
    const uint32_t *cnt_all = reinterpret_cast<const uint32_t *>(own + h->xlay.cnt_base);
    FusedScan fs = h->fused_scan(Sc, 0, cnt + CNT_CHUNK_CORNERS);
    launch_k(scan_local_kernel, dim3((Sc + FS_TILE - 1) / FS_TILE), dim3(FS_THREADS), 0, st, cnt_all, Sc, h->f_off.as<uint32_t>(), fs);
    ++h->stats.n_launches;
    XchgCompactArgs co{};
    co.own = own; co.lay = h->xlay; co.cnt_all = reinterpret_cast<const int *>(cnt_all); co.off = h->f_off.as<uint32_t>();
    co.block_off = fs.block_off;
    co.where = reinterpret_cast<const int2 *>(own + h->xlay.where_base); co.S = Sc; co.sid0 = (int)sid0;
    co.face_off = h->face_off.as<long long>(); co.face_edges = h->face_edges.as<int>();
    co.face_xyz = h->face_xyz.as<double>(); co.counters = cnt;
    launch_k(xchg_compact_kernel, dim3((Sc + 7) / 8), dim3(256), 0, st, co);
    if (tm) h->span_end(e0, 8);
    ++h->stats.n_launches;
    CK(cudaGetLastError());
}

void process_level(am_handle *h, long long lb, long long le, double iso, int flip)
{
    cudaStream_t st = h->stream;
    const long long S = le - lb;
    g_pdl_now = pdl_mode() == 2 || (pdl_mode() == 1 && S / std::max(1, h->shard_world) < h->pdl_below);
    // candidate index = (state within the level << 5) | edge slot under the CAND_TAG bit (frontier.cuh cand_index)
    if (S >= (1LL << 26)) throw CapacityFail{"a BFS level of " + std::to_string(S) + " states exceeds the 2^26 candidate index"};
    unsigned long long *cnt = h->counters.as<unsigned long long>();
    long long corners_upper = (long long)h->h_counters[CNT_CORNERS];
    const size_t per_state = std::max<size_t>((size_t)h->R * 32, 32);

    // Incremental mode: this level's plane rows stay resident (ping-pong with the previous level) so
    // that children copy the rows of the layers before their flipped neuron instead of recomputing.
    // Sharded march: a rank only ever writes the rows of the states it owns, so the level buffer holds just those, at
    // their position in the bucket-sorted permutation ("compact rows"); the parent's rows are found through the
    // previous level's inverse permutation.  The number of owned states is known on the host (bucket histogram).
    const bool sharded = h->shard_world > 1;
    long long n_owned = S;
    if (sharded) {
        if (lb == 0) {
            n_owned = (S - h->shard_rank + h->shard_world - 1) / h->shard_world;
        } else if (h->next_valid) {
            n_owned = 0;
            for (int b = 1; b <= h->D; ++b) n_owned += h->h_next[b];
        }
    }
    const bool compact = sharded && h->p2p && h->compact_rows && (lb == 0 || h->next_valid);   // peer-memory mode only
    const long long n_rows = compact ? std::max<long long>(n_owned, 1) : S;
    const bool resident = h->incremental && h->D >= 2 && S <= (1 << 26) &&
                          2 * (size_t)n_rows * per_state <= h->resident_budget;
    if (resident) {
        const int cur = h->prev_resident ? 1 - h->prev_buf : 0;
        DevBuf &pb = h->lvl_planes[cur];
        pb.reserve((size_t)n_rows * per_state, 0, true);
        double *base = pb.as<double>();
        h->ensure_chunk_scratch_no_planes((size_t)S);
        h->slot_buf = 1 - h->slot_buf;
        if (compact) h->slot_of[h->slot_buf].reserve((size_t)S * 4, 0, false);
        h->cur_rows_by_slot = compact ? 1 : 0;
        h->cur_slot_of = compact ? h->slot_of[h->slot_buf].as<int>() : nullptr;
        h->cur_prev_slot_of = (h->prev_resident && h->prev_compact) ? h->slot_of[1 - h->slot_buf].as<int>() : nullptr;
        Scratch sc = scratch_of(h, (size_t)S);
        h->ensure_corners((size_t)(corners_upper + S * VSLOTS), (size_t)corners_upper);
        const uint32_t *keys0 = h->keys.as<uint32_t>() + (size_t)lb * h->kw;
        const bool timing = h->timing_on();
        size_t t0 = 0;
        if (timing) t0 = h->span_begin();

        // bucket the states by the layer of their flipped neuron, build the sorted permutation
        const int D = h->D;
        h->bucket.reserve((size_t)S * 4, 0, false);
        h->perm.reserve((size_t)S * 4, 0, false);
        h->bcounts.reserve((size_t)3 * (D + 3) * 4, 0, false);
        int *counts = h->bcounts.as<int>(), *cursor = counts + (D + 3), *npre_d = counts + 2 * (D + 3);
        LayerOffs lo{};
        lo.D = D;
        for (int l = 1; l <= D + 1; ++l) lo.off[l] = h->off[l];
        const unsigned sb = (unsigned)((S + 255) / 256);
        std::vector<int> npre(D + 3, 0);
        // launch sizes of this level: predicted by the winner kernel of the previous level (read back
        // together with n_new), computed for the seed level, or -- fallback -- read back now
        bool have = false;
        if (lb == 0) {                       // seeds: every state is recomputed from layer 2 on
            const int mine = sharded ? (int)((S - h->shard_rank + h->shard_world - 1) / h->shard_world) : (int)S;
            for (int b = 1; b <= D; ++b) npre[b] = mine;
            npre[D + 1] = (int)S;
            have = true;
        } else if (h->next_valid) {
            int run = 0;
            for (int b = 1; b <= D; ++b) {
                run += h->h_next[b];
                npre[b] = run;
            }
            if (!h->prev_resident)           // parents' rows are gone: everything owned is recomputed
                for (int b = 1; b <= D; ++b) npre[b] = run;
            npre[D + 1] = (int)S;
            have = true;
        }
        if (have) {      // bucket sizes known: classify + scatter in one launch, bucket starts as launch parameters
            BucketBase bb{};
            for (int b = 1; b <= D + 1; ++b) bb.base[b] = npre[b - 1];
            launch_k(classify_scatter_kernel, dim3(sb), dim3(256), 0, st, h->via.as<int>(), h->parent.as<int>(), (int)lb, (int)S,
                                                        h->prev_resident ? (int)h->prev_lb : 0,
                                                        h->prev_resident ? (int)h->prev_S : 0, lo, bb, h->bucket.as<int>(),
                                                        h->level_cursor.as<int>(), h->perm.as<int>(),
                                                        sharded ? h->owner.as<uint8_t>() : nullptr, h->shard_rank,
                                                        const_cast<int *>(h->cur_slot_of));
            ++h->stats.n_launches;
        } else {
            if (compact) throw CudaFail{"compact level buffers need the bucket sizes on the host"};
            CK(cudaMemsetAsync(counts, 0, (size_t)(D + 3) * 4, st));
            classify_kernel<<<sb, 256, 0, st>>>(h->via.as<int>(), h->parent.as<int>(), (int)lb, (int)S,
                                                h->prev_resident ? (int)h->prev_lb : 0, h->prev_resident ? (int)h->prev_S : 0,
                                                lo, h->bucket.as<int>(), counts, sharded ? h->owner.as<uint8_t>() : nullptr,
                                                h->shard_rank);
            ++h->stats.n_launches;
            bucket_offsets_kernel<<<1, 32, 0, st>>>(counts, cursor, npre_d, D);
            ++h->stats.n_launches;
            scatter_kernel<<<sb, 256, 0, st>>>(h->bucket.as<int>(), (int)S, cursor, h->perm.as<int>());
            ++h->stats.n_launches;
        }
        CK(cudaGetLastError());
        if (!have) CK(cudaMemcpyAsync(h->h_npre, npre_d, (size_t)(D + 2) * 4, cudaMemcpyDeviceToHost, st));
        if (sharded && !h->p2p) CK(cudaMemsetAsync(sc.cnt, 0, (size_t)S * 4, st));   // counts of states owned elsewhere
        h->lazy_prev = nullptr;
        if (h->prev_resident && h->lazy_ok) {        // no copy: slice / level-plane / clip kernels read the parent's rows
            h->lazy_prev = h->lvl_planes[h->prev_buf].as<double>();
            h->lazy_lb = lb;
            h->lazy_prev_lb = h->prev_lb;
        } else if (h->prev_resident) {
            copy_parent_rows_kernel<<<(unsigned)((S + 7) / 8), 256, 0, st>>>(
                h->bucket.as<int>(), h->parent.as<int>(), (int)lb, (int)S, (int)h->prev_lb,
                h->lvl_planes[h->prev_buf].as<double>(), base, 4LL * h->R, lo, h->n1, h->cur_slot_of, h->cur_prev_slot_of);
            ++h->stats.n_launches;
            CK(cudaGetLastError());
        }
        if (!have) {
            CK(cudaStreamSynchronize(st));
            for (int b = 0; b <= D + 1; ++b) npre[b] = h->h_npre[b];
        }
        const int n_mine = npre[D];                       // states this rank composes and clips
        h->shard_owned_states += n_mine;
        // every chain clips its own states right after composing them (perm covers all S states when not sharded)
        const int n_clip = sharded ? n_mine : (int)S;
        // a single chain on one GPU walks the states in id order (siblings are neighbours: better DRAM / TLB locality
        // than the bucket-sorted order); chains and the sharded march go through the permutation
        const bool in_order = !sharded && h->planned_chains() == 1 && !h->force_perm_order;
        const int *order = in_order ? nullptr : h->perm.as<int>();
        const std::function<void(int, int, cudaStream_t)> tail = [&](int c, int nc, cudaStream_t cs) {
            if (nc == 1) {
                if (timing) h->span_end(t0, 1);
                if (timing) t0 = h->span_begin();
            }
            run_clip(h, lb, n_clip, base, flip, order, sc, c, nc, cs);
        };
        h->compose_chunk(keys0, (int)S, iso, base, h->perm.as<int>(), npre.data(), n_clip, order, &tail);
        if (h->last_n_chain > 1 && timing) {     // chained: the compose span covers the chains' clip launches too
            h->span_end(t0, 1);
            t0 = h->span_begin();
        }
        if (sharded && h->p2p) store_faces_p2p(h, lb, (int)S, sc, h->perm.as<int>(), n_mine);   // NVLink peer pushes
        else store_faces(h, lb, (int)S, sc, /*fused=*/!sharded);   // sharded (NCCL scheme): + the level's collectives
        h->lazy_prev = nullptr;
        if (timing) h->span_end(t0, 2);
        h->prev_resident = true;
        h->prev_compact = compact;
        h->cur_rows_by_slot = 0;
        h->cur_slot_of = h->cur_prev_slot_of = nullptr;
        h->prev_buf = cur;
        h->prev_lb = lb;
        h->prev_S = S;
        h->n_incremental_levels++;
    } else {
        if (h->shard_world > 1)
            throw CapacityFail{"sharded mode needs the level's plane rows resident: " + std::to_string(S) +
                               " states do not fit AM_B200_RESIDENT_GIB"};
        h->prev_resident = false;
        const size_t chunk = h->chunk_states();
        for (long long c0 = 0; c0 < S; c0 += (long long)chunk) {
            const int Sc = (int)std::min<long long>((long long)chunk, S - c0);
            const long long sid0 = lb + c0;
            h->ensure_chunk_scratch((size_t)Sc);
            h->ensure_corners((size_t)(corners_upper + (long long)Sc * VSLOTS), (size_t)corners_upper);
            const uint32_t *keys0 = h->keys.as<uint32_t>() + (size_t)sid0 * h->kw;
            size_t t0 = 0;
            const bool timing = h->timing_on();
            if (timing) t0 = h->span_begin();
            h->compose_chunk(keys0, Sc, iso, h->planes.as<double>(), nullptr, nullptr, Sc, nullptr);
            if (timing) h->span_end(t0, 1);
            if (timing) t0 = h->span_begin();
            Scratch sc = scratch_of(h, (size_t)Sc);
            run_clip(h, sid0, Sc, h->planes.as<double>(), flip, nullptr, sc);
            store_faces(h, sid0, Sc, sc);
            if (timing) h->span_end(t0, 2);
            corners_upper += (long long)Sc * VSLOTS;
        }
    }

    // ---- neighbour enumeration + visited set ----------------------------------------------------
    const bool fused_faces = resident && (h->shard_world == 1 || h->p2p);   // corner total advanced by the winner kernel
    size_t t0 = 0;
    const bool timing = h->timing_on();
    if (timing) t0 = h->span_begin();
    h->ensure_table((size_t)h->n_states + (size_t)S * VSLOTS);
    // room for the children before their number is known (a level usually discovers about as many states as it
    // holds); finalize_kernel is guarded and re-run in the rare case that this was not enough
    h->ensure_states((size_t)h->n_states + (size_t)std::min<long long>(S * VSLOTS, 2 * S + 65536));
    h->cand_slot.reserve((size_t)S * VSLOTS * 4, 0, false);
    h->nwin.reserve((size_t)S * 4, 0, false);
    h->wbase.reserve((size_t)S * 4, 0, false);
    LevelArgs a{};
    a.keys = h->keys.as<uint32_t>(); a.hsum = h->hsum.as<unsigned long long>();
    a.face_off = h->face_off.as<long long>(); a.face_edges = h->face_edges.as<int>();
    a.face_xyz = h->face_xyz.as<double>();
    a.kw = h->kw; a.kw4 = h->kw4; a.L = h->L; a.lb = (int)lb; a.S = (int)S;
    a.table = TableRef{h->table.as<unsigned long long>(), h->tcap - 1};
    a.cand_slot = h->cand_slot.as<int>(); a.nwin = h->nwin.as<uint32_t>(); a.win_base = h->wbase.as<uint32_t>();
    a.counters = cnt;
    const int G = h->G;
    const unsigned gb = (unsigned)((S * G + 255) / 256);
    const bool p2p = h->p2p && h->shard_world > 1;
    a.world = p2p ? h->shard_world : 1;
    a.rank = h->shard_rank;
    a.wmask = nullptr;
    if (p2p && S > h->xlay.mask_cap)
        throw CapacityFail{"sharded march: a BFS level of " + std::to_string(S) + " states exceeds the winner-mask region"};
    size_t tk = 0;
    if (timing) tk = h->span_begin();
    h->dispatch_group([&](auto g) { launch_k(expand_insert_kernel<decltype(g)::value>, dim3(gb), dim3(256), 0, st, a); });
    if (timing) h->span_end(tk, 9);
    ++h->stats.n_launches;
    {
        LayerOffs lo{};
        lo.D = h->D;
        for (int l = 1; l <= h->D + 1; ++l) lo.off[l] = h->off[l];
        a.owner = h->shard_world > 1 ? h->owner.as<uint8_t>() : nullptr;
        CK(cudaMemsetAsync(h->next_counts.p, 0, (size_t)(h->D + 3) * 4, st));
        // winners per parent + bucket histogram of the children + the prefix sums that number them: one launch
        // (xchg.cuh winners_scan_kernel).  It also advances the running corner total of the fused CSR path, clears
        // the small cursors of the next level's kernels and, in sharded mode, deals the free children to the ranks.
        h->wmask.reserve((size_t)S * 4, 0, false);
        h->wbase.reserve((size_t)S * 8, 0, false);
        const size_t nb = (size_t)(S + FS_TILE - 1) / FS_TILE + 1;
        h->fs_sums2.reserve(nb * 8, 0, false);
        h->fs_off2.reserve(nb * 8, 0, false);
        WinArgs w{};
        w.lo = lo; w.next_counts = h->next_counts.as<int>(); w.rank = h->shard_rank; w.world = h->shard_world;
        w.wmask = h->wmask.as<uint32_t>(); w.win_base = h->wbase.as<unsigned long long>();
        w.fs.block_sums = h->fs_sums2.as<unsigned long long>(); w.fs.block_off = h->fs_off2.as<unsigned long long>();
        w.fs.ticket = h->fs_ticket.as<unsigned int>() + 1; w.fs.total = cnt + CNT_NEW;
        w.fs.bump_dst = fused_faces ? cnt + CNT_CORNERS : nullptr; w.fs.bump_src = cnt + CNT_CHUNK_CORNERS;
        w.zero_a = h->level_cursor.as<int>(); w.n_zero_a = h->D + 3; w.zero_b = h->xcursor.as<int>();
        w.own = p2p ? h->xpeers.base[h->shard_rank] : nullptr; w.lay = h->xlay;
        w.balance = (h->shard_world > 1 && h->balance) ? 1 : 0;
        w.loads = h->bal_loads.as<unsigned long long>(); w.cuts = h->bal_cuts.as<int>();
        w.unit_clip = 3;
        const unsigned wb = (unsigned)((S + FS_TILE - 1) / FS_TILE);
        if (p2p) {   // winners of the candidates whose hash this rank owns -> every rank; OR after the barrier
            launch_k(xchg_push_masks_kernel, dim3((unsigned)((S + 255) / 256)), dim3(256), 0, st, a, h->xpeers, h->xlay);
            ++h->stats.n_launches;
            xchg_barrier(h);
            launch_k(winners_scan_kernel<true>, dim3(wb), dim3(FS_THREADS), 0, st, a, w);
        } else {
            launch_k(winners_scan_kernel<false>, dim3(wb), dim3(FS_THREADS), 0, st, a, w);
        }
        a.wmask = w.wmask;
        a.win_base64 = w.win_base; a.win_off64 = w.fs.block_off;
        a.cuts = w.balance ? w.cuts : nullptr;
        a.free_below_bit = h->off[2 <= h->D ? 2 : h->D + 1];
        a.n_ranks = h->shard_world;
        ++h->stats.n_launches;
        CK(cudaGetLastError());
    }
    // The winners' appends are launched at once (the arenas were grown speculatively above); the level's counters
    // travel to the host on a second stream meanwhile, so the host learns n_new -- and enqueues the next level --
    // while finalize_kernel is still running: no idle gap around the level's one host synchronisation.
    CK(cudaEventRecord(h->win_event, st));
    CK(cudaStreamWaitEvent(h->copy_stream, h->win_event, 0));
    CK(cudaMemcpyAsync(h->h_next, h->next_counts.p, (size_t)(h->D + 3) * 4, cudaMemcpyDeviceToHost, h->copy_stream));
    CK(cudaMemcpyAsync(h->h_counters, h->counters.p, CNT_NUM * 8, cudaMemcpyDeviceToHost, h->copy_stream));
    auto launch_finalize = [&]() {
        a.keys = h->keys.as<uint32_t>(); a.hsum = h->hsum.as<unsigned long long>();
        a.face_off = h->face_off.as<long long>();
        a.keys_w = h->keys.as<uint32_t>(); a.hsum_w = h->hsum.as<unsigned long long>();
        a.parent = h->parent.as<int>(); a.via_edge = h->via.as<int>(); a.seedpt = h->seedpt.as<double>();
        a.owner = h->shard_world > 1 ? h->owner.as<uint8_t>() : nullptr;
        a.n_states = (int)h->n_states;
        a.cap_states = (int)std::min<size_t>(h->cap_states, (size_t)0x7FFFFFFF);
        if (timing) tk = h->span_begin();
        // the winners' keys are copied by the group: one uint4 per lane and 128 key bits (a parent without winners
        // leaves at once, so the wider group costs nothing there)
        const int Gf = h->finalize_G;
        const unsigned gbf = (unsigned)((S * Gf + 255) / 256);
        switch (Gf) {
            case 32: launch_k(finalize_kernel<32>, dim3(gbf), dim3(256), 0, st, a); break;
            case 16: launch_k(finalize_kernel<16>, dim3(gbf), dim3(256), 0, st, a); break;
            default: h->dispatch_group([&](auto g) { launch_k(finalize_kernel<decltype(g)::value>, dim3(gb), dim3(256), 0, st, a); });
        }
        if (timing) h->span_end(tk, 11);
        ++h->stats.n_launches;
        CK(cudaGetLastError());
    };
    const size_t cap_at_launch = h->cap_states;
    launch_finalize();
    {
        const auto w0 = std::chrono::steady_clock::now();
        CK(cudaStreamSynchronize(h->copy_stream));        // the one host sync of the level
        h->host_wait_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
    }
    h->next_valid = true;
    if (h->h_counters[CNT_XCHG_ERROR] != 0) {
        const unsigned long long e = h->h_counters[CNT_XCHG_ERROR];
        throw CudaFail{"sharded march: the exchange over peer memory failed (" + std::to_string(e & 0xFFFF) +
                       " barrier time-outs, " + std::to_string((e >> 16) & 0xFFFF) + " bad records, " +
                       std::to_string(e >> 32) + " polygons beyond the inbox capacity)"};
    }
    const long long n_new = (long long)h->h_counters[CNT_NEW];
    if (h->n_states + n_new >= (1LL << 31) - 1) throw CapacityFail{"more than 2^31 states"};
    if ((size_t)(h->n_states + n_new) > cap_at_launch) {   // more children than reserved: the guarded kernel skipped them
        h->ensure_states((size_t)(h->n_states + n_new));
        launch_finalize();                                 // idempotent
    }
    if (timing) h->span_end(t0, 3);
    h->n_states += n_new;
}

// After a sharded march the visited set of this rank holds only the keys whose hash it owns; stitching and the
// edge-incidence check look up arbitrary states, so the table is rebuilt from the (replicated) stored hashes.
void ensure_full_table(am_handle *h)
{
    if (!h->table_sharded) return;
    h->table_sharded = false;
    const uint32_t before = h->tcap;
    h->ensure_table((size_t)h->n_states);              // grows + rehashes everything if it was too small
    if (h->tcap == before && h->n_states > 0) {
        CK(cudaMemsetAsync(h->table.p, 0xFF, (size_t)h->tcap * 8, h->stream));
        TableRef t{h->table.as<unsigned long long>(), h->tcap - 1};
        rehash_kernel<<<(unsigned)((h->n_states + 255) / 256), 256, 0, h->stream>>>(h->hsum.as<unsigned long long>(),
                                                                                     (int)h->n_states, t, 1, 0);
        ++h->stats.n_launches;
        CK(cudaGetLastError());
    }
}

void resolve_spans(am_handle *h)
{
    double ms_k[am_handle::N_KINDS] = {};
    h->gemm_ms = 0; h->gemm_flops = 0; h->gemm_launches = 0;
    for (int k = 0; k < am_handle::N_KINDS; ++k) h->kind_ms[k] = h->kind_flops[k] = 0.0, h->kind_launches[k] = 0;
    for (const auto &s : h->spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->ev_pool[s.a], h->ev_pool[s.b]) != cudaSuccess) { cudaGetLastError(); continue; }
        const double sc = (double)h->span_scale;       // sampled levels -> estimate for the whole march
        ms *= (float)sc;
        ms_k[s.kind] += ms;
        h->kind_ms[s.kind] += ms;
        h->kind_flops[s.kind] += s.flops * sc;
        h->kind_launches[s.kind] += h->span_scale;
        if (s.kind == 0) { h->gemm_flops += s.flops * sc; h->gemm_launches += h->span_scale; }
    }
    h->gemm_ms = ms_k[0];
    h->stats.seconds_compose = ms_k[1] * 1e-3;
    h->stats.seconds_clip = ms_k[2] * 1e-3;
    h->stats.seconds_frontier = ms_k[3] * 1e-3;
}

}  // namespace

// =================================================================================================
//                                           C ABI
// =================================================================================================
extern "C" {

const char *am_last_error(const am_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int am_create(am_handle **out, int is_f64, const int *nodes, int n_nodes, const int *arc_table, int arc_rows,
              int arc_cols, int num_extra_constraints)
{
    g_create_error.clear();
    if (!out || !nodes || n_nodes < 3 || nodes[0] != 3 || nodes[n_nodes - 1] != 1 || num_extra_constraints < 0 ||
        arc_rows != n_nodes - 2 || (arc_rows > 0 && (!arc_table || arc_cols < 1 || (arc_cols + 1) % 2 != 0))) {
        g_create_error = "am_create: malformed architecture (need nodes = [3, ..., 1], one arc_table row per hidden "
                         "layer, an odd number of arc_table columns, num_extra_constraints >= 0)";
        return AM_ERR_ARG;
    }
    am_handle *h = new am_handle();
    try {
        h->f64 = is_f64 != 0;
        h->D = n_nodes - 2;
        h->n.assign(nodes, nodes + n_nodes);
        for (int v : h->n)
            if (v <= 0) throw CudaFail{"non-positive layer width"};
        h->off.assign(h->D + 2, 0);
        for (int l = 1; l <= h->D; ++l) h->off[l + 1] = h->off[l] + h->n[l];
        h->L = h->off[h->D + 1];
        h->E = num_extra_constraints;
        h->kw4 = (h->L + 127) / 128;
        h->kw = 4 * h->kw4;
        h->n1 = h->n[1];
        h->R = h->L - h->n1;
        h->G = std::min(pow2_group(h->kw4), 8);      // measured at L = 4096: 8 lanes per state beat 32 (fewer idle lanes)
        if (const char *e = getenv("AM_B200_FRONTIER_G")) {     // lanes per state in the frontier / stitching kernels
            const int g = atoi(e);
            if (g == 1 || g == 2 || g == 4 || g == 8 || g == 16 || g == 32) h->G = std::min(g, h->G);
        }
        h->skips.assign(h->D + 1, {});
        h->n_tm = 0;
        for (int r = 0; r < arc_rows; ++r) {
            const int *row = arc_table + (size_t)r * arc_cols;
            if (row[0] < 0 || 1 + 2 * row[0] > arc_cols) throw CudaFail{"arc_table row " + std::to_string(r) + " is too short"};
            for (int j = 0; j < row[0]; ++j) {
                Skip sk{row[1 + 2 * j], row[2 + 2 * j]};
                if (sk.src < 0 || sk.src > r + 1 || sk.tm < 0)
                    throw CudaFail{"arc_table row " + std::to_string(r) + ": bad source/transform index"};
                h->skips[r + 1].push_back(sk);
                h->n_tm = std::max(h->n_tm, sk.tm + 1);
            }
        }
        for (DevBuf *b : {&h->keys, &h->hsum, &h->parent, &h->via, &h->seedpt, &h->face_off, &h->owner, &h->face_edges,
                          &h->face_xyz, &h->lvl_planes[0], &h->lvl_planes[1], &h->xchg, &h->cand_slot, &h->f_verts,
                          &h->planes, &h->table, &h->cmb_owner, &h->cmb_flag, &h->cmb_vid, &h->cmb_cvid, &h->cmb_verts})
            b->vm = true;
        h->Wt.resize(h->D + 1); h->bias.resize(h->D + 1);
        h->Wrow.resize(h->D + 1); h->Brow.resize(h->D + 1); h->sd_act.resize(h->D + 1);
        h->Mpad.assign(h->D + 1, 0); h->Kpad.assign(h->D + 1, 0);
        CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&h->win_event, cudaEventDisableTiming));
        if (const char *e = getenv("AM_B200_CHAINS")) h->n_chains = std::max(0, std::min(am_handle::MAX_CHAINS, atoi(e)));
        CK(cudaEventCreateWithFlags(&h->fork_event, cudaEventDisableTiming));
        for (int c = 0; c < am_handle::MAX_CHAINS; ++c) {
            CK(cudaStreamCreateWithFlags(&h->chain_stream[c], cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&h->join_event[c], cudaEventDisableTiming));
        }
        CK(cudaMallocHost(&h->h_counters, CNT_NUM * 8));
        memset(h->h_counters, 0, CNT_NUM * 8);
        if (h->D + 2 > MAX_LAYERS) throw CudaFail{"more than 62 hidden layers"};
        CK(cudaMallocHost(&h->h_npre, (size_t)(h->D + 3) * 4));
        CK(cudaMallocHost(&h->h_next, (size_t)(h->D + 3) * 4));
        h->next_counts.reserve((size_t)(h->D + 3) * 4);
        h->level_cursor.reserve((size_t)(h->D + 3) * 4);
        h->fs_ticket.reserve(64);
        h->bal_loads.reserve(XCHG_MAX_WORLD * 8);
        h->bal_cuts.reserve((XCHG_MAX_WORLD + 1) * 4);
        if (const char *e = getenv("AM_B200_BALANCE")) h->balance = atoi(e) != 0;
        if (const char *e = getenv("AM_B200_PERM_ORDER")) h->force_perm_order = atoi(e) != 0;
        if (const char *e = getenv("AM_B200_EQU_WARP")) h->equ_warp = atoi(e) != 0;
        if (const char *e = getenv("AM_B200_PUSH_IN_CLIP")) h->push_in_clip = atoi(e) != 0;
        if (const char *e = getenv("AM_B200_COMPACT_ROWS")) h->compact_rows = atoi(e) != 0;
        h->finalize_G = std::min(32, std::max(h->G, pow2_group(h->kw4)));          // one lane per 128 key bits
        if (const char *e = getenv("AM_B200_FINALIZE_G")) {
            const int g = atoi(e);
            if (g == 8 || g == 16 || g == 32) h->finalize_G = g;
        }
        if (h->finalize_G != 32 && h->finalize_G != 16) h->finalize_G = h->G;
        if (const char *e = getenv("AM_B200_TRACE_EVERY")) h->trace_every = std::max(1, atoi(e));
        if (const char *e = getenv("AM_B200_PDL_BELOW")) h->pdl_below = atoll(e);
        h->xcursor.reserve(64);
        if (const char *e = getenv("AM_B200_INCREMENTAL")) h->incremental = atoi(e) != 0;
        CK(cudaFuncSetAttribute(clip_kernel<2, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)clip_ring_bytes(2, 3)));
        CK(cudaFuncSetAttribute(clip_kernel<3, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)clip_ring_bytes(2, 3)));
        CK(cudaFuncSetAttribute(clip_kernel<2, 2, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)clip_ring_bytes(2, 3)));
        CK(cudaFuncSetAttribute(clip_kernel<2, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)clip_ring_bytes(2, 4)));
        CK(cudaFuncSetAttribute(clip_kernel<2, 2, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)clip_ring_bytes(2, 5)));
        CK(cudaFuncSetAttribute(clip_kernel<3, 1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)clip_ring_bytes(1, 4)));
        CK(cudaFuncSetAttribute(clip_kernel<4, 1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)clip_ring_bytes(1, 4)));
        CK(cudaFuncSetAttribute(clip_kernel<3, 1, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)clip_ring_bytes(1, 6)));
        if (const char *e = getenv("AM_B200_CLIP_MINB")) h->clip_minb = atoi(e);
        h->counters.reserve(CNT_NUM * 8);
        {
            int kmax = 3;
            for (int l = 1; l <= h->D; ++l) kmax = std::max(kmax, h->n[l]);
            // default: wide layers go to the tcgen05 split-integer path, narrow ones (tiles mostly padding) to FP64 DMMA
            h->gemm_variant = (kmax >= 256 && kmax <= 8192) ? 2 : 0;   // measured: 128-wide 16.2 (DMMA) vs 10.0 M faces/s, 500-wide 12.6 vs 15.5
            if (const char *e = getenv("AM_B200_GEMM_VARIANT")) h->gemm_variant = atoi(e);
            auto prep = [&](auto cfg, auto kern) {
                const size_t need = decltype(cfg)::smem_bytes(kmax);
                if (need > 227 * 1024) throw CudaFail{"hidden layers this wide are not supported by the composition kernel"};
                CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
            };
            if (const char *e = getenv("AM_B200_SPLIT_DIGITS")) h->split_digits = std::max(6, std::min(8, atoi(e)));
            // read-through of the parent's rows: the consumers that know how are the tcgen05 digit kernels, the
            // level-plane kernel and the clip kernel; skips that read a hidden layer keep the copy kernel
            h->lazy_ok = (h->gemm_variant == 2);
            for (int l = 1; l <= h->D; ++l)
                for (const Skip &sk : h->skips[l])
                    if (sk.src >= 1) h->lazy_ok = false;
            if (const char *e = getenv("AM_B200_READ_THROUGH")) h->lazy_ok = h->lazy_ok && atoi(e) != 0;
            h->splitW.resize(h->D + 1);
            {
                int dev = 0;
                CK(cudaGetDevice(&dev));
                CK(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, dev));
            }
            switch (h->gemm_variant) {
                case 1: prep(GemmWide{}, compose_gemm_kernel<GemmWide>); break;
                case 2:
                    if (kmax > 8192) throw CudaFail{"hidden layers wider than 8192 overflow the int32 digit accumulators"};
                    if (const char *e = getenv("AM_B200_SPLIT_EPI")) h->split_epi = atoi(e);
                    if (h->split_epi != 2 && h->split_epi != 4) h->split_epi = 1;
                    if (const char *e = getenv("AM_B200_SPLIT_BK")) h->split_bk = (atoi(e) == 32) ? 32 : 64;
                    for_split_digits([&](auto sd) {
                        constexpr int SD = decltype(sd)::value;
                        const int bytes = (int)SplitCfg<SD>::SMEM;
                        CK(cudaFuncSetAttribute(split_gemm_kernel<SD, 1, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
                        CK(cudaFuncSetAttribute(split_gemm_kernel<SD, 2, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
                        CK(cudaFuncSetAttribute(split_gemm_kernel<SD, 4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
                        CK(cudaFuncSetAttribute(split_gemm_kernel<SD, 1, 16, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)SplitCfg<SD, 64>::SMEM));
                    });
                    break;
                default: prep(GemmDefault{}, compose_gemm_kernel<GemmDefault>); break;
            }
        }
    } catch (const CudaFail &f) {
        g_create_error = "am_create: " + f.msg;
        h->free_all();
        if (h->stream) cudaStreamDestroy(h->stream);
        delete h;
        return AM_ERR_CUDA;
    }
    *out = h;
    return AM_OK;
}

void am_destroy(am_handle *h)
{
    if (!h) return;
    if (h->nccl_comm) {
        cudaDeviceSynchronize();
        NcclApi::get().destroy(h->nccl_comm);
        h->nccl_comm = nullptr;
    }
    if (h->xblock) {
        cudaDeviceSynchronize();
        for (int q = 0; q < h->xpeers.world; ++q)
            if (q != h->xpeers.rank && h->xpeers.base[q]) cudaIpcCloseMemHandle(h->xpeers.base[q]);
        cudaFree(h->xblock);
        h->xblock = nullptr;
    }
    h->free_all();
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int am_set_shard(am_handle *h, int rank, int world, am_allreduce_fn fn, void *user)
{
    if (!h) return AM_ERR_ARG;
    if (world < 1 || world > 255 || rank < 0 || rank >= world || (world > 1 && !fn)) {
        h->err = "am_set_shard: need 0 <= rank < world <= 255 and a callback when world > 1";
        return AM_ERR_ARG;
    }
    h->shard_rank = rank;
    h->shard_world = world;
    h->shard_cb = fn;
    h->shard_user = user;
    return AM_OK;
}

int am_nccl_unique_id(void *out128)
{
    NcclApi &api = NcclApi::get();
    if (!api.ok || !out128) return AM_ERR_STATE;
    NcclApi::UniqueId id;
    if (api.get_id(&id) != 0) return AM_ERR_CUDA;
    memcpy(out128, id.internal, 128);
    return AM_OK;
}

int am_set_shard_nccl(am_handle *h, int rank, int world, const void *unique_id128)
{
    if (!h) return AM_ERR_ARG;
    NcclApi &api = NcclApi::get();
    if (!api.ok) {
        h->err = "am_set_shard_nccl: libnccl.so.2 could not be loaded";
        return AM_ERR_STATE;
    }
    if (world < 2 || world > 255 || rank < 0 || rank >= world || !unique_id128) {
        h->err = "am_set_shard_nccl: need 0 <= rank < world, 2 <= world <= 255 and a unique id";
        return AM_ERR_ARG;
    }
    if (h->nccl_comm) {
        api.destroy(h->nccl_comm);
        h->nccl_comm = nullptr;
    }
    NcclApi::UniqueId id;
    memcpy(id.internal, unique_id128, 128);
    const int rc = api.init_rank(&h->nccl_comm, world, id, rank);
    if (rc != 0) {
        h->err = "am_set_shard_nccl: ncclCommInitRank failed (" + std::to_string(rc) + ")";
        h->nccl_comm = nullptr;
        return AM_ERR_CUDA;
    }
    h->shard_rank = rank;
    h->shard_world = world;
    return AM_OK;
}

namespace {

// f(points) for P points [P][3] on the device -> h->sd_val[P]; optionally the packed activation keys
void seed_forward(am_handle *h, const double *pts, int P, uint32_t *keys)
{
    cudaStream_t st = h->stream;
    const int D = h->D;
    auto gemm = [&](const double *in, int ldi, const double *W, const double *b, double *out, int ldo, int M, int K, int acc) {
        dim3 grid((unsigned)((M + SG_T - 1) / SG_T), (unsigned)((P + SG_T - 1) / SG_T));
        seed_gemm_kernel<<<grid, 256, 0, st>>>(in, ldi, W, b, out, ldo, P, M, K, acc);
        ++h->stats.n_launches;
    };
    auto skips_into = [&](int hfc, double *out, int ldo, int M) {      // fc layer hfc: hidden hfc -> hidden hfc + 1 / output
        for (const Skip &sk : h->skips[hfc]) {
            const bool identity = (h->tm_h[sk.tm] == 0 && h->tm_w[sk.tm] == 0);
            const double *src = (sk.src == 0) ? pts : h->sd_act[sk.src].as<double>();
            const int lds = (sk.src == 0) ? 3 : h->n[sk.src];
            if (identity) {
                const int m = std::min(M, lds);
                seed_add_kernel<<<(unsigned)(((long long)P * m + 255) / 256), 256, 0, st>>>(out, ldo, src, lds, P, m);
                ++h->stats.n_launches;
            } else {
                gemm(src, lds, h->TM[sk.tm].as<double>(), nullptr, out, ldo, M, lds, 1);
            }
        }
    };
    if (keys) CK(cudaMemsetAsync(keys, 0, (size_t)P * h->kw * 4, st));
    for (int l = 1; l <= D; ++l) h->sd_act[l].reserve((size_t)P * h->n[l] * 8, 0, false);
    gemm(pts, 3, h->Wrow[0].as<double>(), h->Brow[0].as<double>(), h->sd_act[1].as<double>(), h->n[1], h->n[1], 3, 0);
    for (int l = 1; l <= D; ++l) {
        if (l >= 2) skips_into(l - 1, h->sd_act[l].as<double>(), h->n[l], h->n[l]);
        const int words = (h->n[l] + 31) / 32;
        seed_relu_bits_kernel<<<(unsigned)(((long long)P * words + 255) / 256), 256, 0, st>>>(
            h->sd_act[l].as<double>(), h->n[l], P, h->n[l], h->off[l], keys, h->kw);
        ++h->stats.n_launches;
        if (l < D)
            gemm(h->sd_act[l].as<double>(), h->n[l], h->Wrow[l].as<double>(), h->Brow[l].as<double>(),
                 h->sd_act[l + 1].as<double>(), h->n[l + 1], h->n[l + 1], h->n[l], 0);
    }
    h->sd_val.reserve((size_t)P * 8, 0, false);
    gemm(h->sd_act[D].as<double>(), h->n[D], h->Wrow[D].as<double>(), h->Brow[D].as<double>(), h->sd_val.as<double>(), 1, 1,
         h->n[D], 0);
    skips_into(D, h->sd_val.as<double>(), 1, 1);
    CK(cudaGetLastError());
}

}  // namespace

int am_seed_dichotomy(am_handle *h, const void *const *W, const void *const *B, const void *const *TM, const int *tm_shapes,
                      int n_tm, const void *w_extra, const void *b_extra, int n_extra, double iso, int64_t init_num,
                      int64_t try_pts_num, double ball_radius, int iter_max, double avg_eps, uint64_t seed,
                      am_seed_report *report)
{
    if (!h || !W || !B || init_num < 1 || try_pts_num < 2 || !(ball_radius > 0.0) || iter_max < 0 ||
        (n_tm > 0 && (!TM || !tm_shapes)) || init_num > (1 << 22) || try_pts_num > (1 << 22)) {
        if (h) h->err = "am_seed_dichotomy: bad argument";
        return AM_ERR_ARG;
    }
    if (n_extra != h->E) {
        h->err = "am_seed_dichotomy: the number of extra constraints differs from am_create";
        return AM_ERR_ARG;
    }
    try {
        cudaStream_t st = h->stream;
        CK(cudaDeviceSynchronize());
        const auto t0 = std::chrono::steady_clock::now();
        load_weights(h, W, B, TM, tm_shapes, n_tm);
        {
            auto we = fetch_real(w_extra, (size_t)h->E * 3, h->f64);
            auto be = fetch_real(b_extra, (size_t)h->E, h->f64);
            std::vector<double> ex((size_t)h->E * 4 + 4, 0.0);
            for (int e = 0; e < h->E; ++e) {
                ex[4 * e + 0] = we[3 * e + 0]; ex[4 * e + 1] = we[3 * e + 1]; ex[4 * e + 2] = we[3 * e + 2];
                ex[4 * e + 3] = be[e];
            }
            upload(h->extra, ex.data(), ex.size() * 8, st);
            CK(cudaStreamSynchronize(st));
        }
        const int N = (int)init_num, T = (int)try_pts_num;
        h->sd_pts.reserve((size_t)T * 24, 0, false);
        h->sd_valid.reserve((size_t)T * 4, 0, false);
        h->sd_flags.reserve((size_t)T * 8, 0, false);     // is_pos | is_neg
        h->sd_offs.reserve((size_t)T * 8, 0, false);
        h->sd_lists.reserve((size_t)T * 8, 0, false);
        h->sd_pos.reserve((size_t)N * 24, 0, false);
        h->sd_neg.reserve((size_t)N * 24, 0, false);
        h->sd_mid.reserve((size_t)N * 24, 0, false);
        h->sd_err.reserve((size_t)N * 8 + 64, 0, false);
        h->sd_tot.reserve(64);
        h->sd_keys.reserve((size_t)N * h->kw * 4, 0, false);
        uint32_t *is_pos = h->sd_flags.as<uint32_t>(), *is_neg = is_pos + T;
        uint32_t *off_pos = h->sd_offs.as<uint32_t>(), *off_neg = off_pos + T;
        int *pos_list = h->sd_lists.as<int>(), *neg_list = pos_list + T;
        unsigned long long *tot = h->sd_tot.as<unsigned long long>();
        unsigned long long h_tot[2];
        int n_have = 0, rounds = 0;
        const unsigned tb = (unsigned)((T + 255) / 256);
        while (n_have < N) {
            if (rounds >= 4096)
                throw CudaFail{"no pair of trial points with opposite signs of f - iso within 4096 rounds (the reference "
                               "raises after time_out, backend/main.py:276-277)"};
            seed_sample_kernel<<<tb, 256, 0, st>>>(h->sd_pts.as<double>(), h->sd_valid.as<int>(), T, ball_radius,
                                                   h->extra.as<double>(), h->E, seed, (uint32_t)rounds);
            ++h->stats.n_launches;
            seed_forward(h, h->sd_pts.as<double>(), T, nullptr);
            seed_classify_kernel<<<tb, 256, 0, st>>>(h->sd_val.as<double>(), h->sd_valid.as<int>(), T, iso, is_pos, is_neg);
            ++h->stats.n_launches;
            h->scan(is_pos, off_pos, T, tot);
            h->scan(is_neg, off_neg, T, tot + 1);
            CK(cudaMemcpyAsync(h_tot, tot, 16, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            const long long n_pos = (long long)h_tot[0], n_neg = (long long)h_tot[1];
            if (n_pos > 0 && n_neg > 0) {
                const long long n_take = std::min<long long>(N - n_have, n_pos * n_neg);
                seed_lists_kernel<<<tb, 256, 0, st>>>(is_pos, off_pos, is_neg, off_neg, T, pos_list, neg_list);
                seed_pairs_kernel<<<(unsigned)((n_take + 255) / 256), 256, 0, st>>>(
                    h->sd_pts.as<double>(), pos_list, neg_list, (int)n_pos, (int)n_neg, (int)n_take, n_have,
                    splitmix64(seed ^ (0xA5A5A5A5ull + (uint64_t)rounds)), h->sd_pos.as<double>(), h->sd_neg.as<double>());
                h->stats.n_launches += 2;
                n_have += (int)n_take;
            }
            ++rounds;
        }
        // bisection (reference backend/main.py:314-326): stop when the MEAN |f - iso| over the pairs is below avg_eps
        int it = 0;
        double avg = 0.0;
        bool evaluated = false;
        const unsigned nb = (unsigned)((N + 255) / 256);
        for (; it < iter_max; ++it) {
            seed_mid_kernel<<<(unsigned)((3 * N + 255) / 256), 256, 0, st>>>(h->sd_pos.as<double>(), h->sd_neg.as<double>(),
                                                                           h->sd_mid.as<double>(), 3 * N);
            seed_forward(h, h->sd_mid.as<double>(), N, h->sd_keys.as<uint32_t>());
            seed_err_kernel<<<1, 1024, 0, st>>>(h->sd_val.as<double>(), iso, N, h->sd_err.as<double>());
            h->stats.n_launches += 2;
            double h_err = 0.0;
            CK(cudaMemcpyAsync(&h_err, h->sd_err.p, 8, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            avg = h_err / N;
            evaluated = true;
            if (avg < avg_eps) break;
            seed_bisect_kernel<<<nb, 256, 0, st>>>(h->sd_val.as<double>(), iso, h->sd_mid.as<double>(), h->sd_pos.as<double>(),
                                                   h->sd_neg.as<double>(), N);
            ++h->stats.n_launches;
            evaluated = false;
        }
        if (!evaluated) {      // iter_max reached (or zero): the points are the last midpoints, their states still missing
            seed_mid_kernel<<<(unsigned)((3 * N + 255) / 256), 256, 0, st>>>(h->sd_pos.as<double>(), h->sd_neg.as<double>(),
                                                                           h->sd_mid.as<double>(), 3 * N);
            seed_forward(h, h->sd_mid.as<double>(), N, h->sd_keys.as<uint32_t>());
            seed_err_kernel<<<1, 1024, 0, st>>>(h->sd_val.as<double>(), iso, N, h->sd_err.as<double>());
            double h_err = 0.0;
            CK(cudaMemcpyAsync(&h_err, h->sd_err.p, 8, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            avg = h_err / N;
        }
        CK(cudaStreamSynchronize(st));
        h->n_stored_seeds = N;
        if (report) {
            report->n_points = N;
            report->rounds = rounds;
            report->iterations = it;
            report->avg_abs_error = avg;
            report->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
    } catch (const CudaFail &f) {
        h->err = "am_seed_dichotomy: " + f.msg;
        return f.code;
    }
    return AM_OK;
}

int64_t am_num_seeds(const am_handle *h) { return h ? h->n_stored_seeds : 0; }

int am_copy_seeds(const am_handle *h, double *points, uint8_t *states_bool)
{
    if (!h || h->n_stored_seeds < 1) return AM_ERR_STATE;
    const size_t N = (size_t)h->n_stored_seeds;
    cudaError_t e = cudaSuccess;
    if (points) e = cudaMemcpy(points, h->sd_mid.p, N * 24, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && states_bool) {
        std::vector<uint32_t> keys(N * h->kw);
        e = cudaMemcpy(keys.data(), h->sd_keys.p, keys.size() * 4, cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < N && e == cudaSuccess; ++i)
            for (int j = 0; j < h->L; ++j) states_bool[i * h->L + j] = (keys[i * h->kw + (j >> 5)] >> (j & 31)) & 1u;
    }
    return e == cudaSuccess ? AM_OK : AM_ERR_CUDA;
}

int am_set_shard_p2p(am_handle *h, int rank, int world, am_allgather_fn fn, void *user)
{
    if (!h) return AM_ERR_ARG;
    if (world < 2 || world > XCHG_MAX_WORLD || rank < 0 || rank >= world || !fn) {
        h->err = "am_set_shard_p2p: need 0 <= rank < world, 2 <= world <= 16 and an all-gather callback";
        return AM_ERR_ARG;
    }
    try {
        if (h->xblock) throw CudaFail{"the exchange block of this handle is already set up"};
        double mib = 96.0;
        if (const char *e = getenv("AM_B200_XCHG_MIB")) mib = std::max(1.0, atof(e));
        if (const char *e = getenv("AM_B200_XCHG_TIMEOUT_MS")) h->xtimeout_ns = (unsigned long long)(atof(e) * 1e6);
        XchgLayout lay{};
        lay.cap_corners = (int)std::min<double>(mib * (1 << 20) / 28.0, double(1 << 28));
        lay.region_bytes = (lay.xyz_off() + (size_t)lay.cap_corners * 24 + 255) & ~size_t(255);
        lay.mask_cap = 1 << 21;
        if (const char *e = getenv("AM_B200_XCHG_LEVEL_STATES")) lay.mask_cap = std::max(1024, atoi(e));
        lay.mask_base = XCHG_CTRL_BYTES + (size_t)world * lay.region_bytes;
        lay.cnt_base = lay.mask_base + (size_t)world * lay.mask_cap * 4;
        lay.where_base = lay.cnt_base + (size_t)lay.mask_cap * 4;
        const size_t total = lay.where_base + (size_t)lay.mask_cap * 8;
        CK(cudaMalloc(&h->xblock, total));
        CK(cudaMemset(h->xblock, 0, total));
        CK(cudaDeviceSynchronize());
        cudaIpcMemHandle_t mine;
        CK(cudaIpcGetMemHandle(&mine, h->xblock));
        std::vector<cudaIpcMemHandle_t> all(world);
        if (fn(user, &mine, all.data(), (int64_t)sizeof(mine)) != 0) throw CudaFail{"the all-gather callback failed"};
        XchgPeers p{};
        p.world = world; p.rank = rank;
        for (int q = 0; q < world; ++q) {
            if (q == rank) { p.base[q] = static_cast<unsigned char *>(h->xblock); continue; }
            void *ptr = nullptr;
            CK(cudaIpcOpenMemHandle(&ptr, all[q], cudaIpcMemLazyEnablePeerAccess));
            p.base[q] = static_cast<unsigned char *>(ptr);
        }
        // nobody may write into a block before its owner has zeroed it: one more round trip as a host barrier
        std::vector<char> dummy(world * 8);
        long long tag = rank;
        if (fn(user, &tag, dummy.data(), 8) != 0) throw CudaFail{"the all-gather callback failed"};
        h->xpeers = p;
        h->xlay = lay;
        h->xepoch = 0;
        h->shard_rank = rank;
        h->shard_world = world;
        h->p2p = true;
    } catch (const CudaFail &f) {
        h->err = "am_set_shard_p2p: " + f.msg;
        return f.code;
    }
    return AM_OK;
}

int am_load_weights(am_handle *h, const void *const *W, const void *const *B, const void *const *TM, const int *tm_shapes,
                    int n_tm)
{
    if (!h || !W || !B || (n_tm > 0 && (!TM || !tm_shapes))) {
        if (h) h->err = "am_load_weights: null argument";
        return AM_ERR_ARG;
    }
    try {
        return load_weights(h, W, B, TM, tm_shapes, n_tm);
    } catch (const CudaFail &f) {
        h->err = "am_load_weights: " + f.msg;
        return AM_ERR_CUDA;
    }
}

int am_march(am_handle *h, const void *const *W, const void *const *B, const void *const *TM, const int *tm_shapes,
             int n_tm, const uint8_t *states, const void *points, int64_t n_seeds, const void *w_extra,
             const void *b_extra, int n_extra, double iso, int flip_insideout, void *user_stream)
{
    if (!h) return AM_ERR_ARG;
    const bool stored = (states == nullptr && points == nullptr);      // seeds of the last am_seed_dichotomy
    if (stored) n_seeds = h->n_stored_seeds;
    if (!W || !B || (!stored && (!states || !points)) || n_seeds < 1 || (n_tm > 0 && (!TM || !tm_shapes))) {
        h->err = stored ? "am_march: no stored seeds (call am_seed_dichotomy first)"
                        : "am_march: null argument or no seed states";
        return AM_ERR_ARG;
    }
    if (n_extra != h->E) {
        h->err = "am_march: " + std::to_string(n_extra) + " extra constraints given but the environment was created for " +
                 std::to_string(h->E);
        return AM_ERR_ARG;
    }
    h->has_march = false;
    h->has_mesh = false;
    try {
        cudaStream_t st = h->stream;
        if (user_stream) {
            // The inputs are read with blocking copies on the legacy stream (tensor_changed / fetch_real), which a
            // non-blocking caller stream does not order: wait on the host until the caller's stream has produced them.
            CK(cudaStreamSynchronize((cudaStream_t)user_stream));
        } else {
            CK(cudaDeviceSynchronize());
        }
        h->ev_used = 0;
        h->spans.clear();
        h->stats = am_stats{};
        h->host_wait_s = 0.0;
        h->span_scale = h->trace_every;
        const auto host_t0 = std::chrono::steady_clock::now();
        cudaEvent_t e_begin = h->ev();
        load_weights(h, W, B, TM, tm_shapes, n_tm);
        {
            auto we = fetch_real(w_extra, (size_t)h->E * 3, h->f64);
            auto be = fetch_real(b_extra, (size_t)h->E, h->f64);
            std::vector<double> ex((size_t)h->E * 4 + 4, 0.0);
            for (int e = 0; e < h->E; ++e) {
                ex[4 * e + 0] = we[3 * e + 0]; ex[4 * e + 1] = we[3 * e + 1]; ex[4 * e + 2] = we[3 * e + 2];
                ex[4 * e + 3] = be[e];
            }
            upload(h->extra, ex.data(), ex.size() * 8, st);
            CK(cudaStreamSynchronize(st));
        }
        std::vector<double> pts;
        if (!stored) pts = fetch_real(points, (size_t)n_seeds * 3, h->f64);
        h->n_states = 0;
        h->level_begin.clear();
        h->prev_resident = false;
        h->next_valid = false;
        h->n_incremental_levels = 0;
        h->shard_owned_states = 0;
        {
            size_t free_b = 0, total_b = 0;
            CK(cudaMemGetInfo(&free_b, &total_b));
            // the two resident level buffers may take what is free beyond a reserve for everything that still grows
            // during the march (key arena, CSR, visited set, digit planes of the widest level, chunk scratch: < 30 GB
            // at 8x512); 16 384 seeds make levels of 450 k states = 2 x 51 GB
            double gib = 128.0;
            if (const char *e = getenv("AM_B200_RESIDENT_GIB")) gib = atof(e);
            const size_t have = free_b + h->lvl_planes[0].cap + h->lvl_planes[1].cap;
            const size_t reserve = (size_t)36 << 30;
            h->resident_budget = std::min<size_t>((size_t)(gib * (1ull << 30)), have > 2 * reserve ? have - reserve : have / 2);
        }
        CK(cudaMemsetAsync(h->counters.p, 0, CNT_NUM * 8, st));
        memset(h->h_counters, 0, CNT_NUM * 8);
        if (h->tcap) CK(cudaMemsetAsync(h->table.p, 0xFF, (size_t)h->tcap * 8, st));
        h->table_sharded = h->p2p && h->shard_world > 1;
        CK(cudaMemsetAsync(h->level_cursor.p, 0, (size_t)(h->D + 3) * 4, st));   // re-armed by the kernels themselves,
        CK(cudaMemsetAsync(h->fs_ticket.p, 0, 64, st));                          // cleared here in case a march failed
        CK(cudaMemsetAsync(h->bal_loads.p, 0, XCHG_MAX_WORLD * 8, st));
        CK(cudaMemsetAsync(h->xcursor.p, 0, 64, st));
        insert_seeds(h, stored ? nullptr : states, pts.data(), n_seeds);
        h->stats.n_seeds = n_seeds;
        h->stats.n_unique_seeds = h->n_states;
        long long lb = 0, le = h->n_states;
        while (le > lb) {
            h->level_begin.push_back(lb);
            h->stats.max_level_states = std::max<int64_t>(h->stats.max_level_states, le - lb);
            h->trace_level = ((h->level_begin.size() - 1) % (size_t)h->trace_every) == 0;
            process_level(h, lb, le, iso, flip_insideout);
            lb = le;
            le = h->n_states;
        }
        h->level_begin.push_back(h->n_states);
        cudaEvent_t e_end = h->ev();
        CK(cudaStreamSynchronize(st));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e_begin, e_end));
        h->read_counters();
        resolve_spans(h);
        am_stats &s = h->stats;
        s.seconds_march = ms * 1e-3;
        s.n_states = h->n_states;
        s.n_faces = (int64_t)h->h_counters[CNT_FACES];
        s.n_corners = (int64_t)h->h_counters[CNT_CORNERS];
        s.n_levels = (int64_t)h->level_begin.size() - 1;
        s.n_candidates = (int64_t)h->h_counters[CNT_CANDIDATES];
        s.n_unbounded = (int64_t)h->h_counters[CNT_UNBOUNDED];
        s.n_overflow = (int64_t)h->h_counters[CNT_OVERFLOW];
        s.n_over_vertmax = (int64_t)h->h_counters[CNT_OVER_VERTMAX];
        s.n_inconsistent = (int64_t)h->h_counters[CNT_INCONSISTENT];
        s.n_tensors_reloaded = h->stats_layers_reloaded;
        s.seconds_host_wait = h->host_wait_s;
        s.seconds_host_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - host_t0).count();
        h->has_march = true;
    } catch (const CudaFail &f) {
        h->err = "am_march: " + f.msg;
        return f.code;
    }
    return AM_OK;
}

int am_combine(am_handle *h, double scale, const double center[3])
{
    if (!h) return AM_ERR_ARG;
    if (!h->has_march) {
        h->err = "AnalyticMarching must be done first!";
        return AM_ERR_STATE;
    }
    if (!center) {
        h->err = "am_combine: null center";
        return AM_ERR_ARG;
    }
    try {
        cudaStream_t st = h->stream;
        ensure_full_table(h);
        const long long nC = h->stats.n_corners;
        const long long nS = h->n_states;
        DevBuf &owner = h->cmb_owner, &flag = h->cmb_flag, &vid = h->cmb_vid, &cvid = h->cmb_cvid, &verts = h->cmb_verts;
        h->h_face_off.resize((size_t)nS + 1);
        h->h_face_off[0] = 0;
        if (nS) CK(cudaMemcpyAsync(h->h_face_off.data(), h->face_off.p, (size_t)(nS + 1) * 8, cudaMemcpyDeviceToHost, st));
        h->h_vertices.resize(0);
        h->h_corner_vid.resize((size_t)nC);
        unsigned long long *cnt = h->counters.as<unsigned long long>();
        long long nV = 0;
        if (nC > 0) {
            owner.reserve((size_t)nC * 8); flag.reserve((size_t)nC * 4); vid.reserve((size_t)nC * 4);
            cvid.reserve((size_t)nC * 4);
            StitchArgs sa{};
            sa.keys = h->keys.as<uint32_t>(); sa.hsum = h->hsum.as<unsigned long long>();
            sa.face_off = h->face_off.as<long long>(); sa.face_edges = h->face_edges.as<int>();
            sa.kw = h->kw; sa.kw4 = h->kw4; sa.L = h->L; sa.n_states = (int)nS;
            sa.table = TableRef{h->table.as<unsigned long long>(), h->tcap - 1};
            sa.owner = owner.as<long long>(); sa.counters = cnt;
            const int G = h->G;
            const unsigned gb = (unsigned)((nS * G + 255) / 256);
            h->dispatch_group([&](auto g) { stitch_owner_kernel<decltype(g)::value><<<gb, 256, 0, st>>>(sa); });
            ++h->stats.n_launches;
            const unsigned cb = (unsigned)((nC + 255) / 256);
            owner_flags_kernel<<<cb, 256, 0, st>>>(owner.as<long long>(), nC, flag.as<uint32_t>());
            ++h->stats.n_launches;
            CK(cudaGetLastError());
            h->scan(flag.as<uint32_t>(), vid.as<uint32_t>(), (int)nC, cnt + CNT_VERTS);
            h->read_counters();
            nV = (long long)h->h_counters[CNT_VERTS];
            verts.reserve(std::max<size_t>((size_t)nV * 24, 24));
            index_corners_kernel<<<cb, 256, 0, st>>>(owner.as<long long>(), vid.as<uint32_t>(), flag.as<uint32_t>(), nC,
                                                    h->face_xyz.as<double>(), scale, center[0], center[1], center[2],
                                                    cvid.as<int>(), verts.as<double>());
            ++h->stats.n_launches;
            CK(cudaGetLastError());
            h->h_vertices.resize((size_t)nV * 3);
            CK(cudaMemcpyAsync(h->h_vertices.data(), verts.p, (size_t)nV * 24, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(h->h_corner_vid.data(), cvid.p, (size_t)nC * 4, cudaMemcpyDeviceToHost, st));
        }
        CK(cudaStreamSynchronize(st));
        h->read_counters();
        h->stats.n_vertices = nV;
        h->stats.n_stitch_miss = (int64_t)h->h_counters[CNT_STITCH_MISS];
        h->has_mesh = true;
    } catch (const CudaFail &f) {
        h->err = "am_combine: " + f.msg;
        return AM_ERR_CUDA;
    }
    return AM_OK;
}

int am_export(am_handle *h, const char *path, int is_polymesh, int is_float32)
{
    if (!h) return AM_ERR_ARG;
    if (!h->has_mesh) {
        h->err = "CombineMesh must be done first!";
        return AM_ERR_STATE;
    }
    if (!path) {
        h->err = "am_export: null path";
        return AM_ERR_ARG;
    }
    const long long nS = h->n_states;
    const long long nV = h->stats.n_vertices;
    // The body is assembled by a few host threads: every thread owns a contiguous range of states, counts its
    // faces first (so that its byte offset is known) and then writes its records in place.
    const int T = (int)std::max<long long>(1, std::min<long long>({(long long)std::thread::hardware_concurrency(), 16LL,
                                                                    nS / 65536 + 1}));
    std::vector<long long> cF(T + 1, 0), cT(T + 1, 0), cC(T + 1, 0);       // faces, fan triangles, corners per range
    auto range = [&](int t) { return std::make_pair(nS * t / T, nS * (t + 1) / T); };
    auto run_threads = [&](auto &&fn) {
        std::vector<std::thread> th;
        for (int t = 1; t < T; ++t) th.emplace_back(fn, t);
        fn(0);
        for (auto &x : th) x.join();
    };
    run_threads([&](int t) {
        const auto r = range(t);
        long long f = 0, tr = 0, c = 0;
        for (long long s = r.first; s < r.second; ++s) {
            const long long k = h->h_face_off[s + 1] - h->h_face_off[s];
            if (k >= 3) { ++f; tr += k - 2; c += k; }
        }
        cF[t + 1] = f; cT[t + 1] = tr; cC[t + 1] = c;
    });
    for (int t = 0; t < T; ++t) { cF[t + 1] += cF[t]; cT[t + 1] += cT[t]; cC[t + 1] += cC[t]; }
    const long long nF = cF[T], nT = cT[T];
    // header bytes exactly as reference backend/inc/polymesh.h:368-377
    const char *ft = is_float32 ? "float" : "double";
    std::string head = "ply\nformat binary_little_endian 1.0\nelement vertex " + std::to_string(nV) + "\nproperty " + ft +
                       " x\nproperty " + ft + " y\nproperty " + ft + " z\nelement face " +
                       std::to_string(is_polymesh ? nF : nT) + "\nproperty list uchar int vertex_index\nend_header\n";
    const size_t vbytes = (size_t)nV * 3 * (is_float32 ? 4 : 8);
    const size_t fbytes = is_polymesh ? (size_t)nF + (size_t)cC[T] * 4 : (size_t)nT * 13;
    const size_t total = head.size() + vbytes + fbytes;
    std::unique_ptr<unsigned char[]> buf(new unsigned char[total]);           // not zero-filled: every byte is written
    memcpy(buf.get(), head.data(), head.size());
    unsigned char *vout = buf.get() + head.size();
    unsigned char *fout = vout + vbytes;
    run_threads([&](int t) {
        // vertices: this thread's share of the coordinate array
        const size_t n3 = (size_t)nV * 3, v0 = n3 * t / T, v1 = n3 * (t + 1) / T;
        if (is_float32) {
            float *o = reinterpret_cast<float *>(vout);
            for (size_t i = v0; i < v1; ++i) o[i] = (float)h->h_vertices[i];
        } else if (v1 > v0) {
            memcpy(vout + v0 * 8, h->h_vertices.data() + v0, (v1 - v0) * 8);
        }
        // faces of the states in this thread's range
        const auto r = range(t);
        unsigned char *w = fout + (is_polymesh ? (size_t)cF[t] + (size_t)cC[t] * 4 : (size_t)cT[t] * 13);
        for (long long s = r.first; s < r.second; ++s) {
            const long long fo = h->h_face_off[s];
            const int k = (int)(h->h_face_off[s + 1] - fo);
            if (k < 3) continue;
            const int *idx = h->h_corner_vid.data() + fo;
            if (is_polymesh) {
                *w++ = (unsigned char)k;
                memcpy(w, idx, (size_t)k * 4);
                w += (size_t)k * 4;
            } else {
                for (int j = 0; j + 2 < k; ++j) {   // fan (0, j+1, j+2), reference polymesh.h:405-416
                    *w++ = 3;
                    memcpy(w, idx, 4);
                    memcpy(w + 4, idx + j + 1, 4);
                    memcpy(w + 8, idx + j + 2, 4);
                    w += 12;
                }
            }
        }
    });
    // one sequential write: measured faster than 16 concurrent pwrite() calls into the same file (0.22 vs 0.30 s for the
    // 512 MB mesh of the 8x512 network -- writers to one inode serialise in the file system)
    FILE *f = fopen(path, "wb");
    if (!f) {
        h->err = std::string("am_export: cannot open ") + path;
        return AM_ERR_IO;
    }
    const size_t wr = fwrite(buf.get(), 1, total, f);
    fclose(f);
    if (wr != total) {
        h->err = std::string("am_export: short write to ") + path;
        return AM_ERR_IO;
    }
    return AM_OK;
}

int am_get_stats(const am_handle *h, am_stats *out)
{
    if (!h || !out) return AM_ERR_ARG;
    *out = h->stats;
    return AM_OK;
}
int am_key_words(const am_handle *h) { return h ? h->kw : 0; }
int am_state_len(const am_handle *h) { return h ? h->L : 0; }

int am_copy_states(const am_handle *h, uint32_t *keys, int64_t *face_off, int32_t *parent, int32_t *via_edge)
{
    if (!h || !h->has_march) return AM_ERR_STATE;
    const size_t n = (size_t)h->n_states;
    cudaError_t e = cudaSuccess;
    if (keys && n) e = cudaMemcpy(keys, h->keys.p, n * h->kw * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && face_off) e = cudaMemcpy(face_off, h->face_off.p, (n + 1) * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && parent && n) e = cudaMemcpy(parent, h->parent.p, n * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && via_edge && n) e = cudaMemcpy(via_edge, h->via.p, n * 4, cudaMemcpyDeviceToHost);
    return e == cudaSuccess ? AM_OK : AM_ERR_CUDA;
}

int am_copy_faces(const am_handle *h, int32_t *edge_ids, double *xyz)
{
    if (!h || !h->has_march) return AM_ERR_STATE;
    const size_t n = (size_t)h->stats.n_corners;
    cudaError_t e = cudaSuccess;
    if (edge_ids && n) e = cudaMemcpy(edge_ids, h->face_edges.p, n * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && xyz && n) e = cudaMemcpy(xyz, h->face_xyz.p, n * 24, cudaMemcpyDeviceToHost);
    return e == cudaSuccess ? AM_OK : AM_ERR_CUDA;
}

int am_copy_mesh(const am_handle *h, double *vertices, int32_t *face_sizes, int32_t *face_index)
{
    if (!h || !h->has_mesh) return AM_ERR_STATE;
    if (vertices && !h->h_vertices.empty()) memcpy(vertices, h->h_vertices.data(), h->h_vertices.size() * 8);
    if (face_index && !h->h_corner_vid.empty()) memcpy(face_index, h->h_corner_vid.data(), h->h_corner_vid.size() * 4);
    if (face_sizes) {
        size_t f = 0;
        for (long long s = 0; s < h->n_states; ++s) {
            const long long k = h->h_face_off[s + 1] - h->h_face_off[s];
            if (k >= 3) face_sizes[f++] = (int32_t)k;
        }
    }
    return AM_OK;
}

int am_digest(am_handle *h, uint64_t out[8])
{
    if (!h || !out) return AM_ERR_ARG;
    if (!h->has_march) {
        h->err = "am_digest: AnalyticMarching must be done first!";
        return AM_ERR_STATE;
    }
    try {
        cudaStream_t st = h->stream;
        DevBuf &acc = h->digest_acc;
        acc.reserve(8 * 8);
        CK(cudaMemsetAsync(acc.p, 0, 8 * 8, st));
        unsigned long long *a = acc.as<unsigned long long>();
        const long long nS = h->n_states, nC = h->stats.n_corners;
        const unsigned grid = (unsigned)h->num_sms * 8;
        if (nS > 0) {
            digest_words_kernel<<<grid, 256, 0, st>>>(h->keys.p, nS * h->kw, 4, 0x1000000000ull, a + 0);
            digest_words_kernel<<<grid, 256, 0, st>>>(h->face_off.p, nS + 1, 8, 0x2000000000ull, a + 1);
            if (nC > 0) {
                digest_words_kernel<<<grid, 256, 0, st>>>(h->face_edges.p, nC, 4, 0x3000000000ull, a + 2);
                digest_words_kernel<<<grid, 256, 0, st>>>(h->face_xyz.p, nC * 3, 8, 0x4000000000ull, a + 3);
            }
            digest_states_kernel<<<grid, 256, 0, st>>>(h->keys.as<uint32_t>(), h->kw, (h->L + 31) / 32,
                                                       h->face_off.as<long long>(), h->face_edges.as<int>(),
                                                       h->face_xyz.as<double>(), nS, a + 4);
            CK(cudaGetLastError());
        }
        unsigned long long host[8];
        CK(cudaMemcpyAsync(host, acc.p, 8 * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (int i = 0; i < 6; ++i) out[i] = host[i];
        out[6] = (uint64_t)nS;
        out[7] = (uint64_t)nC;
    } catch (const CudaFail &f) {
        h->err = "am_digest: " + f.msg;
        return AM_ERR_CUDA;
    }
    return AM_OK;
}

int am_edge_incidence(am_handle *h, int64_t out[4])
{
    if (!h || !out) return AM_ERR_ARG;
    if (!h->has_march) {
        h->err = "am_edge_incidence: AnalyticMarching must be done first!";
        return AM_ERR_STATE;
    }
    try {
        cudaStream_t st = h->stream;
        DevBuf &acc = h->digest_acc;
        acc.reserve(8 * 8);
        CK(cudaMemsetAsync(acc.p, 0, 8 * 8, st));
        const long long nS = h->n_states;
        ensure_full_table(h);
        if (nS > 0) {
            StitchArgs sa{};
            sa.keys = h->keys.as<uint32_t>(); sa.hsum = h->hsum.as<unsigned long long>();
            sa.face_off = h->face_off.as<long long>(); sa.face_edges = h->face_edges.as<int>();
            sa.kw = h->kw; sa.kw4 = h->kw4; sa.L = h->L; sa.n_states = (int)nS;
            sa.table = TableRef{h->table.as<unsigned long long>(), h->tcap - 1};
            sa.owner = nullptr; sa.counters = h->counters.as<unsigned long long>();
            const int G = h->G;
            const unsigned gb = (unsigned)((nS * G + 255) / 256);
            unsigned long long *a = acc.as<unsigned long long>();
            h->dispatch_group([&](auto g) { edge_incidence_kernel<decltype(g)::value><<<gb, 256, 0, st>>>(sa, a); });
            CK(cudaGetLastError());
        }
        unsigned long long host[4];
        CK(cudaMemcpyAsync(host, acc.p, 4 * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (int i = 0; i < 4; ++i) out[i] = (int64_t)host[i];
    } catch (const CudaFail &f) {
        h->err = "am_edge_incidence: " + f.msg;
        return AM_ERR_CUDA;
    }
    return AM_OK;
}

int am_gather_states(const am_handle *h, const int64_t *ids, int64_t n, uint32_t *keys, int32_t *counts, int32_t *edges,
                     double *xyz, int32_t *parent, int32_t *via_edge, double *seedpt)
{
    if (!h || !h->has_march || !ids || n < 0) return AM_ERR_STATE;
    cudaError_t e = cudaSuccess;
    for (int64_t i = 0; i < n && e == cudaSuccess; ++i) {
        const int64_t s = ids[i];
        if (s < 0 || s >= h->n_states) return AM_ERR_ARG;
        long long fo[2] = {0, 0};
        e = cudaMemcpy(fo, h->face_off.as<long long>() + s, 16, cudaMemcpyDeviceToHost);
        const int k = (int)(fo[1] - fo[0]);
        if (k < 0 || k > VSLOTS) return AM_ERR_STATE;
        if (counts) counts[i] = k;
        if (e == cudaSuccess && keys)
            e = cudaMemcpy(keys + (size_t)i * h->kw, h->keys.as<uint32_t>() + (size_t)s * h->kw, (size_t)h->kw * 4,
                           cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && edges && k)
            e = cudaMemcpy(edges + (size_t)i * VSLOTS, h->face_edges.as<int>() + fo[0], (size_t)k * 4, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && xyz && k)
            e = cudaMemcpy(xyz + (size_t)i * VSLOTS * 3, h->face_xyz.as<double>() + fo[0] * 3, (size_t)k * 24,
                           cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && parent) e = cudaMemcpy(parent + i, h->parent.as<int>() + s, 4, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && via_edge) e = cudaMemcpy(via_edge + i, h->via.as<int>() + s, 4, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && seedpt)
            e = cudaMemcpy(seedpt + (size_t)i * 4, h->seedpt.as<double>() + (size_t)s * 4, 32, cudaMemcpyDeviceToHost);
    }
    return e == cudaSuccess ? AM_OK : AM_ERR_CUDA;
}

int am_debug_planes(am_handle *h, const uint8_t *states, int64_t n, double iso, void *planes_out, void *equ_out)
{
    if (!h || !states || n < 1) return AM_ERR_ARG;
    if (!h->weights_loaded) {
        h->err = "am_debug_planes: load weights first";
        return AM_ERR_STATE;
    }
    try {
        cudaStream_t st = h->stream;
        h->xstates.reserve((size_t)n * h->L, 0, false);
        CK(cudaMemcpyAsync(h->xstates.p, states, (size_t)n * h->L,
                           classify(states) == PK_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
        h->xkeys.reserve((size_t)n * h->kw * 4, 0, false);
        const long long tot = n * h->kw;
        pack_states_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(h->xstates.as<uint8_t>(), (int)n, h->L, h->kw,
                                                                         h->xkeys.as<uint32_t>());
        ++h->stats.n_launches;
        CK(cudaGetLastError());
        h->ensure_chunk_scratch((size_t)n);
        h->ev_used = 0;
        h->spans.clear();
        h->trace_level = true;
        h->span_scale = 1;
        h->compose_chunk(h->xkeys.as<uint32_t>(), (int)n, iso, h->planes.as<double>(), nullptr, nullptr, (int)n, nullptr);
        CK(cudaStreamSynchronize(st));
        std::vector<double> p1((size_t)h->n1 * 4), pl((size_t)n * h->R * 4), eq((size_t)n * 4);
        CK(cudaMemcpy(p1.data(), h->P1.p, p1.size() * 8, cudaMemcpyDeviceToHost));
        if (!pl.empty()) CK(cudaMemcpy(pl.data(), h->planes.p, pl.size() * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(eq.data(), h->equ.p, eq.size() * 8, cudaMemcpyDeviceToHost));
        for (int64_t s = 0; s < n; ++s)
            for (int r = 0; r < h->L; ++r)
                for (int c = 0; c < 4; ++c) {
                    const double v = (r < h->n1) ? p1[4 * (size_t)r + c] : pl[((size_t)s * h->R + (r - h->n1)) * 4 + c];
                    const size_t o = ((size_t)s * h->L + r) * 4 + c;
                    if (h->f64) static_cast<double *>(planes_out)[o] = v;
                    else static_cast<float *>(planes_out)[o] = (float)v;
                }
        for (size_t i = 0; i < eq.size(); ++i) {
            if (h->f64) static_cast<double *>(equ_out)[i] = eq[i];
            else static_cast<float *>(equ_out)[i] = (float)eq[i];
        }
        resolve_spans(h);
    } catch (const CudaFail &f) {
        h->err = "am_debug_planes: " + f.msg;
        return AM_ERR_CUDA;
    }
    return AM_OK;
}

__global__ void fp64_probe_kernel(double *out, double a, double b, int iters)
{
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

double am_fp64_peak_tflops(void)
{
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1.0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1.0;
    const int threads = 512, blocks = sms * 2, iters = 20000;
    double *out = nullptr;
    if (cudaMalloc(&out, sizeof(double) * threads * blocks) != cudaSuccess) return -1.0;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(a);
        fp64_probe_kernel<<<blocks, threads>>>(out, 1.000001, 1e-9, iters);
        cudaEventRecord(b);
        if (cudaEventSynchronize(b) != cudaSuccess) { best = -1.f; break; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        if (r > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(out);
    if (best <= 0.f) return -1.0;
    return 2.0 * 8 * iters * (double)threads * blocks / (best * 1e-3) / 1e12;
}

int am_gemm_variant(const am_handle *h, int *split_digits)
{
    if (!h) return AM_ERR_ARG;
    if (split_digits) *split_digits = h->split_digits;
    return h->gemm_variant;
}

int am_kernel_profile(const am_handle *h, int kind, double *ms_total, int64_t *launches, double *flops)
{
    if (!h || kind < 0 || kind >= am_handle::N_KINDS) return AM_ERR_ARG;
    if (ms_total) *ms_total = h->kind_ms[kind];
    if (launches) *launches = h->kind_launches[kind];
    if (flops) *flops = h->kind_flops[kind];
    return AM_OK;
}

int am_compose_profile(const am_handle *h, double *ms_total, int64_t *launches, double *flops)
{
    if (!h) return AM_ERR_ARG;
    if (ms_total) *ms_total = h->gemm_ms;
    if (launches) *launches = h->gemm_launches;
    if (flops) *flops = h->gemm_flops;
    return AM_OK;
}

}  // extern "C"
