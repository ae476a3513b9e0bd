/*
 * am_b200.h -- C ABI of the B200-native Analytic Marching engine (libam_b200.so).
 *
 * This is the drop-in boundary for the reference's native extension `cuam`
 * (reference backend/src/cuam.cpp:186-217: Init / AnalyticMarching / CombineMesh / ExportMesh /
 * Destroy).  Plain pointers and sizes only -- no torch types.  The pybind11 module
 * analyticmesh_b200/csrc/cuam_pybind.cpp keeps the reference's five names and keyword arguments
 * and forwards to these entry points; INTEGRATION.md shows the binding a reference maintainer adds.
 *
 * Conventions
 *   - every function returns AM_OK (0) or a negative error code; am_last_error() gives the text.
 *     (The reference prints and exit()s the interpreter on CUDA errors, backend/inc/utilities.h:73-96.)
 *   - data pointers may point to HOST or DEVICE memory; the library classifies them with
 *     cudaPointerGetAttributes and stages host data itself.
 *   - real type: double when the handle was created with is_f64 = 1, float otherwise.
 *   - matrices are row-major (out, in), exactly as torch's nn.Linear.weight.
 *   - bit j of a state <-> 32-bit word j/32, bit j%32 (reference backend/inc/states.h:63,76).
 */
#ifndef AM_B200_H
#define AM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct am_handle am_handle;

enum {
    AM_OK = 0,
    AM_ERR_ARG = -1,        /* malformed argument (shape, null pointer, float type) */
    AM_ERR_STATE = -2,      /* call order violated (e.g. am_combine before am_march) */
    AM_ERR_CUDA = -3,       /* a CUDA call failed */
    AM_ERR_IO = -4,         /* file could not be written */
    AM_ERR_CAPACITY = -5    /* an internal limit was hit (reported, never silent) */
};

/* counters of the last march; all monotone within one am_march call */
typedef struct am_stats {
    int64_t n_seeds;            /* seed states passed in */
    int64_t n_unique_seeds;     /* after de-duplication */
    int64_t n_states;           /* activation patterns visited (= processed) */
    int64_t n_faces;            /* states whose polygon has >= 3 vertices */
    int64_t n_corners;          /* sum of polygon sizes */
    int64_t n_levels;           /* BFS levels */
    int64_t n_candidates;       /* neighbour candidates generated (one per neuron edge) */
    int64_t n_unbounded;        /* polygons still touching the artificial bounding box (dropped) */
    int64_t n_overflow;         /* polygons that exceeded the 32-vertex working capacity (dropped) */
    int64_t n_over_vertmax;     /* polygons with more than 20 vertices (kept; the reference truncates) */
    int64_t n_inconsistent;     /* clip steps whose outside set was not one cyclic run */
    int64_t n_vertices;         /* unique vertices after am_combine */
    int64_t n_stitch_miss;      /* corners whose owner sibling lacked the matching corner */
    int64_t max_level_states;   /* widest BFS level */
    int64_t n_launches;         /* kernels launched by am_march (+ am_combine once it has run) */
    double  seconds_march;      /* device time of am_march (CUDA events) */
    double  seconds_compose;    /* ... spent in the affine-composition kernels */
    double  seconds_clip;       /* ... in the clipping kernel */
    double  seconds_frontier;   /* ... in neighbour enumeration + visited-set kernels */
    double  compose_flops;      /* algorithmic flops executed by the composition kernels */
    int64_t n_tensors_reloaded; /* weight matrices / transforms whose bytes changed since the previous call and were
                                   re-staged on the device (0 when the same network is marched again) */
    double  seconds_host_wait;  /* host wall time blocked in the per-level synchronisation (GPU-bound when close
                                   to seconds_host_total, launch-bound when small) */
    double  seconds_host_total; /* host wall time of am_march */
} am_stats;

/* replaces cuam.Init (reference backend/src/cuam.cpp:58-95, src/cuam_kernel.cu:128-148).
 * nodes: [3, n_1..n_D, 1]; arc_table: arc_rows = D rows of arc_cols ints, row = [k, src0, tm0, ...]. */
int am_create(am_handle **out, int is_f64, const int *nodes, int n_nodes, const int *arc_table,
              int arc_rows, int arc_cols, int num_extra_constraints);

/* replaces cuam.AnalyticMarching (reference backend/src/cuam.cpp:97-184, src/cuam_kernel.cu:30-123).
 * W, B: n_nodes-1 pointers; TM: n_tm pointers, tm_shapes: 2*n_tm ints ((0,0) = identity);
 * states: n_seeds x L bytes (0/1, torch.bool layout); points: n_seeds x 3 reals;
 * w_extra: n_extra x 3, b_extra: n_extra (inside is w.x + b < 0).  n_extra must equal the
 * value given to am_create (the reference does not check this).
 * states = points = NULL (n_seeds ignored): march from the seeds of the last am_seed_dichotomy call.
 * stream: a cudaStream_t (NULL = default stream).  Returns after the march has completed. */
int am_march(am_handle *h, const void *const *W, const void *const *B, const void *const *TM,
             const int *tm_shapes, int n_tm, const uint8_t *states, const void *points, int64_t n_seeds,
             const void *w_extra, const void *b_extra, int n_extra, double iso, int flip_insideout,
             void *stream);

/* replaces cuam.CombineMesh (reference src/cuam_kernel.cu:185-200): shared-vertex stitching +
 * indexing; v_out = v * scale + center. */
int am_combine(am_handle *h, double scale, const double center[3]);

/* replaces cuam.ExportMesh (reference src/cuam_kernel.cu:205-246, inc/polymesh.h:350-423):
 * binary little-endian PLY, polygons or fan triangles, float or double vertices. */
int am_export(am_handle *h, const char *path, int is_polymesh, int is_float32);

/* replaces cuam.Destroy (reference src/cuam_kernel.cu:153-179). NULL is allowed. */
void am_destroy(am_handle *h);

/* Sharded mode (new): ONE march spread over `world` GPUs, one process per GPU, every
 * process calling am_march with identical arguments.  Compose + clip of a state run on its owner
 * rank only (children inherit the owner of their parent, so the parent's plane rows are local);
 * once per BFS level the library hands the level's polygon scratch (device pointer, n_int32 32-bit
 * words, non-zero on exactly one rank per slot) to `fn`, which must sum it over all ranks in place
 * (e.g. ncclAllReduce / torch.distributed.all_reduce on an int32 view) ordered on `cuda_stream` (the
 * engine's stream: no host synchronisation is needed on either side) and return 0.  Frontier, visited
 * set and mesh are replicated and bit-identical. */
typedef int (*am_allreduce_fn)(void *user, void *device_ptr, int64_t n_int32, void *cuda_stream);
int am_set_shard(am_handle *h, int rank, int world, am_allreduce_fn fn, void *user);
/* Same, with the collective issued by the library itself: ncclAllReduce on the engine's stream
 * (libnccl.so.2 is bound with dlopen).  Rank 0 calls am_nccl_unique_id, the host broadcasts the 128
 * bytes (e.g. torch.distributed.broadcast), every rank calls am_set_shard_nccl (collective call). */
int am_nccl_unique_id(void *out128);
int am_set_shard_nccl(am_handle *h, int rank, int world, const void *unique_id128);

/* Same march, with the per-level exchange done by the engine's own kernels over NVLink peer memory
 * (csrc/xchg.cuh): every rank allocates one exchange block, the blocks are mapped into all processes with CUDA
 * IPC, polygons and winner masks are pushed with plain stores and ordered by device-side flag barriers; the
 * visited set is sharded by key hash.  No collective library and no host synchronisation inside a level.
 * `fn` is only used here, during set-up: it must gather `bytes` bytes from every rank into recv (rank order)
 * on every rank, e.g. torch.distributed.all_gather, and return 0.  Collective call. */
typedef int (*am_allgather_fn)(void *user, const void *send, void *recv, int64_t bytes);
int am_set_shard_p2p(am_handle *h, int rank, int world, am_allgather_fn fn, void *user);

/* The surface-point initialiser on the device (csrc/seeds.cuh): replaces `dichotomy` + `init_within_ball` +
 * `constraints_filter` + the state extraction of reference backend/main.py:252-326, 83-91, 70-80, 408-411.
 * try_pts_num trial points per round are drawn in the ball by a counter-based generator keyed by `seed`, points of
 * opposite sign of f - iso are paired without replacement, every pair is bisected until the mean |f - iso| is below
 * avg_eps (or iter_max bisections).  The init_num points and their packed activation keys stay on the device:
 * am_march(..., states = NULL, points = NULL, n_seeds = 0, ...) marches from them.  Deterministic in `seed`. */
typedef struct am_seed_report {
    int64_t n_points, rounds, iterations;
    double avg_abs_error, seconds;
} am_seed_report;
int am_seed_dichotomy(am_handle *h, const void *const *W, const void *const *B, const void *const *TM,
                      const int *tm_shapes, int n_tm, const void *w_extra, const void *b_extra, int n_extra,
                      double iso, int64_t init_num, int64_t try_pts_num, double ball_radius, int iter_max,
                      double avg_eps, uint64_t seed, am_seed_report *report);
/* the stored seeds: points [n][3] double, states [n][L] bytes (0/1); either may be NULL.  Host pointers. */
int am_copy_seeds(const am_handle *h, double *points, uint8_t *states_bool);
int64_t am_num_seeds(const am_handle *h);

int am_get_stats(const am_handle *h, am_stats *out);
const char *am_last_error(const am_handle *h);   /* h may be NULL: error of the last failed am_create */
int am_key_words(const am_handle *h);            /* 32-bit words per stored key (multiple of 4) */
int am_state_len(const am_handle *h);

/* ---- parity accessors (new; the reference exposes only the PLY) ------------------------------ */

/* per visited state, in state-id order (BFS order, deterministic):
 *   keys      [n_states][am_key_words()]
 *   face_off  [n_states + 1]  corner offsets (face i has face_off[i+1]-face_off[i] corners, 0 = no face)
 *   parent    [n_states]      state id that discovered it (-1 for seeds)
 *   via_edge  [n_states]      neuron whose bit was flipped (-1 for seeds)
 * any pointer may be NULL. Host pointers only. */
int am_copy_states(const am_handle *h, uint32_t *keys, int64_t *face_off, int32_t *parent, int32_t *via_edge);

/* per corner: edge_ids[c] = constraint carrying the segment corner c -> next corner; xyz[c][3] in
 * double regardless of the handle's real type (unscaled). Host pointers only. */
int am_copy_faces(const am_handle *h, int32_t *edge_ids, double *xyz);

/* after am_combine: vertices [n_vertices][3] (scaled, double), face_sizes [n_faces],
 * face_index [n_corners of exported faces]. Host pointers only. */
int am_copy_mesh(const am_handle *h, double *vertices, int32_t *face_sizes, int32_t *face_index);

/* the listed states only (ids: state ids, host pointer): keys [n][am_key_words()], counts [n] polygon sizes,
 * edges [n][32], xyz [n][32][3] (first counts[i] entries valid), parent / via_edge [n], seedpt [n][4] = the point
 * on the shared edge handed to the state (x, y, z, size hint).  Any output may be NULL.  Lets a test recompute
 * a sample of a march too large to copy (17.7 M states = 9 GB of keys) with the CPU oracle.
 * Reference counterpart: the popped batch of backend/src/cuam_kernel.cu:53-98. */
int am_gather_states(const am_handle *h, const int64_t *ids, int64_t n, uint32_t *keys, int32_t *counts,
                     int32_t *edges, double *xyz, int32_t *parent, int32_t *via_edge, double *seedpt);

/* Checksums of the last march, computed on the device (nothing is copied to the host):
 *   out[0..3]  positional checksums of keys / face_off / edge ids / vertex bits in state-id order -- equal
 *              between two runs iff numbering, topology and every coordinate bit are equal
 *   out[4]     order-independent checksum over the states of (key, edge loop): the region set + adjacency
 *   out[5]     the same including the vertex bits
 *   out[6], out[7]  n_states, n_corners */
int am_digest(am_handle *h, uint64_t out[8]);

/* Edge incidence of the polygon soup (exact visited-set lookups of the state across every edge):
 * out[0] edges on extra constraints (boundary), out[1] neuron edges matched by the neighbouring state's
 * polygon (counted from both sides: out[1]/2 shared edges), out[2] neighbour never visited, out[3] neighbour
 * visited but without that edge.  Closed 2-manifold <=> out[0] = out[2] = out[3] = 0. */
int am_edge_incidence(am_handle *h, int64_t out[4]);

/* debug/parity: run only the affine-composition kernels for n given states and return the
 * UNSIGNED rows planes[n][L][4] and the level planes equ[n][4] (host pointers, real type).
 * Weights are the ones of the last am_march / am_load_weights call. */
int am_load_weights(am_handle *h, const void *const *W, const void *const *B, const void *const *TM,
                    const int *tm_shapes, int n_tm);
int am_debug_planes(am_handle *h, const uint8_t *states, int64_t n, double iso, void *planes, void *equ);

/* timing hook for bench.py: average device milliseconds of the dominant composition kernel over
 * its launches in the last march, its launch count and the flops of those launches. */
int am_compose_profile(const am_handle *h, double *ms_total, int64_t *launches, double *flops);

/* same, per timed span kind: 0 composition launches of a chunk (= am_compose_profile), 1 compose phase,
 * 2 clip, 3 frontier, 4 the tcgen05 split-integer GEMM kernel alone, 5 its digit-extraction kernel,
 * 6 exchange barriers of the sharded march (= time waiting for the slowest rank), 7 polygon push, 8 unpack +
 * scan + CSR, 9 neighbour enumeration + insert, 11 finalize.
 * Kinds 4/5 are recorded per launch with CUDA events on the engine's stream (single-chain mode). */
int am_kernel_profile(const am_handle *h, int kind, double *ms_total, int64_t *launches, double *flops);
/* 0/1: FP64 tensor-core (DMMA) tiles, 2: tcgen05 int8 split-integer path; digits of the split (6..8) */
int am_gemm_variant(const am_handle *h, int *split_digits);

/* roofline denominator measured on the spot: TFLOP/s of a register-resident DFMA loop that fills
 * every SM of the current device (the same probe as tools/fp64_peak.cu). <= 0 on error. */
double am_fp64_peak_tflops(void);

/* ---- polygon-mesh file helpers (host only; native part of analyticmesh_b200.polymesh.PolyMesh, the
 * counterpart of reference backend/libpolytools/src/polylib.cpp:134-268, 349-393) ------------------ */

/* face records of a binary PLY body: [uchar k][k x int32][3 x uchar colour if has_colors].
 * indices == NULL: sizing pass (counts, *n_indices only).  Returns bytes consumed or -1. */
int64_t am_ply_parse_faces(const uint8_t *body, int64_t nbytes, int64_t n_faces, int has_colors, int32_t *counts,
                           int32_t *indices, int64_t cap, uint8_t *colors, int64_t *n_indices);
/* inverse; out == NULL: returns the size only */
int64_t am_ply_pack_faces(const int32_t *counts, const int32_t *indices, int64_t n_faces, const uint8_t *colors,
                          uint8_t *out);

#ifdef __cplusplus
}
#endif
#endif /* AM_B200_H */
