// polyio.cpp -- host-side helpers for polygon-mesh files (part of libam_b200.so).
// Native counterpart of the record loops in reference backend/libpolytools/src/polylib.cpp:134-268
// (load_ply) and :349-393 (save_ply): the face stream of a binary PLY is a sequence of
// variable-length records [uchar k][k x int32][optional 3 x uchar colour]; numpy cannot slice it
// without a Python loop, so the two loops live here.  The Python class analyticmesh_b200.polymesh.PolyMesh
// does everything else with arrays.
#include <cstdint>
#include <cstring>

extern "C" {

// pass 1 (indices == nullptr): only counts[] and *n_indices are produced.  Returns bytes consumed, or -1
// if the body is truncated / capacity too small.
int64_t am_ply_parse_faces(const uint8_t *body, int64_t nbytes, int64_t n_faces, int has_colors, int32_t *counts,
                           int32_t *indices, int64_t cap, uint8_t *colors, int64_t *n_indices)
{
    int64_t pos = 0, ni = 0;
    for (int64_t f = 0; f < n_faces; ++f) {
        if (pos + 1 > nbytes) return -1;
        const int k = body[pos++];
        if (pos + 4LL * k + (has_colors ? 3 : 0) > nbytes) return -1;
        if (counts) counts[f] = k;
        if (indices) {
            if (ni + k > cap) return -1;
            memcpy(indices + ni, body + pos, 4LL * k);
        }
        ni += k;
        pos += 4LL * k;
        if (has_colors) {
            if (colors) memcpy(colors + 3 * f, body + pos, 3);
            pos += 3;
        }
    }
    if (n_indices) *n_indices = ni;
    return pos;
}

// inverse: writes the record stream, returns its size in bytes (out == nullptr: size only)
int64_t am_ply_pack_faces(const int32_t *counts, const int32_t *indices, int64_t n_faces, const uint8_t *colors,
                          uint8_t *out)
{
    int64_t pos = 0, ni = 0;
    for (int64_t f = 0; f < n_faces; ++f) {
        const int k = counts[f];
        if (out) {
            out[pos] = (uint8_t)k;
            memcpy(out + pos + 1, indices + ni, 4LL * k);
            if (colors) memcpy(out + pos + 1 + 4LL * k, colors + 3 * f, 3);
        }
        pos += 1 + 4LL * k + (colors ? 3 : 0);
        ni += k;
    }
    return pos;
}

}  // extern "C"
