// Host-side state shared by the translation units of libdeodr_b200.so (kernels.cu: device API, host_api.cu: the
// reference-shaped host entry points).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

#include "../../include/deodr_b200.h"
#include "phases.h"

char *deodr_error_buffer();  // thread-local, 512 bytes (defined in kernels.cu)

static inline int set_error(int code, const char *fmt, const char *detail = "") {
    snprintf(deodr_error_buffer(), 512, fmt, detail);
    return code;
}

#define CUDA_TRY(expr)                                                                                \
    do {                                                                                              \
        cudaError_t err__ = (expr);                                                                   \
        if (err__ != cudaSuccess) {                                                                   \
            snprintf(deodr_error_buffer(), 512, "%s failed: %s", #expr, cudaGetErrorString(err__));   \
            return DEODR_B200_ECUDA;                                                                  \
        }                                                                                             \
    } while (0)

// growable device buffer
struct DevBuf {
    void *ptr = nullptr;
    size_t bytes = 0;
    int ensure(size_t need, int64_t *accounted) {
        if (need <= bytes) return DEODR_B200_OK;
        if (ptr) {
            CUDA_TRY(cudaFree(ptr));
            *accounted -= (int64_t)bytes;
            ptr = nullptr;
            bytes = 0;
        }
        size_t want = need + need / 4 + 256;
        CUDA_TRY(cudaMalloc(&ptr, want));
        bytes = want;
        *accounted += (int64_t)want;
        return DEODR_B200_OK;
    }
    template <class T>
    T *as() const { return static_cast<T *>(ptr); }
};

struct DeodrWorkspace {
    int device = 0;
    int64_t bytes = 0;
    int64_t launches = 0;
    // optional per-phase event timing (bench / profiling)
    // independent kernel chains run on two auxiliary streams forked from / joined to the caller's stream
    cudaStream_t aux[2] = {nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
    bool overlap = true;  // DEODR_B200_SERIAL=1 keeps every launch on the caller's stream (profiling, A/B)
    std::vector<cudaEvent_t> ev_start, ev_stop;
    std::vector<int> ev_phase;
    int ev_used = 0;
    int totals_seq = 0;          // sequence number of the flag k_scan_tiles raises in host_totals[16]
    int *host_totals = nullptr;  // pinned: [0] tri refs, [1] selected edges, [2] edge refs, [3] tie counter, [4] flags
    // forward state
    int tiles_x = 0, tiles_y = 0, num_tiles = 0;
    int num_edges = 0;           // silhouette edges of the last forward pass
    double sigma = -1;
    int fwd_valid = 0;
    int any_textured = 1;  // does the scene of the last forward hold a textured triangle? (k_bin_count raises the flag)
    int fwd_T = 0, fwd_H = 0, fwd_W = 0, fwd_C = 0;
    DevBuf zeroed;               // [scalars(8) | 6 per-tile int arrays], one memset per forward
    int *scal = nullptr, *edge_count_ptr = nullptr;  // views into `zeroed`
    deodr::TriBins bins;         // small (pre-masked records) / large (indices) tile lists of the last forward
    DevBuf small_offset, small_recs, tri_offset, tri_refs;
    DevBuf small_ids, large_ids;  // compacted lists of the drawn triangles (count pass), reused by the adjoint
    int num_small = 0, num_large = 0;
    DevBuf large_tiles, edge_tiles;
    int *edge_tiles_ptr = nullptr;  // the two-ended list the edge kernels walk (crowded tiles first)  // compact lists of the tiles with large triangles / silhouette edges
    int num_large_tiles = 0, num_edge_tiles = 0;
    int num_heavy_edge_tiles = 0;  // crowded tiles (more than one chunk of edges) at the front of the two-ended edge_tiles list
    DevBuf edge_recs;            // per-edge band stencils in far-to-near order (k_edge_records)
    DevBuf edge_rank, edge_ids, edge_ids_tmp, edge_keys_in, edge_keys_out, edge_sorted, cub_temp;
    DevBuf edge_offset, edge_refs_tmp, edge_refs;
    DevBuf edge_spans;  // x spans of every (tile, edge, row), written by k_edge_fwd and reused by k_raster_bwd
    DevBuf scalars;              // device ints: [0] tri total, [1] num selected, [2] edge total, [3] tie counter, [4] flags
    DevBuf tie_pairs;
    int tie_capacity = 0;
    DevBuf edge_acc;
    // host-path staging (canonical device copies of a DeodrHostScene)
    DevBuf h_faces, h_faces_uv, h_ij, h_depths, h_uv, h_colors, h_shade, h_edgeflags, h_textured, h_shaded, h_texture,
        h_background, h_image, h_z, h_owner, h_image_b, h_grads;
    struct HostPath *host = nullptr;  // pinned staging, copy threads, cache of the last host forward (host_api.cu)
};


void deodr_host_path_destroy(DeodrWorkspace *ws);  // host_api.cu

// deodr_b200_render with optional on-device index validation (checkSceneValid, DR.h:2703-2714)
int deodr_render_impl(DeodrWorkspace *ws, const DeodrSceneView *scene, double sigma, float *image, double *z_buffer,
                      int32_t *owner, int32_t *face_id, void *stream, bool check_indices);
