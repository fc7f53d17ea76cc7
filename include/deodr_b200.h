/*
 * deodr_b200 - C-ABI of the B200 (sm_100a) differentiable rasteriser.
 *
 * This is the drop-in boundary for the reference's only native entry points,
 *     void renderScene  (Scene, double* image, double* z_buffer, double sigma, bool antialiaseError,
 *                        double* obs, double* err_buffer)                C++/DifferentiableRenderer.h:2717
 *     void renderScene_B(Scene, double* image, double* z_buffer, double* image_b, double sigma, bool,
 *                        double* obs, double* err_buffer, double* err_buffer_b)   C++/DifferentiableRenderer.h:2903
 * which the reference reaches through its Cython shim (deodr/differentiable_renderer_cython.pyx:50 renderSceneCpp,
 * :206 renderSceneBCpp; extern block :11-45).  Plain pointers and sizes only; no torch / numpy types.
 *
 * Two levels:
 *   - deodr_b200_render_host / deodr_b200_render_b_host take the reference's own `struct Scene`
 *     (DifferentiableRenderer.h:56-90: HOST pointers, fp64) and have the reference semantics (image / z_buffer
 *     fully overwritten; gradients ACCUMULATED into scene.*_b; image_b scaled in place is NOT reproduced - it is
 *     left untouched).  Host<->device copies happen inside the call.
 *   - deodr_b200_render / deodr_b200_render_b work on DEVICE-resident buffers in the canonical layout
 *     (fp64 ij / depths / uv / z_buffer, fp32 colours / shade / texture / image / gradients, int32 face ids) on a
 *     caller-supplied CUDA stream; this is what the PyTorch surface uses.
 *
 * All functions return 0 on success or a DEODR_B200_E* code; deodr_b200_last_error() gives the message of the last
 * failure on the calling thread (the reference throws `const char*` instead: DifferentiableRenderer.h:2664-2715,
 * 810, 2924).
 */
#ifndef DEODR_B200_H
#define DEODR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    DEODR_B200_OK = 0,
    DEODR_B200_EINVAL = 1,      /* null pointer / bad size / index out of range (checkSceneValid, DR.h:2664) */
    DEODR_B200_EUNSUPPORTED = 2,/* backward with perspective_correct or without backface_culling (DR.h:810, 2924) */
    DEODR_B200_ECUDA = 3,       /* CUDA runtime failure */
    DEODR_B200_ENOMEM = 4
};

/* Device-resident scene ("SceneView"): the fields of struct Scene (DR.h:56-90) in the canonical device layout. */
typedef struct DeodrSceneView {
    const uint32_t *faces;            /* [T,3] */
    const uint32_t *faces_uv;         /* [T,3] */
    const double *ij;                 /* [V,2] col 0 = x (image column), col 1 = y (image row) */
    const double *depths;             /* [V] */
    const double *uv;                 /* [Nuv,2] texel coordinates, col 0 = texture column */
    const float *colors;              /* [V,C] */
    const float *shade;               /* [V] */
    const uint8_t *edgeflags;         /* [T,3] */
    const uint8_t *textured;          /* [T] */
    const uint8_t *shaded;            /* [T] */
    const float *texture;             /* [Ht,Wt,C] */
    const float *background_image;    /* [H,W,C] or NULL */
    const float *background_color;    /* [C] (device) or NULL; exactly one of the two backgrounds is set */
    int32_t nb_triangles, nb_vertices, nb_uv;
    int32_t height, width, nb_colors;
    int32_t texture_height, texture_width;
    int32_t clockwise, backface_culling, strict_edge, perspective_correct, integer_pixel_centers;
} DeodrSceneView;

/* Device-resident gradient slots (scene.*_b of DR.h:82-86), fp32, ACCUMULATED into.  Any pointer may be NULL. */
typedef struct DeodrGrads {
    float *ij_b;       /* [V,2] */
    float *colors_b;   /* [V,C] */
    float *uv_b;       /* [Nuv,2] */
    float *shade_b;    /* [V] */
    float *texture_b;  /* [Ht,Wt,C]  (summed: the reference's `=` defect DR.h:621-624 is not reproduced) */
} DeodrGrads;

/* The reference's own scene struct (DR.h:56-90) with HOST pointers; flags are one byte each like C++ bool. */
typedef struct DeodrHostScene {
    const uint32_t *faces;
    const uint32_t *faces_uv;
    const double *depths;
    const double *uv;
    const double *ij;
    const double *shade;
    const double *colors;
    const uint8_t *edgeflags;
    const uint8_t *textured;
    const uint8_t *shaded;
    int32_t nb_triangles;
    int32_t nb_vertices;
    int32_t clockwise;
    int32_t backface_culling;
    int32_t nb_uv;
    int32_t height;
    int32_t width;
    int32_t nb_colors;
    const double *texture;
    int32_t texture_height;
    int32_t texture_width;
    const double *background_image;
    const double *background_color;
    double *uv_b;
    double *ij_b;
    double *shade_b;
    double *colors_b;
    double *texture_b;
    int32_t strict_edge;
    int32_t perspective_correct;
    int32_t integer_pixel_centers;
} DeodrHostScene;

/* Opaque per-device workspace: tile lists, sorted silhouette edges, staging buffers.  Not thread-safe: use one
 * workspace per stream.  The forward pass leaves in it the state (tile edge lists) the backward pass replays. */
typedef struct DeodrWorkspace DeodrWorkspace;

int deodr_b200_workspace_create(DeodrWorkspace **ws, int device);
void deodr_b200_workspace_destroy(DeodrWorkspace *ws);
/* bytes of device memory currently held by the workspace */
int64_t deodr_b200_workspace_bytes(const DeodrWorkspace *ws);
/* number of kernels launched by the library on behalf of this workspace since creation */
int64_t deodr_b200_workspace_launches(const DeodrWorkspace *ws);

/* Forward pass on device buffers (replaces renderScene, DR.h:2717, antialiaseError = false).
 *   image   [H,W,C] fp32, fully overwritten
 *   z_buffer[H,W]   fp64, fully overwritten (bit-identical to the reference; +inf = background)
 *   owner   [H,W]   int32, adjoint-owner triangle per pixel (-1 = background); state consumed by the backward pass
 *   face_id [H,W]   int32 or NULL: forward owner (lowest index at the minimum z), i.e. rint() of the reference's
 *                   render_deferred face-id channel
 * `stream` is a cudaStream_t.  No host synchronisation other than the two size read-backs of the binning. */
int deodr_b200_render(DeodrWorkspace *ws, const DeodrSceneView *scene, double sigma, float *image, double *z_buffer,
                      int32_t *owner, int32_t *face_id, void *stream);

/* Adjoint pass on device buffers (replaces renderScene_B, DR.h:2903, antialiaseError = false).
 * Must follow deodr_b200_render on the same workspace, scene and sigma.  image_b [H,W,C] fp32 is read-only.
 * Gradients are accumulated (+=) into `grads`. */
int deodr_b200_render_b(DeodrWorkspace *ws, const DeodrSceneView *scene, double sigma, const double *z_buffer,
                        const int32_t *owner, const float *image_b, const DeodrGrads *grads, void *stream);

/* Reference-shaped host entry points (fp64 host buffers in, fp64 host buffers out). */
int deodr_b200_render_host(DeodrWorkspace *ws, const DeodrHostScene *scene, double *image, double *z_buffer,
                           double sigma, int antialiase_error, const double *obs, double *err_buffer);
int deodr_b200_render_b_host(DeodrWorkspace *ws, const DeodrHostScene *scene, double *image, double *z_buffer,
                             double *image_b, double sigma, int antialiase_error, const double *obs,
                             double *err_buffer, double *err_buffer_b);

/* Zero-fills `n` host buffers (ptrs[i], bytes[i]) with the host path's copy threads: what Scene2D.clear_gradients
 * (deodr/differentiable_renderer.py:599-610, five numpy `fill(0)` calls = 30+ MB on a 1M-triangle scene) does before
 * every adjoint call, at memory speed instead of one core's. */
int deodr_b200_host_zero(DeodrWorkspace *ws, void *const *ptrs, const int64_t *bytes, int n);

/* Index-range validation of a device scene (checkSceneValid, DR.h:2703-2714); synchronises the stream. */
int deodr_b200_check_scene(DeodrWorkspace *ws, const DeodrSceneView *scene, void *stream);

/* Per-kernel device timing with CUDA events recorded on the launching stream (for bench.py's roofline).
 * enable: pre-creates `max_records` event pairs (0 disables and frees them).  Every instrumented launch group of the
 * following render / render_b calls then records one (phase, start, stop) triple until the pool is exhausted.
 * collect: synchronises the recorded events, writes up to `capacity` (phase, milliseconds) pairs in launch order,
 * resets the pool and returns the number of records written (negative on error). */
enum {
    DEODR_B200_PH_BIN_COUNT = 0,    /* memset + (index check) + k_bin_count + k_scan_tiles */
    DEODR_B200_PH_EDGE_ORDER = 1,   /* k_rank_edges + k_scatter_edges (or CUB radix sorts) + k_edge_records */
    DEODR_B200_PH_BIN_FILL = 2,     /* k_bin_fill */
    DEODR_B200_PH_EDGE_TILE_SORT = 3,/* k_sort_tile_edges */
    DEODR_B200_PH_TILE_Z = 4,       /* k_tile_z: z-buffer + owner ids */
    DEODR_B200_PH_SHADE = 5,        /* k_shade: colour of every pixel */
    DEODR_B200_PH_EDGE_FWD = 6,     /* k_edge_fwd: ordered silhouette-edge overdraw */
    DEODR_B200_PH_SMALL_BWD = 7,    /* k_small_tri_bwd: triangle-parallel interior adjoint */
    DEODR_B200_PH_INTERIOR_BWD = 8, /* k_interior_bwd: pixel-parallel interior adjoint (large triangles) */
    DEODR_B200_PH_EDGE_BWD = 9,     /* accumulator memset + k_raster_bwd: tiles with silhouette edges */
    DEODR_B200_PH_EDGE_FINALIZE = 10,/* k_finalize_edges */
    DEODR_B200_PH_COUNT = 11
};
int deodr_b200_timing_enable(DeodrWorkspace *ws, int max_records);
int deodr_b200_timing_collect(DeodrWorkspace *ws, int32_t *phase, float *ms, int capacity);
const char *deodr_b200_phase_name(int phase);

const char *deodr_b200_last_error(void);
const char *deodr_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DEODR_B200_H */
