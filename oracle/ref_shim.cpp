// TEST INFRASTRUCTURE ONLY (oracle/_ref): C-ABI wrapper around the UNMODIFIED reference raster core.
//
// The reference header is NOT copied into this repository: it is included from where it lies
// (/root/reference/C++/DifferentiableRenderer.h, passed with -I by oracle/Makefile) and compiled into
// oracle/_ref/libdeodr_ref.so.  This file only adds extern "C" entry points so that ctypes can call
// renderScene (DifferentiableRenderer.h:2717) and renderScene_B (DifferentiableRenderer.h:2903) without Cython.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load the result.
// <climits>/<cmath>/<cstdlib> come first: the reference header relies on SHRT_MAX and on the double overload of
// abs() being visible (the Cython build gets both through Python.h / libstdc++).
#include <climits>
#include <cmath>
#include <cstdlib>

#include "DifferentiableRenderer.h"

#include <cstdint>
#include <cstring>

extern "C" {

// Mirror of the reference `struct Scene` (DifferentiableRenderer.h:56-90) with fixed-width flag types so
// that a ctypes.Structure can describe it.  `bool` is one byte under g++/x86-64.
struct RefSceneC {
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
};

static char g_ref_error[512] = "";

static Scene to_ref_scene(const RefSceneC *c) {
    static_assert(sizeof(bool) == 1, "bool must be one byte");
    Scene s;
    s.faces = const_cast<unsigned int *>(c->faces);
    s.faces_uv = const_cast<unsigned int *>(c->faces_uv);
    s.depths = const_cast<double *>(c->depths);
    s.uv = const_cast<double *>(c->uv);
    s.ij = const_cast<double *>(c->ij);
    s.shade = const_cast<double *>(c->shade);
    s.colors = const_cast<double *>(c->colors);
    s.edgeflags = reinterpret_cast<bool *>(const_cast<uint8_t *>(c->edgeflags));
    s.textured = reinterpret_cast<bool *>(const_cast<uint8_t *>(c->textured));
    s.shaded = reinterpret_cast<bool *>(const_cast<uint8_t *>(c->shaded));
    s.nb_triangles = c->nb_triangles;
    s.nb_vertices = c->nb_vertices;
    s.clockwise = c->clockwise != 0;
    s.backface_culling = c->backface_culling != 0;
    s.nb_uv = c->nb_uv;
    s.height = c->height;
    s.width = c->width;
    s.nb_colors = c->nb_colors;
    s.texture = const_cast<double *>(c->texture);
    s.texture_height = c->texture_height;
    s.texture_width = c->texture_width;
    s.background_image = const_cast<double *>(c->background_image);
    s.background_color = const_cast<double *>(c->background_color);
    s.uv_b = c->uv_b;
    s.ij_b = c->ij_b;
    s.shade_b = c->shade_b;
    s.colors_b = c->colors_b;
    s.texture_b = c->texture_b;
    s.strict_edge = c->strict_edge != 0;
    s.perspective_correct = c->perspective_correct != 0;
    s.integer_pixel_centers = c->integer_pixel_centers != 0;
    return s;
}

const char *deodr_ref_last_error(void) { return g_ref_error; }

int deodr_ref_render(const RefSceneC *scene, double *image, double *z_buffer, double sigma, int antialiase_error,
                     double *obs, double *err_buffer) {
    try {
        renderScene(to_ref_scene(scene), image, z_buffer, sigma, antialiase_error != 0, obs, err_buffer);
    } catch (const char *msg) {
        strncpy(g_ref_error, msg, sizeof(g_ref_error) - 1);
        return 1;
    }
    return 0;
}

int deodr_ref_render_b(const RefSceneC *scene, double *image, double *z_buffer, double *image_b, double sigma,
                       int antialiase_error, double *obs, double *err_buffer, double *err_buffer_b) {
    try {
        renderScene_B(to_ref_scene(scene), image, z_buffer, image_b, sigma, antialiase_error != 0, obs, err_buffer,
                      err_buffer_b);
    } catch (const char *msg) {
        strncpy(g_ref_error, msg, sizeof(g_ref_error) - 1);
        return 1;
    }
    return 0;
}
}
