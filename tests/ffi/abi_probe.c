/* TEST INFRASTRUCTURE (tests/test_cabi.py): include/deodr_b200.h seen by a C compiler.
 *
 *   abi_probe layout          prints "<struct>.<field> <offset>" and "<struct> <size>" for every struct of the header
 *   abi_probe call            a plain-C caller of the boundary: creates a workspace and renders two triangles through
 *                             deodr_b200_render_host exactly as a cgo / JNI / Cython stub would; prints the status
 *
 * Compiled as C99 with -pedantic -Werror: the header must not need C++.  */
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#include "deodr_b200.h"

#define FIELD(S, F) printf(#S "." #F " %zu\n", offsetof(S, F))
#define SIZE(S) printf(#S " %zu\n", sizeof(S))

static void layout(void) {
    FIELD(DeodrSceneView, faces); FIELD(DeodrSceneView, faces_uv); FIELD(DeodrSceneView, ij);
    FIELD(DeodrSceneView, depths); FIELD(DeodrSceneView, uv); FIELD(DeodrSceneView, colors);
    FIELD(DeodrSceneView, shade); FIELD(DeodrSceneView, edgeflags); FIELD(DeodrSceneView, textured);
    FIELD(DeodrSceneView, shaded); FIELD(DeodrSceneView, texture); FIELD(DeodrSceneView, background_image);
    FIELD(DeodrSceneView, background_color); FIELD(DeodrSceneView, nb_triangles); FIELD(DeodrSceneView, nb_vertices);
    FIELD(DeodrSceneView, nb_uv); FIELD(DeodrSceneView, height); FIELD(DeodrSceneView, width);
    FIELD(DeodrSceneView, nb_colors); FIELD(DeodrSceneView, texture_height); FIELD(DeodrSceneView, texture_width);
    FIELD(DeodrSceneView, clockwise); FIELD(DeodrSceneView, backface_culling); FIELD(DeodrSceneView, strict_edge);
    FIELD(DeodrSceneView, perspective_correct); FIELD(DeodrSceneView, integer_pixel_centers);
    SIZE(DeodrSceneView);
    FIELD(DeodrGrads, ij_b); FIELD(DeodrGrads, colors_b); FIELD(DeodrGrads, uv_b); FIELD(DeodrGrads, shade_b);
    FIELD(DeodrGrads, texture_b);
    SIZE(DeodrGrads);
    FIELD(DeodrHostScene, faces); FIELD(DeodrHostScene, faces_uv); FIELD(DeodrHostScene, depths);
    FIELD(DeodrHostScene, uv); FIELD(DeodrHostScene, ij); FIELD(DeodrHostScene, shade); FIELD(DeodrHostScene, colors);
    FIELD(DeodrHostScene, edgeflags); FIELD(DeodrHostScene, textured); FIELD(DeodrHostScene, shaded);
    FIELD(DeodrHostScene, nb_triangles); FIELD(DeodrHostScene, nb_vertices); FIELD(DeodrHostScene, clockwise);
    FIELD(DeodrHostScene, backface_culling); FIELD(DeodrHostScene, nb_uv); FIELD(DeodrHostScene, height);
    FIELD(DeodrHostScene, width); FIELD(DeodrHostScene, nb_colors); FIELD(DeodrHostScene, texture);
    FIELD(DeodrHostScene, texture_height); FIELD(DeodrHostScene, texture_width);
    FIELD(DeodrHostScene, background_image); FIELD(DeodrHostScene, background_color); FIELD(DeodrHostScene, uv_b);
    FIELD(DeodrHostScene, ij_b); FIELD(DeodrHostScene, shade_b); FIELD(DeodrHostScene, colors_b);
    FIELD(DeodrHostScene, texture_b); FIELD(DeodrHostScene, strict_edge); FIELD(DeodrHostScene, perspective_correct);
    FIELD(DeodrHostScene, integer_pixel_centers);
    SIZE(DeodrHostScene);
    FIELD(DeodrViewIO, image); FIELD(DeodrViewIO, z_buffer); FIELD(DeodrViewIO, owner); FIELD(DeodrViewIO, face_id);
    FIELD(DeodrViewIO, barycentric); FIELD(DeodrViewIO, obs); FIELD(DeodrViewIO, err_buffer);
    FIELD(DeodrViewIO, image_b); FIELD(DeodrViewIO, err_buffer_b);
    SIZE(DeodrViewIO);
    SIZE(DeodrCamera);
    SIZE(DeodrMeshTopology);
}

/* two triangles on a 16 x 12 image, one colour channel: the fields are filled one by one, as the pyx does (pyx:117-171) */
static int call(void) {
    static const uint32_t faces[6] = {0, 1, 2, 3, 4, 5}, faces_uv[6] = {0, 0, 0, 0, 0, 0};
    static const double ij[12] = {1, 1, 12, 2, 3, 10, 4, 3, 14, 5, 6, 11}, depths[6] = {1, 1, 1, 2, 2, 2};
    static const double uv[2] = {0, 0}, shade[6] = {1, 1, 1, 1, 1, 1}, colors[6] = {0.1, 0.2, 0.3, 0.7, 0.8, 0.9};
    static const uint8_t edgeflags[6] = {1, 1, 1, 1, 1, 1}, textured[2] = {0, 0}, shaded[2] = {0, 0};
    static const double texture[1] = {0}, background[1] = {0.5};
    static double image[16 * 12], z_buffer[16 * 12];
    DeodrHostScene s;
    DeodrWorkspace *ws = NULL;
    int rc;
    memset(&s, 0, sizeof(s));
    s.faces = faces; s.faces_uv = faces_uv; s.depths = depths; s.uv = uv; s.ij = ij; s.shade = shade; s.colors = colors;
    s.edgeflags = edgeflags; s.textured = textured; s.shaded = shaded;
    s.nb_triangles = 2; s.nb_vertices = 6; s.clockwise = 0; s.backface_culling = 0; s.nb_uv = 1;
    s.height = 12; s.width = 16; s.nb_colors = 1;
    s.texture = texture; s.texture_height = 1; s.texture_width = 1;
    s.background_image = NULL; s.background_color = background;
    s.strict_edge = 1; s.perspective_correct = 0; s.integer_pixel_centers = 1;
    rc = deodr_b200_workspace_create(&ws, 0);
    if (rc != DEODR_B200_OK) {
        printf("workspace_create %d %s\n", rc, deodr_b200_last_error());
        return 0; /* reported, judged by the test */
    }
    rc = deodr_b200_render_host(ws, &s, image, z_buffer, 1.0, 0, NULL, NULL);
    printf("render_host %d %s\n", rc, rc ? deodr_b200_last_error() : "ok");
    if (rc == DEODR_B200_OK) {
        int covered = 0, i;
        double sum = 0;
        for (i = 0; i < 16 * 12; i++) {
            covered += z_buffer[i] < 1e300;
            sum += image[i];
        }
        printf("covered %d sum %.6f\n", covered, sum);
    }
    deodr_b200_workspace_destroy(ws);
    return 0;
}

int main(int argc, char **argv) {
    if (argc > 1 && !strcmp(argv[1], "layout")) { layout(); return 0; }
    if (argc > 1 && !strcmp(argv[1], "call")) return call();
    return 2;
}
