/*
 * TEST INFRASTRUCTURE ONLY - CPU restatement ("port") of the DEODR raster hot path.
 *
 * This file restates, in plain C99 and with our own structure (setup records + row-span helpers shared by the
 * forward and the adjoint), the algorithm of the reference core C++/DifferentiableRenderer.h (DR.h below):
 * renderScene (DR.h:2717-2901) and renderScene_B (DR.h:2903-3135), both modes (antialiase_error on / off).  Every
 * function cites the DR.h lines it follows.  It keeps the reference's operation ORDER (no FMA: compile with
 * -ffp-contract=off) so that z-buffer, coverage, image and gradients are bit-identical to the compiled reference;
 * tests/test_oracle.py checks that against oracle/_ref and against the reference's own pinned SHA-256 vectors
 * (tests/test_render_mesh.py:34-73, tests/test_triangle_soup_fitting.py:29-68) -> parity is PINNED.
 *
 * Nothing in the product (deodr_b200/) links, loads or calls this file.  It is the checker used by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 *
 * Known reference behaviours restated on purpose:
 *  - bilinear_sample_B writes the texel adjoint with `=` (DR.h:621-624); `texfix != 0` switches to `+=`.
 *  - a pixel whose z ties exactly between several triangles is credited, in the adjoint, to the highest index
 *    (and to every tied textured triangle above the highest tied interpolated one), see DR.h:1024-1031, 1320.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    const uint32_t *faces, *faces_uv;
    const double *depths, *uv, *ij, *shade, *colors;
    const uint8_t *edgeflags, *textured, *shaded;
    int32_t nb_triangles, nb_vertices, clockwise, backface_culling, nb_uv, height, width, nb_colors;
    const double *texture;
    int32_t texture_height, texture_width;
    const double *background_image, *background_color;
    double *uv_b, *ij_b, *shade_b, *colors_b, *texture_b;
    int32_t strict_edge, perspective_correct, integer_pixel_centers;
} OracleScene; /* same layout as RefSceneC in ref_shim.cpp / struct Scene DR.h:56-90 */

static char g_error[256] = "";
static int g_texfix = 0;

const char *deodr_oracle_last_error(void) { return g_error; }
void deodr_oracle_set_texfix(int on) { g_texfix = on; }

static int fail(const char *msg) {
    strncpy(g_error, msg, sizeof(g_error) - 1);
    return 1;
}

/* ---------------------------------------------------------------------------------------------- 3x3 algebra */

/* DR.h:92-117: transposed cofactors scaled by the reciprocal determinant. */
static void inv3x3(const double *S, double *T) {
    T[0] = (S[4] * S[8] - S[7] * S[5]);
    T[3] = -(S[3] * S[8] - S[6] * S[5]);
    T[6] = (S[3] * S[7] - S[6] * S[4]);
    T[1] = -(S[1] * S[8] - S[7] * S[2]);
    T[4] = (S[0] * S[8] - S[6] * S[2]);
    T[7] = -(S[0] * S[7] - S[6] * S[1]);
    T[2] = (S[1] * S[5] - S[4] * S[2]);
    T[5] = -(S[0] * S[5] - S[3] * S[2]);
    T[8] = (S[0] * S[4] - S[3] * S[1]);
    double inv_det = 1 / (S[0] * T[0] + S[1] * T[3] + S[2] * T[6]);
    for (int k = 0; k < 9; k++) T[k] *= inv_det;
}

/* DR.h:124-232: adjoint of inv3x3; S_B is accumulated.  The cofactor adjoints are visited in the reference's
 * order (0,3,6,1,4,7,2,5,8) because S_B entries receive several contributions. */
static void inv3x3_adj(const double *S, double *S_B, const double *T_B) {
    double Tp[9], Tp_B[9];
    Tp[0] = (S[4] * S[8] - S[7] * S[5]);
    Tp[3] = -(S[3] * S[8] - S[6] * S[5]);
    Tp[6] = (S[3] * S[7] - S[6] * S[4]);
    Tp[1] = -(S[1] * S[8] - S[7] * S[2]);
    Tp[4] = (S[0] * S[8] - S[6] * S[2]);
    Tp[7] = -(S[0] * S[7] - S[6] * S[1]);
    Tp[2] = (S[1] * S[5] - S[4] * S[2]);
    Tp[5] = -(S[0] * S[5] - S[3] * S[2]);
    Tp[8] = (S[0] * S[4] - S[3] * S[1]);
    double inv_det = 1 / (S[0] * Tp[0] + S[1] * Tp[3] + S[2] * Tp[6]);
    double inv_det_b = 0;
    for (int k = 0; k < 9; k++) {
        Tp_B[k] = 0;
        inv_det_b += Tp[k] * T_B[k];
        Tp_B[k] += inv_det * T_B[k];
    }
    double t_B = inv_det_b * (-inv_det * inv_det);
    S_B[0] += Tp[0] * t_B; Tp_B[0] += S[0] * t_B;
    S_B[1] += Tp[3] * t_B; Tp_B[3] += S[1] * t_B;
    S_B[2] += Tp[6] * t_B; Tp_B[6] += S[2] * t_B;
    /* each row: {cofactor index, +a, +b, -c, -d, sign} for Tp = sign*(S[a]*S[b] - S[c]*S[d]) */
    static const int cof[9][5] = {
        {0, 4, 8, 7, 5}, {3, 3, 8, 6, 5}, {6, 3, 7, 6, 4}, {1, 1, 8, 7, 2}, {4, 0, 8, 6, 2},
        {7, 0, 7, 6, 1}, {2, 1, 5, 4, 2}, {5, 0, 5, 3, 2}, {8, 0, 4, 3, 1}};
    static const int negated[9] = {0, 1, 0, 1, 0, 1, 0, 1, 0};
    for (int r = 0; r < 9; r++) {
        int q = cof[r][0], a = cof[r][1], b = cof[r][2], c = cof[r][3], d = cof[r][4];
        double g = Tp_B[q];
        if (!negated[r]) {
            S_B[a] += S[b] * g;  S_B[b] += S[a] * g;  S_B[c] += -S[d] * g;  S_B[d] += -S[c] * g;
        } else {
            S_B[a] += -S[b] * g; S_B[b] += -S[a] * g; S_B[c] += S[d] * g;   S_B[d] += S[c] * g;
        }
    }
}

/* ------------------------------------------------------------------------------------- robust integer division */

/* DR.h:440-479: min(x_max, max(x_min, floor(a/b))) with the incremental fall-back for near-degenerate b. */
static short floor_div_clamped(double a, double b, int x_min, int x_max) {
    short x;
    if (fabs(b) * 32767 > fabs(a) + fabs(b)) {
        x = (short)floor(a / b);
        if (x < x_min) x = (short)x_min;
        if (x > x_max) x = (short)x_max;
    } else if (b > 0) {
        x = (short)x_min;
        while (((x + 1) * b <= a) && (x < x_max)) x++;
    } else {
        x = (short)x_min;
        while (((x + 1) * b >= a) && (x < x_max)) x++;
    }
    return x;
}

/* DR.h:481-519 */
static short ceil_div_clamped(double a, double b, int x_min, int x_max) {
    short x;
    if (fabs(b) * 32767 > fabs(a) + fabs(b)) {
        x = (short)ceil(a / b);
        if (x < x_min) x = (short)x_min;
        if (x > x_max) x = (short)x_max;
    } else if (b > 0) {
        x = (short)x_min;
        while (((x + 1) * b < a) && (x < x_max)) x++;
    } else {
        x = (short)x_min;
        while (((x + 1) * b > a) && (x < x_max)) x++;
    }
    return x;
}

/* ------------------------------------------------------------------------------------------- bilinear texture */

typedef struct {
    int i00, i10, i01, i11; /* offsets of the four texels (already multiplied by C) */
    double e0, e1;          /* fractional parts after clamping */
    int out0, out1;         /* axis was clamped */
} Tap;

/* DR.h:527-556 (also 568-602): clamp-to-edge tap for texel coordinate p = (column, row). */
static Tap texture_tap(const double *p, int tex_w, int tex_h, int C) {
    Tap t;
    int fp[2], size[2] = {tex_w, tex_h}, out[2] = {0, 0};
    double e[2];
    for (int k = 0; k < 2; k++) {
        fp[k] = (int)floor(p[k]);
        e[k] = p[k] - fp[k];
    }
    for (int k = 0; k < 2; k++) {
        if (fp[k] < 0) { out[k] = 1; fp[k] = 0; e[k] = 0; }
        if (fp[k] > size[k] - 2) { out[k] = 1; fp[k] = size[k] - 2; e[k] = 1; }
    }
    t.i00 = C * (fp[0] + tex_w * fp[1]);
    t.i10 = C * (fp[0] + 1 + tex_w * fp[1]);
    t.i01 = C * (fp[0] + tex_w * (fp[1] + 1));
    t.i11 = C * (fp[0] + 1 + tex_w * (fp[1] + 1));
    t.e0 = e[0]; t.e1 = e[1]; t.out0 = out[0]; t.out1 = out[1];
    return t;
}

/* DR.h:558-559 */
static void texture_fetch(const Tap *t, const double *I, int C, double *A) {
    for (int k = 0; k < C; k++)
        A[k] = ((1 - t->e0) * I[t->i00 + k] + t->e0 * I[t->i10 + k]) * (1 - t->e1) +
               ((1 - t->e0) * I[t->i01 + k] + t->e0 * I[t->i11 + k]) * t->e1;
}

/* DR.h:607-630: adjoint of the fetch w.r.t. p (p_B accumulated) and the texels (I_B). */
static void texture_fetch_adj(const Tap *t, const double *I, double *I_B, int C, const double *A_B, double *p_B) {
    double e_B[2] = {0, 0};
    for (int k = 0; k < C; k++) {
        double t1 = ((1 - t->e0) * I[t->i00 + k] + t->e0 * I[t->i10 + k]);
        double t2 = ((1 - t->e0) * I[t->i01 + k] + t->e0 * I[t->i11 + k]);
        e_B[1] += -A_B[k] * t1;
        e_B[1] += A_B[k] * t2;
        double t1_B = A_B[k] * (1 - t->e1);
        double t2_B = A_B[k] * t->e1;
        e_B[0] += t1_B * (I[t->i10 + k] - I[t->i00 + k]);
        e_B[0] += t2_B * (I[t->i11 + k] - I[t->i01 + k]);
        double w00 = (1 - t->e0) * (1 - t->e1) * A_B[k], w10 = t->e0 * (1 - t->e1) * A_B[k];
        double w01 = (1 - t->e0) * t->e1 * A_B[k], w11 = t->e0 * t->e1 * A_B[k];
        if (g_texfix) {
            I_B[t->i00 + k] += w00; I_B[t->i10 + k] += w10; I_B[t->i01 + k] += w01; I_B[t->i11 + k] += w11;
        } else { /* reference defect: last writer wins (DR.h:621-624) */
            I_B[t->i00 + k] = w00; I_B[t->i10 + k] = w10; I_B[t->i01 + k] = w01; I_B[t->i11 + k] = w11;
        }
    }
    if (!t->out0) p_B[0] += e_B[0];
    if (!t->out1) p_B[1] += e_B[1];
}

/* ---------------------------------------------------------------------------------------- triangle setup record */

typedef struct {
    double M[9];       /* bary_to_xy1 */
    double Minv[9];    /* xy1_to_bary */
    double eq[3][3];   /* edge equations a x + b y + c */
    int x_min, x_max;
    int y_begin[2], y_end[2];
    int left[2], right[2];
} TriSetup;

static void order3(const double v[3], double sv[3], short idx[3]) { /* DR.h:400-426 */
    for (int k = 0; k < 3; k++) { sv[k] = v[k]; idx[k] = (short)k; }
#define CSWAP(a, b) if (sv[a] > sv[b]) { double td = sv[a]; sv[a] = sv[b]; sv[b] = td; short ti = idx[a]; idx[a] = idx[b]; idx[b] = ti; }
    CSWAP(0, 1) CSWAP(0, 2) CSWAP(1, 2)
#undef CSWAP
}

static void edge_equation(double e[3], const double v1[2], const double v2[2], int cw) { /* DR.h:373-389 */
    if (cw) { e[0] = (v1[1] - v2[1]); e[1] = (v2[0] - v1[0]); }
    else    { e[0] = (v2[1] - v1[1]); e[1] = (v1[0] - v2[0]); }
    e[2] = -0.5 * (e[0] * (v1[0] + v2[0]) + e[1] * (v1[1] + v2[1]));
}

static double signed_area(const double V[3][2], int cw) { /* DR.h:391-398 */
    double ux = V[1][0] - V[0][0], uy = V[1][1] - V[0][1];
    double vx = V[2][0] - V[0][0], vy = V[2][1] - V[0][1];
    return 0.5 * (ux * vy - vx * uy) * (cw ? 1 : -1);
}

/* DR.h:633-739 */
static void tri_setup(const double V[3][2], int strict, TriSetup *s) {
    for (int v = 0; v < 3; v++)
        for (int d = 0; d < 2; d++) s->M[3 * d + v] = V[v][d];
    for (int v = 0; v < 3; v++) s->M[6 + v] = 1;
    inv3x3(s->M, s->Minv);
    int cw = signed_area(V, 1) > 0;
    edge_equation(s->eq[0], V[0], V[1], cw);
    edge_equation(s->eq[1], V[1], V[2], cw);
    edge_equation(s->eq[2], V[2], V[0], cw);
    double xu[3], yu[3], xs[3], ys[3];
    short xo[3], yo[3];
    for (int k = 0; k < 3; k++) { xu[k] = V[k][0]; yu[k] = V[k][1]; }
    order3(xu, xs, xo);
    order3(yu, ys, yo);
    s->x_min = strict ? (short)floor(xs[0]) : (short)ceil(xs[0]);
    s->x_max = (short)floor(xs[2]);
    s->y_begin[0] = strict ? (short)floor(ys[0]) + 1 : (short)ceil(ys[0]);
    s->y_end[0] = (short)floor(ys[1]);
    s->y_begin[1] = strict ? (short)floor(ys[1]) + 1 : (short)ceil(ys[1]);
    s->y_end[1] = (short)floor(ys[2]);
    int id = yo[0];
    if (s->eq[id % 3][0] > 0) { s->right[0] = (id + 2) % 3; s->left[0] = id % 3; }
    else                      { s->right[0] = id % 3;       s->left[0] = (id + 2) % 3; }
    id = yo[2];
    if (s->eq[id % 3][0] < 0) { s->right[1] = id % 3;       s->left[1] = (id + 2) % 3; }
    else                      { s->right[1] = (id + 2) % 3; s->left[1] = id % 3; }
}

/* DR.h:864-906: [x_begin, x_end] of row y for one half of the triangle. */
static void tri_row_span(const TriSetup *s, int half, short y, int width, int strict, short *x_begin, short *x_end) {
    const double *l = s->eq[s->left[half]], *r = s->eq[s->right[half]];
    short x_min = (short)s->x_min, x_max = (short)s->x_max, tmp;
    if (x_min < 0) x_min = 0;
    if (x_max > width - 1) x_max = (short)(width - 1);
    *x_begin = x_min;
    *x_end = x_max;
    double num = -(l[1] * y + l[2]);
    if (strict) tmp = (short)(1 + floor_div_clamped(num, l[0], x_min - 1, x_max));
    else        tmp = ceil_div_clamped(num, l[0], x_min - 1, x_max);
    if (tmp > *x_begin) *x_begin = tmp;
    num = -(r[1] * y + r[2]);
    tmp = floor_div_clamped(num, r[0], x_min - 1, x_max);
    if (tmp < *x_end) *x_end = tmp;
}

/* plane[3i+j] = sum_k attr[k][i] * Minv[3k+j]   (DR.h:779-785, 1081-1087; nv = 3 for faces, 2 for edges) */
static void attribute_planes(int n_attr, int nv, const double *const *attr, const double *w, const double *Minv,
                             double *plane) {
    for (int i = 0; i < n_attr; i++)
        for (int j = 0; j < 3; j++) {
            plane[3 * i + j] = 0;
            for (int k = 0; k < nv; k++) plane[3 * i + j] += (w ? attr[k][i] * w[k] : attr[k][i]) * Minv[k * 3 + j];
        }
}

/* R[i] = sum_j M[3j+i] * V[j]   (mul_vect_matrix3x3 DR.h:272-280, mul_matrix(1,2,3) DR.h:296-309) */
static void vec_times_rows(int nv, const double *V, const double *M, double R[3]) {
    for (int i = 0; i < 3; i++) {
        R[i] = 0;
        for (int j = 0; j < nv; j++) R[i] += M[3 * j + i] * V[j];
    }
}

static double row_value(const double *plane, short y) { /* dot_prod with t = (0, y, 1)  DR.h:359-365, 929-934 */
    double t[3] = {0, (double)y, 1}, R = 0;
    for (int i = 0; i < 3; i++) R += plane[i] * t[i];
    return R;
}

/* --------------------------------------------------------------------------------------------- forward triangles */

typedef struct {
    const OracleScene *sc;
    double *image, *z_buffer;
    int tex_size[2];
    /* antialiase_error mode (DR.h:2066-2618): the silhouette edges overdraw the squared residual |image - obs|^2 held in
     * `err` instead of the colours; NULL in the colour mode */
    const double *obs;
    double *err, *err_b;
} Ctx;

/* DR.h:741-794 + 908-972 (interpolated) and DR.h:1042-1092 + 1159-1258 (textured gouraud). */
static void draw_triangle(const Ctx *c, int k, const double V[3][2], const double Zv[3]) {
    const OracleScene *sc = c->sc;
    const int C = sc->nb_colors, W = sc->width, H = sc->height, strict = sc->strict_edge, persp = sc->perspective_correct;
    const uint32_t *face = &sc->faces[3 * k];
    const int textured = sc->textured[k] && sc->shaded[k];
    if (sc->textured[k] && !textured) return; /* DR.h:2798, 2813: drawn by neither branch */
    TriSetup s;
    tri_setup(V, strict, &s);
    double invZ[3], planeZ[3], planeL[3], planeUV[6];
    double *planeA = (double *)malloc(sizeof(double) * 3 * C), *row = (double *)malloc(sizeof(double) * C);
    const double *w = NULL;
    if (persp) { for (int i = 0; i < 3; i++) invZ[i] = 1 / Zv[i]; w = invZ; }
    vec_times_rows(3, persp ? invZ : Zv, s.Minv, planeZ);
    if (textured) {
        const uint32_t *fuv = &sc->faces_uv[3 * k];
        double shade[3], uvv[3][2];
        const double *uvp[3];
        for (int i = 0; i < 3; i++) {
            shade[i] = sc->shade[face[i]];
            if (persp) shade[i] = invZ[i] * shade[i]; /* elementwise_prod_vec3(inv_Z, Shade) DR.h:1066 */
            uvv[i][0] = sc->uv[fuv[i] * 2]; uvv[i][1] = sc->uv[fuv[i] * 2 + 1];
            uvp[i] = uvv[i];
        }
        vec_times_rows(3, shade, s.Minv, planeL);
        attribute_planes(2, 3, uvp, w, s.Minv, planeUV);
    } else {
        const double *col[3];
        for (int i = 0; i < 3; i++) col[i] = sc->colors + (size_t)face[i] * C;
        attribute_planes(C, 3, col, w, s.Minv, planeA);
    }
    for (int half = 0; half < 2; half++) {
        int y0 = s.y_begin[half], y1 = s.y_end[half];
        if (y0 < 0) y0 = 0;
        if (y1 > H - 1) y1 = H - 1;
        for (short y = (short)y0; y <= y1; y++) {
            double Z0y = row_value(planeZ, y), L0y = 0, UV0y[2] = {0, 0};
            if (textured) {
                L0y = row_value(planeL, y);
                UV0y[0] = row_value(planeUV, y);
                UV0y[1] = row_value(planeUV + 3, y);
            } else {
                for (int i = 0; i < C; i++) row[i] = row_value(planeA + 3 * i, y);
            }
            short xb, xe;
            tri_row_span(&s, half, y, W, strict, &xb, &xe);
            int idx = y * W + xb;
            for (short x = xb; x <= xe; x++, idx++) {
                double Z = Z0y + planeZ[0] * x;
                if (persp) Z = 1 / Z;
                if (Z < c->z_buffer[idx]) {
                    c->z_buffer[idx] = Z;
                    if (textured) {
                        double L = L0y + planeL[0] * x, UV[2];
                        for (int q = 0; q < 2; q++) UV[q] = UV0y[q] + planeUV[3 * q] * x;
                        if (persp) { L = L * Z; UV[0] = UV[0] * Z; UV[1] = UV[1] * Z; }
                        Tap tap = texture_tap(UV, c->tex_size[0], c->tex_size[1], C);
                        texture_fetch(&tap, sc->texture, C, row);
                        for (int i = 0; i < C; i++) c->image[(size_t)C * idx + i] = row[i] * L;
                    } else if (persp) {
                        for (int i = 0; i < C; i++) c->image[(size_t)C * idx + i] = (row[i] + planeA[3 * i] * x) * Z;
                    } else {
                        for (int i = 0; i < C; i++) c->image[(size_t)C * idx + i] = row[i] + planeA[3 * i] * x;
                    }
                }
            }
        }
    }
    free(planeA);
    free(row);
}

/* ------------------------------------------------------------------------------------------------ edge records */

typedef struct {
    double E[9];       /* edge_to_xy1 */
    double Einv[9];    /* xy1_to_edge: rows 0,1 = edge barycentrics, row 2 = signed distance along the normal */
    double transp[3];  /* row 2 / sigma */
    double ineq[12];
    double nt[2], inv_norm;
    int y_begin, y_end;
} EdgeSetup;

/* DR.h:1366-1460 */
static void edge_setup(const double V[2][2], int H, double sigma, int cw, EdgeSetup *s) {
    double n[2];
    if (cw) { n[0] = V[0][1] - V[1][1]; n[1] = V[1][0] - V[0][0]; }
    else    { n[0] = V[1][1] - V[0][1]; n[1] = V[0][0] - V[1][0]; }
    s->nt[0] = n[0]; s->nt[1] = n[1];
    s->inv_norm = 1 / sqrt(n[0] * n[0] + n[1] * n[1]);
    n[0] *= s->inv_norm; n[1] *= s->inv_norm;
    for (int v = 0; v < 2; v++)
        for (int d = 0; d < 2; d++) s->E[3 * d + v] = V[v][d];
    for (int d = 0; d < 2; d++) s->E[3 * d + 2] = n[d];
    s->E[6] = 1; s->E[7] = 1; s->E[8] = 0;
    inv3x3(s->E, s->Einv);
    for (int k = 0; k < 3; k++) s->transp[k] = (1 / sigma) * s->Einv[6 + k];
    for (int k = 0; k < 6; k++) s->ineq[k] = s->Einv[k];
    for (int j = 0; j < 3; j++) s->ineq[6 + j] = s->transp[j];
    for (int j = 0; j < 2; j++) s->ineq[9 + j] = -s->transp[j];
    s->ineq[11] = (1 - s->transp[2]);
    s->y_begin = H - 1;
    for (int k = 0; k < 2; k++)
        if (V[k][1] - sigma < s->y_begin) s->y_begin = (int)floor(V[k][1] - sigma) + 1;
    if (s->y_begin < 0) s->y_begin = 0;
    s->y_end = 0;
    for (int k = 0; k < 2; k++)
        if (V[k][1] + sigma > s->y_end) s->y_end = (int)floor(V[k][1] + sigma);
    if (s->y_end > H - 1) s->y_end = H - 1;
}

/* DR.h:2620-2648 */
static void edge_row_span(const double ineq[12], int width, int y, int *x_begin, int *x_end) {
    *x_begin = 0;
    *x_end = width - 1;
    for (int k = 0; k < 4; k++) {
        double num = -(ineq[3 * k + 1] * y + ineq[3 * k + 2]);
        if (ineq[3 * k] < 0) {
            short t = floor_div_clamped(num, ineq[3 * k], *x_begin - 1, *x_end + 1);
            if (t < *x_end) *x_end = t;
        } else {
            short t = (short)(1 + floor_div_clamped(num, ineq[3 * k], *x_begin - 1, *x_end + 1));
            if (t > *x_begin) *x_begin = t;
        }
    }
}

typedef struct {
    EdgeSetup s;
    double planeZ[3], planeL[3], planeUV[6];
    double *planeA; /* 3*C, interpolated mode only */
    int textured;
} EdgeRec;

static void edge_gather(const OracleScene *sc, int k, int n, double off, double V[2][2], double Zv[3], uint32_t vid[2],
                        uint32_t uvid[2]) {
    static const int sub[3][2] = {{1, 0}, {2, 1}, {0, 2}}; /* DR.h:2822 */
    const uint32_t *face = &sc->faces[3 * k], *fuv = &sc->faces_uv[3 * k];
    for (int i = 0; i < 2; i++) {
        vid[i] = face[sub[n][i]];
        uvid[i] = fuv[sub[n][i]];
        for (int j = 0; j < 2; j++) V[i][j] = sc->ij[vid[i] * 2 + j] - off;
        Zv[i] = sc->depths[vid[i]];
    }
    Zv[2] = 1; /* never used by non-perspective paths; the reference reads past the array here (DR.h:1563) */
}

/* planes of DR.h:1560-1586 (interpolated) / DR.h:1803-1835 (textured) */
static void edge_planes(const OracleScene *sc, const double V[2][2], const double Zv[3], const uint32_t vid[2],
                        const uint32_t uvid[2], int textured, double sigma, EdgeRec *r) {
    const int C = sc->nb_colors, persp = sc->perspective_correct;
    edge_setup(V, sc->height, sigma, sc->clockwise, &r->s);
    r->textured = textured;
    double invZ[2];
    const double *w = NULL;
    if (persp) { invZ[0] = 1 / Zv[0]; invZ[1] = 1 / Zv[1]; w = invZ; }
    vec_times_rows(2, persp ? invZ : Zv, r->s.Einv, r->planeZ);
    if (textured) {
        double shade[2], uvv[2][2];
        const double *uvp[2];
        for (int i = 0; i < 2; i++) {
            shade[i] = sc->shade[vid[i]];
            if (persp) shade[i] = invZ[i] * shade[i];
            uvv[i][0] = sc->uv[uvid[i] * 2]; uvv[i][1] = sc->uv[uvid[i] * 2 + 1];
            uvp[i] = uvv[i];
        }
        vec_times_rows(2, shade, r->s.Einv, r->planeL);
        attribute_planes(2, 2, uvp, w, r->s.Einv, r->planeUV);
    } else {
        const double *col[2] = {sc->colors + (size_t)vid[0] * C, sc->colors + (size_t)vid[1] * C};
        attribute_planes(C, 2, col, w, r->s.Einv, r->planeA);
    }
}

/* DR.h:1541-1649 and DR.h:1781-1907 */
static void draw_edge(const Ctx *c, int k, int n, double sigma, double off) {
    const OracleScene *sc = c->sc;
    const int C = sc->nb_colors, W = sc->width, persp = sc->perspective_correct;
    double V[2][2], Zv[3];
    uint32_t vid[2], uvid[2];
    EdgeRec r;
    r.planeA = (double *)malloc(sizeof(double) * 3 * C);
    double *row = (double *)malloc(sizeof(double) * C);
    edge_gather(sc, k, n, off, V, Zv, vid, uvid);
    edge_planes(sc, V, Zv, vid, uvid, sc->textured[k] && sc->shaded[k], sigma, &r);
    for (short y = (short)r.s.y_begin; y <= r.s.y_end; y++) {
        double T0y = row_value(r.s.transp, y), Z0y = row_value(r.planeZ, y), L0y = 0, UV0y[2] = {0, 0};
        if (r.textured) {
            L0y = row_value(r.planeL, y);
            UV0y[0] = row_value(r.planeUV, y);
            UV0y[1] = row_value(r.planeUV + 3, y);
        } else {
            for (int i = 0; i < C; i++) row[i] = row_value(r.planeA + 3 * i, y);
        }
        int xb, xe;
        edge_row_span(r.s.ineq, W, y, &xb, &xe);
        int idx = y * W + xb;
        for (short x = (short)xb; x <= xe; x++, idx++) {
            double Z = Z0y + r.planeZ[0] * x;
            if (persp) Z = 1 / Z;
            if (Z < c->z_buffer[idx]) {
                double T = T0y + r.s.transp[0] * x;
                double *px = c->image + (size_t)C * idx;
                const double *ob = c->err ? c->obs + (size_t)C * idx : NULL;
                double Err = 0;
                if (r.textured) {
                    double L = L0y + r.planeL[0] * x, UV[2];
                    for (int q = 0; q < 2; q++) UV[q] = UV0y[q] + r.planeUV[3 * q] * x;
                    if (persp) { L *= Z; UV[0] *= Z; UV[1] *= Z; }
                    Tap tap = texture_tap(UV, c->tex_size[0], c->tex_size[1], C);
                    double *A = (double *)malloc(sizeof(double) * C);
                    texture_fetch(&tap, sc->texture, C, A);
                    if (c->err)  /* DR.h:2180-2185 */
                        for (int i = 0; i < C; i++) { double diff = A[i] * L - ob[i]; Err += diff * diff; }
                    else
                        for (int i = 0; i < C; i++) { px[i] *= T; px[i] += (1 - T) * A[i] * L; }
                    free(A);
                } else if (c->err) {  /* DR.h:2457-2466 */
                    for (int i = 0; i < C; i++) {
                        double A = (row[i] + r.planeA[3 * i] * x);
                        if (persp) A *= Z;
                        double diff = A - ob[i];
                        Err += diff * diff;
                    }
                } else {
                    for (int i = 0; i < C; i++) {
                        px[i] *= T;
                        double A = persp ? (row[i] + r.planeA[3 * i] * x) * Z : (row[i] + r.planeA[3 * i] * x);
                        px[i] += (1 - T) * A;
                    }
                }
                if (c->err) { c->err[idx] *= T; c->err[idx] += (1 - T) * Err; }  /* DR.h:2186-2187, 2467-2468 */
            }
        }
    }
    free(r.planeA);
    free(row);
}

/* --------------------------------------------------------------------------------- ordering shared by both passes */

typedef struct { double value; size_t index; } DepthKey;

/* The reference sorts with std::sort (introsort, unstable) and comparator value_left > value_right
 * (DR.h:2656-2662, 2781).  Equal keys are reference-undefined; we break ties by ascending triangle index. */
static int by_depth_desc(const void *a, const void *b) {
    const DepthKey *l = (const DepthKey *)a, *r = (const DepthKey *)b;
    if (l->value > r->value) return -1;
    if (l->value < r->value) return 1;
    return (l->index > r->index) - (l->index < r->index);
}

/* DR.h:2751-2781: depth-sum keys, front test, signed area on the UN-offset ij. */
static void classify(const OracleScene *sc, DepthKey *keys, double *area) {
    for (int k = 0; k < sc->nb_triangles; k++) {
        const uint32_t *face = &sc->faces[3 * k];
        keys[k].value = 0;
        keys[k].index = (size_t)k;
        int in_front = 1;
        for (int i = 0; i < 3; i++) {
            if (sc->depths[face[i]] < 0) in_front = 0;
            keys[k].value += sc->depths[face[i]];
        }
        if (in_front) {
            double V[3][2];
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 2; j++) V[i][j] = sc->ij[face[i] * 2 + j];
            area[k] = signed_area(V, sc->clockwise);
        } else {
            area[k] = 0;
        }
    }
    qsort(keys, (size_t)sc->nb_triangles, sizeof(DepthKey), by_depth_desc);
}

static int check_scene(const OracleScene *sc, int with_grads) { /* DR.h:2664-2715 */
    if (!sc->faces || !sc->faces_uv || !sc->depths || !sc->uv || !sc->ij || !sc->shade || !sc->colors ||
        !sc->edgeflags || !sc->textured || !sc->shaded || !sc->texture)
        return fail("scene array == NULL");
    if (!sc->background_image && !sc->background_color)
        return fail("scene.background == NULL and scene.background_color == NULL");
    if (with_grads && (!sc->uv_b || !sc->ij_b || !sc->shade_b || !sc->colors_b || !sc->texture_b))
        return fail("scene gradient array == NULL");
    for (int k = 0; k < sc->nb_triangles * 3; k++) {
        if (sc->faces[k] >= (uint32_t)sc->nb_vertices) return fail("scene.faces value greater than scene.nb_vertices");
        if (sc->faces_uv[k] >= (uint32_t)sc->nb_uv) return fail("scene.faces_uv value greater than scene.nb_uv");
    }
    return 0;
}

static void gather_face(const OracleScene *sc, int k, double off, double V[3][2], double Zv[3]) {
    const uint32_t *face = &sc->faces[3 * k];
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 2; j++) V[i][j] = sc->ij[face[i] * 2 + j] - off;
        Zv[i] = sc->depths[face[i]];
    }
}

/* DR.h:2717-2901 (antialiaseError == 0) */
int deodr_oracle_render(const OracleScene *sc, double *image, double *z_buffer, double sigma, int antialiase_error,
                        double *obs, double *err_buffer) {
    if (antialiase_error && (!obs || !err_buffer)) return fail("antialiase_error mode needs obs and err_buffer");
    if (check_scene(sc, 0)) return 1;
    const int P = sc->height * sc->width, C = sc->nb_colors, T = sc->nb_triangles;
    Ctx c = {sc, image, z_buffer, {sc->texture_width, sc->texture_height}, NULL, NULL, NULL};
    if (sc->background_image) memcpy(image, sc->background_image, sizeof(double) * (size_t)P * C);
    else
        for (int i = 0; i < P; i++)
            for (int k = 0; k < C; k++) image[(size_t)i * C + k] = sc->background_color[k];
    for (int i = 0; i < P; i++) z_buffer[i] = INFINITY;
    DepthKey *keys = (DepthKey *)malloc(sizeof(DepthKey) * (size_t)(T > 0 ? T : 1));
    double *area = (double *)malloc(sizeof(double) * (size_t)(T > 0 ? T : 1));
    classify(sc, keys, area);
    const float off_f = sc->integer_pixel_centers ? 0 : 0.5; /* a C float in the reference, DR.h:2783 */
    const double off = off_f;
    for (int k = 0; k < T; k++)
        if (area[k] > 0 || !sc->backface_culling) {
            double V[3][2], Zv[3];
            gather_face(sc, k, off, V, Zv);
            draw_triangle(&c, k, V, Zv);
        }
    if (antialiase_error) {  /* DR.h:2824-2837: the residual the edges then overdraw */
        for (int k = 0; k < P; k++) {
            double s = 0;
            for (int i = 0; i < C; i++) {
                double d = image[(size_t)C * k + i] - obs[(size_t)C * k + i];
                s += d * d;
            }
            err_buffer[k] = s;
        }
        c.obs = obs;
        c.err = err_buffer;
    }
    if (sigma > 0)
        for (int it = 0; it < T; it++) {
            int k = (int)keys[it].index;
            if (area[k] > 0)
                for (int n = 0; n < 3; n++)
                    if (sc->edgeflags[n + k * 3]) draw_edge(&c, k, n, sigma, off);
        }
    free(keys);
    free(area);
    return 0;
}

/* ------------------------------------------------------------------------------------------------------ adjoint */

/* DR.h:1462-1539: adjoint of edge_setup w.r.t. the two vertices. */
static void edge_setup_adj(const EdgeSetup *s, double V_B[2][2], double sigma, const double bary_B[6],
                           const double transp_B[3], int cw) {
    double E_B[9] = {0}, Einv_B[9] = {0};
    for (int k = 0; k < 3; k++) Einv_B[6 + k] += transp_B[k] * (1 / sigma);
    for (int k = 0; k < 6; k++) Einv_B[k] += bary_B[k];
    inv3x3_adj(s->E, E_B, Einv_B);
    for (int v = 0; v < 2; v++)
        for (int d = 0; d < 2; d++) V_B[v][d] += E_B[3 * d + v];
    double n_B[2] = {0, 0}, nt_B[2] = {0, 0}, inv_norm_B = 0;
    for (int d = 0; d < 2; d++) n_B[d] += E_B[3 * d + 2];
    for (int k = 0; k < 2; k++) {
        nt_B[k] += n_B[k] * s->inv_norm;
        inv_norm_B += n_B[k] * s->nt[k];
    }
    double nor_B = -inv_norm_B * (s->inv_norm * s->inv_norm);
    double nor_s_B = nor_B * 0.5 * s->inv_norm;
    nt_B[0] += 2 * s->nt[0] * nor_s_B;
    nt_B[1] += 2 * s->nt[1] * nor_s_B;
    if (cw) { V_B[0][1] += nt_B[0]; V_B[1][1] += -nt_B[0]; V_B[1][0] += nt_B[1]; V_B[0][0] += -nt_B[1]; }
    else    { V_B[0][1] += -nt_B[0]; V_B[1][1] += nt_B[0]; V_B[1][0] += -nt_B[1]; V_B[0][0] += nt_B[1]; }
}

/* DR.h:1651-1779 (interpolated) and DR.h:1909-2064 (textured): un-blend + adjoint of one silhouette edge. */
static void edge_adjoint(const Ctx *c, double *image_b, int k, int n, double sigma, double off) {
    const OracleScene *sc = c->sc;
    const int C = sc->nb_colors, W = sc->width;
    double V[2][2], Zv[3], V_B[2][2];
    uint32_t vid[2], uvid[2];
    EdgeRec r;
    r.planeA = (double *)calloc((size_t)3 * C, sizeof(double));
    double *planeA_B = (double *)calloc((size_t)3 * C, sizeof(double));
    double *row = (double *)calloc((size_t)C, sizeof(double)), *row_B = (double *)calloc((size_t)C, sizeof(double));
    double *A = (double *)calloc((size_t)C, sizeof(double)), *A_B = (double *)calloc((size_t)C, sizeof(double));
    edge_gather(sc, k, n, off, V, Zv, vid, uvid);
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) V_B[i][j] = sc->ij_b[vid[i] * 2 + j];
    edge_planes(sc, V, Zv, vid, uvid, sc->textured[k] && sc->shaded[k], sigma, &r);
    double bary_B[6] = {0}, transp_B[3] = {0}, planeL_B[3] = {0}, planeUV_B[6] = {0}, T_inc_B = 0;
    for (short y = (short)r.s.y_begin; y <= r.s.y_end; y++) {
        const double t[3] = {0, (double)y, 1};
        double T0y = row_value(r.s.transp, y), Z0y = row_value(r.planeZ, y), L0y = 0, UV0y[2] = {0, 0};
        double T0y_B = 0, L0y_B = 0, UV0y_B[2] = {0, 0};
        if (r.textured) {
            L0y = row_value(r.planeL, y);
            UV0y[0] = row_value(r.planeUV, y);
            UV0y[1] = row_value(r.planeUV + 3, y);
        } else {
            for (int i = 0; i < C; i++) { row[i] = row_value(r.planeA + 3 * i, y); row_B[i] = 0; }
        }
        int xb, xe;
        edge_row_span(r.s.ineq, W, y, &xb, &xe);
        int idx = y * W + xb;
        for (short x = (short)xb; x <= xe; x++, idx++) {
            double Z = Z0y + r.planeZ[0] * x;
            if (!(Z < c->z_buffer[idx])) continue;
            double T = T0y + r.s.transp[0] * x, T_B = 0;
            double *px = c->image + (size_t)C * idx, *px_b = image_b ? image_b + (size_t)C * idx : NULL;
            if (c->err) {
                /* antialiase_error mode: un-blend the residual, then the adjoint of Err = sum_k (A_k - obs_k)^2
                 * (DR.h:2296-2336 textured, DR.h:2542-2576 interpolated) */
                const double *ob = c->obs + (size_t)C * idx;
                double L = 1, L_B = 0, UV[2] = {0, 0}, UV_B[2] = {0, 0}, Err = 0;
                Tap tap;
                if (r.textured) {
                    L = L0y + r.planeL[0] * x;
                    for (int q = 0; q < 2; q++) UV[q] = UV0y[q] + r.planeUV[3 * q] * x;
                    tap = texture_tap(UV, c->tex_size[0], c->tex_size[1], C);
                    texture_fetch(&tap, sc->texture, C, A);
                    for (int i = 0; i < C; i++) A_B[i] = 0;
                    for (int i = 0; i < C; i++) { double diff = A[i] * L - ob[i]; Err += diff * diff; }
                } else {
                    for (int i = 0; i < C; i++) {
                        double Ai = row[i] + r.planeA[3 * i] * x;
                        double diff = Ai - ob[i];
                        Err += diff * diff;
                    }
                }
                double Err_B = 0;
                T_B += -Err * c->err_b[idx];
                Err_B += (1 - T) * c->err_b[idx];
                c->err[idx] -= (1 - T) * Err;
                c->err[idx] /= T;
                T_B += c->err_b[idx] * c->err[idx];
                c->err_b[idx] *= T;
                if (r.textured) {
                    for (int i = 0; i < C; i++) {
                        double diff = A[i] * L - ob[i];
                        double diff_B = 2 * diff * Err_B;
                        A_B[i] += diff_B * L;
                        L_B += diff_B * A[i];
                    }
                    texture_fetch_adj(&tap, sc->texture, sc->texture_b, C, A_B, UV_B);
                    for (int q = 0; q < 2; q++) { UV0y_B[q] += UV_B[q]; planeUV_B[3 * q] += UV_B[q] * x; }
                    L0y_B += L_B;
                    planeL_B[0] += x * L_B;
                } else {
                    for (int i = 0; i < C; i++) {
                        double Ai = row[i] + r.planeA[3 * i] * x;
                        double diff = Ai - ob[i];
                        double diff_B = 2 * diff * Err_B;
                        row_B[i] += diff_B;  /* A0y_B: accumulated and then DROPPED by the reference, see below */
                        planeA_B[3 * i] += x * diff_B;
                    }
                }
                T0y_B += T_B;
                T_inc_B += x * T_B;
                continue;
            }
            if (r.textured) {
                double L = L0y + r.planeL[0] * x, L_B = 0, UV[2], UV_B[2] = {0, 0};
                for (int q = 0; q < 2; q++) UV[q] = UV0y[q] + r.planeUV[3 * q] * x;
                Tap tap = texture_tap(UV, c->tex_size[0], c->tex_size[1], C);
                texture_fetch(&tap, sc->texture, C, A);
                for (int i = 0; i < C; i++) A_B[i] = 0;
                for (int i = 0; i < C; i++) {
                    T_B += -px_b[i] * A[i] * L;
                    A_B[i] += L * (1 - T) * px_b[i];
                    L_B += px_b[i] * (1 - T) * A[i];
                    px[i] = (px[i] - (1 - T) * A[i] * L) / T;
                    T_B += px_b[i] * px[i];
                    px_b[i] *= T;
                }
                texture_fetch_adj(&tap, sc->texture, sc->texture_b, C, A_B, UV_B);
                for (int q = 0; q < 2; q++) { UV0y_B[q] += UV_B[q]; planeUV_B[3 * q] += UV_B[q] * x; }
                L0y_B += L_B;
                planeL_B[0] += x * L_B;
            } else {
                for (int i = 0; i < C; i++) {
                    double Ai = row[i] + r.planeA[3 * i] * x;
                    T_B += -px_b[i] * Ai;
                    double Ai_B = (1 - T) * px_b[i];
                    px[i] = (px[i] - (1 - T) * Ai) / T;
                    T_B += px_b[i] * px[i];
                    px_b[i] *= T;
                    row_B[i] += Ai_B;
                    planeA_B[3 * i] += x * Ai_B;
                }
            }
            T0y_B += T_B;
            T_inc_B += x * T_B;
        }
        if (r.textured) {
            for (int q = 0; q < 3; q++) transp_B[q] += T0y_B * t[q];
            for (int i = 0; i < 2; i++)
                for (int q = 0; q < 3; q++) planeUV_B[q + 3 * i] += UV0y_B[i] * t[q];
            for (int q = 0; q < 3; q++) planeL_B[q] += L0y_B * t[q];
        } else {
            /* Reference defect kept on purpose (defect #2): rasterize_edge_interpolated_error_B never
             * back-propagates the per-row A0y_B into xy1_to_A_B (no mul_matrixNx3_vect_B after the x loop, DR.h:2577-2583),
             * so in antialiase_error mode the colour planes only receive their x-coefficient adjoint. */
            if (!c->err)
                for (int i = 0; i < C; i++)
                    for (int j = 0; j < 3; j++) planeA_B[3 * i + j] += row_B[i] * t[j];
            for (int q = 0; q < 3; q++) transp_B[q] += T0y_B * t[q];
        }
    }
    if (r.textured) {
        double uv_B[2][2], shade_B[2];
        for (int i = 0; i < 2; i++) {
            uv_B[i][0] = sc->uv_b[uvid[i] * 2]; uv_B[i][1] = sc->uv_b[uvid[i] * 2 + 1];
            shade_B[i] = sc->shade_b[vid[i]];
        }
        for (int i = 0; i < 2; i++)
            for (int j = 0; j < 3; j++)
                for (int q = 0; q < 2; q++) {
                    uv_B[q][i] += planeUV_B[3 * i + j] * r.s.Einv[q * 3 + j];
                    bary_B[q * 3 + j] += planeUV_B[3 * i + j] * sc->uv[uvid[q] * 2 + i];
                }
        /* mul_matrix_B(1,2,3,...) DR.h:311-333, 2056: planeL = Shade . Einv[rows 0,1] */
        for (int q = 0; q < 3; q++)
            for (int j = 0; j < 2; j++) {
                shade_B[j] += planeL_B[q] * r.s.Einv[j * 3 + q];
                bary_B[j * 3 + q] += planeL_B[q] * sc->shade[vid[j]];
            }
        transp_B[0] += T_inc_B;
        edge_setup_adj(&r.s, V_B, sigma, bary_B, transp_B, sc->clockwise);
        for (int i = 0; i < 2; i++) {
            sc->uv_b[uvid[i] * 2] = uv_B[i][0]; sc->uv_b[uvid[i] * 2 + 1] = uv_B[i][1];
        }
        for (int i = 0; i < 2; i++) sc->shade_b[vid[i]] = shade_B[i];
    } else {
        for (int i = 0; i < C; i++)
            for (int j = 0; j < 3; j++)
                for (int q = 0; q < 2; q++) {
                    sc->colors_b[(size_t)vid[q] * C + i] += planeA_B[3 * i + j] * r.s.Einv[q * 3 + j];
                    bary_B[q * 3 + j] += sc->colors[(size_t)vid[q] * C + i] * planeA_B[3 * i + j];
                }
        transp_B[0] += T_inc_B;
        edge_setup_adj(&r.s, V_B, sigma, bary_B, transp_B, sc->clockwise);
    }
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) sc->ij_b[vid[i] * 2 + j] = V_B[i][j];
    free(r.planeA); free(planeA_B); free(row); free(row_B); free(A); free(A_B);
}

/* DR.h:796-862 + 974-1040 (interpolated) and DR.h:1094-1157 + 1260-1364 (textured gouraud). */
static void triangle_adjoint(const Ctx *c, double *image_b, int k, double off) {
    const OracleScene *sc = c->sc;
    const int C = sc->nb_colors, W = sc->width, H = sc->height, strict = sc->strict_edge;
    const uint32_t *face = &sc->faces[3 * k], *fuv = &sc->faces_uv[3 * k];
    const int textured = sc->textured[k] && sc->shaded[k];
    if (sc->textured[k] && !textured) return;
    double V[3][2], Zv[3], V_B[3][2];
    gather_face(sc, k, off, V, Zv);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 2; j++) V_B[i][j] = sc->ij_b[face[i] * 2 + j];
    TriSetup s;
    tri_setup(V, strict, &s);
    double planeZ[3], planeL[3], planeUV[6], planeL_B[3] = {0}, planeUV_B[6] = {0}, Minv_B[9] = {0}, M_B[9] = {0};
    double *planeA = (double *)calloc((size_t)3 * C, sizeof(double)), *planeA_B = (double *)calloc((size_t)3 * C, sizeof(double));
    double *row_B = (double *)calloc((size_t)C, sizeof(double));
    double *A = (double *)calloc((size_t)C, sizeof(double)), *A_B = (double *)calloc((size_t)C, sizeof(double));
    double shade[3] = {0, 0, 0}, uvv[3][2];
    vec_times_rows(3, Zv, s.Minv, planeZ);
    if (textured) {
        const double *uvp[3];
        for (int i = 0; i < 3; i++) {
            shade[i] = sc->shade[face[i]];
            uvv[i][0] = sc->uv[fuv[i] * 2]; uvv[i][1] = sc->uv[fuv[i] * 2 + 1];
            uvp[i] = uvv[i];
        }
        vec_times_rows(3, shade, s.Minv, planeL);
        attribute_planes(2, 3, uvp, NULL, s.Minv, planeUV);
    } else {
        const double *col[3];
        for (int i = 0; i < 3; i++) col[i] = sc->colors + (size_t)face[i] * C;
        attribute_planes(C, 3, col, NULL, s.Minv, planeA);
    }
    for (int half = 0; half < 2; half++) {
        int y0 = s.y_begin[half], y1 = s.y_end[half];
        if (y0 < 0) y0 = 0;
        if (y1 > H - 1) y1 = H - 1;
        for (short y = (short)y0; y <= y1; y++) {
            const double t[3] = {0, (double)y, 1};
            double Z0y = row_value(planeZ, y), L0y = 0, UV0y[2] = {0, 0}, L0y_B = 0, UV0y_B[2] = {0, 0};
            if (textured) {
                L0y = row_value(planeL, y);
                UV0y[0] = row_value(planeUV, y);
                UV0y[1] = row_value(planeUV + 3, y);
            } else {
                for (int i = 0; i < C; i++) row_B[i] = 0;
            }
            short xb, xe;
            tri_row_span(&s, half, y, W, strict, &xb, &xe);
            int idx = y * W + xb;
            for (short x = xb; x <= xe; x++, idx++) {
                double Z = Z0y + planeZ[0] * x;
                if (Z != c->z_buffer[idx]) continue;
                double *px_b = image_b + (size_t)C * idx;
                if (textured) {
                    double L = L0y + planeL[0] * x, L_B = 0, UV[2], UV_B[2] = {0, 0};
                    for (int q = 0; q < 2; q++) UV[q] = UV0y[q] + planeUV[3 * q] * x;
                    Tap tap = texture_tap(UV, c->tex_size[0], c->tex_size[1], C);
                    texture_fetch(&tap, sc->texture, C, A);
                    for (int i = 0; i < C; i++) A_B[i] = 0;
                    for (int i = 0; i < C; i++) { A_B[i] += px_b[i] * L; L_B += px_b[i] * A[i]; }
                    texture_fetch_adj(&tap, sc->texture, sc->texture_b, C, A_B, UV_B);
                    for (int q = 0; q < 2; q++) { UV0y_B[q] += UV_B[q]; planeUV_B[3 * q] += UV_B[q] * x; }
                    L0y_B += L_B;
                    planeL_B[0] += x * L_B;
                } else {
                    for (int i = 0; i < C; i++) {
                        row_B[i] += px_b[i];
                        planeA_B[3 * i] += px_b[i] * x;
                        px_b[i] = 0;
                    }
                }
            }
            if (textured) {
                for (int i = 0; i < 2; i++)
                    for (int q = 0; q < 3; q++) planeUV_B[q + 3 * i] += UV0y_B[i] * t[q];
                for (int q = 0; q < 3; q++) planeL_B[q] += L0y_B * t[q];
            } else {
                for (int i = 0; i < C; i++)
                    for (int j = 0; j < 3; j++) planeA_B[3 * i + j] += row_B[i] * t[j];
            }
        }
    }
    if (textured) {
        double uv_B[3][2], shade_B[3];
        for (int i = 0; i < 3; i++) {
            shade_B[i] = sc->shade_b[face[i]];
            uv_B[i][0] = sc->uv_b[fuv[i] * 2]; uv_B[i][1] = sc->uv_b[fuv[i] * 2 + 1];
        }
        for (int i = 0; i < 2; i++)
            for (int j = 0; j < 3; j++)
                for (int q = 0; q < 3; q++) {
                    uv_B[q][i] += planeUV_B[3 * i + j] * s.Minv[q * 3 + j];
                    Minv_B[q * 3 + j] += planeUV_B[3 * i + j] * uvv[q][i];
                }
        for (int i = 0; i < 3; i++) /* mul_vect_matrix3x3_B DR.h:282-294 */
            for (int j = 0; j < 3; j++) {
                Minv_B[3 * j + i] += planeL_B[i] * shade[j];
                shade_B[j] += planeL_B[i] * s.Minv[3 * j + i];
            }
        inv3x3_adj(s.M, M_B, Minv_B);
        for (int v = 0; v < 3; v++)
            for (int d = 0; d < 2; d++) V_B[v][d] += M_B[3 * d + v];
        for (int i = 0; i < 3; i++) {
            sc->uv_b[fuv[i] * 2] = uv_B[i][0]; sc->uv_b[fuv[i] * 2 + 1] = uv_B[i][1];
        }
        for (int i = 0; i < 3; i++) sc->shade_b[face[i]] = shade_B[i];
    } else {
        for (int i = 0; i < C; i++)
            for (int j = 0; j < 3; j++)
                for (int q = 0; q < 3; q++) {
                    sc->colors_b[(size_t)face[q] * C + i] += planeA_B[3 * i + j] * s.Minv[q * 3 + j];
                    Minv_B[q * 3 + j] += sc->colors[(size_t)face[q] * C + i] * planeA_B[3 * i + j];
                }
        inv3x3_adj(s.M, M_B, Minv_B);
        for (int v = 0; v < 3; v++)
            for (int d = 0; d < 2; d++) V_B[v][d] += M_B[3 * d + v];
    }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 2; j++) sc->ij_b[face[i] * 2 + j] = V_B[i][j];
    free(planeA); free(planeA_B); free(row_B); free(A); free(A_B);
}

/* DR.h:2903-3135 (antialiaseError == 0).  `image` (AA undone) and `image_b` (scaled / zeroed) are MUTATED. */
int deodr_oracle_render_b(const OracleScene *sc, double *image, double *z_buffer, double *image_b, double sigma,
                          int antialiase_error, double *obs, double *err_buffer, double *err_buffer_b) {
    if (antialiase_error && (!obs || !err_buffer || !err_buffer_b))
        return fail("antialiase_error mode needs obs, err_buffer and err_buffer_b");
    if (check_scene(sc, 1)) return 1;
    if (!sc->backface_culling) return fail("You have to use backface_culling true if you ant to compute gradients");
    if (sc->perspective_correct)
        return fail("backward gradient propagation not supported yet with perspective_correct=True");
    const int T = sc->nb_triangles;
    Ctx c = {sc, image, z_buffer, {sc->texture_width, sc->texture_height}, NULL, NULL, NULL};
    if (antialiase_error) { c.obs = obs; c.err = err_buffer; c.err_b = err_buffer_b; }
    DepthKey *keys = (DepthKey *)malloc(sizeof(DepthKey) * (size_t)(T > 0 ? T : 1));
    double *area = (double *)malloc(sizeof(double) * (size_t)(T > 0 ? T : 1));
    classify(sc, keys, area);
    const float off_f = sc->integer_pixel_centers ? 0 : 0.5;
    const double off = off_f;
    if (sigma > 0)
        for (int it = T - 1; it >= 0; it--) {
            int k = (int)keys[it].index;
            if (area[k] > 0)
                for (int n = 2; n >= 0; n--)
                    if (sc->edgeflags[n + k * 3]) edge_adjoint(&c, antialiase_error ? NULL : image_b, k, n, sigma, off);
        }
    double *own_image_b = NULL;
    if (antialiase_error) {  /* DR.h:3054-3060: the adjoint of the residual the edges started from */
        const size_t n = (size_t)sc->width * sc->height * sc->nb_colors;
        own_image_b = (double *)malloc(sizeof(double) * (n > 0 ? n : 1));
        for (int k = 0; k < sc->width * sc->height; k++)
            for (int i = 0; i < sc->nb_colors; i++)
                own_image_b[(size_t)sc->nb_colors * k + i] =
                    -2 * (obs[(size_t)sc->nb_colors * k + i] - image[(size_t)sc->nb_colors * k + i]) * err_buffer_b[k];
        image_b = own_image_b;
    }
    c.err = NULL;  /* the triangle adjoints work on colours in both modes */
    for (int k = T - 1; k >= 0; k--)
        if (area[k] > 0) triangle_adjoint(&c, image_b, k, off);
    free(own_image_b);
    free(keys);
    free(area);
    return 0;
}
