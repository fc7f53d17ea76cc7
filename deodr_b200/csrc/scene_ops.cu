// deodr_b200: the steps either side of the raster, on device buffers.
//
//   f1  Camera.project_points / project_points_backward   deodr/differentiable_renderer.py:341-438
//       Scene3D vertex luminosity + its adjoint             deodr/differentiable_renderer.py:814-850
//   f2  TriMeshAdjacencies.edge_on_silhouette               deodr/triangulated_mesh.py:153-166
//       compute_face_normals / compute_vertex_normals (+ adjoints)   deodr/triangulated_mesh.py:113-151
//
// They produce what the rasteriser consumes (ij, depths, colours, edge flags) and consume what it produces (ij_b,
// colors_b) without a round trip through the host: one thread per point / vertex / face, fp64 like the reference's
// numpy code (the flags are bit-exact: same products and differences as np.cross, never fused).
// All of them are elementwise or small gathers: HBM-bound, coalesced on the point index.
#include <cuda_runtime.h>

#include <cstring>

#include "../../include/deodr_b200.h"
#include "rmath.h"
#include "workspace.h"

namespace {

constexpr int BLOCK = 256;
inline int grid_of(size_t n) { return (int)((n + BLOCK - 1) / BLOCK); }

struct Cam {
    double R[9], t[3], K00, K01, K02, K10, K11, K12, k1, k2, p1, p2, k3;
    int distorted;
};

Cam cam_of(const DeodrCamera *c) {
    Cam k;
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) k.R[3 * i + j] = c->extrinsic[4 * i + j];
        k.t[i] = c->extrinsic[4 * i + 3];
    }
    k.K00 = c->intrinsic[0]; k.K01 = c->intrinsic[1]; k.K02 = c->intrinsic[2];
    k.K10 = c->intrinsic[3]; k.K11 = c->intrinsic[4]; k.K12 = c->intrinsic[5];
    k.distorted = c->has_distortion;
    k.k1 = c->distortion[0]; k.k2 = c->distortion[1]; k.p1 = c->distortion[2]; k.p2 = c->distortion[3];
    k.k3 = c->distortion[4];
    return k;
}

// world_to_camera (:290-292): p . R^T + t
__device__ __forceinline__ void to_camera(const Cam &c, const double *p, double out[3]) {
    for (int i = 0; i < 3; i++) out[i] = p[0] * c.R[3 * i] + p[1] * c.R[3 * i + 1] + p[2] * c.R[3 * i + 2] + c.t[i];
}

__global__ void k_project(const double *points, int n, Cam c, double *ij, double *depths) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double pc[3];
    to_camera(c, points + 3 * (size_t)i, pc);
    const double d = pc[2];
    double x = pc[0] / d, y = pc[1] / d;  // projected (:352)
    if (c.distorted) {                   // (:358-375)
        const double x2 = x * x, y2 = y * y, r2 = x2 + y2, r4 = r2 * r2, r6 = r2 * r4;
        const double radial = 1 + c.k1 * r2 + c.k2 * r4 + c.k3 * r6;
        const double tx = 2 * c.p1 * x * y + c.p2 * (r2 + 2 * x2);
        const double ty = c.p1 * (r2 + 2 * y2) + 2 * c.p2 * x * y;
        const double dx = x * radial + tx, dy = y * radial + ty;
        x = dx;
        y = dy;
    }
    // left_mul_intrinsic (:303-306): projected . K[:2,:2]^T + K[:2,2]
    ij[2 * (size_t)i] = x * c.K00 + y * c.K01 + c.K02;
    ij[2 * (size_t)i + 1] = x * c.K10 + y * c.K11 + c.K12;
    depths[i] = d;
}

__global__ void k_project_b(const double *points, int n, Cam c, const float *ij_b, const double *depths_b,
                            double *points_b, int reference_transpose) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double pc[3];
    to_camera(c, points + 3 * (size_t)i, pc);
    const double d = pc[2];
    const double x = pc[0] / d, y = pc[1] / d;
    const double gi = (double)ij_b[2 * (size_t)i], gj = (double)ij_b[2 * (size_t)i + 1];
    // projected_image_coordinates_b . K[:2,:2]^T  (:402 / :412, as the reference writes it)
    const double qx = gi * c.K00 + gj * c.K01, qy = gi * c.K10 + gj * c.K11;
    double x_b = qx, y_b = qy;
    if (c.distorted) {  // (:404-430)
        const double r2 = x * x + y * y;
        const double radial = 1 + c.k1 * r2 + c.k2 * r2 * r2 + c.k3 * r2 * r2 * r2;
        x_b = qx * radial;
        y_b = qy * radial;
        const double radial_b = qx * x + qy * y;
        x_b += qx * (2 * c.p1 * y + c.p2 * 4 * x);
        y_b += qx * 2 * c.p1 * x;
        x_b += qy * 2 * c.p2 * y;
        y_b += qy * (2 * c.p2 * x + c.p1 * 4 * y);
        double r2_b = qx * c.p2 + qy * c.p1;
        r2_b += radial_b * (c.k1 + 2 * c.k2 * r2 + 3 * c.k3 * r2 * r2);
        x_b += r2_b * 2 * x;
        y_b += r2_b * 2 * y;
    }
    // p_camera_b (:432-436)
    double pb[3] = {x_b / d, y_b / d, -(x_b * pc[0] + y_b * pc[1]) / (d * d)};
    if (depths_b) pb[2] += depths_b[i];
    for (int j = 0; j < 3; j++) {
        // reference (:438): p_camera_b . R^T ; adjoint of world_to_camera: p_camera_b . R
        const double v = reference_transpose ? pb[0] * c.R[3 * j] + pb[1] * c.R[3 * j + 1] + pb[2] * c.R[3 * j + 2]
                                             : pb[0] * c.R[j] + pb[1] * c.R[3 + j] + pb[2] * c.R[6 + j];
        points_b[3 * (size_t)i + j] += v;
    }
}

struct Light {
    double d[3], ambient;
    int directional;
};

__global__ void k_luminosity(const double *normals, const double *vcol, int n, int C, Light L, float *lum, float *colors) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    double dir = 0.0;
    if (L.directional) {
        const double *nv = normals + 3 * (size_t)v;
        const double s = (nv[0] * L.d[0] + nv[1] * L.d[1]) + nv[2] * L.d[2];  // np.sum(axis=1) of three terms
        dir = fmax(0.0, -s);
    }
    const double l = dir + L.ambient;
    if (lum) lum[v] = (float)l;
    if (colors && vcol)
        for (int c = 0; c < C; c++) colors[(size_t)v * C + c] = (float)(vcol[(size_t)v * C + c] * l);
}

// block-level sum of four doubles into out[0..3] (one fp64 atomic per block and component)
__device__ void block_add4(double v0, double v1, double v2, double v3, double *out) {
    __shared__ double part[4][BLOCK / 32];
    double v[4] = {v0, v1, v2, v3};
    for (int k = 0; k < 4; k++)
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
        for (int k = 0; k < 4; k++) part[k][warp] = v[k];
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0;
        for (int w = 0; w < BLOCK / 32; w++) s += part[threadIdx.x][w];
        if (s != 0.0) atomicAdd(out + threadIdx.x, s);
    }
}

__global__ void k_luminosity_b(const double *normals, const double *vcol, int n, int C, Light L, const float *colors_b,
                               const float *lum_b_in, double *normals_b, double *vcol_b, double *light_b) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    double lb = 0.0, g[3] = {0, 0, 0};
    if (v < n) {
        double dir = 0.0;
        const double *nv = normals + 3 * (size_t)v;
        if (L.directional) dir = fmax(0.0, -((nv[0] * L.d[0] + nv[1] * L.d[1]) + nv[2] * L.d[2]));
        const double l = dir + L.ambient;
        if (colors_b && vcol)
            for (int c = 0; c < C; c++) {
                const double cb = (double)colors_b[(size_t)v * C + c];
                lb += vcol[(size_t)v * C + c] * cb;                         // vertices_luminosity_b (:838)
                if (vcol_b) vcol_b[(size_t)v * C + c] = cb * l;             // vertices_colors_b (:839), assigned
            }
        if (lum_b_in) lb += (double)lum_b_in[v];
        const double gate = (L.directional && dir > 0) ? lb : 0.0;          // (:845-849)
        for (int j = 0; j < 3; j++) {
            g[j] = -gate * nv[j];
            if (normals_b) normals_b[3 * (size_t)v + j] = -gate * L.d[j];
        }
    }
    if (light_b) block_add4(g[0], g[1], g[2], lb, light_b);
}

// ---- f2 ------------------------------------------------------------------------------------------------------

// face_visible (:161-163): np.cross(u, v) of the 2-D edges, > 0 (clockwise) or < 0 - two products, one difference
__global__ void k_face_visible(DeodrMeshTopology m, const double *ij, uint8_t *visible) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= m.nb_faces) return;
    const uint32_t a = m.faces[3 * (size_t)f], b = m.faces[3 * (size_t)f + 1], c = m.faces[3 * (size_t)f + 2];
    const double ux = DSUB(ij[2 * (size_t)b], ij[2 * (size_t)a]), uy = DSUB(ij[2 * (size_t)b + 1], ij[2 * (size_t)a + 1]);
    const double vx = DSUB(ij[2 * (size_t)c], ij[2 * (size_t)a]), vy = DSUB(ij[2 * (size_t)c + 1], ij[2 * (size_t)a + 1]);
    const double cr = DSUB(DMUL(ux, vy), DMUL(uy, vx));
    visible[f] = (uint8_t)(m.clockwise ? cr > 0 : cr < 0);
}

// edge_bool = (edges_faces_ones * face_visible) == 1, gathered through faces_edges (:164-165)
__global__ void k_edge_flags(DeodrMeshTopology m, const uint8_t *visible, uint8_t *edgeflags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * m.nb_faces) return;
    const int e = m.faces_edges[i];
    int count = 0;
    for (int k = m.edge_face_offset[e]; k < m.edge_face_offset[e + 1]; k++) count += visible[m.edge_face_index[k]];
    edgeflags[i] = (uint8_t)(count == 1);
}

__device__ __forceinline__ void face_cross(const DeodrMeshTopology &m, const double *vertices, int f, double u[3],
                                           double v[3], double n[3]) {
    const uint32_t a = m.faces[3 * (size_t)f], b = m.faces[3 * (size_t)f + 1], c = m.faces[3 * (size_t)f + 2];
    for (int j = 0; j < 3; j++) {
        u[j] = vertices[3 * (size_t)b + j] - vertices[3 * (size_t)a + j];
        v[j] = vertices[3 * (size_t)c + j] - vertices[3 * (size_t)a + j];
    }
    n[0] = u[1] * v[2] - u[2] * v[1];
    n[1] = u[2] * v[0] - u[0] * v[2];
    n[2] = u[0] * v[1] - u[1] * v[0];
    if (m.clockwise)
        for (int j = 0; j < 3; j++) n[j] = -n[j];
}

// compute_face_normals (:113-123): normalize(+-cross(u, v))
__global__ void k_face_normals(DeodrMeshTopology m, const double *vertices, double *face_normals) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= m.nb_faces) return;
    double u[3], v[3], n[3];
    face_cross(m, vertices, f, u, v, n);
    const double norm = sqrt((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]);
    for (int j = 0; j < 3; j++) face_normals[3 * (size_t)f + j] = n[j] / norm;
}

// compute_vertex_normals (:137-143): normalize(_vertices_faces * face_normals) - a gather over the incident faces
__global__ void k_vertex_normals(DeodrMeshTopology m, const double *face_normals, double *vertex_normals) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= m.nb_vertices) return;
    double n[3] = {0, 0, 0};
    for (int k = m.vertex_face_offset[v]; k < m.vertex_face_offset[v + 1]; k++) {
        const int f = m.vertex_face_index[k];
        for (int j = 0; j < 3; j++) n[j] += face_normals[3 * (size_t)f + j];
    }
    const double norm = sqrt((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]);
    for (int j = 0; j < 3; j++) vertex_normals[3 * (size_t)v + j] = n[j] / norm;
}

// normalize_backward (deodr/tools.py:42-47)
__device__ __forceinline__ void normalize_b(const double x[3], const double xn_b[3], double out[3]) {
    const double n2 = (x[0] * x[0] + x[1] * x[1]) + x[2] * x[2];
    const double inv_n = 1.0 / sqrt(n2);
    const double n_b = -((xn_b[0] * x[0] + xn_b[1] * x[1]) + xn_b[2] * x[2]) * (inv_n * inv_n);
    for (int j = 0; j < 3; j++) out[j] = (xn_b[j] + x[j] * n_b) * inv_n;
}

// compute_vertex_normals_backward (:145-151): face_normals_b = _vertices_faces^T * normalize_backward(n, normals_b);
// first per vertex (the un-normalised sum is recomputed from the unit face normals), then gathered per face
__global__ void k_vertex_normals_b1(DeodrMeshTopology m, const double *vertices, const double *vertex_normals_b,
                                    double *vertex_sum_b) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= m.nb_vertices) return;
    double n[3] = {0, 0, 0};
    for (int k = m.vertex_face_offset[v]; k < m.vertex_face_offset[v + 1]; k++) {
        const int f = m.vertex_face_index[k];
        double u[3], w[3], c[3];
        face_cross(m, vertices, f, u, w, c);
        const double norm = sqrt((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]);
        for (int j = 0; j < 3; j++) n[j] += c[j] / norm;
    }
    double g[3] = {vertex_normals_b[3 * (size_t)v], vertex_normals_b[3 * (size_t)v + 1], vertex_normals_b[3 * (size_t)v + 2]};
    double out[3];
    normalize_b(n, g, out);
    for (int j = 0; j < 3; j++) vertex_sum_b[3 * (size_t)v + j] = out[j];
}

// compute_face_normals_backward (:125-135): per face, n_b = normalize_backward(n, normals_b); u_b, v_b =
// cross_backward(u, v, +-n_b); triangles_b = (-u_b - v_b, u_b, v_b) scattered with np.add.at
__global__ void k_face_normals_b(DeodrMeshTopology m, const double *vertices, const double *vertex_sum_b,
                                 double *face_normals_b, double *vertices_b) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= m.nb_faces) return;
    const uint32_t id[3] = {m.faces[3 * (size_t)f], m.faces[3 * (size_t)f + 1], m.faces[3 * (size_t)f + 2]};
    double fb[3] = {0, 0, 0};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) fb[j] += vertex_sum_b[3 * (size_t)id[i] + j];
    if (face_normals_b)
        for (int j = 0; j < 3; j++) face_normals_b[3 * (size_t)f + j] = fb[j];
    double u[3], v[3], n[3], n_b[3];
    face_cross(m, vertices, f, u, v, n);
    normalize_b(n, fb, n_b);
    if (m.clockwise)
        for (int j = 0; j < 3; j++) n_b[j] = -n_b[j];
    // cross_backward (tools.py:50-53): v_b = cross(c_b, u), u_b = cross(v, c_b)
    const double v_b[3] = {n_b[1] * u[2] - n_b[2] * u[1], n_b[2] * u[0] - n_b[0] * u[2], n_b[0] * u[1] - n_b[1] * u[0]};
    const double u_b[3] = {v[1] * n_b[2] - v[2] * n_b[1], v[2] * n_b[0] - v[0] * n_b[2], v[0] * n_b[1] - v[1] * n_b[0]};
    for (int j = 0; j < 3; j++) {
        atomicAdd(vertices_b + 3 * (size_t)id[0] + j, -u_b[j] - v_b[j]);
        atomicAdd(vertices_b + 3 * (size_t)id[1] + j, u_b[j]);
        atomicAdd(vertices_b + 3 * (size_t)id[2] + j, v_b[j]);
    }
}

int check_launch() {
    CUDA_TRY(cudaGetLastError());
    return DEODR_B200_OK;
}

int check_mesh(const DeodrMeshTopology *m) {
    if (!m) return set_error(DEODR_B200_EINVAL, "mesh == NULL");
    if (m->nb_faces < 0 || m->nb_vertices < 0 || m->nb_edges < 0) return set_error(DEODR_B200_EINVAL, "negative size");
    if (m->nb_faces > 0 && !m->faces) return set_error(DEODR_B200_EINVAL, "mesh.faces == NULL");
    return DEODR_B200_OK;
}

}  // namespace

extern "C" {

int deodr_b200_project_points(const double *points, int n, const DeodrCamera *camera, double *ij, double *depths,
                              void *stream) {
    if (!camera || n < 0 || (n > 0 && (!points || !ij || !depths))) return set_error(DEODR_B200_EINVAL, "bad argument");
    if (n == 0) return DEODR_B200_OK;
    k_project<<<grid_of(n), BLOCK, 0, (cudaStream_t)stream>>>(points, n, cam_of(camera), ij, depths);
    return check_launch();
}

int deodr_b200_project_points_b(const double *points, int n, const DeodrCamera *camera, const float *ij_b,
                                const double *depths_b, double *points_b, int reference_transpose, void *stream) {
    if (!camera || n < 0 || (n > 0 && (!points || !ij_b || !points_b))) return set_error(DEODR_B200_EINVAL, "bad argument");
    if (n == 0) return DEODR_B200_OK;
    k_project_b<<<grid_of(n), BLOCK, 0, (cudaStream_t)stream>>>(points, n, cam_of(camera), ij_b, depths_b, points_b,
                                                               reference_transpose);
    return check_launch();
}

static Light light_of(const double *directional, double ambient) {
    Light L;
    L.directional = directional != nullptr;
    for (int j = 0; j < 3; j++) L.d[j] = directional ? directional[j] : 0.0;
    L.ambient = ambient;
    return L;
}

int deodr_b200_vertex_luminosity(const double *normals, const double *vertex_colors, int n_vertices, int nb_colors,
                                 const double *light_directional, double ambient, float *luminosity, float *colors,
                                 void *stream) {
    if (n_vertices < 0 || nb_colors < 0 || (n_vertices > 0 && light_directional && !normals))
        return set_error(DEODR_B200_EINVAL, "bad argument");
    if (n_vertices == 0) return DEODR_B200_OK;
    k_luminosity<<<grid_of(n_vertices), BLOCK, 0, (cudaStream_t)stream>>>(normals, vertex_colors, n_vertices, nb_colors,
                                                                         light_of(light_directional, ambient),
                                                                         luminosity, colors);
    return check_launch();
}

int deodr_b200_vertex_luminosity_b(const double *normals, const double *vertex_colors, int n_vertices, int nb_colors,
                                   const double *light_directional, double ambient, const float *colors_b,
                                   const float *luminosity_b, double *normals_b, double *vertex_colors_b,
                                   double *light_b, void *stream) {
    if (n_vertices < 0 || nb_colors < 0 || (n_vertices > 0 && !normals))
        return set_error(DEODR_B200_EINVAL, "bad argument");
    if (n_vertices == 0) return DEODR_B200_OK;
    k_luminosity_b<<<grid_of(n_vertices), BLOCK, 0, (cudaStream_t)stream>>>(
        normals, vertex_colors, n_vertices, nb_colors, light_of(light_directional, ambient), colors_b, luminosity_b,
        normals_b, vertex_colors_b, light_b);
    return check_launch();
}

int deodr_b200_edge_on_silhouette(const DeodrMeshTopology *mesh, const double *ij, uint8_t *face_visible,
                                  uint8_t *edgeflags, void *stream) {
    if (int rc = check_mesh(mesh)) return rc;
    if (mesh->nb_faces == 0) return DEODR_B200_OK;
    if (!ij || !face_visible || !edgeflags || !mesh->faces_edges || !mesh->edge_face_offset || !mesh->edge_face_index)
        return set_error(DEODR_B200_EINVAL, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    k_face_visible<<<grid_of(mesh->nb_faces), BLOCK, 0, st>>>(*mesh, ij, face_visible);
    k_edge_flags<<<grid_of(3 * (size_t)mesh->nb_faces), BLOCK, 0, st>>>(*mesh, face_visible, edgeflags);
    return check_launch();
}

int deodr_b200_vertex_normals(const DeodrMeshTopology *mesh, const double *vertices, double *face_normals,
                              double *vertex_normals, void *stream) {
    if (int rc = check_mesh(mesh)) return rc;
    if (mesh->nb_faces == 0) return DEODR_B200_OK;
    if (!vertices || !face_normals) return set_error(DEODR_B200_EINVAL, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    k_face_normals<<<grid_of(mesh->nb_faces), BLOCK, 0, st>>>(*mesh, vertices, face_normals);
    if (vertex_normals) {
        if (!mesh->vertex_face_offset || !mesh->vertex_face_index) return set_error(DEODR_B200_EINVAL, "bad argument");
        k_vertex_normals<<<grid_of(mesh->nb_vertices), BLOCK, 0, st>>>(*mesh, face_normals, vertex_normals);
    }
    return check_launch();
}

int deodr_b200_vertex_normals_b(const DeodrMeshTopology *mesh, const double *vertices, const double *vertex_normals_b,
                                double *vertex_scratch, double *vertices_b, void *stream) {
    if (int rc = check_mesh(mesh)) return rc;
    if (mesh->nb_faces == 0) return DEODR_B200_OK;
    if (!vertices || !vertex_normals_b || !vertex_scratch || !vertices_b || !mesh->vertex_face_offset ||
        !mesh->vertex_face_index)
        return set_error(DEODR_B200_EINVAL, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    k_vertex_normals_b1<<<grid_of(mesh->nb_vertices), BLOCK, 0, st>>>(*mesh, vertices, vertex_normals_b, vertex_scratch);
    k_face_normals_b<<<grid_of(mesh->nb_faces), BLOCK, 0, st>>>(*mesh, vertices, vertex_scratch, nullptr, vertices_b);
    return check_launch();
}

}  // extern "C"
