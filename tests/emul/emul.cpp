// TEST INFRASTRUCTURE ONLY - CPU emulation of the CUDA kernels of deodr_b200/csrc/kernels.cu.
//
// The kernels are written as barrier-separated phases (deodr_b200/csrc/phases.h) that are __host__ __device__.
// This harness compiles the SAME phase functions with g++ and runs each CTA as `for (tid ...)` loops per phase, with
// the same control flow as the __global__ wrappers, so that the kernel logic and numerics can be checked against the
// oracle in a container that has no GPU.  It is never linked into, loaded by or shipped with the product library,
// and it is not a fallback: deodr_b200 fails loudly without CUDA.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../../deodr_b200/csrc/phases.h"
#include "../../include/deodr_b200.h"

using namespace deodr;

struct HostEnv {
    static int atomic_add(int *p, int v) { int old = *p; *p += v; return old; }
    static void atomic_add(float *p, float v) { *p += v; }
    static void atomic_add(double *p, double v) { *p += v; }
    static void atomic_or(int *p, int v) { *p |= v; }
    static int shared_inc(int *p) { return (*p)++; }
    void emit(float *p, float v) const { *p += v; }
};

struct EmulState {
    int tiles_x = 0, tiles_y = 0, nt = 0, E = 0;
    // the plan: segment offsets (counts of a count-only pass, padded with the device's slack) + edge capacity
    std::vector<int> small_offset, large_offset, edge_offset;
    int cap_edges = 0;
    // one forward pass
    int scal[SC_WORDS] = {0};
    std::vector<int> small_cursor, large_cursor, edge_cursor, large_refs;
    std::vector<PreRec> small_recs;
    std::vector<int> small_ids;
    std::vector<int> edge_ids;
    std::vector<uint64_t> edge_keys;
    std::vector<int> edge_refs_tmp, edge_refs;
    std::vector<EdgeRec> edge_recs;
    std::vector<int> tie_pairs;
    TileSegments small_seg() { return TileSegments{small_offset.data(), small_cursor.data()}; }
    TileSegments large_seg() { return TileSegments{large_offset.data(), large_cursor.data()}; }
    TileSegments edge_seg() { return TileSegments{edge_offset.data(), edge_cursor.data()}; }
};

static EmulState g_state;
static int g_small_records = 0;   // the small triangles' adjoint: 0 = triangle-parallel (device default), 1 = record-parallel (k_small_rec_bwd)
static int g_small_textured = 1;  // 0: textured triangles go through the pixel-parallel adjoint (TriBins::small_textured)
static int g_record_rows = RECORD_ROWS;  // the device picks 0 for scenes of a few thousand triangles: both are emulated

// k_scan_tiles: exclusive scan of the counts padded with slack (c + c/4 + 4, kernels.cu SEG_SLACK)
static void scan_tiles(const std::vector<int> &count, std::vector<int> &offset) {
    int run = 0;
    offset.resize(count.size() + 1);
    for (size_t i = 0; i < count.size(); i++) {
        const int raw = count[i] & 0x3fffffff;  // bit 30: launch-hint flag of the count pass (see k_scan_tiles)
        offset[i] = run;
        run += raw + (raw >> 2) + 4;
    }
    offset[count.size()] = run;
}

static void decode_owner(int code, const EmulState &st, int *own, int *bown) {
    if (code <= -2) { *own = st.tie_pairs[2 * (-2 - code)]; *bown = st.tie_pairs[2 * (-2 - code) + 1]; }
    else *own = *bown = code;
}

// k_tile_z + k_shade<MAXC> + k_edge_fwd<MAXC>
template <int MAXC>
static void raster_fwd(const SceneView &s, double sigma, EmulState &st, float *image, double *z_buffer, int *owner,
                       int *face_id, const float *obs = nullptr, float *err = nullptr) {
    TileShared *sh = new TileShared;
    memset(sh, 0, sizeof(TileShared));
    // ---- k_tile_z
    {
        std::vector<PixelState<1>> px(NT);
        for (int tile_id = 0; tile_id < st.nt; tile_id++) {
            const Tile tile = tile_of(tile_id, st.tiles_x);
            for (int tid = 0; tid < NT; tid++) { px[tid].z = std::numeric_limits<double>::infinity(); px[tid].own = px[tid].bown = -1; }
            auto inside = [&](int tid) { return tile.x0 + tid % TS < s.width && tile.y0 + tid / TS < s.height; };
            const int n_small = segment_size(st.small_seg(), tile_id);
            for (int base = 0; base < n_small; base += PRE_CHUNK) {
                const int m = std::min(PRE_CHUNK, n_small - base);
                // the device pulls the chunk into shared memory with one bulk (TMA) copy; same bytes here
                PreRec pre[PRE_CHUNK];
                memcpy(pre, st.small_recs.data() + st.small_offset[tile_id] + base, m * sizeof(PreRec));
                for (int tid = 0; tid < NT; tid++) phase_pre_scatter<HostEnv>(tid, m, pre, sh);
                for (int tid = 0; tid < NT; tid++) phase_pix_test<1>(s, tid, m, tile, pre, sh, &px[tid]);
            }
            const int n_large = segment_size(st.large_seg(), tile_id);
            for (int base = 0; base < n_large; base += LARGE_CHUNK) {
                const int m = std::min(LARGE_CHUNK, n_large - base);
                for (int tid = 0; tid < NT; tid++)
                    phase_tri_setup(s, tid, m, st.large_refs.data() + st.large_offset[tile_id] + base, sh);
                for (int tid = 0; tid < NT; tid++) phase_tri_masks(s, tid, m, tile, sh);
                for (int tid = 0; tid < NT; tid++) if (inside(tid)) phase_tri_test<1>(s, tid, m, tile, sh, &px[tid]);
            }
            for (int tid = 0; tid < NT; tid++) {
                if (!inside(tid)) continue;
                const size_t idx = (size_t)(tile.y0 + tid / TS) * s.width + tile.x0 + tid % TS;
                const PixelState<1> &p = px[tid];
                z_buffer[idx] = p.z;
                int code = p.bown;
                if (p.own != p.bown) {
                    int slot = (int)st.tie_pairs.size() / 2;
                    st.tie_pairs.push_back(p.own);
                    st.tie_pairs.push_back(p.bown);
                    code = -2 - slot;
                }
                owner[idx] = code;
                if (face_id) face_id[idx] = p.own >= 0 ? (p.own & TRI_INDEX_MASK) : -1;
            }
        }
    }
    // ---- k_shade
    for (int y = 0; y < s.height; y++)
        for (int x = 0; x < s.width; x++) {
            const size_t idx = (size_t)y * s.width + x;
            PixelState<MAXC> p;
            decode_owner(owner[idx], st, &p.own, &p.bown);
            p.z = s.perspective_correct && p.own >= 0 ? z_buffer[idx] : 0.0;
            phase_shade<MAXC>(s, x, y, &p);
            for (int k = 0; k < s.nb_colors; k++) image[idx * s.nb_colors + k] = p.col[k];
        }
    // ---- antialiase_error mode: residual of every pixel, which the edges then overdraw instead of the colours
    if (err)
        for (size_t idx = 0; idx < (size_t)s.height * s.width; idx++)
            err[idx] = (float)pixel_residual<MAXC>(s, image + idx * s.nb_colors, obs + idx * s.nb_colors);
    // ---- k_edge_fwd
    std::vector<PixelState<MAXC>> px(NT);
    for (int tile_id = 0; tile_id < st.nt && st.E > 0; tile_id++) {
        const int n_edge = segment_size(st.edge_seg(), tile_id);
        if (n_edge == 0) continue;
        const Tile tile = tile_of(tile_id, st.tiles_x);
        auto inside = [&](int tid) { return tile.x0 + tid % TS < s.width && tile.y0 + tid / TS < s.height; };
        for (int tid = 0; tid < NT; tid++) {
            if (!inside(tid)) continue;
            const size_t idx = (size_t)(tile.y0 + tid / TS) * s.width + tile.x0 + tid % TS;
            px[tid].z = z_buffer[idx];
            for (int k = 0; k < s.nb_colors; k++) px[tid].col[k] = image[idx * s.nb_colors + k];
        }
        const int edge_base = st.edge_offset[tile_id];
        for (int base = 0; base < n_edge; base += EDGE_CHUNK) {
            const int m = std::min(EDGE_CHUNK, n_edge - base);
            for (int tid = 0; tid < NT; tid++)
                phase_edge_setup(tid, NT, m, st.edge_refs.data() + edge_base + base, st.edge_recs.data(), sh);
            for (int tid = 0; tid < NT; tid++) phase_edge_spans(s, tid, NT, m, tile, 0, TS, sh);
            for (int tid = 0; tid < NT; tid++) {
                if (!inside(tid)) continue;
                const int x = tile.x0 + tid % TS, y = tile.y0 + tid / TS;
                const size_t idx = (size_t)y * s.width + x;
                if (err) phase_edge_blend_error<MAXC>(s, x, y, tid / TS, m, sh, px[tid].z, obs + idx * s.nb_colors, err + idx);
                else phase_edge_blend<MAXC>(s, x, y, tid / TS, m, sh, &px[tid]);
            }
        }
        for (int tid = 0; tid < NT && !err; tid++) {
            if (!inside(tid)) continue;
            const size_t idx = (size_t)(tile.y0 + tid / TS) * s.width + tile.x0 + tid % TS;
            for (int k = 0; k < s.nb_colors; k++) image[idx * s.nb_colors + k] = px[tid].col[k];
        }
    }
    delete sh;
}

template <int MAXC>
static void raster_bwd(const SceneView &s, double sigma, EmulState &st, const double *z_buffer, const int *owner,
                       const float *image_b, const DeodrGrads &g, double *edge_acc, const float *obs = nullptr,
                       const float *err_b = nullptr, bool compat = true) {
    std::vector<PixelState<MAXC>> px(NT);
    std::vector<AdjointState<MAXC>> adj(NT);
    std::vector<ErrorAdjointState<MAXC>> eadj(NT);  // antialiase_error mode (obs != nullptr)
    TileShared *sh = new TileShared;
    for (int tile_id = 0; tile_id < st.nt; tile_id++) {
        const Tile tile = tile_of(tile_id, st.tiles_x);
        auto inside = [&](int tid) { return tile.x0 + tid % TS < s.width && tile.y0 + tid / TS < s.height; };
        for (int tid = 0; tid < NT; tid++) {
            PixelState<MAXC> &p = px[tid];
            AdjointState<MAXC> &a = adj[tid];
            a.has_colour = false;
            p.z = std::numeric_limits<double>::infinity();
            p.own = p.bown = -1;
            if (!inside(tid)) continue;
            const size_t idx = (size_t)(tile.y0 + tid / TS) * s.width + tile.x0 + tid % TS;
            p.z = z_buffer[idx];
            int code = owner[idx];
            if (code <= -2) { p.own = st.tie_pairs[2 * (-2 - code)]; p.bown = st.tie_pairs[2 * (-2 - code) + 1]; }
            else p.own = p.bown = code;
            for (int k = 0; k < s.nb_colors; k++) a.g[k] = image_b[idx * s.nb_colors + k];
            eadj[tid].has = false;
            eadj[tid].g = err_b ? (double)err_b[idx] : 0.0;
        }
        const int n_edge = st.E > 0 ? segment_size(st.edge_seg(), tile_id) : 0;
        if (n_edge > 0) {
            const int edge_base = st.edge_offset[tile_id];
            const bool single = n_edge <= EDGE_CHUNK;
            for (int base = 0; base < n_edge; base += EDGE_CHUNK) {
                const int m = std::min(EDGE_CHUNK, n_edge - base);
                for (int tid = 0; tid < NT; tid++)
                    phase_edge_setup(tid, NT, m, st.edge_refs.data() + edge_base + base, st.edge_recs.data(), sh);
                for (int tid = 0; tid < NT; tid++) phase_edge_spans(s, tid, NT, m, tile, 0, TS, sh);
                for (int tid = 0; tid < NT; tid++) {
                    if (!inside(tid)) continue;
                    const int x = tile.x0 + tid % TS, y = tile.y0 + tid / TS;
                    if (obs) {
                        const float *ob = obs + ((size_t)y * s.width + x) * s.nb_colors;
                        phase_edge_replay_error<MAXC>(s, x, y, tid / TS, m, sh, px[tid], ob, &eadj[tid]);
                        if (single && eadj[tid].has)
                            phase_edge_adjoint_error<MAXC, HostEnv>(s, x, y, tid / TS, m, sh, px[tid], ob, &eadj[tid],
                                                                    edge_acc, g.texture_b, compat);
                        continue;
                    }
                    phase_edge_replay<MAXC>(s, x, y, tid / TS, m, sh, px[tid], &adj[tid]);
                    if (single && adj[tid].has_colour)
                        phase_edge_adjoint<MAXC, HostEnv>(s, x, y, tid / TS, m, sh, px[tid], &adj[tid], edge_acc, g.texture_b);
                }
            }
            if (!single) {
                const int last = ((n_edge - 1) / EDGE_CHUNK) * EDGE_CHUNK;
                for (int base = last; base >= 0; base -= EDGE_CHUNK) {
                    const int m = std::min(EDGE_CHUNK, n_edge - base);
                    for (int tid = 0; tid < NT; tid++)
                        phase_edge_setup(tid, NT, m, st.edge_refs.data() + edge_base + base, st.edge_recs.data(), sh);
                    for (int tid = 0; tid < NT; tid++) phase_edge_spans(s, tid, NT, m, tile, 0, TS, sh);
                    for (int tid = 0; tid < NT; tid++) {
                        if (!inside(tid)) continue;
                        const int x = tile.x0 + tid % TS, y = tile.y0 + tid / TS;
                        if (obs) {
                            if (eadj[tid].has)
                                phase_edge_adjoint_error<MAXC, HostEnv>(s, x, y, tid / TS, m, sh, px[tid],
                                                                        obs + ((size_t)y * s.width + x) * s.nb_colors,
                                                                        &eadj[tid], edge_acc, g.texture_b, compat);
                            continue;
                        }
                        if (!adj[tid].has_colour) continue;
                        phase_edge_adjoint<MAXC, HostEnv>(s, x, y, tid / TS, m, sh, px[tid], &adj[tid], edge_acc, g.texture_b);
                    }
                }
            }
        }
        // k_raster_bwd takes every pixel of the tiles with silhouette edges; elsewhere k_interior_bwd takes the pixels
        // owned by large triangles and k_small_tri_bwd (below, triangle-parallel) those owned by small ones
        for (int tid = 0; tid < NT; tid++) {
            if (!inside(tid) || px[tid].bown < 0) continue;
            if (n_edge == 0 && (px[tid].bown & SMALL_FLAG)) continue;
            if (obs && eadj[tid].has) {  // pixel overdrawn by edges: colour adjoint from what is left of the residual's
                const size_t idx = (size_t)(tile.y0 + tid / TS) * s.width + tile.x0 + tid % TS;
                residual_adjoint<MAXC>(s, eadj[tid].col, obs + idx * s.nb_colors, eadj[tid].g, adj[tid].g);
            }
            phase_interior_adjoint<MAXC, HostEnv>(s, tile.x0 + tid % TS, tile.y0 + tid / TS, px[tid], adj[tid].g,
                                                  g.ij_b, g.colors_b, g.uv_b, g.shade_b, g.texture_b, HostEnv());
        }
    }
    delete sh;
    const int *edge_count = st.E > 0 ? st.edge_cursor.data() : nullptr;
    if (g_small_records) {  // k_small_rec_bwd: one thread per record of the tiles without silhouette edges
        for (int tile_id = 0; tile_id < st.nt; tile_id++) {
            if (edge_count && edge_count[tile_id] > 0) continue;
            const Tile tile = tile_of(tile_id, st.tiles_x);
            const GlobalTileFetch fetch{owner, image_b, tile, s.width, s.height, s.nb_colors};
            const int n = segment_size(st.small_seg(), tile_id);
            for (int i = 0; i < n; i++)
                small_record_adjoint<MAXC, HostEnv>(s, st.small_recs[st.small_offset[tile_id] + i], tile, fetch,
                                                    st.tie_pairs.data(), g.ij_b, g.colors_b, g.uv_b, g.shade_b, g.texture_b);
        }
        return;
    }
    // k_small_tri_bwd
    for (int k : st.small_ids)
        small_triangle_adjoint<MAXC, HostEnv>(s, k, st.tiles_x, edge_count, owner, st.tie_pairs.data(), image_b, g.ij_b,
                                              g.colors_b, g.uv_b, g.shade_b, g.texture_b);
}

// Plan building (build_plan of kernels.cu): count-only pass + scans with slack.
static void emul_plan(const SceneView &s, double sigma, EmulState &st) {
    const int T = s.nb_triangles;
    st.tiles_x = (s.width + TS - 1) / TS;
    st.tiles_y = (s.height + TS - 1) / TS;
    st.nt = st.tiles_x * st.tiles_y;
    std::vector<int> small_count(st.nt, 0), large_count(st.nt, 0), edge_count(st.nt, 0);
    int scal[SC_WORDS] = {0};
    TriBins bins{{nullptr, small_count.data()}, nullptr, {nullptr, large_count.data()}, nullptr, scal + SC_OVERFLOW,
                 g_record_rows, g_small_textured};
    EdgeList edges{scal + SC_EDGES, nullptr, nullptr, 0};
    for (int k = 0; k < T; k++)
        bin_triangle<HostEnv, true>(s, k, sigma, st.tiles_x, bins, scal + SC_SMALL, nullptr, edges, edge_count.data());
    scan_tiles(small_count, st.small_offset);
    scan_tiles(large_count, st.large_offset);
    scan_tiles(edge_count, st.edge_offset);
    const int E = scal[SC_EDGES];
    st.cap_edges = E > 0 ? E + E / 4 + 64 : 0;
}

// One forward pass against the current plan (enqueue_forward of kernels.cu); returns the verdict word.
static int emul_forward(const SceneView &s, double sigma, EmulState &st, float *image, double *z_buffer, int32_t *owner,
                        int32_t *face_id, const float *obs, float *err) {
    const int T = s.nb_triangles;
    memset(st.scal, 0, sizeof(st.scal));
    st.small_cursor.assign(st.nt, 0);
    st.large_cursor.assign(st.nt, 0);
    st.edge_cursor.assign(st.nt, 0);
    st.small_recs.assign(st.small_offset[st.nt] + 1, PreRec());
    st.large_refs.assign(st.large_offset[st.nt] + 4, -1);
    st.small_ids.assign(T + 4, -1);
    st.edge_ids.assign(st.cap_edges + 4, -1);
    st.edge_keys.assign(st.cap_edges + 4, 0);
    st.tie_pairs.clear();
    TriBins bins{st.small_seg(), st.small_recs.data(), st.large_seg(), st.large_refs.data(), st.scal + SC_OVERFLOW,
                 g_record_rows, g_small_textured};
    EdgeList edges{st.scal + SC_EDGES, st.edge_ids.data(), st.edge_keys.data(), st.cap_edges};
    // k_bin, in DESCENDING triangle order: the device appends in an arbitrary order, nothing may depend on it
    for (int k = T - 1; k >= 0; k--)
        bin_triangle<HostEnv, false>(s, k, sigma, st.tiles_x, bins, st.scal + SC_SMALL, st.small_ids.data(), edges, nullptr);
    st.small_ids.resize(st.scal[SC_SMALL]);
    st.E = std::min(st.scal[SC_EDGES], st.cap_edges);
    if (st.cap_edges == 0) st.E = 0;
    // k_bin_edges (again in reversed order) + k_sort_tile_edges
    st.edge_recs.assign(st.cap_edges + 1, EdgeRec());
    st.edge_refs_tmp.assign(st.edge_offset[st.nt] + 4, -1);
    st.edge_refs.assign(st.edge_offset[st.nt] + 4, -1);
    if (!st.scal[SC_OVERFLOW]) {
        std::vector<int> tiles_raw(st.nt + 1, -1);
        EdgeBins ebins{st.edge_seg(), st.edge_refs_tmp.data(), st.scal + SC_OVERFLOW, tiles_raw.data(),
                       st.scal + SC_EDGE_TILES, st.nt};
        for (int i = st.E - 1; i >= 0; i--) bin_edge<HostEnv>(s, i, sigma, st.tiles_x, edges, ebins, st.edge_recs.data());
    }
    if (!st.scal[SC_OVERFLOW])
        for (int t = 0; t < st.nt; t++) {
            const int n = segment_size(st.edge_seg(), t), base = st.edge_offset[t];
            for (int i = 0; i < n; i++)
                st.edge_refs[base + tile_edge_position(i, n, st.edge_refs_tmp.data() + base, st.edge_recs.data())] =
                    st.edge_refs_tmp[base + i];
        }
    if (st.scal[SC_OVERFLOW]) return st.scal[SC_OVERFLOW];  // the device kernels return at once: the pass is void
    const int C = s.nb_colors;
    if (C == 1) raster_fwd<1>(s, sigma, st, image, z_buffer, owner, face_id, obs, err);
    else if (C <= 3) raster_fwd<3>(s, sigma, st, image, z_buffer, owner, face_id, obs, err);
    else if (C <= 4) raster_fwd<4>(s, sigma, st, image, z_buffer, owner, face_id, obs, err);
    else raster_fwd<16>(s, sigma, st, image, z_buffer, owner, face_id, obs, err);
    return 0;
}

extern "C" {

// Same contract as deodr_b200_render, with HOST pointers in the canonical layout: plan + pass.
int emul_render(const DeodrSceneView *scene, double sigma, float *image, double *z_buffer, int32_t *owner,
                int32_t *face_id) {
    SceneView s;
    memcpy(&s, scene, sizeof(s));
    EmulState &st = g_state;
    st = EmulState();
    emul_plan(s, sigma, st);
    return emul_forward(s, sigma, st, image, z_buffer, owner, face_id, nullptr, nullptr) ? -1 : 0;
}

// The plan of one scene ...
int emul_build_plan(const DeodrSceneView *scene, double sigma) {
    SceneView s;
    memcpy(&s, scene, sizeof(s));
    g_state = EmulState();
    emul_plan(s, sigma, g_state);
    return 0;
}

// ... used by the pass over ANOTHER scene of the same shape (an optimisation loop moves the vertices between two
// passes): returns the verdict word - 0 = the lists fitted, the outputs are valid; otherwise the pass is void.
int emul_render_planned(const DeodrSceneView *scene, double sigma, float *image, double *z_buffer, int32_t *owner,
                        int32_t *face_id) {
    SceneView s;
    memcpy(&s, scene, sizeof(s));
    return emul_forward(s, sigma, g_state, image, z_buffer, owner, face_id, nullptr, nullptr);
}

// antialiase_error mode: the residual and its overdraw by the edges (the image keeps its aliased edges).
int emul_render_error(const DeodrSceneView *scene, double sigma, const float *obs, float *image, double *z_buffer,
                      int32_t *owner, int32_t *face_id, float *err) {
    SceneView s;
    memcpy(&s, scene, sizeof(s));
    EmulState &st = g_state;
    st = EmulState();
    emul_plan(s, sigma, st);
    return emul_forward(s, sigma, st, image, z_buffer, owner, face_id, obs, err) ? -1 : 0;
}

// adjoint of emul_render_error: `image` is its (aliased) output, err_b the adjoint of the residual buffer
int emul_render_error_b(const DeodrSceneView *scene, double sigma, const double *z_buffer, const int32_t *owner,
                        const float *image, const float *obs, const float *err_b, const DeodrGrads *grads, int compat) {
    SceneView s;
    memcpy(&s, scene, sizeof(s));
    EmulState &st = g_state;
    const int C = s.nb_colors;
    const size_t P = (size_t)s.height * s.width;
    // colour adjoint of the pixels no edge touches (one elementwise kernel on the device): DR.h:3054-3060
    std::vector<float> image_b(P * C);
    for (size_t i = 0; i < P; i++)
        for (int c = 0; c < C; c++)
            image_b[i * C + c] = (float)(-2.0 * ((double)obs[i * C + c] - (double)image[i * C + c]) * (double)err_b[i]);
    std::vector<double> acc((size_t)std::max(st.E, 1) * edge_acc_stride(C), 0.0);
    if (C == 1) raster_bwd<1>(s, sigma, st, z_buffer, owner, image_b.data(), *grads, acc.data(), obs, err_b, compat != 0);
    else if (C <= 3) raster_bwd<3>(s, sigma, st, z_buffer, owner, image_b.data(), *grads, acc.data(), obs, err_b, compat != 0);
    else if (C <= 4) raster_bwd<4>(s, sigma, st, z_buffer, owner, image_b.data(), *grads, acc.data(), obs, err_b, compat != 0);
    else raster_bwd<16>(s, sigma, st, z_buffer, owner, image_b.data(), *grads, acc.data(), obs, err_b, compat != 0);
    for (int r = 0; r < st.E; r++)
        finalize_edge<HostEnv>(s, st.edge_ids[r], sigma, acc.data() + (size_t)r * edge_acc_stride(C), grads->ij_b,
                               grads->colors_b, grads->uv_b, grads->shade_b);
    return 0;
}

// Same contract as deodr_b200_render_b (must follow emul_render on the same scene).
int emul_render_b(const DeodrSceneView *scene, double sigma, const double *z_buffer, const int32_t *owner,
                  const float *image_b, const DeodrGrads *grads) {
    SceneView s;
    memcpy(&s, scene, sizeof(s));
    EmulState &st = g_state;
    const int C = s.nb_colors;
    std::vector<double> acc((size_t)std::max(st.E, 1) * edge_acc_stride(C), 0.0);
    if (C == 1) raster_bwd<1>(s, sigma, st, z_buffer, owner, image_b, *grads, acc.data());
    else if (C <= 3) raster_bwd<3>(s, sigma, st, z_buffer, owner, image_b, *grads, acc.data());
    else if (C <= 4) raster_bwd<4>(s, sigma, st, z_buffer, owner, image_b, *grads, acc.data());
    else raster_bwd<16>(s, sigma, st, z_buffer, owner, image_b, *grads, acc.data());
    for (int r = 0; r < st.E; r++)
        finalize_edge<HostEnv>(s, st.edge_ids[r], sigma, acc.data() + (size_t)r * edge_acc_stride(C), grads->ij_b,
                               grads->colors_b, grads->uv_b, grads->shade_b);
    return 0;
}

// test hooks for the exact-arithmetic helpers of rmath.h
int emul_floor_div(double a, double b, int lo, int hi, int fast) {
    return fast ? floor_div_clamped(a, b, lo, hi) : floor_div_clamped_reference(a, b, lo, hi);
}
int emul_ceil_div(double a, double b, int lo, int hi, int fast) {
    return fast ? ceil_div_clamped(a, b, lo, hi) : ceil_div_clamped_reference(a, b, lo, hi);
}
// batch comparison: returns the number of mismatches between the fast and the reference formulations
long emul_div_mismatches(const double *a, const double *b, long n, int lo, int hi) {
    long bad = 0;
    for (long i = 0; i < n; i++) {
        bad += floor_div_clamped(a[i], b[i], lo, hi) != floor_div_clamped_reference(a[i], b[i], lo, hi);
        bad += ceil_div_clamped(a[i], b[i], lo, hi) != ceil_div_clamped_reference(a[i], b[i], lo, hi);
        int x;  // the branch-free variant used by edge_row_span makes the same decisions
        if (floor_quotient_try(a[i], b[i], &x)) x = floor_quotient_exact(a[i], b[i]);
        bad += x != floor_quotient(a[i], b[i]);
    }
    return bad;
}

// edge_row_span (four quotients side by side) against the line-by-line formulation; ineq = n x 12 doubles
long emul_span_mismatches(const double *ineq, const int *y, long n, int width) {
    long bad = 0;
    for (long i = 0; i < n; i++) {
        EdgeGeom g;
        memcpy(g.ineq, ineq + 12 * i, sizeof(g.ineq));
        int b0, e0, b1, e1;
        edge_row_span(g, width, y[i], &b0, &e0);
        edge_row_span_reference(g, width, y[i], &b1, &e1);
        bad += (b0 != b1) || (e0 != e1);
    }
    return bad;
}

// tri_half_span (two quotients side by side) against the line-by-line formulation, every row of both halves;
// V = n x 6 doubles (x0 y0 x1 y1 x2 y2)
long emul_tri_span_mismatches(const double *V, long n, int width, int height, int strict) {
    long bad = 0;
    for (long i = 0; i < n; i++) {
        double P[3][2] = {{V[6 * i], V[6 * i + 1]}, {V[6 * i + 2], V[6 * i + 3]}, {V[6 * i + 4], V[6 * i + 5]}};
        double Z[3] = {1.0, 2.0, 3.0};
        TriGeom g;
        tri_geom(P, Z, strict != 0, false, &g, nullptr);
        for (int half = 0; half < 2; half++) {
            int y0 = g.y_begin[half] < 0 ? 0 : g.y_begin[half], y1 = g.y_end[half] > height - 1 ? height - 1 : g.y_end[half];
            for (int y = y0; y <= y1; y++) {
                int b0, e0, b1, e1;
                tri_half_span(g, half, y, width, strict != 0, &b0, &e0);
                tri_half_span_reference(g, half, y, width, strict != 0, &b1, &e1);
                bad += (b0 != b1) || (e0 != e1);
            }
        }
    }
    return bad;
}

// Size of the interpolation weights of the forward owner at every pixel (1 where there is none): the scale of the fp32
// rounding of a colour that is a sum w0 a0 + w1 a1 + w2 a2 (sliver or edge-on triangles extrapolate with |w| >> 1).
void emul_weight_scale(const DeodrSceneView *scene, const int32_t *face_id, const double *z_buffer, float *out) {
    SceneView s;
    memcpy(&s, scene, sizeof(s));
    for (int y = 0; y < s.height; y++)
        for (int x = 0; x < s.width; x++) {
            const size_t idx = (size_t)y * s.width + x;
            float m = 1.0f;
            if (face_id[idx] >= 0) {
                TriAttr t;
                tri_attr(s, face_id[idx], &t);
                double w[3];
                tri_weights(s, t, x, y, z_buffer[idx], w);
                for (int i = 0; i < 3; i++) m = std::max(m, (float)std::fabs(w[i]));
            }
            out[idx] = m;
        }
}

void emul_set_record_rows(int rows) { g_record_rows = rows; }
void emul_set_small_textured(int on) { g_small_textured = on; }
void emul_set_small_records(int on) { g_small_records = on; }
int emul_num_ties(void) { return (int)g_state.tie_pairs.size() / 2; }
int emul_num_edges(void) { return g_state.E; }
int emul_tri_refs(void) {
    int n = 0;
    for (int v : g_state.small_cursor) n += v;
    for (int v : g_state.large_cursor) n += v;
    return n;
}
}
