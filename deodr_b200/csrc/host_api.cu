// Reference-shaped host entry points of the C-ABI: deodr_b200_render_host / deodr_b200_render_b_host take the
// reference's `struct Scene` (DR.h:56-90, HOST pointers, fp64) like renderScene (DR.h:2717) / renderScene_B (DR.h:2903).
//
// Data path of one call (everything below is inside the timed region of bench.py's `e2e`):
//   user arrays --(copy threads: convert fp64->fp32 where the device layout is fp32)--> pinned mirror --(DMA)--> HBM
//   kernels (kernels.cu)
//   HBM --(DMA, chunked)--> pinned staging --(copy threads: fp32->fp64)--> user image / z_buffer / gradients
// The pinned mirror of the scene outlives the forward call: renderScene_B re-derives everything from the scene in the
// reference; here the adjoint call compares the caller's arrays with the mirror (exact, multi-threaded memcmp) and,
// when they are identical to the last forward's, reuses the device-resident scene, z-buffer, owner ids and tile edge
// lists instead of re-uploading and re-rendering.  Any difference -> full re-stage + forward, so the call stays
// stateless in its semantics.
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <new>
#include <thread>

#include <chrono>
#include <cstdlib>

#include "workspace.h"

// DEODR_B200_TRACE=1 prints the wall-clock breakdown of the host entry points (development aid)
struct HostTrace {
    bool on;
    std::chrono::steady_clock::time_point t0;
    const char *what;
    explicit HostTrace(const char *w) : on(getenv("DEODR_B200_TRACE") != nullptr), what(w) {
        if (on) t0 = std::chrono::steady_clock::now();
    }
    void lap(const char *label) {
        if (!on) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[deodr_b200 %s] %-28s %8.3f ms\n", what, label,
                std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// ---------------------------------------------------------------------------------------------- copy thread pool

class CopyPool {
   public:
    explicit CopyPool(int n) : stop_(false), generation_(0), pending_(0) {
        for (int i = 0; i < n; i++) threads_.emplace_back([this, i] { loop(i); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lock(mu_);
            stop_ = true;
            generation_++;
        }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    int size() const { return (int)threads_.size() + 1; }
    // runs fn(part, nparts) on every worker and on the caller; returns when all parts are done
    void run(const std::function<void(int, int)> &fn) {
        const int nparts = size();
        {
            std::lock_guard<std::mutex> lock(mu_);
            fn_ = &fn;
            pending_ = (int)threads_.size();
            generation_++;
        }
        cv_.notify_all();
        fn(nparts - 1, nparts);
        std::unique_lock<std::mutex> lock(mu_);
        done_.wait(lock, [this] { return pending_ == 0; });
    }

   private:
    void loop(int index) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int, int)> *fn;
            {
                std::unique_lock<std::mutex> lock(mu_);
                cv_.wait(lock, [&] { return generation_ != seen; });
                seen = generation_;
                if (stop_) return;
                fn = fn_;
            }
            (*fn)(index, size());
            {
                std::lock_guard<std::mutex> lock(mu_);
                pending_--;
            }
            done_.notify_one();
        }
    }
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    const std::function<void(int, int)> *fn_ = nullptr;
    bool stop_;
    uint64_t generation_;
    int pending_;
};

static inline void part_range(size_t n, int part, int nparts, size_t *lo, size_t *hi) {
    size_t per = (n + nparts - 1) / nparts;
    per = (per + 63) & ~(size_t)63;  // whole cache lines per part
    *lo = std::min(n, per * part);
    *hi = std::min(n, *lo + per);
}

struct PinnedBuf {
    void *ptr = nullptr;
    size_t bytes = 0;
    int ensure(size_t need) {
        if (need <= bytes) return DEODR_B200_OK;
        if (ptr) CUDA_TRY(cudaFreeHost(ptr));
        ptr = nullptr;
        bytes = 0;
        size_t want = need + need / 8 + 4096;
        CUDA_TRY(cudaHostAlloc(&ptr, want, cudaHostAllocDefault));
        bytes = want;
        return DEODR_B200_OK;
    }
};

// one scene array as staged: where it lives in the pinned mirror and on the device
struct MirrorSlot {
    size_t offset = 0;  // byte offset in the mirror
    size_t count = 0;   // elements
    bool as_float = false;  // user fp64 -> fp32 on the device
    size_t elem = 0;    // bytes per element in the mirror
};

enum { SL_FACES, SL_FACES_UV, SL_IJ, SL_DEPTHS, SL_UV, SL_COLORS, SL_SHADE, SL_EDGEFLAGS, SL_TEXTURED, SL_SHADED,
       SL_TEXTURE, SL_BACKGROUND, SL_COUNT };

struct HostPath {
    CopyPool pool;
    PinnedBuf mirror;   // canonical-layout copy of the last staged scene
    PinnedBuf staging;  // image_b upload / image, z, gradient download
    MirrorSlot slot[SL_COUNT];
    DeodrHostScene meta;  // scalar fields of the last staged scene (pointers unused)
    double sigma = -1;
    bool valid = false;   // mirror + device state describe a completed forward pass
    DeodrSceneView view;  // device view of the staged scene
    cudaStream_t stream = nullptr;
    cudaEvent_t chunk_event[64];
    HostPath() : pool(std::max(1, std::min(15, (int)std::thread::hardware_concurrency() / 2 - 1))) {
        memset(&meta, 0, sizeof(meta));
        memset(&view, 0, sizeof(view));
    }
};

void deodr_host_path_destroy(DeodrWorkspace *ws) {
    if (!ws || !ws->host) return;
    HostPath *h = ws->host;
    if (h->mirror.ptr) cudaFreeHost(h->mirror.ptr);
    if (h->staging.ptr) cudaFreeHost(h->staging.ptr);
    if (h->stream) {
        for (auto &e : h->chunk_event) cudaEventDestroy(e);
        cudaStreamDestroy(h->stream);
    }
    delete h;
    ws->host = nullptr;
}

static int host_path(DeodrWorkspace *ws, HostPath **out) {
    if (!ws->host) {
        HostPath *h = new (std::nothrow) HostPath();
        if (!h) return set_error(DEODR_B200_ENOMEM, "out of host memory");
        CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        for (auto &e : h->chunk_event) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ws->host = h;
    }
    *out = ws->host;
    return DEODR_B200_OK;
}

// ------------------------------------------------------------------------------------------ parallel host kernels

static void par_copy(CopyPool &pool, void *dst, const void *src, size_t bytes) {
    pool.run([&](int part, int nparts) {
        size_t lo, hi;
        part_range(bytes, part, nparts, &lo, &hi);
        if (hi > lo) memcpy((char *)dst + lo, (const char *)src + lo, hi - lo);
    });
}

static void par_f64_to_f32(CopyPool &pool, float *dst, const double *src, size_t n) {
    pool.run([&](int part, int nparts) {
        size_t lo, hi;
        part_range(n, part, nparts, &lo, &hi);
        for (size_t i = lo; i < hi; i++) dst[i] = (float)src[i];
    });
}

static void par_f32_to_f64(CopyPool &pool, double *dst, const float *src, size_t n) {
    pool.run([&](int part, int nparts) {
        size_t lo, hi;
        part_range(n, part, nparts, &lo, &hi);
        for (size_t i = lo; i < hi; i++) dst[i] = (double)src[i];
    });
}

static void par_f32_add_to_f64(CopyPool &pool, double *dst, const float *src, size_t n) {
    pool.run([&](int part, int nparts) {
        size_t lo, hi;
        part_range(n, part, nparts, &lo, &hi);
        for (size_t i = lo; i < hi; i++) dst[i] += (double)src[i];
    });
}

static bool par_equal_raw(CopyPool &pool, const void *a, const void *b, size_t bytes) {
    std::vector<int> same(pool.size(), 1);
    pool.run([&](int part, int nparts) {
        size_t lo, hi;
        part_range(bytes, part, nparts, &lo, &hi);
        if (hi > lo && memcmp((const char *)a + lo, (const char *)b + lo, hi - lo) != 0) same[part] = 0;
    });
    return std::all_of(same.begin(), same.end(), [](int v) { return v != 0; });
}

static bool par_equal_as_float(CopyPool &pool, const double *user, const float *mirror, size_t n) {
    std::vector<int> same(pool.size(), 1);
    pool.run([&](int part, int nparts) {
        size_t lo, hi;
        part_range(n, part, nparts, &lo, &hi);
        int ok = 1;
        for (size_t i = lo; i < hi; i++) {
            float f = (float)user[i];
            ok &= (memcmp(&f, &mirror[i], sizeof(float)) == 0);
        }
        same[part] = ok;
    });
    return std::all_of(same.begin(), same.end(), [](int v) { return v != 0; });
}

// ------------------------------------------------------------------------------------------------- scene staging

static int check_host_pointers(const DeodrHostScene *h, bool backward) {
    if (!h) return set_error(DEODR_B200_EINVAL, "scene == NULL");
    if (!h->faces) return set_error(DEODR_B200_EINVAL, "scene.faces == NULL");
    if (!h->faces_uv) return set_error(DEODR_B200_EINVAL, "scene.faces_uv == NULL");
    if (!h->depths) return set_error(DEODR_B200_EINVAL, "scene.depths == NULL");
    if (!h->uv) return set_error(DEODR_B200_EINVAL, "scene.uv == NULL");
    if (!h->ij) return set_error(DEODR_B200_EINVAL, "scene.ij == NULL");
    if (!h->shade) return set_error(DEODR_B200_EINVAL, "scene.shade == NULL");
    if (!h->colors) return set_error(DEODR_B200_EINVAL, "scene.colors == NULL");
    if (!h->edgeflags) return set_error(DEODR_B200_EINVAL, "scene.edgeflags == NULL");
    if (!h->textured) return set_error(DEODR_B200_EINVAL, "scene.textured == NULL");
    if (!h->shaded) return set_error(DEODR_B200_EINVAL, "scene.shaded == NULL");
    if (!h->texture) return set_error(DEODR_B200_EINVAL, "scene.texture == NULL");
    if (!h->background_image && !h->background_color)
        return set_error(DEODR_B200_EINVAL, "scene.background == NULL and scene.background_color == NULL");
    if (backward) {
        if (!h->uv_b) return set_error(DEODR_B200_EINVAL, "scene.uv_b == NULL");
        if (!h->ij_b) return set_error(DEODR_B200_EINVAL, "scene.ij_b == NULL");
        if (!h->shade_b) return set_error(DEODR_B200_EINVAL, "scene.shade_b == NULL");
        if (!h->colors_b) return set_error(DEODR_B200_EINVAL, "scene.colors_b == NULL");
        if (!h->texture_b) return set_error(DEODR_B200_EINVAL, "scene.texture_b == NULL");
    }
    if (h->nb_triangles < 0 || h->nb_vertices < 0 || h->nb_uv < 0 || h->height <= 0 || h->width <= 0 ||
        h->nb_colors <= 0 || h->texture_height < 0 || h->texture_width < 0)
        return set_error(DEODR_B200_EINVAL, "negative or zero size");
    return DEODR_B200_OK;
}

struct UserArrays {
    const void *ptr[SL_COUNT];
};

static void describe(const DeodrHostScene *h, MirrorSlot slot[SL_COUNT], UserArrays *u, size_t *total) {
    const size_t T = h->nb_triangles, V = h->nb_vertices, U = h->nb_uv, C = h->nb_colors;
    const size_t P = (size_t)h->height * h->width, tex = (size_t)h->texture_height * h->texture_width * C;
    const size_t counts[SL_COUNT] = {3 * T, 3 * T, 2 * V, V, 2 * U, V * C, V, 3 * T, T, T, tex,
                                     h->background_image ? P * C : C};
    const bool as_float[SL_COUNT] = {false, false, false, false, false, true, true, false, false, false, true, true};
    const size_t elem[SL_COUNT] = {4, 4, 8, 8, 8, 4, 4, 1, 1, 1, 4, 4};
    const void *ptrs[SL_COUNT] = {h->faces, h->faces_uv, h->ij, h->depths, h->uv, h->colors, h->shade, h->edgeflags,
                                  h->textured, h->shaded, h->texture,
                                  h->background_image ? (const void *)h->background_image : (const void *)h->background_color};
    size_t off = 0;
    for (int i = 0; i < SL_COUNT; i++) {
        slot[i].offset = off;
        slot[i].count = counts[i];
        slot[i].as_float = as_float[i];
        slot[i].elem = elem[i];
        off += (counts[i] * elem[i] + 255) & ~(size_t)255;
        u->ptr[i] = ptrs[i];
    }
    *total = off;
}

static bool same_meta(const DeodrHostScene &a, const DeodrHostScene &b) {
    return a.nb_triangles == b.nb_triangles && a.nb_vertices == b.nb_vertices && a.clockwise == b.clockwise &&
           a.backface_culling == b.backface_culling && a.nb_uv == b.nb_uv && a.height == b.height &&
           a.width == b.width && a.nb_colors == b.nb_colors && a.texture_height == b.texture_height &&
           a.texture_width == b.texture_width && (a.background_image != nullptr) == (b.background_image != nullptr) &&
           a.strict_edge == b.strict_edge && a.perspective_correct == b.perspective_correct &&
           a.integer_pixel_centers == b.integer_pixel_centers;
}

// true iff the caller's scene is bit-for-bit what the mirror (hence the device) already holds
static bool scene_matches_mirror(HostPath *hp, const DeodrHostScene *h, double sigma) {
    if (!hp->valid || hp->sigma != sigma || !same_meta(hp->meta, *h)) return false;
    MirrorSlot slot[SL_COUNT];
    UserArrays u;
    size_t total;
    describe(h, slot, &u, &total);
    for (int i = 0; i < SL_COUNT; i++) {
        if (slot[i].count != hp->slot[i].count) return false;
        const char *m = (const char *)hp->mirror.ptr + hp->slot[i].offset;
        bool eq = slot[i].as_float ? par_equal_as_float(hp->pool, (const double *)u.ptr[i], (const float *)m, slot[i].count)
                                   : par_equal_raw(hp->pool, u.ptr[i], m, slot[i].count * slot[i].elem);
        if (!eq) return false;
    }
    return true;
}

static int stage_scene(DeodrWorkspace *ws, HostPath *hp, const DeodrHostScene *h, double sigma) {
    hp->valid = false;
    UserArrays u;
    size_t total;
    describe(h, hp->slot, &u, &total);
    if (int rc = hp->mirror.ensure(total)) return rc;
    DevBuf *dev[SL_COUNT] = {&ws->h_faces, &ws->h_faces_uv, &ws->h_ij, &ws->h_depths, &ws->h_uv, &ws->h_colors,
                             &ws->h_shade, &ws->h_edgeflags, &ws->h_textured, &ws->h_shaded, &ws->h_texture,
                             &ws->h_background};
    cudaStream_t st = hp->stream;
    for (int i = 0; i < SL_COUNT; i++) {
        const MirrorSlot &s = hp->slot[i];
        const size_t bytes = s.count * s.elem;
        if (dev[i]->ensure(bytes + 16, &ws->bytes)) return DEODR_B200_ECUDA;
        if (bytes == 0) continue;
        char *m = (char *)hp->mirror.ptr + s.offset;
        if (s.as_float) par_f64_to_f32(hp->pool, (float *)m, (const double *)u.ptr[i], s.count);
        else par_copy(hp->pool, m, u.ptr[i], bytes);
        CUDA_TRY(cudaMemcpyAsync(dev[i]->ptr, m, bytes, cudaMemcpyHostToDevice, st));  // overlaps the next array's copy
    }
    DeodrSceneView *v = &hp->view;
    memset(v, 0, sizeof(*v));
    v->faces = ws->h_faces.as<uint32_t>();
    v->faces_uv = ws->h_faces_uv.as<uint32_t>();
    v->ij = ws->h_ij.as<double>();
    v->depths = ws->h_depths.as<double>();
    v->uv = ws->h_uv.as<double>();
    v->colors = ws->h_colors.as<float>();
    v->shade = ws->h_shade.as<float>();
    v->edgeflags = ws->h_edgeflags.as<uint8_t>();
    v->textured = ws->h_textured.as<uint8_t>();
    v->shaded = ws->h_shaded.as<uint8_t>();
    v->texture = ws->h_texture.as<float>();
    if (h->background_image) v->background_image = ws->h_background.as<float>();
    else v->background_color = ws->h_background.as<float>();
    v->nb_triangles = h->nb_triangles; v->nb_vertices = h->nb_vertices; v->nb_uv = h->nb_uv;
    v->height = h->height; v->width = h->width; v->nb_colors = h->nb_colors;
    v->texture_height = h->texture_height; v->texture_width = h->texture_width;
    v->clockwise = h->clockwise; v->backface_culling = h->backface_culling; v->strict_edge = h->strict_edge;
    v->perspective_correct = h->perspective_correct; v->integer_pixel_centers = h->integer_pixel_centers;
    hp->meta = *h;
    hp->sigma = sigma;
    return DEODR_B200_OK;
}

// stage + forward into the workspace's device framebuffers; on success the forward state is cached
static int host_forward(DeodrWorkspace *ws, HostPath *hp, const DeodrHostScene *h, double sigma) {
    if (int rc = stage_scene(ws, hp, h, sigma)) return rc;
    const size_t P = (size_t)h->height * h->width, C = h->nb_colors;
    int rc = 0;
    rc |= ws->h_image.ensure(P * C * sizeof(float), &ws->bytes);
    rc |= ws->h_z.ensure(P * sizeof(double), &ws->bytes);
    rc |= ws->h_owner.ensure(P * sizeof(int), &ws->bytes);
    if (rc) return DEODR_B200_ECUDA;
    rc = deodr_render_impl(ws, &hp->view, sigma, ws->h_image.as<float>(), ws->h_z.as<double>(), ws->h_owner.as<int>(),
                           nullptr, hp->stream, /*check_indices=*/true);
    if (rc) return rc;
    hp->valid = true;
    return DEODR_B200_OK;
}

// device -> user, chunked: DMA into pinned staging, copy threads convert / copy out while the next chunk is in flight.
// queue_download enqueues every DMA of one buffer (events [ev0, ev0 + n_chunks)), finish_download consumes them.
struct Download {
    const void *dev;
    void *user;
    size_t count, chunk;
    bool f32_to_f64, accumulate;
    char *staging;
    int ev0, n_chunks;
};

static int queue_download(HostPath *hp, Download *d, int max_events, cudaStream_t st) {
    const size_t elem = d->f32_to_f64 ? 4 : 8;
    d->chunk = std::max((size_t)4 << 20, (d->count + max_events - 1) / max_events);
    d->n_chunks = (int)((d->count + d->chunk - 1) / d->chunk);
    for (int c = 0; c < d->n_chunks; c++) {
        size_t lo = c * d->chunk, n = std::min(d->chunk, d->count - lo);
        CUDA_TRY(cudaMemcpyAsync(d->staging + lo * elem, (const char *)d->dev + lo * elem, n * elem,
                                 cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaEventRecord(hp->chunk_event[d->ev0 + c], st));
    }
    return DEODR_B200_OK;
}

static int finish_download(HostPath *hp, const Download *d) {
    for (int c = 0; c < d->n_chunks; c++) {
        size_t lo = c * d->chunk, n = std::min(d->chunk, d->count - lo);
        CUDA_TRY(cudaEventSynchronize(hp->chunk_event[d->ev0 + c]));
        if (d->f32_to_f64) {
            const float *src = (const float *)(d->staging + lo * 4);
            if (d->accumulate) par_f32_add_to_f64(hp->pool, (double *)d->user + lo, src, n);
            else par_f32_to_f64(hp->pool, (double *)d->user + lo, src, n);
        } else {
            par_copy(hp->pool, (char *)d->user + lo * 8, d->staging + lo * 8, n * 8);
        }
    }
    return DEODR_B200_OK;
}

extern "C" {

int deodr_b200_render_host(DeodrWorkspace *ws, const DeodrHostScene *scene, double *image, double *z_buffer,
                           double sigma, int antialiase_error, const double *obs, double *err_buffer) {
    (void)obs; (void)err_buffer;
    if (!ws) return set_error(DEODR_B200_EINVAL, "ws == NULL");
    if (antialiase_error)
        return set_error(DEODR_B200_EUNSUPPORTED, "antialiase_error mode is not implemented by deodr_b200 yet");
    if (int rc = check_host_pointers(scene, false)) return rc;
    if (!image) return set_error(DEODR_B200_EINVAL, "image_ptr is NULL");
    if (!z_buffer) return set_error(DEODR_B200_EINVAL, "z_buffer_ptr is NULL");
    CUDA_TRY(cudaSetDevice(ws->device));
    HostPath *hp;
    HostTrace trace("render_host");
    if (int rc = host_path(ws, &hp)) return rc;
    if (int rc = host_forward(ws, hp, scene, sigma)) return rc;
    trace.lap("stage + forward (enqueued)");
    const size_t P = (size_t)scene->height * scene->width, C = scene->nb_colors;
    if (int rc = hp->staging.ensure(P * C * 4 + P * 8 + 512)) return rc;
    char *stage_image = (char *)hp->staging.ptr, *stage_z = stage_image + ((P * C * 4 + 255) & ~(size_t)255);
    // both DMAs are queued before the first chunk is consumed
    Download d_image{ws->h_image.ptr, image, P * C, 0, true, false, stage_image, 0, 0};
    Download d_z{ws->h_z.ptr, z_buffer, P, 0, false, false, stage_z, 40, 0};
    if (int rc = queue_download(hp, &d_image, 40, hp->stream)) return rc;
    if (int rc = queue_download(hp, &d_z, 24, hp->stream)) return rc;
    if (int rc = finish_download(hp, &d_image)) return rc;
    trace.lap("image DMA + fp32->fp64");
    if (int rc = finish_download(hp, &d_z)) return rc;
    trace.lap("z DMA + copy");
    return DEODR_B200_OK;
}

int deodr_b200_render_b_host(DeodrWorkspace *ws, const DeodrHostScene *scene, double *image, double *z_buffer,
                             double *image_b, double sigma, int antialiase_error, const double *obs,
                             double *err_buffer, double *err_buffer_b) {
    (void)obs; (void)err_buffer; (void)err_buffer_b; (void)image; (void)z_buffer;
    if (!ws) return set_error(DEODR_B200_EINVAL, "ws == NULL");
    if (antialiase_error)
        return set_error(DEODR_B200_EUNSUPPORTED, "antialiase_error mode is not implemented by deodr_b200 yet");
    if (int rc = check_host_pointers(scene, true)) return rc;
    if (!scene->backface_culling)
        return set_error(DEODR_B200_EUNSUPPORTED, "You have to use backface_culling true if you ant to compute gradients");
    if (scene->perspective_correct)
        return set_error(DEODR_B200_EUNSUPPORTED,
                         "backward gradient propagation not supported yet with perspective_correct=True");
    if (!image_b) return set_error(DEODR_B200_EINVAL, "image_b_ptr is NULL");
    CUDA_TRY(cudaSetDevice(ws->device));
    HostPath *hp;
    HostTrace trace("render_b_host");
    if (int rc = host_path(ws, &hp)) return rc;
    cudaStream_t st = hp->stream;
    const size_t P = (size_t)scene->height * scene->width, C = scene->nb_colors;
    const size_t V = scene->nb_vertices, U = scene->nb_uv;
    const size_t tex = (size_t)scene->texture_height * scene->texture_width * C;
    const size_t n_ij = 2 * V, n_col = V * C, n_uv = 2 * U, n_sh = V, n_grad = n_ij + n_col + n_uv + n_sh + tex;
    if (int rc = hp->staging.ensure(std::max(P * C * 4, n_grad * 4) + 512)) return rc;

    // image_b: fp64 -> fp32 into pinned staging in chunks, each chunk DMA'd while the next is converted
    if (ws->h_image_b.ensure(P * C * sizeof(float), &ws->bytes)) return DEODR_B200_ECUDA;
    {
        const size_t chunk = (size_t)4 << 20, count = P * C;
        for (size_t lo = 0; lo < count; lo += chunk) {
            size_t n = std::min(chunk, count - lo);
            par_f64_to_f32(hp->pool, (float *)hp->staging.ptr + lo, image_b + lo, n);
            CUDA_TRY(cudaMemcpyAsync(ws->h_image_b.as<float>() + lo, (float *)hp->staging.ptr + lo, n * 4,
                                     cudaMemcpyHostToDevice, st));
        }
    }
    trace.lap("image_b fp64->fp32 + DMA");
    // forward state: reuse the cached one iff the caller's scene is bit-identical to the last forward's
    if (!scene_matches_mirror(hp, scene, sigma)) {
        if (int rc = host_forward(ws, hp, scene, sigma)) return rc;
        trace.lap("scene changed: re-forward");
    }
    trace.lap("scene == mirror check");
    if (ws->h_grads.ensure(n_grad * sizeof(float), &ws->bytes)) return DEODR_B200_ECUDA;
    CUDA_TRY(cudaMemsetAsync(ws->h_grads.ptr, 0, n_grad * sizeof(float), st));
    DeodrGrads g;
    g.ij_b = ws->h_grads.as<float>();
    g.colors_b = g.ij_b + n_ij;
    g.uv_b = g.colors_b + n_col;
    g.shade_b = g.uv_b + n_uv;
    g.texture_b = g.shade_b + n_sh;
    if (int rc = deodr_b200_render_b(ws, &hp->view, sigma, ws->h_z.as<double>(), ws->h_owner.as<int>(),
                                     ws->h_image_b.as<float>(), &g, st))
        return rc;
    // gradients are ACCUMULATED into scene.*_b (DR.h:3019-3049, 3126-3128)
    CUDA_TRY(cudaStreamSynchronize(st));  // staging is reused: image_b DMA must have completed
    trace.lap("backward kernels");
    double *dst[5] = {scene->ij_b, scene->colors_b, scene->uv_b, scene->shade_b, scene->texture_b};
    const size_t cnt[5] = {n_ij, n_col, n_uv, n_sh, tex};
    size_t off = 0;
    Download d[5];
    for (int i = 0; i < 5; i++) {
        d[i] = Download{ws->h_grads.as<float>() + off, dst[i], cnt[i], 0, true, true, (char *)hp->staging.ptr + off * 4,
                        12 * i, 0};
        if (cnt[i])
            if (int rc = queue_download(hp, &d[i], 12, st)) return rc;
        off += cnt[i];
    }
    for (int i = 0; i < 5; i++)
        if (cnt[i])
            if (int rc = finish_download(hp, &d[i])) return rc;
    trace.lap("gradients DMA + accumulate");
    return DEODR_B200_OK;
}

}  // extern "C"
