// deodr_b200: sm_100a kernels + C-ABI (include/deodr_b200.h) of the differentiable rasteriser.
//
// Pipeline of one forward pass (all on the caller's stream):
//   bin_tri(count) -> scan -> bin_tri(fill)                       per-tile triangle lists (unordered: the z test is
//                                                                  order independent, see phase_tri_test)
//   select silhouette edges -> depth keys -> stable radix sort    far-to-near order of DR.h:2781 (CUB)
//   bin_edge(count) -> scan -> bin_edge(fill) -> sort_tile_edges  per-tile edge lists in far-to-near order
//   raster_fwd                                                    one CTA per 16x16 tile: z-buffer, owner ids, colour,
//                                                                  ordered edge overdraw, single framebuffer write
// and of one backward pass:
//   raster_bwd                                                    per tile: edge replay + reverse sweep, interior
//                                                                  adjoint scattered to the vertices
//   finalize_edges                                                per edge: plane adjoints -> vertex adjoints
//
// There is no CPU fallback: every entry point fails with DEODR_B200_ECUDA if no device is usable.
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>

#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <limits>
#include <new>
#include <vector>

#include "../../include/deodr_b200.h"
#include "phases.h"
#include "workspace.h"

using namespace deodr;

// ------------------------------------------------------------------------------------------------ device Env

struct DevEnv {
    static __device__ __forceinline__ int atomic_add(int *p, int v) { return atomicAdd(p, v); }
    static __device__ __forceinline__ void atomic_add(float *p, float v) { atomicAdd(p, v); }
    static __device__ __forceinline__ void atomic_add(double *p, double v) { atomicAdd(p, v); }
    static __device__ __forceinline__ int shared_inc(int *p) { return atomicAdd(p, 1); }  // p in shared memory
};

// Interior adjoint of one pixel per lane with a warp-level reduce-by-owner before the scatter: the lanes of a warp that
// share the adjoint owner (neighbouring pixels of a large triangle) first sum their vertex gradients with a tree of
// shuffles over the group (__match_any_sync gives the groups), then ONE lane per group issues the atomics.  Measured on
// the 1M-triangle scene: float-atomic throughput was 38 of k_interior_bwd's 67 us.  Must be called by all 32 lanes.
template <int MAXC>
static __device__ __forceinline__ void interior_adjoint_warp(const SceneView &s, int x, int y, bool has,
                                                             const PixelState<MAXC> &p, const float *g,
                                                             const DeodrGrads &grads) {
    const int lane = (int)(threadIdx.x & 31), C = s.nb_colors;
    TriAttr t;
    VertexGrads<MAXC> acc;
    zero_vertex_grads<MAXC>(s, &acc);
    t.textured = false;
    if (has) {
        tri_attr(s, p.bown & TRI_INDEX_MASK, &t);
        pixel_adjoint<MAXC, DevEnv>(s, t, x, y, g, &acc, grads.texture_b);
    }
    const int key = has ? (p.bown & TRI_INDEX_MASK) : -1 - lane;  // idle lanes: groups of one
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const int rank = __popc(peers & ((1u << lane) - 1u)), size = __popc(peers);
    const int max_size = __reduce_max_sync(0xffffffffu, size);
    const bool any_textured = __any_sync(0xffffffffu, has && t.textured);
    const bool any_plain = __any_sync(0xffffffffu, has && !t.textured);
    for (int stride = 1; stride < max_size; stride <<= 1) {
        // tree over the members of a group: member `rank` (a multiple of 2*stride) takes member rank + stride
        const bool take = (rank & (2 * stride - 1)) == 0 && rank + stride < size;
        const int src = take ? (int)__fns(peers, 0, rank + stride + 1) : lane;
#define DEODR_TAKE(field)                                                  \
        {                                                                      \
            const float other = __shfl_sync(0xffffffffu, (field), src);        \
            if (take) (field) += other;                                        \
        }
#pragma unroll
        for (int i = 0; i < 3; i++) {
            DEODR_TAKE(acc.ij[i][0]);
            DEODR_TAKE(acc.ij[i][1]);
        }
        if (any_textured) {
#pragma unroll
            for (int i = 0; i < 3; i++) {
                DEODR_TAKE(acc.uv[i][0]);
                DEODR_TAKE(acc.uv[i][1]);
                DEODR_TAKE(acc.shade[i]);
            }
        }
        if (any_plain) {
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int c = 0; c < MAXC; c++)
                    if (c < C) DEODR_TAKE(acc.attr[i][c]);
        }
#undef DEODR_TAKE
    }
    if (has && rank == 0)
        flush_vertex_grads<MAXC, AtomicEmit<DevEnv>>(s, t, acc, grads.ij_b, grads.colors_b, grads.uv_b, grads.shade_b,
                                                     AtomicEmit<DevEnv>());
}

// The <1> and <3> instances are only launched for exactly 1 / 3 colour channels: telling the compiler turns every
// `for (c < nb_colors)` loop of the inlined shading code into straight-line code (the <4> and <16> instances keep the
// run-time channel count: 2 or 4, 5..16).
template <int MAXC, bool TEX = true>
static __device__ __forceinline__ void fix_channel_count(SceneView &s) {
    if (MAXC == 1 || MAXC == 3) s.nb_colors = MAXC;
    // TEX = false instances are launched for scenes without a textured triangle (flag raised by k_bin_count): nulling
    // the texture pointer of the kernel's copy of the scene folds every texture branch (tri_attr / edge_hit test it)
    if (!TEX) s.texture = nullptr;
}

static_assert(sizeof(SceneView) == sizeof(DeodrSceneView), "SceneView must mirror DeodrSceneView");

// ------------------------------------------------------------------------------------------------- kernels

// Count pass: one thread per triangle (small / large tile counts, silhouette-edge append, edge tile counts).
__global__ void k_bin_count(SceneView s, double sigma, int tiles_x, TriBins bins, TriLists lists, EdgeList edges,
                            int *edge_tile_count, const int *bad_indices, int *any_textured) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= s.nb_triangles) return;
    if (bad_indices && *bad_indices) return;  // out-of-range face indices found by k_check_scene: touch nothing
    // does the scene hold ANY textured triangle?  (read back with the list sizes: scenes without one run kernel
    // instances compiled without the texture paths - fewer registers, no local memory)
    if (s.textured[k] && s.shaded[k] && *(volatile int *)any_textured == 0) atomicOr(any_textured, 1);
    bin_count_triangle<DevEnv>(s, k, sigma, tiles_x, bins, lists, edges, edge_tile_count);
}

// Exclusive scans of the tile counts: blockIdx 0 small triangles (records), 1 large triangles, 2 silhouette edges.
// 1024 threads per CTA, coalesced loads with a one-chunk prefetch; offsets[n] and totals[] receive the grand totals.
struct ScanJob {
    const int *count[3];
    int *offset[3];
    int total_slot[3];
    int *nonempty[3];        // optional: compact list of the tiles with count > 0 (nullptr to skip)
    int nonempty_slot[3];    // totals[] slot receiving the length of that list
    // two-ended list (threshold > 0): tiles with count > threshold fill the list from the front, the other non-empty
    // tiles from the back (position n-1 downwards), so that the kernels that walk it start the crowded tiles first
    int heavy_threshold[3];
    int heavy_slot[3];       // totals[] slot receiving the number of crowded tiles
};

// Block b of a kernel that walks a two-ended tile list of capacity n with `heavy` crowded tiles at its front.
static __device__ __forceinline__ int two_ended_at(const int *list, int n, int heavy, int b) {
    return b < heavy ? list[b] : list[n - 1 - (b - heavy)];
}

// One 64-bit scan carries both the running sum of the counts (low word) and the number of non-empty tiles (high word).
// Each thread owns SCAN_IPT consecutive tiles, so 16384 tiles take ONE block scan instead of sixteen; counts and
// offsets pass through a padded shared-memory stage so that every global access stays coalesced (the kernel runs on
// three SMs only: 16 uncoalesced wavefronts per load instruction would be its whole duration).
constexpr int SCAN_IPT = 16;
constexpr int SCAN_TILE = 1024 * SCAN_IPT;
constexpr int SCAN_SMEM = (SCAN_TILE + SCAN_TILE / 32) * (int)sizeof(int);
static __device__ __forceinline__ int scan_slot(int i) { return i + (i >> 5); }  // conflict-free for stride-16 readers
__global__ void __launch_bounds__(1024) k_scan_tiles(ScanJob job, int n, int *totals, int *host_totals, int seq) {
    extern __shared__ int stage[];
    __shared__ unsigned long long warp_sums[32];
    __shared__ int heavy_sums[32];
    const int threshold = job.heavy_threshold[blockIdx.x];
    int heavy_carry = 0;
    const int *count = job.count[blockIdx.x];
    int *offset = job.offset[blockIdx.x];
    int *nonempty = job.nonempty[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long carry = 0;
    for (int base = 0; base < n; base += SCAN_TILE) {
        {
            // all sixteen loads first (the compiler cannot move a generic-pointer load above a shared-memory store on
            // its own: interleaved they would cost sixteen serial memory round trips)
            int in[SCAN_IPT];
#pragma unroll
            for (int k = 0; k < SCAN_IPT; k++) in[k] = base + k * 1024 + tid < n ? __ldg(count + base + k * 1024 + tid) : 0;
#pragma unroll
            for (int k = 0; k < SCAN_IPT; k++) stage[scan_slot(k * 1024 + tid)] = in[k];
        }
        __syncthreads();
        const int first = tid * SCAN_IPT;
        int c[SCAN_IPT];
#pragma unroll
        for (int k = 0; k < SCAN_IPT; k++) c[k] = stage[scan_slot(first + k)];
        unsigned long long v = 0;
#pragma unroll
        for (int k = 0; k < SCAN_IPT; k++) v += (unsigned long long)(unsigned)c[k] | ((unsigned long long)(c[k] > 0) << 32);
        int hv = 0;
        if (threshold > 0) {
#pragma unroll
            for (int k = 0; k < SCAN_IPT; k++) hv += (int)(c[k] > threshold);
        }
        unsigned long long incl = v;
        int hincl = hv;
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            int ht = __shfl_up_sync(0xffffffffu, hincl, o);
            if (lane >= o) { incl += t; hincl += ht; }
        }
        if (lane == 31) { warp_sums[warp] = incl; heavy_sums[warp] = hincl; }
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = warp_sums[lane];
            int hw = heavy_sums[lane];
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long t = __shfl_up_sync(0xffffffffu, w, o);
                int ht = __shfl_up_sync(0xffffffffu, hw, o);
                if (lane >= o) { w += t; hw += ht; }
            }
            warp_sums[lane] = w;
            heavy_sums[lane] = hw;
        }
        __syncthreads();
        int hrun = heavy_carry + (warp ? heavy_sums[warp - 1] : 0) + hincl - hv;
        unsigned long long run = carry + (warp ? warp_sums[warp - 1] : 0ull) + incl - v;
#pragma unroll
        for (int k = 0; k < SCAN_IPT; k++) {
            stage[scan_slot(first + k)] = (int)(unsigned)run;
            if (nonempty && c[k] > 0) {
                const int rank = (int)(run >> 32);  // non-empty tiles before this one
                if (threshold <= 0) nonempty[rank] = base + first + k;
                else if (c[k] > threshold) nonempty[hrun++] = base + first + k;
                else nonempty[n - 1 - (rank - hrun)] = base + first + k;
            }
            run += (unsigned long long)(unsigned)c[k] | ((unsigned long long)(c[k] > 0) << 32);
        }
        carry += warp_sums[31];
        heavy_carry += heavy_sums[31];
        __syncthreads();
        {
            int outv[SCAN_IPT];
#pragma unroll
            for (int k = 0; k < SCAN_IPT; k++) outv[k] = stage[scan_slot(k * 1024 + tid)];
#pragma unroll
            for (int k = 0; k < SCAN_IPT; k++)
                if (base + k * 1024 + tid < n) offset[base + k * 1024 + tid] = outv[k];
        }
        __syncthreads();
    }
    if (tid == 0) {
        offset[n] = (int)(unsigned)carry;
        totals[job.total_slot[blockIdx.x]] = (int)(unsigned)carry;
        if (nonempty) totals[job.nonempty_slot[blockIdx.x]] = (int)(carry >> 32);
        if (threshold > 0) totals[job.heavy_slot[blockIdx.x]] = heavy_carry;
        // The last CTA to finish publishes the sixteen scalars straight into the host's (pinned, device-visible) buffer
        // and raises a sequence flag the host is polling: no copy engine, no stream synchronisation on the critical path.
        __threadfence();
        const int ticket = atomicAdd(&totals[11], 1);
        if (host_totals && ticket == (int)gridDim.x - 1) {
            __threadfence();
            for (int i = 0; i < 16; i++) host_totals[i] = ((volatile int *)totals)[i];
            __threadfence_system();
            ((volatile int *)host_totals)[16] = seq;
        }
    }
}

// Far-to-near order of the appended silhouette edges by rank counting (DR.h:2781; ties by id).  2-D grid: CTA (bx, by)
// counts, for its 256 edges i, the edges j of chunk by (256 keys staged in shared memory) that precede them, and adds
// the partial count to rank[i]; k_scatter_edges then writes edge_sorted[rank[i]] = ids[i].
constexpr int RANK_CHUNK = 256;  // keys per CTA: E/256 x E/256 CTAs keep the whole chip busy for a few thousand edges
__global__ void __launch_bounds__(256) k_rank_edges(EdgeList edges, int n, int *rank) {
    __shared__ unsigned long long sk[RANK_CHUNK];
    __shared__ uint32_t shi[RANK_CHUNK];
    __shared__ int si[RANK_CHUNK];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int base = blockIdx.y * RANK_CHUNK, m = min(RANK_CHUNK, n - base);
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        sk[j] = edges.keys[base + j];
        shi[j] = (uint32_t)(sk[j] >> 32);
        si[j] = edges.ids[base + j];
    }
    __syncthreads();
    if (i >= n) return;
    const unsigned long long key = edges.keys[i];
    const uint32_t key_hi = (uint32_t)(key >> 32);
    const int id = edges.ids[i];
    // the high words (sign, exponent, 20 mantissa bits of the depth sum) decide almost every comparison: count them
    // with 32-bit compares; only when another edge shares the high word (same triangle, nearly equal depth sums, the
    // edge itself on the diagonal chunks) is the full (key, id) comparison run
    int partial = 0, equal = 0;
#pragma unroll 8
    for (int j = 0; j < m; j++) {
        partial += (int)(shi[j] < key_hi);
        equal += (int)(shi[j] == key_hi);
    }
    if (equal > 0)
        for (int j = 0; j < m; j++)
            if (shi[j] == key_hi) partial += (int)(sk[j] < key || (sk[j] == key && si[j] < id));
    if (partial) atomicAdd(&rank[i], partial);
}

// edge_sorted[rank] = id, and the edge's band stencil record (DR.h:1366-1460 + z plane) at the same rank: built ONCE
// per forward pass, not per tile.
__global__ void k_scatter_edges(SceneView s, EdgeList edges, int n, double sigma, int *edge_sorted, EdgeRec *recs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = edges.rank[i], id = edges.ids[i];
    edge_sorted[r] = id;
    edge_record(s, id, r, sigma, &recs[r]);
}

// One thread per silhouette edge (far-to-near rank r): band stencil of DR.h:1366-1460 + z plane, once per forward.
__global__ void k_edge_records(SceneView s, const int *edge_sorted, int n, double sigma, EdgeRec *recs) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) edge_record(s, edge_sorted[r], r, sigma, &recs[r]);
}

// Fill pass over the compacted lists: blocks [0, small_blocks) small triangles (pre-masked records), then large
// triangles (indices), then silhouette edges (ranks).
#ifndef DEODR_FILL_MIN_CTAS
#define DEODR_FILL_MIN_CTAS 8  // 64 registers: fill 64.3 us vs 68.9 us at 74 (measured, c5)
#endif
__global__ void __launch_bounds__(128, DEODR_FILL_MIN_CTAS) k_bin_fill(SceneView s, double sigma, int tiles_x, int small_blocks, int large_blocks, TriBins bins,
                           const int *small_ids, int num_small, const int *large_ids, int num_large,
                           const int *edge_sorted, int num_edges, const int *edge_offset, int *edge_cursor,
                           int *edge_refs) {
    int b = blockIdx.x;
    if (b < small_blocks) {
        int i = b * blockDim.x + threadIdx.x;
        if (i < num_small) bin_fill_small<DevEnv>(s, small_ids[i], tiles_x, bins);
        return;
    }
    b -= small_blocks;
    if (b < large_blocks) {
        int i = b * blockDim.x + threadIdx.x;
        if (i < num_large) bin_fill_large<DevEnv>(s, large_ids[i], tiles_x, bins);
        return;
    }
    b -= large_blocks;
    int r = b * blockDim.x + threadIdx.x;
    if (r < num_edges) bin_fill_edge<DevEnv>(s, edge_sorted[r], r, sigma, tiles_x, edge_offset, edge_cursor, edge_refs);
}

// ---------------------------------------------------------------------------- TMA (bulk async copy) + mbarrier

static __device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

static __device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

static __device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// global -> shared bulk copy (SASS: UBLKCP), completion signalled on `bar` as transaction bytes.
// dst, src 16-byte aligned, bytes a multiple of 16.
static __device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

static __device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// Orders every tile's edge list by far-to-near rank (ranks are unique): rank-counting sort, one CTA per tile.
__global__ void k_sort_tile_edges(const int *edge_tiles, int num_tiles, int heavy, const int *count, const int *offset,
                                  const int *refs_in, int *refs_out) {
    const int tile = two_ended_at(edge_tiles, num_tiles, heavy, blockIdx.x);
    const int n = count[tile], base = offset[tile];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int mine = refs_in[base + i], pos = 0;
        for (int j = 0; j < n; j++) pos += refs_in[base + j] < mine;
        refs_out[base + pos] = mine;
    }
}

struct TieTable {
    int *pairs;     // (own, bown) per tie pixel
    int *counter;   // number of entries requested so far (may exceed capacity: overflow)
    int capacity;
};

// Forward, kernel 1 of 3 - z-buffer and owner ids, one 16x16 tile per CTA (no colour work: few registers).
// Small triangles: the tile's pre-masked records are pulled into shared memory by a bulk copy (cp.async.bulk +
// mbarrier, SASS UBLKCP); two threads per record scatter its index into per-pixel candidate lists; each pixel then
// z-tests its own candidates exactly.  Large triangles: stencil per triangle -> coverage masks (8 threads each) ->
// mask test.  The kernel is written as a loop over tiles with a two-stage copy pipeline (the copy of the next tile in
// flight while the current one is tested) and launched with one tile per CTA: measured faster than persistent CTAs
// (DEODR_B200_TILEZ_CTAS_PER_SM / DEODR_B200_TILEZ_TILES_PER_CTA switch the other modes on for A/B runs).
#ifndef DEODR_TILEZ_MIN_CTAS
#define DEODR_TILEZ_MIN_CTAS 5  // 51 registers: 98.5 us vs 102.6 us at 64 and 123 us at 85 (measured, c5)
#endif
// PERSP: perspective_correct as a compile-time constant (the 1/z division and its registers leave the common instance)
template <bool PERSP>
__global__ void __launch_bounds__(NT, DEODR_TILEZ_MIN_CTAS) k_tile_z(SceneView s, TileDiv tiles_x, int num_tiles, TriBins bins, TieTable ties,
                                                  double *z_buffer, int *owner, int *face_id) {
    s.perspective_correct = PERSP ? 1 : 0;
    __shared__ TileShared sh;
    __shared__ alignas(16) PreRec pre[2][PRE_CHUNK];
    __shared__ alignas(8) uint64_t bar[2];
    __shared__ int info[2][2];  // [buffer][0] = number of small records of the tile, [1] = offset of its list
    const int tid = threadIdx.x;
    sh.tri.pix_cnt[tid] = 0;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
    }
    uint32_t parity0 = 0, parity1 = 0;

    // thread 0 runs a two-deep software pipeline: the list size / offset of tile i+2 are LOADED (registers nn, noff)
    // while the copy of tile i+1 is ISSUED from values loaded one iteration earlier, so the loads' latency never stalls
    // warp 0 (which also does pixel work).
    int nn = 0, noff = 0;
    auto load_info = [&](int t) {
        if (tid == 0 && t < num_tiles) { nn = bins.small_cursor[t]; noff = bins.small_offset[t]; }
    };
    auto issue = [&](int t, int b) {  // publish (nn, noff) for tile t and start the copy of its first chunk
        if (tid != 0 || t >= num_tiles) return;
        info[b][0] = nn;
        info[b][1] = noff;
        const int m = min(nn, PRE_CHUNK);
        if (m > 0) {
            mbar_expect_tx(&bar[b], (uint32_t)(m * sizeof(PreRec)));
            bulk_load(pre[b], bins.small_recs + noff, (uint32_t)(m * sizeof(PreRec)), &bar[b]);
        }
    };
    __syncthreads();  // barriers initialised
    int cur = 0;
    load_info(blockIdx.x);
    issue(blockIdx.x, 0);
    load_info(blockIdx.x + gridDim.x);
    for (int tile_id = blockIdx.x; tile_id < num_tiles; tile_id += gridDim.x, cur ^= 1) {
        __syncthreads();  // info[cur] is visible; everybody is done with buffer cur^1 and with sh (previous tile)
        const int n_small = info[cur][0];
        const PreRec *list = bins.small_recs + info[cur][1];
        issue(tile_id + gridDim.x, cur ^ 1);        // overlaps with the work on this tile
        load_info(tile_id + 2 * gridDim.x);         // consumed by the next iteration's issue()

        const Tile tile = tile_of(tile_id, tiles_x);
        const int x = tile.x0 + tid % TS, y = tile.y0 + tid / TS;
        const bool inside = x < s.width && y < s.height;
        PixelState<1> p;
        p.z = __longlong_as_double(0x7ff0000000000000LL);
        p.own = -1;
        p.bown = -1;

        for (int base = 0; base < n_small; base += PRE_CHUNK) {
            const int m = min(PRE_CHUNK, n_small - base);
            if (base > 0) {  // tiles with more than one chunk: the later chunks are fetched synchronously
                __syncthreads();  // buffer cur fully consumed
                if (tid == 0) {
                    mbar_expect_tx(&bar[cur], (uint32_t)(m * sizeof(PreRec)));
                    bulk_load(pre[cur], list + base, (uint32_t)(m * sizeof(PreRec)), &bar[cur]);
                }
            }
            if (cur == 0) { mbar_wait(&bar[0], parity0); parity0 ^= 1u; }
            else          { mbar_wait(&bar[1], parity1); parity1 ^= 1u; }
            phase_pre_scatter<DevEnv>(tid, m, pre[cur], &sh);
            __syncthreads();
            phase_pix_test<1>(s, tid, m, tile, pre[cur], &sh, &p);  // (pixels outside the image have no candidates)
            __syncthreads();
        }
        // large triangles: by index, stencil set-up and row spans computed here, one thread per triangle
        const int n_large = bins.large_count[tile_id];
        if (n_large > 0) {
            const int *large = bins.large_refs + bins.large_offset[tile_id];
            for (int base = 0; base < n_large; base += LARGE_CHUNK) {
                const int m = min(LARGE_CHUNK, n_large - base);
                phase_tri_setup(s, tid, m, large + base, &sh);
                __syncthreads();
                phase_tri_masks(s, tid, m, tile, &sh);
                __syncthreads();
                if (inside) phase_tri_test<1>(s, tid, m, tile, &sh, &p);
                __syncthreads();
            }
        }
        if (inside) {
            const size_t idx = (size_t)y * s.width + x;
            z_buffer[idx] = p.z;
            int code = p.bown;
            if (p.own != p.bown) {  // exact z tie between distinct triangles: keep both ids in the side table
                int slot = atomicAdd(ties.counter, 1);
                if (slot < ties.capacity) {
                    ties.pairs[2 * slot] = p.own;
                    ties.pairs[2 * slot + 1] = p.bown;
                    code = -2 - slot;
                }
            }
            owner[idx] = code;
            if (face_id) face_id[idx] = p.own >= 0 ? (p.own & TRI_INDEX_MASK) : -1;
        }
    }
}

// decodes an owner code into (forward owner, adjoint owner); -1 = background
static __device__ __forceinline__ void decode_owner(int code, const TieTable &ties, int *own, int *bown) {
    if (code <= -2) {
        *own = ties.pairs[2 * (-2 - code)];
        *bown = ties.pairs[2 * (-2 - code) + 1];
    } else {
        *own = *bown = code;
    }
}

// Forward, kernel 2 of 3 - colour of every pixel from its owner (one thread per pixel, tile-shaped blocks for
// locality of the vertex gathers).  Reads owner (and z with perspective_correct), writes image.
template <int MAXC, bool PERSP, bool TEX>
#ifndef DEODR_SHADE_MIN_CTAS
#define DEODR_SHADE_MIN_CTAS 6  // 40 registers: measured 55.5 us vs 58.6 us at 48 (5 CTAs / SM) and 67 us at 56
#endif
__global__ void __launch_bounds__(NT, DEODR_SHADE_MIN_CTAS) k_shade(SceneView s, TileDiv tiles_x, TieTable ties, const int *owner,
                                              const double *z_buffer, float *image) {
    fix_channel_count<MAXC, TEX>(s);
    s.perspective_correct = PERSP ? 1 : 0;
    const Tile tile = tile_of(blockIdx.x, tiles_x);
    const int x = tile.x0 + threadIdx.x % TS, y = tile.y0 + threadIdx.x / TS;
    if (x >= s.width || y >= s.height) return;
    const size_t idx = (size_t)y * s.width + x;
    PixelState<MAXC> p;
    decode_owner(owner[idx], ties, &p.own, &p.bown);
    p.z = s.perspective_correct && p.own >= 0 ? z_buffer[idx] : 0.0;
    phase_shade<MAXC>(s, x, y, &p);
    for (int k = 0; k < s.nb_colors; k++) image[idx * s.nb_colors + k] = p.col[k];
}


#ifndef DEODR_EDGE_MIN_CTAS
#define DEODR_EDGE_MIN_CTAS 3  // 85 registers: edge_bwd 62.9 us vs 68 us at 64, 77 us at 51 (measured, c5)
#endif
static_assert(EDGE_ROWS == TS, "the span cache shared by k_edge_fwd and k_raster_bwd holds whole tiles");

// Forward, kernel 3 of 3 - ordered silhouette-edge overdraw on the tiles that have edges (DR.h:2839-2899).
// One CTA per tile, the tiles with more than one chunk of edges first (two-ended list built by k_scan_tiles): the few
// crowded tiles set the kernel's duration.  Per chunk of <= 64 edges: records copied to shared memory, (edge, row) x
// spans computed (and saved for the adjoint pass), then every pixel blends the edges of its own 64-bit hit mask in
// far-to-near order.  (16x4 strips with four CTAs per tile were measured slower: the set-up work is per edge.)
template <int MAXC, bool PERSP, bool TEX>
__global__ void __launch_bounds__(EDGE_NT, DEODR_EDGE_MIN_CTAS) k_edge_fwd(SceneView s, double sigma, TileDiv tiles_x, const int *edge_tiles, int num_tiles, int heavy,
                                                      const int *edge_count, const int *edge_offset,
                                                      const int *edge_refs, const EdgeRec *edge_recs,
                                                      uint32_t *span_cache, const double *z_buffer, float *image) {
    fix_channel_count<MAXC, TEX>(s);
    s.perspective_correct = PERSP ? 1 : 0;
    __shared__ TileShared sh;
    const int tile_id = two_ended_at(edge_tiles, num_tiles, heavy, blockIdx.x / (TS / EDGE_ROWS)), tid = threadIdx.x;
    const int row0 = (blockIdx.x % (TS / EDGE_ROWS)) * EDGE_ROWS;
    const int n_edge = edge_count[tile_id];
    const Tile tile = tile_of(tile_id, tiles_x);
    const int r = row0 + tid / TS;
    const int x = tile.x0 + tid % TS, y = tile.y0 + r;
    const bool inside = x < s.width && y < s.height;
    const size_t idx = inside ? (size_t)y * s.width + x : 0;
    PixelState<MAXC> p;
    p.z = 0.0;
    p.own = p.bown = -1;
    if (inside) {
        p.z = z_buffer[idx];
        for (int k = 0; k < s.nb_colors; k++) p.col[k] = image[idx * s.nb_colors + k];
    }
    const int edge_base = edge_offset[tile_id];
    for (int base = 0; base < n_edge; base += EDGE_CHUNK) {
        const int m = min(EDGE_CHUNK, n_edge - base);
        phase_edge_setup(tid, EDGE_NT, m, edge_refs + edge_base + base, edge_recs, &sh);
        __syncthreads();
        phase_edge_spans(s, tid, EDGE_NT, m, tile, row0, EDGE_ROWS, &sh);
        __syncthreads();
        // the backward pass reuses the spans (same scene, same sigma): 64 bytes per (tile, edge), coalesced
        for (int item = tid; item < m * TS; item += EDGE_NT)
            span_cache[(size_t)(edge_base + base) * TS + item] = sh.edge.span[item / TS][item % TS];
        if (inside) phase_edge_blend<MAXC>(s, x, y, r, m, &sh, &p);
        __syncthreads();
    }
    if (inside)
        for (int k = 0; k < s.nb_colors; k++) image[idx * s.nb_colors + k] = p.col[k];
}

// (register budgets are pinned: the allocator's own choice moved 64 -> 80 on an unrelated signature change)
template <int MAXC, bool TEX>
__global__ void __launch_bounds__(EDGE_NT, DEODR_EDGE_MIN_CTAS) k_raster_bwd(SceneView s, double sigma, TileDiv tiles_x, const int *edge_tiles, int num_tiles, int heavy,
                                                   const int *edge_count, const int *edge_offset, const int *edge_refs,
                                                   const EdgeRec *edge_recs, const uint32_t *span_cache, TieTable ties,
                                                   const double *z_buffer,
                                                   const int *owner, const float *image_b, DeodrGrads grads,
                                                   double *edge_acc) {
    fix_channel_count<MAXC, TEX>(s);
    s.perspective_correct = 0;  // the adjoint is only defined without it (validate_view rejects it): folds the branches
    __shared__ TileShared sh;
    // one CTA per tile that HAS edges, crowded tiles first (see k_edge_fwd)
    const int tile_id = two_ended_at(edge_tiles, num_tiles, heavy, blockIdx.x / (TS / EDGE_ROWS)), tid = threadIdx.x;
    const int row0 = (blockIdx.x % (TS / EDGE_ROWS)) * EDGE_ROWS;
    const Tile tile = tile_of(tile_id, tiles_x);
    const int c = tid % TS, r = row0 + tid / TS;
    const int x = tile.x0 + c, y = tile.y0 + r;
    const bool inside = x < s.width && y < s.height;
    const size_t idx = inside ? (size_t)y * s.width + x : 0;
    const int n_edge = edge_count[tile_id];

    PixelState<MAXC> p;
    AdjointState<MAXC> a;
    a.has_colour = false;
    p.z = __longlong_as_double(0x7ff0000000000000LL);
    p.own = p.bown = -1;
    if (inside) {
        p.z = z_buffer[idx];
        int code = owner[idx];
        if (code <= -2) {
            p.own = ties.pairs[2 * (-2 - code)];
            p.bown = ties.pairs[2 * (-2 - code) + 1];
        } else {
            p.own = p.bown = code;
        }
        for (int k = 0; k < s.nb_colors; k++) a.g[k] = image_b[idx * s.nb_colors + k];
    }

    {
        const int edge_base = edge_offset[tile_id];
        const bool single = n_edge <= EDGE_CHUNK;
        auto load_spans = [&](int base, int m) {  // (all loads of a pass before its stores, see phase_edge_setup)
            const uint32_t *src = span_cache + (size_t)(edge_base + base) * TS;
            for (int first = tid; first < m * TS; first += 4 * EDGE_NT) {
                uint32_t v[4];
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (first + j * EDGE_NT < m * TS) v[j] = __ldg(src + first + j * EDGE_NT);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int item = first + j * EDGE_NT;
                    if (item < m * TS) sh.edge.span[item / TS][item % TS] = v[j];
                }
            }
        };
        // pass A: forward replay (far to near) to obtain the final colour in fp64
        for (int base = 0; base < n_edge; base += EDGE_CHUNK) {
            const int m = min(EDGE_CHUNK, n_edge - base);
            phase_edge_setup(tid, EDGE_NT, m, edge_refs + edge_base + base, edge_recs, &sh);
            load_spans(base, m);  // computed by the forward pass (k_edge_fwd)
            __syncthreads();
            if (inside) phase_edge_replay<MAXC>(s, x, y, r, m, &sh, p, &a);
            if (single) {
                if (inside && a.has_colour)
                    phase_edge_adjoint<MAXC, DevEnv>(s, x, y, r, m, &sh, p, &a, edge_acc, grads.texture_b);
            }
            __syncthreads();
        }
        // pass B: reverse sweep (near to far), chunks in reverse order
        if (!single) {
            const int last = ((n_edge - 1) / EDGE_CHUNK) * EDGE_CHUNK;
            for (int base = last; base >= 0; base -= EDGE_CHUNK) {
                const int m = min(EDGE_CHUNK, n_edge - base);
                phase_edge_setup(tid, EDGE_NT, m, edge_refs + edge_base + base, edge_recs, &sh);
                load_spans(base, m);
                __syncthreads();
                if (inside && a.has_colour)
                    phase_edge_adjoint<MAXC, DevEnv>(s, x, y, r, m, &sh, p, &a, edge_acc, grads.texture_b);
                __syncthreads();
            }
        }
    }
    interior_adjoint_warp<MAXC>(s, x, y, inside && p.bown >= 0, p, a.g, grads);
}

// Interior adjoint of the pixels owned by LARGE triangles in the tiles without silhouette edges: no shared memory, no
// z-buffer read; the gradients are summed per owner inside each warp before the scatter (interior_adjoint_warp).
template <int MAXC, bool TEX>
#ifndef DEODR_INTERIOR_MIN_CTAS
#define DEODR_INTERIOR_MIN_CTAS 3  // 85 registers: 39.2 us vs 40.3 us at 64 and 47.6 us at 51 (measured, c5)
#endif
__global__ void __launch_bounds__(NT, DEODR_INTERIOR_MIN_CTAS) k_interior_bwd(SceneView s, TileDiv tiles_x, const int *large_tiles,
                                                     const int *edge_count, TieTable ties, const int *owner,
                                                     const float *image_b, DeodrGrads grads) {
    fix_channel_count<MAXC, TEX>(s);
    s.perspective_correct = 0;  // the adjoint is only defined without it (validate_view rejects it): folds the branches
    const int tile_id = large_tiles[blockIdx.x], tid = threadIdx.x;  // one CTA per tile with large triangles binned
    if (edge_count && edge_count[tile_id] > 0) return;  // handled by k_raster_bwd
    const Tile tile = tile_of(tile_id, tiles_x);
    const int x = tile.x0 + tid % TS, y = tile.y0 + tid / TS;
    const bool inside = x < s.width && y < s.height;
    PixelState<MAXC> p;
    p.z = 0.0;  // only read by the perspective-correct forward path
    p.own = p.bown = -1;
    float g[MAXC];
    if (inside) {
        const size_t idx = (size_t)y * s.width + x;
        const int code = owner[idx];
        if (code <= -2) {
            p.own = ties.pairs[2 * (-2 - code)];
            p.bown = ties.pairs[2 * (-2 - code) + 1];
        } else {
            p.own = p.bown = code;
        }
        if (p.bown >= 0 && (p.bown & SMALL_FLAG)) p.bown = -1;  // taken by k_small_tri_bwd (triangle-parallel)
        if (p.bown >= 0)
            for (int k = 0; k < s.nb_colors; k++) g[k] = image_b[idx * s.nb_colors + k];
    }
    interior_adjoint_warp<MAXC>(s, x, y, inside && p.bown >= 0, p, g, grads);
}

// Triangle-parallel interior adjoint of the small triangles (one thread per entry of the compacted small list).
template <int MAXC, bool TEX>
#ifndef DEODR_SMALL_MIN_CTAS
#define DEODR_SMALL_MIN_CTAS 8
#endif
__global__ void __launch_bounds__(128, DEODR_SMALL_MIN_CTAS) k_small_tri_bwd(SceneView s, int tiles_x, const int *small_ids, int num_small,
                                                       const int *edge_count, TieTable ties, const int *owner,
                                                       const float *image_b, DeodrGrads grads) {
    fix_channel_count<MAXC, TEX>(s);
    s.perspective_correct = 0;  // the adjoint is only defined without it (validate_view rejects it): folds the branches
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_small) return;
    small_triangle_adjoint<MAXC, DevEnv>(s, small_ids[i], tiles_x, edge_count, owner, ties.pairs, image_b, grads.ij_b,
                                         grads.colors_b, grads.uv_b, grads.shade_b, grads.texture_b);
}

__global__ void k_finalize_edges(SceneView s, const int *edge_sorted, const int *num_edges, double sigma,
                                 const double *edge_acc, DeodrGrads grads) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= *num_edges) return;
    finalize_edge<DevEnv>(s, edge_sorted[r], sigma, edge_acc + (size_t)r * edge_acc_stride(s.nb_colors), grads.ij_b,
                          grads.colors_b, grads.uv_b, grads.shade_b);
}

__global__ void k_check_scene(SceneView s, int *bad) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * s.nb_triangles) return;
    if (s.faces[i] >= (uint32_t)s.nb_vertices) atomicOr(bad, 1);
    if (s.faces_uv[i] >= (uint32_t)s.nb_uv) atomicOr(bad, 2);
}

// ------------------------------------------------------------------------------------------------- host side

char *deodr_error_buffer() {
    static thread_local char buffer[512] = "";
    return buffer;
}

static int sm_count_cached = 0;

// RAII bracket: records a (start, stop) event pair around a group of launches when timing is enabled
struct PhaseTimer {
    DeodrWorkspace *ws;
    cudaStream_t st;
    int slot;
    PhaseTimer(DeodrWorkspace *w, int phase, cudaStream_t s) : ws(w), st(s), slot(-1) {
        if (ws->ev_used < (int)ws->ev_start.size()) {
            slot = ws->ev_used++;
            ws->ev_phase[slot] = phase;
            cudaEventRecord(ws->ev_start[slot], st);
        }
    }
    ~PhaseTimer() {
        if (slot >= 0) cudaEventRecord(ws->ev_stop[slot], st);
    }
};

static inline int grid_for(size_t n, int block) { return (int)((n + block - 1) / block); }

// Fork: auxiliary stream i continues from the current point of the caller's stream; join: the caller's stream waits
// for it.  Independent kernel chains then overlap (tails of one fill with CTAs of the other; no launch gaps).
static inline cudaStream_t fork_stream(DeodrWorkspace *ws, int i, cudaStream_t st, bool *first) {
    if (!ws->overlap) return st;
    if (*first) { cudaEventRecord(ws->ev_fork, st); *first = false; }
    cudaStreamWaitEvent(ws->aux[i], ws->ev_fork, 0);
    return ws->aux[i];
}
static inline void join_stream(DeodrWorkspace *ws, int i, cudaStream_t st) {
    if (!ws->overlap) return;
    cudaEventRecord(ws->ev_join[i], ws->aux[i]);
    cudaStreamWaitEvent(st, ws->ev_join[i], 0);
}

template <int MAXC>
static void launch_fwd(DeodrWorkspace *ws, const SceneView &s, double sigma, const int *edge_count, TieTable ties,
                       float *image, double *z, int *owner, int *face_id, bool edge_chain, cudaStream_t st) {
    const bool tex = ws->any_textured != 0;
    {
        PhaseTimer timer(ws, DEODR_B200_PH_TILE_Z, st);
        // DEODR_B200_TILEZ_CTAS_PER_SM = k > 0 runs k persistent CTAs per SM with the two-stage TMA pipeline; default 0 =
        // one CTA per tile (the same kernel, its loop runs once).  Measured on B200, 1M-triangle scene: 168 us per-tile
        // vs 199 us persistent x4 vs 300 us persistent x2 - the hardware CTA scheduler balances the very uneven tiles
        // better than a static stride, and 4 resident CTAs already overlap each other's copy latency.
        static const int per_sm = getenv("DEODR_B200_TILEZ_CTAS_PER_SM") ? atoi(getenv("DEODR_B200_TILEZ_CTAS_PER_SM")) : 0;
        // DEODR_B200_TILEZ_TILES_PER_CTA = k: every CTA walks k tiles (stride = grid size) with the copy of the next
        // tile in flight while it tests the current one
        static const int per_cta = getenv("DEODR_B200_TILEZ_TILES_PER_CTA") ? atoi(getenv("DEODR_B200_TILEZ_TILES_PER_CTA")) : 1;
        int persistent = per_sm > 0 ? per_sm * (sm_count_cached > 0 ? sm_count_cached : 148) : ws->num_tiles;
        if (per_sm <= 0 && per_cta > 1) persistent = (ws->num_tiles + per_cta - 1) / per_cta;
        (s.perspective_correct ? k_tile_z<true> : k_tile_z<false>)<<<ws->num_tiles < persistent ? ws->num_tiles : persistent, NT, 0, st>>>(
            s, make_tile_div(ws->tiles_x), ws->num_tiles, ws->bins, ties, z, owner, face_id);
    }
    {
        PhaseTimer timer(ws, DEODR_B200_PH_SHADE, st);
        (s.perspective_correct ? (tex ? k_shade<MAXC, true, true> : k_shade<MAXC, true, false>)
                               : (tex ? k_shade<MAXC, false, true> : k_shade<MAXC, false, false>))<<<ws->num_tiles, NT, 0, st>>>(s, make_tile_div(ws->tiles_x), ties, owner, z, image);
    }
    ws->launches += 2;
    if (edge_chain) join_stream(ws, 0, st);  // the edge lists are ready
    if (edge_count && ws->num_edge_tiles > 0) {
        PhaseTimer timer(ws, DEODR_B200_PH_EDGE_FWD, st);
        (s.perspective_correct ? (tex ? k_edge_fwd<MAXC, true, true> : k_edge_fwd<MAXC, true, false>)
                               : (tex ? k_edge_fwd<MAXC, false, true> : k_edge_fwd<MAXC, false, false>))<<<ws->num_edge_tiles * (TS / EDGE_ROWS), EDGE_NT, 0, st>>>(s, sigma, make_tile_div(ws->tiles_x), ws->edge_tiles_ptr, ws->num_tiles, ws->num_heavy_edge_tiles, edge_count,
                                                            ws->edge_offset.as<int>(), ws->edge_refs.as<int>(),
                                                            ws->edge_recs.as<EdgeRec>(), ws->edge_spans.as<uint32_t>(), z, image);
        ws->launches++;
    }
}

template <int MAXC>
static void launch_bwd(DeodrWorkspace *ws, const SceneView &s, double sigma, const int *edge_count, TieTable ties,
                       const double *z, const int *owner, const float *image_b, const DeodrGrads &g, int *scal,
                       cudaStream_t st) {
    // Three independent chains (disjoint pixel sets, all accumulate with atomics): edge tiles (longest tail: launched
    // first so that its CTAs are dispatched first), tiles with large triangles, small triangles.
    bool first = true;
    const bool tex = ws->any_textured != 0;
    const int E = ws->num_edges, C = s.nb_colors;
    const bool edges = edge_count && ws->num_edge_tiles > 0 && E > 0;
    if (edges) {
        cudaStream_t se = fork_stream(ws, 0, st, &first);
        cudaMemsetAsync(ws->edge_acc.ptr, 0, (size_t)E * edge_acc_stride(C) * sizeof(double), se);
        {
            PhaseTimer timer(ws, DEODR_B200_PH_EDGE_BWD, se);
            (tex ? k_raster_bwd<MAXC, true> : k_raster_bwd<MAXC, false>)<<<ws->num_edge_tiles * (TS / EDGE_ROWS), EDGE_NT, 0, se>>>(
                s, sigma, make_tile_div(ws->tiles_x), ws->edge_tiles_ptr, ws->num_tiles, ws->num_heavy_edge_tiles, edge_count,
                ws->edge_offset.as<int>(),
                ws->edge_refs.as<int>(), ws->edge_recs.as<EdgeRec>(), ws->edge_spans.as<uint32_t>(), ties, z, owner,
                image_b, g,
                ws->edge_acc.as<double>());
        }
        {
            PhaseTimer timer(ws, DEODR_B200_PH_EDGE_FINALIZE, se);
            k_finalize_edges<<<grid_for(E, 128), 128, 0, se>>>(s, ws->edge_sorted.as<int>(), scal + 1, sigma,
                                                               ws->edge_acc.as<double>(), g);
        }
        ws->launches += 2;
    }
    if (ws->num_large_tiles > 0) {  // pixels owned by large triangles, tiles without silhouette edges
        cudaStream_t sl = fork_stream(ws, 1, st, &first);
        PhaseTimer timer(ws, DEODR_B200_PH_INTERIOR_BWD, sl);
        (tex ? k_interior_bwd<MAXC, true> : k_interior_bwd<MAXC, false>)<<<ws->num_large_tiles, NT, 0, sl>>>(s, make_tile_div(ws->tiles_x), ws->large_tiles.as<int>(), edge_count,
                                                                 ties, owner, image_b, g);
        ws->launches++;
    }
    if (ws->num_small > 0) {
        PhaseTimer timer(ws, DEODR_B200_PH_SMALL_BWD, st);
        (tex ? k_small_tri_bwd<MAXC, true> : k_small_tri_bwd<MAXC, false>)<<<grid_for(ws->num_small, 128), 128, 0, st>>>(s, ws->tiles_x, ws->small_ids.as<int>(),
                                                                           ws->num_small, edge_count, ties, owner,
                                                                           image_b, g);
        ws->launches++;
    }
    if (edges) join_stream(ws, 0, st);
    if (ws->num_large_tiles > 0) join_stream(ws, 1, st);
}

static int validate_view(const DeodrSceneView *v, bool backward) {
    if (!v) return set_error(DEODR_B200_EINVAL, "scene == NULL");
    // zero-sized arrays may legitimately come with a null data pointer
    const bool tris = v->nb_triangles > 0, verts = v->nb_vertices > 0;
    if (tris && !v->faces) return set_error(DEODR_B200_EINVAL, "scene.faces == NULL");
    if (tris && !v->faces_uv) return set_error(DEODR_B200_EINVAL, "scene.faces_uv == NULL");
    if (verts && !v->depths) return set_error(DEODR_B200_EINVAL, "scene.depths == NULL");
    if (v->nb_uv > 0 && !v->uv) return set_error(DEODR_B200_EINVAL, "scene.uv == NULL");
    if (verts && !v->ij) return set_error(DEODR_B200_EINVAL, "scene.ij == NULL");
    if (verts && !v->shade) return set_error(DEODR_B200_EINVAL, "scene.shade == NULL");
    if (verts && !v->colors) return set_error(DEODR_B200_EINVAL, "scene.colors == NULL");
    if (tris && !v->edgeflags) return set_error(DEODR_B200_EINVAL, "scene.edgeflags == NULL");
    if (tris && !v->textured) return set_error(DEODR_B200_EINVAL, "scene.textured == NULL");
    if (tris && !v->shaded) return set_error(DEODR_B200_EINVAL, "scene.shaded == NULL");
    if (v->texture_height > 0 && v->texture_width > 0 && !v->texture)
        return set_error(DEODR_B200_EINVAL, "scene.texture == NULL");
    if (!v->background_image && !v->background_color)
        return set_error(DEODR_B200_EINVAL, "scene.background == NULL and scene.background_color == NULL");
    if (v->height <= 0 || v->width <= 0 || v->height > 32767 || v->width > 32767)
        return set_error(DEODR_B200_EINVAL, "image sides must be in [1, 32767] (short loop counters, DR.h:925)");
    if (v->nb_colors < 1 || v->nb_colors > 16)
        return set_error(DEODR_B200_EUNSUPPORTED, "nb_colors must be in [1, 16]");
    if (v->nb_triangles < 0 || v->nb_vertices < 0 || v->nb_uv < 0)
        return set_error(DEODR_B200_EINVAL, "negative size");
    if (backward) {
        if (!v->backface_culling)
            return set_error(DEODR_B200_EUNSUPPORTED,
                             "You have to use backface_culling true if you ant to compute gradients");
        if (v->perspective_correct)
            return set_error(DEODR_B200_EUNSUPPORTED,
                             "backward gradient propagation not supported yet with perspective_correct=True");
    }
    return DEODR_B200_OK;
}

extern "C" {

const char *deodr_b200_last_error(void) { return deodr_error_buffer(); }
const char *deodr_b200_version(void) { return "deodr_b200 0.1 (sm_100a)"; }

const char *deodr_b200_phase_name(int phase) {
    static const char *names[] = {"bin_count", "edge_order",    "bin_fill",     "edge_tile_sort", "tile_z",       "shade",
                                  "edge_fwd",  "small_tri_bwd", "interior_bwd", "edge_bwd",       "edge_finalize"};
    return phase >= 0 && phase < DEODR_B200_PH_COUNT ? names[phase] : "?";
}

int deodr_b200_timing_enable(DeodrWorkspace *ws, int max_records) {
    if (!ws) return set_error(DEODR_B200_EINVAL, "ws == NULL");
    CUDA_TRY(cudaSetDevice(ws->device));
    for (cudaEvent_t e : ws->ev_start) cudaEventDestroy(e);
    for (cudaEvent_t e : ws->ev_stop) cudaEventDestroy(e);
    ws->ev_start.clear();
    ws->ev_stop.clear();
    ws->ev_phase.clear();
    ws->ev_used = 0;
    for (int i = 0; i < max_records; i++) {
        cudaEvent_t a, b;
        CUDA_TRY(cudaEventCreate(&a));
        CUDA_TRY(cudaEventCreate(&b));
        ws->ev_start.push_back(a);
        ws->ev_stop.push_back(b);
        ws->ev_phase.push_back(0);
    }
    return DEODR_B200_OK;
}

int deodr_b200_timing_collect(DeodrWorkspace *ws, int32_t *phase, float *ms, int capacity) {
    if (!ws || !phase || !ms) return -1;
    int n = ws->ev_used < capacity ? ws->ev_used : capacity;
    for (int i = 0; i < n; i++) {
        if (cudaEventSynchronize(ws->ev_stop[i]) != cudaSuccess) return -1;
        float t = 0;
        if (cudaEventElapsedTime(&t, ws->ev_start[i], ws->ev_stop[i]) != cudaSuccess) return -1;
        phase[i] = ws->ev_phase[i];
        ms[i] = t;
    }
    if (getenv("DEODR_B200_TRACE_GAPS")) {  // development aid: idle time between consecutive timed phases
        double gap[DEODR_B200_PH_COUNT] = {0};
        int cnt[DEODR_B200_PH_COUNT] = {0};
        for (int i = 1; i < n; i++) {
            float t = 0;
            if (cudaEventElapsedTime(&t, ws->ev_stop[i - 1], ws->ev_start[i]) != cudaSuccess) continue;
            gap[ws->ev_phase[i]] += t;
            cnt[ws->ev_phase[i]]++;
        }
        for (int k = 0; k < DEODR_B200_PH_COUNT; k++)
            if (cnt[k]) fprintf(stderr, "[deodr_b200] gap before %-14s %.1f us\n", deodr_b200_phase_name(k), 1e3 * gap[k] / cnt[k]);
    }
    ws->ev_used = 0;
    return n;
}

int deodr_b200_workspace_create(DeodrWorkspace **out, int device) {
    if (!out) return set_error(DEODR_B200_EINVAL, "ws == NULL");
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0)
        return set_error(DEODR_B200_ECUDA, "no CUDA device available (%s): deodr_b200 has no CPU fallback",
                         cudaGetErrorString(err));
    if (device < 0 || device >= count) return set_error(DEODR_B200_EINVAL, "bad device index");
    CUDA_TRY(cudaSetDevice(device));
    DeodrWorkspace *ws = new (std::nothrow) DeodrWorkspace();
    if (!ws) return set_error(DEODR_B200_ENOMEM, "out of host memory");
    ws->device = device;
    CUDA_TRY(cudaMallocHost(&ws->host_totals, 32 * sizeof(int)));
    memset(ws->host_totals, 0, 32 * sizeof(int));
    CUDA_TRY(cudaFuncSetAttribute(k_scan_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, SCAN_SMEM));
    ws->overlap = !(getenv("DEODR_B200_SERIAL") && atoi(getenv("DEODR_B200_SERIAL")));
    for (int i = 0; i < 2; i++) {
        CUDA_TRY(cudaStreamCreateWithFlags(&ws->aux[i], cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&ws->ev_join[i], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventCreateWithFlags(&ws->ev_fork, cudaEventDisableTiming));
    if (ws->scalars.ensure(8 * sizeof(int), &ws->bytes)) return DEODR_B200_ECUDA;
    if (!sm_count_cached) cudaDeviceGetAttribute(&sm_count_cached, cudaDevAttrMultiProcessorCount, device);
    *out = ws;
    return DEODR_B200_OK;
}

void deodr_b200_workspace_destroy(DeodrWorkspace *ws) {
    if (!ws) return;
    cudaSetDevice(ws->device);
    DevBuf *bufs[] = {&ws->zeroed, &ws->large_tiles, &ws->edge_tiles, &ws->small_offset, &ws->small_recs, &ws->small_ids, &ws->large_ids, &ws->tri_offset, &ws->tri_refs, &ws->edge_ids,
                      &ws->edge_ids_tmp, &ws->edge_rank, &ws->edge_recs,
                      &ws->edge_keys_in, &ws->edge_keys_out, &ws->edge_sorted, &ws->cub_temp,
                      &ws->edge_offset, &ws->edge_refs_tmp, &ws->edge_refs, &ws->edge_spans, &ws->scalars,
                      &ws->tie_pairs, &ws->edge_acc, &ws->h_faces, &ws->h_faces_uv, &ws->h_ij, &ws->h_depths, &ws->h_uv,
                      &ws->h_colors, &ws->h_shade, &ws->h_edgeflags, &ws->h_textured, &ws->h_shaded, &ws->h_texture,
                      &ws->h_background, &ws->h_image, &ws->h_z, &ws->h_owner, &ws->h_image_b,
                      &ws->h_grads};
    for (DevBuf *b : bufs)
        if (b->ptr) cudaFree(b->ptr);
    deodr_host_path_destroy(ws);
    if (ws->host_totals) cudaFreeHost(ws->host_totals);
    for (cudaEvent_t e : ws->ev_start) cudaEventDestroy(e);
    for (cudaEvent_t e : ws->ev_stop) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++) {
        if (ws->aux[i]) cudaStreamDestroy(ws->aux[i]);
        if (ws->ev_join[i]) cudaEventDestroy(ws->ev_join[i]);
    }
    if (ws->ev_fork) cudaEventDestroy(ws->ev_fork);
    delete ws;
}

int64_t deodr_b200_workspace_bytes(const DeodrWorkspace *ws) { return ws ? ws->bytes : 0; }
int64_t deodr_b200_workspace_launches(const DeodrWorkspace *ws) { return ws ? ws->launches : 0; }

int deodr_b200_check_scene(DeodrWorkspace *ws, const DeodrSceneView *scene, void *stream) {
    if (!ws) return set_error(DEODR_B200_EINVAL, "ws == NULL");
    if (int rc = validate_view(scene, false)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(ws->device));
    SceneView s;
    memcpy(&s, scene, sizeof(s));
    int *flag = ws->scalars.as<int>() + 4;  // private scratch (not the forward pass's zeroed block)
    CUDA_TRY(cudaMemsetAsync(flag, 0, sizeof(int), st));
    if (s.nb_triangles > 0) {
        k_check_scene<<<grid_for(3 * (size_t)s.nb_triangles, 256), 256, 0, st>>>(s, flag);
        ws->launches++;
    }
    CUDA_TRY(cudaMemcpyAsync(ws->host_totals + 4, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (ws->host_totals[4] & 1) return set_error(DEODR_B200_EINVAL, "scene.faces value greater than scene.nb_vertices");
    if (ws->host_totals[4] & 2) return set_error(DEODR_B200_EINVAL, "scene.faces_uv value greater than scene.nb_uv");
    return DEODR_B200_OK;
}

int deodr_b200_render(DeodrWorkspace *ws, const DeodrSceneView *scene, double sigma, float *image, double *z_buffer,
                      int32_t *owner, int32_t *face_id, void *stream) {
    return deodr_render_impl(ws, scene, sigma, image, z_buffer, owner, face_id, stream, false);
}

}  // extern "C"

int deodr_render_impl(DeodrWorkspace *ws, const DeodrSceneView *scene, double sigma, float *image, double *z_buffer,
                      int32_t *owner, int32_t *face_id, void *stream, bool check_indices) {
    if (!ws) return set_error(DEODR_B200_EINVAL, "ws == NULL");
    if (int rc = validate_view(scene, false)) return rc;
    if (!image) return set_error(DEODR_B200_EINVAL, "image_ptr is NULL");
    if (!z_buffer) return set_error(DEODR_B200_EINVAL, "z_buffer_ptr is NULL");
    if (!owner) return set_error(DEODR_B200_EINVAL, "owner is NULL");
    if (!(sigma >= 0)) return set_error(DEODR_B200_EINVAL, "sigma must be >= 0");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(ws->device));
    SceneView s;
    memcpy(&s, scene, sizeof(s));
    ws->fwd_valid = 0;
    const int T = s.nb_triangles;
    ws->tiles_x = (s.width + TS - 1) / TS;
    ws->tiles_y = (s.height + TS - 1) / TS;
    ws->num_tiles = ws->tiles_x * ws->tiles_y;
    const int nt = ws->num_tiles;
    // one zero-initialised block: [scalars(16) | small_count | small_cursor | large_count | large_cursor | edge_count |
    // edge_cursor], (nt+1) ints each
    const size_t tile_ints = (size_t)nt + 1;
    const size_t tile_bytes = tile_ints * sizeof(int);
    int rc = 0;
    rc |= ws->zeroed.ensure((16 + 6 * tile_ints) * sizeof(int), &ws->bytes);
    rc |= ws->small_offset.ensure(tile_bytes, &ws->bytes);
    rc |= ws->large_tiles.ensure(tile_bytes, &ws->bytes);
    rc |= ws->edge_tiles.ensure(tile_bytes, &ws->bytes);
    ws->edge_tiles_ptr = ws->edge_tiles.as<int>();
    rc |= ws->tri_offset.ensure(tile_bytes, &ws->bytes);
    rc |= ws->edge_offset.ensure(tile_bytes, &ws->bytes);
    rc |= ws->edge_ids.ensure(((size_t)3 * T + 4) * sizeof(int), &ws->bytes);
    rc |= ws->edge_keys_in.ensure(((size_t)3 * T + 4) * 8, &ws->bytes);
    rc |= ws->edge_rank.ensure(((size_t)3 * T + 4) * sizeof(int), &ws->bytes);
    rc |= ws->small_ids.ensure(((size_t)T + 4) * sizeof(int), &ws->bytes);
    rc |= ws->large_ids.ensure(((size_t)T + 4) * sizeof(int), &ws->bytes);
    // one (own, bown) slot per pixel: the exact-tie table can never overflow, so the adjoint needs no read-back
    if (ws->tie_capacity < s.height * s.width) {
        ws->tie_capacity = s.height * s.width;
        rc |= ws->tie_pairs.ensure((size_t)2 * ws->tie_capacity * sizeof(int), &ws->bytes);
    }
    if (rc) return DEODR_B200_ECUDA;
    int *scal = ws->zeroed.as<int>();
    int *small_count = scal + 16, *small_cursor = small_count + tile_ints, *large_count = small_cursor + tile_ints,
        *large_cursor = large_count + tile_ints, *edge_count_buf = large_cursor + tile_ints,
        *edge_cursor = edge_count_buf + tile_ints;
    ws->scal = scal;
    ws->edge_count_ptr = edge_count_buf;
    EdgeList edges{scal + 1, ws->edge_ids.as<int>(), (uint64_t *)ws->edge_keys_in.ptr, ws->edge_rank.as<int>()};
    TriBins bins{small_count, ws->small_offset.as<int>(), small_cursor, nullptr,
                 large_count, ws->tri_offset.as<int>(), large_cursor, nullptr};
    TriLists lists{scal + 6, ws->small_ids.as<int>(), scal + 7, ws->large_ids.as<int>()};

    // ---- count pass + scans (triangles and silhouette edges together), then the ONE host read-back of the sizes
    {
        PhaseTimer timer(ws, DEODR_B200_PH_BIN_COUNT, st);
        CUDA_TRY(cudaMemsetAsync(scal, 0, (16 + 6 * tile_ints) * sizeof(int), st));
        if (T > 0) {
            if (check_indices) {  // checkSceneValid (DR.h:2703-2714) on the device, before any index is dereferenced
                k_check_scene<<<grid_for(3 * (size_t)T, 256), 256, 0, st>>>(s, scal + 4);
                ws->launches++;
            }
            k_bin_count<<<grid_for(T, 128), 128, 0, st>>>(s, sigma, ws->tiles_x, bins, lists, edges, edge_count_buf,
                                                          check_indices ? scal + 4 : nullptr, scal + 12);
            ws->launches++;
        }
        ScanJob job{{small_count, large_count, edge_count_buf},
                    {ws->small_offset.as<int>(), ws->tri_offset.as<int>(), ws->edge_offset.as<int>()},
                    {5, 0, 2},
                    {nullptr, ws->large_tiles.as<int>(), ws->edge_tiles.as<int>()},
                    {15, 8, 9},
                    {0, 0, EDGE_CHUNK},
                    {15, 15, 10}};
        ws->totals_seq = ws->totals_seq == 0x7fffffff ? 1 : ws->totals_seq + 1;
        k_scan_tiles<<<3, 1024, SCAN_SMEM, st>>>(job, nt, scal, ws->host_totals, ws->totals_seq);
        ws->launches++;
    }
    {
        // the one host read-back of the forward: poll the flag the scan kernel raises in pinned memory (a few
        // microseconds after the kernel's last store) instead of a copy + stream synchronisation
        volatile int *flag = ws->host_totals + 16;
        for (unsigned spins = 1; *flag != ws->totals_seq; spins++) {
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
            if ((spins & 0xfffu) == 0) {  // every few microseconds: has the stream drained (flag lost) or failed?
                const cudaError_t q = cudaStreamQuery(st);
                if (q == cudaSuccess) {
                    if (*flag != ws->totals_seq) {  // not expected; fall back to an explicit copy
                        CUDA_TRY(cudaMemcpyAsync(ws->host_totals, scal, 16 * sizeof(int), cudaMemcpyDeviceToHost, st));
                        CUDA_TRY(cudaStreamSynchronize(st));
                    }
                    break;
                }
                if (q != cudaErrorNotReady) return set_error(DEODR_B200_ECUDA, "forward pass failed: %s", cudaGetErrorString(q));
            }
        }
        __atomic_thread_fence(__ATOMIC_ACQUIRE);
    }
    if (check_indices && (ws->host_totals[4] & 1))
        return set_error(DEODR_B200_EINVAL, "scene.faces value greater than scene.nb_vertices");
    if (check_indices && (ws->host_totals[4] & 2))
        return set_error(DEODR_B200_EINVAL, "scene.faces_uv value greater than scene.nb_uv");
    const int large_total = ws->host_totals[0], E = ws->host_totals[1], edge_total = ws->host_totals[2],
              small_total = ws->host_totals[5];
    ws->num_edges = E;
    ws->num_small = ws->host_totals[6];
    ws->num_large = ws->host_totals[7];
    ws->num_large_tiles = ws->host_totals[8];
    ws->num_edge_tiles = E > 0 ? ws->host_totals[9] : 0;
    ws->num_heavy_edge_tiles = E > 0 ? ws->host_totals[10] : 0;
    ws->any_textured = ws->host_totals[12] != 0;
    rc = 0;
    rc |= ws->small_recs.ensure(((size_t)small_total + 1) * sizeof(PreRec), &ws->bytes);
    rc |= ws->tri_refs.ensure(((size_t)large_total + 4) * sizeof(int), &ws->bytes);
    rc |= ws->edge_sorted.ensure(((size_t)E + 4) * sizeof(int), &ws->bytes);
    rc |= ws->edge_recs.ensure(((size_t)E + 1) * sizeof(EdgeRec), &ws->bytes);
    rc |= ws->edge_refs_tmp.ensure(((size_t)edge_total + 4) * sizeof(int), &ws->bytes);
    rc |= ws->edge_refs.ensure(((size_t)edge_total + 4) * sizeof(int), &ws->bytes);
    rc |= ws->edge_spans.ensure(((size_t)edge_total + 4) * TS * sizeof(uint32_t), &ws->bytes);
    if (rc) return DEODR_B200_ECUDA;
    bins.small_recs = ws->small_recs.as<PreRec>();
    bins.large_refs = ws->tri_refs.as<int>();
    ws->bins = bins;

    // ---- far-to-near order of the silhouette edges (DR.h:2781) + per-tile edge lists: a chain of its own (stream se)
    // that overlaps the triangle fill, the z pass and the shading; joined before k_edge_fwd
    bool first_fork = true;
    cudaStream_t se = E > 0 ? fork_stream(ws, 0, st, &first_fork) : st;
    if (E > 0) {
        PhaseTimer timer(ws, DEODR_B200_PH_EDGE_ORDER, se);
        if (E <= 65536) {
            dim3 grid(grid_for(E, 256), grid_for(E, RANK_CHUNK));
            k_rank_edges<<<grid, 256, 0, se>>>(edges, E, ws->edge_rank.as<int>());
            k_scatter_edges<<<grid_for(E, 128), 128, 0, se>>>(s, edges, E, sigma, ws->edge_sorted.as<int>(),
                                                               ws->edge_recs.as<EdgeRec>());
            ws->launches += 2;
        } else {
            // large soups: two stable radix sorts (by id, then by depth key) give the same total order
            rc = 0;
            rc |= ws->edge_keys_out.ensure((size_t)E * 8, &ws->bytes);
            rc |= ws->edge_ids_tmp.ensure((size_t)E * sizeof(int), &ws->bytes);
            if (rc) return DEODR_B200_ECUDA;
            auto *k_in = ws->edge_keys_in.as<unsigned long long>(), *k_out = ws->edge_keys_out.as<unsigned long long>();
            int *i_in = ws->edge_ids.as<int>(), *i_tmp = ws->edge_ids_tmp.as<int>();
            size_t temp1 = 0, temp2 = 0;
            CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, temp1, i_in, i_tmp, k_in, k_out, E, 0, 32, se));
            CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, temp2, k_out, k_in, i_tmp, i_in, E, 0, 64, se));
            if (ws->cub_temp.ensure(temp1 > temp2 ? temp1 : temp2, &ws->bytes)) return DEODR_B200_ECUDA;
            CUDA_TRY(cub::DeviceRadixSort::SortPairs(ws->cub_temp.ptr, temp1, i_in, i_tmp, k_in, k_out, E, 0, 32, se));
            CUDA_TRY(cub::DeviceRadixSort::SortPairs(ws->cub_temp.ptr, temp2, k_out, k_in, i_tmp,
                                                     ws->edge_sorted.as<int>(), E, 0, 64, se));
            k_edge_records<<<grid_for(E, 128), 128, 0, se>>>(s, ws->edge_sorted.as<int>(), E, sigma,
                                                             ws->edge_recs.as<EdgeRec>());
            ws->launches++;
        }
    }

    // ---- per-tile edge lists: fill, then order every list by far-to-near rank
    if (E > 0) {
        PhaseTimer timer(ws, DEODR_B200_PH_EDGE_TILE_SORT, se);
        k_bin_fill<<<grid_for(E, 128), 128, 0, se>>>(s, sigma, ws->tiles_x, 0, 0, bins, nullptr, 0, nullptr, 0,
                                                     ws->edge_sorted.as<int>(), E, ws->edge_offset.as<int>(),
                                                     edge_cursor, ws->edge_refs_tmp.as<int>());
        ws->launches++;
        if (ws->num_edge_tiles > 0) {
            k_sort_tile_edges<<<ws->num_edge_tiles, 128, 0, se>>>(ws->edge_tiles.as<int>(), nt, ws->num_heavy_edge_tiles,
                                                                  edge_count_buf,
                                                                  ws->edge_offset.as<int>(),
                                                                  ws->edge_refs_tmp.as<int>(), ws->edge_refs.as<int>());
            ws->launches++;
        }
    }
    // ---- triangle fill: pre-masked records of the small triangles, index lists of the large ones
    if (T > 0) {
        PhaseTimer timer(ws, DEODR_B200_PH_BIN_FILL, st);
        const int small_blocks = grid_for(ws->num_small, 128), large_blocks = grid_for(ws->num_large, 128);
        if (small_blocks + large_blocks > 0) {
            k_bin_fill<<<small_blocks + large_blocks, 128, 0, st>>>(
                s, sigma, ws->tiles_x, small_blocks, large_blocks, bins, ws->small_ids.as<int>(), ws->num_small,
                ws->large_ids.as<int>(), ws->num_large, nullptr, 0, nullptr, nullptr, nullptr);
            ws->launches++;
        }
    }

    // ---- raster
    TieTable ties{ws->tie_pairs.as<int>(), scal + 3, ws->tie_capacity};
    const int *edge_count = E > 0 ? edge_count_buf : nullptr;
    const int C = s.nb_colors;
    {
    if (C == 1) launch_fwd<1>(ws, s, sigma, edge_count, ties, image, z_buffer, owner, face_id, E > 0, st);
    else if (C == 3) launch_fwd<3>(ws, s, sigma, edge_count, ties, image, z_buffer, owner, face_id, E > 0, st);
    else if (C <= 4) launch_fwd<4>(ws, s, sigma, edge_count, ties, image, z_buffer, owner, face_id, E > 0, st);
    else launch_fwd<16>(ws, s, sigma, edge_count, ties, image, z_buffer, owner, face_id, E > 0, st);
    }
    CUDA_TRY(cudaGetLastError());
    ws->sigma = sigma;
    ws->fwd_T = T; ws->fwd_H = s.height; ws->fwd_W = s.width; ws->fwd_C = C;
    ws->fwd_valid = 1;
    return DEODR_B200_OK;
}

extern "C" {

int deodr_b200_render_b(DeodrWorkspace *ws, const DeodrSceneView *scene, double sigma, const double *z_buffer,
                        const int32_t *owner, const float *image_b, const DeodrGrads *grads, void *stream) {
    if (!ws) return set_error(DEODR_B200_EINVAL, "ws == NULL");
    if (int rc = validate_view(scene, true)) return rc;
    if (!z_buffer || !owner) return set_error(DEODR_B200_EINVAL, "z_buffer / owner is NULL");
    if (!image_b) return set_error(DEODR_B200_EINVAL, "image_b_ptr is NULL");
    if (!grads) return set_error(DEODR_B200_EINVAL, "grads == NULL");
    SceneView s;
    memcpy(&s, scene, sizeof(s));
    if (!ws->fwd_valid || ws->sigma != sigma || ws->fwd_T != s.nb_triangles || ws->fwd_H != s.height ||
        ws->fwd_W != s.width || ws->fwd_C != s.nb_colors)
        return set_error(DEODR_B200_EINVAL, "render_b must follow render on the same workspace, scene and sigma");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(ws->device));
    int *scal = ws->scal;
    const int E = ws->num_edges, C = s.nb_colors;
    DeodrGrads g = *grads;
    if ((s.nb_vertices > 0 && (!g.ij_b || !g.colors_b || !g.shade_b)) || (s.nb_uv > 0 && !g.uv_b))
        return set_error(DEODR_B200_EINVAL, "ij_b / colors_b / uv_b / shade_b must be provided");
    if (E > 0) {
        size_t acc_bytes = (size_t)E * edge_acc_stride(C) * sizeof(double);
        if (ws->edge_acc.ensure(acc_bytes, &ws->bytes)) return DEODR_B200_ECUDA;
    }
    TieTable ties{ws->tie_pairs.as<int>(), scal + 3, ws->tie_capacity};
    const int *edge_count = E > 0 ? ws->edge_count_ptr : nullptr;
    if (C == 1) launch_bwd<1>(ws, s, sigma, edge_count, ties, z_buffer, owner, image_b, g, scal, st);
    else if (C == 3) launch_bwd<3>(ws, s, sigma, edge_count, ties, z_buffer, owner, image_b, g, scal, st);
    else if (C <= 4) launch_bwd<4>(ws, s, sigma, edge_count, ties, z_buffer, owner, image_b, g, scal, st);
    else launch_bwd<16>(ws, s, sigma, edge_count, ties, z_buffer, owner, image_b, g, scal, st);
    CUDA_TRY(cudaGetLastError());
    return DEODR_B200_OK;
}

}  // extern "C"
