// Copy-thread crew of the host path (deodr_b200_render_host / render_b_host / host_zero): pure C++, no CUDA, so that
// tests/test_host_crew.py can stress it on the CPU.  Included by host_api.cu only.
#pragma once

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__linux__)
#include <pthread.h>
#include <sched.h>
#endif

#include "host_simd.h"

// workers (besides the coordinator) of a PCIe-bound batch; the others only join the wide, DRAM-bound batches
constexpr int WIDTH_PCIE_BOUND = 7;

// ---------------------------------------------------------------------------------------------- copy thread crew
//
// A call moves a few hundred MB between the caller's arrays and pinned memory, as dozens of chunks that are each tied
// to a DMA (an upload chunk is sent when its conversion is complete, a download chunk may be converted when its DMA has
// landed).  One BATCH describes all of it: chunks in DMA order, cut into ~256 KB tasks that the workers claim with an
// atomic counter.  The calling thread is the coordinator: it opens the gates of download chunks as their events
// complete and enqueues the DMA of upload chunks as their last task retires.  Workers spin for a short while between
// batches (a call issues several batches back to back) before they go to sleep on a condition variable, so that a
// batch starts within microseconds instead of a futex wake-up per thread.

enum OpKind { OP_COPY, OP_F64_TO_F32, OP_F32_TO_F64, OP_F32_ADD_F64, OP_ZERO, OP_EQ_RAW, OP_EQ_F32 };

struct Chunk {
    OpKind kind;
    char *dst;        // OP_EQ_*: the mirror side of the comparison
    const char *src;  // OP_EQ_*: the caller's side
    size_t count;     // elements (bytes for OP_COPY / OP_ZERO / OP_EQ_RAW)
    int n_tasks = 0;
    std::atomic<int> remaining{0};
    Chunk(OpKind k, void *d, const void *s, size_t c) : kind(k), dst((char *)d), src((const char *)s), count(c) {}
};

struct TaskRef {
    int chunk;
    size_t lo, hi;  // element range
};

struct Batch {
    std::deque<Chunk> chunks;
    std::vector<TaskRef> tasks;
    std::atomic<int> next_task{0}, tasks_done{0}, open_chunks{0}, unequal{0};
    // workers allowed on this batch (DEODR_B200_HOST_WIDE=0: development switch that keeps every batch narrow)
    int width = (getenv("DEODR_B200_HOST_WIDE") && atoi(getenv("DEODR_B200_HOST_WIDE")) == 0) ? 7 : 1 << 20;  // (PCIe-bound batches run best with fewer threads than DRAM-bound ones)

    static size_t src_elem(OpKind k) { return k == OP_F64_TO_F32 || k == OP_EQ_F32 ? 8 : k == OP_F32_TO_F64 || k == OP_F32_ADD_F64 ? 4 : 1; }
    static size_t dst_elem(OpKind k) { return k == OP_F64_TO_F32 || k == OP_EQ_F32 ? 4 : k == OP_F32_TO_F64 || k == OP_F32_ADD_F64 ? 8 : 1; }

    // appends one chunk and its tasks (`task_bytes` of the wider side each, whole cache lines)
    int add(OpKind kind, void *dst, const void *src, size_t count, size_t task_bytes = 256 << 10) {
        chunks.emplace_back(kind, dst, src, count);
        Chunk &c = chunks.back();
        const size_t wide = std::max(src_elem(kind), dst_elem(kind));
        const size_t per = std::max<size_t>(64, (task_bytes / wide) & ~(size_t)63);
        const int id = (int)chunks.size() - 1;
        for (size_t lo = 0; lo < count; lo += per) {
            tasks.push_back(TaskRef{id, lo, std::min(count, lo + per)});
            c.n_tasks++;
        }
        c.remaining.store(c.n_tasks, std::memory_order_relaxed);
        return id;
    }
    void open_all() { open_chunks.store((int)chunks.size(), std::memory_order_release); }

    void execute(const TaskRef &t) {
        Chunk &c = chunks[t.chunk];
        const size_t n = t.hi - t.lo;
        switch (c.kind) {
            case OP_COPY: deodr_simd_copy(c.dst + t.lo, c.src + t.lo, n); break;
            case OP_ZERO: deodr_simd_zero(c.dst + t.lo, n); break;
            case OP_F64_TO_F32: deodr_simd_f64_to_f32((float *)c.dst + t.lo, (const double *)c.src + t.lo, n); break;
            case OP_F32_TO_F64: deodr_simd_f32_to_f64((double *)c.dst + t.lo, (const float *)c.src + t.lo, n); break;
            case OP_F32_ADD_F64: deodr_simd_f32_add_f64((double *)c.dst + t.lo, (const float *)c.src + t.lo, n); break;
            case OP_EQ_RAW:
                if (memcmp(c.dst + t.lo, c.src + t.lo, n) != 0) unequal.store(1, std::memory_order_relaxed);
                break;
            case OP_EQ_F32:
                if (!deodr_simd_equal_f32((const double *)c.src + t.lo, (const float *)c.dst + t.lo, n))
                    unequal.store(1, std::memory_order_relaxed);
                break;
        }
        c.remaining.fetch_sub(1, std::memory_order_release);
        tasks_done.fetch_add(1, std::memory_order_release);
    }
    // claims and runs tasks until none is left; a task whose chunk is still gated is waited for (gates open in order)
    void work() {
        const int n = (int)tasks.size();
        for (;;) {
            const int t = next_task.fetch_add(1, std::memory_order_relaxed);
            if (t >= n) return;
            while (open_chunks.load(std::memory_order_acquire) <= tasks[t].chunk) cpu_relax();
            execute(tasks[t]);
        }
    }
    static void cpu_relax() {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
    }
    void wait_chunk(int id) {
        while (chunks[id].remaining.load(std::memory_order_acquire) != 0) cpu_relax();
    }
};

// CPUs of the NUMA node the calling thread runs on (Linux sysfs); empty when it cannot be determined.  The caller's
// arrays and the pinned buffers were first touched from this thread, so the workers are kept on the same node: on a
// two-socket host, workers of the other socket stream through the inter-socket link and slow everybody down.
// `want_node` >= 0 selects that node instead of the caller's (ranks of a multi-GPU job sit next to their GPU).
static std::vector<int> cpus_of_local_node(int want_node = -1) {
    std::vector<int> cpus;
#if defined(__linux__)
    if (getenv("DEODR_B200_HOST_PIN") && atoi(getenv("DEODR_B200_HOST_PIN")) == 0) return cpus;
    const int cpu = sched_getcpu();
    if (cpu < 0) return cpus;
    for (int node = 0; node < 64; node++) {
        char path[96];
        snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
        FILE *f = fopen(path, "r");
        if (!f) break;
        char buf[1024] = "";
        const bool got = fgets(buf, sizeof(buf), f) != nullptr;
        fclose(f);
        if (!got) continue;
        std::vector<int> list;
        for (char *p = buf; *p;) {  // "0-31,64-95"
            char *end;
            long a = strtol(p, &end, 10);
            if (end == p) break;
            long b = a;
            if (*end == '-') b = strtol(end + 1, &end, 10);
            for (long c = a; c <= b && c < CPU_SETSIZE; c++) list.push_back((int)c);
            p = (*end == ',') ? end + 1 : end;
            if (*end != ',' ) break;
        }
        if (want_node >= 0 ? node == want_node : std::find(list.begin(), list.end(), cpu) != list.end()) return list;
    }
#endif
    return cpus;
}

// Where a rank of a multi-GPU job on this host puts its crew: the NUMA node of its GPU, shared with `count` ranks of
// which it is number `index` (host_api.cu works this out from the PCI bus ids).  node < 0: the caller's node, sliced by
// LOCAL_RANK / LOCAL_WORLD_SIZE as a fallback.
struct CrewPlacement {
    int node = -1, index = 0, count = 1;
};

class Crew {
   public:
    explicit Crew(int workers, CrewPlacement place = CrewPlacement()) {
        // one worker per physical core of the caller's node (hyper-thread siblings and the caller's own core are left
        // alone): streaming loops gain nothing from sharing a core
        std::vector<int> cores;
#if defined(__linux__)
        const int self = place.node >= 0 ? -1 : sched_getcpu();
        for (int c : cpus_of_local_node(place.node)) {
            char path[128];
            snprintf(path, sizeof(path), "/sys/devices/system/cpu/cpu%d/topology/thread_siblings_list", c);
            int first = c;
            if (FILE *f = fopen(path, "r")) {
                if (fscanf(f, "%d", &first) != 1) first = c;
                fclose(f);
            }
            bool mine = false;  // is the caller on this core?
            if (first == c) {
                if (FILE *f = fopen(path, "r")) {
                    char buf[128] = "";
                    if (fgets(buf, sizeof(buf), f)) {
                        int a = -1, b = -1;
                        if (sscanf(buf, "%d,%d", &a, &b) >= 1) mine = (a == self || b == self);
                        if (sscanf(buf, "%d-%d", &a, &b) == 2) mine = mine || (self >= a && self <= b);
                    }
                    fclose(f);
                }
                if (!mine) cores.push_back(c);
            }
        }
#endif
        // Several ranks on one host (torchrun: LOCAL_RANK / LOCAL_WORLD_SIZE) must not pin their crews to the same
        // cores: each rank takes its own slice of the node's cores and a crew that fits it.
        int first = 0;
        if (place.node >= 0 && !cores.empty()) {
            // the node's physical cores are split between the ranks whose GPU hangs off it: one core of the slice is
            // left to the rank's own (coordinating) thread, the others run its workers
            const int share = std::max(2, (int)cores.size() / std::max(1, place.count));
            if (!getenv("DEODR_B200_HOST_THREADS")) workers = std::max(1, std::min(workers, share - 1));
            first = (int)(((long)place.index * share + 1) % (long)cores.size());
            caller_core_ = cores[(size_t)(((long)place.index * share) % (long)cores.size())];
            for (int i = 0; i < share; i++) slice_.push_back(cores[(size_t)(((long)place.index * share + i) % (long)cores.size())]);
        } else if (!getenv("DEODR_B200_HOST_THREADS")) {
            const int local_world = getenv("LOCAL_WORLD_SIZE") ? std::max(1, atoi(getenv("LOCAL_WORLD_SIZE"))) : 1;
            const int local_rank = getenv("LOCAL_RANK") ? std::max(0, atoi(getenv("LOCAL_RANK"))) : 0;
            if (local_world > 1 && !cores.empty()) {
                const int share = std::max(3, (int)cores.size() / local_world);
                workers = std::max(1, std::min(workers, share - 1));
                first = (int)(((long)(local_rank % local_world) * share) % (long)cores.size());
            }
        }
        for (int i = 0; i < workers; i++) {
            threads_.emplace_back([this, i] { loop(i); });
#if defined(__linux__)
            if ((int)cores.size() >= workers) {
                cpu_set_t set;
                CPU_ZERO(&set);
                CPU_SET(cores[(first + i) % (int)cores.size()], &set);
                pthread_setaffinity_np(threads_.back().native_handle(), sizeof(set), &set);
            }
#endif
        }
    }
    ~Crew() {
        stop_.store(true);
        publish(nullptr);
        for (auto &t : threads_) t.join();
    }
    int workers() const { return (int)threads_.size(); }
    int caller_core() const { return caller_core_; }  // core set aside for the coordinating thread (-1: none)
    const std::vector<int> &slice() const { return slice_; }  // the rank's share of its node's cores
    // the workers start on `b` at once; the caller coordinates (gates, DMAs) and then calls finish(b)
    void start(Batch *b) { publish(b); }
    // the caller helps with what is left, then waits until every task has retired and no worker still looks at `b`
    void finish(Batch *b) {
        b->work();
        const int n = (int)b->tasks.size();
        while (b->tasks_done.load(std::memory_order_acquire) < n) Batch::cpu_relax();
        current_.store(nullptr, std::memory_order_release);
        while (inside_.load(std::memory_order_acquire) != 0) Batch::cpu_relax();
    }
    void run(Batch *b) {
        b->open_all();
        start(b);
        finish(b);
    }

   private:
    // Workers [0, WIDTH_PCIE_BOUND) follow every batch; the others only the wide (DRAM-bound) ones, through a generation
    // counter of their own, so that a narrow batch neither wakes them nor has them spin next to the busy workers.
    void publish(Batch *b) {
        current_.store(b, std::memory_order_release);
        const bool wide = b == nullptr || b->width > narrow_;
        generation_[0].fetch_add(1, std::memory_order_release);
        if (wide) generation_[1].fetch_add(1, std::memory_order_release);
        if (sleepers_.load(std::memory_order_acquire) > 0 || b == nullptr) {
            std::lock_guard<std::mutex> lock(mu_);
            cv_[0].notify_all();
            if (wide) cv_[1].notify_all();
        }
    }
    void loop(int index) {
        const int group = index < narrow_ ? 0 : 1;
        std::atomic<uint64_t> &generation = generation_[group];
        uint64_t seen = 0;
        for (;;) {
            // wait for a new generation: spin ~200 us (a call issues its batches back to back), then sleep
            const auto t0 = std::chrono::steady_clock::now();
            int spins = 0;
            while (generation.load(std::memory_order_acquire) == seen) {
                Batch::cpu_relax();
                if (++spins % 512 == 0 &&
                    std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(200)) {
                    std::unique_lock<std::mutex> lock(mu_);
                    sleepers_.fetch_add(1, std::memory_order_acq_rel);
                    cv_[group].wait(lock, [&] { return generation.load(std::memory_order_acquire) != seen; });
                    sleepers_.fetch_sub(1, std::memory_order_acq_rel);
                }
            }
            seen = generation.load(std::memory_order_acquire);
            if (stop_.load()) return;
            Batch *b = current_.load(std::memory_order_acquire);
            if (!b) continue;
            inside_.fetch_add(1, std::memory_order_acq_rel);
            if (current_.load(std::memory_order_acquire) == b && index < b->width) b->work();
            inside_.fetch_sub(1, std::memory_order_acq_rel);
        }
    }
    std::vector<std::thread> threads_;
    int caller_core_ = -1;
    std::vector<int> slice_;
    std::mutex mu_;
    std::condition_variable cv_[2];
    std::atomic<Batch *> current_{nullptr};
    std::atomic<uint64_t> generation_[2] = {{0}, {0}};
    const int narrow_ = WIDTH_PCIE_BOUND;
    std::atomic<int> sleepers_{0}, inside_{0};
    std::atomic<bool> stop_{false};
};

