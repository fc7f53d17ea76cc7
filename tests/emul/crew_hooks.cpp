// TEST INFRASTRUCTURE: stress harness for the host path's copy-thread crew (deodr_b200/csrc/host_crew.h) - the same
// batches the host entry points build (uploads whose chunks are consumed in order as they complete, downloads whose
// chunks are gated by a simulated DMA), run back to back with random sizes, widths and pauses, every result checked.
#include <random>

#include "../../deodr_b200/csrc/host_crew.h"

extern "C" long hook_crew_stress(int workers, int batches, int seed, int max_kb) {
    Crew crew(workers);
    std::mt19937_64 rng((uint64_t)seed);
    auto rnd = [&](uint64_t n) { return (size_t)(rng() % n); };
    long bad = 0;
    for (int it = 0; it < batches; it++) {
        const int n_chunks = 1 + (int)rnd(12);
        std::vector<std::vector<double>> src64(n_chunks);
        std::vector<std::vector<float>> dst32(n_chunks), src32(n_chunks);
        std::vector<std::vector<double>> dst64(n_chunks), acc64(n_chunks), acc_ref(n_chunks);
        Batch b;
        if (rnd(2)) b.width = WIDTH_PCIE_BOUND;
        std::vector<int> kind(n_chunks);
        for (int c = 0; c < n_chunks; c++) {
            const size_t n = rnd((size_t)max_kb * 128) + (rnd(4) == 0 ? 0 : 1);
            kind[c] = (int)rnd(4);
            const size_t task = (size_t)(4096 << rnd(6));
            if (kind[c] == 0) {  // fp64 -> fp32 (upload conversion)
                src64[c].resize(n);
                for (size_t i = 0; i < n; i++) src64[c][i] = (double)(int64_t)(rng() % 2000001) * 1e-3 - 1000.0;
                dst32[c].assign(n, -7.0f);
                b.add(OP_F64_TO_F32, dst32[c].data(), src64[c].data(), n, task);
            } else if (kind[c] == 1) {  // fp32 -> fp64 (download conversion)
                src32[c].resize(n);
                for (size_t i = 0; i < n; i++) src32[c][i] = (float)(int)(rng() % 100001) * 0.25f;
                dst64[c].assign(n, -7.0);
                b.add(OP_F32_TO_F64, dst64[c].data(), src32[c].data(), n, task);
            } else if (kind[c] == 2) {  // += (gradient download)
                src32[c].resize(n);
                acc64[c].resize(n);
                for (size_t i = 0; i < n; i++) { src32[c][i] = (float)(int)(rng() % 1001); acc64[c][i] = (double)(int)(rng() % 77); }
                acc_ref[c] = acc64[c];
                for (size_t i = 0; i < n; i++) acc_ref[c][i] += (double)src32[c][i];
                b.add(OP_F32_ADD_F64, acc64[c].data(), src32[c].data(), n, task);
            } else {  // mirror comparison, equal unless poisoned
                src64[c].resize(n);
                dst32[c].resize(n);
                for (size_t i = 0; i < n; i++) { src64[c][i] = (double)(int)(rng() % 5001) * 0.5; dst32[c][i] = (float)src64[c][i]; }
                b.add(OP_EQ_F32, dst32[c].data(), src64[c].data(), n, task);
            }
        }
        const int mode = (int)rnd(3);
        if (mode == 0) {  // plain
            crew.run(&b);
        } else if (mode == 1) {  // upload: the coordinator consumes the chunks in order as they complete
            b.open_all();
            crew.start(&b);
            for (int c = 0; c < n_chunks; c++) {
                b.wait_chunk(c);
                if (kind[c] == 0)  // the data must be complete at this point (this is where the DMA would be enqueued)
                    for (size_t i = 0; i < src64[c].size(); i += 97) bad += dst32[c][i] != (float)src64[c][i];
            }
            crew.finish(&b);
        } else {  // download: gates open one by one, with pauses
            crew.start(&b);
            for (int c = 0; c < n_chunks; c++) {
                if (rnd(3) == 0) std::this_thread::sleep_for(std::chrono::microseconds(rnd(300)));
                b.open_chunks.store(c + 1, std::memory_order_release);
            }
            crew.finish(&b);
        }
        for (int c = 0; c < n_chunks; c++) {
            if (kind[c] == 0) for (size_t i = 0; i < src64[c].size(); i++) bad += dst32[c][i] != (float)src64[c][i];
            if (kind[c] == 1) for (size_t i = 0; i < src32[c].size(); i++) bad += dst64[c][i] != (double)src32[c][i];
            if (kind[c] == 2) for (size_t i = 0; i < acc64[c].size(); i++) bad += acc64[c][i] != acc_ref[c][i];
        }
        bad += b.unequal.load() != 0;
        if (rnd(5) == 0) std::this_thread::sleep_for(std::chrono::microseconds(rnd(1500)));  // let the workers go to sleep
    }
    return bad;
}

// a poisoned comparison must be reported
extern "C" int hook_crew_detects_difference(int workers) {
    Crew crew(workers);
    std::vector<double> user(100000);
    std::vector<float> mirror(user.size());
    for (size_t i = 0; i < user.size(); i++) { user[i] = (double)i * 0.5; mirror[i] = (float)user[i]; }
    user[77777] += 1.0;
    Batch b;
    b.add(OP_EQ_F32, mirror.data(), user.data(), user.size());
    crew.run(&b);
    return b.unequal.load();
}
