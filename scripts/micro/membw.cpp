// membw.cpp -- host memory bandwidth available to T pinned threads that each stream one slice of a buffer
// (scripts/micro; sizing of the staging path).  g++ -O3 -march=native -pthread membw.cpp -o membw
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sched.h>
#include <thread>
#include <vector>
#include <immintrin.h>

int main(int argc, char **argv) {
    const size_t bytes = 85ull << 20;
    std::vector<int> counts;
    for (int a = 1; a < argc; a++) counts.push_back(atoi(argv[a]));
    if (counts.empty()) counts = {8, 16, 32, 64};
    uint8_t *buf = static_cast<uint8_t *>(aligned_alloc(4096, bytes));
    uint8_t *out = static_cast<uint8_t *>(aligned_alloc(4096, bytes / 4));
    memset(buf, 1, bytes);
    memset(out, 0, bytes / 4);
    for (int T : counts) {
        std::atomic<int> go{0}, done{0};
        std::atomic<bool> stop{false};
        std::vector<uint64_t> sums(T * 16);
        std::vector<std::thread> th;
        int mode = 0;
        for (int t = 0; t < T; t++)
            th.emplace_back([&, t] {
                cpu_set_t one;
                CPU_ZERO(&one);
                CPU_SET(t < 32 ? t : 64 + (t - 32), &one);   // 0-31 physical cores of node 0, then their SMT siblings
                sched_setaffinity(0, sizeof(one), &one);
                int seen = 0;
                for (;;) {
                    while (go.load(std::memory_order_acquire) == seen && !stop.load()) _mm_pause();
                    if (stop.load()) return;
                    seen = go.load();
                    const size_t lo = bytes * t / T & ~size_t(63), hi = bytes * (t + 1) / T & ~size_t(63);
                    __m512i acc = _mm512_setzero_si512();
                    if (mode == 0) {
                        for (size_t o = lo; o < hi; o += 64) acc = _mm512_add_epi64(acc, _mm512_load_si512(buf + o));
                    } else {  // read 4 lines, write 1 (the packer's ratio), regular or streaming stores
                        for (size_t o = lo; o + 256 <= hi; o += 256) {
                            __m512i v = _mm512_add_epi64(_mm512_add_epi64(_mm512_load_si512(buf + o), _mm512_load_si512(buf + o + 64)),
                                                         _mm512_add_epi64(_mm512_load_si512(buf + o + 128), _mm512_load_si512(buf + o + 192)));
                            if (mode == 1) _mm512_store_si512(out + o / 4, v); else _mm512_stream_si512(reinterpret_cast<__m512i *>(out + o / 4), v);
                        }
                    }
                    sums[t * 16] = _mm512_reduce_add_epi64(acc);
                    done.fetch_add(1, std::memory_order_release);
                }
            });
        for (mode = 0; mode < 3; mode++) {
            double best = 1e9;
            for (int rep = 0; rep < 8; rep++) {
                done.store(0);
                auto t0 = std::chrono::steady_clock::now();
                go.fetch_add(1, std::memory_order_release);
                while (done.load(std::memory_order_acquire) != T) _mm_pause();
                best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
            }
            printf("T=%2d %s: %.3f ms, %.1f GB/s read\n", T, mode == 0 ? "read only        " : mode == 1 ? "read 4 : write 1  " : "read 4 : stream 1 ",
                   1e3 * best, bytes / best / 1e9);
        }
        stop.store(true);
        for (auto &x : th) x.join();
    }
    return 0;
}
