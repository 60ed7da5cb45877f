// pack_bench.cpp -- host-only timing of the compact-stream packer (scripts/micro; not part of the library).
//   g++ -O3 -std=c++17 -I include -I raft-rs_b200/csrc scripts/micro/pack_bench.cpp raft-rs_b200/csrc/pack_compact.cpp \
//       raft-rs_b200/csrc/synth.cpp -lpthread -o /tmp/pack_bench && /tmp/pack_bench [n_groups] [threads...]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "pack_compact.h"
#include "raftgpu_synth.h"

using namespace raftgpu;
using clk = std::chrono::steady_clock;

int main(int argc, char **argv) {
    const uint32_t n = argc > 1 ? atoi(argv[1]) : 1000000;
    const uint32_t cap = (n + 127u) & ~127u;
    std::vector<uint64_t> m(8ull * cap), nx(8ull * cap), pc(8ull * cap), acked(8ull * cap), com(cap), ts(cap), li(cap), term(cap), sl(cap);
    std::vector<uint8_t> pf(8ull * cap), sf(8ull * cap);
    std::vector<uint32_t> meta(cap);
    raftgpu_synth_columns c{cap, n, m.data(), nx.data(), pc.data(), pf.data(), meta.data(), com.data(), ts.data(), li.data(), term.data(),
                            acked.data(), sl.data(), sf.data()};
    raftgpu_synth_init(&c, 0x5EED0003, 5, 0);
    std::vector<raftgpu_append_resp> recs(5ull * n + 64);
    uint64_t nr = 0;
    raftgpu_synth_round(&c, 0x5EED0003, 0, 5, recs.data(), recs.size(), &nr);
    std::vector<uint8_t> out(raftgpu_compact_bound(nr) + 64);
    uint8_t *o = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(out.data()) + 15) & ~uintptr_t(15));
    printf("impl %s, %llu records\n", pack_impl(), (unsigned long long)nr);
    for (int rep = 0; rep < 5; rep++) {
        uint64_t nb = 0;
        const int inner = std::max<int>(1, static_cast<int>(4000000 / nr));  // small inputs: repeat (cache-resident timing)
        auto t0 = clk::now();
        for (int k = 0; k < inner; k++) raftgpu_pack_compact(recs.data(), nr, o, out.size() - 64, &nb, nullptr);
        double s = std::chrono::duration<double>(clk::now() - t0).count() / inner;
        printf("  1 thread: %.2f ns/record (%llu bytes)\n", 1e9 * s / nr, (unsigned long long)nb);
    }
    for (int a = 2; a < argc; a++) {
        const int T = atoi(argv[a]);
        std::vector<std::vector<uint32_t>> units(T), gb(T);
        for (int t = 0; t < T; t++) {
            units[t].resize(3 * (nr / T + 64) + 4096);
            gb[t].resize(units[t].size() / RAFTGPU_COMPACT_BLOCK + 2);
        }
        std::vector<uint64_t> cut(T + 1, nr);
        cut[0] = 0;
        for (int t = 1; t < T; t++) {
            uint64_t k = std::max(cut[t - 1], nr * t / T);
            while (k < nr && k > 0 && ((recs[k].flags & RAFTGPU_REC_EXT) || recs[k].group == recs[k - 1].group)) k++;
            cut[t] = k;
        }
        double best = 1e9;
        for (int rep = 0; rep < 5; rep++) {
            std::vector<std::thread> th;
            auto t0 = clk::now();
            for (int t = 0; t < T; t++)
                th.emplace_back([&, t] {
                    PackState st;
                    st.units = units[t].data();
                    st.unit_cap = units[t].size();
                    st.g_base = gb[t].data();
                    st.gbase_cap = gb[t].size();
                    pack_range(recs.data(), cut[t], cut[t + 1], nr, st, nullptr, 0);
                });
            for (auto &x : th) x.join();
            best = std::min(best, std::chrono::duration<double>(clk::now() - t0).count());
        }
        printf("  %d threads: %.3f ms per round, %.2f ns/record/thread\n", T, 1e3 * best, 1e9 * best * T / nr);
    }
    return 0;
}
