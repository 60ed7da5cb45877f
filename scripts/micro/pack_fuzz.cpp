// pack_fuzz.cpp -- differential fuzz of the compact-stream packers: pack_range (AVX-512 when the CPU has
// it) against pack_range_scalar (the definition), on random record arrays and random [lo, hi) ranges,
// continuing one PackState over several calls.  Built and run by tests/test_compact_format.py.
//   g++ -O2 -std=c++17 -I include -I raft-rs_b200/csrc scripts/micro/pack_fuzz.cpp raft-rs_b200/csrc/pack_compact.cpp -o pack_fuzz
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pack_compact.h"

using namespace raftgpu;

static uint64_t rng_state = 0x1234567;
static uint64_t rnd() {
    uint64_t x = (rng_state += 0x9E3779B97F4A7C15ull);
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static uint64_t below(uint64_t n) { return rnd() % n; }

struct Out {
    std::vector<uint32_t> units, gb;
    PackState st;
    explicit Out(size_t n) : units(3 * n + 64 + 4096, 0xdeadbeefu), gb((3 * n + 64 + 4096) / RAFTGPU_COMPACT_BLOCK + 2, 0) {
        st.units = units.data();
        st.unit_cap = units.size();
        st.g_base = gb.data();
        st.gbase_cap = gb.size();
        st.want_esc_pos = true;
    }
};

int main(int argc, char **argv) {
    const int rounds = argc > 1 ? atoi(argv[1]) : 300;
    printf("impl %s\n", pack_impl());
    for (int round = 0; round < rounds; round++) {
        rng_state = 0xABCDEF + 7919ull * round;
        const int mode = round % 4;  // 0: mostly clean traffic, 1: many rejects, 2: hostile values, 3: unsorted / stray EXT
        const size_t n = 200 + below(3000);
        std::vector<raftgpu_append_resp> r;
        uint32_t g = below(1000);
        while (r.size() < n) {
            g += (mode == 3 && below(10) == 0) ? static_cast<uint32_t>(rnd()) : 1 + (below(20) == 0 ? below(5000) : 0);
            const uint64_t base = mode == 2 && below(4) == 0 ? rnd() >> below(64) : 1000 + below(1ull << 40);
            const int len = 1 + (below(30) == 0 ? below(12) : below(6));
            for (int k = 0; k < len; k++) {
                raftgpu_append_resp a{};
                a.group = g;
                a.peer_slot = static_cast<uint8_t>(below(50) == 0 ? below(256) : (mode == 0 ? k % 8 : below(8)));
                a.reserved = below(7) == 0 ? static_cast<uint16_t>(rnd()) : 0;
                a.index = base + below(mode == 2 ? 40000 : 64) - (below(9) == 0 ? below(std::min<uint64_t>(base, 20000) + 1) : 0);
                const uint64_t kind = below(100);
                const uint64_t rej_pct = mode == 1 ? 30 : 3;
                if (kind < rej_pct) {
                    a.flags = RAFTGPU_REC_REJECT;
                    a.commit = a.index - std::min<uint64_t>(a.index, below(8) == 0 ? below(1000) : below(4));
                    r.push_back(a);
                    if (below(10) != 0) {
                        raftgpu_append_resp e{};
                        e.group = below(40) == 0 ? g + 1 : g;
                        e.peer_slot = a.peer_slot;
                        e.flags = RAFTGPU_REC_EXT;
                        e.index = below(6) == 0 ? rnd() >> below(64) : a.index - std::min<uint64_t>(a.index, below(16));
                        e.commit = below(8) == 0 ? 1 + below(1000) : 0;
                        r.push_back(e);
                    }
                    continue;
                } else if (kind < rej_pct + 20) {
                    a.flags = RAFTGPU_REC_LOCAL;
                    a.commit = below(5) == 0 ? 0 : a.index + (below(10) == 0 ? below(600) : below(64));
                    if (below(40) == 0) a.commit = a.index - std::min<uint64_t>(a.index, 1);
                } else {
                    a.flags = below(200) == 0 ? static_cast<uint8_t>(rnd()) : 0;
                    a.commit = a.index - std::min<uint64_t>(a.index, below(12) == 0 ? below(600) : below(4));
                    if (below(60) == 0) a.commit = a.index + 1 + below(5);
                }
                r.push_back(a);
                if (mode == 3 && below(15) == 0) {  // stray EXT
                    raftgpu_append_resp e{};
                    e.group = below(3) == 0 ? static_cast<uint32_t>(rnd()) : g;
                    e.flags = RAFTGPU_REC_EXT;
                    e.index = rnd();
                    r.push_back(e);
                }
            }
        }
        const size_t nt = r.size();
        Out a(nt), b(nt);
        // the same sequence of ranges through both
        size_t pos = below(5);
        int32_t rca = RAFTGPU_OK, rcb = RAFTGPU_OK;
        while (pos < nt && rca == RAFTGPU_OK) {
            const size_t end = std::min(nt, pos + 1 + below(below(3) == 0 ? 40 : 2000));
            rca = pack_range(r.data(), pos, end, nt, a.st, nullptr, 0);
            rcb = pack_range_scalar(r.data(), pos, end, nt, b.st, nullptr, 0);
            if (rca != rcb) {
                printf("round %d: rc %d vs %d\n", round, rca, rcb);
                return 1;
            }
            pos = end;
        }
        const PackState &x = a.st, &y = b.st;
        bool same = x.nu == y.nu && x.n_rec == y.n_rec && x.blocks_set == y.blocks_set && x.tileable == y.tileable &&
                    (x.one_wave == y.one_wave || !x.tileable) && x.any == y.any && x.first_group == y.first_group &&
                    x.last_group == y.last_group && x.side.size() == y.side.size() && x.esc_pos == y.esc_pos &&
                    memcmp(a.units.data(), b.units.data(), 4 * x.nu) == 0 && memcmp(a.gb.data(), b.gb.data(), 4 * x.blocks_set) == 0 &&
                    (x.side.empty() || memcmp(x.side.data(), y.side.data(), x.side.size() * sizeof(raftgpu_append_resp)) == 0);
        if (!same) {
            size_t d = 0;
            while (d < x.nu && d < y.nu && a.units[d] == b.units[d]) d++;
            printf("round %d (mode %d): MISMATCH nu %llu/%llu n_rec %llu/%llu tileable %d/%d one_wave %d/%d first diff unit %zu\n", round, mode,
                   (unsigned long long)x.nu, (unsigned long long)y.nu, (unsigned long long)x.n_rec, (unsigned long long)y.n_rec, x.tileable,
                   y.tileable, x.one_wave, y.one_wave, d);
            return 1;
        }
    }
    printf("fuzz ok: %d rounds\n", rounds);
    return 0;
}
