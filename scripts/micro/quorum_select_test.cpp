// Differential test of csrc/quorum_select.h (host build) against a plain sort: random values with many ties, every
// pair of masks.  g++ -O2 -std=c++17 -Iraft-rs_b200/csrc scripts/micro/quorum_select_test.cpp
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "quorum_select.h"

static uint64_t by_sort(const uint64_t (&v)[8], uint32_t mask) {
    std::vector<uint64_t> m;
    for (int s = 0; s < 8; s++)
        if ((mask >> s) & 1u) m.push_back(v[s]);
    if (m.empty()) return UINT64_MAX;                       // majority.rs:71-75
    std::sort(m.begin(), m.end(), std::greater<uint64_t>());  // majority.rs:95
    return m[m.size() / 2];                                   // quorum = n / 2 + 1 -> index quorum - 1 (majority.rs:97-99)
}

int main(int argc, char **argv) {
    const int rounds = argc > 1 ? atoi(argv[1]) : 2000;
    std::mt19937_64 rng(12345);
    uint64_t checked = 0;
    for (int r = 0; r < rounds; r++) {
        uint64_t v[8];
        const int kind = r % 4;
        for (int s = 0; s < 8; s++) {
            if (kind == 0) v[s] = rng() % 4;                      // almost everything ties
            else if (kind == 1) v[s] = rng() % 16;
            else if (kind == 2) v[s] = rng();                     // full 64-bit range
            else v[s] = (rng() % 3 == 0) ? UINT64_MAX - rng() % 2 : rng() % 5;
        }
        for (uint32_t in = 0; in < 256; in += (r % 8 == 0 ? 1 : 1 + rng() % 5))
            for (uint32_t out = 0; out < 256; out += 1 + rng() % 7) {
                uint64_t a, b;
                raftgpu::quorum_index_joint(v, in, out, a, b);
                if (a != by_sort(v, in) || b != by_sort(v, out)) {
                    printf("MISMATCH round %d in %02x out %02x: %llu %llu want %llu %llu\n", r, in, out, (unsigned long long)a,
                           (unsigned long long)b, (unsigned long long)by_sort(v, in), (unsigned long long)by_sort(v, out));
                    return 1;
                }
                checked++;
            }
    }
    printf("quorum select ok: %llu (values, in, out) cases\n", (unsigned long long)checked);
    return 0;
}
