// pack_compact.cpp -- host packer of the compact stream: 24-byte public records (raftgpu_append_resp,
// what Raft::handle_append_response consumes, raft.rs:1559) -> 4-byte units (include/raftgpu.h
// "compact stream").  Plain C++ (g++), no CUDA: this is the CPU work that stands between a caller's
// batch and the H2D copy, so it is the end-to-end cost of a step.
//
// Two implementations with byte-identical output:
//   pack_range_scalar  one record at a time (the definition of the format);
//   pack_range         AVX-512: run boundaries for 64 records at a time from one pass over the group
//                      words, then ONE 8-lane pass per run -- no data-dependent branch per record, and
//                      no loop-carried dependency between runs except the output position.  Runs the
//                      vector form cannot take (a REJECT / EXT inside, values out of the compact
//                      ranges, more than 8 records) go through the scalar state machine.
#include "pack_compact.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace raftgpu {

int32_t pack_range_scalar(const raftgpu_append_resp *records, uint64_t lo, uint64_t hi, uint64_t n_total, PackState &o,
                          uint32_t *unit_of_record, uint32_t unit_base) {
    // One record at a time.  A run starts when the group changes (or 8 units are used up); its base
    // index is taken from its FIRST record (index - 8192), so the header can be written before the
    // rest of the run is seen and every later record of the run within [-8192, +8191] of the first one
    // fits the 14-bit delta -- no look-ahead, no second loop.
    uint32_t *units = o.units;
    uint32_t *g_base = o.g_base;
    const uint64_t unit_cap = o.unit_cap;
    uint64_t nu = o.nu;
    bool in_run = o.in_run, header = o.header, have_prev_group = o.have_prev_group;
    uint32_t cur_g = o.cur_g, run_units = o.run_units, back = o.back, prev_group = o.prev_group, seen_slots = o.seen_slots;
    uint64_t base = o.base;
    auto save = [&] {
        o.nu = nu;
        o.in_run = in_run;
        o.header = header;
        o.have_prev_group = have_prev_group;
        o.cur_g = cur_g;
        o.run_units = run_units;
        o.back = back;
        o.prev_group = prev_group;
        o.seen_slots = seen_slots;
        o.base = base;
    };
    for (uint64_t i = lo; i < hi; i++) {
        const raftgpu_append_resp &r = records[i];
        if (r.flags & RAFTGPU_REC_EXT) {  // rides with the REJECT in front of it (or carries nothing)
            if (unit_of_record) unit_of_record[i] = UINT32_MAX;
            continue;
        }
        const uint32_t need = (r.flags & RAFTGPU_REC_REJECT) ? 2u : 1u;  // a REJECT may take a payload unit
        if (!in_run || r.group != cur_g || run_units + need > 8u) {
            // ---- a new run
            const uint32_t g = r.group;
            if (nu + 2 + 8 > unit_cap) {
                save();
                return RAFTGPU_ERR_FULL;
            }
            base = r.index > 0x2000u ? r.index - 0x2000u : 0;
            header = base < (1ull << 48);
            if (header) {
                const uint64_t b = nu / RAFTGPU_COMPACT_BLOCK;
                if (b >= o.gbase_cap) {
                    save();
                    return RAFTGPU_ERR_FULL;
                }
                while (o.blocks_set <= b) g_base[o.blocks_set++] = g;  // first header of the block names its g_base
                const uint32_t gb = g_base[b];
                if (g < gb || g - gb > 0xfffu) {
                    header = false;
                } else {
                    units[nu++] = kCuHdrA | (static_cast<uint32_t>(base & 0x3fffffffu) << 2);
                    units[nu++] = kCuHdrB | ((g - gb) << 2) | (static_cast<uint32_t>(base >> 30) << 14);
                }
            }
            if (!header || (have_prev_group && g < prev_group)) o.tileable = false;
            if (!have_prev_group || g != prev_group) seen_slots = 0;
            if (!o.any) {
                o.any = true;
                o.first_group = g;
            }
            o.last_group = g;
            prev_group = g;
            have_prev_group = true;
            in_run = true;
            cur_g = g;
            run_units = 0;
            back = 0;
        }
        run_units += need;
        o.n_rec++;
        if (r.peer_slot < RAFTGPU_SLOTS) {
            if ((seen_slots >> r.peer_slot) & 1u) o.one_wave = false;
            seen_slots |= 1u << r.peer_slot;
        }
        const bool is_local = r.flags == RAFTGPU_REC_LOCAL, is_reject = r.flags == RAFTGPU_REC_REJECT;
        bool compact = header && (r.flags == 0 || is_local || is_reject) && r.peer_slot < RAFTGPU_SLOTS &&
                       r.index >= base && r.index - base <= 0x3fffu;
        uint32_t cd = 0, payload = 0;
        if (compact) {
            if (is_local) {
                if (r.commit == 0)
                    cd = kCuNoCommit;
                else if (r.commit >= r.index && r.commit - r.index < kCuNoCommit)
                    cd = static_cast<uint32_t>(r.commit - r.index);
                else
                    compact = false;
            } else if (r.commit <= r.index && r.index - r.commit <= 255u) {
                cd = static_cast<uint32_t>(r.index - r.commit);
            } else {
                compact = false;
            }
        }
        if (compact && is_reject) {
            // the EXT's next_probe_index hint as a signed 29-bit delta from the index; a snapshot
            // request (rare) sends the record to the side table
            const bool has_ext = i + 1 < n_total && (records[i + 1].flags & RAFTGPU_REC_EXT);
            const uint64_t hint = has_ext ? records[i + 1].index : 0;
            const uint64_t snapshot = has_ext ? records[i + 1].commit : RAFTGPU_INVALID_INDEX;
            const int64_t d = static_cast<int64_t>(hint - r.index);
            if (snapshot != RAFTGPU_INVALID_INDEX || d < -(1ll << 28) || d >= (1ll << 28))
                compact = false;
            else
                payload = kCuEsc | ((kCuPayload | (static_cast<uint32_t>(d) & (kCuPayload - 1u))) << 2);
        }
        if (compact) {
            units[nu] = kCuRec | (is_local ? kCuLocal : 0u) | (is_reject ? kCuReject : 0u) | (back << 3) |
                        (static_cast<uint32_t>(r.peer_slot) << 6) | (static_cast<uint32_t>(r.index - base) << 10) | (cd << 24);
            if (unit_of_record) unit_of_record[i] = unit_base + static_cast<uint32_t>(nu);
            nu++;
            back++;
            if (is_reject) {
                units[nu++] = payload;
                back++;
            }
        } else {  // the record (and its EXT) to the side table, one ESC unit
            if (nu >= unit_cap || o.side.size() >= kCuPad - 2) {
                save();
                return RAFTGPU_ERR_FULL;
            }
            units[nu] = kCuEsc | (static_cast<uint32_t>(o.side.size()) << 2);
            if (o.want_esc_pos) o.esc_pos.push_back(static_cast<uint32_t>(nu));
            if (unit_of_record) unit_of_record[i] = unit_base + static_cast<uint32_t>(nu);
            nu++;
            back++;
            o.side.push_back(r);
            if ((r.flags & RAFTGPU_REC_REJECT) && i + 1 < n_total && (records[i + 1].flags & RAFTGPU_REC_EXT))
                o.side.push_back(records[i + 1]);
        }
    }
    save();
    return RAFTGPU_OK;
}

#if defined(__x86_64__)
#define RAFTGPU_AVX512 __attribute__((target("avx512f,avx512bw,avx512dq,avx512vl,bmi,bmi2,lzcnt,popcnt")))

// records[i, i+8) as three vectors of their 64-bit words: W = {group, peer_slot, flags}, I = index, C = commit
RAFTGPU_AVX512 static inline void load8(const raftgpu_append_resp *p, __m512i &W, __m512i &I, __m512i &C) {
    const __m512i z0 = _mm512_loadu_si512(reinterpret_cast<const char *>(p));
    const __m512i z1 = _mm512_loadu_si512(reinterpret_cast<const char *>(p) + 64);
    const __m512i z2 = _mm512_loadu_si512(reinterpret_cast<const char *>(p) + 128);
    // word k of record j is 64-bit word 3j + k of the 24 loaded: z0 holds 0..7, z1 8..15, z2 16..23
    const __m512i w01 = _mm512_setr_epi64(0, 3, 6, 9, 12, 15, 0, 0), w2 = _mm512_setr_epi64(0, 0, 0, 0, 0, 0, 2, 5);
    const __m512i i01 = _mm512_setr_epi64(1, 4, 7, 10, 13, 0, 0, 0), i2 = _mm512_setr_epi64(0, 0, 0, 0, 0, 0, 3, 6);
    const __m512i c01 = _mm512_setr_epi64(2, 5, 8, 11, 14, 0, 0, 0), c2 = _mm512_setr_epi64(0, 0, 0, 0, 0, 1, 4, 7);
    W = _mm512_mask_permutexvar_epi64(_mm512_permutex2var_epi64(z0, w01, z1), 0xC0, w2, z2);
    I = _mm512_mask_permutexvar_epi64(_mm512_permutex2var_epi64(z0, i01, z1), 0xE0, i2, z2);
    C = _mm512_mask_permutexvar_epi64(_mm512_permutex2var_epi64(z0, c01, z1), 0xE0, c2, z2);
}

RAFTGPU_AVX512 static int32_t pack_range_avx512(const raftgpu_append_resp *records, uint64_t lo, uint64_t hi, uint64_t n_total,
                                                PackState &st, uint32_t *unit_of_record, uint32_t unit_base) {
    static_assert(sizeof(raftgpu_append_resp) == 24, "record layout");
    uint64_t i = lo;
    const __m512i gather_idx = _mm512_setr_epi64(0, 3, 6, 9, 12, 15, 18, 21);  // record j's first word, in 8-byte units
    const __m512i lane3 = _mm512_setr_epi64(0 << 3, 1 << 3, 2 << 3, 3 << 3, 4 << 3, 5 << 3, 6 << 3, 7 << 3);
    // a record the vector form may take: flags in {0, LOCAL}, peer_slot < 8 (reserved bits ignored)
    const __m512i dirty_bits = _mm512_set1_epi64(0x0000fdf800000000ll);
    const __m512i lo32 = _mm512_set1_epi64(0xffffffffll);
    // Window of 64 records starting at c0: bit k of `starts` = record c0+k opens a new stretch of equal
    // groups, bit k of `clean` = record c0+k is vector material.  A run's end is searched at most 9
    // records ahead, so a window serves starts up to c0+54 and is then re-based.
    while (i < hi && i + 64 + 8 <= n_total) {
        const uint64_t c0 = i;
        uint64_t starts = 0, clean = 0;
        {
            const long long *w = reinterpret_cast<const long long *>(records + c0);
            // the group before the window: a real record, or (at the very start) anything that differs
            long long before = c0 > 0 ? reinterpret_cast<const long long *>(records + c0 - 1)[0] : ~w[0];
            __m512i prev = _mm512_set1_epi64(before);
            for (int j = 0; j < 8; j++) {
                const __m512i W = _mm512_i64gather_epi64(gather_idx, w + 24 * j, 8);
                const __m512i P = _mm512_alignr_epi64(W, prev, 7);  // lane k = word of record c0 + 8j + k - 1
                const __mmask8 ks = _mm512_cmpneq_epu64_mask(_mm512_and_si512(W, lo32), _mm512_and_si512(P, lo32));
                const __mmask8 kc = _mm512_testn_epi64_mask(W, dirty_bits);
                starts |= static_cast<uint64_t>(ks) << (8 * j);
                clean |= static_cast<uint64_t>(kc) << (8 * j);
                prev = W;
            }
        }
        while (i < hi && i - c0 < 55) {
            const uint32_t off = static_cast<uint32_t>(i - c0);
            // length of the stretch of equal groups that starts at i: 1..8, or 9 = "more than 8"
            uint32_t len = static_cast<uint32_t>(_tzcnt_u64((starts >> (off + 1)) | 0x100u)) + 1u;
            const bool too_long = len > 8;
            if (too_long) len = 8;
            if (i + len > hi) len = static_cast<uint32_t>(hi - i);
            const uint32_t lenmask = (1u << len) - 1u;
            const raftgpu_append_resp *p = records + i;
            const uint32_t g = p->group;
            bool vec = !too_long && ((clean >> off) & lenmask) == lenmask && !(st.in_run && st.cur_g == g);
            if (vec) {
                const uint64_t idx0 = p->index;
                const uint64_t base = idx0 > 0x2000u ? idx0 - 0x2000u : 0;
                const uint64_t b = st.nu / RAFTGPU_COMPACT_BLOCK;
                if (st.nu + 2 + 8 > st.unit_cap) return RAFTGPU_ERR_FULL;  // what the scalar form answers at a run start
                const uint32_t gb = (st.blocks_set <= b || b >= st.gbase_cap) ? g : st.g_base[b];
                vec = b < st.gbase_cap && base < (1ull << 48) && g >= gb && g - gb <= 0xfffu;
                if (vec) {
                    __m512i W, I, C;
                    load8(p, W, I, C);
                    const __mmask8 kloc = _mm512_test_epi64_mask(W, _mm512_set1_epi64(1ll << 41));  // flags & LOCAL
                    const __m512i D = _mm512_sub_epi64(I, _mm512_set1_epi64(static_cast<long long>(base)));
                    const __mmask8 kd = _mm512_cmple_epu64_mask(D, _mm512_set1_epi64(0x3fff));
                    const __m512i CDM = _mm512_sub_epi64(I, C);  // message: index - commit
                    const __mmask8 kcm = _mm512_cmple_epu64_mask(CDM, _mm512_set1_epi64(255));
                    const __m512i CDL = _mm512_sub_epi64(C, I);  // LOCAL: commit - index, or 255 for "no new last_index"
                    const __mmask8 kzero = _mm512_testn_epi64_mask(C, C);
                    const __mmask8 kcl = _mm512_cmplt_epu64_mask(CDL, _mm512_set1_epi64(255)) | kzero;
                    const __m512i CDLv = _mm512_mask_mov_epi64(CDL, kzero, _mm512_set1_epi64(255));
                    const __m512i CD = _mm512_mask_blend_epi64(kloc, CDM, CDLv);
                    const __mmask8 ok = kd & ((kloc & kcl) | (~kloc & kcm));
                    vec = (ok & lenmask) == lenmask;
                    if (vec) {
                        // ---- commit the run: g_base, header, 8 units (lanes >= len are overwritten by the next run)
                        uint32_t *units = st.units;
                        uint64_t nu = st.nu;
                        while (st.blocks_set <= b) st.g_base[st.blocks_set++] = g;
                        units[nu] = kCuHdrA | (static_cast<uint32_t>(base & 0x3fffffffu) << 2);
                        units[nu + 1] = kCuHdrB | ((g - gb) << 2) | (static_cast<uint32_t>(base >> 30) << 14);
                        const __m512i slot = _mm512_and_si512(_mm512_srli_epi64(W, 32), _mm512_set1_epi64(7));
                        __m512i U = _mm512_or_si512(_mm512_slli_epi64(slot, 6), _mm512_slli_epi64(D, 10));
                        U = _mm512_or_si512(U, _mm512_slli_epi64(CD, 24));
                        U = _mm512_or_si512(U, lane3);
                        U = _mm512_mask_or_epi64(U, kloc, U, _mm512_set1_epi64(kCuLocal));
                        _mm256_storeu_si256(reinterpret_cast<__m256i *>(units + nu + 2), _mm512_cvtepi64_epi32(U));
                        if (unit_of_record)
                            for (uint32_t k = 0; k < len; k++) unit_of_record[i + k] = unit_base + static_cast<uint32_t>(nu + 2 + k);
                        // one record per (group, peer) cell?
                        const __m512i bits = _mm512_sllv_epi64(_mm512_set1_epi64(1), slot);
                        const uint32_t seen = static_cast<uint32_t>(_mm512_mask_reduce_or_epi64(static_cast<__mmask8>(lenmask), bits));
                        if (static_cast<uint32_t>(__builtin_popcount(seen)) != len) st.one_wave = false;
                        if (st.have_prev_group && g < st.prev_group) st.tileable = false;
                        if (!st.any) {
                            st.any = true;
                            st.first_group = g;
                        }
                        st.last_group = g;
                        st.prev_group = g;
                        st.have_prev_group = true;
                        st.seen_slots = seen;
                        st.in_run = true;
                        st.header = true;
                        st.cur_g = g;
                        st.run_units = len;
                        st.back = len;
                        st.base = base;
                        st.nu = nu + 2 + len;
                        st.n_rec += len;
                    }
                }
            }
            if (!vec) {
                const int32_t rc = pack_range_scalar(records, i, i + len, n_total, st, unit_of_record, unit_base);
                if (rc != RAFTGPU_OK) return rc;
            }
            i += len;
        }
    }
    if (i < hi) return pack_range_scalar(records, i, hi, n_total, st, unit_of_record, unit_base);
    return RAFTGPU_OK;
}

static bool use_avx512() {
    static const bool on = [] {
        const char *e = getenv("RAFTGPU_PACK_SCALAR");
        if (e && e[0] == '1') return false;
        return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512dq") &&
               __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("bmi");
    }();
    return on;
}
#else
static bool use_avx512() { return false; }
#endif

const char *pack_impl() { return use_avx512() ? "avx512" : "scalar"; }

int32_t pack_range(const raftgpu_append_resp *records, uint64_t lo, uint64_t hi, uint64_t n_total, PackState &st,
                   uint32_t *unit_of_record, uint32_t unit_base) {
#if defined(__x86_64__)
    if (use_avx512()) return pack_range_avx512(records, lo, hi, n_total, st, unit_of_record, unit_base);
#endif
    return pack_range_scalar(records, lo, hi, n_total, st, unit_of_record, unit_base);
}

}  // namespace raftgpu

// ===========================================================================
using namespace raftgpu;

static inline uint64_t align16(uint64_t x) { return (x + 15u) & ~15ull; }

extern "C" {

uint64_t raftgpu_compact_bound(uint64_t n) {
    // worst case per record: its own run (2 header units + 1) or an ESC unit plus 24 side bytes
    const uint64_t units = 3 * n + 8;
    return sizeof(raftgpu_compact_hdr) + align16(4 * (units / RAFTGPU_COMPACT_BLOCK + 2)) + align16(4 * units) + 24 * n + 64;
}

int32_t raftgpu_pack_compact(const raftgpu_append_resp *records, uint64_t n, void *out, uint64_t out_capacity,
                             uint64_t *out_bytes, uint32_t *unit_of_record) {
    if ((!records && n) || !out || !out_bytes) return RAFTGPU_ERR_INVALID;
    if (reinterpret_cast<uintptr_t>(out) & 15u) return RAFTGPU_ERR_INVALID;
    const uint64_t max_units = 3 * n + 8;
    const uint64_t off_blocks = sizeof(raftgpu_compact_hdr);
    const uint64_t off_units = off_blocks + align16(4 * (max_units / RAFTGPU_COMPACT_BLOCK + 2));
    if (off_units > out_capacity) return RAFTGPU_ERR_FULL;
    uint8_t *blob = static_cast<uint8_t *>(out);
    PackState o;
    o.g_base = reinterpret_cast<uint32_t *>(blob + off_blocks);
    o.gbase_cap = max_units / RAFTGPU_COMPACT_BLOCK + 2;
    o.units = reinterpret_cast<uint32_t *>(blob + off_units);
    o.unit_cap = std::min<uint64_t>((out_capacity - off_units) / 4, 0xfffffff0ull);
    const int32_t rc = pack_range(records, 0, n, n, o, unit_of_record, 0);
    if (rc != RAFTGPU_OK) return rc;
    uint64_t nu = o.nu;
    uint32_t *units = o.units;
    const uint64_t n_blocks = (nu + RAFTGPU_COMPACT_BLOCK - 1) / RAFTGPU_COMPACT_BLOCK;
    while (o.blocks_set < n_blocks) o.g_base[o.blocks_set++] = 0;
    while (nu & 3u) {
        if (nu >= o.unit_cap) return RAFTGPU_ERR_FULL;
        units[nu++] = kCuEsc | (kCuPad << 2);  // the fused kernel fetches units in 16-byte pieces
    }
    const uint64_t off_side = off_units + align16(4 * nu);
    const uint64_t total = off_side + align16(o.side.size() * sizeof(raftgpu_append_resp));
    if (total > out_capacity) return RAFTGPU_ERR_FULL;
    if (!o.side.empty()) memcpy(blob + off_side, o.side.data(), o.side.size() * sizeof(raftgpu_append_resp));
    raftgpu_compact_hdr h{};
    h.magic = RAFTGPU_COMPACT_MAGIC;
    h.n_units = static_cast<uint32_t>(nu);
    h.n_blocks = static_cast<uint32_t>(n_blocks);
    h.n_side = static_cast<uint32_t>(o.side.size());
    h.n_records = o.n_rec;
    h.off_blocks = off_blocks;
    h.off_units = off_units;
    h.off_side = off_side;
    h.total_bytes = total;
    h.flags = (o.tileable ? RAFTGPU_COMPACT_TILEABLE : 0u) | (o.tileable && o.one_wave ? RAFTGPU_COMPACT_ONE_WAVE : 0u);
    memcpy(blob, &h, sizeof(h));
    *out_bytes = total;
    return RAFTGPU_OK;
}

}  // extern "C"
