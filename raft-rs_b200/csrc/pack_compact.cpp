// pack_compact.cpp -- host packer of the compact stream: 24-byte public records (raftgpu_append_resp,
// what Raft::handle_append_response consumes, raft.rs:1559) -> 4-byte units (include/raftgpu.h
// "compact stream").  Plain C++ (g++), no CUDA: this is the CPU work that stands between a caller's
// batch and the H2D copy, so it is the end-to-end cost of a step.
//
// Two implementations with byte-identical output:
//   pack_range_scalar  one record at a time (the definition of the format);
//   pack_range         AVX-512: eight records per step as lane arithmetic, whatever runs they belong
//                      to (see pack_range_avx512); blocks the lanes cannot express go through the
//                      scalar state machine.
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
#define RAFTGPU_AVX512 __attribute__((target("avx512f,avx512bw,avx512dq,avx512vl,avx512cd,avx512vpopcntdq,bmi,bmi2,lzcnt,popcnt")))

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

// T_first[ks][k]: the lane (<= k) at which the run of lane k starts, for the 8-bit run-start mask ks
// (bit j = lane j opens a run); 8 = the run was opened before this block (carry).
struct FirstTable {
    alignas(64) uint8_t t[256][8];
    FirstTable() {
        for (int ks = 0; ks < 256; ks++) {
            int cur = 8;
            for (int k = 0; k < 8; k++) {
                if ((ks >> k) & 1) cur = k;
                t[ks][k] = static_cast<uint8_t>(cur);
            }
        }
    }
};
static const FirstTable kFirst;

// The vector form: EIGHT records per step, whatever runs they belong to.  All per-record work (run
// starts, the run's base index, position inside the run, the delta fits, the unit words, the headers
// and REJECT payloads) is lane arithmetic; the units of the block -- [HDR_A HDR_B] rec [payload] per
// lane -- are laid out as 32 candidates and squeezed together with two vpcompressd.  Between blocks only
// a handful of scalars carry over (output position, the open run's group / base / length / slots), so
// consecutive blocks overlap in the pipeline.  Whatever the lanes cannot express (an unknown flag
// combination, a delta that does not fit, a run reaching 8 units, a header that needs a new g_base
// word, ...) sends that block through the scalar state machine, which also defines the result.
RAFTGPU_AVX512 static int32_t pack_range_avx512(const raftgpu_append_resp *records, uint64_t lo, uint64_t hi, uint64_t n_total,
                                                PackState &st, uint32_t *unit_of_record, uint32_t unit_base) {
    static_assert(sizeof(raftgpu_append_resp) == 24, "record layout");
    if (unit_of_record) return pack_range_scalar(records, lo, hi, n_total, st, unit_of_record, unit_base);  // tests only
    const __m512i zero = _mm512_setzero_si512();
    const __m512i lo32 = _mm512_set1_epi64(0xffffffffll);
    const __m512i c255 = _mm512_set1_epi64(255), c8 = _mm512_set1_epi64(8), one = _mm512_set1_epi64(1);
    const __m512i idx0 = _mm512_setr_epi64(0, 8, 1, 9, 2, 10, 3, 11), idx1 = _mm512_setr_epi64(4, 12, 5, 13, 6, 14, 7, 15);
    uint64_t i = lo;
    while (i + 8 <= hi) {
        const raftgpu_append_resp *p = records + i;
        __m512i W, I, C;
        load8(p, W, I, C);
        const __m512i F = _mm512_and_si512(_mm512_srli_epi64(W, 40), c255);  // flags
        const __m512i S = _mm512_and_si512(_mm512_srli_epi64(W, 32), c255);  // peer slot
        const __mmask8 kext = _mm512_cmpeq_epu64_mask(F, _mm512_set1_epi64(RAFTGPU_REC_EXT));
        const __mmask8 kloc = _mm512_cmpeq_epu64_mask(F, _mm512_set1_epi64(RAFTGPU_REC_LOCAL));
        const __mmask8 krej_all = _mm512_cmpeq_epu64_mask(F, one);  // RAFTGPU_REC_REJECT
        const __mmask8 kacc = _mm512_testn_epi64_mask(F, F);
        const __mmask8 kslot = _mm512_cmplt_epu64_mask(S, c8) | kext;
        bool vec = ((kext | kloc | krej_all | kacc) & kslot) == 0xff;
        // a REJECT in the last lane has its EXT in the next block: leave it for the next step
        const uint32_t nproc = (krej_all & 0x80) ? 7u : 8u;
        const __mmask8 live = static_cast<__mmask8>((1u << nproc) - 1u);
        const __mmask8 krej = krej_all & live, kmain = static_cast<__mmask8>(~kext) & live;
        const __m512i G = _mm512_and_si512(W, lo32);
        const __m512i Gprev = _mm512_alignr_epi64(G, _mm512_set1_epi64(st.cur_g), 7);
        __mmask8 ks = _mm512_cmpneq_epu64_mask(G, Gprev) & live;
        if (!st.in_run) ks |= 1;
        vec = vec && !(ks & kext);  // an EXT whose group differs from its predecessor's: the scalar form sorts it out
        const uint64_t nu = st.nu, b = nu / RAFTGPU_COMPACT_BLOCK;
        vec = vec && nu + 32 <= st.unit_cap && (nu + 31) / RAFTGPU_COMPACT_BLOCK == b && st.blocks_set > b;
        const __mmask8 kc = ks ? static_cast<__mmask8>((ks & (0u - ks)) - 1u) : 0xff;  // lanes of the run carried in
        vec = vec && !((kc & kmain) && !(st.in_run && st.header));
        if (vec) {
            // units per lane (0 EXT, 1 record, 2 REJECT + payload) and their running sum
            const __m512i CNT = _mm512_mask_add_epi64(_mm512_maskz_mov_epi64(kmain, one), krej, one, one);
            __m512i x = _mm512_add_epi64(CNT, _mm512_alignr_epi64(CNT, zero, 7));
            x = _mm512_add_epi64(x, _mm512_alignr_epi64(x, zero, 6));
            x = _mm512_add_epi64(x, _mm512_alignr_epi64(x, zero, 4));
            const __m512i E = _mm512_sub_epi64(x, CNT);  // exclusive
            const __m512i idxFirst = _mm512_cvtepu8_epi64(_mm_loadl_epi64(reinterpret_cast<const __m128i *>(kFirst.t[ks])));
            // units of the lane's run in front of it (the `back` field): E - E[first], carried run: + st.back
            const __m512i BACK = _mm512_sub_epi64(
                E, _mm512_permutex2var_epi64(E, idxFirst, _mm512_set1_epi64(-static_cast<long long>(st.back))));
            // the scalar form opens a new run when run_units + need > 8; run_units counts 2 for every REJECT, also
            // one that went to the side table as ONE unit, so the carried run may be ahead of its `back`
            const __m512i RU = _mm512_mask_add_epi64(BACK, kc, BACK, _mm512_set1_epi64(st.run_units - st.back));
            const __mmask8 kover = _mm512_cmpgt_epu64_mask(_mm512_add_epi64(RU, CNT), c8) & live;
            // the run's base: its first record's index - 8192 (saturating)
            const __m512i c2000 = _mm512_set1_epi64(0x2000);
            const __m512i baseRec = _mm512_sub_epi64(_mm512_max_epu64(I, c2000), c2000);
            const __m512i BASE = _mm512_permutex2var_epi64(baseRec, idxFirst, _mm512_set1_epi64(static_cast<long long>(st.base)));
            const __m512i D = _mm512_sub_epi64(I, BASE);
            const __mmask8 kd = _mm512_cmple_epu64_mask(D, _mm512_set1_epi64(0x3fff));
            // headers: base < 2^48, group within 4095 of the block's g_base
            const uint32_t gb = st.g_base[b];
            const __m512i GD = _mm512_sub_epi64(G, _mm512_set1_epi64(gb));
            const __mmask8 khdr = _mm512_cmple_epu64_mask(GD, _mm512_set1_epi64(0xfff)) &
                                  _mm512_cmplt_epu64_mask(BASE, _mm512_set1_epi64(1ll << 48));
            // commit deltas
            const __m512i CDM = _mm512_sub_epi64(I, C);  // message (accept / reject): index - commit
            const __mmask8 kcm = _mm512_cmple_epu64_mask(CDM, c255);
            const __m512i CDL = _mm512_sub_epi64(C, I);  // LOCAL: commit - index, or 255 for "no new last_index"
            const __mmask8 kzero = _mm512_testn_epi64_mask(C, C);
            const __mmask8 kcl = _mm512_cmplt_epu64_mask(CDL, c255) | kzero;
            const __m512i CD = _mm512_mask_blend_epi64(kloc, CDM, _mm512_mask_mov_epi64(CDL, kzero, c255));
            const __mmask8 kcok = (kloc & kcl) | (static_cast<__mmask8>(~kloc) & kcm);
            // REJECT payload: the EXT behind it gives next_probe_index (hint) and request_snapshot
            __m512i DH = zero;
            __mmask8 krejok = 0xff;
            if (krej) {
                const __mmask8 kextnext = static_cast<__mmask8>(kext >> 1);
                const __m512i HINT = _mm512_maskz_mov_epi64(kextnext, _mm512_alignr_epi64(zero, I, 1));
                const __m512i SNAP = _mm512_maskz_mov_epi64(kextnext, _mm512_alignr_epi64(zero, C, 1));
                DH = _mm512_sub_epi64(HINT, I);
                krejok = _mm512_cmplt_epu64_mask(_mm512_add_epi64(DH, _mm512_set1_epi64(1ll << 28)), _mm512_set1_epi64(1ll << 29)) &
                         _mm512_testn_epi64_mask(SNAP, SNAP);
            }
            const __mmask8 good = (static_cast<__mmask8>(~kmain) | (kd & kcok)) & (static_cast<__mmask8>(~krej) | krejok) &
                                  (static_cast<__mmask8>(~ks) | khdr);
            vec = !kover && (good & live) == live;
            if (vec) {
                // ---- the unit words
                __m512i U = _mm512_or_si512(_mm512_slli_epi64(S, 6), _mm512_slli_epi64(D, 10));
                U = _mm512_or_si512(U, _mm512_slli_epi64(CD, 24));
                U = _mm512_or_si512(U, _mm512_slli_epi64(BACK, 3));
                U = _mm512_mask_or_epi64(U, kloc, U, _mm512_set1_epi64(kCuLocal));
                U = _mm512_mask_or_epi64(U, krej, U, _mm512_set1_epi64(kCuReject));
                const __m512i P = _mm512_or_si512(
                    _mm512_set1_epi64(kCuEsc | (static_cast<uint64_t>(kCuPayload) << 2)),
                    _mm512_slli_epi64(_mm512_and_si512(DH, _mm512_set1_epi64(kCuPayload - 1u)), 2));
                const __m512i HA = _mm512_or_si512(_mm512_set1_epi64(kCuHdrA),
                                                   _mm512_slli_epi64(_mm512_and_si512(BASE, _mm512_set1_epi64(0x3fffffff)), 2));
                const __m512i HB = _mm512_or_si512(_mm512_or_si512(_mm512_set1_epi64(kCuHdrB), _mm512_slli_epi64(GD, 2)),
                                                   _mm512_slli_epi64(_mm512_srli_epi64(BASE, 30), 14));
                // lane k's four candidates [HDR_A HDR_B rec payload] side by side: two 64-bit words per lane
                const __m512i V1 = _mm512_or_si512(HA, _mm512_slli_epi64(HB, 32)), V2 = _mm512_or_si512(U, _mm512_slli_epi64(P, 32));
                const __m512i X0 = _mm512_permutex2var_epi64(V1, idx0, V2), X1 = _mm512_permutex2var_epi64(V1, idx1, V2);
                const uint32_t m32 = static_cast<uint32_t>(_pdep_u32(ks, 0x11111111u)) * 3u | static_cast<uint32_t>(_pdep_u32(kmain, 0x44444444u)) |
                                     static_cast<uint32_t>(_pdep_u32(krej, 0x88888888u));
                uint32_t *out = st.units + nu;
                const uint32_t n0 = static_cast<uint32_t>(__builtin_popcount(m32 & 0xffffu));
                _mm512_storeu_si512(out, _mm512_maskz_compress_epi32(static_cast<__mmask16>(m32 & 0xffffu), X0));
                _mm512_storeu_si512(out + n0, _mm512_maskz_compress_epi32(static_cast<__mmask16>(m32 >> 16), X1));
                // ---- one record per (group, peer) cell?  Every main lane sets bit (8 * run ordinal + slot) of a
                // 64-bit word (at most 8 runs touch a block): a cell hit twice leaves fewer bits than lanes.
                const __m512i RI = _mm512_popcnt_epi64(_mm512_and_si512(_mm512_set1_epi64(ks), _mm512_setr_epi64(1, 3, 7, 15, 31, 63, 127, 255)));
                const __m512i sh = _mm512_add_epi64(S, _mm512_slli_epi64(_mm512_and_si512(RI, _mm512_set1_epi64(7)), 3));
                const uint64_t cells = static_cast<uint64_t>(_mm512_reduce_or_epi64(_mm512_maskz_sllv_epi64(kmain, one, sh)));
                if (__builtin_popcountll(cells) != __builtin_popcount(kmain) || (!(ks & 1) && (cells & 0xffu & st.seen_slots)))
                    st.one_wave = false;
                if (kmain) {
                    const uint32_t last = nproc - 1;
                    const uint32_t g_last = p[last].group;
                    if (_mm512_cmplt_epu64_mask(G, Gprev) & ks & (st.have_prev_group ? 0xff : 0xfe)) st.tileable = false;
                    if (!st.any) {
                        st.any = true;
                        st.first_group = p[_tzcnt_u32(kmain)].group;
                    }
                    if (ks) {
                        const uint32_t ls = 31u - static_cast<uint32_t>(__builtin_clz(static_cast<uint32_t>(ks)));
                        const __mmask8 tail = static_cast<__mmask8>(0xffu << ls);
                        st.back = st.run_units = static_cast<uint32_t>(__builtin_popcount(kmain & tail) + __builtin_popcount(krej & tail));
                        const uint64_t idx = p[ls].index;
                        st.base = idx > 0x2000u ? idx - 0x2000u : 0;
                        st.header = true;
                        st.seen_slots = static_cast<uint32_t>(cells >> (8 * (__builtin_popcount(ks) & 7))) & 0xffu;
                    } else {
                        const uint32_t added = static_cast<uint32_t>(__builtin_popcount(kmain) + __builtin_popcount(krej));
                        st.back += added;
                        st.run_units += added;
                        st.seen_slots |= static_cast<uint32_t>(cells & 0xffu);
                    }
                    st.in_run = true;
                    st.have_prev_group = true;
                    st.cur_g = st.prev_group = st.last_group = g_last;
                }
                st.nu = nu + static_cast<uint32_t>(__builtin_popcount(m32));
                st.n_rec += static_cast<uint32_t>(__builtin_popcount(kmain));
                i += nproc;
            }
        }
        if (!vec) {
            const int32_t rc = pack_range_scalar(records, i, i + 8, n_total, st, nullptr, unit_base);
            if (rc != RAFTGPU_OK) return rc;
            i += 8;
        }
    }
    if (i < hi) return pack_range_scalar(records, i, hi, n_total, st, nullptr, unit_base);
    return RAFTGPU_OK;
}

static bool use_avx512() {
    static const bool on = [] {
        const char *e = getenv("RAFTGPU_PACK_SCALAR");
        if (e && e[0] == '1') return false;
        return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512dq") &&
               __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx512vpopcntdq") && __builtin_cpu_supports("bmi2");
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
