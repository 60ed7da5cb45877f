// pack_compact.h -- the host packer of the compact stream (24-byte public records -> 4-byte units).
// Internal to the library: arena.cu's raftgpu_step_begin_records drives it from its staging threads,
// raftgpu_pack_compact (pack_compact.cpp) from the caller's thread.
#pragma once
#include <cstdint>
#include <vector>

#include "raftgpu.h"
#include "wire_format.h"

namespace raftgpu {

// Where one packer pass writes, what it has produced so far, and the state of the run it is in.
// A pass may be continued by another call with the same PackState (the run state carries over).
// Cache-line aligned: the staging threads keep one PackState per slice in an array and update it for
// every block of records -- unaligned neighbours would share lines and ping-pong between cores.
struct alignas(128) PackState {
    // outputs
    uint32_t *units = nullptr;  // unit positions (and g_base blocks) are relative to this pointer
    uint64_t unit_cap = 0;
    uint32_t *g_base = nullptr;
    uint64_t gbase_cap = 0;
    std::vector<raftgpu_append_resp> side;  // records (and their EXT) the compact form cannot hold
    std::vector<uint32_t> esc_pos;          // unit positions of the ESC units pointing into `side`
    bool want_esc_pos = false;
    uint64_t nu = 0, n_rec = 0, blocks_set = 0;
    bool tileable = true, one_wave = true, any = false;
    uint32_t first_group = 0, last_group = 0;
    // the run being written
    bool in_run = false, header = false, have_prev_group = false;
    uint32_t cur_g = 0, run_units = 0, back = 0, prev_group = 0, seen_slots = 0;
    uint64_t base = 0;

    // back to the initial state, keeping the vectors' capacity (the step path reuses its PackStates)
    void reset() {
        std::vector<raftgpu_append_resp> sd;
        std::vector<uint32_t> ep;
        sd.swap(side);
        ep.swap(esc_pos);
        *this = PackState();
        sd.clear();
        ep.clear();
        side.swap(sd);
        esc_pos.swap(ep);
    }
};

// records[lo, hi) -> units.  `n_total` = length of the whole array (a REJECT at hi-1 looks at its EXT
// at hi; the vector path loads whole 8-record blocks and must not run past the array).
// unit_of_record: optional [n_total], unit_base + the unit position of every main record.
// RAFTGPU_OK or RAFTGPU_ERR_FULL.
int32_t pack_range(const raftgpu_append_resp *records, uint64_t lo, uint64_t hi, uint64_t n_total, PackState &st,
                   uint32_t *unit_of_record, uint32_t unit_base);

// The scalar state machine alone (the reference the vector path must match byte for byte).
int32_t pack_range_scalar(const raftgpu_append_resp *records, uint64_t lo, uint64_t hi, uint64_t n_total, PackState &st,
                          uint32_t *unit_of_record, uint32_t unit_base);

// "avx512" or "scalar": what pack_range runs on this CPU (RAFTGPU_PACK_SCALAR=1 forces scalar).
const char *pack_impl();

}  // namespace raftgpu
