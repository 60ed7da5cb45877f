// wire_format.h -- bit layouts of the two record wire forms, shared by the host packers (plain
// C++, compiled by g++) and the device decoders (kernels.cuh).  Layout prose: include/raftgpu.h
// ("compact stream") and DESIGN.md 2.
#pragma once
#include <cstdint>

namespace raftgpu {

// Packed staging record, 16 bytes: what raftgpu_enqueue_* writes into the pinned rings and the
// general step path ships over PCIe.
//   w0: [0,32) group  [32,35) slot  35 REJECT  36 LOCAL  37 EXT  38 WIDE  39 HAS_EXT  [40,64) delta
//   w1: m.index   (EXT: the payload)
// commit is carried as a 24-bit delta: index - commit for a message (a follower's commit never
// exceeds what it acknowledges), commit - index for a LOCAL record (0xFFFFFF = "no new
// last_index"); anything else sets WIDE and the exact value follows in an EXT record.
// EXT kinds (in the delta field): 1 = next_probe_index, 2 = request_snapshot, 3 = wide commit,
// 0 = padding.
struct PackedRec {
    uint64_t w0, w1;
};
constexpr uint64_t kPkReject = 1ull << 35, kPkLocal = 1ull << 36, kPkExt = 1ull << 37, kPkWide = 1ull << 38,
                   kPkHasExt = 1ull << 39;
constexpr uint32_t kPkNoCommit = 0xFFFFFFu;

// The compact stream: 4-byte units, group runs with a two-unit header.
constexpr uint32_t kCuRec = 0, kCuHdrA = 1, kCuHdrB = 2, kCuEsc = 3;
constexpr uint32_t kCuLocal = 4u, kCuReject = 1u << 9, kCuNoCommit = 255u;
constexpr uint32_t kCuPayload = 1u << 29;   // in the ESC field: a REJECT's hint rides here, not a side index
constexpr uint32_t kCuPad = 0x1fffffffu;    // ESC field value of a padding unit (side indexes stay below it)

}  // namespace raftgpu
