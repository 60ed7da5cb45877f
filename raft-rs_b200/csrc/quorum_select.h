// quorum_select.h -- the joint quorum index of an 8-slot group from ONE comparison pass (host + device: the host
// build is what tests/test_quorum_select.py checks against a sort).
//
// MajorityConfig::committed_index without group commit (majority.rs:70-101) is the q-th largest acked index of the
// voters in a mask, q = n/2 + 1 (util.rs:118-120); JointConfig::committed_index (joint.rs:47-51) wants it for two
// masks over the same values.  Sorting twice (two 19-comparator networks on u64 = ~310 instructions per group) is
// what made the general instantiation of the step kernels compute-bound.  Instead: order the 8 slots once -- slot i
// PRECEDES slot j when v[i] > v[j], or they are equal and i < j (a strict total order; equal values select the same
// index whichever comes first, as with the reference's stable sort, majority.rs:95) -- as one bit mask of
// predecessors per slot (28 compares), and the q-th largest member of a mask M is then the member with exactly q - 1
// predecessors inside M.  Non-members need no zeroing and an empty mask yields u64::MAX (majority.rs:71-75).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RAFTGPU_QS_FN __host__ __device__ __forceinline__
#else
#define RAFTGPU_QS_FN inline
#endif
#if defined(__CUDA_ARCH__)
#define RAFTGPU_QS_UNROLL _Pragma("unroll")
#else
#define RAFTGPU_QS_UNROLL
#endif

namespace raftgpu {

RAFTGPU_QS_FN uint32_t qs_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return static_cast<uint32_t>(__popc(x));
#else
    return static_cast<uint32_t>(__builtin_popcount(x));
#endif
}

RAFTGPU_QS_FN void quorum_index_joint(const uint64_t (&v)[8], uint32_t in, uint32_t out, uint64_t &i_idx, uint64_t &o_idx) {
    uint32_t prec[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
RAFTGPU_QS_UNROLL
    for (int i = 0; i < 8; i++) {
RAFTGPU_QS_UNROLL
        for (int j = i + 1; j < 8; j++) {
            const bool i_first = v[i] >= v[j];  // ties: the lower slot first
            prec[j] |= i_first ? (1u << i) : 0u;
            prec[i] |= i_first ? 0u : (1u << j);
        }
    }
    i_idx = UINT64_MAX;
    o_idx = UINT64_MAX;
    const uint32_t before_in = qs_popc(in & 0xffu) >> 1, before_out = qs_popc(out & 0xffu) >> 1;  // q - 1 = n / 2
RAFTGPU_QS_UNROLL
    for (int i = 0; i < 8; i++) {
        if (((in >> i) & 1u) && qs_popc(prec[i] & in) == before_in) i_idx = v[i];
        if (((out >> i) & 1u) && qs_popc(prec[i] & out) == before_out) o_idx = v[i];
    }
}

}  // namespace raftgpu
