// Micro-benchmark: throughput of a TMA load -> (no compute) -> TMA store ring, as a function of
// the bulk-copy size.  One persistent CTA per SM; a stage is C copies of B bytes taken from C
// different rows (row stride = whole row length, like the [slot][cap] columns of the arena).
// Usage: tma_chunks <B bytes> <C copies per stage> <stages> [total MB]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(96, 1) ring(uint8_t *base, uint64_t row_bytes, uint32_t B, uint32_t C, uint32_t n_stages, uint32_t n_tiles) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[16], done[16], empty[16];
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < n_stages; s++) { mbar_init(&full[s], 1); mbar_init(&done[s], 1); mbar_init(&empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t stage_bytes = B * C;
    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
        const uint32_t st = it % n_stages, ph = (it / n_stages) & 1u;
        uint8_t *sb = smem + static_cast<size_t>(st) * stage_bytes;
        if (warp == 0) {        // load
            if (lane == 0) { mbar_wait(&empty[st], ph ^ 1u); mbar_expect_tx(&full[st], stage_bytes); }
            __syncwarp();
            for (uint32_t j = lane; j < C; j += 32) tma_load(sb + j * B, base + j * row_bytes + static_cast<uint64_t>(tile) * B, B, &full[st]);
        } else if (warp == 1) { // "consumer"
            mbar_wait(&full[st], ph);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&done[st]);
        } else {                // store
            mbar_wait(&done[st], ph);
            for (uint32_t j = lane; j < C; j += 32) tma_store(base + j * row_bytes + static_cast<uint64_t>(tile) * B, sb + j * B, B);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
        }
    }
    if (warp == 2) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char **argv) {
    const uint32_t B = argc > 1 ? atoi(argv[1]) : 2048, C = argc > 2 ? atoi(argv[2]) : 25;
    uint32_t stages = argc > 3 ? atoi(argv[3]) : 4;
    const double total_mb = argc > 4 ? atof(argv[4]) : 200.0;
    const uint64_t n_tiles = static_cast<uint64_t>(total_mb * 1e6 / (static_cast<double>(B) * C));
    const uint64_t row_bytes = n_tiles * B;
    uint8_t *d;
    cudaMalloc(&d, row_bytes * C);
    cudaMemset(d, 1, row_bytes * C);
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t smem = static_cast<size_t>(stages) * B * C;
    if (smem > 227 * 1024 || stages > 16) { printf("B=%u C=%u stages=%u: too much smem\n", B, C, stages); return 0; }
    cudaFuncSetAttribute(ring, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    // L2 flush buffer
    uint8_t *fl; cudaMalloc(&fl, 256u << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9, sum = 0; const int reps = 10;
    for (int r = 0; r < reps + 2; r++) {
        cudaMemsetAsync(fl, r, 256u << 20);
        cudaEventRecord(e0);
        ring<<<sms, 96, smem>>>(d, row_bytes, B, C, stages, static_cast<uint32_t>(n_tiles));
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r >= 2) { best = ms < best ? ms : best; sum += ms; }
    }
    cudaError_t err = cudaGetLastError();
    const double bytes = 2.0 * row_bytes * C;
    printf("B=%6u C=%3u stages=%2u stage=%6.1fKB tiles=%8llu : avg %.1f us  best %.1f us  -> %.0f GB/s (rd+wr)  %s\n", B, C, stages, B * C / 1024.0,
           (unsigned long long)n_tiles, 1e3 * sum / reps, 1e3 * best, bytes / (sum / reps * 1e-3) / 1e9, cudaGetErrorString(err));
    return 0;
}
