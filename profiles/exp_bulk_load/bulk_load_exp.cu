// Experiment for north_star stage 3 / N1 ("TMA bulk loads where record spans are long enough to pay off"): does the
// multi-key shape (C4: 66 KB lines, 25-80 byte samples) gain from feeding the warps through cp.async.bulk (global ->
// shared, mbarrier completion) instead of 128-bit loads prefetched three windows ahead in registers?  NOT part of the
// product: two stand-alone kernels with the same per-window work — tab / newline / '0' / '1' byte masks of a 512-byte
// window and three tallies (tabs, tabs followed by '0', tabs followed by '1'), about the integer work of the product's
// skip-ahead loop — that differ only in where a window's 16 bytes per lane come from.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

namespace {

constexpr uint32_t WINDOW = 512, TILE = 64u << 10, WARPS = 8, FULL = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t zero_bytes(uint32_t x) { return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u; }
__device__ __forceinline__ uint32_t eq_bytes(uint32_t w, uint32_t c4) { return zero_bytes(w ^ c4); }

struct Tally { unsigned long long tabs, t0, t1, nl; };

// one window: 16 bytes of this lane, the byte mask work, the tallies (prev_t: 0x80 when the byte in front of the lane's
// 16 bytes was a tab)
__device__ __forceinline__ void window(const uint4 v, uint32_t &carry, Tally &T) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t t[4], z[4], o[4], n = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) { t[j] = eq_bytes(w[j], 0x09090909u); z[j] = eq_bytes(w[j], 0x30303030u); o[j] = eq_bytes(w[j], 0x31313131u); n |= eq_bytes(w[j], 0x0A0A0A0Au); }
    // the tab mask moved up by one byte (the last byte of the lane before comes in through a shuffle)
    const uint32_t last = __shfl_up_sync(FULL, t[3], 1);
    uint32_t prev = (threadIdx.x & 31) ? last : carry;
    carry = __shfl_sync(FULL, t[3], 31);
    uint32_t c_t = 0, c_0 = 0, c_1 = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t up = __funnelshift_l(prev, t[j], 8);
        c_t += __popc(t[j]); c_0 += __popc(up & z[j]); c_1 += __popc(up & o[j]);
        prev = t[j];
    }
    T.tabs += c_t; T.t0 += c_0; T.t1 += c_1; T.nl += __popc(n);
}

__device__ __forceinline__ void finish(Tally T, unsigned long long *out) {
    for (int o = 16; o; o >>= 1) {
        T.tabs += __shfl_down_sync(FULL, T.tabs, o); T.t0 += __shfl_down_sync(FULL, T.t0, o);
        T.t1 += __shfl_down_sync(FULL, T.t1, o); T.nl += __shfl_down_sync(FULL, T.nl, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, T.tabs); atomicAdd(out + 1, T.t0); atomicAdd(out + 2, T.t1); atomicAdd(out + 3, T.nl); }
}

// A: the product's way — 128-bit loads, three windows ahead in registers
__global__ void __launch_bounds__(WARPS * 32, 3)
scan_regs(const uint8_t *__restrict__ in, uint32_t n_tiles, unsigned int *ticket, unsigned long long *out) {
    const int lane = threadIdx.x & 31;
    Tally T = {0, 0, 0, 0};
    for (;;) {
        unsigned int tile = 0;
        if (lane == 0) tile = atomicAdd(ticket, 1u);
        tile = __shfl_sync(FULL, tile, 0);
        if (tile >= n_tiles) break;
        const uint4 *p = reinterpret_cast<const uint4 *>(in + (size_t)tile * TILE) + lane;
        uint32_t carry = 0;
        uint4 cur = __ldg(p), nxt = __ldg(p + 32), nx2 = __ldg(p + 64);
        for (uint32_t wv = 0; wv < TILE / WINDOW; ++wv) {
            const uint4 v = cur;
            cur = nxt; nxt = nx2;
            if (wv + 3 < TILE / WINDOW) nx2 = __ldg(p + 32 * (wv + 3));
            window(v, carry, T);
        }
    }
    finish(T, out);
}

// B: the same work fed by the bulk-copy engine — per warp a ring of STAGES pieces of PIECE bytes in shared memory, each
// filled by ONE cp.async.bulk (global -> shared) that completes on the piece's mbarrier
template <int STAGES, uint32_t PIECE>
__global__ void __launch_bounds__(WARPS * 32, 3)
scan_bulk(const uint8_t *__restrict__ in, uint32_t n_tiles, unsigned int *ticket, unsigned long long *out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long bars[WARPS][STAGES];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint8_t *ring = smem + (size_t)wid * STAGES * PIECE;
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(&bars[wid][0]);
    if (lane == 0) {
        for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_s + 8u * s));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t phase = 0;                       // bit s: the parity stage s is waited on next
    Tally T = {0, 0, 0, 0};
    constexpr uint32_t PIECES = TILE / PIECE;
    for (;;) {
        unsigned int tile = 0;
        if (lane == 0) tile = atomicAdd(ticket, 1u);
        tile = __shfl_sync(FULL, tile, 0);
        if (tile >= n_tiles) break;
        const uint8_t *src = in + (size_t)tile * TILE;
        auto issue = [&](uint32_t piece) {                           // lane 0 only
            const uint32_t s = piece % STAGES;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_s + 8u * s), "r"(PIECE) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(ring_s + s * PIECE), "l"(src + (size_t)piece * PIECE), "r"(PIECE), "r"(bar_s + 8u * s) : "memory");
        };
        if (lane == 0) for (uint32_t q = 0; q < STAGES && q < PIECES; ++q) issue(q);
        uint32_t carry = 0;
        for (uint32_t piece = 0; piece < PIECES; ++piece) {
            const uint32_t s = piece % STAGES;
            uint32_t done = 0;
            while (!done) {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(bar_s + 8u * s), "r"((phase >> s) & 1u) : "memory");
            }
            phase ^= 1u << s;
            const uint4 *sp = reinterpret_cast<const uint4 *>(ring + s * PIECE) + lane;
#pragma unroll
            for (uint32_t wv = 0; wv < PIECE / WINDOW; ++wv) window(sp[32 * wv], carry, T);
            __syncwarp();                                            // every lane has read the piece: it may be filled again
            if (lane == 0 && piece + STAGES < PIECES) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(piece + STAGES);
            }
        }
    }
    finish(T, out);
}

}  // namespace

extern "C" int bulk_exp_run(int variant, const void *d_in, size_t nbytes, int reps, float *ms_out, unsigned long long *tallies) {
    const uint32_t n_tiles = (uint32_t)(nbytes / TILE);
    unsigned int *ticket = nullptr; unsigned long long *out = nullptr;
    cudaMalloc(&ticket, 4); cudaMalloc(&out, 32);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    auto launch = [&](void) {
        const uint8_t *in = static_cast<const uint8_t *>(d_in);
        const int grid = sms * 3;
        switch (variant) {
        case 0: scan_regs<<<grid, WARPS * 32>>>(in, n_tiles, ticket, out); break;
        case 1: cudaFuncSetAttribute(scan_bulk<2, 2048>, cudaFuncAttributeMaxDynamicSharedMemorySize, WARPS * 2 * 2048);
                scan_bulk<2, 2048><<<grid, WARPS * 32, WARPS * 2 * 2048>>>(in, n_tiles, ticket, out); break;
        case 2: cudaFuncSetAttribute(scan_bulk<4, 2048>, cudaFuncAttributeMaxDynamicSharedMemorySize, WARPS * 4 * 2048);
                scan_bulk<4, 2048><<<grid, WARPS * 32, WARPS * 4 * 2048>>>(in, n_tiles, ticket, out); break;
        case 3: cudaFuncSetAttribute(scan_bulk<2, 4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, WARPS * 2 * 4096);
                scan_bulk<2, 4096><<<grid, WARPS * 32, WARPS * 2 * 4096>>>(in, n_tiles, ticket, out); break;
        default: break;
        }
    };
    for (int r = 0; r < reps + 2; ++r) {
        cudaMemset(ticket, 0, 4); cudaMemset(out, 0, 32);
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { fprintf(stderr, "bulk_exp: %s\n", cudaGetErrorString(cudaGetLastError())); return -1; }
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (r >= 2 && ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { fprintf(stderr, "bulk_exp: %s\n", cudaGetErrorString(err)); return -1; }
    cudaMemcpy(tallies, out, 32, cudaMemcpyDeviceToHost);
    *ms_out = best;
    cudaFree(ticket); cudaFree(out); cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}
