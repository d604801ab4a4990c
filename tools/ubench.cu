// Step-0 microbenchmarks: what bounds the scan on this B200?
//   read_ldg    : read-only HBM stream with 128-bit loads (no shared memory)
//   read_tma    : read-only HBM stream through a cp.async.bulk ring (data touched minimally)
//   popc / lop3 : integer pipe issue rates
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/gsb_ubench tools/ubench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../gpusimilarity_b200/csrc/gsb_kernels.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void read_ldg_kernel(const uint4* __restrict__ p, uint64_t n_vec, uint32_t* sink)
{
    uint32_t acc = 0;
    const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n_vec; i += 4 * stride) {
        uint4 a = __ldcs(p + i), b = __ldcs(p + i + stride), c = __ldcs(p + i + 2 * stride), d = __ldcs(p + i + 3 * stride);
        acc += a.x ^ b.y ^ c.z ^ d.w;
    }
    for (; i < n_vec; i += stride) acc += __ldcs(p + i).x;
    if (acc == 0x12345678) *sink = acc;
}

__global__ void __launch_bounds__(288, 1) read_tma_kernel(const uint8_t* base, uint32_t n_tiles, uint32_t tile_bytes, uint32_t stages, uint32_t* sink)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[8], empty[8];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (uint32_t s = 0; s < stages; s++) { gsb::mbar_init(&full[s], 1); gsb::mbar_init(&empty[s], 8); }
        gsb::mbar_fence_init();
    }
    __syncthreads();
    if (warp == 8) {
        if (lane == 0) {
            uint32_t it = 0;
            for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x, it++) {
                const uint32_t s = it % stages, use = it / stages;
                if (use) gsb::mbar_wait(&empty[s], (use - 1) & 1);
                gsb::mbar_arrive_expect_tx(&full[s], tile_bytes);
                gsb::tma_bulk_g2s(smem + (size_t) s * tile_bytes, base + (uint64_t) t * tile_bytes, tile_bytes, &full[s]);
            }
        }
        return;
    }
    uint32_t acc = 0, it = 0;
    for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x, it++) {
        const uint32_t s = it % stages, use = it / stages;
        gsb::mbar_wait(&full[s], use & 1);
        acc += reinterpret_cast<const uint32_t*>(smem + (size_t) s * tile_bytes)[tid];
        __syncwarp();
        if (lane == 0) gsb::mbar_arrive(&empty[s]);
    }
    if (acc == 0x12345678) *sink = acc;
}

// Per-warp private rings (the product kernel's access pattern) with no compute.
__global__ void __launch_bounds__(512, 1) read_tma_warp_kernel(const uint8_t* base, uint32_t n_batches, uint32_t bytes,
                                                              uint32_t stride, uint32_t stages, uint32_t* sink)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[16 * 8];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, cw = blockDim.x >> 5;
    const uint32_t stage_bytes = (bytes + 127) / 128 * 128;
    uint8_t* ring = smem + (size_t) warp * stages * stage_bytes;
    uint64_t* my = full + warp * 8;
    if (lane == 0) { for (uint32_t s = 0; s < stages; s++) gsb::mbar_init(&my[s], 1); gsb::mbar_fence_init(); }
    __syncthreads();
    const uint32_t step = gridDim.x * cw;
    uint32_t b = blockIdx.x * cw + warp;
    if (lane == 0)
        for (uint32_t s = 0; s < stages; s++)
            if (b + s * step < n_batches) { gsb::mbar_arrive_expect_tx(&my[s], bytes); gsb::tma_bulk_g2s(ring + s * stage_bytes, base + (uint64_t)(b + s * step) * stride, bytes, &my[s]); }
    uint32_t acc = 0, st = 0, ph = 0;
    for (; b < n_batches; b += step) {
        gsb::mbar_wait(&my[st], ph);
        acc += reinterpret_cast<const uint32_t*>(ring + st * stage_bytes)[lane];
        __syncwarp();
        const uint32_t nb = b + stages * step;
        if (lane == 0 && nb < n_batches) { gsb::mbar_arrive_expect_tx(&my[st], bytes); gsb::tma_bulk_g2s(ring + st * stage_bytes, base + (uint64_t) nb * stride, bytes, &my[st]); }
        if (++st == stages) { st = 0; ph ^= 1; }
    }
    if (acc == 0x12345678) *sink = acc;
}

template <int MODE> __global__ void alu_kernel(uint32_t* out, int iters)
{
    uint32_t a[8];
    for (int j = 0; j < 8; j++) a[j] = threadIdx.x * 2654435761u + j;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (MODE == 0) a[j] = __popc(a[j]) + a[(j + 1) & 7];          // POPC + IADD
            else if (MODE == 1) a[j] = (a[j] & a[(j + 1) & 7]) ^ a[(j + 2) & 7]; // LOP3
            else a[j] = __popc(a[j] & a[(j + 1) & 7]) + a[(j + 2) & 7];     // LOP3+POPC+IADD
        }
    }
    uint32_t s = 0;
    for (int j = 0; j < 8; j++) s ^= a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

int main(int argc, char** argv)
{
    const size_t gib = argc > 1 ? atoi(argv[1]) : 16;
    const size_t bytes = gib << 30;
    uint8_t* buf; uint32_t* sink;
    CK(cudaMalloc(&buf, bytes)); CK(cudaMalloc(&sink, 1 << 24));
    CK(cudaMemset(buf, 1, bytes));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    printf("device %s, %d SMs, buffer %zu GiB\n", prop.name, prop.multiProcessorCount, gib);

    for (int bpsm : {4, 8, 16, 32}) {
        float best = 1e9;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            read_ldg_kernel<<<prop.multiProcessorCount * bpsm, 256>>>((const uint4*) buf, bytes / 16, sink);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            best = fminf(best, time_ms(e0, e1));
        }
        printf("read_ldg  blocks/SM=%2d           : %8.1f GB/s\n", bpsm, bytes / best / 1e6);
    }
    CK(cudaFuncSetAttribute(read_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    for (uint32_t tile : {8192u, 16384u, 32768u, 65536u}) {
        for (uint32_t stages : {2u, 3u, 4u, 6u}) {
            if ((size_t) tile * stages > 200 * 1024) continue;
            float best = 1e9;
            for (int rep = 0; rep < 4; rep++) {
                cudaEventRecord(e0);
                read_tma_kernel<<<prop.multiProcessorCount, 288, tile * stages>>>(buf, bytes / tile, tile, stages, sink);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                best = fminf(best, time_ms(e0, e1));
            }
            printf("read_tma  tile=%6u stages=%u      : %8.1f GB/s\n", tile, stages, bytes / best / 1e6);
        }
    }
    CK(cudaFuncSetAttribute(read_tma_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    for (uint32_t bytes : {4096u, 4160u, 8192u, 8320u, 16384u}) {
        for (uint32_t cw : {8u, 12u, 16u}) {
            for (uint32_t stages : {2u, 3u, 4u}) {
                const uint32_t sb = (bytes + 127) / 128 * 128;
                if ((size_t) sb * stages * cw > 220 * 1024) continue;
                float best = 1e9;
                const uint32_t nb = bytes / 4096 * 0 + (uint32_t)(((size_t) gib << 30) / bytes);
                for (int rep = 0; rep < 4; rep++) {
                    cudaEventRecord(e0);
                    read_tma_warp_kernel<<<prop.multiProcessorCount, cw * 32, sb * stages * cw>>>(buf, nb, bytes, bytes, stages, sink);
                    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                    best = fminf(best, time_ms(e0, e1));
                }
                printf("read_tma_warp bytes=%5u warps=%2u stages=%u (%3u KB in ring): %8.1f GB/s\n", bytes, cw, stages, sb * stages * cw / 1024, (double) nb * bytes / best / 1e6);
            }
        }
    }
    const int iters = 4096;
    const char* names[3] = {"popc+iadd", "lop3", "lop3+popc+iadd"};
    for (int mode = 0; mode < 3; mode++) {
        float best = 1e9;
        const int blocks = prop.multiProcessorCount * 8, threads = 256;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) alu_kernel<0><<<blocks, threads>>>(sink, iters);
            else if (mode == 1) alu_kernel<1><<<blocks, threads>>>(sink, iters);
            else alu_kernel<2><<<blocks, threads>>>(sink, iters);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            best = fminf(best, time_ms(e0, e1));
        }
        const double ops = (double) blocks * threads * iters * 8;
        printf("alu %-16s : %8.2f Gop/s (thread-ops; per SM per clk @1.9GHz: %.1f)\n", names[mode], ops / best / 1e6,
               ops / best / 1e6 / prop.multiProcessorCount / 1.9);
    }
    return 0;
}
