// tmem_drain.cu -- microbenchmark: how fast can the epilogue warps of one CTA read a 128 x 256 32-bit accumulator
// out of TMEM with tcgen05.ld?  This is the floor DESIGN.md quotes for K2's short-row configs (config E batched:
// 512 MMA cycles per tile against the drain of the tile).  One CTA per SM, 512 TMEM columns allocated (the two
// accumulators K2 double-buffers), G epilogue warpgroups each reading 256 / G columns of every tile in chunks of
// X columns (tcgen05.ld.32x32b.xX), R tiles per run.  Reports SM cycles per tile, with and without the 3-input
// min tree K2 runs on the values.  No MMA is issued: TMEM contents are whatever the allocation holds.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_drain tmem_drain.cu && ./tmem_drain
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s\n", cudaGetErrorString(e), #x); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X> struct Ld;
template <> struct Ld<32> {
    static __device__ __forceinline__ void go(uint32_t taddr, uint32_t* v) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                     "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                       "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                       "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(taddr) : "memory");
    }
};
template <> struct Ld<64> {
    static __device__ __forceinline__ void go(uint32_t taddr, uint32_t* v) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 "
                     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                     "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
                     "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
                     "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                       "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                       "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]),
                       "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]),
                       "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]),
                       "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]),
                       "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
                     : "r"(taddr) : "memory");
    }
};
template <> struct Ld<16> {
    static __device__ __forceinline__ void go(uint32_t taddr, uint32_t* v) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr) : "memory");
    }
};

// G warpgroups, X columns per tcgen05.ld, PIPE loads in flight before the wait, MINTREE: run K2's integer min tree
template <int G, int X, int PIPE, bool MINTREE>
__global__ void __launch_bounds__(128 * G, 1) drain(int tiles, long long* cycles_out, uint32_t* sink) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_slot;
    const int quarter = warp & 3, grp = warp >> 2;
    constexpr int COLS = 256 / G;                        // columns of a tile this warpgroup reads
    static_assert(COLS % (X * PIPE) == 0, "slice must be a whole number of load groups");
    uint32_t acc = 0;
    int ext = 0x7fffffff;
    __syncthreads();
    const long long t0 = clock64();
    for (int t = 0; t < tiles; t++) {
        const uint32_t taddr = base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((t & 1) * 256 + grp * COLS);
#pragma unroll
        for (int c = 0; c < COLS; c += X * PIPE) {
            uint32_t v[PIPE][X];
#pragma unroll
            for (int p = 0; p < PIPE; p++) Ld<X>::go(taddr + c + p * X, v[p]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int p = 0; p < PIPE; p++) {
                if (MINTREE) {
#pragma unroll
                    for (int j = 0; j + 1 < X; j += 2) ext = min(ext, min((int)v[p][j], (int)v[p][j + 1]));   // VIMNMX3: one per two elements
                } else {
                    acc ^= v[p][0] ^ v[p][X - 1];        // keeps the load alive at ~no ALU cost
                }
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles_out[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u || ext == 0x12345678) sink[0] = acc + (uint32_t)ext + lane;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512) : "memory");
}

template <int G, int X, int PIPE, bool MINTREE>
void run(const char* name, int sms, long long* d_cycles, uint32_t* d_sink) {
    const int tiles = 4000;
    drain<G, X, PIPE, MINTREE><<<sms, 128 * G>>>(200, d_cycles, d_sink);          // warm-up
    CHECK(cudaDeviceSynchronize());
    drain<G, X, PIPE, MINTREE><<<sms, 128 * G>>>(tiles, d_cycles, d_sink);
    CHECK(cudaDeviceSynchronize());
    long long* h = (long long*)malloc(sizeof(long long) * sms);
    CHECK(cudaMemcpy(h, d_cycles, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    double sum = 0, mx = 0;
    for (int i = 0; i < sms; i++) { sum += (double)h[i]; if ((double)h[i] > mx) mx = (double)h[i]; }
    const double per_tile = sum / sms / tiles;
    printf("%-44s %8.1f cycles / 128x256 tile  (max SM %8.1f)  %6.1f B/clk/SM\n", name, per_tile, mx / tiles, 128.0 * 256 * 4 / per_tile);
    free(h);
}

int main() {
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    long long* d_cycles; uint32_t* d_sink;
    CHECK(cudaMalloc(&d_cycles, sizeof(long long) * sms));
    CHECK(cudaMalloc(&d_sink, 64));
    printf("device %s, %d SMs; every CTA drains 128 lanes x 256 columns x 4 B = 128 KiB per tile\n", prop.name, sms);
    run<1, 32, 1, false>("1 warpgroup,  x32, wait per load", sms, d_cycles, d_sink);
    run<2, 32, 1, false>("2 warpgroups, x32, wait per load", sms, d_cycles, d_sink);
    run<4, 32, 1, false>("4 warpgroups, x32, wait per load", sms, d_cycles, d_sink);
    run<2, 32, 2, false>("2 warpgroups, x32, 2 loads per wait", sms, d_cycles, d_sink);
    run<4, 32, 2, false>("4 warpgroups, x32, 2 loads per wait", sms, d_cycles, d_sink);
    run<2, 64, 1, false>("2 warpgroups, x64, wait per load", sms, d_cycles, d_sink);
    run<4, 64, 1, false>("4 warpgroups, x64, wait per load", sms, d_cycles, d_sink);
    run<2, 64, 2, false>("2 warpgroups, x64, 2 loads per wait", sms, d_cycles, d_sink);
    run<4, 16, 4, false>("4 warpgroups, x16, 4 loads per wait", sms, d_cycles, d_sink);
    run<2, 32, 1, true>("2 warpgroups, x32, + min tree", sms, d_cycles, d_sink);
    run<4, 32, 1, true>("4 warpgroups, x32, + min tree", sms, d_cycles, d_sink);
    run<4, 32, 2, true>("4 warpgroups, x32, 2 loads per wait, + min tree", sms, d_cycles, d_sink);
    run<4, 64, 1, true>("4 warpgroups, x64, + min tree", sms, d_cycles, d_sink);
    return 0;
}
