// tmem_cp_seed.cu -- microbenchmark: can tcgen05.cp (shared memory -> TMEM, issued by ONE thread) replace the
// epilogue warps' tcgen05.st when every lane of an accumulator column gets the same value (K2's integer L2 seeds,
// tc_kernels.cuh INIT_ACC)?  Two questions:
//   (1) semantics: shape .128x256b reads 128 rows x 32 bytes through a matrix descriptor; with the no-swizzle K-major
//       layout a row r, 16-byte chunk c lives at start + (r % 8) * 16 + (r / 8) * SBO + c * LBO.  SBO = 0 would make all
//       sixteen 8-row groups read the SAME 8 rows, so a 128-byte block (8 copies of 4 seeds) broadcasts 4 columns to all
//       128 lanes.  Mode 1 checks the layout with a full 128-row image (SBO = 256), mode 0 the SBO = 0 aliasing.
//   (2) cost: cycles for the 32 copies that seed one 128 x 256 accumulator (issue -> commit -> mbarrier).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_cp_seed tmem_cp_seed.cu && ./tmem_cp_seed
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s\n", cudaGetErrorString(e), #x); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                   "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                   "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// no-swizzle K-major descriptor: start >> 4 | LBO >> 4 at [16,30) | SBO >> 4 at [32,46) | version 1 at [46,48)
__device__ __forceinline__ uint64_t sdesc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t seed_of(uint32_t col) { return col * 2654435761u + 12345u; }

// mode 0: 8 KiB image, SBO = 0 (every 8-row group aliases the same 8 rows), all 256 columns in 32 copies
// mode 1: one copy (8 columns) from a full 128-row image, SBO = 256
__global__ void __launch_bounds__(128, 1) cp_seed(int mode, int reps, long long* cycles_out, uint32_t* mism_out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint4* img = reinterpret_cast<uint4*>(smem);
    if (mode == 0) {
        // chunk ch (4 columns) -> 128-byte block: 8 identical 16-byte rows
        for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {
            const uint32_t ch = i >> 3;
            img[i] = make_uint4(seed_of(4 * ch), seed_of(4 * ch + 1), seed_of(4 * ch + 2), seed_of(4 * ch + 3));
        }
    } else {
        // rows 0..127, chunks 0..1: group g = r / 8 at g * 256, chunk c at c * 128, row at (r % 8) * 16; lane-dependent values
        for (int i = threadIdx.x; i < 128 * 2; i += blockDim.x) {
            const uint32_t r = i >> 1, c = i & 1;
            img[(r >> 3) * 16 + c * 8 + (r & 7)] = make_uint4(seed_of(4 * c) + r, seed_of(4 * c + 1) + r, seed_of(4 * c + 2) + r, seed_of(4 * c + 3) + r);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_slot;
    long long total = 0;
    if (threadIdx.x == 0) {
        uint32_t phase = 0;
        for (int rep = 0; rep < reps; rep++) {
            const long long t0 = clock64();
            const uint32_t d_tmem = base + (uint32_t)((rep & 1) * 256);
            const int ncp = mode == 0 ? 32 : 1;
            for (int c = 0; c < ncp; c++) {
                const uint64_t desc = mode == 0 ? sdesc(smem_u32(smem) + (uint32_t)c * 256u, 128u, 0u) : sdesc(smem_u32(smem), 128u, 256u);
                asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(d_tmem + (uint32_t)c * 8u), "l"(desc) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            asm volatile(
                "{\n\t.reg .pred p;\n\tW_%=:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                "@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(&bar)), "r"(phase) : "memory");
            phase ^= 1;
            total += clock64() - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // verify the accumulator written last
    const uint32_t d_tmem = base + (uint32_t)(((reps - 1) & 1) * 256);
    uint32_t bad = 0;
    const int ncols = mode == 0 ? 256 : 8;
    for (int c0 = 0; c0 < (mode == 0 ? 256 : 32); c0 += 32) {
        uint32_t v[32];
        ld32(d_tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 32; j++) {
            if (c0 + j >= ncols) continue;
            const uint32_t want = mode == 0 ? seed_of(c0 + j) : seed_of(c0 + j) + (uint32_t)(warp * 32 + lane);
            if (v[j] != want) bad++;
        }
    }
    atomicAdd(mism_out, bad);
    if (threadIdx.x == 0) cycles_out[blockIdx.x] = total;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512) : "memory");
}

int main() {
    int dev = 0, sms = 0;
    CHECK(cudaGetDevice(&dev));
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long* d_cycles; uint32_t* d_mism;
    CHECK(cudaMalloc(&d_cycles, sizeof(long long) * sms));
    CHECK(cudaMalloc(&d_mism, 4));
    CHECK(cudaFuncSetAttribute(cp_seed, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
    for (int mode = 1; mode >= 0; mode--) {
        for (int grid : {1, sms}) {
            const int reps = 200;
            CHECK(cudaMemset(d_mism, 0, 4));
            cp_seed<<<grid, 128, 16384>>>(mode, reps, d_cycles, d_mism);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d grid %d: launch failed: %s\n", mode, grid, cudaGetErrorString(e)); return 1; }
            long long h[1024]; uint32_t mism = 0;
            CHECK(cudaMemcpy(h, d_cycles, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
            CHECK(cudaMemcpy(&mism, d_mism, 4, cudaMemcpyDeviceToHost));
            long long mx = 0, sum = 0;
            for (int i = 0; i < grid; i++) { sum += h[i]; if (h[i] > mx) mx = h[i]; }
            printf("mode %d (%s) grid %3d: mismatching (lane, column) values %u of %d; %.1f cycles per %s (max CTA %.1f), issue + commit + wait\n",
                   mode, mode == 0 ? "SBO = 0 broadcast image, 32 copies = 256 columns" : "full 128-row image, 1 copy = 8 columns", grid, mism,
                   grid * 128 * (mode == 0 ? 256 : 8), (double)sum / grid / reps, mode == 0 ? "128x256 accumulator" : "copy", (double)mx / reps);
        }
    }
    return 0;
}
