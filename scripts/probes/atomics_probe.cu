// Hardware-ceiling probe for the hash-grid backward / large-table forward (not part of libidrk):
//   * device-wide rate of red.global.add.v2/v4.f32 to pseudo-random rows of a table of S bytes,
//   * shared-memory accumulation (float CAS loop, native int32 ATOMS.ADD),
//   * random 8-byte / 16-byte / 32-byte gathers (ld.global.nc) from tables inside and beyond L2.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/probes/atomics_probe scripts/probes/atomics_probe.cu
// Run:    scripts/probes/atomics_probe            (prints one line per experiment)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int VEC, int UNROLL>
__global__ void red_kernel(float* table, uint32_t slot_mask, int iters) {
    uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            s = s * 1664525u + 1013904223u;
            const uint32_t slot = mix(s) & slot_mask;               // slot = VEC floats
            float* p = table + (size_t)slot * VEC;
            const float v = 1.0f;
            if (VEC == 2) asm volatile("red.global.add.v2.f32 [%0], {%1, %1};" :: "l"(p), "f"(v) : "memory");
            else if (VEC == 4) asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" :: "l"(p), "f"(v) : "memory");
            else asm volatile("red.global.add.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
        }
    }
}

// same, but each thread's UNROLL reductions land in one 32-byte sector pair neighbourhood (models the x-pairs of a cell)
template <int MODE>
__global__ void smem_kernel(float* out, int words, int iters) {
    extern __shared__ float sm[];
    for (int i = threadIdx.x; i < words; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
    for (int i = 0; i < iters; ++i) {
        s = s * 1664525u + 1013904223u;
        const uint32_t w = mix(s) % (uint32_t)words;
        if (MODE == 0) atomicAdd(sm + w, 1.0f);                                            // CAS loop
        else atomicAdd(reinterpret_cast<int*>(sm) + w, 1);                                 // native ATOMS.ADD
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = sm[0];
}

template <int BYTES, int UNROLL>
__global__ void gather_kernel(const float* __restrict__ table, uint32_t slot_mask, int iters, float* out) {
    uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
    float acc = 0.f;
    for (int i = 0; i < iters; ++i) {
        float v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            s = s * 1664525u + 1013904223u;
            const uint32_t slot = mix(s) & slot_mask;
            if (BYTES == 8) { const float2 t = __ldg(reinterpret_cast<const float2*>(table) + slot); v[u] = t.x + t.y; }
            else if (BYTES == 16) { const float4 t = __ldg(reinterpret_cast<const float4*>(table) + slot); v[u] = t.x + t.w; }
            else { const float4 a = __ldg(reinterpret_cast<const float4*>(table) + 2 * (size_t)slot);
                   const float4 b = __ldg(reinterpret_cast<const float4*>(table) + 2 * (size_t)slot + 1); v[u] = a.x + b.w; }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc += v[u];
    }
    if (acc == 123.456f) out[0] = acc;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const size_t max_bytes = (size_t)1 << 30;
    float* table; CK(cudaMalloc(&table, max_bytes)); CK(cudaMemset(table, 0, max_bytes));
    float* out; CK(cudaMalloc(&out, 1 << 20));
    const int threads = 256;
    printf("sms %d\n", sms);
    const size_t sizes[] = {(size_t)32 << 10, (size_t)128 << 10, (size_t)512 << 10, (size_t)2 << 20, (size_t)4 << 20, (size_t)32 << 20, (size_t)48 << 20, (size_t)96 << 20, (size_t)256 << 20, (size_t)1 << 30};
    for (int ctas_per_sm : {8}) {
        for (size_t S : sizes) {
            const int grid = sms * ctas_per_sm, iters = 256;
            const double ops = (double)grid * threads * iters * 8;
#define RUN_RED(VEC) { \
            uint32_t mask = (uint32_t)(S / (4 * VEC) - 1); \
            red_kernel<VEC, 8><<<grid, threads>>>(table, mask, 16); CK(cudaDeviceSynchronize()); \
            CK(cudaEventRecord(e0)); red_kernel<VEC, 8><<<grid, threads>>>(table, mask, iters); CK(cudaEventRecord(e1)); \
            CK(cudaDeviceSynchronize()); const float ms = time_ms(e0, e1); \
            printf("red.v%d  table %7zu KB  ctas/sm %d : %7.1f Gops/s  (%.3f cyc/op/SM @1.9GHz)  %7.1f GB/s payload\n", VEC, S >> 10, ctas_per_sm, \
                   ops / ms / 1e6, 1.9e9 * sms / (ops / (ms * 1e-3)), ops * 4 * VEC / ms / 1e6); }
            RUN_RED(1) RUN_RED(2) RUN_RED(4)
        }
    }
    for (size_t S : sizes) {
        for (int ctas_per_sm : {8}) {
            const int grid = sms * ctas_per_sm, iters = 128;
            const double ops = (double)grid * threads * iters * 8;
#define RUN_G(BYTES) { \
            uint32_t mask = (uint32_t)(S / BYTES - 1); \
            gather_kernel<BYTES, 8><<<grid, threads>>>(table, mask, 8, out); CK(cudaDeviceSynchronize()); \
            CK(cudaEventRecord(e0)); gather_kernel<BYTES, 8><<<grid, threads>>>(table, mask, iters, out); CK(cudaEventRecord(e1)); \
            CK(cudaDeviceSynchronize()); const float ms = time_ms(e0, e1); \
            printf("gather %2dB table %5zu MB  ctas/sm %d : %7.1f Gops/s  %7.1f GB/s useful  %7.1f GB/s in 32B sectors\n", BYTES, S >> 20, ctas_per_sm, \
                   ops / ms / 1e6, ops * BYTES / ms / 1e6, ops * 32 / ms / 1e6); }
            RUN_G(8) RUN_G(16) RUN_G(32)
        }
    }
    for (int words : {8192, 32768}) {
        const int grid = sms * 2, iters = 4096;
        const double ops = (double)grid * threads * iters;
        CK(cudaFuncSetAttribute(smem_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, words * 4));
        CK(cudaFuncSetAttribute(smem_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, words * 4));
        CK(cudaEventRecord(e0)); smem_kernel<0><<<grid, threads, words * 4>>>(out, words, iters); CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = time_ms(e0, e1);
        printf("smem float atomicAdd (CAS)  %6d words: %7.1f Gops/s (%.2f cyc/op/SM)\n", words, ops / ms / 1e6, 1.9e9 * sms / (ops / (ms * 1e-3)));
        CK(cudaEventRecord(e0)); smem_kernel<1><<<grid, threads, words * 4>>>(out, words, iters); CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        ms = time_ms(e0, e1);
        printf("smem int32 atomicAdd (ATOMS.ADD) %6d words: %7.1f Gops/s (%.2f cyc/op/SM)\n", words, ops / ms / 1e6, 1.9e9 * sms / (ops / (ms * 1e-3)));
    }
    return 0;
}
