// Stable LSD radix sort of (uint32 key, uint32 value) pairs (csrc/point_sort.cu), shared by the Morton pre-pass and the
// deterministic table-gradient pass.
#pragma once
#include "common.cuh"

namespace idrk {

long long radix_sort_scratch_bytes(long long n);
int radix_sort_pairs(uint32_t* keys[2], uint32_t* vals[2], uint32_t* vals_final, long long n, int bits, void* scratch,
                     cudaStream_t st, int* keys_in);

}  // namespace idrk
