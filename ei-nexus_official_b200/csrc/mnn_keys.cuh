// Packed (similarity, index) keys shared by every MNN precision path.
//   key = orderable_bits(sim) << 32 | (0xffffffff - index)
// so an unsigned 64-bit max is "largest similarity, lowest index on ties" == topk(1).  0 = empty.
#pragma once
#include "common.cuh"

__device__ __forceinline__ unsigned int key_index(unsigned long long k) { return 0xffffffffu - (unsigned int)k; }
__device__ __forceinline__ float key_value(unsigned long long k) { return f32_from_orderable((unsigned int)(k >> 32)); }
