// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3"): the counter-based generator
// behind every dropout mask of the library (rowops.cu: the row kernels; mha.cu: attention weights).  A mask
// is a pure function of (element, call site, step, seed), so the backward recomputes it instead of
// reading a stored mask tensor.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace msda {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

}  // namespace msda
