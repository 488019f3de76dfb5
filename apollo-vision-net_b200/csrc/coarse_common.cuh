// Shared between the fused SCA backward (fused.cu: producer of the coarse-level sample records) and
// the tensor-core scatter of those records (coarse_scatter.cu).
//
// "Coarse patch" = the last one or two pyramid levels of a value map, as long as they hold at most
// kCoarseMaxPx pixels together (BEVFormer-base: 29x50 + 15x25 = 1825 px; the tiny config's single
// 28x48 level = 1344 px).  Those levels have few pixels and many samples (157 / 500 grad_value
// updates per slot at the base config), so instead of one L2 reduction per corner the backward
// writes one 16-byte record per sample and a second kernel accumulates the whole patch of one
// (camera, head) in TENSOR MEMORY with tcgen05.mma: grad_patch[pixel, channel] += W^T[pixel, row] *
// g_out[row, channel], W being the sparse matrix of (attention x bilinear / count) weights.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <type_traits>
#include <stdint.h>

namespace msda {

constexpr int kCoarseMaxPx = 2048;      // 16 M-tiles of 128 pixels x 32 fp32 columns = the 512 TMEM columns
constexpr int kCoarseMaxLevels = 2;
constexpr int kCoarseMaxP = 8;          // points per level the record path supports
constexpr int kCoarseDh = 32;           // head dim (= N of the MMA, TMEM columns per M-tile)

struct CoarsePatch {
  int first_level;                      // == L: no coarse patch for these level tables
  int npix;                             // pixels of levels [first_level, L)
  int start;                            // pixel index of the patch inside a value map
};

// Evaluated on the device from the int64 level tables (no host synchronisation anywhere): the longest
// suffix of at most kCoarseMaxLevels levels that are contiguous in the map and fit kCoarseMaxPx.
__device__ __forceinline__ CoarsePatch coarse_patch(const int64_t* __restrict__ shapes,
                                                    const int64_t* __restrict__ starts, int L) {
  CoarsePatch c{L, 0, 0};
  long long next_start = -1;
  for (int l = L - 1; l >= 0 && L - l <= kCoarseMaxLevels; --l) {
    const long long n = shapes[2 * l] * shapes[2 * l + 1];
    if (n <= 0 || c.npix + n > kCoarseMaxPx) break;
    if (next_start >= 0 && starts[l] + n != next_start) break;     // not contiguous with the level above
    c.first_level = l;
    c.npix += (int)n;
    c.start = (int)starts[l];
    next_start = starts[l];
  }
  return c;
}

// Record of one coarse sample of one (batch, camera, query, head) row: four corners, each
// (position of the patch pixel : 16 | 16-bit weight : 16), weight 0 for corners outside the map.  The
// position is the pixel's byte offset inside one k-group of the MMA's A operand (MN-major core
// matrices of 8 pixels x 8 rows): (pixel / 8) * 128 + (pixel % 8) * 2.
// Layout: rec[((((b * cams + cam) * Nq + q) * M + m) * (kCoarseMaxLevels * P)) + (s - first_level * P)].
// The weight dtype W is fp16 with the fp16 accumulator (and for fp16 models), bf16 for a bf16 model with
// the fp32 accumulator (see coarse_scatter.cu).
template <typename W>
__device__ __forceinline__ uint32_t coarse_corner(int patch_pixel, float weight) {
  const uint32_t pos = ((uint32_t)patch_pixel >> 3) * 128u + ((uint32_t)patch_pixel & 7u) * 2u;
  uint16_t bits;
  if constexpr (sizeof(W) == 2 && std::is_same<W, __half>::value) {
    const __half h = __float2half_rn(weight);
    bits = *reinterpret_cast<const uint16_t*>(&h);
  } else {
    const __nv_bfloat16 h = __float2bfloat16_rn(weight);
    bits = *reinterpret_cast<const uint16_t*>(&h);
  }
  return ((uint32_t)bits << 16) | pos;
}

}  // namespace msda
