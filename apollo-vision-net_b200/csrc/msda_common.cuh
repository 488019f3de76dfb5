// Shared device helpers for the sm_100a deformable-attention kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/msda_b200.h"

namespace msda {

constexpr int kMaxLevels = MSDA_MAX_LEVELS;

// Per-level table staged once per CTA in shared memory (h, w, start as int32).
struct LevelTable {
  int h[kMaxLevels];
  int w[kMaxLevels];
  int start[kMaxLevels];
};

__device__ __forceinline__ void load_level_table(LevelTable& t, const int64_t* __restrict__ shapes,
                                                 const int64_t* __restrict__ starts, int L) {
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    t.h[l] = (int)shapes[2 * l];
    t.w[l] = (int)shapes[2 * l + 1];
    t.start[l] = (int)starts[l];
  }
}

// ---- element type traits: 16-byte vectors of the value dtype ---------------------------
template <typename T> struct Vec16;            // VEC elements in one 128-bit load

template <> struct Vec16<float> {
  static constexpr int N = 4;
  __device__ __forceinline__ static void unpack(const uint4& u, float (&f)[4]) {
    f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y);
    f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
  }
  __device__ __forceinline__ static uint4 pack(const float (&f)[4]) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]),
                      __float_as_uint(f[2]), __float_as_uint(f[3]));
  }
};

template <> struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  // bf16 -> fp32 is a 16-bit shift: lo = w << 16, hi = w & 0xffff0000
  __device__ __forceinline__ static void unpack(const uint4& u, float (&f)[8]) {
    f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
    f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
    f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
    f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
  }
  __device__ __forceinline__ static uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __device__ __forceinline__ static uint4 pack(const float (&f)[8]) {
    return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
  }
};

template <> struct Vec16<__half> {
  static constexpr int N = 8;
  __device__ __forceinline__ static void unpack(const uint4& u, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      float2 p = __half22float2(h);
      f[2 * i] = p.x; f[2 * i + 1] = p.y;
    }
  }
  __device__ __forceinline__ static uint32_t pack2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __device__ __forceinline__ static uint4 pack(const float (&f)[8]) {
    return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
  }
};

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 128-bit read-only global load (value maps are re-read many times: keep them in L1/L2).
__device__ __forceinline__ uint4 ldg128(const void* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}

// Vector reduction: one 16-byte fp32x4 add at L2 (sm_90+), no return value.
__device__ __forceinline__ void red_add_f32x4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_f32x2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" :: "l"(addr), "f"(a), "f"(b) : "memory");
}

// ---- mbarrier + 1-D bulk copy (TMA engine; SASS: UBLKCP) --------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
               :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" :: "r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- bilinear sample set-up ---------------------------------------------------------------
// Pixel convention of the operator: x_pix = loc_x * W - 0.5 (align_corners = False); a sample
// takes part only when -1 < x_pix < W and -1 < y_pix < H, and a corner only when it is inside
// the map (zero padding) -- SURVEY.md section 8a quirk (10).
struct Bilinear {
  int x0, y0;          // top-left corner (may be -1)
  float lw, lh;        // fractional parts
  bool in_range;       // sample contributes at all
  bool vx0, vx1, vy0, vy1;
};

__device__ __forceinline__ Bilinear bilinear_setup(float loc_x, float loc_y, int H, int W) {
  Bilinear b;
  const float x = loc_x * (float)W - 0.5f;
  const float y = loc_y * (float)H - 0.5f;
  b.in_range = (y > -1.0f) && (x > -1.0f) && (y < (float)H) && (x < (float)W);
  const float xf = floorf(x), yf = floorf(y);
  b.x0 = (int)xf;
  b.y0 = (int)yf;
  b.lw = x - xf;
  b.lh = y - yf;
  b.vx0 = b.in_range && (b.x0 >= 0);
  b.vx1 = b.in_range && (b.x0 + 1 <= W - 1);
  b.vy0 = b.in_range && (b.y0 >= 0);
  b.vy1 = b.in_range && (b.y0 + 1 <= H - 1);
  return b;
}

}  // namespace msda
