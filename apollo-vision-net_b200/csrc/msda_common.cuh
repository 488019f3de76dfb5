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
  __device__ __forceinline__ static void unpack2(const uint4& u, float2 (&f)[2]) {
    f[0] = make_float2(__uint_as_float(u.x), __uint_as_float(u.y));
    f[1] = make_float2(__uint_as_float(u.z), __uint_as_float(u.w));
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
  // the two channels of a 32-bit word land in an aligned register pair, ready for FFMA2
  __device__ __forceinline__ static void unpack2(const uint4& u, float2 (&f)[4]) {
    f[0] = make_float2(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u));
    f[1] = make_float2(__uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
    f[2] = make_float2(__uint_as_float(u.z << 16), __uint_as_float(u.z & 0xffff0000u));
    f[3] = make_float2(__uint_as_float(u.w << 16), __uint_as_float(u.w & 0xffff0000u));
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
  __device__ __forceinline__ static void unpack2(const uint4& u, float2 (&f)[4]) {
    f[0] = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
    f[1] = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    f[2] = __half22float2(*reinterpret_cast<const __half2*>(&u.z));
    f[3] = __half22float2(*reinterpret_cast<const __half2*>(&u.w));
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

// Packed fp32x2 arithmetic (Blackwell FFMA2 / FMUL2: two IEEE fp32 operations per instruction,
// each half rounded exactly like the scalar instruction).
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(*reinterpret_cast<unsigned long long*>(&d))
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)),
        "l"(*reinterpret_cast<const unsigned long long*>(&b)),
        "l"(*reinterpret_cast<const unsigned long long*>(&c)));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;"
      : "=l"(*reinterpret_cast<unsigned long long*>(&d))
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)),
        "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  return d;
}
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }

// Mixed-precision FMA of sm_100 (SASS FHFMA): fp32 accumulator += a * b with a, b in the 16-bit
// value dtype.  The product of two bf16 / fp16 numbers is exact in fp32, so nothing is lost
// against unpack + FFMA; the operands are taken straight from the halves of the loaded words.
template <typename T> __device__ __forceinline__ float fma_mixed(uint16_t a, uint16_t b, float c);
template <> __device__ __forceinline__ float fma_mixed<__nv_bfloat16>(uint16_t a, uint16_t b, float c) {
  float d;
  asm("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c));
  return d;
}
template <> __device__ __forceinline__ float fma_mixed<__half>(uint16_t a, uint16_t b, float c) {
  float d;
  asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c));
  return d;
}
__device__ __forceinline__ uint16_t lo16(uint32_t w) { return (uint16_t)(w & 0xffffu); }
__device__ __forceinline__ uint16_t hi16(uint32_t w) { return (uint16_t)(w >> 16); }

// acc[0..7] += (8 channels of u) * w, w in the value dtype
template <typename T>
__device__ __forceinline__ void axpy_mixed(const uint4& u, uint16_t w, float (&acc)[8]) {
  acc[0] = fma_mixed<T>(lo16(u.x), w, acc[0]); acc[1] = fma_mixed<T>(hi16(u.x), w, acc[1]);
  acc[2] = fma_mixed<T>(lo16(u.y), w, acc[2]); acc[3] = fma_mixed<T>(hi16(u.y), w, acc[3]);
  acc[4] = fma_mixed<T>(lo16(u.z), w, acc[4]); acc[5] = fma_mixed<T>(hi16(u.z), w, acc[5]);
  acc[6] = fma_mixed<T>(lo16(u.w), w, acc[6]); acc[7] = fma_mixed<T>(hi16(u.w), w, acc[7]);
}
// sum over the 8 channels of u * g (two chains of four)
template <typename T>
__device__ __forceinline__ float dot_mixed(const uint4& u, const uint4& g) {
  float a = fma_mixed<T>(lo16(u.x), lo16(g.x), 0.f), b = fma_mixed<T>(hi16(u.x), hi16(g.x), 0.f);
  a = fma_mixed<T>(lo16(u.y), lo16(g.y), a); b = fma_mixed<T>(hi16(u.y), hi16(g.y), b);
  a = fma_mixed<T>(lo16(u.z), lo16(g.z), a); b = fma_mixed<T>(hi16(u.z), hi16(g.z), b);
  a = fma_mixed<T>(lo16(u.w), lo16(g.w), a); b = fma_mixed<T>(hi16(u.w), hi16(g.w), b);
  return a + b;
}

// 128-bit read-only global load (value maps are re-read many times: keep them in L1/L2).
__device__ __forceinline__ uint4 ldg128(const void* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}

// Predicated form: zeros where `on` is false (the load is not issued).  Keeps a loop of gathers
// branch-free, so the compiler still batches the loads of several samples.
__device__ __forceinline__ uint4 ldg128_if(const void* p, bool on) {
  uint4 v;
  asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t"
      "mov.u32 %0, 0;\n\tmov.u32 %1, 0;\n\tmov.u32 %2, 0;\n\tmov.u32 %3, 0;\n\t"
      "@q ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];\n\t}"
      : "=&r"(v.x), "=&r"(v.y), "=&r"(v.z), "=&r"(v.w) : "l"(p), "r"((uint32_t)on));
  return v;
}

// Vector reduction: one 16-byte fp32x4 add at L2 (sm_90+), no return value.
__device__ __forceinline__ void red_add_f32x4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// 16-byte reduction of 8 packed fp16 values (4 x f16x2) at L2: half the sectors of the fp32 form
__device__ __forceinline__ void red_add_f16x8(__half* addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("red.global.add.noftz.v4.f16x2 [%0], {%1, %2, %3, %4};"
               :: "l"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// predicated forms: the reduction is issued only where `on` is non-zero (no branch)
__device__ __forceinline__ void red_add_f32x4_if(float* addr, float a, float b, float c, float d, uint32_t on) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t"
               "@p red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n\t}"
               :: "l"(addr), "f"(a), "f"(b), "f"(c), "f"(d), "r"(on) : "memory");
}
__device__ __forceinline__ void red_add_f16x8_if(__half* addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d,
                                                 uint32_t on) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t"
               "@p red.global.add.noftz.v4.f16x2 [%0], {%1, %2, %3, %4};\n\t}"
               :: "l"(addr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(on) : "memory");
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

// `pixel`: the location already is the pixel coordinate (the DCNv3 entry points, which state their sampling
// positions in pixels of the input map: no normalise / de-normalise round trip).
__device__ __forceinline__ Bilinear bilinear_setup(float loc_x, float loc_y, int H, int W, bool pixel = false) {
  Bilinear b;
  const float x = pixel ? loc_x : loc_x * (float)W - 0.5f;
  const float y = pixel ? loc_y : loc_y * (float)H - 0.5f;
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

// Bilinear corner set-up with CLAMPED (always loadable) corner offsets: a corner outside the map
// or a sample outside the open interval keeps a valid address but gets weight zero and a cleared
// `valid` bit, so the loads need no predicates and can be hoisted freely.
struct Corners {
  int o00, o01, o10, o11;      // element offsets from the level base
  float w00, w01, w10, w11;    // bilinear weights (0 for invalid corners)
  float lw, lh, hw, hh;
  unsigned valid;              // bit0..3: corner 00, 01, 10, 11 contributes
};

// `px` receives the (clamped) pixel indices of the four corners inside the level (row-major).
__device__ __forceinline__ Corners corner_setup_px(float loc_x, float loc_y, int H, int W, int pix_stride,
                                                   int (&px)[4], bool pixel = false) {
  Corners c;
  const float x = pixel ? loc_x : loc_x * (float)W - 0.5f;
  const float y = pixel ? loc_y : loc_y * (float)H - 0.5f;
  const bool in = (y > -1.0f) && (x > -1.0f) && (y < (float)H) && (x < (float)W);
  const float xf = floorf(x), yf = floorf(y);
  const int x0 = (int)xf, y0 = (int)yf;             // cvt saturates; NaN -> 0
  c.lw = x - xf;
  c.lh = y - yf;
  c.hw = 1.f - c.lw;
  c.hh = 1.f - c.lh;
  const bool vx0 = in && x0 >= 0, vx1 = in && x0 + 1 <= W - 1;
  const bool vy0 = in && y0 >= 0, vy1 = in && y0 + 1 <= H - 1;
  const int xa = min(max(x0, 0), W - 1), xb = min(max(x0 + 1, 0), W - 1);
  const int ya = min(max(y0, 0), H - 1), yb = min(max(y0 + 1, 0), H - 1);
  const int ra = ya * W, rb = yb * W;
  px[0] = ra + xa; px[1] = ra + xb; px[2] = rb + xa; px[3] = rb + xb;
  c.o00 = px[0] * pix_stride;
  c.o01 = px[1] * pix_stride;
  c.o10 = px[2] * pix_stride;
  c.o11 = px[3] * pix_stride;
  c.w00 = (vy0 && vx0) ? c.hh * c.hw : 0.f;
  c.w01 = (vy0 && vx1) ? c.hh * c.lw : 0.f;
  c.w10 = (vy1 && vx0) ? c.lh * c.hw : 0.f;
  c.w11 = (vy1 && vx1) ? c.lh * c.lw : 0.f;
  c.valid = (unsigned)(vy0 && vx0) | ((unsigned)(vy0 && vx1) << 1) | ((unsigned)(vy1 && vx0) << 2) |
            ((unsigned)(vy1 && vx1) << 3);
  return c;
}
__device__ __forceinline__ Corners corner_setup(float loc_x, float loc_y, int H, int W, int pix_stride,
                                                bool pixel = false) {
  int px[4];
  return corner_setup_px(loc_x, loc_y, H, W, pix_stride, px, pixel);
}

}  // namespace msda
