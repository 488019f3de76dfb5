// Backward of the multi-scale deformable attention operator at the op boundary, sm_100a.
//
// Replaces the reference-era kernel reached through ext_module.ms_deform_attn_backward
// (multi_scale_deformable_attn_function.py:72-82, :148-158), whose structure is visible in-tree
// in the sibling DCNv3 op (ops_dcnv3/src/cuda/dcnv3_im2col_cuda.cuh:82-147, 278-361): one
// 32-thread block per (query, head), 4 scalar atomicAdd per (sample, channel), and a serial
// shared-memory sum by thread 0 for the location / weight gradients.
// Here: one thread owns 16 bytes of the head's channels, so every corner is one 128-bit load
// and one (fp32 value) or two (16-bit value) 16-byte vector reductions `red.global.add.v4.f32`
// into the fp32 grad_value accumulator; grad_loc / grad_attn are reduced over the head's lanes
// with warp shuffles (no barriers, no shared memory) and written exactly once.
#include "msda_common.cuh"
#include "msda_host.h"

namespace msda {

constexpr int kBwdThreads = 256;

template <int TPH>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = TPH / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T, typename CT, int TPH, bool STAGED>
__global__ void __launch_bounds__(kBwdThreads)
msda_bwd_vec_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes,
                    const int64_t* __restrict__ starts, const CT* __restrict__ loc,
                    const CT* __restrict__ attn, const T* __restrict__ grad_out,
                    float* __restrict__ g_value, float* __restrict__ g_loc,
                    float* __restrict__ g_attn,
                    int Nk, int M, int Dh, int L, int Nq, int P, long long total_rows, float pixel_scale) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int ROWS = kBwdThreads / TPH;
  __shared__ LevelTable lv;
  __shared__ __align__(8) uint64_t bar;
  extern __shared__ __align__(16) unsigned char dyn_smem[];

  const int LP = L * P;
  const int tid = threadIdx.x;
  const long long row0 = (long long)blockIdx.x * ROWS;
  const int rows_here = (int)min((long long)ROWS, total_rows - row0);

  load_level_table(lv, shapes, starts, L);

  CT* s_loc = reinterpret_cast<CT*>(dyn_smem);
  CT* s_att = s_loc + (size_t)ROWS * LP * 2;
  if (STAGED) {
    const uint32_t loc_bytes = (uint32_t)(rows_here * LP * 2 * sizeof(CT));
    const uint32_t att_bytes = (uint32_t)(rows_here * LP * sizeof(CT));
    const CT* gl = loc + row0 * LP * 2;
    const CT* ga = attn + row0 * LP;
    const bool bulk_ok = ((loc_bytes | att_bytes) & 15u) == 0 &&
                         ((reinterpret_cast<uintptr_t>(gl) | reinterpret_cast<uintptr_t>(ga)) & 15u) == 0;
    if (bulk_ok) {
      if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
      }
      __syncthreads();
      if (tid == 0) {
        mbar_expect_tx(&bar, loc_bytes + att_bytes);
        bulk_g2s(s_loc, gl, loc_bytes, &bar);
        bulk_g2s(s_att, ga, att_bytes, &bar);
      }
      mbar_wait(&bar, 0);
    } else {
      for (int i = tid; i < rows_here * LP * 2; i += kBwdThreads) s_loc[i] = gl[i];
      for (int i = tid; i < rows_here * LP; i += kBwdThreads) s_att[i] = ga[i];
      __syncthreads();
    }
  } else {
    __syncthreads();
  }

  // Lanes past the end stay in the loop (their shuffles are needed) but touch no memory.
  const int r_raw = tid / TPH;
  const bool active = r_raw < rows_here;
  const int r_local = active ? r_raw : 0;
  const int chunk = tid % TPH;
  const long long row = row0 + r_local;
  const int m = (int)(row % M);
  const int b = (int)(row / M / Nq);

  const int pix_stride = M * Dh;
  const CT* my_loc = STAGED ? (s_loc + (size_t)r_local * LP * 2) : (loc + row * LP * 2);
  const CT* my_att = STAGED ? (s_att + (size_t)r_local * LP) : (attn + row * LP);

  constexpr int V2 = VEC / 2;
  constexpr int SC = VEC / 4;
  // Dot products use the natural channel ownership [VEC*chunk, VEC*(chunk+1)) of the 128-bit value
  // loads; the grad_value scatter of a 16-bit value dtype uses [4c, 4c+4) and [4*TPH+4c, ...), so
  // that the TPH lanes of a head cover 16*TPH contiguous bytes (whole sectors) per reduction.
  float2 g[V2], gs[V2];
  {
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    const T* grow_ptr = grad_out + row * Dh;
    const uint4 ug = active ? ldg128(grow_ptr + chunk * VEC) : z;
    Vec16<T>::unpack2(ug, g);
    if (VEC == 4) {
#pragma unroll
      for (int i = 0; i < V2; ++i) gs[i] = g[i];
    } else {
      const T* lo = grow_ptr + 4 * chunk;
      const T* hi = grow_ptr + 4 * TPH + 4 * chunk;
      const bool ok = active;
      gs[0] = ok ? make_float2(to_f32<T>(lo[0]), to_f32<T>(lo[1])) : make_float2(0.f, 0.f);
      gs[1] = ok ? make_float2(to_f32<T>(lo[2]), to_f32<T>(lo[3])) : make_float2(0.f, 0.f);
      gs[V2 - 2] = ok ? make_float2(to_f32<T>(hi[0]), to_f32<T>(hi[1])) : make_float2(0.f, 0.f);
      gs[V2 - 1] = ok ? make_float2(to_f32<T>(hi[2]), to_f32<T>(hi[3])) : make_float2(0.f, 0.f);
    }
  }
  const size_t head_off = ((size_t)b * Nk * M + m) * Dh;
  const T* vhead = value + head_off + chunk * VEC;
  float* ghead = g_value + head_off + 4 * chunk;

  for (int l = 0; l < L; ++l) {
    const int H = lv.h[l], W = lv.w[l];
    const size_t loff = (size_t)lv.start[l] * pix_stride;
#pragma unroll 2
    for (int p = 0; p < P; ++p) {
      const int s = l * P + p;
      const float lx = to_f32<CT>(my_loc[2 * s]);
      const float ly = to_f32<CT>(my_loc[2 * s + 1]);
      const float a = active ? to_f32<CT>(my_att[s]) : 0.f;
      const Corners c = corner_setup(lx, ly, H, W, pix_stride, pixel_scale > 0.f);
      const T* vb = vhead + loff;
      float* gb = ghead + loff;
      const uint4 u00 = ldg128(vb + c.o00);
      const uint4 u01 = ldg128(vb + c.o01);
      const uint4 u10 = ldg128(vb + c.o10);
      const uint4 u11 = ldg128(vb + c.o11);
      auto scatter = [&](int off, float cw) {
        const float aw = a * cw;
        if (aw == 0.f) return;                           // invalid corner, or a zero contribution
        const float2 aw2 = splat2(aw);
        float* dst = gb + off;
#pragma unroll
        for (int k = 0; k < SC; ++k) {
          const float2 p0 = fmul2(aw2, gs[2 * k]), p1 = fmul2(aw2, gs[2 * k + 1]);
          red_add_f32x4(dst + k * 4 * TPH, p0.x, p0.y, p1.x, p1.y);
        }
      };
      scatter(c.o00, c.w00);
      scatter(c.o01, c.w01);
      scatter(c.o10, c.w10);
      scatter(c.o11, c.w11);
      float2 f[V2];
      float2 d;
      Vec16<T>::unpack2(u00, f);
      d = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < V2; ++i) d = ffma2(f[i], g[i], d);
      const float d00 = (c.valid & 1u) ? d.x + d.y : 0.f;
      Vec16<T>::unpack2(u01, f);
      d = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < V2; ++i) d = ffma2(f[i], g[i], d);
      const float d01 = (c.valid & 2u) ? d.x + d.y : 0.f;
      Vec16<T>::unpack2(u10, f);
      d = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < V2; ++i) d = ffma2(f[i], g[i], d);
      const float d10 = (c.valid & 4u) ? d.x + d.y : 0.f;
      Vec16<T>::unpack2(u11, f);
      d = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < V2; ++i) d = ffma2(f[i], g[i], d);
      const float d11 = (c.valid & 8u) ? d.x + d.y : 0.f;
      float ga = c.w00 * d00 + c.w01 * d01 + c.w10 * d10 + c.w11 * d11;
      float gx = c.hh * (d01 - d00) + c.lh * (d11 - d10);
      float gy = c.hw * (d10 - d00) + c.lw * (d11 - d01);
      ga = group_sum<TPH>(ga);
      gx = group_sum<TPH>(gx);
      gy = group_sum<TPH>(gy);
      if (active && chunk == 0) {
        g_attn[row * LP + s] = ga;
        *reinterpret_cast<float2*>(g_loc + (row * LP + s) * 2) =
            make_float2((pixel_scale > 0.f ? pixel_scale : (float)W) * a * gx,
                        (pixel_scale > 0.f ? pixel_scale : (float)H) * a * gy);
      }
    }
  }
}

// Fully generic fallback (any Dh): one thread per sample (row, l, p), serial over channels.
template <typename T, typename CT>
__global__ void __launch_bounds__(256)
msda_bwd_scalar_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes,
                       const int64_t* __restrict__ starts, const CT* __restrict__ loc,
                       const CT* __restrict__ attn, const T* __restrict__ grad_out,
                       float* __restrict__ g_value, float* __restrict__ g_loc,
                       float* __restrict__ g_attn,
                       int Nk, int M, int Dh, int L, int Nq, int P, long long total, float pixel_scale) {
  __shared__ LevelTable lv;
  load_level_table(lv, shapes, starts, L);
  __syncthreads();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int LP = L * P;
  const int s = (int)(idx % LP);
  const int l = s / P;
  const long long row = idx / LP;
  const int m = (int)(row % M);
  const int b = (int)(row / M / Nq);
  const int H = lv.h[l], W = lv.w[l];
  const int pix_stride = M * Dh;
  const size_t voff = ((size_t)b * Nk * M + m) * Dh + (size_t)lv.start[l] * pix_stride;
  const float a = to_f32<CT>(attn[idx]);
  const Bilinear bl = bilinear_setup(to_f32<CT>(loc[2 * idx]), to_f32<CT>(loc[2 * idx + 1]), H, W,
                                     pixel_scale > 0.f);
  const float hw = 1.f - bl.lw, hh = 1.f - bl.lh;
  const long long o00 = (long long)voff + ((long long)bl.y0 * W + bl.x0) * pix_stride;
  const long long o01 = o00 + pix_stride, o10 = o00 + (long long)W * pix_stride, o11 = o10 + pix_stride;
  const bool c00 = bl.vy0 && bl.vx0, c01 = bl.vy0 && bl.vx1, c10 = bl.vy1 && bl.vx0, c11 = bl.vy1 && bl.vx1;
  const float w00 = hh * hw, w01 = hh * bl.lw, w10 = bl.lh * hw, w11 = bl.lh * bl.lw;
  float d00 = 0.f, d01 = 0.f, d10 = 0.f, d11 = 0.f;
  const T* go = grad_out + row * Dh;
  for (int c = 0; c < Dh; ++c) {
    const float g = to_f32<T>(go[c]);
    if (c00) { d00 = fmaf(to_f32<T>(value[o00 + c]), g, d00); atomicAdd(g_value + o00 + c, a * w00 * g); }
    if (c01) { d01 = fmaf(to_f32<T>(value[o01 + c]), g, d01); atomicAdd(g_value + o01 + c, a * w01 * g); }
    if (c10) { d10 = fmaf(to_f32<T>(value[o10 + c]), g, d10); atomicAdd(g_value + o10 + c, a * w10 * g); }
    if (c11) { d11 = fmaf(to_f32<T>(value[o11 + c]), g, d11); atomicAdd(g_value + o11 + c, a * w11 * g); }
  }
  g_attn[idx] = w00 * d00 + w01 * d01 + w10 * d10 + w11 * d11;
  g_loc[2 * idx] = (pixel_scale > 0.f ? pixel_scale : (float)W) * a * (hh * (d01 - d00) + bl.lh * (d11 - d10));
  g_loc[2 * idx + 1] = (pixel_scale > 0.f ? pixel_scale : (float)H) * a * (hw * (d10 - d00) + bl.lw * (d11 - d01));
}

template <typename T, typename CT, int TPH>
static int launch_vec(const Problem& pr, cudaStream_t st) {
  constexpr int ROWS = kBwdThreads / TPH;
  const long long rows = (long long)pr.B * pr.Nq * pr.M;
  const long long grid = (rows + ROWS - 1) / ROWS;
  if (grid <= 0) return MSDA_OK;
  if (grid > 0x7fffffffLL) return set_error(MSDA_ERR_UNSUPPORTED, "msda_bwd: problem too large for one launch");
  const size_t smem = (size_t)ROWS * pr.L * pr.P * 3 * sizeof(CT);
  const T* v = static_cast<const T*>(pr.value);
  const CT* lo = static_cast<const CT*>(pr.loc);
  const CT* at = static_cast<const CT*>(pr.attn);
  const T* go = static_cast<const T*>(pr.grad_out);
  if (smem <= 40 * 1024) {
    msda_bwd_vec_kernel<T, CT, TPH, true><<<(unsigned)grid, kBwdThreads, smem, st>>>(
        v, pr.shapes, pr.starts, lo, at, go, pr.g_value, pr.g_loc, pr.g_attn,
        pr.Nk, pr.M, pr.Dh, pr.L, pr.Nq, pr.P, rows, pr.pixel_scale);
  } else {
    msda_bwd_vec_kernel<T, CT, TPH, false><<<(unsigned)grid, kBwdThreads, 0, st>>>(
        v, pr.shapes, pr.starts, lo, at, go, pr.g_value, pr.g_loc, pr.g_attn,
        pr.Nk, pr.M, pr.Dh, pr.L, pr.Nq, pr.P, rows, pr.pixel_scale);
  }
  count_launch();
  return check_launch("msda_bwd");
}

template <typename T, typename CT>
static int launch_bwd_typed(const Problem& pr, cudaStream_t st) {
  constexpr int VEC = Vec16<T>::N;
  const bool aligned = (reinterpret_cast<uintptr_t>(pr.value) % 16 == 0) &&
                       (reinterpret_cast<uintptr_t>(pr.grad_out) % 16 == 0) &&
                       (reinterpret_cast<uintptr_t>(pr.g_value) % 16 == 0) &&
                       (reinterpret_cast<uintptr_t>(pr.g_loc) % 8 == 0);
  if (pr.Dh % VEC == 0 && aligned) {
    switch (pr.Dh / VEC) {
      case 1: return launch_vec<T, CT, 1>(pr, st);
      case 2: return launch_vec<T, CT, 2>(pr, st);
      case 4: return launch_vec<T, CT, 4>(pr, st);
      case 8: return launch_vec<T, CT, 8>(pr, st);
      case 16: return launch_vec<T, CT, 16>(pr, st);
      default: break;
    }
  }
  const long long total = (long long)pr.B * pr.Nq * pr.M * pr.L * pr.P;
  const long long grid = (total + 255) / 256;
  if (grid <= 0) return MSDA_OK;
  if (grid > 0x7fffffffLL) return set_error(MSDA_ERR_UNSUPPORTED, "msda_bwd: problem too large for one launch");
  msda_bwd_scalar_kernel<T, CT><<<(unsigned)grid, 256, 0, st>>>(
      static_cast<const T*>(pr.value), pr.shapes, pr.starts, static_cast<const CT*>(pr.loc),
      static_cast<const CT*>(pr.attn), static_cast<const T*>(pr.grad_out), pr.g_value, pr.g_loc,
      pr.g_attn, pr.Nk, pr.M, pr.Dh, pr.L, pr.Nq, pr.P, total, pr.pixel_scale);
  count_launch();
  return check_launch("msda_bwd(scalar)");
}

int launch_msda_bwd(const Problem& pr, cudaStream_t st) {
  if (pr.value_dtype == MSDA_F32) return launch_bwd_typed<float, float>(pr, st);
  if (pr.value_dtype == MSDA_BF16) {
    return pr.coord_dtype == MSDA_F32 ? launch_bwd_typed<__nv_bfloat16, float>(pr, st)
                                      : launch_bwd_typed<__nv_bfloat16, __nv_bfloat16>(pr, st);
  }
  return pr.coord_dtype == MSDA_F32 ? launch_bwd_typed<__half, float>(pr, st)
                                    : launch_bwd_typed<__half, __half>(pr, st);
}

}  // namespace msda
