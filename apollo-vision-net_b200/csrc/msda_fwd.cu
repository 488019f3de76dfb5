// Forward of the multi-scale deformable attention operator at the op boundary
// (materialised sampling_locations / attention_weights), sm_100a.
//
// Replaces the reference-era kernel reached through ext_module.ms_deform_attn_forward
// (multi_scale_deformable_attn_function.py:40-46, :116-122): that one runs one thread per
// output scalar, re-reads locations / weights in every channel thread and issues 4-byte loads.
// Here one thread owns 16 bytes of the output row (4 fp32 / 8 bf16 channels of one head), the
// CTA's slab of locations and weights is staged once in shared memory by a TMA bulk copy, the
// level table is staged once per CTA, and every corner is one 128-bit read-only load.
#include "msda_common.cuh"
#include "msda_host.h"

namespace msda {

constexpr int kFwdThreads = 256;

template <typename CT> struct Coord2 { CT x, y; };

template <typename T, typename CT, int TPH, bool STAGED>
__global__ void __launch_bounds__(kFwdThreads)
msda_fwd_vec_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes,
                    const int64_t* __restrict__ starts, const CT* __restrict__ loc,
                    const CT* __restrict__ attn, T* __restrict__ out,
                    int Nk, int M, int Dh, int L, int Nq, int P, long long total_rows, float pixel_scale) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int ROWS = kFwdThreads / TPH;      // (b, q, m) rows per CTA
  __shared__ LevelTable lv;
  __shared__ __align__(8) uint64_t bar;
  extern __shared__ __align__(16) unsigned char dyn_smem[];

  const int LP = L * P;
  const int tid = threadIdx.x;
  const long long row0 = (long long)blockIdx.x * ROWS;
  const int rows_here = (int)min((long long)ROWS, total_rows - row0);

  load_level_table(lv, shapes, starts, L);

  CT* s_loc = reinterpret_cast<CT*>(dyn_smem);                 // [ROWS][LP][2]
  CT* s_att = s_loc + (size_t)ROWS * LP * 2;                   // [ROWS][LP]
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
      for (int i = tid; i < rows_here * LP * 2; i += kFwdThreads) s_loc[i] = gl[i];
      for (int i = tid; i < rows_here * LP; i += kFwdThreads) s_att[i] = ga[i];
      __syncthreads();
    }
  } else {
    __syncthreads();
  }

  const int r_local = tid / TPH;
  const int chunk = tid % TPH;
  if (r_local >= rows_here) return;
  const long long row = row0 + r_local;                         // (b*Nq + q)*M + m
  const int m = (int)(row % M);
  const long long bq = row / M;
  const int b = (int)(bq / Nq);

  const int pix_stride = M * Dh;                                // elements between pixels
  const T* vbase = value + ((size_t)b * Nk * M + m) * Dh + chunk * VEC;
  const CT* my_loc = STAGED ? (s_loc + (size_t)r_local * LP * 2) : (loc + row * LP * 2);
  const CT* my_att = STAGED ? (s_att + (size_t)r_local * LP) : (attn + row * LP);

  constexpr int V2 = VEC / 2;
  float2 acc2[V2];
#pragma unroll
  for (int i = 0; i < V2; ++i) acc2[i] = make_float2(0.f, 0.f);

  for (int l = 0; l < L; ++l) {
    const int H = lv.h[l], W = lv.w[l];
    const T* lbase = vbase + (size_t)lv.start[l] * pix_stride;
#pragma unroll 4
    for (int p = 0; p < P; ++p) {
      const int s = l * P + p;
      const float lx = to_f32<CT>(my_loc[2 * s]);
      const float ly = to_f32<CT>(my_loc[2 * s + 1]);
      const float a = to_f32<CT>(my_att[s]);
      // clamped corners: unpredicated 128-bit loads, weight zero for corners outside the map.  An
      // invalid corner of a partly valid sample clamps onto the pixel of one of its valid corners;
      // a sample entirely outside the map reads nothing (predicated loads), so a non-finite value only reaches the samples
      // that touch it (the reference kernel never reads such corners)
      const Corners c = corner_setup(lx, ly, H, W, pix_stride, pixel_scale > 0.f);
      const bool touch = c.valid != 0u;
      const uint4 u00 = ldg128_if(lbase + c.o00, touch);
      const uint4 u01 = ldg128_if(lbase + c.o01, touch);
      const uint4 u10 = ldg128_if(lbase + c.o10, touch);
      const uint4 u11 = ldg128_if(lbase + c.o11, touch);
      const float2 w00 = splat2(a * c.w00), w01 = splat2(a * c.w01);
      const float2 w10 = splat2(a * c.w10), w11 = splat2(a * c.w11);
      float2 f[V2];
      Vec16<T>::unpack2(u00, f);
#pragma unroll
      for (int i = 0; i < V2; ++i) acc2[i] = ffma2(w00, f[i], acc2[i]);
      Vec16<T>::unpack2(u01, f);
#pragma unroll
      for (int i = 0; i < V2; ++i) acc2[i] = ffma2(w01, f[i], acc2[i]);
      Vec16<T>::unpack2(u10, f);
#pragma unroll
      for (int i = 0; i < V2; ++i) acc2[i] = ffma2(w10, f[i], acc2[i]);
      Vec16<T>::unpack2(u11, f);
#pragma unroll
      for (int i = 0; i < V2; ++i) acc2[i] = ffma2(w11, f[i], acc2[i]);
    }
  }
  float acc[VEC];
#pragma unroll
  for (int i = 0; i < V2; ++i) {
    acc[2 * i] = acc2[i].x;
    acc[2 * i + 1] = acc2[i].y;
  }
  T* o = out + row * Dh + chunk * VEC;
  *reinterpret_cast<uint4*>(o) = Vec16<T>::pack(acc);
}

// Fully generic fallback (any Dh): one thread per output scalar.
template <typename T, typename CT>
__global__ void __launch_bounds__(256)
msda_fwd_scalar_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes,
                       const int64_t* __restrict__ starts, const CT* __restrict__ loc,
                       const CT* __restrict__ attn, T* __restrict__ out,
                       int Nk, int M, int Dh, int L, int Nq, int P, long long total, float pixel_scale) {
  __shared__ LevelTable lv;
  load_level_table(lv, shapes, starts, L);
  __syncthreads();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % Dh);
  const long long row = idx / Dh;
  const int m = (int)(row % M);
  const int b = (int)(row / M / Nq);
  const int LP = L * P;
  const int pix_stride = M * Dh;
  const T* vbase = value + ((size_t)b * Nk * M + m) * Dh + c;
  const CT* my_loc = loc + row * LP * 2;
  const CT* my_att = attn + row * LP;
  float acc = 0.f;
  for (int l = 0; l < L; ++l) {
    const int H = lv.h[l], W = lv.w[l];
    const T* lbase = vbase + (size_t)lv.start[l] * pix_stride;
    for (int p = 0; p < P; ++p) {
      const int s = l * P + p;
      const Bilinear bl = bilinear_setup(to_f32<CT>(my_loc[2 * s]), to_f32<CT>(my_loc[2 * s + 1]), H, W,
                                         pixel_scale > 0.f);
      const float a = to_f32<CT>(my_att[s]);
      const float hw = 1.f - bl.lw, hh = 1.f - bl.lh;
      const T* p00 = lbase + ((long long)bl.y0 * W + bl.x0) * pix_stride;
      const float v00 = (bl.vy0 && bl.vx0) ? to_f32<T>(p00[0]) : 0.f;
      const float v01 = (bl.vy0 && bl.vx1) ? to_f32<T>(p00[pix_stride]) : 0.f;
      const float v10 = (bl.vy1 && bl.vx0) ? to_f32<T>(p00[(size_t)W * pix_stride]) : 0.f;
      const float v11 = (bl.vy1 && bl.vx1) ? to_f32<T>(p00[(size_t)(W + 1) * pix_stride]) : 0.f;
      const float v = hh * hw * v00 + hh * bl.lw * v01 + bl.lh * hw * v10 + bl.lh * bl.lw * v11;
      acc = fmaf(a, v, acc);
    }
  }
  out[idx] = from_f32<T>(acc);
}

template <typename T, typename CT, int TPH>
static int launch_vec(const Problem& pr, cudaStream_t st) {
  constexpr int ROWS = kFwdThreads / TPH;
  const long long rows = (long long)pr.B * pr.Nq * pr.M;
  const long long grid = (rows + ROWS - 1) / ROWS;
  if (grid <= 0) return MSDA_OK;
  if (grid > 0x7fffffffLL) return set_error(MSDA_ERR_UNSUPPORTED, "msda_fwd: problem too large for one launch");
  const size_t smem = (size_t)ROWS * pr.L * pr.P * 3 * sizeof(CT);
  const T* v = static_cast<const T*>(pr.value);
  const CT* lo = static_cast<const CT*>(pr.loc);
  const CT* at = static_cast<const CT*>(pr.attn);
  T* o = static_cast<T*>(pr.out);
  if (smem <= 40 * 1024) {
    msda_fwd_vec_kernel<T, CT, TPH, true><<<(unsigned)grid, kFwdThreads, smem, st>>>(
        v, pr.shapes, pr.starts, lo, at, o, pr.Nk, pr.M, pr.Dh, pr.L, pr.Nq, pr.P, rows, pr.pixel_scale);
  } else {
    msda_fwd_vec_kernel<T, CT, TPH, false><<<(unsigned)grid, kFwdThreads, 0, st>>>(
        v, pr.shapes, pr.starts, lo, at, o, pr.Nk, pr.M, pr.Dh, pr.L, pr.Nq, pr.P, rows, pr.pixel_scale);
  }
  count_launch();
  return check_launch("msda_fwd");
}

template <typename T, typename CT>
static int launch_fwd_typed(const Problem& pr, cudaStream_t st) {
  constexpr int VEC = Vec16<T>::N;
  const bool aligned = (reinterpret_cast<uintptr_t>(pr.value) % 16 == 0) &&
                       (reinterpret_cast<uintptr_t>(pr.out) % 16 == 0);
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
  const long long total = (long long)pr.B * pr.Nq * pr.M * pr.Dh;
  const long long grid = (total + 255) / 256;
  if (grid <= 0) return MSDA_OK;
  if (grid > 0x7fffffffLL) return set_error(MSDA_ERR_UNSUPPORTED, "msda_fwd: problem too large for one launch");
  msda_fwd_scalar_kernel<T, CT><<<(unsigned)grid, 256, 0, st>>>(
      static_cast<const T*>(pr.value), pr.shapes, pr.starts, static_cast<const CT*>(pr.loc),
      static_cast<const CT*>(pr.attn), static_cast<T*>(pr.out), pr.Nk, pr.M, pr.Dh, pr.L, pr.Nq,
      pr.P, total, pr.pixel_scale);
  count_launch();
  return check_launch("msda_fwd(scalar)");
}

int launch_msda_fwd(const Problem& pr, cudaStream_t st) {
  if (pr.value_dtype == MSDA_F32) return launch_fwd_typed<float, float>(pr, st);
  if (pr.value_dtype == MSDA_BF16) {
    return pr.coord_dtype == MSDA_F32 ? launch_fwd_typed<__nv_bfloat16, float>(pr, st)
                                      : launch_fwd_typed<__nv_bfloat16, __nv_bfloat16>(pr, st);
  }
  return pr.coord_dtype == MSDA_F32 ? launch_fwd_typed<__half, float>(pr, st)
                                    : launch_fwd_typed<__half, __half>(pr, st);
}

}  // namespace msda
