// Fused attention cores for the BEVFormer encoder / MapTRv2 decoder, sm_100a.
//
//  * SCA  (spatial cross-attention, spatial_cross_attention.py:135-170 + :342-396): camera-hit
//    gating from bev_mask, softmax over L*P, Z-anchor sampling locations, 4-level bilinear
//    sampling, sum over the cameras that see the query, division by the hit count.  The reference
//    materialises rebatched queries, sampling_locations (119 MB per layer at 200x200) and a
//    scatter-add; none of those exist here.
//  * TSA / decoder (temporal_self_attention.py:204-279, decoder.py:299-350): optional logit
//    clamp, softmax per queue entry, location = ref + offset / (W_l, H_l), sampling, mean over
//    the queue.
//
// Work decomposition (both, forward and backward): a CTA owns `ROWS` consecutive
// (batch, query, head) rows; the raw Linear outputs for those rows (offsets, logits) are one
// contiguous slab that a single TMA bulk copy stages in shared memory; a short cooperative pass
// turns them into softmax weights and normalised offsets once per row (not once per lane or per
// camera); then every thread owns 16 bytes of one head's channels and walks the row's samples,
// one 128-bit read-only load per bilinear corner.  The backward reduces the location / weight
// gradients over the head's lanes with warp shuffles, finishes the softmax backward in shared
// memory, and sends grad_value to an fp32 accumulator with 16-byte vector reductions.
#include "msda_common.cuh"
#include "msda_host.h"

namespace msda {

constexpr int kFusedThreads = 256;
enum { MODE_SCA = 0, MODE_TSA = 1 };

template <int TPH>
__device__ __forceinline__ float lanes_sum(float v) {
#pragma unroll
  for (int o = TPH / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct FusedArgs {
  const void* value;
  const int64_t* shapes;
  const int64_t* starts;
  const float* offsets;
  const float* logits;
  const float* ref;
  const uint8_t* bev_mask;
  const uint32_t* hit_bits;
  void* out;
  const void* g_out;
  float* g_value;
  float* g_offsets;
  float* g_logits;
  int bs, groups, Nk, M, Dh, L, P, D, Nq;
  float clamp;
  long long total_rows;
};

// Stage offsets / logits of the CTA's rows, then softmax + offset normalisation in place.
// S = samples per row (SCA: L*P, TSA: Q*L*P); softmax segments are LP long.
template <bool WITH_GA>
__device__ __forceinline__ void stage_rows(const FusedArgs& a, const LevelTable& lv, uint64_t* bar,
                                           float* s_off, float* s_w, long long row0, int rows_here,
                                           int S, int LP) {
  const int tid = threadIdx.x;
  const uint32_t off_bytes = (uint32_t)(rows_here * S * 2 * sizeof(float));
  const uint32_t w_bytes = (uint32_t)(rows_here * S * sizeof(float));
  const float* go = a.offsets + row0 * S * 2;
  const float* gl = a.logits + row0 * S;
  const bool bulk_ok = ((off_bytes | w_bytes) & 15u) == 0 &&
                       ((reinterpret_cast<uintptr_t>(go) | reinterpret_cast<uintptr_t>(gl)) & 15u) == 0;
  if (bulk_ok) {
    if (tid == 0) {
      mbar_init(bar, 1);
      fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(bar, off_bytes + w_bytes);
      bulk_g2s(s_off, go, off_bytes, bar);
      bulk_g2s(s_w, gl, w_bytes, bar);
    }
    mbar_wait(bar, 0);
  } else {
    for (int i = tid; i < rows_here * S * 2; i += blockDim.x) s_off[i] = go[i];
    for (int i = tid; i < rows_here * S; i += blockDim.x) s_w[i] = gl[i];
  }
  __syncthreads();
  // softmax per (row, segment)
  const int nseg = rows_here * (S / LP);
  for (int sgi = tid; sgi < nseg; sgi += blockDim.x) {
    float* w = s_w + (size_t)sgi * LP;
    float mx = -INFINITY;
    for (int i = 0; i < LP; ++i) {
      float v = w[i];
      if (a.clamp >= 0.f) v = fminf(fmaxf(v, -a.clamp), a.clamp);
      w[i] = v;
      mx = fmaxf(mx, v);
    }
    float sum = 0.f;
    for (int i = 0; i < LP; ++i) {
      const float e = expf(w[i] - mx);
      w[i] = e;
      sum += e;
    }
    for (int i = 0; i < LP; ++i) w[i] = w[i] / sum;
  }
  // offsets / (W_l, H_l)
  for (int i = tid; i < rows_here * S * 2; i += blockDim.x) {
    const int s = (i >> 1) % S;
    const int l = (s % LP) / a.P;
    const float d = (i & 1) ? (float)lv.h[l] : (float)lv.w[l];
    s_off[i] = s_off[i] / d;
  }
  __syncthreads();
}

template <typename T, int TPH, int MODE>
__global__ void __launch_bounds__(kFusedThreads)
fused_fwd_kernel(const FusedArgs a) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int ROWS = kFusedThreads / TPH;
  __shared__ LevelTable lv;
  __shared__ __align__(8) uint64_t bar;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  const int LP = a.L * a.P;
  const int S = (MODE == MODE_TSA) ? a.groups * LP : LP;
  float* s_off = reinterpret_cast<float*>(dyn_smem);           // [ROWS][S][2]
  float* s_w = s_off + (size_t)ROWS * S * 2;                   // [ROWS][S]

  const int tid = threadIdx.x;
  const long long row0 = (long long)blockIdx.x * ROWS;
  const int rows_here = (int)min((long long)ROWS, a.total_rows - row0);
  load_level_table(lv, a.shapes, a.starts, a.L);
  __syncthreads();
  stage_rows<false>(a, lv, &bar, s_off, s_w, row0, rows_here, S, LP);

  const int r_local = tid / TPH;
  const int chunk = tid % TPH;
  if (r_local >= rows_here) return;
  const long long row = row0 + r_local;
  const int m = (int)(row % a.M);
  const long long bq = row / a.M;
  const int b = (int)(bq / a.Nq);
  const int q = (int)(bq % a.Nq);
  const int pix_stride = a.M * a.Dh;
  const size_t batch_stride = (size_t)a.Nk * pix_stride;
  const T* vhead = static_cast<const T*>(a.value) + (size_t)m * a.Dh + chunk * VEC;
  const float* my_off = s_off + (size_t)r_local * S * 2;
  const float* my_w = s_w + (size_t)r_local * S;

  float acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f;

  uint32_t hits = 0;
  float scale = 1.f;
  if (MODE == MODE_SCA) {
    hits = a.hit_bits[q];                                      // batch element 0 decides (quirk 1)
    const int cnt = __popc(a.hit_bits[(size_t)b * a.Nq + q]);  // count is per sample
    scale = (float)(cnt > 0 ? cnt : 1);
  } else {
    scale = (float)a.groups;
  }

  auto sample = [&](const T* lbase, int H, int W, float lx, float ly, float w) {
    const Bilinear bl = bilinear_setup(lx, ly, H, W);
    const float hw = 1.f - bl.lw, hh = 1.f - bl.lh;
    const T* p00 = lbase + ((long long)bl.y0 * W + bl.x0) * pix_stride;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    const uint4 u00 = (bl.vy0 && bl.vx0) ? ldg128(p00) : z;
    const uint4 u01 = (bl.vy0 && bl.vx1) ? ldg128(p00 + pix_stride) : z;
    const uint4 u10 = (bl.vy1 && bl.vx0) ? ldg128(p00 + (size_t)W * pix_stride) : z;
    const uint4 u11 = (bl.vy1 && bl.vx1) ? ldg128(p00 + (size_t)(W + 1) * pix_stride) : z;
    const float w00 = hh * hw, w01 = hh * bl.lw, w10 = bl.lh * hw, w11 = bl.lh * bl.lw;
    float f00[VEC], f01[VEC], f10[VEC], f11[VEC];
    Vec16<T>::unpack(u00, f00);
    Vec16<T>::unpack(u01, f01);
    Vec16<T>::unpack(u10, f10);
    Vec16<T>::unpack(u11, f11);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float v = w00 * f00[i] + w01 * f01[i] + w10 * f10[i] + w11 * f11[i];
      acc[i] = fmaf(w, v, acc[i]);
    }
  };

  if (MODE == MODE_SCA) {
    if (hits != 0) {
      for (int l = 0; l < a.L; ++l) {
        const int H = lv.h[l], W = lv.w[l];
        const size_t loff = (size_t)lv.start[l] * pix_stride;
        for (int p = 0; p < a.P; ++p) {
          const int s = l * a.P + p;
          const float ox = my_off[2 * s], oy = my_off[2 * s + 1], w = my_w[s];
          const int z = p % a.D;                               // point index p = k*D + z (quirk 5)
          uint32_t h = hits;
          while (h) {
            const int cam = __ffs(h) - 1;
            h &= h - 1;
            const float2 r = *reinterpret_cast<const float2*>(
                a.ref + ((((size_t)cam * a.bs + b) * a.Nq + q) * a.D + z) * 2);
            const T* lbase = vhead + ((size_t)b * a.groups + cam) * batch_stride + loff;
            sample(lbase, H, W, r.x + ox, r.y + oy, w);
          }
        }
      }
    }
  } else {
    for (int j = 0; j < a.groups; ++j) {
      const T* vb = vhead + ((size_t)b * a.groups + j) * batch_stride;
      for (int l = 0; l < a.L; ++l) {
        const int H = lv.h[l], W = lv.w[l];
        const T* lbase = vb + (size_t)lv.start[l] * pix_stride;
        const float2 r = *reinterpret_cast<const float2*>(
            a.ref + ((((size_t)b * a.groups + j) * a.Nq + q) * a.L + l) * 2);
#pragma unroll 4
        for (int p = 0; p < a.P; ++p) {
          const int s = (j * a.L + l) * a.P + p;
          sample(lbase, H, W, r.x + my_off[2 * s], r.y + my_off[2 * s + 1], my_w[s]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = acc[i] / scale;
  T* o = static_cast<T*>(a.out) + row * a.Dh + chunk * VEC;
  *reinterpret_cast<uint4*>(o) = Vec16<T>::pack(acc);
}

template <typename T, int TPH, int MODE>
__global__ void __launch_bounds__(kFusedThreads)
fused_bwd_kernel(const FusedArgs a) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int ROWS = kFusedThreads / TPH;
  __shared__ LevelTable lv;
  __shared__ __align__(8) uint64_t bar;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  const int LP = a.L * a.P;
  const int S = (MODE == MODE_TSA) ? a.groups * LP : LP;
  float* s_off = reinterpret_cast<float*>(dyn_smem);           // [ROWS][S][2]
  float* s_w = s_off + (size_t)ROWS * S * 2;                   // [ROWS][S]
  float* s_ga = s_w + (size_t)ROWS * S;                        // [ROWS][S] grad wrt softmax output

  const int tid = threadIdx.x;
  const long long row0 = (long long)blockIdx.x * ROWS;
  const int rows_here = (int)min((long long)ROWS, a.total_rows - row0);
  load_level_table(lv, a.shapes, a.starts, a.L);
  __syncthreads();
  stage_rows<true>(a, lv, &bar, s_off, s_w, row0, rows_here, S, LP);

  const int r_raw = tid / TPH;
  const bool active = r_raw < rows_here;
  const int r_local = active ? r_raw : 0;
  const int chunk = tid % TPH;
  const long long row = row0 + r_local;
  const int m = (int)(row % a.M);
  const long long bq = row / a.M;
  const int b = (int)(bq / a.Nq);
  const int q = (int)(bq % a.Nq);
  const int pix_stride = a.M * a.Dh;
  const size_t batch_stride = (size_t)a.Nk * pix_stride;
  const size_t head_off = (size_t)m * a.Dh + chunk * VEC;
  const T* vhead = static_cast<const T*>(a.value) + head_off;
  float* ghead = a.g_value + head_off;
  const float* my_off = s_off + (size_t)r_local * S * 2;
  const float* my_w = s_w + (size_t)r_local * S;
  float* my_ga = s_ga + (size_t)r_local * S;

  uint32_t hits = 0;
  float scale = 1.f;
  if (MODE == MODE_SCA) {
    hits = a.hit_bits[q];
    const int cnt = __popc(a.hit_bits[(size_t)b * a.Nq + q]);
    scale = (float)(cnt > 0 ? cnt : 1);
  } else {
    scale = (float)a.groups;
  }
  if (!active) hits = 0;

  float g[VEC];
  {
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    const uint4 ug = active ? ldg128(static_cast<const T*>(a.g_out) + row * a.Dh + chunk * VEC) : z;
    Vec16<T>::unpack(ug, g);
#pragma unroll
    for (int i = 0; i < VEC; ++i) g[i] = g[i] / scale;
  }

  // One sample against one value map: scatters grad_value, returns the three partial dots.
  auto sample = [&](size_t boff, int H, int W, float lx, float ly, float w, float& ga, float& gx,
                    float& gy) {
    const Bilinear bl = bilinear_setup(lx, ly, H, W);
    const float hw = 1.f - bl.lw, hh = 1.f - bl.lh;
    const long long o00 = (long long)boff + ((long long)bl.y0 * W + bl.x0) * pix_stride;
    const long long o01 = o00 + pix_stride;
    const long long o10 = o00 + (long long)W * pix_stride;
    const long long o11 = o10 + pix_stride;
    const bool c00 = bl.vy0 && bl.vx0, c01 = bl.vy0 && bl.vx1;
    const bool c10 = bl.vy1 && bl.vx0, c11 = bl.vy1 && bl.vx1;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    const uint4 u00 = c00 ? ldg128(vhead + o00) : z;
    const uint4 u01 = c01 ? ldg128(vhead + o01) : z;
    const uint4 u10 = c10 ? ldg128(vhead + o10) : z;
    const uint4 u11 = c11 ? ldg128(vhead + o11) : z;
    const float w00 = hh * hw, w01 = hh * bl.lw, w10 = bl.lh * hw, w11 = bl.lh * bl.lw;
    auto scatter = [&](bool ok, long long off, float cw) {
      if (!ok) return;
      const float aw = w * cw;
      float* dst = ghead + off;
#pragma unroll
      for (int i = 0; i < VEC; i += 4)
        red_add_f32x4(dst + i, aw * g[i], aw * g[i + 1], aw * g[i + 2], aw * g[i + 3]);
    };
    scatter(c00, o00, w00);
    scatter(c01, o01, w01);
    scatter(c10, o10, w10);
    scatter(c11, o11, w11);
    float f[VEC];
    float d00 = 0.f, d01 = 0.f, d10 = 0.f, d11 = 0.f;
    Vec16<T>::unpack(u00, f);
#pragma unroll
    for (int i = 0; i < VEC; ++i) d00 = fmaf(f[i], g[i], d00);
    Vec16<T>::unpack(u01, f);
#pragma unroll
    for (int i = 0; i < VEC; ++i) d01 = fmaf(f[i], g[i], d01);
    Vec16<T>::unpack(u10, f);
#pragma unroll
    for (int i = 0; i < VEC; ++i) d10 = fmaf(f[i], g[i], d10);
    Vec16<T>::unpack(u11, f);
#pragma unroll
    for (int i = 0; i < VEC; ++i) d11 = fmaf(f[i], g[i], d11);
    ga += w00 * d00 + w01 * d01 + w10 * d10 + w11 * d11;
    gx += hh * (d01 - d00) + bl.lh * (d11 - d10);
    gy += hw * (d10 - d00) + bl.lw * (d11 - d01);
  };

  float* go_row = a.g_offsets + row * S * 2;
  auto finish_sample = [&](int s, float w, float ga, float gx, float gy) {
    ga = lanes_sum<TPH>(ga);
    gx = lanes_sum<TPH>(gx);
    gy = lanes_sum<TPH>(gy);
    if (active && chunk == 0) {
      my_ga[s] = ga;
      // d loc / d offset = 1 / (W_l, H_l) cancels the (W_l, H_l) factor of d pixel / d loc
      *reinterpret_cast<float2*>(go_row + 2 * s) = make_float2(w * gx, w * gy);
    }
  };

  if (MODE == MODE_SCA) {
    for (int l = 0; l < a.L; ++l) {
      const int H = lv.h[l], W = lv.w[l];
      const size_t loff = (size_t)lv.start[l] * pix_stride;
      for (int p = 0; p < a.P; ++p) {
        const int s = l * a.P + p;
        const float ox = my_off[2 * s], oy = my_off[2 * s + 1], w = my_w[s];
        const int z = p % a.D;
        float ga = 0.f, gx = 0.f, gy = 0.f;
        uint32_t h = hits;
        while (h) {
          const int cam = __ffs(h) - 1;
          h &= h - 1;
          const float2 r = *reinterpret_cast<const float2*>(
              a.ref + ((((size_t)cam * a.bs + b) * a.Nq + q) * a.D + z) * 2);
          sample(((size_t)b * a.groups + cam) * batch_stride + loff, H, W, r.x + ox, r.y + oy, w,
                 ga, gx, gy);
        }
        finish_sample(s, w, ga, gx, gy);
      }
    }
  } else {
    for (int j = 0; j < a.groups; ++j) {
      const size_t boff = ((size_t)b * a.groups + j) * batch_stride;
      for (int l = 0; l < a.L; ++l) {
        const int H = lv.h[l], W = lv.w[l];
        const size_t loff = boff + (size_t)lv.start[l] * pix_stride;
        const float2 r = *reinterpret_cast<const float2*>(
            a.ref + ((((size_t)b * a.groups + j) * a.Nq + q) * a.L + l) * 2);
#pragma unroll 2
        for (int p = 0; p < a.P; ++p) {
          const int s = (j * a.L + l) * a.P + p;
          const float w = my_w[s];
          float ga = 0.f, gx = 0.f, gy = 0.f;
          if (active)
            sample(loff, H, W, r.x + my_off[2 * s], r.y + my_off[2 * s + 1], w, ga, gx, gy);
          finish_sample(s, w, ga, gx, gy);
        }
      }
    }
  }
  __syncthreads();
  // softmax backward per (row, segment): g_logit = a * (ga - sum_t a_t ga_t); zero where clamped
  const int nseg = rows_here * (S / LP);
  for (int sgi = tid; sgi < nseg; sgi += blockDim.x) {
    const float* w = s_w + (size_t)sgi * LP;
    const float* ga = s_ga + (size_t)sgi * LP;
    float dot = 0.f;
    for (int i = 0; i < LP; ++i) dot = fmaf(w[i], ga[i], dot);
    const size_t base = (size_t)row0 * S + (size_t)sgi * LP;
    for (int i = 0; i < LP; ++i) {
      float gl = w[i] * (ga[i] - dot);
      if (a.clamp >= 0.f) {
        const float raw = a.logits[base + i];
        if (raw < -a.clamp || raw > a.clamp) gl = 0.f;
      }
      a.g_logits[base + i] = gl;
    }
  }
}

template <typename T, int TPH, int MODE>
static int launch_fused(const FusedProblem& f, bool bwd, cudaStream_t st, const char* what) {
  constexpr int ROWS = kFusedThreads / TPH;
  FusedArgs a;
  a.value = f.value; a.shapes = f.shapes; a.starts = f.starts; a.offsets = f.offsets;
  a.logits = f.logits; a.ref = f.ref; a.bev_mask = f.bev_mask; a.hit_bits = f.hit_bits;
  a.out = f.out; a.g_out = f.g_out; a.g_value = f.g_value; a.g_offsets = f.g_offsets;
  a.g_logits = f.g_logits;
  a.bs = f.bs; a.groups = f.groups; a.Nk = f.Nk; a.M = f.M; a.Dh = f.Dh; a.L = f.L; a.P = f.P;
  a.D = f.D; a.Nq = f.Nq; a.clamp = f.clamp;
  a.total_rows = (long long)f.bs * f.Nq * f.M;
  const long long grid = (a.total_rows + ROWS - 1) / ROWS;
  if (grid <= 0) return MSDA_OK;
  if (grid > 0x7fffffffLL) return set_error(MSDA_ERR_UNSUPPORTED, "%s: problem too large for one launch", what);
  const int S = (MODE == MODE_TSA ? f.groups : 1) * f.L * f.P;
  const size_t smem = (size_t)ROWS * S * (bwd ? 4 : 3) * sizeof(float);
  if (smem > 200 * 1024)
    return set_error(MSDA_ERR_UNSUPPORTED, "%s: %d samples per row need %zu bytes of shared memory", what, S, smem);
  cudaError_t e = cudaSuccess;
  if (bwd) {
    if (smem > 48 * 1024)
      e = cudaFuncSetAttribute(fused_bwd_kernel<T, TPH, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) fused_bwd_kernel<T, TPH, MODE><<<(unsigned)grid, kFusedThreads, smem, st>>>(a);
  } else {
    if (smem > 48 * 1024)
      e = cudaFuncSetAttribute(fused_fwd_kernel<T, TPH, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) fused_fwd_kernel<T, TPH, MODE><<<(unsigned)grid, kFusedThreads, smem, st>>>(a);
  }
  if (e != cudaSuccess) return set_error(MSDA_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  count_launch();
  return check_launch(what);
}

template <typename T, int MODE>
static int dispatch_tph(const FusedProblem& f, bool bwd, cudaStream_t st, const char* what) {
  constexpr int VEC = Vec16<T>::N;
  if (f.Dh % VEC != 0)
    return set_error(MSDA_ERR_UNSUPPORTED, "%s: head_dim %d must be a multiple of %d for this dtype "
                     "(use the op-boundary msda_fwd/msda_bwd, which has a generic path)", what, f.Dh, VEC);
  switch (f.Dh / VEC) {
    case 1: return launch_fused<T, 1, MODE>(f, bwd, st, what);
    case 2: return launch_fused<T, 2, MODE>(f, bwd, st, what);
    case 4: return launch_fused<T, 4, MODE>(f, bwd, st, what);
    case 8: return launch_fused<T, 8, MODE>(f, bwd, st, what);
    case 16: return launch_fused<T, 16, MODE>(f, bwd, st, what);
    default: break;
  }
  return set_error(MSDA_ERR_UNSUPPORTED, "%s: head_dim %d not supported by the fused kernels", what, f.Dh);
}

template <int MODE>
static int dispatch_dtype(const FusedProblem& f, bool bwd, cudaStream_t st, const char* what) {
  const uintptr_t al = reinterpret_cast<uintptr_t>(f.value) | reinterpret_cast<uintptr_t>(f.out) |
                       reinterpret_cast<uintptr_t>(f.g_out) | reinterpret_cast<uintptr_t>(f.g_value) |
                       reinterpret_cast<uintptr_t>(f.g_offsets) | reinterpret_cast<uintptr_t>(f.ref);
  if (al & 15u) return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: tensors must be 16-byte aligned", what);
  switch (f.value_dtype) {
    case MSDA_F32: return dispatch_tph<float, MODE>(f, bwd, st, what);
    case MSDA_BF16: return dispatch_tph<__nv_bfloat16, MODE>(f, bwd, st, what);
    default: return dispatch_tph<__half, MODE>(f, bwd, st, what);
  }
}

int launch_sca_fwd(const FusedProblem& f, cudaStream_t st) { return dispatch_dtype<MODE_SCA>(f, false, st, "sca_fwd"); }
int launch_sca_bwd(const FusedProblem& f, cudaStream_t st) { return dispatch_dtype<MODE_SCA>(f, true, st, "sca_bwd"); }
int launch_tsa_fwd(const FusedProblem& f, cudaStream_t st) { return dispatch_dtype<MODE_TSA>(f, false, st, "tsa_fwd"); }
int launch_tsa_bwd(const FusedProblem& f, cudaStream_t st) { return dispatch_dtype<MODE_TSA>(f, true, st, "tsa_bwd"); }

}  // namespace msda
