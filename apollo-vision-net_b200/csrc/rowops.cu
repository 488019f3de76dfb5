// Row-wise companions of the attention kernels in a BEVFormer layer (SURVEY.md section 8f rank 3:
// "encoder remainder"): LayerNorm forward / backward and the column sums that give the bias
// gradients of the Linear layers.  All three are pure HBM streams over (rows, C) activations
// (40 000 x 256 at the base config); they exist because the stock PyTorch kernels for exactly
// these shapes (bf16, many short rows) were the largest non-attention items of the measured step
// (profiles/r01_launches.md).
//
// LayerNorm follows torch.nn.LayerNorm (the reference builds its norms through mmcv's
// build_norm_layer(dict(type='LN')), custom_base_transformer_layer.py:156-161): biased variance,
// eps inside the square root, affine weight / bias, statistics in fp32.
#include "msda_common.cuh"
#include "msda_host.h"
#include "philox.cuh"

namespace msda {

constexpr int kRowThreads = 256;
#ifndef LN_BWD_MINBLOCKS
#define LN_BWD_MINBLOCKS 3
#endif
#ifndef LN_BWD_MINBLOCKS_DROP
#define LN_BWD_MINBLOCKS_DROP 2
#endif
// Rows a warp keeps in flight ahead of the one it is working on (register stages).  A warp's row is only
// 1 KB of loads (x + residual, or x + dy), so one row ahead leaves ~32 KB in flight per SM -- below what the
// HBM latency x bandwidth product asks for; see profiles/r02_rowops.md for the sweep.
#ifndef LN_BWD_HOIST_GAMMA
#define LN_BWD_HOIST_GAMMA 0
#endif
#ifndef LN_FWD_DEPTH
#define LN_FWD_DEPTH 1
#endif
#ifndef LN_FWD_MINBLOCKS
#define LN_FWD_MINBLOCKS 4
#endif
#ifndef LN_BWD_DEPTH
#define LN_BWD_DEPTH 1
#endif

template <typename T> struct Vec16IO {
  static constexpr int N = Vec16<T>::N;
  __device__ __forceinline__ static void load(const T* p, float (&f)[N]) { Vec16<T>::unpack(ldg128(p), f); }
  __device__ __forceinline__ static void store(T* p, const float (&f)[N]) {
    *reinterpret_cast<uint4*>(p) = Vec16<T>::pack(f);
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- dropout: counter-based masks, recomputed in the backward (no mask tensor) -------------------
// The reference trains with dropout 0.1 after the output projections of TSA / SCA and twice inside the
// FFN (temporal_self_attention.py:285-289, spatial_cross_attention.py:171-173, mmcv FFN).  Element e of
// a tensor is kept when the 16-bit lane (e % 8) of Philox4x32-10(counter = (e / 8, call site, step),
// key = seed) is >= round(p * 65536); kept values are scaled by 1 / (1 - p).  (seed, step) live in a
// device buffer (the step advances on the device once per training step, so a replayed CUDA graph draws
// new masks); the forward kernel copies the pair it used into `key_save`, which the backward reads.
struct DropArgs {
  const unsigned long long* key_src;   // (seed, step) on the device; NULL = no dropout
  unsigned long long* key_save;        // forward: where block 0 stores the pair (may be NULL)
  uint32_t site;                       // call site (distinct per dropout instance inside a step)
  uint32_t thresh;                     // round(p * 65536)
  float scale;                         // 1 / (1 - p)
};
// (p < 1 is checked by the callers; a p within 2^-17 of 1 would round to 65536, which no 16-bit lane reaches)
static inline uint32_t drop_thresh(float p) {
  const uint32_t t = (uint32_t)(p * 65536.f + 0.5f);
  return t > 65535u ? 65535u : t;
}

// keep bits (bit j = element 8 * group + j is kept)
__device__ __forceinline__ uint32_t drop_keep8(unsigned long long group, unsigned long long seed,
                                               unsigned long long step, uint32_t site, uint32_t thresh) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)group, (uint32_t)(group >> 32), site, (uint32_t)step),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32)));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  uint32_t keep = 0u;
#pragma unroll
  for (int j = 0; j < 8; ++j) keep |= (((w[j >> 1] >> (16 * (j & 1))) & 0xffffu) >= thresh ? 1u : 0u) << j;
  return keep;
}
// the same decisions as predicates, straight from the four words (lane j of word w is kept when its 16 bits
// are >= thresh: the high lane compares the whole word against thresh << 16, the low lane the word shifted up)
template <int VEC>
__device__ __forceinline__ void drop_keep_flags(unsigned long long e0, unsigned long long seed, unsigned long long step,
                                                const uint32_t site, const uint32_t thresh, bool (&keep)[VEC]) {
  const unsigned long long group = e0 >> 3;
  const uint4 r = philox4x32_10(make_uint4((uint32_t)group, (uint32_t)(group >> 32), site, (uint32_t)step),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32)));
  const uint32_t t16 = thresh << 16;                 // thresh <= 65535 (p < 1)
  if (VEC == 8) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      keep[2 * j] = (w[j] << 16) >= t16;
      keep[2 * j + 1] = w[j] >= t16;
    }
  } else {
    const bool hi = (e0 & 4) != 0;
    const uint32_t w0 = hi ? r.z : r.x, w1 = hi ? r.w : r.y;
    keep[0] = (w0 << 16) >= t16;
    keep[1] = w0 >= t16;
    keep[2] = (w1 << 16) >= t16;
    keep[3] = w1 >= t16;
  }
}
// keep bits of the VEC (4 or 8) consecutive elements starting at element e0 (a multiple of VEC)
template <int VEC>
__device__ __forceinline__ uint32_t drop_keep(unsigned long long e0, unsigned long long seed, unsigned long long step,
                                              const DropArgs& d) {
  const uint32_t k = drop_keep8(e0 >> 3, seed, step, d.site, d.thresh);
  return VEC == 8 ? k : (k >> (e0 & 4)) & 0xfu;
}

// Final stage of the column reductions.  Every CTA adds its partial sums into one of kStrips fp32 strips in
// the workspace with reductions at L2 (the strips are all zero on entry); the LAST CTA to finish
// (ticket counter in the workspace header) sums the strips into the output and zeroes strips and
// counter again, so the workspace is reusable without a memset.  One launch, no serial tail.
// Several strips because reductions on one line serialise in L2: with ~600 CTAs adding into a single
// strip the last of them waited ~5 us for the queue in front of it (profiles/r02_rowops.md).
constexpr int kWsHeaderFloats = 64;          // 256-byte header in front of the strips
constexpr int kStrips = 8;                   // workspace: kWsHeaderFloats + kStrips * ncols floats

__device__ __forceinline__ float* my_strip(float* ws, unsigned cta, int ncols) {
  return ws + kWsHeaderFloats + (size_t)(cta % kStrips) * ncols;
}

template <typename TO>
__device__ __forceinline__ void finalize_columns(float* __restrict__ ws, TO* __restrict__ out, int ncols,
                                                 unsigned total_ctas) {
  __shared__ bool last;
  // the CTA barrier orders every thread's strip reductions before thread 0's fence (fences are
  // cumulative), which orders them before the ticket: one device-scope fence per CTA, not one per thread
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned* counter = reinterpret_cast<unsigned*>(ws);
    const unsigned ticket = atomicAdd(counter, 1u);
    last = (ticket == total_ctas - 1);
    if (last) *counter = 0u;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float* strips = ws + kWsHeaderFloats;
  for (int c = threadIdx.x; c < ncols; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kStrips; ++k) {
      s += __ldcg(strips + (size_t)k * ncols + c);
      strips[(size_t)k * ncols + c] = 0.f;
    }
    out[c] = from_f32<TO>(s);
  }
}

// One warp per row; every lane keeps its slice of the row (C / 32 elements, <= 32) in registers.
// With `residual`: normalises s = x + residual, rounded to T first (bit-identical to a separate
// add kernel followed by LayerNorm), and writes s to `sum_out` (may alias x) for the backward.
// DROP: s = dropout(x) + residual (x = the output of the block's last Linear).
// The loads of a warp's next DEPTH rows are in flight (register stages) while it works on the current one.
template <typename T, int PER_LANE, bool DROP>
__global__ void __launch_bounds__(kRowThreads, PER_LANE <= 8 ? LN_FWD_MINBLOCKS : 1)
ln_fwd_kernel(const T* x, const T* __restrict__ residual, const T* __restrict__ gamma,
              const T* __restrict__ beta, T* sum_out, T* __restrict__ y, float* __restrict__ mean,
              float* __restrict__ rstd, long long rows, int C, float eps, const DropArgs drop) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int NV = PER_LANE / VEC;
  constexpr int DEPTH = PER_LANE <= 8 ? LN_FWD_DEPTH : 1;
  const int lane = threadIdx.x & 31;
  unsigned long long seed = 0ull, step = 0ull;
  if (DROP) {
    seed = drop.key_src[0];
    step = drop.key_src[1];
    if (drop.key_save != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
      drop.key_save[0] = seed;
      drop.key_save[1] = step;
    }
  }
  const long long warp = (long long)blockIdx.x * (kRowThreads / 32) + (threadIdx.x >> 5);
  const long long nwarp = (long long)gridDim.x * (kRowThreads / 32);
  constexpr int NVA = NV > 0 ? NV : 1;      // (NV == 0 instances are rejected at dispatch)
  uint4 gpk[NVA], bpk[NVA];                 // gamma / beta stay packed (registers), unpacked per row
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    gpk[v] = ldg128(gamma + (v * 32 + lane) * VEC);
    bpk[v] = ldg128(beta + (v * 32 + lane) * VEC);
  }
  const float inv_c = 1.0f / (float)C;      // (C is a multiple of 128: the quotient below stays the rounded one
                                            //  only when C is a power of two -- see mean_of)
  const bool pow2 = (C & (C - 1)) == 0;
  auto mean_of = [&](float s) { return pow2 ? s * inv_c : s / (float)C; };
  uint4 nx[DEPTH][NVA], nr[DEPTH][NVA];
  auto fetch = [&](int st, long long r) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      nx[st][v] = *reinterpret_cast<const uint4*>(x + r * C + (v * 32 + lane) * VEC);   // x may alias sum_out
      if (residual != nullptr) nr[st][v] = ldg128(residual + r * C + (v * 32 + lane) * VEC);
    }
  };
#pragma unroll
  for (int st = 0; st < DEPTH; ++st)
    if (warp + st * nwarp < rows) fetch(st, warp + st * nwarp);
  for (long long base = warp; base < rows; base += DEPTH * nwarp) {
#pragma unroll
    for (int st = 0; st < DEPTH; ++st) {
      const long long r = base + st * nwarp;
      if (r >= rows) break;
      uint4 cx[NVA], cr[NVA];
#pragma unroll
      for (int v = 0; v < NV; ++v) { cx[v] = nx[st][v]; cr[v] = nr[st][v]; }
      if (r + DEPTH * nwarp < rows) fetch(st, r + DEPTH * nwarp);
      float f[PER_LANE];
      float s = 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        float t[VEC];
        Vec16<T>::unpack(cx[v], t);
        if (residual != nullptr) {
          float rr[VEC];
          Vec16<T>::unpack(cr[v], rr);
          if (DROP) {        // (rounded to T like a separate dropout kernel's output)
            bool keep[VEC];
            drop_keep_flags<VEC>((unsigned long long)r * C + (v * 32 + lane) * VEC, seed, step, drop.site,
                                 drop.thresh, keep);
#pragma unroll
            for (int i = 0; i < VEC; ++i) t[i] = keep[i] ? to_f32<T>(from_f32<T>(t[i] * drop.scale)) : 0.f;
          }
#pragma unroll
          for (int i = 0; i < VEC; ++i) t[i] = to_f32<T>(from_f32<T>(t[i] + rr[i]));
          Vec16IO<T>::store(sum_out + r * C + (v * 32 + lane) * VEC, t);
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) { f[v * VEC + i] = t[i]; s += t[i]; }
      }
      const float mu = mean_of(warp_sum(s));
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < PER_LANE; ++i) { const float d = f[i] - mu; q = fmaf(d, d, q); }
      const float rs = rsqrtf(mean_of(warp_sum(q)) + eps);
      T* yr = y + r * C;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        float t[VEC], g[VEC], bb[VEC];
        Vec16<T>::unpack(gpk[v], g);
        Vec16<T>::unpack(bpk[v], bb);
#pragma unroll
        for (int i = 0; i < VEC; ++i) t[i] = (f[v * VEC + i] - mu) * rs * g[i] + bb[i];
        Vec16IO<T>::store(yr + (v * 32 + lane) * VEC, t);
      }
      if (lane == 0) { mean[r] = mu; rstd[r] = rs; }
    }
  }
}

// dx per row; per-CTA partial sums of dgamma / dbeta into `partial` (gridDim.x, 2, C) fp32.
// DXSUM: additionally the column sums of dx (as rounded to T) -- the bias gradient of a Linear
// layer whose output (plus a residual) this LayerNorm normalised; output rows: d gamma, d beta,
// sum of dx.
// DROP (with DXSUM): the normalised sum was dropout(lin) + residual, so the gradient of the Linear output
// is dx masked and scaled -- written to `dx_masked` and summed over the rows instead of dx.
template <typename T, int PER_LANE, bool DXSUM, bool DROP>
__global__ void __launch_bounds__(kRowThreads, PER_LANE <= 8 ? (DROP ? LN_BWD_MINBLOCKS_DROP : LN_BWD_MINBLOCKS) : 1)
ln_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, const T* __restrict__ gamma,
              const float* __restrict__ mean, const float* __restrict__ rstd, T* __restrict__ dx,
              float* __restrict__ ws, T* __restrict__ dgamma_dbeta, long long rows, int C,
              T* __restrict__ dx_masked, const DropArgs drop) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int NV = PER_LANE / VEC;
  constexpr int NOUT = DXSUM ? 3 : 2;
  extern __shared__ __align__(16) float sm[];                       // [warps][NOUT][C]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long warp = (long long)blockIdx.x * (kRowThreads / 32) + wid;
  const long long nwarp = (long long)gridDim.x * (kRowThreads / 32);
  float dg[PER_LANE], db[PER_LANE], dsum[DXSUM ? PER_LANE : 1];
  unsigned long long seed = 0ull, step = 0ull;
  if (DROP) {
    seed = drop.key_src[0];
    step = drop.key_src[1];
  }
  constexpr int NVG = NV > 0 ? NV : 1;
  uint4 gpk[NVG];                                     // gamma stays packed (registers), unpacked per row
#pragma unroll
  for (int v = 0; v < NV; ++v) gpk[v] = ldg128(gamma + (v * 32 + lane) * VEC);
  constexpr bool HOIST = LN_BWD_HOIST_GAMMA != 0 && PER_LANE <= 8;      // ... or once, when registers allow
  float gf[HOIST ? PER_LANE : 1];
  if (HOIST) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float t[VEC];
      Vec16<T>::unpack(gpk[v], t);
#pragma unroll
      for (int i = 0; i < VEC; ++i) gf[HOIST ? v * VEC + i : 0] = t[i];
    }
  }
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) { dg[i] = 0.f; db[i] = 0.f; }
#pragma unroll
  for (int i = 0; i < (DXSUM ? PER_LANE : 1); ++i) dsum[i] = 0.f;
  constexpr int NVA = NV > 0 ? NV : 1;
  constexpr int DEPTH = PER_LANE <= 8 ? LN_BWD_DEPTH : 1;
  const float inv_c = 1.0f / (float)C;
  const bool pow2 = (C & (C - 1)) == 0;               // exact reciprocal: the product equals the quotient
  auto mean_of = [&](float s) { return pow2 ? s * inv_c : s / (float)C; };
  uint4 nx[DEPTH][NVA], nd[DEPTH][NVA];
  float nmu[DEPTH], nrs[DEPTH];
  auto fetch = [&](int st, long long r) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      nx[st][v] = ldg128(x + r * C + (v * 32 + lane) * VEC);
      nd[st][v] = ldg128(dy + r * C + (v * 32 + lane) * VEC);
    }
    nmu[st] = __ldg(mean + r);
    nrs[st] = __ldg(rstd + r);
  };
#pragma unroll
  for (int st = 0; st < DEPTH; ++st)
    if (warp + st * nwarp < rows) fetch(st, warp + st * nwarp);
  for (long long base = warp; base < rows; base += DEPTH * nwarp) {
#pragma unroll
    for (int st = 0; st < DEPTH; ++st) {
      const long long r = base + st * nwarp;
      if (r >= rows) break;
      const float mu = nmu[st], rs = nrs[st];
      uint4 cx[NVA], cd[NVA];
#pragma unroll
      for (int v = 0; v < NV; ++v) { cx[v] = nx[st][v]; cd[v] = nd[st][v]; }
      if (r + DEPTH * nwarp < rows) fetch(st, r + DEPTH * nwarp);
      float xh[PER_LANE], gd[PER_LANE];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        float tx[VEC], td[VEC], tg[VEC];
        Vec16<T>::unpack(cx[v], tx);
        Vec16<T>::unpack(cd[v], td);
        if (HOIST) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) tg[i] = gf[HOIST ? v * VEC + i : 0];
        } else {
          Vec16<T>::unpack(gpk[v], tg);
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const int k = v * VEC + i;
          xh[k] = (tx[i] - mu) * rs;
          gd[k] = td[i] * tg[i];
          dg[k] = fmaf(td[i], xh[k], dg[k]);
          db[k] += td[i];
          s1 += gd[k];
          s2 = fmaf(gd[k], xh[k], s2);
        }
      }
      s1 = mean_of(warp_sum(s1));
      s2 = mean_of(warp_sum(s2));
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        float t[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const int k = v * VEC + i;
          t[i] = rs * (gd[k] - s1 - xh[k] * s2);
          if (DXSUM && !DROP) dsum[k] += to_f32<T>(from_f32<T>(t[i]));      // what a column sum of dx would read
        }
        Vec16IO<T>::store(dx + r * C + (v * 32 + lane) * VEC, t);
        if (DXSUM && DROP) {
          bool keep[VEC];
          drop_keep_flags<VEC>((unsigned long long)r * C + (v * 32 + lane) * VEC, seed, step, drop.site,
                               drop.thresh, keep);
          float u[VEC];
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            u[i] = keep[i] ? to_f32<T>(from_f32<T>(to_f32<T>(from_f32<T>(t[i])) * drop.scale)) : 0.f;
            dsum[v * VEC + i] += u[i];
          }
          Vec16IO<T>::store(dx_masked + r * C + (v * 32 + lane) * VEC, u);
        }
      }
    }
  }
  // CTA reduction of the per-warp partials, then one row of partials per CTA
  float* mine = sm + (size_t)wid * NOUT * C;
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int c = (v * 32 + lane) * VEC + i;
      mine[c] = dg[v * VEC + i];
      mine[C + c] = db[v * VEC + i];
      if (DXSUM) mine[2 * C + c] = dsum[v * VEC + i];
    }
  __syncthreads();
  // four columns per thread and one 16-byte vector reduction into the strip (C is a multiple of 128)
  for (int c = threadIdx.x * 4; c < NOUT * C; c += kRowThreads * 4) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < kRowThreads / 32; ++w) {
      const float4 v = *reinterpret_cast<const float4*>(sm + (size_t)w * NOUT * C + c);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    red_add_f32x4(my_strip(ws, blockIdx.x, NOUT * C) + c, s.x, s.y, s.z, s.w);
  }
  finalize_columns<T>(ws, dgamma_dbeta, NOUT * C, gridDim.x);
}

// Column sums of a (rows, C) matrix: every CTA reduces its rows, adds into the strip, the last CTA
// converts.  A thread owns 16 bytes of columns; a CTA's threads cover C/VEC column groups x
// (256 / (C/VEC)) row lanes.
// RELU: the matrix summed is dx = (y > 0 ? x : 0) -- the backward of ReLU with upstream gradient x and
// forward output y -- and dx is written out as well: one pass gives the activation's input gradient and
// the bias gradient of the Linear layer in front of it.
// `relu_scale`: factor on the kept entries (1 / (1 - p) when a dropout was fused into the activation: y is
// then zero where the unit was dropped, too, so y > 0 selects exactly the surviving entries).
template <typename T, typename TO, bool RELU>
__global__ void __launch_bounds__(kRowThreads)
colsum_kernel(const T* __restrict__ x, const T* __restrict__ y, T* __restrict__ dx, float* __restrict__ ws,
              TO* __restrict__ out, long long rows, int C, float relu_scale) {
  constexpr int VEC = Vec16<T>::N;
  extern __shared__ __align__(16) float sm[];                       // [row_lanes][C]
  const int groups = C / VEC;                         // <= 256
  const int row_lanes = kRowThreads / groups;
  const int gidx = threadIdx.x % groups, rl = threadIdx.x / groups;
  float acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
  // (relu_scale == 1: the product is the value itself, already representable in T)
  auto finish = [&](long long r, const uint4& xr, const uint4& yr, float (&t)[VEC]) {
    Vec16<T>::unpack(xr, t);
    if (RELU) {
      float m[VEC];
      Vec16<T>::unpack(yr, m);
#pragma unroll
      for (int i = 0; i < VEC; ++i) t[i] = m[i] > 0.f ? to_f32<T>(from_f32<T>(t[i] * relu_scale)) : 0.f;
      Vec16IO<T>::store(dx + r * C + gidx * VEC, t);
    }
  };
  if (rl < row_lanes) {
    const long long step = (long long)gridDim.x * row_lanes;
    long long r = (long long)blockIdx.x * row_lanes + rl;
    // four rows per trip: all their loads (x, and y with RELU) are issued before the first is consumed
    for (; r + 3 * step < rows; r += 4 * step) {
      uint4 xr[4], yr[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        xr[k] = ldg128(x + (r + k * step) * C + gidx * VEC);
        if (RELU) yr[k] = ldg128(y + (r + k * step) * C + gidx * VEC);
        else yr[k] = make_uint4(0u, 0u, 0u, 0u);
      }
      float t0[VEC], t1[VEC], t2[VEC], t3[VEC];
      finish(r, xr[0], yr[0], t0);
      finish(r + step, xr[1], yr[1], t1);
      finish(r + 2 * step, xr[2], yr[2], t2);
      finish(r + 3 * step, xr[3], yr[3], t3);
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] += (t0[i] + t1[i]) + (t2[i] + t3[i]);
    }
    for (; r < rows; r += step) {
      float t[VEC];
      finish(r, ldg128(x + r * C + gidx * VEC), RELU ? ldg128(y + r * C + gidx * VEC) : make_uint4(0u, 0u, 0u, 0u), t);
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] += t[i];
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) sm[(size_t)rl * C + gidx * VEC + i] = acc[i];
  }
  __syncthreads();
  if (C % 4 == 0) {                                   // (always for 16-bit types: VEC = 8)
    for (int c = threadIdx.x * 4; c < C; c += kRowThreads * 4) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int w = 0; w < row_lanes; ++w) {
        const float4 v = *reinterpret_cast<const float4*>(sm + (size_t)w * C + c);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      red_add_f32x4(my_strip(ws, blockIdx.x, C) + c, s.x, s.y, s.z, s.w);
    }
  } else {
    for (int c = threadIdx.x; c < C; c += kRowThreads) {
      float s = 0.f;
      for (int w = 0; w < row_lanes; ++w) s += sm[(size_t)w * C + c];
      atomicAdd(my_strip(ws, blockIdx.x, C) + c, s);
    }
  }
  finalize_columns<TO>(ws, out, C, gridDim.x);
}

static int row_grid(int per_sm = 4) {
  int dev = 0, n = 148;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  return n * per_sm;
}

int rowops_partial_rows() { return row_grid(); }

// Forward: `residual` / `sum_out` optional (both or neither).  Backward: `dxsum` selects the
// three-row output (d gamma, d beta, column sums of dx).
template <typename T, int PER_LANE>
static int ln_launch(bool bwd, const void* x, const void* dy, const void* gamma, const void* beta, void* y,
                     float* mean, float* rstd, void* dx, float* partial, void* dgb, long long rows, int C,
                     float eps, const void* residual, void* sum_out, bool dxsum, const RowDropout& rd,
                     void* dx_masked, cudaStream_t st) {
  // one wave of CTAs: as many per SM as the kernel's launch bound keeps resident
  const long long need = (rows + kRowThreads / 32 - 1) / (kRowThreads / 32);
  const int fgrid_max = row_grid(PER_LANE <= 8 ? LN_FWD_MINBLOCKS : 4);
  const int grid = (int)(need < fgrid_max ? need : fgrid_max);
  if (grid <= 0) return MSDA_OK;
  const bool dropping = rd.key != nullptr && rd.p > 0.f;
  DropArgs da{};
  if (dropping) {
    da.key_src = static_cast<const unsigned long long*>(rd.key);
    da.key_save = static_cast<unsigned long long*>(rd.key_save);
    da.site = rd.site;
    da.thresh = drop_thresh(rd.p);
    da.scale = 1.f / (1.f - rd.p);
  }
  if (!bwd) {
    auto kfwd = dropping ? ln_fwd_kernel<T, PER_LANE, true> : ln_fwd_kernel<T, PER_LANE, false>;
    kfwd<<<grid, kRowThreads, 0, st>>>(
        static_cast<const T*>(x), static_cast<const T*>(residual), static_cast<const T*>(gamma),
        static_cast<const T*>(beta), static_cast<T*>(sum_out), static_cast<T*>(y), mean, rstd, rows, C, eps, da);
    count_launch();
    return check_launch("ln_fwd");
  }
  const size_t smem = (size_t)(kRowThreads / 32) * (dxsum ? 3 : 2) * C * sizeof(float);
  auto kfn = dxsum ? (dropping ? ln_bwd_kernel<T, PER_LANE, true, true> : ln_bwd_kernel<T, PER_LANE, true, false>)
                   : ln_bwd_kernel<T, PER_LANE, false, false>;
  if (smem + 1024 > 48 * 1024 &&        // (+ the kernel's static shared memory)
      cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return set_error(MSDA_ERR_CUDA, "ln_bwd: cannot reserve %zu bytes of shared memory", smem);
  // every CTA ends with a reduction of its column partials into the L2 strip: with few rows (the decoders'
  // 7 000 queries) fewer CTAs, eight rows per warp, keep that fixed cost from dominating
  // (4 / 8 / 16 / 32 rows per warp: MapTRv2 decoder step 4.11 / 4.02 / 4.45 / 5.00 ms -- 4 to 8 is flat)
  const long long want = (rows + 8 * (kRowThreads / 32) - 1) / (8 * (kRowThreads / 32));
  const int bgrid_max = row_grid(PER_LANE <= 8 ? (dropping ? LN_BWD_MINBLOCKS_DROP : LN_BWD_MINBLOCKS) : 4);
  const int bgrid = (int)(want < bgrid_max ? (want > 0 ? want : 1) : bgrid_max);
  kfn<<<bgrid, kRowThreads, smem, st>>>(
      static_cast<const T*>(x), static_cast<const T*>(dy), static_cast<const T*>(gamma), mean, rstd,
      static_cast<T*>(dx), partial, static_cast<T*>(dgb), rows, C, static_cast<T*>(dx_masked), da);
  count_launch();
  return check_launch("ln_bwd");
}

template <typename T>
static int ln_dispatch(bool bwd, const void* x, const void* dy, const void* gamma, const void* beta, void* y,
                       float* mean, float* rstd, void* dx, float* partial, void* dgb, long long rows, int C,
                       float eps, const void* residual, void* sum_out, bool dxsum, const RowDropout& rd,
                       void* dx_masked, cudaStream_t st) {
  constexpr int VEC = Vec16<T>::N;
  if (C % (32 * VEC) != 0 || C / 32 > 32)
    return set_error(MSDA_ERR_UNSUPPORTED, "layer_norm: C=%d must be a multiple of %d and <= 1024", C, 32 * VEC);
#define LN_CASE(n) case n: return ln_launch<T, n>(bwd, x, dy, gamma, beta, y, mean, rstd, dx, partial, dgb, \
                                                  rows, C, eps, residual, sum_out, dxsum, rd, dx_masked, st)
  switch (C / 32) {
    LN_CASE(4);
    LN_CASE(8);
    LN_CASE(16);
    LN_CASE(32);
    default: break;
  }
#undef LN_CASE
  return set_error(MSDA_ERR_UNSUPPORTED, "layer_norm: C=%d not supported (128, 256, 512 or 1024)", C);
}

// ---- fp16 gradient accumulator helpers -------------------------------------------------------
// ws layout (floats): [0] ticket counter, [1] running max bits, [16] scale out.
template <typename T>
__global__ void __launch_bounds__(256)
grad_scale_kernel(const T* __restrict__ g, long long n, float limit, float* __restrict__ ws,
                  uint4* __restrict__ zero, long long zero_chunks) {
  constexpr int VEC = Vec16<T>::N;
  // the accumulator the scale is for starts at zero: cleared here (16-byte chunks), beside the read-only
  // stream of the maximum, instead of by a fill launch of its own
  for (long long z = (long long)blockIdx.x * blockDim.x + threadIdx.x; z < zero_chunks;
       z += (long long)gridDim.x * blockDim.x)
    zero[z] = make_uint4(0u, 0u, 0u, 0u);
  float m = 0.f;
  const long long nv = n / VEC;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  auto take = [&](const uint4& raw) {
    float t[VEC];
    Vec16<T>::unpack(raw, t);
#pragma unroll
    for (int k = 0; k < VEC; ++k) m = fmaxf(m, fabsf(t[k]));
  };
  for (; i + 3 * stride < nv; i += 4 * stride) {             // four independent 16-byte loads in flight
    const uint4 r0 = ldg128(g + i * VEC), r1 = ldg128(g + (i + stride) * VEC);
    const uint4 r2 = ldg128(g + (i + 2 * stride) * VEC), r3 = ldg128(g + (i + 3 * stride) * VEC);
    take(r0);
    take(r1);
    take(r2);
    take(r3);
  }
  for (; i < nv; i += stride) take(ldg128(g + i * VEC));
  if (blockIdx.x == 0)
    for (long long i = nv * VEC + threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(to_f32<T>(g[i])));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  unsigned* wsu = reinterpret_cast<unsigned*>(ws);
  // one atomic per CTA (same-address atomics serialise in L2: one per warp cost ~5 us of a 13 us launch)
  __shared__ float wmax[8];
  __shared__ bool last;
  if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, wmax[w]);
    if (m > 0.f) atomicMax(wsu + 1, __float_as_uint(m));       // non-negative floats order like uints
    __threadfence();
    const unsigned ticket = atomicAdd(wsu, 1u);
    last = ticket == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    const float amax = __uint_as_float(atomicExch(wsu + 1, 0u));
    float s = 1.f;
    if (amax > 0.f && isfinite(amax)) s = exp2f(floorf(log2f(limit / amax)));
    ws[16] = s;
    wsu[0] = 0u;
  }
}

// out = acc / scale in the output dtype.  With tail replicas (see FusedArgs::g_tail): the last
// `tail_elems` elements of every `map_elems`-long value map also sum `copies` replica maps.
// COLSUM: additionally the column sums over all rows of the (rows, C) view of the output (C = 8 * cgroups
// columns: the bias gradient of the Linear layer that produced the value tensor), through the same
// strip + ticket scheme as colsum_kernel.  The grid's x-stride is a multiple of cgroups, so a thread
// always holds the same 8 columns.
template <typename TO, bool COLSUM>
__global__ void __launch_bounds__(256)
unscale_cast_kernel(const __half* __restrict__ acc, TO* __restrict__ out, const float* __restrict__ scale, long long n,
                    const __half* __restrict__ tail, int copies, long long map_elems, long long tail_elems,
                    float* __restrict__ ws, TO* __restrict__ colsum_out, int cgroups, int* __restrict__ overflow,
                    long long out_ld) {
  const float inv = 1.0f / __ldg(scale);             // power of two: exact
  const int cg_shift = (cgroups > 0 && (cgroups & (cgroups - 1)) == 0) ? __ffs(cgroups) - 1 : -1;
  bool bad = false;                                  // a saturated (inf / NaN) accumulator slot was seen
  // grid: x over the 16-byte chunks of a value map, y over the maps (one map of n elements
  // when there are no tail replicas)
  const long long maps = n / map_elems;
  const long long chunks = map_elems / 8, tail_first = (map_elems - tail_elems) / 8;
  float csum[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) csum[k] = 0.f;
  for (long long map = blockIdx.y; map < maps; map += gridDim.y) {
    const __half* a = acc + map * map_elems;
    TO* o = out + map * map_elems;
    const __half* t = copies > 0 ? tail + map * tail_elems - tail_first * 8 : nullptr;
    auto finish = [&](long long i, const uint4& raw) {
      float v[8];
      Vec16<__half>::unpack(raw, v);
      if (copies > 0 && i >= tail_first) {
        for (int c = 0; c < copies; ++c) {
          float u[8];
          Vec16<__half>::unpack(ldg128(t + (long long)c * maps * tail_elems + i * 8), u);
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] += u[k];
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        bad |= !(fabsf(v[k]) <= 3.0e38f);            // fp16 inf / NaN survive the conversion and the sums
        v[k] *= inv;
      }
      if (COLSUM) {
#pragma unroll
        for (int k = 0; k < 8; ++k) csum[k] += v[k];
      }
      // out_ld != 0: `out` is a column block of a wider matrix -- row r of the (n / C, C) view (C = 8 * cgroups)
      // starts at out + r * out_ld (the feature-gradient GEMMs of the hoisted value projections read the
      // blocks of all layers as one matrix)
      TO* dst = o + i * 8;
      if (out_ld != 0) {
        const long long cg = map * chunks + i;
        // (a 64-bit division per 16-byte chunk made the pass issue bound: 38 -> 59 us at 47 M elements)
        const long long row = cg_shift >= 0 ? (cg >> cg_shift) : cg / cgroups;
        const int col = cg_shift >= 0 ? (int)(cg & (cgroups - 1)) : (int)(cg % cgroups);
        dst = out + row * out_ld + col * 8;
      }
      if constexpr (sizeof(TO) == 2) {
        *reinterpret_cast<uint4*>(dst) = Vec16<TO>::pack(v);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) dst[k] = from_f32<TO>(v[k]);
      }
    };
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < chunks; i += 4 * stride) {       // four independent 16-byte loads in flight
      const uint4 r0 = ldg128(a + i * 8), r1 = ldg128(a + (i + stride) * 8);
      const uint4 r2 = ldg128(a + (i + 2 * stride) * 8), r3 = ldg128(a + (i + 3 * stride) * 8);
      finish(i, r0);
      finish(i + stride, r1);
      finish(i + 2 * stride, r2);
      finish(i + 3 * stride, r3);
    }
    for (; i < chunks; i += stride) finish(i, ldg128(a + i * 8));
  }
  if (blockIdx.x == 0 && blockIdx.y == 0)             // (only without replicas: map_elems need not divide by 8)
    for (long long i = maps * chunks * 8 + threadIdx.x; i < n; i += blockDim.x) {
      const float v = __half2float(acc[i]);
      bad |= !(fabsf(v) <= 3.0e38f);
      out[i] = from_f32<TO>(v * inv);
    }
  // sticky flag word: the host may poll it whenever it likes (no synchronisation on the good path)
  if (overflow != nullptr && __any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(overflow, 1);
  if constexpr (COLSUM) {
    // threads with the same (threadIdx.x % cgroups) own the same 8 columns (map_elems, the x-stride and 256
    // are multiples of 8 * cgroups)
    __shared__ float part[256][9];
#pragma unroll
    for (int k = 0; k < 8; ++k) part[threadIdx.x][k] = csum[k];
    __syncthreads();
    const int C = cgroups * 8;
    float* strip = my_strip(ws, blockIdx.y * gridDim.x + blockIdx.x, C);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const int grp = c >> 3, k = c & 7;
      float s = 0.f;
      for (int th = grp; th < 256; th += cgroups) s += part[th][k];
      atomicAdd(strip + c, s);
    }
    finalize_columns<TO>(ws, colsum_out, C, gridDim.x * gridDim.y);       // ticket over the whole (x, y) grid
  }
}

int launch_grad_scale(const void* g, long long n, int dtype, float limit, float* ws, void* zero, long long zero_bytes,
                      cudaStream_t st) {
  const long long zc = zero != nullptr ? zero_bytes / 16 : 0;
  const long long work = n / 8 > zc ? n / 8 : zc;
  const long long need = (work + 255) / 256;
  // a read-only stream: eight CTAs per SM keep enough 16-byte loads in flight (two per SM: 16.6 us, four: 13 us)
  const int grid = (int)(need < row_grid(8) ? (need > 0 ? need : 1) : row_grid(8));
  uint4* z = static_cast<uint4*>(zero);
  if (dtype == MSDA_F32) grad_scale_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(g), n, limit, ws, z, zc);
  else if (dtype == MSDA_BF16) grad_scale_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(g), n, limit, ws, z, zc);
  else grad_scale_kernel<__half><<<grid, 256, 0, st>>>(static_cast<const __half*>(g), n, limit, ws, z, zc);
  count_launch();
  return check_launch("grad_amax_scale");
}

template <typename TO>
static void unscale_launch(dim3 grid, cudaStream_t st, const __half* a, void* out, const float* scale, long long n,
                           const __half* t, int copies, long long map_elems, long long tail_elems, float* ws,
                           void* colsum_out, int cgroups, int* overflow, long long out_ld) {
  if (colsum_out)
    unscale_cast_kernel<TO, true><<<grid, 256, 0, st>>>(a, static_cast<TO*>(out), scale, n, t, copies, map_elems,
                                                         tail_elems, ws, static_cast<TO*>(colsum_out), cgroups, overflow, out_ld);
  else
    unscale_cast_kernel<TO, false><<<grid, 256, 0, st>>>(a, static_cast<TO*>(out), scale, n, t, copies, map_elems,
                                                          tail_elems, nullptr, nullptr, out_ld != 0 ? cgroups : 1, overflow, out_ld);
}

int launch_unscale_cast(const void* acc16, void* out, const float* scale, long long n, int out_dtype,
                        const void* tail, int copies, long long map_elems, long long tail_elems,
                        float* ws, void* colsum_out, int C, int* overflow, long long out_ld, cudaStream_t st) {
  const __half* a = static_cast<const __half*>(acc16);
  const __half* t = static_cast<const __half*>(tail);
  if (!t || copies <= 0) { t = nullptr; copies = 0; map_elems = n; tail_elems = 0; }
  const long long maps = n / map_elems;
  const long long need = (map_elems / 8 + 255) / 256;
  // with column sums every CTA ends with C reductions into the strip: a grid of ~4 CTAs per SM, as colsum
  const long long budget = colsum_out ? row_grid() : (long long)row_grid() * 4;
  long long gx = budget / maps;
  if (gx < 1) gx = 1;
  if (gx > need) gx = need > 0 ? need : 1;
  const dim3 grid((unsigned)gx, (unsigned)(maps < 65535 ? maps : 65535));
  const int cgroups = C / 8;
  if (out_dtype == MSDA_F32) unscale_launch<float>(grid, st, a, out, scale, n, t, copies, map_elems, tail_elems, ws, colsum_out, cgroups, overflow, out_ld);
  else if (out_dtype == MSDA_BF16) unscale_launch<__nv_bfloat16>(grid, st, a, out, scale, n, t, copies, map_elems, tail_elems, ws, colsum_out, cgroups, overflow, out_ld);
  else unscale_launch<__half>(grid, st, a, out, scale, n, t, copies, map_elems, tail_elems, ws, colsum_out, cgroups, overflow, out_ld);
  count_launch();
  return check_launch("unscale_cast");
}

int launch_ln(bool bwd, const void* x, const void* dy, const void* gamma, const void* beta, void* y,
              float* mean, float* rstd, void* dx, void* dgamma_dbeta, float* partial,
              long long rows, int C, float eps, int dtype, const void* residual, void* sum_out,
              bool dxsum, const RowDropout& rd, void* dx_masked, cudaStream_t st) {
  if (rd.key != nullptr && rd.p > 0.f) {
    if (!(rd.p < 1.f)) return set_error(MSDA_ERR_BAD_ARGUMENT, "layer_norm: dropout probability %g out of [0, 1)", rd.p);
    if (!bwd && residual == nullptr)
      return set_error(MSDA_ERR_BAD_ARGUMENT, "layer_norm: the fused dropout acts on the residual form only");
    if (bwd && (!dxsum || dx_masked == nullptr))
      return set_error(MSDA_ERR_BAD_ARGUMENT, "layer_norm backward: dropout needs the dx-sum form and dx_masked");
  }
  if (dtype == MSDA_F32)
    return ln_dispatch<float>(bwd, x, dy, gamma, beta, y, mean, rstd, dx, partial, dgamma_dbeta, rows, C, eps,
                              residual, sum_out, dxsum, rd, dx_masked, st);
  if (dtype == MSDA_BF16)
    return ln_dispatch<__nv_bfloat16>(bwd, x, dy, gamma, beta, y, mean, rstd, dx, partial, dgamma_dbeta, rows, C,
                                      eps, residual, sum_out, dxsum, rd, dx_masked, st);
  return ln_dispatch<__half>(bwd, x, dy, gamma, beta, y, mean, rstd, dx, partial, dgamma_dbeta, rows, C, eps,
                             residual, sum_out, dxsum, rd, dx_masked, st);
}

// ---- ReLU + dropout in place (the FFN's first activation), and the mask itself (test hook) ----
template <typename T, bool RELU>
__global__ void __launch_bounds__(256)
relu_dropout_kernel(T* __restrict__ x, unsigned char* __restrict__ mask_out, long long n, const DropArgs drop) {
  constexpr int VEC = Vec16<T>::N;
  const unsigned long long seed = drop.key_src[0], step = drop.key_src[1];
  if (drop.key_save != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    drop.key_save[0] = seed;
    drop.key_save[1] = step;
  }
  const long long chunks = n / VEC, stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (mask_out != nullptr) {                       // test hook: the keep mask itself
    for (; i < chunks; i += stride) {
      const uint32_t keep = drop_keep<VEC>((unsigned long long)i * VEC, seed, step, drop);
#pragma unroll
      for (int k = 0; k < VEC; ++k) mask_out[i * VEC + k] = (unsigned char)((keep >> k) & 1u);
    }
    return;
  }
  auto finish = [&](long long c, const uint4& raw) {
    bool keep[VEC];
    drop_keep_flags<VEC>((unsigned long long)c * VEC, seed, step, drop.site, drop.thresh, keep);
    float t[VEC];
    Vec16<T>::unpack(raw, t);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float v = RELU ? fmaxf(t[k], 0.f) : t[k];
      t[k] = keep[k] ? v * drop.scale : 0.f;
    }
    *reinterpret_cast<uint4*>(x + c * VEC) = Vec16<T>::pack(t);
  };
  for (; i + 3 * stride < chunks; i += 4 * stride) {         // four independent 16-byte loads in flight
    const uint4 r0 = *reinterpret_cast<const uint4*>(x + i * VEC);
    const uint4 r1 = *reinterpret_cast<const uint4*>(x + (i + stride) * VEC);
    const uint4 r2 = *reinterpret_cast<const uint4*>(x + (i + 2 * stride) * VEC);
    const uint4 r3 = *reinterpret_cast<const uint4*>(x + (i + 3 * stride) * VEC);
    finish(i, r0);
    finish(i + stride, r1);
    finish(i + 2 * stride, r2);
    finish(i + 3 * stride, r3);
  }
  for (; i < chunks; i += stride) finish(i, *reinterpret_cast<const uint4*>(x + i * VEC));
}

int launch_relu_dropout(void* x, void* mask_out, long long n, int dtype, const RowDropout& rd, cudaStream_t st) {
  if (!(rd.p > 0.f && rd.p < 1.f) || rd.key == nullptr)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "relu_dropout: needs 0 < p < 1 and a key");
  const int vec = dtype == MSDA_F32 ? 4 : 8;
  if (n % 8 != 0) return set_error(MSDA_ERR_UNSUPPORTED, "relu_dropout: the element count must be a multiple of 8");
  DropArgs da{};
  da.key_src = static_cast<const unsigned long long*>(rd.key);
  da.key_save = static_cast<unsigned long long*>(rd.key_save);
  da.site = rd.site;
  da.thresh = drop_thresh(rd.p);
  da.scale = 1.f / (1.f - rd.p);
  const long long chunks = n / vec;
  if (chunks == 0) return MSDA_OK;
  long long grid = (chunks + 255) / 256;
  if (grid > (long long)row_grid() * 4) grid = (long long)row_grid() * 4;
  unsigned char* mo = static_cast<unsigned char*>(mask_out);
  if (dtype == MSDA_F32) relu_dropout_kernel<float, true><<<(unsigned)grid, 256, 0, st>>>(static_cast<float*>(x), mo, n, da);
  else if (dtype == MSDA_BF16) relu_dropout_kernel<__nv_bfloat16, true><<<(unsigned)grid, 256, 0, st>>>(static_cast<__nv_bfloat16*>(x), mo, n, da);
  else relu_dropout_kernel<__half, true><<<(unsigned)grid, 256, 0, st>>>(static_cast<__half*>(x), mo, n, da);
  count_launch();
  return check_launch("relu_dropout");
}

template <typename T, bool RELU>
static int colsum_out(const void* x, const void* y, void* dx, void* out, float* ws, long long rows, int C,
                      int out_dtype, int grid, size_t smem, float relu_scale, cudaStream_t st) {
  const T* xi = static_cast<const T*>(x);
  const T* yi = static_cast<const T*>(y);
  T* di = static_cast<T*>(dx);
  if (out_dtype == MSDA_F32) colsum_kernel<T, float, RELU><<<grid, kRowThreads, smem, st>>>(xi, yi, di, ws, static_cast<float*>(out), rows, C, relu_scale);
  else if (out_dtype == MSDA_BF16) colsum_kernel<T, __nv_bfloat16, RELU><<<grid, kRowThreads, smem, st>>>(xi, yi, di, ws, static_cast<__nv_bfloat16*>(out), rows, C, relu_scale);
  else colsum_kernel<T, __half, RELU><<<grid, kRowThreads, smem, st>>>(xi, yi, di, ws, static_cast<__half*>(out), rows, C, relu_scale);
  count_launch();
  return check_launch(RELU ? "relu_bwd_colsum" : "colsum");
}

// y == nullptr: plain column sums of x.  Otherwise dx = relu'(y) * x is written and summed.
int launch_colsum(const void* x, const void* y, void* dx, void* out, float* partial, long long rows, int C,
                  int dtype, int out_dtype, float relu_scale, cudaStream_t st) {
  const int vec = dtype == MSDA_F32 ? 4 : 8;
  if (C % vec != 0 || C / vec > kRowThreads)
    return set_error(MSDA_ERR_UNSUPPORTED, "colsum: C=%d must be a multiple of %d and <= %d", C, vec, vec * kRowThreads);
  const int groups = C / vec, row_lanes = kRowThreads / groups;
  const long long need = (rows + row_lanes - 1) / row_lanes;
  const int grid = (int)(need < row_grid() ? (need > 0 ? need : 1) : row_grid());
  const size_t smem = (size_t)row_lanes * C * sizeof(float);
  const bool relu = y != nullptr;
#define COLSUM_CASE(T) (relu ? colsum_out<T, true>(x, y, dx, out, partial, rows, C, out_dtype, grid, smem, relu_scale, st) \
                             : colsum_out<T, false>(x, y, dx, out, partial, rows, C, out_dtype, grid, smem, 1.f, st))
  if (dtype == MSDA_F32) return COLSUM_CASE(float);
  if (dtype == MSDA_BF16) return COLSUM_CASE(__nv_bfloat16);
  return COLSUM_CASE(__half);
#undef COLSUM_CASE
}

}  // namespace msda
