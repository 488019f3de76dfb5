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
// Work decomposition (both, forward and backward).  A "row" is one (batch, query, head); the TPH
// lanes that own a row (16 bytes of the head's channels each) are always inside one warp, and a
// warp is autonomous: there is no block-level barrier after the level table has been staged.
// Rows are mapped to queries through tiles -- a TW x TH patch of the BEV grid when the queries lie
// on one, so that the rows of a CTA sample neighbouring pixels and share L1 lines.  Per row:
//   1. the row's lanes fetch its raw Linear outputs (offsets, logits) with coalesced loads and
//      turn them, once per row (not once per lane or per camera), into softmax weights and
//      normalised offsets held in a padded (bank-conflict-free) shared-memory row;
//   2. every lane walks the row's samples, camera outermost, one unpredicated 128-bit read-only
//      load per (clamped) bilinear corner, packed FFMA2 accumulation in fp32.
// The backward additionally reduces the location / weight gradients over the row's lanes with
// warp shuffles, accumulates them over cameras in the shared row, finishes the softmax backward
// there, writes both gradient rows with 16-byte stores, and sends grad_value to an fp32
// accumulator with 16-byte vector reductions (red.global.add.v4.f32) laid out so that one warp
// instruction covers whole 32-byte sectors.
#include <cstdlib>
#include <type_traits>
#include "coarse_common.cuh"
#include "msda_common.cuh"
#include "msda_host.h"

namespace msda {

constexpr int kFusedThreads = 256;
#ifndef FUSED_BWD_UNROLL
#define FUSED_BWD_UNROLL 2
#endif
#ifndef FUSED_FWD_MINBLOCKS
#define FUSED_FWD_MINBLOCKS 4
#endif
#ifndef FUSED_BWD_MINBLOCKS
#define FUSED_BWD_MINBLOCKS 3
#endif
// The fixed-shape TSA forward (short rows: latency bound, not L1 bound, and 49 registers once its shapes are
// constants) runs five CTAs per SM: 78.9 -> 67.7 us; six (40 registers): 73.7 us.  The fixed-shape SCA forward
// does not care (235.6 us at four and at five), the backward kernels lose at two or four
// (tsa_bwd 169 -> 197 us either way): profiles/r02_fixed_shapes.md.
#ifndef FUSED_TSA_FWD_MINBLOCKS
#define FUSED_TSA_FWD_MINBLOCKS 5
#endif
#ifndef FUSED_TSA_BWD_MINBLOCKS
#define FUSED_TSA_BWD_MINBLOCKS FUSED_BWD_MINBLOCKS
#endif
#define FUSED_PRAGMA(x) _Pragma(#x)
#define FUSED_UNROLL(n) FUSED_PRAGMA(unroll n)
enum { MODE_SCA = 0, MODE_TSA = 1 };

template <int TPH>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = TPH / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int TPH>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
  for (int o = TPH / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Level table with reciprocals (offset normalisation multiplies by 1/W_l, 1/H_l).
struct FusedLevels {
  LevelTable t;
  float inv_w[kMaxLevels];
  float inv_h[kMaxLevels];
};

struct FusedArgs {
  const void* value;
  const int64_t* shapes;
  const int64_t* starts;
  const void* offsets;         // CT: fp32 or the value dtype
  const void* logits;
  const float* ref;
  const uint32_t* hit_bits;
  void* out;
  const void* g_out;
  void* g_value;               // fp32 accumulator, or fp16 (scaled by *acc_scale) when acc_half
  const float* acc_scale;
  int acc_half;
  // fp16 accumulator only: the LAST `tail_px` pixels of every value map (the coarse pyramid levels,
  // whose slots receive hundreds of updates each) have `tail_copies` extra zero-filled replicas
  // (tail_copies, batch, tail_px, M, Dh); a CTA adds into replica blockIdx % (tail_copies + 1)
  // (0 = the main buffer), which divides the number of roundings a slot's running sum sees.
  void* g_tail;
  int tail_copies, tail_px;
  void* g_offsets;             // CT
  void* g_logits;
  int bs, groups, Nk, M, Dh, L, P, D, Nq;
  float clamp;
  int bev_w, bev_h;            // query grid (0 = queries are not on a grid)
  int tile_w, tile_h;          // 2-D tile (tile_h == 0: linear runs of qpt queries)
  int tiles_x;
  int qpt;                     // queries per tile
  int rpt;                     // rows each lane group walks per tile (tile rows = ROWS * rpt)
  int tiles_per_sample;
  long long num_tiles;         // bs * tiles_per_sample
  int vec_ok;                  // rows can be written with 16-byte stores
  long long off_stride, log_stride;   // elements between the rows of consecutive queries
  int debug;                   // MSDA_DEBUG experiment switches (0 in production)
  // SCA backward, 16-bit value: the samples of the coarse pyramid levels are not scattered here but
  // written as records for the tensor-core pass (coarse_common.cuh); NULL = everything is scattered
  uint4* coarse_rec;
};

// Sample `b` and query index of local slot `lq` in tile `t`; -1 when the slot is outside the grid.
__device__ __forceinline__ int tile_query(const FusedArgs& a, long long t, int lq, int& b) {
  b = (int)(t / a.tiles_per_sample);
  const int tt = (int)(t % a.tiles_per_sample);
  if (a.tile_h > 0) {
    const int x = (tt % a.tiles_x) * a.tile_w + lq % a.tile_w;
    const int y = (tt / a.tiles_x) * a.tile_h + lq / a.tile_w;
    return (x < a.bev_w && y < a.bev_h) ? y * a.bev_w + x : -1;
  }
  const int q = tt * a.qpt + lq;
  return q < a.Nq ? q : -1;
}

__device__ __forceinline__ void load_fused_levels(FusedLevels& lv, const FusedArgs& a) {
  for (int l = threadIdx.x; l < a.L; l += blockDim.x) {
    const int h = (int)a.shapes[2 * l], w = (int)a.shapes[2 * l + 1];
    lv.t.h[l] = h;
    lv.t.w[l] = w;
    lv.t.start[l] = (int)a.starts[l];
    lv.inv_w[l] = 1.0f / (float)w;
    lv.inv_h[l] = 1.0f / (float)h;
  }
}

template <typename CT> __device__ __forceinline__ float2 load_coord2(const CT* p);
template <> __device__ __forceinline__ float2 load_coord2<float>(const float* p) {
  return __ldg(reinterpret_cast<const float2*>(p));
}
template <> __device__ __forceinline__ float2 load_coord2<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p));
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <> __device__ __forceinline__ float2 load_coord2<__half>(const __half* p) {
  const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p));
  return __half22float2(*reinterpret_cast<const __half2*>(&w));
}
template <typename CT> __device__ __forceinline__ float load_coord(const CT* p) { return to_f32<CT>(__ldg(p)); }

// 4 consecutive fp32 values from a shared row -> 4 consecutive CT values in global memory
template <typename CT> __device__ __forceinline__ void store_coord4(CT* dst, const float* src);
template <> __device__ __forceinline__ void store_coord4<float>(float* dst, const float* src) {
  *reinterpret_cast<float4*>(dst) = *reinterpret_cast<const float4*>(src);
}
template <> __device__ __forceinline__ void store_coord4<__nv_bfloat16>(__nv_bfloat16* dst, const float* src) {
  const float4 v = *reinterpret_cast<const float4*>(src);
  *reinterpret_cast<uint2*>(dst) = make_uint2(Vec16<__nv_bfloat16>::pack2(v.x, v.y), Vec16<__nv_bfloat16>::pack2(v.z, v.w));
}
template <> __device__ __forceinline__ void store_coord4<__half>(__half* dst, const float* src) {
  const float4 v = *reinterpret_cast<const float4*>(src);
  *reinterpret_cast<uint2*>(dst) = make_uint2(Vec16<__half>::pack2(v.x, v.y), Vec16<__half>::pack2(v.z, v.w));
}

// a / b given y = RN(1 / b): one Newton step on the quotient (Markstein) -- the correctly rounded
// Compile-time shapes of the two encoder configurations the step spends its time in (BEVFormer-base:
// 8 heads of 32 channels; SCA 4 levels x 8 points over 4 anchors, TSA 2 queue entries x 1 level x 4 points).
// A FIX instance copies the argument block and overwrites these fields with constants, so the sample /
// level / head loops have constant trip counts and the index arithmetic on them folds; FIX = 0 keeps
// every field a run-time value (any other shape).
#ifndef FUSED_FIXED_SHAPES
#define FUSED_FIXED_SHAPES 1
#endif
// The tiling that launch_fused derives from those shapes (rows per lane group, queries per tile, 2-D tile on
// the BEV grid) is part of the fixed set, as are "no logit clamp", 16-byte gradient rows and no debug switch.
template <int MODE, int FIX> struct FixedShape {
  static constexpr int L = 0, P = 0, D = 0, M = 0, Dh = 0, groups = 0, rpt = 0, qpt = 0, tile_w = 0, tile_h = 0;
};
template <> struct FixedShape<MODE_SCA, 1> {
  static constexpr int L = 4, P = 8, D = 4, M = 8, Dh = 32, groups = 0, rpt = 1, qpt = 8, tile_w = 4, tile_h = 2;
};
template <> struct FixedShape<MODE_TSA, 1> {
  static constexpr int L = 1, P = 4, D = 0, M = 8, Dh = 32, groups = 2, rpt = 4, qpt = 32, tile_w = 8, tile_h = 4;
};

// TILING: also the tiling / clamp / store-width constants (measured: -1 % on the forward kernels, +1-2 % on
// the backward kernels, whose register allocation is tighter -- so only the forward takes them).
template <int MODE, int FIX, bool TILING>
__device__ __forceinline__ FusedArgs fix_shape(const FusedArgs& in) {
  FusedArgs a = in;
  if constexpr (FIX != 0) {
    using F = FixedShape<MODE, FIX>;
    a.L = F::L; a.P = F::P; a.M = F::M; a.Dh = F::Dh;
    if constexpr (F::D != 0) a.D = F::D;
    if constexpr (F::groups != 0) a.groups = F::groups;
    if constexpr (TILING) {
      a.rpt = F::rpt; a.qpt = F::qpt; a.tile_w = F::tile_w; a.tile_h = F::tile_h;
      a.clamp = -1.f; a.vec_ok = 1; a.debug = 0;
    }
  }
  return a;
}
// whether the launch described by `a` (tiling already derived) is the fixed configuration
template <int MODE, int FIX>
static bool shape_is_fixed(const FusedArgs& a) {
  using F = FixedShape<MODE, FIX>;
  return FIX != 0 && a.L == F::L && a.P == F::P && a.M == F::M && a.Dh == F::Dh && (F::D == 0 || a.D == F::D) &&
         (F::groups == 0 || a.groups == F::groups) && a.rpt == F::rpt && a.qpt == F::qpt &&
         a.tile_w == F::tile_w && a.tile_h == F::tile_h && a.clamp < 0.f && a.vec_ok == 1 && a.debug == 0;
}

// fp32 quotient for the small integer divisors (map widths / heights) used here, i.e. bit-identical
// to the reference's `sampling_offsets / offset_normalizer`, at 3 FMA-pipe instructions.
__device__ __forceinline__ float div_by(float a, float b, float y) {
  const float q = a * y;
  const float r = fmaf(-q, b, a);
  return fmaf(r, y, q);
}

// Stage one row, executed by the row's TPH lanes (all lanes of the warp take part in the
// shuffles; `live` gates the memory traffic).  raw offsets -> off / (W_l, H_l) in `so`;
// raw logits -> softmax per LP-long segment in `sw`.
template <int TPH, typename CT>
__device__ __forceinline__ void stage_row(const FusedArgs& a, const FusedLevels& lv, bool live,
                                          const CT* __restrict__ g_off, const CT* __restrict__ g_log,
                                          float* so, float* sw, int chunk, int S, int LP,
                                          const uint32_t* s_meta) {
  // offsets: lane handles samples chunk, chunk + TPH, ... (8 contiguous bytes each)
  if (live) {
    for (int base = 0; base < S; base += a.P) {       // one level of one segment at a time
      const int l = (int)(s_meta[base] & 0xffffu);
      const float iw = lv.inv_w[l], ih = lv.inv_h[l];
      const float wf = (float)lv.t.w[l], hf = (float)lv.t.h[l];
      for (int s = base + chunk; s < base + a.P; s += TPH) {
        const float2 o = load_coord2<CT>(g_off + 2 * s);
        *reinterpret_cast<float2*>(so + 2 * s) = make_float2(div_by(o.x, wf, iw), div_by(o.y, hf, ih));
      }
    }
  }
  // softmax
  const int nseg = S / LP;
  const bool use_clamp = a.clamp >= 0.f;
  if (LP % TPH == 0) {
    // every segment is spread evenly over the row's lanes
    for (int sg = 0; sg < nseg; ++sg) {
      float mx = -INFINITY;
      if (live)
        for (int i = chunk; i < LP; i += TPH) {
          float v = load_coord<CT>(g_log + sg * LP + i);
          if (use_clamp) v = fminf(fmaxf(v, -a.clamp), a.clamp);
          sw[sg * LP + i] = v;
          mx = fmaxf(mx, v);
        }
      mx = group_max<TPH>(mx);
      float sum = 0.f;
      if (live)
        for (int i = chunk; i < LP; i += TPH) {
          const float e = expf(sw[sg * LP + i] - mx);
          sw[sg * LP + i] = e;
          sum += e;
        }
      sum = group_sum<TPH>(sum);
      if (live) {
        const float inv = 1.0f / sum;
        for (int i = chunk; i < LP; i += TPH) sw[sg * LP + i] *= inv;
      }
    }
  } else if (live) {
    // generic: a lane takes whole segments
    for (int sg = chunk; sg < nseg; sg += TPH) {
      float mx = -INFINITY;
      for (int i = 0; i < LP; ++i) {
        float v = load_coord<CT>(g_log + sg * LP + i);
        if (use_clamp) v = fminf(fmaxf(v, -a.clamp), a.clamp);
        sw[sg * LP + i] = v;
        mx = fmaxf(mx, v);
      }
      float sum = 0.f;
      for (int i = 0; i < LP; ++i) {
        const float e = expf(sw[sg * LP + i] - mx);
        sw[sg * LP + i] = e;
        sum += e;
      }
      const float inv = 1.0f / sum;
      for (int i = 0; i < LP; ++i) sw[sg * LP + i] *= inv;
    }
  }
  __syncwarp();
}

template <typename T, typename CT, int TPH, int MODE, int FIX>
__global__ void __launch_bounds__(kFusedThreads, (MODE == 1 && FIX != 0) ? FUSED_TSA_FWD_MINBLOCKS : FUSED_FWD_MINBLOCKS)
fused_fwd_kernel(const FusedArgs a_in) {
  const FusedArgs a = fix_shape<MODE, FIX, true>(a_in);
  constexpr int VEC = Vec16<T>::N;
  constexpr int V2 = VEC / 2;
  constexpr int ROWS = kFusedThreads / TPH;
  __shared__ FusedLevels lv;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  const int LP = a.L * a.P;
  const int S = (MODE == MODE_TSA) ? a.groups * LP : LP;
  const int po = 2 * S + 4, pw = S + 4;                // padded row pitches (floats)
  float* s_off = reinterpret_cast<float*>(dyn_smem);   // [ROWS][po]
  float* s_w = s_off + (size_t)ROWS * po;              // [ROWS][pw]
  // per-row exchange buffer: TPH sample records of 32 bytes (4 corner offsets + 4 weights), row
  // pitch padded by 16 bytes so that the rows of a warp read conflict-free
  constexpr int REC_PITCH = 8 * TPH + 4;               // in 4-byte words
  uint32_t* s_rec = reinterpret_cast<uint32_t*>(s_w + (size_t)ROWS * pw);
  // per-sample (level | anchor << 16); the anchor follows the SCA point ordering p = k*D + z
  // (quirk 5); TSA / decoder rows keep the queue entry there instead.  Keeps the integer divisions
  // out of the sample loops.
  uint32_t* s_meta = s_rec + (size_t)ROWS * REC_PITCH;
  const int tid = threadIdx.x;
  const int rows_tile = a.qpt * a.M;

  load_fused_levels(lv, a);
  for (int i = tid; i < S; i += kFusedThreads) {
    const int sl = i % LP, l = sl / a.P;
    const int hi = (MODE == MODE_SCA) ? (sl - l * a.P) % a.D : i / LP;   // Z-anchor | queue entry
    s_meta[i] = (uint32_t)l | ((uint32_t)hi << 16);
  }
  __syncthreads();

  const int chunk = tid % TPH;
  const int r_slot = tid / TPH;
  float* my_off = s_off + (size_t)r_slot * po;
  float* my_w = s_w + (size_t)r_slot * pw;
  uint32_t* my_rec = s_rec + (size_t)r_slot * REC_PITCH;
  const int pix_stride = a.M * a.Dh;
  const size_t batch_stride = (size_t)a.Nk * pix_stride;

  for (long long t = blockIdx.x; t < a.num_tiles; t += gridDim.x) {
    for (int rk = 0; rk < a.rpt; ++rk) {
      const int r_local = rk * ROWS + r_slot;
      const bool in_tile = r_local < rows_tile;
      const int lq = in_tile ? r_local / a.M : 0;
      const int m = in_tile ? r_local % a.M : 0;
      int b;
      int q = tile_query(a, t, lq, b);
      const bool live = in_tile && q >= 0;
      if (!live) { q = 0; b = 0; }
      uint32_t hits = 0;
      float scale = (float)a.groups;
      if (MODE == MODE_SCA) {
        hits = live ? a.hit_bits[q] : 0u;                             // batch element 0 decides (quirk 1)
        const int cnt = __popc(a.hit_bits[(size_t)b * a.Nq + q]);     // the divisor is per sample
        scale = (float)(cnt > 0 ? cnt : 1);
      }
      const bool work = live && (MODE == MODE_TSA || hits != 0);
      const long long grow = ((long long)b * a.Nq + q) * a.M + m;
      __syncwarp();                                                   // previous row's readers are done
      const long long qrow = (long long)b * a.Nq + q;
      const long long off_at = qrow * a.off_stride + (long long)m * S * 2;
      const long long log_at = qrow * a.log_stride + (long long)m * S;
      stage_row<TPH, CT>(a, lv, work, static_cast<const CT*>(a.offsets) + off_at,
                         static_cast<const CT*>(a.logits) + log_at, my_off, my_w, chunk, S, LP, s_meta);

      {
        const T* vhead = static_cast<const T*>(a.value) + (size_t)m * a.Dh + chunk * VEC;
        constexpr bool MIXED = sizeof(T) == 2;   // 16-bit value: FHFMA straight from the loaded words
        float2 acc[V2];
        float accm[8];
#pragma unroll
        for (int i = 0; i < V2; ++i) acc[i] = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) accm[i] = 0.f;

        // a * bilinear weights against the four gathered corners.  fp32 value: FFMA2 with fp32
        // weights.  16-bit value: the weights are rounded to the value dtype (wa = w00|w01,
        // wb = w10|w11 packed), products are exact, accumulation is fp32.
        auto blend = [&](const uint4& u00, const uint4& u01, const uint4& u10, const uint4& u11,
                         float w00f, float w01f, float w10f, float w11f, uint32_t wa, uint32_t wb) {
          if constexpr (MIXED) {
            axpy_mixed<T>(u00, lo16(wa), accm);
            axpy_mixed<T>(u01, hi16(wa), accm);
            axpy_mixed<T>(u10, lo16(wb), accm);
            axpy_mixed<T>(u11, hi16(wb), accm);
          } else {
            const float2 w00 = splat2(w00f), w01 = splat2(w01f), w10 = splat2(w10f), w11 = splat2(w11f);
            float2 f[V2];
            Vec16<T>::unpack2(u00, f);
#pragma unroll
            for (int i = 0; i < V2; ++i) acc[i] = ffma2(w00, f[i], acc[i]);
            Vec16<T>::unpack2(u01, f);
#pragma unroll
            for (int i = 0; i < V2; ++i) acc[i] = ffma2(w01, f[i], acc[i]);
            Vec16<T>::unpack2(u10, f);
#pragma unroll
            for (int i = 0; i < V2; ++i) acc[i] = ffma2(w10, f[i], acc[i]);
            Vec16<T>::unpack2(u11, f);
#pragma unroll
            for (int i = 0; i < V2; ++i) acc[i] = ffma2(w11, f[i], acc[i]);
          }
        };

        // Each of the row's TPH lanes sets up ONE sample of a block of TPH (corner BYTE offsets
        // from the camera's value slab, weights = attention x bilinear, zero for corners outside
        // the map), the lanes exchange the records through the row's shared buffer, and then
        // every lane walks the block: 2 shared loads + 4 unpredicated 128-bit gathers per sample.
        auto accumulate_block = [&](const T* vbase_cam, int count, bool mine) {
          const char* vb = reinterpret_cast<const char*>(vbase_cam);
          __syncwarp();               // the camera loop is warp-uniform (rows not hit idle through it)
#pragma unroll
          for (int j = 0; j < TPH; ++j) {
            if (j < count && mine) {
              const uint4 o = *reinterpret_cast<const uint4*>(my_rec + 8 * j);
              if (!(o.x & 1u)) continue;          // sample entirely outside the map: nothing to read
              const uint4 u00 = ldg128(vb + (o.x & ~15u));
              const uint4 u01 = ldg128(vb + o.y);
              const uint4 u10 = ldg128(vb + o.z);
              const uint4 u11 = ldg128(vb + o.w);
              if constexpr (MIXED) {
                const uint2 wq = *reinterpret_cast<const uint2*>(my_rec + 8 * j + 4);
                blend(u00, u01, u10, u11, 0.f, 0.f, 0.f, 0.f, wq.x, wq.y);
              } else {
                const float4 wq = *reinterpret_cast<const float4*>(my_rec + 8 * j + 4);
                blend(u00, u01, u10, u11, wq.x, wq.y, wq.z, wq.w, 0u, 0u);
              }
            }
          }
          __syncwarp();
        };
        // `map_off`: element offset of the sample's value map from the base the block is walked with
        auto write_record = [&](int l, float lx, float ly, float w, unsigned map_off) {
          const Corners c = corner_setup(lx, ly, lv.t.h[l], lv.t.w[l], pix_stride);
          const unsigned base = map_off + (unsigned)(lv.t.start[l] * pix_stride);
          constexpr unsigned ES = sizeof(T);
          // (bit 0 of the first offset -- offsets are multiples of 16 -- flags a sample with at least
          // one corner inside the map; the others are skipped, so a non-finite value only reaches
          // samples that touch it)
          *reinterpret_cast<uint4*>(my_rec + 8 * chunk) =
              make_uint4(((base + (unsigned)c.o00) * ES) | (c.valid ? 1u : 0u), (base + (unsigned)c.o01) * ES,
                         (base + (unsigned)c.o10) * ES, (base + (unsigned)c.o11) * ES);
          if constexpr (MIXED) {
            *reinterpret_cast<uint2*>(my_rec + 8 * chunk + 4) =
                make_uint2(Vec16<T>::pack2(w * c.w00, w * c.w01), Vec16<T>::pack2(w * c.w10, w * c.w11));
          } else {
            *reinterpret_cast<float4*>(my_rec + 8 * chunk + 4) =
                make_float4(w * c.w00, w * c.w01, w * c.w10, w * c.w11);
          }
        };

        if (MODE == MODE_SCA) {
          // every lane of the warp walks the union of the warp's cameras (with 8 heads per warp
          // that is the query's own list); rows the camera does not hit skip records and gathers
          uint32_t warp_hits = __reduce_or_sync(0xffffffffu, hits);
          while (warp_hits) {
            const int cam = __ffs(warp_hits) - 1;
            warp_hits &= warp_hits - 1;
            const bool mine = (hits >> cam) & 1u;
            const float2* rc = reinterpret_cast<const float2*>(
                a.ref + (((size_t)cam * a.bs + b) * a.Nq + q) * a.D * 2);
            const T* vcam = vhead + ((size_t)b * a.groups + cam) * batch_stride;
            for (int s0 = 0; s0 < LP; s0 += TPH) {
              const int s = s0 + chunk;
              if (s < LP && mine) {
                const uint32_t meta = s_meta[s];
                const float2 r = __ldg(rc + (meta >> 16));
                const float2 o = *reinterpret_cast<const float2*>(my_off + 2 * s);
                write_record((int)(meta & 0xffffu), r.x + o.x, r.y + o.y, my_w[s], 0u);
              }
              accumulate_block(vcam, min(TPH, LP - s0), mine);
            }
          }
        } else {
          // TSA / decoder rows (4-8 samples per queue entry): the same exchange -- lane c sets up
          // sample s0 + c of the row (queue entry and level from the shared table), every lane
          // walks the block -- so a sample is set up once per row, not once per lane
          const T* vb = vhead + (size_t)b * a.groups * batch_stride;
          for (int s0 = 0; s0 < S; s0 += TPH) {
            const int s = s0 + chunk;
            if (s < S && live) {
              const uint32_t meta = s_meta[s];
              const int l = (int)(meta & 0xffffu), j = (int)(meta >> 16);
              const float2 r = __ldg(reinterpret_cast<const float2*>(
                  a.ref + ((((size_t)b * a.groups + j) * a.Nq + q) * a.L + l) * 2));
              const float2 o = *reinterpret_cast<const float2*>(my_off + 2 * s);
              write_record(l, r.x + o.x, r.y + o.y, my_w[s], (unsigned)j * (unsigned)batch_stride);
            }
            accumulate_block(vb, min(TPH, S - s0), live);
          }
        }
        float out[VEC];
#pragma unroll
        for (int i = 0; i < V2; ++i) {
          out[2 * i] = MIXED ? accm[(2 * i) % 8] : acc[i].x;
          out[2 * i + 1] = MIXED ? accm[(2 * i + 1) % 8] : acc[i].y;
        }
        if (scale == 2.f) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) out[i] *= 0.5f;
        } else if (scale != 1.f) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) out[i] = out[i] / scale;
        }
        if (live) {
          T* o = static_cast<T*>(a.out) + grow * a.Dh + chunk * VEC;
          *reinterpret_cast<uint4*>(o) = Vec16<T>::pack(out);
        }
      }
    }
  }
}

template <typename T, typename CT, int TPH, int MODE, bool ACC_HALF, int FIX>
__global__ void __launch_bounds__(kFusedThreads, (MODE == 1 && FIX != 0) ? FUSED_TSA_BWD_MINBLOCKS : FUSED_BWD_MINBLOCKS)
fused_bwd_kernel(const FusedArgs a_in) {
  const FusedArgs a = fix_shape<MODE, FIX, false>(a_in);
  constexpr int VEC = Vec16<T>::N;
  constexpr int V2 = VEC / 2;
  constexpr int ROWS = kFusedThreads / TPH;
  constexpr int SC = VEC / 4;                          // scatter instructions per corner
  constexpr unsigned ES_ACC = sizeof(T);               // byte offsets of the records are in units of T
  __shared__ FusedLevels lv;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  const int LP = a.L * a.P;
  const int S = (MODE == MODE_TSA) ? a.groups * LP : LP;
  const int po = 2 * S + 4, pw = S + 4;
  float* s_off = reinterpret_cast<float*>(dyn_smem);   // [ROWS][po] normalised offsets
  float* s_w = s_off + (size_t)ROWS * po;              // [ROWS][pw] softmax weights
  float* s_go = s_w + (size_t)ROWS * pw;               // [ROWS][po] grad wrt raw offsets
  float* s_ga = s_go + (size_t)ROWS * po;              // [ROWS][pw] grad wrt weights -> logits
  // per-row exchange buffer of TPH sample records, in 4-byte words: 4 corner byte offsets (the
  // validity bits ride in the low 4 bits of the first, the offsets are multiples of 16), lw, lh
  // and the scatter weights attention x bilinear -- two packed fp16 pairs for the fp16 accumulator,
  // four floats otherwise
  constexpr int REC_WORDS = ACC_HALF ? 8 : 12;
  constexpr int REC_PITCH = REC_WORDS * TPH + 4;
  uint32_t* s_rec = reinterpret_cast<uint32_t*>(s_ga + (size_t)ROWS * pw);
  uint32_t* s_meta = s_rec + (size_t)ROWS * REC_PITCH;   // per-sample level | anchor << 16
  // coarse patch of the level table (first level, first pixel); first level == L: none
  constexpr bool COARSE = MODE == MODE_SCA && sizeof(T) == 2;
  __shared__ int s_coarse[2];
  const int tid = threadIdx.x;
  const int rows_tile = a.qpt * a.M;

  load_fused_levels(lv, a);
  for (int i = tid; i < S; i += kFusedThreads) {
    const int sl = i % LP, l = sl / a.P;
    const int hi = (MODE == MODE_SCA) ? (sl - l * a.P) % a.D : i / LP;   // Z-anchor | queue entry
    s_meta[i] = (uint32_t)l | ((uint32_t)hi << 16);
  }
  if (tid == 0) {
    s_coarse[0] = a.L;
    s_coarse[1] = 0;
    if (COARSE && a.coarse_rec != nullptr) {
      const CoarsePatch cp = coarse_patch(a.shapes, a.starts, a.L);
      s_coarse[0] = cp.first_level;
      s_coarse[1] = cp.start;
    }
  }
  __syncthreads();
  const int coarse_first = COARSE ? s_coarse[0] : a.L;
  const int coarse_start = COARSE ? s_coarse[1] : 0;

  const int chunk = tid % TPH;
  const int r_slot = tid / TPH;
  float* my_off = s_off + (size_t)r_slot * po;
  float* my_w = s_w + (size_t)r_slot * pw;
  float* my_go = s_go + (size_t)r_slot * po;
  float* my_ga = s_ga + (size_t)r_slot * pw;
  uint32_t* my_rec = s_rec + (size_t)r_slot * REC_PITCH;
  const int pix_stride = a.M * a.Dh;
  const size_t batch_stride = (size_t)a.Nk * pix_stride;

  for (long long t = blockIdx.x; t < a.num_tiles; t += gridDim.x) {
    for (int rk = 0; rk < a.rpt; ++rk) {
      // Every lane of the warp walks the same loops (the lane sums use full-warp shuffles);
      // lanes without work carry weight zero and touch nothing.
      const int r_local = rk * ROWS + r_slot;
      const bool in_tile = r_local < rows_tile;
      const int lq = in_tile ? r_local / a.M : 0;
      const int m = in_tile ? r_local % a.M : 0;
      int b;
      int q = tile_query(a, t, lq, b);
      const bool live = in_tile && q >= 0;
      if (!live) { q = 0; b = 0; }
      uint32_t hits = 0;
      float scale = (float)a.groups;
      if (MODE == MODE_SCA) {
        hits = live ? a.hit_bits[q] : 0u;
        const int cnt = __popc(a.hit_bits[(size_t)b * a.Nq + q]);
        scale = (float)(cnt > 0 ? cnt : 1);
      }
      const bool work = live && (MODE == MODE_TSA || hits != 0);
      const long long grow = ((long long)b * a.Nq + q) * a.M + m;
      __syncwarp();
      const long long qrow = (long long)b * a.Nq + q;
      const long long off_at = qrow * a.off_stride + (long long)m * S * 2;
      const long long log_at = qrow * a.log_stride + (long long)m * S;
      stage_row<TPH, CT>(a, lv, work, static_cast<const CT*>(a.offsets) + off_at,
                         static_cast<const CT*>(a.logits) + log_at, my_off, my_w, chunk, S, LP, s_meta);
      // clear the row's gradient accumulators
      for (int i = chunk; i < 2 * S; i += TPH) my_go[i] = 0.f;
      for (int i = chunk; i < S; i += TPH) my_ga[i] = 0.f;
      __syncwarp();

      const size_t head_off = (size_t)m * a.Dh;
      const T* vhead = static_cast<const T*>(a.value) + head_off + chunk * VEC;
      // Channel ownership for the grad_value scatter.  With 8 channels per lane (16-bit value)
      // lane c scatters channels [4c, 4c+4) and [4*TPH + 4c, 4*TPH + 4c + 4), so that the TPH lanes
      // of a head cover 16*TPH contiguous bytes per reduction instruction (whole sectors); the dot
      // products use the natural [8c, 8c+8) ownership of the 128-bit value loads.
      float* ghead = static_cast<float*>(a.g_value) + head_off + 4 * chunk;
      // fp16 accumulator (16-bit value dtypes): one 16-byte red.v4.f16x2 carries the lane's 8
      // channels in their natural order, scaled by a power of two chosen from max|g_out|
      __half* ghead16 = static_cast<__half*>(a.g_value) + head_off + chunk * VEC;
      const float acc_scale = ACC_HALF ? __ldg(a.acc_scale) : 1.f;
      // replica of the accumulator's coarse tail this CTA adds into (see FusedArgs::g_tail).  Not
      // for TPH == 4 (head_dim 32 with a 16-bit value): those shapes have the tensor-core pass for
      // the coarse levels, and the address selects cost 24 instructions per sample in the hot loop.
      constexpr bool TAIL = ACC_HALF && MODE == MODE_SCA && TPH != 4;
      const int replica = (TAIL && a.tail_copies > 0) ? (int)(blockIdx.x % (unsigned)(a.tail_copies + 1)) : 0;
      const unsigned tail_from = replica > 0 ? (unsigned)(a.Nk - a.tail_px) * (unsigned)pix_stride * ES_ACC
                                             : 0xffffffffu;      // byte offset inside a value map
      const size_t tail_map = (size_t)a.tail_px * pix_stride;    // elements per map in a replica
      __half* tail16 = static_cast<__half*>(a.g_tail) + head_off + chunk * VEC;
      char* acc_tail = nullptr;     // set per value map: replica base minus tail_from (SCA only)

      // 16-bit value: the dot products take g_out straight from its packed words (FHFMA, exact
      // products) and the 1 / count factor is applied to the per-sample totals instead
      constexpr bool MIXED = sizeof(T) == 2;
      constexpr unsigned ES = sizeof(T);
      const float dscale = !MIXED ? 1.f : (scale == 2.f ? 0.5f : 1.f / scale);
      float2 g[V2], gs[V2];
      uint4 gp;
      {
        const T* grow_ptr = static_cast<const T*>(a.g_out) + grow * a.Dh;
        gp = ldg128(grow_ptr + chunk * VEC);
        Vec16<T>::unpack2(gp, g);
        if (VEC == 4 || ACC_HALF) {
#pragma unroll
          for (int i = 0; i < V2; ++i) gs[i] = g[i];
        } else {
          const T* lo = grow_ptr + 4 * chunk;
          const T* hi = grow_ptr + 4 * TPH + 4 * chunk;
          gs[0] = make_float2(to_f32<T>(lo[0]), to_f32<T>(lo[1]));
          gs[1] = make_float2(to_f32<T>(lo[2]), to_f32<T>(lo[3]));
          gs[V2 - 2] = make_float2(to_f32<T>(hi[0]), to_f32<T>(hi[1]));
          gs[V2 - 1] = make_float2(to_f32<T>(hi[2]), to_f32<T>(hi[3]));
        }
        if (scale != 1.f) {
          const bool half = scale == 2.f;
#pragma unroll
          for (int i = 0; i < V2; ++i) {
            g[i] = half ? make_float2(g[i].x * 0.5f, g[i].y * 0.5f)
                        : make_float2(g[i].x / scale, g[i].y / scale);
            gs[i] = half ? make_float2(gs[i].x * 0.5f, gs[i].y * 0.5f)
                         : make_float2(gs[i].x / scale, gs[i].y / scale);
          }
        }
      }

      __half2 gh[4];
      if (ACC_HALF) {
#pragma unroll
        for (int k = 0; k < 4; ++k) gh[k] = __floats2half2_rn(g[k % V2].x * acc_scale, g[k % V2].y * acc_scale);
      }

      // Core of one sample against one value map, from its record: gathers the four corners,
      // scatters grad_value, returns this lane's partial (d out / d weight, d/dx, d/dy).  o00v =
      // byte offset of corner 00 | validity bits; aw* = attention x bilinear weight per corner
      // (zero for invalid corners and for rows without work), as packed fp16 pairs (awa = 00|01,
      // awb = 10|11) for the fp16 accumulator or as floats.
      auto core = [&](auto scatter_tag, size_t boff, unsigned o00v, unsigned o01, unsigned o10, unsigned o11, float lw,
                      float lh, uint32_t awa, uint32_t awb, float aw00, float aw01, float aw10,
                      float aw11, float& ga, float& gx, float& gy) {
        constexpr bool SCATTER = decltype(scatter_tag)::value;
        const unsigned o00 = o00v & ~15u;
        const char* vb = reinterpret_cast<const char*>(vhead + boff);
        const uint4 u00 = ldg128(vb + o00);
        const uint4 u01 = ldg128(vb + o01);
        const uint4 u10 = ldg128(vb + o10);
        const uint4 u11 = ldg128(vb + o11);
        if constexpr (!SCATTER) {
          // (a block of coarse-level samples: their grad_value goes through the records)
        } else if constexpr (ACC_HALF) {
          // gh = fp16(g * scale) per row; one HMUL2 per channel pair, one predicated 16-byte
          // reduction per corner (skipped where the weight is zero); corners in the coarse tail
          // of the map go to this CTA's replica
          char* gb = reinterpret_cast<char*>(ghead16 + boff);
          char* gt = acc_tail != nullptr ? acc_tail : gb;
          auto scatter = [&](unsigned off, __half2 aw2, uint32_t on) {
            uint32_t h[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const __half2 hk = __hmul2(aw2, gh[k]);
              h[k] = *reinterpret_cast<const uint32_t*>(&hk);
            }
            char* dst = (TAIL ? (off >= tail_from ? gt : gb) : gb) + off;
            red_add_f16x8_if(reinterpret_cast<__half*>(dst), h[0], h[1], h[2], h[3], on);
          };
          const __half2 pa = *reinterpret_cast<const __half2*>(&awa);
          const __half2 pb = *reinterpret_cast<const __half2*>(&awb);
          const uint32_t live_mask = (a.debug & 1) ? 0u : 0x7fffu;
          scatter(o00, __low2half2(pa), awa & live_mask);
          scatter(o01, __high2half2(pa), (awa >> 16) & live_mask);
          scatter(o10, __low2half2(pb), awb & live_mask);
          scatter(o11, __high2half2(pb), (awb >> 16) & live_mask);
        } else {
          char* gb = reinterpret_cast<char*>(ghead + boff);
          auto scatter = [&](unsigned off, float aw) {
            const float2 aw2 = splat2(aw);
            float* dst = reinterpret_cast<float*>(gb + (size_t)off * (sizeof(float) / ES));
            const uint32_t on = (aw != 0.f && !(a.debug & 1)) ? 1u : 0u;
#pragma unroll
            for (int k = 0; k < SC; ++k) {
              const float2 p0 = fmul2(aw2, gs[2 * k]), p1 = fmul2(aw2, gs[2 * k + 1]);
              red_add_f32x4_if(dst + k * 4 * TPH, p0.x, p0.y, p1.x, p1.y, on);
            }
          };
          scatter(o00, aw00);
          scatter(o01, aw01);
          scatter(o10, aw10);
          scatter(o11, aw11);
        }
        float d00, d01, d10, d11;
        if constexpr (MIXED) {
          d00 = dot_mixed<T>(u00, gp);
          d01 = dot_mixed<T>(u01, gp);
          d10 = dot_mixed<T>(u10, gp);
          d11 = dot_mixed<T>(u11, gp);
        } else {
          float2 f[V2];
          float2 d;
          Vec16<T>::unpack2(u00, f);
          d = make_float2(0.f, 0.f);
#pragma unroll
          for (int i = 0; i < V2; ++i) d = ffma2(f[i], g[i], d);
          d00 = d.x + d.y;
          Vec16<T>::unpack2(u01, f);
          d = make_float2(0.f, 0.f);
#pragma unroll
          for (int i = 0; i < V2; ++i) d = ffma2(f[i], g[i], d);
          d01 = d.x + d.y;
          Vec16<T>::unpack2(u10, f);
          d = make_float2(0.f, 0.f);
#pragma unroll
          for (int i = 0; i < V2; ++i) d = ffma2(f[i], g[i], d);
          d10 = d.x + d.y;
          Vec16<T>::unpack2(u11, f);
          d = make_float2(0.f, 0.f);
#pragma unroll
          for (int i = 0; i < V2; ++i) d = ffma2(f[i], g[i], d);
          d11 = d.x + d.y;
        }
        // corners outside the map were gathered from a clamped (in-bounds) address: drop them
        d00 = (o00v & 1u) ? d00 : 0.f;
        d01 = (o00v & 2u) ? d01 : 0.f;
        d10 = (o00v & 4u) ? d10 : 0.f;
        d11 = (o00v & 8u) ? d11 : 0.f;
        const float hw = 1.f - lw, hh = 1.f - lh;
        ga = hh * (hw * d00 + lw * d01) + lh * (hw * d10 + lw * d11);
        gx = hh * (d01 - d00) + lh * (d11 - d10);
        gy = hw * (d10 - d00) + lw * (d11 - d01);
      };
      // record fields after the offsets, from a corner set-up and the sample's attention weight
      auto scatter_weights = [&](const Corners& c, float w, uint32_t (&rw)[REC_WORDS - 4]) {
        rw[0] = __float_as_uint(c.lw);
        rw[1] = __float_as_uint(c.lh);
        if constexpr (ACC_HALF) {
          const __half2 pa = __floats2half2_rn(w * c.w00, w * c.w01);
          const __half2 pb = __floats2half2_rn(w * c.w10, w * c.w11);
          rw[2] = *reinterpret_cast<const uint32_t*>(&pa);
          rw[3] = *reinterpret_cast<const uint32_t*>(&pb);
        } else {
          rw[2] = 0u; rw[3] = 0u;
          rw[4] = __float_as_uint(w * c.w00); rw[5] = __float_as_uint(w * c.w01);
          rw[6] = __float_as_uint(w * c.w10); rw[7] = __float_as_uint(w * c.w11);
        }
      };

      // Blocks of TPH samples: lane c sets up sample s0 + c and publishes its record (corner
      // offsets, fractions, validity, attention weight) in the row's exchange buffer; every lane
      // then walks the block; after the lane sums, lane c owns the totals of sample s0 + c and
      // adds them to the row's accumulators (d loc / d offset = 1 / (W_l, H_l) cancels the
      // (W_l, H_l) factor of d pixel / d loc).  `have`: this lane's sample exists; `keep`: its row takes
      // part (it accumulates the totals); (l, lx, ly, wgt) level, location and attention weight (zero
      // for rows that do not take part: their gathers stay unpredicated, nothing is scattered);
      // `map_off`: element offset of its value map from `coff`, the base the block is walked with.
      // `crow`: the row's records for the tensor-core pass of the coarse levels (indexed by the
      // sample), or NULL
      auto run_block = [&](auto scatter_tag, size_t coff, int s, int count, bool have, bool keep, int l, float lx,
                           float ly, float wgt, unsigned map_off, uint4* crow) {
        float my_wgt = 0.f;
        {
          uint4 ro = make_uint4(0u, 0u, 0u, 0u);
          uint32_t rw[REC_WORDS - 4];
#pragma unroll
          for (int i = 0; i < REC_WORDS - 4; ++i) rw[i] = 0u;
          if (have) {
            int px[4];
            const Corners c = corner_setup_px(lx, ly, lv.t.h[l], lv.t.w[l], pix_stride, px);
            const unsigned base = map_off + (unsigned)(lv.t.start[l] * pix_stride);
            my_wgt = wgt;
            ro = make_uint4(((base + (unsigned)c.o00) * ES) | c.valid, (base + (unsigned)c.o01) * ES,
                            (base + (unsigned)c.o10) * ES, (base + (unsigned)c.o11) * ES);
            scatter_weights(c, my_wgt, rw);
            if constexpr (COARSE) {
              if (l >= coarse_first) {
                // grad_value of this sample is accumulated by the tensor-core pass: no reductions
                // here (zero scatter weights), one 16-byte record instead -- per corner the pixel
                // inside the coarse patch and attention x bilinear / count as a 16-bit float
                if (keep && crow != nullptr) {
                  const int p0 = lv.t.start[l] - coarse_start;
                  const float ws = my_wgt * dscale;
                  using WT = typename std::conditional<ACC_HALF, __half, T>::type;
                  crow[s] = make_uint4(coarse_corner<WT>(p0 + px[0], ws * c.w00), coarse_corner<WT>(p0 + px[1], ws * c.w01),
                                       coarse_corner<WT>(p0 + px[2], ws * c.w10), coarse_corner<WT>(p0 + px[3], ws * c.w11));
                }
#pragma unroll
                for (int i = 2; i < REC_WORDS - 4; ++i) rw[i] = 0u;
              }
            }
          }
          uint32_t* rec = my_rec + REC_WORDS * chunk;
          *reinterpret_cast<uint4*>(rec) = ro;
          *reinterpret_cast<uint4*>(rec + 4) = make_uint4(rw[0], rw[1], rw[2], rw[3]);
          if constexpr (!ACC_HALF)
            *reinterpret_cast<uint4*>(rec + 8) = make_uint4(rw[4], rw[5], rw[6], rw[7]);
        }
        __syncwarp();
        float tga = 0.f, tgx = 0.f, tgy = 0.f;
        if constexpr (TPH <= 8) {
          // Fully unrolled walk (record addresses become immediates) with the lane sums done once per
          // block: every lane keeps its partial (d weight, d x, d y) of each of the block's samples,
          // and a halving exchange -- TPH - 1 shuffles per quantity instead of TPH log2 TPH -- leaves
          // lane c with the totals of sample s0 + c.
          float pa[TPH], px_[TPH], py[TPH];
#pragma unroll
          for (int j = 0; j < TPH; ++j) {
            pa[j] = 0.f; px_[j] = 0.f; py[j] = 0.f;
            if (j < count) {
              const uint32_t* rec = my_rec + REC_WORDS * j;
              const uint4 ro = *reinterpret_cast<const uint4*>(rec);
              const uint4 rf = *reinterpret_cast<const uint4*>(rec + 4);
              float4 ra = make_float4(0.f, 0.f, 0.f, 0.f);
              if constexpr (!ACC_HALF) ra = *reinterpret_cast<const float4*>(rec + 8);
              core(scatter_tag, coff, ro.x, ro.y, ro.z, ro.w, __uint_as_float(rf.x), __uint_as_float(rf.y), rf.z, rf.w,
                   ra.x, ra.y, ra.z, ra.w, pa[j], px_[j], py[j]);
            }
          }
#pragma unroll
          for (int w = TPH / 2; w >= 1; w >>= 1) {
            const bool up = (chunk & w) != 0;          // keeps the upper half of the samples, sends the lower
#pragma unroll
            for (int i = 0; i < w; ++i) {
              const float sa = up ? pa[i] : pa[i + w], ka = up ? pa[i + w] : pa[i];
              const float sx = up ? px_[i] : px_[i + w], kx = up ? px_[i + w] : px_[i];
              const float sy = up ? py[i] : py[i + w], ky = up ? py[i + w] : py[i];
              pa[i] = ka + __shfl_xor_sync(0xffffffffu, sa, w);
              px_[i] = kx + __shfl_xor_sync(0xffffffffu, sx, w);
              py[i] = ky + __shfl_xor_sync(0xffffffffu, sy, w);
            }
          }
          tga = pa[0]; tgx = px_[0]; tgy = py[0];
        } else {
FUSED_UNROLL(FUSED_BWD_UNROLL)
          for (int j = 0; j < TPH; ++j) {
            if (j < count) {
              const uint32_t* rec = my_rec + REC_WORDS * j;
              const uint4 ro = *reinterpret_cast<const uint4*>(rec);
              const uint4 rf = *reinterpret_cast<const uint4*>(rec + 4);
              float4 ra = make_float4(0.f, 0.f, 0.f, 0.f);
              if constexpr (!ACC_HALF) ra = *reinterpret_cast<const float4*>(rec + 8);
              float ga, gx, gy;
              core(scatter_tag, coff, ro.x, ro.y, ro.z, ro.w, __uint_as_float(rf.x), __uint_as_float(rf.y), rf.z, rf.w,
                   ra.x, ra.y, ra.z, ra.w, ga, gx, gy);
              ga = group_sum<TPH>(ga);
              gx = group_sum<TPH>(gx);
              gy = group_sum<TPH>(gy);
              if (j == chunk) { tga = ga; tgx = gx; tgy = gy; }
            }
          }
        }
        if (keep) {
          my_ga[s] += tga * dscale;
          float2* go2 = reinterpret_cast<float2*>(my_go + 2 * s);
          float2 cur = *go2;
          const float ws = my_wgt * dscale;
          cur.x += ws * tgx;
          cur.y += ws * tgy;
          *go2 = cur;
        }
        __syncwarp();
      };

      if (MODE == MODE_SCA) {
        uint32_t warp_hits = __reduce_or_sync(0xffffffffu, hits);
        while (warp_hits) {
          const int cam = __ffs(warp_hits) - 1;
          warp_hits &= warp_hits - 1;
          const bool mine = (hits >> cam) & 1u;
          const float2* rc = reinterpret_cast<const float2*>(
              a.ref + (((size_t)cam * a.bs + b) * a.Nq + q) * a.D * 2);
          const size_t coff = ((size_t)b * a.groups + cam) * batch_stride;
          uint4* crow = nullptr;
          if (COARSE && coarse_first < a.L)
            crow = a.coarse_rec + ((((size_t)b * a.groups + cam) * a.Nq + q) * a.M + m) * (kCoarseMaxLevels * a.P) -
                   (size_t)coarse_first * a.P;
          if (replica > 0)
            acc_tail = reinterpret_cast<char*>(
                           tail16 + ((size_t)(replica - 1) * a.bs * a.groups + (size_t)b * a.groups + cam) * tail_map) -
                       tail_from;
          for (int s0 = 0; s0 < LP; s0 += TPH) {
            const int s = s0 + chunk;
            int l = 0;
            float lx = 0.f, ly = 0.f, wgt = 0.f;
            const bool in_row = s < LP;
            if (in_row) {
              const uint32_t meta = s_meta[s];
              l = (int)(meta & 0xffffu);
              const float2 r = __ldg(rc + (meta >> 16));
              const float2 o = *reinterpret_cast<const float2*>(my_off + 2 * s);
              lx = r.x + o.x;
              ly = r.y + o.y;
              wgt = mine ? my_w[s] : 0.f;
            }
            // (the coarse levels are a suffix of the sample axis: a block that starts inside them has
            // nothing to scatter -- a warp-uniform choice)
            if (COARSE && s0 >= coarse_first * a.P)
              run_block(std::false_type{}, coff, s, min(TPH, LP - s0), in_row, in_row && mine, l, lx, ly, wgt, 0u, crow);
            else
              run_block(std::true_type{}, coff, s, min(TPH, LP - s0), in_row, in_row && mine, l, lx, ly, wgt, 0u, crow);
          }
        }
      } else {
        // TSA / decoder rows: the same exchange; the queue entry and the level of a sample come from
        // the shared table
        const size_t coff = (size_t)b * a.groups * batch_stride;
        for (int s0 = 0; s0 < S; s0 += TPH) {
          const int s = s0 + chunk;
          int l = 0;
          float lx = 0.f, ly = 0.f, wgt = 0.f;
          unsigned map_off = 0u;
          const bool in_row = s < S;
          if (in_row) {
            const uint32_t meta = s_meta[s];
            l = (int)(meta & 0xffffu);
            const int j = (int)(meta >> 16);
            const float2 r = __ldg(reinterpret_cast<const float2*>(
                a.ref + ((((size_t)b * a.groups + j) * a.Nq + q) * a.L + l) * 2));
            const float2 o = *reinterpret_cast<const float2*>(my_off + 2 * s);
            lx = r.x + o.x;
            ly = r.y + o.y;
            wgt = live ? my_w[s] : 0.f;
            map_off = (unsigned)j * (unsigned)batch_stride;
          }
          run_block(std::true_type{}, coff, s, min(TPH, S - s0), in_row, in_row && live, l, lx, ly, wgt, map_off, nullptr);
        }
      }
      __syncwarp();

      // softmax backward per segment: g_logit = a * (ga - sum_t a_t ga_t); zero where the raw
      // logit was clamped.  Result overwrites the ga row.
      {
        const int nseg = S / LP;
        const bool use_clamp = a.clamp >= 0.f;
        const CT* raw = static_cast<const CT*>(a.logits) + log_at;
        if (LP % TPH == 0) {
          for (int sg = 0; sg < nseg; ++sg) {
            float dot = 0.f;
            for (int i = chunk; i < LP; i += TPH) dot = fmaf(my_w[sg * LP + i], my_ga[sg * LP + i], dot);
            dot = group_sum<TPH>(dot);
            for (int i = chunk; i < LP; i += TPH) {
              const int e = sg * LP + i;
              float gl = work ? my_w[e] * (my_ga[e] - dot) : 0.f;   // rows without work hold stale weights
              if (use_clamp && live) {
                const float v = load_coord<CT>(raw + e);
                if (v < -a.clamp || v > a.clamp) gl = 0.f;
              }
              my_ga[e] = gl;
            }
          }
        } else {
          for (int sg = chunk; sg < nseg; sg += TPH) {
            float dot = 0.f;
            for (int i = 0; i < LP; ++i) dot = fmaf(my_w[sg * LP + i], my_ga[sg * LP + i], dot);
            for (int i = 0; i < LP; ++i) {
              const int e = sg * LP + i;
              float gl = work ? my_w[e] * (my_ga[e] - dot) : 0.f;
              if (use_clamp && live) {
                const float v = load_coord<CT>(raw + e);
                if (v < -a.clamp || v > a.clamp) gl = 0.f;
              }
              my_ga[e] = gl;
            }
          }
        }
      }
      __syncwarp();

      // gradient rows back to global memory
      if (live) {
        CT* dst_o = static_cast<CT*>(a.g_offsets) + off_at;
        CT* dst_l = static_cast<CT*>(a.g_logits) + log_at;
        if (a.vec_ok) {
          for (int c = chunk; c < (2 * S) / 4; c += TPH) store_coord4<CT>(dst_o + 4 * c, my_go + 4 * c);
          for (int c = chunk; c < S / 4; c += TPH) store_coord4<CT>(dst_l + 4 * c, my_ga + 4 * c);
        } else {
          for (int i = chunk; i < 2 * S; i += TPH) dst_o[i] = from_f32<CT>(my_go[i]);
          for (int i = chunk; i < S; i += TPH) dst_l[i] = from_f32<CT>(my_ga[i]);
        }
      }
    }
  }
}

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <typename T, typename CT, int TPH, int MODE>
static int launch_fused(const FusedProblem& f, bool bwd, cudaStream_t st, const char* what) {
  constexpr int ROWS = kFusedThreads / TPH;
  FusedArgs a{};
  a.value = f.value; a.shapes = f.shapes; a.starts = f.starts; a.offsets = f.offsets;
  a.logits = f.logits; a.ref = f.ref; a.hit_bits = f.hit_bits;
  a.out = f.out; a.g_out = f.g_out; a.g_value = f.g_value; a.g_offsets = f.g_offsets;
  a.g_logits = f.g_logits; a.acc_scale = f.acc_scale; a.acc_half = f.acc_half;
  a.g_tail = f.g_tail; a.tail_copies = f.g_tail ? f.tail_copies : 0; a.tail_px = f.tail_px;
  a.coarse_rec = (bwd && MODE == MODE_SCA && sca_coarse_active(f)) ? static_cast<uint4*>(f.coarse_rec) : nullptr;
  a.bs = f.bs; a.groups = f.groups; a.Nk = f.Nk; a.M = f.M; a.Dh = f.Dh; a.L = f.L; a.P = f.P;
  a.D = f.D; a.Nq = f.Nq; a.clamp = f.clamp;
  if (f.M > ROWS)
    return set_error(MSDA_ERR_UNSUPPORTED, "%s: %d heads do not fit a %d-row tile", what, f.M, ROWS);
  const int S = (MODE == MODE_TSA ? f.groups : 1) * f.L * f.P;
  // short rows (TSA / decoder: 4-8 samples): several rows per lane group per tile
  a.rpt = S >= 32 ? 1 : (S >= 16 ? 2 : 4);
  a.qpt = (ROWS / f.M) * a.rpt;
  a.bev_w = 0; a.bev_h = 0; a.tile_w = a.qpt; a.tile_h = 0; a.tiles_x = 0;
  if (f.bev_w > 0 && f.Nq % f.bev_w == 0 && a.qpt >= 2) {
    a.bev_w = f.bev_w;
    a.bev_h = f.Nq / f.bev_w;
    a.tile_w = (a.qpt % 8 == 0 && a.qpt >= 32) ? 8 : (a.qpt % 4 == 0) ? 4 : (a.qpt % 2 == 0 ? 2 : a.qpt);
    a.tile_h = a.qpt / a.tile_w;
    a.tiles_x = (a.bev_w + a.tile_w - 1) / a.tile_w;
    a.tiles_per_sample = a.tiles_x * ((a.bev_h + a.tile_h - 1) / a.tile_h);
  } else {
    a.tiles_per_sample = (f.Nq + a.qpt - 1) / a.qpt;
  }
  a.num_tiles = (long long)f.bs * a.tiles_per_sample;
  if (a.num_tiles <= 0) return MSDA_OK;
  // the sample records carry 32-bit byte offsets from the base of a sample's value maps
  if ((unsigned long long)(MODE == MODE_TSA ? f.groups : 1) * f.Nk * f.M * f.Dh * sizeof(T) >= (1ull << 32))
    return set_error(MSDA_ERR_UNSUPPORTED, "%s: value maps of %d pixels exceed the 4 GiB offset range", what, f.Nk);
  a.off_stride = f.off_stride ? f.off_stride : (long long)f.M * S * 2;
  a.log_stride = f.log_stride ? f.log_stride : (long long)f.M * S;
  if (a.off_stride % 2 != 0)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: the offsets stride must be even", what);
  a.vec_ok = (S % 4 == 0) && ((reinterpret_cast<uintptr_t>(f.g_offsets) | reinterpret_cast<uintptr_t>(f.g_logits)) % 16 == 0) &&
             ((a.off_stride * sizeof(CT)) % 16 == 0) && ((a.log_stride * sizeof(CT)) % 16 == 0);
  if ((reinterpret_cast<uintptr_t>(f.offsets) % 16) != 0 || (reinterpret_cast<uintptr_t>(f.logits) % 16) != 0)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: offsets / logits must be 16-byte aligned", what);
  const size_t smem = (size_t)ROWS * ((2 * S + 4) + (S + 4)) * (bwd ? 2 : 1) * sizeof(float) +
                      (size_t)ROWS * ((bwd && !(sizeof(T) == 2 && f.acc_half) ? 12 : 8) * TPH + 4) * sizeof(uint32_t) +
                      (size_t)S * sizeof(uint32_t);
  if (smem > 200 * 1024)
    return set_error(MSDA_ERR_UNSUPPORTED, "%s: %d samples per row need %zu bytes of shared memory", what, S, smem);
  static const int dbg = [] { const char* v = getenv("MSDA_DEBUG"); return v ? atoi(v) : 0; }();
  a.debug = dbg;
  constexpr bool kHalfOk = sizeof(T) == 2;
  // the base-encoder shapes of a 16-bit model run instances with those shapes as compile-time constants
  constexpr int kFix = (FUSED_FIXED_SHAPES != 0 && sizeof(T) == 2 && sizeof(CT) == 2 && TPH == 4) ? 1 : 0;
  const bool fixed = shape_is_fixed<MODE, kFix>(a);
  auto kfn = !bwd ? (fixed ? fused_fwd_kernel<T, CT, TPH, MODE, kFix> : fused_fwd_kernel<T, CT, TPH, MODE, 0>)
                  : (kHalfOk && f.acc_half
                         ? (fixed ? fused_bwd_kernel<T, CT, TPH, MODE, kHalfOk, kFix> : fused_bwd_kernel<T, CT, TPH, MODE, kHalfOk, 0>)
                         : fused_bwd_kernel<T, CT, TPH, MODE, false, 0>);
  cudaError_t e = cudaSuccess;
  if (smem > 48 * 1024)
    e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_error(MSDA_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  // MSDA_FUSED_WAVES=k > 0: persistent grid of k CTAs per SM striding over the tiles;
  // default: one tile per CTA (the hardware scheduler balances the uneven tiles).
  static const int waves = [] { const char* v = getenv("MSDA_FUSED_WAVES"); return v ? atoi(v) : 0; }();
  long long grid_ll = a.num_tiles;
  if (waves > 0 && grid_ll > (long long)sm_count() * waves) grid_ll = (long long)sm_count() * waves;
  if (grid_ll > 0x7fffffffLL) return set_error(MSDA_ERR_UNSUPPORTED, "%s: too many tiles", what);
  kfn<<<(unsigned)grid_ll, kFusedThreads, smem, st>>>(a);
  count_launch();
  return check_launch(what);
}

template <typename T, typename CT, int MODE>
static int dispatch_tph(const FusedProblem& f, bool bwd, cudaStream_t st, const char* what) {
  constexpr int VEC = Vec16<T>::N;
  if (f.Dh % VEC != 0)
    return set_error(MSDA_ERR_UNSUPPORTED, "%s: head_dim %d must be a multiple of %d for this dtype "
                     "(use the op-boundary msda_fwd/msda_bwd, which has a generic path)", what, f.Dh, VEC);
  switch (f.Dh / VEC) {
#ifndef FUSED_DEV_ONLY               // development builds compile the base-config instance only
    case 1: return launch_fused<T, CT, 1, MODE>(f, bwd, st, what);
    case 2: return launch_fused<T, CT, 2, MODE>(f, bwd, st, what);
#endif
    case 4: return launch_fused<T, CT, 4, MODE>(f, bwd, st, what);
#ifndef FUSED_DEV_ONLY
    case 8: return launch_fused<T, CT, 8, MODE>(f, bwd, st, what);
    case 16: return launch_fused<T, CT, 16, MODE>(f, bwd, st, what);
#endif
    default: break;
  }
  return set_error(MSDA_ERR_UNSUPPORTED, "%s: head_dim %d not supported by the fused kernels", what, f.Dh);
}

template <int MODE>
static int dispatch_dtype(const FusedProblem& f, bool bwd, cudaStream_t st, const char* what) {
  const uintptr_t al = reinterpret_cast<uintptr_t>(f.value) | reinterpret_cast<uintptr_t>(f.out) |
                       reinterpret_cast<uintptr_t>(f.g_out) | reinterpret_cast<uintptr_t>(f.g_value) |
                       reinterpret_cast<uintptr_t>(f.ref);
  if (al & 15u) return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: tensors must be 16-byte aligned", what);
  if (f.coord_dtype != MSDA_F32 && f.coord_dtype != f.value_dtype)
    return set_error(MSDA_ERR_UNSUPPORTED, "%s: offsets / logits must be fp32 or the value dtype", what);
  const bool c32 = f.coord_dtype == MSDA_F32;
#ifdef FUSED_DEV_ONLY
  if (f.value_dtype == MSDA_BF16 && !c32) return dispatch_tph<__nv_bfloat16, __nv_bfloat16, MODE>(f, bwd, st, what);
  return set_error(MSDA_ERR_UNSUPPORTED, "%s: development build (bf16 / bf16 only)", what);
#else
  switch (f.value_dtype) {
    case MSDA_F32: return dispatch_tph<float, float, MODE>(f, bwd, st, what);
    case MSDA_BF16: return c32 ? dispatch_tph<__nv_bfloat16, float, MODE>(f, bwd, st, what)
                               : dispatch_tph<__nv_bfloat16, __nv_bfloat16, MODE>(f, bwd, st, what);
    default: return c32 ? dispatch_tph<__half, float, MODE>(f, bwd, st, what)
                        : dispatch_tph<__half, __half, MODE>(f, bwd, st, what);
  }
#endif
}

int launch_sca_fwd(const FusedProblem& f, cudaStream_t st) { return dispatch_dtype<MODE_SCA>(f, false, st, "sca_fwd"); }
bool sca_coarse_active(const FusedProblem& f) {
  return f.coarse_rec != nullptr && f.hit_index != nullptr && f.hit_count != nullptr &&
         coarse_supported(f.Dh, f.P, f.value_dtype);
}
int launch_sca_bwd(const FusedProblem& f, cudaStream_t st) {
  if (int rc = dispatch_dtype<MODE_SCA>(f, true, st, "sca_bwd")) return rc;
  if (!sca_coarse_active(f)) return MSDA_OK;
  // second pass: grad_value of the coarse levels from the records, on the tensor cores
  return launch_coarse_scatter(f.coarse_rec, f.hit_index, f.hit_count, f.g_out, f.g_value, f.acc_scale, f.acc_half,
                               f.shapes, f.starts, f.bs, f.groups, f.Nq, f.Nk, f.M, f.Dh, f.L, f.P, f.value_dtype, st);
}
int launch_tsa_fwd(const FusedProblem& f, cudaStream_t st) { return dispatch_dtype<MODE_TSA>(f, false, st, "tsa_fwd"); }
int launch_tsa_bwd(const FusedProblem& f, cudaStream_t st) { return dispatch_dtype<MODE_TSA>(f, true, st, "tsa_bwd"); }

}  // namespace msda
