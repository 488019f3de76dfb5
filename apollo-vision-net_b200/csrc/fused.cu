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
// Work decomposition (both, forward and backward).  The queries are cut into tiles of
// QPT = ROWS / M queries -- a TW x TH patch of the BEV grid when the queries lie on one (so the
// tile's samples fall on neighbouring pixels and share L1 lines), consecutive queries otherwise.
// The grid is persistent: a few CTAs per SM, each walking runs of consecutive tiles.  Per tile:
//   1. every (query, head) row's raw Linear outputs (offsets, logits) are fetched by one TMA bulk
//      copy per row (cp.async.bulk -> mbarrier) into PADDED shared-memory rows (bank-conflict
//      free for the per-row reads that follow);
//   2. a warp-cooperative pass turns them into softmax weights and normalised offsets once per
//      row (not once per lane or per camera);
//   3. every thread owns 16 bytes of one head's channels and walks the row's samples, camera
//      outermost, one 128-bit read-only load per bilinear corner.
// The backward additionally reduces the location / weight gradients over the head's lanes with
// warp shuffles, accumulates them over cameras in shared memory, finishes the softmax backward
// there, writes both gradient rows back with TMA bulk stores, and sends grad_value to an fp32
// accumulator with 16-byte vector reductions laid out so that one warp instruction covers whole
// 32-byte sectors.
#include <cstdlib>
#include "msda_common.cuh"
#include "msda_host.h"

namespace msda {

constexpr int kFusedThreads = 256;
enum { MODE_SCA = 0, MODE_TSA = 1 };

template <int TPH>
__device__ __forceinline__ float lanes_sum(float v, unsigned gmask) {
#pragma unroll
  for (int o = TPH / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
  return v;
}

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :: "l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

struct FusedArgs {
  const void* value;
  const int64_t* shapes;
  const int64_t* starts;
  const float* offsets;
  const float* logits;
  const float* ref;
  const uint32_t* hit_bits;
  void* out;
  const void* g_out;
  float* g_value;
  float* g_offsets;
  float* g_logits;
  int bs, groups, Nk, M, Dh, L, P, D, Nq;
  float clamp;
  int bev_w, bev_h;            // query grid (0 = queries are not on a grid)
  int tile_w, tile_h;          // 2-D tile (tile_h == 0: linear runs of qpt queries)
  int tiles_x, tiles_y;
  int qpt;                     // queries per tile
  int rpt;                     // rows each thread group walks per tile (tile rows = ROWS * rpt)
  int tiles_per_sample;
  long long num_tiles;         // bs * tiles_per_sample
  int run;                     // consecutive tiles a CTA takes at a time
  int bulk;                    // rows can use TMA bulk copies (16-byte granularity holds)
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

__device__ __forceinline__ int tile_active_queries(const FusedArgs& a, long long t) {
  const int tt = (int)(t % a.tiles_per_sample);
  if (a.tile_h > 0) {
    const int ax = min(a.tile_w, a.bev_w - (tt % a.tiles_x) * a.tile_w);
    const int ay = min(a.tile_h, a.bev_h - (tt / a.tiles_x) * a.tile_h);
    return ax * ay;
  }
  return min(a.qpt, a.Nq - tt * a.qpt);
}

struct Smem {
  float* raw_off;  // [ROWS][2S]  region A: raw offsets as copied by TMA (contiguous rows)
  float* raw_w;    // [ROWS][S]   region A: raw logits
  float* off;      // [ROWS][po]  region B: normalised offsets, padded rows
  float* w;        // [ROWS][pw]  region B: softmax weights, padded rows
  float* ga;       // [ROWS][pw]  (bwd, reuses region A) grad wrt softmax output -> grad wrt logits
  float* go;       // [ROWS][po]  (bwd, reuses region A) grad wrt raw offsets
  int po, pw;
};

__device__ __forceinline__ Smem carve_smem(unsigned char* base, int rows, int S) {
  Smem sm;
  sm.po = 2 * S + 4;
  sm.pw = S + 4;
  float* A = reinterpret_cast<float*>(base);
  float* B = A + (size_t)rows * (sm.po + sm.pw);
  sm.raw_off = A;
  sm.raw_w = A + (size_t)rows * 2 * S;
  sm.go = A;
  sm.ga = A + (size_t)rows * sm.po;
  sm.off = B;
  sm.w = B + (size_t)rows * sm.po;
  return sm;
}

// Stage one tile: a few large TMA bulk copies bring the raw Linear outputs of the tile's rows
// (contiguous in global memory per grid row of the tile) into region A; one cooperative pass
// writes softmax weights and normalised offsets into the padded rows of region B.
template <bool BWD>
__device__ __forceinline__ void stage_tile(const FusedArgs& a, const LevelTable& lv, uint64_t* bar,
                                           uint32_t phase, const Smem& sm, long long t,
                                           int rows_tile, int S, int LP) {
  const int tid = threadIdx.x;
  const int b = (int)(t / a.tiles_per_sample);
  const int tt = (int)(t % a.tiles_per_sample);
  int slabs, slab_q, q0, q_step, active_q;
  if (a.tile_h > 0) {
    const int x0 = (tt % a.tiles_x) * a.tile_w, y0 = (tt / a.tiles_x) * a.tile_h;
    slab_q = min(a.tile_w, a.bev_w - x0);
    slabs = min(a.tile_h, a.bev_h - y0);
    q0 = y0 * a.bev_w + x0;
    q_step = a.bev_w;
  } else {
    slab_q = min(a.qpt, a.Nq - tt * a.qpt);
    slabs = 1;
    q0 = tt * a.qpt;
    q_step = 0;
  }
  active_q = slabs * slab_q;
  const int slab_rows_full = (a.tile_h > 0 ? a.tile_w : a.qpt) * a.M;   // local rows per slab
  if (a.bulk) {
    fence_proxy_async_smem();      // region A was last used through the generic proxy
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(bar, (uint32_t)active_q * (uint32_t)a.M * (uint32_t)S * 12u);
      for (int sl = 0; sl < slabs; ++sl) {
        const long long grow = ((long long)b * a.Nq + q0 + (long long)sl * q_step) * a.M;
        const uint32_t nrow = (uint32_t)slab_q * (uint32_t)a.M;
        bulk_g2s(sm.raw_off + (size_t)sl * slab_rows_full * 2 * S, a.offsets + grow * S * 2,
                 nrow * (uint32_t)S * 8u, bar);
        bulk_g2s(sm.raw_w + (size_t)sl * slab_rows_full * S, a.logits + grow * S,
                 nrow * (uint32_t)S * 4u, bar);
      }
    }
    mbar_wait(bar, phase);
  } else {
    __syncthreads();
    for (int sl = 0; sl < slabs; ++sl) {
      const long long grow = ((long long)b * a.Nq + q0 + (long long)sl * q_step) * a.M;
      const int nrow = slab_q * a.M;
      const float* go = a.offsets + grow * S * 2;
      const float* gl = a.logits + grow * S;
      float* so = sm.raw_off + (size_t)sl * slab_rows_full * 2 * S;
      float* sw = sm.raw_w + (size_t)sl * slab_rows_full * S;
      for (int i = tid; i < nrow * 2 * S; i += blockDim.x) so[i] = go[i];
      for (int i = tid; i < nrow * S; i += blockDim.x) sw[i] = gl[i];
    }
    __syncthreads();
  }
  // softmax per (row, segment), LPp lanes per segment: raw logits (A) -> weights (B).
  // Rows outside the grid hold stale data: harmless, nobody reads their results.
  int LPp = 1;
  while (LPp < LP && LPp < 32) LPp <<= 1;
  const int seg_row = S / LP;
  const int nseg = rows_tile * seg_row;
  const int lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const int spw = 32 / LPp;
  const int j = lane % LPp;
  for (int base = warp * spw; base < nseg; base += nwarp * spw) {
    const int sgi = base + lane / LPp;
    const bool ok = sgi < nseg;
    const int r = ok ? sgi / seg_row : 0, sg = ok ? sgi % seg_row : 0;
    const float* src = sm.raw_w + (size_t)r * S + sg * LP;
    float* dst = sm.w + (size_t)r * sm.pw + sg * LP;
    float mx = -INFINITY;
    for (int i = j; i < LP; i += LPp) {
      float v = src[i];
      if (a.clamp >= 0.f) v = fminf(fmaxf(v, -a.clamp), a.clamp);
      mx = fmaxf(mx, v);
    }
    for (int o = LPp / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int i = j; i < LP; i += LPp) {
      float v = src[i];
      if (a.clamp >= 0.f) v = fminf(fmaxf(v, -a.clamp), a.clamp);
      const float e = expf(v - mx);
      if (ok) dst[i] = e;
      sum += e;
    }
    for (int o = LPp / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (ok)
      for (int i = j; i < LP; i += LPp) dst[i] = dst[i] / sum;
  }
  // offsets / (W_l, H_l): raw (A) -> padded rows (B), 16 bytes (two samples) at a time; a warp
  // takes whole rows, so the only integer division left is by P (a shift when P is a power of 2)
  if ((S & 1) == 0 && (a.P & 1) == 0) {
    const int cpr = S / 2;                               // float4 chunks per row
    const int pshift = (a.P & (a.P - 1)) == 0 ? 31 - __clz(a.P) : -1;
    for (int r = warp; r < rows_tile; r += nwarp) {
      const float4* src = reinterpret_cast<const float4*>(sm.raw_off + (size_t)r * 2 * S);
      float4* dst = reinterpret_cast<float4*>(sm.off + (size_t)r * sm.po);
      for (int c = lane; c < cpr; c += 32) {
        int sl = 2 * c;                                  // sample index within the row
        while (sl >= LP) sl -= LP;                       // at most Q - 1 iterations
        const int l = pshift >= 0 ? (sl >> pshift) : sl / a.P;
        const float dw = (float)lv.w[l], dh = (float)lv.h[l];
        float4 v = src[c];
        v.x = v.x / dw; v.y = v.y / dh; v.z = v.z / dw; v.w = v.w / dh;
        dst[c] = v;
      }
    }
  } else {
    const int per_row = 2 * S;
    for (int i = tid; i < rows_tile * per_row; i += blockDim.x) {
      const int r = i / per_row, e = i % per_row;
      const int l = ((e >> 1) % LP) / a.P;
      const float d = (e & 1) ? (float)lv.h[l] : (float)lv.w[l];
      sm.off[(size_t)r * sm.po + e] = sm.raw_off[i] / d;
    }
  }
  __syncthreads();
  if (BWD) {       // region A becomes the gradient accumulators
    for (int i = tid; i < rows_tile * (sm.po + sm.pw); i += blockDim.x) sm.go[i] = 0.f;
    __syncthreads();
  }
}

template <typename T, int TPH, int MODE>
__global__ void __launch_bounds__(kFusedThreads)
fused_fwd_kernel(const FusedArgs a) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int V2 = VEC / 2;
  constexpr int ROWS = kFusedThreads / TPH;
  __shared__ LevelTable lv;
  __shared__ __align__(8) uint64_t bar;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  const int LP = a.L * a.P;
  const int S = (MODE == MODE_TSA) ? a.groups * LP : LP;
  const Smem sm = carve_smem(dyn_smem, ROWS * a.rpt, S);
  const int tid = threadIdx.x;
  const int rows_tile = a.qpt * a.M;

  load_level_table(lv, a.shapes, a.starts, a.L);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  __syncthreads();

  const int chunk = tid % TPH;

  const int pix_stride = a.M * a.Dh;
  const size_t batch_stride = (size_t)a.Nk * pix_stride;
  uint32_t phase = 0;

  for (long long base = (long long)blockIdx.x * a.run; base < a.num_tiles;
       base += (long long)gridDim.x * a.run) {
    const long long tend = min(base + (long long)a.run, a.num_tiles);
    for (long long t = base; t < tend; ++t) {
      stage_tile<false>(a, lv, &bar, phase, sm, t, rows_tile, S, LP);
      phase ^= 1;

      for (int rk = 0; rk < a.rpt; ++rk) {
      const int r_local = rk * ROWS + tid / TPH;
      const bool in_tile = r_local < rows_tile;
      const int lq = in_tile ? r_local / a.M : 0;
      const int m = in_tile ? r_local % a.M : 0;
      int b;
      const int q = tile_query(a, t, lq, b);
      if (in_tile && q >= 0) {
        const T* vhead = static_cast<const T*>(a.value) + (size_t)m * a.Dh + chunk * VEC;
        const float* my_off = sm.off + (size_t)r_local * sm.po;
        const float* my_w = sm.w + (size_t)r_local * sm.pw;
        float2 acc[V2];
#pragma unroll
        for (int i = 0; i < V2; ++i) acc[i] = make_float2(0.f, 0.f);

        auto sample = [&](const T* lbase, int H, int W, float lx, float ly, float w) {
          const Corners c = corner_setup(lx, ly, H, W, pix_stride);
          const uint4 u00 = ldg128(lbase + c.o00);
          const uint4 u01 = ldg128(lbase + c.o01);
          const uint4 u10 = ldg128(lbase + c.o10);
          const uint4 u11 = ldg128(lbase + c.o11);
          const float2 w00 = splat2(w * c.w00), w01 = splat2(w * c.w01);
          const float2 w10 = splat2(w * c.w10), w11 = splat2(w * c.w11);
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
        };

        float scale;
        if (MODE == MODE_SCA) {
          uint32_t hits = a.hit_bits[q];                             // batch element 0 decides (quirk 1)
          const int cnt = __popc(a.hit_bits[(size_t)b * a.Nq + q]);  // the divisor is per sample
          scale = (float)(cnt > 0 ? cnt : 1);
          while (hits) {
            const int cam = __ffs(hits) - 1;
            hits &= hits - 1;
            const float2* rc = reinterpret_cast<const float2*>(
                a.ref + (((size_t)cam * a.bs + b) * a.Nq + q) * a.D * 2);
            const T* vcam = vhead + ((size_t)b * a.groups + cam) * batch_stride;
            for (int l = 0; l < a.L; ++l) {
              const int H = lv.h[l], W = lv.w[l];
              const T* lbase = vcam + (size_t)lv.start[l] * pix_stride;
              int z = 0;                                             // point index p = k*D + z (quirk 5)
#pragma unroll 4
              for (int p = 0; p < a.P; ++p) {
                const int s = l * a.P + p;
                const float2 r = __ldg(rc + z);
                z = (z + 1 == a.D) ? 0 : z + 1;
                const float2 o = *reinterpret_cast<const float2*>(my_off + 2 * s);
                sample(lbase, H, W, r.x + o.x, r.y + o.y, my_w[s]);
              }
            }
          }
        } else {
          scale = (float)a.groups;
          for (int j = 0; j < a.groups; ++j) {
            const T* vb = vhead + ((size_t)b * a.groups + j) * batch_stride;
            for (int l = 0; l < a.L; ++l) {
              const int H = lv.h[l], W = lv.w[l];
              const T* lbase = vb + (size_t)lv.start[l] * pix_stride;
              const float2 r = __ldg(reinterpret_cast<const float2*>(
                  a.ref + ((((size_t)b * a.groups + j) * a.Nq + q) * a.L + l) * 2));
#pragma unroll 4
              for (int p = 0; p < a.P; ++p) {
                const int s = (j * a.L + l) * a.P + p;
                const float2 o = *reinterpret_cast<const float2*>(my_off + 2 * s);
                sample(lbase, H, W, r.x + o.x, r.y + o.y, my_w[s]);
              }
            }
          }
        }
        float out[VEC];
#pragma unroll
        for (int i = 0; i < V2; ++i) {
          out[2 * i] = acc[i].x;
          out[2 * i + 1] = acc[i].y;
        }
        if (scale == 2.f) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) out[i] *= 0.5f;
        } else if (scale != 1.f) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) out[i] = out[i] / scale;
        }
        T* o = static_cast<T*>(a.out) + (((size_t)b * a.Nq + q) * a.M + m) * a.Dh + chunk * VEC;
        *reinterpret_cast<uint4*>(o) = Vec16<T>::pack(out);
      }
      }
      __syncthreads();    // everyone is done with this tile's shared rows
    }
  }
}

template <typename T, int TPH, int MODE>
__global__ void __launch_bounds__(kFusedThreads)
fused_bwd_kernel(const FusedArgs a) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int V2 = VEC / 2;
  constexpr int ROWS = kFusedThreads / TPH;
  __shared__ LevelTable lv;
  __shared__ __align__(8) uint64_t bar;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  const int LP = a.L * a.P;
  const int S = (MODE == MODE_TSA) ? a.groups * LP : LP;
  const Smem sm = carve_smem(dyn_smem, ROWS * a.rpt, S);
  const int tid = threadIdx.x;
  const int rows_tile = a.qpt * a.M;

  load_level_table(lv, a.shapes, a.starts, a.L);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  __syncthreads();

  const int chunk = tid % TPH;
  const int lane = tid & 31;

  const int pix_stride = a.M * a.Dh;
  const size_t batch_stride = (size_t)a.Nk * pix_stride;
  // Channel ownership for the grad_value scatter.  With 8 channels per lane (16-bit value) lane c
  // scatters channels [4c, 4c+4) and [4*TPH + 4c, 4*TPH + 4c + 4), so that the TPH lanes of a head
  // cover 16*TPH contiguous bytes per reduction instruction (whole sectors); the dot products
  // use the natural [8c, 8c+8) ownership of the 128-bit value loads.
  constexpr int SC = VEC / 4;                       // scatter instructions per corner
  uint32_t phase = 0;

  for (long long base = (long long)blockIdx.x * a.run; base < a.num_tiles;
       base += (long long)gridDim.x * a.run) {
    const long long tend = min(base + (long long)a.run, a.num_tiles);
    for (long long t = base; t < tend; ++t) {
      stage_tile<true>(a, lv, &bar, phase, sm, t, rows_tile, S, LP);
      phase ^= 1;

      // Every lane walks the same loops (the lane sums below use full-warp shuffles); lanes
      // without work carry weight zero and touch nothing.
      for (int rk = 0; rk < a.rpt; ++rk) {
      const int r_raw = rk * ROWS + tid / TPH;
      const bool in_tile = r_raw < rows_tile;
      const int r_local = in_tile ? r_raw : 0;
      const int lq = r_local / a.M;
      const int m = r_local % a.M;
      int b;
      int q = tile_query(a, t, lq, b);
      const bool live = in_tile && q >= 0;
      if (!live) { q = 0; b = 0; }
      {
        const size_t head_off = (size_t)m * a.Dh;
        const T* vhead = static_cast<const T*>(a.value) + head_off + chunk * VEC;
        float* ghead = a.g_value + head_off + 4 * chunk;
        const float* my_off = sm.off + (size_t)r_local * sm.po;
        const float* my_w = sm.w + (size_t)r_local * sm.pw;
        float* my_ga = sm.ga + (size_t)r_local * sm.pw;
        float* my_go = sm.go + (size_t)r_local * sm.po;

        uint32_t hits = 0;
        float scale;
        if (MODE == MODE_SCA) {
          hits = live ? a.hit_bits[q] : 0u;
          const int cnt = __popc(a.hit_bits[(size_t)b * a.Nq + q]);
          scale = (float)(cnt > 0 ? cnt : 1);
        } else {
          scale = (float)a.groups;
        }

        float2 g[V2], gs[V2];
        {
          const T* grow_ptr = static_cast<const T*>(a.g_out) + (((size_t)b * a.Nq + q) * a.M + m) * a.Dh;
          Vec16<T>::unpack2(ldg128(grow_ptr + chunk * VEC), g);
          if (VEC == 4) {
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

        // One sample against one value map: scatters grad_value, accumulates the row's
        // location / weight gradients in shared memory (one writer per (row, sample)).
        auto sample = [&](size_t boff, int H, int W, float lx, float ly, float w, int s, bool mine) {
          const Corners c = corner_setup(lx, ly, H, W, pix_stride);
          const T* vb = vhead + boff;
          float* gb = ghead + boff;
          const uint4 u00 = ldg128(vb + c.o00);
          const uint4 u01 = ldg128(vb + c.o01);
          const uint4 u10 = ldg128(vb + c.o10);
          const uint4 u11 = ldg128(vb + c.o11);
          auto scatter = [&](int off, float cw) {
            const float aw = w * cw;
            if (aw == 0.f) return;                       // invalid corner, or a zero contribution
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
          ga = lanes_sum<TPH>(ga, 0xffffffffu);
          gx = lanes_sum<TPH>(gx, 0xffffffffu);
          gy = lanes_sum<TPH>(gy, 0xffffffffu);
          if (chunk == 0 && mine) {
            // d loc / d offset = 1 / (W_l, H_l) cancels the (W_l, H_l) factor of d pixel / d loc
            my_ga[s] += ga;
            float2* go2 = reinterpret_cast<float2*>(my_go + 2 * s);
            float2 cur = *go2;
            cur.x += w * gx;
            cur.y += w * gy;
            *go2 = cur;
          }
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
            for (int l = 0; l < a.L; ++l) {
              const int H = lv.h[l], W = lv.w[l];
              const size_t loff = coff + (size_t)lv.start[l] * pix_stride;
              int z = 0;
#pragma unroll 2
              for (int p = 0; p < a.P; ++p) {
                const int s = l * a.P + p;
                const float2 r = __ldg(rc + z);
                z = (z + 1 == a.D) ? 0 : z + 1;
                const float2 o = *reinterpret_cast<const float2*>(my_off + 2 * s);
                sample(loff, H, W, r.x + o.x, r.y + o.y, mine ? my_w[s] : 0.f, s, mine);
              }
            }
          }
        } else {
          for (int j = 0; j < a.groups; ++j) {
            const size_t boff = ((size_t)b * a.groups + j) * batch_stride;
            for (int l = 0; l < a.L; ++l) {
              const int H = lv.h[l], W = lv.w[l];
              const size_t loff = boff + (size_t)lv.start[l] * pix_stride;
              const float2 r = __ldg(reinterpret_cast<const float2*>(
                  a.ref + ((((size_t)b * a.groups + j) * a.Nq + q) * a.L + l) * 2));
#pragma unroll 2
              for (int p = 0; p < a.P; ++p) {
                const int s = (j * a.L + l) * a.P + p;
                const float2 o = *reinterpret_cast<const float2*>(my_off + 2 * s);
                sample(loff, H, W, r.x + o.x, r.y + o.y, live ? my_w[s] : 0.f, s, live);
              }
            }
          }
        }
      }
      }
      __syncthreads();

      // softmax backward per (row, segment): g_logit = a * (ga - sum_t a_t ga_t); zero where the
      // raw logit was clamped.  LPp lanes per segment.
      {
        int LPp = 1;
        while (LPp < LP && LPp < 32) LPp <<= 1;
        const int seg_row = S / LP;
        const int nseg = rows_tile * seg_row;
        const int warp = tid >> 5, nwarp = blockDim.x >> 5;
        const int spw = 32 / LPp;
        const int j = lane % LPp;
        for (int sb = warp * spw; sb < nseg; sb += nwarp * spw) {
          const int sgi = sb + lane / LPp;
          const bool ok = sgi < nseg;
          const int r = ok ? sgi / seg_row : 0, sg = ok ? sgi % seg_row : 0;
          const float* w = sm.w + (size_t)r * sm.pw + sg * LP;
          float* ga = sm.ga + (size_t)r * sm.pw + sg * LP;
          float dot = 0.f;
          for (int i = j; i < LP; i += LPp) dot = fmaf(w[i], ga[i], dot);
          for (int o = LPp / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
          if (ok) {
            bool clamped_row = false;
            const float* raw = nullptr;
            if (a.clamp >= 0.f) {
              int rb;
              const int rq = tile_query(a, t, r / a.M, rb);
              clamped_row = rq >= 0;
              if (clamped_row)
                raw = a.logits + (((long long)rb * a.Nq + rq) * a.M + r % a.M) * S + sg * LP;
            }
            for (int i = j; i < LP; i += LPp) {
              float gl = w[i] * (ga[i] - dot);
              if (clamped_row) {
                const float v = raw[i];
                if (v < -a.clamp || v > a.clamp) gl = 0.f;
              }
              ga[i] = gl;
            }
          }
        }
      }
      __syncthreads();

      // gradient rows back to global memory: coalesced 16-byte stores from the padded rows
      {
        const int cpr_o = (2 * S) / 4, cpr_l = S / 4;           // 16-byte chunks per row
        if (a.bulk) {
          for (int i = tid; i < rows_tile * (cpr_o + cpr_l); i += blockDim.x) {
            const int r = i / (cpr_o + cpr_l), c = i % (cpr_o + cpr_l);
            int rb;
            const int rq = tile_query(a, t, r / a.M, rb);
            if (rq < 0) continue;
            const long long grow = ((long long)rb * a.Nq + rq) * a.M + r % a.M;
            if (c < cpr_o) {
              const float4 v = *reinterpret_cast<const float4*>(sm.go + (size_t)r * sm.po + 4 * c);
              *reinterpret_cast<float4*>(a.g_offsets + grow * S * 2 + 4 * c) = v;
            } else {
              const int cc = c - cpr_o;
              const float4 v = *reinterpret_cast<const float4*>(sm.ga + (size_t)r * sm.pw + 4 * cc);
              *reinterpret_cast<float4*>(a.g_logits + grow * S + 4 * cc) = v;
            }
          }
        } else {
          for (int i = tid; i < rows_tile * 3 * S; i += blockDim.x) {
            const int r = i / (3 * S), e = i % (3 * S);
            int rb;
            const int rq = tile_query(a, t, r / a.M, rb);
            if (rq < 0) continue;
            const long long grow = ((long long)rb * a.Nq + rq) * a.M + r % a.M;
            if (e < 2 * S) a.g_offsets[grow * S * 2 + e] = sm.go[(size_t)r * sm.po + e];
            else a.g_logits[grow * S + (e - 2 * S)] = sm.ga[(size_t)r * sm.pw + (e - 2 * S)];
          }
        }
      }
      __syncthreads();
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

template <typename T, int TPH, int MODE>
static int launch_fused(const FusedProblem& f, bool bwd, cudaStream_t st, const char* what) {
  constexpr int ROWS = kFusedThreads / TPH;
  FusedArgs a{};
  a.value = f.value; a.shapes = f.shapes; a.starts = f.starts; a.offsets = f.offsets;
  a.logits = f.logits; a.ref = f.ref; a.hit_bits = f.hit_bits;
  a.out = f.out; a.g_out = f.g_out; a.g_value = f.g_value; a.g_offsets = f.g_offsets;
  a.g_logits = f.g_logits;
  a.bs = f.bs; a.groups = f.groups; a.Nk = f.Nk; a.M = f.M; a.Dh = f.Dh; a.L = f.L; a.P = f.P;
  a.D = f.D; a.Nq = f.Nq; a.clamp = f.clamp;
  if (f.M > ROWS)
    return set_error(MSDA_ERR_UNSUPPORTED, "%s: %d heads do not fit a %d-row tile", what, f.M, ROWS);
  const int S = (MODE == MODE_TSA ? f.groups : 1) * f.L * f.P;
  // short rows (TSA / decoder: 4-8 samples) would leave a tile with a few microseconds of work
  // against its fixed cost (TMA round trip, barriers): give every thread group several rows
  a.rpt = S >= 32 ? 1 : (S >= 16 ? 2 : 4);
  a.qpt = (ROWS / f.M) * a.rpt;
  a.bev_w = 0; a.bev_h = 0; a.tile_w = a.qpt; a.tile_h = 0; a.tiles_x = 0; a.tiles_y = 0;
  if (f.bev_w > 0 && f.Nq % f.bev_w == 0 && a.qpt >= 2) {
    a.bev_w = f.bev_w;
    a.bev_h = f.Nq / f.bev_w;
    a.tile_w = (a.qpt % 8 == 0 && a.qpt >= 32) ? 8 : (a.qpt % 4 == 0) ? 4 : (a.qpt % 2 == 0 ? 2 : a.qpt);
    a.tile_h = a.qpt / a.tile_w;
    a.tiles_x = (a.bev_w + a.tile_w - 1) / a.tile_w;
    a.tiles_y = (a.bev_h + a.tile_h - 1) / a.tile_h;
    a.tiles_per_sample = a.tiles_x * a.tiles_y;
  } else {
    a.tiles_per_sample = (f.Nq + a.qpt - 1) / a.qpt;
  }
  a.num_tiles = (long long)f.bs * a.tiles_per_sample;
  if (a.num_tiles <= 0) return MSDA_OK;
  a.bulk = (S % 4 == 0) && ((reinterpret_cast<uintptr_t>(f.offsets) | reinterpret_cast<uintptr_t>(f.logits) |
                             reinterpret_cast<uintptr_t>(f.g_offsets) | reinterpret_cast<uintptr_t>(f.g_logits)) % 16 == 0);
  const size_t smem = (size_t)ROWS * a.rpt * ((2 * S + 4) + (S + 4)) * 2 * sizeof(float);
  if (smem > 200 * 1024)
    return set_error(MSDA_ERR_UNSUPPORTED, "%s: %d samples per row need %zu bytes of shared memory", what, S, smem);
  auto kfn = bwd ? fused_bwd_kernel<T, TPH, MODE> : fused_fwd_kernel<T, TPH, MODE>;
  cudaError_t e = cudaSuccess;
  if (smem > 48 * 1024)
    e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int per_sm = 1;
  if (e == cudaSuccess)
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, kFusedThreads, smem);
  if (e != cudaSuccess) return set_error(MSDA_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  if (per_sm < 1) per_sm = 1;
  const long long slots = (long long)sm_count() * per_sm;
  // MSDA_FUSED_PERSIST=k > 0: persistent grid, every resident CTA slot walks runs of k
  // consecutive tiles (stays inside one neighbourhood of the BEV grid); default: one tile per CTA.
  static const int persist = [] { const char* e = getenv("MSDA_FUSED_PERSIST"); return e ? atoi(e) : 0; }();
  long long grid_ll = a.num_tiles;
  a.run = 1;
  if (persist > 0) {
    a.run = persist;
    const long long want = (a.num_tiles + a.run - 1) / a.run;
    grid_ll = want < slots ? want : slots;
  }
  if (grid_ll > 0x7fffffffLL) return set_error(MSDA_ERR_UNSUPPORTED, "%s: too many tiles", what);
  const unsigned grid = (unsigned)grid_ll;
  kfn<<<grid, kFusedThreads, smem, st>>>(a);
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
                       reinterpret_cast<uintptr_t>(f.ref);
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
