// Small-sequence multi-head self-attention for the decoders.
//
// What it replaces: the two dense self-attentions of a MapTRv2 decoder layer
// (projects/mmdet3d_plugin/maptrv2/modules/decoder.py:129-188 -- "inter-vector" attention over the
// 350 vectors (50 one-to-one + 300 one-to-many, block mask) for each of the 20 point slots, then
// "intra-vector" attention over the 20 points of each vector) and the 900-query self-attention of the
// detection decoder (mmdet DetrTransformerDecoderLayer, configured at
// projects/configs/bevformer/bev_tiny_det_mapv2.py).  In the reference these run through mmcv's
// MultiheadAttention -> torch.nn.MultiheadAttention: softmax(q k^T / sqrt(d) + mask), dropout on the
// attention weights (p = 0.1 in training), times v.
//
// Work decomposition.  A PROBLEM is one (group, head): S tokens of head_dim Dh.  Tokens are addressed
// in place in the layer's (num_query, bs, C) activations -- token s of group g lives in row
//   s * seq_stride + (g / n_lo) * hi_stride + (g % n_lo) * lo_stride
// of the projected q / k / v matrices -- so the reference's permute + contiguous copies around the second
// self-attention (decoder.py:149-185) do not exist here.
//
//  * 16-bit models: tensor-core kernels (mma.sync m16n8k16, fp32 accumulate).  The problems are far too
//    small for a tcgen05 pipeline to pay (20 x 20 x 32 per problem in the intra-vector attention; the
//    whole 350-token attention of a layer is 2.5 GFLOP): what matters is that K / V of a problem are staged
//    ONCE per CTA in shared memory (padded rows, conflict-free ldmatrix), that a warp owns a 16-row tile
//    end to end (online softmax in registers, P never leaves the register file) and that several small
//    problems share a CTA.  The backward is two passes over the same code: pass A (a warp owns 16 queries)
//    produces dQ and delta = rowsum(dO . O); pass B (a warp owns 16 keys, the score tile is computed
//    transposed) produces dK and dV -- no atomics, no cross-warp reduction, nothing but lse (S floats per
//    problem) saved by the forward.
//  * fp32 models (and shapes the tensor-core path does not cover): a CTA owns 8 consecutive rows of one
//    problem, one warp per row; the other side of the score matrix streams through shared memory in tiles
//    of 32 tokens (coalesced loads, "lane = token" reads of a padded tile), plain FMA, the same passes.
//    This is the precision path the fp32 parity tests run against the oracle; it is 2.5-4x slower than the
//    library's fp32 attention at 350 tokens, so the module sends long fp32 sequences there
//    (modules/maptrv2_decoder.py: FusedSelfAttentionMixin).
//
// Dropout on the attention weights is counter-based (philox.cuh) and recomputed in the backward.  The
// mask is a function of (problem, q, k) built so that one Philox call serves the four elements
// {q, q ^ 8} x {k, k ^ 8}: those four sit in ONE thread's accumulator fragment both in the row-major pass
// and in the transposed pass, so neither pass needs a shuffle or a redundant call.
#include "msda_common.cuh"
#include "msda_host.h"
#include "philox.cuh"

namespace msda {
namespace {

constexpr int kMhaWarps = 8;
constexpr int kMhaThreads = kMhaWarps * 32;
constexpr int kMhaMaxSmem = 227 * 1024;

struct MhaArgs {
  const void *q, *k, *v, *o, *dout;
  void *out, *dq, *dk, *dv;
  long long ldq, ldk, ldv, ldo;
  long long lddq, lddk, lddv;      // row strides of the gradient matrices dq / dk / dv
  float* lse;                      // (G * H, S) natural-log sum of exp of the scaled scores (+inf: empty row)
  float* delta;                    // (G * H, S) rowsum(dO . O), written by pass A, read by pass B
  const uint32_t* mask_bits;       // (S, mask_words): bit k of row q set = q may not attend to k; or NULL
  const uint32_t* mask_bits_t;     // the same matrix transposed (row k, bit q)
  int mask_words;
  int G, H, S;
  long long seq_stride, hi_stride, lo_stride;
  int n_lo;
  float scale;
  const unsigned long long* key;   // (seed, step) on the device; NULL = no dropout
  unsigned long long* key_save;
  uint32_t site, thresh;           // keep an element when its 16-bit lane >= thresh = round(p * 65536)
  float inv_keep;
};

__device__ __forceinline__ long long group_base(const MhaArgs& a, int g) {
  return (long long)(g / a.n_lo) * a.hi_stride + (long long)(g % a.n_lo) * a.lo_stride;
}

// Dropout randomness.  One Philox2x32-10 call (Random123; 64 bits out) serves the four elements
// {q, q ^ 8} x {k, k ^ 8} of problem ph, 16 bits each: element (q, k) reads lane ((q >> 3) & 1) * 2 +
// ((k >> 3) & 1) and is kept when its lane >= round(p * 65536).  The 32-bit key is derived once per kernel
// from (seed, step, call site) with Philox4x32-10, so sites and steps draw independent streams.
__device__ __forceinline__ uint32_t mha_stream_key(unsigned long long seed, unsigned long long step, uint32_t site) {
  return philox4x32_10(make_uint4(site, (uint32_t)step, (uint32_t)(step >> 32), 0x4d484131u),
                       make_uint2((uint32_t)seed, (uint32_t)(seed >> 32))).x;
}
__device__ __forceinline__ uint2 mha_rand4(uint32_t ph, int q, int k, uint32_t key) {
  uint32_t c0 = (uint32_t)((q >> 4) * 8 + (q & 7)) | ((uint32_t)((k >> 4) * 8 + (k & 7)) << 16), c1 = ph;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi = __umulhi(0xD256D193u, c0), lo = 0xD256D193u * c0;
    c0 = hi ^ key ^ c1;
    c1 = lo;
    key += 0x9E3779B9u;
  }
  return make_uint2(c0, c1);
}
__device__ __forceinline__ uint32_t pick_lane(const uint2& r, int idx) {
  return ((idx & 2) ? r.y : r.x) >> (16 * (idx & 1)) & 0xffffu;
}
__device__ __forceinline__ bool mha_keep(uint32_t ph, int q, int k, uint32_t key, uint32_t thresh) {
  return pick_lane(mha_rand4(ph, q, k, key), ((q >> 3) & 1) * 2 + ((k >> 3) & 1)) >= thresh;
}
__device__ __forceinline__ float ex2(float x) {         // 2^x, ex2(-inf) = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---------------------------------------------------------------------------------------------------
// tensor-core path
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

template <typename T> struct Mma;
template <> struct Mma<__nv_bfloat16> {
  static __device__ __forceinline__ void run(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, "
                 "{%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
  }
};
template <> struct Mma<__half> {
  static __device__ __forceinline__ void run(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, "
                 "{%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    const __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
  }
};

// rows of a staged matrix are DH + 8 elements apart: 80 / 144 bytes, so the eight row addresses of an
// 8 x 8 ldmatrix tile fall into eight different 16-byte bank groups
template <int DH> struct TcShape {
  static constexpr int RS = DH + 8;
  static constexpr int CH = DH / 8;        // 16-byte chunks per row
  static constexpr int KS = DH / 16;       // k-steps of a product contracted over the head dim
  static constexpr int NT = DH / 8;        // n-tiles of a product whose columns are the head dim
};

// all S_pad rows of the (group, head) slices of `src` for the CTA's problems -> shared memory, zero rows
// beyond S (0 x garbage must not become NaN)
template <typename T, int DH>
__device__ __forceinline__ void stage_matrix(T* dst, const void* src, long long ld, const MhaArgs& a, int p0,
                                             int npc, int S_pad) {
  using Sh = TcShape<DH>;
  const int P = a.G * a.H;
  for (int i = threadIdx.x; i < npc * S_pad * Sh::CH; i += kMhaThreads) {
    const int ch = i % Sh::CH, s = (i / Sh::CH) % S_pad, pl = i / (Sh::CH * S_pad), p = p0 + pl;
    const T* from = static_cast<const T*>(src);
    const bool on = p < P && s < a.S;
    if (on) {
      const int g = p / a.H, h = p % a.H;
      from += (group_base(a, g) + (long long)s * a.seq_stride) * ld + h * DH + ch * 8;
    }
    // asynchronous 16-byte copy (zero-fill where the row does not exist): all of a thread's copies are in
    // flight together; the caller waits with stage_wait()
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst + ((size_t)pl * S_pad + s) * Sh::RS + ch * 8)),
                 "l"(from), "r"(on ? 16 : 0) : "memory");
  }
}
__device__ __forceinline__ void stage_wait() {
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
}

// 16 rows starting at token s0 of one problem -> the warp's staging tile
template <typename T, int DH>
__device__ __forceinline__ void load_tile(T* st, const void* src, long long ld, const MhaArgs& a, long long gbase,
                                          int h, int s0, int lane) {
  using Sh = TcShape<DH>;
#pragma unroll
  for (int i = lane; i < 16 * Sh::CH; i += 32) {
    const int r = i / Sh::CH, ch = i % Sh::CH, s = s0 + r;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (s < a.S) val = ldg128(static_cast<const T*>(src) + (gbase + (long long)s * a.seq_stride) * ld + h * DH + ch * 8);
    *reinterpret_cast<uint4*>(st + r * Sh::RS + ch * 8) = val;
  }
}
template <typename T, int DH>
__device__ __forceinline__ void store_tile(const T* st, void* dst, long long ld, const MhaArgs& a, long long gbase,
                                           int h, int s0, int lane) {
  using Sh = TcShape<DH>;
#pragma unroll
  for (int i = lane; i < 16 * Sh::CH; i += 32) {
    const int r = i / Sh::CH, ch = i % Sh::CH, s = s0 + r;
    if (s < a.S)
      *reinterpret_cast<uint4*>(static_cast<T*>(dst) + (gbase + (long long)s * a.seq_stride) * ld + h * DH + ch * 8) =
          *reinterpret_cast<const uint4*>(st + r * Sh::RS + ch * 8);
  }
}
// accumulator fragments of a 16 x DH tile -> staging tile (row lane / 4 (+ 8), columns 8 n + 2 (lane % 4))
template <typename T, int DH>
__device__ __forceinline__ void frags_to_tile(T* st, const float (&acc)[DH / 8][4], int lane) {
  using Sh = TcShape<DH>;
  const int g = lane >> 2, c = lane & 3;
#pragma unroll
  for (int n = 0; n < Sh::NT; ++n) {
    *reinterpret_cast<uint32_t*>(st + g * Sh::RS + n * 8 + 2 * c) = Mma<T>::pack(acc[n][0], acc[n][1]);
    *reinterpret_cast<uint32_t*>(st + (g + 8) * Sh::RS + n * 8 + 2 * c) = Mma<T>::pack(acc[n][2], acc[n][3]);
  }
}
// A fragments (16 rows x DH) of a staging tile
template <typename T, int DH>
__device__ __forceinline__ void tile_a_frags(uint32_t (&fa)[DH / 16][4], const T* st, int lane) {
  using Sh = TcShape<DH>;
  const int mat = lane >> 3, r = (lane & 7) + (mat & 1) * 8, cofs = (mat >> 1) * 8;
#pragma unroll
  for (int ks = 0; ks < Sh::KS; ++ks) ldsm4(fa[ks], smem_u32(st + r * Sh::RS + ks * 16 + cofs));
}
// acc(16 x 8) += A(16 x DH) . Y[n0 .. n0 + 7][:]^T   (Y staged row-major, contraction over the head dim)
template <typename T, int DH>
__device__ __forceinline__ void mma_rows(float (&acc)[4], const uint32_t (&fa)[DH / 16][4], const T* Y, int n0,
                                         int lane) {
  using Sh = TcShape<DH>;
#pragma unroll
  for (int kk = 0; kk < DH / 32; ++kk) {
    uint32_t b[4];
    ldsm4(b, smem_u32(Y + (n0 + (lane & 7)) * Sh::RS + kk * 32 + (lane >> 3) * 8));
    Mma<T>::run(acc, fa[2 * kk], b[0], b[1]);
    Mma<T>::run(acc, fa[2 * kk + 1], b[2], b[3]);
  }
}
// acc(16 x DH) += A(16 x 16) . Z[k0 .. k0 + 15][:]       (contraction over 16 staged rows)
template <typename T, int DH>
__device__ __forceinline__ void mma_cols(float (&acc)[DH / 8][4], const uint32_t (&pa)[4], const T* Z, int k0,
                                         int lane) {
  using Sh = TcShape<DH>;
#pragma unroll
  for (int dn = 0; dn < Sh::NT; dn += 2) {
    uint32_t b[4];
    ldsm4t(b, smem_u32(Z + (k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * Sh::RS + (dn + (lane >> 4)) * 8));
    Mma<T>::run(acc[dn], pa, b[0], b[1]);
    Mma<T>::run(acc[dn + 1], pa, b[2], b[3]);
  }
}

// work items of a CTA: CTA (x, y) covers problems [x * npc, (x + 1) * npc) and, of each, the 16-row tiles
// [y * tpc, (y + 1) * tpc); item it = (problem it / tpc, tile it % tpc)
struct WarpItem {
  int pl, p, tile;
  bool valid;
};
__device__ __forceinline__ WarpItem warp_item(const MhaArgs& a, int npc, int tpc, int ntiles, int it) {
  WarpItem w;
  w.pl = it / tpc;
  w.p = blockIdx.x * npc + w.pl;
  w.tile = blockIdx.y * tpc + it % tpc;
  w.valid = w.pl < npc && w.p < a.G * a.H && w.tile < ntiles;
  return w;
}

template <typename T, int DH, bool DROP>
__global__ void __launch_bounds__(kMhaThreads)
mha_fwd_tc_kernel(const MhaArgs a, int npc, int tpc, int ntiles, int S_pad) {
  using Sh = TcShape<DH>;
  extern __shared__ __align__(16) unsigned char mha_smem[];
  T* Ks = reinterpret_cast<T*>(mha_smem);
  T* Vs = Ks + (size_t)npc * S_pad * Sh::RS;
  T* stage = Vs + (size_t)npc * S_pad * Sh::RS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t rkey = 0u;
  if (DROP) {
    const unsigned long long seed = a.key[0], step = a.key[1];
    if (a.key_save != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
      a.key_save[0] = seed;
      a.key_save[1] = step;
    }
    rkey = mha_stream_key(seed, step, a.site);
  }
  stage_matrix<T, DH>(Ks, a.k, a.ldk, a, blockIdx.x * npc, npc, S_pad);
  stage_matrix<T, DH>(Vs, a.v, a.ldv, a, blockIdx.x * npc, npc, S_pad);
  stage_wait();
  // warps are independent from here on; a warp takes the work items warp, warp + 8, ... of the CTA
  for (int it = warp; it < npc * tpc; it += kMhaWarps) {
  const WarpItem w = warp_item(a, npc, tpc, ntiles, it);
  if (!w.valid) continue;

  const int g = w.p / a.H, h = w.p % a.H;
  const long long gbase = group_base(a, g);
  const int q0 = w.tile * 16;
  T* st = stage + warp * 16 * Sh::RS;
  load_tile<T, DH>(st, a.q, a.ldq, a, gbase, h, q0, lane);
  __syncwarp();
  uint32_t qa[Sh::KS][4];
  tile_a_frags<T, DH>(qa, st, lane);

  const T* Kp = Ks + (size_t)w.pl * S_pad * Sh::RS;
  const T* Vp = Vs + (size_t)w.pl * S_pad * Sh::RS;
  const int rg = lane >> 2, c = lane & 3;
  const float sl2 = a.scale * 1.4426950408889634f;
  float o[Sh::NT][4];
#pragma unroll
  for (int n = 0; n < Sh::NT; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};

  for (int k0 = 0; k0 < S_pad; k0 += 32) {
    float s[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      mma_rows<T, DH>(s[nt], qa, Kp, k0 + 8 * nt, lane);
    }
    // inadmissible keys of the block as bits: the attention mask's row words, plus every key beyond S
    uint32_t mw[2];
    mw[0] = mw[1] = (k0 + 32 > a.S) ? (0xffffffffu << (a.S - k0)) : 0u;
    if (a.mask_bits != nullptr) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int q = q0 + rg + 8 * r;
        if (q < a.S) mw[r] |= __ldg(a.mask_bits + (long long)q * a.mask_words + (k0 >> 5));
      }
    }
    if (__any_sync(0xffffffffu, (mw[0] | mw[1]) != 0u)) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const uint32_t bits = mw[r] >> (2 * c);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          if (bits & (1u << (8 * nt))) s[nt][2 * r] = -INFINITY;
          if (bits & (2u << (8 * nt))) s[nt][2 * r + 1] = -INFINITY;
        }
      }
    }
    float mx[2], mu[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(fmaxf(s[0][2 * r], s[0][2 * r + 1]), fmaxf(s[1][2 * r], s[1][2 * r + 1]));
      mx[r] = fmaxf(mx[r], fmaxf(fmaxf(s[2][2 * r], s[2][2 * r + 1]), fmaxf(s[3][2 * r], s[3][2 * r + 1])));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float mn = fmaxf(m[r], mx[r] * sl2);           // running maximum in units of log2
      mu[r] = mn == -INFINITY ? 0.f : mn;
      const float corr = ex2(m[r] - mu[r]);
      m[r] = mn;
      l[r] *= corr;
#pragma unroll
      for (int n = 0; n < Sh::NT; ++n) {
        o[n][2 * r] *= corr;
        o[n][2 * r + 1] *= corr;
      }
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float pv = ex2(fmaf(s[nt][i], sl2, -mu[i >> 1]));
        l[i >> 1] += pv;
        s[nt][i] = pv;
      }
    if (DROP) {
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const uint2 rnd = mha_rand4((uint32_t)w.p, q0 + rg, k0 + 16 * j + 2 * c + e, rkey);
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              float& x = s[2 * j + t][2 * r + e];
              x = pick_lane(rnd, r * 2 + t) >= a.thresh ? x * a.inv_keep : 0.f;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint32_t pa[4];
      pa[0] = Mma<T>::pack(s[2 * j][0], s[2 * j][1]);
      pa[1] = Mma<T>::pack(s[2 * j][2], s[2 * j][3]);
      pa[2] = Mma<T>::pack(s[2 * j + 1][0], s[2 * j + 1][1]);
      pa[3] = Mma<T>::pack(s[2 * j + 1][2], s[2 * j + 1][3]);
      mma_cols<T, DH>(o, pa, Vp, k0 + 16 * j, lane);
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
    const float inv = l[r] > 0.f ? 1.f / l[r] : 0.f;
#pragma unroll
    for (int n = 0; n < Sh::NT; ++n) {
      o[n][2 * r] *= inv;
      o[n][2 * r + 1] *= inv;
    }
    const int q = q0 + rg + 8 * r;
    if (c == 0 && q < a.S)
      a.lse[(long long)w.p * a.S + q] = l[r] > 0.f ? (m[r] + log2f(l[r])) * 0.6931471805599453f : INFINITY;
  }
  __syncwarp();
  frags_to_tile<T, DH>(st, o, lane);
  __syncwarp();
  store_tile<T, DH>(st, a.out, a.ldo, a, gbase, h, q0, lane);
  __syncwarp();
  }
}

// Backward.  TR = false (pass A): the warp's rows are 16 queries, X = Q tile, Xd = dO tile, the staged
// matrices are Y = K, Z = V; result dQ (and delta).  TR = true (pass B): the rows are 16 keys, X = K tile,
// Xd = V tile, Y = Q, Z = dO; the score tile is the transpose of pass A's, the statistics (lse, delta) go
// with the columns; results dK and dV.
template <typename T, int DH, bool TR, bool DROP>
__global__ void __launch_bounds__(kMhaThreads, DH == 32 ? 2 : 1)
mha_bwd_tc_kernel(const MhaArgs a, int npc, int tpc, int ntiles, int S_pad) {
  using Sh = TcShape<DH>;
  extern __shared__ __align__(16) unsigned char mha_smem[];
  T* Ys = reinterpret_cast<T*>(mha_smem);
  T* Zs = Ys + (size_t)npc * S_pad * Sh::RS;
  T* stage = Zs + (size_t)npc * S_pad * Sh::RS;                       // [warps][2][16][RS]
  float* stat = reinterpret_cast<float*>(stage + kMhaWarps * 2 * 16 * Sh::RS);
  // pass A: [warps][16] delta of the warp's rows; pass B: lse then delta of all staged tokens
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t rkey = 0u;
  if (DROP) rkey = mha_stream_key(a.key[0], a.key[1], a.site);
  const int p0 = blockIdx.x * npc;
  stage_matrix<T, DH>(Ys, TR ? a.q : a.k, TR ? a.ldq : a.ldk, a, p0, npc, S_pad);
  stage_matrix<T, DH>(Zs, TR ? a.dout : a.v, TR ? a.ldo : a.ldv, a, p0, npc, S_pad);
  if (TR) {
    const int P = a.G * a.H;
    for (int i = threadIdx.x; i < npc * S_pad; i += kMhaThreads) {
      const int s = i % S_pad, p = p0 + i / S_pad;
      const bool on = p < P && s < a.S;
      stat[i] = on ? a.lse[(long long)p * a.S + s] * 1.4426950408889634f : INFINITY;
      stat[npc * S_pad + i] = on ? a.delta[(long long)p * a.S + s] : 0.f;
    }
  }
  stage_wait();
  for (int it = warp; it < npc * tpc; it += kMhaWarps) {
  const WarpItem w = warp_item(a, npc, tpc, ntiles, it);
  if (!w.valid) continue;

  const int g = w.p / a.H, h = w.p % a.H;
  const long long gbase = group_base(a, g);
  const int r0 = w.tile * 16;
  const int rg = lane >> 2, c = lane & 3;
  T* st0 = stage + warp * 2 * 16 * Sh::RS;
  T* st1 = st0 + 16 * Sh::RS;
  float Lr[2] = {INFINITY, INFINITY}, Dr[2] = {0.f, 0.f};            // row statistics (pass A)
  if (!TR) {
    // delta of the warp's 16 queries from the O and dO tiles: two lanes per row
    load_tile<T, DH>(st0, a.o, a.ldo, a, gbase, h, r0, lane);
    load_tile<T, DH>(st1, a.dout, a.ldo, a, gbase, h, r0, lane);
    __syncwarp();
    const int row = lane >> 1, half = lane & 1;
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < DH / 2; ++i)
      d += to_f32(st0[row * Sh::RS + half * (DH / 2) + i]) * to_f32(st1[row * Sh::RS + half * (DH / 2) + i]);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    float* dl = stat + warp * 16;
    if (half == 0) {
      dl[row] = d;
      if (r0 + row < a.S) a.delta[(long long)w.p * a.S + r0 + row] = d;
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int q = r0 + rg + 8 * r;
      Dr[r] = dl[rg + 8 * r];
      if (q < a.S) Lr[r] = a.lse[(long long)w.p * a.S + q] * 1.4426950408889634f;
    }
    load_tile<T, DH>(st0, a.q, a.ldq, a, gbase, h, r0, lane);       // dO stays in st1
  } else {
    load_tile<T, DH>(st0, a.k, a.ldk, a, gbase, h, r0, lane);
    load_tile<T, DH>(st1, a.v, a.ldv, a, gbase, h, r0, lane);
  }
  __syncwarp();
  uint32_t xa[Sh::KS][4], xda[Sh::KS][4];
  tile_a_frags<T, DH>(xa, st0, lane);
  tile_a_frags<T, DH>(xda, st1, lane);

  const T* Yp = Ys + (size_t)w.pl * S_pad * Sh::RS;
  const T* Zp = Zs + (size_t)w.pl * S_pad * Sh::RS;
  const float* Lc = stat + (size_t)w.pl * S_pad;                      // column statistics (pass B)
  const float* Dc = stat + (size_t)npc * S_pad + (size_t)w.pl * S_pad;
  const uint32_t* mbits = TR ? a.mask_bits_t : a.mask_bits;
  const float sl2 = a.scale * 1.4426950408889634f;
  float acc1[Sh::NT][4], acc2[Sh::NT][4];                             // dQ | dK, dV
#pragma unroll
  for (int n = 0; n < Sh::NT; ++n) {
    acc1[n][0] = acc1[n][1] = acc1[n][2] = acc1[n][3] = 0.f;
    acc2[n][0] = acc2[n][1] = acc2[n][2] = acc2[n][3] = 0.f;
  }

  for (int c0 = 0; c0 < S_pad; c0 += 32) {
    float s[4][4], dp[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
      mma_rows<T, DH>(s[nt], xa, Yp, c0 + 8 * nt, lane);
      mma_rows<T, DH>(dp[nt], xda, Zp, c0 + 8 * nt, lane);
    }
    uint32_t mw[2];
    mw[0] = mw[1] = (!TR && c0 + 32 > a.S) ? (0xffffffffu << (a.S - c0)) : 0u;   // keys beyond S (pass A)
    if (mbits != nullptr) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int row = r0 + rg + 8 * r;
        if (row < a.S) mw[r] |= __ldg(mbits + (long long)row * a.mask_words + (c0 >> 5));
      }
    }
    uint2 rnd[2][2];
    if (DROP) {
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int row = r0 + rg, col = c0 + 16 * j + 2 * c + e;
          rnd[j][e] = TR ? mha_rand4((uint32_t)w.p, col, row, rkey) : mha_rand4((uint32_t)w.p, row, col, rkey);
        }
    }
    float pd[4][4];                                                   // dropped-out probabilities (pass B)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = i >> 1, e = i & 1, kc = 8 * nt + 2 * c + e;
        float L, D;
        if (TR) {
          L = Lc[c0 + kc];
          D = Dc[c0 + kc];
        } else {
          L = Lr[r];
          D = Dr[r];
        }
        const bool masked = (mw[r] >> kc) & 1u;
        const float pv = masked ? 0.f : ex2(fmaf(s[nt][i], sl2, -L));
        float dpm = dp[nt][i], pk = pv;
        if (DROP) {
          // pass A: word (row half r, column half t); pass B: the element is (q = column, k = row)
          const int t = nt & 1;
          const bool keep = pick_lane(rnd[nt >> 1][e], TR ? t * 2 + r : r * 2 + t) >= a.thresh;
          dpm = keep ? dpm * a.inv_keep : 0.f;
          pk = keep ? pv * a.inv_keep : 0.f;
        }
        s[nt][i] = pv * (dpm - D) * a.scale;                          // dS
        pd[nt][i] = pk;
      }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint32_t da[4];
      da[0] = Mma<T>::pack(s[2 * j][0], s[2 * j][1]);
      da[1] = Mma<T>::pack(s[2 * j][2], s[2 * j][3]);
      da[2] = Mma<T>::pack(s[2 * j + 1][0], s[2 * j + 1][1]);
      da[3] = Mma<T>::pack(s[2 * j + 1][2], s[2 * j + 1][3]);
      mma_cols<T, DH>(acc1, da, Yp, c0 + 16 * j, lane);               // dQ += dS K   |   dK += dS^T Q
      if (TR) {
        uint32_t pa[4];
        pa[0] = Mma<T>::pack(pd[2 * j][0], pd[2 * j][1]);
        pa[1] = Mma<T>::pack(pd[2 * j][2], pd[2 * j][3]);
        pa[2] = Mma<T>::pack(pd[2 * j + 1][0], pd[2 * j + 1][1]);
        pa[3] = Mma<T>::pack(pd[2 * j + 1][2], pd[2 * j + 1][3]);
        mma_cols<T, DH>(acc2, pa, Zp, c0 + 16 * j, lane);             // dV += P_drop^T dO
      }
    }
  }
  __syncwarp();
  frags_to_tile<T, DH>(st0, acc1, lane);
  if (TR) frags_to_tile<T, DH>(st1, acc2, lane);
  __syncwarp();
  if (!TR) {
    store_tile<T, DH>(st0, a.dq, a.lddq, a, gbase, h, r0, lane);
  } else {
    store_tile<T, DH>(st0, a.dk, a.lddk, a, gbase, h, r0, lane);
    store_tile<T, DH>(st1, a.dv, a.lddv, a, gbase, h, r0, lane);
  }
  __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------
// one-warp-per-row path (fp32 models, shapes outside the tensor-core path)
// ---------------------------------------------------------------------------------------------------
template <typename T, int DH>
__device__ __forceinline__ void load_row(float (&f)[DH], const void* base, long long row, long long ld, int h) {
  constexpr int VEC = Vec16<T>::N;
  const T* p = static_cast<const T*>(base) + row * ld + h * DH;
#pragma unroll
  for (int i = 0; i < DH / VEC; ++i) {
    float t[VEC];
    Vec16<T>::unpack(ldg128(p + i * VEC), t);
#pragma unroll
    for (int j = 0; j < VEC; ++j) f[i * VEC + j] = t[j];
  }
}
template <int DH> __device__ __forceinline__ float dot_row(const float (&x)[DH], const float (&y)[DH]) {
  float s[4] = {0.f, 0.f, 0.f, 0.f};             // four independent chains instead of DH dependent FMAs
#pragma unroll
  for (int i = 0; i < DH; ++i) s[i & 3] = fmaf(x[i], y[i], s[i & 3]);
  return (s[0] + s[1]) + (s[2] + s[3]);
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// out[d] = sum_t w[t] * M[token t][d] for the lane's dims d = lane + 32 j
template <typename T, int DH>
__device__ __forceinline__ void weighted_rows(float (&acc)[(DH + 31) / 32], const float* w, const void* base,
                                              long long gbase, long long seq_stride, long long ld, int h, int S,
                                              int lane) {
  constexpr int NJ = (DH + 31) / 32;
  float part[4][NJ];                             // four tokens per iteration, four independent chains
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int j = 0; j < NJ; ++j) part[u][j] = 0.f;
  const T* p = static_cast<const T*>(base) + gbase * ld + h * DH;
  const long long step = seq_stride * ld;
  int t = 0;
  for (; t + 4 <= S; t += 4) {
    const float4 wt = *reinterpret_cast<const float4*>(w + t);       // w is 16-byte aligned, t a multiple of 4
    const T* row = p + (long long)t * step;
    const float wv[4] = {wt.x, wt.y, wt.z, wt.w};
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int d = lane + 32 * j;
        if (d < DH) part[u][j] = fmaf(wv[u], to_f32(row[u * step + d]), part[u][j]);
      }
  }
  for (; t < S; ++t) {
    const float wt = w[t];
    const T* row = p + (long long)t * step;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int d = lane + 32 * j;
      if (d < DH) part[0][j] = fmaf(wt, to_f32(row[d]), part[0][j]);
    }
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) acc[j] = (part[0][j] + part[1][j]) + (part[2][j] + part[3][j]);
}
template <typename T, int DH>
__device__ __forceinline__ void store_dims(void* base, long long row, long long ld, int h,
                                           const float (&acc)[(DH + 31) / 32], int lane) {
  T* p = static_cast<T*>(base) + row * ld + h * DH;
#pragma unroll
  for (int j = 0; j < (DH + 31) / 32; ++j) {
    const int d = lane + 32 * j;
    if (d < DH) p[d] = from_f32<T>(acc[j]);
  }
}

// 32 token rows [t0, t0 + 32) of one (group, head) -> a shared tile [32][DH + 1] fp32 (coalesced loads; the
// padded row stride makes "lane = token" reads conflict-free); rows beyond S are zero
template <typename T, int DH>
__device__ __forceinline__ void stage_rows(float* dst, const void* base, long long ld, long long gbase,
                                           long long seq_stride, int h, int t0, int S) {
  const T* p = static_cast<const T*>(base) + gbase * ld + h * DH;
  for (int i = threadIdx.x; i < 32 * DH; i += kMhaThreads) {
    const int row = i / DH, d = i % DH, t = t0 + row;
    dst[row * (DH + 1) + d] = t < S ? to_f32(p[(long long)t * seq_stride * ld + d]) : 0.f;
  }
}
template <int DH> __device__ __forceinline__ float dot_tile(const float (&x)[DH], const float* row) {
  float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < DH; ++i) s[i & 3] = fmaf(x[i], row[i], s[i & 3]);
  return (s[0] + s[1]) + (s[2] + s[3]);
}

// A CTA owns 8 consecutive rows of one problem (one warp per row); the other side of the score matrix
// streams through shared memory in tiles of 32 tokens shared by the 8 warps.
struct SimtItem {
  int p, r, g, h;
  bool valid;
  long long gbase;
};
__device__ __forceinline__ SimtItem simt_item(const MhaArgs& a) {
  const int nblk = (a.S + kMhaWarps - 1) / kMhaWarps;
  SimtItem it;
  it.p = blockIdx.x / nblk;
  it.r = (blockIdx.x % nblk) * kMhaWarps + (threadIdx.x >> 5);
  it.valid = it.r < a.S;
  it.g = it.p / a.H;
  it.h = it.p % a.H;
  it.gbase = group_base(a, it.g);
  return it;
}

template <typename T, int DH, bool DROP>
__global__ void __launch_bounds__(kMhaThreads) mha_fwd_simt_kernel(const MhaArgs a, int S_round) {
  extern __shared__ __align__(16) unsigned char mha_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t rkey = 0u;
  if (DROP) {
    const unsigned long long seed = a.key[0], step = a.key[1];
    if (a.key_save != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
      a.key_save[0] = seed;
      a.key_save[1] = step;
    }
    rkey = mha_stream_key(seed, step, a.site);
  }
  const SimtItem it = simt_item(a);
  const int p = it.p, q = it.r, h = it.h;
  float* tile = reinterpret_cast<float*>(mha_smem);
  float* sc = tile + 32 * (DH + 1) + (size_t)warp * S_round;
  float qr[DH];
#pragma unroll
  for (int i = 0; i < DH; ++i) qr[i] = 0.f;
  if (it.valid) load_row<T, DH>(qr, a.q, it.gbase + (long long)q * a.seq_stride, a.ldq, h);
  const uint32_t* mrow = (a.mask_bits && it.valid) ? a.mask_bits + (long long)q * a.mask_words : nullptr;
  float mx = -INFINITY;
  for (int t0 = 0; t0 < a.S; t0 += 32) {
    __syncthreads();
    stage_rows<T, DH>(tile, a.k, a.ldk, it.gbase, a.seq_stride, h, t0, a.S);
    __syncthreads();
    const int k = t0 + lane;
    if (it.valid && k < a.S) {
      float s = dot_tile<DH>(qr, tile + lane * (DH + 1)) * a.scale;
      if (mrow != nullptr && ((__ldg(mrow + (k >> 5)) >> (k & 31)) & 1u)) s = -INFINITY;
      sc[k] = s;
      mx = fmaxf(mx, s);
    }
  }
  if (!it.valid) return;                       // no block-wide barrier below
  mx = warp_max_f(mx);
  const float mu = mx == -INFINITY ? 0.f : mx;
  float l = 0.f;
  for (int k = lane; k < a.S; k += 32) {
    const float pv = expf(sc[k] - mu);
    sc[k] = pv;
    l += pv;
  }
  l = warp_sum_f(l);
  const float inv = l > 0.f ? 1.f / l : 0.f;
  if (lane == 0) a.lse[(long long)p * a.S + q] = l > 0.f ? mu + logf(l) : INFINITY;
  for (int k = lane; k < a.S; k += 32) {
    float pv = sc[k] * inv;
    if (DROP) pv = mha_keep((uint32_t)p, q, k, rkey, a.thresh) ? pv * a.inv_keep : 0.f;
    sc[k] = pv;
  }
  __syncwarp();
  float acc[(DH + 31) / 32];
  weighted_rows<T, DH>(acc, sc, a.v, it.gbase, a.seq_stride, a.ldv, h, a.S, lane);
  store_dims<T, DH>(a.out, it.gbase + (long long)q * a.seq_stride, a.ldo, h, acc, lane);
}

// pass A: one warp per query -> dQ row and delta
template <typename T, int DH, bool DROP>
__global__ void __launch_bounds__(kMhaThreads) mha_bwd_dq_simt_kernel(const MhaArgs a, int S_round) {
  extern __shared__ __align__(16) unsigned char mha_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t rkey = 0u;
  if (DROP) rkey = mha_stream_key(a.key[0], a.key[1], a.site);
  const SimtItem it = simt_item(a);
  const int p = it.p, q = it.r, h = it.h;
  const long long qrow = it.gbase + (long long)q * a.seq_stride;
  float* tk = reinterpret_cast<float*>(mha_smem);
  float* tv = tk + 32 * (DH + 1);
  float* sc = tv + 32 * (DH + 1) + (size_t)warp * S_round;
  float qr[DH], dor[DH];
#pragma unroll
  for (int i = 0; i < DH; ++i) qr[i] = dor[i] = 0.f;
  float delta = 0.f, lse = INFINITY;
  if (it.valid) {
    load_row<T, DH>(qr, a.q, qrow, a.ldq, h);
    load_row<T, DH>(dor, a.dout, qrow, a.ldo, h);
    float orow[DH];
    load_row<T, DH>(orow, a.o, qrow, a.ldo, h);
    delta = dot_row<DH>(dor, orow);
    lse = a.lse[(long long)p * a.S + q];
    if (lane == 0) a.delta[(long long)p * a.S + q] = delta;
  }
  const uint32_t* mrow = (a.mask_bits && it.valid) ? a.mask_bits + (long long)q * a.mask_words : nullptr;
  for (int t0 = 0; t0 < a.S; t0 += 32) {
    __syncthreads();
    stage_rows<T, DH>(tk, a.k, a.ldk, it.gbase, a.seq_stride, h, t0, a.S);
    stage_rows<T, DH>(tv, a.v, a.ldv, it.gbase, a.seq_stride, h, t0, a.S);
    __syncthreads();
    const int k = t0 + lane;
    if (it.valid && k < a.S) {
      const float s = dot_tile<DH>(qr, tk + lane * (DH + 1)) * a.scale;
      const bool masked = mrow != nullptr && ((__ldg(mrow + (k >> 5)) >> (k & 31)) & 1u);
      const float pv = masked ? 0.f : expf(s - lse);
      float dp = dot_tile<DH>(dor, tv + lane * (DH + 1));
      if (DROP) dp = mha_keep((uint32_t)p, q, k, rkey, a.thresh) ? dp * a.inv_keep : 0.f;
      sc[k] = pv * (dp - delta) * a.scale;
    }
  }
  if (!it.valid) return;
  __syncwarp();
  float acc[(DH + 31) / 32];
  weighted_rows<T, DH>(acc, sc, a.k, it.gbase, a.seq_stride, a.ldk, h, a.S, lane);
  store_dims<T, DH>(a.dq, qrow, a.lddq, h, acc, lane);
}

// pass B: one warp per key -> dK and dV rows
template <typename T, int DH, bool DROP>
__global__ void __launch_bounds__(kMhaThreads) mha_bwd_dkv_simt_kernel(const MhaArgs a, int S_round) {
  extern __shared__ __align__(16) unsigned char mha_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t rkey = 0u;
  if (DROP) rkey = mha_stream_key(a.key[0], a.key[1], a.site);
  const SimtItem it = simt_item(a);
  const int p = it.p, k = it.r, h = it.h;
  const long long krow = it.gbase + (long long)k * a.seq_stride;
  float* tq = reinterpret_cast<float*>(mha_smem);
  float* tdo = tq + 32 * (DH + 1);
  float* scp = tdo + 32 * (DH + 1) + (size_t)warp * 2 * S_round;
  float* scs = scp + S_round;
  float kr[DH], vr[DH];
#pragma unroll
  for (int i = 0; i < DH; ++i) kr[i] = vr[i] = 0.f;
  if (it.valid) {
    load_row<T, DH>(kr, a.k, krow, a.ldk, h);
    load_row<T, DH>(vr, a.v, krow, a.ldv, h);
  }
  for (int t0 = 0; t0 < a.S; t0 += 32) {
    __syncthreads();
    stage_rows<T, DH>(tq, a.q, a.ldq, it.gbase, a.seq_stride, h, t0, a.S);
    stage_rows<T, DH>(tdo, a.dout, a.ldo, it.gbase, a.seq_stride, h, t0, a.S);
    __syncthreads();
    const int q = t0 + lane;
    if (it.valid && q < a.S) {
      const float s = dot_tile<DH>(kr, tq + lane * (DH + 1)) * a.scale;
      const bool masked = a.mask_bits != nullptr &&
                          ((__ldg(a.mask_bits + (long long)q * a.mask_words + (k >> 5)) >> (k & 31)) & 1u);
      const float pv = masked ? 0.f : expf(s - a.lse[(long long)p * a.S + q]);
      float dp = dot_tile<DH>(vr, tdo + lane * (DH + 1)), pk = pv;
      if (DROP) {
        const bool keep = mha_keep((uint32_t)p, q, k, rkey, a.thresh);
        dp = keep ? dp * a.inv_keep : 0.f;
        pk = keep ? pv * a.inv_keep : 0.f;
      }
      scp[q] = pk;
      scs[q] = pv * (dp - a.delta[(long long)p * a.S + q]) * a.scale;
    }
  }
  if (!it.valid) return;
  __syncwarp();
  float acc[(DH + 31) / 32];
  weighted_rows<T, DH>(acc, scp, a.dout, it.gbase, a.seq_stride, a.ldo, h, a.S, lane);
  store_dims<T, DH>(a.dv, krow, a.lddv, h, acc, lane);
  weighted_rows<T, DH>(acc, scs, a.q, it.gbase, a.seq_stride, a.ldq, h, a.S, lane);
  store_dims<T, DH>(a.dk, krow, a.lddk, h, acc, lane);
}

__global__ void mha_pack_mask_kernel(const uint8_t* __restrict__ mask, int S, int W, uint32_t* __restrict__ bits,
                                     uint32_t* __restrict__ bits_t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * W) return;
  const int row = i / W, w = i % W;
  uint32_t b = 0u, bt = 0u;
  for (int j = 0; j < 32; ++j) {
    const int col = 32 * w + j;
    if (col < S) {
      b |= (mask[(long long)row * S + col] ? 1u : 0u) << j;
      bt |= (mask[(long long)col * S + row] ? 1u : 0u) << j;
    }
  }
  bits[i] = b;
  bits_t[i] = bt;
}

__global__ void mha_keep_mask_kernel(uint8_t* __restrict__ keep, int P, int S, const unsigned long long* key,
                                     uint32_t site, uint32_t thresh) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)P * S * S) return;
  const int k = (int)(i % S), q = (int)((i / S) % S), p = (int)(i / ((long long)S * S));
  keep[i] = mha_keep((uint32_t)p, q, k, mha_stream_key(key[0], key[1], site), thresh) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
struct TcPlan {
  int S_pad, ntiles, npc, tpc;
  dim3 grid;
  size_t smem_fwd, smem_bwd_a, smem_bwd_b;
};
TcPlan tc_plan(int P, int S, int Dh) {
  TcPlan t;
  t.S_pad = (S + 31) / 32 * 32;
  t.ntiles = (S + 15) / 16;
  if (t.ntiles >= kMhaWarps) {
    t.npc = 1;
    // one tile per warp; few long problems (the detection decoder: 8 problems x 57 tiles) take half-filled
    // CTAs rather than leave SMs idle (measured: 80 -> 70 us forward + backward at 900 queries)
    t.tpc = ((long long)P * ((t.ntiles + kMhaWarps - 1) / kMhaWarps) < 148) ? kMhaWarps / 2 : kMhaWarps;
  } else {
    t.tpc = t.ntiles;
    t.npc = kMhaWarps / t.ntiles;
  }
  t.grid = dim3((unsigned)((P + t.npc - 1) / t.npc), (unsigned)((t.ntiles + t.tpc - 1) / t.tpc));
  const size_t row = (size_t)(Dh + 8) * 2;
  const size_t mats = 2 * (size_t)t.npc * t.S_pad * row;
  t.smem_fwd = mats + kMhaWarps * 16 * row;
  t.smem_bwd_a = mats + kMhaWarps * 2 * 16 * row + kMhaWarps * 16 * sizeof(float);
  t.smem_bwd_b = mats + kMhaWarps * 2 * 16 * row + 2 * (size_t)t.npc * t.S_pad * sizeof(float);
  return t;
}

template <typename K> int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024 &&
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
    cudaGetLastError();
    return set_error(MSDA_ERR_CUDA, "mha: cannot reserve %zu bytes of shared memory", bytes);
  }
  return MSDA_OK;
}

template <typename T, int DH, bool DROP> int fwd_tc(const MhaArgs& a, const TcPlan& t, cudaStream_t st) {
  if (int rc = set_smem(mha_fwd_tc_kernel<T, DH, DROP>, t.smem_fwd)) return rc;
  mha_fwd_tc_kernel<T, DH, DROP><<<t.grid, kMhaThreads, t.smem_fwd, st>>>(a, t.npc, t.tpc, t.ntiles, t.S_pad);
  count_launch();
  return check_launch("mha_fwd");
}
template <typename T, int DH, bool DROP> int bwd_tc(const MhaArgs& a, const TcPlan& t, cudaStream_t st) {
  if (int rc = set_smem(mha_bwd_tc_kernel<T, DH, false, DROP>, t.smem_bwd_a)) return rc;
  if (int rc = set_smem(mha_bwd_tc_kernel<T, DH, true, DROP>, t.smem_bwd_b)) return rc;
  mha_bwd_tc_kernel<T, DH, false, DROP><<<t.grid, kMhaThreads, t.smem_bwd_a, st>>>(a, t.npc, t.tpc, t.ntiles, t.S_pad);
  count_launch();
  if (int rc = check_launch("mha_bwd (dQ pass)")) return rc;
  mha_bwd_tc_kernel<T, DH, true, DROP><<<t.grid, kMhaThreads, t.smem_bwd_b, st>>>(a, t.npc, t.tpc, t.ntiles, t.S_pad);
  count_launch();
  return check_launch("mha_bwd (dK / dV pass)");
}

template <typename T, int DH, bool DROP> int fwd_simt(const MhaArgs& a, cudaStream_t st) {
  const int S_round = (a.S + 31) / 32 * 32;
  const size_t smem = ((size_t)kMhaWarps * S_round + 32 * (DH + 1)) * sizeof(float);
  const long long grid = (long long)a.G * a.H * ((a.S + kMhaWarps - 1) / kMhaWarps);
  if (grid > 0x7fffffffLL) return set_error(MSDA_ERR_UNSUPPORTED, "mha_fwd: too many rows for one launch");
  if (int rc = set_smem(mha_fwd_simt_kernel<T, DH, DROP>, smem)) return rc;
  mha_fwd_simt_kernel<T, DH, DROP><<<(unsigned)grid, kMhaThreads, smem, st>>>(a, S_round);
  count_launch();
  return check_launch("mha_fwd");
}
template <typename T, int DH, bool DROP> int bwd_simt(const MhaArgs& a, cudaStream_t st) {
  const int S_round = (a.S + 31) / 32 * 32;
  const size_t tiles = 2 * 32 * (DH + 1) * sizeof(float);
  const size_t smem_a = (size_t)kMhaWarps * S_round * sizeof(float) + tiles;
  const size_t smem_b = (size_t)kMhaWarps * 2 * S_round * sizeof(float) + tiles;
  const long long blocks = (long long)a.G * a.H * ((a.S + kMhaWarps - 1) / kMhaWarps);
  if (blocks > 0x7fffffffLL) return set_error(MSDA_ERR_UNSUPPORTED, "mha_bwd: too many rows for one launch");
  const unsigned grid = (unsigned)blocks;
  if (int rc = set_smem(mha_bwd_dq_simt_kernel<T, DH, DROP>, smem_a)) return rc;
  if (int rc = set_smem(mha_bwd_dkv_simt_kernel<T, DH, DROP>, smem_b)) return rc;
  mha_bwd_dq_simt_kernel<T, DH, DROP><<<grid, kMhaThreads, smem_a, st>>>(a, S_round);
  count_launch();
  if (int rc = check_launch("mha_bwd (dQ pass)")) return rc;
  mha_bwd_dkv_simt_kernel<T, DH, DROP><<<grid, kMhaThreads, smem_b, st>>>(a, S_round);
  count_launch();
  return check_launch("mha_bwd (dK / dV pass)");
}

constexpr int kSimtMaxS = 3072;     // 8 warps x 2 score rows x 3072 floats = 192 KB (+ two 32-token tiles)

bool tc_covers(int dtype, int Dh, int P, int S) {
  if (dtype == MSDA_F32 || (Dh != 32 && Dh != 64)) return false;
  const TcPlan t = tc_plan(P, S, Dh);
  return t.smem_bwd_b <= (size_t)kMhaMaxSmem && t.smem_bwd_a <= (size_t)kMhaMaxSmem && t.grid.y <= 65535u;
}

template <typename T, bool BWD, bool DROP> int dispatch_simt(const MhaArgs& a, int Dh, cudaStream_t st) {
#define MHA_SIMT_CASE(D) case D: return BWD ? bwd_simt<T, D, DROP>(a, st) : fwd_simt<T, D, DROP>(a, st);
  switch (Dh) {
    MHA_SIMT_CASE(8)
    MHA_SIMT_CASE(16)
    MHA_SIMT_CASE(32)
    MHA_SIMT_CASE(64)
    default: return set_error(MSDA_ERR_UNSUPPORTED, "mha: head_dim %d is not one of 8, 16, 32, 64", Dh);
  }
#undef MHA_SIMT_CASE
}
template <typename T, bool BWD, bool DROP> int dispatch_tc(const MhaArgs& a, int Dh, cudaStream_t st) {
  const TcPlan t = tc_plan(a.G * a.H, a.S, Dh);
  if (Dh == 32) return BWD ? bwd_tc<T, 32, DROP>(a, t, st) : fwd_tc<T, 32, DROP>(a, t, st);
  return BWD ? bwd_tc<T, 64, DROP>(a, t, st) : fwd_tc<T, 64, DROP>(a, t, st);
}

template <bool BWD> int mha_dispatch(const MhaArgs& a, int Dh, int dtype, int impl, cudaStream_t st) {
  const bool drop = a.key != nullptr;
  bool tc = tc_covers(dtype, Dh, a.G * a.H, a.S);
  if (impl == 2 && !tc)
    return set_error(MSDA_ERR_UNSUPPORTED, "mha: the tensor-core path needs a 16-bit dtype, head_dim 32 or 64 and "
                     "K / V of one problem in shared memory (S = %d, head_dim = %d, dtype %d)", a.S, Dh, dtype);
  if (impl == 1) tc = false;
  if (!tc && a.S > kSimtMaxS)
    return set_error(MSDA_ERR_UNSUPPORTED, "mha: sequence length %d exceeds %d", a.S, kSimtMaxS);
  if (tc) {
    if (dtype == MSDA_BF16)
      return drop ? dispatch_tc<__nv_bfloat16, BWD, true>(a, Dh, st) : dispatch_tc<__nv_bfloat16, BWD, false>(a, Dh, st);
    return drop ? dispatch_tc<__half, BWD, true>(a, Dh, st) : dispatch_tc<__half, BWD, false>(a, Dh, st);
  }
  if (dtype == MSDA_F32)
    return drop ? dispatch_simt<float, BWD, true>(a, Dh, st) : dispatch_simt<float, BWD, false>(a, Dh, st);
  if (dtype == MSDA_BF16)
    return drop ? dispatch_simt<__nv_bfloat16, BWD, true>(a, Dh, st) : dispatch_simt<__nv_bfloat16, BWD, false>(a, Dh, st);
  return drop ? dispatch_simt<__half, BWD, true>(a, Dh, st) : dispatch_simt<__half, BWD, false>(a, Dh, st);
}

int check_common(const char* what, const MhaArgs& a, int Dh, int dtype, float p) {
  if (a.G <= 0 || a.H <= 0 || a.S <= 0 || Dh <= 0 || a.n_lo <= 0)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: sizes must be positive (G=%d H=%d S=%d Dh=%d n_lo=%d)", what, a.G,
                     a.H, a.S, Dh, a.n_lo);
  if (dtype != MSDA_F32 && dtype != MSDA_F16 && dtype != MSDA_BF16)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: unknown dtype %d", what, dtype);
  if (!a.q || !a.k || !a.v || !a.lse) return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: NULL pointer", what);
  const int vec = dtype == MSDA_F32 ? 4 : 8;
  if (Dh % vec != 0 || a.ldq % vec != 0 || a.ldk % vec != 0 || a.ldv % vec != 0 || a.ldo % vec != 0 ||
      a.lddq % vec != 0 || a.lddk % vec != 0 || a.lddv % vec != 0)
    return set_error(MSDA_ERR_UNSUPPORTED, "%s: head_dim and the row strides must be multiples of %d elements", what, vec);
  const uintptr_t bits = (uintptr_t)a.q | (uintptr_t)a.k | (uintptr_t)a.v | (uintptr_t)a.out | (uintptr_t)a.o |
                         (uintptr_t)a.dout | (uintptr_t)a.dq | (uintptr_t)a.dk | (uintptr_t)a.dv;
  if (bits & 15u) return set_error(MSDA_ERR_UNSUPPORTED, "%s: pointers must be 16-byte aligned", what);
  if ((long long)a.G * a.H > 0x7fffffffLL / a.S || (long long)a.G * a.H >= (1LL << 31))
    return set_error(MSDA_ERR_UNSUPPORTED, "%s: too many (group, head, token) rows", what);
  if (a.mask_bits != nullptr && a.mask_words != (a.S + 31) / 32)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: mask_words must be ceil(S / 32)", what);
  if (!(p >= 0.f && p < 1.f)) return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: dropout p = %f outside [0, 1)", what, (double)p);
  return MSDA_OK;
}

void set_dropout(MhaArgs& a, const void* key, void* key_save, uint32_t site, float p) {
  a.key = (p > 0.f) ? static_cast<const unsigned long long*>(key) : nullptr;
  a.key_save = static_cast<unsigned long long*>(key_save);
  a.site = site;
  a.thresh = (uint32_t)((double)p * 65536.0 + 0.5);
  a.inv_keep = 1.f / (1.f - p);
}

}  // namespace
}  // namespace msda

using namespace msda;

extern "C" {

int mha_pack_mask(const uint8_t* mask, int S, uint32_t* bits, uint32_t* bits_t, void* stream) {
  if (!mask || !bits || !bits_t || S <= 0) return set_error(MSDA_ERR_BAD_ARGUMENT, "mha_pack_mask: bad argument");
  const int W = (S + 31) / 32, n = S * W;
  mha_pack_mask_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(mask, S, W, bits, bits_t);
  count_launch();
  return check_launch("mha_pack_mask");
}

int mha_impl(int G, int H, int S, int Dh, int dtype) {
  if (G <= 0 || H <= 0 || S <= 0) return 0;
  if (tc_covers(dtype, Dh, G * H, S)) return 2;
  const int vec = dtype == MSDA_F32 ? 4 : 8;
  return (S <= kSimtMaxS && (Dh == 8 || Dh == 16 || Dh == 32 || Dh == 64) && Dh % vec == 0) ? 1 : 0;
}

int mha_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int64_t ldq, int64_t ldk,
            int64_t ldv, int64_t ldo, const uint32_t* mask_bits, int G, int H, int S, int Dh, int64_t seq_stride,
            int64_t hi_stride, int64_t lo_stride, int n_lo, float scale, int dtype, int impl,
            const void* rng_state, void* key_save, uint32_t site, float p, void* stream) {
  MhaArgs a{};
  a.q = q; a.k = k; a.v = v; a.out = out; a.lse = lse;
  a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo;
  a.mask_bits = mask_bits; a.mask_words = (S + 31) / 32;
  a.G = G; a.H = H; a.S = S;
  a.seq_stride = seq_stride; a.hi_stride = hi_stride; a.lo_stride = lo_stride; a.n_lo = n_lo;
  a.scale = scale;
  if (int rc = check_common("mha_fwd", a, Dh, dtype, p)) return rc;
  if (!out) return set_error(MSDA_ERR_BAD_ARGUMENT, "mha_fwd: NULL output");
  if (p > 0.f && !rng_state) return set_error(MSDA_ERR_BAD_ARGUMENT, "mha_fwd: dropout needs rng_state");
  set_dropout(a, rng_state, key_save, site, p);
  return mha_dispatch<false>(a, Dh, dtype, impl, static_cast<cudaStream_t>(stream));
}

int mha_bwd(const void* q, const void* k, const void* v, const void* out, const void* grad_out, const float* lse,
            float* delta, void* dq, void* dk, void* dv, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
            int64_t ld_dq, int64_t ld_dk, int64_t ld_dv,
            const uint32_t* mask_bits, const uint32_t* mask_bits_t, int G, int H, int S, int Dh,
            int64_t seq_stride, int64_t hi_stride, int64_t lo_stride, int n_lo, float scale, int dtype, int impl,
            const void* key, uint32_t site, float p, void* stream) {
  MhaArgs a{};
  a.q = q; a.k = k; a.v = v; a.o = out; a.dout = grad_out; a.lse = const_cast<float*>(lse); a.delta = delta;
  a.dq = dq; a.dk = dk; a.dv = dv;
  a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo;
  a.lddq = ld_dq > 0 ? ld_dq : ldq; a.lddk = ld_dk > 0 ? ld_dk : ldk; a.lddv = ld_dv > 0 ? ld_dv : ldv;
  a.mask_bits = mask_bits; a.mask_bits_t = mask_bits_t; a.mask_words = (S + 31) / 32;
  a.G = G; a.H = H; a.S = S;
  a.seq_stride = seq_stride; a.hi_stride = hi_stride; a.lo_stride = lo_stride; a.n_lo = n_lo;
  a.scale = scale;
  if (int rc = check_common("mha_bwd", a, Dh, dtype, p)) return rc;
  if (!out || !grad_out || !delta || !dq || !dk || !dv) return set_error(MSDA_ERR_BAD_ARGUMENT, "mha_bwd: NULL pointer");
  if ((mask_bits == nullptr) != (mask_bits_t == nullptr))
    return set_error(MSDA_ERR_BAD_ARGUMENT, "mha_bwd: pass both packed masks or neither");
  if (p > 0.f && !key) return set_error(MSDA_ERR_BAD_ARGUMENT, "mha_bwd: dropout needs the key the forward saved");
  set_dropout(a, key, nullptr, site, p);
  return mha_dispatch<true>(a, Dh, dtype, impl, static_cast<cudaStream_t>(stream));
}

int mha_keep_mask(uint8_t* keep, int P, int S, const void* key, uint32_t site, float p, void* stream) {
  if (!keep || !key || P <= 0 || S <= 0 || !(p >= 0.f && p < 1.f))
    return set_error(MSDA_ERR_BAD_ARGUMENT, "mha_keep_mask: bad argument");
  MhaArgs a{};
  set_dropout(a, key, nullptr, site, p);
  const long long n = (long long)P * S * S;
  mha_keep_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      keep, P, S, static_cast<const unsigned long long*>(key), site, p > 0.f ? a.thresh : 0u);
  count_launch();
  return check_launch("mha_keep_mask");
}

}  // extern "C"
