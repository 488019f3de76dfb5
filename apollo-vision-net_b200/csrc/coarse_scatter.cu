// grad_value of the coarse pyramid levels on the 5th-generation tensor cores (second pass of the fused
// spatial cross-attention backward, see coarse_common.cuh).
//
// The reference-era design adds every bilinear corner of every sample to grad_value with a global
// atomicAdd (ops_dcnv3/src/cuda/dcnv3_im2col_cuda.cuh:106-146; mmcv's ms_deform_attn kernel does the
// same).  On the two coarse levels that is 46 % of all updates landing on 6 % of the pixels.  Here one
// CTA owns the coarse patch (<= 2048 pixels) of ONE (batch, camera, head) for a slice of that
// camera's queries and keeps it as fp32 accumulators in TENSOR MEMORY (<= 16 M-tiles x 32 columns):
//
//   for every 16 queries (one MMA K-step):
//     builder warp:  W[16 rows][patch pixels] (value dtype, MN-major canonical UMMA layout) <- the rows'
//                    sample records, zero elsewhere (the buffer is kept zero: a row's few non-zeros are
//                    written before the MMAs and erased after them);  G[16 rows][32 ch] <- g_out
//     issuing thread: for each M-tile of 128 pixels  D_tile += W_tile^T . G   (tcgen05.mma, M128 N32 K16)
//   epilogue: TMEM -> registers -> one reduction per (pixel, head) into the grad_value accumulator.
//
// No atomics inside the accumulation, no same-address contention, fp32 sums (the fp16 accumulator of
// the main pass only sees the few partial sums of the CTAs that share a (camera, head)).
// Three stages (61-64 KB of W each), FOUR builder warps per stage (a lane owns one bilinear corner of one
// level of one row and walks the level's samples: the corners of a sample are distinct pixels, samples
// of one row may share pixels and follow each other in warp lockstep), one issuing warp; all waits are
// bounded and trap instead of hanging.
#include <cstdlib>
#include <type_traits>

#include "coarse_common.cuh"
#include "msda_common.cuh"
#include "msda_host.h"

namespace msda {

constexpr int CS_BUILDERS = 4;                                          // builder warps per stage (4 rows each)
constexpr int CS_ISSUERS = 4;                                           // issuing warps (M-tiles t = w mod 4)
constexpr int CS_THREADS = 32 * (CS_ISSUERS + 3 * CS_BUILDERS);
constexpr int CS_STAGES = 3;
constexpr int CS_KROWS = 16;
constexpr int CS_TILE_BYTES = 128 / 8 * 128;                          // one M-tile of one k-group: 16 core matrices
constexpr int CS_A_BYTES = 2 * (kCoarseMaxPx / 128) * CS_TILE_BYTES;  // 65 536: two k-groups of the largest patch
constexpr int CS_B_BYTES = 2 * (kCoarseDh / 8) * 128;                 // 1 024
constexpr int CS_STAGE_BYTES = CS_A_BYTES + CS_B_BYTES;
constexpr int CS_SMEM = CS_STAGES * CS_STAGE_BYTES + 1024;
constexpr int CS_TMEM_COLS = 512;

struct CoarseArgs {
  const uint4* rec;
  const int32_t* hit_index;      // (cams, Nq) ascending query indices per camera (batch element 0's lists)
  const int32_t* hit_count;      // (cams,)
  const void* g_out;             // (bs, Nq, M * 32) value dtype
  void* g_value;                 // (bs * cams, Nk, M, 32) accumulator: fp16 scaled by *acc_scale, or fp32
  const float* acc_scale;
  const int64_t* shapes;
  const int64_t* starts;
  int bs, cams, Nq, Nk, M, L, P;
  uint32_t idesc;
  int debug;                     // MSDA_COARSE_DEBUG experiment switches (0 in production): 1 = no build / erase,
                                 // 2 = no MMAs, 4 = no record / gradient fetch
};

__device__ __forceinline__ uint64_t cs_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // no-swizzle shared-memory matrix descriptor (the same encoding as wgrad.cu, which is verified on hardware)
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
         ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ bool cs_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void cs_wait(uint64_t* bar, uint32_t parity) {
  for (unsigned spins = 0; !cs_try_wait(bar, parity); ++spins)
    if (spins > (1u << 22)) __trap();              // a lost arrival must surface as an error, never as a hang
}
__device__ __forceinline__ void cs_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// explicit shared-space accesses (the stage pointers are derived by integer arithmetic, from which the
// compiler cannot tell the address space: plain C++ accesses became generic LD / ST .STRONG.SYS)
__device__ __forceinline__ uint16_t cs_lds16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void cs_sts16(uint32_t addr, uint16_t v) {
  asm volatile("st.shared.u16 [%0], %1;" :: "r"(addr), "h"(v) : "memory");
}
__device__ __forceinline__ void cs_sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// sum of two weights in the weight dtype W (fp16, or bf16 for a bf16 model with the fp32 accumulator)
template <typename W> __device__ __forceinline__ uint16_t cs_add_bits(uint16_t a, uint16_t b);
template <> __device__ __forceinline__ uint16_t cs_add_bits<__nv_bfloat16>(uint16_t a, uint16_t b) {
  const float s = __uint_as_float((uint32_t)a << 16) + __uint_as_float((uint32_t)b << 16);
  const __nv_bfloat16 h = __float2bfloat16_rn(s);
  return *reinterpret_cast<const uint16_t*>(&h);
}
template <> __device__ __forceinline__ uint16_t cs_add_bits<__half>(uint16_t a, uint16_t b) {
  const __half h = __hadd(*reinterpret_cast<const __half*>(&a), *reinterpret_cast<const __half*>(&b));
  return *reinterpret_cast<const uint16_t*>(&h);
}
// 8 values of the model dtype T -> 8 fp16 values scaled by `scale`
template <typename T> __device__ __forceinline__ uint4 cs_to_f16_scaled(const uint4& u, float scale) {
  float f[8];
  Vec16<T>::unpack(u, f);
  return make_uint4(Vec16<__half>::pack2(f[0] * scale, f[1] * scale), Vec16<__half>::pack2(f[2] * scale, f[3] * scale),
                    Vec16<__half>::pack2(f[4] * scale, f[5] * scale), Vec16<__half>::pack2(f[6] * scale, f[7] * scale));
}

template <typename T, bool ACC_HALF>
__global__ void __launch_bounds__(CS_THREADS, 1)
coarse_scatter_kernel(const CoarseArgs a) {
  extern __shared__ unsigned char cs_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(cs_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[CS_STAGES];
  __shared__ uint64_t empty_bar[CS_STAGES];
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int s_patch[3];
  __shared__ uint32_t s_tiles[CS_STAGES][2];     // M-tiles a K-step touches (per stage, double-buffered by use count)
  __shared__ uint32_t s_init[CS_ISSUERS];        // M-tiles each issuing warp has written in the current segment
  // Operand dtypes.  fp16 accumulator: the weights are fp16 (11 significant bits whatever the model
  // dtype) and g_out is staged as fp16(g * scale) -- what the reduction path adds, too -- so the patch
  // sums come out scaled.  fp32 accumulator: both operands in the model dtype, unscaled.
  // (kind::f16 wants A and B of ONE format: fp16 x bf16 raises an illegal-instruction fault.)
  using WT = typename std::conditional<ACC_HALF, __half, T>::type;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m = blockIdx.y, b = blockIdx.z;

  if (tid == 0) {
    const CoarsePatch cp = coarse_patch(a.shapes, a.starts, a.L);
    s_patch[0] = cp.first_level;
    s_patch[1] = cp.npix;
    s_patch[2] = cp.start;
  }
  if (tid < CS_STAGES * 2) s_tiles[tid >> 1][tid & 1] = 0u;
  __syncthreads();
  const int first_level = s_patch[0], npix = s_patch[1], patch_start = s_patch[2];
  if (first_level >= a.L || npix <= 0) return;                       // no coarse patch (whole grid: uniform)
  const int nlev = a.L - first_level;                                // 1 or 2
  // This CTA's share of the (camera, hit query) rows of head m: an equal slice of the cameras' hit
  // lists laid end to end, so the CTAs finish together however unevenly the cameras see the grid.  A
  // slice that crosses a camera boundary is processed as one segment per camera (the patch in tensor
  // memory belongs to one camera at a time).
  long long total = 0;
  for (int c = 0; c < a.cams; ++c) total += a.hit_count[c];
  const long long lo = total * blockIdx.x / gridDim.x, hi = total * (blockIdx.x + 1) / gridDim.x;
  if (hi <= lo) return;                                              // (whole CTA: uniform)
  const int NT = (npix + 127) / 128;
  const uint32_t lbo_a = (uint32_t)NT * CS_TILE_BYTES;               // bytes between the two k-groups of 8 rows

  if (tid == 0) {
    for (int s = 0; s < CS_STAGES; ++s) {
      mbar_init(&full_bar[s], CS_BUILDERS);
      mbar_init(&empty_bar[s], CS_ISSUERS);
    }
    mbar_init(&done_bar, CS_ISSUERS);
    fence_barrier_init();
  }
  for (int i = tid; i < CS_STAGES * CS_STAGE_BYTES / 16; i += CS_THREADS)
    cs_sts128(smem_u32(smem) + 16u * (uint32_t)i, make_uint4(0u, 0u, 0u, 0u));
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(&tmem_slot)), "n"(CS_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // the zero fill is visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_slot;

  // K-steps are numbered through the whole kernel (`g`): step g uses stage g % CS_STAGES for the
  // (g / CS_STAGES)-th time, which fixes the parities of the stage barriers across segments.
  int g0 = 0, seg = 0;
  long long prefix = 0;
  for (int cam = 0; cam < a.cams; ++cam) {
    const int count = a.hit_count[cam];
    const long long s_lo = (lo > prefix ? lo : prefix) - prefix, s_hi = (hi < prefix + count ? hi : prefix + count) - prefix;
    prefix += count;
    if (s_lo >= s_hi) continue;                                      // (uniform)
    const int i0 = (int)s_lo, i1 = (int)s_hi;
    const int nsteps = (i1 - i0 + CS_KROWS - 1) / CS_KROWS;

    if (warp < CS_ISSUERS) {
      // ---------------- issuing warps: the touched M-tiles of a K-step (M128 N32 K16), tile t by warp t mod 4 ----
      // One thread issues an MMA in ~50-90 cycles whatever its shape (descriptor arithmetic, five
      // register-to-uniform moves, the election; tools/micro/umma_shape_bench.cu), several times the
      // tensor-pipe time of these skinny tiles -- so four warps share the tiles.  A tile always belongs to
      // the same warp (its accumulation stays in issue order); each warp walks the loop converged and one
      // ELECTED lane issues (elect.sync names the same lane every time, so a warp's commits track its MMAs).
      // A K-step's 16 neighbouring queries touch a few of the patch's tiles: the builders publish which.
      const uint32_t smem_base = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_d, 0);
      const uint32_t idesc_u = a.idesc;
      const uint32_t desc_hi = (uint32_t)(128 >> 4) | (1u << 14);      // SBO = 128 bytes | descriptor version 1
      const uint32_t lbo_field = ((lbo_a >> 4) & 0x3fffu) << 16;
      const uint32_t b_lbo_field = (uint32_t)(((kCoarseDh / 8) * 128) >> 4) << 16;
      uint32_t inited = 0u;
      for (int k = 0; k < nsteps; ++k) {
        const int g = g0 + k, st = g % CS_STAGES, use = g / CS_STAGES;
        cs_wait(&full_bar[st], (uint32_t)(use & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t tiles = *reinterpret_cast<volatile uint32_t*>(&s_tiles[st][use & 1]);
        if (a.debug & 8) tiles = 0xffffffffu;
        if (a.debug & 2) tiles = 0u;
        const uint32_t a_addr = smem_base + (uint32_t)st * CS_STAGE_BYTES;
        const uint32_t b_lo = (((a_addr + CS_A_BYTES) & 0x3ffffu) >> 4) | b_lbo_field;
        const uint32_t a_lo0 = ((a_addr & 0x3ffffu) >> 4) | lbo_field;
        for (int t = warp; t < NT; t += CS_ISSUERS) {
          if (!((tiles >> t) & 1u)) continue;
          const uint32_t accumulate = (inited >> t) & 1u;
          inited |= 1u << t;
          // descriptor of M-tile t: the start address field advances by CS_TILE_BYTES / 16
          asm volatile("{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
                       "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
                       "elect.sync _|e, 0xffffffff;\n\t"
                       "setp.ne.b32 p, %4, 0;\n\t"
                       "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
                       :: "r"(tmem_u + (uint32_t)t * kCoarseDh), "r"(a_lo0 + (uint32_t)t * (CS_TILE_BYTES >> 4)), "r"(b_lo),
                          "r"(idesc_u), "r"(accumulate), "r"(desc_hi)
                       : "memory");
        }
        // arrives when these MMAs (and all earlier ones) have completed: the stage may be rewritten
        asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                     "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
                     :: "r"(smem_u32(&empty_bar[st])) : "memory");
      }
      asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                   "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
                   :: "r"(smem_u32(&done_bar)) : "memory");

      // ---------------- epilogue: TMEM -> registers -> one reduction per (pixel, head) ----------------
      // (the issuers) warp i reads TMEM lanes [32 i, 32 i + 32); tiles no K-step touched hold nothing
      if (lane == 0) s_init[warp] = inited;
      cs_wait(&done_bar, (uint32_t)(seg & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      asm volatile("bar.sync 1, %0;" :: "n"(32 * CS_ISSUERS) : "memory");
      uint32_t written = 0u;
#pragma unroll
      for (int w = 0; w < CS_ISSUERS; ++w) written |= *reinterpret_cast<volatile uint32_t*>(&s_init[w]);
      const size_t map = (size_t)b * a.cams + cam;
      for (int t = 0; t < NT; ++t) {
        if (!((written >> t) & 1u)) continue;
        uint32_t v[32];
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)t * kCoarseDh;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int p = t * 128 + warp * 32 + lane;
        if (p < npix) {
          const size_t at = ((map * a.Nk + patch_start + p) * a.M + m) * kCoarseDh;
          if constexpr (ACC_HALF) {
            __half* dst = static_cast<__half*>(a.g_value) + at;
#pragma unroll
            for (int c = 0; c < 32; c += 8) {
              uint32_t h[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const __half2 x = __floats2half2_rn(__uint_as_float(v[c + 2 * e]), __uint_as_float(v[c + 2 * e + 1]));
                h[e] = *reinterpret_cast<const uint32_t*>(&x);
              }
              red_add_f16x8(dst + c, h[0], h[1], h[2], h[3]);
            }
          } else {
            float* dst = static_cast<float*>(a.g_value) + at;
#pragma unroll
            for (int c = 0; c < 32; c += 4)
              red_add_f32x4(dst + c, __uint_as_float(v[c]), __uint_as_float(v[c + 1]), __uint_as_float(v[c + 2]),
                            __uint_as_float(v[c + 3]));
          }
        }
      }
      // every issuer has read the patch before the next segment's first MMAs overwrite it
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("bar.sync 1, %0;" :: "n"(32 * CS_ISSUERS) : "memory");
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    } else {
      // ---------------- builder warps: CS_BUILDERS per stage, 4 rows of the K-step each ----------------
      const int st = (warp - CS_ISSUERS) / CS_BUILDERS, sub = (warp - CS_ISSUERS) % CS_BUILDERS;
      unsigned char* A = smem + (size_t)st * CS_STAGE_BYTES;
      unsigned char* B = A + CS_A_BYTES;
      // lane = (row of the warp's four, coarse level slot, bilinear corner)
      const int r = sub * 4 + (lane >> 3), slot = (lane >> 2) & 1, cn = lane & 3;
      const bool lane_on = slot < nlev;
      const uint32_t row_off = (uint32_t)(r >> 3) * lbo_a + (uint32_t)(r & 7) * 16;
      const int recs_per_row = kCoarseMaxLevels * a.P;
      const T* gout = static_cast<const T*>(a.g_out);
      const float g_scale = ACC_HALF ? __ldg(a.acc_scale) : 1.f;
      // one K-step of this lane: word `cn` of the records of its (row, level) -- byte offset of the pixel
      // inside the row's k-group : 16 | weight : 16 -- and, for the corner-0/1 lanes, one 16-byte chunk of
      // the row's upstream gradient (all zero for rows past the slice)
      struct Step {
        uint32_t w[kCoarseMaxP];
        uint4 g;
      };
      // the row's query index is read one use of the stage ahead of its records (no dependent load
      // on the critical path); k = step inside the segment
      auto fetch_index = [&](int k) -> int {
        const int i = i0 + k * CS_KROWS + r;
        return (k < nsteps && i < i1 && !(a.debug & 4)) ? __ldg(a.hit_index + (size_t)cam * a.Nq + i) : -1;
      };
      auto fetch = [&](int q, Step& d) {
        const bool ok = q >= 0;
        const size_t qq = ok ? (size_t)q : 0;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(
            a.rec + ((((size_t)b * a.cams + cam) * a.Nq + qq) * a.M + m) * recs_per_row + (size_t)slot * a.P) + cn;
#pragma unroll
        for (int j = 0; j < kCoarseMaxP; ++j) d.w[j] = (ok && lane_on && j < a.P) ? __ldg(src + 4 * j) : 0u;
        const uint4* gs = reinterpret_cast<const uint4*>(gout + (((size_t)b * a.Nq + qq) * a.M + m) * kCoarseDh) + slot * 2 + cn;
        d.g = (ok && cn < 2) ? __ldg(gs) : make_uint4(0u, 0u, 0u, 0u);
      };
      const uint32_t a_row = smem_u32(A) + row_off;
      const uint32_t b_at = smem_u32(B) + (uint32_t)(r >> 3) * ((kCoarseDh / 8) * 128) + (uint32_t)(slot * 2 + cn) * 128 +
                            (uint32_t)(r & 7) * 16;
      // samples one after the other: the warp's lanes update sample j together (distinct addresses, or
      // nothing where the weight is zero), and the __syncwarp orders sample j's stores before sample
      // j + 1's loads of a pixel that another corner lane of the same row wrote.  Returns the M-tiles
      // this lane touched (position >> 11 = pixel / 128).
      auto build = [&](const Step& d) -> uint32_t {
        uint32_t tiles = 0u;
#pragma unroll
        for (int j = 0; j < kCoarseMaxP; ++j) {
          const uint32_t w = d.w[j];
          if (w & 0x7fff0000u) {
            const uint32_t at = a_row + (w & 0xffffu);
            cs_sts16(at, cs_add_bits<WT>(cs_lds16(at), (uint16_t)(w >> 16)));
            tiles |= 1u << ((w & 0xffffu) >> 11);
          }
          __syncwarp();
        }
        return tiles;
      };
      auto erase = [&](const Step& d) {
#pragma unroll
        for (int j = 0; j < kCoarseMaxP; ++j)
          if (d.w[j] & 0x7fff0000u) cs_sts16(a_row + (d.w[j] & 0xffffu), (uint16_t)0);
      };

      // first step of this segment that falls on this warp's stage
      const int k_first = ((st - g0) % CS_STAGES + CS_STAGES) % CS_STAGES;
      Step cur, nxt;
      fetch(fetch_index(k_first), cur);
      int q_next = fetch_index(k_first + CS_STAGES);
      for (int k = k_first; k < nsteps; k += CS_STAGES) {
        const int use = (g0 + k) / CS_STAGES;
        fetch(q_next, nxt);                                            // in flight while this step is built
        q_next = fetch_index(k + 2 * CS_STAGES);
        uint32_t tiles = 0u;
        if (!(a.debug & 1)) tiles = build(cur);                        // (the stage is all-zero here)
        if (cn < 2) {
          if constexpr (ACC_HALF) cs_sts128(b_at, cs_to_f16_scaled<T>(cur.g, g_scale));
          else cs_sts128(b_at, cur.g);
        }
        tiles = __reduce_or_sync(0xffffffffu, tiles);
        if (lane == 0) {
          if (tiles) atomicOr(&s_tiles[st][use & 1], tiles);
          if (sub == 0) s_tiles[st][(use + 1) & 1] = 0u;               // (read by the issuers one use ago)
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core reads
        __syncwarp();
        if (lane == 0) cs_arrive(&full_bar[st]);
        cs_wait(&empty_bar[st], (uint32_t)(use & 1));                  // this step's MMAs have read the stage
        if (!(a.debug & 1)) erase(cur);                                // back to all-zero
        cur = nxt;
      }
    }
    g0 += nsteps;
    ++seg;
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "n"(CS_TMEM_COLS) : "memory");
}

static int cs_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

long long coarse_record_bytes(int bs, int cams, int Nq, int M, int P) {
  return (long long)bs * cams * Nq * M * kCoarseMaxLevels * P * 16;
}

bool coarse_supported(int Dh, int P, int value_dtype) {
  return Dh == kCoarseDh && P >= 1 && P <= kCoarseMaxP && (value_dtype == MSDA_BF16 || value_dtype == MSDA_F16);
}

template <typename T>
static int coarse_launch_t(const CoarseArgs& a0, bool acc_half, cudaStream_t st) {
  CoarseArgs a = a0;
  // InstrDescriptor: fp32 accumulate; A / B format (0 = f16, 1 = bf16; see the kernel for the choice); both
  // operands MN-major, N / 8, M / 16
  const uint32_t fmt = (!acc_half && std::is_same<T, __nv_bfloat16>::value) ? 1u : 0u;
  a.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(kCoarseDh >> 3) << 17) |
            ((uint32_t)(128 >> 4) << 24);
  auto kfn = acc_half ? coarse_scatter_kernel<T, true> : coarse_scatter_kernel<T, false>;
  if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM) != cudaSuccess)
    return set_error(MSDA_ERR_CUDA, "sca_coarse_scatter: cannot reserve %d bytes of shared memory", CS_SMEM);
  // one CTA per SM (shared memory, all 512 TMEM columns): the (camera, hit query) rows of every
  // (batch, head) are split evenly over enough CTAs to fill the machine once
  const int units = a.bs * a.M;
  int split = cs_sm_count() / units;
  if (split < 1) split = 1;
  const dim3 grid((unsigned)split, (unsigned)a.M, (unsigned)a.bs);
  kfn<<<grid, CS_THREADS, CS_SMEM, st>>>(a);
  count_launch();
  return check_launch("sca_coarse_scatter");
}

int launch_coarse_scatter(const void* rec, const int32_t* hit_index, const int32_t* hit_count, const void* g_out,
                          void* g_value, const float* acc_scale, int acc_half, const int64_t* shapes,
                          const int64_t* starts, int bs, int cams, int Nq, int Nk, int M, int Dh, int L, int P,
                          int value_dtype, cudaStream_t st) {
  if (!coarse_supported(Dh, P, value_dtype))
    return set_error(MSDA_ERR_UNSUPPORTED, "sca_coarse_scatter: needs a 16-bit value dtype, head_dim 32 and <= 8 points");
  CoarseArgs a{};
  a.rec = static_cast<const uint4*>(rec); a.hit_index = hit_index; a.hit_count = hit_count; a.g_out = g_out;
  a.g_value = g_value; a.acc_scale = acc_scale; a.shapes = shapes; a.starts = starts;
  a.bs = bs; a.cams = cams; a.Nq = Nq; a.Nk = Nk; a.M = M; a.L = L; a.P = P;
  static const int dbg = [] { const char* v = getenv("MSDA_COARSE_DEBUG"); return v ? atoi(v) : 0; }();
  a.debug = dbg;
  if (value_dtype == MSDA_BF16) return coarse_launch_t<__nv_bfloat16>(a, acc_half != 0, st);
  return coarse_launch_t<__half>(a, acc_half != 0, st);
}

}  // namespace msda
