// Weight and bias gradients of a Linear layer (16-bit dtypes) on the 5th-generation tensor cores:
//     dW[o, i] = sum_n dy[n, o] * x[n, i]        db[o] = sum_n dy[n, o]
// for the shapes of the BEVFormer layer's backward: a tiny output (O, I <= 768) and a huge
// reduction (N = 40 000 ... 185 000 rows).  The library runs these as a split-K GEMM + a reduction
// kernel, and the bias gradient needs a separate pass over dy.  Here the rows are split over the
// whole grid; every CTA streams its slab once through shared memory (cp.async, 3 stages) into
// tcgen05.mma with the 128 x 256 fp32 tile in tensor memory, the column sums of dy ride along from the
// same shared-memory tiles, the tile goes TMEM -> registers -> shared -> coalesced fp32 reductions into
// an L2-resident scratch, and a second small launch converts scratch -> outputs and re-zeroes it.
//
// Status (profiles/r01_wgrad.md): results equal the library's to the last bit of bf16 rounding on every
// shape tried, but at 40 000 x 256 x 256 the pair of launches takes 28.5 + 5.5 us against 18 + 5 us for
// the library's split-K GEMM (+ 14-17 us for our column-sum kernel where the bias gradient is not
// already fused into LayerNorm's backward): the split-row reduction of 74 partial tiles through L2 and
// the serial prologue / epilogue of a one-CTA-per-SM grid cost more than the fused bias sum saves, and
// the measured step is 0.6 ms slower with it.  It is therefore OPT-IN (APOLLO_B200_WGRAD=1) until the
// epilogue is reworked (cluster-level reduction of the partial tiles, two CTAs per SM); an earlier
// mma.sync version of the same decomposition sustained only ~140 TFLOP/s (a sixteenth of tcgen05).
#include <cstdlib>
#include <type_traits>

#include "msda_common.cuh"
#include "msda_host.h"

namespace msda {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const uint32_t s = smem_u32(smem);
  const int n = valid ? 16 : 0;     // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(s), "l"(gmem), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory");
}
// scratch -> outputs in the layer dtype, scratch re-zeroed (second launch: every SM takes part, where a
// "last CTA finishes" tail would convert the whole O x I tile with one CTA's worth of memory parallelism)
template <typename T>
__global__ void __launch_bounds__(256)
wgrad_finalize_kernel(float* __restrict__ ws, T* __restrict__ dW, T* __restrict__ db, long long nW, int O) {
  float* wsW = ws + kWgHeaderFloats;
  float* wsB = wsW + nW;
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; e < nW; e += stride) {
    const float4 v = __ldcg(reinterpret_cast<const float4*>(wsW + e));
    *reinterpret_cast<float4*>(wsW + e) = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<uint2*>(dW + e) = make_uint2(Vec16<T>::pack2(v.x, v.y), Vec16<T>::pack2(v.z, v.w));
  }
  if (db != nullptr && blockIdx.x == 0)
    for (int o = threadIdx.x; o < O; o += blockDim.x) {
      db[o] = from_f32<T>(__ldcg(wsB + o));
      wsB[o] = 0.f;
    }
}

// =====================================================================================================
// One thread issues tcgen05.mma for the whole CTA; the accumulator tile lives in tensor memory.  Both
// operands are MN-major (the output dimension is the contiguous one of dy and x), staged by cp.async straight into the
// canonical no-swizzle UMMA layout: 16-byte chunk (row k, chunk j) of a tile goes to
// (k / 8) * LBO + j * SBO + (k % 8) * 16 -- core matrices of 8 rows x 16 bytes; SBO carries 16 bytes of
// padding so neither the cp.async writes nor the bias column sums conflict on banks.
// =====================================================================================================
constexpr int UM_BM = 128, UM_BN = 256, UM_BK = 64, UM_STAGES = 3;
constexpr int UM_SBO = 144;                                  // bytes between MN-adjacent core matrices
constexpr int UM_LBO_A = (UM_BM / 8) * UM_SBO;               // bytes between k-groups of 8 rows (A: dy tile)
constexpr int UM_LBO_B = (UM_BN / 8) * UM_SBO;               //                                  (B: x tile)
constexpr int UM_STAGE_A = (UM_BK / 8) * UM_LBO_A;           // 18 432 bytes
constexpr int UM_STAGE_B = (UM_BK / 8) * UM_LBO_B;           // 36 864 bytes
constexpr int UM_STAGE = UM_STAGE_A + UM_STAGE_B;
constexpr int UM_EPI_PITCH = UM_BN + 4;                      // floats per staged output row
constexpr int UM_SMEM = (UM_STAGES * UM_STAGE > UM_BM * UM_EPI_PITCH * 4 ? UM_STAGES * UM_STAGE
                                                                         : UM_BM * UM_EPI_PITCH * 4) + 1024;
constexpr int UM_TMEM_COLS = 256;

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46,
  // no swizzle (layout type 0 in bits 61..63)
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
         ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ bool mbar_try_wait_parity(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  for (unsigned spins = 0; !mbar_try_wait_parity(bar, parity); ++spins)
    if (spins > (1u << 24)) __trap();          // a lost arrival must surface as an error, never as a hang
}

template <typename T>
__global__ void __launch_bounds__(256, 1)
wgrad_umma_kernel(const T* __restrict__ dy, const T* __restrict__ x, T* __restrict__ db, float* __restrict__ ws,
                  long long N, int O, int I, long long rows_per_slab, uint32_t idesc) {
  extern __shared__ unsigned char um_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(um_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t empty_bar[UM_STAGES];
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_base_slot;
  __shared__ float bias_part[16][UM_BM];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int o0 = blockIdx.y * UM_BM, i0 = blockIdx.z * UM_BN;
  const long long n0 = (long long)blockIdx.x * rows_per_slab;
  const long long n1 = n0 + rows_per_slab < N ? n0 + rows_per_slab : N;
  const int chunks = n1 > n0 ? (int)((n1 - n0 + UM_BK - 1) / UM_BK) : 0;
  if (chunks == 0) return;                                     // (whole CTA: uniform)
  const bool do_bias = (db != nullptr) && blockIdx.z == 0;

  if (tid == 0) {
    for (int s = 0; s < UM_STAGES; ++s) mbar_init(&empty_bar[s], 1);
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(&tmem_base_slot)), "n"(UM_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_slot;

  auto load_stage = [&](int stage, int chunk) {
    const long long r0 = n0 + (long long)chunk * UM_BK;
    unsigned char* a = smem + (size_t)stage * UM_STAGE;
    unsigned char* b = a + UM_STAGE_A;
    for (int p = tid; p < UM_BK * (UM_BM / 8); p += 256) {      // dy tile: 64 rows x 16 chunks
      const int r = p / (UM_BM / 8), j = p % (UM_BM / 8);
      const bool ok = (r0 + r < n1) && (o0 + j * 8 < O);
      cp_async16(a + (r >> 3) * UM_LBO_A + j * UM_SBO + (r & 7) * 16, ok ? dy + (r0 + r) * O + o0 + j * 8 : dy, ok);
    }
    for (int p = tid; p < UM_BK * (UM_BN / 8); p += 256) {      // x tile: 64 rows x 32 chunks
      const int r = p / (UM_BN / 8), j = p % (UM_BN / 8);
      const bool ok = (r0 + r < n1) && (i0 + j * 8 < I);
      cp_async16(b + (r >> 3) * UM_LBO_B + j * UM_SBO + (r & 7) * 16, ok ? x + (r0 + r) * I + i0 + j * 8 : x, ok);
    }
  };

  float bsum[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) bsum[i] = 0.f;

#pragma unroll
  for (int s = 0; s < UM_STAGES - 1; ++s) {
    if (s < chunks) load_stage(s, s);
    cp_async_commit();
  }
  for (int ch = 0; ch < chunks; ++ch) {
    const int stage = ch % UM_STAGES;
    cp_async_wait<UM_STAGES - 2>();                             // this thread's copies of chunk ch have landed
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // ... and are visible to the tensor core's proxy
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_addr = smem_u32(smem + (size_t)stage * UM_STAGE);
      const uint32_t b_addr = a_addr + UM_STAGE_A;
#pragma unroll
      for (int ks = 0; ks < UM_BK / 16; ++ks) {                 // one MMA = 16 rows = two k-groups
        const uint64_t adesc = umma_desc(a_addr + ks * 2 * UM_LBO_A, UM_LBO_A, UM_SBO);
        const uint64_t bdesc = umma_desc(b_addr + ks * 2 * UM_LBO_B, UM_LBO_B, UM_SBO);
        const uint32_t accumulate = (ch > 0 || ks > 0) ? 1u : 0u;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
      }
      // arrives on the stage's barrier when these MMAs (and all earlier ones) have completed
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                   :: "r"(smem_u32(&empty_bar[stage])) : "memory");
    }
    if (do_bias) {                                              // column sums of the dy tile (generic-proxy reads)
      const unsigned char* a = smem + (size_t)stage * UM_STAGE;
      const int j = tid & 15;
#pragma unroll
      for (int it = 0; it < UM_BK / 16; ++it) {
        const int r = (tid >> 4) + it * 16;
        const uint4 v = *reinterpret_cast<const uint4*>(a + (r >> 3) * UM_LBO_A + j * UM_SBO + (r & 7) * 16);
        float f[8];
        Vec16<T>::unpack(v, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) bsum[i] += f[i];
      }
    }
    // refill the stage consumed at iteration ch - 1 with chunk ch + STAGES - 1, once its MMAs are done
    const int nxt = ch + UM_STAGES - 1;
    if (nxt < chunks) {
      if (ch >= 1) mbar_wait_bounded(&empty_bar[(ch - 1) % UM_STAGES], (uint32_t)(((ch - 1) / UM_STAGES) & 1));
      load_stage(nxt % UM_STAGES, nxt);
    }
    cp_async_commit();
  }
  cp_async_wait<0>();
  if (tid == 0)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 :: "r"(smem_u32(&done_bar)) : "memory");
  mbar_wait_bounded(&done_bar, 0u);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  __syncthreads();                                              // every stage is free: reuse it for the epilogue

  // TMEM -> registers -> shared (row-major fp32 tile) -> coalesced reductions into the scratch
  float* tile = reinterpret_cast<float*>(smem);
  {
    const int row = (warp & 3) * 32 + lane;                     // TMEM lane == output row
    const int colh = (warp >> 2) * (UM_BN / 2);
#pragma unroll
    for (int cb = 0; cb < UM_BN / 2; cb += 32) {
      uint32_t r[32];
      const uint32_t taddr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(colh + cb);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
            "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
            "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float* dst = tile + (size_t)row * UM_EPI_PITCH + colh + cb;
#pragma unroll
      for (int k = 0; k < 32; k += 4)
        *reinterpret_cast<uint4*>(dst + k) = make_uint4(r[k], r[k + 1], r[k + 2], r[k + 3]);
    }
  }
  if (do_bias) {
#pragma unroll
    for (int i = 0; i < 8; ++i) bias_part[tid >> 4][(tid & 15) * 8 + i] = bsum[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "n"(UM_TMEM_COLS) : "memory");

  float* wsW = ws + kWgHeaderFloats;
  float* wsB = wsW + (size_t)O * I;
  for (int idx = tid; idx < UM_BM * (UM_BN / 4); idx += 256) {
    const int row = idx / (UM_BN / 4), c4 = (idx % (UM_BN / 4)) * 4;
    const int o = o0 + row, i = i0 + c4;
    if (o < O && i < I) {                                       // I is a multiple of 8
      const float4 v = *reinterpret_cast<const float4*>(tile + (size_t)row * UM_EPI_PITCH + c4);
      red_add_f32x4(wsW + (size_t)o * I + i, v.x, v.y, v.z, v.w);
    }
  }
  if (do_bias && tid < UM_BM && o0 + tid < O) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += bias_part[k][tid];
    atomicAdd(wsB + o0 + tid, s);
  }
}

static int wg_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

long long wgrad_workspace_floats(int O, int I) { return kWgHeaderFloats + (long long)O * I + O; }

template <typename T>
static int wgrad_finalize(void* dW, void* db, float* ws, int O, int I, cudaStream_t st) {
  const long long nW = (long long)O * I;
  const int fgrid = (int)((nW / 4 + 255) / 256 < 4 * wg_sm_count() ? (nW / 4 + 255) / 256 : 4 * wg_sm_count());
  wgrad_finalize_kernel<T><<<fgrid, 256, 0, st>>>(ws, static_cast<T*>(dW), static_cast<T*>(db), nW, O);
  count_launch();
  return check_launch("linear_wgrad(finalize)");
}

template <typename T>
static int wgrad_umma_t(const void* dy, const void* x, void* dW, void* db, float* ws, long long N, int O, int I,
                        cudaStream_t st) {
  const int ty = (O + UM_BM - 1) / UM_BM, tz = (I + UM_BN - 1) / UM_BN;
  long long slabs = wg_sm_count() / (ty * tz);                 // one CTA per SM (166 KB of shared memory)
  if (slabs < 1) slabs = 1;
  const long long max_slabs = (N + UM_BK - 1) / UM_BK;
  if (slabs > max_slabs) slabs = max_slabs > 0 ? max_slabs : 1;
  long long rows = (N + slabs - 1) / slabs;
  rows = (rows + UM_BK - 1) / UM_BK * UM_BK;
  slabs = N > 0 ? (N + rows - 1) / rows : 1;
  auto kfn = wgrad_umma_kernel<T>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, UM_SMEM) != cudaSuccess)
      return set_error(MSDA_ERR_CUDA, "linear_wgrad: cannot reserve %d bytes of shared memory", UM_SMEM);
    attr_set = true;
  }
  // InstrDescriptor (cute/arch/mma_sm100_desc.hpp): fp32 accumulate, A/B format (0 = f16, 1 = bf16), both
  // operands MN-major, N / 8, M / 16
  const uint32_t fmt = sizeof(T) == 2 && std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) |
                         ((uint32_t)(UM_BN >> 3) << 17) | ((uint32_t)(UM_BM >> 4) << 24);
  if (N > 0) {
    const dim3 grid((unsigned)slabs, ty, tz);
    kfn<<<grid, 256, UM_SMEM, st>>>(static_cast<const T*>(dy), static_cast<const T*>(x), static_cast<T*>(db), ws, N, O,
                                    I, rows, idesc);
    count_launch();
    if (int rc = check_launch("linear_wgrad")) return rc;
  }
  return wgrad_finalize<T>(dW, db, ws, O, I, st);
}

int launch_wgrad(const void* dy, const void* x, void* dW, void* db, float* ws, long long N, int O, int I,
                 int dtype, cudaStream_t st) {
  if (dtype == MSDA_BF16) return wgrad_umma_t<__nv_bfloat16>(dy, x, dW, db, ws, N, O, I, st);
  if (dtype == MSDA_F16) return wgrad_umma_t<__half>(dy, x, dW, db, ws, N, O, I, st);
  return set_error(MSDA_ERR_UNSUPPORTED, "linear_wgrad: 16-bit dtypes only (fp32 uses the library GEMM)");
}

}  // namespace msda
