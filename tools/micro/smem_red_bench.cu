// Micro-benchmark for the planned tile-level pre-reduction of grad_value (DESIGN.md section 9,
// profiles/r01_scatter_reuse.md): the corner updates of the two coarse pyramid levels of one head of an
// 8x8 BEV tile (64 rows x 64 corner updates, ~224 distinct (pixel, head) slots) either
//   A) go straight to L2 as 16-byte red.global.add.v4.f16x2 (what sca_bwd does today), or
//   B) are added into int32 fixed-point slots in shared memory with native ATOMS.ADD (one IMAD + one
//      ATOMS per channel, channel rotation per row against bank conflicts) and flushed with ONE global
//      reduction per distinct slot.
// sm_100a has no native floating-point shared-memory add (ATOMS.CAST.SPIN loops), hence the integers.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_red_bench smem_red_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

constexpr int kRows = 64;        // rows (queries of one head) per CTA
constexpr int kThreads = 256;    // 4 lanes per row

__device__ __forceinline__ void red_f16x8(__half* dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("red.global.add.noftz.v4.f16x2 [%0], {%1,%2,%3,%4};" :: "l"(dst), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// A: every corner update is a global reduction.  The CTA's updates fall on `nslots` distinct pixels.
__global__ void direct_kernel(__half* acc, int npix, int iters, int nslots, uint32_t seed) {
  const int row = threadIdx.x >> 2, c = threadIdx.x & 3;
  const int head = blockIdx.x & 7;
  const uint32_t tile = blockIdx.x >> 3;
  for (int it = 0; it < iters; ++it) {
    const uint32_t slot = hash32(seed + row * 131u + it * 7919u) % (uint32_t)nslots;
    const uint32_t pix = hash32(tile * 977u + slot) % (uint32_t)npix;
    const __half2 w = __float2half2_rn(1.0f / (float)(1 + (it & 7)));
    const uint32_t v = *reinterpret_cast<const uint32_t*>(&w);
    red_f16x8(acc + ((size_t)pix * 8 + head) * 32 + 8 * c, v, v, v, v);
  }
}

// B: int32 slots in shared memory, flush at the end.
template <bool ROTATE>
__global__ void staged_kernel(__half* acc, int npix, int iters, int nslots, uint32_t seed, float inv_scale) {
  extern __shared__ int slots[];                 // [nslots][32]
  const int row = threadIdx.x >> 2, c = threadIdx.x & 3;
  const int head = blockIdx.x & 7;
  const uint32_t tile = blockIdx.x >> 3;
  for (int i = threadIdx.x; i < nslots * 32; i += kThreads) slots[i] = 0;
  __syncthreads();
  int g[8];                                      // the row's g_out channels in 12-bit fixed point
#pragma unroll
  for (int k = 0; k < 8; ++k) g[k] = (int)(hash32(seed + threadIdx.x * 8 + k) & 4095u) - 2048;
  for (int it = 0; it < iters; ++it) {
    const uint32_t slot = hash32(seed + row * 131u + it * 7919u) % (uint32_t)nslots;
    const int aw = 4096 / (1 + (it & 7));        // corner weight, 12-bit fixed point
    int* s = slots + slot * 32;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int grp = ROTATE ? ((k + row) & 7) : k;      // row r starts at bank group r
      atomicAdd(s + grp * 4 + c, aw * g[k]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nslots * 4; i += kThreads) {
    const int slot = i >> 2, cc = i & 3;
    const int* s = slots + slot * 32 + cc * 8;
    uint32_t h[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __half2 v = __floats2half2_rn((float)s[2 * k] * inv_scale, (float)s[2 * k + 1] * inv_scale);
      h[k] = *reinterpret_cast<const uint32_t*>(&v);
    }
    const uint32_t pix = hash32(tile * 977u + slot) % (uint32_t)npix;
    red_f16x8(acc + ((size_t)pix * 8 + head) * 32 + 8 * cc, h[0], h[1], h[2], h[3]);
  }
}

template <typename F>
float timeit(F launch) {
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  launch(1u);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(s);
    launch(2u + r);
    cudaEventRecord(e); cudaEventSynchronize(e);
    float ms; cudaEventElapsedTime(&ms, s, e);
    best = ms < best ? ms : best;
  }
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) printf("error %s\n", cudaGetErrorString(err));
  return best * 1e3f;
}

int main() {
  const int npix = 6 * 30825;
  __half* acc; cudaMalloc(&acc, (size_t)npix * 256 * 2); cudaMemset(acc, 0, (size_t)npix * 256 * 2);
  const int ctas = 40000 * 8 / kRows;            // (tile, head) CTAs: 5000
  const int iters = 64;                          // corner updates per row on the two coarse levels
  const double upd = (double)ctas * kRows * iters;
  printf("%d CTAs x %d rows x %d corner updates = %.1f M updates\n", ctas, kRows, iters, upd / 1e6);
  for (int nslots : {56, 112, 224, 448}) {
    const size_t smem = (size_t)nslots * 32 * sizeof(int);
    cudaFuncSetAttribute(staged_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(staged_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const float a = timeit([&](uint32_t sd) { direct_kernel<<<ctas, kThreads>>>(acc, npix, iters, nslots, sd); });
    const float b = timeit([&](uint32_t sd) { staged_kernel<true><<<ctas, kThreads, smem>>>(acc, npix, iters, nslots, sd, 1.f / 16777216.f); });
    const float c = timeit([&](uint32_t sd) { staged_kernel<false><<<ctas, kThreads, smem>>>(acc, npix, iters, nslots, sd, 1.f / 16777216.f); });
    printf("distinct slots per CTA %4d (%.3f of the updates, %5.1f KB int32):  direct red.global %7.1f us | "
           "shared int32 + flush %7.1f us (rotated banks) %7.1f us (plain)\n",
           nslots, (double)nslots / (kRows * iters), smem / 1024.0, a, b, c);
  }
  return 0;
}
