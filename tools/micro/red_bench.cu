// Micro-benchmark: throughput of vector reductions into a (pixels, 256 ch) accumulator with the
// access pattern of the fused backward (each 4-lane group targets one (pixel, head) slice).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// mode 0: f32 v4, lane c -> [4c,4c+4) and [16+4c, ...)  (two 16 B reds, whole sectors per instruction)
// mode 1: f32 v4, lane c -> [8c, 8c+4), [8c+4, 8c+8)    (half sectors per instruction)
// mode 2: f16x2 v4 (one 16 B red = 8 channels)
// mode 3: bf16x2 v4
// mode 4: f32 v4 with 8 lanes per head (full 128 B line per instruction)
template <int MODE>
__global__ void red_kernel(void* acc, int npix, int iters, uint32_t seed, int locality) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int lanes_per_row = (MODE == 4) ? 8 : 4;
  const int row = tid / lanes_per_row;      // (query, head) row
  const int c = tid % lanes_per_row;
  const int head = row & 7;
  const int query = row >> 3;
  for (int it = 0; it < iters; ++it) {
    uint32_t h = hash32(seed + (uint32_t)(locality ? (query / locality) : query) * 131u + it * 7919u + (locality ? 0 : head * 17));
    const int pix = h % npix;
    if (MODE == 0 || MODE == 1 || MODE == 4) {
      float* base = (float*)acc + ((size_t)pix * 8 + head) * 32;
      if (MODE == 0) {
        asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" :: "l"(base + 4 * c), "f"(1.0f) : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" :: "l"(base + 16 + 4 * c), "f"(1.0f) : "memory");
      } else if (MODE == 1) {
        asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" :: "l"(base + 8 * c), "f"(1.0f) : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" :: "l"(base + 8 * c + 4), "f"(1.0f) : "memory");
      } else {
        asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" :: "l"(base + 4 * c), "f"(1.0f) : "memory");
      }
    } else if (MODE == 2) {
      __half* base = (__half*)acc + ((size_t)pix * 8 + head) * 32 + 8 * c;
      const uint32_t one = 0x3c003c00u;
      asm volatile("red.global.add.noftz.v4.f16x2 [%0], {%1,%1,%1,%1};" :: "l"(base), "r"(one) : "memory");
    } else {
      __nv_bfloat16* base = (__nv_bfloat16*)acc + ((size_t)pix * 8 + head) * 32 + 8 * c;
      const uint32_t one = 0x3f803f80u;
      asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1,%1,%1,%1};" :: "l"(base), "r"(one) : "memory");
    }
  }
}

template <int MODE>
float run(void* acc, int npix, int rows, int iters, int locality) {
  const int lanes = (MODE == 4) ? 8 : 4;
  const long long threads = (long long)rows * lanes;
  const int grid = (int)((threads + 255) / 256);
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  red_kernel<MODE><<<grid, 256>>>(acc, npix, iters, 1u, locality);
  cudaDeviceSynchronize();
  cudaEventRecord(s);
  red_kernel<MODE><<<grid, 256>>>(acc, npix, iters, 2u, locality);
  cudaEventRecord(e); cudaEventSynchronize(e);
  float ms; cudaEventElapsedTime(&ms, s, e);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) printf("error %s\n", cudaGetErrorString(err));
  return ms;
}

int main() {
  const int npix = 6 * 30825;                 // base SCA value map
  void* acc; cudaMalloc(&acc, (size_t)npix * 256 * 4); cudaMemset(acc, 0, (size_t)npix * 256 * 4);
  const int rows = 40000 * 8;                 // (query, head)
  const int iters = 144;                      // ~ 36 samples x 4 corners -> 46 M corner updates
  const double upd = (double)rows * iters;
  const char* names[5] = {"f32 v4 x2 sector-aligned", "f32 v4 x2 half-sector", "f16x2 v4", "bf16x2 v4", "f32 v4 8 lanes/row"};
  for (int loc = 0; loc <= 64; loc = loc ? loc * 8 : 1) {
    float t[5];
    t[0] = run<0>(acc, npix, rows, iters, loc); t[1] = run<1>(acc, npix, rows, iters, loc);
    t[2] = run<2>(acc, npix, rows, iters, loc); t[3] = run<3>(acc, npix, rows, iters, loc);
    t[4] = run<4>(acc, npix, rows, iters, loc);
    for (int m = 0; m < 5; ++m)
      printf("locality=%d  %-26s %8.1f us  %6.1f G corner-updates/s\n", loc, names[m], t[m] * 1e3, upd / (t[m] * 1e-3) / 1e9);
  }
  return 0;
}
