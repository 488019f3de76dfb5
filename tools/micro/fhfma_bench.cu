// Issue-rate micro-benchmark: FFMA vs FFMA2 vs FHFMA (mixed bf16*bf16+f32) vs unpack+FFMA2.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fhfma_bench fhfma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float fhfma(uint16_t a, uint16_t b, float c) {
  float d; asm volatile("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c)); return d;
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(*reinterpret_cast<unsigned long long*>(&d))
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)),
        "l"(*reinterpret_cast<const unsigned long long*>(&c)));
  return d;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(const uint32_t* in, float* out, int iters) {
  uint32_t u[4]; for (int i = 0; i < 4; ++i) u[i] = in[threadIdx.x * 4 + i];
  uint32_t w = in[1024 + threadIdx.x];
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float2 acc2[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {          // 8 FFMA
      float wf = __uint_as_float(w);
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i]) : "f"(__uint_as_float(u[i & 3])), "f"(wf));
    } else if (MODE == 1) {   // 4 FFMA2
      float2 wf = make_float2(__uint_as_float(w), __uint_as_float(w));
#pragma unroll
      for (int i = 0; i < 4; ++i) acc2[i] = ffma2(make_float2(__uint_as_float(u[i]), __uint_as_float(u[(i + 1) & 3])), wf, acc2[i]);
    } else if (MODE == 2) {   // 8 FHFMA
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[2 * i] = fhfma((uint16_t)(u[i] & 0xffff), (uint16_t)(w & 0xffff), acc[2 * i]);
        acc[2 * i + 1] = fhfma((uint16_t)(u[i] >> 16), (uint16_t)(w & 0xffff), acc[2 * i + 1]);
      }
    } else {                  // 8 unpack + 4 FFMA2 (the old inner loop)
      float2 wf = make_float2(__uint_as_float(w), __uint_as_float(w));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t lo, hi;
        asm volatile("shl.b32 %0, %1, 16;" : "=r"(lo) : "r"(u[i]));
        asm volatile("and.b32 %0, %1, 0xffff0000;" : "=r"(hi) : "r"(u[i]));
        acc2[i] = ffma2(make_float2(__uint_as_float(lo), __uint_as_float(hi)), wf, acc2[i]);
      }
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += acc[i];
  for (int i = 0; i < 4; ++i) s += acc2[i].x + acc2[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> float run(const uint32_t* in, float* out, int blocks, int iters) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<blocks, 256>>>(in, out, iters);
  cudaEventRecord(a);
  k<MODE><<<blocks, 256>>>(in, out, iters);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int blocks = sms * 8, iters = 1 << 15;
  uint32_t* in; float* out;
  cudaMalloc(&in, 8192 * 4); cudaMemset(in, 0x3c, 8192 * 4); cudaMalloc(&out, blocks * 256 * 4);
  const char* names[4] = {"8xFFMA", "4xFFMA2", "8xFHFMA", "8xALU+4xFFMA2"};
  float ms[4] = {run<0>(in, out, blocks, iters), run<1>(in, out, blocks, iters), run<2>(in, out, blocks, iters), run<3>(in, out, blocks, iters)};
  for (int m = 0; m < 4; ++m) {
    // warp-iterations per SM sub-partition: blocks*8 warps / (sms*4) * iters
    double warp_iters = (double)blocks * 8 / (sms * 4) * iters;
    double cyc = ms[m] * 1e-3 * clk * 1e3 / warp_iters;
    printf("%-16s %8.3f ms  ~%.2f cycles per iteration per SMSP (at %d MHz nominal)\n", names[m], ms[m], cyc, clk / 1000);
  }
  return 0;
}
