// tcgen05.mma cost for skinny shapes, shared-memory operands, no swizzle: cycles per MMA (M128, K16, kind::f16)
// as a function of N and of the majorness of A and B.  One CTA per SM, one issuing thread, `tiles` distinct A
// tiles / D tiles walked round-robin, one commit at the end; clock64 around issue -> completion.
// Motivation: coarse_scatter.cu issues 15 MMAs of M128 N32 K16 per K-step with MN-major operands.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_shape_bench umma_shape_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

struct Cfg {
  int N, a_mn, b_mn, tiles, reps, same_d;
};

__global__ void __launch_bounds__(128, 1) k(Cfg c, long long* out) {
  extern __shared__ unsigned char raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < 180 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  // A tile: 128 (M) x 16 (K) fp16 = 4 KB.  MN-major: core(m8, k8) at m8 * 128 + k8 * LBO (LBO = tiles * 2048), SBO 128.
  //                                        K-major:  core(m8, k8) at m8 * 256 + k8 * 128: SBO 256, LBO 128.
  // B tile: 16 (K) x N.  MN-major: core(n8, k8) at n8 * 128 + k8 * (N / 8 * 128).  K-major: n8 * 256 + k8 * 128.
  const uint32_t a_base = smem_u32(smem), b_base = a_base + 160 * 1024;
  const uint32_t idesc = (1u << 4) | ((uint32_t)c.a_mn << 15) | ((uint32_t)c.b_mn << 16) | ((uint32_t)(c.N >> 3) << 17) |
                         ((uint32_t)(128 >> 4) << 24);
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint64_t bdesc = c.b_mn ? desc(b_base, (c.N / 8) * 128, 128) : desc(b_base, 128, 256);
    t0 = clock64();
    for (int r = 0; r < c.reps; ++r) {
      for (int t = 0; t < c.tiles; ++t) {
        const uint64_t adesc = c.a_mn ? desc(a_base + t * 2048, c.tiles * 2048, 128) : desc(a_base + t * 4096, 128, 256);
        const uint32_t d = tmem + (c.same_d ? 0 : (uint32_t)((t * c.N) % 512));
        const uint32_t acc = r > 0;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     :: "r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

int main() {
  long long* out;
  cudaMalloc(&out, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int Ns[] = {16, 32, 64, 128, 256};
  for (int a_mn = 1; a_mn >= 0; --a_mn)
    for (int b_mn = 1; b_mn >= 0; --b_mn)
      for (int N : Ns)
        for (int tiles : {1, 15}) {
          if (tiles * N > 512 && tiles > 1) continue;
          Cfg c{N, a_mn, b_mn, tiles, 64, 0};
          k<<<148, 128, 200 * 1024>>>(c, out);
          long long cyc = 0;
          cudaError_t e = cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
          if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
          printf("A %s-major  B %s-major  N %3d  tiles %2d : %7.1f cycles / MMA\n", a_mn ? "MN" : " K", b_mn ? "MN" : " K", N,
                 tiles, (double)cyc / (c.reps * tiles));
        }
  return 0;
}
