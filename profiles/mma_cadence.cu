// Micro-benchmark: how fast can ONE thread feed the sm_100a tensor core with small tcgen05.mma instructions?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mma_cadence profiles/mma_cadence.cu && build/mma_cadence
// For M = 128 (cta_group::1), K = 16, N in {64, 128, 256}; A from shared memory or from TMEM; B K-major or MN-major;
// 1 or 2 resident CTAs per SM.  Prints SM cycles per MMA instruction (issue loop + completion, 512 back-to-back MMAs)
// next to the ideal 128*N*16*2 / 8192 FLOP-per-clock figure.  Operands are zero-filled shared memory (timing only).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int m, int n, int b_mn_major) {
  return (1u << 4) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);   // f16 x f16 -> f32
}

template <int N, bool A_TMEM, bool B_MN>
__global__ void __launch_bounds__(128) cadence_kernel(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = make_idesc(128, N, B_MN ? 1 : 0);
    const uint64_t da = make_smem_desc(base), db = make_smem_desc(base + 16384);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t d = tmem + (uint32_t)((i & 1) * (N > 64 ? 0 : 64));      // alternate accumulators when they fit next to A
      if (A_TMEM)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                     ::"r"(d), "r"(tmem + 192), "l"(db), "r"(idesc), "r"(1u) : "memory");
      else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
  }
}

template <int N, bool A_TMEM, bool B_MN>
static void run(const char* name, int ctas_per_sm) {
  long long* d;
  cudaMalloc(&d, 16);
  const int smem = 16384 + 32768 + 1024 + (ctas_per_sm == 1 ? 64 * 1024 : 0);     // pad so that only one CTA fits when asked
  cudaFuncSetAttribute(cadence_kernel<N, A_TMEM, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const int iters = 512;
  const int grid = 148 * ctas_per_sm;
  const int sm_bytes = ctas_per_sm == 1 ? 150 * 1024 : smem;
  for (int rep = 0; rep < 2; ++rep) cadence_kernel<N, A_TMEM, B_MN><<<grid, 128, sm_bytes>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-34s ctas/SM %d : issue %.1f clk/MMA, issue+drain %.1f clk/MMA, ideal %.1f  (%s)\n", name, ctas_per_sm, (double)h[0] / iters,
         (double)h[1] / iters, 128.0 * N * 16 * 2 / 8192.0, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int c = 1; c <= 2; ++c) {
    run<64, false, false>("N=64  A smem  B K-major", c);
    run<64, false, true>("N=64  A smem  B MN-major", c);
    run<64, true, true>("N=64  A TMEM  B MN-major", c);
    run<128, false, false>("N=128 A smem  B K-major", c);
    run<128, true, true>("N=128 A TMEM  B MN-major", c);
    run<256, false, false>("N=256 A smem  B K-major", c);
  }
  return 0;
}
