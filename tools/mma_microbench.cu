// Micro-benchmark of tcgen05.mma issue patterns used by the convolution kernels
// (not part of the library).  One CTA per SM, one thread issues a long stream of
// kind::f16 MMAs (M = 128, K = 16) on zero-filled shared memory; prints cycles
// per MMA for several N / operand-layout variants.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_microbench mma_microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}


// Fully unrolled issue loop: descriptors are advanced by adding compile-time
// constants to the low word, so one MMA costs ~3 SASS instructions of issue.
template <int N, int NT, int PW, int M, int LBO16 = -1>     // LBO16 >= 0: the second K half is ANOTHER TAP, that many pixels away (8-channel layers)
__global__ void __launch_bounds__(128, 1) bench(int iters, unsigned long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) |
                               ((uint32_t)(M >> 4) << 24);
    const uint32_t a_base = smem_u32(smem), b_base = a_base + 64 * 1024;
    const uint64_t a0 = umma_desc(a_base, LBO16 >= 0 ? LBO16 * 16 : 18 * PW * 16, PW * 16);
    const uint64_t b0 = umma_desc(b_base, N * 16, 128);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const uint64_t bd = b0 + (uint64_t)((tap * 2 * N * 16) >> 4);
#pragma unroll
        for (int i = 0; i < NT; ++i) {
          const uint64_t ad = a0 + (uint64_t)((((tap / 3) * PW + 8 * i + tap % 3) * 16) >> 4);
          mma(tmem + i * N, ad, bd, idesc, 1);
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    const long long t1 = clock64();
    out[blockIdx.x] = (unsigned long long)(t1 - t0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// The issue pattern of conv_tcg for split operands (S = 2): per entry and tile, a_hi x [w_hi | w_lo]
// (2N columns) followed by a_lo x w_hi (N columns) into the SECOND half of the same accumulator.
// OVERLAP = false: the second MMA goes to a separate accumulator instead (same work, no dependency).
template <int N, int NT, bool OVERLAP>
__global__ void __launch_bounds__(128, 1) bench_split(int iters, unsigned long long* out, int random_data) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u; h ^= h >> 13; h *= 0x5bd1e995u;
    // two fp16 values in [-2, 2): sign | exponent 01100..01111 | random mantissa
    const uint32_t lo = (h & 0x83ffu) | (((h >> 16) & 3u) + 12u) << 10, hi = ((h >> 3) & 0x83ffu) | (((h >> 20) & 3u) + 12u) << 10;
    ((uint32_t*)smem)[i] = random_data ? (lo | (hi << 16)) : 0u;
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t base = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t id0 = base | ((uint32_t)(2 * N >> 3) << 17), id1 = base | ((uint32_t)(N >> 3) << 17);
    constexpr int PW = 10;
    const uint32_t a_base = smem_u32(smem), b_base = a_base + 96 * 1024;
    const uint64_t a0 = umma_desc(a_base, 16, PW * 16);                  // tap pair: second K half one pixel away
    const uint64_t b0 = umma_desc(b_base, 2 * N * 16, 128);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int e = 0; e < 9; ++e) {
        const uint64_t bd = b0 + (uint64_t)((e * 2 * 2 * N * 16) >> 4);
#pragma unroll
        for (int i = 0; i < NT; ++i) {
          const uint64_t ad = a0 + (uint64_t)((e / 3) * PW + e % 3 + i * 180);
          mma(tmem + i * 2 * N, ad, bd, id0, 1);
          mma(tmem + (OVERLAP ? i * 2 * N + N : NT * 2 * N + i * N), ad + 1080, bd, id1, 1);
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    out[blockIdx.x] = (unsigned long long)(clock64() - t0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <int N, int NT, bool OVERLAP>
void run_split(unsigned long long* out, int random_data = 0) {
  const int iters = 200;
  cudaFuncSetAttribute(bench_split<N, NT, OVERLAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  bench_split<N, NT, OVERLAP><<<148, 128, 200 * 1024>>>(iters, out, random_data);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("variant failed: %s\n", cudaGetErrorString(e)); exit(1); }
  unsigned long long h[148];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  double sum = 0;
  for (int i = 0; i < 148; ++i) sum += h[i];
  printf("split pair N=%3d+%3d NT=%d %s %s | cycles per PAIR avg %.1f\n", 2 * N, N, NT, random_data ? "random fp16 data" : "zero data       ",
         OVERLAP ? "second MMA into the first one's upper half" : "second MMA into its own accumulator  ",
         sum / 148 / ((double)iters * 9 * NT));
}

template <int N, int NT, int PW, int M, int LBO16 = -1>
void run(unsigned long long* out) {
  const int iters = 200;
  cudaFuncSetAttribute(bench<N, NT, PW, M, LBO16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  bench<N, NT, PW, M, LBO16><<<148, 128, 200 * 1024>>>(iters, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("variant failed: %s\n", cudaGetErrorString(e)); exit(1); }
  unsigned long long h[148];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  double mx = 0, sum = 0;
  for (int i = 0; i < 148; ++i) { sum += h[i]; if (h[i] > mx) mx = (double)h[i]; }
  const double n = (double)iters * 9 * NT;
  printf("N=%3d NT=%2d PW=%2d M=%3d LBO=%3d | cycles/MMA avg %.1f max %.1f  (math floor %d)\n", N, NT, PW, M, LBO16,
         sum / 148 / n, mx / n, N / 2);
}

int main() {
  unsigned long long* out;
  cudaMalloc(&out, 148 * sizeof(unsigned long long));
  run<64, 1, 26, 128>(out);  run<64, 2, 26, 128>(out);  run<64, 3, 26, 128>(out);  run<64, 4, 26, 128>(out);
  run<64, 8, 26, 128>(out);  run<128, 1, 26, 128>(out); run<128, 2, 26, 128>(out); run<128, 4, 26, 128>(out);
  run<192, 1, 26, 128>(out); run<192, 2, 26, 128>(out); run<256, 1, 26, 128>(out); run<256, 2, 26, 128>(out);
  run<32, 8, 26, 128>(out);  run<16, 8, 26, 128>(out);  run<64, 3, 24, 128>(out);  run<64, 3, 32, 128>(out);
  run<64, 4, 26, 64>(out);   run<128, 2, 26, 64>(out);  run<128, 2, 18, 128>(out); run<128, 1, 10, 128>(out);
  // tap pairs on K (8-channel layers): second K half 1 / PW / 0 pixels away
  run<32, 4, 18, 128, 1>(out);  run<32, 4, 18, 128, 16>(out); run<32, 4, 18, 128, 0>(out);
  run<64, 4, 10, 128, 1>(out);  run<64, 4, 10, 128, 10>(out); run<64, 4, 10, 128, 0>(out);  run<64, 4, 10, 128>(out);
  run<32, 4, 10, 128, 1>(out);  run<32, 4, 10, 128, 10>(out); run<64, 4, 18, 128, 1>(out);  run<64, 4, 12, 128, 1>(out);
  run_split<32, 4, true>(out); run_split<32, 4, false>(out); run_split<16, 8, true>(out); run_split<16, 8, false>(out);
  run_split<64, 2, true>(out); run_split<64, 2, false>(out);
  run_split<32, 4, true>(out, 1); run_split<16, 8, true>(out, 1); run_split<64, 2, true>(out, 1);
  return 0;
}
