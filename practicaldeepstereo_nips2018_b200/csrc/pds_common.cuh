// Shared helpers of the sm_100a kernel library (device + host side).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pds_b200.h"

namespace pds {

// Thread-local error message returned by pds_last_error().
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define PDS_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      ::pds::set_error(__VA_ARGS__);             \
      return PDS_ERR_INVALID_ARGUMENT;           \
    }                                            \
  } while (0)

#define PDS_CUDA(call)                                          \
  do {                                                          \
    cudaError_t e__ = (call);                                   \
    if (e__ != cudaSuccess) return ::pds::cuda_fail(e__, #call); \
  } while (0)

#define PDS_LAUNCH_CHECK(name)                                     \
  do {                                                             \
    cudaError_t e__ = cudaGetLastError();                          \
    if (e__ != cudaSuccess) return ::pds::cuda_fail(e__, name);    \
  } while (0)

// Launch accounting (pds_launch_count) and the optional per-kernel CUDA-event
// profiler (pds_profiler_*): every kernel launch in the library goes through a
// KernelScope so bench.py can count launches and time each kernel class live.
struct KernelScope {
  KernelScope(const char* name, cudaStream_t st);
  ~KernelScope();
  // algorithmic work of this launch (reference FLOPs / minimum HBM bytes), summed per kernel class
  void work(double flops, double bytes) { flops_ = flops; bytes_ = bytes; }
  const char* name_;
  cudaStream_t st_;
  cudaEvent_t start_ = nullptr;
  double flops_ = 0.0, bytes_ = 0.0;
};
#define PDS_KERNEL(name, st) ::pds::KernelScope pds_kernel_scope__(name, st)
#define PDS_KERNEL_WORK(flops, bytes) pds_kernel_scope__.work((double)(flops), (double)(bytes))

// Programmatic dependent launch: kernels launched through launch_pdl may START (prologue: barrier
// initialisation, TMEM allocation, loads of constant data such as weights) while the previous
// kernel of the stream is still draining; they must call pdl_wait() before touching any global
// memory the stream's earlier work produces or still reads, and call pdl_trigger() early so that
// their own successor can be scheduled as their CTAs retire.  The attribute is OFF by default
// (PDS_B200_PDL=1 enables it; measured slightly slower at C2); the device-side calls are then no-ops.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: set once per (kernel,
// device), so that one process may drive several GPUs.
cudaError_t allow_dynamic_smem_impl(const void* kernel, int bytes);
template <typename F>
inline cudaError_t allow_dynamic_smem(F kernel, int bytes) {
  return allow_dynamic_smem_impl(reinterpret_cast<const void*>(kernel), bytes);
}

// Streaming 128-bit accesses that do not pollute L1 (data touched once).
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ldg_stream(const uint2* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];"
               : "=r"(r.x), "=r"(r.y)
               : "l"(p));
  return r;
}
// 256-bit accesses (sm_100: LDG/STG.256): one full 32-byte sector per lane, for 8-float records
// (32-byte aligned)
__device__ __forceinline__ void stg256(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
               "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void ldg256_stream(const float* p, float* v) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p),
               "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

}  // namespace pds
