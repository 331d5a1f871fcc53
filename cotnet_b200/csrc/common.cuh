// Shared helpers for libcotb200 (sm_100a).  Host-side error plumbing + device-side element traits.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include <type_traits>

#include "../../include/cotb200.h"

namespace cotb200 {

// ---------------------------------------------------------------- host: errors + launch accounting
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

// Optional per-launch timing (cotb200_prof_enable): CUDA events recorded on the launch stream around each kernel.
// Disabled by default (zero overhead beyond one relaxed atomic load); never enable under graph capture.
struct ProfScope {
  const char* name; cudaStream_t st; cudaEvent_t e0; bool on; double bytes;
  ProfScope(const char* n, cudaStream_t s, double algorithmic_bytes = 0.0);
  ~ProfScope();
};
// `bytes` = ALGORITHMIC bytes of the launch: what the kernel must move, not what it happened to.
#define COTB200_PROF(name) cotb200::ProfScope _prof_scope(name, st)
#define COTB200_PROF_B(name, bytes) cotb200::ProfScope _prof_scope(name, st, (double)(bytes))

// Per-device memo slots: a process may drive several GPUs (the reference memoises its kernels per device too,
// cupy_layers/utils.py:14), so nothing that depends on the device is cached in a plain static.
inline int cur_dev_slot() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev & 63;
}
struct PerDevFlag {
  bool v[64] = {};
  bool& get() { return v[cur_dev_slot()]; }
};
inline int num_sms() {
  static int sms[64] = {};
  int& s = sms[cur_dev_slot()];
  if (!s) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&s, cudaDevAttrMultiProcessorCount, dev);
    if (s <= 0) s = 148;
  }
  return s;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---------------------------------------------------------------- device: element traits
template <typename T> struct Elem;
template <> struct Elem<float> {
  using Acc = float;
  __device__ __forceinline__ static float ld(const float* p) { return __ldg(p); }
  __device__ __forceinline__ static float from(float a) { return a; }
};
template <> struct Elem<double> {
  using Acc = double;
  __device__ __forceinline__ static double ld(const double* p) { return __ldg(p); }
  __device__ __forceinline__ static double from(double a) { return a; }
};
template <> struct Elem<__nv_bfloat16> {
  using Acc = float;
  __device__ __forceinline__ static float ld(const __nv_bfloat16* p) {
    return __bfloat162float(__ldg(p));
  }
  __device__ __forceinline__ static __nv_bfloat16 from(float a) { return __float2bfloat16_rn(a); }
};
template <> struct Elem<__half> {
  using Acc = float;
  __device__ __forceinline__ static float ld(const __half* p) { return __half2float(__ldg(p)); }
  __device__ __forceinline__ static __half from(float a) { return __float2half_rn(a); }
};

// A VEC-wide packet of T moved with one load/store instruction (VEC*sizeof(T) <= 16 bytes).
template <typename T, int VEC> struct alignas(sizeof(T) * VEC) Pack { T v[VEC]; };

template <typename T, int VEC>
__device__ __forceinline__ Pack<T, VEC> ld_pack(const T* p) {
  Pack<T, VEC> r;
  constexpr int B = sizeof(T) * VEC;
  if constexpr (B == 16) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    r = *reinterpret_cast<Pack<T, VEC>*>(&u);
  } else if constexpr (B == 8) {
    uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    r = *reinterpret_cast<Pack<T, VEC>*>(&u);
  } else if constexpr (B == 4) {
    unsigned u = __ldg(reinterpret_cast<const unsigned*>(p));
    r = *reinterpret_cast<Pack<T, VEC>*>(&u);
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) r.v[i] = __ldg(p + i);
  }
  return r;
}

template <typename T, int VEC>
__device__ __forceinline__ void st_pack(T* p, const Pack<T, VEC>& r) {
  constexpr int B = sizeof(T) * VEC;
  if constexpr (B == 16) {
    *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(&r);
  } else if constexpr (B == 8) {
    *reinterpret_cast<uint2*>(p) = *reinterpret_cast<const uint2*>(&r);
  } else if constexpr (B == 4) {
    *reinterpret_cast<unsigned*>(p) = *reinterpret_cast<const unsigned*>(&r);
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) p[i] = r.v[i];
  }
}

template <typename T> __device__ __forceinline__ typename Elem<T>::Acc to_acc(T v);
template <> __device__ __forceinline__ float to_acc<float>(float v) { return v; }
template <> __device__ __forceinline__ double to_acc<double>(double v) { return v; }
template <> __device__ __forceinline__ float to_acc<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_acc<__half>(__half v) { return __half2float(v); }

// d = a*b + c with 16-bit a, b consumed directly (sm_100 mixed-precision FMA, SASS FHFMA with .H0/.H1 operand selectors):
// no unpack instructions in reductions over bf16/fp16 tensors.  Exact products, fp32 accumulation.
template <typename T> __device__ __forceinline__ float mfma(T a, T b, float c);
template <> __device__ __forceinline__ float mfma<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <> __device__ __forceinline__ float mfma<__nv_bfloat16>(__nv_bfloat16 a, __nv_bfloat16 b, float c) {
  float d;
  asm("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(d) : "h"(__bfloat16_as_ushort(a)), "h"(__bfloat16_as_ushort(b)), "f"(c));
  return d;
}
template <> __device__ __forceinline__ float mfma<__half>(__half a, __half b, float c) {
  float d;
  asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"(__half_as_ushort(a)), "h"(__half_as_ushort(b)), "f"(c));
  return d;
}
template <typename T> __device__ __forceinline__ T one_of();
template <> __device__ __forceinline__ float one_of<float>() { return 1.f; }
template <> __device__ __forceinline__ __nv_bfloat16 one_of<__nv_bfloat16>() { return __ushort_as_bfloat16((unsigned short)0x3F80); }
template <> __device__ __forceinline__ __half one_of<__half>() { return __ushort_as_half((unsigned short)0x3C00); }

// dtype dispatch on the host
#define COTB200_DISPATCH_DTYPE(dtype, ...)                                   \
  switch (dtype) {                                                           \
    case COTB200_F32:  { using T = float;          __VA_ARGS__; } break;     \
    case COTB200_F64:  { using T = double;         __VA_ARGS__; } break;     \
    case COTB200_BF16: { using T = __nv_bfloat16;  __VA_ARGS__; } break;     \
    case COTB200_F16:  { using T = __half;         __VA_ARGS__; } break;     \
    default: cotb200::set_error("unknown dtype %d", (int)(dtype)); return COTB200_EDTYPE; \
  }

}  // namespace cotb200
