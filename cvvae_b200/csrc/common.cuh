// Shared host/device helpers of libcvvae_b200: error reporting, launch accounting, dtype conversion,
// cuTensorMapEncodeTiled access without linking libcuda.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/cvvae_b200.h"

namespace cvvae {

// ---- error plumbing (api.cu owns the storage)
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define CVVAE_CHECK_ARG(cond, ...)              \
  do {                                          \
    if (!(cond)) {                              \
      ::cvvae::set_error(__VA_ARGS__);          \
      return CVVAE_E_ARG;                       \
    }                                           \
  } while (0)

#define CVVAE_CUDA(expr)                                                                        \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      ::cvvae::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CVVAE_E_CUDA;                                                                      \
    }                                                                                           \
  } while (0)

#define CVVAE_LAUNCH_CHECK()                                                                    \
  do {                                                                                          \
    cudaError_t _e = cudaGetLastError();                                                        \
    if (_e != cudaSuccess) {                                                                    \
      ::cvvae::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CVVAE_E_CUDA;                                                                      \
    }                                                                                           \
    ::cvvae::count_launch();                                                                    \
  } while (0)

// ---- driver entry point for TMA descriptors
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_tiled();

// ---- 16-bit storage <-> fp32
template <int DT>
struct Elem;
template <>
struct Elem<CVVAE_F16> {
  using T = __half;
  using T2 = __half2;
  static __device__ __forceinline__ float to_f(T v) { return __half2float(v); }
  static __device__ __forceinline__ T from_f(float v) { return __float2half_rn(v); }
  static __device__ __forceinline__ float2 to_f2(uint32_t u) {
    return __half22float2(*reinterpret_cast<const __half2*>(&u));
  }
  static __device__ __forceinline__ uint32_t pack2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
};
template <>
struct Elem<CVVAE_BF16> {
  using T = __nv_bfloat16;
  using T2 = __nv_bfloat162;
  static __device__ __forceinline__ float to_f(T v) { return __bfloat162float(v); }
  static __device__ __forceinline__ T from_f(float v) { return __float2bfloat16_rn(v); }
  static __device__ __forceinline__ float2 to_f2(uint32_t u) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
  }
  static __device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
};

#define CVVAE_DISPATCH_DTYPE(dt, ...)                          \
  do {                                                         \
    if ((dt) == CVVAE_F16) {                                   \
      constexpr int DT = CVVAE_F16;                            \
      __VA_ARGS__                                              \
    } else if ((dt) == CVVAE_BF16) {                           \
      constexpr int DT = CVVAE_BF16;                           \
      __VA_ARGS__                                              \
    } else {                                                   \
      ::cvvae::set_error("unsupported dtype %d", (int)(dt));   \
      return CVVAE_E_ARG;                                      \
    }                                                          \
  } while (0)

// GroupNorm statistics are accumulated as 64-bit FIXED-POINT integers (sum * 2^20, sum of squares * 2^18) so that
// the many atomic contributions add up to the same bits in any order: results stay deterministic run to run.
// Range: |sum| < 8.8e12, sum^2 < 3.5e13 per (sample, group) - e.g. 22.6 M elements of rms magnitude up to 1.2e3.
constexpr double kGnSumScale = 1048576.0;   // 2^20
constexpr double kGnSqScale = 262144.0;     // 2^18
__device__ __forceinline__ unsigned long long gn_fix(float v, double scale) {
  return static_cast<unsigned long long>(__double2ll_rn(static_cast<double>(v) * scale));
}

// x * sigmoid(x); fast reciprocal (2 ulp) is far below the 16-bit output rounding
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

// ---- per-device state.  The library may be driven on several GPUs from one process (a model per device): everything
// cached about "the device" is keyed by the CURRENT device ordinal at the time of the call (the caller guarantees the
// current device is the one the pointers live on; cvvae_b200/ops.py enforces that).
constexpr int kMaxDevices = 64;
inline int cur_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
inline int num_sms() {
  static std::atomic<int> n[kMaxDevices];
  const int dev = cur_device();
  int v = n[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
    n[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}
// "has this one-time set-up (cudaFuncSetAttribute ...) been done on the current device?"  One static instance per call site.
struct PerDeviceOnce {
  std::atomic<bool> done[kMaxDevices];
  bool need() const { return !done[cur_device()].load(std::memory_order_acquire); }
  void mark() { done[cur_device()].store(true, std::memory_order_release); }
};

inline bool tensor_ok(const cvvae_tensor5* t) {
  return t && t->ptr && t->B > 0 && t->T > 0 && t->H > 0 && t->W > 0 && t->C > 0;
}

}  // namespace cvvae
