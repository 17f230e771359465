// fsr1_easu_common.cuh — pieces shared by the tiled EASU kernels (fp16 and fp32 storage): PTX wrappers for
// mbarrier + TMA, the per-texel and per-pixel fp32 analysis, and the host-side tensor-map helpers.
#pragma once
#include <cuda.h>
#include <stdlib.h>
#include "fsr1_common.cuh"

namespace fsr1 {

#ifdef FSR1_CPU_EMU  // tests/emu: this device code compiled for the host; the emulator supplies the PTX wrappers
}  // namespace fsr1
#include "fsr1_emu_ptx.h"
namespace fsr1 {
#else
// ---- PTX wrappers: mbarrier + TMA ------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LAB_DONE;\n"
      "bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}

#endif  // FSR1_CPU_EMU

// ---- small helpers ------------------------------------------------------------------------------------
__device__ __forceinline__ __half2 u2h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t h22u(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 h2c(float v) { return __float2half2_rn(v); }
#ifndef FSR1_CPU_EMU
__device__ __forceinline__ float rcp_approx(float a) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
#endif

// FsrEasuSetF without the bilinear weight: (dirX, dirY, lenX^2 + lenY^2) of the texel whose luma is lC.
__device__ __forceinline__ float4 texel_terms(float lA, float lB, float lC, float lD, float lE) {
  const float dirX = lD - lB, dirY = lE - lA;
  const float lenX = sat(fabsf(dirX) * prx_lo_rcp(fmaxf(fabsf(lD - lC), fabsf(lC - lB))));
  const float lenY = sat(fabsf(dirY) * prx_lo_rcp(fmaxf(fabsf(lE - lC), fabsf(lC - lA))));
  return make_float4(dirX, dirY, fmaf(lenX, lenX, lenY * lenY), 0.0f);
}

#ifndef FSR1_CPU_EMU
// ---- host side ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn resolve_encode_fn() {
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
      qres == cudaDriverEntryPointSuccess)
    return reinterpret_cast<EncodeTiledFn>(sym);
  return nullptr;
}
static inline EncodeTiledFn get_encode_fn() {
  static const EncodeTiledFn fn = resolve_encode_fn();  // thread-safe one-time initialisation
  return fn;
}

// Same float arithmetic as easu_pos on the device.
static inline int host_fp(int o, float scale, float offset) {
  volatile float m = (float)o * scale;
  volatile float s = m + offset;
  return (int)floorf(s);
}

static inline int sm_count() {  // of the CURRENT device (one process may drive several); queried once per device
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = cache[dev];
  if (n <= 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev] = n;  // benign race: every thread writes the same value
  }
  return n;
}

#endif  // FSR1_CPU_EMU

}  // namespace fsr1
