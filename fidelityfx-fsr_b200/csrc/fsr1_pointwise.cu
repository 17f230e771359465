// fsr1_pointwise.cu — the pointwise companions of the scaling path: the steps the sample runs
// directly before EASU and after RCAS, as image-level passes on sm_100a.
//
//   SRTM / SRTM-inverse   FsrSrtmF / FsrSrtmInvF      ffx-fsr/ffx_fsr1.h:1044,1046   reversible tone-mapper around the filter
//   LFGA                  FsrLfgaF                     ffx-fsr/ffx_fsr1.h:1014        film grain after scaling
//   TEPD 8 / 10 bit       FsrTepdC8F / FsrTepdC10F     ffx-fsr/ffx_fsr1.h:1100-1126   dithered linear -> gamma 2.0
//                         FsrTepdDitF                  ffx-fsr/ffx_fsr1.h:1086-1095   positional dither value
//   SQUARE                `c *= c`                     sample/src/DX12/FSR_Pass.hlsl:78-79,84-85,93-94,99-100
//                                                      (the Sample.x hook: gamma 2.0 back to linear on the last pass)
//
// These are streaming passes: 2 x bytes-per-pixel of compulsory traffic and a few dozen flops per pixel, i.e.
// HBM-bound by a wide margin.  One thread = kRowsPerThread pixels of ONE row, 256 apart (a CTA covers a contiguous
// 1024-pixel stretch of a row): the loads of all of them are issued before any arithmetic (memory-level
// parallelism), every warp access is one fully coalesced 128/256/512-byte segment, and there is no shared memory.  Because the arithmetic is free here, it is always the
// EXACT policy: separate roundings, IEEE sqrt and division — the fp32 results are bit-identical to the reference
// source compiled with -ffp-contract=off, for every storage format (fp16/unorm storage rounds that fp32 result once).
// Alpha is carried through unchanged (the reference functions take RGB).
#include <stdlib.h>
#include "fsr1_common.cuh"

namespace fsr1 {

enum { kOpSrtm = 1, kOpSrtmInv = 2, kOpLfga = 3, kOpTepd8 = 4, kOpTepd10 = 5, kOpSquare = 6 };

constexpr int kRowsPerThread = 4;
constexpr int kPointThreads = 256;

template <typename S> __device__ __forceinline__ float4 load4(const ImgView& im, int x, int y);
template <> __device__ __forceinline__ float4 load4<float>(const ImgView& im, int x, int y) {
  return __ldg(reinterpret_cast<const float4*>(im.base + (long long)(y - im.row0) * im.pitch) + x);
}
template <> __device__ __forceinline__ float4 load4<__half>(const ImgView& im, int x, int y) {
  const uint2 v = __ldg(reinterpret_cast<const uint2*>(im.base + (long long)(y - im.row0) * im.pitch) + x);
  const float2 rg = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
  const float2 ba = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
  return make_float4(rg.x, rg.y, ba.x, ba.y);
}
template <> __device__ __forceinline__ float4 load4<Unorm8>(const ImgView& im, int x, int y) {
  const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(im.base + (long long)(y - im.row0) * im.pitch) + x);
  return make_float4(__fdiv_rn((float)(v & 255u), 255.0f), __fdiv_rn((float)((v >> 8) & 255u), 255.0f),
                     __fdiv_rn((float)((v >> 16) & 255u), 255.0f), __fdiv_rn((float)(v >> 24), 255.0f));
}
template <> __device__ __forceinline__ float4 load4<Unorm10>(const ImgView& im, int x, int y) {
  const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(im.base + (long long)(y - im.row0) * im.pitch) + x);
  return make_float4(__fdiv_rn((float)(v & 1023u), 1023.0f), __fdiv_rn((float)((v >> 10) & 1023u), 1023.0f),
                     __fdiv_rn((float)((v >> 20) & 1023u), 1023.0f), __fdiv_rn((float)(v >> 30), 3.0f));
}

// aux tiles (grain, dither) are small and L1/L2-resident: runtime format; (x, y) already wrapped into the tile
__device__ __forceinline__ float4 load_aux(const ImgView& im, int fmt, int x, int y) {
  switch (fmt) {
    case 1: return load4<__half>(im, x, y);
    case 2: return load4<float>(im, x, y);
    case 3: return load4<Unorm8>(im, x, y);
    default: return load4<Unorm10>(im, x, y);
  }
}

struct PointParams {
  ImgView in, out, aux;
  int aux_format;  // 0: no aux image
  int op;
  float amount;    // LFGA
  uint32_t frame;  // TEPD positional dither
  int y0, y1;
};

// APrxMedRcpF1 (ffx-fsr/ffx_a.h:1844), separate roundings; the integer subtract wraps for negative arguments
__device__ __forceinline__ float prx_med_rcp_exact(float a) {
  const float b = __uint_as_float(0x7ef19fffu - __float_as_uint(a));
  return __fmul_rn(b, __fadd_rn(__fmul_rn(-b, a), 2.0f));
}

__device__ __forceinline__ float tepd_dit(uint32_t px, uint32_t py, uint32_t frame) {
  const float x = (float)(px + frame), y = (float)py;
  const float a = 1.61803398874989484820f, b = (float)(1.0 / 3.69);
  const float v = __fadd_rn(__fmul_rn(x, a), __fmul_rn(y, b));
  return __fsub_rn(v, floorf(v));
}

__device__ __forceinline__ float tepd_channel(float c, float dit, float q, float rq) {
  float n = __fsqrt_rn(c);
  n = __fmul_rn(floorf(__fmul_rn(n, q)), rq);
  const float a = __fmul_rn(n, n);
  float b = __fadd_rn(n, rq);
  b = __fmul_rn(b, b);
  const float r = __fmul_rn(__fsub_rn(c, b), prx_med_rcp_exact(__fsub_rn(a, b)));
  // AGtZeroF1(m) = saturate(m * +INF): 1 for m > 0, else 0 (0 * INF = NaN saturates to 0)
  const float gt = sat(__fmul_rn(__fsub_rn(dit, r), __uint_as_float(0x7f800000u)));
  return sat(__fadd_rn(n, __fmul_rn(gt, rq)));
}

// (ax, ay) = (x mod aux width, y mod aux height), maintained by the caller (one division per thread, not per pixel)
__device__ __forceinline__ float4 apply_op(const PointParams& p, float4 c, int x, int y, int ax, int ay) {
  switch (p.op) {
    case kOpSrtm: {
      const float r = __fdiv_rn(1.0f, __fadd_rn(fmaxf(c.x, fmaxf(c.y, c.z)), 1.0f));
      return make_float4(__fmul_rn(c.x, r), __fmul_rn(c.y, r), __fmul_rn(c.z, r), c.w);
    }
    case kOpSrtmInv: {
      const float r = __fdiv_rn(1.0f, fmaxf((float)(1.0 / 32768.0), __fsub_rn(1.0f, fmaxf(c.x, fmaxf(c.y, c.z)))));
      return make_float4(__fmul_rn(c.x, r), __fmul_rn(c.y, r), __fmul_rn(c.z, r), c.w);
    }
    case kOpLfga: {
      const float4 t = load_aux(p.aux, p.aux_format, ax, ay);
      const float a = p.amount;
      return make_float4(__fadd_rn(c.x, __fmul_rn(__fmul_rn(t.x, a), fminf(__fsub_rn(1.0f, c.x), c.x))),
                         __fadd_rn(c.y, __fmul_rn(__fmul_rn(t.y, a), fminf(__fsub_rn(1.0f, c.y), c.y))),
                         __fadd_rn(c.z, __fmul_rn(__fmul_rn(t.z, a), fminf(__fsub_rn(1.0f, c.z), c.z))), c.w);
    }
    case kOpTepd8:
    case kOpTepd10: {
      const float q = p.op == kOpTepd8 ? 255.0f : 1023.0f;
      const float rq = p.op == kOpTepd8 ? (float)(1.0 / 255.0) : (float)(1.0 / 1023.0);
      const float dit = p.aux_format ? sat(load_aux(p.aux, p.aux_format, ax, ay).w) : tepd_dit((uint32_t)x, (uint32_t)y, p.frame);
      return make_float4(tepd_channel(c.x, dit, q, rq), tepd_channel(c.y, dit, q, rq), tepd_channel(c.z, dit, q, rq), c.w);
    }
    default:  // kOpSquare
      return make_float4(__fmul_rn(c.x, c.x), __fmul_rn(c.y, c.y), __fmul_rn(c.z, c.z), c.w);
  }
}

// kCols = false: a thread owns one column position and kRowsPerThread rows (CTA = 256 x 4 pixels);
// kCols = true : a thread owns kRowsPerThread positions 256 apart in ONE row (CTA = 1024 x 1 pixels: one contiguous
//               8-16 KB stretch of a row per CTA).  Same arithmetic, different DRAM access shape (FSR1_POINT_LAYOUT).
template <typename SI, typename SO, bool kCols, int N>
__global__ void __launch_bounds__(kPointThreads) pointwise_kernel(const PointParams p, const int aux_step) {
  int xs[N], ys[N];
#pragma unroll
  for (int r = 0; r < N; r++) {
    xs[r] = kCols ? blockIdx.x * (kPointThreads * N) + threadIdx.x + kPointThreads * r : blockIdx.x * kPointThreads + threadIdx.x;
    ys[r] = kCols ? p.y0 + (int)blockIdx.y : p.y0 + (int)blockIdx.y * N + r;
  }
  float4 c[N];
#pragma unroll
  for (int r = 0; r < N; r++)
    if (xs[r] < p.out.w && ys[r] < p.y1) c[r] = load4<SI>(p.in, xs[r], ys[r]);
  // position inside the aux tile: one division per thread, then incremental with wrap
  int ax = 0, ay = 0;
  if (p.aux_format) {
    ax = xs[0] % p.aux.w;
    ay = ys[0] % p.aux.h;
  }
#pragma unroll
  for (int r = 0; r < N; r++) {
    if (xs[r] < p.out.w && ys[r] < p.y1) {
      const float4 o = apply_op(p, c[r], xs[r], ys[r], ax, ay);
      Px<SO>::store(p.out, xs[r], ys[r], o.x, o.y, o.z, o.w);
    }
    if (kCols) {
      ax += aux_step;  // aux_step = 256 mod aux width (host)
      if (ax >= p.aux.w) ax -= p.aux.w;
    } else if (++ay >= p.aux.h) {
      ay = 0;
    }
  }
}

template <typename SI, typename SO>
static cudaError_t launch_one(const PointParams& p, cudaStream_t s) {
  // a CTA covers a contiguous stretch of one row (measured: +8 % over a 256x4-pixel CTA on RGBA16F; 8 pixels per thread
  // is slower: registers)
  const int aux_step = p.aux_format ? kPointThreads % p.aux.w : 0;
  const int per_cta = kPointThreads * kRowsPerThread;
  const dim3 grid((p.out.w + per_cta - 1) / per_cta, p.y1 - p.y0, 1);
  pointwise_kernel<SI, SO, true, kRowsPerThread><<<grid, kPointThreads, 0, s>>>(p, aux_step);
  return cudaGetLastError();
}

// in_format == out_format for every op; TEPD may also write its 8/10-bit code values straight into a UNORM image
// from a float image (the conversion it exists for).
cudaError_t launch_pointwise(int op, const ImgView& in, int in_format, const ImgView& out, int out_format, const ImgView* aux,
                             int aux_format, float amount, uint32_t frame, int y0, int y1, cudaStream_t s, const char** name) {
  PointParams p;
  p.in = in;
  p.out = out;
  p.aux_format = aux ? aux_format : 0;
  if (aux) p.aux = *aux; else p.aux = in;
  p.op = op;
  p.amount = amount;
  p.frame = frame;
  p.y0 = y0;
  p.y1 = y1;
  static const char* const names[] = {"", "pointwise<srtm>", "pointwise<srtm_inv>", "pointwise<lfga>", "pointwise<tepd8>",
                                      "pointwise<tepd10>", "pointwise<square>"};
  if (op < kOpSrtm || op > kOpSquare) return cudaErrorInvalidValue;
  *name = names[op];
  if (in_format == out_format) {
    switch (in_format) {
      case 1: return launch_one<__half, __half>(p, s);
      case 2: return launch_one<float, float>(p, s);
      case 3: return launch_one<Unorm8, Unorm8>(p, s);
      case 4: return launch_one<Unorm10, Unorm10>(p, s);
    }
    return cudaErrorNotSupported;
  }
  if (op == kOpTepd8 && out_format == 3) {
    if (in_format == 1) return launch_one<__half, Unorm8>(p, s);
    if (in_format == 2) return launch_one<float, Unorm8>(p, s);
  }
  if (op == kOpTepd10 && out_format == 4) {
    if (in_format == 1) return launch_one<__half, Unorm10>(p, s);
    if (in_format == 2) return launch_one<float, Unorm10>(p, s);
  }
  return cudaErrorNotSupported;
}

}  // namespace fsr1
