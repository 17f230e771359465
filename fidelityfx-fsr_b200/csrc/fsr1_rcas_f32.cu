// fsr1_rcas_f32.cu — RCAS for RGBA32F images (the SAMPLE_SLOW_FALLBACK precision) organised like the packed
// half kernel: a lane owns two adjacent pixels and walks kRows rows with every row's 2 x 128-bit loads issued
// up front; left/right taps come from the neighbouring lanes by warp shuffle; a warp loads a 64-pixel span and
// produces its inner 60, so there are no edge fetches; warps strictly inside the image skip all bounds checks.
// Arithmetic is the F path (ffx-fsr/ffx_fsr1.h:684-769) in fp32 with FMA contraction (tolerance 1e-5 against the
// oracle; FSR1_FLAG_EXACT selects the bit-exact direct kernel instead).  32 bytes of compulsory traffic per pixel.
#include "fsr1_common.cuh"

namespace fsr1 {

constexpr int kFRows = 4, kFWarps = 8, kFSpan = 60;

struct PxF { float r, g, b, a; };
struct PairF { PxF a, b; };  // pixels x, x+1

template <bool kChecked>
__device__ __forceinline__ PxF load_one(const RcasParams& p, int x, int y) {
  if (kChecked) {
    if (p.clamp) { x = clampi(x, 0, p.in.w - 1); y = clampi(y, 0, p.in.h - 1); }
    else if (x < 0 || y < 0 || x >= p.in.w || y >= p.in.h) return PxF{0.f, 0.f, 0.f, 0.f};
    if (!row_stored(p.in, y)) return PxF{0.f, 0.f, 0.f, 0.f};  // prefetched past the row range of a window: never used
  }
  const float4 v = __ldg(reinterpret_cast<const float4*>(p.in.base + (long long)(y - p.in.row0) * p.in.pitch) + x);
  return PxF{v.x, v.y, v.z, v.w};
}
template <bool kChecked> __device__ __forceinline__ PairF load_two(const RcasParams& p, int x, int y) {
  return PairF{load_one<kChecked>(p, x, y), load_one<kChecked>(p, x + 1, y)};
}

// Reciprocals: MUFU.RCP (rcp.approx.f32, 1 ulp, denormals handled) instead of the IEEE-rounded __frcp_rn, whose
// refinement and slow-path call were a quarter of this kernel's instructions (measured on B200: 65.1 -> 52.6 us per 4K
// frame, profiles/r02_variants.log); the fast fp32 path is held to 1e-5 of the oracle, not to bit-exactness (that is
// FSR1_FLAG_EXACT).
__device__ __forceinline__ float rcp_f(float a) {
#ifdef FSR1_CPU_EMU
  return 1.0f / a;
#else
  float r;
  asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
#endif
}

__device__ __forceinline__ float lobe_f(float b, float d, float e, float f, float h) {
  const float mn4 = fminf(fminf(b, d), fminf(f, h)), mx4 = fmaxf(fmaxf(b, d), fmaxf(f, h));
  const float hitMin = fminf(mn4, e) * rcp_f(4.0f * mx4);
  const float hitMax = (1.0f - fmaxf(mx4, e)) * rcp_f(fmaf(4.0f, mn4, -4.0f));
  return fmaxf(-hitMin, hitMax);
}

__device__ __forceinline__ float luma2x_f(PxF c) { return fmaf(c.b, 0.5f, fmaf(c.r, 0.5f, c.g)); }  // ffx_fsr1.h:725-729

// kOpt: the reference's compile-time options as template bits (enum kRcas* in fsr1_rcas_math.cuh): 1 FSR_RCAS_DENOISE
// (ffx_fsr1.h:731-739,761-763), 2 FSR_RCAS_PASSTHROUGH_ALPHA (:688-702), 4 the Sample.x output square (FSR_Pass.hlsl:93-94)
template <int kOpt>
__device__ __forceinline__ void rcas_px(const RcasParams& p, PxF b, PxF d, PxF e, PxF f, PxF h, float4& out) {
  float lobe = fmaxf(-0.1875f, fminf(fmaxf(lobe_f(b.r, d.r, e.r, f.r, h.r),
                                           fmaxf(lobe_f(b.g, d.g, e.g, f.g, h.g), lobe_f(b.b, d.b, e.b, f.b, h.b))), 0.0f)) * p.sharp;
  if (kOpt & 1) {
    const float bL = luma2x_f(b), dL = luma2x_f(d), eL = luma2x_f(e), fL = luma2x_f(f), hL = luma2x_f(h);
    float nz = fmaf(0.25f, hL, fmaf(0.25f, fL, fmaf(0.25f, dL, 0.25f * bL))) - eL;
    const float range = fmaxf(fmaxf(fmaxf(bL, dL), eL), fmaxf(fL, hL)) - fminf(fminf(fminf(bL, dL), eL), fminf(fL, hL));
    const float sr = __uint_as_float(0x7ef19fffu - __float_as_uint(range));  // APrxMedRcpF1
    nz = sat(fabsf(nz) * (sr * fmaf(-sr, range, 2.0f)));
    lobe *= fmaf(-0.5f, nz, 1.0f);
  }
  const float a = fmaf(4.0f, lobe, 1.0f);
  const float s = __uint_as_float(0x7ef19fffu - __float_as_uint(a));  // APrxMedRcpF1 (ffx_a.h:1844)
  const float rcpL = s * fmaf(-s, a, 2.0f);
  out.x = fmaf(lobe, f.r, fmaf(lobe, h.r, fmaf(lobe, d.r, lobe * b.r))) + e.r;
  out.y = fmaf(lobe, f.g, fmaf(lobe, h.g, fmaf(lobe, d.g, lobe * b.g))) + e.g;
  out.z = fmaf(lobe, f.b, fmaf(lobe, h.b, fmaf(lobe, d.b, lobe * b.b))) + e.b;
  out.x *= rcpL; out.y *= rcpL; out.z *= rcpL;
  if (kOpt & 4) { out.x *= out.x; out.y *= out.y; out.z *= out.z; }
  out.w = (kOpt & 2) ? e.a : 1.0f;
}

__device__ __forceinline__ PxF shfl_px(PxF v, int delta_up) {
  PxF o;
  o.a = 0.f;
  if (delta_up) {
    o.r = __shfl_up_sync(0xffffffffu, v.r, 1); o.g = __shfl_up_sync(0xffffffffu, v.g, 1); o.b = __shfl_up_sync(0xffffffffu, v.b, 1);
  } else {
    o.r = __shfl_down_sync(0xffffffffu, v.r, 1); o.g = __shfl_down_sync(0xffffffffu, v.g, 1); o.b = __shfl_down_sync(0xffffffffu, v.b, 1);
  }
  return o;
}

template <bool kChecked, int kOpt>
__device__ __forceinline__ void rcas_rows_f32(const RcasParams& p, int x, int ys, int lane) {
  const bool writer = lane >= 1 && lane <= 30 && (!kChecked || x < p.out.w);
  PairF rows[kFRows + 2];
#pragma unroll
  for (int r = 0; r < kFRows + 2; r++) rows[r] = load_two<kChecked>(p, x, ys - 1 + r);
#pragma unroll
  for (int r = 0; r < kFRows; r++) {
    const int y = ys + r;
    if (kChecked && y >= p.y1) break;
    const PairF prev = rows[r], cur = rows[r + 1], next = rows[r + 2];
    const PxF left = shfl_px(cur.b, 1), right = shfl_px(cur.a, 0);  // left lane's pixel1, right lane's pixel0
    float4 o0, o1;
    rcas_px<kOpt>(p, prev.a, left, cur.a, cur.b, next.a, o0);
    rcas_px<kOpt>(p, prev.b, cur.a, cur.b, right, next.b, o1);
    if (writer) {
      float4* o = reinterpret_cast<float4*>(p.out.base + (long long)(y - p.out.row0) * p.out.pitch) + x;
      o[0] = o0;
      if (!kChecked || x + 1 < p.out.w) o[1] = o1;
    }
  }
}

template <int kOpt>
__global__ void __launch_bounds__(32 * kFWarps) rcas_f32_packed_kernel(const RcasParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x0 = blockIdx.x * kFSpan - 2, x = x0 + lane * 2;
  const int ys = p.y0 + (blockIdx.y * kFWarps + warp) * kFRows;
  if (ys >= p.y1) return;
  const bool interior = x0 >= 0 && x0 + 64 <= p.in.w && ys >= 1 && ys + kFRows < p.in.h && ys + kFRows <= p.y1;
  if (interior) rcas_rows_f32<false, kOpt>(p, x, ys, lane);
  else rcas_rows_f32<true, kOpt>(p, x, ys, lane);
}

#ifndef FSR1_CPU_EMU
cudaError_t launch_rcas_f32_packed(const RcasParams& p, cudaStream_t s, const char** name) {
  if ((reinterpret_cast<uintptr_t>(p.in.base) & 15) || (p.in.pitch & 15) || (reinterpret_cast<uintptr_t>(p.out.base) & 15) ||
      (p.out.pitch & 15))
    return cudaErrorNotSupported;
  const dim3 grid((p.out.w + kFSpan - 1) / kFSpan, (p.y1 - p.y0 + kFWarps * kFRows - 1) / (kFWarps * kFRows), 1);
  switch (p.options & 7) {
    case 0: rcas_f32_packed_kernel<0><<<grid, 32 * kFWarps, 0, s>>>(p); break;
    case 1: rcas_f32_packed_kernel<1><<<grid, 32 * kFWarps, 0, s>>>(p); break;
    case 2: rcas_f32_packed_kernel<2><<<grid, 32 * kFWarps, 0, s>>>(p); break;
    case 3: rcas_f32_packed_kernel<3><<<grid, 32 * kFWarps, 0, s>>>(p); break;
    case 4: rcas_f32_packed_kernel<4><<<grid, 32 * kFWarps, 0, s>>>(p); break;
    case 5: rcas_f32_packed_kernel<5><<<grid, 32 * kFWarps, 0, s>>>(p); break;
    case 6: rcas_f32_packed_kernel<6><<<grid, 32 * kFWarps, 0, s>>>(p); break;
    default: rcas_f32_packed_kernel<7><<<grid, 32 * kFWarps, 0, s>>>(p); break;
  }
  *name = "rcas_f32_packed<2px,4rows,shfl60>";
  return cudaGetLastError();
}

#endif  // FSR1_CPU_EMU

}  // namespace fsr1
