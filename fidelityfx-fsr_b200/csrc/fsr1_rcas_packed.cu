// fsr1_rcas_packed.cu — the production RCAS kernel for RGBA16F images on sm_100a.
//
// RCAS is a 5-tap cross (b above, d left, e centre, f right, h below; ffx-fsr/ffx_fsr1.h:693-707) with 16
// bytes of compulsory traffic per pixel: the HBM-bound half of the path.  The kernel is organised around
// the memory system and the issue slots, not shared memory:
//   * a lane owns TWO horizontally adjacent pixels (one 128-bit load and one 128-bit store per row) and
//     walks kRows rows downwards with a rolling 3-row register window;
//   * left/right neighbours (d, f) come from the adjacent lanes by warp shuffle.  A warp loads a 64-pixel
//     span but produces only its inner 60 pixels: lanes 0 and 31 exist to feed their neighbours, so there is
//     no per-row edge fetch and no divergence (spans overlap by 4 pixels; the re-read hits L1/L2);
//   * warps whose span and rows lie strictly inside the image take a path with no bounds checks at all;
//   * arithmetic is half2 over the lane's two pixels, structure-of-arrays like the reference's FsrRcasHx2
//     (ffx_fsr1.h:888-984): (R0,R1) (G0,G1) (B0,B1).
// Numerics: the six "high precision" reciprocals (ffx_fsr1.h:750-755) are rcp.approx.f32 on the unpacked
// halves (h2rcp, MUFU; a packed Newton iteration on the fp16 pipe measured slower); the resolve reciprocal is the packed APrxMedRcpH2 (ffx_a.h:1815).  Against the fp32
// oracle on the same half input: <= 2e-3 (tolerance 1e-2).  min/max are the non-propagating half2 forms, so
// the 0*inf NaNs of flat black / white neighbourhoods drop out exactly as with HLSL min/max (:756-759).
#include <stdlib.h>
#include "fsr1_common.cuh"

namespace fsr1 {

constexpr int kNW = 4;     // warps per CTA, stacked vertically: CTA = 60 x (kNW*kRows) output pixels
constexpr int kRows = 4;   // rows walked by one lane
constexpr int kSpan = 60;   // output pixels per warp per row (lanes 1..30)

struct Row3 { __half2 r, g, b; };  // (pixel0, pixel1) per channel

__device__ __forceinline__ __half2 uh2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t hu2(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

// AoS (RG0,BA0,RG1,BA1) -> SoA
__device__ __forceinline__ Row3 to_soa(uint4 v) {
  Row3 o;
  o.r = uh2(__byte_perm(v.x, v.z, 0x5410));
  o.g = uh2(__byte_perm(v.x, v.z, 0x7632));
  o.b = uh2(__byte_perm(v.y, v.w, 0x5410));
  return o;
}

// Pixels (x, x+1) of logical row y.  kChecked applies the out-of-image rule (0, or clamp with kClamp).
template <bool kChecked, bool kClamp>
__device__ __forceinline__ Row3 load_pair(const RcasParams& p, int x, int y) {
  if (!kChecked) {
    return to_soa(__ldg(reinterpret_cast<const uint4*>(p.in.base + (long long)(y - p.in.row0) * p.in.pitch + (long long)x * 8)));
  }
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (kClamp) y = clampi(y, 0, p.in.h - 1);
  if (row_stored(p.in, y)) {  // rows outside the stored window are prefetched past the row range, never used
    const uint2* row = reinterpret_cast<const uint2*>(p.in.base + (long long)(y - p.in.row0) * p.in.pitch);
    if (x >= 0 && x + 1 < p.in.w) {
      v = __ldg(reinterpret_cast<const uint4*>(row + x));
    } else if (kClamp) {
      const uint2 t0 = __ldg(row + clampi(x, 0, p.in.w - 1)), t1 = __ldg(row + clampi(x + 1, 0, p.in.w - 1));
      v = make_uint4(t0.x, t0.y, t1.x, t1.y);
    } else {
      if (x >= 0 && x < p.in.w) { const uint2 t = __ldg(row + x); v.x = t.x; v.y = t.y; }
      if (x + 1 >= 0 && x + 1 < p.in.w) { const uint2 t = __ldg(row + x + 1); v.z = t.x; v.w = t.y; }
    }
  }
  return to_soa(v);
}

// lobe of one channel for two pixels:  max(-hitMin, hitMax) = -min( min(mn4,e)/(4 mx4), (1-max(mx4,e))/(4-4 mn4) )
__device__ __forceinline__ __half2 lobe_channel(__half2 b, __half2 d, __half2 e, __half2 f, __half2 h) {
  const __half2 mn4 = __hmin2(__hmin2(b, d), __hmin2(f, h));
  const __half2 mx4 = __hmax2(__hmax2(b, d), __hmax2(f, h));
  const __half2 k4 = __float2half2_rn(4.0f), k1 = __float2half2_rn(1.0f), km4 = __float2half2_rn(-4.0f);
  const __half2 hitMin = __hmul2(__hmin2(mn4, e), h2rcp(__hmul2(k4, mx4)));
  const __half2 negHitMax = __hmul2(__hsub2(k1, __hmax2(mx4, e)), h2rcp(__hfma2(km4, mn4, k4)));
  return __hneg2(__hmin2(hitMin, negHitMax));  // __hmin2 drops the 0*inf NaN of a flat black / white ring
}

__device__ __forceinline__ __half2 resolve_channel(__half2 lobe, __half2 rcpL, __half2 b, __half2 d, __half2 e,
                                                   __half2 f, __half2 h) {
  const __half2 ring = __hadd2(__hadd2(b, d), __hadd2(h, f));
  return __hmul2(__hfma2(lobe, ring, e), rcpL);
}

template <bool kChecked, bool kClamp>
__device__ __forceinline__ void rcas_rows(const RcasParams& p, int x, int ys, int lane) {
  const __half2 sharp = uh2(p.sharp_h2);
  const __half2 kLimit = __float2half2_rn(-0.1875f), kZero = __float2half2_rn(0.0f);
  const uint32_t one = 0x3c003c00u;
  const bool writer = lane >= 1 && lane <= 30 && (!kChecked || x < p.out.w);
  // all kRows+2 rows are requested up front: kRows+2 independent 16-byte loads in flight per lane
  Row3 rows[kRows + 2];
  if (!kChecked) {  // one 64-bit address, then += pitch: no per-row address arithmetic
    const unsigned char* src = p.in.base + (long long)(ys - 1 - p.in.row0) * p.in.pitch + (long long)x * 8;
#pragma unroll
    for (int r = 0; r < kRows + 2; r++) rows[r] = to_soa(__ldg(reinterpret_cast<const uint4*>(src + (long long)r * p.in.pitch)));
  } else {
#pragma unroll
    for (int r = 0; r < kRows + 2; r++) rows[r] = load_pair<kChecked, kClamp>(p, x, ys - 1 + r);
  }
  unsigned char* dst = p.out.base + (long long)(ys - p.out.row0) * p.out.pitch + (long long)x * 8;
#pragma unroll
  for (int r = 0; r < kRows; r++) {
    const int y = ys + r;
    if (kChecked && y >= p.y1) break;  // warp-uniform
    const Row3 prev = rows[r], cur = rows[r + 1], next = rows[r + 2];
    // d = (left lane's pixel1, my pixel0), f = (my pixel1, right lane's pixel0)
    const __half2 dR = uh2(__byte_perm(__shfl_up_sync(0xffffffffu, hu2(cur.r), 1), hu2(cur.r), 0x5432));
    const __half2 dG = uh2(__byte_perm(__shfl_up_sync(0xffffffffu, hu2(cur.g), 1), hu2(cur.g), 0x5432));
    const __half2 dB = uh2(__byte_perm(__shfl_up_sync(0xffffffffu, hu2(cur.b), 1), hu2(cur.b), 0x5432));
    const __half2 fR = uh2(__byte_perm(hu2(cur.r), __shfl_down_sync(0xffffffffu, hu2(cur.r), 1), 0x5432));
    const __half2 fG = uh2(__byte_perm(hu2(cur.g), __shfl_down_sync(0xffffffffu, hu2(cur.g), 1), 0x5432));
    const __half2 fB = uh2(__byte_perm(hu2(cur.b), __shfl_down_sync(0xffffffffu, hu2(cur.b), 1), 0x5432));

    const __half2 lR = lobe_channel(prev.r, dR, cur.r, fR, next.r);
    const __half2 lG = lobe_channel(prev.g, dG, cur.g, fG, next.g);
    const __half2 lB = lobe_channel(prev.b, dB, cur.b, fB, next.b);
    const __half2 lobe = __hmul2(__hmax2(kLimit, __hmin2(__hmax2(lR, __hmax2(lG, lB)), kZero)), sharp);
    // APrxMedRcpH2(4*lobe+1): packed 16-bit magic subtract (no borrow: both lanes' bits <= 0x3c00) + one Newton step
    const __half2 a = __hfma2(__float2half2_rn(4.0f), lobe, __float2half2_rn(1.0f));
    const __half2 s = uh2(0x778d778du - hu2(a));
    const __half2 rcpL = __hmul2(s, __hfma2(__hneg2(s), a, __float2half2_rn(2.0f)));
    const __half2 oR = resolve_channel(lobe, rcpL, prev.r, dR, cur.r, fR, next.r);
    const __half2 oG = resolve_channel(lobe, rcpL, prev.g, dG, cur.g, fG, next.g);
    const __half2 oB = resolve_channel(lobe, rcpL, prev.b, dB, cur.b, fB, next.b);

    if (writer) {
      unsigned char* o = dst + (long long)r * p.out.pitch;
      const uint32_t rg0 = __byte_perm(hu2(oR), hu2(oG), 0x5410), b0 = __byte_perm(hu2(oB), one, 0x5410);
      if (!kChecked || x + 1 < p.out.w) {
        const uint32_t rg1 = __byte_perm(hu2(oR), hu2(oG), 0x7632), b1 = __byte_perm(hu2(oB), one, 0x7632);
        *reinterpret_cast<uint4*>(o) = make_uint4(rg0, b0, rg1, b1);
      } else {
        *reinterpret_cast<uint2*>(o) = make_uint2(rg0, b0);
      }
    }
  }
}

template <bool kClamp>
__global__ void __launch_bounds__(32 * kNW) rcas_h_packed_kernel(const RcasParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x0 = blockIdx.x * kSpan - 2;  // even -> every lane's pair is 16-byte aligned
  const int x = x0 + lane * 2;
  const int ys = p.y0 + (blockIdx.y * kNW + warp) * kRows;
  if (ys >= p.y1) return;  // whole warp
  const bool interior = x0 >= 0 && x0 + 64 <= p.in.w && ys >= 1 && ys + kRows < p.in.h && ys + kRows <= p.y1;
  if (interior)
    rcas_rows<false, kClamp>(p, x, ys, lane);
  else
    rcas_rows<true, kClamp>(p, x, ys, lane);
}

// ---- RCAS for the UNORM formats (experimental, FSR1_UNORM_TILED=1) ---------------------------------------------------
// The structure of rcas_rows / rcas_h_packed_kernel with 4-byte texels: a lane still owns two adjacent pixels (one 64-bit
// load and store per row), decodes them to the (pixel0, pixel1)-per-channel half2 form the arithmetic above works on
// (c / (2^n - 1) in fp32, one rounding to half), and re-encodes the saturated result in the half domain
// (x * (2^n - 1) + 1024 leaves round(x * (2^n - 1)) in the low mantissa bits).
template <int kBits> __device__ __forceinline__ Row3 decode_pair(uint2 v) {
  Row3 o;
  if (kBits == 8) {
    const float k = 1.0f / 255.0f;
    o.r = __floats2half2_rn((float)(v.x & 255u) * k, (float)(v.y & 255u) * k);
    o.g = __floats2half2_rn((float)((v.x >> 8) & 255u) * k, (float)((v.y >> 8) & 255u) * k);
    o.b = __floats2half2_rn((float)((v.x >> 16) & 255u) * k, (float)((v.y >> 16) & 255u) * k);
  } else {
    const float k = 1.0f / 1023.0f;
    o.r = __floats2half2_rn((float)(v.x & 1023u) * k, (float)(v.y & 1023u) * k);
    o.g = __floats2half2_rn((float)((v.x >> 10) & 1023u) * k, (float)((v.y >> 10) & 1023u) * k);
    o.b = __floats2half2_rn((float)((v.x >> 20) & 1023u) * k, (float)((v.y >> 20) & 1023u) * k);
  }
  return o;
}

template <int kBits> __device__ __forceinline__ uint2 encode_pair(__half2 r, __half2 g, __half2 b) {  // inputs in [0,1]
  const __half2 sc = __float2half2_rn(kBits == 8 ? 255.0f : 1023.0f), k1024 = __float2half2_rn(1024.0f);
  const uint32_t tr = hu2(__hfma2(r, sc, k1024)), tg = hu2(__hfma2(g, sc, k1024)), tb = hu2(__hfma2(b, sc, k1024));
  if (kBits == 8) {
    const uint32_t rg0 = __byte_perm(tr, tg, 0x0040), rg1 = __byte_perm(tr, tg, 0x0062);  // (R, G) of pixel 0 / pixel 1
    const uint32_t ba0 = __byte_perm(tb, 0xffffffffu, 0x0040), ba1 = __byte_perm(tb, 0xffffffffu, 0x0042);  // (B, 255)
    return make_uint2(__byte_perm(rg0, ba0, 0x5410), __byte_perm(rg1, ba1, 0x5410));
  }
  return make_uint2((tr & 0x3ffu) | ((tg & 0x3ffu) << 10) | ((tb & 0x3ffu) << 20) | 0xC0000000u,
                    ((tr >> 16) & 0x3ffu) | (((tg >> 16) & 0x3ffu) << 10) | (((tb >> 16) & 0x3ffu) << 20) | 0xC0000000u);
}

// Pixels (x, x+1) of logical row y as raw words; out-of-image pixels read 0 (D3D12 Load) or are clamped.
template <bool kClamp> __device__ __forceinline__ uint2 load_words(const RcasParams& p, int x, int y) {
  uint2 v = make_uint2(0u, 0u);
  if (kClamp) y = clampi(y, 0, p.in.h - 1);
  if (row_stored(p.in, y)) {
    const uint32_t* row = reinterpret_cast<const uint32_t*>(p.in.base + (long long)(y - p.in.row0) * p.in.pitch);
    if (kClamp) {
      v.x = __ldg(row + clampi(x, 0, p.in.w - 1));
      v.y = __ldg(row + clampi(x + 1, 0, p.in.w - 1));
    } else {
      if (x >= 0 && x < p.in.w) v.x = __ldg(row + x);
      if (x + 1 >= 0 && x + 1 < p.in.w) v.y = __ldg(row + x + 1);
    }
  }
  return v;
}

template <bool kChecked, bool kClamp, int kBits>
__device__ __forceinline__ void rcas_rows_u(const RcasParams& p, int x, int ys, int lane) {
  const __half2 sharp = uh2(p.sharp_h2);
  const __half2 kLimit = __float2half2_rn(-0.1875f), kZero = __float2half2_rn(0.0f);
  const bool writer = lane >= 1 && lane <= 30 && (!kChecked || x < p.out.w);
  Row3 rows[kRows + 2];
  if (!kChecked) {
    const unsigned char* src = p.in.base + (long long)(ys - 1 - p.in.row0) * p.in.pitch + (long long)x * 4;
#pragma unroll
    for (int r = 0; r < kRows + 2; r++) rows[r] = decode_pair<kBits>(__ldg(reinterpret_cast<const uint2*>(src + (long long)r * p.in.pitch)));
  } else {
#pragma unroll
    for (int r = 0; r < kRows + 2; r++) rows[r] = decode_pair<kBits>(load_words<kClamp>(p, x, ys - 1 + r));
  }
  unsigned char* dst = p.out.base + (long long)(ys - p.out.row0) * p.out.pitch + (long long)x * 4;
#pragma unroll
  for (int r = 0; r < kRows; r++) {
    const int y = ys + r;
    if (kChecked && y >= p.y1) break;  // warp-uniform
    const Row3 prev = rows[r], cur = rows[r + 1], next = rows[r + 2];
    const __half2 dR = uh2(__byte_perm(__shfl_up_sync(0xffffffffu, hu2(cur.r), 1), hu2(cur.r), 0x5432));
    const __half2 dG = uh2(__byte_perm(__shfl_up_sync(0xffffffffu, hu2(cur.g), 1), hu2(cur.g), 0x5432));
    const __half2 dB = uh2(__byte_perm(__shfl_up_sync(0xffffffffu, hu2(cur.b), 1), hu2(cur.b), 0x5432));
    const __half2 fR = uh2(__byte_perm(hu2(cur.r), __shfl_down_sync(0xffffffffu, hu2(cur.r), 1), 0x5432));
    const __half2 fG = uh2(__byte_perm(hu2(cur.g), __shfl_down_sync(0xffffffffu, hu2(cur.g), 1), 0x5432));
    const __half2 fB = uh2(__byte_perm(hu2(cur.b), __shfl_down_sync(0xffffffffu, hu2(cur.b), 1), 0x5432));
    const __half2 lR = lobe_channel(prev.r, dR, cur.r, fR, next.r);
    const __half2 lG = lobe_channel(prev.g, dG, cur.g, fG, next.g);
    const __half2 lB = lobe_channel(prev.b, dB, cur.b, fB, next.b);
    const __half2 lobe = __hmul2(__hmax2(kLimit, __hmin2(__hmax2(lR, __hmax2(lG, lB)), kZero)), sharp);
    const __half2 a = __hfma2(__float2half2_rn(4.0f), lobe, __float2half2_rn(1.0f));
    const __half2 s = uh2(0x778d778du - hu2(a));
    const __half2 rcpL = __hmul2(s, __hfma2(__hneg2(s), a, __float2half2_rn(2.0f)));
    // resolve, saturated to [0,1] for the encode (a UNORM store clamps anyway)
    const __half2 one = __float2half2_rn(1.0f);
    const __half2 oR = __hmin2(one, __hmax2(kZero, resolve_channel(lobe, rcpL, prev.r, dR, cur.r, fR, next.r)));
    const __half2 oG = __hmin2(one, __hmax2(kZero, resolve_channel(lobe, rcpL, prev.g, dG, cur.g, fG, next.g)));
    const __half2 oB = __hmin2(one, __hmax2(kZero, resolve_channel(lobe, rcpL, prev.b, dB, cur.b, fB, next.b)));
    if (writer) {
      unsigned char* o = dst + (long long)r * p.out.pitch;
      const uint2 w = encode_pair<kBits>(oR, oG, oB);
      if (!kChecked || x + 1 < p.out.w) *reinterpret_cast<uint2*>(o) = w;
      else *reinterpret_cast<uint32_t*>(o) = w.x;
    }
  }
}

template <bool kClamp, int kBits>
__global__ void __launch_bounds__(32 * 4) rcas_u_packed_kernel(const RcasParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x0 = blockIdx.x * kSpan - 2;  // even -> every lane's pair is 8-byte aligned
  const int x = x0 + lane * 2;
  const int ys = p.y0 + (blockIdx.y * kNW + warp) * kRows;
  if (ys >= p.y1) return;  // whole warp
  const bool interior = x0 >= 0 && x0 + 64 <= p.in.w && ys >= 1 && ys + kRows < p.in.h && ys + kRows <= p.y1;
  if (interior) rcas_rows_u<false, kClamp, kBits>(p, x, ys, lane);
  else rcas_rows_u<true, kClamp, kBits>(p, x, ys, lane);
}

#ifndef FSR1_CPU_EMU  // tests/emu compiles the device code above for the host and supplies its own launcher
cudaError_t launch_rcas_u_packed(const RcasParams& p, int format, cudaStream_t s, const char** name) {
  if (format != 3 && format != 4) return cudaErrorNotSupported;
  if ((reinterpret_cast<uintptr_t>(p.in.base) & 7) || (p.in.pitch & 7) || (reinterpret_cast<uintptr_t>(p.out.base) & 7) || (p.out.pitch & 7))
    return cudaErrorNotSupported;
  const dim3 grid((p.out.w + kSpan - 1) / kSpan, (p.y1 - p.y0 + 15) / 16, 1);
  if (format == 3) {
    if (p.clamp) rcas_u_packed_kernel<true, 8><<<grid, 128, 0, s>>>(p);
    else rcas_u_packed_kernel<false, 8><<<grid, 128, 0, s>>>(p);
    *name = "rcas_u8_packed<2px,4rows,shfl60>";
  } else {
    if (p.clamp) rcas_u_packed_kernel<true, 10><<<grid, 128, 0, s>>>(p);
    else rcas_u_packed_kernel<false, 10><<<grid, 128, 0, s>>>(p);
    *name = "rcas_u10_packed<2px,4rows,shfl60>";
  }
  return cudaGetLastError();
}

cudaError_t launch_rcas_h_packed(const RcasParams& p, cudaStream_t s, const char** name) {
  if ((reinterpret_cast<uintptr_t>(p.in.base) & 15) || (p.in.pitch & 15) || (reinterpret_cast<uintptr_t>(p.out.base) & 15) ||
      (p.out.pitch & 15))
    return cudaErrorNotSupported;
  // 4-warp CTAs (60 x 16 pixels), 4 rows per lane, MUFU reciprocals: same kernel time as 8-warp CTAs, but the smaller
  // CTA starts earlier in the tail of the preceding EASU and shares SMs with it when frames are pipelined
  // (round 1: 92.1 -> 90.7 us per frame back to back; the 8-row and Newton-reciprocal variants measured slower)
  const dim3 grid((p.out.w + kSpan - 1) / kSpan, (p.y1 - p.y0 + kNW * kRows - 1) / (kNW * kRows), 1);
  if (p.clamp) rcas_h_packed_kernel<true><<<grid, 32 * kNW, 0, s>>>(p);
  else rcas_h_packed_kernel<false><<<grid, 32 * kNW, 0, s>>>(p);
  *name = "rcas_h_packed<2px,4rows,shfl60>";
  return cudaGetLastError();
}

#endif  // FSR1_CPU_EMU

}  // namespace fsr1
