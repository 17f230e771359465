// fsr1_hx2.cu — the reference's PACKED calling convention ("...Hx2"): two pixels, ip and ip + (8,0), per lane, held as
// structure-of-arrays half2 registers (R0,R1) (G0,G1) (B0,B1), every operation a packed half operation.
//
//   FsrRcasHx2 + FsrRcasDepackHx2      ffx-fsr/ffx_fsr1.h:880-984     (FSR1_FLAG_RCAS_HX2 on fsr1_rcas)
//   FsrLfgaHx2                         ffx-fsr/ffx_fsr1.h:1022-1024   (fsr1_lfga_h)
//   FsrSrtmHx2 / FsrSrtmInvHx2         ffx-fsr/ffx_fsr1.h:1052-1055   (fsr1_srtm_h)
//   FsrTepdDitHx2                      ffx-fsr/ffx_fsr1.h:1156-1164   (fsr1_tepd_h, positional dither)
//   FsrTepdC8Hx2 / FsrTepdC10Hx2       ffx-fsr/ffx_fsr1.h:1166-1199   (fsr1_tepd_h)
//
// Lane for lane the packed functions perform the operations of the scalar H functions (FsrRcasH :782-866, FsrLfgaH :1019,
// FsrSrtmH / FsrSrtmInvH :1049-1050, FsrTepdC8H / C10H :1137-1153), so one kernel serves both spellings; oracle tests
// (tests/test_oracle.py) show Hx2 == H bit for bit on the reference's own source.
//
// Arithmetic policy = fsr1_href.cu's: every operation rounds to half ONCE and nothing is fused (add / sub / mul are the packed
// HADD2 / HMUL2 forms with .rn and no contraction; min / max / saturate / reciprocal / sqrt / floor go through fp32 per lane,
// which is exact for half operands), so results are BIT-IDENTICAL to the reference's half source executed with per-operation
// half rounding (oracle/_ref built with A_HALF).  RGBA16F images only.  Like the H-reference kernels these are parity /
// calling-convention paths with direct loads, not the fast path: the production RCAS kernel keeps the same SoA layout but
// feeds it from 128-bit row loads and warp shuffles (fsr1_rcas_packed.cu).
#include "fsr1_common.cuh"

namespace fsr1 {

struct H2 {  // two halves whose every operator rounds each lane once, never fuses
  __half2 v;
};
__device__ __forceinline__ H2 h2_of(float lo, float hi) { return H2{__floats2half2_rn(lo, hi)}; }
__device__ __forceinline__ H2 h2_all(float a) { return H2{__float2half2_rn(a)}; }
__device__ __forceinline__ float2 f2_of(H2 a) { return __half22float2(a.v); }
__device__ __forceinline__ uint32_t h2_bits(H2 a) { return *reinterpret_cast<const uint32_t*>(&a.v); }
__device__ __forceinline__ H2 h2_from_bits(uint32_t u) {
  H2 r;
  *reinterpret_cast<uint32_t*>(&r.v) = u;
  return r;
}
__device__ __forceinline__ H2 operator+(H2 a, H2 b) { return H2{__hadd2_rn(a.v, b.v)}; }
__device__ __forceinline__ H2 operator-(H2 a, H2 b) { return H2{__hsub2_rn(a.v, b.v)}; }
__device__ __forceinline__ H2 operator*(H2 a, H2 b) { return H2{__hmul2_rn(a.v, b.v)}; }
__device__ __forceinline__ H2 operator-(H2 a) { return h2_from_bits(h2_bits(a) ^ 0x80008000u); }
__device__ __forceinline__ H2 h2_abs(H2 a) { return h2_from_bits(h2_bits(a) & 0x7fff7fffu); }
__device__ __forceinline__ H2 h2_min(H2 a, H2 b) {
  const float2 x = f2_of(a), y = f2_of(b);
  return h2_of(fminf(x.x, y.x), fminf(x.y, y.y));
}
__device__ __forceinline__ H2 h2_max(H2 a, H2 b) {
  const float2 x = f2_of(a), y = f2_of(b);
  return h2_of(fmaxf(x.x, y.x), fmaxf(x.y, y.y));
}
__device__ __forceinline__ H2 h2_min3(H2 a, H2 b, H2 c) { return h2_min(a, h2_min(b, c)); }  // AMin3H2 (ffx_a.h:1356)
__device__ __forceinline__ H2 h2_max3(H2 a, H2 b, H2 c) { return h2_max(a, h2_max(b, c)); }  // AMax3H2 (ffx_a.h:1346)
__device__ __forceinline__ H2 h2_sat(H2 a) {  // saturate(NaN) = 0
  const float2 x = f2_of(a);
  return h2_of(fminf(fmaxf(x.x, 0.0f), 1.0f), fminf(fmaxf(x.y, 0.0f), 1.0f));
}
__device__ __forceinline__ H2 h2_rcp(H2 a) {  // ARcpH2 = rcp: correctly rounded 1/a
  const float2 x = f2_of(a);
  return h2_of(__fdiv_rn(1.0f, x.x), __fdiv_rn(1.0f, x.y));
}
__device__ __forceinline__ H2 h2_sqrt(H2 a) {
  const float2 x = f2_of(a);
  return h2_of(__fsqrt_rn(x.x), __fsqrt_rn(x.y));
}
__device__ __forceinline__ H2 h2_floor(H2 a) {
  const float2 x = f2_of(a);
  return h2_of(floorf(x.x), floorf(x.y));
}
// APrxMedRcpH2 (ffx_a.h:1815): the 16-bit integer subtract wraps inside each lane
__device__ __forceinline__ H2 h2_prx_med_rcp(H2 a) {
  const uint32_t u = h2_bits(a);
  const H2 b = h2_from_bits(((0x778du - (u & 0xffffu)) & 0xffffu) | ((0x778du - (u >> 16)) << 16));
  return b * (-b * a + h2_all(2.0f));
}

struct Px2 {  // two pixels, structure of arrays
  H2 r, g, b;
};
// arrays of structures -> structure of arrays (ffx_fsr1.h:927-941): texel a -> lane .x, texel b -> lane .y
__device__ __forceinline__ Px2 soa_of(uint2 a, uint2 b) {
  Px2 o;
  o.r = h2_from_bits((a.x & 0xffffu) | (b.x << 16));
  o.g = h2_from_bits((a.x >> 16) | (b.x & 0xffff0000u));
  o.b = h2_from_bits((a.y & 0xffffu) | (b.y << 16));
  return o;
}
// FsrRcasDepackHx2 (:880-886) + the store of both pixels; alpha bits are passed separately
__device__ __forceinline__ void store_pair(const ImgView& im, int x0, int x1, bool has1, int y, const Px2& c, uint32_t alpha0, uint32_t alpha1) {
  uint2* row = reinterpret_cast<uint2*>(im.base + (long long)(y - im.row0) * im.pitch);
  const uint32_t r = h2_bits(c.r), g = h2_bits(c.g), b = h2_bits(c.b);
  row[x0] = make_uint2((r & 0xffffu) | (g << 16), (b & 0xffffu) | (alpha0 << 16));
  if (has1) row[x1] = make_uint2((r >> 16) | (g & 0xffff0000u), (b >> 16) | (alpha1 << 16));
}

// FsrRcasLoadHx2: integer-coordinate load of the EASU output, out-of-image policy as the other RCAS kernels
__device__ __forceinline__ uint2 rcas_texel(const RcasParams& p, int x, int y) {
  if (p.clamp) {
    x = clampi(x, 0, p.in.w - 1);
    y = clampi(y, 0, p.in.h - 1);
  } else if (x < 0 || y < 0 || x >= p.in.w || y >= p.in.h) {
    return make_uint2(0u, 0u);
  }
  return __ldg(reinterpret_cast<const uint2*>(p.in.base + (long long)(y - p.in.row0) * p.in.pitch) + x);
}

constexpr int kHx2Threads = 128;                    // 16 groups of 8 lanes
constexpr int kHx2Span = (kHx2Threads / 8) * 16;    // pixels of one row a CTA covers: each group owns a 16-pixel strip

// lane -> its two pixel columns: lane i of group g handles x0 = strip + i and x1 = x0 + 8 (ffx_fsr1.h:891-894, 915)
__device__ __forceinline__ int hx2_x0() { return (int)blockIdx.x * kHx2Span + ((int)threadIdx.x >> 3) * 16 + ((int)threadIdx.x & 7); }

__global__ void __launch_bounds__(kHx2Threads) rcas_hx2_kernel(const RcasParams p) {
  const int x0 = hx2_x0(), x1 = x0 + 8;
  const int y = p.y0 + (int)blockIdx.y;
  if (x0 >= p.out.w || y >= p.y1) return;
  const bool has1 = x1 < p.out.w;
  //    b
  //  d e f     for both pixels (:902-925)
  //    h
  const uint2 e0 = rcas_texel(p, x0, y), e1 = rcas_texel(p, x1, y);
  const Px2 b = soa_of(rcas_texel(p, x0, y - 1), rcas_texel(p, x1, y - 1));
  const Px2 d = soa_of(rcas_texel(p, x0 - 1, y), rcas_texel(p, x1 - 1, y));
  const Px2 e = soa_of(e0, e1);
  const Px2 f = soa_of(rcas_texel(p, x0 + 1, y), rcas_texel(p, x1 + 1, y));
  const Px2 h = soa_of(rcas_texel(p, x0, y + 1), rcas_texel(p, x1, y + 1));
  const H2 four = h2_all(4.0f), one = h2_all(1.0f);
  // min and max of the ring (:958-963), limiters with exact reciprocals (:967-972), lobe (:973-976)
  H2 lobe;
  {
    H2 lobeC[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const H2 bb = k == 0 ? b.r : (k == 1 ? b.g : b.b), dd = k == 0 ? d.r : (k == 1 ? d.g : d.b), ee = k == 0 ? e.r : (k == 1 ? e.g : e.b);
      const H2 ff = k == 0 ? f.r : (k == 1 ? f.g : f.b), hh = k == 0 ? h.r : (k == 1 ? h.g : h.b);
      const H2 mn4 = h2_min(h2_min3(bb, dd, ff), hh);
      const H2 mx4 = h2_max(h2_max3(bb, dd, ff), hh);
      const H2 hitMin = h2_min(mn4, ee) * h2_rcp(four * mx4);
      const H2 hitMax = (one - h2_max(mx4, ee)) * h2_rcp(four * mn4 + h2_all(-4.0f));
      lobeC[k] = h2_max(-hitMin, hitMax);
    }
    const H2 sharp = h2_from_bits((p.sharp_h2 & 0xffffu) * 0x10001u);  // AH2_(AH2_AU1(con.y).x) (:976)
    lobe = h2_max(h2_all(-0.1875f), h2_min(h2_max3(lobeC[0], lobeC[1], lobeC[2]), h2_all(0.0f))) * sharp;
  }
  if (p.options & 1) {  // FSR_RCAS_DENOISE: luma times 2 (:948-952), noise detection (:954-956), lobe *= nz (:978-980)
    const H2 hf = h2_all(0.5f), q = h2_all(0.25f);
    const H2 bL = b.b * hf + (b.r * hf + b.g), dL = d.b * hf + (d.r * hf + d.g), eL = e.b * hf + (e.r * hf + e.g);
    const H2 fL = f.b * hf + (f.r * hf + f.g), hL = h.b * hf + (h.r * hf + h.g);
    H2 nz = q * bL + q * dL + q * fL + q * hL - eL;
    nz = h2_sat(h2_abs(nz) * h2_prx_med_rcp(h2_max3(h2_max3(bL, dL, eL), fL, hL) - h2_min3(h2_min3(bL, dL, eL), fL, hL)));
    nz = h2_all(-0.5f) * nz + one;
    lobe = lobe * nz;
  }
  // resolve with the medium-precision reciprocal (:982-985)
  const H2 rcpL = h2_prx_med_rcp(four * lobe + one);
  Px2 o;
  o.r = (lobe * b.r + lobe * d.r + lobe * h.r + lobe * f.r + e.r) * rcpL;
  o.g = (lobe * b.g + lobe * d.g + lobe * h.g + lobe * f.g + e.g) * rcpL;
  o.b = (lobe * b.b + lobe * d.b + lobe * h.b + lobe * f.b + e.b) * rcpL;
  uint32_t a0 = 0x3c00u, a1 = 0x3c00u;
  if (p.options & 2) {  // FSR_RCAS_PASSTHROUGH_ALPHA (:907-909, 918-920): the centre texels' alpha
    a0 = e0.y >> 16;
    a1 = e1.y >> 16;
  }
  store_pair(p.out, x0, x1, has1, y, o, a0, a1);
}

// ---- pointwise companions, packed half ---------------------------------------------------------------------------------
enum { kHOpSrtm = 1, kHOpSrtmInv = 2, kHOpLfga = 3, kHOpTepd8 = 4, kHOpTepd10 = 5 };

struct PointHParams {
  ImgView in, out, aux;  // aux: grain (LFGA) or dither (TEPD) tile, RGBA16F, wraps over the frame
  int has_aux;
  int op;
  float amount;          // LFGA: converted to half once (AH1 a)
  uint32_t frame;        // TEPD positional dither
  int y0, y1;
};

__device__ __forceinline__ uint2 texel_of(const ImgView& im, int x, int y) {
  return __ldg(reinterpret_cast<const uint2*>(im.base + (long long)(y - im.row0) * im.pitch) + x);
}

// FsrTepdC8Hx2 / FsrTepdC10Hx2 for one channel pair (:1166-1199); q = 255 or 1023, rq = half(1/q)
__device__ __forceinline__ H2 tepd_hx2_channel(H2 c, H2 dit, H2 q, H2 rq) {
  H2 n = h2_sqrt(c);
  n = h2_floor(n * q) * rq;
  const H2 a = n * n;
  H2 b = n + rq;
  b = b * b;
  const H2 r = (c - b) * h2_prx_med_rcp(a - b);
  // AGtZeroH2(m) = saturate(m * +INF) (ffx_a.h:1525): 1 for m > 0, else 0 (0 * INF = NaN saturates to 0)
  const H2 gt = h2_sat((dit - r) * h2_from_bits(0x7c007c00u));
  return h2_sat(n + gt * rq);
}

__global__ void __launch_bounds__(kHx2Threads) pointwise_hx2_kernel(const PointHParams p) {
  const int x0 = hx2_x0(), x1 = x0 + 8;
  const int y = p.y0 + (int)blockIdx.y;
  if (x0 >= p.out.w || y >= p.y1) return;
  const bool has1 = x1 < p.out.w;
  const int x1l = has1 ? x1 : x0;  // the idle lane of a strip that crosses the right edge repeats pixel 0
  const uint2 t0 = texel_of(p.in, x0, y), t1 = texel_of(p.in, x1l, y);
  Px2 c = soa_of(t0, t1);
  const H2 one = h2_all(1.0f);
  switch (p.op) {
    case kHOpSrtm: {  // FsrSrtmHx2 (:1052-1053)
      const H2 rcp = h2_rcp(h2_max3(c.r, c.g, c.b) + one);
      c.r = c.r * rcp; c.g = c.g * rcp; c.b = c.b * rcp;
      break;
    }
    case kHOpSrtmInv: {  // FsrSrtmInvHx2 (:1054-1055): the extra max solves c = 1.0
      const H2 rcp = h2_rcp(h2_max(h2_from_bits(0x02000200u) /* 1/32768 */, one - h2_max3(c.r, c.g, c.b)));
      c.r = c.r * rcp; c.g = c.g * rcp; c.b = c.b * rcp;
      break;
    }
    case kHOpLfga: {  // FsrLfgaHx2 (:1022-1024): c += (t * a) * min(1 - c, c)
      const int ay = y % p.aux.h;
      const Px2 t = soa_of(texel_of(p.aux, x0 % p.aux.w, ay), texel_of(p.aux, x1l % p.aux.w, ay));
      const H2 a = h2_all(p.amount);
      c.r = c.r + (t.r * a) * h2_min(one - c.r, c.r);
      c.g = c.g + (t.g * a) * h2_min(one - c.g, c.g);
      c.b = c.b + (t.b * a) * h2_min(one - c.b, c.b);
      break;
    }
    default: {  // kHOpTepd8 / kHOpTepd10
      H2 dit;
      if (p.has_aux) {  // a blue-noise tile's .w (sample/src/DX12/FSR_Tonemapping.hlsl:87), saturated
        const int ay = y % p.aux.h;
        const uint2 d0 = texel_of(p.aux, x0 % p.aux.w, ay), d1 = texel_of(p.aux, x1l % p.aux.w, ay);
        dit = h2_sat(h2_from_bits((d0.y >> 16) | (d1.y & 0xffff0000u)));
      } else {  // FsrTepdDitHx2 (:1156-1164): fp32 position hash of p and p + (8,0), converted to half once
        const float fx0 = (float)((uint32_t)x0 + p.frame), fx1 = __fadd_rn(fx0, 8.0f), fy = (float)y;
        const float a = 1.61803398874989484820f, b = (float)(1.0 / 3.69);
        const float yb = __fmul_rn(fy, b);
        const float v0 = __fadd_rn(__fmul_rn(fx0, a), yb), v1 = __fadd_rn(__fmul_rn(fx1, a), yb);
        dit = h2_of(__fsub_rn(v0, floorf(v0)), __fsub_rn(v1, floorf(v1)));
      }
      const bool c8 = p.op == kHOpTepd8;
      const H2 q = h2_all(c8 ? 255.0f : 1023.0f);
      const H2 rq = h2_from_bits(c8 ? 0x1c041c04u : 0x14011401u);  // AH2_(1.0/255.0), AH2_(1.0/1023.0)
      c.r = tepd_hx2_channel(c.r, dit, q, rq);
      c.g = tepd_hx2_channel(c.g, dit, q, rq);
      c.b = tepd_hx2_channel(c.b, dit, q, rq);
      break;
    }
  }
  store_pair(p.out, x0, x1, has1, y, c, t0.y >> 16, t1.y >> 16);  // alpha is carried through
}

#ifndef FSR1_CPU_EMU
cudaError_t launch_rcas_hx2(const RcasParams& p, cudaStream_t s, const char** name) {
  rcas_hx2_kernel<<<dim3((p.out.w + kHx2Span - 1) / kHx2Span, p.y1 - p.y0, 1), kHx2Threads, 0, s>>>(p);
  *name = "rcas_hx2<FsrRcasHx2 calling convention>";
  return cudaGetLastError();
}

// op: 1 SRTM, 2 SRTM inverse, 3 LFGA, 4 TEPD 8 bit, 5 TEPD 10 bit (the numbering of launch_pointwise); RGBA16F everywhere
cudaError_t launch_pointwise_hx2(int op, const ImgView& in, const ImgView& out, const ImgView* aux, float amount, uint32_t frame, int y0,
                                 int y1, cudaStream_t s, const char** name) {
  static const char* const names[] = {"", "pointwise_hx2<FsrSrtmHx2>", "pointwise_hx2<FsrSrtmInvHx2>", "pointwise_hx2<FsrLfgaHx2>",
                                      "pointwise_hx2<FsrTepdC8Hx2>", "pointwise_hx2<FsrTepdC10Hx2>"};
  if (op < kHOpSrtm || op > kHOpTepd10) return cudaErrorInvalidValue;
  if (op == kHOpLfga && !aux) return cudaErrorInvalidValue;
  PointHParams p;
  p.in = in;
  p.out = out;
  p.has_aux = aux ? 1 : 0;
  p.aux = aux ? *aux : in;
  p.op = op;
  p.amount = amount;
  p.frame = frame;
  p.y0 = y0;
  p.y1 = y1;
  *name = names[op];
  pointwise_hx2_kernel<<<dim3((out.w + kHx2Span - 1) / kHx2Span, y1 - y0, 1), kHx2Threads, 0, s>>>(p);
  return cudaGetLastError();
}
#endif

}  // namespace fsr1
