// fsr1_href.cu — literal packed-half semantics of the reference (FSR1_FLAG_H_REFERENCE, RGBA16F only).
//
// The production fp16 kernels run the F algorithm in mixed precision (DESIGN.md "numerics"), because the
// reference's own H path differs from its F path by up to 0.1.  Callers who need what the reference's fp16
// shader computes — FsrEasuH (ffx-fsr/ffx_fsr1.h:452-593: exact ARcpH2 in the Set step, the 0x7784/0x59a3 half
// magic numbers, two taps per packed lane, the (-x,x) min/max trick) and FsrRcasH (:782-866, APrxMedRcpH1
// 0x778d) — get it here: every operation is an IEEE half operation with its own rounding (no FMA contraction),
// in the reference's order, so the result is BIT-IDENTICAL to the reference's H source executed with
// per-operation half rounding (oracle/_ref built with A_HALF; tests/test_gpu_parity.py::test_h_reference_*).
// One thread per output pixel, direct loads: a parity path, not a fast path.
#include "fsr1_common.cuh"

namespace fsr1 {

struct H {  // a half whose every operator rounds once, never fuses
  __half v;
};
__device__ __forceinline__ H hf(float f) { return H{__float2half_rn(f)}; }
__device__ __forceinline__ float fh(H a) { return __half2float(a.v); }
__device__ __forceinline__ H operator+(H a, H b) { return H{__hadd_rn(a.v, b.v)}; }
__device__ __forceinline__ H operator-(H a, H b) { return H{__hsub_rn(a.v, b.v)}; }
__device__ __forceinline__ H operator*(H a, H b) { return H{__hmul_rn(a.v, b.v)}; }
__device__ __forceinline__ H operator-(H a) { return H{__hneg(a.v)}; }
__device__ __forceinline__ H hmin_(H a, H b) { return hf(fminf(fh(a), fh(b))); }
__device__ __forceinline__ H hmax_(H a, H b) { return hf(fmaxf(fh(a), fh(b))); }
__device__ __forceinline__ H habs_(H a) { return H{__habs(a.v)}; }
__device__ __forceinline__ H hsat_(H a) { return hf(fminf(fmaxf(fh(a), 0.0f), 1.0f)); }
__device__ __forceinline__ H hrcp_(H a) { return hf(__fdiv_rn(1.0f, fh(a))); }  // correctly rounded 1/a
__device__ __forceinline__ H hbits(unsigned short u) { return H{__ushort_as_half(u)}; }
__device__ __forceinline__ unsigned short bitsh(H a) { return __half_as_ushort(a.v); }
__device__ __forceinline__ H prx_lo_rcp_hh(H a) { return hbits((unsigned short)(0x7784u - bitsh(a))); }
__device__ __forceinline__ H prx_lo_rsq_hh(H a) { return hbits((unsigned short)(0x59a3u - (bitsh(a) >> 1))); }
__device__ __forceinline__ H prx_med_rcp_hh(H a) {
  const H b = hbits((unsigned short)(0x778du - bitsh(a)));
  return b * (-b * a + hf(2.0f));
}

struct H3 { H r, g, b; };
__device__ __forceinline__ H3 load_h3(const ImgView& im, int x, int y) {
  const uint2 v = __ldg(reinterpret_cast<const uint2*>(im.base + (long long)(y - im.row0) * im.pitch) + x);
  return H3{hbits((unsigned short)(v.x & 0xffffu)), hbits((unsigned short)(v.x >> 16)), hbits((unsigned short)(v.y & 0xffffu))};
}
__device__ __forceinline__ void store_h3(const ImgView& im, int x, int y, H r, H g, H b, unsigned short a = 0x3c00) {
  uint2 v;
  v.x = (uint32_t)bitsh(r) | ((uint32_t)bitsh(g) << 16);
  v.y = (uint32_t)bitsh(b) | ((uint32_t)a << 16);
  reinterpret_cast<uint2*>(im.base + (long long)(y - im.row0) * im.pitch)[x] = v;
}

// FsrEasuSetH for one packed lane (ffx_fsr1.h:476-503)
__device__ __forceinline__ void set_h(H& dx, H& dy, H& len, H w, H lA, H lB, H lC, H lD, H lE) {
  const H dc = lD - lC, cb = lC - lB;
  H lenX = hrcp_(hmax_(habs_(dc), habs_(cb)));
  const H dirX = lD - lB;
  dx = dx + dirX * w;
  lenX = hsat_(habs_(dirX) * lenX);
  lenX = lenX * lenX;
  len = len + lenX * w;
  const H ec = lE - lC, ca = lC - lA;
  H lenY = hrcp_(hmax_(habs_(ec), habs_(ca)));
  const H dirY = lE - lA;
  dy = dy + dirY * w;
  lenY = hsat_(habs_(dirY) * lenY);
  lenY = lenY * lenY;
  len = len + lenY * w;
}

// FsrEasuTapH for one packed lane (ffx_fsr1.h:452-473)
__device__ __forceinline__ void tap_h(H3& aC, H& aW, H ox, H oy, H dx, H dy, H l2x, H l2y, H lob, H clp, H3 c) {
  H vx = ox * dx + oy * dy;
  H vy = ox * (-dy) + oy * dx;
  vx = vx * l2x;
  vy = vy * l2y;
  H d2 = vx * vx + vy * vy;
  d2 = hmin_(d2, clp);
  H wB = hf((float)(2.0 / 5.0)) * d2 + hf(-1.0f);
  H wA = lob * d2 + hf(-1.0f);
  wB = wB * wB;
  wA = wA * wA;
  wB = hf(1.5625f) * wB + hf(-0.5625f);
  const H w = wB * wA;
  aC.r = aC.r + c.r * w;
  aC.g = aC.g + c.g * w;
  aC.b = aC.b + c.b * w;
  aW = aW + w;
}

__global__ void __launch_bounds__(256) easu_href_kernel(const EasuParams p) {
  const int ox = blockIdx.x * 32 + threadIdx.x;
  const int oy = p.y0 + blockIdx.y * 8 + threadIdx.y;
  if (ox >= p.out.w || oy >= p.y1) return;
  int fx, fy;
  float fpx, fpy;
  easu_pos(ox, p.c0x, p.c0z, fx, fpx);
  easu_pos(oy, p.c0y, p.c0w, fy, fpy);
  const H ppx = hf(fpx), ppy = hf(fpy);
  H3 t[4][4];
  H L[4][4];
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int c = 0; c < 4; c++) {
      if ((r == 0 || r == 3) && (c == 0 || c == 3)) continue;
      t[r][c] = load_h3(p.in, clampi(fx - 1 + c, 0, p.in.w - 1), clampi(fy - 1 + r, 0, p.in.h - 1));
      L[r][c] = t[r][c].b * hf(0.5f) + (t[r][c].r * hf(0.5f) + t[r][c].g);
    }
  // packed accumulators: lane .x takes texels f then j, lane .y takes g then k (ffx_fsr1.h:555-558)
  const H zero = hf(0.0f), one = hf(1.0f);
  H dxa = zero, dya = zero, lena = zero, dxb = zero, dyb = zero, lenb = zero;
  const H wl = one + (-ppx), wr = zero + ppx, wt = one - ppy;
  set_h(dxa, dya, lena, wl * wt, L[0][1], L[1][0], L[1][1], L[1][2], L[2][1]);
  set_h(dxb, dyb, lenb, wr * wt, L[0][2], L[1][1], L[1][2], L[1][3], L[2][2]);
  set_h(dxa, dya, lena, wl * ppy, L[1][1], L[2][0], L[2][1], L[2][2], L[3][1]);
  set_h(dxb, dyb, lenb, wr * ppy, L[1][2], L[2][1], L[2][2], L[2][3], L[3][2]);
  H dx = dxa + dxb, dy = dya + dyb, len = lena + lenb;
  H dirR = dx * dx + dy * dy;
  const bool zro = __hlt(dirR.v, __float2half_rn(1.0f / 32768.0f));
  dirR = prx_lo_rsq_hh(dirR);
  dirR = zro ? one : dirR;
  dx = zro ? one : dx;
  dx = dx * dirR;
  dy = dy * dirR;
  len = len * hf(0.5f);
  len = len * len;
  const H stretch = (dx * dx + dy * dy) * prx_lo_rcp_hh(hmax_(habs_(dx), habs_(dy)));
  const H l2x = one + (stretch - one) * len;
  const H l2y = one + hf(-0.5f) * len;
  const H lob = hf(0.5f) + hf((float)((1.0 / 4.0 - 0.04) - 0.5)) * len;
  const H clp = prx_lo_rcp_hh(lob);
  // six tap pairs, lane .x / lane .y accumulated separately: (b,c) (i,j) (f,e) (k,l) (h,g) (o,n)  (:583-590)
  H3 aCx{zero, zero, zero}, aCy{zero, zero, zero};
  H aWx = zero, aWy = zero;
#define FSR1_HT(ACC, AW, R, K) tap_h(ACC, AW, hf((float)((K)-1)) - ppx, hf((float)((R)-1)) - ppy, dx, dy, l2x, l2y, lob, clp, t[R][K]);
  FSR1_HT(aCx, aWx, 0, 1) FSR1_HT(aCy, aWy, 0, 2)
  FSR1_HT(aCx, aWx, 2, 0) FSR1_HT(aCy, aWy, 2, 1)
  FSR1_HT(aCx, aWx, 1, 1) FSR1_HT(aCy, aWy, 1, 0)
  FSR1_HT(aCx, aWx, 2, 2) FSR1_HT(aCy, aWy, 2, 3)
  FSR1_HT(aCx, aWx, 1, 3) FSR1_HT(aCy, aWy, 1, 2)
  FSR1_HT(aCx, aWx, 3, 2) FSR1_HT(aCy, aWy, 3, 1)
#undef FSR1_HT
  const H rW = hrcp_(aWx + aWy);
  H o[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const H f = k == 0 ? t[1][1].r : (k == 1 ? t[1][1].g : t[1][1].b), g = k == 0 ? t[1][2].r : (k == 1 ? t[1][2].g : t[1][2].b);
    const H j = k == 0 ? t[2][1].r : (k == 1 ? t[2][1].g : t[2][1].b), kk = k == 0 ? t[2][2].r : (k == 1 ? t[2][2].g : t[2][2].b);
    const H negmin = hmax_(hmax_(-f, -g), hmax_(-j, -kk));
    const H mx = hmax_(hmax_(f, g), hmax_(j, kk));
    const H acc = k == 0 ? aCx.r + aCy.r : (k == 1 ? aCx.g + aCy.g : aCx.b + aCy.b);
    o[k] = hmin_(mx, hmax_(-negmin, acc * rW));
  }
  store_h3(p.out, ox, oy, o[0], o[1], o[2]);
}

__device__ __forceinline__ H3 rcas_fetch_h(const RcasParams& p, int x, int y) {
  if (p.clamp) {
    x = clampi(x, 0, p.in.w - 1);
    y = clampi(y, 0, p.in.h - 1);
  } else if (x < 0 || y < 0 || x >= p.in.w || y >= p.in.h) {
    return H3{hf(0.f), hf(0.f), hf(0.f)};
  }
  return load_h3(p.in, x, y);
}

__device__ __forceinline__ H lobe_h(H b, H d, H e, H f, H h) {
  const H mn4 = hmin_(hmin_(b, hmin_(d, f)), h);
  const H mx4 = hmax_(hmax_(b, hmax_(d, f)), h);
  const H hitMin = hmin_(mn4, e) * hrcp_(hf(4.0f) * mx4);
  const H hitMax = (hf(1.0f) - hmax_(mx4, e)) * hrcp_(hf(4.0f) * mn4 + hf(-4.0f));
  return hmax_(-hitMin, hitMax);
}
__device__ __forceinline__ H resolve_h(H lobe, H rcpL, H b, H d, H e, H f, H h) {
  return (lobe * b + lobe * d + lobe * h + lobe * f + e) * rcpL;
}

__global__ void __launch_bounds__(256) rcas_href_kernel(const RcasParams p) {
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = p.y0 + blockIdx.y * 8 + threadIdx.y;
  if (x >= p.out.w || y >= p.y1) return;
  const H3 b = rcas_fetch_h(p, x, y - 1), d = rcas_fetch_h(p, x - 1, y), e = rcas_fetch_h(p, x, y);
  const H3 f = rcas_fetch_h(p, x + 1, y), h = rcas_fetch_h(p, x, y + 1);
  const H lR = lobe_h(b.r, d.r, e.r, f.r, h.r), lG = lobe_h(b.g, d.g, e.g, f.g, h.g), lB = lobe_h(b.b, d.b, e.b, f.b, h.b);
  const H sharp = hbits((unsigned short)(p.sharp_h2 & 0xffffu));  // AH2_AU1(con.y).x (ffx_fsr1.h:857)
  H lobe = hmax_(hf(-0.1875f), hmin_(hmax_(lR, hmax_(lG, lB)), hf(0.0f))) * sharp;
  if (p.options & 1) {  // FSR_RCAS_DENOISE (ffx_fsr1.h:829-837, 859-861)
    const H hh = hf(0.5f), q = hf(0.25f);
    const H bL = b.b * hh + (b.r * hh + b.g), dL = d.b * hh + (d.r * hh + d.g), eL = e.b * hh + (e.r * hh + e.g);
    const H fL = f.b * hh + (f.r * hh + f.g), hL = h.b * hh + (h.r * hh + h.g);
    H nz = q * bL + q * dL + q * fL + q * hL - eL;
    const H mx = hmax_(hmax_(bL, hmax_(dL, eL)), hmax_(fL, hL)), mn = hmin_(hmin_(bL, hmin_(dL, eL)), hmin_(fL, hL));
    nz = hsat_(habs_(nz) * prx_med_rcp_hh(mx - mn));
    nz = hf(-0.5f) * nz + hf(1.0f);
    lobe = lobe * nz;
  }
  const H rcpL = prx_med_rcp_hh(hf(4.0f) * lobe + hf(1.0f));
  unsigned short alpha = 0x3c00;
  if (p.options & 2)  // FSR_RCAS_PASSTHROUGH_ALPHA (:786-800): centre texel's alpha
    alpha = (unsigned short)(__ldg(reinterpret_cast<const uint2*>(p.in.base + (long long)(y - p.in.row0) * p.in.pitch) + x).y >> 16);
  store_h3(p.out, x, y, resolve_h(lobe, rcpL, b.r, d.r, e.r, f.r, h.r), resolve_h(lobe, rcpL, b.g, d.g, e.g, f.g, h.g),
           resolve_h(lobe, rcpL, b.b, d.b, e.b, f.b, h.b), alpha);
}

cudaError_t launch_easu_href(const EasuParams& p, cudaStream_t s, const char** name) {
  easu_href_kernel<<<dim3((p.out.w + 31) / 32, (p.y1 - p.y0 + 7) / 8, 1), dim3(32, 8, 1), 0, s>>>(p);
  *name = "easu_href<FsrEasuH semantics>";
  return cudaGetLastError();
}
cudaError_t launch_rcas_href(const RcasParams& p, cudaStream_t s, const char** name) {
  rcas_href_kernel<<<dim3((p.out.w + 31) / 32, (p.y1 - p.y0 + 7) / 8, 1), dim3(32, 8, 1), 0, s>>>(p);
  *name = "rcas_href<FsrRcasH semantics>";
  return cudaGetLastError();
}

}  // namespace fsr1
