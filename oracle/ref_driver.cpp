// TEST INFRASTRUCTURE — not product code.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load the library built from this file.
//
// Host driver around the reference's OWN per-pixel source.  build_ref.sh produces, in a temp
// directory, a mechanically rewritten copy of /root/reference/ffx-fsr/{ffx_a.h,ffx_fsr1.h}
// (parameter qualifiers `in/out/inout` -> C++ value/reference; nothing else) and compiles this
// file against it with oracle/shim/hlsl_on_cpp.h standing in for the HLSL vector language.
// The statements executed per pixel are therefore the reference's own, in their own order:
//   FsrEasuF  ffx-fsr/ffx_fsr1.h:315-437     FsrRcasF  ffx-fsr/ffx_fsr1.h:684-769
//   FsrEasuH  ffx-fsr/ffx_fsr1.h:505-593     FsrRcasH  ffx-fsr/ffx_fsr1.h:782-866
// The callbacks below mirror the sample's shader wrapper:
//   gather4 + linear/clamp sampler   sample/src/DX12/FSR_Pass.hlsl:39-41,55-57, FSR_Filter.cpp:48-53
//   integer Load (OOB -> 0 in D3D12) sample/src/DX12/FSR_Pass.hlsl:45-46,61-62
//   store float4(c,1)                sample/src/DX12/FSR_Pass.hlsl:80,86,95,101
#include "hlsl_on_cpp.h"

// Variants: the RCAS compile-time options of the reference (ffx-fsr/ffx_fsr1.h:647-651) are selected with
//   -DFSR_RCAS_DENOISE=1 and/or -DFSR_RCAS_PASSTHROUGH_ALPHA=1, together with -DREFNS=<namespace> -DREFSUF=<suffix>;
// each variant lives in its own namespace and exports only fsr1ref_rcas_f<suffix> / fsr1ref_rcas_h<suffix>.
#ifndef REFNS
#define REFNS fsr1ref_base
#define REF_BASE 1
#endif
#ifndef REFSUF
#define REFSUF
#endif
#define REF_CAT2(a, b) a##b
#define REF_CAT(a, b) REF_CAT2(a, b)
#define REF_NAME(n) REF_CAT(n, REFSUF)

namespace REFNS {

#define A_GPU 1
#define A_HLSL 1
#ifdef FSR1_REF_HALF
#define A_HALF 1
#endif
#include "ffx_a.h"  // rewritten copy, found through -I<tmpdir>

struct ImgF { const float* p; int w, h; size_t pitch; };        // RGBA32F, pitch in floats
struct ImgH { const uint16_t* p; int w, h; size_t pitch; };     // RGBA16F, pitch in halves
static thread_local ImgF g_f;
static thread_local int g_rcas_clamp;
#ifdef FSR1_REF_HALF
static thread_local ImgH g_h;
#endif

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// gather4 of channel c at normalised p: (x,y,z,w) = texels (i,j+1),(i+1,j+1),(i+1,j),(i,j)
static inline AF4 gatherF(AF2 p, int c) {
  int i = (int)floorf(p.x * (float)g_f.w - 0.5f), j = (int)floorf(p.y * (float)g_f.h - 0.5f);
  int x0 = clampi(i, 0, g_f.w - 1), x1 = clampi(i + 1, 0, g_f.w - 1);
  int y0 = clampi(j, 0, g_f.h - 1), y1 = clampi(j + 1, 0, g_f.h - 1);
  const float* b = g_f.p;
  return AF4(b[y1 * g_f.pitch + x0 * 4 + c], b[y1 * g_f.pitch + x1 * 4 + c],
             b[y0 * g_f.pitch + x1 * 4 + c], b[y0 * g_f.pitch + x0 * 4 + c]);
}
#define FSR_EASU_F 1
AF4 FsrEasuRF(AF2 p) { return gatherF(p, 0); }
AF4 FsrEasuGF(AF2 p) { return gatherF(p, 1); }
AF4 FsrEasuBF(AF2 p) { return gatherF(p, 2); }
#define FSR_RCAS_F 1
AF4 FsrRcasLoadF(ASU2 p) {
  int x = p.x, y = p.y;
  if (g_rcas_clamp) { x = clampi(x, 0, g_f.w - 1); y = clampi(y, 0, g_f.h - 1); }
  else if (x < 0 || y < 0 || x >= g_f.w || y >= g_f.h) return AF4(0.0f, 0.0f, 0.0f, 0.0f);
  const float* t = g_f.p + y * g_f.pitch + x * 4;
  return AF4(t[0], t[1], t[2], t[3]);
}
void FsrRcasInputF(AF1& r, AF1& g, AF1& b) {}

#ifdef FSR1_REF_HALF
static inline _Float16 h_from_bits(uint16_t b) { _Float16 h; memcpy(&h, &b, 2); return h; }
static inline uint16_t h_to_bits(_Float16 h) { uint16_t b; memcpy(&b, &h, 2); return b; }
static inline AH4 gatherH(AF2 p, int c) {
  int i = (int)floorf(p.x * (float)g_h.w - 0.5f), j = (int)floorf(p.y * (float)g_h.h - 0.5f);
  int x0 = clampi(i, 0, g_h.w - 1), x1 = clampi(i + 1, 0, g_h.w - 1);
  int y0 = clampi(j, 0, g_h.h - 1), y1 = clampi(j + 1, 0, g_h.h - 1);
  const uint16_t* b = g_h.p;
  return AH4(h_from_bits(b[y1 * g_h.pitch + x0 * 4 + c]), h_from_bits(b[y1 * g_h.pitch + x1 * 4 + c]),
             h_from_bits(b[y0 * g_h.pitch + x1 * 4 + c]), h_from_bits(b[y0 * g_h.pitch + x0 * 4 + c]));
}
#define FSR_EASU_H 1
AH4 FsrEasuRH(AF2 p) { return gatherH(p, 0); }
AH4 FsrEasuGH(AF2 p) { return gatherH(p, 1); }
AH4 FsrEasuBH(AF2 p) { return gatherH(p, 2); }
#define FSR_RCAS_H 1
AH4 FsrRcasLoadH(ASW2 p) {
  int x = p.x, y = p.y;
  if (g_rcas_clamp) { x = clampi(x, 0, g_h.w - 1); y = clampi(y, 0, g_h.h - 1); }
  else if (x < 0 || y < 0 || x >= g_h.w || y >= g_h.h) return AH4(0.0f, 0.0f, 0.0f, 0.0f);
  const uint16_t* t = g_h.p + y * g_h.pitch + x * 4;
  return AH4(h_from_bits(t[0]), h_from_bits(t[1]), h_from_bits(t[2]), h_from_bits(t[3]));
}
void FsrRcasInputH(AH1& r, AH1& g, AH1& b) {}
// the packed calling convention (ffx-fsr/ffx_fsr1.h:874-984): same loads, two pixels (ip, ip+(8,0)) per call
#define FSR_RCAS_HX2 1
AH4 FsrRcasLoadHx2(ASW2 p) { return FsrRcasLoadH(p); }
void FsrRcasInputHx2(AH2& r, AH2& g, AH2& b) {}
#endif

#include "ffx_fsr1.h"  // rewritten copy

}  // namespace REFNS
using namespace REFNS;

extern "C" {

#ifdef REF_BASE
// Constants as the reference computes them in its A_GPU build (FsrRcasCon's packed half uses
// round-to-nearest f32tof16 here; the A_CPU build truncates — see ref_cpu.c for that one).
void fsr1ref_gpu_easu_con(uint32_t* con /*16*/, float inVpW, float inVpH, float inW, float inH, float outW, float outH) {
  AU4 c0, c1, c2, c3;
  FsrEasuCon(c0, c1, c2, c3, inVpW, inVpH, inW, inH, outW, outH);
  for (int i = 0; i < 4; i++) { con[i] = c0[i]; con[4 + i] = c1[i]; con[8 + i] = c2[i]; con[12 + i] = c3[i]; }
}
void fsr1ref_gpu_rcas_con(uint32_t* con /*4*/, float sharpness) {
  AU4 c; FsrRcasCon(c, sharpness);
  for (int i = 0; i < 4; i++) con[i] = c[i];
}

// EASU fp32 over output rows [y0,y1).  Images are RGBA32F, pitches in floats.
void fsr1ref_easu_f(const float* in, int inW, int inH, size_t inPitch, float* out, int outW, int outH,
                    size_t outPitch, const uint32_t* con, int y0, int y1) {
  AU4 c0(con[0], con[1], con[2], con[3]), c1(con[4], con[5], con[6], con[7]);
  AU4 c2(con[8], con[9], con[10], con[11]), c3(con[12], con[13], con[14], con[15]);
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = y0; y < y1; y++) {
    g_f = ImgF{in, inW, inH, inPitch};
    for (int x = 0; x < outW; x++) {
      AF3 pix;
      FsrEasuF(pix, AU2((uint)x, (uint)y), c0, c1, c2, c3);
      float* o = out + (size_t)y * outPitch + (size_t)x * 4;
      o[0] = pix.r; o[1] = pix.g; o[2] = pix.b; o[3] = 1.0f;
    }
  }
}


// ---- the pointwise companions of the scaling path, fp32 ------------------------------------------------
//   FsrLfgaF ffx-fsr/ffx_fsr1.h:1014   FsrSrtmF/FsrSrtmInvF :1044,1046   FsrTepdDitF :1086-1095   FsrTepdC8F/C10F :1100-1126
// Pixels are n interleaved RGBA32F quadruples; alpha is not touched (the functions take AF3).
void fsr1ref_lfga_f(float* c, const float* t, size_t n, float amount) {
  for (size_t i = 0; i < n; i++) {
    AF3 v(c[4 * i], c[4 * i + 1], c[4 * i + 2]);
    FsrLfgaF(v, AF3(t[4 * i], t[4 * i + 1], t[4 * i + 2]), amount);
    c[4 * i] = v.r; c[4 * i + 1] = v.g; c[4 * i + 2] = v.b;
  }
}
void fsr1ref_srtm_f(float* c, size_t n, int inverse) {
  for (size_t i = 0; i < n; i++) {
    AF3 v(c[4 * i], c[4 * i + 1], c[4 * i + 2]);
    if (inverse) FsrSrtmInvF(v); else FsrSrtmF(v);
    c[4 * i] = v.r; c[4 * i + 1] = v.g; c[4 * i + 2] = v.b;
  }
}
void fsr1ref_tepd_dit_f(float* dit, int w, int h, uint32_t frame) {
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) dit[(size_t)y * w + x] = FsrTepdDitF(AU2((uint)x, (uint)y), (AU1)frame);
}
void fsr1ref_tepd_f(float* c, const float* dit, size_t n, int bits) {
  for (size_t i = 0; i < n; i++) {
    AF3 v(c[4 * i], c[4 * i + 1], c[4 * i + 2]);
    if (bits == 8) FsrTepdC8F(v, dit[i]); else FsrTepdC10F(v, dit[i]);
    c[4 * i] = v.r; c[4 * i + 1] = v.g; c[4 * i + 2] = v.b;
  }
}

#endif  // REF_BASE

// RCAS fp32 over rows [y0,y1); oob_clamp=0 -> D3D12 Load semantics (OOB reads 0), 1 -> clamp.
void REF_NAME(fsr1ref_rcas_f)(const float* in, int W, int H, size_t inPitch, float* out, size_t outPitch,
                    const uint32_t* con, int oob_clamp, int y0, int y1) {
  AU4 c(con[0], con[1], con[2], con[3]);
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = y0; y < y1; y++) {
    g_f = ImgF{in, W, H, inPitch}; g_rcas_clamp = oob_clamp;
    for (int x = 0; x < W; x++) {
      AF1 r, g, b, a = 1.0f;
#ifdef FSR_RCAS_PASSTHROUGH_ALPHA
      FsrRcasF(r, g, b, a, AU2((uint)x, (uint)y), c);
#else
      FsrRcasF(r, g, b, AU2((uint)x, (uint)y), c);
#endif
      float* o = out + (size_t)y * outPitch + (size_t)x * 4;
      o[0] = r; o[1] = g; o[2] = b; o[3] = a;
    }
  }
}

#ifdef FSR1_REF_HALF
// The packed-half variants: images are RGBA16F (raw half bits), pitches in halves.
#ifdef REF_BASE
void fsr1ref_easu_h(const uint16_t* in, int inW, int inH, size_t inPitch, uint16_t* out, int outW, int outH,
                    size_t outPitch, const uint32_t* con, int y0, int y1) {
  AU4 c0(con[0], con[1], con[2], con[3]), c1(con[4], con[5], con[6], con[7]);
  AU4 c2(con[8], con[9], con[10], con[11]), c3(con[12], con[13], con[14], con[15]);
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = y0; y < y1; y++) {
    g_h = ImgH{in, inW, inH, inPitch};
    for (int x = 0; x < outW; x++) {
      AH3 pix;
      FsrEasuH(pix, AU2((uint)x, (uint)y), c0, c1, c2, c3);
      uint16_t* o = out + (size_t)y * outPitch + (size_t)x * 4;
      o[0] = h_to_bits(pix.r); o[1] = h_to_bits(pix.g); o[2] = h_to_bits(pix.b); o[3] = 0x3c00;
    }
  }
}
#endif  // REF_BASE
void REF_NAME(fsr1ref_rcas_h)(const uint16_t* in, int W, int H, size_t inPitch, uint16_t* out, size_t outPitch,
                    const uint32_t* con, int oob_clamp, int y0, int y1) {
  AU4 c(con[0], con[1], con[2], con[3]);
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = y0; y < y1; y++) {
    g_h = ImgH{in, W, H, inPitch}; g_rcas_clamp = oob_clamp;
    for (int x = 0; x < W; x++) {
      AH1 r, g, b, a = (AH1)1.0;
#ifdef FSR_RCAS_PASSTHROUGH_ALPHA
      FsrRcasH(r, g, b, a, AU2((uint)x, (uint)y), c);
#else
      FsrRcasH(r, g, b, AU2((uint)x, (uint)y), c);
#endif
      uint16_t* o = out + (size_t)y * outPitch + (size_t)x * 4;
      o[0] = h_to_bits(r); o[1] = h_to_bits(g); o[2] = h_to_bits(b); o[3] = h_to_bits(a);
    }
  }
}
// FsrRcasHx2 (ffx-fsr/ffx_fsr1.h:888-984): lane i of a 16x1 strip produces pixels x0+i and x0+i+8.
void REF_NAME(fsr1ref_rcas_hx2)(const uint16_t* in, int W, int H, size_t inPitch, uint16_t* out, size_t outPitch,
                      const uint32_t* con, int oob_clamp, int y0, int y1) {
  AU4 c(con[0], con[1], con[2], con[3]);
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = y0; y < y1; y++) {
    g_h = ImgH{in, W, H, inPitch}; g_rcas_clamp = oob_clamp;
    for (int x0 = 0; x0 < W; x0 += 16)
      for (int i = 0; i < 8; i++) {
        AH2 pR, pG, pB, pA((AH1)1.0, (AH1)1.0);
#ifdef FSR_RCAS_PASSTHROUGH_ALPHA
        FsrRcasHx2(pR, pG, pB, pA, AU2((uint)(x0 + i), (uint)y), c);
#else
        FsrRcasHx2(pR, pG, pB, AU2((uint)(x0 + i), (uint)y), c);
#endif
        AH4 q0, q1;
        FsrRcasDepackHx2(q0, q1, pR, pG, pB);
        if (x0 + i < W) {
          uint16_t* o = out + (size_t)y * outPitch + (size_t)(x0 + i) * 4;
          o[0] = h_to_bits(q0.r); o[1] = h_to_bits(q0.g); o[2] = h_to_bits(q0.b); o[3] = h_to_bits(pA.x);
        }
        if (x0 + i + 8 < W) {
          uint16_t* o = out + (size_t)y * outPitch + (size_t)(x0 + i + 8) * 4;
          o[0] = h_to_bits(q1.r); o[1] = h_to_bits(q1.g); o[2] = h_to_bits(q1.b); o[3] = h_to_bits(pA.y);
        }
      }
  }
}
#ifdef REF_BASE
// ---- the pointwise companions in half precision: the scalar H functions and the packed Hx2 calling convention --------
//   FsrLfgaH / FsrLfgaHx2 ffx-fsr/ffx_fsr1.h:1019-1024   FsrSrtmH / InvH / Hx2 / InvHx2 :1049-1056
//   FsrTepdDitH :1129-1135   FsrTepdC8H / C10H :1137-1153   FsrTepdDitHx2 :1156-1164   FsrTepdC8Hx2 / C10Hx2 :1166-1199
// Pixels are n interleaved RGBA16F quadruples (raw half bits); alpha is not touched.  The Hx2 entry points pack pixels
// (2k, 2k+1) into the two lanes (n odd: the last pixel is paired with itself).
static inline AH3 h3_at(const uint16_t* c, size_t i) { return AH3(h_from_bits(c[4 * i]), h_from_bits(c[4 * i + 1]), h_from_bits(c[4 * i + 2])); }
static inline void h3_to(uint16_t* c, size_t i, AH3 v) { c[4 * i] = h_to_bits(v.r); c[4 * i + 1] = h_to_bits(v.g); c[4 * i + 2] = h_to_bits(v.b); }
struct PairH { AH2 r, g, b; };
static inline PairH pair_at(const uint16_t* c, size_t i, size_t j) {
  return PairH{AH2(h_from_bits(c[4 * i]), h_from_bits(c[4 * j])), AH2(h_from_bits(c[4 * i + 1]), h_from_bits(c[4 * j + 1])),
               AH2(h_from_bits(c[4 * i + 2]), h_from_bits(c[4 * j + 2]))};
}
static inline void pair_to(uint16_t* c, size_t i, size_t j, const PairH& v) {
  c[4 * j] = h_to_bits(v.r.y); c[4 * j + 1] = h_to_bits(v.g.y); c[4 * j + 2] = h_to_bits(v.b.y);
  c[4 * i] = h_to_bits(v.r.x); c[4 * i + 1] = h_to_bits(v.g.x); c[4 * i + 2] = h_to_bits(v.b.x);
}
void fsr1ref_lfga_h(uint16_t* c, const uint16_t* t, size_t n, float amount) {
  for (size_t i = 0; i < n; i++) { AH3 v = h3_at(c, i); FsrLfgaH(v, h3_at(t, i), (AH1)amount); h3_to(c, i, v); }
}
void fsr1ref_lfga_hx2(uint16_t* c, const uint16_t* t, size_t n, float amount) {
  for (size_t i = 0; i < n; i += 2) {
    const size_t j = i + 1 < n ? i + 1 : i;
    PairH v = pair_at(c, i, j), g = pair_at(t, i, j);
    FsrLfgaHx2(v.r, v.g, v.b, g.r, g.g, g.b, (AH1)amount);
    pair_to(c, i, j, v);
  }
}
void fsr1ref_srtm_h(uint16_t* c, size_t n, int inverse) {
  for (size_t i = 0; i < n; i++) { AH3 v = h3_at(c, i); if (inverse) FsrSrtmInvH(v); else FsrSrtmH(v); h3_to(c, i, v); }
}
void fsr1ref_srtm_hx2(uint16_t* c, size_t n, int inverse) {
  for (size_t i = 0; i < n; i += 2) {
    const size_t j = i + 1 < n ? i + 1 : i;
    PairH v = pair_at(c, i, j);
    if (inverse) FsrSrtmInvHx2(v.r, v.g, v.b); else FsrSrtmHx2(v.r, v.g, v.b);
    pair_to(c, i, j, v);
  }
}
void fsr1ref_tepd_dit_h(uint16_t* dit, int w, int h, uint32_t frame) {
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) dit[(size_t)y * w + x] = h_to_bits(FsrTepdDitH(AU2((uint)x, (uint)y), (AU1)frame));
}
// FsrTepdDitHx2 produces the values of positions p and p+(8,0): columns [16k, 16k+8) are lane x, [16k+8, 16k+16) lane y
void fsr1ref_tepd_dit_hx2(uint16_t* dit, int w, int h, uint32_t frame) {
  for (int y = 0; y < h; y++)
    for (int x0 = 0; x0 < w; x0 += 16)
      for (int i = 0; i < 8; i++) {
        const AH2 d = FsrTepdDitHx2(AU2((uint)(x0 + i), (uint)y), (AU1)frame);
        if (x0 + i < w) dit[(size_t)y * w + x0 + i] = h_to_bits(d.x);
        if (x0 + i + 8 < w) dit[(size_t)y * w + x0 + i + 8] = h_to_bits(d.y);
      }
}
void fsr1ref_tepd_h(uint16_t* c, const uint16_t* dit, size_t n, int bits) {
  for (size_t i = 0; i < n; i++) {
    AH3 v = h3_at(c, i);
    if (bits == 8) FsrTepdC8H(v, h_from_bits(dit[i])); else FsrTepdC10H(v, h_from_bits(dit[i]));
    h3_to(c, i, v);
  }
}
void fsr1ref_tepd_hx2(uint16_t* c, const uint16_t* dit, size_t n, int bits) {
  for (size_t i = 0; i < n; i += 2) {
    const size_t j = i + 1 < n ? i + 1 : i;
    PairH v = pair_at(c, i, j);
    const AH2 d(h_from_bits(dit[i]), h_from_bits(dit[j]));
    if (bits == 8) FsrTepdC8Hx2(v.r, v.g, v.b, d); else FsrTepdC10Hx2(v.r, v.g, v.b, d);
    pair_to(c, i, j, v);
  }
}
int fsr1ref_has_half(void) { return 1; }
#endif
#else
#ifdef REF_BASE
int fsr1ref_has_half(void) { return 0; }
#endif
#endif

}  // extern "C"
