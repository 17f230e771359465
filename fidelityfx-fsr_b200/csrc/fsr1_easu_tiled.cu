// fsr1_easu_tiled.cu — the production EASU kernels for RGBA16F images on sm_100a.
//
// Common structure.  A CTA produces a tile of the output.  The tile's input footprint (plus the 4x4
// tap window's halo) is fetched by ONE TMA 2D tile load (cp.async.bulk.tensor.2d, elected thread,
// mbarrier complete_tx) into shared memory; parts of the box outside the image arrive as zeros and are
// rewritten to clamp-to-edge (the reference samples through a CLAMP sampler,
// sample/src/DX12/FSR_Filter.cpp:48-53).  The arithmetic is then split the way it factors:
//
//   phase 1  per INPUT texel   2*luma, fp32                                         (ffx_fsr1.h:363-366)
//   phase 2  per INPUT texel   the FsrEasuSetF terms that do not depend on the output pixel:
//                              dirX, dirY, lenX^2+lenY^2 (fp32, F-path bit tricks)    (ffx_fsr1.h:295-313)
//   phase 3  per OUTPUT pixel  fp32: bilinear blend of the 4 nearest texels' terms (:383-386), normalise,
//                              stretch, lobe, clip (:389-409);  packed half2 over TWO output pixels:
//                              the 12 taps (:423-434) and the de-ringing clamp (:416-419,437).
//
// Precision split (DESIGN.md "numerics"): everything that decides the filter's ORIENTATION is fp32 — in
// half it is ill-conditioned where gradients nearly cancel and differs from the fp32 algorithm by up to
// 0.1; the taps (the bulk of the arithmetic) are half2 and stay within ~5e-3 of the fp32 oracle.
// Issue model (tools/ubench_pipes.cu, profiles/r01_ubench_pipes.txt): on B200 HFMA2, FFMA2 and scalar FFMA all
// sustain ~2 warp-instructions per cycle per SM and mixed streams 2.1-2.6, so half2 buys registers and shared-memory
// bytes, not issue slots; this kernel runs at 2.53 inst/cycle/SM, i.e. instruction count is what is left to cut.
//
// Tap weights use the expanded quadratic form of the rotated, anisotropically scaled distance
//   d2(ox,oy) = qa*ox^2 + qb*ox*oy + qc*oy^2,  qa = l2x^2 dx^2 + l2y^2 dy^2,  qc = l2x^2 dy^2 + l2y^2 dx^2,
//   qb = 2 dx dy (l2x^2 - l2y^2)
// and the window polynomial (25/16 (2/5 d2 - 1)^2 - 9/16)(lob d2 - 1)^2 = ((d2/4 - 5/4) d2 + 1)(lob d2 - 1)^2.
//
// Two kernels:
//   easu_h_pairs_kernel   any scale; 64x16 output tile per CTA; lane = 2 horizontally adjacent pixels.
//   easu_h_quad2x_kernel  exactly 2x (con0 = {.5,.5,-.25,-.25}, BASELINE configs[1]): the four output pixels
//                         (2k+1,2k+2)x(2m+1,2m+2) share one 4x4 window; a lane owns that quad, loads the 12
//                         taps and the 4 term vectors once, and every tap offset is a compile-time constant.
//                         Persistent CTAs, TMA double-buffered: tile i+1 loads while tile i computes.
#include "fsr1_easu_common.cuh"

namespace fsr1 {

constexpr int kThreads = 256;

// Zero-filled out-of-image texels of a TMA box -> clamp-to-edge.  Sources are always in-image positions
// (never rewritten), destinations always out-of-image ones (never read), so no intermediate barrier.
__device__ __forceinline__ void clamp_fixup(uint2* tile, int BW, int BH, int gx0, int gy0, int W, int H, int lane,
                                            int warp) {
  for (int j = warp; j < BH; j += kThreads / 32) {
    const int cy = clampi(gy0 + j, 0, H - 1) - gy0;
    for (int i = lane; i < BW; i += 32) {
      const int cx = clampi(gx0 + i, 0, W - 1) - gx0;
      if ((cx != i || cy != j) && cx >= 0 && cx < BW && cy >= 0 && cy < BH) tile[j * BW + i] = tile[cy * BW + cx];
    }
  }
}

__device__ __forceinline__ float texel_luma(uint2 t) {  // 2*luma = 0.5 B + (0.5 R + G); exact in fp32
  const float2 rg = __half22float2(u2h2(t.x));
  return fmaf(__low2float(u2h2(t.y)), 0.5f, fmaf(rg.x, 0.5f, rg.y));
}

// window weight of (up to) two pixels at squared distance d2
__device__ __forceinline__ __half2 tap_weight(__half2 d2, __half2 lob, __half2 clp) {
  d2 = __hmin2(d2, clp);
  const __half2 wb = __hfma2(__hfma2(h2c(0.25f), d2, h2c(-1.25f)), d2, h2c(1.0f));
  __half2 wa = __hfma2(lob, d2, h2c(-1.0f));
  wa = __hmul2(wa, wa);
  return __hmul2(wb, wa);
}

// experimental (FSR1_EASU_QUAD_VARIANT=8, unmeasured): the clamp as a packed 16-bit INTEGER min (VIMNMX.S16x2, integer
// ALU pipe — HMNMX2 does not overlap with HFMA2 in the microbenchmark, integer ops do).  clp > 0, so comparing the bit
// patterns as signed 16-bit integers orders every non-negative d2 correctly and returns d2 itself when rounding made
// it slightly negative: identical results to __hmin2 for finite inputs.
__device__ __forceinline__ __half2 tap_weight_iclamp(__half2 d2, __half2 lob, __half2 clp) {
#ifdef FSR1_CPU_EMU
  const uint32_t r = emu_min_s16x2(h22u(d2), h22u(clp));
#else
  uint32_t r;
  asm("min.s16x2 %0, %1, %2;" : "=r"(r) : "r"(h22u(d2)), "r"(h22u(clp)));
#endif
  d2 = u2h2(r);
  const __half2 wb = __hfma2(__hfma2(h2c(0.25f), d2, h2c(-1.25f)), d2, h2c(1.0f));
  __half2 wa = __hfma2(lob, d2, h2c(-1.0f));
  wa = __hmul2(wa, wa);
  return __hmul2(wb, wa);
}

template <int kTap> __device__ __forceinline__ __half2 tap_weight_sel(__half2 d2, __half2 lob, __half2 clp) {
  if constexpr (kTap == 3) return tap_weight_iclamp(d2, lob, clp);
  else return tap_weight(d2, lob, clp);
}

// the same without the distance clamp, for taps that provably never reach it (see quad_pair<.., 1>)
__device__ __forceinline__ __half2 tap_weight_unclamped(__half2 d2, __half2 lob) {
  const __half2 wb = __hfma2(__hfma2(h2c(0.25f), d2, h2c(-1.25f)), d2, h2c(1.0f));
  __half2 wa = __hfma2(lob, d2, h2c(-1.0f));
  wa = __hmul2(wa, wa);
  return __hmul2(wb, wa);
}

// ---- experimental (FSR1_EASU_QUAD_VARIANT=7, not yet measured): the per-pixel fp32 analysis of a pixel PAIR in
// packed f32x2 (FFMA2 / FMUL2 / FADD2 issue at the scalar FFMA rate on B200, profiles/r01_ubench_pipes.txt, so this
// halves the fp32 FMA-class instructions of phase 3).  Lane .x = pixel A, .y = pixel B; the operations per lane are
// exactly those of pixel_shape().
__device__ __forceinline__ float2 mk2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 bc2(float a) { return make_float2(a, a); }
struct Shape2 { float2 qa, qb, qc, lob, clp; };
__device__ __forceinline__ Shape2 pixel_shape2(float2 dx, float2 dy, float2 len) {
  const float2 dirR = __ffma2_rn(dx, dx, __fmul2_rn(dy, dy));
  const bool zx = dirR.x < (1.0f / 32768.0f), zy = dirR.y < (1.0f / 32768.0f);
  const float2 rs = mk2(zx ? 1.0f : prx_lo_rsq(dirR.x), zy ? 1.0f : prx_lo_rsq(dirR.y));
  dx = __fmul2_rn(mk2(zx ? 1.0f : dx.x, zy ? 1.0f : dx.y), rs);
  dy = __fmul2_rn(dy, rs);
  len = __fmul2_rn(len, bc2(0.5f));
  len = __fmul2_rn(len, len);
  const float2 dx2 = __fmul2_rn(dx, dx), dy2 = __fmul2_rn(dy, dy);
  const float2 rmax = mk2(prx_lo_rcp(fmaxf(fabsf(dx.x), fabsf(dy.x))), prx_lo_rcp(fmaxf(fabsf(dx.y), fabsf(dy.y))));
  const float2 stretch = __fmul2_rn(__fadd2_rn(dx2, dy2), rmax);
  const float2 l2x = __ffma2_rn(__fadd2_rn(stretch, bc2(-1.0f)), len, bc2(1.0f));
  const float2 l2y = __ffma2_rn(bc2(-0.5f), len, bc2(1.0f));
  Shape2 s;
  s.lob = __ffma2_rn(bc2((float)((1.0 / 4.0 - 0.04) - 0.5)), len, bc2(0.5f));
  s.clp = mk2(prx_lo_rcp(s.lob.x), prx_lo_rcp(s.lob.y));
  const float2 X2 = __fmul2_rn(l2x, l2x), Y2 = __fmul2_rn(l2y, l2y);
  s.qa = __ffma2_rn(X2, dx2, __fmul2_rn(Y2, dy2));
  s.qc = __ffma2_rn(X2, dy2, __fmul2_rn(Y2, dx2));
  s.qb = __fmul2_rn(__fmul2_rn(__fmul2_rn(dx, dy), bc2(2.0f)), __ffma2_rn(Y2, bc2(-1.0f), X2));
  return s;
}
__device__ __forceinline__ Shape lane_x(const Shape2& s) { return Shape{s.qa.x, s.qb.x, s.qc.x, s.lob.x, s.clp.x}; }
__device__ __forceinline__ Shape lane_y(const Shape2& s) { return Shape{s.qa.y, s.qb.y, s.qc.y, s.lob.y, s.clp.y}; }

// =======================================================================================================
//  generic kernel: any scale, lane = pixel pair (2*lane, 2*lane+1), rows warp and warp+8 of a 64x16 tile
// =======================================================================================================
constexpr int kTileW = 64, kTileH = 32;

// One vertical pixel pair of the generic kernel: pixel A (row oy) and B (row oy+1) in the same output column.
// t0/q0 point at tap (0,0) / texel f of pixel A; DR = fy(B) - fy(A) in {0,1}.  Packed lanes are (A, B).
// kVar = 1 (experimental, FSR1_EASU_PAIRS_VARIANT=1, validated on the CPU emulator, not yet timed): the fp32 analysis of
// the pair packed in f32x2, the factored tap distance ox (qa ox + qb oy) + qc oy^2 and the integer distance clamp.
template <int DR, int kVar = 0>
__device__ __forceinline__ void vpair(const uint2* __restrict__ t0, const float4* __restrict__ q0, int BW, int SW, float ppx,
                                      float ppyA, float ppyB, uint2& outA, uint2& outB) {
  // fp32: blend of the f,g,j,k terms (reference order) and the filter shape, per pixel
  const float4 f = q0[0], g = q0[1], j = q0[SW], k = q0[SW + 1];
  const float ipx = 1.0f - ppx;
  Shape sA, sB;
  if constexpr (kVar == 1) {
    const float4 f2 = DR ? j : f, g2 = DR ? k : g, j2 = DR ? q0[2 * SW] : j, k2 = DR ? q0[2 * SW + 1] : k;
    const float2 ppy = mk2(ppyA, ppyB), ipy = __ffma2_rn(ppy, bc2(-1.0f), bc2(1.0f));
    const float2 wf = __fmul2_rn(bc2(ipx), ipy), wg = __fmul2_rn(bc2(ppx), ipy);
    const float2 wj = __fmul2_rn(bc2(ipx), ppy), wk = __fmul2_rn(bc2(ppx), ppy);
#define FSR1_BLEND2(C)                                                                                                  \
  __ffma2_rn(mk2(k.C, k2.C), wk, __ffma2_rn(mk2(j.C, j2.C), wj, __ffma2_rn(mk2(g.C, g2.C), wg, __fmul2_rn(mk2(f.C, f2.C), wf))))
    const Shape2 s2 = pixel_shape2(FSR1_BLEND2(x), FSR1_BLEND2(y), FSR1_BLEND2(z));
#undef FSR1_BLEND2
    sA = lane_x(s2);
    sB = lane_y(s2);
  } else {
  {
    const float ipy = 1.0f - ppyA, wf = ipx * ipy, wg = ppx * ipy, wj = ipx * ppyA, wk = ppx * ppyA;
    sA = pixel_shape(fmaf(k.x, wk, fmaf(j.x, wj, fmaf(g.x, wg, f.x * wf))), fmaf(k.y, wk, fmaf(j.y, wj, fmaf(g.y, wg, f.y * wf))),
                     fmaf(k.z, wk, fmaf(j.z, wj, fmaf(g.z, wg, f.z * wf))));
  }
  {
    const float4 f2 = DR ? j : f, g2 = DR ? k : g, j2 = DR ? q0[2 * SW] : j, k2 = DR ? q0[2 * SW + 1] : k;
    const float ipy = 1.0f - ppyB, wf = ipx * ipy, wg = ppx * ipy, wj = ipx * ppyB, wk = ppx * ppyB;
    sB = pixel_shape(fmaf(k2.x, wk, fmaf(j2.x, wj, fmaf(g2.x, wg, f2.x * wf))), fmaf(k2.y, wk, fmaf(j2.y, wj, fmaf(g2.y, wg, f2.y * wf))),
                     fmaf(k2.z, wk, fmaf(j2.z, wj, fmaf(g2.z, wg, f2.z * wf))));
  }
  }
  const __half2 qa = __floats2half2_rn(sA.qa, sB.qa), qb = __floats2half2_rn(sA.qb, sB.qb);
  const __half2 qc = __floats2half2_rn(sA.qc, sB.qc), lob = __floats2half2_rn(sA.lob, sB.lob);
  const __half2 clp = __floats2half2_rn(sA.clp, sB.clp);
  // d2(k,r) = PX[k] + QY[r] + SB[k]*OY[r]; the column offset is the same for both pixels, the row offset is not
  const __half2 ppy2 = __floats2half2_rn(ppyA, ppyB), ppx2 = __float2half2_rn(ppx);
  // kVar = 0: PX/SB per column, QY/OY per row.  kVar = 1: PX holds ox_K, SB is unused, OY holds qb*oy_R, QY qc*oy_R^2.
  __half2 PX[4], SB[4], QY[4], OY[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const __half2 oxk = __hsub2(h2c((float)(i - 1)), ppx2);
    if constexpr (kVar == 1) {
      const __half2 oyr = __hsub2(h2c((float)(i - 1)), ppy2);
      PX[i] = oxk;
      SB[i] = oxk;
      OY[i] = __hmul2(qb, oyr);
      QY[i] = __hmul2(__hmul2(qc, oyr), oyr);
    } else {
    SB[i] = __hmul2(qb, oxk);
    PX[i] = __hmul2(__hmul2(qa, oxk), oxk);
    OY[i] = __hsub2(h2c((float)(i - 1)), ppy2);
    QY[i] = __hmul2(__hmul2(qc, OY[i]), OY[i]);
    }
  }
  // squared tap distance of tap (R,K) and its window weight
  auto tapw = [&](int R, int K) -> __half2 {
    if constexpr (kVar == 1) return tap_weight_iclamp(__hfma2(__hfma2(qa, PX[K], OY[R]), PX[K], QY[R]), lob, clp);
    else return tap_weight(__hfma2(SB[K], OY[R], __hadd2(PX[K], QY[R])), lob, clp);
  };
  const __half2 kZero = h2c(0.0f), one = h2c(1.0f);
  if (DR == 0) {
    // same window for both pixels: colour accumulators in structure-of-arrays form (A,B) per channel
    __half2 aR = kZero, aG = kZero, aB = kZero, aW = kZero;
#define FSR1_VTAP0(R, K)                                                                     \
    {                                                                                        \
      const uint2 c = t0[(R) * BW + (K)];                                                    \
      const __half2 w = tapw(R, K);                                                          \
      aR = __hfma2(__low2half2(u2h2(c.x)), w, aR);                                           \
      aG = __hfma2(__high2half2(u2h2(c.x)), w, aG);                                          \
      aB = __hfma2(__low2half2(u2h2(c.y)), w, aB);                                           \
      aW = __hadd2(aW, w);                                                                   \
    }
    FSR1_VTAP0(0, 1) FSR1_VTAP0(0, 2) FSR1_VTAP0(1, 0) FSR1_VTAP0(1, 3)   // far taps first (see quad_pair)
    FSR1_VTAP0(2, 0) FSR1_VTAP0(2, 3) FSR1_VTAP0(3, 1) FSR1_VTAP0(3, 2)
    FSR1_VTAP0(1, 1) FSR1_VTAP0(1, 2) FSR1_VTAP0(2, 1) FSR1_VTAP0(2, 2)
#undef FSR1_VTAP0
    const uint2 cf = t0[BW + 1], cg = t0[BW + 2], cj = t0[2 * BW + 1], ck = t0[2 * BW + 2];
    const __half2 mnRG = __hmin2(__hmin2(u2h2(cf.x), u2h2(cg.x)), __hmin2(u2h2(cj.x), u2h2(ck.x)));
    const __half2 mxRG = __hmax2(__hmax2(u2h2(cf.x), u2h2(cg.x)), __hmax2(u2h2(cj.x), u2h2(ck.x)));
    const __half2 mnBA = __hmin2(__hmin2(u2h2(cf.y), u2h2(cg.y)), __hmin2(u2h2(cj.y), u2h2(ck.y)));
    const __half2 mxBA = __hmax2(__hmax2(u2h2(cf.y), u2h2(cg.y)), __hmax2(u2h2(cj.y), u2h2(ck.y)));
    const float2 aWf = __half22float2(aW);
    const __half2 r = __floats2half2_rn(rcp_approx(aWf.x), rcp_approx(aWf.y));
    const __half2 oR = __hmin2(__low2half2(mxRG), __hmax2(__low2half2(mnRG), __hmul2(aR, r)));
    const __half2 oG = __hmin2(__high2half2(mxRG), __hmax2(__high2half2(mnRG), __hmul2(aG, r)));
    const __half2 oB = __hmin2(__low2half2(mxBA), __hmax2(__low2half2(mnBA), __hmul2(aB, r)));
    outA = make_uint2(h22u(__lows2half2(oR, oG)), h22u(__lows2half2(oB, one)));
    outB = make_uint2(h22u(__highs2half2(oR, oG)), h22u(__highs2half2(oB, one)));
  } else {
    // pixel B's window is one input row further down: tap (R,K) of A is texel row R, of B texel row R+1
    __half2 aRG_A = kZero, aBA_A = kZero, aRG_B = kZero, aBA_B = kZero, aW = kZero;
#define FSR1_VTAP1(R, K)                                                                     \
    {                                                                                        \
      const uint2 ca = t0[(R) * BW + (K)], cb = t0[((R) + 1) * BW + (K)];                    \
      const __half2 w = tapw(R, K);                                                          \
      const __half2 wA2 = __low2half2(w), wB2 = __high2half2(w);                             \
      aRG_A = __hfma2(u2h2(ca.x), wA2, aRG_A);                                               \
      aBA_A = __hfma2(u2h2(ca.y), wA2, aBA_A);                                               \
      aRG_B = __hfma2(u2h2(cb.x), wB2, aRG_B);                                               \
      aBA_B = __hfma2(u2h2(cb.y), wB2, aBA_B);                                               \
      aW = __hadd2(aW, w);                                                                   \
    }
    FSR1_VTAP1(0, 1) FSR1_VTAP1(0, 2) FSR1_VTAP1(1, 0) FSR1_VTAP1(1, 3)
    FSR1_VTAP1(2, 0) FSR1_VTAP1(2, 3) FSR1_VTAP1(3, 1) FSR1_VTAP1(3, 2)
    FSR1_VTAP1(1, 1) FSR1_VTAP1(1, 2) FSR1_VTAP1(2, 1) FSR1_VTAP1(2, 2)
#undef FSR1_VTAP1
    const uint2 r1a = t0[BW + 1], r1b = t0[BW + 2], r2a = t0[2 * BW + 1], r2b = t0[2 * BW + 2];
    const uint2 r3a = t0[3 * BW + 1], r3b = t0[3 * BW + 2];
    const __half2 midMnRG = __hmin2(u2h2(r2a.x), u2h2(r2b.x)), midMxRG = __hmax2(u2h2(r2a.x), u2h2(r2b.x));
    const __half2 midMnBA = __hmin2(u2h2(r2a.y), u2h2(r2b.y)), midMxBA = __hmax2(u2h2(r2a.y), u2h2(r2b.y));
    const __half2 mnRG_A = __hmin2(__hmin2(u2h2(r1a.x), u2h2(r1b.x)), midMnRG), mxRG_A = __hmax2(__hmax2(u2h2(r1a.x), u2h2(r1b.x)), midMxRG);
    const __half2 mnBA_A = __hmin2(__hmin2(u2h2(r1a.y), u2h2(r1b.y)), midMnBA), mxBA_A = __hmax2(__hmax2(u2h2(r1a.y), u2h2(r1b.y)), midMxBA);
    const __half2 mnRG_B = __hmin2(__hmin2(u2h2(r3a.x), u2h2(r3b.x)), midMnRG), mxRG_B = __hmax2(__hmax2(u2h2(r3a.x), u2h2(r3b.x)), midMxRG);
    const __half2 mnBA_B = __hmin2(__hmin2(u2h2(r3a.y), u2h2(r3b.y)), midMnBA), mxBA_B = __hmax2(__hmax2(u2h2(r3a.y), u2h2(r3b.y)), midMxBA);
    const float2 aWf = __half22float2(aW);
    const __half2 rA = __float2half2_rn(rcp_approx(aWf.x)), rB = __float2half2_rn(rcp_approx(aWf.y));
    const __half2 oRG_A = __hmin2(mxRG_A, __hmax2(mnRG_A, __hmul2(aRG_A, rA)));
    const __half2 oBA_A = __hmin2(mxBA_A, __hmax2(mnBA_A, __hmul2(aBA_A, rA)));
    const __half2 oRG_B = __hmin2(mxRG_B, __hmax2(mnRG_B, __hmul2(aRG_B, rB)));
    const __half2 oBA_B = __hmin2(mxBA_B, __hmax2(mnBA_B, __hmul2(aBA_B, rB)));
    outA = make_uint2(h22u(oRG_A), h22u(__lows2half2(oBA_A, one)));   // alpha = 1 (FSR_Pass.hlsl:95)
    outB = make_uint2(h22u(oRG_B), h22u(__lows2half2(oBA_B, one)));
  }
}

// dynamic shared memory: [tile0][tile1][luma][terms][2 mbarriers], every part 128-byte aligned
__host__ __device__ inline size_t pairs_tile_stride(int BW, int BH) { return ((size_t)BW * BH * 8 + 127) & ~(size_t)127; }
__host__ __device__ inline size_t pairs_smem_bytes(int BW, int BH) {
  size_t off = 2 * pairs_tile_stride(BW, BH);
  off += ((size_t)BW * BH * 4 + 127) & ~(size_t)127;
  off += ((size_t)(BW - 2) * (BH - 2) * 16 + 127) & ~(size_t)127;
  return off + 16 + 128;  // + barriers + slack for the manual 128B alignment
}

template <int kVar>
__global__ void __launch_bounds__(kThreads, 3)
easu_h_pairs_kernel(const EasuParams p, const __grid_constant__ CUtensorMap tmap, const int BW, const int BH,
                    const int tiles_x, const int n_tiles) {
#ifdef FSR1_CPU_EMU
  unsigned char* smem_raw = fsr1_emu_dynamic_smem();
#else
  extern __shared__ unsigned char smem_raw[];
#endif
  // 128-byte align by OFFSET (pointer arithmetic on the shared array keeps the address space -> LDS/STS)
  unsigned char* base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  const int n = BW * BH;
  const size_t tstride = pairs_tile_stride(BW, BH);
  float* L = reinterpret_cast<float*>(base + 2 * tstride);
  float4* S = reinterpret_cast<float4*>(base + 2 * tstride + (((size_t)n * 4 + 127) & ~(size_t)127));
  uint64_t* bar = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(S) +
                                              (((size_t)(BW - 2) * (BH - 2) * 16 + 127) & ~(size_t)127));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  // box origin of tile t = first tap column/row of its first pixel; the column is rounded down to even because
  // TMA traps unless the box starts on a 16-byte boundary (2 texels)
  auto origin = [&](int t, int& ox0, int& oy0, int& fx0, int& fy0) {
    ox0 = (t % tiles_x) * kTileW;
    oy0 = p.y0 + (t / tiles_x) * kTileH;
    float dummy;
    easu_pos(ox0, p.c0x, p.c0z, fx0, dummy);
    easu_pos(oy0, p.c0y, p.c0w, fy0, dummy);
    fx0 = (fx0 - 1) & ~1;
    fy0 -= 1;
  };
  int t = blockIdx.x;
  if (tid == 0 && t < n_tiles) {
    int a, b, fx, fy;
    origin(t, a, b, fx, fy);
    mbar_expect_tx(&bar[0], (uint32_t)n * 8u);
    tma_load_2d(base, &tmap, fx, fy - p.in.row0, &bar[0]);
  }
  for (int it = 0; t < n_tiles; t += gridDim.x, it++) {
  const int bsel = it & 1;
  if (tid == 0 && t + (int)gridDim.x < n_tiles) {  // prefetch the next tile into the other buffer
    int a, b, fx, fy;
    origin(t + gridDim.x, a, b, fx, fy);
    fence_proxy_async();
    mbar_expect_tx(&bar[bsel ^ 1], (uint32_t)n * 8u);
    tma_load_2d(base + (bsel ^ 1) * tstride, &tmap, fx, fy - p.in.row0, &bar[bsel ^ 1]);
  }
  uint2* tile = reinterpret_cast<uint2*>(base + bsel * tstride);
  int ox0, oy0, fx0, fy0;
  origin(t, ox0, oy0, fx0, fy0);
  mbar_wait(&bar[bsel], (it >> 1) & 1);

  if (fx0 < 0 || fy0 < 0 || fx0 + BW > p.in.w || fy0 + BH > p.in.h) {  // border tiles only (CTA-uniform)
    clamp_fixup(tile, BW, BH, fx0, fy0, p.in.w, p.in.h, lane, warp);
    fence_proxy_async();
    __syncthreads();
  }

  for (int i = tid; i < n; i += kThreads) L[i] = texel_luma(tile[i]);  // phase 1
  __syncthreads();

  // phase 2 over the (BW-2)x(BH-2) inner texels, flattened so that all lanes stay busy
  const int SW = BW - 2, nS = SW * (BH - 2);
  {
    int j = tid / SW, i = tid - j * SW;
    const int dj = kThreads / SW, di = kThreads - dj * SW;
    for (int idx = tid; idx < nS; idx += kThreads) {
      const float* c = L + (j + 1) * BW + (i + 1);
      S[idx] = texel_terms(c[-BW], c[-1], c[0], c[1], c[BW]);
      i += di; j += dj;
      if (i >= SW) { i -= SW; j += 1; }
    }
  }
  __syncthreads();

  // phase 3: a lane owns one output column and the VERTICAL pixel pair (oy, oy+1).  Whether the two rows fall in
  // the same input cell row (DR = 0) or in consecutive ones (DR = 1) depends on oy only, so it is warp-uniform:
  // DR = 0 loads 12 taps + 4 term vectors once for both pixels, DR = 1 loads a 5-row window (16 + 6).  Lanes walk
  // the input row at < 1 texel per lane: shared-memory reads are (nearly) conflict-free, which the earlier
  // horizontal pairing (1.3-2 texels per lane, two windows per lane) was not — it was LSU-bound (ncu: 72 %).
#pragma unroll 1
  for (int job = warp; job < (kTileW / 32) * (kTileH / 2); job += kThreads / 32) {
    const int oyA = oy0 + (job >> 1) * 2;
    if (oyA >= p.y1) continue;  // warp-uniform
    const bool hasB = oyA + 1 < p.y1;
    const int oxr = ox0 + (job & 1) * 32 + lane;
    const bool active = oxr < p.out.w;
    const int ox = active ? oxr : p.out.w - 1;
    int fx, fyA, fyB;
    float ppx, ppyA, ppyB;
    easu_pos(ox, p.c0x, p.c0z, fx, ppx);
    easu_pos(oyA, p.c0y, p.c0w, fyA, ppyA);
    easu_pos(hasB ? oyA + 1 : oyA, p.c0y, p.c0w, fyB, ppyB);
    const uint2* t0 = tile + (fyA - fy0 - 1) * BW + (fx - fx0 - 1);       // window origin: tap (0,0) of pixel A
    const float4* q0 = S + (fyA - fy0 - 1) * SW + (fx - fx0 - 1);          // term vector of texel f of pixel A
    uint2 oA, oB;
    if (fyB == fyA) vpair<0, kVar>(t0, q0, BW, SW, ppx, ppyA, ppyB, oA, oB);
    else vpair<1, kVar>(t0, q0, BW, SW, ppx, ppyA, ppyB, oA, oB);
    if (active) {
      unsigned char* o = p.out.base + (long long)(oyA - p.out.row0) * p.out.pitch + (long long)ox * 8;
      *reinterpret_cast<uint2*>(o) = oA;
      if (hasB) *reinterpret_cast<uint2*>(o + p.out.pitch) = oB;
    }
  }
  __syncthreads();  // L, S and this tile buffer are free again
  }  // persistent tile loop
}

// =======================================================================================================
//  2x kernel: lane = the quad of output pixels sharing input cell (k,m); persistent, double-buffered TMA
// =======================================================================================================
constexpr int kQCX = 32;        // cells per tile in x (= 64 output pixels); one lane per cell
constexpr int kQBW = kQCX + 4;  // TMA box width: 36 texels (35 needed, even width)
constexpr int kQSW = kQBW - 2;  // inner texels carrying terms: 34 per row
// NW warps per CTA, each warp owns 2 cell rows: cells per tile 32 x 2NW, box 36 x (2NW+3), terms 34 x (2NW+1)
template <int NW> struct QuadCfg {
  static constexpr int kCY = 2 * NW, kBH = kCY + 3, kSH = kBH - 2, kElems = kQBW * kBH;
  static constexpr int kPad = ((kElems * 8 + 127) / 128) * 128 / 8;  // buffer stride keeping 128B alignment
};

// One pixel pair (A: px=.25, B: px=.75) of the quad; kBottom selects py=.75.  t = the 12 taps (RG,BA);
// every tap offset is a constant, so d2 is three half2 FMAs against immediates.
// kTap = 1 (default; kTap = 0 is FSR1_EASU_QUAD_VARIANT=2, 3 % slower on B200): rows 1 and 2 (four taps each) use the factored form
// d2 = ox (qa ox + qb oy) + qc oy^2 with the row terms qb oy, qc oy^2 hoisted (10 instead of 12 half2 ops per row), and
// the four nearest taps f g j k skip the min(d2, clp): at exactly 2x their offsets are <= .75 per axis, so
// d2 <= 1.125 len2.x^2 (1 + eps) < clp = 1/lob for every len in [0,1] (1.24 (1 + .56 len)^2 vs .94 / (.5 - .29 len)).
template <bool kBottom, int kTap>
__device__ __forceinline__ void quad_pair(const uint2 (&t)[4][4], const Shape& sA, const Shape& sB, __half2 mnR,
                                          __half2 mnG, __half2 mnB, __half2 mxR, __half2 mxG, __half2 mxB,
                                          uint2& outA, uint2& outB) {
  const __half2 qa = __floats2half2_rn(sA.qa, sB.qa), qb = __floats2half2_rn(sA.qb, sB.qb);
  const __half2 qc = __floats2half2_rn(sA.qc, sB.qc), lob = __floats2half2_rn(sA.lob, sB.lob);
  const __half2 clp = __floats2half2_rn(sA.clp, sB.clp);
  const __half2 kZero = h2c(0.0f);
  __half2 aR = kZero, aG = kZero, aB = kZero, aW = kZero;
  constexpr float py = kBottom ? 0.75f : 0.25f;
#define FSR1_QTAP(R, K)                                                                                     \
  {                                                                                                         \
    constexpr float oxA = (float)((K)-1) - 0.25f, oxB = (float)((K)-1) - 0.75f, oy = (float)((R)-1) - py;     \
    const __half2 d2 = __hfma2(qa, __floats2half2_rn(oxA * oxA, oxB * oxB),                                  \
                               __hfma2(qc, __floats2half2_rn(oy * oy, oy * oy),                              \
                                       __hmul2(qb, __floats2half2_rn(oxA * oy, oxB * oy))));                 \
    const __half2 w = tap_weight_sel<kTap>(d2, lob, clp);                                                   \
    const __half2 rg = u2h2(t[R][K].x), ba = u2h2(t[R][K].y);                                               \
    aR = __hfma2(__low2half2(rg), w, aR);                                                                   \
    aG = __hfma2(__high2half2(rg), w, aG);                                                                  \
    aB = __hfma2(__low2half2(ba), w, aB);                                                                   \
    aW = __hadd2(aW, w);                                                                                    \
  }
#define FSR1_QTAP_ROW(R, K, INNER)                                                                          \
  {                                                                                                         \
    constexpr float oxA = (float)((K)-1) - 0.25f, oxB = (float)((K)-1) - 0.75f;                               \
    const __half2 ox = __floats2half2_rn(oxA, oxB);                                                         \
    const __half2 d2 = __hfma2(__hfma2(qa, ox, rowB##R), ox, rowC##R);                                       \
    const __half2 w = (INNER) ? tap_weight_unclamped(d2, lob) : tap_weight_sel<kTap>(d2, lob, clp);                   \
    const __half2 rg = u2h2(t[R][K].x), ba = u2h2(t[R][K].y);                                               \
    aR = __hfma2(__low2half2(rg), w, aR);                                                                   \
    aG = __hfma2(__high2half2(rg), w, aG);                                                                  \
    aB = __hfma2(__low2half2(ba), w, aB);                                                                   \
    aW = __hadd2(aW, w);                                                                                    \
  }
  // far taps first, near taps (f g j k, the large weights) last: less rounding error in the half accumulators
  if (kTap == 0) {
    FSR1_QTAP(0, 1) FSR1_QTAP(0, 2) FSR1_QTAP(1, 0) FSR1_QTAP(1, 3)
    FSR1_QTAP(2, 0) FSR1_QTAP(2, 3) FSR1_QTAP(3, 1) FSR1_QTAP(3, 2)
    FSR1_QTAP(1, 1) FSR1_QTAP(1, 2) FSR1_QTAP(2, 1) FSR1_QTAP(2, 2)
  } else {
    constexpr float oy1 = 0.0f - py, oy2 = 1.0f - py;
    const __half2 rowB1 = __hmul2(qb, h2c(oy1)), rowC1 = __hmul2(qc, h2c(oy1 * oy1));
    const __half2 rowB2 = __hmul2(qb, h2c(oy2)), rowC2 = __hmul2(qc, h2c(oy2 * oy2));
    FSR1_QTAP(0, 1) FSR1_QTAP(0, 2) FSR1_QTAP_ROW(1, 0, false) FSR1_QTAP_ROW(1, 3, false)
    FSR1_QTAP_ROW(2, 0, false) FSR1_QTAP_ROW(2, 3, false) FSR1_QTAP(3, 1) FSR1_QTAP(3, 2)
    FSR1_QTAP_ROW(1, 1, true) FSR1_QTAP_ROW(1, 2, true) FSR1_QTAP_ROW(2, 1, true) FSR1_QTAP_ROW(2, 2, true)
  }
#undef FSR1_QTAP_ROW
#undef FSR1_QTAP
  const float2 aWf = __half22float2(aW);
  const __half2 r = __floats2half2_rn(rcp_approx(aWf.x), rcp_approx(aWf.y));
  const __half2 oR = __hmin2(mxR, __hmax2(mnR, __hmul2(aR, r)));
  const __half2 oG = __hmin2(mxG, __hmax2(mnG, __hmul2(aG, r)));
  const __half2 oB = __hmin2(mxB, __hmax2(mnB, __hmul2(aB, r)));
  const __half2 one = h2c(1.0f);
  outA = make_uint2(h22u(__lows2half2(oR, oG)), h22u(__lows2half2(oB, one)));
  outB = make_uint2(h22u(__highs2half2(oR, oG)), h22u(__highs2half2(oB, one)));
}

// ---- output storage of the experimental 2x kernels: RGBA16F (8 B/px) or UNORM (4 B/px) --------------------------
struct StoreHalf {
  static constexpr int kBpp = 8;
  static __device__ __forceinline__ void put(unsigned char* o, uint2 v, bool ok) {
    if (ok) *reinterpret_cast<uint2*>(o) = v;
  }
};
// kBits = 8: R8G8B8A8_UNORM, 10: R10G10B10A2_UNORM.  x * (2^n - 1) + 1024 lands in [1024, 2048), where the ulp of a half
// is 1: one HFMA2 rounds to the nearest code value and leaves it in the low mantissa bits (alpha 1.0 -> all ones).
template <int kBits> struct StoreUnorm {
  static constexpr int kBpp = 4;
  static __device__ __forceinline__ void put(unsigned char* o, uint2 v, bool ok) {
    const __half2 sc = h2c(kBits == 8 ? 255.0f : 1023.0f), k1024 = h2c(1024.0f);
    const uint32_t t0 = h22u(__hfma2(u2h2(v.x), sc, k1024)), t1 = h22u(__hfma2(u2h2(v.y), sc, k1024));
    uint32_t w;
    if (kBits == 8) w = __byte_perm(t0, t1, 0x6420);  // R, G = bytes 0, 2 of t0; B, A = bytes 0, 2 of t1
    else w = (t0 & 0x3ffu) | (((t0 >> 16) & 0x3ffu) << 10) | ((t1 & 0x3ffu) << 20) | 0xC0000000u;
    if (ok) *reinterpret_cast<uint32_t*>(o) = w;
  }
};

// Phase 3 for one lane and one cell row r of a 2x tile: the quad of output pixels (2k+1,2k+2)x(2m+1,2m+2) of cell
// k = gx0+1+lane, m = gy0+1+r.  tile/S = the tile's texels and per-texel terms in shared memory.
// kFast (experimental, FSR1_EASU_QUAD_VARIANT=9): the tile lies strictly inside the image and the row range, so every
// bounds predicate is true and is dropped at compile time.
template <int kTap = 0, bool kFast = false, typename ST = StoreHalf>
__device__ __forceinline__ void quad_cell(const EasuParams& p, const uint2* __restrict__ tile, const float4* __restrict__ S,
                                          int gx0, int gy0, int lane, int r) {
  const int oxA = (gx0 + 1 + lane) * 2 + 1;  // cell k = gx0 + 1 + lane -> output columns 2k+1, 2k+2
    const int oyT = (gy0 + 1 + r) * 2 + 1;     // output rows 2m+1 (top pair), 2m+2 (bottom pair)
    const bool rowT = kFast || (oyT >= p.y0 && oyT < p.y1), rowB = kFast || (oyT + 1 >= p.y0 && oyT + 1 < p.y1);
    if (!kFast && (oxA >= p.out.w || !(rowT || rowB))) return;
    uint2 tp[4][4];
    const uint2* t0 = tile + r * kQBW + lane;
#pragma unroll
    for (int R = 0; R < 4; R++)
#pragma unroll
      for (int K = 0; K < 4; K++)
        if (!((R == 0 || R == 3) && (K == 0 || K == 3))) tp[R][K] = t0[R * kQBW + K];
    const float4* s0 = S + r * kQSW + lane;
    const float4 f = s0[0], g = s0[1], j = s0[kQSW], k = s0[kQSW + 1];
    // de-ringing bounds of the quad: min/max of f,g,j,k per channel, broadcast to both lanes
    const __half2 mnRG = __hmin2(__hmin2(u2h2(tp[1][1].x), u2h2(tp[1][2].x)), __hmin2(u2h2(tp[2][1].x), u2h2(tp[2][2].x)));
    const __half2 mxRG = __hmax2(__hmax2(u2h2(tp[1][1].x), u2h2(tp[1][2].x)), __hmax2(u2h2(tp[2][1].x), u2h2(tp[2][2].x)));
    const __half2 mnBA = __hmin2(__hmin2(u2h2(tp[1][1].y), u2h2(tp[1][2].y)), __hmin2(u2h2(tp[2][1].y), u2h2(tp[2][2].y)));
    const __half2 mxBA = __hmax2(__hmax2(u2h2(tp[1][1].y), u2h2(tp[1][2].y)), __hmax2(u2h2(tp[2][1].y), u2h2(tp[2][2].y)));
    const __half2 mnR = __low2half2(mnRG), mnG = __high2half2(mnRG), mnB = __low2half2(mnBA);
    const __half2 mxR = __low2half2(mxRG), mxG = __high2half2(mxRG), mxB = __low2half2(mxBA);
    if constexpr (kTap == 2 || kTap == 3) {
      constexpr int kPairTap = kTap == 2 ? 1 : 3;
      // packed pair (A: px=.25, B: px=.75): T = top texel row f,g blended horizontally, Bm = bottom texel row j,k
      const float2 wF = mk2(0.75f, 0.25f), wG = mk2(0.25f, 0.75f);
      const float2 Tx = __ffma2_rn(bc2(g.x), wG, __fmul2_rn(bc2(f.x), wF)), Ty = __ffma2_rn(bc2(g.y), wG, __fmul2_rn(bc2(f.y), wF));
      const float2 Tz = __ffma2_rn(bc2(g.z), wG, __fmul2_rn(bc2(f.z), wF));
      const float2 Bx = __ffma2_rn(bc2(k.x), wG, __fmul2_rn(bc2(j.x), wF)), By = __ffma2_rn(bc2(k.y), wG, __fmul2_rn(bc2(j.y), wF));
      const float2 Bz = __ffma2_rn(bc2(k.z), wG, __fmul2_rn(bc2(j.z), wF));
      unsigned char* orow2 = p.out.base + (long long)(oyT - p.out.row0) * p.out.pitch + (long long)oxA * ST::kBpp;
      const bool okA2 = kFast || oxA >= 0, okB2 = kFast || oxA + 1 < p.out.w;
      uint2 oA2, oB2;
      if (rowT) {
        const Shape2 s2 = pixel_shape2(__ffma2_rn(Bx, bc2(0.25f), __fmul2_rn(Tx, bc2(0.75f))),
                                       __ffma2_rn(By, bc2(0.25f), __fmul2_rn(Ty, bc2(0.75f))),
                                       __ffma2_rn(Bz, bc2(0.25f), __fmul2_rn(Tz, bc2(0.75f))));
        quad_pair<false, kPairTap>(tp, lane_x(s2), lane_y(s2), mnR, mnG, mnB, mxR, mxG, mxB, oA2, oB2);
        ST::put(orow2, oA2, okA2);
        ST::put(orow2 + ST::kBpp, oB2, okB2);
      }
      if (rowB) {
        const Shape2 s2 = pixel_shape2(__ffma2_rn(Bx, bc2(0.75f), __fmul2_rn(Tx, bc2(0.25f))),
                                       __ffma2_rn(By, bc2(0.75f), __fmul2_rn(Ty, bc2(0.25f))),
                                       __ffma2_rn(Bz, bc2(0.75f), __fmul2_rn(Tz, bc2(0.25f))));
        quad_pair<true, kPairTap>(tp, lane_x(s2), lane_y(s2), mnR, mnG, mnB, mxR, mxG, mxB, oA2, oB2);
        ST::put(orow2 + p.out.pitch, oA2, okA2);
        ST::put(orow2 + p.out.pitch + ST::kBpp, oB2, okB2);
      }
      return;
    }
    // bilinear blends with the four constant weight sets (pp = .25/.75): horizontal first
    const float fx = 0.75f, gx = 0.25f;
    const float3 t25 = make_float3(fmaf(g.x, gx, f.x * fx), fmaf(g.y, gx, f.y * fx), fmaf(g.z, gx, f.z * fx));
    const float3 t75 = make_float3(fmaf(g.x, fx, f.x * gx), fmaf(g.y, fx, f.y * gx), fmaf(g.z, fx, f.z * gx));
    const float3 b25 = make_float3(fmaf(k.x, gx, j.x * fx), fmaf(k.y, gx, j.y * fx), fmaf(k.z, gx, j.z * fx));
    const float3 b75 = make_float3(fmaf(k.x, fx, j.x * gx), fmaf(k.y, fx, j.y * gx), fmaf(k.z, fx, j.z * gx));
    unsigned char* orow = p.out.base + (long long)(oyT - p.out.row0) * p.out.pitch + (long long)oxA * 8;
    const bool okA = kFast || oxA >= 0, okB = kFast || oxA + 1 < p.out.w;
    uint2 oA, oB;
    if (rowT) {
      const Shape sA = pixel_shape(fmaf(b25.x, gx, t25.x * fx), fmaf(b25.y, gx, t25.y * fx), fmaf(b25.z, gx, t25.z * fx));
      const Shape sB = pixel_shape(fmaf(b75.x, gx, t75.x * fx), fmaf(b75.y, gx, t75.y * fx), fmaf(b75.z, gx, t75.z * fx));
      quad_pair<false, (kTap == 4 ? 3 : (kTap >= 2 ? 1 : kTap))>(tp, sA, sB, mnR, mnG, mnB, mxR, mxG, mxB, oA, oB);
      if (okA) *reinterpret_cast<uint2*>(orow) = oA;
      if (okB) *reinterpret_cast<uint2*>(orow + 8) = oB;
    }
    if (rowB) {
      const Shape sA = pixel_shape(fmaf(b25.x, fx, t25.x * gx), fmaf(b25.y, fx, t25.y * gx), fmaf(b25.z, fx, t25.z * gx));
      const Shape sB = pixel_shape(fmaf(b75.x, fx, t75.x * gx), fmaf(b75.y, fx, t75.y * gx), fmaf(b75.z, fx, t75.z * gx));
      quad_pair<true, (kTap == 4 ? 3 : (kTap >= 2 ? 1 : kTap))>(tp, sA, sB, mnR, mnG, mnB, mxR, mxG, mxB, oA, oB);
      if (okA) *reinterpret_cast<uint2*>(orow + p.out.pitch) = oA;
      if (okB) *reinterpret_cast<uint2*>(orow + p.out.pitch + 8) = oB;
    }
}

template <int NW> struct __align__(128) QuadSmem {
  uint2 tile[2][QuadCfg<NW>::kPad];
  float4 S[kQSW * QuadCfg<NW>::kSH];
  float L[QuadCfg<NW>::kElems];
  uint64_t bar[2];
};

template <int NW, int MINB, int kTap = 0, bool kFastPath = false>
__global__ void __launch_bounds__(NW * 32, MINB)
easu_h_quad2x_kernel(const EasuParams p, const __grid_constant__ CUtensorMap tmap, const int tiles_x,
                     const int n_tiles, const int mbase) {
  using C = QuadCfg<NW>;
  constexpr int NT = NW * 32;
  __shared__ QuadSmem<NW> sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&sm.bar[0], 1);
    mbar_init(&sm.bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  // tile t: cells k in [32 tx - 1, +32), m in [mbase + kCY ty, +kCY); box origin = (first cell) - 1
  auto box_x = [&](int t) { return (t % tiles_x) * kQCX - 2; };
  auto box_y = [&](int t) { return mbase + (t / tiles_x) * C::kCY - 1; };
  int t = blockIdx.x;
  // experimental build (kFastPath): tile coordinates advance incrementally — one division per kernel instead of one
  // per tile and thread (the persistent loop's bookkeeping is ~10 % of the instructions executed)
  int tx = 0, ty = 0, step_x = 0, step_y = 0;
  if constexpr (kFastPath) {
    tx = t % tiles_x;
    ty = t / tiles_x;
    step_x = (int)gridDim.x % tiles_x;
    step_y = (int)gridDim.x / tiles_x;
  }
  if (tid == 0 && t < n_tiles) {
    mbar_expect_tx(&sm.bar[0], C::kElems * 8u);
    tma_load_2d(sm.tile[0], &tmap, box_x(t), box_y(t) - p.in.row0, &sm.bar[0]);
  }
  for (int it = 0; t < n_tiles; t += gridDim.x, it++) {
    const int b = it & 1;
    const int tn = t + gridDim.x;
    int txn = tx + step_x, tyn = ty + step_y;  // coordinates of tile tn (kFastPath only)
    if (txn >= tiles_x) { txn -= tiles_x; tyn++; }
    if (tid == 0 && tn < n_tiles) {  // prefetch the next tile into the other buffer (its readers all passed
      fence_proxy_async();           // the barrier that closed the previous iteration)
      mbar_expect_tx(&sm.bar[b ^ 1], C::kElems * 8u);
      if constexpr (kFastPath)
        tma_load_2d(sm.tile[b ^ 1], &tmap, txn * kQCX - 2, mbase + tyn * C::kCY - 1 - p.in.row0, &sm.bar[b ^ 1]);
      else
        tma_load_2d(sm.tile[b ^ 1], &tmap, box_x(tn), box_y(tn) - p.in.row0, &sm.bar[b ^ 1]);
    }
    uint2* tile = sm.tile[b];
    const int gx0 = kFastPath ? tx * kQCX - 2 : box_x(t), gy0 = kFastPath ? mbase + ty * C::kCY - 1 : box_y(t);
    tx = txn;
    ty = tyn;
    mbar_wait(&sm.bar[b], (it >> 1) & 1);
    if (gx0 < 0 || gy0 < 0 || gx0 + kQBW > p.in.w || gy0 + C::kBH > p.in.h) {
      for (int j = warp; j < C::kBH; j += NW) {
        const int cy = clampi(gy0 + j, 0, p.in.h - 1) - gy0;
        for (int i = lane; i < kQBW; i += 32) {
          const int cx = clampi(gx0 + i, 0, p.in.w - 1) - gx0;
          if ((cx != i || cy != j) && cx >= 0 && cx < kQBW && cy >= 0 && cy < C::kBH) tile[j * kQBW + i] = tile[cy * kQBW + cx];
        }
      }
      fence_proxy_async();  // these generic-proxy writes are later overwritten by a TMA (async proxy) load
      __syncthreads();
    }
    for (int i = tid; i < C::kElems; i += NT) sm.L[i] = texel_luma(tile[i]);
    __syncthreads();
    for (int idx = tid; idx < kQSW * C::kSH; idx += NT) {
      const int j = idx / kQSW, i = idx - j * kQSW;
      const float* c = sm.L + (j + 1) * kQBW + (i + 1);
      sm.S[idx] = texel_terms(c[-kQBW], c[-1], c[0], c[1], c[kQBW]);
    }
    __syncthreads();

    if constexpr (kFastPath) {
      // every output pixel of the tile (columns 2(gx0+1)+1 .. 2(gx0+32)+2, rows 2(gy0+1)+1 .. 2(gy0+2NW)+2) is in range
      const bool inside = 2 * (gx0 + 1) + 1 >= 0 && 2 * (gx0 + 32) + 2 < p.out.w && 2 * (gy0 + 1) + 1 >= p.y0 &&
                          2 * (gy0 + C::kCY) + 2 < p.y1;
      if (inside) {
#pragma unroll 1
        for (int q = 0; q < 2; q++) quad_cell<kTap, true>(p, tile, sm.S, gx0, gy0, lane, warp + q * NW);
      } else {
#pragma unroll 1
        for (int q = 0; q < 2; q++) quad_cell<kTap, false>(p, tile, sm.S, gx0, gy0, lane, warp + q * NW);
      }
    } else {
#pragma unroll 1
      for (int q = 0; q < 2; q++) quad_cell<kTap>(p, tile, sm.S, gx0, gy0, lane, warp + q * NW);
    }
    __syncthreads();  // L, S and this tile buffer are free again
  }
}

// ---- 2x EASU for the UNORM formats the sample renders into (experimental, FSR1_UNORM_TILED=1) --------------------------
// Same phases as easu_h_quad2x_kernel.  The TMA box holds 4-byte texels (origin rounded down to 4 texels = 16 bytes, so
// the box is 40 wide); one pass decodes it (c / (2^n - 1), fp32) into the half tile the tap loop reads and into the
// fp32 luma plane; the epilogue re-encodes in the half domain (StoreUnorm).  Against quantise(oracle(dequantise(in))) the
// stored codes differ by at most one (half tile + half taps), like the default arithmetic of the direct kernels.
constexpr int kUBW = kQBW + 4;
template <int NW> struct __align__(128) QuadSmemU {
  uint32_t stage[2][((kUBW * QuadCfg<NW>::kBH * 4 + 127) / 128) * 128 / 4];
  uint2 tile[QuadCfg<NW>::kPad];
  float4 S[kQSW * QuadCfg<NW>::kSH];
  float L[QuadCfg<NW>::kElems];
  uint64_t bar[2];
};

template <int kBits> __device__ __forceinline__ float3 unorm_decode(uint32_t u) {
  if (kBits == 8) {
    const float k = 1.0f / 255.0f;
    return make_float3((float)(u & 255u) * k, (float)((u >> 8) & 255u) * k, (float)((u >> 16) & 255u) * k);
  }
  const float k = 1.0f / 1023.0f;
  return make_float3((float)(u & 1023u) * k, (float)((u >> 10) & 1023u) * k, (float)((u >> 20) & 1023u) * k);
}

template <int NW, int MINB, int kBits>
__global__ void __launch_bounds__(NW * 32, MINB)
easu_u_quad2x_kernel(const EasuParams p, const __grid_constant__ CUtensorMap tmap, const int tiles_x, const int n_tiles,
                     const int mbase) {
  using C = QuadCfg<NW>;
  constexpr int NT = NW * 32;
  constexpr uint32_t kBoxBytes = kUBW * C::kBH * 4u;
  __shared__ QuadSmemU<NW> sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&sm.bar[0], 1);
    mbar_init(&sm.bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  int t = blockIdx.x;
  int tx = t % tiles_x, ty = t / tiles_x;
  const int step_x = (int)gridDim.x % tiles_x, step_y = (int)gridDim.x / tiles_x;
  // tile (tx, ty): half tile origin gx0 = 32 tx - 2 (as in the half kernel); the staging box starts 2 texels further left
  if (tid == 0 && t < n_tiles) {
    mbar_expect_tx(&sm.bar[0], kBoxBytes);
    tma_load_2d(sm.stage[0], &tmap, tx * kQCX - 4, mbase + ty * C::kCY - 1 - p.in.row0, &sm.bar[0]);
  }
  for (int it = 0; t < n_tiles; t += gridDim.x, it++) {
    const int b = it & 1;
    int txn = tx + step_x, tyn = ty + step_y;
    if (txn >= tiles_x) { txn -= tiles_x; tyn++; }
    if (tid == 0 && t + (int)gridDim.x < n_tiles) {
      fence_proxy_async();
      mbar_expect_tx(&sm.bar[b ^ 1], kBoxBytes);
      tma_load_2d(sm.stage[b ^ 1], &tmap, txn * kQCX - 4, mbase + tyn * C::kCY - 1 - p.in.row0, &sm.bar[b ^ 1]);
    }
    const int gx0 = tx * kQCX - 2, gy0 = mbase + ty * C::kCY - 1;
    tx = txn;
    ty = tyn;
    uint32_t* stage = sm.stage[b];
    mbar_wait(&sm.bar[b], (it >> 1) & 1);
    if (gx0 < 0 || gy0 < 0 || gx0 + kQBW > p.in.w || gy0 + C::kBH > p.in.h) {  // zero fill -> clamp-to-edge, on the raw texels
      for (int j = warp; j < C::kBH; j += NW) {
        const int cy = clampi(gy0 + j, 0, p.in.h - 1) - gy0;
        for (int i = lane; i < kQBW; i += 32) {
          const int cx = clampi(gx0 + i, 0, p.in.w - 1) - gx0;
          if ((cx != i || cy != j) && cx >= 0 && cx < kQBW && cy >= 0 && cy < C::kBH) stage[j * kUBW + 2 + i] = stage[cy * kUBW + 2 + cx];
        }
      }
      fence_proxy_async();
      __syncthreads();
    }
    for (int i = tid; i < C::kElems; i += NT) {  // decode: half tile for the taps, fp32 luma for the analysis
      const int j = i / kQBW, c = i - j * kQBW;
      const float3 v = unorm_decode<kBits>(stage[j * kUBW + 2 + c]);
      sm.tile[i] = make_uint2(h22u(__floats2half2_rn(v.x, v.y)), h22u(__floats2half2_rn(v.z, 1.0f)));
      sm.L[i] = fmaf(v.z, 0.5f, fmaf(v.x, 0.5f, v.y));
    }
    __syncthreads();
    for (int idx = tid; idx < kQSW * C::kSH; idx += NT) {
      const int j = idx / kQSW, i = idx - j * kQSW;
      const float* c = sm.L + (j + 1) * kQBW + (i + 1);
      sm.S[idx] = texel_terms(c[-kQBW], c[-1], c[0], c[1], c[kQBW]);
    }
    __syncthreads();
    const bool inside = 2 * (gx0 + 1) + 1 >= 0 && 2 * (gx0 + 32) + 2 < p.out.w && 2 * (gy0 + 1) + 1 >= p.y0 &&
                        2 * (gy0 + C::kCY) + 2 < p.y1;
    if (inside) {
#pragma unroll 1
      for (int q = 0; q < 2; q++) quad_cell<3, true, StoreUnorm<kBits>>(p, sm.tile, sm.S, gx0, gy0, lane, warp + q * NW);
    } else {
#pragma unroll 1
      for (int q = 0; q < 2; q++) quad_cell<3, false, StoreUnorm<kBits>>(p, sm.tile, sm.S, gx0, gy0, lane, warp + q * NW);
    }
    __syncthreads();  // stage, tile, L and S are free again
  }
}

// ---- warp-specialised variant: one producer warp prepares tile i+1 (TMA wait, clamp fix-up, luma, terms) while
// NWC consumer warps run phase 3 of tile i.  No CTA-wide barrier in the steady state: the hand-offs are mbarriers
// (ready[b]: producer -> consumers, free_[b]: consumers -> producer), tile/L/S are all double-buffered.
#ifndef FSR1_CPU_EMU
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#endif

template <int NWC> struct __align__(128) QuadWsSmem {
  uint2 tile[2][QuadCfg<NWC>::kPad];
  float4 S[2][kQSW * QuadCfg<NWC>::kSH];
  float L[2][QuadCfg<NWC>::kElems];
  uint64_t tma[2], ready[2], free_[2];
};

template <int NWC, int MINB>
__global__ void __launch_bounds__((NWC + 1) * 32, MINB)
easu_h_quad2x_ws_kernel(const EasuParams p, const __grid_constant__ CUtensorMap tmap, const int tiles_x,
                        const int n_tiles, const int mbase) {
  using C = QuadCfg<NWC>;
  __shared__ QuadWsSmem<NWC> sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int b = 0; b < 2; b++) {
      mbar_init(&sm.tma[b], 1);
      mbar_init(&sm.ready[b], 1);
      mbar_init(&sm.free_[b], NWC);
    }
    mbar_fence_init();
  }
  __syncthreads();
  auto box_x = [&](int t) { return (t % tiles_x) * kQCX - 2; };
  auto box_y = [&](int t) { return mbase + (t / tiles_x) * C::kCY - 1; };
  const int t0 = blockIdx.x, stride = gridDim.x;
  const int my_tiles = t0 < n_tiles ? (n_tiles - t0 + stride - 1) / stride : 0;

  if (warp == NWC) {
    // ------------------------------------------------------------------ producer warp
    if (lane == 0 && my_tiles > 0) {
      mbar_expect_tx(&sm.tma[0], C::kElems * 8u);
      tma_load_2d(sm.tile[0], &tmap, box_x(t0), box_y(t0) - p.in.row0, &sm.tma[0]);
    }
    for (int it = 0; it < my_tiles; it++) {
      const int b = it & 1, t = t0 + it * stride;
      uint2* tile = sm.tile[b];
      float* L = sm.L[b];
      float4* S = sm.S[b];
      const int gx0 = box_x(t), gy0 = box_y(t);
      mbar_wait(&sm.tma[b], (it >> 1) & 1);
      if (gx0 < 0 || gy0 < 0 || gx0 + kQBW > p.in.w || gy0 + C::kBH > p.in.h) {
        for (int idx = lane; idx < C::kElems; idx += 32) {
          const int j = idx / kQBW, i = idx - j * kQBW;
          const int cy = clampi(gy0 + j, 0, p.in.h - 1) - gy0, cx = clampi(gx0 + i, 0, p.in.w - 1) - gx0;
          if ((cx != i || cy != j) && cx >= 0 && cx < kQBW && cy >= 0 && cy < C::kBH) tile[idx] = tile[cy * kQBW + cx];
        }
        fence_proxy_async();
        __syncwarp();
      }
      for (int i = lane; i < C::kElems; i += 32) L[i] = texel_luma(tile[i]);
      __syncwarp();
      for (int idx = lane; idx < kQSW * C::kSH; idx += 32) {
        const int j = idx / kQSW, i = idx - j * kQSW;
        const float* c = L + (j + 1) * kQBW + (i + 1);
        S[idx] = texel_terms(c[-kQBW], c[-1], c[0], c[1], c[kQBW]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.ready[b]);  // release: the warp's shared-memory writes are visible to waiters
      // refill the OTHER buffer with tile it+1 once the consumers have finished tile it-1 (which used it)
      if (it + 1 < my_tiles) {
        if (it >= 1) mbar_wait(&sm.free_[b ^ 1], ((it - 1) >> 1) & 1);
        if (lane == 0) {
          fence_proxy_async();
          mbar_expect_tx(&sm.tma[b ^ 1], C::kElems * 8u);
          tma_load_2d(sm.tile[b ^ 1], &tmap, box_x(t + stride), box_y(t + stride) - p.in.row0, &sm.tma[b ^ 1]);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ consumer warps
    for (int it = 0; it < my_tiles; it++) {
      const int b = it & 1, t = t0 + it * stride;
      mbar_wait(&sm.ready[b], (it >> 1) & 1);
      const int gx0 = box_x(t), gy0 = box_y(t);
#pragma unroll 1
      for (int q = 0; q < 2; q++) quad_cell(p, sm.tile[b], sm.S[b], gx0, gy0, lane, warp + q * NWC);
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.free_[b]);
    }
  }
}

#ifndef FSR1_CPU_EMU
// ---- host side ----------------------------------------------------------------------------------------
// one RGBA16F texel = one 64-bit TMA element; tensor = the stored window of the image
static bool make_tmap(CUtensorMap* tmap, const ImgView& in, int BW, int BH) {
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)in.w, (cuuint64_t)in.rows};
  const cuuint64_t strides[1] = {(cuuint64_t)in.pitch};
  const cuuint32_t box[2] = {(cuuint32_t)BW, (cuuint32_t)BH};
  const cuuint32_t estr[2] = {1, 1};
  return encode(tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, in.base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Largest footprint (in texels) any tile of `tile` output pixels needs along one axis; with even_origin the
// box starts at the even texel at or before the first tap (see the kernel).
static int max_footprint(int n_out, int first, int tile, float scale, float offset, bool even_origin) {
  int best = 4;
  for (int o0 = first; o0 < n_out; o0 += tile) {
    const int o1 = (o0 + tile - 1 < n_out - 1) ? o0 + tile - 1 : n_out - 1;
    int origin = host_fp(o0, scale, offset) - 1;
    if (even_origin) origin &= ~1;
    const int span = host_fp(o1, scale, offset) + 2 - origin + 1;
    if (span > best) best = span;
  }
  return best;
}

// UNORM images, exactly 2x (experimental: FSR1_UNORM_TILED=1; otherwise the caller uses the direct kernels)
cudaError_t launch_easu_u_tiled(const EasuParams& p, int format, cudaStream_t s, const char** name) {
  static const int enabled = env_knob("FSR1_UNORM_TILED", 0);  // 1: R8G8B8A8 only, 2: also R10G10B10A2 (coarser than its codes)
  if (!(enabled >= 1 && format == 3) && !(enabled >= 2 && format == 4)) return cudaErrorNotSupported;
  if ((reinterpret_cast<uintptr_t>(p.in.base) & 15) || (p.in.pitch & 15)) return cudaErrorNotSupported;  // TMA
  if (!(p.c0x == 0.5f && p.c0y == 0.5f && p.c0z == -0.25f && p.c0w == -0.25f)) return cudaErrorNotSupported;
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return cudaErrorNotSupported;
  constexpr int NW = 4, CY = 2 * NW;
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {(cuuint64_t)p.in.w, (cuuint64_t)p.in.rows};
  const cuuint64_t strides[1] = {(cuuint64_t)p.in.pitch};
  const cuuint32_t box[2] = {(cuuint32_t)kUBW, (cuuint32_t)(CY + 3)};
  const cuuint32_t estr[2] = {1, 1};
  if (encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, p.in.base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return cudaErrorNotSupported;
  const int k_first = -1, k_last = host_fp(p.out.w - 1, 0.5f, -0.25f);
  const int m_first = host_fp(p.y0, 0.5f, -0.25f), m_last = host_fp(p.y1 - 1, 0.5f, -0.25f);
  const int tiles_x = (k_last - k_first + 1 + kQCX - 1) / kQCX;
  const int tiles_y = (m_last - m_first + 1 + CY - 1) / CY, n_tiles = tiles_x * tiles_y;
  const int grid = n_tiles < 6 * sm_count() ? n_tiles : 6 * sm_count();
  if (format == 3) {
    easu_u_quad2x_kernel<NW, 6, 8><<<grid, NW * 32, 0, s>>>(p, tmap, tiles_x, n_tiles, m_first);
    *name = "easu_u8_quad2x<4w,6/sm,tma2>";
  } else {
    easu_u_quad2x_kernel<NW, 6, 10><<<grid, NW * 32, 0, s>>>(p, tmap, tiles_x, n_tiles, m_first);
    *name = "easu_u10_quad2x<4w,6/sm,tma2>";
  }
  return cudaGetLastError();
}

cudaError_t launch_easu_h_tiled(const EasuParams& p, cudaStream_t s, const char** name) {
  // layout requirements of TMA and of the vector stores
  if ((reinterpret_cast<uintptr_t>(p.in.base) & 15) || (p.in.pitch & 15) || (reinterpret_cast<uintptr_t>(p.out.base) & 15) ||
      (p.out.pitch & 15))
    return cudaErrorNotSupported;
  CUtensorMap tmap;

  if (p.c0x == 0.5f && p.c0y == 0.5f && p.c0z == -0.25f && p.c0w == -0.25f) {  // exactly 2x
    // development knob: FSR1_EASU_QUAD_VARIANT = 0 (8 warps x2), 1 (8 warps x3), 2 (4 warps x6, plain tap form),
    // 3/4 (warp-specialised), 5 (4 warps x7), 6 = default (4 warps x6, factored tap distance),
    // 7 (experimental, unmeasured: as 6 with the per-pixel fp32 analysis packed in f32x2), 8 (7 + integer distance clamp), 9 (8 + predicate-free path for interior tiles + incremental tile coordinates),
    // 10 (the default's scalar fp32 analysis with 9's integer clamp, interior path and incremental coordinates),
    // 12 (9 at 7 CTAs per SM: variant 9 needs 72 registers, which is exactly what 7 x 128 threads allow)
    static const int variant = env_knob("FSR1_EASU_QUAD_VARIANT", 6);
    const int k_first = -1, k_last = host_fp(p.out.w - 1, 0.5f, -0.25f);
    const int m_first = host_fp(p.y0, 0.5f, -0.25f), m_last = host_fp(p.y1 - 1, 0.5f, -0.25f);
    const int tiles_x = (k_last - k_first + 1 + kQCX - 1) / kQCX;
    // FSR1_EASU_CTAS_PER_SM: leave room on every SM for a concurrently running kernel (measured: slower, with and without pipelining)
    static const int cap = env_knob("FSR1_EASU_CTAS_PER_SM", 0);
    auto launch = [&](auto kernel, int nw, int per_sm, const char* nm) -> cudaError_t {
      const int cy = 2 * nw;
      if (cap > 0 && cap < per_sm) per_sm = cap;
      if (!make_tmap(&tmap, p.in, kQBW, cy + 3)) return cudaErrorNotSupported;
      const int tiles_y = (m_last - m_first + 1 + cy - 1) / cy, n_tiles = tiles_x * tiles_y;
      const int grid = n_tiles < per_sm * sm_count() ? n_tiles : per_sm * sm_count();
      kernel<<<grid, nw * 32, 0, s>>>(p, tmap, tiles_x, n_tiles, m_first);
      *name = nm;
      return cudaGetLastError();
    };
    if (variant == 3 || variant == 4) {  // warp-specialised: 4 consumer warps + 1 producer warp
      constexpr int NWC = 4;
      if (!make_tmap(&tmap, p.in, kQBW, QuadCfg<NWC>::kBH)) return cudaErrorNotSupported;
      const int cy = QuadCfg<NWC>::kCY, tiles_y = (m_last - m_first + 1 + cy - 1) / cy, n_tiles = tiles_x * tiles_y;
      const int per_sm = variant == 3 ? 5 : 4;
      const int grid = n_tiles < per_sm * sm_count() ? n_tiles : per_sm * sm_count();
      if (variant == 3) easu_h_quad2x_ws_kernel<NWC, 5><<<grid, (NWC + 1) * 32, 0, s>>>(p, tmap, tiles_x, n_tiles, m_first);
      else easu_h_quad2x_ws_kernel<NWC, 4><<<grid, (NWC + 1) * 32, 0, s>>>(p, tmap, tiles_x, n_tiles, m_first);
      *name = variant == 3 ? "easu_h_quad2x_ws<4+1w,5/sm,tma2>" : "easu_h_quad2x_ws<4+1w,4/sm,tma2>";
      return cudaGetLastError();
    }
    if (variant == 5) return launch(easu_h_quad2x_kernel<4, 7>, 4, 7, "easu_h_quad2x<4w,7/sm,tma2>");
    if (variant == 2) return launch(easu_h_quad2x_kernel<4, 6, 0>, 4, 6, "easu_h_quad2x<4w,6/sm,tma2,plain>");
    if (variant == 7) return launch(easu_h_quad2x_kernel<4, 6, 2>, 4, 6, "easu_h_quad2x<4w,6/sm,tma2,f32x2shape>");  // experimental
    if (variant == 8) return launch(easu_h_quad2x_kernel<4, 6, 3>, 4, 6, "easu_h_quad2x<4w,6/sm,tma2,f32x2shape,iclamp>");  // experimental
    if (variant == 9) return launch(easu_h_quad2x_kernel<4, 6, 3, true>, 4, 6, "easu_h_quad2x<4w,6/sm,tma2,f32x2shape,iclamp,interior>");  // experimental
    if (variant == 10) return launch(easu_h_quad2x_kernel<4, 6, 4, true>, 4, 6, "easu_h_quad2x<4w,6/sm,tma2,iclamp,interior>");  // experimental
    if (variant == 12) return launch(easu_h_quad2x_kernel<4, 7, 3, true>, 4, 7, "easu_h_quad2x<4w,7/sm,tma2,f32x2shape,iclamp,interior>");  // experimental
    if (variant == 0) return launch(easu_h_quad2x_kernel<8, 2>, 8, 2, "easu_h_quad2x<8w,2/sm,tma2>");
    if (variant == 1) return launch(easu_h_quad2x_kernel<8, 3>, 8, 3, "easu_h_quad2x<8w,3/sm,tma2>");
    return launch(easu_h_quad2x_kernel<4, 6, 1>, 4, 6, "easu_h_quad2x<4w,6/sm,tma2>");
  }

  if (!(p.c0x > 0.0f && p.c0x <= 1.0f && p.c0y > 0.0f && p.c0y <= 1.0f)) return cudaErrorNotSupported;  // upscaling only
  int BW = max_footprint(p.out.w, 0, kTileW, p.c0x, p.c0z, true);
  int BH = max_footprint(p.y1, p.y0, kTileH, p.c0y, p.c0w, false);
  BW = (BW + 1) & ~1;  // inner box extent must be a multiple of 16 bytes
  if (BW > 256 || BH > 256) return cudaErrorNotSupported;
  const size_t smem = pairs_smem_bytes(BW, BH);
  if (smem > 200 * 1024) return cudaErrorNotSupported;
  if (!make_tmap(&tmap, p.in, BW, BH)) return cudaErrorNotSupported;
  static const int pairs_variant = env_knob("FSR1_EASU_PAIRS_VARIANT", 0);  // 1: experimental (see vpair)
  if (smem > 48 * 1024) {  // per device and cheap: set on every launch that needs the opt-in
    cudaError_t e = pairs_variant == 1
                        ? cudaFuncSetAttribute(easu_h_pairs_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                        : cudaFuncSetAttribute(easu_h_pairs_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const int tiles_x = (p.out.w + kTileW - 1) / kTileW, n_tiles = tiles_x * ((p.y1 - p.y0 + kTileH - 1) / kTileH);
  int per_sm = 3;
  while (per_sm > 1 && (size_t)per_sm * (smem + 1024) > 220 * 1024) per_sm--;
  const int grid = n_tiles < per_sm * sm_count() ? n_tiles : per_sm * sm_count();
  if (pairs_variant == 1) {
    easu_h_pairs_kernel<1><<<grid, kThreads, smem, s>>>(p, tmap, BW, BH, tiles_x, n_tiles);
    *name = "easu_h_vpairs<64x32,persistent,tma2,f32x2shape,iclamp>";
  } else {
    easu_h_pairs_kernel<0><<<grid, kThreads, smem, s>>>(p, tmap, BW, BH, tiles_x, n_tiles);
    *name = "easu_h_vpairs<64x32,persistent,tma2>";
  }
  return cudaGetLastError();
}

#endif  // FSR1_CPU_EMU

}  // namespace fsr1
