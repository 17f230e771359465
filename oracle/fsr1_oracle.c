/* TEST INFRASTRUCTURE — the CPU oracle.  Not product code: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may build, load or call this file.
 *
 * A plain-C restatement of the FSR 1.0 hot path, written from the algorithm (not copied) and
 * PINNED against the reference itself: oracle/_ref/libfsr1_ref.so is the reference's own
 * ffx_fsr1.h source compiled for the host (oracle/build_ref.sh), and tests/test_oracle.py checks
 * this file bit-for-bit against it and against the known-answer table of SURVEY.md §8(c)
 * (committed as tests/golden/).  Build with -ffp-contract=off: every a*b+c below is then two
 * roundings, exactly like the reference compiled the same way.
 *
 * What is restated, with the reference lines each function follows:
 *   fsr1o_easu_con / _offset   FsrEasuCon / FsrEasuConOffset   ffx-fsr/ffx_fsr1.h:156-202, 205-225
 *   fsr1o_rcas_con             FsrRcasCon                      ffx-fsr/ffx_fsr1.h:662-672
 *   f32_to_f16_trunc           AU1_AH1_AF1 (table packer)      ffx-fsr/ffx_a.h:482-549
 *   prx_lo_rcp/prx_lo_rsq/prx_med_rcp   APrx*F1                ffx-fsr/ffx_a.h:1843-1845
 *   easu_set / easu_tap / fsr1o_easu_f32   FsrEasuSetF/TapF/F  ffx-fsr/ffx_fsr1.h:275-313, 239-272, 315-437
 *   fsr1o_rcas_f32             FsrRcasF                        ffx-fsr/ffx_fsr1.h:684-769
 *   fsr1o_easu_h16 / fsr1o_rcas_h16   FsrEasuH / FsrRcasH      ffx-fsr/ffx_fsr1.h:452-593, 782-866
 *     (half approximations ffx-fsr/ffx_a.h:1808,1814,1820)
 *   fsr1o_lfga_f32             FsrLfgaF                        ffx-fsr/ffx_fsr1.h:1014
 *   fsr1o_srtm_f32             FsrSrtmF / FsrSrtmInvF          ffx-fsr/ffx_fsr1.h:1044-1046
 *   fsr1o_tepd_dit / fsr1o_tepd_f32   FsrTepdDitF / FsrTepdC8F / FsrTepdC10F   ffx-fsr/ffx_fsr1.h:1086-1126
 *     (AGtZeroF1 = saturate(m * +INF) ffx-fsr/ffx_a.h:1499; APrxMedRcpF3 :1854)
 * Addressing follows the sample's shader wrapper: gather4 through a linear/clamp sampler
 * (sample/src/DX12/FSR_Pass.hlsl:39-41, FSR_Filter.cpp:48-53) = per-texel clamp-to-edge of the
 * integer tap coordinate; RCAS uses integer Load, out of bounds -> 0 (D3D12) or clamp (switch).
 * Outputs are stored like the sample does: (r,g,b,1)  (FSR_Pass.hlsl:80,95).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline float satf(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); } /* saturate(NaN)=0 */

/* ------------------------------------------------------------------ constants ---------------- */
void fsr1o_easu_con(uint32_t* con, float inVpW, float inVpH, float inW, float inH, float outW, float outH) {
  float rW = 1.0f / inW, rH = 1.0f / inH;
  con[0] = f2u(inVpW * (1.0f / outW));
  con[1] = f2u(inVpH * (1.0f / outH));
  con[2] = f2u(0.5f * inVpW * (1.0f / outW) - 0.5f);
  con[3] = f2u(0.5f * inVpH * (1.0f / outH) - 0.5f);
  con[4] = f2u(rW);            con[5] = f2u(rH);
  con[6] = f2u(1.0f * rW);     con[7] = f2u(-1.0f * rH);
  con[8] = f2u(-1.0f * rW);    con[9] = f2u(2.0f * rH);
  con[10] = f2u(1.0f * rW);    con[11] = f2u(2.0f * rH);
  con[12] = f2u(0.0f * rW);    con[13] = f2u(4.0f * rH);
  con[14] = 0; con[15] = 0;
}
void fsr1o_easu_con_offset(uint32_t* con, float inVpW, float inVpH, float inW, float inH, float outW,
                           float outH, float offX, float offY) {
  fsr1o_easu_con(con, inVpW, inVpH, inW, inH, outW, outH);
  con[2] = f2u(0.5f * inVpW * (1.0f / outW) - 0.5f + offX);
  con[3] = f2u(0.5f * inVpH * (1.0f / outH) - 0.5f + offY);
}
/* float -> half, TRUNCATING mantissa, subnormals kept, inf/NaN -> +-65504, like the table packer. */
uint32_t fsr1o_f32_to_f16_trunc(float f) {
  uint32_t u = f2u(f), s = (u >> 16) & 0x8000u, e = (u >> 23) & 0xffu, m = u & 0x7fffffu;
  if (e < 103) return s;                                   /* < 2^-24: flush */
  if (e < 113) return s + (0x0400u >> (113 - e)) + (m >> (126 - e)); /* half subnormal */
  if (e < 143) return s + ((e - 112) << 10) + (m >> 13);    /* normal */
  return s + 0x7bffu;                                        /* overflow, inf, NaN */
}
void fsr1o_rcas_con(uint32_t* con, float sharpness) {
  float s = exp2f(-sharpness);
  uint32_t h = fsr1o_f32_to_f16_trunc(s);
  con[0] = f2u(s); con[1] = h + (h << 16); con[2] = 0; con[3] = 0;
}

/* ------------------------------------------------------------------ fp32 path ---------------- */
static inline float prx_lo_rcp(float a) { return u2f(0x7ef07ebbu - f2u(a)); }
static inline float prx_lo_rsq(float a) { return u2f(0x5f347d74u - (f2u(a) >> 1)); }
static inline float prx_med_rcp(float a) { float b = u2f(0x7ef19fffu - f2u(a)); return b * (-b * a + 2.0f); }

/* One texel's contribution to the edge direction/length estimate; lA..lE = up,left,centre,right,down. */
static inline void easu_set(float* dx, float* dy, float* len, float w, float lA, float lB, float lC, float lD, float lE) {
  float dc = lD - lC, cb = lC - lB;
  float lenX = prx_lo_rcp(fmaxf(fabsf(dc), fabsf(cb)));
  float dirX = lD - lB;
  *dx += dirX * w;
  lenX = satf(fabsf(dirX) * lenX); lenX *= lenX;
  *len += lenX * w;
  float ec = lE - lC, ca = lC - lA;
  float lenY = prx_lo_rcp(fmaxf(fabsf(ec), fabsf(ca)));
  float dirY = lE - lA;
  *dy += dirY * w;
  lenY = satf(fabsf(dirY) * lenY); lenY *= lenY;
  *len += lenY * w;
}
static inline void easu_tap(float aC[3], float* aW, float ox, float oy, float dx, float dy, float l2x, float l2y,
                            float lob, float clp, const float c[3]) {
  float vx = (ox * dx) + (oy * dy);
  float vy = (ox * (-dy)) + (oy * dx);
  vx *= l2x; vy *= l2y;
  float d2 = vx * vx + vy * vy;
  d2 = fminf(d2, clp);
  float wB = (float)(2.0 / 5.0) * d2 + -1.0f;
  float wA = lob * d2 + -1.0f;
  wB *= wB; wA *= wA;
  wB = (float)(25.0 / 16.0) * wB + (float)(-(25.0 / 16.0 - 1.0));
  float w = wB * wA;
  aC[0] += c[0] * w; aC[1] += c[1] * w; aC[2] += c[2] * w; *aW += w;
}

/* One EASU output pixel.  t[r][c] = texel (fx-1+c, fy-1+r), RGB, already clamp-fetched. */
static inline void easu_pixel(float pix[3], float ppx, float ppy, float t[4][4][3]) {
  float L[4][4];
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++)
    L[r][c] = t[r][c][2] * 0.5f + (t[r][c][0] * 0.5f + t[r][c][1]);
  /*    b c        L[0][1] L[0][2]
   *  e f g h   =  L[1][0..3]
   *  i j k l      L[2][0..3]
   *    n o        L[3][1] L[3][2]                                                   */
  float dx = 0.0f, dy = 0.0f, len = 0.0f;
  easu_set(&dx, &dy, &len, (1.0f - ppx) * (1.0f - ppy), L[0][1], L[1][0], L[1][1], L[1][2], L[2][1]); /* f */
  easu_set(&dx, &dy, &len, ppx * (1.0f - ppy),          L[0][2], L[1][1], L[1][2], L[1][3], L[2][2]); /* g */
  easu_set(&dx, &dy, &len, (1.0f - ppx) * ppy,          L[1][1], L[2][0], L[2][1], L[2][2], L[3][1]); /* j */
  easu_set(&dx, &dy, &len, ppx * ppy,                   L[1][2], L[2][1], L[2][2], L[2][3], L[3][2]); /* k */
  float dirR = dx * dx + dy * dy;
  int zro = dirR < (float)(1.0 / 32768.0);
  dirR = prx_lo_rsq(dirR);
  dirR = zro ? 1.0f : dirR;
  dx = zro ? 1.0f : dx;
  dx *= dirR; dy *= dirR;
  len = len * 0.5f; len *= len;
  float stretch = (dx * dx + dy * dy) * prx_lo_rcp(fmaxf(fabsf(dx), fabsf(dy)));
  float l2x = 1.0f + (stretch - 1.0f) * len;
  float l2y = 1.0f + -0.5f * len;
  float lob = 0.5f + (float)((1.0 / 4.0 - 0.04) - 0.5) * len;
  float clp = prx_lo_rcp(lob);
  float mn[3], mx[3];
  for (int k = 0; k < 3; k++) {
    float f = t[1][1][k], g = t[1][2][k], j = t[2][1][k], kk = t[2][2][k];
    mn[k] = fminf(fminf(f, fminf(g, j)), kk);
    mx[k] = fmaxf(fmaxf(f, fmaxf(g, j)), kk);
  }
  /* accumulation order of the reference: b c i j f e k l h g o n */
  static const int tr[12] = {0, 0, 2, 2, 1, 1, 2, 2, 1, 1, 3, 3};
  static const int tc[12] = {1, 2, 0, 1, 1, 0, 2, 3, 3, 2, 2, 1};
  float aC[3] = {0.0f, 0.0f, 0.0f}, aW = 0.0f;
  for (int i = 0; i < 12; i++)
    easu_tap(aC, &aW, (float)(tc[i] - 1) - ppx, (float)(tr[i] - 1) - ppy, dx, dy, l2x, l2y, lob, clp, t[tr[i]][tc[i]]);
  float rW = 1.0f / aW;
  for (int k = 0; k < 3; k++) pix[k] = fminf(mx[k], fmaxf(mn[k], aC[k] * rW));
}

/* EASU over output rows [y0,y1).  RGBA32F images, pitches in floats.  con = 16 words from *_easu_con. */
void fsr1o_easu_f32(const float* in, int inW, int inH, size_t inPitch, float* out, int outW, int outH,
                    size_t outPitch, const uint32_t* con, int y0, int y1) {
  (void)outH;
  float c0x = u2f(con[0]), c0y = u2f(con[1]), c0z = u2f(con[2]), c0w = u2f(con[3]);
#pragma omp parallel for schedule(static)
  for (int y = y0; y < y1; y++) {
    float ppy = (float)y * c0y + c0w;
    float fy = floorf(ppy); ppy -= fy;
    int ry[4];
    for (int r = 0; r < 4; r++) ry[r] = clampi((int)fy - 1 + r, 0, inH - 1);
    for (int x = 0; x < outW; x++) {
      float ppx = (float)x * c0x + c0z;
      float fx = floorf(ppx); ppx -= fx;
      float t[4][4][3];
      for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) {
        const float* s = in + (size_t)ry[r] * inPitch + (size_t)clampi((int)fx - 1 + c, 0, inW - 1) * 4;
        t[r][c][0] = s[0]; t[r][c][1] = s[1]; t[r][c][2] = s[2];
      }
      float* o = out + (size_t)y * outPitch + (size_t)x * 4;
      easu_pixel(o, ppx, ppy, t);
      o[3] = 1.0f;
    }
  }
}

/* options: bit 0 = FSR_RCAS_DENOISE (ffx_fsr1.h:731-739,761-763), bit 1 = FSR_RCAS_PASSTHROUGH_ALPHA (:688-702) */
void fsr1o_rcas_f32_opt(const float* in, int W, int H, size_t inPitch, float* out, size_t outPitch,
                        const uint32_t* con, int oob_clamp, int y0, int y1, int options) {
  float sharp = u2f(con[0]);
#pragma omp parallel for schedule(static)
  for (int y = y0; y < y1; y++) {
    for (int x = 0; x < W; x++) {
      static const int ox[5] = {0, -1, 0, 1, 0}, oy[5] = {-1, 0, 0, 0, 1}; /* b d e f h */
      float t[5][4];
      for (int i = 0; i < 5; i++) {
        int sx = x + ox[i], sy = y + oy[i];
        if (oob_clamp) { sx = clampi(sx, 0, W - 1); sy = clampi(sy, 0, H - 1); }
        if (sx < 0 || sy < 0 || sx >= W || sy >= H) { t[i][0] = t[i][1] = t[i][2] = t[i][3] = 0.0f; continue; }
        const float* s = in + (size_t)sy * inPitch + (size_t)sx * 4;
        t[i][0] = s[0]; t[i][1] = s[1]; t[i][2] = s[2]; t[i][3] = s[3];
      }
      float lobeC[3];
      for (int k = 0; k < 3; k++) {
        float b = t[0][k], d = t[1][k], e = t[2][k], f = t[3][k], h = t[4][k];
        float mn4 = fminf(fminf(b, fminf(d, f)), h);
        float mx4 = fmaxf(fmaxf(b, fmaxf(d, f)), h);
        float hitMin = fminf(mn4, e) * (1.0f / (4.0f * mx4));
        float hitMax = (1.0f - fmaxf(mx4, e)) * (1.0f / (4.0f * mn4 + -4.0f));
        lobeC[k] = fmaxf(-hitMin, hitMax);
      }
      float lobe = fmaxf(-(float)(0.25 - (1.0 / 16.0)), fminf(fmaxf(lobeC[0], fmaxf(lobeC[1], lobeC[2])), 0.0f)) * sharp;
      if (options & 1) {
        float L[5];
        for (int i = 0; i < 5; i++) L[i] = t[i][2] * 0.5f + (t[i][0] * 0.5f + t[i][1]);
        float bL = L[0], dL = L[1], eL = L[2], fL = L[3], hL = L[4];
        float nz = 0.25f * bL + 0.25f * dL + 0.25f * fL + 0.25f * hL - eL;
        float mx = fmaxf(fmaxf(bL, fmaxf(dL, eL)), fmaxf(fL, hL));   /* AMax3F1(AMax3F1(bL,dL,eL),fL,hL) */
        float mn = fminf(fminf(bL, fminf(dL, eL)), fminf(fL, hL));
        nz = satf(fabsf(nz) * prx_med_rcp(mx - mn));
        nz = -0.5f * nz + 1.0f;
        lobe *= nz;
      }
      float rcpL = prx_med_rcp(4.0f * lobe + 1.0f);
      float* o = out + (size_t)y * outPitch + (size_t)x * 4;
      for (int k = 0; k < 3; k++)
        o[k] = (lobe * t[0][k] + lobe * t[1][k] + lobe * t[4][k] + lobe * t[3][k] + t[2][k]) * rcpL;
      o[3] = (options & 2) ? t[2][3] : 1.0f;
    }
  }
}
void fsr1o_rcas_f32(const float* in, int W, int H, size_t inPitch, float* out, size_t outPitch,
                    const uint32_t* con, int oob_clamp, int y0, int y1) {
  fsr1o_rcas_f32_opt(in, W, H, inPitch, out, outPitch, con, oob_clamp, y0, y1, 0);
}

/* ------------------------------------------------------------------ packed-half path model --- */
/* _Float16 with -fexcess-precision=16: every operation rounds to half, like the H entry points. */
typedef _Float16 h16;
static inline uint16_t h2w(h16 h) { uint16_t w; memcpy(&w, &h, 2); return w; }
static inline h16 w2h(uint16_t w) { h16 h; memcpy(&h, &w, 2); return h; }
static inline h16 hmin(h16 a, h16 b) { return (h16)fminf((float)a, (float)b); }
static inline h16 hmax(h16 a, h16 b) { return (h16)fmaxf((float)a, (float)b); }
static inline h16 habs(h16 a) { return (h16)fabsf((float)a); }
static inline h16 hsat(h16 a) { return (h16)fminf(fmaxf((float)a, 0.0f), 1.0f); }
static inline h16 hrcp(h16 a) { return (h16)1.0 / a; }
static inline h16 hprx_lo_rcp(h16 a) { return w2h((uint16_t)(0x7784u - h2w(a))); }
static inline h16 hprx_lo_rsq(h16 a) { return w2h((uint16_t)(0x59a3u - (h2w(a) >> 1))); }
static inline h16 hprx_med_rcp(h16 a) { h16 b = w2h((uint16_t)(0x778du - h2w(a))); return b * (-b * a + (h16)2.0); }

static inline void easu_set_h(h16* dx, h16* dy, h16* len, h16 w, h16 lA, h16 lB, h16 lC, h16 lD, h16 lE) {
  h16 dc = lD - lC, cb = lC - lB;
  h16 lenX = hrcp(hmax(habs(dc), habs(cb)));          /* exact rcp here, not the bit trick */
  h16 dirX = lD - lB;
  *dx += dirX * w;
  lenX = hsat(habs(dirX) * lenX); lenX *= lenX;
  *len += lenX * w;
  h16 ec = lE - lC, ca = lC - lA;
  h16 lenY = hrcp(hmax(habs(ec), habs(ca)));
  h16 dirY = lE - lA;
  *dy += dirY * w;
  lenY = hsat(habs(dirY) * lenY); lenY *= lenY;
  *len += lenY * w;
}
static inline void easu_tap_h(h16 aC[3], h16* aW, h16 ox, h16 oy, h16 dx, h16 dy, h16 l2x, h16 l2y, h16 lob,
                              h16 clp, const h16 c[3]) {
  h16 vx = ox * dx + oy * dy;
  h16 vy = ox * (-dy) + oy * dx;
  vx *= l2x; vy *= l2y;
  h16 d2 = vx * vx + vy * vy;
  d2 = hmin(d2, clp);
  h16 wB = (h16)(2.0 / 5.0) * d2 + (h16)(-1.0);
  h16 wA = lob * d2 + (h16)(-1.0);
  wB *= wB; wA *= wA;
  wB = (h16)(25.0 / 16.0) * wB + (h16)(-(25.0 / 16.0 - 1.0));
  h16 w = wB * wA;
  aC[0] += c[0] * w; aC[1] += c[1] * w; aC[2] += c[2] * w; *aW += w;
}

/* RGBA16F images (raw half bits), pitches in halves. */
void fsr1o_easu_h16(const uint16_t* in, int inW, int inH, size_t inPitch, uint16_t* out, int outW, int outH,
                    size_t outPitch, const uint32_t* con, int y0, int y1) {
  (void)outH;
  float c0x = u2f(con[0]), c0y = u2f(con[1]), c0z = u2f(con[2]), c0w = u2f(con[3]);
#pragma omp parallel for schedule(static)
  for (int y = y0; y < y1; y++) {
    float fpy = (float)y * c0y + c0w;
    float fy = floorf(fpy); fpy -= fy;
    h16 ppy = (h16)fpy;
    int ry[4];
    for (int r = 0; r < 4; r++) ry[r] = clampi((int)fy - 1 + r, 0, inH - 1);
    for (int x = 0; x < outW; x++) {
      float fpx = (float)x * c0x + c0z;
      float fx = floorf(fpx); fpx -= fx;
      h16 ppx = (h16)fpx;
      h16 t[4][4][3], L[4][4];
      for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) {
        const uint16_t* s = in + (size_t)ry[r] * inPitch + (size_t)clampi((int)fx - 1 + c, 0, inW - 1) * 4;
        t[r][c][0] = w2h(s[0]); t[r][c][1] = w2h(s[1]); t[r][c][2] = w2h(s[2]);
        L[r][c] = t[r][c][2] * (h16)0.5 + (t[r][c][0] * (h16)0.5 + t[r][c][1]);
      }
      /* two packed accumulators: lane .x takes texels f then j, lane .y takes g then k */
      h16 dxa = 0, dya = 0, lena = 0, dxb = 0, dyb = 0, lenb = 0;
      h16 wl = (h16)1.0 + (-ppx), wr = (h16)0.0 + ppx, wt = (h16)1.0 - ppy;
      easu_set_h(&dxa, &dya, &lena, wl * wt, L[0][1], L[1][0], L[1][1], L[1][2], L[2][1]);  /* f */
      easu_set_h(&dxb, &dyb, &lenb, wr * wt, L[0][2], L[1][1], L[1][2], L[1][3], L[2][2]);  /* g */
      easu_set_h(&dxa, &dya, &lena, wl * ppy, L[1][1], L[2][0], L[2][1], L[2][2], L[3][1]); /* j */
      easu_set_h(&dxb, &dyb, &lenb, wr * ppy, L[1][2], L[2][1], L[2][2], L[2][3], L[3][2]); /* k */
      h16 dx = dxa + dxb, dy = dya + dyb, len = lena + lenb;
      h16 dirR = dx * dx + dy * dy;
      int zro = dirR < (h16)(1.0 / 32768.0);
      dirR = hprx_lo_rsq(dirR);
      dirR = zro ? (h16)1.0 : dirR;
      dx = zro ? (h16)1.0 : dx;
      dx *= dirR; dy *= dirR;
      len = len * (h16)0.5; len *= len;
      h16 stretch = (dx * dx + dy * dy) * hprx_lo_rcp(hmax(habs(dx), habs(dy)));
      h16 l2x = (h16)1.0 + (stretch - (h16)1.0) * len;
      h16 l2y = (h16)1.0 + (h16)(-0.5) * len;
      h16 lob = (h16)0.5 + (h16)((1.0 / 4.0 - 0.04) - 0.5) * len;
      h16 clp = hprx_lo_rcp(lob);
      /* six tap PAIRS; lane .x / lane .y accumulate separately then add: (b,c) (i,j) (f,e) (k,l) (h,g) (o,n) */
      static const int pr[6][2] = {{0, 0}, {2, 2}, {1, 1}, {2, 2}, {1, 1}, {3, 3}};
      static const int pc[6][2] = {{1, 2}, {0, 1}, {1, 0}, {2, 3}, {3, 2}, {2, 1}};
      h16 aCx[3] = {0, 0, 0}, aCy[3] = {0, 0, 0}, aWx = 0, aWy = 0;
      for (int i = 0; i < 6; i++) {
        easu_tap_h(aCx, &aWx, (h16)(float)(pc[i][0] - 1) - ppx, (h16)(float)(pr[i][0] - 1) - ppy, dx, dy, l2x, l2y, lob, clp, t[pr[i][0]][pc[i][0]]);
        easu_tap_h(aCy, &aWy, (h16)(float)(pc[i][1] - 1) - ppx, (h16)(float)(pr[i][1] - 1) - ppy, dx, dy, l2x, l2y, lob, clp, t[pr[i][1]][pc[i][1]]);
      }
      h16 rW = hrcp(aWx + aWy);
      uint16_t* o = out + (size_t)y * outPitch + (size_t)x * 4;
      for (int k = 0; k < 3; k++) {
        h16 f = t[1][1][k], g = t[1][2][k], j = t[2][1][k], kk = t[2][2][k];
        /* the (-x,x) packed trick: both.x = max of negatives = -min, both.y = max */
        h16 negmin = hmax(hmax(-f, -g), hmax(-j, -kk));
        h16 mx = hmax(hmax(f, g), hmax(j, kk));
        o[k] = h2w(hmin(mx, hmax(-negmin, (aCx[k] + aCy[k]) * rW)));
      }
      o[3] = 0x3c00;
    }
  }
}

void fsr1o_rcas_h16_opt(const uint16_t* in, int W, int H, size_t inPitch, uint16_t* out, size_t outPitch,
                        const uint32_t* con, int oob_clamp, int y0, int y1, int options) {
  h16 sharp = w2h((uint16_t)(con[1] & 0xffffu));
#pragma omp parallel for schedule(static)
  for (int y = y0; y < y1; y++) {
    for (int x = 0; x < W; x++) {
      static const int ox[5] = {0, -1, 0, 1, 0}, oy[5] = {-1, 0, 0, 0, 1};
      h16 t[5][4];
      for (int i = 0; i < 5; i++) {
        int sx = x + ox[i], sy = y + oy[i];
        if (oob_clamp) { sx = clampi(sx, 0, W - 1); sy = clampi(sy, 0, H - 1); }
        if (sx < 0 || sy < 0 || sx >= W || sy >= H) { t[i][0] = t[i][1] = t[i][2] = t[i][3] = 0; continue; }
        const uint16_t* s = in + (size_t)sy * inPitch + (size_t)sx * 4;
        t[i][0] = w2h(s[0]); t[i][1] = w2h(s[1]); t[i][2] = w2h(s[2]); t[i][3] = w2h(s[3]);
      }
      h16 lobeC[3];
      for (int k = 0; k < 3; k++) {
        h16 b = t[0][k], d = t[1][k], e = t[2][k], f = t[3][k], h = t[4][k];
        h16 mn4 = hmin(hmin(b, hmin(d, f)), h);
        h16 mx4 = hmax(hmax(b, hmax(d, f)), h);
        h16 hitMin = hmin(mn4, e) * hrcp((h16)4.0 * mx4);
        h16 hitMax = ((h16)1.0 - hmax(mx4, e)) * hrcp((h16)4.0 * mn4 + (h16)(-4.0));
        lobeC[k] = hmax(-hitMin, hitMax);
      }
      h16 lobe = hmax((h16)(-(0.25 - (1.0 / 16.0))), hmin(hmax(lobeC[0], hmax(lobeC[1], lobeC[2])), (h16)0.0)) * sharp;
      if (options & 1) {
        h16 L[5];
        for (int i = 0; i < 5; i++) L[i] = t[i][2] * (h16)0.5 + (t[i][0] * (h16)0.5 + t[i][1]);
        h16 bL = L[0], dL = L[1], eL = L[2], fL = L[3], hL = L[4];
        h16 nz = (h16)0.25 * bL + (h16)0.25 * dL + (h16)0.25 * fL + (h16)0.25 * hL - eL;
        h16 mx = hmax(hmax(bL, hmax(dL, eL)), hmax(fL, hL));
        h16 mn = hmin(hmin(bL, hmin(dL, eL)), hmin(fL, hL));
        nz = hsat(habs(nz) * hprx_med_rcp(mx - mn));
        nz = (h16)(-0.5) * nz + (h16)1.0;
        lobe *= nz;
      }
      h16 rcpL = hprx_med_rcp((h16)4.0 * lobe + (h16)1.0);
      uint16_t* o = out + (size_t)y * outPitch + (size_t)x * 4;
      for (int k = 0; k < 3; k++)
        o[k] = h2w((lobe * t[0][k] + lobe * t[1][k] + lobe * t[4][k] + lobe * t[3][k] + t[2][k]) * rcpL);
      o[3] = (options & 2) ? h2w(t[2][3]) : 0x3c00;
    }
  }
}
void fsr1o_rcas_h16(const uint16_t* in, int W, int H, size_t inPitch, uint16_t* out, size_t outPitch,
                    const uint32_t* con, int oob_clamp, int y0, int y1) {
  fsr1o_rcas_h16_opt(in, W, H, inPitch, out, outPitch, con, oob_clamp, y0, y1, 0);
}

/* ------------------------------------------------------------------ synthetic frames --------- */
/* The LCG frame of SURVEY.md §8(c): s=s*1664525+1013904223; v=(s>>8)*2^-24, row-major RGBA. */
void fsr1o_lcg_fill(float* dst, size_t n, uint32_t seed) {
  uint32_t s = seed;
  for (size_t i = 0; i < n; i++) { s = s * 1664525u + 1013904223u; dst[i] = (float)(s >> 8) * (1.0f / 16777216.0f); }
}
/* IEEE round-to-nearest-even conversions for building RGBA16F test frames. */
void fsr1o_f32_to_f16_rne(const float* src, uint16_t* dst, size_t n) { for (size_t i = 0; i < n; i++) dst[i] = h2w((h16)src[i]); }
void fsr1o_f16_to_f32(const uint16_t* src, float* dst, size_t n) { for (size_t i = 0; i < n; i++) dst[i] = (float)w2h(src[i]); }

/* ------------------------------------------------ pointwise companions (fp32) ------------------ */
/* Images are RGBA32F, pitches in floats; alpha is copied through (the reference functions take RGB).
 * aux images (grain, dither) tile with wrap addressing: texel (x mod aw, y mod ah). */
void fsr1o_lfga_f32(const float* in, size_t inPitch, const float* grain, int gw, int gh, size_t gPitch, float* out,
                    size_t outPitch, int W, int H, float amount) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      const float* c = in + (size_t)y * inPitch + (size_t)x * 4;
      const float* t = grain + (size_t)(y % gh) * gPitch + (size_t)(x % gw) * 4;
      float* o = out + (size_t)y * outPitch + (size_t)x * 4;
      for (int k = 0; k < 3; k++) o[k] = c[k] + (t[k] * amount) * fminf(1.0f - c[k], c[k]);
      o[3] = c[3];
    }
}

void fsr1o_srtm_f32(const float* in, size_t inPitch, float* out, size_t outPitch, int W, int H, int inverse) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      const float* c = in + (size_t)y * inPitch + (size_t)x * 4;
      float* o = out + (size_t)y * outPitch + (size_t)x * 4;
      const float m = fmaxf(c[0], fmaxf(c[1], c[2]));
      const float r = inverse ? 1.0f / fmaxf((float)(1.0 / 32768.0), 1.0f - m) : 1.0f / (m + 1.0f);
      for (int k = 0; k < 3; k++) o[k] = c[k] * r;
      o[3] = c[3];
    }
}

float fsr1o_tepd_dit(uint32_t px, uint32_t py, uint32_t frame) {
  float x = (float)(px + frame), y = (float)py;
  const float a = (float)((1.0 + sqrt(5.0)) / 2.0), b = (float)(1.0 / 3.69);
  x = x * a + (y * b);
  return x - floorf(x);
}

static inline float gt_zero(float m) { return satf(m * u2f(0x7f800000u)); } /* 0*inf = NaN -> 0 */

/* dither == NULL: FsrTepdDitF(position, frame); else the .w channel of the tiled dither image, saturated
 * (sample/src/DX12/FSR_Tonemapping.hlsl:87).  bits = 8 or 10. */
void fsr1o_tepd_f32(const float* in, size_t inPitch, const float* dither, int dw, int dh, size_t dPitch, float* out,
                    size_t outPitch, int W, int H, int bits, uint32_t frame) {
  const float q = bits == 8 ? 255.0f : 1023.0f, rq = bits == 8 ? (float)(1.0 / 255.0) : (float)(1.0 / 1023.0);
#pragma omp parallel for schedule(static)
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      const float* c = in + (size_t)y * inPitch + (size_t)x * 4;
      float* o = out + (size_t)y * outPitch + (size_t)x * 4;
      const float dit = dither ? satf(dither[(size_t)(y % dh) * dPitch + (size_t)(x % dw) * 4 + 3])
                               : fsr1o_tepd_dit((uint32_t)x, (uint32_t)y, frame);
      for (int k = 0; k < 3; k++) {
        float n = sqrtf(c[k]);
        n = floorf(n * q) * rq;
        const float a = n * n;
        float b = n + rq;
        b = b * b;
        const float r = (c[k] - b) * prx_med_rcp(a - b);
        o[k] = satf(n + gt_zero(dit - r) * rq);
      }
      o[3] = c[3];
    }
}

/* ------------------------------------------- pointwise companions, half precision -------------- */
/* The H entry points (ffx-fsr/ffx_fsr1.h: FsrLfgaH :1019, FsrSrtmH / FsrSrtmInvH :1049-1050, FsrTepdDitH :1129-1135,
 * FsrTepdC8H / FsrTepdC10H :1137-1153).  The packed Hx2 forms (:1022-1024, :1052-1055, :1156-1199) apply the same
 * operations to two pixels per lane pair, so they produce these bits too (tests/test_oracle.py checks that on the
 * reference build).  Images are RGBA16F (raw half bits), pitches in halves; alpha is copied through; every operation
 * rounds to half (-fexcess-precision=16, no contraction). */
static inline h16 hmax3(h16 a, h16 b, h16 c) { return hmax(a, hmax(b, c)); }

void fsr1o_lfga_h16(const uint16_t* in, size_t inPitch, const uint16_t* grain, int gw, int gh, size_t gPitch, uint16_t* out,
                    size_t outPitch, int W, int H, float amount) {
  const h16 a = (h16)amount, one = (h16)1.0;
#pragma omp parallel for schedule(static)
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      const uint16_t* c = in + (size_t)y * inPitch + (size_t)x * 4;
      const uint16_t* t = grain + (size_t)(y % gh) * gPitch + (size_t)(x % gw) * 4;
      uint16_t* o = out + (size_t)y * outPitch + (size_t)x * 4;
      for (int k = 0; k < 3; k++) {
        const h16 v = w2h(c[k]);
        o[k] = h2w(v + (w2h(t[k]) * a) * hmin(one - v, v));
      }
      o[3] = c[3];
    }
}

void fsr1o_srtm_h16(const uint16_t* in, size_t inPitch, uint16_t* out, size_t outPitch, int W, int H, int inverse) {
  const h16 one = (h16)1.0, tiny = (h16)(1.0 / 32768.0);
#pragma omp parallel for schedule(static)
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      const uint16_t* c = in + (size_t)y * inPitch + (size_t)x * 4;
      uint16_t* o = out + (size_t)y * outPitch + (size_t)x * 4;
      const h16 m = hmax3(w2h(c[0]), w2h(c[1]), w2h(c[2]));
      const h16 r = inverse ? hrcp(hmax(tiny, one - m)) : hrcp(m + one);
      for (int k = 0; k < 3; k++) o[k] = h2w(w2h(c[k]) * r);
      o[3] = c[3];
    }
}

/* the position hash is computed in fp32 (only 32-bit has the precision) and converted once */
uint16_t fsr1o_tepd_dit_h16(uint32_t px, uint32_t py, uint32_t frame) { return h2w((h16)fsr1o_tepd_dit(px, py, frame)); }

void fsr1o_tepd_h16(const uint16_t* in, size_t inPitch, const uint16_t* dither, int dw, int dh, size_t dPitch, uint16_t* out,
                    size_t outPitch, int W, int H, int bits, uint32_t frame) {
  const h16 q = bits == 8 ? (h16)255.0 : (h16)1023.0, rq = bits == 8 ? (h16)(1.0 / 255.0) : (h16)(1.0 / 1023.0);
  const h16 inf = w2h(0x7c00u);
#pragma omp parallel for schedule(static)
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      const uint16_t* c = in + (size_t)y * inPitch + (size_t)x * 4;
      uint16_t* o = out + (size_t)y * outPitch + (size_t)x * 4;
      const h16 dit = dither ? hsat(w2h(dither[(size_t)(y % dh) * dPitch + (size_t)(x % dw) * 4 + 3]))
                             : w2h(fsr1o_tepd_dit_h16((uint32_t)x, (uint32_t)y, frame));
      for (int k = 0; k < 3; k++) {
        const h16 v = w2h(c[k]);
        h16 n = (h16)sqrtf((float)v);
        n = (h16)floorf((float)(n * q)) * rq;
        const h16 a = n * n;
        h16 b = n + rq;
        b = b * b;
        const h16 r = (v - b) * hprx_med_rcp(a - b);
        o[k] = h2w(hsat(n + hsat((dit - r) * inf) * rq));
      }
      o[3] = c[3];
    }
}
