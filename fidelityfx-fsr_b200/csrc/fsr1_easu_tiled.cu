// fsr1_easu_tiled.cu — the production EASU kernel for RGBA16F images on sm_100a.
//
// One CTA produces a 64x16 tile of the output.  Its input footprint (tile*scale + 3..4 texel halo,
// the box size is fixed per launch) is fetched by ONE TMA 2D tile load (cp.async.bulk.tensor, elected
// thread, mbarrier completion) into shared memory; out-of-image parts of the box arrive as zeros and
// are rewritten to clamp-to-edge (the reference samples through a CLAMP sampler,
// sample/src/DX12/FSR_Filter.cpp:48-53).  The work is then split the way the arithmetic wants it:
//
//   phase 1  per INPUT texel   2*luma in fp32                                    (ffx_fsr1.h:363-366)
//   phase 2  per INPUT texel   the FsrEasuSetF terms that do not depend on the output pixel:
//                              dirX, dirY, lenX+lenY (fp32, F-path bit tricks)      (ffx_fsr1.h:295-313)
//   phase 3  per OUTPUT pixel  bilinear blend of the 4 nearest texels' terms in fp32 (ffx_fsr1.h:383-386),
//                              then everything else in packed half2 with TWO horizontally adjacent
//                              output pixels per lane: normalise / stretch / lobe / clip
//                              (ffx_fsr1.h:389-409), the 12 taps (ffx_fsr1.h:423-434) and the
//                              de-ringing clamp (:416-419,437); one 128-bit store per pixel pair.
//
// Why fp32 for phases 1-2 and the blend: the edge direction is a normalised difference of lumas; in
// half precision it is ill-conditioned wherever the gradient nearly cancels, and the result then
// differs from the fp32 algorithm by up to 0.1 (measured on the CPU model, DESIGN.md "numerics").  With
// fp32 analysis and half2 taps the kernel stays within 4e-3 of the fp32 oracle (tolerance 1e-2).
// Hoisting phases 1-2 to once per input texel also removes ~40% of the per-pixel arithmetic at 2x.
//
// Tap weights use the expanded quadratic form of the rotated/scaled distance
//   d2(ox,oy) = qa*ox^2 + qb*ox*oy + qc*oy^2,   qa = (dx*l2x)^2+(dy*l2y)^2, qc = (dy*l2x)^2+(dx*l2y)^2,
//   qb = 2*dx*dy*(l2x^2-l2y^2)
// so the 12 taps need 2 half2 operations each for d2 instead of 6; the window polynomial
//   (25/16*(2/5*d2-1)^2 - 9/16) * (lob*d2-1)^2  is evaluated as ((d2/4-5/4)*d2+1) * (lob*d2-1)^2.
#include <cuda.h>
#include "fsr1_common.cuh"

namespace fsr1 {

constexpr int kTileW = 64;  // output pixels per CTA in x
constexpr int kTileH = 16;  // ... in y
constexpr int kThreads = 256;

// ---- PTX wrappers: mbarrier + TMA ------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
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

// ---- half2 helpers ----------------------------------------------------------------------------------
__device__ __forceinline__ __half2 u2h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t h22u(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 h2c(float v) { return __float2half2_rn(v); }
// F-path bit tricks applied per lane through fp32 (the H-path magic numbers give a different
// approximation and break the 1e-2 bound)
__device__ __forceinline__ __half2 prx_lo_rcp_h2(__half2 a) {
  const float2 f = __half22float2(a);
  return __floats2half2_rn(prx_lo_rcp(f.x), prx_lo_rcp(f.y));
}
__device__ __forceinline__ __half2 prx_lo_rsq_h2(__half2 a) {
  const float2 f = __half22float2(a);
  return __floats2half2_rn(prx_lo_rsq(f.x), prx_lo_rsq(f.y));
}

struct PixelTerms { float dx, dy, len; };

// fp32 bilinear blend of the per-texel terms of f,g,j,k (reference order f,g,j,k)
__device__ __forceinline__ PixelTerms blend_terms(const float4* __restrict__ S, int idx, int stride, float ppx,
                                                  float ppy) {
  const float4 f = S[idx], g = S[idx + 1], j = S[idx + stride], k = S[idx + stride + 1];
  const float ipx = 1.0f - ppx, ipy = 1.0f - ppy;
  const float wf = ipx * ipy, wg = ppx * ipy, wj = ipx * ppy, wk = ppx * ppy;
  PixelTerms t;
  t.dx = fmaf(k.x, wk, fmaf(j.x, wj, fmaf(g.x, wg, f.x * wf)));
  t.dy = fmaf(k.y, wk, fmaf(j.y, wj, fmaf(g.y, wg, f.y * wf)));
  t.len = fmaf(k.z, wk, fmaf(j.z, wj, fmaf(g.z, wg, f.z * wf)));
  return t;
}

// Shared-memory carve-up (dynamic): [tile BH*BW uint2][luma BH*BW float][terms (BH-2)*(BW-2) float4][mbarrier]
struct Smem {
  uint2* tile;
  float* luma;
  float4* terms;
  uint64_t* bar;
};
__host__ __device__ inline size_t smem_bytes(int BW, int BH) {
  size_t n = (size_t)BW * BH;
  size_t off = n * 8;                 // tile, 128B aligned at 0
  off = (off + 15) & ~(size_t)15;
  off += n * 4;                       // luma
  off = (off + 15) & ~(size_t)15;
  off += (size_t)(BW - 2) * (BH - 2) * 16;  // terms
  off = (off + 15) & ~(size_t)15;
  off += 16;                          // barrier
  return off + 128;                   // slack for manual 128B alignment of the base
}

__global__ void __launch_bounds__(kThreads, 2)
easu_h_tiled_kernel(const EasuParams p, const __grid_constant__ CUtensorMap tmap, const int BW, const int BH) {
  extern __shared__ unsigned char smem_raw[];
  // 128-byte align the carve-up by OFFSET (pointer arithmetic on the shared array keeps the
  // address space, so the accesses below compile to LDS/STS, not generic LD/ST)
  unsigned char* base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  const int n = BW * BH;
  uint2* tile = reinterpret_cast<uint2*>(base);
  size_t off = ((size_t)n * 8 + 15) & ~(size_t)15;
  float* L = reinterpret_cast<float*>(base + off);
  off = (off + (size_t)n * 4 + 15) & ~(size_t)15;
  float4* S = reinterpret_cast<float4*>(base + off);
  off = (off + (size_t)(BW - 2) * (BH - 2) * 16 + 15) & ~(size_t)15;
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + off);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ox0 = blockIdx.x * kTileW, oy0 = p.y0 + blockIdx.y * kTileH;
  int fx0, fy0;
  float dummy;
  easu_pos(ox0, p.c0x, p.c0z, fx0, dummy);
  easu_pos(oy0, p.c0y, p.c0w, fy0, dummy);
  fx0 -= 1;  // box origin = first tap column/row of the tile's first pixel
  fy0 -= 1;

  if (tid == 0) {
    mbar_init(bar, 1);
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(bar, (uint32_t)n * 8u);
    tma_load_2d(tile, &tmap, fx0, fy0 - p.in.row0, bar);
  }
  mbar_wait(bar, 0);

  // clamp-to-edge fix-up of the zero-filled out-of-image part of the box (border tiles only)
  const bool border = fx0 < 0 || fy0 < 0 || fx0 + BW > p.in.w || fy0 + BH > p.in.h;
  if (border) {
    for (int j = warp; j < BH; j += kThreads / 32) {
      const int gy = fy0 + j, cy = clampi(gy, 0, p.in.h - 1) - fy0;
      for (int i = lane; i < BW; i += 32) {
        const int gx = fx0 + i, cx = clampi(gx, 0, p.in.w - 1) - fx0;
        if ((cx != i || cy != j) && cx >= 0 && cx < BW && cy >= 0 && cy < BH) tile[j * BW + i] = tile[cy * BW + cx];
      }
    }
    __syncthreads();
  }

  // phase 1: 2*luma per texel, fp32 (exact: inputs are halves)
  for (int i = tid; i < n; i += kThreads) {
    const uint2 t = tile[i];
    const float2 rg = __half22float2(u2h2(t.x));
    const float b = __low2float(u2h2(t.y));
    L[i] = fmaf(b, 0.5f, fmaf(rg.x, 0.5f, rg.y));
  }
  __syncthreads();

  // phase 2: per-texel direction / length terms for the inner texels (those that can be f,g,j,k)
  const int SW = BW - 2;
  for (int j = 1 + warp; j < BH - 1; j += kThreads / 32) {
    for (int i = 1 + lane; i < BW - 1; i += 32) {
      const float lC = L[j * BW + i], lB = L[j * BW + i - 1], lD = L[j * BW + i + 1];
      const float lA = L[(j - 1) * BW + i], lE = L[(j + 1) * BW + i];
      const float dirX = lD - lB, dirY = lE - lA;
      float lenX = sat(fabsf(dirX) * prx_lo_rcp(fmaxf(fabsf(lD - lC), fabsf(lC - lB))));
      float lenY = sat(fabsf(dirY) * prx_lo_rcp(fmaxf(fabsf(lE - lC), fabsf(lC - lA))));
      S[(j - 1) * SW + (i - 1)] = make_float4(dirX, dirY, fmaf(lenX, lenX, lenY * lenY), 0.0f);
    }
  }
  __syncthreads();

  // phase 3: two pixel pairs per thread: columns (2*lane, 2*lane+1), rows warp and warp+8
  const __half2 kZero = h2c(0.0f), kOne = h2c(1.0f);
#pragma unroll 1
  for (int pr = 0; pr < kTileH / 8; pr++) {
    const int ox = ox0 + lane * 2, oy = oy0 + warp + pr * 8;
    if (ox >= p.out.w || oy >= p.y1) continue;
    int fxA, fxB, fy;
    float ppxA, ppxB, ppy;
    easu_pos(ox, p.c0x, p.c0z, fxA, ppxA);
    easu_pos(ox + 1 < p.out.w ? ox + 1 : ox, p.c0x, p.c0z, fxB, ppxB);  // odd width: B duplicates A
    easu_pos(oy, p.c0y, p.c0w, fy, ppy);
    const int cy = fy - fy0, cxA = fxA - fx0, cxB = fxB - fx0;  // >= 1 by construction
    const PixelTerms tA = blend_terms(S, (cy - 1) * SW + (cxA - 1), SW, ppxA, ppy);
    const PixelTerms tB = blend_terms(S, (cy - 1) * SW + (cxB - 1), SW, ppxB, ppy);

    // ---- per-pixel kernel shape, packed (A,B) ----
    __half2 dx = __floats2half2_rn(tA.dx, tB.dx), dy = __floats2half2_rn(tA.dy, tB.dy);
    __half2 len = __floats2half2_rn(tA.len, tB.len);
    __half2 dirR = __hfma2(dx, dx, __hmul2(dy, dy));
    const uint32_t zro = __hlt2_mask(dirR, h2c(1.0f / 32768.0f));
    __half2 rs = prx_lo_rsq_h2(dirR);
    rs = u2h2((h22u(rs) & ~zro) | (h22u(kOne) & zro));
    dx = u2h2((h22u(dx) & ~zro) | (h22u(kOne) & zro));
    dx = __hmul2(dx, rs);
    dy = __hmul2(dy, rs);
    len = __hmul2(len, h2c(0.5f));
    len = __hmul2(len, len);
    const __half2 dx2 = __hmul2(dx, dx), dy2 = __hmul2(dy, dy);
    const __half2 stretch =
        __hmul2(__hadd2(dx2, dy2), prx_lo_rcp_h2(__hmax2(__habs2(dx), __habs2(dy))));
    const __half2 l2x = __hfma2(__hsub2(stretch, kOne), len, kOne);
    const __half2 l2y = __hfma2(h2c(-0.5f), len, kOne);
    const __half2 lob = __hfma2(h2c((float)((1.0 / 4.0 - 0.04) - 0.5)), len, h2c(0.5f));
    const __half2 clp = prx_lo_rcp_h2(lob);
    const __half2 X2 = __hmul2(l2x, l2x), Y2 = __hmul2(l2y, l2y);
    const __half2 qa = __hfma2(X2, dx2, __hmul2(Y2, dy2));
    const __half2 qc = __hfma2(X2, dy2, __hmul2(Y2, dx2));
    const __half2 qb = __hmul2(__hmul2(__hadd2(dx, dx), dy), __hsub2(X2, Y2));

    // per-column / per-row pieces of d2:  d2(k,r) = PX[k] + QY[r] + SB[k]*oy[r]
    const __half2 ppx2 = __floats2half2_rn(ppxA, ppxB), ppy2 = __float2half2_rn(ppy);
    __half2 PX[4], SB[4], QY[4], OY[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const __half2 oxk = __hsub2(h2c((float)(k - 1)), ppx2);
      SB[k] = __hmul2(qb, oxk);
      PX[k] = __hmul2(__hmul2(qa, oxk), oxk);
      OY[k] = __hsub2(h2c((float)(k - 1)), ppy2);
      QY[k] = __hmul2(__hmul2(qc, OY[k]), OY[k]);
    }

    const uint2* tA0 = tile + (cy - 1) * BW + (cxA - 1);
    const uint2* tB0 = tile + (cy - 1) * BW + (cxB - 1);
    __half2 aRG_A = kZero, aBA_A = kZero, aRG_B = kZero, aBA_B = kZero, aW = kZero;
    __half2 mnRG_A, mnBA_A, mxRG_A, mxBA_A, mnRG_B, mnBA_B, mxRG_B, mxBA_B;
    const __half2 c025 = h2c(0.25f), cm125 = h2c(-1.25f), cm1 = h2c(-1.0f);

#define FSR1_TAP(R, K)                                                                             \
    {                                                                                              \
      const uint2 ca = tA0[(R) * BW + (K)], cb = tB0[(R) * BW + (K)];                              \
      __half2 d2 = __hfma2(SB[K], OY[R], __hadd2(PX[K], QY[R]));                                   \
      d2 = __hmin2(d2, clp);                                                                       \
      const __half2 wb = __hfma2(__hfma2(c025, d2, cm125), d2, kOne);                              \
      __half2 wa = __hfma2(lob, d2, cm1);                                                          \
      wa = __hmul2(wa, wa);                                                                        \
      const __half2 w = __hmul2(wb, wa);                                                           \
      const __half2 wA2 = __low2half2(w), wB2 = __high2half2(w);                                   \
      aRG_A = __hfma2(u2h2(ca.x), wA2, aRG_A);                                                     \
      aBA_A = __hfma2(u2h2(ca.y), wA2, aBA_A);                                                     \
      aRG_B = __hfma2(u2h2(cb.x), wB2, aRG_B);                                                     \
      aBA_B = __hfma2(u2h2(cb.y), wB2, aBA_B);                                                     \
      aW = __hadd2(aW, w);                                                                         \
      if ((R) == 1 && (K) == 1) {                                                                  \
        mnRG_A = mxRG_A = u2h2(ca.x); mnBA_A = mxBA_A = u2h2(ca.y);                                \
        mnRG_B = mxRG_B = u2h2(cb.x); mnBA_B = mxBA_B = u2h2(cb.y);                                \
      } else if (((R) == 1 || (R) == 2) && ((K) == 1 || (K) == 2)) {                               \
        mnRG_A = __hmin2(mnRG_A, u2h2(ca.x)); mxRG_A = __hmax2(mxRG_A, u2h2(ca.x));                \
        mnBA_A = __hmin2(mnBA_A, u2h2(ca.y)); mxBA_A = __hmax2(mxBA_A, u2h2(ca.y));                \
        mnRG_B = __hmin2(mnRG_B, u2h2(cb.x)); mxRG_B = __hmax2(mxRG_B, u2h2(cb.x));                \
        mnBA_B = __hmin2(mnBA_B, u2h2(cb.y)); mxBA_B = __hmax2(mxBA_B, u2h2(cb.y));                \
      }                                                                                            \
    }
    FSR1_TAP(1, 1) FSR1_TAP(1, 2) FSR1_TAP(2, 1) FSR1_TAP(2, 2)   // f g j k
    FSR1_TAP(0, 1) FSR1_TAP(0, 2)                                 // b c
    FSR1_TAP(1, 0) FSR1_TAP(1, 3)                                 // e h
    FSR1_TAP(2, 0) FSR1_TAP(2, 3)                                 // i l
    FSR1_TAP(3, 1) FSR1_TAP(3, 2)                                 // n o
#undef FSR1_TAP

    const float2 aWf = __half22float2(aW);
    const __half2 rA = __float2half2_rn(__frcp_rn(aWf.x)), rB = __float2half2_rn(__frcp_rn(aWf.y));
    __half2 oRG_A = __hmin2(mxRG_A, __hmax2(mnRG_A, __hmul2(aRG_A, rA)));
    __half2 oBA_A = __hmin2(mxBA_A, __hmax2(mnBA_A, __hmul2(aBA_A, rA)));
    __half2 oRG_B = __hmin2(mxRG_B, __hmax2(mnRG_B, __hmul2(aRG_B, rB)));
    __half2 oBA_B = __hmin2(mxBA_B, __hmax2(mnBA_B, __hmul2(aBA_B, rB)));
    oBA_A = __halves2half2(__low2half(oBA_A), __float2half_rn(1.0f));  // alpha = 1 (FSR_Pass.hlsl:95)
    oBA_B = __halves2half2(__low2half(oBA_B), __float2half_rn(1.0f));
    unsigned char* orow = p.out.base + (long long)(oy - p.out.row0) * p.out.pitch;
    if (ox + 1 < p.out.w) {
      *reinterpret_cast<uint4*>(orow + (size_t)ox * 8) = make_uint4(h22u(oRG_A), h22u(oBA_A), h22u(oRG_B), h22u(oBA_B));
    } else {
      *reinterpret_cast<uint2*>(orow + (size_t)ox * 8) = make_uint2(h22u(oRG_A), h22u(oBA_A));
    }
  }
}

// ---- host side ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// Same float arithmetic as easu_pos on the device.
static inline int host_fp(int o, float scale, float offset) {
  volatile float m = (float)o * scale;
  volatile float s = m + offset;
  return (int)floorf(s);
}

// Largest footprint (in texels) any tile of `tile` output pixels needs along one axis.
static int max_footprint(int n_out, int first, int tile, float scale, float offset) {
  int best = 4;
  for (int o0 = first; o0 < n_out; o0 += tile) {
    const int o1 = (o0 + tile - 1 < n_out - 1) ? o0 + tile - 1 : n_out - 1;
    const int span = host_fp(o1, scale, offset) - host_fp(o0, scale, offset) + 4;
    if (span > best) best = span;
  }
  return best;
}

cudaError_t launch_easu_h_tiled(const EasuParams& p, cudaStream_t s, const char** name) {
  // layout requirements of TMA and of the 128-bit stores
  if ((reinterpret_cast<uintptr_t>(p.in.base) & 15) || (p.in.pitch & 15) || (reinterpret_cast<uintptr_t>(p.out.base) & 15) ||
      (p.out.pitch & 15))
    return cudaErrorNotSupported;
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return cudaErrorNotSupported;
  int BW = max_footprint(p.out.w, 0, kTileW, p.c0x, p.c0z);
  int BH = max_footprint(p.y1, p.y0, kTileH, p.c0y, p.c0w);
  BW = (BW + 1) & ~1;  // inner box extent must be a multiple of 16 bytes
  if (BW > 256 || BH > 256) return cudaErrorNotSupported;
  const size_t smem = smem_bytes(BW, BH);
  if (smem > 200 * 1024) return cudaErrorNotSupported;

  CUtensorMap tmap;
  const cuuint64_t dims[2] = {(cuuint64_t)p.in.w, (cuuint64_t)p.in.rows};
  const cuuint64_t strides[1] = {(cuuint64_t)p.in.pitch};
  const cuuint32_t box[2] = {(cuuint32_t)BW, (cuuint32_t)BH};
  const cuuint32_t estr[2] = {1, 1};
  // one RGBA16F texel = one 64-bit element
  CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, p.in.base, dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return cudaErrorNotSupported;

  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(easu_h_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  const dim3 grid((p.out.w + kTileW - 1) / kTileW, (p.y1 - p.y0 + kTileH - 1) / kTileH, 1);
  easu_h_tiled_kernel<<<grid, kThreads, smem, s>>>(p, tmap, BW, BH);
  *name = "easu_h_tiled<64x16,tma>";
  return cudaGetLastError();
}

}  // namespace fsr1
