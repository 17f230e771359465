// fsr1_rcas_packed.cu — the production RCAS kernel for RGBA16F images on sm_100a.
//
// RCAS is a 5-tap cross (b above, d left, e centre, f right, h below; ffx-fsr/ffx_fsr1.h:693-707) with
// ~45 packed operations per pixel and 16 bytes of compulsory traffic per pixel: it is the HBM-bound
// half of the path, so the kernel is organised around the memory system, not shared memory:
//   * each lane owns TWO horizontally adjacent pixels (one 128-bit load / store per row) and walks
//     kRows rows downwards keeping a rolling 3-row window in registers, so a row is loaded once
//     per (kRows+2)/kRows of its uses and every global access is a full 16-byte vector;
//   * the left/right neighbours (d, f) come from the adjacent lanes by warp shuffle; only lane 0 /
//     lane 31 fetch their missing neighbour pixel from memory;
//   * all arithmetic is half2 over the lane's two pixels, in the structure-of-arrays form of the
//     reference's FsrRcasHx2 (ffx_fsr1.h:888-984): (R0,R1) (G0,G1) (B0,B1).
// Numerics: the six "high precision" reciprocals (ffx_fsr1.h:750-755) are rcp.approx.f32 on the
// unpacked halves (h2rcp); the resolve reciprocal is the packed APrxMedRcpH2 (ffx_a.h:1815).  Measured
// against the fp32 oracle on the same half input: <= 2e-3 (tolerance 1e-2).
// min/max are the non-propagating half2 forms, so the 0*inf NaNs of flat black/white neighbourhoods
// drop out exactly as with HLSL min/max (ffx_fsr1.h:756-759).
#include "fsr1_common.cuh"

namespace fsr1 {

constexpr int kRows = 4;       // rows walked by one lane
constexpr int kWarps = 8;      // warps per CTA -> CTA covers 64 x 32 pixels

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

// Two pixels (x, x+1) of logical row y with the RCAS out-of-image rule applied.
__device__ __forceinline__ Row3 load_pair(const RcasParams& p, int x, int y) {
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (p.clamp) y = clampi(y, 0, p.in.h - 1);
  if (y >= 0 && y < p.in.h) {
    const unsigned char* row = p.in.base + (long long)(y - p.in.row0) * p.in.pitch;
    if (x + 1 < p.in.w) {
      v = __ldg(reinterpret_cast<const uint4*>(row + (size_t)x * 8));
    } else if (p.clamp) {  // right edge: both pixels clamp into the row
      const uint2 t0 = __ldg(reinterpret_cast<const uint2*>(row) + min(x, p.in.w - 1));
      const uint2 t1 = __ldg(reinterpret_cast<const uint2*>(row) + (p.in.w - 1));
      v = make_uint4(t0.x, t0.y, t1.x, t1.y);
    } else if (x < p.in.w) {  // odd width: second pixel is outside and reads 0
      const uint2 t = __ldg(reinterpret_cast<const uint2*>(row) + x);
      v.x = t.x; v.y = t.y;
    }
  }
  return to_soa(v);
}

// One pixel (x,y) broadcast to both halves of each channel register (used for the warp-edge neighbours).
__device__ __forceinline__ Row3 load_single(const RcasParams& p, int x, int y) {
  uint2 t = make_uint2(0u, 0u);
  if (p.clamp) { x = clampi(x, 0, p.in.w - 1); y = clampi(y, 0, p.in.h - 1); }
  if (x >= 0 && x < p.in.w && y >= 0 && y < p.in.h)
    t = __ldg(reinterpret_cast<const uint2*>(p.in.base + (long long)(y - p.in.row0) * p.in.pitch + (size_t)x * 8));
  Row3 o;
  o.r = uh2(__byte_perm(t.x, t.x, 0x1010));
  o.g = uh2(__byte_perm(t.x, t.x, 0x3232));
  o.b = uh2(__byte_perm(t.y, t.y, 0x1010));
  return o;
}

__device__ __forceinline__ __half2 lobe_channel(__half2 b, __half2 d, __half2 e, __half2 f, __half2 h) {
  const __half2 mn4 = __hmin2(__hmin2(b, d), __hmin2(f, h));
  const __half2 mx4 = __hmax2(__hmax2(b, d), __hmax2(f, h));
  const __half2 k4 = __float2half2_rn(4.0f), k1 = __float2half2_rn(1.0f), km4 = __float2half2_rn(-4.0f);
  const __half2 hitMin = __hmul2(__hmin2(mn4, e), h2rcp(__hmul2(k4, mx4)));
  const __half2 hitMax = __hmul2(__hsub2(k1, __hmax2(mx4, e)), h2rcp(__hfma2(k4, mn4, km4)));
  return __hmax2(__hneg2(hitMin), hitMax);
}

__device__ __forceinline__ __half2 resolve_channel(__half2 lobe, __half2 rcpL, __half2 b, __half2 d, __half2 e,
                                                   __half2 f, __half2 h) {
  const __half2 ring = __hadd2(__hadd2(b, d), __hadd2(h, f));
  return __hmul2(__hfma2(lobe, ring, e), rcpL);
}

__global__ void __launch_bounds__(32 * kWarps) rcas_h_packed_kernel(const RcasParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x = (blockIdx.x * 32 + lane) * 2;
  const int ys = p.y0 + (blockIdx.y * kWarps + warp) * kRows;
  if (ys >= p.y1) return;  // whole warp
  const __half2 sharp = uh2(p.sharp_h2);
  const __half2 kLimit = __float2half2_rn(-0.1875f), kZero = __float2half2_rn(0.0f);
  const uint32_t one = 0x3c003c00u;

  Row3 prev = load_pair(p, x, ys - 1), cur = load_pair(p, x, ys);
#pragma unroll
  for (int r = 0; r < kRows; r++) {
    const int y = ys + r;
    if (y >= p.y1) break;  // warp-uniform
    const Row3 next = load_pair(p, x, y + 1);
    // neighbours from the adjacent lanes; the warp's outer lanes read theirs from memory
    Row3 lf, rt;
    lf.r = uh2(__shfl_up_sync(0xffffffffu, hu2(cur.r), 1));
    lf.g = uh2(__shfl_up_sync(0xffffffffu, hu2(cur.g), 1));
    lf.b = uh2(__shfl_up_sync(0xffffffffu, hu2(cur.b), 1));
    rt.r = uh2(__shfl_down_sync(0xffffffffu, hu2(cur.r), 1));
    rt.g = uh2(__shfl_down_sync(0xffffffffu, hu2(cur.g), 1));
    rt.b = uh2(__shfl_down_sync(0xffffffffu, hu2(cur.b), 1));
    if (lane == 0) lf = load_single(p, x - 1, y);
    if (lane == 31) rt = load_single(p, x + 2, y);
    // d = (left.px1, e.px0), f = (e.px1, right.px0)
    const __half2 dR = uh2(__byte_perm(hu2(lf.r), hu2(cur.r), 0x5432)), fR = uh2(__byte_perm(hu2(cur.r), hu2(rt.r), 0x5432));
    const __half2 dG = uh2(__byte_perm(hu2(lf.g), hu2(cur.g), 0x5432)), fG = uh2(__byte_perm(hu2(cur.g), hu2(rt.g), 0x5432));
    const __half2 dB = uh2(__byte_perm(hu2(lf.b), hu2(cur.b), 0x5432)), fB = uh2(__byte_perm(hu2(cur.b), hu2(rt.b), 0x5432));

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

    if (x < p.out.w) {
      unsigned char* orow = p.out.base + (long long)(y - p.out.row0) * p.out.pitch + (size_t)x * 8;
      const uint32_t rg0 = __byte_perm(hu2(oR), hu2(oG), 0x5410), b0 = __byte_perm(hu2(oB), one, 0x5410);
      if (x + 1 < p.out.w) {
        const uint32_t rg1 = __byte_perm(hu2(oR), hu2(oG), 0x7632), b1 = __byte_perm(hu2(oB), one, 0x7632);
        *reinterpret_cast<uint4*>(orow) = make_uint4(rg0, b0, rg1, b1);
      } else {
        *reinterpret_cast<uint2*>(orow) = make_uint2(rg0, b0);
      }
    }
    prev = cur;
    cur = next;
  }
}

cudaError_t launch_rcas_h_packed(const RcasParams& p, cudaStream_t s, const char** name) {
  if ((reinterpret_cast<uintptr_t>(p.in.base) & 15) || (p.in.pitch & 15) || (reinterpret_cast<uintptr_t>(p.out.base) & 15) ||
      (p.out.pitch & 15))
    return cudaErrorNotSupported;
  const dim3 grid((p.out.w + 63) / 64, (p.y1 - p.y0 + kWarps * kRows - 1) / (kWarps * kRows), 1);
  rcas_h_packed_kernel<<<grid, 32 * kWarps, 0, s>>>(p);
  *name = "rcas_h_packed<2px,4rows,shfl>";
  return cudaGetLastError();
}

}  // namespace fsr1
