// fsr1_fused.cu — EASU -> RCAS in ONE kernel for RGBA16F images at exactly 2x.
//
// The reference runs two dispatches through a display-sized intermediate texture (sample/src/DX12/FSR_Filter.cpp:121-131,
// intermediate at :72-73).  Here the intermediate never leaves the SM: a CTA walks DOWN a column strip of the output,
// step by step; each step
//   1. takes a TMA box of the input (double-buffered, the next step's box is in flight),
//   2. runs EASU phases 1-3 exactly as easu_h_quad2x_kernel does (same device functions, fsr1_easu_quad.cuh) but stores
//      the 64 x 2n pixels into a shared-memory "mid" tile — rounded to fp16, i.e. the very bits the intermediate image
//      would hold — in the (pixel0, pixel1)-per-channel half2 layout RCAS wants,
//   3. runs RCAS (fsr1_rcas_math.cuh, the arithmetic of rcas_packed_kernel) on the rows of the mid tile whose upper and
//      lower neighbours are present and stores the result with 128-bit stores.
// The last two mid rows of a step stay in shared memory for the next step, so nothing is recomputed vertically inside a
// strip run; horizontally a strip is 31 cells (62 output pixels) of the 32 a warp computes (the 1-pixel apron RCAS needs).
// The result is bit-identical to fsr1_easu + fsr1_rcas (tests/test_gpu_parity.py::test_fused_*): same operations on the same
// rounded intermediate.  Compulsory HBM traffic: bpp (Pin + Pout) = 10 B per output pixel instead of 26.
//
// Work distribution: the image is n_strips x rows; in a linear order (strip-major, units of two rows) CTA c takes the c-th
// of gridDim.x equal shares, i.e. 1-2 "runs" (strip, [ya, yb)).  A run starts like a row slab: its first step also computes the
// cell row above ya (RCAS's upper neighbour), which is the only vertical redundancy (~1 cell row per ~65).
//
// Geometry of a step with cell rows m0 .. m0+n-1 (n <= CY) of strip tx (cells k0 .. k0+31, k0 = 31 tx - 1):
//   mid row index i  <->  pixel row 2 m0 - 1 + i   (i = 0, 1: kept from the previous step; cell row r -> i = 2r+2, 2r+3)
//   mid pair k (lane) <-> pixels (2k+1, 2k+2)
//   RCAS output pair of lane l >= 1: pixels (2k, 2k+1), k = k0 + l:  e = (P[k-1].hi, P[k].lo), d = P[k-1], f = P[k]
//   output rows o in [2 m0, 2 m0 + 2n - 1] ∩ [ya, yb): mid index o - 2 m0 + 1, needs indices -1/+1 around it.
// Out-of-image mid pixels hold 0 (D3D12 Load semantics, ffx_fsr1.h:698-707 through FSR_Pass.hlsl:61); FSR1_FLAG_RCAS_CLAMP is not
// implemented here (the caller falls back to the two-kernel path).
#include "fsr1_easu_quad.cuh"
#include "fsr1_rcas_math.cuh"

namespace fsr1 {

constexpr int kFBW = kQBW + 2;   // box width 38: the strip origin 31 tx - 2 is odd for odd tx and TMA boxes start on 16 bytes
constexpr int kFSW = kFBW - 2;   // texels carrying terms per row
constexpr int kStripCells = 31;  // cells per strip (lanes 1..31 produce RCAS output; lane 0 only feeds its right neighbour)

struct FusedParams {
  ImgView in, out;
  int y0, y1;          // output rows
  uint32_t sharp_h2;   // RCAS con.y
  int n_strips;
  HaloSync sync = {};
};

template <int NW> struct FusedCfg {
  static constexpr int kCY = 2 * NW, kBH = kCY + 3, kSH = kCY + 1, kElems = kFBW * kBH;
  static constexpr int kPad = ((kElems * 8 + 127) / 128) * 128 / 8;
  static constexpr int kMidRows = 2 * kCY + 2;
};

template <int NW> struct __align__(128) FusedSmem {
  uint2 tile[2][FusedCfg<NW>::kPad];
  float4 S[kFSW * FusedCfg<NW>::kSH];
  uint4 mid[FusedCfg<NW>::kMidRows][32];  // (R0R1, G0G1, B0B1, -) of pixel pair (2k+1, 2k+2)
  float L[FusedCfg<NW>::kElems];
  uint64_t bar[2];
};

// EASU output of a quad into the mid tile; pixels outside the image become 0 (what an out-of-image Load returns)
struct MidSink {
  uint4* top;          // &mid[2r+2][lane]
  bool zT, zB;         // pixel row outside the image
  uint32_t keep;       // per-half mask of the pair: 0xffff low = pixel 2k+1 inside, high = pixel 2k+2 inside
  __device__ __forceinline__ void put(bool bottom, __half2 oR, __half2 oG, __half2 oB) const {
    const uint32_t m = (bottom ? zB : zT) ? 0u : keep;
    top[bottom ? 32 : 0] = make_uint4(h22u(oR) & m, h22u(oG) & m, h22u(oB) & m, 0u);
  }
};

struct FusedStep { int tx, m0, n, ya, yb; };

// the CTA's share of the (strip, row-pair) space, cut into runs and steps
template <int CY> struct FusedIter {
  long long pos, hi;
  int rows2, y0, y1;
  int strip, ya, yb, m, mend;
  __device__ void init(const FusedParams& p, int cta, int ctas) {
    y0 = p.y0; y1 = p.y1;
    rows2 = (p.y1 - p.y0 + 1) >> 1;
    const long long total = (long long)p.n_strips * rows2;
    pos = total * cta / ctas;
    hi = total * (cta + 1) / ctas;
    m = 0; mend = -1; strip = 0; ya = yb = 0;
  }
  __device__ bool next(FusedStep& s) {
    if (m > mend) {  // next run
      if (pos >= hi) return false;
      strip = (int)(pos / rows2);
      const int a2 = (int)(pos - (long long)strip * rows2);
      const long long left = hi - pos, room = rows2 - a2;
      const int take = (int)(left < room ? left : room);
      ya = y0 + 2 * a2;
      yb = ya + 2 * take < y1 ? ya + 2 * take : y1;
      pos += take;
      m = (ya - 2) >> 1;     // cell row holding pixel row ya-1 (arithmetic shift = floor)
      mend = (yb - 1) >> 1;  // cell row holding pixel row yb
    }
    s.tx = strip; s.m0 = m; s.ya = ya; s.yb = yb;
    s.n = mend - m + 1 < CY ? mend - m + 1 : CY;
    m += s.n;
    return true;
  }
};

template <int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB)
fused_h_quad2x_kernel(const FusedParams p, const __grid_constant__ CUtensorMap tmap) {
  using C = FusedCfg<NW>;
  constexpr int NT = NW * 32, CY = C::kCY;
  __shared__ FusedSmem<NW> sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&sm.bar[0], 1);
    mbar_init(&sm.bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  halo_sync_begin(p.sync);
  FusedIter<CY> iter;
  iter.init(p, blockIdx.x, gridDim.x);
  FusedStep cur, nxt;
  bool has = iter.next(cur);
  auto box_x = [](const FusedStep& s) { return (kStripCells * s.tx - 2) & ~1; };  // even texel at or before the first tap column
  if (tid == 0 && has) {
    mbar_expect_tx(&sm.bar[0], C::kElems * 8u);
    tma_load_2d(sm.tile[0], &tmap, box_x(cur), cur.m0 - 1 - p.in.row0, &sm.bar[0]);
  }
  const __half2 sharp = uh2(p.sharp_h2);
  for (int it = 0; has; it++) {
    const int b = it & 1;
    const bool hasn = iter.next(nxt);
    if (tid == 0 && hasn) {  // prefetch the next step's box into the other buffer (its readers passed the closing barrier)
      fence_proxy_async();
      mbar_expect_tx(&sm.bar[b ^ 1], C::kElems * 8u);
      tma_load_2d(sm.tile[b ^ 1], &tmap, box_x(nxt), nxt.m0 - 1 - p.in.row0, &sm.bar[b ^ 1]);
    }
    uint2* tile = sm.tile[b];
    const int k0 = kStripCells * cur.tx - 1;       // first cell of the strip (lane 0)
    const int gxe = box_x(cur), dx = (k0 - 1) - gxe;  // box origin; offset of tap column 0 of lane 0 inside it (0 or 1)
    const int gy0 = cur.m0 - 1, n = cur.n;
    mbar_wait(&sm.bar[b], (it >> 1) & 1);
    if (gxe < 0 || gy0 < 0 || gxe + kFBW > p.in.w || gy0 + C::kBH > p.in.h) {
      clamp_fixup(tile, kFBW, kFBW, C::kBH, gxe, gy0, p.in.w, p.in.h, lane, warp, NW);
      fence_proxy_async();
      __syncthreads();
    }
    // phases 1 and 2 on the rows this step needs (n + 3 texel rows, n + 1 rows of terms)
    for (int i = tid; i < kFBW * (n + 3); i += NT) sm.L[i] = texel_luma(tile[i]);
    __syncthreads();
    for (int idx = tid; idx < kFSW * (n + 1); idx += NT) {
      const int j = idx / kFSW, i = idx - j * kFSW;
      const float* c = sm.L + (j + 1) * kFBW + (i + 1);
      sm.S[idx] = texel_terms(c[-kFBW], c[-1], c[0], c[1], c[kFBW]);
    }
    __syncthreads();
    // phase 3: EASU of the step's cells into the mid tile
#pragma unroll 1
    for (int q = 0; q < 2; q++) {
      const int r = warp + q * NW;
      if (r >= n) break;  // warp-uniform
      const int pyT = 2 * (cur.m0 + r) + 1, pxA = 2 * (k0 + lane) + 1;
      MidSink sink;
      sink.top = &sm.mid[2 * r + 2][lane];
      sink.zT = pyT < 0 || pyT >= p.out.h;
      sink.zB = pyT + 1 < 0 || pyT + 1 >= p.out.h;
      sink.keep = ((pxA >= 0 && pxA < p.out.w) ? 0x0000ffffu : 0u) | ((pxA + 1 >= 0 && pxA + 1 < p.out.w) ? 0xffff0000u : 0u);
      quad_compute<MidSink, kFBW, kFSW>(tile + dx, sm.S + dx, lane, r, true, true, sink);
    }
    __syncthreads();
    // RCAS on the rows whose neighbours are in the mid tile: warp w takes output rows 2 m0 + 4w .. + 3 (mid index 4w+1 ..)
    {
      constexpr int kR = 2 * CY / NW;  // rows per warp
      const int o_first = 2 * cur.m0 + warp * kR;
      const int i0 = warp * kR;        // mid index of the row above the warp's first output row
      const int lm = lane > 0 ? lane - 1 : 0;
      const int ox = 2 * (k0 + lane);  // first pixel of the lane's output pair
      const bool writer = lane >= 1 && ox < p.out.w;
      Row3 E[kR + 2], D[kR], Fv[kR];
#pragma unroll
      for (int r = 0; r < kR + 2; r++) {
        const uint4 a = sm.mid[i0 + r][lm], c = sm.mid[i0 + r][lane];
        E[r].r = uh2(__byte_perm(a.x, c.x, 0x5432));
        E[r].g = uh2(__byte_perm(a.y, c.y, 0x5432));
        E[r].b = uh2(__byte_perm(a.z, c.z, 0x5432));
        if (r >= 1 && r <= kR) {
          D[r - 1].r = uh2(a.x); D[r - 1].g = uh2(a.y); D[r - 1].b = uh2(a.z);
          Fv[r - 1].r = uh2(c.x); Fv[r - 1].g = uh2(c.y); Fv[r - 1].b = uh2(c.z);
        }
      }
      unsigned char* dst = p.out.base + (long long)(o_first - p.out.row0) * p.out.pitch + (long long)ox * 8;
#pragma unroll
      for (int r = 0; r < kR; r++) {
        const int o = o_first + r;
        if (o >= cur.ya && o < cur.yb && o < 2 * cur.m0 + 2 * n) {  // warp-uniform
          __half2 oR, oG, oB;
          rcas_pair<0>(E[r], D[r], E[r + 1], Fv[r], E[r + 2], sharp, oR, oG, oB);
          if (writer) {
            const uint4 w = pack_pair_half(oR, oG, oB, 0x3c003c00u);
            unsigned char* o8 = dst + (long long)r * p.out.pitch;
            if (ox + 1 < p.out.w) *reinterpret_cast<uint4*>(o8) = w;
            else *reinterpret_cast<uint2*>(o8) = make_uint2(w.x, w.y);
          }
        }
      }
    }
    __syncthreads();
    if (n == CY && tid < 64) {  // the run may continue: its last two mid rows become rows 0, 1 of the next step
      const int rr = tid >> 5;
      sm.mid[rr][lane] = sm.mid[2 * CY + rr][lane];
    }
    // (no barrier needed here: the next writers of mid rows >= 2 and the next readers of rows 0, 1 are both behind the
    //  two __syncthreads of the next iteration's phases 1 and 2)
    cur = nxt;
    has = hasn;
  }
  halo_sync_end(p.sync);
}

#ifndef FSR1_CPU_EMU
cudaError_t launch_fused_h(const EasuParams& e, uint32_t sharp_h2, int clamp, cudaStream_t s, const char** name) {
  if (clamp) return cudaErrorNotSupported;
  if (!(e.c0x == 0.5f && e.c0y == 0.5f && e.c0z == -0.25f && e.c0w == -0.25f)) return cudaErrorNotSupported;
  if ((reinterpret_cast<uintptr_t>(e.in.base) & 15) || (e.in.pitch & 15) || (reinterpret_cast<uintptr_t>(e.out.base) & 15) || (e.out.pitch & 15))
    return cudaErrorNotSupported;
  constexpr int NW = 4;
  using C = FusedCfg<NW>;
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return cudaErrorNotSupported;
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {(cuuint64_t)e.in.w, (cuuint64_t)e.in.rows};
  const cuuint64_t strides[1] = {(cuuint64_t)e.in.pitch};
  const cuuint32_t box[2] = {(cuuint32_t)kFBW, (cuuint32_t)C::kBH};
  const cuuint32_t estr[2] = {1, 1};
  if (encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, e.in.base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return cudaErrorNotSupported;
  FusedParams p;
  p.in = e.in; p.out = e.out; p.y0 = e.y0; p.y1 = e.y1; p.sharp_h2 = sharp_h2; p.sync = e.sync;
  // output pairs (2k, 2k+1), k = 0 .. (w-1)/2, 31 per strip
  p.n_strips = ((e.out.w + 1) / 2 + kStripCells - 1) / kStripCells;
  const long long units = (long long)p.n_strips * ((e.y1 - e.y0 + 1) / 2);
  constexpr int kPerSM = 6;
  long long grid = (long long)kPerSM * sm_count();
  if (grid > (units + 7) / 8) grid = (units + 7) / 8;  // at least ~16 rows of a strip per CTA
  if (grid < 1) grid = 1;
  fused_h_quad2x_kernel<NW, kPerSM><<<(int)grid, NW * 32, 0, s>>>(p, tmap);
  *name = "fused_easu_rcas_h_quad2x<4w,6/sm,tma2,strips>";
  return cudaGetLastError();
}
#endif

}  // namespace fsr1
