// TEST INFRASTRUCTURE — runs the 2x EASU kernel of csrc/fsr1_easu_tiled.cu on CPU threads (see include/cuda_emu.h).
// The .cu file is compiled AS IS with -DFSR1_CPU_EMU (which only swaps the PTX wrappers and drops the host launcher);
// this harness re-creates the launcher's geometry (launch_easu_h_tiled, 2x branch) and the CTA/thread structure.
#include <pthread.h>
#include <thread>
#include <vector>

#include "cuda_emu.h"

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 gridDim, blockDim;
static pthread_barrier_t g_cta_barrier;
void __syncthreads() { pthread_barrier_wait(&g_cta_barrier); }
// warp shuffles: per-warp exchange buffer between two warp-wide barriers (initialised per CTA by run_cta)
static pthread_barrier_t g_warp_barrier[32];
static uint32_t g_shfl_buf[32][32];
uint32_t fsr1_emu_shfl(uint32_t v, int src) {
  const int w = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
  g_shfl_buf[w][lane] = v;
  pthread_barrier_wait(&g_warp_barrier[w]);
  const uint32_t r = src < 0 ? v : g_shfl_buf[w][src];
  pthread_barrier_wait(&g_warp_barrier[w]);
  return r;
}
alignas(128) static unsigned char g_dynamic_smem[232448];
unsigned char* fsr1_emu_dynamic_smem() { return g_dynamic_smem; }

#include "../../fidelityfx-fsr_b200/csrc/fsr1_easu_tiled.cu"
#include "../../fidelityfx-fsr_b200/csrc/fsr1_rcas_packed.cu"
#include "../../fidelityfx-fsr_b200/csrc/fsr1_rcas_f32.cu"
#include "../../fidelityfx-fsr_b200/csrc/fsr1_fused.cu"
#include "../../fidelityfx-fsr_b200/csrc/fsr1_easu_f32.cu"
#include "../../fidelityfx-fsr_b200/csrc/fsr1_hx2.cu"

using namespace fsr1;

static int cell_of(int o, float scale, float offset) {  // = host_fp of the launcher = easu_pos on the device
  volatile float m = (float)o * scale;
  volatile float s = m + offset;
  return (int)floorf(s);
}

template <typename Kernel>
static void run_grid(Kernel kernel, int grid, int threads, const EasuParams& p, const CUtensorMap& tmap, int tiles_x, int n_tiles,
                     int mbase) {
  for (int b = 0; b < grid; b++) {
    pthread_barrier_init(&g_cta_barrier, nullptr, (unsigned)threads);
    std::vector<std::thread> ts;
    for (int t = 0; t < threads; t++)
      ts.emplace_back([=, &p, &tmap]() {
        threadIdx = uint3{(unsigned)t, 0, 0};
        blockIdx = uint3{(unsigned)b, 0, 0};
        gridDim.x = (unsigned)grid;
        blockDim.x = (unsigned)threads;
        kernel(p, tmap, tiles_x, n_tiles, mbase);
      });
    for (auto& th : ts) th.join();
    pthread_barrier_destroy(&g_cta_barrier);
  }
}

// variant: 12 = the production kernel (the number it carried among the round-1 candidates; the others were deleted).
// Images are RGBA16F, whole frames (row0 = 0).  Returns 0, or -1 for an unknown variant / not exactly 2x.
extern "C" int emu_easu_h_quad2x(int variant, const void* in, int iw, int ih, long long in_pitch, void* out, int ow, int oh,
                                 long long out_pitch, const uint32_t* con, int y0, int y1, int max_ctas) {
  EasuParams p;
  p.in = ImgView{(unsigned char*)in, in_pitch, iw, ih, 0, ih};
  p.out = ImgView{(unsigned char*)out, out_pitch, ow, oh, 0, oh};
  memcpy(&p.c0x, &con[0], 4); memcpy(&p.c0y, &con[1], 4); memcpy(&p.c0z, &con[2], 4); memcpy(&p.c0w, &con[3], 4);
  p.y0 = y0; p.y1 = y1;
  if (!(p.c0x == 0.5f && p.c0y == 0.5f && p.c0z == -0.25f && p.c0w == -0.25f)) return -1;
  constexpr int NW = 4, CY = 2 * NW;
  const int k_first = -1, k_last = cell_of(ow - 1, 0.5f, -0.25f);
  const int m_first = cell_of(y0, 0.5f, -0.25f), m_last = cell_of(y1 - 1, 0.5f, -0.25f);
  const int tiles_x = (k_last - k_first + 1 + kQCX - 1) / kQCX;
  const int tiles_y = (m_last - m_first + 1 + CY - 1) / CY, n_tiles = tiles_x * tiles_y;
  const int grid = n_tiles < max_ctas ? n_tiles : max_ctas;
  CUtensorMap tmap{(const unsigned char*)in, iw, ih, in_pitch, kQBW, CY + 3, 8};
  if (variant != 12) return -1;  // the production kernel: launch_easu_h_tiled's 2x branch (7 CTAs per SM on the GPU)
  run_grid(easu_h_quad2x_kernel<4, 7>, grid, NW * 32, p, tmap, tiles_x, n_tiles, m_first);
  return 0;
}

// launch_easu_h_tiled's generic branch: largest footprint (in texels) any tile needs along one axis
static int max_footprint(int n_out, int first, int tile, float scale, float offset, bool even_origin) {
  int best = 4;
  for (int o0 = first; o0 < n_out; o0 += tile) {
    const int o1 = (o0 + tile - 1 < n_out - 1) ? o0 + tile - 1 : n_out - 1;
    int origin = cell_of(o0, scale, offset) - 1;
    if (even_origin) origin &= ~1;
    const int span = cell_of(o1, scale, offset) + 2 - origin + 1;
    if (span > best) best = span;
  }
  return best;
}

// The any-scale kernel (easu_h_pairs_kernel: vertical pixel pairs, 64x32 tiles).  Returns 0, -1 if unsupported.
extern "C" int emu_easu_h_pairs(int variant, const void* in, int iw, int ih, long long in_pitch, void* out, int ow, int oh,
                                long long out_pitch, const uint32_t* con, int y0, int y1, int max_ctas) {
  if (variant != 1) return -1;  // the production any-scale kernel
  EasuParams p;
  p.in = ImgView{(unsigned char*)in, in_pitch, iw, ih, 0, ih};
  p.out = ImgView{(unsigned char*)out, out_pitch, ow, oh, 0, oh};
  memcpy(&p.c0x, &con[0], 4); memcpy(&p.c0y, &con[1], 4); memcpy(&p.c0z, &con[2], 4); memcpy(&p.c0w, &con[3], 4);
  p.y0 = y0; p.y1 = y1;
  if (!(p.c0x > 0.0f && p.c0x <= 1.0f && p.c0y > 0.0f && p.c0y <= 1.0f)) return -1;
  int BW = max_footprint(ow, 0, kTileW, p.c0x, p.c0z, true);
  const int BH = max_footprint(y1, y0, kTileH, p.c0y, p.c0w, false);
  BW = (BW + 1) & ~1;
  if (BW > 256 || BH > 256 || pairs_smem_bytes(BW, BH) > sizeof g_dynamic_smem) return -1;
  const int tiles_x = (ow + kTileW - 1) / kTileW, n_tiles = tiles_x * ((y1 - y0 + kTileH - 1) / kTileH);
  const int grid = n_tiles < max_ctas ? n_tiles : max_ctas;
  CUtensorMap tmap{(const unsigned char*)in, iw, ih, in_pitch, BW, BH, 8};
  for (int b = 0; b < grid; b++) {
    pthread_barrier_init(&g_cta_barrier, nullptr, (unsigned)kThreads);
    std::vector<std::thread> ts;
    for (int t = 0; t < kThreads; t++)
      ts.emplace_back([=, &p, &tmap]() {
        threadIdx = uint3{(unsigned)t, 0, 0};
        blockIdx = uint3{(unsigned)b, 0, 0};
        gridDim.x = (unsigned)grid;
        blockDim.x = (unsigned)kThreads;
        easu_h_pairs_kernel(p, tmap, BW, BH, tiles_x, n_tiles);
      });
    for (auto& th : ts) th.join();
    pthread_barrier_destroy(&g_cta_barrier);
  }
  return 0;
}

// The production RCAS kernel (rcas_h_packed_kernel<kClamp>: 4 rows per lane, 4 warps): grid = 60-pixel spans x 16-row bands.
// `in` points at logical row in_row0 and holds in_rows rows (a row-slab window; in_row0 = 0, in_rows = h for a whole image).
// opts: bit 0 FSR_RCAS_DENOISE, bit 1 FSR_RCAS_PASSTHROUGH_ALPHA, bit 2 the Sample.x output square (kRcas* of fsr1_rcas_math.cuh)
extern "C" int emu_rcas_h_packed_opt(const void* in, int in_row0, int in_rows, void* out, int w, int h, long long in_pitch,
                                     long long out_pitch, const uint32_t* con, int clamp, int y0, int y1, int opts);
extern "C" int emu_rcas_h_packed(const void* in, void* out, int w, int h, long long in_pitch, long long out_pitch,
                                 const uint32_t* con, int clamp, int y0, int y1) {
  return emu_rcas_h_packed_opt(in, 0, h, out, w, h, in_pitch, out_pitch, con, clamp, y0, y1, 0);
}
extern "C" int emu_rcas_h_packed_win(const void* in, int in_row0, int in_rows, void* out, int w, int h, long long in_pitch,
                                     long long out_pitch, const uint32_t* con, int clamp, int y0, int y1) {
  return emu_rcas_h_packed_opt(in, in_row0, in_rows, out, w, h, in_pitch, out_pitch, con, clamp, y0, y1, 0);
}
template <typename FM, bool kClamp> static void emu_rcas_dispatch(const RcasParams& p, int opts) {
  switch (opts & 7) {
    case 0: rcas_packed_kernel<FM, kClamp, 0>(p); break;
    case 1: rcas_packed_kernel<FM, kClamp, 1>(p); break;
    case 2: rcas_packed_kernel<FM, kClamp, 2>(p); break;
    case 3: rcas_packed_kernel<FM, kClamp, 3>(p); break;
    case 4: rcas_packed_kernel<FM, kClamp, 4>(p); break;
    case 5: rcas_packed_kernel<FM, kClamp, 5>(p); break;
    case 6: rcas_packed_kernel<FM, kClamp, 6>(p); break;
    default: rcas_packed_kernel<FM, kClamp, 7>(p); break;
  }
}
extern "C" int emu_rcas_h_packed_opt(const void* in, int in_row0, int in_rows, void* out, int w, int h, long long in_pitch,
                                     long long out_pitch, const uint32_t* con, int clamp, int y0, int y1, int opts) {
  RcasParams p;
  p.in = ImgView{(unsigned char*)in, in_pitch, w, h, in_row0, in_rows};
  p.out = ImgView{(unsigned char*)out, out_pitch, w, h, 0, h};
  memcpy(&p.sharp, &con[0], 4);
  p.sharp_h2 = con[1];
  p.y0 = y0; p.y1 = y1; p.clamp = clamp; p.options = 0;
  constexpr int NWARP = 4, ROWS = 4, threads = 32 * NWARP;
  const int gx = (w + kSpan - 1) / kSpan, gy = (y1 - y0 + NWARP * ROWS - 1) / (NWARP * ROWS);
  for (int by = 0; by < gy; by++)
    for (int bx = 0; bx < gx; bx++) {
      for (int i = 0; i < NWARP; i++) pthread_barrier_init(&g_warp_barrier[i], nullptr, 32);
      std::vector<std::thread> ts;
      for (int t = 0; t < threads; t++)
        ts.emplace_back([=, &p]() {
          threadIdx = uint3{(unsigned)t, 0, 0};
          blockIdx = uint3{(unsigned)bx, (unsigned)by, 0};
          gridDim.x = (unsigned)gx; gridDim.y = (unsigned)gy;
          blockDim.x = (unsigned)threads;
          if (clamp) emu_rcas_dispatch<FmtHalf, true>(p, opts);
          else emu_rcas_dispatch<FmtHalf, false>(p, opts);
        });
      for (auto& th : ts) th.join();
      for (int i = 0; i < NWARP; i++) pthread_barrier_destroy(&g_warp_barrier[i]);
    }
  return 0;
}

// easu_u_quad2x_kernel: 2x EASU on R8G8B8A8_UNORM (bits = 8) / R10G10B10A2_UNORM (bits = 10) images, 4 bytes per texel.
extern "C" int emu_easu_u_quad2x(int bits, const void* in, int iw, int ih, long long in_pitch, void* out, int ow, int oh,
                                 long long out_pitch, const uint32_t* con, int y0, int y1, int max_ctas) {
  EasuParams p;
  p.in = ImgView{(unsigned char*)in, in_pitch, iw, ih, 0, ih};
  p.out = ImgView{(unsigned char*)out, out_pitch, ow, oh, 0, oh};
  memcpy(&p.c0x, &con[0], 4); memcpy(&p.c0y, &con[1], 4); memcpy(&p.c0z, &con[2], 4); memcpy(&p.c0w, &con[3], 4);
  p.y0 = y0; p.y1 = y1;
  if (!(p.c0x == 0.5f && p.c0y == 0.5f && p.c0z == -0.25f && p.c0w == -0.25f) || (bits != 8 && bits != 10)) return -1;
  constexpr int NW = 4, CY = 2 * NW;
  const int k_first = -1, k_last = cell_of(ow - 1, 0.5f, -0.25f);
  const int m_first = cell_of(y0, 0.5f, -0.25f), m_last = cell_of(y1 - 1, 0.5f, -0.25f);
  const int tiles_x = (k_last - k_first + 1 + kQCX - 1) / kQCX;
  const int tiles_y = (m_last - m_first + 1 + CY - 1) / CY, n_tiles = tiles_x * tiles_y;
  const int grid = n_tiles < max_ctas ? n_tiles : max_ctas;
  CUtensorMap tmap{(const unsigned char*)in, iw, ih, in_pitch, kUBW, CY + 3, 4};
  if (bits == 8) run_grid(easu_u_quad2x_kernel<4, 6, 8>, grid, NW * 32, p, tmap, tiles_x, n_tiles, m_first);
  else run_grid(easu_u_quad2x_kernel<4, 6, 10>, grid, NW * 32, p, tmap, tiles_x, n_tiles, m_first);
  return 0;
}

// rcas_u_packed_kernel: RCAS on UNORM images (bits = 8 or 10), 4 bytes per texel.
extern "C" int emu_rcas_u_packed_opt(int bits, const void* in, void* out, int w, int h, long long in_pitch, long long out_pitch,
                                     const uint32_t* con, int clamp, int y0, int y1, int opts);
extern "C" int emu_rcas_u_packed(int bits, const void* in, void* out, int w, int h, long long in_pitch, long long out_pitch,
                                 const uint32_t* con, int clamp, int y0, int y1) {
  return emu_rcas_u_packed_opt(bits, in, out, w, h, in_pitch, out_pitch, con, clamp, y0, y1, 0);
}
extern "C" int emu_rcas_u_packed_opt(int bits, const void* in, void* out, int w, int h, long long in_pitch, long long out_pitch,
                                     const uint32_t* con, int clamp, int y0, int y1, int opts) {
  if (bits != 8 && bits != 10) return -1;
  RcasParams p;
  p.in = ImgView{(unsigned char*)in, in_pitch, w, h, 0, h};
  p.out = ImgView{(unsigned char*)out, out_pitch, w, h, 0, h};
  memcpy(&p.sharp, &con[0], 4);
  p.sharp_h2 = con[1];
  p.y0 = y0; p.y1 = y1; p.clamp = clamp; p.options = 0;
  constexpr int NWARP = 4, threads = 32 * NWARP;
  const int gx = (w + kSpan - 1) / kSpan, gy = (y1 - y0 + 15) / 16;
  for (int by = 0; by < gy; by++)
    for (int bx = 0; bx < gx; bx++) {
      for (int i = 0; i < NWARP; i++) pthread_barrier_init(&g_warp_barrier[i], nullptr, 32);
      std::vector<std::thread> ts;
      for (int t = 0; t < threads; t++)
        ts.emplace_back([=, &p]() {
          threadIdx = uint3{(unsigned)t, 0, 0};
          blockIdx = uint3{(unsigned)bx, (unsigned)by, 0};
          gridDim.x = (unsigned)gx; gridDim.y = (unsigned)gy;
          blockDim.x = (unsigned)threads;
          if (bits == 8) { if (clamp) emu_rcas_dispatch<FmtUnorm<8>, true>(p, opts); else emu_rcas_dispatch<FmtUnorm<8>, false>(p, opts); }
          else { if (clamp) emu_rcas_dispatch<FmtUnorm<10>, true>(p, opts); else emu_rcas_dispatch<FmtUnorm<10>, false>(p, opts); }
        });
      for (auto& th : ts) th.join();
      for (int i = 0; i < NWARP; i++) pthread_barrier_destroy(&g_warp_barrier[i]);
    }
  return 0;
}

// rcas_f32_packed_kernel: RCAS on RGBA32F images (MUFU reciprocals).
extern "C" int emu_rcas_f32_packed(int variant, const void* in, void* out, int w, int h, long long in_pitch, long long out_pitch,
                                   const uint32_t* con, int clamp, int y0, int y1) {
  RcasParams p;
  p.in = ImgView{(unsigned char*)in, in_pitch, w, h, 0, h};
  p.out = ImgView{(unsigned char*)out, out_pitch, w, h, 0, h};
  memcpy(&p.sharp, &con[0], 4);
  p.sharp_h2 = con[1];
  p.y0 = y0; p.y1 = y1; p.clamp = clamp; p.options = 0;
  const int threads = 32 * kFWarps;
  const int gx = (w + kFSpan - 1) / kFSpan, gy = (y1 - y0 + kFWarps * kFRows - 1) / (kFWarps * kFRows);
  for (int by = 0; by < gy; by++)
    for (int bx = 0; bx < gx; bx++) {
      for (int i = 0; i < kFWarps; i++) pthread_barrier_init(&g_warp_barrier[i], nullptr, 32);
      std::vector<std::thread> ts;
      for (int t = 0; t < threads; t++)
        ts.emplace_back([=, &p]() {
          threadIdx = uint3{(unsigned)t, 0, 0};
          blockIdx = uint3{(unsigned)bx, (unsigned)by, 0};
          gridDim.x = (unsigned)gx; gridDim.y = (unsigned)gy;
          blockDim.x = (unsigned)threads;
          rcas_f32_packed_kernel<0>(p);
        });
      for (auto& th : ts) th.join();
      for (int i = 0; i < kFWarps; i++) pthread_barrier_destroy(&g_warp_barrier[i]);
    }
  return 0;
}


// fused_h_quad2x_kernel: EASU -> RCAS in one kernel (RGBA16F, 2x), launch geometry of launch_fused_h with `ctas` CTAs.
extern "C" int emu_fused_h(const void* in, int iw, int ih, long long in_pitch, void* out, int ow, int oh, long long out_pitch,
                           const uint32_t* rcon, int y0, int y1, int ctas) {
  constexpr int NW = 4;
  using C = FusedCfg<NW>;
  FusedParams p;
  p.in = ImgView{(unsigned char*)in, in_pitch, iw, ih, 0, ih};
  p.out = ImgView{(unsigned char*)out, out_pitch, ow, oh, 0, oh};
  p.y0 = y0; p.y1 = y1; p.sharp_h2 = rcon[1];
  p.n_strips = ((ow + 1) / 2 + kStripCells - 1) / kStripCells;
  CUtensorMap tmap{(const unsigned char*)in, iw, ih, in_pitch, kFBW, C::kBH, 8};
  const int threads = NW * 32;
  for (int b = 0; b < ctas; b++) {
    pthread_barrier_init(&g_cta_barrier, nullptr, (unsigned)threads);
    std::vector<std::thread> ts;
    for (int t = 0; t < threads; t++)
      ts.emplace_back([=, &p, &tmap]() {
        threadIdx = uint3{(unsigned)t, 0, 0};
        blockIdx = uint3{(unsigned)b, 0, 0};
        gridDim.x = (unsigned)ctas;
        blockDim.x = (unsigned)threads;
        fused_h_quad2x_kernel<NW, 6>(p, tmap);
      });
    for (auto& th : ts) th.join();
    pthread_barrier_destroy(&g_cta_barrier);
  }
  return 0;
}

// easu_f32_pairs_kernel<S>: any-scale EASU with fp32 arithmetic; storage 0 = RGBA32F (16-byte texels), 1 = RGBA16F.
template <typename S>
static int run_f32_pairs(const void* in, int iw, int ih, long long in_pitch, void* out, int ow, int oh, long long out_pitch,
                         const uint32_t* con, int y0, int y1, int max_ctas) {
  constexpr int kB = Tex<S>::kBytes;
  EasuParams p;
  p.in = ImgView{(unsigned char*)in, in_pitch, iw, ih, 0, ih};
  p.out = ImgView{(unsigned char*)out, out_pitch, ow, oh, 0, oh};
  memcpy(&p.c0x, &con[0], 4); memcpy(&p.c0y, &con[1], 4); memcpy(&p.c0z, &con[2], 4); memcpy(&p.c0w, &con[3], 4);
  p.y0 = y0; p.y1 = y1;
  if (!(p.c0x > 0.0f && p.c0x <= 1.0f && p.c0y > 0.0f && p.c0y <= 1.0f)) return -1;
  int BW = max_footprint(ow, 0, kFTileW, p.c0x, p.c0z, kB == 8);
  const int BH = max_footprint(y1, y0, kFTileH, p.c0y, p.c0w, false);
  if (kB == 8) BW = (BW + 1) & ~1;
  if (BW > 256 || BH > 256 || fpairs_smem_bytes<S>(BW, BH) > sizeof g_dynamic_smem) return -1;
  const int tiles_x = (ow + kFTileW - 1) / kFTileW, n_tiles = tiles_x * ((y1 - y0 + kFTileH - 1) / kFTileH);
  const int grid = n_tiles < max_ctas ? n_tiles : max_ctas;
  CUtensorMap tmap{(const unsigned char*)in, iw, ih, in_pitch, BW, BH, kB};
  for (int b = 0; b < grid; b++) {
    pthread_barrier_init(&g_cta_barrier, nullptr, (unsigned)kFThreads);
    std::vector<std::thread> ts;
    for (int t = 0; t < kFThreads; t++)
      ts.emplace_back([=, &p, &tmap]() {
        threadIdx = uint3{(unsigned)t, 0, 0};
        blockIdx = uint3{(unsigned)b, 0, 0};
        gridDim.x = (unsigned)grid;
        blockDim.x = (unsigned)kFThreads;
        easu_f32_pairs_kernel<S>(p, tmap, BW, BH, tiles_x, n_tiles);
      });
    for (auto& th : ts) th.join();
    pthread_barrier_destroy(&g_cta_barrier);
  }
  return 0;
}
extern "C" int emu_easu_f32_pairs(int half_storage, const void* in, int iw, int ih, long long in_pitch, void* out, int ow, int oh,
                                  long long out_pitch, const uint32_t* con, int y0, int y1, int max_ctas) {
  return half_storage ? run_f32_pairs<__half>(in, iw, ih, in_pitch, out, ow, oh, out_pitch, con, y0, y1, max_ctas)
                      : run_f32_pairs<float>(in, iw, ih, in_pitch, out, ow, oh, out_pitch, con, y0, y1, max_ctas);
}

// ---- the packed Hx2 calling convention (csrc/fsr1_hx2.cu): RCAS and the pointwise companions, RGBA16F ------------------------
template <typename Body> static void run_hx2_grid(int w, int rows, Body body) {
  const int gx = (w + kHx2Span - 1) / kHx2Span;
  for (int by = 0; by < rows; by++)
    for (int bx = 0; bx < gx; bx++) {
      std::vector<std::thread> ts;
      for (int t = 0; t < kHx2Threads; t++)
        ts.emplace_back([=]() {
          threadIdx = uint3{(unsigned)t, 0, 0};
          blockIdx = uint3{(unsigned)bx, (unsigned)by, 0};
          gridDim.x = (unsigned)gx; gridDim.y = (unsigned)rows;
          blockDim.x = (unsigned)kHx2Threads;
          body();
        });
      for (auto& th : ts) th.join();
    }
}

// `in` points at logical row in_row0 and holds in_rows rows (a row-slab window); opts: bit 0 DENOISE, bit 1 PASSTHROUGH_ALPHA
extern "C" int emu_rcas_hx2(const void* in, int in_row0, int in_rows, void* out, int w, int h, long long in_pitch, long long out_pitch,
                            const uint32_t* con, int clamp, int y0, int y1, int opts) {
  RcasParams p;
  p.in = ImgView{(unsigned char*)in, in_pitch, w, h, in_row0, in_rows};
  p.out = ImgView{(unsigned char*)out, out_pitch, w, h, 0, h};
  memcpy(&p.sharp, &con[0], 4);
  p.sharp_h2 = con[1];
  p.y0 = y0; p.y1 = y1; p.clamp = clamp; p.options = opts & 3;
  run_hx2_grid(w, y1 - y0, [&p]() { rcas_hx2_kernel(p); });
  return 0;
}

// op: 1 SRTM, 2 SRTM inverse, 3 LFGA, 4 TEPD 8 bit, 5 TEPD 10 bit; aux = grain / dither tile (aw x ah) or null
extern "C" int emu_pointwise_hx2(int op, const void* in, long long in_pitch, void* out, long long out_pitch, int w, int h, const void* aux,
                                 int aw, int ah, long long aux_pitch, float amount, uint32_t frame, int y0, int y1) {
  if (op < 1 || op > 5 || (op == 3 && !aux)) return -1;
  PointHParams p;
  p.in = ImgView{(unsigned char*)in, in_pitch, w, h, 0, h};
  p.out = ImgView{(unsigned char*)out, out_pitch, w, h, 0, h};
  p.has_aux = aux ? 1 : 0;
  p.aux = aux ? ImgView{(unsigned char*)aux, aux_pitch, aw, ah, 0, ah} : p.in;
  p.op = op; p.amount = amount; p.frame = frame; p.y0 = y0; p.y1 = y1;
  run_hx2_grid(w, y1 - y0, [&p]() { pointwise_hx2_kernel(p); });
  return 0;
}
