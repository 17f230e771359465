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
alignas(128) static unsigned char g_dynamic_smem[232448];
unsigned char* fsr1_emu_dynamic_smem() { return g_dynamic_smem; }

#include "../../fidelityfx-fsr_b200/csrc/fsr1_easu_tiled.cu"

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

// variant: the FSR1_EASU_QUAD_VARIANT numbering of launch_easu_h_tiled (2 plain, 6 default, 7, 8, 9).
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
  switch (variant) {
    case 2: run_grid(easu_h_quad2x_kernel<4, 6, 0, false>, grid, NW * 32, p, tmap, tiles_x, n_tiles, m_first); break;
    case 6: run_grid(easu_h_quad2x_kernel<4, 6, 1, false>, grid, NW * 32, p, tmap, tiles_x, n_tiles, m_first); break;
    case 7: run_grid(easu_h_quad2x_kernel<4, 6, 2, false>, grid, NW * 32, p, tmap, tiles_x, n_tiles, m_first); break;
    case 8: run_grid(easu_h_quad2x_kernel<4, 6, 3, false>, grid, NW * 32, p, tmap, tiles_x, n_tiles, m_first); break;
    case 9: run_grid(easu_h_quad2x_kernel<4, 6, 3, true>, grid, NW * 32, p, tmap, tiles_x, n_tiles, m_first); break;
    default: return -1;
  }
  return 0;
}
