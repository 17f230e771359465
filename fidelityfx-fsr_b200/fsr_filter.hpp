// fsr_filter.hpp — C++ twin of the reference's FSR_Filter (sample/src/DX12/FSR_Filter.h:27-45,
// FSR_Filter.cpp:41-141) over the C ABI of include/fsr1_b200.h.  Header-only; link libfsr1_b200.so.
// Same method names and call pattern; D3D12 resources become device pointers, the command list a CUDA stream.
#pragma once
#include <stdexcept>
#include <string>

#include "../include/fsr1_b200.h"
#include "../include/fsr1_host.h"

namespace fsr1 {

// The fields of the sample's State that Upscale reads (sample/src/DX12/SampleRenderer.h).
struct State {
  int renderWidth = 0, renderHeight = 0;
  bool bUseRcas = true;
  float rcasAttenuation = 0.25f;
  int m_nUpscaleType = 1;  // 1 = FSR 1.0; 0 (bilinear comparison mode) is out of scope
};

// A device image view handed to Upscale (replaces the SRV/UAV pair of the sample).
struct Texture {
  void* data = nullptr;
  uint64_t pitchBytes = 0;
};

class FSR_Filter {
 public:
  // FSR_Filter::OnCreate: choose the fp16 kernels or the fp32 "slow fallback" (FSR_Filter.cpp:41-68).
  void OnCreate(bool slowFallback = false) { m_format = slowFallback ? FSR1_FORMAT_RGBA32F : FSR1_FORMAT_RGBA16F; }

  // FSR_Filter::OnCreateWindowSizeDependentResources: the display-sized intermediate (FSR_Filter.cpp:70-90).
  void OnCreateWindowSizeDependentResources(uint32_t renderWidth, uint32_t renderHeight, uint32_t displayWidth,
                                            uint32_t displayHeight) {
    OnDestroyWindowSizeDependentResources();
    check(fsr1_context_create(&m_ctx, renderWidth, renderHeight, displayWidth, displayHeight, m_format));
    m_displayWidth = displayWidth;
    m_displayHeight = displayHeight;
  }
  void OnDestroyWindowSizeDependentResources() {
    if (m_ctx) fsr1_context_destroy(m_ctx);
    m_ctx = nullptr;
  }
  void OnDestroy() { OnDestroyWindowSizeDependentResources(); }
  ~FSR_Filter() { OnDestroy(); }

  // FSR_Filter::Upscale (FSR_Filter.cpp:101-141): FsrEasuCon, EASU dispatch, FsrRcasCon, RCAS dispatch.
  void Upscale(void* stream, int displayWidth, int displayHeight, const State* pState, const Texture& input,
               const Texture& output, bool hdr = false) {
    if (!m_ctx || (uint32_t)displayWidth != m_displayWidth || (uint32_t)displayHeight != m_displayHeight)
      throw std::runtime_error("FSR_Filter: OnCreateWindowSizeDependentResources not called for this display size");
    if (pState->m_nUpscaleType != 1)
      throw std::runtime_error("FSR_Filter: only the FSR 1.0 path is implemented (no bilinear comparison mode)");
    // hdr = the sample's Sample.x (FSR_Filter.cpp:107,125): the last pass squares its output
    const uint32_t flags = (pState->bUseRcas ? 0u : FSR1_FLAG_NO_RCAS) | (hdr ? FSR1_FLAG_OUTPUT_SQUARE : 0u);
    // the constants are rebuilt from THIS frame's render size, as the reference does on every call (FSR_Filter.cpp:106)
    if (pState->renderWidth <= 0 || pState->renderHeight <= 0) throw std::runtime_error("FSR_Filter: pState->renderWidth/renderHeight not set");
    check(fsr1_context_upscale_render(m_ctx, input.data, input.pitchBytes, (uint32_t)pState->renderWidth, (uint32_t)pState->renderHeight,
                                      output.data, output.pitchBytes, pState->rcasAttenuation, flags, stream));
  }

 private:
  static void check(int rc) {
    if (rc != FSR1_OK)
      throw std::runtime_error(std::string("fsr1: ") + fsr1_error_string(rc) + " (cuda error " +
                               std::to_string(fsr1_last_cuda_error()) + ")");
  }
  fsr1_context* m_ctx = nullptr;
  uint32_t m_format = FSR1_FORMAT_RGBA16F, m_displayWidth = 0, m_displayHeight = 0;
};

}  // namespace fsr1
