// fsr1_capi.cu — the C ABI declared in include/fsr1_b200.h: argument validation, kernel selection,
// the resource-owning context, and the constant-setup entry points.
#include <atomic>
#include <math.h>
#include <new>
#include <string.h>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX 3: ranges "EASU" / "RCAS" / "FSR1" around the launches, as the sample's
                               // user markers do (sample/src/DX12/FSR_Filter.cpp:118,128); no-ops unless a profiler is attached

#include "../../include/fsr1_b200.h"
#include "../../include/fsr1_host.h"
#include "fsr1_common.cuh"

using namespace fsr1;

namespace {

struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

thread_local int t_last_cuda = 0;
thread_local const char* t_last_kernel = "";
std::atomic<unsigned long long> g_launches{0};

}  // namespace
namespace fsr1 {
void set_last_detail(int v) { t_last_cuda = v; }  // fsr1_shard_status: which wait timed out, reported through fsr1_last_cuda_error()
static thread_local const HaloSync* t_sync = nullptr;
static thread_local bool t_sync_used = false;
void set_halo_sync(const HaloSync* hs) { t_sync = hs; t_sync_used = false; }
bool halo_sync_consumed() { return t_sync_used; }
}
namespace {

int cuda_fail(cudaError_t e) {
  t_last_cuda = (int)e;
  return FSR1_ERR_CUDA;
}

int bytes_per_pixel(uint32_t fmt) {
  switch (fmt) {
    case FSR1_FORMAT_RGBA16F: return 8;
    case FSR1_FORMAT_RGBA32F: return 16;
    case FSR1_FORMAT_RGBA8_UNORM: case FSR1_FORMAT_RGB10A2_UNORM: return 4;
    default: return 0;
  }
}

int check_image(const fsr1_image* im) {
  if (!im || !im->data || im->width == 0 || im->height == 0 || im->rows == 0) return FSR1_ERR_INVALID_ARGUMENT;
  const int bpp = bytes_per_pixel(im->format);
  if (!bpp) return FSR1_ERR_INVALID_ARGUMENT;
  if (im->width > 32768u || im->height > 32768u) return FSR1_ERR_INVALID_ARGUMENT;
  if (im->pitch_bytes < (uint64_t)im->width * bpp || (im->pitch_bytes % bpp) != 0) return FSR1_ERR_INVALID_ARGUMENT;
  if ((uintptr_t)im->data % bpp != 0) return FSR1_ERR_INVALID_ARGUMENT;
  if ((uint64_t)im->row0 + im->rows > im->height) return FSR1_ERR_INVALID_ARGUMENT;
  return FSR1_OK;
}

ImgView view_of(const fsr1_image* im) {
  ImgView v;
  v.base = static_cast<unsigned char*>(im->data);
  v.pitch = (long long)im->pitch_bytes;
  v.w = (int)im->width;
  v.h = (int)im->height;
  v.row0 = (int)im->row0;
  v.rows = (int)im->rows;
  return v;
}

// identical float arithmetic to easu_pos() on the device and to the oracle
int host_cell(uint32_t o, float scale, float offset) {
  volatile float m = (float)o * scale;
  volatile float s = m + offset;
  return (int)floorf(s);
}

float as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

// every flag include/fsr1_b200.h defines
constexpr uint32_t kAllFlags = FSR1_FLAG_RCAS_CLAMP | FSR1_FLAG_EXACT | FSR1_FLAG_FORCE_DIRECT | FSR1_FLAG_NO_RCAS |
                               FSR1_FLAG_H_REFERENCE | FSR1_FLAG_PRECISE | FSR1_FLAG_RCAS_DENOISE |
                               FSR1_FLAG_RCAS_PASSTHROUGH_ALPHA | FSR1_FLAG_OUTPUT_SQUARE | FSR1_FLAG_FUSED | FSR1_FLAG_RCAS_HX2;

bool window_holds(const fsr1_image* im, int first, int last) {  // logical rows [first,last]
  return first >= (int)im->row0 && last < (int)(im->row0 + im->rows);
}

// the Sample.x hook: `c *= c` in place on the rows the last pass wrote (a separate streaming pass)
int square_rows(const fsr1_image* img, uint32_t y0, uint32_t y1, cudaStream_t s) {
  const ImgView v = view_of(img);
  const char* name = "";
  cudaError_t e = launch_pointwise(6, v, (int)img->format, v, (int)img->format, nullptr, 0, 0.0f, 0u, (int)y0, (int)y1, s, &name);
  if (e != cudaSuccess) return cuda_fail(e);
  t_last_kernel = name;
  g_launches.fetch_add(1);
  return FSR1_OK;
}

int pointwise(int op, const fsr1_image* in, const fsr1_image* aux, const fsr1_image* out, float amount, uint32_t frame,
              uint32_t y0, uint32_t y1, void* stream) {
  int rc;
  if ((rc = check_image(in)) != FSR1_OK || (rc = check_image(out)) != FSR1_OK) return rc;
  if (aux && (rc = check_image(aux)) != FSR1_OK) return rc;
  if (aux && (aux->row0 != 0 || aux->rows != aux->height)) return FSR1_ERR_INVALID_ARGUMENT;  // tiles are whole images
  if (in->width != out->width || in->height != out->height) return FSR1_ERR_INVALID_ARGUMENT;
  if (y1 == 0) y1 = out->height;
  if (y0 >= y1 || y1 > out->height) return FSR1_ERR_INVALID_ARGUMENT;
  if (!window_holds(out, (int)y0, (int)y1 - 1) || !window_holds(in, (int)y0, (int)y1 - 1)) return FSR1_ERR_WINDOW;
  const ImgView vi = view_of(in), vo = view_of(out);
  ImgView va;
  if (aux) va = view_of(aux);
  const char* name = "";
  cudaError_t e = launch_pointwise(op, vi, (int)in->format, vo, (int)out->format, aux ? &va : nullptr, aux ? (int)aux->format : 0,
                                   amount, frame, (int)y0, (int)y1, static_cast<cudaStream_t>(stream), &name);
  if (e == cudaErrorNotSupported) return FSR1_ERR_UNSUPPORTED;
  if (e != cudaSuccess) return cuda_fail(e);
  t_last_kernel = name;
  g_launches.fetch_add(1);
  return FSR1_OK;
}

// the half-precision (H / Hx2) forms: RGBA16F everywhere
int pointwise_h(int op, const fsr1_image* in, const fsr1_image* aux, const fsr1_image* out, float amount, uint32_t frame, uint32_t y0,
                uint32_t y1, void* stream) {
  int rc;
  if ((rc = check_image(in)) != FSR1_OK || (rc = check_image(out)) != FSR1_OK) return rc;
  if (aux && (rc = check_image(aux)) != FSR1_OK) return rc;
  if (aux && (aux->row0 != 0 || aux->rows != aux->height)) return FSR1_ERR_INVALID_ARGUMENT;  // tiles are whole images
  if (in->width != out->width || in->height != out->height) return FSR1_ERR_INVALID_ARGUMENT;
  if (in->format != FSR1_FORMAT_RGBA16F || out->format != FSR1_FORMAT_RGBA16F || (aux && aux->format != FSR1_FORMAT_RGBA16F))
    return FSR1_ERR_UNSUPPORTED;
  if (y1 == 0) y1 = out->height;
  if (y0 >= y1 || y1 > out->height) return FSR1_ERR_INVALID_ARGUMENT;
  if (!window_holds(out, (int)y0, (int)y1 - 1) || !window_holds(in, (int)y0, (int)y1 - 1)) return FSR1_ERR_WINDOW;
  const ImgView vi = view_of(in), vo = view_of(out);
  ImgView va;
  if (aux) va = view_of(aux);
  const char* name = "";
  cudaError_t e = launch_pointwise_hx2(op, vi, vo, aux ? &va : nullptr, amount, frame, (int)y0, (int)y1, static_cast<cudaStream_t>(stream), &name);
  if (e != cudaSuccess) return cuda_fail(e);
  t_last_kernel = name;
  g_launches.fetch_add(1);
  return FSR1_OK;
}

}  // namespace

extern "C" {

int fsr1_abi_version(void) { return FSR1_ABI_VERSION; }

const char* fsr1_error_string(int err) {
  switch (err) {
    case FSR1_OK: return "ok";
    case FSR1_ERR_INVALID_ARGUMENT: return "invalid argument";
    case FSR1_ERR_UNSUPPORTED: return "unsupported format combination";
    case FSR1_ERR_WINDOW: return "image window does not hold the rows this pass touches";
    case FSR1_ERR_CUDA: return "CUDA error (see fsr1_last_cuda_error)";
    case FSR1_ERR_NO_DEVICE: return "no usable CUDA device";
    case FSR1_ERR_TIMEOUT: return "a neighbouring rank's halo rows or credit did not arrive in time";
    default: return "unknown fsr1 error";
  }
}
int fsr1_last_cuda_error(void) { return t_last_cuda; }
uint64_t fsr1_launch_count(void) { return g_launches.load(); }
const char* fsr1_last_kernel_name(void) { return t_last_kernel; }

void fsr1_easu_con(uint32_t con[16], float vw, float vh, float sw, float sh, float ow, float oh) {
  FsrEasuCon(con, con + 4, con + 8, con + 12, vw, vh, sw, sh, ow, oh);
}
void fsr1_easu_con_offset(uint32_t con[16], float vw, float vh, float sw, float sh, float ow, float oh, float ox,
                          float oy) {
  FsrEasuConOffset(con, con + 4, con + 8, con + 12, vw, vh, sw, sh, ow, oh, ox, oy);
}
void fsr1_rcas_con(uint32_t con[4], float sharpness_stops) { FsrRcasCon(con, sharpness_stops); }

int fsr1_easu_input_rows(const uint32_t con[16], uint32_t in_height, uint32_t y0, uint32_t y1, uint32_t* first_row,
                         uint32_t* last_row) {
  if (!con || !first_row || !last_row || in_height == 0 || y1 <= y0) return FSR1_ERR_INVALID_ARGUMENT;
  const float sy = as_float(con[1]), oy = as_float(con[3]);
  int lo = host_cell(y0, sy, oy) - 1, hi = host_cell(y1 - 1, sy, oy) + 2;
  const int H = (int)in_height;
  lo = lo < 0 ? 0 : (lo > H - 1 ? H - 1 : lo);
  hi = hi < 0 ? 0 : (hi > H - 1 ? H - 1 : hi);
  *first_row = (uint32_t)lo;
  *last_row = (uint32_t)hi;
  return FSR1_OK;
}

int fsr1_easu(const fsr1_image* in, const fsr1_image* out, const uint32_t con[16], uint32_t y0, uint32_t y1,
              uint32_t flags, void* stream) {
  NvtxRange range("EASU");
  int rc;
  if ((rc = check_image(in)) != FSR1_OK || (rc = check_image(out)) != FSR1_OK) return rc;
  if (!con || (flags & ~kAllFlags)) return FSR1_ERR_INVALID_ARGUMENT;
  if (in->format != out->format) return FSR1_ERR_UNSUPPORTED;
  if (y1 == 0) y1 = out->height;
  if (y0 >= y1 || y1 > out->height) return FSR1_ERR_INVALID_ARGUMENT;
  if (!window_holds(out, (int)y0, (int)y1 - 1)) return FSR1_ERR_WINDOW;
  uint32_t r0, r1;
  fsr1_easu_input_rows(con, in->height, y0, y1, &r0, &r1);
  if (!window_holds(in, (int)r0, (int)r1)) return FSR1_ERR_WINDOW;

  EasuParams p;
  p.in = view_of(in);
  p.out = view_of(out);
  p.c0x = as_float(con[0]); p.c0y = as_float(con[1]); p.c0z = as_float(con[2]); p.c0w = as_float(con[3]);
  p.y0 = (int)y0; p.y1 = (int)y1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool exact = (flags & FSR1_FLAG_EXACT) != 0;
  cudaError_t e = cudaErrorNotSupported;
  const char* name = "";
  if (flags & FSR1_FLAG_H_REFERENCE) {
    if (in->format != FSR1_FORMAT_RGBA16F || exact) return FSR1_ERR_UNSUPPORTED;
    e = launch_easu_href(p, s, &name);
  } else if (in->format == FSR1_FORMAT_RGBA16F && !exact && !(flags & FSR1_FLAG_FORCE_DIRECT)) {
    if (flags & FSR1_FLAG_PRECISE) e = launch_easu_h_precise(p, s, &name);
    if (e == cudaErrorNotSupported) {
      if (t_sync) p.sync = *t_sync;  // sharded frame: the neighbour hand-shake rides inside the kernel
      e = launch_easu_h_tiled(p, s, &name);
      if (e == cudaSuccess && t_sync) t_sync_used = true;
      p.sync = HaloSync{};
    }
  } else if (in->format == FSR1_FORMAT_RGBA32F && !exact && !(flags & FSR1_FLAG_FORCE_DIRECT)) {
    e = launch_easu_f32_tiled(p, s, &name);
  } else if ((in->format == FSR1_FORMAT_RGBA8_UNORM || in->format == FSR1_FORMAT_RGB10A2_UNORM) && !exact &&
             !(flags & (FSR1_FLAG_FORCE_DIRECT | FSR1_FLAG_PRECISE))) {
    e = launch_easu_u_tiled(p, (int)in->format, s, &name);  // half2 taps; FSR1_FLAG_PRECISE keeps the fp32 direct kernel
  }
  if (e == cudaErrorNotSupported) e = launch_easu_direct(p, (int)in->format, exact, s, &name);
  if (e != cudaSuccess) return cuda_fail(e);
  t_last_kernel = name;
  g_launches.fetch_add(1);
  if (flags & FSR1_FLAG_OUTPUT_SQUARE) return square_rows(out, y0, y1, s);
  return FSR1_OK;
}

int fsr1_rcas(const fsr1_image* in, const fsr1_image* out, const uint32_t con[4], uint32_t y0, uint32_t y1,
              uint32_t flags, void* stream) {
  NvtxRange range("RCAS");
  int rc;
  if ((rc = check_image(in)) != FSR1_OK || (rc = check_image(out)) != FSR1_OK) return rc;
  if (!con || (flags & ~kAllFlags)) return FSR1_ERR_INVALID_ARGUMENT;
  if (in->format != out->format) return FSR1_ERR_UNSUPPORTED;
  if (in->width != out->width || in->height != out->height) return FSR1_ERR_INVALID_ARGUMENT;
  if (y1 == 0) y1 = out->height;
  if (y0 >= y1 || y1 > out->height) return FSR1_ERR_INVALID_ARGUMENT;
  if (!window_holds(out, (int)y0, (int)y1 - 1)) return FSR1_ERR_WINDOW;
  const int need0 = y0 == 0 ? 0 : (int)y0 - 1, need1 = y1 >= out->height ? (int)out->height - 1 : (int)y1;
  if (!window_holds(in, need0, need1)) return FSR1_ERR_WINDOW;
  {  // RCAS reads neighbours of every pixel it writes: in place is a race
    const uintptr_t a0 = (uintptr_t)in->data, a1 = a0 + (uintptr_t)in->pitch_bytes * in->rows;
    const uintptr_t b0 = (uintptr_t)out->data, b1 = b0 + (uintptr_t)out->pitch_bytes * out->rows;
    if (a0 < b1 && b0 < a1) return FSR1_ERR_INVALID_ARGUMENT;
  }

  RcasParams p;
  p.in = view_of(in);
  p.out = view_of(out);
  p.sharp = as_float(con[0]);
  p.sharp_h2 = con[1];
  p.y0 = (int)y0; p.y1 = (int)y1;
  p.clamp = (flags & FSR1_FLAG_RCAS_CLAMP) ? 1 : 0;
  p.options = ((flags & FSR1_FLAG_RCAS_DENOISE) ? 1 : 0) | ((flags & FSR1_FLAG_RCAS_PASSTHROUGH_ALPHA) ? 2 : 0);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool exact = (flags & FSR1_FLAG_EXACT) != 0;
  cudaError_t e = cudaErrorNotSupported;
  const char* name = "";
  bool squared = false;  // the Sample.x hook folded into the kernel's store (c *= c before the one rounding)
  const int fused_square = (flags & FSR1_FLAG_OUTPUT_SQUARE) ? 4 : 0;
  if (flags & FSR1_FLAG_RCAS_HX2) {  // the packed calling convention: FsrRcasHx2 + FsrRcasDepackHx2
    if (in->format != FSR1_FORMAT_RGBA16F || exact) return FSR1_ERR_UNSUPPORTED;
    e = launch_rcas_hx2(p, s, &name);
  } else if (flags & FSR1_FLAG_H_REFERENCE) {
    if (in->format != FSR1_FORMAT_RGBA16F || exact) return FSR1_ERR_UNSUPPORTED;
    e = launch_rcas_href(p, s, &name);
  } else if (in->format == FSR1_FORMAT_RGBA16F && !exact && !(flags & FSR1_FLAG_FORCE_DIRECT)) {
    p.options |= fused_square;  // the reference's options are template bits of the packed kernels: no slower fallback
    e = launch_rcas_h_packed(p, s, &name);
    squared = e == cudaSuccess;
  } else if (in->format == FSR1_FORMAT_RGBA32F && !exact && !(flags & FSR1_FLAG_FORCE_DIRECT)) {
    p.options |= fused_square;
    e = launch_rcas_f32_packed(p, s, &name);
    squared = e == cudaSuccess;
  } else if ((in->format == FSR1_FORMAT_RGBA8_UNORM || in->format == FSR1_FORMAT_RGB10A2_UNORM) && !exact &&
             !(flags & (FSR1_FLAG_FORCE_DIRECT | FSR1_FLAG_PRECISE))) {
    p.options |= fused_square;
    e = launch_rcas_u_packed(p, (int)in->format, s, &name);
    squared = e == cudaSuccess;
  }
  if (e == cudaErrorNotSupported) p.options &= 3;
  if (e == cudaErrorNotSupported) e = launch_rcas_direct(p, (int)in->format, exact, s, &name);
  if (e != cudaSuccess) return cuda_fail(e);
  t_last_kernel = name;
  g_launches.fetch_add(1);
  if ((flags & FSR1_FLAG_OUTPUT_SQUARE) && !squared) return square_rows(out, y0, y1, s);  // direct / H-reference kernels: separate pass
  return FSR1_OK;
}

int fsr1_upscale(const fsr1_image* in, const fsr1_image* tmp, const fsr1_image* out, const uint32_t easu_con[16],
                 const uint32_t rcas_con[4], uint32_t y0, uint32_t y1, uint32_t flags, void* stream) {
  NvtxRange range("FSR1 upscale");
  if (!out) return FSR1_ERR_INVALID_ARGUMENT;
  if (y1 == 0) y1 = out->height;
  if (flags & FSR1_FLAG_NO_RCAS) return fsr1_easu(in, out, easu_con, y0, y1, flags, stream);
  // EASU also produces the one-row apron RCAS reads, so a row slab needs no second exchange
  const uint32_t e0 = y0 == 0 ? 0 : y0 - 1, e1 = y1 >= out->height ? out->height : y1 + 1;
  if ((flags & FSR1_FLAG_FUSED) && in && easu_con && rcas_con && in->format == FSR1_FORMAT_RGBA16F && out->format == FSR1_FORMAT_RGBA16F &&
      !(flags & (FSR1_FLAG_EXACT | FSR1_FLAG_FORCE_DIRECT | FSR1_FLAG_H_REFERENCE | FSR1_FLAG_PRECISE | FSR1_FLAG_RCAS_CLAMP |
                 FSR1_FLAG_RCAS_DENOISE | FSR1_FLAG_RCAS_PASSTHROUGH_ALPHA | FSR1_FLAG_OUTPUT_SQUARE | FSR1_FLAG_RCAS_HX2))) {
    int rc;
    if ((rc = check_image(in)) != FSR1_OK || (rc = check_image(out)) != FSR1_OK) return rc;
    if (flags & ~kAllFlags) return FSR1_ERR_INVALID_ARGUMENT;
    if (y0 >= y1 || y1 > out->height) return FSR1_ERR_INVALID_ARGUMENT;
    if (!window_holds(out, (int)y0, (int)y1 - 1)) return FSR1_ERR_WINDOW;
    uint32_t r0, r1;
    fsr1_easu_input_rows(easu_con, in->height, e0, e1, &r0, &r1);
    if (!window_holds(in, (int)r0, (int)r1)) return FSR1_ERR_WINDOW;
    EasuParams p;
    p.in = view_of(in);
    p.out = view_of(out);
    p.c0x = as_float(easu_con[0]); p.c0y = as_float(easu_con[1]); p.c0z = as_float(easu_con[2]); p.c0w = as_float(easu_con[3]);
    p.y0 = (int)y0; p.y1 = (int)y1;
    const char* name = "";
    if (t_sync) p.sync = *t_sync;
    const cudaError_t e = launch_fused_h(p, rcas_con[1], 0, static_cast<cudaStream_t>(stream), &name);
    if (e == cudaSuccess && t_sync) t_sync_used = true;
    if (e == cudaSuccess) {
      t_last_kernel = name;
      g_launches.fetch_add(1);
      return FSR1_OK;
    }
    if (e != cudaErrorNotSupported) return cuda_fail(e);
  }
  flags &= ~(uint32_t)FSR1_FLAG_FUSED;
  if (!tmp) return FSR1_ERR_INVALID_ARGUMENT;
  int rc = fsr1_easu(in, tmp, easu_con, e0, e1, flags & ~(uint32_t)FSR1_FLAG_OUTPUT_SQUARE, stream);  // last pass only
  if (rc != FSR1_OK) return rc;
  return fsr1_rcas(tmp, out, rcas_con, y0, y1, flags, stream);
}

// ---- pointwise companions ------------------------------------------------------------------------------
int fsr1_srtm(const fsr1_image* in, const fsr1_image* out, int inverse, uint32_t y0, uint32_t y1, void* stream) {
  return pointwise(inverse ? 2 : 1, in, nullptr, out, 0.0f, 0u, y0, y1, stream);
}

int fsr1_lfga(const fsr1_image* in, const fsr1_image* grain, const fsr1_image* out, float amount, uint32_t y0,
              uint32_t y1, void* stream) {
  if (!grain) return FSR1_ERR_INVALID_ARGUMENT;
  if (grain->format != FSR1_FORMAT_RGBA16F && grain->format != FSR1_FORMAT_RGBA32F) return FSR1_ERR_UNSUPPORTED;  // signed values
  return pointwise(3, in, grain, out, amount, 0u, y0, y1, stream);
}

int fsr1_tepd(const fsr1_image* in, const fsr1_image* dither, const fsr1_image* out, int bits, uint32_t frame,
              uint32_t y0, uint32_t y1, void* stream) {
  if (bits != 8 && bits != 10) return FSR1_ERR_INVALID_ARGUMENT;
  return pointwise(bits == 8 ? 4 : 5, in, dither, out, 0.0f, frame, y0, y1, stream);
}

int fsr1_srtm_h(const fsr1_image* in, const fsr1_image* out, int inverse, uint32_t y0, uint32_t y1, void* stream) {
  return pointwise_h(inverse ? 2 : 1, in, nullptr, out, 0.0f, 0u, y0, y1, stream);
}

int fsr1_lfga_h(const fsr1_image* in, const fsr1_image* grain, const fsr1_image* out, float amount, uint32_t y0, uint32_t y1,
                void* stream) {
  if (!grain) return FSR1_ERR_INVALID_ARGUMENT;
  return pointwise_h(3, in, grain, out, amount, 0u, y0, y1, stream);
}

int fsr1_tepd_h(const fsr1_image* in, const fsr1_image* dither, const fsr1_image* out, int bits, uint32_t frame, uint32_t y0,
                uint32_t y1, void* stream) {
  if (bits != 8 && bits != 10) return FSR1_ERR_INVALID_ARGUMENT;
  return pointwise_h(bits == 8 ? 4 : 5, in, dither, out, 0.0f, frame, y0, y1, stream);
}

// ---- context --------------------------------------------------------------------------------------
struct fsr1_context {
  uint32_t in_w, in_h, out_w, out_h, format;
  void* tmp;          // intermediate, out_w x out_h
  uint64_t tmp_pitch;
  void* dev_in;       // staging for the host-frame entry point (allocated on first use)
  void* dev_out;
  uint64_t in_pitch, out_pitch;
};

int fsr1_context_create(fsr1_context** ctx, uint32_t in_w, uint32_t in_h, uint32_t out_w, uint32_t out_h,
                        uint32_t format) {
  if (!ctx || !in_w || !in_h || !out_w || !out_h || !bytes_per_pixel(format)) return FSR1_ERR_INVALID_ARGUMENT;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return FSR1_ERR_NO_DEVICE;
  fsr1_context* c = new (std::nothrow) fsr1_context();
  if (!c) return FSR1_ERR_INVALID_ARGUMENT;
  memset(c, 0, sizeof *c);
  c->in_w = in_w; c->in_h = in_h; c->out_w = out_w; c->out_h = out_h; c->format = format;
  const uint64_t bpp = (uint64_t)bytes_per_pixel(format);
  c->tmp_pitch = ((uint64_t)out_w * bpp + 127) & ~(uint64_t)127;
  cudaError_t e = cudaMalloc(&c->tmp, c->tmp_pitch * out_h);
  if (e != cudaSuccess) { delete c; return cuda_fail(e); }
  *ctx = c;
  return FSR1_OK;
}

void fsr1_context_destroy(fsr1_context* c) {
  if (!c) return;
  cudaFree(c->tmp);
  cudaFree(c->dev_in);
  cudaFree(c->dev_out);
  delete c;
}

static int context_run(fsr1_context* c, void* in_dev, uint64_t in_pitch, void* out_dev, uint64_t out_pitch,
                       float sharpness, uint32_t flags, void* stream, uint32_t render_w = 0, uint32_t render_h = 0) {
  if (render_w == 0) render_w = c->in_w;
  if (render_h == 0) render_h = c->in_h;
  fsr1_image in = {in_dev, in_pitch, render_w, render_h, 0, render_h, c->format, 0};
  fsr1_image tmp = {c->tmp, c->tmp_pitch, c->out_w, c->out_h, 0, c->out_h, c->format, 0};
  fsr1_image out = {out_dev, out_pitch, c->out_w, c->out_h, 0, c->out_h, c->format, 0};
  uint32_t econ[16], rcon[4];
  // exactly what FSR_Filter::Upscale passes (sample/src/DX12/FSR_Filter.cpp:106,124)
  fsr1_easu_con(econ, (float)render_w, (float)render_h, (float)render_w, (float)render_h, (float)c->out_w, (float)c->out_h);
  fsr1_rcas_con(rcon, sharpness);
  return fsr1_upscale(&in, &tmp, &out, econ, rcon, 0, c->out_h, flags, stream);
}

int fsr1_context_upscale_render(fsr1_context* c, const void* in_dev, uint64_t in_pitch, uint32_t render_w, uint32_t render_h,
                                void* out_dev, uint64_t out_pitch, float sharpness, uint32_t flags, void* stream) {
  if (!c || !in_dev || !out_dev || !render_w || !render_h) return FSR1_ERR_INVALID_ARGUMENT;
  return context_run(c, const_cast<void*>(in_dev), in_pitch, out_dev, out_pitch, sharpness, flags, stream, render_w, render_h);
}

int fsr1_context_upscale(fsr1_context* c, const void* in_dev, uint64_t in_pitch, void* out_dev, uint64_t out_pitch,
                         float sharpness, uint32_t flags, void* stream) {
  if (!c || !in_dev || !out_dev) return FSR1_ERR_INVALID_ARGUMENT;
  return context_run(c, const_cast<void*>(in_dev), in_pitch, out_dev, out_pitch, sharpness, flags, stream);
}

int fsr1_context_upscale_host(fsr1_context* c, const void* in_host, uint64_t in_pitch, void* out_host,
                              uint64_t out_pitch, float sharpness, uint32_t flags, void* stream) {
  if (!c || !in_host || !out_host) return FSR1_ERR_INVALID_ARGUMENT;
  const uint64_t bpp = (uint64_t)bytes_per_pixel(c->format);
  if (in_pitch < c->in_w * bpp || out_pitch < c->out_w * bpp) return FSR1_ERR_INVALID_ARGUMENT;
  cudaError_t e;
  if (!c->dev_in || !c->dev_out) {  // both or neither: a failed second allocation leaves nothing half-initialised
    c->in_pitch = ((uint64_t)c->in_w * bpp + 127) & ~(uint64_t)127;
    c->out_pitch = ((uint64_t)c->out_w * bpp + 127) & ~(uint64_t)127;
    void *din = nullptr, *dout = nullptr;
    if ((e = cudaMalloc(&din, c->in_pitch * c->in_h)) != cudaSuccess) return cuda_fail(e);
    if ((e = cudaMalloc(&dout, c->out_pitch * c->out_h)) != cudaSuccess) { cudaFree(din); return cuda_fail(e); }
    c->dev_in = din;
    c->dev_out = dout;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  e = cudaMemcpy2DAsync(c->dev_in, c->in_pitch, in_host, in_pitch, c->in_w * bpp, c->in_h, cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return cuda_fail(e);
  int rc = context_run(c, c->dev_in, c->in_pitch, c->dev_out, c->out_pitch, sharpness, flags, stream);
  if (rc != FSR1_OK) return rc;
  e = cudaMemcpy2DAsync(out_host, out_pitch, c->dev_out, c->out_pitch, c->out_w * bpp, c->out_h, cudaMemcpyDeviceToHost, s);
  if (e != cudaSuccess) return cuda_fail(e);
  return FSR1_OK;
}

}  // extern "C"
