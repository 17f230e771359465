// fsr1_shard.cu — row-slab sharding of the EASU+RCAS path across the GPUs of one box, with the EASU input halo
// moved by DIRECT NVLink stores into the neighbour's memory (SURVEY.md §8(e) "Alternative": P2P mapped slabs).
//
// The reference has no multi-GPU path (sample/src/DX12/FSRSample.cpp:901 is a comment); this is new.  One fsr1_shard
// per rank (= per GPU; ranks may be processes, attached through CUDA IPC handles, or live in one process, attached
// by pointer).  The OUTPUT image is cut into `world` row slabs; rank k owns input rows [k inH/world, (k+1) inH/world)
// and needs 2-3 more rows each side (the EASU footprint of its slab plus the one-row apron RCAS reads).
//
// Data plane per frame (no NCCL, no host round trip, no collective, and NO extra launch on the streams the big kernels use):
//   comm stream   halo_push_kernel, 2 x 8 one-warp CTAs on a HIGH-PRIORITY stream (a 32-thread CTA fits beside the seven resident EASU
//   (high prio)   CTAs of an SM and is dispatched ahead of pending RCAS CTAs): half of them copy my top rows into the upper neighbour's
//                 window (its bottom halo), the others my bottom rows into the lower neighbour's — 128-bit stores over NVLink,
//                 __threadfence_system, then a release store of the frame's sequence number into the neighbour's `ready` flag.
//   EASU stream   the EASU (or fused) kernel itself carries the hand-shake (HaloSync, fsr1_common.cuh): every CTA's thread 0
//                 acquires MY `ready` flags before the CTA's first load, the last CTA to finish release-stores the sequence
//                 number into the neighbours' `credit` flags ("your rows in my window may be overwritten").  Kernels without
//                 that hook (UNORM, fp32, direct) get the same protocol from two one-warp kernels around them.
//   two compute streams, whole frames in turn: RCAS of frame i overlaps EASU of frame i+1 exactly as on one GPU.
// Flow control is by sequence numbers in device memory, so it is independent of host timing on either side: a push
// for the q-th use of a slot waits for credit q-1, EASU of use q waits for ready q.  Every spin is bounded (a wall
// clock timeout sets an error word instead of hanging the GPU).
#include <new>
#include <string.h>

#include "../../include/fsr1_b200.h"
#include "fsr1_common.cuh"

namespace {

using fsr1::spin_until;
using fsr1::st_release_sys;

constexpr uint32_t kFlagBytes = 4096;          // flags page at the start of the arena
constexpr uint32_t kMaxSlots = 128;

enum { kFromUp = 0, kFromDown = 1 };
// flag word index inside an arena's flags page
__host__ __device__ inline uint32_t ready_idx(uint32_t slot, int from) { return slot * 4 + from; }
__host__ __device__ inline uint32_t credit_idx(uint32_t slot, int from) { return slot * 4 + 2 + from; }
constexpr uint32_t kStatusIdx = kMaxSlots * 4;  // != 0: a spin timed out (value = 1 + which)
__host__ __device__ inline uint32_t counter_idx(uint32_t slot) { return kStatusIdx + 1 + slot; }  // CTAs of the slot's EASU that finished
__host__ __device__ inline uint32_t push_cnt_idx(uint32_t slot, int side) { return kStatusIdx + 1 + kMaxSlots + slot * 2 + side; }  // push parts done
constexpr int kPushParts = 8;  // one-warp CTAs per direction
constexpr uint32_t kTraceFrames = 256, kTraceWords = 8;  // fsr1_shard_trace: device timestamps of the last frames

struct PushSide {
  const uint4* src;        // my rows (local)
  uint4* dst;              // the neighbour's window rows (peer memory); nullptr = no neighbour on this side
  uint32_t n16;            // 16-byte units
  const uint32_t* credit;  // local: the neighbour has finished reading the previous use of this slot
  uint32_t* ready;         // peer: "your halo rows for use q are in place"
  uint32_t* parts_done;    // local: one-warp CTAs of this push that have finished
  unsigned long long* trace;  // optional: [3 + 2 side] first part started copying, [4 + 2 side] published
};

// kPushParts one-warp CTAs per direction (blockIdx.x / kPushParts = 0: up, 1: down), each moving 1/kPushParts of the rows.  32 threads
// and < 32 registers: such a CTA fits into what seven resident EASU CTAs leave of an SM, so a push never waits for a big kernel to
// finish.  Stores over NVLink are fire-and-forget; two 16-byte loads per lane are kept in flight.  The last part to finish (counter in
// local memory) publishes the sequence number.
__global__ void __launch_bounds__(32) halo_push_kernel(const PushSide up, const PushSide down, const uint32_t q, uint32_t* status) {
  const int side = blockIdx.x / kPushParts, part = blockIdx.x - side * kPushParts;
  const PushSide s = side == 0 ? up : down;
  if (!s.dst) return;
  int ok = 1;
  if (threadIdx.x == 0) {
    ok = spin_until(s.credit, q - 1) ? 1 : 0;
    if (!ok) atomicExch(status, 1u + side);
  }
  ok = __shfl_sync(0xffffffffu, ok, 0);
  if (s.trace && part == 0 && threadIdx.x == 0) s.trace[3 + 2 * side] = fsr1::global_ns();
  if (ok) {
    const uint32_t per = (s.n16 + kPushParts - 1) / kPushParts, first = part * per;
    const uint32_t end = first + per < s.n16 ? first + per : s.n16;
    const uint32_t t0 = first + threadIdx.x;
    uint32_t left = end > t0 ? (end - t0 + 31) / 32 : 0;  // 16-byte units this lane moves
    const uint4* src = s.src + t0;
    uint4* dst = s.dst + t0;
#pragma unroll 1
    for (; left >= 2; left -= 2, src += 64, dst += 64) {
      const uint4 a = src[0], b = src[32];
      dst[0] = a;
      dst[32] = b;
    }
    if (left) dst[0] = src[0];
    __threadfence_system();
  }
  __syncwarp();
  if (threadIdx.x == 0 && atomicAdd(s.parts_done, 1u) == kPushParts - 1) {  // last part (a timed-out push publishes nothing)
    atomicExch(s.parts_done, 0u);
    __threadfence_system();  // acquire side of the counter: the other parts' stores (fenced before their increment) precede the flag
    if (ok) st_release_sys(s.ready, q);
    if (s.trace) s.trace[4 + 2 * side] = fsr1::global_ns();
  }
}

__global__ void __launch_bounds__(32) halo_wait_kernel(const uint32_t* ready_up, const uint32_t* ready_down, const uint32_t q, uint32_t* status) {
  const uint32_t* f = threadIdx.x == 0 ? ready_up : (threadIdx.x == 1 ? ready_down : nullptr);
  if (f && !spin_until(f, q)) atomicExch(status, 3u + threadIdx.x);
}

__global__ void __launch_bounds__(32) credit_signal_kernel(uint32_t* credit_up, uint32_t* credit_down, const uint32_t q) {
  uint32_t* f = threadIdx.x == 0 ? credit_up : (threadIdx.x == 1 ? credit_down : nullptr);
  if (f) st_release_sys(f, q);
}

int bpp_of(uint32_t fmt) {
  switch (fmt) {
    case FSR1_FORMAT_RGBA16F: return 8;
    case FSR1_FORMAT_RGBA32F: return 16;
    case FSR1_FORMAT_RGBA8_UNORM: case FSR1_FORMAT_RGB10A2_UNORM: return 4;
    default: return 0;
  }
}

struct Rows { uint32_t a, b; };  // [a, b)

}  // namespace

struct fsr1_shard {
  uint32_t in_w, in_h, out_w, out_h, format, world, rank, slots, flags;
  int device;
  uint32_t econ[16], rcon[4];
  // geometry of THIS rank
  Rows out_rows, easu_rows, owned, needed, window;
  // arena: [flags page][slot 0 window][slot 1 window]...; identical layout on every rank
  uint64_t pitch, slot_stride, arena_bytes;
  uint32_t win_rows_max;
  unsigned char* arena;
  unsigned char* peer[2];      // [kFromUp] = arena of rank-1, [kFromDown] = arena of rank+1 (mapped), nullptr = none
  bool peer_is_ipc[2];
  uint32_t peer_win0[2];       // first logical row of the neighbour's window
  Rows send[2];                // my rows the neighbour needs
  unsigned char* tmp;          // slots x rows easu_rows
  unsigned char* out;          // slots x rows out_rows
  uint64_t out_pitch, tmp_slot_stride, out_slot_stride;
  uint32_t seq[kMaxSlots];
  cudaStream_t s_comm, s_easu, s_rcas;
  cudaEvent_t ev_in[kMaxSlots], ev_push[kMaxSlots], ev_rcas[kMaxSlots];
  bool attached;
  unsigned long long* trace;  // FSR1_SHARD_TRACE: kTraceFrames x kTraceWords globaltimer stamps (device memory), else null
  unsigned long long frames;  // frames submitted
  bool inkernel_sync;  // the EASU / fused kernel of this configuration carries the hand-shake itself (HaloSync)
};

namespace {

Rows plan_out_rows(const fsr1_shard* s, uint32_t r) {
  return Rows{(uint32_t)((uint64_t)r * s->out_h / s->world), (uint32_t)((uint64_t)(r + 1) * s->out_h / s->world)};
}
Rows plan_easu_rows(const fsr1_shard* s, uint32_t r) {
  const Rows o = plan_out_rows(s, r);
  return Rows{o.a == 0 ? 0 : o.a - 1, o.b >= s->out_h ? s->out_h : o.b + 1};
}
Rows plan_owned(const fsr1_shard* s, uint32_t r) {
  return Rows{(uint32_t)((uint64_t)r * s->in_h / s->world), (uint32_t)((uint64_t)(r + 1) * s->in_h / s->world)};
}
Rows plan_needed(const fsr1_shard* s, uint32_t r) {
  const Rows e = plan_easu_rows(s, r);
  uint32_t first = 0, last = 0;
  fsr1_easu_input_rows(s->econ, s->in_h, e.a, e.b, &first, &last);
  return Rows{first, last + 1};
}
Rows plan_window(const fsr1_shard* s, uint32_t r) {
  const Rows o = plan_owned(s, r), n = plan_needed(s, r);
  return Rows{o.a < n.a ? o.a : n.a, o.b > n.b ? o.b : n.b};
}

int cuda_rc(cudaError_t e) { return e == cudaSuccess ? FSR1_OK : FSR1_ERR_CUDA; }

struct DeviceGuard {  // calls may come from a thread whose current device is another GPU (one process, several ranks)
  int prev;
  explicit DeviceGuard(int dev) : prev(-1) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace

extern "C" {

int fsr1_shard_create(fsr1_shard** out_sh, uint32_t in_w, uint32_t in_h, uint32_t out_w, uint32_t out_h, uint32_t format,
                      uint32_t world, uint32_t rank, uint32_t slots, float sharpness_stops, uint32_t flags) {
  if (!out_sh || !in_w || !in_h || !out_w || !out_h || !world || rank >= world || !slots || slots > kMaxSlots) return FSR1_ERR_INVALID_ARGUMENT;
  const int bpp = bpp_of(format);
  if (!bpp) return FSR1_ERR_INVALID_ARGUMENT;
  if (world > in_h || world > out_h) return FSR1_ERR_INVALID_ARGUMENT;  // no empty slabs
  fsr1_shard* s = new (std::nothrow) fsr1_shard();
  if (!s) return FSR1_ERR_INVALID_ARGUMENT;
  memset(s, 0, sizeof *s);
  s->in_w = in_w; s->in_h = in_h; s->out_w = out_w; s->out_h = out_h; s->format = format;
  s->world = world; s->rank = rank; s->slots = slots; s->flags = flags;
  if (cudaGetDevice(&s->device) != cudaSuccess) { delete s; return FSR1_ERR_NO_DEVICE; }
  fsr1_easu_con(s->econ, (float)in_w, (float)in_h, (float)in_w, (float)in_h, (float)out_w, (float)out_h);
  fsr1_rcas_con(s->rcon, sharpness_stops);
  s->out_rows = plan_out_rows(s, rank);
  s->easu_rows = plan_easu_rows(s, rank);
  s->owned = plan_owned(s, rank);
  s->needed = plan_needed(s, rank);
  s->window = plan_window(s, rank);
  // every rank's halo must come from its direct neighbours only (true whenever a slab is taller than the halo)
  s->win_rows_max = 0;
  for (uint32_t r = 0; r < world; r++) {
    const Rows n = plan_needed(s, r), w = plan_window(s, r);
    const uint32_t lo = r == 0 ? 0 : plan_owned(s, r - 1).a, hi = r + 1 == world ? in_h : plan_owned(s, r + 1).b;
    if (n.a < lo || n.b > hi) { delete s; return FSR1_ERR_UNSUPPORTED; }
    if (w.b - w.a > s->win_rows_max) s->win_rows_max = w.b - w.a;
  }
  for (int side = 0; side < 2; side++) {
    const bool has = side == kFromUp ? rank > 0 : rank + 1 < world;
    s->send[side] = Rows{0, 0};
    if (!has) continue;
    const uint32_t peer = side == kFromUp ? rank - 1 : rank + 1;
    const Rows pn = plan_needed(s, peer);
    const uint32_t a = s->owned.a > pn.a ? s->owned.a : pn.a, b = s->owned.b < pn.b ? s->owned.b : pn.b;
    if (b > a) s->send[side] = Rows{a, b};
    s->peer_win0[side] = plan_window(s, peer).a;
  }
  s->pitch = ((uint64_t)in_w * bpp + 127) & ~(uint64_t)127;
  s->slot_stride = ((uint64_t)s->win_rows_max * s->pitch + 255) & ~(uint64_t)255;
  s->arena_bytes = kFlagBytes + s->slot_stride * slots;
  s->out_pitch = ((uint64_t)out_w * bpp + 127) & ~(uint64_t)127;
  s->tmp_slot_stride = (uint64_t)(s->easu_rows.b - s->easu_rows.a) * s->out_pitch;
  s->out_slot_stride = (uint64_t)(s->out_rows.b - s->out_rows.a) * s->out_pitch;
  cudaError_t e;
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);  // numerically lowest = most urgent
  // cudaMalloc (not a pool / VMM allocation): the arena must be exportable through cudaIpcGetMemHandle
  if ((e = cudaMalloc((void**)&s->arena, s->arena_bytes)) != cudaSuccess || (e = cudaMemset(s->arena, 0, s->arena_bytes)) != cudaSuccess ||
      (e = cudaMalloc((void**)&s->tmp, s->tmp_slot_stride * slots)) != cudaSuccess ||
      (e = cudaMalloc((void**)&s->out, s->out_slot_stride * slots)) != cudaSuccess ||
      (e = cudaStreamCreateWithPriority(&s->s_comm, cudaStreamNonBlocking, prio_hi)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&s->s_easu, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&s->s_rcas, cudaStreamNonBlocking)) != cudaSuccess) {
    fsr1_shard_destroy(s);
    return FSR1_ERR_CUDA;
  }
  if ((flags & FSR1_SHARD_TRACE) && (cudaMalloc((void**)&s->trace, sizeof(unsigned long long) * kTraceFrames * kTraceWords) != cudaSuccess ||
                                     cudaMemset(s->trace, 0, sizeof(unsigned long long) * kTraceFrames * kTraceWords) != cudaSuccess)) {
    fsr1_shard_destroy(s);
    return FSR1_ERR_CUDA;
  }
  for (uint32_t i = 0; i < slots; i++) {
    if (cudaEventCreateWithFlags(&s->ev_in[i], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_push[i], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_rcas[i], cudaEventDisableTiming) != cudaSuccess) {
      fsr1_shard_destroy(s);
      return FSR1_ERR_CUDA;
    }
  }
  if (world > 1) {
    // Load every kernel a frame uses NOW.  CUDA loads kernels lazily, on first launch, and loading may synchronise the context: a
    // first-use load issued while a flag-waiting kernel spins would wait for that kernel, which (several ranks in ONE process) may be
    // waiting for work this very host thread has not submitted yet.  One dry frame on the zero-filled slot 0 (no halo protocol).
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, halo_push_kernel) != cudaSuccess || cudaFuncGetAttributes(&fa, halo_wait_kernel) != cudaSuccess ||
        cudaFuncGetAttributes(&fa, credit_signal_kernel) != cudaSuccess) {
      fsr1_shard_destroy(s);
      return FSR1_ERR_CUDA;
    }
    fsr1_image win, out;
    fsr1_shard_window(s, 0, &win);
    fsr1_shard_output(s, 0, &out);
    fsr1_image tmp0 = {s->tmp, s->out_pitch, s->out_w, s->out_h, s->easu_rows.a, s->easu_rows.b - s->easu_rows.a, s->format, 0};
    const uint32_t kflags = flags & ~(uint32_t)(FSR1_SHARD_ONE_STREAM | FSR1_SHARD_SKIP_HALO | FSR1_SHARD_TRACE);
    const fsr1::HaloSync none = {};
    fsr1::set_halo_sync(&none);  // does this configuration's kernel take the hand-shake? (null pointers: a no-op inside the kernel)
    int rc = fsr1_upscale(&win, &tmp0, &out, s->econ, s->rcon, s->out_rows.a, s->out_rows.b, kflags, s->s_easu);
    s->inkernel_sync = fsr1::halo_sync_consumed();
    fsr1::set_halo_sync(nullptr);
    if (rc == FSR1_OK && (kflags & FSR1_FLAG_FUSED))  // the fused path may fall back to the two kernels for other frames: load those too
      rc = fsr1_upscale(&win, &tmp0, &out, s->econ, s->rcon, s->out_rows.a, s->out_rows.b, kflags & ~(uint32_t)FSR1_FLAG_FUSED, s->s_easu);
    if (rc != FSR1_OK) { fsr1_shard_destroy(s); return rc; }
  }
  if ((e = cudaDeviceSynchronize()) != cudaSuccess) { fsr1_shard_destroy(s); return FSR1_ERR_CUDA; }  // flags are zero before anyone attaches
  s->attached = world == 1 || (flags & FSR1_SHARD_SKIP_HALO);
  *out_sh = s;
  return FSR1_OK;
}

void fsr1_shard_destroy(fsr1_shard* s) {
  if (!s) return;
  DeviceGuard g(s->device);
  cudaDeviceSynchronize();
  for (int side = 0; side < 2; side++)
    if (s->peer[side] && s->peer_is_ipc[side]) cudaIpcCloseMemHandle(s->peer[side]);
  for (uint32_t i = 0; i < s->slots && i < kMaxSlots; i++) {
    if (s->ev_in[i]) cudaEventDestroy(s->ev_in[i]);
    if (s->ev_push[i]) cudaEventDestroy(s->ev_push[i]);
    if (s->ev_rcas[i]) cudaEventDestroy(s->ev_rcas[i]);
  }
  if (s->s_comm) cudaStreamDestroy(s->s_comm);
  if (s->s_easu) cudaStreamDestroy(s->s_easu);
  if (s->s_rcas) cudaStreamDestroy(s->s_rcas);
  cudaFree(s->arena);
  cudaFree(s->trace);
  cudaFree(s->tmp);
  cudaFree(s->out);
  delete s;
}

int fsr1_shard_geometry(const fsr1_shard* s, fsr1_shard_info* info) {
  if (!s || !info) return FSR1_ERR_INVALID_ARGUMENT;
  info->out_row0 = s->out_rows.a; info->out_row1 = s->out_rows.b;
  info->easu_row0 = s->easu_rows.a; info->easu_row1 = s->easu_rows.b;
  info->owned_row0 = s->owned.a; info->owned_row1 = s->owned.b;
  info->needed_row0 = s->needed.a; info->needed_row1 = s->needed.b;
  info->window_row0 = s->window.a; info->window_row1 = s->window.b;
  info->send_up_row0 = s->send[kFromUp].a; info->send_up_row1 = s->send[kFromUp].b;
  info->send_down_row0 = s->send[kFromDown].a; info->send_down_row1 = s->send[kFromDown].b;
  info->halo_recv_bytes = (uint64_t)((s->owned.a - s->window.a) + (s->window.b - s->owned.b)) * s->in_w * bpp_of(s->format);
  info->arena_bytes = s->arena_bytes;
  return FSR1_OK;
}

int fsr1_shard_export(const fsr1_shard* s, void* handle) {
  if (!s || !handle) return FSR1_ERR_INVALID_ARGUMENT;
  static_assert(sizeof(cudaIpcMemHandle_t) == FSR1_SHARD_HANDLE_BYTES, "handle size");
  DeviceGuard g(s->device);
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, s->arena) != cudaSuccess) return FSR1_ERR_CUDA;
  memcpy(handle, &h, sizeof h);
  return FSR1_OK;
}

void* fsr1_shard_arena(const fsr1_shard* s) { return s ? s->arena : nullptr; }

static int attach_done(fsr1_shard* s) {
  s->attached = (s->rank == 0 || s->peer[kFromUp]) && (s->rank + 1 == s->world || s->peer[kFromDown]);
  return s->attached ? FSR1_OK : FSR1_ERR_INVALID_ARGUMENT;
}

int fsr1_shard_attach(fsr1_shard* s, const void* handles, uint32_t count) {
  if (!s || !handles || count != s->world) return FSR1_ERR_INVALID_ARGUMENT;
  DeviceGuard g(s->device);
  for (int side = 0; side < 2; side++) {
    const bool has = side == kFromUp ? s->rank > 0 : s->rank + 1 < s->world;
    if (!has || s->peer[side]) continue;
    const uint32_t peer = side == kFromUp ? s->rank - 1 : s->rank + 1;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const unsigned char*)handles + (size_t)peer * FSR1_SHARD_HANDLE_BYTES, sizeof h);
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) return FSR1_ERR_CUDA;
    s->peer[side] = (unsigned char*)p;
    s->peer_is_ipc[side] = true;
  }
  return attach_done(s);
}

int fsr1_shard_attach_local(fsr1_shard* s, fsr1_shard* up, fsr1_shard* down) {
  if (!s) return FSR1_ERR_INVALID_ARGUMENT;
  DeviceGuard g(s->device);
  fsr1_shard* nb[2] = {up, down};
  for (int side = 0; side < 2; side++) {
    const bool has = side == kFromUp ? s->rank > 0 : s->rank + 1 < s->world;
    if (!has) continue;
    fsr1_shard* n = nb[side];
    if (!n || n->world != s->world || n->rank != (side == kFromUp ? s->rank - 1 : s->rank + 1) || n->arena_bytes != s->arena_bytes)
      return FSR1_ERR_INVALID_ARGUMENT;
    if (n->device != s->device) {
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, s->device, n->device) != cudaSuccess || !can) return FSR1_ERR_UNSUPPORTED;
      cudaError_t e = cudaDeviceEnablePeerAccess(n->device, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else if (e != cudaSuccess) return FSR1_ERR_CUDA;
    }
    s->peer[side] = n->arena;
    s->peer_is_ipc[side] = false;
  }
  return attach_done(s);
}

static fsr1_image make_img(void* data, uint64_t pitch, uint32_t w, uint32_t h, uint32_t row0, uint32_t rows, uint32_t fmt) {
  fsr1_image im;
  im.data = data; im.pitch_bytes = pitch; im.width = w; im.height = h; im.row0 = row0; im.rows = rows; im.format = fmt; im.reserved = 0;
  return im;
}
static unsigned char* window_of(const fsr1_shard* s, unsigned char* arena, uint32_t slot) { return arena + kFlagBytes + (uint64_t)slot * s->slot_stride; }

int fsr1_shard_input(const fsr1_shard* s, uint32_t slot, fsr1_image* owned) {
  if (!s || !owned || slot >= s->slots) return FSR1_ERR_INVALID_ARGUMENT;
  *owned = make_img(window_of(s, s->arena, slot) + (uint64_t)(s->owned.a - s->window.a) * s->pitch, s->pitch, s->in_w, s->in_h, s->owned.a,
                    s->owned.b - s->owned.a, s->format);
  return FSR1_OK;
}

int fsr1_shard_window(const fsr1_shard* s, uint32_t slot, fsr1_image* window) {
  if (!s || !window || slot >= s->slots) return FSR1_ERR_INVALID_ARGUMENT;
  *window = make_img(window_of(s, s->arena, slot), s->pitch, s->in_w, s->in_h, s->window.a, s->window.b - s->window.a, s->format);
  return FSR1_OK;
}

int fsr1_shard_output(const fsr1_shard* s, uint32_t slot, fsr1_image* out) {
  if (!s || !out || slot >= s->slots) return FSR1_ERR_INVALID_ARGUMENT;
  *out = make_img(s->out + (uint64_t)slot * s->out_slot_stride, s->out_pitch, s->out_w, s->out_h, s->out_rows.a, s->out_rows.b - s->out_rows.a,
                  s->format);
  return FSR1_OK;
}

int fsr1_shard_submit(fsr1_shard* s, uint32_t slot, void* stream) {
  if (!s || slot >= s->slots) return FSR1_ERR_INVALID_ARGUMENT;
  if (!s->attached) return FSR1_ERR_INVALID_ARGUMENT;
  DeviceGuard g(s->device);
  cudaStream_t caller = static_cast<cudaStream_t>(stream);
  const bool one_stream = (s->flags & FSR1_SHARD_ONE_STREAM) != 0;
  cudaStream_t se = s->s_easu, sr = one_stream ? s->s_easu : s->s_rcas;
  const uint32_t q = ++s->seq[slot];
  struct FrameCount { fsr1_shard* s; ~FrameCount() { s->frames++; } } frame_count{s};
  uint32_t* flags = reinterpret_cast<uint32_t*>(s->arena);
  cudaError_t e;
  if ((e = cudaEventRecord(s->ev_in[slot], caller)) != cudaSuccess) return cuda_rc(e);
  const bool skip_halo = (s->flags & FSR1_SHARD_SKIP_HALO) != 0;  // measurement only: what the frame costs without the exchange
  const bool up = s->rank > 0 && !skip_halo, down = s->rank + 1 < s->world && !skip_halo;
  if (up || down) {
    PushSide ps[2];
    for (int side = 0; side < 2; side++) {
      ps[side] = PushSide{nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr};
      const bool has = side == kFromUp ? up : down;
      if (!has) continue;
      const Rows r = s->send[side];
      uint32_t* pf = reinterpret_cast<uint32_t*>(s->peer[side]);
      ps[side].src = reinterpret_cast<const uint4*>(window_of(s, s->arena, slot) + (uint64_t)(r.a - s->window.a) * s->pitch);
      ps[side].dst = reinterpret_cast<uint4*>(window_of(s, s->peer[side], slot) + (uint64_t)(r.a - s->peer_win0[side]) * s->pitch);
      ps[side].n16 = (uint32_t)((uint64_t)(r.b - r.a) * s->pitch / 16);
      ps[side].credit = flags + credit_idx(slot, side);
      ps[side].parts_done = flags + push_cnt_idx(slot, side);
      ps[side].trace = s->trace ? s->trace + (size_t)(s->frames % kTraceFrames) * kTraceWords : nullptr;
      // I am the neighbour's lower (upper) peer when I push up (down)
      ps[side].ready = pf + ready_idx(slot, side == kFromUp ? kFromDown : kFromUp);
    }
    if ((e = cudaStreamWaitEvent(s->s_comm, s->ev_in[slot], 0)) != cudaSuccess) return cuda_rc(e);
    halo_push_kernel<<<2 * kPushParts, 32, 0, s->s_comm>>>(ps[kFromUp], ps[kFromDown], q, flags + kStatusIdx);
    if ((e = cudaGetLastError()) != cudaSuccess) return cuda_rc(e);
    if ((e = cudaEventRecord(s->ev_push[slot], s->s_comm)) != cudaSuccess) return cuda_rc(e);
  }
  fsr1_image win, out;
  fsr1_shard_window(s, slot, &win);
  fsr1_shard_output(s, slot, &out);
  fsr1_image tmp = make_img(s->tmp + (uint64_t)slot * s->tmp_slot_stride, s->out_pitch, s->out_w, s->out_h, s->easu_rows.a,
                            s->easu_rows.b - s->easu_rows.a, s->format);
  const uint32_t kflags = s->flags & ~(uint32_t)(FSR1_SHARD_ONE_STREAM | FSR1_SHARD_SKIP_HALO | FSR1_SHARD_TRACE);
  const bool fused = (kflags & FSR1_FLAG_FUSED) != 0;
  // The whole frame runs on ONE stream, consecutive frames on the two streams in turn: RCAS of frame i (ALU / XU / HBM-bound) overlaps
  // EASU of frame i+1 (FMA-pipe-bound) without an event between the two kernels of a frame.  Measured on B200 against "EASU of every
  // frame on one stream, RCAS on the other": 74.0 vs 76.2 us per frame at N = 1, 77.6 vs 81.5 at N = 2 (three driver calls fewer).
  cudaStream_t sk = (!one_stream && (s->frames & 1)) ? sr : se;
  if ((e = cudaStreamWaitEvent(sk, s->ev_in[slot], 0)) != cudaSuccess) return cuda_rc(e);
  if (q > 1 && !one_stream && (e = cudaStreamWaitEvent(sk, s->ev_rcas[slot], 0)) != cudaSuccess) return cuda_rc(e);  // the slot's intermediate / output are free
  fsr1::HaloSync hs = {};
  hs.ready[kFromUp] = up ? flags + ready_idx(slot, kFromUp) : nullptr;
  hs.ready[kFromDown] = down ? flags + ready_idx(slot, kFromDown) : nullptr;
  hs.credit[kFromUp] = up ? reinterpret_cast<uint32_t*>(s->peer[kFromUp]) + credit_idx(slot, kFromDown) : nullptr;
  hs.credit[kFromDown] = down ? reinterpret_cast<uint32_t*>(s->peer[kFromDown]) + credit_idx(slot, kFromUp) : nullptr;
  hs.counter = flags + counter_idx(slot);
  hs.status = flags + kStatusIdx;
  hs.trace = s->trace ? s->trace + (size_t)(s->frames % kTraceFrames) * kTraceWords : nullptr;
  hs.seq = q;
  const bool shake = up || down, inkernel = shake && s->inkernel_sync;
  if (shake && !inkernel) {
    halo_wait_kernel<<<1, 32, 0, sk>>>(hs.ready[kFromUp], hs.ready[kFromDown], q, flags + kStatusIdx);
    if ((e = cudaGetLastError()) != cudaSuccess) return cuda_rc(e);
  }
  if (inkernel) fsr1::set_halo_sync(&hs);
  int rc = fused ? fsr1_upscale(&win, &tmp, &out, s->econ, s->rcon, s->out_rows.a, s->out_rows.b, kflags, sk)
                 : fsr1_easu(&win, &tmp, s->econ, s->easu_rows.a, s->easu_rows.b, kflags & ~(uint32_t)FSR1_FLAG_OUTPUT_SQUARE, sk);
  const bool took = inkernel && fsr1::halo_sync_consumed();
  fsr1::set_halo_sync(nullptr);
  if (rc != FSR1_OK) return rc;
  if (inkernel && !took) return FSR1_ERR_UNSUPPORTED;  // cannot happen: the capability was probed with this configuration
  if (shake && !inkernel) {
    credit_signal_kernel<<<1, 32, 0, sk>>>(hs.credit[kFromUp], hs.credit[kFromDown], q);
    if ((e = cudaGetLastError()) != cudaSuccess) return cuda_rc(e);
  }
  if (!fused) {
    rc = fsr1_rcas(&tmp, &out, s->rcon, s->out_rows.a, s->out_rows.b, kflags, sk);
    if (rc != FSR1_OK) return rc;
  }
  if ((e = cudaEventRecord(s->ev_rcas[slot], sk)) != cudaSuccess) return cuda_rc(e);
  return FSR1_OK;
}

int fsr1_shard_wait(fsr1_shard* s, uint32_t slot, void* stream) {
  if (!s || slot >= s->slots) return FSR1_ERR_INVALID_ARGUMENT;
  if (s->seq[slot] == 0) return FSR1_OK;
  DeviceGuard g(s->device);
  cudaStream_t caller = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  if ((e = cudaStreamWaitEvent(caller, s->ev_rcas[slot], 0)) != cudaSuccess) return cuda_rc(e);
  if (s->world > 1 && !(s->flags & FSR1_SHARD_SKIP_HALO) && (e = cudaStreamWaitEvent(caller, s->ev_push[slot], 0)) != cudaSuccess)
    return cuda_rc(e);  // my rows have left
  return FSR1_OK;
}

int fsr1_shard_trace(fsr1_shard* s, uint64_t* out, uint32_t max_frames, uint32_t* n_frames) {
  if (!s || !out || !n_frames) return FSR1_ERR_INVALID_ARGUMENT;
  *n_frames = 0;
  if (!s->trace) return FSR1_OK;
  DeviceGuard g(s->device);
  uint32_t n = (uint32_t)(s->frames < kTraceFrames ? s->frames : kTraceFrames);
  if (n > max_frames) n = max_frames;
  // oldest first: frames [frames - n, frames)
  for (uint32_t i = 0; i < n; i++) {
    const unsigned long long f = s->frames - n + i;
    if (cudaMemcpy(out + (size_t)i * kTraceWords, s->trace + (size_t)(f % kTraceFrames) * kTraceWords, sizeof(unsigned long long) * kTraceWords,
                   cudaMemcpyDeviceToHost) != cudaSuccess)
      return FSR1_ERR_CUDA;
  }
  *n_frames = n;
  return FSR1_OK;
}

int fsr1_shard_status(fsr1_shard* s) {
  if (!s) return FSR1_ERR_INVALID_ARGUMENT;
  DeviceGuard g(s->device);
  uint32_t st = 0;
  if (cudaMemcpy(&st, reinterpret_cast<uint32_t*>(s->arena) + kStatusIdx, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return FSR1_ERR_CUDA;
  if (st != 0) fsr1::set_last_detail(1000 + (int)st + 10 * (int)s->rank);  // 1 / 2: push up / down starved of credit; 3 / 4: halo from above / below never came
  return st == 0 ? FSR1_OK : FSR1_ERR_TIMEOUT;
}

}  // extern "C"
