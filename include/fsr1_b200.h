/* fsr1_b200.h — C ABI of the B200-native FSR 1.0 hot path (EASU upsample + RCAS sharpen).
 *
 * Plain C, plain pointers and sizes; no CUDA or torch types appear in any signature (a stream is
 * passed as void* = cudaStream_t, NULL = the legacy default stream).  This is what a binding in
 * the reference's host code (or any FFI: ctypes, cgo, JNI ...) attaches to; see INTEGRATION.md.
 *
 * What each entry point replaces in the reference (GPUOpen-Effects/FidelityFX-FSR):
 *   fsr1_easu      the EASU dispatch: shader FsrEasuF/FsrEasuH (ffx-fsr/ffx_fsr1.h:315-437, 505-593)
 *                  entered from mainCS/CurrFilter (sample/src/DX12/FSR_Pass.hlsl:68-118), recorded by
 *                  FSR_Filter::Upscale -> m_easu.Draw (sample/src/DX12/FSR_Filter.cpp:121,135)
 *   fsr1_rcas      the RCAS dispatch: FsrRcasF/FsrRcasH (ffx-fsr/ffx_fsr1.h:684-769, 782-866),
 *                  m_rcas.Draw (sample/src/DX12/FSR_Filter.cpp:131)
 *   fsr1_upscale   the whole of FSR_Filter::Upscale (sample/src/DX12/FSR_Filter.cpp:101-141):
 *                  EASU -> (barrier) -> RCAS through a display-sized intermediate
 *   fsr1_context_* FSR_Filter::OnCreateWindowSizeDependentResources / OnDestroy... (FSR_Filter.cpp:70-99):
 *                  owns the intermediate image (and, for the *_host call, device staging buffers)
 *   fsr1_upscale_host  same as fsr1_upscale for callers whose frames live in HOST memory: copies the
 *                  input up, runs both passes, copies the result back, all on one stream
 * The constant blocks (con0..con3, rcas con) are EXACTLY the uint32[4] words FsrEasuCon /
 * FsrEasuConOffset / FsrRcasCon produce (include/fsr1_host.h, or the reference's own header).
 *
 * Semantics fixed by the reference and reproduced here:
 *   - images are row-major RGBA, 4 x fp16 (FSR1_FORMAT_RGBA16F, the reference's rgba16f path) or
 *     4 x fp32 (FSR1_FORMAT_RGBA32F, the SAMPLE_SLOW_FALLBACK path); EASU/RCAS read RGB, ignore A,
 *     and store A = 1 (FSR_Pass.hlsl:80,95)
 *   - EASU taps are clamped to the edge of the input RESOURCE (linear/clamp sampler, FSR_Filter.cpp:48-53)
 *   - RCAS taps outside the image read 0 (D3D12 Load); FSR1_FLAG_RCAS_CLAMP selects clamp instead
 *   - fp32 images run the F algorithm in fp32; fp16 images run a packed-half implementation whose
 *     results stay within 1e-2 of the fp32 algorithm on the same (quantised) input; UNORM images (the
 *     formats the sample renders into, FSR_Filter.cpp:72-73) run the F algorithm in fp32 on the D3D
 *     unorm<->float conversions (c/(2^n-1); clamp, scale, +0.5, truncate)
 * All launch calls are asynchronous with respect to the host and allocate nothing
 * (fsr1_context_create and fsr1_upscale_host's first use are the only allocating calls).
 * Thread-safe for distinct contexts/streams.  Every function returns FSR1_OK or a negative fsr1 error;
 * CUDA failures are reported as FSR1_ERR_CUDA and the CUDA error is kept for fsr1_last_cuda_error().
 */
#ifndef FSR1_B200_H
#define FSR1_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSR1_ABI_VERSION 3  /* additions that leave existing callers untouched keep the number: FSR1_FLAG_RCAS_HX2, fsr1_srtm_h / fsr1_lfga_h / fsr1_tepd_h */

enum {
  FSR1_OK = 0,
  FSR1_ERR_INVALID_ARGUMENT = -1, /* null pointer, zero size, bad row range, unknown format/flag   */
  FSR1_ERR_UNSUPPORTED = -2,      /* format combination the kernels do not implement              */
  FSR1_ERR_WINDOW = -3,           /* an image window (row0/rows) does not hold the rows the pass reads/writes */
  FSR1_ERR_CUDA = -4,             /* a CUDA call failed; see fsr1_last_cuda_error()                */
  FSR1_ERR_NO_DEVICE = -5,        /* no usable sm_100 device / driver                              */
  FSR1_ERR_TIMEOUT = -6           /* fsr1_shard_*: a neighbour's halo or credit did not arrive in time */
};

enum {
  FSR1_FORMAT_RGBA16F = 1,
  FSR1_FORMAT_RGBA32F = 2,
  FSR1_FORMAT_RGBA8_UNORM = 3,    /* 4 B/px, byte order R,G,B,A (DXGI_FORMAT_R8G8B8A8_UNORM)                      */
  FSR1_FORMAT_RGB10A2_UNORM = 4   /* 4 B/px, bits 0-9 R, 10-19 G, 20-29 B, 30-31 A (DXGI_FORMAT_R10G10B10A2_UNORM) */
};

enum {
  FSR1_FLAG_RCAS_CLAMP = 1u << 0,   /* RCAS out-of-image taps clamp instead of reading 0               */
  FSR1_FLAG_EXACT = 1u << 1,        /* fp32 images only: no FMA contraction, IEEE division — bit-identical
                                       to the reference source compiled with -ffp-contract=off          */
  FSR1_FLAG_FORCE_DIRECT = 1u << 2, /* skip the TMA/shared-memory kernels, use the direct-load kernels   */
  FSR1_FLAG_NO_RCAS = 1u << 3,      /* fsr1_upscale*: EASU straight to the output (bUseRcas == false)   */
  FSR1_FLAG_PRECISE = 1u << 5,      /* fp16 images: fp32 arithmetic on fp16 storage where a tiled kernel exists for it
                                       (EASU at exactly 2x: packed FFMA2 taps); ~10x closer to the fp32 algorithm */
  FSR1_FLAG_RCAS_DENOISE = 1u << 6, /* the reference's FSR_RCAS_DENOISE compile-time option (ffx_fsr1.h:651,731-763)      */
  FSR1_FLAG_RCAS_PASSTHROUGH_ALPHA = 1u << 7, /* FSR_RCAS_PASSTHROUGH_ALPHA (:648,688-702): output alpha = input alpha */
  FSR1_FLAG_OUTPUT_SQUARE = 1u << 8, /* the sample's Sample.x hook (sample/src/DX12/FSR_Pass.hlsl:78-79,93-94): `c *= c` on the
                                       output of the LAST pass (gamma 2.0, as produced by TEPD, back to linear)      */
  FSR1_FLAG_FUSED = 1u << 9,        /* fsr1_upscale*: EASU and RCAS in ONE kernel where one exists (RGBA16F, exactly 2x, out-of-image
                                       RCAS taps read 0, no RCAS options): the intermediate stays in shared memory, `tmp` is not
                                       touched, HBM traffic drops from 26 to 10 bytes per output pixel; results are bit-identical
                                       to the two-kernel path.  Falls back to the two kernels otherwise. */
  FSR1_FLAG_RCAS_HX2 = 1u << 10,    /* fsr1_rcas, fp16 images only: the reference's PACKED calling convention FsrRcasHx2 +
                                       FsrRcasDepackHx2 (ffx_fsr1.h:880-984): each lane sharpens pixels ip and ip + (8,0) held as
                                       half2 structure-of-arrays registers, every operation a packed half operation with its own
                                       rounding.  Bit-identical to FSR1_FLAG_H_REFERENCE (the reference's Hx2 and H sources agree
                                       bit for bit); honours RCAS_CLAMP, RCAS_DENOISE, RCAS_PASSTHROUGH_ALPHA.  A parity path. */
  FSR1_FLAG_H_REFERENCE = 1u << 4   /* fp16 images only: the literal FsrEasuH / FsrRcasH arithmetic (packed-half
                                       algorithm, half magic numbers, per-operation half rounding), bit-identical
                                       to the reference's H source; a parity path, slower and LESS accurate than
                                       the default fp16 kernels (see DESIGN.md "numerics")                  */
};

/* A (window of a) device image.  `width`/`height` are the logical size of the whole image; `data`
 * points at logical row `row0` and holds `rows` rows (row0 = 0, rows = height for a whole image).
 * Windows exist for row-slab sharding: a GPU holds only the rows it needs (plus halo) but clamping and
 * out-of-image rules still refer to the whole image.  pitch_bytes >= width * bytes-per-pixel. */
typedef struct fsr1_image {
  void* data;
  uint64_t pitch_bytes;
  uint32_t width, height;
  uint32_t row0, rows;
  uint32_t format;
  uint32_t reserved;
} fsr1_image;

/* EASU over output rows [y0, y1) (y1 == 0 means "to the last row").  con = con0..con3, 16 words. */
int fsr1_easu(const fsr1_image* in, const fsr1_image* out, const uint32_t con[16], uint32_t y0, uint32_t y1,
              uint32_t flags, void* stream);

/* RCAS over rows [y0, y1); in and out have the same logical size and format.  con = 4 words. */
int fsr1_rcas(const fsr1_image* in, const fsr1_image* out, const uint32_t con[4], uint32_t y0, uint32_t y1,
              uint32_t flags, void* stream);

/* First and last input row EASU reads to produce output rows [y0,y1) (clamped to the image): what a
 * slab must hold, and what must be exchanged as halo when the output is sharded by rows. */
int fsr1_easu_input_rows(const uint32_t con[16], uint32_t in_height, uint32_t y0, uint32_t y1,
                         uint32_t* first_row, uint32_t* last_row);

/* EASU -> RCAS for output rows [y0,y1).  `tmp` is the display-sized intermediate (same format as out);
 * it must hold rows [y0-1, y1+1) clipped to the image.  easu_con 16 words, rcas_con 4 words. */
int fsr1_upscale(const fsr1_image* in, const fsr1_image* tmp, const fsr1_image* out, const uint32_t easu_con[16],
                 const uint32_t rcas_con[4], uint32_t y0, uint32_t y1, uint32_t flags, void* stream);

/* ---- resource-owning context (the FSR_Filter object of the sample) ---------------------------- */
typedef struct fsr1_context fsr1_context;

/* Allocates the intermediate image for (out_width x out_height, format) on the current device. */
int fsr1_context_create(fsr1_context** ctx, uint32_t in_width, uint32_t in_height, uint32_t out_width,
                        uint32_t out_height, uint32_t format);
void fsr1_context_destroy(fsr1_context* ctx);

/* Device-resident frames: constants are derived inside exactly as FSR_Filter::Upscale does
 * (FsrEasuCon(renderW,renderH,renderW,renderH,displayW,displayH); FsrRcasCon(sharpness_stops)). */
int fsr1_context_upscale(fsr1_context* ctx, const void* in_dev, uint64_t in_pitch, void* out_dev, uint64_t out_pitch,
                         float sharpness_stops, uint32_t flags, void* stream);

/* The same with the render size of THIS frame (dynamic resolution / a preset change without re-creating the resources):
 * FSR_Filter::Upscale rebuilds FsrEasuCon from pState->renderWidth/renderHeight on every call (FSR_Filter.cpp:106). */
int fsr1_context_upscale_render(fsr1_context* ctx, const void* in_dev, uint64_t in_pitch, uint32_t render_width,
                                uint32_t render_height, void* out_dev, uint64_t out_pitch, float sharpness_stops,
                                uint32_t flags, void* stream);

/* Host-resident frames (pinned memory recommended): H2D copy, EASU, RCAS, D2H copy on `stream`. */
int fsr1_context_upscale_host(fsr1_context* ctx, const void* in_host, uint64_t in_pitch, void* out_host,
                              uint64_t out_pitch, float sharpness_stops, uint32_t flags, void* stream);

/* ---- row-slab sharding across the GPUs of one box (the reference has no multi-GPU path: new) ------
 * The output image is cut into `world` contiguous row slabs, one per rank (= per GPU).  Rank k owns input rows
 * [k*in_h/world, (k+1)*in_h/world) and needs 2-3 more rows each side: the EASU footprint of its slab plus the
 * one-row apron RCAS reads, so there is exactly ONE neighbour exchange per frame and no collective.
 * A fsr1_shard owns the rank's share of a ring of `slots` frames: input window (own rows + halo), intermediate,
 * output slab, three internal streams.  The halo moves by direct NVLink stores into the neighbour's window
 * (CUDA IPC between processes, peer access inside one process), flow-controlled by sequence numbers in device
 * memory: no NCCL call, no host synchronisation and no allocation per frame.  Per frame the caller writes its
 * input rows into fsr1_shard_input(slot) on `stream`, calls fsr1_shard_submit(slot, stream), and orders its
 * consumer after fsr1_shard_wait(slot, stream).  RCAS of frame i overlaps EASU of frame i+1 on every rank.
 * All ranks must create shards with identical arguments (except rank) and submit slots in the same order.
 * Set-up between processes: every rank calls fsr1_shard_export, the 64-byte handles are gathered in rank order by
 * any means (torch.distributed all_gather, MPI, a pipe), every rank calls fsr1_shard_attach.  In one process
 * driving several devices: fsr1_shard_attach_local(shard, shard_of_rank-1, shard_of_rank+1). */
typedef struct fsr1_shard fsr1_shard;
#define FSR1_SHARD_HANDLE_BYTES 64
#define FSR1_SHARD_ONE_STREAM (1u << 16) /* fsr1_shard_create flag: every frame on one stream (no overlap of consecutive frames; default: two
                                            streams, whole frames in turn, so RCAS of frame i overlaps EASU of frame i+1)                  */
#define FSR1_SHARD_TRACE (1u << 18)      /* keep device timestamps of the last 256 frames (fsr1_shard_trace)                          */
#define FSR1_SHARD_SKIP_HALO (1u << 17)  /* MEASUREMENT ONLY: no halo exchange (slab borders are wrong); times the frame without it */

typedef struct fsr1_shard_info {   /* logical row ranges [row0, row1) of this rank */
  uint32_t out_row0, out_row1;       /* output slab                                                 */
  uint32_t easu_row0, easu_row1;     /* rows EASU produces (slab + RCAS apron)                      */
  uint32_t owned_row0, owned_row1;   /* input rows this rank owns (the caller writes them)          */
  uint32_t needed_row0, needed_row1; /* input rows EASU reads                                       */
  uint32_t window_row0, window_row1; /* input rows resident on this rank (owned + halo)             */
  uint32_t send_up_row0, send_up_row1, send_down_row0, send_down_row1; /* rows pushed to rank-1 / rank+1 */
  uint64_t halo_recv_bytes;          /* halo payload received per frame                             */
  uint64_t arena_bytes;
} fsr1_shard_info;

/* `flags`: FSR1_FLAG_* for the kernels, optionally FSR1_SHARD_ONE_STREAM.  FSR1_ERR_UNSUPPORTED when a slab is
 * shorter than the halo it must supply (the halo would come from beyond the direct neighbours). */
int fsr1_shard_create(fsr1_shard** shard, uint32_t in_width, uint32_t in_height, uint32_t out_width, uint32_t out_height,
                      uint32_t format, uint32_t world, uint32_t rank, uint32_t slots, float sharpness_stops, uint32_t flags);
void fsr1_shard_destroy(fsr1_shard* shard);
int fsr1_shard_geometry(const fsr1_shard* shard, fsr1_shard_info* info);
int fsr1_shard_export(const fsr1_shard* shard, void* handle /* FSR1_SHARD_HANDLE_BYTES */);
int fsr1_shard_attach(fsr1_shard* shard, const void* handles /* world x FSR1_SHARD_HANDLE_BYTES, rank order */, uint32_t count);
int fsr1_shard_attach_local(fsr1_shard* shard, fsr1_shard* up /* rank-1 or NULL */, fsr1_shard* down /* rank+1 or NULL */);
int fsr1_shard_input(const fsr1_shard* shard, uint32_t slot, fsr1_image* owned);   /* where the caller writes its rows */
int fsr1_shard_window(const fsr1_shard* shard, uint32_t slot, fsr1_image* window); /* owned rows + halo (read-only)     */
int fsr1_shard_output(const fsr1_shard* shard, uint32_t slot, fsr1_image* out);    /* the rank's output slab            */
void* fsr1_shard_arena(const fsr1_shard* shard);
/* Upscale the frame in `slot`: ordered after everything already on `stream`; returns without waiting. */
int fsr1_shard_submit(fsr1_shard* shard, uint32_t slot, void* stream);
/* Orders `stream` after the slot's result (and after this rank's halo rows have left: the input may be rewritten). */
int fsr1_shard_wait(fsr1_shard* shard, uint32_t slot, void* stream);
/* FSR1_SHARD_TRACE: 8 GPU globaltimer stamps (ns) per frame, oldest first: [0] EASU began waiting for its halo, [1] halo present,
 * [2] last EASU CTA done (credit sent), [3]/[5] push up/down started, [4]/[6] push up/down published, [7] unused. */
int fsr1_shard_trace(fsr1_shard* shard, uint64_t* out /* max_frames x 8 */, uint32_t max_frames, uint32_t* n_frames);
int fsr1_shard_status(fsr1_shard* shard);  /* FSR1_OK, or FSR1_ERR_TIMEOUT if a neighbour never answered (call after a sync) */

/* ---- pointwise companions of the scaling path (ffx-fsr/ffx_fsr1.h:986-1199) ----------------------
 * The passes the sample runs either side of EASU/RCAS, as whole-image streaming kernels over rows [y0,y1)
 * (y1 == 0: to the last row).  `in` and `out` have the same logical size and may be the same image (in place).
 * Arithmetic is fp32 with separate roundings for every storage format: RGBA32F results are bit-identical to the
 * reference's F functions, other formats round that result once on store.  Alpha is carried through.
 *   fsr1_srtm   FsrSrtmF (inverse == 0) / FsrSrtmInvF (inverse != 0)   ffx_fsr1.h:1044,1046
 *   fsr1_lfga   FsrLfgaF: c += (grain * amount) * min(1 - c, c)         ffx_fsr1.h:1014
 *               `grain` = RGB image of {-0.5..0.5} values (float formats), tiled over the frame with wrap addressing
 *   fsr1_tepd   FsrTepdC8F (bits == 8) / FsrTepdC10F (bits == 10)       ffx_fsr1.h:1100-1126
 *               dither == NULL: FsrTepdDitF(pixel, frame) (ffx_fsr1.h:1086-1095); else the saturated .w channel of the
 *               tiled `dither` image (sample/src/DX12/FSR_Tonemapping.hlsl:87).  `out` may be the same format as `in`
 *               or, from a float image, RGBA8_UNORM (bits 8) / RGB10A2_UNORM (bits 10): the code values themselves. */
int fsr1_srtm(const fsr1_image* in, const fsr1_image* out, int inverse, uint32_t y0, uint32_t y1, void* stream);
int fsr1_lfga(const fsr1_image* in, const fsr1_image* grain, const fsr1_image* out, float amount, uint32_t y0,
              uint32_t y1, void* stream);
int fsr1_tepd(const fsr1_image* in, const fsr1_image* dither, const fsr1_image* out, int bits, uint32_t frame,
              uint32_t y0, uint32_t y1, void* stream);

/* The same passes in the reference's HALF arithmetic, through its packed calling convention (two pixels, p and p + (8,0), per lane
 * in half2 registers): RGBA16F images only (in, out, grain, dither), every operation rounded to half once, results bit-identical
 * to the reference's H and Hx2 functions (which agree with each other bit for bit).  Same arguments and rules as above.
 *   fsr1_srtm_h   FsrSrtmH / FsrSrtmHx2 (inverse == 0), FsrSrtmInvH / FsrSrtmInvHx2     ffx_fsr1.h:1049-1055
 *   fsr1_lfga_h   FsrLfgaH / FsrLfgaHx2; `amount` is converted to half once               ffx_fsr1.h:1019-1024
 *   fsr1_tepd_h   FsrTepdC8H / C8Hx2 (bits == 8), FsrTepdC10H / C10Hx2 (bits == 10)      ffx_fsr1.h:1137-1153,1166-1199
 *                 dither == NULL: FsrTepdDitH / FsrTepdDitHx2(pixel, frame) (:1129-1135,1156-1164); `out` is RGBA16F */
int fsr1_srtm_h(const fsr1_image* in, const fsr1_image* out, int inverse, uint32_t y0, uint32_t y1, void* stream);
int fsr1_lfga_h(const fsr1_image* in, const fsr1_image* grain, const fsr1_image* out, float amount, uint32_t y0,
                uint32_t y1, void* stream);
int fsr1_tepd_h(const fsr1_image* in, const fsr1_image* dither, const fsr1_image* out, int bits, uint32_t frame,
                uint32_t y0, uint32_t y1, void* stream);

/* ---- constants through the ABI (for FFIs that cannot include fsr1_host.h) ---------------------- */
void fsr1_easu_con(uint32_t con[16], float in_viewport_w, float in_viewport_h, float in_size_w, float in_size_h,
                   float out_w, float out_h);
void fsr1_easu_con_offset(uint32_t con[16], float in_viewport_w, float in_viewport_h, float in_size_w,
                          float in_size_h, float out_w, float out_h, float in_off_x, float in_off_y);
void fsr1_rcas_con(uint32_t con[4], float sharpness_stops);

/* ---- introspection ----------------------------------------------------------------------------- */
int fsr1_abi_version(void);
const char* fsr1_error_string(int err);
int fsr1_last_cuda_error(void);          /* cudaError_t of the last failed CUDA call on this thread   */
uint64_t fsr1_launch_count(void);        /* kernels launched by this library since load (all threads) */
const char* fsr1_last_kernel_name(void); /* which kernel variant the last launch on this thread used  */

#ifdef __cplusplus
}
#endif
#endif /* FSR1_B200_H */
