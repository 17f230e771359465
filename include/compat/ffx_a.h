/* Drop-in shim: code written against the reference's `#define A_CPU` + `#include "ffx_a.h"` keeps compiling.
 * The host-side type/math surface the FSR1 constant setup needs lives in ../fsr1_host.h. */
#ifndef FSR1_COMPAT_FFX_A_H
#define FSR1_COMPAT_FFX_A_H
#ifndef A_CPU
#error "this build of ffx_a.h provides the host (A_CPU) surface only; the per-pixel passes are CUDA kernels behind fsr1_b200.h"
#endif
#include "../fsr1_host.h"
#endif
