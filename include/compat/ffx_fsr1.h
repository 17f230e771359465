/* Drop-in shim: FsrEasuCon / FsrEasuConOffset / FsrRcasCon with the reference's signatures (see ../fsr1_host.h).
 * The device entry points FsrEasuF/H and FsrRcasF/H are replaced by fsr1_easu / fsr1_rcas of ../fsr1_b200.h. */
#ifndef FSR1_COMPAT_FFX_FSR1_H
#define FSR1_COMPAT_FFX_FSR1_H
#include "../fsr1_host.h"
#define FSR_RCAS_LIMIT (0.25 - (1.0 / 16.0))
#endif
