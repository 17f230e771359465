/* TEST INFRASTRUCTURE — not product code.
 *
 * The reference's constant-setup functions compiled from its UNMODIFIED headers with
 * `#define A_CPU` (the only thing A_CPU provides):
 *   FsrEasuCon        ffx-fsr/ffx_fsr1.h:156-202
 *   FsrEasuConOffset  ffx-fsr/ffx_fsr1.h:205-225
 *   FsrRcasCon        ffx-fsr/ffx_fsr1.h:662-672   (packed half via the truncating table packer,
 *                                                    ffx-fsr/ffx_a.h:482-552)
 * The headers are read from /root/reference through -I at build time; nothing is copied.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#define A_CPU 1
#include "ffx_a.h"
#include "ffx_fsr1.h"

void fsr1ref_cpu_easu_con(uint32_t* con, float inVpW, float inVpH, float inW, float inH, float outW, float outH) {
  FsrEasuCon(con, con + 4, con + 8, con + 12, inVpW, inVpH, inW, inH, outW, outH);
}
void fsr1ref_cpu_easu_con_offset(uint32_t* con, float inVpW, float inVpH, float inW, float inH, float outW,
                                 float outH, float offX, float offY) {
  FsrEasuConOffset(con, con + 4, con + 8, con + 12, inVpW, inVpH, inW, inH, outW, outH, offX, offY);
}
void fsr1ref_cpu_rcas_con(uint32_t* con, float sharpness) { FsrRcasCon(con, sharpness); }
uint32_t fsr1ref_cpu_f32_to_f16(float f) { return AU1_AH1_AF1(f); }
