/* fsr1_host.h — host-side constant setup of the FSR 1.0 hot path, call-compatible with the reference.
 *
 * The reference is header-only; its host entry points exist when the application does
 *     #define A_CPU
 *     #include "ffx_a.h"
 *     #include "ffx_fsr1.h"
 * and are then  FsrEasuCon        (reference ffx-fsr/ffx_fsr1.h:156-202)
 *               FsrEasuConOffset  (reference ffx-fsr/ffx_fsr1.h:205-225)
 *               FsrRcasCon        (reference ffx-fsr/ffx_fsr1.h:662-672)
 * over the scalar types AF1/AU1/AP1 ... (reference ffx-fsr/ffx_a.h:121-131) and the helpers
 * AU1_AF1, ARcpF1, AExp2F1, AU1_AH1_AF1, AU1_AH2_AF2 (ffx_a.h:141,326,283-286,482-552).
 * This header supplies the same names with the same signatures, argument meaning and bit-identical
 * results (tests/test_constants.py checks them against the reference's unmodified header).
 * Like the reference it is plain C (C99) or C++, needs <math.h>, validates nothing and returns void.
 *
 * Written from the published formulas; the float->half packer is arithmetic instead of the
 * reference's two 512-entry tables but produces the same bits for every one of the 2^32 inputs
 * (truncating mantissa, subnormals kept, +-inf/NaN -> +-65504).
 */
#ifndef FSR1_HOST_H
#define FSR1_HOST_H
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifndef A_STATIC
#define A_STATIC static
#endif
#ifndef A_RESTRICT
#define A_RESTRICT __restrict
#endif

#ifndef FSR1_HOST_NO_TYPES
typedef uint32_t AP1;
typedef float AF1;
typedef double AD1;
typedef uint8_t AB1;
typedef uint16_t AW1;
typedef uint32_t AU1;
typedef uint64_t AL1;
typedef int8_t ASB1;
typedef int16_t ASW1;
typedef int32_t ASU1;
typedef int64_t ASL1;
#define AF1_(a) ((AF1)(a))
#define AU1_(a) ((AU1)(a))
#define A_TRUE 1
#define A_FALSE 0
/* vector arguments are restrict pointers on the CPU, exactly as in the reference's porting layer */
#define outAU4 AU1* A_RESTRICT
#define inAF2 AF1* A_RESTRICT
#define varAF2(x) AF1 x[2]
#define initAF2(x, y) {x, y}
#endif

A_STATIC AU1 AU1_AF1(AF1 a) { AU1 u; memcpy(&u, &a, sizeof u); return u; }
A_STATIC AF1 AF1_AU1(AU1 a) { AF1 f; memcpy(&f, &a, sizeof f); return f; }
A_STATIC AF1 ARcpF1(AF1 a) { return 1.0f / a; }
A_STATIC AF1 AExp2F1(AF1 a) { return exp2f(a); }

/* float -> half in the low 16 bits: mantissa truncated, half subnormals kept, overflow/inf/NaN -> 0x7bff. */
A_STATIC AU1 AU1_AH1_AF1(AF1 f) {
  AU1 u = AU1_AF1(f), s = (u >> 16) & 0x8000u, e = (u >> 23) & 0xffu, m = u & 0x7fffffu;
  if (e < 103u) return s;
  if (e < 113u) return s + (0x0400u >> (113u - e)) + (m >> (126u - e));
  if (e < 143u) return s + ((e - 112u) << 10) + (m >> 13);
  return s + 0x7bffu;
}
A_STATIC AU1 AU1_AH2_AF2(inAF2 a) { return AU1_AH1_AF1(a[0]) + (AU1_AH1_AF1(a[1]) << 16); }

/* con0 = {inVp/out (x,y), 0.5*inVp/out-0.5 (x,y)}   output pixel -> input position of tap 'f'
 * con1 = {1/inW, 1/inH, 1/inW, -1/inH}  con2 = {-1/inW, 2/inH, 1/inW, 2/inH}  con3 = {0, 4/inH, 0, 0}
 *        (gather4 centres in normalised coordinates; our kernels address texels directly and use
 *         only con0, but all 16 words are produced so the block is interchangeable) */
A_STATIC void FsrEasuCon(outAU4 con0, outAU4 con1, outAU4 con2, outAU4 con3,
                         AF1 inputViewportInPixelsX, AF1 inputViewportInPixelsY,
                         AF1 inputSizeInPixelsX, AF1 inputSizeInPixelsY,
                         AF1 outputSizeInPixelsX, AF1 outputSizeInPixelsY) {
  con0[0] = AU1_AF1(inputViewportInPixelsX * ARcpF1(outputSizeInPixelsX));
  con0[1] = AU1_AF1(inputViewportInPixelsY * ARcpF1(outputSizeInPixelsY));
  con0[2] = AU1_AF1(AF1_(0.5) * inputViewportInPixelsX * ARcpF1(outputSizeInPixelsX) - AF1_(0.5));
  con0[3] = AU1_AF1(AF1_(0.5) * inputViewportInPixelsY * ARcpF1(outputSizeInPixelsY) - AF1_(0.5));
  con1[0] = AU1_AF1(ARcpF1(inputSizeInPixelsX));
  con1[1] = AU1_AF1(ARcpF1(inputSizeInPixelsY));
  con1[2] = AU1_AF1(AF1_(1.0) * ARcpF1(inputSizeInPixelsX));
  con1[3] = AU1_AF1(AF1_(-1.0) * ARcpF1(inputSizeInPixelsY));
  con2[0] = AU1_AF1(AF1_(-1.0) * ARcpF1(inputSizeInPixelsX));
  con2[1] = AU1_AF1(AF1_(2.0) * ARcpF1(inputSizeInPixelsY));
  con2[2] = AU1_AF1(AF1_(1.0) * ARcpF1(inputSizeInPixelsX));
  con2[3] = AU1_AF1(AF1_(2.0) * ARcpF1(inputSizeInPixelsY));
  con3[0] = AU1_AF1(AF1_(0.0) * ARcpF1(inputSizeInPixelsX));
  con3[1] = AU1_AF1(AF1_(4.0) * ARcpF1(inputSizeInPixelsY));
  con3[2] = con3[3] = 0;
}

A_STATIC void FsrEasuConOffset(outAU4 con0, outAU4 con1, outAU4 con2, outAU4 con3,
                               AF1 inputViewportInPixelsX, AF1 inputViewportInPixelsY,
                               AF1 inputSizeInPixelsX, AF1 inputSizeInPixelsY,
                               AF1 outputSizeInPixelsX, AF1 outputSizeInPixelsY,
                               AF1 inputOffsetInPixelsX, AF1 inputOffsetInPixelsY) {
  FsrEasuCon(con0, con1, con2, con3, inputViewportInPixelsX, inputViewportInPixelsY, inputSizeInPixelsX,
             inputSizeInPixelsY, outputSizeInPixelsX, outputSizeInPixelsY);
  con0[2] = AU1_AF1(AF1_(0.5) * inputViewportInPixelsX * ARcpF1(outputSizeInPixelsX) - AF1_(0.5) + inputOffsetInPixelsX);
  con0[3] = AU1_AF1(AF1_(0.5) * inputViewportInPixelsY * ARcpF1(outputSizeInPixelsY) - AF1_(0.5) + inputOffsetInPixelsY);
}

/* sharpness is in stops: 0 = maximum, N halves it N times.  con = {bits(s), half2(s,s), 0, 0}. */
A_STATIC void FsrRcasCon(outAU4 con, AF1 sharpness) {
  sharpness = AExp2F1(-sharpness);
  varAF2(hSharp) = initAF2(sharpness, sharpness);
  con[0] = AU1_AF1(sharpness);
  con[1] = AU1_AH2_AF2(hSharp);
  con[2] = 0;
  con[3] = 0;
}

#endif /* FSR1_HOST_H */
