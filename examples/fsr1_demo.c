/* fsr1_demo.c — the C ABI from plain C: what a host application written against the reference's headers does after
 * switching.  Constants come from FsrEasuCon / FsrRcasCon exactly as in the reference's FSR_Filter::Upscale
 * (sample/src/DX12/FSR_Filter.cpp:101-141); the two dispatches become one fsr1_upscale call on device images.
 *
 *   gcc -std=c99 -Iinclude examples/fsr1_demo.c -o fsr1_demo -Lfidelityfx-fsr_b200/lib -lfsr1_b200 \
 *       -L/usr/local/cuda/lib64 -lcudart -lm
 *   ./fsr1_demo            (prints a checksum of the upscaled frame; needs a GPU)
 */
#include <stdio.h>
#include <stdlib.h>
#include <cuda_runtime_api.h>

#define A_CPU 1
#include "compat/ffx_a.h"   /* the reference's include lines, served by include/compat -> fsr1_host.h */
#include "compat/ffx_fsr1.h"
#include "fsr1_b200.h"

int main(void) {
  const uint32_t renderW = 640, renderH = 360, displayW = 1280, displayH = 720;
  AU1 con0[4], con1[4], con2[4], con3[4], rcas[4], easu[16];
  FsrEasuCon(con0, con1, con2, con3, (AF1)renderW, (AF1)renderH, (AF1)renderW, (AF1)renderH, (AF1)displayW, (AF1)displayH);
  FsrRcasCon(rcas, (AF1)0.25);
  for (int i = 0; i < 4; i++) { easu[i] = con0[i]; easu[4 + i] = con1[i]; easu[8 + i] = con2[i]; easu[12 + i] = con3[i]; }

  /* an RGBA32F gradient frame on the host */
  size_t inBytes = (size_t)renderW * renderH * 16, outBytes = (size_t)displayW * displayH * 16;
  float* h = (float*)malloc(inBytes);
  for (uint32_t y = 0; y < renderH; y++)
    for (uint32_t x = 0; x < renderW; x++) {
      float* p = h + ((size_t)y * renderW + x) * 4;
      p[0] = (float)x / renderW; p[1] = (float)y / renderH; p[2] = ((x / 16 + y / 16) & 1) ? 1.0f : 0.0f; p[3] = 1.0f;
    }
  void *dIn, *dTmp, *dOut;
  if (cudaMalloc(&dIn, inBytes) || cudaMalloc(&dTmp, outBytes) || cudaMalloc(&dOut, outBytes)) { fprintf(stderr, "no CUDA device\n"); return 2; }
  cudaMemcpy(dIn, h, inBytes, cudaMemcpyHostToDevice);
  fsr1_image in = {dIn, renderW * 16ull, renderW, renderH, 0, renderH, FSR1_FORMAT_RGBA32F, 0};
  fsr1_image tmp = {dTmp, displayW * 16ull, displayW, displayH, 0, displayH, FSR1_FORMAT_RGBA32F, 0};
  fsr1_image out = {dOut, displayW * 16ull, displayW, displayH, 0, displayH, FSR1_FORMAT_RGBA32F, 0};
  int rc = fsr1_upscale(&in, &tmp, &out, easu, rcas, 0, 0, 0, NULL);
  if (rc != FSR1_OK) { fprintf(stderr, "fsr1_upscale: %s\n", fsr1_error_string(rc)); return 1; }
  float* r = (float*)malloc(outBytes);
  cudaMemcpy(r, dOut, outBytes, cudaMemcpyDeviceToHost);   /* also synchronises */
  double sum = 0.0;
  for (size_t i = 0; i < (size_t)displayW * displayH; i++) sum += r[i * 4] + r[i * 4 + 1] + r[i * 4 + 2];
  printf("upscaled %ux%u -> %ux%u with %s, checksum %.3f, alpha %.1f\n", renderW, renderH, displayW, displayH,
         fsr1_last_kernel_name(), sum, r[3]);
  cudaFree(dIn); cudaFree(dTmp); cudaFree(dOut); free(h); free(r);
  return 0;
}
