// micro_pipes.cu — measures issue rates that the Harris/Canny kernel designs depend on:
// FFMA (3-reg), FFMA with a constant-bank operand, packed FFMA2 (fma.rn.f32x2), FADD+FFMA mix,
// DFMA, and un-fused DMUL+DADD.  Prints G lane-ops/s per variant.
#include <cstdio>
#include <cuda_runtime.h>

struct W { float w[8]; };

template <int MODE>
__global__ void k(float *out, float a, float b, double da, double db, const __grid_constant__ W cw, int iters) {
  float x[8]; float2 p[4]; double d[8];
  for (int i = 0; i < 8; i++) { x[i] = threadIdx.x * 0.001f + i; d[i] = x[i]; }
  for (int i = 0; i < 4; i++) p[i] = make_float2(x[2 * i], x[2 * i + 1]);
  float2 aa = make_float2(a, a), bb = make_float2(b, b);
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 8; i++) x[i] = fmaf(x[i], a, b);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; i++) x[i] = fmaf(x[i], cw.w[i], b);
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 4; i++) p[i] = __ffma2_rn(p[i], aa, bb);
    } else if (MODE == 3) {   // symmetric-tap pattern: add then fma
#pragma unroll
      for (int i = 0; i < 4; i++) { float s = x[i] + x[i + 4]; x[i] = fmaf(s, a, x[i]); x[i + 4] = fmaf(s, b, x[i + 4]); }
    } else if (MODE == 4) {
#pragma unroll
      for (int i = 0; i < 8; i++) d[i] = fma(d[i], da, db);
    } else if (MODE == 5) {
#pragma unroll
      for (int i = 0; i < 8; i++) d[i] = __dadd_rn(__dmul_rn(d[i], da), db);
    } else if (MODE == 6) {   // packed add + packed fma (symmetric taps, packed)
#pragma unroll
      for (int i = 0; i < 2; i++) { float2 s = __fadd2_rn(p[i], p[i + 2]); p[i] = __ffma2_rn(s, aa, p[i]); p[i + 2] = __ffma2_rn(s, bb, p[i + 2]); }
    }
  }
  float r = 0;
  for (int i = 0; i < 8; i++) r += x[i] + (float)d[i];
  for (int i = 0; i < 4; i++) r += p[i].x + p[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE> void run(const char *name, double ops_per_iter) {
  int blocks = 148 * 8, threads = 256, iters = 4096;
  float *out; cudaMalloc(&out, blocks * threads * 4);
  W cw; for (int i = 0; i < 8; i++) cw.w[i] = 0.999f + i * 1e-4f;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(out, 0.9999f, 1e-3f, 0.9999, 1e-3, cw, 16);
  cudaDeviceSynchronize();
  float best = 1e9;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, 0.9999f, 1e-3f, 0.9999, 1e-3, cw, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  double ops = (double)blocks * threads * iters * ops_per_iter;
  printf("%-28s %8.3f ms  %9.1f G lane-ops/s\n", name, best, ops / best / 1e6);
  cudaFree(out);
}

int main() {
  run<0>("FFMA reg,reg,reg", 8);
  run<1>("FFMA reg,const,reg", 8);
  run<2>("FFMA2 packed (8 lanes)", 8);
  run<3>("FADD+2xFFMA mix (12 ops)", 12);
  run<6>("FADD2+2xFFMA2 mix (12 ops)", 12);
  run<4>("DFMA", 8);
  run<5>("DMUL+DADD (16 ops)", 16);
  return 0;
}
