// harris_kernels.cuh — device code for the Harris corner path.
//
// Two implementations of the response map R = measure(G_si * (grad(G_sd * I) grad^T)):
//
//  (1) harris_fused_kernel  — the production path.  ONE kernel, one HBM read of the frame
//      (u8 or f32) and one HBM write of R.  A CTA owns a TWxTH output tile and walks the
//      whole chain in shared memory:  tile(+11 halo) -> row blur sd -> column blur sd ->
//      gradient+products+row blur si (fused, products never touch memory) -> column blur si
//      + corner measure -> coalesced store.  fp32 FMA arithmetic with register-blocked
//      sliding windows (4 outputs / thread in x, 5-8 in y).  Reference semantics kept exactly:
//      asymmetric reflect padding of every blur (gaussian.cpp:345-349,376-380) and the
//      replicate rule of the gradient (gradient.cpp:40-55) — realised by reflecting the input
//      tile at load time and by remapping product coordinates in border tiles.
//      Follows image.CornerDetectionHarris/src/harris.cpp:511-520 (stages 1-4 of harris()).
//
//  (2) exact_* kernels — one kernel per reference stage, double accumulation in the
//      reference's own operation order without FMA contraction (gaussian.cpp:353-358), the
//      sequential float prefix sums of the SII mode (gaussian.cpp:179-215) and correctly
//      rounded float products.  Bit-identical to the reference's R; used for params.exact=1,
//      for the SII / no-Gaussian modes, for sigmas the fused kernel is not instantiated for,
//      and for frames too small to tile.
#pragma once
#include "common.cuh"

namespace b2f {

constexpr int HARRIS_MAX_TAPS = 32;

struct HarrisConsts {
  float wd[HARRIS_MAX_TAPS];    // sigma_d taps, wd[0] centre (float-rounded reference weights)
  float wir[HARRIS_MAX_TAPS];   // sigma_i taps of the ROW pass: x0.25 for central differences (the
                                // two 0.5 factors of gradient.cpp:35-36 commute exactly with the
                                // float roundings), x1 for Sobel
  float wic[HARRIS_MAX_TAPS];   // sigma_i taps of the COLUMN pass (unscaled)
  float k;
  int measure;   // 0 Harris, 1 Shi-Tomasi, 2 harmonic mean
};

// Corner measure from the smoothed structure tensor, every float operation rounded separately
// and in the reference's order (harris.cpp:100-103, :113-116, :126-129) — no FMA contraction.
__device__ __forceinline__ float corner_measure(float A, float B, float C, float k, int measure) {
  if (measure == 1) {
    float s = __fmul_rn(A, A);
    s = __fsub_rn(s, __fmul_rn(__fmul_rn(2.f, A), C));
    s = __fadd_rn(s, __fmul_rn(__fmul_rn(4.f, B), B));
    s = __fadd_rn(s, __fmul_rn(C, C));
    float D = __fsqrt_rn(s);
    return __fsub_rn(__fmul_rn(0.5f, __fadd_rn(A, C)), __fmul_rn(0.5f, D));
  }
  float det = __fsub_rn(__fmul_rn(A, C), __fmul_rn(B, B));
  float tr = __fadd_rn(A, C);
  if (measure == 2) return (float)__ddiv_rn((double)__fmul_rn(2.f, det), __dadd_rn((double)tr, 0.0001));
  return __fsub_rn(det, __fmul_rn(__fmul_rn(k, tr), tr));
}

// ------------------------------------------------------------------------------------------
// fused kernel
// ------------------------------------------------------------------------------------------
template <int RD, int RI> struct FusedCfg {
  static constexpr int TW = 64, TH = 64, NT = 256;
  static constexpr int H = RD + 1 + RI;             // total halo
  static constexpr int G = RI + 1;                  // halo of the blurred image Is
  static constexpr int IN_W = TW + 2 * H, IN_H = TH + 2 * H;
  static constexpr int IN_P = (IN_W + 3) & ~3;      // pitch (floats), multiple of 4
  static constexpr int R1_W = TW + 2 * G, R1_H = IN_H, R1_P = (R1_W + 3) & ~3;
  static constexpr int IS_W = R1_W, IS_H = TH + 2 * G, IS_P = R1_P;
  static constexpr int AR_H = TH + 2 * RI, AR_P = TW;
  static constexpr int REGION_X = (IN_H * IN_P > IS_H * IS_P) ? IN_H * IN_P : IS_H * IS_P;
  static constexpr int REGION_Y = (R1_H * R1_P > 3 * AR_H * AR_P) ? R1_H * R1_P : 3 * AR_H * AR_P;
  static constexpr int MAP_N = 2 * (TW + 2 * RI) + 2 * (TH + 2 * RI);
  static constexpr size_t SMEM = sizeof(float) * (REGION_X + REGION_Y) + sizeof(short) * MAP_N;
};

__device__ __forceinline__ int reflect_index(int p, int n) {
  // padding rule of discrete_gaussian: -k -> k ; n-1+k -> n-k   (gaussian.cpp:345-349)
  if (p < 0) p = -p;
  else if (p >= n) p = 2 * n - 1 - p;
  return min(max(p, 0), n - 1);
}

// tiles whose +-12 halo lies inside the frame are taken by harris_fused2_kernel (harris_kernels2.cuh)
__device__ __forceinline__ bool harris_tile_is_interior(int x0, int y0, int nx, int ny) {
  return x0 >= 12 && y0 >= 12 && x0 + 64 + 12 <= nx && y0 + 64 + 12 <= ny;
}

template <int RD, int RI, bool U8, int GRAD>
__global__ void __launch_bounds__(256, 2)
harris_fused_kernel(const void *__restrict__ frames, float *__restrict__ Rout, int nx, int ny,
                    const __grid_constant__ HarrisConsts kc, int border_ring_only) {
  using C = FusedCfg<RD, RI>;
  if (border_ring_only && harris_tile_is_interior(blockIdx.x * C::TW, blockIdx.y * C::TH, nx, ny)) return;
  extern __shared__ __align__(16) float smem[];
  float *sIN = smem;                    // region X (input tile, later Is)
  float *sIS = smem;
  float *sR1 = smem + C::REGION_X;      // region Y (row-blurred input, later row-blurred A,B,C)
  float *sAR = smem + C::REGION_X;
  short *mapx = reinterpret_cast<short *>(smem + C::REGION_X + C::REGION_Y);
  short *mapy = mapx + (C::TW + 2 * RI);
  // mapx/mapy hold, for product column/row q of this tile, the Is-tile index of the pixel whose
  // gradient the reference would use there (reflect for the blur padding, then replicate for
  // the gradient border).  mapx2/mapy2 unused slots keep the struct simple.

  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * C::TW, y0 = blockIdx.y * C::TH;
  const size_t plane = (size_t)nx * ny;
  const size_t fofs = (size_t)blockIdx.z * plane;
  const bool border = (x0 - C::H < 0) || (y0 - C::H < 0) || (x0 + C::TW + C::H > nx) || (y0 + C::TH + C::H > ny);

  // ---- stage A: load tile (+halo) as float, reflecting at image borders ---------------------
  if (!border) {
    if (U8) {
      const unsigned char *src = static_cast<const unsigned char *>(frames) + fofs + (size_t)(y0 - C::H) * nx + (x0 - C::H);
      for (int i = tid; i < C::IN_H * C::IN_W; i += C::NT) {
        int r = i / C::IN_W, c = i - r * C::IN_W;
        sIN[r * C::IN_P + c] = (float)__ldg(src + (size_t)r * nx + c);
      }
    } else {
      const float *src = static_cast<const float *>(frames) + fofs + (size_t)(y0 - C::H) * nx + (x0 - C::H);
      for (int i = tid; i < C::IN_H * C::IN_W; i += C::NT) {
        int r = i / C::IN_W, c = i - r * C::IN_W;
        sIN[r * C::IN_P + c] = __ldg(src + (size_t)r * nx + c);
      }
    }
  } else {
    for (int i = tid; i < C::IN_H * C::IN_W; i += C::NT) {
      int r = i / C::IN_W, c = i - r * C::IN_W;
      int gy = reflect_index(y0 - C::H + r, ny), gx = reflect_index(x0 - C::H + c, nx);
      float v;
      if (U8) v = (float)__ldg(static_cast<const unsigned char *>(frames) + fofs + (size_t)gy * nx + gx);
      else v = __ldg(static_cast<const float *>(frames) + fofs + (size_t)gy * nx + gx);
      sIN[r * C::IN_P + c] = v;
    }
    // product-coordinate remap tables (only border tiles use them)
    for (int q = tid; q < C::TW + 2 * RI; q += C::NT) {
      int gx = reflect_index(x0 - RI + q, nx);            // blur padding of the product planes
      gx = min(max(gx, 1), nx - 2);                        // gradient replicate rule
      int li = gx - (x0 - C::G);                           // Is-tile column
      mapx[q] = (short)min(max(li, 1), C::IS_W - 2);
    }
    for (int q = tid; q < C::TH + 2 * RI; q += C::NT) {
      int gy = reflect_index(y0 - RI + q, ny);
      gy = min(max(gy, 1), ny - 2);
      int li = gy - (y0 - C::G);
      mapy[q] = (short)min(max(li, 1), C::IS_H - 2);
    }
  }
  __syncthreads();

  // ---- stage B: row blur sigma_d : sIN (IN_H x IN_W) -> sR1 (R1_H x R1_W) -------------------
  {
    constexpr int GROUPS = C::R1_W / 4;   // R1_W is a multiple of 4 for the instantiated radii
    static_assert(C::R1_W % 4 == 0, "row-blur width must be a multiple of 4");
    for (int it = tid; it < C::R1_H * GROUPS; it += C::NT) {
      int r = it / GROUPS, g = it - r * GROUPS;
      const float *p = sIN + r * C::IN_P + 4 * g;   // output col j uses input cols j .. j+2RD
      float v[4 + 2 * RD];
#pragma unroll
      for (int q = 0; q < (4 + 2 * RD + 3) / 4; q++) {
        float4 t = *reinterpret_cast<const float4 *>(p + 4 * q);
        if (4 * q + 0 < 4 + 2 * RD) v[4 * q + 0] = t.x;
        if (4 * q + 1 < 4 + 2 * RD) v[4 * q + 1] = t.y;
        if (4 * q + 2 < 4 + 2 * RD) v[4 * q + 2] = t.z;
        if (4 * q + 3 < 4 + 2 * RD) v[4 * q + 3] = t.w;
      }
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        float acc = kc.wd[0] * v[j + RD];
#pragma unroll
        for (int t = 1; t <= RD; t++) acc = fmaf(kc.wd[t], v[j + RD - t] + v[j + RD + t], acc);
        o[j] = acc;
      }
      *reinterpret_cast<float4 *>(sR1 + r * C::R1_P + 4 * g) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
  __syncthreads();

  // ---- stage C: column blur sigma_d : sR1 -> sIS (IS_H x IS_W), overwrites the input tile ----
  {
    constexpr int RB = (C::IS_H % 5 == 0) ? 5 : 4;
    static_assert(C::IS_H % RB == 0, "Is height must be a multiple of the column register block");
    for (int it = tid; it < C::IS_W * (C::IS_H / RB); it += C::NT) {
      int rg = it / C::IS_W, c = it - rg * C::IS_W;
      const float *p = sR1 + (rg * RB) * C::R1_P + c;   // output row j uses rows j .. j+2RD
      float acc[RB];
#pragma unroll
      for (int j = 0; j < RB; j++) acc[j] = 0.f;
#pragma unroll
      for (int q = 0; q < RB + 2 * RD; q++) {
        float v = p[q * C::R1_P];
#pragma unroll
        for (int j = 0; j < RB; j++) {
          int t = q - j - RD;                       // tap index relative to output j
          if (t >= -RD && t <= RD) acc[j] = fmaf(kc.wd[t < 0 ? -t : t], v, acc[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < RB; j++) sIS[(rg * RB + j) * C::IS_P + c] = acc[j];
    }
  }
  __syncthreads();

  // ---- stage D: gradient + products + row blur sigma_i : sIS -> sAR (3 x AR_H x TW) ----------
  {
    constexpr int GROUPS = C::TW / 4;
    constexpr int NP = 4 + 2 * RI;                  // product positions per item
    for (int it = tid; it < C::AR_H * GROUPS; it += C::NT) {
      int r = it / GROUPS, g = it - r * GROUPS;
      float pa[NP], pb[NP], pc[NP];
      if (!border) {
        // product row r <-> Is row r+1 ; product col q (0..TW+2RI) <-> Is col q+1
        const float *rm = sIS + (r + 1) * C::IS_P + 4 * g;
        float m[NP + 2], u[NP + 4], d[NP + 4];
#pragma unroll
        for (int q = 0; q < (NP + 2 + 3) / 4; q++) {
          float4 t = *reinterpret_cast<const float4 *>(rm + 4 * q);
          if (4 * q + 0 < NP + 2) m[4 * q + 0] = t.x;
          if (4 * q + 1 < NP + 2) m[4 * q + 1] = t.y;
          if (4 * q + 2 < NP + 2) m[4 * q + 2] = t.z;
          if (4 * q + 3 < NP + 2) m[4 * q + 3] = t.w;
          float4 a = *reinterpret_cast<const float4 *>(rm - C::IS_P + 4 * q);
          float4 b = *reinterpret_cast<const float4 *>(rm + C::IS_P + 4 * q);
          u[4 * q + 0] = a.x; u[4 * q + 1] = a.y; u[4 * q + 2] = a.z; u[4 * q + 3] = a.w;
          d[4 * q + 0] = b.x; d[4 * q + 1] = b.y; d[4 * q + 2] = b.z; d[4 * q + 3] = b.w;
        }
#pragma unroll
        for (int q = 0; q < NP; q++) {
          float gx, gy;
          if (GRAD == 0) {            // central differences; the 0.5 factors live in kc.wi
            gx = m[q + 2] - m[q];
            gy = d[q + 1] - u[q + 1];
          } else {                       // Sobel/8  (gradient.cpp:77-82), one rounding like the reference
            gx = fmaf(0.25f, m[q + 2] - m[q], 0.125f * (u[q + 2] + d[q + 2] - u[q] - d[q]));
            gy = fmaf(0.25f, d[q + 1] - u[q + 1], 0.125f * (d[q + 2] + d[q] - u[q + 2] - u[q]));
          }
          pa[q] = gx * gx; pb[q] = gx * gy; pc[q] = gy * gy;
        }
      } else {
        const int ry = mapy[r];
#pragma unroll
        for (int q = 0; q < NP; q++) {
          const int rx = mapx[4 * g + q];
          const float *c = sIS + ry * C::IS_P + rx;
          float gx, gy;
          if (GRAD == 0) {
            gx = c[1] - c[-1];
            gy = c[C::IS_P] - c[-C::IS_P];
          } else {
            gx = fmaf(0.25f, c[1] - c[-1], 0.125f * (c[-C::IS_P + 1] + c[C::IS_P + 1] - c[-C::IS_P - 1] - c[C::IS_P - 1]));
            gy = fmaf(0.25f, c[C::IS_P] - c[-C::IS_P], 0.125f * (c[C::IS_P + 1] + c[C::IS_P - 1] - c[-C::IS_P + 1] - c[-C::IS_P - 1]));
          }
          pa[q] = gx * gx; pb[q] = gx * gy; pc[q] = gy * gy;
        }
      }
      float oa[4], ob[4], oc[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        float a = kc.wir[0] * pa[j + RI], b = kc.wir[0] * pb[j + RI], c = kc.wir[0] * pc[j + RI];
#pragma unroll
        for (int t = 1; t <= RI; t++) {
          a = fmaf(kc.wir[t], pa[j + RI - t] + pa[j + RI + t], a);
          b = fmaf(kc.wir[t], pb[j + RI - t] + pb[j + RI + t], b);
          c = fmaf(kc.wir[t], pc[j + RI - t] + pc[j + RI + t], c);
        }
        oa[j] = a; ob[j] = b; oc[j] = c;
      }
      float *o = sAR + r * C::AR_P + 4 * g;
      *reinterpret_cast<float4 *>(o) = make_float4(oa[0], oa[1], oa[2], oa[3]);
      *reinterpret_cast<float4 *>(o + C::AR_H * C::AR_P) = make_float4(ob[0], ob[1], ob[2], ob[3]);
      *reinterpret_cast<float4 *>(o + 2 * C::AR_H * C::AR_P) = make_float4(oc[0], oc[1], oc[2], oc[3]);
    }
  }
  __syncthreads();

  // ---- stage E: column blur sigma_i + corner measure + store ---------------------------------
  {
    constexpr int RB = 8;
    static_assert(C::TH % RB == 0, "tile height must be a multiple of the register block");
    for (int it = tid; it < C::TW * (C::TH / RB); it += C::NT) {
      int rg = it / C::TW, c = it - rg * C::TW;
      float aa[RB], ab[RB], ac[RB];
#pragma unroll
      for (int j = 0; j < RB; j++) { aa[j] = 0.f; ab[j] = 0.f; ac[j] = 0.f; }
      const float *p = sAR + (rg * RB) * C::AR_P + c;
#pragma unroll
      for (int q = 0; q < RB + 2 * RI; q++) {
        float va = p[q * C::AR_P], vb = p[q * C::AR_P + C::AR_H * C::AR_P], vc = p[q * C::AR_P + 2 * C::AR_H * C::AR_P];
#pragma unroll
        for (int j = 0; j < RB; j++) {
          int t = q - j - RI;
          if (t >= -RI && t <= RI) {
            const float w = kc.wic[t < 0 ? -t : t];
            aa[j] = fmaf(w, va, aa[j]); ab[j] = fmaf(w, vb, ab[j]); ac[j] = fmaf(w, vc, ac[j]);
          }
        }
      }
      const int gx = x0 + c;
      if (gx < nx) {
#pragma unroll
        for (int j = 0; j < RB; j++) {
          int gy = y0 + rg * RB + j;
          if (gy < ny) {
            Rout[fofs + (size_t)gy * nx + gx] = corner_measure(aa[j], ab[j], ac[j], kc.k, kc.measure);
          }
        }
      }
    }
  }
}

}  // namespace b2f
