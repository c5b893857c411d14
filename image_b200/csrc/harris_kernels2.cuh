// harris_kernels2.cuh — second-generation fused Harris response kernel (interior tiles).
//
// Same chain and arithmetic class as harris_fused_kernel (fp32 FMA), restructured around Blackwell's
// packed fp32 instructions (fma.rn.f32x2 / add / mul -> SASS FFMA2/FADD2/FMUL2): every stage works on
// float2 lanes so that one issue slot advances two pixels.  The packed partner alternates with the
// pass direction, and the shared-memory layouts are chosen so that the partner pair is always one
// aligned 8-byte word:
//     row pass    (taps along x): partner = the pixel one ROW below  -> tile stored row-pair interleaved
//     column pass (taps along y): partner = the pixel one COLUMN right -> tile stored plain row-major
//   stage A  u8/f32 tile (+12 halo) -> sINp   [44][88(+2)] float2 (rows 2k,2k+1 interleaved)
//   stage B  row blur sigma_d        -> sR1    [88][80(+2)] float  (plain)
//   stage C  column blur sigma_d     -> sISp   [41][80(+2)] float2 (row-pair interleaved), aliases sINp
//   stage D  gradient, products, row blur sigma_i, streamed position by position (each product is
//            scattered into the <= 4 outputs it contributes to; nothing but accumulators stays live)
//                                    -> sAR    [3][78][64] float (plain), aliases sR1
//   stage E  column blur sigma_i + corner measure -> global R (8-byte stores)
// Global loads are issued in batches (all loads of a thread first, then the stores) so that memory
// latency is paid once per stage instead of once per element.
// This kernel only takes tiles whose +-12 halo lies inside the frame (no reflection / replicate logic
// at all); the remaining ring of border tiles is processed by harris_fused_kernel (v1) in a second
// launch that skips interior tiles.  Requires nx % 4 == 0 (aligned uchar4 / float2 global accesses).
#pragma once
#include "harris_kernels.cuh"

namespace b2f {

template <int RD, int RI> struct Fused2Cfg {
  static constexpr int TW = 64, TH = 64, NT = 256;
  static constexpr int HALO = 12;                               // >= RD + 1 + RI + 1, multiple of 4
  static constexpr int IN_W = TW + 2 * HALO, IN_H = TH + 2 * HALO;      // 88 x 88
  static constexpr int G = RI + 1;                              // halo of Is needed by the products (8)
  static constexpr int R1_W = TW + 2 * G;                       // 80 : global x0-G .. x0+TW+G-1
  static constexpr int R1_H = IN_H;                             // 88 : same rows as IN
  static constexpr int IS_H = TH + 2 * G;                       // 80 : global y0-G .. ; AR row a <-> Is row a+1, so an AR row pair straddles two Is pair lines
  static constexpr int IS_W = R1_W;
  static constexpr int AR_H = TH + 2 * RI;                      // 78 : global y0-RI ..
  // Pitches.  Whenever consecutive threads walk consecutive ROWS (row passes B and D: 16-byte loads of
  // float2 pairs, 16-byte stores of 4 outputs) the row stride must be an odd multiple of 16 bytes
  // modulo 128 so that a quarter warp covers all 32 banks: pitch % 4 == 2 for float2 rows, and for the
  // plain float tiles written two rows at a time.
  static constexpr int IN_P = IN_W + 2;                         // 90 float2
  static constexpr int IS_P = IS_W + 2;                         // 82 float2
  static constexpr int R1_P = R1_W + 2;                         // 82 float
  static constexpr int AR_P = TW + 2;                           // 66 float
  static_assert(IN_P % 4 == 2 && IS_P % 4 == 2 && R1_P % 4 == 2 && AR_P % 4 == 2, "bank-conflict-free pitches");
  static constexpr int REGION_X = (IN_H * IN_P > IS_H * IS_P) ? IN_H * IN_P : IS_H * IS_P;   // floats
  static constexpr int REGION_Y = (R1_H * R1_P > 3 * AR_H * AR_P) ? R1_H * R1_P : 3 * AR_H * AR_P;
  static constexpr size_t SMEM = sizeof(float) * (REGION_X + REGION_Y);
  static_assert(RD + 1 + RI + 1 <= HALO, "halo too small");
  static_assert(HALO - G - RD >= 0, "row-blur taps must stay inside the input tile");
};

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2s(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __ffma2_rn(b, f2s(-1.f), a); }   // a - b, one rounding

template <int RD, int RI, bool U8, int GRAD>
__global__ void __launch_bounds__(256, 2)
harris_fused2_kernel(const void *__restrict__ frames, float *__restrict__ Rout, int nx, int ny,
                     const __grid_constant__ HarrisConsts kc) {
  using C = Fused2Cfg<RD, RI>;
  const int x0 = blockIdx.x * C::TW, y0 = blockIdx.y * C::TH;
  if (!harris_tile_is_interior(x0, y0, nx, ny)) return;       // border ring: second launch (v1 kernel)
  extern __shared__ __align__(16) float smem[];
  float2 *sINp = reinterpret_cast<float2 *>(smem);             // [IN_H/2][IN_W]
  float2 *sISp = reinterpret_cast<float2 *>(smem);             // [IS_H/2][IS_W]   (aliases sINp)
  float *sR1 = smem + C::REGION_X;                             // [R1_H][R1_W]
  float *sAR = smem + C::REGION_X;                             // [3][AR_H][TW]    (aliases sR1)
  const int tid = threadIdx.x;
  const size_t plane = (size_t)nx * ny;
  const size_t fofs = (size_t)blockIdx.z * plane;

  // ---- stage A: batched vector loads -> row-pair interleaved float2 tile ----------------------
  {
    constexpr int VW = C::IN_W / 4;                            // 22 vectors of 4 pixels per row
    constexpr int ITEMS = (C::IN_H / 2) * VW;                  // 968 (row pair, vector)
    constexpr int PER = (ITEMS + C::NT - 1) / C::NT;           // 4
    if (U8) {
      const unsigned char *base = static_cast<const unsigned char *>(frames) + fofs + (size_t)(y0 - C::HALO) * nx + (x0 - C::HALO);
      unsigned a[PER], b[PER];
#pragma unroll
      for (int k = 0; k < PER; k++) {
        const int it = tid + k * C::NT;
        const int rp = min(it, ITEMS - 1) / VW, v = min(it, ITEMS - 1) - rp * VW;
        const unsigned char *p = base + (size_t)(2 * rp) * nx + 4 * v;
        a[k] = __ldg(reinterpret_cast<const unsigned *>(p));
        b[k] = __ldg(reinterpret_cast<const unsigned *>(p + nx));
      }
#pragma unroll
      for (int k = 0; k < PER; k++) {
        const int it = tid + k * C::NT;
        if (it < ITEMS) {
          const int rp = it / VW, v = it - rp * VW;
          float4 *d = reinterpret_cast<float4 *>(sINp + rp * C::IN_P + 4 * v);
          d[0] = make_float4((float)(a[k] & 0xff), (float)(b[k] & 0xff), (float)((a[k] >> 8) & 0xff), (float)((b[k] >> 8) & 0xff));
          d[1] = make_float4((float)((a[k] >> 16) & 0xff), (float)((b[k] >> 16) & 0xff), (float)(a[k] >> 24), (float)(b[k] >> 24));
        }
      }
    } else {
      const float *base = static_cast<const float *>(frames) + fofs + (size_t)(y0 - C::HALO) * nx + (x0 - C::HALO);
      float4 a[PER], b[PER];
#pragma unroll
      for (int k = 0; k < PER; k++) {
        const int it = tid + k * C::NT;
        const int rp = min(it, ITEMS - 1) / VW, v = min(it, ITEMS - 1) - rp * VW;
        const float *p = base + (size_t)(2 * rp) * nx + 4 * v;
        a[k] = __ldg(reinterpret_cast<const float4 *>(p));
        b[k] = __ldg(reinterpret_cast<const float4 *>(p + nx));
      }
#pragma unroll
      for (int k = 0; k < PER; k++) {
        const int it = tid + k * C::NT;
        if (it < ITEMS) {
          const int rp = it / VW, v = it - rp * VW;
          float4 *d = reinterpret_cast<float4 *>(sINp + rp * C::IN_P + 4 * v);
          d[0] = make_float4(a[k].x, b[k].x, a[k].y, b[k].y);
          d[1] = make_float4(a[k].z, b[k].z, a[k].w, b[k].w);
        }
      }
    }
  }
  __syncthreads();

  // ---- stage B: row blur sigma_d, packed over row pairs -> sR1 (plain) ------------------------
  {
    constexpr int GROUPS = C::R1_W / 4;                        // 20
    constexpr int OFF = C::HALO - C::G - RD;                   // first tap of R1 col 0 sits at IN col OFF (=1)
    constexpr int NP = 4 + 2 * RD + OFF;                       // positions loaded from the 4-aligned start (11) -> round up even
    constexpr int NL = (NP + 1) / 2 * 2;                       // 12
    constexpr int RPS = C::R1_H / 2;                           // 44 row pairs; consecutive threads = consecutive row pairs
    // software pipelined: the loads of item k+1 are issued before the stores of item k (the two tiles do not alias)
    constexpr int NIT = (RPS * GROUPS + C::NT - 1) / C::NT;    // 4
    float2 v[NL];
    {
      const int g = tid / RPS, rp = tid - g * RPS;             // (tid < RPS * GROUPS)
      const float2 *p = sINp + rp * C::IN_P + 4 * g;
#pragma unroll
      for (int q = 0; q < NL / 2; q++) { const float4 t = *reinterpret_cast<const float4 *>(p + 2 * q); v[2 * q] = f2(t.x, t.y); v[2 * q + 1] = f2(t.z, t.w); }
    }
#pragma unroll
    for (int k = 0; k < NIT; k++) {
      const int it = tid + k * C::NT;
      if (it < RPS * GROUPS) {
        const int g = it / RPS, rp = it - g * RPS;
        float2 o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int c = j + OFF + RD;
          float2 acc = __fmul2_rn(f2s(kc.wd[0]), v[c]);
#pragma unroll
          for (int t = 1; t <= RD; t++) acc = __ffma2_rn(f2s(kc.wd[t]), __fadd2_rn(v[c - t], v[c + t]), acc);
          o[j] = acc;
        }
        const int itn = it + C::NT;
        if (k + 1 < NIT && itn < RPS * GROUPS) {
          const int gn = itn / RPS, rpn = itn - gn * RPS;
          const float2 *p = sINp + rpn * C::IN_P + 4 * gn;
#pragma unroll
          for (int q = 0; q < NL / 2; q++) { const float4 t = *reinterpret_cast<const float4 *>(p + 2 * q); v[2 * q] = f2(t.x, t.y); v[2 * q + 1] = f2(t.z, t.w); }
        }
        float *d = sR1 + (2 * rp) * C::R1_P + 4 * g;
        *reinterpret_cast<float4 *>(d) = make_float4(o[0].x, o[1].x, o[2].x, o[3].x);
        *reinterpret_cast<float2 *>(d + C::R1_P) = make_float2(o[0].y, o[1].y);      // odd rows are only 8-byte aligned
        *reinterpret_cast<float2 *>(d + C::R1_P + 2) = make_float2(o[2].y, o[3].y);
      }
    }
  }
  __syncthreads();

  // ---- stage C: column blur sigma_d, packed over column pairs -> sISp (row-pair interleaved) --
  {
    constexpr int CP = C::IS_W / 2;                            // 40 column pairs
    constexpr int RB = 10;                                     // Is rows per item (5 row pairs)
    constexpr int RGS = (C::IS_H + RB - 1) / RB;               // 8
    constexpr int OFFR = C::HALO - C::G - RD;                  // Is row 0 (global y0-G) uses R1 rows OFFR .. OFFR+2RD (=1)
    static_assert(OFFR >= 0, "R1 must cover the taps of Is row 0");
    for (int it = tid; it < CP * RGS; it += C::NT) {
      const int rg = it / CP, cp = it - rg * CP;
      const float *p = sR1 + (rg * RB + OFFR) * C::R1_P + 2 * cp;
      float2 acc[RB];
#pragma unroll
      for (int j = 0; j < RB; j++) acc[j] = f2s(0.f);
#pragma unroll
      for (int q = 0; q < RB + 2 * RD; q++) {
        const int row = min(rg * RB + OFFR + q, C::R1_H - 1) - (rg * RB + OFFR);      // clamp (last group overruns)
        const float2 v = *reinterpret_cast<const float2 *>(p + row * C::R1_P);
#pragma unroll
        for (int j = 0; j < RB; j++) {
          const int t = q - j - RD;
          if (t >= -RD && t <= RD) acc[j] = __ffma2_rn(f2s(kc.wd[t < 0 ? -t : t]), v, acc[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < RB; j += 2) {
        const int i = rg * RB + j;                             // Is row (even)
        if (i < C::IS_H)
          *reinterpret_cast<float4 *>(sISp + (i >> 1) * C::IS_P + 2 * cp) = make_float4(acc[j].x, acc[j + 1].x, acc[j].y, acc[j + 1].y);
      }
    }
  }
  __syncthreads();

  // ---- stage D: gradient + products + row blur sigma_i, streamed -> sAR (plain) ---------------
  // One item = one AR row pair x 4 output columns.  AR rows (2r, 2r+1) <-> Is rows (2r+1, 2r+2); with Is
  // rows paired (even, odd), the two pair lines L0 = (2r, 2r+1) and L1 = (2r+2, 2r+3) hold everything:
  //     gy = (Is[2r+2] - Is[2r], Is[2r+3] - Is[2r+1]) = L1 - L0 (packed),   gx from L0.y and L1.x.
  // Consecutive threads take consecutive row pairs (line stride = 4 banks: a quarter warp of 16-byte loads
  // covers all 32 banks); 39 row pairs are padded to 40 so that quarter warps never straddle two column groups.
  {
    constexpr int GROUPS = C::TW / 4;                          // 16
    constexpr int NPOS = 4 + 2 * RI;                           // 18 product positions per item
    constexpr int RPS = C::AR_H / 2, RPS_PAD = (RPS + 7) & ~7; // 39 -> 40
    constexpr int PL = C::AR_H * C::AR_P;                      // plane stride (multiple of 4 floats)
    static_assert(PL % 4 == 0, "plane stride keeps 16-byte alignment");
    static_assert(C::AR_H % 2 == 0 && RPS + 1 <= C::IS_H / 2, "row pairs and their two pair lines");
    float2 w2[RI + 1];
#pragma unroll
    for (int t = 0; t <= RI; t++) w2[t] = f2s(kc.wir[t]);
    for (int it = tid; it < RPS_PAD * GROUPS; it += C::NT) {
      const int g = it / RPS_PAD, r = it - g * RPS_PAD;
      if (r >= RPS) continue;
      const float2 *l0 = sISp + r * C::IS_P + 4 * g;           // pair line r   = Is rows 2r,   2r+1
      const float2 *l1 = l0 + C::IS_P;                         // pair line r+1 = Is rows 2r+2, 2r+3
      float2 a0[4], b0[4], c0[4];
#pragma unroll
      for (int j = 0; j < 4; j++) { a0[j] = b0[j] = c0[j] = f2s(0.f); }
      float2 p0[3], p1[3];                                     // sliding window of Is columns q, q+1, q+2
      {
        const float4 t0 = *reinterpret_cast<const float4 *>(l0), t1 = *reinterpret_cast<const float4 *>(l1);
        p0[0] = f2(t0.x, t0.y); p0[1] = f2(t0.z, t0.w); p1[0] = f2(t1.x, t1.y); p1[1] = f2(t1.z, t1.w);
      }
#pragma unroll
      for (int q = 0; q < NPOS; q++) {
        float2 n0 = f2s(0.f), n1 = f2s(0.f);
        if ((q & 1) == 0) {                                    // columns q+2, q+3 arrive as one 16-byte load per line
          const float4 t0 = *reinterpret_cast<const float4 *>(l0 + q + 2), t1 = *reinterpret_cast<const float4 *>(l1 + q + 2);
          p0[2] = f2(t0.x, t0.y); p1[2] = f2(t1.x, t1.y);
          n0 = f2(t0.z, t0.w); n1 = f2(t1.z, t1.w);
        }
        // ---- products at column q+1
        float2 gx, gy;
        if (GRAD == 0) {
          gy = sub2(p1[1], p0[1]);
          gx = f2(p0[2].y - p0[0].y, p1[2].x - p1[0].x);
        } else {
          gx = f2(fmaf(0.25f, p0[2].y - p0[0].y, 0.125f * (p0[2].x + p1[2].x - p0[0].x - p1[0].x)),
                  fmaf(0.25f, p1[2].x - p1[0].x, 0.125f * (p0[2].y + p1[2].y - p0[0].y - p1[0].y)));
          gy = f2(fmaf(0.25f, p1[1].x - p0[1].x, 0.125f * (p1[2].x + p1[0].x - p0[2].x - p0[0].x)),
                  fmaf(0.25f, p1[1].y - p0[1].y, 0.125f * (p1[2].y + p1[0].y - p0[2].y - p0[0].y)));
        }
        const float2 pa = __fmul2_rn(gx, gx), pb = __fmul2_rn(gx, gy), pc = __fmul2_rn(gy, gy);
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int t = q - j - RI;
          if (t >= -RI && t <= RI) {
            const float2 w = w2[t < 0 ? -t : t];
            a0[j] = __ffma2_rn(w, pa, a0[j]); b0[j] = __ffma2_rn(w, pb, b0[j]); c0[j] = __ffma2_rn(w, pc, c0[j]);
          }
        }
        // slide: (q, q+1, q+2) -> (q+1, q+2, q+3)
        p0[0] = p0[1]; p0[1] = p0[2]; p1[0] = p1[1]; p1[1] = p1[2];
        if ((q & 1) == 0) { p0[2] = n0; p1[2] = n1; }
      }
      float *o = sAR + (2 * r) * C::AR_P + 4 * g;
#define ST_ROWPAIR(dst, v)                                                                   \
      *reinterpret_cast<float4 *>(dst) = make_float4(v[0].x, v[1].x, v[2].x, v[3].x);            \
      *reinterpret_cast<float2 *>((dst) + C::AR_P) = make_float2(v[0].y, v[1].y);                \
      *reinterpret_cast<float2 *>((dst) + C::AR_P + 2) = make_float2(v[2].y, v[3].y);
      ST_ROWPAIR(o, a0) ST_ROWPAIR(o + PL, b0) ST_ROWPAIR(o + 2 * PL, c0)
#undef ST_ROWPAIR
    }
  }
  __syncthreads();

  // ---- stage E: column blur sigma_i, packed over column pairs, + corner measure + store -------
  {
    constexpr int RB = 8;
    constexpr int CP = C::TW / 2;                              // 32
    constexpr int PL = C::AR_H * C::AR_P;
    float2 w2[RI + 1];
#pragma unroll
    for (int t = 0; t <= RI; t++) w2[t] = f2s(kc.wic[t]);
    for (int it = tid; it < CP * (C::TH / RB); it += C::NT) {
      const int rg = it / CP, cp = it - rg * CP;
      const float *p = sAR + (rg * RB) * C::AR_P + 2 * cp;    // output row j uses AR rows j .. j+2RI
      float2 aa[RB], ab[RB], ac[RB];
#pragma unroll
      for (int j = 0; j < RB; j++) { aa[j] = f2s(0.f); ab[j] = f2s(0.f); ac[j] = f2s(0.f); }
#pragma unroll
      for (int q = 0; q < RB + 2 * RI; q++) {
        const float2 va = *reinterpret_cast<const float2 *>(p + q * C::AR_P);
        const float2 vb = *reinterpret_cast<const float2 *>(p + q * C::AR_P + PL);
        const float2 vc = *reinterpret_cast<const float2 *>(p + q * C::AR_P + 2 * PL);
#pragma unroll
        for (int j = 0; j < RB; j++) {
          const int t = q - j - RI;
          if (t >= -RI && t <= RI) {
            const float2 w = w2[t < 0 ? -t : t];
            aa[j] = __ffma2_rn(w, va, aa[j]); ab[j] = __ffma2_rn(w, vb, ab[j]); ac[j] = __ffma2_rn(w, vc, ac[j]);
          }
        }
      }
      float *dst = Rout + fofs + (size_t)(y0 + rg * RB) * nx + x0 + 2 * cp;
#pragma unroll
      for (int j = 0; j < RB; j++) {
        float2 r;
        if (kc.measure == 0) {                                 // Harris: (A*C - B*B) - (k*tr)*tr, each op rounded (harris.cpp:100-103)
          const float2 det = sub2(__fmul2_rn(aa[j], ac[j]), __fmul2_rn(ab[j], ab[j]));
          const float2 tr = __fadd2_rn(aa[j], ac[j]);
          r = sub2(det, __fmul2_rn(__fmul2_rn(f2s(kc.k), tr), tr));
        } else {
          r = f2(corner_measure(aa[j].x, ab[j].x, ac[j].x, kc.k, kc.measure), corner_measure(aa[j].y, ab[j].y, ac[j].y, kc.k, kc.measure));
        }
        *reinterpret_cast<float2 *>(dst + (size_t)j * nx) = r;
      }
    }
  }
}

}  // namespace b2f
