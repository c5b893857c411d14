// harris_kernels3.cuh — third-generation fused Harris response kernel: ONE launch for the whole frame.
//
// Same arithmetic as the second generation (fp32 FMA on packed f32x2 lanes, SASS FFMA2/FMUL2/FADD2; the
// shared-memory layouts alternate between row-pair interleaved float2 tiles for row passes and plain
// row-major tiles for column passes so that a packed partner pair is always one aligned 8-byte word), plus
//   * the border ring is handled by the same kernel: a border tile (or any tile of a frame whose rows are not
//     4-pixel aligned) takes the GENERIC variants of stage A (scalar loads, reflected at load time:
//     gaussian.cpp:345-349) and stage D (product coordinates remapped: reflect padding of the product planes,
//     then the replicate rule of the gradient, gradient.cpp:40-55) and guarded stores — CTA-uniform branches;
//   * a certified error bound: stage E tracks max(trace) per 8x8 pixel block and writes, per block, a bound
//     eps >= |R_fp32 - R_reference| (harris_eps below) that the tolerant NMS uses to decide which key points
//     need the exact recomputation (harris_exact_patch_kernel);
//   * plain tiles store their odd rows shifted by two floats, so that both rows of a row-pair store are
//     16-byte aligned and conflict free (the second generation's odd rows were two 8-byte stores, 2-way);
//   * tile height, CTA size and the stage-D item width are template parameters: <64,256,4> is the second
//     generation's shape (2 CTAs / SM), <108,512,8> a tall tile (1 CTA / SM, 20 tile rows per 2160-row frame)
//     with 13 % vertical halo over-compute instead of 22-38 % and stage-D items of 8 outputs that re-compute
//     each gradient product 2.75 times instead of 4.5 times;
//   * optional TMA staging of the u8 tile (cp.async.bulk.tensor.3d + mbarrier) in a persistent loop: the
//     bytes of tile k+1 stream into shared memory while tile k is computed.
// Reference chain: image.CornerDetectionHarris/src/harris.cpp:511-520 (gaussian.cpp:289-395, gradient.cpp:17-128,
// harris.cpp:44-133).
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace b2f {

constexpr int HARRIS_MAX_TAPS = 32;

struct HarrisConsts {
  float wd[HARRIS_MAX_TAPS];    // sigma_d taps, wd[0] centre (float-rounded reference weights)
  float wir[HARRIS_MAX_TAPS];   // sigma_i taps of the ROW pass: x0.25 for central differences (the two 0.5
                                // factors of gradient.cpp:35-36 commute exactly with the float roundings), x1 for Sobel
  float wic[HARRIS_MAX_TAPS];   // sigma_i taps of the COLUMN pass (unscaled)
  float k;
  int measure;   // 0 Harris, 1 Shi-Tomasi, 2 harmonic mean
  float tr_cut;  // > 0 (certified corner path, u8 frames, Harris measure, k >= 0): pixels whose trace is below it are stored
                 // as -FLT_MAX — their reference response is certainly below the threshold (harris_trace_cut)
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

__device__ __forceinline__ int reflect_index(int p, int n) {
  // padding rule of discrete_gaussian: -k -> k ; n-1+k -> n-k   (gaussian.cpp:345-349)
  if (p < 0) p = -p;
  else if (p >= n) p = 2 * n - 1 - p;
  return min(max(p, 0), n - 1);
}

// Bound on |R_fused - R_reference| for the Harris measure (measure 0), from the trace T = A+C of the smoothed
// tensor (an upper bound of each of A, C and 2|B|), the largest |pixel| M that can reach the block and k.
// Derivation (DESIGN.md "certified Harris"): u = 2^-24.
//   sigma_d blur: fp32 FMA chain + float-rounded taps vs double accumulation rounded once:  e_I <= 16 u M
//   gradient (central or Sobel, weights sum to 1):                                           e_G <= e_I + 8 u |G|
//   products, 2-D sigma_i blur (non-negative taps summing to 1, Jensen: blur|G| <= sqrt(T)):
//                                       e_T <= 2 e_I sqrt(T) + e_I^2 + 64 u T      (each of A, B, C)
//   measure (AC - BB) - (k tr) tr with both sides' roundings:
//                                       eps <= (2+4|k|)(T e_T + e_T^2) + 2(1+3|k|) u T^2
// evaluated with 17 for 16 and a final factor 1.25 (fp32 evaluation, T taken from the fp32 planes).
__device__ __forceinline__ float harris_eps(float T, float M, float k) {
  const float u = 5.9604645e-8f;
  const float eI = 17.f * u * M;
  const float eT = fmaf(2.f * eI, sqrtf(T), fmaf(eI, eI, 64.f * u * T));
  const float kk = fabsf(k);
  const float e = fmaf(2.f + 4.f * kk, fmaf(T, eT, eT * eT), 2.f * (1.f + 3.f * kk) * u * T * T);
  return fmaf(e, 1.25f, 1e-30f);
}

// Certificate "the reference response of this pixel is below Th" from the fp32 trace t = A + C alone (Harris measure,
// k >= 0, u8 frames).  The reference computes R = fl(fl(fl(A C) - fl(B B)) - fl(fl(k tr) tr)) from non-negative A, C:
// both subtrahends are >= 0 and rounding is monotone, so R <= fl(A C) <= ((A + C) / 2)^2 (1 + u).  Its trace differs from
// the fused one by at most 2 e_T, and e_T <= 2 e_I sqrt(T) + e_I^2 + 64 u T is largest at the largest trace u8 frames can
// produce (|gradient| <= 127.5: T <= 32512, e_T <= 0.2172).  Hence t < 2 sqrt(Th) - 2 e_T  =>  R_reference < Th; the cut
// keeps the bound's factor 1.25 on e_T and a relative 1e-6 on the root.  Such a pixel cannot be a corner, and as a
// neighbour it is below every kept corner (whose response is >= Th), so the non-maximum test may see -FLT_MAX in its
// place: beside strong edges — where the per-block eps is wide and the response small — this removes the band of
// undecided candidates the block bound would otherwise send to the exact patches.
static inline float harris_trace_cut(float Th, float k, int measure, bool u8) {
  if (!u8 || measure != 0 || !(k >= 0.f) || !(Th > 0.f)) return 0.f;
  const float cut = 2.f * sqrtf(Th) * (1.f - 1e-6f) - 2.f * 0.2172f * 1.25f;
  return cut > 0.f ? cut : 0.f;
}

template <int RD_, int RI_, int TH_, int NT_, int DW_, int CTAS_> struct Fused3Cfg {
  static constexpr int RD = RD_, RI = RI_;
  static constexpr int TW = 64, TH = TH_, NT = NT_, DW = DW_, CTAS = CTAS_;
  static constexpr int HALO = 12;                               // >= RD + 1 + RI + 1, multiple of 4
  static constexpr int IN_W = TW + 2 * HALO, IN_H = TH + 2 * HALO;
  static constexpr int G = RI + 1;                              // halo of Is needed by the products
  static constexpr int R1_W = TW + 2 * G;                       // global x0-G .. x0+TW+G-1
  static constexpr int R1_H = IN_H;                             // same rows as IN
  static constexpr int IS_H = TH + 2 * G, IS_W = R1_W;          // AR row a <-> Is row a+1
  static constexpr int AR_H = TH + 2 * RI;                      // global y0-RI ..
  // Pitches.  Whenever consecutive threads walk consecutive ROW PAIRS (row passes B and D: 16-byte loads of float2
  // pairs, 16-byte stores of 4 outputs) the pair stride must be an odd multiple of 16 bytes modulo 128 so that a
  // quarter warp covers all 32 banks: pitch % 4 == 2 for the float2 tiles and for the plain float tiles (whose odd
  // rows are shifted by 2 floats to be 16-byte aligned as well).
  static constexpr int IN_P = IN_W + 2;                         // float2
  static constexpr int IS_P = IS_W + 2;                         // float2
  static constexpr int R1_P = R1_W + 2;                         // float (+2 holds the odd-row shift)
  static constexpr int AR_P = TW + 2;                           // float
  static_assert(IN_P % 4 == 2 && IS_P % 4 == 2 && R1_P % 4 == 2 && AR_P % 4 == 2, "bank-conflict-free pitches");
  static_assert(IN_H % 2 == 0 && IS_H % 2 == 0 && AR_H % 2 == 0, "row pairs");
  // Shared-memory layout (floats).  Live ranges: IN (A-B), R1 (B-C), IS (C-D), AR (D-E).  IS sits at 0 and AR right
  // behind it (both live in stage D); IN also starts at 0 and may run into AR's space; R1 takes the tail of AR's
  // space, behind IN (both live in stage B) and behind IS (both live in stage C).
  static constexpr int IN_N = IN_H / 2 * IN_P * 2, IS_N = IS_H / 2 * IS_P * 2, R1_N = R1_H * R1_P, AR_N = 3 * AR_H * AR_P;
  static constexpr int OFF_AR = (IS_N + 3) & ~3;
  static constexpr int OFF_R1 = ((((OFF_AR + AR_N - R1_N) > IN_N ? (OFF_AR + AR_N - R1_N) : IN_N) + 3) & ~3);
  static constexpr int TILE_N = ((OFF_AR + AR_N > OFF_R1 + R1_N ? OFF_AR + AR_N : OFF_R1 + R1_N) + 3) & ~3;
  static_assert(OFF_R1 >= IN_N && OFF_R1 >= IS_N && OFF_AR >= IS_N, "live tiles must not overlap");
  static constexpr int MAP_N = (TW + 2 * RI) + (TH + 2 * RI);   // product-coordinate remap tables (generic tiles)
  // TMA staging buffer: box 96 x IN_H bytes starting at column x0-16 (the global start of every box row must be
  // 16-byte aligned), so input column c = x0-12+c' sits at box column c'+4
  static constexpr int U8_P = 96, U8_BYTES = IN_H * U8_P, U8_X0 = 16, U8_SKIP = U8_X0 - HALO;
  static_assert(U8_SKIP >= 0 && U8_SKIP + IN_W <= U8_P, "the box covers the input tile");
  static constexpr size_t SMEM_CORE = sizeof(float) * TILE_N + sizeof(short) * ((MAP_N + 7) & ~7) + sizeof(float) * (NT / 32);
  static constexpr size_t SMEM = SMEM_CORE;
  static constexpr size_t SMEM_TMA = ((SMEM_CORE + 127) & ~size_t(127)) + U8_BYTES + 128;
  static_assert(RD + 1 + RI + 1 <= HALO, "halo too small");
  static_assert(HALO - G - RD >= 0, "row-blur taps must stay inside the input tile");
  static_assert(DW == 4 || DW == 8, "stage-D item width");
  // stage C: RB Is rows per item (even), stage E: RB output rows per item
  static constexpr int C_RB = (IS_H % 10 == 0) ? 10 : ((IS_H % 8 == 0) ? 8 : 12);
  static constexpr int E_NRG = NT / (TW / 2);                   // row groups of stage E
  static constexpr int E_RB = (TH + E_NRG - 1) / E_NRG;
  static_assert(E_RB <= 8, "a stage-E item must not span more than two 8-row eps blocks");
};

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2s(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __ffma2_rn(b, f2s(-1.f), a); }   // a - b, one rounding

template <class C>
__device__ __forceinline__ bool harris3_tile_is_interior(int x0, int y0, int nx, int ny) {
  return x0 >= C::HALO && y0 >= C::HALO && x0 + C::TW + C::HALO <= nx && y0 + C::TH + C::HALO <= ny;
}

// ---- TMA / mbarrier helpers (raw PTX; see cute/arch/copy_sm90_tma.hpp for the same strings) -------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, unsigned long long *bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// TMA = 0: grid (tiles_x, tiles_y, frames), one tile per CTA, register-staged global loads.
// TMA = 1: persistent grid; u8 frames only; interior tiles get their bytes through the tensor map (prefetched one
//          tile ahead), generic tiles load with scalar reflected loads.
template <class C, bool U8, int GRAD, bool TMA>
__global__ void __launch_bounds__(C::NT, C::CTAS)
harris_fused3_kernel(const void *__restrict__ frames, float *__restrict__ Rout, unsigned *__restrict__ eps_blk,
                     int nx, int ny, int n_frames, int generic_all, const __grid_constant__ HarrisConsts kc,
                     const __grid_constant__ CUtensorMap tmap) {
  constexpr int RD = C::RD, RI = C::RI;
  extern __shared__ __align__(128) float smem[];
  float2 *sINp = reinterpret_cast<float2 *>(smem);             // [IN_H/2][IN_P]
  float2 *sISp = reinterpret_cast<float2 *>(smem);             // [IS_H/2][IS_P]   (aliases sINp)
  float *sR1 = smem + C::OFF_R1;                               // [R1_H][R1_P], odd rows shifted by 2 (tail of AR's space)
  float *sAR = smem + C::OFF_AR;                               // [3][AR_H][AR_P], odd rows shifted by 2
  short *mapx = reinterpret_cast<short *>(smem + C::TILE_N);
  short *mapy = mapx + (C::TW + 2 * RI);
  float *sM = reinterpret_cast<float *>(mapx + ((C::MAP_N + 7) & ~7));         // per-warp max |pixel| of the tile
  const int tid = threadIdx.x;
  const size_t plane = (size_t)nx * ny;
  const int tiles_x = (nx + C::TW - 1) / C::TW, tiles_y = (ny + C::TH - 1) / C::TH;
  const int ebx = (nx + 7) >> 3, eby = (ny + 7) >> 3;          // eps grid
  unsigned char *sU8 = nullptr;
  unsigned long long *bar = nullptr;
  if (TMA) {
    sU8 = reinterpret_cast<unsigned char *>(smem) + ((C::SMEM_CORE + 127) & ~size_t(127));
    bar = reinterpret_cast<unsigned long long *>(sU8 + C::U8_BYTES);
  }

  int tile = TMA ? (int)blockIdx.x : 0;
  const int n_tiles = TMA ? tiles_x * tiles_y * n_frames : 1;
  const int tile_step = TMA ? (int)gridDim.x : 1;
  unsigned phase = 0;
  auto decode = [&](int t, int &tx0, int &ty0, int &tf) {
    const int per = tiles_x * tiles_y;
    tf = t / per;
    const int r = t - tf * per;
    const int ty = r / tiles_x;
    tx0 = (r - ty * tiles_x) * C::TW;
    ty0 = ty * C::TH;
  };
  if (TMA) {
    if (tid == 0) {
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0 && tile < n_tiles) {
      int tx0, ty0, tf;
      decode(tile, tx0, ty0, tf);
      if (!generic_all && harris3_tile_is_interior<C>(tx0, ty0, nx, ny)) {
        mbar_expect_tx(bar, C::U8_BYTES);
        tma_load_3d(sU8, &tmap, bar, tx0 - C::U8_X0, ty0 - C::HALO, tf);
      }
    }
  }

  for (; tile < n_tiles; tile += tile_step) {
    int x0, y0, frame;
    if (TMA) decode(tile, x0, y0, frame);
    else { x0 = blockIdx.x * C::TW; y0 = blockIdx.y * C::TH; frame = blockIdx.z; }
    const size_t fofs = (size_t)frame * plane;
    const bool generic = generic_all || !harris3_tile_is_interior<C>(x0, y0, nx, ny);
    float mloc = 0.f;

    // ---- stages A+B fused (interior u8 tiles, register-staged): the sigma_d row blur straight from the u8 words ----
    // A thread takes (row pair, 4 output columns): three 4-byte words per row cover its 4 + 2*RD + OFF input columns;
    // consecutive lanes = consecutive column groups, so the loads of a warp are contiguous and the row-pair stores are
    // 16 contiguous bytes per lane.  Bytes become floats with one PRMT into the mantissa of 2^23 and one packed subtract
    // (no I2F, no float input tile in shared memory: a quarter of the kernel's shared-memory traffic goes away).
    constexpr bool FUSED_AB = U8 && !TMA;
    if (FUSED_AB && !generic) {
      constexpr int GROUPS = C::R1_W / 4;                      // 20
      constexpr int OFF = C::HALO - C::G - RD;                 // R1 col j uses input cols j+OFF .. j+OFF+2RD
      constexpr int NW = (4 + 2 * RD + OFF + 3) / 4;           // words per row and item (3)
      constexpr int RPS = C::R1_H / 2, ITEMS = RPS * GROUPS, PER = (ITEMS + C::NT - 1) / C::NT;
      static_assert(4 * (GROUPS - 1) + 4 * NW <= C::IN_W, "the last item's words stay inside the input tile");
      static_assert(C::NT % GROUPS + GROUPS - 1 < 2 * GROUPS, "one carry per step");
      // 32-bit word offsets from the tile origin; item k of a thread is item k-1 plus NT: (rp, g) advance by
      // (NT / GROUPS, NT % GROUPS) with one carry — no division per item
      const unsigned *org = reinterpret_cast<const unsigned *>(static_cast<const unsigned char *>(frames) + fofs + (size_t)(y0 - C::HALO) * nx + (x0 - C::HALO));
      const int nxw = nx >> 2;                                 // words per frame row
      constexpr int DRP = C::NT / GROUPS, DG = C::NT % GROUPS;
      const int rp0 = tid / GROUPS, g0 = tid - rp0 * GROUPS;
      unsigned wa[PER][NW], wb[PER][NW];
      {
        int rp = rp0, g = g0;
#pragma unroll
        for (int k = 0; k < PER; k++) {
          const int rpc = min(rp, RPS - 1);                    // (threads past the last item re-read the last row pair)
          const unsigned *p = org + (2 * rpc) * nxw + g;
#pragma unroll
          for (int q = 0; q < NW; q++) { wa[k][q] = __ldg(p + q); wb[k][q] = __ldg(p + q + nxw); }
          g += DG; rp += DRP;
          if (g >= GROUPS) { g -= GROUPS; rp++; }
        }
      }
      int rp = rp0, g = g0;
#pragma unroll
      for (int k = 0; k < PER; k++) {
        if (rp < RPS) {
          float2 v[4 * NW];                                    // (row 2rp, row 2rp+1) per input column
#pragma unroll
          for (int q = 0; q < NW; q++) {
#pragma unroll
            for (int bsel = 0; bsel < 4; bsel++) {
              const float lo = __uint_as_float(__byte_perm(wa[k][q], 0x4B000000u, 0x7540 + bsel));
              const float hi = __uint_as_float(__byte_perm(wb[k][q], 0x4B000000u, 0x7540 + bsel));
              v[4 * q + bsel] = __fadd2_rn(f2(lo, hi), f2s(-8388608.f));
            }
          }
          float2 o[4];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int c = j + OFF + RD;
            float2 acc = __fmul2_rn(f2s(kc.wd[0]), v[c]);
#pragma unroll
            for (int t = 1; t <= RD; t++) acc = __ffma2_rn(f2s(kc.wd[t]), __fadd2_rn(v[c - t], v[c + t]), acc);
            o[j] = acc;
          }
          float *d = sR1 + (2 * rp) * C::R1_P + 4 * g;
          *reinterpret_cast<float4 *>(d) = make_float4(o[0].x, o[1].x, o[2].x, o[3].x);
          *reinterpret_cast<float4 *>(d + C::R1_P + 2) = make_float4(o[0].y, o[1].y, o[2].y, o[3].y);
        }
        g += DG; rp += DRP;
        if (g >= GROUPS) { g -= GROUPS; rp++; }
      }
      if ((tid & 31) == 0) sM[tid >> 5] = 255.f;               // largest |pixel| a u8 tile can hold (the bound's M; a byte-wise maximum
                                                               // costs more integer instructions than the bound gains in dark tiles)
    } else {
    // ---- stage A: tile (+12 halo) -> row-pair interleaved float2 tile --------------------------------
    if (!generic) {
      if (TMA) {
        // bytes are (or will be) in sU8: [IN_H][96]; a thread converts 2 pixels x 2 rows -> one 16-byte store,
        // consecutive threads = consecutive 16-byte words: conflict free
        mbar_wait(bar, phase);
        phase ^= 1;
        constexpr int CPR = C::IN_W / 2;                       // 44 column pairs
        constexpr int ITEMS = (C::IN_H / 2) * CPR;
        unsigned mx = 0;
        for (int it = tid; it < ITEMS; it += C::NT) {
          const int rp = it / CPR, cpi = it - rp * CPR;
          const unsigned a = *reinterpret_cast<const unsigned short *>(sU8 + (2 * rp) * C::U8_P + C::U8_SKIP + 2 * cpi);
          const unsigned b = *reinterpret_cast<const unsigned short *>(sU8 + (2 * rp + 1) * C::U8_P + C::U8_SKIP + 2 * cpi);
          mx = max(mx, max(max(a & 0xff, a >> 8), max(b & 0xff, b >> 8)));
          *reinterpret_cast<float4 *>(sINp + rp * C::IN_P + 2 * cpi) =
              make_float4((float)(a & 0xff), (float)(b & 0xff), (float)(a >> 8), (float)(b >> 8));
        }
        mloc = (float)mx;
      } else {
        constexpr int VW = C::IN_W / 4;                        // 22 vectors of 4 pixels per row
        constexpr int ITEMS = (C::IN_H / 2) * VW;              // (row pair, vector)
        constexpr int PER = (ITEMS + C::NT - 1) / C::NT;
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
          unsigned mx = 0;
#pragma unroll
          for (int k = 0; k < PER; k++) {
            const int it = tid + k * C::NT;
            if (it < ITEMS) {
              const int rp = it / VW, v = it - rp * VW;
              float4 *d = reinterpret_cast<float4 *>(sINp + rp * C::IN_P + 4 * v);
              d[0] = make_float4((float)(a[k] & 0xff), (float)(b[k] & 0xff), (float)((a[k] >> 8) & 0xff), (float)((b[k] >> 8) & 0xff));
              d[1] = make_float4((float)((a[k] >> 16) & 0xff), (float)((b[k] >> 16) & 0xff), (float)(a[k] >> 24), (float)(b[k] >> 24));
              mx = __vmaxu4(mx, __vmaxu4(a[k], b[k]));
            }
          }
          mloc = (float)max(max(mx & 0xff, (mx >> 8) & 0xff), max((mx >> 16) & 0xff, mx >> 24));
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
              mloc = fmaxf(mloc, fmaxf(fmaxf(fmaxf(fabsf(a[k].x), fabsf(a[k].y)), fmaxf(fabsf(a[k].z), fabsf(a[k].w))),
                                       fmaxf(fmaxf(fabsf(b[k].x), fabsf(b[k].y)), fmaxf(fabsf(b[k].z), fabsf(b[k].w)))));
            }
          }
        }
      }
    } else {
      // generic tile: scalar loads, reflected at the frame border (the blur's padding rule applied to the INPUT,
      // which makes every later plane of the tile the reference's plane at the reflected coordinate); all loads of a
      // thread are issued before its stores, so that the memory latency is paid once
      constexpr int N = C::IN_H * C::IN_W, PER = (N + C::NT - 1) / C::NT;
      float v[PER];
#pragma unroll
      for (int k = 0; k < PER; k++) {
        const int i = min(tid + k * C::NT, N - 1);
        const int r = i / C::IN_W, c = i - r * C::IN_W;
        const int gy = reflect_index(y0 - C::HALO + r, ny), gx = reflect_index(x0 - C::HALO + c, nx);
        if (U8) v[k] = (float)__ldg(static_cast<const unsigned char *>(frames) + fofs + (size_t)gy * nx + gx);
        else v[k] = __ldg(static_cast<const float *>(frames) + fofs + (size_t)gy * nx + gx);
      }
#pragma unroll
      for (int k = 0; k < PER; k++) {
        const int i = tid + k * C::NT;
        if (i < N) {
          const int r = i / C::IN_W, c = i - r * C::IN_W;
          reinterpret_cast<float *>(sINp + (r >> 1) * C::IN_P + c)[r & 1] = v[k];
          mloc = fmaxf(mloc, fabsf(v[k]));
        }
      }
      // product-coordinate remap tables: Is-tile index of the pixel whose gradient the reference uses there
      for (int q = tid; q < C::TW + 2 * RI; q += C::NT) {
        int gx = reflect_index(x0 - RI + q, nx);               // blur padding of the product planes
        gx = min(max(gx, 1), nx - 2);                          // gradient replicate rule
        mapx[q] = (short)min(max(gx - (x0 - C::G), 1), C::IS_W - 2);
      }
      for (int q = tid; q < C::TH + 2 * RI; q += C::NT) {
        int gy = reflect_index(y0 - RI + q, ny);
        gy = min(max(gy, 1), ny - 2);
        mapy[q] = (short)min(max(gy - (y0 - C::G), 1), C::IS_H - 2);
      }
    }
    for (int o = 16; o; o >>= 1) mloc = fmaxf(mloc, __shfl_xor_sync(0xffffffffu, mloc, o));
    if ((tid & 31) == 0) sM[tid >> 5] = mloc;                  // per-warp max |pixel|; combined in stage E
    __syncthreads();
    if (TMA) {   // the staging buffer is free again: request the bytes of this CTA's next tile
      const int nt = tile + tile_step;
      if (tid == 0 && nt < n_tiles) {
        int tx0, ty0, tf;
        decode(nt, tx0, ty0, tf);
        if (!generic_all && harris3_tile_is_interior<C>(tx0, ty0, nx, ny)) {
          mbar_expect_tx(bar, C::U8_BYTES);
          tma_load_3d(sU8, &tmap, bar, tx0 - C::U8_X0, ty0 - C::HALO, tf);
        }
      }
    }

    // ---- stage B: row blur sigma_d, packed over row pairs -> sR1 (plain, odd rows +2) ---------------
    {
      constexpr int GROUPS = C::R1_W / 4;                      // 20
      constexpr int OFF = C::HALO - C::G - RD;                 // first tap of R1 col 0 sits at IN col OFF
      constexpr int NP = 4 + 2 * RD + OFF;                     // positions loaded from the 4-aligned start
      constexpr int NL = (NP + 1) / 2 * 2;
      constexpr int RPS = C::R1_H / 2;                         // row pairs; consecutive threads = consecutive row pairs
      constexpr int NIT = (RPS * GROUPS + C::NT - 1) / C::NT;
      // software pipelined: the loads of item k+1 are issued before the stores of item k (the two tiles do not alias)
      float2 v[NL];
      if (tid < RPS * GROUPS) {
        const int g = tid / RPS, rp = tid - g * RPS;
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
          *reinterpret_cast<float4 *>(d + C::R1_P + 2) = make_float4(o[0].y, o[1].y, o[2].y, o[3].y);
        }
      }
    }
    }   // separate stages A and B
    __syncthreads();

    // ---- stage C: column blur sigma_d, packed over column pairs -> sISp (row-pair interleaved) -------
    {
      constexpr int CP = C::IS_W / 2;                          // 40 column pairs
      constexpr int RB = C::C_RB;                              // Is rows per item (even)
      constexpr int RGS = (C::IS_H + RB - 1) / RB;
      constexpr int OFFR = C::HALO - C::G - RD;                // Is row 0 (global y0-G) uses R1 rows OFFR .. OFFR+2RD
      static_assert(OFFR >= 0 && RB % 2 == 0, "R1 must cover the taps of Is row 0; item rows start even");
      // items: first the column pairs 0..31 of every row group (a warp = one row group: conflict-free 8-byte loads),
      // then the remaining CP-32 column pairs (a warp spans several row groups there: a small conflicted tail)
      constexpr int MAIN = 32 * RGS, TAILW = CP - 32;
      for (int it = tid; it < CP * RGS; it += C::NT) {
        int rg, cp;
        if (it < MAIN) { rg = it >> 5; cp = it & 31; }
        else { const int u = it - MAIN; rg = u / TAILW; cp = 32 + (u - rg * TAILW); }
        const int row0 = rg * RB + OFFR;                       // rg * RB is even: the parity of row0 + q is known per q
        constexpr bool OVERRUN = RGS * RB + OFFR + 2 * RD > C::R1_H;   // does the last group read past the tile?
        const float *col = sR1 + row0 * C::R1_P + 2 * cp;
        float2 acc[RB];
#pragma unroll
        for (int j = 0; j < RB; j++) acc[j] = f2s(0.f);
#pragma unroll
        for (int q = 0; q < RB + 2 * RD; q++) {
          float2 v;
          if (OVERRUN) {
            const int row = min(row0 + q, C::R1_H - 1);        // clamp (the overrun rows feed outputs that are not stored)
            v = *reinterpret_cast<const float2 *>(sR1 + row * C::R1_P + ((row & 1) << 1) + 2 * cp);
          } else {
            v = *reinterpret_cast<const float2 *>(col + q * C::R1_P + (((OFFR + q) & 1) << 1));
          }
#pragma unroll
          for (int j = 0; j < RB; j++) {
            const int t = q - j - RD;
            if (t >= -RD && t <= RD) acc[j] = __ffma2_rn(f2s(kc.wd[t < 0 ? -t : t]), v, acc[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < RB; j += 2) {
          const int i = rg * RB + j;                           // Is row (even)
          if (i < C::IS_H)
            *reinterpret_cast<float4 *>(sISp + (i >> 1) * C::IS_P + 2 * cp) = make_float4(acc[j].x, acc[j + 1].x, acc[j].y, acc[j + 1].y);
        }
      }
    }
    __syncthreads();

    // ---- stage D: gradient + products + row blur sigma_i, streamed -> sAR (plain, odd rows +2) -------
    // One item = one AR row pair x DW output columns.  AR rows (2r, 2r+1) <-> Is rows (2r+1, 2r+2); with Is
    // rows paired (even, odd), the two pair lines L0 = (2r, 2r+1) and L1 = (2r+2, 2r+3) hold everything:
    //     gy = (Is[2r+2] - Is[2r], Is[2r+3] - Is[2r+1]) = L1 - L0 (packed),   gx from L0.y and L1.x.
    // Consecutive threads take consecutive row pairs (line stride = 4 banks: a quarter warp of 16-byte loads
    // covers all 32 banks); row pairs are padded to a multiple of 8 so that quarter warps never straddle groups.
    {
      constexpr int DW = C::DW;
      constexpr int GROUPS = C::TW / DW;
      constexpr int NPOS = DW + 2 * RI;                        // product positions per item
      constexpr int RPS = C::AR_H / 2, RPS_PAD = (RPS + 7) & ~7;
      constexpr int PL = C::AR_H * C::AR_P;                    // plane stride (multiple of 4 floats)
      static_assert(PL % 4 == 0, "plane stride keeps 16-byte alignment");
      static_assert(RPS + 1 <= C::IS_H / 2, "row pairs and their two pair lines");
      auto store_item = [&](int g, int r, const float2 (&a0)[DW], const float2 (&b0)[DW], const float2 (&c0)[DW]) {
        float *o = sAR + (2 * r) * C::AR_P + DW * g;
#define ST_ROWPAIR(dst, v)                                                                          \
        *reinterpret_cast<float4 *>(dst) = make_float4(v[0].x, v[1].x, v[2].x, v[3].x);             \
        *reinterpret_cast<float4 *>((dst) + C::AR_P + 2) = make_float4(v[0].y, v[1].y, v[2].y, v[3].y); \
        if (DW == 8) {                                                                              \
          *reinterpret_cast<float4 *>((dst) + 4) = make_float4(v[DW - 4].x, v[DW - 3].x, v[DW - 2].x, v[DW - 1].x); \
          *reinterpret_cast<float4 *>((dst) + C::AR_P + 6) = make_float4(v[DW - 4].y, v[DW - 3].y, v[DW - 2].y, v[DW - 1].y); \
        }
        ST_ROWPAIR(o, a0) ST_ROWPAIR(o + PL, b0) ST_ROWPAIR(o + 2 * PL, c0)
#undef ST_ROWPAIR
      };
      // regular item: every product position lies inside the frame's gradient domain -> sliding windows over Is
      auto fast_item = [&](int g, int r) {
        float2 a0[DW], b0[DW], c0[DW];
#pragma unroll
        for (int j = 0; j < DW; j++) { a0[j] = b0[j] = c0[j] = f2s(0.f); }
        const float2 *l0 = sISp + r * C::IS_P + DW * g;      // pair line r   = Is rows 2r,   2r+1
        const float2 *l1 = l0 + C::IS_P;                     // pair line r+1 = Is rows 2r+2, 2r+3
        float2 p0[3], p1[3];                                 // sliding window of Is columns q, q+1, q+2
        {
          const float4 t0 = *reinterpret_cast<const float4 *>(l0), t1 = *reinterpret_cast<const float4 *>(l1);
          p0[0] = f2(t0.x, t0.y); p0[1] = f2(t0.z, t0.w); p1[0] = f2(t1.x, t1.y); p1[1] = f2(t1.z, t1.w);
        }
#pragma unroll
        for (int q = 0; q < NPOS; q++) {
          float2 n0 = f2s(0.f), n1 = f2s(0.f);
          if ((q & 1) == 0) {                                // columns q+2, q+3 arrive as one 16-byte load per line
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
          for (int j = 0; j < DW; j++) {
            const int t = q - j - RI;
            if (t >= -RI && t <= RI) {
              const float2 w = f2s(kc.wir[t < 0 ? -t : t]);
              a0[j] = __ffma2_rn(w, pa, a0[j]); b0[j] = __ffma2_rn(w, pb, b0[j]); c0[j] = __ffma2_rn(w, pc, c0[j]);
            }
          }
          // slide: (q, q+1, q+2) -> (q+1, q+2, q+3)
          p0[0] = p0[1]; p0[1] = p0[2]; p1[0] = p1[1]; p1[1] = p1[2];
          if ((q & 1) == 0) { p0[2] = n0; p1[2] = n1; }
        }
        store_item(g, r, a0, b0, c0);
      };
      // irregular item of a border tile: the product at virtual position (row a, column q) is the product at the
      // remapped pixel (reflect padding of the product planes, then the gradient's replicate rule)
      auto remapped_item = [&](int g, int r) {
        float2 a0[DW], b0[DW], c0[DW];
#pragma unroll
        for (int j = 0; j < DW; j++) { a0[j] = b0[j] = c0[j] = f2s(0.f); }
        auto IS = [&](int row, int col) -> float { return reinterpret_cast<const float *>(sISp + (row >> 1) * C::IS_P + col)[row & 1]; };
        const int ry0 = mapy[2 * r], ry1 = mapy[2 * r + 1];
#pragma unroll 2
        for (int q = 0; q < NPOS; q++) {
          const int rx = mapx[DW * g + q];
          float2 gx, gy;
          if (GRAD == 0) {
            gx = f2(IS(ry0, rx + 1) - IS(ry0, rx - 1), IS(ry1, rx + 1) - IS(ry1, rx - 1));
            gy = f2(IS(ry0 + 1, rx) - IS(ry0 - 1, rx), IS(ry1 + 1, rx) - IS(ry1 - 1, rx));
          } else {
            gx = f2(fmaf(0.25f, IS(ry0, rx + 1) - IS(ry0, rx - 1), 0.125f * (IS(ry0 - 1, rx + 1) + IS(ry0 + 1, rx + 1) - IS(ry0 - 1, rx - 1) - IS(ry0 + 1, rx - 1))),
                    fmaf(0.25f, IS(ry1, rx + 1) - IS(ry1, rx - 1), 0.125f * (IS(ry1 - 1, rx + 1) + IS(ry1 + 1, rx + 1) - IS(ry1 - 1, rx - 1) - IS(ry1 + 1, rx - 1))));
            gy = f2(fmaf(0.25f, IS(ry0 + 1, rx) - IS(ry0 - 1, rx), 0.125f * (IS(ry0 + 1, rx + 1) + IS(ry0 + 1, rx - 1) - IS(ry0 - 1, rx + 1) - IS(ry0 - 1, rx - 1))),
                    fmaf(0.25f, IS(ry1 + 1, rx) - IS(ry1 - 1, rx), 0.125f * (IS(ry1 + 1, rx + 1) + IS(ry1 + 1, rx - 1) - IS(ry1 - 1, rx + 1) - IS(ry1 - 1, rx - 1))));
          }
          const float2 pa = __fmul2_rn(gx, gx), pb = __fmul2_rn(gx, gy), pc = __fmul2_rn(gy, gy);
#pragma unroll
          for (int j = 0; j < DW; j++) {
            const int t = q - j - RI;
            if (t >= -RI && t <= RI) {
              const float2 w = f2s(kc.wir[t < 0 ? -t : t]);
              a0[j] = __ffma2_rn(w, pa, a0[j]); b0[j] = __ffma2_rn(w, pb, b0[j]); c0[j] = __ffma2_rn(w, pc, c0[j]);
            }
          }
        }
        store_item(g, r, a0, b0, c0);
      };
      // Regular ranges of a border tile: group g is regular in x iff all its NPOS product columns x0-RI+DW*g+q lie in
      // [1, nx-2]; row pair r iff its two product rows y0-RI+2r, +1 lie in [1, ny-2].  Interior tiles: everything.
      int gb0 = 0, gb1 = GROUPS, rb0 = 0, rb1 = RPS;
      if (generic) {
        gb0 = max(0, (RI + 1 - x0 + DW - 1) / DW);
        const int gx_hi = nx - 2 - (NPOS - 1) - x0 + RI;          // largest regular DW*g
        gb1 = gx_hi < 0 ? 0 : min(GROUPS, gx_hi / DW + 1);
        rb0 = max(0, (RI + 1 - y0 + 1) / 2);
        const int ry_hi = ny - 3 - y0 + RI;                        // largest regular 2r
        rb1 = ry_hi < 0 ? 0 : min(RPS, ry_hi / 2 + 1);
        if (gb1 < gb0) gb1 = gb0;
        if (rb1 < rb0) rb1 = rb0;
      }
      for (int it = tid; it < RPS_PAD * GROUPS; it += C::NT) {
        const int g = it / RPS_PAD, r = it - g * RPS_PAD;
        if (r >= rb1 || r < rb0 || g < gb0 || g >= gb1) continue;
        fast_item(g, r);
      }
      if (generic) {   // the irregular items, densely packed over the threads: whole groups first, then the irregular row pairs of regular groups
        const int ng = gb1 - gb0, nbg = GROUPS - ng, nbr = RPS - (rb1 - rb0);
        const int n_irr = nbg * RPS + ng * nbr;
        for (int j = tid; j < n_irr; j += C::NT) {
          int g, r;
          if (j < nbg * RPS) {
            const int gi = j / RPS;
            r = j - gi * RPS;
            g = gi < gb0 ? gi : gb1 + (gi - gb0);
          } else {
            const int j2 = j - nbg * RPS, ri = j2 / ng;
            g = gb0 + (j2 - ri * ng);
            r = ri < rb0 ? ri : rb1 + (ri - rb0);
          }
          remapped_item(g, r);
        }
      }
    }
    __syncthreads();

    // ---- stage E: column blur sigma_i, packed over column pairs, + corner measure + store + eps -----
    {
      constexpr int RB = C::E_RB;
      constexpr int CP = C::TW / 2;                            // 32 column pairs: one warp = one row group
      constexpr int PL = C::AR_H * C::AR_P;
      float Mtile = 0.f;
#pragma unroll
      for (int w = 0; w < C::NT / 32; w++) Mtile = fmaxf(Mtile, sM[w]);
      for (int it = tid; it < CP * C::E_NRG; it += C::NT) {
        const int rg = it / CP, cp = it - rg * CP;
        // the last groups may overlap when TH is not a multiple of RB; otherwise r0 is even at compile time (row parity known)
        const int r0 = (C::TH % RB == 0) ? rg * RB : min(rg * RB, C::TH - RB);
        float2 aa[RB], ab[RB], ac[RB];
#pragma unroll
        for (int j = 0; j < RB; j++) { aa[j] = f2s(0.f); ab[j] = f2s(0.f); ac[j] = f2s(0.f); }
#pragma unroll
        for (int q = 0; q < RB + 2 * RI; q++) {                // output row j uses AR rows j .. j+2RI
          const int row = r0 + q;
          const float *p = sAR + row * C::AR_P + ((row & 1) << 1) + 2 * cp;
          const float2 va = *reinterpret_cast<const float2 *>(p);
          const float2 vb = *reinterpret_cast<const float2 *>(p + PL);
          const float2 vc = *reinterpret_cast<const float2 *>(p + 2 * PL);
#pragma unroll
          for (int j = 0; j < RB; j++) {
            const int t = q - j - RI;
            if (t >= -RI && t <= RI) {
              const float2 w = f2s(kc.wic[t < 0 ? -t : t]);
              aa[j] = __ffma2_rn(w, va, aa[j]); ab[j] = __ffma2_rn(w, vb, ab[j]); ac[j] = __ffma2_rn(w, vc, ac[j]);
            }
          }
        }
        const int gx = x0 + 2 * cp, gy0 = y0 + r0;
        float *dst = Rout + fofs + (size_t)gy0 * nx + gx;
        float2 rr[RB];
        float tmax = 0.f;                                      // largest trace of the item's 2 x RB pixels
#pragma unroll
        for (int j = 0; j < RB; j++) {
          const float2 tr = __fadd2_rn(aa[j], ac[j]);
          if (kc.measure == 0) {                               // Harris: (A*C - B*B) - (k*tr)*tr, each op rounded (harris.cpp:100-103)
            const float2 det = sub2(__fmul2_rn(aa[j], ac[j]), __fmul2_rn(ab[j], ab[j]));
            rr[j] = sub2(det, __fmul2_rn(__fmul2_rn(f2s(kc.k), tr), tr));
            if (kc.tr_cut > 0.f) {                             // certainly below the threshold (harris_trace_cut)
              rr[j].x = tr.x < kc.tr_cut ? -3.402823466e+38f : rr[j].x;
              rr[j].y = tr.y < kc.tr_cut ? -3.402823466e+38f : rr[j].y;
            }
          } else {
            rr[j] = f2(corner_measure(aa[j].x, ab[j].x, ac[j].x, kc.k, kc.measure), corner_measure(aa[j].y, ab[j].y, ac[j].y, kc.k, kc.measure));
          }
          tmax = fmaxf(tmax, fmaxf(tr.x, tr.y));
        }
        if (!generic) {
#pragma unroll
          for (int j = 0; j < RB; j++) *reinterpret_cast<float2 *>(dst + (size_t)j * nx) = rr[j];
        } else {
#pragma unroll
          for (int j = 0; j < RB; j++)
            if (gy0 + j < ny) {
              if (gx < nx) dst[(size_t)j * nx] = rr[j].x;
              if (gx + 1 < nx) dst[(size_t)j * nx + 1] = rr[j].y;
            }
        }
        if (eps_blk) {
          // The item's RB <= 8 rows touch at most two 8-row eps blocks; both receive the bound of the item's largest trace
          // (a block's bound may then include a few rows of its neighbour: larger, never smaller).  4 lanes = 8 columns =
          // one eps block column; non-negative floats order like their bit patterns.
          tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
          tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 2));
          const int bx = gx >> 3, blk0 = gy0 >> 3, blk1 = (gy0 + RB - 1) >> 3;
          if ((cp & 3) == 0 && bx < ebx) {
            const unsigned e32 = __float_as_uint(harris_eps(tmax, Mtile, kc.k));
            unsigned *e = eps_blk + ((size_t)frame * eby + blk0) * ebx + bx;
            if (blk0 < eby) atomicMax(e, e32);
            if (blk1 != blk0 && blk1 < eby) atomicMax(e + ebx, e32);
          }
        }
      }
    }
    if (TMA) __syncthreads();                                  // all of sAR / sISp consumed before the next tile's stage A
  }
}

}  // namespace b2f
