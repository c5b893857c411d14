// harris_fused.cu — launcher of the fused Harris response kernel (harris_kernels3.cuh): picks the tile
// configuration, builds the TMA tensor map of the u8 frames and launches ONE kernel per batch.
#include "harris_kernels3.cuh"
#include "harris_host.h"
#include <cmath>

namespace b2f {

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library links cudart only)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// u8 frames [n][ny][nx] as a 3-D tensor, box = 96 x in_h x 1 bytes; out-of-range bytes are zero filled (never used:
// only tiles whose halo lies inside the frame go through the map)
static bool make_u8_tile_map(CUtensorMap *map, const void *frames, int n_frames, int nx, int ny, int in_h) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)n_frames};
  const cuuint64_t strides[2] = {(cuuint64_t)nx, (cuuint64_t)nx * ny};       // bytes, dims 1 and 2
  const cuuint32_t box[3] = {96u, (cuuint32_t)in_h, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(frames), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <class C, bool U8, int GRAD, bool TMA>
static int launch_cfg(b2f_ctx *ctx, const void *d_frames, int n_frames, int nx, int ny, float *d_R, unsigned *d_eps,
                      int generic_all, const HarrisConsts &kc, cudaStream_t st) {
  auto kern = harris_fused3_kernel<C, U8, GRAD, TMA>;
  const size_t smem = TMA ? C::SMEM_TMA : C::SMEM;
  // function attributes are per device: set on every launch (a few hundred ns), never cached per process
  B2F_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  const int tiles_x = ceil_div(nx, C::TW), tiles_y = ceil_div(ny, C::TH);
  if (TMA) {
    if (!make_u8_tile_map(&map, d_frames, n_frames, nx, ny, C::IN_H)) { set_error("harris: cuTensorMapEncodeTiled failed"); return B2F_ECUDA; }
    const long long tiles = (long long)tiles_x * tiles_y * n_frames;
    const int per_sm = C::CTAS;
    const int grid = (int)std::min<long long>(tiles, (long long)ctx->sm_count * per_sm);
    kern<<<grid, C::NT, smem, st>>>(d_frames, d_R, d_eps, nx, ny, n_frames, generic_all, kc, map);
  } else {
    kern<<<dim3(tiles_x, tiles_y, n_frames), C::NT, smem, st>>>(d_frames, d_R, d_eps, nx, ny, n_frames, generic_all, kc, map);
  }
  B2F_LAUNCH_CHECK(ctx);
  return B2F_OK;
}

static int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

template <class C, bool ALLOW_TMA>
static int launch_shape(b2f_ctx *ctx, const void *d_frames, bool u8, int grad, int n_frames, int nx, int ny, float *d_R,
                        unsigned *d_eps, int generic_all, bool tma, const HarrisConsts &kc, cudaStream_t st) {
  if (u8) {
    if (ALLOW_TMA && tma)
      return grad ? launch_cfg<C, true, 1, ALLOW_TMA>(ctx, d_frames, n_frames, nx, ny, d_R, d_eps, generic_all, kc, st)
                  : launch_cfg<C, true, 0, ALLOW_TMA>(ctx, d_frames, n_frames, nx, ny, d_R, d_eps, generic_all, kc, st);
    return grad ? launch_cfg<C, true, 1, false>(ctx, d_frames, n_frames, nx, ny, d_R, d_eps, generic_all, kc, st)
                : launch_cfg<C, true, 0, false>(ctx, d_frames, n_frames, nx, ny, d_R, d_eps, generic_all, kc, st);
  }
  return grad ? launch_cfg<C, false, 1, false>(ctx, d_frames, n_frames, nx, ny, d_R, d_eps, generic_all, kc, st)
              : launch_cfg<C, false, 0, false>(ctx, d_frames, n_frames, nx, ny, d_R, d_eps, generic_all, kc, st);
}

bool harris_fused_supported(int nx, int ny, float sigma_d, float sigma_i, int gaussian) {
  if (gaussian != 0) return false;
  if (sigma_d <= 0 || sigma_i <= 0) return false;
  int rd = (int)(3 * sigma_d), ri = (int)(3 * sigma_i);
  if (!(rd == 3 && (ri == 7 || ri == 3))) return false;
  // frames must be large enough that every reflection stays inside its own tile window
  return nx >= 32 && ny >= 32 && (long long)nx * ny < (1ll << 31);
}

int harris_taps_double(float sigma, double *B) {     // gaussian.cpp:306-329
  int size = (int)(3 * sigma) + 1;
  if (size > HARRIS_MAX_TAPS) return -1;
  float den_f = 2 * sigma * sigma;
  double den = den_f, s = sigma;
  for (int i = 0; i < size; i++) B[i] = 1 / (s * sqrt(2.0 * 3.1415926)) * exp(-i * i / den);
  double norm = 0;
  for (int i = 0; i < size; i++) norm += B[i];
  norm *= 2;
  norm -= B[0];
  for (int i = 0; i < size; i++) B[i] /= norm;
  return size;
}

// Fused response for a batch resident on the device.  d_eps (may be NULL) receives the per-8x8-block error bound
// (it must be zero filled: blocks are combined with atomicMax).  corners_only: the plane feeds the certified corner path
// only, so pixels certainly below the threshold may be stored as -FLT_MAX (harris_trace_cut).
int harris_fused_launch(b2f_ctx *ctx, const void *d_frames, bool u8, int n_frames, int nx, int ny, const b2f_harris_params *p,
                        float *d_R, unsigned *d_eps, bool corners_only, cudaStream_t st) {
  HarrisConsts kc;
  memset(&kc, 0, sizeof(kc));
  double Bd[HARRIS_MAX_TAPS], Bi[HARRIS_MAX_TAPS];
  const int sd = harris_taps_double(p->sigma_d, Bd), si = harris_taps_double(p->sigma_i, Bi);
  const float gscale = (p->gradient == 1) ? 1.f : 0.25f;
  for (int i = 0; i < sd; i++) kc.wd[i] = (float)Bd[i];
  for (int i = 0; i < si; i++) { kc.wic[i] = (float)Bi[i]; kc.wir[i] = gscale * (float)Bi[i]; }
  kc.k = p->k;
  kc.measure = p->measure;
  kc.tr_cut = corners_only ? harris_trace_cut(p->threshold, p->k, p->measure, u8) : 0.f;
  const int ri = si - 1, grad = p->gradient == 1;
  // vector loads / stores need 4-pixel aligned rows and aligned base pointers; anything else runs every tile generic
  const bool aligned = (nx % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_frames) & 15) == 0) && ((reinterpret_cast<uintptr_t>(d_R) & 7) == 0);
  const int generic_all = aligned ? 0 : 1;
  if (ri == 3) {
    using C = Fused3Cfg<3, 3, 64, 256, 4, 2>;
    return launch_shape<C, false>(ctx, d_frames, u8, grad, n_frames, nx, ny, d_R, d_eps, generic_all, false, kc, st);
  }
  // Measured on 16 x 3840x2160 u8 frames (tools/harris_timing.py): 64x48 tiles with 3 CTAs / SM 61 us per
  // frame, 64x64 tiles with 2 CTAs / SM 64 us, the same with TMA-staged input in a persistent grid 78 us (the resident
  // CTAs of an SM then walk their tiles in lock step and stop overlapping each other's stages).  Default: 64x48 / 3 CTAs;
  // B2F_HARRIS_TILE=64 and B2F_HARRIS_TMA=1 select the alternatives (kept because the tests cover every shape).
  const int cfg_env = env_int("B2F_HARRIS_TILE", 0), tma_env = env_int("B2F_HARRIS_TMA", 0);
  const bool tall = cfg_env != 64;
  const bool tma_ok = u8 && aligned && (nx % 16 == 0) && (((size_t)nx * ny) % 16 == 0) && encode_tiled_fn() != nullptr;   // 16-byte row strides
  const bool tma = tma_ok && tma_env != 0;
  if (tall) {
    using C = Fused3Cfg<3, 7, 48, 256, 4, 3>;
    return launch_shape<C, false>(ctx, d_frames, u8, grad, n_frames, nx, ny, d_R, d_eps, generic_all, false, kc, st);
  }
  using C = Fused3Cfg<3, 7, 64, 256, 4, 2>;
  return launch_shape<C, true>(ctx, d_frames, u8, grad, n_frames, nx, ny, d_R, d_eps, generic_all, tma, kc, st);
}

}  // namespace b2f
