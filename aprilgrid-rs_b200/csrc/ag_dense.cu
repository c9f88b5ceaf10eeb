// Dense per-pixel stages of the aprilgrid front end, sm_100a.
//
//   K1  k_blur_hessian_*   gray conversion -> 7-tap separable Gaussian (sigma 1.5) ->
//                          3x3 determinant-of-Hessian -> per-frame min
//                          reference: src/detector.rs:409-418, src/image_util.rs:72-206
//   K2  k_threshold_mask   mask = resp < 0.05 * min, packed 32 px / word
//                          reference: src/detector.rs:418, :176-177 (the `v < threshold` test)
//
// All arithmetic is bit-exact with the reference: products and sums are rounded separately
// (no FMA), taps are accumulated in tap order 0..6 starting from 0, and the taps themselves
// are the glibc-expf values the reference computes at run time.
#include "ag_common.cuh"
#include "ag_kernels.h"

namespace ag {

// Normalised Gaussian taps for sigma = 1.5 (src/image_util.rs:111-124, f32::exp = expf).
// tests/test_oracle_pins.py checks these bit patterns against the oracle's run-time taps.
__constant__ float c_taps[kBlurTaps] = {
    0x1.2c18a6p-5f /*0x3d160c53*/, 0x1.c7ce56p-4f /*0x3de3e72b*/, 0x1.bbe4fap-3f /*0x3e5df27d*/,
    0x1.152db4p-2f /*0x3e8a96da*/, 0x1.bbe4fap-3f, 0x1.c7ce56p-4f, 0x1.2c18a6p-5f};

// ---- gray conversion (image 0.25: to_luma32f) ------------------------------------------
// v / 255 and v / 65535 correctly rounded, as two FP ops: with r_hi + r_lo = 1/max split in
// two floats, fma(v, r_hi, v * r_lo) equals RN(v / max) for every u8 / u16 v (checked
// exhaustively by tests/test_gpu_dense.py::test_unorm_conversion_exhaustive).
AG_D float unorm8_to_f32(float v) {
  const float rh = __uint_as_float(0x3b808081u), rl = __uint_as_float(0xaf7efeffu);
  return __fmaf_rn(v, rh, __fmul_rn(v, rl));
}
AG_D float unorm16_to_f32(float v) {
  const float rh = __uint_as_float(0x37800080u), rl = __uint_as_float(0x27800080u);
  return __fmaf_rn(v, rh, __fmul_rn(v, rl));
}
// image::color::rgb_to_luma on u8: integer sRGB weights, truncating division by 10000.
AG_D uint32_t rgb_luma_u8(uint32_t r, uint32_t g, uint32_t b) {
  return (2126u * r + 7152u * g + 722u * b) / 10000u;
}

template <int FMT>
AG_D float load_luma(const uint8_t* __restrict__ frame, size_t row_stride, int x, int y) {
  const uint8_t* row = frame + (size_t)y * row_stride;
  if (FMT == AG_L8) {
    return unorm8_to_f32((float)row[x]);
  } else if (FMT == AG_L16) {
    return unorm16_to_f32((float)reinterpret_cast<const uint16_t*>(row)[x]);
  } else {
    const uint8_t* p = row + 3 * x;
    return unorm8_to_f32((float)rgb_luma_u8(p[0], p[1], p[2]));
  }
}

AG_D float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// -----------------------------------------------------------------------------------------
// K1, generic tile version (any format, any size, any stride).  One CTA = 64 x 32 outputs.
// Used for L16 / RGB8 inputs and as the cross-check of the streaming L8 kernel.
// -----------------------------------------------------------------------------------------
constexpr int T_W = 64, T_H = 32;
constexpr int IN_W = T_W + 8, IN_H = T_H + 8;    // input halo: 3 (blur) + 1 (Hessian)
constexpr int TMP_W = T_W + 2, TMP_H = T_H + 8;  // H-pass output
constexpr int BL_W = T_W + 2, BL_H = T_H + 2;    // blurred tile incl. Hessian halo
constexpr int TMP_LD = TMP_W + 1, BL_LD = BL_W + 1;

template <int FMT, bool WRITE_BLUR>
__global__ void __launch_bounds__(256)
k_blur_hessian_tile(const uint8_t* __restrict__ frames, FrameGeom g, float* __restrict__ blur,
                    float* __restrict__ resp, uint32_t* __restrict__ frame_min) {
  __shared__ float s_in[IN_H][IN_W];
  __shared__ float s_tmp[TMP_H][TMP_LD];
  __shared__ float s_bl[BL_H][BL_LD];
  __shared__ float s_red[8];

  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * T_W, y0 = blockIdx.y * T_H;
  const int f = blockIdx.z;
  const uint8_t* frame = frames + (size_t)f * g.frame_stride;

  // 1. gray conversion of the clamped (edge-replicated) input window
  for (int i = tid; i < IN_H * IN_W; i += 256) {
    int iy = i / IN_W, ix = i - iy * IN_W;
    int gx = min(max(x0 - 4 + ix, 0), g.w - 1);
    int gy = min(max(y0 - 4 + iy, 0), g.h - 1);
    s_in[iy][ix] = load_luma<FMT>(frame, g.row_stride, gx, gy);
  }
  __syncthreads();
  // 2. horizontal pass: val = 0; val += px * k[i], i = 0..6   (image_util.rs:138-185)
  for (int i = tid; i < TMP_H * TMP_W; i += 256) {
    int iy = i / TMP_W, jx = i - iy * TMP_W;
    float val = 0.0f;
#pragma unroll
    for (int t = 0; t < kBlurTaps; ++t) val = __fadd_rn(val, __fmul_rn(s_in[iy][jx + t], c_taps[t]));
    s_tmp[iy][jx] = val;
  }
  __syncthreads();
  // 3. vertical pass: out = 0; out += temp[ky] * k[i]           (image_util.rs:188-203)
  for (int i = tid; i < BL_H * BL_W; i += 256) {
    int jy = i / BL_W, jx = i - jy * BL_W;
    float val = 0.0f;
#pragma unroll
    for (int t = 0; t < kBlurTaps; ++t) val = __fadd_rn(val, __fmul_rn(s_tmp[jy + t][jx], c_taps[t]));
    s_bl[jy][jx] = val;
  }
  __syncthreads();
  // 4. Hessian response on the interior, 0 on the image border  (image_util.rs:83-106)
  float mn = 3.40282347e+38f;
  for (int i = tid; i < T_H * T_W; i += 256) {
    int ty = i / T_W, tx = i - ty * T_W;
    int gx = x0 + tx, gy = y0 + ty;
    if (gx >= g.w || gy >= g.h) continue;
    float r = 0.0f;
    if (gx >= 1 && gx < g.w - 1 && gy >= 1 && gy < g.h - 1) {
      float v11 = s_bl[ty][tx], v12 = s_bl[ty][tx + 1], v13 = s_bl[ty][tx + 2];
      float v21 = s_bl[ty + 1][tx], v22 = s_bl[ty + 1][tx + 1], v23 = s_bl[ty + 1][tx + 2];
      float v31 = s_bl[ty + 2][tx], v32 = s_bl[ty + 2][tx + 1], v33 = s_bl[ty + 2][tx + 2];
      float t2 = __fmul_rn(v22, 2.0f);
      float lxx = __fadd_rn(__fsub_rn(v21, t2), v23);
      float lyy = __fadd_rn(__fsub_rn(v12, t2), v32);
      float lxy = __fmul_rn(__fsub_rn(__fadd_rn(__fsub_rn(v13, v11), v31), v33), 0.25f);
      r = __fsub_rn(__fmul_rn(lxx, lyy), __fmul_rn(lxy, lxy));
    }
    size_t o = (size_t)f * g.n_px + (size_t)gy * g.w + gx;
    resp[o] = r;
    if (WRITE_BLUR) blur[o] = s_bl[ty + 1][tx + 1];
    mn = fminf(mn, r);
  }
  // 5. per-frame min (detector.rs:414-417)
  mn = warp_min(mn);
  if ((tid & 31) == 0) s_red[tid >> 5] = mn;
  __syncthreads();
  if (tid < 32) {
    float v = tid < 8 ? s_red[tid] : 3.40282347e+38f;
    v = warp_min(v);
    if (tid == 0) atomicMin(&frame_min[f], float_to_ordered(v));
  }
}

__global__ void k_fill_u32(uint32_t* p, int n, uint32_t v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// -----------------------------------------------------------------------------------------
// K2: threshold.  One warp handles runs of 32 consecutive pixels of a row with one ballot
// per word; each thread keeps several independent 4-byte loads in flight.
// -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_threshold_mask(const float* __restrict__ resp, FrameGeom g, const uint32_t* __restrict__ frame_min,
                 uint32_t* __restrict__ mask) {
  const int f = blockIdx.y;
  const float thr = __fmul_rn(ordered_to_float(frame_min[f]), 0.05f);  // detector.rs:418
  const float* R = resp + (size_t)f * g.n_px;
  uint32_t* M = mask + (size_t)f * g.n_words;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  constexpr int U = 4;
  for (int wi0 = warp * U; wi0 < g.n_words; wi0 += n_warps * U) {
    float v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int wi = wi0 + u;
      v[u] = 3.40282347e+38f;
      if (wi < g.n_words) {
        int row = wi / g.wpr, wc = wi - row * g.wpr;
        int x = wc * 32 + lane;
        if (x < g.w) v[u] = __ldcs(R + (size_t)row * g.w + x);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      uint32_t b = __ballot_sync(0xffffffffu, v[u] < thr);
      if (lane == 0 && wi0 + u < g.n_words) M[wi0 + u] = b;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Standalone operators mirroring the reference's public image_util functions (f32 -> f32).
// ---------------------------------------------------------------------------------------
// image_util::gaussian_blur_f32(img, sigma): general radius, taps passed from the host.
__global__ void __launch_bounds__(256)
k_blur_f32_h(const float* __restrict__ in, float* __restrict__ tmp, int w, int h,
             const float* __restrict__ taps, int radius) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w || y >= h) return;
  const float* r = in + (size_t)y * w;
  float val = 0.0f;
  for (int i = 0; i <= 2 * radius; ++i) {
    int kx = min(max(x + i - radius, 0), w - 1);
    val = __fadd_rn(val, __fmul_rn(r[kx], taps[i]));
  }
  tmp[(size_t)y * w + x] = val;
}
__global__ void __launch_bounds__(256)
k_blur_f32_v(const float* __restrict__ tmp, float* __restrict__ out, int w, int h,
             const float* __restrict__ taps, int radius) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w || y >= h) return;
  float val = 0.0f;
  for (int i = 0; i <= 2 * radius; ++i) {
    int ky = min(max(y + i - radius, 0), h - 1);
    val = __fadd_rn(val, __fmul_rn(tmp[(size_t)ky * w + x], taps[i]));
  }
  out[(size_t)y * w + x] = val;
}
// image_util::hessian_response(img)
__global__ void __launch_bounds__(256)
k_hessian_f32(const float* __restrict__ img, float* __restrict__ out, int w, int h) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w || y >= h) return;
  float r = 0.0f;
  if (x >= 1 && x < w - 1 && y >= 1 && y < h - 1) {
    const float* p = img + (size_t)(y - 1) * w + x;
    const float* c = p + w;
    const float* n = c + w;
    float t2 = __fmul_rn(c[0], 2.0f);
    float lxx = __fadd_rn(__fsub_rn(c[-1], t2), c[1]);
    float lyy = __fadd_rn(__fsub_rn(p[0], t2), n[0]);
    float lxy = __fmul_rn(__fsub_rn(__fadd_rn(__fsub_rn(p[1], p[-1]), n[-1]), n[1]), 0.25f);
    r = __fsub_rn(__fmul_rn(lxx, lyy), __fmul_rn(lxy, lxy));
  }
  out[(size_t)y * w + x] = r;
}

// Exhaustive check helper for the unorm conversions (test only): out[v] for v in [0, n).
__global__ void k_unorm_table(float* out8, float* out16, float* ref8, float* ref16) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < 256) {
    out8[v] = unorm8_to_f32((float)v);
    ref8[v] = __fdiv_rn((float)v, 255.0f);
  }
  if (v < 65536) {
    out16[v] = unorm16_to_f32((float)v);
    ref16[v] = __fdiv_rn((float)v, 65535.0f);
  }
}

// ---------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------
template <int FMT>
static void launch_tile(const uint8_t* frames, const FrameGeom& g, int n_frames, float* blur,
                        float* resp, uint32_t* frame_min, bool write_blur, cudaStream_t s) {
  dim3 grid((g.w + T_W - 1) / T_W, (g.h + T_H - 1) / T_H, n_frames);
  if (write_blur)
    k_blur_hessian_tile<FMT, true><<<grid, 256, 0, s>>>(frames, g, blur, resp, frame_min);
  else
    k_blur_hessian_tile<FMT, false><<<grid, 256, 0, s>>>(frames, g, blur, resp, frame_min);
}

int launch_blur_hessian(const uint8_t* frames, const FrameGeom& g, int n_frames, float* blur,
                        float* resp, uint32_t* frame_min, bool write_blur, int variant,
                        cudaStream_t s) {
  int launches = 0;
  k_fill_u32<<<(n_frames + 255) / 256, 256, 0, s>>>(frame_min, n_frames, kOrderedFltMax);
  ++launches;
  (void)variant;
  switch (g.format) {
    case AG_L8: launch_tile<AG_L8>(frames, g, n_frames, blur, resp, frame_min, write_blur, s); break;
    case AG_L16: launch_tile<AG_L16>(frames, g, n_frames, blur, resp, frame_min, write_blur, s); break;
    default: launch_tile<AG_RGB8>(frames, g, n_frames, blur, resp, frame_min, write_blur, s); break;
  }
  return launches + 1;
}

int launch_threshold(const float* resp, const FrameGeom& g, int n_frames, const uint32_t* frame_min,
                     uint32_t* mask, cudaStream_t s) {
  int blocks_x = (g.n_words + 8 * 4 - 1) / (8 * 4);  // 8 warps x 4 words per block iteration
  if (blocks_x > 4096) blocks_x = 4096;
  if (blocks_x < 1) blocks_x = 1;
  dim3 grid(blocks_x, n_frames);
  k_threshold_mask<<<grid, 256, 0, s>>>(resp, g, frame_min, mask);
  return 1;
}

int launch_blur_f32(const float* in, float* tmp, float* out, int w, int h, const float* d_taps,
                    int radius, cudaStream_t s) {
  dim3 grid((w + 255) / 256, h);
  k_blur_f32_h<<<grid, 256, 0, s>>>(in, tmp, w, h, d_taps, radius);
  k_blur_f32_v<<<grid, 256, 0, s>>>(tmp, out, w, h, d_taps, radius);
  return 2;
}
int launch_hessian_f32(const float* in, float* out, int w, int h, cudaStream_t s) {
  dim3 grid((w + 255) / 256, h);
  k_hessian_f32<<<grid, 256, 0, s>>>(in, out, w, h);
  return 1;
}
int launch_unorm_table(float* out8, float* out16, float* ref8, float* ref16, cudaStream_t s) {
  k_unorm_table<<<256, 256, 0, s>>>(out8, out16, ref8, ref16);
  return 1;
}

}  // namespace ag
