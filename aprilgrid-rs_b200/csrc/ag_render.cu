// Synthetic AprilGrid frame generator (benchmark / test input, not part of the detect path).
//
// Board geometry follows scripts/generate_aprilgrid.py:1086-1167 of the reference repository:
// a cols x rows lattice of tags with side 1 separated by gaps of 0.3, a black square of side
// 0.3 at every lattice corner, tag ids row-major starting from the BOTTOM row, each tag
// (edge + 2*border)^2 cells with a black border and code bit "1" = white, MSB first, rows
// from the top of the tag.  Each frame views the board under a seeded random pose
// (in-plane rotation +-45 deg, tilt <= 30 deg about both axes, pinhole projection), is
// area-sampled 4x4 per pixel, and gets +-2-sigma-ish additive noise.  Paper = 200, ink = 30.
#include "ag_common.cuh"
#include "ag_kernels.h"

namespace ag {

struct Pose {
  float hinv[9];  // image (x, y, 1) -> page (X, Y, W), page units = tag sides
  float wb, hb;   // board extent in page units
};

__device__ __forceinline__ uint64_t splitmix64(uint64_t& s) {
  uint64_t z = (s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ float urand(uint64_t& s) {  // [0, 1)
  return (float)(splitmix64(s) >> 40) * (1.0f / 16777216.0f);
}

__device__ void invert3(const float* m, float* o) {
  float a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
  float A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
  float det = a * A + b * B + c * C;
  float r = 1.0f / det;
  o[0] = A * r; o[1] = -(b * i - c * h) * r; o[2] = (b * f - c * e) * r;
  o[3] = B * r; o[4] = (a * i - c * g) * r;  o[5] = -(a * f - c * d) * r;
  o[6] = C * r; o[7] = -(a * h - b * g) * r; o[8] = (a * e - b * d) * r;
}

__device__ void make_pose(uint64_t seed, int frame, int w, int h, int cols, int rows, Pose* P) {
  uint64_t s = seed * 0x2545f4914f6cdd1dull + (uint64_t)frame * 0x9e3779b97f4a7c15ull + 12345u;
  const float sp = 0.3f;
  const float wb = cols * (1.0f + sp) + sp, hb = rows * (1.0f + sp) + sp;
  const float alpha = (urand(s) * 2.0f - 1.0f) * 0.7853982f;  // +-45 deg
  const float bx = (urand(s) * 2.0f - 1.0f) * 0.5235988f;     // +-30 deg
  const float by = (urand(s) * 2.0f - 1.0f) * 0.5235988f;
  float tag_px = 60.0f + urand(s) * 60.0f;  // tag side at the board centre
  const float focal = 1.2f * (float)w;
  const float ca = cosf(alpha), sa = sinf(alpha), cx = cosf(bx), sx = sinf(bx), cy = cosf(by),
              sy = sinf(by);
  // R = Rz(alpha) * Rx(bx) * Ry(by); only the first two columns matter (board plane z = 0)
  float r00 = cy, r01 = 0.0f, r10 = sx * sy, r11 = cx, r20 = -cx * sy, r21 = sx;
  float q00 = ca * r00 - sa * r10, q01 = ca * r01 - sa * r11;
  float q10 = sa * r00 + ca * r10, q11 = sa * r01 + ca * r11;
  float q20 = r20, q21 = r21;
  float tx = 0.0f, ty = 0.0f;
  float H[9];
  for (int iter = 0; iter < 8; ++iter) {
    // page (X, Y) in tag sides, centred; camera depth = focal (unit magnification at centre)
    // x = tx + focal * t*(q00 X' + q01 Y') / (focal + t*(q20 X' + q21 Y')),  X' = X - wb/2
    const float t = tag_px;
    H[0] = focal * t * q00; H[1] = focal * t * q01; H[2] = -focal * t * (q00 * wb + q01 * hb) * 0.5f;
    H[3] = focal * t * q10; H[4] = focal * t * q11; H[5] = -focal * t * (q10 * wb + q11 * hb) * 0.5f;
    H[6] = t * q20;         H[7] = t * q21;         H[8] = focal - t * (q20 * wb + q21 * hb) * 0.5f;
    float minx = 1e9f, maxx = -1e9f, miny = 1e9f, maxy = -1e9f;
    for (int c = 0; c < 4; ++c) {
      float X = (c == 1 || c == 2) ? wb : 0.0f, Y = (c >= 2) ? hb : 0.0f;
      float W = H[6] * X + H[7] * Y + H[8];
      float x = (H[0] * X + H[1] * Y + H[2]) / W, y = (H[3] * X + H[4] * Y + H[5]) / W;
      minx = fminf(minx, x); maxx = fmaxf(maxx, x);
      miny = fminf(miny, y); maxy = fmaxf(maxy, y);
    }
    const float margin = 12.0f;
    const float bw = maxx - minx, bh = maxy - miny;
    if (bw > w - 2 * margin || bh > h - 2 * margin) {
      tag_px *= 0.9f * fminf((w - 2 * margin) / bw, (h - 2 * margin) / bh);
      continue;
    }
    tx = margin - minx + urand(s) * ((float)w - 2 * margin - bw);
    ty = margin - miny + urand(s) * ((float)h - 2 * margin - bh);
    break;
  }
  // translate: x' = x + tx  =>  row0 += tx * row2, row1 += ty * row2
  for (int c = 0; c < 3; ++c) {
    H[c] += tx * H[6 + c];
    H[3 + c] += ty * H[6 + c];
  }
  invert3(H, P->hinv);
  P->wb = wb;
  P->hb = hb;
}

__device__ __forceinline__ bool page_is_white(float X, float Y, const Pose& P, int cols, int rows,
                                              const uint64_t* codes, int edge, int border) {
  if (X < 0.0f || Y < 0.0f || X >= P.wb || Y >= P.hb) return true;
  const float sp = 0.3f, pitch = 1.3f;
  const float Yb = P.hb - Y;  // from the bottom of the page
  const int i = (int)floorf(X / pitch), j = (int)floorf(Yb / pitch);
  const float fx = X - i * pitch, fy = Yb - j * pitch;
  if (fx < sp && fy < sp) return false;  // corner square
  if (fx >= sp && fy >= sp && i < cols && j < rows) {
    const int cells = edge + 2 * border;
    const float tx = fx - sp, ty = 1.0f - (fy - sp);  // ty from the top of the tag
    int cc = (int)(tx * cells), cr = (int)(ty * cells);
    cc = min(max(cc, 0), cells - 1);
    cr = min(max(cr, 0), cells - 1);
    if (cc < border || cr < border || cc >= border + edge || cr >= border + edge) return false;
    const int idx = (cr - border) * edge + (cc - border);
    const uint64_t code = codes[j * cols + i];
    return (code >> (edge * edge - 1 - idx)) & 1ull;
  }
  return true;
}

// fixed_hinv: null = seeded random pose per frame; else ONE given pose (image -> page, row-major 3x3)
// for every frame, and noise only if `noise` is set (renderer tests).
__global__ void __launch_bounds__(256)
k_render_boards(uint8_t* __restrict__ frames, int w, int h, int cols, int rows,
                const uint64_t* __restrict__ codes, int edge, int border, uint64_t seed,
                const float* __restrict__ fixed_hinv, int noise) {
  __shared__ Pose s_pose;
  const int f = blockIdx.z;
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    if (fixed_hinv) {
      for (int i = 0; i < 9; ++i) s_pose.hinv[i] = fixed_hinv[i];
      s_pose.wb = cols * 1.3f + 0.3f;
      s_pose.hb = rows * 1.3f + 0.3f;
    } else {
      make_pose(seed, f, w, h, cols, rows, &s_pose);
    }
  }
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const Pose& P = s_pose;
  int white = 0;
#pragma unroll
  for (int sy = 0; sy < 4; ++sy)
#pragma unroll
    for (int sx = 0; sx < 4; ++sx) {
      const float px = (float)x + ((float)sx + 0.5f) * 0.35f - 0.7f;  // 1.4 px footprint
      const float py = (float)y + ((float)sy + 0.5f) * 0.35f - 0.7f;
      const float W = P.hinv[6] * px + P.hinv[7] * py + P.hinv[8];
      const float X = (P.hinv[0] * px + P.hinv[1] * py + P.hinv[2]) / W;
      const float Y = (P.hinv[3] * px + P.hinv[4] * py + P.hinv[5]) / W;
      white += page_is_white(X, Y, P, cols, rows, codes, edge, border) ? 1 : 0;
    }
  float v = 30.0f + 170.0f * (float)white * (1.0f / 16.0f);
  uint64_t s = seed ^ (((uint64_t)f << 40) + ((uint64_t)y << 20) + (uint64_t)x);
  const uint64_t r = splitmix64(s);
  const int nsum = (int)(r & 255) + (int)((r >> 8) & 255) + (int)((r >> 16) & 255) + (int)((r >> 24) & 255);
  if (noise) v += ((float)nsum - 510.0f) * (2.0f / 147.8f);  // ~N(0, 2^2)
  v = fminf(fmaxf(rintf(v), 0.0f), 255.0f);
  frames[((size_t)f * h + y) * w + x] = (uint8_t)v;
}

int launch_render_boards(uint8_t* frames, int n_frames, int w, int h, int cols, int rows,
                         const uint64_t* d_codes, int edge, int border, uint64_t seed,
                         const float* d_fixed_hinv, int noise, cudaStream_t s) {
  dim3 block(32, 8);
  dim3 grid((w + 31) / 32, (h + 7) / 8, n_frames);
  k_render_boards<<<grid, block, 0, s>>>(frames, w, h, cols, rows, d_codes, edge, border, seed, d_fixed_hinv, noise);
  return 1;
}

}  // namespace ag
