// K6/K7: board assembly + tag decoding, one warp per frame (see ag_board_core.h).
// reference: src/detector.rs:505-639, src/board.rs, src/saddle.rs.
#include "ag_board_core.h"
#include "ag_common.cuh"
#include "ag_kernels.h"

namespace ag {

__constant__ uint64_t c_codes[agb::kMaxCodes];

int upload_codes(const uint64_t* codes, int n) {
  if (n > agb::kMaxCodes) return -1;
  return (int)cudaMemcpyToSymbol(c_codes, codes, sizeof(uint64_t) * n);
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

BoardWsLayout make_board_layout(int max_saddles) {
  BoardWsLayout L;
  const int N = max_saddles;
  const int Q = N / 4 + 2;
  L.max_saddles = N;
  L.max_quads = Q;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t r = o;
    o = align_up(o + bytes, 16);
    return r;
  };
  for (int i = 0; i < 6; ++i) L.off_pos[i] = take(sizeof(float) * N);
  for (int i = 0; i < 2; ++i) {
    L.off_cell[i] = take(sizeof(int16_t) * agb::kCells);
    L.off_quads[i] = take(sizeof(int16_t) * 4 * Q);
    L.off_touched[i] = take(sizeof(int16_t) * agb::kCells);
    L.off_active[i] = take(N);
  }
  L.off_stack = take(sizeof(int16_t) * 2 * (Q + 1));
  L.off_seeds = take(sizeof(int16_t) * N);
  L.off_nn = take(sizeof(int16_t) * 64);
  L.off_same = take(sizeof(int16_t) * 64);
  L.off_diff = take(sizeof(int16_t) * 64);
  L.off_samp = take(sizeof(int16_t) * 64);
  L.off_hist = take(sizeof(int) * agb::kHistBins);
  L.off_remove = take(N);
  L.off_tag_valid = take(agb::kMaxCodes);
  L.off_tag_by_id = take(sizeof(agb::TagRec) * agb::kMaxCodes);
  L.bytes_per_frame = align_up(o, 256);
  return L;
}

constexpr int kBoardWarpsPerBlock = 4;

__global__ void __launch_bounds__(kBoardWarpsPerBlock * 32)
k_boards_decode(const uint8_t* __restrict__ frames, FrameGeom g, int n_frames,
                const ag_saddle* __restrict__ refined, const int* __restrict__ n_refined,
                uint8_t* __restrict__ ws, BoardWsLayout L, int n_codes, int edge, int border,
                int hamming, int max_boards, ag_tag* __restrict__ out, int cap,
                int* __restrict__ n_out, uint32_t* __restrict__ frame_status,
                int32_t* __restrict__ tap_quads, int* __restrict__ tap_n_quads, int tap_cap) {
  const int lane = threadIdx.x & 31;
  const int f = blockIdx.x * kBoardWarpsPerBlock + (threadIdx.x >> 5);
  if (f >= n_frames) return;
  uint8_t* W = ws + (size_t)f * L.bytes_per_frame;

  agb::Frame F;
  F.lane = lane;
  F.n = n_refined[f];
  F.sx = (float*)(W + L.off_pos[0]);
  F.sy = (float*)(W + L.off_pos[1]);
  F.st = (float*)(W + L.off_pos[2]);
  F.sx2 = (float*)(W + L.off_pos[3]);
  F.sy2 = (float*)(W + L.off_pos[4]);
  F.st2 = (float*)(W + L.off_pos[5]);
  for (int i = 0; i < 2; ++i) {
    F.bs[i].cell = (int16_t*)(W + L.off_cell[i]);
    F.bs[i].quads = (int16_t*)(W + L.off_quads[i]);
    F.bs[i].touched = (int16_t*)(W + L.off_touched[i]);
    F.bs[i].active = W + L.off_active[i];
    F.bs[i].n_quads = F.bs[i].n_touched = F.bs[i].score = 0;
  }
  F.stack = (int16_t*)(W + L.off_stack);
  F.seeds = (int16_t*)(W + L.off_seeds);
  F.nn_idx = (int16_t*)(W + L.off_nn);
  F.same = (int16_t*)(W + L.off_same);
  F.diff = (int16_t*)(W + L.off_diff);
  F.samp = (int16_t*)(W + L.off_samp);
  F.hist = (int*)(W + L.off_hist);
  F.remove = W + L.off_remove;
  F.max_quads = L.max_quads;
  F.img = frames + (size_t)f * g.frame_stride;
  F.w = g.w;
  F.h = g.h;
  F.format = g.format;
  F.row_stride = g.row_stride;
  F.codes = c_codes;
  F.n_codes = n_codes;
  F.edge = edge;
  F.border = border;
  F.hamming = hamming;
  F.tag_valid = W + L.off_tag_valid;
  F.tag_by_id = (agb::TagRec*)(W + L.off_tag_by_id);
  F.tap_quads = tap_quads ? tap_quads + (size_t)f * tap_cap * 4 : nullptr;
  F.tap_n_quads = tap_n_quads ? tap_n_quads + f : nullptr;
  F.tap_cap = tap_cap;
  F.status = 0;

  // workspace init: lattice cells 0, every saddle active, no tags
  {
    uint4 z = make_uint4(0, 0, 0, 0);
    for (int b = 0; b < 2; ++b) {
      uint4* c = (uint4*)F.bs[b].cell;
      for (int i = lane; i < agb::kCells * 2 / 16; i += 32) c[i] = z;
      uint32_t* a = (uint32_t*)F.bs[b].active;
      for (int i = lane; i < (L.max_saddles + 3) / 4; i += 32) a[i] = 0x01010101u;
    }
    uint32_t* tv = (uint32_t*)F.tag_valid;
    for (int i = lane; i < agb::kMaxCodes / 4; i += 32) tv[i] = 0;
    if (F.tap_n_quads && lane == 0) *F.tap_n_quads = 0;
  }
  // saddles AoS -> SoA
  const ag_saddle* S = refined + (size_t)f * L.max_saddles;
  for (int i = lane; i < F.n; i += 32) {
    ag_saddle s = S[i];
    F.sx[i] = s.x;
    F.sy[i] = s.y;
    F.st[i] = s.theta;
  }
  __syncwarp();

  agb::detect_boards(F, max_boards);

  // emit the map in ascending id order
  ag_tag* O = out + (size_t)f * cap;
  int n = 0;
  for (int base = 0; base < n_codes; base += 32) {
    const int id = base + lane;
    const bool v = id < n_codes && F.tag_valid[id];
    const unsigned m = __ballot_sync(0xffffffffu, v);
    const int dst = n + __popc(m & ((1u << lane) - 1u));
    if (v && dst < cap) {
      const agb::TagRec t = F.tag_by_id[id];
      ag_tag o;
      o.id = t.id;
#pragma unroll
      for (int j = 0; j < 8; ++j) o.xy[j] = t.xy[j];
      O[dst] = o;
    }
    n += __popc(m);
  }
  if (lane == 0) {
    n_out[f] = n;
    uint32_t st = F.status;
    if (n > cap) st |= (uint32_t)AG_FRAME_TAG_OVERFLOW;
    if (st) atomicOr(frame_status + f, st);
  }
}

int launch_boards_decode(const uint8_t* frames, const FrameGeom& g, int n_frames,
                         const ag_saddle* refined, const int* n_refined, uint8_t* ws,
                         const BoardWsLayout& L, int n_codes, int edge, int border, int hamming,
                         int max_boards, ag_tag* out, int cap, int* n_out, uint32_t* frame_status,
                         int32_t* tap_quads, int* tap_n_quads, int tap_cap, cudaStream_t s) {
  int blocks = (n_frames + kBoardWarpsPerBlock - 1) / kBoardWarpsPerBlock;
  k_boards_decode<<<blocks, kBoardWarpsPerBlock * 32, 0, s>>>(
      frames, g, n_frames, refined, n_refined, ws, L, n_codes, edge, border, hamming, max_boards,
      out, cap, n_out, frame_status, tap_quads, tap_n_quads, tap_cap);
  return 1;
}

}  // namespace ag
