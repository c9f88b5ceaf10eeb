// Throughput path of the board search (device only).  Same results as ag_board_core.h's
// find_best_board, different mapping onto the warp:
//
//   * try_find_best_board (detector.rs:588-639) visits seeds one after the other; the seed's
//     candidate quads (init_quads, detector.rs:543-586) are ENUMERATED first, in the
//     reference's order, into a list;
//   * every quad of the list is then SCORED by growing its board (Board::new + try_expand,
//     board.rs:27-48, :114-176).  Board growth is a chain of dependent steps with four small
//     neighbour searches each, so a whole warp per board wastes 28 lanes: here a GROUP OF FOUR
//     LANES grows one board (lane j runs neighbour search j of try_expand_one), eight boards
//     per warp side by side, each with a 1 KB state in shared memory (16 x 16 tag window,
//     64 quads).  Only the score (= number of quads) is kept;
//   * a board that leaves the small window or has more than 64 quads is re-scored by the
//     general warp-wide board_build of ag_board_core.h (same result, any size);
//   * the winning (seed, quad) -- first maximum in the reference's visiting order, with its
//     strict `score > best_score`, `>= 36` early exit and 30-seed limit -- is rebuilt once
//     with the general board_build, so everything downstream (try_fix_missing, decoding)
//     runs on the same state as before.
//
// The enumeration of init_quads is reorganised so that the cheap gates of is_valid_quad
// (saddle.rs:18-66) run first and only their survivors reach the transcendental gates; every
// gate only ever rejects, so the set and the order of valid quads are those of the reference.
#pragma once
#include "ag_board_core.h"

namespace agb {

constexpr int kGridCapCells = 1408;  // buckets of the neighbour grid (BoardWsLayout::grid_cap_cells)
constexpr int kGridStartBytes = ((kGridCapCells + 2) * 2 + 15) & ~15;  // its bucket-start array, 16-byte padded
#ifndef AGB_QCACHE_BITS
#define AGB_QCACHE_BITS 11
#endif
constexpr int kQCacheBits = AGB_QCACHE_BITS;  // neighbour-search cache: 2^11 entries of 8 bytes per frame (group_query)
constexpr int kQCacheEntries = 1 << kQCacheBits;
constexpr int kFastMaxSaddles = 1024;  // saddle list and bucket grid live in shared memory (two
                                       // layout tiers: 512 and 1024 saddles, see make_board_layout)
constexpr int kGroupLanes = 4;
constexpr int kGroupsPerWarp = 8;
constexpr int kGroupBytes = 1024;
constexpr int kGWin = 16;           // tag window per group: x, y in [-8, 7]
constexpr int kGQuads = 64;         // quads per group board
constexpr int kQListCap = 64;       // candidate quads a warp buffers between enumeration and scoring
constexpr int kScoreRedo = 0xffff;  // group board overflowed: score it with the general path
// group state layout (bytes)
constexpr int kGOffCell = 0;        // u8 [256]: 0 unvisited, 0xff None, q + 1 Some(q)
constexpr int kGOffQuads = 256;     // i16 [64][4]
constexpr int kGOffStack = 768;     // u16 [64]: cell | next direction << 8
constexpr int kGOffActive = 896;    // u32 [32]: up to 1024 saddles
constexpr int kDiffCap = 52;        // `diff` / `same` entries of a seed: at most 49 (50-NN minus the seed)

}  // namespace agb

#if AGB_DEVICE
namespace agb {

// ---- bucket grid, built by the whole warp (grid_build_parallel, ag_board_core.h) -------------------
__device__ __noinline__ void grid_build_warp(Frame& F) { grid_build_parallel(F, true); }

// ---- one neighbour search of try_expand_one per lane, the whole warp in step -----------------------
// find_closest_potential_saddle_idxs (board.rs:177-234) for the edge a -> b seen from `self`:
// the up to three nearest saddles within the radius, ascending by (d2, index), then filtered by
// the board's active mask and the theta gate.  Returns the survivors packed as
// count | c0 << 2 | c1 << 12 | c2 << 22 (indices < 1024).
//
// EVERY lane of the warp calls this (lanes with `on` == false search nothing): the candidate
// loop is kept convergent with one vote per iteration, so the warp pays the longest search of
// its 32 lanes once instead of running the lanes' loops one after the other.
//
// Best-first scan of the window's bucket rows: the row of the query first, then alternately
// the rows above and below (a row of buckets is one contiguous range of the grid-ordered
// arrays).  Once three saddles are known, a row whose vertical distance from the query exceeds
// the third-best distance ends the search, and the column range of a row shrinks to that
// distance -- a long edge a -> b gives a radius that covers most of the board although only
// the three nearest saddles matter.  All bounds are conservative, so the result equals the
// scan of the whole window.
// The frame's fields the neighbour search needs, read once per warp_score_quads call and kept in
// registers (the Frame itself lives in local memory; the search runs hundreds of times per call).
// A pointer into the block's dynamic shared memory, re-derived from the shared array itself: the
// compiler then knows the address space and emits LDS / STS with 32-bit addresses instead of
// generic loads and stores through 64-bit pointers (the Frame holds generic pointers because the
// general path may point them at global memory).  Only for pointers known to be in shared memory.
extern __shared__ __align__(16) uint8_t agb_dyn_smem[];
template <class T>
__device__ __forceinline__ T* as_shared(T* p) {
  const unsigned off = (unsigned)__cvta_generic_to_shared(p) - (unsigned)__cvta_generic_to_shared(agb_dyn_smem);
  return (T*)(agb_dyn_smem + off);
}

// The saddle list of a throughput-path frame lives in shared memory (F.fast_on requires it): x, y
// and theta arrays `stride` bytes apart, read with 32-bit shared addresses.
struct Pts {
  unsigned base, stride;
  __device__ __forceinline__ float ld(unsigned addr) const {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
  }
  __device__ __forceinline__ float x(int i) const { return ld(base + 4u * (unsigned)i); }
  __device__ __forceinline__ float y(int i) const { return ld(base + stride + 4u * (unsigned)i); }
  __device__ __forceinline__ float t(int i) const { return ld(base + 2u * stride + 4u * (unsigned)i); }
};
// is_valid_quad (saddle.rs:17-67) on the shared-memory saddle list: no call, no generic loads
__device__ __forceinline__ bool is_valid_quad_s(const Pts& P, int s0, int d0, int s1, int d1) {
  const float x0 = P.x(s0), y0 = P.y(s0), x1 = P.x(s1), y1 = P.y(s1);
  if (!quad_diag_ok_v(x0, y0, P.t(s0), x1, y1)) return false;
  return quad_rest_ok_v(x0, y0, P.x(d0), P.y(d0), P.t(d0), x1, y1, P.x(d1), P.y(d1), P.t(d1));
}
struct QueryCtx {
  Pts P;
  float g_inv;
  int g_nx, g_ny, g_on, n;
  // Shared-memory address of the bucket starts (F.g_start = array base + one entry).  The
  // grid-ordered positions follow that array and the grid-ordered item indices follow the positions
  // (BoardWsLayout; checked where F.fast_on is set), so their addresses are a_start plus a constant
  // / plus two saddle-array strides: nothing else has to stay live across the search loops.
  unsigned a_start;
  __device__ __forceinline__ unsigned a_pos() const { return a_start + (unsigned)(kGridStartBytes - 2); }
  __device__ __forceinline__ unsigned a_item() const { return a_pos() + 2u * P.stride; }
};
__device__ __forceinline__ QueryCtx make_query_ctx(const Frame& F) {
  QueryCtx C;
  C.P.base = (unsigned)__cvta_generic_to_shared(F.sx);
  C.P.stride = (unsigned)((const char*)F.sy - (const char*)F.sx);  // F.st = F.sy + the same stride
  C.g_inv = F.g_inv; C.g_nx = F.g_nx; C.g_ny = F.g_ny; C.g_on = F.g_on; C.n = F.n;
  C.a_start = F.g_on ? (unsigned)__cvta_generic_to_shared(F.g_start) : 0u;
  return C;
}
// group_knn: the search proper (the up to three nearest saddles within the radius, unfiltered,
// packed as above).  It depends on (a, b, self == b) and the frame's saddle list only -- not on the
// board being grown -- so its result is shared by every board of the frame through a cache
// (group_query below).
__device__ __forceinline__ unsigned group_knn(const QueryCtx& F, bool on, int a, int b, int self, uint32_t* tm = nullptr) {
  const unsigned full = 0xffffffffu;
  // key = squared distance bits << 32 | index; "none" carries distance +inf, so the distance of the
  // third-best candidate is the high word of k2 whether or not three are known
  const unsigned long long kInf = (0x7f800000ull << 32) | 0xffffffffull;
  unsigned long long k0 = kInf, k1 = kInf, k2 = kInf;
  float d3f = __uint_as_float(0x7f800000u);  // distance of the current third-best candidate (cheap first reject)
  auto insert = [&](float d, int i) {
    if (d > d3f) return;
    const unsigned long long k = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)i;
    if (k < k2) {
      k2 = k;
      if (k2 < k1) { const unsigned long long t = k1; k1 = k2; k2 = t; }
      if (k1 < k0) { const unsigned long long t = k0; k0 = k1; k1 = t; }
      d3f = __uint_as_float((unsigned)(k2 >> 32));
    }
  };
  float qx = 0.0f, qy = 0.0f, r2 = -1.0f, r = 0.0f;
  if (on) {
    const float ratio0 = fadd(1.0f, 0.3f);  // 1.0 + spacing_ratio; detector.rs:621 passes 0.3
    const float ax = F.P.x(a), ay = F.P.y(a), bx = F.P.x(b), by = F.P.y(b);
    const float dx = fsub(ax, bx), dy = fsub(ay, by);
    r2 = fmul(0.5f, fadd(fmul(dx, dx), fmul(dy, dy)));
    const float v10x = fsub(bx, ax), v10y = fsub(by, ay);
    const float sfx = self == a ? ax : bx, sfy = self == a ? ay : by;
    qx = fadd(sfx, fmul(v10x, ratio0));
    qy = fadd(sfy, fmul(v10y, ratio0));
  }
  if (F.g_on) {  // block-uniform
    const bool windowed = on && r2 >= 0.0f && r2 < 1.0e12f;
    int x0 = 0, x1 = -1, y0 = 0, y1 = -1, cy = 0, t = 0, t_end = 0;
    if (windowed) {
      r = sqrtf(r2) * 1.0001f + 0.01f;
      x0 = (int)floorf((qx - r) * F.g_inv); x1 = (int)floorf((qx + r) * F.g_inv);
      y0 = (int)floorf((qy - r) * F.g_inv); y1 = (int)floorf((qy + r) * F.g_inv);
      // both ends clamped INTO the grid (saddles outside the image sit in the border buckets)
      x0 = x0 < 0 ? 0 : (x0 >= F.g_nx ? F.g_nx - 1 : x0); y0 = y0 < 0 ? 0 : (y0 >= F.g_ny ? F.g_ny - 1 : y0);
      x1 = x1 >= F.g_nx ? F.g_nx - 1 : (x1 < 0 ? 0 : x1); y1 = y1 >= F.g_ny ? F.g_ny - 1 : (y1 < 0 ? 0 : y1);
      if (x0 <= x1 && y0 <= y1) {
        cy = (int)floorf(qy * F.g_inv);
        cy = cy < y0 ? y0 : (cy > y1 ? y1 : cy);
        t_end = 2 * max(cy - y0, y1 - cy) + 1;  // rows are visited at offsets 0, -1, +1, -2, ...
      }
    }
    const float bsz = 1.0f / F.g_inv;  // bucket side, a power of two: bucket indices are exact
    // shared-memory addresses of the grid arrays (32-bit, LDS instead of generic loads)
    const unsigned a_start = F.a_start, a_pos = F.a_pos(), a_item = F.a_item();
    auto lds_u16 = [](unsigned addr) -> int {
      unsigned short v;
      asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
      return (int)v;
    };
    // Outer loop: row offset t is the same for every lane (each lane's rows are counted from its own
    // query row), so the row set-up is executed once per t for the whole warp; it is issued one
    // row AHEAD (its shared-memory loads overlap the candidate loop of the current row; the bounds
    // it uses are those known before the current row, i.e. only looser).  Inner loop: two
    // candidates per lane and iteration.
    // Column range and row cut-off follow the third-best distance; they are recomputed only when
    // the third-best candidate changes (sqrt / floor are the expensive part of the row set-up).
    float d3_seen = d3f;
    int bx0 = x0, bx1 = x1;
    float stop_d3 = 3.0e38f;  // a row whose lower bound lb has lb^2 * 0.9999 > stop_d3 ends the search
    auto row_setup = [&](int tt, int& re, int& re1) {
      re = re1 = 0;
      if (tt >= t_end) return;
      const int j = (tt + 1) >> 1, yy = (tt & 1) ? cy - j : cy + j;
      if (yy < y0 || yy > y1) return;
      if (d3f != d3_seen) {
        d3_seen = d3f;
        const float d3 = d3f;
        stop_d3 = d3;
        const float R = fminf(r, sqrtf(d3) * 1.0001f + 0.01f);
        bx0 = (int)floorf((qx - R) * F.g_inv);
        bx1 = (int)floorf((qx + R) * F.g_inv);
        bx0 = bx0 < x0 ? x0 : (bx0 > x1 ? x1 : bx0);
        bx1 = bx1 > x1 ? x1 : (bx1 < x0 ? x0 : bx1);
      }
      const float lb = (float)(j > 0 ? j - 1 : 0) * bsz;  // every saddle of the row is farther than this
      if (lb * lb * 0.9999f > stop_d3) {                   // ... and so are the remaining rows
        t_end = 0;
      } else if (bx0 <= bx1) {
        const int b0 = yy * F.g_nx;
        re = lds_u16(a_start + 2u * (unsigned)(b0 + bx0));
        re1 = lds_u16(a_start + 2u * (unsigned)(b0 + bx1 + 1));
      }
    };
    int e = 0, e1 = 0;
    row_setup(0, e, e1);
    t = 1;
    while (__any_sync(full, e < e1 || t < t_end)) {
      int ne, ne1;
#ifdef AGB_LOOP_STATS
      if (tm && (threadIdx.x & 31) == 0) tm[23] += 1;
#endif
      row_setup(t, ne, ne1);
      ++t;
      while (__any_sync(full, e < e1)) {
#ifdef AGB_LOOP_STATS
        if (tm) { const unsigned mm = __ballot_sync(full, e < e1); if ((threadIdx.x & 31) == 0) { tm[24] += 1; tm[25] += __popc(mm); } }
#endif
        if (e < e1) {
          const bool two = e + 1 < e1;
          const unsigned ea = a_pos + 8u * (unsigned)e, eb = two ? ea + 8u : ea;
          float pax, pay, pbx, pby;
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(pax), "=f"(pay) : "r"(ea));
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(pbx), "=f"(pby) : "r"(eb));
          const float dax = fsub(qx, pax), day = fsub(qy, pay);  // dist2(): (0 + dx*dx) + dy*dy
          const float dbx = fsub(qx, pbx), dby = fsub(qy, pby);
          const float da = fadd(fmul(dax, dax), fmul(day, day));
          const float db = fadd(fmul(dbx, dbx), fmul(dby, dby));
          if (da <= r2) insert(da, lds_u16(a_item + 2u * (unsigned)e));
          if (two && db <= r2) insert(db, lds_u16(a_item + 2u * (unsigned)(e + 1)));
          e += 2;
        }
      }
      e = ne;
      e1 = ne1;
    }
    if (on && !windowed)  // degenerate radius (NaN / huge): exhaustive
      for (int i = 0; i < F.n; ++i) {
        const float ddx = fsub(qx, F.P.x(i)), ddy = fsub(qy, F.P.y(i));
        const float d = fadd(fmul(ddx, ddx), fmul(ddy, ddy));
        if (d <= r2) insert(d, i);
      }
  } else {
    for (int i = 0; __any_sync(full, on && i < F.n); ++i)
      if (on && i < F.n) {
        const float ddx = fsub(qx, F.P.x(i)), ddy = fsub(qy, F.P.y(i));
        const float d = fadd(fmul(ddx, ddx), fmul(ddy, ddy));
        if (d <= r2) insert(d, i);
      }
  }
  unsigned packed = 0, cnt = 0;
  if (k0 != kInf) { packed |= ((unsigned)k0 & 0x3ffu) << 2; ++cnt; }
  if (k1 != kInf) { packed |= ((unsigned)k1 & 0x3ffu) << 12; ++cnt; }
  if (k2 != kInf) { packed |= ((unsigned)k2 & 0x3ffu) << 22; ++cnt; }
  return packed | cnt;
}
// The frame's neighbour-search cache: direct-mapped, in global memory (L2), one 64-bit entry
// = tag (valid bit, search round, a, b, self == b) << 32 | packed result.  Boards grown from
// different quads of a frame walk the same tags, so most searches after the first board of a
// warp are repeats.  An entry is written and read with single 64-bit accesses; a reader sees
// the old or the new entry, both are complete, and a hit returns exactly what the search
// would, so races between the warps of a frame do not change the result.
__device__ __forceinline__ unsigned group_query(const QueryCtx& F, unsigned long long* qcache, unsigned round_tag,
                                                const uint32_t* active, bool on, int a, int b, int self,
                                                uint32_t* tm) {
  unsigned knn = 0, tag = 0;
  unsigned long long* slot = nullptr;
  bool search = on;
  if (on && qcache) {
    const unsigned key = ((unsigned)a << 11) | ((unsigned)b << 1) | (self == a ? 0u : 1u);
    tag = round_tag | key;
    slot = qcache + ((key * 0x9E3779B1u) >> (32 - kQCacheBits));
    const unsigned long long e = __ldcg(slot);
    if ((unsigned)(e >> 32) == tag) {
      knn = (unsigned)e;
      search = false;
    }
  }
  if (tm) {
    const unsigned ms = __ballot_sync(0xffffffffu, search), mo = __ballot_sync(0xffffffffu, on);
    if ((threadIdx.x & 31) == 0) { tm[20] += __popc(mo); tm[21] += __popc(ms); tm[22] += ms != 0u; }
  }
  if (__any_sync(0xffffffffu, search)) {
    const unsigned r = group_knn(F, search, a, b, self, tm);
    if (search) {
      knn = r;
      if (slot) __stcg(slot, ((unsigned long long)tag << 32) | r);
    }
  }
  // active mask of the board and theta gate (board.rs:218-231), order kept
  unsigned packed = 0, cnt = 0;
  if (on) {
    const float ts = F.P.t(self);
    const int n = (int)(knn & 3u);
    for (int t = 0; t < 3; ++t) {
      if (t >= n) break;
      const int i = (int)((knn >> (2 + 10 * t)) & 0x3ffu);
      if (((active[i >> 5] >> (i & 31)) & 1u) && theta_distance_degree(ts, F.P.t(i)) < 5.0f) {
        packed |= (unsigned)i << (2 + 10 * cnt);
        ++cnt;
      }
    }
  }
  return packed | cnt;
}
__device__ __forceinline__ int packed_count(unsigned p) { return (int)(p & 3u); }
__device__ __forceinline__ int packed_cand(unsigned p, int k) { return (int)((p >> (2 + 10 * k)) & 0x3ffu); }

// ---- Board::new + try_expand (board.rs:27-48, :114-176), eight boards per warp --------------------
// Scores the quads qlist[0 .. nq) into qscore[]: score = number of quads of the grown board, or kScoreRedo
// when the board leaves the group window / quad capacity.  A group of four lanes grows one board;
// the eight groups of the warp advance in LOCKSTEP, one expansion attempt per iteration (cheap
// bookkeeping steps -- returning to a parent cell, neighbours that already hold a tag -- are
// skipped inside the iteration), so the warp issues each instruction of the expensive part once
// for all eight boards.  A group that finishes its board takes the next quad of the list
// of the warp's list, so long and short boards overlap.
constexpr int kCtlNext = 6;  // slot of F.ctl: next wave slot (seed) to hand to a warp
// A board of at least kSaveMin quads (a real board, not the few-quad debris of tag interiors) is
// worth keeping: each warp keeps the state of its best such board (first one of the highest score)
// in global memory, and when the search's winner is one of them it is converted instead of being
// grown a second time by the general path.  Layout: int score, i16 quad[4], pad to 16 bytes, then
// the group state's cell window (256 B) and quads (512 B).
constexpr int kSaveMin = 12;
constexpr int kSaveBytes = 16 + 256 + 512;
// TM: the instantiation with the timing taps (option board_timing); the production one carries none.
template <bool TM>
__device__ __noinline__ void warp_score_quads_t(const Frame& F, const int16_t* qlist, uint16_t* qscore, int nq) {
  const QueryCtx QC = make_query_ctx(F);
  const int max_quads = F.max_quads;
  // searches are cached per (round, a, b, self); the 7-bit round tag bounds the rounds that may use it
  unsigned long long* const qcache = F.round < 127 ? F.fx_qcache : nullptr;
  const unsigned round_tag = 0x80000000u | ((unsigned)F.round << 24);
  uint8_t* const save = F.fx_save0 + (size_t)F.warp * F.fx_save_stride;
  int save_score = *(const int*)save;  // warp-uniform
  const unsigned full = 0xffffffffu;
  const int grp = F.lane >> 2, jl = F.lane & 3, gshift = grp * 4;
  const unsigned gmask = 0xfu << gshift;
  uint8_t* const gstate = as_shared(F.fx_gstate);
  uint8_t* gs = gstate + grp * kGroupBytes;
  uint8_t* cell = gs + kGOffCell;
  int16_t* quads = (int16_t*)(gs + kGOffQuads);
  uint16_t* stack = (uint16_t*)(gs + kGOffStack);
  uint32_t* active = (uint32_t*)(gs + kGOffActive);
  const int c_origin = 8 * kGWin + 8;
  bool alive = false, list_empty = false;
  int k = 0, n_quads = 0, depth = 0, cur_ci = 0, cur_i = 0;
  int next = 0;  // list cursor (warp-uniform)
  const bool tmon = TM && F.tm && F.warp == 0 && F.lane == 0;
  for (;;) {
    long long tc0 = TM ? clock64() : 0ll;
    // (0) idle groups take the next quads of the list
    const unsigned idle = __ballot_sync(full, !alive) & 0x11111111u;
    if (idle != 0u && !list_empty) {
      const int cnt = __popc(idle);
      const int base = next;
      next += cnt;
      if (base + cnt >= nq) list_empty = true;
      const int kk = base + __popc(idle & ((1u << gshift) - 1u));
      if (!alive && kk < nq) {
        k = kk;
        alive = true;
        uint32_t* c32 = (uint32_t*)cell;  // window empty, every saddle active
#pragma unroll
        for (int t = 0; t < 16; ++t) c32[jl + 4 * t] = 0u;
#pragma unroll
        for (int t = 0; t < 8; ++t) active[jl + 4 * t] = 0xffffffffu;
        __syncwarp(gmask);
        if (jl == 0) {
          const int16_t* quad = qlist + 4 * k;
          for (int j = 1; j < 4; ++j) {  // quad[0] stays active (board.rs:35-37)
            const int sdl = quad[j];
            active[sdl >> 5] &= ~(1u << (sdl & 31));
          }
          for (int j = 0; j < 4; ++j) quads[j] = quad[j];
          cell[c_origin] = 1;
        }
        n_quads = 1;
        depth = 1;
        cur_ci = c_origin;
        cur_i = 0;
      }
      __syncwarp();
    }
    if (!__any_sync(full, alive)) break;
    if (tmon) { const long long tc = clock64(); F.tm[16] += (uint32_t)(tc - tc0); tc0 = tc; }
    // (1) every running board advances to its next expansion attempt (or finishes)
    bool need = false;
    int dir = 0, nci = 0, result = -1;
    if (alive) {
      for (;;) {
        if (cur_i == 4) {  // all four directions of this cell done: return to the parent
          if (--depth == 0) { result = n_quads; break; }
          const unsigned e = stack[depth - 1];
          cur_ci = (int)(e & 0xffu);
          cur_i = (int)(e >> 8);
          continue;
        }
        dir = cur_i++;
        const int bx = (cur_ci >> 4) - 8, by = (cur_ci & 15) - 8;
        int nx = bx, ny = by;
        if (dir == 0) nx = bx + 1;
        else if (dir == 1) ny = by - 1;
        else if (dir == 2) nx = bx - 1;
        else ny = by + 1;
        if (nx < -8 || nx > 7 || ny < -8 || ny > 7) { result = kScoreRedo; break; }
        nci = (nx + 8) * kGWin + (ny + 8);
        const int cur = cell[nci];
        if (cur != 0 && cur != 0xff) continue;  // already Some (board.rs:131-135)
        if (n_quads >= max_quads) {             // no room: the attempt fails, the cell becomes None
          if (jl == 0) cell[nci] = 0xff;
          __syncwarp(gmask);
          continue;
        }
        if (n_quads >= kGQuads) { result = kScoreRedo; break; }
        need = true;
        break;
      }
      if (result >= 0) {
        if (jl == 0) qscore[k] = (uint16_t)result;
        alive = false;
      }
    }
    __syncwarp();
    {
      // keep the state of the warp's best large board: among the boards that finished in this
      // iteration the highest score wins, ties go to the earlier list entry; an earlier
      // iteration's board is replaced only by a strictly higher score
      const bool big = result >= kSaveMin && result != kScoreRedo;
      const unsigned key = (big && jl == 0) ? (((unsigned)result << 16) | (unsigned)(0xffff - k)) : 0u;
      const unsigned best = __reduce_max_sync(full, key);
      if ((int)(best >> 16) > save_score) {  // warp-uniform
        const int src_lane = __ffs((int)__ballot_sync(full, key == best)) - 1;  // leader of the winning group
        const uint8_t* src = gstate + (src_lane >> 2) * kGroupBytes;
        save_score = (int)(best >> 16);
        const int kk = 0xffff - (int)(best & 0xffffu);
        if (F.lane == 0) *(int*)save = save_score;
        if (F.lane < 4) ((int16_t*)(save + 4))[F.lane] = qlist[4 * kk + F.lane];
        const uint32_t* s32 = (const uint32_t*)src;  // cell window [0, 256) and quads [256, 768)
        uint32_t* d32 = (uint32_t*)(save + 16);
#pragma unroll
        for (int t = 0; t < 6; ++t) d32[F.lane + 32 * t] = s32[F.lane + 32 * t];
        __syncwarp();
      }
    }
    {
      const unsigned nb = __ballot_sync(full, need);
      if (nb == 0u) continue;
      if (tmon) {
        F.tm[13] += 1;
        F.tm[14] += __popc(nb) >> 2;
      }
    }
    // (2) the four neighbour searches of try_expand_one, one per lane
    if (tmon) { const long long tc = clock64(); F.tm[17] += (uint32_t)(tc - tc0); tc0 = tc; }
    int qa = 0, qb = 0, qself = 0;
    if (need) {
      const int qi = cell[cur_ci] - 1;
      const int2 qv = *(const int2*)(quads + 4 * qi);  // 4 x i16
      const int qq0 = (int)(int16_t)(qv.x & 0xffff), qq1 = qv.x >> 16;
      const int qq2 = (int)(int16_t)(qv.y & 0xffff), qq3 = qv.y >> 16;
      // rotate_left(dir): qs[j] = quad[(j + dir) & 3]; lanes 0, 1 use the edge qs[0] -> qs[1],
      // lanes 2, 3 the edge qs[3] -> qs[2]
      const int ia = (jl < 2 ? dir : dir + 3) & 3, ib = (jl < 2 ? dir + 1 : dir + 2) & 3;
      qa = ia == 0 ? qq0 : (ia == 1 ? qq1 : (ia == 2 ? qq2 : qq3));
      qb = ib == 0 ? qq0 : (ib == 1 ? qq1 : (ib == 2 ? qq2 : qq3));
      qself = (jl == 1 || jl == 2) ? qb : qa;
    }
    __syncwarp();
    const unsigned mine = group_query(QC, qcache, round_tag, active, need, qa, qb, qself, (TM && F.warp == 0) ? F.tm : nullptr);  // whole warp, convergent
    if (tmon) { const long long tc = clock64(); F.tm[18] += (uint32_t)(tc - tc0); tc0 = tc; }
    const unsigned p0 = __shfl_sync(full, mine, 0, 4), p1 = __shfl_sync(full, mine, 1, 4);
    const unsigned p2 = __shfl_sync(full, mine, 2, 4), p3 = __shfl_sync(full, mine, 3, 4);
    const int n0 = packed_count(p0), n1 = packed_count(p1), n2 = packed_count(p2), n3 = packed_count(p3);
    const int total = need ? n0 * n1 * n2 * n3 : 0;
    // (3) candidate 4-tuples in the reference's nested-loop order (i0 outermost), four at a time.
    // A tuple number t < 81 is split into its digits with divisors 1..3: t / n = (t * M[n]) >> 8
    // for M = 256, 128, 86 (exact for t <= 80), instead of three integer divisions.
    const int m1 = n1 == 3 ? 86 : (256 >> (n1 >> 1)), m2 = n2 == 3 ? 86 : (256 >> (n2 >> 1)),
              m3 = n3 == 3 ? 86 : (256 >> (n3 >> 1));
    auto split3 = [](int& r, int& digit, int n, int mul) {
      const int q = (r * mul) >> 8;
      digit = r - q * n;
      r = q;
    };
    bool ok = false;
    int nq0 = 0, nq1 = 0, nq2 = 0, nq3 = 0;
    for (int base = 0; __any_sync(full, !ok && base < total); base += 4) {
      if (tmon) F.tm[15] += 1;
      const int t = base + jl;
      bool valid = false;
      if (!ok && t < total) {
        int r = t, i1, i2, i3;
        split3(r, i3, n3, m3); split3(r, i2, n2, m2); split3(r, i1, n1, m1);
        valid = is_valid_quad_s(QC.P, packed_cand(p0, r), packed_cand(p1, i1), packed_cand(p2, i2),
                                packed_cand(p3, i3));
      }
      const unsigned m = (__ballot_sync(full, valid) >> gshift) & 0xfu;
      if (!ok && m) {
        int r = base + __ffs((int)m) - 1, i1, i2, i3;
        split3(r, i3, n3, m3); split3(r, i2, n2, m2); split3(r, i1, n1, m1);
        nq0 = packed_cand(p0, r); nq1 = packed_cand(p1, i1);
        nq2 = packed_cand(p2, i2); nq3 = packed_cand(p3, i3);
        ok = true;
      }
    }
    if (tmon) { const long long tc = clock64(); F.tm[19] += (uint32_t)(tc - tc0); tc0 = tc; }
    // (4) update the board
    if (need) {
      if (ok) {
        if (jl == 0) {
          // rotate_right(dir): v[(j + dir) & 3] = nq[j]
          int v[4];
          v[dir & 3] = nq0; v[(dir + 1) & 3] = nq1; v[(dir + 2) & 3] = nq2; v[(dir + 3) & 3] = nq3;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            active[v[j] >> 5] &= ~(1u << (v[j] & 31));
            quads[n_quads * 4 + j] = (int16_t)v[j];
          }
          cell[nci] = (uint8_t)(n_quads + 1);
          stack[depth - 1] = (uint16_t)(cur_ci | (cur_i << 8));
        }
        ++n_quads;
        ++depth;
        cur_ci = nci;
        cur_i = 0;
      } else if (jl == 0) {
        cell[nci] = 0xff;
      }
    }
    __syncwarp();
  }
}

__device__ __forceinline__ void warp_score_quads(const Frame& F, const int16_t* qlist, uint16_t* qscore, int nq) {
  if (F.tm) warp_score_quads_t<true>(F, qlist, qscore, nq);
  else warp_score_quads_t<false>(F, qlist, qscore, nq);
}

// ---- init_quads, enumerated by one warp ----------------------------------------------------------
struct SeedEnum {
  int s0, n_same, n_diff;
  unsigned diag_ok[2];  // bit a: quad_diag_ok(s0, same[a])
  int a;                // `same` entry being expanded (-1 before the first)
  bool a_open;          // cand[] holds pairs of entry a that are not queued yet
  // PER LANE: for diff entry i = lane (cand[0]) / lane + 32 (cand[1]), the partners j > i that
  // passed the cheap gates and still have to be queued
  unsigned long long cand[2];
  int q_head, q_n;      // ring of cheap-gate survivors
  bool exhausted;
};

// The k (<= 64) nearest saddles of a point for n <= 512 saddles, ascending by (d2, index), into
// F.nn_idx -- same result as nearest_k (kdtree `nearest`, ties -> lower index).  Every lane keeps
// the squared distances of its <= 16 saddles in registers; the k-th smallest distance is found
// by bisection on its bit pattern (d2 >= 0: the IEEE bits are monotone) with one warp reduction
// per bit, the saddles at or below it are gathered and sorted by a 64-wide bitonic network.
__device__ __noinline__ int nearest_k_fast(Frame& F, float qx, float qy, int k, unsigned long long* scratch64) {
  // the throughput path keeps these arrays in shared memory: address them as such (as_shared)
  int16_t* const nn_idx = as_shared(F.nn_idx);
  const float* const sx = as_shared(F.sx);
  const float* const sy = as_shared(F.sy);
  scratch64 = as_shared(scratch64);
  const int n = F.n;
  const int kk = k < n ? k : n;
  if (kk <= 0) return 0;
  if (n > 512) return -1;  // 16 distances per lane; larger frames use nearest_k
  unsigned db[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int i = F.lane + 32 * j;
    db[j] = i < n ? __float_as_uint(fadd(fmul(fsub(qx, sx[i]), fsub(qx, sx[i])), fmul(fsub(qy, sy[i]), fsub(qy, sy[i])))) : 0xffffffffu;
  }
  // smallest T with |{d <= T}| >= kk  (NaN / negative patterns cannot occur: d2 = x*x + y*y)
  unsigned T = 0;
#pragma unroll 1
  for (int bit = 30; bit >= 0; --bit) {
    const unsigned trial = T | (1u << bit);
    int c = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) c += db[j] < trial ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c < kk) T = trial;  // fewer than kk strictly below trial: the answer is >= trial
  }
  // gather {d <= T} (>= kk of them; more only on exact ties at T)
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const bool in = db[j] <= T && (F.lane + 32 * j) < n;
    const unsigned m = __ballot_sync(0xffffffffu, in);
    if (in) {
      const int dst = cnt + __popc(m & ((1u << F.lane) - 1u));
      if (dst < 64) scratch64[dst] = ((unsigned long long)db[j] << 32) | (unsigned)(F.lane + 32 * j);
    }
    cnt += __popc(m);
  }
  if (cnt > 64) return -1;  // > 14 exact ties at the k-th distance: let the caller use nearest_k
  __syncwarp();
  unsigned long long v0 = F.lane < cnt ? scratch64[F.lane] : ~0ull;
  unsigned long long v1 = F.lane + 32 < cnt ? scratch64[F.lane + 32] : ~0ull;
  // bitonic sort of 64 keys, element e = lane (v0) or lane + 32 (v1)
#pragma unroll
  for (int size = 2; size <= 64; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride == 32) {
        // partner of element e is e ^ 32: the lane's own other register; ascending for size 64
        const unsigned long long lo = v0 < v1 ? v0 : v1, hi = v0 < v1 ? v1 : v0;
        v0 = lo;
        v1 = hi;
      } else {
        const unsigned long long o0 = __shfl_xor_sync(0xffffffffu, v0, stride);
        const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, v1, stride);
        const bool lower = (F.lane & stride) == 0;
        const bool up0 = (F.lane & size) == 0 || size == 64;          // direction of element `lane`
        const bool up1 = ((F.lane + 32) & size) == 0 || size == 64;   // ... of element `lane + 32`
        const bool keep_min0 = lower == up0, keep_min1 = lower == up1;
        v0 = keep_min0 ? (v0 < o0 ? v0 : o0) : (v0 < o0 ? o0 : v0);
        v1 = keep_min1 ? (v1 < o1 ? v1 : o1) : (v1 < o1 ? o1 : v1);
      }
    }
  }
  if (F.lane < kk) nn_idx[F.lane] = (int16_t)(unsigned)v0;
  if (F.lane + 32 < kk) nn_idx[F.lane + 32] = (int16_t)(unsigned)v1;
  __syncwarp();
  return kk;
}

// 50-NN of the seed, same / diff classification (detector.rs:550-563), per-diff vectors.
__device__ __noinline__ void seed_enum_begin(Frame& F, SeedEnum& E, int s0) {
  // the throughput path keeps these arrays in shared memory: address them as such (as_shared)
  const float* const sx = as_shared(F.sx);
  const float* const sy = as_shared(F.sy);
  const float* const st = as_shared(F.st);
  int16_t* const nn_idx = as_shared(F.nn_idx);
  int16_t* const same = as_shared(F.same);
  int16_t* const diff = as_shared(F.diff);
  float* const dvx = as_shared(F.fx_dvx);
  float* const dvy = as_shared(F.fx_dvy);
  unsigned long long* const tmask = as_shared(F.fx_tmask);
  E.s0 = s0;
  // the 1 KB at fx_dvx (dvx, dvy, tmask) is free until the classification below fills it
  int n_nn = nearest_k_fast(F, sx[s0], sy[s0], 50, (unsigned long long*)F.fx_dvx);
  if (n_nn < 0) n_nn = nearest_k(F, sx[s0], sy[s0], 50);
  const float t0 = st[s0], x0 = sx[s0], y0 = sy[s0];
  int n_same = 0, n_diff = 0;
  for (int base = 1; base < n_nn; base += 32) {  // nearest[1..]: the first hit is the seed itself
    const int j = base + F.lane;
    int si = 0;
    bool is_same = false, is_diff = false;
    if (j < n_nn) {
      si = nn_idx[j];
      const float td = theta_distance_degree(t0, st[si]);
      is_same = td < 5.0f;
      is_diff = !is_same && td > 80.0f;
    }
    const unsigned ms = __ballot_sync(0xffffffffu, is_same), md = __ballot_sync(0xffffffffu, is_diff);
    const unsigned lt = (1u << F.lane) - 1u;
    if (is_same) same[n_same + __popc(ms & lt)] = (int16_t)si;
    if (is_diff) {
      const int d = n_diff + __popc(md & lt);
      diff[d] = (int16_t)si;
      dvx[d] = fsub(sx[si], x0);
      dvy[d] = fsub(sy[si], y0);
    }
    n_same += __popc(ms);
    n_diff += __popc(md);
  }
  __syncwarp();
  E.n_same = n_same;
  E.n_diff = n_diff;
  E.diag_ok[0] = E.diag_ok[1] = 0u;
  for (int blk = 0; blk * 32 < n_same; ++blk) {  // n_same <= 49
    const int a = blk * 32 + F.lane;
    bool ok = false;
    if (a < n_same) {
      const int s1 = same[a];
      ok = quad_diag_ok_v(x0, y0, t0, sx[s1], sy[s1]);
    }
    E.diag_ok[blk & 1] = __ballot_sync(0xffffffffu, ok);
  }
  // theta gate of is_valid_quad (saddle.rs:21-24) for every pair (i < j) of diff entries: it does
  // not depend on s1, so it is evaluated once per seed.  tmask[i] bit j = pair (i, j) passes.
  for (int base = 0; base < n_diff; base += 32) {
    const int i = base + F.lane;
    if (i < n_diff) {
      const float ti = st[diff[i]];
      unsigned long long m = 0ull;
      for (int j = i + 1; j < n_diff; ++j)
        if (!(theta_distance_degree(ti, st[diff[j]]) > 5.0f)) m |= 1ull << j;
      tmask[i] = m;
    }
  }
  __syncwarp();
  E.a = -1;
  E.a_open = false;
  E.cand[0] = E.cand[1] = 0ull;
  E.q_head = E.q_n = 0;
  E.exhausted = n_diff < 2;
}

// Appends valid quads to qlist (capacity qcap quads, from *list_n on) until the seed is exhausted
// (returns true) or the list may not hold another batch (returns false; call again after draining
// the list).
__device__ __noinline__ bool seed_enum_fill(Frame& F, SeedEnum& E, int16_t* qlist, int qcap, int* list_n) {
  // the throughput path keeps these arrays in shared memory: address them as such (as_shared)
  const float* const sx = as_shared(F.sx);
  const float* const sy = as_shared(F.sy);
  uint32_t* const squeue = as_shared(F.fx_squeue);
  int16_t* const same = as_shared(F.same);
  int16_t* const diff = as_shared(F.diff);
  const float* const st = as_shared(F.st);
  float* const dvx = as_shared(F.fx_dvx);
  float* const dvy = as_shared(F.fx_dvy);
  unsigned long long* const tmask = as_shared(F.fx_tmask);
  const unsigned lt = (1u << F.lane) - 1u;
  const float x0 = sx[E.s0], y0 = sy[E.s0];
  for (;;) {
    if (E.q_n >= 32 || (E.exhausted && E.q_n > 0)) {
      // expensive gates for up to 32 survivors, in order
      if (*list_n + 32 > qcap) return false;
      const int take = E.q_n < 32 ? E.q_n : 32;
      bool valid = false;
      int s1 = 0, d0 = 0, d1 = 0;
      if (F.lane < take) {
        const unsigned e = squeue[(E.q_head + F.lane) & 63];
        s1 = same[e & 0xffu];
        d0 = diff[(e >> 8) & 0xffu];
        d1 = diff[(e >> 16) & 0xffu];
        valid = quad_rest_ok_v(x0, y0, sx[d0], sy[d0], st[d0], sx[s1], sy[s1], sx[d1], sy[d1],
                               st[d1]);
      }
      const unsigned m = __ballot_sync(0xffffffffu, valid);
      if (valid) {
        // winding (detector.rs:571-583)
        const float c0 = cross2(fsub(sx[d0], x0), fsub(sy[d0], y0), fsub(sx[s1], x0), fsub(sy[s1], y0));
        int16_t* q = qlist + 4 * (*list_n + __popc(m & lt));
        q[0] = (int16_t)E.s0;
        q[2] = (int16_t)s1;
        if (c0 > 0.0f) { q[1] = (int16_t)d0; q[3] = (int16_t)d1; }
        else { q[1] = (int16_t)d1; q[3] = (int16_t)d0; }
      }
      *list_n += __popc(m);
      E.q_head = (E.q_head + take) & 63;
      E.q_n -= take;
      __syncwarp();
      continue;
    }
    if (E.exhausted) return true;
    if (!E.a_open) {
      // next `same` entry that passes the (s0, s1)-only gate
      int a = E.a + 1;
      while (a < E.n_same && !((E.diag_ok[a >> 5] >> (a & 31)) & 1u)) ++a;
      E.a = a;
      if (a >= E.n_same) {
        E.exhausted = true;
        continue;
      }
      const int s1 = same[a];
      const float v02x = fsub(sx[s1], x0), v02y = fsub(sy[s1], y0);
      // Per diff entry d: c = cross(v0d, v02) (side of the diagonal s0 -> s1) and the dot gate
      // (saddle.rs:55-59).  The side gate (:40-45) rejects a pair iff c0 * c1 < 0 with
      // c0 = cross(v01, v02) = c[i] and c1 = cross(v02, v03) = -c[j] exactly (the products commute,
      // x - y = -(y - x)), i.e. iff c[i] * c[j] > 0.  Pairs whose c have the same sign and are far
      // from underflowing the product are dropped here; everything else is re-checked exactly by
      // quad_rest_ok, so the prefilter can only pass too much, never too little.
      unsigned long long el = 0ull, pos = 0ull, neg = 0ull;
      for (int base = 0; base < E.n_diff; base += 32) {
        const int d = base + F.lane;
        bool e = false, ps = false, ng = false;
        if (d < E.n_diff) {
          const float vx = dvx[d], vy = dvy[d];
          const float c = cross2(vx, vy, v02x, v02y);
          e = !(dot2(vx, vy, v02x, v02y) < 0.0f);
          ps = c > 1.0e-18f;
          ng = c < -1.0e-18f;
        }
        el |= (unsigned long long)__ballot_sync(0xffffffffu, e) << base;
        pos |= (unsigned long long)__ballot_sync(0xffffffffu, ps) << base;
        neg |= (unsigned long long)__ballot_sync(0xffffffffu, ng) << base;
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = F.lane + 32 * h;
        unsigned long long m = 0ull;
        if (i < E.n_diff && ((el >> i) & 1ull)) {
          const unsigned long long same_side = ((pos >> i) & 1ull) ? pos : (((neg >> i) & 1ull) ? neg : 0ull);
          m = tmask[i] & el & ~same_side;
        }
        E.cand[h] = m;
      }
      // most diagonals leave no pair at all: nothing to queue
      E.a_open = __any_sync(0xffffffffu, (E.cand[0] | E.cand[1]) != 0ull);
      if (!E.a_open) continue;
    }
    // queue the open entry's pairs in combinations(2) order (i ascending, then j ascending), as
    // many as the ring holds
    {
      int space = 64 - E.q_n;
      bool left = false;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        unsigned long long m = E.cand[h];
        const int cnt = __popcll(m);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, o);
          if (F.lane >= o) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        int pos_q = incl - cnt;  // position of this lane's first pair among the half's pairs
        const int i = F.lane + 32 * h;
        while (m != 0ull && pos_q < space) {
          const int j = __ffsll((long long)m) - 1;
          m &= m - 1ull;
          squeue[(E.q_head + E.q_n + pos_q) & 63] = (unsigned)E.a | ((unsigned)i << 8) | ((unsigned)j << 16);
          ++pos_q;
        }
        E.cand[h] = m;
        const int put = total < space ? total : space;
        E.q_n += put;
        space -= put;
        left |= __any_sync(0xffffffffu, m != 0ull);
        if (left) break;  // the second half must wait until the first is queued completely
      }
      if (!left) left = __any_sync(0xffffffffu, E.cand[1] != 0ull);
      E.a_open = left;
      __syncwarp();
    }
  }
}

// Clear this warp's general board state (its lattice region doubles as the group states).
__device__ __forceinline__ void general_state_init(Frame& F) {
  uint32_t* c = (uint32_t*)F.bs.cell;
  const int words = F.lat * F.lat / 2;
  for (int i = F.lane; i < words; i += 32) c[i] = 0u;
  for (int i = F.lane; i < (F.active_words); i += 32) F.bs.active[i] = 0xffffffffu;
  F.bs.n_quads = F.bs.n_touched = F.bs.score = 0;
  __syncwarp();
}

// try_find_best_board (detector.rs:588-639).  Every warp of the block calls it.  Returns 1 with
// the best board (after try_fix_missing) live in warp 0's F.bs, or -1 for None.
// timing taps (warp 0, lane 0 of the frame's block; F.tm may be null)
#define AGB_TM_ADD(slot, v) do { if (F.tm && F.warp == 0 && F.lane == 0) F.tm[slot] += (uint32_t)(v); } while (0)

constexpr int kWaveMax = 16;  // seeds whose quads may be scored side by side
constexpr int kMaxRanges = 32;  // seeds whose quads may share one list batch of a warp (a wave has at most 30)

// try_find_best_board (detector.rs:588-639).  Every warp of the block calls it.  Returns 1 with
// the best board (after try_fix_missing) live in warp 0's F.bs, or -1 for None.
//
// Seeds are taken in the reference's pop order, in WAVES (2, 4, 8, 16, 16 ... seeds; 1 first
// for a single warp).  Inside a wave the warps work independently: a warp claims the next
// seed (slot) of the wave, enumerates its quads into its own list, scores list batches with
// its eight lane groups and keeps the seed's first-maximum quad in the slot.  After the wave
// the slots are merged in seed order with the reference's `score > best_score`, `>= 36` early
// exit and 30-seed limit; seeds of the wave that lie past the early exit were scored in vain
// and are ignored, so the outcome equals the sequential loop.
__device__ __noinline__ int find_best_board_fast(Frame& F) {
  if (F.n == 0) return -1;
  // shared-memory arrays of the block, addressed as such (as_shared)
  // Candidate quads between enumeration and scoring.  First round: 64 quads in shared memory (one
  // or two seeds per warp).  Later rounds visit every seed, each with a dozen short-lived boards:
  // the list moves to this warp's global scratch (the seed-best arrays, free during the search) and
  // holds the quads of ALL the seeds the warp claims in the wave, so the eight lane groups run dry
  // once per wave instead of once per 64 quads.
  uint16_t* qscore = as_shared(F.fx_qscore);
  int16_t* qlist = as_shared(F.fx_qlist);
  int qcap = kQListCap;
  if (F.round > 0) {
    int cap = F.max_quads < F.lat * F.lat ? F.max_quads : F.lat * F.lat;  // entries of the two scratch arrays
    cap = (cap < 512 ? cap : 512) & ~31;
    if (cap > kQListCap) {
      qcap = cap;
      qlist = F.seedbest.quads;
      qscore = (uint16_t*)F.seedbest.vals;
    }
  }
  uint16_t* const wscore = as_shared(F.fx_wscore);
  int16_t* const wquad = as_shared(F.fx_wquad);
  int* const ctl = as_shared(F.ctl);
  long long t0 = clock64();
  if (F.warp == 0) {
    grid_build_warp(F);
    select_seeds(F);
  }
  if (F.lane == 0) *(int*)(F.fx_save0 + (size_t)F.warp * F.fx_save_stride) = 0;  // no board kept yet
  __syncthreads();
  AGB_TM_ADD(1, clock64() - t0);
  int seeds_left = ctl[0];
  F.g_on = ctl[5];
  if (F.g_on) F.g_start = F.g_base + 1;
  // first wave: two seeds (the first board is usually found by the first or second seed); after the
  // first board the leftovers rarely hold another one and every seed will be visited: all (up to
  // 30) seeds form ONE wave, so the warps of the block meet at a barrier once instead of once per
  // wave (the warps claim seeds dynamically; each barrier costs up to one seed of waiting)
  int best_score = 0, count = 0;
  int wave = F.round > 0 ? 30 : (F.n_warps > 1 ? 2 : 1);
  int best_quad[4] = {0, 0, 0, 0};
  SeedEnum E;
  while (seeds_left > 0 && count < 30 && best_score < 36) {
    int nw = wave < seeds_left ? wave : seeds_left;
    nw = nw < 30 - count ? nw : 30 - count;
    // slot t of the wave holds seed F.seeds[seeds_left - 1 - t]
    if (F.warp == 0) {
      if (F.lane < nw) wscore[F.lane] = 0;
      if (F.lane == 0) ctl[kCtlNext] = 0;
    }
    __syncthreads();
    t0 = clock64();
    {
      int list_n = 0, n_rng = 0;
      int rng_slot[kMaxRanges], rng_lo[kMaxRanges], rng_hi[kMaxRanges];
      int t_cur = -1;
      bool begun = false, no_more = false;
      for (;;) {
        // ---- fill the list with the quads of one or more seeds
        for (;;) {
          if (!begun) {
            if (no_more || n_rng == kMaxRanges) break;
            int t = 0;
            if (F.lane == 0) t = atomicAdd(&ctl[kCtlNext], 1);
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t >= nw) { no_more = true; break; }
            t_cur = t;
            seed_enum_begin(F, E, F.seeds[seeds_left - 1 - t]);
            AGB_TM_ADD(8, 1);
            begun = true;
            rng_slot[n_rng] = t;
            rng_lo[n_rng] = rng_hi[n_rng] = list_n;
            ++n_rng;
          }
          const bool seed_done = seed_enum_fill(F, E, qlist, qcap, &list_n);
          rng_hi[n_rng - 1] = list_n;
          if (!seed_done) break;  // list full: score it, then continue with this seed
          begun = false;
        }
        if (list_n == 0 && !begun) break;  // nothing left for this warp in this wave
        AGB_TM_ADD(9, list_n);
        // ---- score the listed quads: eight boards side by side, four lanes each
        const long long ts = clock64();
        warp_score_quads(F, qlist, qscore, list_n);
        __syncwarp();
        bool redo = false;
        for (int k = F.lane; k < list_n; k += 32) redo |= qscore[k] == kScoreRedo;
        if (__any_sync(0xffffffffu, redo)) {
          // boards too large for a group: the general warp-wide build gives the same score
          general_state_init(F);
          for (int k = 0; k < list_n; ++k) {
            if (qscore[k] != kScoreRedo) continue;  // warp-uniform
            int quad[4];
            for (int j = 0; j < 4; ++j) quad[j] = qlist[4 * k + j];
            board_build(F, F.bs, quad);
            const int sc = F.bs.score < kScoreRedo ? F.bs.score : kScoreRedo - 1;
            __syncwarp();
            if (F.lane == 0) qscore[k] = (uint16_t)sc;
            __syncwarp();
          }
          board_reset(F, F.bs);
          AGB_TM_ADD(10, 1);
        }
        AGB_TM_ADD(4, clock64() - ts);
        // ---- per seed: first maximum of its quads in this batch (list order); `>` keeps the
        //      earliest across batches.  The slot belongs to this warp alone during the wave.
        for (int r = 0; r < n_rng; ++r) {
          int bs = -1, bk = kNone;
          for (int k = rng_lo[r] + F.lane; k < rng_hi[r]; k += 32) {
            const int sc = qscore[k];
            if (sc > bs) { bs = sc; bk = k; }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const int os = __shfl_xor_sync(0xffffffffu, bs, o), ok = __shfl_xor_sync(0xffffffffu, bk, o);
            if (os > bs || (os == bs && ok < bk)) { bs = os; bk = ok; }
          }
          const int t = rng_slot[r];
          if (bs > (int)wscore[t]) {
            __syncwarp();
            if (F.lane == 0) wscore[t] = (uint16_t)bs;
            if (F.lane < 4) wquad[4 * t + F.lane] = qlist[4 * bk + F.lane];
          }
          __syncwarp();
        }
        // ---- next batch; a seed that is not finished continues at the head of the list
        list_n = 0;
        n_rng = 0;
        if (begun) {
          rng_slot[0] = t_cur;
          rng_lo[0] = rng_hi[0] = 0;
          n_rng = 1;
        }
      }
    }
    __syncthreads();
    AGB_TM_ADD(3, clock64() - t0);
    // ---- merge the wave in seed order
    for (int t = 0; t < nw; ++t) {
      const int sc = wscore[t];
      if (sc > best_score) {  // best_board_option = Some(board)
        best_score = sc;
        for (int j = 0; j < 4; ++j) best_quad[j] = wquad[4 * t + j];
      }
      if (best_score >= 36) break;
      ++count;
    }
    seeds_left -= nw;
    wave = wave < 4 ? 4 : (wave < kWaveMax ? wave * 2 : wave);
    __syncthreads();  // fx_wscore / fx_wquad are reused by the next wave
  }
  if (best_score == 0) return -1;
  t0 = clock64();
  if (F.warp == 0) {
    general_state_init(F);
    // the winner may be one of the boards whose state a warp kept: convert it
    const uint8_t* hit = nullptr;
    if (best_score >= kSaveMin)
      for (int w = 0; w < F.n_warps; ++w) {
        const uint8_t* sv = F.fx_save0 + (size_t)w * F.fx_save_stride;
        const int16_t* sq = (const int16_t*)(sv + 4);
        if (*(const int*)sv == best_score && sq[0] == best_quad[0] && sq[1] == best_quad[1] &&
            sq[2] == best_quad[2] && sq[3] == best_quad[3]) {
          hit = sv + 16;
          break;
        }
      }
    if (hit) {
      BoardState& B = F.bs;
      const int16_t* sq = (const int16_t*)(hit + 256);
      for (int i = F.lane; i < best_score * 4; i += 32) B.quads[i] = sq[i];
      int n_t = 0;
      for (int base = 0; base < kGWin * kGWin; base += 32) {  // window cell (x + 8) * 16 + (y + 8)
        const int wc = base + F.lane;
        const int v = hit[wc];
        const unsigned m = __ballot_sync(0xffffffffu, v != 0);
        if (v != 0) {
          const int ci = cell_index(F, (wc >> 4) - 8, (wc & 15) - 8);
          B.cell[ci] = v == 0xff ? (int16_t)-1 : (int16_t)v;
          B.touched[n_t + __popc(m & ((1u << F.lane) - 1u))] = (int16_t)ci;
        }
        n_t += __popc(m);
      }
      B.n_touched = n_t;
      B.n_quads = best_score;
      B.score = best_score;
      __syncwarp();
      AGB_TM_ADD(12, 100);
    } else {
      board_build(F, F.bs, best_quad);
    }
    board_fix_missing(F, F.bs);
  }
  AGB_TM_ADD(6, clock64() - t0);
  return 1;
}

}  // namespace agb
#endif  // AGB_DEVICE
