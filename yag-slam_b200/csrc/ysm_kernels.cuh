// ysm_kernels.cuh -- sm_100a device code of the correlative scan matcher.
//
// Every kernel cites the Karto function it replaces (the Karto C++
// itself is not in the reference tree -- it ships as the wheel karto_scanmatcher==1.0.0,
// reference setup.py:46 -- so citations are to the survey's restatement and to the reference's
// python twins in yag_slam/helpers.py).
//
// Compile with -fmad=false: cell indices come from Round(double) and must see the same
// separately-rounded IEEE-754 operations as an x86-64 build of Karto (no FMA contraction).
// No device sin/cos/atan2 anywhere: every transcendental is evaluated by the host runtime with
// libm and handed in as a table, so results are bit-identical to the CPU reference.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ysm {

#define YSM_KT_TOLERANCE 1e-06
#define YSM_INVALID_CELL 0xFFFFFFFFu
#define YSM_TILE 32  // correlation-grid tile edge (cells) of the build / clear kernels

// Sizes of one correlation grid (ScanMatcher::Create / CorrelationGrid).
struct GridC {
  int roi, border, stride, width, height, data_size;
  int half_kernel, K, Wt;  // Wt = row width (cells) of the pre-shifted stamp table
  int stride4;             // stride / 4
  double scale;            // 1 / resolution
  long long grid_bytes;    // bytes between consecutive slots (16-B aligned)
};

struct PenaltyC {
  double distance_variance_penalty, angle_variance_penalty;
  double minimum_distance_penalty, minimum_angle_penalty;
};

// One MatchScan call of the current wave (grid build inputs).
struct MatchDev {
  int slot;
  int base_begin, base_end;  // range in the wave's base_idx array (pool scan ids)
  int cells_off;             // offset of this match's cell list / point scratch
  int gbox_off;              // offset of this match's per-32-cell bounding boxes
  int pad0;
  double vpx, vpy;           // FindValidPoints viewpoint = query sensor position
  double gox, goy;           // CorrelationGrid converter offset
};

// One GridIndexLookup::ComputeOffsets table.
struct TableDev {
  int q_start, P, Ppad, nA;
  int trig_off;  // cos/sin of each search angle: trig[2*(trig_off+a)], trig[2*(trig_off+a)+1]
  int out_off;   // int32 element offset into the offsets buffer (multiple of 4)
  double px, py;                // scan sensor position
  double r00, r01, r10, r11;    // Transform(sensorPose).m_InverseRotation (host libm)
  double gox, goy;
};

// One CorrelateScan call.
struct PassDev {
  int slot, table, nA, nX, nY, P, Ppad, fine, penalize;
  int sums_off;   // u32 element offset into the sums buffer, layout [iy][ix][a]
  int htrig_off;  // cos/sin of NormalizeAngle(angle_a), for the tie average
  int ang_off;    // int element offset into the angular-covariance sums buffer
  // speculative fine pass of a coarse pass (latency path): pass id (-1: none) and where, in the
  // trig array, the host left the per-winning-angle data: heading table [nA] (double index),
  // cos/sin blocks [nA][nAf] of the fine search angles and of their normalised headings (pair index)
  int spec, spec_nAf, spec_h_off, spec_trig_off, spec_htrig_off;
  int cmax_off;   // element offset of this (coarse) pass's per-cell maxima [nY][nX]; -1: none
  double cx, cy, ch;              // search centre
  double offx, offy, resx, resy;  // search space offset / resolution
  double angle_offset, angle_res;
  double gox, goy;
};

// Per-pass reduction result, finished on the host (atan2 via libm).
struct PassOut {
  double best;          // best response before the clamp to 1
  double avg_x, avg_y;  // tie-averaged position
  double tx, ty;        // mean cos / mean sin of the tied headings
  double norm, axx, axy, ayy;  // ComputePositionalCovariance accumulators
  int n_ties;     // -1: speculative pass that was not run
  int first_idx;  // storage index (iy, ix, a) of the first tied pose
};

__device__ __forceinline__ double kt_round(double v) { return v >= 0.0 ? floor(v + 0.5) : ceil(v - 0.5); }
__device__ __forceinline__ bool kt_double_equal(double a, double b) { return fabs(a - b) <= YSM_KT_TOLERANCE; }
__device__ __forceinline__ int world_to_grid1(double w, double off, double scale) {
  return (int)kt_round((w - off) * scale);
}

// Byte-wise max of two packed u8x4 words whose bytes are all < 128 (grid values are <= 100).
__device__ __forceinline__ uint32_t vmax4_lt128(uint32_t a, uint32_t b) {
  uint32_t d = (a | 0x80808080u) - b;  // per byte: a + 128 - b, never borrows
  uint32_t m = ((d & 0x80808080u) >> 7) * 0xFFu;  // 0xFF where a >= b
  return (a & m) | (b & ~m);
}

// ---------------------------------------------------------------------------------------------
// K1a  ScanMatcher::FindValidPoints + the WorldToGrid/ROI test of AddScan (python
// analogue yag_slam/helpers.py:298-329). One CTA per match, one warp per base scan.
// The sequential "trailing iterator" filter is restated as: next[i] = first j>i farther than
// 10 cm from point i (parallel), the trigger chain 0 -> next[0] -> ... (one lane, shared
// memory pointer chase), then every segment [t_k, t_k+1) is kept iff the side test ss >= 0
// (parallel). Output: the match's occupied cells in Karto's processing order.
// ---------------------------------------------------------------------------------------------
// warp-aggregated bump of the 16-bit counter of tile t (t < 0: this lane has none): lanes naming the same
// tile are served by ONE shared-memory atomic; returns this lane's slot = old counter value + its rank
__device__ __forceinline__ unsigned tile_counter_bump(unsigned* s_cnt, int t, int lane) {
  const unsigned grp = __match_any_sync(0xffffffffu, t);
  const int leader = __ffs(grp) - 1, sh = 16 * (t & 1);
  unsigned old = 0u;
  if (t >= 0 && lane == leader) old = atomicAdd(&s_cnt[t >> 1], (unsigned)__popc(grp) << sh);
  old = __shfl_sync(0xffffffffu, old, leader);  // full-mask shuffle: every lane reads its group's leader
  return ((old >> sh) & 0xFFFFu) + (unsigned)__popc(grp & ((1u << lane) - 1u));
}

// the same for the (up to) 2 x 2 tiles of one cell at once: the four match / atomic / shuffle chains are
// issued back to back so their latencies overlap
__device__ __forceinline__ void tile_counter_bump4(unsigned* s_cnt, const int (&t)[4], int lane, unsigned (&slot)[4]) {
  unsigned grp[4], old[4];
#pragma unroll
  for (int k = 0; k < 4; k++) grp[k] = __match_any_sync(0xffffffffu, t[k]);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    old[k] = 0u;
    if (t[k] >= 0 && lane == __ffs(grp[k]) - 1)
      old[k] = atomicAdd(&s_cnt[t[k] >> 1], (unsigned)__popc(grp[k]) << (16 * (t[k] & 1)));
  }
#pragma unroll
  for (int k = 0; k < 4; k++) old[k] = __shfl_sync(0xffffffffu, old[k], __ffs(grp[k]) - 1);
#pragma unroll
  for (int k = 0; k < 4; k++)
    slot[k] = ((old[k] >> (16 * (t[k] & 1))) & 0xFFFFu) + (unsigned)__popc(grp[k] & ((1u << lane) - 1u));
}

__device__ __forceinline__ void
find_valid_body(const GridC& g, const MatchDev* matches, const int* base_idx, const int* scan_start,
                const int* scan_count, const double* pool, uint32_t* pt_cell, uint32_t* cells, int* cell_count,
                uint2* gbox, int2* work, int* work_count, int pmax, int nbase_max, int stage, int vbx,
                unsigned char* smem_raw, int fv_warps = 0, uint32_t* cand = nullptr, uint2* wcand = nullptr,
                int stamp_tiles_axis = 0) {
  // cand != nullptr (throughput path): besides the work list, every touched tile gets the exact list of
  // the cells whose stamp reaches it (wcand[work item] = {offset, count} into cand), so the stamping
  // kernel needs no candidate search. The per-tile flags are then 16-bit counters / fill cursors.
  __shared__ int s_tile_total, s_tile_base;
  __shared__ int s_wsum_t[32], s_wsum_c[32];
  // fv_warps > 0: only that many warps take scans (the shared arrays are sized for them)
  const int nwarps = fv_warps > 0 ? min(fv_warps, (int)(blockDim.x >> 5)) : (int)(blockDim.x >> 5);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned short* s_next = reinterpret_cast<unsigned short*>(smem_raw) + (size_t)warp * 2 * pmax;
  unsigned short* s_trig = s_next + pmax;
  int* s_scan_emit = reinterpret_cast<int*>(smem_raw + (size_t)nwarps * 4 * pmax);  // [nbase_max]
  int* s_dir = s_scan_emit + nbase_max;  // [3][nbase_max]: point count, pool start, prefix of counts of the base scans
  unsigned* s_bits = reinterpret_cast<unsigned*>(s_dir + 3 * nbase_max);           // touched-tile flags, one BYTE per tile
  unsigned char* s_flag = reinterpret_cast<unsigned char*>(s_bits);                // (plain stores: no atomics)
  const int tnx = (g.width + YSM_TILE - 1) / YSM_TILE;
  const int nbitw = cand ? (tnx * tnx + 1) >> 1 : (tnx * tnx + 3) >> 2;  // words of 4 byte flags / 2 u16 counters
  // small waves: each warp stages its scan's points in shared memory (SoA) so the filter's
  // dependent loads are LDS instead of L2 round trips
  double* s_px = reinterpret_cast<double*>(smem_raw + (((size_t)nwarps * 4 * pmax + 16 * (size_t)nbase_max + 4 * (size_t)nbitw + 15) & ~(size_t)15)) + (size_t)warp * 2 * pmax;
  double* s_py = s_px + pmax;
  for (int i = threadIdx.x; i < nbitw; i += blockDim.x) s_bits[i] = 0u;
  if (threadIdx.x == 0) s_tile_total = 0;
  __syncthreads();
#define YSM_PX(i) (stage ? s_px[i] : pts[2 * (i)])
#define YSM_PY(i) (stage ? s_py[i] : pts[2 * (i) + 1])

  const MatchDev m = matches[vbx];
  const int nbase = m.base_end - m.base_begin;
  const double msd = 0.1 * 0.1;  // math::Square(0.1)

  // directory of the match's base scans (one round of loads), scan-local offsets = prefix of counts
  for (int b = threadIdx.x; b < nbase; b += blockDim.x) {
    const int s = base_idx[m.base_begin + b];
    s_dir[b] = scan_count[s];
    s_dir[nbase_max + b] = scan_start[s];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int off = 0;
    for (int b = 0; b < nbase; b++) {
      s_dir[2 * nbase_max + b] = off;
      off += s_dir[b];
    }
  }
  __syncthreads();
  for (int b = warp; b < nbase && warp < nwarps; b += nwarps) {
    const int off = s_dir[2 * nbase_max + b];
    const int n = s_dir[b];
    const double* pts = pool + 2 * (size_t)s_dir[nbase_max + b];
    uint32_t* out = pt_cell + m.cells_off + off;
    for (int i = lane; i < n; i += 32) out[i] = YSM_INVALID_CELL;
    int emitted = 0;
    if (n > 0) {
      if (stage) {
        for (int i = lane; i < n; i += 32) {
          s_px[i] = pts[2 * i];
          s_py[i] = pts[2 * i + 1];
        }
        __syncwarp();
      }
      for (int i = lane; i < n; i += 32) {
        const double fx = YSM_PX(i), fy = YSM_PY(i);
        int j = i + 1;
        while (j < n) {
          const double dx = fx - YSM_PX(j), dy = fy - YSM_PY(j);
          if (dx * dx + dy * dy > msd) break;
          j++;
        }
        s_next[i] = (unsigned short)j;
      }
      __syncwarp();
      // the trigger chain 0 -> next[0] -> ... , 32 points at a time: inside a block of 32 the points the chain
      // visits are found by pointer doubling on the lanes (five rounds of SHFL + REDUX.OR), the chain leaves the
      // block through the next[] of its last point (r02zj: one lane chasing the pointers through shared memory
      // was 14 % of the kernel's warp time)
      int ntrig = 0;
      for (int e = 0; e < n;) {  // e: the chain's entry into the block that holds it (warp-uniform)
        const int b0 = e & ~31;
        const int i = b0 + lane;
        const int nx = i < n ? (int)s_next[i] : n;  // > i
        int J = nx < n ? nx - b0 : 64;              // in-block target lane, or >= 32: the chain leaves the block / ends
        unsigned reach = 1u << (e - b0);
#pragma unroll
        for (int r = 0; r < 5; r++) {
          const bool on = (reach >> lane) & 1u;
          reach |= __reduce_or_sync(0xffffffffu, (on && J < 32) ? (1u << J) : 0u);
          const int JJ = __shfl_sync(0xffffffffu, J, J & 31);
          if (J < 32) J = JJ;
        }
        if ((reach >> lane) & 1u) s_trig[ntrig + __popc(reach & ((1u << lane) - 1u))] = (unsigned short)i;
        ntrig += __popc(reach);
        e = __shfl_sync(0xffffffffu, nx, 31 - __clz(reach));  // where the chain's last point in this block points
      }
      __syncwarp();
      for (int k = lane; k < ntrig - 1; k += 32) {
        const int f = s_trig[k], c = s_trig[k + 1];
        const double fx = YSM_PX(f), fy = YSM_PY(f);
        const double cx = YSM_PX(c), cy = YSM_PY(c);
        const double a = m.vpy - fy;
        const double b2 = fx - m.vpx;
        const double cc = fy * m.vpx - fx * m.vpy;
        const double ss = cx * a + cy * b2 + cc;
        if (!(ss < 0.0)) {
          for (int j = f; j < c; j++) {
            const double vx = (YSM_PX(j) - m.gox) * g.scale;
            const double vy = (YSM_PY(j) - m.goy) * g.scale;
            if (vx > -1.0 && vy > -1.0 && vx < 1e9 && vy < 1e9) {
              const int gx = (int)kt_round(vx), gy = (int)kt_round(vy);
              if (gx >= 0 && gx < g.roi && gy >= 0 && gy < g.roi) {
                const int ax = gx + g.border, ay = gy + g.border;
                out[j] = (uint32_t)ax | ((uint32_t)ay << 16);
                emitted++;
                // every tile the K x K stamp of this cell overlaps
                const int tx0 = (ax - g.half_kernel) / YSM_TILE, tx1 = (ax + g.half_kernel) / YSM_TILE;
                const int ty0 = (ay - g.half_kernel) / YSM_TILE, ty1 = (ay + g.half_kernel) / YSM_TILE;
                for (int ty = ty0; ty <= ty1; ty++)
                  for (int tx = tx0; tx <= tx1; tx++) {
                    if (!cand) s_flag[ty * tnx + tx] = 1;
                  }
              }
            }
          }
        }
      }
    }
    for (int o = 16; o > 0; o >>= 1) emitted += __shfl_xor_sync(0xffffffffu, emitted, o);
    if (lane == 0) s_scan_emit[b] = emitted;
    __syncwarp();
  }
  __syncthreads();
  // ordered compaction (scan order, then point order)
  for (int b = warp; b < nbase && warp < nwarps; b += nwarps) {
    int dst = 0;
    for (int bb = 0; bb < b; bb++) dst += s_scan_emit[bb];
    const int off = s_dir[2 * nbase_max + b];
    const int n = s_dir[b];
    const uint32_t* in = pt_cell + m.cells_off + off;
    uint32_t* outc = cells + m.cells_off;
    for (int i0 = 0; i0 < n; i0 += 32) {
      const int i = i0 + lane;
      const uint32_t c = (i < n) ? in[i] : YSM_INVALID_CELL;
      const unsigned bal = __ballot_sync(0xffffffffu, c != YSM_INVALID_CELL);
      if (c != YSM_INVALID_CELL) outc[dst + __popc(bal & ((1u << lane) - 1u))] = c;
      dst += __popc(bal);
    }
  }
  int tot = 0;
  for (int b = 0; b < nbase; b++) tot += s_scan_emit[b];
  if (threadIdx.x == 0) cell_count[vbx] = tot;
  if (cand) {
    // ---- exact per-tile candidate lists ------------------------------------------------------------
    // every thread owns a contiguous chunk of counter words: block-wide exclusive scan of
    // (touched tiles, candidates), then work items + list offsets, then the fill
    const int tps1 = stamp_tiles_axis;  // tiles a stamp can span per axis
    const uint32_t* mc = cells + m.cells_off;
    uint32_t* rank01 = pt_cell + m.cells_off;                                // (tps1 == 2)
    uint32_t* rank23 = reinterpret_cast<uint32_t*>(gbox) + m.cells_off;      // sized for it by the host in this path
    __syncthreads();  // the compacted cells (global, written by this CTA) are complete
    for (int i0 = 0; i0 < tot; i0 += (int)blockDim.x) {  // count pass (block-uniform trip count)
      const int i = i0 + (int)threadIdx.x;
      int ax = 0, ay = 0;
      if (i < tot) {
        const uint32_t c = mc[i];
        ax = (int)(c & 0xFFFFu);
        ay = (int)(c >> 16);
      }
      const int tx0 = (ax - g.half_kernel) / YSM_TILE, tx1 = (ax + g.half_kernel) / YSM_TILE;
      const int ty0 = (ay - g.half_kernel) / YSM_TILE, ty1 = (ay + g.half_kernel) / YSM_TILE;
      if (tps1 == 2) {
        int t4[4];
        unsigned sl[4];
#pragma unroll
        for (int k = 0; k < 4; k++)
          t4[k] = (i < tot && ty0 + (k >> 1) <= ty1 && tx0 + (k & 1) <= tx1) ? (ty0 + (k >> 1)) * tnx + tx0 + (k & 1) : -1;
        tile_counter_bump4(s_bits, t4, lane, sl);
        // the slot a cell got in each tile's count IS its rank in that tile's list: kept (the point scratch and
        // the bounding-box buffer are free in this path) so the fill pass needs no second round of match / atomics
        if (i < tot) {
          rank01[i] = (sl[0] & 0xFFFFu) | (sl[1] << 16);
          rank23[i] = (sl[2] & 0xFFFFu) | (sl[3] << 16);
        }
      } else {
        for (int dy = 0; dy < tps1; dy++)
          for (int dx = 0; dx < tps1; dx++) {
            const int t = (i < tot && ty0 + dy <= ty1 && tx0 + dx <= tx1) ? (ty0 + dy) * tnx + tx0 + dx : -1;
            tile_counter_bump(s_bits, t, lane);
          }
      }
    }
    __syncthreads();
    const int per = (nbitw + (int)blockDim.x - 1) / (int)blockDim.x;
    const int w0 = min(nbitw, (int)threadIdx.x * per), w1 = min(nbitw, w0 + per);
    int nt = 0, nc = 0;
    for (int w = w0; w < w1; w++) {
      const unsigned v = s_bits[w], a = v & 0xFFFFu, bb = v >> 16;
      nt += (a != 0u) + (bb != 0u);
      nc += (int)(a + bb);
    }
    int it = nt, ic = nc;  // inclusive scans inside the warp
    for (int o = 1; o < 32; o <<= 1) {
      const int ut = __shfl_up_sync(0xffffffffu, it, o), uc = __shfl_up_sync(0xffffffffu, ic, o);
      if (lane >= o) { it += ut; ic += uc; }
    }
    if (lane == 31) { s_wsum_t[warp] = it; s_wsum_c[warp] = ic; }
    __syncthreads();
    int pt = 0, pc = 0, tt = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) {
      if (w < warp) { pt += s_wsum_t[w]; pc += s_wsum_c[w]; }
      tt += s_wsum_t[w];
    }
    if (threadIdx.x == 0) s_tile_base = tt ? atomicAdd(work_count, tt) : 0;
    __syncthreads();
    int pos = s_tile_base + pt + it - nt;
    unsigned off = (unsigned)(pc + ic - nc);
    const unsigned cbase = (unsigned)m.cells_off * (unsigned)(tps1 * tps1);
    for (int w = w0; w < w1; w++) {
      const unsigned v = s_bits[w], a = v & 0xFFFFu, bb = v >> 16;
      unsigned cur = 0u;
      if (a) {
        work[pos] = make_int2(vbx, 2 * w);
        wcand[pos] = make_uint2(cbase + off, a);
        cur |= off;
        off += a;
        pos++;
      }
      if (bb) {
        work[pos] = make_int2(vbx, 2 * w + 1);
        wcand[pos] = make_uint2(cbase + off, bb);
        cur |= off << 16;
        off += bb;
        pos++;
      }
      s_bits[w] = cur;  // fill cursors (list-relative; < 65536 by the host's guard)
    }
    __syncthreads();
    for (int i0 = 0; i0 < tot; i0 += (int)blockDim.x) {  // fill pass
      const int i = i0 + (int)threadIdx.x;
      uint32_t c = 0u;
      int ax = 0, ay = 0;
      if (i < tot) {
        c = mc[i];
        ax = (int)(c & 0xFFFFu);
        ay = (int)(c >> 16);
      }
      const int tx0 = (ax - g.half_kernel) / YSM_TILE, tx1 = (ax + g.half_kernel) / YSM_TILE;
      const int ty0 = (ay - g.half_kernel) / YSM_TILE, ty1 = (ay + g.half_kernel) / YSM_TILE;
      if (tps1 == 2) {
        if (i < tot) {
          const uint32_t r01 = rank01[i], r23 = rank23[i];
#pragma unroll
          for (int k = 0; k < 4; k++) {
            if (ty0 + (k >> 1) <= ty1 && tx0 + (k & 1) <= tx1) {
              const int t = (ty0 + (k >> 1)) * tnx + tx0 + (k & 1);
              const unsigned off = (s_bits[t >> 1] >> (16 * (t & 1))) & 0xFFFFu;  // the tile's list offset (read-only now)
              const unsigned rk = ((k < 2 ? r01 : r23) >> (16 * (k & 1))) & 0xFFFFu;
              cand[cbase + off + rk] = c;
            }
          }
        }
      } else {
        for (int dy = 0; dy < tps1; dy++)
          for (int dx = 0; dx < tps1; dx++) {
            const int t = (i < tot && ty0 + dy <= ty1 && tx0 + dx <= tx1) ? (ty0 + dy) * tnx + tx0 + dx : -1;
            const unsigned slot = tile_counter_bump(s_bits, t, lane);
            if (t >= 0) cand[cbase + slot] = c;
          }
      }
    }
    return;
  }
  // touched tiles -> the wave's (match, tile) work list
  int cnt = 0;
  for (int i = threadIdx.x; i < nbitw; i += blockDim.x) cnt += __popc(s_bits[i]);
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0 && cnt) atomicAdd(&s_tile_total, cnt);
  __syncthreads();  // also orders the compacted cells before the bounding-box pass below
  if (threadIdx.x == 0) s_tile_base = s_tile_total ? atomicAdd(work_count, s_tile_total) : 0;
  if (threadIdx.x == 0) s_tile_total = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < nbitw; i += blockDim.x) {
    unsigned b = s_bits[i];
    if (!b) continue;
    int pos = s_tile_base + atomicAdd(&s_tile_total, __popc(b));
    while (b) {
      const int bit = __ffs(b) - 1;  // flags are 0 / 1: bit 8k set <=> flag k
      b &= b - 1;
      work[pos++] = make_int2(vbx, i * 4 + (bit >> 3));
    }
  }
  // bounding box of every group of 32 consecutive cells (scan order keeps them spatially close)
  __threadfence_block();
  const uint32_t* mcells = cells + m.cells_off;
  for (int g0 = warp * 32; g0 < tot && warp < nwarps; g0 += nwarps * 32) {
    const int i = g0 + lane;
    int xlo = 0xFFFF, xhi = 0, ylo = 0xFFFF, yhi = 0;
    if (i < tot) {
      const uint32_t c = mcells[i];
      xlo = xhi = (int)(c & 0xFFFFu);
      ylo = yhi = (int)(c >> 16);
    }
    for (int o = 16; o > 0; o >>= 1) {
      xlo = min(xlo, __shfl_xor_sync(0xffffffffu, xlo, o));
      xhi = max(xhi, __shfl_xor_sync(0xffffffffu, xhi, o));
      ylo = min(ylo, __shfl_xor_sync(0xffffffffu, ylo, o));
      yhi = max(yhi, __shfl_xor_sync(0xffffffffu, yhi, o));
    }
    if (lane == 0) gbox[m.gbox_off + (g0 >> 5)] = make_uint2((uint32_t)xlo | ((uint32_t)xhi << 16), (uint32_t)ylo | ((uint32_t)yhi << 16));
  }
}

// ---------------------------------------------------------------------------------------------
// K1a (latency path)  FindValidPoints with a whole CTA per base scan, then a CTA per match.
// fv_scan_body: points staged in shared memory; next[i] in parallel; the trigger chain
// 0 -> next[0] -> ... is found by pointer doubling (log2 n rounds) instead of a serial walk; every
// point finds its trigger by stepping back, evaluates the side test and its cell; touched tiles are
// flagged with plain byte stores; ordered block-wide compaction into the scan's region.
// fv_match_body: concatenates the scans' cells in scan order, bounding boxes, (match, tile) work list.
// ---------------------------------------------------------------------------------------------
struct ScanRef {
  int match, b, off, pad;  // match of the wave, base-scan ordinal, prefix of the point counts inside the match
};

__host__ __device__ __forceinline__ size_t fv_scan_smem(int pmax) {
  return ((size_t)pmax * (8 + 8 + 2 + 2 + 2 + 1 + 1) + 256 + 15) & ~(size_t)15;
}

__device__ __forceinline__ void
fv_scan_body(const GridC& g, const MatchDev* matches, const int* base_idx, const int* scan_start,
             const int* scan_count, const double* pool, const ScanRef& sr, uint32_t* pt_cell, int* scan_emit, int v,
             unsigned char* tileflag, int tiles_per_grid, int pmax, unsigned char* dsm) {
  __shared__ int s_wtot[32];
  __shared__ int s_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  double* s_px = reinterpret_cast<double*>(dsm);
  double* s_py = s_px + pmax;
  unsigned short* s_next = reinterpret_cast<unsigned short*>(s_py + pmax);
  unsigned short* s_ja = s_next + pmax;
  unsigned short* s_jb = s_ja + pmax;
  unsigned char* s_mark = reinterpret_cast<unsigned char*>(s_jb + pmax);
  const MatchDev m = matches[sr.match];
  const int s = base_idx[m.base_begin + sr.b];
  const int n = scan_count[s];
  const double* pts = pool + 2 * (size_t)scan_start[s];
  const int tnx = (g.width + YSM_TILE - 1) / YSM_TILE;
  const double msd = 0.1 * 0.1;  // math::Square(0.1)
  for (int i = tid; i < n; i += blockDim.x) {
    const double2 w = *reinterpret_cast<const double2*>(pts + 2 * (size_t)i);
    s_px[i] = w.x;
    s_py[i] = w.y;
    s_mark[i] = i == 0;
  }
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) {
    const double fx = s_px[i], fy = s_py[i];
    int j = i + 1;
    while (j < n) {
      const double dx = fx - s_px[j], dy = fy - s_py[j];
      if (dx * dx + dy * dy > msd) break;
      j++;
    }
    s_next[i] = (unsigned short)j;
    s_ja[i] = (unsigned short)j;
  }
  __syncthreads();
  // reachability from point 0 by pointer doubling: after round k every point within 2^(k+1) - 1 hops is marked
  unsigned short* ja = s_ja;
  unsigned short* jb = s_jb;
  for (int span = 1; span < n; span <<= 1) {
    for (int i = tid; i < n; i += blockDim.x) {
      const int j = ja[i];
      if (j < n) {
        // (racecheck reports this byte as a read/write hazard: another thread may set s_mark[i] in this
        // same round. The race is benign -- marks only ever go 0 -> 1 and an early-seen mark is a true one
        // (i reachable => ja[i] reachable), so the set reached after the last round is the same.)
        if (s_mark[i]) s_mark[j] = 1;
        jb[i] = ja[j];
      } else {
        jb[i] = (unsigned short)n;
      }
    }
    __syncthreads();
    unsigned short* t = ja; ja = jb; jb = t;
  }
  // per point: its trigger, the side test of the trigger's segment, its cell; ordered compaction
  uint32_t* out = pt_cell + m.cells_off + sr.off;
  for (int j0 = 0; j0 < n; j0 += blockDim.x) {
    const int j = j0 + tid;
    uint32_t cell = YSM_INVALID_CELL;
    if (j < n) {
      int f = j;
      while (!s_mark[f]) f--;
      const int c = s_next[f];
      if (c < n) {  // points after the last trigger are never emitted
        const double fx = s_px[f], fy = s_py[f];
        const double a = m.vpy - fy;
        const double b2 = fx - m.vpx;
        const double cc = fy * m.vpx - fx * m.vpy;
        const double ss = s_px[c] * a + s_py[c] * b2 + cc;
        if (!(ss < 0.0)) {
          const double vx = (s_px[j] - m.gox) * g.scale;
          const double vy = (s_py[j] - m.goy) * g.scale;
          if (vx > -1.0 && vy > -1.0 && vx < 1e9 && vy < 1e9) {
            const int gx = (int)kt_round(vx), gy = (int)kt_round(vy);
            if (gx >= 0 && gx < g.roi && gy >= 0 && gy < g.roi) {
              const int ax = gx + g.border, ay = gy + g.border;
              cell = (uint32_t)ax | ((uint32_t)ay << 16);
              const int tx0 = (ax - g.half_kernel) / YSM_TILE, tx1 = (ax + g.half_kernel) / YSM_TILE;
              const int ty0 = (ay - g.half_kernel) / YSM_TILE, ty1 = (ay + g.half_kernel) / YSM_TILE;
              unsigned char* tf = tileflag + (size_t)sr.match * tiles_per_grid;
              for (int ty = ty0; ty <= ty1; ty++)
                for (int tx = tx0; tx <= tx1; tx++) tf[ty * tnx + tx] = 1;
            }
          }
        }
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, cell != YSM_INVALID_CELL);
    if (lane == 0) s_wtot[warp] = __popc(bal);
    __syncthreads();
    int before = s_base;
    for (int w = 0; w < warp; w++) before += s_wtot[w];
    if (cell != YSM_INVALID_CELL) out[before + __popc(bal & ((1u << lane) - 1u))] = cell;
    __syncthreads();
    if (tid == 0) {
      int tot = s_base;
      for (int w = 0; w < nwarps; w++) tot += s_wtot[w];
      s_base = tot;
    }
    __syncthreads();
  }
  if (tid == 0) scan_emit[v] = s_base;
}

// Grid-wide: every CTA rebuilds the (tiny) emit-count prefix of match vbx, then its warps take
// tasks -- a group of 32 final cells (gathered from the scans' regions, bounding box) or a slice of
// the tile flags (work-list entries). vblock / nblocks: this CTA's rank among those working on the match.
__device__ __forceinline__ void
fv_match_body(const GridC& g, const MatchDev* matches, const ScanRef* scanlist, int nscans_total, const int* scan_emit,
              const uint32_t* pt_cell, uint32_t* cells, int* cell_count, uint2* gbox, int2* work, int* work_count,
              const unsigned char* tileflag, int tiles_per_grid, int vbx, int vblock, int nblocks) {
  __shared__ int s_src[64], s_dst[65], s_first;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const MatchDev m = matches[vbx];
  if (tid == 0) s_first = nscans_total;
  __syncthreads();
  // scans of this match are consecutive in the scan list, in base order
  if (tid < nscans_total && scanlist[tid].match == vbx) atomicMin(&s_first, tid);
  __syncthreads();
  const int v0 = s_first;
  const int nb = min(64, m.base_end - m.base_begin);
  if (tid < nb) {
    s_src[tid] = scanlist[v0 + tid].off;
    s_dst[tid + 1] = __ldcg(scan_emit + v0 + tid);
  }
  __syncthreads();
  if (tid == 0) {
    s_dst[0] = 0;
    for (int b = 0; b < nb; b++) s_dst[b + 1] += s_dst[b];
  }
  __syncthreads();
  const int tot = nb > 0 ? s_dst[nb] : 0;
  if (vblock == 0 && tid == 0) cell_count[vbx] = tot;
  const int ngroups = (tot + 31) >> 5;
  const unsigned* tf = reinterpret_cast<const unsigned*>(tileflag + (size_t)vbx * tiles_per_grid);
  const int nw4 = (tiles_per_grid + 3) >> 2, nflag = (nw4 + 31) >> 5;
  for (int t = vblock * nwarps + warp; t < ngroups + nflag; t += nblocks * nwarps) {
    if (t < ngroups) {
      // 32 consecutive cells of the match (scan order, then point order) + their bounding box
      const int i = t * 32 + lane;
      int xlo = 0xFFFF, xhi = 0, ylo = 0xFFFF, yhi = 0;
      if (i < tot) {
        int b = 0;
        while (s_dst[b + 1] <= i) b++;
        const uint32_t c = __ldcg(pt_cell + m.cells_off + s_src[b] + (i - s_dst[b]));
        cells[m.cells_off + i] = c;
        xlo = xhi = (int)(c & 0xFFFFu);
        ylo = yhi = (int)(c >> 16);
      }
      for (int o = 16; o > 0; o >>= 1) {
        xlo = min(xlo, __shfl_xor_sync(0xffffffffu, xlo, o));
        xhi = max(xhi, __shfl_xor_sync(0xffffffffu, xhi, o));
        ylo = min(ylo, __shfl_xor_sync(0xffffffffu, ylo, o));
        yhi = max(yhi, __shfl_xor_sync(0xffffffffu, yhi, o));
      }
      if (lane == 0) gbox[m.gbox_off + t] = make_uint2((uint32_t)xlo | ((uint32_t)xhi << 16), (uint32_t)ylo | ((uint32_t)yhi << 16));
    } else {
      // 32 words of tile flags -> (match, tile) work-list entries
      const int i = (t - ngroups) * 32 + lane;
      unsigned b = i < nw4 ? __ldcg(tf + i) : 0u;
      int cnt = __popc(b), pre = cnt;
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, pre, o);
        if (lane >= o) pre += v;
      }
      const int wtot = __shfl_sync(0xffffffffu, pre, 31);
      int base = 0;
      if (lane == 0 && wtot) base = atomicAdd(work_count, wtot);
      base = __shfl_sync(0xffffffffu, base, 0);
      int pos = base + pre - cnt;
      while (b) {
        const int bit = __ffs(b) - 1;  // flags are 0 / 1: bit 8k set <=> flag k
        b &= b - 1;
        work[pos++] = make_int2(vbx, i * 4 + (bit >> 3));
      }
    }
  }
}

__global__ void __launch_bounds__(1024)
k_find_valid(GridC g, const MatchDev* __restrict__ matches, const int* __restrict__ base_idx,
             const int* __restrict__ scan_start, const int* __restrict__ scan_count,
             const double* __restrict__ pool, uint32_t* __restrict__ pt_cell,
             uint32_t* __restrict__ cells, int* __restrict__ cell_count, uint2* __restrict__ gbox,
             int2* __restrict__ work, int* __restrict__ work_count, int pmax, int nbase_max, int stage,
             uint32_t* __restrict__ cand, uint2* __restrict__ wcand, int stamp_tiles_axis) {
  extern __shared__ __align__(16) unsigned char dsm_fv[];
  find_valid_body(g, matches, base_idx, scan_start, scan_count, pool, pt_cell, cells, cell_count, gbox, work,
                  work_count, pmax, nbase_max, stage, (int)blockIdx.x, dsm_fv, 0, cand, wcand, stamp_tiles_axis);
}

// ---------------------------------------------------------------------------------------------
// K1a'  AddScan's "cell already occupied -> skip" rule for wide smears. When
// smear_deviation >= ~9.99 * resolution the stamp is 100 not only at its centre but also at its
// four edge neighbours, so a point whose cell an EARLIER point (or its neighbour) already set to
// 100 is skipped by Karto and never smeared: which stamps exist depends on the processing order.
// One warp per match replays that order over the compacted cell list: 32 cells per step, a
// shared-memory hash set of the stamped cells for the earlier steps, an in-order resolution
// inside the step. Skipped cells are overwritten with YSM_INVALID_CELL (k_tile_stamp ignores them).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool hash_has(const uint32_t* __restrict__ tab, uint32_t mask, int shift, uint32_t key) {
  uint32_t hsh = (key * 2654435761u) >> shift;
  while (true) {
    const uint32_t v = tab[hsh];
    if (v == key) return true;
    if (v == YSM_INVALID_CELL) return false;
    hsh = (hsh + 1) & mask;
  }
}

// (one warp; `lane` = lane id)
__device__ __forceinline__ void
stamp_order_body(const MatchDev* matches, uint32_t* cells, const int* cell_count, int log2cap, int vbx, int lane,
                 uint32_t* s_tab) {
  const MatchDev m = matches[vbx];
  const int n = cell_count[vbx];
  const uint32_t cap = 1u << log2cap, mask = cap - 1u;
  const int shift = 32 - log2cap;
  for (uint32_t i = lane; i < cap; i += 32) s_tab[i] = YSM_INVALID_CELL;
  __syncwarp();
  uint32_t* mc = cells + m.cells_off;
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    const bool valid = i < n;
    const uint32_t c = valid ? mc[i] : YSM_INVALID_CELL;
    const int x = (int)(c & 0xFFFFu), y = (int)(c >> 16);
    // already 100 because of a point stamped in an earlier step? (its cell or a 4-neighbour)
    bool blocked = !valid;
    if (valid) {
      blocked = hash_has(s_tab, mask, shift, c) || hash_has(s_tab, mask, shift, c - 1u) ||
                hash_has(s_tab, mask, shift, c + 1u) || hash_has(s_tab, mask, shift, c - 0x10000u) ||
                hash_has(s_tab, mask, shift, c + 0x10000u);
    }
    // in-order resolution inside the step
    bool stamped = false;
    for (int k = 0; k < 32; k++) {
      const bool mine = !__shfl_sync(0xffffffffu, (int)blocked, k);  // lane k stamps iff nothing earlier blocked it
      if (!mine) continue;                                            // warp-uniform
      const uint32_t ck = __shfl_sync(0xffffffffu, c, k);
      if (lane == k) stamped = true;
      const int dx = x - (int)(ck & 0xFFFFu), dy = y - (int)(ck >> 16);
      if (lane > k && (dx < 0 ? -dx : dx) + (dy < 0 ? -dy : dy) <= 1) blocked = true;
    }
    if (valid && !stamped) mc[i] = YSM_INVALID_CELL;
    // insert this step's stamped cells (distinct by construction)
    if (stamped) {
      uint32_t hsh = (c * 2654435761u) >> shift;
      while (atomicCAS(&s_tab[hsh], YSM_INVALID_CELL, c) != YSM_INVALID_CELL) hsh = (hsh + 1) & mask;
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(32)
k_stamp_order(const MatchDev* __restrict__ matches, uint32_t* __restrict__ cells,
              const int* __restrict__ cell_count, int log2cap) {
  extern __shared__ uint32_t dsm_so[];
  stamp_order_body(matches, cells, cell_count, log2cap, (int)blockIdx.x, (int)threadIdx.x, dsm_so);
}

// ---------------------------------------------------------------------------------------------
// K1b  CorrelationGrid::SmearPoint over every occupied cell (python twin
// yag_slam/helpers.py:105-119). The smear is a pure max of a K x K stamp, so each touched
// 32 x 32 tile of the grid is OWNED by one warp and lives in its REGISTERS: lane r holds tile
// row r as 16 u16x2 words. The warp collects the match's cells whose stamp reaches the tile
// (groups of 32 consecutive cells are pre-filtered by their bounding box; the per-candidate
// addressing is prepared lane-parallel while the list is compacted), then every candidate is one
// warp-uniform step: lane r reads the stamp row that crosses its tile row -- already shifted to
// the tile's columns, from a shared-memory table holding the stamp rows at the 8 cell alignments
// a 16-byte load allows -- and maxes it in with VIMNMX.U16x2, skipping the 8-cell groups the
// stamp does not overlap. No shared-memory read-modify-write, no atomics; the tile is written
// to the grid exactly once. Warps stride the wave's (match, tile) work list.
// Stamp table (host-built, ysm.cu build_stamp_table): u16 [8][K][Wt]; row j of alignment a holds
// the stamp's row j at cells [24 + a, 24 + a + K), zero elsewhere; Wt = 8 (mod 16) keeps the
// lanes' 16-byte loads (consecutive rows) on distinct banks.
// ---------------------------------------------------------------------------------------------
#define YSM_TILE_LIST 224   // candidate cells staged per warp between flushes

// candidate -> packed step descriptor: bits 31..12 q (signed; u16 index of tile row 0's window in
// the table), bit 11 zero, 10..7 mask of the 8-cell groups the stamp overlaps, 6..0 dy + 31 (dy = stamp
// row of tile row 0)
__device__ __forceinline__ uint32_t stamp_step(uint32_t c, int h, int K, int Wt, int x0t, int y0t) {
  const int ax = (int)(c & 0xFFFFu), ay = (int)(c >> 16);
  const int xr = ax - h - x0t;      // tile column of the stamp's first column: [-(K-1), 31]
  const int dy = y0t - (ay - h);    // stamp row that lands on tile row 0: [-31, K-1]
  const int a = xr & 7;
  const int ws = 24 + a - xr;       // multiple of 8
  const int q = (a * K + dy) * Wt + ws;
  const int g0 = max(0, xr) >> 3, g1 = min(31, xr + K - 1) >> 3;
  const uint32_t mask = ((2u << g1) - 1u) & ~((1u << g0) - 1u);
  return ((uint32_t)q << 12) | (mask << 7) | (uint32_t)(dy + 31);
}

// t[0..3] = max(t[0..3], the 8 cells at shared address saddr)   (VIMNMX.U16x2: two cells per instruction)
__device__ __forceinline__ void max_group(uint32_t& t0, uint32_t& t1, uint32_t& t2, uint32_t& t3, uint32_t saddr) {
  uint32_t a, b, c, d;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(saddr));
  t0 = __vmaxu2(t0, a);
  t1 = __vmaxu2(t1, b);
  t2 = __vmaxu2(t2, c);
  t3 = __vmaxu2(t3, d);
}

// lane_tab_s: shared-space address of this lane's row of the stamp table (tile row `lane`, stamp row 0)
__device__ __forceinline__ void tile_scatter_rows(uint32_t (&t)[16], const uint32_t* __restrict__ list, int n,
                                                  uint32_t lane_tab_s, int K) {
  const int lane31 = (int)(threadIdx.x & 31) - 31;
#pragma unroll 2
  for (int k = 0; k < n; k++) {
    const uint32_t w = list[k];  // warp-uniform (broadcast)
    const uint32_t saddr = lane_tab_s + (uint32_t)((int)w >> 11);  // + 2 * q bytes
    // group mask, cleared when this tile row does not cross the stamp
    const uint32_t gm = (unsigned)(lane31 + (int)(w & 0x7Fu)) < (unsigned)K ? w : 0u;
    if (gm & 0x080u) max_group(t[0], t[1], t[2], t[3], saddr);
    if (gm & 0x100u) max_group(t[4], t[5], t[6], t[7], saddr + 16u);
    if (gm & 0x200u) max_group(t[8], t[9], t[10], t[11], saddr + 32u);
    if (gm & 0x400u) max_group(t[12], t[13], t[14], t[15], saddr + 48u);
  }
}

__host__ __device__ __forceinline__ size_t stamp_table_bytes(int K, int Wt) { return (size_t)8 * K * Wt * 2; }

// dynamic smem: stamp table | per warp: byte tile staging [256] u32 | step list [YSM_TILE_LIST + 32] u32
__host__ __device__ __forceinline__ size_t tile_stamp_smem(int K, int Wt, int nwarps) {
  return stamp_table_bytes(K, Wt) + (size_t)nwarps * (YSM_TILE * YSM_TILE + (YSM_TILE_LIST + 32) * 4);
}

__device__ __forceinline__ void
tile_stamp_body(const GridC& g, const MatchDev* matches, const uint32_t* cells, const int* cell_count,
                const uint2* gbox, const int2* work, const int* work_count, const uint16_t* stamp_tab, uint8_t* grids,
                uint32_t* rowmask, int rm_words, int vbx, int vgx, unsigned char* dsm, int S = 1,
                const uint32_t* cand = nullptr, const uint2* wcand = nullptr) {
  // S > 1 (latency path): S warps share one tile -- each scans its share of the cell groups into a
  // private copy of the tile, the copies are max-combined at write-out (named barrier per group)
  __shared__ unsigned s_rows[16];
  if (threadIdx.x < 16) s_rows[threadIdx.x] = 0u;
  const int K = g.K, Wt = g.Wt, h = g.half_kernel;
  const int wpb = blockDim.x >> 5;  // warps per block
  const int ntab4 = (int)(stamp_table_bytes(K, Wt) / 16);
  uint4* s_tab4 = reinterpret_cast<uint4*>(dsm);
  for (int t = threadIdx.x; t < ntab4; t += blockDim.x) s_tab4[t] = __ldg(reinterpret_cast<const uint4*>(stamp_tab) + t);
  uint32_t* s_stage = reinterpret_cast<uint32_t*>(s_tab4 + ntab4);
  uint32_t* s_list = s_stage + (size_t)wpb * (YSM_TILE * YSM_TILE / 4);
  __syncthreads();
  const int tnx = (g.width + YSM_TILE - 1) / YSM_TILE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t lane_tab_s = (uint32_t)__cvta_generic_to_shared(dsm) + (uint32_t)(lane * Wt * 2);
  uint32_t* stage = s_stage + (size_t)warp * (YSM_TILE * YSM_TILE / 4);
  uint32_t* list = s_list + (size_t)warp * (YSM_TILE_LIST + 32);
  const int nwork = *work_count;
  const int gpc = wpb / S, group = warp / S, sub = warp - group * S;  // tile groups per CTA
  const int ngroups_total = vgx * gpc;
  for (int wi = vbx * gpc + group; wi < nwork; wi += ngroups_total) {
    const int2 wk = work[wi];
    const MatchDev m = matches[wk.x];
    const int ncells = cell_count[wk.x];
    const uint32_t* mc = cells + m.cells_off;
    const uint2* mb = gbox + m.gbox_off;
    const int ty = wk.y / tnx, tx = wk.y - ty * tnx;
    const int x0t = tx * YSM_TILE, y0t = ty * YSM_TILE;
    const int ngroups = (ncells + 31) >> 5;
    uint32_t t[16];
#pragma unroll
    for (int k = 0; k < 16; k++) t[k] = 0u;
    int n = 0;
    if (wcand) {
      // exact candidate list of this tile (k_find_valid): no search
      const uint2 wc = wcand[wi];
      const uint32_t* cl = cand + wc.x;
      const int cnt = (int)wc.y;
      for (int i0 = 0; i0 < cnt; i0 += 32) {
        if (i0 + lane < cnt) list[n + lane] = stamp_step(cl[i0 + lane], h, K, Wt, x0t, y0t);
        n += min(32, cnt - i0);
        if (n > YSM_TILE_LIST) {  // warp-uniform
          __syncwarp();
          tile_scatter_rows(t, list, n, lane_tab_s, K);
          __syncwarp();
          n = 0;
        }
      }
    } else
    for (int gs = 0; gs < ngroups; gs += 32 * S) {
      // groups of 32 cells whose bounding box (grown by the stamp) reaches the tile
      const int g0 = gs + sub * 32;
      bool ghit = false;
      if (g0 + lane < ngroups) {
        const uint2 bb = mb[g0 + lane];
        const int xlo = (int)(bb.x & 0xFFFFu), xhi = (int)(bb.x >> 16);
        const int ylo = (int)(bb.y & 0xFFFFu), yhi = (int)(bb.y >> 16);
        ghit = !(xhi + h < x0t || xlo - h > x0t + YSM_TILE - 1 || yhi + h < y0t || ylo - h > y0t + YSM_TILE - 1);
      }
      unsigned gm = __ballot_sync(0xffffffffu, ghit);
      while (gm) {
        const int gi = g0 + __ffs(gm) - 1;
        gm &= gm - 1;
        const int i = gi * 32 + lane;
        uint32_t c = 0;
        bool hit = false;
        if (i < ncells) {
          c = mc[i];
          const int ax = (int)(c & 0xFFFFu), ay = (int)(c >> 16);
          hit = c != YSM_INVALID_CELL && ax + h >= x0t && ax - h <= x0t + YSM_TILE - 1 && ay + h >= y0t &&
                ay - h <= y0t + YSM_TILE - 1;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (hit) list[n + __popc(bal & ((1u << lane) - 1u))] = stamp_step(c, h, K, Wt, x0t, y0t);
        n += __popc(bal);
        if (n > YSM_TILE_LIST) {  // warp-uniform
          __syncwarp();
          tile_scatter_rows(t, list, n, lane_tab_s, K);
          __syncwarp();
          n = 0;
        }
      }
    }
    __syncwarp();
    tile_scatter_rows(t, list, n, lane_tab_s, K);
    // lane r's row as bytes -> this warp's staging tile (16-byte chunks swizzled: conflict-free both ways)
    {
      uint32_t b[8];
#pragma unroll
      for (int k = 0; k < 8; k++) b[k] = __byte_perm(t[2 * k], t[2 * k + 1], 0x6420);  // u16 lanes -> bytes
      const int sw = (lane >> 2) & 1;
      uint4* st4 = reinterpret_cast<uint4*>(stage);
      st4[lane * 2 + (0 ^ sw)] = make_uint4(b[0], b[1], b[2], b[3]);
      st4[lane * 2 + (1 ^ sw)] = make_uint4(b[4], b[5], b[6], b[7]);
    }
    // the tile is written exactly once
    uint32_t* gout = reinterpret_cast<uint32_t*>(grids + (size_t)m.slot * g.grid_bytes);
    const int dr = lane >> 3, wd = lane & 7;
    const int gw = (x0t >> 2) + wd;
    uint32_t rows = 0;  // bit r: row r of this tile holds a non-zero cell (the sweep skips the others)
    if (S > 1) asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(32 * S) : "memory");
    else __syncwarp();
    const uint32_t* copies = s_stage + (size_t)(group * S) * (YSM_TILE * YSM_TILE / 4);
    for (int k = sub; k < 8; k += S) {
      const int r = k * 4 + dr, row = y0t + r;
      const int widx = r * 8 + 4 * ((wd >> 2) ^ (k & 1)) + (wd & 3);
      uint32_t v = copies[widx];
      for (int c = 1; c < S; c++) v = vmax4_lt128(v, copies[(size_t)c * (YSM_TILE * YSM_TILE / 4) + widx]);
      const unsigned nz = __ballot_sync(0xffffffffu, v != 0u);
      // all-zero rows are not written (the slot is all-zero between matches; 8 lanes = one 32-byte sector)
      if ((nz & (0xFFu << (8 * dr))) && row < g.height && gw < g.stride4) gout[(size_t)row * g.stride4 + gw] = v;
#pragma unroll
      for (int d = 0; d < 4; d++)
        if (nz & (0xFFu << (8 * d))) rows |= 1u << (k * 4 + d);
    }
    if (S > 1) {
      if (lane == 0 && rows) atomicOr(&s_rows[group], rows);
      asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(32 * S) : "memory");
      if (sub == 0 && lane == 0) {
        rowmask[(size_t)m.slot * rm_words + wk.y] = s_rows[group];
        s_rows[group] = 0u;
      }
    } else if (lane == 0) {
      rowmask[(size_t)m.slot * rm_words + wk.y] = rows;
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256)
k_tile_stamp(GridC g, const MatchDev* __restrict__ matches, const uint32_t* __restrict__ cells,
             const int* __restrict__ cell_count, const uint2* __restrict__ gbox,
             const int2* __restrict__ work, const int* __restrict__ work_count,
             const uint16_t* __restrict__ stamp_tab, uint8_t* __restrict__ grids, uint32_t* __restrict__ rowmask,
             int rm_words, const uint32_t* __restrict__ cand, const uint2* __restrict__ wcand) {
  extern __shared__ __align__(16) unsigned char dsm_ts[];
  tile_stamp_body(g, matches, cells, cell_count, gbox, work, work_count, stamp_tab, grids, rowmask, rm_words,
                  (int)blockIdx.x, (int)gridDim.x, dsm_ts, 1, cand, wcand);
}

// ---------------------------------------------------------------------------------------------
// K1b (throughput form)  k_tile_stamp_lists: the same register-tile stamping for tiles whose exact
// candidate lists k_find_valid made, with the step loop rebuilt around what ncu showed for k_tile_stamp
// (r02p: half of all instructions were the per-candidate group tests / reconvergence, and only 12 of 32
// lanes were live in the max instructions because a K-row stamp crosses ~13 of a tile's 32 rows):
//   * a candidate is expanded into one step per 8-column group it overlaps, kept in a list of that group,
//     so a group's loop names its four tile registers statically and carries no group test;
//   * the tile is treated as two 16-row halves with their own lists: lanes 0-15 walk the upper half's
//     list while lanes 16-31 walk the lower half's, so a stamp that only reaches one half costs the other
//     half nothing (steps per group = max of the two lengths instead of their sum);
//   * a step is one word: bits 31..16 = byte offset of tile row 0's window in the stamp table (biased to be
//     non-negative), bits 15..0 = which of the half's 16 rows the stamp crosses; the lane's row bit
//     selects between the stamp's window and 16 zero bytes for one LDS.128 + four VIMNMX.U16x2, no branch.
// Results are identical to k_tile_stamp (a pure max; tests/test_gpu_parity.py compares grid bytes).
// ---------------------------------------------------------------------------------------------
#ifndef YSM_STAMP_DEDUP
#define YSM_STAMP_DEDUP 1
#endif
#define YSM_HL_CAP 96  // steps per (column group, tile half) list between flushes; multiple of 4

__host__ __device__ __forceinline__ int stamp_lists_bias(int Wt) { return 31 * Wt + 8; }  // cells; multiple of 8
// does every step offset fit 16 bits?
__host__ __device__ __forceinline__ bool stamp_lists_fit(int K, int Wt) {
  return 2 * ((8 * K - 1) * Wt + 56 + stamp_lists_bias(Wt)) <= 65535;
}
__host__ __device__ __forceinline__ size_t tile_stamp_lists_smem(int K, int Wt, int nwarps) {
  return stamp_table_bytes(K, Wt) + (size_t)nwarps * (8 * YSM_HL_CAP * 4);
}

// t0..t3 = max(t0..t3, the 8 cells of this lane's row of step w); a lane whose row the stamp does not cross
// reads 16 zero bytes (zaddr: the left margin of the table's first row) instead of branching around the
// load -- ptxas then pairs two steps into VIMNMX3.U16x2 (7.25 instructions per step)
__device__ __forceinline__ void half_step(uint32_t& t0, uint32_t& t1, uint32_t& t2, uint32_t& t3, uint32_t w,
                                          uint32_t tab_s, uint32_t lanebit, uint32_t zaddr) {
  const uint32_t saddr = (w & lanebit) ? tab_s + (w >> 16) : zaddr;
  uint32_t a, b, c, d;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(saddr));
  t0 = __vmaxu2(t0, a);
  t1 = __vmaxu2(t1, b);
  t2 = __vmaxu2(t2, c);
  t3 = __vmaxu2(t3, d);
}

__device__ __forceinline__ void half_group_run(uint32_t& t0, uint32_t& t1, uint32_t& t2, uint32_t& t3,
                                               const uint32_t* __restrict__ lp, int n4, uint32_t tab_s, uint32_t lanebit,
                                               uint32_t zaddr) {
#pragma unroll 1
  for (int k = 0; k < n4; k += 4) {
    const uint4 ww = *reinterpret_cast<const uint4*>(lp + k);
    half_step(t0, t1, t2, t3, ww.x, tab_s, lanebit, zaddr);
    half_step(t0, t1, t2, t3, ww.y, tab_s, lanebit, zaddr);
    half_step(t0, t1, t2, t3, ww.z, tab_s, lanebit, zaddr);
    half_step(t0, t1, t2, t3, ww.w, tab_s, lanebit, zaddr);
  }
}

// run and empty the warp's eight lists (nU / nL: four byte counters each, one per column group)
__device__ __forceinline__ void half_lists_flush(uint32_t (&t)[16], uint32_t* wl, uint32_t& nU, uint32_t& nL,
                                                 uint32_t tab_s, uint32_t lanebit, uint32_t zaddr, int lane) {
  // pad the shorter list of every group with null steps up to the common length (multiple of 4)
#pragma unroll
  for (int gq = 0; gq < 4; gq++) {
    const int a = (int)((nU >> (8 * gq)) & 0xFFu), b = (int)((nL >> (8 * gq)) & 0xFFu);
    const int n4 = (max(a, b) + 3) & ~3;
    for (int k = a + lane; k < n4; k += 32) wl[(2 * gq) * YSM_HL_CAP + k] = 0u;
    for (int k = b + lane; k < n4; k += 32) wl[(2 * gq + 1) * YSM_HL_CAP + k] = 0u;
  }
  __syncwarp();
  const uint32_t* lp = wl + (lane >> 4) * YSM_HL_CAP;
#pragma unroll
  for (int gq = 0; gq < 4; gq++) {
    const int a = (int)((nU >> (8 * gq)) & 0xFFu), b = (int)((nL >> (8 * gq)) & 0xFFu);
    const int n4 = (max(a, b) + 3) & ~3;
    half_group_run(t[4 * gq], t[4 * gq + 1], t[4 * gq + 2], t[4 * gq + 3], lp + (2 * gq) * YSM_HL_CAP, n4,
                   tab_s + 16u * gq, lanebit, zaddr);
  }
  __syncwarp();
  nU = 0u;
  nL = 0u;
}

__global__ void __launch_bounds__(256, 4)
k_tile_stamp_lists(GridC g, const MatchDev* __restrict__ matches, const int2* __restrict__ work,
                   const int* __restrict__ work_count, const uint16_t* __restrict__ stamp_tab,
                   uint8_t* __restrict__ grids, uint32_t* __restrict__ rowmask, int rm_words,
                   const uint32_t* __restrict__ cand, const uint2* __restrict__ wcand) {
  extern __shared__ __align__(16) unsigned char dsm_tl[];
  const int K = g.K, Wt = g.Wt, h = g.half_kernel;
  const int ntab4 = (int)(stamp_table_bytes(K, Wt) / 16);
  uint4* s_tab4 = reinterpret_cast<uint4*>(dsm_tl);
  for (int t = threadIdx.x; t < ntab4; t += blockDim.x) s_tab4[t] = __ldg(reinterpret_cast<const uint4*>(stamp_tab) + t);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  uint32_t* wl = reinterpret_cast<uint32_t*>(s_tab4 + ntab4) + (size_t)warp * (8 * YSM_HL_CAP);  // lists; staging tile at the end
  const int bias = stamp_lists_bias(Wt);
  const uint32_t tab_s = (uint32_t)__cvta_generic_to_shared(dsm_tl) + (uint32_t)(lane * Wt * 2) - (uint32_t)(2 * bias);
  const uint32_t lanebit = 1u << (lane & 15);
  const uint32_t zaddr = (uint32_t)__cvta_generic_to_shared(dsm_tl);  // cells 0..7 of table row 0: zeros
  const uint32_t ltmask = (1u << lane) - 1u;
  const int tnx = (g.width + YSM_TILE - 1) / YSM_TILE;
  const int nwork = *work_count;
  const int nwarps_total = (int)gridDim.x * wpb;
  int wi = (int)blockIdx.x * wpb + warp;
  int2 wk = make_int2(0, 0);
  uint2 wc = make_uint2(0u, 0u);
  uint32_t c0 = 0u;
  int slot = 0;
  if (wi < nwork) {
    wk = work[wi];
    wc = wcand[wi];
    slot = matches[wk.x].slot;
    if (lane < (int)wc.y) c0 = cand[wc.x + lane];
  }
  for (; wi < nwork; wi += nwarps_total) {
    // the next tile's descriptors are fetched under this tile's work
    const int nxt = wi + nwarps_total;
    int2 wk_n = make_int2(0, 0);
    uint2 wc_n = make_uint2(0u, 0u);
    if (nxt < nwork) {
      wk_n = work[nxt];
      wc_n = wcand[nxt];
    }
    const int ty = wk.y / tnx, tx = wk.y - ty * tnx;
    const int x0t = tx * YSM_TILE, y0t = ty * YSM_TILE;
    const uint32_t* cl = cand + wc.x;
    const int cnt = (int)wc.y;
    uint32_t t[16];
#pragma unroll
    for (int k = 0; k < 16; k++) t[k] = 0u;
    uint32_t nU = 0u, nL = 0u;
    const uint32_t dummy = (uint32_t)x0t | ((uint32_t)y0t << 16);  // lanes past the list: in-range arithmetic, no steps
    uint32_t c = c0;
#pragma unroll 1
    for (int i0 = 0; i0 < cnt; i0 += 32) {
      if (i0 && i0 + lane < cnt) c = cl[i0 + lane];
      if (i0 + lane >= cnt) c = dummy;
      // candidate -> steps
      const int ax = (int)(c & 0xFFFFu), ay = (int)(c >> 16);
      const int xr = ax - h - x0t;    // tile column of the stamp's first column: [-(K-1), 31]
      const int dy = y0t - (ay - h);  // stamp row that lands on tile row 0: [-31, K-1]
      const int a = xr & 7;
      const int q = (a * K + dy) * Wt + 24 + a - xr + bias;  // > 0, multiple of 8
      const int g0 = max(0, xr) >> 3, g1 = min(31, xr + K - 1) >> 3;
      uint32_t gmask = ((2u << g1) - 1u) & ~((1u << g0) - 1u);
      if (i0 + lane >= cnt) gmask = 0u;
#if YSM_STAMP_DEDUP
      // base points of different running scans often fall into the same cell (13 % on the bench workload): of the
      // lanes of this chunk naming one cell only the first makes steps (the smear is a max: a repeat adds nothing)
      if (lane != __ffs(__match_any_sync(0xffffffffu, c)) - 1) gmask = 0u;
#endif
      const int rlo = max(0, -dy), rhi = min(31, K - 1 - dy);
      const uint32_t rows = (0xFFFFFFFFu >> (31 - rhi)) & (0xFFFFFFFFu << rlo);
      const uint32_t wU = ((uint32_t)(2 * q) << 16) | (rows & 0xFFFFu), wL = ((uint32_t)(2 * q) << 16) | (rows >> 16);
      const bool up = (rows & 0xFFFFu) != 0u, lo = (rows >> 16) != 0u;
#pragma unroll
      for (int gq = 0; gq < 4; gq++) {
        const bool ing = (gmask >> gq) & 1u;
        const unsigned bu = __ballot_sync(0xffffffffu, ing && up), bl = __ballot_sync(0xffffffffu, ing && lo);
        if (ing && up) wl[(2 * gq) * YSM_HL_CAP + ((nU >> (8 * gq)) & 0xFFu) + __popc(bu & ltmask)] = wU;
        if (ing && lo) wl[(2 * gq + 1) * YSM_HL_CAP + ((nL >> (8 * gq)) & 0xFFu) + __popc(bl & ltmask)] = wL;
        nU += (uint32_t)__popc(bu) << (8 * gq);
        nL += (uint32_t)__popc(bl) << (8 * gq);
      }
      // another chunk could overflow a list (a counter above CAP - 32): run what is there
      if (i0 + 32 < cnt && (((nU + 0x3F3F3F3Fu) | (nL + 0x3F3F3F3Fu)) & 0x80808080u))
        half_lists_flush(t, wl, nU, nL, tab_s, lanebit, zaddr, lane);
    }
    int slot_n = 0;
    if (nxt < nwork) {
      slot_n = matches[wk_n.x].slot;
      if (lane < (int)wc_n.y) c0 = cand[wc_n.x + lane];
    }
    half_lists_flush(t, wl, nU, nL, tab_s, lanebit, zaddr, lane);
    // lane r's row as bytes -> this warp's staging tile (16-byte chunks swizzled: conflict-free both ways)
    uint32_t rows_nz;  // bit r: row r of this tile holds a non-zero cell (the sweep skips the others)
    {
      uint32_t b[8];
#pragma unroll
      for (int k = 0; k < 8; k++) b[k] = __byte_perm(t[2 * k], t[2 * k + 1], 0x6420);  // u16 lanes -> bytes
      rows_nz = __ballot_sync(0xffffffffu, ((b[0] | b[1]) | (b[2] | b[3]) | (b[4] | b[5]) | (b[6] | b[7])) != 0u);
      const int sw = (lane >> 2) & 1;
      uint4* st4 = reinterpret_cast<uint4*>(wl);
      st4[lane * 2 + (0 ^ sw)] = make_uint4(b[0], b[1], b[2], b[3]);
      st4[lane * 2 + (1 ^ sw)] = make_uint4(b[4], b[5], b[6], b[7]);
    }
    __syncwarp();
    // the tile is written exactly once: four lanes per row (8 bytes each: the row stride is a multiple of 8), eight
    // rows per store instruction; all-zero rows are not written (the slot is all-zero between matches)
    {
      uint2* gout = reinterpret_cast<uint2*>(grids + (size_t)slot * g.grid_bytes);
      const int stride8 = g.stride4 >> 1;
      const int dr = lane >> 2, wd = lane & 3;
      const int gw = (x0t >> 3) + wd;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int r = k * 8 + dr, row = y0t + r;
        if (((rows_nz >> r) & 1u) && row < g.height && gw < stride8)
          gout[(size_t)row * stride8 + gw] =
              *reinterpret_cast<const uint2*>(wl + r * 8 + 4 * ((wd >> 1) ^ ((r >> 2) & 1)) + 2 * (wd & 1));
      }
    }
    if (lane == 0) rowmask[(size_t)slot * rm_words + wk.y] = rows_nz;
    __syncwarp();
    wk = wk_n;
    wc = wc_n;
    slot = slot_n;
  }
}

// zero the tiles a wave touched (the slot grids are kept all-zero between matches). A warp takes 32
// work items at a time: the (match, tile) pairs and their slots are fetched lane-parallel (no
// dependent-load chain per tile), then each tile is zeroed with four 8-byte stores per lane
// (row stride is a multiple of 8 bytes).
#define YSM_CLEAR_BATCH 8  // work items a warp takes per round (r02z: with 32, half the launched warps had no work)
__global__ void __launch_bounds__(256)
k_tile_clear(GridC g, const MatchDev* __restrict__ matches, const int2* __restrict__ work,
             const int* __restrict__ work_count, uint8_t* __restrict__ grids, uint32_t* __restrict__ rowmask,
             int rm_words) {
  const int tnx = (g.width + YSM_TILE - 1) / YSM_TILE;
  const int lane = threadIdx.x & 31, dr = lane >> 2, wd = lane & 3;
  const int nwork = *work_count;
  const int nwarps_total = gridDim.x * (blockDim.x >> 5);
  const int stride8 = g.stride4 >> 1;
  for (int base = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * YSM_CLEAR_BATCH; base < nwork; base += nwarps_total * YSM_CLEAR_BATCH) {
    const int cnt = min(YSM_CLEAR_BATCH, nwork - base);
    int my_tile = 0, my_slot = 0;
    uint32_t my_rows = 0u;  // rows of the tile that hold a non-zero cell: the only ones the stamp kernel wrote
    if (lane < cnt) {
      const int2 wk = work[base + lane];
      my_tile = wk.y;
      my_slot = matches[wk.x].slot;
      uint32_t* rm = rowmask + (size_t)my_slot * rm_words + wk.y;
      my_rows = *rm;
      if (my_rows) *rm = 0u;
    }
    for (int j = 0; j < cnt; j++) {
      const uint32_t rows = __shfl_sync(0xffffffffu, my_rows, j);
      if (rows == 0u) continue;  // warp-uniform
      const int tile = __shfl_sync(0xffffffffu, my_tile, j), slot = __shfl_sync(0xffffffffu, my_slot, j);
      const int ty = tile / tnx, tx = tile - ty * tnx;
      uint2* gout = reinterpret_cast<uint2*>(grids + (size_t)slot * g.grid_bytes);
      const int gw = ((tx * YSM_TILE) >> 3) + wd;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int row = ty * YSM_TILE + k * 8 + dr;
        if (((rows >> (k * 8 + dr)) & 1u) && row < g.height && gw < stride8) gout[(size_t)row * stride8 + gw] = make_uint2(0u, 0u);
      }
    }
  }
}

// cell offset (gx, gy) of query point (wx, wy) rotated by the search angle (cosine, sine):
// the arithmetic of GridIndexLookup::ComputeOffsets, shared by k_offsets and the fused sweep
__device__ __forceinline__ void offset_cell(const TableDev& t, double scale, double wx, double wy, double cosine,
                                            double sine, int& gx, int& gy) {
  const double dx = wx - t.px, dy = wy - t.py;
  const double lx = t.r00 * dx + t.r01 * dy;
  const double ly = t.r10 * dx + t.r11 * dy;
  const double ox = cosine * lx - sine * ly;
  const double oy = sine * lx + cosine * ly;
  gx = world_to_grid1(ox + t.gox, t.gox, scale);
  gy = world_to_grid1(oy + t.goy, t.goy, scale);
}

// (thread `tid0` of `nthreads` cooperating on table t)
__device__ __forceinline__ void offsets_body(const GridC& g, const TableDev& t, const double* trig, const double* pool,
                                             int* offsets, int tid0, int nthreads) {
  const int p4n = t.Ppad >> 2;
  const int work = t.nA * p4n;
  for (int it = tid0; it < work; it += nthreads) {
    const int a = it / p4n, p0 = (it - a * p4n) << 2;
    const double cosine = trig[2 * (t.trig_off + a)], sine = trig[2 * (t.trig_off + a) + 1];
    int o[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int p = p0 + k;
      int v = 0;
      if (p < t.P) {
        const double wx = pool[2 * (size_t)(t.q_start + p)], wy = pool[2 * (size_t)(t.q_start + p) + 1];
        int gx, gy;
        offset_cell(t, g.scale, wx, wy, cosine, sine, gx, gy);
        v = gx + gy * g.stride;
      }
      o[k] = v;
    }
    *reinterpret_cast<int4*>(offsets + t.out_off + (size_t)a * t.Ppad + p0) = make_int4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void __launch_bounds__(256)
k_offsets(GridC g, const TableDev* __restrict__ tables, const double* __restrict__ trig,
          const double* __restrict__ pool, int* __restrict__ offsets, int table_base) {
  const TableDev t = tables[table_base + blockIdx.y];
  offsets_body(g, t, trig, pool, offsets, (int)(blockIdx.x * blockDim.x + threadIdx.x), (int)(gridDim.x * blockDim.x));
}

// ---------------------------------------------------------------------------------------------
// K3  CorrelateScan / GetResponse sweep, lattice form (python twin
// find_best_pose, yag_slam/helpers.py:156-295). CTA = one (pass, angle); the angle's lookup
// offsets are staged in shared memory; ONE WARP PER LATTICE ROW (lanes = adjacent x poses, so a
// warp load touches one contiguous span of a grid row) and all rows of the lattice walk the
// query points together, so the rows' overlapping windows are served by L1 and only a dozen
// matches are in flight chip-wide (their grid lines stay L2-resident). Each lane accumulates
// its pose's response as an exact integer. grid = (n_pass_angles, task_chunks, p_chunks).
// ---------------------------------------------------------------------------------------------
struct PassAngle {
  int pass, a;
  int table;   // = passes[pass].table      (copies: a sweep CTA fetches its pass, its table and its angle's
  int trig_i;  // = tables[table].trig_off + a   cos/sin in ONE round of loads instead of a chain of three)
};

// GetResponse normalisation + the odometry penalty of CorrelateScan for the
// pose (ix, iy, a) whose integer lookup sum is `sum`.
// the two halves of the odometry penalty (dp: distance, ap: angle); the same expressions, in the same order,
// as CorrelateScan (A.7)
__device__ __forceinline__ double penalty_distance(const PassDev& ps, const PenaltyC& pen, int ix, int iy) {
  const double x = -ps.offx + (double)ix * ps.resx;
  const double y = -ps.offy + (double)iy * ps.resy;
  const double sqd = x * x + y * y;
  const double dp = 1.0 - (0.2 * sqd / pen.distance_variance_penalty);
  return dp > pen.minimum_distance_penalty ? dp : pen.minimum_distance_penalty;
}
__device__ __forceinline__ double penalty_angle(const PassDev& ps, const PenaltyC& pen, int a) {
  const double angle = (ps.ch - ps.angle_offset) + (double)a * ps.angle_res;
  const double da = angle - ps.ch;
  const double sqa = da * da;
  const double ap = 1.0 - (0.2 * sqa / pen.angle_variance_penalty);
  return ap > pen.minimum_angle_penalty ? ap : pen.minimum_angle_penalty;
}
// response of a pose from its integer lookup sum and the penalty product dp * ap computed beforehand
__device__ __forceinline__ double response_from(const PassDev& ps, unsigned sum, double dpap) {
  double r = (double)sum / (double)((unsigned)ps.P * 100u);
  if (ps.penalize && !kt_double_equal(r, 0.0)) r *= dpap;
  return r;
}

__device__ __forceinline__ double response_of(const PassDev& ps, const PenaltyC& pen, unsigned sum, int ix,
                                              int iy, int a) {
  double r = (double)sum / (double)((unsigned)ps.P * 100u);
  if (ps.penalize && !kt_double_equal(r, 0.0)) r *= (penalty_distance(ps, pen, ix, iy) * penalty_angle(ps, pen, a));
  return r;
}

// m_pSearchSpaceProbs of CorrelateScan: per lattice cell the maximum response over the
// angles, accumulated by the sweep epilogues (responses are >= 0: bit patterns order like integers)
__device__ __forceinline__ void cell_max_update(unsigned long long* cellmax, const PassDev& ps, int ix, int iy, double v) {
  if (ps.cmax_off >= 0)
    atomicMax(cellmax + ps.cmax_off + (size_t)iy * ps.nX + ix, (unsigned long long)__double_as_longlong(v));
}

// responses are >= 0, so their IEEE bit patterns order like unsigned integers
__device__ __forceinline__ void pass_max_update(double* passmax, int pass, double v) {
  atomicMax(reinterpret_cast<unsigned long long*>(passmax + pass), (unsigned long long)__double_as_longlong(v));
}


// s_off holds offsets biased by the CTA-wide minimum (so they are non-negative and extend to
// 64 bits for free); gp / base already include that minimum.
template <bool kChecked>
__device__ __forceinline__ unsigned sweep_row(const uint8_t* __restrict__ gp, const unsigned* __restrict__ s_off,
                                              int np, int base, unsigned dsz) {
  unsigned sum0 = 0, sum1 = 0;
  int p = 0;
  for (; p + 8 <= np; p += 8) {
    const uint4 o0 = *reinterpret_cast<const uint4*>(s_off + p);
    const uint4 o1 = *reinterpret_cast<const uint4*>(s_off + p + 4);
    unsigned v0, v1, v2, v3, v4, v5, v6, v7;
    if (kChecked) {
      v0 = ((unsigned)(base + o0.x) < dsz) ? (unsigned)__ldg(gp + o0.x) : 0u;
      v1 = ((unsigned)(base + o0.y) < dsz) ? (unsigned)__ldg(gp + o0.y) : 0u;
      v2 = ((unsigned)(base + o0.z) < dsz) ? (unsigned)__ldg(gp + o0.z) : 0u;
      v3 = ((unsigned)(base + o0.w) < dsz) ? (unsigned)__ldg(gp + o0.w) : 0u;
      v4 = ((unsigned)(base + o1.x) < dsz) ? (unsigned)__ldg(gp + o1.x) : 0u;
      v5 = ((unsigned)(base + o1.y) < dsz) ? (unsigned)__ldg(gp + o1.y) : 0u;
      v6 = ((unsigned)(base + o1.z) < dsz) ? (unsigned)__ldg(gp + o1.z) : 0u;
      v7 = ((unsigned)(base + o1.w) < dsz) ? (unsigned)__ldg(gp + o1.w) : 0u;
    } else {
      v0 = __ldg(gp + o0.x); v1 = __ldg(gp + o0.y); v2 = __ldg(gp + o0.z); v3 = __ldg(gp + o0.w);
      v4 = __ldg(gp + o1.x); v5 = __ldg(gp + o1.y); v6 = __ldg(gp + o1.z); v7 = __ldg(gp + o1.w);
    }
    sum0 += v0 + v1 + v2 + v3;
    sum1 += v4 + v5 + v6 + v7;
  }
  for (; p < np; p++) {
    const unsigned o = s_off[p];
    if (!kChecked || (unsigned)(base + o) < dsz) sum0 += (unsigned)__ldg(gp + o);
  }
  return sum0 + sum1;
}

// offsets == nullptr: the lookup offsets of the angle are computed here from the query points
// (ComputeOffsets fused; trig / pool must be given), else read from the k_offsets table.
__device__ __forceinline__ void
sweep_lattice_body(const GridC& g, const PenaltyC& pen, const PassDev* passes, const PassAngle* pa_list,
                   const TableDev* tables, const int* offsets, const double* trig, const double* pool,
                   const uint8_t* grids, double* resp, double* passmax, unsigned long long* cellmax,
                   int tasks_per_cta, int psplit, int vbx, int vby, int* s_i) {
  __shared__ int s_minmax[4];  // min off, max off, min base, max base
  __shared__ double s_wmax[32];
  const int p_chunk = (passes[pa_list[vbx].pass].P + 7) & ~7;
  const PassAngle pa = pa_list[vbx];
  const PassDev ps = passes[pa.pass];
  const int nxc = (ps.nX + 31) >> 5;
  const int ntasks = ps.nY * nxc;
  const int task0 = vby * tasks_per_cta;
  if (task0 >= ntasks) return;
  const int task1 = min(ntasks, task0 + tasks_per_cta);
  const int pbeg = 0;
  const int pend = ps.P;
  const int pc4 = (p_chunk + 3) & ~3;
  int* s_off = s_i;             // [pc4]
  int* s_col = s_i + pc4;       // [nX]
  int* s_row = s_col + ps.nX;   // [nY]
  if (threadIdx.x == 0) {
    s_minmax[0] = 0x7fffffff; s_minmax[1] = (int)0x80000000;
    s_minmax[2] = 0x7fffffff; s_minmax[3] = (int)0x80000000;
  }
  __syncthreads();
  const TableDev tb = tables[ps.table];
  const int* goff = offsets ? offsets + tb.out_off + (size_t)pa.a * tb.Ppad : nullptr;
  double cosine = 0.0, sine = 0.0;
  if (!goff) {
    cosine = trig[2 * (tb.trig_off + pa.a)];
    sine = trig[2 * (tb.trig_off + pa.a) + 1];
  }
  int mn = 0x7fffffff, mx = (int)0x80000000;
  for (int p = pbeg + threadIdx.x; p < pend; p += blockDim.x) {
    int o;
    if (goff) {
      o = goff[p];
    } else {
      const double2 w = *reinterpret_cast<const double2*>(pool + 2 * (size_t)(tb.q_start + p));
      int gx, gy;
      offset_cell(tb, g.scale, w.x, w.y, cosine, sine, gx, gy);
      o = gx + gy * g.stride;
    }
    s_off[p - pbeg] = o;
    mn = min(mn, o);
    mx = max(mx, o);
  }
  const double startX = -ps.offx, startY = -ps.offy;
  int bmn = 0x7fffffff, bmx = (int)0x80000000;
  for (int i = threadIdx.x; i < ps.nX; i += blockDim.x) {
    const double x = startX + (double)i * ps.resx;
    const int c = world_to_grid1(ps.cx + x, ps.gox, g.scale) + g.border;
    s_col[i] = c;
  }
  for (int i = threadIdx.x; i < ps.nY; i += blockDim.x) {
    const double y = startY + (double)i * ps.resy;
    const int r = (world_to_grid1(ps.cy + y, ps.goy, g.scale) + g.border) * g.stride;
    s_row[i] = r;
    bmn = min(bmn, r);
    bmx = max(bmx, r);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    bmn = min(bmn, __shfl_xor_sync(0xffffffffu, bmn, o));
    bmx = max(bmx, __shfl_xor_sync(0xffffffffu, bmx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&s_minmax[0], mn);
    atomicMax(&s_minmax[1], mx);
    if (bmn != 0x7fffffff) { atomicMin(&s_minmax[2], bmn); atomicMax(&s_minmax[3], bmx); }
  }
  __syncthreads();
  // column extents from shared memory (nX is small); rows + columns + offsets give a conservative bound
  int colmin = 0x7fffffff, colmax = (int)0x80000000;
  for (int i = (int)(threadIdx.x & 31); i < ps.nX; i += 32) {
    colmin = min(colmin, s_col[i]);
    colmax = max(colmax, s_col[i]);
  }
  for (int o = 16; o > 0; o >>= 1) {
    colmin = min(colmin, __shfl_xor_sync(0xffffffffu, colmin, o));
    colmax = max(colmax, __shfl_xor_sync(0xffffffffu, colmax, o));
  }
  const int minoff = s_minmax[0];
  for (int p = threadIdx.x; p < pend - pbeg; p += blockDim.x) s_off[p] -= minoff;  // own entries only
  __syncthreads();
  const long long lo = (long long)s_minmax[0] + s_minmax[2] + colmin;
  const long long hi = (long long)s_minmax[1] + s_minmax[3] + colmax;
  const bool safe = lo >= 0 && hi < (long long)g.data_size;  // CTA-uniform: no lookup can leave the grid

  const uint8_t* grid = grids + (size_t)ps.slot * g.grid_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int np = pend - pbeg;
  const unsigned dsz = (unsigned)g.data_size;
  double wmax = 0.0;
  // psplit > 1 (small batches): `psplit` warps share one row-task, each sums a slice of the
  // points; the slices are combined through shared memory (exact: integer sums)
  const int chunk = warp % psplit, wtask = warp / psplit, ntw = nwarps / psplit;
  const int plen = (((np + psplit - 1) / psplit) + 7) & ~7;
  const int pb = min(np, chunk * plen), pe = min(np, pb + plen);
  unsigned* s_part = reinterpret_cast<unsigned*>(s_row + ps.nY);  // [nwarps][32] when psplit > 1
  for (int task = task0 + wtask; task < task1; task += ntw) {
    const int iy = task / nxc, xc = task - iy * nxc;
    const int ix = (xc << 5) + lane;
    const bool active = ix < ps.nX;
    const int base = s_row[iy] + s_col[active ? ix : 0] + minoff;
    const uint8_t* gp = grid + base;
    const unsigned* uoff = reinterpret_cast<const unsigned*>(s_off) + pb;
    unsigned sum = safe ? sweep_row<false>(gp, uoff, pe - pb, base, dsz) : sweep_row<true>(gp, uoff, pe - pb, base, dsz);
    if (psplit > 1) {
      // all warps of the CTA run the same number of task iterations (host sizes tasks_per_cta == ntw)
      s_part[warp * 32 + lane] = sum;
      __syncthreads();
      if (chunk == 0)
        for (int c = 1; c < psplit; c++) sum += s_part[(warp + c) * 32 + lane];
      __syncthreads();
      if (chunk != 0) continue;
    }
    if (active) {
      const double rr = response_of(ps, pen, sum, ix, iy, pa.a);
      resp[ps.sums_off + ((size_t)iy * ps.nX + ix) * ps.nA + pa.a] = rr;
      cell_max_update(cellmax, ps, ix, iy, rr);
      wmax = rr > wmax ? rr : wmax;
    }
  }
  // best response of this CTA -> the pass maximum (CorrelateScan's bestResponse)
  for (int o = 16; o > 0; o >>= 1) {
    const double t = __shfl_xor_sync(0xffffffffu, wmax, o);
    wmax = t > wmax ? t : wmax;
  }
  if (lane == 0) s_wmax[warp] = wmax;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = s_wmax[0];
    for (int w = 1; w < nwarps; w++) m = s_wmax[w] > m ? s_wmax[w] : m;
    pass_max_update(passmax, pa.pass, m);
  }
}

__global__ void __launch_bounds__(1024, 2)
k_sweep_lattice(GridC g, PenaltyC pen, const PassDev* __restrict__ passes,
                const PassAngle* __restrict__ pa_list, const TableDev* __restrict__ tables,
                const int* __restrict__ offsets, const uint8_t* __restrict__ grids,
                double* __restrict__ resp, double* __restrict__ passmax, unsigned long long* __restrict__ cellmax,
                int tasks_per_cta, int psplit) {
  extern __shared__ __align__(16) int dsm_sl[];
  sweep_lattice_body(g, pen, passes, pa_list, tables, offsets, nullptr, nullptr, grids, resp, passmax, cellmax,
                     tasks_per_cta, psplit, (int)blockIdx.x, (int)blockIdx.y, dsm_sl);
}

// ---------------------------------------------------------------------------------------------
// K3p  CorrelateScan / GetResponse sweep with exact zero-row pruning (throughput form).
// A lookup only contributes when its cell is non-zero, and most of the (angle, point, lattice
// row) spans a sweep touches lie in free space. k_tile_stamp leaves one 32-bit word per 32 x 32
// grid tile saying which tile rows hold a non-zero cell; this kernel
//   A. computes the angle's lookup offsets itself (ComputeOffsets fused: no table round trip),
//      and for every query point ORs the <= 3 x 3 tile words under the point's window into a
//      per-point mask "lattice row r can see a non-zero cell";
//   B. lets every warp (= one lattice row, lanes = adjacent x poses as in k_sweep_lattice)
//      compact the offsets of the points whose bit is set into its private shared-memory list;
//   C. sums only those (identical result: the skipped cells are all zero).
// CTA = (pass, angle) x (group of <= 28 lattice rows, chunk of <= 32 lattice columns).
// Windows that could leave the grid (or wrap a row) disable pruning for that batch of points and
// take Karto's flat bounds check instead.
// smem: s_pm[PB] {offset, row mask} of the surviving points | lists[nrows][PB]   (PB = points per batch)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t even_bits64(unsigned long long x) {
  x &= 0x5555555555555555ull;
  x = (x | (x >> 1)) & 0x3333333333333333ull;
  x = (x | (x >> 2)) & 0x0F0F0F0F0F0F0F0Full;
  x = (x | (x >> 4)) & 0x00FF00FF00FF00FFull;
  x = (x | (x >> 8)) & 0x0000FFFF0000FFFFull;
  x = (x | (x >> 16)) & 0x00000000FFFFFFFFull;
  return (uint32_t)x;
}

// unchecked: entries are offsets biased by data_size (non-negative), gp already holds -data_size.
// k32: the caller has checked that the low 32 bits of gp cannot carry when an entry is added, so
// every address is one 32-bit add (the high word is shared).
template <bool k32>
__device__ __forceinline__ const uint8_t* sweep_addr(const uint8_t* gp, unsigned lo, unsigned hi, unsigned o) {
  if (k32) return reinterpret_cast<const uint8_t*>(((unsigned long long)hi << 32) | (unsigned long long)(lo + o));
  return gp + o;
}

template <bool k32>
__device__ __forceinline__ unsigned sweep_list(const uint8_t* __restrict__ gp, const unsigned* __restrict__ s_e, int n) {
  unsigned sum0 = 0, sum1 = 0;
  const unsigned lo = (unsigned)reinterpret_cast<unsigned long long>(gp);
  const unsigned hi = (unsigned)(reinterpret_cast<unsigned long long>(gp) >> 32);
  int p = 0;
  for (; p + 8 <= n; p += 8) {
    const uint4 o0 = *reinterpret_cast<const uint4*>(s_e + p);
    const uint4 o1 = *reinterpret_cast<const uint4*>(s_e + p + 4);
    const unsigned v0 = __ldg(sweep_addr<k32>(gp, lo, hi, o0.x)), v1 = __ldg(sweep_addr<k32>(gp, lo, hi, o0.y));
    const unsigned v2 = __ldg(sweep_addr<k32>(gp, lo, hi, o0.z)), v3 = __ldg(sweep_addr<k32>(gp, lo, hi, o0.w));
    const unsigned v4 = __ldg(sweep_addr<k32>(gp, lo, hi, o1.x)), v5 = __ldg(sweep_addr<k32>(gp, lo, hi, o1.y));
    const unsigned v6 = __ldg(sweep_addr<k32>(gp, lo, hi, o1.z)), v7 = __ldg(sweep_addr<k32>(gp, lo, hi, o1.w));
    sum0 += v0 + v1 + v2 + v3;
    sum1 += v4 + v5 + v6 + v7;
  }
  for (; p < n; p++) sum0 += (unsigned)__ldg(sweep_addr<k32>(gp, lo, hi, s_e[p]));
  return sum0 + sum1;
}

// checked: entries are raw (signed) flat offsets; Karto's IsUpTo(index, dataSize) test per lookup
__device__ __forceinline__ unsigned sweep_list_checked(const uint8_t* __restrict__ grid, int base,
                                                       const unsigned* __restrict__ s_e, int n, unsigned dsz) {
  unsigned sum = 0;
  for (int p = 0; p < n; p++) {
    const unsigned idx = (unsigned)(base + (int)s_e[p]);
    if (idx < dsz) sum += (unsigned)__ldg(grid + idx);
  }
  return sum;
}

// CTA constants of k_sweep_pruned, made once by warp 0 (r02p: with every warp deriving them itself the
// prologue was 19 % of the kernel's instructions)
struct SweepCta {
  PassDev ps;
  TableDev tb;
  double cosine, sine;
  double ap;  // odometry angle penalty of this CTA's search angle
  int regular, sx, sy, xspan, yspan;
};

__global__ void __launch_bounds__(896, 2)
k_sweep_pruned(GridC g, PenaltyC pen, const PassDev* __restrict__ passes, const PassAngle* __restrict__ pa_list,
               const TableDev* __restrict__ tables, const double* __restrict__ trig,
               const double* __restrict__ pool, const uint8_t* __restrict__ grids,
               const uint32_t* __restrict__ rowmask, int rm_words, int tnx, double* __restrict__ resp,
               double* __restrict__ passmax, unsigned long long* __restrict__ cellmax, int rows_per_cta, int cw,
               int PB, unsigned long long* __restrict__ issued, const double* __restrict__ dpen, int dpen_nx) {
  // dpen: the matcher's coarse distance-penalty table [nY][nX] (host-made with the device's own IEEE operations;
  // every coarse pass of a handle shares the lattice), used when a pass has exactly that lattice width
  extern __shared__ __align__(16) uint32_t s_u[];
  __shared__ __align__(16) SweepCta s_c;
  __shared__ unsigned long long s_cmax;
  __shared__ int s_col[32], s_row[32];
  __shared__ unsigned s_issued;
  __shared__ int s_nsurv;
  const PassAngle pa = pa_list[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    // pass, table and the angle's cos / sin in one round of loads
    const int nps = (int)(sizeof(PassDev) / 4), ntb = (int)(sizeof(TableDev) / 4);
    for (int i = threadIdx.x; i < nps + ntb + 2; i += blockDim.x) {
      if (i < nps) reinterpret_cast<uint32_t*>(&s_c.ps)[i] = __ldg(reinterpret_cast<const uint32_t*>(passes + pa.pass) + i);
      else if (i < nps + ntb) reinterpret_cast<uint32_t*>(&s_c.tb)[i - nps] = __ldg(reinterpret_cast<const uint32_t*>(tables + pa.table) + (i - nps));
      else if (i == nps + ntb) s_c.cosine = trig[2 * pa.trig_i];
      else s_c.sine = trig[2 * pa.trig_i + 1];
    }
    if (threadIdx.x == 0) {
      s_issued = 0u;
      s_cmax = 0ull;
    }
  }
  __syncthreads();
  const int nX = s_c.ps.nX, nY = s_c.ps.nY;
  const int nxc = (nX + cw - 1) / cw;
  const int rg = blockIdx.y / nxc, xc = blockIdx.y - rg * nxc;
  const int iy0 = rg * rows_per_cta, ix0 = xc * cw;
  if (iy0 >= nY || ix0 >= nX) return;  // block-uniform
  const int nr = min(rows_per_cta, nY - iy0), nxl = min(cw, nX - ix0);
  // lattice cells of this CTA, exactly as CorrelateScan rounds them (A.7)
  if (warp == 0) {
    const PassDev& ps = s_c.ps;
    if (lane < nxl) {
      const double x = -ps.offx + (double)(ix0 + lane) * ps.resx;
      s_col[lane] = world_to_grid1(ps.cx + x, ps.gox, g.scale) + g.border;
    }
    if (lane < nr) {
      const double y = -ps.offy + (double)(iy0 + lane) * ps.resy;
      s_row[lane] = world_to_grid1(ps.cy + y, ps.goy, g.scale) + g.border;
    }
    __syncwarp();
    // pruning needs a regular lattice (step 1 or 2 cells): col(i) = col(0) + i*sx, row(i) = row(0) + i*sy
    const int sx = nxl > 1 ? s_col[1] - s_col[0] : 1, sy = nr > 1 ? s_row[1] - s_row[0] : 1;
    bool regular = (sx == 1 || sx == 2) && (sy == 1 || sy == 2);
    if (lane < nxl) regular = regular && s_col[lane] == s_col[0] + lane * sx;
    if (lane < nr) regular = regular && s_row[lane] == s_row[0] + lane * sy;
    regular = __all_sync(0xffffffffu, regular);
    if (lane == 0) {
      s_c.regular = regular ? 1 : 0;
      s_c.sx = sx;
      s_c.sy = sy;
      s_c.xspan = s_col[nxl - 1] - s_col[0];
      s_c.yspan = s_row[nr - 1] - s_row[0];
      s_c.ap = penalty_angle(ps, pen, pa.a);
    }
  }
  __syncthreads();
  uint2* s_pm = reinterpret_cast<uint2*>(s_u);
  uint32_t* s_list = s_u + 2 * PB + (size_t)warp * PB;  // warp-private
  const int P = s_c.ps.P;
  const int slot = s_c.ps.slot;
  const uint32_t* rm = rowmask + (size_t)slot * rm_words;
  const uint32_t rows_all = nr >= 32 ? 0xFFFFFFFFu : ((1u << nr) - 1u);
  const uint8_t* grid = grids + (size_t)slot * g.grid_bytes;
  const unsigned dsz = (unsigned)g.data_size;
  const bool row_warp = warp < nr;
  const bool active = row_warp && lane < nxl;
  const int base = s_row[row_warp ? warp : 0] * g.stride + s_col[active ? lane : 0];
  unsigned sum = 0;

  // can every lane's address be formed with a 32-bit add? (low word of gp + largest entry must not carry)
  const uint8_t* gp = grid + (base - (long long)dsz);
  const bool lo32 = __all_sync(0xffffffffu, (reinterpret_cast<unsigned long long>(gp) & 0xFFFFFFFFull) + 2ull * dsz <
                                               0x100000000ull);
  for (int pb = 0; pb < P; pb += PB) {
    const int nb = min(PB, P - pb);
    __syncthreads();  // the previous batch's s_pm is still being read; s_nsurv reset below; s_c complete
    if (threadIdx.x == 0) s_nsurv = 0;
    __syncthreads();
    // ---- A: offsets + per-point row masks; points that can see a non-zero cell survive ----------
    int ok = 1;
    {
      const double* qpts = pool + 2 * (size_t)s_c.tb.q_start;
      const double cosine = s_c.cosine, sine = s_c.sine;
      const int regular = s_c.regular, sy = s_c.sy, xspan = s_c.xspan, yspan = s_c.yspan;
      const int xa0 = s_col[0], ya0 = s_row[0];
      for (int i0 = 0; i0 < nb; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        uint32_t mask = 0u;
        int flat = 0;
        if (i < nb) {
          const double2 w = *reinterpret_cast<const double2*>(qpts + 2 * (size_t)(pb + i));
          int gx, gy;
          offset_cell(s_c.tb, g.scale, w.x, w.y, cosine, sine, gx, gy);
          flat = gx + gy * g.stride;
          const int xa = xa0 + gx, ya = ya0 + gy;
          mask = rows_all;
          if (regular && xa >= 0 && ya >= 0 && xa + xspan < g.width && ya + yspan < g.height) {
            const int txa = xa >> 5, txb = (xa + xspan) >> 5, tya = ya >> 5, tyb = (ya + yspan) >> 5;
            uint32_t m0 = 0u, m1 = 0u, m2 = 0u;
            {
              const uint32_t* r = rm + (size_t)tya * tnx;
              uint32_t v = __ldg(r + txa);
              if (txa + 1 <= txb) v |= __ldg(r + txa + 1);
              if (txa + 2 <= txb) v |= __ldg(r + txa + 2);
              m0 = v;
            }
            if (tya + 1 <= tyb) {
              const uint32_t* r = rm + (size_t)(tya + 1) * tnx;
              uint32_t v = __ldg(r + txa);
              if (txa + 1 <= txb) v |= __ldg(r + txa + 1);
              if (txa + 2 <= txb) v |= __ldg(r + txa + 2);
              m1 = v;
            }
            if (tya + 2 <= tyb) {
              const uint32_t* r = rm + (size_t)(tya + 2) * tnx;
              uint32_t v = __ldg(r + txa);
              if (txa + 1 <= txb) v |= __ldg(r + txa + 1);
              if (txa + 2 <= txb) v |= __ldg(r + txa + 2);
              m2 = v;
            }
            const int sft = ya & 31;
            const unsigned long long lo = (unsigned long long)m0 | ((unsigned long long)m1 << 32);
            const unsigned long long S = sft ? ((lo >> sft) | ((unsigned long long)m2 << (64 - sft))) : lo;
            mask = (sy == 2 ? even_bits64(S) : (uint32_t)S) & rows_all;
          } else {
            ok = 0;  // window may leave the grid: all rows, and Karto's flat bounds check for this batch
          }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, mask != 0u);
        int wbase = 0;
        if (lane == 0 && bal) wbase = atomicAdd(&s_nsurv, __popc(bal));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (mask) s_pm[wbase + __popc(bal & ((1u << lane) - 1u))] = make_uint2((unsigned)flat, mask);
      }
    }
    const int safe = __syncthreads_and(ok);
    const int nsurv = s_nsurv;
    // ---- B: this warp's list, C: its sums ------------------------------------------------------
    if (row_warp) {
      int cnt = 0;
      const unsigned bias = safe ? dsz : 0u;
      for (int i0 = 0; i0 < nsurv; i0 += 32) {
        const int i = i0 + lane;
        uint2 pm = make_uint2(0u, 0u);
        if (i < nsurv) pm = s_pm[i];
        const bool take = (pm.y >> warp) & 1u;
        const unsigned bal = __ballot_sync(0xffffffffu, take);
        if (take) s_list[cnt + __popc(bal & ((1u << lane) - 1u))] = pm.x + bias;
        cnt += __popc(bal);
      }
      __syncwarp();
      if (issued && lane == 0) atomicAdd(&s_issued, (unsigned)(cnt * nxl));
      if (!safe) sum += sweep_list_checked(grid, base, s_list, cnt, dsz);
      else if (lo32) sum += sweep_list<true>(gp, s_list, cnt);
      else sum += sweep_list<false>(gp, s_list, cnt);
      __syncwarp();
    }
  }
  // ---- epilogue: normalise, penalise, store; CTA maximum through the responses' bit patterns (>= 0) ----------
  unsigned long long bits = 0ull;
  if (active) {
    const PassDev& ps = s_c.ps;
    const int ix = ix0 + lane, iy = iy0 + warp;
    double rr = (double)sum / (double)((unsigned)P * 100u);
    if (ps.penalize && !kt_double_equal(rr, 0.0)) {
      const double dp = (dpen && dpen_nx == nX) ? __ldg(dpen + (size_t)iy * nX + ix) : penalty_distance(ps, pen, ix, iy);
      rr *= (dp * s_c.ap);
    }
    resp[ps.sums_off + ((size_t)iy * nX + ix) * ps.nA + pa.a] = rr;
    cell_max_update(cellmax, ps, ix, iy, rr);
    bits = (unsigned long long)__double_as_longlong(rr);
  }
  {
    const unsigned hi = (unsigned)(bits >> 32);
    const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? (unsigned)bits : 0u);
    if (lane == 0 && row_warp) atomicMax(&s_cmax, ((unsigned long long)mhi << 32) | mlo);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicMax(reinterpret_cast<unsigned long long*>(passmax + pa.pass), s_cmax);
    if (issued) atomicAdd(issued, (unsigned long long)s_issued);  // lookups actually performed (bench accounting)
  }
}

// ---------------------------------------------------------------------------------------------
// K3'  GetResponse sweep, point-parallel form for small search volumes (the fine pass,
// 3 x 3 x nA poses): one warp per pose, lanes stride the query points, integer partial sums
// combined with warp shuffles. grid = (ceil(nposes / warps_per_cta), n_fine_passes)
// ---------------------------------------------------------------------------------------------
// (one warp evaluates pose `pose` of pass `pid`)
__device__ __forceinline__ void
sweep_points_body(const GridC& g, const PenaltyC& pen, const PassDev& ps, const TableDev& tb, int pid, int pose,
                  const int* offsets, const uint8_t* grids, double* resp, double* passmax) {
  const int nposes = ps.nX * ps.nY * ps.nA;
  const int lane = threadIdx.x & 31;
  if (pose >= nposes) return;
  const int iy = pose / (ps.nX * ps.nA);
  const int rem = pose - iy * ps.nX * ps.nA;
  const int ix = rem / ps.nA, a = rem - ix * ps.nA;
  const double x = -ps.offx + (double)ix * ps.resx;
  const double y = -ps.offy + (double)iy * ps.resy;
  const int gx = world_to_grid1(ps.cx + x, ps.gox, g.scale) + g.border;
  const int gy = world_to_grid1(ps.cy + y, ps.goy, g.scale) + g.border;
  const int base = gx + gy * g.stride;
  const int* goff = offsets + tb.out_off + (size_t)a * tb.Ppad;
  const uint8_t* grid = grids + (size_t)ps.slot * g.grid_bytes;
  const unsigned dsz = (unsigned)g.data_size;
  unsigned sum = 0;
  for (int p0 = lane; p0 < ps.P; p0 += 128) {
    unsigned idx[4];
#pragma unroll
    for (int u = 0; u < 4; u++) idx[u] = (p0 + 32 * u < ps.P) ? (unsigned)(base + goff[p0 + 32 * u]) : 0xFFFFFFFFu;
#pragma unroll
    for (int u = 0; u < 4; u++) if (idx[u] < dsz) sum += (unsigned)__ldg(grid + idx[u]);
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) {
    const double rr = response_of(ps, pen, sum, ix, iy, a);
    resp[ps.sums_off + pose] = rr;
    pass_max_update(passmax, pid, rr);
  }
}

__global__ void __launch_bounds__(256)
k_sweep_points(GridC g, PenaltyC pen, const PassDev* __restrict__ passes, const int* __restrict__ pass_ids,
               const TableDev* __restrict__ tables, const int* __restrict__ offsets,
               const uint8_t* __restrict__ grids, double* __restrict__ resp, double* __restrict__ passmax) {
  const int pid = pass_ids[blockIdx.y];
  const PassDev ps = passes[pid];
  const TableDev tb = tables[ps.table];
  sweep_points_body(g, pen, ps, tb, pid, (int)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)), offsets, grids,
                    resp, passmax);
}

// ---------------------------------------------------------------------------------------------
// K3' (3 x 3 form)  the fine pass of the throughput path: its lattice is 3 x 3 cells, so a warp takes one
// (pass, angle): every lane reads a point's lookup offset ONCE and sums the nine cells around it (k_sweep_points
// read the offset once per pose and kept one pose per warp: nine times the table traffic, 99 short dependent
// chains per match). One CTA per fine pass; nine REDUX per warp, lanes 0-8 finish their pose.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(384)
k_sweep_fine9(GridC g, PenaltyC pen, const PassDev* __restrict__ passes, const int* __restrict__ pass_ids,
              const TableDev* __restrict__ tables, const int* __restrict__ offsets,
              const uint8_t* __restrict__ grids, double* __restrict__ resp, double* __restrict__ passmax,
              unsigned* __restrict__ isums) {
  __shared__ PassDev s_ps;
  __shared__ int s_out_off, s_ppad;
  const int pid = pass_ids[blockIdx.x];
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(passes + pid);
    if (threadIdx.x < sizeof(PassDev) / 4) reinterpret_cast<uint32_t*>(&s_ps)[threadIdx.x] = __ldg(src + threadIdx.x);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const TableDev* tb = tables + s_ps.table;
    s_out_off = tb->out_off;
    s_ppad = tb->Ppad;
  }
  __syncthreads();
  const PassDev& ps = s_ps;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  // the nine lattice cells, rounded exactly as CorrelateScan rounds them (A.7)
  int cb[9];
#pragma unroll
  for (int iy = 0; iy < 3; iy++) {
    const double y = -ps.offy + (double)iy * ps.resy;
    const int gy = world_to_grid1(ps.cy + y, ps.goy, g.scale) + g.border;
#pragma unroll
    for (int ix = 0; ix < 3; ix++) {
      const double x = -ps.offx + (double)ix * ps.resx;
      const int gx = world_to_grid1(ps.cx + x, ps.gox, g.scale) + g.border;
      cb[iy * 3 + ix] = gx + gy * g.stride;
    }
  }
  const uint8_t* grid = grids + (size_t)ps.slot * g.grid_bytes;
  const unsigned dsz = (unsigned)g.data_size;
  const int P = ps.P, nA = ps.nA;
  for (int a = warp; a < nA; a += nwarps) {
    const int* goff = offsets + s_out_off + (size_t)a * s_ppad;
    unsigned sum[9];
#pragma unroll
    for (int k = 0; k < 9; k++) sum[k] = 0u;
    for (int p = lane; p < P; p += 32) {
      const int o = __ldg(goff + p);
#pragma unroll
      for (int k = 0; k < 9; k++) {
        const unsigned idx = (unsigned)(cb[k] + o);
        if (idx < dsz) sum[k] += (unsigned)__ldg(grid + idx);  // Karto's IsUpTo(index, dataSize)
      }
    }
    unsigned mine = 0u;
#pragma unroll
    for (int k = 0; k < 9; k++) {
      const unsigned t = __reduce_add_sync(0xffffffffu, sum[k]);
      if (lane == k) mine = t;
    }
    unsigned long long bits = 0ull;
    if (lane < 9) {
      const int iy = lane / 3, ix = lane - iy * 3;
      const double rr = response_of(ps, pen, mine, ix, iy, a);
      resp[ps.sums_off + (iy * 3 + ix) * nA + a] = rr;
      isums[ps.sums_off + (iy * 3 + ix) * nA + a] = mine;  // (the reduce's angular-covariance sums come from here)
      bits = (unsigned long long)__double_as_longlong(rr);
    }
    const unsigned hi = (unsigned)(bits >> 32);
    const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? (unsigned)bits : 0u);
    if (lane == 0) atomicMax(reinterpret_cast<unsigned long long*>(passmax + pid), ((unsigned long long)mhi << 32) | mlo);
  }
}

// ---------------------------------------------------------------------------------------------
// K3b/K4  CorrelateScan epilogue: response normalisation + odometry penalty,
// best response, Karto's average-of-ties best pose (sequential, in y/x/angle storage order so
// the double sums round exactly as the CPU's), then ComputePositionalCovariance accumulators
// (coarse, A.9) or the per-angle response sums ComputeAngularCovariance needs (fine, A.9).
// One CTA per pass.
// ---------------------------------------------------------------------------------------------
#define YSM_TIE_CAP 1024

__device__ __forceinline__ double block_reduce_max(double v, double* s_tmp) {
  for (int o = 16; o > 0; o >>= 1) {
    const double t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t > v ? t : v;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_tmp[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = s_tmp[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); w++) r = s_tmp[w] > r ? s_tmp[w] : r;
  return r;
}

__device__ __forceinline__ double block_reduce_sum(double v, double* s_tmp) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_tmp[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); w++) r += s_tmp[w];
  return r;
}

// (one CTA of <= 512 threads reduces pass `ps`; *po may live in shared memory)
__device__ __forceinline__ void
reduce_body(const GridC& g, const PassDev& ps, const TableDev& tb, const int* offsets, const double* resp, double best,
            const unsigned long long* cellmax, const double* trig, const uint8_t* grids, PassOut* po, int* angsums,
            const unsigned* isums = nullptr) {
  __shared__ int s_list[YSM_TIE_CAP];
  __shared__ int s_sorted[YSM_TIE_CAP];
  __shared__ int s_count;
  __shared__ unsigned s_bits[16];
  __shared__ double s_acc[4];
  __shared__ int s_n, s_first;

  const double* pr = resp + ps.sums_off;
  const int nposes = ps.nX * ps.nY * ps.nA;
  const int tid = threadIdx.x;
  if (nposes == 0) {  // speculative fine pass whose coarse pass did not end in a single winner
    if (tid == 0) {
      po->best = 0.0; po->avg_x = 0.0; po->avg_y = 0.0; po->tx = 0.0; po->ty = 0.0;
      po->norm = 0.0; po->axx = 0.0; po->axy = 0.0; po->ayy = 0.0;
      po->n_ties = -1; po->first_idx = -1;
    }
    return;
  }
  if (tid == 0) s_count = 0;
  // best response: accumulated by the sweep kernels (max over all poses; Karto's init of -1
  // never survives because every pass has at least one pose and responses are >= 0)
  __syncthreads();

  // per-cell maxima of a coarse pass (accumulated by the sweep): only cells whose maximum is within
  // the tie tolerance of the best can hold a tied pose, and the maxima ARE the search-space probs
  const unsigned long long* cm = (!ps.fine && cellmax && ps.cmax_off >= 0) ? cellmax + ps.cmax_off : nullptr;
  if (cm) {
    const int ncell = ps.nX * ps.nY;
    for (int c = tid; c < ncell; c += blockDim.x) {
      const double m = __longlong_as_double((long long)__ldcg(cm + c));
      if (m >= best - YSM_KT_TOLERANCE) {
        const double* pc = pr + (size_t)c * ps.nA;
        for (int a0 = 0; a0 < ps.nA; a0 += 8) {
          double r[8];
#pragma unroll
          for (int u = 0; u < 8; u++) r[u] = a0 + u < ps.nA ? pc[a0 + u] : -1.0;  // independent loads
#pragma unroll
          for (int u = 0; u < 8; u++) {
            if (r[u] >= 0.0 && kt_double_equal(r[u], best)) {
              const int pos = atomicAdd(&s_count, 1);
              if (pos < YSM_TIE_CAP) s_list[pos] = c * ps.nA + a0 + u;
            }
          }
        }
      }
    }
  }
  // poses tied with the best, in storage order (loads batched 4 deep to overlap L2 latency)
  for (int i0 = tid; i0 < (cm ? 0 : nposes); i0 += 4 * blockDim.x) {
    double r[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = i0 + u * blockDim.x;
      r[u] = i < nposes ? pr[i] : -1.0;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (r[u] >= 0.0 && kt_double_equal(r[u], best)) {
        const int pos = atomicAdd(&s_count, 1);
        if (pos < YSM_TIE_CAP) s_list[pos] = i0 + u * blockDim.x;
      }
    }
  }
  __syncthreads();
  const int nt = s_count;
  const double startX = -ps.offx, startY = -ps.offy;
  const double* htrig = trig + 2 * (size_t)ps.htrig_off;
  if (nt <= YSM_TIE_CAP) {
    for (int e = tid; e < nt; e += blockDim.x) {
      const int v = s_list[e];
      int rank = 0;
      for (int k = 0; k < nt; k++) rank += (s_list[k] < v);
      s_sorted[rank] = v;
    }
    __syncthreads();
    if (tid == 0) {
      double sx = 0.0, sy = 0.0, tx = 0.0, ty = 0.0;
      for (int e = 0; e < nt; e++) {
        const int idx = s_sorted[e];
        const int iy = idx / (ps.nX * ps.nA);
        const int rem = idx - iy * ps.nX * ps.nA;
        const int ix = rem / ps.nA, a = rem - ix * ps.nA;
        sx += ps.cx + (startX + (double)ix * ps.resx);
        sy += ps.cy + (startY + (double)iy * ps.resy);
        tx += htrig[2 * a];
        ty += htrig[2 * a + 1];
      }
      s_acc[0] = sx; s_acc[1] = sy; s_acc[2] = tx; s_acc[3] = ty;
      s_n = nt;
      s_first = nt > 0 ? s_sorted[0] : -1;
    }
  } else if (nt == nposes) {
    // EVERY pose ties (best == 0: an empty grid, the response-expansion case). The four ordered sums are
    // independent of each other: four threads of four different warps each run one of them as plain nested
    // loops in storage order (y, x, angle) -- the same additions in the same order as the CPU loop, without
    // the per-pose index arithmetic and memory reads of the general path below.
    __shared__ double s_ht[2 * 128];
    const bool ht_smem = ps.nA <= 128;
    if (ht_smem)
      for (int i = tid; i < 2 * ps.nA; i += blockDim.x) s_ht[i] = htrig[i];
    __syncthreads();
    if ((tid & 31) == 0 && tid < 128) {
      const int which = tid >> 5;
      double acc = 0.0;
      if (which == 0) {
        for (int iy = 0; iy < ps.nY; iy++)
          for (int ix = 0; ix < ps.nX; ix++) {
            const double v = ps.cx + (startX + (double)ix * ps.resx);
            for (int a = 0; a < ps.nA; a++) acc += v;
          }
      } else if (which == 1) {
        for (int iy = 0; iy < ps.nY; iy++) {
          const double v = ps.cy + (startY + (double)iy * ps.resy);
          for (int k = 0; k < ps.nX * ps.nA; k++) acc += v;
        }
      } else {
        const int o = which - 2;
        for (int c = 0; c < ps.nY * ps.nX; c++) {
          if (ht_smem)
            for (int a = 0; a < ps.nA; a++) acc += s_ht[2 * a + o];
          else
            for (int a = 0; a < ps.nA; a++) acc += htrig[2 * a + o];
        }
      }
      s_acc[which] = acc;
    }
    if (tid == 0) {
      s_n = nt;
      s_first = 0;
    }
  } else {
    // many (not all) poses tie: ordered chunks of blockDim poses,
    // one thread accumulates sequentially so the rounding matches the CPU loop
    double sx = 0.0, sy = 0.0, tx = 0.0, ty = 0.0;
    int n = 0, first = -1;
    for (int c0 = 0; c0 < nposes; c0 += blockDim.x) {
      const int i = c0 + tid;
      bool tie = false;
      if (i < nposes) tie = kt_double_equal(pr[i], best);
      const unsigned bal = __ballot_sync(0xffffffffu, tie);
      if ((tid & 31) == 0) s_bits[tid >> 5] = bal;
      __syncthreads();
      if (tid == 0) {
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) {
          unsigned b = s_bits[w];
          while (b) {
            const int bit = __ffs(b) - 1;
            b &= b - 1;
            const int idx = c0 + w * 32 + bit;
            const int iy = idx / (ps.nX * ps.nA);
            const int rem = idx - iy * ps.nX * ps.nA;
            const int ix = rem / ps.nA, a = rem - ix * ps.nA;
            sx += ps.cx + (startX + (double)ix * ps.resx);
            sy += ps.cy + (startY + (double)iy * ps.resy);
            tx += htrig[2 * a];
            ty += htrig[2 * a + 1];
            if (n == 0) first = idx;
            n++;
          }
        }
      }
      __syncthreads();
    }
    if (tid == 0) {
      s_acc[0] = sx; s_acc[1] = sy; s_acc[2] = tx; s_acc[3] = ty;
      s_n = n;
      s_first = first;
    }
  }
  __syncthreads();
  const int n = s_n;
  const double cnt = (double)n;
  const double avg_x = n > 0 ? s_acc[0] / cnt : 0.0;
  const double avg_y = n > 0 ? s_acc[1] / cnt : 0.0;
  if (tid == 0) {
    po->best = best;
    po->avg_x = avg_x;
    po->avg_y = avg_y;
    po->tx = n > 0 ? s_acc[2] / cnt : 0.0;
    po->ty = n > 0 ? s_acc[3] / cnt : 0.0;
    po->n_ties = n;
    po->first_idx = s_first;
  }
  if (!ps.fine) {
    // ComputePositionalCovariance accumulators over the (y, x) lattice; probs(x, y) is the
    // max response over angles of that lattice cell (m_pSearchSpaceProbs).
    double norm = 0.0, axx = 0.0, axy = 0.0, ayy = 0.0;
    if (!(best < YSM_KT_TOLERANCE)) {
      const double dx = avg_x - ps.cx, dy = avg_y - ps.cy;
      const int ncell = ps.nX * ps.nY;
      for (int c = tid; c < ncell; c += blockDim.x) {
        const int iy = c / ps.nX, ix = c - iy * ps.nX;
        double pm = 0.0;  // probs grid is cleared to 0 and max'ed with every response
        if (cm) {
          pm = __longlong_as_double((long long)__ldcg(cm + c));
        } else {
          const double* pc = pr + (size_t)c * ps.nA;
          int a = 0;
          for (; a + 4 <= ps.nA; a += 4) {
            const double r0 = pc[a], r1 = pc[a + 1], r2 = pc[a + 2], r3 = pc[a + 3];
            const double m01 = r0 > r1 ? r0 : r1, m23 = r2 > r3 ? r2 : r3;
            const double m = m01 > m23 ? m01 : m23;
            pm = m > pm ? m : pm;
          }
          for (; a < ps.nA; a++) {
            const double r = pc[a];
            pm = r > pm ? r : pm;
          }
        }
        if (pm >= (best - 0.1)) {
          const double x = startX + (double)ix * ps.resx;
          const double y = startY + (double)iy * ps.resy;
          norm += pm;
          axx += ((x - dx) * (x - dx) * pm);
          axy += ((x - dx) * (y - dy) * pm);
          ayy += ((y - dy) * (y - dy) * pm);
        }
      }
    }
    // the four sums in one pass (same order per sum as block_reduce_sum: xor-shuffles, then the warps in order)
    {
      __shared__ double s_tmp4[32][4];
      for (int o = 16; o > 0; o >>= 1) {
        norm += __shfl_xor_sync(0xffffffffu, norm, o);
        axx += __shfl_xor_sync(0xffffffffu, axx, o);
        axy += __shfl_xor_sync(0xffffffffu, axy, o);
        ayy += __shfl_xor_sync(0xffffffffu, ayy, o);
      }
      __syncthreads();
      if ((tid & 31) == 0) {
        s_tmp4[tid >> 5][0] = norm; s_tmp4[tid >> 5][1] = axx; s_tmp4[tid >> 5][2] = axy; s_tmp4[tid >> 5][3] = ayy;
      }
      __syncthreads();
      if (tid < 4) {
        double r = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) r += s_tmp4[w][tid];
        s_acc[tid] = r;  // (s_acc is free again: avg_x / avg_y were read above)
      }
      __syncthreads();
      norm = s_acc[0]; axx = s_acc[1]; axy = s_acc[2]; ayy = s_acc[3];
    }
    if (tid == 0) {
      po->norm = norm; po->axx = axx; po->axy = axy; po->ayy = ayy;
    }
  } else {
    // ComputeAngularCovariance: un-penalised GetResponse at the best cell for every fine angle
    if (tid == 0) { po->norm = 0.0; po->axx = 0.0; po->axy = 0.0; po->ayy = 0.0; }
    if (n > 0) {
      const int gx = world_to_grid1(avg_x, ps.gox, g.scale) + g.border;
      const int gy = world_to_grid1(avg_y, ps.goy, g.scale) + g.border;
      // The best cell is one of the fine lattice's 3 x 3 cells whenever the (tie-averaged) best pose rounds onto
      // one -- nearly always -- and then the sums asked for are exactly the integer sums k_sweep_fine9 made for
      // that cell (same offsets, same bounds check). r02z: re-gathering them cost 78 % of the fine reduce's warp
      // time, 7,920 scattered bytes per match fetched from DRAM again.
      int cell = -1;
      if (isums != nullptr && ps.nX == 3 && ps.nY == 3) {  // block-uniform
        __shared__ int s_cell;
        if (tid == 0) s_cell = -1;
        __syncthreads();
        if (tid < 9) {
          const int iy = tid / 3, ix = tid - iy * 3;
          const double x = -ps.offx + (double)ix * ps.resx;
          const double y = -ps.offy + (double)iy * ps.resy;
          const int cx = world_to_grid1(ps.cx + x, ps.gox, g.scale) + g.border;
          const int cy = world_to_grid1(ps.cy + y, ps.goy, g.scale) + g.border;
          if (cx == gx && cy == gy) atomicMax(&s_cell, tid);
        }
        __syncthreads();
        cell = s_cell;
        if (cell >= 0)
          for (int a = tid; a < ps.nA; a += blockDim.x) angsums[ps.ang_off + a] = (int)isums[ps.sums_off + cell * ps.nA + a];
      }
      const int base = gx + gy * g.stride;
      const uint8_t* grid = grids + (size_t)ps.slot * g.grid_bytes;
      const unsigned dsz = (unsigned)g.data_size;
      const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
      for (int a = warp; a < (cell >= 0 ? 0 : ps.nA); a += nwarps) {
        const int* goff = offsets + tb.out_off + (size_t)a * tb.Ppad;
        unsigned sum = 0;
        for (int p0 = lane; p0 < ps.P; p0 += 384) {
          unsigned idx[12];
#pragma unroll
          for (int u = 0; u < 12; u++) idx[u] = (p0 + 32 * u < ps.P) ? (unsigned)(base + goff[p0 + 32 * u]) : 0xFFFFFFFFu;
#pragma unroll
          for (int u = 0; u < 12; u++) if (idx[u] < dsz) sum += (unsigned)__ldg(grid + idx[u]);
        }
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) angsums[ps.ang_off + a] = (int)sum;
      }
    }
  }
}

// latency path: point the speculative fine pass of coarse pass `ps` at its winner (result *po).
// Valid only for a single winning pose with a non-zero response (then MatchScan goes straight to
// the fine pass, its centre is that lattice pose and the heading the atan2(sin, cos) the host
// tabulated); otherwise nA = 0 and the host reschedules the match through the general path.
__device__ __forceinline__ void spec_resolve(const PassDev& ps, const PassOut& po, const double* trig, PassDev* f,
                                             TableDev* ft) {
  if (po.n_ties == 1 && po.best > YSM_KT_TOLERANCE) {
    const int a = po.first_idx % ps.nA;
    f->cx = po.avg_x;
    f->cy = po.avg_y;
    f->ch = trig[ps.spec_h_off + a];
    f->htrig_off = ps.spec_htrig_off + a * ps.spec_nAf;
    ft->trig_off = ps.spec_trig_off + a * ps.spec_nAf;
  } else {
    f->nA = 0;
    ft->nA = 0;
  }
}

__global__ void __launch_bounds__(512)
k_reduce(GridC g, PassDev* passes, TableDev* tables,
         const int* __restrict__ offsets, const double* __restrict__ resp,
         const double* __restrict__ passmax, const unsigned long long* __restrict__ cellmax,
         const double* __restrict__ trig, const uint8_t* __restrict__ grids, PassOut* __restrict__ outs,
         int* __restrict__ angsums, int pass_base, const unsigned* __restrict__ isums) {
  // isums: integer sums of the fine passes' poses when k_sweep_fine9 made them ([sums_off + pose]); else null
  const int pid = pass_base + blockIdx.x;
  const PassDev ps = passes[pid];
  const TableDev tb = tables[ps.table];
  reduce_body(g, ps, tb, offsets, resp, passmax[pid], cellmax, trig, grids, outs + pid, angsums, isums);
  if (ps.spec >= 0 && threadIdx.x == 0) {
    PassDev* f = passes + ps.spec;
    spec_resolve(ps, outs[pid], trig, f, tables + f->table);  // thread 0 wrote outs[pid] itself
  }
}

// ---------------------------------------------------------------------------------------------
// KL  latency path: ONE cooperative kernel runs a whole MatchScan for a handful of matches --
// P0 pull the staging blob (points, descriptors, pass tables) from mapped host memory,
// P1 FindValidPoints (+ the ordered-stamp filter), P2 tile stamping, P3 the coarse lattice sweep
// with ComputeOffsets fused, P4 per match: reduce -> resolve the fine pass at the winner ->
// fine offsets -> fine sweep -> reduce, results and a completion flag written straight to
// mapped host memory (the host polls the flag: no copy engine, no stream synchronisation).
// Phases are separated by grid-wide barriers; every phase is the body of the stand-alone kernel.
// ---------------------------------------------------------------------------------------------
// cooperative copy of a descriptor another CTA of the same kernel just rewrote (L2 reads: the L1 of
// this SM may still hold the line from an earlier phase)
template <typename T>
__device__ __forceinline__ void load_struct_cg(T* dst_smem, const T* src, int tid) {
  static_assert(sizeof(T) % 8 == 0, "descriptor size");
  if (tid < (int)(sizeof(T) / 8))
    reinterpret_cast<unsigned long long*>(dst_smem)[tid] = __ldcg(reinterpret_cast<const unsigned long long*>(src) + tid);
}

// fine CorrelateScan sweep of ONE angle of pass f by one CTA: the angle's lookup offsets go to shared
// memory (and to the table row the angular covariance reads later), then one warp per (x, y) pose
__device__ __forceinline__ void
fine_angle_body(const GridC& g, const PenaltyC& pen, const PassDev& f, const TableDev& ft, int fid, int a,
                const double* trig, const double* pool, int* offsets_tab, const uint8_t* grids, double* resp,
                double* passmax, int* s_off) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const double cosine = trig[2 * (ft.trig_off + a)], sine = trig[2 * (ft.trig_off + a) + 1];
  int* row = offsets_tab + ft.out_off + (size_t)a * ft.Ppad;
  for (int p = tid; p < ft.P; p += blockDim.x) {
    const double2 w = *reinterpret_cast<const double2*>(pool + 2 * (size_t)(ft.q_start + p));
    int gx, gy;
    offset_cell(ft, g.scale, w.x, w.y, cosine, sine, gx, gy);
    const int o = gx + gy * g.stride;
    s_off[p] = o;
    row[p] = o;
  }
  __syncthreads();
  const uint8_t* grid = grids + (size_t)f.slot * g.grid_bytes;
  const unsigned dsz = (unsigned)g.data_size;
  const int nxy = f.nX * f.nY;
  for (int c = warp; c < nxy; c += nwarps) {
    const int iy = c / f.nX, ix = c - iy * f.nX;
    const double x = -f.offx + (double)ix * f.resx;
    const double y = -f.offy + (double)iy * f.resy;
    const int gx = world_to_grid1(f.cx + x, f.gox, g.scale) + g.border;
    const int gy = world_to_grid1(f.cy + y, f.goy, g.scale) + g.border;
    const int base = gx + gy * g.stride;
    unsigned sum = 0;
    for (int p0 = lane; p0 < f.P; p0 += 256) {
      unsigned idx[8];
#pragma unroll
      for (int u = 0; u < 8; u++) idx[u] = (p0 + 32 * u < f.P) ? (unsigned)(base + s_off[p0 + 32 * u]) : 0xFFFFFFFFu;
#pragma unroll
      for (int u = 0; u < 8; u++) if (idx[u] < dsz) sum += (unsigned)__ldg(grid + idx[u]);
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) {
      const double rr = response_of(f, pen, sum, ix, iy, a);
      resp[f.sums_off + (size_t)c * f.nA + a] = rr;
      pass_max_update(passmax, fid, rr);
    }
  }
}

struct SmallArgs {
  const uint4* blob_src;  // mapped host memory
  uint4* blob_dst;        // device copy
  int blob_vec;           // 16-byte units
  int pool_in_blob;
  const double* pool_dev; // caller's device pool when not in the blob
  unsigned o_pool, o_scan_start, o_scan_count, o_matches, o_base, o_workcount, o_tab, o_pass, o_pa, o_trig, o_pmax;
  unsigned o_scanlist;
  int nscans_total, tiles_per_grid;
  int* scan_emit;
  unsigned char* tileflag;  // [nw][tiles_per_grid] bytes
  int nw, npa, ncoarse, nspec, nAf;
  int pmax, nbase_max, stage, fv_warps, ordered, log2cap;
  int tpc, psplit, task_chunks;
  uint32_t* ptcell;
  uint32_t* cells;
  int* cellcount;
  uint2* gbox;
  int2* work;
  const uint16_t* stamp_tab;
  uint8_t* grids;
  uint32_t* rowmask;
  int rm_words;
  int epoch;
  int* offsets;
  double* resp;
  unsigned long long* cellmax;
  int cellmax_n;          // elements to zero
  PassOut* outs_host;     // mapped host memory
  int* angs_host;         // mapped host memory
  volatile int* flags_host;  // mapped host memory: flags_host[pid] = epoch when pass pid's results are visible
  unsigned long long* tstamps;  // optional (tracing): %globaltimer at the phase boundaries, CTA 0
};

__global__ void __launch_bounds__(512, 2)
k_match_small(GridC g, PenaltyC pen, SmallArgs A) {
  extern __shared__ __align__(16) unsigned char dsm[];
  __shared__ PassOut s_po[2];
  __shared__ PassDev s_f, s_ps;
  __shared__ TableDev s_ft, s_tb;
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#define YSM_TSTAMP(k)                                                                  \
  if (A.tstamps && blockIdx.x == 0 && tid == 0) {                                      \
    unsigned long long t_;                                                             \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                             \
    A.tstamps[k] = t_;                                                                 \
  }
  YSM_TSTAMP(0)
  // ---- P0: staging blob, host -> device ---------------------------------------------------------
  for (int i = blockIdx.x * blockDim.x + tid; i < A.blob_vec; i += gridDim.x * blockDim.x) A.blob_dst[i] = A.blob_src[i];
  for (int i = blockIdx.x * blockDim.x + tid; i < A.cellmax_n; i += gridDim.x * blockDim.x) A.cellmax[i] = 0ull;
  for (int i = blockIdx.x * blockDim.x + tid; i < (A.nw * A.tiles_per_grid + 3) / 4; i += gridDim.x * blockDim.x)
    reinterpret_cast<unsigned*>(A.tileflag)[i] = 0u;
  grid.sync();
  YSM_TSTAMP(1)
  unsigned char* db = reinterpret_cast<unsigned char*>(A.blob_dst);
  const double* pool = A.pool_in_blob ? reinterpret_cast<const double*>(db + A.o_pool) : A.pool_dev;
  const int* scan_start = reinterpret_cast<const int*>(db + A.o_scan_start);
  const int* scan_count = reinterpret_cast<const int*>(db + A.o_scan_count);
  const MatchDev* matches = reinterpret_cast<const MatchDev*>(db + A.o_matches);
  const int* base_idx = reinterpret_cast<const int*>(db + A.o_base);
  int* work_count = reinterpret_cast<int*>(db + A.o_workcount);
  TableDev* tables = reinterpret_cast<TableDev*>(db + A.o_tab);
  PassDev* passes = reinterpret_cast<PassDev*>(db + A.o_pass);
  const PassAngle* pa_list = reinterpret_cast<const PassAngle*>(db + A.o_pa);
  const double* trig = reinterpret_cast<const double*>(db + A.o_trig);
  double* passmax = reinterpret_cast<double*>(db + A.o_pmax);
  // The coarse reduce of pass p runs on CTA gridDim-1-p (CTAs from the end of the grid have no scan to
  // filter and usually no sweep task): it fetches its descriptors -- and the unresolved ones of the
  // pass's speculative fine pass -- now, off the critical path (two dependent L2 round trips each).
  const int my_pid = (int)gridDim.x - 1 - (int)blockIdx.x;
  const bool pre = my_pid < A.ncoarse;
  if (pre) {
    load_struct_cg(&s_ps, passes + my_pid, tid);
    __syncthreads();
    load_struct_cg(&s_tb, tables + s_ps.table, tid);
    if (s_ps.spec >= 0) load_struct_cg(&s_f, passes + s_ps.spec, tid);
    __syncthreads();
    if (s_ps.spec >= 0) load_struct_cg(&s_ft, tables + s_f.table, tid);
    __syncthreads();
  }
  // ---- P1a: FindValidPoints, CTA per base scan ------------------------------------------------------
  const ScanRef* scanlist = reinterpret_cast<const ScanRef*>(db + A.o_scanlist);
  for (int v = blockIdx.x; v < A.nscans_total; v += gridDim.x) {
    const ScanRef sr = scanlist[v];
    fv_scan_body(g, matches, base_idx, scan_start, scan_count, pool, sr, A.ptcell, A.scan_emit, v, A.tileflag,
                 A.tiles_per_grid, A.pmax, dsm);
    __syncthreads();
  }
  __threadfence();
  YSM_TSTAMP(13)
  grid.sync();
  YSM_TSTAMP(14)
  // ---- P1b: per match: cells in scan order, bounding boxes, tile work list (all CTAs) --------------
  {
    // the CTAs are dealt to the matches round-robin
    const int per = max(1, (int)gridDim.x / A.nw);
    for (int vb = 0; vb < A.nw; vb++) {
      const int lo = vb * per, hi = vb == A.nw - 1 ? (int)gridDim.x : lo + per;
      if ((int)blockIdx.x >= lo && (int)blockIdx.x < hi)
        fv_match_body(g, matches, scanlist, A.nscans_total, A.scan_emit, A.ptcell, A.cells, A.cellcount, A.gbox, A.work,
                      work_count, A.tileflag, A.tiles_per_grid, vb, (int)blockIdx.x - lo, hi - lo);
    }
  }
  if (A.ordered) {
    // the ordered-stamp filter needs the whole cell list of a match: one more barrier
    __threadfence();
    grid.sync();
    for (int vb = blockIdx.x; vb < A.nw; vb += gridDim.x) {
      if (warp == 0) stamp_order_body(matches, A.cells, A.cellcount, A.log2cap, vb, lane, reinterpret_cast<uint32_t*>(dsm));
      __syncthreads();
    }
  }
  YSM_TSTAMP(2)
  grid.sync();
  YSM_TSTAMP(3)
  // ---- P2: SmearPoint ---------------------------------------------------------------------------------
  {
    // as many warps per tile as the machine has to spare (<= 8: named barriers 1..8)
    const int nwork = *work_count, wtot = (int)gridDim.x * (int)(blockDim.x >> 5);
    int S = 8;
    while (S > 1 && (long long)nwork * S > wtot) S >>= 1;
    tile_stamp_body(g, matches, A.cells, A.cellcount, A.gbox, A.work, work_count, A.stamp_tab, A.grids, A.rowmask,
                    A.rm_words, (int)blockIdx.x, (int)gridDim.x, dsm, S);
  }
  YSM_TSTAMP(4)
  grid.sync();
  YSM_TSTAMP(5)
  // ---- P3: coarse CorrelateScan sweep -----------------------------------------------------------------
  const int nv = A.npa * A.task_chunks;
  for (int v = blockIdx.x; v < nv; v += gridDim.x) {
    sweep_lattice_body(g, pen, passes, pa_list, tables, nullptr, trig, pool, A.grids, A.resp, passmax, A.cellmax,
                       A.tpc, A.psplit, v % A.npa, v / A.npa, reinterpret_cast<int*>(dsm));
    __syncthreads();
  }
  YSM_TSTAMP(6)
  grid.sync();
  YSM_TSTAMP(7)
  // ---- P4a: per coarse pass: reduce, resolve its fine pass at the winner ------------------------------
  const int nw8 = (int)(sizeof(PassOut) / 8);
  for (int pid = my_pid; pid < A.ncoarse; pid += gridDim.x) {  // (one iteration: the grid is larger than ncoarse)
    const PassDev& ps = s_ps;
    const TableDev& tb = s_tb;
    reduce_body(g, ps, tb, A.offsets, A.resp, __ldcg(passmax + pid), A.cellmax, trig, A.grids, &s_po[0], A.angs_host);
    __syncthreads();
    if (tid < nw8) reinterpret_cast<double*>(A.outs_host + pid)[tid] = reinterpret_cast<const double*>(&s_po[0])[tid];
    if (ps.spec >= 0) {
      if (tid == 0) spec_resolve(ps, s_po[0], trig, &s_f, &s_ft);  // s_f / s_ft: preloaded above
      __syncthreads();
      // publish the resolved descriptors for the CTAs of the next phases
      if (tid < (int)(sizeof(PassDev) / 8))
        reinterpret_cast<unsigned long long*>(passes + ps.spec)[tid] = reinterpret_cast<const unsigned long long*>(&s_f)[tid];
      if (tid >= 32 && tid < 32 + (int)(sizeof(TableDev) / 8))
        reinterpret_cast<unsigned long long*>(tables + s_f.table)[tid - 32] =
            reinterpret_cast<const unsigned long long*>(&s_ft)[tid - 32];
      __threadfence();
    } else {
      __threadfence_system();
      __syncthreads();
      if (tid == 0) {
        __threadfence_system();
        A.flags_host[pid] = A.epoch;
      }
    }
    __syncthreads();
  }
  YSM_TSTAMP(8)
  grid.sync();
  YSM_TSTAMP(9)
  // ---- P4b: fine sweep, CTA = (fine pass, angle) -----------------------------------------------------
  int held_fid = -1;  // fine pass whose resolved descriptors sit in s_f / s_ft
  for (int v = blockIdx.x; v < A.nspec * A.nAf; v += gridDim.x) {
    const int fid = A.ncoarse + v / A.nAf, a = v % A.nAf;
    load_struct_cg(&s_f, passes + fid, tid);
    __syncthreads();
    load_struct_cg(&s_ft, tables + s_f.table, tid);
    __syncthreads();
    held_fid = fid;
    if (a < s_f.nA)
      fine_angle_body(g, pen, s_f, s_ft, fid, a, trig, pool, A.offsets, A.grids, A.resp, passmax, reinterpret_cast<int*>(dsm));
    __syncthreads();
  }
  YSM_TSTAMP(10)
  grid.sync();
  YSM_TSTAMP(11)
  // ---- P4c: per fine pass: reduce + angular covariance sums, publish --------------------------------
  for (int fi = blockIdx.x; fi < A.nspec; fi += gridDim.x) {
    const int fid = A.ncoarse + fi;
    if (held_fid != fid) {  // (the fine sweep of this CTA already fetched them when fi == 0)
      load_struct_cg(&s_f, passes + fid, tid);
      __syncthreads();
      load_struct_cg(&s_ft, tables + s_f.table, tid);
      __syncthreads();
    }
    reduce_body(g, s_f, s_ft, A.offsets, A.resp, __ldcg(passmax + fid), A.cellmax, trig, A.grids, &s_po[1], A.angs_host);
    __syncthreads();
    if (tid < nw8) reinterpret_cast<double*>(A.outs_host + fid)[tid] = reinterpret_cast<const double*>(&s_po[1])[tid];
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
      __threadfence_system();
      A.flags_host[fid] = A.epoch;
    }
    __syncthreads();
  }
  YSM_TSTAMP(12)
#undef YSM_TSTAMP
}

// ---------------------------------------------------------------------------------------------
// K5  trace_ray / run_raytracing_sweep (reference yag_slam/raytracing.py:63-92, numeric model
// SURVEY Appendix F): one thread per (start, angle) ray; float32 state advanced in float64,
// round-half-even cell index, uint8 map read through the read-only path.
// cs = per-angle (cos, sin) evaluated on the host with libm.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_raywalk(const uint8_t* __restrict__ img, int h, int w, const double* __restrict__ cs,
          int n_angles, const double* __restrict__ starts, int n_rays, float* __restrict__ out) {
  const int ray = blockIdx.x * blockDim.x + threadIdx.x;
  if (ray >= n_rays) return;
  const int s = ray / n_angles, a = ray - s * n_angles;
  const double c = cs[2 * a], sn = cs[2 * a + 1];
  const float spx = (float)starts[2 * s], spy = (float)starts[2 * s + 1];
  float x = spx, y = spy;
  bool run = true;
  int guard = 4 * (w + h) + 16;
  while (run && guard-- > 0) {
    int yi = __float2int_rn(y), xi = __float2int_rn(x);
    unsigned val = 0;
    if (xi >= 0 && xi < w && yi >= 0 && yi < h) val = __ldg(img + (size_t)yi * w + xi);
    if (val < 210u) run = false;
    x = __double2float_rn((double)x + c);
    y = __double2float_rn((double)y + sn);
    if (val > 180u && val < 210u && !run) {
      x = __double2float_rn((double)x + 1000.0 * c);
      y = __double2float_rn((double)y + 1000.0 * sn);
    }
    yi = __float2int_rn(y);
    xi = __float2int_rn(x);
    if (yi < 1 || xi < 1 || xi >= w - 1 || yi >= h - 1) run = false;
  }
  const float dx = x - spx, dy = y - spy;
  float* o = out + (size_t)ray * 5;
  o[0] = spx; o[1] = spy; o[2] = x; o[3] = y;
  o[4] = sqrtf(dx * dx + dy * dy);
}

}  // namespace ysm
