// ysm_resident.cuh -- KR, the resident single-query form of ScanMatcher::MatchScan (SURVEY.md A.5;
// reference call site yag_slam/scan_matching.py:40-42 -> Wrapper.match_scan, one synchronous call per
// scan from yag_slam/graph_slam.py:326).
//
// One cooperative kernel stays resident on the device (one CTA per SM) and serves match requests
// that the host posts through a 16-byte doorbell in mapped host memory; it leaves after
// `idle_ns` without a request (so device-wide synchronisations of other code never wait longer than
// that) and the host relaunches it on demand. Per request:
//   detect    CTAs 0..16 poll the doorbell (one 16-B PCIe read per poll)
//   A         CTA 1+s: FindValidPoints of base scan s, points pulled straight from host memory into
//             shared memory (fv logic of fv_scan_body); CTA 1+nbase copies the query points to HBM;
//             CTA 0 copies the control block (pass descriptors, search-angle cos/sin from host libm)
//   barrier 1
//   B         every CTA walks ALL occupied cells and keeps the (tile, cell) pairs of the tiles it owns
//             (tile -> CTA by hash), stamps them in registers (tile_scatter_rows) and writes each tile
//             once; meanwhile the lookup offsets of the CTA's search angle are already in shared memory
//   barrier 2
//   C         coarse CorrelateScan sweep, CTA = (angle, group of lattice rows), warps = row x point slice
//   arrive    CTA 0 waits for the sweep CTAs, then runs the whole tail alone: coarse max / ties / A.9
//             accumulators -> fine pass at the winner (offsets, 3 x 3 x nAf sweep, max / ties, angular
//             covariance sums) -> results as epoch-tagged 16-byte chunks in mapped host memory
//   barrier 3, then every CTA zeroes the tiles it stamped (the slot is all-zero between matches)
// Anything the tail cannot finish exactly (tied coarse winners before a fine pass, response 0 with
// response expansion, list overflows) is reported as RES_ST_FALLBACK and the host reruns the match
// through the general path.
//
// Barriers are monotone counters in HBM (red.release.gpu / ld.acquire.gpu by thread 0 of each CTA,
// __syncthreads around); the acquire invalidates the SM's L1 (CCTL.IVALL), which is what makes data
// other SMs rewrote since the previous request visible.
#pragma once
#include "ysm_kernels.cuh"

namespace ysm {

#define YSM_RES_THREADS 1024
#define YSM_RES_MAXBASE 64
#define YSM_RES_POLLERS 16   // CTAs 1..16 poll the doorbell too (they own the scans of phase A)
#define YSM_RES_MAXT 32      // tiles one CTA can own per request
#define YSM_RES_CAND 256     // stamps per owned tile
#define YSM_RES_TIECAP 256
#define YSM_RES_MAXNA 128
#define YSM_RES_CHUNKS 256   // 16-byte result chunks
#define YSM_RES_TS 24        // trace timestamps

enum { RES_CMD_NONE = 0, RES_CMD_MATCH = 1, RES_CMD_QUIT = 2, RES_CMD_PING = 3 };
enum { RES_ST_OK = 0, RES_ST_FALLBACK = 1, RES_ST_PONG = 2 };

// Control block of one request (mapped host memory -> HBM copy). Plain data, multiple of 16 bytes.
struct ResReq {
  unsigned seq, cmd;
  int nbase, Pq, pstride, do_refine;
  int nA, nAf, tpc, psplit, task_chunks, trace;
  // CTA 0's dynamic shared memory (bytes past the stamp table and the scratch every CTA has): query points
  // (double2) | spec tables | fine lookup offsets [nAf][Ppad] | fine sums (u32) + fine responses (f64)
  unsigned o_q, o_spec, o_foff, o_fsum;
  MatchDev m;
  PassDev coarse, fine;   // fine: everything but the search centre / trig rows (set by the tail)
  TableDev ctab, ftab;
  unsigned short counts[YSM_RES_MAXBASE + 12];  // point readings of base scan s; [nbase] = query
  double trig4[YSM_RES_MAXNA][4];              // per coarse angle: cos, sin, cos / sin of the normalised heading
};
static_assert(sizeof(ResReq) % 16 == 0, "ResReq must be a multiple of 16 bytes");
static_assert(offsetof(ResReq, trig4) % 16 == 0, "ResReq header must be a multiple of 16 bytes");

// Speculative fine tables (host libm, written while the GPU runs phases A-C): for every possible
// winning coarse angle the heading atan2 returns and the fine search angles around it.
struct ResSpecHdr {
  unsigned seq, pad[3];
};

struct ResArgs {
  // mapped host memory
  const uint4* db;              // doorbell: {seq, cmd | nbase << 8 | nA << 16 | nAf << 24, Pq | pstride << 16, ctl bytes}
  const unsigned char* req;     // ResReq
  const double* pts;            // scan s at pts + 2 * s * pstride ([nbase] = query)
  const unsigned char* spec;    // ResSpecHdr | heading[nA] | ftrig4[nA][nAf][4]
  uint4* out;                   // result chunks {payload lo, payload hi, seq, index}
  unsigned* exit_line;          // {last seq served, exit code}
  // HBM
  unsigned char* ctl;           // device copy of ResReq
  unsigned* bars;               // counters, 32 words apart: [0] barrier 1, [32] barrier 2, [64] sweep done, [96] barrier 3
  unsigned* quit_round;         // CTA 0 -> pollers: round number that ends the kernel
  int* abort_flag;
  uint32_t* cells;              // occupied cells of the match (any order)
  int* ncells;
  int cells_cap;
  double* qpts;                 // query point readings
  double* resp;                 // [iy][ix][a] coarse responses
  unsigned long long* cellmax;  // [iy][ix]
  double* passmax;
  int* fail;
  const uint16_t* stamp_tab;
  uint8_t* grid;                // slot 0
  unsigned last_seq;
  unsigned long long idle_ns;     // CTA 0 ends the kernel after this long without a request
  unsigned long long stall_ns;    // a barrier that waits longer than this aborts the kernel
  // workers' dynamic shared memory past the stamp table: [0, o_off) phase A / B / sweep scratch, then the
  // lookup offsets of the CTA's angle | lattice columns | rows
  unsigned o_off;
};

__device__ __forceinline__ unsigned long long res_timer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint4 res_ld_volatile_v4(const void* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void res_st_volatile_v4(void* p, uint4 v) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ unsigned res_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void res_red_release(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void res_st_release(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// thread 0 of a CTA: wait until *ctr has reached `target` (wrap-safe). false: aborted / timed out.
__device__ __forceinline__ bool res_wait(const unsigned* ctr, unsigned target, int* abort_flag, unsigned long long stall_ns) {
  unsigned long long t0 = 0ull;
  unsigned spins = 0u;
  while ((int)(res_ld_acquire(ctr) - target) < 0) {
    if ((++spins & 0x3FFu) == 0u) {
      const unsigned long long now = res_timer();
      if (t0 == 0ull) t0 = now;
      if (now - t0 > stall_ns) atomicExch(abort_flag, 1);
      if (res_ld_acquire(reinterpret_cast<const unsigned*>(abort_flag)) != 0u) return false;
    }
  }
  return true;
}

// all threads of the CTA; s_ok is a shared int. Returns false when the kernel must stop.
__device__ __forceinline__ bool res_barrier(unsigned* ctr, unsigned target, int* abort_flag, int* s_ok, unsigned long long stall_ns) {
  __syncthreads();
  if (threadIdx.x == 0) {
    res_red_release(ctr, 1u);
    *s_ok = res_wait(ctr, target, abort_flag, stall_ns) ? 1 : 0;
  }
  __syncthreads();
  return *s_ok != 0;
}

__host__ __device__ __forceinline__ size_t res_fv_smem(int pstride) {
  return ((size_t)pstride * (8 + 8 + 2 + 2 + 2 + 1) + 64 + 15) & ~(size_t)15;
}
__host__ __device__ __forceinline__ size_t res_stamp_smem() {
  return (size_t)YSM_RES_MAXT * YSM_RES_CAND * 4 + (size_t)(YSM_RES_THREADS / 32) * (YSM_TILE * YSM_TILE);
}

__device__ __forceinline__ unsigned res_tile_hash(int t) { return (unsigned)t * 0x9E3779B1u; }

// ---- phase A: FindValidPoints of one base scan, points pulled from mapped host memory ----
__device__ __forceinline__ void
res_filter_scan(const GridC& g, const ResArgs& A, const ResReq* hreq, int s, int pstride, unsigned char* scratch,
                int* s_misc) {
  const int tid = threadIdx.x, lane = tid & 31, T = blockDim.x;
  double* s_px = reinterpret_cast<double*>(scratch);
  double* s_py = s_px + pstride;
  unsigned short* s_next = reinterpret_cast<unsigned short*>(s_py + pstride);
  unsigned short* s_ja = s_next + pstride;
  unsigned short* s_jb = s_ja + pstride;
  unsigned char* s_mark = reinterpret_cast<unsigned char*>(s_jb + pstride);
  double* s_m = reinterpret_cast<double*>(s_misc);  // [4] vpx, vpy, gox, goy; s_misc[8] = n
  // header fields and points in one PCIe round trip (the point count is only known afterwards: pstride
  // points are fetched)
  if (tid == 0) {
    s_misc[8] = (int)__ldcv(&hreq->counts[s]);
  } else if (tid >= 32 && tid < 36) {
    const double* src = tid == 32 ? &hreq->m.vpx : tid == 33 ? &hreq->m.vpy : tid == 34 ? &hreq->m.gox : &hreq->m.goy;
    s_m[tid - 32] = __ldcv(src);
  }
  const double2* src = reinterpret_cast<const double2*>(A.pts + 2 * (size_t)s * pstride);
  for (int i = tid; i < pstride; i += T) {
    const double2 w = __ldcv(src + i);
    s_px[i] = w.x;
    s_py[i] = w.y;
    s_mark[i] = i == 0;
  }
  __syncthreads();
  const int n = min(s_misc[8], pstride);
  const double vpx = s_m[0], vpy = s_m[1], gox = s_m[2], goy = s_m[3];
  const double msd = 0.1 * 0.1;  // math::Square(0.1)
  for (int i = tid; i < n; i += T) {
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
  // the trigger chain 0 -> next[0] -> ... by pointer doubling (see fv_scan_body)
  unsigned short* ja = s_ja;
  unsigned short* jb = s_jb;
  for (int span = 1; span < n; span <<= 1) {
    for (int i = tid; i < n; i += T) {
      const int j = ja[i];
      if (j < n) {
        if (s_mark[i]) s_mark[j] = 1;  // benign race: marks only go 0 -> 1 (see fv_scan_body)
        jb[i] = ja[j];
      } else {
        jb[i] = (unsigned short)n;
      }
    }
    __syncthreads();
    unsigned short* t = ja; ja = jb; jb = t;
  }
  for (int j0 = 0; j0 < n; j0 += T) {
    const int j = j0 + tid;
    uint32_t cell = YSM_INVALID_CELL;
    if (j < n) {
      int f = j;
      while (!s_mark[f]) f--;
      const int c = s_next[f];
      if (c < n) {  // points after the last trigger are never emitted
        const double fx = s_px[f], fy = s_py[f];
        const double a = vpy - fy;
        const double b2 = fx - vpx;
        const double cc = fy * vpx - fx * vpy;
        const double ss = s_px[c] * a + s_py[c] * b2 + cc;
        if (!(ss < 0.0)) {
          const double vx = (s_px[j] - gox) * g.scale;
          const double vy = (s_py[j] - goy) * g.scale;
          if (vx > -1.0 && vy > -1.0 && vx < 1e9 && vy < 1e9) {
            const int gx = (int)kt_round(vx), gy = (int)kt_round(vy);
            if (gx >= 0 && gx < g.roi && gy >= 0 && gy < g.roi)
              cell = (uint32_t)(gx + g.border) | ((uint32_t)(gy + g.border) << 16);
          }
        }
      }
    }
    // the smear is a pure max: the cells may be listed in any order
    const unsigned bal = __ballot_sync(0xffffffffu, cell != YSM_INVALID_CELL);
    int base = 0;
    if (lane == 0 && bal) base = atomicAdd(A.ncells, __popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (cell != YSM_INVALID_CELL) {
      const int pos = base + __popc(bal & ((1u << lane) - 1u));
      if (pos < A.cells_cap) A.cells[pos] = cell;
      else *A.fail = 1;
    }
  }
}

// ---- phase B: the tiles this CTA owns ------------------------------------------------------------------
__device__ __forceinline__ void
res_collect(const GridC& g, const ResArgs& A, int total, int G, int bid, int* s_tile, int* s_cnt, uint32_t* s_steps) {
  const int tid = threadIdx.x, T = blockDim.x;
  const int tnx = (g.width + YSM_TILE - 1) / YSM_TILE;
  const int h = g.half_kernel;
  for (int i0 = 0; i0 < total; i0 += 4 * T) {
    uint32_t c[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int i = i0 + k * T + tid;
      c[k] = i < total ? __ldcg(A.cells + i) : YSM_INVALID_CELL;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (c[k] == YSM_INVALID_CELL) continue;
      const int ax = (int)(c[k] & 0xFFFFu), ay = (int)(c[k] >> 16);
      const int tx0 = (ax - h) / YSM_TILE, tx1 = (ax + h) / YSM_TILE;
      const int ty0 = (ay - h) / YSM_TILE, ty1 = (ay + h) / YSM_TILE;
      for (int ty = ty0; ty <= ty1; ty++)
        for (int tx = tx0; tx <= tx1; tx++) {
          const int t = ty * tnx + tx;
          const unsigned hsh = res_tile_hash(t);
          if ((int)((hsh >> 12) % (unsigned)G) != bid) continue;
          bool placed = false;
          for (int probe = 0; probe < YSM_RES_MAXT && !placed; probe++) {
            const int sl = (int)((hsh >> 27) + (unsigned)probe) & (YSM_RES_MAXT - 1);
            const int old = atomicCAS(&s_tile[sl], -1, t);
            if (old == -1 || old == t) {
              const int k2 = atomicAdd(&s_cnt[sl], 1);
              if (k2 < YSM_RES_CAND) s_steps[sl * YSM_RES_CAND + k2] = stamp_step(c[k], h, g.K, g.Wt, tx * YSM_TILE, ty * YSM_TILE);
              else *A.fail = 1;
              placed = true;
            }
          }
          if (!placed) *A.fail = 1;
        }
    }
  }
}

// stamps the owned tiles (up to 8 warps share a tile) and writes each once
__device__ __forceinline__ void
res_stamp(const GridC& g, const ResArgs& A, const int* s_tile, const int* s_cnt, const uint32_t* s_steps, int* s_slots,
          int* s_nt, uint32_t* s_stage, uint32_t lane_tab_s) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const int tnx = (g.width + YSM_TILE - 1) / YSM_TILE;
  if (warp == 0) {
    const bool used = s_tile[lane] != -1;  // YSM_RES_MAXT == 32
    const unsigned bal = __ballot_sync(0xffffffffu, used);
    if (used) s_slots[__popc(bal & ((1u << lane) - 1u))] = lane;
    if (lane == 0) *s_nt = __popc(bal);
  }
  __syncthreads();
  const int nt = *s_nt;
  if (nt == 0) return;  // CTA-uniform
  const int S = max(1, min(8, nwarps / nt));
  const int ti = warp / S, part = warp - ti * S;
  const bool active = ti < nt;
  int slot = 0;
  if (active) {
    slot = s_slots[ti];
    const int cnt = min(s_cnt[slot], YSM_RES_CAND);
    const int per = (cnt + S - 1) / S;
    const int lo = min(cnt, part * per), hi = min(cnt, lo + per);
    uint32_t t[16];
#pragma unroll
    for (int k = 0; k < 16; k++) t[k] = 0u;
    tile_scatter_rows(t, s_steps + slot * YSM_RES_CAND + lo, hi - lo, lane_tab_s, g.K);
    uint32_t b[8];
#pragma unroll
    for (int k = 0; k < 8; k++) b[k] = __byte_perm(t[2 * k], t[2 * k + 1], 0x6420);  // u16 lanes -> bytes
    const int sw = (lane >> 2) & 1;
    uint4* st4 = reinterpret_cast<uint4*>(s_stage + (size_t)warp * (YSM_TILE * YSM_TILE / 4));
    st4[lane * 2 + (0 ^ sw)] = make_uint4(b[0], b[1], b[2], b[3]);
    st4[lane * 2 + (1 ^ sw)] = make_uint4(b[4], b[5], b[6], b[7]);
  }
  __syncthreads();
  if (active) {
    const int tile = s_tile[slot];
    const int ty = tile / tnx, tx = tile - ty * tnx;
    const int x0t = tx * YSM_TILE, y0t = ty * YSM_TILE;
    uint32_t* gout = reinterpret_cast<uint32_t*>(A.grid);
    const int dr = lane >> 3, wd = lane & 7;
    const int gw = (x0t >> 2) + wd;
    const uint32_t* copies = s_stage + (size_t)(ti * S) * (YSM_TILE * YSM_TILE / 4);
    for (int k = part; k < 8; k += S) {
      const int r = k * 4 + dr, row = y0t + r;
      const int widx = r * 8 + 4 * ((wd >> 2) ^ (k & 1)) + (wd & 3);
      uint32_t v = copies[widx];
      for (int c = 1; c < S; c++) v = vmax4_lt128(v, copies[(size_t)c * (YSM_TILE * YSM_TILE / 4) + widx]);
      const unsigned nz = __ballot_sync(0xffffffffu, v != 0u);
      if ((nz & (0xFFu << (8 * dr))) && row < g.height && gw < g.stride4) gout[(size_t)row * g.stride4 + gw] = v;
    }
  }
}

// zero the tiles this CTA stamped
__device__ __forceinline__ void res_clear(const GridC& g, const ResArgs& A, const int* s_tile, const int* s_slots, int nt) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int tnx = (g.width + YSM_TILE - 1) / YSM_TILE;
  uint32_t* gout = reinterpret_cast<uint32_t*>(A.grid);
  for (int i = warp; i < nt * 8; i += nwarps) {
    const int tile = s_tile[s_slots[i >> 3]];
    const int ty = tile / tnx, tx = tile - ty * tnx;
    const int k = (i & 7) * 32 + lane;  // word of the tile: row k >> 3, word k & 7
    const int row = ty * YSM_TILE + (k >> 3), gw = ((tx * YSM_TILE) >> 2) + (k & 7);
    if (row < g.height && gw < g.stride4) gout[(size_t)row * g.stride4 + gw] = 0u;
  }
}

// ---- phase C: coarse sweep of one (angle, task chunk) ---------------------------------------------------
// prep: lookup offsets of angle a (ComputeOffsets fused), lattice columns / rows -> shared memory
__device__ __forceinline__ void
res_sweep_prep(const GridC& g, const ResReq& rq, const ResArgs& A, int a, int* s_i, int* s_minmax) {
  const PassDev& ps = rq.coarse;
  const TableDev& tb = rq.ctab;
  const int tid = threadIdx.x, T = blockDim.x;
  const int pc4 = (ps.P + 7) & ~7;
  int* s_off = s_i;
  int* s_col = s_i + pc4;
  int* s_row = s_col + ps.nX;
  if (tid == 0) {
    s_minmax[0] = 0x7fffffff; s_minmax[1] = (int)0x80000000;
    s_minmax[2] = 0x7fffffff; s_minmax[3] = (int)0x80000000;
    s_minmax[4] = 0x7fffffff; s_minmax[5] = (int)0x80000000;
  }
  __syncthreads();
  const double cosine = rq.trig4[a][0], sine = rq.trig4[a][1];
  int mn = 0x7fffffff, mx = (int)0x80000000;
  for (int p = tid; p < ps.P; p += T) {
    const double2 w = __ldcg(reinterpret_cast<const double2*>(A.qpts) + p);
    int gx, gy;
    offset_cell(tb, g.scale, w.x, w.y, cosine, sine, gx, gy);
    const int o = gx + gy * g.stride;
    s_off[p] = o;
    mn = min(mn, o);
    mx = max(mx, o);
  }
  int bmn = 0x7fffffff, bmx = (int)0x80000000, cmn = 0x7fffffff, cmx = (int)0x80000000;
  for (int i = tid; i < ps.nX; i += T) {
    const double x = -ps.offx + (double)i * ps.resx;
    const int c = world_to_grid1(ps.cx + x, ps.gox, g.scale) + g.border;
    s_col[i] = c;
    cmn = min(cmn, c);
    cmx = max(cmx, c);
  }
  for (int i = tid; i < ps.nY; i += T) {
    const double y = -ps.offy + (double)i * ps.resy;
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
    cmn = min(cmn, __shfl_xor_sync(0xffffffffu, cmn, o));
    cmx = max(cmx, __shfl_xor_sync(0xffffffffu, cmx, o));
  }
  if ((tid & 31) == 0) {
    if (mn != 0x7fffffff) { atomicMin(&s_minmax[0], mn); atomicMax(&s_minmax[1], mx); }
    if (bmn != 0x7fffffff) { atomicMin(&s_minmax[2], bmn); atomicMax(&s_minmax[3], bmx); }
    if (cmn != 0x7fffffff) { atomicMin(&s_minmax[4], cmn); atomicMax(&s_minmax[5], cmx); }
  }
  __syncthreads();
  const int minoff = s_minmax[0];
  for (int p = tid; p < ps.P; p += T) s_off[p] -= minoff;  // own entries only: non-negative offsets
  __syncthreads();
}

__device__ __forceinline__ void
res_sweep_run(const GridC& g, const PenaltyC& pen, const ResReq& rq, const ResArgs& A, int a, int chunk, const int* s_i,
              const int* s_minmax, unsigned* s_part, double* s_wmax) {
  const PassDev& ps = rq.coarse;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const int pc4 = (ps.P + 7) & ~7;
  const unsigned* s_off = reinterpret_cast<const unsigned*>(s_i);
  const int* s_col = s_i + pc4;
  const int* s_row = s_col + ps.nX;
  const int nxc = (ps.nX + 31) >> 5;
  const int ntasks = ps.nY * nxc;
  const int tpc = rq.tpc, psplit = rq.psplit;
  const int task0 = chunk * tpc, task1 = min(ntasks, task0 + tpc);
  const int minoff = s_minmax[0];
  const long long lo = (long long)s_minmax[0] + s_minmax[2] + s_minmax[4];
  const long long hi = (long long)s_minmax[1] + s_minmax[3] + s_minmax[5];
  const bool safe = lo >= 0 && hi < (long long)g.data_size;  // CTA-uniform: no lookup can leave the grid
  const unsigned dsz = (unsigned)g.data_size;
  const int np = ps.P;
  const int slice = warp % psplit, wtask = warp / psplit, ntw = nwarps / psplit;
  const int plen = (((np + psplit - 1) / psplit) + 7) & ~7;
  const int pb = min(np, slice * plen), pe = min(np, pb + plen);
  const int iters = (task1 - task0 + ntw - 1) / ntw;  // CTA-uniform
  double wmax = 0.0;
  for (int it = 0; it < iters; it++) {
    const int task = task0 + wtask + it * ntw;
    const bool tv = wtask < ntw && task < task1;
    const int iy = tv ? task / nxc : 0, xc = tv ? task - iy * nxc : 0;
    const int ix = (xc << 5) + lane;
    const bool active = tv && ix < ps.nX;
    unsigned sum = 0;
    if (tv) {
      const int base = s_row[iy] + s_col[active ? ix : 0] + minoff;
      const uint8_t* gp = A.grid + base;
      sum = safe ? sweep_row<false>(gp, s_off + pb, pe - pb, base, dsz) : sweep_row<true>(gp, s_off + pb, pe - pb, base, dsz);
    }
    if (psplit > 1) {
      s_part[warp * 32 + lane] = sum;
      __syncthreads();
      if (slice == 0 && tv)
        for (int c = 1; c < psplit; c++) sum += s_part[(warp + c) * 32 + lane];
      __syncthreads();
    }
    if (active && slice == 0) {
      const double rr = response_of(ps, pen, sum, ix, iy, a);
      A.resp[((size_t)iy * ps.nX + ix) * ps.nA + a] = rr;
      atomicMax(A.cellmax + (size_t)iy * ps.nX + ix, (unsigned long long)__double_as_longlong(rr));
      wmax = rr > wmax ? rr : wmax;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double t = __shfl_xor_sync(0xffffffffu, wmax, o);
    wmax = t > wmax ? t : wmax;
  }
  if (lane == 0) s_wmax[warp] = wmax;
  __syncthreads();
  if (tid == 0) {
    double m = s_wmax[0];
    for (int w = 1; w < nwarps; w++) m = s_wmax[w] > m ? s_wmax[w] : m;
    atomicMax(reinterpret_cast<unsigned long long*>(A.passmax), (unsigned long long)__double_as_longlong(m));
  }
}

// ---- tail (CTA 0) -----------------------------------------------------------------------------------------
struct ResTail {
  int list[YSM_RES_TIECAP];
  int sorted[YSM_RES_TIECAP];
  int count, n, first, status;
  double acc[4];
  double tmp4[32][4];
  PassOut po[2];
  unsigned long long fbest;
  int angs[YSM_RES_MAXNA];
};

// CorrelateScan epilogue of the coarse pass (reduce_body's logic for one CTA of any size): ties in
// storage order, sequential sums, A.9 accumulators. Returns false when the tie list overflows.
__device__ __forceinline__ bool
res_reduce_coarse(const ResReq& rq, const ResArgs& A, ResTail& S) {
  const PassDev& ps = rq.coarse;
  const int tid = threadIdx.x, T = blockDim.x;
  const double best = __longlong_as_double((long long)__ldcg(reinterpret_cast<const unsigned long long*>(A.passmax)));
  if (tid == 0) S.count = 0;
  __syncthreads();
  const int ncell = ps.nX * ps.nY;
  for (int c = tid; c < ncell; c += T) {
    const double m = __longlong_as_double((long long)__ldcg(A.cellmax + c));
    if (m >= best - YSM_KT_TOLERANCE) {
      const double* pc = A.resp + (size_t)c * ps.nA;
      for (int a0 = 0; a0 < ps.nA; a0 += 8) {
        double r[8];
#pragma unroll
        for (int u = 0; u < 8; u++) r[u] = a0 + u < ps.nA ? __ldcg(pc + a0 + u) : -1.0;
#pragma unroll
        for (int u = 0; u < 8; u++)
          if (r[u] >= 0.0 && kt_double_equal(r[u], best)) {
            const int pos = atomicAdd(&S.count, 1);
            if (pos < YSM_RES_TIECAP) S.list[pos] = c * ps.nA + a0 + u;
          }
      }
    }
  }
  __syncthreads();
  const int nt = S.count;
  if (nt > YSM_RES_TIECAP) return false;
  const double startX = -ps.offx, startY = -ps.offy;
  for (int e = tid; e < nt; e += T) {
    const int v = S.list[e];
    int rank = 0;
    for (int k = 0; k < nt; k++) rank += (S.list[k] < v);
    S.sorted[rank] = v;
  }
  __syncthreads();
  if (tid == 0) {
    double sx = 0.0, sy = 0.0, tx = 0.0, ty = 0.0;
    for (int e = 0; e < nt; e++) {
      const int idx = S.sorted[e];
      const int iy = idx / (ps.nX * ps.nA);
      const int rem = idx - iy * ps.nX * ps.nA;
      const int ix = rem / ps.nA, a = rem - ix * ps.nA;
      sx += ps.cx + (startX + (double)ix * ps.resx);
      sy += ps.cy + (startY + (double)iy * ps.resy);
      tx += rq.trig4[a][2];
      ty += rq.trig4[a][3];
    }
    const double cnt = (double)nt;
    PassOut& po = S.po[0];
    po.best = best;
    po.avg_x = nt > 0 ? sx / cnt : 0.0;
    po.avg_y = nt > 0 ? sy / cnt : 0.0;
    po.tx = nt > 0 ? tx / cnt : 0.0;
    po.ty = nt > 0 ? ty / cnt : 0.0;
    po.n_ties = nt;
    po.first_idx = nt > 0 ? S.sorted[0] : -1;
  }
  __syncthreads();
  // ComputePositionalCovariance accumulators (A.9); probs(x, y) = max response over the angles
  double norm = 0.0, axx = 0.0, axy = 0.0, ayy = 0.0;
  if (!(best < YSM_KT_TOLERANCE)) {
    const double dx = S.po[0].avg_x - ps.cx, dy = S.po[0].avg_y - ps.cy;
    for (int c = tid; c < ncell; c += T) {
      const int iy = c / ps.nX, ix = c - iy * ps.nX;
      const double pm = __longlong_as_double((long long)__ldcg(A.cellmax + c));
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
  for (int o = 16; o > 0; o >>= 1) {
    norm += __shfl_xor_sync(0xffffffffu, norm, o);
    axx += __shfl_xor_sync(0xffffffffu, axx, o);
    axy += __shfl_xor_sync(0xffffffffu, axy, o);
    ayy += __shfl_xor_sync(0xffffffffu, ayy, o);
  }
  if ((tid & 31) == 0) {
    S.tmp4[tid >> 5][0] = norm; S.tmp4[tid >> 5][1] = axx; S.tmp4[tid >> 5][2] = axy; S.tmp4[tid >> 5][3] = ayy;
  }
  __syncthreads();
  if (tid < 4) {
    double r = 0.0;
    for (int w = 0; w < (int)(T >> 5); w++) r += S.tmp4[w][tid];
    if (tid == 0) S.po[0].norm = r;
    else if (tid == 1) S.po[0].axx = r;
    else if (tid == 2) S.po[0].axy = r;
    else S.po[0].ayy = r;
  }
  __syncthreads();
  return true;
}

// fine CorrelateScan at the coarse winner: offsets, 3 x 3 x nAf sweep, max / ties, angular covariance sums.
// s_q: query points; s_ft: [nAf][4] cos / sin of the fine angles and of their normalised headings.
__device__ __forceinline__ bool
res_fine(const GridC& g, const PenaltyC& pen, const ResReq& rq, const ResArgs& A, ResTail& S, PassDev& f, const TableDev& ft,
         const double2* s_q, const double* s_ft, int* s_foff, unsigned* s_fsum, double* s_fr) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, T = blockDim.x, nwarps = T >> 5;
  const int nAf = f.nA, P = f.P, Ppad = f.Ppad;
  for (int it = tid; it < nAf * P; it += T) {
    const int a = it / P, p = it - a * P;
    int gx, gy;
    offset_cell(ft, g.scale, s_q[p].x, s_q[p].y, s_ft[4 * a], s_ft[4 * a + 1], gx, gy);
    s_foff[a * Ppad + p] = gx + gy * g.stride;
  }
  if (tid == 0) S.fbest = 0ull;
  __syncthreads();
  const int nxy = f.nX * f.nY, nposes = nxy * nAf;
  const unsigned dsz = (unsigned)g.data_size;
  for (int pose = warp; pose < nposes; pose += nwarps) {
    const int c = pose / nAf, a = pose - c * nAf;
    const int iy = c / f.nX, ix = c - iy * f.nX;
    const double x = -f.offx + (double)ix * f.resx;
    const double y = -f.offy + (double)iy * f.resy;
    const int gx = world_to_grid1(f.cx + x, f.gox, g.scale) + g.border;
    const int gy = world_to_grid1(f.cy + y, f.goy, g.scale) + g.border;
    const int base = gx + gy * g.stride;
    const int* off = s_foff + a * Ppad;
    unsigned sum = 0;
    for (int p0 = lane; p0 < P; p0 += 384) {
      unsigned idx[12];
#pragma unroll
      for (int u = 0; u < 12; u++) idx[u] = (p0 + 32 * u < P) ? (unsigned)(base + off[p0 + 32 * u]) : 0xFFFFFFFFu;
#pragma unroll
      for (int u = 0; u < 12; u++) if (idx[u] < dsz) sum += (unsigned)A.grid[idx[u]];
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) {
      const double rr = response_of(f, pen, sum, ix, iy, a);
      s_fsum[pose] = sum;
      s_fr[pose] = rr;  // storage order [iy][ix][a] == pose index
      atomicMax(&S.fbest, (unsigned long long)__double_as_longlong(rr));
    }
  }
  __syncthreads();
  const double best = __longlong_as_double((long long)S.fbest);
  // ties in storage order: ordered compaction warp by warp
  if (tid == 0) S.count = 0;
  __syncthreads();
  for (int c0 = 0; c0 < nposes; c0 += T) {
    const int i = c0 + tid;
    const bool tie = i < nposes && kt_double_equal(s_fr[i], best);
    const unsigned bal = __ballot_sync(0xffffffffu, tie);
    if (lane == 0) S.sorted[warp] = __popc(bal);  // (sorted[] is free here; T / 32 <= TIECAP)
    __syncthreads();
    int before = S.count;
    for (int w = 0; w < warp; w++) before += S.sorted[w];
    const int pos = before + __popc(bal & ((1u << lane) - 1u));
    if (tie && pos < YSM_RES_TIECAP) S.list[pos] = i;
    __syncthreads();
    if (tid == 0) {
      int tot = S.count;
      for (int w = 0; w < nwarps; w++) tot += S.sorted[w];
      S.count = tot;
    }
    __syncthreads();
  }
  const int nt = S.count;
  if (nt > YSM_RES_TIECAP) return false;
  if (tid == 0) {
    const double startX = -f.offx, startY = -f.offy;
    double sx = 0.0, sy = 0.0, tx = 0.0, ty = 0.0;
    for (int e = 0; e < nt; e++) {
      const int idx = S.list[e];
      const int iy = idx / (f.nX * nAf);
      const int rem = idx - iy * f.nX * nAf;
      const int ix = rem / nAf, a = rem - ix * nAf;
      sx += f.cx + (startX + (double)ix * f.resx);
      sy += f.cy + (startY + (double)iy * f.resy);
      tx += s_ft[4 * a + 2];
      ty += s_ft[4 * a + 3];
    }
    const double cnt = (double)nt;
    PassOut& po = S.po[1];
    po.best = best;
    po.avg_x = nt > 0 ? sx / cnt : 0.0;
    po.avg_y = nt > 0 ? sy / cnt : 0.0;
    po.tx = nt > 0 ? tx / cnt : 0.0;
    po.ty = nt > 0 ? ty / cnt : 0.0;
    po.norm = 0.0; po.axx = 0.0; po.axy = 0.0; po.ayy = 0.0;
    po.n_ties = nt;
    po.first_idx = nt > 0 ? S.list[0] : -1;
  }
  __syncthreads();
  // ComputeAngularCovariance sums: un-penalised GetResponse at the best cell for every fine angle. The
  // best cell is one of the 3 x 3 lattice cells whenever the mean rounds onto it: reuse those sums.
  if (nt > 0) {
    const int gx = world_to_grid1(S.po[1].avg_x, f.gox, g.scale) + g.border;
    const int gy = world_to_grid1(S.po[1].avg_y, f.goy, g.scale) + g.border;
    if (tid == 0) S.first = -1;
    __syncthreads();
    if (tid < nxy) {
      const int iy = tid / f.nX, ix = tid - iy * f.nX;
      const double x = -f.offx + (double)ix * f.resx;
      const double y = -f.offy + (double)iy * f.resy;
      const int cx = world_to_grid1(f.cx + x, f.gox, g.scale) + g.border;
      const int cy = world_to_grid1(f.cy + y, f.goy, g.scale) + g.border;
      if (cx == gx && cy == gy) atomicMax(&S.first, tid);
    }
    __syncthreads();
    const int cell = S.first;
    if (cell >= 0) {
      for (int a = tid; a < nAf; a += T) S.angs[a] = (int)s_fsum[cell * nAf + a];
    } else {
      const int base = gx + gy * g.stride;
      for (int a = warp; a < nAf; a += nwarps) {
        const int* off = s_foff + a * Ppad;
        unsigned sum = 0;
        for (int p = lane; p < P; p += 32) {
          const unsigned idx = (unsigned)(base + off[p]);
          if (idx < dsz) sum += (unsigned)A.grid[idx];
        }
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) S.angs[a] = (int)sum;
      }
    }
  } else {
    for (int a = tid; a < nAf; a += T) S.angs[a] = 0;
  }
  __syncthreads();
  return true;
}

// results -> mapped host memory as 16-byte chunks {payload, seq, index}: every chunk is valid on its own, so
// no fence and no completion flag are needed
__device__ __forceinline__ void res_publish(const ResArgs& A, const ResTail& S, unsigned seq, int status, int has_fine,
                                            int nAf, const unsigned long long* ts) {
  const int tid = threadIdx.x;
  const int np = (int)(sizeof(PassOut) / 8);
  const int nang = has_fine ? (nAf + 1) / 2 : 0;
  const int nts = ts ? YSM_RES_TS : 0;
  const int total = 1 + 2 * np + nang + nts;
  if (tid < total && tid < YSM_RES_CHUNKS) {
    unsigned long long v;
    if (tid == 0) v = (unsigned long long)(unsigned)status | ((unsigned long long)(unsigned)has_fine << 8) | ((unsigned long long)(unsigned)total << 32);
    else if (tid < 1 + np) v = reinterpret_cast<const unsigned long long*>(&S.po[0])[tid - 1];
    else if (tid < 1 + 2 * np) v = reinterpret_cast<const unsigned long long*>(&S.po[1])[tid - 1 - np];
    else if (tid < 1 + 2 * np + nang) {
      const int k = 2 * (tid - 1 - 2 * np);
      v = (unsigned long long)(unsigned)S.angs[k] | ((unsigned long long)(unsigned)(k + 1 < nAf ? S.angs[k + 1] : 0) << 32);
    } else v = ts[tid - 1 - 2 * np - nang];
    res_st_volatile_v4(A.out + tid, make_uint4((unsigned)v, (unsigned)(v >> 32), seq, (unsigned)tid));
  }
}

__global__ void __launch_bounds__(YSM_RES_THREADS, 1)
k_match_resident(GridC g, PenaltyC pen, ResArgs A) {
  extern __shared__ __align__(16) unsigned char dsm[];
  __shared__ __align__(16) ResReq s_rq;
  __shared__ ResTail s_tail;
  __shared__ int s_tile[YSM_RES_MAXT], s_cnt[YSM_RES_MAXT], s_slots[YSM_RES_MAXT];
  __shared__ int s_nt, s_ok, s_cmd;
  __shared__ uint4 s_db;
  __shared__ int s_minmax[8];
  __shared__ __align__(8) int s_misc[16];
  __shared__ double s_wmax[32];
  __shared__ unsigned long long s_ts[YSM_RES_TS];
  const int tid = threadIdx.x, lane = tid & 31, T = blockDim.x;
  const int G = (int)gridDim.x, bid = (int)blockIdx.x;
  // the stamp table stays in shared memory for the kernel's lifetime
  const int ntab4 = (int)(stamp_table_bytes(g.K, g.Wt) / 16);
  {
    uint4* s_tab4 = reinterpret_cast<uint4*>(dsm);
    for (int t = tid; t < ntab4; t += T) s_tab4[t] = __ldg(reinterpret_cast<const uint4*>(A.stamp_tab) + t);
  }
  unsigned char* dyn = dsm + (size_t)ntab4 * 16;
  const uint32_t lane_tab_s = (uint32_t)__cvta_generic_to_shared(dsm) + (uint32_t)(lane * g.Wt * 2);
  const bool poller = bid <= YSM_RES_POLLERS;
  const ResReq* hreq = reinterpret_cast<const ResReq*>(A.req);
  unsigned seq_done = A.last_seq;
  unsigned round = 0u, t1 = 0u, t2 = 0u, t3 = 0u, t4 = 0u;
  unsigned long long idle0 = res_timer();
  int exit_code = 0;
  if (tid < YSM_RES_MAXT) { s_tile[tid] = -1; s_cnt[tid] = 0; }
  __syncthreads();
#define RES_TS(k) if (bid == 0 && tid == 0) s_ts[k] = res_timer();
  for (;;) {
    round++;
    // ---- wait for the doorbell -----------------------------------------------------------------------
    int cmd = RES_CMD_NONE;
    if (poller) {
      if (tid == 0) {
        uint4 d = make_uint4(0u, 0u, 0u, 0u);
        int c = RES_CMD_NONE;
        for (unsigned spins = 1u;; spins++) {
          d = res_ld_volatile_v4(A.db);
          if (d.x != seq_done) { c = (int)(d.y & 0xFFu); break; }
          if (bid == 0) {
            if ((spins & 7u) == 0u && res_timer() - idle0 > A.idle_ns) {
              c = RES_CMD_QUIT;
              res_st_release(A.quit_round, round);
              break;
            }
          } else if ((spins & 3u) == 0u && res_ld_acquire(A.quit_round) == round) {
            c = RES_CMD_QUIT;
            break;
          }
        }
        s_db = d;
        s_cmd = c;
      }
      __syncthreads();
      cmd = s_cmd;
    }
    if (bid == 0 && tid == 0) s_ts[0] = res_timer();
    if (poller && cmd == RES_CMD_MATCH) {
      const uint4 d = s_db;
      const int nbase = (int)((d.y >> 8) & 0xFFu), pstride = (int)(d.z >> 16);
      if (bid == 0) {
        // control block -> HBM (every CTA reads it after barrier 1)
        const int nvec = (int)(d.w / 16u);
        const uint4* src = reinterpret_cast<const uint4*>(A.req);
        uint4* dst = reinterpret_cast<uint4*>(A.ctl);
        for (int i = tid; i < nvec; i += T) dst[i] = __ldcv(src + i);
      } else {
        for (int s = bid - 1; s <= nbase; s += YSM_RES_POLLERS) {
          if (s < nbase) {
            res_filter_scan(g, A, hreq, s, pstride, dyn, s_misc);
          } else {
            const int Pq = (int)(d.z & 0xFFFFu);
            const double2* src = reinterpret_cast<const double2*>(A.pts + 2 * (size_t)s * pstride);
            double2* dst = reinterpret_cast<double2*>(A.qpts);
            for (int i = tid; i < Pq; i += T) dst[i] = __ldcv(src + i);
          }
          __syncthreads();
        }
      }
    } else if (bid == 0) {
      // quit / ping / idle: a 32-byte control block made here
      if (tid == 0) {
        unsigned* c = reinterpret_cast<unsigned*>(A.ctl);
        c[0] = s_db.x;
        c[1] = (unsigned)cmd;
      }
      if (cmd == RES_CMD_PING && tid == 0)
        res_st_volatile_v4(A.out, make_uint4((unsigned)RES_ST_PONG, 1u, s_db.x, 0u));
    }
    RES_TS(1)
    // ---- barrier 1: cells, query points and control block are in HBM ----------------------------
    t1 += (unsigned)G;
    if (!res_barrier(A.bars, t1, A.abort_flag, &s_ok, A.stall_ns)) { exit_code = 2; break; }
    RES_TS(2)
    {
      const unsigned* c = reinterpret_cast<const unsigned*>(A.ctl);
      if (tid == 0) { s_misc[0] = (int)__ldcg(c); s_misc[1] = (int)__ldcg(c + 1); }
      __syncthreads();
    }
    const unsigned seq = (unsigned)s_misc[0];
    cmd = s_misc[1];
    __syncthreads();
    if (cmd == RES_CMD_QUIT || cmd == RES_CMD_NONE) break;
    if (cmd == RES_CMD_PING) {
      seq_done = seq;
      idle0 = res_timer();
      continue;
    }
    // the request's control block -> shared memory (header, descriptors, the nA trig rows)
    {
      const int hdr4 = (int)(offsetof(ResReq, trig4) / 16);
      const uint4* src = reinterpret_cast<const uint4*>(A.ctl);
      uint4* dst = reinterpret_cast<uint4*>(&s_rq);
      for (int i = tid; i < hdr4; i += T) dst[i] = __ldcg(src + i);
      __syncthreads();
      const int n4 = hdr4 + 2 * s_rq.nA;
      for (int i = hdr4 + tid; i < n4; i += T) dst[i] = __ldcg(src + i);
      __syncthreads();
    }
    const ResReq& rq = s_rq;
    const PassDev& ps = rq.coarse;
    const int nv = rq.nA * rq.task_chunks;
    int v = bid - 1;
    const bool worker = bid >= 1;
    // lookup offsets of this CTA's first angle (they do not depend on the grid): before the stamping
    if (worker && v < nv) res_sweep_prep(g, rq, A, v % rq.nA, reinterpret_cast<int*>(dyn + A.o_off), s_minmax);
    if (bid == 0) {
      // the tail needs the query points in shared memory
      double2* s_q = reinterpret_cast<double2*>(dyn + A.o_off + rq.o_q);
      for (int i = tid; i < ps.P; i += T) s_q[i] = __ldcg(reinterpret_cast<const double2*>(A.qpts) + i);
    }
    // ---- phase B: stamp the tiles this CTA owns -------------------------------------------------------
    {
      const int total = min(__ldcg(A.ncells), A.cells_cap);
      uint32_t* s_steps = reinterpret_cast<uint32_t*>(dyn);
      uint32_t* s_stage = s_steps + YSM_RES_MAXT * YSM_RES_CAND;
      res_collect(g, A, total, G, bid, s_tile, s_cnt, s_steps);
      __syncthreads();
      res_stamp(g, A, s_tile, s_cnt, s_steps, s_slots, &s_nt, s_stage, lane_tab_s);
    }
    RES_TS(3)
    t2 += (unsigned)G;
    if (!res_barrier(A.bars + 32, t2, A.abort_flag, &s_ok, A.stall_ns)) { exit_code = 2; break; }
    RES_TS(4)
    const bool failed = __ldcg(A.fail) != 0;
    // ---- phase C: coarse sweep ------------------------------------------------------------------------------
    if (worker && !failed) {
      bool first = true;
      for (; v < nv; v += G - 1) {
        if (!first) {
          __syncthreads();
          res_sweep_prep(g, rq, A, v % rq.nA, reinterpret_cast<int*>(dyn + A.o_off), s_minmax);
        }
        first = false;
        res_sweep_run(g, pen, rq, A, v % rq.nA, v / rq.nA, reinterpret_cast<const int*>(dyn + A.o_off), s_minmax,
                      reinterpret_cast<unsigned*>(dyn), s_wmax);
      }
    }
    if (worker) {
      __syncthreads();
      if (tid == 0) res_red_release(A.bars + 64, 1u);
    }
    if (bid == 0) {
      // spec tables: the host writes them while the GPU runs phases A-C
      int status = failed ? RES_ST_FALLBACK : RES_ST_OK;
      int has_fine = 0;
      double* s_spec = reinterpret_cast<double*>(dyn + A.o_off + rq.o_spec);
      const int spec_doubles = rq.nA + 4 * rq.nA * rq.nAf;
      if (rq.do_refine && !failed) {
        if (tid == 0) {
          const unsigned long long t0 = res_timer();
          unsigned spins = 0u;
          int ok = 1;
          while (res_ld_volatile_v4(A.spec).x != seq) {
            if ((++spins & 0xFFu) == 0u && res_timer() - t0 > 2000000000ull) { ok = 0; break; }
          }
          s_ok = ok;
        }
        __syncthreads();
        if (!s_ok) status = RES_ST_FALLBACK;
        else {
          const double2* src = reinterpret_cast<const double2*>(A.spec + sizeof(ResSpecHdr));
          double2* dst = reinterpret_cast<double2*>(s_spec);
          for (int i = tid; i < (spec_doubles + 1) / 2; i += T) dst[i] = __ldcv(src + i);
        }
      }
      RES_TS(5)
      t3 += (unsigned)(G - 1);
      if (tid == 0) s_ok = res_wait(A.bars + 64, t3, A.abort_flag, A.stall_ns) ? 1 : 0;
      __syncthreads();
      if (!s_ok) { exit_code = 2; break; }
      RES_TS(6)
      if (status == RES_ST_OK) {
        if (!res_reduce_coarse(rq, A, s_tail)) status = RES_ST_FALLBACK;
      }
      RES_TS(7)
      if (status == RES_ST_OK && rq.do_refine) {
        const PassOut& po = s_tail.po[0];
        if (po.n_ties == 1 && po.best > YSM_KT_TOLERANCE) {
          // MatchScan goes straight to the fine pass: its centre is the winning lattice pose, the heading
          // the atan2(sin, cos) the host tabulated for that coarse angle
          const int a = po.first_idx % ps.nA;
          PassDev& f = s_rq.fine;
          if (tid == 0) {
            f.cx = po.avg_x;
            f.cy = po.avg_y;
            f.ch = s_spec[a];
          }
          __syncthreads();
          const double* s_ft = s_spec + rq.nA + (size_t)4 * a * rq.nAf;
          unsigned* s_fsum = reinterpret_cast<unsigned*>(dyn + A.o_off + rq.o_fsum);
          double* s_fr = reinterpret_cast<double*>(s_fsum + ((f.nX * f.nY * f.nA + 1) & ~1));
          if (res_fine(g, pen, rq, A, s_tail, f, rq.ftab, reinterpret_cast<const double2*>(dyn + A.o_off + rq.o_q), s_ft,
                       reinterpret_cast<int*>(dyn + A.o_off + rq.o_foff), s_fsum, s_fr))
            has_fine = 1;
          else
            status = RES_ST_FALLBACK;
        } else {
          status = RES_ST_FALLBACK;  // tied winners / response 0: the host reschedules (expansion, ties)
        }
      }
      RES_TS(8)
      __syncthreads();
      res_publish(A, s_tail, seq, status, has_fine, rq.nAf, rq.trace ? s_ts : nullptr);
      // reset the per-request accumulators for the next request
      const int ncell = ps.nX * ps.nY;
      for (int i = tid; i < ncell; i += T) A.cellmax[i] = 0ull;
      if (tid == 0) {
        *A.passmax = 0.0;
        *A.ncells = 0;
        *A.fail = 0;
      }
    }
    // ---- barrier 3: the tail has read the grid; zero the tiles this CTA stamped ---------------------
    t4 += (unsigned)G;
    if (!res_barrier(A.bars + 96, t4, A.abort_flag, &s_ok, A.stall_ns)) { exit_code = 2; break; }
    res_clear(g, A, s_tile, s_slots, s_nt);
    __syncthreads();
    if (tid < YSM_RES_MAXT) { s_tile[tid] = -1; s_cnt[tid] = 0; }
    if (tid == 0) s_nt = 0;
    __syncthreads();
    seq_done = seq;
    idle0 = res_timer();
  }
#undef RES_TS
  // leave: CTA 0 tells the host which request was the last one served
  if (bid == 0 && tid == 0) {
    __threadfence_system();
    res_st_volatile_v4(A.exit_line, make_uint4(seq_done, (unsigned)exit_code, 0x45584954u, round));
  }
}

}  // namespace ysm
