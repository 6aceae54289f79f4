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

#ifndef YSM_RES_THREADS
#define YSM_RES_THREADS 1024
#endif
#define YSM_RES_MAXBASE 64
#define YSM_RES_POLLERS 16   // CTAs 1..16 poll the doorbell too (they own the scans of phase A)
#define YSM_RES_MAXT 32      // tiles one CTA can own per request
#define YSM_RES_CAND 256     // stamps per owned tile
#define YSM_RES_TIECAP 256
#define YSM_RES_MAXNA 128
#define YSM_RES_CHUNKS 256   // 16-byte result chunks
#define YSM_RES_TS 24        // trace timestamps
#define YSM_RES_DB_STRIDE 256
#define YSM_RES_PROF 24      // per-CTA trace timestamps
#define YSM_RES_PMAX 4096    // point readings per scan
#define YSM_RES_CACHE_SLOTS 32
#define YSM_RES_CM_SMEM 4096 // coarse lattice cells whose maxima the tail keeps in shared memory

enum { RES_CMD_NONE = 0, RES_CMD_MATCH = 1, RES_CMD_QUIT = 2, RES_CMD_PING = 3 };
enum { RES_ST_OK = 0, RES_ST_FALLBACK = 1, RES_ST_PONG = 2 };

// Control block of one request (mapped host memory -> HBM copy). Plain data, multiple of 16 bytes.
struct ResReq {
  unsigned seq, cmd;
  int nbase, Pq, pstride, do_refine;
  int nA, nAf, tpc, psplit, task_chunks, trace;
  int maxt, cand, tail_warps, pad1;  // tiles a CTA can own x stamps per tile (maxt * cand = YSM_RES_MAXT * YSM_RES_CAND)
  // CTA 0's dynamic shared memory (bytes past the stamp table and the scratch every CTA has): point stash |
  // spec tables | fine sums (u32) + fine responses (f64)   (o_foff: unused)
  unsigned o_q, o_spec, o_foff, o_fsum;
  unsigned o_cm, pad2[3];          // ... | per-cell maxima of the coarse pass (f64, up to YSM_RES_CM_SMEM cells)
  MatchDev m;
  PassDev coarse, fine;   // fine: everything but the search centre / trig rows (set by the tail)
  TableDev ctab, ftab;
  unsigned short counts[YSM_RES_MAXBASE + 12];  // point readings of base scan s; [nbase] = query
  double ap[YSM_RES_MAXNA];                     // angle penalty of every coarse angle (host: plain IEEE arithmetic)
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
  const uint4* db;              // doorbell lines, one per polling CTA, YSM_RES_DB_STRIDE bytes apart (concurrent reads of
                                // ONE host address are served one PCIe round trip after the other). Five tagged 16-byte
                                // words, valid together when all carry the same new seq:
                                //   v0 {seq, cmd | nbase << 8 | nA << 16 | nAf << 24, Pq | pstride << 16, ctl bytes}
                                //   v1 {seq, points of this CTA's scan, source cache slot + 1 (0: mailbox), slot to fill + 1}
                                //   v2..v4 {w, w, w, seq}: viewpoint x, y and grid offset x, y as eight 32-bit halves
  const unsigned char* req;     // ResReq
  const double* pts;            // scan s at pts + 2 * s * pstride ([nbase] = query)
  const unsigned char* spec;    // ResSpecHdr | heading[nA] | ftrig4[nA][nAf][4]
  uint4* out;                   // result chunks {payload lo, payload hi, seq, index}
  unsigned long long* prof;     // [gridDim][YSM_RES_PROF] per-CTA phase timestamps of traced requests
  unsigned* exit_line;          // {last seq served, exit code}
  // HBM
  unsigned char* ctl;           // device copy of ResReq
  unsigned long long* bars;     // counters, YSM_RES_BAR_STRIDE words apart: barrier 1, barrier 2, sweep done, fine done, barrier 3
  unsigned long long* win;      // CTA 0 -> workers: the coarse winner {seq, angle | ix << 8 | iy << 20} (or "no fine pass")
  unsigned* fsum;               // fine lookup sums [iy][ix][a], accumulated by the workers
  double* spec_dev;             // device copy of the spec tables (CTA 0 -> workers)
  unsigned* quit_round;         // CTA 0 -> pollers: round number that ends the kernel
  int* abort_flag;
  uint32_t* cells;              // [nbase][pstride] cell of every base point reading (YSM_INVALID_CELL: dropped)
  int cells_cap;
  double* qpts;                 // query point readings
  double* cache;                // device-resident scan store: YSM_RES_CACHE_SLOTS x YSM_RES_PMAX points
  double* resp;                 // [iy][ix][a] coarse responses
  unsigned long long* cellmax;  // [iy][ix]
  const double* dp_coarse;      // distance penalty of every coarse lattice cell [nY][nX] (depends on the configuration only)
  const double* dp_fine;        // ... of the fine pass's 3 x 3 cells
  int* winrec;                  // winner record for the fine items: flat index of the 3 x 3 cells | (cos, sin) rows as doubles at +16 ints
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
__device__ __forceinline__ unsigned long long res_ld_acquire64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void res_red_release64(unsigned long long* p, unsigned long long v) {
  asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void res_st_release(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void res_st_release64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// A barrier is a monotone 64-bit counter in HBM: the low word counts arrivals, the high word the arrivals
// that report a failure (so the verdict of a phase travels with the barrier: no extra round trip).
#define YSM_RES_BAR_STRIDE 16  // u64 words between counters (128 bytes)
__device__ __forceinline__ void res_arrive(unsigned long long* ctr, bool fail) {
  res_red_release64(ctr, 1ull + (fail ? (1ull << 32) : 0ull));
}
// thread 0 of a CTA: wait until the arrivals have reached `target` (wrap-safe); *hi receives the failure
// count. false: aborted / stalled.
__device__ __forceinline__ bool res_wait(const unsigned long long* ctr, unsigned target, unsigned* hi, int* abort_flag,
                                         unsigned long long stall_ns) {
  unsigned long long t0 = 0ull;
  unsigned spins = 0u;
  for (;;) {
    const unsigned long long v = res_ld_acquire64(ctr);
    if ((int)((unsigned)v - target) >= 0) {
      *hi = (unsigned)(v >> 32);
      return true;
    }
    if ((++spins & 0x3FFu) == 0u) {
      const unsigned long long now = res_timer();
      if (t0 == 0ull) t0 = now;
      if (now - t0 > stall_ns) atomicExch(abort_flag, 1);
      if (res_ld_acquire(reinterpret_cast<const unsigned*>(abort_flag)) != 0u) return false;
    }
  }
}

// all threads of the CTA; s_io is a shared int[2]: [0] = ok, [1] = failure count seen. `fail` is read from
// thread 0. Returns false when the kernel must stop.
__device__ __forceinline__ bool res_barrier(unsigned long long* ctr, unsigned target, bool fail, int* abort_flag, int* s_io,
                                            unsigned long long stall_ns) {
  __syncthreads();
  if (threadIdx.x == 0) {
    res_arrive(ctr, fail);
    unsigned hi = 0u;
    s_io[0] = res_wait(ctr, target, &hi, abort_flag, stall_ns) ? 1 : 0;
    s_io[1] = (int)hi;
  }
  __syncthreads();
  return s_io[0] != 0;
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
res_filter_scan(const GridC& g, const ResArgs& A, int s, int pstride, int n_in, int src_slot, int store_slot,
                const double* mf, unsigned char* scratch, unsigned long long* pf) {
  const int tid = threadIdx.x, T = blockDim.x;
  double* s_px = reinterpret_cast<double*>(scratch);
  double* s_py = s_px + pstride;
  unsigned short* s_next = reinterpret_cast<unsigned short*>(s_py + pstride);
  unsigned short* s_ja = s_next + pstride;
  unsigned short* s_jb = s_ja + pstride;
  unsigned char* s_mark = reinterpret_cast<unsigned char*>(s_jb + pstride);
  const int n = min(n_in, pstride);
  // the scan's point readings: from the device-resident scan store when the host found them there, else
  // straight from mapped host memory (one PCIe round trip)
  if (src_slot > 0) {
    const double2* src = reinterpret_cast<const double2*>(A.cache) + (size_t)(src_slot - 1) * YSM_RES_PMAX;
    for (int i = tid; i < n; i += T) {
      const double2 w = __ldcg(src + i);
      s_px[i] = w.x;
      s_py[i] = w.y;
      s_mark[i] = 0;
    }
  } else {
    const double2* src = reinterpret_cast<const double2*>(A.pts + 2 * (size_t)s * pstride);
    for (int i = tid; i < n; i += T) {
      const double2 w = __ldcv(src + i);
      s_px[i] = w.x;
      s_py[i] = w.y;
      s_mark[i] = 0;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) pf[9] = res_timer();
  const double vpx = mf[0], vpy = mf[1], gox = mf[2], goy = mf[3];
  const double msd = 0.1 * 0.1;  // math::Square(0.1)
  // next[i] = first later point farther than 10 cm from point i (four candidates per step: the loads of a
  // step are independent)
  for (int i = tid; i < n; i += T) {
    const double fx = s_px[i], fy = s_py[i];
    int j = i + 1;
    for (;;) {
      if (j >= n) { j = n; break; }
      const int j1 = min(j + 1, n - 1), j2 = min(j + 2, n - 1), j3 = min(j + 3, n - 1);
      const double x0 = fx - s_px[j], y0 = fy - s_py[j], x1 = fx - s_px[j1], y1 = fy - s_py[j1];
      const double x2 = fx - s_px[j2], y2 = fy - s_py[j2], x3 = fx - s_px[j3], y3 = fy - s_py[j3];
      if (x0 * x0 + y0 * y0 > msd) break;
      if (j + 1 >= n) { j = n; break; }
      if (x1 * x1 + y1 * y1 > msd) { j += 1; break; }
      if (j + 2 >= n) { j = n; break; }
      if (x2 * x2 + y2 * y2 > msd) { j += 2; break; }
      if (j + 3 >= n) { j = n; break; }
      if (x3 * x3 + y3 * y3 > msd) { j += 3; break; }
      j += 4;
    }
    s_next[i] = (unsigned short)j;
  }
  __syncthreads();
  if (threadIdx.x == 0) pf[10] = res_timer();
  // The trigger chain 0 -> next[0] -> next[next[0]] ... (FindValidPoints' firstPoint sequence), block-wise
  // over blocks of 32 points (= warps): (1) in-warp pointer doubling (5 shuffle rounds) tells every point
  // where the chain LEAVES the block if it enters there; (2) one thread hops from block to block through those
  // exits, which gives every block its entry point; (3) in-warp doubling again spreads the mark from the
  // entry over the chain's points inside the block.
  const int nblk = (n + 31) >> 5;
  const int lane_ = tid & 31;
  for (int b = tid >> 5; b < nblk; b += T >> 5) {
    const int i = b * 32 + lane_, bstart = b * 32, bend = min(n, bstart + 32);
    int cur = i < n ? (int)s_next[i] : n;
#pragma unroll
    for (int r = 0; r < 5; r++) {
      const int src = cur < bend ? cur - bstart : lane_;
      const int nxt = __shfl_sync(0xffffffffu, cur, src);
      if (cur < bend) cur = nxt;
    }
    if (i < n) s_ja[i] = (unsigned short)cur;
  }
  __syncthreads();
  if (tid == 0) {
    int e = 0;
    for (int b = 0; b < nblk; b++) {
      if (e < n && e < (b + 1) * 32) {
        s_jb[b] = (unsigned short)e;
        e = s_ja[e];
      } else {
        s_jb[b] = 0xFFFFu;  // the chain jumps over this block
      }
    }
  }
  __syncthreads();
  for (int b = tid >> 5; b < nblk; b += T >> 5) {
    const int i = b * 32 + lane_, bstart = b * 32, bend = min(n, bstart + 32);
    const int entry = s_jb[b];
    int jmp = i < n ? (int)s_next[i] : n;  // 2^r-th successor (absorbing outside the block)
    bool mk = i == entry;
    if (mk) s_mark[i] = 1;
#pragma unroll
    for (int r = 0; r < 5; r++) {
      if (mk && jmp < bend) s_mark[jmp] = 1;
      __syncwarp();
      if (i < n) mk = mk || s_mark[i] != 0;
      const int src = jmp < bend ? jmp - bstart : lane_;
      const int nxt = __shfl_sync(0xffffffffu, jmp, src);
      if (jmp < bend) jmp = nxt;
      __syncwarp();
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) pf[11] = res_timer();
  // one entry per point reading (YSM_INVALID_CELL: dropped); the smear is a pure max, so phase B may take the
  // cells in any order and no compaction is needed
  uint32_t* out = A.cells + (size_t)s * pstride;
  for (int j = tid; j < pstride; j += T) {
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
    out[j] = cell;
  }
  if (store_slot > 0) {  // keep the readings for the next requests (the host marks the slot valid afterwards)
    double2* dst = reinterpret_cast<double2*>(A.cache) + (size_t)(store_slot - 1) * YSM_RES_PMAX;
    for (int i = tid; i < n; i += T) dst[i] = make_double2(s_px[i], s_py[i]);
  }
}

// ---- phase B: the tiles this CTA owns ------------------------------------------------------------------
// Every CTA walks all cells and keeps the (tile, cell) pairs of the tiles it owns in a small shared-memory hash
// table. Cells arrive in scan order, so the lanes of a warp mostly name the SAME tile: one lane per (warp, tile)
// finds the slot and bumps its counter for the whole group (MATCH.ANY), the others only store their step.
// (r02zp, cfg-2 shape: with one atomicCAS + atomicAdd per pair the CTA owning the densest tiles -- a wall corner
// all ten running scans see -- spent 18 us here, serialised on four shared-memory words, and 147 CTAs waited for
// it at barrier 2.)
__device__ __forceinline__ void
res_collect(const GridC& g, const ResArgs& A, int total, int G, int bid, int maxt, int cand, int* s_tile, int* s_cnt,
            uint32_t* s_steps, int* s_fail, const uint32_t (&c_first)[4]) {
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31;
  const unsigned ltmask = (1u << lane) - 1u;
  const int tnx = (g.width + YSM_TILE - 1) / YSM_TILE;
  const int h = g.half_kernel;
  const bool two = 2 * h <= YSM_TILE;  // a stamp spans at most 2 x 2 tiles
  for (int i0 = 0; i0 < total; i0 += 4 * T) {
    uint32_t c[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int i = i0 + k * T + tid;
      c[k] = i >= total ? YSM_INVALID_CELL : (i0 == 0 ? c_first[k] : __ldcg(A.cells + i));  // (first round: preloaded)
    }
    if (two) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const bool cv = c[k] != YSM_INVALID_CELL;
        if (!__any_sync(0xffffffffu, cv)) continue;  // (warps past the end of the cell list: nothing to look at)
        const int ax = (int)(c[k] & 0xFFFFu), ay = (int)(c[k] >> 16);
        const int tx0 = (ax - h) / YSM_TILE, tx1 = (ax + h) / YSM_TILE;
        const int ty0 = (ay - h) / YSM_TILE, ty1 = (ay + h) / YSM_TILE;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int ty = ty0 + (q >> 1), tx = tx0 + (q & 1);
          const int t = ty * tnx + tx;
          const unsigned hsh = res_tile_hash(t);
          // owner = hash scaled to [0, G) (a multiply, not a division); the slot probe starts from other bits
          const bool mine = cv && ty <= ty1 && tx <= tx1 && (int)__umulhi(hsh, (unsigned)G) == bid;
          if (!__any_sync(0xffffffffu, mine)) continue;  // warp-uniform (a CTA owns ~ 1 tile in 148)
          const unsigned grp = __match_any_sync(0xffffffffu, mine ? t : -1);
          const int leader = __ffs(grp) - 1;
          int sl = -1, base = 0;
          if (mine && lane == leader) {
            for (int probe = 0; probe < maxt; probe++) {
              const int s2 = (int)((hsh >> 9) + (unsigned)probe) & (maxt - 1);
              const int old = atomicCAS(&s_tile[s2], -1, t);
              if (old == -1 || old == t) {
                sl = s2;
                base = atomicAdd(&s_cnt[s2], __popc(grp));
                break;
              }
            }
            if (sl < 0) *s_fail = 1;
          }
          sl = __shfl_sync(0xffffffffu, sl, leader);
          base = __shfl_sync(0xffffffffu, base, leader);
          if (mine && sl >= 0) {
            const int k2 = base + __popc(grp & ltmask);
            if (k2 < cand) s_steps[sl * cand + k2] = stamp_step(c[k], h, g.K, g.Wt, tx * YSM_TILE, ty * YSM_TILE);
            else *s_fail = 1;
          }
        }
      }
      continue;
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
          for (int probe = 0; probe < maxt && !placed; probe++) {
            const int sl = (int)((hsh >> 27) + (unsigned)probe) & (maxt - 1);
            const int old = atomicCAS(&s_tile[sl], -1, t);
            if (old == -1 || old == t) {
              const int k2 = atomicAdd(&s_cnt[sl], 1);
              if (k2 < cand) s_steps[sl * cand + k2] = stamp_step(c[k], h, g.K, g.Wt, tx * YSM_TILE, ty * YSM_TILE);
              else *s_fail = 1;
              placed = true;
            }
          }
          if (!placed) *s_fail = 1;
        }
    }
  }
}

// stamps the owned tiles (up to 8 warps share a tile) and writes each once
__device__ __forceinline__ void
res_stamp(const GridC& g, const ResArgs& A, int cand, const int* s_tile, const int* s_cnt, const uint32_t* s_steps,
          int* s_slots, int* s_nt, uint32_t* s_stage, uint32_t lane_tab_s, unsigned long long* pf) {
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
  if (threadIdx.x == 0) { pf[16] = res_timer(); pf[17] = pf[18] = pf[16]; }
  if (nt == 0) return;  // CTA-uniform
  const int S = max(1, min(8, nwarps / nt));
  const int ti = warp / S, part = warp - ti * S;
  const bool active = ti < nt;
  int slot = 0;
  if (active) {
    slot = s_slots[ti];
    const int cnt = min(s_cnt[slot], cand);
    const int per = (cnt + S - 1) / S;
    const int lo = min(cnt, part * per), hi = min(cnt, lo + per);
    uint32_t t[16];
#pragma unroll
    for (int k = 0; k < 16; k++) t[k] = 0u;
    tile_scatter_rows(t, s_steps + slot * cand + lo, hi - lo, lane_tab_s, g.K);
    uint32_t b[8];
#pragma unroll
    for (int k = 0; k < 8; k++) b[k] = __byte_perm(t[2 * k], t[2 * k + 1], 0x6420);  // u16 lanes -> bytes
    const int sw = (lane >> 2) & 1;
    uint4* st4 = reinterpret_cast<uint4*>(s_stage + (size_t)warp * (YSM_TILE * YSM_TILE / 4));
    st4[lane * 2 + (0 ^ sw)] = make_uint4(b[0], b[1], b[2], b[3]);
    st4[lane * 2 + (1 ^ sw)] = make_uint4(b[4], b[5], b[6], b[7]);
  }
  if (threadIdx.x == 0) pf[17] = res_timer();
  __syncthreads();
  if (threadIdx.x == 0) pf[18] = res_timer();
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
res_sweep_prep(const GridC& g, const ResReq& rq, const ResArgs& A, int a, int* s_i, int* s_minmax, const double2* s_qp,
               int n_qp) {
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
    const double2 w = p < n_qp ? s_qp[p] : __ldcg(reinterpret_cast<const double2*>(A.qpts) + p);
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

// one lattice row-task: the integer lookup sum of this lane's pose over points [0, np); 16 independent byte
// loads in flight per lane (s_off holds offsets biased to be non-negative, gp includes the bias)
__device__ __forceinline__ unsigned res_sweep_row(const uint8_t* gp, const unsigned* s_off, int np) {
  unsigned sum0 = 0, sum1 = 0;
  int p = 0;
  for (; p + 16 <= np; p += 16) {
    const uint4 o0 = *reinterpret_cast<const uint4*>(s_off + p);
    const uint4 o1 = *reinterpret_cast<const uint4*>(s_off + p + 4);
    const uint4 o2 = *reinterpret_cast<const uint4*>(s_off + p + 8);
    const uint4 o3 = *reinterpret_cast<const uint4*>(s_off + p + 12);
    const unsigned v0 = gp[o0.x], v1 = gp[o0.y], v2 = gp[o0.z], v3 = gp[o0.w];
    const unsigned v4 = gp[o1.x], v5 = gp[o1.y], v6 = gp[o1.z], v7 = gp[o1.w];
    const unsigned v8 = gp[o2.x], v9 = gp[o2.y], v10 = gp[o2.z], v11 = gp[o2.w];
    const unsigned v12 = gp[o3.x], v13 = gp[o3.y], v14 = gp[o3.z], v15 = gp[o3.w];
    sum0 += (v0 + v1 + v2 + v3) + (v4 + v5 + v6 + v7);
    sum1 += (v8 + v9 + v10 + v11) + (v12 + v13 + v14 + v15);
  }
  for (; p + 4 <= np; p += 4) {
    const uint4 o0 = *reinterpret_cast<const uint4*>(s_off + p);
    sum0 += (unsigned)gp[o0.x] + gp[o0.y] + gp[o0.z] + gp[o0.w];
  }
  for (; p < np; p++) sum1 += (unsigned)gp[s_off[p]];
  return sum0 + sum1;
}

// returns false (CTA-uniform) when a lookup could leave the grid: Karto's per-lookup bounds check is the
// general path's business
__device__ __forceinline__ bool
res_sweep_run(const GridC& g, const PenaltyC& pen, const ResReq& rq, const ResArgs& A, int a, int chunk, const int* s_i,
              const int* s_minmax, unsigned* s_part, unsigned long long* pf) {
  const PassDev& ps = rq.coarse;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  if (threadIdx.x == 0) pf[19] = res_timer();
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
  if (!(lo >= 0 && hi < (long long)g.data_size)) return false;
  const int np = ps.P;
  const int slice = warp % psplit, wtask = warp / psplit, ntw = nwarps / psplit;
  const int plen = (((np + psplit - 1) / psplit) + 3) & ~3;
  const int pb = min(np, slice * plen), pe = min(np, pb + plen);
  const int iters = (task1 - task0 + ntw - 1) / ntw;  // CTA-uniform
  const double ap = rq.ap[a];
  if (threadIdx.x == 0) pf[20] = res_timer();
  for (int it = 0; it < iters; it++) {
    const int task = task0 + wtask + it * ntw;
    const bool tv = wtask < ntw && task < task1;
    const int iy = tv ? task / nxc : 0, xc = tv ? task - iy * nxc : 0;
    const int ix = (xc << 5) + lane;
    const bool active = tv && ix < ps.nX;
    // the odometry penalty of the pose: distance part from the per-configuration table, angle part from the
    // request (both made by the host with the same IEEE operations as CorrelateScan); the load overlaps the lookups
    double dp = 1.0;
    if (ps.penalize && active && slice == 0) dp = __ldg(A.dp_coarse + (size_t)iy * ps.nX + ix);
    if (threadIdx.x == 0) pf[13] = res_timer();
    unsigned sum = 0;
    if (tv) {
      const int base = s_row[iy] + s_col[active ? ix : 0] + minoff;
      sum = res_sweep_row(A.grid + base, s_off + pb, pe - pb);
    }
    if (threadIdx.x == 0) pf[14] = res_timer();
    if (psplit > 1) {
      s_part[warp * 32 + lane] = sum;
      __syncthreads();
      if (slice == 0 && tv)
        for (int c = 1; c < psplit; c++) sum += s_part[(warp + c) * 32 + lane];
      __syncthreads();
    }
    if (active && slice == 0) {
      const double rr = response_from(ps, sum, dp * ap);
      A.resp[((size_t)iy * ps.nX + ix) * ps.nA + a] = rr;
      atomicMax(A.cellmax + (size_t)iy * ps.nX + ix, (unsigned long long)__double_as_longlong(rr));
    }
    if (threadIdx.x == 0) pf[15] = res_timer();
  }
  return true;
}

// The tail runs on a TEAM of a few warps of CTA 0 (named barrier 1): its steps are tiny, and every warp that
// merely walks through them costs issue slots and barrier time.
__device__ __forceinline__ void res_team_sync(int team_threads) {
  asm volatile("bar.sync 1, %0;" ::"r"(team_threads) : "memory");
}

// ---- tail (CTA 0) -----------------------------------------------------------------------------------------
struct ResTail {
  int list[YSM_RES_TIECAP];
  int sorted[YSM_RES_TIECAP];
  int count, n, first, status, has_fine, stop;
  unsigned hi3;
  double acc[4];
  double tmp4[32][4];
  PassOut po[2];
  int angs[YSM_RES_MAXNA];
};

// CorrelateScan epilogue of the coarse pass (reduce_body's logic, restated for a small team with few dependent
// L2 round trips), in two steps so that the winner can be published to the fine-pass workers in between:
//   res_reduce_winner  per-cell maxima -> shared memory, best response = their maximum, the responses of the
//                      few cells that can hold a tied pose, ties in storage order, sequential sums
//   res_reduce_cov     ComputePositionalCovariance accumulators (A.9) from the maxima in shared memory
// res_reduce_winner returns false when a list overflows (the host takes the general path).
__device__ __forceinline__ bool
res_reduce_winner(const ResReq& rq, const ResArgs& A, ResTail& S, int T, double* s_cm) {
  const PassDev& ps = rq.coarse;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
  const int ncell = ps.nX * ps.nY;
  if (tid == 0) { S.count = 0; S.n = 0; }
  double mx = 0.0;
  for (int c = tid; c < ncell; c += T) {
    const double m = __longlong_as_double((long long)__ldcg(A.cellmax + c));
    if (c < YSM_RES_CM_SMEM) s_cm[c] = m;
    mx = m > mx ? m : mx;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double t = __shfl_xor_sync(0xffffffffu, mx, o);
    mx = t > mx ? t : mx;
  }
  if (lane == 0) S.tmp4[warp][0] = mx;
  res_team_sync(T);
  double best = 0.0;  // (responses are >= 0; CorrelateScan's -1 initial value never survives)
  for (int w = 0; w < nwarps; w++) best = S.tmp4[w][0] > best ? S.tmp4[w][0] : best;
  // cells whose maximum is within the tie tolerance of the best (S.sorted doubles as the cell list)
  for (int c = tid; c < ncell; c += T) {
    const double m = c < YSM_RES_CM_SMEM ? s_cm[c] : __longlong_as_double((long long)__ldcg(A.cellmax + c));
    if (m >= best - YSM_KT_TOLERANCE) {
      const int pos = atomicAdd(&S.n, 1);
      if (pos < YSM_RES_TIECAP) S.sorted[pos] = c;
    }
  }
  res_team_sync(T);
  const int ncand = S.n;
  if (ncand > YSM_RES_TIECAP) return false;
  for (int it = tid; it < ncand * ps.nA; it += T) {
    const int j = it / ps.nA, a = it - j * ps.nA;
    const int c = S.sorted[j];
    const double r = __ldcg(A.resp + (size_t)c * ps.nA + a);
    if (kt_double_equal(r, best)) {
      const int pos = atomicAdd(&S.count, 1);
      if (pos < YSM_RES_TIECAP) S.list[pos] = c * ps.nA + a;
    }
  }
  res_team_sync(T);
  const int nt = S.count;
  if (nt > YSM_RES_TIECAP) return false;
  const double startX = -ps.offx, startY = -ps.offy;
  for (int e = tid; e < nt; e += T) {
    const int v = S.list[e];
    int rank = 0;
    for (int k = 0; k < nt; k++) rank += (S.list[k] < v);
    S.sorted[rank] = v;
  }
  res_team_sync(T);
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
    if (nt == 1) {  // (x / 1.0 == x: skip four f64 divisions on the usual path)
      po.avg_x = sx; po.avg_y = sy; po.tx = tx; po.ty = ty;
    } else {
      po.avg_x = nt > 0 ? sx / cnt : 0.0;
      po.avg_y = nt > 0 ? sy / cnt : 0.0;
      po.tx = nt > 0 ? tx / cnt : 0.0;
      po.ty = nt > 0 ? ty / cnt : 0.0;
    }
    po.norm = 0.0; po.axx = 0.0; po.axy = 0.0; po.ayy = 0.0;
    po.n_ties = nt;
    po.first_idx = nt > 0 ? S.sorted[0] : -1;
  }
  res_team_sync(T);
  return true;
}

__device__ __forceinline__ void
res_reduce_cov(const ResReq& rq, const ResArgs& A, ResTail& S, int T, const double* s_cm) {
  const PassDev& ps = rq.coarse;
  const int tid = threadIdx.x, lane = tid & 31;
  const int ncell = ps.nX * ps.nY;
  const double best = S.po[0].best;
  const double startX = -ps.offx, startY = -ps.offy;
  // probs(x, y) = max response over the angles
  double norm = 0.0, axx = 0.0, axy = 0.0, ayy = 0.0;
  if (!(best < YSM_KT_TOLERANCE)) {
    const double dx = S.po[0].avg_x - ps.cx, dy = S.po[0].avg_y - ps.cy;
    for (int c = tid; c < ncell; c += T) {
      const double pm = c < YSM_RES_CM_SMEM ? s_cm[c] : __longlong_as_double((long long)__ldcg(A.cellmax + c));
      if (pm >= (best - 0.1)) {
        const int iy = c / ps.nX, ix = c - iy * ps.nX;
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
  if (lane == 0) {
    S.tmp4[tid >> 5][0] = norm; S.tmp4[tid >> 5][1] = axx; S.tmp4[tid >> 5][2] = axy; S.tmp4[tid >> 5][3] = ayy;
  }
  res_team_sync(T);
  if (tid < 32) {
    const int nw = (int)(T >> 5);
    double v0 = lane < nw ? S.tmp4[lane][0] : 0.0, v1 = lane < nw ? S.tmp4[lane][1] : 0.0;
    double v2 = lane < nw ? S.tmp4[lane][2] : 0.0, v3 = lane < nw ? S.tmp4[lane][3] : 0.0;
    for (int o = 16; o > 0; o >>= 1) {
      v0 += __shfl_xor_sync(0xffffffffu, v0, o);
      v1 += __shfl_xor_sync(0xffffffffu, v1, o);
      v2 += __shfl_xor_sync(0xffffffffu, v2, o);
      v3 += __shfl_xor_sync(0xffffffffu, v3, o);
    }
    if (lane == 0) { S.po[0].norm = v0; S.po[0].axx = v1; S.po[0].axy = v2; S.po[0].ayy = v3; }
  }
  res_team_sync(T);
}

// ---- fine CorrelateScan at the coarse winner, spread over the machine ------------------------------------
// CTA 0 publishes the winner (angle, lattice cell); the workers, idle after the coarse sweep, each take
// (fine angle, 32 query points) items: rotate the points (ComputeOffsets), look up the 3 x 3 cells, one warp
// reduction per pose, integer atomics into fsum[iy][ix][a]. CTA 0 then finishes: responses, max / ties,
// angular covariance sums. (On ONE SM the 3 x 3 x nAf x P scattered byte lookups cost ~0.5 cycle each.)
#define YSM_RES_WIN_NOFINE 0xFFFFFFFFu

// the fine pass's search centre from the winning coarse lattice cell: with a single winner the tie average
// is (cx + (startX + ix * resx)) / 1.0 -- the same expression everywhere
__device__ __forceinline__ void res_fine_centre(const PassDev& ps, int ix, int iy, double& cx, double& cy) {
  cx = (0.0 + (ps.cx + (-ps.offx + (double)ix * ps.resx))) / 1.0;
  cy = (0.0 + (ps.cy + (-ps.offy + (double)iy * ps.resy))) / 1.0;
}

// one warp, one item. The winner record (written by CTA 0 before it released the winner word) holds what does not
// depend on the item: the flat grid index of each of the 3 x 3 lattice cells, then the (cos, sin) of every fine angle.
__device__ __forceinline__ void
res_fine_item(const GridC& g, const ResReq& rq, const ResArgs& A, int item, double2 pt, bool valid) {
  const PassDev& f = rq.fine;
  const TableDev& ft = rq.ftab;
  const int lane = threadIdx.x & 31;
  const int nAf = f.nA, nxy = f.nX * f.nY;
  const int a = item % nAf;  // fine angle (the chunk of points came with `pt`)
  const double* rows = reinterpret_cast<const double*>(A.winrec + 64);
  const double cosine = __ldcg(rows + 2 * a), sine = __ldcg(rows + 2 * a + 1);
  int cb = 0;
  if (lane < nxy) cb = __ldcg(A.winrec + lane);  // (nxy <= 32: one cell base per lane, handed round by shuffles)
  int gx, gy;
  offset_cell(ft, g.scale, pt.x, pt.y, cosine, sine, gx, gy);
  const int o = gx + gy * g.stride;
  const unsigned dsz = (unsigned)g.data_size;
  for (int c0 = 0; c0 < nxy; c0 += 9) {  // (3 x 3 cells: one round of independent loads)
    unsigned v[9];
#pragma unroll
    for (int u = 0; u < 9; u++) {
      const int base = __shfl_sync(0xffffffffu, cb, (c0 + u) & 31);
      const unsigned idx = c0 + u < nxy ? (unsigned)(base + o) : 0xFFFFFFFFu;
      v[u] = (valid && idx < dsz) ? (unsigned)A.grid[idx] : 0u;
    }
#pragma unroll
    for (int u = 0; u < 9; u++) {
      const unsigned r = __reduce_add_sync(0xffffffffu, v[u]);
      if (lane == 0 && c0 + u < nxy && r) atomicAdd(A.fsum + (c0 + u) * nAf + a, r);
    }
  }
}

// CTA 0, after the workers' sums: responses, max / ties (storage order), angular covariance sums.
// s_fsum: the sums copied to shared memory; s_ft: [nAf][4] trig rows of the winning angle.
__device__ __forceinline__ bool
res_fine_finish(const GridC& g, const ResArgs& A, ResTail& S, const PassDev& f, const double* s_ft, const double* s_fap,
                const unsigned* s_fsum, double* s_fr, int T) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = T >> 5;
  const int nAf = f.nA, nxy = f.nX * f.nY, nposes = nxy * nAf;
  // odometry penalty: distance part from the per-configuration table (A.dp_fine), angle part from the spec
  // tables (s_fap: the host evaluated it for this winner's heading)
  if (tid == 0) S.count = 0;
  res_team_sync(T);
  double mx = 0.0;
  for (int pose = tid; pose < nposes; pose += T) {
    const int c = pose / nAf, a = pose - c * nAf;
    const double rr = response_from(f, s_fsum[pose], f.penalize ? __ldg(A.dp_fine + c) * s_fap[a] : 1.0);
    s_fr[pose] = rr;
    mx = rr > mx ? rr : mx;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double t = __shfl_xor_sync(0xffffffffu, mx, o);
    mx = t > mx ? t : mx;
  }
  if (lane == 0) S.tmp4[warp][0] = mx;
  res_team_sync(T);
  double best = 0.0;
  for (int w = 0; w < nwarps; w++) best = S.tmp4[w][0] > best ? S.tmp4[w][0] : best;
  // ties (few): collected in any order, then rank-sorted into storage order
  for (int i = tid; i < nposes; i += T)
    if (kt_double_equal(s_fr[i], best)) {
      const int pos = atomicAdd(&S.count, 1);
      if (pos < YSM_RES_TIECAP) S.sorted[pos] = i;
    }
  res_team_sync(T);
  const int nt = S.count;
  if (nt > YSM_RES_TIECAP) return false;
  for (int e = tid; e < nt; e += T) {
    const int v = S.sorted[e];
    int rank = 0;
    for (int k = 0; k < nt; k++) rank += (S.sorted[k] < v);
    S.list[rank] = v;
  }
  res_team_sync(T);
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
    if (nt == 1) {
      po.avg_x = sx; po.avg_y = sy; po.tx = tx; po.ty = ty;
    } else {
      po.avg_x = nt > 0 ? sx / cnt : 0.0;
      po.avg_y = nt > 0 ? sy / cnt : 0.0;
      po.tx = nt > 0 ? tx / cnt : 0.0;
      po.ty = nt > 0 ? ty / cnt : 0.0;
    }
    po.norm = 0.0; po.axx = 0.0; po.axy = 0.0; po.ayy = 0.0;
    po.n_ties = nt;
    po.first_idx = nt > 0 ? S.list[0] : -1;
    S.first = -1;
  }
  res_team_sync(T);
  // ComputeAngularCovariance sums: un-penalised GetResponse at the best cell for every fine angle. The best
  // cell is one of the lattice cells whenever the mean rounds onto one (else: general path).
  if (nt > 0) {
    const int gx = world_to_grid1(S.po[1].avg_x, f.gox, g.scale) + g.border;
    const int gy = world_to_grid1(S.po[1].avg_y, f.goy, g.scale) + g.border;
    if (tid < nxy) {
      const int iy = tid / f.nX, ix = tid - iy * f.nX;
      const double x = -f.offx + (double)ix * f.resx;
      const double y = -f.offy + (double)iy * f.resy;
      const int cx = world_to_grid1(f.cx + x, f.gox, g.scale) + g.border;
      const int cy = world_to_grid1(f.cy + y, f.goy, g.scale) + g.border;
      if (cx == gx && cy == gy) atomicMax(&S.first, tid);
    }
    res_team_sync(T);
    const int cell = S.first;
    if (cell < 0) return false;
    for (int a = tid; a < nAf; a += T) S.angs[a] = (int)s_fsum[cell * nAf + a];
  } else {
    for (int a = tid; a < nAf; a += T) S.angs[a] = 0;
  }
  res_team_sync(T);
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
  __shared__ int s_nt, s_cmd, s_fail;
  __shared__ int s_io[2];
  __shared__ uint4 s_db;
  __shared__ __align__(8) unsigned s_line[12];  // this poller's scan: count, source slot, slot to fill, pad, m fields
  __shared__ int s_minmax[8];
  __shared__ __align__(8) int s_misc[16];
  __shared__ double s_wmax[32];
  __shared__ unsigned long long s_ts[YSM_RES_TS];
  __shared__ unsigned long long s_pf[YSM_RES_PROF];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, T = blockDim.x;
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
  const bool worker = bid >= 1;
  const ResReq* hreq = reinterpret_cast<const ResReq*>(A.req);
  unsigned long long* bar1 = A.bars;
  unsigned long long* bar2 = A.bars + YSM_RES_BAR_STRIDE;
  unsigned long long* bar_sweep = A.bars + 2 * YSM_RES_BAR_STRIDE;
  unsigned long long* bar_fine = A.bars + 3 * YSM_RES_BAR_STRIDE;
  unsigned long long* bar3 = A.bars + 4 * YSM_RES_BAR_STRIDE;
  unsigned seq_done = A.last_seq;
  unsigned round = 0u, t1 = 0u, t2 = 0u, t3 = 0u, t4 = 0u, t5 = 0u;
  unsigned fails2 = 0u, fails3 = 0u;  // failure counts seen so far at barrier 2 / the sweep-done counter
  unsigned long long idle0 = res_timer();
  int exit_code = 0;
  if (tid < YSM_RES_MAXT) { s_tile[tid] = -1; s_cnt[tid] = 0; }
  if (tid == 0) { s_nt = 0; s_fail = 0; }
  __syncthreads();
#define RES_TS(k) if (bid == 0 && tid == 0) s_ts[k] = res_timer();
  for (;;) {
    round++;
    unsigned long long* pf = s_pf;  // per-CTA phase timestamps (thread 0)
    // ---- wait for the doorbell -----------------------------------------------------------------------
    int cmd = RES_CMD_NONE;
    if (poller) {
      if (warp == 0) {
        // Lanes 0..4 read the five tagged words of this CTA's line with ONE warp-wide load (one thread's
        // system-scope loads are served one after the other; five lanes' loads travel together).
        const unsigned char* line = reinterpret_cast<const unsigned char*>(A.db) + (size_t)bid * YSM_RES_DB_STRIDE;
        int c = RES_CMD_NONE;
        for (unsigned spins = 1u;; spins++) {
          uint4 v = make_uint4(0u, 0u, 0u, 0u);
          if (lane < 5) v = res_ld_volatile_v4(line + 16 * lane);
          const unsigned tag = lane < 2 ? v.x : v.w;
          const unsigned tag0 = __shfl_sync(0xffffffffu, tag, 0);
          const bool fresh = __all_sync(0xffffffffu, lane >= 5 || tag == tag0) && tag0 != seq_done;
          if (fresh) {
            if (lane == 0) { s_db = v; c = (int)(v.y & 0x7Fu); }
            else if (lane == 1) { s_line[0] = v.y; s_line[1] = v.z; s_line[2] = v.w; }
            else if (lane == 2) { s_line[4] = v.x; s_line[5] = v.y; s_line[6] = v.z; }
            else if (lane == 3) { s_line[7] = v.x; s_line[8] = v.y; s_line[9] = v.z; }
            else if (lane == 4) { s_line[10] = v.x; s_line[11] = v.y; }
            break;
          }
          int quit = 0;
          if (lane == 0) {
            if (bid == 0) {
              if ((spins & 7u) == 0u && res_timer() - idle0 > A.idle_ns) {
                quit = 1;
                res_st_release(A.quit_round, round);
              }
            } else if ((spins & 3u) == 0u && res_ld_acquire(A.quit_round) == round) {
              quit = 1;
            }
          }
          if (__shfl_sync(0xffffffffu, quit, 0)) {
            if (lane == 0) { s_db = v; c = RES_CMD_QUIT; }
            break;
          }
        }
        if (lane == 0) s_cmd = c;
      }
      __syncthreads();
      cmd = s_cmd;
    }
    if (bid == 0 && tid == 0) { s_ts[0] = res_timer(); s_ts[20] = (unsigned long long)clock64(); }
    if (threadIdx.x == 0) pf[8] = res_timer();
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
        const double* mf = reinterpret_cast<const double*>(s_line + 4);  // viewpoint x, y, grid offset x, y
        for (int s = bid - 1; s <= nbase; s += YSM_RES_POLLERS) {
          // the poller's own line describes its first scan; a second round (more than 16 scans) asks the
          // control block in host memory
          int cnt = (int)s_line[0], src_slot = (int)s_line[1], store_slot = (int)s_line[2];
          if (s != bid - 1) {
            if (tid == 0) s_misc[8] = (int)__ldcv(&hreq->counts[s]);
            __syncthreads();
            cnt = s_misc[8];
            src_slot = store_slot = 0;
          }
          if (s < nbase) {
            res_filter_scan(g, A, s, pstride, cnt, src_slot, store_slot, mf, dyn, pf);
          } else {
            const int Pq = (int)(d.z & 0xFFFFu);
            double2* dst = reinterpret_cast<double2*>(A.qpts);
            if (src_slot > 0) {
              const double2* src = reinterpret_cast<const double2*>(A.cache) + (size_t)(src_slot - 1) * YSM_RES_PMAX;
              for (int i = tid; i < Pq; i += T) dst[i] = __ldcg(src + i);
            } else {
              const double2* src = reinterpret_cast<const double2*>(A.pts + 2 * (size_t)s * pstride);
              double2* keep = store_slot > 0 ? reinterpret_cast<double2*>(A.cache) + (size_t)(store_slot - 1) * YSM_RES_PMAX : nullptr;
              for (int i = tid; i < Pq; i += T) {
                const double2 w = __ldcv(src + i);
                dst[i] = w;
                if (keep) keep[i] = w;
              }
            }
          }
          __syncthreads();
        }
      }
    } else if (bid == 0) {
      // quit / ping / idle: a two-word control block made here
      if (tid == 0) {
        unsigned* c = reinterpret_cast<unsigned*>(A.ctl);
        c[0] = s_db.x;
        c[1] = (unsigned)cmd;
      }
      if (cmd == RES_CMD_PING && tid == 0)
        res_st_volatile_v4(A.out, make_uint4((unsigned)RES_ST_PONG, 1u, s_db.x, 0u));
    }
    if (threadIdx.x == 0) pf[12] = res_timer();
    RES_TS(1)
    // ---- barrier 1: cells, query points and control block are in HBM ----------------------------
    t1 += (unsigned)G;
    if (!res_barrier(bar1, t1, false, A.abort_flag, s_io, A.stall_ns)) { exit_code = 2; break; }
    RES_TS(2)
    // ONE round of loads after the barrier: the control block (header, descriptors, the first 32 trig rows = the
    // usual 21 search angles) -> shared memory; speculatively, the first 1024 query points -> shared memory and
    // the first 4096 cell entries -> registers (how many are real is only known once the header is here)
    double2* s_qp = reinterpret_cast<double2*>(dyn + A.o_off);  // [T] query points (CTA 0: the tail's s_q)
    int* s_offs = reinterpret_cast<int*>(dyn + A.o_off + 16 * (size_t)YSM_RES_THREADS);
    uint32_t c_first[4];
    {
      const int hdr4 = (int)(offsetof(ResReq, trig4) / 16);
      const uint4* src = reinterpret_cast<const uint4*>(A.ctl);
      uint4* dst = reinterpret_cast<uint4*>(&s_rq);
      for (int i = tid; i < hdr4 + 64; i += T) dst[i] = __ldcg(src + i);
      s_qp[tid] = __ldcg(reinterpret_cast<const double2*>(A.qpts) + tid);
#pragma unroll
      for (int k = 0; k < 4; k++) c_first[k] = __ldcg(A.cells + tid + k * T);
      __syncthreads();
    }
    const unsigned seq = s_rq.seq;
    cmd = (int)s_rq.cmd;
    if (cmd == RES_CMD_QUIT || cmd == RES_CMD_NONE) break;
    if (cmd == RES_CMD_PING) {
      seq_done = seq;
      idle0 = res_timer();
      __syncthreads();
      continue;
    }
    if (s_rq.nA > 32) {
      const int hdr4 = (int)(offsetof(ResReq, trig4) / 16);
      const uint4* src = reinterpret_cast<const uint4*>(A.ctl);
      uint4* dst = reinterpret_cast<uint4*>(&s_rq);
      for (int i = hdr4 + 64 + tid; i < hdr4 + 2 * s_rq.nA; i += T) dst[i] = __ldcg(src + i);
      __syncthreads();
    }
    if (threadIdx.x == 0) pf[0] = res_timer();
    const ResReq& rq = s_rq;
    const PassDev& ps = rq.coarse;
    const int nv = rq.nA * rq.task_chunks;
    int v = bid - 1;
    if (threadIdx.x == 0) pf[1] = res_timer();
    // ---- phase B: stamp the tiles this CTA owns -------------------------------------------------------
    {
      const int total = min(rq.nbase * rq.pstride, A.cells_cap);
      uint32_t* s_steps = reinterpret_cast<uint32_t*>(dyn);
      uint32_t* s_stage = s_steps + YSM_RES_MAXT * YSM_RES_CAND;
      res_collect(g, A, total, G, bid, rq.maxt, rq.cand, s_tile, s_cnt, s_steps, &s_fail, c_first);
      __syncthreads();
      if (threadIdx.x == 0) pf[2] = res_timer();
      res_stamp(g, A, rq.cand, s_tile, s_cnt, s_steps, s_slots, &s_nt, s_stage, lane_tab_s, pf);
    }
    if (threadIdx.x == 0) pf[3] = res_timer();
    RES_TS(3)
    // ---- barrier 2, split: arrive, do what does not depend on the grid, then wait -----------------
    t2 += (unsigned)G;
    __syncthreads();
    if (tid == 0) res_arrive(bar2, s_fail != 0);
    // lookup offsets of this CTA's first angle
    if (worker && v < nv) res_sweep_prep(g, rq, A, v % rq.nA, s_offs, s_minmax, s_qp, (int)T);
    // the fine pass's items of this CTA (fine angle x 32 query points; warp w takes item bid - 1 + w (G - 1)):
    // the points wait in registers
    const int fine_chunks = (ps.P + 31) >> 5, fine_items = rq.do_refine ? rq.nAf * fine_chunks : 0;
    const int my_item = (bid - 1) + warp * (G - 1);
    const bool have_item = worker && my_item < fine_items;
    double2 my_pt = make_double2(0.0, 0.0);
    bool my_valid = false;
    if (have_item) {
      const int p = (my_item / rq.nAf) * 32 + lane;
      my_valid = p < ps.P;
      if (my_valid) my_pt = p < (int)T ? s_qp[p] : __ldcg(reinterpret_cast<const double2*>(A.qpts) + p);
    }
    if (tid == 0) {
      unsigned hi = 0u;
      s_io[0] = res_wait(bar2, t2, &hi, A.abort_flag, A.stall_ns) ? 1 : 0;
      s_io[1] = (int)hi;
      s_fail = 0;
    }
    __syncthreads();
    if (!s_io[0]) { exit_code = 2; break; }
    RES_TS(4)
    if (threadIdx.x == 0) pf[4] = res_timer();
    const bool failed = (unsigned)s_io[1] != fails2;  // some CTA overflowed its tile lists
    fails2 = (unsigned)s_io[1];
    // ---- phase C: coarse sweep ------------------------------------------------------------------------------
    if (worker && !failed) {
      bool first = true;
      for (; v < nv; v += G - 1) {
        if (!first) {
          __syncthreads();
          res_sweep_prep(g, rq, A, v % rq.nA, s_offs, s_minmax, s_qp, (int)T);
        }
        first = false;
        if (!res_sweep_run(g, pen, rq, A, v % rq.nA, v / rq.nA, s_offs, s_minmax,
                           reinterpret_cast<unsigned*>(dyn), pf) && tid == 0)
          s_fail = 1;
      }
    }
    if (threadIdx.x == 0) pf[5] = res_timer();
    if (worker) {
      __syncthreads();
      if (tid == 0) {
        res_arrive(bar_sweep, s_fail != 0);
        s_fail = 0;
      }
      // ---- fine pass items: every warp that holds one waits for the winner, sums, arrives on its own -----
      if (have_item) {
        unsigned long long w = 0ull;
        if (lane == 0) {
          unsigned spins = 0u;
          unsigned long long t0 = 0ull;
          for (;;) {
            w = res_ld_acquire64(A.win);
            if ((unsigned)(w >> 32) == seq) break;
            if ((++spins & 0x3FFu) == 0u) {
              const unsigned long long now = res_timer();
              if (t0 == 0ull) t0 = now;
              if (now - t0 > A.stall_ns) atomicExch(A.abort_flag, 1);
              if (res_ld_acquire(reinterpret_cast<const unsigned*>(A.abort_flag)) != 0u) {
                w = ((unsigned long long)seq << 32) | YSM_RES_WIN_NOFINE;  // (barrier 3 ends the kernel)
                break;
              }
            }
          }
        }
        w = __shfl_sync(0xffffffffu, w, 0);
        const unsigned winw = (unsigned)w;
        if (winw != YSM_RES_WIN_NOFINE) res_fine_item(g, rq, A, my_item, my_pt, my_valid);
        __syncwarp();
        if (lane == 0) res_arrive(bar_fine, false);
      }
    }
    if (bid == 0) {
      // spec tables: the host writes them while the GPU runs phases A-C; a copy goes to HBM for the workers
      int status = failed ? RES_ST_FALLBACK : RES_ST_OK;
      int has_fine = 0;
      double* s_spec = reinterpret_cast<double*>(dyn + A.o_off + rq.o_spec);
      const int spec_doubles = rq.nA + 5 * rq.nA * rq.nAf;  // headings | trig rows | fine angle penalties
      if (rq.do_refine && !failed) {
        if (tid == 0) {
          const unsigned long long t0 = res_timer();
          unsigned spins = 0u;
          int ok = 1;
          while (res_ld_volatile_v4(A.spec).x != seq) {
            if ((++spins & 0xFFu) == 0u && res_timer() - t0 > 2000000000ull) { ok = 0; break; }
          }
          s_io[0] = ok;
        }
        __syncthreads();
        if (!s_io[0]) status = RES_ST_FALLBACK;
        else {
          const double2* src = reinterpret_cast<const double2*>(A.spec + sizeof(ResSpecHdr));
          double2* dst = reinterpret_cast<double2*>(s_spec);
          double2* dst2 = reinterpret_cast<double2*>(A.spec_dev);
          for (int i = tid; i < (spec_doubles + 1) / 2; i += T) {
            const double2 w = __ldcv(src + i);
            dst[i] = w;
            dst2[i] = w;
          }
        }
      }
      __syncthreads();
      RES_TS(5)
      // ---- the tail proper: a team of a few warps (named barrier), the other warps wait below ----------
      const int TT = min((int)T, max(128, rq.tail_warps * 32));
      t3 += (unsigned)(G - 1);
      if (rq.do_refine) t4 += (unsigned)fine_items;
      if (tid == 0) { s_tail.status = status; s_tail.has_fine = 0; s_tail.stop = 0; }
      if (tid < TT) {
        if (tid == 0) {
          unsigned hi = 0u;
          if (!res_wait(bar_sweep, t3, &hi, A.abort_flag, A.stall_ns)) s_tail.stop = 1;
          s_tail.hi3 = hi;
          if (hi != fails3) s_tail.status = RES_ST_FALLBACK;  // (a sweep CTA met a window that can leave the grid)
        }
        res_team_sync(TT);
        RES_TS(6)
        bool stop = s_tail.stop != 0;
        double* s_cm = reinterpret_cast<double*>(dyn + A.o_off + rq.o_cm);
        if (!stop && s_tail.status == RES_ST_OK) {
          if (!res_reduce_winner(rq, A, s_tail, TT, s_cm) && tid == 0) s_tail.status = RES_ST_FALLBACK;
          res_team_sync(TT);
        }
        RES_TS(7)
        const bool coarse_ok = !stop && s_tail.status == RES_ST_OK;
        bool go = false;
        int a = 0;
        if (!stop && rq.do_refine) {
          // MatchScan goes straight to the fine pass when the coarse pass has ONE winner with a non-zero
          // response: its centre is the winning lattice pose, the heading the atan2(sin, cos) the host tabulated
          // for that coarse angle. Otherwise the host reschedules the match (ties, response expansion).
          const PassOut& po = s_tail.po[0];
          go = coarse_ok && po.n_ties == 1 && po.best > YSM_KT_TOLERANCE;
          a = go ? po.first_idx % ps.nA : 0;
          PassDev& f = s_rq.fine;
          if (go) {
            // the winner record for the workers: flat index of the 3 x 3 cells, (cos, sin) of the fine angles
            const int nxy = f.nX * f.nY;
            if (tid < nxy) {
              const int iy = tid / f.nX, ix = tid - iy * f.nX;
              const double x = -f.offx + (double)ix * f.resx;
              const double y = -f.offy + (double)iy * f.resy;
              const int bx = world_to_grid1(po.avg_x + x, f.gox, g.scale) + g.border;
              const int by = world_to_grid1(po.avg_y + y, f.goy, g.scale) + g.border;
              A.winrec[tid] = bx + by * g.stride;
            } else if (tid >= 32 && tid < 32 + 2 * rq.nAf) {
              const int k = tid - 32, fa = k >> 1;
              reinterpret_cast<double*>(A.winrec + 64)[k] = s_spec[rq.nA + 4 * ((size_t)a * rq.nAf + fa) + (k & 1)];
            }
          }
          res_team_sync(TT);  // (record written, status read by everyone)
          if (tid == 0) {
            res_st_release64(A.win, ((unsigned long long)seq << 32) | (go ? (unsigned)a : YSM_RES_WIN_NOFINE));
            if (!go) s_tail.status = RES_ST_FALLBACK;
            f.cx = po.avg_x;
            f.cy = po.avg_y;
            f.ch = s_spec[a];
          }
          RES_TS(9)
        }
        // the A.9 accumulators, while the workers sum the fine pass
        if (coarse_ok) res_reduce_cov(rq, A, s_tail, TT, s_cm);
        if (!stop && rq.do_refine) {
          if (tid == 0) {
            unsigned hi = 0u;
            if (!res_wait(bar_fine, t4, &hi, A.abort_flag, A.stall_ns)) s_tail.stop = 1;
          }
          res_team_sync(TT);
          RES_TS(10)
          stop = s_tail.stop != 0;
          if (go && !stop) {
            const PassDev& f = s_rq.fine;
            const int nposes = f.nX * f.nY * f.nA;
            unsigned* s_fsum = reinterpret_cast<unsigned*>(dyn + A.o_off + rq.o_fsum);
            double* s_fr = reinterpret_cast<double*>(s_fsum + ((nposes + 1) & ~1));
            for (int i = tid; i < nposes; i += TT) {
              s_fsum[i] = __ldcg(A.fsum + i);
              A.fsum[i] = 0u;  // (for the next request)
            }
            res_team_sync(TT);
            const double* s_ft = s_spec + rq.nA + (size_t)4 * a * rq.nAf;
            const double* s_fap = s_spec + rq.nA + (size_t)4 * rq.nA * rq.nAf + (size_t)a * rq.nAf;
            const bool ok = res_fine_finish(g, A, s_tail, f, s_ft, s_fap, s_fsum, s_fr, TT);
            if (tid == 0) {
              if (ok) s_tail.has_fine = 1;
              else s_tail.status = RES_ST_FALLBACK;
            }
          }
        }
        RES_TS(8)
        if (tid == 0) s_ts[21] = (unsigned long long)clock64();
      }
      __syncthreads();
      fails3 = s_tail.hi3;
      if (s_tail.stop) { exit_code = 2; break; }
      status = s_tail.status;
      has_fine = s_tail.has_fine;
      res_publish(A, s_tail, seq, status, has_fine, rq.nAf, rq.trace ? s_ts : nullptr);
      // reset the per-request accumulators for the next request
      const int ncell = ps.nX * ps.nY;
      for (int i = tid; i < ncell; i += T) A.cellmax[i] = 0ull;
    }
    // ---- barrier 3: the fine pass has read the grid; zero the tiles this CTA stamped ----------------
    t5 += (unsigned)G;
    if (!res_barrier(bar3, t5, false, A.abort_flag, s_io, A.stall_ns)) { exit_code = 2; break; }
    res_clear(g, A, s_tile, s_slots, s_nt);
    if (rq.trace && A.prof && tid == 0) {
      pf[6] = res_timer();
      pf[7] = (unsigned long long)s_nt;
      for (int k = 0; k < YSM_RES_PROF; k++) A.prof[(size_t)bid * YSM_RES_PROF + k] = pf[k];
    }
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
