// ysm.cu -- host runtime + C ABI (include/ysm.h) of the B200-native correlative scan matcher.
//
// Host side of ScanMatcher::MatchScan: it owns the correlation-grid slots and
// workspaces, schedules CorrelateScan passes (coarse -> response expansion -> fine) over a
// wave of independent matches, evaluates every transcendental (cos/sin/atan2/exp/hypot) with
// libm so results are bit-identical to the CPU reference, and launches the sm_100a kernels in
// ysm_kernels.cuh for all of the gather/reduce work. There is no CPU compute fallback.
#include "../../include/ysm.h"
#include "ysm_internal.h"
#include "ysm_kernels.cuh"
#include "ysm_resident.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <sched.h>

#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <new>
#include <exception>
#include <thread>
#include <string>
#include <unordered_map>
#include <vector>

using namespace ysm;

static_assert(sizeof(ysm_result) == 128, "ysm_result must be a 128-byte record");

#define KT_PI 3.14159265358979323846
#define KT_2PI 6.28318530717958647692
#define KT_PI_180 0.01745329251994329577
#define KT_TOLERANCE 1e-06
#define MAX_VARIANCE 500.0

static std::string g_create_error;

#include <chrono>
struct KernelTrace {
  bool on = false;
  cudaEvent_t ev[32];
  const char* name[32];
  int n = 0;
  cudaStream_t st = nullptr;
  void init(bool enable, cudaStream_t s) {
    on = enable; st = s; n = 0;
    if (on) for (int i = 0; i < 32; i++) cudaEventCreate(&ev[i]);
  }
  void mark(const char* what) {
    if (!on || n >= 32) return;
    name[n] = what;
    cudaEventRecord(ev[n++], st);
  }
  void dump() {
    if (!on) return;
    cudaEventSynchronize(ev[n - 1]);
    for (int i = 1; i < n; i++) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
      fprintf(stderr, "[ysm-gpu] %-22s %8.1f us\n", name[i], ms * 1e3);
    }
    for (int i = 0; i < 32; i++) cudaEventDestroy(ev[i]);
    n = 0; on = false;
  }
};
struct PhaseTrace {
  bool on;
  std::chrono::steady_clock::time_point t0, last;
  static bool enabled() {
    static const bool e = getenv("YSM_TRACE") != nullptr;  // read once: no getenv on the hot path
    return e;
  }
  PhaseTrace() : on(enabled()) {
    if (on) t0 = last = std::chrono::steady_clock::now();
  }
  void mark(const char* what) {
    if (!on) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[ysm %03x] %-28s +%8.1f us (t=%8.1f)\n", (unsigned)(((uintptr_t)pthread_self()) >> 12) & 0xfffu, what,
            std::chrono::duration<double, std::micro>(now - last).count(),
            std::chrono::duration<double, std::micro>(now - t0).count());
    last = now;
  }
};

namespace {

inline double h_round(double v) { return v >= 0.0 ? floor(v + 0.5) : ceil(v - 0.5); }
inline bool h_double_equal(double a, double b) { return fabs(a - b) <= KT_TOLERANCE; }
inline double h_square(double v) { return v * v; }
inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

double h_normalize_angle(double angle) {
  while (angle < -KT_PI) {
    if (angle < -KT_2PI) angle += (uint32_t)(angle / -KT_2PI) * KT_2PI;
    else angle += KT_2PI;
  }
  while (angle > KT_PI) {
    if (angle > KT_2PI) angle -= (uint32_t)(angle / KT_2PI) * KT_2PI;
    else angle -= KT_2PI;
  }
  return angle;
}

double h_normalize_angle_difference(double minuend, double subtrahend) {
  while (minuend - subtrahend < -KT_PI) minuend += KT_2PI;
  while (minuend - subtrahend > KT_PI) minuend -= KT_2PI;
  return minuend;
}

// growable device buffer
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

// growable pinned host buffer
struct PinBuf {
  void* p = nullptr;
  void* dptr = nullptr;  // device view of the mapped allocation
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    // mapped: the latency kernel reads its inputs from / writes its results to this memory directly
    cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocMapped);
    if (e == cudaSuccess) {
      cap = want;
      void* d = nullptr;
      dptr = (cudaHostGetDevicePointer(&d, p, 0) == cudaSuccess) ? d : nullptr;
    }
    return e;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

struct TableKey {
  int q;
  uint64_t b[6];  // pose x,y,h; angle centre, offset, res (bit patterns)
  bool operator==(const TableKey& o) const { return q == o.q && memcmp(b, o.b, sizeof(b)) == 0; }
};
struct TableKeyHash {
  size_t operator()(const TableKey& k) const {
    uint64_t h = 1469598103934665603ull ^ (uint64_t)k.q;
    for (int i = 0; i < 6; i++) {
      h ^= k.b[i];
      h *= 1099511628211ull;
    }
    return (size_t)h;
  }
};
inline uint64_t dbits(double d) {
  uint64_t u;
  memcpy(&u, &d, 8);
  return u;
}

// host state of one match while its passes are scheduled
struct MatchState {
  int idx;        // index in the batch
  int slot;       // grid slot == index in the wave
  int q;          // pool scan id of the query
  int P;          // query point readings
  double pose[3];
  double gox, goy;
  int stage;      // 0 coarse, 1..3 expansion k, 4 fine, 5 done
  double angle_offset_cur;
  double mean[3];
  double cov[9];
  double best;
  int n_passes, n_ties;
  int status;
  // current pass
  int pass_id;
};

struct PassHost {
  int match;  // index into wave states
  bool fine;
  double cx, cy, ch, offx, offy, resx, resy, angle_offset, angle_res;
  int nA, nX, nY;
  int ang_off;
  bool spec;  // speculative fine pass resolved on the device
};

}  // namespace

// --------------------------------------------------------------------------------------------
// One iteration of CorrelateScan passes over the wave (host-side description).
namespace {

struct PassPlan {
  std::vector<TableDev> tab;
  std::vector<PassDev> pass;
  std::vector<PassHost> ph;
  std::vector<PassAngle> pa;
  std::vector<int> fine;       // ids of the fine passes scheduled by the host
  std::vector<int> spec_fine;  // ids of the speculative fine passes (latency path)
  std::vector<double> trig;
  std::vector<int> spec_of;    // coarse pass id -> its speculative fine pass id (-1)
  std::vector<int> spec_h;     // coarse pass id -> index of its heading table in trig (doubles)
  std::unordered_map<TableKey, int, TableKeyHash> tab_index;
  size_t off_elems = 0, sums_elems = 0, cmax_elems = 0;
  int ang_elems = 0;
  int max_lat_P = 0, max_lat_nx = 0, max_lat_ny = 0, max_lat_tasks = 0, max_fine_poses = 0;
  bool fine_not9 = false;  // some fine pass is not a 3 x 3 lattice (k_sweep_points instead of k_sweep_fine9)
  int first_spec_table = -1;
  void clear() {
    tab.clear(); pass.clear(); ph.clear(); pa.clear(); fine.clear(); spec_fine.clear(); trig.clear();
    spec_of.clear(); spec_h.clear(); tab_index.clear();
    off_elems = sums_elems = cmax_elems = 0;
    ang_elems = 0;
    max_lat_P = max_lat_nx = max_lat_ny = max_lat_tasks = max_fine_poses = 0;
    fine_not9 = false;
    first_spec_table = -1;
  }
};

// byte layout of one staging blob (host pinned mirror == device copy)
struct BlobLayout {
  size_t pool = 0, scan_start = 0, scan_count = 0, matches = 0, base = 0, workcount = 0, scanlist = 0;  // wave-static part
  size_t tab = 0, pass = 0, pa = 0, fine = 0, trig = 0, pmax = 0, total = 0;
};

inline size_t a16(size_t v) { return (v + 15) / 16 * 16; }

}  // namespace


// Host state of the resident latency kernel of one handle (ysm_resident.cuh).
struct Resident {
  bool alive = false;        // launched and not known to have left the device
  cudaStream_t st = nullptr; // non-blocking stream the kernel runs on
  unsigned char* mb = nullptr;    // mapped host memory: doorbell | exit line | spec | result chunks | request | points
  unsigned char* mb_dev = nullptr;
  size_t o_db = 0, o_exit = 0, o_prof = 0, o_spec = 0, o_out = 0, o_req = 0, o_pts = 0, mb_bytes = 0;
  unsigned char* d_small = nullptr;  // barrier counters | quit_round | abort | passmax | winner word (2 KB)
  unsigned* d_fsum = nullptr;        // fine lookup sums
  double* d_spec = nullptr;          // device copy of the spec tables
  double* d_dp = nullptr;            // distance-penalty tables: coarse lattice [nY][nX], then the fine 3 x 3
  size_t dp_fine_off = 0;
  int* d_winrec = nullptr;           // winner record CTA 0 -> fine-pass workers
  unsigned char* d_ctl = nullptr;
  uint32_t* d_cells = nullptr;
  double* d_qpts = nullptr;
  double* d_resp = nullptr;
  unsigned long long* d_cellmax = nullptr;
  size_t resp_cap = 0, cellmax_cap = 0;
  // device-resident scan store (ysm_batch::scan_tag): slot -> content tag
  double* d_cache = nullptr;
  struct Slot { uint64_t tag = 0; int count = 0; uint64_t last_use = 0; bool valid = false; };
  Slot slots[YSM_RES_CACHE_SLOTS];
  uint64_t use_clock = 0;
  unsigned seq = 0;
  size_t smem = 0;           // dynamic shared memory of the running instance
  size_t scratch = 0;        // ... of which worker scratch (fixes where the lookup offsets start)
  int G = 0;
  unsigned long long idle_ns = 0;
  int64_t served = 0, launches = 0;
};
#define YSM_RES_PTS_CAP 131072   // points the mailbox holds
#define YSM_RES_CELLS_CAP 131072
#define YSM_RES_SPEC_DOUBLES (YSM_RES_MAXNA + 5 * 4096)

struct ysm_handle {
  ysm_params prm;
  int device = 0;
  Resident* res = nullptr;
  bool res_enabled = true;
  size_t res_smem_limit = 0;
  GridC g;
  PenaltyC pen;
  int side = 0, margin = 0;
  double res_eff = 0.0;
  int slots = 0;
  uint8_t* d_grids = nullptr;
  uint32_t* d_rowmask = nullptr;  // [slots][tny*tnx]: which rows of every 32x32 tile hold a non-zero cell
  int tnx = 0, rm_words = 0;
  uint8_t* d_kernel = nullptr;
  uint16_t* d_stamp_tab = nullptr;  // pre-shifted stamp rows of k_tile_stamp: u16 [8][K][Wt]
  double* d_dpc = nullptr;          // coarse distance-penalty table [nY][nX] of k_sweep_pruned (made at the first batch)
  int dpc_nx = 0;
  std::vector<uint8_t> h_kernel;
  std::string err;
  int debug = 0;
  int64_t launches = 0;
  int num_sms = 148;
  // device workspaces
  DevBuf d_pool, d_scan_start, d_scan_count, d_base_idx, d_matches, d_cells, d_ptcell, d_cellcount;
  DevBuf d_gbox, d_work, d_workcount;
  DevBuf d_tables, d_passes, d_palist, d_fineids, d_trig, d_offsets, d_sums, d_outs, d_angsums, d_blob;
  DevBuf d_wblob, d_cellmax, d_scan_emit, d_tileflag, d_cand, d_wcand, d_isums;
  PinBuf h_blob, h_wblob, h_outs, h_angsums, h_flags;
  int epoch = 0;           // completion-flag value of the current latency-kernel launch
  size_t mega_occ_smem = ~(size_t)0;  // dynamic shared memory size mega_ctas_per_sm was computed for
  int mega_ctas_per_sm = 0;
  PassPlan plan;
  const MatchDev* cur_matches = nullptr;  // device views of the last wave (deferred clear)
  int* cur_workcount = nullptr;
  // last-batch debug info
  std::vector<int> last_slot_of_match;    // match -> slot (only for the last wave)
  std::vector<int> last_coarse_table_off; // match -> offsets element offset of its first coarse table
  std::vector<int> last_coarse_nA, last_coarse_P, last_coarse_Ppad;
  int last_wave_begin = 0, last_wave_end = 0;
  bool grids_dirty = false;
  // timing
  cudaEvent_t ev[8];
  bool ev_ok = false;
  double t_sweep = 0, t_build = 0, t_reduce = 0, t_total = 0;
  // lanes: extra matcher instances (own grid slots, workspaces and stream) that take contiguous
  // shares of a large batch on their own host threads, so one lane's host-side pass planning and
  // result handling overlap the other lanes' kernels
  std::vector<ysm_handle*> lanes;
  cudaStream_t lane_stream = nullptr;
  cudaEvent_t lane_event = nullptr;
  DevBuf d_pool_shared;
  bool ordered_stamps = false;  // wide smear: AddScan's skip rule is order dependent
  // match-against-map handles (ysm_create_map): ONE resident correlation grid built from a map image,
  // every match of a batch sweeps slot 0; nothing is built or cleared per match
  bool static_grid = false;
  double map_ox = 0.0, map_oy = 0.0;
  // lanes with a host-resident pool: the pool is uploaded wave by wave on the caller's stream; a lane
  // waits, before wave [lo, hi) of its share, for the event recorded behind that wave's scans
  struct SliceWait { int lo, hi; cudaEvent_t ev; bool waited; int seq; std::atomic<int>* seq_done; };
  std::vector<SliceWait> slice_wait;
  std::vector<cudaEvent_t> slice_events;  // pool of events (main handle)
  cudaStream_t upload_stream = nullptr;   // the uploader thread's stream (main handle)
  std::atomic<int> upload_seq{0};         // slices whose event the uploader has recorded so far
  int64_t work[16] = {0};
  unsigned long long* d_issued = nullptr;  // device counter: lookups the pruned sweep really issued
};

static void res_free(ysm_handle* h);

#define CK(call)                                                                     \
  do {                                                                               \
    cudaError_t _e = (call);                                                         \
    if (_e != cudaSuccess) {                                                         \
      h->err = std::string(#call) + ": " + cudaGetErrorString(_e);                   \
      return YSM_ECUDA;                                                              \
    }                                                                                \
  } while (0)

static int fail(ysm_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  else g_create_error = msg;
  return code;
}

extern "C" const char* ysm_last_error(const ysm_handle* h) {
  return h ? h->err.c_str() : g_create_error.c_str();
}

// CorrelationGrid::CalculateKernel (python twin yag_slam/helpers.py:86-97)
static int calculate_kernel(double res_eff, double smear, std::vector<uint8_t>& out, int& half, int& K) {
  if (!(smear >= 0.5 * res_eff && smear <= 10 * res_eff)) return -1;
  half = (int)h_round(2.0 * smear / res_eff);
  K = 2 * half + 1;
  out.assign((size_t)K * K, 0);
  const int hk = K / 2;
  for (int i = -hk; i <= hk; i++) {
    for (int j = -hk; j <= hk; j++) {
      double d = hypot(i * res_eff, j * res_eff);
      double z = exp(-0.5 * pow(d / smear, 2));
      uint32_t kv = (uint32_t)h_round(z * 100);
      out[(size_t)(i + hk) + (size_t)K * (j + hk)] = (uint8_t)kv;
    }
  }
  return 0;
}

// Stamp rows pre-shifted to the 8 cell alignments of a 16-byte shared-memory load (k_tile_stamp):
// row j of alignment a holds kernel row j at cells [24 + a, 24 + a + K) of a Wt-cell row of zeros.
static void build_stamp_table(const std::vector<uint8_t>& kern, int K, int& Wt, std::vector<uint16_t>& tab) {
  Wt = K + 62;
  while (Wt % 16 != 8) Wt++;  // row stride = 16 B (mod 32 B): consecutive rows on distinct banks
  tab.assign((size_t)8 * K * Wt, 0);
  for (int a = 0; a < 8; a++)
    for (int j = 0; j < K; j++)
      for (int i = 0; i < K; i++) tab[((size_t)a * K + j) * Wt + 24 + a + i] = kern[(size_t)i + (size_t)K * j];
}

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the (device, kernel) pair, not to a matcher handle:
// it is raised ONCE per device to the opt-in maximum for every kernel that takes dynamic shared memory and
// never lowered, so handles (and lanes on concurrent host threads) cannot undercut each other.
static std::mutex g_attr_mu;
static bool g_attr_done[64] = {false};
static size_t g_res_smem_limit[64] = {0};

static cudaError_t init_kernel_attrs(int device) {
  std::lock_guard<std::mutex> lk(g_attr_mu);
  if (device < 0 || device >= 64) return cudaErrorInvalidDevice;
  if (g_attr_done[device]) return cudaSuccess;
  int optin = 0;
  cudaError_t e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  if (e != cudaSuccess) return e;
  const void* fns[] = {(const void*)k_find_valid,    (const void*)k_stamp_order, (const void*)k_tile_stamp,
                       (const void*)k_tile_stamp_lists,
                       (const void*)k_sweep_pruned,  (const void*)k_sweep_lattice, (const void*)k_match_small,
                       (const void*)k_match_resident};
  for (const void* fn : fns) {
    cudaFuncAttributes fa;
    if ((e = cudaFuncGetAttributes(&fa, fn)) != cudaSuccess) return e;
    const int dyn = optin - (int)fa.sharedSizeBytes;
    if ((e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn)) != cudaSuccess) return e;
    if (fn == (const void*)k_match_resident) g_res_smem_limit[device] = dyn > 0 ? (size_t)dyn : 0;
  }
  g_attr_done[device] = true;
  return cudaSuccess;
}

static int debug_env_or();
static int create_one(const ysm_params* p, int device, ysm_handle** out, int roi_override = 0) {
  if (!p || !out) return fail(nullptr, YSM_EINVAL, "null argument");
  *out = nullptr;
  if (!(p->resolution > 0) || !(p->search_size > 0) || p->smear_deviation < 0 || !(p->range_threshold > 0))
    return fail(nullptr, YSM_EINVAL, "invalid matcher parameters");
  if (!(p->coarse_angle_resolution > 0) || !(p->fine_search_angle_resolution > 0) ||
      !(p->coarse_search_angle_offset > 0))
    return fail(nullptr, YSM_EINVAL, "angle offsets/resolutions must be positive");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0)
    return fail(nullptr, YSM_ECUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(nullptr, YSM_EINVAL, "bad device index");
  if ((e = cudaSetDevice(device)) != cudaSuccess)
    return fail(nullptr, YSM_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));

  if ((e = init_kernel_attrs(device)) != cudaSuccess)
    return fail(nullptr, YSM_ECUDA, std::string("kernel attributes: ") + cudaGetErrorString(e));
  ysm_quiesce_device(device);  // (allocations below must not wait for another handle's resident kernel)

  ysm_handle* h = new ysm_handle();
  h->prm = *p;
  h->device = device;
  {
    static const bool no_res = getenv("YSM_NO_RESIDENT") != nullptr;
    h->res_enabled = !no_res && p->resident_idle_us >= 0;
    // the sweep reads the grid through L1: leave it at least a third of the SM's 228 KB
    h->res_smem_limit = std::min<size_t>(g_res_smem_limit[device], 144 * 1024);
  }
  // ScanMatcher::Create sizing
  h->side = (int)(uint32_t)(h_round(p->search_size / p->resolution) + 1);
  h->margin = (int)(uint32_t)ceil(p->range_threshold / p->resolution);
  GridC& g = h->g;
  g.roi = roi_override > 0 ? roi_override : h->side + 2 * h->margin;
  g.border = (int)h_round(2.0 * p->smear_deviation / p->resolution) + 1;
  g.width = g.roi + 2 * g.border;
  g.height = g.width;
  g.stride = align_up(g.width, 8);
  g.stride4 = g.stride / 4;
  g.scale = 1.0 / p->resolution;
  h->res_eff = 1.0 / g.scale;
  const long long dsz = (long long)g.stride * g.height;
  if (dsz > 0x7fffffffLL || g.width > 65535) {
    delete h;
    return fail(nullptr, YSM_EUNSUP, "correlation grid too large (>= 2 GiB or > 65535 cells wide)");
  }
  g.data_size = (int)dsz;
  g.grid_bytes = (dsz + 255) / 256 * 256;
  if (calculate_kernel(h->res_eff, p->smear_deviation, h->h_kernel, g.half_kernel, g.K) != 0) {
    char buf[160];
    snprintf(buf, sizeof(buf), "Mapper Error:  Smear deviation too small:  Must be between %g and %g",
             0.5 * h->res_eff, 10 * h->res_eff);
    delete h;
    return fail(nullptr, YSM_EINVAL, buf);
  }
  if (g.half_kernel + 1 > g.border) {
    delete h;
    return fail(nullptr, YSM_EUNSUP, "kernel larger than the grid border");
  }
  // The parallel smear is order-independent only while the stamp is 100 at its centre alone. For
  // smear_deviation >= ~9.99 * resolution its four edge neighbours are 100 as well and Karto's
  // "cell already occupied -> skip" test makes the set of stamps depend on the point order:
  // k_stamp_order replays that order before the stamping kernel.
  {
    int hot = 0;
    for (uint8_t v : h->h_kernel) hot += (v == 100);
    const int hk = g.half_kernel, K = g.K;
    const bool plus = hot == 5 && h->h_kernel[(size_t)(hk - 1) + (size_t)K * hk] == 100 &&
                      h->h_kernel[(size_t)(hk + 1) + (size_t)K * hk] == 100 &&
                      h->h_kernel[(size_t)hk + (size_t)K * (hk - 1)] == 100 &&
                      h->h_kernel[(size_t)hk + (size_t)K * (hk + 1)] == 100;
    if (hot != 1 && !plus) {
      delete h;
      return fail(nullptr, YSM_EUNSUP, "smear kernel with an unexpected set of 100-valued taps");
    }
    h->ordered_stamps = (hot == 5);
  }
  std::vector<uint16_t> stamp_tab;
  build_stamp_table(h->h_kernel, g.K, g.Wt, stamp_tab);
  h->pen.distance_variance_penalty = p->distance_variance_penalty;
  h->pen.angle_variance_penalty = p->angle_variance_penalty;
  h->pen.minimum_distance_penalty = p->minimum_distance_penalty;
  h->pen.minimum_angle_penalty = p->minimum_angle_penalty;

  long long budget = p->max_grid_bytes > 0 ? p->max_grid_bytes : (16LL << 30);
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  if ((long long)free_b / 2 < budget) budget = (long long)free_b / 2;
  long long slots = budget / g.grid_bytes;
  if (p->max_slots > 0 && slots > p->max_slots) slots = p->max_slots;
  if (slots > 8192) slots = 8192;
  {
    // developer A/B: waves a multiple of YSM_WAVE_ALIGN matches (k_find_valid runs one CTA per match)
    static const int wave_align = getenv("YSM_WAVE_ALIGN") ? atoi(getenv("YSM_WAVE_ALIGN")) : 0;
    if (wave_align > 0 && slots > wave_align) slots -= slots % wave_align;
  }
  if (slots < 1) {
    delete h;
    return fail(nullptr, YSM_ENOMEM, "not enough device memory for one correlation grid");
  }
  h->slots = (int)slots;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->num_sms = prop.multiProcessorCount;
  h->tnx = (g.width + YSM_TILE - 1) / YSM_TILE;
  h->rm_words = h->tnx * h->tnx;
  e = cudaMalloc((void**)&h->d_grids, (size_t)slots * g.grid_bytes);
  if (e == cudaSuccess) e = cudaMemset(h->d_grids, 0, (size_t)slots * g.grid_bytes);
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_rowmask, (size_t)slots * h->rm_words * 4);
  if (e == cudaSuccess) e = cudaMemset(h->d_rowmask, 0, (size_t)slots * h->rm_words * 4);
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_issued, 8);
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_kernel, h->h_kernel.size());
  if (e == cudaSuccess) e = cudaMemcpy(h->d_kernel, h->h_kernel.data(), h->h_kernel.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_stamp_tab, stamp_tab.size() * 2);
  if (e == cudaSuccess) e = cudaMemcpy(h->d_stamp_tab, stamp_tab.data(), stamp_tab.size() * 2, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    h->ev_ok = true;
    for (int i = 0; i < 8; i++)
      if (cudaEventCreate(&h->ev[i]) != cudaSuccess) h->ev_ok = false;
  }
  if (e != cudaSuccess) {
    std::string msg = std::string("device allocation failed: ") + cudaGetErrorString(e);
    if (h->d_grids) cudaFree(h->d_grids);
    if (h->d_rowmask) cudaFree(h->d_rowmask);
    if (h->d_issued) cudaFree(h->d_issued);
    if (h->d_kernel) cudaFree(h->d_kernel);
    if (h->d_stamp_tab) cudaFree(h->d_stamp_tab);
    delete h;
    return fail(nullptr, YSM_ECUDA, msg);
  }
  h->debug = debug_env_or();
  *out = h;
  return YSM_OK;
}

extern "C" void ysm_destroy(ysm_handle* h);

extern "C" int ysm_create(const ysm_params* p, int device, ysm_handle** out) {
  if (!p || !out) return fail(nullptr, YSM_EINVAL, "null argument");
  *out = nullptr;
  int lanes = std::max(1, std::min(4, (int)p->lanes));
  if (p->max_slots > 0 && p->max_slots < 64 * lanes) lanes = 1;  // small handles (tests, latency use) stay single
  ysm_params q = *p;
  if (lanes > 1) {
    q.max_grid_bytes = (p->max_grid_bytes > 0 ? p->max_grid_bytes : (16LL << 30)) / lanes;
    if (p->max_slots > 0) q.max_slots = p->max_slots / lanes;
  }
  ysm_handle* h = nullptr;
  int rc = create_one(&q, device, &h);
  if (rc != YSM_OK) return rc;
  for (int l = 1; l < lanes; l++) {
    ysm_handle* sub = nullptr;
    rc = create_one(&q, device, &sub);
    if (rc != YSM_OK) {
      ysm_destroy(h);
      return rc;
    }
    h->lanes.push_back(sub);
  }
  if (lanes > 1) {
    cudaError_t e = cudaStreamCreateWithFlags(&h->lane_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->lane_event, cudaEventDisableTiming);
    for (ysm_handle* sub : h->lanes)
      if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&sub->lane_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      ysm_destroy(h);
      return fail(nullptr, YSM_ECUDA, std::string("lane stream creation failed: ") + cudaGetErrorString(e));
    }
  }
  *out = h;
  return YSM_OK;
}

// ---- match against a map image ----------------------------------------------------
// occupancy_grid_map_to_correlation_grid (reference yag_slam/helpers.py:24-34) with Karto's grid
// semantics: image cell (u, v) is ROI cell (u, v); cells equal to occupied_value become 100 and every
// one of them is smeared with the kernel (pure max, no skip rule -- order independent). Gather form:
// one thread per grid cell takes the max over the K x K window of source cells.
__global__ void __launch_bounds__(256)
k_map_smear(GridC g, const uint8_t* __restrict__ img, int ih, int iw, int occ, const uint8_t* __restrict__ kern,
            uint8_t* __restrict__ grid) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= g.stride || y >= g.height) return;
  const int tx = x - g.border, ty = y - g.border, hk = g.half_kernel, K = g.K;
  unsigned v = 0;
  if (x < g.width) {
    for (int j = -hk; j <= hk; j++) {
      const int sv = ty - j;
      if (sv < 0 || sv >= ih) continue;
      const uint8_t* row = img + (size_t)sv * iw;
      for (int i = -hk; i <= hk; i++) {
        const int su = tx - i;
        if (su < 0 || su >= iw) continue;
        if ((int)__ldg(row + su) == occ) v = max(v, (unsigned)__ldg(kern + (i + hk) + K * (j + hk)));
      }
    }
  }
  grid[(size_t)y * g.stride + x] = (uint8_t)v;
}

// row masks of the resident grid (one word per 32 x 32 tile: which tile rows hold a non-zero cell) --
// what k_tile_stamp leaves for scan-built grids; the pruned sweep reads them
__global__ void __launch_bounds__(256)
k_map_rowmask(GridC g, const uint8_t* __restrict__ grid, uint32_t* __restrict__ rowmask, int tnx, int ntiles) {
  const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (tile >= ntiles) return;
  const int ty = tile / tnx, tx = tile - ty * tnx;
  const int row = ty * YSM_TILE + lane;
  bool nz = false;
  if (row < g.height) {
    for (int k = 0; k < YSM_TILE; k++) {
      const int x = tx * YSM_TILE + k;
      if (x < g.stride && grid[(size_t)row * g.stride + x] != 0) nz = true;
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, nz);
  if (lane == 0) rowmask[tile] = m;
}

extern "C" int ysm_create_map(const ysm_params* p, const uint8_t* img, int32_t ih, int32_t iw, int32_t occupied_value,
                              double offset_x, double offset_y, int device, ysm_handle** out) {
  if (!p || !out) return fail(nullptr, YSM_EINVAL, "null argument");
  *out = nullptr;
  if (!img || ih <= 0 || iw <= 0) return fail(nullptr, YSM_EINVAL, "ysm_create_map: empty map image");
  ysm_params q = *p;
  q.max_slots = 1;
  q.lanes = 1;
  ysm_handle* h = nullptr;
  int rc = create_one(&q, device, &h, std::max(ih, iw));
  if (rc != YSM_OK) return rc;
  h->static_grid = true;
  h->ordered_stamps = false;  // every occupied cell is smeared: no order dependence in map mode
  h->map_ox = offset_x;
  h->map_oy = offset_y;
  const GridC& g = h->g;
  uint8_t* d_img = nullptr;
  cudaError_t e = cudaMalloc((void**)&d_img, (size_t)ih * iw);
  if (e == cudaSuccess) e = cudaMemcpy(d_img, img, (size_t)ih * iw, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    dim3 grid((g.stride + 255) / 256, g.height);
    k_map_smear<<<grid, 256>>>(g, d_img, ih, iw, occupied_value, h->d_kernel, h->d_grids);
    const int ntiles = h->tnx * h->tnx;
    k_map_rowmask<<<(ntiles + 7) / 8, 256>>>(g, h->d_grids, h->d_rowmask, h->tnx, ntiles);
    h->launches += 2;
    e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaGetLastError();
  }
  if (d_img) cudaFree(d_img);
  if (e != cudaSuccess) {
    std::string msg = std::string("ysm_create_map: ") + cudaGetErrorString(e);
    ysm_destroy(h);
    return fail(nullptr, YSM_ECUDA, msg);
  }
  *out = h;
  return YSM_OK;
}

extern "C" void ysm_destroy(ysm_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  res_free(h);
  ysm_quiesce_device(h->device);
  cudaDeviceSynchronize();
  for (ysm_handle* sub : h->lanes) ysm_destroy(sub);
  h->lanes.clear();
  if (h->lane_stream) cudaStreamDestroy(h->lane_stream);
  if (h->lane_event) cudaEventDestroy(h->lane_event);
  for (cudaEvent_t ev : h->slice_events) cudaEventDestroy(ev);
  if (h->upload_stream) cudaStreamDestroy(h->upload_stream);
  h->d_pool_shared.release();
  if (h->d_grids) cudaFree(h->d_grids);
  if (h->d_rowmask) cudaFree(h->d_rowmask);
  if (h->d_issued) cudaFree(h->d_issued);
  if (h->d_kernel) cudaFree(h->d_kernel);
  if (h->d_stamp_tab) cudaFree(h->d_stamp_tab);
  if (h->d_dpc) cudaFree(h->d_dpc);
  DevBuf* bufs[] = {&h->d_pool, &h->d_scan_start, &h->d_scan_count, &h->d_base_idx, &h->d_matches,
                    &h->d_cells, &h->d_ptcell, &h->d_cellcount, &h->d_gbox, &h->d_work, &h->d_workcount,
                    &h->d_tables, &h->d_passes,
                    &h->d_palist, &h->d_fineids, &h->d_trig, &h->d_offsets, &h->d_sums, &h->d_outs,
                    &h->d_angsums, &h->d_blob, &h->d_wblob, &h->d_cellmax, &h->d_scan_emit,
                    &h->d_tileflag, &h->d_cand, &h->d_wcand, &h->d_isums};
  for (DevBuf* b : bufs) b->release();
  h->h_blob.release();
  h->h_wblob.release();
  h->h_flags.release();
  h->h_outs.release();
  h->h_angsums.release();
  if (h->ev_ok)
    for (int i = 0; i < 8; i++) cudaEventDestroy(h->ev[i]);
  delete h;
}

extern "C" int ysm_get_dims(const ysm_handle* h, ysm_dims* out) {
  if (!h || !out) return YSM_EINVAL;
  out->side = h->side; out->margin = h->margin; out->roi = h->g.roi;
  out->half_kernel = h->g.half_kernel; out->kernel_size = h->g.K; out->border = h->g.border;
  out->width = h->g.width; out->height = h->g.height; out->stride = h->g.stride;
  out->slots = h->slots; out->grid_bytes = h->g.data_size;
  for (const ysm_handle* sub : h->lanes) out->slots += sub->slots;
  return YSM_OK;
}

static int debug_env_or() {
  static const int v = getenv("YSM_DEBUG_OR") ? atoi(getenv("YSM_DEBUG_OR")) : 0;
  return v;
}

extern "C" int ysm_set_debug(ysm_handle* h, int32_t flags) {
  if (!h) return YSM_EINVAL;
  // YSM_DEBUG_OR: developer A/B switch (read once), ORed into every set_debug / create: e.g. 64 | 128
  flags |= debug_env_or();
  h->debug = flags;
  for (ysm_handle* sub : h->lanes) sub->debug = flags;
  return YSM_OK;
}

extern "C" int64_t ysm_launch_count(const ysm_handle* h) {
  if (!h) return 0;
  int64_t n = h->launches;
  for (const ysm_handle* sub : h->lanes) n += sub->launches;
  return n;
}

extern "C" int ysm_last_work(const ysm_handle* h, int64_t* out, int32_t n) {
  if (!h || !out || n < 0) return YSM_EINVAL;
  for (int i = 0; i < n; i++) out[i] = i < 16 ? h->work[i] : 0;
  return YSM_OK;
}

extern "C" int ysm_last_kernel_ms(const ysm_handle* h, double* sweep_ms, double* build_ms,
                                  double* reduce_ms, double* total_ms) {
  if (!h) return YSM_EINVAL;
  if (sweep_ms) *sweep_ms = h->t_sweep;
  if (build_ms) *build_ms = h->t_build;
  if (reduce_ms) *reduce_ms = h->t_reduce;
  if (total_ms) *total_ms = h->t_total;
  return YSM_OK;
}

// LocalizedRangeScan::Update
extern "C" int ysm_point_readings(const double* ranges, int32_t n, double min_angle,
                                  double angular_resolution, double min_range,
                                  double range_threshold, double x, double y, double heading,
                                  double* out_xy, int32_t* n_out) {
  if (!ranges || !out_xy || !n_out || n < 0) return YSM_EINVAL;
  int k = 0;
  for (int i = 0; i < n; i++) {
    const double r = ranges[i];
    if (!(r >= min_range && r <= range_threshold)) continue;
    const double angle = heading + min_angle + (double)(uint32_t)i * angular_resolution;
    out_xy[2 * k] = x + (r * cos(angle));
    out_xy[2 * k + 1] = y + (r * sin(angle));
    k++;
  }
  *n_out = k;
  return YSM_OK;
}

// LocalizedRangeScan::Update for a batch of scans: the same libm loop, spread over host threads
extern "C" int ysm_point_readings_batch(const double* ranges, const int32_t* beam_ptr, int32_t n_src, const int32_t* src,
                                        const double* pose, int32_t n, double min_angle, double angular_resolution,
                                        double min_range, double range_threshold, double* out_xy, int32_t* out_start,
                                        int32_t* out_count, int64_t* n_points) {
  if (!ranges || !beam_ptr || !src || !pose || !out_xy || !out_start || !out_count || n < 0 || n_src < 0) return YSM_EINVAL;
  for (int i = 0; i < n; i++)
    if (src[i] < 0 || src[i] >= n_src) return YSM_EINVAL;
  try {
    // pass 1: how many readings every scan keeps (range test only), then the offsets
    int64_t tot = 0;
    for (int i = 0; i < n; i++) {
      int k = 0;
      for (int j = beam_ptr[src[i]]; j < beam_ptr[src[i] + 1]; j++) k += (ranges[j] >= min_range && ranges[j] <= range_threshold);
      if (tot > 0x7fffffffLL - k) return YSM_EUNSUP;
      out_start[i] = (int32_t)tot;
      out_count[i] = k;
      tot += k;
    }
    if (n_points) *n_points = tot;
    // pass 2: the readings, scans dealt to the host threads in contiguous blocks
    int nthreads = 1;
    {
      cpu_set_t set;
      CPU_ZERO(&set);
      if (sched_getaffinity(0, sizeof(set), &set) == 0) nthreads = std::max(1, CPU_COUNT(&set));
      nthreads = std::min(nthreads, std::max(1, n / 64));
    }
    auto work = [&](int lo, int hi) {
      for (int i = lo; i < hi; i++) {
        const double x = pose[3 * (size_t)i], y = pose[3 * (size_t)i + 1], heading = pose[3 * (size_t)i + 2];
        const int b0 = beam_ptr[src[i]], nb = beam_ptr[src[i] + 1] - b0;
        double* o = out_xy + 2 * (size_t)out_start[i];
        for (int j = 0; j < nb; j++) {
          const double r = ranges[b0 + j];
          if (!(r >= min_range && r <= range_threshold)) continue;
          const double angle = heading + min_angle + (double)(uint32_t)j * angular_resolution;
          o[0] = x + (r * cos(angle));
          o[1] = y + (r * sin(angle));
          o += 2;
        }
      }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; t++) {
      const int lo = (int)((long long)n * t / nthreads), hi = (int)((long long)n * (t + 1) / nthreads);
      try {
        th.emplace_back(work, lo, hi);
      } catch (const std::exception&) {
        work(lo, hi);
      }
    }
    work(0, (int)((long long)n / nthreads));
    for (std::thread& t : th) t.join();
  } catch (const std::exception&) {
    return YSM_ENOMEM;
  }
  return YSM_OK;
}

// --------------------------------------------------------------------------------------------
// zero the tiles of the last built wave (its work list is still resident) -- the grids
// return to all-zero
static int clear_wave(ysm_handle* h, const MatchDev* d_matches, cudaStream_t st) {
  k_tile_clear<<<h->num_sms * 8, 256, 0, st>>>(h->g, d_matches, (const int2*)h->d_work.p, h->cur_workcount, h->d_grids,
                                               h->d_rowmask, h->rm_words);
  h->launches++;
  return YSM_OK;
}

static void finalize_positional(const ysm_handle* h, const PassHost& ph, const PassOut& po, double* cov) {
  // ComputePositionalCovariance tail
  memset(cov, 0, 9 * sizeof(double));
  cov[0] = cov[4] = cov[8] = 1.0;
  const double best = po.best;
  if (best < KT_TOLERANCE) {
    cov[0] = MAX_VARIANCE;
    cov[4] = MAX_VARIANCE;
    cov[8] = 4 * h_square(ph.angle_res);
    return;
  }
  if (po.norm > KT_TOLERANCE) {
    double vxx = po.axx / po.norm, vxy = po.axy / po.norm, vyy = po.ayy / po.norm;
    const double vthth = 4 * h_square(ph.angle_res);
    const double min_vxx = 0.1 * h_square(ph.resx);
    const double min_vyy = 0.1 * h_square(ph.resy);
    vxx = vxx > min_vxx ? vxx : min_vxx;
    vyy = vyy > min_vyy ? vyy : min_vyy;
    const double mult = 1.0 / best;
    cov[0] = vxx * mult;
    cov[1] = vxy * mult;
    cov[3] = vxy * mult;
    cov[4] = vyy * mult;
    cov[8] = vthth;
  }
  if (h_double_equal(cov[0], 0.0)) cov[0] = MAX_VARIANCE;
  if (h_double_equal(cov[4], 0.0)) cov[4] = MAX_VARIANCE;
}

static void finalize_angular(const PassHost& ph, const PassOut& po, double heading, const int* angsums,
                             int P, double* cov) {
  // ComputeAngularCovariance; the per-angle GetResponse sums come from the GPU
  const double best_angle = h_normalize_angle_difference(heading, ph.ch);
  const double start_angle = ph.ch - ph.angle_offset;
  double norm = 0.0, acc = 0.0;
  const double denom = (double)((uint32_t)P * 100u);
  for (int a = 0; a < ph.nA; a++) {
    const double angle = start_angle + (double)(uint32_t)a * ph.angle_res;
    double response = (double)angsums[a];
    response /= denom;
    if (response >= (po.best - 0.1)) {
      norm += response;
      acc += (h_square(angle - best_angle) * response);
    }
  }
  if (norm > KT_TOLERANCE) {
    if (acc < KT_TOLERANCE) acc = h_square(ph.angle_res);
    acc /= norm;
  } else {
    acc = 1000 * h_square(ph.angle_res);
  }
  cov[8] = acc;
}

// Matches per wave: the batch split into the fewest waves the slots allow, of equal size -- a runt last wave
// (12 waves of 344 and one of 39) costs nearly a whole wave's pipeline latency (r02ze: ~2 ms per step at N = 8).
static inline int balanced_wave(int n_matches, int slots) {
  if (slots <= 0 || n_matches <= slots) return std::max(1, slots);
  const int waves = (n_matches + slots - 1) / slots;
  return (n_matches + waves - 1) / waves;
}
static inline int n_steps(double off, double res) { return (int)(uint32_t)(h_round(off * 2.0 / res) + 1); }

static void fill_inverse_rotation(TableDev& t, const double* pose) {
  // Transform(sensorPose): m_InverseRotation = FromAxisAngle(0,0,1, 0 - heading)
  if (pose[0] == 0.0 && pose[1] == 0.0 && pose[2] == 0.0) {
    t.r00 = 1.0; t.r01 = 0.0; t.r10 = 0.0; t.r11 = 1.0;
  } else {
    const double radians = 0.0 - pose[2];
    const double c = cos(radians), sn = sin(radians);
    const double omc = 1.0 - c;
    t.r00 = 0.0 * omc + c;
    t.r01 = (0.0 * 0.0 * omc) - (1.0 * sn);
    t.r10 = (0.0 * 0.0 * omc) + (1.0 * sn);
    t.r11 = 0.0 * omc + c;
  }
}

// --------------------------------------------------------------------------------------------
// Resident latency path, host side (kernel: ysm_resident.cuh). One resident kernel per device at a
// time; g_res_owner[device] is the handle it belongs to.
static std::mutex g_res_mu;
static ysm_handle* g_res_owner[64] = {nullptr};
static std::atomic<int> g_res_alive{0};

// One tagged 16-byte word of a doorbell line: the word that carries the tag is stored last (x86 keeps the store
// order; the GPU reads the 16 bytes with one request and accepts a line only when all its tags agree).
static inline void res_put(unsigned char* line, int k, unsigned a, unsigned b, unsigned c, unsigned d, int tag_pos) {
  uint32_t* w = reinterpret_cast<uint32_t*>(line + 16 * k);
  const unsigned v[4] = {a, b, c, d};
  for (int i = 0; i < 4; i++)
    if (i != tag_pos) __atomic_store_n(&w[i], v[i], __ATOMIC_RELAXED);
  __atomic_store_n(&w[tag_pos], v[tag_pos], __ATOMIC_RELEASE);
}

struct ResScanInfo { int count, src_slot, store_slot; };  // slots are 1-based, 0 = none

// Rings the doorbell of every polling CTA: v0 = {seq, w1, w2, w3}; CTA 1 + s also learns about scan s
// (scans[s], s <= nbase, the last one being the query) and every line carries the viewpoint / grid offset m[4].
static void res_ring(Resident& R, unsigned seq, unsigned w1, unsigned w2, unsigned w3, const ResScanInfo* scans = nullptr,
                     int nscans = 0, const double* m = nullptr) {
  uint32_t mw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (m) memcpy(mw, m, 32);
  for (int k = YSM_RES_POLLERS; k >= 0; k--) {
    unsigned char* line = R.mb + R.o_db + (size_t)k * YSM_RES_DB_STRIDE;
    ResScanInfo si = {0, 0, 0};
    if (k >= 1 && k - 1 < nscans) si = scans[k - 1];
    res_put(line, 4, mw[6], mw[7], 0u, seq, 3);
    res_put(line, 3, mw[3], mw[4], mw[5], seq, 3);
    res_put(line, 2, mw[0], mw[1], mw[2], seq, 3);
    res_put(line, 1, seq, (unsigned)si.count, (unsigned)si.src_slot, (unsigned)si.store_slot, 0);
    res_put(line, 0, seq, w1, w2, w3, 0);
  }
}

// (g_res_mu held) ends the handle's resident kernel and waits until it has left the device
static int res_stop_locked(ysm_handle* h) {
  Resident* R = h->res;
  if (!R || !R->alive) return YSM_OK;
  int cur = -1;
  cudaGetDevice(&cur);
  if (cur != h->device) cudaSetDevice(h->device);
  R->seq++;
  res_ring(*R, R->seq, RES_CMD_QUIT, 0u, 0u);
  const cudaError_t e = cudaStreamSynchronize(R->st);
  R->alive = false;
  g_res_alive.fetch_sub(1);
  if (g_res_owner[h->device] == h) g_res_owner[h->device] = nullptr;
  if (cur >= 0 && cur != h->device) cudaSetDevice(cur);
  return e == cudaSuccess ? YSM_OK : YSM_ECUDA;
}

void ysm_quiesce_device(int device) {
  if (device < 0 || device >= 64 || g_res_alive.load() == 0) return;
  std::lock_guard<std::mutex> lk(g_res_mu);
  if (g_res_owner[device]) res_stop_locked(g_res_owner[device]);
}

static void res_free(ysm_handle* h) {
  Resident* R = h->res;
  if (!R) return;
  {
    std::lock_guard<std::mutex> lk(g_res_mu);
    res_stop_locked(h);
  }
  if (R->st) cudaStreamDestroy(R->st);
  if (R->mb) cudaFreeHost(R->mb);
  if (R->d_small) cudaFree(R->d_small);
  if (R->d_fsum) cudaFree(R->d_fsum);
  if (R->d_spec) cudaFree(R->d_spec);
  if (R->d_dp) cudaFree(R->d_dp);
  if (R->d_winrec) cudaFree(R->d_winrec);
  if (R->d_ctl) cudaFree(R->d_ctl);
  if (R->d_cells) cudaFree(R->d_cells);
  if (R->d_qpts) cudaFree(R->d_qpts);
  if (R->d_cache) cudaFree(R->d_cache);
  if (R->d_resp) cudaFree(R->d_resp);
  if (R->d_cellmax) cudaFree(R->d_cellmax);
  delete R;
  h->res = nullptr;
}

// The two halves of CorrelateScan's odometry penalty, evaluated on the host with the same IEEE
// operations, in the same order, as the device functions penalty_distance / penalty_angle: plain arithmetic
// (no libm), so host and device agree bit for bit -- and f64 division is the most expensive thing an SM of this
// machine can be asked to do.
static double h_penalty_distance(double offx, double resx, double offy, double resy, int ix, int iy, const PenaltyC& pen) {
  const double x = -offx + (double)ix * resx;
  const double y = -offy + (double)iy * resy;
  const double sqd = x * x + y * y;
  const double dp = 1.0 - (0.2 * sqd / pen.distance_variance_penalty);
  return dp > pen.minimum_distance_penalty ? dp : pen.minimum_distance_penalty;
}
static double h_penalty_angle(double ch, double angle_offset, double angle_res, int a, const PenaltyC& pen) {
  const double angle = (ch - angle_offset) + (double)a * angle_res;
  const double da = angle - ch;
  const double sqa = da * da;
  const double ap = 1.0 - (0.2 * sqa / pen.angle_variance_penalty);
  return ap > pen.minimum_angle_penalty ? ap : pen.minimum_angle_penalty;
}

static int res_alloc(ysm_handle* h) {
  if (h->res) return YSM_OK;
  ysm_quiesce_device(h->device);  // cudaMalloc / cudaHostAlloc below must not wait for another handle's kernel
  Resident* R = new Resident();
  h->res = R;
  size_t o = 0;
  R->o_db = o; o += (size_t)(YSM_RES_POLLERS + 1) * YSM_RES_DB_STRIDE;
  R->o_exit = o; o += 128;
  R->o_prof = o; o += 8 * (size_t)YSM_RES_PROF * 256;
  R->o_spec = o; o += 64 + sizeof(double) * (size_t)YSM_RES_SPEC_DOUBLES;
  o = (o + 127) & ~(size_t)127;
  R->o_out = o; o += 16 * (size_t)YSM_RES_CHUNKS;
  R->o_req = o; o += (sizeof(ResReq) + 127) & ~(size_t)127;
  R->o_pts = o; o += 16 * (size_t)YSM_RES_PTS_CAP;
  R->mb_bytes = o;
  CK(cudaStreamCreateWithFlags(&R->st, cudaStreamNonBlocking));
  CK(cudaHostAlloc((void**)&R->mb, R->mb_bytes, cudaHostAllocMapped));
  memset(R->mb, 0, R->mb_bytes);
  void* d = nullptr;
  CK(cudaHostGetDevicePointer(&d, R->mb, 0));
  R->mb_dev = (unsigned char*)d;
  CK(cudaMalloc((void**)&R->d_small, 2048));
  CK(cudaMalloc((void**)&R->d_fsum, 4 * 4096));
  CK(cudaMalloc((void**)&R->d_spec, 8 * (size_t)YSM_RES_SPEC_DOUBLES));
  CK(cudaMalloc((void**)&R->d_ctl, sizeof(ResReq)));
  CK(cudaMalloc((void**)&R->d_cells, 4 * (size_t)YSM_RES_CELLS_CAP));
  CK(cudaMalloc((void**)&R->d_qpts, 16 * (size_t)YSM_RES_PMAX));
  CK(cudaMalloc((void**)&R->d_cache, 16 * (size_t)YSM_RES_PMAX * YSM_RES_CACHE_SLOTS));
  CK(cudaMalloc((void**)&R->d_winrec, 4096));
  {
    // distance-penalty tables: they depend on the matcher's configuration only
    const double csx = 0.5 * (h->side - 1) * h->res_eff, crx = 2 * h->res_eff;
    const int nX = n_steps(csx, crx), fnX = n_steps(crx * 0.5, h->res_eff);
    std::vector<double> dp((size_t)nX * nX + (size_t)fnX * fnX);
    for (int iy = 0; iy < nX; iy++)
      for (int ix = 0; ix < nX; ix++) dp[(size_t)iy * nX + ix] = h_penalty_distance(csx, crx, csx, crx, ix, iy, h->pen);
    R->dp_fine_off = (size_t)nX * nX;
    for (int iy = 0; iy < fnX; iy++)
      for (int ix = 0; ix < fnX; ix++)
        dp[R->dp_fine_off + (size_t)iy * fnX + ix] = h_penalty_distance(crx * 0.5, h->res_eff, crx * 0.5, h->res_eff, ix, iy, h->pen);
    CK(cudaMalloc((void**)&R->d_dp, dp.size() * 8));
    CK(cudaMemcpy(R->d_dp, dp.data(), dp.size() * 8, cudaMemcpyHostToDevice));
  }
  const int idle_us = h->prm.resident_idle_us > 0 ? h->prm.resident_idle_us : 2000;
  R->idle_ns = (unsigned long long)idle_us * 1000ull;
  R->G = h->num_sms;
  return YSM_OK;
}

// shared-memory plan of a launch: worker scratch (phase A for `pstride` points / phase B / sweep slices)
static size_t res_scratch_bytes(int pstride) {
  return (std::max(std::max(res_fv_smem(pstride), res_stamp_smem()), (size_t)YSM_RES_THREADS * 4) + 127) & ~(size_t)127;
}

static int res_launch(ysm_handle* h, size_t smem, unsigned last_seq) {
  Resident& R = *h->res;
  const GridC& g = h->g;
  CK(cudaMemsetAsync(R.d_small, 0, 2048, R.st));  // barrier counters and per-request accumulators start at zero
  CK(cudaMemsetAsync(R.d_fsum, 0, 4 * 4096, R.st));
  if (R.cellmax_cap) CK(cudaMemsetAsync(R.d_cellmax, 0, R.cellmax_cap * 8, R.st));
  memset(R.mb + R.o_exit, 0, 16);
  ResArgs A;
  memset(&A, 0, sizeof(A));
  A.db = (const uint4*)(R.mb_dev + R.o_db);
  A.req = R.mb_dev + R.o_req;
  A.pts = (const double*)(R.mb_dev + R.o_pts);
  A.spec = R.mb_dev + R.o_spec;
  A.out = (uint4*)(R.mb_dev + R.o_out);
  A.exit_line = (unsigned*)(R.mb_dev + R.o_exit);
  A.prof = R.G <= 256 ? (unsigned long long*)(R.mb_dev + R.o_prof) : nullptr;
  A.ctl = R.d_ctl;
  A.bars = (unsigned long long*)R.d_small;  // 5 counters, 128 bytes apart
  A.quit_round = (unsigned*)(R.d_small + 1024);
  A.abort_flag = (int*)(R.d_small + 1088);
  A.win = (unsigned long long*)(R.d_small + 1216);
  A.fsum = R.d_fsum;
  A.spec_dev = R.d_spec;
  A.dp_coarse = R.d_dp;
  A.dp_fine = R.d_dp + R.dp_fine_off;
  A.winrec = R.d_winrec;
  A.cells = R.d_cells;
  A.cells_cap = YSM_RES_CELLS_CAP;
  A.qpts = R.d_qpts;
  A.cache = R.d_cache;
  A.resp = R.d_resp;
  A.cellmax = R.d_cellmax;
  A.stamp_tab = h->d_stamp_tab;
  A.grid = h->d_grids;  // slot 0
  A.last_seq = last_seq;
  A.idle_ns = R.idle_ns;
  A.stall_ns = R.idle_ns + 4000000000ull;
  A.o_off = (unsigned)R.scratch;
  GridC gg = g;
  PenaltyC pp = h->pen;
  void* kargs[] = {&gg, &pp, &A};
  CK(cudaLaunchCooperativeKernel((const void*)k_match_resident, dim3(R.G), dim3(YSM_RES_THREADS), kargs, smem, R.st));
  if (!R.alive) g_res_alive.fetch_add(1);
  R.alive = true;
  R.smem = smem;
  R.launches++;
  h->launches++;
  g_res_owner[h->device] = h;
  return YSM_OK;
}

static inline const volatile uint32_t* res_chunk(const Resident& R, int i) {
  return reinterpret_cast<const volatile uint32_t*>(R.mb + R.o_out + 16 * (size_t)i);
}

// Waits for result chunk 0 of request `seq`; relaunches the kernel if it left the device (idle exit racing
// the doorbell). Returns YSM_OK or an error.
static int res_await(ysm_handle* h, unsigned seq, size_t smem) {
  Resident& R = *h->res;
  const volatile uint32_t* c0 = res_chunk(R, 0);
  const volatile uint32_t* ex = reinterpret_cast<const volatile uint32_t*>(R.mb + R.o_exit);
  const auto t_start = std::chrono::steady_clock::now();
  for (unsigned spins = 1;; spins++) {
    if (c0[2] == seq && c0[3] == 0u) return YSM_OK;
    if ((spins & 0x3F) == 0) {
      if (ex[2] == 0x45584954u) {  // the kernel has written its exit line
        const unsigned served = ex[0], code = ex[1];
        cudaError_t e = cudaStreamSynchronize(R.st);
        {
          std::lock_guard<std::mutex> lk(g_res_mu);
          if (R.alive) { R.alive = false; g_res_alive.fetch_sub(1); }
          if (g_res_owner[h->device] == h) g_res_owner[h->device] = nullptr;
        }
        if (e != cudaSuccess) return fail(h, YSM_ECUDA, std::string("resident kernel: ") + cudaGetErrorString(e));
        if (code != 0) return fail(h, YSM_ECUDA, "resident kernel aborted (barrier stall)");
        if (c0[2] == seq && c0[3] == 0u) return YSM_OK;
        if (served != seq) {
          std::lock_guard<std::mutex> lk(g_res_mu);
          if (g_res_owner[h->device] && g_res_owner[h->device] != h) res_stop_locked(g_res_owner[h->device]);
          const int rc = res_launch(h, smem, seq - 1);
          if (rc != YSM_OK) return rc;
        }
      }
      if ((spins & 0xFFFF) == 0) {
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() > 20.0)
          return fail(h, YSM_ECUDA, "resident kernel did not answer within 20 s");
        const cudaError_t q = cudaStreamQuery(R.st);
        if (q != cudaErrorNotReady && q != cudaSuccess)
          return fail(h, YSM_ECUDA, std::string("resident kernel: ") + cudaGetErrorString(q));
      }
    }
    _mm_pause();
  }
}

// One MatchScan through the resident kernel. Returns YSM_OK (result written), 1 when the request is not
// eligible or the kernel asked for the general path, or an error code.
static int res_match(ysm_handle* h, const ysm_batch* b, ysm_result* out, PhaseTrace& tr) {
  if (!h->res_enabled || h->static_grid || h->ordered_stamps || h->debug != 0 || b->n_matches != 1 || b->pool_on_device)
    return 1;
  const GridC& g = h->g;
  const int q = b->query_scan[0];
  if (q < 0 || q >= b->n_scans) return 1;  // (the general path reports it)
  const int P = b->scan_count[q];
  const int nbase = b->base_ptr[1] - b->base_ptr[0];
  if (P < 1 || P > YSM_RES_PMAX || nbase < 1 || nbase > YSM_RES_MAXBASE) return 1;
  int pmax = P;
  long long cells = 0;
  for (int k = b->base_ptr[0]; k < b->base_ptr[1]; k++) {
    const int s = b->base_idx[k];
    if (s < 0 || s >= b->n_scans) return 1;
    const int c = b->scan_count[s];
    if (c < 0 || (int64_t)b->scan_start[s] + c > b->n_points || b->scan_start[s] < 0) return 1;
    pmax = std::max(pmax, c);
    cells += c;
  }
  if ((int64_t)b->scan_start[q] + P > b->n_points || b->scan_start[q] < 0) return 1;
  const int pstride = (pmax + 7) & ~7;
  if (pmax > YSM_RES_PMAX || (long long)(nbase + 1) * pstride > YSM_RES_PTS_CAP || (long long)nbase * pstride > YSM_RES_CELLS_CAP) return 1;

  // ---- the two passes: coarse now, fine resolved on the device at the coarse winner ----
  const double pose[3] = {b->query_pose[0], b->query_pose[1], b->query_pose[2]};
  const double csx = 0.5 * (h->side - 1) * h->res_eff, crx = 2 * h->res_eff;
  const double a_off = h->prm.coarse_search_angle_offset, a_res = h->prm.coarse_angle_resolution;
  const double fo = 0.5 * h->prm.coarse_angle_resolution, fr = h->prm.fine_search_angle_resolution;
  const int nA = n_steps(a_off, a_res), nAf = n_steps(fo, fr);
  const int nX = n_steps(csx, crx), nY = nX;
  const int fnX = n_steps(crx * 0.5, h->res_eff), fnY = fnX;
  if (nA < 1 || nA > YSM_RES_MAXNA || nAf < 1 || nA * nAf > 4096 || nA > 255 || nAf > 64 || fnX * fnY > 32) return 1;
  if ((long long)nX * nY * nA > (1 << 21) || fnX * fnY * nAf > 4096) return 1;
  if ((long long)nAf * ((P + 31) / 32) > (long long)(YSM_RES_THREADS / 32) * (h->num_sms - 1)) return 1;  // fine-pass items: one per warp
  const int Ppad = align_up(P, 4);
  const size_t tab_bytes = stamp_table_bytes(g.K, g.Wt);
  // shared memory: stamp table | scratch | workers: offsets + lattice / CTA 0: query points, spec, fine offsets, fine sums
  const size_t scratch = res_scratch_bytes(pstride);
  const size_t worker = 16 * (size_t)YSM_RES_THREADS + (size_t)(((P + 7) & ~7) + nX + nY) * 4;  // point stash | offsets + lattice
  const size_t o_q = 0, o_spec = 16 * (size_t)YSM_RES_THREADS;  // (CTA 0 has the point stash too)
  const size_t o_foff = 0;                                       // (unused: the workers rotate the points)
  const size_t o_fsum = a16(o_spec + 8 * (size_t)(nA + 5 * nA * nAf));
  const size_t o_cm = a16(o_fsum + 12 * (size_t)(fnX * fnY * nAf + 2));
  const size_t tail = a16(o_cm + 8 * (size_t)std::min(nX * nY, YSM_RES_CM_SMEM));
  size_t need = tab_bytes + scratch + std::max(worker, tail);
  need = (need + 16383) & ~(size_t)16383;
  if (need > h->res_smem_limit) return 1;
  {
    const int rc = res_alloc(h);
    if (rc != YSM_OK) return rc;
  }
  Resident& R = *h->res;
  CK(cudaSetDevice(h->device));
  const size_t resp_need = (size_t)nX * nY * nA, cm_need = (size_t)nX * nY;
  const bool regrow = resp_need > R.resp_cap || cm_need > R.cellmax_cap;
  {
    // one resident kernel per device; restart ours when this request does not fit the running instance
    std::lock_guard<std::mutex> lk(g_res_mu);
    ysm_handle* owner = g_res_owner[h->device];
    if (owner && owner != h) res_stop_locked(owner);
    if (R.alive && (need > R.smem || scratch > R.scratch || regrow)) res_stop_locked(h);
  }
  if (regrow) {
    if (R.d_resp) cudaFree(R.d_resp);
    if (R.d_cellmax) cudaFree(R.d_cellmax);
    R.d_resp = nullptr; R.d_cellmax = nullptr;
    R.resp_cap = std::max(resp_need, (size_t)65536);
    R.cellmax_cap = std::max(cm_need, (size_t)4096);
    CK(cudaMalloc((void**)&R.d_resp, R.resp_cap * 8));
    CK(cudaMalloc((void**)&R.d_cellmax, R.cellmax_cap * 8));
  }

  ResReq* rq = reinterpret_cast<ResReq*>(R.mb + R.o_req);
  const unsigned seq = ++R.seq;
  const double gox = pose[0] - (0.5 * (g.roi - 1) * h->res_eff), goy = pose[1] - (0.5 * (g.roi - 1) * h->res_eff);
  rq->seq = seq; rq->cmd = RES_CMD_MATCH;
  rq->nbase = nbase; rq->Pq = P; rq->pstride = pstride; rq->do_refine = b->do_refine ? 1 : 0;
  rq->nA = nA; rq->nAf = nAf; rq->trace = tr.on ? 1 : 0;
  {
    // sweep shape: CTAs per angle x tasks per CTA x point slices = the machine (32 warps per CTA)
    const int nxc = (nX + 31) / 32, tasks = nY * nxc;
    const int cpa = std::max(1, (R.G - 1) / nA);
    const int nwarps = YSM_RES_THREADS / 32;
    int tpc = std::min(nwarps, (tasks + cpa - 1) / cpa);
    int psplit = 1;
    while (psplit * 2 * tpc <= nwarps && P / (psplit * 2) >= 16) psplit *= 2;
    rq->tpc = tpc; rq->psplit = psplit; rq->task_chunks = (tasks + tpc - 1) / tpc;
  }
  {
    // tile lists: coarse-resolution grids have few tiles per CTA but many stamps per tile
    const int tiles_per_grid = h->tnx * h->tnx;
    rq->maxt = tiles_per_grid / R.G >= 16 ? YSM_RES_MAXT : YSM_RES_MAXT / 2;
    rq->cand = YSM_RES_MAXT * YSM_RES_CAND / rq->maxt;
    rq->tail_warps = 4;
    rq->pad1 = 0;
  }
  rq->o_q = (unsigned)o_q; rq->o_spec = (unsigned)o_spec; rq->o_foff = (unsigned)o_foff; rq->o_fsum = (unsigned)o_fsum;
  rq->o_cm = (unsigned)o_cm;
  MatchDev& m = rq->m;
  m.slot = 0; m.base_begin = 0; m.base_end = nbase; m.cells_off = 0; m.gbox_off = 0; m.pad0 = 0;
  m.vpx = pose[0]; m.vpy = pose[1]; m.gox = gox; m.goy = goy;
  TableDev& ct = rq->ctab;
  ct.q_start = 0; ct.P = P; ct.Ppad = Ppad; ct.nA = nA; ct.trig_off = 0; ct.out_off = 0;
  ct.px = pose[0]; ct.py = pose[1];
  fill_inverse_rotation(ct, pose);
  ct.gox = gox; ct.goy = goy;
  rq->ftab = ct;
  rq->ftab.nA = nAf;
  PassDev& cp = rq->coarse;
  memset(&cp, 0, sizeof(cp));
  cp.slot = 0; cp.table = 0; cp.nA = nA; cp.nX = nX; cp.nY = nY; cp.P = P; cp.Ppad = Ppad; cp.fine = 0;
  cp.penalize = b->do_penalize ? 1 : 0; cp.spec = -1; cp.cmax_off = 0;
  cp.cx = pose[0]; cp.cy = pose[1]; cp.ch = pose[2];
  cp.offx = csx; cp.offy = csx; cp.resx = crx; cp.resy = crx;
  cp.angle_offset = a_off; cp.angle_res = a_res; cp.gox = gox; cp.goy = goy;
  PassDev& fp = rq->fine;
  memset(&fp, 0, sizeof(fp));
  fp.slot = 0; fp.table = 1; fp.nA = nAf; fp.nX = fnX; fp.nY = fnY; fp.P = P; fp.Ppad = Ppad; fp.fine = 1;
  fp.penalize = b->do_penalize ? 1 : 0; fp.spec = -1; fp.cmax_off = -1;
  fp.offx = crx * 0.5; fp.offy = crx * 0.5; fp.resx = h->res_eff; fp.resy = h->res_eff;
  fp.angle_offset = fo; fp.angle_res = fr; fp.gox = gox; fp.goy = goy;
  {
    const double start_angle = pose[2] - a_off;
    for (int a = 0; a < nA; a++) {
      const double angle = start_angle + (double)(uint32_t)a * a_res;
      const double ca = cos(angle), sa = sin(angle);
      const double hn = h_normalize_angle(angle);
      const bool same = dbits(hn) == dbits(angle);
      rq->trig4[a][0] = ca; rq->trig4[a][1] = sa;
      rq->trig4[a][2] = same ? ca : cos(hn);
      rq->trig4[a][3] = same ? sa : sin(hn);
      rq->ap[a] = b->do_penalize ? h_penalty_angle(pose[2], a_off, a_res, a, h->pen) : 1.0;
    }
  }
  // points: base scan s at mailbox slot s, the query at slot nbase -- unless the device-resident scan store
  // already holds them (ysm_batch::scan_tag): then nothing is copied and the kernel reads them from HBM.
  // An untagged or new scan goes through the mailbox; a new TAGGED scan is also kept by the kernel in a free /
  // least recently used slot, valid once this request has completed.
  double* pts = reinterpret_cast<double*>(R.mb + R.o_pts);
  ResScanInfo info[YSM_RES_MAXBASE + 1];
  int pending[YSM_RES_MAXBASE + 1];
  int npending = 0;
  const bool use_cache = b->scan_tag != nullptr && nbase + 1 <= YSM_RES_POLLERS;
  R.use_clock++;
  for (int k = 0; k <= nbase; k++) {
    const int s = k < nbase ? b->base_idx[b->base_ptr[0] + k] : q;
    const int c = b->scan_count[s];
    rq->counts[k] = (unsigned short)c;
    info[k].count = c; info[k].src_slot = 0; info[k].store_slot = 0;
    const uint64_t tag = use_cache ? b->scan_tag[s] : 0;
    if (tag != 0 && c > 0) {
      int hit = -1, victim = -1;
      for (int j = 0; j < YSM_RES_CACHE_SLOTS; j++) {
        const Resident::Slot& sl = R.slots[j];
        if (sl.valid && sl.tag == tag && sl.count == c) { hit = j; break; }
      }
      if (hit >= 0) {
        R.slots[hit].last_use = R.use_clock;
        info[k].src_slot = hit + 1;
        h->work[14]++;
        continue;  // resident: no copy
      }
      bool dup = false;  // the same new scan twice in one request: the first occurrence fills the slot
      for (int j = 0; j < npending; j++) dup = dup || R.slots[pending[j]].tag == tag;
      if (!dup) {
        for (int j = 0; j < YSM_RES_CACHE_SLOTS; j++) {
          const Resident::Slot& sl = R.slots[j];
          if (sl.last_use == R.use_clock) continue;  // in use by this request
          if (victim < 0 || !sl.valid || (R.slots[victim].valid && sl.last_use < R.slots[victim].last_use)) {
            victim = j;
            if (!sl.valid) break;
          }
        }
        if (victim >= 0) {
          Resident::Slot& sl = R.slots[victim];
          sl.tag = tag; sl.count = c; sl.valid = false; sl.last_use = R.use_clock;
          pending[npending++] = victim;
          info[k].store_slot = victim + 1;
        }
      }
    }
    if (c) memcpy(pts + 2 * (size_t)k * pstride, b->pool_xy + 2 * (size_t)b->scan_start[s], 16 * (size_t)c);
  }
  const unsigned ctl_bytes = (unsigned)(offsetof(ResReq, trig4) + 32 * (size_t)nA);
  tr.mark("resident: request staged");
  {
    const double mf[4] = {m.vpx, m.vpy, m.gox, m.goy};
    res_ring(R, seq, (unsigned)RES_CMD_MATCH | ((unsigned)nbase << 8) | ((unsigned)nA << 16) | ((unsigned)nAf << 24),
             (unsigned)P | ((unsigned)pstride << 16), ctl_bytes, info, nbase + 1, mf);
  }
  if (!R.alive) {
    std::lock_guard<std::mutex> lk(g_res_mu);
    R.scratch = scratch;
    const int rc = res_launch(h, need, seq - 1);
    if (rc != YSM_OK) return rc;
    tr.mark("resident: kernel launch");
  }
  // ---- speculative fine tables while the GPU builds and sweeps (host libm, off the critical path) ----
  double* spec = reinterpret_cast<double*>(R.mb + R.o_spec + sizeof(ResSpecHdr));
  if (b->do_refine) {
    double* ft = spec + nA;
    for (int a = 0; a < nA; a++) {
      const double heading = atan2(rq->trig4[a][3] / 1.0, rq->trig4[a][2] / 1.0);
      spec[a] = heading;
      const double start_angle = heading - fo;
      for (int f = 0; f < nAf; f++) {
        const double angle = start_angle + (double)(uint32_t)f * fr;
        const double ca = cos(angle), sa = sin(angle);
        const double hn = h_normalize_angle(angle);
        const bool same = dbits(hn) == dbits(angle);
        double* e = ft + 4 * ((size_t)a * nAf + f);
        e[0] = ca; e[1] = sa;
        e[2] = same ? ca : cos(hn);
        e[3] = same ? sa : sin(hn);
        // the fine pass's angle penalty, should this coarse angle win (its search centre heading is `heading`)
        ft[4 * (size_t)nA * nAf + (size_t)a * nAf + f] = b->do_penalize ? h_penalty_angle(heading, fo, fr, f, h->pen) : 1.0;
      }
    }
    __atomic_store_n(reinterpret_cast<uint32_t*>(R.mb + R.o_spec), seq, __ATOMIC_RELEASE);
    tr.mark("resident: spec tables");
  }
  {
    const int rc = res_await(h, seq, need);
    if (rc != YSM_OK) {
      for (int j = 0; j < YSM_RES_CACHE_SLOTS; j++) R.slots[j].valid = false;  // (the store may be half written)
      return rc;
    }
    // phase A has run to completion for every answered request: the slots it filled are valid now
    for (int j = 0; j < npending; j++) R.slots[pending[j]].valid = true;
  }
  // ---- read the chunks -------------------------------------------------------------------------------
  const volatile uint32_t* c0 = res_chunk(R, 0);
  const unsigned status = c0[0] & 0xFFu, has_fine = (c0[0] >> 8) & 0xFFu, total = c0[1];
  if (total < 1 || total > YSM_RES_CHUNKS) return fail(h, YSM_ECUDA, "resident kernel: malformed result");
  uint64_t payload[YSM_RES_CHUNKS];
  for (unsigned i = 1; i < total; i++) {
    const volatile uint32_t* c = res_chunk(R, (int)i);
    unsigned spins = 0;
    while (!(c[2] == seq && c[3] == i)) {
      if (++spins > 400000000u) return fail(h, YSM_ECUDA, "resident kernel: incomplete result");
      _mm_pause();
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    payload[i] = (uint64_t)c[0] | ((uint64_t)c[1] << 32);
  }
  tr.mark("resident: results");
  R.served++;
  h->work[13]++;
  const int np = (int)(sizeof(PassOut) / 8);
  if (tr.on && total >= (unsigned)(1 + 2 * np + YSM_RES_TS)) {
    const uint64_t* ts = payload + total - YSM_RES_TS;
    static const char* names[] = {"detect -> phase A done", "barrier 1", "ctl + offsets + stamp", "barrier 2",
                                  "spec tables in smem", "sweep done (wait)", "coarse reduce", "fine pass"};
    for (int k = 0; k < 8; k++)
      fprintf(stderr, "[ysm-resident] %-26s %8.2f us\n", names[k], (double)(ts[k + 1] - ts[k]) * 1e-3);
    fprintf(stderr, "[ysm-resident]   SM clock over the request: %.0f MHz\n", (double)(ts[21] - ts[20]) / ((double)(ts[8] - ts[0]) * 1e-3));
    if (has_fine)
      fprintf(stderr, "[ysm-resident]   fine: winner published +%.2f, workers' sums arrived +%.2f, finish %.2f us\n",
              (double)(ts[9] - ts[7]) * 1e-3, (double)(ts[10] - ts[9]) * 1e-3, (double)(ts[8] - ts[10]) * 1e-3);
    // per-CTA phases (written after barrier 3: give the stores a moment)
    std::this_thread::sleep_for(std::chrono::microseconds(200));
    const uint64_t* pf = reinterpret_cast<const uint64_t*>(R.mb + R.o_prof);
    static const char* pn[] = {"ctl in smem -> offsets", "collect", "stamp", "barrier 2 (wait)", "sweep", "tail + barrier 3 + clear"};
    for (int k = 0; k < 6; k++) {
      std::vector<double> v;
      for (int c = 1; c < R.G; c++) v.push_back((double)(int64_t)(pf[(size_t)c * YSM_RES_PROF + k + 1] - pf[(size_t)c * YSM_RES_PROF + k]) * 1e-3);
      std::sort(v.begin(), v.end());
      fprintf(stderr, "[ysm-resident]   workers %-26s min %6.2f  med %6.2f  max %6.2f us\n", pn[k], v.front(), v[v.size() / 2], v.back());
    }
    {
      // inside the sweep (pf[4] barrier-2 exit, [13] penalties done, [14] row sums done, [15] responses stored, [5] end)
      // and inside the stamp (pf[2] start, [16] slots known, [17] scattered + staged, [18] barrier, [3] written)
      static const int seq[3][5] = {{4, 13, 14, 15, 5}, {2, 16, 17, 18, 3}, {4, 19, 20, 13, 14}};
      static const char* nm[3][4] = {{"sweep: penalties", "sweep: row sums", "sweep: combine + response", "sweep: block max"},
                                     {"stamp: slot list", "stamp: scatter + stage", "stamp: barrier", "stamp: combine + write"},
                                     {"sweep: barrier exit -> entry", "sweep: setup", "sweep: dp load issue", "sweep: row sums"}};
      for (int q = 0; q < 3; q++)
        for (int k = 0; k < 4; k++) {
          std::vector<double> v;
          for (int c = 1; c < R.G; c++)
            v.push_back((double)(int64_t)(pf[(size_t)c * YSM_RES_PROF + seq[q][k + 1]] - pf[(size_t)c * YSM_RES_PROF + seq[q][k]]) * 1e-3);
          std::sort(v.begin(), v.end());
          fprintf(stderr, "[ysm-resident]     %-28s min %6.2f  med %6.2f  max %6.2f us\n", nm[q][k], v.front(), v[v.size() / 2], v.back());
        }
    }
    {
      const uint64_t* f1 = pf + (size_t)1 * YSM_RES_PROF;  // CTA 1: base scan 0
      fprintf(stderr, "[ysm-resident]   CTA 1 phase A: pull %.2f, next[] %.2f, chain %.2f, cells %.2f us; detect skew vs CTA 0 %+.2f us\n",
              (double)(int64_t)(f1[9] - f1[8]) * 1e-3, (double)(int64_t)(f1[10] - f1[9]) * 1e-3, (double)(int64_t)(f1[11] - f1[10]) * 1e-3,
              (double)(int64_t)(f1[12] - f1[11]) * 1e-3, (double)(int64_t)(f1[8] - pf[8]) * 1e-3);
    }
    {
      // the CTA every other one waits for at barrier 2
      int slow = 1;
      for (int c = 1; c < R.G; c++)
        if ((int64_t)(pf[(size_t)c * YSM_RES_PROF + 3] - pf[(size_t)c * YSM_RES_PROF + 2]) >
            (int64_t)(pf[(size_t)slow * YSM_RES_PROF + 3] - pf[(size_t)slow * YSM_RES_PROF + 2])) slow = c;
      const uint64_t* q = pf + (size_t)slow * YSM_RES_PROF;
      fprintf(stderr, "[ysm-resident]   slowest stamp: CTA %d, %d tiles: collect %.2f, slot list %.2f, scatter %.2f, barrier %.2f, write %.2f us\n",
              slow, (int)q[7], (double)(int64_t)(q[2] - q[1]) * 1e-3, (double)(int64_t)(q[16] - q[2]) * 1e-3,
              (double)(int64_t)(q[17] - q[16]) * 1e-3, (double)(int64_t)(q[18] - q[17]) * 1e-3, (double)(int64_t)(q[3] - q[18]) * 1e-3);
    }
    int tmax = 0, tsum = 0;
    for (int c = 0; c < R.G; c++) { const int t = (int)pf[(size_t)c * YSM_RES_PROF + 7]; tmax = std::max(tmax, t); tsum += t; }
    uint64_t b1min = ~0ull, b1max = 0;
    for (int c = 0; c < R.G; c++) { b1min = std::min(b1min, pf[(size_t)c * YSM_RES_PROF]); b1max = std::max(b1max, pf[(size_t)c * YSM_RES_PROF]); }
    fprintf(stderr, "[ysm-resident]   tiles stamped %d (max %d per CTA); barrier-1 exit skew %.2f us\n", tsum, tmax, (double)(b1max - b1min) * 1e-3);
  }
  if (status != RES_ST_OK) return 1;
  PassOut po, fo_out;
  memcpy(&po, payload + 1, sizeof(PassOut));
  memcpy(&fo_out, payload + 1 + np, sizeof(PassOut));
  // ---- finish exactly as CorrelateScan / MatchScan do (host libm) -----------------------------------
  if (po.n_ties <= 0) return 1;  // ("Unable to find best position": let the general path report it)
  PassHost ph;
  ph.match = 0; ph.fine = false; ph.spec = false;
  ph.cx = pose[0]; ph.cy = pose[1]; ph.ch = pose[2];
  ph.offx = csx; ph.offy = csx; ph.resx = crx; ph.resy = crx;
  ph.angle_offset = a_off; ph.angle_res = a_res; ph.nA = nA; ph.nX = nX; ph.nY = nY; ph.ang_off = 0;
  double cov[9];
  const double heading = atan2(po.ty, po.tx);
  finalize_positional(h, ph, po, cov);
  double mean[3] = {po.avg_x, po.avg_y, heading};
  double best = po.best > 1.0 ? 1.0 : po.best;
  int n_passes = 1, n_ties = po.n_ties;
  if (h->prm.use_response_expansion && h_double_equal(best, 0.0)) return 1;  // response expansion: general path
  if (b->do_refine) {
    const int a = po.first_idx % nA;
    if (!has_fine || po.n_ties != 1 || fo_out.n_ties <= 0 || dbits(spec[a]) != dbits(heading)) return 1;
    PassHost fph = ph;
    fph.fine = true;
    fph.cx = po.avg_x; fph.cy = po.avg_y; fph.ch = spec[a];
    fph.offx = crx * 0.5; fph.offy = crx * 0.5; fph.resx = h->res_eff; fph.resy = h->res_eff;
    fph.angle_offset = fo; fph.angle_res = fr; fph.nA = nAf; fph.nX = fnX; fph.nY = fnY;
    int angs[256];
    const uint64_t* ap = payload + 1 + 2 * np;
    for (int k = 0; k < nAf; k++) angs[k] = (int)(uint32_t)(ap[k >> 1] >> (32 * (k & 1)));
    const double fheading = atan2(fo_out.ty, fo_out.tx);
    finalize_angular(fph, fo_out, fheading, angs, P, cov);
    mean[0] = fo_out.avg_x; mean[1] = fo_out.avg_y; mean[2] = fheading;
    best = fo_out.best > 1.0 ? 1.0 : fo_out.best;
    n_passes = 2;
    n_ties = fo_out.n_ties;
    h->work[10]++;
  }
  ysm_result& r = out[0];
  r.response = best;
  r.x = mean[0]; r.y = mean[1]; r.heading = mean[2];
  memcpy(r.cov, cov, sizeof(cov));
  r.n_passes = n_passes; r.n_ties = n_ties; r.status = YSM_OK; r._pad = 0; r._reserved = 0.0;
  tr.mark("resident: host finalize");
  return YSM_OK;
}

extern "C" int ysm_debug_ping(ysm_handle* h, int32_t n, double* rtt_us) {
  if (!h || n < 0 || (n > 0 && !rtt_us)) return YSM_EINVAL;
  if (!h->res_enabled || h->static_grid) return fail(h, YSM_EUNSUP, "no resident kernel on this handle");
  CK(cudaSetDevice(h->device));
  {
    const int rc = res_alloc(h);
    if (rc != YSM_OK) return rc;
  }
  Resident& R = *h->res;
  const size_t need = (stamp_table_bytes(h->g.K, h->g.Wt) + res_scratch_bytes(1024) + 16 * (size_t)YSM_RES_THREADS + 8192 + 16383) & ~(size_t)16383;
  if (need > h->res_smem_limit) return fail(h, YSM_EUNSUP, "resident kernel does not fit this configuration");
  if (!R.d_resp) {
    R.resp_cap = 65536; R.cellmax_cap = 4096;
    CK(cudaMalloc((void**)&R.d_resp, R.resp_cap * 8));
    CK(cudaMalloc((void**)&R.d_cellmax, R.cellmax_cap * 8));
  }
  for (int i = 0; i < n; i++) {
    const auto t0 = std::chrono::steady_clock::now();
    const unsigned seq = ++R.seq;
    res_ring(R, seq, RES_CMD_PING, 0u, 0u);
    if (!R.alive) {
      std::lock_guard<std::mutex> lk(g_res_mu);
      if (g_res_owner[h->device] && g_res_owner[h->device] != h) res_stop_locked(g_res_owner[h->device]);
      R.scratch = res_scratch_bytes(1024);
      const int rc = res_launch(h, need, seq - 1);
      if (rc != YSM_OK) return rc;
    }
    const int rc = res_await(h, seq, R.smem);
    if (rc != YSM_OK) return rc;
    rtt_us[i] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
  }
  return YSM_OK;
}

// MatchScan for a query scan that has range readings but not one point reading: the
// lookup table is empty, GetResponse returns 0 for every pose ("if (nPoints == 0) return response"), so in
// every CorrelateScan pass all poses tie with the best response 0 and the result is the ordered average of the
// whole search lattice; response expansion (if enabled) widens the angle window three times, a fine pass
// follows when asked. There is not a single grid lookup in this schedule -- nothing for the GPU to do -- so the
// host runtime walks the lattices in Karto's storage order (y, x, angle) with the same additions.
static void zero_point_schedule(const ysm_handle* h, const double* pose, bool do_refine, MatchState& s) {
  auto tie_average = [](const double* center, double offx, double offy, double resx, double resy, double angle_offset,
                        double angle_res, double* mean) {
    const int nX = n_steps(offx, resx), nY = n_steps(offy, resy), nA = n_steps(angle_offset, angle_res);
    const double startX = -offx, startY = -offy, start_angle = center[2] - angle_offset;
    std::vector<double> hc((size_t)nA), hs((size_t)nA);
    for (int a = 0; a < nA; a++) {
      const double hn = h_normalize_angle(start_angle + (double)(uint32_t)a * angle_res);
      hc[a] = cos(hn);
      hs[a] = sin(hn);
    }
    double sx = 0.0, sy = 0.0, tx = 0.0, ty = 0.0;
    for (int iy = 0; iy < nY; iy++) {
      const double ny = center[1] + (startY + (double)(uint32_t)iy * resy);
      for (int ix = 0; ix < nX; ix++) {
        const double nx = center[0] + (startX + (double)(uint32_t)ix * resx);
        for (int a = 0; a < nA; a++) {
          sx += nx;
          sy += ny;
          tx += hc[a];
          ty += hs[a];
        }
      }
    }
    const double cnt = (double)((long long)nX * nY * nA);
    mean[0] = sx / cnt;
    mean[1] = sy / cnt;
    mean[2] = atan2(ty / cnt, tx / cnt);
    return nX * nY * nA;
  };
  const double csx = 0.5 * (h->side - 1) * h->res_eff, crx = 2 * h->res_eff;
  double angle_offset = h->prm.coarse_search_angle_offset;
  double mean[3] = {pose[0], pose[1], pose[2]};
  int passes = 0, ties = 0;
  for (int stage = 0; stage < 4; stage++) {
    ties = tie_average(pose, csx, csx, crx, crx, angle_offset, h->prm.coarse_angle_resolution, mean);
    passes++;
    if (!h->prm.use_response_expansion || stage == 3) break;
    angle_offset += 20 * KT_PI_180;  // (best == 0: the next response expansion)
  }
  memset(s.cov, 0, sizeof(s.cov));
  s.cov[0] = MAX_VARIANCE;  // ComputePositionalCovariance with best < KT_TOLERANCE
  s.cov[4] = MAX_VARIANCE;
  s.cov[8] = 4 * h_square(h->prm.coarse_angle_resolution);
  if (do_refine) {
    double fmean[3];
    ties = tie_average(mean, crx * 0.5, crx * 0.5, h->res_eff, h->res_eff, 0.5 * h->prm.coarse_angle_resolution,
                       h->prm.fine_search_angle_resolution, fmean);
    passes++;
    mean[0] = fmean[0]; mean[1] = fmean[1]; mean[2] = fmean[2];
    s.cov[8] = 1000 * h_square(h->prm.fine_search_angle_resolution);  // ComputeAngularCovariance with norm == 0
  }
  s.mean[0] = mean[0]; s.mean[1] = mean[1]; s.mean[2] = mean[2];
  s.best = 0.0;
  s.n_passes = passes;
  s.n_ties = ties;
  s.status = YSM_OK;
  s.stage = 5;
}

// --------------------------------------------------------------------------------------------
static int match_batch_impl(ysm_handle* h, const ysm_batch* b, ysm_result* out, cudaStream_t st) {
  if (!h || !b || !out) return YSM_EINVAL;
  if (b->n_matches < 0 || b->n_scans < 0) return fail(h, YSM_EINVAL, "negative sizes");
  PhaseTrace tr;
  KernelTrace kt;
  static const bool trace_gpu = getenv("YSM_TRACE_GPU") != nullptr;
  kt.init(trace_gpu, st);
  kt.mark("start");
  CK(cudaSetDevice(h->device));
  const GridC& g = h->g;
  const bool timing = (h->debug & YSM_DEBUG_TIME_KERNELS) && h->ev_ok;
  h->t_sweep = h->t_build = h->t_reduce = h->t_total = 0.0;
  for (int i = 0; i < 16; i++) h->work[i] = 0;
  if (b->n_matches == 0) return YSM_OK;
  if (timing) CK(cudaMemsetAsync(h->d_issued, 0, 8, st));

  // deferred clear from a previous KEEP_GRIDS batch
  if (h->grids_dirty) {
    clear_wave(h, h->cur_matches, st);
    h->grids_dirty = false;
  }

  // validate + sizes
  int pmax = 1;
  for (int s = 0; s < b->n_scans; s++) {
    if (b->scan_count[s] < 0 || b->scan_start[s] < 0 ||
        (int64_t)b->scan_start[s] + b->scan_count[s] > b->n_points)
      return fail(h, YSM_EINVAL, "scan range outside the point pool");
    pmax = std::max(pmax, b->scan_count[s]);
  }
  if (pmax > 16384) return fail(h, YSM_EUNSUP, "more than 16384 point readings in one scan");
  for (int i = 0; i < b->n_matches; i++) {
    if (b->query_scan[i] < 0 || b->query_scan[i] >= b->n_scans) return fail(h, YSM_EINVAL, "bad query scan index");
    if (b->base_ptr[i + 1] < b->base_ptr[i]) return fail(h, YSM_EINVAL, "base_ptr not monotone");
    for (int k = b->base_ptr[i]; k < b->base_ptr[i + 1]; k++)
      if (b->base_idx[k] < 0 || b->base_idx[k] >= b->n_scans) return fail(h, YSM_EINVAL, "bad base scan index");
  }

  // Latency path: a handful of matches over a small pool. Everything the GPU needs (points,
  // descriptors, pass tables) travels in ONE host->device copy, the fine pass is chained on the
  // device behind the coarse pass (k_reduce points it at the winner), and the host synchronises once.
  const bool small = b->n_matches <= 8 && (h->static_grid || b->n_matches <= h->slots) &&
                     (b->pool_on_device || (size_t)b->n_points * 16 <= (1u << 20)) && b->n_scans <= 4096;
  const bool speculate = small && b->do_refine && !(h->debug & YSM_DEBUG_NO_SPECULATE);

  // pool / scan directory -> device (throughput path: once per call, straight from the caller's memory)
  const double* d_pool = nullptr;
  const int *d_scan_start = nullptr, *d_scan_count = nullptr;
  if (b->pool_on_device) d_pool = b->pool_xy;
  if (!small) {
    if (!b->pool_on_device) {
      CK(h->d_pool.ensure(std::max<size_t>(16, (size_t)b->n_points * 16)));
      if (b->n_points > 0)
        CK(cudaMemcpyAsync(h->d_pool.p, b->pool_xy, (size_t)b->n_points * 16, cudaMemcpyHostToDevice, st));
      h->work[6] += (int64_t)b->n_points * 16;
      d_pool = (const double*)h->d_pool.p;
    }
    CK(h->d_scan_start.ensure((size_t)std::max(1, b->n_scans) * 4));
    CK(h->d_scan_count.ensure((size_t)std::max(1, b->n_scans) * 4));
    if (b->n_scans > 0) {
      CK(cudaMemcpyAsync(h->d_scan_start.p, b->scan_start, (size_t)b->n_scans * 4, cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(h->d_scan_count.p, b->scan_count, (size_t)b->n_scans * 4, cudaMemcpyHostToDevice, st));
      h->work[6] += (int64_t)b->n_scans * 8;
    }
    d_scan_start = (const int*)h->d_scan_start.p;
    d_scan_count = (const int*)h->d_scan_count.p;
  }

  if (timing) CK(cudaEventRecord(h->ev[6], st));
  tr.mark("validate + pool H2D");

  const int tiles_per_grid = h->tnx * h->tnx;
  const int tps1 = (2 * g.half_kernel + YSM_TILE - 1) / YSM_TILE + 1;  // tiles a stamp can span per axis
  const int tiles_per_stamp = tps1 * tps1;
  const double csx = 0.5 * (h->side - 1) * h->res_eff;
  const double crx = 2 * h->res_eff;
  if (!h->d_dpc) {
    // distance half of the odometry penalty for the coarse lattice: it depends on the matcher's configuration
    // only, and its f64 division per pose is the dearest thing in the sweep's epilogue
    const int nX = n_steps(csx, crx);
    std::vector<double> dp((size_t)nX * nX);
    for (int iy = 0; iy < nX; iy++)
      for (int ix = 0; ix < nX; ix++) dp[(size_t)iy * nX + ix] = h_penalty_distance(csx, crx, csx, crx, ix, iy, h->pen);
    CK(cudaMalloc((void**)&h->d_dpc, dp.size() * 8));
    CK(cudaMemcpy(h->d_dpc, dp.data(), dp.size() * 8, cudaMemcpyHostToDevice));
    h->dpc_nx = nX;
  }
  const int S = h->static_grid ? 4096 : balanced_wave(b->n_matches, h->slots);  // matches per wave
  const int nAf = n_steps(0.5 * h->prm.coarse_angle_resolution, h->prm.fine_search_angle_resolution);

  std::vector<MatchState> states;
  std::vector<MatchDev> hm;
  std::vector<int> hbase;
  std::vector<ScanRef> hscans;  // latency path: one entry per (match, base scan)
  PassPlan& pl = h->plan;

  h->last_slot_of_match.assign(b->n_matches, -1);
  h->last_coarse_table_off.assign(b->n_matches, -1);
  h->last_coarse_nA.assign(b->n_matches, 0);
  h->last_coarse_P.assign(b->n_matches, 0);
  h->last_coarse_Ppad.assign(b->n_matches, 0);

  for (int w0 = 0; w0 < b->n_matches; w0 += S) {
    const int w1 = std::min(b->n_matches, w0 + S);
    const int nw = w1 - w0;
    for (ysm_handle::SliceWait& sw : h->slice_wait)
      if (!sw.waited && sw.lo < w1 && sw.hi > w0) {  // this wave's scans are (being) uploaded behind sw.ev
        while (sw.seq_done->load(std::memory_order_acquire) <= sw.seq) std::this_thread::yield();  // until recorded
        CK(cudaStreamWaitEvent(st, sw.ev, 0));
        sw.waited = true;
      }
    h->last_wave_begin = w0;
    h->last_wave_end = w1;
    // ---- wave setup: match descriptors, base lists ------------------------------------------
    states.assign(nw, MatchState());
    hm.assign(nw, MatchDev());
    hbase.clear();
    hscans.clear();
    long long cells_total = 0, gbox_total = 0, work_cap = 0, max_match_cells = 1;
    int nbase_max = 1;
    for (int i = 0; i < nw; i++) {
      const int mi = w0 + i;
      MatchState& s = states[i];
      s.idx = mi;
      s.slot = h->static_grid ? 0 : i;
      s.q = b->query_scan[mi];
      s.P = b->scan_count[s.q];
      s.pose[0] = b->query_pose[3 * mi];
      s.pose[1] = b->query_pose[3 * mi + 1];
      s.pose[2] = b->query_pose[3 * mi + 2];
      s.gox = s.pose[0] - (0.5 * (g.roi - 1) * h->res_eff);
      s.goy = s.pose[1] - (0.5 * (g.roi - 1) * h->res_eff);
      if (h->static_grid) { s.gox = h->map_ox; s.goy = h->map_oy; }
      s.stage = 0;
      s.angle_offset_cur = h->prm.coarse_search_angle_offset;
      s.n_passes = 0;
      s.n_ties = 0;
      s.status = YSM_OK;
      s.best = 0.0;
      memset(s.cov, 0, sizeof(s.cov));
      s.mean[0] = s.pose[0]; s.mean[1] = s.pose[1]; s.mean[2] = s.pose[2];
      MatchDev& m = hm[i];
      m.slot = h->static_grid ? 0 : i;
      m.base_begin = (int)hbase.size();
      long long mc = 0;
      if (s.P > 0 && !h->static_grid) {
        for (int k = b->base_ptr[mi]; k < b->base_ptr[mi + 1]; k++) {
          hbase.push_back(b->base_idx[k]);
          if (small) hscans.push_back(ScanRef{i, k - b->base_ptr[mi], (int)mc, 0});
          mc += b->scan_count[b->base_idx[k]];
        }
      }
      m.base_end = (int)hbase.size();
      nbase_max = std::max(nbase_max, m.base_end - m.base_begin);
      m.cells_off = (int)cells_total;
      m.gbox_off = (int)gbox_total;
      m.pad0 = 0;
      cells_total += mc;
      max_match_cells = std::max(max_match_cells, mc);
      gbox_total += (mc + 31) / 32;
      work_cap += std::min<long long>((long long)tiles_per_grid, mc * tiles_per_stamp);
      m.vpx = s.pose[0]; m.vpy = s.pose[1];
      m.gox = s.gox; m.goy = s.goy;
      if (s.P == 0 && b->scan_raw_count && b->scan_raw_count[s.q] > 0) {
        // Beams, but none inside [min_range, range_threshold]: Karto does NOT return early (it tests the number
        // of RANGE readings) -- see zero_point_schedule.
        zero_point_schedule(h, s.pose, b->do_refine != 0, s);
      } else if (s.P == 0) {
        // scan has no readings (MatchScan early return)
        s.stage = 5;
        s.cov[0] = MAX_VARIANCE;
        s.cov[4] = MAX_VARIANCE;
        s.cov[8] = 4 * h_square(h->prm.coarse_angle_resolution);
        s.best = 0.0;
      }
      h->last_slot_of_match[mi] = h->static_grid ? 0 : i;
    }
    if (cells_total > 0x7fffff00LL) return fail(h, YSM_ENOMEM, "wave has too many base points");
    CK(h->d_cells.ensure(std::max<size_t>(4, (size_t)cells_total * 4)));
    CK(h->d_ptcell.ensure(std::max<size_t>(4, (size_t)cells_total * 4)));
    CK(h->d_cellcount.ensure((size_t)nw * 4));
    CK(h->d_gbox.ensure(std::max<size_t>(8, (size_t)gbox_total * 8)));
    CK(h->d_work.ensure(std::max<size_t>(8, (size_t)work_cap * 8)));
    h->work[5] += cells_total;

    // device views of the wave-static data (set by upload below)
    const MatchDev* d_matches = nullptr;
    const int* d_base_idx = nullptr;
    int* d_workcount = nullptr;

    // ---- pass planning (host libm) -----------------------------------------------------------
    // Builds the tables / passes of the next iteration for every match that is not done. With
    // `spec`, every coarse pass also gets a speculative fine pass resolved on the device.
    auto plan_passes = [&](bool spec) -> int {
      pl.clear();
      for (int i = 0; i < nw; i++) {
        MatchState& s = states[i];
        if (s.stage >= 5) continue;
        PassHost ph;
        ph.match = i;
        ph.fine = (s.stage == 4);
        if (!ph.fine) {
          ph.cx = s.pose[0]; ph.cy = s.pose[1]; ph.ch = s.pose[2];
          ph.offx = csx; ph.offy = csx; ph.resx = crx; ph.resy = crx;
          ph.angle_offset = s.angle_offset_cur;
          ph.angle_res = h->prm.coarse_angle_resolution;
        } else {
          ph.cx = s.mean[0]; ph.cy = s.mean[1]; ph.ch = s.mean[2];
          ph.offx = crx * 0.5; ph.offy = crx * 0.5; ph.resx = h->res_eff; ph.resy = h->res_eff;
          ph.angle_offset = 0.5 * h->prm.coarse_angle_resolution;
          ph.angle_res = h->prm.fine_search_angle_resolution;
        }
        ph.nX = n_steps(ph.offx, ph.resx);
        ph.nY = n_steps(ph.offy, ph.resy);
        ph.nA = n_steps(ph.angle_offset, ph.angle_res);
        ph.spec = false;
        if (ph.nA <= 0 || ph.nA > 100000) return fail(h, YSM_EINVAL, "bad angle search window");
        // lookup table (deduplicated: same query scan, pose and angle window share one table)
        TableKey key;
        key.q = s.q;
        key.b[0] = dbits(s.pose[0]); key.b[1] = dbits(s.pose[1]); key.b[2] = dbits(s.pose[2]);
        key.b[3] = dbits(ph.ch); key.b[4] = dbits(ph.angle_offset); key.b[5] = dbits(ph.angle_res);
        int tid;
        auto it = pl.tab_index.find(key);
        const int Ppad = align_up(s.P, 4);
        if (it == pl.tab_index.end()) {
          TableDev t;
          t.q_start = b->scan_start[s.q];
          t.P = s.P; t.Ppad = Ppad; t.nA = ph.nA;
          t.trig_off = (int)(pl.trig.size() / 2);
          t.out_off = (int)pl.off_elems;
          pl.off_elems += (size_t)ph.nA * Ppad;
          t.px = s.pose[0]; t.py = s.pose[1];
          fill_inverse_rotation(t, s.pose);
          t.gox = s.gox; t.goy = s.goy;
          const double start_angle = ph.ch - ph.angle_offset;
          for (int a = 0; a < ph.nA; a++) {
            const double angle = start_angle + (double)(uint32_t)a * ph.angle_res;
            pl.trig.push_back(cos(angle));
            pl.trig.push_back(sin(angle));
          }
          tid = (int)pl.tab.size();
          pl.tab.push_back(t);
          pl.tab_index.emplace(key, tid);
        } else {
          tid = it->second;
        }
        PassDev pd;
        memset(&pd, 0, sizeof(pd));
        pd.slot = s.slot; pd.table = tid; pd.nA = ph.nA; pd.nX = ph.nX; pd.nY = ph.nY;
        pd.P = s.P; pd.Ppad = Ppad; pd.fine = ph.fine ? 1 : 0; pd.penalize = b->do_penalize ? 1 : 0;
        pd.sums_off = (int)pl.sums_elems;
        pl.sums_elems += (size_t)ph.nX * ph.nY * ph.nA;
        // cos/sin of the normalised headings (tie average); the table's own entries when
        // normalisation leaves every angle unchanged
        {
          const double start_angle = ph.ch - ph.angle_offset;
          bool same = true;
          for (int a = 0; a < ph.nA && same; a++) {
            const double angle = start_angle + (double)(uint32_t)a * ph.angle_res;
            same = dbits(h_normalize_angle(angle)) == dbits(angle);
          }
          if (same) {
            pd.htrig_off = pl.tab[tid].trig_off;
          } else {
            pd.htrig_off = (int)(pl.trig.size() / 2);
            for (int a = 0; a < ph.nA; a++) {
              const double hn = h_normalize_angle(start_angle + (double)(uint32_t)a * ph.angle_res);
              pl.trig.push_back(cos(hn));
              pl.trig.push_back(sin(hn));
            }
          }
        }
        pd.ang_off = pl.ang_elems;
        ph.ang_off = pl.ang_elems;
        if (ph.fine) pl.ang_elems += ph.nA;
        pd.spec = -1;
        pd.cmax_off = -1;
        if (!ph.fine) {
          pd.cmax_off = (int)pl.cmax_elems;
          pl.cmax_elems += (size_t)ph.nX * ph.nY;
        }
        pd.cx = ph.cx; pd.cy = ph.cy; pd.ch = ph.ch;
        pd.offx = ph.offx; pd.offy = ph.offy; pd.resx = ph.resx; pd.resy = ph.resy;
        pd.angle_offset = ph.angle_offset; pd.angle_res = ph.angle_res;
        pd.gox = s.gox; pd.goy = s.goy;
        const int pid = (int)pl.pass.size();
        s.pass_id = pid;
        pl.pass.push_back(pd);
        pl.ph.push_back(ph);
        pl.spec_of.push_back(-1);
        pl.spec_h.push_back(-1);
        if (!ph.fine) {
          for (int a = 0; a < ph.nA; a++) pl.pa.push_back(PassAngle{pid, a, tid, pl.tab[tid].trig_off + a});
          h->work[0] += (int64_t)ph.nX * ph.nY * ph.nA * s.P;
          pl.max_lat_P = std::max(pl.max_lat_P, s.P);
          pl.max_lat_nx = std::max(pl.max_lat_nx, ph.nX);
          pl.max_lat_ny = std::max(pl.max_lat_ny, ph.nY);
          pl.max_lat_tasks = std::max(pl.max_lat_tasks, ph.nY * ((ph.nX + 31) / 32));
          if (s.stage == 0) {
            h->last_coarse_table_off[s.idx] = pl.tab[tid].out_off;
            h->last_coarse_nA[s.idx] = ph.nA;
            h->last_coarse_P[s.idx] = s.P;
            h->last_coarse_Ppad[s.idx] = Ppad;
          }
        } else {
          pl.fine.push_back(pid);
          h->work[4] += (int64_t)(ph.nX * ph.nY + 1) * ph.nA * s.P;
          pl.max_fine_poses = std::max(pl.max_fine_poses, ph.nX * ph.nY * ph.nA);
          if (ph.nX != 3 || ph.nY != 3) pl.fine_not9 = true;
        }
      }
      if (spec) {
        // speculative fine passes, one per coarse pass: everything but the search centre is known;
        // for each possible winning coarse angle the host tabulates the heading atan2 would return
        // and the fine search angles around it (cos/sin of the angle and of its normalisation).
        const int ncoarse = (int)pl.pass.size();
        pl.first_spec_table = (int)pl.tab.size();
        for (int pid = 0; pid < ncoarse; pid++) {
          if (pl.ph[pid].fine) continue;
          const PassHost cph = pl.ph[pid];
          MatchState& s = states[cph.match];
          PassDev& cpd = pl.pass[pid];
          const int nA = cph.nA;
          const double fo = 0.5 * h->prm.coarse_angle_resolution, fr = h->prm.fine_search_angle_resolution;
          const int h_off = (int)pl.trig.size();
          pl.trig.resize(pl.trig.size() + (size_t)((nA + 1) & ~1));
          const int t_off = (int)(pl.trig.size() / 2);
          pl.trig.resize(pl.trig.size() + (size_t)2 * nA * nAf);
          const int ht_off = (int)(pl.trig.size() / 2);
          pl.trig.resize(pl.trig.size() + (size_t)2 * nA * nAf);
          const double* ctrig = pl.trig.data() + 2 * (size_t)cpd.htrig_off;
          for (int a = 0; a < nA; a++) {
            const double heading = atan2(ctrig[2 * a + 1] / 1.0, ctrig[2 * a] / 1.0);
            pl.trig[h_off + a] = heading;
            const double start_angle = heading - fo;
            for (int f = 0; f < nAf; f++) {
              const double angle = start_angle + (double)(uint32_t)f * fr;
              const double ca = cos(angle), sa = sin(angle);
              pl.trig[2 * ((size_t)t_off + (size_t)a * nAf + f)] = ca;
              pl.trig[2 * ((size_t)t_off + (size_t)a * nAf + f) + 1] = sa;
              const double hn = h_normalize_angle(angle);
              const bool same = dbits(hn) == dbits(angle);  // (the usual case: nothing to re-evaluate)
              pl.trig[2 * ((size_t)ht_off + (size_t)a * nAf + f)] = same ? ca : cos(hn);
              pl.trig[2 * ((size_t)ht_off + (size_t)a * nAf + f) + 1] = same ? sa : sin(hn);
            }
          }
          const int Ppad = align_up(s.P, 4);
          TableDev t;
          t.q_start = b->scan_start[s.q];
          t.P = s.P; t.Ppad = Ppad; t.nA = nAf;
          t.trig_off = t_off;  // set by k_reduce to the winner's block
          t.out_off = (int)pl.off_elems;
          pl.off_elems += (size_t)nAf * Ppad;
          t.px = s.pose[0]; t.py = s.pose[1];
          fill_inverse_rotation(t, s.pose);
          t.gox = s.gox; t.goy = s.goy;
          const int tid = (int)pl.tab.size();
          pl.tab.push_back(t);
          PassDev fd;
          memset(&fd, 0, sizeof(fd));
          fd.slot = s.slot; fd.table = tid; fd.nA = nAf;
          fd.offx = crx * 0.5; fd.offy = crx * 0.5; fd.resx = h->res_eff; fd.resy = h->res_eff;
          fd.nX = n_steps(fd.offx, fd.resx); fd.nY = n_steps(fd.offy, fd.resy);
          fd.P = s.P; fd.Ppad = Ppad; fd.fine = 1; fd.penalize = b->do_penalize ? 1 : 0;
          fd.sums_off = (int)pl.sums_elems;
          pl.sums_elems += (size_t)fd.nX * fd.nY * nAf;
          fd.htrig_off = ht_off;
          fd.ang_off = pl.ang_elems;
          pl.ang_elems += nAf;
          fd.spec = -1;
          fd.cmax_off = -1;
          fd.angle_offset = fo; fd.angle_res = fr;
          fd.gox = s.gox; fd.goy = s.goy;
          const int fid = (int)pl.pass.size();
          PassHost fph;
          fph.match = cph.match; fph.fine = true; fph.spec = true;
          fph.cx = fph.cy = fph.ch = 0.0;  // known after the coarse pass
          fph.offx = fd.offx; fph.offy = fd.offy; fph.resx = fd.resx; fph.resy = fd.resy;
          fph.angle_offset = fo; fph.angle_res = fr;
          fph.nA = nAf; fph.nX = fd.nX; fph.nY = fd.nY; fph.ang_off = fd.ang_off;
          pl.pass.push_back(fd);
          pl.ph.push_back(fph);
          pl.spec_of.push_back(-1);
          pl.spec_h.push_back(-1);
          pl.spec_fine.push_back(fid);
          pl.spec_of[pid] = fid;
          pl.spec_h[pid] = h_off;
          PassDev& c2 = pl.pass[pid];
          c2.spec = fid; c2.spec_nAf = nAf; c2.spec_h_off = h_off; c2.spec_trig_off = t_off; c2.spec_htrig_off = ht_off;
          pl.max_fine_poses = std::max(pl.max_fine_poses, fd.nX * fd.nY * nAf);
          if (fd.nX != 3 || fd.nY != 3) pl.fine_not9 = true;
        }
      }
      return YSM_OK;
    };

    // ---- staging + upload --------------------------------------------------------------------
    // wave_static: also carry the pool / scan directory / match descriptors (latency path)
    BlobLayout L;
    auto stage_blob = [&](bool wave_static, bool with_passes, PinBuf& hb, DevBuf& db, bool upload = true) -> int {
      size_t o = 0;
      if (wave_static) {
        if (small) {
          if (!b->pool_on_device) { L.pool = o; o = a16(o + (size_t)b->n_points * 16); }
          L.scan_start = o; o = a16(o + (size_t)b->n_scans * 4);
          L.scan_count = o; o = a16(o + (size_t)b->n_scans * 4);
        }
        L.matches = o; o = a16(o + sizeof(MatchDev) * (size_t)nw);
        L.base = o; o = a16(o + hbase.size() * 4);
        L.workcount = o; o = a16(o + 16);
        if (small) { L.scanlist = o; o = a16(o + sizeof(ScanRef) * hscans.size()); }
      }
      if (with_passes) {
        L.tab = o; o = a16(o + sizeof(TableDev) * pl.tab.size());
        L.pass = o; o = a16(o + sizeof(PassDev) * pl.pass.size());
        L.pa = o; o = a16(o + sizeof(PassAngle) * pl.pa.size());
        L.fine = o; o = a16(o + sizeof(int) * (pl.fine.size() + pl.spec_fine.size()));
        L.trig = o; o = a16(o + sizeof(double) * pl.trig.size());
        L.pmax = o; o = a16(o + sizeof(double) * pl.pass.size());
      }
      L.total = std::max<size_t>(o, 16);
      CK(hb.ensure(L.total));
      CK(db.ensure(L.total));
      char* p = (char*)hb.p;
      if (wave_static) {
        if (small) {
          if (!b->pool_on_device && b->n_points > 0) memcpy(p + L.pool, b->pool_xy, (size_t)b->n_points * 16);
          if (b->n_scans > 0) {
            memcpy(p + L.scan_start, b->scan_start, (size_t)b->n_scans * 4);
            memcpy(p + L.scan_count, b->scan_count, (size_t)b->n_scans * 4);
          }
        }
        memcpy(p + L.matches, hm.data(), sizeof(MatchDev) * (size_t)nw);
        if (!hbase.empty()) memcpy(p + L.base, hbase.data(), hbase.size() * 4);
        memset(p + L.workcount, 0, 16);
        if (small && !hscans.empty()) memcpy(p + L.scanlist, hscans.data(), sizeof(ScanRef) * hscans.size());
      }
      if (with_passes) {
        memcpy(p + L.tab, pl.tab.data(), sizeof(TableDev) * pl.tab.size());
        memcpy(p + L.pass, pl.pass.data(), sizeof(PassDev) * pl.pass.size());
        if (!pl.pa.empty()) memcpy(p + L.pa, pl.pa.data(), sizeof(PassAngle) * pl.pa.size());
        if (!pl.fine.empty()) memcpy(p + L.fine, pl.fine.data(), sizeof(int) * pl.fine.size());
        if (!pl.spec_fine.empty())
          memcpy(p + L.fine + sizeof(int) * pl.fine.size(), pl.spec_fine.data(), sizeof(int) * pl.spec_fine.size());
        memcpy(p + L.trig, pl.trig.data(), sizeof(double) * pl.trig.size());
        memset(p + L.pmax, 0, sizeof(double) * pl.pass.size());  // per-pass best response, max-accumulated on the GPU
      }
      if (upload) CK(cudaMemcpyAsync(db.p, hb.p, L.total, cudaMemcpyHostToDevice, st));
      h->work[6] += (int64_t)L.total;
      if (wave_static) {
        const char* d = (const char*)db.p;
        if (small) {
          if (!b->pool_on_device) d_pool = (const double*)(d + L.pool);
          d_scan_start = (const int*)(d + L.scan_start);
          d_scan_count = (const int*)(d + L.scan_count);
        }
        d_matches = (const MatchDev*)(d + L.matches);
        d_base_idx = (const int*)(d + L.base);
        d_workcount = (int*)((char*)db.p + L.workcount);
        h->cur_matches = d_matches;
        h->cur_workcount = d_workcount;
      }
      return YSM_OK;
    };

    // ---- K1: grid build ------------------------------------------------------------------------
    auto launch_build = [&]() -> int {
      if (timing) CK(cudaEventRecord(h->ev[0], st));
      // exact per-tile candidate lists (16-bit counters / cursors in shared memory): not for the ordered
      // (wide-smear) filter, which edits the cell list afterwards, nor for grids / matches that overflow them
      const bool use_cand = !h->ordered_stamps && !(h->debug & YSM_DEBUG_NO_CANDLISTS) &&
                            max_match_cells * tiles_per_stamp <= 65535 && (size_t)tiles_per_grid * 2 <= 64 * 1024;
      if (use_cand) {
        CK(h->d_cand.ensure(std::max<size_t>(16, (size_t)cells_total * tiles_per_stamp * 4)));
        CK(h->d_wcand.ensure(std::max<size_t>(16, (size_t)work_cap * 8)));
        CK(h->d_gbox.ensure(std::max<size_t>(8, (size_t)cells_total * 4)));  // second half of the cells' list ranks
      }
      if (h->static_grid) {  // the map grid is resident: nothing to build
        if (timing) CK(cudaEventRecord(h->ev[1], st));
        return YSM_OK;
      }
      const size_t bits_bytes = use_cand ? (size_t)((tiles_per_grid + 1) / 2) * 4   // a 16-bit counter per tile
                                         : (size_t)((tiles_per_grid + 3) / 4) * 4;  // one byte per tile
      const size_t fixed = 16 * (size_t)nbase_max + bits_bytes;
      // small waves: one warp per base scan (up to 32) so the scans are filtered concurrently
      // (large waves: a warp per base scan up to 16, so no warp sits out a second round of scans)
      int nwarps = nw >= 2 * h->num_sms ? std::min(16, std::max(8, nbase_max)) : std::min(32, std::max(8, nbase_max));
      while (nwarps > 1 && (size_t)nwarps * 4 * pmax + fixed > 200 * 1024) nwarps >>= 1;
      size_t smem = (size_t)nwarps * 4 * pmax + fixed;
      int stage = 0;
      if (nw < 2 * h->num_sms) {  // latency mode: stage scan points in shared memory if they fit
        const size_t with_pts = ((smem + 15) & ~(size_t)15) + (size_t)nwarps * 16 * pmax;
        if (with_pts <= 200 * 1024) {
          smem = with_pts;
          stage = 1;
        }
      }
      if (smem > 220 * 1024) return fail(h, YSM_EUNSUP, "too many base scans / points / tiles for the filter kernel");
      k_find_valid<<<nw, nwarps * 32, smem, st>>>(g, d_matches, d_base_idx, d_scan_start, d_scan_count, d_pool,
                                                   (uint32_t*)h->d_ptcell.p, (uint32_t*)h->d_cells.p,
                                                   (int*)h->d_cellcount.p, (uint2*)h->d_gbox.p, (int2*)h->d_work.p,
                                                   d_workcount, pmax, nbase_max, stage,
                                                   use_cand ? (uint32_t*)h->d_cand.p : nullptr,
                                                   use_cand ? (uint2*)h->d_wcand.p : nullptr, tps1);
      h->launches++;
      kt.mark("k_find_valid");
      if (h->ordered_stamps) {
        int log2cap = 6;
        while ((1ll << log2cap) < 2 * max_match_cells) log2cap++;
        const size_t osmem = (size_t)4 << log2cap;
        if (osmem > 200 * 1024)
          return fail(h, YSM_EUNSUP, "too many base points per match for the ordered-stamp filter (wide smear)");
        k_stamp_order<<<nw, 32, osmem, st>>>(d_matches, (uint32_t*)h->d_cells.p, (const int*)h->d_cellcount.p, log2cap);
        h->launches++;
        kt.mark("k_stamp_order");
      }
      const size_t ksmem = tile_stamp_smem(g.K, g.Wt, 8);
      if (ksmem > 200 * 1024) return fail(h, YSM_EUNSUP, "smear kernel too large for the stamping kernel");
      const long long ctas = std::max<long long>(1, std::min<long long>(work_cap, (long long)h->num_sms * 8));
      const size_t lsmem = tile_stamp_lists_smem(g.K, g.Wt, 8);
      if (use_cand && stamp_lists_fit(g.K, g.Wt) && lsmem <= 100 * 1024 && !(h->debug & YSM_DEBUG_NO_HALF_LISTS)) {
        // throughput form: per-column-group, per-tile-half step lists
        const long long lctas = std::max<long long>(1, std::min<long long>((work_cap + 7) / 8, (long long)h->num_sms * 4));
        k_tile_stamp_lists<<<(unsigned)lctas, 256, lsmem, st>>>(g, d_matches, (const int2*)h->d_work.p, d_workcount,
                                                                h->d_stamp_tab, h->d_grids, h->d_rowmask, h->rm_words,
                                                                (const uint32_t*)h->d_cand.p, (const uint2*)h->d_wcand.p);
      } else
      k_tile_stamp<<<(unsigned)ctas, 256, ksmem, st>>>(g, d_matches, (const uint32_t*)h->d_cells.p,
                                                       (const int*)h->d_cellcount.p, (const uint2*)h->d_gbox.p,
                                                       (const int2*)h->d_work.p, d_workcount, h->d_stamp_tab, h->d_grids,
                                                       h->d_rowmask, h->rm_words,
                                                       use_cand ? (const uint32_t*)h->d_cand.p : nullptr,
                                                       use_cand ? (const uint2*)h->d_wcand.p : nullptr);
      h->launches++;
      kt.mark("k_tile_stamp");
      if (timing) CK(cudaEventRecord(h->ev[1], st));
      CK(cudaGetLastError());
      if (timing) {  // P_valid of the wave (roofline accounting): the cells FindValidPoints + the ROI test kept
        std::vector<int> cc((size_t)nw);
        CK(cudaMemcpyAsync(cc.data(), h->d_cellcount.p, (size_t)nw * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (int v : cc) h->work[15] += v;
      }
      return YSM_OK;
    };

    // ---- pass iterations ------------------------------------------------------------------------
    int iter = 0;
    bool built = false;
    while (true) {
      const bool spec = speculate && iter == 0;
      if (!small && !built) {
        // throughput path: descriptors + build first, so the build overlaps the host's pass planning
        int rc = stage_blob(true, false, h->h_wblob, h->d_wblob);
        if (rc != YSM_OK) return rc;
        rc = launch_build();
        if (rc != YSM_OK) return rc;
        built = true;
      }
      {
        const int rc = plan_passes(spec);
        if (rc != YSM_OK) return rc;
      }
      const int npass = (int)pl.pass.size();
      if (npass == 0) break;
      if (pl.off_elems > 0x7fffff00ull || pl.sums_elems > 0x7fffff00ull)
        return fail(h, YSM_ENOMEM, "wave workspace exceeds 2^31 elements; lower max_slots");
      tr.mark("pass plan (host libm)");
      const char* db = nullptr;
      // latency path, first iteration: everything in ONE cooperative kernel (k_match_small)
      bool mega = iter == 0 && !built && small && !timing && !h->static_grid && !(h->debug & (YSM_DEBUG_NO_MEGA | YSM_DEBUG_KEEP_GRIDS)) &&
                  !pl.pa.empty() && pl.fine.empty() && nbase_max <= 64 && hscans.size() <= 512;
      int mega_tpc = 0, mega_psplit = 1, mega_chunks = 0, mega_stage = 0, mega_fvw = 0, mega_log2cap = 6;
      size_t mega_smem = 0;
      if (mega) {
        // phase shapes for a 512-thread CTA (16 warps)
        const size_t fv = fv_scan_smem(pmax);
        (void)mega_fvw; (void)mega_stage;
        size_t so = 0;
        if (h->ordered_stamps) {
          while ((1ll << mega_log2cap) < 2 * max_match_cells) mega_log2cap++;
          so = (size_t)4 << mega_log2cap;
        }
        // sweep: tpc row-tasks x psplit point slices = 16 warps; every CTA sees whole task groups
        bool uniform = true;
        for (const PassHost& q : pl.ph)
          if (!q.fine && q.nY * ((q.nX + 31) / 32) != pl.max_lat_tasks) uniform = false;
        mega_tpc = 16;
        if (uniform) {
          for (int t : {2, 4, 8, 1})
            if (pl.max_lat_tasks % t == 0 && pl.max_lat_P / 24 >= 16 / t) {
              mega_tpc = t;
              break;
            }
        }
        mega_psplit = 16 / mega_tpc;
        if (mega_psplit > 1 && (!uniform || pl.max_lat_tasks % mega_tpc != 0)) { mega_tpc = 16; mega_psplit = 1; }
        mega_chunks = (pl.max_lat_tasks + mega_tpc - 1) / mega_tpc;
        const size_t sw = (size_t)(((pl.max_lat_P + 7) & ~7) + pl.max_lat_nx + pl.max_lat_ny + (mega_psplit > 1 ? 512 : 0)) * 4;
        mega_smem = std::max(std::max(fv, so), std::max(tile_stamp_smem(g.K, g.Wt, 16), sw));
        mega_smem = std::max(mega_smem, (size_t)(pmax + 8) * 4);
        mega_smem = (mega_smem + 1023) & ~(size_t)1023;
        if (mega_smem > 110 * 1024) mega = false;
      }
      if (mega) {
        if (mega_smem != h->mega_occ_smem) {
          h->mega_occ_smem = mega_smem;
          int occ = 0;
          CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_match_small, 512, mega_smem));
          h->mega_ctas_per_sm = std::min(occ, 2);
        }
        if (h->mega_ctas_per_sm < 1) mega = false;
      }
      if (mega) {
        int rc = stage_blob(true, true, h->h_wblob, h->d_wblob, /*upload=*/false);
        if (rc != YSM_OK) return rc;
        db = (const char*)h->d_wblob.p;
        built = true;
      } else if (!built) {
        // latency path: ONE copy carries the wave-static data and the first iteration's pass tables
        int rc = stage_blob(true, true, h->h_wblob, h->d_wblob);
        if (rc != YSM_OK) return rc;
        db = (const char*)h->d_wblob.p;
        rc = launch_build();
        if (rc != YSM_OK) return rc;
        built = true;
      } else {
        const int rc = stage_blob(false, true, h->h_blob, h->d_blob);
        if (rc != YSM_OK) return rc;
        db = (const char*)h->d_blob.p;
      }
      const int nhostfine = (int)pl.fine.size(), nspec = (int)pl.spec_fine.size();
      const int ncoarse_total = npass - nspec;  // passes scheduled by the host (coarse + host fine)
      h->work[7] += (int64_t)(sizeof(PassOut) * (size_t)npass + (size_t)pl.ang_elems * 4);
      h->work[2] += (int64_t)pl.off_elems;
      h->work[3] += (int64_t)pl.sums_elems;
      TableDev* d_tab = (TableDev*)(db + L.tab);
      PassDev* d_pass = (PassDev*)(db + L.pass);
      const PassAngle* d_pa = (const PassAngle*)(db + L.pa);
      const int* d_fine = (const int*)(db + L.fine);
      const double* d_trig = (const double*)(db + L.trig);
      double* d_pmax = (double*)(db + L.pmax);

      CK(h->d_offsets.ensure(std::max<size_t>(16, pl.off_elems * 4)));
      CK(h->d_sums.ensure(std::max<size_t>(16, pl.sums_elems * 8)));  // penalised responses, f64 [iy][ix][a]
      CK(h->d_outs.ensure(sizeof(PassOut) * (size_t)npass));
      CK(h->d_angsums.ensure(std::max<size_t>(16, (size_t)pl.ang_elems * 4)));
      CK(h->d_cellmax.ensure(std::max<size_t>(16, pl.cmax_elems * 8)));
      CK(h->h_outs.ensure(sizeof(PassOut) * (size_t)npass));
      CK(h->h_angsums.ensure(std::max<size_t>(16, (size_t)pl.ang_elems * 4)));
      tr.mark("blob H2D");

      if (mega) {
        CK(h->h_flags.ensure((size_t)npass * 4 + 256));
        if (!h->h_wblob.dptr || !h->h_outs.dptr || !h->h_angsums.dptr || !h->h_flags.dptr)
          return fail(h, YSM_ECUDA, "mapped host memory is not available");
        h->epoch++;
        for (int pid = 0; pid < npass; pid++) ((volatile int*)h->h_flags.p)[pid] = 0;
        SmallArgs A;
        memset(&A, 0, sizeof(A));
        A.blob_src = (const uint4*)h->h_wblob.dptr;
        A.blob_dst = (uint4*)h->d_wblob.p;
        A.blob_vec = (int)((L.total + 15) / 16);
        A.pool_in_blob = b->pool_on_device ? 0 : 1;
        A.pool_dev = b->pool_on_device ? b->pool_xy : nullptr;
        A.o_pool = (unsigned)L.pool; A.o_scan_start = (unsigned)L.scan_start; A.o_scan_count = (unsigned)L.scan_count;
        A.o_matches = (unsigned)L.matches; A.o_base = (unsigned)L.base; A.o_workcount = (unsigned)L.workcount;
        A.o_tab = (unsigned)L.tab; A.o_pass = (unsigned)L.pass; A.o_pa = (unsigned)L.pa; A.o_trig = (unsigned)L.trig;
        A.o_pmax = (unsigned)L.pmax;
        A.o_scanlist = (unsigned)L.scanlist;
        A.nscans_total = (int)hscans.size();
        A.tiles_per_grid = (tiles_per_grid + 3) & ~3;
        CK(h->d_scan_emit.ensure(std::max<size_t>(16, hscans.size() * 4)));
        CK(h->d_tileflag.ensure((size_t)nw * A.tiles_per_grid + 16));
        A.scan_emit = (int*)h->d_scan_emit.p;
        A.tileflag = (unsigned char*)h->d_tileflag.p;
        A.nw = nw; A.npa = (int)pl.pa.size(); A.ncoarse = ncoarse_total; A.nspec = nspec; A.nAf = std::max(1, nAf);
        A.pmax = pmax; A.nbase_max = nbase_max; A.stage = mega_stage; A.fv_warps = mega_fvw;
        A.ordered = h->ordered_stamps ? 1 : 0; A.log2cap = mega_log2cap;
        A.tpc = mega_tpc; A.psplit = mega_psplit; A.task_chunks = mega_chunks;
        A.ptcell = (uint32_t*)h->d_ptcell.p; A.cells = (uint32_t*)h->d_cells.p; A.cellcount = (int*)h->d_cellcount.p;
        A.gbox = (uint2*)h->d_gbox.p; A.work = (int2*)h->d_work.p;
        A.stamp_tab = h->d_stamp_tab; A.grids = h->d_grids; A.rowmask = h->d_rowmask; A.rm_words = h->rm_words;
        A.epoch = h->epoch;
        A.offsets = (int*)h->d_offsets.p; A.resp = (double*)h->d_sums.p;
        A.cellmax = (unsigned long long*)h->d_cellmax.p; A.cellmax_n = (int)pl.cmax_elems;
        A.outs_host = (PassOut*)h->h_outs.dptr; A.angs_host = (int*)h->h_angsums.dptr;
        A.flags_host = (volatile int*)h->h_flags.dptr;
        unsigned long long* ts_host = nullptr;
        if (tr.on) {
          ts_host = (unsigned long long*)((char*)h->h_flags.p + (((size_t)npass * 4 + 15) & ~(size_t)15));
          A.tstamps = (unsigned long long*)((char*)h->h_flags.dptr + (((size_t)npass * 4 + 15) & ~(size_t)15));
        }
        GridC gg = g;
        PenaltyC pp = h->pen;
        void* kargs[] = {&gg, &pp, &A};
        const int ctas = h->num_sms * h->mega_ctas_per_sm;
        CK(cudaLaunchCooperativeKernel((const void*)k_match_small, dim3(ctas), dim3(512), kargs, mega_smem, st));
        h->launches++;
        h->work[12]++;
        // The tile clear is queued behind the kernel right away: the host would only spin on the flags
        // meanwhile, so the launch leaves the call's critical path. A match the kernel could not finish
        // (tied coarse winners, response expansion) has its grid rebuilt by the general path in the next
        // iteration (built = false; the clear after the loop then belongs to that rebuild).
        clear_wave(h, d_matches, st);
        built = false;
        tr.mark("latency kernel launch");
        // wait for the per-pass completion flags the kernel writes after its results
        volatile int* flags = (volatile int*)h->h_flags.p;
        const auto t_start = std::chrono::steady_clock::now();
        long spins = 0;
        for (int cp = 0; cp < ncoarse_total; cp++) {
          const int pid = pl.spec_of[cp] >= 0 ? pl.spec_of[cp] : cp;  // the pass that publishes last for this match
          while (flags[pid] != h->epoch) {
            if ((++spins & 0xFFFF) == 0) {
              if (cudaStreamQuery(st) != cudaErrorNotReady) break;  // finished (or failed) without the flag
              if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() > 30.0) break;
            }
          }
          if (flags[pid] != h->epoch) {
            cudaError_t e = cudaStreamSynchronize(st);
            if (e == cudaSuccess) e = cudaGetLastError();
            if (e != cudaSuccess) return fail(h, YSM_ECUDA, std::string("latency kernel: ") + cudaGetErrorString(e));
            if (flags[pid] != h->epoch) return fail(h, YSM_ECUDA, "latency kernel finished without publishing its results");
          }
        }
        tr.mark("poll");
        if (ts_host) {
          cudaStreamSynchronize(st);
          static const char* names[] = {"P0 blob copy", "P1a+P1b", "sync", "P2 stamp", "sync", "P3 sweep", "sync",
                                        "P4a reduce", "sync", "P4b fine sweep", "sync", "P4c fine reduce"};
          for (int k = 0; k < 12; k++)
            fprintf(stderr, "[ysm-kernel] %-14s %8.1f us\n", names[k], (double)(ts_host[k + 1] - ts_host[k]) * 1e-3);
          fprintf(stderr, "[ysm-kernel]   P1a scans %.1f us, sync %.1f us, P1b %.1f us\n", (double)(ts_host[13] - ts_host[1]) * 1e-3,
                  (double)(ts_host[14] - ts_host[13]) * 1e-3, (double)(ts_host[2] - ts_host[14]) * 1e-3);
        }
      } else {
      // ---- K2 offsets (tables of the passes the host scheduled) ------------------------------------
      if (pl.cmax_elems > 0) CK(cudaMemsetAsync(h->d_cellmax.p, 0, pl.cmax_elems * 8, st));
      const int ntab_host = spec ? pl.first_spec_table : (int)pl.tab.size();
      // The pruned sweep computes its own lookup offsets (ComputeOffsets fused): the tables are only needed by
      // the unpruned lattice sweep, the fine sweep / angular covariance, and the offset introspection of tests.
      const bool will_prune = !pl.pa.empty() && !(h->debug & YSM_DEBUG_NO_PRUNE) &&
                              (long long)pl.pa.size() * ((pl.max_lat_ny + 27) / 28) * ((pl.max_lat_nx + 31) / 32) >= h->num_sms * 2;
      const bool need_tables = !will_prune || nhostfine > 0 || (h->debug & YSM_DEBUG_KEEP_GRIDS);
      if (need_tables) {
        int maxwork = 1;
        for (int t = 0; t < ntab_host; t++) maxwork = std::max(maxwork, pl.tab[t].nA * (pl.tab[t].Ppad / 4));
        dim3 grid((maxwork + 255) / 256, (unsigned)ntab_host);
        k_offsets<<<grid, 256, 0, st>>>(g, d_tab, d_trig, d_pool, (int*)h->d_offsets.p, 0);
        h->launches++;
        kt.mark("k_offsets");
      }
      // ---- K3 sweeps ----------------------------------------------------------------------------
      if (timing) CK(cudaEventRecord(h->ev[2], st));
      if (!pl.pa.empty()) {
        const int npa = (int)pl.pa.size();
        const int target = h->num_sms * 2;
        static const int sweep_rows = getenv("YSM_SWEEP_ROWS") ? std::max(1, std::min(28, atoi(getenv("YSM_SWEEP_ROWS")))) : 28;
        const int nrg = (pl.max_lat_ny + sweep_rows - 1) / sweep_rows, rows_per_cta = (pl.max_lat_ny + nrg - 1) / nrg;
        const int nxc = (pl.max_lat_nx + 31) / 32, cw = (pl.max_lat_nx + nxc - 1) / nxc;
        const bool pruned = !(h->debug & YSM_DEBUG_NO_PRUNE) && (long long)npa * nrg * nxc >= target;
        if (pruned) {
          // throughput form: zero-row pruning, offsets fused (k_sweep_pruned)
          // shared memory per CTA decides the L1 the lookups get: 2 CTAs x 82 KB leave 60 KB of L1, 2 x 41 KB 156 KB
          static const int sweep_kb = getenv("YSM_SWEEP_SMEM_KB") ? atoi(getenv("YSM_SWEEP_SMEM_KB")) : 96;
          int PB = (int)((sweep_kb * 1024 / 4 / (2 + rows_per_cta)) & ~31);
          PB = std::max(32, std::min(PB, (pl.max_lat_P + 31) & ~31));
          const size_t smem = (size_t)(2 + rows_per_cta) * PB * 4;
          dim3 grid(npa, nrg * nxc, 1);
          k_sweep_pruned<<<grid, 32 * rows_per_cta, smem, st>>>(g, h->pen, d_pass, d_pa, d_tab, d_trig, d_pool, h->d_grids,
                                                               h->d_rowmask, h->rm_words, h->tnx, (double*)h->d_sums.p,
                                                               d_pmax, (unsigned long long*)h->d_cellmax.p, rows_per_cta, cw, PB,
                                                               timing ? h->d_issued : nullptr, h->d_dpc, h->dpc_nx);
          h->work[8]++;
        } else {
          // one warp per lattice row-task; small batches: fewer row-tasks per CTA and several warps
          // per row-task (point slices) until the machine is full
          int tpc = std::min(pl.max_lat_tasks, 32);
          int task_chunks = (pl.max_lat_tasks + tpc - 1) / tpc;
          int psplit = 1;
          if (npa * task_chunks < target) {
            const int want = (target + npa - 1) / npa;            // CTAs wanted per (pass, angle)
            tpc = std::max(2, std::min(tpc, (pl.max_lat_tasks + want - 1) / want));
            task_chunks = (pl.max_lat_tasks + tpc - 1) / tpc;
            psplit = std::max(1, std::min(std::min(8, 32 / tpc), pl.max_lat_P / 64));
            // every CTA must see exactly one task iteration per warp group (barriers inside the loop)
            bool uniform = true;
            for (const PassHost& q : pl.ph)
              if (!q.fine && q.nY * ((q.nX + 31) / 32) != pl.max_lat_tasks) uniform = false;
            if (!uniform || pl.max_lat_tasks % tpc != 0) psplit = 1;
          }
          const int threads = 32 * std::min(tpc, 32) * psplit;
          const size_t smem = (size_t)(((pl.max_lat_P + 7) & ~7) + pl.max_lat_nx + pl.max_lat_ny + (psplit > 1 ? threads : 0)) * 4;
          if (smem > 200 * 1024) return fail(h, YSM_EUNSUP, "search lattice too large for the sweep kernel");
          dim3 grid(npa, task_chunks, 1);
          k_sweep_lattice<<<grid, threads, smem, st>>>(g, h->pen, d_pass, d_pa, d_tab, (const int*)h->d_offsets.p,
                                                       h->d_grids, (double*)h->d_sums.p, d_pmax,
                                                       (unsigned long long*)h->d_cellmax.p, tpc, psplit);
        }
        h->launches++;
        h->work[1]++;
        kt.mark("k_sweep_lattice");
      }
      if (timing) CK(cudaEventRecord(h->ev[3], st));
      bool fine9_ran = false;
      if (nhostfine > 0) {
        dim3 grid((pl.max_fine_poses + 7) / 8, (unsigned)nhostfine);
        if (!pl.fine_not9 && nhostfine >= 64 && !(h->debug & YSM_DEBUG_NO_FINE9)) {
          // waves of 3 x 3 fine passes: a warp per (pass, angle), nine cells per offset; the integer sums are kept
          // for the reduce's angular-covariance sums
          CK(h->d_isums.ensure(std::max<size_t>(16, pl.sums_elems * 4)));
          k_sweep_fine9<<<(unsigned)nhostfine, 32 * std::min(12, std::max(1, nAf)), 0, st>>>(
              g, h->pen, d_pass, d_fine, d_tab, (const int*)h->d_offsets.p, h->d_grids, (double*)h->d_sums.p, d_pmax,
              (unsigned*)h->d_isums.p);
          fine9_ran = true;
        } else
        k_sweep_points<<<grid, 256, 0, st>>>(g, h->pen, d_pass, d_fine, d_tab, (const int*)h->d_offsets.p, h->d_grids,
                                             (double*)h->d_sums.p, d_pmax);
        h->launches++;
        kt.mark("k_sweep_points");
      }
      // ---- K3b/K4 reduce ------------------------------------------------------------------------
      if (timing) CK(cudaEventRecord(h->ev[4], st));
      k_reduce<<<ncoarse_total, 512, 0, st>>>(g, d_pass, d_tab, (const int*)h->d_offsets.p, (const double*)h->d_sums.p,
                                             d_pmax, (const unsigned long long*)h->d_cellmax.p, d_trig, h->d_grids,
                                             (PassOut*)h->d_outs.p, (int*)h->d_angsums.p, 0,
                                             fine9_ran ? (const unsigned*)h->d_isums.p : nullptr);
      h->launches++;
      kt.mark("k_reduce");
      if (nspec > 0) {
        // speculative fine passes: their centre / angle tables were selected by the reduce above
        int maxwork = 1;
        for (size_t t = (size_t)pl.first_spec_table; t < pl.tab.size(); t++)
          maxwork = std::max(maxwork, pl.tab[t].nA * (pl.tab[t].Ppad / 4));
        dim3 og((maxwork + 255) / 256, (unsigned)(pl.tab.size() - (size_t)pl.first_spec_table));
        k_offsets<<<og, 256, 0, st>>>(g, d_tab, d_trig, d_pool, (int*)h->d_offsets.p, pl.first_spec_table);
        dim3 sg((pl.max_fine_poses + 7) / 8, (unsigned)nspec);
        k_sweep_points<<<sg, 256, 0, st>>>(g, h->pen, d_pass, d_fine + nhostfine, d_tab, (const int*)h->d_offsets.p,
                                           h->d_grids, (double*)h->d_sums.p, d_pmax);
        k_reduce<<<nspec, 512, 0, st>>>(g, d_pass, d_tab, (const int*)h->d_offsets.p, (const double*)h->d_sums.p,
                                       d_pmax, (const unsigned long long*)h->d_cellmax.p, d_trig, h->d_grids,
                                       (PassOut*)h->d_outs.p, (int*)h->d_angsums.p, ncoarse_total, nullptr);
        h->launches += 3;
        kt.mark("speculative fine");
      }
      if (timing) CK(cudaEventRecord(h->ev[5], st));
      CK(cudaMemcpyAsync(h->h_outs.p, h->d_outs.p, sizeof(PassOut) * (size_t)npass, cudaMemcpyDeviceToHost, st));
      if (pl.ang_elems > 0)
        CK(cudaMemcpyAsync(h->h_angsums.p, h->d_angsums.p, (size_t)pl.ang_elems * 4, cudaMemcpyDeviceToHost, st));
      tr.mark("pass launches");
      kt.mark("d2h");
      CK(cudaStreamSynchronize(st));
      tr.mark("sync");
      CK(cudaGetLastError());
      }  // !mega
      if (timing) {
        float ms = 0;
        if (iter == 0) {
          cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
          h->t_build += ms;
        }
        if (!pl.pa.empty()) {
          cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]);
          h->t_sweep += ms;
        }
        cudaEventElapsedTime(&ms, h->ev[4], h->ev[5]);
        h->t_reduce += ms;
      }

      // ---- host: finish every pass exactly as CorrelateScan / MatchScan do ----------------------
      const PassOut* outs = (const PassOut*)h->h_outs.p;
      const int* angs = (const int*)h->h_angsums.p;
      for (int pid = 0; pid < ncoarse_total; pid++) {
        const PassHost& ph = pl.ph[pid];
        const PassOut& po = outs[pid];
        MatchState& s = states[ph.match];
        s.n_passes++;
        s.n_ties = po.n_ties;
        if (po.n_ties <= 0) {
          s.status = YSM_EMATCH;  // "Mapper FATAL ERROR - Unable to find best position"
          s.stage = 5;
          continue;
        }
        const double heading = atan2(po.ty, po.tx);
        if (!ph.fine) {
          finalize_positional(h, ph, po, s.cov);
        } else {
          finalize_angular(ph, po, heading, angs + ph.ang_off, s.P, s.cov);
        }
        s.mean[0] = po.avg_x; s.mean[1] = po.avg_y; s.mean[2] = heading;
        double best = po.best;
        if (best > 1.0) best = 1.0;
        s.best = best;
        // MatchScan schedule
        if (ph.fine) {
          s.stage = 5;
        } else {
          bool expand = false;
          if (h->prm.use_response_expansion && h_double_equal(best, 0.0) && s.stage < 3) {
            // stage 0 -> expansion 1, ..., stage 2 -> expansion 3
            expand = true;
          }
          if (expand) {
            s.stage += 1;
            s.angle_offset_cur += 20 * KT_PI_180;
          } else {
            s.stage = b->do_refine ? 4 : 5;
          }
          // the fine pass already ran on the device behind this coarse pass?
          const int fid = pl.spec_of[pid];
          if (fid >= 0 && s.stage == 4 && outs[fid].n_ties > 0) {
            const PassOut& fo = outs[fid];
            PassHost fph = pl.ph[fid];
            // the centre k_reduce gave the fine pass: the coarse mean (single winner)
            const int a = po.first_idx % ph.nA;
            fph.cx = po.avg_x; fph.cy = po.avg_y; fph.ch = pl.trig[(size_t)pl.spec_h[pid] + a];
            if (po.n_ties == 1 && dbits(fph.ch) == dbits(heading)) {
              s.n_passes++;
              s.n_ties = fo.n_ties;
              const double fheading = atan2(fo.ty, fo.tx);
              finalize_angular(fph, fo, fheading, angs + fph.ang_off, s.P, s.cov);
              s.mean[0] = fo.avg_x; s.mean[1] = fo.avg_y; s.mean[2] = fheading;
              double fbest = fo.best;
              if (fbest > 1.0) fbest = 1.0;
              s.best = fbest;
              s.stage = 5;
              h->work[10]++;
            }
          }
        }
      }
      iter++;
    }

    tr.mark("host finalize");
    // ---- results + clear ------------------------------------------------------------------------
    for (int i = 0; i < nw; i++) {
      const MatchState& s = states[i];
      ysm_result& r = out[s.idx];
      r.response = s.best;
      r.x = s.mean[0]; r.y = s.mean[1]; r.heading = s.mean[2];
      memcpy(r.cov, s.cov, sizeof(s.cov));
      r.n_passes = s.n_passes;
      r.n_ties = s.n_ties;
      r.status = s.status;
      r._pad = 0;
      r._reserved = 0.0;
    }
    if (built && !h->static_grid) {
      if ((h->debug & YSM_DEBUG_KEEP_GRIDS) && w1 >= b->n_matches) {
        h->grids_dirty = true;  // cleared at the start of the next call (the work list stays resident)
      } else {
        clear_wave(h, d_matches, st);
        kt.mark("k_tile_clear");
      }
    }
    kt.dump();
    CK(cudaGetLastError());
  }
  if (timing) {
    CK(cudaEventRecord(h->ev[7], st));
    CK(cudaEventSynchronize(h->ev[7]));
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev[6], h->ev[7]);
    h->t_total = ms;
    unsigned long long iss = 0;
    CK(cudaMemcpy(&iss, h->d_issued, 8, cudaMemcpyDeviceToHost));
    h->work[9] = (int64_t)iss;
  }
  for (int i = 0; i < b->n_matches; i++)
    if (out[i].status != YSM_OK) return fail(h, YSM_EMATCH, "Mapper FATAL ERROR - Unable to find best position");
  return YSM_OK;
}

// --------------------------------------------------------------------------------------------
static int match_batch_lanes(ysm_handle* h, const ysm_batch* b, ysm_result* out, void* stream);

extern "C" int ysm_match_batch(ysm_handle* h, const ysm_batch* b, ysm_result* out, void* stream) {
  if (!h || !b || !out) return YSM_EINVAL;
  try {  // no exception crosses the C ABI
    if (b->n_matches == 1 && stream == nullptr) {
      // single query: the resident latency kernel (falls through when not eligible / not finished there)
      PhaseTrace tr;
      for (int i = 0; i < 16; i++) h->work[i] = 0;
      const int rc = res_match(h, b, out, tr);
      if (rc != 1) return rc;
    }
    ysm_quiesce_device(h->device);
    return match_batch_lanes(h, b, out, stream);
  } catch (const std::bad_alloc&) {
    return fail(h, YSM_ENOMEM, "host allocation failed");
  } catch (const std::exception& e) {
    return fail(h, YSM_ECUDA, std::string("internal error: ") + e.what());
  }
}

static int match_batch_lanes(ysm_handle* h, const ysm_batch* b, ysm_result* out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int nl = 1 + (int)h->lanes.size();
  // (kernel timing / grid introspection are per-stream: those debug modes stay on the main lane)
  if (nl == 1 || b->n_matches < 64 * nl || (h->debug & (YSM_DEBUG_KEEP_GRIDS | YSM_DEBUG_TIME_KERNELS)))
    return match_batch_impl(h, b, out, st);
  if (b->n_matches < 0 || b->n_scans < 0 || b->n_points < 0) return fail(h, YSM_EINVAL, "negative sizes");
  // lanes: the point pool is made resident once, then every lane matches a contiguous share of the
  // batch on its own stream and host thread
  CK(cudaSetDevice(h->device));
  ysm_batch sb = *b;
  int64_t h2d = 0;
  std::vector<ysm_handle*> hs;
  hs.push_back(h);
  for (ysm_handle* sub : h->lanes) hs.push_back(sub);
  for (ysm_handle* x : hs) x->slice_wait.clear();
  bool lane_event_recorded = false;
  // Wave-sliced pool upload (host-resident pool). The pool is sent wave by wave (lane 0 wave 0, lane 1
  // wave 0, lane 0 wave 1, ...): only the scans a wave references and no earlier wave brought, as merged
  // contiguous ranges, with an event behind each wave -- so the first waves start after a fraction of the
  // pool and the rest of the copy overlaps their kernels. The ranges are worked out by the uploader
  // thread itself (below), off the lanes' critical path.
  struct UploadSlice { int lane, r0, r1; cudaEvent_t ev; };
  std::vector<UploadSlice> plan;
  if (!b->pool_on_device) {
    CK(h->d_pool_shared.ensure(std::max<size_t>(16, (size_t)b->n_points * 16)));
    sb.pool_xy = (const double*)h->d_pool_shared.p;
    sb.pool_on_device = 1;
    static const bool no_sliced = getenv("YSM_NO_SLICED_UPLOAD") != nullptr;
    const bool sliced = b->n_points > 0 && b->n_scans > 0 && !no_sliced;
    if (sliced) {
      CK(cudaEventRecord(h->lane_event, st));  // lanes and uploader start behind the caller's earlier work
      lane_event_recorded = true;
      if (!h->upload_stream) CK(cudaStreamCreateWithFlags(&h->upload_stream, cudaStreamNonBlocking));
      h->upload_seq.store(0, std::memory_order_release);
      int max_waves = 0;
      for (int l = 0; l < nl; l++) {
        const int lo = (int)((long long)b->n_matches * l / nl), hi = (int)((long long)b->n_matches * (l + 1) / nl);
        max_waves = std::max(max_waves, (hi - lo + hs[l]->slots - 1) / hs[l]->slots);  // (balancing keeps the count)
      }
      for (int w = 0; w < max_waves; w++) {
        for (int l = 0; l < nl; l++) {
          const int lo = (int)((long long)b->n_matches * l / nl), hi = (int)((long long)b->n_matches * (l + 1) / nl);
          const int S = balanced_wave(hi - lo, hs[l]->slots), r0 = w * S, r1 = std::min(hi - lo, (w + 1) * S);
          if (r0 >= r1) continue;
          const size_t k = plan.size();
          if (k >= h->slice_events.size()) {
            cudaEvent_t ev;
            CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            h->slice_events.push_back(ev);
          }
          hs[l]->slice_wait.push_back({r0, r1, h->slice_events[k], false, (int)k, &h->upload_seq});
          plan.push_back({l, r0, r1, h->slice_events[k]});
        }
      }
    } else {
      if (b->n_points > 0)
        CK(cudaMemcpyAsync(h->d_pool_shared.p, b->pool_xy, (size_t)b->n_points * 16, cudaMemcpyHostToDevice, st));
      h2d = (int64_t)b->n_points * 16;
    }
  }
  if (!lane_event_recorded) CK(cudaEventRecord(h->lane_event, st));
  std::vector<int> rcs(nl, YSM_OK);
  std::vector<ysm_batch> subs(nl, sb);
  std::vector<std::thread> threads;
  for (int l = 0; l < nl; l++) {
    const int lo = (int)((long long)b->n_matches * l / nl), hi = (int)((long long)b->n_matches * (l + 1) / nl);
    ysm_batch& q = subs[l];
    q.n_matches = hi - lo;
    q.query_scan = b->query_scan + lo;
    q.query_pose = b->query_pose + 3 * (size_t)lo;
    q.base_ptr = b->base_ptr + lo;  // absolute offsets into the shared base_idx
    CK(cudaStreamWaitEvent(hs[l]->lane_stream, h->lane_event, 0));
  }
  auto run = [&](int l) {
    const int lo = (int)((long long)b->n_matches * l / nl);
    rcs[l] = match_batch_impl(hs[l], &subs[l], out + lo, hs[l]->lane_stream);
  };
  // The uploader sends the pool in 4 MB pieces on its own stream and waits for each piece before it
  // submits the next: the H2D copy engine is a FIFO, and the lanes' small descriptor copies must not
  // queue behind the whole pool. Behind every wave's scans it records that wave's event. Bad indices
  // (the lanes will report them) or a fragmented access pattern fall back to sending everything at once.
  cudaError_t up_err = cudaSuccess;
  int64_t h2d_up = 0;
  auto uploader = [&]() {
    cudaError_t e = cudaSetDevice(h->device);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(h->upload_stream, h->lane_event, 0);
    const char* pe = getenv("YSM_UPLOAD_PIECE_KB");
    const int64_t piece = std::max<int64_t>(4096, ((pe ? atoll(pe) : 4096) << 10) / 16);  // points (default 4 MB)
    bool valid = true;
    for (int sc = 0; valid && sc < b->n_scans; sc++)
      if (b->scan_count[sc] < 0 || b->scan_start[sc] < 0 || (int64_t)b->scan_start[sc] + b->scan_count[sc] > b->n_points)
        valid = false;
    std::vector<uint8_t> up((size_t)b->n_scans, 0);
    std::vector<int> news;
    std::vector<std::pair<int64_t, int64_t>> ranges;
    bool all_sent = false;
    for (size_t k = 0; k < plan.size(); k++) {
      const UploadSlice& sl = plan[k];
      const int lo = (int)((long long)b->n_matches * sl.lane / nl);
      news.clear();
      ranges.clear();
      for (int i = lo + sl.r0; valid && i < lo + sl.r1; i++) {
        if (b->query_scan[i] < 0 || b->query_scan[i] >= b->n_scans || b->base_ptr[i + 1] < b->base_ptr[i]) { valid = false; break; }
        const int q = b->query_scan[i];
        if (!up[q]) { up[q] = 1; if (b->scan_count[q] > 0) news.push_back(q); }
        for (int j = b->base_ptr[i]; j < b->base_ptr[i + 1]; j++) {
          const int sc = b->base_idx[j];
          if (sc < 0 || sc >= b->n_scans) { valid = false; break; }
          if (!up[sc]) { up[sc] = 1; if (b->scan_count[sc] > 0) news.push_back(sc); }
        }
      }
      if (valid && !all_sent) {
        std::sort(news.begin(), news.end(), [&](int a, int c) { return b->scan_start[a] < b->scan_start[c]; });
        // merge scans that are adjacent (or overlapping) in the pool; when that leaves many ranges -- a copy costs a
        // stream synchronise -- also those less than a quarter piece (1 MB) apart: a few scans sent twice cost less
        // than a wave that waits for the whole pool (r02zm: 321-match waves of the relocalisation batch made 65+
        // exact ranges and fell back to that)
        for (int64_t gap = 0;; gap = piece / 4) {
          ranges.clear();
          for (int sc : news) {
            const int64_t a = b->scan_start[sc], z = a + b->scan_count[sc];
            if (!ranges.empty() && a <= ranges.back().second + gap) ranges.back().second = std::max(ranges.back().second, z);
            else ranges.push_back({a, z});
          }
          if (ranges.size() <= 48 || gap > 0) break;
        }
      }
      if ((!valid || ranges.size() > 256) && !all_sent) {
        ranges.assign(1, {0, b->n_points});
        all_sent = true;
      }
      for (const auto& r : ranges)
        for (int64_t a = r.first; a < r.second && e == cudaSuccess; a += piece) {
          const int64_t n = std::min(piece, r.second - a);
          e = cudaMemcpyAsync((char*)h->d_pool_shared.p + a * 16, (const char*)b->pool_xy + a * 16, (size_t)n * 16,
                              cudaMemcpyHostToDevice, h->upload_stream);
          if (e == cudaSuccess) e = cudaStreamSynchronize(h->upload_stream);
          h2d_up += n * 16;
        }
      if (e == cudaSuccess) e = cudaEventRecord(sl.ev, h->upload_stream);
      // (on an error the lanes are released anyway; the call fails below)
      h->upload_seq.store((int)k + 1, std::memory_order_release);
    }
    up_err = e;
  };
  // (no exception may cross the C ABI: if a thread cannot be started, its work runs on this thread)
  auto spawn = [&](auto&& fn) -> bool {
    try {
      threads.emplace_back(fn);
      return true;
    } catch (const std::exception&) {
      return false;
    }
  };
  if (!plan.empty() && !spawn(uploader)) uploader();  // (inline: the whole pool is sent before any lane starts)
  std::vector<int> inline_lanes;
  for (int l = 1; l < nl; l++)
    if (!spawn([&run, l] { run(l); })) inline_lanes.push_back(l);
  run(0);
  for (int l : inline_lanes) run(l);
  for (std::thread& t : threads) t.join();
  if (up_err != cudaSuccess) return fail(h, YSM_ECUDA, std::string("pool upload: ") + cudaGetErrorString(up_err));
  for (ysm_handle* x : hs) x->slice_wait.clear();
  // merge the lanes' accounting into the main handle
  for (int l = 1; l < nl; l++) {
    for (int i = 0; i < 16; i++) h->work[i] += hs[l]->work[i];
    h->t_sweep += hs[l]->t_sweep; h->t_build += hs[l]->t_build; h->t_reduce += hs[l]->t_reduce;
    h->t_total = std::max(h->t_total, hs[l]->t_total);
  }
  h->work[6] += h2d + h2d_up;
  h->work[11] = nl;
  for (int l = 0; l < nl; l++)
    if (rcs[l] != YSM_OK) {
      if (l) h->err = hs[l]->err;
      return rcs[l];
    }
  return YSM_OK;
}

// --------------------------------------------------------------------------------------------
extern "C" int ysm_debug_copy_grid(ysm_handle* h, int32_t match, uint8_t* out_host) {
  if (!h || !out_host) return YSM_EINVAL;
  if (h->static_grid) {  // the resident map grid
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out_host, h->d_grids, (size_t)h->g.data_size, cudaMemcpyDeviceToHost));
    return YSM_OK;
  }
  if (match < h->last_wave_begin || match >= h->last_wave_end || !h->grids_dirty)
    return fail(h, YSM_EINVAL, "grid of that match is not resident (set YSM_DEBUG_KEEP_GRIDS; last wave only)");
  CK(cudaSetDevice(h->device));
  CK(cudaDeviceSynchronize());
  const int slot = h->last_slot_of_match[match];
  CK(cudaMemcpy(out_host, h->d_grids + (size_t)slot * h->g.grid_bytes, (size_t)h->g.data_size, cudaMemcpyDeviceToHost));
  return YSM_OK;
}

extern "C" int ysm_debug_copy_kernel(ysm_handle* h, uint8_t* out_host) {
  if (!h || !out_host) return YSM_EINVAL;
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpy(out_host, h->d_kernel, h->h_kernel.size(), cudaMemcpyDeviceToHost));
  return YSM_OK;
}

extern "C" int ysm_debug_copy_offsets(ysm_handle* h, int32_t match, int32_t* out_host, int32_t* n_angles,
                                      int32_t* n_points) {
  if (!h || !n_angles || !n_points) return YSM_EINVAL;
  if (match < 0 || match >= (int)h->last_coarse_table_off.size() || h->last_coarse_table_off[match] < 0)
    return fail(h, YSM_EINVAL, "no coarse table recorded for that match");
  *n_angles = h->last_coarse_nA[match];
  *n_points = h->last_coarse_P[match];
  if (!out_host) return YSM_OK;
  // valid only if the match needed a single pass iteration after the coarse one did not
  // overwrite the offsets buffer: callers use do_refine=0 and non-degenerate inputs
  CK(cudaSetDevice(h->device));
  CK(cudaDeviceSynchronize());
  const int Ppad = h->last_coarse_Ppad[match];
  std::vector<int32_t> tmp((size_t)(*n_angles) * Ppad);
  CK(cudaMemcpy(tmp.data(), (const int*)h->d_offsets.p + h->last_coarse_table_off[match], tmp.size() * 4,
                cudaMemcpyDeviceToHost));
  for (int a = 0; a < *n_angles; a++)
    memcpy(out_host + (size_t)a * (*n_points), tmp.data() + (size_t)a * Ppad, (size_t)(*n_points) * 4);
  return YSM_OK;
}

// --------------------------------------------------------------------------------------------
// run_raytracing_sweep (reference yag_slam/raytracing.py:90-92) for many start cells.
// The reference's caller does one sweep per centroid (yag_slam/splicing.py:90-94): no allocation on the call --
// a per-device workspace (map copy, trig table, starts, rays, pinned staging) grows on demand and is reused.
namespace {
struct RayWorkspace {
  std::mutex mu;
  DevBuf img, cs, starts, out;
  PinBuf h_in;  // cos/sin table + starts, staged for one async copy
};
RayWorkspace g_ray_ws[64];
}  // namespace

extern "C" int ysm_raytrace(const uint8_t* img, int32_t hh, int32_t ww, int32_t img_on_device,
                            const double* angles_deg, int32_t n_angles, const double* starts_xy,
                            int32_t n_starts, float* out, int device, void* stream) {
  if (!img || !angles_deg || !starts_xy || !out || hh < 3 || ww < 3 || n_angles < 0 || n_starts < 0)
    return fail(nullptr, YSM_EINVAL, "ysm_raytrace: bad argument");
  if (n_angles == 0 || n_starts == 0) return YSM_OK;
  if (device < 0 || device >= 64) return fail(nullptr, YSM_EINVAL, "ysm_raytrace: bad device index");
  const long long n_rays = (long long)n_angles * n_starts;
  if (n_rays > 0x7fffffffLL / 5) return fail(nullptr, YSM_EUNSUP, "ysm_raytrace: too many rays in one call");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(nullptr, YSM_ECUDA, cudaGetErrorString(e));
  RayWorkspace& W = g_ray_ws[device];
  std::lock_guard<std::mutex> lk(W.mu);
  const size_t in_bytes = ((size_t)2 * n_angles + (size_t)2 * n_starts) * 8;
  const bool grow = (!img_on_device && (size_t)hh * ww > W.img.cap) || (size_t)2 * n_angles * 8 > W.cs.cap ||
                    (size_t)n_starts * 16 > W.starts.cap || (size_t)n_rays * 20 > W.out.cap || in_bytes > W.h_in.cap;
  if (grow) ysm_quiesce_device(device);  // (allocations below must not wait for a resident latency kernel)
  do {
    if (!img_on_device && (e = W.img.ensure((size_t)hh * ww)) != cudaSuccess) break;
    if ((e = W.cs.ensure((size_t)2 * n_angles * 8)) != cudaSuccess) break;
    if ((e = W.starts.ensure((size_t)n_starts * 16)) != cudaSuccess) break;
    if ((e = W.out.ensure((size_t)n_rays * 20)) != cudaSuccess) break;
    if ((e = W.h_in.ensure(in_bytes)) != cudaSuccess) break;
    double* cs = (double*)W.h_in.p;
    for (int a = 0; a < n_angles; a++) {
      const double ang = angles_deg[a] * (3.141592653589793 / 180.0);  // np.deg2rad
      cs[2 * a] = cos(ang);
      cs[2 * a + 1] = sin(ang);
    }
    memcpy(cs + 2 * (size_t)n_angles, starts_xy, (size_t)n_starts * 16);
    if (!img_on_device &&
        (e = cudaMemcpyAsync(W.img.p, img, (size_t)hh * ww, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(W.cs.p, cs, (size_t)2 * n_angles * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(W.starts.p, cs + 2 * (size_t)n_angles, (size_t)n_starts * 16, cudaMemcpyHostToDevice, st)) !=
        cudaSuccess) break;
    k_raywalk<<<(unsigned)((n_rays + 127) / 128), 128, 0, st>>>(img_on_device ? img : (const uint8_t*)W.img.p, hh, ww,
                                                                (const double*)W.cs.p, n_angles, (const double*)W.starts.p,
                                                                (int)n_rays, (float*)W.out.p);
    if ((e = cudaGetLastError()) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(out, W.out.p, (size_t)n_rays * 5 * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
    e = cudaStreamSynchronize(st);
  } while (0);
  if (e != cudaSuccess) return fail(nullptr, YSM_ECUDA, std::string("ysm_raytrace: ") + cudaGetErrorString(e));
  return YSM_OK;
}
