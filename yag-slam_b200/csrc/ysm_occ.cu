// ysm_occ.cu -- karto_scanmatcher.create_occupancy_grid on the B200.
//
// Replaces Karto's OccupancyGrid::CreateFromScans (the C++ ships in the
// wheel karto_scanmatcher==1.0.0, reference setup.py:46) behind the reference call sites
// yag_slam/graph_slam.py:341-342 and ros1/slam_node_ros1:188-209.
//
// Pipeline (one ysm_occ_create call; everything stays in HBM, the image is copied on request):
//   K6a k_occ_points      thread per beam: world point of every raw beam (device sincos) and the
//                         bounding box of the filtered points + sensor positions
//   K6b k_occ_candidates  beams whose point lies within 1e-9 m of a box face -> host re-evaluates
//                         those few with libm, so ComputeDimensions (width, height, offset) is
//                         bit-identical to the CPU reference
//   K6c k_occ_cells       thread per beam: ray shortening at range_threshold, WorldToGrid of the
//                         end point; an end point within 1e-6 cell of a rounding boundary is
//                         flagged and re-evaluated on the host with libm (exactness guard: device
//                         sincos may differ from libm in the last bits, the cell index may not)
//   K6d k_occ_trace       warp per ray, lanes = consecutive Bresenham steps in closed form
//                         (y_k = y0 + ystep * floor((2 k dy + dx) / (2 dx))), pass/hit counts by
//                         integer reductions (RED.ADD) -- order independent, hence exact
//   K6e k_occ_classify    UpdateCell: pass > 2 ? (hit/pass > 0.1 ? occupied : free) : unknown
//
// Compile with -fmad=false (see build.py): every double operation is separately rounded.
#include "ysm_internal.h"
#include "../../include/ysm.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

namespace {

constexpr double kBoxEps = 1e-9;    // metres: device-vs-libm slack for the bounding-box candidates
constexpr double kCellEps = 1e-6;   // cells: distance from a rounding boundary that triggers the libm re-evaluation
constexpr int kMaxList = 1 << 16;   // capacity of the candidate / fix-up lists

constexpr int F_SKIP = 1, F_ENDVALID = 2, F_UNSURE = 4;

struct OccScanDev {
  double px, py, heading;
  double min_angle, ares, min_range, max_range;
  int beam0, nbeams;
};

struct RayRec {  // what K6d consumes
  int x1, y1;    // end cell
  int scan;
  int flags;
};

thread_local std::string t_err;

__host__ __device__ inline double occ_round(double v) { return v >= 0.0 ? floor(v + 0.5) : ceil(v - 0.5); }

__device__ __forceinline__ long long ord_key(double v) {
  const long long k = __double_as_longlong(v);
  return k >= 0 ? k : k ^ 0x7fffffffffffffffLL;
}
inline double ord_val(long long k) {
  const long long b = k >= 0 ? k : k ^ 0x7fffffffffffffffLL;
  double v;
  memcpy(&v, &b, 8);
  return v;
}

__device__ __forceinline__ int scan_of_beam(const int* __restrict__ beam_ptr, int n_scans, int beam) {
  int lo = 0, hi = n_scans;  // last s with beam_ptr[s] <= beam
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(beam_ptr + mid) <= beam) lo = mid; else hi = mid;
  }
  return lo;
}

// K6a: LocalizedRangeScan::Update point readings of every raw beam + bounding box
// bbox[0..3] = ordered keys of min x, min y, max x, max y
__global__ void __launch_bounds__(256)
k_occ_points(const OccScanDev* __restrict__ scans, int n_scans, const int* __restrict__ beam_ptr,
             const double* __restrict__ ranges, int n_beams, double range_threshold,
             double2* __restrict__ pts, long long* __restrict__ bbox) {
  const int beam = blockIdx.x * blockDim.x + threadIdx.x;
  double mnx = 1e300, mny = 1e300, mxx = -1e300, mxy = -1e300;
  if (beam < n_beams) {
    const int s = scan_of_beam(beam_ptr, n_scans, beam);
    const OccScanDev sc = scans[s];
    const int i = beam - sc.beam0;
    const double r = ranges[beam];
    const double angle = sc.heading + sc.min_angle + (double)(unsigned)i * sc.ares;
    double sn, cs;
    sincos(angle, &sn, &cs);
    const double x = sc.px + (r * cs), y = sc.py + (r * sn);
    pts[beam] = make_double2(x, y);
    if (r >= sc.min_range && r <= range_threshold) {
      mnx = mxx = x;
      mny = mxy = y;
    }
    if (i == 0) {  // the sensor position belongs to the scan's box
      mnx = fmin(mnx, sc.px); mxx = fmax(mxx, sc.px);
      mny = fmin(mny, sc.py); mxy = fmax(mxy, sc.py);
    }
  }
  for (int d = 16; d > 0; d >>= 1) {
    mnx = fmin(mnx, __shfl_xor_sync(0xffffffffu, mnx, d));
    mny = fmin(mny, __shfl_xor_sync(0xffffffffu, mny, d));
    mxx = fmax(mxx, __shfl_xor_sync(0xffffffffu, mxx, d));
    mxy = fmax(mxy, __shfl_xor_sync(0xffffffffu, mxy, d));
  }
  __shared__ double s_r[4][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_r[0][warp] = mnx; s_r[1][warp] = mny; s_r[2][warp] = mxx; s_r[3][warp] = mxy; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double v = s_r[threadIdx.x][0];
    for (int k = 1; k < 8; k++) v = threadIdx.x < 2 ? fmin(v, s_r[threadIdx.x][k]) : fmax(v, s_r[threadIdx.x][k]);
    if (threadIdx.x < 2) { if (v < 1e299) atomicMin(bbox + threadIdx.x, ord_key(v)); }
    else if (v > -1e299) atomicMax(bbox + threadIdx.x, ord_key(v));
  }
}

// K6b: filtered beams whose device point is within kBoxEps of a face of the device box
__global__ void __launch_bounds__(256)
k_occ_candidates(const OccScanDev* __restrict__ scans, int n_scans, const int* __restrict__ beam_ptr,
                 const double* __restrict__ ranges, int n_beams, double range_threshold,
                 const double2* __restrict__ pts, double mnx, double mny, double mxx, double mxy,
                 int* __restrict__ list, int* __restrict__ count) {
  const int beam = blockIdx.x * blockDim.x + threadIdx.x;
  if (beam >= n_beams) return;
  const double2 p = pts[beam];
  if (!(p.x <= mnx + kBoxEps || p.x >= mxx - kBoxEps || p.y <= mny + kBoxEps || p.y >= mxy - kBoxEps)) return;
  const int s = scan_of_beam(beam_ptr, n_scans, beam);
  const double r = ranges[beam];
  if (!(r >= scans[s].min_range && r <= range_threshold)) return;
  const int k = atomicAdd(count, 1);
  if (k < kMaxList) list[k] = beam;
}

// K6c: OccupancyGrid::AddScan per-beam decisions + WorldToGrid of the (shortened) end point
__global__ void __launch_bounds__(256)
k_occ_cells(const OccScanDev* __restrict__ scans, int n_scans, const int* __restrict__ beam_ptr,
            const double* __restrict__ ranges, int n_beams, double range_threshold,
            const double2* __restrict__ pts, double offx, double offy, double scale,
            RayRec* __restrict__ rays, int* __restrict__ list, int* __restrict__ count) {
  const int beam = blockIdx.x * blockDim.x + threadIdx.x;
  if (beam >= n_beams) return;
  const int s = scan_of_beam(beam_ptr, n_scans, beam);
  const OccScanDev sc = scans[s];
  const double r = ranges[beam];
  RayRec rec;
  rec.scan = s;
  rec.x1 = rec.y1 = 0;
  rec.flags = 0;
  if (r <= sc.min_range || r >= sc.max_range || isnan(r)) {
    rec.flags = F_SKIP;
  } else {
    if (r < (range_threshold - 1e-6)) rec.flags |= F_ENDVALID;
    double2 p = pts[beam];
    if (r >= range_threshold) {
      const double ratio = range_threshold / r;
      const double dx = p.x - sc.px, dy = p.y - sc.py;
      p.x = sc.px + ratio * dx;
      p.y = sc.py + ratio * dy;
    }
    const double tx = (p.x - offx) * scale, ty = (p.y - offy) * scale;
    rec.x1 = (int)occ_round(tx);
    rec.y1 = (int)occ_round(ty);
    const double ux = tx + 0.5, uy = ty + 0.5;
    const double fx = ux - floor(ux), fy = uy - floor(uy);
    if (fx < kCellEps || fx > 1.0 - kCellEps || fy < kCellEps || fy > 1.0 - kCellEps) {
      rec.flags |= F_UNSURE;
      const int k = atomicAdd(count, 1);
      if (k < kMaxList) list[k] = beam;
    }
  }
  rays[beam] = rec;
}

// K6d: Grid<T>::TraceLine + OccupancyGrid::RayTrace. One warp per ray; lane l takes steps
// l, l+32, ... of the Bresenham walk, whose cell at step k is known in closed form:
//   major = m0 + k,  minor = n0 + nstep * floor((2 k dmin + dmaj) / (2 dmaj))
// (the error term of Karto's loop stays in [-dmaj/2, dmaj/2), ties decrement). The quotient is
// advanced incrementally by divmod(64 dmin, 2 dmaj), so a lane divides twice per ray.
__global__ void __launch_bounds__(256)
k_occ_trace(const OccScanDev* __restrict__ scans, const RayRec* __restrict__ rays, int n_beams,
            double offx, double offy, double scale, int w, int h, unsigned* __restrict__ pass,
            unsigned* __restrict__ hit, unsigned long long* __restrict__ visited) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  unsigned long long nvis = 0;
  for (int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; ray < n_beams; ray += warps) {
    const RayRec rc = rays[ray];
    if (rc.flags & F_SKIP) continue;
    const OccScanDev* sc = scans + rc.scan;
    int x0 = (int)occ_round((sc->px - offx) * scale), y0 = (int)occ_round((sc->py - offy) * scale);
    int x1 = rc.x1, y1 = rc.y1;
    if (lane == 0 && (rc.flags & F_ENDVALID) && x1 >= 0 && x1 < w && y1 >= 0 && y1 < h) {
      atomicAdd(pass + (size_t)y1 * w + x1, 1u);
      atomicAdd(hit + (size_t)y1 * w + x1, 1u);
    }
    const bool steep = abs(y1 - y0) > abs(x1 - x0);
    if (steep) { int t = x0; x0 = y0; y0 = t; t = x1; x1 = y1; y1 = t; }
    if (x0 > x1) { int t = x0; x0 = x1; x1 = t; t = y0; y0 = y1; y1 = t; }
    const int dmaj = x1 - x0, dmin = abs(y1 - y0);
    const int nstep = y0 < y1 ? 1 : -1;
    // bounds of the (major, minor) frame
    const int wmaj = steep ? h : w, wmin = steep ? w : h;
    const int smaj = steep ? w : 1, smin = steep ? 1 : w;  // address strides
    unsigned q = 0, rem = 0, q32 = 0, r32 = 0;
    const unsigned den = 2u * (unsigned)dmaj;
    if (dmaj > 0) {
      const unsigned num = 2u * (unsigned)lane * (unsigned)dmin + (unsigned)dmaj;
      q = num / den; rem = num - q * den;
      const unsigned n32 = 64u * (unsigned)dmin;
      q32 = n32 / den; r32 = n32 - q32 * den;
    }
    for (int k = lane; k <= dmaj; k += 32) {
      const int a = x0 + k, b = y0 + nstep * (int)q;
      if (a >= 0 && a < wmaj && b >= 0 && b < wmin) atomicAdd(pass + (size_t)a * smaj + (size_t)b * smin, 1u);
      q += q32; rem += r32;
      if (rem >= den) { rem -= den; q++; }
    }
    nvis += (unsigned)(dmaj + 1);
  }
  if (visited && lane == 0 && nvis) atomicAdd(visited, nvis);
}

// K6e: OccupancyGrid::Update / UpdateCell; image values as ros1/slam_node_ros1:199-202 decodes them
__global__ void __launch_bounds__(256)
k_occ_classify(const unsigned* __restrict__ pass, const unsigned* __restrict__ hit, long long n,
               uint8_t* __restrict__ image) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned p = pass[i];
  uint8_t v = 200;
  if (p > 2u) v = ((double)hit[i] / (double)p) > 0.1 ? 0 : 255;
  image[i] = v;
}

int occ_fail(int code, const std::string& msg) {
  t_err = msg;
  return code;
}

}  // namespace

struct ysm_occ {
  int device = 0;
  ysm_occ_info info{};
  unsigned* d_pass = nullptr;
  unsigned* d_hit = nullptr;
  uint8_t* d_image = nullptr;
};

extern "C" const char* ysm_occ_last_error(void) { return t_err.c_str(); }

extern "C" void ysm_occ_destroy(ysm_occ* o) {
  if (!o) return;
  cudaSetDevice(o->device);
  ysm_quiesce_device(o->device);
  if (o->d_pass) cudaFree(o->d_pass);
  if (o->d_hit) cudaFree(o->d_hit);
  if (o->d_image) cudaFree(o->d_image);
  delete o;
}

#define OCK(x)                                                                                   \
  do {                                                                                           \
    cudaError_t e_ = (x);                                                                        \
    if (e_ != cudaSuccess) {                                                                     \
      rc = occ_fail(YSM_ECUDA, std::string("ysm_occ_create: ") + #x + ": " + cudaGetErrorString(e_)); \
      goto done;                                                                                 \
    }                                                                                            \
  } while (0)

extern "C" int ysm_occ_create(const ysm_occ_scans* in, int device, void* stream, ysm_occ** out) {
  ysm_quiesce_device(device);  // (a resident latency kernel would make the allocations / syncs below wait)
  if (!in || !out) return occ_fail(YSM_EINVAL, "ysm_occ_create: null argument");
  *out = nullptr;
  if (in->n_scans <= 0 || !in->pose || !in->laser || !in->beam_ptr)
    return occ_fail(YSM_EINVAL, "ysm_occ_create: no scans (Karto's CreateFromScans returns NULL)");
  if (!(in->resolution > 0.0) || !(in->range_threshold > 0.0))
    return occ_fail(YSM_EINVAL, "ysm_occ_create: resolution and range_threshold must be positive");
  const int ns = in->n_scans;
  const long long nb_ll = in->beam_ptr[ns];
  if (nb_ll < 0 || nb_ll > 0x7fffff00LL) return occ_fail(YSM_EUNSUP, "ysm_occ_create: too many beams in one call");
  const int nb = (int)nb_ll;
  if (nb > 0 && !in->ranges) return occ_fail(YSM_EINVAL, "ysm_occ_create: null ranges");
  cudaStream_t st = (cudaStream_t)stream;
  {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return occ_fail(YSM_ECUDA, std::string("ysm_occ_create: ") + cudaGetErrorString(e));
  }
  const double thr = in->range_threshold, scale = 1.0 / in->resolution;
  std::vector<OccScanDev> hs((size_t)ns);
  for (int s = 0; s < ns; s++) {
    OccScanDev& d = hs[s];
    d.px = in->pose[3 * s]; d.py = in->pose[3 * s + 1]; d.heading = in->pose[3 * s + 2];
    d.min_angle = in->laser[4 * s]; d.ares = in->laser[4 * s + 1];
    d.min_range = in->laser[4 * s + 2]; d.max_range = in->laser[4 * s + 3];
    d.beam0 = in->beam_ptr[s]; d.nbeams = in->beam_ptr[s + 1] - in->beam_ptr[s];
    if (d.nbeams < 0) return occ_fail(YSM_EINVAL, "ysm_occ_create: beam_ptr must be non-decreasing");
  }
  // libm evaluation of one beam's point (the CPU reference's arithmetic, A.4)
  auto host_point = [&](int beam, double& x, double& y, int& s_out) {
    int s = (int)(std::upper_bound(in->beam_ptr, in->beam_ptr + ns + 1, beam) - in->beam_ptr) - 1;
    s = std::min(std::max(s, 0), ns - 1);
    const OccScanDev& d = hs[s];
    const double r = in->ranges[beam];
    const double angle = d.heading + d.min_angle + (double)(uint32_t)(beam - d.beam0) * d.ares;
    x = d.px + (r * cos(angle));
    y = d.py + (r * sin(angle));
    s_out = s;
  };

  int rc = YSM_OK;
  ysm_occ* o = new ysm_occ();
  o->device = device;
  OccScanDev* d_scans = nullptr;
  int *d_bptr = nullptr, *d_list = nullptr, *d_count = nullptr;
  double* d_ranges = nullptr;
  double2* d_pts = nullptr;
  long long* d_bbox = nullptr;
  RayRec* d_rays = nullptr;
  unsigned long long* d_vis = nullptr;
  const int nblk = std::max(1, (nb + 255) / 256);
  long long hb[4];
  double mnx, mny, mxx, mxy;
  int n_cand = 0, n_fix = 0, w = 0, h = 0;
  long long ncell = 0;
  unsigned long long visited = 0;
  std::vector<int> list;

  OCK(cudaMalloc((void**)&d_scans, sizeof(OccScanDev) * (size_t)ns));
  OCK(cudaMalloc((void**)&d_bptr, 4 * (size_t)(ns + 1)));
  OCK(cudaMalloc((void**)&d_ranges, 8 * (size_t)std::max(nb, 1)));
  OCK(cudaMalloc((void**)&d_pts, 16 * (size_t)std::max(nb, 1)));
  OCK(cudaMalloc((void**)&d_rays, sizeof(RayRec) * (size_t)std::max(nb, 1)));
  OCK(cudaMalloc((void**)&d_list, 4 * (size_t)kMaxList));
  OCK(cudaMalloc((void**)&d_count, 4));
  OCK(cudaMalloc((void**)&d_bbox, 32));
  OCK(cudaMalloc((void**)&d_vis, 8));
  OCK(cudaMemcpyAsync(d_scans, hs.data(), sizeof(OccScanDev) * (size_t)ns, cudaMemcpyHostToDevice, st));
  OCK(cudaMemcpyAsync(d_bptr, in->beam_ptr, 4 * (size_t)(ns + 1), cudaMemcpyHostToDevice, st));
  if (nb) OCK(cudaMemcpyAsync(d_ranges, in->ranges, 8 * (size_t)nb, cudaMemcpyHostToDevice, st));
  hb[0] = hb[1] = 0x7fffffffffffffffLL;
  hb[2] = hb[3] = (long long)0x8000000000000000ULL;
  OCK(cudaMemcpyAsync(d_bbox, hb, 32, cudaMemcpyHostToDevice, st));
  OCK(cudaMemsetAsync(d_count, 0, 4, st));
  OCK(cudaMemsetAsync(d_vis, 0, 8, st));
  k_occ_points<<<nblk, 256, 0, st>>>(d_scans, ns, d_bptr, d_ranges, nb, thr, d_pts, d_bbox);
  OCK(cudaGetLastError());
  OCK(cudaMemcpyAsync(hb, d_bbox, 32, cudaMemcpyDeviceToHost, st));
  OCK(cudaStreamSynchronize(st));
  o->info.launches++;
  // sensor positions are exact on the host (no trigonometry)
  mnx = mny = 999999999999999999.99999;  // BoundingBox2's default corners (Karto.h)
  mxx = mxy = -999999999999999999.99999;
  for (int s = 0; s < ns; s++) {
    mnx = std::min(mnx, hs[s].px); mxx = std::max(mxx, hs[s].px);
    mny = std::min(mny, hs[s].py); mxy = std::max(mxy, hs[s].py);
  }
  if (nb) {
    const double dmnx = std::min(ord_val(hb[0]), mnx), dmny = std::min(ord_val(hb[1]), mny);
    const double dmxx = std::max(ord_val(hb[2]), mxx), dmxy = std::max(ord_val(hb[3]), mxy);
    k_occ_candidates<<<nblk, 256, 0, st>>>(d_scans, ns, d_bptr, d_ranges, nb, thr, d_pts, dmnx, dmny, dmxx, dmxy,
                                           d_list, d_count);
    OCK(cudaGetLastError());
    OCK(cudaMemcpyAsync(&n_cand, d_count, 4, cudaMemcpyDeviceToHost, st));
    OCK(cudaStreamSynchronize(st));
    o->info.launches++;
    if (n_cand <= kMaxList) {
      list.resize((size_t)n_cand);
      if (n_cand) OCK(cudaMemcpy(list.data(), d_list, 4 * (size_t)n_cand, cudaMemcpyDeviceToHost));
      for (int beam : list) {
        double x, y; int s;
        host_point(beam, x, y, s);
        mnx = std::min(mnx, x); mxx = std::max(mxx, x);
        mny = std::min(mny, y); mxy = std::max(mxy, y);
      }
    } else {  // degenerate input (tens of thousands of points on a box face): evaluate every beam with libm
      for (int beam = 0; beam < nb; beam++) {
        double x, y; int s;
        host_point(beam, x, y, s);
        const double r = in->ranges[beam];
        if (!(r >= hs[s].min_range && r <= thr)) continue;
        mnx = std::min(mnx, x); mxx = std::max(mxx, x);
        mny = std::min(mny, y); mxy = std::max(mxy, y);
      }
    }
    o->info.box_candidates = n_cand;
  }
  // OccupancyGrid::ComputeDimensions
  w = (int)occ_round((mxx - mnx) * scale);
  h = (int)occ_round((mxy - mny) * scale);
  o->info.width = w; o->info.height = h;
  o->info.offset_x = mnx; o->info.offset_y = mny;
  o->info.resolution = in->resolution;
  o->info.rays = nb;
  ncell = (long long)w * h;
  if (w < 0 || h < 0 || ncell > (1ll << 33)) {
    rc = occ_fail(YSM_EUNSUP, "ysm_occ_create: occupancy grid too large");
    goto done;
  }
  if (ncell > 0) {
    OCK(cudaMalloc((void**)&o->d_pass, 4 * (size_t)ncell));
    OCK(cudaMalloc((void**)&o->d_hit, 4 * (size_t)ncell));
    OCK(cudaMalloc((void**)&o->d_image, (size_t)ncell));
    OCK(cudaMemsetAsync(o->d_pass, 0, 4 * (size_t)ncell, st));
    OCK(cudaMemsetAsync(o->d_hit, 0, 4 * (size_t)ncell, st));
  }
  if (nb && ncell > 0) {
    OCK(cudaMemsetAsync(d_count, 0, 4, st));
    k_occ_cells<<<nblk, 256, 0, st>>>(d_scans, ns, d_bptr, d_ranges, nb, thr, d_pts, mnx, mny, scale, d_rays, d_list,
                                      d_count);
    OCK(cudaGetLastError());
    OCK(cudaMemcpyAsync(&n_fix, d_count, 4, cudaMemcpyDeviceToHost, st));
    OCK(cudaStreamSynchronize(st));
    o->info.launches++;
    o->info.cell_fixups = n_fix;
    auto fix_one = [&](int beam) -> RayRec {  // the CPU reference's arithmetic for one beam (A.10)
      double x, y; int s;
      host_point(beam, x, y, s);
      const OccScanDev& d = hs[s];
      const double r = in->ranges[beam];
      RayRec rec; rec.scan = s; rec.x1 = rec.y1 = 0; rec.flags = 0;
      if (r <= d.min_range || r >= d.max_range || std::isnan(r)) { rec.flags = F_SKIP; return rec; }
      if (r < (thr - 1e-6)) rec.flags |= F_ENDVALID;
      if (r >= thr) {
        const double ratio = thr / r;
        const double dx = x - d.px, dy = y - d.py;
        x = d.px + ratio * dx;
        y = d.py + ratio * dy;
      }
      rec.x1 = (int)occ_round((x - mnx) * scale);
      rec.y1 = (int)occ_round((y - mny) * scale);
      return rec;
    };
    if (n_fix > 0 && n_fix <= kMaxList) {
      list.resize((size_t)n_fix);
      OCK(cudaMemcpy(list.data(), d_list, 4 * (size_t)n_fix, cudaMemcpyDeviceToHost));
      for (int beam : list) {
        const RayRec rec = fix_one(beam);
        OCK(cudaMemcpy(d_rays + beam, &rec, sizeof(RayRec), cudaMemcpyHostToDevice));
      }
    } else if (n_fix > kMaxList) {
      std::vector<RayRec> all((size_t)nb);
      for (int beam = 0; beam < nb; beam++) all[beam] = fix_one(beam);
      OCK(cudaMemcpy(d_rays, all.data(), sizeof(RayRec) * (size_t)nb, cudaMemcpyHostToDevice));
    }
    {
      int dev_sms = 148;
      cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, device);
      const int blocks = (int)std::min<long long>(((long long)nb + 7) / 8, (long long)dev_sms * 8 * 4);
      k_occ_trace<<<std::max(blocks, 1), 256, 0, st>>>(d_scans, d_rays, nb, mnx, mny, scale, w, h, o->d_pass, o->d_hit,
                                                       d_vis);
      OCK(cudaGetLastError());
      o->info.launches++;
    }
  }
  if (ncell > 0) {
    k_occ_classify<<<(unsigned)((ncell + 255) / 256), 256, 0, st>>>(o->d_pass, o->d_hit, ncell, o->d_image);
    OCK(cudaGetLastError());
    o->info.launches++;
  }
  OCK(cudaMemcpyAsync(&visited, d_vis, 8, cudaMemcpyDeviceToHost, st));
  OCK(cudaStreamSynchronize(st));
  o->info.cells_visited = (long long)visited;

done:
  if (d_scans) cudaFree(d_scans);
  if (d_bptr) cudaFree(d_bptr);
  if (d_ranges) cudaFree(d_ranges);
  if (d_pts) cudaFree(d_pts);
  if (d_rays) cudaFree(d_rays);
  if (d_list) cudaFree(d_list);
  if (d_count) cudaFree(d_count);
  if (d_bbox) cudaFree(d_bbox);
  if (d_vis) cudaFree(d_vis);
  if (rc != YSM_OK) {
    ysm_occ_destroy(o);
    return rc;
  }
  *out = o;
  return YSM_OK;
}

extern "C" int ysm_occ_get_info(const ysm_occ* o, ysm_occ_info* out) {
  if (!o || !out) return occ_fail(YSM_EINVAL, "ysm_occ_get_info: null argument");
  *out = o->info;
  return YSM_OK;
}

extern "C" int ysm_occ_copy_image(const ysm_occ* o, uint8_t* out_host) {
  if (!o || !out_host) return occ_fail(YSM_EINVAL, "ysm_occ_copy_image: null argument");
  const size_t n = (size_t)o->info.width * o->info.height;
  if (!n) return YSM_OK;
  cudaSetDevice(o->device);
  cudaError_t e = cudaMemcpy(out_host, o->d_image, n, cudaMemcpyDeviceToHost);
  return e == cudaSuccess ? YSM_OK : occ_fail(YSM_ECUDA, cudaGetErrorString(e));
}

extern "C" int ysm_occ_copy_counts(const ysm_occ* o, uint32_t* pass_host, uint32_t* hit_host) {
  if (!o || !pass_host || !hit_host) return occ_fail(YSM_EINVAL, "ysm_occ_copy_counts: null argument");
  const size_t n = (size_t)o->info.width * o->info.height;
  if (!n) return YSM_OK;
  cudaSetDevice(o->device);
  cudaError_t e = cudaMemcpy(pass_host, o->d_pass, 4 * n, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(hit_host, o->d_hit, 4 * n, cudaMemcpyDeviceToHost);
  return e == cudaSuccess ? YSM_OK : occ_fail(YSM_ECUDA, cudaGetErrorString(e));
}

extern "C" const uint8_t* ysm_occ_device_image(const ysm_occ* o) { return o ? o->d_image : nullptr; }
