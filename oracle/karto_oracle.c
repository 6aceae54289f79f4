/*
 * karto_oracle.c -- CPU restatement of Karto's correlative scan matcher.
 *
 * THIS FILE IS TEST INFRASTRUCTURE. It is the parity checker for the CUDA path
 * in yag-slam_b200/csrc and the `cpu_baseline` / `--impl reference` leg of
 * bench.py. Nothing in the product path may import, link or call it.
 *
 * PARITY UNPINNED for the matcher: the arithmetic of the reference path lives in
 * the third-party wheel `karto_scanmatcher==1.0.0` (reference setup.py:46,
 * imported at yag_slam/scan_matching.py:22), whose source is not in
 * /root/reference and is not installed here. The reference ships no golden
 * vectors for it (test.py:23-43 only prints). This file restates the published
 * algorithm of ros-perception/open_karto (src/Mapper.cpp ScanMatcher::*,
 * include/open_karto/{Mapper.h,Karto.h,Math.h}) as summarised in
 * SURVEY.md Appendix A, and is anchored on the reference's call sites:
 *   - yag_slam/scan_matching.py:40-42 (Wrapper.match_scan(query, bases, penalty, do_fine))
 *   - yag_slam/helpers.py:339-361     (parameter names and defaults)
 *   - test.py:27-38                   (constructor/argument order)
 * (The ray-walk oracle in raywalk_oracle.c IS pinned, against the reference's
 *  own numba code, see tests/golden/.)
 *
 * Build: gcc -O2 -ffp-contract=off (no -march=native, no -ffast-math) so every
 * double operation is a separately rounded IEEE-754 operation, as in an
 * x86-64 manylinux build of Karto.
 *
 * Section tags "A.n" refer to SURVEY.md Appendix A.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define KT_PI 3.14159265358979323846
#define KT_2PI 6.28318530717958647692
#define KT_PI_180 0.01745329251994329577
#define KT_TOLERANCE 1e-06
#define MAX_VARIANCE 500.0
#define DISTANCE_PENALTY_GAIN 0.2
#define ANGLE_PENALTY_GAIN 0.2
#define GRIDSTATES_OCCUPIED 100

typedef struct {
  double search_size;
  double resolution;
  double smear_deviation;
  double range_threshold;
  double coarse_search_angle_offset;
  double coarse_angle_resolution;
  double fine_search_angle_resolution;
  double distance_variance_penalty;
  double angle_variance_penalty;
  double minimum_angle_penalty;
  double minimum_distance_penalty;
  int use_response_expansion;
  int _pad;
} ko_params;

typedef struct {
  ko_params p;
  /* A.1 sizing */
  int side;        /* search-space side in cells */
  int margin;      /* point reading margin */
  int roi;         /* ROI width == height */
  int half_kernel; /* from the EFFECTIVE resolution (CalculateKernel) */
  int kernel_size;
  int border;      /* from the RAW resolution (CreateGrid) */
  int width, height, stride;
  int data_size;
  double scale;    /* 1/resolution */
  double res_eff;  /* 1/scale == CoordinateConverter::GetResolution() */
  uint8_t *grid;
  uint8_t *kernel;
  double grid_off_x, grid_off_y; /* CorrelationGrid converter offset */
  /* search space probabilities: Grid<double>(side, side) */
  int probs_stride;
  double *probs;
  double probs_off_x, probs_off_y;
  /* GridIndexLookup */
  int n_angles;
  int n_points;
  int lookup_cap;
  int32_t *lookup; /* [n_angles][n_points] */
  /* debug / introspection */
  int last_num_ties;
  int last_num_passes;
  long probs_collisions;
} ko_matcher;

/* ---- A.0 primitives (open_karto Math.h) ---------------------------------- */
static inline double kt_round(double v) { return v >= 0.0 ? floor(v + 0.5) : ceil(v - 0.5); }
static inline int kt_double_equal(double a, double b) {
  double d = a - b;
  return d < 0.0 ? d >= -KT_TOLERANCE : d <= KT_TOLERANCE;
}
static inline int kt_in_range(double v, double lo, double hi) { return v >= lo && v <= hi; }
static inline int kt_is_up_to_i(int v, int maximum) { return v >= 0 && v < maximum; }
static inline double kt_square(double v) { return v * v; }
static inline double kt_max(double a, double b) { return a >= b ? a : b; }
static inline int kt_align8(int v) { return (v + 7) & ~7; }

static double kt_normalize_angle(double angle) {
  while (angle < -KT_PI) {
    if (angle < -KT_2PI) {
      angle += (uint32_t)(angle / -KT_2PI) * KT_2PI;
    } else {
      angle += KT_2PI;
    }
  }
  while (angle > KT_PI) {
    if (angle > KT_2PI) {
      angle -= (uint32_t)(angle / KT_2PI) * KT_2PI;
    } else {
      angle -= KT_2PI;
    }
  }
  return angle;
}

static double kt_normalize_angle_difference(double minuend, double subtrahend) {
  while (minuend - subtrahend < -KT_PI) minuend += KT_2PI;
  while (minuend - subtrahend > KT_PI) minuend -= KT_2PI;
  return minuend;
}

/* CoordinateConverter::WorldToGrid, one axis */
static inline int world_to_grid1(double w, double off, double scale) {
  return (int)kt_round((w - off) * scale);
}

/* ---- A.4 point readings (LocalizedRangeScan::Update) ----------------------
 * Python twin: yag_slam/helpers.py:58-68 via yag_slam/models.py:100-102.
 * Keeps beams with min_range <= r <= range_threshold. Returns count. */
int ko_point_readings(const double *ranges, int n, double min_angle, double angular_resolution,
                      double min_range, double range_threshold, double px, double py,
                      double heading, double *out_xy) {
  int k = 0;
  for (int i = 0; i < n; i++) {
    double r = ranges[i];
    if (!kt_in_range(r, min_range, range_threshold)) continue;
    double angle = heading + min_angle + (double)(uint32_t)i * angular_resolution;
    out_xy[2 * k] = px + (r * cos(angle));
    out_xy[2 * k + 1] = py + (r * sin(angle));
    k++;
  }
  return k;
}

/* ---- A.2 kernel (CorrelationGrid::CalculateKernel) ------------------------
 * Python twin: yag_slam/helpers.py:86-97. */
static int calculate_kernel(ko_matcher *m) {
  double resolution = m->res_eff;
  double smear = m->p.smear_deviation;
  if (!kt_in_range(smear, 0.5 * resolution, 10 * resolution)) return -1;
  m->half_kernel = (int)kt_round(2.0 * smear / resolution);
  m->kernel_size = 2 * m->half_kernel + 1;
  m->kernel = (uint8_t *)malloc((size_t)m->kernel_size * m->kernel_size);
  int half = m->kernel_size / 2;
  for (int i = -half; i <= half; i++) {
    for (int j = -half; j <= half; j++) {
      double d = hypot(i * resolution, j * resolution);
      double z = exp(-0.5 * pow(d / smear, 2));
      uint32_t kv = (uint32_t)kt_round(z * GRIDSTATES_OCCUPIED);
      m->kernel[(i + half) + m->kernel_size * (j + half)] = (uint8_t)kv;
    }
  }
  return 0;
}

/* ---- A.1 ScanMatcher::Create ---------------------------------------------- */
static ko_matcher *create_with_roi(const ko_params *p, int roi_override) {
  if (p->resolution <= 0 || p->search_size <= 0 || p->smear_deviation < 0 || p->range_threshold <= 0)
    return NULL;
  ko_matcher *m = (ko_matcher *)calloc(1, sizeof(ko_matcher));
  m->p = *p;
  m->side = (int)(uint32_t)(kt_round(p->search_size / p->resolution) + 1);
  m->margin = (int)(uint32_t)ceil(p->range_threshold / p->resolution);
  m->roi = roi_override > 0 ? roi_override : m->side + 2 * m->margin;
  /* CorrelationGrid::CreateGrid: border from the raw resolution */
  m->border = (int)kt_round(2.0 * p->smear_deviation / p->resolution) + 1;
  m->width = m->roi + 2 * m->border;
  m->height = m->roi + 2 * m->border;
  m->stride = kt_align8(m->width);
  m->data_size = m->stride * m->height;
  m->scale = 1.0 / p->resolution;
  m->res_eff = 1.0 / m->scale;
  if (calculate_kernel(m) != 0) {
    free(m);
    return NULL;
  }
  m->grid = (uint8_t *)malloc((size_t)m->data_size);
  memset(m->grid, 0, (size_t)m->data_size);
  m->probs_stride = kt_align8(m->side);
  m->probs = (double *)calloc((size_t)m->probs_stride * m->side, sizeof(double));
  m->lookup = NULL;
  m->lookup_cap = 0;
  return m;
}

ko_matcher *ko_create(const ko_params *p) { return create_with_roi(p, 0); }

void ko_destroy(ko_matcher *m) {
  if (!m) return;
  free(m->grid);
  free(m->kernel);
  free(m->probs);
  free(m->lookup);
  free(m);
}

/* introspection for parity tests: [side, margin, roi, half_kernel, kernel_size,
 * border, width, height, stride, data_size, n_angles, n_points, last_ties, last_passes] */
void ko_get_dims(const ko_matcher *m, int *out) {
  out[0] = m->side; out[1] = m->margin; out[2] = m->roi; out[3] = m->half_kernel;
  out[4] = m->kernel_size; out[5] = m->border; out[6] = m->width; out[7] = m->height;
  out[8] = m->stride; out[9] = m->data_size; out[10] = m->n_angles; out[11] = m->n_points;
  out[12] = m->last_num_ties; out[13] = m->last_num_passes;
}
const uint8_t *ko_grid_ptr(const ko_matcher *m) { return m->grid; }
const uint8_t *ko_kernel_ptr(const ko_matcher *m) { return m->kernel; }
const int32_t *ko_lookup_ptr(const ko_matcher *m) { return m->lookup; }
long ko_probs_collisions(const ko_matcher *m) { return m->probs_collisions; }

/* ---- A.3 FindValidPoints ---------------------------------------------------
 * Python analogue (different constants): yag_slam/helpers.py:298-329.
 * Writes a 0/1 mask over the scan's point readings; returns #valid. */
int ko_find_valid_points(const double *pts, int n, double vpx, double vpy, uint8_t *mask) {
  const double min_sq = kt_square(0.1);
  int trailing = 0;
  int count = 0;
  double fpx = 0.0, fpy = 0.0;
  int first_time = 1;
  memset(mask, 0, (size_t)n);
  for (int i = 0; i < n; i++) {
    double cx = pts[2 * i], cy = pts[2 * i + 1];
    if (first_time && !isnan(cx) && !isnan(cy)) {
      fpx = cx; fpy = cy;
      first_time = 0;
    }
    double dx = fpx - cx, dy = fpy - cy;
    if (kt_square(dx) + kt_square(dy) > min_sq) {
      double a = vpy - fpy;
      double b = fpx - vpx;
      double c = fpy * vpx - fpx * vpy;
      double ss = cx * a + cy * b + c;
      fpx = cx; fpy = cy;
      if (ss < 0.0) {
        trailing = i;
      } else {
        for (; trailing != i; ++trailing) {
          mask[trailing] = 1;
          count++;
        }
      }
    }
  }
  return count;
}

/* CorrelationGrid::SmearPoint (g is ROI-relative) */
static void smear_point(ko_matcher *m, int gx, int gy) {
  int half = m->kernel_size / 2;
  for (int j = -half; j <= half; j++) {
    uint8_t *row = m->grid + (gx + m->border) + (gy + j + m->border) * m->stride;
    int kc = half + m->kernel_size * (j + half);
    for (int i = -half; i <= half; i++) {
      uint8_t kv = m->kernel[i + kc];
      if (kv > row[i]) row[i] = kv;
    }
  }
}

/* ScanMatcher::AddScan */
static void add_scan(ko_matcher *m, const double *pts, int n, double vpx, double vpy, uint8_t *mask) {
  ko_find_valid_points(pts, n, vpx, vpy, mask);
  for (int i = 0; i < n; i++) {
    if (!mask[i]) continue;
    int gx = world_to_grid1(pts[2 * i], m->grid_off_x, m->scale);
    int gy = world_to_grid1(pts[2 * i + 1], m->grid_off_y, m->scale);
    if (!kt_is_up_to_i(gx, m->roi) || !kt_is_up_to_i(gy, m->roi)) continue;
    int gi = (gx + m->border) + (gy + m->border) * m->stride;
    if (m->grid[gi] == GRIDSTATES_OCCUPIED) continue;
    m->grid[gi] = GRIDSTATES_OCCUPIED;
    smear_point(m, gx, gy);
  }
}

/* ---- A.6 GridIndexLookup::ComputeOffsets ---------------------------------- */
static void compute_offsets(ko_matcher *m, const double *qpts, int nq, const double *pose,
                            double angle_center, double angle_offset, double angle_res) {
  int n_angles = (int)(uint32_t)(kt_round(angle_offset * 2.0 / angle_res) + 1);
  m->n_angles = n_angles;
  m->n_points = nq;
  size_t need = (size_t)n_angles * (size_t)(nq > 0 ? nq : 1);
  if ((size_t)m->lookup_cap < need) {
    free(m->lookup);
    m->lookup = (int32_t *)malloc(need * sizeof(int32_t));
    m->lookup_cap = (int)need;
  }
  /* Transform(sensorPose): SetTransform(Pose2(), pose) */
  double r00, r01, r10, r11; /* m_InverseRotation */
  if (pose[0] == 0.0 && pose[1] == 0.0 && pose[2] == 0.0) {
    r00 = 1.0; r01 = 0.0; r10 = 0.0; r11 = 1.0;
  } else {
    double radians = 0.0 - pose[2];
    double c = cos(radians), s = sin(radians);
    double omc = 1.0 - c;
    /* FromAxisAngle(0,0,1,radians) */
    r00 = 0.0 * omc + c;
    r01 = (0.0 * 0.0 * omc) - (1.0 * s);
    r10 = (0.0 * 0.0 * omc) + (1.0 * s);
    r11 = 0.0 * omc + c;
  }
  double *lx = (double *)malloc(sizeof(double) * (size_t)(nq > 0 ? nq : 1));
  double *ly = (double *)malloc(sizeof(double) * (size_t)(nq > 0 ? nq : 1));
  double dh = 0.0 - pose[2];
  for (int i = 0; i < nq; i++) {
    double dx = qpts[2 * i] - pose[0];
    double dy = qpts[2 * i + 1] - pose[1];
    /* Matrix3 * Pose2, third column of a z-rotation is (+0, +0) */
    lx[i] = r00 * dx + r01 * dy + 0.0 * dh;
    ly[i] = r10 * dx + r11 * dy + 0.0 * dh;
  }
  double start_angle = angle_center - angle_offset;
  for (int a = 0; a < n_angles; a++) {
    double angle = start_angle + (double)(uint32_t)a * angle_res;
    double cosine = cos(angle), sine = sin(angle);
    int32_t *out = m->lookup + (size_t)a * nq;
    for (int i = 0; i < nq; i++) {
      double ox = cosine * lx[i] - sine * ly[i];
      double oy = sine * lx[i] + cosine * ly[i];
      /* WorldToGrid(offset + gridOffset): keep the add-then-subtract */
      int gx = world_to_grid1(ox + m->grid_off_x, m->grid_off_x, m->scale);
      int gy = world_to_grid1(oy + m->grid_off_y, m->grid_off_y, m->scale);
      out[i] = gx + gy * m->stride; /* base Grid::GridIndex, no ROI shift, no bounds check */
    }
  }
  free(lx);
  free(ly);
}

/* ---- A.8 GetResponse ------------------------------------------------------- */
static double get_response(const ko_matcher *m, int angle_index, int grid_position_index) {
  double response = 0.0;
  const uint8_t *byte = m->grid + grid_position_index;
  int n = m->n_points;
  if (n == 0) return response;
  const int32_t *off = m->lookup + (size_t)angle_index * n;
  for (int i = 0; i < n; i++) {
    int idx = grid_position_index + off[i];
    if (!kt_is_up_to_i(idx, m->data_size)) continue;
    response += byte[off[i]];
  }
  response /= (double)((uint32_t)n * GRIDSTATES_OCCUPIED);
  return response;
}

typedef struct {
  double response;
  double x, y, heading;
} pose_response;

/* A.9 positional covariance */
static void positional_covariance(ko_matcher *m, const double *best_pose, double best_response,
                                  const double *center, double off_x, double off_y, double res_x,
                                  double res_y, double angle_res, double *cov) {
  memset(cov, 0, 9 * sizeof(double));
  cov[0] = cov[4] = cov[8] = 1.0;
  if (best_response < KT_TOLERANCE) {
    cov[0] = MAX_VARIANCE;
    cov[4] = MAX_VARIANCE;
    cov[8] = 4 * kt_square(angle_res);
    return;
  }
  double axx = 0, axy = 0, ayy = 0, norm = 0;
  double dx = best_pose[0] - center[0];
  double dy = best_pose[1] - center[1];
  uint32_t nx = (uint32_t)(kt_round(off_x * 2.0 / res_x) + 1);
  uint32_t ny = (uint32_t)(kt_round(off_y * 2.0 / res_y) + 1);
  double start_x = -off_x, start_y = -off_y;
  for (uint32_t yi = 0; yi < ny; yi++) {
    double y = start_y + yi * res_y;
    for (uint32_t xi = 0; xi < nx; xi++) {
      double x = start_x + xi * res_x;
      int gx = world_to_grid1(center[0] + x, m->probs_off_x, m->scale);
      int gy = world_to_grid1(center[1] + y, m->probs_off_y, m->scale);
      if (!kt_is_up_to_i(gx, m->side) || !kt_is_up_to_i(gy, m->side)) continue; /* Karto would throw */
      double response = m->probs[gx + gy * m->probs_stride];
      if (response >= (best_response - 0.1)) {
        norm += response;
        axx += (kt_square(x - dx) * response);
        axy += ((x - dx) * (y - dy) * response);
        ayy += (kt_square(y - dy) * response);
      }
    }
  }
  if (norm > KT_TOLERANCE) {
    double vxx = axx / norm, vxy = axy / norm, vyy = ayy / norm;
    double vthth = 4 * kt_square(angle_res);
    double min_vxx = 0.1 * kt_square(res_x);
    double min_vyy = 0.1 * kt_square(res_y);
    vxx = kt_max(vxx, min_vxx);
    vyy = kt_max(vyy, min_vyy);
    double mult = 1.0 / best_response;
    cov[0] = vxx * mult;
    cov[1] = vxy * mult;
    cov[3] = vxy * mult;
    cov[4] = vyy * mult;
    cov[8] = vthth;
  }
  if (kt_double_equal(cov[0], 0.0)) cov[0] = MAX_VARIANCE;
  if (kt_double_equal(cov[4], 0.0)) cov[4] = MAX_VARIANCE;
}

/* A.9 angular covariance */
static void angular_covariance(ko_matcher *m, const double *best_pose, double best_response,
                               const double *center, double angle_offset, double angle_res,
                               double *cov) {
  double best_angle = kt_normalize_angle_difference(best_pose[2], center[2]);
  int gx = world_to_grid1(best_pose[0], m->grid_off_x, m->scale);
  int gy = world_to_grid1(best_pose[1], m->grid_off_y, m->scale);
  int gi = (gx + m->border) + (gy + m->border) * m->stride;
  uint32_t n_angles = (uint32_t)(kt_round(angle_offset * 2 / angle_res) + 1);
  double start_angle = center[2] - angle_offset;
  double norm = 0.0, acc = 0.0;
  for (uint32_t a = 0; a < n_angles; a++) {
    double angle = start_angle + a * angle_res;
    double response = get_response(m, (int)a, gi);
    if (response >= (best_response - 0.1)) {
      norm += response;
      acc += (kt_square(angle - best_angle) * response);
    }
  }
  if (norm > KT_TOLERANCE) {
    if (acc < KT_TOLERANCE) acc = kt_square(angle_res);
    acc /= norm;
  } else {
    acc = 1000 * kt_square(angle_res);
  }
  cov[8] = acc;
}

/* ---- A.7 CorrelateScan ----------------------------------------------------- */
static double correlate_scan(ko_matcher *m, const double *qpts, int nq, const double *scan_pose,
                             const double *center_in, double off_x, double off_y, double res_x,
                             double res_y, double angle_offset, double angle_res, int do_penalize,
                             double *mean, double *cov, int fine) {
  double center[3] = {center_in[0], center_in[1], center_in[2]}; /* rSearchCenter may alias rMean */
  compute_offsets(m, qpts, nq, scan_pose, center[2], angle_offset, angle_res);
  if (!fine) {
    memset(m->probs, 0, sizeof(double) * (size_t)m->probs_stride * m->side);
    m->probs_off_x = center[0] - off_x;
    m->probs_off_y = center[1] - off_y;
  }
  uint32_t nx = (uint32_t)(kt_round(off_x * 2.0 / res_x) + 1);
  uint32_t ny = (uint32_t)(kt_round(off_y * 2.0 / res_y) + 1);
  double start_x = -off_x, start_y = -off_y;
  uint32_t n_angles = (uint32_t)(kt_round(angle_offset * 2.0 / angle_res) + 1);
  uint32_t total = nx * ny * n_angles;
  pose_response *pr = (pose_response *)malloc(sizeof(pose_response) * (size_t)total);
  uint32_t counter = 0;
  for (uint32_t yi = 0; yi < ny; yi++) {
    double y = start_y + yi * res_y;
    double new_y = center[1] + y;
    double square_y = kt_square(y);
    for (uint32_t xi = 0; xi < nx; xi++) {
      double x = start_x + xi * res_x;
      double new_x = center[0] + x;
      double square_x = kt_square(x);
      int gx = world_to_grid1(new_x, m->grid_off_x, m->scale);
      int gy = world_to_grid1(new_y, m->grid_off_y, m->scale);
      int gi = (gx + m->border) + (gy + m->border) * m->stride;
      double start_angle = center[2] - angle_offset;
      for (uint32_t a = 0; a < n_angles; a++) {
        double angle = start_angle + a * angle_res;
        double response = get_response(m, (int)a, gi);
        if (do_penalize && !kt_double_equal(response, 0.0)) {
          double sqd = square_x + square_y;
          double dp = 1.0 - (DISTANCE_PENALTY_GAIN * sqd / m->p.distance_variance_penalty);
          dp = kt_max(dp, m->p.minimum_distance_penalty);
          double sqa = kt_square(angle - center[2]);
          double ap = 1.0 - (ANGLE_PENALTY_GAIN * sqa / m->p.angle_variance_penalty);
          ap = kt_max(ap, m->p.minimum_angle_penalty);
          response *= (dp * ap);
        }
        pr[counter].response = response;
        pr[counter].x = new_x;
        pr[counter].y = new_y;
        pr[counter].heading = kt_normalize_angle(angle);
        counter++;
      }
    }
  }
  double best = -1;
  for (uint32_t i = 0; i < total; i++) {
    best = kt_max(best, pr[i].response);
    if (!fine) {
      int gx = world_to_grid1(pr[i].x, m->probs_off_x, m->scale);
      int gy = world_to_grid1(pr[i].y, m->probs_off_y, m->scale);
      if (kt_is_up_to_i(gx, m->side) && kt_is_up_to_i(gy, m->side)) {
        double *ptr = m->probs + gx + gy * m->probs_stride;
        *ptr = kt_max(pr[i].response, *ptr);
      }
    }
  }
  double sum_x = 0.0, sum_y = 0.0, theta_x = 0.0, theta_y = 0.0;
  int count = 0;
  for (uint32_t i = 0; i < total; i++) {
    if (kt_double_equal(pr[i].response, best)) {
      sum_x += pr[i].x;
      sum_y += pr[i].y;
      double h = pr[i].heading;
      theta_x += cos(h);
      theta_y += sin(h);
      count++;
    }
  }
  free(pr);
  m->last_num_ties = count;
  m->last_num_passes++;
  double avg[3];
  if (count > 0) {
    sum_x /= count;
    sum_y /= count;
    theta_x /= count;
    theta_y /= count;
    avg[0] = sum_x;
    avg[1] = sum_y;
    avg[2] = atan2(theta_y, theta_x);
  } else {
    return -2.0; /* "Unable to find best position" */
  }
  if (!fine) {
    positional_covariance(m, avg, best, center, off_x, off_y, res_x, res_y, angle_res, cov);
  } else {
    angular_covariance(m, avg, best, center, angle_offset, angle_res, cov);
  }
  mean[0] = avg[0];
  mean[1] = avg[1];
  mean[2] = avg[2];
  if (best > 1.0) best = 1.0;
  return best;
}

/* ---- A.5 ScanMatcher::MatchScan --------------------------------------------
 * query_pts: the query's filtered world point readings (nq xy pairs) at query_pose.
 * base_pts: concatenated filtered world point readings of the base scans.
 * out[13] = {response, x, y, heading, cov[0..8]} */
static int match_schedule(ko_matcher *m, const double *query_pts, int nq, const double *query_pose,
                          int do_penalize, int do_refine, double *out);

/* n_raw = number of RAW range readings of the query scan. Karto's early return tests that count
 * (pScan->GetNumberOfRangeReadings() == 0), not the number of filtered point readings: a scan with beams but
 * none inside [min_range, range_threshold] runs the schedule on an empty lookup table, GetResponse divides
 * 0.0 by (0 * 100), every response is NaN, nothing compares equal to the best response (-1) and CorrelateScan
 * throws "Unable to find best position" (rc != 0 here). */
int ko_match_raw(ko_matcher *m, const double *query_pts, int nq, int n_raw, const double *query_pose,
                 const double *base_pts, const int *base_counts, int nbase, int do_penalize,
                 int do_refine, double *out);

int ko_match(ko_matcher *m, const double *query_pts, int nq, const double *query_pose,
             const double *base_pts, const int *base_counts, int nbase, int do_penalize,
             int do_refine, double *out) {
  return ko_match_raw(m, query_pts, nq, nq, query_pose, base_pts, base_counts, nbase, do_penalize, do_refine, out);
}

int ko_match_raw(ko_matcher *m, const double *query_pts, int nq, int n_raw, const double *query_pose,
                 const double *base_pts, const int *base_counts, int nbase, int do_penalize,
                 int do_refine, double *out) {
  double cov[9];
  memset(cov, 0, sizeof(cov));
  cov[0] = cov[4] = cov[8] = 1.0; /* Matrix3 default-constructs... the wrapper hands in identity */
  m->last_num_passes = 0;
  if (n_raw == 0 || (nq == 0 && n_raw < 0)) {
    out[0] = 0.0;
    out[1] = query_pose[0]; out[2] = query_pose[1]; out[3] = query_pose[2];
    cov[0] = MAX_VARIANCE;
    cov[4] = MAX_VARIANCE;
    cov[8] = 4 * kt_square(m->p.coarse_angle_resolution);
    memcpy(out + 4, cov, sizeof(cov));
    return 0;
  }
  m->grid_off_x = query_pose[0] - (0.5 * (m->roi - 1) * m->res_eff);
  m->grid_off_y = query_pose[1] - (0.5 * (m->roi - 1) * m->res_eff);
  /* AddScans */
  memset(m->grid, 0, (size_t)m->data_size);
  {
    int maxn = 1;
    for (int b = 0; b < nbase; b++) if (base_counts[b] > maxn) maxn = base_counts[b];
    uint8_t *mask = (uint8_t *)malloc((size_t)maxn);
    const double *p = base_pts;
    for (int b = 0; b < nbase; b++) {
      add_scan(m, p, base_counts[b], query_pose[0], query_pose[1], mask);
      p += 2 * (size_t)base_counts[b];
    }
    free(mask);
  }
  return match_schedule(m, query_pts, nq, query_pose, do_penalize, do_refine, out);
}

/* coarse -> response expansion -> fine (A.5 steps 1-3) on the grid m currently holds */
static int match_schedule(ko_matcher *m, const double *query_pts, int nq, const double *query_pose,
                          int do_penalize, int do_refine, double *out) {
  double mean[3];
  double cov[9];
  memset(cov, 0, sizeof(cov));
  cov[0] = cov[4] = cov[8] = 1.0;
  double csx = 0.5 * (m->side - 1) * m->res_eff;
  double csy = 0.5 * (m->side - 1) * m->res_eff;
  double crx = 2 * m->res_eff, cry = 2 * m->res_eff;
  double best = correlate_scan(m, query_pts, nq, query_pose, query_pose, csx, csy, crx, cry,
                               m->p.coarse_search_angle_offset, m->p.coarse_angle_resolution,
                               do_penalize, mean, cov, 0);
  if (best < -1.5) return -1;
  if (m->p.use_response_expansion) {
    if (kt_double_equal(best, 0.0)) {
      double new_offset = m->p.coarse_search_angle_offset;
      for (int i = 0; i < 3; i++) {
        new_offset += 20 * KT_PI_180;
        best = correlate_scan(m, query_pts, nq, query_pose, query_pose, csx, csy, crx, cry,
                              new_offset, m->p.coarse_angle_resolution, do_penalize, mean, cov, 0);
        if (best < -1.5) return -1;
        if (!kt_double_equal(best, 0.0)) break;
      }
    }
  }
  if (do_refine) {
    double fx = crx * 0.5, fy = cry * 0.5;
    best = correlate_scan(m, query_pts, nq, query_pose, mean, fx, fy, m->res_eff, m->res_eff,
                          0.5 * m->p.coarse_angle_resolution, m->p.fine_search_angle_resolution,
                          do_penalize, mean, cov, 1);
    if (best < -1.5) return -1;
  }
  out[0] = best;
  out[1] = mean[0]; out[2] = mean[1]; out[3] = mean[2];
  memcpy(out + 4, cov, sizeof(cov));
  return 0;
}

/* ---- match against a map image --------------------------------------------------
 * The reference sketches this with its numba twin: occupancy_grid_map_to_correlation_grid
 * (yag_slam/helpers.py:24-34: every cell equal to occupied_value is set to full occupancy and smeared,
 * no skip rule) and Scan2DMatcherPy.match_scan_sets_with_map (yag_slam/scan_matching.py:124-173, which
 * calls an un-imported function and cannot run). Restated here with Karto's grid semantics: the
 * correlation grid's ROI is the map (cell (u, v) of the image = ROI cell (u, v), world offset (ox, oy) =
 * position of image cell (0, 0)), side = max(w, h) cells, usual border; occupied cells are 100 and
 * every one of them is smeared (pure max: order independent); MatchScan then runs its usual schedule
 * (A.5) on that fixed grid. Parity unpinned (no runnable reference). */
ko_matcher *ko_create_map(const ko_params *p, const uint8_t *img, int h, int w, int occupied_value,
                          double ox, double oy) {
  if (!img || h <= 0 || w <= 0) return NULL;
  ko_matcher *m = create_with_roi(p, h > w ? h : w);
  if (!m) return NULL;
  m->grid_off_x = ox;
  m->grid_off_y = oy;
  for (int v = 0; v < h; v++)
    for (int u = 0; u < w; u++)
      if (img[(size_t)v * w + u] == (uint8_t)occupied_value) {
        m->grid[(u + m->border) + (size_t)(v + m->border) * m->stride] = GRIDSTATES_OCCUPIED;
        smear_point(m, u, v);
      }
  return m;
}

int ko_match_map(ko_matcher *m, const double *query_pts, int nq, const double *query_pose,
                 int do_penalize, int do_refine, double *out) {
  m->last_num_passes = 0;
  if (nq == 0) {
    double cov[9];
    memset(cov, 0, sizeof(cov));
    out[0] = 0.0;
    out[1] = query_pose[0]; out[2] = query_pose[1]; out[3] = query_pose[2];
    cov[0] = MAX_VARIANCE;
    cov[4] = MAX_VARIANCE;
    cov[8] = 4 * kt_square(m->p.coarse_angle_resolution);
    memcpy(out + 4, cov, sizeof(cov));
    return 0;
  }
  return match_schedule(m, query_pts, nq, query_pose, do_penalize, do_refine, out);
}

/* Only the grid build (AddScans) -- for byte-exact grid parity tests. */
int ko_build_grid(ko_matcher *m, const double *query_pose, const double *base_pts,
                  const int *base_counts, int nbase) {
  m->grid_off_x = query_pose[0] - (0.5 * (m->roi - 1) * m->res_eff);
  m->grid_off_y = query_pose[1] - (0.5 * (m->roi - 1) * m->res_eff);
  memset(m->grid, 0, (size_t)m->data_size);
  int maxn = 1;
  for (int b = 0; b < nbase; b++) if (base_counts[b] > maxn) maxn = base_counts[b];
  uint8_t *mask = (uint8_t *)malloc((size_t)maxn);
  const double *p = base_pts;
  for (int b = 0; b < nbase; b++) {
    add_scan(m, p, base_counts[b], query_pose[0], query_pose[1], mask);
    p += 2 * (size_t)base_counts[b];
  }
  free(mask);
  return 0;
}

/* Only ComputeOffsets, after ko_build_grid/ko_match set the grid offset. */
int ko_compute_offsets(ko_matcher *m, const double *qpts, int nq, const double *pose,
                       double angle_center, double angle_offset, double angle_res) {
  compute_offsets(m, qpts, nq, pose, angle_center, angle_offset, angle_res);
  return m->n_angles;
}

/* ---- batch driver: one matcher per thread over independent matches ---------
 * The point pool holds every scan's filtered world readings; a match names its
 * query scan and base scans by pool index. Used by bench.py's cpu legs. */
int ko_match_batch(const ko_params *p, int n_matches, const double *pool_xy, const int *scan_start,
                   const int *scan_count, const int *query_scan, const double *query_poses,
                   const int *base_ptr, const int *base_idx, int do_penalize, int do_refine,
                   double *out /* [n_matches][13] */, int n_threads) {
  int err = 0;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel
#endif
  {
    ko_matcher *m = ko_create(p);
    double *bbuf = NULL;
    size_t bcap = 0;
    int *cnt = NULL;
    size_t ccap = 0;
    if (!m) {
      err = 1;
    } else {
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
      for (int i = 0; i < n_matches; i++) {
        int nb = base_ptr[i + 1] - base_ptr[i];
        size_t tot = 0;
        for (int b = 0; b < nb; b++) tot += (size_t)scan_count[base_idx[base_ptr[i] + b]];
        if (tot > bcap) { bcap = tot * 2 + 16; bbuf = (double *)realloc(bbuf, bcap * 2 * sizeof(double)); }
        if ((size_t)nb > ccap) { ccap = (size_t)nb * 2 + 4; cnt = (int *)realloc(cnt, ccap * sizeof(int)); }
        size_t w = 0;
        for (int b = 0; b < nb; b++) {
          int s = base_idx[base_ptr[i] + b];
          memcpy(bbuf + 2 * w, pool_xy + 2 * (size_t)scan_start[s], sizeof(double) * 2 * (size_t)scan_count[s]);
          cnt[b] = scan_count[s];
          w += (size_t)scan_count[s];
        }
        int q = query_scan[i];
        int rc = ko_match(m, pool_xy + 2 * (size_t)scan_start[q], scan_count[q], query_poses + 3 * (size_t)i,
                          bbuf, cnt, nb, do_penalize, do_refine, out + 13 * (size_t)i);
        if (rc != 0) err = 2;
      }
    }
    free(bbuf);
    free(cnt);
    ko_destroy(m);
  }
  return err;
}

int ko_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
