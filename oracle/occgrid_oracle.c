/*
 * occgrid_oracle.c -- CPU restatement of karto_scanmatcher.create_occupancy_grid
 * (TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg as the checker; never by the product path).
 *
 * Reference call sites: yag_slam/graph_slam.py:341-342 (make_occupancy_grid),
 * ros1/slam_node_ros1:188-209 (_make_map: image values 0 occupied / 200 unknown / 255 free,
 * .width/.height/.offset), yag_slam/helpers.py:590-603.
 *
 * PARITY UNPINNED: the implementation lives in the third-party wheel karto_scanmatcher==1.0.0
 * (reference setup.py:46), absent from /root/reference and not installable here; the reference
 * holds no golden occupancy image. The algorithm restated is open_karto's
 * OccupancyGrid::CreateFromScans (Karto.h: ComputeDimensions, AddScan, RayTrace,
 * Grid<T>::TraceLine, UpdateCell/Update):
 *
 *   bounding box  = union over scans of {sensor position} + {filtered point readings
 *                   (min_range <= r <= range_threshold)}                (LocalizedRangeScan::Update)
 *   width,height  = Round(size * (1/res)), offset = box minimum         (ComputeDimensions)
 *   per scan, per raw beam r:                                           (AddScan)
 *       skip if r <= min_range || r >= max_range || isnan(r)
 *       endpoint valid iff r < range_threshold - 1e-6 (KT_TOLERANCE)
 *       if r >= range_threshold: endpoint = sensor + (range_threshold / r) * (point - sensor)
 *       Bresenham from WorldToGrid(sensor) to WorldToGrid(endpoint): pass++ on every in-bounds
 *       cell, end cell included                                         (Grid::TraceLine)
 *       valid endpoint in bounds: pass++ and hit++ once more            (RayTrace)
 *   cell = pass > 2 ? (hit/pass > 0.1 ? Occupied : Free) : Unknown      (UpdateCell)
 *
 * Decisions the absent wheel leaves open (stated in DESIGN.md):
 *   - create_occupancy_grid's range_threshold argument replaces the laser's range_threshold
 *     for the whole construction (bounding box filter, ray shortening, endpoint validity);
 *   - the image is [height][width] uint8 with Occupied -> 0, Unknown -> 200, Free -> 255
 *     (the values ros1/slam_node_ros1:199-202 decodes), row 0 = minimum y.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline double oc_round(double v) { return v >= 0.0 ? floor(v + 0.5) : ceil(v - 0.5); }

typedef struct oc_dims {
  int32_t width, height;
  double offset_x, offset_y;
} oc_dims;

/* scans: pose[n][3] sensor pose; laser[n][4] = min_angle, angular_resolution, min_range, max_range;
 * ranges concatenated, beam_ptr[n+1]. */
int oc_compute_dims(int n_scans, const double *pose, const double *laser, const double *ranges,
                    const int32_t *beam_ptr, double resolution, double range_threshold, oc_dims *out) {
  /* BoundingBox2 default corners (Karto.h) */
  double minx = 999999999999999999.99999, miny = 999999999999999999.99999;
  double maxx = -999999999999999999.99999, maxy = -999999999999999999.99999;
  if (n_scans <= 0) return -1;
  for (int s = 0; s < n_scans; s++) {
    const double px = pose[3 * s], py = pose[3 * s + 1], heading = pose[3 * s + 2];
    const double min_angle = laser[4 * s], ares = laser[4 * s + 1], min_range = laser[4 * s + 2];
    if (px < minx) minx = px;
    if (px > maxx) maxx = px;
    if (py < miny) miny = py;
    if (py > maxy) maxy = py;
    const int nb = beam_ptr[s + 1] - beam_ptr[s];
    const double *r = ranges + beam_ptr[s];
    for (int i = 0; i < nb; i++) {
      if (!(r[i] >= min_range && r[i] <= range_threshold)) continue;
      const double angle = heading + min_angle + (double)(uint32_t)i * ares;
      const double x = px + (r[i] * cos(angle));
      const double y = py + (r[i] * sin(angle));
      if (x < minx) minx = x;
      if (x > maxx) maxx = x;
      if (y < miny) miny = y;
      if (y > maxy) maxy = y;
    }
  }
  const double scale = 1.0 / resolution;
  out->width = (int32_t)oc_round((maxx - minx) * scale);
  out->height = (int32_t)oc_round((maxy - miny) * scale);
  out->offset_x = minx;
  out->offset_y = miny;
  return 0;
}

static void trace_line(uint32_t *pass, int w, int h, int x0, int y0, int x1, int y1) {
  const int steep = abs(y1 - y0) > abs(x1 - x0);
  int t;
  if (steep) {
    t = x0; x0 = y0; y0 = t;
    t = x1; x1 = y1; y1 = t;
  }
  if (x0 > x1) {
    t = x0; x0 = x1; x1 = t;
    t = y0; y0 = y1; y1 = t;
  }
  const int dx = x1 - x0, dy = abs(y1 - y0);
  int error = 0, y = y0;
  const int ystep = y0 < y1 ? 1 : -1;
  for (int x = x0; x <= x1; x++) {
    const int px = steep ? y : x, py = steep ? x : y;
    error += dy;
    if (2 * error >= dx) {
      y += ystep;
      error -= dx;
    }
    if (px >= 0 && px < w && py >= 0 && py < h) pass[(size_t)py * w + px]++;
  }
}

/* pass/hit: [height][width] uint32, zeroed by the caller; image: [height][width] uint8. */
int oc_render(int n_scans, const double *pose, const double *laser, const double *ranges,
              const int32_t *beam_ptr, double resolution, double range_threshold, const oc_dims *d,
              uint32_t *pass, uint32_t *hit, uint8_t *image) {
  const int w = d->width, h = d->height;
  const double scale = 1.0 / resolution;
  for (int s = 0; s < n_scans; s++) {
    const double px = pose[3 * s], py = pose[3 * s + 1], heading = pose[3 * s + 2];
    const double min_angle = laser[4 * s], ares = laser[4 * s + 1];
    const double min_range = laser[4 * s + 2], max_range = laser[4 * s + 3];
    const int nb = beam_ptr[s + 1] - beam_ptr[s];
    const double *r = ranges + beam_ptr[s];
    const int gx0 = (int)oc_round((px - d->offset_x) * scale);
    const int gy0 = (int)oc_round((py - d->offset_y) * scale);
    for (int i = 0; i < nb; i++) {
      const double rr = r[i];
      const int end_valid = rr < (range_threshold - 1e-6);
      if (rr <= min_range || rr >= max_range || isnan(rr)) continue;
      const double angle = heading + min_angle + (double)(uint32_t)i * ares;
      double x = px + (rr * cos(angle));
      double y = py + (rr * sin(angle));
      if (rr >= range_threshold) {
        const double ratio = range_threshold / rr;
        const double ddx = x - px, ddy = y - py;
        x = px + ratio * ddx;
        y = py + ratio * ddy;
      }
      const int gx1 = (int)oc_round((x - d->offset_x) * scale);
      const int gy1 = (int)oc_round((y - d->offset_y) * scale);
      trace_line(pass, w, h, gx0, gy0, gx1, gy1);
      if (end_valid && gx1 >= 0 && gx1 < w && gy1 >= 0 && gy1 < h) {
        pass[(size_t)gy1 * w + gx1]++;
        hit[(size_t)gy1 * w + gx1]++;
      }
    }
  }
  for (size_t c = 0; c < (size_t)w * h; c++) {
    uint8_t v = 200; /* Unknown */
    if (pass[c] > 2) {
      const double ratio = (double)hit[c] / (double)pass[c];
      v = ratio > 0.1 ? 0 /* Occupied */ : 255 /* Free */;
    }
    image[c] = v;
  }
  return 0;
}
