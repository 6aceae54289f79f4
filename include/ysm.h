/*
 * ysm.h -- C ABI of the B200-native correlative scan matcher (libysm_b200.so).
 *
 * Drop-in boundary for ONE hot path of safijari/yag-slam: Karto's
 * ScanMatcher::MatchScan behind
 *     yag_slam/scan_matching.py:40-42   Scan2DMatcherCpp.match_scan
 *       -> karto_scanmatcher.Wrapper.match_scan(query._scan, [b._scan...], penalty, do_fine)
 * plus the numba ray-walk yag_slam/raytracing.py:63-92.
 *
 * The reference's boundary is a pybind11 module (external wheel
 * karto_scanmatcher==1.0.0, reference setup.py:46); it has no C ABI today. These entry
 * points are what a maintainer would bind instead (ctypes stub: INTEGRATION.md).
 * Plain pointers and sizes only; no exceptions cross the ABI; every function
 * returns 0 on success or a negative YSM_E* code (text via ysm_last_error).
 * There is NO CPU fallback: ysm_create fails if no CUDA device is usable.
 */
#ifndef YSM_H_
#define YSM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YSM_OK 0
#define YSM_EINVAL (-1)   /* bad argument / parameter out of bounds (e.g. smear deviation) */
#define YSM_ECUDA (-2)    /* CUDA runtime error */
#define YSM_ENOMEM (-3)   /* workspace does not fit */
#define YSM_EMATCH (-4)   /* "Unable to find best position" (Karto runtime_error) */
#define YSM_EUNSUP (-5)   /* parameter regime not supported yet */

/* Replaces karto_scanmatcher.ScanMatcherConfig (attribute names: reference
 * yag_slam/helpers.py:339-351; consumed by ScanMatcher::Create).
 * minimum_distance_penalty is Karto's fixed default 0.5 (not in yag's config). */
typedef struct ysm_params {
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
  int32_t use_response_expansion;
  int32_t max_slots;       /* correlation grids kept resident in HBM at once; 0 = auto */
  int64_t max_grid_bytes;  /* HBM budget for those grids; 0 = 16 GiB */
  int32_t lanes;           /* 0/1: single; 2..4: large batches are split over that many internal matcher
                              instances (own slots, stream, host thread) so host work overlaps kernels */
  int32_t resident_idle_us; /* single-query latency path: the resident kernel leaves the device after this many
                              microseconds without a request (0 = 2000; < 0 = never use the resident kernel) */
} ysm_params;

/* Derived sizes (ScanMatcher::Create / CorrelationGrid::CreateGrid). */
typedef struct ysm_dims {
  int32_t side, margin, roi, half_kernel, kernel_size, border;
  int32_t width, height, stride, slots;
  int64_t grid_bytes;
} ysm_dims;

/* One batch of independent MatchScan calls over a pool of scans.
 * A scan is its filtered world-frame point readings (what Karto's
 * LocalizedRangeScan::GetPointReadings() caches) -- compute them with
 * ysm_point_readings. Match i matches pool scan query_scan[i], whose sensor pose is
 * query_pose[i], against pool scans base_idx[base_ptr[i] .. base_ptr[i+1]). */
typedef struct ysm_batch {
  int32_t n_matches;
  int32_t n_scans;
  int64_t n_points;            /* total points in the pool */
  const double *pool_xy;       /* [n_points][2]; host, or device if pool_on_device */
  const int32_t *scan_start;   /* [n_scans]  host */
  const int32_t *scan_count;   /* [n_scans]  host */
  const int32_t *query_scan;   /* [n_matches] host */
  const double *query_pose;    /* [n_matches][3] x, y, heading; host */
  const int32_t *base_ptr;     /* [n_matches+1] host */
  const int32_t *base_idx;     /* [base_ptr[n_matches]] host */
  int32_t do_penalize;         /* Wrapper.match_scan arg 3 */
  int32_t do_refine;           /* Wrapper.match_scan arg 4 */
  int32_t pool_on_device;      /* 1: pool_xy is a device pointer already resident in HBM */
  int32_t _pad;
  const uint64_t *scan_tag;    /* [n_scans] host, or NULL. A non-zero tag names the CONTENT of a scan: the caller promises
                                  that two scans with the same tag (and point count) hold the same point readings. The
                                  single-query path keeps tagged scans in a device-resident store, so the running scans
                                  of sequential mapping (graph_slam.py:326) are uploaded once, not once per match.
                                  0 = untagged (always uploaded). */
  const int32_t *scan_raw_count; /* [n_scans] host, or NULL: RAW range readings of every scan (before the min_range /
                                  range_threshold filter). Karto's MatchScan returns early only for a scan without any
                                  range reading; a query with beams but no in-range reading goes through the whole
                                  schedule with an empty lookup table: GetResponse returns 0 for every pose, so every
                                  pose ties in every pass (all response expansions run) and the result is the ordered
                                  average of the search lattices -- response 0, a pose equal to the search centre up to
                                  rounding, covariance 500 / 500 / 4 coarse_res^2 (1000 fine_res^2 after a fine pass).
                                  That schedule has no grid lookup at all; the library's host runtime evaluates it.
                                  NULL: every scan without point readings is taken to have no beams (early return). */
} ysm_batch;

/* 128-byte result record (what Wrapper.match_scan returns: response, best_pose, covariance). */
typedef struct ysm_result {
  double response;
  double x, y, heading;
  double cov[9];          /* row-major 3x3 */
  int32_t n_passes;       /* CorrelateScan calls made (coarse + expansions + fine) */
  int32_t n_ties;         /* poses averaged in the last pass */
  int32_t status;         /* YSM_OK or YSM_EMATCH */
  int32_t _pad;
  double _reserved;       /* pads the record to 128 bytes */
} ysm_result;

typedef struct ysm_handle ysm_handle;

/* Replaces Wrapper(config) -> ScanMatcher::Create: allocates the correlation-grid slots,
 * kernel and workspaces once; reused by every call (reference ownership rule). */
int ysm_create(const ysm_params *params, int device, ysm_handle **out);
void ysm_destroy(ysm_handle *h);
const char *ysm_last_error(const ysm_handle *h); /* h may be NULL: last create error */
int ysm_get_dims(const ysm_handle *h, ysm_dims *out);

/* Replaces Wrapper.match_scan (ScanMatcher::MatchScan), batched. `out` is host memory,
 * [n_matches]. `stream` is a cudaStream_t (0 = default). Synchronous w.r.t. the host:
 * results are valid on return. Not re-entrant per handle (as the reference matcher). */
int ysm_match_batch(ysm_handle *h, const ysm_batch *batch, ysm_result *out, void *stream);

/* Replaces LocalizedRangeScan::Update (point readings; python twin
 * yag_slam/helpers.py:58-68): host libm, bit-identical to the CPU reference.
 * out_xy has room for n pairs; *n_out receives the number kept. */
int ysm_point_readings(const double *ranges, int32_t n, double min_angle,
                       double angular_resolution, double min_range, double range_threshold,
                       double x, double y, double heading, double *out_xy, int32_t *n_out);

/* The same for a whole batch of scans in ONE call (host libm on all the host threads the process may use):
 * scan i takes the range readings of source scan src[i] (ranges[beam_ptr[src[i]] .. beam_ptr[src[i] + 1]), all
 * sources share one laser) at sensor pose pose[i]. The kept readings are written back to back: scan i owns
 * out_xy[out_start[i] .. out_start[i] + out_count[i]). out_xy needs room for the sum of the source beam counts;
 * *n_points receives the total written. */
int ysm_point_readings_batch(const double *ranges, const int32_t *beam_ptr, int32_t n_src, const int32_t *src,
                             const double *pose, int32_t n, double min_angle, double angular_resolution,
                             double min_range, double range_threshold, double *out_xy, int32_t *out_start,
                             int32_t *out_count, int64_t *n_points);

/* Replaces yag_slam/raytracing.py:90-92 run_raytracing_sweep for n_starts start cells.
 * img: uint8 [h][w] (host, or device if img_on_device); angles in degrees (float64);
 * starts_xy [n_starts][2]; out (host) [n_starts][n_angles][5] float32 rows
 * start.x, start.y, end.x, end.y, length. */
int ysm_raytrace(const uint8_t *img, int32_t h, int32_t w, int32_t img_on_device,
                 const double *angles_deg, int32_t n_angles, const double *starts_xy,
                 int32_t n_starts, float *out, int device, void *stream);

/* ---- occupancy grid: replaces karto_scanmatcher.create_occupancy_grid(scans, resolution,
 * range_threshold) (reference yag_slam/graph_slam.py:341-342, ros1/slam_node_ros1:188-209;
 * Karto OccupancyGrid::CreateFromScans). Scans cross as raw range readings +
 * sensor pose + laser parameters (what the wheel's LocalizedRangeScan holds, models.py:37-39).
 * range_threshold replaces the lasers' own threshold for the whole construction. ---- */
typedef struct ysm_occ_scans {
  int32_t n_scans;
  int32_t _pad;
  const double *pose;      /* [n_scans][3] sensor pose x, y, heading; host */
  const double *laser;     /* [n_scans][4] min_angle, angular_resolution, min_range, max_range; host */
  const double *ranges;    /* raw range readings of all scans, concatenated; host */
  const int32_t *beam_ptr; /* [n_scans+1] first reading of each scan; host */
  double resolution;
  double range_threshold;
} ysm_occ_scans;

typedef struct ysm_occ_info {
  int32_t width, height;     /* cells (OccupancyGrid::ComputeDimensions) */
  double offset_x, offset_y; /* world position of cell (0,0) = bounding-box minimum */
  double resolution;
  int64_t rays;              /* raw beams examined */
  int64_t cells_visited;     /* Bresenham steps taken (pass-count increments incl. out-of-bounds steps) */
  int32_t box_candidates;    /* beams re-evaluated with host libm for the bounding box */
  int32_t cell_fixups;       /* beams whose end cell was re-evaluated with host libm */
  int32_t launches;          /* kernels launched */
  int32_t _pad;
} ysm_occ_info;

typedef struct ysm_occ ysm_occ;

/* Builds the grid on `device`; counts and image stay resident in HBM until ysm_occ_destroy.
 * Synchronous w.r.t. the host. EINVAL for an empty scan list (Karto returns NULL). */
int ysm_occ_create(const ysm_occ_scans *scans, int device, void *stream, ysm_occ **out);
void ysm_occ_destroy(ysm_occ *o);
int ysm_occ_get_info(const ysm_occ *o, ysm_occ_info *out);
/* image [height][width] uint8, row 0 = minimum y: 0 occupied, 200 unknown, 255 free
 * (the values ros1/slam_node_ros1:199-202 decodes) */
int ysm_occ_copy_image(const ysm_occ *o, uint8_t *out_host);
/* pass / hit counters [height][width] uint32 (parity tests) */
int ysm_occ_copy_counts(const ysm_occ *o, uint32_t *pass_host, uint32_t *hit_host);
/* device pointer of the image, for ysm_raytrace(img_on_device=1) without a host round trip */
const uint8_t *ysm_occ_device_image(const ysm_occ *o);
const char *ysm_occ_last_error(void); /* thread-local text of the last failing ysm_occ_* call */

/* ---- match against a map image: replaces the reference's (numba, unfinished) map path --
 * occupancy_grid_map_to_correlation_grid (yag_slam/helpers.py:24-34) +
 * Scan2DMatcherPy.match_scan_sets_with_map (yag_slam/scan_matching.py:124-173) -- with Karto's grid
 * semantics. The handle holds ONE resident correlation grid: image cell (u, v) is ROI cell (u, v) whose
 * world position is (offset_x + u*resolution, offset_y + v*resolution); cells equal to occupied_value
 * are 100 and smeared with params->smear_deviation. params->resolution must be the map's resolution.
 * ysm_match_batch on such a handle runs the usual MatchScan schedule of every query against that grid
 * (base lists are ignored; nothing is built or cleared per match). img: uint8 [h][w], host. ---- */
int ysm_create_map(const ysm_params *params, const uint8_t *img, int32_t h, int32_t w,
                   int32_t occupied_value, double offset_x, double offset_y, int device,
                   ysm_handle **out);

/* ---- loop-closure chain finder: replaces, for a batch of query scans, the reference's Python
 * GraphSlam.find_possible_loop_closure_chains (yag_slam/graph_slam.py:274-304) with its
 * breadth-first "near linked" traversal (yag_slam/graph.py:71-98, graph_slam.py:32-39) and
 * RadiusHashSearch.crude_radius_search (yag_slam/helpers.py:395-431). Vertex ids are scan numbers
 * (graph.vertices[scan.num], graph_slam.py:275). The result is CSR: chains of query i are
 * query_chain_ptr[i] .. query_chain_ptr[i+1]; chain c holds vertices members[chain_ptr[c] ..
 * chain_ptr[c+1]) -- i.e. directly the base_ptr / base_idx lists of ysm_batch. ---- */
typedef struct ysm_chain_query {
  int32_t n_vertices;
  int32_t n_queries;
  const double *pose_xy;        /* [n_vertices][2] current corrected poses; host */
  const double *hash_xy;        /* [n_vertices][2] poses the vertices had when RadiusHashSearch hashed them
                                   (add_vertex / last run_opt); NULL = pose_xy */
  const int32_t *adj_ptr;       /* [n_vertices+1] CSR adjacency over graph edges, both directions; host */
  const int32_t *adj_idx;       /* [adj_ptr[n_vertices]] */
  const int32_t *query_vertex;  /* [n_queries] scan numbers of the query scans */
  double loop_search_dist;      /* GraphSlam.loop_search_dist (also the hash resolution, graph_slam.py:67) */
  double crude_r2;              /* (radius + res)**2 as the caller's language evaluates it (helpers.py:423) */
  double near_dist_sq;          /* distance**2 of make_near_scan_visitor (graph_slam.py:33) */
  int32_t min_chain_size;       /* GraphSlam.loop_search_min_chain_size, >= 1 */
  int32_t _pad;
} ysm_chain_query;

typedef struct ysm_chains ysm_chains;

/* Runs the search on `device` (synchronous w.r.t. the host); the CSR result is held by *out. */
int ysm_chains_find(const ysm_chain_query *q, int device, void *stream, ysm_chains **out);
/* sizes of the result (+ kernels launched and their device time in ms; any pointer may be NULL) */
int ysm_chains_get_counts(const ysm_chains *c, int32_t *n_chains, int32_t *n_members, int32_t *launches,
                          double *kernel_ms);
/* query_chain_ptr [n_queries+1], chain_ptr [n_chains+1], members [n_members]; host, any may be NULL */
int ysm_chains_copy(const ysm_chains *c, int32_t *query_chain_ptr, int32_t *chain_ptr, int32_t *members);
void ysm_chains_destroy(ysm_chains *c);
const char *ysm_chains_last_error(void); /* thread-local text of the last failing ysm_chains_* call */

/* ---- introspection for parity tests (not part of the reference surface) ---- */
#define YSM_DEBUG_KEEP_GRIDS 1 /* do not clear the slot grids after a batch */
int ysm_set_debug(ysm_handle *h, int32_t flags);
/* copies the correlation grid built for match `i` of the last batch (stride*height bytes) */
int ysm_debug_copy_grid(ysm_handle *h, int32_t match, uint8_t *out_host);
/* copies the smear kernel (kernel_size^2 bytes) */
int ysm_debug_copy_kernel(ysm_handle *h, uint8_t *out_host);
/* copies the COARSE lookup-offset table of match `i` of the last batch: [n_angles][n_points] */
int ysm_debug_copy_offsets(ysm_handle *h, int32_t match, int32_t *out_host, int32_t *n_angles,
                           int32_t *n_points);
/* number of kernels launched by this handle since creation */
int64_t ysm_launch_count(const ysm_handle *h);
/* device milliseconds spent in the sweep kernel (coarse lattice) during the last batch,
 * from CUDA events on the call's stream (only when YSM_DEBUG_TIME_KERNELS is set) */
#define YSM_DEBUG_TIME_KERNELS 2
#define YSM_DEBUG_NO_PRUNE 4 /* always use the unpruned lattice sweep (A/B checks of the zero-row pruning) */
#define YSM_DEBUG_NO_SPECULATE 8 /* latency path: do not chain the fine pass on the device (A/B checks) */
#define YSM_DEBUG_NO_MEGA 16     /* latency path: separate kernels instead of the single cooperative kernel */
#define YSM_DEBUG_NO_CANDLISTS 32 /* grid build: search the cell list per tile instead of exact per-tile candidate lists (A/B checks) */
#define YSM_DEBUG_NO_HALF_LISTS 64 /* grid build: k_tile_stamp's per-candidate step loop instead of k_tile_stamp_lists (A/B checks) */
#define YSM_DEBUG_NO_FINE9 128 /* fine pass: k_sweep_points (a warp per pose) instead of k_sweep_fine9 (A/B checks) */
int ysm_last_kernel_ms(const ysm_handle *h, double *sweep_ms, double *build_ms, double *reduce_ms,
                       double *total_ms);

/* Round trip of an empty request through the resident latency kernel (doorbell in mapped host memory ->
 * kernel -> tagged result chunk): the floor under every single-query latency. Starts the kernel if
 * needed; rtt_us receives n wall-clock round trips in microseconds. */
int ysm_debug_ping(ysm_handle *h, int32_t n, double *rtt_us);

/* work done by the last ysm_match_batch call (for roofline accounting in bench.py):
 * out[0] grid lookups of the coarse lattice sweeps (L), out[1] lattice sweep launches,
 * out[2] lookup-offset table entries computed (T), out[3] poses evaluated,
 * out[4] grid lookups of the fine / angular-covariance passes, out[5] base points stamped (upper bound),
 * out[6] host->device bytes, out[7] device->host bytes, out[8] sweep launches that used zero-row pruning,
 * out[9] lattice lookups actually issued after pruning (counted only with YSM_DEBUG_TIME_KERNELS),
 * out[10] fine passes that ran chained on the device behind their coarse pass (latency path),
 * out[11] lanes used by the call, out[12] launches of the single-kernel latency path,
 * out[13] requests served by the resident latency kernel, out[14] scans it took from the device-resident
 * scan store instead of host memory, out[15] base points that survived FindValidPoints and the ROI test
 * (P_valid; counted only with YSM_DEBUG_TIME_KERNELS);
 * fills out[0..n), n <= 16 */
int ysm_last_work(const ysm_handle *h, int64_t *out, int32_t n);

#ifdef __cplusplus
}
#endif
#endif /* YSM_H_ */
