"""Seeded synthetic 2-D world, LiDAR scans and trajectories.

Used by tests/ and bench.py to make the same inputs for the CUDA path and the CPU
oracle. Pure numpy; no reference code involved (the reference ships no datasets,
yag_slam/helpers.py:607-610).
"""
import numpy as np

WORLD_SEED = 1234
ROOM_W, ROOM_H = 40.0, 30.0


class World:
    def __init__(self, segments):
        self.segments = np.ascontiguousarray(segments, dtype=np.float64)  # (S, 4): x1 y1 x2 y2


def loop_path(n, step=0.25, half_w=12.0, half_h=7.0, radius=3.0, start_s=0.0):
    """n poses (x, y, heading) on a closed rounded rectangle, arc-length step `step`."""
    sx, sy = half_w - radius, half_h - radius
    segs = [("L", (-sx, -half_h), 0.0, 2 * sx), ("A", (sx, -sy), -np.pi / 2, radius * np.pi / 2),
            ("L", (half_w, -sy), np.pi / 2, 2 * sy), ("A", (sx, sy), 0.0, radius * np.pi / 2),
            ("L", (sx, half_h), np.pi, 2 * sx), ("A", (-sx, sy), np.pi / 2, radius * np.pi / 2),
            ("L", (-half_w, sy), -np.pi / 2, 2 * sy), ("A", (-sx, -sy), np.pi, radius * np.pi / 2)]
    per = sum(s[3] for s in segs)
    out = np.zeros((n, 3))
    for i in range(n):
        s = (start_s + i * step) % per
        for kind, p0, ang, length in segs:
            if s <= length:
                if kind == "L":
                    out[i] = (p0[0] + s * np.cos(ang), p0[1] + s * np.sin(ang), ang)
                else:
                    a = ang + s / radius
                    out[i] = (p0[0] + radius * np.cos(a), p0[1] + radius * np.sin(a), a + np.pi / 2)
                break
            s -= length
    out[:, 2] = (out[:, 2] + np.pi) % (2 * np.pi) - np.pi
    return out


def make_world(seed=WORLD_SEED, n_pillars=24, keep_clear=None, clear_dist=1.2):
    """Outer 40x30 m rectangle + axis-aligned rectangular pillars (side U[0.5,3])."""
    rng = np.random.default_rng(seed)
    hw, hh = ROOM_W / 2, ROOM_H / 2
    segs = [(-hw, -hh, hw, -hh), (hw, -hh, hw, hh), (hw, hh, -hw, hh), (-hw, hh, -hw, -hh)]
    if keep_clear is None:
        keep_clear = loop_path(600, step=0.125)[:, :2]
    placed = 0
    while placed < n_pillars:
        cx, cy = rng.uniform(-hw + 2, hw - 2), rng.uniform(-hh + 2, hh - 2)
        w, h = rng.uniform(0.5, 3.0), rng.uniform(0.5, 3.0)
        d = np.abs(keep_clear - np.array([cx, cy]))
        if np.any((d[:, 0] < w / 2 + clear_dist) & (d[:, 1] < h / 2 + clear_dist)):
            continue
        x0, x1, y0, y1 = cx - w / 2, cx + w / 2, cy - h / 2, cy + h / 2
        segs += [(x0, y0, x1, y0), (x1, y0, x1, y1), (x1, y1, x0, y1), (x0, y1, x0, y0)]
        placed += 1
    return World(np.array(segs))


def cast_scan(world, pose, n_beams, noise_rng=None, noise=0.01, max_range=30.0, min_angle=-np.pi,
              angle_increment=None):
    """Analytic ray-segment ranges in float64 (+ optional N(0, noise)); all finite."""
    if angle_increment is None:
        angle_increment = 2 * np.pi / n_beams
    x, y, h = pose
    ang = h + min_angle + np.arange(n_beams) * angle_increment
    dx, dy = np.cos(ang)[:, None], np.sin(ang)[:, None]
    s = world.segments
    ex, ey = (s[:, 2] - s[:, 0])[None, :], (s[:, 3] - s[:, 1])[None, :]
    wx, wy = (s[:, 0] - x)[None, :], (s[:, 1] - y)[None, :]
    den = dx * ey - dy * ex
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (wx * ey - wy * ex) / den
        u = (wx * dy - wy * dx) / den
    ok = (np.abs(den) > 1e-12) & (t > 1e-9) & (u >= 0.0) & (u <= 1.0)
    t = np.where(ok, t, np.inf)
    r = t.min(axis=1)
    r = np.where(np.isfinite(r), r, max_range)
    if noise_rng is not None and noise > 0:
        r = r + noise_rng.normal(0.0, noise, size=n_beams)
    return np.clip(r, 0.06, max_range)


def laser_params(n_beams, range_threshold=20.0):
    """(min_angle, max_angle, angle_increment, min_range, max_range, range_threshold)."""
    inc = 2 * np.pi / n_beams
    return (-np.pi, -np.pi + inc * (n_beams - 1), inc, 0.05, 30.0, range_threshold)


def noisy_odometry(poses, rng, sigma_xy=0.02, sigma_t=0.01):
    """odom = truth + cumulative N(0, sigma) drift (SURVEY 8d cfg 2)."""
    n = len(poses)
    drift = np.cumsum(np.column_stack([rng.normal(0, sigma_xy, n), rng.normal(0, sigma_xy, n),
                                       rng.normal(0, sigma_t, n)]), axis=0)
    drift[0] = 0
    out = poses + drift
    out[:, 2] = (out[:, 2] + np.pi) % (2 * np.pi) - np.pi
    return out


def occupancy_image(world, resolution=0.05, pad=1.0, unknown_outside=True):
    """Rasterise the world into the reference's map convention (0 occupied, 200 unknown,
    255 free; ros1/slam_node_ros1:199-202). Returns (img uint8 HxW, origin_xy)."""
    hw, hh = ROOM_W / 2 + pad, ROOM_H / 2 + pad
    w, h = int(round(2 * hw / resolution)), int(round(2 * hh / resolution))
    img = np.full((h, w), 200 if unknown_outside else 255, dtype=np.uint8)
    x0, y0 = -hw, -hh
    ix0, ix1 = int(round((-ROOM_W / 2 - x0) / resolution)), int(round((ROOM_W / 2 - x0) / resolution))
    iy0, iy1 = int(round((-ROOM_H / 2 - y0) / resolution)), int(round((ROOM_H / 2 - y0) / resolution))
    img[iy0:iy1 + 1, ix0:ix1 + 1] = 255
    for (ax, ay, bx, by) in world.segments:
        n = int(np.hypot(bx - ax, by - ay) / (resolution * 0.5)) + 2
        xs, ys = np.linspace(ax, bx, n), np.linspace(ay, by, n)
        img[np.clip(np.round((ys - y0) / resolution).astype(int), 0, h - 1),
            np.clip(np.round((xs - x0) / resolution).astype(int), 0, w - 1)] = 0
    # pillar interiors are unknown (never observed)
    segs = world.segments[4:].reshape(-1, 4, 4)
    for p in segs:
        xa, xb = p[:, [0, 2]].min(), p[:, [0, 2]].max()
        ya, yb = p[:, [1, 3]].min(), p[:, [1, 3]].max()
        i0, i1 = int(round((xa - x0) / resolution)) + 1, int(round((xb - x0) / resolution))
        j0, j1 = int(round((ya - y0) / resolution)) + 1, int(round((yb - y0) / resolution))
        if i1 > i0 and j1 > j0:
            img[j0:j1, i0:i1] = 200
    return img, (x0, y0)


# ---- match problems -------------------------------------------------------
def scan_points(world, pose, n_beams, rng, sense_pose=None, range_threshold=20.0):
    """Point readings of a scan taken at `sense_pose` (truth) but localised at `pose`."""
    lp = laser_params(n_beams, range_threshold)
    r = cast_scan(world, pose if sense_pose is None else sense_pose, n_beams, rng)
    from . import _capi
    return _capi.point_readings(r, lp[0], lp[2], lp[3], lp[5], pose[0], pose[1], pose[2])


def make_match_batch(world, n_matches, n_beams, n_base, seed, perturb=(0.2, 0.15), degenerate_frac=0.0,
               range_threshold=20.0, shared_query=False, path_step=0.25, path_start=0.0):
    """n_matches independent (query, n_base running scans) problems along the loop path.
    Returns dict(pool, starts, counts, query_scan, query_pose, base_ptr, base_idx, points=list)."""
    rng = np.random.default_rng(seed)
    path = loop_path(n_matches + n_base + 1, step=path_step, start_s=path_start)
    pts = []
    base_of = {}

    def base_scan(k):
        if k not in base_of:
            base_of[k] = len(pts)
            pts.append(scan_points(world, path[k], n_beams, rng, range_threshold=range_threshold))
        return base_of[k]

    query_scan, query_pose, base_ptr, base_idx = [], [], [0], []
    shared_q = None
    for i in range(n_matches):
        k = i + n_base
        true_pose = path[k]
        guess = true_pose + np.array([rng.uniform(-perturb[0], perturb[0]), rng.uniform(-perturb[0], perturb[0]),
                                      rng.uniform(-perturb[1], perturb[1])])
        if shared_query:
            if shared_q is None:
                shared_q = (len(pts), guess)
                pts.append(scan_points(world, guess, n_beams, rng, sense_pose=true_pose,
                                       range_threshold=range_threshold))
            qid, guess = shared_q
        else:
            qid = len(pts)
            pts.append(scan_points(world, guess, n_beams, rng, sense_pose=true_pose,
                                   range_threshold=range_threshold))
        query_scan.append(qid)
        query_pose.append(guess)
        if rng.random() < degenerate_frac:
            # chain with no point inside the ROI: an empty base scan
            eid = len(pts)
            pts.append(np.zeros((0, 2)))
            base_idx.append(eid)
        else:
            j0 = 0 if shared_query else i
            for j in range(j0, j0 + n_base):
                base_idx.append(base_scan(j if not shared_query else (i % 7) + (j - j0)))
        base_ptr.append(len(base_idx))
    from .matcher import pack_pool
    pool, starts, counts = pack_pool(pts)
    return dict(pool=pool, starts=starts, counts=counts, query_scan=np.array(query_scan, np.int32),
                query_pose=np.array(query_pose, np.float64), base_ptr=np.array(base_ptr, np.int32),
                base_idx=np.array(base_idx, np.int32), points=pts)




# ---- raw scan logs for the occupancy grid --------------------------------
def make_scan_log(world, n_scans, n_beams, seed, step=0.25, defects=0.0):
    """n_scans raw scans along the loop path: dict(poses [n][3], lasers [n][4] = min_angle,
    angular_resolution, min_range, max_range, ranges (concatenated), beam_ptr [n+1]).
    `defects` = fraction of readings replaced by NaN / inf / below-min / at-max values."""
    rng = np.random.default_rng(seed)
    path = loop_path(n_scans, step=step)
    lp = laser_params(n_beams)
    ranges = np.empty((n_scans, n_beams))
    for i in range(n_scans):
        ranges[i] = cast_scan(world, path[i], n_beams, rng)
    if defects > 0:
        m = rng.random(ranges.shape)
        ranges[m < defects * 0.25] = np.nan
        ranges[(m >= defects * 0.25) & (m < defects * 0.5)] = np.inf
        ranges[(m >= defects * 0.5) & (m < defects * 0.75)] = 0.01
        ranges[(m >= defects * 0.75) & (m < defects)] = lp[4]
    lasers = np.tile(np.array([lp[0], lp[2], lp[3], lp[4]]), (n_scans, 1))
    beam_ptr = (np.arange(n_scans + 1) * n_beams).astype(np.int32)
    return dict(poses=np.ascontiguousarray(path), lasers=lasers, ranges=ranges.reshape(-1), beam_ptr=beam_ptr)


# ---- batched relocalisation (SURVEY.md 8d cfg 5) -------------------------------------------
def make_relocalisation_batch(world, n_matches, n_beams, n_base, seed, n_log=2000, perturb=(0.2, 0.15),
                              range_threshold=20.0, path_step=0.25, n_use=None, readings=None):
    """n_matches independent (query, n_base-scan base set) pairs sampled from an n_log-scan log
    along the loop path: match i takes log scan k_i (uniform in [n_base, n_log)), localises it at
    its true pose + U(+-perturb) and matches it against log scans k_i-n_base .. k_i-1 at their true
    poses. The log's ranges are cast once; only the query point readings are per match.
    n_use: build only the first n_use matches (the SAME matches as the first n_use of the full batch: the random
    draws do not depend on it). readings: None = the library's batched LocalizedRangeScan::Update
    (ysm_point_readings_batch), or a callable (ranges, min_angle, angular_resolution, min_range, range_threshold,
    x, y, heading) -> (k, 2) array, e.g. the oracle's (the CPU reference arm must not load the product library).
    Same dict as make_match_batch (without `points`) plus `log_scan` [n] and `truth`."""
    rng = np.random.default_rng(seed)
    path = loop_path(n_log, step=path_step)
    lp = laser_params(n_beams, range_threshold)
    ranges = np.stack([cast_scan(world, path[k], n_beams, rng) for k in range(n_log)])
    k = rng.integers(n_base, n_log, size=n_matches)
    d = np.column_stack([rng.uniform(-perturb[0], perturb[0], n_matches), rng.uniform(-perturb[0], perturb[0], n_matches),
                         rng.uniform(-perturb[1], perturb[1], n_matches)])
    n = n_matches if n_use is None else min(int(n_use), n_matches)
    k, guess = k[:n], (path[k] + d)[:n]
    src = np.concatenate([np.arange(n_log), k]).astype(np.int32)  # scans 0 .. n_log-1 = the log, then the queries
    poses = np.concatenate([path, guess])
    beam_ptr = (np.arange(n_log + 1) * n_beams).astype(np.int32)
    if readings is None:
        from . import _capi
        pool, starts, counts = _capi.point_readings_batch(ranges.reshape(-1), beam_ptr, src, poses, lp[0], lp[2], lp[3], lp[5])
    else:
        from .matcher import pack_pool
        need = np.zeros(n_log, bool)
        need[(k[:, None] - n_base + np.arange(n_base)[None, :]).reshape(-1)] = True
        pts = [readings(ranges[j], lp[0], lp[2], lp[3], lp[5], *path[j]) if need[j] else np.zeros((0, 2)) for j in range(n_log)]
        pts += [readings(ranges[k[i]], lp[0], lp[2], lp[3], lp[5], *guess[i]) for i in range(n)]
        pool, starts, counts = pack_pool(pts)
    base_idx = (k[:, None] - n_base + np.arange(n_base)[None, :]).astype(np.int32).reshape(-1)
    return dict(pool=pool, starts=starts, counts=counts, query_scan=(n_log + np.arange(n)).astype(np.int32),
                query_pose=np.ascontiguousarray(guess), base_ptr=(np.arange(n + 1) * n_base).astype(np.int32),
                base_idx=base_idx, log_scan=k.astype(np.int32), truth=path[k])
