"""CPU tests of the occupancy-grid oracle (oracle/occgrid_oracle.c) on
hand-computable cases, and of the closed-form Bresenham step the CUDA kernel k_occ_trace uses
(checked here against a plain restatement of Karto's Grid::TraceLine loop)."""
import math

import numpy as np

from oracle import oracle

LASER = (-math.pi, 2 * math.pi / 4, 0.05, 30.0)  # 4 beams: -x, -y, +x, +y


def _grid(poses, ranges, res=1.0, thr=10.0, laser=LASER):
    poses = np.atleast_2d(np.asarray(poses, float))
    n = len(poses)
    ranges = np.asarray(ranges, float).reshape(n, -1)
    bp = (np.arange(n + 1) * ranges.shape[1]).astype(np.int32)
    return oracle.occupancy_grid(poses, np.tile(laser, (n, 1)), ranges.reshape(-1), bp, res, thr)


def test_dimensions_are_the_bounding_box_of_filtered_points_and_poses():
    g = _grid([(0.0, 0.0, 0.0)], [[3.0, 2.0, 5.0, 4.0]])
    # beams: (-3,0) (0,-2) (5,0) (0,4) up to cos/sin rounding of multiples of pi/2
    assert (g["width"], g["height"]) == (8, 6)
    assert abs(g["offset_x"] + 3.0) < 1e-12 and abs(g["offset_y"] + 2.0) < 1e-12
    # a reading beyond the threshold does not grow the box
    g2 = _grid([(0.0, 0.0, 0.0)], [[3.0, 2.0, 25.0, 4.0]], thr=10.0)
    assert g2["width"] == 3 and g2["height"] == 6


def test_pass_and_hit_counts_single_scan():
    g = _grid([(0.0, 0.0, 0.0)], [[3.0, 2.0, 5.0, 4.0]])
    p, h = g["passes"], g["hits"]
    ox, oy = 3, 2  # sensor cell
    assert p[oy, ox] == 4  # every ray starts here
    # the +x / +y rays end at cell x = 8 == width / y = 6 == height: out of bounds (the box maximum
    # always is), so no hit is recorded for them
    assert h.sum() == 2 and h[oy, 0] == 1 and h[0, ox] == 1 and h[oy + 3, ox] == 0
    assert p[oy, 0] == 2  # traced once + end point once
    assert (p[oy, ox + 1:8] == 1).all()
    # nothing is classified with pass <= 2 except the sensor cell (free: 0 hits of 4 passes)
    img = g["image"]
    assert img[oy, ox] == 255 and (img[img != 200].size == 1)


def test_classification_threshold():
    # three identical scans: end cells get pass 6 / hit 3 -> occupied; ray cells pass 3 -> free
    g = _grid([(0.0, 0.0, 0.0)] * 3, [[3.0, 2.0, 5.0, 4.0]] * 3)
    img = g["image"]
    assert img[2, 0] == 0 and img[0, 3] == 0
    assert img[2, 1] == 255 and img[2, 2] == 255 and img[1, 3] == 255
    assert img[0, 0] == 200


def test_long_reading_is_shortened_and_has_no_hit():
    g = _grid([(0.0, 0.0, 0.0)] * 3, [[3.0, 2.0, 25.0, 4.0]] * 3, thr=10.0)
    # the 25 m beam is traced for 10 m (clipped by the 3-cell-wide box) and leaves no hit
    assert g["width"] == 3 and g["hits"][2].sum() == 3  # only the -x end point, three times
    g = _grid([(0.0, 0.0, 0.0)] * 2, [[3.0, 2.0, 5.0, 4.0], [3.0, 2.0, float("nan"), 0.01]])
    assert g["hits"].sum() == 4  # NaN and below-min readings are ignored
    assert g["passes"][2, 4] == 1 and g["passes"][3, 3] == 1  # only the first scan's +x / +y rays


def _trace_serial(x0, y0, x1, y1):
    steep = abs(y1 - y0) > abs(x1 - x0)
    if steep:
        x0, y0, x1, y1 = y0, x0, y1, x1
    if x0 > x1:
        x0, x1, y0, y1 = x1, x0, y1, y0
    dx, dy, err, y = x1 - x0, abs(y1 - y0), 0, y0
    ystep = 1 if y0 < y1 else -1
    out = []
    for x in range(x0, x1 + 1):
        out.append((y, x) if steep else (x, y))
        err += dy
        if 2 * err >= dx:
            y += ystep
            err -= dx
    return out


def _trace_closed_form(x0, y0, x1, y1):
    steep = abs(y1 - y0) > abs(x1 - x0)
    if steep:
        x0, y0, x1, y1 = y0, x0, y1, x1
    if x0 > x1:
        x0, x1, y0, y1 = x1, x0, y1, y0
    dmaj, dmin = x1 - x0, abs(y1 - y0)
    ystep = 1 if y0 < y1 else -1
    out = []
    for k in range(dmaj + 1):
        q = (2 * k * dmin + dmaj) // (2 * dmaj) if dmaj else 0
        out.append((y0 + ystep * q, x0 + k) if steep else (x0 + k, y0 + ystep * q))
    return out


def test_closed_form_bresenham_equals_the_serial_loop():
    rng = np.random.default_rng(0)
    cases = [(0, 0, 0, 0), (0, 0, 5, 0), (0, 0, 0, 5), (0, 0, 4, 2), (0, 0, 2, 4), (3, 3, -3, 0), (0, 0, 7, 7),
             (0, 0, 6, 3), (0, 0, -6, 3), (0, 0, 3, -6), (2, -1, -9, -5)]
    cases += [tuple(int(v) for v in rng.integers(-400, 400, 4)) for _ in range(3000)]
    for c in cases:
        assert _trace_serial(*c) == _trace_closed_form(*c), c


def test_oracle_trace_matches_python_restatement():
    # one scan, many beams: pass counts equal the python serial Bresenham accumulation
    n = 64
    laser = (-math.pi, 2 * math.pi / n, 0.05, 30.0)
    r = 3.0 + 4.0 * np.random.default_rng(1).random(n)
    g = _grid([(0.3, -0.2, 0.4)], [r], res=0.1, thr=12.0, laser=laser)
    w, h = g["width"], g["height"]
    acc = np.zeros((h, w), np.uint32)
    hits = np.zeros((h, w), np.uint32)
    rnd = lambda v: math.floor(v + 0.5) if v >= 0 else math.ceil(v - 0.5)  # noqa: E731
    scale = 1.0 / 0.1
    gx0, gy0 = rnd((0.3 - g["offset_x"]) * scale), rnd((-0.2 - g["offset_y"]) * scale)
    for i in range(n):
        a = 0.4 + laser[0] + float(i) * laser[1]
        x, y = 0.3 + r[i] * math.cos(a), -0.2 + r[i] * math.sin(a)
        gx1, gy1 = rnd((x - g["offset_x"]) * scale), rnd((y - g["offset_y"]) * scale)
        for (cx, cy) in _trace_serial(gx0, gy0, gx1, gy1):
            if 0 <= cx < w and 0 <= cy < h:
                acc[cy, cx] += 1
        if 0 <= gx1 < w and 0 <= gy1 < h:
            acc[gy1, gx1] += 1
            hits[gy1, gx1] += 1
    assert (acc == g["passes"]).all() and (hits == g["hits"]).all()
