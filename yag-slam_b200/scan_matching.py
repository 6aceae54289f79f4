"""Mirror of the reference's matcher API (yag_slam/scan_matching.py:29-42).

`Scan2DMatcherCpp(config_dict, loop=False).match_scan(query, base_scans, penalty, do_fine)`
keeps the reference signature, argument meaning and return type
(ScanMatcherResult(response, covariance, best_pose, meta)); `match_scan_batch` is the
batched form used for loop-closure candidate chains / relocalisation / offline re-matching.
`query` / `base_scans` are yag_slam.models.LocalizedRangeScan-like objects exposing `._scan`.
"""
from collections import namedtuple

from .karto_compat import ScanMatcherConfig, Wrapper
from .matcher import DEFAULTS, DEFAULTS_LOOP
from .tf import Transform

ScanMatcherResult = namedtuple("ScanMatcherResult", ["response", "covariance", "best_pose", "meta"])

# reference yag_slam/helpers.py:339-361
default_config = {k: v for k, v in DEFAULTS.items() if k != "minimum_distance_penalty"}
default_config_loop = {k: v for k, v in DEFAULTS_LOOP.items() if k != "minimum_distance_penalty"}


def make_config(d):
    """reference yag_slam/helpers.py:364-376 (same assertion, same setattr loop)."""
    config = ScanMatcherConfig()
    config_params = default_config.copy()
    if d:
        config_params.update(d)
    assert 0.5 * config_params["resolution"] <= config_params["smear_deviation"] <= 10 * config_params["resolution"], \
        f"Smear deviation must be between {0.5 * config_params['resolution']} and {10 * config_params['resolution']}"
    for key, value in config_params.items():
        config.__setattr__(key, value)
    return config


def _transform_from_pose2(p):
    try:  # the reference returns a tiny_tf Transform when that package is present
        from tiny_tf.tf import Transform as T
        return T.from_pose2d(p)
    except ImportError:
        return Transform.from_pose2d(p)


class Scan2DMatcherCpp(object):
    def __init__(self, config_dict=None, loop=False, device=0, max_slots=0, max_grid_bytes=0):
        cfg = default_config if not loop else default_config_loop
        cfg = cfg.copy()
        cfg.update(config_dict)  # raises on None, like the reference (scan_matching.py:36)
        self.config = make_config(cfg)
        self._matcher = Wrapper(self.config, device=device, max_slots=max_slots, max_grid_bytes=max_grid_bytes)

    def match_scan(self, query, base_scans, penalty=True, do_fine=False):
        res = self._matcher.match_scan(query._scan, [b._scan for b in base_scans], penalty, do_fine)
        return ScanMatcherResult(res.response, res.covariance, _transform_from_pose2(res.best_pose), None)

    def match_scan_batch(self, queries, base_sets, penalty=True, do_fine=False):
        rs = self._matcher.match_scan_batch([q._scan for q in queries],
                                            [[b._scan for b in bs] for bs in base_sets], penalty, do_fine)
        return [ScanMatcherResult(r.response, r.covariance, _transform_from_pose2(r.best_pose), None) for r in rs]


Scan2DMatcherB200 = Scan2DMatcherCpp
