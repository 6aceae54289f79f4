"""Seeded match scenarios shared by the CPU and GPU tests (SURVEY.md 8d shapes, scaled)."""
from yag_slam_b200.synth import make_match_batch as make_batch  # noqa: F401
from yag_slam_b200.synth import scan_points  # noqa: F401


def oracle_results(cfg, batch, penalty, do_fine, n_threads=0):
    from oracle import oracle
    return oracle.match_batch(cfg, batch["pool"], batch["starts"], batch["counts"], batch["query_scan"],
                              batch["query_pose"], batch["base_ptr"], batch["base_idx"], penalty, do_fine,
                              n_threads)
