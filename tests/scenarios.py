"""Seeded match scenarios shared by the CPU and GPU tests (SURVEY.md 8d shapes, scaled)."""
import numpy as np

from yag_slam_b200 import _capi, synth


def scan_points(world, pose, n_beams, rng, sense_pose=None, range_threshold=20.0):
    """Point readings of a scan taken at `sense_pose` (truth) but localised at `pose`."""
    lp = synth.laser_params(n_beams, range_threshold)
    r = synth.cast_scan(world, pose if sense_pose is None else sense_pose, n_beams, rng)
    return _capi.point_readings(r, lp[0], lp[2], lp[3], lp[5], pose[0], pose[1], pose[2])


def make_batch(world, n_matches, n_beams, n_base, seed, perturb=(0.2, 0.15), degenerate_frac=0.0,
               range_threshold=20.0, shared_query=False, path_step=0.25):
    """n_matches independent (query, n_base running scans) problems along the loop path.
    Returns dict(pool, starts, counts, query_scan, query_pose, base_ptr, base_idx, points=list)."""
    rng = np.random.default_rng(seed)
    path = synth.loop_path(n_matches + n_base + 1, step=path_step)
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
    from yag_slam_b200.matcher import pack_pool
    pool, starts, counts = pack_pool(pts)
    return dict(pool=pool, starts=starts, counts=counts, query_scan=np.array(query_scan, np.int32),
                query_pose=np.array(query_pose, np.float64), base_ptr=np.array(base_ptr, np.int32),
                base_idx=np.array(base_idx, np.int32), points=pts)


def oracle_results(cfg, batch, penalty, do_fine, n_threads=0):
    from oracle import oracle
    return oracle.match_batch(cfg, batch["pool"], batch["starts"], batch["counts"], batch["query_scan"],
                              batch["query_pose"], batch["base_ptr"], batch["base_idx"], penalty, do_fine,
                              n_threads)
