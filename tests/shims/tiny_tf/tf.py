from yag_slam_b200.tf import Transform  # noqa: F401
