"""Test-only stand-in for the absent `sba_cpp` (sparse_bundle_adjustment) wheel: records nodes
and constraints, `compute` is a no-op (graph optimisation is out of scope)."""


class _Node(object):
    def __init__(self, x, y, yaw, idx):
        self.x, self.y, self.yaw, self.id = x, y, yaw, idx


class SPA2d(object):
    def __init__(self):
        self.nodes = []
        self.constraints = []

    def add_node(self, x, y, yaw, idx):
        self.nodes.append(_Node(x, y, yaw, idx))

    def add_constraint(self, i, j, dx, dy, dyaw, precision):
        self.constraints.append((i, j, dx, dy, dyaw, precision))

    def compute(self, *args):
        return 0
