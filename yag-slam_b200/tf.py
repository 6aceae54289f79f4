"""Minimal SE(2)-in-SE(3) Transform with the subset of tiny_tf.tf.Transform's interface that
yag_slam's consumers use: x, y, z, qx..qw, euler, from_pose2d,
from_position_euler, from_xyt, `a + b` (compose), `a - b` (b^-1 o a). Used only when tiny_tf
is not installed, and as the test shim for it; not on the hot path."""
import math


class Transform(object):
    def __init__(self, x=0.0, y=0.0, z=0.0, qx=0.0, qy=0.0, qz=0.0, qw=1.0):
        self.x, self.y, self.z = float(x), float(y), float(z)
        self.qx, self.qy, self.qz, self.qw = float(qx), float(qy), float(qz), float(qw)

    @property
    def yaw(self):
        return math.atan2(2.0 * (self.qw * self.qz + self.qx * self.qy),
                          1.0 - 2.0 * (self.qy * self.qy + self.qz * self.qz))

    @property
    def euler(self):
        return (0.0, 0.0, self.yaw)

    @property
    def position(self):
        return (self.x, self.y, self.z)

    @property
    def quaternion(self):
        return (self.qx, self.qy, self.qz, self.qw)

    @classmethod
    def from_position_euler(cls, x, y, z, roll, pitch, yaw):
        assert roll == 0 and pitch == 0, "planar transforms only"
        return cls(x, y, z, 0.0, 0.0, math.sin(yaw / 2.0), math.cos(yaw / 2.0))

    @classmethod
    def from_xyt(cls, x, y, t):
        return cls.from_position_euler(x, y, 0, 0, 0, t)

    @classmethod
    def from_xyt_deg(cls, x, y, t):
        return cls.from_xyt(x, y, math.radians(t))

    @classmethod
    def from_pose2d(cls, p):
        return cls.from_position_euler(p.x, p.y, 0, 0, 0, p.yaw)

    def inverse(self):
        t = self.yaw
        c, s = math.cos(-t), math.sin(-t)
        return Transform.from_xyt(-(c * self.x - s * self.y), -(s * self.x + c * self.y), -t)

    def __add__(self, other):  # compose: self o other
        t = self.yaw
        c, s = math.cos(t), math.sin(t)
        return Transform.from_position_euler(self.x + c * other.x - s * other.y,
                                             self.y + s * other.x + c * other.y, self.z + other.z, 0, 0,
                                             t + other.yaw)

    def __sub__(self, other):  # pose of self expressed in other: other^-1 o self
        return other.inverse() + self

    def __repr__(self):
        return "Transform(x={:.6f}, y={:.6f}, yaw={:.6f})".format(self.x, self.y, self.yaw)
