"""Stand-in for the slice of `shapely.geometry` (Shapely ~=1.7, GEOS 3.8) the
CAV-Gym reference calls (library/geometry.py:4,74-93).  TEST INFRASTRUCTURE ONLY.

Shapely/GEOS is not vendored under /root/reference and is not installable
here, so its published semantics are restated for CONVEX polygons only:

* `intersects`  = not disjoint, closed sets (DE-9IM: boundary touch counts).
* `contains(B)` = no point of B in the exterior of A and interiors meet.
* `intersection(B).area`, `area` = area of the convex clip / of the ring.
* `exterior.distance(Point)` = distance from a point to the ring.

Predicates are EXACT on the fp64 coordinates they are given (GEOS's orientation
predicate is robust, `CGAlgorithmsDD`): a float filter decides the clear cases,
`fractions.Fraction` decides the rest.  PARITY UNPINNED against upstream
Shapely: the reference has no tests or golden vectors at this boundary.
"""
import math
from fractions import Fraction

_EPS = 2.0 ** -52


def _orient_sign(a, b, c):
    """Sign of the exact cross product (b-a) x (c-a) on fp64 inputs."""
    l = (b[0] - a[0]) * (c[1] - a[1])
    r = (b[1] - a[1]) * (c[0] - a[0])
    det = l - r
    bound = 8.0 * _EPS * (abs(l) + abs(r))
    if det > bound:
        return 1
    if det < -bound:
        return -1
    ax, ay, bx, by, cx, cy = (Fraction(v) for v in (a[0], a[1], b[0], b[1], c[0], c[1]))
    exact = (bx - ax) * (cy - ay) - (by - ay) * (cx - ax)
    return (exact > 0) - (exact < 0)


def _signed_area2_exact(pts):
    acc = Fraction(0)
    n = len(pts)
    for i in range(n):
        x0, y0 = pts[i]
        x1, y1 = pts[(i + 1) % n]
        acc += Fraction(x0) * Fraction(y1) - Fraction(x1) * Fraction(y0)
    return acc


class Point:
    def __init__(self, x, y=None):
        if y is None:
            x, y = x
        self.x = float(x)
        self.y = float(y)


class _Ring:
    def __init__(self, pts):
        self._pts = pts

    def distance(self, point):
        px, py = point.x, point.y
        best = math.inf
        n = len(self._pts)
        for i in range(n):
            ax, ay = self._pts[i]
            bx, by = self._pts[(i + 1) % n]
            dx, dy = bx - ax, by - ay
            denom = dx * dx + dy * dy
            t = 0.0 if denom == 0 else max(0.0, min(1.0, ((px - ax) * dx + (py - ay) * dy) / denom))
            qx, qy = ax + t * dx, ay + t * dy
            best = min(best, math.hypot(px - qx, py - qy))
        return best


class _Area:
    """Result object of `intersection`: only `.area` is consumed by the reference."""

    def __init__(self, area):
        self.area = area


class Polygon:
    def __init__(self, shell):
        pts = [(float(x), float(y)) for x, y in shell]
        if len(pts) > 1 and pts[0] == pts[-1]:
            pts = pts[:-1]
        self._pts = pts
        # counter-clockwise copy for the half-plane predicates
        self._ccw = pts if _signed_area2_exact(pts) >= 0 else pts[::-1]

    @property
    def exterior(self):
        return _Ring(self._pts)

    @property
    def area(self):
        return float(abs(_signed_area2_exact(self._pts)) / 2)

    def _separates(self, other):
        """True if some edge line of self has every vertex of other strictly outside."""
        pts = self._ccw
        n = len(pts)
        for i in range(n):
            a, b = pts[i], pts[(i + 1) % n]
            if a == b:
                continue
            if all(_orient_sign(a, b, q) < 0 for q in other._ccw):
                return True
        return False

    def intersects(self, other):
        return not (self._separates(other) or other._separates(self))

    def contains(self, other):
        pts = self._ccw
        n = len(pts)
        for i in range(n):
            a, b = pts[i], pts[(i + 1) % n]
            if a == b:
                continue
            if any(_orient_sign(a, b, q) < 0 for q in other._ccw):
                return False
        return _signed_area2_exact(other._pts) != 0

    def intersection(self, other):
        """Exact Sutherland-Hodgman clip of self by convex `other`; area correctly rounded."""
        subject = [(Fraction(x), Fraction(y)) for x, y in self._ccw]
        clip = [(Fraction(x), Fraction(y)) for x, y in other._ccw]
        m = len(clip)
        for i in range(m):
            if not subject:
                break
            a, b = clip[i], clip[(i + 1) % m]
            if a == b:
                continue

            def side(p):
                return (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0])

            out = []
            k = len(subject)
            for j in range(k):
                p, q = subject[j], subject[(j + 1) % k]
                sp, sq = side(p), side(q)
                if sp >= 0:
                    out.append(p)
                if (sp > 0 and sq < 0) or (sp < 0 and sq > 0):
                    t = sp / (sp - sq)
                    out.append((p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])))
            subject = out
        if len(subject) < 3:
            return _Area(0.0)
        acc = Fraction(0)
        k = len(subject)
        for j in range(k):
            x0, y0 = subject[j]
            x1, y1 = subject[(j + 1) % k]
            acc += x0 * y1 - x1 * y0
        return _Area(float(abs(acc) / 2))
