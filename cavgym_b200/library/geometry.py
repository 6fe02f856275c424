"""Host-side geometry vocabulary of the CAV-Gym plugin surface.

Mirrors the names the reference's scenario modules and agents use
(library/geometry.py:7-61 Point, :96-238 ConvexQuadrilateral, :241-251
make_rectangle, :305-377 Triangle, :380-385 normalise_angle, :388-428 Line) so
that reference-style scenario definitions run unchanged.  These objects only
DESCRIBE a scenario: the per-step geometry (oriented boxes, SAT, clipping) runs
in the CUDA engine (csrc/geometry.cuh).  The predicates kept here
(`intersects`, `contains`, `percentage_intersects`, reference :74-87) are exact
rational-arithmetic host helpers for scenario authoring, not a step fallback.
Curved shapes (CircleSegment/Arc/Arrow/Zone, reference :254-302,431-578) are dead
code in the reference and are not provided.
"""
import math
from fractions import Fraction
from typing import NamedTuple


class Point(NamedTuple):
    x: float
    y: float

    def distance_x(self, other):
        return abs(other.x - self.x)

    def distance_y(self, other):
        return abs(other.y - self.y)

    def distance(self, other):
        return math.sqrt(((other.y - self.y) ** 2) + ((other.x - self.x) ** 2))

    def translate(self, anchor):
        return Point(anchor.x + self.x, anchor.y + self.y)

    def rotate(self, angle):
        if angle == 0:
            return self
        c, s = math.cos(angle), math.sin(angle)
        return Point((c * self.x) - (s * self.y), (s * self.x) + (c * self.y))

    def transform(self, angle, anchor):
        return self.rotate(angle).translate(anchor)

    def rescale(self, center, x_scale=1, y_scale=1):
        return Point(center.x + ((self.x - center.x) * x_scale), center.y + ((self.y - center.y) * y_scale))

    def enlarge(self, center, scale=100):
        return self.rescale(center, x_scale=scale, y_scale=scale)

    def __add__(self, other):
        return Point(self.x + other.x, self.y + other.y)

    def __sub__(self, other):
        return Point(self.x - other.x, self.y - other.y)

    def __mul__(self, factor):
        return Point(self.x * factor, self.y * factor)

    def __copy__(self):
        return Point(self.x, self.y)


def _lerp(a, b, t):
    return Point((a.x * (1 - t)) + (b.x * t), (a.y * (1 - t)) + (b.y * t))


def _mid(a, b):
    return Point((a.x + b.x) * 0.5, (a.y + b.y) * 0.5)


def _exact_ring(shape):
    ring = [(Fraction(x), Fraction(y)) for x, y in shape]
    twice_area = sum(ring[i][0] * ring[(i + 1) % len(ring)][1] - ring[(i + 1) % len(ring)][0] * ring[i][1]
                     for i in range(len(ring)))
    return ring if twice_area >= 0 else ring[::-1]


def _side(a, b, p):
    return (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0])


def _has_separating_edge(ring, other):
    n = len(ring)
    return any(ring[i] != ring[(i + 1) % n] and all(_side(ring[i], ring[(i + 1) % n], q) < 0 for q in other)
               for i in range(n))


class Shape:
    """Convex-polygon predicates with Shapely's closed-set semantics (reference :64-93)."""

    def intersects(self, other):
        a, b = _exact_ring(self), _exact_ring(other)
        return not (_has_separating_edge(a, b) or _has_separating_edge(b, a))

    def contains(self, other):
        a, b = _exact_ring(self), _exact_ring(other)
        n = len(a)
        return all(_side(a[i], a[(i + 1) % n], q) >= 0 for i in range(n) if a[i] != a[(i + 1) % n] for q in b)

    def percentage_intersects(self, other):
        if not self.intersects(other):
            return 0
        if other.contains(self):
            return 1
        subject, clip = _exact_ring(self), _exact_ring(other)
        own_area = _ring_area(subject)
        for i in range(len(clip)):
            a, b = clip[i], clip[(i + 1) % len(clip)]
            if a == b or not subject:
                continue
            kept = []
            for j in range(len(subject)):
                p, q = subject[j], subject[(j + 1) % len(subject)]
                sp, sq = _side(a, b, p), _side(a, b, q)
                if sp >= 0:
                    kept.append(p)
                if sp * sq < 0:
                    t = sp / (sp - sq)
                    kept.append((p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])))
            subject = kept
        return float(_ring_area(subject) / own_area) if len(subject) >= 3 else 0.0

    def mostly_intersects(self, other):
        return self.percentage_intersects(other) > 0.5

    def distance(self, point):
        ring = [Point(*p) for p in self]
        px, py = point
        best = math.inf
        for i, a in enumerate(ring):
            b = ring[(i + 1) % len(ring)]
            dx, dy = b.x - a.x, b.y - a.y
            denom = dx * dx + dy * dy
            t = 0.0 if denom == 0 else max(0.0, min(1.0, ((px - a.x) * dx + (py - a.y) * dy) / denom))
            best = min(best, math.hypot(px - (a.x + t * dx), py - (a.y + t * dy)))
        return best


def _ring_area(ring):
    return abs(sum(ring[i][0] * ring[(i + 1) % len(ring)][1] - ring[(i + 1) % len(ring)][0] * ring[i][1]
                   for i in range(len(ring)))) / 2


class ConvexQuadrilateral(Shape):
    """Corners in the reference's order: rear_left, front_left, front_right, rear_right."""
    __slots__ = ("rear_left", "front_left", "front_right", "rear_right")

    def __init__(self, rear_left, front_left, front_right, rear_right):
        self.rear_left, self.front_left = Point(*rear_left), Point(*front_left)
        self.front_right, self.rear_right = Point(*front_right), Point(*rear_right)

    def corners(self):
        return (self.rear_left, self.front_left, self.front_right, self.rear_right)

    def __iter__(self):
        return iter(tuple(p) for p in self.corners())

    def __eq__(self, other):
        return isinstance(other, ConvexQuadrilateral) and self.corners() == other.corners()

    def __hash__(self):
        return hash(self.corners())

    def __repr__(self):
        return "ConvexQuadrilateral(rear_left={}, front_left={}, front_right={}, rear_right={})".format(*self.corners())

    def _map(self, fn):
        return ConvexQuadrilateral(*(fn(p) for p in self.corners()))

    def translate(self, position):
        return self._map(lambda p: p.translate(position))

    def transform(self, orientation, position):
        if orientation == 0:
            return self.translate(position)
        return self._map(lambda p: p.transform(orientation, position))

    def rescale(self, **kwargs):
        centre = self.centre()
        return self._map(lambda p: p.rescale(centre, **kwargs))

    def centre(self):
        return _mid(self.front_left, self.rear_right)

    def front_centre(self):
        return _mid(self.front_left, self.front_right)

    def rear_centre(self):
        return _mid(self.rear_left, self.rear_right)

    def left_centre(self):
        return _mid(self.rear_left, self.front_left)

    def right_centre(self):
        return _mid(self.rear_right, self.front_right)

    def split_laterally(self, left_percentage=0.5):
        front, rear = _lerp(self.front_left, self.front_right, left_percentage), _lerp(self.rear_left, self.rear_right, left_percentage)
        return (ConvexQuadrilateral(self.rear_left, self.front_left, front, rear),
                ConvexQuadrilateral(rear, front, self.front_right, self.rear_right))

    def split_longitudinally(self, rear_percentage=0.5):
        left, right = _lerp(self.rear_left, self.front_left, rear_percentage), _lerp(self.rear_right, self.front_right, rear_percentage)
        return (ConvexQuadrilateral(self.rear_left, left, right, self.rear_right),
                ConvexQuadrilateral(left, self.front_left, self.front_right, right))

    def divide_laterally(self, segments):
        remainder = self
        for k in range(segments, 1, -1):
            segment, remainder = remainder.split_laterally(left_percentage=1 / k)
            yield segment
        yield remainder

    def flip_laterally(self):
        return ConvexQuadrilateral(self.rear_right, self.front_right, self.front_left, self.rear_left)

    def flip_longitudinally(self):
        return ConvexQuadrilateral(self.front_left, self.rear_left, self.rear_right, self.front_right)

    def flip(self):
        return self.flip_longitudinally().flip_laterally()

    def longitudinal_line(self):
        return Line(self.rear_centre(), self.front_centre())

    def triangles(self):
        return (Triangle(self.front_left, self.front_right, self.rear_left),
                Triangle(self.rear_right, self.rear_left, self.front_right))

    def area(self):
        return sum(triangle.area() for triangle in self.triangles())

    def random_point(self, np_random):
        left, right = self.triangles()
        left_fraction = left.area() / (left.area() + right.area())
        return np_random.choice([left, right], p=[left_fraction, 1 - left_fraction]).random_point(np_random)


def make_rectangle(length, width, anchor=Point(0, 0), rear_offset=0.5, left_offset=0.5):
    rear, front = anchor.x - (length * rear_offset), anchor.x + (length * (1 - rear_offset))
    left, right = anchor.y + (width * left_offset), anchor.y - (width * (1 - left_offset))
    return ConvexQuadrilateral((rear, left), (front, left), (front, right), (rear, right))


class Triangle(Shape):
    __slots__ = ("rear", "front_left", "front_right")

    def __init__(self, rear, front_left, front_right):
        self.rear, self.front_left, self.front_right = Point(*rear), Point(*front_left), Point(*front_right)

    def __iter__(self):
        return iter((tuple(self.rear), tuple(self.front_left), tuple(self.front_right)))

    def angle(self):
        def bearing(p):
            return math.atan2(p.y - self.rear.y, p.x - self.rear.x)
        angle = bearing(self.front_left) - bearing(self.front_right)
        if angle > math.pi:
            angle -= 2 * math.pi
        elif angle <= -math.pi:
            angle += 2 * math.pi
        return angle

    def normalise(self):
        return self if self.angle() >= 0 else Triangle(self.rear, self.front_right, self.front_left)

    def translate(self, position):
        return Triangle(*(p.translate(position) for p in (self.rear, self.front_left, self.front_right)))

    def transform(self, orientation, position):
        return Triangle(*(p.transform(orientation, position) for p in (self.rear, self.front_left, self.front_right)))

    def area(self):
        r, fl, fr = self.rear, self.front_left, self.front_right
        return abs((r.x * (fl.y - fr.y) + fl.x * (fr.y - r.y) + fr.x * (r.y - fl.y)) / 2)

    def includes(self, point):
        signs = [_side(a, b, point) for a, b in ((self.rear, self.front_left), (self.front_left, self.front_right),
                                                 (self.front_right, self.rear))]
        return not (any(s < 0 for s in signs) and any(s > 0 for s in signs))

    def random_point(self, np_random):
        u, v = np_random.uniform(0.0, 1.0), np_random.uniform(0.0, 1.0)
        if u + v > 1:
            u, v = 1 - u, 1 - v
        return self.rear + (self.front_left - self.rear) * u + (self.front_right - self.rear) * v


def normalise_angle(radians):
    while radians <= -math.pi:
        radians += 2 * math.pi
    while radians > math.pi:
        radians -= 2 * math.pi
    return radians + 0.0 if radians == 0 else radians


class Line(Shape):
    __slots__ = ("start", "end")

    def __init__(self, start, end):
        self.start, self.end = Point(*start), Point(*end)

    def __iter__(self):
        return iter((tuple(self.start), tuple(self.end)))

    def translate(self, position):
        return Line(self.start.translate(position), self.end.translate(position))

    def transform(self, orientation, position):
        return Line(self.start.transform(orientation, position), self.end.transform(orientation, position))

    def closest_point_from(self, point):
        dx, dy = self.end.x - self.start.x, self.end.y - self.start.y
        a = (dy * (point.y - self.start.y) + dx * (point.x - self.start.x)) / ((dx * dx) + (dy * dy))
        return Point(self.start.x + a * dx, self.start.y + a * dy)

    def orientation(self):
        return math.atan2(self.end.y - self.start.y, self.end.x - self.start.x)

    def random_point(self, np_random):
        deviation = np_random.uniform(0.0, 1.0)
        return _lerp(self.start, self.end, deviation)
