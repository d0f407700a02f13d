"""Synthetic floor plans for the BASELINE.json configs.

Every plan is a list of axis-aligned wall segments (x1, y1, x2, y2) on half-integer
coordinates, to be gridded at spacing 1.0: the outer box (0.5,0.5)-(W+0.5,H+0.5) gives a
(W+2)x(H+2) grid (PointMap::setGrid, salalib/pointdata.cpp:122-171) whose W x H interior cells
are open unless a closed room excludes them.  No wall passes through a cell centre.
Geometry is deterministic given (name, seed); `random.Random(seed)` only places doors.

These are inputs only -- the same segments go to the reference (`depthmapXcli -m IMPORT` of an
x1,y1,x2,y2 CSV, or oracle/ref_harness.cpp) and to this package.
"""
from __future__ import annotations

import random
from dataclasses import dataclass, field
from typing import List, Tuple

Seg = Tuple[float, float, float, float]


@dataclass
class Plan:
    name: str
    width: int
    height: int
    walls: List[Seg] = field(default_factory=list)
    seeds: List[Tuple[float, float]] = field(default_factory=list)
    spacing: float = 1.0

    def csv(self) -> str:
        """x1,y1,x2,y2 CSV accepted by `depthmapXcli -m IMPORT -it drawing`."""
        rows = ["x1,y1,x2,y2"]
        rows += [f"{a:.10g},{b:.10g},{c:.10g},{d:.10g}" for a, b, c, d in self.walls]
        return "\n".join(rows) + "\n"


def _box(w: int, h: int) -> List[Seg]:
    x0, y0, x1, y1 = 0.5, 0.5, w + 0.5, h + 0.5
    return [(x0, y0, x0, y1), (x0, y1, x1, y1), (x1, y1, x1, y0), (x1, y0, x0, y0)]


def _wall_with_doors(fixed: float, a: float, b: float, doors, vertical: bool) -> List[Seg]:
    """A wall along one axis from a to b at coordinate `fixed`, with openings [(d0,d1),...]."""
    out: List[Seg] = []
    cur = a
    for d0, d1 in sorted(doors):
        if d0 > cur:
            out.append((fixed, cur, fixed, d0) if vertical else (cur, fixed, d0, fixed))
        cur = max(cur, d1)
    if cur < b:
        out.append((fixed, cur, fixed, b) if vertical else (cur, fixed, b, fixed))
    return out


def room(w: int = 100, h: int = 100, seed: int = 0) -> Plan:
    """C1: one W x H room, two full-height partitions with one door each and one half wall."""
    rng = random.Random(seed)
    p = Plan(f"room{w}x{h}", w, h, _box(w, h), [(1.0, 1.0)])
    x1 = 0.5 + w // 3
    x2 = 0.5 + (2 * w) // 3
    for x in (x1, x2):
        d = 0.5 + rng.randrange(2, h - 6)
        p.walls += _wall_with_doors(x, 0.5, h + 0.5, [(d, d + 3.0)], True)
    yh = 0.5 + h // 2
    p.walls.append((x1, yh, x1 + int((x2 - x1) // 2), yh))
    return p


def office(w: int = 256, h: int = 256, seed: int = 1, room_size: int = 16, corridor: int = 4, door: int = 2,
           closed_fraction: float = 0.0) -> Plan:
    """C2/C3: bands of square rooms separated by corridors; every open room has one door onto the
    corridor below or above it; a vertical spine corridor on the left joins the corridors.
    `closed_fraction` of the rooms get no door (closed cores -> unfilled cells)."""
    rng = random.Random(seed)
    p = Plan(f"office{w}x{h}", w, h, _box(w, h), [(1.0, 1.0)])
    xs = 0.5 + corridor  # room bands start right of the spine corridor
    y = 0.5
    band = 0
    while y + room_size <= h + 0.5 + 1e-9:
        y0, y1 = y, y + room_size
        has_below = band > 0
        has_above = (y1 + corridor) <= h + 0.5 + 1e-9
        # room cells along x
        edges = []
        x = xs
        while x < w + 0.5 - 1e-9:
            edges.append((x, min(x + room_size, w + 0.5)))
            x += room_size
        doors_bottom, doors_top = [], []
        for (xa, xb) in edges:
            span = int(xb - xa)
            closed = rng.random() < closed_fraction
            off = rng.randrange(1, max(2, span - door))
            side_top = has_above and (not has_below or rng.random() < 0.5)
            if closed or span < door + 2:
                continue
            if side_top:
                doors_top.append((xa + off, xa + off + door))
            elif has_below:
                doors_bottom.append((xa + off, xa + off + door))
            else:
                doors_top.append((xa + off, xa + off + door))
        # horizontal walls of the band (skip those coinciding with the outer box)
        if has_below:
            p.walls += _wall_with_doors(y0, xs, w + 0.5, doors_bottom, False)
        if y1 < h + 0.5 - 1e-9:
            p.walls += _wall_with_doors(y1, xs, w + 0.5, doors_top, False)
        # vertical walls between rooms, and the wall against the spine
        p.walls.append((xs, y0, xs, y1))
        for (xa, xb) in edges[:-1]:
            p.walls.append((xb, y0, xb, y1))
        y = y1 + corridor
        band += 1
    return p


def gallery(w: int = 512, h: int = 512, seed: int = 3, door: int = 6) -> Plan:
    """C4: large halls (64-128 cells) joined by enfilade doors (doors aligned on a common axis)."""
    rng = random.Random(seed)
    p = Plan(f"gallery{w}x{h}", w, h, _box(w, h), [(1.0, 1.0)])

    def cuts(total: int) -> List[int]:
        out, cur = [], 0
        while total - cur > 128:
            step = rng.choice([64, 96, 128])
            if total - (cur + step) < 64:
                break
            cur += step
            out.append(cur)
        return out

    xc = cuts(w)
    yc = cuts(h)
    xb = [0] + xc + [w]
    yb = [0] + yc + [h]
    # enfilade axes: one door position per row of halls (for vertical walls) and per column
    row_axis = [rng.randrange(yb[j] + 8, yb[j + 1] - 8 - door) for j in range(len(yb) - 1)]
    col_axis = [rng.randrange(xb[i] + 8, xb[i + 1] - 8 - door) for i in range(len(xb) - 1)]
    for xcut in xc:
        doors = [(0.5 + a, 0.5 + a + door) for a in row_axis]
        p.walls += _wall_with_doors(0.5 + xcut, 0.5, h + 0.5, doors, True)
    for ycut in yc:
        doors = [(0.5 + a, 0.5 + a + door) for a in col_axis]
        p.walls += _wall_with_doors(0.5 + ycut, 0.5, w + 0.5, doors, False)
    return p


def urban(w: int = 1024, h: int = 1024, seed: int = 4, street: int = 8, door: int = 3) -> Plan:
    """C5: rectangular blocks with walled perimeters; block interiors are open and reachable through
    one door onto a street, so (almost) every cell is filled."""
    rng = random.Random(seed)
    p = Plan(f"urban{w}x{h}", w, h, _box(w, h), [(1.0, 1.0)])

    def blocks(total: int) -> List[Tuple[int, int]]:
        out, cur = [], street
        while cur + 24 <= total - street:
            size = rng.choice([40, 56, 72, 88])
            size = min(size, total - street - cur)
            if size < 24:
                break
            out.append((cur, cur + size))
            cur += size + street
        return out

    bx = blocks(w)
    by = blocks(h)
    for (xa, xb) in bx:
        for (ya, yb) in by:
            x0, x1, y0, y1 = 0.5 + xa, 0.5 + xb, 0.5 + ya, 0.5 + yb
            side = rng.randrange(4)
            d_h = rng.randrange(2, int(x1 - x0) - door - 2)
            d_v = rng.randrange(2, int(y1 - y0) - door - 2)
            p.walls += _wall_with_doors(y0, x0, x1, [(x0 + d_h, x0 + d_h + door)] if side == 0 else [], False)
            p.walls += _wall_with_doors(y1, x0, x1, [(x0 + d_h, x0 + d_h + door)] if side == 1 else [], False)
            p.walls += _wall_with_doors(x0, y0, y1, [(y0 + d_v, y0 + d_v + door)] if side == 2 else [], True)
            p.walls += _wall_with_doors(x1, y0, y1, [(y0 + d_v, y0 + d_v + door)] if side == 3 else [], True)
    return p


def oblique(w: int = 30, h: int = 30, seed: int = 7, n_axis: int = 8, n_oblique: int = 3, spacing: float = 1.0) -> Plan:
    """Small parity case with oblique walls (exercises crop / tanify / tolerance paths that the
    axis-aligned plans never reach).  Endpoints are kept off cell centres."""
    rng = random.Random(seed)
    p = Plan(f"oblique{w}x{h}s{spacing}", w, h, _box(w, h), [(1.0, 1.0)], spacing)
    for _ in range(n_axis):
        if rng.random() < 0.5:
            x = 0.5 + rng.randrange(2, w - 2)
            a = 0.5 + rng.randrange(0, h - 6)
            p.walls.append((x, a, x, a + rng.randrange(3, 6 + h // 3)))
        else:
            y = 0.5 + rng.randrange(2, h - 2)
            a = 0.5 + rng.randrange(0, w - 6)
            p.walls.append((a, y, a + rng.randrange(3, 6 + w // 3), y))
    for _ in range(n_oblique):
        x1 = 0.5 + rng.randrange(2, w - 2) + 0.25
        y1 = 0.5 + rng.randrange(2, h - 2) + 0.125
        x2 = min(w + 0.25, max(0.75, x1 + rng.randrange(-10, 11) + 0.5))
        y2 = min(h + 0.25, max(0.75, y1 + rng.randrange(-10, 11) + 0.25))
        p.walls.append((x1, y1, x2, y2))
    # clip everything to the box
    p.walls = [(min(max(a, 0.5), w + 0.5), min(max(b, 0.5), h + 0.5), min(max(c, 0.5), w + 0.5), min(max(d, 0.5), h + 0.5))
               for a, b, c, d in p.walls]
    return p


CONFIGS = {
    "C1": lambda: room(100, 100, seed=0),
    "C2": lambda: office(256, 256, seed=1),
    "C3": lambda: office(256, 256, seed=1),
    "C4": lambda: gallery(512, 512, seed=3),
    "C5": lambda: urban(1024, 1024, seed=4),
}


def by_name(name: str) -> Plan:
    """'C1'..'C5' or 'room:W:H:seed', 'office:W:H:seed', 'gallery:W:H:seed', 'urban:W:H:seed',
    'oblique:W:H:seed[:spacing]'."""
    if name in CONFIGS:
        return CONFIGS[name]()
    parts = name.split(":")
    kind = parts[0]
    w, h, seed = int(parts[1]), int(parts[2]), int(parts[3])
    if kind == "room":
        return room(w, h, seed)
    if kind == "office":
        return office(w, h, seed)
    if kind == "gallery":
        return gallery(w, h, seed)
    if kind == "urban":
        return urban(w, h, seed)
    if kind == "oblique":
        return oblique(w, h, seed, spacing=float(parts[4]) if len(parts) > 4 else 1.0)
    raise ValueError(f"unknown plan {name!r}")
