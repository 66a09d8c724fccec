"""Host-side GeoArrow containers (numpy buffers) and the GeometryType enum.

GeometryType mirrors py-geopolars/python/geopolars/enums.py:4-15.  The buffer nesting is the GeoArrow
layout the reference's Python side assembles from shapely ragged arrays
(py-geopolars/python/geopolars/internals/geoseries.py:82-107, 164-214): coordinates + up to three
levels of List offsets + an Arrow validity bitmap.
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import IntEnum
from typing import Optional

import numpy as np


class GeometryType(IntEnum):
    """The enumeration of GEOS geometry types (reference: enums.py:4-15)."""

    MISSING = -1
    POINT = 0
    LINESTRING = 1
    LINEARRING = 2
    POLYGON = 3
    MULTIPOINT = 4
    MULTILINESTRING = 5
    MULTIPOLYGON = 6
    GEOMETRYCOLLECTION = 7


_LEVELS = {
    GeometryType.POINT: (),
    GeometryType.LINESTRING: ("geom_off",),
    GeometryType.MULTIPOINT: ("geom_off",),
    GeometryType.POLYGON: ("geom_off", "ring_off"),
    GeometryType.MULTILINESTRING: ("geom_off", "ring_off"),
    GeometryType.MULTIPOLYGON: ("geom_off", "part_off", "ring_off"),
}


@dataclass
class GeoArrowArray:
    """One chunk of a geometry column on the host: interleaved xy (n,2) f64, int64 offsets, bool validity."""

    type: GeometryType
    xy: np.ndarray
    geom_off: Optional[np.ndarray] = None
    part_off: Optional[np.ndarray] = None
    ring_off: Optional[np.ndarray] = None
    valid: Optional[np.ndarray] = None  # bool per geometry, None = all valid

    def __post_init__(self):
        self.type = GeometryType(int(self.type))
        if self.type not in _LEVELS:
            raise TypeError(f"unsupported geometry type {self.type!r}")
        self.xy = np.ascontiguousarray(self.xy, dtype=np.float64).reshape(-1, 2)
        for k in ("geom_off", "part_off", "ring_off"):
            v = getattr(self, k)
            if k in _LEVELS[self.type]:
                if v is None:
                    raise ValueError(f"{k} is required for {self.type.name}")
                setattr(self, k, np.ascontiguousarray(v, dtype=np.int64))
            else:
                setattr(self, k, None)
        if self.valid is not None:
            self.valid = np.ascontiguousarray(self.valid, dtype=bool)
            if self.valid.shape[0] != len(self):
                raise ValueError("validity length does not match the number of geometries")

    def __len__(self) -> int:
        if self.type == GeometryType.POINT:
            return int(self.xy.shape[0])
        return int(self.geom_off.shape[0] - 1)

    @property
    def n_coords(self) -> int:
        return int(self.xy.shape[0])

    @property
    def n_rings(self) -> int:
        return 0 if self.ring_off is None else int(self.ring_off.shape[0] - 1)

    @property
    def n_parts(self) -> int:
        return 0 if self.part_off is None else int(self.part_off.shape[0] - 1)

    def validity_bitmap(self) -> Optional[np.ndarray]:
        if self.valid is None:
            return None
        return np.packbits(self.valid, bitorder="little")

    # -- constructors ------------------------------------------------------------------------------
    @classmethod
    def points(cls, xy, valid=None) -> "GeoArrowArray":
        return cls(GeometryType.POINT, xy, valid=valid)

    @classmethod
    def linestrings(cls, xy, geom_off, valid=None) -> "GeoArrowArray":
        return cls(GeometryType.LINESTRING, xy, geom_off=geom_off, valid=valid)

    @classmethod
    def polygons(cls, xy, ring_off, geom_off, valid=None) -> "GeoArrowArray":
        return cls(GeometryType.POLYGON, xy, geom_off=geom_off, ring_off=ring_off, valid=valid)

    @classmethod
    def multipolygons(cls, xy, ring_off, part_off, geom_off, valid=None) -> "GeoArrowArray":
        return cls(GeometryType.MULTIPOLYGON, xy, geom_off=geom_off, part_off=part_off, ring_off=ring_off, valid=valid)

    @classmethod
    def from_shapes(cls, type: GeometryType, shapes) -> "GeoArrowArray":
        """Build from nested Python lists (tests / small inputs).

        POINT: [(x,y)|None]; LINESTRING/MULTIPOINT: [[(x,y),...]]; POLYGON/MULTILINESTRING:
        [[ring,...]]; MULTIPOLYGON: [[[ring,...],...]].  None = null row.
        """
        type = GeometryType(int(type))
        xy, geom, part, ring, valid = [], [0], [0], [0], []
        for s in shapes:
            valid.append(s is not None)
            if type == GeometryType.POINT:
                xy.append((np.nan, np.nan) if s is None else tuple(s))
                continue
            s = [] if s is None else s
            if type in (GeometryType.LINESTRING, GeometryType.MULTIPOINT):
                xy.extend(tuple(c) for c in s)
                geom.append(len(xy))
            elif type in (GeometryType.POLYGON, GeometryType.MULTILINESTRING):
                for r in s:
                    xy.extend(tuple(c) for c in r)
                    ring.append(len(xy))
                geom.append(len(ring) - 1)
            else:
                for poly in s:
                    for r in poly:
                        xy.extend(tuple(c) for c in r)
                        ring.append(len(xy))
                    part.append(len(ring) - 1)
                geom.append(len(part) - 1)
        kw = {}
        if "geom_off" in _LEVELS[type]:
            kw["geom_off"] = np.array(geom, dtype=np.int64)
        if "part_off" in _LEVELS[type]:
            kw["part_off"] = np.array(part, dtype=np.int64)
        if "ring_off" in _LEVELS[type]:
            kw["ring_off"] = np.array(ring, dtype=np.int64)
        v = np.array(valid, dtype=bool)
        return cls(type, np.array(xy, dtype=np.float64).reshape(-1, 2), valid=None if v.all() else v, **kw)

    def take_rows(self, lo: int, hi: int) -> "GeoArrowArray":
        """Contiguous row range [lo,hi) with rebased offsets (row-range sharding)."""
        t = self.type
        valid = None if self.valid is None else self.valid[lo:hi]
        if t == GeometryType.POINT:
            return GeoArrowArray(t, self.xy[lo:hi], valid=valid)
        g = self.geom_off[lo : hi + 1]
        if t in (GeometryType.LINESTRING, GeometryType.MULTIPOINT):
            return GeoArrowArray(t, self.xy[g[0] : g[-1]], geom_off=g - g[0], valid=valid)
        if t in (GeometryType.POLYGON, GeometryType.MULTILINESTRING):
            r = self.ring_off[g[0] : g[-1] + 1]
            return GeoArrowArray(t, self.xy[r[0] : r[-1]], geom_off=g - g[0], ring_off=r - r[0], valid=valid)
        p = self.part_off[g[0] : g[-1] + 1]
        r = self.ring_off[p[0] : p[-1] + 1]
        return GeoArrowArray(t, self.xy[r[0] : r[-1]], geom_off=g - g[0], part_off=p - p[0], ring_off=r - r[0], valid=valid)
