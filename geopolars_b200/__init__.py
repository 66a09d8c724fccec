"""geopolars_b200 — B200-native execution engine for the GeoPolars GeoSeries hot path.

The public surface mirrors the reference (py-geopolars/python/geopolars/__init__.py): GeoSeries with a
`.geo` accessor, `from_arrow`, GeometryType.  All geometry arithmetic runs in hand-written sm_100a CUDA
kernels behind the C ABI in include/geopolars_b200.h (libgeopolars_b200.so); there is no CPU fallback.
"""
from .geoarrow import GeoArrowArray, GeometryType  # noqa: F401
from ._lib import GeopolarsError, MismatchedGeometry, ShapeError  # noqa: F401

__all__ = ["GeoArrowArray", "GeometryType", "GeopolarsError", "MismatchedGeometry", "ShapeError"]
