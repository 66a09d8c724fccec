"""Bundled example datasets — mirrors py-geopolars/python/geopolars/datasets/__init__.py:1-42 (`available`,
`get_path`, `read_dataset`).  The reference reads its Arrow IPC files with `polars.read_ipc`; polars is not installed
here, so the reader is `pyarrow.ipc` and the result is the pyarrow-backed GeoDataFrame of geopolars_b200.geoseries.

The `.arrow` files next to this module are Arrow IPC *file* format tables with a `geometry: binary` column holding the
same ISO WKB bytes as the reference's fixtures (and the numeric attribute columns kept by tests/golden/make_golden.py);
they are written by tests/golden/make_datasets.py from tests/golden/*.npz.  `cities` is the table of the reference's
data/cities.arrow (BASELINE config 1)."""
from __future__ import annotations

from pathlib import Path

__all__ = ["available", "get_path", "read_dataset"]

HERE = Path(__file__).parent.resolve()
available = ("naturalearth_cities", "nybb", "naturalearth_lowres", "cities")


def get_path(dataset: str) -> Path:
    """path of the Arrow IPC file of `dataset` (see `available`)"""
    if dataset in available:
        return HERE / (dataset + ".arrow")
    msg = f"The dataset '{dataset}' is not available. "
    msg += f"Available datasets are {', '.join(available)}"
    raise ValueError(msg)


def read_dataset(dataset: str):
    """GeoDataFrame of a bundled dataset: `read_dataset("nybb").geometry.geo.area` runs on the GPU"""
    import pyarrow.ipc as ipc

    from ..geoseries import GeoDataFrame

    path = get_path(dataset)
    with open(path, "rb") as f:
        table = ipc.open_file(f).read_all()  # memory_map=False like the reference
    return GeoDataFrame(table)
