"""Regenerates tests/golden/*.npz from the reference's bundled Arrow IPC fixtures.

Run in the authoring container only (it reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
The fixtures are DATA files of the reference (ISO WKB geometry columns), not source code:
  data/cities.arrow                                             202 Points    (BASELINE config 1)
  py-geopolars/python/geopolars/datasets/naturalearth_cities.arrow   243 Points
  py-geopolars/python/geopolars/datasets/naturalearth_lowres.arrow   148 Polygon + 29 MultiPolygon rows
  py-geopolars/python/geopolars/datasets/nybb.arrow             5 MultiPolygons + Shape_Area / Shape_Leng
Each .npz holds the WKB column as (offsets int32, bytes uint8) plus any numeric attribute columns.
The only numeric pin the reference's own tests hold for this path is the 9-point contains vector of
geopolars/src/spatial_index.rs:432-484; it is written to contains_golden.npz.
"""
import os

import numpy as np
import pyarrow.ipc as ipc

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
FILES = {
    "cities": "data/cities.arrow",
    "naturalearth_cities": "py-geopolars/python/geopolars/datasets/naturalearth_cities.arrow",
    "naturalearth_lowres": "py-geopolars/python/geopolars/datasets/naturalearth_lowres.arrow",
    "nybb": "py-geopolars/python/geopolars/datasets/nybb.arrow",
}


def main():
    for name, rel in FILES.items():
        t = ipc.open_file(os.path.join(REF, rel)).read_all()
        col = t.column("geometry").combine_chunks()
        bufs = col.buffers()
        off = np.frombuffer(bufs[1], dtype=np.int32)[col.offset : col.offset + len(col) + 1].copy()
        data = np.frombuffer(bufs[2], dtype=np.uint8)[off[0] : off[-1]].copy()
        off -= off[0]
        extra = {}
        for c in t.column_names:
            if c in ("Shape_Area", "Shape_Leng", "pop_est"):
                extra[c] = np.asarray(t.column(c).to_numpy(), dtype=np.float64)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), offsets=off, bytes=data, **extra)
        print(name, len(col), "rows", len(data), "bytes")
    # the reference's contains vector (spatial_index.rs:432-484): 9 points x square [(0,0),(20,0),(20,20),(0,20)]
    pts = np.array([(0, 10), (1, 1), (10, 1), (1, -1), (0, -10), (-1, -1), (-10, 0), (-1, 1), (0, 10)], dtype=np.float64)
    square = np.array([(0, 0), (20, 0), (20, 20), (0, 20)], dtype=np.float64)  # polygon! macro closes it
    np.savez(os.path.join(OUT, "contains_golden.npz"), points=pts, square=square, inner_rows=np.array([1, 2]), inner_shape=np.array([2, 4]),
             left_shape=np.array([9, 4]))


if __name__ == "__main__":
    main()
