"""Writes geopolars_b200/datasets/*.arrow (Arrow IPC file format, `geometry: binary` + numeric attributes) from the
golden fixtures tests/golden/*.npz (themselves extracted from the reference's data files by make_golden.py).
    python tests/golden/make_datasets.py
"""
import os

import numpy as np
import pyarrow as pa
import pyarrow.ipc as ipc

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(os.path.dirname(HERE)), "geopolars_b200", "datasets")


def main():
    for name in ("cities", "naturalearth_cities", "naturalearth_lowres", "nybb"):
        z = np.load(os.path.join(HERE, name + ".npz"))
        off, data = z["offsets"].astype(np.int32), z["bytes"]
        geom = pa.Array.from_buffers(pa.binary(), len(off) - 1, [None, pa.py_buffer(off.tobytes()), pa.py_buffer(data.tobytes())])
        cols, names = [geom], ["geometry"]
        for k in z.files:
            if k not in ("offsets", "bytes"):
                cols.append(pa.array(z[k]))
                names.append(k)
        t = pa.Table.from_arrays(cols, names=names)
        path = os.path.join(OUT, name + ".arrow")
        with ipc.new_file(path, t.schema) as w:
            w.write_table(t)
        print(name, t.num_rows, "rows ->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
