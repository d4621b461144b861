"""o3d.io.read_point_cloud for PLY files (icp.py:112) -- vertex positions only."""
from __future__ import annotations

import numpy as np

from .geometry import PointCloud

_PLY_TYPES = {
    "char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2",
    "ushort": "u2", "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4",
    "float": "f4", "float32": "f4", "double": "f8", "float64": "f8",
}


def read_ply_vertices(path: str) -> np.ndarray:
    """Minimal PLY reader: returns the x,y,z of the vertex element as float64 [N,3]."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise RuntimeError(f"{path}: not a PLY file")
        fmt = None
        elements = []  # (name, count, [(prop, type) | (prop, 'list', count_t, item_t)])
        while True:
            line = f.readline()
            if not line:
                raise RuntimeError(f"{path}: truncated PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                elements.append((tok[1], int(tok[2]), []))
            elif tok[0] == "property":
                if tok[1] == "list":
                    elements[-1][2].append((tok[4], "list", tok[2], tok[3]))
                else:
                    elements[-1][2].append((tok[2], tok[1]))
            elif tok[0] == "end_header":
                break
        if fmt is None:
            raise RuntimeError(f"{path}: PLY header without format")
        for name, count, props in elements:
            if name != "vertex":
                # the reference's models list vertices first; anything before them would
                # need skipping, which only fixed-size elements allow
                if any(p[1] == "list" for p in props):
                    raise RuntimeError(f"{path}: list element '{name}' precedes vertices")
                if fmt == "ascii":
                    for _ in range(count):
                        f.readline()
                else:
                    size = sum(np.dtype(_PLY_TYPES[p[1]]).itemsize for p in props)
                    f.seek(size * count, 1)
                continue
            names = [p[0] for p in props]
            if any(p[1] == "list" for p in props) or not all(c in names for c in "xyz"):
                raise RuntimeError(f"{path}: vertex element must hold scalar x, y, z")
            if fmt == "ascii":
                rows = np.loadtxt(f, max_rows=count, ndmin=2) if count else np.zeros((0, len(names)))
                cols = [names.index(c) for c in "xyz"]
                return np.ascontiguousarray(rows[:, cols], dtype=np.float64)
            end = "<" if fmt == "binary_little_endian" else ">"
            dt = np.dtype([(p[0], end + _PLY_TYPES[p[1]]) for p in props])
            rec = np.frombuffer(f.read(dt.itemsize * count), dtype=dt, count=count)
            return np.stack([rec["x"], rec["y"], rec["z"]], axis=1).astype(np.float64)
    raise RuntimeError(f"{path}: no vertex element")


def read_point_cloud(filename: str, *args, **kwargs) -> PointCloud:
    if not str(filename).lower().endswith(".ply"):
        raise RuntimeError("read_point_cloud: only .ply is supported (icp.py:112 reads the CAD .ply)")
    return PointCloud(read_ply_vertices(filename))
