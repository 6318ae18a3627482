"""Gaussian PLY files -- SURVEY.md section 8(f4): the on-disk format GS-SR's three Gaussian families save and load.

Attribute order (all float32, one `vertex` element), as written by the reference's save_gaussians:
  vanilla / 2DGS / PGSR  x y z nx ny nz f_dc_* f_rest_* opacity scale_* rot_*
                         (/root/reference/gssr/gaussian/vanilla_gaussian.py:140-169)
  scaffold               x y z nx ny nz f_offset_* f_anchor_feat_* opacity scale_* rot_*
                         (/root/reference/gssr/gaussian/scaffold_gaussian.py:388-416)
  octree                 x y z nx ny nz level extra_level info f_offset_* f_anchor_feat_* opacity scale_* rot_*
                         (/root/reference/gssr/gaussian/octree_gaussian.py:276-310; info[0] = voxel_size, info[1] = standard_dist)
Multi-channel blocks are stored CHANNEL-major: features (P, K, 3) are transposed to (P, 3, K) before flattening, so
f_rest_0..K-1 are the K coefficients of the red channel, and so on (same for f_offset_*).

The reference goes through the `plyfile` package (not installed here); this module writes and parses the PLY container
itself (header + packed little-endian records; ASCII and big-endian files are read too).
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2",
              "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4",
              "double": "f8", "float64": "f8"}


# ---- container --------------------------------------------------------------------------------------------------------
def write_vertex_ply(path, names: Sequence[str], table):
    """One `vertex` element with float32 properties `names`; table is (P, len(names))."""
    table = np.ascontiguousarray(table, dtype="<f4")
    if table.ndim != 2 or table.shape[1] != len(names):
        raise ValueError(f"table shape {table.shape} does not match {len(names)} properties")
    head = ["ply", "format binary_little_endian 1.0", f"element vertex {table.shape[0]}"]
    head += [f"property float {n}" for n in names] + ["end_header"]
    with open(path, "wb") as f:
        f.write(("\n".join(head) + "\n").encode("ascii"))
        f.write(table.tobytes())


def read_vertex_ply(path) -> Dict[str, np.ndarray]:
    """{property name: (P,) array} of the first element of a PLY file (scalar properties only)."""
    with open(path, "rb") as f:
        data = f.read()
    end = data.index(b"end_header")
    end = data.index(b"\n", end) + 1
    fmt, count, props, in_first = None, None, [], False
    for line in data[:end].decode("ascii", "replace").split("\n"):
        t = line.split()
        if not t:
            continue
        if t[0] == "format":
            fmt = t[1]
        elif t[0] == "element":
            if count is None:
                count, in_first = int(t[2]), True
            else:
                in_first = False
        elif t[0] == "property" and in_first:
            if t[1] == "list":
                raise ValueError("list properties are not supported (Gaussian PLYs have none)")
            props.append((t[2], _PLY_TYPES[t[1]]))
    if fmt is None or count is None:
        raise ValueError(f"{path}: not a PLY file")
    if fmt == "ascii":
        rows = np.array(data[end:].split()[: count * len(props)], dtype=np.float64).reshape(count, len(props))
        return {n: rows[:, i].astype(ty) for i, (n, ty) in enumerate(props)}
    order = "<" if fmt == "binary_little_endian" else ">"
    rec = np.dtype([(n, order + ty) for n, ty in props])
    arr = np.frombuffer(data, dtype=rec, count=count, offset=end)
    return {n: np.ascontiguousarray(arr[n]) for n, _ in props}


def _numbered(cols: Dict[str, np.ndarray], prefix: str) -> np.ndarray:
    """(P, k) block of the properties prefix0, prefix1, ... ordered by their numeric suffix (the reference sorts the same way)."""
    names = sorted((n for n in cols if n.startswith(prefix)), key=lambda x: int(x.split("_")[-1]))
    P = len(next(iter(cols.values()))) if cols else 0
    return np.stack([cols[n] for n in names], axis=1).astype(np.float32) if names else np.zeros((P, 0), np.float32)


def _channel_major(a):
    """(P, K, C) -> (P, C*K): transpose(1, 2).flatten(start_dim=1) of the reference."""
    a = np.asarray(a, np.float32)
    return np.ascontiguousarray(a.transpose(0, 2, 1)).reshape(a.shape[0], -1)


def _from_channel_major(flat, channels=3):
    """(P, C*K) -> (P, K, C)."""
    flat = np.asarray(flat, np.float32)
    return np.ascontiguousarray(flat.reshape(flat.shape[0], channels, -1).transpose(0, 2, 1))


def _names(prefix, n) -> List[str]:
    return [f"{prefix}{i}" for i in range(n)]


# ---- vanilla (3DGS / 2DGS / PGSR) -------------------------------------------------------------------------------------
def vanilla_attributes(n_dc=3, n_rest=45, n_scale=3, n_rot=4) -> List[str]:
    return (["x", "y", "z", "nx", "ny", "nz"] + _names("f_dc_", n_dc) + _names("f_rest_", n_rest) + ["opacity"]
            + _names("scale_", n_scale) + _names("rot_", n_rot))


def save_vanilla_ply(path, xyz, features_dc, features_rest, opacity, scaling, rotation):
    """features_dc (P,1,3), features_rest (P,M-1,3) as the model holds them; opacity (P,1); scaling (P,2|3); rotation (P,4)."""
    xyz = np.asarray(xyz, np.float32)
    f_dc, f_rest = _channel_major(features_dc), _channel_major(features_rest)
    opacity = np.asarray(opacity, np.float32).reshape(-1, 1)
    scaling, rotation = np.asarray(scaling, np.float32), np.asarray(rotation, np.float32)
    names = vanilla_attributes(f_dc.shape[1], f_rest.shape[1], scaling.shape[1], rotation.shape[1])
    write_vertex_ply(path, names, np.concatenate([xyz, np.zeros_like(xyz), f_dc, f_rest, opacity, scaling, rotation], axis=1))


def load_vanilla_ply(path, max_sh_degree=3):
    c = read_vertex_ply(path)
    rest = _numbered(c, "f_rest_")
    if rest.shape[1] != 3 * (max_sh_degree + 1) ** 2 - 3:
        raise AssertionError(f"{path}: {rest.shape[1]} f_rest_* properties, expected {3 * (max_sh_degree + 1) ** 2 - 3}")
    return dict(xyz=np.stack([c["x"], c["y"], c["z"]], axis=1).astype(np.float32),
                features_dc=_from_channel_major(np.stack([c["f_dc_0"], c["f_dc_1"], c["f_dc_2"]], axis=1)),
                features_rest=_from_channel_major(rest), opacity=c["opacity"][:, None].astype(np.float32),
                scaling=_numbered(c, "scale_"), rotation=_numbered(c, "rot"))


# ---- scaffold / octree ------------------------------------------------------------------------------------------------
def scaffold_attributes(n_offset=30, n_feat=32, n_scale=6, n_rot=4, octree=False) -> List[str]:
    return (["x", "y", "z", "nx", "ny", "nz"] + (["level", "extra_level", "info"] if octree else []) + _names("f_offset_", n_offset)
            + _names("f_anchor_feat_", n_feat) + ["opacity"] + _names("scale_", n_scale) + _names("rot_", n_rot))


def save_scaffold_ply(path, anchor, offset, anchor_feat, opacity, scaling, rotation, level=None, extra_level=None,
                      voxel_size=None, standard_dist=None):
    """offset (P,k,3); anchor_feat (P,F); scaling (P,6).  With `level` (P,1 int) / `extra_level` (P,) / voxel_size /
    standard_dist the octree layout is written."""
    anchor = np.asarray(anchor, np.float32)
    off = _channel_major(offset)
    feat = np.asarray(anchor_feat, np.float32)
    opacity = np.asarray(opacity, np.float32).reshape(-1, 1)
    scaling, rotation = np.asarray(scaling, np.float32), np.asarray(rotation, np.float32)
    octree = level is not None
    blocks = [anchor, np.zeros_like(anchor)]
    if octree:
        lv = np.asarray(level, np.float32).reshape(-1, 1)
        info = np.zeros_like(lv)
        info[0, 0], info[1, 0] = float(voxel_size), float(standard_dist)
        blocks += [lv, np.asarray(extra_level, np.float32).reshape(-1, 1), info]
    blocks += [off, feat, opacity, scaling, rotation]
    write_vertex_ply(path, scaffold_attributes(off.shape[1], feat.shape[1], scaling.shape[1], rotation.shape[1], octree),
                     np.concatenate(blocks, axis=1))


def load_scaffold_ply(path):
    c = read_vertex_ply(path)
    out = dict(anchor=np.stack([c["x"], c["y"], c["z"]], axis=1).astype(np.float32), opacity=c["opacity"][:, None].astype(np.float32),
               scaling=_numbered(c, "scale_"), rotation=_numbered(c, "rot"), anchor_feat=_numbered(c, "f_anchor_feat"),
               offset=_from_channel_major(_numbered(c, "f_offset")))
    if "level" in c:
        out.update(level=c["level"][:, None].astype(np.int32), extra_level=c["extra_level"].astype(np.float32),
                   voxel_size=float(c["info"][0]), standard_dist=float(c["info"][1]))
    return out
