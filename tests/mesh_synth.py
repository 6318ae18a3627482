"""Seeded TSDF-like lattices for the marching-cubes tests (oracle/mcubes_oracle.py, gs-sr_b200/csrc/mcubes.cu)."""
import numpy as np

F = np.float32


def lattice_coords(nz, ny, nx):
    z, y, x = np.meshgrid(np.arange(nz, dtype=F), np.arange(ny, dtype=F), np.arange(nx, dtype=F), indexing="ij")
    return x, y, z


def sphere(n=40, r=12.3):
    x, y, z = lattice_coords(n, n, n)
    c = F((n - 1) / 2)
    return (np.sqrt((x - c) ** 2 + (y - c) ** 2 + (z - c) ** 2) - F(r)).astype(F)


def torus(nz=30, ny=48, nx=52, R=14.2, r=5.1):
    x, y, z = lattice_coords(nz, ny, nx)
    cx, cy, cz = F((nx - 1) / 2), F((ny - 1) / 2), F((nz - 1) / 2)
    q = np.sqrt((x - cx) ** 2 + (y - cy) ** 2) - F(R)
    return (np.sqrt(q * q + (z - cz) ** 2) - F(r)).astype(F)


def noise(shape, seed):
    return np.random.default_rng(seed).standard_normal(shape).astype(F)


def observed_blob(shape, seed):
    """A noisy field with TSDF-fusion style weights: 1 = never observed, > 1 observed, in irregular patches."""
    rng = np.random.default_rng(seed)
    f = noise(shape, seed + 1)
    w = np.where(rng.random(shape) < 0.8, rng.integers(2, 9, size=shape), 1).astype(F)
    rgb = rng.random(shape + (3,)).astype(F)
    return f, w, rgb


def is_lattice_boundary_vertex(v, origin, voxel, dims, eps=1e-4):
    """Vertices lying on the outer faces of the lattice box (open boundary of an unclosed surface)."""
    nx, ny, nz = dims
    g = (np.asarray(v, dtype=np.float64) - np.asarray(origin, dtype=np.float64)) / float(voxel)
    hi = np.array([nx - 1, ny - 1, nz - 1], dtype=np.float64)
    return ((np.abs(g) < eps) | (np.abs(g - hi) < eps)).any(axis=1)
