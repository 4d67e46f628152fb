"""numpy/ctypes front-ends of the voxelizer oracle.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

``voxelize_py``  - pure-Python loop restatement (small cases only), independent of the C file.
``voxelize_c``   - the C restatement (oracle/voxelize.c -> oracle/_build/libvoxel_oracle.so).
``collate``      - sp_voxel_preprocessor.py:145-174 (prepend agent index, concatenate agents).
PARITY UNPINNED for all of them (spconv absent) - see oracle/voxelize.c header.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libvoxel_oracle.so")


def grid_size(lidar_range, voxel_size):
    """sp_voxel_preprocessor.py:41-43 / yaml_utils.py:113-116."""
    g = (np.array(lidar_range[3:6]) - np.array(lidar_range[0:3])) / np.array(voxel_size)
    return np.round(g).astype(np.int64)


def build_c(force=False):
    src = os.path.join(_HERE, "voxelize.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _SO, src, "-lm"])
    return _SO


def voxelize_c(points, lidar_range, voxel_size, max_pts, max_voxels):
    lib = ctypes.CDLL(build_c())
    pts = np.ascontiguousarray(points, dtype=np.float32)
    P, nf = pts.shape
    g = grid_size(lidar_range, voxel_size).astype(np.int32)
    rng = np.asarray(lidar_range, dtype=np.float32)
    vs = np.asarray(voxel_size, dtype=np.float32)
    cap = int(min(max_voxels, P))
    voxels = np.empty((cap, max_pts, nf), dtype=np.float32)
    coords = np.empty((cap, 3), dtype=np.int32)
    nump = np.empty((cap,), dtype=np.int32)
    table = np.empty(int(g[0]) * int(g[1]) * int(g[2]), dtype=np.int32)
    fp = ctypes.POINTER(ctypes.c_float)
    ip = ctypes.POINTER(ctypes.c_int32)
    lib.oracle_voxelize.restype = ctypes.c_int
    m = lib.oracle_voxelize(pts.ctypes.data_as(fp), P, nf, rng.ctypes.data_as(fp), vs.ctypes.data_as(fp),
                            g.ctypes.data_as(ip), int(max_pts), int(max_voxels),
                            voxels.ctypes.data_as(fp), coords.ctypes.data_as(ip), nump.ctypes.data_as(ip),
                            table.ctypes.data_as(ip))
    return voxels[:m].copy(), coords[:m].copy(), nump[:m].copy()


def voxelize_py(points, lidar_range, voxel_size, max_pts, max_voxels):
    pts = np.asarray(points, dtype=np.float32)
    g = grid_size(lidar_range, voxel_size)
    rng = np.asarray(lidar_range, dtype=np.float32)
    vs = np.asarray(voxel_size, dtype=np.float32)
    table = {}
    voxels, coords, nump = [], [], []
    for i in range(pts.shape[0]):
        c = np.floor((pts[i, :3] - rng[:3]) / vs)          # float32 arithmetic
        if np.any(c < 0) or np.any(c >= g.astype(np.float32)):
            continue
        key = (int(c[2]), int(c[1]), int(c[0]))
        v = table.get(key)
        if v is None:
            if len(voxels) >= max_voxels:
                continue
            v = len(voxels)
            table[key] = v
            voxels.append(np.zeros((max_pts, pts.shape[1]), np.float32))
            coords.append(key)
            nump.append(0)
        if nump[v] < max_pts:
            voxels[v][nump[v]] = pts[i]
            nump[v] += 1
    if not voxels:
        return (np.zeros((0, max_pts, pts.shape[1]), np.float32), np.zeros((0, 3), np.int32),
                np.zeros((0,), np.int32))
    return np.stack(voxels), np.asarray(coords, np.int32).reshape(-1, 3), np.asarray(nump, np.int32)


def collate(per_agent):
    """per_agent: list of (voxels, coords_zyx, num_points) -> concatenated with agent index prepended."""
    vf = np.concatenate([a[0] for a in per_agent])
    vc = np.concatenate([np.pad(a[1], ((0, 0), (1, 0)), mode="constant", constant_values=i)
                         for i, a in enumerate(per_agent)]).astype(np.int32)
    vn = np.concatenate([a[2] for a in per_agent])
    return vf, vc, vn
