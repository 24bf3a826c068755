"""Seeded synthetic point-cloud scenes (numpy, CPU) — the workloads of BASELINE.json configs 1-5.

Town_Test.pcd (README.md:20, Task_File_VGS.txt:16 of the reference) is not distributed with the
reference, so every measured workload is synthetic: area-uniform surface samples of planes, boxes,
cylinders and noisy ellipsoids with i.i.d. Gaussian noise (sigma 5 mm) on all axes, so that no
centroid / normal component is exactly zero (the reference treats exact zeros as "empty
attribute", voxel_segmentation.h:1829,1840).  The point order is a seeded permutation: insertion
order drives PCL's dynamic bounding box.
"""
from __future__ import annotations

import numpy as np


class _Prims:
    def __init__(self):
        self.items = []  # (area, sampler(n, rng) -> (n,3) float64)

    def rect(self, o, u, v):
        o, u, v = (np.asarray(x, dtype=np.float64) for x in (o, u, v))
        area = float(np.linalg.norm(np.cross(u, v)))

        def s(n, rng):
            ab = rng.random((n, 2))
            return o + ab[:, :1] * u + ab[:, 1:] * v
        self.items.append((area, s))

    def box(self, cx, cy, z0, lx, ly, h, yaw, top=True):
        c, s_ = np.cos(yaw), np.sin(yaw)
        ex = np.array([c, s_, 0.0]) * lx
        ey = np.array([-s_, c, 0.0]) * ly
        ez = np.array([0.0, 0.0, h])
        o = np.array([cx, cy, z0]) - 0.5 * ex - 0.5 * ey
        self.rect(o, ex, ez)
        self.rect(o + ey, ex, ez)
        self.rect(o, ey, ez)
        self.rect(o + ex, ey, ez)
        if top:
            self.rect(o + ez, ex, ey)

    def cylinder(self, base, axis, r, h):
        base = np.asarray(base, dtype=np.float64)
        axis = np.asarray(axis, dtype=np.float64)
        axis = axis / np.linalg.norm(axis)
        t = np.array([1.0, 0.0, 0.0]) if abs(axis[0]) < 0.9 else np.array([0.0, 1.0, 0.0])
        e1 = np.cross(axis, t)
        e1 /= np.linalg.norm(e1)
        e2 = np.cross(axis, e1)
        area = float(2 * np.pi * r * h)

        def s(n, rng):
            th = rng.random(n) * 2 * np.pi
            tt = rng.random(n) * h
            return base + tt[:, None] * axis + r * (np.cos(th)[:, None] * e1 + np.sin(th)[:, None] * e2)
        self.items.append((area, s))

    def blob(self, c, radii, rough):
        c = np.asarray(c, dtype=np.float64)
        radii = np.asarray(radii, dtype=np.float64)
        a, b, cc = radii
        area = float(4 * np.pi * (((a * b) ** 1.6 + (a * cc) ** 1.6 + (b * cc) ** 1.6) / 3) ** (1 / 1.6))

        def s(n, rng):
            g = rng.standard_normal((n, 3))
            g /= np.linalg.norm(g, axis=1, keepdims=True)
            return c + g * radii * (1.0 + rough * rng.standard_normal((n, 1)))
        self.items.append((area, s))

    def gable_roof(self, cx, cy, z0, lx, ly, rise, yaw):
        c, s_ = np.cos(yaw), np.sin(yaw)
        ex = np.array([c, s_, 0.0]) * lx
        ey = np.array([-s_, c, 0.0])
        o = np.array([cx, cy, z0]) - 0.5 * ex - 0.5 * ly * ey
        up = np.array([0.0, 0.0, rise])
        self.rect(o, ex, 0.5 * ly * ey + up)
        self.rect(o + ly * ey, ex, -0.5 * ly * ey + up)

    def sample(self, n_points, rng, noise):
        areas = np.array([a for a, _ in self.items])
        cnt = np.floor(n_points * areas / areas.sum()).astype(np.int64)
        cnt[0] += n_points - cnt.sum()
        parts = [s(int(k), rng) for k, (_, s) in zip(cnt, self.items) if k > 0]
        # ground truth for segmentation-quality metrics: index of the surface (rectangle / cylinder / blob) a point
        # was sampled from; does not consume random numbers, so the clouds are unchanged
        self.ids = np.concatenate([np.full(int(k), i, np.int32) for i, k in enumerate(cnt) if k > 0])
        pts = np.concatenate(parts, axis=0)
        pts += noise * rng.standard_normal(pts.shape)
        return pts


def _finish(pts, rng, shuffle, prims=None, return_ids=False):
    perm = rng.permutation(pts.shape[0]) if shuffle else None
    if perm is not None:
        pts = pts[perm]
    out = np.ascontiguousarray(pts, dtype=np.float32)
    if return_ids:
        ids = prims.ids if perm is None else prims.ids[perm]
        return out, np.ascontiguousarray(ids)
    return out


def construction_site(n_points=10_000_000, seed=1, extent=70.0, noise=0.005, shuffle=True,
                      offset=(0.0, 0.0, 0.0), return_ids=False):
    """BASELINE.json config 3: slab + boxes (walls/containers, random yaw) + vertical/horizontal
    cylinders + scaffolding planes.  `extent` scales the whole site; the default 70 m with 10 M
    points gives ~1 500 pts/m^2 (~30 points per 0.15 m voxel)."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    k = extent / 70.0
    P = _Prims()
    h = extent / 2
    P.rect([-h, -h, 0.0], [extent, 0, 0], [0, extent, 0])
    for _ in range(25):
        cx, cy = rng.uniform(-h * 0.85, h * 0.85, 2)
        P.box(cx, cy, 0.0, rng.uniform(2.5, 12.0) * k, rng.uniform(2.0, 3.0) * k, rng.uniform(2.2, 3.2) * k,
              rng.uniform(0, np.pi))
    for _ in range(20):
        cx, cy = rng.uniform(-h * 0.9, h * 0.9, 2)
        P.cylinder([cx, cy, 0.0], [0, 0, 1], rng.uniform(0.15, 0.6) * k, rng.uniform(3.0, 8.0) * k)
    for _ in range(10):
        cx, cy = rng.uniform(-h * 0.9, h * 0.9, 2)
        yaw = rng.uniform(0, np.pi)
        r = rng.uniform(0.15, 0.6) * k
        P.cylinder([cx, cy, r + 0.05 * k], [np.cos(yaw), np.sin(yaw), 0], r, rng.uniform(4.0, 12.0) * k)
    for _ in range(10):
        cx, cy = rng.uniform(-h * 0.8, h * 0.8, 2)
        yaw = rng.uniform(0, np.pi)
        ex = np.array([np.cos(yaw), np.sin(yaw), 0.0]) * rng.uniform(4.0, 10.0) * k
        for lvl in range(3):
            P.rect([cx, cy, (2.0 + 2.0 * lvl) * k], ex, np.array([-np.sin(yaw), np.cos(yaw), 0.0]) * 1.2 * k)
    pts = P.sample(n_points, rng, noise) + np.asarray(offset, dtype=np.float64)
    return _finish(pts, rng, shuffle, P, return_ids)


def town(n_points=2_000_000, seed=20170610, extent=60.0, noise=0.005, shuffle=True, offset=(0.0, 0.0, 0.0)):
    """Stand-in for Town_Test.pcd (configs 1/2): ground + gabled houses + trees + poles."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    k = extent / 60.0
    P = _Prims()
    h = extent / 2
    P.rect([-h, -h, 0.0], [extent, 0, 0], [0, extent, 0])
    for _ in range(12):
        cx, cy = rng.uniform(-h * 0.8, h * 0.8, 2)
        lx, ly, hh = rng.uniform(6, 12) * k, rng.uniform(5, 8) * k, rng.uniform(3, 6) * k
        yaw = rng.uniform(0, np.pi)
        P.box(cx, cy, 0.0, lx, ly, hh, yaw, top=False)
        P.gable_roof(cx, cy, hh, lx, ly, rng.uniform(1.5, 3.0) * k, yaw)
    for _ in range(30):
        cx, cy = rng.uniform(-h * 0.9, h * 0.9, 2)
        th = rng.uniform(2.0, 4.0) * k
        P.cylinder([cx, cy, 0.0], [0, 0, 1], rng.uniform(0.1, 0.25) * k, th)
        P.blob([cx, cy, th + 1.2 * k], np.array([1.5, 1.5, 1.8]) * k * rng.uniform(0.7, 1.3), 0.12)
    for _ in range(15):
        cx, cy = rng.uniform(-h * 0.9, h * 0.9, 2)
        P.cylinder([cx, cy, 0.0], [0, 0, 1], 0.08 * k, rng.uniform(4.0, 7.0) * k)
    pts = P.sample(n_points, rng, noise) + np.asarray(offset, dtype=np.float64)
    return _finish(pts, rng, shuffle)


def urban(n_points=100_000_000, seed=2, extent=320.0, noise=0.005, shuffle=True, offset=(0.0, 0.0, 0.0)):
    """BASELINE.json config 4: Semantic3D-scale urban scene (buildings, trees, cars, furniture)."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    k = extent / 320.0
    P = _Prims()
    h = extent / 2
    P.rect([-h, -h, 0.0], [extent, 0, 0], [0, extent, 0])
    for _ in range(80):
        cx, cy = rng.uniform(-h * 0.9, h * 0.9, 2)
        lx, ly, hh = rng.uniform(8, 25) * k, rng.uniform(8, 15) * k, rng.uniform(4, 15) * k
        yaw = rng.uniform(0, np.pi)
        P.box(cx, cy, 0.0, lx, ly, hh, yaw, top=False)
        P.gable_roof(cx, cy, hh, lx, ly, rng.uniform(1.5, 4.0) * k, yaw)
    for _ in range(400):
        cx, cy = rng.uniform(-h * 0.95, h * 0.95, 2)
        th = rng.uniform(2.0, 5.0) * k
        P.cylinder([cx, cy, 0.0], [0, 0, 1], rng.uniform(0.1, 0.3) * k, th)
        P.blob([cx, cy, th + 1.5 * k], np.array([2.0, 2.0, 2.4]) * k * rng.uniform(0.6, 1.4), 0.12)
    for _ in range(300):
        cx, cy = rng.uniform(-h * 0.95, h * 0.95, 2)
        P.box(cx, cy, 0.3 * k, 4.2 * k, 1.8 * k, 1.2 * k, rng.uniform(0, np.pi))
    for _ in range(200):
        cx, cy = rng.uniform(-h * 0.95, h * 0.95, 2)
        P.cylinder([cx, cy, 0.0], [0, 0, 1], 0.08 * k, rng.uniform(3.0, 8.0) * k)
    pts = P.sample(n_points, rng, noise) + np.asarray(offset, dtype=np.float64)
    return _finish(pts, rng, shuffle)


def two_planes(n_points=40_000, seed=7, noise=0.004, shuffle=True, return_ids=False):
    """Small parity-test scene: a floor patch, a wall and a 30-degree ramp meeting it."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    P = _Prims()
    P.rect([0.3, 0.2, 0.1], [4.0, 0, 0], [0, 3.0, 0])
    P.rect([0.3, 3.2, 0.1], [4.0, 0, 0], [0, 0, 2.5])
    P.rect([4.3, 0.2, 0.1], [2.0, 0, 1.1547], [0, 3.0, 0])
    pts = P.sample(n_points, rng, noise)
    return _finish(pts, rng, shuffle, P, return_ids)


def supervoxel_labels_grid(xyz, seed_size=0.25):
    """Deterministic stand-in for pcl::SupervoxelClustering labels (SV.h:265-284): one label per
    occupied seed-resolution grid cell, labels 1..K in order of first appearance.  Used only to
    feed the SVGS post-processing identically to the oracle and to the CUDA path (the VCCS
    generator itself is third-party code with unpinned parity, SURVEY.md §0 finding 9)."""
    q = np.floor(xyz.astype(np.float64) / float(seed_size)).astype(np.int64)
    q -= q.min(axis=0)
    key = (q[:, 0] << 42) | (q[:, 1] << 21) | q[:, 2]
    _, first, inv = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(np.argsort(first))
    return (order[inv] + 1).astype(np.int32)
