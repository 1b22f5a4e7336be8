"""Seeded synthetic clouds of the shapes BASELINE.json names.

numpy on the host, so the same arrays feed the CUDA path, the CPU oracle and the CPU
baseline.  No dataset or network access anywhere.
"""
import numpy as np


def surface_clouds(B, N, seed):
    """C1 / C2: points on a noisy unit-sphere shell, scaled to max-norm 1 (NormalizeScale)."""
    rng = np.random.default_rng(seed)
    v = rng.standard_normal((B, N, 3))
    v /= np.linalg.norm(v, axis=-1, keepdims=True) + 1e-12
    v *= rng.uniform(0.9, 1.0, (B, N, 1))
    v += rng.normal(0, 0.01, (B, N, 3))
    v /= np.abs(np.linalg.norm(v, axis=-1)).max(axis=1)[:, None, None]
    return v.astype(np.float32)


def _voxel_subsample(p, size):
    key = np.floor(p / size).astype(np.int64)
    key -= key.min(0)
    dims = key.max(0) + 1
    lin = (key[:, 0] * dims[1] + key[:, 1]) * dims[2] + key[:, 2]
    _, first = np.unique(lin, return_index=True)
    return p[np.sort(first)]


def indoor_spheres(num_spheres=8, raw_points=24000, radius=0.62, grid=0.04, seed=1003):
    """C3 / C5: S3DIS-like spheres: points on 6 random planes + clutter inside a ball,
    voxel-subsampled at `grid` (the reference's first_subsampling = 0.04) and centred.  The ball
    radius is chosen so that 8 spheres hold the ~50k points per batch BASELINE.json names at
    that grid size, with realistic neighbour counts (mean ~27 of 34 at r = 0.1, ~40% of rows saturated).
    Returns pos [N,3] float32, batch [N] int64."""
    rng = np.random.default_rng(seed)
    clouds, batch = [], []
    for b in range(num_spheres):
        pts = []
        per = raw_points // 8
        for _ in range(6):
            nrm = rng.standard_normal(3)
            nrm /= np.linalg.norm(nrm)
            off = rng.uniform(-0.6, 0.6) * radius
            a = np.cross(nrm, rng.standard_normal(3))
            a /= np.linalg.norm(a)
            c = np.cross(nrm, a)
            uv = rng.uniform(-radius, radius, (per, 2))
            pts.append(off * nrm + uv[:, :1] * a + uv[:, 1:] * c + rng.normal(0, 0.005, (per, 3)))
        pts.append(rng.uniform(-radius, radius, (raw_points - 6 * per, 3)))
        p = np.concatenate(pts)
        p = p[np.linalg.norm(p, axis=1) <= radius]
        p = _voxel_subsample(p, grid)
        p = p - p.mean(0)
        clouds.append(p.astype(np.float32))
        batch.append(np.full(len(p), b, dtype=np.int64))
    return np.concatenate(clouds), np.concatenate(batch)


def lidar_scans(B=8, rings=64, azimuths=1875, seed=1004):
    """C4: SemanticKITTI-like scans: 64 rings x 1875 azimuths = 120 000 points per scan, range from
    a ground plane and a few random boxes, |xyz| <= 50 m.  Returns [B, N, 3] float32."""
    rng = np.random.default_rng(seed)
    out = np.empty((B, rings * azimuths, 3), dtype=np.float32)
    az = np.linspace(-np.pi, np.pi, azimuths, endpoint=False)
    el = np.deg2rad(np.linspace(-24.8, 2.0, rings))
    for b in range(B):
        A, E = np.meshgrid(az, el)
        d = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], -1).reshape(-1, 3)
        h = 1.73
        with np.errstate(divide="ignore"):
            r = np.where(d[:, 2] < -1e-3, h / -d[:, 2], 50.0)
        r = np.minimum(r, 50.0)
        for _ in range(12):  # boxes: ray/slab intersection
            c = np.array([rng.uniform(-30, 30), rng.uniform(-30, 30), rng.uniform(-1.5, 0.5)])
            half = rng.uniform(0.5, 4.0, 3)
            lo, hi = c - half + [0, 0, h], c + half + [0, 0, h]
            lo[2] -= h
            hi[2] -= h
            with np.errstate(divide="ignore", invalid="ignore"):
                t1, t2 = lo / d, hi / d
            tmin = np.nanmax(np.minimum(t1, t2), axis=1)
            tmax = np.nanmin(np.maximum(t1, t2), axis=1)
            hit = (tmax >= np.maximum(tmin, 0.5)) & (tmin > 0.5)
            r = np.where(hit, np.minimum(r, tmin), r)
        r = r * (1 + rng.normal(0, 0.002, r.shape))
        out[b] = (d * r[:, None]).astype(np.float32)
    return out


# PointNet++ MSG part-segmentation layer shapes (conf/models/segmentation/pointnet2.yaml:95-130)
C2_SA = [
    dict(npoint=512, radii=[0.1, 0.2, 0.4], nsample=[32, 64, 128], cin=3,
         mlps=[[6, 32, 32, 64], [6, 64, 64, 128], [6, 64, 96, 128]]),
    dict(npoint=128, radii=[0.4, 0.8], nsample=[64, 128], cin=320,
         mlps=[[323, 128, 128, 256], [323, 128, 196, 256]]),
]
# (unknown n, known m, channels) of the two three_nn / three_interpolate layers
C2_FP = [dict(n=512, m=128, C=256), dict(n=2048, m=512, C=128)]

# KPConvPaper encoder blocks after the bottleneck (conf/models/segmentation/kpconv.yaml:14-57):
# (Cin, Cout, strided, max_num_neighbors clipped to 34)
C3_LAYERS = [
    dict(level=0, cin=5, cout=64, strided=False, M=25),
    dict(level=0, cin=32, cout=32, strided=False, M=25),
    dict(level=1, cin=32, cout=32, strided=True, M=25),
    dict(level=1, cin=64, cout=64, strided=False, M=30),
    dict(level=2, cin=64, cout=64, strided=True, M=30),
    dict(level=2, cin=128, cout=128, strided=False, M=34),
    dict(level=3, cin=128, cout=128, strided=True, M=34),
    dict(level=3, cin=256, cout=256, strided=False, M=34),
    dict(level=4, cin=256, cout=256, strided=True, M=34),
    dict(level=4, cin=512, cout=512, strided=False, M=34),
]
