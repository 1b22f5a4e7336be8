"""GridSampling3D: the voxel-clustering kernels against the numpy oracle
(bit-exact cluster ids, representative indices, fp32 means) and against what the reference's own tests hold
(test/test_grid_sampling.py: majority label, idempotence of a second pass, int32 quantised coordinates;
test/test_transform.py:67-93: MultiScaleTransform consistency is covered through the strided KPConv blocks)."""
import numpy as np
import pytest
import torch

import oracle

DEV = "cuda:0"


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def N(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------- oracle (CPU) ---------
def test_oracle_reference_label_vote_case():
    """test/test_grid_sampling.py:18-36: five points inside one 0.04 voxel -> the majority label"""
    pos = np.array([[0, 0, 0.01], [0.01, 0, 0], [0, 0.01, 0], [0, 0.01, 0], [0.01, 0, 0.01]], np.float32)
    for seed in range(8):
        y = np.random.default_rng(seed).integers(0, 2, 5)
        uniq, counts = np.unique(y, return_counts=True)
        cluster, perm, _ = oracle.grid_cluster(pos, 0.04, np.zeros(5, np.int64))
        assert cluster.tolist() == [0] * 5 and perm.tolist() == [4]
        assert oracle.label_vote(y, cluster, 1)[0] == uniq[np.argmax(counts)]


def test_oracle_matches_torch_op_sequence():
    """the op-for-op torch sequence of grid_cluster / consecutive_cluster / scatter_mean on CPU tensors"""
    torch.manual_seed(0)
    pos = torch.randn(4000, 3) * 0.3
    batch = torch.sort(torch.randint(0, 4, (4000,))).values
    coords = torch.round(pos / 0.04)
    full = torch.cat([coords, batch[:, None].type_as(coords)], 1)
    start, end = full.min(0).values, full.max(0).values
    nv = ((end - start) / torch.ones(4)).long() + 1
    nv = torch.cat([torch.ones(1, dtype=torch.long), nv.cumprod(0)])[:4]
    key = (((full - start[None]) / torch.ones(1, 4)).long() * nv[None]).sum(1)
    uniq, inv = torch.unique(key, sorted=True, return_inverse=True)
    perm = inv.new_empty(uniq.size(0)).scatter_(0, inv, torch.arange(inv.size(0)))
    mean = torch.zeros(uniq.size(0), 3).index_add_(0, inv, pos) / torch.bincount(inv).float()[:, None]
    c, pm, _ = oracle.grid_cluster(pos.numpy(), 0.04, batch.numpy())
    assert np.array_equal(c, inv.numpy()) and np.array_equal(pm, perm.numpy())
    assert np.array_equal(oracle.scatter_mean(pos.numpy(), c, len(pm)), mean.numpy())


def test_oracle_idempotent_second_pass():
    """test/test_grid_sampling.py:38-58"""
    rng = np.random.default_rng(5)
    pos = (rng.standard_normal((1000, 3)) * 0.1).astype(np.float32)
    counts = []
    for _ in range(3):
        c, pm, _ = oracle.grid_cluster(pos, 0.02)
        pos = oracle.scatter_mean(pos, c, len(pm))
        counts.append(len(pm))
    assert counts[0] == counts[1] == counts[2] < 1000


# --------------------------------------------------------------------- kernels (GPU) ---------
def _cloud(n, clouds, spread, seed):
    rng = np.random.default_rng(seed)
    pos = (rng.standard_normal((n, 3)) * spread).astype(np.float32)
    batch = np.sort(rng.integers(0, clouds, n)).astype(np.int64)
    return pos, batch


@pytest.mark.gpu
@pytest.mark.parametrize("n,clouds,spread,size", [(5, 1, 0.01, 0.04), (1000, 1, 0.1, 0.02), (50000, 8, 0.6, 0.04),
                                                  (3000, 3, 30.0, 0.05), (70001, 5, 1.0, 0.5), (2048, 2, 1.0, 100.0),
                                                  (400000, 8, 2.0, 0.06)])
def test_grid_cluster_bit_exact(n, clouds, spread, size):
    from torch_points3d_b200 import ops
    pos, batch = _cloud(n, clouds, spread, 100 + n)
    want_c, want_perm, want_coords = oracle.grid_cluster(pos, size, batch)
    vc = ops.grid_cluster(T(pos), size, T(batch), want_mean=True, want_coords=True)
    assert vc.n == len(want_perm)
    assert np.array_equal(N(vc.cluster), want_c)
    assert np.array_equal(N(vc.unique_pos_indices), want_perm)
    assert np.array_equal(N(vc.mean_pos), oracle.scatter_mean(pos, want_c, vc.n))
    assert np.array_equal(N(vc.batch), batch[want_perm])
    assert vc.coords.dtype == torch.int32 and np.array_equal(N(vc.coords), want_coords[want_perm].astype(np.int32))
    # segment list: points sorted by (voxel, index)
    order, seg = N(vc.order).astype(np.int64), N(vc.seg_start)
    assert np.array_equal(order, np.lexsort((np.arange(n), want_c)))
    assert seg[0] == 0 and seg[-1] == n and np.array_equal(np.diff(seg), np.bincount(want_c, minlength=vc.n))
    # features and labels
    x = np.random.default_rng(1).standard_normal((n, 7)).astype(np.float32)
    assert np.array_equal(N(ops.segment_mean(T(x), vc)), oracle.scatter_mean(x, want_c, vc.n))
    y = np.random.default_rng(2).integers(-3, 9, n)
    assert np.array_equal(N(ops.segment_vote(T(y), vc)), oracle.label_vote(y, want_c, vc.n))


@pytest.mark.gpu
def test_grid_cluster_without_batch_and_edge_cases():
    from torch_points3d_b200 import ops
    pos, _ = _cloud(9000, 1, 0.5, 7)
    want_c, want_perm, _ = oracle.grid_cluster(pos, 0.03)
    vc = ops.grid_cluster(T(pos), 0.03)
    assert np.array_equal(N(vc.cluster), want_c) and np.array_equal(N(vc.unique_pos_indices), want_perm)
    assert vc.batch is None
    # half-way coordinates round to even, negative zero and positive zero share a voxel
    pos = np.array([[0.5, 0.5, 0.5], [1.5, 1.5, 1.5], [-0.2, 0.2, 0.0], [2.5, -0.5, -1.5], [0.49, 2.0, 2.0]], np.float32)
    want_c, want_perm, _ = oracle.grid_cluster(pos, 1.0)
    vc = ops.grid_cluster(T(pos), 1.0)
    assert np.array_equal(N(vc.cluster), want_c) and np.array_equal(N(vc.unique_pos_indices), want_perm)
    # empty input and a single voxel
    vc = ops.grid_cluster(torch.zeros(0, 3, device=DEV), 0.1, torch.zeros(0, dtype=torch.int64, device=DEV))
    assert vc.n == 0 and vc.cluster.numel() == 0
    vc = ops.grid_cluster(torch.rand(3000, 3, device=DEV) * 0.01 + 0.02, 1.0)
    assert vc.n == 1 and int(vc.unique_pos_indices[0]) == 2999
    with pytest.raises(RuntimeError):
        ops.grid_cluster(torch.zeros(4, 3), 0.1)  # no CPU fallback
    with pytest.raises(ValueError):
        ops.grid_cluster(torch.zeros(4, 2, device=DEV), 0.1)


@pytest.mark.gpu
def test_grid_sampling_3d_reference_cases():
    """the three cases of the reference's test/test_grid_sampling.py through the mirrored transform"""
    from torch_points3d_b200.grid_transform import Data, GridSampling3D
    pos = T(np.array([[0, 0, 0.01], [0.01, 0, 0], [0, 0.01, 0], [0, 0.01, 0], [0.01, 0, 0.01]], np.float32))
    for seed in range(4):
        y = np.random.default_rng(seed).integers(0, 2, 5)
        uniq, counts = np.unique(y, return_counts=True)
        out = GridSampling3D(0.04)(Data(pos=pos, batch=torch.zeros(5, dtype=torch.int64, device=DEV), y=T(y)))
        assert N(out.y).tolist() == [uniq[np.argmax(counts)]] and out.pos.shape == (1, 3)
    # test_double_grid_sampling
    torch.manual_seed(1)
    u = Data(pos=torch.randn(1000, 3, device=DEV) * 0.1, x=torch.ones(1000, 1, device=DEV))
    gr, shapes = GridSampling3D(0.02), []
    for _ in range(3):
        u = gr(u)
        shapes.append(u.pos.shape[0])
        assert u.x.shape == (u.pos.shape[0], 1) and torch.equal(u.x, torch.ones_like(u.x))
    assert shapes[0] == shapes[1] == shapes[2] < 1000
    # test_quantize
    q = GridSampling3D(0.2, quantize_coords=True)(Data(pos=torch.randn(100, 3, device=DEV) * 0.1,
                                                       x=torch.ones(100, 1, device=DEV)))
    assert q.coords.dtype == torch.int and q.x.shape[0] == q.coords.shape[0] == q.num_nodes
    assert len(np.unique(N(q.coords), axis=0)) == q.num_nodes
    # mode="last": one existing point per voxel
    d = Data(pos=torch.randn(500, 3, device=DEV) * 0.1, origin_id=torch.arange(500, device=DEV))
    src = d.pos.clone()
    out = GridSampling3D(0.05, mode="last")(d)
    assert torch.equal(out.pos, src[out.origin_id]) and out.pos.shape[0] == GridSampling3D(0.05)(Data(pos=src)).pos.shape[0]
