"""CPU tests of the host-side mirrors added for the decoder / chains / wire format: parameter names a reference
checkpoint uses, the index bookkeeping of the reference's pyramid collation (datasets/multiscale_data.py:107-165), the
gathers of BasePartialDenseConvolutionDown.forward (core/base_conv/partial_dense.py:60-81) and the call order of the
RandLA chain -- no device needed (searches are replaced by numpy restatements where a chain needs one)."""
import numpy as np
import pytest
import torch

from torch_points3d_b200 import kpconv, message_passing as mp, multiscale, partial_dense as pd, spatial_ops


def test_state_dict_keys_are_the_reference_names():
    """key names written down from the reference's module tree: UnwrappedUnetBasedModel (`down_modules`, `up_modules`),
    KPDualBlock.blocks, SimpleBlock (`kp_conv`, `bn` = FastBatchNorm1d -> `.batch_norm`), ResnetBBlock (`unary_1`,
    `unary_2`, `shortcut_op` Sequentials), FPModule_PD.nn = MLP (Sequential of Sequential(Linear, FastBatchNorm1d,
    act)), KPConvPaper.FC_layer (kpconv.py:52-71)."""
    model = kpconv.KPConvPaper(4, 13)
    keys = set(model.state_dict().keys())
    for k in ["down_modules.0.blocks.0.kp_conv.K_points", "down_modules.0.blocks.0.kp_conv.weight",
              "down_modules.0.blocks.0.bn.batch_norm.weight", "down_modules.0.blocks.0.bn.batch_norm.running_var",
              "down_modules.0.blocks.1.unary_1.0.weight", "down_modules.0.blocks.1.unary_1.1.batch_norm.bias",
              "down_modules.0.blocks.1.kp_conv.kp_conv.weight", "down_modules.0.blocks.1.kp_conv.bn.batch_norm.weight",
              "down_modules.0.blocks.1.unary_2.1.batch_norm.running_mean",
              "down_modules.0.blocks.1.shortcut_op.0.weight", "down_modules.0.blocks.1.shortcut_op.1.batch_norm.weight",
              "down_modules.4.blocks.1.kp_conv.kp_conv.K_points", "up_modules.0.nn.0.0.weight",
              "up_modules.3.nn.0.1.batch_norm.num_batches_tracked", "FC_layer.1.0.weight",
              "FC_layer.1.1.batch_norm.weight", "FC_layer.Class.weight"]:
        assert k in keys, k
    assert not any(".bn.weight" in k or "unary_1.1.weight" in k for k in keys)  # plain BatchNorm1d names
    # shapes of the yaml (in_feat = 64): last down block 1024 -> 2048 with a 512-wide KPConv, first up 3072 -> 512
    sd = model.state_dict()
    assert sd["down_modules.4.blocks.1.kp_conv.kp_conv.weight"].shape == (15, 512, 512)
    assert sd["up_modules.0.nn.0.0.weight"].shape == (512, 2048 + 1024)
    assert sd["up_modules.3.nn.0.0.weight"].shape == (64, 128 + 128)
    assert sd["down_modules.0.blocks.0.kp_conv.weight"].shape == (15, 5, 64)     # FEAT + 1 (the ones column)
    # a reference-shaped state dict loads strictly
    other = kpconv.KPConvPaper(4, 13)
    other.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    st = model.strategies()
    assert len(st["sampler"]) == len(st["neighbour_finder"]) == 10 and len(st["upsample_op"]) == 4
    assert all(u.k == 1 for u in st["upsample_op"])
    rl = mp.RandLANetRes(3, 32, [1, 1], [[10, 8, 3], [10, 16, 16]], [[6, 8, 6], [32, 64, 32]], [[6, 8, 16], [32, 64, 32]])
    rk = set(rl.state_dict().keys())
    assert "_conv.conv1._conv.point_pos_nn.0.0.weight" in rk and "_conv.conv2._conv.global_nn.1.1.batch_norm.bias" in rk
    assert "_conv.shortcut_feature_resize_nn.0.0.weight" in rk


def test_fast_batch_norm_is_the_reference_arithmetic():
    """FastBatchNorm1d._forward_sparse (base_modules.py:136-145) reshapes [N,C] -> [1,C,N] around BatchNorm1d; the direct
    2-D call gives the same numbers"""
    torch.manual_seed(0)
    bn = pd.FastBatchNorm1d(7, momentum=0.02)
    x = torch.randn(50, 7)
    ref = torch.nn.BatchNorm1d(7, momentum=0.02)
    y_ref = ref(x.unsqueeze(2).transpose(0, 2)).transpose(0, 2).squeeze(2)
    assert torch.allclose(bn(x), y_ref, atol=1e-6)
    assert torch.allclose(bn.batch_norm.running_var, ref.running_var, atol=1e-6)
    with pytest.raises(ValueError):
        bn(torch.randn(7))


def _level(pos, table, inc):
    d = pd._Data(pos=pos, idx_neighboors=table)
    d._inc = {"idx_neighboors": inc}
    return d


def test_pyramid_collation_index_rule():
    """from_data_list_token: non-negative entries of sample i move by the cumulated SUPPORT counts -- the per-sample
    shadow index included, so it aliases the next sample's first point; negative entries stay; x_idx / y_idx move
    by query / support counts; `batch` numbers the samples"""
    rng = np.random.default_rng(0)
    sup = [5, 3, 4]      # support points per sample
    qry = [2, 2, 1]      # query points per sample
    levels, ups = [], []
    for s, q in zip(sup, qry):
        t = torch.from_numpy(rng.integers(0, s + 1, (q, 3)))   # entries == s are shadows
        t[0, 0] = -1
        levels.append(_level(torch.rand(q, 3), t, s))
        u = pd._Data(x_idx=torch.from_numpy(rng.integers(0, q, (s,))), y_idx=torch.arange(s),
                     weights=torch.rand(s, 1), normalisation=torch.rand(s, 1), num_nodes=s)
        u._inc = {"x_idx": q, "y_idx": s}
        ups.append(u)
    before = [l.idx_neighboors.clone() for l in levels]
    out = multiscale.from_data_list_token(levels)
    assert all(torch.equal(a, l.idx_neighboors) for a, l in zip(before, levels))   # inputs untouched
    off = np.concatenate([[0], np.cumsum(sup)])
    want = torch.cat([torch.where(b >= 0, b + int(o), b) for b, o in zip(before, off[:-1])])
    assert torch.equal(out.idx_neighboors, want)
    assert out.batch.tolist() == [0, 0, 1, 1, 2] and out.pos.shape == (5, 3)
    # the aliasing: a shadow of sample 0 (value 5) is now index 5 = first support point of sample 1
    shadow0 = before[0] == sup[0]
    if shadow0.any():
        assert (out.idx_neighboors[:qry[0]][shadow0] == off[1]).all()
    assert (out.idx_neighboors[0, 0] == -1)
    last_shadow = before[2] == sup[2]
    assert (out.idx_neighboors[-qry[2]:][last_shadow] == sum(sup)).all()          # only the last one is the real N
    up = multiscale.from_data_list_token(ups)
    qoff = np.concatenate([[0], np.cumsum(qry)])
    assert torch.equal(up.x_idx, torch.cat([u.x_idx + int(o) for u, o in zip(ups, qoff[:-1])]))
    assert torch.equal(up.y_idx, torch.cat([u.y_idx + int(o) for u, o in zip(ups, off[:-1])]))
    assert up.num_nodes == sum(sup) and up.batch.shape[0] == sum(sup) and up.weights.shape == (sum(sup), 1)
    # MultiScaleBatch: per-scale lists collated scale by scale, plain attributes like a Batch
    samples = []
    for i, (s, q) in enumerate(zip(sup, qry)):
        d = multiscale.MultiScaleData(pos=torch.rand(s, 3), x=torch.rand(s, 2), y=torch.full((s,), i))
        d.multiscale, d.upsample = [levels[i]], [ups[i]]
        samples.append(d)
    b = multiscale.MultiScaleBatch.from_data_list(samples)
    assert b.num_scales == 1 and b.num_upsample == 1 and b.pos.shape == (sum(sup), 3)
    assert b.batch.tolist() == np.repeat(np.arange(3), sup).tolist() and b.y.tolist() == b.batch.tolist()
    assert torch.equal(b.multiscale[0].idx_neighboors, want)
    with pytest.raises(AssertionError):
        samples[1].multiscale = []
        multiscale.MultiScaleBatch.from_data_list(samples)


def _np_ball(pos, batch, r, M):
    n = pos.shape[0]
    out = np.full((n, M), n, dtype=np.int64)
    for i in range(n):
        c = 0
        for j in range(n):
            if batch[j] == batch[i] and ((pos[i] - pos[j]) ** 2).sum() <= r * r and c < M:
                out[i, c] = j
                c += 1
    return out


def test_partial_dense_conv_down_gathers():
    """BasePartialDenseConvolutionDown.forward with a numpy neighbour finder: shadow rows, centred neighbour positions,
    sampled outputs (partial_dense.py:60-81)"""
    rng = np.random.default_rng(1)
    pos = rng.random((12, 3)).astype(np.float32)
    batch = np.repeat([0, 1], 6)
    x = rng.standard_normal((12, 4)).astype(np.float32)
    seen = {}

    class Conv(pd.BasePartialDenseConvolutionDown):
        def conv(self, x, pos, x_neighbour, pos_centered_neighbour, idx_neighbour, idx_sampler):
            seen.update(x=x, pos=pos, xn=x_neighbour, pc=pos_centered_neighbour, idx=idx_neighbour)
            return x_neighbour.sum(1)[idx_sampler]

    finder = lambda a, b, batch_x, batch_y: torch.from_numpy(_np_ball(a.numpy(), batch_x.numpy(), 0.5, 5))  # noqa: E731
    sampler = lambda pos, x, batch: torch.tensor([0, 3, 7, 11])  # noqa: E731
    out = Conv(sampler, finder)(pd._Data(x=torch.from_numpy(x), pos=torch.from_numpy(pos), batch=torch.from_numpy(batch),
                                         y=torch.arange(12)))
    nb = _np_ball(pos, batch, 0.5, 5)
    xs = np.concatenate([x, np.zeros((1, 4), np.float32)])
    ps = np.concatenate([pos, np.full((1, 3), 1e6, np.float32)])
    assert seen["x"].shape == (13, 4) and seen["pos"].shape == (13, 3)
    assert np.array_equal(seen["xn"].numpy(), xs[nb])
    assert np.array_equal(seen["pc"].numpy(), ps[nb] - pos[:, None])
    assert np.allclose(out.x.numpy(), xs[nb].sum(1)[[0, 3, 7, 11]])
    assert out.pos.shape == (4, 3) and out.batch.tolist() == [0, 0, 1, 1] and out.y.shape == (12,)  # copy_from_to


def test_randla_chain_call_order_and_shapes(monkeypatch):
    """BaseConvolutionDown.forward as RandlaConv drives it (message_passing.py:44-58, RandLANet/modules.py:57-67):
    sampler -> knn(pos, pos[idx]) -> edge_index [support, query] -> kernel on (pos, pos[idx])"""
    calls = []
    n, k = 40, 4
    pos = torch.rand(n, 3)
    batch = torch.zeros(n, dtype=torch.long)

    def fake_knn(x, y, kk, bx, by, *a, **kw):
        calls.append(("knn", x.shape[0], y.shape[0], kk))
        d = torch.cdist(y, x)
        col = d.topk(kk, largest=False).indices
        row = torch.arange(y.shape[0]).view(-1, 1).expand_as(col)
        return torch.stack([row.reshape(-1), col.reshape(-1)])

    monkeypatch.setattr(spatial_ops.ops, "knn_edges", fake_knn)
    conv = mp.RandlaConv(ratio=0.5, k=k, point_pos_nn=[10, 8, 3], attention_nn=[6, 8, 6], down_conv_nn=[6, 8, 16])
    got = {}

    def fake_kernel(x, p, edge_index):
        got.update(x=x, pos=p, ei=edge_index)
        return torch.zeros(p[1].shape[0], 16)

    conv._conv.forward = fake_kernel
    out = conv(pd._Data(x=torch.rand(n, 3), pos=pos, batch=batch))
    assert calls == [("knn", n, n // 2, k)]
    assert out.idx.shape == (n // 2,) and (out.idx[1:] >= out.idx[:-1]).all()      # sorted draw: batch stays sorted
    assert got["pos"][0] is pos and torch.equal(got["pos"][1], pos[out.idx])        # (source, target) positions
    assert got["ei"].shape == (2, n // 2 * k) and int(got["ei"][0].max()) < n and int(got["ei"][1].max()) < n // 2
    assert torch.equal(out.edge_index, got["ei"]) and out.x.shape == (n // 2, 16)
    lit = mp.RandlaConv(ratio=1.0, k=k, point_pos_nn=[10, 8, 3], attention_nn=[6, 8, 6], down_conv_nn=[6, 8, 16],
                        reference_pos_order=True)
    lit._conv.forward = fake_kernel
    lit(pd._Data(x=torch.rand(n, 3), pos=pos, batch=batch))
    assert got["pos"][1] is pos                                                      # the reference's literal order


def test_linear_module_is_a_drop_in_for_nn_linear_off_the_device_path():
    """ops.Linear keeps torch.nn.Linear's parameters, keys and CPU behaviour (blocks.py:176-194 construct it as
    `Lin(in, out, bias=False)`); only [n, C] fp32 CUDA tables take the library's kernels."""
    import torch
    from torch_points3d_b200 import ops

    torch.manual_seed(0)
    ref = torch.nn.Linear(12, 7, bias=True)
    lin = ops.Linear(12, 7, bias=True)
    lin.load_state_dict(ref.state_dict())
    assert set(lin.state_dict()) == set(ref.state_dict()) == {"weight", "bias"}
    x = torch.randn(5, 12)
    assert torch.equal(lin(x), ref(x))                      # CPU tensors: the stock implementation
    x3 = torch.randn(2, 5, 12)
    assert torch.equal(lin(x3), ref(x3))                    # not a [n, C] table: the stock implementation
    assert ops.rowgemm_supported(64, 128) and not ops.rowgemm_supported(64, 13)
    assert ops.LINEAR_MIN_ROWS >= 1024


def test_sampler_constructor_precedence_is_the_reference_rule():
    """sampling.py:18-31: num_to_sample excludes the other two; else ratio wins over subsampling_param; only the
    attribute that was chosen exists afterwards (pointnet2.py tests for `_num_to_sample` with hasattr)"""
    s = spatial_ops.DenseFPSSampler(ratio=0.25, subsampling_param=7)
    assert s._ratio == 0.25 and not hasattr(s, "_subsampling_param") and not hasattr(s, "_num_to_sample")
    assert s._get_num_to_sample(10) == 2                                 # floor(10 * 0.25)
    s = spatial_ops.DenseFPSSampler(subsampling_param=7)
    assert s._subsampling_param == 7 and not hasattr(s, "_ratio")
    s = spatial_ops.DenseFPSSampler(num_to_sample=0)                     # 0 is a value, not "absent"
    assert s._get_num_to_sample(10) == 0
    for bad in ({"num_to_sample": 3, "ratio": 0.5}, {"num_to_sample": 3, "subsampling_param": 1}):
        with pytest.raises(ValueError, match="not several"):
            spatial_ops.DenseFPSSampler(**bad)
    with pytest.raises(Exception, match="should be defined"):
        spatial_ops.DenseFPSSampler()
    with pytest.raises(ValueError, match="dense data"):
        spatial_ops.DenseRandomSampler(num_to_sample=2).sample(torch.rand(8, 3))
    with pytest.raises(ValueError, match="sparse data"):
        spatial_ops.RandomSampler(num_to_sample=2).sample(torch.rand(2, 8, 3))
    draw = spatial_ops.RandomSampler(ratio=0.5, sorted_indices=True).sample(torch.rand(40, 3))
    assert draw.shape == (20,) and bool((draw[1:] >= draw[:-1]).all()) and int(draw.max()) < 40
    assert spatial_ops.DenseRandomSampler(num_to_sample=5).sample(torch.rand(2, 9, 3)).shape == (5,)


def test_multiscale_finder_broadcasting_and_repr():
    """neighbour_finder.py:110-135: scalar against list on either side, equal-length lists, two scalars"""
    M = spatial_ops.MultiscaleRadiusNeighbourFinder
    assert (M((0.1, 0.2, 0.4), 16)._radius, M((0.1, 0.2, 0.4), 16)._max_num_neighbors) == ([0.1, 0.2, 0.4], [16] * 3)
    assert (M(0.5, (4, 8))._radius, M(0.5, (4, 8))._max_num_neighbors) == ([0.5, 0.5], [4, 8])
    assert (M([0.1, 0.2], [4, 8])._radius, M([0.1, 0.2], [4, 8])._max_num_neighbors) == ([0.1, 0.2], [4, 8])
    one = M(0.3)
    assert one.num_scales == 1 and one._radius == [0.3] and one._max_num_neighbors == [64]
    with pytest.raises(ValueError, match="same length"):
        M([0.1, 0.2, 0.3], [4, 8])
    with pytest.raises(ValueError, match="out of bounds"):
        M([0.1, 0.2], 8).find_neighbours(None, None, scale_idx=2)
    assert repr(spatial_ops.KNNNeighbourFinder(3)) == "KNNNeighbourFinder {'k': 3}"


def test_dilated_knn_draws_inside_each_query_block():
    """neighbour_finder.py:50-68: query i draws k of ITS k * dilation candidates (block i of the sorted edge list)"""
    k, dilation, n = 3, 4, 5
    finder = spatial_ops.DilatedKNNNeighbourFinder(k, dilation)
    wide = k * dilation

    class Stub:
        def find_neighbours(self, x, y, batch_x, batch_y):
            row = torch.arange(n).repeat_interleave(wide)               # query index, sorted
            col = torch.arange(n * wide) + 1000                         # candidate id = 1000 + position
            return row, col

    finder.initialFinder = Stub()
    row, col = finder.find_neighbours(torch.rand(50, 3), torch.rand(n, 3), None, None)
    assert row.tolist() == np.repeat(np.arange(n), k).tolist()
    pos_in_list = col - 1000
    assert bool((pos_in_list // wide == row).all())                      # every pick comes from its own block


def test_knn_interpolate_applies_a_precomputed_table():
    """interpolate.py:36-57: y[s] = sum_e w_e x[x_idx_e] / sum_e w_e over the edges e of support point s; a table made
    for another support size is refused"""
    from types import SimpleNamespace
    g = torch.Generator().manual_seed(3)
    nq, ns, k, C = 7, 11, 3, 4
    x = torch.randn((nq, C), generator=g)
    y_idx = torch.arange(ns).repeat_interleave(k)
    x_idx = torch.randint(0, nq, (ns * k,), generator=g)
    w = torch.rand((ns * k, 1), generator=g) + 0.1
    norm = torch.zeros((ns, 1)).index_add_(0, y_idx, w)
    table = SimpleNamespace(num_nodes=ns, x_idx=x_idx, y_idx=y_idx, weights=w, normalisation=norm)
    query, support = SimpleNamespace(x=x, pos=torch.rand(nq, 3)), SimpleNamespace(x=None, pos=torch.rand(ns, 3))
    got = spatial_ops.KNNInterpolate(k)(query, support, precomputed=table)
    want = torch.stack([(x[x_idx[s * k:(s + 1) * k]] * w[s * k:(s + 1) * k]).sum(0) / w[s * k:(s + 1) * k].sum()
                        for s in range(ns)])
    assert torch.allclose(got, want, atol=1e-6)
    with pytest.raises(ValueError, match="do not match"):
        spatial_ops.KNNInterpolate(k)(query, SimpleNamespace(x=None, pos=torch.rand(ns + 1, 3)), precomputed=table)


def test_fp_module_pd_innermost_and_table_counter():
    """partial_dense.py:113-146: a coarse level of one row per cloud is broadcast by the skip level's batch vector; the
    position in the list of precomputed tables travels on the data (`up_idx`), starting at 0"""
    torch.manual_seed(0)
    fp = pd.FPModule_PD(1, [6 + 2, 5])
    fp.eval()
    skip = pd._Data(x=torch.randn(9, 2), pos=torch.rand(9, 3), batch=torch.tensor([0, 0, 0, 0, 1, 1, 1, 2, 2]))
    pooled = pd._Data(x=torch.randn(3, 6), pos=torch.zeros(3, 3), batch=torch.arange(3))
    out = fp((pooled, skip), has_innermost=True)
    want = fp.nn(torch.cat([pooled.x[skip.batch], skip.x], dim=1))
    assert torch.equal(out.x, want) and out.pos is skip.pos and not hasattr(out, "up_idx")
    assert torch.equal(fp((pooled, skip)).x, want)                       # decided from the data: 3 rows, 3 clouds

    calls = []

    class Up:
        def __call__(self, data, data_skip, precomputed=None):
            calls.append(precomputed)
            return torch.zeros(data_skip.pos.shape[0], 6)

    fp.upsample_op = Up()
    coarse = pd._Data(x=torch.randn(4, 6), pos=torch.rand(4, 3), batch=torch.tensor([0, 0, 1, 2]))
    tables = ["t0", "t1"]
    first = fp((coarse, skip), precomputed=tables, has_innermost=False)
    assert calls == ["t0"] and first.up_idx == 1
    coarse.up_idx = first.up_idx
    second = fp((coarse, skip), precomputed=tables, has_innermost=False)
    assert calls == ["t0", "t1"] and second.up_idx == 2
    fp((coarse, skip), has_innermost=False)                              # no tables: the search runs on the fly
    assert calls[-1] is None


def test_mlp_and_fast_batch_norm_against_the_reference_classes():
    """tests/golden/base_modules.npz holds outputs of the REFERENCE's own MLP / FastBatchNorm1d
    (core/common_modules/base_modules.py:35-45,128-153, run by tests/golden/make_golden_modules.py): same state_dict
    keys in the same order, a reference state_dict loads strictly, and two training steps, the running statistics,
    the input gradient and the eval output agree to 1e-6 (fp32, same arithmetic up to the [N,C] vs [1,C,N] layout)"""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "base_modules.npz"))
    mlp = pd.MLP([7, 12, 5], bn_momentum=0.2, bias=False)
    assert list(mlp.state_dict().keys()) == [str(k) for k in g["mlp_keys"]]
    mlp.load_state_dict({str(k): torch.from_numpy(g["mlp_init/" + str(k)]) for k in g["mlp_keys"]}, strict=True)
    x = torch.from_numpy(g["mlp_x"]).requires_grad_(True)
    mlp.train()
    y1 = mlp(x)
    y2 = mlp(x * 1.5 + 0.25)
    (gx,) = torch.autograd.grad(y2, x, torch.from_numpy(g["mlp_gy"]))
    assert np.allclose(y1.detach().numpy(), g["mlp_y1"], atol=1e-6)
    assert np.allclose(y2.detach().numpy(), g["mlp_y2"], atol=1e-6)
    assert np.allclose(gx.numpy(), g["mlp_gx"], atol=1e-6)
    for k, v in mlp.state_dict().items():
        assert np.allclose(v.numpy(), g["mlp_after/" + k], atol=1e-6), k
    mlp.eval()
    assert np.allclose(mlp(x).detach().numpy(), g["mlp_eval"], atol=1e-6)

    bn = pd.FastBatchNorm1d(6, momentum=0.05)
    assert list(bn.state_dict().keys()) == [str(k) for k in g["bn_keys"]]
    with torch.no_grad():
        bn.batch_norm.weight.copy_(torch.from_numpy(g["bn_weight"]))
        bn.batch_norm.bias.copy_(torch.from_numpy(g["bn_bias"]))
    bn.train()
    assert np.allclose(bn(torch.from_numpy(g["bn_x"])).detach().numpy(), g["bn_y"], atol=1e-6)   # dense [B, N, C]
    assert np.allclose(bn.batch_norm.running_mean.numpy(), g["bn_running_mean"], atol=1e-7)
    assert np.allclose(bn.batch_norm.running_var.numpy(), g["bn_running_var"], atol=1e-7)
    with pytest.raises(ValueError) as err:
        bn(torch.zeros(2, 2, 2, 2))
    assert str(err.value) == str(g["bn_error"])


def test_compat_cluster_serves_the_reference_import_lines(monkeypatch):
    """`from torch_geometric.nn import (knn_interpolate, fps, radius, global_max_pool, global_mean_pool, knn)`
    (core/base_conv/message_passing.py:10-17, partial_dense.py:10-17) and `PointConv`
    (modules/pointnet2/message_passing.py:1) resolve against compat.cluster; the pools are the per-cloud reductions."""
    from torch_points3d_b200 import ops
    from torch_points3d_b200.compat import cluster
    from torch_points3d_b200.compat.cluster import (PointConv, fps, global_max_pool, global_mean_pool, knn,  # noqa: F401
                                                   knn_interpolate, radius)
    assert PointConv is mp.PointConv and knn is ops.knn_edges and set(cluster.__all__) <= set(dir(cluster)) | {"PointConv"}

    def segment_max(msg, dst, T):
        out = torch.zeros((T, msg.shape[1]))
        for t in torch.unique(dst).tolist():
            out[t] = msg[dst == t].max(dim=0).values
        return out

    monkeypatch.setattr(ops, "segment_max", segment_max)
    monkeypatch.setattr(ops, "segment_sum", lambda msg, dst, T: torch.zeros((T, msg.shape[1])).index_add_(0, dst, msg))
    x = torch.randn(9, 4)
    batch = torch.tensor([0, 0, 0, 1, 1, 3, 3, 3, 3])           # cloud 2 is empty
    want_max = torch.stack([x[batch == b].max(0).values if (batch == b).any() else torch.zeros(4) for b in range(4)])
    want_mean = torch.stack([x[batch == b].mean(0) if (batch == b).any() else torch.zeros(4) for b in range(4)])
    assert torch.equal(global_max_pool(x, batch), want_max)
    assert torch.allclose(global_mean_pool(x, batch, size=4), want_mean, atol=1e-6)
    with pytest.raises(AttributeError):
        cluster.voxel_grid
