"""GPU tests of the callers either side of the operators: the KPConv decoder
(FPModule_PD) and the whole KPConvPaper segmentation network, the PARTIAL_DENSE and MESSAGE_PASSING down-convolution
chains (BasePartialDenseConvolutionDown, RandlaConv / DilatedResidualBlock), the reference-format pyramid, and the C5
data-parallel training step on two ranks.  Operators are checked against the CPU oracle composed the way the
reference composes them; MLPs / BatchNorm are the same stock torch modules on both sides."""
import os
import socket

import numpy as np
import pytest
import torch

import oracle
from torch_points3d_b200 import _cabi, kpconv, message_passing as mp, multiscale, partial_dense as pd, spatial_ops, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def N(t):
    return t.detach().cpu().numpy()


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.fixture(scope="module", autouse=True)
def _lib_loaded():
    import __graft_entry__
    __graft_entry__.build()
    assert torch.cuda.is_available()
    before = _cabi.launch_count()
    yield
    assert _cabi.launch_count() > before, "no kernel of libp3d_b200.so was launched"


def _spheres(num=2, raw=6000, radius=0.45, seed=93):
    pos, batch = synthetic.indoor_spheres(num_spheres=num, raw_points=raw, radius=radius, seed=seed)
    return pos, batch


def test_fp_module_pd_against_composed_oracle():
    """FPModule_PD (core/base_conv/partial_dense.py:103-146) with up_k = 1 and 3: knn of every skip point among the
    coarse points (oracle.knn_csr), w = 1 / clamp(d2, 1e-16), sum(w x) / sum(w) (interpolate.py:25-30,55-57), concat
    skip, MLP (eval-mode BN)"""
    torch.manual_seed(0)
    pos_h, batch_h = _spheres()
    pos, batch = T(pos_h), T(batch_h)
    q_pos, q_batch = kpconv.grid_subsample(pos, batch, 0.12)
    nq = q_pos.shape[0]
    assert 50 < nq < pos.shape[0] // 4
    x_coarse = torch.randn(nq, 24, device=DEV)
    x_skip = torch.randn(pos.shape[0], 8, device=DEV)
    ptr_x, ptr_y = oracle.batch_to_ptr(N(q_batch)), oracle.batch_to_ptr(batch_h)
    for k in (1, 3):
        fp = pd.FPModule_PD(k, [32, 16], bn_momentum=0.02).to(DEV).eval()
        coarse = kpconv.Cloud(q_pos, x_coarse, q_batch)
        skip = kpconv.Cloud(pos, x_skip, batch)
        with torch.no_grad():
            out = fp((coarse, skip))
            col, d2 = oracle.knn_csr(N(q_pos), pos_h, k, ptr_x, ptr_y)
            w = 1.0 / np.maximum(d2, 1e-16)
            interp = (N(x_coarse)[col] * w[..., None]).sum(1) / w.sum(1, keepdims=True)
            want = fp.nn(torch.cat([T(interp.astype(np.float32)), x_skip], 1))
        assert out.pos is pos and out.x.shape == (pos.shape[0], 16)
        assert relerr(N(out.x), N(want)) <= 2e-5
        # the precomputed table (MultiScaleTransform.upsample -> KNNInterpolate.precompute) gives the same features
        pre = fp.upsample_op.precompute(coarse, skip)
        with torch.no_grad():
            out2 = fp((coarse, skip), precomputed=[pre], has_innermost=False)
        assert relerr(N(out2.x), N(want)) <= 2e-5 and out2.up_idx == 1
    # innermost case: one pooled row per cloud is broadcast by the batch vector (partial_dense.py:135-136)
    pooled = kpconv.Cloud(torch.zeros(2, 3, device=DEV), torch.randn(2, 24, device=DEV), torch.arange(2, device=DEV))
    with torch.no_grad():
        out = fp((pooled, skip))
        want = fp.nn(torch.cat([pooled.x[batch], x_skip], 1))
    assert torch.equal(out.x, want)


def test_kpconv_paper_network_trains_and_pyramid_paths_agree():
    """KPConvPaper (models/segmentation/kpconv.py:106-140 over kpconv.yaml): per-point log-probabilities, NLL loss
    decreases under SGD, the precomputed pyramid (down tables + upsample tables) reproduces the on-the-fly path"""
    torch.manual_seed(1)
    pos_h, batch_h = _spheres(num=2, raw=24000, radius=0.62, seed=94)
    pos, batch = T(pos_h), T(batch_h)
    n = pos.shape[0]
    model = kpconv.KPConvPaper(4, 13, max_nn_cap=34, random_rotation=False).to(DEV)
    feats = torch.randn(n, 4, device=DEV)
    labels = (torch.arange(n, device=DEV) * 7 // 3) % 13
    out = model(pos, feats, batch)
    assert out.shape == (n, 13) and torch.isfinite(out).all()
    assert torch.allclose(out.exp().sum(1), torch.ones(n, device=DEV), atol=1e-4)   # LogSoftmax head
    tr = kpconv.MultiScaleTransform(model.strategies())
    ms = tr(pos, batch)
    up = tr.upsample(pos, batch, ms)
    assert len(ms) == 10 and len(up) == 4
    model.eval()
    with torch.no_grad():
        a = model(pos, feats, batch)
        b = model(pos, feats, batch, multiscale=ms, upsample=up)
    # the encoder half is bit-identical (test_multiscale_transform_precomputed_path_equals_on_the_fly); the decoder's
    # two routes weight with 1/d2 from different fp32 expressions (kernel d2 vs the (diff*diff).sum of
    # interpolate.py:27-28), so (x*w)/w rounds differently in the last bit -- in the reference as well
    assert relerr(N(a), N(b)) <= 1e-5
    model.train()
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9)
    losses = []
    for _ in range(6):
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.nll_loss(model(pos, feats, batch, multiscale=ms, upsample=up), labels)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]
    missing = [k for k, p in model.named_parameters() if p.requires_grad and p.grad is None]
    assert not missing, missing


def test_reference_format_pyramid_feeds_the_blocks():
    """A pyramid built the reference's way -- per sample (MultiScaleTransform, transforms.py:579-651), then collated
    (MultiScaleBatch.from_data_list, multiscale_data.py:56-165) -- against the whole-batch tables: identical except that
    sample i's shadow entries carry the collation's offset (the first point of sample i+1, Appendix C.4); with the
    shadows restored the network output equals the on-the-fly path bit for bit, and as-is the blocks run on it."""
    torch.manual_seed(2)
    pos_h, batch_h = _spheres(num=3, raw=5000, radius=0.4, seed=95)
    pos, batch = T(pos_h), T(batch_h)
    model = kpconv.KPConvPaper(3, 5, in_feat=32, levels=3, max_nn_cap=34, random_rotation=False).to(DEV).eval()
    st = model.strategies()
    per_sample = multiscale.MultiScaleTransform(st)
    samples = []
    for b in range(3):
        sel = batch == b
        samples.append(per_sample(multiscale.MultiScaleData(pos=pos[sel], x=torch.randn(int(sel.sum()), 3, device=DEV))))
    mb = multiscale.MultiScaleBatch.from_data_list(samples)
    assert torch.equal(mb.pos, pos) and torch.equal(mb.batch, batch)
    assert mb.num_scales == 6 and mb.num_upsample == 2
    whole = kpconv.MultiScaleTransform(st)
    ms = whole(pos, batch)
    up = whole.upsample(pos, batch, ms)
    sup_counts = torch.bincount(batch)
    repaired = []
    for lvl, (ref_lvl, ours) in enumerate(zip(mb.multiscale, ms)):
        assert torch.equal(ref_lvl.pos, ours.pos) and torch.equal(ref_lvl.batch, ours.batch)
        n_sup = int(sup_counts.sum())
        ends = torch.cumsum(sup_counts, 0)                       # per-sample shadow value after collation
        shadow_val = ends[ref_lvl.batch].unsqueeze(1)
        is_shadow = ours.idx_neighboors == n_sup
        assert torch.equal(torch.where(is_shadow, shadow_val.expand_as(is_shadow), ours.idx_neighboors),
                           ref_lvl.idx_neighboors)
        fixed = pd._Data(pos=ref_lvl.pos, batch=ref_lvl.batch,
                         idx_neighboors=torch.where(is_shadow, torch.full_like(ours.idx_neighboors, n_sup),
                                                    ref_lvl.idx_neighboors))
        repaired.append(fixed)
        if st["sampler"][lvl] is not None:
            sup_counts = torch.bincount(ours.batch)
    for r, o in zip(mb.upsample, up):
        assert torch.equal(r.x_idx, o.x_idx) and torch.equal(r.y_idx, o.y_idx) and r.num_nodes == o.num_nodes
        assert torch.equal(r.weights, o.weights) and torch.equal(r.normalisation, o.normalisation)
    x = mb.x
    with torch.no_grad():
        fly = model(pos, x, batch)
        rep = model(pos, x, batch, multiscale=repaired, upsample=mb.upsample)
        raw = model(pos, x, batch, multiscale=mb.multiscale, upsample=mb.upsample)
    assert relerr(N(fly), N(rep)) <= 1e-5
    # encoder only (no interpolation arithmetic in between): bit for bit
    d_fly = kpconv.Cloud(pos, kpconv.add_ones(pos, x, True), batch)
    d_rep = d_fly.clone()
    with torch.no_grad():
        for m in model.down_modules:
            d_fly, d_rep = m(d_fly), m(d_rep, precomputed=repaired)
    assert torch.equal(d_fly.x, d_rep.x)
    assert torch.isfinite(raw).all() and raw.shape == fly.shape


def test_partial_dense_conv_down_chain():
    """BasePartialDenseConvolutionDown.forward (partial_dense.py:60-81) on the device: FPS sampler, self ball query,
    shadow-padded gathers -- against the oracle's tables and numpy gathers"""
    pos_h, batch_h = _spheres(num=2, raw=3000, radius=0.3, seed=96)
    rng = np.random.default_rng(3)
    x_h = rng.standard_normal((pos_h.shape[0], 6)).astype(np.float32)
    seen = {}

    class Conv(pd.BasePartialDenseConvolutionDown):
        def conv(self, x, pos, x_neighbour, pos_centered_neighbour, idx_neighbour, idx_sampler):
            seen.update(xn=x_neighbour, pc=pos_centered_neighbour, idx=idx_neighbour)
            return (x_neighbour * torch.exp(-pos_centered_neighbour.square().sum(-1, keepdim=True))).sum(1)[idx_sampler]

    conv = Conv(spatial_ops.FPSSampler(ratio=0.25, random_start=False),
                spatial_ops.RadiusNeighbourFinder(0.1, 20, conv_type="partial_dense"))
    out = conv(pd._Data(x=T(x_h), pos=T(pos_h), batch=T(batch_h)))
    ptr = oracle.batch_to_ptr(batch_h)
    nb = oracle.ball_query_csr(0.1, 20, pos_h, pos_h, ptr, ptr)[0]
    idx = oracle.fps_csr(pos_h, ptr, oracle.fps_counts(ptr, 0.25))
    assert np.array_equal(N(seen["idx"]), nb)
    xs = np.concatenate([x_h, np.zeros((1, 6), np.float32)])
    ps = np.concatenate([pos_h, np.full((1, 3), 1e6, np.float32)])
    assert np.array_equal(N(seen["xn"]), xs[nb]) and np.array_equal(N(seen["pc"]), ps[nb] - pos_h[:, None])
    assert np.array_equal(N(out.pos), pos_h[idx]) and np.array_equal(N(out.batch), batch_h[idx])
    assert out.x.shape == (idx.shape[0], 6)


def test_randla_conv_chain_against_composed_oracle():
    """RandlaConv = BaseConvolutionDown.forward + RandlaKernel (RandLANet/modules.py:9-67): sampler -> knn k = 16 ->
    [pos_i, pos_j, pos_i - pos_j, |.|] encoding -> MLPs -> attentive sum per query; oracle: knn_edges, edge_encoding,
    segment_sum with the same torch MLPs (eval-mode BN).  Then the residual block and FPModule on top."""
    torch.manual_seed(4)
    pos_h, batch_h = _spheres(num=2, raw=4000, radius=0.35, seed=97)
    pos, batch = T(pos_h), T(batch_h)
    n = pos.shape[0]
    x = torch.randn(n, 8, device=DEV)
    conv = mp.RandlaConv(ratio=0.25, k=16, sampler=spatial_ops.FPSSampler(ratio=0.25, random_start=False),
                         point_pos_nn=[10, 8, 8], attention_nn=[16, 32, 16], down_conv_nn=[16, 32, 24]).to(DEV).eval()
    with torch.no_grad():
        out = conv(pd._Data(x=x, pos=pos, batch=batch))
    ptr = oracle.batch_to_ptr(batch_h)
    idx = oracle.fps_csr(pos_h, ptr, oracle.fps_counts(ptr, 0.25))
    assert np.array_equal(N(out.idx), idx)
    q, qb = pos_h[idx], batch_h[idx]
    row, col = oracle.knn_edges(pos_h, q, 16, ptr, oracle.batch_to_ptr(qb))
    assert np.array_equal(N(out.edge_index), np.stack([col, row]))
    k = conv._conv
    with torch.no_grad():
        rel = T(oracle.edge_encoding(pos_h, q, col, row))
        fij = torch.cat([x[T(col)], k.point_pos_nn(rel)], 1)
        msg = torch.softmax(k.attention_nn(fij), -1) * fij
        want = k.global_nn(T(oracle.segment_sum(N(msg), row, idx.shape[0])))
    assert relerr(N(out.x), N(want)) <= 2e-5
    # random sampler (the reference's default): sorted draw, every query finds k neighbours inside its own cloud
    rc = mp.RandlaConv(ratio=0.5, k=16, point_pos_nn=[10, 8, 8], attention_nn=[16, 32, 16],
                       down_conv_nn=[16, 32, 24]).to(DEV).eval()
    with torch.no_grad():
        o2 = rc(pd._Data(x=x, pos=pos, batch=batch))
    src, dst = o2.edge_index
    assert o2.idx.shape[0] == n // 2 and bool((batch[src] == o2.batch[dst]).all()) and src.shape[0] == 16 * (n // 2)
    # residual block of two convs + upsampling back (FPModule, knn_interpolate k = 1)
    blk = mp.RandLANetRes(8, 32, [0.5, 0.5], [[10, 8, 8], [10, 16, 16]], [[16, 32, 16], [32, 64, 32]],
                          [[16, 32, 16], [32, 64, 32]]).to(DEV).eval()
    with torch.no_grad():
        down = blk(pd._Data(x=x, pos=pos, batch=batch))
        assert down.x.shape == (n // 4, 32) and torch.isfinite(down.x).all()
        fp = mp.FPModule(1, [32 + 8, 16]).to(DEV).eval()
        upo = fp((down, pd._Data(x=x, pos=pos, batch=batch)))
        col1, _ = oracle.knn_csr(N(down.pos), pos_h, 1, oracle.batch_to_ptr(N(down.batch)), ptr)
        want = fp.nn(torch.cat([down.x[T(col1[:, 0])], x], 1))
    assert relerr(N(upo.x), N(want)) <= 2e-5


# ------------------------------------------------------------------------------------ C5: two ranks ----------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ddp_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)  # both ranks share cuda:0 here: NCCL needs one GPU per rank
    try:
        from torch_points3d_b200 import segmentation
        torch.manual_seed(11)                                   # same initial weights on both ranks
        trainer = segmentation.KPConvSegmentationTrainer(feat_dim=4, num_classes=13, device=torch.device(DEV), in_feat=32,
                                                         levels=3, ddp=True, seed=200 + rank, points_raw=6000, radius=0.4)
        losses = [trainer.step() for _ in range(4)]
        flat = torch.cat([p.detach().reshape(-1) for p in trainer.model.parameters()])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        out[rank] = dict(losses=losses, equal=bool(all(torch.equal(gathered[0], g) for g in gathered)),
                         n=int(trainer.pos.shape[0]))
    finally:
        dist.destroy_process_group()


def test_c5_ddp_two_ranks_parameters_stay_equal_and_loss_decreases():
    """the C5 training step (KPConvPaper + NLL + SGD under DistributedDataParallel) on two ranks with different
    spheres: after every step all ranks hold bit-identical parameters (the gradient all-reduce happened) and the
    loss goes down"""
    import torch.multiprocessing as tmp
    ctx = tmp.get_context("spawn")
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert out[0]["equal"] and out[1]["equal"]
    assert out[0]["n"] != out[1]["n"]                      # different data per rank
    for r in range(2):
        ls = out[r]["losses"]
        assert all(np.isfinite(ls)) and ls[-1] < ls[0], ls
