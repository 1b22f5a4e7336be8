"""The KPConv segmentation training step of BASELINE.json configs[4] ("C5"): KPConvPaper
(models/segmentation/kpconv.py:106-140 over conf/models/segmentation/kpconv.yaml) on synthetic S3DIS-shaped spheres,
negative log-likelihood on per-point labels (kpconv.py:160-166), SGD(momentum 0.98, lr 0.01) as
conf/training/kpconv.yaml:8-16, one process per GPU with torch's DistributedDataParallel (one NCCL gradient all-reduce
per step; BatchNorm statistics stay per rank, the reference has no SyncBN).

The geometry of a batch (grid subsampling, neighbour tables, upsampling tables) depends on positions only.  It is
built once per batch on the device by kpconv.MultiScaleTransform -- the reference builds the same pyramid in its
DataLoader workers (core/data_transform/transforms.py:579-651) -- and the network step reads it through
`precomputed=`.  With the pyramid given every shape of the step is static, so forward + backward + optimizer can be
replayed from one CUDA graph (`capture()`); under DDP the gradient all-reduce is captured with it (NCCL supports
stream capture), which removes the per-step launch and hook overhead the eager DDP step pays.
"""
import numpy as np
import torch

from . import kpconv, synthetic


# DistributedDataParallel bucket size (MB).  The all-reduce of the LAST bucket (the first layers' gradients) cannot
# hide behind the backward pass; with 56 MB of gradients the default 25 MB leaves a large tail exposed.
DDP_BUCKET_MB = 8
IGNORE_LABEL = -1  # datasets/segmentation/__init__.py:1; rows with this label do not count in the loss (kpconv.py:150)
# (Measured on 2 GPUs, graph-captured step: 25 MB buckets 6.38 ms, 8 MB buckets 6.31-6.34 ms; capping NCCL at 6
# channels gave 6.02 ms on one box and 6.34 ms on another -- inside the box-to-box spread, so no default is set.)

class KPConvSegmentationTrainer:
    def __init__(self, feat_dim=4, num_classes=13, device=None, in_feat=64, levels=5, ddp=False, seed=1003,
                 points_raw=24000, radius=0.62, num_spheres=8, max_nn_cap=34, lr=0.01, momentum=0.98):
        self.dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        pos_h, batch_h = synthetic.indoor_spheres(num_spheres=num_spheres, raw_points=points_raw, radius=radius, seed=seed)
        self.pos = torch.from_numpy(pos_h).to(self.dev)
        self.batch = torch.from_numpy(batch_h).to(self.dev)
        n = self.pos.shape[0]
        g = torch.Generator(device=self.dev).manual_seed(seed)
        self.feats = torch.randn(n, feat_dim, device=self.dev, generator=g)
        # labels follow the geometry (height bands x sphere parity) so that the loss can actually go down
        band = torch.clamp(((self.pos[:, 2] / radius + 1.0) * 0.5 * num_classes).long(), 0, num_classes - 1)
        self.labels = (band + self.batch) % num_classes
        self.model = kpconv.KPConvPaper(feat_dim, num_classes, in_feat=in_feat, levels=levels, max_nn_cap=max_nn_cap,
                                        random_rotation=False).to(self.dev)
        self.net = self.model
        if ddp:
            # gradients live in the all-reduce buckets (no copy in / out); buffers (BN statistics) are per rank.
            # Built under a side stream: a requirement for capturing the DDP step in a CUDA graph later.
            side = torch.cuda.Stream(device=self.dev)
            side.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(side):
                self.net = torch.nn.parallel.DistributedDataParallel(
                    self.model, device_ids=[self.dev.index] if torch.distributed.get_backend() == "nccl" else None,
                    broadcast_buffers=False, gradient_as_bucket_view=True, bucket_cap_mb=DDP_BUCKET_MB)
            torch.cuda.current_stream(self.dev).wait_stream(side)
        self.opt = torch.optim.SGD([p for p in self.model.parameters() if p.requires_grad], lr=lr, momentum=momentum)
        self.transform = kpconv.MultiScaleTransform(self.model.strategies())
        self.pyramid = None
        self.graph = None
        self.num_params = sum(p.numel() for p in self.model.parameters() if p.requires_grad)

    def build_pyramid(self):
        ms = self.transform(self.pos, self.batch)
        up = self.transform.upsample(self.pos, self.batch, ms)
        self.pyramid = (ms, up)
        return self.pyramid

    def _fwd_bwd(self, pyramid, set_to_none):
        self.opt.zero_grad(set_to_none=set_to_none)
        ms, up = pyramid
        logp = self.net(self.pos, self.feats, self.batch, multiscale=ms, upsample=up)
        loss = torch.nn.functional.nll_loss(logp, self.labels, ignore_index=IGNORE_LABEL)
        loss.backward()
        self.opt.step()
        return loss

    def step(self, rebuild_geometry=True):
        """one training step; `rebuild_geometry` = pay for the pyramid of a fresh batch inside the step"""
        if rebuild_geometry or self.pyramid is None:
            self.build_pyramid()
        return float(self._fwd_bwd(self.pyramid, True).detach())

    def step_async(self, rebuild_geometry=True):
        """as `step` without the device-to-host read of the loss (timing loops)"""
        if rebuild_geometry or self.pyramid is None:
            self.build_pyramid()
        return self._fwd_bwd(self.pyramid, True).detach()

    def capture(self):
        """capture forward + backward (+ the DDP all-reduce) + SGD on the current pyramid into one CUDA graph"""
        if self.pyramid is None:
            self.build_pyramid()
        quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if quiet is not None:
            quiet(False)  # AccumulateGrad nodes were created on the eager stream: intended here
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            # DDP needs 11 eager iterations before a whole-backward capture (bucket rebuild, reducer warm-up)
            for _ in range(11 if self.net is not self.model else 2):
                self._fwd_bwd(self.pyramid, False)
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.graph_loss = self._fwd_bwd(self.pyramid, False).detach()
        return self.graph

    def replay(self):
        self.graph.replay()
        return self.graph_loss


def param_checksum(model):
    """order-independent fingerprint of the parameters (rank-equality checks)"""
    with torch.no_grad():
        return float(sum(p.double().sum() for p in model.parameters()))


def exposed_allreduce_ms(trainer, reps=10):
    """time the training step with and without gradient synchronisation (DDP.no_sync): the difference is the part of
    the all-reduce that does not hide behind the backward pass.  Returns (ms_with, ms_without)."""
    import contextlib
    dev = trainer.dev
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def run(ctx):
        for _ in range(2):
            with ctx():
                trainer.step_async(rebuild_geometry=False)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(reps):
            with ctx():
                trainer.step_async(rebuild_geometry=False)
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps

    with_sync = run(contextlib.nullcontext)
    no_sync = run(trainer.net.no_sync) if hasattr(trainer.net, "no_sync") else with_sync
    return with_sync, no_sync
