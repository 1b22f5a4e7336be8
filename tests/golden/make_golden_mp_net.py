"""Run the data path of the REFERENCE's own MESSAGE_PASSING PointNet++ network (`pointnet2` of
conf/models/segmentation/pointnet2.yaml: two SAModule, GlobalBaseModule, three FPModule) and store its output.

Executed from /root/reference by file path (nothing is copied): modules/pointnet2/message_passing.py over
core/base_conv/message_passing.py (the loader of make_golden_randla.py), nested by the reference's own
`_init_from_compact_format` / `UnetSkipConnectionBlock` (make_golden_pointnet2_net.reference_unet_scope) from the entry
resolved by the reference's resolver.  torch_geometric is absent: `fps` / `radius` are oracle-backed, `knn` /
`knn_interpolate` / `global_max_pool` brute force, `PointConv` the restatement of make_golden_mp_chain.py.

    python tests/golden/make_golden_mp_net.py      ->  tests/golden/mp_net.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_golden_kpconv_arch as arch  # noqa: E402
import make_golden_mp_chain as mp_loader  # noqa: E402
import make_golden_pointnet2_net as nest  # noqa: E402
import make_golden_randla as loader  # noqa: E402
import mp_net_cases as cases  # noqa: E402


def main():
    loader.load_reference_base(radius=cases.radius)
    tg = sys.modules["torch_geometric.nn"]
    tg.PointConv, tg.fps, tg.knn_interpolate = mp_loader.PointConv, cases.fps, cases.knn_interpolate
    tg.global_max_pool, tg.global_mean_pool = cases.global_max_pool, cases.global_mean_pool
    for name in ("sampling", "interpolate"):          # loaded before the stand-ins above were complete: rebind
        mod = sys.modules["torch_points3d.core.spatial_ops." + name]
        for attr in ("fps", "knn_interpolate", "knn"):
            if hasattr(mod, attr):
                setattr(mod, attr, getattr(tg, attr))
    basemp = sys.modules["torch_points3d.core.base_conv.message_passing"]
    basemp.knn_interpolate, basemp.global_max_pool, basemp.global_mean_pool = (
        cases.knn_interpolate, cases.global_max_pool, cases.global_mean_pool)
    modules_lib = loader._load("ref_pointnet2_mp", os.path.join(loader.REF, "modules/pointnet2/message_passing.py"))
    base = sys.modules["torch_points3d.core.common_modules.base_modules"]
    resolver = arch.reference_resolver()
    entry = arch.load_entry("/root/reference/conf/models/segmentation/pointnet2.yaml", "pointnet2", {})
    dataset = type("Dataset", (), {"feature_dimension": cases.FEAT, "num_classes": 13})()
    resolver.resolve_model(entry, dataset, "segmentation")
    scope = nest.reference_unet_scope(base.Identity)
    Net = type("Net", (torch.nn.Module,), {name: scope[name] for name in nest.BUILDERS})
    net = Net()
    net._spatial_ops_dict = {"neighbour_finder": [], "sampler": [], "upsample_op": []}
    net._init_from_compact_format(entry, "pointnet2", dataset, modules_lib)
    cases.seeded_parameters(net, 900)
    net.eval()
    pos, x, batch = cases.make_input()
    with torch.no_grad():
        res = net.model(loader.Bag(pos=pos, x=x, batch=batch))
    out = {"state_dict_shapes": np.array(["%s" % (tuple(v.shape),) for v in net.state_dict().values()]),
           "x": res.x.numpy(), "pos": res.pos.numpy(), "batch": res.batch.numpy()}
    np.savez_compressed(os.path.join(HERE, "mp_net.npz"), **out)
    for k, v in out.items():
        print(k, v.shape)


if __name__ == "__main__":
    main()
