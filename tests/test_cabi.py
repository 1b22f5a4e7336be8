"""CPU tests of the drop-in boundary: the C-ABI library loads, exports exactly what
include/p3d_b200.h declares, and the host-side mirror keeps the reference's argument checking
and error behaviour.  No compute calls (no GPU here)."""
import os
import re
import subprocess

import pytest
import torch

import __graft_entry__ as entry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cabi():
    entry.build()
    from torch_points3d_b200 import _cabi
    return _cabi


def _header_functions():
    text = open(os.path.join(ROOT, "include", "p3d_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(p3d_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_the_library_exports(cabi):
    declared = _header_functions()
    assert len(declared) >= 20
    out = subprocess.run(["nm", "-D", "--defined-only", cabi.LIB_PATH], capture_output=True, text=True, check=True)
    exported = sorted(set(re.findall(r" T (p3d_[a-z0-9_]+)", out.stdout)))
    assert exported == declared
    assert sorted(cabi.SIGNATURES) == declared  # ctypes binding covers every entry point


def test_library_loads_and_reports_arch(cabi):
    lib = cabi.lib()
    assert lib.p3d_arch() == b"sm_100a"
    assert lib.p3d_version() >= 100
    assert lib.p3d_launch_count() >= 0
    assert lib.p3d_fps_dense_ws_bytes(4, 1024) == 0
    assert lib.p3d_fps_dense_ws_bytes(2, 200000) == 2 * 200000 * 4
    # wf + dwf buffers, plus the bf16 hi/lo image of W^T for the tensor-core dwf GEMM (15 x 32 outputs / 32 per chunk x 4 pieces x 16 KB)
    # ... plus the workspace of the weight gradient on the tcgen05 row GEMM: the bf16 hi/lo image of dout (one 64-point
    # chunk x 32 columns = 8 KB) and one partial K*Cin x Cout block (the reduction over 10 points is not split)
    assert lib.p3d_kpconv_bwd_ws_bytes(10, 34, 15, 32, 32) == 2 * 10 * 15 * 32 * 4 + 15 * 4 * 16384 + (8192 + 480 * 32 * 4)
    # Cin = 5: no tensor-core dwf image (K * Cin = 75 is not a multiple of 16); the weight gradient's workspace is sized
    # for the row GEMM (Cout = 64 is): a 16 KB image of dout + one 75 x 64 partial block
    assert lib.p3d_kpconv_bwd_ws_bytes(10, 34, 15, 5, 64) == 2 * 10 * 15 * 5 * 4 + (2 * 64 * 128 + 75 * 64 * 4)


def test_sass_is_sm100a_only(cabi):
    out = subprocess.run(["cuobjdump", "-lelf", cabi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}


def test_argument_validation_is_reported_not_crashed(cabi):
    lib = cabi.lib()
    # S > N is rejected on the host side of the ABI before any CUDA call
    assert lib.p3d_fps_dense_f32(None, 1, 4, 5, None, 1, None, None, 0, None) == -1
    assert b"npoint" in lib.p3d_last_error()
    assert lib.p3d_knn_batch_f32(None, None, None, None, 4, 4, 0, None, None, None) == -1
    assert lib.p3d_three_nn_f32(None, None, 1, 4, 2, None, None, None, None) == -1
    assert lib.p3d_kpconv_fwd_f32(None, None, None, None, None, None, 4, 4, 200, 15, 8, 8, 1.0, 1, 0, 1, None, None,
                                  0, None) == -1
    assert lib.p3d_kpconv_fwd_ws_bytes(100000, 15, 32, 64) == 1 * 2 * 2 * 32768 and lib.p3d_kpconv_fwd_ws_bytes(100000, 15, 5, 64) == 0
    # few query points: the slabs of the contraction are split over CTAs, partial outputs live in the workspace
    assert lib.p3d_kpconv_fwd_ws_bytes(100, 15, 512, 512) > 16 * 16 * 4 * 16384
    # empty problems are a successful no-op
    assert lib.p3d_fps_dense_f32(None, 0, 4, 2, None, 1, None, None, 0, None) == 0
    assert lib.p3d_ball_query_dense_f32(None, None, 0, 0, 0, 0.1, 4, 1, None, None, None) == 0


def test_no_cpu_fallback():
    from torch_points3d_b200 import ops
    xyz = torch.rand(2, 16, 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.furthest_point_sample(xyz, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.ball_query(0.2, 4, xyz, xyz[:, :4])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.grouping_operation(torch.rand(2, 3, 16), torch.zeros(2, 4, 4, dtype=torch.long))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.KPConv_ops(xyz[0], xyz[0], torch.zeros(16, 4, dtype=torch.long), torch.rand(16, 8), torch.rand(15, 3),
                       torch.rand(15, 8, 8), 0.1, "linear", "sum")


def test_reference_error_behaviour():
    """same exception types as torch_points_kernels / the reference wrappers"""
    from torch_points3d_b200 import ops, spatial_ops as so
    xyz = torch.rand(2, 16, 3)
    with pytest.raises(ValueError):
        ops.furthest_point_sample(xyz, 17)                      # tpk: npoint > N
    with pytest.raises(Exception, match="batch_x and batch_y should be provided"):
        ops.ball_query(0.1, 4, xyz[0], xyz[0], mode="partial_dense")
    with pytest.raises(Exception, match="should not be provided"):
        ops.ball_query(0.1, 4, xyz, xyz, mode="dense", batch_x=torch.zeros(16, dtype=torch.long))
    with pytest.raises(NotImplementedError):
        ops.ball_query(0.1, 4, xyz, xyz, sort=True)
    with pytest.raises(ValueError):
        ops.three_nn(xyz, xyz[:, :2])
    with pytest.raises(ValueError):
        ops.KPConv_ops(xyz[0], xyz[0], None, torch.rand(16, 8), torch.rand(15, 3), torch.rand(15, 8, 8), 0.1,
                       "cubic", "sum")
    with pytest.raises(ValueError):
        so.DenseFPSSampler(num_to_sample=4).sample(xyz[0])     # sampling.py:97-98
    with pytest.raises(ValueError):
        so.FPSSampler(ratio=0.5).sample(xyz, None)             # sampling.py:60-61
    with pytest.raises(ValueError):
        so.BaseSampler.__init__(so.FPSSampler.__new__(so.FPSSampler), ratio=0.5, num_to_sample=3)
    with pytest.raises(Exception):
        so.DenseFPSSampler()
    with pytest.raises(ValueError):
        so.MultiscaleRadiusNeighbourFinder([0.1, 0.2], [4])
    with pytest.raises(ValueError):
        so.DenseRadiusNeighbourFinder([0.1], [4]).find_neighbours(xyz, xyz, scale_idx=1)


def test_sampler_count_rules():
    from torch_points3d_b200 import spatial_ops as so
    s = so.FPSSampler(ratio=0.5)
    assert s._get_num_to_sample(10) == 5 and s._get_ratio_to_sample(10) == 0.5   # test_sampler.py:16-23
    s = so.FPSSampler(num_to_sample=5)
    assert s._get_num_to_sample(10) == 5 and s._get_ratio_to_sample(10) == 0.5
    f = so.MultiscaleRadiusNeighbourFinder([0.1, 0.2], 8)
    assert f.num_scales == 2 and f._max_num_neighbors == [8, 8]
    f = so.MultiscaleRadiusNeighbourFinder(0.3, [8, 16])
    assert f._radius == [0.3, 0.3]
    assert so.RadiusNeighbourFinder(0.1, 4, conv_type="PARTIAL_DENSE")._conv_type == "partial_dense"


def test_kernel_points_fixture_matches_reference_disposition(golden_dir):
    import numpy as np
    from torch_points3d_b200 import kpconv
    kp = kpconv.load_kernels(1.0, random_rotation=False)
    ref = np.load(os.path.join(golden_dir, "k_015_center.npy"))
    assert kp.shape == (15, 3) and np.allclose(kp, ref.astype(np.float32))
    assert np.allclose(kp[0], 0)                       # 'center' disposition: first kernel point at the origin
    rot = kpconv.load_kernels(0.06)                    # rotation + 1% jitter keeps the radii
    assert abs(np.linalg.norm(rot, axis=1)[1:].mean() - 0.06 * np.linalg.norm(ref, axis=1)[1:].mean()) < 0.01
    layer = kpconv.KPConvLayer(8, 16, point_influence=0.04, add_one=True)
    assert layer.weight.shape == (15, 9, 16) and not layer.K_points.requires_grad
    assert "K_points" in layer.state_dict()            # checkpoints carry the kernel points (kernels.py:66-68)


def test_new_entry_points_validate_before_any_cuda_call(cabi):
    lib = cabi.lib()
    assert lib.p3d_grid_cluster_f32(None, None, 4, 0.0, None, None, None, None, None, None, None, None, None, 0, None) == -1
    assert b"voxel size" in lib.p3d_last_error()
    assert lib.p3d_grid_cluster_ws_bytes(0) == 256 and lib.p3d_grid_cluster_ws_bytes(1000) >= 40 * 1000
    assert lib.p3d_ball_query_dense_multi_f32(None, None, 1, 4, 2, 5, None, None, 1, None, None) == -1
    assert b"scales" in lib.p3d_last_error()
    assert lib.p3d_segment_reduce_f32(None, None, 4, 2, 3, 7, None, None, None) == -1
    assert lib.p3d_kpconv_deform_fwd_f32(None, None, None, None, None, None, None, None, 4, 4, 500, 15, 8, 8, 0.1, 1,
                                         None, None, None, None, None, None) == -1
    assert lib.p3d_kpconv_deform_bwd_ws_bytes(10, 15, 32) == 10 * 15 * 32 * 4
    assert lib.p3d_group_xyz_feat_f32(None, None, None, None, 1, 0, 4, 2, 2, 1.0, None, None) == -1
    # narrow first-level backward: per-split partial tables; wide tables: the inverse-index layout
    assert lib.p3d_group_bwd_ws_bytes(16, 3, 2048, 512, 32) == 9 * 16 * 3 * 2048 * 4
    assert lib.p3d_group_bwd_ws_bytes(16, 320, 512, 128, 128) > 16 * 16384 * 2
    # empty problems are a successful no-op
    assert lib.p3d_edge_message_f32(None, None, None, None, None, 0, 8, None, None) == 0
    assert lib.p3d_kpconv_wf_f32(None, None, None, None, None, 0, 0, 4, 15, 8, 0.1, 1, 0, None, None) == 0


def test_host_logic_of_the_next_row_mirrors():
    """argument checking and bookkeeping that never reaches the device"""
    from torch_points3d_b200 import grid_transform as gt, kpconv, message_passing as mp, ops
    d = gt.Data(pos=torch.rand(6, 3), x=torch.rand(6, 2), y=torch.zeros(6, dtype=torch.long))
    assert sorted(d.keys) == ["pos", "x", "y"] and d.num_nodes == 6 and "batch" not in d
    c = d.clone()
    c["x"] = c["x"] + 1
    assert not torch.equal(c.x, d.x)
    with pytest.raises(ValueError):
        gt.group_data(d, mode="mean")                          # grid_transform.py:53-54
    with pytest.raises(ValueError):
        gt.group_data(d, cluster=torch.zeros(6, dtype=torch.long), mode="last")   # :55-56
    with pytest.raises(AssertionError):
        gt.group_data(d, mode="median")
    with pytest.raises(ValueError, match="Edges not supported"):
        gt.group_data(gt.Data(pos=torch.rand(3, 3), edge_index=torch.zeros(2, 3)), cluster=torch.zeros(3),
                      unique_pos_indices=torch.zeros(1, dtype=torch.long), mode="last")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        gt.GridSampling3D(0.1)(d)
    assert "grid_size=0.1" in repr(gt.GridSampling3D(0.1))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.ball_query_multi([0.1], [4], torch.rand(1, 8, 3), torch.rand(1, 2, 3))
    with pytest.raises(TypeError):
        ops.KPConv_deform_ops(None, None, None, torch.rand(4, 8), torch.rand(15, 3), torch.rand(4, 15, 3), None,
                              torch.rand(15, 8, 8), 0.1, "linear", "closest")
    with pytest.raises(AssertionError):
        ops.KPConv_deform_ops(None, None, None, torch.rand(4, 8), torch.rand(15, 3), torch.rand(4, 14, 3), None,
                              torch.rand(15, 8, 8), 0.1, "linear", "sum")
    enc = kpconv.KPConvEncoder(5, max_nn_cap=34)
    st = enc.strategies()
    assert len(st["sampler"]) == len(st["neighbour_finder"]) == 10
    assert [s is not None for s in st["sampler"]] == [False, False, True, False, True, False, True, False, True, False]
    assert st["neighbour_finder"][2]._radius == pytest.approx(2.5 * 0.04)   # blocks.py:52: 2.5 * sigma * prev_grid
    layer = kpconv.KPConvDeformableLayer(8, 16, point_influence=0.04, modulated=True)
    assert layer.offset_weights.shape == (15, 8, 60) and layer.weight.shape == (15, 8, 16)
    assert set(layer.get_internal_losses()) == {"permissive_loss", "fitting_loss", "repulsion_loss"}
    conv = mp.PointConv(local_nn=mp.MLP([6, 8]), global_nn=None)
    assert conv.add_self_loops and isinstance(conv.local_nn, torch.nn.Sequential)
    sa = mp.SAModule(ratio=0.5, radius=[0.1, 0.2], radius_num_point=[8, 16], down_conv_nn=[6, 8])
    assert sa.neighbour_finder.num_scales == 2


def test_sass_uses_the_blackwell_units_the_design_names(cabi):
    """DESIGN.md section 4 in SASS (B200_PROFILING.md mnemonics): tcgen05 MMAs with TMEM loads in the KPConv tensor
    path, TMA bulk copies + mbarrier transaction waits in the staged kernels, warp redux in FPS, cluster barriers
    in the large-cloud FPS"""
    out = subprocess.run(["cuobjdump", "-sass", cabi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    sass = out.stdout
    per_kernel = {}
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            per_kernel[name] = []
        elif name is not None:
            per_kernel[name].append(line)

    def has(kernel_substr, mnemonic):
        return any(kernel_substr in k and any(mnemonic in l for l in body) for k, body in per_kernel.items())

    assert has("kpconv_fwd_tc_kernel", "UTCHMMA") and has("kpconv_fwd_tc_kernel", "LDTM")      # tcgen05.mma / tcgen05.ld
    assert has("kpconv_fwd_tc_kernel", "UBLKCP") and has("kpconv_fwd_tc_kernel", "UTCBAR")     # TMA weights, tcgen05.commit
    assert has("gemm_tc_kernel", "UTCHMMA")
    assert has("rowgemm_tc_kernel", "UTCHMMA") and has("rowgemm_tc_kernel", "UBLKCP") and has("rowgemm_tc_kernel", "LDTM")
    assert has("bwd_sorted_kernel", "UBLKCP") and has("bwd_sorted_kernel", "SYNCS.PHASECHK")   # TMA ring + mbarrier wait
    assert has("fps_cta_kernel", "REDUX") and has("fps_reg_kernel", "UCGABAR")                 # warp redux, cluster barrier
    assert has("kpconv_fwd_tc_kernel", "LDGSTS")                                               # cp.async row staging


def test_kpconv_autograd_plumbing_with_stubbed_device_calls(monkeypatch):
    """which entry points a training / inference / frozen-weight KPConv call reaches, and that every autograd
    Function returns one gradient per input -- checked on the CPU with the device calls stubbed out"""
    from torch_points3d_b200 import ops
    calls = []
    monkeypatch.setattr(ops, "_call", lambda dev, name, *args: calls.append(name))
    monkeypatch.setattr(ops, "_require_cuda", lambda name, *t: torch.device("cpu"))

    class FakeLib:
        def p3d_kpconv_fwd_ws_bytes(self, n, K, Cin, Cout):
            return 1024 if Cin % 32 == 0 and Cout % 32 == 0 else 0

        def p3d_kpconv_bwd_ws_bytes(self, *a):
            return 1024

        def p3d_kpconv_deform_bwd_ws_bytes(self, *a):
            return 1024

    monkeypatch.setattr(ops._cabi, "lib", lambda: FakeLib())
    monkeypatch.setattr(ops, "KPCONV_SPLIT_MAX_POINTS", 0)
    q, nb, kp = torch.rand(6, 3), torch.zeros(6, 4, dtype=torch.long), torch.rand(15, 3)
    x = torch.rand(6, 32, requires_grad=True)
    W = torch.rand(15, 32, 32, requires_grad=True)
    out = ops.KPConv_ops(q, q, nb, x, kp, W, 0.1, "linear", "sum")
    out.sum().backward()
    assert calls == ["p3d_kpconv_fwd_wf_f32", "p3d_kpconv_bwd_wf_f32"]      # training: the forward keeps wf
    assert x.grad.shape == x.shape and W.grad.shape == W.shape
    calls.clear()
    with torch.no_grad():
        ops.KPConv_ops(q, q, nb, x, kp, W, 0.1, "linear", "sum")
    assert calls == ["p3d_kpconv_fwd_f32"]                                  # inference: nothing extra is written
    calls.clear()
    ops.KPConv_ops(q, q, nb, x, kp, W.detach(), 0.1, "linear", "sum").sum().backward()
    assert calls == ["p3d_kpconv_fwd_f32", "p3d_kpconv_bwd_wf_f32"]         # frozen weights: wf is not needed
    calls.clear()
    ops.KPConv_ops(q, q, nb, x, kp, W, 0.1, "linear", "sum", precision="fp32").sum().backward()
    assert calls == ["p3d_kpconv_fwd_f32", "p3d_kpconv_bwd_wf_f32"]
    calls.clear()
    off = torch.zeros(6, 15, 3, requires_grad=True)
    mods = torch.ones(6, 15, requires_grad=True)
    o, sqd, dk = ops.KPConv_deform_ops(q, q, nb, x, kp, off, mods, W, 0.1, "linear", "sum")
    (o.sum() + sqd.sum() + dk.sum()).backward()
    assert calls == ["p3d_kpconv_deform_fwd_f32", "p3d_kpconv_deform_bwd_f32"]
    assert off.grad.shape == off.shape and mods.grad.shape == mods.shape


def test_integration_md_stub_matches_the_header(cabi):
    """INTEGRATION.md section 2 shows the ctypes binding a maintainer would add.  Execute that text against the
    built library: the argtypes must have one entry per parameter of the header's declaration and the documented
    call must reach the library's own argument validation (npoint > N is rejected before any CUDA call)."""
    import types
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"```python\n(# torch_points3d/core/spatial_ops/_p3d_b200\.py.*?)```", text, flags=re.S)
    assert m, "INTEGRATION.md lost its binding stub"
    src = m.group(1).replace('ctypes.CDLL("libp3d_b200.so")', "ctypes.CDLL(%r)" % cabi.LIB_PATH)
    mod = types.ModuleType("_p3d_b200_stub")
    exec(compile(src, "INTEGRATION.md:stub", "exec"), mod.__dict__)
    header = open(os.path.join(ROOT, "include", "p3d_b200.h")).read()
    decl = re.search(r"int p3d_fps_dense_f32\((.*?)\);", header, flags=re.S).group(1)
    assert len(mod._lib.p3d_fps_dense_f32.argtypes) == len(decl.split(","))
    assert cabi.SIGNATURES["p3d_fps_dense_f32"][1] == list(mod._lib.p3d_fps_dense_f32.argtypes)
    with pytest.raises(ValueError):
        mod.furthest_point_sample(torch.zeros(1, 4, 3), 5)          # the stub's own check (reference message)
    rc = mod._lib.p3d_fps_dense_f32(None, 1, 4, 5, None, 1, None, None, 0, None)
    assert rc == -1 and b"npoint" in mod._lib.p3d_last_error()
    # an empty problem goes through the documented call path end to end without a device
    assert mod.furthest_point_sample(torch.zeros(0, 4, 3), 2).shape == (0, 2)
