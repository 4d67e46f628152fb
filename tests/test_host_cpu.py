"""CPU-side checks: the C-ABI library loads and exports every symbol of include/coalign_b200.h, the ctypes
descriptor mirrors the C struct, host-side weight packing / K-step logic, plugin registry contract."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_lib():
    from coalign_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol():
    path = _ensure_lib()
    hdr = open(os.path.join(ROOT, "include", "coalign_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|size_t)\s+(cb_\w+)\s*\(", hdr, flags=re.M))
    assert len(declared) >= 12
    lib = ctypes.CDLL(path)
    for name in declared:
        assert hasattr(lib, name), name
    from coalign_b200 import _lib
    assert set(_lib.EXPORTS) == declared
    assert _lib.load().cb_version() == 100


def test_ctypes_descriptor_matches_c_struct(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "coalign_b200.h"\n'
                   'int main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(cb_conv_desc), offsetof(cb_conv_desc, ksteps),'
                   'offsetof(cb_conv_desc, head_out), offsetof(cb_conv_desc, out_plane_rows), sizeof(cb_kstep));return 0;}')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    vals = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    from coalign_b200._lib import ConvDesc, KStep
    assert vals == [ctypes.sizeof(ConvDesc), ConvDesc.ksteps.offset, ConvDesc.head_out.offset,
                    ConvDesc.out_plane_rows.offset, ctypes.sizeof(KStep)]


def test_missing_gpu_fails_loudly():
    from coalign_b200 import synth
    from coalign_b200.model import PointPillarCoalignB200
    from tests import golden_cases as G
    m = PointPillarCoalignB200(G.small_args()).eval()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        m({"processed_lidar": {}, "record_len": torch.tensor([1]), "pairwise_t_matrix": torch.zeros(1, 5, 5, 4, 4)})
    with pytest.raises(RuntimeError):
        from coalign_b200.engine import CoAlignEngine
        CoAlignEngine(G.small_args(), synth.random_state_dict(G.small_args(), 0), 1, 1)


def test_state_dict_contract():
    from coalign_b200 import synth
    from coalign_b200.model import PointPillarCoalignB200
    args = synth.opv2v_args()
    m = PointPillarCoalignB200(args)
    sd = synth.random_state_dict(args, 0)
    assert set(m.state_dict()) == set(sd) and len(sd) == 244
    assert sum(p.numel() for p in m.parameters()) == 12901524
    m.load_state_dict(sd, strict=True)


def test_single_agent_state_dict_contract():
    """PointPillarB200 (reference core_method `point_pillar`, BaseBEVBackbone): same keys / shapes as the reference
    model (the synthetic state_dict was checked key-by-key against the reference in tests/golden/gen_golden_single.py)."""
    from coalign_b200 import synth
    from coalign_b200.model import PointPillarB200
    args = synth.single_args()
    m = PointPillarB200(args)
    sd = synth.random_state_dict(args, 0, backbone="plain")
    assert set(m.state_dict()) == set(sd)
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    m.load_state_dict(sd, strict=True)
    with pytest.raises(RuntimeError):                  # CPU module: no fallback
        m.eval()({"processed_lidar": {"voxel_features": torch.zeros(1, 32, 4), "voxel_coords": torch.zeros(1, 4, dtype=torch.int32),
                                      "voxel_num_points": torch.ones(1, dtype=torch.int32)}})


def test_weight_packing_is_a_gemm_restatement_of_conv():
    """pack_conv_weight + tap shifts reproduce F.conv2d (the host logic behind the K-step tables)."""
    from coalign_b200.engine import CoAlignEngine, pack_conv_weight
    g = torch.Generator().manual_seed(0)
    cin, cout, H, W = 64, 8, 5, 7
    x = torch.randn(1, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g)
    ref = torch.nn.functional.conv2d(x, w, padding=1)
    Hp, Wp = H + 2, W + 2
    xp = torch.zeros(Hp * Wp, cin)
    xp.view(Hp, Wp, cin)[1:-1, 1:-1] = x[0].permute(1, 2, 0)
    wp = pack_conv_weight(w, None).float()
    out = torch.zeros(Hp * Wp, cout)
    for (ro, col, wk, sel) in CoAlignEngine._steps_3x3_s1(cin, Wp):
        rows = torch.arange(Hp * Wp) + ro
        ok = (rows >= 0) & (rows < Hp * Wp)
        a = torch.zeros(Hp * Wp, 64)
        a[ok] = xp[rows[ok], col:col + 64]
        out += a @ wp[:, wk:wk + 64].t()
    got = out.view(Hp, Wp, cout)[1:-1, 1:-1].permute(2, 0, 1)
    assert torch.allclose(got, ref[0], atol=1e-3)


def _to_ps(x, n_cap):
    """dense (N,C,H,W) -> PS rows [4*plane_rows, C] (include/coalign_b200.h layout)."""
    N, C_, H, W = x.shape
    Hq, Wq = (H + 1) // 2 + 2, (W + 1) // 2 + 2
    plane_rows = n_cap * Hq * Wq
    buf = torch.zeros(4 * plane_rows, C_)
    for n in range(N):
        for h in range(H):
            for w in range(W):
                ph = (h & 1) * 2 + (w & 1)
                buf[ph * plane_rows + (n * Hq + (h >> 1) + 1) * Wq + (w >> 1) + 1] = x[n, :, h, w]
    return buf, Hq, Wq, plane_rows


@pytest.mark.parametrize("H,W", [(6, 8), (5, 7)])
def test_stride2_phase_split_ksteps(H, W):
    """K-step table of the k3/s2/p1 conv over a PS input == F.conv2d(stride=2) (+ the fused 1x1/s2 branch)."""
    from coalign_b200.engine import Act, CoAlignEngine, pack_conv_weight
    g = torch.Generator().manual_seed(1)
    cin, cout, N, cap = 64, 4, 2, 3
    x = torch.randn(N, cin, H, W, generator=g)
    w3 = torch.randn(cout, cin, 3, 3, generator=g)
    w1 = torch.randn(cout, cin, 1, 1, generator=g)
    ref = torch.nn.functional.conv2d(x, w3, stride=2, padding=1) + torch.nn.functional.conv2d(x, w1, stride=2)
    buf, Hq, Wq, plane_rows = _to_ps(x, cap)

    class Src:                    # the attributes _steps_3x3_s2 reads from an Act
        pass
    src = Src()
    src.plane_rows, src.Wp = plane_rows, Wq
    steps = CoAlignEngine._steps_3x3_s2(cin, src) + CoAlignEngine._steps_1x1(cin, 0, 9 * cin)
    wp = torch.cat([pack_conv_weight(w3, None), pack_conv_weight(w1, None)], 1).float()
    rows_total = N * Hq * Wq
    out = torch.zeros(rows_total, cout)
    for (ro, col, wk, sel) in steps:
        rows = torch.arange(rows_total) + ro
        ok = (rows >= 0) & (rows < buf.shape[0])
        a = torch.zeros(rows_total, 64)
        a[ok] = buf[rows[ok], col:col + 64]
        out += a @ wp[:, wk:wk + 64].t()
    Ho, Wo = (H + 1) // 2, (W + 1) // 2
    got = out.view(N, Hq, Wq, cout)[:, 1:Ho + 1, 1:Wo + 1].permute(0, 3, 1, 2)
    assert torch.allclose(got, ref, atol=1e-3), (got - ref).abs().max()


def test_stage1_uncertainty_state_dict_contract_and_registry():
    """PointPillarUncertaintyB200 (reference core_method `point_pillar_uncertainty`): same keys / shapes as the reference
    model (checked key-by-key against the reference in tests/golden/gen_golden_single.py::main_uncertainty); the plugin
    module name resolves to the class the way train_utils.create_model does (train_utils.py:127-146)."""
    import importlib
    from coalign_b200 import synth, PLUGIN_DIR
    from coalign_b200.model import PointPillarUncertaintyB200
    args = synth.uncertainty_args()
    m = PointPillarUncertaintyB200(args)
    sd = synth.random_state_dict(args, 0, backbone="plain")
    assert set(m.state_dict()) == set(sd)
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    m.load_state_dict(sd, strict=True)
    assert m.unc_head.weight.shape == (6, 384, 1, 1)
    import sys
    sys.path.insert(0, PLUGIN_DIR)
    try:
        mod = importlib.import_module("point_pillar_uncertainty_b200")
    finally:
        sys.path.remove(PLUGIN_DIR)
    target = "point_pillar_uncertainty_b200".replace("_", "").lower()
    found = [c for name, c in mod.__dict__.items() if name.lower() == target]
    assert found == [PointPillarUncertaintyB200]
    with pytest.raises(KeyError):
        PointPillarUncertaintyB200(synth.single_args())


def test_c_abi_rejects_bad_arguments_without_a_device():
    """Every entry point validates its arguments before the first CUDA call (returns CB_ERR_ARG = -1, never throws), so
    this runs without a GPU; workspace queries are pure host arithmetic."""
    import ctypes as C
    import numpy as np
    from coalign_b200 import _lib
    lib = _lib.load(check_device=False)
    grid = np.array([704, 200, 1], np.int32)
    grid3 = np.array([704, 200, 2], np.int32)
    rng = np.array([-140.8, -40, -3, 140.8, 40, 1], np.float32)
    vs = np.array([0.4, 0.4, 4], np.float32)
    off = np.array([0, 10], np.int32)
    # workspace sizes: monotone in the agent count; the fused path (nz == 1) needs the per-cell slot arrays
    w1 = lib.cb_voxelize_workspace_bytes(1, 60000, grid.ctypes.data, 70000)
    w5 = lib.cb_voxelize_workspace_bytes(5, 300000, grid.ctypes.data, 70000)
    assert w5 > w1 >= 200 * 704 * 32 * 16
    assert lib.cb_voxelize_workspace_bytes(1, 60000, grid3.ctypes.data, 70000) < w1      # nz > 1: ordered pipeline only
    dummy = C.c_void_p(256)                                     # non-null, 16-byte aligned, never dereferenced
    # fused front-end: nz must be 1, canvas capacity >= agents, workspace / canvas non-null
    assert lib.cb_points_to_canvas(dummy, off.ctypes.data, 1, rng.ctypes.data, vs.ctypes.data, grid3.ctypes.data, 32, 70000,
                                   dummy, dummy, dummy, vs.ctypes.data, 1, dummy, 0, None, None, dummy, 1 << 20, None) == -1
    assert lib.cb_points_to_canvas(dummy, off.ctypes.data, 2, rng.ctypes.data, vs.ctypes.data, grid.ctypes.data, 32, 70000,
                                   dummy, dummy, dummy, vs.ctypes.data, 1, dummy, 0, None, None, dummy, 1 << 20, None) == -1
    assert lib.cb_points_to_canvas(dummy, off.ctypes.data, 1, rng.ctypes.data, vs.ctypes.data, grid.ctypes.data, 33, 70000,
                                   dummy, dummy, dummy, vs.ctypes.data, 1, dummy, 0, None, None, dummy, 1 << 20, None) == -1
    # voxeliser: max_pts in [1, 32], workspace required
    assert lib.cb_voxelize(dummy, off.ctypes.data, 1, rng.ctypes.data, vs.ctypes.data, grid.ctypes.data, 0, 70000, dummy,
                           dummy, dummy, dummy, dummy, 1 << 20, None) == -1
    assert lib.cb_voxelize(dummy, off.ctypes.data, 1, rng.ctypes.data, vs.ctypes.data, grid.ctypes.data, 32, 70000, dummy,
                           dummy, dummy, dummy, None, 0, None) == -1
    # fusion: channel count multiple of 64, method in {0, 1}
    assert lib.cb_warp_att_fuse(dummy, 0, 0, 5, dummy, dummy, 1, 5, 100, 352, 48, 0, dummy, 0, None) == -1
    assert lib.cb_warp_att_fuse(dummy, 0, 0, 5, dummy, dummy, 1, 5, 100, 352, 64, 2, dummy, 0, None) == -1
    # post-processing: top_k <= 1024, outputs required
    gt = np.zeros(6, np.float64)
    assert lib.cb_postprocess(dummy, dummy, None, 1, 100, 352, 2, 0, dummy, dummy, 0.2, 0.0, 0.15, gt.ctypes.data, 1, 2000,
                              dummy, dummy, dummy, dummy, 1 << 20, None) == -1
    assert lib.cb_postprocess_stage1(dummy, dummy, None, 1, 100, 352, 2, 0, dummy, 0.2, 0.0, 0.15, 1, 1000,
                                     dummy, None, None, dummy, dummy, dummy, 1 << 20, None) == -1
    assert lib.cb_postprocess_workspace_bytes(0, 100, 352, 2) == 0


def test_unmodified_reference_registries_resolve_the_plugins():
    """With the reference tree present (build container), `train_utils.create_model` / `create_loss` of the UNMODIFIED
    reference return our classes once `coalign_b200.register()` ran and only `core_method` changed in the hypes
    (SURVEY 8b registry contract).  Runs in a subprocess (import-only stubs for open3d / matplotlib / box_overlaps)."""
    import os
    import subprocess
    import sys
    if not os.path.isdir("/root/reference/opencood"):
        pytest.skip("reference tree not present on this machine")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys, types
import numpy as np
sys.modules["open3d"] = types.ModuleType("open3d")
import matplotlib
cm = types.ModuleType("matplotlib.cm"); cm.get_cmap = lambda name: types.SimpleNamespace(colors=np.zeros((256, 3)))
matplotlib.cm = cm; sys.modules["matplotlib.cm"] = cm
bo = types.ModuleType("opencood.utils.box_overlaps"); bo.bbox_overlaps = None; sys.modules["opencood.utils.box_overlaps"] = bo
import coalign_b200
from coalign_b200 import synth
from opencood.tools import train_utils
coalign_b200.register()
hypes = {"model": {"core_method": "point_pillar_coalign_b200", "args": synth.opv2v_args()},
         "loss": {"core_method": "point_pillar_loss_b200", "args": synth.loss_args()}}
names = [type(train_utils.create_model(hypes)).__name__, type(train_utils.create_loss(hypes)).__name__]
hypes["model"] = {"core_method": "point_pillar_b200", "args": synth.single_args()}
names.append(type(train_utils.create_model(hypes)).__name__)
hypes["model"] = {"core_method": "point_pillar_uncertainty_b200", "args": synth.uncertainty_args()}
names.append(type(train_utils.create_model(hypes)).__name__)
print(",".join(names))
'''
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(root, "tests", "golden", "_stubs"), "/root/reference", root]))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip().splitlines()[-1] == \
        "PointPillarCoalignB200,PointPillarLossB200,PointPillarB200,PointPillarUncertaintyB200"


def test_collate_restatement_matches_the_reference_static_method():
    """A2 (collate_batch_dict, sp_voxel_preprocessor.py:145-174) is plain numpy inside the reference tree: with the tree
    present, the oracle's `collate` is checked against the unmodified static method (A1, the spconv generator itself,
    stays unpinned - the package is absent).  Subprocess: import-only stubs (icecream, open3d, pypcd, matplotlib, box_overlaps)."""
    import os
    import subprocess
    import sys
    if not os.path.isdir("/root/reference/opencood"):
        pytest.skip("reference tree not present on this machine")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys, types
import numpy as np
sys.modules["open3d"] = types.ModuleType("open3d")
pp = types.ModuleType("pypcd"); pp.pypcd = types.ModuleType("pypcd.pypcd"); sys.modules["pypcd"] = pp; sys.modules["pypcd.pypcd"] = pp.pypcd
import matplotlib
cm = types.ModuleType("matplotlib.cm"); cm.get_cmap = lambda name: types.SimpleNamespace(colors=np.zeros((256, 3)))
matplotlib.cm = cm; sys.modules["matplotlib.cm"] = cm
bo = types.ModuleType("opencood.utils.box_overlaps"); bo.bbox_overlaps = None; sys.modules["opencood.utils.box_overlaps"] = bo
from opencood.data_utils.pre_processor.sp_voxel_preprocessor import SpVoxelPreprocessor
from coalign_b200 import synth
from oracle import voxelize_np as V
rng = np.random.default_rng(0)
R, VS = [-11.2, -4.8, -3, 11.2, 4.8, 1], [0.4, 0.4, 4]
per_agent = [V.voxelize_c(synth.lidar_cloud(rng, n, R, sigma=5.0), R, VS, 32, 70000) for n in (700, 1, 1500)]
ref = SpVoxelPreprocessor.collate_batch_dict({"voxel_features": [a[0] for a in per_agent],
                                              "voxel_coords": [a[1] for a in per_agent],
                                              "voxel_num_points": [a[2] for a in per_agent]})
vf, vc, vn = V.collate(per_agent)
assert np.array_equal(ref["voxel_features"].numpy(), vf)
assert np.array_equal(ref["voxel_coords"].numpy(), vc) and ref["voxel_coords"].shape[1] == 4
assert np.array_equal(ref["voxel_num_points"].numpy(), vn)
print("collate ok", vc.shape)
'''
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(root, "tests", "golden", "_stubs"), "/root/reference", root]))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "collate ok" in r.stdout
