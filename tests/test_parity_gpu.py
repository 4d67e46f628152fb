"""GPU parity tests: the CUDA path (through the C ABI / ctypes) against the CPU oracle and the golden
vectors of the unmodified reference.  Run on the B200 box: pytest -m gpu."""
import os

import numpy as np
import pytest
import torch

from coalign_b200 import synth
from tests import golden_cases as G

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def assert_close(a, b, rtol, atol_rms, what=""):
    """|a-b| <= rtol*|b| + atol_rms*rms(b): rtol on every element, with the absolute floor tied to the tensor's
    scale (elements near zero carry the accumulated error of ~35 stacked layers, not a relative one)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b)
    tol = rtol * np.abs(b) + atol_rms * np.sqrt((b * b).mean())
    bad = err > tol
    assert not bad.any(), f"{what}: {bad.sum()}/{bad.size} out of tol, max err {err.max():.3e}, rel_l2 {rel_l2(a, b):.3e}"


def make_engine(args, sd, n_agents, n_scenes, **kw):
    from coalign_b200.engine import CoAlignEngine
    return CoAlignEngine(args, sd, n_agents, n_scenes, device="cuda", **kw)


def cuda_batch(inp):
    return (torch.from_numpy(inp["voxel_features"]).cuda(), torch.from_numpy(inp["voxel_coords"]).cuda(),
            torch.from_numpy(inp["voxel_num_points"]).cuda(), [int(v) for v in inp["record_len"]],
            torch.from_numpy(inp["pairwise_t_matrix"]).cuda())


def oracle_stages(args, sd, inp):
    from oracle import coalign_oracle as O
    st = {}
    out = O.forward(sd, args, G.to_torch_batch(inp), st)
    return out, st


def engine_stages(eng, n_img, n_scenes):
    s = {"canvas": eng.read_act(eng.canvas, n_img).cpu().numpy()}
    for i in range(len(eng.levels)):
        s[f"feat{i}"] = eng.read_act(eng.lvl[i]["out"], n_img).cpu().numpy()
        s[f"fused{i}"] = eng.read_act(eng.lvl[i]["fused"], n_scenes).cpu().numpy()
    s["decoded"] = eng.read_act(eng.cat, n_scenes).cpu().numpy()
    s["shrunk"] = eng.read_act(eng.shrink_bufs[-1], n_scenes).cpu().numpy()
    return s


@pytest.mark.parametrize("name,fusion", [("model_small_att", "att"), ("model_small_single", "att"),
                                         ("model_small_max", "max")])
def test_precise_mode_matches_reference_golden(name, fusion):
    """fp32-class path (bf16x3 split on tensor cores): heads within rtol 1e-3 of the reference's own outputs."""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    seed = int(g["seed"])
    args = G.small_args(fusion)
    sd = synth.random_state_dict(args, seed)
    rl = [int(v) for v in g["record_len"]]
    inp = G.small_case_inputs(rl, seed0=100 + seed)
    eng = make_engine(args, sd, sum(rl), len(rl), precise=True)
    out = eng.forward_voxels(*cuda_batch(inp))
    torch.cuda.synchronize()
    st = engine_stages(eng, sum(rl), len(rl))
    for i in range(3):
        assert_close(st[f"feat{i}"], g[f"feat{i}"], 1e-3, 1e-3, f"feat{i}")
        assert_close(st[f"fused{i}"], g[f"fused{i}"], 1e-3, 1e-3, f"fused{i}")
    assert_close(st["shrunk"], g["shrunk"], 1e-3, 1e-3, "shrunk")
    for k in ("cls_preds", "reg_preds", "dir_preds"):
        assert_close(out[k].cpu().numpy(), g[k], 1e-3, 1e-3, k)


def test_bf16_mode_tensor_core_vs_simt_and_oracle():
    """bf16 path: (1) tcgen05 kernel == SIMT evaluation of the same descriptors (same bf16 operands, fp32
    accumulate) up to accumulation order; (2) drift vs the fp32 oracle stays at bf16 level."""
    seed = 1
    args = G.small_args("att")
    sd = synth.random_state_dict(args, seed)
    rl = [3, 2]
    inp = G.small_case_inputs(rl, seed0=100 + seed)
    ref_out, ref_st = oracle_stages(args, sd, inp)
    res = {}
    for simt in (True, False):
        eng = make_engine(args, sd, 5, 2, precise=False, simt_conv=simt, use_graph=not simt)
        out = eng.forward_voxels(*cuda_batch(inp))
        torch.cuda.synchronize()
        res[simt] = ({k: v.cpu().numpy() for k, v in out.items()}, engine_stages(eng, 5, 2))
    # canvas: PFN is fp32 math rounded to bf16 once
    assert_close(res[False][1]["canvas"], ref_st["canvas"].numpy(), 1e-2, 1e-2, "canvas")
    for k in res[True][1]:
        assert rel_l2(res[False][1][k], res[True][1][k]) < 2e-2, (k, rel_l2(res[False][1][k], res[True][1][k]))
    for k in ("cls_preds", "reg_preds", "dir_preds"):
        assert rel_l2(res[False][0][k], res[True][0][k]) < 2e-2, k
        assert rel_l2(res[False][0][k], ref_out[k].numpy()) < 5e-2, (k, rel_l2(res[False][0][k], ref_out[k].numpy()))


@pytest.mark.parametrize("block_n,pair", [(64, False), (128, False), (256, False), (64, True), (128, True), (256, True)])
def test_conv_tile_shapes_agree(block_n, pair):
    """Every tcgen05 tile shape, single-CTA and CTA-pair (cta_group::2), against the SIMT evaluation."""
    seed = 2
    args = G.small_args("att")
    sd = synth.random_state_dict(args, seed)
    rl = [2]
    inp = G.small_case_inputs(rl, seed0=300)
    outs = []
    for bn, simt in ((block_n, False), (128, True)):
        eng = make_engine(args, sd, 2, 1, precise=True, block_n_cap=bn, simt_conv=simt, use_graph=False, pair=pair)
        eng.pair_min_bn = 64            # exercise the CTA-pair kernel at every tile width
        o = eng.forward_voxels(*cuda_batch(inp))
        torch.cuda.synchronize()
        outs.append({k: v.cpu().numpy() for k, v in o.items()})
    for k in outs[0]:
        assert_close(outs[0][k], outs[1][k], 1e-3, 1e-3, k)


def test_channel_major_conv_kernel_matches_pixel_major():
    """cb_conv_gemm_t (Cout=128 layers, weights on the UMMA M side, 256-pixel tiles) against cb_conv_gemm on the same
    descriptors: identical bf16 operands and fp32 accumulation, so the stages agree to accumulation order (a few bf16
    roundings flip).  Small config: every K-step kind (stride 2 + fused downsample, residual, PS output); full-size
    DAIR-V2X shape: odd 50x126 level, many tiles per CTA, partial last tile."""
    seed = 4
    args = G.small_args("att")
    sd = synth.random_state_dict(args, seed)
    inp = G.small_case_inputs([3, 2], seed0=100 + seed)
    res = {}
    for cm in (True, False):
        eng = make_engine(args, sd, 5, 2, precise=False, use_graph=False)
        eng.chan_major = cm
        out = eng.forward_voxels(*cuda_batch(inp))
        torch.cuda.synchronize()
        res[cm] = ({k: v.cpu().numpy() for k, v in out.items()}, engine_stages(eng, 5, 2))
    for k in res[True][1]:
        assert rel_l2(res[True][1][k], res[False][1][k]) < 3e-3, (k, rel_l2(res[True][1][k], res[False][1][k]))
    for k in res[True][0]:
        assert rel_l2(res[True][0][k], res[False][0][k]) < 3e-3, k
    args = synth.dairv2x_args()
    sd = synth.random_state_dict(args, 0)
    sc = synth.make_scene(21, 3, 60000, args["lidar_range"], pose_noise=True)
    pts = torch.from_numpy(np.concatenate(sc["points"])).cuda()
    off = (np.arange(4) * 60000).astype(np.int32)
    pw = torch.from_numpy(sc["pairwise_t_matrix"][None]).cuda()
    big = {}
    for cm in (True, False):
        eng = make_engine(args, sd, 3, 1, precise=False)
        eng.chan_major = cm
        out = eng.forward_points(pts, off, [3], pw)
        big[cm] = ({k: v.cpu().numpy() for k, v in out.items()}, engine_stages(eng, 3, 1))
        del eng
    for k in big[True][1]:
        assert rel_l2(big[True][1][k], big[False][1][k]) < 3e-3, (k, rel_l2(big[True][1][k], big[False][1][k]))
    for k in big[True][0]:
        assert rel_l2(big[True][0][k], big[False][0][k]) < 3e-3, k


def _run_small_and_big(configure):
    """Small config (every K-step kind: stride 2 + fused downsample, residual, PS output) and a full-size DAIR-V2X
    shape scene (odd 50x126 level, many tiles per CTA, partial last tile) through an engine set up by `configure`."""
    seed = 4
    args = G.small_args("att")
    sd = synth.random_state_dict(args, seed)
    inp = G.small_case_inputs([3, 2], seed0=100 + seed)
    eng = make_engine(args, sd, 5, 2, precise=False, use_graph=False)
    configure(eng)
    out = eng.forward_voxels(*cuda_batch(inp))
    torch.cuda.synchronize()
    small = ({k: v.cpu().numpy() for k, v in out.items()}, engine_stages(eng, 5, 2))
    del eng
    args = synth.dairv2x_args()
    sd = synth.random_state_dict(args, 0)
    sc = synth.make_scene(21, 3, 60000, args["lidar_range"], pose_noise=True)
    pts = torch.from_numpy(np.concatenate(sc["points"])).cuda()
    off = (np.arange(4) * 60000).astype(np.int32)
    pw = torch.from_numpy(sc["pairwise_t_matrix"][None]).cuda()
    eng = make_engine(args, sd, 3, 1, precise=False)
    configure(eng)
    out = eng.forward_points(pts, off, [3], pw)
    big = ({k: v.cpu().numpy() for k, v in out.items()}, engine_stages(eng, 3, 1))
    del eng
    return small, big


def test_halo_conv_kernels_match_plain_kernels():
    """cb_conv_gemm_halo (Cout=64: one 130-row TMA box per filter row, resident weights) and cb_conv_gemm_t_halo
    (Cout=128: 258-row pixel boxes) against the plain tcgen05 kernels on the same descriptors: identical bf16 operands
    and fp32 accumulation in the same K order, so every stage agrees to a few flipped bf16 roundings."""
    def plain(eng):
        eng.halo = False

    def halo(eng):
        eng.halo = True

    ref = _run_small_and_big(plain)
    got = _run_small_and_big(halo)
    for (r_out, r_st), (g_out, g_st) in zip(ref, got):
        for k in r_st:
            assert rel_l2(g_st[k], r_st[k]) < 3e-3, (k, rel_l2(g_st[k], r_st[k]))
        for k in r_out:
            assert rel_l2(g_out[k], r_out[k]) < 3e-3, (k, rel_l2(g_out[k], r_out[k]))


def test_channel_major_256_channel_layers_match_pair_kernel():
    """Cout = 256 layers through the channel-major kernels (two 128-channel work items per pixel tile; plain and halo
    variants, `chan_major_256` = 2) against the default CTA-pair kernel on the same descriptors."""
    def pair(eng):
        eng.chan_major_256 = 0

    def cm(eng):
        eng.chan_major_256 = 2

    ref = _run_small_and_big(pair)
    got = _run_small_and_big(cm)
    for (r_out, r_st), (g_out, g_st) in zip(ref, got):
        for k in r_st:
            assert rel_l2(g_st[k], r_st[k]) < 3e-3, (k, rel_l2(g_st[k], r_st[k]))
        for k in r_out:
            assert rel_l2(g_out[k], r_out[k]) < 3e-3, (k, rel_l2(g_out[k], r_out[k]))


def test_voxelize_bit_exact_and_fused_path():
    """Integer pillar indices bit-exact with the serial generator restatement (oracle/voxelize.c);
    fused points->canvas == voxels->canvas."""
    from oracle import voxelize_np as V
    args = G.small_args("att")
    sd = synth.random_state_dict(args, 5)
    eng = make_engine(args, sd, 4, 1, precise=True)
    rng = np.random.default_rng(11)
    clouds = [synth.lidar_cloud(rng, n, G.SMALL_RANGE, sigma=s) for n, s in ((3000, 5.0), (1, 5.0), (5000, 1.5), (257, 8.0))]
    edge = np.array([[-11.2, -4.8, -3.0, 0.5], [11.2, 0, 0, 0.5], [0, 4.8, 0, 0.5], [0, 0, 1.0, 0.5],
                     [0.4, 0.8, -1.0, 0.1], [11.199999, 4.799999, 0.999, 0.2], [-11.2000001, 0, 0, 0.3]], np.float32)
    clouds[0] = np.concatenate([edge, clouds[0]])
    off = np.concatenate([[0], np.cumsum([c.shape[0] for c in clouds])]).astype(np.int32)
    pts = torch.from_numpy(np.concatenate(clouds)).cuda()
    for max_pts, max_vox in ((32, 70000), (5, 70000), (32, 150)):
        vox, crd, npt, nv = eng.voxelize(pts, off, max_pts, max_vox)
        ref = V.collate([V.voxelize_c(c, G.SMALL_RANGE, G.SMALL_VOXEL, max_pts, max_vox) for c in clouds])
        assert np.array_equal(crd.cpu().numpy(), ref[1]), (max_pts, max_vox)
        assert np.array_equal(npt.cpu().numpy(), ref[2])
        assert np.array_equal(vox.cpu().numpy(), ref[0])          # bit-exact incl. slot order and zero padding
    # fused path (v2 front-end: arrival-order slots, caps handled by the rare-path kernels): same canvas as the staged
    # path, bit for bit (max is order-independent, the pillar mean is accumulated in fp64), incl. the 32-point cap on the
    # dense cloud (sigma 1.5 m: hundreds of points per cell), a 5-point cap and a 150-voxel cap
    pw = torch.from_numpy(np.tile(np.eye(4), (1, 5, 5, 1, 1))).cuda()
    for max_pts, max_vox in ((32, 70000), (5, 70000), (32, 150), (7, 40)):
        vox, crd, npt, nv = eng.voxelize(pts, off, max_pts, max_vox)
        assert max_vox < 70000 or int(npt.max()) == max_pts               # the point cap is exercised
        eng.forward_voxels(vox, crd, npt, [4], pw)
        c1 = eng.read_act(eng.canvas, 4).cpu().numpy()
        for rep in range(2):                                              # twice: sparse canvas clear between frames
            eng.forward_points(pts, off, [4], pw, max_pts, max_vox)
            c2 = eng.read_act(eng.canvas, 4).cpu().numpy()
            assert np.array_equal(c1, c2), (max_pts, max_vox, rep, np.abs(c1 - c2).max(), (c1 != c2).sum())


@pytest.mark.parametrize("version", [9, 8, 1])
def test_warp_fuse_op_golden(version):
    """A11/A12 op-level: fused warp+attention kernel vs the reference's AttFusion/MaxFusion on 64-channel maps - the tiled
    v9 kernel (product path) and the two earlier kernels it falls back to (v8: buffers that are not 32-byte aligned or
    too large for 32-bit byte offsets; v1: channel counts other than 64/128/256), selected through cb_set_option."""
    from coalign_b200 import _lib
    from oracle import coalign_oracle as O
    lib = _lib.load(True)
    prev = lib.cb_set_option(_lib.CB_OPT_FUSE_VERSION, version)
    try:
        _warp_fuse_op_case(lib)
    finally:
        lib.cb_set_option(_lib.CB_OPT_FUSE_VERSION, prev)
    assert lib.cb_get_option(_lib.CB_OPT_FUSE_VERSION) == prev


def _warp_fuse_op_case(lib):
    from coalign_b200 import _lib
    from oracle import coalign_oracle as O
    g = torch.Generator().manual_seed(3)
    H, W, Cc = 9, 14, 64
    x = torch.randn(5, Cc, H, W, generator=g)
    aff = torch.zeros(2, 5, 5, 2, 3, dtype=torch.float64)
    aff[..., 0, 0] = 1
    aff[..., 1, 1] = 1
    aff[0, 0, 1] = torch.tensor([[0.8, -0.5, 0.2], [0.6, 0.9, -0.3]])
    aff[0, 0, 2] = torch.tensor([[-1.0, 0.05, 1.7], [0.02, -1.0, 0.4]])
    aff[1, 0, 1] = torch.tensor([[1.0, 0, 5.0], [0, 1.0, 5.0]])     # fully out of view
    sp = torch.cuda.current_stream().cuda_stream
    for ps in (0, 1):
        for method, name in ((0, "att"), (1, "max")):
            ref = O.att_fusion(x, [3, 2], aff, name).numpy()
            Hq, Wq = ((H + 1) // 2 + 2, (W + 1) // 2 + 2) if ps else (H + 2, W + 2)
            rows = 5 * Hq * Wq * (4 if ps else 1)
            buf = torch.zeros(2 * rows, Cc, dtype=torch.bfloat16, device="cuda")
            xd = x.cuda().contiguous()
            _lib.check(lib.cb_nchw_to_layout(xd.data_ptr(), 5, Cc, H, W, ps, buf.data_ptr(), rows * Cc, sp))
            out = torch.zeros(2 * 2 * (H + 2) * (W + 2), Cc, dtype=torch.bfloat16, device="cuda")
            affd = aff[:, 0].contiguous().cuda()
            offd = torch.tensor([0, 3, 5], dtype=torch.int32, device="cuda")
            _lib.check(lib.cb_warp_att_fuse(buf.data_ptr(), ps, rows * Cc, 5, affd.data_ptr(), offd.data_ptr(), 2, 5,
                                            H, W, Cc, method, out.data_ptr(), 2 * (H + 2) * (W + 2) * Cc, sp))
            dense = torch.empty(2, Cc, H, W, device="cuda")
            _lib.check(lib.cb_layout_to_nchw(out.data_ptr(), 2 * (H + 2) * (W + 2) * Cc, 0, 2, Cc, H, W, Cc, 0,
                                             dense.data_ptr(), sp))
            torch.cuda.synchronize()
            assert_close(dense.cpu().numpy(), ref, 1e-4, 1e-4, f"fuse ps={ps} {name}")


def test_normalize_affine_matches_reference_golden():
    from coalign_b200 import _lib
    lib = _lib.load(True)
    g = np.load(os.path.join(GOLD, "ops.npz"))
    pw = torch.from_numpy(g["pairwise"][None].copy()).cuda()
    out = torch.zeros(1, 5, 2, 3, dtype=torch.float64, device="cuda")
    _lib.check(lib.cb_normalize_affine(pw.data_ptr(), 1, 5, 200, 704, 0.4, out.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    np.testing.assert_allclose(out.cpu().numpy(), g["affine_200_704"][:, 0], rtol=0, atol=1e-13)


def test_plugin_through_reference_style_call():
    """The nn.Module twin: state_dict round trip + forward(data_dict) with the reference batch schema."""
    from coalign_b200.model import PointPillarCoalignB200
    g = np.load(os.path.join(GOLD, "model_small_att.npz"))
    seed = int(g["seed"])
    args = G.small_args("att")
    args["b200_precise"] = True
    sd = synth.random_state_dict(args, seed)
    m = PointPillarCoalignB200(args)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    inp = G.small_case_inputs([3, 2], seed0=100 + seed)
    batch = G.to_torch_batch(inp)
    batch = {"processed_lidar": {k: v.cuda() for k, v in batch["processed_lidar"].items()},
             "record_len": batch["record_len"].cuda(), "pairwise_t_matrix": batch["pairwise_t_matrix"].cuda()}
    with torch.no_grad():
        out = m(batch)
    for k in ("cls_preds", "reg_preds", "dir_preds"):
        assert_close(out[k].cpu().numpy(), g[k], 1e-3, 1e-3, k)
    # .train() runs the device training path (tests/test_train_gpu.py); the single-agent twins stay inference-only
    from coalign_b200.model import PointPillarB200
    single = PointPillarB200(synth.single_args(G.SMALL_RANGE, G.SMALL_VOXEL)).cuda()
    with pytest.raises(NotImplementedError):
        single.train()(batch)


def test_ragged_batches_and_empty_agent_same_engine():
    """One engine, changing batch signatures (graph re-capture), an agent whose cloud is entirely out of range,
    and a 1-point cloud: every result equals the oracle's."""
    from oracle import coalign_oracle as O
    args = G.small_args("att")
    sd = synth.random_state_dict(args, 4)
    eng = make_engine(args, sd, 8, 3, precise=True)
    for case, (rl, seed0) in enumerate((([3, 2], 500), ([1, 5, 2], 600), ([2], 700), ([3, 2], 800))):
        scenes = G.small_case_scenes(rl, seed0)
        if case == 1:
            scenes[1]["points"][2] = scenes[1]["points"][2] + np.array([1000, 0, 0, 0], np.float32)   # all out of range
            scenes[1]["points"][3] = scenes[1]["points"][3][:1]                                       # single point
        inp = G.scenes_to_batch(scenes, G.SMALL_RANGE, G.SMALL_VOXEL)
        ref = O.forward(sd, args, G.to_torch_batch(inp))
        pts = np.concatenate([p for sc in scenes for p in sc["points"]]).astype(np.float32)
        off = np.concatenate([[0], np.cumsum([p.shape[0] for sc in scenes for p in sc["points"]])]).astype(np.int32)
        pw = torch.from_numpy(inp["pairwise_t_matrix"]).cuda()
        out_p = eng.forward_points(torch.from_numpy(pts).cuda(), off, rl, pw)
        out_v = eng.forward_voxels(*cuda_batch(inp))
        torch.cuda.synchronize()
        for k in ref:
            assert_close(out_p[k].cpu().numpy(), ref[k].numpy(), 1e-3, 1e-3, f"case{case} points {k}")
            assert_close(out_v[k].cpu().numpy(), ref[k].numpy(), 1e-3, 1e-3, f"case{case} voxels {k}")


def test_full_size_properties_opv2v():
    """BASELINE-size (200x704 canvas, 60k points/agent) size-independent properties, no oracle needed:
    (a) permuting the non-ego agents leaves the fused output unchanged; (b) N identical co-located copies of the
    ego cloud give the single-agent result (uniform attention over identical vectors)."""
    args = synth.opv2v_args()
    sd = synth.random_state_dict(args, 0)
    eng = make_engine(args, sd, 3, 1, precise=True)
    sc = synth.make_scene(7, 3, 60000, args["lidar_range"], pose_noise=True)
    P = 60000
    off = np.array([0, P, 2 * P, 3 * P], np.int32)

    def run(order, pw):
        pts = torch.from_numpy(np.concatenate([sc["points"][i] for i in order])).cuda()
        o = eng.forward_points(pts, off, [3], torch.from_numpy(pw[None]).cuda())
        return {k: v.cpu().numpy() for k, v in o.items()}

    pw = sc["pairwise_t_matrix"]
    a = run([0, 1, 2], pw)
    pw_sw = pw.copy()
    perm = [0, 2, 1, 3, 4]
    pw_sw = pw[np.ix_(perm, perm)]
    b = run([0, 2, 1], pw_sw)
    for k in a:
        assert_close(b[k], a[k], 1e-3, 1e-3, f"permutation {k}")
    # (b) three identical agents at the ego pose vs one agent
    ident = np.tile(np.eye(4), (5, 5, 1, 1))
    pts3 = torch.from_numpy(np.concatenate([sc["points"][0]] * 3)).cuda()
    o3 = eng.forward_points(pts3, off, [3], torch.from_numpy(ident[None]).cuda())
    o3 = {k: v.cpu().numpy() for k, v in o3.items()}
    o1 = eng.forward_points(torch.from_numpy(sc["points"][0]).cuda(), np.array([0, P], np.int32), [1],
                            torch.from_numpy(ident[None]).cuda())
    for k in o3:
        assert_close(o3[k], o1[k].cpu().numpy(), 1e-3, 1e-3, f"identical agents {k}")
    assert all(np.isfinite(v).all() for v in o3.values())


def test_pipelined_runner_matches_direct_forward():
    """Public serving API (pinned host in/out, overlapped copies): same numbers as a direct forward, for a stream of
    different batches."""
    from coalign_b200.runtime import PipelinedRunner
    args = G.small_args("att")
    sd = synth.random_state_dict(args, 6)
    eng = make_engine(args, sd, 5, 2, precise=False)
    batches = []
    for s in range(5):
        scenes = G.small_case_scenes([3, 2], 900 + 10 * s)
        pts = np.concatenate([p for sc in scenes for p in sc["points"]]).astype(np.float32)
        off = np.concatenate([[0], np.cumsum([p.shape[0] for sc in scenes for p in sc["points"]])]).astype(np.int32)
        pw = np.stack([sc["pairwise_t_matrix"] for sc in scenes])
        batches.append((torch.from_numpy(pts).pin_memory(), off, torch.from_numpy(pw).pin_memory()))
    direct = []
    for pts, off, pw in batches:
        o = eng.forward_points(pts.cuda(), off, [3, 2], pw.cuda())
        direct.append({k: v.cpu().numpy() for k, v in o.items()})
    runner = PipelinedRunner(eng, max_points=int(max(b[1][-1] for b in batches)))
    got, prev = [], None
    for pts, off, pw in batches:
        t = runner.submit(pts, off, [3, 2], pw)
        if prev is not None:
            got.append({k: v.clone().numpy() for k, v in runner.result(prev).items()})
        prev = t
    got.append({k: v.clone().numpy() for k, v in runner.result(prev).items()})
    runner.drain()
    for a, b in zip(got, direct):
        for k in a:
            assert np.array_equal(a[k], b[k]), k


def _full_size_case(args, n_agents, seed, pose_noise):
    from oracle import coalign_oracle as O
    torch.set_num_threads(min(32, os.cpu_count() or 1))
    sd = synth.random_state_dict(args, 0)
    sc = synth.make_scene(seed, n_agents, 60000, args["lidar_range"], pose_noise=pose_noise)
    inp = G.scenes_to_batch([sc], args["lidar_range"], args["voxel_size"])
    ref = O.forward(sd, args, G.to_torch_batch(inp))
    pts = torch.from_numpy(np.concatenate(sc["points"])).cuda()
    off = (np.arange(n_agents + 1) * 60000).astype(np.int32)
    pw = torch.from_numpy(sc["pairwise_t_matrix"][None]).cuda()
    res = {}
    for precise in (True, False):
        eng = make_engine(args, sd, n_agents, 1, precise=precise)
        out = eng.forward_points(pts, off, [n_agents], pw)
        res[precise] = {k: v.cpu().numpy() for k, v in out.items()}
        del eng
    for k in ref:
        # ~35 stacked layers at K up to 10368: the tensor cores' truncating fp32 accumulation leaves rel-L2 ~9e-5 and a
        # handful of elements (2 of 70400 measured) at 1.7e-3*rms, hence the slightly wider absolute floor at full size
        assert_close(res[True][k], ref[k].numpy(), 1e-3, 2.5e-3, f"precise {k}")
        assert rel_l2(res[True][k], ref[k].numpy()) < 2e-4, (k, rel_l2(res[True][k], ref[k].numpy()))
        assert rel_l2(res[False][k], ref[k].numpy()) < 3e-2, (k, rel_l2(res[False][k], ref[k].numpy()))
    if tuple(res[True]["cls_preds"].shape[1:]) == (2, 100, 352):
        _detection_level_agreement(res, ref)


def _detection_level_agreement(res, ref):
    """What the bf16 head maps mean for detections: decode + NMS (cb_postprocess, voxel_postprocessor.py:292-400) of the
    reference's, the precise and the bf16 head maps with the same score offset (random weights give no confident anchors by
    themselves: the offset puts the 300 best reference anchors above the score threshold).  Every box kept from the
    reference maps must have a partner in the GPU result: centre within 0.3 m (the two anchors of a cell share a centre),
    score within 0.05; and the GPU result may not invent more than a tenth more boxes."""
    from coalign_b200.postprocess import VoxelPostprocessorB200
    params = G.post_params()
    pp = VoxelPostprocessorB200(params, train=False)
    anchors = pp.generate_anchor_box()
    thr = float(params["target_args"]["score_threshold"])
    # random-weight head maps are not on the scale of a trained head (logit std of tens): the same scale for all three
    # brings the logits to std 2 and - below - the regression deltas to std 0.3
    cscale = np.float32(2.0 / float(ref["cls_preds"].numpy().std()))
    ref_cls = ref["cls_preds"].numpy() * cscale
    kth = np.sort(ref_cls.reshape(-1))[-300]
    shift = float(np.log(thr / (1.0 - thr)) - kth)
    rscale = np.float32(0.3 / float(ref["reg_preds"].numpy().std()))
    tfm = torch.eye(4)

    def run(o):
        cls = torch.from_numpy(np.asarray(o["cls_preds"]) * cscale + shift).cuda()
        out = pp.post_process_batch(cls, torch.from_numpy(np.asarray(o["reg_preds"]) * rscale).cuda(),
                                    torch.from_numpy(np.asarray(o["dir_preds"])).cuda(), anchors, tfm)[0]
        assert out[0] is not None
        return out[0].cpu().numpy().mean(axis=1), out[1].cpu().numpy()       # box centres (K,3), scores (K,)

    want_c, want_s = run({k: v.numpy() for k, v in ref.items()})
    assert want_c.shape[0] >= 5, want_c.shape
    stats = {}
    for mode in (True, False):
        got_c, got_s = run(res[mode])
        d = np.linalg.norm(want_c[:, None, :2] - got_c[None, :, :2], axis=2)
        j = d.argmin(axis=1)
        ok = (d[np.arange(len(j)), j] < 0.3) & (np.abs(got_s[j] - want_s) < 0.05)
        stats["precise" if mode else "bf16"] = (int((~ok).sum()), len(want_s), len(got_s))
        # measured: the only reference boxes without a partner sit on the score threshold (scores 0.200-0.208 at 0.20)
        assert all(want_s[i] < thr + 0.02 for i in np.where(~ok)[0]), [float(want_s[i]) for i in np.where(~ok)[0]]
    # an anchor sitting exactly on the score threshold may flip in any arithmetic: allow 2 boxes in precise mode;
    # bf16 mode (head logits rel-L2 1-2e-2): at least 90 % of the reference's boxes, at most 10 % + 2 extra ones
    assert stats["precise"][0] <= 2, stats
    assert stats["bf16"][0] <= max(2, 0.1 * stats["bf16"][1]), stats
    for k in stats:
        assert stats[k][2] <= 1.1 * stats[k][1] + 2, stats
    print("detection-level agreement (missed, reference boxes, ours):", stats)


def test_full_size_opv2v_two_agents_vs_oracle():
    """BASELINE configs[1]: OPV2V-shape 2-agent intermediate fusion, no pose noise, full 200x704 canvas, 60k points."""
    _full_size_case(synth.opv2v_args(), 2, seed=11, pose_noise=False)


def test_full_size_dairv2x_two_agents_pose_noise_vs_oracle():
    """BASELINE configs[3]: DAIR-V2X shape (200x504 canvas, voxel z 5 m, odd 25x63 top level), 2 agents, pose noise."""
    _full_size_case(synth.dairv2x_args(), 2, seed=12, pose_noise=True)


def test_full_size_opv2v_five_agents_vs_oracle():
    """BASELINE configs[2], the benchmarked configuration: 5 agents, pose noise, full 200x704 canvas, 60k points/agent."""
    _full_size_case(synth.opv2v_args(), 5, seed=13, pose_noise=True)


def test_bench_launch_60_agents_equals_single_scene_launches():
    """The bench runs 12 scenes x 5 agents in ONE launch sequence.  Scene by scene it must give what a single-scene launch
    gives: bit for bit in bf16 mode (every output pixel sees the same operands in the same K order whatever tile it falls
    into), and the first scene additionally within bf16 drift of the fp32 oracle."""
    from oracle import coalign_oracle as O
    torch.set_num_threads(min(32, os.cpu_count() or 1))
    args = synth.opv2v_args()
    sd = synth.random_state_dict(args, 0)
    B, NA, P = 12, 5, 60000
    scenes = [synth.make_scene(2000 + s, NA, P, args["lidar_range"], max_cav=5, pose_noise=True) for s in range(B)]
    pts = torch.from_numpy(np.concatenate([p for sc in scenes for p in sc["points"]]).astype(np.float32)).cuda()
    pw = torch.from_numpy(np.stack([sc["pairwise_t_matrix"] for sc in scenes])).cuda()
    off = (np.arange(B * NA + 1) * P).astype(np.int32)
    eng = make_engine(args, sd, B * NA, B, precise=False, block_n_cap=256)
    big = {k: v.cpu().numpy() for k, v in eng.forward_points(pts, off, [NA] * B, pw).items()}
    for s in (0, 5, 11):
        one = eng.forward_points(pts[s * NA * P:(s + 1) * NA * P].contiguous(), off[:NA + 1], [NA], pw[s:s + 1])
        for k in big:
            assert np.array_equal(one[k][0].cpu().numpy(), big[k][s]), (s, k)
    inp = G.scenes_to_batch([scenes[0]], args["lidar_range"], args["voxel_size"])
    ref = O.forward(sd, args, G.to_torch_batch(inp))
    for k in ref:
        assert rel_l2(big[k][0], ref[k][0].numpy()) < 3e-2, (k, rel_l2(big[k][0], ref[k][0].numpy()))


def test_bf16_mode_against_bf16_rounded_interpreter():
    """Like-for-like check of the production (bf16) mode: the launch-plan interpreter (tests/plan_interpreter.py) evaluates
    the SAME descriptors on the CPU with the same bf16-packed weights and a bf16 rounding at every activation store, fp32
    accumulation in between.  What is left between it and the GPU is accumulation order inside a K loop, the packed-bf16
    tap blend of the fusion kernel and roundings that flip on ties: well below the 5e-2 drift bound against the fp32 oracle
    (roundings that flip on ties are amplified by the 16 stacked layers of the deepest level: 7e-3 measured there)."""
    from coalign_b200.engine import CoAlignEngine
    from tests import plan_interpreter as PI
    seed = 1
    args = G.small_args("att")
    sd = synth.random_state_dict(args, seed)
    rl = [3, 2]
    inp = G.small_case_inputs(rl, seed0=100 + seed)
    cpu_eng = CoAlignEngine(args, sd, 5, 2, device="cpu", precise=False, plan_only=True)
    ref = PI.run_plan(cpu_eng, sd, args, G.to_torch_batch(inp))
    eng = make_engine(args, sd, 5, 2, precise=False)
    out = eng.forward_voxels(*cuda_batch(inp))
    torch.cuda.synchronize()
    for i in range(3):
        got = eng.read_act(eng.lvl[i]["out"], 5).cpu().numpy()
        want = PI.act_to_nchw(cpu_eng.lvl[i]["out"], cpu_eng.lvl[i]["out"].n_cap)[:5].numpy()
        assert rel_l2(got, want) < 1.2e-2, (f"feat{i}", rel_l2(got, want))      # measured 7e-3 at the deepest level
    for k in ("cls_preds", "reg_preds", "dir_preds"):
        r = rel_l2(out[k].cpu().numpy(), ref[k].numpy())
        assert r < 1.5e-2, (k, r)


def test_varying_cloud_sizes_reuse_one_graph():
    """Frames whose clouds differ in size (real sweeps do): the per-agent offsets travel through a device array, so the
    second and later frames replay the graph captured for the first, and every frame equals the result of the
    host-offset entry point (cb_points_to_canvas + voxel-tensor path) on the same clouds."""
    args = G.small_args("att")
    sd = synth.random_state_dict(args, 8)
    eng = make_engine(args, sd, 5, 2, precise=False)
    ref_eng = make_engine(args, sd, 5, 2, precise=False, use_graph=False)
    rng = np.random.default_rng(0)
    cap = None
    for frame in range(5):
        scenes = G.small_case_scenes([3, 2], 700 + 10 * frame, n_points=1800)
        clouds = [p[:int(rng.integers(900, 1801))] for sc in scenes for p in sc["points"]]
        if frame == 3:
            clouds[1] = clouds[1][:0]                                   # an agent with an empty cloud
        off = np.concatenate([[0], np.cumsum([c.shape[0] for c in clouds])]).astype(np.int32)
        flat = np.concatenate(clouds).astype(np.float32)
        if cap is None:
            cap = torch.zeros(5 * 1800, 4, device="cuda")               # the serving loop's staging buffer: fixed capacity
        cap[:flat.shape[0]] = torch.from_numpy(flat).cuda()
        pw = torch.from_numpy(np.stack([sc["pairwise_t_matrix"] for sc in scenes])).cuda()
        out = {k: v.clone() for k, v in eng.forward_points(cap, off, [3, 2], pw).items()}
        n_graphs = len(eng._graphs)
        assert n_graphs == 1, n_graphs
        # host-offset reference: reference-format voxel tensors from cb_voxelize, then the voxel entry
        vf, vc, vn, _ = ref_eng.voxelize(torch.from_numpy(flat).cuda(), off)
        want = ref_eng.forward_voxels(vf, vc, vn, [3, 2], pw)
        torch.cuda.synchronize()
        for k in out:
            assert torch.equal(out[k], want[k]), (frame, k)


def test_two_engines_with_different_pfn_weights_on_concurrent_streams():
    """The fused points -> canvas path keeps its PFN coefficients in a constant-memory table; each engine (workspace) has its
    own table slot, so two engines with different weights replaying their graphs on different streams at the same time
    must each reproduce their own single-stream result, every time."""
    args = G.small_args("att")
    scenes = G.small_case_scenes([3, 2], 900, n_points=1800)
    clouds = [p for sc in scenes for p in sc["points"]]
    off = np.concatenate([[0], np.cumsum([c.shape[0] for c in clouds])]).astype(np.int32)
    pts = torch.from_numpy(np.concatenate(clouds).astype(np.float32)).cuda()
    pw = torch.from_numpy(np.stack([sc["pairwise_t_matrix"] for sc in scenes])).cuda()
    engs, want, streams = [], [], []
    for seed in (11, 12):
        eng = make_engine(args, synth.random_state_dict(args, seed), 5, 2, precise=False)
        eng.forward_points(pts, off, [3, 2], pw)                       # captures the graph
        want.append({k: v.clone() for k, v in eng.forward_points(pts, off, [3, 2], pw).items()})
        engs.append(eng)
        streams.append(torch.cuda.Stream())
    torch.cuda.synchronize()
    assert not torch.equal(want[0]["cls_preds"], want[1]["cls_preds"])
    for _ in range(20):
        got = []
        for eng, st in zip(engs, streams):
            with torch.cuda.stream(st):
                got.append({k: v.clone() for k, v in eng.forward_points(pts, off, [3, 2], pw).items()})
        torch.cuda.synchronize()
        for g, w in zip(got, want):
            for k in w:
                assert torch.equal(g[k], w[k]), k


def test_single_agent_pointpillar_matches_reference_golden():
    """BASELINE configs[0]: single-agent `point_pillar` (BaseBEVBackbone, no fusion) through the nn.Module twin:
    precise mode within rtol 1e-3 of the unmodified reference's outputs; bf16 mode at bf16-level drift; raw-point entry
    == voxel entry."""
    from coalign_b200.model import PointPillarB200
    g = np.load(os.path.join(GOLD, "model_single_plain.npz"))
    seed, n = int(g["seed"]), int(g["n_frames"])
    inp = G.single_case_inputs(n, seed0=100 + seed)
    batch = G.to_torch_batch(inp)
    dev = {"processed_lidar": {k: v.cuda() for k, v in batch["processed_lidar"].items()}}
    outs = {}
    for precise in (True, False):
        args = synth.single_args(G.SMALL_RANGE, G.SMALL_VOXEL)
        args["b200_precise"] = precise
        sd = synth.random_state_dict(args, seed, backbone="plain")
        m = PointPillarB200(args)
        m.load_state_dict(sd, strict=True)
        m = m.cuda().eval()
        with torch.no_grad():
            out = m(dev)
        torch.cuda.synchronize()
        outs[precise] = {k: v.cpu().numpy() for k, v in out.items()}
        if precise:
            eng = m.engine(n)
            for i in range(3):
                assert_close(eng.read_act(eng.lvl[i]["out"], n).cpu().numpy(), g[f"feat{i}"], 1e-3, 1e-3, f"feat{i}")
            assert_close(eng.read_act(eng.cat, n).cpu().numpy(), g["decoded"], 1e-3, 1e-3, "decoded")
            pts = torch.from_numpy(np.concatenate(inp["points"]).astype(np.float32)).cuda()
            off = np.concatenate([[0], np.cumsum([p.shape[0] for p in inp["points"]])]).astype(np.int32)
            out_p = m.forward_points(pts, off)
            for k in out:
                assert_close(out_p[k].cpu().numpy(), outs[True][k], 1e-5, 1e-5, f"points vs voxels {k}")
    for k in ("cls_preds", "reg_preds", "dir_preds"):
        assert_close(outs[True][k], g[k], 1e-3, 1e-3, k)
        assert rel_l2(outs[False][k], g[k]) < 5e-2, (k, rel_l2(outs[False][k], g[k]))


def test_stage1_uncertainty_detector_matches_reference_golden():
    """the stage-1 `point_pillar_uncertainty` detector through the nn.Module twin - precise mode within
    rtol 1e-3 of the unmodified reference's cls/reg/unc/dir outputs, bf16 mode at bf16-level drift, raw-point entry ==
    voxel entry."""
    from coalign_b200.model import PointPillarUncertaintyB200
    g = np.load(os.path.join(GOLD, "model_single_uncertainty.npz"))
    seed, n = int(g["seed"]), int(g["n_frames"])
    inp = G.single_case_inputs(n, seed0=100 + seed)
    batch = G.to_torch_batch(inp)
    dev = {"processed_lidar": {k: v.cuda() for k, v in batch["processed_lidar"].items()}}
    outs = {}
    for precise in (True, False):
        args = synth.uncertainty_args(G.SMALL_RANGE, G.SMALL_VOXEL)
        args["b200_precise"] = precise
        sd = synth.random_state_dict(args, seed, backbone="plain")
        m = PointPillarUncertaintyB200(args)
        m.load_state_dict(sd, strict=True)
        m = m.cuda().eval()
        with torch.no_grad():
            out = m(dev)
        torch.cuda.synchronize()
        assert list(out) == ["cls_preds", "reg_preds", "unc_preds", "dir_preds"]
        outs[precise] = {k: v.cpu().numpy() for k, v in out.items()}
        if precise:
            eng = m.engine(n)
            assert_close(eng.read_act(eng.cat, n).cpu().numpy(), g["decoded"], 1e-3, 1e-3, "decoded")
            pts = torch.from_numpy(np.concatenate(inp["points"]).astype(np.float32)).cuda()
            off = np.concatenate([[0], np.cumsum([p.shape[0] for p in inp["points"]])]).astype(np.int32)
            out_p = m.forward_points(pts, off)
            for k in out:
                assert_close(out_p[k].cpu().numpy(), outs[True][k], 1e-5, 1e-5, f"points vs voxels {k}")
    for k in ("cls_preds", "reg_preds", "unc_preds", "dir_preds"):
        assert_close(outs[True][k], g[k], 1e-3, 1e-3, k)
        assert rel_l2(outs[False][k], g[k]) < 5e-2, (k, rel_l2(outs[False][k], g[k]))
