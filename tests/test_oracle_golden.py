"""CPU: the oracle replays the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py) -- this is what pins the oracle when /root/reference is not around."""
import numpy as np
import torch

import eslam_oracle as O
from conftest import GOLDEN_CAM, TRUNC, golden_field, load_npz, recorded_draws, rel_err

CAM = O.Camera(*GOLDEN_CAM)
RC = O.RenderCfg(32, 8, TRUNC)


def test_pose_conversions_golden():
    d = load_npz("pose.npz")
    poses = torch.from_numpy(d["poses"])
    M = O.cam_pose_to_matrix(poses)
    assert torch.equal(M, torch.from_numpy(d["mats"]))
    assert torch.equal(O.matrix_to_cam_pose(M), torch.from_numpy(d["back"]))


def test_quaternion_against_scipy():
    from scipy.spatial.transform import Rotation

    g = torch.Generator().manual_seed(5)
    q = torch.randn(200, 4, generator=g) * 3.7
    R = O.quaternion_to_matrix(q)
    Rs = Rotation.from_quat(q[:, [1, 2, 3, 0]].numpy()).as_matrix()
    assert rel_err(R, torch.from_numpy(Rs).float()) < 2e-6
    back = O.matrix_to_quaternion(R)
    qn = q / q.norm(dim=-1, keepdim=True)
    sign = torch.sign((back * qn).sum(-1, keepdim=True))
    assert rel_err(back * sign, qn) < 1e-5
    # round trip identity
    assert rel_err(O.quaternion_to_matrix(back), R) < 1e-5


def test_decoders_golden():
    fld, d = golden_field(), load_npz("decoders.npz")
    pts = torch.from_numpy(d["pts"])
    assert rel_err(O.decode(pts.clone(), fld), d["raw"]) < 1e-6
    pn = O.normalize_pts(pts.clone(), fld.bound)
    assert rel_err(O.plane_features(pn, *fld.planes[:3]), d["feat_sdf"]) < 1e-6


def test_render_golden_forward_backward():
    fld, d = golden_field().clone(requires_grad=True), load_npz("render.npz")
    ro = torch.from_numpy(d["rays_o"]).requires_grad_(True)
    rd = torch.from_numpy(d["rays_d"]).requires_grad_(True)
    gt = torch.from_numpy(d["gt_depth"])
    depth, rgb, sdf, z = O.render_rays(fld, ro, rd, gt, TRUNC, 32, 8, O.ReplayDraws(recorded_draws(d)))
    has = gt > 0
    assert torch.equal(z[has], torch.from_numpy(d["z"])[has])
    assert rel_err(z, d["z"]) < 1e-6 and rel_err(depth, d["depth"]) < 1e-6 and rel_err(rgb, d["rgb"]) < 1e-6
    (depth * torch.from_numpy(d["g_depth"])).sum().add((rgb * torch.from_numpy(d["g_rgb"])).sum()).add(
        (sdf * torch.from_numpy(d["g_sdf"])).sum()).backward()
    assert rel_err(ro.grad, d["d_rays_o"]) < 1e-5 and rel_err(rd.grad, d["d_rays_d"]) < 1e-5
    for k, leaf in enumerate(fld.leaves()[:12]):
        assert rel_err(leaf.grad, d[f"d_plane.{k}"]) < 1e-5
    for name in O.DECODER_KEYS:
        assert rel_err(fld.dec[name].grad, d[f"d_dec.{name}"]) < 1e-5
    assert rel_err(fld.beta.grad, d["d_dec.beta"]) < 1e-5


def test_tracking_golden():
    fld, d = golden_field(), load_npz("tracking.npz")
    best, final, losses = O.track_frame(fld, CAM, RC, O.TRACK_W, torch.from_numpy(d["pose0"]),
                                        torch.from_numpy(d["gt_color"]), torch.from_numpy(d["gt_depth"]),
                                        int(d["n_pix"]), int(d["edge_h"]), int(d["edge_w"]), int(d["iters"]),
                                        float(d["lr_T"]), float(d["lr_R"]), O.ReplayDraws(recorded_draws(d)))
    assert rel_err(torch.tensor(losses), d["losses"]) < 1e-6
    assert rel_err(final, d["pose_trace"][-1:]) < 1e-6


def test_mapping_golden():
    fld, d = golden_field(), load_npz("mapping.npz")
    c2ws, losses = O.map_window(fld, CAM, RC, O.MAP_W, torch.from_numpy(d["c2ws0"]), torch.from_numpy(d["gt_colors"]),
                                torch.from_numpy(d["gt_depths"]), int(d["n_pixels"]), int(d["iters"]), 0.001, 0.005,
                                0.005, True, 0.001, O.ReplayDraws(recorded_draws(d)))
    assert rel_err(c2ws, d["c2ws_after"]) < 1e-6
    names = ("xy", "xz", "yz", "c_xy", "c_xz", "c_yz")
    for n, g in zip(names, fld.planes):
        for s in range(2):
            assert rel_err(g[s], d[f"after.plane.{n}.{s}"]) < 1e-6
    for name in O.DECODER_KEYS:
        assert rel_err(fld.dec[name], d[f"after.dec.{name}"]) < 1e-6
    assert rel_err(fld.beta, d["after.beta"]) < 1e-6


def test_mesh_grid_golden():
    fld, d = golden_field(), load_npz("mesh.npz")
    axes = O.grid_axes(d["mc_bound"], float(d["resolution"]))
    assert [len(a) for a in axes] == list(d["n"])
    ret = O.query_points(fld, O.grid_points(axes))
    assert rel_err(ret[:, -1], d["sdf"]) < 1e-6 and rel_err(ret[:, :3], d["rgb"]) < 1e-6


def test_factored_lattice_query_matches_the_reference_lattice():
    """The algebra behind eslam_grid_preact + eslam_grid_sdf_factored (first decoder layer applied on the lattice's
    faces) against the reference's own lattice values: same -1 mask, values within 1e-5."""
    fld, d = golden_field(), load_npz("mesh.npz")
    axes = O.grid_axes(d["mc_bound"], float(d["resolution"]))
    sdf = O.query_lattice_factored(fld, axes)
    ref = torch.from_numpy(d["sdf"])
    assert torch.equal(sdf == -1, ref == -1)
    assert (sdf - ref).abs().max().item() < 1e-5


def test_adam_formula_matches_torch_optim():
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(1000, generator=g)
    p_ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([p_ref], lr=5e-3)
    p, m, v = p0.clone(), torch.zeros(1000), torch.zeros(1000)
    for step in range(1, 4):
        grad = torch.randn(1000, generator=g) * 0.1
        p_ref.grad = grad.clone()
        opt.step()
        O.adam_update(p, grad, m, v, step, 5e-3)
    assert rel_err(p, p_ref) < 1e-6


def test_empty_band_gives_nan_loss_but_finite_grads():
    # mean over an empty mask is NaN in the loss value only
    sdf = torch.zeros(4, 8, requires_grad=True)
    z = torch.full((4, 8), 5.0)  # everything behind the surface: no front, no center samples
    d = torch.ones(4)
    loss = O.sdf_loss(sdf, z, d, 0.06, 10, 200, 50)
    assert torch.isnan(loss)


def test_keyframe_overlap_golden():
    """Oracle restatement of Mapper.keyframe_selection_overlap (Mapper.py:146-203) against the fixture the
    unmodified reference produced (tests/golden/make_golden_kfsel.py)."""
    d = load_npz("kfsel.npz")
    cam = O.Camera(int(d["H"]), int(d["W"]), float(d["fx"]), float(d["fy"]), float(d["cx"]), float(d["cy"]))
    frac, cnt, n_pts = O.keyframe_overlap(cam, torch.from_numpy(d["cur_c2w"]), torch.from_numpy(d["depth"]),
                                          torch.from_numpy(d["color"]), torch.from_numpy(d["kf_c2ws"])[:-2],
                                          O.ReplayDraws([torch.from_numpy(d["idx"])]))
    assert torch.equal(frac, torch.from_numpy(d["percent_inside"]))
    assert torch.nonzero(frac).squeeze(-1).tolist() == d["selected"].tolist()
    assert int(n_pts) == int(d["n_pts"]) and cnt.tolist() == d["counts"].tolist()


def test_render_img_golden():
    """Oracle restatement of Renderer.render_img (Renderer.py:155-204) against the reference's output
    (tests/golden/make_golden_img.py), replaying its random draws."""
    from conftest import GOLDEN_CAM, TRUNC, golden_field

    d = load_npz("img.npz")
    draws = [torch.from_numpy(d[f"draw.{k}"]) for k in range(int(d["n_draws"]))]
    dep, col = O.render_image(golden_field(), O.Camera(*GOLDEN_CAM), torch.from_numpy(d["c2w"]),
                              torch.from_numpy(d["gt_depth"]), TRUNC, 32, 8, O.ReplayDraws(draws),
                              ray_batch_size=int(d["ray_batch_size"]))
    assert dep.dtype == torch.float64
    assert rel_err(dep, d["depth"]) < 1e-6 and rel_err(col, d["color"]) < 1e-6


def test_ingest_golden():
    """Oracle restatement of BaseDataset.__getitem__'s arithmetic (datasets.py:88-112) against what the reference's
    Replica loader returned for the fixture frame (tests/golden/make_golden_ingest.py)."""
    d = load_npz("ingest.npz")
    color, depth = O.ingest_frame(d["bgr"], d["depth_u16"], float(d["png_depth_scale"]), int(d["crop_edge"]))
    assert color.dtype == torch.float64 and depth.dtype == torch.float32
    assert torch.equal(color, torch.from_numpy(d["color"])) and torch.equal(depth, torch.from_numpy(d["depth"]))


def test_ingest_tum_golden():
    """Oracle restatement of the TUM-shaped loader path (datasets.py:79-112: cv2.undistort of the uint8 colour image,
    crop_size = bilinear align_corners / nearest resize, crop_edge) against what the reference's TUM_RGBD loader returned
    (make_golden_ingest_tum.py), and the undistortion alone against cv2's own output kept in the fixture."""
    d = load_npz("ingest_tum.npz")
    assert np.array_equal(O.undistort_u8(d["bgr"], *d["cam"], d["distortion"]), d["undistorted"])
    color, depth = O.ingest_frame_tum(d["bgr"], d["depth_u16"], float(d["png_depth_scale"]), tuple(d["cam"]),
                                      d["distortion"], tuple(int(v) for v in d["crop_size"]), int(d["crop_edge"]))
    assert color.dtype == torch.float64 and depth.dtype == torch.float32
    assert torch.equal(color, torch.from_numpy(d["color"])) and torch.equal(depth, torch.from_numpy(d["depth"]))


def test_ingest_scannet_golden():
    """Oracle restatement of the ScanNet-shaped loader path (datasets.py:88-112 with cv2.resize of the float64 colour
    image to the depth's size) against what the reference's ScanNet loader returned (make_golden_ingest_scannet.py)."""
    d = load_npz("ingest_scannet.npz")
    assert d["bgr"].shape[0] > d["depth_u16"].shape[0] and d["bgr"].shape[1] > d["depth_u16"].shape[1]
    color, depth = O.ingest_frame_resized(d["bgr"], d["depth_u16"], float(d["png_depth_scale"]), int(d["crop_edge"]))
    assert color.dtype == torch.float64 and depth.dtype == torch.float32
    assert torch.equal(color, torch.from_numpy(d["color"])) and torch.equal(depth, torch.from_numpy(d["depth"]))
