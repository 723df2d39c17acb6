"""GPU: SURVEY.md 8(f)-1 -- keyframe selection by view overlap (one projection kernel) and the frame table that
replaces the reference's per-call torch.stack of the window's frames."""
import os

import numpy as np
import pytest
import torch

import eslam_oracle as O
from conftest import GOLDEN_CAM, SimpleEslam, base_cfg, golden_field, load_npz, rel_err, to_device_scene

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _mapper(cam, dev=DEV):
    import myslam_b200 as M

    fld = golden_field()
    planes, dec = to_device_scene(fld, dev)
    cfg = base_cfg()
    rnd = M.Renderer(cfg, SimpleEslam(fld.bound.clone(), cam, dev))
    return M.MapperStep(cfg, rnd, dec, planes, fld.bound.clone(), cam, dev), fld


def test_keyframe_overlap_counts_golden(monkeypatch):
    """Counts per keyframe equal the unmodified reference's percent_inside * n_pts on the committed fixture, and the
    drop-in method returns the reference's list (randperm replaced by the identity, as in the generator)."""
    import myslam_b200 as M

    d = load_npz("kfsel.npz")
    cam = (int(d["H"]), int(d["W"]), float(d["fx"]), float(d["fy"]), float(d["cx"]), float(d["cy"]))
    mp, _ = _mapper(cam)
    kfs = torch.from_numpy(d["kf_c2ws"]).to(DEV)
    mp.keyframe_list = list(range(kfs.shape[0]))
    mp.estimate_c2w_list = kfs
    mp.draws = M.ReplayDraws([torch.from_numpy(d["idx"])], DEV)
    monkeypatch.setattr(torch, "randperm", lambda n, *a, **k: torch.arange(n))
    sel = mp.keyframe_selection_overlap(torch.from_numpy(d["color"]).to(DEV), torch.from_numpy(d["depth"]).to(DEV),
                                        torch.from_numpy(d["cur_c2w"]).to(DEV), kfs.shape[0])
    inside, n_pts = mp._last_overlap
    assert int(n_pts) == int(d["n_pts"])
    assert inside.cpu().tolist() == d["counts"].tolist()
    assert [int(s) for s in sel] == d["selected"].tolist()
    frac = inside.float() / n_pts.float()
    assert torch.equal(frac.cpu(), torch.from_numpy(d["percent_inside"]))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_keyframe_overlap_counts_vs_oracle_random(seed):
    """Seeded random cameras at Replica image size; counts may differ from the oracle only for points within float
    rounding of the 20-pixel margin (the oracle inverts with LU, the kernel with cofactors)."""
    import myslam_b200 as M

    g = torch.Generator().manual_seed(seed)
    H, W = 340, 600
    cam = (H, W, 300.0, 300.0, 299.5, 169.5)
    mp, _ = _mapper(cam)
    K = 40
    q = torch.randn(K + 2, 4, generator=g) * 0.3 + torch.tensor([1.0, 0, 0, 0])
    t = torch.randn(K + 2, 3, generator=g) * 1.5
    kfs = O.cam_pose_to_matrix(torch.cat([q, t], -1))
    cur = O.cam_pose_to_matrix(torch.tensor([[1.0, 0.02, 0.01, -0.03, 0.1, 0.2, -0.1]]))[0]
    depth = 1.0 + 3.0 * torch.rand(H, W, generator=g)
    depth[torch.rand(H, W, generator=g) < 0.2] = 0.0
    color = torch.zeros(H, W, 3, dtype=torch.float64)
    idx = torch.randint(H * W, (50,), generator=g)
    ocam = O.Camera(*cam)
    frac, cnt, n_pts = O.keyframe_overlap(ocam, cur, depth, color, kfs[:-2], O.ReplayDraws([idx]))
    mp.keyframe_list = list(range(K + 2))
    mp.estimate_c2w_list = kfs.to(DEV)
    mp.draws = M.ReplayDraws([idx], DEV)
    mp.keyframe_selection_overlap(color.to(DEV), depth.to(DEV), cur.to(DEV), K)
    inside, n = mp._last_overlap
    assert int(n) == int(n_pts)
    diff = (inside.cpu().long() - cnt.long()).abs()
    assert int(diff.max()) <= 1 and int((diff > 0).sum()) <= 2, diff.tolist()
    assert int(cnt.sum()) > 0


def test_frame_table_equals_stacked_frames():
    """The same mapping iteration with the window given as a FrameTable (frames read where they live) and as the
    reference's stacked tensors: identical kept rays, gradients and loss."""
    from myslam_b200 import ReplayDraws
    from myslam_b200.decoders import synced_store
    from myslam_b200.hotpath import FrameTable, mapping_iteration
    from myslam_b200.mapper import _mapper_state

    mp, fld = _mapper(GOLDEN_CAM)
    d = load_npz("mapping.npz")
    st = _mapper_state(mp, 400, 4)
    store = synced_store((mp.planes_xy, mp.planes_xz, mp.planes_yz, mp.c_planes_xy, mp.c_planes_xz, mp.c_planes_yz),
                         mp.decoders, fld.bound)
    c2ws = torch.from_numpy(d["c2ws0"]).to(DEV)
    cols, deps = torch.from_numpy(d["gt_colors"]).to(DEV), torch.from_numpy(d["gt_depths"]).to(DEV)
    draws = [torch.from_numpy(d[f"draw.{k}"]) for k in range(int(d["n_draws"]))]
    outs = []
    for mode in ("stack", "table"):
        store.reset_adam()
        if mode == "stack":
            gc, gd = cols, deps
        else:  # separate allocations in a shuffled order of creation: nothing contiguous about them
            frames = [(cols[k].clone(), deps[k].clone()) for k in (2, 0, 3, 1)]
            frames = [frames[[2, 0, 3, 1].index(k)] for k in range(4)]
            gc = gd = FrameTable([f[0] for f in frames], [f[1] for f in frames], st["sc"].cam, DEV)
        mapping_iteration(st["ws"], store, st["sc"], c2ws, None, gc, gd, 100, 1, 1e-3, 5e-3, 5e-3, 1e-3,
                          draws=ReplayDraws(draws, DEV), strict_rng=True, want_loss=True, apply_adam=False)
        ws = st["ws"]
        R = int(ws.counters[0])
        outs.append((R, ws.src[:R].clone(), ws.gt_color[:R].clone(), ws.gt_depth[:R].clone(), store.grad.clone(),
                     ws.loss_acc[5].item()))
    a, b = outs
    assert a[0] == b[0] and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])
    assert rel_err(b[4], a[4]) < 1e-5 and abs(a[5] - b[5]) <= 1e-6 * abs(a[5])


def test_keyframe_device_cache():
    """`keyframe_device: cpu` keyframes are uploaded once and re-used across optimize_mapping calls."""
    from myslam_b200.mapper import _device_frame

    mp, _ = _mapper(GOLDEN_CAM)
    t = torch.rand(GOLDEN_CAM[0], GOLDEN_CAM[1])
    a = _device_frame(mp, t, torch.float32)
    b = _device_frame(mp, t, torch.float32)
    assert a.is_cuda and a.data_ptr() == b.data_ptr() and torch.equal(a.cpu(), t)
    t.add_(1.0)  # in-place change -> new version -> re-upload
    c = _device_frame(mp, t, torch.float32)
    assert torch.equal(c.cpu(), t)
    g = torch.rand(4, 4, device=DEV)
    assert _device_frame(mp, g, torch.float32) is g


def test_pose_conversion_kernels_match_host_mirror():
    """eslam_matrix_to_pose / eslam_pose_to_matrix == common.matrix_to_cam_pose / cam_pose_to_matrix (which the CPU
    tests pin against the reference's functions through the golden pose fixture), bit for bit on the device."""
    from myslam_b200._lib import call, ptr, stream
    from myslam_b200.common import cam_pose_to_matrix, matrix_to_cam_pose

    d = load_npz("pose.npz")
    poses = torch.from_numpy(d["poses"]).to(DEV)
    mats = torch.from_numpy(d["mats"]).to(DEV).contiguous()
    n = poses.shape[0]
    out_p = torch.empty(n, 7, device=DEV)
    call("eslam_matrix_to_pose", ptr(mats), ptr(out_p), n, stream())
    assert torch.equal(out_p, matrix_to_cam_pose(mats))
    assert rel_err(out_p, d["back"]) < 1e-6  # the fixture was computed by torch on the CPU
    out_m = torch.empty(n, 4, 4, device=DEV)
    call("eslam_pose_to_matrix", ptr(poses.contiguous()), ptr(out_m), n, stream())
    # (torch's device reduction adds |q|^2 in a different order than its CPU loop, which the kernel follows)
    assert rel_err(out_m, cam_pose_to_matrix(poses)) < 1e-6
    assert rel_err(out_m, d["mats"]) < 1e-6
    # every branch of the argmax: rotations by ~180 degrees about each axis
    q = torch.tensor([[1e-3, 1, 0, 0], [1e-3, 0, 1, 0], [1e-3, 0, 0, 1], [1, 0, 0, 0], [0.5, 0.5, 0.5, 0.5]], device=DEV)
    m = cam_pose_to_matrix(torch.cat([q, torch.zeros(5, 3, device=DEV)], -1)).contiguous()
    o = torch.empty(5, 7, device=DEV)
    call("eslam_matrix_to_pose", ptr(m), ptr(o), 5, stream())
    assert torch.equal(o, matrix_to_cam_pose(m))



def test_ingest_frame_bit_exact():
    """eslam_ingest_frame == the reference's Replica loader on the fixture frame, bit for bit."""
    from myslam_b200.ingest import ingest_frame

    d = load_npz("ingest.npz")
    color, depth = ingest_frame(d["bgr"], d["depth_u16"], float(d["png_depth_scale"]), int(d["crop_edge"]), DEV)
    assert color.dtype == torch.float64 and depth.dtype == torch.float32
    assert torch.equal(color.cpu(), torch.from_numpy(d["color"]))
    assert torch.equal(depth.cpu(), torch.from_numpy(d["depth"]))
    # no crop, large values, full uint16 range
    rng = np.random.default_rng(3)
    bgr = rng.integers(0, 256, size=(33, 47, 3), dtype=np.uint8)
    dep = rng.integers(0, 65536, size=(33, 47), dtype=np.uint16)
    c2, d2 = ingest_frame(bgr, dep, 1000.0, 0, DEV)
    oc, od = O.ingest_frame(bgr, dep, 1000.0, 0)
    assert torch.equal(c2.cpu(), oc) and torch.equal(d2.cpu(), od)


def test_ingest_scannet_shaped_frame_bit_exact():
    """eslam_ingest_frame_resized == the reference's ScanNet loader (colour 2x the depth's size, resized by cv2.resize
    on the float64 image, crop_edge) on the fixture frame, bit for bit; and the oracle on a second, odd-sized case."""
    from myslam_b200.ingest import ingest_frame

    d = load_npz("ingest_scannet.npz")
    color, depth = ingest_frame(d["bgr"], d["depth_u16"], float(d["png_depth_scale"]), int(d["crop_edge"]), DEV)
    assert color.dtype == torch.float64 and depth.dtype == torch.float32
    assert torch.equal(depth.cpu(), torch.from_numpy(d["depth"]))
    assert torch.equal(color.cpu(), torch.from_numpy(d["color"]))
    rng = np.random.default_rng(5)
    bgr = rng.integers(0, 256, size=(61, 83, 3), dtype=np.uint8)
    dep = rng.integers(0, 65536, size=(29, 40), dtype=np.uint16)
    c2, d2 = ingest_frame(bgr, dep, 1000.0, 1, DEV)
    oc, od = O.ingest_frame_resized(bgr, dep, 1000.0, 1)
    assert torch.equal(c2.cpu(), oc) and torch.equal(d2.cpu(), od)


def test_ingest_tum_shaped_frame_bit_exact():
    """Undistortion + crop_size + crop_edge on the device == the reference's TUM_RGBD loader on the fixture frame, bit
    for bit (eslam_undistort_u8 against cv2.undistort's own output, eslam_ingest_frame_crop against the loader's
    F.interpolate steps); and the oracle on a second, odd-sized case with and without each step."""
    from myslam_b200.ingest import ingest_frame, undistort

    d = load_npz("ingest_tum.npz")
    cam, dist, crop = tuple(d["cam"]), d["distortion"], tuple(int(v) for v in d["crop_size"])
    und = undistort(torch.from_numpy(d["bgr"]).to(DEV), cam, dist)
    assert torch.equal(und.cpu(), torch.from_numpy(d["undistorted"]))
    color, depth = ingest_frame(d["bgr"], d["depth_u16"], float(d["png_depth_scale"]), int(d["crop_edge"]), DEV, cam=cam,
                                distortion=dist, crop_size=crop)
    assert color.dtype == torch.float64 and depth.dtype == torch.float32
    assert torch.equal(depth.cpu(), torch.from_numpy(d["depth"]))
    assert torch.equal(color.cpu(), torch.from_numpy(d["color"]))
    rng = np.random.default_rng(9)
    bgr = rng.integers(0, 256, size=(45, 61, 3), dtype=np.uint8)
    dep = rng.integers(0, 65536, size=(45, 61), dtype=np.uint16)
    cam2, dist2 = (50.3, 49.1, 30.2, 22.6), (0.2312, -0.7849, -0.0033, -0.0001, 0.9172)
    for kw in (dict(cam=cam2, distortion=dist2, crop_size=(37, 50)), dict(crop_size=(36, 49)),
               dict(cam=cam2, distortion=dist2)):
        c2, d2 = ingest_frame(bgr, dep, 5000.0, 3, DEV, **kw)
        oc, od = O.ingest_frame_tum(bgr, dep, 5000.0, kw.get("cam"), kw.get("distortion"), kw.get("crop_size"), 3)
        assert torch.equal(d2.cpu(), od), kw
        assert torch.equal(c2.cpu(), oc), kw


def test_grid_query_with_convex_mesh_bound():
    """Mesher.get_mesh forces sdf = -1 outside the convex hull of the observed region (Mesher.py:206-217); here the
    half-space test runs inside the grid query.  Against the plain query + a float64 half-space mask, away from the
    faces' rounding band."""
    from scipy.spatial import ConvexHull
    from myslam_b200 import grid_axes, hull_planes, query_grid_sdf

    mp, fld = _mapper(GOLDEN_CAM)
    planes = (mp.planes_xy, mp.planes_xz, mp.planes_yz, mp.c_planes_xy, mp.c_planes_xz, mp.c_planes_yz)
    b = fld.bound.numpy().astype(np.float64)
    rng = np.random.default_rng(1)
    pts = b[:, 0] + (b[:, 1] - b[:, 0]) * (0.15 + 0.7 * rng.random((40, 3)))  # a blob strictly inside the bound
    hull = ConvexHull(pts)
    hp = hull_planes(pts, hull.simplices)
    assert hp.shape == (len(hull.simplices), 4)
    axes = grid_axes(b, 0.03)
    plain = query_grid_sdf(planes, mp.decoders, axes, fld.bound, separable=False).cpu()
    bounded = query_grid_sdf(planes, mp.decoders, axes, fld.bound, hull=hp, separable=False).cpu()
    # the separable form (plane features resampled once on the lattice's faces) is bit-identical to the direct one
    assert torch.equal(query_grid_sdf(planes, mp.decoders, axes, fld.bound, separable=True).cpu(), plain)
    assert torch.equal(query_grid_sdf(planes, mp.decoders, axes, fld.bound, hull=hp, separable=True).cpu(), bounded)
    gx, gy, gz = np.meshgrid(*axes, indexing="xy")  # Mesher.get_grid_uniform's point order
    p = np.stack([gx.reshape(-1), gy.reshape(-1), gz.reshape(-1)], 1)
    margin = (p @ hp[:, :3].double().numpy().T + hp[:, 3].double().numpy()).max(1)  # <= 0 inside
    sure_in, sure_out = torch.from_numpy(margin < -1e-4), torch.from_numpy(margin > 1e-4)
    assert int(sure_in.sum()) > 100 and int(sure_out.sum()) > 100
    assert torch.equal(bounded[sure_in], plain[sure_in])
    assert (bounded[sure_out] == -1.0).all()
    # sharded ranges give the same lattice values
    n = bounded.numel()
    parts = [query_grid_sdf(planes, mp.decoders, axes, fld.bound, start=s, count=c, hull=hp, separable=True).cpu()
             for s, c in ((0, n // 3), (n // 3, n - n // 3))]
    assert torch.equal(torch.cat(parts), bounded)
    # the factored form (first layer applied on the faces) re-associates one sum: same mask, values within 1e-5
    fac = query_grid_sdf(planes, mp.decoders, axes, fld.bound, factored=True).cpu()
    fac_b = query_grid_sdf(planes, mp.decoders, axes, fld.bound, hull=hp, factored=True).cpu()
    assert (fac - plain).abs().max().item() < 1e-5 and (fac_b - bounded).abs().max().item() < 1e-5
    assert torch.equal(fac_b == -1.0, bounded == -1.0)
    cuts = (0, 77, n // 3 + 5, n - 300, n)  # ragged ranges: CTAs straddle z columns and range ends
    ragged = lambda: torch.cat([query_grid_sdf(planes, mp.decoders, axes, fld.bound, start=a, count=b - a, hull=hp,
                                               factored=True).cpu() for a, b in zip(cuts[:-1], cuts[1:])])
    # whole lattice rows run eslam_grid_sdf_rows (a warp walks y with the xz values in registers), the ragged ends of a
    # range the per-voxel kernel: the same operation order (packed FP32 is two exact FMAs), so the same bits
    assert torch.equal(ragged(), fac_b)
    os.environ["ESLAM_B200_GRID_ROWS"] = "0"
    try:  # the per-voxel kernel alone
        voxel_b = query_grid_sdf(planes, mp.decoders, axes, fld.bound, hull=hp, factored=True).cpu()
        voxel = query_grid_sdf(planes, mp.decoders, axes, fld.bound, factored=True).cpu()
        assert torch.equal(ragged(), voxel_b)
    finally:
        del os.environ["ESLAM_B200_GRID_ROWS"]
    assert torch.equal(voxel_b, fac_b) and torch.equal(voxel, fac)
