"""GPU parity: the CUDA path through the C ABI vs the oracle / golden vectors on identical inputs and draws.
Tolerances are north_star's: indices, masks and depth-guided z_vals bit-exact; rendered depth/colour 1e-4
relative; gradients 1e-3 relative (relative to the tensor's max magnitude)."""
import numpy as np
import pytest
import torch

import eslam_oracle as O
from conftest import (GOLDEN_CAM, TRUNC, SimpleEslam, arena_index, base_cfg, golden_field, load_npz, recorded_draws, rel_err,
                      to_device_scene)

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL_VAL, TOL_GRAD = 1e-4, 1e-3


def make_renderer(fld, n_strat=32, n_imp=8):
    from myslam_b200 import Renderer

    cfg = base_cfg(n_strat, n_imp)
    return Renderer(cfg, SimpleEslam(fld.bound.clone(), GOLDEN_CAM, DEV))


# ------------------------------------------------------------------------------------------- layout
def test_plane_layout_round_trip():
    from myslam_b200 import FieldStore

    fld = golden_field()
    planes, dec = to_device_scene(fld)
    store = FieldStore.from_planes(planes, fld.bound, DEV)
    store.pull_planes(planes)
    from myslam_b200.field import flatten_planes
    for i, p in enumerate(flatten_planes(planes)):
        assert torch.equal(store.plane_view(i), p[0].permute(1, 2, 0))
    out = tuple([torch.zeros_like(p) for p in g] for g in planes)
    store.push_planes(out)
    for a, b in zip(flatten_planes(out), flatten_planes(planes)):
        assert torch.equal(a, b)


# ------------------------------------------------------------------------------------------- decoders
def test_decoders_forward_golden():
    fld, d = golden_field(), load_npz("decoders.npz")
    planes, dec = to_device_scene(fld)
    pts = torch.from_numpy(d["pts"]).to(DEV)
    with torch.no_grad():
        raw = dec(pts, planes)
    assert raw.shape == (pts.shape[0], 4)
    assert rel_err(raw, d["raw"]) < TOL_VAL
    pn = O.normalize_pts(torch.from_numpy(d["pts"]).clone(), fld.bound).to(DEV)
    with torch.no_grad():
        feat = dec.sample_plane_feature(pn, planes[0], planes[1], planes[2])
        sdf = dec.get_raw_sdf(pn, planes)
        rgb = dec.get_raw_rgb(pn, planes)
    assert rel_err(feat, d["feat_sdf"]) < TOL_VAL
    assert rel_err(sdf, d["raw"][:, 3]) < TOL_VAL and rel_err(rgb, d["raw"][:, :3]) < TOL_VAL


def test_decoders_backward_vs_oracle():
    fld = golden_field()
    g = torch.Generator().manual_seed(3)
    lo, hi = fld.bound[:, 0], fld.bound[:, 1]
    pts = lo + (hi - lo) * (torch.rand(300, 3, generator=g) * 1.1 - 0.05)
    gr = torch.randn(300, 4, generator=g)
    f2 = fld.clone(requires_grad=True)
    p_o = pts.clone().requires_grad_(True)
    (O.decode(p_o, f2) * gr).sum().backward()
    planes, dec = to_device_scene(fld, requires_grad=True)
    p_c = pts.clone().to(DEV).requires_grad_(True)
    (dec(p_c, planes) * gr.to(DEV)).sum().backward()
    assert rel_err(p_c.grad, p_o.grad) < TOL_GRAD
    from myslam_b200.field import flatten_planes
    flat = flatten_planes(planes)
    for k, b in enumerate(f2.leaves()[:12]):
        assert rel_err(flat[arena_index(k)].grad, b.grad) < TOL_GRAD, f"plane {k}"
    named = dict(dec.named_parameters())
    for name in O.DECODER_KEYS:
        assert rel_err(named[name].grad, f2.dec[name].grad) < TOL_GRAD, name


# ------------------------------------------------------------------------------------------- render_batch_ray
def test_render_batch_ray_golden():
    from myslam_b200 import ReplayDraws
    from myslam_b200.field import flatten_planes

    fld, d = golden_field(), load_npz("render.npz")
    planes, dec = to_device_scene(fld, requires_grad=True)
    rnd = make_renderer(fld)
    rnd.draws = ReplayDraws(recorded_draws(d), DEV)
    ro = torch.from_numpy(d["rays_o"]).to(DEV).requires_grad_(True)
    rd = torch.from_numpy(d["rays_d"]).to(DEV).requires_grad_(True)
    gt = torch.from_numpy(d["gt_depth"]).to(DEV)
    depth, rgb, sdf, z = rnd.render_batch_ray(planes, dec, rd, ro, DEV, TRUNC, gt_depth=gt)
    has = (gt > 0).cpu()
    assert torch.equal(z.cpu()[has], torch.from_numpy(d["z"])[has]), "depth-guided z_vals must be bit-exact"
    assert rel_err(z, d["z"]) < TOL_VAL
    assert rel_err(depth, d["depth"]) < TOL_VAL and rel_err(rgb, d["rgb"]) < TOL_VAL and rel_err(sdf, d["sdf"]) < TOL_VAL
    loss = (depth * torch.from_numpy(d["g_depth"]).to(DEV)).sum() + (rgb * torch.from_numpy(d["g_rgb"]).to(DEV)).sum() \
        + (sdf * torch.from_numpy(d["g_sdf"]).to(DEV)).sum()
    loss.backward()
    assert rel_err(ro.grad, d["d_rays_o"]) < TOL_GRAD and rel_err(rd.grad, d["d_rays_d"]) < TOL_GRAD
    flat = flatten_planes(planes)
    for k in range(12):
        assert rel_err(flat[arena_index(k)].grad, d[f"d_plane.{k}"]) < TOL_GRAD, f"plane {k}"
    named = dict(dec.named_parameters())
    for name in list(O.DECODER_KEYS) + ["beta"]:
        assert rel_err(named[name].grad, d[f"d_dec.{name}"]) < TOL_GRAD, name


@pytest.mark.parametrize("n_strat,n_imp,R", [(48, 8, 37), (32, 8, 1), (32, 8, 130), (16, 4, 50)])
def test_render_shapes_vs_oracle(n_strat, n_imp, R):
    """ScanNet's 56 samples, single ray, ragged last CTA, and a small-S configuration."""
    from myslam_b200 import ReplayDraws

    fld = golden_field()
    g = torch.Generator().manual_seed(R)
    c2w = O.cam_pose_to_matrix(torch.tensor([[1.0, 0.02, -0.01, 0.03, 0.02, -0.03, 0.25]]))
    ii = torch.randint(0, 64, (R,), generator=g).float()
    jj = torch.randint(0, 48, (R,), generator=g).float()
    ro, rd = O.rays_from_pixels(ii[None], jj[None], c2w, 50.0, 50.0, 31.5, 23.5)
    ro, rd = ro.reshape(-1, 3).contiguous(), rd.reshape(-1, 3).contiguous()
    gt = 0.3 + 0.2 * torch.rand(R, generator=g)
    if R > 4:
        gt[::5] = 0.0
    draws = O.LiveDraws(g)
    depth, rgb, sdf, z = O.render_rays(fld, ro, rd, gt, TRUNC, n_strat, n_imp, draws)
    planes, dec = to_device_scene(fld)
    rnd = make_renderer(fld, n_strat, n_imp)
    rnd.draws = ReplayDraws(draws.log, DEV)
    with torch.no_grad():
        d2, c2, s2, z2 = rnd.render_batch_ray(planes, dec, rd.to(DEV), ro.to(DEV), DEV, TRUNC, gt_depth=gt.to(DEV))
    has = gt > 0
    assert torch.equal(z2.cpu()[has], z[has])
    assert rel_err(z2, z) < TOL_VAL and rel_err(d2, depth) < TOL_VAL and rel_err(c2, rgb) < TOL_VAL
    assert rel_err(s2, sdf) < TOL_VAL


def test_render_empty_batch():
    fld = golden_field()
    planes, dec = to_device_scene(fld)
    rnd = make_renderer(fld)
    e = torch.zeros(0, 3, device=DEV)
    d, c, s, z = rnd.render_batch_ray(planes, dec, e, e, DEV, TRUNC, gt_depth=torch.zeros(0, device=DEV))
    assert d.shape == (0,) and c.shape == (0, 3) and s.shape == (0, 40) and z.shape == (0, 40)


# ------------------------------------------------------------------------------------------- tracking
def make_tracker(fld, d):
    from myslam_b200 import TrackerStep

    planes, dec = to_device_scene(fld)
    cfg = base_cfg()
    cfg["tracking"].update(ignore_edge_H=int(d["edge_h"]), ignore_edge_W=int(d["edge_w"]), pixels=int(d["n_pix"]),
                           iters=int(d["iters"]), lr_T=float(d["lr_T"]), lr_R=float(d["lr_R"]))
    trk = TrackerStep(cfg, make_renderer(fld), dec, planes, fld.bound.clone(), GOLDEN_CAM, DEV)
    trk.strict_rng = True
    return trk


def test_tracking_sampling_bit_exact():
    from myslam_b200 import ReplayDraws
    from myslam_b200.hotpath import tracking_iteration
    from myslam_b200.tracker import _tracker_state, _tracker_store

    fld, d = golden_field(), load_npz("tracking.npz")
    trk = make_tracker(fld, d)
    draws = recorded_draws(d)
    st = _tracker_state(trk, int(d["n_pix"]))
    store = _tracker_store(trk, st)
    pose = torch.from_numpy(d["pose0"]).to(DEV).contiguous()
    gc, gd = torch.from_numpy(d["gt_color"]).to(DEV), torch.from_numpy(d["gt_depth"]).to(DEV)
    tracking_iteration(st["ws"], store, st["sc"], pose, gc, gd, int(d["n_pix"]), draws=ReplayDraws(draws[:2], DEV),
                       strict_rng=True)
    ws = st["ws"]
    keep = torch.from_numpy(d["it0_keep"])
    R = int(keep.sum())
    assert int(ws.counters[0]) == R
    # which pixels survived, in the reference's order
    assert torch.equal(ws.src[:R].cpu().long(), torch.nonzero(keep).squeeze(-1))
    o1 = O.tracking_forward(fld, O.Camera(*GOLDEN_CAM), O.RenderCfg(32, 8, TRUNC), O.TRACK_W,
                            torch.from_numpy(d["pose0"]), torch.from_numpy(d["gt_color"]),
                            torch.from_numpy(d["gt_depth"]), int(d["n_pix"]), int(d["edge_h"]), int(d["edge_w"]),
                            O.ReplayDraws(draws[:2]))
    assert torch.equal(ws.gt_depth[:R].cpu(), o1.gt_depth)
    assert torch.equal(ws.gt_color[:R].cpu(), o1.gt_color)
    assert torch.equal(ws.rays_o[:R].cpu(), o1.rays_o.detach())
    assert torch.equal(ws.rays_d[:R].cpu(), o1.rays_d.detach()), "ray directions must be bit-exact"
    assert torch.equal(ws.z[:R].cpu(), torch.from_numpy(d["it0_z"])), "z_vals must be bit-exact"
    assert rel_err(ws.depth[:R], d["it0_depth"]) < TOL_VAL and rel_err(ws.rgb[:R], d["it0_rgb"]) < TOL_VAL
    assert torch.equal(ws.ray_mask[:R].cpu().bool(), torch.from_numpy(d["it0_mask"]))
    assert abs(ws.loss_acc[5].item() - d["losses"][0]) / abs(d["losses"][0]) < TOL_VAL
    assert rel_err(ws.grad7[0:1], d["pose_grads"][0:1]) < TOL_GRAD


def test_optimize_tracking_trace_golden():
    """The drop-in method driven like Tracker.run drives it (Tracker.py:291-307)."""
    from myslam_b200 import ReplayDraws

    fld, d = golden_field(), load_npz("tracking.npz")
    trk = make_tracker(fld, d)
    trk.draws = ReplayDraws(recorded_draws(d), DEV)
    pose0 = torch.from_numpy(d["pose0"]).to(DEV)
    gc, gd = torch.from_numpy(d["gt_color"]).to(DEV), torch.from_numpy(d["gt_depth"]).to(DEV)
    T = torch.nn.Parameter(pose0[:, -3:].clone())
    Rq = torch.nn.Parameter(pose0[:, :4].clone())
    opt = torch.optim.Adam([{"params": [T], "lr": float(d["lr_T"]), "betas": (0.5, 0.999)},
                            {"params": [Rq], "lr": float(d["lr_R"]), "betas": (0.5, 0.999)}])
    losses, grads = [], []
    for _ in range(int(d["iters"])):
        pose = torch.cat([Rq, T], -1)
        losses.append(trk.optimize_tracking(pose, gc, gd, int(d["n_pix"]), opt))
        grads.append(torch.cat([Rq.grad, T.grad], -1).clone())
    assert rel_err(torch.tensor(losses), d["losses"]) < TOL_VAL
    assert rel_err(torch.cat(grads, 0), d["pose_grads"]) < TOL_GRAD
    assert rel_err(torch.cat([Rq, T], -1), d["pose_trace"][-1:]) < 1e-5


def test_track_frame_fused_adam_golden():
    from myslam_b200 import ReplayDraws

    fld, d = golden_field(), load_npz("tracking.npz")
    trk = make_tracker(fld, d)
    trk.draws = ReplayDraws(recorded_draws(d), DEV)
    gc, gd = torch.from_numpy(d["gt_color"]).to(DEV), torch.from_numpy(d["gt_depth"]).to(DEV)
    best, losses, final = trk.track_frame(torch.from_numpy(d["pose0"]).to(DEV), gc, gd)
    assert rel_err(losses, d["losses"]) < TOL_VAL
    assert rel_err(final, d["pose_trace"][-1:]) < 1e-5
    k = int(np.argmin(d["losses"]))
    assert rel_err(best, d["pose_trace"][k:k + 1]) < 1e-5


def test_track_mask_median_rule():
    """Lower median and the 10x rule on hand-made errors, even and odd counts (Tracker.py:193-195)."""
    import ctypes as C
    from myslam_b200._lib import call, ptr, stream

    for R in (1, 2, 7, 200, 2049):
        g = torch.Generator().manual_seed(R)
        gt = torch.rand(R, generator=g) + 0.5
        dep = gt + torch.randn(R, generator=g) * 0.01
        dep[::9] += 1.0  # outliers
        band = torch.randint(0, 20, (R, 4), generator=g, dtype=torch.uint8)
        err = (gt - dep).abs()
        med = err.median()
        mask = err < 10 * med
        cnt = torch.zeros(8, dtype=torch.int32, device=DEV)
        cnt[0] = R
        rm = torch.zeros(R, dtype=torch.uint8, device=DEV)
        scratch = torch.zeros(R + 1, device=DEV)
        gt_d, dep_d, band_d = gt.to(DEV), dep.to(DEV), band.to(DEV)  # keep alive: ptr() does not hold a reference
        call("eslam_track_mask", ptr(gt_d), ptr(dep_d), ptr(band_d), R, ptr(cnt), ptr(rm), ptr(scratch), stream())
        torch.cuda.synchronize()
        assert scratch[R].item() == med.item(), f"R={R}: median {scratch[R].item()} vs {med.item()}"
        assert torch.equal(rm.cpu().bool(), mask), f"R={R}: mask differs in {(rm.cpu().bool() != mask).sum()} rays"
        exp = [int(mask.sum())] + [int(band[mask][:, k].long().sum()) for k in range(3)]
        assert cnt[2:6].tolist() == exp, f"R={R}: {cnt.tolist()} vs {exp}"


# ------------------------------------------------------------------------------------------- mapping
def make_mapper(fld, d):
    from myslam_b200 import MapperStep

    planes, dec = to_device_scene(fld)
    cfg = base_cfg()
    mp = MapperStep(cfg, make_renderer(fld), dec, planes, fld.bound.clone(), GOLDEN_CAM, DEV)
    mp.strict_rng = True
    return mp


def test_mapping_iteration_gradients_golden():
    from myslam_b200 import ReplayDraws
    from myslam_b200.common import matrix_to_cam_pose
    from myslam_b200.decoders import synced_store
    from myslam_b200.hotpath import mapping_iteration
    from myslam_b200.mapper import _mapper_state

    fld, d = golden_field(), load_npz("mapping.npz")
    mp = make_mapper(fld, d)
    all_planes = (mp.planes_xy, mp.planes_xz, mp.planes_yz, mp.c_planes_xy, mp.c_planes_xz, mp.c_planes_yz)
    st = _mapper_state(mp, 400, 4)
    store = synced_store(all_planes, mp.decoders, mp.bound)
    store.reset_adam()
    c2ws = torch.from_numpy(d["c2ws0"]).to(DEV)
    poses7 = torch.zeros(4, 7, device=DEV)
    poses7[1:] = matrix_to_cam_pose(c2ws[1:])
    gc, gd = torch.from_numpy(d["gt_colors"]).to(DEV), torch.from_numpy(d["gt_depths"]).to(DEV)
    draws = recorded_draws(d)
    mapping_iteration(st["ws"], store, st["sc"], c2ws, poses7, gc, gd, 100, 1, 1e-3, 5e-3, 5e-3, 1e-3,
                      draws=ReplayDraws(draws[:4], DEV), strict_rng=True, want_loss=True, apply_adam=False)
    ws = st["ws"]
    keep = torch.from_numpy(d["it0_keep"])
    R = int(keep.sum())
    assert int(ws.counters[0]) == R
    assert torch.equal(ws.src[:R].cpu().long(), torch.nonzero(keep).squeeze(-1))
    has = (ws.gt_depth[:R] > 0).cpu()
    zg = torch.from_numpy(d["it0_z"])
    assert torch.equal(ws.z[:R].cpu()[has], zg[has]), "depth-guided z_vals must be bit-exact"
    assert rel_err(ws.z[:R], zg) < TOL_VAL
    assert abs(ws.loss_acc[5].item() - float(d["it0_loss"])) / abs(float(d["it0_loss"])) < TOL_VAL
    for k in range(12):
        assert rel_err(store.export_plane(arena_index(k), store.grad), d[f"it0_d_plane.{k}"]) < TOL_GRAD, f"plane {k}"
    gdec = store.dec_grad_dict(store.grad)
    for name in O.DECODER_KEYS:
        assert rel_err(gdec[name].reshape(d[f"it0_d_dec.{name}"].shape), d[f"it0_d_dec.{name}"]) < TOL_GRAD, name
    assert rel_err(gdec["beta"], d["it0_beta_grad"]) < TOL_GRAD
    assert rel_err(ws.grad7[1:4], d["it0_pose_grad"]) < TOL_GRAD


def test_optimize_mapping_call_golden():
    """The drop-in method on the golden window (b=4, joint_opt): post-Adam planes, decoders and poses."""
    from myslam_b200 import ReplayDraws

    fld, d = golden_field(), load_npz("mapping.npz")
    mp = make_mapper(fld, d)
    mp.draws = ReplayDraws(recorded_draws(d), DEV)
    mp.joint_opt = True
    c2ws0 = torch.from_numpy(d["c2ws0"]).to(DEV)
    gc, gd = torch.from_numpy(d["gt_colors"]).to(DEV), torch.from_numpy(d["gt_depths"]).to(DEV)
    kf = [{"gt_c2w": c2ws0[k], "idx": torch.tensor(4 * k), "color": gc[k], "depth": gd[k], "est_c2w": c2ws0[k].clone()}
          for k in range(3)]
    mp.keyframe_dict = kf
    np.random.seed(3)  # same window draw as the generator: [0] + last two
    new_cur = mp.optimize_mapping(int(d["iters"]), 1.0, torch.tensor(12), gc[3], gd[3], c2ws0[3].clone(), kf, [0, 4, 8],
                                  c2ws0[3].clone())
    after = torch.stack([c2ws0[0], kf[1]["est_c2w"], kf[2]["est_c2w"], new_cur], 0)
    assert rel_err(after, d["c2ws_after"]) < 1e-5
    # the optimiser's moments come from the oracle (pinned to the reference at 0.0) run on the same recorded draws
    from myslam_b200.decoders import synced_store
    from myslam_b200.field import arena_slot

    f_o, state = fld.clone(), {}
    lr = mp.cfg["mapping"]["lr"]
    O.map_window(f_o, O.Camera(*GOLDEN_CAM), O.RenderCfg(32, 8, TRUNC), O.MAP_W, torch.from_numpy(d["c2ws0"]),
                 torch.from_numpy(d["gt_colors"]), torch.from_numpy(d["gt_depths"]), 400, int(d["iters"]),
                 lr["decoders_lr"], lr["planes_lr"], lr["c_planes_lr"], True, 1e-3, O.ReplayDraws(recorded_draws(d)),
                 state_out=state)
    names = ("xy", "xz", "yz", "c_xy", "c_xz", "c_yz")
    groups = (mp.planes_xy, mp.planes_xz, mp.planes_yz, mp.c_planes_xy, mp.c_planes_xz, mp.c_planes_yz)
    store = synced_store(groups, mp.decoders, mp.bound)
    for gi, (n, g) in enumerate(zip(names, groups)):
        for s in range(2):
            ref_after = torch.from_numpy(d[f"after.plane.{n}.{s}"])
            assert rel_err(f_o.planes[gi][s], ref_after) < 1e-6, "the oracle run must reproduce the reference's planes"
            assert rel_err(g[s], ref_after) < 1e-3
            # Adam's state: first and second moments in the arena against torch.optim.Adam's (1e-3)
            slot = arena_slot(gi, s)
            m_ref, v_ref = state["exp_avg"][gi * 2 + s], state["exp_avg_sq"][gi * 2 + s]
            assert rel_err(store.export_plane(slot, store.exp_avg), m_ref) < 1e-3, f"plane {n}[{s}] exp_avg"
            assert rel_err(store.export_plane(slot, store.exp_avg_sq), v_ref) < 1e-3, f"plane {n}[{s}] exp_avg_sq"
            # Adam's first steps move every touched texel by ~lr whatever the size of its gradient: compare the UPDATE
            # where the gradient is above a floor (below it the direction m / sqrt(v) amplifies rounding noise)
            before = fld.planes[gi][s]
            upd_ref = ref_after - before
            upd = g[s].detach().cpu() - before
            big = m_ref.abs() > 1e-2 * m_ref.abs().max()
            assert big.any()
            err = ((upd - upd_ref).abs()[big].max() / upd_ref.abs().max()).item()
            assert err < 1e-3, f"plane {n}[{s}] update: {err}"
            assert rel_err(upd, upd_ref) < 2e-2, f"plane {n}[{s}] update (all texels)"
    sd = mp.decoders.state_dict()
    for name in O.DECODER_KEYS:
        assert rel_err(sd[name], d[f"after.dec.{name}"]) < 1e-3, name
    assert rel_err(sd["beta"], d["after.beta"]) < 1e-4


def test_adam_kernel_matches_torch():
    import ctypes as C
    from myslam_b200._lib import call, ptr, stream

    n = 4096
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(n, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([{"params": [ref], "lr": 5e-3}])
    p, m, v = p0.clone().to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for step in range(1, 5):
        grad = torch.randn(n, generator=g) * 0.1
        grad[::3] = 0
        ref.grad = grad.clone()
        opt.step()
        gdev = grad.clone().to(DEV)
        call("eslam_adam_step", ptr(p), ptr(gdev), ptr(m), ptr(v), n, (C.c_int64 * 1)(n), (C.c_double * 1)(5e-3), 1,
             step, 0.9, 0.999, 1e-8, stream())
        assert float(gdev.abs().max()) == 0.0, "gradient must be zeroed by the step"
    assert rel_err(p, ref) < 1e-6


# ------------------------------------------------------------------------------------------- mesh query
def test_grid_sdf_golden():
    from myslam_b200 import eval_points, grid_axes, query_grid_sdf

    fld, d = golden_field(), load_npz("mesh.npz")
    planes, dec = to_device_scene(fld)
    axes = grid_axes(d["mc_bound"], float(d["resolution"]))
    sdf = query_grid_sdf(planes, dec, axes)
    assert rel_err(sdf, d["sdf"]) < TOL_VAL
    # forced -1 outside the open box must be exact
    out = torch.from_numpy(d["sdf"]) == -1
    assert torch.equal(sdf.cpu() == -1, out)
    # sharded query == whole query
    n = sdf.numel()
    a = query_grid_sdf(planes, dec, axes, start=0, count=n // 3)
    b = query_grid_sdf(planes, dec, axes, start=n // 3, count=n - n // 3)
    assert torch.equal(torch.cat([a, b]), sdf)
    pts = O.grid_points(O.grid_axes(d["mc_bound"], float(d["resolution"]))).to(DEV)
    raw = eval_points(pts, planes, dec)
    assert rel_err(raw[:, 3], d["sdf"]) < TOL_VAL and rel_err(raw[:, :3], d["rgb"]) < TOL_VAL


def test_cached_pose_backward_equals_recomputing_backward():
    """The tracker's backward pass on the activations the forward kernel kept (eslam_pose_backward_q on the Q images:
    no forward MLPs) gives the pose gradient and loss of the recomputing parameter-form kernel (eslam_loss_backward)
    on the same rays, samples and outlier mask, up to the re-association of the first layer's sums."""
    import ctypes as C
    from myslam_b200 import ReplayDraws
    from myslam_b200._lib import call, ptr, stream
    from myslam_b200.hotpath import tracking_iteration
    from myslam_b200.tracker import _tracker_state, _tracker_store

    fld, d = golden_field(), load_npz("tracking.npz")
    trk = make_tracker(fld, d)
    draws = recorded_draws(d)
    n_pix = int(d["n_pix"])
    st = _tracker_state(trk, n_pix)
    store = _tracker_store(trk, st)
    pose = torch.from_numpy(d["pose0"]).to(DEV).contiguous()
    gc, gd = torch.from_numpy(d["gt_color"]).to(DEV), torch.from_numpy(d["gt_depth"]).to(DEV)
    tracking_iteration(st["ws"], store, st["sc"], pose, gc, gd, n_pix, draws=ReplayDraws(draws[:2], DEV), strict_rng=True)
    ws, sc = st["ws"], st["sc"]
    g_cached, loss_cached = ws.grad7[0].clone(), ws.loss_acc[5].item()
    idx = draws[0].to(DEV)
    ws.loss_acc.zero_()
    call("eslam_loss_backward", store.ref(), ptr(store.arena), C.byref(sc.cam), C.byref(sc.render), ptr(ws.rays_o),
         ptr(ws.rays_d), ptr(ws.z), ptr(ws.gt_depth), ptr(ws.gt_color), ptr(ws.src), ptr(idx), n_pix, ptr(ws.ray_mask),
         ptr(ws.counters), None, n_pix, None, ptr(ws.pose_grad), ptr(ws.loss_acc), stream())
    call("eslam_finalize_loss", C.byref(sc.render), ptr(ws.counters), 1, ptr(ws.loss_acc), ptr(ws.loss_out), stream())
    call("eslam_pose_adam_step", ptr(pose), ptr(ws.pose_grad), None, None, 1, 0, 0.0, 0.0, 1, 0.5, 0.999, 1e-8,
         ptr(ws.grad7), 0, stream())
    assert rel_err(g_cached, ws.grad7[0]) < 1e-4
    assert abs(loss_cached - ws.loss_acc[5].item()) <= 1e-5 * abs(loss_cached)


def test_track_frame_graph_replay_matches_eager_loop(monkeypatch):
    """track_frame replays its iterations as one CUDA graph; with the same map, frame and initial pose it must
    behave like the launch-by-launch loop: finite, decreasing-on-average losses of the same size and a best pose
    within the step size of the eager one (the random pixels differ, so not bit for bit), and a replay on a
    DIFFERENT frame tensor must read the new frame (pointer table rewritten)."""
    fld, d = golden_field(), load_npz("tracking.npz")
    pose0 = torch.from_numpy(d["pose0"]).to(DEV)
    gc, gd = torch.from_numpy(d["gt_color"]).to(DEV), torch.from_numpy(d["gt_depth"]).to(DEV)
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("ESLAM_B200_GRAPH", mode)
        trk = make_tracker(fld, d)
        trk.strict_rng = False  # graphs need fixed draw shapes
        torch.manual_seed(5)
        best, losses, final = trk.track_frame(pose0, gc, gd, iters=6, batch_size=int(d["n_pix"]), lr_T=2e-3, lr_R=1e-3)
        torch.cuda.synchronize()
        assert torch.isfinite(losses).all() and torch.isfinite(final).all()
        assert (trk._b200.get("graph") is not None) == (mode == "1")
        res[mode] = (best.clone(), losses.clone(), final.clone(), trk)
    l0, l1 = res["0"][1], res["1"][1]
    assert abs(l0.mean().item() - l1.mean().item()) < 0.25 * abs(l0.mean().item())
    assert (res["0"][2] - res["1"][2]).abs().max().item() < 6 * 2e-3 * 2  # both within iters x lr of the start
    # replay on another frame object whose depth is all zero: the tracker keeps depth > 0 rays only (Tracker.py:182), so
    # if the rewritten pointer table is what the graph reads, no ray survives
    trk = res["1"][3]
    assert int(trk._b200["ws"].counters[0]) > 0
    trk.track_frame(pose0, gc.clone(), torch.zeros_like(gd), iters=6, batch_size=int(d["n_pix"]), lr_T=2e-3, lr_R=1e-3)
    assert int(trk._b200["ws"].counters[0]) == 0
    _, losses3, _ = trk.track_frame(pose0, gc, gd, iters=6, batch_size=int(d["n_pix"]), lr_T=2e-3, lr_R=1e-3)
    assert torch.isfinite(losses3).all() and int(trk._b200["ws"].counters[0]) > 0


def test_optimize_tracking_fused_step_keeps_the_callers_optimizer_consistent(monkeypatch):
    """optimize_tracking takes the step of the caller's torch.optim.Adam with the fused kernel when it recognises the
    (R, T) set-up of Tracker.run; parameters, .grad, moments and step counters must end up where autograd +
    optimizer.step() (ESLAM_B200_FUSED_OPT=0) puts them, and an optimizer it does not recognise goes the generic way."""
    from myslam_b200 import ReplayDraws
    from myslam_b200.tracker import _fused_adam_plan

    fld, d = golden_field(), load_npz("tracking.npz")
    pose0 = torch.from_numpy(d["pose0"]).to(DEV)
    gc, gd = torch.from_numpy(d["gt_color"]).to(DEV), torch.from_numpy(d["gt_depth"]).to(DEV)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("ESLAM_B200_FUSED_OPT", mode)
        trk = make_tracker(fld, d)
        trk.draws = ReplayDraws(recorded_draws(d), DEV)
        T = torch.nn.Parameter(pose0[:, -3:].clone())
        Rq = torch.nn.Parameter(pose0[:, :4].clone())
        opt = torch.optim.Adam([{"params": [T], "lr": float(d["lr_T"]), "betas": (0.5, 0.999)},
                                {"params": [Rq], "lr": float(d["lr_R"]), "betas": (0.5, 0.999)}])
        for _ in range(int(d["iters"])):
            pose = torch.cat([Rq, T], -1)
            if mode == "1":
                assert _fused_adam_plan(pose, opt) is not None
            trk.optimize_tracking(pose, gc, gd, int(d["n_pix"]), opt)
        out[mode] = (Rq.detach().clone(), T.detach().clone(), Rq.grad.clone(), T.grad.clone(),
                     opt.state[Rq]["exp_avg"].clone(), opt.state[T]["exp_avg_sq"].clone(), float(opt.state[Rq]["step"]),
                     float(opt.state[T]["step"]))
    for a, b in zip(out["0"][:6], out["1"][:6]):
        assert rel_err(b, a) < 1e-5
    assert out["0"][6:] == out["1"][6:] == (float(d["iters"]), float(d["iters"]))
    # not the (R, T) pattern: a single [1,7] parameter, SGD -> generic path
    P = torch.nn.Parameter(pose0.clone())
    assert _fused_adam_plan(P * 1.0, torch.optim.SGD([P], lr=1e-3)) is None
    assert _fused_adam_plan(torch.cat([P[:, :4], P[:, 4:]], -1), torch.optim.Adam([P], lr=1e-3)) is None


def test_optimize_tracking_iteration_graphs(monkeypatch):
    """Without injected draws the drop-in replays one CUDA graph per Adam step index.  Driven like Tracker.run (a new
    Adam per frame): the optimizer's state advances, the graphs captured for the first frame are re-used for the next,
    the losses are those of the launch-by-launch path on average, and a frame of zero depth really is read."""
    fld, d = golden_field(), load_npz("tracking.npz")
    pose0 = torch.from_numpy(d["pose0"]).to(DEV)
    gc, gd = torch.from_numpy(d["gt_color"]).to(DEV), torch.from_numpy(d["gt_depth"]).to(DEV)
    iters, n_pix = 6, int(d["n_pix"])

    def run_frame(trk, color, depth):
        T = torch.nn.Parameter(pose0[:, -3:].clone())
        Rq = torch.nn.Parameter(pose0[:, :4].clone())
        opt = torch.optim.Adam([{"params": [T], "lr": 2e-3, "betas": (0.5, 0.999)},
                                {"params": [Rq], "lr": 1e-3, "betas": (0.5, 0.999)}])
        losses = [trk.optimize_tracking(torch.cat([Rq, T], -1), color, depth, n_pix, opt) for _ in range(iters)]
        return torch.tensor(losses), torch.cat([Rq, T], -1).detach().clone(), opt, Rq, T

    means = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("ESLAM_B200_GRAPH", mode)
        trk = make_tracker(fld, d)
        trk.strict_rng = False  # graphs need fixed draw shapes
        torch.manual_seed(9)
        l_a, pose_a, opt, Rq, T = run_frame(trk, gc, gd)
        assert torch.isfinite(l_a).all() and torch.isfinite(pose_a).all()
        assert float(opt.state[Rq]["step"]) == iters and float(opt.state[T]["step"]) == iters
        assert opt.state[Rq]["exp_avg"].abs().sum() > 0 and torch.isfinite(Rq.grad).all()
        assert (pose_a - pose0).abs().max() > 0
        means[mode] = l_a.mean().item()
        if mode == "1":
            ig = trk._b200["iter_graphs"]
            n_graphs = len(ig.graphs)
            assert n_graphs == iters - 1  # the first iteration of the first frame ran launch by launch
            del opt
            l_b, _, opt2, Rq2, _ = run_frame(trk, gc.clone(), gd.clone())  # next frame: new optimizer, new tensors
            assert len(ig.graphs) == iters and torch.isfinite(l_b).all()
            assert opt2.state[Rq2]["exp_avg"].data_ptr() == ig.m7.data_ptr()
            del opt2
            run_frame(trk, gc, torch.zeros_like(gd))
            assert int(trk._b200["ws"].counters[0]) == 0  # depth > 0 rays only: the new frame was read
    assert abs(means["0"] - means["1"]) < 0.25 * abs(means["0"])


def test_sparse_adam_is_bit_identical_to_dense():
    """eslam_adam_step_sparse skips groups of 128 parameters whose gradient has been zero since the optimiser was
    created (torch's update leaves them unchanged); everything else -- parameters, moments, the zeroed gradient -- must
    equal the dense kernel bit for bit, over steps in which groups wake up at different times."""
    import ctypes as C
    from myslam_b200._lib import call, ptr, stream

    n = 128 * 700 + 64  # a ragged last group
    g = torch.Generator().manual_seed(3)
    p0 = torch.randn(n, generator=g)
    seg_end, seg_lr = (C.c_int64 * 2)(128 * 300 + 4, n), (C.c_double * 2)(5e-3, 1e-3)
    state = {k: [p0.clone().to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)] for k in ("dense", "sparse")}
    touched = torch.zeros((n + 127) // 128, dtype=torch.uint8, device=DEV)
    awake = torch.zeros(n // 128 + 1, dtype=torch.bool)
    for step in range(1, 7):
        awake |= torch.rand(awake.shape, generator=g) < 0.2      # more groups receive gradients over time
        grad = torch.randn(n, generator=g) * 0.1
        grad[~awake.repeat_interleave(128)[:n]] = 0
        grad[torch.rand(n, generator=g) < 0.3] = 0                  # zeros inside live groups too
        if step == 4:
            grad.zero_()                                            # a step without any gradient: momentum only
        for kind in ("dense", "sparse"):
            p, m, v = state[kind]
            gd = grad.clone().to(DEV)
            if kind == "dense":
                call("eslam_adam_step", ptr(p), ptr(gd), ptr(m), ptr(v), n, seg_end, seg_lr, 2, step, 0.9, 0.999, 1e-8,
                     stream())
            else:
                call("eslam_adam_step_sparse", ptr(p), ptr(gd), ptr(m), ptr(v), n, seg_end, seg_lr, 2, step, 0.9, 0.999,
                     1e-8, ptr(touched), stream())
            assert float(gd.abs().max()) == 0.0
        for a, b in zip(state["dense"], state["sparse"]):
            assert torch.equal(a, b), f"step {step}"
    t = touched.cpu().bool()
    assert 0 < int(t.sum()) < t.numel()
    assert torch.equal(state["sparse"][0].cpu()[~t.repeat_interleave(128)[:n]], p0[~t.repeat_interleave(128)[:n]])
