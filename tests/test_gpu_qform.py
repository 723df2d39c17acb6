"""The Q form (the first decoder layer applied to the planes) against the parameter form of the same
kernels and against the reference's golden gradients: render forward, the tracker's pose-only backward, the mapping
backward into gradient images, and the optimiser tail that consumes them."""
import pytest
import torch

from conftest import golden_field, to_device_scene

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rays(fld, n, S, seed=0):
    g = torch.Generator().manual_seed(seed)
    b = fld.bound
    o = b[:, 0] + (b[:, 1] - b[:, 0]) * (0.3 + 0.4 * torch.rand(n, 3, generator=g))
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    z = torch.sort(0.05 + 2.5 * torch.rand(n, S, generator=g), dim=-1).values  # leaves the bound for part of the rays
    return o.float().contiguous().to(DEV), d.float().contiguous().to(DEV), z.float().contiguous().to(DEV)


@pytest.mark.parametrize("S", [40, 56, 13])
def test_render_forward_on_preactivated_planes_matches_render_forward(S):
    from myslam_b200._lib import call, ptr, stream
    from myslam_b200.decoders import synced_store

    fld = golden_field()
    planes, dec = to_device_scene(fld)
    store = synced_store(planes, dec, fld.bound)
    n = 257
    o, d, z = _rays(fld, n, S)
    out = {}
    for name in ("ref", "q"):
        out[name] = dict(depth=torch.empty(n, device=DEV), rgb=torch.empty(n, 3, device=DEV),
                         sdf=torch.empty(n, S, device=DEV), act4=torch.empty(n, S, 4, device=DEV),
                         actm=torch.empty(n, S, dtype=torch.int32, device=DEV))
    r = out["ref"]
    call("eslam_render_forward_act", store.ref(), ptr(store.arena), ptr(o), ptr(d), ptr(z), n, S, None, ptr(r["depth"]),
         ptr(r["rgb"]), ptr(r["sdf"]), ptr(r["act4"]), ptr(r["actm"]), stream())
    q_arena = torch.zeros(store.n_planes_end // 2, dtype=torch.float32, device=DEV)
    call("eslam_q_build", store.ref(), ptr(store.arena), ptr(q_arena), stream())
    r = out["q"]
    call("eslam_render_forward_q", store.ref(), ptr(q_arena), ptr(o), ptr(d), ptr(z), n, S, None, ptr(r["depth"]),
         ptr(r["rgb"]), ptr(r["sdf"]), ptr(r["act4"]), ptr(r["actm"]), stream())
    torch.cuda.synchronize()
    for k in ("depth", "rgb", "sdf"):
        err = (out["q"][k] - out["ref"][k]).abs().max().item()
        assert err < 1e-5, (k, err)
    assert (out["q"]["act4"][..., :3] - out["ref"]["act4"][..., :3]).abs().max().item() < 1e-5
    # ReLU masks may flip only where a pre-activation sits within rounding of zero
    flips = (out["q"]["actm"] != out["ref"]["actm"]).float().mean().item()
    assert flips < 1e-3, flips


def test_pose_backward_on_preactivated_planes_matches_the_parameter_form():
    """Tracker iteration on the golden frame twice: the product path (eslam_render_forward_q + eslam_pose_backward_q
    on the Q images) and the parameter-form kernels (eslam_render_forward_act + eslam_pose_backward_act) on the same
    rays, samples and outlier mask; the pose gradient and the loss must agree to the re-association of the first
    layer's sums."""
    import ctypes as C

    from conftest import load_npz, recorded_draws, rel_err
    from myslam_b200 import ReplayDraws
    from myslam_b200._lib import call, ptr, stream
    from myslam_b200.hotpath import tracking_iteration
    from myslam_b200.tracker import _tracker_state, _tracker_store
    from test_gpu_parity import make_tracker

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
    g_q, loss_q = ws.grad7[0].clone(), ws.loss_acc[5].item()
    mask_q = ws.ray_mask.clone()
    idx = draws[0].to(DEV)
    S = sc.render.n_stratified + sc.render.n_importance
    ws.loss_acc.zero_()
    call("eslam_render_forward_act", store.ref(), ptr(store.arena), ptr(ws.rays_o), ptr(ws.rays_d), ptr(ws.z), n_pix, S,
         ptr(ws.counters), ptr(ws.depth), ptr(ws.rgb), ptr(ws.sdf), ptr(ws.act4), ptr(ws.actm), stream())
    call("eslam_track_mask", ptr(ws.gt_depth), ptr(ws.depth), ptr(ws.band), n_pix, ptr(ws.counters), ptr(ws.ray_mask),
         ptr(ws.scratch), stream())
    call("eslam_pose_backward_act", store.ref(), ptr(store.arena), C.byref(sc.cam), C.byref(sc.render),
         ptr(ws.rays_o), ptr(ws.rays_d), ptr(ws.z), ptr(ws.gt_depth), ptr(ws.gt_color), ptr(ws.src), ptr(idx), n_pix,
         ptr(ws.ray_mask), ptr(ws.counters), n_pix, ptr(ws.sdf), ptr(ws.act4), ptr(ws.actm), ptr(ws.pose_grad),
         ptr(ws.loss_acc), stream())
    call("eslam_finalize_loss", C.byref(sc.render), ptr(ws.counters), 1, ptr(ws.loss_acc), ptr(ws.loss_out), stream())
    call("eslam_pose_adam_step", ptr(pose), ptr(ws.pose_grad), None, None, 1, 0, 0.0, 0.0, 1, 0.5, 0.999, 1e-8,
         ptr(ws.grad7), 0, stream())
    torch.cuda.synchronize()
    assert torch.equal(ws.ray_mask, mask_q), "the outlier mask must not depend on the form of the forward"
    assert rel_err(g_q, ws.grad7[0]) < 1e-4
    assert abs(ws.loss_acc[5].item() - loss_q) <= 1e-5 * abs(loss_q)


def test_q_adam_planes_matches_the_dense_chain_rule_and_adam():
    """Two optimiser steps of the Q form's dense tail on sparse gradient images against plain torch:
    dplane = GQ . W1_half, dW1_half = sum GQ (x) plane, Adam on the planes (oracle formula = torch.optim.Adam), gradient
    images zeroed, untouched texels bit-identical."""
    import eslam_oracle as O
    from myslam_b200._lib import call, load, ptr, stream
    from myslam_b200.decoders import synced_store

    fld = golden_field()
    planes, dec = to_device_scene(fld)
    store = synced_store(planes, dec, fld.bound)
    store.reset_adam()
    lib = load()
    touched = torch.zeros(lib.eslam_q_touched_bytes(store.ref()), dtype=torch.uint8, device=DEV)
    gq = torch.zeros(store.n_planes_end // 2, dtype=torch.float32, device=DEV)
    W1 = {0: store.dec[0:1024].view(16, 64).clone(), 1: store.dec[1332:1332 + 1024].view(16, 64).clone()}
    ref_p = [store.plane_view(i).clone() for i in range(12)]
    ref_m = [torch.zeros_like(p) for p in ref_p]
    ref_v = [torch.zeros_like(p) for p in ref_p]
    arena0 = store.arena.clone()
    g = torch.Generator(device="cpu").manual_seed(3)
    lr = (5e-3, 2e-3)
    hit = [torch.zeros(p.shape[:2], dtype=torch.bool, device=DEV) for p in ref_p]
    for step in (1, 2):
        dW1 = {0: torch.zeros(16, 64, device=DEV), 1: torch.zeros(16, 64, device=DEV)}
        for i in range(12):
            h, w = store.shapes[i]
            view = gq[store.plane_off[i] // 2: store.plane_off[i] // 2 + h * w * 16].view(h, w, 16)
            mask = (torch.rand(h, w, generator=g) < (0.2 if step == 1 else 0.1)).to(DEV)
            view[mask] = torch.randn(int(mask.sum()), 16, generator=g).to(DEV) * 1e-2
            hit[i] |= mask
            fld_i, sc = i // 6, (i % 6) // 3
            Wh = W1[fld_i][:, sc * 32:(sc + 1) * 32]
            grad = view @ Wh
            dW1[fld_i][:, sc * 32:(sc + 1) * 32] += torch.einsum("hwj,hwc->jc", view, ref_p[i])
            O.adam_update(ref_p[i], grad, ref_m[i], ref_v[i], step, lr[fld_i])
        store.grad.zero_()
        call("eslam_q_adam_planes", store.ref(), ptr(store.arena), ptr(gq), ptr(store.exp_avg), ptr(store.exp_avg_sq),
             ptr(store.grad), ptr(touched), lr[0], lr[1], step, 0.9, 0.999, 1e-8, stream())
        torch.cuda.synchronize()
        assert float(gq.abs().max()) == 0.0, "consumed gradient images must be zeroed"
        gdec = store.grad[store.dec_off:]
        for fld_i, off in ((0, 0), (1, 1332)):
            got = gdec[off:off + 1024].view(16, 64)
            assert (got - dW1[fld_i]).abs().max().item() <= 1e-5 * dW1[fld_i].abs().max().item()
    for i in range(12):
        assert (store.plane_view(i) - ref_p[i]).abs().max().item() <= 1e-5 * ref_p[i].abs().max().item()
        assert (store.plane_view(i, store.exp_avg) - ref_m[i]).abs().max().item() <= 1e-5 * ref_m[i].abs().max().item()
        assert (store.plane_view(i, store.exp_avg_sq) - ref_v[i]).abs().max().item() <= 1e-5 * ref_v[i].abs().max().item()
        # texels never hit keep their initial bits (exact skip)
        before = arena0[store.plane_off[i]: store.plane_off[i] + ref_p[i].numel()].view_as(ref_p[i])
        assert torch.equal(store.plane_view(i)[~hit[i]], before[~hit[i]])


def test_mapping_backward_in_the_q_form_matches_the_reference_gradients():
    """The golden mapping iteration (reference gradients from tests/golden/mapping.npz) through the Q form:
    eslam_q_build -> eslam_loss_backward_q.  Plane gradients are recovered from the gradient images as GQ . W1_half,
    dW1 as sum GQ (x) plane (what eslam_q_adam_planes consumes), the other decoder gradients, beta, poses and the loss
    come out of the kernel directly; all held to the product path's bars (1e-4 loss, 1e-3 gradients)."""
    import ctypes as C

    import eslam_oracle as O
    from conftest import arena_index, load_npz, recorded_draws, rel_err
    from myslam_b200 import ReplayDraws
    from myslam_b200._lib import call, ptr, stream
    from myslam_b200.common import matrix_to_cam_pose
    from myslam_b200.decoders import synced_store
    from myslam_b200.hotpath import mapping_iteration
    from myslam_b200.mapper import _mapper_state
    from test_gpu_parity import make_mapper

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
    # the iteration itself first: it leaves the compacted rays, samples and counters in the workspace
    mapping_iteration(st["ws"], store, st["sc"], c2ws, poses7, gc, gd, 100, 1, 1e-3, 5e-3, 5e-3, 1e-3,
                      draws=ReplayDraws(draws[:4], DEV), strict_rng=True, want_loss=True, apply_adam=False)
    ws, sc = st["ws"], st["sc"]
    idx = draws[0].to(DEV)
    N = 400
    q_arena = torch.zeros(store.n_planes_end // 2, dtype=torch.float32, device=DEV)
    gq = torch.zeros_like(q_arena)
    call("eslam_q_build", store.ref(), ptr(store.arena), ptr(q_arena), stream())
    store.grad.zero_()
    ws.pose_grad.zero_()
    ws.loss_acc.zero_()
    call("eslam_loss_backward_q", store.ref(), ptr(store.arena), ptr(q_arena), ptr(gq), C.byref(sc.cam),
         C.byref(sc.render), ptr(ws.rays_o), ptr(ws.rays_d), ptr(ws.z), ptr(ws.gt_depth), ptr(ws.gt_color), ptr(ws.src),
         ptr(idx), 100, None, ptr(ws.counters), None, N, ptr(store.grad), ptr(ws.pose_grad), ptr(ws.loss_acc), stream())
    call("eslam_finalize_loss", C.byref(sc.render), ptr(ws.counters), 0, ptr(ws.loss_acc), ptr(ws.loss_out), stream())
    call("eslam_pose_adam_step", ptr(poses7), ptr(ws.pose_grad), None, None, 4, 1, 0.0, 0.0, 1, 0.9, 0.999, 1e-8,
         ptr(ws.grad7), 0, stream())
    torch.cuda.synchronize()
    assert abs(ws.loss_acc[5].item() - float(d["it0_loss"])) / abs(float(d["it0_loss"])) < 1e-4
    W1 = {0: store.dec[0:1024].view(16, 64), 1: store.dec[1332:1332 + 1024].view(16, 64)}
    dW1 = {0: torch.zeros(16, 64, device=DEV), 1: torch.zeros(16, 64, device=DEV)}
    for k in range(12):
        i = arena_index(k)
        h, w = store.shapes[i]
        G = gq[store.plane_off[i] // 2: store.plane_off[i] // 2 + h * w * 16].view(h, w, 16)
        f_i, s_i = i // 6, (i % 6) // 3
        dplane = (G @ W1[f_i][:, s_i * 32:(s_i + 1) * 32]).permute(2, 0, 1)[None]  # NCHW like the reference's gradient
        assert rel_err(dplane, d[f"it0_d_plane.{k}"]) < 1e-3, f"plane {k}"
        dW1[f_i][:, s_i * 32:(s_i + 1) * 32] += torch.einsum("hwj,hwc->jc", G, store.plane_view(i))
    gdec = store.dec_grad_dict(store.grad)
    for name in O.DECODER_KEYS:
        ref = d[f"it0_d_dec.{name}"]
        if name == "linears.0.weight":
            got = dW1[0]
        elif name == "c_linears.0.weight":
            got = dW1[1]
        else:
            got = gdec[name].reshape(ref.shape)
        assert rel_err(got, ref) < 1e-3, name
    assert rel_err(gdec["beta"], d["it0_beta_grad"]) < 1e-3
    assert rel_err(ws.grad7[1:4], d["it0_pose_grad"]) < 1e-3
    # ---- the same backward as a PAIR of launches (eslam_loss_backward_q_part: tiles without / with a depth-less ray),
    # the way the pipelined window loop issues it: every tile belongs to exactly one part, so the sums are the same
    n_dl = int(ws.counters[0]) - int((ws.gt_depth[: int(ws.counters[0])] > 0).sum())
    assert n_dl > 0, "the fixture must hold depth-less rays, or part 2 is empty"
    def launch(parts):
        out = torch.zeros_like(gq)
        store.grad.zero_()
        ws.pose_grad.zero_()
        after = []
        for part in parts:
            call("eslam_loss_backward_q_part", store.ref(), ptr(store.arena), ptr(q_arena), ptr(out), C.byref(sc.cam),
                 C.byref(sc.render), ptr(ws.rays_o), ptr(ws.rays_d), ptr(ws.z), ptr(ws.gt_depth), ptr(ws.gt_color),
                 ptr(ws.src), ptr(idx), 100, ptr(ws.dl_list), ptr(ws.counters), None, N, ptr(store.grad),
                 ptr(ws.pose_grad), part, stream())
            torch.cuda.synchronize()
            after.append(out.clone())
        return out, store.grad.clone(), ws.pose_grad.clone(), after

    gq_one, grad_one, pose_one, _ = launch((0,))
    gq_pair, grad_pair, pose_pair, seen = launch((1, 2))
    assert rel_err(gq_one, gq) < 1e-5
    assert float(seen[0].abs().max()) > 0 and float((seen[1] - seen[0]).abs().max()) > 0, "both parts must contribute"
    assert rel_err(gq_pair, gq_one) < 1e-5
    assert rel_err(grad_pair[store.dec_off:], grad_one[store.dec_off:]) < 1e-5
    assert rel_err(pose_pair, pose_one) < 1e-5
