"""GPU, BASELINE.json's full shapes (Replica room0 1200x680 S=40; ScanNet scene0000 620x460 S=56): parity against the
oracle on a bounded number of rays with seeded random fields at the DEFAULT plane resolutions (24/6/3 cm), and
size-independent properties at the full ray counts (gradient linearity in the upstream gradient, Adam zero-gradient
fixed point, exact mesh-query sharding)."""
import numpy as np
import pytest
import torch

import eslam_oracle as O
from conftest import arena_index, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def build(spec, seed=0):
    import myslam_b200 as M
    from myslam_b200 import synthetic as S

    gen = torch.Generator().manual_seed(seed)
    bound = O.rounded_bound(spec["bound"], spec["bound_dividable"])
    fld = O.make_field(bound, spec["planes_res"], spec["c_planes_res"], generator=gen, std=0.05, dec_scale=1.5)
    planes = tuple([p.clone().to(DEV) for p in g] for g in fld.planes)
    dec = M.Decoders(c_dim=32, truncation=spec["truncation"], learnable_beta=True)
    dec.load_state_dict({**fld.dec, "beta": fld.beta})
    dec = dec.to(DEV)
    dec.bound = bound.clone()
    cam = (spec["H"], spec["W"], spec["fx"], spec["fy"], spec["cx"], spec["cy"])
    cfg = S.run_cfg(spec)

    class E:
        pass

    e = E()
    e.bound, e.device = bound, DEV
    e.H, e.W, e.fx, e.fy, e.cx, e.cy = cam
    rnd = M.Renderer(cfg, e)
    return fld, planes, dec, cam, cfg, rnd, gen


@pytest.mark.parametrize("name", ["REPLICA_ROOM0", "SCANNET_0000"])
def test_mapping_iteration_full_shapes_vs_oracle(name):
    """One mapping iteration (b=3 frames, 3 x 400 rays) at the dataset's real image size, plane shapes and sample
    count: kept set and depth-guided z bit-exact, loss 1e-4, all plane / decoder / pose gradients 1e-3."""
    import myslam_b200 as M
    from myslam_b200 import synthetic as S
    from myslam_b200.common import matrix_to_cam_pose
    from myslam_b200.decoders import synced_store
    from myslam_b200.hotpath import mapping_iteration
    from myslam_b200.mapper import _mapper_state

    spec = getattr(S, name)
    fld, planes, dec, cam, cfg, rnd, gen = build(spec)
    poses = S.trajectory(3, spec["room"], step_deg=6.0)
    frames = [S.render_box_room(p, *cam, spec["room"], "cpu", hole_frac=0.05, generator=gen) for p in poses]
    cols, deps = torch.stack([f[0] for f in frames], 0), torch.stack([f[1] for f in frames], 0)
    ocam = O.Camera(*cam)
    rc = O.RenderCfg(spec["n_stratified"], spec["n_importance"], spec["truncation"])
    n_per = 400
    f2 = fld.clone(requires_grad=True)
    pp = O.matrix_to_cam_pose(poses[1:]).clone().requires_grad_(True)
    cw = torch.cat([poses[0:1], O.cam_pose_to_matrix(pp)], 0)
    live = O.LiveDraws(gen)
    out = O.mapping_forward(f2, ocam, rc, O.MAP_W, cw, cols, deps, n_per, live)
    out.loss.backward()

    mp = M.MapperStep(cfg, rnd, dec, planes, fld.bound.clone(), cam, DEV)
    st = _mapper_state(mp, 3 * n_per, 3)
    store = synced_store(planes, dec, fld.bound)
    store.reset_adam()
    poses7 = torch.zeros(3, 7, device=DEV)
    poses7[1:] = matrix_to_cam_pose(poses[1:].to(DEV))
    mapping_iteration(st["ws"], store, st["sc"], poses.to(DEV), poses7, cols.to(DEV), deps.to(DEV), n_per, 1, 1e-3, 5e-3,
                      5e-3, 1e-3, draws=M.ReplayDraws(live.log, DEV), strict_rng=True, want_loss=True, apply_adam=False)
    ws = st["ws"]
    R = int(out.keep.sum())
    assert int(ws.counters[0]) == R
    assert torch.equal(ws.src[:R].cpu().long(), torch.nonzero(out.keep).squeeze(-1))
    has = out.gt_depth > 0
    assert torch.equal(ws.z[:R].cpu()[has], out.z[has])
    assert rel_err(ws.z[:R], out.z) < 1e-4
    assert abs(ws.loss_acc[5].item() - out.loss.item()) / abs(out.loss.item()) < 1e-4
    for k in range(12):
        assert rel_err(store.export_plane(arena_index(k), store.grad), f2.leaves()[k].grad) < 1e-3, f"plane {k}"
    gdec = store.dec_grad_dict(store.grad)
    for key in O.DECODER_KEYS:
        assert rel_err(gdec[key].reshape(f2.dec[key].shape), f2.dec[key].grad) < 1e-3, key
    assert rel_err(gdec["beta"], f2.beta.grad) < 1e-3
    assert rel_err(ws.grad7[1:3], pp.grad) < 1e-3


def test_tracking_iteration_replica_full_shape_vs_oracle():
    import myslam_b200 as M
    from myslam_b200 import synthetic as S
    from myslam_b200.hotpath import tracking_iteration
    from myslam_b200.tracker import _tracker_state, _tracker_store

    spec = S.REPLICA_ROOM0
    fld, planes, dec, cam, cfg, rnd, gen = build(spec, seed=3)
    pose_m = S.trajectory(1, spec["room"])
    col, dep = S.render_box_room(pose_m[0], *cam, spec["room"], "cpu", hole_frac=0.03, generator=gen)
    pose0 = O.matrix_to_cam_pose(pose_m)
    t = spec["tracking"]
    live = O.LiveDraws(gen)
    p_o = pose0.clone().requires_grad_(True)
    out = O.tracking_forward(fld, O.Camera(*cam), O.RenderCfg(32, 8, 0.06), O.TRACK_W, p_o, col[None], dep[None],
                             t["pixels"], t["ignore_edge_H"], t["ignore_edge_W"], live)
    out.loss.backward()
    trk = M.TrackerStep(cfg, rnd, dec, planes, fld.bound.clone(), cam, DEV)
    st = _tracker_state(trk, t["pixels"])
    store = _tracker_store(trk, st)
    tracking_iteration(st["ws"], store, st["sc"], pose0.to(DEV).contiguous(), col[None].to(DEV), dep[None].to(DEV),
                       t["pixels"], draws=M.ReplayDraws(live.log, DEV), strict_rng=True)
    ws = st["ws"]
    R = int(out.keep.sum())
    assert int(ws.counters[0]) == R and R > 1500
    assert torch.equal(ws.z[:R].cpu(), out.z)
    assert torch.equal(ws.ray_mask[:R].cpu().bool(), out.mask)
    assert rel_err(ws.depth[:R], out.depth) < 1e-4 and rel_err(ws.rgb[:R], out.rgb) < 1e-4
    assert abs(ws.loss_acc[5].item() - out.loss.item()) / abs(out.loss.item()) < 1e-4
    assert rel_err(ws.grad7[0:1], p_o.grad) < 1e-3


def test_backward_is_linear_in_upstream_gradient_at_full_batch():
    """Property at the full 4000-ray batch: grad(a*g1 + b*g2) == a*grad(g1) + b*grad(g2) for the plane arena."""
    import ctypes as C
    from myslam_b200 import synthetic as S
    from myslam_b200._lib import call, ptr, stream
    from myslam_b200.decoders import synced_store

    spec = S.REPLICA_ROOM0
    fld, planes, dec, cam, cfg, rnd, gen = build(spec, seed=5)
    store = synced_store(planes, dec, fld.bound)
    R, Sn = 4000, 40
    g = torch.Generator().manual_seed(1)
    lo, hi = fld.bound[:, 0], fld.bound[:, 1]
    ro = (lo + (hi - lo) * (0.3 + 0.4 * torch.rand(R, 3, generator=g))).to(DEV)
    rd = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1).to(DEV)
    z = torch.sort(torch.rand(R, Sn, generator=g) * 1.5, -1)[0].to(DEV).contiguous()

    def grads(gd, gc, gs):
        ga = torch.zeros_like(store.arena)
        call("eslam_render_backward", store.ref(), ptr(store.arena), ptr(ro), ptr(rd), ptr(z), R, Sn, ptr(gd), ptr(gc),
             ptr(gs), ptr(ga), None, None, stream())
        return ga

    mk = lambda *s: torch.randn(*s, generator=g).to(DEV)
    g1, g2 = (mk(R), mk(R, 3), mk(R, Sn)), (mk(R), mk(R, 3), mk(R, Sn))
    a, b = 0.7, -1.3
    lhs = grads(*[a * x + b * y for x, y in zip(g1, g2)])
    rhs = a * grads(*g1) + b * grads(*g2)
    assert float(rhs.abs().max()) > 0
    assert rel_err(lhs, rhs) < 1e-4


def test_adam_zero_gradient_is_a_fixed_point_and_dense_momentum_keeps_moving():
    """Adam semantics the mapper relies on (SURVEY 7 'Adam semantics'): untouched parameters never move, touched
    ones keep moving on later zero-gradient steps (dense update)."""
    import ctypes as C
    from myslam_b200._lib import call, ptr, stream

    n = 1 << 20
    p = torch.randn(n, device=DEV)
    p0 = p.clone()
    m, v, g = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    g[: n // 2] = 0.01
    for step in (1, 2, 3):
        call("eslam_adam_step", ptr(p), ptr(g), ptr(m), ptr(v), n, (C.c_int64 * 1)(n), (C.c_double * 1)(5e-3), 1, step,
             0.9, 0.999, 1e-8, stream())
    assert torch.equal(p[n // 2:], p0[n // 2:])
    moved = (p[: n // 2] - p0[: n // 2]).abs()
    assert float(moved.min()) > 5e-3 * 1.5, "momentum must keep moving touched parameters after the gradient is zeroed"


def test_mesh_query_sharding_is_exact_at_replica_resolution_slab():
    """A 4-slab of the 1 cm Replica lattice (990x680x490): 8-way sharded query == single query, bit for bit."""
    from myslam_b200 import grid_axes, query_grid_sdf, synthetic as S
    from myslam_b200.dist import shard_range

    spec = S.REPLICA_ROOM0
    fld, planes, dec, cam, cfg, rnd, gen = build(spec, seed=7)
    axes = grid_axes(spec["bound"], 0.01)
    assert [len(a) for a in axes] == [990, 680, 490]
    start, count = 990 * 490 * 300, 990 * 490 * 4  # four iy-rows of the lattice in the middle of the volume
    whole = query_grid_sdf(planes, dec, axes, start=start, count=count)
    parts = [query_grid_sdf(planes, dec, axes, start=start + s, count=c)
             for s, c in (shard_range(count, r, 8) for r in range(8))]
    assert torch.equal(torch.cat(parts), whole)
    # against the oracle on a random subset of those lattice points
    idx = torch.randint(0, count, (3000,)) + start
    iz, t = idx % 490, idx // 490
    ix, iy = t % 990, t // 990
    pts = torch.stack([torch.from_numpy(axes[0]).float()[ix], torch.from_numpy(axes[1]).float()[iy],
                       torch.from_numpy(axes[2]).float()[iz]], 1)
    ref = O.query_points(fld, pts)[:, -1]
    assert rel_err(whole.cpu()[idx - start], ref) < 1e-4
