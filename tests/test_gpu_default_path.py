"""GPU parity of the DEFAULT path -- the one bench.py times: `strict_rng=False` (fixed-shape draws, no host sync) and
CUDA-graph replay.  The oracle is fed the SAME padded draws (tests/conftest.py: PaddedOracleDraws), so the comparison is
deterministic and held to north_star's bars: kept set / pixel indices / depth-guided z bit-exact, rendered values and
loss 1e-4, gradients 1e-3.  Also: consecutive launches with different decoders must each read their own (constant bank
refreshed by cudaMemcpyToSymbolAsync), eagerly and inside one graph."""
import numpy as np
import pytest
import torch

import eslam_oracle as O
from conftest import (GOLDEN_CAM, TRUNC, PaddedDeviceDraws, PaddedOracleDraws, arena_index, golden_field, load_npz,
                      rel_err, to_device_scene)
from test_gpu_parity import make_mapper, make_tracker

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _padded(n, seed, n_crop):
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(n_crop, (n,), generator=g)
    return idx, [torch.rand(n, 40, generator=g), torch.rand(n, 32, generator=g), torch.rand(n, 8, generator=g)]


def test_default_mapping_iteration_vs_oracle_on_the_same_padded_draws():
    from myslam_b200.common import matrix_to_cam_pose
    from myslam_b200.decoders import synced_store
    from myslam_b200.hotpath import mapping_iteration
    from myslam_b200.mapper import _mapper_state

    fld, d = golden_field(), load_npz("mapping.npz")
    H, W = GOLDEN_CAM[:2]
    idx, blocks = _padded(400, 17, H * W)
    c2ws = torch.from_numpy(d["c2ws0"])
    cols, deps = torch.from_numpy(d["gt_colors"]), torch.from_numpy(d["gt_depths"])
    f2 = fld.clone(requires_grad=True)
    pp = O.matrix_to_cam_pose(c2ws[1:]).clone().requires_grad_(True)
    cw = torch.cat([c2ws[0:1], O.cam_pose_to_matrix(pp)], 0)
    out = O.mapping_forward(f2, O.Camera(*GOLDEN_CAM), O.RenderCfg(32, 8, TRUNC), O.MAP_W, cw, cols, deps, 100,
                            PaddedOracleDraws(idx, blocks))
    out.loss.backward()
    assert int((out.gt_depth <= 0).sum()) > 0, "the fixture must exercise the depth-less branch"

    mp = make_mapper(fld, d)
    mp.strict_rng = False
    all_planes = (mp.planes_xy, mp.planes_xz, mp.planes_yz, mp.c_planes_xy, mp.c_planes_xz, mp.c_planes_yz)
    st = _mapper_state(mp, 400, 4)
    store = synced_store(all_planes, mp.decoders, mp.bound)
    store.reset_adam()
    poses7 = torch.zeros(4, 7, device=DEV)
    poses7[1:] = matrix_to_cam_pose(c2ws[1:].to(DEV))
    mapping_iteration(st["ws"], store, st["sc"], c2ws.to(DEV), poses7, cols.to(DEV), deps.to(DEV), 100, 1, 1e-3, 5e-3,
                      5e-3, 1e-3, draws=PaddedDeviceDraws(idx, blocks, DEV), strict_rng=False, want_loss=True,
                      apply_adam=False)
    ws = st["ws"]
    R = int(out.keep.sum())
    assert int(ws.counters[0]) == R
    assert torch.equal(ws.src[:R].cpu().long(), torch.nonzero(out.keep).squeeze(-1)), "kept pixels must be bit-exact"
    has = out.gt_depth > 0
    assert torch.equal(ws.z[:R].cpu()[has], out.z[has]), "depth-guided z_vals must be bit-exact"
    assert rel_err(ws.z[:R], out.z) < 1e-4
    assert abs(ws.loss_acc[5].item() - out.loss.item()) / abs(out.loss.item()) < 1e-4
    for k in range(12):
        assert rel_err(store.export_plane(arena_index(k), store.grad), f2.leaves()[k].grad) < 1e-3, f"plane {k}"
    gdec = store.dec_grad_dict(store.grad)
    for key in O.DECODER_KEYS:
        assert rel_err(gdec[key].reshape(f2.dec[key].shape), f2.dec[key].grad) < 1e-3, key
    assert rel_err(gdec["beta"], f2.beta.grad) < 1e-3
    assert rel_err(ws.grad7[1:4], pp.grad) < 1e-3


def _tracking_inputs():
    fld, d = golden_field(), load_npz("tracking.npz")
    n_pix = int(d["n_pix"])
    eh, ew = int(d["edge_h"]), int(d["edge_w"])
    H, W = GOLDEN_CAM[:2]
    idx, blocks = _padded(n_pix, 23, (H - 2 * eh) * (W - 2 * ew))
    return fld, d, n_pix, eh, ew, idx, blocks[:1]


def test_default_tracking_iteration_vs_oracle_and_graph_replay():
    """One default-path tracking iteration against the oracle on the same padded draws, then the SAME iteration captured
    as a CUDA graph and replayed: everything deterministic (kept rays, samples, rendered values, outlier mask) must be
    bit-identical to the eager run, loss and pose gradient (floating-point atomics) equal to 1e-6."""
    from myslam_b200.hotpath import tracking_iteration
    from myslam_b200.tracker import _tracker_state, _tracker_store

    fld, d, n_pix, eh, ew, idx, blocks = _tracking_inputs()
    pose0 = torch.from_numpy(d["pose0"])
    gc, gd = torch.from_numpy(d["gt_color"]), torch.from_numpy(d["gt_depth"])
    p_o = pose0.clone().requires_grad_(True)
    out = O.tracking_forward(fld, O.Camera(*GOLDEN_CAM), O.RenderCfg(32, 8, TRUNC), O.TRACK_W, p_o, gc, gd, n_pix, eh, ew,
                             PaddedOracleDraws(idx, blocks))
    out.loss.backward()

    trk = make_tracker(fld, d)
    trk.strict_rng = False
    st = _tracker_state(trk, n_pix)
    store = _tracker_store(trk, st)
    ws, sc = st["ws"], st["sc"]
    pose = pose0.to(DEV).contiguous()
    gcd, gdd = gc.to(DEV), gd.to(DEV)
    draws = PaddedDeviceDraws(idx, blocks, DEV)

    def run():
        tracking_iteration(ws, store, sc, pose, gcd, gdd, n_pix, draws=draws, strict_rng=False)

    run()
    torch.cuda.synchronize()
    R = int(out.keep.sum())
    assert int(ws.counters[0]) == R
    assert torch.equal(ws.src[:R].cpu().long(), torch.nonzero(out.keep).squeeze(-1))
    assert torch.equal(ws.z[:R].cpu(), out.z), "z_vals must be bit-exact"
    assert torch.equal(ws.ray_mask[:R].cpu().bool(), out.mask)
    assert rel_err(ws.depth[:R], out.depth) < 1e-4 and rel_err(ws.rgb[:R], out.rgb) < 1e-4
    assert abs(ws.loss_acc[5].item() - out.loss.item()) / abs(out.loss.item()) < 1e-4
    assert rel_err(ws.grad7[0:1], p_o.grad) < 1e-3
    keep = {k: getattr(ws, k).clone() for k in ("src", "z", "depth", "rgb", "sdf", "ray_mask", "counters")}
    loss_e, grad_e = ws.loss_acc[5].item(), ws.grad7[0].clone()
    # ---- the same launches as one CUDA graph
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        run()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    for name in ("z", "depth", "rgb", "sdf", "ray_mask"):
        getattr(ws, name).zero_()
    g.replay()
    torch.cuda.synchronize()
    for name, ref in keep.items():
        got = getattr(ws, name)
        n = R if name not in ("counters",) else ref.numel()
        assert torch.equal(got[:n], ref[:n]), f"graph replay differs from the eager run in {name}"
    assert abs(ws.loss_acc[5].item() - loss_e) <= 1e-6 * abs(loss_e)
    assert rel_err(ws.grad7[0], grad_e) < 1e-6


def test_consecutive_launches_read_their_own_decoders_eager_and_in_a_graph():
    """Regression for the constant-bank refresh (eslam_bind_decoders): decoders A, render, decoders B (very different),
    render -- each render must use its own weights, launched one by one and inside ONE CUDA graph, and equal the oracle
    to 1e-4; the two orders of execution must agree bit for bit."""
    from myslam_b200._lib import call, ptr, stream
    from myslam_b200.decoders import synced_store

    fld = golden_field()
    gsd = torch.Generator().manual_seed(5)
    decB = {k: (v * 3.0 + 0.05 * torch.randn(v.shape, generator=gsd)) for k, v in fld.dec.items()}
    fldB = O.Field(fld.planes, decB, fld.beta * 0.5, fld.bound)
    planes, dec = to_device_scene(fld)
    store = synced_store(planes, dec, fld.bound)
    n, S = 96, 40
    g = torch.Generator().manual_seed(1)
    b = fld.bound
    o = (b[:, 0] + (b[:, 1] - b[:, 0]) * (0.3 + 0.4 * torch.rand(n, 3, generator=g))).float()
    dd = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).float()
    z = torch.sort(0.05 + 0.6 * torch.rand(n, S, generator=g), dim=-1).values.float()
    ref = {name: O.composite(f, o, dd, z)[:2] for name, f in (("A", fld), ("B", fldB))}  # depth, rgb
    od, ddv, zd = o.to(DEV).contiguous(), dd.to(DEV).contiguous(), z.to(DEV).contiguous()
    arenaA = store.arena.clone()
    arenaB = store.arena.clone()
    dB = arenaB[store.dec_off:]
    from myslam_b200.field import DEC_LAYOUT
    for key, off, cnt in DEC_LAYOUT:
        src = (fldB.beta.reshape(1) if key == "beta" else decB[key].reshape(-1)).to(DEV)
        dB[off:off + cnt].copy_(src)
    outs = {k: (torch.empty(n, device=DEV), torch.empty(n, 3, device=DEV)) for k in ("A", "B")}

    def both():
        for name, arena in (("A", arenaA), ("B", arenaB)):
            call("eslam_bind_decoders", ptr(arena[store.dec_off:]), stream())
            call("eslam_render_forward", store.ref(), ptr(arena), ptr(od), ptr(ddv), ptr(zd), n, S, None,
                 ptr(outs[name][0]), ptr(outs[name][1]), None, stream())

    both()
    torch.cuda.synchronize()
    eager = {k: (v[0].clone(), v[1].clone()) for k, v in outs.items()}
    for name in ("A", "B"):
        assert rel_err(eager[name][0], ref[name][0]) < 1e-4 and rel_err(eager[name][1], ref[name][1]) < 1e-4, name
    assert not torch.equal(eager["A"][0], eager["B"][0])
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        both()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        both()
    for v in outs.values():
        v[0].zero_()
        v[1].zero_()
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    for name in ("A", "B"):
        assert torch.equal(outs[name][0], eager[name][0]) and torch.equal(outs[name][1], eager[name][1]), name


def test_mapping_window_graph_replay_matches_kernel_by_kernel(monkeypatch):
    """map_window on the default path three times from the SAME parameters, window and generator seed: launched kernel
    by kernel (ESLAM_B200_GRAPH=0), then twice with graphs on (the first call of a shape is launched kernel by kernel,
    the second is captured and replayed).  Torch's generator is graph safe, so all three draw the same pixels and
    uniforms; planes, decoders and optimised poses must agree to the noise of the float atomics."""
    from myslam_b200.decoders import synced_store
    from myslam_b200.hotpath import FrameTable
    from myslam_b200.mapper import _WindowGraph, _mapper_state, map_window

    fld, d = golden_field(), load_npz("mapping.npz")
    c2ws = torch.from_numpy(d["c2ws0"]).to(DEV)
    cols, deps = torch.from_numpy(d["gt_colors"]).to(DEV), torch.from_numpy(d["gt_depths"]).to(DEV)
    mp = make_mapper(fld, d)
    mp.strict_rng = False
    all_planes = (mp.planes_xy, mp.planes_xz, mp.planes_yz, mp.c_planes_xy, mp.c_planes_xz, mp.c_planes_yz)
    st = _mapper_state(mp, 400, 4)
    store = synced_store(all_planes, mp.decoders, mp.bound)
    arena0 = store.arena.clone()

    def run(frames):
        store.arena.copy_(arena0)
        store.gen += 1
        torch.manual_seed(99)
        out = map_window(store, st["ws"], st["sc"], c2ws, frames, frames, 400, 3, 1e-3, 5e-3, 5e-3, True, 1e-3)
        torch.cuda.synchronize()
        return store.arena.clone(), out.clone()

    table = lambda: FrameTable([cols[k] for k in range(4)], [deps[k] for k in range(4)], st["sc"].cam, DEV)
    monkeypatch.setenv("ESLAM_B200_GRAPH", "0")
    a_ref, p_ref = run(table())
    a_ref2, p_ref2 = run(table())  # the same again: run-to-run noise of the float atomics (amplified by Adam's g/|g|)
    assert not getattr(st["ws"], "_window_graphs", None)
    monkeypatch.setenv("ESLAM_B200_GRAPH", "1")
    a_1, p_1 = run(table())
    a_2, p_2 = run(table())  # captured + replayed; a different table object whose pointers are copied in
    graphs = st["ws"]._window_graphs
    assert len(graphs) == 1 and isinstance(next(iter(graphs.values())), _WindowGraph)
    a_3, p_3 = run(table())  # replayed again
    upd = (a_ref - arena0).abs()
    assert float(upd.max()) > 1e-4, "the call must have moved the parameters"
    noise_max = float((a_ref2 - a_ref).abs().max())
    noise_mean = float((a_ref2 - a_ref).abs().mean())
    for a, p in ((a_1, p_1), (a_2, p_2), (a_3, p_3)):
        diff = (a - a_ref).abs()
        assert float(diff.max()) <= max(8 * noise_max, 1e-5 * float(upd.max())), (float(diff.max()), noise_max)
        assert float(diff.mean()) <= max(4 * noise_mean, 1e-6 * float(upd.mean())), (float(diff.mean()), noise_mean)
        assert float(diff.mean()) < 1e-3 * float(upd.mean())
        assert float((p - p_ref).abs().max()) < max(1e-5, 8 * float((p_ref2 - p_ref).abs().max()))
    assert float((p_ref[1:] - c2ws[1:]).abs().max()) > 0, "joint optimisation must have moved the poses"
    assert torch.equal(p_ref[0], c2ws[0])
