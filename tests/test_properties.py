"""CPU: size-independent properties of the path (oracle side) and of the host logic, beyond the golden vectors:
sortedness and ranges of the depth samples on ragged inputs (rays with and without depth, none of either),
compositing weights, the convex mesh bound's half-spaces, store signatures, shard ranges."""
import numpy as np
import torch

import eslam_oracle as O
from conftest import golden_field


def _rays(fld, n, seed):
    g = torch.Generator().manual_seed(seed)
    b = fld.bound
    o = b[:, 0] + (b[:, 1] - b[:, 0]) * (0.3 + 0.4 * torch.rand(n, 3, generator=g))
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    return o.float(), d.float(), g


def test_depth_samples_are_sorted_and_bounded_on_ragged_input():
    fld = golden_field()
    tr, ns, ni = 0.06, 32, 8
    for n_with, n_without in ((7, 5), (9, 0), (0, 6), (0, 0)):
        n = n_with + n_without
        o, d, g = _rays(fld, n, seed=n_with * 10 + n_without)
        gt = torch.cat([0.5 + 2.0 * torch.rand(n_with, generator=g), torch.zeros(n_without)])
        perm = torch.randperm(n, generator=g) if n else torch.zeros(0, dtype=torch.long)
        gt = gt[perm]
        z = O.ray_depths(fld, o, d, gt, tr, ns, ni, O.LiveDraws(g))
        assert z.shape == (n, ns + ni)
        if n == 0:
            continue
        assert bool((z[:, 1:] >= z[:, :-1]).all()), "samples along a ray are sorted"
        has = gt > 0
        if bool(has.any()):
            zh, dh = z[has], gt[has][:, None]
            assert bool((zh >= torch.minimum(torch.zeros_like(dh), dh - 1.5 * tr) - 1e-6).all())
            assert bool((zh <= torch.maximum(1.2 * dh, dh + 1.5 * tr) + 1e-6).all())
        if bool((~has).any()):
            assert bool((z[~has] >= 0).all()) and bool(torch.isfinite(z[~has]).all())


def test_compositing_weights_are_a_sub_probability():
    g = torch.Generator().manual_seed(1)
    sdf = torch.randn(50, 40, generator=g) * 0.3
    for beta in (2.0, 10.0, 40.0):
        w = O.transmittance_weights(O.sdf2alpha(sdf, torch.tensor([beta])))
        assert bool((w >= 0).all()) and bool((w.sum(-1) <= 1.0 + 1e-5).all())
    z = torch.sort(torch.rand(50, 40, generator=g) * 3, -1).values
    depth = (w * z).sum(-1)
    assert bool((depth <= z[:, -1] + 1e-5).all()) and bool((depth >= 0).all())


def test_hull_half_spaces_contain_exactly_the_hull():
    from scipy.spatial import ConvexHull
    from myslam_b200.mesher import hull_planes

    rng = np.random.default_rng(0)
    pts = rng.normal(size=(60, 3))
    hull = ConvexHull(pts)
    hp = hull_planes(pts, hull.simplices).double().numpy()
    assert hp.shape == (len(hull.simplices), 4)
    assert np.allclose(np.linalg.norm(hp[:, :3], axis=1), 1.0, atol=1e-6)
    margin = lambda q: (q @ hp[:, :3].T + hp[:, 3]).max(1)
    assert (margin(pts) <= 1e-5).all(), "every input point is inside or on its own hull"
    inside = 0.3 * pts[:20] + 0.7 * pts.mean(0)
    assert (margin(inside) < 0).all()
    outside = pts.mean(0) + 10.0 * rng.normal(size=(20, 3)) / 1.0
    far = np.linalg.norm(outside - pts.mean(0), axis=1) > np.linalg.norm(pts - pts.mean(0), axis=1).max()
    assert (margin(outside[far]) > 0).all()


def test_store_signature_tracks_identity_and_version():
    from myslam_b200.field import Signature

    a, b = torch.zeros(4), torch.zeros(4)
    s0 = Signature([a, b, 1.5])
    assert s0 == Signature([a, b, 1.5])
    assert s0 != Signature([a, b, 2.5]) and s0 != Signature([a, 1.5]) and s0 != Signature([a, a, 1.5])
    a.add_(1.0)  # in-place update bumps the autograd version: the mirror must be refreshed
    assert s0 != Signature([a, b, 1.5])
    c = torch.zeros(4)
    assert Signature([c]) != Signature([torch.zeros(4)]), "equal values in another tensor are not the same map"


def test_shard_ranges_tile_the_lattice_for_every_world_size():
    from myslam_b200.dist import shard_range

    for total in (0, 1, 7, 990 * 680 * 490):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (s0, c0), (s1, _) in zip(spans[:-1], spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
