"""EXPERIMENTAL entry points (include/eslam_b200.h, "pre-activated planes"): not part of the
product path yet; held to 1e-5 of the product's render forward on the same rays."""
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
