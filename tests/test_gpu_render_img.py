"""GPU: full-image inference (Renderer.render_img) against the reference's output, in the
reference's chunks and in one pass."""
import pytest
import torch

from conftest import GOLDEN_CAM, TRUNC, SimpleEslam, base_cfg, golden_field, load_npz, rel_err, to_device_scene

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _setup():
    import myslam_b200 as M

    fld = golden_field()
    planes, dec = to_device_scene(fld, DEV)
    rnd = M.Renderer(base_cfg(), SimpleEslam(fld.bound.clone(), GOLDEN_CAM, DEV), ray_batch_size=1000)
    d = load_npz("img.npz")
    draws = [torch.from_numpy(d[f"draw.{k}"]) for k in range(int(d["n_draws"]))]
    return M, rnd, planes, dec, d, draws


def test_render_img_chunked_golden():
    M, rnd, planes, dec, d, draws = _setup()
    rnd.strict_rng = True
    rnd.draws = M.ReplayDraws(draws, DEV)
    dep, col = rnd.render_img(planes, dec, torch.from_numpy(d["c2w"]).to(DEV), TRUNC, DEV,
                              gt_depth=torch.from_numpy(d["gt_depth"]).to(DEV))
    assert dep.dtype == torch.float64 and tuple(dep.shape) == GOLDEN_CAM[:2] and tuple(col.shape) == GOLDEN_CAM[:2] + (3,)
    assert rel_err(dep, d["depth"]) < 1e-4 and rel_err(col, d["color"]) < 1e-4


def test_render_img_single_pass_equals_chunks():
    """One sampling + one render launch over all H*W rays gives the chunked result when fed the same uniforms
    (the per-chunk draws concatenated in ray order)."""
    M, rnd, planes, dec, d, draws = _setup()
    gt = torch.from_numpy(d["gt_depth"]).reshape(-1)
    us, ucs, ufs, k = [], [], [], 0
    for i in range(0, gt.numel(), 1000):
        r0 = int((gt[i:i + 1000] <= 0).sum())
        us.append(draws[k])
        k += 1
        if r0 > 0:
            ucs.append(draws[k])
            ufs.append(draws[k + 1])
            k += 2
    assert k == len(draws) and len(ucs) > 0
    rnd.strict_rng = False
    rnd.draws = M.ReplayDraws([torch.cat(us, 0), torch.cat(ucs, 0), torch.cat(ufs, 0)], DEV)
    dep, col = rnd.render_img(planes, dec, torch.from_numpy(d["c2w"]).to(DEV), TRUNC, DEV,
                              gt_depth=torch.from_numpy(d["gt_depth"]).to(DEV))
    assert rel_err(dep, d["depth"]) < 1e-4 and rel_err(col, d["color"]) < 1e-4
