import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the shared library is a build artefact (git-ignored): build it in-tree if this checkout does not have it yet
    # (nvcc cross-compiles sm_100a without a GPU, ~35 s); the tests themselves never fall back to anything else
    from myslam_b200.build import OUT, build_library

    if not os.path.exists(OUT):
        build_library()


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


PLANE_NAMES = ("xy", "xz", "yz", "c_xy", "c_xz", "c_yz")


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name))


def golden_field():
    """The fixture scene as an oracle Field (CPU)."""
    import eslam_oracle as O

    d = load_npz("field.npz")
    planes = tuple([torch.from_numpy(d[f"plane.{n}.{s}"]) for s in range(2)] for n in PLANE_NAMES)
    dec = {k: torch.from_numpy(d[f"dec.{k}"]) for k in O.DECODER_KEYS}
    return O.Field(planes, dec, torch.from_numpy(d["beta"]), torch.from_numpy(d["bound"]))


def recorded_draws(d):
    return [torch.from_numpy(d[f"draw.{k}"]) for k in range(int(d["n_draws"]))]


def rel_err(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.numel() == 0:
        return 0.0
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def to_device_scene(fld, device="cuda", requires_grad=False):
    """Oracle Field -> (all_planes on device, myslam_b200.Decoders on device)."""
    from myslam_b200 import Decoders

    all_planes = tuple([p.clone().to(device).requires_grad_(requires_grad) for p in g] for g in fld.planes)
    dec = Decoders(c_dim=32, truncation=0.06, learnable_beta=True)
    dec.load_state_dict({**{k: v.clone() for k, v in fld.dec.items()}, "beta": fld.beta.clone()})
    dec = dec.to(device)
    dec.bound = fld.bound.clone()
    for p in dec.parameters():
        p.requires_grad_(requires_grad)
    return all_planes, dec


class SimpleEslam:
    """What Renderer.__init__ reads from ESLAM (Renderer.py:34-44)."""

    def __init__(self, bound, cam, device):
        self.bound, self.device = bound, device
        self.H, self.W, self.fx, self.fy, self.cx, self.cy = cam


GOLDEN_CAM = (48, 64, 50.0, 50.0, 31.5, 23.5)
TRUNC = 0.06


def base_cfg(n_strat=32, n_imp=8, trunc=TRUNC):
    return {
        "scale": 1,
        "rendering": {"perturb": True, "n_stratified": n_strat, "n_importance": n_imp, "learnable_beta": True},
        "model": {"c_dim": 32, "truncation": trunc},
        "tracking": {"ignore_edge_W": 6, "ignore_edge_H": 5, "lr_T": 0.002, "lr_R": 0.001, "pixels": 200, "iters": 3,
                     "w_sdf_fs": 10, "w_sdf_center": 200, "w_sdf_tail": 50, "w_depth": 1, "w_color": 5},
        "mapping": {"pixels": 400, "iters": 2, "mapping_window_size": 20, "keyframe_selection_method": "global",
                    "joint_opt": True, "joint_opt_cam_lr": 0.001, "w_sdf_fs": 5, "w_sdf_center": 200,
                    "w_sdf_tail": 10, "w_depth": 0.1, "w_color": 5,
                    "lr": {"decoders_lr": 0.001, "planes_lr": 0.005, "c_planes_lr": 0.005}},
    }


def arena_index(k):
    """Index in the oracle's leaf order ([xy_c, xy_f, xz_c, xz_f, ...], group-major) -> arena plane slot."""
    from myslam_b200.field import arena_slot

    return arena_slot(k // 2, k % 2)


class PaddedOracleDraws:
    """Oracle-side view of the DEFAULT (non-strict) draw convention of myslam_b200.hotpath: the uniforms are drawn
    as fixed [N,S] / [N,n_strat] / [N,n_imp] blocks and a kept ray uses the row of its ordinal among the rays of its
    kind, i.e. the oracle's [R1,S] / [R0,n_strat] / [R0,n_imp] requests are the first rows of the padded blocks."""

    def __init__(self, idx, blocks):
        self.idx, self.blocks, self.pos = idx, list(blocks), 0

    def randint(self, high, n):
        assert self.idx.numel() == n
        return self.idx

    def rand(self, rows, cols):
        t = self.blocks[self.pos]
        self.pos += 1
        assert t.shape[1] == cols and rows <= t.shape[0], (tuple(t.shape), rows, cols)
        return t[:rows]


class PaddedDeviceDraws:
    """The same padded blocks handed to the CUDA path (TorchDraws' interface incl. rand_many)."""

    def __init__(self, idx, blocks, device):
        self.idx = idx.to(device).contiguous()
        self.blocks = [b.to(device).float().contiguous() for b in blocks]
        self.pos = 0

    def randint(self, high, n):
        assert self.idx.numel() == n
        self.pos = 0
        return self.idx

    def rand(self, rows, cols):
        t = self.blocks[self.pos]
        self.pos += 1
        assert tuple(t.shape) == (rows, cols), (tuple(t.shape), rows, cols)
        return t

    def rand_many(self, shapes):
        return [self.rand(r, c) for r, c in shapes]
