"""CPU: host-side mirror logic (pose maths, decoder packing table, state_dict surface)."""
import numpy as np
import torch

import eslam_oracle as O
from conftest import load_npz, rel_err


def test_pose_mirror_matches_oracle_and_golden():
    from myslam_b200 import common

    d = load_npz("pose.npz")
    poses = torch.from_numpy(d["poses"])
    assert rel_err(common.cam_pose_to_matrix(poses), d["mats"]) < 1e-6
    M = torch.from_numpy(d["mats"])
    assert rel_err(common.matrix_to_cam_pose(M), d["back"]) < 1e-6
    assert torch.equal(common.matrix_to_cam_pose(M), O.matrix_to_cam_pose(M))


def test_decoders_surface_matches_reference_keys():
    from myslam_b200 import Decoders

    dec = Decoders(c_dim=32, truncation=0.06, learnable_beta=True)
    keys = set(dec.state_dict().keys())
    expect = {f"{p}linears.{i}.{w}" for p in ("", "c_") for i in (0, 1) for w in ("weight", "bias")}
    expect |= {f"{p}output_linear.{w}" for p in ("", "c_") for w in ("weight", "bias")} | {"beta"}
    assert keys == expect
    assert sum(p.numel() for p in dec.parameters()) == 2693
    dec2 = Decoders(learnable_beta=False)
    assert dec2.beta == 10 and "beta" not in dec2.state_dict()
    import copy, pickle
    copy.deepcopy(dec)
    pickle.loads(pickle.dumps(dec))
    dec.share_memory()


def test_decoder_packing_table_is_consistent():
    from myslam_b200.field import DEC_LAYOUT
    from myslam_b200._lib import DEC_FLOATS

    shapes = {"linears.0.weight": 1024, "linears.0.bias": 16, "linears.1.weight": 256, "linears.1.bias": 16,
              "output_linear.weight": 16, "output_linear.bias": 1, "c_linears.0.weight": 1024, "c_linears.0.bias": 16,
              "c_linears.1.weight": 256, "c_linears.1.bias": 16, "c_output_linear.weight": 48,
              "c_output_linear.bias": 3, "beta": 1}
    end = 0
    for key, off, n in DEC_LAYOUT:
        assert shapes[key] == n and off % 4 == 0 and off >= end
        end = off + n
    assert end <= DEC_FLOATS


def test_grid_axes_match_oracle():
    from myslam_b200.mesher import grid_axes

    b = [[-1.9, 7.9], [-2.2, 4.5], [-2.5, 2.3]]
    a1, a2 = grid_axes(b, 0.01), O.grid_axes(b, 0.01)
    assert [len(a) for a in a1] == [990, 680, 490]
    for x, y in zip(a1, a2):
        assert np.array_equal(x, y)


def test_plane_shapes_room0_float32_truncation():
    # 6.7199993/0.06 -> 111, not 112
    b = O.rounded_bound([[-1.9, 7.9], [-2.2, 4.5], [-2.5, 2.3]], 0.24)
    assert O.plane_shapes(b, 0.06) == ((111, 164), (84, 164), (84, 111))
    assert O.plane_shapes(b, 0.24) == ((27, 41), (21, 41), (21, 27))
    assert O.plane_shapes(b, 0.03) == ((223, 328), (168, 328), (168, 223))


def test_install_patches_the_reference_package():
    """`myslam_b200.install("src")` on the UNMODIFIED reference (only where /root/reference exists: the build
    container): the four redirections of INTEGRATION.md land on the reference's own classes and keep their
    signatures."""
    import inspect
    import os
    import sys

    import pytest

    ref = os.environ.get("ESLAM_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "src")):
        pytest.skip("the reference tree is not available here")
    from conftest import ROOT

    standins = os.path.join(ROOT, "oracle", "standins")
    added = [p for p in (standins, ref) if p not in sys.path]
    for p in added:
        sys.path.insert(0, p)
    try:
        import src.Mapper as RM
        import src.Tracker as RT
        import src.networks.decoders as RD
        import src.utils.Renderer as RR

        ref_sigs = {"ot": inspect.signature(RT.Tracker.optimize_tracking), "om": inspect.signature(RM.Mapper.optimize_mapping),
                    "ks": inspect.signature(RM.Mapper.keyframe_selection_overlap),
                    "rb": inspect.signature(RR.Renderer.render_batch_ray), "ri": inspect.signature(RR.Renderer.render_img),
                    "fw": inspect.signature(RD.Decoders.forward), "di": inspect.signature(RD.Decoders.__init__)}
        import myslam_b200 as M

        M.install("src")
        assert RD.Decoders is M.Decoders and RR.Renderer is M.Renderer
        assert RT.Tracker.optimize_tracking is M.tracker.optimize_tracking
        assert RM.Mapper.optimize_mapping is M.mapper.optimize_mapping
        new_sigs = {"ot": inspect.signature(RT.Tracker.optimize_tracking), "om": inspect.signature(RM.Mapper.optimize_mapping),
                    "ks": inspect.signature(RM.Mapper.keyframe_selection_overlap),
                    "rb": inspect.signature(M.Renderer.render_batch_ray), "ri": inspect.signature(M.Renderer.render_img),
                    "fw": inspect.signature(M.Decoders.forward), "di": inspect.signature(M.Decoders.__init__)}
        for k in ref_sigs:
            assert list(ref_sigs[k].parameters) == list(new_sigs[k].parameters), (k, ref_sigs[k], new_sigs[k])
    finally:
        for p in added:
            sys.path.remove(p)
        for name in [n for n in sys.modules if n == "src" or n.startswith("src.")]:
            del sys.modules[name]
