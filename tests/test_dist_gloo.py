"""CPU, world_size 2, gloo: the two exchange points of the ray-sharded mapping (myslam_b200/dist.py) and the
claim they rest on -- with ALL-REDUCED normalisers, the sum of the ranks' gradients is the gradient of the
global-batch loss (checked with the oracle, which is only the checker here)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN_CAM, ROOT, TRUNC, golden_field, load_npz, rel_err


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)


def _unnormalised_terms(O, fld, ro, rd, d, c, z, w):
    """Per-rank SUMS of the five loss terms and their element counts (the kernel's loss_acc / counters)."""
    depth, rgb, sdf, _, _ = O.composite(fld, ro, rd, z)
    m = d > 0
    front, center, tail = O.sdf_band_masks(z[m], d[m], TRUNC)
    resid = z[m] + sdf[m] * TRUNC - d[m][:, None]
    sums = [torch.square(sdf[m][front] - 1.0).sum(), torch.square(resid[center]).sum(), torch.square(resid[tail]).sum(),
            torch.square(d[m] - depth[m]).sum(), torch.square(c - rgb).sum()]
    counts = torch.tensor([ro.shape[0], 0, int(m.sum()), int(front.sum()), int(center.sum()), int(tail.sum()), 0, 0],
                          dtype=torch.int32)
    return sums, counts


def _worker(rank, world, port, out):
    _init(rank, world, port)
    import eslam_oracle as O
    from myslam_b200.dist import MappingExchange, shard_range

    torch.set_num_threads(1)
    fld = golden_field().clone(requires_grad=True)
    d = load_npz("mapping.npz")
    cam = O.Camera(*GOLDEN_CAM)
    c2ws = torch.from_numpy(d["c2ws0"])
    cols, deps = torch.from_numpy(d["gt_colors"]), torch.from_numpy(d["gt_depths"])
    g = torch.Generator().manual_seed(5)
    n_per = 60  # pixels per frame in the GLOBAL batch; each rank takes a contiguous half of every frame's draw
    idx = torch.randint(cam.H * cam.W, (4 * n_per,), generator=g)
    u = torch.rand(4 * n_per, 40, generator=g)

    def rays_for(sel):
        draws = O.ReplayDraws([idx.reshape(4, n_per)[:, sel].reshape(-1)])
        ro, rd, dd, cc, _ = O.sample_rays(0, cam.H, 0, cam.W, len(range(*sel.indices(n_per))), cam.fx, cam.fy, cam.cx,
                                          cam.cy, c2ws, deps, cols, draws)
        keep = O.bbox_keep(ro, rd, dd, fld.bound, False) & (dd > 0)  # depth>0 only: keeps z independent of the field
        uu = u.reshape(4, n_per, 40)[:, sel].reshape(-1, 40)[keep]
        ro, rd, dd, cc = ro[keep], rd[keep], dd[keep], cc[keep]
        z = O.ray_depths(fld, ro, rd, dd, TRUNC, 32, 8, O.ReplayDraws([uu]))
        return ro, rd, dd, cc, z

    w = O.MAP_W
    start, count = shard_range(n_per, rank, world)
    ro, rd, dd, cc, z = rays_for(slice(start, start + count))
    sums, counts = _unnormalised_terms(O, fld, ro, rd, dd, cc, z, w)
    ex = MappingExchange()
    norm = ex.reduce_counters(counts)
    assert int(norm[0]) >= int(counts[0]) and norm.dtype == torch.int32
    loss = (w.fs * sums[0] / norm[3] + w.center * sums[1] / norm[4] + w.tail * sums[2] / norm[5]
            + w.color * sums[4] / (3.0 * norm[0]) + w.depth * sums[3] / norm[2])
    loss.backward()
    flat = torch.cat([t.grad.reshape(-1) for t in fld.leaves()]).float().contiguous()
    ex.reduce_grads(flat)
    if rank == 0:
        # single-process reference: the same global batch through the oracle's own (mean-based) loss
        f2 = golden_field().clone(requires_grad=True)
        fld = f2
        ro, rd, dd, cc, z = rays_for(slice(0, n_per))
        depth, rgb, sdf, _, _ = O.composite(f2, ro, rd, z)
        ref, _ = O.mapping_loss(depth, rgb, sdf, z, dd, cc, TRUNC, w)
        ref.backward()
        flat_ref = torch.cat([t.grad.reshape(-1) for t in f2.leaves()]).float()
        torch.save({"err": rel_err(flat, flat_ref), "n": int(norm[0])}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_mapping_gradients_equal_global_batch(tmp_path):
    out = str(tmp_path / "res.pt")
    port = 29600 + os.getpid() % 200
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["n"] > 100
    assert res["err"] < 1e-5, res


def test_shard_range_partitions_exactly():
    from myslam_b200.dist import shard_range

    for total in (0, 1, 7, 4000, 329_868_000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
