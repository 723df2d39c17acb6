#!/usr/bin/env python
"""bench.py -- the reference's headline metric on the B200-native hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Metric (BASELINE.json): mapping rays*iters/s (and tracking frames/s) on a synthetic Replica-room0-shaped
1200x680 stream, ESLAM.yaml defaults.  One "step" = one `Mapper.optimize_mapping` call: 15 iterations x 4000
rays over a 20-keyframe window (per GPU; with N GPUs the window's ray batch is N x 4000, sharded, weak scaling).

  value      device-resident: frames, window stack and parameters already in HBM, the per-call loop only
  e2e        the same metric through the reference-facing `optimize_mapping` drop-in, with the current frame
             coming from pinned HOST memory every step (H2D inside the timed region) and the pose read back
  roofline   the dominant kernel (fused loss+backward) timed with CUDA events, algorithmic bytes / time
  cpu_baseline   the oracle port of the reference's PyTorch path on the host cores (bounded sample)
  tracking   frames/s of the per-frame camera loop (8 iterations x 2000 rays), device-resident and e2e

`--impl reference` times the reference's algorithm (oracle port; the reference is Python and cannot travel to
the GPU box) on the host cores for the same metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_RAY_ITER = 40 * 6144 * 2  # S x (12 taps x 4 corners x 128 B) x (gather + scatter)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", d
    return 6650.0, "fallback (B200_PROFILING.md)", {}


def l2_ceilings():
    """Measured L2-resident ceilings of the hot path's access shape (tools/microbench/l2_gather_red.cu, run here):
    random 64-byte / 128-byte line gathers (LDG.128) and red.global.add.v4.f32, alone and interleaved, plus the
    bulk-async forms (TMA gather4, bulk reduce).  None if the binary is missing and cannot be built."""
    exe = os.path.join(ROOT, "tools", "microbench", "l2_gather_red")
    if not os.path.exists(exe):
        try:
            subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-o", exe,
                            exe + ".cu"], check=True, capture_output=True, timeout=300)
        except Exception:
            return None
    try:
        out = subprocess.run([exe, "--json"], capture_output=True, text=True, timeout=120).stdout
        for line in out.splitlines():
            if line.startswith("{"):
                return json.loads(line)
    except Exception:
        return None
    return None


def ncu_traffic(kernel_key):
    """DRAM bytes per launch of a kernel from this round's committed `ncu --set full` capture (profiles/)."""
    prof = os.path.join(ROOT, "profiles", "r02_ncu_summary.json")
    try:
        d = json.load(open(prof))
        return d[kernel_key]["dram_bytes_per_launch"], "profiles/r02_ncu_summary.json (ncu --set full, cold cache)"
    except Exception:
        return None, None


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 3 + k and r[3 + k] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def build_inputs(spec, device, n_keyframes, seed):
    from myslam_b200 import synthetic as S

    gen = torch.Generator().manual_seed(seed)
    poses = S.trajectory(n_keyframes, spec["room"], step_deg=2.0)
    cam = (spec["H"], spec["W"], spec["fx"], spec["fy"], spec["cx"], spec["cy"])
    cols, deps = [], []
    for k in range(n_keyframes):
        c, d = S.render_box_room(poses[k], *cam, spec["room"], device, hole_frac=0.03, generator=gen)
        cols.append(c)
        deps.append(d)
    return poses, torch.stack(cols, 0).contiguous(), torch.stack(deps, 0).contiguous()


def time_region(fn, steps, warmup, dist_on):
    """W warm-up + K timed steps, barrier + synchronize on both sides, CUDA events, max over ranks (ms)."""
    import torch.distributed as dist

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if dist_on:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item()


class FramePrefetcher:
    """Host -> device staging of the step's input frame from PINNED host memory on a copy stream, one step ahead,
    as a frame loader does for a live stream: every step's copy is issued and waited for inside the timed region,
    but it overlaps the previous step's kernels instead of preceding its own."""

    def __init__(self, host_tensors, dev):
        self.host, self.dev = host_tensors, dev
        self.stream = torch.cuda.Stream(device=dev)
        self.slots = [[torch.empty_like(h, device=dev) for h in host_tensors] for _ in range(2)]
        self.events = [torch.cuda.Event(), torch.cuda.Event()]
        self.consumed = [torch.cuda.Event(), torch.cuda.Event()]
        self.k = 0
        self._issue(0)

    def _issue(self, slot):
        self.stream.wait_event(self.consumed[slot])  # no-op until the slot has been used once
        with torch.cuda.stream(self.stream):
            for h, d in zip(self.host, self.slots[slot]):
                d.copy_(h, non_blocking=True)
            self.events[slot].record(self.stream)

    def next(self):
        """Device tensors of this step's frame (the copy was issued during the previous step); issues the next."""
        slot = self.k & 1
        torch.cuda.current_stream().wait_event(self.events[slot])
        self.k += 1
        self._issue(self.k & 1)
        return self.slots[slot]

    def done(self, slot_tensors):
        """Call after the step's kernels are enqueued: the slot may be overwritten once they have run."""
        slot = 0 if slot_tensors is self.slots[0] else 1
        self.consumed[slot].record()


class OracleRun:
    """The reference's algorithm (oracle port) set up once on `device`; run_mapping / run_tracking time
    `iters` iterations and return seconds per iteration."""

    def __init__(self, spec, n_frames, device="cpu", seed=0):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import eslam_oracle as O

        self.O, self.spec, self.device = O, spec, device
        gen = torch.Generator().manual_seed(seed)
        bound = O.rounded_bound(spec["bound"], spec["bound_dividable"])
        fld = O.make_field(bound, spec["planes_res"], spec["c_planes_res"], generator=gen)
        if device != "cpu":
            fld = O.Field(tuple([p.to(device) for p in g] for g in fld.planes),
                          {k: v.to(device) for k, v in fld.dec.items()}, fld.beta.to(device), fld.bound.to(device))
        self.fld = fld
        poses, self.cols, self.deps = build_inputs(spec, device, n_frames, seed)
        self.poses = poses.to(device)
        self.cam = O.Camera(spec["H"], spec["W"], spec["fx"], spec["fy"], spec["cx"], spec["cy"])
        self.rc = O.RenderCfg(spec["n_stratified"], spec["n_importance"], spec["truncation"])
        self.draws = O.LiveDraws(None, device)

    def _timed(self, fn):
        if self.device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        if self.device != "cpu":
            torch.cuda.synchronize()
        self.draws.log.clear()
        return time.perf_counter() - t0

    def run_mapping(self, iters):
        O, m = self.O, self.spec["mapping"]
        return self._timed(lambda: O.map_window(self.fld, self.cam, self.rc, O.MAP_W, self.poses, self.cols, self.deps,
                                                m["pixels"], iters, 1e-3, 5e-3, 5e-3, True, 1e-3, self.draws)) / iters

    def run_tracking(self, iters):
        O, t = self.O, self.spec["tracking"]
        pose0 = O.matrix_to_cam_pose(self.poses[:1])
        return self._timed(lambda: O.track_frame(self.fld, self.cam, self.rc, O.TRACK_W, pose0, self.cols[:1],
                                                 self.deps[:1], t["pixels"], t["ignore_edge_H"], t["ignore_edge_W"],
                                                 iters, t["lr_T"], t["lr_R"], self.draws)) / iters


class ReferenceRun:
    """The UNMODIFIED reference (MohammadJohari/myslam, Python) on the host cores, when its sources are reachable
    (ESLAM_REFERENCE or /root/reference: the build container; the GPU box does not have them and falls back to the
    oracle port).  Its Mapper / Tracker are instantiated with object.__new__ and given exactly the attributes
    optimize_mapping / optimize_tracking read (as tests/golden/make_golden.py does); import-only stand-ins for
    colorama / matplotlib / trimesh / open3d / skimage and the restated pytorch3d.transforms come from oracle/standins."""

    @staticmethod
    def path():
        p = os.environ.get("ESLAM_REFERENCE", "/root/reference")
        return p if os.path.exists(os.path.join(p, "src", "Mapper.py")) else None

    def __init__(self, spec, n_frames, seed=0):
        import types

        for q in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "oracle", "standins"), self.path()):
            if q not in sys.path:
                sys.path.insert(0, q)
        import eslam_oracle as O
        from src.Mapper import Mapper
        from src.networks.decoders import Decoders
        from src.utils.Renderer import Renderer

        self.O, self.spec = O, spec
        gen = torch.Generator().manual_seed(seed)
        bound = O.rounded_bound(spec["bound"], spec["bound_dividable"])
        fld = O.make_field(bound, spec["planes_res"], spec["c_planes_res"], generator=gen)
        poses, cols, deps = build_inputs(spec, "cpu", n_frames, seed)
        dec = Decoders(c_dim=32, truncation=spec["truncation"], learnable_beta=True)
        dec.load_state_dict({**{k: v.clone() for k, v in fld.dec.items()}, "beta": fld.beta.clone()})
        dec.bound = bound.clone()
        cam = dict(H=spec["H"], W=spec["W"], fx=spec["fx"], fy=spec["fy"], cx=spec["cx"], cy=spec["cy"])
        es = types.SimpleNamespace(bound=bound.clone(), device="cpu", **cam)
        rnd = Renderer({"rendering": {"perturb": True, "n_stratified": spec["n_stratified"],
                                      "n_importance": spec["n_importance"]}, "scale": 1}, es)
        m = spec["mapping"]
        mp = object.__new__(Mapper)
        (mp.planes_xy, mp.planes_xz, mp.planes_yz, mp.c_planes_xy, mp.c_planes_xz, mp.c_planes_yz) = \
            tuple([p.clone() for p in g] for g in fld.planes)
        mp.device, mp.bound, mp.renderer, mp.decoders, mp.truncation = "cpu", bound.clone(), rnd, dec, spec["truncation"]
        for k, v in cam.items():
            setattr(mp, k, v)
        w = O.MAP_W
        mp.w_sdf_fs, mp.w_sdf_center, mp.w_sdf_tail, mp.w_depth, mp.w_color = w.fs, w.center, w.tail, w.depth, w.color
        mp.cfg = {"mapping": {"lr": dict(m["lr"])}}
        mp.keyframe_selection_method, mp.mapping_window_size, mp.mapping_pixels = "global", n_frames, m["pixels"]
        mp.joint_opt, mp.joint_opt_cam_lr, mp.no_vis_on_first_frame = True, m["joint_opt_cam_lr"], True
        mp.visualizer = types.SimpleNamespace(save_imgs=lambda *a, **k: None)
        self.kf = [{"gt_c2w": poses[k], "idx": torch.tensor(4 * k), "color": cols[k], "depth": deps[k],
                    "est_c2w": poses[k].clone()} for k in range(n_frames - 1)]
        mp.keyframe_dict = self.kf
        self.mp, self.cur = mp, (cols[-1], deps[-1], poses[-1])
        self.kf_list = list(range(0, 4 * (n_frames - 1), 4))

    def run_mapping(self, iters):
        """seconds per iteration of one Mapper.optimize_mapping call of `iters` iterations (Mapper.py:211-364)."""
        col, dep, c2w = self.cur
        t0 = time.perf_counter()
        self.mp.optimize_mapping(iters, 1.0, torch.tensor(4 * len(self.kf) + 4), col, dep, c2w.clone(), self.kf,
                                 self.kf_list, c2w.clone())
        return (time.perf_counter() - t0) / iters


def extras(scene, rnd, spec, dev, rank, world, dist_on, poses, deps, hbm_peak):
    """Secondary workloads of BASELINE.json (configs 3 and 5) and the full-frame render, each a few launches."""
    import torch.distributed as dist
    import myslam_b200 as M
    from myslam_b200 import synthetic as S
    from myslam_b200.dist import shard_range
    from myslam_b200.mesher import grid_axes, query_grid_sdf

    out = {}
    # ---- config 5: marching-cubes SDF query at 1 cm over the room0 bound, flat lattice ranges sharded over ranks
    axes = grid_axes(spec["bound"], 0.01)
    total = len(axes[0]) * len(axes[1]) * len(axes[2])
    start, count = shard_range(total, rank, world)
    buf = torch.empty(count, dtype=torch.float32, device=dev)
    q = lambda **kw: (lambda: query_grid_sdf(scene.all_planes, scene.decoders, axes, scene.bound, start=start,
                                             count=count, out=buf, **kw))
    ms = time_region(q(), 3, 1, dist_on) / 3  # the default form (factored)
    os.environ["ESLAM_B200_GRID_ROWS"] = "0"
    ms_voxel = time_region(q(), 3, 1, dist_on) / 3  # the factored form without the tensor-core rows kernel
    del os.environ["ESLAM_B200_GRID_ROWS"]
    forms = {"factored_ms": ms, "factored_per_voxel_ms": ms_voxel,
             "separable_ms": time_region(q(separable=True), 3, 1, dist_on) / 3,
             "direct_ms": time_region(q(separable=False), 3, 1, dist_on) / 3}
    out["mesh_query"] = {"value": total / (ms * 1e-3), "unit": "points/s", "ms": ms, "points": total,
                         "lattice": [len(a) for a in axes], "n_gpus": world, "scaling": "strong", "forms": forms,
                         "bytes_per_point": {"reference_gather": 3076, "separable": 772, "factored": 196},
                         "factored_GBps_per_gpu": count * 196 / (ms * 1e-3) / 1e9,  # served by L2 (faces: 95 MB)
                         "hbm_write_GBps_per_gpu": count * 4 / (ms * 1e-3) / 1e9,
                         "what": "Mesher.get_grid_uniform + eval_points (Mesher.py:130-186), SDF head only, coordinates "
                                 "generated in-kernel.  direct: every voxel gathers its 24 corners; separable: the "
                                 "planes are resampled once on the lattice's faces (768 B per point, bit-identical); "
                                 "factored (default): the first decoder layer is applied on the faces too (equal to "
                                 "1e-5); whole lattice rows run the 16->16 layer as mma.sync 3xTF32 with the xz face "
                                 "values in registers (eslam_grid_sdf_rows), ragged range ends the per-voxel kernel "
                                 "(192 B and 272 FMA per point: factored_per_voxel_ms).  Face resampling is inside "
                                 "every timed call."}
    del buf
    if rank != 0:
        return out
    # ---- marching cubes over the same 1 cm lattice on the device (Mesher.py:219-247).  The synthetic map is untrained
    # (its sdf has no surface), so the volume here is the analytic signed distance of the box room the frames were
    # rendered from: a surface of the size a trained room0 map has.
    try:
        from myslam_b200.mesher import marching_cubes

        xs, ys, zs = (torch.from_numpy(a).float().to(dev) for a in axes)
        room = torch.tensor(spec["room"], dtype=torch.float32, device=dev)
        dx = torch.minimum(xs - room[0, 0], room[0, 1] - xs)[None, :, None]
        dy = torch.minimum(ys - room[1, 0], room[1, 1] - ys)[:, None, None]
        dz = torch.minimum(zs - room[2, 0], room[2, 1] - zs)[None, None, :]
        vol = torch.minimum(torch.minimum(dx.expand(len(ys), len(xs), len(zs)), dy.expand(len(ys), len(xs), len(zs))),
                            dz.expand(len(ys), len(xs), len(zs))).contiguous().reshape(-1)  # (iy*nx + ix)*nz + iz
        del dx, dy, dz
        res_mc = {}
        ms_soup = time_region(lambda: res_mc.__setitem__("m", marching_cubes(vol, axes, 0.0, weld=False)), 3, 1, False) / 3
        ms_weld = time_region(lambda: res_mc.__setitem__("m", marching_cubes(vol, axes, 0.0, weld=True)), 2, 1, False) / 2
        v, f = res_mc["m"]
        out["mesh_extract"] = {"ms_marching_cubes": ms_soup, "ms_with_welding": ms_weld, "triangles": int(f.shape[0]),
                               "vertices": int(v.shape[0]), "cells": (len(xs) - 1) * (len(ys) - 1) * (len(zs) - 1),
                               "GBps_volume_read": vol.numel() * 4 / (ms_soup * 1e-3) / 1e9,
                               "what": "device marching cubes over the 990x680x490 lattice (two passes over the "
                                       "1.32 GB volume, which never leaves HBM; the reference copies it to the host "
                                       "for skimage): count + emit kernels; welding = torch.unique on the vertices' "
                                       "lattice-edge keys"}
        del vol, v, f, res_mc
    except Exception as exc:  # noqa: BLE001
        out["mesh_extract"] = {"error": repr(exc)}
    # ---- full-frame inference (Renderer.render_img, 1200x680 = 816 k rays, one pass)
    fn = lambda: rnd.render_img(scene.all_planes, scene.decoders, poses[0], spec["truncation"], dev, gt_depth=deps[0])
    ms = time_region(fn, 3, 1, False) / 3
    out["render_img"] = {"ms": ms, "rays": spec["H"] * spec["W"], "rays_per_s": spec["H"] * spec["W"] / (ms * 1e-3),
                         "what": "Renderer.render_img (Renderer.py:155-204) in one sampling + one render launch "
                                 "(the reference: 82 chunks of 10 k rays)"}
    # ---- config 3: ScanNet scene0000 shape (620x460 after crop, S = 48 + 8, 70.5 MB of planes)
    sp = S.SCANNET_0000
    sc3 = S.make_scene(sp, dev, seed=0)
    cfg3 = S.run_cfg(sp)

    class E:
        pass

    e = E()
    e.bound, e.device = sc3.bound, dev
    e.H, e.W, e.fx, e.fy, e.cx, e.cy = sc3.cam
    rnd3 = M.Renderer(cfg3, e)
    m3, t3 = sp["mapping"], sp["tracking"]
    nf = m3["mapping_window_size"]
    p3, c3, d3 = build_inputs(sp, dev, nf, seed=2)
    p3 = p3.to(dev)
    from myslam_b200.common import matrix_to_cam_pose
    from myslam_b200.decoders import synced_store
    from myslam_b200.mapper import _mapper_state, map_window

    mp3 = M.MapperStep(cfg3, rnd3, sc3.decoders, sc3.all_planes, sc3.bound, sc3.cam, dev)
    st3 = _mapper_state(mp3, m3["pixels"], nf)
    store3 = synced_store(sc3.all_planes, sc3.decoders, sc3.bound)
    lr = m3["lr"]
    iters = 15
    fn = lambda: map_window(store3, st3["ws"], st3["sc"], p3, c3, d3, m3["pixels"], iters, lr["decoders_lr"],
                            lr["planes_lr"], lr["c_planes_lr"], True, m3["joint_opt_cam_lr"])
    ms = time_region(fn, 3, 1, False) / 3
    trk3 = M.TrackerStep(cfg3, rnd3, sc3.decoders, sc3.all_planes, sc3.bound, sc3.cam, dev)
    pose0 = matrix_to_cam_pose(p3[:1])
    col1, dep1 = c3[:1].contiguous(), d3[:1].contiguous()
    ms_t = time_region(lambda: trk3.track_frame(pose0, col1, dep1), 3, 1, False) / 3
    S56 = sp["n_stratified"] + sp["n_importance"]
    out["scannet_scene0000_shape"] = {
        "mapping_rays_iters_per_s": iters * (m3["pixels"] // nf) * nf / (ms * 1e-3), "mapping_ms_per_iter": ms / iters,
        "tracking_frames_per_s": 1.0 / (ms_t * 1e-3), "tracking_iters_per_frame": t3["iters"], "samples_per_ray": S56,
        "plane_MB": store3.n_floats * 4 / 1e6,
        "mapping_frac_of_hbm_peak": iters * (m3["pixels"] // nf) * nf * S56 * 6144 * 2 / (ms * 1e-3) / 1e9 / hbm_peak}
    return out


def replicas_identical(t, world):
    """True iff the bits of tensor `t` are the same on every rank (checksum of the int32 view, all-gathered)."""
    import torch.distributed as dist

    h = t.view(torch.int32).to(torch.int64).sum().reshape(1)
    hs = [torch.zeros_like(h) for _ in range(world)]
    dist.all_gather(hs, h)
    return all(int(x) == int(hs[0]) for x in hs)


def workload_config(world, pix, n_frames, exchange_name):
    return {"workload": "replica-room0-shaped 1200x680, 20-keyframe window, 4000 rays/iter per GPU, "
                        "15 iterations per optimize_mapping call, ESLAM.yaml defaults, joint pose optimisation",
            "rays_per_iter_total": world * pix * n_frames, "exchange": exchange_name,
            "l2": "inputs larger than L2 (471 MB of window frames + 109 MB parameter/optimiser arenas); no flush",
            "launch": ("one CUDA-graph replay per call (reset + 15 pipelined iterations over three streams + pose "
                       "conversion; captured on the second call of a shape)" if world == 1 and
                       os.environ.get("ESLAM_B200_GRAPH", "1") == "1" else "kernel by kernel over three streams"),
            "seed": 0}


def run_reference(args, spec):
    """--impl reference: the reference's CPU implementation of the path (oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m = spec["mapping"]
    n_frames = m["mapping_window_size"]
    kind = "port"
    run = OracleRun(spec, n_frames)
    if ReferenceRun.path():
        try:
            ref = ReferenceRun(spec, n_frames)
            ref.run_tracking = run.run_tracking  # the tracking extra stays on the port
            run, kind = ref, "reference"
        except Exception as exc:  # noqa: BLE001
            print(f"reference sources found but not usable ({exc!r}); timing the oracle port", file=sys.stderr)
    iters = m["iters"]
    for _ in range(min(max(args.warmup, 1), 1)):
        run.run_mapping(2)
    # one step = one optimize_mapping-shaped call of 15 iterations, as on the GPU arm (fresh Adam per call)
    per_iter = [run.run_mapping(iters) for _ in range(max(args.steps, 1))]
    s_iter = sum(per_iter) / len(per_iter)
    value = m["pixels"] / s_iter
    run.run_tracking(1)
    trk = run.run_tracking(4)
    line = {
        "impl": "reference", "metric": "mapping rays*iters/s (Replica room0 shape)", "value": value,
        "unit": "rays*iters/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * s_iter * iters, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(1, m["pixels"] // n_frames, n_frames, "none (host cores)"),
        "cpu_baseline": {"value": value, "unit": "rays*iters/s", "cores": cores, "kind": kind,
                         "sample": f"{len(per_iter)} optimize_mapping-shaped calls of {iters} iterations x 4000 rays "
                                   + ("(the unmodified reference's Mapper.optimize_mapping" if kind == "reference"
                                      else "(oracle port of the reference's PyTorch path") +
                                   f", torch CPU, {cores} threads)"},
        "e2e": {"value": value, "unit": "rays*iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "tracking": {"value": 1.0 / (trk * spec["tracking"]["iters"]), "unit": "frames/s",
                     "ms_per_iter": 1e3 * trk},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-baselines", action="store_true", help="skip the CPU / torch-GPU baseline legs")
    ap.add_argument("--profile-only", action="store_true", help="few iterations, no baselines (for ncu)")
    args = ap.parse_args()

    from myslam_b200 import synthetic as S

    spec = S.REPLICA_ROOM0
    if args.impl == "reference":
        args.steps = min(args.steps, 4)  # bounded: ~10-30 s of host work per 15-iteration step
        run_reference(args, spec)
        return

    import torch.distributed as dist
    import myslam_b200 as M
    from myslam_b200 import _lib
    from myslam_b200.decoders import synced_store
    from myslam_b200.dist import MappingExchange, PeerExchange
    from myslam_b200.hotpath import mapping_iteration
    from myslam_b200.mapper import _mapper_state, map_window
    from myslam_b200.common import matrix_to_cam_pose

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist_on = world > 1
    assert torch.cuda.is_available(), "bench.py --impl b200 needs a GPU: there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if dist_on:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    if args.profile_only:
        args.no_baselines = True
    if os.environ.get("ESLAM_B200_DEBUG"):
        _lib.load().eslam_set_debug(int(os.environ["ESLAM_B200_DEBUG"]))
    m, t = spec["mapping"], spec["tracking"]
    n_frames = m["mapping_window_size"]

    scene = S.make_scene(spec, dev, seed=0)  # identical parameters on every rank
    cfg = S.run_cfg(spec)

    class E:
        pass

    e = E()
    e.bound, e.device = scene.bound, dev
    e.H, e.W, e.fx, e.fy, e.cx, e.cy = scene.cam
    rnd = M.Renderer(cfg, e)
    poses, cols, deps = build_inputs(spec, dev, n_frames, seed=1)
    poses = poses.to(dev)
    torch.manual_seed(1234 + rank)  # every rank draws its own pixels: the union is the N x 4000 batch

    # ------------------------------------------------------------------ mapping, device-resident
    mp = M.MapperStep(cfg, rnd, scene.decoders, scene.all_planes, scene.bound, scene.cam, dev)
    st = _mapper_state(mp, m["pixels"], n_frames)
    store = synced_store(scene.all_planes, scene.decoders, scene.bound)
    ws, sc = st["ws"], st["sc"]
    from myslam_b200.hotpath import FrameTable
    # the window as optimize_mapping stages it: a table of per-frame pointers, the frames read where they live
    window = FrameTable([cols[k] for k in range(n_frames)], [deps[k] for k in range(n_frames)], sc.cam, dev)
    ex, exchange_name = None, "none (1 GPU)"
    if dist_on:
        if os.environ.get("ESLAM_B200_EXCHANGE", "peer") == "nccl":
            ex, exchange_name = MappingExchange(), "NCCL all-reduce of the gradient images + replicated optimiser step"
        else:
            # symmetric (peer-mapped) memory needs P2P between all GPUs of the job: agree on it, and measure the NCCL
            # path (saying so) rather than nothing if the box does not offer it
            err = None
            try:
                ex = PeerExchange(store, ws)
            except Exception as exc:  # noqa: BLE001
                err = repr(exc)
            ok = torch.tensor([0 if err else 1], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 1:
                exchange_name = ("own kernels over symmetric NVLink peer memory, no NCCL call per iteration: push "
                                 "reduce-scatter of the 16-channel gradient images + plane Adam on the owned tiles + "
                                 "all-gather of the updated texels (" +
                                 ("multimem.st" if ex.multimem else "P2P stores") + ") + zero_grad, replicated decoder "
                                 "step on the published decoder gradients, normalisers summed over peer loads; "
                                 "limiter: the exchange runs after the backward with nothing overlapped")
            else:
                ex = MappingExchange()
                exchange_name = ("NCCL all-reduce of the gradient images + replicated optimiser step (symmetric memory "
                                 f"unavailable: {err})")
    lr = m["lr"]
    pix = m["pixels"] // n_frames

    def mapping_call():
        # Mapper.optimize_mapping's per-call loop (fresh Adam, joint pose optimisation, 15 iterations)
        map_window(store, ws, sc, poses, window, window, m["pixels"], m["iters"], lr["decoders_lr"], lr["planes_lr"],
                   lr["c_planes_lr"], True, m["joint_opt_cam_lr"], exchange=ex)

    sampler = ClockSampler(local).start() if rank == 0 else None
    stages = {}
    if dist_on:
        stages["start"] = replicas_identical(store.arena, world)
    l0 = _lib.LAUNCHES
    ms_map = time_region(mapping_call, args.steps, args.warmup, dist_on)
    if dist_on:
        stages["after_timed_mapping_calls"] = replicas_identical(store.arena, world)
    launches = (_lib.LAUNCHES - l0) * args.steps // (args.steps + args.warmup)
    rays_per_step = m["iters"] * pix * n_frames
    value = world * rays_per_step * args.steps / (ms_map * 1e-3)

    # ------------------------------------------------------------------ dominant kernel alone (roofline)
    import ctypes as C
    from myslam_b200._lib import call, ptr, stream

    store.reset_adam()
    poses7 = torch.zeros(n_frames, 7, device=dev)
    poses7[1:] = matrix_to_cam_pose(poses[1:])
    mapping_iteration(ws, store, sc, poses, poses7, cols, deps, pix, 1, 1e-3, 5e-3, 5e-3, 1e-3, apply_adam=False)
    idx = torch.randint(spec["H"] * spec["W"], (pix * n_frames,), device=dev)
    R = int(ws.counters[0])

    q_img, gq_img = store.ensure_q(), store.ensure_q_grad()

    def bwd_kernel():
        call("eslam_loss_backward_q", store.ref(), ptr(store.arena), ptr(q_img), ptr(gq_img), C.byref(sc.cam),
             C.byref(sc.render), ptr(ws.rays_o), ptr(ws.rays_d), ptr(ws.z), ptr(ws.gt_depth), ptr(ws.gt_color),
             ptr(ws.src), ptr(idx), pix, None, ptr(ws.counters), None, pix * n_frames, ptr(store.grad),
             ptr(ws.pose_grad), None, stream())

    ms_k = time_region(bwd_kernel, 50, 5, False) / 50
    gq_img.zero_()
    store.grad.zero_()
    hbm_peak, peak_src, _ = measured_peaks()
    achieved = R * ALGO_BYTES_PER_RAY_ITER / (ms_k * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic("k_map_bwd_q")
    # (not under --profile-only: ncu would replay every launch of the microbenchmark subprocess as well)
    ceil = l2_ceilings() if rank == 0 and not args.profile_only else None
    # what the Q-form kernel itself moves through L2: 64-byte lines (16 channels) where the algorithmic figure of
    # SURVEY 8d counts 128-byte lines, i.e. half of it (before the run-length merge of the coarse reductions)
    q_bytes = R * ALGO_BYTES_PER_RAY_ITER / 2
    roofline = {"bound": "hbm", "kernel": "k_map_bwd_q<poses> (eslam_loss_backward_q)",
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "us_per_launch": 1e3 * ms_k, "rays_per_launch": R,
                "dram_frac": (traffic / (ms_k * 1e-3) / 1e9 / hbm_peak) if traffic else None,
                "l2": None,
                "note": "achieved = SURVEY 8d's algorithmic bytes (rays x 40 samples x 6144 B x (gather+scatter)) / "
                        "kernel time, as the contract defines it; it is NOT a bandwidth the kernel sustains: the 27 MB "
                        "plane set is L2-resident (see dram_frac) and the kernel works on 16-channel images, moving half "
                        "those bytes.  The limit that applies is the L2 gather + reduction ceiling: see l2."}
    if ceil:
        peak_l2 = ceil["line64"]["ldg_red"]
        roofline["l2"] = {"kernel_bytes": q_bytes, "achieved": q_bytes / (ms_k * 1e-3) / 1e9, "peak": peak_l2,
                          "unit": "GB/s", "frac": q_bytes / (ms_k * 1e-3) / 1e9 / peak_l2,
                          "peak_source": "tools/microbench/l2_gather_red (run inside this bench): random 64-byte "
                                         "lines out of a 27 MB set, LDG.128 gathers interleaved with "
                                         "red.global.add.v4.f32, best of 2/4/8 CTAs per SM",
                          "ceilings_GBps": ceil}

    # ------------------------------------------------------------------ tracking (1 GPU: too few rays to shard)
    # the tracker shares this process's FieldStore and re-imports the reference-layout planes when it first runs
    # (Tracker.py:222-232 semantics): publish the arena to them first, as optimize_mapping does at the end of a call,
    # so rank 0's parameters stay what the other ranks hold
    store.push_planes(scene.all_planes)
    store.push_decoders(scene.decoders)
    tracking = None
    e2e = None
    if rank == 0:
        trk = M.TrackerStep(cfg, rnd, scene.decoders, scene.all_planes, scene.bound, scene.cam, dev)
        pose0 = matrix_to_cam_pose(poses[:1])
        col1, dep1 = cols[:1].contiguous(), deps[:1].contiguous()

        def track_call():
            trk.track_frame(pose0, col1, dep1)

        l1 = _lib.LAUNCHES
        ms_trk = time_region(track_call, args.steps, args.warmup, False)
        trk_launches = (_lib.LAUNCHES - l1) // (args.steps + args.warmup)
        fps = args.steps / (ms_trk * 1e-3)

        # e2e tracking: drop-in optimize_tracking + torch optimizer, frame from pinned host memory
        h_col = cols[0].cpu().pin_memory()
        h_dep = deps[0].cpu().pin_memory()

        pre_t = FramePrefetcher([h_col, h_dep], dev)

        def track_e2e():
            frame = pre_t.next()
            gc, gd = frame[0][None], frame[1][None]
            T = torch.nn.Parameter(pose0[:, -3:].clone())
            Rq = torch.nn.Parameter(pose0[:, :4].clone())
            opt = torch.optim.Adam([{"params": [T], "lr": t["lr_T"], "betas": (0.5, 0.999)},
                                    {"params": [Rq], "lr": t["lr_R"], "betas": (0.5, 0.999)}])
            best = float("inf")
            for _ in range(t["iters"]):
                pose = torch.cat([Rq, T], -1)
                loss = trk.optimize_tracking(pose, gc, gd, t["pixels"], opt)  # .item() inside: D2H each iteration
                best = min(best, loss)
            pre_t.done(frame)

        def track_e2e_fused():
            frame = pre_t.next()
            best_pose, losses, _ = trk.track_frame(pose0, frame[0][None], frame[1][None])
            pre_t.done(frame)
            best_pose.cpu()  # the pose the caller stores (Tracker.py:309): one D2H per frame

        ms_trk_e2e = time_region(track_e2e, max(args.steps // 2, 2), 2, False)
        fps_e2e = max(args.steps // 2, 2) / (ms_trk_e2e * 1e-3)
        ms_trk_fused = time_region(track_e2e_fused, max(args.steps // 2, 2), 2, False)
        fps_fused = max(args.steps // 2, 2) / (ms_trk_fused * 1e-3)
        tracking = {"value": fps, "unit": "frames/s", "ms_per_frame": ms_trk / args.steps,
                    "iters_per_frame": t["iters"], "rays_per_iter": t["pixels"], "gpu_launches_per_frame": trk_launches,
                    "e2e": {"value": fps_e2e, "unit": "frames/s",
                            "h2d_bytes_per_step": h_col.numel() * 8 + h_dep.numel() * 4,
                            "d2h_bytes_per_step": 8 * t["iters"],
                            "api": "Tracker.optimize_tracking drop-in x 8 with the caller's torch.optim.Adam and a "
                                   "loss.item() per iteration (reference signature); frame staged from pinned host "
                                   "memory one step ahead on a copy stream"},
                    "e2e_track_frame": {"value": fps_fused, "unit": "frames/s",
                                        "h2d_bytes_per_step": h_col.numel() * 8 + h_dep.numel() * 4,
                                        "d2h_bytes_per_step": 28,
                                        "api": "TrackerStep.track_frame: the same 8 iterations with the pose Adam "
                                               "fused on the device, one host sync per frame"},
                    "roofline_frac_hbm": fps * t["iters"] * t["pixels"] * ALGO_BYTES_PER_RAY_ITER / 1e9 / hbm_peak}

    # ------------------------------------------------------------------ e2e mapping through the drop-in
    kf = [{"gt_c2w": poses[k], "idx": torch.tensor(4 * k), "color": cols[k], "depth": deps[k],
           "est_c2w": poses[k].clone()} for k in range(n_frames - 1)]
    mp.keyframe_dict = kf
    mp.joint_opt = True
    mp.cfg["mapping"]["mapping_window_size"] = n_frames
    mp.mapping_window_size = n_frames
    h_col = cols[-1].cpu().pin_memory()
    h_dep = deps[-1].cpu().pin_memory()
    h_c2w = poses[-1].cpu().pin_memory()
    kf_list = list(range(0, 4 * (n_frames - 1), 4))
    if dist_on:
        dist.barrier()  # rank 0 has been timing the tracker; line the ranks up before kernels that wait on peers
    mp.exchange = ex  # N > 1: every rank stages the same window from its own host memory and draws its own rays

    pre_m = FramePrefetcher([h_col, h_dep, h_c2w], dev)

    def mapping_e2e():
        frame = pre_m.next()
        gc, gd, cw = frame
        out = mp.optimize_mapping(m["iters"], 1.0, torch.tensor(4 * n_frames), gc, gd, cw, kf, kf_list, cw)
        pre_m.done(frame)
        out.cpu()

    n_e2e = max(args.steps // 2, 2)
    ms_e2e = time_region(mapping_e2e, n_e2e, 2, dist_on)
    if dist_on:
        stages["after_e2e_dropin_calls"] = replicas_identical(store.arena, world)
    e2e = {"value": world * rays_per_step * n_e2e / (ms_e2e * 1e-3), "unit": "rays*iters/s",
           "h2d_bytes_per_step": h_col.numel() * 8 + h_dep.numel() * 4 + 64, "d2h_bytes_per_step": 64,
           "ms_per_step": ms_e2e / n_e2e,
           "api": "MapperStep.optimize_mapping (reference signature): window selection + 20-frame table + "
                  "15 fused iterations + write-back of planes/decoders to the reference's NCHW tensors; the current "
                  "frame is staged from pinned host memory one step ahead on a copy stream, the returned pose is "
                  "read back every step"
                  + ("; per rank, bytes are per rank" if dist_on else "")}
    # the same drop-in call with the reference's random-draw SHAPES (strict_rng: seed-compatible with the reference,
    # one host sync per iteration)
    mp.strict_rng = True
    ms_strict = time_region(mapping_e2e, max(n_e2e // 2, 2), 1, dist_on)
    mp.strict_rng = False
    e2e_strict = {"value": world * rays_per_step * max(n_e2e // 2, 2) / (ms_strict * 1e-3), "unit": "rays*iters/s",
                  "ms_per_step": ms_strict / max(n_e2e // 2, 2),
                  "what": "e2e with ESLAM_B200_STRICT_RNG semantics: uniforms drawn as [R1,S] / [R0,32] / [R0,8] like "
                          "Renderer.py:59 / common.py:59, so a seeded run consumes torch's generator exactly as the "
                          "reference does (the default draws fixed [N,S] blocks and never syncs)"}
    if hasattr(ex, "check"):
        ex.check()
    clocks = sampler.stop() if sampler else None
    # ---- config 4 of BASELINE.json: 8 x 4000 = 32 000 rays per iteration, SHARDED over the ranks (strong scaling)
    strong = None
    try:
        from myslam_b200.hotpath import Workspace

        pix32 = (32000 // n_frames) // world
        ws32 = Workspace(dev, pix32 * n_frames, sc.render.n_stratified + sc.render.n_importance, max(32, n_frames))

        def call32():
            map_window(store, ws32, sc, poses, cols, deps, pix32 * n_frames, m["iters"], lr["decoders_lr"],
                       lr["planes_lr"], lr["c_planes_lr"], True, m["joint_opt_cam_lr"], exchange=ex)

        ms32 = time_region(call32, max(args.steps // 4, 2), 1, dist_on)
        n32 = max(args.steps // 4, 2)
        strong = {"rays_per_iter_total": pix32 * n_frames * world, "value": world * m["iters"] * pix32 * n_frames * n32
                  / (ms32 * 1e-3), "unit": "rays*iters/s", "ms_per_call": ms32 / n32, "scaling": "strong",
                  "n_gpus": world, "what": "SURVEY 8d config 4: the 32 000-ray batch of a 20-keyframe window split "
                  "evenly over the ranks (every rank renders 1600 / world pixels of every frame), same exchange"}
        del ws32
    except Exception as exc:  # noqa: BLE001
        strong = {"error": repr(exc)}
    # ---- multi-GPU parity inside the driver's run (its test box has one GPU): two iterations from the same state
    # through the peer-memory exchange and through NCCL all-reduce + replicated step
    exchange_check = None
    if dist_on and isinstance(ex, PeerExchange):
        a0 = store.arena.clone()

        def two_iters(exch, seed):
            store.arena.copy_(a0)
            store.gen += 1
            torch.manual_seed(seed + rank)
            map_window(store, ws, sc, poses, cols, deps, m["pixels"], 2, lr["decoders_lr"], lr["planes_lr"],
                       lr["c_planes_lr"], True, m["joint_opt_cam_lr"], exchange=exch)
            torch.cuda.synchronize()
            return store.arena.clone()

        peer = two_iters(ex, 4242)
        nccl = two_iters(MappingExchange(), 4242)
        store.arena.copy_(a0)
        store.gen += 1
        exchange_check = {"bit_identical_replicas": replicas_identical(peer, world),
                          "bit_identical_replicas_nccl_path": replicas_identical(nccl, world),
                          "replicas_identical_at": {**stages, "check_start": replicas_identical(a0, world)},
                          "max_rel_vs_nccl": ((peer - nccl).abs().max() / nccl.abs().max()).item(),
                          "what": "2 joint-opt mapping iterations (own rays per rank) from the same state: peer-memory "
                                  "exchange vs NCCL all-reduce of the gradient images + replicated optimiser step; "
                                  "replicas compared by a checksum of the parameter arena's bits over all ranks"}
        ex.check()
    extra = {}
    if not args.profile_only:
        try:
            extra = extras(scene, rnd, spec, dev, rank, world, dist_on, poses, deps, hbm_peak)
        except Exception as exc:  # secondary numbers never take the headline line down
            extra = {"error": repr(exc)}

    # ------------------------------------------------------------------ baselines (rank 0, N=1 only)
    cpu_baseline = None
    torch_gpu = None
    if rank == 0 and world == 1 and not args.no_baselines:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        n_it = 5
        cpu = OracleRun(spec, n_frames)
        cpu.run_mapping(1)
        s_iter = cpu.run_mapping(n_it)
        cpu.run_tracking(1)
        s_trk = cpu.run_tracking(4)
        del cpu
        cpu_baseline = {"value": m["pixels"] / s_iter, "unit": "rays*iters/s", "cores": cores, "kind": "port",
                        "sample": f"{n_it} mapping iterations of 4000 rays over the same 20-frame window "
                                  f"(oracle port of the reference's PyTorch path, torch CPU, {cores} threads)",
                        "tracking_frames_per_s": 1.0 / (s_trk * t["iters"])}
        gpu = OracleRun(spec, n_frames, device=dev)
        gpu.run_mapping(5)
        g_runs = sorted(gpu.run_mapping(30) for _ in range(3))
        gpu.run_tracking(8)
        t_runs = sorted(gpu.run_tracking(32) for _ in range(3))
        del gpu
        g_iter, g_trk = g_runs[1], t_runs[1]
        torch_gpu = {"mapping_rays_iters_per_s": m["pixels"] / g_iter, "tracking_frames_per_s": 1.0 / (g_trk * t["iters"]),
                     "mapping_best_of_3": m["pixels"] / g_runs[0], "mapping_worst_of_3": m["pixels"] / g_runs[2],
                     "tracking_best_of_3": 1.0 / (t_runs[0] * t["iters"]),
                     "what": "the reference's PyTorch path (oracle port, stock ATen kernels, eager) on the same B200: "
                             "median of 3 x 30 mapping iterations / 3 x 32 tracking iterations after warm-up",
                     "vs_torch_gpu": {"mapping": value / (m["pixels"] / g_iter),
                                      "tracking": tracking["value"] / (1.0 / (g_trk * t["iters"])) if tracking else None}}

    if rank == 0:
        line = {
            "metric": "mapping rays*iters/s (Replica room0 shape)", "value": value, "unit": "rays*iters/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_map / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world, pix, n_frames, exchange_name),
            "e2e": e2e, "e2e_strict": e2e_strict, "gpu_launches": launches, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "torch_gpu_baseline": torch_gpu, "tracking": tracking, "clocks": clocks,
            "mapping_32k": strong, "exchange_check": exchange_check, **extra,
        }
        print(json.dumps(line))
    if dist_on:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
