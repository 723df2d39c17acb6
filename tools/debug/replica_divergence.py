"""2-rank debug: where do the replicas of the e2e drop-in path diverge?  (torchrun, 2 GPUs)"""
import os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import myslam_b200 as M
from bench import build_inputs, replicas_identical
from myslam_b200 import synthetic as S
from myslam_b200.decoders import synced_store
from myslam_b200.dist import PeerExchange
from myslam_b200.mapper import _mapper_state, map_window

def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local); dev = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    spec = S.REPLICA_ROOM0; m = spec["mapping"]; nf = m["mapping_window_size"]
    scene = S.make_scene(spec, dev, seed=0); cfg = S.run_cfg(spec)
    class E: pass
    e = E(); e.bound, e.device = scene.bound, dev; e.H, e.W, e.fx, e.fy, e.cx, e.cy = scene.cam
    rnd = M.Renderer(cfg, e)
    poses, cols, deps = build_inputs(spec, dev, nf, seed=1); poses = poses.to(dev)
    torch.manual_seed(1234 + rank)
    mp = M.MapperStep(cfg, rnd, scene.decoders, scene.all_planes, scene.bound, scene.cam, dev)
    st = _mapper_state(mp, m["pixels"], nf)
    store = synced_store(scene.all_planes, scene.decoders, scene.bound)
    ex = PeerExchange(store, st["ws"])
    lr = m["lr"]
    def chk(tag, extra=None):
        torch.cuda.synchronize()
        pl = replicas_identical(store.arena[:store.n_planes_end], world)
        dc = replicas_identical(store.arena[store.dec_off:store.dec_off + 2700], world)
        ex_ = replicas_identical(extra, world) if extra is not None else None
        if rank == 0: print(f"{tag:40s} planes {pl} decoders {dc} extra {ex_}", flush=True)
    chk("start")
    for i in range(2):
        map_window(store, st["ws"], st["sc"], poses, cols, deps, m["pixels"], 15, lr["decoders_lr"], lr["planes_lr"], lr["c_planes_lr"], True, m["joint_opt_cam_lr"], exchange=ex)
        chk(f"map_window {i}")
    kf = [{"gt_c2w": poses[k], "idx": torch.tensor(4 * k), "color": cols[k], "depth": deps[k], "est_c2w": poses[k].clone()} for k in range(nf - 1)]
    mp.keyframe_dict = kf; mp.joint_opt = True; mp.cfg["mapping"]["mapping_window_size"] = nf; mp.mapping_window_size = nf
    mp.exchange = ex
    kf_list = list(range(0, 4 * (nf - 1), 4))
    for i in range(4):
        out = mp.optimize_mapping(15, 1.0, torch.tensor(4 * nf), cols[-1], deps[-1], poses[-1], kf, kf_list, poses[-1])
        allp = torch.stack([k["est_c2w"] for k in kf] + [out]).contiguous().float()
        chk(f"optimize_mapping {i}", allp)
    ex.check()
    dist.barrier(); dist.destroy_process_group()
main()
