"""Ray-sharded multi-GPU mapping (not in the reference, which is single-GPU):
one process per GPU, parameters replicated, every rank renders its own slice of the keyframe-ray
batch, and two exchanges per iteration make all ranks differentiate the SAME global-batch loss:

  1. all-reduce of the 8 int32 loss normalisers (ray / band counts) -> `norm` for eslam_loss_backward_q
  2. all-reduce (sum, fp32) of the map gradients (the 16-channel gradient images of the 12 planes, the decoder
     block) and of the [frames,12] pose-gradient block, then the identical Adam step.

`MappingExchange` does both with NCCL collectives (and with gloo on CPU tensors for the world_size-2 tests).
`PeerExchange` is the B200 path: parameters, a gradient-image staging block, published decoder-gradient /
pose-gradient / loss / counter blocks and a flag block live in ONE symmetric (NVLink peer-mapped) allocation per
rank, the normalisers are summed by a one-CTA kernel over peer loads, and `eslam_q_adam_exchange` does reduce-scatter
of the gradient images + plane Adam + all-gather + zero_grad over peer memory (P2P loads/stores, or multimem.st through
the NVSwitch), followed by the decoders' replicated step on the sum of the published decoder gradients.
No NCCL call is left on the per-iteration path; torch.distributed only sets the allocation up.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist


class MappingExchange:
    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._norm = None

    def reduce_counters(self, counters: torch.Tensor) -> torch.Tensor:
        if self._norm is None or self._norm.device != counters.device:
            self._norm = torch.empty_like(counters)
        self._norm.copy_(counters)
        if self.world > 1:
            dist.all_reduce(self._norm, op=dist.ReduceOp.SUM, group=self.group)
        return self._norm

    def reduce_grads(self, grads, pose_grad=None, loss_acc=None) -> None:
        """grads: the tensors holding this rank's map gradients (the gradient images and the decoder block)."""
        if self.world == 1:
            return
        for g in ([grads] if torch.is_tensor(grads) else grads):
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
        if pose_grad is not None:
            dist.all_reduce(pose_grad, op=dist.ReduceOp.SUM, group=self.group)
        if loss_acc is not None:
            dist.all_reduce(loss_acc, op=dist.ReduceOp.SUM, group=self.group)


    def after_step(self, store) -> None:
        """Keep the replicas bit-identical: the plane update is deterministic given the all-reduced gradient images,
        but dW1 is summed with floating-point atomics whose order differs from rank to rank, so the decoders (2 700
        floats) are taken from rank 0 after every step."""
        if self.world > 1:
            dist.broadcast(store.dec, src=0, group=self.group)


class PeerExchange(MappingExchange):
    """Symmetric-memory exchange for one FieldStore.  Construct it on every rank at the same point (it is
    collective); afterwards `store.arena` is a view of the symmetric allocation.  The gradient arena stays in
    ordinary device memory (reductions into peer-mapped memory are slower); peers receive it through a staging
    block."""

    def __init__(self, store, ws=None, group=None, multimem=None, max_frames=None):
        super().__init__(group)
        import torch.distributed._symmetric_memory as symm

        from . import _lib

        if not dist.is_initialized() or self.world < 2:
            raise RuntimeError("PeerExchange needs an initialised process group with at least 2 ranks")
        if self.world > _lib.MAX_PEERS:
            raise RuntimeError(f"PeerExchange supports up to {_lib.MAX_PEERS} GPUs of one box")
        dev = store.device
        self.store, self._side = store, None
        n = store.n_floats
        frames = max_frames or (ws.max_frames if ws is not None else 32)
        self.n_pose = frames * 12
        pose_blk = ((self.n_pose + 3) // 4) * 4
        flag_words = _lib.load().eslam_exchange_flag_words()
        n_stage = int(_lib.load().eslam_q_exchange_stage_floats(store.ref(), self.world))
        if n_stage <= 0:
            raise RuntimeError("PeerExchange: the plane layout cannot be sliced over the ranks")
        self.off_stage = n
        o = n + n_stage
        self.off_pose = [o, o + pose_blk]
        o += 2 * pose_blk
        self.off_loss = [o, o + 2 * _lib.N_LOSS]  # 16-byte aligned, so the float64 view is aligned
        o += 4 * _lib.N_LOSS
        self.off_cnt = [o, o + _lib.N_COUNTERS]
        o += 2 * _lib.N_COUNTERS
        self.off_dec = [o, o + _lib.DEC_FLOATS]  # published decoder gradients, two copies
        o += 2 * _lib.DEC_FLOATS
        # reduce_small() publishes through its own two copies (its calls interleave with adam_exchange's on another stream)
        self.off_pose_s = [o, o + pose_blk]
        o += 2 * pose_blk
        self.off_loss_s = [o, o + 2 * _lib.N_LOSS]
        o += 4 * _lib.N_LOSS
        self.off_flags = o
        total = o + flag_words
        grp = group if group is not None else dist.group.WORLD
        self.buf = symm.empty(total, dtype=torch.float32, device=dev)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, grp.group_name)
        self.rank = dist.get_rank(group)
        bases = [int(b) for b in self.handle.buffer_ptrs]
        mc = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        if multimem is None:
            # The all-gather of the updated texels: world_size point-to-point stores per element, or ONE multimem.st
            # through the NVSwitch.  Measured (tools/exchange_times.py): no difference at 2 GPUs, 14 us per iteration
            # less at 8 (235 vs 249 us for backward + exchange): the switch by default from 4 ranks up;
            # ESLAM_B200_MULTIMEM=0 / 1 forces either.
            env = os.environ.get("ESLAM_B200_MULTIMEM", "")
            multimem = env == "1" if env in ("0", "1") else self.world >= 4
        self.multimem = bool(multimem and mc)
        self._mc = mc

        def table(off_floats):
            arr = (C.c_void_p * _lib.MAX_PEERS)()
            for r in range(self.world):
                arr[r] = bases[r] + 4 * off_floats
            return arr

        self._param = table(0)
        self._stage = table(self.off_stage)
        self._pose = [table(o_) for o_ in self.off_pose]
        self._loss = [table(o_) for o_ in self.off_loss]
        self._cnt = [table(o_) for o_ in self.off_cnt]
        self._dec = [table(o_) for o_ in self.off_dec]
        self._pose_s = [table(o_) for o_ in self.off_pose_s]
        self._loss_s = [table(o_) for o_ in self.off_loss_s]
        self.peers = _lib.Peers()
        self.peers.rank, self.peers.world, self.peers.epoch, self.peers.adam_seq = self.rank, self.world, 0, 0
        for r in range(self.world):
            self.peers.flags[r] = bases[r] + 4 * self.off_flags
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        self.local_sync = torch.zeros(2, dtype=torch.int64, device=dev)
        self.peers.status = self.status.data_ptr()
        self.peers.local_sync = self.local_sync.data_ptr()
        self.norm = torch.zeros(_lib.N_COUNTERS, dtype=torch.int32, device=dev)
        self.pose_sum = torch.zeros(frames, 12, dtype=torch.float32, device=dev)
        self.loss_sum = torch.zeros(_lib.N_LOSS, dtype=torch.float64, device=dev)
        self._n_cnt = 0
        self._n_aux = 0
        # move the store's arenas into the symmetric allocation
        arena = self.buf[:n]
        arena.copy_(store.arena)
        store.arena = arena
        store.gen += 1
        store.ensure_grad()
        store.ensure_q_grad()
        torch.cuda.synchronize(dev)
        dist.barrier(group)  # nobody signals into a flag block that is not zeroed yet

    def _next_epoch(self):
        self.peers.epoch = (self.peers.epoch + 1) & 0x7FFFFFFF
        return C.byref(self.peers)

    def _counters_call(self, counters, cuda_stream):
        from ._lib import N_COUNTERS, call, ptr

        par = self._n_cnt & 1
        self._n_cnt += 1
        call("eslam_exchange_counters", self._next_epoch(), ptr(counters), self._cnt[par], N_COUNTERS, ptr(self.norm),
             cuda_stream)
        return self.norm

    def reduce_counters(self, counters: torch.Tensor) -> torch.Tensor:
        from ._lib import stream

        return self._counters_call(counters, stream())

    def begin_counters(self, counters: torch.Tensor) -> torch.Tensor:
        """reduce_counters on a side stream, so the peer round trip overlaps the kernels launched until
        end_counters() (the importance sampling of depth-less rays, which only reads the LOCAL counters)."""
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.store.device)
            self._ev_fork, self._ev_join = torch.cuda.Event(), torch.cuda.Event()
        self._ev_fork.record()
        self._side.wait_event(self._ev_fork)
        norm = self._counters_call(counters, self._side.cuda_stream)
        self._ev_join.record(self._side)
        return norm

    def end_counters(self) -> None:
        torch.cuda.current_stream().wait_event(self._ev_join)

    def adam_exchange(self, step: int, lr_dec: float, lr_planes: float, lr_cplanes: float, pose_grad=None,
                      n_pose_frames: int = 0, loss_acc=None, betas=(0.9, 0.999), eps=1e-8):
        """The fused optimiser step of the ray-sharded mapping on what eslam_loss_backward_q left on this rank
        (`store.gq_arena`, the decoder block of `store.grad`).  pose_grad [frames,12] / loss_acc float64[>=5] are this
        rank's local blocks: published, zeroed and summed over the ranks.  Returns (pose_grad_sum, loss_sum).  The
        gradient images and the decoder gradients are zero afterwards."""
        from ._lib import call, ptr, stream

        st = self.store
        if n_pose_frames * 12 > self.n_pose:
            raise RuntimeError("PeerExchange: more frames than the published pose block was sized for")
        par = self.peers.adam_seq & 1
        self.peers.adam_seq += 1
        n_aux = n_pose_frames * 12 if pose_grad is not None else 0
        n_auxd = 5 if loss_acc is not None else 0
        call("eslam_q_adam_exchange", self._next_epoch(), st.ref(), self._param, self._stage, ptr(st.gq_arena),
             ptr(st.grad), self._mc if self.multimem else None, ptr(st.exp_avg), ptr(st.exp_avg_sq), ptr(st.touched_q),
             lr_planes, lr_cplanes, lr_dec, step, betas[0], betas[1], eps, self._dec[par],
             ptr(pose_grad) if n_aux else None, self._pose[par], ptr(self.pose_sum), n_aux,
             ptr(loss_acc) if n_auxd else None, self._loss[par], ptr(self.loss_sum), n_auxd, stream())
        st.gen += 1
        return self.pose_sum, self.loss_sum

    def reduce_small(self, pose_grad=None, n_pose_frames: int = 0, loss_acc=None):
        """Sum of the ranks' pose-gradient blocks / loss terms on the CURRENT stream, apart from the plane exchange (the
        pipelined window loop runs it on its side stream).  Returns (pose_grad_sum, loss_sum); local blocks zeroed."""
        from ._lib import call, ptr, stream

        if n_pose_frames * 12 > self.n_pose:
            raise RuntimeError("PeerExchange: more frames than the published pose block was sized for")
        par = self._n_aux & 1
        self._n_aux += 1
        n_aux = n_pose_frames * 12 if pose_grad is not None else 0
        n_auxd = 5 if loss_acc is not None else 0
        call("eslam_exchange_aux", self._next_epoch(), ptr(pose_grad) if n_aux else None, self._pose_s[par],
             ptr(self.pose_sum), n_aux, ptr(loss_acc) if n_auxd else None, self._loss_s[par], ptr(self.loss_sum), n_auxd,
             stream())
        return self.pose_sum, self.loss_sum

    def check(self) -> None:
        """Raise if a peer ever failed to arrive at a barrier (host sync)."""
        if int(self.status.item()) != 0:
            raise RuntimeError("PeerExchange: a peer did not reach an exchange barrier within 4 s")


def shard_range(total: int, rank: int, world: int):
    """Contiguous [start, count) slice of `total` independent units for `rank` (mesh voxel blocks)."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)
