"""Sharded runs of the drop-in `Solver` (no reference counterpart -- SURVEY.md 8(e)).

Launched as one process per GPU (`FDTD_SHARDED=1 torchrun --nproc-per-node N script.py`, the script unchanged, or
`Solver(..., sharded=True)`), every rank builds the same `Solver`; the grid's rows are split into y-slabs (slab.split_rows), each rank's `Engine` owns one slab and the C
library exchanges the halo rows over NCCL by itself (fdtd_comm_init).  `torch.distributed` is only the control plane:
it carries the 128-byte NCCL id, the slabs when the script reads `solver.h`, and the probe records of the owning rank.

All helpers are COLLECTIVE: every rank must call them in the same order, which an SPMD script does by construction.
"""
import os


class Shard:
    def __init__(self, dist, rank, world, local_rank):
        self.dist, self.rank, self.world, self.local_rank = dist, rank, world, local_rank

    def all_gather(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def broadcast(self, obj, src=0):
        box = [obj if self.rank == src else None]
        self.dist.broadcast_object_list(box, src=src)
        return box[0]

    def barrier(self):
        self.dist.barrier()


def context(enable=None):
    """The `Shard` of this process, or None for an ordinary single-process run.

    enable=None : opt-in by environment -- with FDTD_SHARDED=1 shard iff a process group with more than one rank is up or
                  the process was started by torchrun (WORLD_SIZE > 1), so `FDTD_SHARDED=1 torchrun ... script.py` shards
                  an unchanged script; without it never (a torchrun launch of INDEPENDENT simulations per rank stays
                  independent).
    enable=False: never.
    enable=True : shard; a single-process launch is an error.
    Under torchrun the process group is initialised here if it is not up yet (NCCL with a GPU, gloo without)."""
    if enable is False:
        return None
    strict = enable is True
    if enable is None and os.environ.get("FDTD_SHARDED", "") in ("", "0", "false", "no"):
        return None
    import torch
    import torch.distributed as dist
    if not dist.is_available():
        if strict:
            raise RuntimeError("sharded=True needs torch.distributed")
        return None
    if not dist.is_initialized():
        if int(os.environ.get("WORLD_SIZE", "1")) <= 1:
            if strict:
                raise RuntimeError("sharded=True needs a torchrun launch (WORLD_SIZE > 1) or an initialised process group")
            return None
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        local = int(os.environ.get("LOCAL_RANK", "0"))
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo")
    world = dist.get_world_size()
    if world <= 1:
        if strict:
            raise RuntimeError("sharded=True needs more than one rank")
        return None
    local = int(os.environ.get("LOCAL_RANK", str(dist.get_rank())))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)              # object collectives over NCCL stage through the current device
    return Shard(dist, dist.get_rank(), world, local)


def owner_of_row(ranges, row):
    """Rank whose [begin, end) holds the global h/ey row `row`."""
    for r, (b, e) in enumerate(ranges):
        if b <= row < e:
            return r
    raise IndexError("row %d outside the grid" % row)
