"""One process per GPU, data parallel over NCCL / NVLink.

The reference's only parallelism is nn.DataParallel threads (models/base_model.py:104-108) and one
tmux pane per material (experiments/tmux_launcher.py:87-124).  Here every rank owns a full replica
of G / D / D2 and its own (material, augmentation) sample; the only exchange is one all-reduce of
each net's flat fp32 gradient bucket right after its backward (3 per step), and the mean is folded
into the Adam kernel (grad_scale = 1/world_size).  BatchNorm statistics stay rank-local, like
DataParallel replicas.  Works with the gloo backend on CPU tensors for the host-logic tests."""
import os

import torch
import torch.distributed as dist


def launched_by_torchrun():
    """True when the process was started by torchrun / torch.distributed.run with more than one rank."""
    return int(os.environ.get("WORLD_SIZE", "1")) > 1 and "RANK" in os.environ


class DistContext:
    def __init__(self, backend=None):
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        self.backend = backend
        if self.world_size > 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
            dist.init_process_group(backend, rank=self.rank, world_size=self.world_size)

    def allreduce_grads(self, flat_grad):
        """Sum the flat gradient bucket over ranks (the mean is applied by the optimiser kernel)."""
        if self.world_size > 1:
            dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)

    def barrier(self):
        if self.world_size > 1:
            dist.barrier()

    def max_over_ranks(self, value):
        if self.world_size == 1:
            return float(value)
        t = torch.tensor([float(value)], dtype=torch.float64, device="cuda" if self.backend == "nccl" else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sample_indices(self, n_items):
        """Round-robin partition: rank r takes items r, r+W, ... (DistributedSampler without shuffling;
        matches `material_index = index % len(material_list)`, data/skit_dataset.py:240)."""
        return list(range(self.rank, n_items, self.world_size))

    def broadcast_params(self, nets):
        """Make every replica start from rank 0's parameters and buffers."""
        if self.world_size > 1:
            for net in nets:
                dist.broadcast(net.flat_param, src=0)
                for b in net.buffers():
                    if b.dtype.is_floating_point:
                        dist.broadcast(b, src=0)

    def shutdown(self):
        if self.world_size > 1 and dist.is_initialized():
            dist.destroy_process_group()
