"""CPU, world_size 2, gloo: the data-parallel host logic (vts_b200/dist.py) — flat-bucket gradient
all-reduce, 1/W folded into the optimiser, round-robin sample partition, parameter broadcast — driven
through the oracle's train step (grad_hook) so that both replicas must end bit-identical and equal to a
single process that averages the two samples' gradients."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_state(seed):
    import argparse
    import vts_b200
    torch.manual_seed(seed)
    opt = argparse.Namespace(gan_mode="nonsaturating")
    G = vts_b200.networks.define_G(9, 5, 8, "resnet_4blocks", "instance", False, "xavier", 0.5, False, False, [], opt)
    D = vts_b200.networks.define_D(4, 8, "multiscale", 3, "batch", "xavier", 0.5, False, 3, [], opt)
    D2 = vts_b200.networks.define_D(7, 8, "multiscale", 3, "batch", "xavier", 0.5, False, 3, [], opt)
    return [{k: v.detach().clone() for k, v in n.state_dict().items()} for n in (G, D, D2)], (G, D, D2)


def _rand(rank):
    return dict(real_b=[0.3 + 0.1 * rank], real_s=[0.8], fake_b=[0.6], fake_s=[0.2 + 0.1 * rank],
                fake_ox=np.array([3, 10], dtype=np.int32), fake_oy=np.array([5, 1 + rank], dtype=np.int32))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    from vts_b200.dist import DistContext
    from oracle import skit_oracle as O
    ctx = DistContext(backend="gloo")
    assert ctx.sample_indices(5) == list(range(rank, 5, world))
    sds, nets = _make_state(seed=100 + rank)          # replicas start DIFFERENT; broadcast must fix that
    for net in nets:
        net.flatten_parameters()
    ctx.broadcast_params(nets)
    sds = [{k: v.detach().clone() for k, v in n.state_dict().items()} for n in nets]
    cfg = O.StepConfig(netG="resnet_9blocks", n_blocks=4, batch_size_G2=4, add_fake_T_sample_size=2)
    batch = O.step_inputs_from_batch(O.synthetic_batch(32, NT=4, seed=rank))

    def hook(name, grads):   # what SinSKITGModel._allreduce + _adam(grad_scale=1/W) do on the flat bucket
        keys = sorted(grads)
        flat = torch.cat([grads[k].reshape(-1) for k in keys])
        ctx.allreduce_grads(flat)
        flat /= ctx.world_size
        off = 0
        for k in keys:
            n = grads[k].numel()
            grads[k].copy_(flat[off:off + n].view_as(grads[k]))
            off += n

    O.train_step(cfg, sds[0], sds[1], sds[2], {}, batch, _rand(rank), step=1, grad_hook=hook)
    assert ctx.max_over_ranks(float(rank)) == float(world - 1)
    ctx.barrier()
    torch.save({"G": sds[0], "D": sds[1]}, os.path.join(out, "rank%d.pt" % rank))
    ctx.shutdown()


@pytest.mark.timeout(600)
def test_two_rank_data_parallel_step(tmp_path):
    world, port = 2, _free_port()
    mp.start_processes(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    r0 = torch.load(os.path.join(tmp_path, "rank0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "rank1.pt"))
    for net in ("G", "D"):
        for k, v in r0[net].items():
            if "running" in k or "num_batches" in k:
                continue  # BatchNorm statistics stay rank-local (like DataParallel replicas)
            assert torch.equal(v, r1[net][k]), (net, k)   # identical reduced grads -> bit-identical replicas
    # and the replicas moved: the step really applied an update
    sys.path.insert(0, ROOT)
    sds, _ = _make_state(seed=100)
    moved = sum(float((r0["G"][k] - sds[0][k]).abs().sum()) for k in sds[0] if sds[0][k].dtype.is_floating_point)
    assert moved > 0


def test_single_process_context_is_a_noop():
    sys.path.insert(0, ROOT)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        os.environ.pop(k, None)
    from vts_b200.dist import DistContext
    ctx = DistContext(backend="gloo")
    g = torch.ones(8)
    ctx.allreduce_grads(g)
    assert ctx.world_size == 1 and float(g.sum()) == 8.0 and ctx.max_over_ranks(3.5) == 3.5
    assert ctx.sample_indices(4) == [0, 1, 2, 3]
