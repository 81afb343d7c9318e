"""Data-parallel host logic on CPU: world_size 2, gloo backend (the GPU path uses the same code over NCCL).

Checks the flat parameter/gradient buffers of trainer._FlatNet, the rank-0 parameter broadcast, the summed
gradient all-reduce with the 1/world scale handed to Adam, and the per-rank seeding."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dcvgan_b200 import trainer
    torch.manual_seed(100 + rank)                      # deliberately different initial weights per rank
    model = torch.nn.Sequential(torch.nn.Conv2d(3, 5, 3, bias=False), torch.nn.BatchNorm2d(5), torch.nn.Conv2d(5, 7, 1, bias=False))
    opt = torch.optim.Adam(model.parameters(), lr=2e-4, betas=(0.5, 0.999), weight_decay=1e-5)
    flat = trainer._FlatNet(model, opt)
    # parameters are views into the flat buffer, gradients too, Adam moments registered in opt.state
    for p in model.parameters():
        assert p.data_ptr() >= flat.flat_p.data_ptr() and p.grad.data_ptr() >= flat.flat_g.data_ptr()
        assert opt.state[p]["exp_avg"].shape == p.shape
    trainer.dp_sync_params([flat])
    w0 = flat.flat_p.clone()
    for i, p in enumerate(model.parameters()):
        p.grad.fill_(float(rank + 1) * (i + 1))        # rank-specific gradient written through the views
    scale = trainer.dp_allreduce_grads([flat])
    # the bucketed reducer of the fused step: launch per network, wait before Adam (synchronous on CPU tensors)
    red = trainer.GradReducer()
    flat2 = trainer._FlatNet(torch.nn.Linear(3, 2), torch.optim.Adam(torch.nn.Linear(3, 2).parameters()))
    flat2.flat_g.fill_(float(rank + 1))
    red.launch(flat2)
    assert red.wait() == 0.5 and float(flat2.flat_g[0]) == 3.0
    seed = trainer.dp_seed(15)
    draw = torch.randn(3)
    out[rank] = (w0, flat.flat_g.clone(), scale, seed, draw, [float(p.grad.flatten()[0]) for p in model.parameters()])
    dist.destroy_process_group()


def test_dp_flat_buffers_allreduce_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    (w0a, ga, sa, seeda, da, firsta), (w0b, gb, sb, seedb, db, firstb) = out[0], out[1]
    assert torch.equal(w0a, w0b)                                   # rank 0's parameters were broadcast
    assert torch.equal(ga, gb) and sa == sb == 0.5                 # summed gradients, averaged later by Adam's grad_scale
    assert firsta == [3.0 * (i + 1) for i in range(len(firsta))]   # (1 + 2) * (i + 1) seen through every p.grad view
    assert (seeda, seedb) == (15, 16) and not torch.equal(da, db)  # replicas draw different noise


def test_single_process_is_identity():
    from dcvgan_b200 import trainer
    model = torch.nn.Linear(4, 4)
    opt = torch.optim.Adam(model.parameters())
    flat = trainer._FlatNet(model, opt)
    model.weight.grad.fill_(2.0)
    assert trainer.dp_allreduce_grads([flat]) == 1.0 and float(flat.flat_g[0]) == 2.0
    x = torch.randn(2, 4)
    assert torch.allclose(model(x), torch.nn.functional.linear(x, model.weight, model.bias))   # views still drive the module
