"""CPU, world_size 2, gloo: the multi-rank host logic of the data-parallel path — constructor
broadcast + flat gradient averaging of FlatAllReduceDDP, the rank-averaged GradNorm backward
(vae_trainer.py:27-60) and avg_scalar_over_nodes."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q, use_store, ranges):
    sys.path.insert(0, os.path.join(ROOT, "vqgan-training_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      VQB_OFFLINE="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import vae_trainer as vt

    torch.manual_seed(100 + rank)  # different init per rank: the wrapper must broadcast rank 0's weights
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    ddp = vt.FlatAllReduceDDP(net)
    if use_store:  # the Trainer's configuration: gradients live in one flat buffer, reduced in place (in k async ranges)
        import flat

        store = flat.FlatParams(net.parameters())
        ddp.attach_store(store, overlap_ranges=ranges)
        assert (ddp._ranges is not None) == (ranges >= 2)
        if ranges >= 2:
            assert len(ddp._ranges) == ranges and sum(len(r["idx"]) for r in ddp._ranges) == 4
    w0 = [p.detach().clone() for p in net.parameters()]
    g = torch.Generator().manual_seed(7 + rank)
    x = torch.randn(4, 6, generator=g)
    # the rank-local gradient, from an unwrapped twin (with overlap on, p.grad is already being averaged after backward)
    import copy
    twin = copy.deepcopy(net)
    yt = twin(x)
    (vt.gradnorm(yt, 0.5).pow(2).sum() * (rank + 1)).backward()
    local = [p.grad.detach().clone() for p in twin.parameters()]
    y = ddp.module(x)  # calling .module directly, like the reference loop does (vae_trainer.py:538,624)
    loss = vt.gradnorm(y, 0.5).pow(2).sum() * (rank + 1)
    loss.backward()
    ddp.allreduce_grads()
    avg = [p.grad.detach().clone() for p in net.parameters()]
    if use_store:  # every gradient now lives in its slot of the flat buffer
        assert all(p.grad.data_ptr() == store.grads.data_ptr() + 4 * o for p, o in zip(store.plist, store.offsets))
    s = vt.avg_scalar_over_nodes(float(rank + 1), "cpu")
    st = vt.avg_scalar_over_nodes(torch.tensor(float(rank + 1)), "cpu")
    # local norm of the gradient entering GradNorm (for the analytic check in the parent)
    gn_in = (2 * y.detach() * (rank + 1))
    tl = lambda ts: [t_.tolist() for t_ in ts]  # plain lists: tensors in an mp.Queue die with the worker
    q.put((rank, tl(w0), tl(local), tl(avg), s, float(st), gn_in.norm().item(), y.detach().tolist(), x.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("use_store,ranges", [(False, 0), (True, 0), (True, 2), (True, 4)])
def test_flat_allreduce_gradnorm_world2(use_store, ranges):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200) + ranges + int(use_store)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, use_store, ranges)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    T = lambda ls: [torch.tensor(l) for l in ls]
    (r0, w0a, loc0, avg0, s0, st0, n0, y0, x0), (r1, w0b, loc1, avg1, s1, st1, n1, y1, x1) = res
    w0a, w0b, loc0, loc1, avg0, avg1 = T(w0a), T(w0b), T(loc0), T(loc1), T(avg0), T(avg1)
    x0 = torch.tensor(x0)
    for a, b in zip(w0a, w0b):  # constructor broadcast
        assert torch.equal(a, b)
    for l0, l1, a0, a1 in zip(loc0, loc1, avg0, avg1):  # gradient averaging
        assert torch.allclose(a0, (l0 + l1) / 2, atol=1e-6) and torch.equal(a0, a1)
        assert not torch.allclose(l0, l1)
    assert s0 == s1 == 1.5 and st0 == st1 == 1.5
    # GradNorm: both ranks divided by the SAME rank-averaged norm
    nbar = (n0 + n1) / 2
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    with torch.no_grad():
        for p, w in zip(net.parameters(), w0a):
            p.copy_(w)
    yy = net(x0)
    yy.backward(0.5 * (2 * yy.detach() * 1) / (nbar + 1e-8))
    for p, l in zip(net.parameters(), loc0):
        assert torch.allclose(p.grad, l, atol=1e-5)
