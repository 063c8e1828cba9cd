"""N > 1 host logic on CPU: two processes over the gloo backend build their subdomains
independently and run the interface sum-exchange + owned-DOF dot product with the index sets the
device path uses (halo_sum_exchange / halo_allreduce do the same over NCCL)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import jots_b200 as J
    m = J.Mesh.brick(4, 3, 2, 1.0, 1.0, 1.0)
    sp = J.Space(m, 2)
    prt = J.Partition(m, sp, world, rank)
    g, el = prt.global_ids()
    lsp = prt.space()
    # local partial sums of a global "assembly": each element adds (1 + global element id) to its DOFs
    w = np.repeat(1.0 + el.astype(float), lsp.dofs_per_elem)
    v = np.bincount(lsp.gather().ravel(), weights=w, minlength=prt.n_local)
    # pack / exchange / canonical-order reduction
    nb = prt.neighbors()
    recv = [torch.empty(idx.size, dtype=torch.float64) for _, idx in nb]
    reqs = []
    for (r, idx), buf in zip(nb, recv):
        reqs.append(dist.isend(torch.from_numpy(v[idx].copy()), dst=r))
        reqs.append(dist.irecv(buf, src=r))
    for rq in reqs:
        rq.wait()
    rb = np.concatenate([b.numpy() for b in recv]) if recv else np.zeros(0)
    dof, ptr, src = prt.reduction()
    for s in range(dof.size):
        acc = 0.0
        for k in range(ptr[s], ptr[s + 1]):
            acc += v[dof[s]] if src[k] < 0 else rb[src[k]]
        v[dof[s]] = acc
    # global reference
    gg = sp.gather()
    ref = np.bincount(gg.ravel(), weights=np.repeat(1.0 + np.arange(sp.n_elem, dtype=float), sp.dofs_per_elem), minlength=sp.n_dof)
    ok_vec = bool(np.array_equal(v, ref[g]))
    # dot product over owned DOFs + all-reduce == global dot
    t = torch.tensor([float(v[:prt.n_owned] @ v[:prt.n_owned])], dtype=torch.float64)
    dist.all_reduce(t)
    ok_dot = bool(abs(t.item() - float(ref @ ref)) <= 1e-12 * float(ref @ ref))
    q.put((rank, ok_vec, ok_dot, prt.n_owned, prt.n_local))
    dist.destroy_process_group()


def test_two_rank_interface_exchange_over_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] and r[2] for r in res), res
    assert sum(r[3] for r in res) == (4 * 2 + 1) * (3 * 2 + 1) * (2 * 2 + 1)
