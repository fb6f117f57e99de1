"""world_size-2 `gloo` test (CPU) of the N > 1 host logic: index sharding, RNG keying by global env
index, and the gradient / normalisation decomposition the multi-GPU path relies on — local sums
scaled by 1/(B_local * G), one sum all-reduce — checked with the oracle against the single-process
result on the concatenated batch."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as g
    import oracle_lib as O
    O.lib().orc_set_threads(2)
    pkg = g.load_package()
    sh = pkg.sharding
    n_total, T = 256, 8
    lo, hi = sh.shard_range(n_total, rank, world)
    seeds = sh.splitmix_states(0xABCDEF, lo, hi)
    # (1) shards tile the single-process seeding exactly
    full = sh.splitmix_states(0xABCDEF, 0, n_total)
    ok_seeds = np.array_equal(seeds, full[lo:hi]) and np.array_equal(full, O.splitmix_states_fast(n_total, 0xABCDEF))
    # (2) synthetic rollout on the shard; global quantities through gloo all-reduces
    desc = O.ac_desc(4, 64, 2)
    params = sh.glorot_actor_critic(123, 4, 64, 2)
    rng = np.random.default_rng(7)  # same stream on both ranks, each takes its slice -> identical global data
    nt = n_total * T
    states = rng.standard_normal((4, n_total, T)).astype(np.float32)
    actions = rng.integers(1, 3, (n_total, T)).astype(np.int32)
    logp = (-0.7 + 0.1 * rng.standard_normal((n_total, T))).astype(np.float32)
    adv = rng.standard_normal((n_total, T)).astype(np.float32); ret = rng.standard_normal((n_total, T)).astype(np.float32)
    flat = lambda a, sl: np.ascontiguousarray(a[..., sl, :].reshape(*a.shape[:-2], -1, order="F") if a.ndim == 3 else a[sl].ravel(order="F"))
    sl = slice(lo, hi)
    a_l = adv[sl].ravel(order="F")
    sums = torch.tensor([a_l.astype(np.float64).sum(), (a_l.astype(np.float64) ** 2).sum()], dtype=torch.float64)
    dist.all_reduce(sums)
    mean = sums[0].item() / nt
    var = (sums[1].item() - nt * mean * mean) / (nt - 1)
    mean_f, inv_std_f = float(np.float32(mean)), float(np.float32(1.0) / np.float32(np.sqrt(var)))
    g_l, l_l = O.ac_loss_grad(0, desc, O.hyper_array(), params, np.asfortranarray(states[:, sl, :]).reshape(4, -1, order="F"),
                              actions[sl].ravel(order="F"), logp[sl].ravel(order="F"), a_l, ret[sl].ravel(order="F"), None, mean_f, inv_std_f)
    gt = torch.tensor(g_l / world)           # local mean * 1/G == local sum * 1/(B_local * G)
    dist.all_reduce(gt)
    lt = torch.tensor([l_l["actor_loss"] / world, l_l["critic_loss"] / world, l_l["entropy"] / world], dtype=torch.float64)
    dist.all_reduce(lt)
    if rank == 0:
        m_ref, s_ref = O.adv_norm(adv.ravel(order="F"))
        g_ref, l_ref = O.ac_loss_grad(0, desc, O.hyper_array(), params, np.asfortranarray(states).reshape(4, -1, order="F"),
                                      actions.ravel(order="F"), logp.ravel(order="F"), adv.ravel(order="F"), ret.ravel(order="F"), None, m_ref, s_ref)
        q.put(dict(ok_seeds=ok_seeds, norm=(abs(mean_f - m_ref) < 1e-6 and abs(inv_std_f - s_ref) < 1e-5 * s_ref),
                   grad_err=float(np.linalg.norm(gt.numpy() - g_ref) / np.linalg.norm(g_ref)),
                   loss_err=float(abs(lt[1].item() - l_ref["critic_loss"]) / l_ref["critic_loss"])))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_decomposition_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["ok_seeds"] and res["norm"]
    assert res["grad_err"] < 1e-6 and res["loss_err"] < 1e-6


def test_shard_range_rules():
    import __graft_entry__ as g
    sh = g.load_package().sharding
    assert [sh.shard_range(65536, r, 8) for r in (0, 7)] == [(0, 8192), (57344, 65536)]
    with pytest.raises(ValueError):
        sh.shard_range(10, 0, 3)
