"""world_size-2 gloo test of the episode sharding + head-gradient all-reduce (host logic only;
the per-shard compute is a stand-in — here the oracle — injected by the test)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, B, q):
    for p in (ROOT, os.path.join(ROOT, "lite-mkd_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from lmkd.dist import shard_range, sharded_step
        from lmkd.episodes import make_episodes
        torch.manual_seed(0)
        ep = make_episodes(B, 3, 2, 2, 6, 32, teacher_dim=32, seed=3)
        d = 16
        W = {k: torch.randn(*s, generator=torch.Generator().manual_seed(i)).mul_(0.1).requires_grad_(True)
             for i, (k, s) in enumerate({"Wk": (d, 64), "bk": (d,), "Wv": (d, 64), "bv": (d,), "gk": (d,), "bek": (d,)}.items())}
        params = list(W.values())

        def compute(lo, hi):
            for p in params:
                p.grad = None
            tot, correct = 0.0, 0
            for b in range(lo, hi):
                lg = oracle.trx_logits(ep.support[b], ep.support_labels[b], ep.query[b], W["Wk"], W["bk"], W["Wv"],
                                       W["bv"], W["gk"] + 1.0, W["bek"], 2, 3)
                loss = oracle.cross_entropy(lg, ep.query_labels[b])
                loss.backward()
                tot += loss.item()
                correct += int((lg.argmax(1) == ep.query_labels[b]).sum())
            return tot, correct

        loss, correct, n = sharded_step(compute, B, params)
        grads = [p.grad.clone() for p in params]
        # single-process truth on every rank
        t_loss, t_correct = compute(0, B)
        ok = abs(float(loss) - t_loss) < 1e-6 * max(1.0, abs(t_loss)) and int(correct) == t_correct and int(n) == B
        for g, p in zip(grads, params):
            ok = ok and torch.allclose(g, p.grad, rtol=1e-5, atol=1e-7)
        lo, hi = shard_range(B, rank, world)
        q.put((rank, bool(ok), lo, hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [4, 5])
def test_world2_gloo_sharded_step_matches_single_process(B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + B
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(r[1] for r in res), res
    spans = sorted((r[2], r[3]) for r in res)
    assert spans[0][0] == 0 and spans[0][1] == spans[1][0] and spans[1][1] == B


def test_shard_range_partitions():
    from lmkd.dist import shard_range
    for total in (1, 7, 64, 4096):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker_views(rank, world, port, q):
    """Bucket-view gradients: backward ACCUMULATES straight into slices of the all-reduce bucket over several
    micro-batches, the all-reduce runs in place, nothing is copied (what bench.py does on the GPUs)."""
    for p in (ROOT, os.path.join(ROOT, "lite-mkd_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lmkd.dist import HeadGradReducer, shard_range
        g = torch.Generator().manual_seed(1)
        W = [torch.randn(5, 7, generator=g).requires_grad_(True), torch.randn(7, generator=g).requires_grad_(True)]
        X = torch.randn(8, 3, 5, generator=g)                 # 8 micro-batches of 3 rows
        red = HeadGradReducer(W, side_stream=False, grads_as_views=True)
        base = [p.grad.data_ptr() for p in W]
        lo, hi = shard_range(8, rank, world)
        for step in range(2):                                  # the views survive zero() and a second step
            red.zero()
            for m in range(lo, hi):
                ((X[m] @ W[0] + W[1]) ** 2).sum().backward()   # accumulates into the bucket slices
            red.reduce(0.0, 0, hi - lo)
            s = red.finish()
        ok = [p.grad.data_ptr() for p in W] == base and int(s[2]) == 8
        ref = [torch.zeros_like(p) for p in W]
        Wd = [p.detach().clone().requires_grad_(True) for p in W]
        for m in range(8):
            ((X[m] @ Wd[0] + Wd[1]) ** 2).sum().backward()
        for p, r in zip(W, Wd):
            ok = ok and torch.allclose(p.grad, r.grad, rtol=1e-5, atol=1e-6)
        ok = ok and red.bucket.data_ptr() == W[0].grad.data_ptr()
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_bucket_view_gradients_accumulate_and_reduce_in_place():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31700 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_views, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(r[1] for r in res), res
