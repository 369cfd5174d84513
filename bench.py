#!/usr/bin/env python
"""Benchmark of the Lite-MKD episodic matching + D2M distillation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[3], the configuration the metric's "1/2/4/8 B200" is quoted on; episodes shaped as
configs[1]): one optimizer step over a GLOBAL batch of 4096 5-way 5-shot episodes (25 queries, 8 frames x
2048-d, shuffled support labels), sharded over the ranks -- STRONG scaling, rank r owns 4096/N episodes and
walks them in 64-episode micro-batches.  Per micro-batch: TRX head with tuple cardinalities {2,3} (TrxBranch
semantics) student forward+backward, teacher forward (no grad) on the 3-modality teacher features, OTAM head
student forward+backward and teacher forward, SupportDK on both, the D2M losses (CE/16 + temperature-KL + 0.5 *
inter-class relation on the TRX logits; CE/16 + 2 * KL on the OTAM probabilities), backward to head parameters
AND student features.  Head gradients accumulate straight into the all-reduce bucket (p.grad are views of it);
per step ONE in-place NCCL all-reduce of that 94.4 MB bucket + 3 scalars, then Adam on the 23.6 M head parameters.

metric = episodes/sec (fwd+bwd matching + D2M loss), whole job.  `value`: episode tensors resident in HBM, every
micro-batch replayed from a CUDA graph (the NCCL call stays outside the graphs, so graphs work at any N).  `e2e`:
the same step fed from the HOST every micro-batch -- student features from pinned memory in bf16 (stand-in for
the on-GPU backbone's output), teacher features as row indices into the HBM-resident teacher-feature store,
labels -- with the loss read back every step.  `roofline` is for the tcgen05 kernels (every contraction of the
step incl. the fused attention kernel); `roofline_hbm_loss`, `roofline_tuple`, `roofline_otam_dp` carry the other
kernel classes, all timed with CUDA events inside this program.  `cpu_baseline` / `--impl reference` time the
CPU restatement of the reference (oracle/) on the host cores: the reference is pure Python and does not travel
to the GPU box.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "lite-mkd_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

WAY, SHOT, QPC, L, D, DOUT, CARDS = 5, 5, 5, 8, 2048, 1152, [2, 3]
NS, NQ = WAY * SHOT, WAY * QPC
CFG = dict(temperature=4, soft_loss_weight=2, hard_loss_weight=1, feature_loss_weight=1,
           soft_loss_weight_support=1, soft_loss_weight_query=1)
WORKLOAD = ("cfg4: 4096 episodes per optimizer step (5-way 5-shot, 25 queries, 8x2048-d, shuffled supports) sharded "
            "over the ranks in 64-episode micro-batches; TRX{2,3} + OTAM student fwd+bwd, teacher fwd, SupportDK, "
            "D2M losses, Adam")
METRIC = "episodes/sec (fwd+bwd matching+D2M loss)"
# class-centroid scale of the synthetic episodes: small enough that the random-initialised heads give soft
# predictions, i.e. non-zero loss gradients for the head parameters (see lmkd/episodes.py)
SEPARATION = 0.1


def head_args():
    return types.SimpleNamespace(seq_len=L, trans_dropout=0.1, trans_linear_out_dim=DOUT, trans_linear_in_dim=D,
                                 way=WAY, shot=SHOT, temp_set=CARDS)


def algorithmic_flops_per_episode(with_otam=True):
    """factored projection + class-grouped attention, student fwd+bwd = 3x, teacher fwd = 1x;
    frame similarity of the OTAM head 3 F_sim (student) + 1 F_sim (teacher)."""
    fwd = 0.0
    for c in CARDS:
        T = math.comb(L, c)
        fwd += 2.0 * (NS + NQ) * L * D * (2 * c * DOUT) + 4.0 * NQ * NS * T * T * DOUT
    f = 4.0 * fwd
    if with_otam:
        f += 4.0 * 2.0 * (NQ * L) * (NS * L) * D
    return f


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path, all host threads
# --------------------------------------------------------------------------------------------
def cpu_episode_step(ep, b, heads_s, heads_t, recipes):
    """Same work as one GPU episode: TRX{2,3} + OTAM student fwd+bwd, teacher fwd, SupportDK x2, the two losses."""
    import oracle
    s = ep.support[b].clone().requires_grad_(True)
    q = ep.query[b].clone().requires_grad_(True)
    for h in heads_s:
        for k in ("Wk", "bk", "Wv", "bv", "gk", "bek"):
            h[k].grad = None
    lab = ep.support_labels[b]
    lg = oracle.trx_branch_logits(s, lab, q, heads_s, WAY)
    sup = oracle.support_dk(s, WAY, SHOT, L)
    ot = oracle.otam_logits(s, lab, q)
    with torch.no_grad():
        tl = oracle.trx_branch_logits(ep.teacher_support[b], lab, ep.teacher_query[b], heads_t, WAY)
        tsup = oracle.support_dk(ep.teacher_support[b], WAY, SHOT, L)
        tot = oracle.otam_logits(ep.teacher_support[b], lab, ep.teacher_query[b])
    loss = recipes.fc_1_sup({"kl": lg, "sup": sup}, {"kl": tl, "sup": tsup}, ep.query_labels[b]) + \
        recipes.KD(ot, tot, ep.query_labels[b])
    loss.backward()
    return float(loss.detach())


def make_cpu_heads(seed):
    g = torch.Generator().manual_seed(seed)
    heads = []
    for c in CARDS:
        bound = 1.0 / math.sqrt(c * D)
        mk = lambda *s: (torch.rand(*s, generator=g) * 2 - 1).mul_(bound).requires_grad_(True)
        heads.append(dict(Wk=mk(DOUT, c * D), bk=mk(DOUT), Wv=mk(DOUT, c * D), bv=mk(DOUT),
                          gk=torch.ones(DOUT, requires_grad=True), bek=torch.zeros(DOUT, requires_grad=True), card=c))
    return heads


def run_cpu(n_warm, n_timed_min, budget_s):
    """cfg4-step episodes on the host cores.  Returns (episodes/s, episodes timed, cores, seconds)."""
    from oracle.losses import Recipes
    from lmkd.episodes import make_episodes
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ep = make_episodes(4, WAY, SHOT, QPC, L, D, modalities=3, separation=SEPARATION)
    hs, ht, rec = make_cpu_heads(1), make_cpu_heads(2), Recipes(CFG)
    for i in range(n_warm):
        cpu_episode_step(ep, i % 4, hs, ht, rec)
    t0, n = time.perf_counter(), 0
    while n < n_timed_min or (time.perf_counter() - t0 < budget_s and n < 64):
        cpu_episode_step(ep, n % 4, hs, ht, rec)
        n += 1
        if time.perf_counter() - t0 > 3 * budget_s:
            break
    dt = time.perf_counter() - t0
    return n / dt, n, cores, dt


def run_cpu_cfg1(n_warm=3, n_timed=20):
    """SURVEY.md §8(d) CPU baseline config: 5-way 1-shot, 25 queries, 8 x 512-d, OTAM head + Distiller.KD, fwd+bwd."""
    import oracle
    from oracle.losses import Recipes
    from lmkd.episodes import make_episodes
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ep = make_episodes(4, 5, 1, 5, 8, 512, teacher_dim=512, seed=11)
    rec = Recipes(CFG)
    tl = torch.randn(4, 25, 5, generator=torch.Generator().manual_seed(1))

    def one(b):
        s, q = ep.support[b].clone().requires_grad_(True), ep.query[b].clone().requires_grad_(True)
        rec.KD(oracle.otam_logits(s, ep.support_labels[b], q), tl[b], ep.query_labels[b]).backward()
    for i in range(n_warm):
        one(i % 4)
    t0 = time.perf_counter()
    for i in range(n_timed):
        one(i % 4)
    dt = time.perf_counter() - t0
    return {"value": n_timed / dt, "unit": "episodes/s", "cores": cores, "kind": "port",
            "sample": f"{n_timed} episodes of cfg1 (5-way 1-shot, 8x512-d, OTAM + Distiller.KD, fwd+bwd) in {dt:.2f} s "
                      f"after {n_warm} warm-ups (oracle port, torch {torch.__version__} CPU fp32)"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port), all host threads; each
    step is a bounded sample (2 episodes) of the same per-episode workload.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.losses import Recipes
    from lmkd.episodes import make_episodes
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    eps = 2
    ep = make_episodes(4, WAY, SHOT, QPC, L, D, modalities=3, separation=SEPARATION)
    hs, ht, rec = make_cpu_heads(1), make_cpu_heads(2), Recipes(CFG)
    for i in range(args.warmup):
        for j in range(eps):
            cpu_episode_step(ep, (i * eps + j) % 4, hs, ht, rec)
    t0 = time.perf_counter()
    for i in range(args.steps):
        for j in range(eps):
            cpu_episode_step(ep, (i * eps + j) % 4, hs, ht, rec)
    dt = time.perf_counter() - t0
    value = args.steps * eps / dt
    sample = (f"{eps} episodes per step x {args.steps} steps of the cfg4 per-episode workload "
              f"(oracle port, torch {torch.__version__} CPU fp32)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value,
        "unit": "episodes/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "episodes_per_step": eps, "way": WAY, "shot": SHOT, "frames": L, "dim": D,
                   "note": "bounded CPU sample of the per-episode work; episodes/s is per-episode comparable"},
        "cpu_baseline": {"value": value, "unit": "episodes/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "episodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def read_timing(lib, cat):
    ms, work, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
    lib.lmkd_kernel_timing_read(cat, ctypes.byref(ms), ctypes.byref(work), ctypes.byref(n))
    return ms.value, work.value, n.value


def run_ours(args):
    import torch.distributed as dist
    import distillers
    import model.classifiers as C
    from lmkd import _ffi, ops
    from lmkd.dist import HeadGradReducer, shard_range
    from lmkd.episodes import make_episodes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _ffi.lib()
    B = args.episodes                                   # micro-batch
    G = args.global_episodes
    lo, hi = shard_range(G // B, rank, world)           # micro-batches of this rank
    M = hi - lo
    assert G % B == 0 and M >= 1, "global batch must be a multiple of the micro-batch and give every rank work"
    peaks = load_peaks()

    torch.manual_seed(3483)
    student = C.TrxBranch(head_args()).to(dev).train()
    teacher = C.TrxBranch(head_args()).to(dev).train()      # the reference never calls teacher.eval()
    otam = C.OTAM(head_args())
    supdk = C.SupportDK(head_args())
    d_trx = distillers.Distiller("fc_1_sup", dict(CFG), dev)
    d_otam = distillers.Distiller("KD", dict(CFG), dev)
    params = [p for p in student.parameters() if p.requires_grad]
    # p.grad are views of ONE fp32 bucket: backward accumulates into it over the micro-batches, NCCL reduces it in
    # place, Adam reads it -- no gradient copies anywhere (round 1: 2 x 12 copy kernels + a copy-in/out per step)
    reducer = HeadGradReducer(params, side_stream=False, grads_as_views=True)
    ops.ACCUMULATE_PARAM_GRADS_IN_PLACE = not args.autograd_accumulate   # the backward kernels add into the bucket
    opt = torch.optim.Adam(params, lr=1e-4, fused=True)
    use_graph = not args.no_graph

    # two distinct resident micro-batches (each 4 x 105 MB of fp32 features > 126 MB L2), alternated
    batches = [make_episodes(B, WAY, SHOT, QPC, L, D, modalities=3, seed=3483 + 17 * rank + i, device=dev,
                             separation=SEPARATION) for i in range(2)]
    loss_acc = torch.zeros((), device=dev)

    def micro(ep, with_otam=True, accumulate_loss=True):
        """Forward + backward of one micro-batch; head gradients accumulate into the bucket views."""
        sup = ep.support.requires_grad_(True)
        qry = ep.query.requires_grad_(True)
        sup.grad = qry.grad = None
        lab = ep.support_labels
        lg = student(sup, lab, qry)["logits"]
        ssup = supdk(sup, lab, None)["logits"]
        with torch.no_grad():
            tl = teacher(ep.teacher_support, lab, ep.teacher_query)["logits"]
            tsup = supdk(ep.teacher_support, lab, None)["logits"]
        loss = d_trx.fc_1_sup({"kl": lg, "sup": ssup}, {"kl": tl, "sup": tsup}, ep.query_labels)["loss"]
        if with_otam:
            po = otam(sup, lab, qry)["logits"]
            with torch.no_grad():
                pt = otam(ep.teacher_support, lab, ep.teacher_query)["logits"]
            loss = loss + d_otam.KD(po, pt, ep.query_labels)["loss"]
        loss.backward()
        if os.environ.get("LMKD_BENCH_DEBUG") == "2":
            torch.cuda.synchronize()
            print("DEBUG micro: loss", float(loss), "requires_grad", loss.requires_grad, "sup.grad", float(sup.grad.norm()),
                  "Wk.grad", float(params[0].grad.norm()), "grad enabled", torch.is_grad_enabled(), "lg rg", lg.requires_grad,
                  file=sys.stderr)
        if accumulate_loss:
            loss_acc.add_(loss.detach())
        return loss.detach()

    class Graphed:
        """One micro-batch captured in a CUDA graph (all graphs share one memory pool: they replay back to back)."""
        pool = None

        def __init__(self, ep, **kw):
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    micro(ep, **kw)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            lib.lmkd_launch_count(1)
            self.graph = torch.cuda.CUDAGraph()
            if Graphed.pool is None:
                Graphed.pool = torch.cuda.graph_pool_handle()
            with torch.cuda.graph(self.graph, pool=Graphed.pool, capture_error_mode="thread_local"):
                self.loss = micro(ep, **kw)
            self.launches = int(lib.lmkd_launch_count(0))

        def __call__(self):
            self.graph.replay()
            return self.loss

    graph_note = "eager"
    runners = [lambda ep=ep: micro(ep) for ep in batches]
    launches_per_micro = None
    if use_graph:
        try:
            graphed = [Graphed(ep) for ep in batches]
            runners = graphed
            launches_per_micro = graphed[0].launches
            graph_note = ("every micro-batch replayed from a CUDA graph (one per resident batch, shared pool); "
                          "NCCL all-reduce and Adam outside the graphs")
        except Exception as ex:      # fail loudly in the output, keep measuring eagerly
            graph_note = f"eager (graph capture failed: {type(ex).__name__}: {str(ex)[:120]})"
            torch.cuda.synchronize()

    def global_step(run=None):
        """One optimizer step over this rank's share of the global batch; returns the device scalars
        [loss_sum, 0, episodes] of the WHOLE job."""
        run = run or (lambda m: runners[m % 2]())
        reducer.zero()
        loss_acc.zero_()
        for m in range(M):
            run(m)
        reducer.reduce(loss_acc, 0, M * B)
        s = reducer.finish()
        opt.step()
        return s

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- gradient equality across world sizes (SURVEY.md §8e "Verification") ------------------------------
    # V episodes with seeds tied to their GLOBAL micro-batch index, dropout off: the sharded, all-reduced head
    # gradient must equal the one a single rank computes over all V episodes.
    def grad_check():
        V = 8 * B
        student.eval()
        teacher.eval()       # the teacher's PE dropout would otherwise draw different targets in the two evaluations
        chunks = list(range(V // B))
        mine = [c for i, c in enumerate(chunks) if i % world == rank]

        def accumulate(which):
            reducer.zero()
            for c in which:
                micro(make_episodes(B, WAY, SHOT, QPC, L, D, modalities=3, seed=900000 + c, device=dev,
                                    separation=SEPARATION), accumulate_loss=False)
        accumulate(mine)
        if os.environ.get("LMKD_BENCH_DEBUG"):
            torch.cuda.synchronize()
            print("DEBUG after accumulate: bucket norm", float(reducer.bucket.norm()), "views",
                  all(p.grad is not None and p.grad.data_ptr() >= reducer.bucket.data_ptr() and
                      p.grad.data_ptr() < reducer.bucket.data_ptr() + reducer.nbytes for p in params),
                  "grad norms", [float(p.grad.norm()) for p in params[:3]], file=sys.stderr)
        if world > 1:
            dist.all_reduce(reducer.bucket, op=dist.ReduceOp.SUM)
        sharded = reducer.bucket.double().clone()
        out = {"episodes": V, "world": world, "l2": float(sharded.norm()), "sum": float(sharded.sum())}
        assert out["l2"] > 0.0, "head gradients are identically zero: the gradient check would be vacuous"
        if world > 1:
            if rank == 0:
                accumulate(chunks)          # the same V episodes on ONE rank
                single = reducer.bucket.double()
                out["single_rank_l2"] = float(single.norm())
                out["rel_l2_diff"] = float((sharded - single).norm() / single.norm())
                out["max_abs_diff_over_max_abs"] = float((sharded - single).abs().max() / single.abs().max())
                single = single.clone()
                accumulate(chunks)          # and once more: the run-to-run noise floor of the fp32 atomics
                again = reducer.bucket.double()
                out["single_rank_repeat_rel_l2_diff"] = float((again - single).norm() / single.norm())
            dist.barrier()
        student.train()
        teacher.train()
        reducer.zero()
        return out

    grad_note = grad_check()

    for i in range(args.warmup):
        global_step()
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    lib.lmkd_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        scal = global_step()
    e1.record()
    torch.cuda.synchronize()
    _ffi.check_device_status(dev)
    launches = int(lib.lmkd_launch_count(0)) if launches_per_micro is None else launches_per_micro * M * args.steps
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.summary() if sampler else None
    total_ms = float(ms.item())
    value = G * args.steps / (total_ms / 1e3)
    last_loss = float(scal[0].item()) / G

    # ---- e2e: every micro-batch fed from the host, loss read back every step ---------------------------------
    # student features: pinned host bf16 -> staging -> widened into the fp32 tensors the heads take;
    # teacher features: row indices into the HBM-resident teacher-feature store -> gather;
    # double-buffered on a copy stream so the upload of micro-batch m+1 overlaps the compute of micro-batch m.
    store = torch.cat([t.reshape(-1, L * D) for ep in batches for t in (ep.teacher_support, ep.teacher_query)]).bfloat16()
    host = []
    for j, ep in enumerate(batches):
        base = j * B * (NS + NQ)
        host.append(dict(
            sup=ep.support.detach().bfloat16().cpu().pin_memory(), qry=ep.query.detach().bfloat16().cpu().pin_memory(),
            slab=ep.support_labels.cpu().pin_memory(), qlab=ep.query_labels.cpu().pin_memory(),
            ts_idx=(base + torch.arange(B * NS)).reshape(B, NS).pin_memory(),
            tq_idx=(base + B * NS + torch.arange(B * NQ)).reshape(B, NQ).pin_memory()))
    stage = [dict(sup=torch.empty(B, NS, L, D, dtype=torch.bfloat16, device=dev),
                  qry=torch.empty(B, NQ, L, D, dtype=torch.bfloat16, device=dev),
                  ts_idx=torch.empty(B, NS, dtype=torch.int64, device=dev),
                  tq_idx=torch.empty(B, NQ, dtype=torch.int64, device=dev)) for _ in range(2)]
    h2d_micro = sum(t.numel() * t.element_size() for t in host[0].values())
    copy_stream = torch.cuda.Stream(device=dev)
    consumed = [None, None]           # event: the micro-batch that read static buffers j has finished

    def upload(m):
        j = m % 2
        with torch.cuda.stream(copy_stream):
            if consumed[j] is not None:
                copy_stream.wait_event(consumed[j])
            with torch.no_grad():
                for k in ("sup", "qry", "ts_idx", "tq_idx"):
                    stage[j][k].copy_(host[j][k], non_blocking=True)
                batches[j].support_labels.copy_(host[j]["slab"], non_blocking=True)
                batches[j].query_labels.copy_(host[j]["qlab"], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    def e2e_step():
        state = {"ev": upload(0)}

        def run(m):
            j = m % 2
            nxt = upload(m + 1) if m + 1 < M else None          # overlaps this micro-batch's compute
            torch.cuda.current_stream().wait_event(state["ev"])
            with torch.no_grad():
                ops.upcast_into(stage[j]["sup"], batches[j].support.detach())
                ops.upcast_into(stage[j]["qry"], batches[j].query.detach())
                ops.episode_gather(store, stage[j]["ts_idx"], L, out=batches[j].teacher_support)
                ops.episode_gather(store, stage[j]["tq_idx"], L, out=batches[j].teacher_query)
            runners[j]()
            done = torch.cuda.Event()
            done.record()
            consumed[j] = done
            state["ev"] = nxt
        s = global_step(run)
        return float(s[0].item())                                # device->host read of the step's loss

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(2):
        e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    _ffi.check_device_status(dev)
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3], device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = G * e2e_steps / (float(e2e_ms.item()) / 1e3)

    # ---- rooflines: per-launch CUDA-event timing inside the step, by kernel category ---------------------------
    lib.lmkd_gemm_timing_enable(1)
    nroof = 2
    for i in range(nroof):
        # park the GPU behind a spin kernel while the host enqueues the micro-batch, so the per-launch event
        # pairs bracket back-to-back kernels and never a host-side enqueue gap
        torch.cuda._sleep(int(1.0e8))
        micro(batches[i % 2], accumulate_loss=False)
    torch.cuda.synchronize()
    t_ms, t_flops, t_n = read_timing(lib, 0)
    u_ms, u_bytes, u_n = read_timing(lib, 1)
    o_ms, o_cells, o_n = read_timing(lib, 2)
    roofline = rf_tuple = rf_dp = rf_loss = extras = None
    micro_ms = total_ms / args.steps / M
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
    except Exception:
        pass
    if rank == 0:
        peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        src = "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"
        algo = algorithmic_flops_per_episode() * B * nroof
        achieved = algo / (t_ms / 1e3) / 1e12 if t_ms > 0 else 0.0
        roofline = {"bound": "tensor",
                    "kernel": "gemm_tcgen05_kernel + trx_attn_fwd_kernel (every tcgen05 contraction of the micro-batch)",
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": traffic.get("trx_attn_fwd_kernel_c3_train_bytes"),
                    "traffic_of": "one training launch of trx_attn_fwd_kernel<18> (c = 3, 64 episodes): the longest single "
                                  "tcgen05 launch; per-kernel DRAM bytes of the whole pass are in profiles/r02_ncu_traffic.json",
                    "dram_bytes_student_pass": (traffic.get("student_trx_pass") or {}).get("dram_bytes"),
                    "peak_source": f"{src} bf16_tflops_sustained",
                    "launches_per_micro_batch": t_n // nroof, "kernel_ms_per_micro_batch": t_ms / nroof,
                    "executed_tflops": t_flops / (t_ms / 1e3) / 1e12 if t_ms > 0 else 0.0,
                    "share_of_micro_batch": (t_ms / nroof) / micro_ms,
                    "algorithmic_gflop_per_episode": algorithmic_flops_per_episode() / 1e9,
                    "whole_step_frac": algorithmic_flops_per_episode() * G / (total_ms / args.steps / 1e3) / 1e12 / peak / world}
        rf_tuple = {"bound": "hbm", "kernel": "tuple_ln_fwd2_kernel + ln_gather_bwd3_kernel (tuple assembly + LayerNorm, fwd and bwd)",
                    "achieved": u_bytes / (u_ms / 1e3) / 1e9 if u_ms > 0 else 0.0, "peak": hbm, "unit": "GB/s",
                    "frac": (u_bytes / (u_ms / 1e3) / 1e9 / hbm) if u_ms > 0 else 0.0,
                    "traffic": traffic.get("tuple_kernels_bytes_per_micro_batch"), "peak_source": f"{src} hbm_gbs",
                    "launches_per_micro_batch": u_n // nroof, "kernel_ms_per_micro_batch": u_ms / nroof,
                    "share_of_micro_batch": (u_ms / nroof) / micro_ms}
        rf_dp = {"bound": "issue (neither roofline: transcendental recurrence)", "kernel": "otam_dp_fwd_kernel + otam_dp_bwd_kernel",
                 "achieved": o_cells / (o_ms / 1e3) / 1e9 if o_ms > 0 else 0.0, "unit": "G cells/s", "peak": None, "frac": None,
                 "launches_per_micro_batch": o_n // nroof, "kernel_ms_per_micro_batch": o_ms / nroof,
                 "share_of_micro_batch": (o_ms / nroof) / micro_ms,
                 "occupancy": traffic.get("otam_dp_occupancy")}

    # ---- the other kernel classes at their own BASELINE configs (rank 0, a few hundred ms each) -----------------
    if rank == 0 and not args.no_extras:
        extras = {}
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def timed(fn, iters=5, warm=2):
            for _ in range(warm):
                fn()
            ts = []
            for _ in range(iters):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            return sorted(ts)[len(ts) // 2]

        # cfg3: fused D2M feature-MSE fwd+bwd over 1024 episodes (kernel time from the library's own event pairs)
        n = 1024 * (NS + NQ) * L * D
        s_, t_ = torch.randn(n, device=dev), torch.randn(n, device=dev)
        ds = torch.empty_like(s_)
        part = torch.empty(lib.lmkd_mse_partials(), dtype=torch.float32, device=dev)
        lo_ = torch.zeros(1, device=dev)
        read_timing(lib, 3)
        for _ in range(7):
            flush.zero_()
            _ffi.check(lib.lmkd_d2m_feature_mse_fwdbwd(_ffi.ptr(s_), _ffi.ptr(t_), _ffi.ptr(ds), n, 0, 1.0 / n, 2.0 / n,
                                                       _ffi.ptr(part), _ffi.ptr(lo_), 0, _ffi.stream()))
        torch.cuda.synchronize()
        l_ms, l_bytes, l_n = read_timing(lib, 3)
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        gbs = l_bytes / (l_ms / 1e3) / 1e9
        rf_loss = {"bound": "hbm", "kernel": "feat_mse_kernel (fused D2M feature-MSE fwd+bwd, cfg3: 1024 episodes, fp32)",
                   "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                   "traffic": traffic.get("feat_mse_kernel_bytes"), "algorithmic_bytes_per_launch": l_bytes / l_n,
                   "ms_per_launch": l_ms / l_n, "episodes_per_s": 1024 / (l_ms / l_n / 1e3),
                   "peak_source": ("MEASURED_PEAKS.json" if peaks else "fallback") + " hbm_gbs", "l2": "flushed between launches"}
        del s_, t_, ds
        # cfg4: OTAM alone over 4096 episodes
        ep4 = make_episodes(4096, WAY, SHOT, QPC, L, D, teacher_dim=8, device=dev)
        s4, q4 = ep4.support.requires_grad_(True), ep4.query.requires_grad_(True)
        up4 = torch.randn(4096, NQ, WAY, device=dev)

        def otam_fb():
            s4.grad = q4.grad = None
            (ops.otam_probs(s4, ep4.support_labels, q4, WAY) * up4).sum().backward()
        ms_f = timed(lambda: ops.otam_probs(s4.detach(), ep4.support_labels, q4.detach(), WAY), iters=3, warm=1)
        read_timing(lib, 2)
        ms_fb = timed(otam_fb, iters=3, warm=1)
        d_ms, d_cells, d_n = read_timing(lib, 2)
        extras["otam_cfg4_4096_episodes"] = {"fwd_ms": ms_f, "fwd_bwd_ms": ms_fb, "episodes_per_s_fwd_bwd": 4096 / (ms_fb / 1e3),
                                             "dp_kernels_G_cells_per_s": d_cells / (d_ms / 1e3) / 1e9 if d_ms > 0 else None,
                                             "dp_kernels_ms_per_fwd_bwd": d_ms / 4 if d_n else None}
        del ep4, s4, q4, up4
        lib.lmkd_gemm_timing_enable(0)
        # continuity with round 1: the cfg2 micro-batch (TRX{2,3} only, 64 episodes), graph replay
        try:
            g2 = [Graphed(ep, with_otam=False, accumulate_loss=False) for ep in batches] if use_graph else None
            r2 = g2 if g2 else [lambda ep=ep: micro(ep, with_otam=False, accumulate_loss=False) for ep in batches]
            for i in range(4):
                r2[i % 2]()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(10):
                r2[i % 2]()
            b.record()
            torch.cuda.synchronize()
            ms2 = a.elapsed_time(b) / 10
            extras["cfg2_trx_only_64_episodes"] = {"ms_per_micro_batch": ms2, "episodes_per_s": B / (ms2 / 1e3),
                                                   "note": "BASELINE configs[1] as benched in round 1 (no OTAM, no optimizer)"}
        except Exception as ex:
            extras["cfg2_trx_only_64_episodes"] = {"error": f"{type(ex).__name__}: {str(ex)[:160]}"}
        # the shipped recipe: Student TRX_2fcsup (two feature heads) + Teacher TRX_2fcsup_fixed + fc_2_sup_dist
        try:
            a2 = head_args()
            stu2 = C.TRX_2fcsup(a2).to(dev).train()
            tea2 = C.TRX_2fcsup_fixed(a2).to(dev).train()
            d2 = distillers.Distiller("fc_2_sup_dist", dict(CFG), dev)
            ep = batches[0]
            f2 = ep.support.detach().clone().requires_grad_(True)
            q2 = ep.query.detach().clone().requires_grad_(True)

            def shipped():
                sup = ep.support.requires_grad_(True)
                qry = ep.query.requires_grad_(True)
                lg = stu2({"context_features_1": sup, "context_features_2": f2}, ep.support_labels,
                          {"target_features_1": qry, "target_features_2": q2})["logits"]
                tl = tea2(ep.teacher_support, ep.support_labels, ep.teacher_query)["logits"]
                d2.fc_2_sup_dist(lg, tl, ep.query_labels)["loss"].backward()
            ms3 = timed(shipped, iters=5, warm=2)
            extras["shipped_recipe_fc_2_sup_dist_shuffled_supports"] = {
                "ms_per_micro_batch": ms3, "episodes_per_s": B / (ms3 / 1e3),
                "note": "TRX_2fcsup student (two c=2 passes) + TRX_2fcsup_fixed teacher + SupportDK, eager launches"}
        except Exception as ex:
            extras["shipped_recipe_fc_2_sup_dist_shuffled_supports"] = {"error": f"{type(ex).__name__}: {str(ex)[:160]}"}
    lib.lmkd_gemm_timing_enable(0)

    cpu = cpu1 = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, n, cores, dt = run_cpu(1, 3, 12.0)
        cpu = {"value": v, "unit": "episodes/s", "cores": cores, "kind": "port",
               "sample": f"{n} episodes of the cfg4 per-episode workload in {dt:.1f} s (oracle port of the reference, torch CPU fp32)"}
        cpu1 = run_cpu_cfg1()

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "episodes/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_episodes_per_step": G, "episodes_per_micro_batch": B,
                       "micro_batches_per_rank_per_step": M, "ms_per_micro_batch": micro_ms,
                       "way": WAY, "shot": SHOT, "queries": NQ, "frames": L, "dim": D, "key_dim": DOUT,
                       "cardinalities": CARDS, "parallelism": f"episode-sharded x{world}, one in-place NCCL all-reduce of the 94.4 MB head-gradient bucket per step",
                       "l2_policy": "inputs (420 MB of features per micro-batch, 2 alternating batches) exceed the 126 MB L2",
                       "optimizer": "Adam(fused) on the 23.6 M head parameters inside the timed region, once per global step",
                       "launch": graph_note, "mean_loss_per_episode": last_loss, "class_separation": SEPARATION},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "episodes/s", "h2d_bytes_per_step": h2d_micro * M * world,
                    "d2h_bytes_per_step": 4 * world, "h2d_bytes_per_micro_batch_per_rank": h2d_micro, "steps": e2e_steps,
                    "inputs": "student features pinned-host bf16 (stand-in for the on-GPU backbone output), teacher features "
                              "as row indices into the HBM-resident bf16 feature store, labels; loss read back per step"},
            "gpu_launches": launches,
            "roofline": roofline, "roofline_hbm_loss": rf_loss, "roofline_tuple": rf_tuple, "roofline_otam_dp": rf_dp,
            "grad_check": grad_note,
            "extras": extras,
            "cpu_baseline": cpu, "cpu_baseline_cfg1": cpu1,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        # leave together, then exit hard: NCCL teardown must never keep a rank (and torchrun) alive
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--episodes", type=int, default=64, help="episodes per micro-batch")
    ap.add_argument("--global-episodes", type=int, default=4096, help="episodes per optimizer step, whole job")
    ap.add_argument("--e2e-steps", type=int, default=8, help="upper bound of the timed end-to-end steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the per-kernel-class measurements at their own configs")
    ap.add_argument("--autograd-accumulate", action="store_true",
                    help="let autograd add the head gradients (12 extra elementwise kernels per micro-batch) instead of "
                         "the backward kernels accumulating into the bucket")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every micro-batch eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3      # timing hygiene: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
