"""Per-kernel roofline measurements on one B200 (CUDA events, L2 flushed between iterations):

  * fused D2M feature-MSE fwd+bwd at BASELINE config 3 size (1024 episodes) -> GB/s vs measured HBM peak
  * OTAM at config 4 size (4096 episodes, 5-way 5-shot): forward / forward+backward, DP cells/s
  * the tcgen05 GEMM on each contraction shape of config 2 -> TFLOP/s vs measured bf16 peak
  * full KL_feature step at config 3 (TRX{2} head + feature MSE), episodes/s

Prints one JSON object; `python tools/kernel_bench.py > profiles/<name>.json`.
"""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lite-mkd_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

from lmkd import _ffi, ops  # noqa: E402


def timed(fn, iters=5, warm=2, flush=None):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()                       # > L2: evicts the previous iteration's data
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def gemm_section(out, dev, flush, tf_sus, tf_burst):
    # ---- GEMM shapes of config 2 (B = 64) -----------------------------------------------------------
    shapes = {
        "proj_c2   X~.Wcat^T   M25600 N4608 K2048": (25600, 4608, 2048, 1, 0, 0),
        "proj_c3   X~.Wcat^T   M25600 N6912 K2048": (25600, 6912, 2048, 1, 0, 0),
        "dX_c3     dP.Wcat     M25600 N2048 K6912": (25600, 2048, 6912, 1, 0, 1),
        "dW_c3     dP^T.X~     M6912 N2048 K25600": (6912, 2048, 25600, 1, 1, 1),
        "scores_c2 Kq.Ks^T     M700 N720 K1152 x64": (700, 720, 1152, 64, 0, 0),
        "scores_c3 Kq.Ks^T     M1400 N1440 K1152 x64": (1400, 1440, 1152, 64, 0, 0),
        "PV_c3     P.V         M1400 N1152 K288 x320": (1400, 1152, 288, 320, 0, 1),
        "dVs_c3    P^T.D       M288 N1152 K1400 x320": (288, 1152, 1400, 320, 1, 1),
        "dKs_c3    dS^T.Kq     M1440 N1152 K1400 x64": (1440, 1152, 1400, 64, 1, 1),
        "sim_cfg4  Xq.Xs^T     M200 N200 K2048 x4096": (200, 200, 2048, 4096, 0, 0),
    }
    g = {}
    for name, (M, N, K, nb, amn, bmn) in shapes.items():
        A = torch.randn((nb, K, M) if amn else (nb, M, K), device=dev).bfloat16()
        Bm = torch.randn((nb, K, N) if bmn else (nb, N, K), device=dev).bfloat16()
        C = torch.empty(nb, M, N, device=dev)
        ms = timed(lambda: ops.gemm_bf16(A, Bm, a_mn=bool(amn), b_mn=bool(bmn), out=C), iters=5, flush=flush)
        tf = 2.0 * M * N * K * nb / (ms / 1e3) / 1e12
        g[name] = {"ms": ms, "TFLOPs": tf, "frac_of_sustained_peak": tf / tf_sus, "frac_of_burst_peak": tf / tf_burst}
        del A, Bm, C
    out["gemm_tcgen05_store_f32_epilogue"] = g



def main():
    only = sys.argv[1] if len(sys.argv) > 1 else "all"
    dev = torch.device("cuda:0")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    tf_burst = float(peaks.get("bf16_tflops", 1590.0))
    tf_sus = float(peaks.get("bf16_tflops_sustained", 1400.0))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {"peaks": {"hbm_gbs": hbm, "bf16_tflops_burst": tf_burst, "bf16_tflops_sustained": tf_sus,
                     "source": "MEASURED_PEAKS.json" if peaks else "fallback"}}
    lib = _ffi.lib()

    # ---- fused feature MSE (config 3: 1024 episodes x 50 videos x 8 frames x 2048) ----------------
    res = {}
    for name, dtype, esz in ((("fp32", torch.float32, 4), ("bf16_storage", torch.bfloat16, 2)) if only == "all" else ()):
        n = 1024 * 50 * 8 * 2048
        s = torch.randn(n, device=dev, dtype=dtype)
        t = torch.randn(n, device=dev, dtype=dtype)
        ds = torch.empty_like(s)
        partials = torch.empty(lib.lmkd_mse_partials(), dtype=torch.float32, device=dev)
        loss = torch.zeros(1, device=dev)
        st = _ffi.stream()
        fn = lambda: _ffi.check(lib.lmkd_d2m_feature_mse_fwdbwd(_ffi.ptr(s), _ffi.ptr(t), _ffi.ptr(ds), n, 0 if esz == 4 else 1,
                                                               1.0 / n, 2.0 / n, _ffi.ptr(partials), _ffi.ptr(loss), 0, st))
        ms = timed(fn, iters=7)
        gbs = 3.0 * n * esz / (ms / 1e3) / 1e9
        res[name] = {"ms": ms, "algorithmic_bytes": 3 * n * esz, "GBps": gbs, "frac_of_measured_hbm": gbs / hbm,
                     "episodes_per_s": 1024 / (ms / 1e3)}
        del s, t, ds
    out["feature_mse_cfg3_1024_episodes"] = res

    # ---- student feature heads at config 2: 64 episodes x 50 videos x 8 frames ----
    if only in ("all", "heads"):
        rows = 64 * 50 * 8
        fmap = torch.randn(rows, 512, 7, 7, device=dev)
        x = torch.randn(rows, 512, device=dev).requires_grad_(True)
        Wt = (torch.randn(2, 2048, 512, device=dev) * 0.04).requires_grad_(True)
        bt = torch.zeros(2, 2048, device=dev).requires_grad_(True)
        up = torch.randn(2, rows, 2048, device=dev)
        ms_pool = timed(lambda: ops.frame_pool(fmap, 4), iters=5)

        def fh_fb():
            x.grad = Wt.grad = bt.grad = None
            (ops.feature_heads(x, Wt, bt) * up).sum().backward()
        ms_f = timed(lambda: ops.feature_heads(x.detach(), Wt.detach(), bt.detach()), iters=5)
        ms_fb = timed(fh_fb, iters=5)
        fl = 2.0 * rows * 512 * 2048 * 2
        out["feature_heads_cfg2_64_episodes"] = {
            "pool_ms": ms_pool, "pool_GBps": (fmap.numel() + rows * 512) * 4 / (ms_pool / 1e3) / 1e9,
            "pool_frac_of_measured_hbm": (fmap.numel() + rows * 512) * 4 / (ms_pool / 1e3) / 1e9 / hbm,
            "fc_fwd_ms": ms_f, "fc_fwd_TFLOPs": fl / (ms_f / 1e3) / 1e12,
            "fc_fwd_output_GBps": 2 * rows * 2048 * 4 / (ms_f / 1e3) / 1e9,
            "fc_fwd_bwd_ms_incl_autograd_mul": ms_fb, "episodes_per_s_fwd_bwd": 64 / (ms_fb / 1e3)}
        del fmap, x, Wt, bt, up
        if only == "heads":
            print(json.dumps(out, indent=1))
            return

    # ---- teacher-feature store at config 3: 1024 episodes x 50 videos -------------
    if only in ("all", "store"):
        res = {}
        nvid, row, Bst = 20000, 8 * 2048, 1024
        idx = torch.randint(0, nvid, (Bst, 50), device=dev)
        student = torch.randn(Bst, 50, 8, 2048, device=dev)
        n = student.numel()
        for name, dt, esz in (("fp32_store", torch.float32, 4), ("bf16_store", torch.bfloat16, 2)):
            store = torch.randn(nvid, row, device=dev).to(dt)
            ms_g = timed(lambda: ops.episode_gather(store, idx, 8), iters=5)
            s_ = student.requires_grad_(True)
            ms_l = timed(lambda: ops.feature_mse_from_store(s_, store, idx, 1.0, 50 * row), iters=5)
            res[name] = {"store_GB": store.numel() * esz / 1e9,
                         "gather_ms": ms_g, "gather_GBps": n * (esz + 4) / (ms_g / 1e3) / 1e9,
                         "gather_frac_of_measured_hbm": n * (esz + 4) / (ms_g / 1e3) / 1e9 / hbm,
                         "store_fed_mse_ms": ms_l, "store_fed_mse_GBps": n * (8 + esz) / (ms_l / 1e3) / 1e9,
                         "store_fed_mse_frac_of_measured_hbm": n * (8 + esz) / (ms_l / 1e3) / 1e9 / hbm,
                         "episodes_per_s": Bst / (ms_l / 1e3)}
            del store
        out["teacher_feature_store_cfg3_1024_episodes"] = res
        del student, idx
        if only == "store":
            print(json.dumps(out, indent=1))
            return

    # ---- OTAM at config 4 (4096 episodes, 5-way 5-shot, 25 queries, L=8, D=2048) -----------------
    from lmkd.episodes import make_episodes
    if only == "gemm":
        gemm_section(out, dev, flush, tf_sus, tf_burst)
        print(json.dumps(out, indent=1))
        return
    B = 4096
    ep = make_episodes(B, 5, 5, 5, 8, 2048, teacher_dim=8, device=dev)      # teacher feats unused here
    sup, qry = ep.support.requires_grad_(True), ep.query.requires_grad_(True)
    up = torch.randn(B, 25, 5, device=dev)
    fwd = lambda: ops.otam_probs(sup.detach(), ep.support_labels, qry.detach(), 5)

    def fwdbwd():
        sup.grad = qry.grad = None
        (ops.otam_probs(sup, ep.support_labels, qry, 5) * up).sum().backward()
    ms_f, ms_fb = timed(fwd, iters=3, warm=1), timed(fwdbwd, iters=3, warm=1)
    cells = 2 * 25 * 25 * 8 * 9 * B
    out["otam_cfg4_4096_episodes"] = {
        "fwd_ms": ms_f, "fwd_bwd_ms": ms_fb, "episodes_per_s_fwd_bwd": B / (ms_fb / 1e3),
        "dp_cells_fwd": cells, "sim_gemm_gflop_fwd_bwd": 3 * 2 * 200 * 200 * 2048 * B / 1e9}
    # DP kernels alone, through the raw recurrence entry point on the same number of tables
    d = torch.rand(B * 25 * 25, 8, 8, device=dev)
    go = torch.ones(B * 25 * 25, device=dev)
    ms_dp = timed(lambda: ops.otam_cum_dist(d, 0.1), iters=3, warm=1)
    ms_dpb = timed(lambda: ops.otam_cum_dist(d, 0.1, grad_out=go), iters=3, warm=1)
    out["otam_cfg4_4096_episodes"].update({"dp_one_direction_fwd_ms": ms_dp, "dp_one_direction_fwd_bwd_ms": ms_dpb,
                                           "dp_cells_per_s_fwd": (cells / 2) / (ms_dp / 1e3)})
    del ep, sup, qry, d, go
    if only == "otam":
        print(json.dumps(out, indent=1))
        return

    gemm_section(out, dev, flush, tf_sus, tf_burst)

    # ---- config 3 step: TRX{2} student fwd+bwd + teacher fwd + KL_feature (CE/16 + 2 T^2 KL + MSE) ----
    import distillers
    import model.classifiers as Cm
    args = types.SimpleNamespace(seq_len=8, trans_dropout=0.1, trans_linear_out_dim=1152, trans_linear_in_dim=2048,
                                 way=5, shot=5, temp_set=[2])
    cfg = dict(temperature=4, soft_loss_weight=2, hard_loss_weight=1, feature_loss_weight=1,
               soft_loss_weight_support=1, soft_loss_weight_query=1)
    B3 = 128
    ep = make_episodes(B3, 5, 5, 5, 8, 2048, modalities=3, device=dev)
    student, teacher = Cm.TRX(args).to(dev).train(), Cm.TRX_fixed(args).to(dev).train()
    dist = distillers.Distiller("KL_feature", cfg, dev)

    def step3():
        sup, qry = ep.support.requires_grad_(True), ep.query.requires_grad_(True)
        sup.grad = qry.grad = None
        lg = student(sup, ep.support_labels, qry)["logits"]
        tl = teacher(ep.teacher_support, ep.support_labels, ep.teacher_query)["logits"]
        sf = torch.cat([sup, qry], 1)
        tf_ = torch.cat([ep.teacher_support, ep.teacher_query], 1)
        dist.KL_feature({"logits": lg, "feature": sf}, {"logits": tl, "feature": tf_}, ep.query_labels)["loss"].backward()
    ms3 = timed(step3, iters=3, warm=2)
    out["cfg3_step_KL_feature_TRX2"] = {"episodes_per_step": B3, "ms_per_step": ms3, "episodes_per_s": B3 / (ms3 / 1e3),
                                        "note": "1024-episode config run as 8 steps of 128 (workspace of the head)"}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
