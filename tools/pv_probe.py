"""PV-shaped contraction (short K, MN-major B) timed across column-tile widths: is it bound by operand bytes
per tile (time ~ bytes) or by per-tile latency (time ~ tiles)?   python tools/pv_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lite-mkd_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

from lmkd import ops  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=5):
    ts = []
    for i in range(iters + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


for name, (M, N, K, nb) in {"PV_c3": (1400, 1152, 288, 320), "PV_c2": (700, 1152, 144, 320),
                            "PV_c3_K320": (1400, 1152, 320, 320), "PV_c3_K416": (1400, 1152, 416, 320)}.items():
    A = torch.randn(nb, M, K, device=dev).bfloat16()
    B = torch.randn(nb, K, N, device=dev).bfloat16()
    C = torch.empty(nb, M, N, device=dev)
    for bn in (128, 192, 256):
        ms = timed(lambda: ops.gemm_bf16(A, B, b_mn=True, out=C, block_n=bn))
        tiles = -(-M // 128) * -(-N // bn) * nb
        kb = -(-K // 64)
        kbytes = kb * (16 + bn // 64 * 8) + 128 * bn * 4 / 1024
        print(f"{name:12s} BN {bn:3d}  {ms:7.3f} ms  {2.0 * M * N * K * nb / ms / 1e9:7.1f} TFLOP/s  "
              f"{ms * 1e3 / (tiles / 148):6.2f} us/tile/SM  {kbytes:6.0f} KB/tile  {kbytes / (ms * 1e3 / (tiles / 148)):6.1f} GB/s/SM")
    del A, B, C
