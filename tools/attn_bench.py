"""Times the fused TRX attention kernel alone (lmkd_trx_attn_fwd) on the config-2 episode shape.
usage: python tools/attn_bench.py [B=64] [card=3] [train=1] [iters=20]"""
import ctypes as C
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lite-mkd_b200")]
import torch
from lmkd._ffi import TrxShape, check, lib, ptr, stream

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
card = int(sys.argv[2]) if len(sys.argv) > 2 else 3
train = int(sys.argv[3]) if len(sys.argv) > 3 else 1
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 20
way, shot, Nq, L, d = 5, 5, 25, 8, 1152
dev = torch.device("cuda:0")
T = math.comb(L, card)
KTp = (shot * T + 15) // 16 * 16
NqT = Nq * T
sh = TrxShape(B, way * shot, Nq, L, 64, d, card, way, shot, 0.0, 0, None, 1e-5)
g = torch.Generator(device="cpu").manual_seed(1)
kq = torch.randn(B, NqT, d, device=dev).bfloat16()
vq = torch.randn(B, NqT, d, device=dev).bfloat16()
ks = torch.randn(B, way, KTp, d, device=dev).bfloat16()
vs = torch.randn(B, way, KTp, d, device=dev).bfloat16()
ks[:, :, shot * T:] = 0
vs[:, :, shot * T:] = 0
cnt = torch.full((B, way), shot, dtype=torch.int32, device=dev)
dq = torch.zeros(B, way, NqT, d, device=dev).bfloat16() if train else None
patt = torch.zeros(B, NqT, way * KTp, device=dev).bfloat16() if train else None
rowred = torch.zeros(B, way, NqT, device=dev)
rowdot = torch.zeros(B, way, NqT, device=dev) if train else None
linv = torch.zeros(B, way, NqT, device=dev) if train else None


def call():
    check(lib().lmkd_trx_attn_fwd(C.byref(sh), ptr(kq), ptr(vq), ptr(ks), ptr(vs), ptr(cnt), ptr(dq), ptr(patt),
                                  ptr(rowred), ptr(rowdot), ptr(linv), stream()), "attn")


for _ in range(3):
    call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    call()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
flops = 4.0 * B * way * NqT * (shot * T) * d
print(f"attn B={B} card={card} train={train}: {ms * 1e3:.1f} us  {flops / ms / 1e9:.0f} TFLOP/s (algorithmic, valid columns)")
