"""Bring-up / regression check of the tcgen05 GEMM against torch.matmul (fp32 on bf16-rounded inputs).
Each case runs in its own process (a device trap must not poison the other cases):
    python tools/gemm_check.py            # all cases, one subprocess each
    python tools/gemm_check.py <index>    # one case in-process
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lite-mkd_b200"))

# (M, N, K, batch, a_mn, b_mn, block_n)
CASES = [
    (128, 128, 64, 1, 0, 0, 128),
    (128, 128, 256, 1, 0, 0, 128),
    (256, 256, 512, 1, 0, 0, 128),
    (200, 40, 512, 1, 0, 0, 0),
    (700, 720, 1152, 2, 0, 0, 0),
    (1000, 512, 200, 3, 0, 0, 256),
    (128, 128, 64, 1, 0, 1, 128),
    (128, 128, 64, 1, 1, 0, 128),
    (128, 128, 128, 1, 1, 1, 128),
    (700, 1152, 144, 5, 0, 1, 0),
    (144, 1152, 700, 5, 1, 1, 0),
    (720, 1152, 700, 2, 1, 1, 0),
    (4608, 2048, 3200, 1, 1, 1, 0),
    (3200, 4608, 2048, 1, 0, 0, 0),
    (3200, 2048, 4608, 1, 0, 1, 0),
    (200, 512, 40, 4, 0, 1, 0),
    (40, 512, 200, 4, 1, 1, 0),
    (50, 16, 24, 2, 0, 0, 0),
    (50, 16, 24, 2, 1, 1, 0),
]


def run_case(i):
    import torch
    from lmkd import ops
    M, N, K, nb, a_mn, b_mn, bn = CASES[i]
    g = torch.Generator(device="cuda").manual_seed(i)
    pad = lambda x: (x + 7) // 8 * 8
    # allocate with padded pitches so every case satisfies the 16-byte stride rule
    if a_mn:
        A_full = torch.randn(nb, K, pad(M), generator=g, device="cuda").bfloat16()
        A = A_full[:, :, :M]
        Af = A.float().transpose(1, 2)
    else:
        A_full = torch.randn(nb, M, pad(K), generator=g, device="cuda").bfloat16()
        A = A_full[:, :, :K]
        Af = A.float()
    if b_mn:
        B_full = torch.randn(nb, K, pad(N), generator=g, device="cuda").bfloat16()
        B = B_full[:, :, :N]
        Bf = B.float().transpose(1, 2)
    else:
        B_full = torch.randn(nb, N, pad(K), generator=g, device="cuda").bfloat16()
        B = B_full[:, :, :K]
        Bf = B.float()
    ref = torch.matmul(Af, Bf.transpose(1, 2))
    out = ops.gemm_bf16(A, B, a_mn=bool(a_mn), b_mn=bool(b_mn), block_n=bn)
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item()
    # accumulate path
    out2 = ops.gemm_bf16(A, B, a_mn=bool(a_mn), b_mn=bool(b_mn), block_n=bn, out=out.clone(), accumulate=True, alpha=0.5)
    torch.cuda.synchronize()
    err2 = (out2 - 1.5 * ref).abs().max().item()
    ok = err <= 2e-3 * max(scale, 1.0) and err2 <= 3e-3 * max(scale, 1.0)
    print(f"case {i:2d} M{M} N{N} K{K} b{nb} a_mn{a_mn} b_mn{b_mn} bn{bn}: max|err| {err:.3e} (acc {err2:.3e}) "
          f"scale {scale:.2f} -> {'OK' if ok else 'FAIL'}", flush=True)
    if not ok:
        bad = ((out - ref).abs() > 2e-3 * max(scale, 1.0)).nonzero()
        print("   first bad idx:", bad[:5].tolist(), "n_bad", bad.shape[0], "of", out.numel(), flush=True)
    return ok


if __name__ == "__main__":
    if len(sys.argv) > 1:
        sys.exit(0 if run_case(int(sys.argv[1])) else 1)
    fails = 0
    for i in range(len(CASES)):
        try:
            r = subprocess.run([sys.executable, __file__, str(i)], timeout=120)
            fails += r.returncode != 0
            if r.returncode not in (0, 1):
                print(f"case {i}: process exit {r.returncode}", flush=True)
        except subprocess.TimeoutExpired:
            print(f"case {i}: TIMEOUT", flush=True)
            fails += 1
    print(f"gemm_check: {len(CASES) - fails}/{len(CASES)} passed")
    sys.exit(1 if fails else 0)
