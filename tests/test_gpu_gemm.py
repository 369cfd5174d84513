"""The tcgen05 GEMM against torch.matmul on bf16-rounded inputs: all four operand major-ness
combinations, ragged M/N/K, batches, accumulate epilogue."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def test_all_gemm_cases():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import gemm_check
    bad = [i for i in range(len(gemm_check.CASES)) if not gemm_check.run_case(i)]
    assert not bad, f"failing gemm cases: {bad}"
