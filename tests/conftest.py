"""pytest configuration: markers and import paths.

`-m "not gpu"` runs everywhere (oracle vs golden fixtures, host logic, C-ABI symbol checks);
`-m gpu` needs a B200 and exercises the CUDA path through the C-ABI.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "lite-mkd_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


# ---- achieved-error ledger ------------------------------------------------------------------------
# Every GPU parity test reports the error it actually measured (not only pass/fail against the tolerance);
# the session writes them to gpurun_out/parity_errors.json, which is copied to profiles/rNN_parity_errors.json.
_ERRORS = {}


def record_error(tag, **values):
    entry = _ERRORS.setdefault(tag, {})
    for k, v in values.items():
        entry[k] = float(v)


def pytest_sessionfinish(session, exitstatus):
    if not _ERRORS:
        return
    import json
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        path = os.path.join(out, "parity_errors.json")
        merged = {}
        if os.path.exists(path):
            try:
                merged = json.load(open(path))
            except Exception:
                merged = {}
        merged.update(_ERRORS)
        json.dump(merged, open(path, "w"), indent=1, sort_keys=True)
    except OSError:
        pass


# comparison helpers that LOG what they measured (drop-in for np.testing.assert_allclose / torch.allclose / rel-L2)
_calls = {}


def _tag(what=None):
    t = os.environ.get("PYTEST_CURRENT_TEST", "unknown").split(" ")[0].split("::")[-1]
    if what is None:
        n = _calls.get(t, 0)
        _calls[t] = n + 1
        what = f"cmp{n:02d}"
    return t, what


def _as_f64(x):
    import numpy as np
    import torch
    if torch.is_tensor(x):
        return x.detach().double().cpu()
    return torch.from_numpy(np.asarray(x)).double()


def rel_l2(a, b, what=None):
    """|a - b|_2 / |b|_2, recorded in the error ledger under the running test's name."""
    a, b = _as_f64(a), _as_f64(b)
    err = float((a - b).norm() / b.norm().clamp_min(1e-30))
    t, w = _tag(what)
    record_error(t, **{f"{w}.rel_l2": err})
    return err


def _close_stats(a, b, rtol, atol):
    a, b = _as_f64(a), _as_f64(b)
    diff = (a - b).abs()
    finite = diff[diff == diff]
    max_abs = float(finite.max()) if finite.numel() else 0.0
    # how much of the allowed band |a-b| <= atol + rtol |b| was used (1.0 = at the limit)
    band = float((diff / (atol + rtol * b.abs()).clamp_min(1e-300)).nan_to_num(0.0).max()) if diff.numel() else 0.0
    scale = float(b.abs().max()) if b.numel() else 0.0
    return max_abs, band, scale


def assert_close(actual, desired, rtol=1e-7, atol=0.0, what=None, **kw):
    import numpy as np
    max_abs, band, scale = _close_stats(actual, desired, rtol, atol)
    t, w = _tag(what)
    record_error(t, **{f"{w}.max_abs": max_abs, f"{w}.ref_max": scale, f"{w}.band_used": band})
    a = actual.detach().cpu().numpy() if hasattr(actual, "detach") else np.asarray(actual)
    d = desired.detach().cpu().numpy() if hasattr(desired, "detach") else np.asarray(desired)
    np.testing.assert_allclose(a, d, rtol=rtol, atol=atol, **kw)


def allclose(a, b, rtol=1e-5, atol=1e-8, what=None):
    import torch
    max_abs, band, scale = _close_stats(a, b, rtol, atol)
    t, w = _tag(what)
    record_error(t, **{f"{w}.max_abs": max_abs, f"{w}.ref_max": scale, f"{w}.band_used": band})
    return bool(torch.allclose(torch.as_tensor(a).detach().cpu().double(), torch.as_tensor(b).detach().cpu().double(),
                               rtol=rtol, atol=atol))
