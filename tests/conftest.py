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
