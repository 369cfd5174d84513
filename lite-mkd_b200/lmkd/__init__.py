"""Host-side glue of the B200-native Lite-MKD matching + D2M path.

`_ffi` binds liblmkd.so (C ABI in include/lmkd.h), `ops` wraps it in autograd Functions,
`episodes` makes the synthetic workloads of BASELINE.json, `dist` shards episodes over ranks.
The reference-facing modules (`distillers`, `model.classifiers`, `model.model_select`, `utils`)
live one directory up so the reference's drivers import them by their original names.
"""
from . import _ffi, ops  # noqa: F401
from ._ffi import check_device_status  # noqa: F401
