"""Backbones are OUT OF SCOPE of this build (BASELINE.json north_star: the ResNet student stays
stock PyTorch).  To run the reference's drivers unchanged, point LITE_MKD_REFERENCE at a checkout
of the reference; its model/backbone/*.py files are then importable as model.backbone.<name>.
`precomputed` is a pass-through for episodes whose features already exist (synthetic benches)."""
import os

_ref = os.environ.get("LITE_MKD_REFERENCE")
if _ref and os.path.isdir(os.path.join(_ref, "model", "backbone")):
    __path__.append(os.path.join(_ref, "model", "backbone"))
