"""Import the UNMODIFIED reference (read-only tree at /root/reference) for oracle
validation and golden-vector generation.  Runs only in the build container: the
GPU box has no /root/reference, and nothing under `-m gpu`, smoke() or bench.py
imports this file.

Shims — none of them edits the reference:
  * empty `matplotlib` / `matplotlib.pyplot` / `torchvision...` modules so root `utils.py`
    imports;
  * `sys.dont_write_bytecode` (the tree is read-only);
  * GDLNet: a per-instance wrapper giving `_output_padding` the `num_spatial_dims`
    argument torch >= 2.x requires.
"""
import os
import sys
import types

REF = os.environ.get("CDL_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "model", "net.py"))


def load():
    """Returns the reference's `model.net` module."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    sys.dont_write_bytecode = True
    for name in ("matplotlib", "matplotlib.pyplot", "PIL", "PIL.Image",
                 "torchvision", "torchvision.transforms", "torchvision.transforms.functional"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                if name.endswith("functional"):
                    m.to_tensor = lambda *a, **k: None
                sys.modules[name] = m
                parent, _, child = name.rpartition(".")
                if parent:
                    setattr(sys.modules[parent], child, m)
    # our own `model` package (the drop-in) must not shadow the reference's
    saved = {k: v for k, v in sys.modules.items() if k == "model" or k.startswith("model.") or k == "utils"}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REF)
    try:
        import model.net as ref_net          # noqa
        import model.utils as ref_utils      # noqa
        import utils as ref_root_utils       # noqa
    finally:
        sys.path.remove(REF)
    mods = {k: sys.modules[k] for k in list(sys.modules) if k == "model" or k.startswith("model.") or k == "utils"}
    for k in mods:
        del sys.modules[k]
    sys.modules.update(saved)
    ref_net._ref_utils = ref_utils
    ref_net._ref_root_utils = ref_root_utils
    return ref_net


def load_nle(filter_bank):
    """Returns the reference's `model.nle` module.  Its `model.wvlt` imports PyWavelets (absent here): a stub
    `pywt` whose Wavelet('bior4.4').filter_bank is `filter_bank` stands in, so that everything the reference does WITH
    the coefficients (outer products, flips, conv2d, median) is the reference's own code."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    sys.dont_write_bytecode = True
    stub = types.ModuleType("pywt")

    class Wavelet:
        def __init__(self, name):
            if name != "bior4.4":
                raise ValueError(name)
            self.filter_bank = filter_bank
    stub.Wavelet = Wavelet
    saved = {k: v for k, v in sys.modules.items() if k == "model" or k.startswith("model.") or k in ("utils", "pywt")}
    for k in saved:
        del sys.modules[k]
    sys.modules["pywt"] = stub
    sys.path.insert(0, REF)
    try:
        import model.nle as ref_nle          # noqa
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == "model" or k.startswith("model.") or k in ("utils", "pywt")]:
            del sys.modules[k]
        sys.modules.update(saved)
    return ref_nle


def fix_gdlnet(net):
    """wrap the private `_output_padding` of every Gabor layer."""
    for mod in list(net.A) + list(net.B):
        orig = mod._output_padding
        if getattr(orig, "_cdl_wrapped", False):
            continue

        def wrapped(x, output_size, stride, padding, kernel_size, _orig=orig):
            return _orig(x, output_size, list(stride), list(padding), list(kernel_size), 2)
        wrapped._cdl_wrapped = True
        mod._output_padding = wrapped
    return net
