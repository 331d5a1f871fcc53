"""Zero-edit drop-in: make an unmodified JDAI-CV/CoTNet checkout run its CoT path on libcotb200.

    import cotnet_b200.patch as p
    p.patch_reference()            # before `import models`
    import models                  # reference zoo, no CuPy needed for cotnet*/se_cotnetd*

The reference picks its LocalConv by hard import (`models/cotnet.py:12`); this aliases the `cupy_layers` modules it
imports to the mirrors in this package and (optionally) swaps the layer classes for the fused ones.
"""
import sys
import types

import importlib

# NB: the package re-exports functions named like these sub-modules, so fetch the MODULES explicitly
_agg = importlib.import_module(__package__ + ".aggregation_zeropad")
_mix = importlib.import_module(__package__ + ".aggregation_zeropad_mix")

# variants imported by lr_net / botnet / flops_counter but constructed by no registered model
_UNUSED_VARIANTS = {
    "cupy_layers.aggregation_zeropad_dilate": ("LocalConvolutionDilate", "aggregation_zeropad_dilate"),
    "cupy_layers.aggregation_zeropad_mix_merge": ("LocalConvolutionMixMerge", "aggregation_zeropad_mix_merge"),
    "cupy_layers.aggregation_refpad": ("aggregation_refpad",),
}


def _unsupported(name):
    def fn(*a, **k):
        raise NotImplementedError("%s is outside the CoT hot path implemented by cotnet_b200 (SURVEY.md section 8f)" % name)
    fn.__name__ = name
    return fn


def patch_reference(swap_layers=True, stub_unused_variants=True):
    """Alias `cupy_layers.*` to this package.  Call BEFORE importing the reference's `models` package."""
    pkg = sys.modules.get("cupy_layers")
    if pkg is None or not isinstance(pkg, types.ModuleType):
        pkg = types.ModuleType("cupy_layers")
        pkg.__path__ = []
        sys.modules["cupy_layers"] = pkg
    sys.modules["cupy_layers.aggregation_zeropad"] = _agg
    sys.modules["cupy_layers.aggregation_zeropad_mix"] = _mix
    pkg.aggregation_zeropad = _agg
    pkg.aggregation_zeropad_mix = _mix
    if stub_unused_variants:
        for modname, names in _UNUSED_VARIANTS.items():
            if modname in sys.modules:
                continue
            m = types.ModuleType(modname)
            for n in names:
                if n[0].isupper():
                    setattr(m, n, type(n, (object,), {"__init__": _unsupported(n)}))
                else:
                    setattr(m, n, _unsupported(n))
            sys.modules[modname] = m
            setattr(pkg, modname.split(".")[-1], m)
    if swap_layers:
        swap_layer_classes()


def swap_layer_classes():
    """If the reference's model modules are (or get) imported, point their layer classes at the fused ones."""
    from .cot_layer import CotLayer, CoXtLayer
    for modname, attrs in (("models.cotnet", (("CotLayer", CotLayer), ("CoXtLayer", CoXtLayer))),
                           ("models.cotnet_hybrid", (("CoTLayer", CotLayer),))):
        mod = sys.modules.get(modname)
        if mod is not None:
            for a, cls in attrs:
                setattr(mod, a, cls)
