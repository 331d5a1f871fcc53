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
import importlib.abc

# NB: the package re-exports functions named like these sub-modules, so fetch the MODULES explicitly
_agg = importlib.import_module(__package__ + ".aggregation_zeropad")
_mix = importlib.import_module(__package__ + ".aggregation_zeropad_mix")

_VARIANTS = {name: importlib.import_module(__package__ + "." + name)
             for name in ("aggregation_refpad", "aggregation_zeropad_dilate", "aggregation_zeropad_mix_merge")}


def patch_reference(swap_layers=True):
    """Alias `cupy_layers.*` (all five operator modules) to this package.  Call BEFORE importing the reference's `models`
    package; the layer classes are swapped by a post-import hook."""
    pkg = sys.modules.get("cupy_layers")
    if pkg is None or not isinstance(pkg, types.ModuleType):
        pkg = types.ModuleType("cupy_layers")
        pkg.__path__ = []
        sys.modules["cupy_layers"] = pkg
    sys.modules["cupy_layers.aggregation_zeropad"] = _agg
    sys.modules["cupy_layers.aggregation_zeropad_mix"] = _mix
    pkg.aggregation_zeropad = _agg
    pkg.aggregation_zeropad_mix = _mix
    for name, mod in _VARIANTS.items():
        sys.modules["cupy_layers." + name] = mod
        setattr(pkg, name, mod)
    if swap_layers:
        swap_layer_classes()


def _swap_table():
    from .cot_layer import CotLayer, CoXtLayer
    return {"models.cotnet": (("CotLayer", CotLayer), ("CoXtLayer", CoXtLayer)),
            "models.cotnet_hybrid": (("CoTLayer", CotLayer),)}


class _SwapOnImport(importlib.abc.MetaPathFinder):
    """Post-import hook: when `models.cotnet` / `models.cotnet_hybrid` are imported AFTER patch_reference(), their
    layer classes are replaced as soon as the module body has run (the zoo's factory functions look the class names up
    at call time, models/cotnet.py:199-204, models/cotnet_hybrid.py:138-153)."""

    def find_spec(self, fullname, path, target=None):
        if fullname not in ("models.cotnet", "models.cotnet_hybrid"):
            return None
        for finder in sys.meta_path:
            if finder is self or not hasattr(finder, "find_spec"):
                continue
            spec = finder.find_spec(fullname, path, target)
            if spec is not None and spec.loader is not None and hasattr(spec.loader, "exec_module"):
                inner = spec.loader.exec_module

                def exec_module(module, _inner=inner, _name=fullname):
                    _inner(module)
                    for a, cls in _swap_table()[_name]:
                        setattr(module, a, cls)
                spec.loader.exec_module = exec_module
                return spec
        return None


_HOOK = _SwapOnImport()


def swap_layer_classes():
    """Point the reference's layer classes at the fused ones: immediately for model modules that are already imported,
    and through a post-import hook for those imported later (so the documented order -- patch_reference() BEFORE
    `import models` -- really swaps them)."""
    for modname, attrs in _swap_table().items():
        mod = sys.modules.get(modname)
        if mod is not None:
            for a, cls in attrs:
                setattr(mod, a, cls)
    if _HOOK not in sys.meta_path:
        sys.meta_path.insert(0, _HOOK)
