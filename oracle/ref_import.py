"""Import the *unmodified* reference modules on a CPU-only box.  TEST INFRASTRUCTURE ONLY.

Works only where ``/root/reference`` exists (the build container); used by
``oracle/make_golden.py`` to pin the restatements in ``cot_ref.py`` / ``agg_ref.py``
and by a few ``-m "not gpu"`` tests that skip when the tree is absent.

Two import shims are needed:
  * ``cupy``  -- ``cupy_layers/utils.py:4`` imports it; only ``cupy.memoize`` is touched
    at import time.  No arithmetic lives there.
  * ``yacs``  -- ``config/config.py:2``; only ``CfgNode`` as an attribute dict.
and the module global ``cupy_layers.aggregation_zeropad.aggregation_zeropad`` (looked up
at call time by ``LocalConvolution.forward``, ``aggregation_zeropad.py:221``) is replaced
by the Unfold identity the reference's own self-test equates it to (``:249-251``).
"""
import os
import sys
import types

REF = os.environ.get("COTB200_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "cupy_layers"))


def _install_shims():
    if "cupy" not in sys.modules:
        cupy = types.ModuleType("cupy")

        def memoize(for_each_device=False):
            def deco(f):
                return f
            return deco

        cupy.memoize = memoize
        cupy.cuda = types.SimpleNamespace(compile_with_cache=None)
        sys.modules["cupy"] = cupy
    if "yacs" not in sys.modules:
        yacs = types.ModuleType("yacs")
        yacs_config = types.ModuleType("yacs.config")

        class CfgNode(dict):
            def __init__(self, *a, **k):
                super().__init__(*a, **k)

            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError as e:
                    raise AttributeError(k) from e

            def __setattr__(self, k, v):
                self[k] = v

        yacs_config.CfgNode = CfgNode
        yacs.config = yacs_config
        sys.modules["yacs"] = yacs
        sys.modules["yacs.config"] = yacs_config


_loaded = {}


def load():
    """Returns a namespace with the reference's CotLayer / CoXtLayer / CoTLayer classes and
    model entry points, running on CPU with the op replaced by the Unfold identity."""
    if _loaded:
        return _loaded["ns"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF)
    _install_shims()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import cupy_layers.aggregation_zeropad as ref_agg  # noqa
    from oracle import agg_ref

    def _unfold_op(input, weight, kernel_size=3, stride=1, padding=0, dilation=1):
        return agg_ref.agg_zeropad_unfold(input, weight, kernel_size, stride, padding, dilation)

    ref_agg.aggregation_zeropad = _unfold_op
    import models.cotnet as ref_cotnet  # noqa
    import models.cotnet_hybrid as ref_hybrid  # noqa
    ns = types.SimpleNamespace(
        CotLayer=ref_cotnet.CotLayer,
        CoXtLayer=ref_cotnet.CoXtLayer,
        CoTLayer=ref_hybrid.CoTLayer,
        Bottleneck=ref_cotnet.Bottleneck,
        cotnet50=ref_cotnet.cotnet50,
        cotnext50_2x48d=ref_cotnet.cotnext50_2x48d,
        cotnet101=ref_cotnet.cotnet101,
        hybrid=ref_hybrid,
        agg_module=ref_agg,
    )
    _loaded["ns"] = ns
    return ns
