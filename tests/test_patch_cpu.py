"""The zero-edit drop-in (cotnet_b200/patch.py): the unmodified reference model zoo imports and builds on top of the
mirrors, with identical parameter names.  Needs /root/reference, so it only runs in the build container."""
import subprocess
import sys
import os

import pytest

from oracle import ref_import

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys, types
sys.path.insert(0, %r); sys.path.insert(0, %r)
# yacs is absent from the image (config only, no arithmetic): the same shim the oracle uses
yacs = types.ModuleType("yacs"); yc = types.ModuleType("yacs.config")
class CfgNode(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__
yc.CfgNode = CfgNode; yacs.config = yc; sys.modules["yacs"] = yacs; sys.modules["yacs.config"] = yc
import cotnet_b200.patch as p
p.patch_reference()
import models.cotnet as mc            # imported AFTER patch_reference(): the post-import hook swaps the layer classes
import models.cotnet_hybrid as mh
from cotnet_b200.cot_layer import CotLayer
from cotnet_b200.aggregation_zeropad import LocalConvolution
import cupy_layers.aggregation_zeropad as ca
assert ca.LocalConvolution is LocalConvolution and "cupy" not in sys.modules
m = mc.cotnet50()
layers = [x for x in m.modules() if isinstance(x, CotLayer)]
assert len(layers) == 16, len(layers)
assert isinstance(layers[0].local_conv, LocalConvolution)
assert sum(q.numel() for q in m.parameters()) == 22222416
assert mc.CotLayer is CotLayer and mh.CoTLayer is CotLayer
# the rest of the zoo and the flops counter import all five cupy_layers operator modules: every one resolves to a mirror
import models.lr_net, models.san_lowrank, models.botnet, utils.flops_counter
import cupy_layers.aggregation_zeropad_dilate as cd, cupy_layers.aggregation_zeropad_mix_merge as cm, cupy_layers.aggregation_refpad as cr
assert cd.__name__.startswith("cotnet_b200") and cm.__name__.startswith("cotnet_b200") and cr.__name__.startswith("cotnet_b200")
assert "cupy" not in sys.modules
print("OK")
'''


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_reference_zoo_builds_on_the_mirrors():
    r = subprocess.run([sys.executable, "-c", SCRIPT % (ROOT, ref_import.REF)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stderr[-2000:]
