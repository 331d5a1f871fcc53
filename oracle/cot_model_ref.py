"""CPU restatement of the CoTNet / CoTNeXt / SE-CoTNetD models (the callers of the hot path).  TEST INFRASTRUCTURE ONLY.

Used as (i) the checker for the product backbone in tests and (ii) the timed CPU baseline of ``bench.py``
(``cpu_baseline`` / ``--impl reference``): the reference has no CPU implementation of the LocalConv op
(SURVEY.md D5), so "the reference's CPU forward" is its module graph on CPU with the op replaced by the
Unfold identity of its own self-test -- exactly what this file restates:
    trunk        /root/reference/models/resnet.py:448-611  (7x7 stem, max-pool, 4 stages, avg-pool, fc)
    bottleneck   /root/reference/models/cotnet.py:181-264  (1x1 -> [avgpool 3x3/2] -> CoT -> 1x1 -> +res -> ReLU)
    CoT layer    oracle/cot_ref.py (pinned to golden vectors of the reference's own module code)
Parameter names equal the reference's, so one state dict drives the reference, this oracle and the product.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import cot_ref


class OracleCotLayer(nn.Module):
    """Holds the reference-named parameters; forward = cot_ref.cot_layer / coxt_layer."""

    def __init__(self, kind, dim):
        super().__init__()
        self.kind, self.dim = kind, dim
        sd = cot_ref.init_state_dict(kind, dim, torch.Generator().manual_seed(0), dtype=torch.float32, perturb=False)
        self._names = []
        for k, v in sd.items():
            flat = k.replace(".", "__")
            self._names.append((k, flat))
            if "running" in k or "num_batches" in k:
                self.register_buffer(flat, v)
            else:
                self.register_parameter(flat, nn.Parameter(v))

    def _sd(self):
        return {k: getattr(self, flat) for k, flat in self._names}

    def load_reference_state(self, sd, prefix=""):
        with torch.no_grad():
            for k, flat in self._names:
                getattr(self, flat).copy_(sd[prefix + k])

    def forward(self, x):
        fn = cot_ref.cot_layer if self.kind == "cot" else cot_ref.coxt_layer
        return fn(x, self._sd(), training=self.training)


class OracleBottleneck(nn.Module):
    def __init__(self, inplanes, planes, stride, down, cardinality, base_width):
        super().__init__()
        width = int(math.floor(planes * (base_width / 64)) * cardinality)
        self.conv1 = nn.Conv2d(inplanes, width, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        self.stride = stride
        self.conv2 = OracleCotLayer("cot" if cardinality == 1 else "coxt", width)
        self.conv3 = nn.Conv2d(width, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = down

    def forward(self, x):
        r = x
        x = F.relu(self.bn1(self.conv1(x)))
        if self.stride > 1:
            x = F.avg_pool2d(x, 3, 2, 1)                      # models/cotnet.py:199-202,237-238
        x = self.bn3(self.conv3(self.conv2(x)))
        if self.downsample is not None:
            r = self.downsample(r)
        return F.relu(x + r)


class OracleCoTResNet(nn.Module):
    def __init__(self, layers=(3, 4, 6, 3), cardinality=1, base_width=64, num_classes=1000):
        super().__init__()
        inpl = 64
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        for i, (planes, n) in enumerate(zip((64, 128, 256, 512), layers)):
            blocks = []
            for b in range(n):
                s = 2 if (b == 0 and i > 0) else 1
                down = None
                if b == 0 and (s != 1 or inpl != planes * 4):
                    down = nn.Sequential(nn.Conv2d(inpl, planes * 4, 1, stride=s, bias=False), nn.BatchNorm2d(planes * 4))
                blocks.append(OracleBottleneck(inpl, planes, s, down, cardinality, base_width))
                inpl = planes * 4
            setattr(self, "layer%d" % (i + 1), nn.Sequential(*blocks))
        self.fc = nn.Linear(inpl, num_classes)

    def load_reference_state(self, sd):
        """Load a reference / product state dict (names of models/resnet.py + models/cotnet.py)."""
        own = {}
        for name, mod in self.named_modules():
            if isinstance(mod, OracleCotLayer):
                mod.load_reference_state(sd, name + ".")
        for k, v in self.state_dict().items():
            if "__" in k:
                continue
            own[k] = sd[k]
        self.load_state_dict(own, strict=False)

    def forward(self, x):
        x = F.max_pool2d(F.relu(self.bn1(self.conv1(x))), 3, 2, 1)
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.fc(x.mean((2, 3)))


# ------------------------------------------------------------------------------------------------ SE-CoTNetD (hybrid)
class OracleSplitAttn(nn.Module):
    """models/layers/split_attn.py:31-88 with radix 1, groups 1, swish: conv -> bn0 -> SiLU -> GAP -> fc1 -> bn1 -> SiLU ->
    fc2 -> sigmoid (RadixSoftmax with radix 1, :26-27) -> rescale."""

    def __init__(self, ch, stride):
        super().__init__()
        attn = max(ch // 4, 32)
        self.conv = nn.Conv2d(ch, ch, 3, stride, 1, bias=False)
        self.bn0 = nn.BatchNorm2d(ch)
        self.fc1 = nn.Conv2d(ch, attn, 1)
        self.bn1 = nn.BatchNorm2d(attn)
        self.fc2 = nn.Conv2d(attn, ch, 1)

    def forward(self, x):
        x = F.silu(self.bn0(self.conv(x)))
        g = x.mean((2, 3), keepdim=True)
        return x * torch.sigmoid(self.fc2(F.silu(self.bn1(self.fc1(g)))))


def _blur_pool(x):
    """models/layers/blur_pool.py:19-58, filt_size 3, stride 2: reflect-pad 1, depthwise binomial [1,2,1]^2/16."""
    c = torch.tensor([0.25, 0.5, 0.25], dtype=x.dtype, device=x.device)
    f = (c[:, None] * c[None, :])[None, None].repeat(x.shape[1], 1, 1, 1)
    return F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), f, stride=2, groups=x.shape[1])


class OracleHybridBottleneck(nn.Module):
    """models/cotnet_hybrid.py:118-204 (cardinality 1, no drop block / drop path)."""

    def __init__(self, inpl, planes, stride, down, use_cot, blur, avd, avd_first):
        super().__init__()
        self.conv1 = nn.Conv2d(inpl, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.pool, self.avd_first, self.blur = False, avd_first, blur
        if use_cot:
            self.conv2 = OracleCotLayer("cot", planes)
            self.pool = stride > 1
        else:
            if stride > 1 and avd:
                self.pool, stride = True, 1
            self.conv2 = OracleSplitAttn(planes, stride)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = down

    def _avd(self, x):
        return _blur_pool(x) if self.blur else F.avg_pool2d(x, 3, 2, 1)

    def forward(self, x):
        r = x
        x = F.relu(self.bn1(self.conv1(x)))
        if self.pool and self.avd_first:
            x = self._avd(x)
        x = self.conv2(x)
        if self.pool and not self.avd_first:
            x = self._avd(x)
        x = self.bn3(self.conv3(x))
        if self.downsample is not None:
            r = self.downsample(r)
        return F.relu(x + r)


class OracleHybridNet(nn.Module):
    """models/cotnet_hybrid.py:338-452 as configured by se_cotnetd_{50,101,152} (:458-482): deep stem, avg-pool down-sampling,
    stride 2 in every stage, SplitAttn in layers 1-2 and the odd blocks of layer 3, CoT layer elsewhere."""

    def __init__(self, layers, stem_width, blur=False, avd=False, avd_first=True, num_classes=1000):
        super().__init__()
        inpl = stem_width * 2
        self.conv1 = nn.Sequential(
            nn.Conv2d(3, stem_width, 3, 2, 1, bias=False), nn.BatchNorm2d(stem_width), nn.ReLU(),
            nn.Conv2d(stem_width, stem_width, 3, 1, 1, bias=False), nn.BatchNorm2d(stem_width), nn.ReLU(),
            nn.Conv2d(stem_width, inpl, 3, 1, 1, bias=False))
        self.bn1 = nn.BatchNorm2d(inpl)
        for i, (planes, n) in enumerate(zip((64, 128, 256, 512), layers)):
            blocks = []
            for b in range(n):
                s = 2 if b == 0 else 1
                down = None
                if b == 0:
                    down = nn.Sequential(nn.AvgPool2d(2, 2, ceil_mode=True, count_include_pad=False),
                                         nn.Conv2d(inpl, planes * 4, 1, bias=False), nn.BatchNorm2d(planes * 4))
                use_cot = planes == 512 or (planes == 256 and b % 2 == 0)         # conv_dim={64,128}, c4_dim=256, c4_idx=even (:138)
                blocks.append(OracleHybridBottleneck(inpl, planes, s, down, use_cot, blur, avd, avd_first))
                inpl = planes * 4
            setattr(self, "layer%d" % (i + 1), nn.Sequential(*blocks))
        self.fc = nn.Linear(inpl, num_classes)

    load_reference_state = OracleCoTResNet.load_reference_state

    def forward(self, x):
        x = F.relu(self.bn1(self.conv1(x)))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.fc(x.mean((2, 3)))


def build(name="cotnet50"):
    cfg = {"cotnet50": ((3, 4, 6, 3), 1, 64), "cotnext50_2x48d": ((3, 4, 6, 3), 2, 48),
           "cotnet101": ((3, 4, 23, 3), 1, 64), "cotnext101_2x48d": ((3, 4, 23, 3), 2, 48)}
    hyb = {"se_cotnetd_50": dict(layers=(3, 4, 6, 3), stem_width=32), "se_cotnetd_101": dict(layers=(3, 4, 23, 3), stem_width=64),
           "se_cotnetd_152": dict(layers=(3, 8, 36, 3), stem_width=64, blur=True, avd=True, avd_first=False)}
    if name in hyb:
        return OracleHybridNet(**hyb[name])
    return OracleCoTResNet(*cfg[name])
