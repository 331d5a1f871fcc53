"""Launch the reference's own LocalConv kernels (cubins built by oracle/build_ref_kernels.py) on the GPU.
TEST INFRASTRUCTURE ONLY -- used by tests/test_ref_kernels_gpu.py and tools/bench_ref_kernels.py.

The launch mirrors cupy_layers/aggregation_zeropad.py:140-143 (block (1024,1,1), grid (GET_BLOCKS(n),1,1), raw device
pointers, torch's current stream) through the CUDA driver API (ctypes on libcuda; no CuPy, no /root/reference at run time).
"""
import ctypes
import json
import os

import torch

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_TORCH = {"float": torch.float32, "double": torch.float64}


def available():
    return os.path.exists(os.path.join(HERE, "manifest.json"))


class RefKernels:
    def __init__(self):
        if not available():
            raise RuntimeError("oracle/_ref/manifest.json missing: run `python -m oracle.build_ref_kernels` where /root/reference exists")
        self.manifest = json.load(open(os.path.join(HERE, "manifest.json")))
        self.cu = ctypes.CDLL("libcuda.so.1")
        torch.cuda.init()
        torch.zeros(1, device="cuda")                     # make torch's primary context current
        self._fn = {}

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError("%s failed with CUresult %d" % (what, rc))

    def entry(self, kernel, tag, dtype):
        for e in self.manifest["kernels"]:
            if e["kernel"] == kernel and e["tag"] == tag and e["dtype"] == dtype:
                return e
        raise KeyError((kernel, tag, dtype))

    def tags(self, op):
        seen = []
        for e in self.manifest["kernels"]:
            if e["op"] == op and (e["tag"], e["dtype"]) not in seen:
                seen.append((e["tag"], e["dtype"]))
        return seen

    def function(self, e):
        key = e["file"]
        if key not in self._fn:
            blob = open(os.path.join(HERE, e["file"]), "rb").read()
            mod, fn = ctypes.c_void_p(), ctypes.c_void_p()
            self._check(self.cu.cuModuleLoadData(ctypes.byref(mod), blob), "cuModuleLoadData")
            self._check(self.cu.cuModuleGetFunction(ctypes.byref(fn), mod, e["kernel"].encode()), "cuModuleGetFunction")
            self._fn[key] = (mod, fn, blob)
        return self._fn[key][1]

    def launch(self, e, *tensors):
        fn = self.function(e)
        vals = [ctypes.c_void_p(t.data_ptr()) for t in tensors]
        argv = (ctypes.c_void_p * len(vals))(*[ctypes.cast(ctypes.pointer(v), ctypes.c_void_p) for v in vals])
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        self._check(self.cu.cuLaunchKernel(fn, e["grid"], 1, 1, self.manifest["block"], 1, 1, 0, st, argv, None), "cuLaunchKernel")

    # ---- the three calls of AggregationZeropad.forward / backward (aggregation_zeropad.py:112-186) on NCHW-contiguous tensors
    def agg_forward(self, tag, x, w):
        e = self.entry("aggregation_zeropad_forward_kernel", tag, "float" if x.dtype == torch.float32 else "double")
        y = torch.empty(e["N"], e["heads"] * e["C"], e["Ho"], e["Wo"], dtype=x.dtype, device=x.device)
        self.launch(e, x, w, y)
        return y

    def agg_backward(self, tag, dy, x, w):
        dt = "float" if x.dtype == torch.float32 else "double"
        dx, dw = torch.empty_like(x), torch.empty_like(w)
        self.launch(self.entry("aggregation_zeropad_input_backward_kernel", tag, dt), dy, w, dx)
        self.launch(self.entry("aggregation_zeropad_weight_backward_kernel", tag, dt), dy, x, dw)
        return dx, dw

    # ---- AggregationZeropadMix (aggregation_zeropad_mix.py:209-288)
    def mix_forward(self, tag, x, w1, w2):
        e = self.entry("aggregation_zeropad_mix_forward_kernel", tag, "float" if x.dtype == torch.float32 else "double")
        y = torch.empty(e["N"], 2 * e["heads"] * e["C"], e["H"], e["W"], dtype=x.dtype, device=x.device)
        self.launch(e, x, w1, w2, y)
        return y

    def mix_backward(self, tag, dy, x, w1, w2):
        dt = "float" if x.dtype == torch.float32 else "double"
        dx, dw1, dw2 = torch.empty_like(x), torch.empty_like(w1), torch.empty_like(w2)
        self.launch(self.entry("aggregation_zeropad_mix_input_backward_kernel", tag, dt), dy, w1, w2, dx)
        self.launch(self.entry("aggregation_zeropad_mix_weight_backward_kernel", tag, dt), dy, x, dw1, dw2)
        return dx, dw1, dw2

    # ---- AggregationRefpad (aggregation_refpad.py:129-208): dX is computed on the padded grid, then folded (:188-199)
    def refpad_forward(self, tag, x, w):
        e = self.entry("aggregation_refpad_forward_kernel", tag, "float" if x.dtype == torch.float32 else "double")
        y = torch.empty(e["N"], e["heads"] * e["C"], e["Ho"], e["Wo"], dtype=x.dtype, device=x.device)
        self.launch(e, x, w, y)
        return y

    def refpad_backward(self, tag, dy, x, w):
        dt = "float" if x.dtype == torch.float32 else "double"
        e = self.entry("aggregation_refpad_input_backward_kernel", tag, dt)
        p, H, W = e["pad"], e["H"], e["W"]
        gi = torch.empty(e["N"], e["C"], H + 2 * p, W + 2 * p, dtype=x.dtype, device=x.device)
        self.launch(e, dy, w, gi)
        # the reference's host-side border fold, aggregation_refpad.py:195-199
        gi[:, :, p + 1:2 * p + 1, :] += torch.flip(gi[:, :, :p, :], dims=[2])
        gi[:, :, H - 1:H + p - 1, :] += torch.flip(gi[:, :, H + p:, :], dims=[2])
        gi[:, :, :, p + 1:2 * p + 1] += torch.flip(gi[:, :, :, :p], dims=[3])
        gi[:, :, :, W - 1:W + p - 1] += torch.flip(gi[:, :, :, W + p:], dims=[3])
        dx = gi[:, :, p:p + H, p:p + W].contiguous()
        dw = torch.empty_like(w)
        self.launch(self.entry("aggregation_refpad_weight_backward_kernel", tag, dt), dy, x, dw)
        return dx, dw

    # ---- AggregationZeropadDilate (aggregation_zeropad_dilate.py:148-218)
    def dilate_forward(self, tag, x, w, dil):
        e = self.entry("aggregation_zeropad_dilate_forward_kernel", tag, "float" if x.dtype == torch.float32 else "double")
        y = torch.empty(e["N"], e["heads"] * e["C"], e["H"], e["W"], dtype=x.dtype, device=x.device)
        self.launch(e, x, w, dil, y)
        return y

    def dilate_backward(self, tag, dy, x, w, dil):
        dt = "float" if x.dtype == torch.float32 else "double"
        dx, dw = torch.empty_like(x), torch.empty_like(w)
        self.launch(self.entry("aggregation_zeropad_dilate_input_backward_kernel", tag, dt), dy, w, dil, dx)
        self.launch(self.entry("aggregation_zeropad_dilate_weight_backward_kernel", tag, dt), dy, x, dil, dw)
        return dx, dw

    # ---- AggregationZeropadMixMerge (aggregation_zeropad_mix_merge.py:180-274)
    def merge_forward(self, tag, x, w):
        e = self.entry("aggregation_zeropad_mix_merge_forward_kernel", tag, "float" if x.dtype == torch.float32 else "double")
        y = torch.empty(e["N"], 2 * e["heads"] * e["C"], e["H"], e["W"], dtype=x.dtype, device=x.device)
        self.launch(e, x, w, y)
        return y

    def merge_backward(self, tag, dy, x, w):
        dt = "float" if x.dtype == torch.float32 else "double"
        dx, dw = torch.empty_like(x), torch.empty_like(w)
        self.launch(self.entry("aggregation_zeropad_mix_merge_input_backward_kernel", tag, dt), dy, w, dx)
        self.launch(self.entry("aggregation_zeropad_mix_merge_weight_backward_kernel", tag, dt), dy, x, dw)
        return dx, dw

    def make_variant_inputs(self, tag, dtype, op, seed=0):
        e = next(k for k in self.manifest["kernels"] if k["tag"] == tag and k["op"] == op)
        g = torch.Generator(device="cuda").manual_seed(seed)
        td = _TORCH[dtype]
        N, C, H, W, heads, wc = e["N"], e["C"], e["H"], e["W"], e["heads"], e["wc"]
        x = torch.randn(N, C, H, W, generator=g, device="cuda", dtype=td)
        if op == "refpad":
            w = torch.randn(N, heads, wc, e["k"] ** 2, e["Ho"], e["Wo"], generator=g, device="cuda", dtype=td)
            dy = torch.randn(N, heads * C, e["Ho"], e["Wo"], generator=g, device="cuda", dtype=td)
            return e, x, w, dy
        if op == "dilate":
            w = torch.randn(N, heads, wc, 9, H, W, generator=g, device="cuda", dtype=td)
            dil = torch.tensor([(1, 1, 2, 4, 3, 2, 1, 5)[i % 8] for i in range(wc)], device="cuda", dtype=td)
            dy = torch.randn(N, heads * C, H, W, generator=g, device="cuda", dtype=td)
            return e, x, w, dil, dy
        w = torch.randn(N, heads * wc * 34, H, W, generator=g, device="cuda", dtype=td)
        dy = torch.randn(N, 2 * heads * C, H, W, generator=g, device="cuda", dtype=td)
        return e, x, w, dy

    def make_inputs(self, tag, dtype, op="agg", seed=0):
        e = next(k for k in self.manifest["kernels"] if k["tag"] == tag and k["op"] == op)
        g = torch.Generator(device="cuda").manual_seed(seed)
        td = _TORCH[dtype]
        x = torch.randn(e["N"], e["C"], e["H"], e["W"], generator=g, device="cuda", dtype=td)
        if op == "agg":
            w = torch.randn(e["N"], e["heads"], e["wc"], e["k"] ** 2, e["Ho"], e["Wo"], generator=g, device="cuda", dtype=td)
            dy = torch.randn(e["N"], e["heads"] * e["C"], e["Ho"], e["Wo"], generator=g, device="cuda", dtype=td)
            return e, x, w, dy
        w1 = torch.randn(e["N"], e["heads"], e["wc"], 9, e["H"], e["W"], generator=g, device="cuda", dtype=td)
        w2 = torch.randn(e["N"], e["heads"], e["wc"], 25, e["H"], e["W"], generator=g, device="cuda", dtype=td)
        dy = torch.randn(e["N"], 2 * e["heads"] * e["C"], e["H"], e["W"], generator=g, device="cuda", dtype=td)
        return e, x, w1, w2, dy
