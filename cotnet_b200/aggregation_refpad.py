"""Drop-in for the reference's ``cupy_layers/aggregation_refpad.py`` (LocalConv with reflect padding).

    AggregationRefpad.apply(input, weight, kernel_size, stride, padding, dilation)   /root/reference/cupy_layers/aggregation_refpad.py:129-208
    aggregation_refpad(input, weight, kernel_size=3, stride=1, padding=0, dilation=1)  :211-221

Same names, signatures, asserts and shapes; the kernels are libcotb200's (cotb200_agg_refpad_{fwd,bwd}, csrc/agg_variants.cu).
The reference computes dX on the padded grid and folds the borders with four flip/add torch ops (:188-199); here dX comes out
of one kernel.  fp32 / fp64 like the reference, plus bf16 / fp16.
"""
import torch
from torch.autograd import Function
from torch.nn.modules.utils import _pair

from . import _lib
from .aggregation_zeropad import _desc, _out_hw


class AggregationRefpad(Function):
    @staticmethod
    def forward(ctx, input, weight, kernel_size, stride, padding, dilation):
        kernel_size, stride, padding, dilation = _pair(kernel_size), _pair(stride), _pair(padding), _pair(dilation)
        ctx.cfg = (kernel_size, stride, padding, dilation)
        assert input.dim() == 4 and input.is_cuda and weight.is_cuda
        batch_size, input_channels, input_height, input_width = input.size()
        _, weight_heads, weight_channels, weight_kernels, weight_height, weight_width = weight.size()
        output_height, output_width = _out_hw(input_height, input_width, kernel_size, stride, padding, dilation)
        assert output_height * output_width == weight_height * weight_width
        input, weight = input.detach().contiguous(), weight.detach().contiguous()
        output = input.new_empty((batch_size, weight_heads * input_channels, output_height, output_width))
        dsc = _desc(input, weight, kernel_size, stride, padding, dilation, output_height, output_width, _lib.NCHW)
        if output.numel():
            with torch.cuda.device_of(input):
                rc = _lib.load().cotb200_agg_refpad_fwd(dsc, input.data_ptr(), weight.data_ptr(), output.data_ptr(),
                                                        _lib.stream_ptr(input))
            _lib.check(rc, "agg_refpad_fwd")
        ctx.save_for_backward(input, weight)
        ctx.out_hw = (output_height, output_width)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        kernel_size, stride, padding, dilation = ctx.cfg
        input, weight = ctx.saved_tensors
        assert grad_output.is_cuda
        grad_output = grad_output.contiguous()
        grad_input = torch.empty_like(input) if ctx.needs_input_grad[0] else None
        grad_weight = torch.empty_like(weight) if ctx.needs_input_grad[1] else None
        if (grad_input is not None or grad_weight is not None) and grad_output.numel():
            dsc = _desc(input, weight, kernel_size, stride, padding, dilation, ctx.out_hw[0], ctx.out_hw[1], _lib.NCHW)
            with torch.cuda.device_of(input):
                rc = _lib.load().cotb200_agg_refpad_bwd(dsc, grad_output.data_ptr(), input.data_ptr(), weight.data_ptr(),
                                                        _lib.ptr(grad_input), _lib.ptr(grad_weight), _lib.stream_ptr(input))
            _lib.check(rc, "agg_refpad_bwd")
        return grad_input, grad_weight, None, None, None, None


def aggregation_refpad(input, weight, kernel_size=3, stride=1, padding=0, dilation=1):
    assert input.shape[0] == weight.shape[0] and (input.shape[1] % weight.shape[2] == 0)
    if input.is_cuda:
        out = AggregationRefpad.apply(input, weight, kernel_size, stride, padding, dilation)
    else:
        if not torch.cuda.is_available():
            raise RuntimeError("cotb200 aggregation_refpad: no CUDA device (there is no CPU implementation)")
        out = AggregationRefpad.apply(input.cuda(), weight.cuda(), kernel_size, stride, padding, dilation)
        torch.cuda.synchronize()
        out = out.cpu()
    return out
