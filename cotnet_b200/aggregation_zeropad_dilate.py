"""Drop-in for the reference's ``cupy_layers/aggregation_zeropad_dilate.py`` (3x3 LocalConv, per-weight-channel dilation).

    AggregationZeropadDilate.apply(input, weight, dilation, kernel_size, stride)    /root/reference/cupy_layers/aggregation_zeropad_dilate.py:148-218
    aggregation_zeropad_dilate(input, weight, dilation, kernel_size=3, stride=1)    :221-231
    LocalConvolutionDilate(in_channels, out_channels, kernel_size, stride=1)        :233-256

``dilation`` is a tensor of ``weight_channels`` values in the input's dtype (the reference indexes it with
``c % weight_channels`` and truncates to int, :31-33); padding == dilation, output size == input size.
"""
import torch
from torch import Tensor
from torch.autograd import Function
from torch.nn.modules.utils import _pair

from . import _lib
from .aggregation_zeropad import _desc


class AggregationZeropadDilate(Function):
    @staticmethod
    def forward(ctx, input, weight, dilation, kernel_size, stride):
        kernel_size, stride = _pair(kernel_size), _pair(stride)
        ctx.kernel_size, ctx.stride = kernel_size, stride
        assert input.dim() == 4 and input.is_cuda and weight.is_cuda and dilation.is_cuda
        batch_size, input_channels, input_height, input_width = input.size()
        _, weight_heads, weight_channels, weight_kernels, weight_height, weight_width = weight.size()
        output_height, output_width = input_height, input_width
        assert output_height * output_width == weight_height * weight_width
        input, weight = input.detach().contiguous(), weight.detach().contiguous()
        dilation = dilation.detach().to(input.dtype).contiguous()
        output = input.new_empty((batch_size, weight_heads * input_channels, output_height, output_width))
        dsc = _desc(input, weight, kernel_size, (1, 1), (0, 0), (1, 1), output_height, output_width, _lib.NCHW)
        if output.numel():
            with torch.cuda.device_of(input):
                rc = _lib.load().cotb200_agg_zeropad_dilate_fwd(dsc, input.data_ptr(), weight.data_ptr(), dilation.data_ptr(),
                                                                output.data_ptr(), _lib.stream_ptr(input))
            _lib.check(rc, "agg_zeropad_dilate_fwd")
        ctx.save_for_backward(input, weight, dilation)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        input, weight, dilation = ctx.saved_tensors
        assert grad_output.is_cuda
        grad_output = grad_output.contiguous()
        grad_input = torch.empty_like(input) if ctx.needs_input_grad[0] else None
        grad_weight = torch.empty_like(weight) if ctx.needs_input_grad[1] else None
        if (grad_input is not None or grad_weight is not None) and grad_output.numel():
            dsc = _desc(input, weight, ctx.kernel_size, (1, 1), (0, 0), (1, 1), input.shape[2], input.shape[3], _lib.NCHW)
            with torch.cuda.device_of(input):
                rc = _lib.load().cotb200_agg_zeropad_dilate_bwd(dsc, grad_output.data_ptr(), input.data_ptr(), weight.data_ptr(),
                                                                dilation.data_ptr(), _lib.ptr(grad_input), _lib.ptr(grad_weight),
                                                                _lib.stream_ptr(input))
            _lib.check(rc, "agg_zeropad_dilate_bwd")
        return grad_input, grad_weight, None, None, None


def aggregation_zeropad_dilate(input, weight, dilation, kernel_size=3, stride=1):
    assert (input.shape[0] == weight.shape[0]) and (input.shape[1] % weight.shape[2] == 0) and (dilation.shape[0] == weight.shape[2])
    if input.is_cuda:
        out = AggregationZeropadDilate.apply(input, weight, dilation, kernel_size, stride)
    else:
        if not torch.cuda.is_available():
            raise RuntimeError("cotb200 aggregation_zeropad_dilate: no CUDA device (there is no CPU implementation)")
        out = AggregationZeropadDilate.apply(input.cuda(), weight.cuda(), dilation.cuda(), kernel_size, stride)
        torch.cuda.synchronize()
        out = out.cpu()
    return out


class LocalConvolutionDilate(torch.nn.Module):
    def __init__(self, in_channels: int, out_channels: int, kernel_size: int, stride: int = 1):
        super(LocalConvolutionDilate, self).__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = kernel_size
        self.stride = stride
        assert kernel_size == 3

    def forward(self, input: Tensor, weight: Tensor, dilation: Tensor):
        return aggregation_zeropad_dilate(input, weight, dilation, kernel_size=self.kernel_size, stride=self.stride)
