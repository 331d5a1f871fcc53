"""Drop-in for the reference's ``cupy_layers/aggregation_zeropad_mix.py`` (3x3 + 5x5 LocalConv).

    AggregationZeropadMix.apply(...)         /root/reference/cupy_layers/aggregation_zeropad_mix.py:209-290
    aggregation_zeropad_mix(...)             :292-302
    LocalConvolutionMix(...)                 :304-342

Output channel order is the reference's: ``[n, (kernel_idx*heads + head)*C + c]`` (:26-31, :220).
The reference's dX kernel only visits head 0 (:88); here dX is the full gradient for any ``heads``
(identical for heads == 1, the only configuration the reference's self-test and callers use).
"""
import torch
from torch import Tensor
from torch.autograd import Function
from torch.nn.modules.utils import _pair

from . import _lib
from .aggregation_zeropad import _desc, _out_hw


class AggregationZeropadMix(Function):
    @staticmethod
    def forward(ctx, input, weight1, weight2, kernel_size1, kernel_size2, stride, padding1, padding2, dilation):
        kernel_size1, kernel_size2, stride = _pair(kernel_size1), _pair(kernel_size2), _pair(stride)
        padding1, padding2, dilation = _pair(padding1), _pair(padding2), _pair(dilation)
        ctx.cfg = (kernel_size1, kernel_size2, stride, padding1, padding2, dilation)
        assert input.dim() == 4 and input.is_cuda and weight1.is_cuda and weight2.is_cuda
        batch_size, input_channels, input_height, input_width = input.size()
        _, weight_heads, weight_channels, weight_kernels, weight_height, weight_width = weight1.size()
        assert weight2.shape[1] == weight_heads and weight2.shape[2] == weight_channels
        output_height, output_width = _out_hw(input_height, input_width, kernel_size1, stride, padding1, dilation)
        assert output_height * output_width == weight_height * weight_width
        assert (output_height, output_width) == _out_hw(input_height, input_width, kernel_size2, stride, padding2, dilation)
        input, weight1, weight2 = input.detach().contiguous(), weight1.detach().contiguous(), weight2.detach().contiguous()
        output = input.new_empty((batch_size, weight_heads * input_channels * 2, output_height, output_width))
        dsc = _desc(input, weight1, kernel_size1, stride, padding1, dilation, output_height, output_width, _lib.NCHW)
        if output.numel():
            with torch.cuda.device_of(input):
                rc = _lib.load().cotb200_agg_zeropad_mix_fwd(
                    dsc, kernel_size2[0], kernel_size2[1], padding2[0], padding2[1], input.data_ptr(), weight1.data_ptr(),
                    weight2.data_ptr(), output.data_ptr(), _lib.stream_ptr(input))
            _lib.check(rc, "agg_zeropad_mix_fwd")
        ctx.save_for_backward(input, weight1, weight2)
        ctx.out_hw = (output_height, output_width)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        kernel_size1, kernel_size2, stride, padding1, padding2, dilation = ctx.cfg
        input, weight1, weight2 = ctx.saved_tensors
        assert grad_output.is_cuda
        grad_output = grad_output.contiguous()
        grad_input = grad_weight1 = grad_weight2 = None
        if ctx.needs_input_grad[0]:
            grad_input = torch.empty_like(input)
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            grad_weight1 = torch.empty_like(weight1)
            grad_weight2 = torch.empty_like(weight2)
        if (grad_input is not None or grad_weight1 is not None) and grad_output.numel():
            dsc = _desc(input, weight1, kernel_size1, stride, padding1, dilation, ctx.out_hw[0], ctx.out_hw[1], _lib.NCHW)
            with torch.cuda.device_of(input):
                rc = _lib.load().cotb200_agg_zeropad_mix_bwd(
                    dsc, kernel_size2[0], kernel_size2[1], padding2[0], padding2[1], grad_output.data_ptr(),
                    input.data_ptr(), weight1.data_ptr(), weight2.data_ptr(), _lib.ptr(grad_input),
                    _lib.ptr(grad_weight1), _lib.ptr(grad_weight2), _lib.stream_ptr(input))
            _lib.check(rc, "agg_zeropad_mix_bwd")
        return grad_input, grad_weight1, grad_weight2, None, None, None, None, None, None


def aggregation_zeropad_mix(input, weight1, weight2, kernel_size1=3, kernel_size2=5, stride=1, padding1=0, padding2=0,
                            dilation=1):
    assert input.shape[0] == weight1.shape[0] and (input.shape[1] % weight1.shape[2] == 0)
    assert input.shape[0] == weight2.shape[0] and (input.shape[1] % weight2.shape[2] == 0)
    if input.is_cuda:
        out = AggregationZeropadMix.apply(input, weight1, weight2, kernel_size1, kernel_size2, stride, padding1,
                                          padding2, dilation)
    else:
        if not torch.cuda.is_available():
            raise RuntimeError("cotb200 aggregation_zeropad_mix: no CUDA device (there is no CPU implementation)")
        out = AggregationZeropadMix.apply(input.cuda(), weight1.cuda(), weight2.cuda(), kernel_size1, kernel_size2,
                                          stride, padding1, padding2, dilation)
        torch.cuda.synchronize()
        out = out.cpu()
    return out


class LocalConvolutionMix(torch.nn.Module):
    def __init__(self, in_channels: int, out_channels: int, kernel_size1: int, kernel_size2: int, stride: int = 1,
                 padding1: int = 0, padding2: int = 0, dilation: int = 1, pad_mode: int = 0):
        super(LocalConvolutionMix, self).__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size1 = kernel_size1
        self.kernel_size2 = kernel_size2
        self.stride = stride
        self.padding1 = padding1
        self.padding2 = padding2
        self.dilation = dilation
        self.pad_mode = pad_mode
        assert self.kernel_size1 == 3        # aggregation_zeropad_mix.py:328-329
        assert self.kernel_size2 == 5

    def forward(self, input: Tensor, weight1: Tensor, weight2: Tensor):
        return aggregation_zeropad_mix(input, weight1, weight2, kernel_size1=self.kernel_size1,
                                       kernel_size2=self.kernel_size2, stride=self.stride, padding1=self.padding1,
                                       padding2=self.padding2, dilation=self.dilation)
