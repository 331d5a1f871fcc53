"""Drop-in for the reference's ``cupy_layers/aggregation_zeropad_mix_merge.py`` (3x3 + 5x5 LocalConv, packed weights).

    AggregationZeropadMixMerge.apply(input, weight, head_num, w_channels, k1, k2, stride, p1, p2, dilation)   /root/reference/cupy_layers/aggregation_zeropad_mix_merge.py:180-274
    aggregation_zeropad_mix_merge(...)          :276-287
    LocalConvolutionMixMerge(...)               :289-330

``weight`` is [N, head_num*w_channels*(k1^2 + k2^2), Ho, Wo]: the two weight sets of the mix op concatenated along the
channel axis (:35-36,:52-54).  The kernels read / write the packed tensor in place.
"""
import torch
from torch import Tensor
from torch.autograd import Function
from torch.nn.modules.utils import _pair

from . import _lib
from .aggregation_zeropad import _out_hw


def _merge_desc(input, head_num, w_channels, k1, s, p1, d, Ho, Wo):
    dsc = _lib.AggDesc()
    dsc.n, dsc.c, dsc.h, dsc.w = input.shape
    dsc.heads, dsc.wc = head_num, w_channels
    dsc.kh, dsc.kw = k1
    dsc.sh, dsc.sw = s
    dsc.ph, dsc.pw = p1
    dsc.dh, dsc.dw = d
    dsc.ho, dsc.wo = Ho, Wo
    dsc.dtype = _lib.dtype_code(input)
    dsc.layout = _lib.NCHW
    dsc.fold = 1
    return dsc


class AggregationZeropadMixMerge(Function):
    @staticmethod
    def forward(ctx, input, weight, head_num, w_channels, kernel_size1, kernel_size2, stride, padding1, padding2, dilation):
        kernel_size1, kernel_size2, stride = _pair(kernel_size1), _pair(kernel_size2), _pair(stride)
        padding1, padding2, dilation = _pair(padding1), _pair(padding2), _pair(dilation)
        ctx.cfg = (head_num, w_channels, kernel_size1, kernel_size2, stride, padding1, padding2, dilation)
        assert input.dim() == 4 and input.is_cuda and weight.is_cuda
        batch_size, input_channels, input_height, input_width = input.size()
        weight_height, weight_width = weight.size()[-2], weight.size()[-1]
        output_height, output_width = _out_hw(input_height, input_width, kernel_size1, stride, padding1, dilation)
        assert output_height * output_width == weight_height * weight_width
        input, weight = input.detach().contiguous(), weight.detach().contiguous()
        output = input.new_empty((batch_size, head_num * input_channels * 2, output_height, output_width))
        dsc = _merge_desc(input, head_num, w_channels, kernel_size1, stride, padding1, dilation, output_height, output_width)
        if output.numel():
            with torch.cuda.device_of(input):
                rc = _lib.load().cotb200_agg_zeropad_mix_merge_fwd(dsc, kernel_size2[0], kernel_size2[1], padding2[0], padding2[1],
                                                                   input.data_ptr(), weight.data_ptr(), output.data_ptr(),
                                                                   _lib.stream_ptr(input))
            _lib.check(rc, "agg_zeropad_mix_merge_fwd")
        ctx.save_for_backward(input, weight)
        ctx.out_hw = (output_height, output_width)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        head_num, w_channels, kernel_size1, kernel_size2, stride, padding1, padding2, dilation = ctx.cfg
        input, weight = ctx.saved_tensors
        assert grad_output.is_cuda
        grad_output = grad_output.contiguous()
        grad_input = torch.empty_like(input) if ctx.needs_input_grad[0] else None
        grad_weight = torch.empty_like(weight) if ctx.needs_input_grad[1] else None
        if (grad_input is not None or grad_weight is not None) and grad_output.numel():
            dsc = _merge_desc(input, head_num, w_channels, kernel_size1, stride, padding1, dilation, ctx.out_hw[0], ctx.out_hw[1])
            with torch.cuda.device_of(input):
                rc = _lib.load().cotb200_agg_zeropad_mix_merge_bwd(dsc, kernel_size2[0], kernel_size2[1], padding2[0], padding2[1],
                                                                   grad_output.data_ptr(), input.data_ptr(), weight.data_ptr(),
                                                                   _lib.ptr(grad_input), _lib.ptr(grad_weight), _lib.stream_ptr(input))
            _lib.check(rc, "agg_zeropad_mix_merge_bwd")
        return grad_input, grad_weight, None, None, None, None, None, None, None, None


def aggregation_zeropad_mix_merge(input, weight, head_num, w_channels, kernel_size1=3, kernel_size2=5, stride=1, padding1=0,
                                  padding2=0, dilation=1):
    assert input.shape[0] == weight.shape[0] and (input.shape[1] % w_channels == 0)
    assert weight.shape[1] == head_num * w_channels * (kernel_size1 * kernel_size1 + kernel_size2 * kernel_size2)
    if input.is_cuda:
        out = AggregationZeropadMixMerge.apply(input, weight, head_num, w_channels, kernel_size1, kernel_size2, stride,
                                               padding1, padding2, dilation)
    else:
        if not torch.cuda.is_available():
            raise RuntimeError("cotb200 aggregation_zeropad_mix_merge: no CUDA device (there is no CPU implementation)")
        out = AggregationZeropadMixMerge.apply(input.cuda(), weight.cuda(), head_num, w_channels, kernel_size1, kernel_size2,
                                               stride, padding1, padding2, dilation)
        torch.cuda.synchronize()
        out = out.cpu()
    return out


class LocalConvolutionMixMerge(torch.nn.Module):
    def __init__(self, in_channels: int, out_channels: int, head_num: int, w_channels: int, kernel_size1: int,
                 kernel_size2: int, stride: int = 1, padding1: int = 0, padding2: int = 0, dilation: int = 1, pad_mode: int = 0):
        super(LocalConvolutionMixMerge, self).__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.head_num = head_num
        self.w_channels = w_channels
        self.kernel_size1 = kernel_size1
        self.kernel_size2 = kernel_size2
        self.stride = stride
        self.padding1 = padding1
        self.padding2 = padding2
        self.dilation = dilation
        self.pad_mode = pad_mode

    def forward(self, input: Tensor, weight: Tensor):
        return aggregation_zeropad_mix_merge(input, weight, head_num=self.head_num, w_channels=self.w_channels,
                                             kernel_size1=self.kernel_size1, kernel_size2=self.kernel_size2, stride=self.stride,
                                             padding1=self.padding1, padding2=self.padding2, dilation=self.dilation)
