/*
 * cotb200.h -- C ABI of libcotb200.so: the B200-native (sm_100a) kernels of the CoT-block hot path.
 *
 * This is the drop-in boundary.  The reference has no C ABI: its "FFI" is CuPy --
 *   load_kernel(name, code, **literals) -> cupy.cuda.compile_with_cache(code).get_function(name)
 *                                                       (/root/reference/cupy_layers/utils.py:14-18)
 *   f(block=(1024,1,1), grid=(GET_BLOCKS(n),1,1), args=[ptr,...], stream=Stream(ptr=current_stream))
 *                                                       (/root/reference/cupy_layers/aggregation_zeropad.py:130-143)
 * with every dimension baked into the NVRTC source.  Each entry point below names the reference launch it
 * replaces; dimensions are runtime arguments (no per-shape JIT), pointers are raw device pointers, the
 * stream is a cudaStream_t passed as void*.  No torch types, no allocation, no hidden synchronisation:
 * every call is asynchronous on the given stream and safe under CUDA-graph capture.
 *
 * Return value: 0 on success; >0 = cudaError_t from the launch; <0 = COTB200_E* argument error.
 * cotb200_last_error() returns a thread-local, human-readable description of the last failure.
 */
#ifndef COTB200_H_
#define COTB200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define COTB200_VERSION 100

/* element types (the reference supports only float/double: cupy_layers/utils.py:8-12) */
enum { COTB200_F32 = 0, COTB200_F64 = 1, COTB200_BF16 = 2, COTB200_F16 = 3 };

/* memory layouts of activations / weights
 *   NCHW  : x [N,C,H,W], w [N,heads,wc,kh*kw,Ho,Wo], y [N,heads*C,Ho,Wo]   -- the reference contract
 *           (cupy_layers/aggregation_zeropad.py:37-38)
 *   NHWC  : x [N,H,W,C], w [N,Ho,Wo,heads,wc,kh*kw], y [N,Ho,Wo,heads*C]   -- torch channels_last
 *   NHWC_TAP : as NHWC but weight channels are stored tap-major inside chunks of `gc` weight channels:
 *           index of (head, g, tap) = head*wc*K2 + ((g/gc)*K2 + tap)*gc + g%gc.  This is the layout the fused
 *           CoT block uses internally (the logits GEMM emits it for free by permuting its weight rows). */
enum { COTB200_NCHW = 0, COTB200_NHWC = 1, COTB200_NHWC_TAP = 2 };

enum {
  COTB200_EINVAL = -1,      /* inconsistent / unsupported dimensions */
  COTB200_EDTYPE = -2,      /* unknown dtype */
  COTB200_ELAYOUT = -3,     /* unknown layout */
  COTB200_EALIGN = -4,      /* pointer / stride alignment required by the selected kernel not met */
  COTB200_ENULL = -5,       /* required pointer is NULL */
  COTB200_ETOOBIG = -6      /* tensor exceeds the 2^31-element index range of the fast kernels */
};

/* Geometry of one LocalConv call.  Mirrors the literals substituted into the reference kernels
 * (cupy_layers/aggregation_zeropad.py:131-139). */
typedef struct cotb200_agg_desc {
  int n, c, h, w;            /* input  [n, c, h, w] */
  int heads, wc;             /* weight heads, weight channels; c % wc == 0; input channel ch uses weight channel ch % wc */
  int kh, kw;                /* kernel size */
  int sh, sw, ph, pw, dh, dw;/* stride, zero padding, dilation */
  int ho, wo;                /* output spatial size = weight spatial size */
  int dtype;                 /* COTB200_F32 ... */
  int layout;                /* COTB200_NCHW / NHWC / NHWC_TAP */
  int gc;                    /* NHWC_TAP only: weight-channel chunk width (divides wc); ignored otherwise */
  int fold;                  /* CoXt channel fold (models/cotnet.py:157-162): 0/1 = none; F>1: channel c uses weight
                              * channel (c/(C/F))*(wc/F) + (c%(C/F))%(wc/F), i.e. the F channel groups own disjoint
                              * wc/F-wide weight ranges -- the un-folded form of the reference's view(B*F, ...) trick */
  /* NHWC layouts only: element strides of the batch and pixel dimensions (channel stride is 1).
   * 0 selects the dense default.  They let the CoXt "fold the two channel halves into the batch"
   * view (models/cotnet.py:157-162) run without a copy. */
  long long x_sn, x_sp;
  long long w_sn, w_sp;
  long long y_sn, y_sp;
} cotb200_agg_desc;

int cotb200_version(void);
const char* cotb200_last_error(void);
/* number of kernels launched by this library in this process (the bench's gpu_launches counter) */
long long cotb200_launch_count(void);

/* Per-kernel device timing for the bench's roofline line: when enabled every launch of this library is
 * bracketed by CUDA events on its stream; cotb200_prof_report writes "<kernel> <launches> <total_ms>" lines
 * (returns the length needed).  Off by default; do not enable during CUDA-graph capture. */
void cotb200_prof_enable(int on);
int cotb200_prof_report(char* buf, int len);

/* Replaces aggregation_zeropad_forward_kernel (cupy_layers/aggregation_zeropad.py:20-46, launch :130-143).
 * y[n, head*C + c, ho, wo] = sum_{kh,kw} w[n, head, c % wc, kh*KW+kw, ho, wo] * x[n, c, ho*s-p+kh*d, wo*s-p+kw*d] */
int cotb200_agg_zeropad_fwd(const cotb200_agg_desc* d, const void* x, const void* w, void* y, void* stream);

/* Replaces aggregation_zeropad_input_backward_kernel (:48-79) and ..._weight_backward_kernel (:81-110),
 * launches :168-185.  dx and/or dw may be NULL (ctx.needs_input_grad, :168,:177); when both are wanted
 * and the fast path applies they are produced by ONE fused kernel that reads dy once. */
int cotb200_agg_zeropad_bwd(const cotb200_agg_desc* d, const void* dy, const void* x, const void* w,
                            void* dx, void* dw, void* stream);

/* Replaces aggregation_zeropad_mix_forward_kernel (cupy_layers/aggregation_zeropad_mix.py:20-74).
 * d describes the FIRST kernel (kh,kw,ph,pw = kernel_size1/padding1); the second kernel is (k2h,k2w,p2h,p2w).
 * y = cat_channels[ agg(x, w1; k1,p1), agg(x, w2; k2,p2) ]  -> [n, 2*heads*c, ho, wo].  NCHW layout only. */
int cotb200_agg_zeropad_mix_fwd(const cotb200_agg_desc* d, int k2h, int k2w, int p2h, int p2w,
                                const void* x, const void* w1, const void* w2, void* y, void* stream);

/* Replaces aggregation_zeropad_mix_{input,weight}_backward_kernel (aggregation_zeropad_mix.py:76-207).
 * dx, or dw1 and dw2 together, may be NULL. */
int cotb200_agg_zeropad_mix_bwd(const cotb200_agg_desc* d, int k2h, int k2w, int p2h, int p2w,
                                const void* dy, const void* x, const void* w1, const void* w2,
                                void* dx, void* dw1, void* dw2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COTB200_H_ */
