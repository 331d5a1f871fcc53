// LocalConv 3x3 forward on NHWC tensors, third generation: persistent CTAs fed by a TMA pipeline (sm_100a).
//
// ncu of the second generation (profiles/r01_agg_v2_ncu.md) shows the register-resident kernels are LATENCY bound:
// 105 registers -> 2 CTAs/SM, 8 resident warps, every warp stalled on `long_scoreboard` (global loads), DRAM at 18 %.
// This kernel decouples the memory stream from the math:
//   * warp 0 is a TMA producer: per tile it issues cp.async.bulk.tensor loads of the input band WITH its halo
//     (4-D box {128 B of channels, W+2, TH+2, 1} fetched at (c0, -1, h0-1, n): out-of-bounds coordinates are
//     zero-filled by the TMA unit == the operator's zero padding) and of the weight band, into a multi-stage
//     shared-memory ring guarded by mbarriers (complete_tx);
//   * the other warps consume: one thread per (pixel, 16-byte channel packet), operands read from shared memory
//     with conflict-free 16-byte accesses (input rows are 128B-swizzled by TMA; the 8 lanes sharing a weight packet
//     broadcast), bf16/fp16 through the mixed-precision FMA (FHFMA), result stored straight to global, coalesced;
//   * CTAs are persistent (grid = resident CTAs), tiles = (sample, band of TH image rows) taken round-robin.
// Weight layout: COTB200_NHWC_TAP (tap-major chunks of gc weight channels) -- the block-internal layout.
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"

namespace cotb200 {

template <typename T> struct MixT;
template <> struct MixT<float> {
  __device__ __forceinline__ static float fma(float a, float b, float c) { return fmaf(a, b, c); }
};
template <> struct MixT<__nv_bfloat16> {
  __device__ __forceinline__ static float fma(__nv_bfloat16 a, __nv_bfloat16 b, float c) {
    float d;
    asm("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(d) : "h"(__bfloat16_as_ushort(a)), "h"(__bfloat16_as_ushort(b)), "f"(c));
    return d;
  }
};
template <> struct MixT<__half> {
  __device__ __forceinline__ static float fma(__half a, __half b, float c) {
    float d;
    asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"(__half_as_ushort(a)), "h"(__half_as_ushort(b)), "f"(c));
    return d;
  }
};

static constexpr int AT_MAX_STAGES = 4;

struct AggTmaP {
  int N, C, H, W, wc, Cf, wcf, gc, J;
  int TH;                 // image rows per tile
  int slabs;              // 128-byte channel slabs per pixel = C * sizeof(T) / 128
  int jboxes, jbox;       // weight row split into jboxes TMA boxes of jbox elements (<= 256)
  int stages;
  int slab_bytes;         // (TH+2)*(W+2)*128 rounded up to 1024
  int x_bytes_tx;         // bytes TMA reports for the input boxes of one stage
  int w_stage_bytes;      // TH*W*J*sizeof(T) rounded up to 128
  int w_bytes_tx;
  int stage_bytes;
  int bands;              // ceil(H / TH)
  int total_tiles;
  long long y_sn;         // output batch stride (elements)
  int y_sp;               // output pixel stride (elements)
};

__device__ __forceinline__ uint32_t at_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void at_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void at_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void at_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void at_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (int spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && spin > (1 << 22)) __trap();
  }
}
__device__ __forceinline__ void at_tma_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

template <typename T, int VEC>
__device__ __forceinline__ Pack<T, VEC> lds_pack(const uint8_t* p) {
  Pack<T, VEC> r;
  *reinterpret_cast<uint4*>(&r) = *reinterpret_cast<const uint4*>(p);
  return r;
}

template <typename T>
__global__ void __launch_bounds__(1024, 1)
agg3_fwd_tma_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW, T* __restrict__ y,
                    const AggTmaP p) {
  constexpr int VEC = 16 / (int)sizeof(T);
  extern __shared__ __align__(1024) uint8_t at_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t s_full[AT_MAX_STAGES], s_empty[AT_MAX_STAGES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncw = (blockDim.x >> 5) - 1;               // consumer warps
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
    for (int s = 0; s < p.stages; ++s) { at_mbar_init(at_smem_u32(&s_full[s]), 1); at_mbar_init(at_smem_u32(&s_empty[s]), ncw); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1;
        at_mbar_wait(at_smem_u32(&s_empty[s]), ph ^ 1);
        const int n = tile / p.bands, h0 = (tile - n * p.bands) * p.TH;
        const uint32_t full = at_smem_u32(&s_full[s]);
        const uint32_t base = at_smem_u32(smem + (size_t)s * p.stage_bytes);
        at_mbar_expect_tx(full, (uint32_t)(p.x_bytes_tx + p.w_bytes_tx));
        for (int sl = 0; sl < p.slabs; ++sl)      // input band + halo; OOB (w = -1 / W, h = -1 / H) zero-filled
          at_tma_4d(base + sl * p.slab_bytes, &mapX, full, sl * (128 / (int)sizeof(T)), -1, h0 - 1, n);
        const uint32_t wbase = base + p.slabs * p.slab_bytes;
        for (int jb = 0; jb < p.jboxes; ++jb)
          at_tma_4d(wbase + jb * (p.TH * p.W * p.jbox * (int)sizeof(T)), &mapW, full, jb * p.jbox, 0, h0, n);
      }
    }
  } else {
    // ===================== consumers: thread = (pixel of the band, 16-byte channel packet) =====================
    const int ct = threadIdx.x - 32, nct = ncw * 32;
    const int CQ = p.C / VEC;
    const int items = p.TH * p.W * CQ;
    const int Wp = p.W + 2;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (it / p.stages) & 1;
      at_mbar_wait(at_smem_u32(&s_full[s]), ph);
      const int n = tile / p.bands, h0 = (tile - n * p.bands) * p.TH;
      const uint8_t* xs = smem + (size_t)s * p.stage_bytes;
      const uint8_t* ws = xs + (size_t)p.slabs * p.slab_bytes;
      for (int item = ct; item < items; item += nct) {
        const int q = item % CQ, px = item / CQ;
        const int hl = px / p.W, wl = px - hl * p.W;
        if (h0 + hl >= p.H) continue;
        const int c0 = q * VEC;
        const int g0 = (c0 / p.Cf) * p.wcf + (c0 % p.Cf) % p.wcf;
        const int cb = c0 * (int)sizeof(T);                 // byte offset of the packet inside the pixel
        const int slab = cb >> 7, chunk = (cb >> 4) & 7;
        // weights: [jb][TH*W][jbox] boxes; packet of tap t at element ((g0/gc)*9 + t)*gc + g0%gc of the pixel's J row
        const int e0 = (g0 / p.gc) * 9 * p.gc + g0 % p.gc;
        float acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int e = e0 + t * p.gc;
          const int jb = e / p.jbox, ej = e - jb * p.jbox;
          const Pack<T, VEC> wv = lds_pack<T, VEC>(ws + ((size_t)(jb * p.TH * p.W + px) * p.jbox + ej) * sizeof(T));
          const int r = (hl + t / 3) * Wp + (wl + t % 3);   // row of the haloed band: (hl + dh + 1, wl + dw + 1)
          const Pack<T, VEC> xv = lds_pack<T, VEC>(xs + (size_t)slab * p.slab_bytes + r * 128 + ((chunk ^ (r & 7)) << 4));
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc[i] = MixT<T>::fma(wv.v[i], xv.v[i], acc[i]);
        }
        Pack<T, VEC> o;
#pragma unroll
        for (int i = 0; i < VEC; ++i) o.v[i] = Elem<T>::from(acc[i]);
        st_pack<T, VEC>(y + n * p.y_sn + (long long)((h0 + hl) * p.W + wl) * p.y_sp + c0, o);
      }
      __syncwarp();
      if (lane == 0) at_mbar_arrive(at_smem_u32(&s_empty[s]));   // this warp is done reading the stage
    }
  }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*AtEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static AtEncodeFn at_encode_fn() {
  static AtEncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (AtEncodeFn)p;
  }
  return fn;
}

template <typename T> static CUtensorMapDataType at_dtype();
template <> CUtensorMapDataType at_dtype<float>() { return CU_TENSOR_MAP_DATA_TYPE_FLOAT32; }
template <> CUtensorMapDataType at_dtype<__nv_bfloat16>() { return CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; }
template <> CUtensorMapDataType at_dtype<__half>() { return CU_TENSOR_MAP_DATA_TYPE_FLOAT16; }

// NHWC tensor [N,H,W,Cdim] with pixel pitch sp / batch pitch sn (elements); box {b0, b1, b2, 1}
template <typename T>
static bool at_make_map(CUtensorMap* m, const void* base, int N, int H, int W, int Cdim, long long sp, long long sn, int b0, int b1,
                        int b2, bool swizzle) {
  AtEncodeFn enc = at_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[4] = {(cuuint64_t)Cdim, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)sp * sizeof(T), (cuuint64_t)sp * sizeof(T) * W, (cuuint64_t)sn * sizeof(T)};
  cuuint32_t box[4] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  return enc(m, at_dtype<T>(), 4, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct Nhwc2Args {
  int N, C, H, W, wc, fold, layout, gc, dtype;
  long long x_sn, x_sp, w_sn, w_sp, y_sn, y_sp;
};

// returns 1 if handled (rc in *rc), 0 if the caller should use the register-resident kernels
template <typename T>
int agg_tma_fwd(const Nhwc2Args& a, const T* x, const T* w, T* y, cudaStream_t st, int* rc) {
  if constexpr (std::is_same<T, double>::value) { return 0; } else {
    constexpr int VEC = 16 / (int)sizeof(T);
    static int disabled = -1;
    if (disabled < 0) { const char* e = getenv("COTB200_AGG_TMA"); disabled = (e && e[0] == '0') ? 1 : 0; }
    if (disabled) return 0;
    if (a.layout != COTB200_NHWC_TAP) return 0;
    const int Cf = a.C / a.fold, wcf = a.wc / a.fold;
    if ((a.C * (int)sizeof(T)) % 128 || wcf % VEC || a.gc < VEC || a.gc % VEC || a.wc % a.gc) return 0;
    if (((uintptr_t)x | (uintptr_t)w | (uintptr_t)y) & 15) return 0;
    if ((a.x_sp * sizeof(T)) % 16 || (a.w_sp * sizeof(T)) % 16 || (a.y_sp * sizeof(T)) % 16 || (a.x_sn * sizeof(T)) % 16 ||
        (a.w_sn * sizeof(T)) % 16 || (a.y_sn * sizeof(T)) % 16) return 0;
    if (a.W + 2 > 256 || a.y_sp > 2147483647LL) return 0;
    AggTmaP p{};
    p.N = a.N; p.C = a.C; p.H = a.H; p.W = a.W; p.wc = a.wc; p.Cf = Cf; p.wcf = wcf; p.gc = a.gc; p.J = 9 * a.wc;
    p.slabs = a.C * (int)sizeof(T) / 128;
    p.jboxes = (p.J + 255) / 256;
    while (p.J % p.jboxes) ++p.jboxes;
    p.jbox = p.J / p.jboxes;
    if ((p.jbox * (int)sizeof(T)) % 16) return 0;
    p.y_sn = a.y_sn; p.y_sp = (int)a.y_sp;
    // choose TH: about 100 KB per stage at most, at least ~512 work items, bands tile H as evenly as possible
    const int CQ = a.C / VEC;
    int best_th = 0;
    for (int th = 1; th <= a.H && th <= 254; ++th) {
      const long long xb = (long long)p.slabs * ((((long long)(th + 2) * (a.W + 2) * 128) + 1023) / 1024 * 1024);
      const long long wb = (((long long)th * a.W * p.J * sizeof(T)) + 127) / 128 * 128;
      if (xb + wb > 72 * 1024) break;
      best_th = th;
      if ((long long)th * a.W * CQ >= 896) break;
    }
    if (!best_th) return 0;
    p.TH = best_th;
    p.slab_bytes = (int)((((long long)(p.TH + 2) * (a.W + 2) * 128) + 1023) / 1024 * 1024);
    p.x_bytes_tx = p.slabs * (p.TH + 2) * (a.W + 2) * 128;
    p.w_bytes_tx = p.TH * a.W * p.J * (int)sizeof(T);
    p.w_stage_bytes = (p.w_bytes_tx + 1023) / 1024 * 1024;
    p.stage_bytes = p.slabs * p.slab_bytes + p.w_stage_bytes;
    p.stages = (int)((200 * 1024) / p.stage_bytes);
    if (p.stages > AT_MAX_STAGES) p.stages = AT_MAX_STAGES;
    if (p.stages < 2) return 0;
    if (p.jboxes > 1 && ((long long)p.TH * a.W * p.jbox * sizeof(T)) % 128) return 0;   // TMA destinations are 128-byte aligned
    p.bands = (a.H + p.TH - 1) / p.TH;
    p.total_tiles = a.N * p.bands;
    CUtensorMap mx, mw;
    if (!at_make_map<T>(&mx, x, a.N, a.H, a.W, a.C, a.x_sp, a.x_sn, 128 / (int)sizeof(T), a.W + 2, p.TH + 2, true)) return 0;
    if (!at_make_map<T>(&mw, w, a.N, a.H, a.W, p.J, a.w_sp, a.w_sn, p.jbox, a.W, p.TH, false)) return 0;
    const int items = p.TH * a.W * CQ;
    int cw = (items + 31) / 32;                          // consumer warps
    if (cw > 28) cw = 28;
    if (cw < 4) cw = 4;
    const int threads = (cw + 1) * 32;
    const int smem = p.stages * p.stage_bytes + 1024;
    static int configured = 0;
    if (configured < smem) {
      cudaError_t e = cudaFuncSetAttribute(agg3_fwd_tma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
      if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(agg3_fwd_tma): %s", cudaGetErrorString(e)); *rc = (int)e; return 1; }
      configured = 220 * 1024;
    }
    int grid = num_sms();
    if (grid > p.total_tiles) grid = p.total_tiles;
    COTB200_PROF("agg3_fwd_tma");
    agg3_fwd_tma_kernel<T><<<grid, threads, smem, st>>>(mx, mw, y, p);
    *rc = check_launch("agg3_fwd_tma");
    return 1;
  }
}

#define COTB200_INST3(T) template int agg_tma_fwd<T>(const Nhwc2Args&, const T*, const T*, T*, cudaStream_t, int*);
COTB200_INST3(float) COTB200_INST3(double) COTB200_INST3(__nv_bfloat16) COTB200_INST3(__half)

}  // namespace cotb200
