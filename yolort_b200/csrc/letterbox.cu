// Letterbox pre-processing: aspect-preserving bilinear resize + centred pad, whole batch in one launch.
//
// Replaces YOLOTransform.forward (yolort/models/transform.py:143-221):
//   _resize_image_and_masks :53-97  -> F.interpolate(bilinear, align_corners=False,
//                                      recompute_scale_factor=True), i.e. ATen upsample_bilinear2d
//   batch_images            :297-330 -> new_full(fill) + centred copy
// and fuses the uint8 -> [0,1] conversion of the default loader (yolov5.py:218-228) plus the layout
// change the first convolution wants (space-to-depth NHWC, see conv_sm100.cu / engine.py).
//
// HBM-bound: per image it reads 3*h*w source bytes (each source texel is touched by <= ~4 output
// pixels and stays in L1/L2) and writes the canvas once.  Threads map to consecutive output x so
// both the source reads (consecutive sx) and the destination writes are coalesced.
#include <cmath>

#include "common.cuh"

namespace yb {
namespace {

constexpr int kMaxImagesPerLaunch = 64;
constexpr int kRowsPerBlock = 8;   // s2d rows per CTA (amortises the 1 KB LUT staging)

struct ImgGeom {
  const void* src;
  int src_h, src_w, new_h, new_w, top, left;
  float ratio_h, ratio_w;
};
// Element strides (channel, row, pixel) of the source: planar CHW = (h*w, w, 1); interleaved HWC (what image
// decoders emit) = (1, 3w, 3).  A template parameter, so the planar path keeps its constant-stride addressing.
template <bool kHwc>
struct SrcStrides {
  size_t cs, rs, ps;
  __device__ __forceinline__ explicit SrcStrides(const ImgGeom& g)
      : cs(kHwc ? 1 : static_cast<size_t>(g.src_h) * g.src_w), rs(kHwc ? 3 * static_cast<size_t>(g.src_w) : g.src_w),
        ps(kHwc ? 3 : 1) {}
};
struct BatchGeom {
  ImgGeom img[kMaxImagesPerLaunch];
};

template <typename SrcT>
__device__ __forceinline__ float load_src(const SrcT* p, const float* lut);
template <>
__device__ __forceinline__ float load_src<uint8_t>(const uint8_t* p, const float* lut) {
  return lut[__ldg(p)];   // `lut` points to the shared-memory copy staged by the kernel
}
template <>
__device__ __forceinline__ float load_src<float>(const float* p, const float*) {
  return __ldg(p);
}
template <>
__device__ __forceinline__ float load_src<__half>(const __half* p, const float*) {
  return __half2float(__ldg(p));
}
template <>
__device__ __forceinline__ float load_src<__nv_bfloat16>(const __nv_bfloat16* p, const float*) {
  return __bfloat162float(*p);
}

// Source index / interpolation weight exactly as ATen's area_pixel_compute_source_index +
// guard_index_and_lambda (align_corners=False, no antialias), all in fp32.
__device__ __forceinline__ void src_coord(int dst, float ratio, int size, int& i0, int& i1, float& l1) {
  float real = __fsub_rn(__fmul_rn(ratio, static_cast<float>(dst) + 0.5f), 0.5f);
  if (real < 0.f) real = 0.f;
  int idx = static_cast<int>(real);
  if (idx > size - 1) idx = size - 1;
  float lam = __fsub_rn(real, static_cast<float>(idx));
  lam = fminf(fmaxf(lam, 0.f), 1.f);
  i0 = idx;
  i1 = idx + (idx < size - 1 ? 1 : 0);
  l1 = lam;
}

template <typename SrcT, bool kHwc>
__device__ __forceinline__ void sample_rgb(const ImgGeom& g, const float* lut, int y, int x, float fill,
                                           float (&rgb)[3]) {
  const int yy = y - g.top, xx = x - g.left;
  if (yy < 0 || yy >= g.new_h || xx < 0 || xx >= g.new_w) {
    rgb[0] = rgb[1] = rgb[2] = fill;
    return;
  }
  const SrcStrides<kHwc> st(g);
  if (g.new_h == g.src_h && g.new_w == g.src_w) {
    // identity resize (ratios are exactly 1, all interpolation weights exactly 0/1): plain copy, same bits
    const SrcT* base = static_cast<const SrcT*>(g.src);
    const size_t o = static_cast<size_t>(yy) * st.rs + static_cast<size_t>(xx) * st.ps;
#pragma unroll
    for (int c = 0; c < 3; ++c) rgb[c] = load_src<SrcT>(base + static_cast<size_t>(c) * st.cs + o, lut);
    return;
  }
  int y0, y1, x0, x1;
  float ly, lx;
  src_coord(yy, g.ratio_h, g.src_h, y0, y1, ly);
  src_coord(xx, g.ratio_w, g.src_w, x0, x1, lx);
  const float wy0 = 1.f - ly, wx0 = 1.f - lx;
  const SrcT* base = static_cast<const SrcT*>(g.src);
  const size_t r0 = static_cast<size_t>(y0) * st.rs, r1 = static_cast<size_t>(y1) * st.rs;
  const size_t c0 = static_cast<size_t>(x0) * st.ps, c1 = static_cast<size_t>(x1) * st.ps;
  // Taps with zero weight are not fetched (w*p + 0*q == w*p exactly for finite q): an identity resize
  // (the 640x640 headline case) touches one source texel per output pixel instead of four.
  const bool need_x1 = lx != 0.f, need_y1 = ly != 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const SrcT* p = base + static_cast<size_t>(c) * st.cs;
    const float p00 = load_src<SrcT>(p + r0 + c0, lut);
    const float p01 = need_x1 ? load_src<SrcT>(p + r0 + c1, lut) : p00;
    float bot = 0.f;
    if (need_y1) {
      const float p10 = load_src<SrcT>(p + r1 + c0, lut);
      const float p11 = need_x1 ? load_src<SrcT>(p + r1 + c1, lut) : p10;
      bot = __fadd_rn(__fmul_rn(wx0, p10), __fmul_rn(lx, p11));
    }
    const float top = __fadd_rn(__fmul_rn(wx0, p00), __fmul_rn(lx, p01));
    rgb[c] = __fadd_rn(__fmul_rn(wy0, top), __fmul_rn(ly, bot));
  }
}

template <typename DstT>
__device__ __forceinline__ DstT cvt_out(float v);
template <>
__device__ __forceinline__ float cvt_out<float>(float v) {
  return v;
}
template <>
__device__ __forceinline__ __half cvt_out<__half>(float v) {
  return __float2half_rn(v);
}
template <>
__device__ __forceinline__ __nv_bfloat16 cvt_out<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// NCHW destination (reference layout): thread per (y, x), three planes.
template <typename SrcT, typename DstT, bool kHwc>
__global__ void letterbox_nchw_kernel(const __grid_constant__ BatchGeom bg, int img0, int Hb, int Wb,
                                      float fill, const float* lut, DstT* __restrict__ dst) {
  __shared__ float s_lut[256];
  if (lut != nullptr) {   // uint8 sources: one coalesced 1 KB read per CTA instead of a global lookup per texel
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = lut[i];
    __syncthreads();
    lut = s_lut;
  }
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int li = blockIdx.z;
  if (x >= Wb) return;
  float rgb[3];
  sample_rgb<SrcT, kHwc>(bg.img[li], lut, y, x, fill, rgb);
  const size_t plane = static_cast<size_t>(Hb) * Wb;
  DstT* o = dst + static_cast<size_t>(img0 + li) * 3 * plane + static_cast<size_t>(y) * Wb + x;
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c * plane] = cvt_out<DstT>(rgb[c]);
}

// Space-to-depth NHWC destination [N, Hb/2, Wb/2, 16]: thread per 2x2 pixel block, one 32-byte store.
template <typename SrcT, typename DstT, bool kHwc>
__global__ void letterbox_s2d_kernel(const __grid_constant__ BatchGeom bg, int img0, int Hb, int Wb,
                                     float fill, const float* lut, DstT* __restrict__ dst) {
  __shared__ float s_lut[256];
  if (lut != nullptr) {   // uint8 sources: one coalesced 1 KB read per CTA instead of a global lookup per texel
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = lut[i];
    __syncthreads();
    lut = s_lut;
  }
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  const int li = blockIdx.z;
  const int W2 = Wb >> 1, H2 = Hb >> 1;
  if (X >= W2) return;
  for (int Y = blockIdx.y * kRowsPerBlock; Y < min(H2, (blockIdx.y + 1) * kRowsPerBlock); ++Y) {
    __align__(16) DstT v[16];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        float rgb[3];
        sample_rgb<SrcT, kHwc>(bg.img[li], lut, 2 * Y + dy, 2 * X + dx, fill, rgb);
        const int q = (dy * 2 + dx) * 4;
        v[q + 0] = cvt_out<DstT>(rgb[0]);
        v[q + 1] = cvt_out<DstT>(rgb[1]);
        v[q + 2] = cvt_out<DstT>(rgb[2]);
        v[q + 3] = cvt_out<DstT>(0.f);
      }
    }
    DstT* o = dst + ((static_cast<size_t>(img0 + li) * H2 + Y) * W2 + X) * 16;
    reinterpret_cast<uint4*>(o)[0] = reinterpret_cast<const uint4*>(v)[0];
    reinterpret_cast<uint4*>(o)[1] = reinterpret_cast<const uint4*>(v)[1];
  }
}

template <typename SrcT, typename DstT, bool kHwc>
int launch_typed(const BatchGeom& bg, int img0, int count, int Hb, int Wb, float fill, const float* lut,
                 void* dst, int layout, cudaStream_t stream) {
  const int threads = 128;
  if (layout == YB_LAYOUT_NCHW) {
    dim3 grid((Wb + threads - 1) / threads, Hb, count);
    letterbox_nchw_kernel<SrcT, DstT, kHwc><<<grid, threads, 0, stream>>>(bg, img0, Hb, Wb, fill, lut,
                                                                    static_cast<DstT*>(dst));
  } else {
    dim3 grid((Wb / 2 + threads - 1) / threads, (Hb / 2 + kRowsPerBlock - 1) / kRowsPerBlock, count);
    letterbox_s2d_kernel<SrcT, DstT, kHwc><<<grid, threads, 0, stream>>>(bg, img0, Hb, Wb, fill, lut,
                                                                   static_cast<DstT*>(dst));
  }
  YB_CHECK_CUDA(cudaGetLastError());
  return YB_OK;
}

template <typename SrcT, bool kHwc>
int launch_src(const BatchGeom& bg, int img0, int count, int Hb, int Wb, float fill, const float* lut,
               void* dst, int dst_dtype, int layout, cudaStream_t stream) {
  switch (dst_dtype) {
    case YB_F32:
      if (layout == YB_LAYOUT_S2D16) break;
      return launch_typed<SrcT, float, kHwc>(bg, img0, count, Hb, Wb, fill, lut, dst, layout, stream);
    case YB_F16:
      return launch_typed<SrcT, __half, kHwc>(bg, img0, count, Hb, Wb, fill, lut, dst, layout, stream);
    case YB_BF16:
      return launch_typed<SrcT, __nv_bfloat16, kHwc>(bg, img0, count, Hb, Wb, fill, lut, dst, layout, stream);
    default:
      break;
  }
  set_error("letterbox: unsupported destination dtype %d for layout %d", dst_dtype, layout);
  return YB_ERR_INVALID;
}

}  // namespace
}  // namespace yb

extern "C" int yb_letterbox_geometry(int n, const int32_t* src_hw, float min_size, float max_size,
                                     int size_divisible, const int32_t* fixed_shape,
                                     yb_letterbox_geom* geom, int32_t* batch_hw) {
  YB_REQUIRE(n > 0 && src_hw && geom && batch_hw, "letterbox_geometry: null/empty arguments");
  YB_REQUIRE(size_divisible > 0, "letterbox_geometry: size_divisible must be positive");
  int max_h = 0, max_w = 0;
  for (int i = 0; i < n; ++i) {
    const int h = src_hw[2 * i], w = src_hw[2 * i + 1];
    YB_REQUIRE(h > 0 && w > 0, "letterbox_geometry: image %d has non-positive size", i);
    const int lo = h < w ? h : w, hi = h < w ? w : h;
    // transform.py:66-73: python-float / fp32 0-dim tensor == tensor.reciprocal() * scalar, in fp32.
    volatile float ra = 1.0f / static_cast<float>(lo);
    volatile float rb = 1.0f / static_cast<float>(hi);
    volatile float a = ra * min_size;
    volatile float b = rb * max_size;
    const double s = static_cast<double>(a < b ? a : b);
    const int nh = static_cast<int>(static_cast<double>(h) * s);
    const int nw = static_cast<int>(static_cast<double>(w) * s);
    YB_REQUIRE(nh > 0 && nw > 0, "letterbox_geometry: image %d resizes to an empty image", i);
    geom[i].src_h = h;
    geom[i].src_w = w;
    geom[i].new_h = nh;
    geom[i].new_w = nw;
    geom[i].ratio_h = static_cast<float>(h) / static_cast<float>(nh);
    geom[i].ratio_w = static_cast<float>(w) / static_cast<float>(nw);
    if (nh > max_h) max_h = nh;
    if (nw > max_w) max_w = nw;
  }
  int Hb, Wb;
  if (fixed_shape) {
    Hb = fixed_shape[0];
    Wb = fixed_shape[1];
    YB_REQUIRE(Hb >= max_h && Wb >= max_w, "letterbox_geometry: fixed_shape (%d,%d) smaller than resized (%d,%d)",
               Hb, Wb, max_h, max_w);
  } else {
    const double d = static_cast<double>(size_divisible);
    Hb = static_cast<int>(std::ceil(static_cast<double>(max_h) / d) * d);
    Wb = static_cast<int>(std::ceil(static_cast<double>(max_w) / d) * d);
  }
  for (int i = 0; i < n; ++i) {
    // transform.py:322-326: int(round(d/2 - 0.1)), Python round == round-half-even on doubles.
    geom[i].top = static_cast<int>(std::nearbyint((Hb - geom[i].new_h) / 2.0 - 0.1));
    geom[i].left = static_cast<int>(std::nearbyint((Wb - geom[i].new_w) / 2.0 - 0.1));
  }
  batch_hw[0] = Hb;
  batch_hw[1] = Wb;
  return YB_OK;
}

extern "C" int yb_scale_coords_params(int Hb, int Wb, int src_h, int src_w, float* out3) {
  YB_REQUIRE(out3 && src_h > 0 && src_w > 0, "scale_coords_params: bad arguments");
  // transform.py:358-359 on fp32 tensors: gain = min(Hb/h, Wb/w); pad = (Wb - w*gain)/2, (Hb - h*gain)/2
  volatile float gh = static_cast<float>(Hb) / static_cast<float>(src_h);
  volatile float gw = static_cast<float>(Wb) / static_cast<float>(src_w);
  volatile float gain = gh < gw ? gh : gw;
  volatile float wx = static_cast<float>(src_w) * gain;
  volatile float hy = static_cast<float>(src_h) * gain;
  volatile float px = (static_cast<float>(Wb) - wx) / 2.0f;
  volatile float py = (static_cast<float>(Hb) - hy) / 2.0f;
  out3[0] = gain;
  out3[1] = px;
  out3[2] = py;
  return YB_OK;
}

extern "C" int yb_letterbox(int n, const void* const* src_dev, int src_dtype, const yb_letterbox_geom* geom,
                            int Hb, int Wb, float fill, const float* u8_lut_dev, void* dst_dev,
                            int dst_dtype, int dst_layout, void* stream_) {
  return yb_letterbox_strided(n, src_dev, src_dtype, YB_SRC_CHW, geom, Hb, Wb, fill, u8_lut_dev, dst_dev, dst_dtype,
                              dst_layout, stream_);
}

extern "C" int yb_letterbox_strided(int n, const void* const* src_dev, int src_dtype, int src_layout,
                                    const yb_letterbox_geom* geom, int Hb, int Wb, float fill,
                                    const float* u8_lut_dev, void* dst_dev, int dst_dtype, int dst_layout,
                                    void* stream_) {
  using namespace yb;
  YB_REQUIRE(src_layout == YB_SRC_CHW || src_layout == YB_SRC_HWC, "letterbox: bad source layout %d", src_layout);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YB_REQUIRE(n > 0 && src_dev && geom && dst_dev, "letterbox: null/empty arguments");
  YB_REQUIRE(dst_layout == YB_LAYOUT_NCHW || dst_layout == YB_LAYOUT_S2D16, "letterbox: bad layout");
  YB_REQUIRE(dst_layout != YB_LAYOUT_S2D16 || (Hb % 2 == 0 && Wb % 2 == 0),
             "letterbox: S2D16 layout needs even canvas size");
  YB_REQUIRE(src_dtype != YB_U8 || u8_lut_dev != nullptr, "letterbox: uint8 sources need the 256-entry LUT");
  for (int i0 = 0; i0 < n; i0 += kMaxImagesPerLaunch) {
    const int count = (n - i0) < kMaxImagesPerLaunch ? (n - i0) : kMaxImagesPerLaunch;
    BatchGeom bg;
    for (int j = 0; j < count; ++j) {
      const yb_letterbox_geom& g = geom[i0 + j];
      YB_REQUIRE(g.top >= 0 && g.left >= 0 && g.top + g.new_h <= Hb && g.left + g.new_w <= Wb,
                 "letterbox: image %d does not fit the canvas", i0 + j);
      YB_REQUIRE(static_cast<long long>(g.src_h) * g.src_w * 3 < (1ll << 31), "letterbox: image %d too large", i0 + j);
      bg.img[j] = ImgGeom{src_dev[i0 + j], g.src_h, g.src_w, g.new_h, g.new_w, g.top, g.left, g.ratio_h, g.ratio_w};
    }
    int rc;
    const bool hwc = src_layout == YB_SRC_HWC;
    switch (src_dtype) {
      case YB_U8:
        rc = hwc ? launch_src<uint8_t, true>(bg, i0, count, Hb, Wb, fill, u8_lut_dev, dst_dev, dst_dtype, dst_layout, stream)
                 : launch_src<uint8_t, false>(bg, i0, count, Hb, Wb, fill, u8_lut_dev, dst_dev, dst_dtype, dst_layout, stream);
        break;
      case YB_F32:
        rc = hwc ? launch_src<float, true>(bg, i0, count, Hb, Wb, fill, u8_lut_dev, dst_dev, dst_dtype, dst_layout, stream)
                 : launch_src<float, false>(bg, i0, count, Hb, Wb, fill, u8_lut_dev, dst_dev, dst_dtype, dst_layout, stream);
        break;
      case YB_F16:
        rc = hwc ? launch_src<__half, true>(bg, i0, count, Hb, Wb, fill, u8_lut_dev, dst_dev, dst_dtype, dst_layout, stream)
                 : launch_src<__half, false>(bg, i0, count, Hb, Wb, fill, u8_lut_dev, dst_dev, dst_dtype, dst_layout, stream);
        break;
      case YB_BF16:
        rc = hwc ? launch_src<__nv_bfloat16, true>(bg, i0, count, Hb, Wb, fill, u8_lut_dev, dst_dev, dst_dtype, dst_layout, stream)
                 : launch_src<__nv_bfloat16, false>(bg, i0, count, Hb, Wb, fill, u8_lut_dev, dst_dev, dst_dtype, dst_layout, stream);
        break;
      default:
        set_error("letterbox: unsupported source dtype %d", src_dtype);
        return YB_ERR_INVALID;
    }
    if (rc != YB_OK) return rc;
  }
  return YB_OK;
}
