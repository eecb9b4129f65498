// preprocess.cu -- device-side image pre-processing: uint8 RGB in, the two towers' input
// tensors out, bit-identical to the host processors the reference calls.
//
//   primary tower  Qwen2VLImageProcessor.preprocess (called at vlm_fo1/mm_utils.py:615): smart-resize to multiples of 28 with
//                  PIL's bicubic filter, x/255, CLIP mean/std, patchify in 2x2-merge order with the frame repeated over
//                  the temporal patch -> fp32 [gh*gw][3*2*14*14]
//   aux tower      CLIPImageProcessor.preprocess (davit/image_processing_clip.py:222-367, :344-358; called at mm_utils.py:596):
//                  optional bicubic squash to size x size, x*(1/255), ImageNet mean/std, CHW -> fp32 [3][H][W]
//
// The resize is PIL's ImagingResample for 8-bit images restated in integers: per output column (row) a window
// [xmin, xmin + n) of the input and n coefficients of the bicubic kernel (a = -0.5, support scaled by the shrink factor),
// normalised and rounded to 22 fractional bits ON THE HOST in double precision exactly as PIL does; the pass is then
// out = clip8((sum_i in[xmin + i] * k[i] + 2^21) >> 22), horizontal first, the uint8 intermediate in between.  Integer
// arithmetic, so the result equals PIL's byte for byte (tests/test_preprocess_model.py pins the restatement against PIL
// itself on the CPU; tests/test_gpu_preprocess.py pins these kernels against the host processors).
// Memory-bound byte work: one thread per output byte (channels innermost, so a warp reads / writes consecutive bytes),
// the normalise + patchify kernel writes each 1176-float token row contiguously.
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace fo1 {

constexpr int kResizePrecisionBits = 32 - 8 - 2;   // PIL: PRECISION_BITS

// horizontal pass: in [H][W][3] -> out [H][OW][3]
__global__ void __launch_bounds__(256) resize_h_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int H, int W, int OW,
                                                          const int* __restrict__ bounds, const int* __restrict__ coef, int ksize) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)H * OW * 3;
  if (i >= total) return;
  const int c = (int)(i % 3);
  const int xx = (int)((i / 3) % OW);
  const int y = (int)(i / (3LL * OW));
  const int xmin = __ldg(bounds + 2 * xx), n = __ldg(bounds + 2 * xx + 1);
  const uint8_t* row = in + ((long long)y * W + xmin) * 3 + c;
  const int* k = coef + (long long)xx * ksize;
  int acc = 1 << (kResizePrecisionBits - 1);
  for (int j = 0; j < n; ++j) acc += (int)row[3 * j] * __ldg(k + j);
  out[i] = (uint8_t)min(max(acc >> kResizePrecisionBits, 0), 255);
}

// vertical pass: in [H][W][3] -> out [OH][W][3]
__global__ void __launch_bounds__(256) resize_v_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int H, int W, int OH,
                                                          const int* __restrict__ bounds, const int* __restrict__ coef, int ksize) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long row_bytes = (long long)W * 3;
  if (i >= row_bytes * OH) return;
  const int yy = (int)(i / row_bytes);
  const long long col = i - (long long)yy * row_bytes;
  const int ymin = __ldg(bounds + 2 * yy), n = __ldg(bounds + 2 * yy + 1);
  const uint8_t* p = in + (long long)ymin * row_bytes + col;
  const int* k = coef + (long long)yy * ksize;
  int acc = 1 << (kResizePrecisionBits - 1);
  for (int j = 0; j < n; ++j) acc += (int)p[(long long)j * row_bytes] * __ldg(k + j);
  out[i] = (uint8_t)min(max(acc >> kResizePrecisionBits, 0), 255);
}

// (x / 255 - mean) / std in numpy's float32 operation order, then the 2x2-merge patch order:
// token = ((by * (gw/2) + bx) * 2 + sy) * 2 + sx, feature = (c * T + t) * p*p + py * p + px  (the frame repeated over t)
__global__ void __launch_bounds__(256) primary_patchify_kernel(const uint8_t* __restrict__ img, float* __restrict__ out, int gh, int gw, int patch,
                                                               int merge, int temporal, float m0, float m1, float m2, float s0, float s1, float s2) {
  const int token = blockIdx.x;
  const int unit = merge * merge;
  const int cell = token / unit, sub = token - cell * unit;
  const int lw = gw / merge;
  const int by = cell / lw, bx = cell - by * lw;
  const int ty = by * merge + sub / merge, tx = bx * merge + sub % merge;     // patch coordinates in the image grid
  const int W = gw * patch;
  const int pp = patch * patch;
  const int feat = 3 * temporal * pp;
  float* o = out + (long long)token * feat;
  for (int f = threadIdx.x; f < feat; f += blockDim.x) {
    const int c = f / (temporal * pp);
    const int r = f % pp;
    const int py = r / patch, px = r - py * patch;
    const uint8_t v = img[((long long)(ty * patch + py) * W + (tx * patch + px)) * 3 + c];
    const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
    o[f] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.0f), mean), sd);
  }
}

// (x * float32(1/255) - mean) / std, HWC -> CHW (davit/image_processing_clip.py:344-358)
__global__ void __launch_bounds__(256) aux_normalize_kernel(const uint8_t* __restrict__ img, float* __restrict__ out, int H, int W, float m0, float m1,
                                                            float m2, float s0, float s1, float s2) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long plane = (long long)H * W;
  if (i >= 3 * plane) return;
  const int c = (int)(i / plane);
  const long long px = i - (long long)c * plane;
  const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
  const float r255 = (float)(1.0 / 255.0);
  out[i] = __fdiv_rn(__fsub_rn(__fmul_rn((float)img[px * 3 + c], r255), mean), sd);
}

}  // namespace fo1

using namespace fo1;

extern "C" int fo1_resize_bicubic_u8(const void* img, int32_t H, int32_t W, int32_t out_h, int32_t out_w, const int32_t* h_bounds,
                                     const int32_t* h_coef, int32_t h_ksize, const int32_t* v_bounds, const int32_t* v_coef, int32_t v_ksize,
                                     void* tmp, void* out, void* stream) {
  FO1_CHECK_ARG(img && out && H > 0 && W > 0 && out_h > 0 && out_w > 0, "fo1_resize_bicubic_u8: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const uint8_t* src = static_cast<const uint8_t*>(img);
  const bool do_h = out_w != W, do_v = out_h != H;
  FO1_CHECK_ARG(!do_h || (h_bounds && h_coef && h_ksize > 0), "fo1_resize_bicubic_u8: horizontal coefficients missing");
  FO1_CHECK_ARG(!do_v || (v_bounds && v_coef && v_ksize > 0), "fo1_resize_bicubic_u8: vertical coefficients missing");
  FO1_CHECK_ARG(!(do_h && do_v) || tmp, "fo1_resize_bicubic_u8: both passes need the [H][out_w][3] intermediate");
  if (!do_h && !do_v) {
    FO1_CUDA(cudaMemcpyAsync(out, img, (size_t)H * W * 3, cudaMemcpyDeviceToDevice, s));
    return FO1_OK;
  }
  if (do_h) {
    uint8_t* dst = static_cast<uint8_t*>(do_v ? tmp : out);
    const long long n = (long long)H * out_w * 3;
    resize_h_u8_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, s>>>(src, dst, H, W, out_w, h_bounds, h_coef, h_ksize);
    FO1_LAUNCH_CHECK();
    src = dst;
  }
  if (do_v) {
    const long long n = (long long)out_h * out_w * 3;
    resize_v_u8_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, s>>>(src, static_cast<uint8_t*>(out), H, out_w, out_h, v_bounds, v_coef, v_ksize);
    FO1_LAUNCH_CHECK();
  }
  return FO1_OK;
}

extern "C" int fo1_preprocess_primary_u8(const void* img, int32_t H, int32_t W, int32_t patch, int32_t merge, int32_t temporal, const float* mean,
                                         const float* std, float* pixel_values, void* stream) {
  FO1_CHECK_ARG(img && pixel_values && mean && std, "fo1_preprocess_primary_u8: null argument");
  FO1_CHECK_ARG(patch > 0 && merge > 0 && temporal > 0 && H % (patch * merge) == 0 && W % (patch * merge) == 0,
                "fo1_preprocess_primary_u8: %dx%d is not a multiple of patch*merge = %d (resize first)", H, W, patch * merge);
  const int gh = H / patch, gw = W / patch;
  primary_patchify_kernel<<<gh * gw, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint8_t*>(img), pixel_values, gh, gw, patch, merge,
                                                                                   temporal, mean[0], mean[1], mean[2], std[0], std[1], std[2]);
  FO1_LAUNCH_CHECK();
  return FO1_OK;
}

extern "C" int fo1_preprocess_aux_u8(const void* img, int32_t H, int32_t W, const float* mean, const float* std, float* out, void* stream) {
  FO1_CHECK_ARG(img && out && mean && std && H > 0 && W > 0, "fo1_preprocess_aux_u8: bad argument");
  const long long n = 3LL * H * W;
  aux_normalize_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint8_t*>(img), out, H, W, mean[0],
                                                                                                    mean[1], mean[2], std[0], std[1], std[2]);
  FO1_LAUNCH_CHECK();
  return FO1_OK;
}
