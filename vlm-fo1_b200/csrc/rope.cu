// rope.cu -- rotary position embeddings: ViT 2-D RoPE and the LLM's multimodal 3-axis M-RoPE.
#include "kernels.cuh"

namespace fo1 {

// cos_sin layout: [T][half] cos followed by [T][half] sin, half = head_dim / 2.
// rot_pos_emb (modeling_qwen2_5_vl.py:436-463): rotary dim = head_dim/2; inv_freq[j] = theta^(-2j/rot), j < rot/2;
// entries [0, rot/2) use the h position, [rot/2, rot) the w position; emb = cat(rotary, rotary) so the
// pair (x[j], x[j + half]) is rotated by angle[j], j < half (qwen2_5_vl_encoder.py:110-116).
__global__ void vit_rope_table_kernel(const int* __restrict__ pos_hw, float* __restrict__ cs, int T, int half, float theta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * half) return;
  const int t = i / half, j = i - t * half;
  const int q = half / 2;
  const int jj = j < q ? j : j - q;
  const float inv = 1.0f / powf(theta, (float)(2 * jj) / (float)half);
  const float ang = (float)pos_hw[2 * t + (j < q ? 0 : 1)] * inv;
  cs[i] = cosf(ang);
  cs[(long long)T * half + i] = sinf(ang);
}

__global__ void vit_rope_apply_kernel(bf16* __restrict__ qkv, const float* __restrict__ cs, int T, int heads, int hd) {
  const int half = hd / 2;
  const long long n = (long long)T * 2 * heads * half;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % half);
    long long r = i / half;
    const int h = (int)(r % heads); r /= heads;
    const int which = (int)(r % 2);  // 0: q, 1: k
    const int t = (int)(r / 2);
    bf16* p = qkv + (long long)t * 3 * heads * hd + (long long)which * heads * hd + (long long)h * hd;
    const float c = cs[(long long)t * half + j], s = cs[(long long)T * half + (long long)t * half + j];
    const float x1 = __bfloat162float(p[j]), x2 = __bfloat162float(p[j + half]);
    p[j] = __float2bfloat16_rn(x1 * c - x2 * s);
    p[j + half] = __float2bfloat16_rn(x2 * c + x1 * s);
  }
}

int vit_rope_table(const int* pos_hw, float* cos_sin, int T, int head_dim, float theta, cudaStream_t s) {
  FO1_CHECK_ARG(head_dim % 4 == 0, "vit_rope_table: head_dim %d must be a multiple of 4", head_dim);
  if (T == 0) return FO1_OK;
  const int half = head_dim / 2;
  vit_rope_table_kernel<<<ceil_div(T * half, 256), 256, 0, s>>>(pos_hw, cos_sin, T, half, theta);
  FO1_LAUNCH_CHECK();
  return FO1_OK;
}

int vit_rope_apply(bf16* qkv, const float* cos_sin, int T, int heads, int head_dim, cudaStream_t s) {
  if (T == 0) return FO1_OK;
  const long long n = (long long)T * 2 * heads * (head_dim / 2);
  const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  vit_rope_apply_kernel<<<blocks, 256, 0, s>>>(qkv, cos_sin, T, heads, head_dim);
  FO1_LAUNCH_CHECK();
  return FO1_OK;
}

// M-RoPE: inv_freq[j] = theta^(-2j/hd), j < hd/2; frequency j takes its position from axis
// t / h / w according to the sections [sec_t, sec_h, sec_w] (sum = hd/2); cos/sin duplicated over the two
// halves (apply_multimodal_rotary_pos_emb, modeling_qwen2_5_vl.py:675-685).
__global__ void mrope_apply_kernel(bf16* __restrict__ q, bf16* __restrict__ k, long long ld, const int* __restrict__ pos3, int T,
                                   int qh, int kvh, int hd, int sec_t, int sec_h, float theta) {
  const int half = hd / 2;
  const int heads = qh + kvh;
  const long long n = (long long)T * heads * half;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % half);
    long long r = i / half;
    const int h = (int)(r % heads);
    const int t = (int)(r / heads);
    const int axis = j < sec_t ? 0 : (j < sec_t + sec_h ? 1 : 2);
    const float inv = 1.0f / powf(theta, (float)(2 * j) / (float)hd);
    const float ang = (float)pos3[(long long)axis * T + t] * inv;
    const float c = cosf(ang), s = sinf(ang);
    bf16* p = (h < qh) ? q + (long long)t * ld + (long long)h * hd : k + (long long)t * ld + (long long)(h - qh) * hd;
    const float x1 = __bfloat162float(p[j]), x2 = __bfloat162float(p[j + half]);
    p[j] = __float2bfloat16_rn(x1 * c - x2 * s);
    p[j + half] = __float2bfloat16_rn(x2 * c + x1 * s);
  }
}

int mrope_apply(bf16* q, bf16* k, long long ld, const int* pos3, int T, int q_heads, int kv_heads, int head_dim, int sec_t,
                int sec_h, int sec_w, float theta, cudaStream_t s) {
  FO1_CHECK_ARG(sec_t + sec_h + sec_w == head_dim / 2, "mrope_apply: sections %d+%d+%d != head_dim/2", sec_t, sec_h, sec_w);
  if (T == 0) return FO1_OK;
  const long long n = (long long)T * (q_heads + kv_heads) * (head_dim / 2);
  const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  mrope_apply_kernel<<<blocks, 256, 0, s>>>(q, k, ld, pos3, T, q_heads, kv_heads, head_dim, sec_t, sec_h, theta);
  FO1_LAUNCH_CHECK();
  return FO1_OK;
}

}  // namespace fo1
