// skinny_gemm.cu -- D[M<=32][N] = epilogue(X[M][K] . W[N][K]^T): the decode-step contraction (M = batch).
//
// With M <= 32 the op is pure weight streaming: 2 FLOP per weight byte x M, i.e. HBM-bound by two orders of
// magnitude.  A 128-row UMMA tile would spend its time on launch / TMEM / barrier fixed costs and cannot keep enough
// bytes in flight per SM (measured: 0.45-1.1 TB/s with the tcgen05 kernel), so this path is built for memory-level
// parallelism instead: each warp owns 16 weight rows x a K slice and streams them with unrolled 16-byte
// L1-bypassing loads (8 KB in flight per warp, ~30 warps per SM); the batch activations sit in shared memory; the
// multiply-accumulate rides on warp-level HMMA (mma.sync m16n8k16, weights as the 16-row operand, the batch as the
// 8-column operand) purely because SIMT FMAs would be the bottleneck at M = 32.  K splits park their fp32 partials in
// a scratch with plain stores; a small second kernel adds them in fixed order (bitwise reproducible, no float
// atomics) and applies bias / activation / gating / residual.
#include <map>
#include <mutex>

#include "kernels.cuh"

namespace fo1 {

constexpr int kSgWarps = 4;                 // warps per block, 16 weight rows each
constexpr int kSgRows = kSgWarps * 16;
constexpr int kSgMaxM = 32;

__device__ __forceinline__ void hmma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// partial layout: [split][N][32] fp32 (weight row major, batch contiguous); every split writes its slot with plain
// stores and the finish kernel adds the splits in order, so results are bitwise reproducible (no float atomics)
template <int KS>
__global__ void __launch_bounds__(kSgWarps * 32) skinny_gemm_kernel(const bf16* __restrict__ X, long long ldx, const bf16* __restrict__ W,
                                                                    long long ldw, float* __restrict__ part, int M, int N, int K,
                                                                    int slices_per_split) {
  constexpr int PITCH = KS + 8;             // +16 B: the 8 batch rows of a fragment load land in distinct 16-byte bank groups
  extern __shared__ __align__(16) bf16 xs[];  // [32][PITCH]
  const int n0 = blockIdx.x * kSgRows;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int row_a = n0 + warp * 16 + g, row_b = row_a + 8;
  float c[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
  for (int sl = 0; sl < slices_per_split; ++sl) {
  const int k0 = (blockIdx.y * slices_per_split + sl) * KS;
  if (k0 >= K) break;
  const int klen = min(KS, K - k0);         // multiple of 32 (checked on the host)
  __syncthreads();                          // the previous slice's activations are no longer read
  // ---- stage the batch activations of this K slice (rows >= M are zero) ----
  for (int i = threadIdx.x; i < kSgMaxM * (KS / 8); i += blockDim.x) {
    const int r = i / (KS / 8), cc = (i % (KS / 8)) * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < M && cc < klen) v = *reinterpret_cast<const uint4*>(X + (long long)r * ldx + k0 + cc);
    *reinterpret_cast<uint4*>(xs + r * PITCH + cc) = v;
  }
  __syncthreads();
  const bf16* wa = W + (long long)min(row_a, N - 1) * ldw + k0 + 8 * t;   // clamped rows are computed but never stored
  const bf16* wb = W + (long long)min(row_b, N - 1) * ldw + k0 + 8 * t;
  const int nchunk = klen / 32;
  int ch = 0;
  for (; ch + 4 <= nchunk; ch += 4) {
    uint4 va[4], vb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      va[u] = ldg_nc_v4(wa + (ch + u) * 32);
      vb[u] = ldg_nc_v4(wb + (ch + u) * 32);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 xv = *reinterpret_cast<const uint4*>(xs + (8 * j + g) * PITCH + (ch + u) * 32 + 8 * t);
        hmma_16816(c[j], va[u].x, vb[u].x, va[u].y, vb[u].y, xv.x, xv.y);
        hmma_16816(c[j], va[u].z, vb[u].z, va[u].w, vb[u].w, xv.z, xv.w);
      }
    }
  }
  for (; ch < nchunk; ++ch) {
    const uint4 va = ldg_nc_v4(wa + ch * 32), vb = ldg_nc_v4(wb + ch * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 xv = *reinterpret_cast<const uint4*>(xs + (8 * j + g) * PITCH + ch * 32 + 8 * t);
      hmma_16816(c[j], va.x, vb.x, va.y, vb.y, xv.x, xv.y);
      hmma_16816(c[j], va.z, vb.z, va.w, vb.w, xv.z, xv.w);
    }
  }
  }  // K slices of this split
  // c[j][0,1] = (weight row g, batch 8j + 2t, +1); c[j][2,3] = (weight row g + 8, same batch columns)
  float* out = part + (long long)blockIdx.y * N * 32;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int b = 8 * j + 2 * t;
    if (row_a < N) *reinterpret_cast<float2*>(out + (long long)row_a * 32 + b) = make_float2(c[j][0], c[j][1]);
    if (row_b < N) *reinterpret_cast<float2*>(out + (long long)row_b * 32 + b) = make_float2(c[j][2], c[j][3]);
  }
}

struct SgEpi {
  void* D; long long ldd; int d_dtype;
  const void* bias; int bias_dtype; int act;
  const bf16* residual; long long ldr;
  int gated;
};
__device__ __forceinline__ float sg_bias(const SgEpi& e, int n) {
  if (e.bias == nullptr) return 0.f;
  return e.bias_dtype == FO1_F32 ? static_cast<const float*>(e.bias)[n] : __bfloat162float(static_cast<const bf16*>(e.bias)[n]);
}
__device__ __forceinline__ float sg_act(float x, int act) { return act == FO1_EPI_GELU ? gelu_erf(x) : (act == FO1_EPI_SILU ? silu(x) : x); }

// one thread per output column: adds the column's partial sums over the K splits (fixed order), finishes M rows
__global__ void __launch_bounds__(256) skinny_finish_kernel(const float* __restrict__ part, int ksplit, const SgEpi e, int M, int N) {
  const int n_out = e.gated ? N / 2 : N;
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  int n_a = o, n_b = 0;
  if (e.gated) { n_a = (o / 32) * 64 + (o % 32); n_b = n_a + 32; }   // [32 gate | 32 up] row interleave
  const float ba = sg_bias(e, n_a), bb = e.gated ? sg_bias(e, n_b) : 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    if (q * 4 >= M) break;
    float4 va = make_float4(0, 0, 0, 0), vb = make_float4(0, 0, 0, 0);
    for (int sp = 0; sp < ksplit; ++sp) {
      const float4 ta = *reinterpret_cast<const float4*>(part + ((long long)sp * N + n_a) * 32 + q * 4);
      va.x += ta.x; va.y += ta.y; va.z += ta.z; va.w += ta.w;
      if (e.gated) {
        const float4 tb = *reinterpret_cast<const float4*>(part + ((long long)sp * N + n_b) * 32 + q * 4);
        vb.x += tb.x; vb.y += tb.y; vb.z += tb.z; vb.w += tb.w;
      }
    }
    const float xa[4] = {va.x, va.y, va.z, va.w}, xb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = q * 4 + i;
      if (m >= M) break;
      float v = sg_act(xa[i] + ba, e.act);
      if (e.gated) v *= (xb[i] + bb);
      if (e.residual) v += __bfloat162float(e.residual[(long long)m * e.ldr + o]);
      if (e.d_dtype == FO1_BF16) static_cast<bf16*>(e.D)[(long long)m * e.ldd + o] = __float2bfloat16_rn(v);
      else static_cast<float*>(e.D)[(long long)m * e.ldd + o] = v;
    }
  }
}

struct SgScratch { float* acc = nullptr; size_t floats = 0; };
static int sg_scratch(cudaStream_t stream, size_t need, float** out) {
  static std::map<cudaStream_t, SgScratch> pool;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  SgScratch& sc = pool[stream];
  if (need > sc.floats) {
    FO1_CUDA(cudaStreamSynchronize(stream));
    if (sc.acc) FO1_CUDA(cudaFree(sc.acc));
    sc.acc = nullptr; sc.floats = 0;
    const size_t want = std::max(need, (size_t)160000 * 32);   // covers the 151,936-row LM head
    FO1_CUDA(cudaMalloc(reinterpret_cast<void**>(&sc.acc), want * sizeof(float)));
    sc.floats = want;
  }
  *out = sc.acc;
  return FO1_OK;
}

bool skinny_gemm_supported(const fo1_gemm_desc* d) {
  return d->M >= 1 && d->M <= kSgMaxM && d->K % 32 == 0 && d->lda % 8 == 0 && d->ldw % 8 == 0 && (!d->gated || d->N % 64 == 0) &&
         getenv("FO1_NO_SKINNY") == nullptr;
}

int skinny_gemm(const fo1_gemm_desc* d, cudaStream_t stream) {
  // K slicing: `ksplit` blocks share a row block (>= ~4 blocks per SM overall), each walking `per` slices of KS
  const long long row_blocks = ceil_div(d->N, kSgRows);
  const bool wide = d->K > 4096;
  const int KSv = wide ? 1024 : 512;
  const int n_slices = ceil_div(d->K, KSv);
  int ksplit = (int)std::min<long long>(n_slices, std::max<long long>(1, ceil_div(4 * device_sm_count(), (int)std::min<long long>(row_blocks, 1 << 20))));
  const int per = ceil_div(n_slices, ksplit);
  ksplit = ceil_div(n_slices, per);
  float* part = nullptr;
  FO1_TRY(sg_scratch(stream, (size_t)ksplit * d->N * 32, &part));
  char tag[96] = "gemm_skinny";
  if (g_prof_on) snprintf(tag, sizeof(tag), "gemm_skinny:%dx%dx%d%s", d->M, d->N, d->K, d->gated ? ":gated" : "");
  ProfScope prof(tag, 2.0 * d->M * (double)d->N * d->K,
                 2.0 * ((double)d->M * d->K + (double)d->N * d->K + (double)d->M * (d->gated ? d->N / 2 : d->N)), stream);
  const bf16* X = static_cast<const bf16*>(d->A);
  const bf16* W = static_cast<const bf16*>(d->W);
  dim3 grid((unsigned)row_blocks, ksplit);
  if (wide) {
    constexpr int KS = 1024;
    static bool set = false;
    const int smem = kSgMaxM * (KS + 8) * 2;
    if (!set) { FO1_CUDA(cudaFuncSetAttribute(skinny_gemm_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); set = true; }
    skinny_gemm_kernel<KS><<<grid, kSgWarps * 32, smem, stream>>>(X, d->lda, W, d->ldw, part, d->M, d->N, d->K, per);
  } else {
    constexpr int KS = 512;
    const int smem = kSgMaxM * (KS + 8) * 2;
    skinny_gemm_kernel<KS><<<grid, kSgWarps * 32, smem, stream>>>(X, d->lda, W, d->ldw, part, d->M, d->N, d->K, per);
  }
  FO1_LAUNCH_CHECK();
  SgEpi e;
  e.D = d->D; e.ldd = d->ldd; e.d_dtype = d->d_dtype;
  e.bias = d->bias; e.bias_dtype = d->bias_dtype; e.act = d->act;
  e.residual = static_cast<const bf16*>(d->residual); e.ldr = d->ldr;
  e.gated = d->gated;
  const int n_out = d->gated ? d->N / 2 : d->N;
  skinny_finish_kernel<<<ceil_div(n_out, 256), 256, 0, stream>>>(part, ksplit, e, d->M, d->N);
  FO1_LAUNCH_CHECK();
  return FO1_OK;
}

}  // namespace fo1
