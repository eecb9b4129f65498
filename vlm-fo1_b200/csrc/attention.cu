// attention.cu -- variable-length softmax attention over packed token rows.
//
// One kernel family serves the ViT's windowed / full attention (head_dim 80, cu_seqlens = 64-token
// windows or whole images; modeling_qwen2_5_vl.py:172-209), DaViT's 12x12 window attention (head_dim 32;
// modeling_davit.py:225-282) and the LLM's causal GQA prefill (head_dim 128, 16 q / 2 kv heads;
// modeling_qwen2_5_vl.py:738-802).  Flash-style: 64-query x 64-key tiles, online softmax in fp32 (exp2),
// K/V tiles double-buffered through shared memory with cp.async, S = QK^T and O += PV on the warp-level
// tensor path (mma.sync m16n8k16 bf16 -> fp32, operands via ldmatrix).
// NOTE (round 1): this is the legacy HMMA path, not tcgen05; it carries ~6 % of the path's FLOPs.  The
// tcgen05/TMEM port of the two large-sequence users (ViT full attention, LLM prefill) is the next step.
#include "kernels.cuh"

namespace fo1 {

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
  const int sz = valid ? 16 : 0;  // src-size 0 -> 16 bytes of zeros
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int kTileN = 64;   // keys per iteration; queries per CTA = 16 x WARPS (64 for windows, 128 for long sequences)

// copy `rows_valid` rows of HD bf16 (zero-fill up to 64 rows) into a padded smem tile
template <int HD, int ROWS, int THREADS>
__device__ __forceinline__ void load_tile_async(bf16* dst, const bf16* src, long long ld, int rows_valid) {
  constexpr int PITCH = HD + 8;
  constexpr int CHUNKS = HD / 8;  // 16-byte chunks per row
  for (int i = threadIdx.x; i < ROWS * CHUNKS; i += THREADS) {
    const int r = i / CHUNKS, c = i - r * CHUNKS;
    const bool ok = r < rows_valid;
    cp_async16(dst + r * PITCH + c * 8, src + (long long)(ok ? r : 0) * ld + c * 8, ok);
  }
}

// two resident CTAs per SM when the accumulators leave room (head_dim <= 80): with one 8-warp CTA the tensor pipe sat at
// 38 % with the warps waiting on fixed-latency dependencies (ncu, profiles/)
template <int HD, bool CAUSAL, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, (HD <= 80 ? 2 : 1)) attn_varlen_kernel(const AttnArgs a) {
  constexpr int PITCH = HD + 8;
  constexpr int TILE = 64 * PITCH;
  constexpr int kTileM = 16 * WARPS, kAttnThreads = WARPS * 32;
  constexpr int QTILE = kTileM * PITCH;
  constexpr int KC = HD / 16;  // k-chunks of the QK^T contraction == d-tile pairs of the PV product
  extern __shared__ __align__(16) bf16 smem_attn[];
  bf16* sQ = smem_attn;
  bf16* sK = smem_attn + QTILE;             // 2 buffers
  bf16* sV = smem_attn + QTILE + 2 * TILE;  // 2 buffers

  const int seq = blockIdx.y, head = blockIdx.z;
  const int kvh = head / (a.q_heads / a.kv_heads);
  const int s0 = a.cu_seqlens[seq];
  const int len = a.cu_seqlens[seq + 1] - s0;
  const int q0 = blockIdx.x * kTileM;
  if (q0 >= len) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;

  int n_kv = (len + kTileN - 1) / kTileN;
  if (CAUSAL) n_kv = min(n_kv, (q0 + kTileM + kTileN - 1) / kTileN);

  const bf16* qp = a.q + (long long)(s0 + q0) * a.ldq + (long long)head * HD;
  const bf16* kp = a.k + (long long)s0 * a.ldk + (long long)kvh * HD;
  const bf16* vp = a.v + (long long)s0 * a.ldv + (long long)kvh * HD;

  load_tile_async<HD, kTileM, kAttnThreads>(sQ, qp, a.ldq, min(kTileM, len - q0));
  load_tile_async<HD, 64, kAttnThreads>(sK, kp, a.ldk, min(64, len));
  load_tile_async<HD, 64, kAttnThreads>(sV, vp, a.ldv, min(64, len));
  cp_async_commit();

  uint32_t qf[KC][4];
  float o[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float row_m[2] = {-INFINITY, -INFINITY}, row_l[2] = {0.f, 0.f};
  const float sl2 = a.scale * 1.4426950408889634f;

  for (int j = 0; j < n_kv; ++j) {
    const int buf = j & 1;
    if (j + 1 < n_kv) {
      const int k0 = (j + 1) * kTileN;
      load_tile_async<HD, 64, kAttnThreads>(sK + (buf ^ 1) * TILE, kp + (long long)k0 * a.ldk, a.ldk, min(64, len - k0));
      load_tile_async<HD, 64, kAttnThreads>(sV + (buf ^ 1) * TILE, vp + (long long)k0 * a.ldv, a.ldv, min(64, len - k0));
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (j == 0) {
#pragma unroll
      for (int kc = 0; kc < KC; ++kc) {
        const int jm = lane >> 3, r = lane & 7;
        ldsm_x4(qf[kc], sQ + (warp * 16 + (jm & 1) * 8 + r) * PITCH + kc * 16 + (jm >> 1) * 8);
      }
    }
    const bf16* cK = sK + buf * TILE;
    const bf16* cV = sV + buf * TILE;

    // ---- S = Q K^T (16 x 64 per warp) ----
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int np = 0; np < 4; ++np) {
#pragma unroll
      for (int kc = 0; kc < KC; ++kc) {
        uint32_t b[4];
        const int jm = lane >> 3, r = lane & 7;
        ldsm_x4(b, cK + (np * 16 + (jm >> 1) * 8 + r) * PITCH + kc * 16 + (jm & 1) * 8);
        mma_bf16(s[2 * np], qf[kc], b[0], b[1]);
        mma_bf16(s[2 * np + 1], qf[kc], b[2], b[3]);
      }
    }
    // ---- mask + online softmax ----
    const int kbase = j * kTileN;
    const int qrow0 = q0 + warp * 16 + g;  // rows qrow0 and qrow0 + 8
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = kbase + nt * 8 + 2 * t + (e & 1);
        const int row = qrow0 + (e >> 1) * 8;
        const bool dead = (col >= len) || (CAUSAL && col > row);
        const float v = dead ? -INFINITY : s[nt][e] * sl2;
        s[nt][e] = v;
        mx[e >> 1] = fmaxf(mx[e >> 1], v);
      }
    }
    float alpha[2], mnew[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float m = mx[h];
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
      mnew[h] = fmaxf(row_m[h], m);
      const float base = (mnew[h] == -INFINITY) ? 0.f : mnew[h];
      alpha[h] = (row_m[h] == -INFINITY) ? 0.f : exp2f(row_m[h] - base);
      row_m[h] = mnew[h];
      mnew[h] = base;
    }
    float psum[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float p = exp2f(s[nt][e] - mnew[e >> 1]);
        s[nt][e] = p;
        psum[e >> 1] += p;
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) row_l[h] = row_l[h] * alpha[h] + psum[h];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      o[i][0] *= alpha[0]; o[i][1] *= alpha[0];
      o[i][2] *= alpha[1]; o[i][3] *= alpha[1];
    }
    // ---- O += P V ----
#pragma unroll
    for (int kc2 = 0; kc2 < 4; ++kc2) {
      uint32_t pa[4];
      pa[0] = pack_bf16(s[2 * kc2][0], s[2 * kc2][1]);
      pa[1] = pack_bf16(s[2 * kc2][2], s[2 * kc2][3]);
      pa[2] = pack_bf16(s[2 * kc2 + 1][0], s[2 * kc2 + 1][1]);
      pa[3] = pack_bf16(s[2 * kc2 + 1][2], s[2 * kc2 + 1][3]);
#pragma unroll
      for (int dp = 0; dp < KC; ++dp) {
        uint32_t b[4];
        const int jm = lane >> 3, r = lane & 7;
        ldsm_x4_trans(b, cV + (kc2 * 16 + (jm & 1) * 8 + r) * PITCH + dp * 16 + (jm >> 1) * 8);
        mma_bf16(o[2 * dp], pa, b[0], b[1]);
        mma_bf16(o[2 * dp + 1], pa, b[2], b[3]);
      }
    }
    __syncthreads();  // everyone is done with this buffer before the next prefetch overwrites it
  }

  // ---- normalise, stage through this warp's slice of the Q tile, coalesced 16-byte stores ----
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float l = row_l[h];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    row_l[h] = l > 0.f ? 1.0f / l : 0.f;
  }
  bf16* stage = sQ + warp * 16 * PITCH;
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    *reinterpret_cast<uint32_t*>(stage + g * PITCH + i * 8 + 2 * t) = pack_bf16(o[i][0] * row_l[0], o[i][1] * row_l[0]);
    *reinterpret_cast<uint32_t*>(stage + (g + 8) * PITCH + i * 8 + 2 * t) = pack_bf16(o[i][2] * row_l[1], o[i][3] * row_l[1]);
  }
  __syncwarp();
  constexpr int CHUNKS = HD / 8;
  bf16* op = a.o + (long long)(s0 + q0 + warp * 16) * a.ldo + (long long)head * HD;
  for (int i = lane; i < 16 * CHUNKS; i += 32) {
    const int r = i / CHUNKS, c = i - r * CHUNKS;
    if (q0 + warp * 16 + r < len)
      *reinterpret_cast<uint4*>(op + (long long)r * a.ldo + c * 8) = *reinterpret_cast<const uint4*>(stage + r * PITCH + c * 8);
  }
}

template <int HD, int WARPS>
static int launch_attn(const AttnArgs& a, cudaStream_t s) {
  constexpr int smem = (16 * WARPS + 4 * 64) * (HD + 8) * 2;
  dim3 grid(ceil_div(a.max_seqlen, 16 * WARPS), a.n_seqs, a.q_heads);
  ProfScope prof(a.causal ? "attn_causal" : "attn", 0.0, 0.0, s);
  if (a.causal) {
    static bool set = false;
    if (!set) { FO1_CUDA(cudaFuncSetAttribute(attn_varlen_kernel<HD, true, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); set = true; }
    attn_varlen_kernel<HD, true, WARPS><<<grid, WARPS * 32, smem, s>>>(a);
  } else {
    static bool set = false;
    if (!set) { FO1_CUDA(cudaFuncSetAttribute(attn_varlen_kernel<HD, false, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); set = true; }
    attn_varlen_kernel<HD, false, WARPS><<<grid, WARPS * 32, smem, s>>>(a);
  }
  FO1_LAUNCH_CHECK();
  return FO1_OK;
}

int attention_varlen(const AttnArgs& a, cudaStream_t s) {
  FO1_CHECK_ARG(a.q && a.k && a.v && a.o && a.cu_seqlens, "attention: null pointer");
  FO1_CHECK_ARG(a.q_heads > 0 && a.kv_heads > 0 && a.q_heads % a.kv_heads == 0, "attention: heads %d/%d", a.q_heads, a.kv_heads);
  FO1_CHECK_ARG((a.ldq % 8) == 0 && (a.ldk % 8) == 0 && (a.ldv % 8) == 0 && (a.ldo % 8) == 0, "attention: pitches must be multiples of 8");
  if (a.n_seqs == 0 || a.max_seqlen == 0) return FO1_OK;
  FO1_CHECK_ARG(a.n_seqs <= 65535 && a.q_heads <= 65535, "attention: grid too large (%d seqs, %d heads)", a.n_seqs, a.q_heads);
  const bool long_seq = a.max_seqlen >= 512;   // 128-query tiles halve the K/V re-reads of long sequences
  switch (a.head_dim) {
    case 32: return launch_attn<32, 4>(a, s);
    case 80: return long_seq ? launch_attn<80, 8>(a, s) : launch_attn<80, 4>(a, s);
    case 128: return launch_attn<128, 4>(a, s);   // 64-query tiles, 2 CTAs per SM: 16 % faster than one 8-warp CTA (178 registers) on the causal prefill
    default:
      set_error("attention: head_dim %d unsupported (32, 80, 128)", a.head_dim);
      return FO1_ERR_UNSUPPORTED;
  }
}

}  // namespace fo1

extern "C" int fo1_attention_varlen(const fo1_attn_desc* d, void* stream) {
  using namespace fo1;
  FO1_CHECK_ARG(d != nullptr, "fo1_attention_varlen: null descriptor");
  AttnArgs a;
  a.q = static_cast<const bf16*>(d->q); a.k = static_cast<const bf16*>(d->k); a.v = static_cast<const bf16*>(d->v);
  a.o = static_cast<bf16*>(d->o);
  a.ldq = d->ldq; a.ldk = d->ldk; a.ldv = d->ldv; a.ldo = d->ldo;
  a.cu_seqlens = d->cu_seqlens; a.n_seqs = d->n_seqs; a.max_seqlen = d->max_seqlen;
  a.q_heads = d->q_heads; a.kv_heads = d->kv_heads; a.head_dim = d->head_dim;
  a.scale = d->scale; a.causal = d->causal;
  return attention_varlen(a, static_cast<cudaStream_t>(stream));
}
