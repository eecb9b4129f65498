// decode_mega.cuh -- argument block of the persistent decode kernel (decode_mega.cu), filled by llm.cu.
#pragma once
#include "engine.cuh"

namespace fo1 {

constexpr int kMgMaxSplits = 8;

struct MegaLayer { const bf16 *qkv_w, *qkv_b, *o_w, *gu_w, *down_w; };

struct MegaArgs {
  int B, layers, H, QD, KD, hd, q_heads, kv_heads, I, V;
  float eps, theta;
  int sec_t, sec_h;
  const MegaLayer* layer;     // [layers] (device)
  const bf16* head_w;         // [V][H], rows scaled by the final norm gain
  const bf16* embed;          // [V][H]
  bf16 *kc, *vc;              // K / V cache [layers][B][cap][KD] (K rotated)
  long long kv_layer_stride;
  int cap;
  // activations (global, L2 resident)
  bf16 *x, *x_mid, *qkv, *att, *h;
  float* ssq;                 // [2][grid][32] partial row sums of squares: [0] residual stream entering a layer, [1] after attention
  float* cs;                  // [2][B][hd/2] cos | sin of the current positions
  float* att_part;            // [B][q_heads][kMgMaxSplits][hd + 4]
  int* att_count;             // [B][kv_heads] arrival counters (self-resetting)
  float* amax_val; int* amax_idx;   // [grid][32]
  // loop state (DecodeState of llm.cu) and outputs
  int *cache_len, *pos3, *cur_tok, *finished, *n_active, *step;
  const int* stop_ids; int n_stop, pad_id, max_new;
  int* out_tokens; int* out_lens;
  int n_steps;                // decode iterations to run at most
  unsigned* bar;              // grid barrier counter (zeroed by the host before the launch)
  int n_splits;
  unsigned long long* prof;   // optional [grid][prof_slots][2] barrier enter / leave stamps of the first iteration (nullptr: off)
  int prof_slots;
};


int decode_mega_grid();                                   // CTAs of the cooperative launch (0: not available on this device)
int decode_mega_run(const MegaArgs& a, cudaStream_t s);   // the whole greedy loop, one launch

}  // namespace fo1
