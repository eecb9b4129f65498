// prompt.cu -- batched, device-side integer bookkeeping around the language model:
//   * fo1_splice_plan_batch : the splice of prepare_inputs_labels_for_qwen2_5_vl_multimodal (omchat_qwen2_5_vl.py:291-373,
//                             434-458) + get_rope_index (modeling_qwen2_5_vl.py:1546-1721) for a whole batch of prompts in one
//                             launch -- the same rules as the per-sample host function fo1_splice_plan (llm.cu), bit-exact;
//   * fo1_parse_predictions : "<ground>label</ground><objects><regionN>...</objects>" -> (label span, N) records straight
//                             from the decoded token ids of a batch (extract_predictions_to_indexes, mm_utils.py:346-369),
//                             for tokenizers that hold the five markers as single tokens.
// One CTA per sample; thread 0 walks the (short) prompt to build offset / segment tables in shared memory, all threads then
// fill the rows.  Pure integer work: results must equal the host functions bit for bit (tests/test_gpu_prompt.py).
#include "kernels.cuh"

namespace fo1 {

constexpr int kMaxPromptIds = 8192;     // ids of one prompt (before the image rows are expanded)
constexpr int kMaxSegs = 64;            // text / image segments of one spliced sequence

struct SpliceBatchArgs {
  const int* ids; const int* id_off;            // concatenated prompt ids [sum n], offsets [B + 1]
  const int* grids; const int* img_off;         // image grids [sum n_img][2] (gh, gw), offsets [B + 1]
  const int* n_regions;                         // [B] region feature rows available per sample
  const int* out_off;                           // [B + 1] row offsets of the spliced sequences (host arithmetic)
  const int* img_row_off; const int* reg_row_off;   // [B] first row of the sample in the batch's image- / region-feature matrices
  int* new_ids; int* kind; int* index;          // [T]
  int* pos;                                     // [3][T]
  int* rope_delta;                              // [B]
  int* status;                                  // [B] 0 ok, else a fo1_status
  long long T;
  fo1_splice_cfg cf;
};

__global__ void __launch_bounds__(256) splice_plan_batch_kernel(const SpliceBatchArgs a) {
  __shared__ int s_off[kMaxPromptIds];          // output row of every input id
  __shared__ int s_seg[kMaxSegs][5];            // (first row, length, base, lh (0 = text), lw)
  __shared__ int s_nseg, s_err, s_delta;
  const int b = blockIdx.x;
  const int i0 = a.id_off[b], n = a.id_off[b + 1] - i0;
  const int g0 = a.img_off[b], n_img = a.img_off[b + 1] - g0;
  const int r0 = a.out_off[b], L = a.out_off[b + 1] - r0;
  const int unit = a.cf.merge * a.cf.merge;
  const int* ids = a.ids + i0;
  // ---- pass 1 (thread 0): output offsets; validation (omchat_qwen2_5_vl.py:318-368) ----
  if (threadIdx.x == 0) {
    int err = FO1_OK, row = 0, img = 0, reg = 0;
    if (n > kMaxPromptIds) err = FO1_ERR_UNSUPPORTED;
    for (int i = 0; i < n && err == FO1_OK; ++i) {
      s_off[i] = row;
      const int t = ids[i];
      if (t == a.cf.image_placeholder) {
        if (img >= n_img) { err = FO1_ERR_INVALID_ARG; break; }
        row += a.grids[2 * (g0 + img)] * a.grids[2 * (g0 + img) + 1] / unit;
        ++img;
      } else {
        if (t == a.cf.region_placeholder && reg++ >= a.n_regions[b]) { err = FO1_ERR_INVALID_ARG; break; }
        ++row;
      }
    }
    if (err == FO1_OK && row != L) err = FO1_ERR_WORKSPACE;    // the host's length arithmetic disagrees
    s_err = err;
  }
  __syncthreads();
  if (s_err != FO1_OK) {
    if (threadIdx.x == 0) { a.status[b] = s_err; a.rope_delta[b] = 0; }
    return;
  }
  // ---- pass 2 (all threads): rows of kind / index / new id ----
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int t = ids[i], row = r0 + s_off[i];
    if (t == a.cf.image_placeholder) {
      int img = 0, base = 0;                    // rank of this placeholder among the sample's images, rows before it
      for (int j = 0; j < i; ++j)
        if (ids[j] == a.cf.image_placeholder) { base += a.grids[2 * (g0 + img)] * a.grids[2 * (g0 + img) + 1] / unit; ++img; }
      const int cnt = a.grids[2 * (g0 + img)] * a.grids[2 * (g0 + img) + 1] / unit;
      for (int k = 0; k < cnt; ++k) {
        a.new_ids[row + k] = a.cf.image_token_id; a.kind[row + k] = 1; a.index[row + k] = a.img_row_off[b] + base + k;
      }
    } else if (t == a.cf.region_placeholder) {
      int reg = 0;
      for (int j = 0; j < i; ++j) reg += (ids[j] == a.cf.region_placeholder);
      a.new_ids[row] = t; a.kind[row] = 2; a.index[row] = a.reg_row_off[b] + reg;
    } else {
      a.new_ids[row] = t; a.kind[row] = 0; a.index[row] = t;
    }
  }
  __syncthreads();
  // ---- pass 3 (thread 0): segment table of get_rope_index (modeling_qwen2_5_vl.py:1625-1699), attention mask all ones ----
  if (threadIdx.x == 0) {
    const int* nid = a.new_ids + r0;
    int image_nums = 0;
    for (int i = 0; i + 1 < L; ++i)
      if (nid[i] == a.cf.vision_start_token_id && nid[i + 1] == a.cf.image_token_id) ++image_nums;
    int st = 0, image_index = 0, nseg = 0, err = FO1_OK;
    long long next_base = 0;
    bool have_any = false;
    for (int it = 0; it < image_nums && err == FO1_OK; ++it) {
      int ed = -1;
      for (int i = st; i < L; ++i) if (nid[i] == a.cf.image_token_id) { ed = i; break; }
      if (ed < 0 || image_index >= n_img || nseg + 2 > kMaxSegs) { err = FO1_ERR_INVALID_ARG; break; }
      const int lh = a.grids[2 * (g0 + image_index)] / a.cf.merge, lw = a.grids[2 * (g0 + image_index) + 1] / a.cf.merge;
      ++image_index;
      const int text_len = ed - st;
      const long long st_idx = have_any ? next_base : 0;
      if (text_len > 0) { s_seg[nseg][0] = st; s_seg[nseg][1] = text_len; s_seg[nseg][2] = (int)st_idx; s_seg[nseg][3] = 0; s_seg[nseg][4] = 0; ++nseg; }
      const long long vb = st_idx + text_len;
      if ((long long)ed + (long long)lh * lw > L) { err = FO1_ERR_INVALID_ARG; break; }
      s_seg[nseg][0] = ed; s_seg[nseg][1] = lh * lw; s_seg[nseg][2] = (int)vb; s_seg[nseg][3] = lh; s_seg[nseg][4] = lw; ++nseg;
      long long mx = vb + max(lh, lw) - 1;
      if (text_len > 0) mx = max(mx, st_idx + text_len - 1);
      next_base = mx + 1;
      have_any = true;
      st = ed + lh * lw;
    }
    long long maxpos = have_any ? next_base - 1 : -1;
    if (err == FO1_OK && st < L) {
      if (nseg + 1 > kMaxSegs) err = FO1_ERR_INVALID_ARG;
      else {
        const long long st_idx = have_any ? next_base : 0;
        s_seg[nseg][0] = st; s_seg[nseg][1] = L - st; s_seg[nseg][2] = (int)st_idx; s_seg[nseg][3] = 0; s_seg[nseg][4] = 0; ++nseg;
        maxpos = st_idx + (L - st) - 1;
      }
    }
    s_nseg = nseg; s_err = err; s_delta = (int)(maxpos + 1 - L);
  }
  __syncthreads();
  if (s_err != FO1_OK) {
    if (threadIdx.x == 0) { a.status[b] = s_err; a.rope_delta[b] = 0; }
    return;
  }
  // ---- pass 4 (all threads): position ids ----
  for (int sg = 0; sg < s_nseg; ++sg) {
    const int first = s_seg[sg][0], len = s_seg[sg][1], base = s_seg[sg][2], lh = s_seg[sg][3], lw = s_seg[sg][4];
    for (int k = threadIdx.x; k < len; k += blockDim.x) {
      const long long p = r0 + first + k;
      if (lh == 0) {
        a.pos[p] = base + k; a.pos[a.T + p] = base + k; a.pos[2 * a.T + p] = base + k;
      } else {
        a.pos[p] = base;                          // t index = 0 for an image
        a.pos[a.T + p] = base + k / lw;
        a.pos[2 * a.T + p] = base + k % lw;
      }
    }
  }
  if (threadIdx.x == 0) { a.status[b] = FO1_OK; a.rope_delta[b] = s_delta; }
}

// ---- <ground>label</ground><objects><regionN>...</objects> over token ids --------------------------------------------
// The reference matches the DECODED TEXT with r"<ground>(.*?)</ground><objects>(.*?)</objects>" and then r"<region(\d+)>" inside
// the second group (mm_utils.py:346-369).  With the five markers held as single tokens that is a scan over ids: from a
// <ground> at p, the label ends at the first "</ground>" that is IMMEDIATELY followed by "<objects>", the body at the first
// "</objects>" after it; "." does not match a newline, so a token whose text contains '\n' inside either group kills the
// match from p (the scan resumes at p + 1); after a match the scan resumes behind "</objects>" (re.findall).
struct ParseArgs {
  const int* tokens; long long ld; const int* lens; int B;
  int ground_s, ground_e, objects_s, objects_e;
  const int* region_ids; int n_region_ids;        // token id of <region0>, <region1>, ...
  const unsigned* newline_bitmap; int vocab;      // bit t set: the text of token t contains '\n'
  int* records; int max_records; int* n_records;  // per sample [max_records][3] = (label first token, label end token (excl.), N)
};

__device__ __forceinline__ bool has_newline(const ParseArgs& a, int t) {
  return t >= 0 && t < a.vocab && a.newline_bitmap != nullptr && ((a.newline_bitmap[t >> 5] >> (t & 31)) & 1u);
}

__global__ void __launch_bounds__(32) parse_predictions_kernel(const ParseArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  const int* tk = a.tokens + (long long)b * a.ld;
  const int n = a.lens[b];
  int* rec = a.records + (long long)b * a.max_records * 3;
  int nrec = 0, i = 0;
  while (i < n) {
    if (tk[i] != a.ground_s) { ++i; continue; }
    // label: first j > i with tk[j] == </ground> and tk[j+1] == <objects>, no newline token in (i, j)
    int j = i + 1, le = -1;
    for (; j + 1 < n; ++j) {
      if (tk[j] == a.ground_e && tk[j + 1] == a.objects_s) { le = j; break; }
      if (has_newline(a, tk[j])) break;
    }
    if (le < 0) { ++i; continue; }
    int k = le + 2, oe = -1;
    for (; k < n; ++k) {
      if (tk[k] == a.objects_e) { oe = k; break; }
      if (has_newline(a, tk[k])) break;
    }
    if (oe < 0) { ++i; continue; }
    bool any = false;
    for (int q = le + 2; q < oe; ++q) {
      const int t = tk[q];
      int idx = -1;
      for (int r = 0; r < a.n_region_ids; ++r) if (a.region_ids[r] == t) { idx = r; break; }
      if (idx >= 0 && nrec < a.max_records) { rec[3 * nrec] = i + 1; rec[3 * nrec + 1] = le; rec[3 * nrec + 2] = idx; ++nrec; any = true; }
    }
    if (!any && nrec < a.max_records) { rec[3 * nrec] = i + 1; rec[3 * nrec + 1] = le; rec[3 * nrec + 2] = -1; ++nrec; }   // a label without regions still defines a key (empty set)
    i = oe + 1;
  }
  a.n_records[b] = nrec;
}

}  // namespace fo1

using namespace fo1;

extern "C" int fo1_splice_plan_batch(const int32_t* ids, const int32_t* id_off, const int32_t* grids, const int32_t* img_off,
                                     const int32_t* n_regions, const int32_t* out_off, const int32_t* img_row_off,
                                     const int32_t* reg_row_off, int32_t n_samples, int64_t total_rows, const fo1_splice_cfg* cfg,
                                     int32_t* new_ids, int32_t* src_kind, int32_t* src_index, int32_t* position_ids, int32_t* rope_delta,
                                     int32_t* status, void* stream) {
  FO1_CHECK_ARG(ids && id_off && grids && img_off && n_regions && out_off && img_row_off && reg_row_off && cfg, "fo1_splice_plan_batch: null input");
  FO1_CHECK_ARG(new_ids && src_kind && src_index && position_ids && rope_delta && status, "fo1_splice_plan_batch: null output");
  FO1_CHECK_ARG(cfg->merge > 0 && n_samples >= 0 && total_rows >= 0, "fo1_splice_plan_batch: bad sizes");
  if (n_samples == 0) return FO1_OK;
  SpliceBatchArgs a;
  a.ids = ids; a.id_off = id_off; a.grids = grids; a.img_off = img_off; a.n_regions = n_regions; a.out_off = out_off;
  a.img_row_off = img_row_off; a.reg_row_off = reg_row_off;
  a.new_ids = new_ids; a.kind = src_kind; a.index = src_index; a.pos = position_ids; a.rope_delta = rope_delta; a.status = status;
  a.T = total_rows; a.cf = *cfg;
  splice_plan_batch_kernel<<<n_samples, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  FO1_LAUNCH_CHECK();
  return FO1_OK;
}

extern "C" int fo1_parse_predictions(const int32_t* tokens, int64_t ld, const int32_t* lens, int32_t n_samples, int32_t ground_start_id,
                                     int32_t ground_end_id, int32_t objects_start_id, int32_t objects_end_id, const int32_t* region_ids,
                                     int32_t n_region_ids, const uint32_t* newline_bitmap, int32_t vocab, int32_t* records, int32_t max_records,
                                     int32_t* n_records, void* stream) {
  FO1_CHECK_ARG(tokens && lens && records && n_records && (n_region_ids == 0 || region_ids), "fo1_parse_predictions: null argument");
  FO1_CHECK_ARG(max_records > 0 && ld >= 0, "fo1_parse_predictions: bad sizes");
  if (n_samples <= 0) return FO1_OK;
  ParseArgs a;
  a.tokens = tokens; a.ld = ld; a.lens = lens; a.B = n_samples;
  a.ground_s = ground_start_id; a.ground_e = ground_end_id; a.objects_s = objects_start_id; a.objects_e = objects_end_id;
  a.region_ids = region_ids; a.n_region_ids = n_region_ids; a.newline_bitmap = newline_bitmap; a.vocab = vocab;
  a.records = records; a.max_records = max_records; a.n_records = n_records;
  parse_predictions_kernel<<<ceil_div(n_samples, 32), 32, 0, static_cast<cudaStream_t>(stream)>>>(a);
  FO1_LAUNCH_CHECK();
  return FO1_OK;
}
