// engine.cu -- model handle lifecycle, weight table, arena.
#include "engine.cuh"

namespace fo1 {

int arena_ensure(Model* m, size_t bytes) {
  if (bytes <= m->arena.cap) return FO1_OK;
  // growth happens only when a larger problem than ever before is planned (not on the steady-state path)
  FO1_CUDA(cudaDeviceSynchronize());
  if (m->arena.base) FO1_CUDA(cudaFree(m->arena.base));
  m->arena.base = nullptr;
  m->arena.cap = 0;
  const size_t want = align_up(bytes + (bytes >> 4), 1 << 21);
  FO1_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->arena.base), want));
  m->arena.cap = want;
  return FO1_OK;
}

// Eviction happens ONLY here, at the entry of a forward (before its first lookup): callers hold the raw device pointers
// of several tables across one forward, so freeing inside cached_ints() would hand a kernel a dangling table.
int int_cache_trim(Model* m, size_t keep_below) {
  if (m->int_cache.size() < keep_below) return FO1_OK;
  FO1_CUDA(cudaDeviceSynchronize());   // tables of earlier forwards may still be read by kernels in flight
  for (auto& kv : m->int_cache) cudaFree(kv.second.dev);
  m->int_cache.clear();
  return FO1_OK;
}

int cached_ints(Model* m, const std::string& key, const std::vector<int>& host, const int** dev, cudaStream_t s) {
  auto it = m->int_cache.find(key);
  if (it == m->int_cache.end()) {
    DeviceInts d;
    d.n = host.size();
    FO1_CUDA(cudaMalloc(reinterpret_cast<void**>(&d.dev), (host.size() + 1) * sizeof(int)));
    FO1_CUDA(cudaMemcpyAsync(d.dev, host.data(), host.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    FO1_CUDA(cudaStreamSynchronize(s));  // `host` may be a temporary
    it = m->int_cache.emplace(key, d).first;
  }
  *dev = it->second.dev;
  return FO1_OK;
}

}  // namespace fo1

using namespace fo1;


extern "C" int fo1_int_cache_entries(fo1_model* m) { return m ? (int)m->int_cache.size() : 0; }
extern "C" int fo1_last_decode_path(fo1_model* m) { return m ? m->last_decode_path : -1; }

extern "C" int fo1_model_create(const fo1_model_config* cfg, fo1_model** out) {
  FO1_CHECK_ARG(cfg && out, "fo1_model_create: null argument");
  fo1_model* m = new (std::nothrow) fo1_model();
  if (!m) { set_error("fo1_model_create: out of host memory"); return FO1_ERR_STATE; }
  m->cfg = *cfg;
  *out = m;
  return FO1_OK;
}

namespace fo1 { void llm_destroy_state(Model* m); }

extern "C" void fo1_model_destroy(fo1_model* m) {
  if (!m) return;
  cudaDeviceSynchronize();
  fo1::llm_destroy_state(m);
  if (m->arena.base) cudaFree(m->arena.base);
  if (m->kv_cache) cudaFree(m->kv_cache);
  for (auto& kv : m->int_cache) cudaFree(kv.second.dev);
  delete m;
}

extern "C" int fo1_model_set_weight(fo1_model* m, const char* name, const void* dev_ptr, int32_t dtype, int32_t ndim,
                                    const int64_t* shape) {
  FO1_CHECK_ARG(m && name && dev_ptr && (ndim == 0 || shape), "fo1_model_set_weight: null argument");
  FO1_CHECK_ARG(ndim >= 0 && ndim <= 8, "fo1_model_set_weight(%s): ndim %d", name, ndim);
  FO1_CHECK_ARG((reinterpret_cast<uintptr_t>(dev_ptr) & 15) == 0, "fo1_model_set_weight(%s): pointer not 16-byte aligned", name);
  WeightRef w;
  w.ptr = dev_ptr;
  w.dtype = dtype;
  w.shape.assign(shape, shape + ndim);
  m->weights[name] = w;
  m->finalized = false;
  return FO1_OK;
}
