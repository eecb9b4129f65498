// llm.cu -- Qwen2.5 decoder: prefill / decode (placeholder until the LLM stage lands).
#include "engine.cuh"

namespace fo1 {
int llm_finalize(Model* m) { (void)m; return FO1_OK; }
void llm_destroy_state(Model* m) { (void)m; }
}  // namespace fo1
