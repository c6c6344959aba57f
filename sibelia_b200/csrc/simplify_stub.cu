#include "context.h"
extern "C" int sibgpu_simplify(sibgpu_ctx *, char **, uint32_t **, uint64_t *, uint32_t, uint32_t, uint32_t, uint32_t,
	sibgpu_progress_fn, void *, uint64_t *)
{
	sibgpu::set_error("invalid: sibgpu_simplify not implemented yet");
	return SIBGPU_ERR_INVALID;
}
