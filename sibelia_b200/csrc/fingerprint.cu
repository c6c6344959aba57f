// k > 32: per-position canonical fingerprints + lexicographic ranking of the vertex k-mers (stub, filled in below)
#include "context.h"

namespace sibgpu {
struct TextDesc;
int fingerprint_positions(sibgpu_ctx *, const TextDesc &, uint32_t, uint32_t)
{
	set_error("invalid: k > 32 not implemented yet");
	return SIBGPU_ERR_INVALID;
}
int rank_fingerprint_vertices(sibgpu_ctx *, const TextDesc &, uint32_t, uint64_t, uint32_t, uint32_t, uint32_t, uint32_t *, bool *)
{
	set_error("invalid: k > 32 not implemented yet");
	return SIBGPU_ERR_INVALID;
}
}
