/* surtr_debug.h -- development aids exported by libsurtr_b200.so, NOT part of the drop-in ABI. */
#ifndef SURTR_DEBUG_H
#define SURTR_DEBUG_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
struct surtr_ctx;
/* K3 writes 8 words per candidate pair: cycles {load, clip, moments, write}, sequential cuts, cuts, V_in, planes. */
int surtr_debug_enable(struct surtr_ctx* ctx, int on);
int surtr_debug_read(struct surtr_ctx* ctx, uint32_t* out, uint64_t n_cand);
/* The cudaStream_t the asynchronous downloads run on (timeline tools record their own events on it). */
void* surtr_debug_copy_stream(struct surtr_ctx* ctx);
#ifdef __cplusplus
}
#endif
#endif
