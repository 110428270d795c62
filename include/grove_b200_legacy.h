/* grove_b200 — TEST-ONLY entry points: the warp-level mma.sync forward attention kernels of round 1 (csrc/attention.cu), built into
 * libgrove_b200_legacy.so and loaded by the parity tests as an independent cross-check of the tcgen05 kernels.  The product library
 * (libgrove_b200.so) does not contain them and the modules never call them. */
#ifndef GROVE_B200_LEGACY_H
#define GROVE_B200_LEGACY_H
#include "grove_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
/* Windowed attention with decomposed rel-pos bias on the UNPARTITIONED token-major qkv[F,G,G,3,heads,hd]
 * (bf16): window_partition's zero padding after norm1 (image_encoder.py:245-249,344-348) is reproduced by
 * giving pad tokens k = b_k, v = b_v (qkv_bias, bf16) — they receive softmax mass like in the reference —
 * and window_unpartition's crop (:382-383) by not computing pad queries.  rel_pos_h/w: [2*ws-1, hd] bf16.
 * out[F,G,G,heads*hd] bf16.  Replaces Attention.forward :301-326 + add_decomposed_rel_pos :420-458.
 * (legacy warp-level mma.sync implementation, kept as an independent cross-check for the tests) */
int grove_attn_window_relpos_fwd(const void* qkv, const void* qkv_bias_bf16, const void* rel_pos_h, const void* rel_pos_w,
                                 void* out, int F, int G, int heads, int hd, int ws, grove_stream_t stream);
/* Same contract on the legacy warp-level tensor path (mma.sync flash kernel, attention.cu) — kept as an independent
 * cross-check for the tests; the modules never call it. */
int grove_attn_global_relpos_fwd_mma(const void* qkv, const void* rel_pos_h, const void* rel_pos_w, void* out, int F, int G,
                                     int heads, int hd, grove_stream_t stream);
#ifdef __cplusplus
}
#endif
#endif
