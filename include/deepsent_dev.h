/* deepsent_dev.h — development-only entry points of libdeepsent_dev.so (built with -DDS_DEV next to the product library).
 * Nothing on the product path calls these; libdeepsent.so does not export them.  They exist for the kernel tuning tools under
 * tools/ and for the per-kernel tests that force a launch mode (tests/test_split_gpu.py). */
#ifndef DEEPSENT_DEV_H_
#define DEEPSENT_DEV_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* launch-policy overrides (0 = default): key 0 = im2col base-coordinate convention of ds_conv_tc, 1 = force N tile, 2 = force
 * ring stages, 3 = shared-memory budget per CTA in KB (ds_conv_tc), 7 = output rows per stem band, 8 = 1: generic (per-pixel
 * gather) pooling kernels instead of the row-walking / 2x2-block ones, 9 = rows per segment of the row-walking pool kernels,
 * 10 = CTA-pair mode of ds_conv_bf16x3 (1 force pairs, 2 force single CTAs), 11 = 3x3 staging of ds_conv_bf16x3 (1 force the
 * per-tap im2col TMA path, 2 force halo-tile staging), 12 = 1: single epilogue staging tile in the stem kernel */
int ds_debug_set(int key, int value);
int ds_debug_get(int key);

/* hardware probe (csrc/probe.cu): D[128,64] = rows [row_shift, row_shift + 128) of the 128B-swizzled shared-memory tile
 * A[256,64] (bf16) times B[64,64]^T, with the UMMA descriptor's base-offset field left 0 (mode 0) or set to (start >> 7) & 7
 * (mode 1) - decides whether 3x3 taps can be read as shifted views of one staged halo tile */
int ds_probe_umma_row_shift(const uint16_t* a, const uint16_t* b, int row_shift, int mode, float* d, void* stream);

/* hardware probe 2 (csrc/probe.cu): TMA cycles per box transfer for the access patterns of the contraction kernels - mode 0 2-D tiled
 * load (64 x 128 rows), 1 4-D tiled halo load, 2 the same halo tile through an im2col-mode load with a padded bounding box, 3 the
 * im2col kernel's 128-pixel tap load, 4 2-D tiled store, 5 4-D tiled clipped store.  `buf` is a device buffer viewed as a
 * [images, h, w, ld] bf16 activation; out[grid] receives cycles per transfer of each CTA. */
int ds_probe_tma_rate(void* buf, int mode, int64_t images, int64_t h, int64_t w, int64_t ld, int reps, int grid, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPSENT_DEV_H_ */
