/*
 * j40b.h -- batch / device-resident extension of the j40 API (SURVEY.md §8b "Batch extension").
 *
 * The ten j40_* functions decode one image per handle into host memory. Throughput work (BASELINE
 * configs C3/C5: batches of independent 4K frames, sharded over GPUs) needs many images per kernel
 * launch and pixels that can stay in HBM, so this header adds a batch object. Nothing here changes the
 * behaviour of the j40_* functions. One batch lives on one CUDA device; shard a list of images over
 * several devices by creating one batch per device/process (images are independent: no collective).
 */
#ifndef J40_B200_J40B_H_INCLUDED
#define J40_B200_J40B_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct j40b_batch j40b_batch;

/* creates a batch on CUDA device `device` (ordinal). Returns NULL if no usable GPU exists. */
j40b_batch *j40b_batch_create(int device);
void j40b_batch_destroy(j40b_batch *b);

/* host-side parse of one image (container, headers, TOC, LfGlobal, HfGlobal: the part the reference does in
 * j40.h:8175-8192). `buf` must stay valid until j40b_batch_wait has returned for the decode that uses it (or the
 * batch is reset / destroyed). Returns the image index or -1. */
int j40b_batch_add(j40b_batch *b, const void *buf, size_t size);

/* the same for n images at once, parsed on `threads` host threads (<= 0: all cores, at most 16); the thread
 * count is also used for the copies into the staging buffer at upload. Returns the index of the first. */
int j40b_batch_add_many(j40b_batch *b, const void *const *bufs, const size_t *sizes, int n, int threads);

/* lays the added images out and enqueues ONE H2D transfer of codestreams + tables on the batch's stream
 * (asynchronous; the staging buffer is pinned memory owned by the batch). 0 on success. */
int j40b_batch_upload(j40b_batch *b);

/* forgets all images but keeps the device and pinned allocations, so that a batch object can be reused for
 * the next set of images without cudaMalloc / cudaHostAlloc. Waits for work in flight. */
int j40b_batch_reset(j40b_batch *b);

/* enqueues all decode kernels (asynchronous). May be called repeatedly on the same uploaded batch. */
int j40b_batch_decode(j40b_batch *b);

/* waits for the kernels and gathers per-image error codes. Returns the number of failed images. */
int j40b_batch_wait(j40b_batch *b);

int j40b_batch_count(const j40b_batch *b);
uint32_t j40b_batch_error(const j40b_batch *b, int index);            /* four-character code or 0 */
int j40b_batch_info(const j40b_batch *b, int index, int32_t *width, int32_t *height, int32_t *stride_bytes);
const void *j40b_batch_device_pixels(const j40b_batch *b, int index); /* RGBA8 in device memory */
int j40b_batch_read_pixels(j40b_batch *b, int index, void *dst);     /* D2H of stride*height bytes */
/* enqueues, behind the decode kernels on the batch's stream, the D2H copy of every image to dst + index*pitch
 * (stride*height bytes each, clipped to pitch). Asynchronous when dst is pinned; complete after j40b_batch_wait.
 * With several batch objects in flight the copies of one overlap the kernels of the others. */
int j40b_batch_read_all_async(j40b_batch *b, void *dst, size_t pitch);

/* device time in milliseconds of the most recent j40b_batch_decode, measured with CUDA events on the
 * batch's stream (valid after j40b_batch_wait) */
float j40b_batch_last_decode_ms(const j40b_batch *b);
/* per-kernel device times of the last decode: 0 lf_group (= 6 + 7 + 8), 1 hf_group, 2 back, 3 back_big, 4 modular,
 * 5 render, 6 k_lf_decode<1> (LF image), 7 k_lf_post + k_lf_decode<2> (HF metadata), 8 k_lf_llf */
float j40b_batch_kernel_ms(const j40b_batch *b, int which);

/* statistics: 0 device bytes allocated, 1 bytes uploaded (H2D), 2 kernels launched by the last decode,
 * 3 compressed bytes, 4 pixels, 5 lanes per LF group in the last decode's serial LF kernels (32, 16 or 8; 1: one
 * LF group per lane) */
int64_t j40b_batch_stat(const j40b_batch *b, int what);

/* Timing several batches in flight at once (each batch owns a CUDA stream; decodes of different batches
 * overlap, e.g. the latency-bound LF-group kernel of one with the HF/back kernels of another):
 *   j40b_batch_mark(b0, 0); enqueue decodes on any batches; j40b_batch_join(b0, bi) for every other batch;
 *   j40b_batch_mark(b0, 1); wait for all batches; j40b_batch_mark_ms(b0) = device time of the region. */
int j40b_batch_mark(j40b_batch *b, int which);
int j40b_batch_join(j40b_batch *b, j40b_batch *other);
float j40b_batch_mark_ms(j40b_batch *b);

/* Phase control for a serving loop with several batches in flight: work enqueued on b after this call waits
 * until `other`'s most recently enqueued decode has finished stage `stage` (0 LF-group stage, 1 pass-group
 * entropy decode, 2 everything). Keeping half of the batches one stage behind the others lets the serial,
 * latency-bound LF decoders of one half run under the throughput kernels of the other half. */
int j40b_batch_after(j40b_batch *b, j40b_batch *other, int stage);

/* diagnostic: device time in ms from `ref`'s mark 0 to event `which` of b's most recent decode (0 LF start,
 * 5 LF image decoded, 6 HF metadata decoded, 1 LF stage done, 2 HF done, 3 tiles done, 4 all done); -1 if unknown */
float j40b_batch_event_ms(const j40b_batch *b, const j40b_batch *ref, int which);

/* diagnostics: copies an intermediate array of LF group `lf_group` of VarDCT image `index` (after a decode) to dst, in
 * the layout of the reference's j40__lf_group_st members (j40.h:6360-6391): what = 0 blocks, 1 varblocks, 2 lfindices,
 * 3/4/5 llfcoeffs X/Y/B, 6/7/8 coeffs as decoded, 9/10 xfromy/bfromy, 11 sharpness, 12/13/14 coeffs after
 * j40__dequant_hf, 15/16/17 LF planes after smoothing. Returns the bytes written (0: unavailable / cap too small).
 * The parity tests compare these with the reference's own arrays (float intermediates within 1e-5). */
size_t j40b_batch_debug_dump(j40b_batch *b, int index, int lf_group, int what, void *dst, size_t cap);

/* writes image `index` (after a decode) as a PAM file (P7, RGB_ALPHA) straight from device memory, rows without the stride
 * padding: the counterpart of dj40.c's stbi_write_png for device-resident output. 0 on success. */
int j40b_batch_write_pam(j40b_batch *b, int index, const char *path);

/* 1 if a CUDA device is usable by this library, else 0 */
int j40b_gpu_available(void);

#ifdef __cplusplus
}
#endif

#endif
