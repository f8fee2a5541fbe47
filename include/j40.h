/*
 * j40.h -- drop-in public header of j40-b200, the Blackwell-native JPEG XL group decoder.
 *
 * Declares the same public API (types, constants, functions) as lifthrasiir/j40's j40.h
 * (reference j40.h:171-272), so that programs written against the reference -- including its own
 * dj40.c, which defines J40_IMPLEMENTATION and includes this header twice (dj40.c:3-6) -- compile
 * unchanged and link against libj40b200.so. The implementation macros of the reference
 * (J40_IMPLEMENTATION, J40_CONFIRM_THAT_THIS_IS_EXPERIMENTAL_AND_POTENTIALLY_UNSAFE, J40_DEBUG, ...)
 * are accepted and ignored: there is no header-only implementation here.
 *
 * Each prototype cites the reference declaration it replaces.
 */
#ifndef J40_B200_J40_H_INCLUDED
#define J40_B200_J40_H_INCLUDED

#define J40_VERSION 2270 /* API level of the reference this header mirrors (j40.h:77) */

#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef J40_IMPLEMENTATION
#include <stdio.h> /* the reference pulls these in for implementation users (j40.h:94-100); dj40.c relies on it */
#include <string.h>
#include <errno.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#ifndef J40_API
#define J40_API
#endif

/* error code; non-zero = failure, readable as four ASCII characters (j40.h:171-172) */
typedef uint32_t j40_err;
#define J40_MIN_RESERVED_ERR (j40_err) (1 << 24)

/* caller-allocated image handle (j40.h:174-182); 16 bytes on LP64 */
typedef struct {
	uint32_t magic;
	union {
		struct j40__inner *inner;
		j40_err err;
		int saved_errno;
	} u;
} j40_image;

/* frame handle, returned by value (j40.h:184-188) */
typedef struct {
	uint32_t magic;
	uint32_t reserved;
	struct j40__inner *inner;
} j40_frame;

typedef void (*j40_memory_free_func)(void *data); /* j40.h:190 */

#define J40_U8X4 0x0f33 /* j40.h:202 */
#define J40_RGBA 0x1755 /* j40.h:228 */

J40_API j40_err j40_error(const j40_image *image);                 /* j40.h:233 */
J40_API const char *j40_error_string(const j40_image *image);      /* j40.h:234 */

/* does not copy `buf`: it must stay valid until j40_free, which calls freefunc(buf) if given (j40.h:236) */
J40_API j40_err j40_from_memory(j40_image *image, void *buf, size_t size, j40_memory_free_func freefunc);
J40_API j40_err j40_from_file(j40_image *image, const char *path); /* j40.h:237 */

/* only (J40_RGBA, J40_U8X4) is accepted, like the reference (j40.h:239, 8363-8375) */
J40_API j40_err j40_output_format(j40_image *image, int32_t channel, int32_t format);

J40_API int j40_next_frame(j40_image *image);                      /* j40.h:241; returns 1 once, then 0 */
J40_API j40_frame j40_current_frame(j40_image *image);             /* j40.h:242 */

typedef struct {
	int32_t width, height;
	int32_t stride_bytes;
	const void *data; /* host memory, rows 32-byte aligned, valid until j40_free */
} j40_pixels_u8x4;                                                 /* j40.h:244-249 */
J40_API j40_pixels_u8x4 j40_frame_pixels_u8x4(const j40_frame *frame, int32_t channel); /* j40.h:250 */

typedef uint8_t j40_u8x4[4];                                       /* j40.h:253 */
typedef float j40_f32x4[4];                                        /* j40.h:256 */
J40_API const j40_u8x4 *j40_row_u8x4(j40_pixels_u8x4 pixels, int32_t y); /* j40.h:251 */

J40_API void j40_free(j40_image *image);                           /* j40.h:272 */

#ifdef __cplusplus
}
#endif

#endif /* J40_B200_J40_H_INCLUDED */
