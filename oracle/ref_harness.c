/*
 * TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or executed from the product
 * (libj40b200.so / the j40_b200 package).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load the library built from this file.
 *
 * This translation unit compiles the UNMODIFIED reference decoder (lifthrasiir/j40, j40.h) straight
 * from where it lies (-DJ40_REF_HEADER="/root/reference/j40.h"); nothing of it is copied into this
 * repository.  All reference symbols are made static (J40_API=static) and re-exported under a
 * `ref_` prefix so that the oracle can live in the same process as the product library.
 *
 * Besides the plain public-API decode (ref_decode_rgba) it replays the reference's own call
 * sequence (j40.h:8146-8220, j40__advance) step by step so that tests can look at the reference's
 * intermediate state (varblock map, LLF coefficients, quantised/dequantised HF coefficients, LF
 * indices, CfL maps, i16 planes) -- these are the function-level parity pins SURVEY.md §4 asks for.
 *
 * Build flags are pinned by oracle/Makefile: -O3 -ffp-contract=off, no -march.
 */
#define J40_CONFIRM_THAT_THIS_IS_EXPERIMENTAL_AND_POTENTIALLY_UNSAFE
#define J40_IMPLEMENTATION
#define J40_API static
#ifndef J40_REF_HEADER
#error "define J40_REF_HEADER to the absolute path of the reference j40.h"
#endif
#define J40_FILENAME J40_REF_HEADER
#include J40_REF_HEADER

#include <time.h>

#define REF_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------- */
/* public-API decode: exactly what dj40.c does (dj40.c:29-49), from memory                      */

REF_EXPORT int ref_decode_rgba(
	const uint8_t *buf, size_t size,
	uint8_t **out_pixels, int32_t *out_w, int32_t *out_h, int32_t *out_stride,
	uint32_t *out_err, char *out_errstr /* >= 256 bytes or NULL */
) {
	j40_image image;
	int ok = 0;
	*out_pixels = NULL;
	*out_w = *out_h = *out_stride = 0;
	j40_from_memory(&image, (void *) buf, size, NULL);
	j40_output_format(&image, J40_RGBA, J40_U8X4);
	if (j40_next_frame(&image)) {
		j40_frame frame = j40_current_frame(&image);
		j40_pixels_u8x4 pixels = j40_frame_pixels_u8x4(&frame, J40_RGBA);
		size_t total = (size_t) pixels.stride_bytes * (size_t) pixels.height;
		*out_pixels = (uint8_t *) malloc(total ? total : 1);
		if (*out_pixels) {
			memcpy(*out_pixels, pixels.data, total);
			*out_w = pixels.width;
			*out_h = pixels.height;
			*out_stride = pixels.stride_bytes;
			ok = 1;
		}
	}
	*out_err = j40_error(&image);
	if (out_errstr) {
		if (*out_err) {
			strncpy(out_errstr, j40_error_string(&image), 255);
			out_errstr[255] = 0;
		} else {
			out_errstr[0] = 0;
		}
	}
	j40_free(&image);
	return ok;
}

REF_EXPORT void ref_free(void *p) { free(p); }

static double ref_now(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

/* times `reps` complete decodes through the public API (from_memory .. frame_pixels .. free);
 * returns the number of successful decodes, fills seconds[i] for each repetition. */
REF_EXPORT int ref_time_decode(const uint8_t *buf, size_t size, int reps, double *seconds) {
	int i, good = 0;
	for (i = 0; i < reps; ++i) {
		j40_image image;
		double t0 = ref_now();
		volatile uint8_t sink = 0;
		j40_from_memory(&image, (void *) buf, size, NULL);
		j40_output_format(&image, J40_RGBA, J40_U8X4);
		if (j40_next_frame(&image)) {
			j40_frame frame = j40_current_frame(&image);
			j40_pixels_u8x4 pixels = j40_frame_pixels_u8x4(&frame, J40_RGBA);
			sink = (uint8_t) (sink ^ ((const uint8_t *) pixels.data)[0]);
			if (!j40_error(&image)) ++good;
		}
		j40_free(&image);
		seconds[i] = ref_now() - t0;
	}
	return good;
}

/* ------------------------------------------------------------------------------------------- */
/* staged decode: replays j40__advance (j40.h:8146-8220) without the coroutine, keeping the      */
/* reference's internal state alive so that intermediates can be copied out                      */

typedef struct {
	j40__inner *inner;
	j40__st st;
	int stage; /* 0 = sections decoded (quantised coeffs), 1 = dequantised, 2 = combined+rendered */
	j40_err err;
} ref_staged;

#define REF_TRY(expr) do { if ((expr)) goto fail; } while (0)

REF_EXPORT ref_staged *ref_staged_open(const uint8_t *buf, size_t size, uint32_t *out_err) {
	ref_staged *s = (ref_staged *) calloc(1, sizeof(ref_staged));
	j40__st *st;
	j40__frame_st *f;
	j40__inner *inner;
	if (!s) return NULL;
	inner = s->inner = (j40__inner *) calloc(1, sizeof(j40__inner));
	if (!inner) { free(s); return NULL; }
	inner->magic = J40__INNER_MAGIC;
	st = &s->st;
	j40__init_state(st, inner);
	REF_TRY(j40__init_memory_source(st, (uint8_t *) buf, size, NULL, &inner->source));
	f = st->frame;
	REF_TRY(j40__init_buffer(st, 0, INT64_MAX));
	REF_TRY(j40__signature(st));
	REF_TRY(j40__image_metadata(st));
	if (st->image->want_icc) REF_TRY(j40__icc(st));
	REF_TRY(j40__frame_header(st));
	if (!f->is_last) REF_TRY(J40__ERR("TODO: multiple frames"));
	if (f->type != J40__FRAME_REGULAR) REF_TRY(J40__ERR("TODO: non-regular frame"));
	REF_TRY(j40__read_toc(st, &inner->toc));
	REF_TRY(j40__lf_global_in_section(st, &inner->toc));
	REF_TRY(j40__hf_global_in_section(st, &inner->toc));
	REF_TRY(j40__allocate_lf_groups(st, &inner->lf_groups));
	if (inner->toc.single_size) {
		REF_TRY(j40__lf_group(st, &inner->lf_groups[0]));
		inner->lf_groups[0].loaded = 1;
		REF_TRY(j40__prepare_dq_matrices(st));
		REF_TRY(j40__prepare_orders(st));
		REF_TRY(j40__pass_group(st, 0, 0, 0, f->width, f->height, 0, &inner->lf_groups[0]));
		REF_TRY(j40__zero_pad_to_byte(st));
	} else {
		while (inner->toc.nsections_read < inner->toc.nsections) {
			REF_TRY(j40__lf_or_pass_group_in_section(st, &inner->toc, inner->lf_groups));
		}
	}
	REF_TRY(j40__end_of_frame(st, &inner->toc));
	REF_TRY(j40__inverse_transform(st, &f->gmodular));
	s->stage = 0;
	*out_err = 0;
	return s;
fail:
	s->err = st->err;
	*out_err = st->err;
	return s;
}

REF_EXPORT uint32_t ref_staged_error(const ref_staged *s) { return s->err; }

/* frame-level facts: out[0..15] */
REF_EXPORT void ref_staged_info(const ref_staged *s, int64_t *out) {
	const j40__frame_st *f = &s->inner->frame;
	const j40__image_st *im = &s->inner->image;
	out[0] = f->width; out[1] = f->height; out[2] = f->is_modular; out[3] = f->group_size_shift;
	out[4] = f->num_groups; out[5] = f->num_lf_groups; out[6] = f->num_passes;
	out[7] = f->global_scale; out[8] = f->quant_lf; out[9] = f->nb_block_ctx;
	out[10] = f->num_hf_presets; out[11] = im->bpp; out[12] = im->xyb_encoded;
	out[13] = f->dct_select_used; out[14] = f->order_used; out[15] = im->num_extra_channels;
}

/* LF-group geometry: out[0..7] = left, top, width, height, width8, height8, nb_varblocks, loaded */
REF_EXPORT int ref_staged_lf_group_info(const ref_staged *s, int64_t ggidx, int32_t *out) {
	const j40__lf_group_st *gg;
	if (s->err || ggidx < 0 || ggidx >= s->inner->frame.num_lf_groups) return 0;
	gg = &s->inner->lf_groups[ggidx];
	out[0] = gg->left; out[1] = gg->top; out[2] = gg->width; out[3] = gg->height;
	out[4] = gg->width8; out[5] = gg->height8; out[6] = gg->nb_varblocks; out[7] = gg->loaded;
	return 1;
}

/* copies a per-LF-group array; `what`:
 *   0 blocks        int32 [height8][width8]          (j40.h:6374-6377)
 *   1 varblocks     int32 pairs {coeffoff_qfidx, bits of hfmul.inv}  [nb_varblocks][2]
 *   2 lfindices     uint8 [height8][width8]
 *   3/4/5 llfcoeffs float [width8*height8]            for X/Y/B
 *   6/7/8 coeffs    float [width8*height8*64]         for X/Y/B (quantised at stage 0, dequantised at stage 1)
 *   9/10 xfromy/bfromy int16 [height64][width64]
 *   11 sharpness    int16 [height8][width8]
 * returns the number of bytes written (0 on error); `cap` is the capacity of `out` in bytes. */
REF_EXPORT int64_t ref_staged_lf_group_array(const ref_staged *s, int64_t ggidx, int what, void *out, int64_t cap) {
	const j40__lf_group_st *gg;
	int64_t n = 0;
	int32_t y;
	if (s->err || ggidx < 0 || ggidx >= s->inner->frame.num_lf_groups) return 0;
	gg = &s->inner->lf_groups[ggidx];
	if (!gg->loaded) return 0;
	switch (what) {
	case 0:
		n = (int64_t) gg->width8 * gg->height8 * 4;
		if (n > cap) return 0;
		for (y = 0; y < gg->height8; ++y) {
			memcpy((char *) out + (size_t) y * (size_t) gg->width8 * 4, J40__I32_PIXELS(&gg->blocks, y), (size_t) gg->width8 * 4);
		}
		return n;
	case 1:
		n = (int64_t) gg->nb_varblocks * 8;
		if (n > cap) return 0;
		memcpy(out, gg->varblocks, (size_t) n);
		return n;
	case 2:
		n = (int64_t) gg->width8 * gg->height8;
		if (n > cap) return 0;
		for (y = 0; y < gg->height8; ++y) {
			memcpy((char *) out + (size_t) y * (size_t) gg->width8, J40__U8_PIXELS(&gg->lfindices, y), (size_t) gg->width8);
		}
		return n;
	case 3: case 4: case 5:
		n = (int64_t) gg->width8 * gg->height8 * 4;
		if (n > cap) return 0;
		memcpy(out, gg->llfcoeffs[what - 3], (size_t) n);
		return n;
	case 6: case 7: case 8:
		n = (int64_t) gg->width8 * gg->height8 * 64 * 4;
		if (n > cap) return 0;
		memcpy(out, gg->coeffs[what - 6], (size_t) n);
		return n;
	case 9: case 10: {
		const j40__plane *p = what == 9 ? &gg->xfromy : &gg->bfromy;
		n = (int64_t) p->width * p->height * 2;
		if (n > cap || p->type != J40__PLANE_I16) return 0;
		for (y = 0; y < p->height; ++y) {
			memcpy((char *) out + (size_t) y * (size_t) p->width * 2, J40__I16_PIXELS(p, y), (size_t) p->width * 2);
		}
		return n;
	}
	case 11: {
		const j40__plane *p = &gg->sharpness;
		n = (int64_t) p->width * p->height * 2;
		if (n > cap || p->type != J40__PLANE_I16) return 0;
		for (y = 0; y < p->height; ++y) {
			memcpy((char *) out + (size_t) y * (size_t) p->width * 2, J40__I16_PIXELS(p, y), (size_t) p->width * 2);
		}
		return n;
	}
	default: return 0;
	}
}

/* advances the staged decode: stage 1 = j40__dequant_hf on every LF group (j40.h:7053);
 * stage 2 = the rest of j40__combine_vardct (j40.h:7862) + j40__no_more_bytes.  returns err. */
REF_EXPORT uint32_t ref_staged_advance(ref_staged *s, int to_stage) {
	j40__st *st = &s->st;
	j40__frame_st *f = st->frame;
	int64_t i;
	if (s->err) return s->err;
	if (f->is_modular) { s->stage = to_stage; return 0; }
	if (s->stage < 1 && to_stage >= 1) {
		for (i = 0; i < f->num_lf_groups; ++i) j40__dequant_hf(st, &s->inner->lf_groups[i]);
		s->stage = 1;
	}
	if (s->stage < 2 && to_stage >= 2) {
		/* j40__combine_vardct minus the dequantisation already done above */
		if (f->do_ycbcr || st->image->cspace == J40__CS_GREY) { s->err = J40__4("TODO"); return s->err; }
		f->gmodular.num_channels = 3;
		f->gmodular.channel = (j40__plane *) calloc(3, sizeof(j40__plane));
		for (i = 0; i < 3; ++i) {
			REF_TRY(j40__init_plane(st, J40__PLANE_I16, f->width, f->height, J40__PLANE_FORCE_PAD, &f->gmodular.channel[i]));
		}
		for (i = 0; i < f->num_lf_groups; ++i) {
			REF_TRY(j40__combine_vardct_from_lf_group(st, &s->inner->lf_groups[i]));
		}
		s->stage = 2;
	}
	return 0;
fail:
	s->err = st->err;
	return s->err;
}

/* copies global modular channel `c` (int16, tight rows) after stage 2 (VarDCT) or stage 0 (modular) */
REF_EXPORT int64_t ref_staged_plane_i16(const ref_staged *s, int c, int16_t *out, int64_t cap_elems) {
	const j40__frame_st *f = &s->inner->frame;
	const j40__plane *p;
	int32_t y;
	if (s->err || c < 0 || c >= f->gmodular.num_channels) return 0;
	p = &f->gmodular.channel[c];
	if (p->type != J40__PLANE_I16 || (int64_t) p->width * p->height > cap_elems) return 0;
	for (y = 0; y < p->height; ++y) memcpy(out + (size_t) y * (size_t) p->width, J40__I16_PIXELS(p, y), (size_t) p->width * 2);
	return (int64_t) p->width * p->height;
}

REF_EXPORT void ref_staged_close(ref_staged *s) {
	if (!s) return;
	if (s->inner) {
		s->inner->source.free_func = NULL; /* the caller owns the buffer */
		j40__free_inner(s->inner);
	}
	free(s);
}

/* ------------------------------------------------------------------------------------------- */
/* function-level pins: the reference's own static functions on caller-supplied data            */

/* j40__inverse_dct2d / special 8x8 transforms, dispatched exactly like j40.h:7178-7191.
 * buf holds 1 << (log_rows+log_columns) floats in the reference's coefficient layout. */
REF_EXPORT void ref_inverse_transform(int dctsel, float *buf) {
	const j40__dct_select *dct = &J40__DCT_SELECT[dctsel];
	float *scratch2 = (float *) malloc(sizeof(float) * 65536);
	switch (dctsel) {
	case 1: j40__inverse_hornuss(buf); break;
	case 2: j40__inverse_dct11(buf); break;
	case 3: j40__inverse_dct22(buf); break;
	case 12: j40__inverse_dct23(buf); break;
	case 13: j40__inverse_dct32(buf); break;
	case 14: j40__inverse_afv(buf, 0, 0); break;
	case 15: j40__inverse_afv(buf, 1, 0); break;
	case 16: j40__inverse_afv(buf, 0, 1); break;
	case 17: j40__inverse_afv(buf, 1, 1); break;
	default: j40__inverse_dct2d(buf, scratch2, dct->log_rows, dct->log_columns); break;
	}
	free(scratch2);
}

/* j40__forward_dct2d_scaled_for_llf (j40.h:5944); buf = (1<<log_rows) x (1<<log_columns) floats */
REF_EXPORT void ref_forward_llf(float *buf, int log_rows, int log_columns) {
	float scratch[1024];
	j40__forward_dct2d_scaled_for_llf(buf, scratch, log_rows, log_columns);
}

/* known-answer constants of the reference */
REF_EXPORT void ref_constants(float *half_secants /*256*/, float *lf2llf /*64*/, float *afv /*256, [i*16+j]*/) {
	int i, j;
	memcpy(half_secants, J40__HALF_SECANTS, sizeof(float) * 256);
	memcpy(lf2llf, J40__LF2LLF_SCALES, sizeof(float) * 64);
	/* AFV_BASIS is function-local (j40.h:6108); recover it exactly by feeding unit vectors */
	for (j = 0; j < 16; ++j) {
		float in[16] = {0}, out[16];
		in[j] = 1.0f;
		j40__inverse_afv22(out, in);
		for (i = 0; i < 16; ++i) afv[i * 16 + j] = out[i];
	}
}

/* dequantisation matrix for parameter set idx (0..16) with the library defaults (j40.h:4828);
 * out = rows*columns*3 floats, [i][c]. returns rows*columns or 0. */
REF_EXPORT int ref_default_dq_matrix(int idx, float *out, int cap) {
	j40__st st;
	j40__dq_matrix dq;
	int n, i, c;
	memset(&st, 0, sizeof(st));
	memset(&dq, 0, sizeof(dq));
	dq.mode = J40__DQ_ENC_LIBRARY;
	if (j40__load_dq_matrix(&st, idx, &dq)) return 0;
	n = dq.n * dq.m;
	if (n * 3 > cap) { j40__free_dq_matrix(&dq); return 0; }
	for (i = 0; i < n; ++i) for (c = 0; c < 3; ++c) out[i * 3 + c] = dq.params[i][c];
	j40__free_dq_matrix(&dq);
	return n;
}

/* natural coefficient order (j40.h:4980) */
REF_EXPORT int ref_natural_order(int log_rows, int log_columns, int32_t *out) {
	j40__st st;
	int32_t *order = NULL;
	memset(&st, 0, sizeof(st));
	if (j40__natural_order(&st, log_rows, log_columns, &order)) return 0;
	memcpy(out, order, sizeof(int32_t) << (log_rows + log_columns));
	free(order);
	return 1 << (log_rows + log_columns);
}

/* the sRGB quantiser of j40.h:7233-7235 for bpp bits, as the reference computes it */
REF_EXPORT int32_t ref_srgb_quant(float v, int bpp) {
	v = (v <= 0.0031308f ? 12.92f * v : 1.055f * powf(v, 1.0f / 2.4f) - 0.055f);
	return (int32_t) (int16_t) ((float) ((1 << bpp) - 1) * v + 0.5f);
}
