/* oracle/oracle_video.c — TEST INFRASTRUCTURE (see msb200_oracle.h).
 * NV12/NV21 -> I420 with rotation and nearest 1/2 decimation, restated from
 * copy_ycbcrbiplanar_to_true_yuv_with_rotation_and_down_scale_by_2() /root/reference/src/voip/msvideo.c:787-919 and
 * rotate_plane_down_scale_by_2() :734-776; pinned against the unmodified reference in
 * tests/test_oracle_vs_reference.py (and by the reference's own framework test pattern,
 * tester/mediastreamer2_framework_tester.c:219-367). */
#include "msb200_oracle.h"

#include <string.h>

/* :734-776 — rotate a plane by 90 degrees; step=2 for the interleaved CbCr plane */
static void rotate_plane(int wDest, int hDest, int full_width, const uint8_t *src, uint8_t *dst, int step, int clockwise,
                         int downscale) {
	int factor = downscale ? 2 : 1;
	int hSrc = wDest * factor, wSrc = hDest * factor;
	int src_stride = full_width * step * factor;
	int signed_dst_stride, incr;
	if (clockwise) {
		dst += wDest - 1;
		incr = 1;
		signed_dst_stride = wDest;
	} else {
		dst += wDest * (hDest - 1);
		incr = -1;
		signed_dst_stride = -wDest;
	}
	for (int y = 0; y < hSrc; y += factor) {
		uint8_t *dst2 = dst;
		for (int x = 0; x < step * wSrc; x += step * factor) {
			*dst2 = src[x];
			dst2 += signed_dst_stride;
		}
		dst -= incr;
		src += src_stride;
	}
}

int orc_nv12_to_i420(const uint8_t *y, const uint8_t *cbcr, int rotation, int w, int h, int y_stride, int cbcr_stride,
                     int u_first, int down_scale, uint8_t *out) {
	int uv_w = w / 2, uv_h = h / 2;
	int factor = down_scale ? 2 : 1;
	uint8_t *py = out, *pu = out + (size_t)w * h, *pv = pu + (size_t)uv_w * uv_h; /* ms_yuv_buf_init, tight I420 */
	if (!u_first) { /* :822-826 */
		uint8_t *t = pu;
		pu = pv;
		pv = t;
	}
	if (rotation % 180 == 0) {
		uint8_t *u_dest = pu, *v_dest = pv;
		if (rotation == 0) { /* :832-857 */
			for (int i = 0; i < h; i++) {
				if (down_scale) {
					for (int j = 0; j < w; j++)
						py[i * w + j] = y[i * 2 * y_stride + j * 2];
				} else {
					memcpy(&py[i * w], &y[i * y_stride], (size_t)w);
				}
			}
			for (int i = 0; i < uv_h; i++)
				for (int j = 0; j < uv_w; j++) {
					*u_dest++ = cbcr[cbcr_stride * i * factor + 2 * j * factor];
					*v_dest++ = cbcr[cbcr_stride * i * factor + 2 * j * factor + 1];
				}
		} else { /* 180: :866-880 */
			for (int i = 0; i < h; i++)
				for (int j = 0; j < w; j++)
					py[i * w + j] = y[(h - 1 - i) * y_stride * factor + (w - 1 - j) * factor];
			for (int i = 0; i < uv_h; i++)
				for (int j = 0; j < uv_w; j++) {
					*u_dest++ = cbcr[cbcr_stride * (uv_h - 1 - i) * factor + 2 * (uv_w - 1 - j) * factor];
					*v_dest++ = cbcr[cbcr_stride * (uv_h - 1 - i) * factor + 2 * (uv_w - 1 - j) * factor + 1];
				}
		}
	} else { /* 90 / 270: :883-915 */
		int clockwise = rotation == 90;
		rotate_plane(w, h, y_stride, y, py, 1, clockwise, down_scale);
		rotate_plane(uv_w, uv_h, cbcr_stride / 2, cbcr, pu, 2, clockwise, down_scale);
		rotate_plane(uv_w, uv_h, cbcr_stride / 2, cbcr + 1, pv, 2, clockwise, down_scale);
	}
	return w * h * 3 / 2;
}

/* ================================================================================================================
 * Pixel-format conversion + bilinear scaling: restatement of the swscale pipeline the reference's ffmpeg scaler
 * back-end runs — ff_create_swscale_context() / ff_sws_scale(), /root/reference/src/voip/msvideo.c:651-681:
 *     sws_getContext(src_w, src_h, src_fmt, dst_w, dst_h, dst_fmt, SWS_BILINEAR, NULL, NULL, NULL); sws_scale(...)
 * FFmpeg is an external un-vendored dependency (/root/reference/CMakeLists.txt:242); libswscale 9.1.100 (FFmpeg 8) is
 * available in this container inside the opencv wheel and PINS this restatement: tests/golden/ holds frames produced by
 * the real library (tests/golden/make_swscale_golden.py) and tests/test_oracle_video.py compares against them.
 * Restated pieces: initFilter() (bilinear branch, cut-off reduction, border fix, normalisation with error diffusion),
 * hScale8To15, chroma de-interleave, yuv2planeX_8 / yuv2plane1_8, yuv2rgb_X / yuv2rgb_2 / yuv2rgb_1 templates and the
 * ff_yuv2rgb_c_init_tables() clipping tables (ITU-601, limited range source, brightness 0, contrast/saturation 1.0).
 * ================================================================================================================ */
#include <stdlib.h>

#define PIX_YUV420P 0
#define PIX_RGB24 2
#define PIX_BGR24 3
#define PIX_YUYV 1
#define PIX_RGBA32 7
#define PIX_BGRA32 11
#define PIX_RGB565 8 /* MS_RGB565 -> AV_PIX_FMT_RGB565 (little endian), src/voip/msvideo.c:610-611 */
#define PIX_UYVY 5
#define PIX_YUY2 6
#define PIX_NV12 100
#define PIX_NV21 101

typedef struct {
	int size;
	int32_t *pos;
	int16_t *coef;
} sws_filter;

static int av_log2_i(unsigned v) {
	int n = 0;
	while (v >>= 1) n++;
	return n;
}
static int64_t rounded_div(int64_t a, int64_t b) {
	return ((a > 0) == (b > 0)) ? (a + (b >> 1)) / b : (a - (b >> 1)) / b; /* ROUNDED_DIV */
}

/* libswscale/utils.c initFilter(), SWS_BILINEAR, no src/dst user filters */
static int init_filter(sws_filter *out, int xInc, int srcW, int dstW, int filterAlign, int one, int srcPos, int dstPos) {
	int filterSize, filter2Size, minFilterSize;
	int64_t *filter = NULL;
	int32_t *filterPos = (int32_t *)calloc((size_t)dstW + 3, sizeof(int32_t));
	const int64_t fone = 1LL << (54 - (av_log2_i((unsigned)(srcW / dstW)) < 8 ? av_log2_i((unsigned)(srcW / dstW)) : 8));
	if (abs(xInc - 0x10000) < 10 && srcPos == dstPos) { /* unscaled */
		filterSize = 1;
		filter = (int64_t *)calloc((size_t)dstW * filterSize, sizeof(int64_t));
		for (int i = 0; i < dstW; i++) {
			filter[i * filterSize] = fone;
			filterPos[i] = i;
		}
	} else {
		const int sizeFactor = 2; /* bilinear */
		int64_t xDstInSrc;
		if (xInc <= 1 << 16) filterSize = 1 + sizeFactor;
		else filterSize = 1 + (sizeFactor * srcW + dstW - 1) / dstW;
		if (filterSize > srcW - 2) filterSize = srcW - 2;
		if (filterSize < 1) filterSize = 1;
		filter = (int64_t *)calloc((size_t)dstW * filterSize, sizeof(int64_t));
		xDstInSrc = ((dstPos * (int64_t)xInc) >> 7) - ((srcPos * 0x10000LL) >> 7);
		for (int i = 0; i < dstW; i++) {
			int xx = (int)((xDstInSrc - (filterSize - 2) * (1LL << 16)) / (1 << 17));
			filterPos[i] = xx;
			for (int j = 0; j < filterSize; j++) {
				int64_t d = llabs(((int64_t)xx * (1 << 17)) - xDstInSrc) << 13;
				int64_t coeff;
				if (xInc > 1 << 16) d = d * dstW / srcW;
				coeff = (1 << 30) - d;
				if (coeff < 0) coeff = 0;
				coeff *= fone >> 30;
				filter[i * filterSize + j] = coeff;
				xx++;
			}
			xDstInSrc += 2 * (int64_t)xInc;
		}
	}
	filter2Size = filterSize;
	/* reduce the filter size: shift near-zero leading taps out, count near-zero trailing taps */
	minFilterSize = 0;
	for (int i = dstW - 1; i >= 0; i--) {
		int min = filter2Size;
		int64_t cutOff = 0;
		for (int j = 0; j < filter2Size; j++) {
			cutOff += llabs(filter[i * filter2Size]);
			if ((double)cutOff > 0.002 * (double)fone) break;
			if (i < dstW - 1 && filterPos[i] >= filterPos[i + 1]) break;
			for (int k = 1; k < filter2Size; k++)
				filter[i * filter2Size + k - 1] = filter[i * filter2Size + k];
			filter[i * filter2Size + filter2Size - 1] = 0;
			filterPos[i]++;
		}
		cutOff = 0;
		for (int j = filter2Size - 1; j > 0; j--) {
			cutOff += llabs(filter[i * filter2Size + j]);
			if ((double)cutOff > 0.002 * (double)fone) break;
			min--;
		}
		if (min > minFilterSize) minFilterSize = min;
	}
	if (minFilterSize == 1 && filterAlign == 2) filterAlign = 1; /* x86: special case for unscaled vertical filtering */
	{
		const int newSize = (minFilterSize + (filterAlign - 1)) & (~(filterAlign - 1));
		int64_t *f2 = (int64_t *)calloc((size_t)dstW * newSize, sizeof(int64_t));
		for (int i = 0; i < dstW; i++)
			for (int j = 0; j < newSize; j++)
				f2[i * newSize + j] = j >= filter2Size ? 0 : filter[i * filter2Size + j];
		free(filter);
		filter = f2;
		filterSize = newSize;
	}
	/* fix borders */
	for (int i = 0; i < dstW; i++) {
		if (filterPos[i] < 0) {
			for (int j = 1; j < filterSize; j++) {
				int left = j + filterPos[i] > 0 ? j + filterPos[i] : 0;
				filter[i * filterSize + left] += filter[i * filterSize + j];
				filter[i * filterSize + j] = 0;
			}
			filterPos[i] = 0;
		}
		if (filterPos[i] + filterSize > srcW) {
			int shift = filterPos[i] + (filterSize - srcW < 0 ? filterSize - srcW : 0);
			int64_t acc = 0;
			for (int j = filterSize - 1; j >= 0; j--) {
				if (filterPos[i] + j >= srcW) {
					acc += filter[i * filterSize + j];
					filter[i * filterSize + j] = 0;
				}
			}
			for (int j = filterSize - 1; j >= 0; j--) {
				if (j < shift) filter[i * filterSize + j] = 0;
				else filter[i * filterSize + j] = filter[i * filterSize + j - shift];
			}
			filterPos[i] -= shift;
			filter[i * filterSize + srcW - 1 - filterPos[i]] += acc;
		}
	}
	/* normalise to `one` with error diffusion */
	out->size = filterSize;
	out->pos = filterPos;
	out->coef = (int16_t *)calloc((size_t)(dstW + 3) * filterSize, sizeof(int16_t));
	for (int i = 0; i < dstW; i++) {
		int64_t error = 0, sum = 0;
		for (int j = 0; j < filterSize; j++)
			sum += filter[i * filterSize + j];
		sum = (sum + one / 2) / one;
		if (!sum) sum = 1;
		for (int j = 0; j < filterSize; j++) {
			int64_t v = filter[i * filterSize + j] + error;
			int intV = (int)rounded_div(v, sum);
			out->coef[i * filterSize + j] = (int16_t)intV;
			error = v - intV * sum;
		}
	}
	free(filter);
	return 0;
}

struct orc_scaler {
	int src_w, src_h, src_fmt, dst_w, dst_h, dst_fmt;
	int chr_src_w, chr_src_h, chr_dst_w, chr_dst_h;
	sws_filter hLum, hChr, vLum, vChr;
	/* yuv2rgb tables in closed form */
	int64_t cy, crv, cbu, cgu, cgv, yb0;
	int yoffs;
	int x86_vertical; /* orc_scaler_set_x86_vertical(): planar output as the library's x86 SIMD vertical scaler rounds it */
};

static int get_local_pos(int chr_subsample, int pos) {
	if (pos == -1 || pos <= -513) pos = (128 << chr_subsample) - 128;
	pos += 128;
	return pos >> chr_subsample;
}

static int is_packed422(int fmt) {
	return fmt == PIX_YUYV || fmt == PIX_UYVY || fmt == PIX_YUY2;
}

orc_scaler *orc_scaler_new(int src_w, int src_h, int src_fmt, int dst_w, int dst_h, int dst_fmt) {
	orc_scaler *s;
	if (is_packed422(src_fmt)) {
		/* MSPixConv's job (pixconv.c:62-94): packed 4:2:2 -> YUV420P at the SAME size. libswscale takes its unscaled
		 * special converter for this (yuyvToYuv420Wrapper / uyvyToYuv420Wrapper): luma copied, chroma = rounded average of
		 * the two source lines. Pinned against the real library in tests/golden (cases yuyv422 / uyvy422). */
		if (dst_fmt != PIX_YUV420P || src_w != dst_w || src_h != dst_h || (src_w & 1) || (src_h & 1)) return NULL;
		s = (orc_scaler *)calloc(1, sizeof(*s));
		s->src_w = src_w; s->src_h = src_h; s->src_fmt = src_fmt;
		s->dst_w = dst_w; s->dst_h = dst_h; s->dst_fmt = dst_fmt;
		return s;
	}
	if (src_fmt == PIX_RGB24 || src_fmt == PIX_BGR24 || src_fmt == PIX_RGBA32 || src_fmt == PIX_BGRA32 || src_fmt == PIX_RGB565) {
		/* MSPixConv's RGB inputs (pixconv.c:62-94, MS_RGB24 -> AV_PIX_FMT_RGB24, MS_RGB24_REV -> AV_PIX_FMT_BGR24 read bottom-up
		 * through a negative stride, :78-81): packed RGB -> YUV420P at the SAME size. See rgb_to_i420() below. */
		if (dst_fmt != PIX_YUV420P || src_w != dst_w || src_h != dst_h || (src_w & 1) || (src_h & 1)) return NULL;
		s = (orc_scaler *)calloc(1, sizeof(*s));
		s->src_w = src_w; s->src_h = src_h; s->src_fmt = src_fmt;
		s->dst_w = dst_w; s->dst_h = dst_h; s->dst_fmt = dst_fmt;
		return s;
	}
	const int src_ok = src_fmt == PIX_YUV420P || src_fmt == PIX_NV12 || src_fmt == PIX_NV21;
	const int dst_ok = dst_fmt == PIX_YUV420P || dst_fmt == PIX_RGB24 || dst_fmt == PIX_BGR24;
	if (!src_ok || !dst_ok || src_w < 8 || src_h < 8 || dst_w < 8 || dst_h < 8) return NULL;
	/* an odd RGB output width makes the library switch to full horizontal chroma interpolation (SWS_FULL_CHR_H_INT is forced,
	 * utils.c sws_init_context): another algorithm, not restated — refused here and by msb200_scaler_create alike */
	if (dst_fmt != PIX_YUV420P && (dst_w & 1)) return NULL;
	s = (orc_scaler *)calloc(1, sizeof(*s));
	s->src_w = src_w; s->src_h = src_h; s->src_fmt = src_fmt;
	s->dst_w = dst_w; s->dst_h = dst_h; s->dst_fmt = dst_fmt;
	const int dst_rgb = dst_fmt != PIX_YUV420P;
	/* RGB output without SWS_FULL_CHR_H_INT reuses one chroma sample for two pixels: chrDstHSubSample=1, VSub=0 */
	const int chrDstH = 1, chrDstV = dst_rgb ? 0 : 1;
	s->chr_src_w = (src_w + 1) >> 1;
	s->chr_src_h = (src_h + 1) >> 1;
	s->chr_dst_w = (dst_w + 1) >> chrDstH;
	s->chr_dst_h = (dst_h + (1 << chrDstV) - 1) >> chrDstV;
	const int lumXInc = (int)((((int64_t)src_w << 16) + (dst_w >> 1)) / dst_w);
	const int lumYInc = (int)((((int64_t)src_h << 16) + (dst_h >> 1)) / dst_h);
	const int chrXInc = (int)((((int64_t)s->chr_src_w << 16) + (s->chr_dst_w >> 1)) / s->chr_dst_w);
	const int chrYInc = (int)((((int64_t)s->chr_src_h << 16) + (s->chr_dst_h >> 1)) / s->chr_dst_h);
	/* x86 build of the library: filterAlign 4 horizontally, 2 vertically (padding taps are zero: results unchanged) */
	init_filter(&s->hLum, lumXInc, src_w, dst_w, 4, 1 << 14, get_local_pos(0, 0), get_local_pos(0, 0));
	init_filter(&s->hChr, chrXInc, s->chr_src_w, s->chr_dst_w, 4, 1 << 14, get_local_pos(1, -513), get_local_pos(chrDstH, -513));
	init_filter(&s->vLum, lumYInc, src_h, dst_h, 2, 1 << 12, get_local_pos(0, 0), get_local_pos(0, 0));
	init_filter(&s->vChr, chrYInc, s->chr_src_h, s->chr_dst_h, 2, 1 << 12, get_local_pos(1, -513), get_local_pos(chrDstV, -513));
	/* ff_yuv2rgb_c_init_tables(): ITU-601 coefficients, limited-range source */
	{
		int64_t crv = 104597, cbu = 132201, cgu = -25675, cgv = -53279, cy = 1 << 16, oy;
		cy = (cy * 255) / 219;
		oy = 16 << 16;
		crv = ((crv * (1 << 16)) + 0x8000) / cy;
		cbu = ((cbu * (1 << 16)) + 0x8000) / cy;
		cgu = ((cgu * (1 << 16)) + 0x8000) / cy;
		cgv = ((cgv * (1 << 16)) + 0x8000) / cy;
		s->cy = cy; s->crv = crv; s->cbu = cbu; s->cgu = cgu; s->cgv = cgv;
		s->yb0 = -(384LL << 16) - 512 * cy - oy; /* YUVRGB_TABLE_LUMA_HEADROOM = 512 */
		s->yoffs = 326 + 512;
	}
	return s;
}
/* planar (YUV420P) output only: 0 = the library's C reference arithmetic (SWS_BITEXACT; what the GPU kernels compute),
 * 1 = its x86 SIMD vertical scaler, i.e. what sws_scale returns for the reference's plain SWS_BILINEAR call on x86 */
void orc_scaler_set_x86_vertical(orc_scaler *s, int on) {
	if (s) s->x86_vertical = on;
}
void orc_scaler_free(orc_scaler *s) {
	if (!s) return;
	if (is_packed422(s->src_fmt) || s->src_fmt == PIX_RGB24 || s->src_fmt == PIX_BGR24 || s->src_fmt == PIX_RGBA32 || s->src_fmt == PIX_BGRA32 || s->src_fmt == PIX_RGB565) {
		free(s);
		return;
	}
	sws_filter *f[4] = {&s->hLum, &s->hChr, &s->vLum, &s->vChr};
	for (int i = 0; i < 4; ++i) {
		free(f[i]->pos);
		free(f[i]->coef);
	}
	free(s);
}
static size_t fmt_bytes(int fmt, int w, int h) {
	if (fmt == PIX_RGB24 || fmt == PIX_BGR24) return (size_t)w * h * 3;
	if (fmt == PIX_RGBA32 || fmt == PIX_BGRA32) return (size_t)w * h * 4;
	if (fmt == PIX_YUYV || fmt == PIX_UYVY || fmt == PIX_YUY2 || fmt == PIX_RGB565) return (size_t)w * h * 2;
	return (size_t)w * h + 2 * (size_t)((w + 1) / 2) * ((h + 1) / 2);
}
size_t orc_scaler_src_bytes(orc_scaler *s) { return fmt_bytes(s->src_fmt, s->src_w, s->src_h); }
size_t orc_scaler_dst_bytes(orc_scaler *s) { return fmt_bytes(s->dst_fmt, s->dst_w, s->dst_h); }
int orc_scaler_get_filter(orc_scaler *s, int which, int32_t *pos, int16_t *coef, int max_entries) {
	sws_filter *f = which == 0 ? &s->hLum : which == 1 ? &s->hChr : which == 2 ? &s->vLum : &s->vChr;
	int n = which == 0 ? s->dst_w : which == 1 ? s->chr_dst_w : which == 2 ? s->dst_h : s->chr_dst_h;
	if (n > max_entries) n = max_entries;
	if (pos) memcpy(pos, f->pos, sizeof(int32_t) * (size_t)n);
	if (coef) memcpy(coef, f->coef, sizeof(int16_t) * (size_t)n * f->size);
	return f->size;
}

static inline int clip_u8(int64_t v) {
	return v < 0 ? 0 : (v > 255 ? 255 : (int)v);
}
/* y_table[idx] of ff_yuv2rgb_c_init_tables (bpp 24) in closed form */
static inline uint8_t ytab(const orc_scaler *s, int64_t idx) {
	return (uint8_t)clip_u8((s->yb0 + idx * s->cy + 0x8000) >> 16);
}
static inline void yuv_to_rgb_px(const orc_scaler *s, int Y, int U, int V, uint8_t *r, uint8_t *g, uint8_t *b) {
	const int Uc = clip_u8(U), Vc = clip_u8(V);
	const int64_t ir = s->yoffs - (s->crv >> 9) + ((Vc * s->crv) >> 16);
	const int64_t ig = s->yoffs - (s->cgu >> 9) + ((Uc * s->cgu) >> 16) + (-(s->cgv >> 9) + ((Vc * s->cgv) >> 16));
	const int64_t ib = s->yoffs - (s->cbu >> 9) + ((Uc * s->cbu) >> 16);
	*r = ytab(s, ir + Y);
	*g = ytab(s, ig + Y);
	*b = ytab(s, ib + Y);
}

/* hScale8To15_c */
static void hscale(int16_t *dst, int dstW, const uint8_t *src, const sws_filter *f) {
	for (int i = 0; i < dstW; i++) {
		int val = 0;
		for (int j = 0; j < f->size; j++)
			val += ((int)src[f->pos[i] + j]) * f->coef[f->size * i + j];
		val >>= 7;
		dst[i] = (int16_t)(val < (1 << 15) - 1 ? val : (1 << 15) - 1);
	}
}

/* Packed RGB -> YUV420P at the same size, as libswscale 9.1 (FFmpeg 8; the reference's ffmpeg back-end calls
 * sws_getContext(..., SWS_BILINEAR), /root/reference/src/voip/msvideo.c:651-670) does it. libswscale is not under
 * /root/reference: the arithmetic below restates its published C code and is pinned against the real library
 * (tests/golden cases rgb24 / bgr24; tests/test_oracle_video.py).
 *
 * ITU-601 limited-range coefficients in Q15 (libswscale's input_rgb2yuv_table).
 *
 * RGB24 takes the GENERIC scaler path with an RGB input stage:
 *   luma   rgb24ToY_c:        y14 = (RY r + GY g + BY b + (32 << 14) + (1 << 8)) >> 9
 *          hScale16To15_c:    y15 = min((y14 * 16384) >> 13, 32767)          (unscaled: one tap of 1 << 14, sh = 13)
 *          yuv2plane1_c:      Y   = clip_u8((y15 + 64) >> 7)
 *   chroma rgb24ToUV_half_c:  sums of two horizontally adjacent pixels, u14 = (RU r2 + GU g2 + BU b2 + (256 << 15) + (1 << 9)) >> 10
 *          hScale16To15_c     as above; the chroma plane keeps the full source height (chrSrcVSubSample = 0)
 *          vertical 2:1 bilinear filter of initFilter: taps {512, 1536, 1536, 512} / 4096 on rows 2y-1 .. 2y+2, rows
 *          outside the picture folded onto the border row; yuv2planeX_8_c: clip_u8(((64 << 12) + sum) >> 19)
 * BGR24 takes the unscaled special converter bgr24ToYv12Wrapper -> ff_rgb24toyv12_c:
 *   Y = ((RY r + GY g + BY b) >> 15) + 16 per pixel; U/V from the 2x2 block's component means ((sum of 4) >> 2):
 *   U = ((RU r + GU g + BU b) >> 15) + 128 (arithmetic shifts). */
/* bpp / ro / go / bo: bytes per pixel and byte offsets of R, G, B for the GENERIC path (RGB24: 3,0,1,2; RGBA: 4,0,1,2;
 * BGRA: 4,2,1,0 — the 32-bit formats take the generic path too and give the RGB24 result for equal colours, alpha
 * ignored: verified against the real library); bgr = 1 selects BGR24's special converter instead. */
static void rgb_to_i420(const uint8_t *src, int w, int h, int bgr, int bpp, int ro, int go, int bo, uint8_t *dst, int x86_vertical) {
	enum { RY = 8414, GY = 16519, BY = 3208, RU = -4865, GU = -9528, BU = 14392, RV = 14392, GV = -12061, BV = -2332 };
	uint8_t *dy = dst, *du = dst + (size_t)w * h, *dv = du + (size_t)(w / 2) * (h / 2);
	const int cw = w / 2;
	if (bgr) {
		for (int y = 0; y < h; ++y)
			for (int x = 0; x < w; ++x) {
				const uint8_t *p = src + ((size_t)y * w + x) * 3;
				dy[(size_t)y * w + x] = (uint8_t)(((RY * p[2] + GY * p[1] + BY * p[0]) >> 15) + 16);
			}
		for (int y = 0; y < h / 2; ++y)
			for (int x = 0; x < cw; ++x) {
				int c[3];
				for (int k = 0; k < 3; ++k) {
					const uint8_t *p = src + ((size_t)(2 * y) * w + 2 * x) * 3 + k;
					c[k] = (p[0] + p[3] + p[(size_t)w * 3] + p[(size_t)w * 3 + 3]) >> 2;
				}
				du[(size_t)y * cw + x] = (uint8_t)(((RU * c[2] + GU * c[1] + BU * c[0]) >> 15) + 128);
				dv[(size_t)y * cw + x] = (uint8_t)(((RV * c[2] + GV * c[1] + BV * c[0]) >> 15) + 128);
			}
		return;
	}
	for (int y = 0; y < h; ++y)
		for (int x = 0; x < w; ++x) {
			const uint8_t *p = src + ((size_t)y * w + x) * bpp;
			int v = (RY * p[ro] + GY * p[go] + BY * p[bo] + (32 << 14) + (1 << 8)) >> 9;
			v = (v * 16384) >> 13;
			if (v > 32767) v = 32767;
			v = (v + 64) >> 7;
			dy[(size_t)y * w + x] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
		}
	int16_t *u15 = (int16_t *)malloc(sizeof(int16_t) * (size_t)h * cw), *v15 = (int16_t *)malloc(sizeof(int16_t) * (size_t)h * cw);
	for (int y = 0; y < h; ++y)
		for (int x = 0; x < cw; ++x) {
			const uint8_t *p = src + ((size_t)y * w + 2 * x) * bpp;
			const int r = p[ro] + p[bpp + ro], g = p[go] + p[bpp + go], b = p[bo] + p[bpp + bo];
			int u = (RU * r + GU * g + BU * b + (256 << 15) + (1 << 9)) >> 10, v = (RV * r + GV * g + BV * b + (256 << 15) + (1 << 9)) >> 10;
			u = (u * 16384) >> 13;
			v = (v * 16384) >> 13;
			u15[(size_t)y * cw + x] = (int16_t)(u > 32767 ? 32767 : u);
			v15[(size_t)y * cw + x] = (int16_t)(v > 32767 ? 32767 : v);
		}
	static const int tap[4] = {512, 1536, 1536, 512};
	sws_filter vf = {0, NULL, NULL};
	if (x86_vertical) /* the library's own vertical chroma filter (h rows -> h / 2): border rows carry FOLDED coefficients,
	                   * which the SIMD scaler's per-tap truncation sees differently from replicated rows */
		init_filter(&vf, (int)((((int64_t)h << 16) + ((h / 2) >> 1)) / (h / 2)), h, h / 2, 2, 1 << 12, get_local_pos(0, -513),
		            get_local_pos(1, -513));
	for (int y = 0; y < h / 2; ++y)
		for (int x = 0; x < cw; ++x) {
			if (x86_vertical && y < h / 2 - 1) { /* x86/yuv2yuvX.asm, see orc_scaler_set_x86_vertical; last line: C functions */
				const int16_t *cf = vf.coef + (size_t)y * vf.size;
				int au = (64 + 8 * (vf.size - 1)) >> 4, av = au;
				for (int j = 0; j < vf.size; ++j) {
					au += (u15[(size_t)(vf.pos[y] + j) * cw + x] * cf[j]) >> 16;
					av += (v15[(size_t)(vf.pos[y] + j) * cw + x] * cf[j]) >> 16;
				}
				au = (int16_t)au >> 3;
				av = (int16_t)av >> 3;
				du[(size_t)y * cw + x] = (uint8_t)(au < 0 ? 0 : (au > 255 ? 255 : au));
				dv[(size_t)y * cw + x] = (uint8_t)(av < 0 ? 0 : (av > 255 ? 255 : av));
				continue;
			}
			int au = 64 << 12, av = 64 << 12;
			for (int j = 0; j < 4; ++j) {
				int row = 2 * y - 1 + j;
				row = row < 0 ? 0 : (row > h - 1 ? h - 1 : row);
				au += tap[j] * u15[(size_t)row * cw + x];
				av += tap[j] * v15[(size_t)row * cw + x];
			}
			au >>= 19;
			av >>= 19;
			du[(size_t)y * cw + x] = (uint8_t)(au < 0 ? 0 : (au > 255 ? 255 : au));
			dv[(size_t)y * cw + x] = (uint8_t)(av < 0 ? 0 : (av > 255 ? 255 : av));
		}
	free(u15);
	free(v15);
	free(vf.pos);
	free(vf.coef);
}

int orc_scaler_process(orc_scaler *s, const uint8_t *src, uint8_t *dst) {
	const int sw = s->src_w, sh = s->src_h, dw = s->dst_w, dh = s->dst_h;
	if (s->src_fmt == PIX_RGB565) {
		/* libswscale's 16-bit RGB readers (input.c rgb16_32ToY_c_template / rgb16_32ToUV_half_c_template, instantiated for
		 * rgb16le with masks 0xF800 / 0x07E0 / 0x001F, coefficient shifts 0 / 5 / 11 and S = RGB2YUV_SHIFT + 8) multiply the
		 * masked fields where they stand; term by term that is the RGB24 reader's arithmetic scaled by 2^8 on
		 * r = r5 << 3, g = g6 << 2, b = b5 << 3 (plain shifts, no bit replication) — so: expand, then the RGB24 path.
		 * Pinned against the live library (tests/test_oracle_video_live.py). */
		uint8_t *tmp = (uint8_t *)malloc((size_t)sw * sh * 3);
		for (size_t i = 0; i < (size_t)sw * sh; ++i) {
			const unsigned px = src[2 * i] | ((unsigned)src[2 * i + 1] << 8);
			tmp[3 * i] = (uint8_t)((px >> 11) << 3);
			tmp[3 * i + 1] = (uint8_t)(((px >> 5) & 63) << 2);
			tmp[3 * i + 2] = (uint8_t)((px & 31) << 3);
		}
		rgb_to_i420(tmp, sw, sh, 0, 3, 0, 1, 2, dst, s->x86_vertical);
		free(tmp);
		return 0;
	}
	if (s->src_fmt == PIX_RGB24 || s->src_fmt == PIX_BGR24 || s->src_fmt == PIX_RGBA32 || s->src_fmt == PIX_BGRA32) {
		const int f = s->src_fmt, four = f == PIX_RGBA32 || f == PIX_BGRA32;
		rgb_to_i420(src, sw, sh, f == PIX_BGR24, four ? 4 : 3, f == PIX_BGRA32 ? 2 : 0, 1, f == PIX_BGRA32 ? 0 : 2, dst, s->x86_vertical);
		return 0;
	}
	if (is_packed422(s->src_fmt)) {
		const int yo = s->src_fmt == PIX_UYVY ? 1 : 0, uo = s->src_fmt == PIX_UYVY ? 0 : 1, vo = uo + 2;
		uint8_t *dy = dst, *du = dst + (size_t)sw * sh, *dv = du + (size_t)(sw / 2) * (sh / 2);
		for (int y = 0; y < sh; ++y)
			for (int x = 0; x < sw; ++x)
				dy[(size_t)y * sw + x] = src[(size_t)y * sw * 2 + 2 * x + yo];
		for (int y = 0; y < sh / 2; ++y) {
			const uint8_t *r0 = src + (size_t)(2 * y) * sw * 2, *r1 = r0 + (size_t)sw * 2;
			for (int x = 0; x < sw / 2; ++x) {
				/* the library's x86 row function averages whole groups of 8 chroma samples with a rounding SIMD average and
				 * the rest of the row with a truncating scalar loop (rgb2rgb extract_odd2avg): reproduced as observed on the
				 * live library (tests/test_oracle_video_live.py) */
				const int rnd = x < ((sw / 2) & ~7) ? 1 : 0;
				du[(size_t)y * (sw / 2) + x] = (uint8_t)((r0[4 * x + uo] + r1[4 * x + uo] + rnd) >> 1);
				dv[(size_t)y * (sw / 2) + x] = (uint8_t)((r0[4 * x + vo] + r1[4 * x + vo] + rnd) >> 1);
			}
		}
		return 0;
	}
	const int csw = s->chr_src_w, csh = s->chr_src_h, cdw = s->chr_dst_w, cdh = s->chr_dst_h;
	/* +16 bytes of padding after each source row copy: the aligned filter may read (zero-weighted) past the row end */
	int16_t *lum = (int16_t *)malloc(sizeof(int16_t) * (size_t)sh * dw);
	int16_t *chu = (int16_t *)malloc(sizeof(int16_t) * (size_t)csh * cdw);
	int16_t *chv = (int16_t *)malloc(sizeof(int16_t) * (size_t)csh * cdw);
	uint8_t *row = (uint8_t *)calloc((size_t)sw + 32, 1), *ru = (uint8_t *)calloc((size_t)csw + 32, 1),
	        *rv = (uint8_t *)calloc((size_t)csw + 32, 1);
	const uint8_t *sy = src, *sc = src + (size_t)sw * sh;
	for (int y = 0; y < sh; ++y) {
		memcpy(row, sy + (size_t)y * sw, (size_t)sw);
		hscale(lum + (size_t)y * dw, dw, row, &s->hLum);
	}
	for (int y = 0; y < csh; ++y) {
		if (s->src_fmt == PIX_YUV420P) {
			memcpy(ru, sc + (size_t)y * csw, (size_t)csw);
			memcpy(rv, sc + (size_t)csw * csh + (size_t)y * csw, (size_t)csw);
		} else { /* nvXXtoUV_c */
			const uint8_t *p = sc + (size_t)y * csw * 2;
			for (int x = 0; x < csw; ++x) {
				ru[x] = p[2 * x + (s->src_fmt == PIX_NV21)];
				rv[x] = p[2 * x + (s->src_fmt == PIX_NV12)];
			}
		}
		hscale(chu + (size_t)y * cdw, cdw, ru, &s->hChr);
		hscale(chv + (size_t)y * cdw, cdw, rv, &s->hChr);
	}
	if (s->dst_fmt == PIX_YUV420P) {
		/* yuv2planeX_8_c / yuv2plane1_8_c with the constant 64 "dither" */
		uint8_t *dy = dst, *du = dst + (size_t)dw * dh, *dv = du + (size_t)cdw * cdh;
		for (int pl = 0; pl < 3; ++pl) {
			const sws_filter *vf = pl == 0 ? &s->vLum : &s->vChr;
			const int16_t *plane = pl == 0 ? lum : pl == 1 ? chu : chv;
			uint8_t *out = pl == 0 ? dy : pl == 1 ? du : dv;
			const int W = pl == 0 ? dw : cdw, H = pl == 0 ? dh : cdh;
			for (int y = 0; y < H; ++y) {
				const int16_t *cf = vf->coef + (size_t)y * vf->size;
				for (int x = 0; x < W; ++x) {
					int val;
					if (vf->size == 1) {
						val = (plane[(size_t)vf->pos[y] * W + x] + 64) >> 7;
					} else if (s->x86_vertical && y < (pl == 0 ? dh - 2 : cdh - 1)) { /* the last two output lines (and the chroma line that goes with them) are done by the C functions (swscale.c: "can't use MMX here without overwriting this array's tail") */
						/* the library's x86 vertical scaler for planar output without SWS_ACCURATE_RND / SWS_BITEXACT
						 * (x86/yuv2yuvX.asm): 16-bit accumulators, one pmulhw per tap (the fraction of every product is
						 * dropped), a rounder that pays the expected loss back: ((64 + 8 (taps - 1)) >> 4), final >> 3 */
						int acc = (64 + 8 * (vf->size - 1)) >> 4;
						for (int j = 0; j < vf->size; ++j)
							acc += (plane[(size_t)(vf->pos[y] + j) * W + x] * cf[j]) >> 16;
						val = (int16_t)acc >> 3;
					} else {
						val = 64 << 12;
						for (int j = 0; j < vf->size; ++j)
							val += plane[(size_t)(vf->pos[y] + j) * W + x] * cf[j];
						val >>= 19;
					}
					out[(size_t)y * W + x] = (uint8_t)clip_u8(val);
				}
			}
		}
	} else {
		const int bgr = s->dst_fmt == PIX_BGR24;
		const sws_filter *vl = &s->vLum, *vc = &s->vChr;
		for (int y = 0; y < dh; ++y) {
			const int16_t *lf = vl->coef + (size_t)y * vl->size, *cf = vc->coef + (size_t)y * vc->size;
			uint8_t *o = dst + (size_t)y * dw * 3;
			for (int i = 0; i < (dw + 1) >> 1; ++i) {
				int Y1, Y2, U, V;
				if (vl->size == 1 && vc->size <= 2) {
					/* yuv2rgb_1_c_template (unscaled vertically) */
					const int16_t *b0 = lum + (size_t)vl->pos[y] * dw;
					Y1 = (b0[i * 2] + 64) >> 7;
					Y2 = (i * 2 + 1 < dw ? b0[i * 2 + 1] + 64 : 64) >> 7;
					const int16_t *u0 = chu + (size_t)vc->pos[y] * cdw, *v0 = chv + (size_t)vc->pos[y] * cdw;
					const int uvalpha = vc->size == 1 ? 0 : cf[1];
					if (uvalpha == 0) {
						U = (u0[i] + 64) >> 7;
						V = (v0[i] + 64) >> 7;
					} else { /* libswscale >= 8: true interpolation between the two chroma lines, rounded */
						const int16_t *u1 = u0 + cdw, *v1 = v0 + cdw;
						const int uvalpha1 = 4096 - uvalpha;
						U = (u0[i] * uvalpha1 + u1[i] * uvalpha + (128 << 11)) >> 19;
						V = (v0[i] * uvalpha1 + v1[i] * uvalpha + (128 << 11)) >> 19;
					}
				} else if (vl->size == 2 && vc->size == 2) {
					/* yuv2rgb_2_c_template (bilinear upscale) */
					const int16_t *b0 = lum + (size_t)vl->pos[y] * dw, *b1 = b0 + dw;
					const int16_t *u0 = chu + (size_t)vc->pos[y] * cdw, *u1 = u0 + cdw;
					const int16_t *v0 = chv + (size_t)vc->pos[y] * cdw, *v1 = v0 + cdw;
					const int yalpha = lf[1], uvalpha = cf[1];
					const int yalpha1 = 4096 - yalpha, uvalpha1 = 4096 - uvalpha;
					Y1 = (b0[i * 2] * yalpha1 + b1[i * 2] * yalpha) >> 19;
					Y2 = i * 2 + 1 < dw ? (b0[i * 2 + 1] * yalpha1 + b1[i * 2 + 1] * yalpha) >> 19 : 0;
					U = (u0[i] * uvalpha1 + u1[i] * uvalpha) >> 19;
					V = (v0[i] * uvalpha1 + v1[i] * uvalpha) >> 19;
				} else {
					/* yuv2rgb_X_c_template */
					Y1 = Y2 = U = V = 1 << 18;
					for (int j = 0; j < vl->size; ++j) {
						const int16_t *r = lum + (size_t)(vl->pos[y] + j) * dw;
						Y1 += r[i * 2] * lf[j];
						if (i * 2 + 1 < dw) Y2 += r[i * 2 + 1] * lf[j];
					}
					for (int j = 0; j < vc->size; ++j) {
						U += chu[(size_t)(vc->pos[y] + j) * cdw + i] * cf[j];
						V += chv[(size_t)(vc->pos[y] + j) * cdw + i] * cf[j];
					}
					Y1 >>= 19;
					Y2 >>= 19;
					U >>= 19;
					V >>= 19;
				}
				uint8_t r, g, b;
				yuv_to_rgb_px(s, Y1, U, V, &r, &g, &b);
				o[i * 6 + 0] = bgr ? b : r;
				o[i * 6 + 1] = g;
				o[i * 6 + 2] = bgr ? r : b;
				if (i * 2 + 1 < dw) {
					yuv_to_rgb_px(s, Y2, U, V, &r, &g, &b);
					o[i * 6 + 3] = bgr ? b : r;
					o[i * 6 + 4] = g;
					o[i * 6 + 5] = bgr ? r : b;
				}
			}
		}
	}
	free(lum);
	free(chu);
	free(chv);
	free(row);
	free(ru);
	free(rv);
	return 0;
}
