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
