// video_copy.cu — ms_yuv_buf_copy_with_pix_strides() (/root/reference/src/voip/msvideo.c:245-270 over plane_copy :204-227
// and row_copy :188-202) batched over frames: copy a region of interest of a three-plane YUV picture into a region of
// another picture, each plane with its own row stride AND pixel stride — which is how the reference converts between
// planar I420 and semi-planar NV12 / NV21 (pixel stride 2 on the chroma planes) and slides a picture inside a larger
// buffer (a compositor's tile). Integer byte moves: bit-exact by construction; pinned by the reference's own eight
// patterns (tester/mediastreamer2_framework_tester.c:393-500) in tests/test_gpu_video_copy.py.
//
// Semantics kept from the reference: the LUMA plane copies src_roi.w x src_roi.h elements from (src_roi.x, src_roi.y) to
// (dst_roi.x, dst_roi.y); for the two chroma planes every field of both rectangles is halved (integer division) first.
// dst_roi.w / dst_roi.h only bound the copy (row_copy stops at whichever end comes first, :196).
//
// Kernels (HBM-bound byte moves):
//   yuv_copy_rows_kernel    both pixel strides 1: 16-byte vectors when source and destination rows are co-aligned
//   yuv_copy_chroma_kernel  the two chroma planes of an interleaved side handled together: one 16-bit access per (U, V)
//                           pair on the semi-planar side instead of two byte accesses with stride 2
//   yuv_copy_generic_kernel anything else: one element per thread, consecutive threads along a row
#include "msb200_internal.h"

struct CopyPlane {
	const unsigned char *src;
	unsigned char *dst;
	long src_frame, dst_frame; // bytes between consecutive frames
	int src_row, dst_row, src_pix, dst_pix;
	int w, h;
};

__global__ void __launch_bounds__(256) yuv_copy_generic_kernel(const CopyPlane P, int n_frames) {
	const long per_frame = (long)P.w * P.h, total = per_frame * n_frames;
	for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
		const int f = (int)(i / per_frame);
		const int r = (int)(i - (long)f * per_frame);
		const int y = r / P.w, x = r - y * P.w;
		P.dst[(long)f * P.dst_frame + (long)y * P.dst_row + (long)x * P.dst_pix] =
		    P.src[(long)f * P.src_frame + (long)y * P.src_row + (long)x * P.src_pix];
	}
}

// rows of w contiguous bytes on both sides. Each row is cut into: a head up to the first 16-byte boundary of the
// DESTINATION, 16-byte vectors, a tail. The vector path needs source and destination to share their alignment mod 16
// (decided per launch on the host); otherwise everything goes through the byte path.
__global__ void __launch_bounds__(256) yuv_copy_rows_kernel(const CopyPlane P, int n_frames, int vec_ok) {
	const int vec_per_row = (P.w + 15) / 16 + 1; // work items per row (head, vectors, tail folded into "chunks of <= 16 bytes")
	const long per_frame = (long)vec_per_row * P.h, total = per_frame * n_frames;
	for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
		const int f = (int)(i / per_frame);
		const int r = (int)(i - (long)f * per_frame);
		const int y = r / vec_per_row, c = r - y * vec_per_row;
		const unsigned char *s = P.src + (long)f * P.src_frame + (long)y * P.src_row;
		unsigned char *d = P.dst + (long)f * P.dst_frame + (long)y * P.dst_row;
		const int head = (int)((16 - ((uintptr_t)d & 15)) & 15); // bytes before the destination's first 16-byte boundary
		int b0, b1;
		if (c == 0) {
			b0 = 0;
			b1 = min(head, P.w);
		} else {
			b0 = head + (c - 1) * 16;
			b1 = min(b0 + 16, P.w);
		}
		if (b0 >= b1) continue;
		if (vec_ok && b1 - b0 == 16) {
			*reinterpret_cast<int4 *>(d + b0) = *reinterpret_cast<const int4 *>(s + b0);
		} else {
			for (int b = b0; b < b1; ++b) d[b] = s[b];
		}
	}
}

// both chroma planes at once. semi_src / semi_dst: that side stores (first, second) byte pairs at pixel stride 2 with
// plane[2] == plane[1] + 1 (NV12: U first) or plane[1] == plane[2] + 1 (NV21); `a` is the plane at the LOWER address.
struct CopyChroma {
	const unsigned char *src_a, *src_b; // source of the value that goes to dst_a / dst_b
	unsigned char *dst_a, *dst_b;
	long src_frame, dst_frame;
	int src_row, dst_row;
	int w, h;
	int semi_src, semi_dst;
};
__global__ void __launch_bounds__(256) yuv_copy_chroma_kernel(const CopyChroma P, int n_frames) {
	const long per_frame = (long)P.w * P.h, total = per_frame * n_frames;
	for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
		const int f = (int)(i / per_frame);
		const int r = (int)(i - (long)f * per_frame);
		const int y = r / P.w, x = r - y * P.w;
		const long so = (long)f * P.src_frame + (long)y * P.src_row, dof = (long)f * P.dst_frame + (long)y * P.dst_row;
		unsigned char a, b;
		if (P.semi_src) {
			const unsigned short v = *reinterpret_cast<const unsigned short *>(P.src_a + so + 2 * x); // src_a is 2-byte aligned (host check)
			a = (unsigned char)(v & 0xff);
			b = (unsigned char)(v >> 8);
		} else {
			a = P.src_a[so + x];
			b = P.src_b[so + x];
		}
		if (P.semi_dst) {
			*reinterpret_cast<unsigned short *>(P.dst_a + dof + 2 * x) = (unsigned short)(a | (b << 8));
		} else {
			P.dst_a[dof + x] = a;
			P.dst_b[dof + x] = b;
		}
	}
}

// eight (a, b) pairs per thread: 8 + 8 planar bytes <-> 16 interleaved bytes, every access one aligned vector. The host
// takes this path when exactly one side is interleaved, w % 8 == 0 and all pointers / strides keep the alignment.
__global__ void __launch_bounds__(256) yuv_copy_chroma_vec_kernel(const CopyChroma P, int n_frames) {
	const int wv = P.w >> 3;
	const long per_frame = (long)wv * P.h, total = per_frame * n_frames;
	for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
		const int f = (int)(i / per_frame);
		const int r = (int)(i - (long)f * per_frame);
		const int y = r / wv, x = (r - y * wv) << 3;
		const long so = (long)f * P.src_frame + (long)y * P.src_row, dof = (long)f * P.dst_frame + (long)y * P.dst_row;
		if (P.semi_src) { // 16 interleaved bytes -> 8 + 8
			const uint4 v = *reinterpret_cast<const uint4 *>(P.src_a + so + 2 * x);
			uint2 a, b;
			a.x = __byte_perm(v.x, v.y, 0x6420); b.x = __byte_perm(v.x, v.y, 0x7531);
			a.y = __byte_perm(v.z, v.w, 0x6420); b.y = __byte_perm(v.z, v.w, 0x7531);
			*reinterpret_cast<uint2 *>(P.dst_a + dof + x) = a;
			*reinterpret_cast<uint2 *>(P.dst_b + dof + x) = b;
		} else { // 8 + 8 -> 16 interleaved bytes
			const uint2 a = *reinterpret_cast<const uint2 *>(P.src_a + so + x), b = *reinterpret_cast<const uint2 *>(P.src_b + so + x);
			uint4 v;
			v.x = __byte_perm(a.x, b.x, 0x5140); v.y = __byte_perm(a.x, b.x, 0x7362);
			v.z = __byte_perm(a.y, b.y, 0x5140); v.w = __byte_perm(a.y, b.y, 0x7362);
			*reinterpret_cast<uint4 *>(P.dst_a + dof + 2 * x) = v;
		}
	}
}

static int grid_for(msb200_ctx *ctx, long items) {
	long g = (items + 255) / 256;
	const long cap = (long)ctx->sm_count * 8;
	return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

static int copy_launch(msb200_ctx *ctx, int n_frames, const unsigned char *src, const msb200_yuv_layout *sl, msb200_rect sr,
                       unsigned char *dst, const msb200_yuv_layout *dl, msb200_rect dr) {
	CopyPlane pl[3];
	for (int p = 0; p < 3; ++p) {
		if (p == 1) { // msvideo.c:256-264
			sr.x /= 2; sr.y /= 2; sr.w /= 2; sr.h /= 2;
			dr.x /= 2; dr.y /= 2; dr.w /= 2; dr.h /= 2;
		}
		CopyPlane &c = pl[p];
		c.src_row = sl->row_stride[p]; c.src_pix = sl->pix_stride[p];
		c.dst_row = dl->row_stride[p]; c.dst_pix = dl->pix_stride[p];
		c.src = src + sl->plane_offset[p] + (long)sr.y * c.src_row + (long)sr.x * c.src_pix;
		c.dst = dst + dl->plane_offset[p] + (long)dr.y * c.dst_row + (long)dr.x * c.dst_pix;
		c.src_frame = (long)sl->frame_bytes;
		c.dst_frame = (long)dl->frame_bytes;
		c.w = sr.w < dr.w ? sr.w : dr.w; // row_copy stops at whichever end comes first (:196)
		c.h = sr.h;
		// plane_copy's shortcut (:216-218): equal row strides, unit pixel strides and equal rectangles are ONE memcpy of
		// row_stride x h bytes from the region's first byte — the bytes between the rows' ends and the next rows' starts
		// travel too. Kept (when it stays inside both frames) so that results equal the reference's byte for byte.
		if (c.src_row == c.dst_row && c.src_pix == 1 && c.dst_pix == 1 && sr.x == dr.x && sr.y == dr.y && sr.w == dr.w && sr.h == dr.h &&
		    dr.h > 0) {
			const long bytes = (long)c.dst_row * dr.h;
			const long s_end = (long)sl->plane_offset[p] + (long)sr.y * c.src_row + sr.x + bytes;
			const long d_end = (long)dl->plane_offset[p] + (long)dr.y * c.dst_row + dr.x + bytes;
			if (s_end <= (long)sl->frame_bytes && d_end <= (long)dl->frame_bytes && bytes < (1L << 30)) {
				c.w = (int)bytes;
				c.h = 1;
			}
		}
	}
	for (int p = 0; p < 3; ++p) {
		CopyPlane &c = pl[p];
		if (c.w <= 0 || c.h <= 0) continue;
		if (p == 1) {
			// the two chroma planes together when a side interleaves them
			const CopyPlane &u = pl[1], &v = pl[2];
			const bool s_semi = u.src_pix == 2 && v.src_pix == 2 && u.src_row == v.src_row && (v.src == u.src + 1 || u.src == v.src + 1);
			const bool d_semi = u.dst_pix == 2 && v.dst_pix == 2 && u.dst_row == v.dst_row && (v.dst == u.dst + 1 || u.dst == v.dst + 1);
			const bool s_plan = u.src_pix == 1 && v.src_pix == 1 && u.src_row == v.src_row;
			const bool d_plan = u.dst_pix == 1 && v.dst_pix == 1 && u.dst_row == v.dst_row;
			if ((s_semi || d_semi) && (s_semi || s_plan) && (d_semi || d_plan) && u.w == v.w && u.h == v.h) {
				// `a` = the value stored at the lower address of whichever side is interleaved
				const bool u_low_src = !s_semi || v.src == u.src + 1, u_low_dst = !d_semi || v.dst == u.dst + 1;
				bool ok = true;
				CopyChroma q;
				q.src_frame = u.src_frame; q.dst_frame = u.dst_frame; q.src_row = u.src_row; q.dst_row = u.dst_row;
				q.w = u.w; q.h = u.h; q.semi_src = s_semi; q.semi_dst = d_semi;
				if (s_semi && d_semi) {
					if (u_low_src != u_low_dst) ok = false; // NV12 <-> NV21 swap: leave it to the generic kernel
					q.src_a = u_low_src ? u.src : v.src; q.src_b = nullptr;
					q.dst_a = u_low_dst ? u.dst : v.dst; q.dst_b = nullptr;
				} else if (s_semi) { // pairs (a, b) in the source; a is U when U is low
					q.src_a = u_low_src ? u.src : v.src; q.src_b = nullptr;
					q.dst_a = u_low_src ? u.dst : v.dst; q.dst_b = u_low_src ? v.dst : u.dst;
				} else {
					q.dst_a = u_low_dst ? u.dst : v.dst; q.dst_b = nullptr;
					q.src_a = u_low_dst ? u.src : v.src; q.src_b = u_low_dst ? v.src : u.src;
				}
				if (s_semi && (((uintptr_t)q.src_a | (uintptr_t)q.src_row | (uintptr_t)q.src_frame) & 1)) ok = false;
				if (d_semi && (((uintptr_t)q.dst_a | (uintptr_t)q.dst_row | (uintptr_t)q.dst_frame) & 1)) ok = false;
				if (ok) {
					auto al = [](const void *p, long a, long b, int m) { return (((uintptr_t)p | (uintptr_t)a | (uintptr_t)b) & (uintptr_t)(m - 1)) == 0; };
					const bool vec = s_semi != d_semi && q.w % 8 == 0 &&
					                 (s_semi ? al(q.src_a, q.src_row, q.src_frame, 16) && al(q.dst_a, q.dst_row, q.dst_frame, 8) &&
					                               al(q.dst_b, 0, 0, 8)
					                         : al(q.dst_a, q.dst_row, q.dst_frame, 16) && al(q.src_a, q.src_row, q.src_frame, 8) &&
					                               al(q.src_b, 0, 0, 8));
					if (vec) MSB200_LAUNCH(ctx, yuv_copy_chroma_vec_kernel, grid_for(ctx, (long)(q.w / 8) * q.h * n_frames), 256, 0, q, n_frames);
					else MSB200_LAUNCH(ctx, yuv_copy_chroma_kernel, grid_for(ctx, (long)q.w * q.h * n_frames), 256, 0, q, n_frames);
					break; // planes 1 and 2 done
				}
			}
		}
		if (c.src_pix == 1 && c.dst_pix == 1) {
			const int vec_ok = ((((uintptr_t)c.src ^ (uintptr_t)c.dst) | (uintptr_t)(c.src_row ^ c.dst_row) |
			                     (uintptr_t)(c.src_frame ^ c.dst_frame)) & 15) == 0 &&
			                   (c.dst_row & 15) == 0 && (c.dst_frame & 15) == 0;
			const long items = (long)((c.w + 15) / 16 + 1) * c.h * n_frames;
			MSB200_LAUNCH(ctx, yuv_copy_rows_kernel, grid_for(ctx, items), 256, 0, c, n_frames, vec_ok);
		} else {
			MSB200_LAUNCH(ctx, yuv_copy_generic_kernel, grid_for(ctx, (long)c.w * c.h * n_frames), 256, 0, c, n_frames);
		}
	}
	return MSB200_OK;
}

static bool layout_ok(const msb200_yuv_layout *l, msb200_rect r) {
	if (!l || r.x < 0 || r.y < 0 || r.w < 0 || r.h < 0) return false;
	for (int p = 0; p < 3; ++p) {
		const int d = p ? 2 : 1;
		if (l->row_stride[p] <= 0 || l->pix_stride[p] <= 0) return false;
		// the last element touched stays inside the frame
		const long last = (long)l->plane_offset[p] + (long)(r.y / d + (r.h / d > 0 ? r.h / d - 1 : 0)) * l->row_stride[p] +
		                  (long)(r.x / d + (r.w / d > 0 ? r.w / d - 1 : 0)) * l->pix_stride[p];
		if (last >= (long)l->frame_bytes) return false;
	}
	return true;
}

extern "C" {

int msb200_yuv_copy_strided_dev(msb200_ctx *ctx, int n_frames, const void *d_src, const msb200_yuv_layout *src_layout,
                                msb200_rect src_roi, void *d_dst, const msb200_yuv_layout *dst_layout, msb200_rect dst_roi) {
	MSB200_CHECK_ARG(ctx && d_src && d_dst && n_frames > 0);
	MSB200_CHECK_ARG(layout_ok(src_layout, src_roi) && layout_ok(dst_layout, dst_roi));
	return copy_launch(ctx, n_frames, (const unsigned char *)d_src, src_layout, src_roi, (unsigned char *)d_dst, dst_layout, dst_roi);
}

int msb200_yuv_copy_strided(msb200_ctx *ctx, int n_frames, const uint8_t *src, const msb200_yuv_layout *src_layout,
                            msb200_rect src_roi, uint8_t *dst, const msb200_yuv_layout *dst_layout, msb200_rect dst_roi) {
	MSB200_CHECK_ARG(ctx && src && dst && n_frames > 0);
	MSB200_CHECK_ARG(layout_ok(src_layout, src_roi) && layout_ok(dst_layout, dst_roi));
	void *ds = nullptr, *dd = nullptr;
	const size_t sb = src_layout->frame_bytes * (size_t)n_frames, db = dst_layout->frame_bytes * (size_t)n_frames;
	MSB200_CUDA(cudaSetDevice(ctx->device));
	MSB200_CUDA(cudaMalloc(&ds, sb));
	cudaError_t e = cudaMalloc(&dd, db);
	if (e != cudaSuccess) {
		cudaFree(ds);
		msb200_set_error("cudaMalloc(%zu) -> %s", db, cudaGetErrorString(e));
		return MSB200_ENOMEM;
	}
	int r = MSB200_OK;
	cudaStream_t st = ctx->stream;
	// only the region is written: the destination's other bytes go up and come back unchanged, as on the CPU
	if ((e = cudaMemcpyAsync(ds, src, sb, cudaMemcpyHostToDevice, st)) != cudaSuccess ||
	    (e = cudaMemcpyAsync(dd, dst, db, cudaMemcpyHostToDevice, st)) != cudaSuccess) {
		msb200_set_error("yuv_copy_strided: %s", cudaGetErrorString(e));
		r = MSB200_ECUDA;
	}
	if (r == MSB200_OK)
		r = copy_launch(ctx, n_frames, (const unsigned char *)ds, src_layout, src_roi, (unsigned char *)dd, dst_layout, dst_roi);
	if (r == MSB200_OK && (e = cudaMemcpyAsync(dst, dd, db, cudaMemcpyDeviceToHost, st)) != cudaSuccess) {
		msb200_set_error("yuv_copy_strided: %s", cudaGetErrorString(e));
		r = MSB200_ECUDA;
	}
	cudaStreamSynchronize(st);
	cudaFree(ds);
	cudaFree(dd);
	return r;
}

} // extern "C"
