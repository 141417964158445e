// video_nv12.cu — NV12/NV21 -> I420 (+ rotation 0/90/180/270, + nearest 1/2 decimation), batched over frames.
//
// Bit-exact replacement of copy_ycbcrbiplanar_to_true_yuv_with_rotation_and_down_scale_by_2()
// /root/reference/src/voip/msvideo.c:787-919 (rotate_plane_down_scale_by_2 :734-776). Pure byte movement: the
// kernel is destination-centric (each thread produces 4 consecutive output bytes so stores are coalesced 32-bit
// words); the rotation-0 full-size path moves 16 bytes per thread when strides and bases allow it.
#include "msb200_internal.h"

struct Nv12Params {
	int rotation, w, h, ys, cs, u_first, f; // (w,h) = destination size; f = decimation factor (1 or 2)
	size_t src_frame_bytes, cbcr_offset, dst_frame_bytes;
};

// source offset of destination luma pixel (i,j) — derived from msvideo.c:832-857 (0), :866-872 (180), :734-776 (90/270)
__device__ __forceinline__ size_t nv12_src_y(const Nv12Params &p, int i, int j) {
	switch (p.rotation) {
		case 0: return (size_t)(i * p.f) * p.ys + (size_t)j * p.f;
		case 180: return (size_t)(p.h - 1 - i) * p.ys * p.f + (size_t)(p.w - 1 - j) * p.f;
		case 90: return (size_t)(p.w - 1 - j) * p.f * p.ys + (size_t)i * p.f;          // clockwise
		default: return (size_t)j * p.f * p.ys + (size_t)(p.h - 1 - i) * p.f;          // 270, anticlockwise
	}
}
// source offset (of the Cb byte) for destination chroma pixel (i,j)
__device__ __forceinline__ size_t nv12_src_c(const Nv12Params &p, int i, int j) {
	const int uw = p.w / 2, uh = p.h / 2;
	switch (p.rotation) {
		case 0: return (size_t)p.cs * i * p.f + (size_t)2 * j * p.f;
		case 180: return (size_t)p.cs * (uh - 1 - i) * p.f + (size_t)2 * (uw - 1 - j) * p.f;
		case 90: return (size_t)(uw - 1 - j) * ((size_t)(p.cs / 2) * 2 * p.f) + (size_t)i * 2 * p.f;
		default: return (size_t)j * ((size_t)(p.cs / 2) * 2 * p.f) + (size_t)(uh - 1 - i) * 2 * p.f;
	}
}

// generic path: grid.y = frame, one thread per 4 destination bytes of the frame (luma quads, then U quads, then V quads)
__global__ void __launch_bounds__(256) nv12_generic_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, Nv12Params p) {
	const int uw = p.w / 2, uh = p.h / 2;
	const long yq = ((long)p.w * p.h + 3) / 4, cq = ((long)uw * uh + 3) / 4;
	const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= yq + 2 * cq) return;
	const uint8_t *fs = src + (size_t)blockIdx.y * p.src_frame_bytes;
	uint8_t *fd = dst + (size_t)blockIdx.y * p.dst_frame_bytes;
	uint8_t v[4];
	if (q < yq) {
		const long base = q * 4, total = (long)p.w * p.h;
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			long idx = base + k;
			if (idx < total) v[k] = fs[nv12_src_y(p, (int)(idx / p.w), (int)(idx % p.w))];
		}
		uint8_t *o = fd + base;
		if (base + 3 < total && (((uintptr_t)o) & 3) == 0) *reinterpret_cast<uchar4 *>(o) = make_uchar4(v[0], v[1], v[2], v[3]);
		else
			for (int k = 0; k < 4 && base + k < total; ++k) o[k] = v[k];
		return;
	}
	const bool is_v = (q - yq) >= cq;
	const long base = ((q - yq) - (is_v ? cq : 0)) * 4, total = (long)uw * uh;
	const uint8_t *cb = fs + p.cbcr_offset + (is_v ? 1 : 0);
	// uFirstvSecond == FALSE swaps the destination planes (:822-826)
	const int plane = (is_v ? 1 : 0) ^ (p.u_first ? 0 : 1);
	uint8_t *o = fd + (size_t)p.w * p.h + (size_t)plane * total + base;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		long idx = base + k;
		if (idx < total) v[k] = cb[nv12_src_c(p, (int)(idx / uw), (int)(idx % uw))];
	}
	if (base + 3 < total && (((uintptr_t)o) & 3) == 0) *reinterpret_cast<uchar4 *>(o) = make_uchar4(v[0], v[1], v[2], v[3]);
	else
		for (int k = 0; k < 4 && base + k < total; ++k) o[k] = v[k];
}

// fast path (rotation 0, no decimation, w % 32 == 0, 16-byte aligned rows): one thread copies 16 luma bytes, or
// de-interleaves 32 CbCr bytes into 16 U + 16 V.
__global__ void __launch_bounds__(256) nv12_fast_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, Nv12Params p) {
	const int uw = p.w / 2, uh = p.h / 2;
	const int yv = p.w / 16, cv = uw / 16;
	const long ny = (long)yv * p.h, nc = (long)cv * uh;
	const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= ny + nc) return;
	const uint8_t *fs = src + (size_t)blockIdx.y * p.src_frame_bytes;
	uint8_t *fd = dst + (size_t)blockIdx.y * p.dst_frame_bytes;
	if (t < ny) {
		const int row = (int)(t / yv), col = (int)(t % yv);
		const int4 v = __ldg(reinterpret_cast<const int4 *>(fs + (size_t)row * p.ys) + col);
		reinterpret_cast<int4 *>(fd + (size_t)row * p.w)[col] = v;
		return;
	}
	const long c = t - ny;
	const int row = (int)(c / cv), col = (int)(c % cv);
	const int4 *s = reinterpret_cast<const int4 *>(fs + p.cbcr_offset + (size_t)row * p.cs) + 2 * col;
	const int4 a = __ldg(s), b = __ldg(s + 1);
	// bytes: Cb0 Cr0 Cb1 Cr1 ... ; __byte_perm picks even (0x6420-style) / odd bytes of a register pair
	int4 u, v;
	u.x = __byte_perm(a.x, a.y, 0x6420); v.x = __byte_perm(a.x, a.y, 0x7531);
	u.y = __byte_perm(a.z, a.w, 0x6420); v.y = __byte_perm(a.z, a.w, 0x7531);
	u.z = __byte_perm(b.x, b.y, 0x6420); v.z = __byte_perm(b.x, b.y, 0x7531);
	u.w = __byte_perm(b.z, b.w, 0x6420); v.w = __byte_perm(b.z, b.w, 0x7531);
	uint8_t *pu = fd + (size_t)p.w * p.h + (p.u_first ? 0 : (size_t)uw * uh);
	uint8_t *pv = fd + (size_t)p.w * p.h + (p.u_first ? (size_t)uw * uh : 0);
	reinterpret_cast<int4 *>(pu + (size_t)row * uw)[col] = u;
	reinterpret_cast<int4 *>(pv + (size_t)row * uw)[col] = v;
}

// ------------------------------------------------------------------------------------------------ tiled rotation
// Rotation by 90 / 270 degrees at full size (msvideo.c:734-776 with down-scale off) is a byte-matrix transpose plus a
// flip. The destination-centric generic kernel reads one byte per source ROW per thread (32 sectors per warp load);
// here a CTA moves a 64 x 64 tile through shared memory instead: phase 1 reads the source rows the tile needs with
// aligned 32-bit (luma) / 64-bit (CbCr pairs, de-interleaved on the way) loads, flipped as the rotation asks, into
// T[b][a] = dst(i0 + a, j0 + b); phase 2 gives every thread one 4 x 4 byte block, transposes it in registers (8 PRMT) and
// writes four coalesced 32-bit words, one per destination row. Both directions of the copy are sector-coalesced.
//   rotation  90: dst(i, j) = src(row = W - 1 - j, col = i)      rotation 270: dst(i, j) = src(row = j, col = H - 1 - i)
// with (W, H) the destination plane size; chroma planes likewise on (W/2, H/2) with 2-byte source pixels.
#define ROT_T 64
#define ROT_PITCH 17 // words per tile row: 64 bytes + 4 of padding
__device__ __forceinline__ unsigned rot_prmt(unsigned a, unsigned b, unsigned sel) {
	unsigned d;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
	return d;
}
// phase 2: tile words -> destination plane (tight rows of `W` bytes)
__device__ __forceinline__ void rot_store_tile(const unsigned *T, uint8_t *plane, int W, int H, int i0, int j0) {
	const int t = threadIdx.x, b4 = t & 15, a4 = t >> 4; // 16 x 16 blocks of 4 x 4 bytes
	const unsigned r0 = T[(4 * b4 + 0) * ROT_PITCH + a4], r1 = T[(4 * b4 + 1) * ROT_PITCH + a4];
	const unsigned r2 = T[(4 * b4 + 2) * ROT_PITCH + a4], r3 = T[(4 * b4 + 3) * ROT_PITCH + a4];
	const unsigned t0 = rot_prmt(r0, r1, 0x5140), t1 = rot_prmt(r2, r3, 0x5140);
	const unsigned t2 = rot_prmt(r0, r1, 0x7362), t3 = rot_prmt(r2, r3, 0x7362);
	const unsigned o[4] = {rot_prmt(t0, t1, 0x5410), rot_prmt(t0, t1, 0x7632), rot_prmt(t2, t3, 0x5410), rot_prmt(t2, t3, 0x7632)};
	const int j = j0 + 4 * b4;
	if (j < W) { // W % 4 == 0: a block is inside or outside as a whole
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			const int i = i0 + 4 * a4 + k;
			if (i < H) *reinterpret_cast<unsigned *>(plane + (size_t)i * W + j) = o[k];
		}
	}
}
template <int ROT>
__global__ void __launch_bounds__(256) nv12_rot_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, Nv12Params p) {
	__shared__ unsigned TU[ROT_T * ROT_PITCH], TV[ROT_T * ROT_PITCH];
	const int W = p.w, H = p.h, uw = W / 2, uh = H / 2;
	const int ltx = (W + ROT_T - 1) / ROT_T, lty = (H + ROT_T - 1) / ROT_T, n_lt = ltx * lty;
	const int ctx_ = (uw + ROT_T - 1) / ROT_T;
	const uint8_t *fs = src + (size_t)blockIdx.y * p.src_frame_bytes;
	uint8_t *fd = dst + (size_t)blockIdx.y * p.dst_frame_bytes;
	const int t = threadIdx.x;
	if ((int)blockIdx.x < n_lt) { // ---- luma tile
		const int i0 = ((int)blockIdx.x / ltx) * ROT_T, j0 = ((int)blockIdx.x % ltx) * ROT_T;
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			const int idx = t + 256 * q, b = idx >> 4, a4 = idx & 15; // tile row b (dst column j0 + b), word a4 (dst rows i0 + 4a4 ..)
			const int j = j0 + b, i = i0 + 4 * a4;
			unsigned v = 0;
			if (j < W && i < H) {
				if (ROT == 90) v = __ldg(reinterpret_cast<const unsigned *>(fs + (size_t)(W - 1 - j) * p.ys + i));
				else v = rot_prmt(__ldg(reinterpret_cast<const unsigned *>(fs + (size_t)j * p.ys + (H - 4 - i))), 0u, 0x0123);
			}
			TU[b * ROT_PITCH + a4] = v;
		}
		__syncthreads();
		rot_store_tile(TU, fd, W, H, i0, j0);
		return;
	}
	// ---- chroma tile: the same walk over (uw, uh) with CbCr pairs as source pixels, two destination planes
	const int c = (int)blockIdx.x - n_lt;
	const int i0 = (c / ctx_) * ROT_T, j0 = (c % ctx_) * ROT_T;
	const uint8_t *cb = fs + p.cbcr_offset;
	const size_t cpitch = (size_t)(p.cs / 2) * 2;
#pragma unroll
	for (int q = 0; q < 4; ++q) {
		const int idx = t + 256 * q, b = idx >> 4, a4 = idx & 15;
		const int j = j0 + b, i = i0 + 4 * a4;
		unsigned u = 0, v = 0;
		if (j < uw && i < uh) {
			uint2 w2;
			if (ROT == 90) w2 = __ldg(reinterpret_cast<const uint2 *>(cb + (size_t)(uw - 1 - j) * cpitch + (size_t)i * 2));
			else w2 = __ldg(reinterpret_cast<const uint2 *>(cb + (size_t)j * cpitch + (size_t)(uh - 4 - i) * 2));
			u = rot_prmt(w2.x, w2.y, 0x6420); // Cb of the four pixels, in address order
			v = rot_prmt(w2.x, w2.y, 0x7531);
			if (ROT != 90) {
				u = rot_prmt(u, 0u, 0x0123);
				v = rot_prmt(v, 0u, 0x0123);
			}
		}
		TU[b * ROT_PITCH + a4] = u;
		TV[b * ROT_PITCH + a4] = v;
	}
	__syncthreads();
	uint8_t *pu = fd + (size_t)W * H + (p.u_first ? 0 : (size_t)uw * uh);
	uint8_t *pv = fd + (size_t)W * H + (p.u_first ? (size_t)uw * uh : 0);
	rot_store_tile(TU, pu, uw, uh, i0, j0);
	rot_store_tile(TV, pv, uw, uh, i0, j0);
}

// ------------------------------------------------------------------------------------------------ packed 4:2:2 -> I420
// MSPixConv's YUYV / UYVY / YUY2 inputs (src/videofilters/pixconv.c:62-94 -> ms_scaler_process at the same size): luma
// copied, chroma = rounded average of the two source lines (what libswscale's unscaled yuyv/uyvy -> yuv420p converters
// produce; pinned in tests/golden). One thread: 8 luma pixels x 2 rows (two 16-byte loads) -> 2 x 8 Y bytes, 4 U, 4 V.
__global__ void __launch_bounds__(256) packed422_to_i420_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                                int w, int h, int uyvy) {
	const int groups = w / 8, rows2 = h / 2;
	const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= (long)groups * rows2) return;
	const int cy = (int)(t / groups), gx = (int)(t % groups);
	const size_t frame = blockIdx.y;
	const uint8_t *fs = src + frame * ((size_t)w * h * 2);
	uint8_t *fd = dst + frame * ((size_t)w * h * 3 / 2);
	const uint4 a = __ldg(reinterpret_cast<const uint4 *>(fs + (size_t)(2 * cy) * w * 2) + gx);
	const uint4 b = __ldg(reinterpret_cast<const uint4 *>(fs + (size_t)(2 * cy + 1) * w * 2) + gx);
	// bytes of a word: YUYV = Y0 U Y1 V, UYVY = U Y0 V Y1
	const unsigned ysel = uyvy ? 0x7531 : 0x6420, csel = uyvy ? 0x6420 : 0x7531;
	const unsigned ya0 = __byte_perm(a.x, a.y, ysel), ya1 = __byte_perm(a.z, a.w, ysel);
	const unsigned yb0 = __byte_perm(b.x, b.y, ysel), yb1 = __byte_perm(b.z, b.w, ysel);
	const unsigned ca0 = __byte_perm(a.x, a.y, csel), ca1 = __byte_perm(a.z, a.w, csel); // U V U V
	const unsigned cb0 = __byte_perm(b.x, b.y, csel), cb1 = __byte_perm(b.z, b.w, csel);
	// per-byte (x + y + 1) >> 1 — except in the row's last group when w % 16 == 8: the library's x86 row function rounds over
	// whole groups of 8 chroma samples only and truncates in its scalar tail (oracle_video.c, pinned on the live library)
	const bool tail = gx * 4 >= ((w / 2) & ~7);
	const unsigned m0 = tail ? __vhaddu4(ca0, cb0) : __vavgu4(ca0, cb0), m1 = tail ? __vhaddu4(ca1, cb1) : __vavgu4(ca1, cb1);
	const unsigned u4 = __byte_perm(m0, m1, 0x6420), v4 = __byte_perm(m0, m1, 0x7531);
	*reinterpret_cast<uint2 *>(fd + (size_t)(2 * cy) * w + (size_t)gx * 8) = make_uint2(ya0, ya1);
	*reinterpret_cast<uint2 *>(fd + (size_t)(2 * cy + 1) * w + (size_t)gx * 8) = make_uint2(yb0, yb1);
	uint8_t *pu = fd + (size_t)w * h, *pv = pu + (size_t)(w / 2) * (h / 2);
	*reinterpret_cast<unsigned *>(pu + (size_t)cy * (w / 2) + (size_t)gx * 4) = u4;
	*reinterpret_cast<unsigned *>(pv + (size_t)cy * (w / 2) + (size_t)gx * 4) = v4;
}

int msb200i_packed422_to_i420(msb200_ctx *ctx, int n_frames, const void *d_src, int w, int h, int uyvy, void *d_dst) {
	MSB200_CHECK_ARG(ctx && d_src && d_dst && n_frames > 0 && n_frames <= 65535 && w > 0 && h > 0 && (w % 8) == 0 && (h % 2) == 0);
	MSB200_CHECK_ARG(((uintptr_t)d_src % 16) == 0 && ((uintptr_t)d_dst % 8) == 0);
	const long threads = (long)(w / 8) * (h / 2);
	dim3 grid((unsigned)((threads + 255) / 256), (unsigned)n_frames);
	MSB200_LAUNCH(ctx, packed422_to_i420_kernel, grid, 256, 0, (const uint8_t *)d_src, (uint8_t *)d_dst, w, h, uyvy);
	return MSB200_OK;
}

// ------------------------------------------------------------------------------------------------ packed RGB -> I420
// MSPixConv's MS_RGB24 / MS_RGB24_REV inputs (src/videofilters/pixconv.c:62-94 -> ms_scaler_process at the same size)
// as libswscale 9.1 converts them (arithmetic and its pinning: oracle/oracle_video.c rgb_to_i420()):
//   RGB24: generic scaler path — Y through rgb24ToY / hScale16To15 / yuv2plane1, chroma from horizontal pixel pairs at
//          full height, then the 2:1 vertical bilinear filter {512,1536,1536,512}/4096 over rows 2y-1..2y+2 (clamped);
//   BGR24: the unscaled special converter (ff_rgb24toyv12): truncating Q15 luma, chroma from the 2x2 component means.
// One thread: 4 pixels x 2 rows of luma (two chroma samples); rows are read as three aligned 32-bit words (w % 4 == 0).
// HBM-bound: 3 B/pixel in + 1.5 B/pixel out; the two halo rows of the RGB24 path come from L1/L2 (neighbour threads).
#define RGB_RY 8414
#define RGB_GY 16519
#define RGB_BY 3208
#define RGB_RU (-4865)
#define RGB_GU (-9528)
#define RGB_BU 14392
#define RGB_RV 14392
#define RGB_GV (-12061)
#define RGB_BV (-2332)
__device__ __forceinline__ void rgb_unpack4(const uint8_t *row, int c[4][3]) { // 12 bytes = 4 pixels x 3 components
	const uint3 wd = *reinterpret_cast<const uint3 *>(row);
	const unsigned w3[3] = {wd.x, wd.y, wd.z};
#pragma unroll
	for (int i = 0; i < 12; ++i) c[i / 3][i % 3] = (int)((w3[i >> 2] >> (8 * (i & 3))) & 255u);
}
// 16 bytes = 4 pixels x 4 bytes (RGBA / BGRA): components moved to (R, G, B) order, alpha dropped
template <bool SWAP_RB>
__device__ __forceinline__ void rgbx_unpack4(const uint8_t *row, int c[4][3]) {
	const uint4 wd = *reinterpret_cast<const uint4 *>(row);
	const unsigned w4[4] = {wd.x, wd.y, wd.z, wd.w};
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		c[k][SWAP_RB ? 2 : 0] = (int)(w4[k] & 255u);
		c[k][1] = (int)((w4[k] >> 8) & 255u);
		c[k][SWAP_RB ? 0 : 2] = (int)((w4[k] >> 16) & 255u);
	}
}
// 8 bytes = 4 pixels of RGB565 (little endian): the fields widened by plain shifts (r5 << 3, g6 << 2, b5 << 3), which is
// what libswscale's 16-bit reader amounts to (oracle/oracle_video.c, pinned against the live library)
__device__ __forceinline__ void rgb565_unpack4(const uint8_t *row, int c[4][3]) {
	const uint2 wd = *reinterpret_cast<const uint2 *>(row);
	const unsigned px[4] = {wd.x & 0xffffu, wd.x >> 16, wd.y & 0xffffu, wd.y >> 16};
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		c[k][0] = (int)((px[k] >> 11) << 3);
		c[k][1] = (int)(((px[k] >> 5) & 63u) << 2);
		c[k][2] = (int)((px[k] & 31u) << 3);
	}
}
// FMT: 0 = RGB24 (generic path), 1 = BGR24 (special converter), 2 = RGBA, 3 = BGRA (generic path, alpha ignored),
// 4 = RGB565 (generic path)
template <int FMT, bool X86 = false>
__global__ void __launch_bounds__(256) rgb24_to_i420_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int w, int h) {
	constexpr bool BGR = FMT == 1;
	constexpr int BPP = FMT == 4 ? 2 : (FMT >= 2 ? 4 : 3);
	const int groups = w / 4, rows2 = h / 2;
	const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= (long)groups * rows2) return;
	const int cy = (int)(t / groups), gx = (int)(t % groups);
	const size_t frame = blockIdx.y, pitch = (size_t)w * BPP;
	const uint8_t *fs = src + frame * (pitch * h) + (size_t)gx * 4 * BPP;
	auto rgb_unpack4 = [](const uint8_t *row, int(&c)[4][3]) {
		if (FMT == 2) rgbx_unpack4<false>(row, c);
		else if (FMT == 3) rgbx_unpack4<true>(row, c);
		else if (FMT == 4) rgb565_unpack4(row, c);
		else ::rgb_unpack4(row, c);
	};
	uint8_t *fd = dst + frame * ((size_t)w * h * 3 / 2);
	uint8_t *pu = fd + (size_t)w * h, *pv = pu + (size_t)(w / 2) * (h / 2);
	int a[4][3], b[4][3]; // the block's two rows
	rgb_unpack4(fs + (size_t)(2 * cy) * pitch, a);
	rgb_unpack4(fs + (size_t)(2 * cy + 1) * pitch, b);
	unsigned ya = 0, yb = 0;
	if (BGR) {
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			ya |= (unsigned)(((RGB_RY * a[k][2] + RGB_GY * a[k][1] + RGB_BY * a[k][0]) >> 15) + 16) << (8 * k);
			yb |= (unsigned)(((RGB_RY * b[k][2] + RGB_GY * b[k][1] + RGB_BY * b[k][0]) >> 15) + 16) << (8 * k);
		}
		unsigned u2 = 0, v2 = 0;
#pragma unroll
		for (int s = 0; s < 2; ++s) {
			int m[3];
#pragma unroll
			for (int k = 0; k < 3; ++k) m[k] = (a[2 * s][k] + a[2 * s + 1][k] + b[2 * s][k] + b[2 * s + 1][k]) >> 2;
			u2 |= (unsigned)((((RGB_RU * m[2] + RGB_GU * m[1] + RGB_BU * m[0]) >> 15) + 128) & 255) << (8 * s);
			v2 |= (unsigned)((((RGB_RV * m[2] + RGB_GV * m[1] + RGB_BV * m[0]) >> 15) + 128) & 255) << (8 * s);
		}
		*reinterpret_cast<unsigned short *>(pu + (size_t)cy * (w / 2) + (size_t)gx * 2) = (unsigned short)u2;
		*reinterpret_cast<unsigned short *>(pv + (size_t)cy * (w / 2) + (size_t)gx * 2) = (unsigned short)v2;
	} else {
		auto luma = [](const int(&c)[3]) {
			int v = (RGB_RY * c[0] + RGB_GY * c[1] + RGB_BY * c[2] + (32 << 14) + (1 << 8)) >> 9;
			v = min(v * 2, 32767); // hScale16To15 with its single 1 << 14 tap: (v * 16384) >> 13
			v = (v + 64) >> 7;
			return (unsigned)min(max(v, 0), 255);
		};
		// 15-bit chroma of a pixel pair
		auto chroma = [](const int(&p)[3], const int(&q)[3], int &u, int &v) {
			const int r = p[0] + q[0], g = p[1] + q[1], bl = p[2] + q[2];
			u = min(((RGB_RU * r + RGB_GU * g + RGB_BU * bl + (256 << 15) + (1 << 9)) >> 10) * 2, 32767);
			v = min(((RGB_RV * r + RGB_GV * g + RGB_BV * bl + (256 << 15) + (1 << 9)) >> 10) * 2, 32767);
		};
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			ya |= luma(a[k]) << (8 * k);
			yb |= luma(b[k]) << (8 * k);
		}
		int above[4][3], below[4][3]; // rows 2cy-1 and 2cy+2, folded onto the picture's border rows
		rgb_unpack4(fs + (size_t)max(2 * cy - 1, 0) * pitch, above);
		rgb_unpack4(fs + (size_t)min(2 * cy + 2, h - 1) * pitch, below);
		unsigned u2 = 0, v2 = 0;
#pragma unroll
		for (int s = 0; s < 2; ++s) {
			int u[4], v[4];
			chroma(above[2 * s], above[2 * s + 1], u[0], v[0]);
			chroma(a[2 * s], a[2 * s + 1], u[1], v[1]);
			chroma(b[2 * s], b[2 * s + 1], u[2], v[2]);
			chroma(below[2 * s], below[2 * s + 1], u[3], v[3]);
			int su, sv;
			if (X86 && cy < h / 2 - 1) {
				// libswscale's x86 SIMD vertical scaler (what a plain SWS_BILINEAR call runs): every product loses its low 16
				// bits before the sum, rounder (64 + 8 * 3) >> 4, final >> 3; the last chroma row is done by the C function.
				// At the top border the library's filter is FOLDED (taps 2048, 1536, 512 on rows 0, 1, 2), not replicated rows
				if (cy == 0) {
					su = (5 + ((u[1] * 2048) >> 16) + ((u[2] * 1536) >> 16) + ((u[3] * 512) >> 16)) >> 3;
					sv = (5 + ((v[1] * 2048) >> 16) + ((v[2] * 1536) >> 16) + ((v[3] * 512) >> 16)) >> 3;
				} else {
					su = (5 + ((u[0] * 512) >> 16) + ((u[1] * 1536) >> 16) + ((u[2] * 1536) >> 16) + ((u[3] * 512) >> 16)) >> 3;
					sv = (5 + ((v[0] * 512) >> 16) + ((v[1] * 1536) >> 16) + ((v[2] * 1536) >> 16) + ((v[3] * 512) >> 16)) >> 3;
				}
			} else {
				su = ((64 << 12) + 512 * (u[0] + u[3]) + 1536 * (u[1] + u[2])) >> 19;
				sv = ((64 << 12) + 512 * (v[0] + v[3]) + 1536 * (v[1] + v[2])) >> 19;
			}
			u2 |= (unsigned)min(max(su, 0), 255) << (8 * s);
			v2 |= (unsigned)min(max(sv, 0), 255) << (8 * s);
		}
		*reinterpret_cast<unsigned short *>(pu + (size_t)cy * (w / 2) + (size_t)gx * 2) = (unsigned short)u2;
		*reinterpret_cast<unsigned short *>(pv + (size_t)cy * (w / 2) + (size_t)gx * 2) = (unsigned short)v2;
	}
	*reinterpret_cast<unsigned *>(fd + (size_t)(2 * cy) * w + (size_t)gx * 4) = ya;
	*reinterpret_cast<unsigned *>(fd + (size_t)(2 * cy + 1) * w + (size_t)gx * 4) = yb;
}

// The generic path again, marching: one thread walks RGB_MARCH row pairs of its 4-pixel column group and carries the
// 15-bit chroma of the two rows it shares with the next pair in registers, so every source row is unpacked (and its chroma
// computed) once — plus one halo row at each end of the walk — instead of twice (the kernel above reads rows 2cy-1 .. 2cy+2
// for every cy: 4 rows per 2). Same arithmetic, same bytes.
#define RGB_MARCH 4
template <int FMT, bool X86>
__global__ void __launch_bounds__(256) rgb_to_i420_march_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int w, int h) {
	constexpr int BPP = FMT == 4 ? 2 : (FMT >= 2 ? 4 : 3);
	const int groups = w / 4, rows2 = h / 2, walks = (rows2 + RGB_MARCH - 1) / RGB_MARCH;
	const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= (long)groups * walks) return;
	const int cy0 = (int)(t / groups) * RGB_MARCH, gx = (int)(t % groups), cy1 = min(cy0 + RGB_MARCH, rows2);
	const size_t frame = blockIdx.y, pitch = (size_t)w * BPP;
	const uint8_t *fs = src + frame * (pitch * h) + (size_t)gx * 4 * BPP;
	uint8_t *fd = dst + frame * ((size_t)w * h * 3 / 2);
	uint8_t *py = fd + (size_t)gx * 4, *pu = fd + (size_t)w * h + (size_t)gx * 2, *pv = pu + (size_t)(w / 2) * (h / 2);
	auto unpack = [&](int row, int(&c)[4][3]) {
		const uint8_t *p = fs + (size_t)row * pitch;
		if (FMT == 2) rgbx_unpack4<false>(p, c);
		else if (FMT == 3) rgbx_unpack4<true>(p, c);
		else if (FMT == 4) rgb565_unpack4(p, c);
		else rgb_unpack4(p, c);
	};
	auto luma4 = [](const int(&c)[4][3]) {
		unsigned y = 0;
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			int v = (RGB_RY * c[k][0] + RGB_GY * c[k][1] + RGB_BY * c[k][2] + (32 << 14) + (1 << 8)) >> 9;
			v = (min(v * 2, 32767) + 64) >> 7; // hScale16To15 with its single 1 << 14 tap, then yuv2plane1
			y |= (unsigned)min(max(v, 0), 255) << (8 * k);
		}
		return y;
	};
	auto chroma2 = [](const int(&c)[4][3], int(&u)[2], int(&v)[2]) { // 15-bit chroma of the two pixel pairs
#pragma unroll
		for (int s = 0; s < 2; ++s) {
			const int r = c[2 * s][0] + c[2 * s + 1][0], g = c[2 * s][1] + c[2 * s + 1][1], bl = c[2 * s][2] + c[2 * s + 1][2];
			u[s] = min(((RGB_RU * r + RGB_GU * g + RGB_BU * bl + (256 << 15) + (1 << 9)) >> 10) * 2, 32767);
			v[s] = min(((RGB_RV * r + RGB_GV * g + RGB_BV * bl + (256 << 15) + (1 << 9)) >> 10) * 2, 32767);
		}
	};
	int px[4][3], up[2], vp[2], ua[2], va[2], ub[2], vb[2], un[2], vn[2];
	unpack(max(2 * cy0 - 1, 0), px);
	chroma2(px, up, vp);
	unpack(2 * cy0, px);
	chroma2(px, ua, va);
	*reinterpret_cast<unsigned *>(py + (size_t)(2 * cy0) * w) = luma4(px);
#pragma unroll
	for (int k = 0; k < RGB_MARCH; ++k) {
		const int cy = cy0 + k;
		if (cy >= cy1) break;
		unpack(2 * cy + 1, px);
		chroma2(px, ub, vb);
		*reinterpret_cast<unsigned *>(py + (size_t)(2 * cy + 1) * w) = luma4(px);
		unpack(min(2 * cy + 2, h - 1), px);
		chroma2(px, un, vn);
		if (cy + 1 < cy1) *reinterpret_cast<unsigned *>(py + (size_t)(2 * cy + 2) * w) = luma4(px);
		unsigned u2 = 0, v2 = 0;
#pragma unroll
		for (int s = 0; s < 2; ++s) {
			int su, sv;
			if (X86 && cy < rows2 - 1) { // see rgb24_to_i420_kernel
				if (cy == 0) {
					su = (5 + ((ua[s] * 2048) >> 16) + ((ub[s] * 1536) >> 16) + ((un[s] * 512) >> 16)) >> 3;
					sv = (5 + ((va[s] * 2048) >> 16) + ((vb[s] * 1536) >> 16) + ((vn[s] * 512) >> 16)) >> 3;
				} else {
					su = (5 + ((up[s] * 512) >> 16) + ((ua[s] * 1536) >> 16) + ((ub[s] * 1536) >> 16) + ((un[s] * 512) >> 16)) >> 3;
					sv = (5 + ((vp[s] * 512) >> 16) + ((va[s] * 1536) >> 16) + ((vb[s] * 1536) >> 16) + ((vn[s] * 512) >> 16)) >> 3;
				}
			} else {
				su = ((64 << 12) + 512 * (up[s] + un[s]) + 1536 * (ua[s] + ub[s])) >> 19;
				sv = ((64 << 12) + 512 * (vp[s] + vn[s]) + 1536 * (va[s] + vb[s])) >> 19;
			}
			u2 |= (unsigned)min(max(su, 0), 255) << (8 * s);
			v2 |= (unsigned)min(max(sv, 0), 255) << (8 * s);
			up[s] = ub[s]; vp[s] = vb[s];
			ua[s] = un[s]; va[s] = vn[s];
		}
		*reinterpret_cast<unsigned short *>(pu + (size_t)cy * (w / 2)) = (unsigned short)u2;
		*reinterpret_cast<unsigned short *>(pv + (size_t)cy * (w / 2)) = (unsigned short)v2;
	}
}

// fmt: 0 RGB24, 1 BGR24, 2 RGBA, 3 BGRA, 4 RGB565
int msb200i_rgb24_to_i420(msb200_ctx *ctx, int n_frames, const void *d_src, int w, int h, int fmt, void *d_dst, int x86_vertical) {
	MSB200_CHECK_ARG(ctx && d_src && d_dst && n_frames > 0 && n_frames <= 65535 && w > 0 && h > 0 && (w % 4) == 0 && (h % 2) == 0);
	MSB200_CHECK_ARG(fmt >= 0 && fmt <= 4 && ((uintptr_t)d_src % (fmt == 2 || fmt == 3 ? 16 : fmt == 4 ? 8 : 4)) == 0 && ((uintptr_t)d_dst % 4) == 0);
	const long threads = (long)(w / 4) * (h / 2);
	dim3 grid((unsigned)((threads + 255) / 256), (unsigned)n_frames);
	static const bool block_kernel = getenv("MSB200_RGB_BLOCK_KERNEL") != nullptr;
#define RGB_LAUNCH(F)                                                                                                  \
	do {                                                                                                               \
		const long walks = (long)(w / 4) * ((h / 2 + RGB_MARCH - 1) / RGB_MARCH);                                      \
		const dim3 gm((unsigned)((walks + 255) / 256), (unsigned)n_frames);                                            \
		if (block_kernel) { /* A/B: the one-block-per-thread kernel */                                                 \
			if (x86_vertical) MSB200_LAUNCH(ctx, (rgb24_to_i420_kernel<F, true>), grid, 256, 0, (const uint8_t *)d_src, (uint8_t *)d_dst, w, h); \
			else MSB200_LAUNCH(ctx, (rgb24_to_i420_kernel<F, false>), grid, 256, 0, (const uint8_t *)d_src, (uint8_t *)d_dst, w, h); \
		} else if (x86_vertical) MSB200_LAUNCH(ctx, (rgb_to_i420_march_kernel<F, true>), gm, 256, 0, (const uint8_t *)d_src, (uint8_t *)d_dst, w, h); \
		else MSB200_LAUNCH(ctx, (rgb_to_i420_march_kernel<F, false>), gm, 256, 0, (const uint8_t *)d_src, (uint8_t *)d_dst, w, h); \
	} while (0)
	switch (fmt) {
		case 0: RGB_LAUNCH(0); break;
		case 1: MSB200_LAUNCH(ctx, (rgb24_to_i420_kernel<1, false>), grid, 256, 0, (const uint8_t *)d_src, (uint8_t *)d_dst, w, h); break;
		case 2: RGB_LAUNCH(2); break;
		case 3: RGB_LAUNCH(3); break;
		default: RGB_LAUNCH(4); break;
	}
#undef RGB_LAUNCH
	return MSB200_OK;
}

extern "C" {

int msb200_nv12_to_i420_dev(msb200_ctx *ctx, int n_frames, const void *d_src, size_t src_frame_bytes, size_t cbcr_offset,
                            int rotation, int w, int h, int y_stride, int cbcr_stride, int u_first, int down_scale,
                            void *d_dst) {
	MSB200_CHECK_ARG(ctx && d_src && d_dst && n_frames > 0 && n_frames <= 65535 && w > 0 && h > 0 && (w % 2) == 0 && (h % 2) == 0);
	MSB200_CHECK_ARG(rotation == 0 || rotation == 90 || rotation == 180 || rotation == 270);
	Nv12Params p;
	p.rotation = rotation;
	p.w = w;
	p.h = h;
	p.ys = y_stride;
	p.cs = cbcr_stride;
	p.u_first = u_first;
	p.f = down_scale ? 2 : 1;
	p.src_frame_bytes = src_frame_bytes;
	p.cbcr_offset = cbcr_offset;
	p.dst_frame_bytes = (size_t)w * h * 3 / 2;
	const int sw = (rotation % 180 == 0 ? w : h) * p.f, sh = (rotation % 180 == 0 ? h : w) * p.f;
	MSB200_CHECK_ARG(y_stride >= sw && cbcr_stride >= sw && cbcr_offset >= (size_t)y_stride * (sh - 1) + sw);
	MSB200_CHECK_ARG(src_frame_bytes >= cbcr_offset + (size_t)cbcr_stride * (sh / 2 - 1) + sw);
	const bool fast = rotation == 0 && !down_scale && (w % 32) == 0 && (y_stride % 16) == 0 && (cbcr_stride % 16) == 0 &&
	                  (cbcr_offset % 16) == 0 && (src_frame_bytes % 16) == 0 && ((uintptr_t)d_src % 16) == 0 &&
	                  ((uintptr_t)d_dst % 16) == 0 && (((size_t)w * h / 4) % 16) == 0;
	const bool tiled_rot = (rotation == 90 || rotation == 270) && !down_scale && (w % 8) == 0 && (h % 8) == 0 && (y_stride % 4) == 0 &&
	                       (cbcr_stride % 8) == 0 && (cbcr_offset % 8) == 0 && (src_frame_bytes % 8) == 0 &&
	                       ((uintptr_t)d_src % 8) == 0 && ((uintptr_t)d_dst % 4) == 0;
	if (tiled_rot) {
		const int n_lt = ((w + ROT_T - 1) / ROT_T) * ((h + ROT_T - 1) / ROT_T);
		const int n_ct = ((w / 2 + ROT_T - 1) / ROT_T) * ((h / 2 + ROT_T - 1) / ROT_T);
		dim3 grid((unsigned)(n_lt + n_ct), (unsigned)n_frames);
		if (rotation == 90) MSB200_LAUNCH(ctx, nv12_rot_kernel<90>, grid, 256, 0, (const uint8_t *)d_src, (uint8_t *)d_dst, p);
		else MSB200_LAUNCH(ctx, nv12_rot_kernel<270>, grid, 256, 0, (const uint8_t *)d_src, (uint8_t *)d_dst, p);
	} else if (fast) {
		long threads = (long)(w / 16) * h + (long)(w / 32) * (h / 2);
		dim3 grid((unsigned)((threads + 255) / 256), (unsigned)n_frames);
		MSB200_LAUNCH(ctx, nv12_fast_kernel, grid, 256, 0, (const uint8_t *)d_src, (uint8_t *)d_dst, p);
	} else {
		long quads = ((long)w * h + 3) / 4 + 2 * (((long)(w / 2) * (h / 2) + 3) / 4);
		dim3 grid((unsigned)((quads + 255) / 256), (unsigned)n_frames);
		MSB200_LAUNCH(ctx, nv12_generic_kernel, grid, 256, 0, (const uint8_t *)d_src, (uint8_t *)d_dst, p);
	}
	return MSB200_OK;
}

int msb200_nv12_to_i420(msb200_ctx *ctx, int n_frames, const uint8_t *src, size_t src_frame_bytes, size_t cbcr_offset,
                        int rotation, int w, int h, int y_stride, int cbcr_stride, int u_first, int down_scale, uint8_t *dst) {
	MSB200_CHECK_ARG(ctx && src && dst && n_frames > 0);
	void *d_src = nullptr, *d_dst = nullptr;
	size_t in_bytes = src_frame_bytes * (size_t)n_frames, out_bytes = (size_t)w * h * 3 / 2 * (size_t)n_frames;
	MSB200_CUDA(cudaMalloc(&d_src, in_bytes));
	MSB200_CUDA(cudaMalloc(&d_dst, out_bytes));
	cudaStream_t s = ctx->stream;
	int r = MSB200_OK;
	if (cudaMemcpyAsync(d_src, src, in_bytes, cudaMemcpyHostToDevice, s) != cudaSuccess) r = MSB200_ECUDA;
	if (!r) r = msb200_nv12_to_i420_dev(ctx, n_frames, d_src, src_frame_bytes, cbcr_offset, rotation, w, h, y_stride, cbcr_stride, u_first, down_scale, d_dst);
	if (!r && cudaMemcpyAsync(dst, d_dst, out_bytes, cudaMemcpyDeviceToHost, s) != cudaSuccess) r = MSB200_ECUDA;
	if (cudaStreamSynchronize(s) != cudaSuccess && !r) r = MSB200_ECUDA;
	cudaFree(d_src);
	cudaFree(d_dst);
	if (r == MSB200_ECUDA) msb200_set_error("nv12_to_i420: CUDA failure: %s", cudaGetErrorString(cudaGetLastError()));
	return r;
}

} // extern "C"
