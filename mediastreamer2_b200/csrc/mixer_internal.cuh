// mixer_internal.cuh — pieces of the MSAudioMixer bank shared by audio.cu (single-GPU kernels) and mixer_xchg.cu (the
// cross-GPU exchange): the bank object, the reference's saturate / gain arithmetic, s16 vector packing.
// reference: /root/reference/src/audiofilters/audiomixer.c  (accumulate :33-38, saturate :40-44, apply_gain :46-51)
#pragma once
#include "msb200_internal.h"

__device__ __forceinline__ int mix_sat(int s) { // audiomixer.c:40-44 — clamps to [-32767, 32767]
	return s > 32767 ? 32767 : (s < -32767 ? -32767 : s);
}
__device__ __forceinline__ int mix_contrib(int s, float gain) { // :46-51, only when gain != 1.0 (:82)
	if (gain != 1.0f) return mix_sat(__float2int_rz(__fmul_rn(gain, (float)s)));
	return s;
}

struct msb200_mixer {
	msb200_ctx *ctx;
	int n_rooms, n_pins, nwords, conf_mode;
	int live;          // rooms [0, live) are processed (msb200_mixer_set_live); == n_rooms by default
	float *d_gain;     // [room][pin]
	uint8_t *d_active; // [room][pin]
	std::vector<float> h_gain;
	std::vector<uint8_t> h_active;
	bool dirty;
	msb200_devbuf in, present, out;
};

// One thread owns a column of VEC consecutive samples of one room and walks the pins twice: first to build the
// int32 sums, then to emit sat(sum - own) per pin. The second walk re-reads lines the same thread just touched
// (L1/L2 hits), so DRAM traffic stays at the algorithmic P*2n in + P*2n out.
template <int VEC> struct s16vec;
template <> struct s16vec<8> { typedef int4 type; };
template <> struct s16vec<4> { typedef int2 type; };
template <> struct s16vec<2> { typedef int type; };
template <> struct s16vec<1> { typedef short type; };

template <int VEC> __device__ __forceinline__ void unpack_s16(const typename s16vec<VEC>::type &v, int (&s)[VEC]);
template <> __device__ __forceinline__ void unpack_s16<8>(const int4 &v, int (&s)[8]) {
	s[0] = (short)(v.x & 0xffff); s[1] = v.x >> 16; s[2] = (short)(v.y & 0xffff); s[3] = v.y >> 16;
	s[4] = (short)(v.z & 0xffff); s[5] = v.z >> 16; s[6] = (short)(v.w & 0xffff); s[7] = v.w >> 16;
}
template <> __device__ __forceinline__ void unpack_s16<4>(const int2 &v, int (&s)[4]) {
	s[0] = (short)(v.x & 0xffff); s[1] = v.x >> 16; s[2] = (short)(v.y & 0xffff); s[3] = v.y >> 16;
}
template <> __device__ __forceinline__ void unpack_s16<2>(const int &v, int (&s)[2]) {
	s[0] = (short)(v & 0xffff); s[1] = v >> 16;
}
template <> __device__ __forceinline__ void unpack_s16<1>(const short &v, int (&s)[1]) {
	s[0] = v;
}
__device__ __forceinline__ int pack2(int lo, int hi) {
	return (lo & 0xffff) | (hi << 16);
}
template <int VEC> __device__ __forceinline__ typename s16vec<VEC>::type pack_s16(const int (&s)[VEC]);
template <> __device__ __forceinline__ int4 pack_s16<8>(const int (&s)[8]) {
	return make_int4(pack2(s[0], s[1]), pack2(s[2], s[3]), pack2(s[4], s[5]), pack2(s[6], s[7]));
}
template <> __device__ __forceinline__ int2 pack_s16<4>(const int (&s)[4]) {
	return make_int2(pack2(s[0], s[1]), pack2(s[2], s[3]));
}
template <> __device__ __forceinline__ int pack_s16<2>(const int (&s)[2]) {
	return pack2(s[0], s[1]);
}
template <> __device__ __forceinline__ short pack_s16<1>(const int (&s)[1]) {
	return (short)s[0];
}


int msb200i_mixer_upload(msb200_mixer *m); // pushes pending gain / active changes to the device (stream-ordered)
