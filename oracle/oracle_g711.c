/* oracle_g711.c — TEST INFRASTRUCTURE (CPU checker; only tests/, smoke() and bench.py's CPU legs may load it).
 *
 * G.711 A-law / mu-law companding as the reference's MSAlawEnc/Dec and MSUlawEnc/Dec filters compute it:
 *   /root/reference/src/audiofilters/g711.c:119-146 (Snack_Lin2Alaw), :152-172 (Snack_Alaw2Lin),
 *   :208-238 (Snack_Lin2Mulaw), :249-262 (Snack_Mulaw2Lin); callers alaw.c:84-87, :207-209 and ulaw.c likewise.
 * Restated in closed form (segment number from the position of the leading one instead of the reference's table
 * search). Pinned: exhaustively (all 65536 PCM values, all 256 code words) against the UNMODIFIED reference functions
 * compiled into oracle/_ref/libms2ref.so — tests/test_oracle_vs_reference.py::test_g711_*. */
#include "msb200_oracle.h"

static int top_bit(unsigned v) { /* index of the most significant set bit, v > 0 */
	int n = 0;
	while (v >>= 1) ++n;
	return n;
}

static uint8_t lin2alaw(int16_t pcm) {
	int x = pcm >> 3; /* 13-bit signed magnitude domain */
	const int mask = x >= 0 ? 0xD5 : 0x55;
	if (x < 0) x = -x - 1;
	/* segments end at 0x1F, 0x3F, ..., 0xFFF: segment = leading-one position - 4, segment 0 below 32 */
	const int seg = x < 32 ? 0 : top_bit((unsigned)x) - 4;
	const int q = (x >> (seg < 2 ? 1 : seg)) & 0xF;
	return (uint8_t)(((seg << 4) | q) ^ mask);
}
static int16_t alaw2lin(uint8_t a) {
	a ^= 0x55;
	const int seg = (a & 0x70) >> 4;
	int t = (a & 0x0F) << 4;
	if (seg == 0) t += 8;
	else t = (t + 0x108) << (seg - 1);
	return (int16_t)((a & 0x80) ? t : -t);
}
static uint8_t lin2ulaw(int16_t pcm) {
	int x = pcm >> 2; /* 14-bit */
	const int mask = x < 0 ? 0x7F : 0xFF;
	if (x < 0) x = -x;
	if (x > 8159) x = 8159;
	x += 0x84 >> 2; /* bias 33 */
	/* segments end at 0x3F, 0x7F, ..., 0x1FFF */
	const int seg = x < 64 ? 0 : top_bit((unsigned)x) - 5;
	if (seg >= 8) return (uint8_t)(0x7F ^ mask);
	return (uint8_t)(((seg << 4) | ((x >> (seg + 1)) & 0xF)) ^ mask);
}
static int16_t ulaw2lin(uint8_t u) {
	u = (uint8_t)~u;
	int t = (((u & 0x0F) << 3) + 0x84) << ((u & 0x70) >> 4);
	return (int16_t)((u & 0x80) ? (0x84 - t) : (t - 0x84));
}

/* law: 0 = A-law (PCMA), 1 = mu-law (PCMU) */
void orc_g711_encode(int law, const int16_t *pcm, uint8_t *code, size_t n) {
	for (size_t i = 0; i < n; ++i) code[i] = law ? lin2ulaw(pcm[i]) : lin2alaw(pcm[i]);
}
void orc_g711_decode(int law, const uint8_t *code, int16_t *pcm, size_t n) {
	for (size_t i = 0; i < n; ++i) pcm[i] = law ? ulaw2lin(code[i]) : alaw2lin(code[i]);
}
