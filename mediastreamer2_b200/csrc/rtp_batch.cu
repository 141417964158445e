// rtp_batch.cu — batched RTP payload hand-off either side of the G.711 banks.
//
// The reference moves every packet through its own MSFilter and mblk_t: MSRtpRecv takes a packet from the session, copies
// the header fields into the block's meta data and advances b_rptr to the payload (receiver_process,
// /root/reference/src/otherfilters/msrtp.c:1050-1092: mblk_set_timestamp_info(rtp_get_timestamp), mblk_set_marker_info,
// mblk_set_cseq, rtp_get_payload); the decoder then allocates one output block per packet (alaw.c:199-211). On the way
// out MSRtpSend prepends a header block per packet (_sender_process :617-705). With thousands of streams per ticker that
// per-packet work — not the companding — is what the host spends its tick on.
//
// Here one object per direction holds a pinned arena for ALL streams of a ticker:
//   rx  push(stream, packet): the RTP header is parsed in place (RFC 3550 §5.1: version, padding, extension, CSRC count,
//       marker, payload type, sequence number, timestamp, SSRC; §5.3.1 header extension) and ONLY the payload bytes are
//       copied, straight into the stream's row of the arena; decode(): one H2D of the rows in use, one g711_decode_kernel
//       launch over the whole arena, one D2H (or none: decode_dev leaves the PCM in HBM for the next bank)
//   tx  encode(): one H2D of the PCM rows, one g711_encode_kernel launch, one strided D2H that lands every payload right
//       behind its 12-byte header in the packet arena; the headers (sequence + 1, timestamp + samples per packet, marker)
//       are written by the host while the GPU works
// No sockets, no jitter buffer, no RTCP: those stay the host's (oRTP's) business.
#include "msb200_internal.h"

struct msb200_rtp_rx {
	msb200_ctx *ctx;
	int n, law, max_payload, hi; // hi: highest stream pushed this tick + 1
	uint8_t *h_codes;            // pinned [n][max_payload]
	void *d_codes, *d_pcm;       // device [n][max_payload] u8 / s16
	std::vector<msb200_rtp_meta> meta;
};
struct msb200_rtp_tx {
	msb200_ctx *ctx;
	int n, law, samples;
	uint8_t *h_packets; // pinned [n][12 + samples]
	void *d_pcm, *d_codes;
	std::vector<uint32_t> ssrc, ts;
	std::vector<uint16_t> seq;
	std::vector<uint8_t> pt;
};

static inline uint16_t be16(const uint8_t *p) {
	return (uint16_t)((p[0] << 8) | p[1]);
}
static inline uint32_t be32(const uint8_t *p) {
	return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}

extern "C" {

// RFC 3550 §5.1 / §5.3.1: where the payload of a packet starts and how long it is. Returns 0 and fills *meta (payload_len
// >= 0), or MSB200_EINVAL for anything that is not a well-formed RTP version 2 packet.
int msb200_rtp_parse(const uint8_t *packet, size_t len, msb200_rtp_meta *meta, size_t *payload_offset) {
	MSB200_CHECK_ARG(packet && meta && payload_offset);
	if (len < 12 || (packet[0] >> 6) != 2) {
		msb200_set_error("rtp: not a version 2 packet of at least 12 bytes (%zu bytes, first byte 0x%02x)", len, len ? packet[0] : 0);
		return MSB200_EINVAL;
	}
	const int padding = (packet[0] >> 5) & 1, extension = (packet[0] >> 4) & 1, cc = packet[0] & 15;
	size_t off = 12 + 4 * (size_t)cc;
	if (off > len) {
		msb200_set_error("rtp: %d CSRC entries do not fit %zu bytes", cc, len);
		return MSB200_EINVAL;
	}
	if (extension) {
		if (off + 4 > len) {
			msb200_set_error("rtp: truncated header extension");
			return MSB200_EINVAL;
		}
		off += 4 + 4 * (size_t)be16(packet + off + 2);
		if (off > len) {
			msb200_set_error("rtp: header extension longer than the packet");
			return MSB200_EINVAL;
		}
	}
	size_t end = len;
	if (padding) {
		const size_t pad = packet[len - 1];
		if (pad == 0 || off + pad > len) {
			msb200_set_error("rtp: bad padding count %zu", pad);
			return MSB200_EINVAL;
		}
		end -= pad;
	}
	meta->marker = (uint8_t)(packet[1] >> 7);
	meta->payload_type = (uint8_t)(packet[1] & 127);
	meta->seq = be16(packet + 2);
	meta->timestamp = be32(packet + 4);
	meta->ssrc = be32(packet + 8);
	meta->payload_len = (int32_t)(end - off);
	*payload_offset = off;
	return MSB200_OK;
}

// ------------------------------------------------------------------------------------------------ receive
int msb200_rtp_rx_create(msb200_ctx *ctx, int n_streams, int law, int max_payload, msb200_rtp_rx **out) {
	MSB200_CHECK_ARG(ctx && out && n_streams > 0 && max_payload > 0 && (law == MSB200_G711_ALAW || law == MSB200_G711_ULAW));
	msb200_rtp_rx *r = new msb200_rtp_rx();
	r->ctx = ctx;
	r->n = n_streams;
	r->law = law;
	r->max_payload = (max_payload + 15) & ~15; // rows stay 16-byte aligned for the kernel's vector accesses
	r->hi = 0;
	r->meta.assign((size_t)n_streams, msb200_rtp_meta{});
	const size_t cells = (size_t)n_streams * r->max_payload;
	MSB200_CUDA(cudaSetDevice(ctx->device));
	MSB200_CUDA(cudaHostAlloc((void **)&r->h_codes, cells, cudaHostAllocDefault));
	MSB200_CUDA(cudaMalloc(&r->d_codes, cells));
	MSB200_CUDA(cudaMalloc(&r->d_pcm, cells * 2));
	memset(r->h_codes, law == MSB200_G711_ALAW ? 0xD5 : 0xFF, cells); // the codes of silence
	*out = r;
	return MSB200_OK;
}
void msb200_rtp_rx_destroy(msb200_rtp_rx *r) {
	if (!r) return;
	cudaStreamSynchronize(r->ctx->stream);
	cudaFreeHost(r->h_codes);
	cudaFree(r->d_codes);
	cudaFree(r->d_pcm);
	delete r;
}
int msb200_rtp_rx_row_samples(const msb200_rtp_rx *r) {
	return r ? r->max_payload : 0;
}
int msb200_rtp_rx_begin_tick(msb200_rtp_rx *r) {
	MSB200_CHECK_ARG(r);
	for (int i = 0; i < r->hi; ++i) r->meta[(size_t)i].payload_len = 0;
	r->hi = 0;
	return MSB200_OK;
}
int msb200_rtp_rx_push_payload(msb200_rtp_rx *r, int stream, const uint8_t *payload, int len, const msb200_rtp_meta *meta) {
	MSB200_CHECK_ARG(r && stream >= 0 && stream < r->n && (payload || len == 0) && len >= 0 && meta);
	if (len > r->max_payload) {
		msb200_set_error("rtp: payload of %d bytes exceeds the arena row of %d", len, r->max_payload);
		return MSB200_EINVAL;
	}
	memcpy(r->h_codes + (size_t)stream * r->max_payload, payload, (size_t)len);
	r->meta[(size_t)stream] = *meta;
	r->meta[(size_t)stream].payload_len = len;
	if (stream + 1 > r->hi) r->hi = stream + 1;
	return len;
}
// one packet of one stream: header parsed in place, payload copied once (into the arena). expected_pt < 0: any.
// A packet of another payload type is ignored (returns 0), as receiver_check_payload_type does (msrtp.c:1005-1020).
int msb200_rtp_rx_push(msb200_rtp_rx *r, int stream, const uint8_t *packet, size_t len, int expected_pt) {
	MSB200_CHECK_ARG(r && stream >= 0 && stream < r->n);
	msb200_rtp_meta m;
	size_t off = 0;
	int rc = msb200_rtp_parse(packet, len, &m, &off);
	if (rc) return rc;
	if (expected_pt >= 0 && m.payload_type != expected_pt) return 0;
	return msb200_rtp_rx_push_payload(r, stream, packet + off, m.payload_len, &m);
}
int msb200_rtp_rx_decode_dev(msb200_rtp_rx *r, void **d_pcm, const msb200_rtp_meta **meta) {
	MSB200_CHECK_ARG(r && d_pcm);
	*d_pcm = r->d_pcm;
	if (meta) *meta = r->meta.data();
	if (r->hi == 0) return MSB200_OK;
	const size_t cells = (size_t)r->hi * r->max_payload;
	MSB200_CUDA(cudaMemcpyAsync(r->d_codes, r->h_codes, cells, cudaMemcpyHostToDevice, r->ctx->stream));
	return msb200_g711_decode_dev(r->ctx, r->law, r->d_codes, r->d_pcm, cells);
}
int msb200_rtp_rx_decode(msb200_rtp_rx *r, int16_t *pcm, msb200_rtp_meta *meta) {
	MSB200_CHECK_ARG(r && pcm);
	void *d = nullptr;
	int rc = msb200_rtp_rx_decode_dev(r, &d, nullptr);
	if (rc) return rc;
	if (r->hi > 0) {
		MSB200_CUDA(cudaMemcpyAsync(pcm, d, (size_t)r->hi * r->max_payload * 2, cudaMemcpyDeviceToHost, r->ctx->stream));
		MSB200_CUDA(cudaStreamSynchronize(r->ctx->stream));
	}
	if (meta) memcpy(meta, r->meta.data(), sizeof(msb200_rtp_meta) * (size_t)r->n);
	return MSB200_OK;
}

// ------------------------------------------------------------------------------------------------ send
int msb200_rtp_tx_create(msb200_ctx *ctx, int n_streams, int law, int samples_per_packet, msb200_rtp_tx **out) {
	MSB200_CHECK_ARG(ctx && out && n_streams > 0 && samples_per_packet > 0 && samples_per_packet % 16 == 0 &&
	                 (law == MSB200_G711_ALAW || law == MSB200_G711_ULAW));
	msb200_rtp_tx *t = new msb200_rtp_tx();
	t->ctx = ctx;
	t->n = n_streams;
	t->law = law;
	t->samples = samples_per_packet;
	t->ssrc.assign((size_t)n_streams, 0);
	t->ts.assign((size_t)n_streams, 0);
	t->seq.assign((size_t)n_streams, 0);
	t->pt.assign((size_t)n_streams, law == MSB200_G711_ALAW ? 8 : 0); // static payload types PCMA / PCMU (RFC 3551)
	const size_t cells = (size_t)n_streams * samples_per_packet;
	MSB200_CUDA(cudaSetDevice(ctx->device));
	MSB200_CUDA(cudaHostAlloc((void **)&t->h_packets, (size_t)n_streams * (12 + (size_t)samples_per_packet), cudaHostAllocDefault));
	MSB200_CUDA(cudaMalloc(&t->d_pcm, cells * 2));
	MSB200_CUDA(cudaMalloc(&t->d_codes, cells));
	*out = t;
	return MSB200_OK;
}
void msb200_rtp_tx_destroy(msb200_rtp_tx *t) {
	if (!t) return;
	cudaStreamSynchronize(t->ctx->stream);
	cudaFreeHost(t->h_packets);
	cudaFree(t->d_pcm);
	cudaFree(t->d_codes);
	delete t;
}
int msb200_rtp_tx_set_stream(msb200_rtp_tx *t, int stream, uint32_t ssrc, int payload_type, uint16_t next_seq, uint32_t next_ts) {
	MSB200_CHECK_ARG(t && stream >= 0 && stream < t->n && payload_type >= 0 && payload_type < 128);
	t->ssrc[(size_t)stream] = ssrc;
	t->pt[(size_t)stream] = (uint8_t)payload_type;
	t->seq[(size_t)stream] = next_seq;
	t->ts[(size_t)stream] = next_ts;
	return MSB200_OK;
}
size_t msb200_rtp_tx_packet_bytes(const msb200_rtp_tx *t) {
	return t ? 12 + (size_t)t->samples : 0;
}
static int tx_run(msb200_rtp_tx *t, const void *d_pcm, const uint8_t *marker, const uint8_t *send, const uint8_t **packets) {
	const size_t pkt = 12 + (size_t)t->samples;
	int rc = msb200_g711_encode_dev(t->ctx, t->law, d_pcm, t->d_codes, (size_t)t->n * t->samples);
	if (rc) return rc;
	// every payload lands right behind its header: device rows of `samples` bytes -> host rows of 12 + samples
	MSB200_CUDA(cudaMemcpy2DAsync(t->h_packets + 12, pkt, t->d_codes, (size_t)t->samples, (size_t)t->samples, (size_t)t->n,
	                              cudaMemcpyDeviceToHost, t->ctx->stream));
	for (int s = 0; s < t->n; ++s) { // headers while the GPU works (RFC 3550 §5.1: V=2, P=0, X=0, CC=0)
		if (send && !send[s]) continue; // a stream with nothing to send this tick keeps its sequence number and clock
		uint8_t *h = t->h_packets + (size_t)s * pkt;
		const uint16_t seq = t->seq[(size_t)s]++;
		const uint32_t ts = t->ts[(size_t)s], ssrc = t->ssrc[(size_t)s];
		t->ts[(size_t)s] += (uint32_t)t->samples;
		h[0] = 0x80;
		h[1] = (uint8_t)((marker && marker[s] ? 0x80 : 0) | t->pt[(size_t)s]);
		h[2] = (uint8_t)(seq >> 8); h[3] = (uint8_t)seq;
		h[4] = (uint8_t)(ts >> 24); h[5] = (uint8_t)(ts >> 16); h[6] = (uint8_t)(ts >> 8); h[7] = (uint8_t)ts;
		h[8] = (uint8_t)(ssrc >> 24); h[9] = (uint8_t)(ssrc >> 16); h[10] = (uint8_t)(ssrc >> 8); h[11] = (uint8_t)ssrc;
	}
	MSB200_CUDA(cudaStreamSynchronize(t->ctx->stream));
	if (packets) *packets = t->h_packets;
	return MSB200_OK;
}
// pcm: host [n][samples] s16. *packets: the pinned packet arena [n][12 + samples], valid until the next call.
int msb200_rtp_tx_encode(msb200_rtp_tx *t, const int16_t *pcm, const uint8_t *marker, const uint8_t *send, const uint8_t **packets) {
	MSB200_CHECK_ARG(t && pcm && packets);
	MSB200_CUDA(cudaMemcpyAsync(t->d_pcm, pcm, (size_t)t->n * t->samples * 2, cudaMemcpyHostToDevice, t->ctx->stream));
	return tx_run(t, t->d_pcm, marker, send, packets);
}
// the PCM is already in HBM (the output of the previous bank), rows `samples` apart
int msb200_rtp_tx_encode_dev(msb200_rtp_tx *t, const void *d_pcm, const uint8_t *marker, const uint8_t *send, const uint8_t **packets) {
	MSB200_CHECK_ARG(t && d_pcm && packets);
	return tx_run(t, d_pcm, marker, send, packets);
}

} // extern "C"
