/*
 * bjxa_oracle.c -- CPU restatement of libbjxa's block transform.
 *
 * TEST INFRASTRUCTURE ONLY (see bjxa_oracle.h).  Plain C99, libc only, one
 * scalar thread.  Written from the reference's behaviour, not from its text:
 * every function names the reference lines it restates ("ref:" = path:line
 * under /root/reference/).  Parity status: PINNED (see header).
 */
#include "bjxa_oracle.h"

#include <string.h>

/* ---------------------------------------------------------------------- */
/* little-endian field access (ref: src/libbjxa.c:99-166)                  */

static uint32_t
get_le(const uint8_t *p, unsigned nbytes)
{
	uint32_t v = 0;
	unsigned i;

	for (i = 0; i < nbytes; i++)
		v |= (uint32_t)p[i] << (8 * i);
	return (v);
}

static void
put_le(uint8_t *p, uint32_t v, unsigned nbytes)
{
	unsigned i;

	for (i = 0; i < nbytes; i++)
		p[i] = (uint8_t)(v >> (8 * i));
}

/* ---------------------------------------------------------------------- */
/* bit unpack / pack                                                        */

/*
 * ref: src/libbjxa.c:286-345.  Every ADPCM code lands in the TOP bits of an
 * int16: 4-bit codes are the high then the low nibble of each byte (:296-297),
 * 6-bit codes are the four fields of a big-endian 24-bit group (:315-321),
 * 8-bit codes are one byte each (:339).
 */
uint8_t
xao_inflate(unsigned bits, int16_t *dst, unsigned stride, const uint8_t *block)
{
	const uint8_t *pay = block + 1;
	unsigned n;

	for (n = 0; n < XAO_BLOCK_SAMPLES; n++) {
		unsigned bitpos = n * bits;	/* MSB-first position in payload */
		unsigned byte = bitpos >> 3;
		unsigned off = bitpos & 7;
		/* 16-bit big-endian window is enough: off + bits <= 6 + 6 */
		uint32_t win = (uint32_t)pay[byte] << 8;
		uint32_t code;

		if (off + bits > 8)
			win |= pay[byte + 1];
		code = (win >> (16 - off - bits)) & ((1u << bits) - 1u);
		dst[n * stride] = (int16_t)(uint16_t)(code << (16 - bits));
	}
	return (block[0]);
}

/*
 * ref: src/libbjxa.c:349-391.  Keeps the top `bits` bits of each sample
 * (logical shift of the uint16 image: truncation, no rounding) and lays the
 * codes out MSB-first, the exact inverse of xao_inflate.
 */
void
xao_deflate(unsigned bits, uint8_t *payload, const int16_t *src)
{
	unsigned n;

	memset(payload, 0, 4 * bits);
	for (n = 0; n < XAO_BLOCK_SAMPLES; n++) {
		uint32_t code = (uint32_t)(uint16_t)src[n] >> (16 - bits);
		unsigned bitpos = n * bits;
		unsigned byte = bitpos >> 3;
		unsigned off = bitpos & 7;
		uint32_t win = code << (16 - off - bits);

		payload[byte] |= (uint8_t)(win >> 8);
		if (off + bits > 8)
			payload[byte + 1] |= (uint8_t)win;
	}
}

/* ---------------------------------------------------------------------- */
/* predictor                                                                */

/* ref: src/libbjxa.c:525-531 and bjxa.5.rst:123-129 (K0, K1 scaled by 256) */
static const int16_t xao_gain[5][2] = {
	{   0,    0 },
	{ 240,    0 },
	{ 460, -208 },
	{ 392, -220 },
	{ 488, -240 },
};

/*
 * ref: src/libbjxa.c:533-578.
 *   factor = profile >> 4 (>= 5 is a protocol error, checked before any
 *   sample is touched, :550), range = profile & 15;
 *   ranged = (int16)(code >> range)            arithmetic shift, :558
 *   gain   = prev0*k0 + prev1*k1               int32, :559
 *   sample = ranged + gain / 256               C division: toward zero, :560
 *   clamp to int16, store, shift the state     :563-571
 */
int
xao_predict(int16_t *samples, unsigned stride, uint8_t profile,
    int16_t prev[2])
{
	unsigned factor = profile >> 4;
	unsigned range = profile & 15u;
	int32_t k0, k1;
	unsigned n;

	if (factor >= 5)
		return (-1);
	k0 = xao_gain[factor][0];
	k1 = xao_gain[factor][1];

	for (n = 0; n < XAO_BLOCK_SAMPLES; n++) {
		int16_t *s = samples + n * stride;
		int32_t code = *s;
		int32_t ranged, gain, q, v;

		/* arithmetic right shift written out so it does not depend on
		 * the compiler's treatment of negative operands */
		if (code >= 0)
			ranged = code >> range;
		else
			ranged = ~((~code) >> range);
		gain = (int32_t)prev[0] * k0 + (int32_t)prev[1] * k1;
		/* truncating division, again written out */
		q = gain >= 0 ? gain >> 8 : -((-gain) >> 8);
		v = ranged + q;
		if (v < -32768)
			v = -32768;
		if (v > 32767)
			v = 32767;
		*s = (int16_t)v;
		prev[1] = prev[0];
		prev[0] = (int16_t)v;
	}
	return (0);
}

/* ---------------------------------------------------------------------- */
/* block loops                                                              */

/* ref: src/libbjxa.c:602-661 */
long
xao_decode_blocks(unsigned bits, unsigned channels, int16_t prev[2][2],
    const uint8_t *xa, uint32_t blocks, int16_t *pcm, uint32_t *pcm_left,
    int *bad)
{
	const unsigned bsize = 4 * bits + 1;		/* :431 */
	const uint32_t full = 64u * channels;		/* :593 */
	int16_t frame[2 * XAO_BLOCK_SAMPLES];
	uint8_t *out = (uint8_t *)pcm;
	long done = 0;
	unsigned c;

	*bad = 0;
	while (blocks > 0 && *pcm_left > 0) {
		uint32_t take = *pcm_left < full ? *pcm_left : full; /* :622-624 */

		for (c = 0; c < channels; c++) {
			uint8_t profile = xao_inflate(bits, frame + c,
			    channels, xa);			/* :633,:640 */
			if (xao_predict(frame + c, channels, profile,
			    prev[c]) < 0) {			/* :634,:642 */
				*bad = 1;
				return (done);
			}
			xa += bsize;
		}
		memcpy(out, frame, take);			/* :648 */
		out += take;
		*pcm_left -= take;				/* :654 */
		blocks--;					/* :655 */
		done++;
	}
	return (done);
}

/* ref: src/libbjxa.c:665-691 (gather, zero-pad, profile 0), :759-819 */
long
xao_encode_blocks(unsigned bits, unsigned channels, const int16_t *pcm,
    uint32_t pcm_bytes, uint8_t *xa)
{
	const unsigned bsize = 4 * bits + 1;
	uint32_t frames = pcm_bytes / (2u * channels);
	long done = 0;
	unsigned c, n;

	while (frames > 0) {
		uint32_t take = frames < XAO_BLOCK_SAMPLES ? frames :
		    XAO_BLOCK_SAMPLES;

		for (c = 0; c < channels; c++) {
			int16_t one[XAO_BLOCK_SAMPLES];

			for (n = 0; n < take; n++)		/* :680-684 */
				one[n] = pcm[n * channels + c];
			for (; n < XAO_BLOCK_SAMPLES; n++)	/* :686-690 */
				one[n] = 0;
			xa[0] = 0;				/* :679,:793 */
			xao_deflate(bits, xa + 1, one);		/* :794 */
			xa += bsize;
		}
		pcm += take * channels;
		frames -= take;
		done++;
	}
	return (done);
}

/* ---------------------------------------------------------------------- */
/* searching encoder -- an EXTENSION, see bjxa_oracle.h                      */

static long
floor_div(long a, long b)				/* b > 0 */
{
	return (a >= 0 ? a / b : -((-a + b - 1) / b));
}

/* MSB-first bit stream of 32 codes of `bits` bits each (bjxa.5.rst) */
static void
pack_codes(unsigned bits, uint8_t *payload, const int codes[XAO_BLOCK_SAMPLES])
{
	unsigned n, k, pos = 0;

	memset(payload, 0, 4 * bits);
	for (n = 0; n < XAO_BLOCK_SAMPLES; n++)
		for (k = 0; k < bits; k++, pos++)
			if (((unsigned)codes[n] >> (bits - 1 - k)) & 1u)
				payload[pos >> 3] |= (uint8_t)(0x80u >> (pos & 7));
}

long
xao_encode_search_blocks(unsigned bits, unsigned channels, int16_t prev[2][2],
    const int16_t *pcm, uint32_t pcm_bytes, uint8_t *xa)
{
	const unsigned bsize = 4 * bits + 1;
	const int lo = -(1 << (bits - 1)), hi = (1 << (bits - 1)) - 1;
	uint32_t frames = pcm_bytes / (2u * channels);
	long done = 0;
	unsigned c, n, f, r;

	while (frames > 0) {
		uint32_t take = frames < XAO_BLOCK_SAMPLES ? frames :
		    XAO_BLOCK_SAMPLES;

		for (c = 0; c < channels; c++) {
			int16_t one[XAO_BLOCK_SAMPLES], back[XAO_BLOCK_SAMPLES];
			int16_t best_rec[XAO_BLOCK_SAMPLES], st[2];
			int best_codes[XAO_BLOCK_SAMPLES];
			uint64_t best_err = UINT64_MAX;
			unsigned best_profile = 0;

			for (n = 0; n < take; n++)
				one[n] = pcm[n * channels + c];
			for (; n < XAO_BLOCK_SAMPLES; n++)
				one[n] = 0;

			/* ascending profile byte, strict "<": ties keep the lowest */
			for (f = 0; f < 5; f++) {
				for (r = 0; r <= 16 - bits; r++) {
					const long step = 1L << (16 - bits - r);
					long q0 = prev[c][0], q1 = prev[c][1];
					int codes[XAO_BLOCK_SAMPLES];
					int16_t rec[XAO_BLOCK_SAMPLES];
					uint64_t err = 0;

					for (n = 0; n < XAO_BLOCK_SAMPLES; n++) {
						long pred = (q0 * xao_gain[f][0] +
						    q1 * xao_gain[f][1]) / 256;	/* C: toward 0 */
						long code = floor_div(one[n] - pred + step / 2, step);
						long s, e;

						if (code < lo)
							code = lo;
						if (code > hi)
							code = hi;
						s = code * step + pred;
						if (s < INT16_MIN)
							s = INT16_MIN;
						if (s > INT16_MAX)
							s = INT16_MAX;
						e = one[n] - s;
						err += (uint64_t)(e * e);
						codes[n] = (int)code;
						rec[n] = (int16_t)s;
						q1 = q0;
						q0 = s;
					}
					if (err < best_err) {
						best_err = err;
						best_profile = f << 4 | r;
						memcpy(best_codes, codes, sizeof codes);
						memcpy(best_rec, rec, sizeof rec);
					}
				}
			}

			xa[0] = (uint8_t)best_profile;
			pack_codes(bits, xa + 1, best_codes);

			/* what the (pinned) decoder makes of this block must be what
			 * the search thought it would be */
			st[0] = prev[c][0];
			st[1] = prev[c][1];
			if (xao_inflate(bits, back, 1, xa) != best_profile ||
			    xao_predict(back, 1, xa[0], st) != 0 ||
			    memcmp(back, best_rec, sizeof back) != 0)
				return (-1);
			prev[c][0] = st[0];
			prev[c][1] = st[1];
			xa += bsize;
		}
		pcm += take * channels;
		frames -= take;
		done++;
	}
	return (done);
}

/* ---------------------------------------------------------------------- */
/* containers                                                               */

/* ref: src/libbjxa.c:395-453 */
int
xao_parse_xa_header(xao_stream_t *st, const uint8_t hdr[32])
{
	xao_stream_t t;
	uint32_t bsize, nblocks, max_samples;

	memset(&t, 0, sizeof t);
	if (memcmp(hdr, "KWD1", 4) != 0)			/* :410 */
		return (-1);
	t.data_len = get_le(hdr + 4, 4);
	t.samples = get_le(hdr + 8, 4);
	t.rate = (uint16_t)get_le(hdr + 12, 2);
	t.bits = hdr[14];
	t.channels = hdr[15];
	/* hdr+16: nLoopPtr, ignored (:416,:446) */
	t.prev[0][0] = (int16_t)get_le(hdr + 20, 2);		/* :417-420 */
	t.prev[0][1] = (int16_t)get_le(hdr + 22, 2);
	t.prev[1][0] = (int16_t)get_le(hdr + 24, 2);
	t.prev[1][1] = (int16_t)get_le(hdr + 26, 2);
	/* hdr+28: pad, ignored */

	if (t.data_len == 0 || t.samples == 0 || t.rate == 0)	/* :425-427 */
		return (-1);
	if (t.bits != 4 && t.bits != 6 && t.bits != 8)		/* :428 */
		return (-1);
	if (t.channels != 1 && t.channels != 2)			/* :429 */
		return (-1);

	bsize = 4u * t.bits + 1u;				/* :431 */
	nblocks = t.data_len / bsize;
	/* uint32 arithmetic on purpose: the reference's product wraps */
	max_samples = (uint32_t)(32u * t.data_len) /
	    (uint32_t)(bsize * t.channels);			/* :433-434 */
	if (nblocks * bsize != t.data_len)			/* :435 */
		return (-1);
	if (max_samples < t.samples)				/* :436 */
		return (-1);
	if (max_samples - t.samples >= XAO_BLOCK_SAMPLES)	/* :437 */
		return (-1);
	*st = t;
	return (0);
}

/* ref: src/libbjxa.c:478-503 */
void
xao_write_xa_header(const xao_stream_t *st, uint8_t hdr[32])
{
	memset(hdr, 0, 32);
	memcpy(hdr, "KWD1", 4);
	put_le(hdr + 4, st->data_len, 4);
	put_le(hdr + 8, st->samples, 4);
	put_le(hdr + 12, st->rate, 2);
	hdr[14] = st->bits;
	hdr[15] = st->channels;
}

/* ref: src/libbjxa.c:898-927 */
void
xao_write_riff_header(const xao_stream_t *st, uint8_t hdr[44])
{
	uint32_t pcm = st->samples * st->channels * 2u;		/* :588 */

	memcpy(hdr, "RIFF", 4);
	put_le(hdr + 4, 36u + pcm, 4);				/* :911 */
	memcpy(hdr + 8, "WAVEfmt ", 8);
	put_le(hdr + 16, 16, 4);
	put_le(hdr + 20, 1, 2);
	put_le(hdr + 22, st->channels, 2);
	put_le(hdr + 24, st->rate, 4);
	put_le(hdr + 28, (uint32_t)st->rate * 2u * st->channels, 4); /* :917 */
	put_le(hdr + 32, 2u * st->channels, 2);
	put_le(hdr + 34, 16, 2);
	memcpy(hdr + 36, "data", 4);
	put_le(hdr + 40, pcm, 4);
}

/* ref: src/libbjxa.c:826-873 */
int
xao_parse_riff_header(xao_stream_t *st, uint32_t *pcm_bytes,
    const uint8_t hdr[44])
{
	uint32_t riff, fmtlen, rate, bps, data;
	uint16_t tag, chan, align, bits;

	if (memcmp(hdr, "RIFF", 4) != 0 || memcmp(hdr + 8, "WAVEfmt ", 8) != 0 ||
	    memcmp(hdr + 36, "data", 4) != 0)
		return (-1);
	riff = get_le(hdr + 4, 4);
	fmtlen = get_le(hdr + 16, 4);
	tag = (uint16_t)get_le(hdr + 20, 2);
	chan = (uint16_t)get_le(hdr + 22, 2);
	rate = get_le(hdr + 24, 4);
	bps = get_le(hdr + 28, 4);
	align = (uint16_t)get_le(hdr + 32, 2);
	bits = (uint16_t)get_le(hdr + 34, 2);
	data = get_le(hdr + 40, 4);

	if (riff < 36u + data)					/* :855 */
		return (-1);
	if (fmtlen != 16 || tag != 1)				/* :856-857 */
		return (-1);
	if (chan != 1 && chan != 2)				/* :858 */
		return (-1);
	if (rate == 0 || rate >= 65535u)			/* :859 */
		return (-1);
	if (align != chan * 2u)					/* :860 */
		return (-1);
	if (bps != rate * align)				/* :861 */
		return (-1);
	if (data % align != 0 || bits != 16)			/* :862-863 */
		return (-1);

	memset(st, 0, sizeof *st);
	st->channels = (uint8_t)chan;
	st->rate = (uint16_t)rate;
	*pcm_bytes = data;
	return (0);
}

/* ref: src/libbjxa.c:693-735 */
int
xao_encode_geometry(xao_stream_t *st, uint32_t pcm_bytes, unsigned bits,
    unsigned channels, unsigned rate)
{
	uint32_t bsize, nblocks;

	memset(st, 0, sizeof *st);
	if (bits != 4 && bits != 6 && bits != 8)		/* :701 */
		return (-1);
	if (channels != 1 && channels != 2)			/* :706 */
		return (-1);
	st->bits = (uint8_t)bits;
	st->channels = (uint8_t)channels;
	st->rate = (uint16_t)rate;
	st->samples = pcm_bytes / (channels * 2u);		/* :708 */
	if (st->samples == 0 || st->rate == 0)			/* :710-711 */
		return (-1);
	if (pcm_bytes % st->samples != 0)			/* :712 */
		return (-1);
	bsize = 4u * bits + 1u;					/* :721 */
	nblocks = st->samples / XAO_BLOCK_SAMPLES;		/* :725 */
	if (st->samples % XAO_BLOCK_SAMPLES != 0)		/* :727-730 */
		nblocks++;
	st->data_len = nblocks * bsize * channels;
	return (0);
}

/* ---------------------------------------------------------------------- */
/* whole files                                                              */

long
xao_xa_to_wav(const uint8_t *xa, size_t xa_len, uint8_t *wav, size_t wav_cap)
{
	xao_stream_t st;
	uint32_t pcm_bytes, left, nblocks;
	int bad;

	if (xa_len < 32 || xao_parse_xa_header(&st, xa) < 0)
		return (-1);
	if (xa_len - 32 < st.data_len)
		return (-1);
	pcm_bytes = st.samples * st.channels * 2u;
	if (wav_cap < 44u + (size_t)pcm_bytes)
		return (-1);
	xao_write_riff_header(&st, wav);
	nblocks = st.data_len / ((4u * st.bits + 1u) * st.channels);
	left = pcm_bytes;
	/* the PCM area of a WAV image is only 4-byte aligned; decode in place
	 * is fine because int16 stores need 2-byte alignment */
	xao_decode_blocks(st.bits, st.channels, st.prev, xa + 32, nblocks,
	    (int16_t *)(void *)(wav + 44), &left, &bad);
	if (bad || left != 0)
		return (-1);
	return (44 + (long)pcm_bytes);
}

long
xao_wav_to_xa(const uint8_t *wav, size_t wav_len, unsigned bits, uint8_t *xa,
    size_t xa_cap)
{
	xao_stream_t st, geo;
	uint32_t pcm_bytes;

	if (wav_len < 44 || xao_parse_riff_header(&st, &pcm_bytes, wav) < 0)
		return (-1);
	if (wav_len - 44 < pcm_bytes)
		return (-1);
	if (xao_encode_geometry(&geo, pcm_bytes, bits, st.channels, st.rate) < 0)
		return (-1);
	if (xa_cap < 32u + (size_t)geo.data_len)
		return (-1);
	xao_write_xa_header(&geo, xa);
	xao_encode_blocks(bits, geo.channels,
	    (const int16_t *)(const void *)(wav + 44), pcm_bytes, xa + 32);
	return (32 + (long)geo.data_len);
}
