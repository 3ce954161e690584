/*
 * bjxa_host.c -- the host side of libbjxa for the B200 backend, in C.
 *
 * Implements the 19 public functions of the reference API (include/bjxa.h,
 * i.e. /root/reference/src/bjxa.h:36-65) plus the host-buffer batch calls and
 * the codec<->descriptor glue of include/bjxa_batch.h.  Everything cold --
 * codec objects, the 32-byte XA header, the 44-byte RIFF header, the errno
 * contract -- is done here on the CPU exactly as the reference does it
 * (citations per function).  The block transform itself is NOT done here:
 * bjxa_decode() / bjxa_encode() stage their buffers to the GPU and run the
 * batch-of-one through the same kernels as bjxa_plan_run(); with no CUDA
 * device they fail with ENODEV.  There is deliberately no CPU implementation
 * of the transform in this library.
 */
#define _POSIX_C_SOURCE 200809L

#include <errno.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/bjxa.h"
#include "../../include/bjxa_batch.h"
#include "bjxa_internal.h"

#define XA_SAMPLES_PER_BLOCK	32u

/* Same tags as the reference so that a junk pointer is told apart from a
 * codec the same way (src/libbjxa.c:219,232). */
#define DECODER_TAG	0x234ec0c2u
#define ENCODER_TAG	0xac12f1ddu

struct bjxa_decoder {
	uint32_t	tag;
	uint32_t	xa_bytes;	/* nDataLen */
	uint32_t	samples;	/* nSamples, per channel */
	uint16_t	rate;
	uint8_t		bits;
	uint8_t		channels;
	uint8_t		block_bytes;	/* 4*bits+1, 0 = header not parsed yet */
	int16_t		prev[2][2];	/* [channel][n-1, n-2] */
	bjxa_format_t	left;		/* what remains to be decoded */
};

struct bjxa_encoder {
	uint32_t	tag;
	uint32_t	xa_bytes;
	uint32_t	samples;
	uint16_t	rate;
	uint8_t		bits;
	uint8_t		channels;
	uint8_t		block_bytes;
	bjxa_format_t	left;		/* what remains to be encoded */
};

/* ---- errno contract (src/libbjxa.c:53-97, bjxa.3.rst.in:205-276) -------- */

#define FAIL(e)		do { errno = (e); return (-1); } while (0)
#define NEED_PTR(p)	do { if ((p) == NULL) FAIL(EFAULT); } while (0)
#define NEED_OBJ(o, t)	do { NEED_PTR(o); if ((o)->tag != (t)) FAIL(EINVAL); } while (0)

/* ---- little-endian fields (src/libbjxa.c:99-166) ------------------------- */

static uint32_t
rd_le(const uint8_t *p, unsigned n)
{
	uint32_t v = 0;
	unsigned i;

	for (i = 0; i < n; i++)
		v |= (uint32_t)p[i] << (8 * i);
	return (v);
}

static void
wr_le(uint8_t *p, uint32_t v, unsigned n)
{
	unsigned i;

	for (i = 0; i < n; i++)
		p[i] = (uint8_t)(v >> (8 * i));
}

/* ---- codec lifetime (src/libbjxa.c:246-282) ------------------------------ */

bjxa_decoder_t *
bjxa_decoder(void)
{
	bjxa_decoder_t *dec;

	errno = 0;
	dec = calloc(1, sizeof *dec);
	if (dec != NULL)
		dec->tag = DECODER_TAG;
	return (dec);
}

int
bjxa_free_decoder(bjxa_decoder_t **decp)
{
	bjxa_decoder_t *dec;

	NEED_PTR(decp);
	dec = *decp;
	NEED_OBJ(dec, DECODER_TAG);
	*decp = NULL;
	memset(dec, 0, sizeof *dec);
	free(dec);
	return (0);
}

bjxa_encoder_t *
bjxa_encoder(void)
{
	bjxa_encoder_t *enc;

	errno = 0;
	enc = calloc(1, sizeof *enc);
	if (enc != NULL)
		enc->tag = ENCODER_TAG;
	return (enc);
}

int
bjxa_free_encoder(bjxa_encoder_t **encp)
{
	bjxa_encoder_t *enc;

	NEED_PTR(encp);
	enc = *encp;
	NEED_OBJ(enc, ENCODER_TAG);
	*encp = NULL;
	memset(enc, 0, sizeof *enc);
	free(enc);
	return (0);
}

/* ---- XA header (src/libbjxa.c:395-521) ----------------------------------- */

static void
decoder_geometry(const bjxa_decoder_t *dec, bjxa_format_t *fmt)
{
	/* src/libbjxa.c:588-595 */
	fmt->data_len_pcm = dec->samples * dec->channels * 2u;
	fmt->samples_rate = dec->rate;
	fmt->sample_bits = 16;
	fmt->channels = dec->channels;
	fmt->block_size_xa = (uint8_t)(dec->block_bytes * dec->channels);
	fmt->block_size_pcm = (uint8_t)(XA_SAMPLES_PER_BLOCK * dec->channels * 2u);
	fmt->blocks = dec->xa_bytes / fmt->block_size_xa;
}

ssize_t
bjxa_parse_header(bjxa_decoder_t *dec, const void *src, size_t len)
{
	const uint8_t *h = src;
	bjxa_decoder_t t;
	uint32_t nblocks, max_samples;

	NEED_OBJ(dec, DECODER_TAG);
	NEED_PTR(src);
	if (len < BJXA_HEADER_SIZE_XA)
		FAIL(ENOBUFS);

	/* layout: bjxa.5.rst:63-105; field reads: libbjxa.c:410-421 */
	memset(&t, 0, sizeof t);
	t.tag = DECODER_TAG;
	if (memcmp(h, "KWD1", 4) != 0)
		FAIL(EPROTO);
	t.xa_bytes = rd_le(h + 4, 4);
	t.samples = rd_le(h + 8, 4);
	t.rate = (uint16_t)rd_le(h + 12, 2);
	t.bits = h[14];
	t.channels = h[15];
	/* h+16 nLoopPtr and h+28 pad are read and ignored (libbjxa.c:446-447) */
	t.prev[0][0] = (int16_t)rd_le(h + 20, 2);
	t.prev[0][1] = (int16_t)rd_le(h + 22, 2);
	t.prev[1][0] = (int16_t)rd_le(h + 24, 2);
	t.prev[1][1] = (int16_t)rd_le(h + 26, 2);

	/* validation, in the reference's order (libbjxa.c:425-437) */
	if (t.xa_bytes == 0 || t.samples == 0 || t.rate == 0)
		FAIL(EPROTO);
	if (t.bits != 4 && t.bits != 6 && t.bits != 8)
		FAIL(EPROTO);
	if (t.channels != 1 && t.channels != 2)
		FAIL(EPROTO);
	t.block_bytes = (uint8_t)(t.bits * 4 + 1);
	nblocks = t.xa_bytes / t.block_bytes;
	/* the product deliberately wraps in 32 bits, as it does upstream */
	max_samples = (uint32_t)(XA_SAMPLES_PER_BLOCK * t.xa_bytes) /
	    (uint32_t)(t.block_bytes * t.channels);
	if (nblocks * t.block_bytes != t.xa_bytes)
		FAIL(EPROTO);
	if (max_samples < t.samples)
		FAIL(EPROTO);
	if (max_samples - t.samples >= XA_SAMPLES_PER_BLOCK)
		FAIL(EPROTO);

	decoder_geometry(&t, &t.left);
	/* a stereo file with an odd number of blocks: the reference stops on it
	 * (assert in bjxa_decode_format, libbjxa.c:596, reached from :449); a
	 * library that serves batches reports the malformed header instead */
	if (t.left.blocks * t.left.block_size_xa != t.xa_bytes)
		FAIL(EPROTO);
	*dec = t;		/* all or nothing (libbjxa.c:409,451) */
	return (BJXA_HEADER_SIZE_XA);
}

ssize_t
bjxa_fread_header(bjxa_decoder_t *dec, FILE *file)
{
	uint8_t buf[BJXA_HEADER_SIZE_XA];

	NEED_OBJ(dec, DECODER_TAG);
	NEED_PTR(file);
	if (fread(buf, sizeof buf, 1, file) != 1) {
		if (feof(file))
			errno = EIO;		/* libbjxa.c:464-468 */
		return (-1);
	}
	return (bjxa_parse_header(dec, buf, sizeof buf));
}

int
bjxa_decode_format(bjxa_decoder_t *dec, bjxa_format_t *fmt)
{
	NEED_OBJ(dec, DECODER_TAG);
	NEED_PTR(fmt);
	if (dec->block_bytes == 0)
		FAIL(EINVAL);
	decoder_geometry(dec, fmt);
	return (0);
}

ssize_t
bjxa_dump_header(bjxa_encoder_t *enc, void *dst, size_t len)
{
	uint8_t *h = dst;

	NEED_OBJ(enc, ENCODER_TAG);
	NEED_PTR(dst);
	if (len < BJXA_HEADER_SIZE_XA)
		FAIL(ENOBUFS);
	if (enc->xa_bytes == 0)
		FAIL(EINVAL);
	/* loop pointer, predictor state and pad are all zero (libbjxa.c:495-500) */
	memset(h, 0, BJXA_HEADER_SIZE_XA);
	memcpy(h, "KWD1", 4);
	wr_le(h + 4, enc->xa_bytes, 4);
	wr_le(h + 8, enc->samples, 4);
	wr_le(h + 12, enc->rate, 2);
	h[14] = enc->bits;
	h[15] = enc->channels;
	return (BJXA_HEADER_SIZE_XA);
}

ssize_t
bjxa_fwrite_header(bjxa_encoder_t *enc, FILE *file)
{
	uint8_t buf[BJXA_HEADER_SIZE_XA];

	NEED_OBJ(enc, ENCODER_TAG);
	NEED_PTR(file);
	if (bjxa_dump_header(enc, buf, sizeof buf) < 0)
		return (-1);
	if (fwrite(buf, sizeof buf, 1, file) != 1)
		return (-1);
	return (BJXA_HEADER_SIZE_XA);
}

/* ---- RIFF/WAVE header (src/libbjxa.c:821-945) ----------------------------- */

ssize_t
bjxa_parse_riff_header(bjxa_format_t *fmt, const void *src, size_t len)
{
	const uint8_t *h = src;
	uint32_t riff_len, fmt_len, rate, byte_rate, data_len;
	uint16_t tag, chans, align, width;
	bjxa_format_t t;

	NEED_PTR(fmt);
	NEED_PTR(src);
	if (len < BJXA_HEADER_SIZE_RIFF)
		FAIL(ENOBUFS);

	if (memcmp(h, "RIFF", 4) != 0)
		FAIL(EPROTO);
	riff_len = rd_le(h + 4, 4);
	if (memcmp(h + 8, "WAVEfmt ", 8) != 0)
		FAIL(EPROTO);
	fmt_len = rd_le(h + 16, 4);
	tag = (uint16_t)rd_le(h + 20, 2);
	chans = (uint16_t)rd_le(h + 22, 2);
	rate = rd_le(h + 24, 4);
	byte_rate = rd_le(h + 28, 4);
	align = (uint16_t)rd_le(h + 32, 2);
	width = (uint16_t)rd_le(h + 34, 2);
	if (memcmp(h + 36, "data", 4) != 0)
		FAIL(EPROTO);
	data_len = rd_le(h + 40, 4);

	/* libbjxa.c:855-863, same order */
	if (riff_len < BJXA_HEADER_SIZE_RIFF - 8 + data_len)
		FAIL(EPROTO);
	if (fmt_len != 16 || tag != 1)
		FAIL(EPROTO);
	if (chans != 1 && chans != 2)
		FAIL(EPROTO);
	if (rate == 0 || rate >= UINT16_MAX)
		FAIL(EPROTO);
	if (align != chans * 2u)
		FAIL(EPROTO);
	if (byte_rate != rate * align)
		FAIL(EPROTO);
	if (data_len % align != 0)
		FAIL(EPROTO);
	if (width != 16)
		FAIL(EPROTO);

	memset(&t, 0, sizeof t);
	t.data_len_pcm = data_len;
	t.samples_rate = (uint16_t)rate;
	t.sample_bits = 16;
	t.channels = (uint8_t)chans;
	*fmt = t;
	return (BJXA_HEADER_SIZE_RIFF);
}

ssize_t
bjxa_fread_riff_header(bjxa_format_t *fmt, FILE *file)
{
	uint8_t buf[BJXA_HEADER_SIZE_RIFF];

	NEED_PTR(fmt);
	NEED_PTR(file);
	if (fread(buf, sizeof buf, 1, file) != 1) {
		if (feof(file))
			errno = EIO;
		return (-1);
	}
	return (bjxa_parse_riff_header(fmt, buf, sizeof buf));
}

ssize_t
bjxa_dump_riff_header(bjxa_decoder_t *dec, void *dst, size_t len)
{
	bjxa_format_t f;
	uint8_t *h = dst;

	NEED_OBJ(dec, DECODER_TAG);
	NEED_PTR(dst);
	if (len < BJXA_HEADER_SIZE_RIFF)
		FAIL(ENOBUFS);
	if (bjxa_decode_format(dec, &f) < 0)
		return (-1);

	/* libbjxa.c:909-922 */
	memcpy(h, "RIFF", 4);
	wr_le(h + 4, BJXA_HEADER_SIZE_RIFF - 8 + f.data_len_pcm, 4);
	memcpy(h + 8, "WAVEfmt ", 8);
	wr_le(h + 16, 16, 4);
	wr_le(h + 20, 1, 2);
	wr_le(h + 22, f.channels, 2);
	wr_le(h + 24, f.samples_rate, 4);
	wr_le(h + 28, (uint32_t)f.samples_rate * f.block_size_pcm /
	    XA_SAMPLES_PER_BLOCK, 4);
	wr_le(h + 32, (uint32_t)f.channels * f.sample_bits / 8u, 2);
	wr_le(h + 34, f.sample_bits, 2);
	memcpy(h + 36, "data", 4);
	wr_le(h + 40, f.data_len_pcm, 4);
	return (BJXA_HEADER_SIZE_RIFF);
}

ssize_t
bjxa_fwrite_riff_header(bjxa_decoder_t *dec, FILE *file)
{
	uint8_t buf[BJXA_HEADER_SIZE_RIFF];

	NEED_OBJ(dec, DECODER_TAG);
	NEED_PTR(file);
	if (bjxa_dump_riff_header(dec, buf, sizeof buf) < 0)
		return (-1);
	if (fwrite(buf, sizeof buf, 1, file) != 1)
		return (-1);
	return (BJXA_HEADER_SIZE_RIFF);
}

/* ---- PCM serialisation (src/libbjxa.c:947-996) ---------------------------- */

int
bjxa_dump_pcm(void *dst, const int16_t *src, size_t len)
{
	uint8_t *out = dst;
	size_t i;

	NEED_PTR(dst);
	NEED_PTR(src);
	if (len == 0 || (len & 1) != 0)
		FAIL(ENOBUFS);
	for (i = 0; i < len / 2; i++)
		wr_le(out + 2 * i, (uint16_t)src[i], 2);
	return (0);
}

int
bjxa_fwrite_pcm(const int16_t *src, size_t len, FILE *file)
{
	uint8_t chunk[64];

	NEED_PTR(src);
	NEED_PTR(file);
	if (len == 0 || (len & 1) != 0)
		FAIL(ENOBUFS);
	while (len > 0) {
		size_t n = len < sizeof chunk ? len : sizeof chunk;

		(void)bjxa_dump_pcm(chunk, src, n);
		if (fwrite(chunk, n, 1, file) != 1)
			return (-1);
		src += n / 2;
		len -= n;
	}
	return (0);
}

/* ---- encoder set-up (src/libbjxa.c:693-757) ------------------------------- */

int
bjxa_encode_init(bjxa_encoder_t *enc, bjxa_format_t *fmt, uint8_t bits)
{
	bjxa_encoder_t t;

	NEED_OBJ(enc, ENCODER_TAG);
	NEED_PTR(fmt);
	if (fmt->sample_bits != 16)
		FAIL(EINVAL);
	if (bits != 4 && bits != 6 && bits != 8)
		FAIL(EINVAL);

	memset(&t, 0, sizeof t);
	t.tag = ENCODER_TAG;
	t.bits = bits;
	t.channels = fmt->channels;
	if (t.channels != 1 && t.channels != 2)
		FAIL(EPROTO);
	t.samples = fmt->data_len_pcm / (t.channels * 2u);
	t.rate = fmt->samples_rate;
	if (t.samples == 0 || t.rate == 0)
		FAIL(EPROTO);
	if (fmt->data_len_pcm % t.samples != 0)		/* libbjxa.c:712 */
		FAIL(EPROTO);

	t.block_bytes = (uint8_t)(bits * 4 + 1);
	/* written back to the caller (libbjxa.c:722-730) */
	fmt->block_size_xa = (uint8_t)(t.block_bytes * t.channels);
	fmt->block_size_pcm = (uint8_t)(XA_SAMPLES_PER_BLOCK * t.channels * 2u);
	fmt->blocks = t.samples / XA_SAMPLES_PER_BLOCK;
	if (t.samples % XA_SAMPLES_PER_BLOCK != 0)
		fmt->blocks++;
	t.xa_bytes = fmt->blocks * fmt->block_size_xa;
	t.left = *fmt;
	*enc = t;
	return (0);
}

int
bjxa_encode_format(bjxa_encoder_t *enc, bjxa_format_t *fmt)
{
	NEED_OBJ(enc, ENCODER_TAG);
	NEED_PTR(fmt);
	if (enc->block_bytes == 0)
		FAIL(EINVAL);
	/* libbjxa.c:745-752: note sample_bits is the XA width here */
	fmt->data_len_pcm = enc->samples * enc->channels * 2u;
	fmt->samples_rate = enc->rate;
	fmt->sample_bits = enc->bits;
	fmt->channels = enc->channels;
	fmt->block_size_xa = (uint8_t)(enc->block_bytes * enc->channels);
	fmt->block_size_pcm = (uint8_t)(XA_SAMPLES_PER_BLOCK * enc->channels * 2u);
	fmt->blocks = enc->xa_bytes / fmt->block_size_xa;
	return (0);
}

/* ---- how many blocks one call covers -------------------------------------- */

/*
 * The reference loops "while blocks remain and both buffers have room"
 * (libbjxa.c:629-630, 787-788); the PCM side shrinks to the stream's
 * remaining bytes on the last block (:622-624,656-657).  Same count, closed
 * form: `pcm_room` is the caller's PCM buffer, `xa_room` the XA buffer.
 */
static uint32_t
blocks_this_call(const bjxa_format_t *left, size_t pcm_room, size_t xa_room,
    uint32_t *pcm_bytes)
{
	const uint32_t full = left->block_size_pcm;
	uint32_t whole = left->data_len_pcm / full;	/* full-size blocks left */
	uint32_t rest = left->data_len_pcm % full;	/* short last block */
	uint64_t by_pcm, by_xa, n;

	by_pcm = pcm_room / full;
	if (by_pcm >= whole) {
		by_pcm = whole;
		if (rest != 0 && pcm_room - (size_t)whole * full >= rest)
			by_pcm++;
	}
	by_xa = xa_room / left->block_size_xa;
	n = left->blocks;
	if (by_pcm < n)
		n = by_pcm;
	if (by_xa < n)
		n = by_xa;
	if (n > whole)
		*pcm_bytes = whole * full + rest;
	else
		*pcm_bytes = (uint32_t)n * full;
	return ((uint32_t)n);
}

/* ---- codec <-> descriptor -------------------------------------------------- */

int
bjxa_decoder_describe(bjxa_decoder_t *dec, bjxa_stream_desc_t *d)
{
	NEED_OBJ(dec, DECODER_TAG);
	NEED_PTR(d);
	if (dec->left.sample_bits != 16)
		FAIL(EINVAL);
	memset(d, 0, sizeof *d);
	d->blocks = dec->left.blocks;
	d->pcm_len = dec->left.data_len_pcm;
	memcpy(d->prev, dec->prev, sizeof d->prev);
	d->bits = dec->bits;
	d->channels = dec->channels;
	return (0);
}

int
bjxa_decoder_commit(bjxa_decoder_t *dec, const bjxa_stream_desc_t *d)
{
	uint32_t bytes;

	NEED_OBJ(dec, DECODER_TAG);
	NEED_PTR(d);
	if (dec->left.sample_bits != 16 || d->done > dec->left.blocks ||
	    d->done > d->blocks)
		FAIL(EINVAL);
	/* libbjxa.c:654-655 per block; 570-571 for the state */
	bytes = d->done == d->blocks ? d->pcm_len :
	    d->done * dec->left.block_size_pcm;
	if (bytes > dec->left.data_len_pcm)
		bytes = dec->left.data_len_pcm;
	dec->left.blocks -= d->done;
	dec->left.data_len_pcm -= bytes;
	memcpy(dec->prev, d->prev, sizeof dec->prev);
	return (0);
}

int
bjxa_encoder_describe(bjxa_encoder_t *enc, bjxa_stream_desc_t *d)
{
	NEED_OBJ(enc, ENCODER_TAG);
	NEED_PTR(d);
	if (enc->left.sample_bits != 16)
		FAIL(EINVAL);
	memset(d, 0, sizeof *d);
	d->blocks = enc->left.blocks;
	d->pcm_len = enc->left.data_len_pcm;
	d->bits = enc->bits;
	d->channels = enc->channels;
	return (0);
}

int
bjxa_encoder_commit(bjxa_encoder_t *enc, const bjxa_stream_desc_t *d)
{
	uint32_t bytes;

	NEED_OBJ(enc, ENCODER_TAG);
	NEED_PTR(d);
	if (enc->left.sample_bits != 16 || d->done > enc->left.blocks ||
	    d->done > d->blocks)
		FAIL(EINVAL);
	bytes = d->done == d->blocks ? d->pcm_len :
	    d->done * enc->left.block_size_pcm;
	if (bytes > enc->left.data_len_pcm)
		bytes = enc->left.data_len_pcm;
	enc->left.blocks -= d->done;		/* libbjxa.c:812-813 */
	enc->left.data_len_pcm -= bytes;
	return (0);
}

/* ---- per-thread staging for the host-buffer calls -------------------------- */

/*
 * Host batches are cut into chunks of streams and pushed through a software
 * pipeline PIPE_DEPTH deep, one CUDA stream per slot:
 *     upload(c) -> kernels(c) -> [results(c)] -> download(c)
 * so that, with pinned caller buffers, the upload of one chunk overlaps the
 * download of another (PCIe is full duplex) and both overlap the kernels.
 * The results of a chunk are fetched BEFORE its download is queued because a
 * stream that met a bad profile must only deliver the blocks in front of it
 * (libbjxa.c:634-648 leaves the rest of dst untouched); that happens as soon as
 * the next chunk has been queued, so the downloads of chunk c run under the
 * upload and the kernels of chunk c+1.
 */
#define PIPE_DEPTH	3
#define CHUNK_BYTES	((uint64_t)96 << 20)	/* PCM bytes per chunk, about */

/*
 * What a thread keeps from one host-buffer call to the next: per pipeline slot a
 * plan, a CUDA stream and two device arenas of chunk size (so device memory goes
 * with the pipeline's depth, not with the size of the batch).  All of it lives
 * on ONE device: a thread that has since selected another one (bjxa_gpu_select)
 * gets a fresh cache there.  Released by bjxa_thread_release(), and by itself
 * when the thread exits.
 */
struct stage {
	int		 device;	/* + 1, so that zero-initialised = nothing yet */
	bjxa_plan_t	*plan[PIPE_DEPTH];
	void		*stream[PIPE_DEPTH];
	void		*d_xa[PIPE_DEPTH], *d_pcm[PIPE_DEPTH];
	size_t		 cap_xa[PIPE_DEPTH], cap_pcm[PIPE_DEPTH];
};

static __thread struct stage tls_stage;

static void
stage_release(struct stage *sg)
{
	int k, cur;

	if (sg->device == 0)
		return;
	cur = bjxa_gpu_current();
	if (cur >= 0 && cur != sg->device - 1)
		(void)bjxa_gpu_select(sg->device - 1);
	for (k = 0; k < PIPE_DEPTH; k++) {
		if (sg->plan[k] != NULL)
			(void)bjxa_plan_free(&sg->plan[k]);
		if (sg->stream[k] != NULL)
			(void)bjxa_gpu_stream_destroy(sg->stream[k]);
		if (sg->d_xa[k] != NULL)
			(void)bjxa_gpu_free(sg->d_xa[k]);
		if (sg->d_pcm[k] != NULL)
			(void)bjxa_gpu_free(sg->d_pcm[k]);
	}
	if (cur >= 0 && cur != sg->device - 1)
		(void)bjxa_gpu_select(cur);
	memset(sg, 0, sizeof *sg);
}

/* thread exit: give the caches back */
static pthread_key_t cache_key;
static pthread_once_t cache_once = PTHREAD_ONCE_INIT;

static void
cache_at_exit(void *unused)
{
	(void)unused;
	bjxa_thread_release();
}

static void
cache_key_make(void)
{
	(void)pthread_key_create(&cache_key, cache_at_exit);
}

void
bjxa_thread_cache_used(void)
{
	(void)pthread_once(&cache_once, cache_key_make);
	if (pthread_getspecific(cache_key) == NULL)
		(void)pthread_setspecific(cache_key, &tls_stage);
}

void
bjxa_thread_release(void)
{
	int e = errno;

	stage_release(&tls_stage);
	bjxa_corpus_release();
	bjxa_small_release();
	errno = e;
}

static int
stage_reserve(void **p, size_t *cap, size_t need)
{
	if (need <= *cap)
		return (0);
	if (*p != NULL)
		(void)bjxa_gpu_free(*p);
	*p = NULL;
	*cap = 0;
	need += need / 4 + 4096;
	*p = bjxa_gpu_alloc(need);
	if (*p == NULL)
		return (-1);
	*cap = need;
	return (0);
}

static int
stage_plan(struct stage *sg, int slot, int kind, const bjxa_stream_desc_t *d,
    size_t n)
{
	if (sg->stream[slot] == NULL) {
		sg->stream[slot] = bjxa_gpu_stream_create();
		if (sg->stream[slot] == NULL)
			return (-1);
	}
	if (sg->plan[slot] == NULL) {
		sg->plan[slot] = bjxa_plan_create(kind, d, n);
		return (sg->plan[slot] == NULL ? -1 : 0);
	}
	return (bjxa_plan_reset(sg->plan[slot], kind, d, n));
}

#define ALIGN16(x)	(((x) + 15u) & ~(uint64_t)15u)

struct chunk {
	size_t		first, count;
	int		slot;
};

/* queue the upload and the kernels of one chunk on its slot's stream */
static int
chunk_start(struct stage *sg, int kind, const struct chunk *ck,
    bjxa_stream_desc_t *work, const void *const *srcs, const uint32_t *xa_bytes)
{
	const int slot = ck->slot;
	uint64_t xa_total = 0, pcm_total = 0;
	void *st;
	size_t i;

	/* the chunk's streams, back to back in the slot's two arenas */
	for (i = ck->first; i < ck->first + ck->count; i++) {
		if (work[i].blocks == 0)
			continue;
		work[i].xa_off = xa_total;
		work[i].pcm_off = pcm_total;
		xa_total += xa_bytes[i];
		pcm_total = ALIGN16(pcm_total + (uint64_t)work[i].blocks * 64u *
		    work[i].channels);
	}
	xa_total = ALIGN16(xa_total) + 16;
	pcm_total += 16;
	/* (growing an arena frees the old one, which waits for the slot's earlier
	 * download to have left it) */
	if (stage_reserve(&sg->d_xa[slot], &sg->cap_xa[slot], xa_total) < 0 ||
	    stage_reserve(&sg->d_pcm[slot], &sg->cap_pcm[slot], pcm_total) < 0)
		return (-1);
	if (stage_plan(sg, slot, kind, work + ck->first, ck->count) < 0)
		return (-1);
	st = sg->stream[slot];
	/* one copy per run of streams whose source buffers follow each other in
	 * host memory exactly as their places do in the device arena (callers
	 * that keep a batch in one allocation get one copy per chunk) */
	{
		const uint8_t *run_src = NULL;
		uint8_t *run_dst = NULL;
		size_t run_len = 0;

		for (i = ck->first; i < ck->first + ck->count; i++) {
			uint8_t *d;
			size_t len;

			if (work[i].blocks == 0)
				continue;
			if (kind == BJXA_PLAN_DECODE) {
				d = (uint8_t *)sg->d_xa[slot] + work[i].xa_off;
				len = xa_bytes[i];
			} else {
				d = (uint8_t *)sg->d_pcm[slot] + work[i].pcm_off;
				len = work[i].pcm_len;
			}
			if (run_len != 0 && (const uint8_t *)srcs[i] == run_src + run_len &&
			    d == run_dst + run_len) {
				run_len += len;
				continue;
			}
			if (run_len != 0 && bjxa_gpu_upload_async(run_dst, run_src, run_len, st) < 0)
				return (-1);
			run_src = srcs[i];
			run_dst = d;
			run_len = len;
		}
		if (run_len != 0 && bjxa_gpu_upload_async(run_dst, run_src, run_len, st) < 0)
			return (-1);
	}
	if (kind == BJXA_PLAN_DECODE)
		return (bjxa_plan_run(sg->plan[slot], sg->d_pcm[slot], sg->cap_pcm[slot],
		    sg->d_xa[slot], sg->cap_xa[slot], st));
	return (bjxa_plan_run(sg->plan[slot], sg->d_xa[slot], sg->cap_xa[slot],
	    sg->d_pcm[slot], sg->cap_pcm[slot], st));
}

/* wait for the chunk's kernels, take its results, queue its download */
static int
chunk_finish(struct stage *sg, int kind, const struct chunk *ck,
    bjxa_stream_desc_t *work, void *const *dsts)
{
	const int slot = ck->slot;
	void *st = sg->stream[slot];
	size_t i;

	if (bjxa_plan_fetch(sg->plan[slot], work + ck->first, ck->count) < 0)
		return (-1);
	/* downloads, again one copy per run of buffers that follow each other in
	 * host memory as their contents do on the device -- and never a byte
	 * more than a stream delivers */
	{
		uint8_t *run_dst = NULL;
		const uint8_t *run_src = NULL;
		size_t run_len = 0;

		for (i = ck->first; i < ck->first + ck->count; i++) {
			const uint8_t *d;
			size_t bytes;

			if (work[i].blocks == 0)
				continue;
			if (kind == BJXA_PLAN_DECODE) {
				bytes = work[i].done == work[i].blocks ? work[i].pcm_len :
				    (size_t)work[i].done * 64u * work[i].channels;
				d = (const uint8_t *)sg->d_pcm[slot] + work[i].pcm_off;
			} else {
				bytes = (size_t)work[i].done * (4u * work[i].bits + 1u) *
				    work[i].channels;
				d = (const uint8_t *)sg->d_xa[slot] + work[i].xa_off;
			}
			if (bytes == 0)
				continue;
			if (run_len != 0 && (uint8_t *)dsts[i] == run_dst + run_len &&
			    d == run_src + run_len) {
				run_len += bytes;
				continue;
			}
			if (run_len != 0 && bjxa_gpu_download_async(run_dst, run_src, run_len, st) < 0)
				return (-1);
			run_dst = dsts[i];
			run_src = d;
			run_len = bytes;
		}
		if (run_len != 0 && bjxa_gpu_download_async(run_dst, run_src, run_len, st) < 0)
			return (-1);
	}
	return (0);
}

/*
 * The common engine behind bjxa_decode, bjxa_encode and their batch forms.
 * work[i].blocks == 0 marks a stream that takes no part (already failed its
 * argument checks).  On return work[] holds the fetched results and every
 * delivered byte has landed in the caller's buffers.
 */
static int
run_host_batch(int kind, bjxa_stream_desc_t *work, void *const *dsts,
    const void *const *srcs, const uint32_t *xa_bytes, size_t n)
{
	struct stage *sg = &tls_stage;
	struct chunk ring[PIPE_DEPTH];
	uint64_t in_chunk = 0;
	size_t i, first = 0, queued = 0, finished = 0;
	int k, rc = 0, dev;

	if (bjxa_gpu_count() <= 0)
		FAIL(ENODEV);
	if ((dev = bjxa_gpu_current()) < 0)
		FAIL(ENODEV);
	if (sg->device != dev + 1) {
		/* first call of the thread, or the thread has moved to another device */
		stage_release(sg);
		sg->device = dev + 1;
		bjxa_thread_cache_used();
	}

	for (i = 0; i <= n && rc == 0; i++) {
		if (i < n) {
			in_chunk += (uint64_t)work[i].blocks * 64u * work[i].channels;
			if (in_chunk < CHUNK_BYTES && i + 1 < n)
				continue;
		} else if (first >= n) {
			break;
		}
		/* streams [first, i] form the next chunk */
		if (queued - finished == PIPE_DEPTH) {
			rc = chunk_finish(sg, kind, &ring[finished % PIPE_DEPTH], work, dsts);
			finished++;
			if (rc < 0)
				break;
		}
		ring[queued % PIPE_DEPTH].first = first;
		ring[queued % PIPE_DEPTH].count = (i < n ? i + 1 : n) - first;
		ring[queued % PIPE_DEPTH].slot = (int)(queued % PIPE_DEPTH);
		rc = chunk_start(sg, kind, &ring[queued % PIPE_DEPTH], work, srcs, xa_bytes);
		queued++;
		first = i + 1;
		in_chunk = 0;
		/* while this chunk uploads: the results of the one before it, and
		 * its downloads queued -- the device-to-host engine is the
		 * bottleneck of the whole call and must never wait for this loop */
		while (rc == 0 && queued - finished >= 2) {
			rc = chunk_finish(sg, kind, &ring[finished % PIPE_DEPTH], work, dsts);
			finished++;
		}
	}
	while (rc == 0 && finished < queued) {
		rc = chunk_finish(sg, kind, &ring[finished % PIPE_DEPTH], work, dsts);
		finished++;
	}
	/* drain every slot, also on the error path: the caller's buffers must not
	 * be written to after we return */
	{
		int e = errno;

		for (k = 0; k < PIPE_DEPTH; k++)
			if (sg->stream[k] != NULL && bjxa_gpu_sync(sg->stream[k]) < 0 && rc == 0) {
				rc = -1;
				e = errno;
			}
		errno = e;
	}
	return (rc);
}

/* argument checks of bjxa_decode, in the reference's order (libbjxa.c:612-620) */
static int
decode_precheck(bjxa_decoder_t *dec, void *dst, size_t dst_len, const void *src,
    size_t src_len)
{
	NEED_OBJ(dec, DECODER_TAG);
	NEED_PTR(dst);
	NEED_PTR(src);
	if (dec->left.sample_bits != 16)
		FAIL(EINVAL);
	if (dec->left.blocks == 0)
		FAIL(EPROTO);
	if (dst_len < dec->left.block_size_pcm)
		FAIL(ENOBUFS);
	if (src_len < dec->left.block_size_xa)
		FAIL(ENOBUFS);
	return (0);
}

/* libbjxa.c:770-778 */
static int
encode_precheck(bjxa_encoder_t *enc, void *dst, size_t dst_len, const void *src,
    size_t src_len)
{
	NEED_OBJ(enc, ENCODER_TAG);
	NEED_PTR(dst);
	NEED_PTR(src);
	if (enc->left.sample_bits != 16)
		FAIL(EINVAL);
	if (enc->left.blocks == 0)
		FAIL(EPROTO);
	if (dst_len < enc->left.block_size_xa)
		FAIL(ENOBUFS);
	if (src_len < enc->left.block_size_pcm)
		FAIL(ENOBUFS);
	return (0);
}

/* ---- the hot-path entry points -------------------------------------------- */

int
bjxa_batch_decode(bjxa_decoder_t *const *decs, void *const *dsts,
    const size_t *dst_lens, const void *const *srcs, const size_t *src_lens,
    int *results, int *errnos, size_t n)
{
	bjxa_stream_desc_t *work;
	uint32_t *xa_bytes;
	size_t i, live = 0;
	int rc = 0;

	if (n == 0)
		return (0);
	NEED_PTR(decs);
	NEED_PTR(dsts);
	NEED_PTR(dst_lens);
	NEED_PTR(srcs);
	NEED_PTR(src_lens);
	NEED_PTR(results);
	NEED_PTR(errnos);

	work = calloc(n, sizeof *work);
	xa_bytes = calloc(n, sizeof *xa_bytes);
	if (work == NULL || xa_bytes == NULL) {
		free(work);
		free(xa_bytes);
		FAIL(ENOMEM);
	}
	for (i = 0; i < n; i++) {
		uint32_t nb, pcm;

		errnos[i] = 0;
		if (decode_precheck(decs[i], dsts[i], dst_lens[i], srcs[i],
		    src_lens[i]) < 0) {
			results[i] = -1;
			errnos[i] = errno;
			continue;
		}
		nb = blocks_this_call(&decs[i]->left, dst_lens[i], src_lens[i], &pcm);
		(void)bjxa_decoder_describe(decs[i], &work[i]);
		work[i].blocks = nb;
		work[i].pcm_len = pcm;
		xa_bytes[i] = nb * decs[i]->left.block_size_xa;
		results[i] = 0;
		live += nb != 0;
	}
	if (live != 0)
		rc = run_host_batch(BJXA_PLAN_DECODE, work, dsts, srcs, xa_bytes, n);
	if (rc == 0) {
		for (i = 0; i < n; i++) {
			if (work[i].blocks == 0)
				continue;
			(void)bjxa_decoder_commit(decs[i], &work[i]);
			results[i] = work[i].result;
			errnos[i] = work[i].error;
		}
	}
	i = (size_t)errno;
	free(work);
	free(xa_bytes);
	errno = (int)i;
	return (rc);
}

int
bjxa_decode(bjxa_decoder_t *dec, void *dst, size_t dst_len, const void *src,
    size_t src_len)
{
	int result = -1, err = 0;
	bjxa_decoder_t *decs[1];
	void *dsts[1];
	const void *srcs[1];

	if (decode_precheck(dec, dst, dst_len, src, src_len) < 0)
		return (-1);
	{
		/* a call of a few blocks (the reference CLI's default mode makes one
		 * per block, src/bjxa_decode.c:102-161): one launch, no plan */
		bjxa_stream_desc_t d;
		uint32_t pcm, nb = blocks_this_call(&dec->left, dst_len, src_len, &pcm);
		int rc;

		(void)bjxa_decoder_describe(dec, &d);
		d.blocks = nb;
		d.pcm_len = pcm;
		if (nb != 0 && (rc = bjxa_small_call(BJXA_PLAN_DECODE, &d, dst, src)) <= 0) {
			if (rc < 0)
				return (-1);
			bjxa_thread_cache_used();
			(void)bjxa_decoder_commit(dec, &d);
			if (d.result < 0)
				errno = d.error;
			return (d.result);
		}
	}
	decs[0] = dec;
	dsts[0] = dst;
	srcs[0] = src;
	if (bjxa_batch_decode(decs, dsts, &dst_len, srcs, &src_len, &result, &err,
	    1) < 0)
		return (-1);
	if (result < 0)
		errno = err;
	return (result);
}

int
bjxa_batch_encode(bjxa_encoder_t *const *encs, void *const *dsts,
    const size_t *dst_lens, const void *const *srcs, const size_t *src_lens,
    int *results, int *errnos, size_t n)
{
	bjxa_stream_desc_t *work;
	uint32_t *xa_bytes;
	size_t i, live = 0;
	int rc = 0;

	if (n == 0)
		return (0);
	NEED_PTR(encs);
	NEED_PTR(dsts);
	NEED_PTR(dst_lens);
	NEED_PTR(srcs);
	NEED_PTR(src_lens);
	NEED_PTR(results);
	NEED_PTR(errnos);

	work = calloc(n, sizeof *work);
	xa_bytes = calloc(n, sizeof *xa_bytes);
	if (work == NULL || xa_bytes == NULL) {
		free(work);
		free(xa_bytes);
		FAIL(ENOMEM);
	}
	for (i = 0; i < n; i++) {
		uint32_t nb, pcm;

		errnos[i] = 0;
		if (encode_precheck(encs[i], dsts[i], dst_lens[i], srcs[i],
		    src_lens[i]) < 0) {
			results[i] = -1;
			errnos[i] = errno;
			continue;
		}
		/* PCM is the source here, XA the destination */
		nb = blocks_this_call(&encs[i]->left, src_lens[i], dst_lens[i], &pcm);
		(void)bjxa_encoder_describe(encs[i], &work[i]);
		work[i].blocks = nb;
		work[i].pcm_len = pcm;
		xa_bytes[i] = nb * encs[i]->left.block_size_xa;
		results[i] = 0;
		live += nb != 0;
	}
	if (live != 0)
		rc = run_host_batch(BJXA_PLAN_ENCODE, work, dsts, srcs, xa_bytes, n);
	if (rc == 0) {
		for (i = 0; i < n; i++) {
			if (work[i].blocks == 0)
				continue;
			(void)bjxa_encoder_commit(encs[i], &work[i]);
			results[i] = work[i].result;
			errnos[i] = work[i].error;
		}
	}
	i = (size_t)errno;
	free(work);
	free(xa_bytes);
	errno = (int)i;
	return (rc);
}

int
bjxa_encode(bjxa_encoder_t *enc, void *dst, size_t dst_len, const void *src,
    size_t src_len)
{
	int result = -1, err = 0;
	bjxa_encoder_t *encs[1];
	void *dsts[1];
	const void *srcs[1];

	if (encode_precheck(enc, dst, dst_len, src, src_len) < 0)
		return (-1);
	{
		bjxa_stream_desc_t d;
		uint32_t pcm, nb = blocks_this_call(&enc->left, src_len, dst_len, &pcm);
		int rc;

		(void)bjxa_encoder_describe(enc, &d);
		d.blocks = nb;
		d.pcm_len = pcm;
		if (nb != 0 && (rc = bjxa_small_call(BJXA_PLAN_ENCODE, &d, dst, src)) <= 0) {
			if (rc < 0)
				return (-1);
			bjxa_thread_cache_used();
			(void)bjxa_encoder_commit(enc, &d);
			return (d.result);
		}
	}
	encs[0] = enc;
	dsts[0] = dst;
	srcs[0] = src;
	if (bjxa_batch_encode(encs, dsts, &dst_len, srcs, &src_len, &result, &err,
	    1) < 0)
		return (-1);
	if (result < 0)
		errno = err;
	return (result);
}
