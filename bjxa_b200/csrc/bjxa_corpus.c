/*
 * bjxa_corpus.c -- whole files in, whole files out (include/bjxa_batch.h).
 *
 * The container step on either side of the block transform: what the
 * reference's CLI does around bjxa_decode / bjxa_encode for one file
 * (/root/reference/src/bjxa_decode.c:38-93, src/bjxa_encode.c) -- parse and
 * validate the header, emit the other header, run the block loop -- done for a
 * corpus of files that lie in one host arena.  Headers are parsed and composed
 * by the library's own bjxa_parse_header / bjxa_dump_riff_header /
 * bjxa_parse_riff_header / bjxa_encode_init / bjxa_dump_header (bjxa_host.c,
 * the reference's semantics and bytes); the data moves as one host-to-device
 * copy and one device-to-host copy per chunk of consecutive files, the files
 * being assembled on the device (block-loop kernels for the data,
 * bjxa_gpu_scatter_async for the headers).  No file data is touched by the CPU.
 */
#define _POSIX_C_SOURCE 200809L

#include <errno.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/bjxa.h"
#include "../../include/bjxa_batch.h"
#include "bjxa_internal.h"

#define DEPTH		3			/* chunks in flight */
#define CHUNK_IN	((uint64_t)48 << 20)	/* input bytes per chunk, about */

/* BJXA_B200_CORPUS_CHUNK=bytes: smaller chunks, for tests of the pipeline */
static uint64_t
chunk_in(void)
{
	const char *e = getenv("BJXA_B200_CORPUS_CHUNK");
	uint64_t v = e != NULL ? strtoull(e, NULL, 0) : 0;

	return (v != 0 ? v : CHUNK_IN);
}
#define ALIGN16(x)	(((x) + 15u) & ~(uint64_t)15u)

struct rec {				/* bjxa_gpu_scatter_async's 64-byte slot */
	uint64_t	off;
	uint8_t		bytes[56];
};

struct slot {
	void		*stream;
	bjxa_plan_t	*plan;
	void		*d_in, *d_out, *d_tab;
	size_t		 cap_in, cap_out, cap_tab;
	struct rec	*tab;			/* pinned */
	size_t		 cap_rec;
	bjxa_stream_desc_t *desc;
	size_t		 cap_desc;
	/* the chunk in flight */
	size_t		 first, count;
	uint64_t	 out0, out_len;		/* its range of the output arena */
	int		 busy;
};

/* one file's header, as the reference reads it */
struct parsed {
	bjxa_stream_desc_t d;		/* blocks, pcm_len, bits, channels, prev */
	uint8_t		hdr[56];	/* the header of the produced file */
	uint32_t	hdr_len;	/* 44 or 32 */
	uint64_t	data_in;	/* data bytes the input file must hold */
	uint64_t	data_out;	/* data bytes of the produced file */
	uint16_t	rate;
	int		error;
};

static void
parse_one(int kind, bjxa_decoder_t *dec, bjxa_encoder_t *enc, const uint8_t *file,
    uint64_t len, unsigned bits, struct parsed *p)
{
	bjxa_format_t fmt;

	memset(p, 0, sizeof *p);
	if (kind == BJXA_CORPUS_XA_TO_WAV) {
		if (len < BJXA_HEADER_SIZE_XA) {
			p->error = EIO;		/* bjxa_fread_header: short read */
			return;
		}
		if (bjxa_parse_header(dec, file, BJXA_HEADER_SIZE_XA) < 0 ||
		    bjxa_decode_format(dec, &fmt) < 0 ||
		    bjxa_decoder_describe(dec, &p->d) < 0 ||
		    bjxa_dump_riff_header(dec, p->hdr, BJXA_HEADER_SIZE_RIFF) < 0) {
			p->error = errno;
			return;
		}
		p->hdr_len = BJXA_HEADER_SIZE_RIFF;
		p->data_in = (uint64_t)fmt.blocks * fmt.block_size_xa;
		p->data_out = fmt.data_len_pcm;
		p->rate = fmt.samples_rate;
		if (len < BJXA_HEADER_SIZE_XA + p->data_in)
			p->error = EIO;		/* src/bjxa_decode.c:78-83 */
		return;
	}
	if (len < BJXA_HEADER_SIZE_RIFF) {
		p->error = EIO;
		return;
	}
	if (bjxa_parse_riff_header(&fmt, file, BJXA_HEADER_SIZE_RIFF) < 0 ||
	    bjxa_encode_init(enc, &fmt, (uint8_t)bits) < 0 ||
	    bjxa_encoder_describe(enc, &p->d) < 0 ||
	    bjxa_dump_header(enc, p->hdr, BJXA_HEADER_SIZE_XA) < 0) {
		p->error = errno;
		return;
	}
	p->hdr_len = BJXA_HEADER_SIZE_XA;
	p->data_in = fmt.data_len_pcm;
	p->data_out = (uint64_t)fmt.blocks * fmt.block_size_xa;
	p->rate = fmt.samples_rate;
	if (len < BJXA_HEADER_SIZE_RIFF + p->data_in)
		p->error = EIO;
}

/*
 * Where the produced file goes: `cur` is the first free byte of the output
 * arena.  A WAV file's PCM must start 16-byte aligned (the decode kernels store
 * 128-bit units), an XA file may start anywhere.
 */
static uint64_t
place(int kind, uint64_t cur)
{
	if (kind == BJXA_CORPUS_XA_TO_WAV)
		return (ALIGN16(cur + BJXA_HEADER_SIZE_RIFF) - BJXA_HEADER_SIZE_RIFF);
	return (cur);
}

/* *packed: every file's data keeps its 16-byte phase when the arena travels as
 * one piece (always true for XA input, whose blocks may start anywhere) */
static int
check_args(int kind, const void *in, const bjxa_file_desc_t *files, size_t n,
    int *packed)
{
	size_t i;

	*packed = 1;
	if (kind != BJXA_CORPUS_XA_TO_WAV && kind != BJXA_CORPUS_WAV_TO_XA) {
		errno = EINVAL;
		return (-1);
	}
	if ((in == NULL || files == NULL) && n != 0) {
		errno = EFAULT;
		return (-1);
	}
	for (i = 0; i < n; i++) {
		if (i > 0 && files[i].in_off < files[i - 1].in_off + files[i - 1].in_len) {
			errno = EINVAL;		/* ascending, non-overlapping */
			return (-1);
		}
		if (kind == BJXA_CORPUS_WAV_TO_XA &&
		    (files[i].in_off + BJXA_HEADER_SIZE_RIFF) % 16 != 0)
			*packed = 0;	/* the encode kernels load PCM 16 bytes at a time */
	}
	return (0);
}

int
bjxa_corpus_extent(int kind, const void *in_arena, size_t in_bytes,
    const bjxa_file_desc_t *files, size_t n, uint64_t *out_bytes)
{
	bjxa_decoder_t *dec;
	bjxa_encoder_t *enc;
	struct parsed p;
	uint64_t cur = 0;
	size_t i;
	int packed;

	if (out_bytes == NULL) {
		errno = EFAULT;
		return (-1);
	}
	if (check_args(kind, in_arena, files, n, &packed) < 0)
		return (-1);
	dec = bjxa_decoder();
	enc = bjxa_encoder();
	if (dec == NULL || enc == NULL) {
		(void)bjxa_free_decoder(&dec);
		(void)bjxa_free_encoder(&enc);
		errno = ENOMEM;
		return (-1);
	}
	for (i = 0; i < n; i++) {
		if (files[i].in_off + files[i].in_len > in_bytes) {
			(void)bjxa_free_decoder(&dec);
			(void)bjxa_free_encoder(&enc);
			errno = ENOBUFS;
			return (-1);
		}
		parse_one(kind, dec, enc, (const uint8_t *)in_arena + files[i].in_off,
		    files[i].in_len, files[i].bits, &p);
		if (p.error != 0)
			continue;
		cur = place(kind, cur) + p.hdr_len + p.data_out;
	}
	(void)bjxa_free_decoder(&dec);
	(void)bjxa_free_encoder(&enc);
	*out_bytes = cur + 16;		/* the kernels may be asked for a 16-byte multiple */
	return (0);
}

static int
grow_dev(void **p, size_t *cap, size_t need)
{
	if (need <= *cap)
		return (0);
	if (*p != NULL)
		(void)bjxa_gpu_free(*p);
	*cap = 0;
	need += need / 4 + 4096;
	*p = bjxa_gpu_alloc(need);
	if (*p == NULL)
		return (-1);
	*cap = need;
	return (0);
}

/*
 * The pipeline's buffers, streams and plans stay with the calling thread from
 * one call to the next (allocating and freeing a few hundred MB of device and
 * pinned memory costs more than moving a small corpus); a failed call gives
 * them back.
 */
static __thread struct slot tls_ring[DEPTH];
static __thread int tls_ring_dev;	/* the device they live on, + 1; 0 = nothing yet */

static void
slot_release(struct slot *s)
{
	if (s->plan != NULL)
		(void)bjxa_plan_free(&s->plan);
	if (s->d_in != NULL)
		(void)bjxa_gpu_free(s->d_in);
	if (s->d_out != NULL)
		(void)bjxa_gpu_free(s->d_out);
	if (s->d_tab != NULL)
		(void)bjxa_gpu_free(s->d_tab);
	if (s->tab != NULL)
		(void)bjxa_host_free(s->tab);
	free(s->desc);
	if (s->stream != NULL)
		(void)bjxa_gpu_stream_destroy(s->stream);
	memset(s, 0, sizeof *s);
}

/* the calling thread's pipeline, on whichever device it was built for */
void
bjxa_corpus_release(void)
{
	int i, cur;

	if (tls_ring_dev == 0)
		return;
	cur = bjxa_gpu_current();
	if (cur >= 0 && cur != tls_ring_dev - 1)
		(void)bjxa_gpu_select(tls_ring_dev - 1);
	for (i = 0; i < DEPTH; i++)
		slot_release(&tls_ring[i]);
	if (cur >= 0 && cur != tls_ring_dev - 1)
		(void)bjxa_gpu_select(cur);
	tls_ring_dev = 0;
}

/* what bjxa_plan_create would refuse (xa_plan.h:build_plan): such a file gets
 * its own error instead of failing the whole call */
static int
desc_ok(const bjxa_stream_desc_t *d)
{
	return ((d->bits == 4 || d->bits == 6 || d->bits == 8) &&
	    (d->channels == 1 || d->channels == 2) && d->blocks != 0 &&
	    d->pcm_len % (2u * d->channels) == 0 &&
	    (uint64_t)d->pcm_len <= (uint64_t)d->blocks * 64u * d->channels);
}

/* the chunk's results: per-file status, then its range of the output arena */
static int
slot_finish(struct slot *s, int kind, void *out_arena, bjxa_file_desc_t *files)
{
	size_t k;

	if (!s->busy)
		return (0);
	s->busy = 0;
	if (bjxa_plan_fetch(s->plan, s->desc, s->count) < 0)
		return (-1);
	for (k = 0; k < s->count; k++) {
		bjxa_file_desc_t *f = &files[s->first + k];
		const bjxa_stream_desc_t *d = &s->desc[k];

		if (f->error != 0)
			continue;		/* rejected at the header: not in the plan */
		f->blocks = d->done;
		if (d->error != 0) {
			/* bad profile (src/libbjxa.c:550): the header and the blocks
			 * in front of it are what the reference CLI has written */
			f->error = d->error;
			f->out_len = BJXA_HEADER_SIZE_RIFF +
			    (uint64_t)d->done * 64u * d->channels;
		}
	}
	if (s->out_len != 0 && bjxa_gpu_download_async((uint8_t *)out_arena + s->out0,
	    (const uint8_t *)s->d_out + (s->out0 & 15u), s->out_len, s->stream) < 0)
		return (-1);
	(void)kind;
	return (0);
}

int
bjxa_corpus_run(int kind, const void *in_arena, size_t in_bytes, void *out_arena,
    size_t out_bytes, bjxa_file_desc_t *files, size_t n)
{
	struct slot *ring = tls_ring;
	bjxa_decoder_t *dec = NULL;
	bjxa_encoder_t *enc = NULL;
	struct parsed p;
	uint64_t cur = 0;
	size_t first = 0, turn = 0, i, k;
	const uint64_t chunk = chunk_in();
	int rc = -1, d, packed;

	if (check_args(kind, in_arena, files, n, &packed) < 0)
		return (-1);
	if (out_arena == NULL && n != 0) {
		errno = EFAULT;
		return (-1);
	}
	if (n == 0)
		return (0);
	if (bjxa_gpu_count() <= 0) {
		errno = ENODEV;		/* no CPU path */
		return (-1);
	}
	if ((d = bjxa_gpu_current()) < 0) {
		errno = ENODEV;
		return (-1);
	}
	if (tls_ring_dev != d + 1) {
		/* first call of the thread, or the thread has moved to another device */
		bjxa_corpus_release();
		tls_ring_dev = d + 1;
		bjxa_thread_cache_used();
	}
	for (i = 0; i < DEPTH; i++)
		ring[i].busy = 0;
	dec = bjxa_decoder();
	enc = bjxa_encoder();
	if (dec == NULL || enc == NULL) {
		errno = ENOMEM;
		goto out;
	}

	while (first < n) {
		struct slot *s = &ring[turn % DEPTH];
		uint64_t in0, in_end, want_in, want_out;
		size_t count = 0, live = 0;
		uint64_t dcur = 0;		/* unpacked input: next free device byte */

		/* the slot's previous chunk: results, download, and only then reuse */
		if (slot_finish(s, kind, out_arena, files) < 0)
			goto out;
		if (s->stream == NULL && (s->stream = bjxa_gpu_stream_create()) == NULL)
			goto out;
		if (bjxa_gpu_sync(s->stream) < 0)
			goto out;

		/* consecutive files up to about CHUNK_IN bytes of input */
		in0 = files[first].in_off;
		in_end = in0;
		while (first + count < n &&
		    (count == 0 || files[first + count].in_off +
		    files[first + count].in_len - in0 <= chunk)) {
			const bjxa_file_desc_t *f = &files[first + count];
			if (f->in_off + f->in_len > in_bytes) {
				errno = ENOBUFS;
				goto out;
			}
			in_end = f->in_off + f->in_len;
			count++;
		}
		if (count > s->cap_desc) {
			free(s->desc);
			s->desc = calloc(count + count / 4 + 16, sizeof *s->desc);
			s->cap_desc = s->desc == NULL ? 0 : count + count / 4 + 16;
			if (s->desc == NULL) {
				errno = ENOMEM;
				goto out;
			}
		}
		if (count > s->cap_rec) {
			if (s->tab != NULL)
				(void)bjxa_host_free(s->tab);
			s->cap_rec = count + count / 4 + 16;
			s->tab = bjxa_host_alloc(s->cap_rec * sizeof *s->tab);
			if (s->tab == NULL) {
				s->cap_rec = 0;
				goto out;
			}
		}

		/* headers: descriptors, the produced files' places and headers */
		s->first = first;
		s->count = count;
		s->out0 = place(kind, cur);
		for (k = 0; k < count; k++) {
			bjxa_file_desc_t *f = &files[first + k];
			bjxa_stream_desc_t *sd = &s->desc[k];
			const uint64_t rel = f->in_off - in0 + (in0 & 15u);

			parse_one(kind, dec, enc, (const uint8_t *)in_arena + f->in_off,
			    f->in_len, f->bits, &p);
			if (p.error == 0 && !desc_ok(&p.d))
				p.error = EINVAL;	/* a header the block loop cannot serve */
			f->error = p.error;
			f->out_off = f->out_len = 0;
			f->blocks = 0;
			f->channels = p.d.channels;
			f->rate = p.rate;
			if (kind == BJXA_CORPUS_XA_TO_WAV)
				f->bits = p.d.bits;
			memset(sd, 0, sizeof *sd);
			if (p.error != 0)
				continue;	/* blocks == 0: takes no part in the plan */
			*sd = p.d;
			f->out_off = place(kind, cur);
			f->out_len = p.hdr_len + p.data_out;
			if (f->out_off + f->out_len + 16 > out_bytes) {
				errno = ENOBUFS;
				goto out;
			}
			/* arena offsets inside the slot's device buffers, which keep the
			 * host arenas' alignment modulo 16 */
			if (kind == BJXA_CORPUS_XA_TO_WAV) {
				sd->xa_off = rel + BJXA_HEADER_SIZE_XA;
				sd->pcm_off = f->out_off - s->out0 + (s->out0 & 15u) + p.hdr_len;
			} else {
				if (packed) {
					sd->pcm_off = rel + BJXA_HEADER_SIZE_RIFF;
				} else {
					sd->pcm_off = dcur;	/* its own, aligned place */
					dcur += ALIGN16(p.data_in);
				}
				sd->xa_off = f->out_off - s->out0 + (s->out0 & 15u) + p.hdr_len;
			}
			s->tab[live].off = f->out_off - s->out0 + (s->out0 & 15u);
			memcpy(s->tab[live].bytes, p.hdr, p.hdr_len);
			live++;
			cur = f->out_off + f->out_len;
		}
		s->out_len = live != 0 ? cur - s->out0 : 0;

		if (live != 0) {
			want_in = (packed ? (in0 & 15u) + (in_end - in0) : dcur) + 64;
			want_out = (s->out0 & 15u) + s->out_len + 64;
			if (grow_dev(&s->d_in, &s->cap_in, want_in) < 0 ||
			    grow_dev(&s->d_out, &s->cap_out, want_out) < 0 ||
			    grow_dev(&s->d_tab, &s->cap_tab, live * sizeof *s->tab) < 0)
				goto out;
			if (packed) {
				if (bjxa_gpu_upload_async((uint8_t *)s->d_in + (in0 & 15u),
				    (const uint8_t *)in_arena + in0, in_end - in0, s->stream) < 0)
					goto out;
			} else {
				/* WAV data that is not 16-byte aligned in the arena: every
				 * file's PCM travels on its own to an aligned place */
				for (k = 0; k < count; k++)
					if (s->desc[k].blocks != 0 && bjxa_gpu_upload_async(
					    (uint8_t *)s->d_in + s->desc[k].pcm_off,
					    (const uint8_t *)in_arena + files[first + k].in_off +
					    BJXA_HEADER_SIZE_RIFF, s->desc[k].pcm_len, s->stream) < 0)
						goto out;
			}
			if (bjxa_gpu_upload_async(s->d_tab, s->tab, live * sizeof *s->tab,
			    s->stream) < 0)
				goto out;
			d = kind == BJXA_CORPUS_XA_TO_WAV ? BJXA_PLAN_DECODE : BJXA_PLAN_ENCODE;
			if (s->plan == NULL) {
				s->plan = bjxa_plan_create(d, s->desc, count);
				if (s->plan == NULL)
					goto out;
			} else if (bjxa_plan_reset(s->plan, d, s->desc, count) < 0) {
				goto out;
			}
			if (bjxa_plan_run(s->plan, s->d_out, s->cap_out, s->d_in, s->cap_in,
			    s->stream) < 0 ||
			    bjxa_gpu_scatter_async(s->d_out, s->d_tab,
			    kind == BJXA_CORPUS_XA_TO_WAV ? BJXA_HEADER_SIZE_RIFF :
			    BJXA_HEADER_SIZE_XA, live, s->stream) < 0)
				goto out;
			s->busy = 1;
		}
		first += count;
		turn++;
		/* while this chunk's upload is under way: the results of the chunk
		 * before it, and its download queued, so that the device-to-host
		 * engine never waits for this loop */
		if (turn >= 2 &&
		    slot_finish(&ring[(turn - 2) % DEPTH], kind, out_arena, files) < 0)
			goto out;
	}
	rc = 0;
out:
	{
		int saved = errno;
		for (i = 0; i < DEPTH; i++) {
			struct slot *s = &ring[(turn + i) % DEPTH];	/* oldest first */
			if (rc == 0 && slot_finish(s, kind, out_arena, files) < 0) {
				rc = -1;
				saved = errno;
			}
		}
		for (i = 0; i < DEPTH; i++) {
			if (ring[i].stream != NULL && bjxa_gpu_sync(ring[i].stream) < 0 &&
			    rc == 0) {
				rc = -1;
				saved = errno;
			}
			if (rc != 0)
				slot_release(&ring[i]);	/* start clean after a failure */
		}
		(void)bjxa_free_decoder(&dec);
		(void)bjxa_free_encoder(&enc);
		errno = saved;
	}
	return (rc);
}
