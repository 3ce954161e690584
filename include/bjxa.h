/*
 * bjxa.h -- public C API of libbjxa, as served by the B200-native backend.
 *
 * This header declares exactly the interface of the reference library
 * (/root/reference/src/bjxa.h:18-65, exported through src/libbjxa.map:16-47,
 * contract in bjxa.3.rst.in): same names, argument order, return values and
 * errno behaviour, so a program written against the reference relinks
 * unchanged.  What differs is behind bjxa_decode() and bjxa_encode(): the
 * block transform runs on the GPU (CUDA kernels for sm_100a); there is no CPU
 * fallback -- without a usable CUDA device those two calls fail with ENODEV.
 *
 * Unlike the reference header this one is self-contained (it includes what it
 * needs; the reference asks the caller to include <stdint.h>, <stdio.h> and
 * <unistd.h> first, bjxa.3.rst.in:44-48 -- doing so remains harmless).
 *
 * The additive batched / device-resident entry points are in bjxa_batch.h.
 */
#ifndef BJXA_H
#define BJXA_H

#include <stdint.h>
#include <stdio.h>
#include <sys/types.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ref: src/bjxa.h:18-19 */
#define BJXA_HEADER_SIZE_XA	32
#define BJXA_HEADER_SIZE_RIFF	44

/* ref: src/bjxa.h:21-22 -- opaque, magic-checked codec objects */
typedef struct bjxa_decoder bjxa_decoder_t;
typedef struct bjxa_encoder bjxa_encoder_t;

/*
 * ref: src/bjxa.h:24-32 -- field order and widths are ABI.
 * sample_bits is 16 (PCM) after bjxa_decode_format / bjxa_parse_riff_header
 * and the XA width (4, 6, 8) after bjxa_encode_format
 * (src/libbjxa.c:590,747).
 */
typedef struct {
	uint32_t	data_len_pcm;
	uint32_t	blocks;
	uint8_t		block_size_pcm;
	uint8_t		block_size_xa;
	uint16_t	samples_rate;
	uint8_t		sample_bits;
	uint8_t		channels;
} bjxa_format_t;

/* ---- decoder: symbol node LIBBJXA_0.1 (src/libbjxa.map:16-32) ---------- */

/* ref: src/bjxa.h:36-37, src/libbjxa.c:246-263 */
bjxa_decoder_t	*bjxa_decoder(void);
int		 bjxa_free_decoder(bjxa_decoder_t **decp);

/* ref: src/bjxa.h:39-40, src/libbjxa.c:395-476 -- host, 32 bytes per stream */
ssize_t		 bjxa_parse_header(bjxa_decoder_t *dec, const void *src, size_t len);
ssize_t		 bjxa_fread_header(bjxa_decoder_t *dec, FILE *file);

/* ref: src/bjxa.h:42, src/libbjxa.c:580-600 */
int		 bjxa_decode_format(bjxa_decoder_t *dec, bjxa_format_t *fmt);

/*
 * ref: src/bjxa.h:43, src/libbjxa.c:602-661 -- THE HOT PATH.  dst and src are
 * host buffers; returns the number of effective blocks decoded by this call
 * or -1 with errno set (EFAULT, EINVAL, EPROTO, ENOBUFS in the reference's
 * order, src/libbjxa.c:612-620; ENODEV when no CUDA device can be used).
 */
int		 bjxa_decode(bjxa_decoder_t *dec, void *dst, size_t dst_len,
		    const void *src, size_t src_len);

/* ref: src/bjxa.h:45-46, src/libbjxa.c:898-945 */
ssize_t		 bjxa_dump_riff_header(bjxa_decoder_t *dec, void *dst, size_t len);
ssize_t		 bjxa_fwrite_riff_header(bjxa_decoder_t *dec, FILE *file);

/* ref: src/bjxa.h:48-49, src/libbjxa.c:947-996 */
int		 bjxa_dump_pcm(void *dst, const int16_t *src, size_t len);
int		 bjxa_fwrite_pcm(const int16_t *src, size_t len, FILE *file);

/* ---- encoder: symbol node LIBBJXA_0.5 (src/libbjxa.map:34-47) ---------- */

/* ref: src/bjxa.h:53-54, src/libbjxa.c:265-282 */
bjxa_encoder_t	*bjxa_encoder(void);
int		 bjxa_free_encoder(bjxa_encoder_t **encp);

/* ref: src/bjxa.h:56, src/libbjxa.c:693-735 -- writes blocks and block sizes
 * back into *fmt */
int		 bjxa_encode_init(bjxa_encoder_t *enc, bjxa_format_t *fmt, uint8_t bits);

/* ref: src/bjxa.h:58-59, src/libbjxa.c:826-896 */
ssize_t		 bjxa_parse_riff_header(bjxa_format_t *fmt, const void *src, size_t len);
ssize_t		 bjxa_fread_riff_header(bjxa_format_t *fmt, FILE *file);

/* ref: src/bjxa.h:61, src/libbjxa.c:737-757 */
int		 bjxa_encode_format(bjxa_encoder_t *enc, bjxa_format_t *fmt);

/*
 * ref: src/bjxa.h:62, src/libbjxa.c:759-819 -- THE HOT PATH (encode side).
 * Reference-exact: profile byte 0 and top-bits truncation
 * (src/libbjxa.c:679,349-391).
 */
int		 bjxa_encode(bjxa_encoder_t *enc, void *dst, size_t dst_len,
		    const void *src, size_t src_len);

/* ref: src/bjxa.h:64-65, src/libbjxa.c:478-521 */
ssize_t		 bjxa_dump_header(bjxa_encoder_t *enc, void *dst, size_t len);
ssize_t		 bjxa_fwrite_header(bjxa_encoder_t *enc, FILE *file);

#ifdef __cplusplus
}
#endif
#endif /* BJXA_H */
