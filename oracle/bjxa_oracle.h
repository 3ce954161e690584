/*
 * bjxa_oracle.h -- CPU restatement of libbjxa's block transform.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it, and only as the checker.  The product
 * library (bjxa_b200/lib/libbjxa_b200.so) never links or calls this file.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 * every golden the reference's own tests hold for the path (the six SHA-1s and
 * the saturation vector of test/test_decode.sh:24-122, the EPROTO cases of
 * test/test_decode_error.sh:221-282) and, when oracle/_ref/ has been built,
 * differentially against the unmodified reference compiled from
 * /root/reference/src/libbjxa.c.
 *
 * All "ref:" citations are path:line under /root/reference/.
 */
#ifndef BJXA_ORACLE_H
#define BJXA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XAO_BLOCK_SAMPLES 32

/* Stream geometry, as the reference keeps it in struct bjxa_decoder /
 * struct bjxa_encoder (ref: src/libbjxa.c:217-242). */
typedef struct {
	uint32_t data_len;	/* XA payload bytes (nDataLen) */
	uint32_t samples;	/* samples per channel (nSamples) */
	uint16_t rate;		/* nSamplesPerSec */
	uint8_t  bits;		/* 4, 6 or 8 */
	uint8_t  channels;	/* 1 or 2 */
	int16_t  prev[2][2];	/* [channel][0 = s(n-1), 1 = s(n-2)] */
} xao_stream_t;

/* One 32-sample block of one channel. */

/* ref: src/libbjxa.c:286-345 -- returns the profile byte; writes 32 int16
 * (code in the top bits) to dst with the given stride. */
uint8_t xao_inflate(unsigned bits, int16_t *dst, unsigned stride,
    const uint8_t *block);

/* ref: src/libbjxa.c:349-391 -- writes 4*bits payload bytes. */
void xao_deflate(unsigned bits, uint8_t *payload, const int16_t *src);

/* ref: src/libbjxa.c:525-578 -- in-place predictor over 32 strided samples.
 * Returns 0, or -1 (nothing touched) when profile>>4 >= 5. */
int xao_predict(int16_t *samples, unsigned stride, uint8_t profile,
    int16_t prev[2]);

/* Block loops. */

/* ref: src/libbjxa.c:602-661.  Decodes `blocks` effective blocks (L then R
 * when stereo) from xa into pcm (interleaved host-endian int16), never writing
 * more than *pcm_left bytes in total; the last block is truncated exactly like
 * the reference's pcm_block logic.  prev[][] is advanced in place, including
 * the reference's partial update when the right block of a stereo pair has a
 * bad profile.  Returns the number of effective blocks fully decoded; when a
 * bad profile stops the loop, *bad is set to 1 (else 0) -- the reference
 * returns -1/EPROTO in that case but keeps the blocks already copied.
 * On return *pcm_left has been reduced by the bytes written. */
long xao_decode_blocks(unsigned bits, unsigned channels, int16_t prev[2][2],
    const uint8_t *xa, uint32_t blocks, int16_t *pcm, uint32_t *pcm_left,
    int *bad);

/* ref: src/libbjxa.c:665-691, 759-819.  Encodes interleaved PCM (pcm_bytes
 * bytes, a whole number of frames) to ceil(frames/32) effective blocks with
 * profile byte 0 and top-bits truncation; the short last block is
 * zero-padded.  Returns the number of effective blocks written. */
long xao_encode_blocks(unsigned bits, unsigned channels, const int16_t *pcm,
    uint32_t pcm_bytes, uint8_t *xa);

/* EXTENSION, not in the reference (the reference
 * hard-codes profile 0, src/libbjxa.c:679): per-block search over filter 0..4 x
 * range 0..16-bits, closed loop against the reference decoder's arithmetic
 * (src/libbjxa.c:556-571), error = sum (x - decoded)^2, ties -> lowest profile
 * byte; the exact rule is spelled out in bjxa_b200/csrc/xa_core.h.  prev is the
 * decoder state on entry and on return.  Every emitted block is decoded again
 * with xao_inflate + xao_predict and compared with the search's own
 * reconstruction: returns -1 if they ever differ, else the number of effective
 * blocks written.  Parity for this function is pinned by properties only
 * (tests/test_search.py): there is no reference implementation. */
long xao_encode_search_blocks(unsigned bits, unsigned channels,
    int16_t prev[2][2], const int16_t *pcm, uint32_t pcm_bytes, uint8_t *xa);

/* Container helpers (cold path; only here so the oracle can be checked
 * against the reference's whole-file goldens). */

/* ref: src/libbjxa.c:395-453.  0 on success, -1 on a header the reference
 * rejects with EPROTO. */
int xao_parse_xa_header(xao_stream_t *st, const uint8_t hdr[32]);

/* ref: src/libbjxa.c:478-503 (always writes zero loop/prev/pad). */
void xao_write_xa_header(const xao_stream_t *st, uint8_t hdr[32]);

/* ref: src/libbjxa.c:898-927. */
void xao_write_riff_header(const xao_stream_t *st, uint8_t hdr[44]);

/* ref: src/libbjxa.c:826-873.  0 on success and fills channels/rate and
 * *pcm_bytes; -1 on EPROTO. */
int xao_parse_riff_header(xao_stream_t *st, uint32_t *pcm_bytes,
    const uint8_t hdr[44]);

/* ref: src/libbjxa.c:693-735 -- geometry of the XA stream that encoding
 * pcm_bytes of PCM produces.  0, or -1 where the reference says EPROTO. */
int xao_encode_geometry(xao_stream_t *st, uint32_t pcm_bytes, unsigned bits,
    unsigned channels, unsigned rate);

/* Whole-file conveniences used by the tests and the CPU baseline. */

/* .xa image -> .wav image.  Returns bytes written to wav (44 + PCM), or -1. */
long xao_xa_to_wav(const uint8_t *xa, size_t xa_len, uint8_t *wav,
    size_t wav_cap);

/* .wav image -> .xa image.  Returns bytes written (32 + payload), or -1. */
long xao_wav_to_xa(const uint8_t *wav, size_t wav_len, unsigned bits,
    uint8_t *xa, size_t xa_cap);

#ifdef __cplusplus
}
#endif
#endif
