/*
 * bjxa_batch.h -- additive batched C ABI of the B200-native libbjxa backend.
 *
 * Nothing in this header exists in the reference; it is what a binding (cgo,
 * JNI, ctypes ...) uses to run MANY independent XA streams per GPU launch.
 * Plain pointers and sizes only.  Every stream behaves exactly as if it had
 * been passed alone to bjxa_decode() / bjxa_encode()
 * (/root/reference/src/libbjxa.c:602-661, 759-819): same bytes, same per-stream
 * return value, same errno, same codec-object bookkeeping.
 *
 * Three levels:
 *   1. bjxa_batch_decode / bjxa_batch_encode  -- arrays of codec objects and
 *      HOST buffers; the library stages, launches and copies back.
 *   2. bjxa_plan_*   -- DEVICE-resident arenas + a table of stream
 *      descriptors; nothing crosses PCIe in bjxa_plan_run().  This is the path
 *      the throughput benchmark times.
 *   3. bjxa_decoder_describe / _commit (and the encoder twins) -- glue between
 *      the reference's codec objects and level 2.
 *
 * All functions return 0 (or a count) on success and -1 with errno on failure,
 * like the rest of the API; they never print.  ENODEV: no usable CUDA device.
 */
#ifndef BJXA_BATCH_H
#define BJXA_BATCH_H

#include <stddef.h>
#include <stdint.h>

#include "bjxa.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- level 1: host buffers, one call per batch ------------------------- */

/*
 * For i in [0, n): results[i] = what bjxa_decode(decs[i], dsts[i], dst_lens[i],
 * srcs[i], src_lens[i]) returns, errnos[i] = the errno it would leave (0 when
 * results[i] >= 0).  Each decoder is updated exactly as the single call would
 * (fmt->blocks, fmt->data_len_pcm, predictor state).  A codec may appear only
 * once per batch.  Returns 0 when the batch ran (individual streams may still
 * have failed), -1/errno when it could not run at all.
 */
int bjxa_batch_decode(bjxa_decoder_t *const *decs, void *const *dsts,
    const size_t *dst_lens, const void *const *srcs, const size_t *src_lens,
    int *results, int *errnos, size_t n);

int bjxa_batch_encode(bjxa_encoder_t *const *encs, void *const *dsts,
    const size_t *dst_lens, const void *const *srcs, const size_t *src_lens,
    int *results, int *errnos, size_t n);

/* ---- level 2: device-resident arenas ----------------------------------- */

typedef struct bjxa_stream_desc {
	uint64_t xa_off;	/* first XA block: byte offset into the XA arena */
	uint64_t pcm_off;	/* PCM: byte offset into the PCM arena, multiple of 16 */
	uint32_t blocks;	/* effective blocks (L+R = 1) to process */
	uint32_t pcm_len;	/* decode: PCM bytes owed (truncates the last block
				 * like fmt->data_len_pcm, libbjxa.c:622-624);
				 * encode: PCM bytes available (rest is zero padding,
				 * libbjxa.c:686-690).  Multiple of 2*channels. */
	int16_t  prev[2][2];	/* decode: predictor state [channel][n-1, n-2];
				 * in = on entry (header befL/befR, libbjxa.c:417-420),
				 * out (bjxa_plan_fetch) = after the last block */
	uint8_t  bits;		/* 4, 6 or 8 */
	uint8_t  channels;	/* 1 or 2 */
	uint16_t reserved;
	uint32_t done;		/* out: effective blocks completed */
	int32_t  result;	/* out: what bjxa_decode/encode returns: done, or -1 */
	int32_t  error;		/* out: 0, or EPROTO (bad profile, libbjxa.c:550) */
} bjxa_stream_desc_t;

#define BJXA_PLAN_DECODE 0
#define BJXA_PLAN_ENCODE 1
/*
 * EXTENSION, not in the reference (whose encoder writes profile 0 and the top
 * bits of every sample, src/libbjxa.c:679): encode with a per-block search
 * over filter 0..4 x range 0..16-bits, closed loop against the reference
 * decoder's arithmetic (src/libbjxa.c:556-571), smallest squared error wins,
 * ties go to the lowest profile byte.  Same arenas and descriptors as
 * BJXA_PLAN_ENCODE; `prev` is the DEcoder state the stream's first block will
 * be decoded from (0 for a file: bjxa_dump_header writes no history) and, out
 * of bjxa_plan_fetch, the state after the last block, so a stream may be
 * encoded in several calls.  The output decodes with the reference decoder and
 * never has a larger error than BJXA_PLAN_ENCODE's.  Opt-in only: bjxa_encode
 * and bjxa_batch_encode stay reference-exact.
 */
#define BJXA_PLAN_ENCODE_SEARCH 2

typedef struct bjxa_plan bjxa_plan_t;

/* Validates descs (EINVAL), builds and uploads the tile tables. */
bjxa_plan_t *bjxa_plan_create(int kind, const bjxa_stream_desc_t *descs, size_t n);

/* Re-targets an existing plan at a new table, reusing its device allocations. */
int bjxa_plan_reset(bjxa_plan_t *plan, int kind, const bjxa_stream_desc_t *descs,
    size_t n);

/*
 * Launches the batch on `cuda_stream` (a cudaStream_t, NULL = default stream)
 * and returns without waiting.  dst/src are DEVICE pointers (16-byte aligned)
 * to the arenas the descriptors index: decode reads the XA arena `src` and
 * writes the PCM arena `dst`; encode the other way round.  ENOBUFS when an
 * arena is smaller than the descriptors need.
 * After a bad profile, PCM at and after the failing block is unspecified.
 */
int bjxa_plan_run(bjxa_plan_t *plan, void *dst, size_t dst_bytes, const void *src,
    size_t src_bytes, void *cuda_stream);

/* Waits for the last run and fills prev/done/result/error of descs[0..n). */
int bjxa_plan_fetch(bjxa_plan_t *plan, bjxa_stream_desc_t *descs, size_t n);

/*
 * Checksums, computed on the device, of what the last bjxa_plan_run() produced
 * for each stream -- decode: its pcm_len bytes of PCM; encode: its blocks *
 * channels * (4 * bits + 1) bytes of XA -- so that a whole batch can be verified
 * without bringing it back to the host:
 *     sums[s] = sum over i of word_i * ((i * 0x9E3779B97F4A7C15 +
 *               0xD1B54A32D192ED03) | 1)   mod 2^64,
 * word_i = bytes 4i .. 4i+3 of the stream's output, little endian, bytes past
 * its end taken as zero.  Waits for the run; streams without blocks give 0.
 */
int bjxa_plan_checksum(bjxa_plan_t *plan, uint64_t *sums, size_t n);

/* Kernel launches one bjxa_plan_run() issues (for launch accounting). */
int bjxa_plan_launches(const bjxa_plan_t *plan);
/*
 * Kernels bjxa_plan_run() has launched for this plan so far (counted launch by
 * launch).  With BJXA_B200_CENSUS_EVERY=n > 1 in the environment when the plan is
 * built, a decode plan asks the census -- and launches every candidate form -- only
 * in every n-th run; the runs in between launch the form chosen last alone.  Every
 * form decodes every batch, the census only picks the fastest: meant for callers
 * that run one plan over and over on data of one kind.  Default: every run.
 */
unsigned long long bjxa_plan_launched(const bjxa_plan_t *plan);

/* Arena bytes the plan reads / writes (so callers can size allocations). */
int bjxa_plan_extent(const bjxa_plan_t *plan, uint64_t *src_bytes, uint64_t *dst_bytes);

int bjxa_plan_free(bjxa_plan_t **planp);

/* ---- level 3: codec objects <-> descriptors ---------------------------- */

/*
 * Fills bits/channels/prev and the remaining blocks / pcm_len of a parsed
 * decoder (offsets are left to the caller).  EINVAL when the decoder is not
 * ready.  bjxa_decoder_commit() applies a fetched descriptor back: advances
 * fmt->blocks / fmt->data_len_pcm by `done` blocks and stores the state.
 */
int bjxa_decoder_describe(bjxa_decoder_t *dec, bjxa_stream_desc_t *desc);
int bjxa_decoder_commit(bjxa_decoder_t *dec, const bjxa_stream_desc_t *desc);
int bjxa_encoder_describe(bjxa_encoder_t *enc, bjxa_stream_desc_t *desc);
int bjxa_encoder_commit(bjxa_encoder_t *enc, const bjxa_stream_desc_t *desc);

/* ---- device helpers (so a C caller need not link the CUDA runtime) ----- */

int    bjxa_gpu_count(void);			/* usable CUDA devices, 0 if none */
int    bjxa_gpu_select(int device);		/* cudaSetDevice for this thread */
void  *bjxa_gpu_alloc(size_t bytes);		/* device memory, 256-byte aligned */
int    bjxa_gpu_free(void *dptr);
void  *bjxa_host_alloc(size_t bytes);		/* pinned host memory */
int    bjxa_host_free(void *hptr);
int    bjxa_gpu_upload(void *dptr, const void *hptr, size_t bytes);
int    bjxa_gpu_download(void *hptr, const void *dptr, size_t bytes);
int    bjxa_gpu_sync(void *cuda_stream);
/* asynchronous forms on a stream of the caller's (overlap needs pinned memory) */
void  *bjxa_gpu_stream_create(void);		/* a non-blocking cudaStream_t */
int    bjxa_gpu_stream_destroy(void *cuda_stream);
int    bjxa_gpu_upload_async(void *dptr, const void *hptr, size_t bytes,
	    void *cuda_stream);
int    bjxa_gpu_download_async(void *hptr, const void *dptr, size_t bytes,
	    void *cuda_stream);

/*
 * Scatter of small fixed-size records on the device: record i is a 64-byte
 * slot of d_table, { uint64 offset; uint8 bytes[56] }, whose first rec_len
 * (<= 56) bytes go to (uint8_t *)dst + offset.  Used to put the 44-byte RIFF /
 * 32-byte XA headers in front of the data a plan produced, so that whole files
 * leave the device in one copy.  All pointers are device pointers.
 */
int    bjxa_gpu_scatter_async(void *dst, const void *d_table, uint32_t rec_len,
	    size_t n, void *cuda_stream);

/*
 * The host-buffer calls (bjxa_decode, bjxa_encode, bjxa_batch_*, bjxa_corpus_run)
 * keep their staging -- device arenas, pinned buffers, CUDA streams, plans -- with
 * the calling thread, on the device that was current at the time; a thread that
 * selects another device gets a fresh set there.  This gives it all back now (it
 * also happens by itself when the thread exits).
 */
void   bjxa_thread_release(void);

/* ---- whole files: a corpus in, a corpus out -------------------------------- */

/*
 * What `bjxa decode` / `bjxa encode` do to one file (src/bjxa_decode.c:38-93,
 * src/bjxa_encode.c: read the header, write the other header, run the block
 * loop, write the data), for n files that lie in one host arena, producing n
 * files in another: header parse and validation are bjxa_parse_header /
 * bjxa_parse_riff_header (src/libbjxa.c:396-453, 827-896), the emitted headers
 * are bjxa_dump_riff_header / bjxa_dump_header (:899-927, :479-503), byte for
 * byte.  Consecutive files travel to the device in ONE copy per chunk, the
 * produced files are assembled on the device (data by the block-loop kernels,
 * headers by bjxa_gpu_scatter_async) and come back in ONE copy per chunk;
 * chunks are pipelined over CUDA streams.  Pinned arenas (bjxa_host_alloc)
 * make the copies asynchronous.
 *
 * in:  in_off / in_len -- where file i lies in the input arena; files must be
 *      in ascending, non-overlapping order.  wav_to_xa is fastest when the PCM
 *      data of every WAV file is 16-byte aligned in the arena, (in_off + 44) %
 *      16 == 0: then a chunk travels in one copy; otherwise every file's PCM is
 *      copied on its own.
 * out: out_off / out_len -- where the produced file lies in the output arena
 *      and how long it is.  Decoded WAV files start at offsets = 4 (mod 16), so
 *      there are up to 15 bytes of padding between two files (content
 *      unspecified).
 *      error -- 0, or the errno the reference would have stopped at: EPROTO /
 *      EINVAL for a header it rejects, EIO for a file shorter than its header
 *      says, EPROTO for a bad block profile (out_len then covers the header and
 *      the blocks in front of the bad one, which is what the reference CLI has
 *      written when it gives up).
 * Returns 0 when every file was looked at (see error for each), -1 with errno
 * for a failure of the call itself (EFAULT, ENOBUFS: output arena too small --
 * bjxa_corpus_extent tells how much is needed --, ENODEV, ENOMEM).
 */
typedef struct bjxa_file_desc {
	uint64_t in_off, in_len;
	uint64_t out_off, out_len;
	int32_t  error;
	uint32_t blocks;	/* out: effective blocks processed */
	uint8_t  bits;		/* wav_to_xa in: 4, 6 or 8; xa_to_wav out: the file's */
	uint8_t  channels;	/* out */
	uint16_t rate;		/* out: samples per second */
	uint32_t reserved;
} bjxa_file_desc_t;

#define BJXA_CORPUS_XA_TO_WAV 0
#define BJXA_CORPUS_WAV_TO_XA 1

/* bytes of output arena bjxa_corpus_run will use for these files (headers are
 * read; nothing is decoded) */
int bjxa_corpus_extent(int kind, const void *in_arena, size_t in_bytes,
    const bjxa_file_desc_t *files, size_t n, uint64_t *out_bytes);
int bjxa_corpus_run(int kind, const void *in_arena, size_t in_bytes,
    void *out_arena, size_t out_bytes, bjxa_file_desc_t *files, size_t n);

/*
 * Contiguous shard of n streams for `rank` of `world` (one process per GPU),
 * balanced by bytes[] (per-stream work; NULL = by count).  No collective is
 * involved: shards are independent.
 */
int bjxa_shard_range(const uint64_t *bytes, size_t n, int rank, int world,
    size_t *first, size_t *count);

#ifdef __cplusplus
}
#endif
#endif /* BJXA_BATCH_H */
