/*
 * bjxa_internal.h -- what the objects of the library share besides the public
 * headers (nothing here is exported: src/libbjxa.map's `local: *`).
 */
#ifndef BJXA_INTERNAL_H
#define BJXA_INTERNAL_H

#include "../../include/bjxa_batch.h"

#ifdef __cplusplus
extern "C" {
#endif

/* xa_kernels.cu */
int  bjxa_gpu_current(void);		/* the calling thread's current device, -1 on error */
int  bjxa_small_call(int kind, bjxa_stream_desc_t *d, void *dst, const void *src);
void bjxa_small_release(void);

/* bjxa_corpus.c */
void bjxa_corpus_release(void);		/* the calling thread's pipeline buffers */

/* bjxa_host.c */
void bjxa_thread_cache_used(void);	/* arms the thread-exit release of the caches */

#ifdef __cplusplus
}
#endif
#endif
