/*
 * xa_tile.h -- the tile-level decode and encode algorithms.
 *
 * A "tile" is a run of consecutive effective blocks of ONE stream that one CTA
 * stages in shared memory: contiguous XA bytes on one side, contiguous
 * interleaved PCM on the other.  The code below is the per-thread body of each
 * phase between two CTA barriers; xa_kernels.cu wraps the phases in the
 * sm_100a kernel (bulk-async loads, mbarrier, __syncthreads), tests/emul wraps
 * the same phases in plain loops to single-step them on a CPU (tests only).
 *
 * Decode schedule (replaces the serial block loop of
 * /root/reference/src/libbjxa.c:629-658):
 *   phase A   every filter-0 ("cut") block is decoded at once -- its output
 *             does not depend on predictor state (k0 = k1 = 0,
 *             libbjxa.c:526,559) -- one thread per block-channel;
 *   rounds    blocks with filters 1..4 form chains behind a cut block; round r
 *             decodes the r-th block of every chain in the tile, compacted so
 *             that live chains fill whole warps; a chain's state is the last
 *             two samples of the previous block, read back from the staged
 *             output;
 *   carry     a chain that crosses a tile boundary takes its state from the
 *             previous tile of the stream through a 64-bit mailbox in global
 *             memory (epoch-tagged, so no clearing between launches); tiles
 *             are handed out by an atomic ticket in an order that puts every
 *             tile after its predecessor, so waiting is deadlock-free;
 *   store     staged PCM rows are interleaved (stereo) and written with
 *             16-byte stores; the last block of a stream is truncated to the
 *             PCM bytes owed (libbjxa.c:622-624,648).
 */
#ifndef XA_TILE_H
#define XA_TILE_H

#include "xa_core.h"

namespace xa {

/* ---- records shared by host and device --------------------------------- */

struct StreamDev {		/* one stream of a batch, device resident */
	uint64_t xa_off;	/* first XA block, bytes into the XA arena */
	uint64_t pcm_off;	/* PCM, bytes into the PCM arena; multiple of 16 */
	uint32_t blocks;	/* effective blocks to process */
	uint32_t pcm_len;	/* PCM bytes owed (decode) / available (encode) */
	uint32_t slot_base;	/* first carry mailbox of this stream */
	uint32_t reserved;
	int16_t  prev[2][2];	/* decode: predictor state on entry [ch][n-1,n-2] */
};

struct StreamRes {		/* decode results, device resident */
	int16_t  prev[2][2];	/* predictor state after the last block */
};

struct LiveRec;

/*
 * Relay form: a chain the tile's walker warp did not finish -- it ran past the
 * end of the strip, or the warp wound down while it was still going -- goes on
 * in the second pass (xa_walk.h) from this record: where its next item lies,
 * and the predictor state in front of it.
 */
struct RelayRec {
	uint32_t xa_lo, xa_hi;	/* arena address of the next item */
	uint32_t out_lo, out_hi;	/* arena address of its PCM */
	uint32_t left;		/* items of the stream behind it */
	uint32_t stream;	/* | chain channels of the item in front << 30 */
	uint32_t st[2];		/* per channel: state n-1 | n-2 << 16 */
};
#ifndef XA_RELAY_WIND
#define XA_RELAY_WIND 8
#endif
#ifndef XA_RELAY_WALKERS
#define XA_RELAY_WALKERS 1
#endif
constexpr uint32_t kRelayWind = XA_RELAY_WIND;	/* lanes at or below which a walker warp winds down */
constexpr uint32_t kRelayWalkers = XA_RELAY_WALKERS;	/* walker warps of a tile with many chains
						 * (2 and 3 measured slower: profiles/history_r2.md) */
constexpr uint32_t kRelayManyHeads = 48;	/* ... from this many chains */
/* records a tile can leave at most: each walker warp's stragglers + the strip's end */
constexpr uint32_t kRelayPerTile = kRelayWalkers * kRelayWind + 1;

/*
 * Segment form (xa_walk.h): every lane of a warp decodes kSegItems consecutive
 * items of one stream, whatever their filters; the state in front of the segment
 * is recomputed from at most kSegBack items before it, else waited for.
 */
#ifndef XA_SEG_ITEMS
#define XA_SEG_ITEMS 128
#endif
#ifndef XA_SEG_BACK
#define XA_SEG_BACK 48
#endif
constexpr uint32_t kSegItems = XA_SEG_ITEMS;
constexpr uint32_t kSegBack = XA_SEG_BACK;

struct TileEnt {		/* decode: NS strips of consecutive streams in issue order */
	uint32_t first;		/* index into order[] of the first strip's stream;
				 * encode: the stream itself */
	uint32_t count;		/* strips in this tile (<= NS); encode: unused */
	uint32_t j;		/* strip index inside each stream;
				 * encode: first effective block of the tile */
	uint32_t pad;
};

struct DecodeParams {
	const uint8_t *src;	/* XA arena */
	uint64_t src_bytes;
	uint8_t *dst;		/* PCM arena */
	const StreamDev *streams;
	StreamRes *results;
	uint32_t *first_bad;	/* per stream: lowest bad block-channel index */
	const TileEnt *tiles;
	uint32_t n_tiles;
	const uint32_t *order;	/* streams in issue order (longest first) */
	unsigned long long *carry;	/* [slot][2] mailboxes */
	unsigned long long *ticket;	/* preset to ~0 before every launch */
	uint32_t *fault;		/* set if a carry never arrived */
	unsigned long long carry_timeout_ns;
	uint32_t epoch;
	/* stereo: two tile forms are launched and the census decides which one
	 * runs; a kernel whose `want` differs from *choice returns at once */
	const uint32_t *choice;		/* NULL = run unconditionally */
	uint32_t want;
	/* split form (xa_walk.h): pass 1 lists the heads of chains instead of walking
	 * them.  0 = never, 1 = when the census says so (*choice == kFormSplit),
	 * 2 = always */
	uint32_t split;
	struct LiveRec *live;		/* one record per tile that has heads */
	uint32_t *live_count;		/* records written so far (pass 1), preset to ~0 */
	/* relay form: the tiles walk their own chains but hand the stragglers to
	 * the second pass.  0 = never, 1 = when the census says so, 2 = always */
	uint32_t relay;
	RelayRec *relay_recs;		/* kRelayPerTile per tile */
	uint32_t *relay_count;		/* records written so far, preset to ~0 */
};

struct EncodeParams {
	const uint8_t *src;	/* PCM arena */
	uint64_t src_bytes;
	uint8_t *dst;		/* XA arena */
	uint64_t dst_bytes;
	const StreamDev *streams;
	const TileEnt *tiles;
	uint32_t n_tiles;
	/* searching encoder only */
	StreamRes *results;	/* decoder state after the last block */
	const uint32_t *order;	/* the class's streams, longest first */
	uint32_t n_streams;
};

/* ---- environment shims -------------------------------------------------- */

/* how long a strip waits for its predecessor's state: wall-clock seconds
 * (BJXA_B200_CARRY_TIMEOUT_S overrides the default when a plan runs) */
#ifndef XA_CARRY_TIMEOUT_S
#define XA_CARRY_TIMEOUT_S 60
#endif

#if defined(__CUDA_ARCH__)
XA_HD void global_min_u32(uint32_t *p, uint32_t v) { atomicMin(p, v); }
XA_HD void mailbox_put(unsigned long long *p, unsigned long long v)
{
	/* the 64-bit word carries its own tag and payload: no ordering with any
	 * other store is needed, so no fence (a release store costs a MEMBAR that
	 * stalls the publishing warp -- and the CTA barrier behind it) */
	asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
XA_HD unsigned long long mailbox_get(const unsigned long long *p, uint32_t epoch,
    uint32_t *fault, unsigned long long timeout_ns)
{
	unsigned long long v, t0 = 0;
	/* The predecessor tile holds a lower ticket, so it is running or done
	 * and this wait is short.  It is still bounded -- by elapsed time on the
	 * device's clock, not by a count of polls, so that a context that is time
	 * sliced, preempted or held in a debugger does not fail a correct run: a
	 * kernel must never hang the device; on expiry the launch is flagged as
	 * failed (bjxa_plan_fetch: EIO, the PCM is not valid). */
	for (uint32_t spins = 0;; spins++) {
		asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
		if ((uint32_t)(v >> 32) == epoch)
			return v;
		__nanosleep(spins < 64 ? 32 : 512);
		if ((spins & 1023u) == 1023u) {
			unsigned long long now;
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
			if (t0 == 0)
				t0 = now;
			else if (now - t0 > timeout_ns)
				break;
		}
	}
	atomicExch(fault, 1u);
	return 0;
}
/* one look, no waiting (the pooled walkers come back to it on their next turn) */
XA_HD bool mailbox_try(const unsigned long long *p, uint32_t epoch, unsigned long long &v)
{
	asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return (uint32_t)(v >> 32) == epoch;
}
#else
XA_HD void global_min_u32(uint32_t *p, uint32_t v) { if (v < *p) *p = v; }
XA_HD void mailbox_put(unsigned long long *p, unsigned long long v) { *p = v; }
XA_HD unsigned long long mailbox_get(const unsigned long long *p, uint32_t epoch,
    uint32_t *, unsigned long long)
{
	/* the emulator runs tiles in ticket order: the value must be there */
	if ((uint32_t)(*p >> 32) != epoch)
		__builtin_trap();
	return *p;
}
XA_HD bool mailbox_try(const unsigned long long *p, uint32_t epoch, unsigned long long &v)
{
	if ((uint32_t)(*p >> 32) != epoch)
		__builtin_trap();
	v = *p;
	return true;
}
#endif

/* ---- decode -------------------------------------------------------------- */

/*
 * A decode tile is NS "strips" of SBQ block-channels each (NS * SBQ = TBQ):
 *   NS == 1   one long run of one stream -- used when a batch has few streams,
 *             so the parallelism must come from inside each stream;
 *   NS == 32  32 short runs of 32 different streams at the same position --
 *             used when a batch has many streams: even a stream without a
 *             single cut block (filters 1..4 throughout) then still gives every
 *             tile 32 independent chains, one per lane of a warp.
 * A strip is contiguous in both arenas; what a strip needs from the stream's
 * previous strip (the predictor state) travels through the carry mailbox.
 */
struct StripCtx {
	uint64_t a0;		/* first source byte rounded down to 16 */
	uint64_t out0;		/* first destination byte (16-byte aligned) */
	uint32_t stream;
	uint32_t first_eb;	/* first effective block of the strip */
	uint32_t nq;		/* block-channels in the strip */
	uint32_t in_base;	/* stage-buffer offset of the strip's first block */
	uint32_t in_need;	/* source bytes needed, counted from a0 */
	uint32_t bulk;		/* of which fetched by the bulk-async engine */
	uint32_t out_valid;	/* PCM bytes the strip owes */
	uint32_t slot;		/* carry mailbox index */
	uint32_t flags;
	uint32_t blocks;	/* effective blocks of the whole stream */
};
enum {
	kCtxFirst = 1u, kCtxLast = 2u, kCtxTail = 4u,
	/* split form: the block of channel 0 / 1 in front of the strip is a chain block */
	kCtxPrevShift = 3, kCtxPrev0 = 8u, kCtxPrev1 = 16u,
	kCtxEnd = 0x80000000u
};

template <int BITS, int CH, int TBQ, int NS>
struct DecGeom {
	static constexpr int kBits = BITS, kCh = CH, kTBQ = TBQ, kNS = NS;
	static constexpr int BS = block_bytes(BITS);
	static constexpr int SBQ = TBQ / NS;		/* block-channels per strip */
	static constexpr int SBE = SBQ / CH;		/* effective blocks per strip */
	/* one strip's slot in the stage buffer: payload + up to 15 bytes of
	 * misalignment in 16-byte units, plus a unit of slack for
	 * load_payload's one-word over-read */
	static constexpr int SLOT = ((SBQ * BS + 15 + 15) / 16) * 16 + 16;
	static constexpr int IN_BYTES = NS * SLOT;
	static_assert(TBQ % NS == 0 && SBQ % CH == 0 && SBQ >= 2, "strip geometry");
	static_assert((SBQ & (SBQ - 1)) == 0, "SBQ must be a power of two");
};

template <int BITS, int CH, int TBQ, int NS>
XA_HD void make_strip_ctx(StripCtx &c, const DecodeParams &p, uint32_t stream,
    uint32_t j, uint32_t strip, bool want_prev = false)
{
	typedef DecGeom<BITS, CH, TBQ, NS> G;
	const StreamDev &s = p.streams[stream];
	c.stream = stream;
	c.first_eb = j * G::SBE;
	uint32_t blocks = s.blocks;
	uint32_t rem = blocks - c.first_eb;
	uint32_t neb = rem < (uint32_t)G::SBE ? rem : (uint32_t)G::SBE;
	c.nq = neb * CH;
	uint64_t g0 = s.xa_off + (uint64_t)c.first_eb * (G::BS * CH);
	c.a0 = g0 & ~(uint64_t)15;
	uint32_t in_off = (uint32_t)(g0 - c.a0);
	c.in_base = strip * G::SLOT + in_off;
	c.in_need = in_off + c.nq * G::BS;
	/* whole 16-byte units that lie inside the arena */
	uint64_t end = c.a0 + ((c.in_need + 15u) & ~15u);
	uint64_t lim = p.src_bytes & ~(uint64_t)15;
	if (end > lim)
		end = lim > c.a0 ? lim : c.a0;
	c.bulk = (uint32_t)(end - c.a0);
	uint64_t pcm_done = (uint64_t)c.first_eb * (64 * CH);
	c.out0 = s.pcm_off + pcm_done;
	uint64_t owed = s.pcm_len > pcm_done ? s.pcm_len - pcm_done : 0;
	uint64_t full = (uint64_t)neb * (64 * CH);
	c.out_valid = (uint32_t)(owed < full ? owed : full);
	c.slot = s.slot_base + j;
	c.flags = (c.first_eb == 0 ? kCtxFirst : 0u) |
	    (c.first_eb + neb == blocks ? kCtxLast : 0u) |
	    (c.in_need > c.bulk ? kCtxTail : 0u);
	c.blocks = blocks;
	if (want_prev && c.first_eb != 0) {
		/* split form: which channels' chains run into this strip (their heads
		 * lie in a strip in front; xa_walk.h follows them to their ends) */
#pragma unroll
		for (int ch = 0; ch < CH; ch++)
			if (block_kind(p.src[g0 - (uint64_t)((CH - ch) * G::BS)]) == kChain)
				c.flags |= kCtxPrev0 << ch;
	}
}

/* the next free record of a list whose counter was preset to ~0 */
XA_HD uint32_t next_record(uint32_t *ctr)
{
#if defined(__CUDA_ARCH__)
	return atomicAdd(ctr, 1u) + 1u;
#else
	return ++(*ctr);
#endif
}

/* draw the next index from a shared-memory counter */
XA_HD uint32_t take_next(uint32_t *ctr)
{
#if defined(__CUDA_ARCH__)
	return atomicAdd(ctr, 1u);
#else
	return (*ctr)++;
#endif
}

/* ---- decode, direct form (mono) -------------------------------------------- */

template <int BITS, int CH, int TBQ, int NS, int STAGES>
struct DecSmem {
	typedef DecGeom<BITS, CH, TBQ, NS> G;

	alignas(16) uint8_t in[STAGES][G::IN_BYTES];
	StripCtx ctx[STAGES][NS];
	uint32_t tile_flags[STAGES];		/* kCtxEnd, kCtxTail (any strip) */
	uint32_t n_strips[STAGES];
	uint16_t heads[STAGES][TBQ];		/* first block of every chain of the tile */
	uint32_t n_heads[STAGES];
	uint32_t next_head[STAGES];		/* first chain no walker lane has taken yet */
	alignas(8) unsigned long long full[STAGES];	/* source bytes have landed */
	alignas(8) unsigned long long ready[STAGES];	/* ... and have been scanned */
	alignas(8) unsigned long long empty[STAGES];	/* every consumer warp is done */
};

/*
 * Work decomposition of a mono decode tile.  Nothing is staged and no thread
 * ever waits for another thread of its CTA:
 *
 *   scan     (producer warp, after the tile's bytes have landed) finds the
 *            head of every chain: a filter-1..4 block whose predecessor is not
 *            one (or lies in the stream's previous strip);
 *   units    one consumer thread per 16-byte unit of output (8 samples) of
 *            every CUT block (filter 0, or an invalid filter, which is decoded
 *            as filter 0 and reported): it unpacks the 2 * BITS/2 payload bytes
 *            it needs and stores the unit straight from registers --
 *            neighbouring threads write neighbouring 16 bytes;
 *   walkers  one consumer WARP per tile (the warps take turns), one lane per
 *            chain, lanes drawing the next chain as they finish one: predictor
 *            state in registers, starting from the last two samples of the cut
 *            block in front of the chain -- which the walker recomputes itself
 *            from that block's bytes -- or from the carry mailbox at the start
 *            of a strip; it stores its own 64-byte rows.
 *
 * A consumer warp that is done with its share of a tile arrives on the stage's
 * "empty" mbarrier and moves on; warps may drift up to STAGES tiles apart, so
 * one long chain delays nobody until the ring of source buffers wraps.
 */
template <int BITS, int TBQ, int NS, int STAGES, bool RELAY = false>
struct DecTile {
	typedef DecGeom<BITS, 1, TBQ, NS> G;
	typedef DecSmem<BITS, 1, TBQ, NS, STAGES> Smem;
	static constexpr int BS = G::BS;
	static constexpr int QB = BITS / 2;		/* payload bytes of 4 samples */
	static constexpr uint32_t SBQ = G::SBQ;		/* blocks per strip */
	static constexpr uint32_t SCAN = G::SBQ;	/* items the scanner looks at per strip */
	static constexpr int kLag = 1;
	static constexpr bool kStaged = false;
	static constexpr int kStages = STAGES;
	static constexpr int kMinCtas = 4;	/* per SM: caps the registers at 48 */
	static constexpr uint32_t UPS = SBQ * 4;	/* 16-byte units per strip */

	const DecodeParams &p;
	const uint8_t *in;		/* this tile's stage buffer */
	const StripCtx *ctx;		/* this tile's strips */
	const uint32_t n_strips;
	/* relay form: no carries, stragglers handed on.  A template parameter, not a
	 * flag: the plain form's kernels are capped at 48 registers, and code they
	 * never run must not cost them one */
	static constexpr bool relay = RELAY;

	XA_HD DecTile(const DecodeParams &p_, const Smem &sm_, int stage)
	    : p(p_), in(sm_.in[stage]), ctx(sm_.ctx[stage]), n_strips(sm_.n_strips[stage])
	{
	}

	/* relay form (NS == 1): the state in front of the strip's first block, from
	 * the arena -- the block there is a cut block, or the stream starts here */
	XA_HD void seed_from_arena(const StripCtx &c, int &p0, int &p1) const
	{
		if (c.flags & kCtxFirst) {
			p0 = p.streams[c.stream].prev[0][0];
			p1 = p.streams[c.stream].prev[0][1];
			return;
		}
		const uint8_t *b = p.src + (c.a0 + c.in_base);
		const int sh = 16 + (int)(b[-BS] & 15u);
		const uint32_t w = (uint32_t)b[-4] | (uint32_t)b[-3] << 8 | (uint32_t)b[-2] << 16 |
		    (uint32_t)b[-1] << 24;
		int x[4];
		quad_codes<BITS>(w >> (8 * (4 - QB)), x);
		p1 = x[2] >> sh;
		p0 = x[3] >> sh;
	}

	/* relay form: the chain goes on at block lq of the strip (lq == nq: the first
	 * block behind it) in the second pass */
	XA_HD void relay_put(const StripCtx &c, uint32_t lq, int p0, int p1) const
	{
		const uint64_t g = c.a0 + c.in_base + (uint64_t)lq * BS;
		const uint64_t o = c.out0 + (uint64_t)lq * 64u;
		RelayRec r;
		r.xa_lo = (uint32_t)g;
		r.xa_hi = (uint32_t)(g >> 32);
		r.out_lo = (uint32_t)o;
		r.out_hi = (uint32_t)(o >> 32);
		r.left = c.blocks - 1u - (c.first_eb + lq);
		r.stream = c.stream | 1u << 30;
		r.st[0] = (uint32_t)(uint16_t)p0 | (uint32_t)(uint16_t)p1 << 16;
		r.st[1] = 0;
		p.relay_recs[next_record(p.relay_count)] = r;
	}

	/* bytes past a strip's `bulk` fetched one by one (only at the arena's end) */
	XA_HD void load_tail(uint32_t tid, uint32_t nt, uint8_t *in_w) const
	{
		for (uint32_t st = 0; st < n_strips; st++) {
			const StripCtx &c = ctx[st];
			uint32_t slot0 = st * G::SLOT;
			for (uint32_t i = c.bulk + tid; i < c.in_need; i += nt)
				in_w[slot0 + i] = p.src[c.a0 + i];
		}
	}

	/* stage-buffer address of block lq of a strip */
	XA_HD uint32_t block_at(const StripCtx &c, uint32_t lq) const
	{
		return c.in_base + lq * BS;
	}

	/* 4 bytes starting at an arbitrary stage-buffer address */
	XA_HD uint32_t bytes_at(uint32_t at) const
	{
		const uint32_t *w = reinterpret_cast<const uint32_t *>(in) + (at >> 2);
		return funnel_r(w[0], w[1], (at & 3u) * 8u);
	}

	XA_HD void fetch_block(uint32_t at, uint32_t (&pw)[BITS]) const
	{
		uint32_t pay = at + 1;		/* first payload byte */
		const uint32_t *w = reinterpret_cast<const uint32_t *>(in) + (pay >> 2);
		load_payload<BITS>(pw, w, (pay & 3u) * 8u);
	}

	/* scanner, one strip: the chain channels of item q (a block: bit 0) */
	XA_HD uint32_t chain_mask(uint32_t q) const
	{
		return q < ctx[0].nq && block_kind(in[ctx[0].in_base + q * BS]) == kChain;
	}

	/* is block q of the tile (strip-major numbering) the head of a chain? */
	XA_HD bool is_head(uint32_t q) const
	{
		const StripCtx &c = ctx[q / SBQ];
		const uint32_t lq = q % SBQ;
		if (lq >= c.nq)
			return false;
		const uint32_t at = block_at(c, lq);
		return block_kind(in[at]) == kChain &&
		    (lq == 0 || block_kind(in[at - BS]) != kChain);
	}

	/* hand the stream's state to whoever continues it */
	XA_HD void publish(const StripCtx &c, int p0, int p1) const
	{
		if (c.flags & kCtxLast) {
			p.results[c.stream].prev[0][0] = (int16_t)p0;
			p.results[c.stream].prev[0][1] = (int16_t)p1;
		} else if (!relay) {
			unsigned long long v = ((unsigned long long)p.epoch << 32) |
			    ((unsigned long long)(uint16_t)p1 << 16) | (uint16_t)p0;
			mailbox_put(&p.carry[(uint64_t)c.slot * 2], v);
		}
	}

	XA_HD void carried_in(const StripCtx &c, int &p0, int &p1) const
	{
		if (c.flags & kCtxFirst) {
			p0 = p.streams[c.stream].prev[0][0];
			p1 = p.streams[c.stream].prev[0][1];
		} else {
			unsigned long long v = mailbox_get(
			    &p.carry[(uint64_t)(c.slot - 1) * 2], p.epoch, p.fault, p.carry_timeout_ns);
			p0 = (int16_t)(uint16_t)v;
			p1 = (int16_t)(uint16_t)(v >> 16);
		}
	}

	/* up to 16 bytes to global memory, honouring the strip's PCM length */
	XA_HD void put_unit(const StripCtx &c, uint32_t boff, const uint4 &v) const
	{
		uint8_t *d8 = p.dst + c.out0 + boff;
		if (boff + 16u <= c.out_valid) {
			*reinterpret_cast<uint4 *>(d8) = v;
		} else if (boff < c.out_valid) {
			/* the truncated last block of a stream (libbjxa.c:622-624) */
			const uint32_t w[4] = { v.x, v.y, v.z, v.w };
			uint16_t *d = reinterpret_cast<uint16_t *>(d8);
			uint32_t n16 = (c.out_valid - boff) / 2u;
			for (uint32_t k = 0; k < n16; k++)
				d[k] = (uint16_t)(w[k >> 1] >> (16u * (k & 1u)));
		}
	}

	/* the two quads (8 samples) that start at stage-buffer address a: their
	 * 2 * BITS/2 payload bytes, each quad in the low bytes of its word */
	XA_HD static void load_quads(const uint32_t *w, uint32_t sh, uint32_t &qa, uint32_t &qb)
	{
		if (BITS == 4) {
			qa = funnel_r(w[0], w[1], sh);		/* 4 bytes: 2 + 2 */
			qb = qa >> 16;
		} else if (BITS == 6) {
			const uint32_t t0 = funnel_r(w[0], w[1], sh);	/* bytes 0..3 */
			const uint32_t t1 = funnel_r(w[1], w[2], sh);	/* bytes 4..7 */
			qa = t0;				/* 3 bytes */
			qb = funnel_r(t0, t1, 24);		/* bytes 3..5 */
		} else {
			qa = funnel_r(w[0], w[1], sh);
			qb = funnel_r(w[1], w[2], sh);
		}
	}

	/* 8 samples from the payload window (aligned words w, bit offset sh) */
	XA_HD static uint4 unit_from(const uint32_t *w, uint32_t bit, uint32_t prof)
	{
		uint32_t qa, qb;
		int x[4], y[4];
		load_quads(w, bit, qa, qb);
		quad_codes<BITS>(qa, x);
		quad_codes<BITS>(qb, y);
		const int sh = 16 + (int)(prof & 15u);
		uint4 v;
		v.x = pack2(x[0] >> sh, x[1] >> sh);
		v.y = pack2(x[2] >> sh, x[3] >> sh);
		v.z = pack2(y[0] >> sh, y[1] >> sh);
		v.w = pack2(y[2] >> sh, y[3] >> sh);
		return v;
	}

	XA_HD uint4 decode_unit(uint32_t a, uint32_t prof) const
	{
		return unit_from(reinterpret_cast<const uint32_t *>(in) + (a >> 2), (a & 3u) * 8u, prof);
	}

	/* a block with a filter nibble set: false = a walker's, true = an invalid
	 * filter, which is recorded (by the k == 0 thread) and decoded as cut */
	XA_HD bool unit_uncut(const StripCtx &c, uint32_t lq, uint32_t k, uint32_t prof) const
	{
		const uint32_t f = prof >> 4;
		if (f - 1u < 4u)
			return false;
		if (k == 0)
			global_min_u32(&p.first_bad[c.stream], c.first_eb + lq);
		return true;
	}

	/* units: one thread per 16-byte unit of output of every cut block */
	XA_HD void phase_units(uint32_t tid, uint32_t nt) const
	{
		if (NS == 1 && nt % 16u == 0) {
			/*
			 * One strip: thread t owns units t, t + nt, ...: always the same
			 * quad pair of blocks nt / 4 apart.  That is (nt / 4) * BS bytes
			 * of stage buffer, a multiple of 4, so the word alignment of the
			 * thread's payload window never changes and the loop advances two
			 * pointers and a counter.  Whole units first; the (at most one)
			 * ragged unit and the carry out of the strip come after the loop.
			 */
			const StripCtx &c = ctx[0];
			const uint32_t total = c.nq * 4u;
			const uint32_t nfull = c.out_valid / 16u < total ? c.out_valid / 16u : total;
			const uint32_t k = tid & 3u, stepb = (nt >> 2) * BS;
			const uint32_t at = c.in_base + (tid >> 2) * BS;
			const uint32_t a = at + 1 + k * (2 * QB);
			const uint8_t *pin = in + at;
			const uint32_t *w = reinterpret_cast<const uint32_t *>(in + (a & ~3u));
			const uint32_t bit = (a & 3u) * 8u;
			uint8_t *const out = p.dst + c.out0;
			uint32_t u = tid;
			for (; u < nfull; u += nt, pin += stepb, w += stepb / 4) {
				const uint32_t prof = pin[0];
				if ((prof & 0xf0u) != 0 && !unit_uncut(c, u >> 2, k, prof))
					continue;
				*reinterpret_cast<uint4 *>(out + (uint64_t)u * 16u) = unit_from(w, bit, prof);
			}
			if (u < total) {
				const uint32_t prof = pin[0];
				if ((prof & 0xf0u) == 0 || unit_uncut(c, u >> 2, k, prof))
					put_unit(c, u * 16u, unit_from(w, bit, prof));
			}
			if (total != 0 && (total - 1u) % nt == tid) {
				/* the last block of the strip hands its last two samples on */
				const uint32_t lat = block_at(c, c.nq - 1u);
				const uint32_t prof = in[lat];
				if (block_kind(prof) != kChain) {
					const uint4 v = decode_unit(lat + 1 + 3 * (2 * QB), prof);
					publish(c, (int)(int16_t)(v.w >> 16), (int)(int16_t)(v.w & 0xffffu));
				}
			}
			return;
		}
		const uint32_t total = n_strips * UPS;
		for (uint32_t u = tid; u < total; u += nt) {
			const uint32_t st = u / UPS, lu = u % UPS;
			const StripCtx &c = ctx[st];
			const uint32_t lq = lu / 4u, k = lu % 4u;
			if (lq >= c.nq)
				continue;
			const uint32_t at0 = block_at(c, lq);
			const uint32_t prof = in[at0];
			const int kind = block_kind(prof);
			if (kind == kChain)
				continue;
			if (k == 0 && kind == kBad)
				global_min_u32(&p.first_bad[c.stream], c.first_eb + lq);
			const uint4 v = decode_unit(at0 + 1 + k * (2 * QB), prof);
			put_unit(c, lu * 16u, v);
			if (k == 3 && lq + 1 >= c.nq)
				publish(c, (int)(int16_t)(v.w >> 16), (int)(int16_t)(v.w & 0xffffu));
		}
	}

	/*
	 * One warp walks all the chains of a tile: lane l starts with chain l and,
	 * whenever its chain ends, draws the next one nobody has taken (*next, preset
	 * to 32 by the scanner).  One flat loop -- pick up a chain if idle, then
	 * decode ONE block -- keeps the lanes that still have blocks converged on the
	 * block decode while the others pick up their next chain; a loop over whole
	 * chains would idle every lane until the longest of 32 chains is through.
	 */
	XA_HD void phase_walk_warp(uint32_t lane, const uint16_t *heads, uint32_t n,
	    uint32_t *next) const
	{
		uint32_t i = lane, lq = 0, at = 0;
		const StripCtx *c = ctx;
		int p0 = 0, p1 = 0;
		bool have = false;
		for (;;) {
			if (!have) {
				if (i >= n)
					break;
				const uint32_t q = heads[i];
				c = &ctx[q / SBQ];
				lq = q % SBQ;
				at = block_at(*c, lq);
				if (lq == 0) {
					carried_in(*c, p0, p1);
				} else {
					/* the block in front is a cut block: its last two
					 * samples are codes 2 and 3 of its last quad */
					const uint32_t pa = at - BS;
					const int sh = 16 + (int)(in[pa] & 15u);
					int x[4];
					quad_codes<BITS>(bytes_at(pa + 1 + 7 * QB), x);
					p1 = x[2] >> sh;
					p0 = x[3] >> sh;
				}
				have = true;
			}
			uint32_t pw[BITS], o[16];
			fetch_block(at, pw);
			decode_block_chain<BITS>(o, pw, in[at], p0, p1);
#pragma unroll
			for (int j = 0; j < 4; j++) {
				uint4 v;
				v.x = o[4 * j]; v.y = o[4 * j + 1];
				v.z = o[4 * j + 2]; v.w = o[4 * j + 3];
				put_unit(*c, (lq * 4u + (uint32_t)j) * 16u, v);
			}
			bool more = false;
			if (lq + 1 >= c->nq) {
				publish(*c, p0, p1);
			} else {
				lq++;
				at += BS;
				more = block_kind(in[at]) == kChain;
			}
			if (!more) {
				have = false;
				i = take_next(next);
			}
		}
	}

	/*
	 * The same walk cut in two -- "start at a head" and "decode ONE block" -- for
	 * the pooled form (xa_decode_pool_kernel), whose walker lanes hold chains of
	 * different tiles of the ring at the same time.
	 */
	struct Walk {
		const StripCtx *c;
		uint32_t lq, at;
		uint32_t need;		/* the carry has not arrived yet */
		int p0, p1;
	};

	/* the state carried into the strip, if it is there: never waits, because
	 * the lane that has to produce it may sit in the same warp */
	XA_HD void walk_carry(Walk &w) const
	{
		if (w.c->flags & kCtxFirst) {
			w.p0 = p.streams[w.c->stream].prev[0][0];
			w.p1 = p.streams[w.c->stream].prev[0][1];
			w.need = 0;
			return;
		}
		unsigned long long v;
		if (mailbox_try(&p.carry[(uint64_t)(w.c->slot - 1) * 2], p.epoch, v)) {
			w.p0 = (int16_t)(uint16_t)v;
			w.p1 = (int16_t)(uint16_t)(v >> 16);
			w.need = 0;
		}
	}

	/* a carry that never arrives: flag the launch (bjxa_plan_fetch: EIO), go on */
	XA_HD void walk_give_up(Walk &w) const
	{
#if defined(__CUDA_ARCH__)
		atomicExch(p.fault, 1u);
#endif
		w.p0 = w.p1 = 0;
		w.need = 0;
	}

	XA_HD void walk_begin(Walk &w, uint32_t q) const
	{
		w.c = &ctx[q / SBQ];
		w.lq = q % SBQ;
		w.at = block_at(*w.c, w.lq);
		w.need = 0;
		if (w.lq == 0 && relay) {
			seed_from_arena(*w.c, w.p0, w.p1);
		} else if (w.lq == 0) {
			w.need = 1;
			walk_carry(w);
		} else {
			/* the block in front is a cut block (see phase_walk_warp) */
			const uint32_t pa = w.at - BS;
			const int sh = 16 + (int)(in[pa] & 15u);
			int x[4];
			quad_codes<BITS>(bytes_at(pa + 1 + 7 * QB), x);
			w.p1 = x[2] >> sh;
			w.p0 = x[3] >> sh;
		}
	}

	/* one block of the chain; false when it was the chain's last */
	XA_HD bool walk_block(Walk &w) const
	{
		uint32_t pw[BITS], o[16];
		fetch_block(w.at, pw);
		decode_block_chain<BITS>(o, pw, in[w.at], w.p0, w.p1);
#pragma unroll
		for (int j = 0; j < 4; j++) {
			uint4 v;
			v.x = o[4 * j]; v.y = o[4 * j + 1];
			v.z = o[4 * j + 2]; v.w = o[4 * j + 3];
			put_unit(*w.c, (w.lq * 4u + (uint32_t)j) * 16u, v);
		}
		if (w.lq + 1 >= w.c->nq) {
			publish(*w.c, w.p0, w.p1);
			/* relay form: whether the strip behind goes on with this chain is for
			 * the second pass to see */
			if (relay && !(w.c->flags & kCtxLast))
				relay_put(*w.c, w.c->nq, w.p0, w.p1);
			return false;
		}
		w.lq++;
		w.at += BS;
		return block_kind(in[w.at]) == kChain;
	}

	/* relay form: the walker warp winds down; the chain goes on at its next block */
	XA_HD void walk_hand_on(const Walk &w) const
	{
		relay_put(*w.c, w.lq, w.p0, w.p1);
	}
};

/* ---- decode, direct form (stereo) ------------------------------------------ */
/*
 * The same barrier-free decomposition for two channels.  The unit of work is
 * the EFFECTIVE block (left block + right block = 32 frames = 128 bytes of
 * PCM):
 *   units    one thread per 16-byte unit (4 frames) of every effective block
 *            whose two blocks are both cut blocks: quad k of the left and of
 *            the right block, interleaved in registers, one 128-bit store;
 *   walkers  one thread per maximal run of effective blocks in which at least
 *            one channel is a chain block.  The walker carries the predictor
 *            state of BOTH channels and decodes them sample by sample side by
 *            side (two independent dependency chains: tools/lat_bench.cu
 *            measures 19 instead of 34 cycles per sample for two interleaved
 *            chains); a cut block inside a run simply has k0 = k1 = 0.  It
 *            writes whole interleaved 16-byte units, so nothing ever has to
 *            be written around the other channel.
 */
template <int BITS, int TBQ, int NS, int STAGES, bool RELAY = false>
struct DecTileStereo {
	typedef DecGeom<BITS, 2, TBQ, NS> G;
	typedef DecSmem<BITS, 2, TBQ, NS, STAGES> Smem;
	static constexpr int BS = G::BS;
	static constexpr int QB = BITS / 2;
	static constexpr uint32_t SBE = G::SBE;		/* effective blocks per strip */
	static constexpr uint32_t SCAN = G::SBE;
	static constexpr int kLag = 1;
	static constexpr bool kStaged = false;
	static constexpr int kStages = STAGES;
	/* CTAs per SM the registers must allow: 4 (48 registers) for 4-bit streams,
	 * whose units are the most work per byte; 3 (64 registers) otherwise, where
	 * the cap costs more in the walkers than the fourth CTA brings (measured) */
	static constexpr int kMinCtas = BITS == 4 ? 4 : 3;
	static constexpr uint32_t UPS = SBE * 8;	/* 16-byte units per strip */

	const DecodeParams &p;
	const uint8_t *in;
	const StripCtx *ctx;
	const uint32_t n_strips;
	static constexpr bool relay = RELAY;	/* see DecTile */

	XA_HD DecTileStereo(const DecodeParams &p_, const Smem &sm_, int stage)
	    : p(p_), in(sm_.in[stage]), ctx(sm_.ctx[stage]), n_strips(sm_.n_strips[stage])
	{
	}

	/* relay form (NS == 1): both channels' states in front of the strip's first
	 * effective block, from the arena (a chain channel's block there is a cut
	 * block; the other channel's value is not used) */
	XA_HD void seed_from_arena(const StripCtx &c, int (&p0)[2], int (&p1)[2]) const
	{
#pragma unroll
		for (int ch = 0; ch < 2; ch++) {
			if (c.flags & kCtxFirst) {
				p0[ch] = p.streams[c.stream].prev[ch][0];
				p1[ch] = p.streams[c.stream].prev[ch][1];
				continue;
			}
			const uint8_t *b = p.src + (c.a0 + c.in_base) - (2 - ch) * BS;
			const int sh = 16 + (int)(b[0] & 15u);
			const uint32_t w = (uint32_t)b[BS - 4] | (uint32_t)b[BS - 3] << 8 |
			    (uint32_t)b[BS - 2] << 16 | (uint32_t)b[BS - 1] << 24;
			int x[4];
			quad_codes<BITS>(w >> (8 * (4 - QB)), x);
			p1[ch] = x[2] >> sh;
			p0[ch] = x[3] >> sh;
		}
	}

	/* relay form: the run goes on at effective block eb of the strip (eb == the
	 * strip's length: the first one behind it) in the second pass, if that item
	 * continues a chain of `m_front`, the chain channels of the item in front */
	XA_HD void relay_put(const StripCtx &c, uint32_t eb, uint32_t m_front, const int (&p0)[2],
	    const int (&p1)[2]) const
	{
		const uint64_t g = c.a0 + c.in_base + (uint64_t)eb * (2 * BS);
		const uint64_t o = c.out0 + (uint64_t)eb * 128u;
		RelayRec r;
		r.xa_lo = (uint32_t)g;
		r.xa_hi = (uint32_t)(g >> 32);
		r.out_lo = (uint32_t)o;
		r.out_hi = (uint32_t)(o >> 32);
		r.left = c.blocks - 1u - (c.first_eb + eb);
		r.stream = c.stream | m_front << 30;
		r.st[0] = (uint32_t)(uint16_t)p0[0] | (uint32_t)(uint16_t)p1[0] << 16;
		r.st[1] = (uint32_t)(uint16_t)p0[1] | (uint32_t)(uint16_t)p1[1] << 16;
		p.relay_recs[next_record(p.relay_count)] = r;
	}

	XA_HD void load_tail(uint32_t tid, uint32_t nt, uint8_t *in_w) const
	{
		for (uint32_t st = 0; st < n_strips; st++) {
			const StripCtx &c = ctx[st];
			uint32_t slot0 = st * G::SLOT;
			for (uint32_t i = c.bulk + tid; i < c.in_need; i += nt)
				in_w[slot0 + i] = p.src[c.a0 + i];
		}
	}

	/* stage-buffer address of the LEFT block of effective block eb */
	XA_HD uint32_t eb_at(const StripCtx &c, uint32_t eb) const
	{
		return c.in_base + eb * (2 * BS);
	}

	XA_HD uint32_t bytes_at(uint32_t at) const
	{
		const uint32_t *w = reinterpret_cast<const uint32_t *>(in) + (at >> 2);
		return funnel_r(w[0], w[1], (at & 3u) * 8u);
	}

	XA_HD void fetch_block(uint32_t at, uint32_t (&pw)[BITS]) const
	{
		uint32_t pay = at + 1;
		const uint32_t *w = reinterpret_cast<const uint32_t *>(in) + (pay >> 2);
		load_payload<BITS>(pw, w, (pay & 3u) * 8u);
	}

	/* which channels of effective block eb (inside the strip) are chain blocks:
	 * bit 0 left, bit 1 right */
	XA_HD uint32_t chains(const StripCtx &c, uint32_t eb) const
	{
		const uint32_t at = eb_at(c, eb);
		return (uint32_t)(block_kind(in[at]) == kChain) |
		    (uint32_t)(block_kind(in[at + BS]) == kChain) << 1;
	}

	XA_HD uint32_t chain_mask(uint32_t q) const	/* scanner, one strip */
	{
		return q * 2 < ctx[0].nq ? chains(ctx[0], q) : 0u;
	}

	/*
	 * A walker starts at an effective block with a chain block whose chain
	 * channels all follow a block that is not a chain block (its last two
	 * samples can be read off its bytes), and goes on for as long as the next
	 * effective block continues a chain in either channel.  Splitting runs
	 * wherever the state can be recomputed keeps them short: with isolated
	 * chain blocks (the usual case) every run is one effective block and the
	 * walker lanes of a warp stay converged.   q = strip * SBE + eb
	 */
	XA_HD bool is_head(uint32_t q) const
	{
		const StripCtx &c = ctx[q / SBE];
		const uint32_t eb = q % SBE;
		if (eb * 2 >= c.nq)
			return false;
		const uint32_t m = chains(c, eb);
		return m != 0 && (eb == 0 || (m & chains(c, eb - 1)) == 0);
	}

	XA_HD void publish(const StripCtx &c, uint32_t ch, int p0, int p1) const
	{
		if (c.flags & kCtxLast) {
			p.results[c.stream].prev[ch][0] = (int16_t)p0;
			p.results[c.stream].prev[ch][1] = (int16_t)p1;
		} else if (!relay) {
			unsigned long long v = ((unsigned long long)p.epoch << 32) |
			    ((unsigned long long)(uint16_t)p1 << 16) | (uint16_t)p0;
			mailbox_put(&p.carry[(uint64_t)c.slot * 2 + ch], v);
		}
	}

	XA_HD void carried_in(const StripCtx &c, uint32_t ch, int &p0, int &p1) const
	{
		if (c.flags & kCtxFirst) {
			p0 = p.streams[c.stream].prev[ch][0];
			p1 = p.streams[c.stream].prev[ch][1];
		} else {
			unsigned long long v = mailbox_get(
			    &p.carry[(uint64_t)(c.slot - 1) * 2 + ch], p.epoch, p.fault, p.carry_timeout_ns);
			p0 = (int16_t)(uint16_t)v;
			p1 = (int16_t)(uint16_t)(v >> 16);
		}
	}

	XA_HD void put_unit(const StripCtx &c, uint32_t boff, const uint4 &v) const
	{
		uint8_t *d8 = p.dst + c.out0 + boff;
		if (boff + 16u <= c.out_valid) {
			*reinterpret_cast<uint4 *>(d8) = v;
		} else if (boff < c.out_valid) {
			const uint32_t w[4] = { v.x, v.y, v.z, v.w };
			uint16_t *d = reinterpret_cast<uint16_t *>(d8);
			uint32_t n16 = (c.out_valid - boff) / 2u;
			for (uint32_t k = 0; k < n16; k++)
				d[k] = (uint16_t)(w[k >> 1] >> (16u * (k & 1u)));
		}
	}

	/* quad k of both blocks of the effective block at `at`, as 4 frames */
	XA_HD uint4 decode_unit(uint32_t at, uint32_t k, uint32_t profl, uint32_t profr) const
	{
		int x[4], y[4];
		const uint32_t a = at + 1 + k * QB;
		quad_codes<BITS>(bytes_at(a), x);
		quad_codes<BITS>(bytes_at(a + BS), y);
		const int shl = 16 + (int)(profl & 15u), shr = 16 + (int)(profr & 15u);
		uint4 v;
		v.x = pack2(x[0] >> shl, y[0] >> shr);
		v.y = pack2(x[1] >> shl, y[1] >> shr);
		v.z = pack2(x[2] >> shl, y[2] >> shr);
		v.w = pack2(x[3] >> shl, y[3] >> shr);
		return v;
	}

	/* one unit of a tile of several strips (the one-strip loop is phase_units') */
	XA_HD void unit_body(const StripCtx &c, uint32_t eb, uint32_t k, uint32_t at,
	    uint32_t u_in_strip) const
	{
		const uint32_t profl = in[at], profr = in[at + BS];
		const uint32_t fl = profl >> 4, fr = profr >> 4;
		if (fl - 1u < 4u || fr - 1u < 4u)
			return;			/* a walker's */
		if (k == 0) {
			if (fl >= 5u)
				global_min_u32(&p.first_bad[c.stream], (c.first_eb + eb) * 2);
			if (fr >= 5u)
				global_min_u32(&p.first_bad[c.stream], (c.first_eb + eb) * 2 + 1);
		}
		const uint4 v = decode_unit(at, k, profl, profr);
		put_unit(c, u_in_strip * 16u, v);
		if (k == 7 && (eb + 1) * 2 >= c.nq) {
			/* frames 30 and 31: v.z = L30 | R30 << 16, v.w = L31 | R31 << 16 */
			publish(c, 0, (int)(int16_t)(v.w & 0xffffu), (int)(int16_t)(v.z & 0xffffu));
			publish(c, 1, (int)(int16_t)(v.w >> 16), (int)(int16_t)(v.z >> 16));
		}
	}

	/* the four frames of quad k from the two payload windows of an effective block */
	XA_HD static uint4 frames(uint32_t lw, uint32_t rw, uint32_t profl, uint32_t profr)
	{
		int x[4], y[4];
		quad_codes<BITS>(lw, x);
		quad_codes<BITS>(rw, y);
		const int shl = 16 + (int)(profl & 15u), shr = 16 + (int)(profr & 15u);
		uint4 v;
		v.x = pack2(x[0] >> shl, y[0] >> shr);
		v.y = pack2(x[1] >> shl, y[1] >> shr);
		v.z = pack2(x[2] >> shl, y[2] >> shr);
		v.w = pack2(x[3] >> shl, y[3] >> shr);
		return v;
	}

	/* an effective block with a filter nibble set: false = a walker's, true = an
	 * invalid filter, which is recorded (by the k == 0 thread) and decoded as cut */
	XA_HD bool unit_uncut(const StripCtx &c, uint32_t eb, uint32_t k, uint32_t profl,
	    uint32_t profr) const
	{
		const uint32_t fl = profl >> 4, fr = profr >> 4;
		if (fl - 1u < 4u || fr - 1u < 4u)
			return false;
		if (k == 0) {
			if (fl >= 5u)
				global_min_u32(&p.first_bad[c.stream], (c.first_eb + eb) * 2);
			if (fr >= 5u)
				global_min_u32(&p.first_bad[c.stream], (c.first_eb + eb) * 2 + 1);
		}
		return true;
	}

	XA_HD void phase_units(uint32_t tid, uint32_t nt) const
	{
		if (NS == 1 && nt % 16u == 0) {
			/*
			 * One strip: thread t owns units t, t + nt, ...  The byte distance
			 * between two of them in the stage buffer, (nt / 8) * 2 * BS, is a
			 * multiple of 4, so the word alignment of a thread's payload
			 * windows never changes and the loop only advances three
			 * pointers; whole units first, the (at most one) ragged unit and
			 * the carry out of the strip after the loop.
			 */
			const StripCtx &c = ctx[0];
			const uint32_t total = (c.nq / 2u) * 8u;
			const uint32_t nfull = c.out_valid / 16u < total ? c.out_valid / 16u : total;
			const uint32_t k = tid & 7u, stepb = (nt >> 3) * (2 * BS);
			const uint32_t at = c.in_base + (tid >> 3) * (2 * BS);
			const uint32_t al = at + 1 + k * QB, ar = al + BS;
			const uint8_t *pin = in + at;
			const uint32_t *wl = reinterpret_cast<const uint32_t *>(in + (al & ~3u));
			const uint32_t *wr = reinterpret_cast<const uint32_t *>(in + (ar & ~3u));
			const uint32_t bl = (al & 3u) * 8u, br = (ar & 3u) * 8u;
			uint8_t *const out = p.dst + c.out0;
			uint32_t u = tid;
			for (; u < nfull; u += nt, pin += stepb, wl += stepb / 4, wr += stepb / 4) {
				const uint32_t profl = pin[0], profr = pin[BS];
				if (((profl | profr) & 0xf0u) != 0 && !unit_uncut(c, u >> 3, k, profl, profr))
					continue;
				*reinterpret_cast<uint4 *>(out + (uint64_t)u * 16u) =
				    frames(funnel_r(wl[0], wl[1], bl), funnel_r(wr[0], wr[1], br), profl, profr);
			}
			if (u < total) {
				const uint32_t profl = pin[0], profr = pin[BS];
				if (((profl | profr) & 0xf0u) == 0 || unit_uncut(c, u >> 3, k, profl, profr))
					put_unit(c, u * 16u, frames(funnel_r(wl[0], wl[1], bl),
					    funnel_r(wr[0], wr[1], br), profl, profr));
			}
			if (total != 0 && (total - 1u) % nt == tid) {
				/* last unit of the strip (k == 7): frames 30, 31 go on as the carry */
				const uint32_t lat = eb_at(c, total / 8u - 1u);
				const uint32_t profl = in[lat], profr = in[lat + BS];
				if (block_kind(profl) != kChain && block_kind(profr) != kChain) {
					const uint4 v = decode_unit(lat, 7, profl, profr);
					publish(c, 0, (int)(int16_t)(v.w & 0xffffu), (int)(int16_t)(v.z & 0xffffu));
					publish(c, 1, (int)(int16_t)(v.w >> 16), (int)(int16_t)(v.z >> 16));
				}
			}
			return;
		}
		const uint32_t total = n_strips * UPS;
		for (uint32_t u = tid; u < total; u += nt) {
			const uint32_t st = u / UPS, lu = u % UPS;
			const StripCtx &c = ctx[st];
			const uint32_t eb = lu / 8u, k = lu % 8u;
			if (eb * 2 >= c.nq)
				continue;
			unit_body(c, eb, k, eb_at(c, eb), lu);
		}
	}

	/*
	 * One effective block of a run: both channels side by side -- two
	 * independent dependency chains (a cut block inside a run simply has
	 * k0 = k1 = 0) -- stored as eight interleaved 16-byte units.
	 */
	XA_HD void decode_pair(const StripCtx &c, uint32_t eb, uint32_t at, uint32_t profl,
	    uint32_t profr, int (&p0)[2], int (&p1)[2]) const
	{
		uint32_t pl[BITS], pr[BITS];
		fetch_block(at, pl);
		fetch_block(at + BS, pr);
		const int shl = 16 + (int)(profl & 15u), shr = 16 + (int)(profr & 15u);
		const int k0l = gain_k0(profl >> 4), k1l = gain_k1(profl >> 4);
		const int k0r = gain_k0(profr >> 4), k1r = gain_k1(profr >> 4);
		/* the plain step (sample_chain), not the biased one of the mono walkers:
		 * two chains side by side already hide its latency, and the four extra
		 * constants cost this register-capped kernel more than the shorter
		 * dependent path brings (6-bit P1 84.5 -> 85.5 %, C20 50.7 -> 53.6 %) */
#pragma unroll
		for (int j = 0; j < 8; j++) {
			int l[4], r[4];
#pragma unroll
			for (int k = 0; k < 4; k++) {
				l[k] = sample_chain(top_code<BITS>(pl, 4 * j + k), shl, k0l, k1l,
				    p0[0], p1[0]);
				r[k] = sample_chain(top_code<BITS>(pr, 4 * j + k), shr, k0r, k1r,
				    p0[1], p1[1]);
			}
			uint4 v;
			v.x = pack2(l[0], r[0]);
			v.y = pack2(l[1], r[1]);
			v.z = pack2(l[2], r[2]);
			v.w = pack2(l[3], r[3]);
			put_unit(c, (eb * 8u + (uint32_t)j) * 16u, v);
		}
	}

	/*
	 * One warp walks all the runs of a tile, lanes drawing the next run as they
	 * finish one (see DecTile::phase_walk_warp).  A run starts at the effective
	 * block heads[i] = strip * SBE + eb and is decoded one effective block per
	 * turn of the loop.
	 */
	XA_HD void phase_walk_warp(uint32_t lane, const uint16_t *heads, uint32_t n,
	    uint32_t *next) const
	{
		uint32_t i = lane, eb = 0, at = 0, m = 0;
		const StripCtx *c = ctx;
		int p0[2] = { 0, 0 }, p1[2] = { 0, 0 };
		bool have = false;
		for (;;) {
			if (!have) {
				if (i >= n)
					break;
				const uint32_t q = heads[i];
				c = &ctx[q / SBE];
				eb = q % SBE;
				at = eb_at(*c, eb);
				m = chains(*c, eb);
				if (eb == 0) {
					/* a channel that starts with a cut block needs no history */
					if (m & 1u)
						carried_in(*c, 0, p0[0], p1[0]);
					if (m & 2u)
						carried_in(*c, 1, p0[1], p1[1]);
				} else {
					/* the blocks in front of this run's chain blocks are not
					 * chain blocks: their last two samples = codes 2, 3 of
					 * the last quad (read for both channels; only a chain
					 * channel's matter) */
#pragma unroll
					for (int ch = 0; ch < 2; ch++) {
						const uint32_t pa = at - 2 * BS + ch * BS;
						const int sh = 16 + (int)(in[pa] & 15u);
						int x[4];
						quad_codes<BITS>(bytes_at(pa + 1 + 7 * QB), x);
						p1[ch] = x[2] >> sh;
						p0[ch] = x[3] >> sh;
					}
				}
				have = true;
			}
			const uint32_t profl = in[at], profr = in[at + BS];
			if (profl >> 4 >= 5u)
				global_min_u32(&p.first_bad[c->stream], (c->first_eb + eb) * 2);
			if (profr >> 4 >= 5u)
				global_min_u32(&p.first_bad[c->stream], (c->first_eb + eb) * 2 + 1);
			decode_pair(*c, eb, at, profl, profr, p0, p1);
			bool more = false;
			if ((eb + 1) * 2 >= c->nq) {
				publish(*c, 0, p0[0], p1[0]);
				publish(*c, 1, p0[1], p1[1]);
			} else {
				eb++;
				at += 2 * BS;
				const uint32_t nm = chains(*c, eb);
				more = (nm & m) != 0;	/* else all cut, or the head of another run */
				m = nm;
			}
			if (!more) {
				have = false;
				i = take_next(next);
			}
		}
	}

	/* the same walk cut in two for the pooled form (see DecTile::Walk) */
	struct Walk {
		const StripCtx *c;
		uint32_t eb, at, m;
		uint32_t need;		/* channels whose carry has not arrived yet */
		int p0[2], p1[2];
	};

	XA_HD void walk_carry(Walk &w) const
	{
#pragma unroll
		for (uint32_t ch = 0; ch < 2; ch++) {
			if (!(w.need >> ch & 1u))
				continue;
			if (w.c->flags & kCtxFirst) {
				w.p0[ch] = p.streams[w.c->stream].prev[ch][0];
				w.p1[ch] = p.streams[w.c->stream].prev[ch][1];
				w.need &= ~(1u << ch);
				continue;
			}
			unsigned long long v;
			if (mailbox_try(&p.carry[(uint64_t)(w.c->slot - 1) * 2 + ch], p.epoch, v)) {
				w.p0[ch] = (int16_t)(uint16_t)v;
				w.p1[ch] = (int16_t)(uint16_t)(v >> 16);
				w.need &= ~(1u << ch);
			}
		}
	}

	XA_HD void walk_give_up(Walk &w) const
	{
#if defined(__CUDA_ARCH__)
		atomicExch(p.fault, 1u);
#endif
		w.need = 0;
	}

	XA_HD void walk_begin(Walk &w, uint32_t q) const
	{
		w.c = &ctx[q / SBE];
		w.eb = q % SBE;
		w.at = eb_at(*w.c, w.eb);
		w.m = chains(*w.c, w.eb);
		w.p0[0] = w.p0[1] = w.p1[0] = w.p1[1] = 0;
		w.need = 0;
		if (w.eb == 0 && relay) {
			seed_from_arena(*w.c, w.p0, w.p1);
		} else if (w.eb == 0) {
			/* a channel that starts with a cut block needs no history */
			w.need = w.m;
			walk_carry(w);
		} else {
#pragma unroll
			for (int ch = 0; ch < 2; ch++) {
				const uint32_t pa = w.at - 2 * BS + ch * BS;
				const int sh = 16 + (int)(in[pa] & 15u);
				int x[4];
				quad_codes<BITS>(bytes_at(pa + 1 + 7 * QB), x);
				w.p1[ch] = x[2] >> sh;
				w.p0[ch] = x[3] >> sh;
			}
		}
	}

	/* one effective block of the run; false when it was the run's last */
	XA_HD bool walk_block(Walk &w) const
	{
		const uint32_t profl = in[w.at], profr = in[w.at + BS];
		if (profl >> 4 >= 5u)
			global_min_u32(&p.first_bad[w.c->stream], (w.c->first_eb + w.eb) * 2);
		if (profr >> 4 >= 5u)
			global_min_u32(&p.first_bad[w.c->stream], (w.c->first_eb + w.eb) * 2 + 1);
		decode_pair(*w.c, w.eb, w.at, profl, profr, w.p0, w.p1);
		if ((w.eb + 1) * 2 >= w.c->nq) {
			publish(*w.c, 0, w.p0[0], w.p1[0]);
			publish(*w.c, 1, w.p0[1], w.p1[1]);
			/* relay form: whether the strip behind goes on with this run is for
			 * the second pass to see */
			if (relay && !(w.c->flags & kCtxLast))
				relay_put(*w.c, w.eb + 1, w.m, w.p0, w.p1);
			return false;
		}
		w.eb++;
		w.at += 2 * BS;
		const uint32_t nm = chains(*w.c, w.eb);
		const bool more = (nm & w.m) != 0;
		w.m = nm;
		return more;
	}

	/* relay form: the walker warp winds down; the run goes on at its next item
	 * (which continues it: any of its chain channels will do as "in front") */
	XA_HD void walk_hand_on(const Walk &w) const
	{
		relay_put(*w.c, w.eb, 3u, w.p0, w.p1);
	}
};

/* ---- decode, staged variant (used for stereo) ------------------------------ */
/*
 * The first complete form of the tile algorithm, kept for chain-heavy STEREO
 * data: every block-channel is decoded into a 64-byte row of a shared-memory
 * image of the tile (cut blocks in phase A, chains by one walker lane each, all
 * warps at once), and a store phase interleaves left and right rows into
 * 16-byte units.  Walking per channel keeps twice as many chains in flight as
 * the pair walkers of the direct form; the price is two CTA barriers per tile
 * and the rows' shared-memory traffic, so the census (xa_kernels.cu) picks it
 * only above a measured share of chain blocks (profiles/history_r1.md).
 */
template <int BITS, int CH, int TBQ, int NS, int STAGES>
struct DecSmemStaged {
	typedef DecGeom<BITS, CH, TBQ, NS> G;

	alignas(16) uint8_t in[STAGES][G::IN_BYTES];
	alignas(16) uint32_t out[TBQ * 16];	/* planar rows, 64 B each, swizzled */
	StripCtx ctx[STAGES][NS];
	uint32_t tile_flags[STAGES];		/* kCtxEnd, kCtxTail (any strip) */
	uint32_t n_strips[STAGES];
	uint16_t heads[STAGES][TBQ];		/* first block of every chain of the tile */
	uint32_t n_heads[STAGES];
	uint32_t next_head[STAGES];		/* unused here: heads are dealt out statically */
	alignas(8) unsigned long long full[STAGES];
	alignas(8) unsigned long long ready[STAGES];
	alignas(8) unsigned long long empty[STAGES];
};

template <int BITS, int CH, int TBQ, int NS, int STAGES>
struct DecTileStaged {
	typedef DecGeom<BITS, CH, TBQ, NS> G;
	typedef DecSmemStaged<BITS, CH, TBQ, NS, STAGES> Smem;
	static constexpr int BS = G::BS;
	static constexpr int QB = BITS / 2;
	static constexpr uint32_t SBQ = G::SBQ;
	static constexpr uint32_t SCAN = G::SBQ;	/* scanner items: block-channels */
	static constexpr int kLag = CH;			/* predecessor of item q is q - CH */
	static constexpr bool kStaged = true;
	static constexpr int kStages = STAGES;
	static constexpr int kMinCtas = 3;	/* per SM: caps the registers at 64 */

	const DecodeParams &p;
	Smem &sm;
	const uint8_t *in;		/* this tile's stage buffer */
	const StripCtx *ctx;		/* this tile's strips */
	const uint32_t n_strips;

	XA_HD DecTileStaged(const DecodeParams &p_, Smem &sm_, int stage)
	    : p(p_), sm(sm_), in(sm_.in[stage]), ctx(sm_.ctx[stage]),
	      n_strips(sm_.n_strips[stage])
	{
	}

	/* bytes past a strip's `bulk` fetched one by one (only at the arena's end) */
	XA_HD void load_tail(uint32_t tid, uint32_t nt, uint8_t *in_w)
	{
		for (uint32_t st = 0; st < n_strips; st++) {
			const StripCtx &c = ctx[st];
			uint32_t slot0 = st * G::SLOT;
			for (uint32_t i = c.bulk + tid; i < c.in_need; i += nt)
				in_w[slot0 + i] = p.src[c.a0 + i];
		}
	}

	XA_HD static int row_word(uint32_t q, int chunk, int w)
	{
		return (int)(q * 16 + (uint32_t)((chunk ^ (int)((q >> 1) & 3u)) * 4 + w));
	}

	XA_HD void store_row(uint32_t q, const uint32_t (&o)[16])
	{
#pragma unroll
		for (int j = 0; j < 4; j++) {
			uint4 *d = reinterpret_cast<uint4 *>(&sm.out[row_word(q, j, 0)]);
			uint4 v;
			v.x = o[4 * j]; v.y = o[4 * j + 1]; v.z = o[4 * j + 2]; v.w = o[4 * j + 3];
			*d = v;
		}
	}

	/* byte address (in the stage buffer) of block lq of strip st */
	XA_HD uint32_t block_at(const StripCtx &c, uint32_t lq) const
	{
		return c.in_base + lq * BS;
	}

	XA_HD void fetch_block(uint32_t at, uint32_t (&pw)[BITS]) const
	{
		uint32_t pay = at + 1;		/* first payload byte */
		const uint32_t *w = reinterpret_cast<const uint32_t *>(in) + (pay >> 2);
		load_payload<BITS>(pw, w, (pay & 3u) * 8u);
	}

	/* hand the channel's state to whoever continues it */
	XA_HD void publish(const StripCtx &c, uint32_t ch, int p0, int p1)
	{
		if (c.flags & kCtxLast) {
			p.results[c.stream].prev[ch][0] = (int16_t)p0;
			p.results[c.stream].prev[ch][1] = (int16_t)p1;
		} else {
			unsigned long long v = ((unsigned long long)p.epoch << 32) |
			    ((unsigned long long)(uint16_t)p1 << 16) | (uint16_t)p0;
			mailbox_put(&p.carry[(uint64_t)c.slot * 2 + ch], v);
		}
	}

	XA_HD void carried_in(const StripCtx &c, uint32_t ch, int &p0, int &p1) const
	{
		if (c.flags & kCtxFirst) {
			p0 = p.streams[c.stream].prev[ch][0];
			p1 = p.streams[c.stream].prev[ch][1];
		} else {
			unsigned long long v = mailbox_get(
			    &p.carry[(uint64_t)(c.slot - 1) * 2 + ch], p.epoch, p.fault, p.carry_timeout_ns);
			p0 = (int16_t)(uint16_t)v;
			p1 = (int16_t)(uint16_t)(v >> 16);
		}
	}

	/*
	 * phase A: every cut block is decoded; every chain block whose
	 * predecessor in its channel is not a chain block (or lies in the
	 * stream's previous strip) is queued as the head of a chain.  Needs
	 * n_heads == 0 on entry.
	 */
	/* scanner, one strip: is item q (a block-channel) a chain block?  (bit 0) */
	XA_HD uint32_t chain_mask(uint32_t q) const
	{
		return q < ctx[0].nq && block_kind(in[ctx[0].in_base + q * BS]) == kChain;
	}

	/* is block-channel q of the tile (strip-major) the head of a chain? */
	XA_HD bool is_head(uint32_t q) const
	{
		const StripCtx &c = ctx[q / SBQ];
		const uint32_t lq = q % SBQ;
		if (lq >= c.nq)
			return false;
		const uint32_t at = block_at(c, lq);
		return block_kind(in[at]) == kChain &&
		    (lq < (uint32_t)CH || block_kind(in[at - CH * BS]) != kChain);
	}

	/* phase A: every cut block-channel is decoded into its row */
	XA_HD void phase_a(uint32_t tid, uint32_t nt)
	{
		const uint32_t nq_all = n_strips * SBQ;
		for (uint32_t q = tid; q < nq_all; q += nt) {
			const StripCtx &c = ctx[q / SBQ];
			const uint32_t lq = q % SBQ;
			if (lq >= c.nq)
				continue;
			const uint32_t at = block_at(c, lq);
			const uint32_t prof = in[at];
			const int kind = block_kind(prof);
			if (kind == kChain)
				continue;		/* a walker's */
			if (kind == kBad)
				global_min_u32(&p.first_bad[c.stream], c.first_eb * CH + lq);
			/* a bad block is decoded as if it were a cut so that
			 * nothing downstream waits for it; what lies at and after
			 * it is not part of the result */
			uint32_t pw[BITS], o[16];
			fetch_block(at, pw);
			decode_block_cut<BITS>(o, pw, prof);
			store_row(q, o);
			if (lq + CH >= c.nq)
				publish(c, lq % CH, (int)(int16_t)(o[15] >> 16),
				    (int)(int16_t)(o[15] & 0xffffu));
		}
	}

	/*
	 * walkers: chain i of the tile goes to thread (i + rot) mod nt.  The
	 * state at the head is recomputed from the bytes of the cut block in
	 * front (codes 2 and 3 of its last quad), so walkers depend on nothing
	 * phase A produces and both run in the same phase.
	 */
	XA_HD void phase_walk(uint32_t tid, uint32_t nt, const uint16_t *heads, uint32_t n,
	    uint32_t rot)
	{
		for (uint32_t i = (tid + nt - rot % nt) % nt; i < n; i += nt) {
			uint32_t q = heads[i];
			const StripCtx &c = ctx[q / SBQ];
			uint32_t lq = q % SBQ;
			uint32_t at = block_at(c, lq);
			int p0, p1;
			if (lq < (uint32_t)CH) {
				carried_in(c, lq, p0, p1);
			} else {
				const uint32_t pa = at - CH * BS;
				const int sh = 16 + (int)(in[pa] & 15u);
				const uint32_t a = pa + 1 + 7 * QB;
				const uint32_t *w = reinterpret_cast<const uint32_t *>(in) + (a >> 2);
				int x[4];
				quad_codes<BITS>(funnel_r(w[0], w[1], (a & 3u) * 8u), x);
				p1 = x[2] >> sh;
				p0 = x[3] >> sh;
			}
			for (;;) {
				uint32_t pw[BITS], o[16];
				fetch_block(at, pw);
				decode_block_chain<BITS>(o, pw, in[at], p0, p1);
				store_row(q, o);
				if (lq + CH >= c.nq) {
					publish(c, lq % CH, p0, p1);
					break;
				}
				q += CH;
				lq += CH;
				at += CH * BS;
				if (block_kind(in[at]) != kChain)
					break;
			}
		}
	}

	/* one 16-byte unit (index li within its strip) of interleaved PCM */
	XA_HD void gather_chunk(uint32_t row0, uint32_t li, uint32_t (&w)[4]) const
	{
		if (CH == 1) {
			const uint32_t *s = &sm.out[row_word(row0 + (li >> 2), (int)(li & 3u), 0)];
			w[0] = s[0]; w[1] = s[1]; w[2] = s[2]; w[3] = s[3];
		} else {
			uint32_t eb = li >> 3, jj = li & 7u;
			int j = (int)(jj >> 1), h = (int)(jj & 1u) * 2;
			const uint32_t *l = &sm.out[row_word(row0 + 2 * eb, j, h)];
			const uint32_t *r = &sm.out[row_word(row0 + 2 * eb + 1, j, h)];
			w[0] = byte_perm(l[0], r[0], 0x5410);
			w[1] = byte_perm(l[0], r[0], 0x7632);
			w[2] = byte_perm(l[1], r[1], 0x5410);
			w[3] = byte_perm(l[1], r[1], 0x7632);
		}
	}

	/*
	 * staged rows -> interleaved PCM, 16 bytes per step, for a tile that is
	 * one strip (NS == 1); takes what it needs of the context by value so
	 * that the stage can be handed back before the store.  nt % 32 == 0.
	 */
	XA_HD void phase_store_one(uint32_t tid, uint32_t nt, uint64_t out0, uint32_t nq,
	    uint32_t out_valid)
	{
		const uint32_t nchunk = nq * 4;
		uint8_t *dst = p.dst + out0;
		if (out_valid == nchunk * 16u) {
			/* every unit is whole.  With nt a multiple of 32 the swizzle
			 * term of a thread's units is the same for all of them, so
			 * both addresses advance by nt*16 bytes per step. */
			uint4 *g = reinterpret_cast<uint4 *>(dst) + tid;
			if (CH == 1) {
				const uint4 *s = reinterpret_cast<const uint4 *>(
				    &sm.out[row_word(tid >> 2, (int)(tid & 3u), 0)]);
				for (uint32_t i = tid; i < nchunk; i += nt) {
					*g = *s;
					g += nt;
					s += nt;
				}
			} else {
				const uint32_t jj = tid & 7u;
				const int j = (int)(jj >> 1), h = (int)(jj & 1u) * 2;
				const uint2 *l = reinterpret_cast<const uint2 *>(
				    &sm.out[row_word(2 * (tid >> 3), j, h)]);
				const uint2 *r = reinterpret_cast<const uint2 *>(
				    &sm.out[row_word(2 * (tid >> 3) + 1, j, h)]);
				for (uint32_t i = tid; i < nchunk; i += nt) {
					uint2 a = *l, b = *r;
					uint4 v;
					v.x = byte_perm(a.x, b.x, 0x5410);
					v.y = byte_perm(a.x, b.x, 0x7632);
					v.z = byte_perm(a.y, b.y, 0x5410);
					v.w = byte_perm(a.y, b.y, 0x7632);
					*g = v;
					g += nt;
					l += nt * 2;	/* nt units = nt*16 B of rows */
					r += nt * 2;
				}
			}
			return;
		}
		/* the truncated last block of a stream */
		for (uint32_t i = tid; i < nchunk; i += nt) {
			const uint32_t boff = i * 16u;
			if (boff >= out_valid)
				continue;
			uint32_t w[4];
			gather_chunk(0, i, w);
			uint32_t n16 = (out_valid - boff) / 2u;
			if (n16 > 8u)
				n16 = 8u;
			uint16_t *d = reinterpret_cast<uint16_t *>(dst + boff);
			for (uint32_t k = 0; k < n16; k++)
				d[k] = (uint16_t)(w[k >> 1] >> (16u * (k & 1u)));
		}
	}

	/* the same for any number of strips, reading the contexts in place */
	XA_HD void phase_store(uint32_t tid, uint32_t nt)
	{
		constexpr uint32_t CPS = SBQ * 4;	/* 16-byte units per full strip */
		if (NS == 1) {
			phase_store_one(tid, nt, ctx[0].out0, ctx[0].nq, ctx[0].out_valid);
			return;
		}
		const uint32_t total = n_strips * CPS;
		for (uint32_t i = tid; i < total; i += nt) {
			const uint32_t st = i / CPS, li = i % CPS;
			const StripCtx &c = ctx[st];
			const uint32_t boff = li * 16u;
			if (boff >= c.out_valid)
				continue;
			uint32_t w[4];
			gather_chunk(st * SBQ, li, w);
			uint8_t *d8 = p.dst + c.out0 + boff;
			if (boff + 16u <= c.out_valid) {
				uint4 v;
				v.x = w[0]; v.y = w[1]; v.z = w[2]; v.w = w[3];
				*reinterpret_cast<uint4 *>(d8) = v;
			} else {
				uint16_t *d = reinterpret_cast<uint16_t *>(d8);
				uint32_t n16 = (c.out_valid - boff) / 2u;
				for (uint32_t k = 0; k < n16; k++)
					d[k] = (uint16_t)(w[k >> 1] >> (16u * (k & 1u)));
			}
		}
	}
};

/* ---- encode -------------------------------------------------------------- */
/*
 * Reference-exact encoder (/root/reference/src/libbjxa.c:665-691,759-819):
 * profile byte 0, top-bits truncation, zero-padded last block.  One thread per
 * effective block: it reads 64*CH contiguous PCM bytes and produces CH
 * consecutive XA blocks.
 */

template <int BITS, int CH, int TBE>
struct EncSmem {
	static constexpr int BS = block_bytes(BITS);
	static constexpr int OUT_BYTES = ((TBE * CH * BS + 15 + 15) / 16) * 16 + 16;

	alignas(16) uint8_t in[TBE * CH * 64 + 16];
	alignas(16) uint8_t out[OUT_BYTES];
	alignas(8) unsigned long long mbar;
};

template <int BITS, int CH, int TBE>
struct EncTile {
	typedef EncSmem<BITS, CH, TBE> Smem;
	static constexpr int BS = block_bytes(BITS);

	const EncodeParams &p;
	Smem &sm;
	uint32_t stream, first_eb, neb;
	uint64_t in0;		/* first PCM byte (16-byte aligned) */
	uint32_t in_valid;	/* PCM bytes that exist for this tile */
	uint64_t o0, o1;	/* destination byte range */
	uint64_t oa;		/* o0 rounded down to 16 */
	uint32_t out_off;

	XA_HD EncTile(const EncodeParams &p_, Smem &sm_, uint32_t tile)
	    : p(p_), sm(sm_)
	{
		const TileEnt te = p.tiles[tile];
		const StreamDev &s = p.streams[te.first];
		stream = te.first;
		first_eb = te.j;
		uint32_t rem = s.blocks - first_eb;
		neb = rem < (uint32_t)TBE ? rem : (uint32_t)TBE;
		uint64_t done = (uint64_t)first_eb * (64 * CH);
		in0 = s.pcm_off + done;
		uint64_t have = s.pcm_len > done ? s.pcm_len - done : 0;
		uint64_t full = (uint64_t)neb * (64 * CH);
		in_valid = (uint32_t)(have < full ? have : full);
		o0 = s.xa_off + (uint64_t)first_eb * (BS * CH);
		o1 = o0 + (uint64_t)neb * (BS * CH);
		oa = o0 & ~(uint64_t)15;
		out_off = (uint32_t)(o0 - oa);
	}

	XA_HD uint32_t bulk_bytes() const
	{
		uint64_t end = in0 + ((in_valid + 15u) & ~15u);
		uint64_t lim = p.src_bytes & ~(uint64_t)15;
		if (end > lim)
			end = lim > in0 ? lim : in0;
		return (uint32_t)(end - in0);
	}

	XA_HD void load_tail(uint32_t tid, uint32_t nt)
	{
		for (uint32_t i = bulk_bytes() + tid; i < in_valid; i += nt)
			sm.in[i] = p.src[in0 + i];
	}

	/* n <= 4 bytes of v to an arbitrary byte address of the output image */
	XA_HD void put_bytes(uint32_t at, uint32_t v, int n)
	{
#pragma unroll
		for (int b = 0; b < 4; b++)
			if (b < n)
				sm.out[at + b] = (uint8_t)(v >> (8 * b));
	}

	/*
	 * One thread per 16-byte unit of PCM (8 mono samples, or 4 stereo
	 * frames): neighbouring threads read neighbouring units, so the shared
	 * memory reads are conflict free and nothing diverges.  Every unit maps
	 * to BITS/2 payload bytes per channel-quad at a fixed place of its
	 * block; the profile bytes (0, libbjxa.c:679) are written separately.
	 */
	XA_HD void phase_pack(uint32_t tid, uint32_t nt)
	{
		constexpr uint32_t UPE = 4 * CH;	/* units per effective block */
		constexpr int QB = BITS / 2;		/* payload bytes per 4 samples */
		const uint32_t nunits = neb * UPE;
		const uint32_t whole = in_valid / 16u;	/* units with nothing missing */
		for (uint32_t q = tid; q < neb * CH; q += nt)
			sm.out[out_off + q * BS] = 0;
		for (uint32_t u = tid; u < nunits; u += nt) {
			uint4 w;
			if (u < whole) {
				w = *reinterpret_cast<const uint4 *>(sm.in + u * 16u);
			} else {
				/* the stream ends inside or before this unit: take the
				 * bytes that exist, zero the rest (libbjxa.c:686-690) */
				uint32_t t[4] = { 0, 0, 0, 0 };
				for (uint32_t k = 0; k < 16u && u * 16u + k < in_valid; k++)
					t[k >> 2] |= (uint32_t)sm.in[u * 16u + k] << (8 * (k & 3u));
				w.x = t[0]; w.y = t[1]; w.z = t[2]; w.w = t[3];
			}
			const uint32_t eb = u / UPE, k = u % UPE;
			if (CH == 1) {
				/* 8 samples of one block: two quads */
				uint32_t a = pack4<BITS>(w.x, w.x >> 16, w.y, w.y >> 16);
				uint32_t b = pack4<BITS>(w.z, w.z >> 16, w.w, w.w >> 16);
				uint32_t at = out_off + eb * BS + 1 + k * (2 * QB);
				put_bytes(at, a, QB);
				put_bytes(at + QB, b, QB);
			} else {
				/* 4 frames: one quad of the left, one of the right block */
				uint32_t l = pack4<BITS>(w.x, w.y, w.z, w.w);
				uint32_t r = pack4<BITS>(w.x >> 16, w.y >> 16, w.z >> 16, w.w >> 16);
				uint32_t at = out_off + (eb * 2) * BS + 1 + k * QB;
				put_bytes(at, l, QB);
				put_bytes(at + BS, r, QB);
			}
		}
	}

	XA_HD void phase_store(uint32_t tid, uint32_t nt)
	{
		/* smem byte i mirrors global byte oa + i */
		uint64_t v0 = (o0 + 15) & ~(uint64_t)15;	/* vector part */
		uint64_t v1 = o1 & ~(uint64_t)15;
		if (v0 > v1) {			/* tile smaller than one unit */
			v0 = o1;
			v1 = o1;
		}
		uint8_t *dst = p.dst;
		uint32_t nvec = (uint32_t)((v1 - v0) >> 4);
		for (uint32_t i = tid; i < nvec; i += nt) {
			uint64_t g = v0 + (uint64_t)i * 16;
			const uint32_t *s = reinterpret_cast<const uint32_t *>(sm.out + (g - oa));
#if defined(__CUDA_ARCH__)
			*reinterpret_cast<uint4 *>(dst + g) = *reinterpret_cast<const uint4 *>(s);
#else
			uint32_t *d = reinterpret_cast<uint32_t *>(dst + g);
			d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3];
#endif
		}
		uint32_t head = (uint32_t)(v0 - o0), tail = (uint32_t)(o1 - v1);
		for (uint32_t i = tid; i < head + tail; i += nt) {
			uint64_t g = i < head ? o0 + i : v1 + (i - head);
			dst[g] = sm.out[g - oa];
		}
	}
};

} /* namespace xa */
#endif
