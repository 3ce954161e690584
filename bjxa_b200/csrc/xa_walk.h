/*
 * xa_walk.h -- decode forms that do not walk chains tile by tile: the dense chain
 * walkers (second pass of the SPLIT and RELAY forms) and, further down, the SEGMENT
 * form (a fixed segment per lane, every block through the chain step).
 *
 * The tile forms of xa_tile.h walk a tile's chains (runs of filter-1..4 blocks,
 * /root/reference/src/libbjxa.c:556-575) while the tile sits in shared memory,
 * so a tile's longest chain decides how long its bytes stay there, and only a
 * few hundred chains are in flight per SM.  On chain-rich data that is the
 * bound (profiles/history_r1.md, steps 15-18).  The split form takes the
 * chains out of the tiles:
 *
 *   pass 1  xa_decode_kernel (direct form, `split` set): units of cut blocks as
 *           before; the scanner does not queue the heads of chains for the
 *           tile's walker warp but writes them out -- one 128-byte LiveRec per
 *           tile that has any: a bitmap of its heads and what a walker needs to
 *           know of the strip;
 *   pass 2  xa_walk_kernel: every LANE of every warp walks one chain (stereo:
 *           one run of effective blocks, both channels side by side) at a
 *           time, straight from global memory through a private shared-memory
 *           window, and draws the next head as soon as its chain ends.  A
 *           chain is followed across tile boundaries to its end, so there are
 *           no carries and no ordering between tiles; the state at a head is
 *           recomputed from the bytes of the cut block in front of it.
 *
 * The RELAY form is the hybrid of the two: the tiles walk their own chains as
 * before -- cheaply, out of the staged bytes -- but a walker warp does not wait
 * for its longest chains: once its heads are dealt out and few lanes are still
 * busy it hands their chains on (RelayRec: next item, state), as it does with
 * any chain that reaches the end of the strip, and pass 2 finishes them.  No
 * carry mailboxes, no order among tiles, and a tile's bytes leave shared memory
 * as soon as the bulk of its chains is through.
 *
 * This header holds what both passes and the CPU single-stepper of the tests
 * (tests/emul) share: the record and the per-item arithmetic.  The GPU-only
 * part (window ring, staged copy-out, the warp's draw) is in xa_kernels.cu.
 */
#ifndef XA_WALK_H
#define XA_WALK_H

#include "xa_tile.h"

namespace xa {

/* one tile with at least one head; written by pass 1's scanner warp, 32 lanes x 4 bytes */
struct LiveRec {
	uint32_t heads[16];	/* bit q: item q of the tile starts a chain (mono: block q;
				 * stereo: effective block q, the first of a run) */
	uint32_t xa_lo, xa_hi;	/* arena address of the tile's first block */
	uint32_t out_lo, out_hi;	/* arena address of its PCM */
	uint32_t stream;
	uint32_t first_eb;	/* the tile's first effective block */
	uint32_t blocks;	/* effective blocks of the stream */
	uint32_t pad[9];
};
static_assert(sizeof(LiveRec) == 128, "one record = one 128-byte line");
enum { kRecXaLo = 16, kRecXaHi, kRecOutLo, kRecOutHi, kRecStream, kRecFirstEb, kRecBlocks };

/*
 * The arithmetic of one walker.  An ITEM is one effective block: CH blocks of
 * BS bytes, contiguous in the XA arena, 64 * CH bytes of interleaved PCM.
 */
template <int BITS, int CH>
struct Walk {
	static constexpr int BS = block_bytes(BITS);
	static constexpr int STEP = CH * BS;		/* XA bytes per item */
	static constexpr int OUT = 64 * CH;		/* PCM bytes per item */
	static constexpr int QB = BITS / 2;		/* payload bytes of 4 samples */
	static constexpr int UNITS = 4 * CH;		/* 16-byte units of output per item */
	/* what a walker looks at beyond its item: the profile byte(s) of the next */
	static constexpr int PEEK = (CH - 1) * BS + 1;

	struct Item {
		uint32_t prof[CH];
		uint32_t pw[CH][BITS];
	};

	/* chain channels of an item: bit 0 left / mono, bit 1 right */
	XA_HD static uint32_t mask_of(const uint32_t (&prof)[CH])
	{
		uint32_t m = 0;
#pragma unroll
		for (int c = 0; c < CH; c++)
			m |= (uint32_t)(block_kind(prof[c]) == kChain) << c;
		return m;
	}

	/*
	 * The item in front of a head, as far as the walker needs it: per channel the
	 * profile byte and its last quad of codes -- fetched as the two aligned words
	 * around the block's last four bytes and put together in seed_apply, so that
	 * on the GPU nothing waits for these loads before the turn's decode is over.
	 * `first`: there is no such item, the state comes with the stream
	 * (libbjxa.c:417-420).
	 */
	struct Seed {
		uint32_t prof[CH];
		uint32_t lo[CH], hi[CH];
	};

	/* arena address of the last four bytes of channel c's block of the item in front of a */
	XA_HD static uint64_t tail_addr(uint64_t a, int c)
	{
		return a - STEP + (uint64_t)(c * BS + BS - 4);
	}

	XA_HD static uint32_t word_at(const uint8_t *src, uint64_t aligned)
	{
#if defined(__CUDA_ARCH__)
		return *reinterpret_cast<const uint32_t *>(src + aligned);
#else
		const uint8_t *b = src + aligned;
		return (uint32_t)b[0] | (uint32_t)b[1] << 8 | (uint32_t)b[2] << 16 | (uint32_t)b[3] << 24;
#endif
	}

	XA_HD static void seed_fetch(const DecodeParams &p, uint32_t stream, bool first,
	    uint64_t a, Seed &s)
	{
		if (first) {
			const StreamDev &sd = p.streams[stream];
#pragma unroll
			for (int c = 0; c < CH; c++) {
				s.prof[c] = 0;
				s.lo[c] = (uint32_t)(uint16_t)sd.prev[c][0] |
				    (uint32_t)(uint16_t)sd.prev[c][1] << 16;
				s.hi[c] = 0;
			}
			return;
		}
#pragma unroll
		for (int c = 0; c < CH; c++) {
			const uint64_t t = tail_addr(a, c);
			s.prof[c] = p.src[a - STEP + (uint64_t)(c * BS)];
			s.lo[c] = word_at(p.src, t & ~(uint64_t)3);
			/* the word behind holds bytes of the head item itself, or is not
			 * needed (t aligned): never past the arena */
			s.hi[c] = (t & 3u) ? word_at(p.src, (t & ~(uint64_t)3) + 4u) : 0u;
		}
	}

	/* the state at the head, into p0/p1 */
	XA_HD static void seed_apply(const Seed &s, bool first, uint64_t a,
	    int (&p0)[CH], int (&p1)[CH])
	{
		if (first) {
#pragma unroll
			for (int c = 0; c < CH; c++) {
				p0[c] = (int16_t)(uint16_t)s.lo[c];
				p1[c] = (int16_t)(uint16_t)(s.lo[c] >> 16);
			}
			return;
		}
#pragma unroll
		for (int c = 0; c < CH; c++) {
			/* a cut (or invalid, decoded as cut) block: its last two samples
			 * are codes 2 and 3 of its last quad, shifted by its range */
			const uint32_t tail = funnel_r(s.lo[c], s.hi[c], (uint32_t)(tail_addr(a, c) & 3u) * 8u);
			const int sh = 16 + (int)(s.prof[c] & 15u);
			int x[4];
			quad_codes<BITS>(tail >> (8 * (4 - QB)), x);
			p1[c] = x[2] >> sh;
			p0[c] = x[3] >> sh;
		}
	}

	/*
	 * One item.  Mono: the walkers' biased predictor step.  Stereo: both channels
	 * side by side, two independent dependency chains; a cut or invalid block
	 * inside a run simply has k0 = k1 = 0.  Units go out through out(j, v),
	 * j = 0 .. UNITS-1 in PCM order.
	 */
	template <bool RANGED = false, class Out>
	XA_HD static void decode(const Item &it, int (&p0)[CH], int (&p1)[CH], Out &out)
	{
		if (CH == 1) {
			uint32_t o[16];
			decode_block_chain<BITS, RANGED>(o, it.pw[0], it.prof[0], p0[0], p1[0]);
#pragma unroll
			for (int j = 0; j < 4; j++) {
				uint4 v;
				v.x = o[4 * j]; v.y = o[4 * j + 1];
				v.z = o[4 * j + 2]; v.w = o[4 * j + 3];
				out(j, v);
			}
		} else {
			/* the biased step for both channels (xa_core.h:sample_chain_b): in a
			 * kernel whose registers are not capped the two extra constants per
			 * channel cost nothing, and the step keeps the ALU pipe -- which
			 * bounds this kernel (72 % busy with the plain step) -- two
			 * instructions a sample shorter */
			const uint32_t pl = it.prof[0], pr = it.prof[CH - 1];
			const int shl = 16 + (int)(pl & 15u), shr = 16 + (int)(pr & 15u);
			const int k0l = gain_k0(pl >> 4), k1l = gain_k1(pl >> 4);
			const int k0r = gain_k0(pr >> 4), k1r = gain_k1(pr >> 4);
			const int cl = chain_bias_c(k0l, k1l), cr = chain_bias_c(k0r, k1r);
			int bl0 = p0[0] + 32768, bl1 = p1[0] + 32768;
			int br0 = p0[CH - 1] + 32768, br1 = p1[CH - 1] + 32768;
#pragma unroll
			for (int j = 0; j < 8; j++) {
				int l[4], r[4];
#pragma unroll
				for (int k = 0; k < 4; k++) {
					if (RANGED) {
						l[k] = sample_chain_r(top_code<BITS>(it.pw[0], 4 * j + k), shl,
						    k0l, k1l, cl, bl0, bl1);
						r[k] = sample_chain_r(top_code<BITS>(it.pw[CH - 1], 4 * j + k), shr,
						    k0r, k1r, cr, br0, br1);
					} else {
						l[k] = sample_chain_b(top_code<BITS>(it.pw[0], 4 * j + k), shl,
						    k0l, k1l, cl, bl0, bl1);
						r[k] = sample_chain_b(top_code<BITS>(it.pw[CH - 1], 4 * j + k), shr,
						    k0r, k1r, cr, br0, br1);
					}
				}
				uint4 v;
				v.x = pack2_biased(l[0], r[0]);
				v.y = pack2_biased(l[1], r[1]);
				v.z = pack2_biased(l[2], r[2]);
				v.w = pack2_biased(l[3], r[3]);
				out(j, v);
			}
			p0[0] = bl0 - 32768;
			p1[0] = bl1 - 32768;
			p0[CH - 1] = br0 - 32768;
			p1[CH - 1] = br1 - 32768;
		}
	}

	/* PCM bytes the stream's LAST item owes (libbjxa.c:622-624,648) */
	XA_HD static uint32_t last_valid(const StreamDev &sd)
	{
		const uint64_t before = (uint64_t)(sd.blocks - 1u) * OUT;
		const uint64_t owed = sd.pcm_len > before ? sd.pcm_len - before : 0;
		return (uint32_t)(owed < (uint64_t)OUT ? owed : (uint64_t)OUT);
	}

	/* invalid filters met inside a stereo run are recorded like the units' */
	XA_HD static void note_bad(const DecodeParams &p, uint32_t stream, uint32_t left,
	    const uint32_t (&prof)[CH])
	{
#pragma unroll
		for (int c = 0; c < CH; c++)
			if (prof[c] >> 4 >= 5u) {
				const uint32_t eb = p.streams[stream].blocks - 1u - left;
				global_min_u32(&p.first_bad[stream], eb * CH + c);
			}
	}

	XA_HD static void put_result(const DecodeParams &p, uint32_t stream,
	    const int (&p0)[CH], const int (&p1)[CH])
	{
#pragma unroll
		for (int c = 0; c < CH; c++) {
			p.results[stream].prev[c][0] = (int16_t)p0[c];
			p.results[stream].prev[c][1] = (int16_t)p1[c];
		}
	}
};

/*
 * One chain (stereo: one run) from its first item to its end, reading the arena
 * directly: the reference semantics of a walker lane, used by the CPU
 * single-stepper of the tests.  `m_front` != 0: the chain channels of the item in
 * front -- the first item then only counts if it goes on with one of them (the
 * relay form does not know what lies behind a strip); 0: a listed head.
 */
template <int BITS, int CH>
XA_HD void walk_chain_serial(const DecodeParams &p, uint32_t stream, uint64_t a, uint64_t o,
    uint32_t left, int (&p0)[CH], int (&p1)[CH], uint32_t m_front)
{
	typedef Walk<BITS, CH> W;
	uint32_t m = 0;
	for (bool fresh = true;; fresh = false) {
		typename W::Item it;
#pragma unroll
		for (int c = 0; c < CH; c++) {
			const uint8_t *b = p.src + a + c * W::BS;
			it.prof[c] = b[0];
#pragma unroll
			for (int i = 0; i < BITS; i++)
				it.pw[c][i] = (uint32_t)b[1 + 4 * i] | (uint32_t)b[2 + 4 * i] << 8 |
				    (uint32_t)b[3 + 4 * i] << 16 | (uint32_t)b[4 + 4 * i] << 24;
		}
		if (fresh) {
			m = W::mask_of(it.prof);
			if (m_front != 0 && (m & m_front) == 0)
				return;		/* the chain ended with the strip */
		}
		const uint32_t valid = left != 0 ? (uint32_t)W::OUT : W::last_valid(p.streams[stream]);
		uint8_t *dst = p.dst + o;
		auto out = [&](int j, const uint4 &v) {
			const uint32_t boff = (uint32_t)j * 16u;
			const uint32_t w[4] = { v.x, v.y, v.z, v.w };
			for (uint32_t k = 0; k < 8u && boff + 2u * k + 2u <= valid; k++) {
				const uint16_t h = (uint16_t)(w[k >> 1] >> (16u * (k & 1u)));
				dst[boff + 2u * k] = (uint8_t)h;
				dst[boff + 2u * k + 1u] = (uint8_t)(h >> 8);
			}
		};
		if (CH == 2)
			W::note_bad(p, stream, left, it.prof);
		W::decode(it, p0, p1, out);
		if (left == 0) {
			W::put_result(p, stream, p0, p1);
			return;
		}
		uint32_t nprof[CH];
#pragma unroll
		for (int c = 0; c < CH; c++)
			nprof[c] = p.src[a + W::STEP + c * W::BS];
		const uint32_t nm = W::mask_of(nprof);
		if ((nm & m) == 0)
			return;
		m = nm;
		a += W::STEP;
		o += W::OUT;
		left--;
	}
}

/* the whole of pass 2 for one record of the split form */
template <int BITS, int CH>
XA_HD void walk_record_serial(const DecodeParams &p, const LiveRec &r)
{
	typedef Walk<BITS, CH> W;
	const uint64_t xa0 = (uint64_t)r.xa_hi << 32 | r.xa_lo;
	const uint64_t out0 = (uint64_t)r.out_hi << 32 | r.out_lo;
	for (uint32_t q = 0; q < 512; q++) {
		if (!(r.heads[q >> 5] >> (q & 31u) & 1u))
			continue;
		const uint64_t a = xa0 + (uint64_t)q * W::STEP;
		const bool first = r.first_eb + q == 0;
		typename W::Seed seed;
		int p0[CH], p1[CH];
		W::seed_fetch(p, r.stream, first, a, seed);
		W::seed_apply(seed, first, a, p0, p1);
		walk_chain_serial<BITS, CH>(p, r.stream, a, out0 + (uint64_t)q * W::OUT,
		    r.blocks - 1u - (r.first_eb + q), p0, p1, 0u);
	}
}

/* ... and for one record of the relay form */
template <int BITS, int CH>
XA_HD void walk_relay_serial(const DecodeParams &p, const RelayRec &r)
{
	int p0[CH], p1[CH];
#pragma unroll
	for (int c = 0; c < CH; c++) {
		p0[c] = (int16_t)(uint16_t)r.st[c];
		p1[c] = (int16_t)(uint16_t)(r.st[c] >> 16);
	}
	walk_chain_serial<BITS, CH>(p, r.stream & 0x3fffffffu, (uint64_t)r.xa_hi << 32 | r.xa_lo,
	    (uint64_t)r.out_hi << 32 | r.out_lo, r.left, p0, p1, r.stream >> 30);
}

/*
 * The SEGMENT form: no chains, no heads, one pass.  A stream is cut into segments
 * of kSegItems items; a lane decodes one segment from its first item to its last
 * with the chain step throughout -- a cut or invalid block is the step with
 * k0 = k1 = 0 (libbjxa.c:526), which forgets the state by itself.  So the only
 * thing a lane has to find is the state in front of its segment (seg_front):
 *   - the stream's own, in front of item 0 (libbjxa.c:417-420);
 *   - else it looks back over at most kSegBack items for the nearest cut (or
 *     invalid) block of every channel and starts decoding there, storing nothing
 *     until it reaches its segment: whatever state it starts with is forgotten
 *     at that block.  On data that has cut blocks at all this costs a few items
 *     per segment, and nothing is ever waited for;
 *   - if a channel has no such block within reach, the state comes from whoever
 *     decodes the segment in front (see "Lanes and streams" below).  On data
 *     without cut blocks this is the serial walk such data demands.
 * All 32 lanes of a warp decode a block in every turn but the few at either end.
 */
template <int BITS, int CH>
struct SegFront {
	uint32_t back;		/* items in front of the segment to decode first */
	bool mail;		/* state from the mailbox (back == 0) */
	bool own;		/* state from the stream descriptor */
};

/* `prof(k, c)`: profile byte of channel c of the item k + 1 in front of the segment */
template <int BITS, int CH, class Prof>
XA_HD SegFront<BITS, CH> seg_front(uint32_t n0, Prof prof)
{
	SegFront<BITS, CH> f = { 0u, false, false };
	if (n0 == 0) {
		f.own = true;
		return f;
	}
	const uint32_t lim = n0 < kSegBack ? n0 : kSegBack;
	uint32_t found[CH], missing = CH;
#pragma unroll
	for (int c = 0; c < CH; c++)
		found[c] = 0;
	for (uint32_t k = 0; k < lim && missing != 0; k++)
#pragma unroll
		for (int c = 0; c < CH; c++)
			if (found[c] == 0 && block_kind(prof(k, c)) != kChain) {
				found[c] = k + 1u;
				missing--;
			}
	if (missing == 0) {
#pragma unroll
		for (int c = 0; c < CH; c++)
			f.back = found[c] > f.back ? found[c] : f.back;
	} else if (lim == n0) {
		f.back = n0;
		f.own = true;
	} else {
		f.mail = true;
	}
	return f;
}

/*
 * Lanes and streams (xa_plan.h: emit_seg_tiles).  The class's segments, streams in
 * arena order, are dealt out 32 to a tile: lane L of a tile decodes the L-th segment
 * counted from segment te.count of stream order[te.first].  A lane that cannot
 * recompute its state waits for the segment in front: the lane before it -- the
 * warp then makes another pass over the tile for the lanes that had to wait -- or,
 * lane 0, the mailbox that the last lane of the tile before leaves (a lower ticket).
 */
template <int BITS, int CH>
struct SegLane {
	uint32_t stream, seg, n0, n, slot;
	uint64_t a0, o0;
	bool valid, ends;
};

template <int BITS, int CH>
XA_HD SegLane<BITS, CH> seg_lane(const DecodeParams &p, const TileEnt &te, uint32_t lane)
{
	typedef Walk<BITS, CH> W;
	SegLane<BITS, CH> l;
	l.valid = lane < te.j;
	l.stream = 0; l.seg = 0; l.n0 = 0; l.n = 0; l.slot = 0; l.a0 = 0; l.o0 = 0; l.ends = false;
	if (!l.valid)
		return l;
	uint32_t k = te.first, at = te.count + lane;
	for (;;) {
		const uint32_t ns = (p.streams[p.order[k]].blocks + kSegItems - 1u) / kSegItems;
		if (at < ns)
			break;
		at -= ns;
		k++;
	}
	l.stream = p.order[k];
	l.seg = at;
	const StreamDev &sd = p.streams[l.stream];
	l.n0 = at * kSegItems;
	l.n = sd.blocks - l.n0 < kSegItems ? sd.blocks - l.n0 : kSegItems;
	l.ends = l.n0 + l.n == sd.blocks;
	l.a0 = sd.xa_off + (uint64_t)l.n0 * W::STEP;
	l.o0 = sd.pcm_off + (uint64_t)l.n0 * W::OUT;
	l.slot = sd.slot_base + at;
	return l;
}

/* one lane's segment from `back` items in front of it to its end, p0/p1 the state there */
template <int BITS, int CH>
XA_HD void walk_seg_lane_serial(const DecodeParams &p, const SegLane<BITS, CH> &l, uint32_t back,
    int (&p0)[CH], int (&p1)[CH])
{
	typedef Walk<BITS, CH> W;
	const StreamDev &sd = p.streams[l.stream];
	for (int cur = -(int)back; cur < (int)l.n; cur++) {
		const uint64_t a = l.a0 + (int64_t)cur * W::STEP;
		typename W::Item it;
#pragma unroll
		for (int c = 0; c < CH; c++) {
			const uint8_t *b = p.src + a + c * W::BS;
			it.prof[c] = b[0];
#pragma unroll
			for (int i = 0; i < BITS; i++)
				it.pw[c][i] = (uint32_t)b[1 + 4 * i] | (uint32_t)b[2 + 4 * i] << 8 |
				    (uint32_t)b[3 + 4 * i] << 16 | (uint32_t)b[4 + 4 * i] << 24;
		}
		const bool last = l.n0 + (uint32_t)cur + 1u == sd.blocks;
		const uint32_t valid = cur < 0 ? 0u : last ? W::last_valid(sd) : (uint32_t)W::OUT;
		uint8_t *dst = p.dst + l.o0 + (int64_t)cur * W::OUT;
		auto out = [&](int j, const uint4 &v) {
			const uint32_t boff = (uint32_t)j * 16u;
			const uint32_t w[4] = { v.x, v.y, v.z, v.w };
			for (uint32_t k = 0; k < 8u && boff + 2u * k + 2u <= valid; k++) {
				const uint16_t h = (uint16_t)(w[k >> 1] >> (16u * (k & 1u)));
				dst[boff + 2u * k] = (uint8_t)h;
				dst[boff + 2u * k + 1u] = (uint8_t)(h >> 8);
			}
		};
		if (cur >= 0) {
#pragma unroll
			for (int c = 0; c < CH; c++)
				if (it.prof[c] >> 4 >= 5u)
					global_min_u32(&p.first_bad[l.stream], (l.n0 + (uint32_t)cur) * CH + c);
		}
		W::decode(it, p0, p1, out);
	}
	if (l.ends)
		W::put_result(p, l.stream, p0, p1);
}

/*
 * One tile, pass by pass: the reference semantics of xa_seg_kernel, used by the CPU
 * single-stepper of the tests.  `visit(i)`: the order in which the lanes of a pass
 * are taken (any order must do).
 */
template <int BITS, int CH, class Visit>
XA_HD void walk_seg_tile_serial(const DecodeParams &p, const TileEnt &te, Visit visit)
{
	typedef Walk<BITS, CH> W;
	SegLane<BITS, CH> ln[32];
	int st0[32][CH], st1[32][CH];
	uint32_t back[32];
	enum { kIdle, kReady, kPending, kDone } state[32];
	for (uint32_t lane = 0; lane < 32; lane++) {
		SegLane<BITS, CH> &l = ln[lane];
		l = seg_lane<BITS, CH>(p, te, lane);
		state[lane] = kIdle;
		back[lane] = 0;
		if (!l.valid)
			continue;
		const SegFront<BITS, CH> f = seg_front<BITS, CH>(l.n0, [&](uint32_t k, int c) {
			return (uint32_t)p.src[l.a0 - (uint64_t)(k + 1u) * W::STEP + (uint64_t)(c * W::BS)];
		});
		const StreamDev &sd = p.streams[l.stream];
		back[lane] = f.back;
		state[lane] = kReady;
#pragma unroll
		for (int c = 0; c < CH; c++) {
			st0[lane][c] = st1[lane][c] = 0;
			if (f.own) {
				st0[lane][c] = sd.prev[c][0];
				st1[lane][c] = sd.prev[c][1];
			} else if (f.mail && lane == 0) {
				const unsigned long long v = mailbox_get(
				    &p.carry[(uint64_t)(l.slot - 1u) * 2 + c], p.epoch, p.fault,
				    p.carry_timeout_ns);
				st0[lane][c] = (int16_t)(uint16_t)v;
				st1[lane][c] = (int16_t)(uint16_t)(v >> 16);
			} else if (f.mail) {
				state[lane] = kPending;		/* lane - 1 of this tile decodes the segment in front */
			}
		}
	}
	for (;;) {
		bool any = false;
		for (uint32_t i = 0; i < 32; i++) {
			const uint32_t lane = visit(i);
			if (state[lane] != kReady)
				continue;
			walk_seg_lane_serial<BITS, CH>(p, ln[lane], back[lane], st0[lane], st1[lane]);
			any = true;
		}
		if (!any)
			break;
		/* what the pass leaves: the state for a lane that waited, or for the next tile */
		for (uint32_t lane = 0; lane < 32; lane++)
			if (state[lane] == kReady) {
				state[lane] = kDone;
				const SegLane<BITS, CH> &l = ln[lane];
				if (!l.ends && lane == 31u) {
#pragma unroll
					for (int c = 0; c < CH; c++)
						mailbox_put(&p.carry[(uint64_t)l.slot * 2 + c],
						    ((unsigned long long)p.epoch << 32) |
						    ((unsigned long long)(uint16_t)st1[lane][c] << 16) |
						    (uint16_t)st0[lane][c]);
				}
			}
		for (uint32_t lane = 31; lane >= 1; lane--)
			if (state[lane] == kPending && state[lane - 1] == kDone) {
				/* (a lane that became ready in this very step does not hand on yet) */
#pragma unroll
				for (int c = 0; c < CH; c++) {
					st0[lane][c] = st0[lane - 1][c];
					st1[lane][c] = st1[lane - 1][c];
				}
				back[lane] = 0;
				state[lane] = kReady;
			}
	}
}

} /* namespace xa */
#endif
