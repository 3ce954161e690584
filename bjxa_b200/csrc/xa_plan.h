/*
 * xa_plan.h -- host-side batch planning: turns a table of stream descriptors
 * into per-(bits,channels) tile lists and the device records of xa_tile.h.
 * Pure host C++ (no CUDA calls) so the same code feeds the sm_100a launcher
 * (xa_kernels.cu) and the CPU single-stepper used by the tests (tests/emul).
 *
 * The reference has no counterpart: it walks one stream serially
 * (/root/reference/src/libbjxa.c:629-658).  Here every stream is cut into
 * tiles of TBE effective blocks; tiles are issued "time-major" (tile j of every
 * stream before tile j+1 of any) so that concurrently resident CTAs work on
 * different streams and a tile's predecessor in its stream has long finished
 * when its carry is needed.
 */
#ifndef XA_PLAN_H
#define XA_PLAN_H

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "xa_tile.h"

/* public descriptor, see include/bjxa_batch.h */
struct bjxa_stream_desc;

namespace xa {

/* tile geometry: block-channels per decode tile, effective blocks per encode tile */
#ifndef XA_DEC_TBQ
#define XA_DEC_TBQ 512
#endif
#ifndef XA_DEC_NT
#define XA_DEC_NT 256
#endif
#ifndef XA_DEC_WIDE
#define XA_DEC_WIDE 32
#endif
constexpr int kDecTBQ = XA_DEC_TBQ;		/* block-channels per tile */
constexpr int kDecThreads = XA_DEC_NT;		/* consumer threads (+1 producer warp) */
/*
 * Source buffers (tiles) in flight per CTA of the direct forms.  A tile with
 * chains stays in flight for as long as its longest chain takes, so the deeper
 * the ring the more chains run at once; the depth is what fits next to the
 * occupancy the registers allow (3 or 4 CTAs of 320 threads per SM) in 227 KB
 * without costing the cut-only case anything (measured: profiles/history_r1.md).
 * -DXA_DEC_STAGES=n overrides it for geometry sweeps.
 */
constexpr int dec_stages(int bits, int ch)
{
#ifdef XA_DEC_STAGES
	return XA_DEC_STAGES;
#else
	return ch == 1 ? (bits == 4 ? 4 : 3) : (bits == 4 ? 3 : bits == 6 ? 5 : 4);
#endif
}
/*
 * Ring depth of the pooled form (xa_decode_pool_kernel): as many tiles as fit
 * next to each other when `ctas` CTAs share the 227 KB of an SM -- their chains
 * are one pool, and the deeper the ring the longer a long chain may take before
 * it holds the loader up.  -DXA_POOL_STAGES=n overrides it.
 */
#ifndef XA_POOL_CTAS
#define XA_POOL_CTAS 3
#endif
constexpr int pool_stages(int bits, int ch, int ctas = XA_POOL_CTAS)
{
#ifdef XA_POOL_STAGES
	return XA_POOL_STAGES;
#else
	/* stage buffer (DecGeom::IN_BYTES for one strip) + heads + context + words */
	const int stage = ((XA_DEC_TBQ * (4 * bits + 1) + 30) / 16) * 16 + 16 + XA_DEC_TBQ * 2 + 48 + 56;
	const int n = ((227 * 1024) / ctas - 1024 - 64) / stage;
	return n > 12 ? 12 : n;
#endif
}
constexpr int kDecWide = XA_DEC_WIDE;		/* strips per tile in "wide" mode */
constexpr int kEncTBE = 256;
constexpr int kEncThreads = 128;
static_assert(kDecThreads % 32 == 0 && kDecTBQ % 2 == 0, "tile geometry");
static_assert(kDecWide <= 32, "one producer lane per strip");

enum { kKindDecode = 0, kKindEncode = 1, kKindSearch = 2 };	/* = BJXA_PLAN_* */

XA_HD int bucket_of(int bits, int ch) { return (bits / 2 - 2) * 2 + (ch - 1); }
XA_HD int bucket_bits(int b) { return 4 + 2 * (b / 2); }
XA_HD int bucket_ch(int b) { return 1 + (b & 1); }

struct HostPlan {
	int kind;
	std::vector<StreamDev> streams;
	std::vector<uint8_t> bucket;		/* per stream */
	std::vector<uint32_t> order;		/* streams in issue order, by bucket */
	std::vector<TileEnt> tiles;		/* all buckets, concatenated */
	uint32_t tile_begin[7];			/* bucket b owns [b], [b+1]) */
	uint32_t alt_begin[7];			/* decode: its list in the other tile shape */
	int alt_ns[6];				/* strips per tile there, 0 = no such list */
	uint32_t order_begin[7];		/* likewise, into order */
	uint32_t seg_begin[7];			/* decode: its list for the segment form (xa_walk.h),
						 * empty for classes too small for it */
	int ns[6];				/* decode: strips per tile of bucket b */
	uint32_t n_slots;
	uint64_t src_need, dst_need;		/* arena bytes the batch touches */
};

/*
 * Strips per decode tile.  One long strip (kDecTBQ block-channels of one
 * stream) is the shape for streams with cut blocks; kDecWide short strips of
 * different streams give every tile kDecWide independent chains even when the
 * streams have no cut block at all, which is the only parallelism such data
 * has.  Measured on B200 (profiles/history_r1.md): the wide shape wins only
 * when nearly every block is a chain block (about 6x on all-chain data) and
 * loses 1.3-1.8x otherwise, so a class of at least kDecWideMinStreams streams
 * gets BOTH tile lists and the census kernel picks per launch; smaller classes
 * already have every stream in flight with one strip per tile.  `force`
 * (1 or kDecWide; 0 = automatic) builds that shape only: a tuning/testing
 * override.
 */
constexpr size_t kDecWideMinStreams = 1024;
/* the segment form (xa_walk.h): classes with enough items for its tiles (32 lanes x
 * kSegItems) to fill the device get its list as well.  (BJXA_B200_SEG=on lifts
 * that: tests force the form on small batches.) */
constexpr uint64_t kSegMinItems = 1ull << 22;

/* effective blocks per strip */
inline uint32_t strip_blocks(int ns, int ch)
{
	return (uint32_t)(kDecTBQ / ns / ch);
}

/*
 * The decode tiles of one class in one shape, time-major: strip j of every
 * stream before strip j+1 of any, so a strip's predecessor in its stream always
 * holds a lower ticket.  `o` = the class's streams, longest first (the streams
 * still active at step j are a prefix); they sit at order[order0...].
 */
inline void emit_decode_tiles(HostPlan &hp, const std::vector<uint32_t> &o, uint32_t order0,
    int ns, int ch)
{
	const uint32_t sbe = strip_blocks(ns, ch);
	size_t active = o.size();
	for (uint32_t j = 0; active > 0; j++) {
		while (active > 0 && (uint64_t)j * sbe >= hp.streams[o[active - 1]].blocks)
			active--;
		for (size_t base = 0; base < active; base += (size_t)ns) {
			size_t cnt = std::min(active - base, (size_t)ns);
			TileEnt te = { order0 + (uint32_t)base, (uint32_t)cnt, j, 0u };
			hp.tiles.push_back(te);
		}
	}
}

/*
 * The segment form's list (xa_walk.h): all segments (kSegItems items) of all streams
 * of the class, the streams in the order they lie in the XA arena, dealt out 32 to a
 * tile -- a lane each.  A long stream fills many tiles by itself, short ones share
 * a tile; either way the lanes of a warp, and the warps that run at the same time,
 * read and write next to each other (a warp's turn touches a handful of pages of the
 * arenas; with one stream per lane it was 32 + 32, and on long streams the address
 * translation, not the arithmetic, then set the pace: profiles/history_r2.md).
 *   te.first  index into order[] of the stream of lane 0 (the class's streams in
 *             arena order are appended to order[] for this list)
 *   te.count  segment of that stream lane 0 decodes
 *   te.j      lanes in use
 *   te.pad    streams the tile touches
 * The segment in front of a lane's is the lane before it, or -- lane 0 -- the last
 * lane of the tile before: a lower ticket.
 */
inline uint32_t seg_count(uint32_t blocks)
{
	return (blocks + kSegItems - 1) / kSegItems;
}

inline void emit_seg_tiles(HostPlan &hp, const std::vector<uint32_t> &members)
{
	std::vector<uint32_t> o(members);
	std::stable_sort(o.begin(), o.end(), [&](uint32_t x, uint32_t y) {
		return hp.streams[x].xa_off < hp.streams[y].xa_off;
	});
	const uint32_t order0 = (uint32_t)hp.order.size();
	hp.order.insert(hp.order.end(), o.begin(), o.end());
	TileEnt te = { 0u, 0u, 0u, 0u };
	for (size_t k = 0; k < o.size(); k++) {
		const uint32_t ns = seg_count(hp.streams[o[k]].blocks);
		for (uint32_t sg = 0; sg < ns; ) {
			if (te.j == 0) {
				te.first = order0 + (uint32_t)k;
				te.count = sg;
				te.pad = 0;
			}
			const uint32_t take = std::min(ns - sg, 32u - te.j);
			te.j += take;
			te.pad++;
			sg += take;
			if (te.j == 32) {
				hp.tiles.push_back(te);
				te.j = 0;
			}
		}
	}
	if (te.j != 0)
		hp.tiles.push_back(te);
}

/*
 * Validates the descriptors and builds the plan.  Returns 0, or an errno
 * value (EINVAL) with *bad_index set to the offending stream.
 */
template <class Desc>
inline int build_plan(HostPlan &hp, int kind, const Desc *d, size_t n,
    size_t *bad_index, int force_strips = 0, uint64_t seg_min_items = kSegMinItems)
{
	hp.kind = kind;
	hp.streams.resize(n);
	hp.bucket.resize(n);
	hp.tiles.clear();
	hp.order.clear();
	hp.n_slots = 0;
	hp.src_need = hp.dst_need = 0;

	std::vector<uint32_t> members[6];
	bool has_seg[6] = { false, false, false, false, false, false };
	for (size_t i = 0; i < n; i++) {
		const Desc &s = d[i];
		if (s.blocks == 0) {
			/* takes no part: no tiles, nothing validated */
			std::memset(&hp.streams[i], 0, sizeof hp.streams[i]);
			hp.bucket[i] = 0;
			continue;
		}
		bool ok = (s.bits == 4 || s.bits == 6 || s.bits == 8) &&
		    (s.channels == 1 || s.channels == 2) && (s.pcm_off % 16 == 0) &&
		    (s.pcm_len % (2u * s.channels) == 0) &&
		    (uint64_t)s.pcm_len <= (uint64_t)s.blocks * 64u * s.channels;
		if (!ok) {
			if (bad_index)
				*bad_index = i;
			return 22;	/* EINVAL */
		}
		StreamDev &sd = hp.streams[i];
		std::memset(&sd, 0, sizeof sd);
		sd.xa_off = s.xa_off;
		sd.pcm_off = s.pcm_off;
		sd.blocks = s.blocks;
		sd.pcm_len = s.pcm_len;
		std::memcpy(sd.prev, s.prev, sizeof sd.prev);
		int b = bucket_of(s.bits, s.channels);
		hp.bucket[i] = (uint8_t)b;
		members[b].push_back((uint32_t)i);
		uint64_t xa_end = s.xa_off + (uint64_t)s.blocks *
		    (uint64_t)(block_bytes(s.bits) * s.channels);
		uint64_t pcm_end = s.pcm_off + s.pcm_len;
		uint64_t &src = kind == kKindDecode ? hp.src_need : hp.dst_need;
		uint64_t &dst = kind == kKindDecode ? hp.dst_need : hp.src_need;
		src = std::max(src, xa_end);
		dst = std::max(dst, pcm_end);
	}

	for (int b = 0; b < 6; b++) {
		hp.tile_begin[b] = (uint32_t)hp.tiles.size();
		hp.ns[b] = 1;
		hp.alt_ns[b] = 0;
		hp.order_begin[b] = (uint32_t)hp.order.size();
		std::vector<uint32_t> &o = members[b];
		if (o.empty())
			continue;
		/* longest first: the streams still active at step j are a prefix */
		std::stable_sort(o.begin(), o.end(), [&](uint32_t x, uint32_t y) {
			return hp.streams[x].blocks > hp.streams[y].blocks;
		});
		const uint32_t order0 = (uint32_t)hp.order.size();
		hp.order.insert(hp.order.end(), o.begin(), o.end());

		if (kind == kKindSearch)
			continue;	/* one warp per stream-channel, in this order: no tiles */
		if (kind == kKindEncode) {
			const uint32_t tbe = (uint32_t)kEncTBE;
			size_t active = o.size();
			for (uint32_t j = 0; active > 0; j++) {
				while (active > 0 &&
				    (uint64_t)j * tbe >= hp.streams[o[active - 1]].blocks)
					active--;
				for (size_t k = 0; k < active; k++) {
					TileEnt te = { o[k], 1u, j * tbe, 0u };
					hp.tiles.push_back(te);
				}
			}
			continue;
		}

		const bool forced = force_strips == 1 || force_strips == kDecWide;
		hp.ns[b] = forced ? force_strips : 1;
		hp.alt_ns[b] = !forced && o.size() >= kDecWideMinStreams ? kDecWide : 0;
		/* carry slots: one per strip of the finest shape in use; strip j of
		 * a stream uses slot_base + j in either shape */
		uint64_t items = 0;
		for (size_t k = 0; k < o.size(); k++)
			items += hp.streams[o[k]].blocks;
		const bool seg = items >= seg_min_items;
		has_seg[b] = seg;
		const uint32_t fine = std::min(strip_blocks(std::max(hp.ns[b], hp.alt_ns[b]), bucket_ch(b)),
		    seg ? kSegItems : ~0u);
		for (size_t k = 0; k < o.size(); k++) {
			StreamDev &sd = hp.streams[o[k]];
			sd.slot_base = hp.n_slots;
			hp.n_slots += (sd.blocks + fine - 1) / fine;
		}
		emit_decode_tiles(hp, o, order0, hp.ns[b], bucket_ch(b));
	}
	hp.tile_begin[6] = (uint32_t)hp.tiles.size();
	hp.order_begin[6] = (uint32_t)hp.order.size();
	for (int b = 0; b < 6; b++) {
		hp.alt_begin[b] = (uint32_t)hp.tiles.size();
		if (kind == kKindDecode && hp.alt_ns[b] != 0)
			emit_decode_tiles(hp, members[b], hp.order_begin[b], hp.alt_ns[b], bucket_ch(b));
	}
	hp.alt_begin[6] = (uint32_t)hp.tiles.size();
	for (int b = 0; b < 6; b++) {
		hp.seg_begin[b] = (uint32_t)hp.tiles.size();
		if (kind == kKindDecode && has_seg[b])
			emit_seg_tiles(hp, members[b]);
	}
	hp.seg_begin[6] = (uint32_t)hp.tiles.size();
	return 0;
}

} /* namespace xa */
#endif
