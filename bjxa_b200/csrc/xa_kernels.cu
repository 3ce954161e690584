/*
 * xa_kernels.cu -- sm_100a kernels and the device-resident batch layer
 * (bjxa_plan_*, bjxa_gpu_* of include/bjxa_batch.h).
 *
 * Replaces the serial block loops of the reference
 * (/root/reference/src/libbjxa.c:602-661 decode, :759-819 encode) with one
 * launch per (bits, channels) class of a batch.  The tile algorithm is in
 * xa_tile.h; this file adds what only exists on the GPU: the bulk-async
 * (TMA, cp.async.bulk) load of a tile's contiguous source bytes into shared
 * memory behind an mbarrier, the CTA barriers between phases, the atomic
 * ticket that orders tiles, and the CUDA runtime plumbing.
 *
 * There is no CPU fallback: without a CUDA device every entry point fails
 * with ENODEV.
 */
#include <cuda_runtime.h>

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <type_traits>
#include <vector>

#include "../../include/bjxa_batch.h"
#include "bjxa_internal.h"
#include "xa_plan.h"
#include "xa_walk.h"

using namespace xa;

/* ---- PTX helpers: mbarrier + bulk async copy (Hopper/Blackwell) ---------- */

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
	return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
	/* make the init visible to the async proxy before any bulk copy uses it */
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
	    :: "r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "XA_WAIT_%=:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	    "@p bra XA_DONE_%=;\n"
	    "bra XA_WAIT_%=;\n"
	    "XA_DONE_%=:\n"
	    "}\n" :: "r"(bar), "r"(parity) : "memory");
}

/* the same for a warp that has time: it sleeps between looks, so that its polling
 * does not take shared-memory pipe cycles from a warp on the critical path (the
 * chain form's stepper ran 1.7 x slower next to two tightly polling warps) */
__device__ __forceinline__ void mbar_wait_idle(uint32_t bar, uint32_t parity, uint32_t ns)
{
	for (;;) {
		uint32_t done;
		asm volatile(
		    "{\n"
		    ".reg .pred p;\n"
		    "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
		    "selp.u32 %0, 1, 0, p;\n"
		    "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
		if (done)
			return;
		if (ns != 0)
			__nanosleep(ns);
	}
}

/* global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned;
 * completion is signalled on the mbarrier as transaction bytes */
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src,
    uint32_t bytes, uint32_t bar)
{
	asm volatile(
	    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes"
	    " [%0], [%1], %2, [%3];"
	    :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

/* ---- kernels -------------------------------------------------------------- */

/* consumer-only CTA barrier (the producer warp never takes part) */
__device__ __forceinline__ void consumer_sync()
{
	asm volatile("bar.sync 1, %0;" :: "n"(kDecThreads) : "memory");
}

/*
 * Persistent, warp-specialised decode kernel, generic over the tile form.
 * Grid = (CTAs that fit one SM) x (SM count).  Per CTA:
 *   loader     (1 warp) lane 0 draws tile tickets in order; lane i builds the
 *              context of strip i and starts the bulk-async (TMA) copy of that
 *              strip's contiguous XA bytes into the next free stage buffer,
 *              Tile::kStages tiles ahead;
 *   scanner    (1 warp) when a tile's bytes have landed, scans its profile
 *              bytes for the heads of chains and publishes the tile on its
 *              "ready" mbarrier;
 *   consumers  (kDecThreads threads) wait for "ready" and run the tile's
 *              phases (xa_tile.h).  Direct forms: the tile's walker warp
 *              (the warps take turns) walks the chains, every warp decodes its
 *              share of the 16-byte units of the cut blocks, everything is
 *              stored straight from registers, then one arrival per warp on the
 *              stage's "empty" mbarrier; no CTA barrier at all.  Staged form
 *              (chain-heavy stereo): rows, barrier, interleaving store, barrier.
 */
#ifndef XA_DEC_STAGED_STAGES
#define XA_DEC_STAGED_STAGES 2
#endif
constexpr int kDecStagedStages = XA_DEC_STAGED_STAGES;
constexpr int kDecBlock = kDecThreads + 64;	/* consumers + loader warp + scanner warp */

/*
 * Census of one class of a batch: looks at 8192 pseudo-randomly chosen blocks
 * and writes which of the launched tile forms is to run, from the share of
 * chain blocks (filters 1..4) among them:
 *   bit 0  the staged stereo form instead of the direct one (share >= staged)
 *   bit 1  the wide tile list instead of the long-strip one  (share >= wide)
 *   bit 2  (alone) the pooled form over the long-strip list  (pool <= share < wide)
 *   bit 3  (alone) the split form: direct form without walkers, then the dense
 *          chain walkers of xa_walk_kernel                    (split <= share < wide)
 *   bit 4  (alone) the relay form: direct form whose walker warps hand their
 *          stragglers to xa_walk_kernel                       (relay <= share < split)
 *   bit 5  (alone) the segment form, xa_seg_kernel            (seg <= share < seg_below;
 *          asked first)
 *   bit 6  (alone) the chain form, xa_chain_kernel: mono data (all but) without cut
 *          blocks                                             (chain <= share; asked second)
 * A threshold above 1000 permille switches that choice off.  One CTA, no
 * atomics, nothing to clear; all of a thread's loads in flight together.  Measured crossovers:
 * profiles/history_r1.md.
 */
constexpr uint32_t kCensusThreads = 1024, kCensusPerThread = 8;
/* stereo: direct form below this share of chain blocks, staged form above */
constexpr uint32_t staged_permille(int bits) { return bits == 4 ? 300u : bits == 6 ? 500u : 650u; }
/* wide tiles: the more streams, the more chains they keep in flight, so the
 * share of chain blocks from which they win drops with the size of the class */
constexpr uint32_t kWideManyStreams = 8192;
constexpr uint32_t kWidePermilleMono[2] = { 985, 930 }, kWidePermilleStereo[2] = { 920, 700 };
constexpr uint32_t kNever = 1001;
enum { kFormStaged = 1, kFormWide = 2, kFormPool = 4, kFormSplit = 8, kFormRelay = 16, kFormSeg = 32, kFormChain = 64 };
/* pooled walkers instead of one walker warp per tile: from this share of chain
 * blocks (measured crossovers: profiles/history_r1.md), for classes of at least
 * kPoolMinTiles tiles -- below that the launch is too short to care */
#ifndef XA_POOL_PERMILLE_MONO
#define XA_POOL_PERMILLE_MONO 300
#endif
#ifndef XA_POOL_PERMILLE_STEREO
#define XA_POOL_PERMILLE_STEREO 300
#endif
constexpr uint32_t kPoolPermilleMono = XA_POOL_PERMILLE_MONO, kPoolPermilleStereo = XA_POOL_PERMILLE_STEREO;
constexpr uint32_t kPoolMinTiles = 296;

__global__ void __launch_bounds__(kCensusThreads)
xa_census_kernel(const uint8_t *src, const StreamDev *streams, const uint32_t *order,
    uint32_t n_streams, uint32_t block_bytes_one, uint32_t ch, uint32_t staged_permille,
    uint32_t wide_permille, uint32_t pool_permille, uint32_t split_permille,
    uint32_t relay_permille, uint32_t seg_permille, uint32_t seg_below, uint32_t chain_permille,
    uint32_t *choice)
{
	__shared__ uint32_t warp_sum[kCensusThreads / 32];
	const uint32_t tid = threadIdx.x;
	uint32_t chains = 0, h[kCensusPerThread], who[kCensusPerThread], prof[kCensusPerThread];
#pragma unroll
	for (uint32_t k = 0; k < kCensusPerThread; k++) {
		uint32_t x = (tid * kCensusPerThread + k) * 2654435761u;
		x ^= x >> 15;
		x *= 2246822519u;
		h[k] = x ^ x >> 13;
		who[k] = order[h[k] % n_streams];
	}
#pragma unroll
	for (uint32_t k = 0; k < kCensusPerThread; k++) {
		const StreamDev &s = streams[who[k]];
		/* any block-channel of the stream */
		const uint32_t q = (h[k] >> 7) % (s.blocks * ch);
		prof[k] = src[s.xa_off + (uint64_t)q * block_bytes_one];
	}
#pragma unroll
	for (uint32_t k = 0; k < kCensusPerThread; k++)
		chains += block_kind(prof[k]) == kChain;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
		chains += __shfl_xor_sync(0xffffffffu, chains, o);
	if ((tid & 31u) == 0)
		warp_sum[tid >> 5] = chains;
	__syncthreads();
	if (tid == 0) {
		uint32_t total = 0;
		for (uint32_t w = 0; w < kCensusThreads / 32; w++)
			total += warp_sum[w];
		const uint32_t permille = total * 1000u / (kCensusThreads * kCensusPerThread);
		const uint32_t staged = permille >= staged_permille ? kFormStaged : 0u;
		*choice = permille >= seg_permille && permille < seg_below ? (uint32_t)kFormSeg :
		    permille >= chain_permille ? (uint32_t)kFormChain :
		    permille >= wide_permille ? kFormWide | staged :
		    permille >= split_permille ? (uint32_t)kFormSplit :
		    permille >= relay_permille ? (uint32_t)kFormRelay :
		    permille >= pool_permille ? (uint32_t)kFormPool : staged;
	}
}

/*
 * Relay form: the tile's walker warp, every lane a chain at a time as in
 * phase_walk_warp, but it does not sit out the tile's longest chains: once the
 * heads are dealt out and no more than kRelayWind lanes are still walking, their
 * chains are handed to the second pass (xa_walk_kernel) and the warp lets go of
 * the stage.  The first turns are exempt, so that a tile with a handful of short
 * chains walks them itself.
 */
template <class Tile>
__device__ __forceinline__ void
relay_walk_warp(const Tile &t, const uint16_t *heads, uint32_t n, uint32_t *next, uint32_t first)
{
	typename Tile::Walk w;
	bool have = false;
	uint32_t i = first;		/* this lane's first chain: lane + 32 * (which walker warp) */
	for (uint32_t turn = 0;; turn++) {
		if (!have && i < n) {
			t.walk_begin(w, heads[i]);
			have = true;
		}
		const uint32_t am = __ballot_sync(0xffffffffu, have);
		if (am == 0)
			break;
		if (turn >= 4u && (uint32_t)__popc(am) <= kRelayWind &&
		    *(volatile uint32_t *)next >= n) {
			if (have)
				t.walk_hand_on(w);
			break;
		}
		if (have && !t.walk_block(w)) {
			have = false;
			i = take_next(next);
		}
	}
}

/* one tile, direct forms: walkers and units, no CTA barrier */
template <class Tile>
__device__ __forceinline__ typename std::enable_if<!Tile::kStaged>::type
consume_tile(Tile &t, typename Tile::Smem &sm, int s, uint32_t it, uint32_t tid)
{
	/* the warps take turns at walking a tile's chains: up to kStages tiles, and
	 * so kStages walker warps, are in flight in a CTA.  A tile with chains has
	 * its units decoded by the other seven warps, so that the walker warp is
	 * not the last one to let go of the stage by a whole share of units. */
	constexpr uint32_t NW = kDecThreads / 32u;
	const uint32_t walker = it % NW, warp = tid >> 5;
	/* relay form: a tile with many chains gets kRelayWalkers walker warps (the
	 * scanner has set the draw counter accordingly), which share its list */
	const uint32_t nw = t.relay && sm.n_heads[s] > kRelayManyHeads ? kRelayWalkers : 1u;
	const uint32_t rel = (warp + NW - walker) % NW;		/* 0 .. nw-1: a walker */
	if (sm.n_heads[s] == 0)
		t.phase_units(tid, kDecThreads);
	else if (rel < nw && t.relay)
		relay_walk_warp(t, sm.heads[s], sm.n_heads[s], &sm.next_head[s], (tid & 31u) + 32u * rel);
	else if (rel < nw)
		t.phase_walk_warp(tid & 31u, sm.heads[s], sm.n_heads[s], &sm.next_head[s]);
	else
		t.phase_units((rel - nw) * 32u + (tid & 31u), kDecThreads - 32u * nw);
	__syncwarp();
	if ((tid & 31u) == 0)
		mbar_arrive(smem_u32(&sm.empty[s]));
}

/* one tile, staged form: rows (walkers and cut blocks in the same phase),
 * barrier, interleaving store, barrier */
template <class Tile>
__device__ __forceinline__ typename std::enable_if<Tile::kStaged>::type
consume_tile(Tile &t, typename Tile::Smem &sm, int s, uint32_t it, uint32_t tid)
{
	/* one strip per tile: the store needs three words of the context; every
	 * thread takes them BEFORE the barrier so that the stage can go back to the
	 * loader ahead of the store */
	const uint64_t c0_out0 = sm.ctx[s][0].out0;
	const uint32_t c0_nq = sm.ctx[s][0].nq, c0_valid = sm.ctx[s][0].out_valid;
	t.phase_walk(tid, kDecThreads, sm.heads[s], sm.n_heads[s], it * 96u);
	t.phase_a(tid, kDecThreads);
	consumer_sync();
	if (Tile::G::kNS == 1) {
		if (tid == 0)
			mbar_arrive(smem_u32(&sm.empty[s]));
		t.phase_store_one(tid, kDecThreads, c0_out0, c0_nq, c0_valid);
		consumer_sync();	/* rows are free for the next tile */
	} else {
		t.phase_store(tid, kDecThreads);
		consumer_sync();
		if (tid == 0)
			mbar_arrive(smem_u32(&sm.empty[s]));
	}
}

/* ---- the pooled form's shared-memory protocol (xa_decode_pool_kernel below) ---- */
#ifndef XA_POOL_UNIT_WARPS
#define XA_POOL_UNIT_WARPS 3
#endif
#ifndef XA_POOL_WALK_WARPS
#define XA_POOL_WALK_WARPS 3
#endif
constexpr int kPoolUnitWarps = XA_POOL_UNIT_WARPS, kPoolWalkWarps = XA_POOL_WALK_WARPS;
constexpr int kPoolCtas = XA_POOL_CTAS;			/* per SM */
constexpr int kPoolConsumers = (kPoolUnitWarps + kPoolWalkWarps) * 32;
constexpr int kPoolBlock = kPoolConsumers + 64;		/* + scanner warp + loader warp */

template <class Tile>
struct PoolSmem : Tile::Smem {
	uint32_t draw[Tile::kStages];
	uint32_t pending[Tile::kStages];	/* chains of the tile not finished yet */
	uint32_t oldest;			/* the stage the loader waits for */
	uint32_t done;				/* every tile has been offered */
};

template <class SM>
__device__ __forceinline__ void pool_set_oldest(SM &sm, uint32_t s)
{
	*(volatile uint32_t *)&sm.oldest = s;
}

template <class SM>
__device__ __forceinline__ void pool_finish(SM &sm)
{
	__threadfence_block();
	*(volatile uint32_t *)&sm.done = 1u;
}

/* scanner lane 0: the tile in stage s has `count` chains (heads[] written) */
template <class SM>
__device__ __forceinline__ void pool_offer(SM &sm, int s, uint32_t gen, uint32_t count)
{
	if (count == 0) {
		mbar_arrive(smem_u32(&sm.empty[s]));	/* on behalf of the chains */
		return;
	}
	*(volatile uint32_t *)&sm.pending[s] = count;
	__threadfence_block();		/* heads[], pending before the offer */
	*(volatile uint32_t *)&sm.draw[s] = (gen & 0xffu) << 24 | count << 12;
}

/*
 * Loader warp: tickets, strip contexts, bulk-async copies.  The ticket and the
 * records of the NEXT tile (ticket -> tile table -> issue order -> stream
 * record: four dependent global round trips) are fetched while the warp waits
 * for a free stage.  POOL: the stage the loader wants back next is made known to
 * the pooled walkers, who then finish that tile's chains first.
 */
template <class Tile, bool POOL, class SM>
__device__ __forceinline__ void
loader_warp(const DecodeParams &p, SM &sm, const uint32_t lane, const bool want_prev = false)
{
	typedef typename Tile::G G;
	constexpr int NS = G::kNS;
	constexpr int kStages = Tile::kStages;
	unsigned long long t = 0;
	TileEnt te = { 0u, 0u, 0u, 0u };
	StripCtx c;
	auto prefetch = [&]() {
		/* the counter is preset to ~0 with first_bad[]: old + 1 = ticket */
		if (lane == 0)
			t = atomicAdd(p.ticket, 1ULL) + 1ULL;
		t = __shfl_sync(0xffffffffu, t, 0);
		if (t < p.n_tiles) {
			te = p.tiles[t];
			if (lane < te.count)
				make_strip_ctx<G::kBits, G::kCh, G::kTBQ, NS>(c, p,
				    p.order[te.first + lane], te.j, lane, want_prev);
		}
	};
	prefetch();
	for (uint32_t it = 0;; it++) {
		const int s = (int)(it % kStages);
		if (it >= (uint32_t)kStages) {
			if constexpr (POOL) {
				if (lane == 0)
					pool_set_oldest(sm, (uint32_t)s);
			}
			mbar_wait(smem_u32(&sm.empty[s]), (it / kStages - 1) & 1);
		}
		const uint32_t full = smem_u32(&sm.full[s]);
		if (t >= p.n_tiles) {
			if (lane == 0) {
				sm.tile_flags[s] = kCtxEnd;
				mbar_arrive(full);
			}
			return;
		}
		uint32_t bulk = 0, tail = 0;
		const unsigned char *src = p.src;
		if (lane < te.count) {
			sm.ctx[s][lane] = c;
			bulk = c.bulk;
			tail = c.flags & kCtxTail;
			src += c.a0;
		}
		uint32_t total = bulk;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1)
			total += __shfl_xor_sync(0xffffffffu, total, o);
		const uint32_t any_tail = __ballot_sync(0xffffffffu, tail != 0);
		if (lane == 0) {
			sm.tile_flags[s] = any_tail ? kCtxTail : 0u;
			sm.n_strips[s] = te.count;
			if (total)
				mbar_expect_tx(full, total);
			else
				mbar_arrive(full);
		}
		__syncwarp();
		if (bulk)
			bulk_g2s(smem_u32(sm.in[s]) + lane * G::SLOT, src, bulk, full);
		prefetch();
	}
}

/*
 * Scanner warp: the heads of chains of every landed tile.  POOL: the tile's
 * chains are then put up for the pooled walkers (pool_offer).
 */
template <class Tile, bool POOL, class SM>
__device__ __forceinline__ void
scanner_warp(const DecodeParams &p, SM &sm, const uint32_t lane, const bool split = false,
    const bool relay = false)
{
	typedef typename Tile::G G;
	constexpr int NS = G::kNS;
	constexpr int kStages = Tile::kStages;
	for (uint32_t it = 0;; it++) {
		const int s = (int)(it % kStages);
		mbar_wait(smem_u32(&sm.full[s]), (it / kStages) & 1);
		const uint32_t tf = sm.tile_flags[s];
		if (tf & kCtxEnd) {
			if (lane == 0) {
				if constexpr (POOL)
					pool_finish(sm);
				mbar_arrive(smem_u32(&sm.ready[s]));
			}
			return;
		}
		Tile t(p, sm, s);
		if (tf & kCtxTail) {	/* only at the very end of the arena */
			t.load_tail(lane, 32, sm.in[s]);
			__syncwarp();
		}
		uint32_t count = 0, mine = 0;
		if (NS == 1) {
			/* one strip: one item (block, or pair of blocks) per lane
			 * and step; "the item in front is a walker's too" comes out
			 * of the ballots */
			constexpr int LAG = Tile::kLag;	/* item q follows item q - LAG */
			const uint32_t nq = Tile::kStaged ? sm.ctx[s][0].nq :
			    sm.ctx[s][0].nq / G::kCh;
			/* items -LAG..-1: no chain channels -- or, in the split form, the
			 * ones the loader found in front of the strip */
			uint32_t prev = split || relay ? (sm.ctx[s][0].flags >> kCtxPrevShift & 3u) : 0u;
			for (uint32_t base = 0; base < nq; base += 32) {
				const uint32_t q = base + lane;
				/* chain channels of item q and of the item in front;
				 * a head continues none of that one's chains */
				const uint32_t cm = t.chain_mask(q);
				const uint32_t up = __shfl_up_sync(0xffffffffu, cm, LAG);
				const uint32_t old = __shfl_sync(0xffffffffu, prev,
				    (32 - LAG + lane) & 31u);
				const uint32_t before = lane >= (uint32_t)LAG ? up : old;
				const bool h = cm != 0 && (cm & before) == 0;
				const uint32_t mh = __ballot_sync(0xffffffffu, h);
				if (split) {
					/* split form: the heads go out as a bitmap, word
					 * base / 32 of it in lane base / 32 */
					if (lane == base / 32u)
						mine = mh;
				} else if (h) {
					sm.heads[s][count + __popc(mh & ((1u << lane) - 1u))] =
					    (uint16_t)q;
				}
				count += __popc(mh);
				prev = cm;
			}
			if (split) {
				if (count != 0) {
					/* one record per tile with heads, one 128-byte store
					 * (xa_walk.h); the counter is preset to ~0 */
					uint32_t slot = 0;
					if (lane == 0)
						slot = atomicAdd(p.live_count, 1u) + 1u;
					slot = __shfl_sync(0xffffffffu, slot, 0);
					const StripCtx &c = sm.ctx[s][0];
					const uint64_t g0 = c.a0 + c.in_base;
					uint32_t w = mine;
					if (lane == kRecXaLo) w = (uint32_t)g0;
					if (lane == kRecXaHi) w = (uint32_t)(g0 >> 32);
					if (lane == kRecOutLo) w = (uint32_t)c.out0;
					if (lane == kRecOutHi) w = (uint32_t)(c.out0 >> 32);
					if (lane == kRecStream) w = c.stream;
					if (lane == kRecFirstEb) w = c.first_eb;
					if (lane == kRecBlocks) w = c.blocks;
					if (lane > kRecBlocks) w = 0;
					reinterpret_cast<uint32_t *>(&p.live[slot])[lane] = w;
				}
				count = 0;	/* the consumers: units only */
			}
		} else {
			const uint32_t nq = t.n_strips * Tile::SCAN;
			for (uint32_t base = 0; base < nq; base += 32) {
				const uint32_t q = base + lane;
				const bool h = q < nq && t.is_head(q);
				const uint32_t m = __ballot_sync(0xffffffffu, h);
				if (h)
					sm.heads[s][count + __popc(m & ((1u << lane) - 1u))] =
					    (uint16_t)q;
				count += __popc(m);
			}
		}
		__syncwarp();
		if (lane == 0) {
			sm.n_heads[s] = count;
			if constexpr (POOL) {
				pool_offer(sm, s, it / kStages, count);
			} else {
				sm.next_head[s] = 32;	/* chains 0..31 start with the lanes */
			}
			mbar_arrive(smem_u32(&sm.ready[s]));
		}
	}
}

/*
 * MODE: kModePlain -- the tile form as it stands; kModeSplit -- the direct form as
 * the split form's first pass (no walkers, heads written out); kModeRelay -- the
 * direct form over relay tiles (DecTile<..., true>: stragglers handed on).  Compile
 * time, so that the plain kernels carry nothing of the other two.
 */
enum { kModePlain = 0, kModeSplit = 1, kModeRelay = 2 };

template <class Tile, int MODE>
__global__ void __launch_bounds__(kDecBlock, Tile::kMinCtas)
xa_decode_kernel(const DecodeParams p)
{
	constexpr int kStages = Tile::kStages;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	typename Tile::Smem &sm = *reinterpret_cast<typename Tile::Smem *>(smem_raw);
	const uint32_t tid = threadIdx.x;

	constexpr bool split = MODE == kModeSplit, relay = MODE == kModeRelay;
	if (MODE == kModeSplit ? p.split != 2 && (p.choice == NULL || *p.choice != kFormSplit) :
	    MODE == kModeRelay ? p.relay != 2 && (p.choice == NULL || *p.choice != kFormRelay) :
	    p.choice != NULL && *p.choice != p.want)
		return;		/* the census picked another tile form */

	if (tid == 0) {
		for (int s = 0; s < kStages; s++) {
			mbar_init(smem_u32(&sm.full[s]), 1);
			mbar_init(smem_u32(&sm.ready[s]), 1);
			/* direct forms: one arrival per consumer warp; staged form: one
			 * arrival after the CTA barrier that ends the row phase */
			mbar_init(smem_u32(&sm.empty[s]), Tile::kStaged ? 1 : kDecThreads / 32);
		}
	}
	__syncthreads();

	if (tid >= kDecThreads + 32) {
		loader_warp<Tile, false>(p, sm, tid - (kDecThreads + 32), split || relay);
		return;
	}
	if (tid >= kDecThreads) {
		scanner_warp<Tile, false>(p, sm, tid - kDecThreads, split, relay);
		return;
	}

	/* ---- consumer warps ---- */
	for (uint32_t it = 0;; it++) {
		const int s = (int)(it % kStages);
		mbar_wait(smem_u32(&sm.ready[s]), (it / kStages) & 1);
		if (sm.tile_flags[s] & kCtxEnd)
			return;
		/* already complete; waiting on it orders the bulk copy's bytes for us */
		mbar_wait(smem_u32(&sm.full[s]), (it / kStages) & 1);
		Tile t(p, sm, s);
		consume_tile(t, sm, s, it, tid);
	}
}

/*
 * Pooled form, for data with many chain blocks.  In xa_decode_kernel a tile's
 * chains are walked by ONE warp, and a CTA has as many walker warps at work as
 * it has tiles in its ring; on chain-heavy data most lanes of those warps idle
 * behind the longest chains while five of eight warps wait for a stage.  Here
 * the roles are fixed and the chains of ALL tiles in the ring form one pool:
 *   unit warps    (kPoolUnitWarps) decode the cut blocks of every tile, in order;
 *   walker warps  (kPoolWalkWarps) never look at tiles: every idle LANE draws
 *                 the next chain of any tile on offer (oldest stage first) and
 *                 walks it one block per turn of the warp's loop, so lanes of
 *                 one warp hold chains of different tiles side by side.
 * A stage goes back to the loader when the unit warps have arrived and the
 * tile's last chain has been finished (a countdown in shared memory).
 *
 * sm.draw[s] = generation:8 | chains:12 | next:12 -- one atomicAdd both draws a
 * chain and tells whether the draw was good, whichever tile the stage holds by
 * then; a good draw pins the tile (its countdown cannot reach zero before the
 * drawn chain is finished).
 */
template <class Tile>
__device__ __forceinline__ void
pool_walker_warp(const DecodeParams &p, PoolSmem<Tile> &sm)
{
	constexpr int kStages = Tile::kStages;
	typename Tile::Walk w;
	int ls = 0;			/* the stage this lane's chain lives in */
	uint32_t vain = 0;		/* turns spent waiting for a carry */
	bool have = false, saw_done = false;

	for (;;) {
		if (!have) {
			const uint32_t s0 = *(volatile uint32_t *)&sm.oldest;
			for (int k = 0; k < kStages; k++) {
				int s = (int)s0 + k;
				if (s >= kStages)
					s -= kStages;
				const uint32_t peek = *(volatile uint32_t *)&sm.draw[s];
				if ((peek & 0xfffu) >= (peek >> 12 & 0xfffu))
					continue;
				const uint32_t got = atomicAdd(&sm.draw[s], 1u);
				if ((got & 0xfffu) >= (got >> 12 & 0xfffu))
					continue;
				/* a good draw: the tile of generation got >> 24 is pinned.
				 * Its barriers have completed; waiting on "full" orders the
				 * bulk copy's bytes for this thread, the fence the scanner's */
				mbar_wait(smem_u32(&sm.full[s]), (got >> 24) & 1u);
				__threadfence_block();
				ls = s;
				Tile t(p, sm, s);
				t.walk_begin(w, sm.heads[s][got & 0xfffu]);
				have = true;
				vain = 0;
				break;
			}
		}
		if (!__any_sync(0xffffffffu, have)) {
			if (saw_done)
				return;		/* nothing on offer after the last tile was */
			saw_done = *(volatile uint32_t *)&sm.done != 0;
			if (!saw_done)
				__nanosleep(64);
			continue;
		}
		bool moved = false;
		if (have) {
			Tile t(p, sm, ls);
			if (w.need != 0) {
				/* the strip's carry: looked for once per turn, never waited
				 * for -- the lane that produces it may be in this very warp.
				 * Bounded like mailbox_get. */
				t.walk_carry(w);
				if (w.need != 0 && ++vain > (1u << 25))
					t.walk_give_up(w);
			}
			if (w.need == 0) {
				moved = true;
				if (!t.walk_block(w)) {
					have = false;
					/* the chain's reads of the stage are over */
					if (atomicSub(&sm.pending[ls], 1u) == 1u)
						mbar_arrive(smem_u32(&sm.empty[ls]));
				}
			}
		}
		if (!__any_sync(0xffffffffu, moved))
			__nanosleep(128);	/* every chain of the warp waits for its carry */
	}
}

template <class Tile>
__global__ void __launch_bounds__(kPoolBlock, kPoolCtas)
xa_decode_pool_kernel(const DecodeParams p)
{
	constexpr int kStages = Tile::kStages;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	PoolSmem<Tile> &sm = *reinterpret_cast<PoolSmem<Tile> *>(smem_raw);
	const uint32_t tid = threadIdx.x;

	if (p.choice != NULL && *p.choice != p.want)
		return;		/* the census picked another tile form */

	if (tid == 0) {
		for (int s = 0; s < kStages; s++) {
			mbar_init(smem_u32(&sm.full[s]), 1);
			mbar_init(smem_u32(&sm.ready[s]), 1);
			/* one arrival per unit warp, one for the tile's chains */
			mbar_init(smem_u32(&sm.empty[s]), kPoolUnitWarps + 1);
			sm.draw[s] = 0;
			sm.pending[s] = 0;
		}
		sm.oldest = 0;
		sm.done = 0;
	}
	__syncthreads();

	if (tid >= kPoolConsumers + 32) {
		loader_warp<Tile, true>(p, sm, tid - (kPoolConsumers + 32));
		return;
	}
	if (tid >= kPoolConsumers) {
		scanner_warp<Tile, true>(p, sm, tid - kPoolConsumers);
		return;
	}
	if (tid >= kPoolUnitWarps * 32) {
		pool_walker_warp<Tile>(p, sm);
		return;
	}

	/* ---- unit warps ---- */
	for (uint32_t it = 0;; it++) {
		const int s = (int)(it % kStages);
		mbar_wait(smem_u32(&sm.ready[s]), (it / kStages) & 1);
		if (sm.tile_flags[s] & kCtxEnd)
			return;
		mbar_wait(smem_u32(&sm.full[s]), (it / kStages) & 1);
		Tile t(p, sm, s);
		t.phase_units(tid, kPoolUnitWarps * 32);
		__syncwarp();
		if ((tid & 31u) == 0)
			mbar_arrive(smem_u32(&sm.empty[s]));
	}
}

/* ---- split form, pass 2: dense chain walkers (xa_walk.h) ---------------------- */
/*
 * Every lane walks one chain (stereo: one run) at a time and draws the next head
 * the moment its chain ends, so all 32 lanes of a warp decode a block in every
 * turn of the loop whatever the chains' lengths.  A lane reads its chain straight
 * from the arena: 16-byte cp.async chunks into a private ring in shared memory
 * (its address bits pick the slot), always one item ahead of the decode.  The
 * item's PCM goes to a row of shared memory and leaves the warp eight (stereo:
 * four) rows per store instruction -- a lane storing its own 64 bytes would
 * touch 32 different lines per instruction, four times (tools/walk_proto.cu has
 * the measurements behind both choices).
 *
 * A warp works through the records of pass 1 (one per tile with heads) at stride
 * "all warps of the grid"; a record's bitmap is spread into a list of head
 * indices in shared memory, from which idle lanes take the next in order.
 */
template <int BITS, int CH>
struct WalkCfg {
	typedef Walk<BITS, CH> W;
	/* bytes the ring holds, from the current item's first: the item in front of
	 * the decode and the profile byte(s) behind it */
	static constexpr int LOOK = 2 * W::STEP + W::PEEK;
	static constexpr int SPAN = (15 + LOOK + 15) / 16 * 16;
	static constexpr int RING = SPAN <= 64 ? 64 : SPAN <= 128 ? 128 : 256;
	static constexpr int SLOT = RING + 16;	/* lane stride: spreads equal offsets over the banks */
	/* most chunks a lane asks for in one turn: a new chain's first item */
	static constexpr int KMAX = (15 + W::STEP + W::PEEK + 15) / 16;
	static constexpr int kThreads = CH == 1 ? 512 : 256;
	static constexpr int kCtas = 2;			/* per SM */
	static constexpr int kWarps = kThreads / 32;
	static constexpr int kListLen = 256;		/* heads per tile at most */
	static constexpr size_t kSmem = (size_t)kThreads * (SLOT + W::OUT) +
	    (size_t)kWarps * (kListLen * 2 + 32);
	static_assert(SPAN <= 256, "ring");
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}

/* the same past the L1: for kernels whose shared memory leaves the L1 too small to
 * hold a chunk's neighbour until it is asked for (the segment form's 6/8-bit stereo
 * classes: +19-30 %; with a larger L1 it costs 1-6 %, profiles/history_r2.md) */
__device__ __forceinline__ void cp_async16_cg(uint32_t dst, const void *src)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}

/* loads whose results must not be asked for before the turn's decode is over: as
 * PTX, so that the compiler neither re-extends the byte nor moves its first use up */
__device__ __forceinline__ uint32_t ldg_u8(const uint8_t *p)
{
	uint32_t v;
	asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}
__device__ __forceinline__ uint32_t ldg_u32(const uint8_t *p)
{
	uint32_t v;
	asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}

/* 16 bytes from global memory, as PTX for the same reason */
__device__ __forceinline__ uint4 ldg_u128(const void *p)
{
	uint4 v;
	asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
	    : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
	return v;
}

/*
 * RELAY = false: the split form's second pass -- the heads of ALL chains, as
 * bitmaps, one record per tile; a warp spreads a record into a list and its lanes
 * draw from it.  RELAY = true: the relay form's -- one RelayRec per unfinished
 * chain, with the state in front of it; lane L of the grid takes records L,
 * L + lanes, ... and always has the next one on its way.
 */
template <int BITS, int CH, bool RELAY>
__global__ void __launch_bounds__(WalkCfg<BITS, CH>::kThreads, WalkCfg<BITS, CH>::kCtas)
xa_walk_kernel(const DecodeParams p)
{
	typedef Walk<BITS, CH> W;
	typedef WalkCfg<BITS, CH> C;
	constexpr int BS = W::BS, STEP = W::STEP, OUT = W::OUT, RING = C::RING;
	constexpr uint32_t FULL = 0xffffffffu;
	constexpr uint32_t UPR = W::UNITS;		/* 16-byte units per row of PCM */
	constexpr uint32_t RPI = 32u / UPR;		/* rows per store instruction */

	if (RELAY ? p.relay != 2 && (p.choice == NULL || *p.choice != kFormRelay) :
	    p.split != 2 && (p.choice == NULL || *p.choice != kFormSplit))
		return;		/* the census picked a tile form that finishes its own chains */

	extern __shared__ __align__(16) unsigned char smem_raw[];
	/* read once, and opaque: the compiler would otherwise recompute them from
	 * the special register wherever registers are short */
	uint32_t tid, lane;
	asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
	asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
	const uint32_t warp = tid >> 5;
	uint8_t *const ringp = smem_raw + (size_t)tid * C::SLOT;
	const uint32_t ring = smem_u32(ringp);
	uint8_t *const rows = smem_raw + (size_t)C::kThreads * C::SLOT + (size_t)warp * 32 * OUT;
	uint16_t *const list = reinterpret_cast<uint16_t *>(smem_raw +
	    (size_t)C::kThreads * (C::SLOT + OUT)) + warp * C::kListLen;
	uint32_t *const ctx = reinterpret_cast<uint32_t *>(smem_raw +
	    (size_t)C::kThreads * (C::SLOT + OUT) + (size_t)C::kWarps * C::kListLen * 2) + warp * 8;
	/* this lane's row, and the rows it copies out: both swizzled by row */
	uint8_t *const row = rows + (size_t)lane * OUT;
	const uint32_t co_c = lane % UPR, co_r0 = lane / UPR;

	/* records: the counter was preset to ~0 */
	const uint32_t n_live = (RELAY ? *p.relay_count : *p.live_count) + 1u;
	const uint32_t stride = RELAY ? gridDim.x * C::kThreads : gridDim.x * C::kWarps;
	uint32_t next = RELAY ? blockIdx.x * C::kThreads + tid : blockIdx.x * C::kWarps + warp;
	bool rec_ready = next < n_live;
	uint32_t nrec = !RELAY && rec_ready ?
	    reinterpret_cast<const uint32_t *>(&p.live[next])[lane] : 0u;
	uint32_t pos = 0, cnt = 0;		/* the current record's list */
	/* relay: this lane's next record, fetched while the chain before it is walked */
	uint4 ra = make_uint4(0, 0, 0, 0), rb = ra;
	if (RELAY && rec_ready) {
		ra = ldg_u128(&p.relay_recs[next]);
		rb = ldg_u128(reinterpret_cast<const uint4 *>(&p.relay_recs[next]) + 1);
	}
	uint32_t m_front = 0;		/* relay: the new chain's first item must go on with these */
	bool fresh = false;
	/* whole 16-byte chunks of the arena; what lies behind is fetched bytewise */
	const uint64_t safe = p.src_bytes & ~(uint64_t)15;

	/* this lane's chain */
	uint64_t a = 0, o = 0;		/* arena addresses of the current item: XA, PCM */
	uint32_t left = 0;		/* items of the stream behind it */
	uint32_t stream = 0, m = 0;
	int fd = 0;			/* ring filled up to a + fd */
	int p0[CH], p1[CH];
	bool act = false;
#pragma unroll
	for (int c = 0; c < CH; c++)
		p0[c] = p1[c] = 0;

	for (;;) {
		/* (1) the chunks asked for in the turn before have landed */
		asm volatile("cp.async.wait_group 0;" ::: "memory");

		/* (2) this turn's item out of the ring; does the chain go on? */
		typename W::Item it;
		uint32_t nm = 0;
		bool more = false;
		if (act) {
#pragma unroll
			for (int c = 0; c < CH; c++) {
				const uint32_t at = (uint32_t)a + (uint32_t)(c * BS);
				it.prof[c] = ringp[at & (RING - 1)];
				const uint32_t pay = at + 1u, w0 = pay & ~3u, sh = (pay & 3u) * 8u;
				uint32_t prev = *reinterpret_cast<const uint32_t *>(ringp + (w0 & (RING - 1)));
#pragma unroll
				for (int i = 0; i < BITS; i++) {
					const uint32_t nx = *reinterpret_cast<const uint32_t *>(
					    ringp + ((w0 + 4u * (i + 1)) & (RING - 1)));
					it.pw[c][i] = __funnelshift_r(prev, nx, sh);
					prev = nx;
				}
			}
			if (RELAY && fresh) {
				/* what lies behind a strip was not known when the record was
				 * written: the chain may have ended with the strip */
				fresh = false;
				m = W::mask_of(it.prof);
				if ((m & m_front) == 0)
					act = false;
			} else if (CH == 2 && m == 0) {
				m = W::mask_of(it.prof);	/* a run's first item */
			}
			if (act && left != 0) {
				uint32_t nprof[CH];
#pragma unroll
				for (int c = 0; c < CH; c++)
					nprof[c] = ringp[((uint32_t)a + (uint32_t)(STEP + c * BS)) & (RING - 1)];
				nm = W::mask_of(nprof);
				more = CH == 1 ? nm != 0 : (nm & m) != 0;
			}
		}

		/* what the decode below needs of the chain, before a draw replaces it */
		const bool last = act && left == 0;
		/* bit 0: a whole row to copy out (a stream's last item goes by itself) */
		const unsigned long long d64 = act ? (unsigned long long)(p.dst + o) | (last ? 0ULL : 1ULL) : 0ULL;
		const uint32_t cur_stream = stream;
		if (CH == 2 && act)
			W::note_bad(p, stream, left, it.prof);

		/* (3) lanes whose chain ends here draw the next head */
		uint32_t needm = __ballot_sync(FULL, !more);
		bool got = false, first = false;
		uint32_t s_prof[CH], s_lo[CH], s_hi[CH];
		if (RELAY) {
			if (!more && rec_ready) {
				a = (uint64_t)ra.y << 32 | ra.x;
				o = (uint64_t)ra.w << 32 | ra.z;
				left = rb.x;
				stream = rb.y & 0x3fffffffu;
				m_front = rb.y >> 30;
				s_lo[0] = rb.z;
				s_lo[CH - 1] = CH == 2 ? rb.w : rb.z;
				got = true;
				next += stride;
				rec_ready = next < n_live;
				if (rec_ready) {
					ra = ldg_u128(&p.relay_recs[next]);
					rb = ldg_u128(reinterpret_cast<const uint4 *>(&p.relay_recs[next]) + 1);
				}
			}
			needm = 0;
		}
#pragma unroll 1
		for (int round = 0; round < 2 && needm != 0; round++) {
			if (pos >= cnt) {
				if (!rec_ready)
					break;
				/* the next record: its bitmap becomes a list of head indices */
				const uint32_t bits = lane < 16u ? nrec : 0u;
				const uint32_t c1 = __popc(bits);
				uint32_t incl = c1;
#pragma unroll
				for (int d = 1; d < 32; d <<= 1) {
					const uint32_t t = __shfl_up_sync(FULL, incl, d);
					if (lane >= (uint32_t)d)
						incl += t;
				}
				cnt = __shfl_sync(FULL, incl, 31);
				__syncwarp();		/* the old list and context have been read */
				uint32_t b = bits, k = incl - c1;
				while (b) {
					list[k++] = (uint16_t)(lane * 32u + (uint32_t)__ffs(b) - 1u);
					b &= b - 1u;
				}
				if (lane >= 16u && lane < 24u)
					ctx[lane - 16u] = nrec;
				__syncwarp();
				pos = 0;
				next += stride;
				rec_ready = next < n_live;
				nrec = rec_ready ? reinterpret_cast<const uint32_t *>(&p.live[next])[lane] : 0u;
			}
			const uint32_t rank = __popc(needm & ((1u << lane) - 1u));
			const bool mine = (needm >> lane & 1u) != 0 && rank < cnt - pos;
			if (mine) {
				const uint32_t q = list[pos + rank];
				const uint2 xa0 = *reinterpret_cast<const uint2 *>(ctx);
				const uint2 out0 = *reinterpret_cast<const uint2 *>(ctx + 2);
				const uint4 sfb = *reinterpret_cast<const uint4 *>(ctx + 4);
				const uint32_t eb = sfb.y + q;
				stream = sfb.x;
				a = ((uint64_t)xa0.y << 32 | xa0.x) + (uint64_t)(q * (uint32_t)STEP);
				o = ((uint64_t)out0.y << 32 | out0.x) + (uint64_t)(q * (uint32_t)OUT);
				left = sfb.z - 1u - eb;
				first = eb == 0;
				got = true;
			}
			const uint32_t took = __ballot_sync(FULL, mine);
			pos += __popc(took);
			needm &= ~took;
		}
		if (!RELAY && got) {
			/* the block(s) in front of the head (Walk::seed_fetch, as PTX loads) */
			if (first) {
				const StreamDev &sd = p.streams[stream];
#pragma unroll
				for (int c = 0; c < CH; c++) {
					s_prof[c] = 0;
					s_lo[c] = (uint32_t)(uint16_t)sd.prev[c][0] |
					    (uint32_t)(uint16_t)sd.prev[c][1] << 16;
					s_hi[c] = 0;
				}
			} else {
#pragma unroll
				for (int c = 0; c < CH; c++) {
					const uint64_t t = W::tail_addr(a, c);
					s_prof[c] = ldg_u8(p.src + (a - STEP + (uint64_t)(c * BS)));
					s_lo[c] = ldg_u32(p.src + (t & ~(uint64_t)3));
					s_hi[c] = ldg_u32(p.src + (t & ~(uint64_t)3) + ((t & 3u) ? 4u : 0u));
				}
			}
		}

		/* (4) ask for what the next turn reads: a new chain's first item, or the
		 * item behind the next one; never past the stream's last block */
		{
			int want = fd;
			if (got) {
				fd = -(int)((uint32_t)a & 15u);
				want = STEP + W::PEEK;
			} else if (more) {
				want = C::LOOK;
			}
			const int lim = (int)(left < 3u ? left + 1u : 4u) * STEP;
			if (want > lim)
				want = lim;
			/* whole chunks of the arena only (a < safe: an item is longer than a chunk) */
			const uint64_t room = safe - a;
			const int wantc = room < (uint64_t)want ? (int)room : want;
			/* up to KMAX chunks, each from its own address: nothing for the
			 * copies to wait for but the queue */
			const uint8_t *const gp = p.src + a + (int64_t)fd;
			const uint32_t a32 = (uint32_t)a, g32 = a32 + (uint32_t)fd;
			const int fd0 = fd;
#pragma unroll
			for (int k = 0; k < C::KMAX; k++) {
				if (fd0 + 16 * k < wantc) {
					cp_async16(ring + ((g32 + 16u * k) & (RING - 1)), gp + 16 * k);
					fd = fd0 + 16 * (k + 1);
				}
			}
			asm volatile("cp.async.commit_group;" ::: "memory");
			if (__any_sync(FULL, wantc < want)) {
				/* the arena's last, partial chunk: bytewise */
				if (wantc < want && fd < want) {
					for (int b = fd; b < want && a + (uint64_t)b < p.src_bytes; b++)
						ringp[(a32 + (uint32_t)b) & (RING - 1)] = p.src[a + (uint64_t)b];
					fd = (want + 15) & ~15;
				}
			}
		}

		/* (5) decode into this lane's row, then the warp's rows leave together */
		if (act) {
			auto out = [&](int j, const uint4 &v) {
				const uint32_t at = CH == 1 ? ((uint32_t)j ^ (lane >> 1 & 3u)) : ((uint32_t)j ^ (lane & 7u));
				*reinterpret_cast<uint4 *>(row + at * 16u) = v;
			};
			W::decode(it, p0, p1, out);
		}
		__syncwarp();
		{
			/* all the addresses, then all the rows, then the stores: three
			 * latencies instead of three per round */
			unsigned long long rd[UPR];
			uint4 v[UPR];
#pragma unroll
			for (uint32_t r = 0; r < UPR; r++)
				rd[r] = __shfl_sync(FULL, d64, r * RPI + co_r0);
#pragma unroll
			for (uint32_t r = 0; r < UPR; r++) {
				const uint32_t sl = r * RPI + co_r0;
				const uint32_t at = CH == 1 ? (co_c ^ (sl >> 1 & 3u)) : (co_c ^ (sl & 7u));
				v[r] = *reinterpret_cast<const uint4 *>(rows + (size_t)sl * OUT + at * 16u);
			}
#pragma unroll
			for (uint32_t r = 0; r < UPR; r++)
				if (rd[r] & 1ULL)
					*reinterpret_cast<uint4 *>(rd[r] - 1ULL + co_c * 16u) = v[r];
		}
		if (__any_sync(FULL, last)) {
			/* the last item of a stream: its own lane stores what the stream still
			 * owes (libbjxa.c:622-624,648) and leaves the final state */
			if (last) {
				const uint32_t valid = W::last_valid(p.streams[cur_stream]);
				uint16_t *d = reinterpret_cast<uint16_t *>(d64);
				for (uint32_t k = 0; k < valid / 2u; k++) {
					const uint32_t j = k >> 3;
					const uint32_t at = CH == 1 ? (j ^ (lane >> 1 & 3u)) : (j ^ (lane & 7u));
					d[k] = *reinterpret_cast<const uint16_t *>(row + at * 16u + (k & 7u) * 2u);
				}
				W::put_result(p, cur_stream, p0, p1);
			}
		}
		__syncwarp();

		/* (6) on to the next item, or into the chain just drawn */
		if (more) {
			a += STEP;
			o += OUT;
			left--;
			fd -= STEP;
			m = nm;
		} else if (got && RELAY) {
#pragma unroll
			for (int c = 0; c < CH; c++) {
				p0[c] = (int16_t)(uint16_t)s_lo[c];
				p1[c] = (int16_t)(uint16_t)(s_lo[c] >> 16);
			}
			m = 0;
			act = true;
			fresh = true;
		} else if (got) {
			typename W::Seed seed;
#pragma unroll
			for (int c = 0; c < CH; c++) {
				seed.prof[c] = s_prof[c];
				seed.lo[c] = s_lo[c];
				seed.hi[c] = s_hi[c];
			}
			W::seed_apply(seed, first, a, p0, p1);
			m = 0;
			act = true;
		} else {
			act = false;
		}
		if (RELAY ? !__any_sync(FULL, act || rec_ready) :
		    !__any_sync(FULL, act) && !rec_ready && pos >= cnt)
			break;
	}
}

/* ---- segment form (xa_walk.h) ---------------------------------------------------- */
/*
 * One warp per tile of the segment list (32 segments, streams in arena order: xa_plan.h
 * emit_seg_tiles), a segment of kSegItems items per lane, every item decoded with
 * the chain step (walk_seg_tile_serial in xa_walk.h is the same thing in plain
 * loops).  The data path is the dense walkers': a lane reads
 * its items straight from the arena, 16-byte cp.async chunks into a private ring in
 * shared memory, one item ahead of the decode; the PCM goes to a row of shared
 * memory and leaves the warp eight (stereo: four) rows per store instruction.
 * What the dense walkers spend on drawing chains -- lists, ballots, seeds, a peek at
 * the next profile byte every turn -- is gone: a lane knows its range from the
 * start, and the turns of a warp differ only at the two ends of the segments.
 */
#ifndef XA_SEG_CTAS
#define XA_SEG_CTAS 0
#endif
#ifndef XA_SEG_WARPS
#define XA_SEG_WARPS 0
#endif
#ifndef XA_SEG_RANGED
#define XA_SEG_RANGED 1		/* the step with the bias on the ranged code (xa_core.h: sample_chain_r) */
#endif
#ifndef XA_SEG_AHEAD
#define XA_SEG_AHEAD 1		/* items a lane's copies run ahead of its decode, at least */
#endif
template <int BITS, int CH>
struct SegCfg {
	typedef Walk<BITS, CH> W;
	/* the ring holds the item being decoded and the D behind it, from any alignment:
	 * the smallest power of two that gives D >= XA_SEG_AHEAD.  One item ahead (the
	 * dense walkers' depth) is a turn of one warp, which hides the arena's latency
	 * only while other warps have turns to run -- not on data without cut blocks,
	 * where a few warps per SM do all the work */
	static constexpr int ahead_of(int ring) { return (ring - 30) / W::STEP - 1; }
	static constexpr int RING = ahead_of(64) >= XA_SEG_AHEAD ? 64 : ahead_of(128) >= XA_SEG_AHEAD ? 128 : 256;
	static constexpr int D = XA_SEG_AHEAD;
	static constexpr int SLOT = RING + 16;	/* lane stride: spreads equal offsets over the banks */
	/* the big rings leave the L1 some 20 KB: their copies go past it (cp_async16_cg) */
	static constexpr bool kPastL1 = RING > 128 || CH == 2;	/* 4-bit stereo: the same either way alone,
							 * but it often shares an SM with a big-ring class */
	/* chunks a lane asks for in one turn at most: one item's worth */
	static constexpr int KMAX = (15 + W::STEP + 15) / 16;
	/* mono: 2 CTAs of 10 warps (96 registers a thread) measured 3-6 % faster than 3 of 8
	 * (80 registers); stereo: 2 of 8 (128 registers, two chains side by side) */
	static constexpr int kWarps = XA_SEG_WARPS != 0 ? XA_SEG_WARPS : CH == 1 ? 10 : 8;
	static constexpr int kThreads = kWarps * 32;
	static constexpr size_t kSmem = (size_t)kThreads * (SLOT + W::OUT);
	/* CTAs per SM: what the shared memory allows, but no more than leaves a thread its
	 * registers -- 80 for mono, 128 for stereo (two chains side by side).  Measured
	 * (profiles/history_r2.md): deeper rings (more items in flight per lane) at the
	 * price of fewer warps lose on every mix */
	static constexpr int kFit = (int)((227 * 1024) / (kSmem + 1024));
	static constexpr int kWant = 2;
	static constexpr int kCtas = XA_SEG_CTAS != 0 ? XA_SEG_CTAS : kFit > kWant ? kWant : kFit;
	static_assert(D >= 1 && 30 + (D + 1) * W::STEP <= RING, "ring");
};

/*
 * The state a lane could not recompute: from the mailbox of the lane that decodes
 * the segment in front (slot - 1).  While even the segment before that one is not
 * through (slot - 2), the wait is a whole segment away and the lane looks rarely:
 * on data without cut blocks all but a few warps of the grid are waiting here, and
 * their looks would otherwise load the L2 more than the decode does.
 */
__device__ __forceinline__ void
seg_wait_front(const DecodeParams &p, uint32_t slot, bool far, uint32_t ch, int &p0, int &p1)
{
	unsigned long long v;
	if (far) {
		unsigned long long t0 = 0;
		for (uint32_t spins = 0;; spins++) {
			if (mailbox_try(&p.carry[(uint64_t)(slot - 2u) * 2 + ch], p.epoch, v) ||
			    mailbox_try(&p.carry[(uint64_t)(slot - 1u) * 2 + ch], p.epoch, v))
				break;
			__nanosleep(8000);
			if ((spins & 255u) == 255u) {
				unsigned long long now;
				asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
				if (t0 == 0)
					t0 = now;
				else if (now - t0 > p.carry_timeout_ns)
					break;		/* mailbox_get below flags the launch */
			}
		}
	}
	v = mailbox_get(&p.carry[(uint64_t)(slot - 1u) * 2 + ch], p.epoch, p.fault, p.carry_timeout_ns);
	p0 = (int16_t)(uint16_t)v;
	p1 = (int16_t)(uint16_t)(v >> 16);
}

template <int BITS, int CH>
__global__ void __launch_bounds__(SegCfg<BITS, CH>::kThreads, SegCfg<BITS, CH>::kCtas)
xa_seg_kernel(const DecodeParams p)
{
	typedef Walk<BITS, CH> W;
	typedef SegCfg<BITS, CH> C;
	constexpr int BS = W::BS, STEP = W::STEP, OUT = W::OUT, RING = C::RING;
	constexpr uint32_t FULL = 0xffffffffu;
	constexpr uint32_t UPR = W::UNITS;		/* 16-byte units per row of PCM */
	constexpr uint32_t RPI = 32u / UPR;		/* rows per store instruction */

	if (p.choice != NULL && *p.choice != p.want)
		return;		/* the census picked another form */

	extern __shared__ __align__(16) unsigned char smem_raw[];
	uint32_t tid, lane;
	asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
	asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
	const uint32_t warp = tid >> 5;
	uint8_t *const ringp = smem_raw + (size_t)tid * C::SLOT;
	const uint32_t ring = smem_u32(ringp);
	uint8_t *const rows = smem_raw + (size_t)C::kThreads * C::SLOT + (size_t)warp * 32 * OUT;
	uint8_t *const row = rows + (size_t)lane * OUT;
	const uint32_t co_c = lane % UPR, co_r0 = lane / UPR;
	/* whole 16-byte chunks of the arena; what lies behind is fetched bytewise */
	const uint64_t safe = p.src_bytes & ~(uint64_t)15;

	for (;;) {
		/* ---- the warp's next tile ---- */
		uint32_t ticket = 0;
		if (lane == 0)
			ticket = (uint32_t)(atomicAdd(p.ticket, 1ULL) + 1ULL);
		ticket = __shfl_sync(FULL, ticket, 0);
		if (ticket >= p.n_tiles)
			break;
		const uint4 te = ldg_u128(&p.tiles[ticket]);	/* first stream, its segment, lanes, streams */
		/* lane L: the L-th segment from segment te.y of stream order[te.x] on (xa_walk.h:
		 * seg_lane); lane k looks up how many segments the tile's k-th stream has */
		bool valid = lane < te.z;
		uint32_t stream = 0, seg = 0, n = 0, n0 = 0, slot = 0;
		{
			uint32_t mine = lane < te.w ? p.order[te.x + lane] : 0u;
			uint32_t ns = lane < te.w ? (p.streams[mine].blocks + kSegItems - 1u) / kSegItems : 0u;
			if (lane == 0)
				ns -= te.y;		/* the first stream's segments in front of the tile */
			uint32_t incl = ns;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t t = __shfl_up_sync(FULL, incl, d);
				if (lane >= (uint32_t)d)
					incl += t;
			}
			/* streams whose segments all lie in front of this lane's */
			uint32_t k = 0;
#pragma unroll 4
			for (uint32_t q = 0; q < 32u; q++) {
				if (q >= te.w)
					break;
				k += __shfl_sync(FULL, incl, q) <= lane;
			}
			k = k < te.w ? k : te.w - 1u;
			stream = __shfl_sync(FULL, mine, k);
			const uint32_t excl = __shfl_sync(FULL, incl - ns, k);
			seg = lane - excl + (k == 0 ? te.y : 0u);
			n0 = seg * kSegItems;
		}
		uint64_t a0 = 0, o0 = 0;
		bool ends = false;		/* the segment is its stream's last */
		bool pending = false;		/* the state comes from lane - 1, after a pass */
		int p0[CH], p1[CH];
		uint32_t back = 0;
#pragma unroll
		for (int c = 0; c < CH; c++)
			p0[c] = p1[c] = 0;
		if (valid) {
			const StreamDev &sd = p.streams[stream];
			const uint32_t blocks = sd.blocks;
			n = blocks - n0 < kSegItems ? blocks - n0 : kSegItems;
			ends = n0 + n == blocks;
			a0 = sd.xa_off + (uint64_t)n0 * STEP;
			o0 = sd.pcm_off + (uint64_t)n0 * OUT;
			slot = sd.slot_base + seg;

			/* ---- the state in front of the segment (seg_front) ---- */
			bool own = n0 == 0, mail = false;
			if (!own) {
				const uint32_t lim = n0 < kSegBack ? n0 : kSegBack;
				uint32_t found[CH], missing = CH;
#pragma unroll
				for (int c = 0; c < CH; c++)
					found[c] = 0;
#pragma unroll 1
				for (uint32_t base = 0; base < lim && missing != 0; base += 8u) {
					/* eight items' profile bytes at a time, all loads in flight together */
					uint32_t pr[8][CH];
#pragma unroll
					for (uint32_t k = 0; k < 8u; k++)
#pragma unroll
						for (int c = 0; c < CH; c++)
							pr[k][c] = base + k < lim ? ldg_u8(p.src + (a0 -
							    (uint64_t)(base + k + 1u) * STEP + (uint64_t)(c * BS))) : 0x10u;
#pragma unroll
					for (uint32_t k = 0; k < 8u; k++)
#pragma unroll
						for (int c = 0; c < CH; c++)
							if (found[c] == 0 && block_kind(pr[k][c]) != kChain) {
								found[c] = base + k + 1u;
								missing--;
							}
				}
				if (missing == 0) {
#pragma unroll
					for (int c = 0; c < CH; c++)
						back = found[c] > back ? found[c] : back;
				} else if (lim == n0) {
					back = n0;
					own = true;
				} else {
					mail = true;
				}
			}
			if (own) {
#pragma unroll
				for (int c = 0; c < CH; c++) {
					p0[c] = sd.prev[c][0];
					p1[c] = sd.prev[c][1];
				}
			}
			if (mail && lane == 0) {
				/* the tile before holds a lower ticket: running or done */
#pragma unroll
				for (int c = 0; c < CH; c++)
					seg_wait_front(p, slot, seg >= 2u, (uint32_t)c, p0[c], p1[c]);
			} else if (mail) {
				pending = true;		/* lane - 1 decodes the segment in front */
			}
		}
		__syncwarp();
		bool ready = valid && !pending;

		/* ---- passes: one, unless lanes have to wait for their neighbours ---- */
		for (;;) {
		const bool run = ready;

		/* ---- the turns, the warp in lockstep: turn t is item t of every lane's segment,
		 * t < 0 the items in front that a lane decodes only for their state ---- */
		const int maxback = (int)__reduce_max_sync(FULL, run ? back : 0u);
		const int nmax = (int)__reduce_max_sync(FULL, run ? n : 0u);
		const int first = run ? -(int)back : 0;		/* this lane's first turn */
		const int nn = run ? (int)n : 0;			/* 0 on a lane that sits this pass out */
		/* the ring, in bytes from a0: filled up to fe; chunks lie at arena offsets that
		 * are multiples of 16, and at the same offsets modulo RING in the ring */
		/* ... counted from a0 rounded down to a chunk, NOT from the arena: lanes in
		 * lockstep then keep their items at (nearly) the same places of their rings,
		 * and the eight lanes of a quarter warp -- SLOT apart, 16 bytes past a multiple
		 * of 128 -- touch eight different bank groups per 16-byte copy or store instead
		 * of random ones (ncu: 19.6 wavefronts per LDGSTS before, profiles/history_r2.md) */
		const uint32_t a32 = (uint32_t)a0 & 15u;
		int fe = first * STEP;
		fe -= (int)((a32 + (uint32_t)fe) & 15u);
		const uint8_t *gq = p.src + a0 + (int64_t)fe;	/* the next chunk to ask for */
		/* whole chunks end here; a lane whose segment reaches further gets the arena's
		 * last few bytes one by one (tail) */
		const uint64_t room = safe > a0 ? safe - a0 : 0;
		const int fe_safe = run ? (room > 0x40000000ULL ? 0x40000000 : (int)room) : 0;
		const bool tail = __any_sync(FULL, run && nn * STEP > fe_safe);
		/* one group of copies: this lane's item u */
		auto request = [&](int u) {
			if (u >= first && u < nn) {
				const int want = (u + 1) * STEP < fe_safe ? (u + 1) * STEP : fe_safe;
				const uint32_t r32 = a32 + (uint32_t)fe;
				int k = 0;
#pragma unroll
				for (int i = 0; i < C::KMAX; i++)
					if (fe + 16 * i < want) {
						if (C::kPastL1)
							cp_async16_cg(ring + ((r32 + 16u * i) & (RING - 1)), gq + 16 * i);
						else
							cp_async16(ring + ((r32 + 16u * i) & (RING - 1)), gq + 16 * i);
						k = i + 1;
					}
				fe += 16 * k;
				gq += 16 * k;
			}
			asm volatile("cp.async.commit_group;" ::: "memory");
			if (tail && u >= first && u < nn && (u + 1) * STEP > fe_safe) {
				const int from = u * STEP > fe_safe ? u * STEP : fe_safe < first * STEP ? first * STEP : fe_safe;
				for (int b = from; b < (u + 1) * STEP; b++)
					if (a0 + (uint64_t)(int64_t)b < p.src_bytes)
						ringp[(a32 + (uint32_t)b) & (RING - 1)] = p.src[a0 + (uint64_t)(int64_t)b];
			}
		};
#pragma unroll
		for (int d = 0; d < C::D; d++)
			request(-maxback + d);

		/* the rows this lane copies out: row r * RPI + co_r0 of every turn; where that
		 * lane's PCM of turn 0 goes, and how many whole rows it stores */
		unsigned long long ob[UPR];
		{
			const unsigned long long mine = (unsigned long long)(p.dst + o0);
#pragma unroll
			for (uint32_t r = 0; r < UPR; r++)
				ob[r] = __shfl_sync(FULL, mine, r * RPI + co_r0) + co_c * 16u;
		}
		/* a stream's last item goes by itself (it may owe less than a row) */
		const int full = run && ends ? nn - 1 : nn;
		const int last_t = run && ends ? nn - 1 : -1;
		uint32_t at32 = a32 + (uint32_t)(-maxback * STEP);	/* low address bits of turn t's item */

#ifdef XA_SEG_PROF
		long long pc[6] = { 0, 0, 0, 0, 0, 0 }, pt = clock64();
#define XA_SEG_TICK(i) do { long long now_ = clock64(); pc[i] += now_ - pt; pt = now_; } while (0)
#else
#define XA_SEG_TICK(i) do { } while (0)
#endif
#pragma unroll 1
		for (int t = -maxback; t < nmax; t++) {
			XA_SEG_TICK(5);
			/* (1) the copies of this turn's item have landed: all groups but the
			 * C::D - 1 asked for last */
			asm volatile("cp.async.wait_group %0;" :: "n"(C::D - 1) : "memory");
			const bool act = t >= first && t < nn;
			XA_SEG_TICK(0);

			/* (2) this turn's item out of the ring */
			typename W::Item it;
			if (act) {
#pragma unroll
				for (int c = 0; c < CH; c++) {
					const uint32_t at = at32 + (uint32_t)(c * BS);
					it.prof[c] = ringp[at & (RING - 1)];
					/*
					 * The payload out of the ring as whole 16-byte chunks: 32 lanes, each
					 * with a 16-byte aligned ring of its own, read 4-byte words 4-way bank
					 * conflicted whatever the lane stride, but 16-byte words at full rate
					 * (a quarter warp covers all 32 banks).  So: the NW chunks from the
					 * one that holds payload byte 0, then the word the payload starts in
					 * selected in two steps (by 2 words, by 1), then the byte shift.
					 */
					constexpr int NW = (15 + 4 * BITS + 15) / 16;
					const uint32_t pay = at + 1u, cb = pay & ~15u, sh = (pay & 3u) * 8u;
					uint32_t w[4 * NW + 2];
#pragma unroll
					for (int k = 0; k < NW; k++) {
						const uint4 q = *reinterpret_cast<const uint4 *>(
						    ringp + ((cb + 16u * k) & (RING - 1)));
						w[4 * k] = q.x; w[4 * k + 1] = q.y; w[4 * k + 2] = q.z; w[4 * k + 3] = q.w;
					}
					w[4 * NW] = w[4 * NW + 1] = 0u;
					const bool by2 = (pay & 8u) != 0, by1 = (pay & 4u) != 0;
					uint32_t v[BITS + 2], u[BITS + 1];
#pragma unroll
					for (int k = 0; k < BITS + 2; k++)
						v[k] = by2 ? w[k + 2] : w[k];
#pragma unroll
					for (int k = 0; k < BITS + 1; k++)
						u[k] = by1 ? v[k + 1] : v[k];
#pragma unroll
					for (int k = 0; k < BITS; k++)
						it.pw[c][k] = __funnelshift_r(u[k], u[k + 1], sh);
				}
				if (t >= 0) {
#pragma unroll
					for (int c = 0; c < CH; c++)
						if (it.prof[c] >> 4 >= 5u)
							global_min_u32(&p.first_bad[stream], (n0 + (uint32_t)t) * CH + c);
				}
			}

			XA_SEG_TICK(1);
			/* (3) ask for the item C::D turns on */
			request(t + C::D);
			XA_SEG_TICK(2);

			/* (4) decode into this lane's row, then the warp's rows leave together */
			if (act) {
				auto out = [&](int j, const uint4 &v) {
					const uint32_t at = CH == 1 ? ((uint32_t)j ^ (lane >> 1 & 3u)) : ((uint32_t)j ^ (lane & 7u));
					*reinterpret_cast<uint4 *>(row + at * 16u) = v;
				};
				W::template decode<XA_SEG_RANGED != 0>(it, p0, p1, out);
			}
			XA_SEG_TICK(3);
			if (t >= 0) {
				__syncwarp();
				uint4 v[UPR];
#pragma unroll
				for (uint32_t r = 0; r < UPR; r++) {
					const uint32_t sl = r * RPI + co_r0;
					const uint32_t at = CH == 1 ? (co_c ^ (sl >> 1 & 3u)) : (co_c ^ (sl & 7u));
					v[r] = *reinterpret_cast<const uint4 *>(rows + (size_t)sl * OUT + at * 16u);
				}
				/* bit r * RPI: the lane whose row r this one copies has a whole row */
				const uint32_t rows_on = __ballot_sync(FULL, t < full) >> co_r0;
#pragma unroll
				for (uint32_t r = 0; r < UPR; r++) {
					if (rows_on & (1u << (r * RPI)))
						*reinterpret_cast<uint4 *>(ob[r]) = v[r];
					ob[r] += OUT;
				}
				if (__any_sync(FULL, t == last_t)) {
					/* the last item of a stream: its own lane stores what the stream still
					 * owes (libbjxa.c:622-624,648) and leaves the final state */
					if (t == last_t) {
						const uint32_t owed = W::last_valid(p.streams[stream]);
						uint16_t *d = reinterpret_cast<uint16_t *>(p.dst + o0 + (uint64_t)t * OUT);
						for (uint32_t k = 0; k < owed / 2u; k++) {
							const uint32_t j = k >> 3;
							const uint32_t at = CH == 1 ? (j ^ (lane >> 1 & 3u)) : (j ^ (lane & 7u));
							d[k] = *reinterpret_cast<const uint16_t *>(row + at * 16u + (k & 7u) * 2u);
						}
						W::put_result(p, stream, p0, p1);
					}
				}
				__syncwarp();
			}
			at32 += (uint32_t)STEP;
			XA_SEG_TICK(4);
		}
#ifdef XA_SEG_PROF
		if (lane == 0 && (ticket & 1023u) == 5u)
			printf("seg prof ticket %u turns %d: wait %lld read %lld request %lld decode %lld out %lld loop %lld\n",
			    ticket, nmax + maxback, pc[0], pc[1], pc[2], pc[3], pc[4], pc[5]);
#endif
		if (run && !ends && lane == 31u) {
			/* the state behind the tile, for the next tile's first lane */
#pragma unroll
			for (int c = 0; c < CH; c++)
				mailbox_put(&p.carry[(uint64_t)slot * 2 + c],
				    ((unsigned long long)p.epoch << 32) |
				    ((unsigned long long)(uint16_t)p1[c] << 16) | (uint16_t)p0[c]);
		}
		/* lanes that waited for the lane in front take its state and go next */
		if (!__any_sync(FULL, pending))
			break;
		{
			const bool front_ran = __shfl_up_sync(FULL, (int)run, 1) != 0 && lane != 0;
			int q0[CH], q1[CH];
#pragma unroll
			for (int c = 0; c < CH; c++) {
				q0[c] = __shfl_up_sync(FULL, p0[c], 1);
				q1[c] = __shfl_up_sync(FULL, p1[c], 1);
			}
			ready = pending && front_ran;
			if (ready) {
#pragma unroll
				for (int c = 0; c < CH; c++) {
					p0[c] = q0[c];
					p1[c] = q1[c];
				}
				back = 0;
				pending = false;
			}
		}
		if (!__any_sync(FULL, ready))
			break;		/* (cannot happen: a pending lane's front either ran or is pending) */
		}
	}
}

/* ---- chain form: mono data without cut blocks --------------------------------------- */
/*
 * A stream without cut blocks is one dependent chain from its first sample to its
 * last: nothing but more streams makes such data faster, and what a stream costs is
 * the latency of the step -- 31 cycles a sample for a warp that does nothing else
 * (tools/step_bench.cu).  In the tile forms and the segment form a lone warp's turn
 * takes twice that: it also feeds itself (ring, payload, ranged codes) and takes
 * its samples away (rows, stores) in between.  Here a CTA owns 32 streams from
 * their first block to their last, a stream per lane, and three warps share the
 * work:
 *   loader   keeps the lanes' rings filled (the segment form's data path) and puts
 *            every block's payload, word-aligned, and profile byte into shared
 *            memory [word][lane];
 *   stepper  does nothing but the block step (xa_core.h: decode_block_chain -- the
 *            ranged codes and the packing ride in the gaps of the dependent chain),
 *            BITS + 1 loads and 16 stores a block beside it, the state in registers
 *            for the whole stream -- no hand-overs, no carries;
 *   storer   takes the packed samples, a 64-byte row per lane, and stores them 16
 *            bytes a lane.
 * Stages of kChainK blocks go round between them on mbarriers.  The census sends
 * mono classes here from 98.5 % chain blocks (where the wide tiles used to go).
 */
#ifndef XA_CHAIN_K
#define XA_CHAIN_K 8
#endif
#ifndef XA_CHAIN_UNROLL
#define XA_CHAIN_UNROLL 8
#endif
constexpr int kChainUnroll = XA_CHAIN_UNROLL;
#ifndef XA_CHAIN_LOADER_UNROLL
#define XA_CHAIN_LOADER_UNROLL 1
#endif
constexpr int kChainLoaderUnroll = XA_CHAIN_LOADER_UNROLL;
constexpr int kChainK = XA_CHAIN_K, kChainS = 2, kChainThreads = 96, kChainRing = 256;

/* blocks (stereo: pairs) per stage: what lets three CTAs share an SM's shared memory for
 * mono; stereo rows are twice as long */
template <int CH> struct ChainK { static constexpr int K = CH == 1 ? kChainK : kChainK / 2; };

template <int BITS, int CH>
struct ChainSmem {
	static constexpr int K = ChainK<CH>::K;
	static constexpr int ROW = 20;			/* words: rows 80 bytes apart, so that eight lanes'
							 * 16-byte accesses hit all 32 banks */
	uint32_t pw[kChainS][K][CH][BITS + 4][32];	/* per channel: payload words, the range shift,
							 * k0, k1 and c (xa_core.h: chain_bias_c), [word][lane] */
	__align__(16) uint32_t out[kChainS][K][CH][32][ROW];	/* per channel: a row of 16 packed pairs
							 * of its samples per lane */
	__align__(16) unsigned char ring[CH][32][kChainRing + 16];	/* a loader warp per channel, each with
							 * rings of its own */
	unsigned long long bar[4][kChainS];		/* xs full / empty, out full / empty */
	uint32_t blocks[32];
	uint32_t maxblocks;
};

template <int BITS, int CH>
__global__ void __launch_bounds__(kChainThreads + 64 * (CH - 1))
xa_chain_kernel(const DecodeParams p, const uint32_t *order, uint32_t n_streams)
{
	typedef Walk<BITS, CH> W;
	typedef SegCfg<BITS, CH> C;
	/* the loader alone hides the arena's latency here: up to four items ahead */
	constexpr int BS = W::BS, STEP = W::STEP, OUT = W::OUT, RING = kChainRing, K = ChainK<CH>::K, S = kChainS;
	constexpr int D = (RING - 30) / STEP - 1 > 4 ? 4 : (RING - 30) / STEP - 1;
	static_assert(D >= 1, "ring");
	constexpr uint32_t FULL = 0xffffffffu;
	enum { kXsFull, kXsEmpty, kOutFull, kOutEmpty };

	if (p.choice != NULL && *p.choice != p.want)
		return;		/* the census picked another form */

	extern __shared__ __align__(16) unsigned char smem_raw[];
	ChainSmem<BITS, CH> &sm = *reinterpret_cast<ChainSmem<BITS, CH> *>(smem_raw);
	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	const uint32_t first = blockIdx.x * 32u;
	const bool have = first + lane < n_streams;
	const uint32_t stream = have ? order[first + lane] : 0u;
	const uint32_t nblk = have ? p.streams[stream].blocks : 0u;
	if (warp == 0) {
		sm.blocks[lane] = nblk;
		const uint32_t mx = __reduce_max_sync(FULL, nblk);
		if (lane == 0)
			sm.maxblocks = mx;
	}
	if (tid == 0)
		for (int b = 0; b < 4; b++)
			for (int s = 0; s < S; s++)	/* a loader and a stepper per channel; one storer */
				mbar_init(smem_u32(&sm.bar[b][s]), b == kOutEmpty ? 1 : CH);
	__syncthreads();
	const uint32_t nst = (sm.maxblocks + K - 1) / K;
#ifdef XA_CHAIN_PROF
	long long wt = 0, t_begin = clock64(), g_begin;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_begin));
#define XA_CHAIN_WAIT(b, ph) do { long long w0_ = clock64(); mbar_wait(b, ph); wt += clock64() - w0_; } while (0)
#else
#define XA_CHAIN_WAIT(b, ph) mbar_wait_idle(b, ph, 0)
#endif
	/* loader and storer have time to spare: they sleep between looks */
#define XA_CHAIN_IDLE(b, ph) mbar_wait_idle(b, ph, 64)

	if (warp == 0 || warp == 4) {
		/* ---- loader: warp 0 the left (or only) channel, warp 4 the right -- each reads the
		 * whole stream through rings of its own and aligns its channel's words ---- */
		const int lc = warp == 0 ? 0 : CH - 1;
		const uint64_t a0 = have ? p.streams[stream].xa_off : 0;
		uint8_t *const ringp = sm.ring[lc][lane];
		const uint32_t ring = smem_u32(ringp);
		const uint64_t safe = p.src_bytes & ~(uint64_t)15;
		const uint32_t a32 = (uint32_t)a0 & 15u;
		const int nn = (int)nblk;
		int fe = -(int)a32;
		const uint8_t *gq = p.src + a0 + (int64_t)fe;
		const uint64_t room = safe > a0 ? safe - a0 : 0;
		const int fe_safe = have ? (room > 0x40000000ULL ? 0x40000000 : (int)room) : 0;
		const bool tail = __any_sync(FULL, have && nn * STEP > fe_safe);
		auto request = [&](int u) {
			if (u < nn) {
				const int want = (u + 1) * STEP < fe_safe ? (u + 1) * STEP : fe_safe;
				const uint32_t r32 = a32 + (uint32_t)fe;
				int k = 0;
#pragma unroll
				for (int i = 0; i < C::KMAX; i++)
					if (fe + 16 * i < want) {
						cp_async16(ring + ((r32 + 16u * i) & (RING - 1)), gq + 16 * i);
						k = i + 1;
					}
				fe += 16 * k;
				gq += 16 * k;
			}
			asm volatile("cp.async.commit_group;" ::: "memory");
			if (tail && u < nn && (u + 1) * STEP > fe_safe) {
				const int from = u * STEP > fe_safe ? u * STEP : fe_safe < 0 ? 0 : fe_safe;
				for (int b = from; b < (u + 1) * STEP; b++)
					if (a0 + (uint64_t)(int64_t)b < p.src_bytes)
						ringp[(a32 + (uint32_t)b) & (RING - 1)] = p.src[a0 + (uint64_t)(int64_t)b];
			}
		};
#pragma unroll
		for (int d = 0; d < D; d++)
			request(d);
		uint32_t at0 = a32;
#pragma unroll 1
		for (uint32_t st = 0; st < nst; st++) {
			const uint32_t s = st % S;
			if (st >= (uint32_t)S)
				XA_CHAIN_IDLE(smem_u32(&sm.bar[kXsEmpty][s]), (st / S - 1u) & 1u);
#pragma unroll kChainLoaderUnroll
			for (int k = 0; k < K; k++) {
				const int t = (int)st * K + k;
				asm volatile("cp.async.wait_group %0;" :: "n"(D - 1) : "memory");
				uint32_t prof = 0, pw[BITS];
				if (t < nn) {
					const uint32_t at = at0 + (uint32_t)(lc * BS);
					prof = ringp[at & (RING - 1)];
					constexpr int NW = (15 + 4 * BITS + 15) / 16;
					const uint32_t pay = at + 1u, cb = pay & ~15u, sh = (pay & 3u) * 8u;
					uint32_t w[4 * NW];
#pragma unroll
					for (int i = 0; i < NW; i++) {
						const uint4 q = *reinterpret_cast<const uint4 *>(
						    ringp + ((cb + 16u * i) & (RING - 1)));
						w[4 * i] = q.x; w[4 * i + 1] = q.y; w[4 * i + 2] = q.z; w[4 * i + 3] = q.w;
					}
					const bool by2 = (pay & 8u) != 0, by1 = (pay & 4u) != 0;
					uint32_t v[BITS + 2], u[BITS + 1];
#pragma unroll
					for (int i = 0; i < BITS + 2; i++)
						v[i] = by2 ? w[i + 2] : w[i];
#pragma unroll
					for (int i = 0; i < BITS + 1; i++)
						u[i] = by1 ? v[i + 1] : v[i];
#pragma unroll
					for (int i = 0; i < BITS; i++)
						pw[i] = __funnelshift_r(u[i], u[i + 1], sh);
				}
				request(t + D);
				if (t < nn) {
					if (prof >> 4 >= 5u)
						global_min_u32(&p.first_bad[stream], (uint32_t)t * CH + lc);
#pragma unroll
					for (int i = 0; i < BITS; i++)
						sm.pw[s][k][lc][i][lane] = pw[i];
				}
				{
					/* a lane whose stream is through: filter 0, whatever the payload */
					const int k0 = gain_k0(prof >> 4), k1 = gain_k1(prof >> 4);
					sm.pw[s][k][lc][BITS][lane] = 16u + (prof & 15u);
					sm.pw[s][k][lc][BITS + 1][lane] = (uint32_t)k0;
					sm.pw[s][k][lc][BITS + 2][lane] = (uint32_t)k1;
					sm.pw[s][k][lc][BITS + 3][lane] = (uint32_t)chain_bias_c(k0, k1);
				}
				at0 += (uint32_t)STEP;
			}
			__syncwarp();
			if (lane == 0)
				mbar_arrive(smem_u32(&sm.bar[kXsFull][s]));
		}
	} else if (warp != 2) {
		/* ---- stepper: warp 1 the left (or only) channel, warp 3 the right ---- */
		/*
		 * One block a turn: its words and constants out of shared memory, then nothing
		 * but the step -- ranged codes and packing ride in the gaps of the dependent
		 * chain; the state stays biased from block to block (xa_core.h:
		 * sample_chain_r).  Stereo: a warp per channel -- both chains of a pair side
		 * by side in one warp took 63 cycles a frame, a lone warp issuing one
		 * instruction every three cycles (profiles/history_r2.md).  (Fetching block
		 * t + 1's words while block t steps, or ranging it in the same turn, measured
		 * 3-12 % slower: ptxas keeps neither early.)
		 */
		const int c = warp == 1 ? 0 : CH - 1;
		int b0 = 32768, b1 = 32768;	/* the state, n-1 and n-2, biased */
		if (have) {
			b0 += p.streams[stream].prev[c][0];
			b1 += p.streams[stream].prev[c][1];
		}
#pragma unroll 1
		for (uint32_t st = 0; st < nst; st++) {
			const uint32_t s = st % S;
			XA_CHAIN_WAIT(smem_u32(&sm.bar[kXsFull][s]), (st / S) & 1u);
			if (st >= (uint32_t)S)
				XA_CHAIN_WAIT(smem_u32(&sm.bar[kOutEmpty][s]), (st / S - 1u) & 1u);
#pragma unroll (CH == 1 ? kChainUnroll : kChainUnroll / 2)
			for (int k = 0; k < K; k++) {
				uint32_t pw[BITS], o[16];
#pragma unroll
				for (int i = 0; i < BITS; i++)
					pw[i] = sm.pw[s][k][c][i][lane];
				const int sh = (int)sm.pw[s][k][c][BITS][lane];
				const int k0 = (int)sm.pw[s][k][c][BITS + 1][lane], k1 = (int)sm.pw[s][k][c][BITS + 2][lane];
				const int cc = (int)sm.pw[s][k][c][BITS + 3][lane];
#pragma unroll
				for (int i = 0; i < 16; i++) {
					const int x0 = sample_chain_r(top_code<BITS>(pw, 2 * i), sh, k0, k1, cc, b0, b1);
					const int x1 = sample_chain_r(top_code<BITS>(pw, 2 * i + 1), sh, k0, k1, cc, b0, b1);
					o[i] = pack2_biased(x0, x1);
				}
#pragma unroll
				for (int i = 0; i < 4; i++)
					*reinterpret_cast<uint4 *>(&sm.out[s][k][c][lane][4 * i]) =
					    make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
				if (st * K + k + 1u == nblk) {
					p.results[stream].prev[c][0] = (int16_t)(b0 - 32768);
					p.results[stream].prev[c][1] = (int16_t)(b1 - 32768);
				}
			}
			__syncwarp();
			if (lane == 0) {
				mbar_arrive(smem_u32(&sm.bar[kXsEmpty][s]));
				mbar_arrive(smem_u32(&sm.bar[kOutFull][s]));
			}
		}
	} else {
		/* ---- storer: lane = 16-byte units `unit` (+ 4 for stereo) of the rows of streams
		 * 8q + lane % 8 (a quarter warp reads one unit of eight rows: all 32 banks) ---- */
		const uint32_t unit = lane >> 3, r0 = lane & 7u;
		uint32_t nb[4], owed[4];
		uint8_t *base[4];
#pragma unroll
		for (int q = 0; q < 4; q++) {
			const uint32_t r = 8u * q + r0;
			nb[q] = sm.blocks[r];
			owed[q] = 0;
			base[q] = NULL;
			if (nb[q] != 0) {
				const StreamDev &sd = p.streams[order[first + r]];
				base[q] = p.dst + sd.pcm_off + unit * 16u;
				owed[q] = W::last_valid(sd);
			}
		}
#pragma unroll 1
		for (uint32_t st = 0; st < nst; st++) {
			const uint32_t s = st % S;
			XA_CHAIN_IDLE(smem_u32(&sm.bar[kOutFull][s]), (st / S) & 1u);
#pragma unroll 1
			for (int k = 0; k < K; k++) {
				const uint32_t t = st * K + k;
#pragma unroll
				for (int uh = 0; uh < CH; uh++) {
					const uint32_t un = unit + 4u * uh;	/* this lane's unit of the row */
					uint4 v[4];
#pragma unroll
					for (int q = 0; q < 4; q++) {
						if (CH == 1) {
							v[q] = *reinterpret_cast<const uint4 *>(&sm.out[s][k][0][8 * q + r0][4 * un]);
						} else {
							/* four frames = two packed pairs of either channel, interleaved */
							const uint2 l = *reinterpret_cast<const uint2 *>(&sm.out[s][k][0][8 * q + r0][2 * un]);
							const uint2 r = *reinterpret_cast<const uint2 *>(&sm.out[s][k][CH - 1][8 * q + r0][2 * un]);
							v[q].x = __byte_perm(l.x, r.x, 0x5410);
							v[q].y = __byte_perm(l.x, r.x, 0x7632);
							v[q].z = __byte_perm(l.y, r.y, 0x5410);
							v[q].w = __byte_perm(l.y, r.y, 0x7632);
						}
					}
#pragma unroll
					for (int q = 0; q < 4; q++) {
						if (t < nb[q]) {
							uint8_t *d = base[q] + (uint64_t)t * OUT + 64u * uh;
							if (t + 1u < nb[q] || un * 16u + 16u <= owed[q]) {
								*reinterpret_cast<uint4 *>(d) = v[q];
							} else if (un * 16u < owed[q]) {
								/* the stream's last block owes less than this unit */
								const uint32_t n16 = (owed[q] - un * 16u) / 2u;
#pragma unroll
								for (uint32_t h = 0; h < 8u; h++) {
									const uint32_t w4 = h < 2 ? v[q].x : h < 4 ? v[q].y : h < 6 ? v[q].z : v[q].w;
									if (h < n16)
										reinterpret_cast<uint16_t *>(d)[h] = (uint16_t)(w4 >> (16u * (h & 1u)));
								}
							}
						}
					}
				}
			}
			__syncwarp();
			if (lane == 0)
				mbar_arrive(smem_u32(&sm.bar[kOutEmpty][s]));
		}
	}
#ifdef XA_CHAIN_PROF
	if (lane == 0 && blockIdx.x == 1) {
		long long g_end;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_end));
		printf("chain prof warp %u: %lld cycles in %lld ns, waited %lld, %u stages\n", warp, clock64() - t_begin,
		    g_end - g_begin, wt, nst);
	}
#endif
}

template <int BITS, int CH>
__global__ void __launch_bounds__(kEncThreads)
xa_encode_kernel(const EncodeParams p)
{
	typedef EncTile<BITS, CH, kEncTBE> Tile;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	typename Tile::Smem &sm = *reinterpret_cast<typename Tile::Smem *>(smem_raw);
	const uint32_t tid = threadIdx.x;
	const uint32_t bar = smem_u32(&sm.mbar);

	if (tid == 0)
		mbar_init(bar, 1);
	__syncthreads();

	Tile t(p, sm, blockIdx.x);
	if (tid == 0) {
		const uint32_t nb = t.bulk_bytes();
		if (nb) {
			mbar_expect_tx(bar, nb);
			bulk_g2s(smem_u32(sm.in), p.src + t.in0, nb, bar);
		} else {
			mbar_arrive(bar);
		}
	}
	t.load_tail(tid, kEncThreads);
	mbar_wait(bar, 0);
	__syncthreads();

	t.phase_pack(tid, kEncThreads);
	__syncthreads();
	t.phase_store(tid, kEncThreads);
}

/*
 * Searching encoder (an extension, not in the reference: xa_core.h spells out
 * the rule, the test suite holds a plain-C restatement).  A stream-channel is one serial
 * chain over all of its blocks -- every block starts from the decoder state
 * its predecessor's winner left -- so the parallelism is streams x channels x
 * candidates: ONE WARP PER STREAM-CHANNEL, lane l simulating candidates l,
 * l + 32, (l + 64) side by side (independent dependency chains, which hides
 * the latency of the predictor step), the block's 32 samples going round by
 * shuffle; warp-wide argmin on (error, candidate); the winner's packed codes
 * are handed to the lanes that store them, one payload byte per lane.
 * Compute-bound by design: 45-65 closed-loop simulations per sample.
 */
constexpr int kSearchThreads = 128;

template <int BITS, int CH>
__global__ void __launch_bounds__(kSearchThreads)
xa_search_kernel(const EncodeParams p)
{
	constexpr int NR = search_ranges(BITS), NC = search_candidates(BITS);
	constexpr int SLOTS = (NC + 31) / 32, BS = block_bytes(BITS);
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t warp = (blockIdx.x * kSearchThreads + threadIdx.x) >> 5;
	if (warp >= p.n_streams * CH)
		return;
	const uint32_t stream = p.order[warp / CH], ch = warp % CH;
	const StreamDev sd = p.streams[stream];
	const uint8_t *pcm = p.src + sd.pcm_off;
	uint8_t *xa = p.dst + sd.xa_off + ch * BS;
	const uint32_t frames = sd.pcm_len / (2u * CH);
	/* decoder state and samples carry +32768 throughout (xa_core.h:search_sample_b) */
	int s0 = sd.prev[ch][0] + 32768, s1 = sd.prev[ch][1] + 32768;

	SearchK<BITS> K[SLOTS];
	uint32_t unbias[BITS];
	search_code_bias<BITS>(unbias);
#pragma unroll
	for (int j = 0; j < SLOTS; j++) {
		const uint32_t c = lane + 32u * j;
		const uint32_t f = c < (uint32_t)NC ? c / NR : 0u, r = c < (uint32_t)NC ? c % NR : 0u;
		search_setup<BITS>(K[j], f, 16 - BITS - (int)r);
	}
	/* lane i holds sample i of the block; frames past the end are zero
	 * (src/libbjxa.c:686-690) */
	auto sample = [&](uint32_t eb) -> int {
		const uint32_t fr = eb * 32u + lane;
		return 32768 + (fr < frames ?
		    *reinterpret_cast<const int16_t *>(pcm + ((uint64_t)fr * CH + ch) * 2u) : 0);
	};
	int xn = sample(0);
	for (uint32_t eb = 0; eb < sd.blocks; eb++) {
		const int x = xn;
		if (eb + 1 < sd.blocks)
			xn = sample(eb + 1);
		int q0[SLOTS], q1[SLOTS];
		unsigned long long err[SLOTS];
		uint32_t w[SLOTS][BITS];
#pragma unroll
		for (int j = 0; j < SLOTS; j++) {
			q0[j] = s0;
			q1[j] = s1;
			err[j] = 0;
#pragma unroll
			for (int k = 0; k < BITS; k++)
				w[j][k] = 0;
		}
#pragma unroll
		for (int i = 0; i < 32; i++) {
			const int xi = __shfl_sync(0xffffffffu, x, i);
#pragma unroll
			for (int j = 0; j < SLOTS; j++)
				put_code<BITS>(w[j], i,
				    search_sample_b<BITS>(xi, K[j], q0[j], q1[j], err[j]));
		}
		/* this lane's best, then the warp's: smallest (error, candidate) */
		unsigned long long be = ~0ULL;
		uint32_t bc = 0xffffffffu;
#pragma unroll
		for (int j = 0; j < SLOTS; j++) {
			const uint32_t c = lane + 32u * j;
			if (c < (uint32_t)NC && (err[j] < be || (err[j] == be && c < bc))) {
				be = err[j];
				bc = c;
			}
		}
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			const unsigned long long oe = __shfl_xor_sync(0xffffffffu, be, o);
			const uint32_t oc = __shfl_xor_sync(0xffffffffu, bc, o);
			if (oe < be || (oe == be && oc < bc)) {
				be = oe;
				bc = oc;
			}
		}
		const uint32_t wl = bc & 31u, ws = bc >> 5;
		int wq0 = q0[0], wq1 = q1[0];
		uint32_t ww[BITS];
#pragma unroll
		for (int k = 0; k < BITS; k++)
			ww[k] = w[0][k];
#pragma unroll
		for (int j = 1; j < SLOTS; j++) {
			if (ws == (uint32_t)j) {
				wq0 = q0[j];
				wq1 = q1[j];
#pragma unroll
				for (int k = 0; k < BITS; k++)
					ww[k] = w[j][k];
			}
		}
		s0 = __shfl_sync(0xffffffffu, wq0, wl);
		s1 = __shfl_sync(0xffffffffu, wq1, wl);
		/* payload byte `lane` (and lane + 32 is never needed: 4 * BITS <= 32) */
		uint32_t mine = 0;
#pragma unroll
		for (int k = 0; k < BITS; k++) {
			const uint32_t wk = __shfl_sync(0xffffffffu, ww[k], wl) ^ unbias[k];
			if (lane >> 2 == (uint32_t)k)
				mine = wk;
		}
		uint8_t *blk = xa + (uint64_t)eb * (CH * BS);
		if (lane < 4u * BITS)
			blk[1 + lane] = (uint8_t)(mine >> (8u * (lane & 3u)));
		if (lane == 0)
			blk[0] = (uint8_t)((bc / NR) << 4 | (bc % NR));
	}
	if (lane == 0) {
		p.results[stream].prev[ch][0] = (int16_t)(s0 - 32768);
		p.results[stream].prev[ch][1] = (int16_t)(s1 - 32768);
	}
}

/* ---- error mapping -------------------------------------------------------- */

static int
cuda_errno(cudaError_t e)
{
	switch (e) {
	case cudaSuccess:
		return 0;
	case cudaErrorNoDevice:
	case cudaErrorInsufficientDriver:
	case cudaErrorInitializationError:
	case cudaErrorInvalidDevice:
	case cudaErrorDevicesUnavailable:
	case cudaErrorSystemDriverMismatch:
	case cudaErrorCompatNotSupportedOnDevice:
		return ENODEV;
	case cudaErrorMemoryAllocation:
		return ENOMEM;
	default:
		return EIO;
	}
}

#define XA_CUDA(call)					\
	do {						\
		cudaError_t e_ = (call);		\
		if (e_ != cudaSuccess) {		\
			(void)cudaGetLastError();	\
			errno = cuda_errno(e_);		\
			return (-1);			\
		}					\
	} while (0)

#define XA_CUDA_NULL(call)				\
	do {						\
		cudaError_t e_ = (call);		\
		if (e_ != cudaSuccess) {		\
			(void)cudaGetLastError();	\
			errno = cuda_errno(e_);		\
			return (NULL);			\
		}					\
	} while (0)

/* ---- device helpers ------------------------------------------------------- */

extern "C" int
bjxa_gpu_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		(void)cudaGetLastError();
		return (0);
	}
	return (n);
}

extern "C" int
bjxa_gpu_current(void)
{
	int dev = -1;
	if (cudaGetDevice(&dev) != cudaSuccess) {
		(void)cudaGetLastError();
		return (-1);
	}
	return (dev);
}

extern "C" int
bjxa_gpu_select(int device)
{
	XA_CUDA(cudaSetDevice(device));
	return (0);
}

extern "C" void *
bjxa_gpu_alloc(size_t bytes)
{
	void *p = NULL;
	XA_CUDA_NULL(cudaMalloc(&p, bytes ? bytes : 16));
	return (p);
}

extern "C" int
bjxa_gpu_free(void *dptr)
{
	XA_CUDA(cudaFree(dptr));
	return (0);
}

extern "C" void *
bjxa_host_alloc(size_t bytes)
{
	void *p = NULL;
	XA_CUDA_NULL(cudaMallocHost(&p, bytes ? bytes : 16));
	return (p);
}

extern "C" int
bjxa_host_free(void *hptr)
{
	XA_CUDA(cudaFreeHost(hptr));
	return (0);
}

extern "C" int
bjxa_gpu_upload(void *dptr, const void *hptr, size_t bytes)
{
	XA_CUDA(cudaMemcpy(dptr, hptr, bytes, cudaMemcpyHostToDevice));
	/* from pageable memory the call returns when the bytes are staged: wait
	 * for the DMA, streams that do not sync with the default one may follow */
	XA_CUDA(cudaStreamSynchronize(0));
	return (0);
}

extern "C" int
bjxa_gpu_download(void *hptr, const void *dptr, size_t bytes)
{
	XA_CUDA(cudaMemcpy(hptr, dptr, bytes, cudaMemcpyDeviceToHost));
	return (0);
}

extern "C" void *
bjxa_gpu_stream_create(void)
{
	cudaStream_t st = NULL;
	XA_CUDA_NULL(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
	return ((void *)st);
}

extern "C" int
bjxa_gpu_stream_destroy(void *cuda_stream)
{
	XA_CUDA(cudaStreamDestroy((cudaStream_t)cuda_stream));
	return (0);
}

extern "C" int
bjxa_gpu_upload_async(void *dptr, const void *hptr, size_t bytes, void *cuda_stream)
{
	XA_CUDA(cudaMemcpyAsync(dptr, hptr, bytes, cudaMemcpyHostToDevice,
	    (cudaStream_t)cuda_stream));
	return (0);
}

extern "C" int
bjxa_gpu_download_async(void *hptr, const void *dptr, size_t bytes, void *cuda_stream)
{
	XA_CUDA(cudaMemcpyAsync(hptr, dptr, bytes, cudaMemcpyDeviceToHost,
	    (cudaStream_t)cuda_stream));
	return (0);
}

extern "C" int
bjxa_gpu_sync(void *cuda_stream)
{
	XA_CUDA(cudaStreamSynchronize((cudaStream_t)cuda_stream));
	return (0);
}

/* ---- plans ---------------------------------------------------------------- */

template <class T>
struct DevBuf {
	T *p;
	size_t cap;
	DevBuf() : p(NULL), cap(0) {}
	int reserve(size_t n, bool zero)
	{
		if (n <= cap)
			return 0;
		if (p)
			cudaFree(p);
		p = NULL;
		cap = 0;
		size_t want = n + n / 4 + 16;
		cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
		if (e != cudaSuccess) {
			(void)cudaGetLastError();
			return cuda_errno(e);
		}
		if (zero) {
			/* cudaMemset only QUEUES the fill on the default stream, and
			 * the plan's kernels run on streams that do not wait for
			 * that one: a fill that lands late would wipe mailboxes
			 * already in use -- so wait for it here */
			e = cudaMemset(p, 0, want * sizeof(T));
			if (e == cudaSuccess)
				e = cudaStreamSynchronize(0);
			if (e != cudaSuccess) {
				(void)cudaGetLastError();
				(void)cudaFree(p);
				p = NULL;
				return cuda_errno(e);
			}
		}
		cap = want;
		return 0;
	}
	void release()
	{
		if (p)
			cudaFree(p);
		p = NULL;
		cap = 0;
	}
};

/* stereo tile form: 0 = direct, 1 = staged, 2 = let the census decide per launch
 * (BJXA_B200_STEREO=direct|staged|auto, default auto) */
static int
stereo_mode(void)
{
	const char *e = getenv("BJXA_B200_STEREO");
	return e == NULL ? 2 : strcmp(e, "direct") == 0 ? 0 : strcmp(e, "staged") == 0 ? 1 : 2;
}

static int pool_mode(void);
static int pool_candidate(int pool, int ns, uint32_t n_tiles);
static int split_mode(void);
static int split_candidate(int split, int bits, int ch, int stereo, int ns, uint32_t n_tiles);
static int relay_mode(void);
static int relay_candidate(int relay, int bits, int ch, int stereo, int ns, uint32_t n_tiles);
static int seg_mode(void);
static int stereo_effective(int stereo, int ch, int segc, int chainc);
static int chain_mode(void);
static int chain_candidate(int chain, int ch, uint32_t n_streams);
static int seg_candidate(int seg, uint32_t n_seg_tiles);
static int decode_class_launches(int ch, int stereo, bool alt, int poolc, int splitc, int relayc, int segc, int chainc);

struct bjxa_plan {
	uint32_t magic;
#define BJXA_PLAN_MAGIC 0x706c414eu
	HostPlan hp;
	std::vector<bjxa_stream_desc_t> descs;
	DevBuf<StreamDev> d_streams;
	DevBuf<StreamRes> d_results;
	DevBuf<uint32_t> d_first_bad;
	DevBuf<TileEnt> d_tiles;
	DevBuf<uint32_t> d_order;
	DevBuf<unsigned long long> d_carry;
	DevBuf<uint32_t> d_fault;		/* set when a carry never arrived */
	DevBuf<LiveRec> d_live;			/* split form: pass 1's records, one per tile */
	DevBuf<RelayRec> d_relay;		/* relay form: kRelayPerTile per tile */
	uint32_t epoch;
	/* last run */
	bool ran;
	uint8_t *last_dst;
	const uint8_t *last_src;
	uint64_t last_src_bytes;
	cudaStream_t last_stream;
	int launches;
	int stereo;			/* stereo_mode() when the plan was built */
	int pool;			/* pool_mode() likewise */
	int split;			/* split_mode() likewise */
	int relay;			/* relay_mode() likewise */
	int seg;			/* seg_mode() likewise */
	int chain;			/* chain_mode() likewise */
	/* a plan with several classes runs them side by side (bjxa_plan_run) */
	cudaStream_t cls_stream[6];
	cudaEvent_t ev_start, ev_done[6];
	/* BJXA_B200_CENSUS_EVERY=n > 1 (off by default): what the census chose is read back
	 * without waiting, the next runs launch that form alone (any form decodes any data;
	 * the census only picks the fastest) and every n-th run asks again.  For callers
	 * that run one plan over and over on data of one kind: a batch of another kind
	 * in the same arenas runs through the old choice until the census is asked again */
	uint32_t *h_choice;		/* pinned, 6 words */
	cudaEvent_t ev_choice;
	bool choice_pending;
	bool choice_asked[6];
	int cached[6];			/* -1: not known */
	uint32_t cached_age[6];
	uint32_t census_every;
	unsigned long long n_launched;
};

#include <atomic>
/* kernels launched by this thread (bjxa_plan_launched sums them per plan) */
static thread_local unsigned long long tls_launched = 0;

/* devices whose kernels have their shared-memory attributes set (bit = device) */
static std::atomic<unsigned long long> g_attr_devices(0);

template <class Tile, int MODE = kModePlain>
static cudaError_t
set_dec_attr(void)
{
	return cudaFuncSetAttribute(xa_decode_kernel<Tile, MODE>,
	    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(typename Tile::Smem));
}

/* the pooled form's tile type: the direct tile over a deeper ring */
template <int BITS, int CH> struct PoolTile;
template <int BITS> struct PoolTile<BITS, 1> {
	typedef DecTile<BITS, kDecTBQ, 1, pool_stages(BITS, 1)> type;
};
template <int BITS> struct PoolTile<BITS, 2> {
	typedef DecTileStereo<BITS, kDecTBQ, 1, pool_stages(BITS, 2)> type;
};

template <int BITS, int CH>
static cudaError_t
set_attrs_one(void)
{
	cudaError_t e;
	typedef typename PoolTile<BITS, CH>::type PT;
	if ((e = cudaFuncSetAttribute(xa_decode_pool_kernel<PT>,
	    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PoolSmem<PT>))) != cudaSuccess)
		return e;
	if (CH == 1) {
		if ((e = set_dec_attr<DecTile<BITS, kDecTBQ, 1, dec_stages(BITS, 1)> >()) != cudaSuccess ||
		    (e = set_dec_attr<DecTile<BITS, kDecTBQ, 1, dec_stages(BITS, 1)>, kModeSplit>()) != cudaSuccess ||
		    (e = set_dec_attr<DecTile<BITS, kDecTBQ, 1, dec_stages(BITS, 1), true>, kModeRelay>()) != cudaSuccess ||
		    (e = set_dec_attr<DecTile<BITS, kDecTBQ, kDecWide, dec_stages(BITS, 1)> >()) != cudaSuccess)
			return e;
	} else {
		if ((e = set_dec_attr<DecTileStereo<BITS, kDecTBQ, 1, dec_stages(BITS, 2)> >()) != cudaSuccess ||
		    (e = set_dec_attr<DecTileStereo<BITS, kDecTBQ, 1, dec_stages(BITS, 2)>, kModeSplit>()) != cudaSuccess ||
		    (e = set_dec_attr<DecTileStereo<BITS, kDecTBQ, 1, dec_stages(BITS, 2), true>, kModeRelay>()) != cudaSuccess ||
		    (e = set_dec_attr<DecTileStereo<BITS, kDecTBQ, kDecWide, dec_stages(BITS, 2)> >()) != cudaSuccess ||
		    (e = set_dec_attr<DecTileStaged<BITS, 2, kDecTBQ, 1, kDecStagedStages> >()) != cudaSuccess ||
		    (e = set_dec_attr<DecTileStaged<BITS, 2, kDecTBQ, kDecWide, kDecStagedStages> >()) != cudaSuccess)
			return e;
	}
	if ((e = cudaFuncSetAttribute(xa_walk_kernel<BITS, CH, false>,
	    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WalkCfg<BITS, CH>::kSmem)) != cudaSuccess ||
	    (e = cudaFuncSetAttribute(xa_walk_kernel<BITS, CH, true>,
	    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WalkCfg<BITS, CH>::kSmem)) != cudaSuccess)
		return e;
	if ((e = cudaFuncSetAttribute(xa_seg_kernel<BITS, CH>,
	    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SegCfg<BITS, CH>::kSmem)) != cudaSuccess)
		return e;
	if ((e = cudaFuncSetAttribute(xa_chain_kernel<BITS, CH>,
	    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ChainSmem<BITS, CH>))) != cudaSuccess)
		return e;
	return cudaFuncSetAttribute(xa_encode_kernel<BITS, CH>,
	    cudaFuncAttributeMaxDynamicSharedMemorySize,
	    (int)sizeof(EncSmem<BITS, CH, kEncTBE>));
}

static cudaError_t
set_attrs(void)
{
	/* the attribute belongs to the device's context: once per device; two
	 * threads that get here together both set it, which is harmless */
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess)
		return e;
	const unsigned long long bit = 1ULL << (dev & 63);
	if (g_attr_devices.load(std::memory_order_acquire) & bit)
		return cudaSuccess;
	if ((e = set_attrs_one<4, 1>()) != cudaSuccess) return e;
	if ((e = set_attrs_one<4, 2>()) != cudaSuccess) return e;
	if ((e = set_attrs_one<6, 1>()) != cudaSuccess) return e;
	if ((e = set_attrs_one<6, 2>()) != cudaSuccess) return e;
	if ((e = set_attrs_one<8, 1>()) != cudaSuccess) return e;
	if ((e = set_attrs_one<8, 2>()) != cudaSuccess) return e;
	g_attr_devices.fetch_or(bit, std::memory_order_release);
	return cudaSuccess;
}

static int
plan_upload(bjxa_plan *pl)
{
	HostPlan &hp = pl->hp;
	size_t n = hp.streams.size();
	int rc;

	if ((rc = pl->d_streams.reserve(n, false)) ||
	    (rc = pl->d_results.reserve(n, false)) ||
	    (rc = pl->d_first_bad.reserve(n + 40, false)) ||	/* + 6 ticket counters, 6 census words, 2 x 6 record counters */
	    (rc = pl->d_live.reserve(hp.kind == kKindDecode && pl->split != 0 ? hp.tile_begin[6] : 0, false)) ||
	    (rc = pl->d_relay.reserve(hp.kind == kKindDecode && pl->relay != 0 ?
	    (size_t)hp.tile_begin[6] * kRelayPerTile : 0, false)) ||
	    (rc = pl->d_fault.reserve(4, true)) ||
	    (rc = pl->d_tiles.reserve(hp.tiles.size(), false)) ||
	    (rc = pl->d_order.reserve(hp.order.size(), false)) ||
	    (rc = pl->d_carry.reserve((size_t)hp.n_slots * 2, true))) {
		errno = rc;
		return (-1);
	}
	if (n)
		XA_CUDA(cudaMemcpy(pl->d_streams.p, hp.streams.data(),
		    n * sizeof(StreamDev), cudaMemcpyHostToDevice));
	if (!hp.tiles.empty())
		XA_CUDA(cudaMemcpy(pl->d_tiles.p, hp.tiles.data(),
		    hp.tiles.size() * sizeof(TileEnt), cudaMemcpyHostToDevice));
	if (!hp.order.empty())
		XA_CUDA(cudaMemcpy(pl->d_order.p, hp.order.data(),
		    hp.order.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
	/* a copy from pageable memory returns once the bytes are staged; the DMA
	 * runs on the default stream, which the caller's (non-blocking) stream does
	 * not wait for: the tables must have landed before a kernel reads them */
	XA_CUDA(cudaStreamSynchronize(0));
	return (0);
}

static int
plan_build(bjxa_plan *pl, int kind, const bjxa_stream_desc_t *descs, size_t n)
{
	size_t bad = 0;

	if (kind != BJXA_PLAN_DECODE && kind != BJXA_PLAN_ENCODE && kind != BJXA_PLAN_ENCODE_SEARCH) {
		errno = EINVAL;
		return (-1);
	}
	if (descs == NULL && n != 0) {
		errno = EFAULT;
		return (-1);
	}
	/* BJXA_B200_STRIPS=1|32 forces the decode tile shape (tuning aid) */
	const char *env = getenv("BJXA_B200_STRIPS");
	int rc;
	try {
		rc = build_plan(pl->hp, kind, descs, n, &bad, env ? atoi(env) : 0,
		    seg_mode() == 1 ? 0 : kSegMinItems);
		if (rc == 0)
			pl->descs.assign(descs, descs + n);
	} catch (const std::bad_alloc &) {
		rc = ENOMEM;	/* nothing of C++ may unwind into the C callers */
	}
	if (rc) {
		errno = rc;
		return (-1);
	}
	pl->ran = false;
	pl->launches = 0;
	pl->choice_pending = false;
	for (int b = 0; b < 6; b++) {
		pl->cached[b] = -1;
		pl->cached_age[b] = 0;
		pl->choice_asked[b] = false;
	}
	{
		const char *e = getenv("BJXA_B200_CENSUS_EVERY");
		pl->census_every = e != NULL && atoi(e) > 0 ? (uint32_t)atoi(e) : 1u;
	}
	pl->stereo = stereo_mode();
	pl->pool = pool_mode();
	pl->split = split_mode();
	pl->relay = relay_mode();
	pl->seg = seg_mode();
	pl->chain = chain_mode();
	for (int b = 0; b < 6; b++)
		if (pl->hp.order_begin[b + 1] > pl->hp.order_begin[b]) {
			const uint32_t nt = pl->hp.tile_begin[b + 1] - pl->hp.tile_begin[b];
			const int segc = seg_candidate(pl->seg, pl->hp.seg_begin[b + 1] - pl->hp.seg_begin[b]);
			const int chainc = chain_candidate(pl->chain, bucket_ch(b),
			    pl->hp.order_begin[b + 1] - pl->hp.order_begin[b]);
			const int stereo = stereo_effective(pl->stereo, bucket_ch(b), segc, chainc);
			pl->launches += kind == kKindDecode ? decode_class_launches(bucket_ch(b),
			    stereo, pl->hp.alt_ns[b] != 0, pool_candidate(pl->pool, pl->hp.ns[b], nt),
			    split_candidate(pl->split, bucket_bits(b), bucket_ch(b), stereo, pl->hp.ns[b], nt),
			    relay_candidate(pl->relay, bucket_bits(b), bucket_ch(b), stereo, pl->hp.ns[b], nt),
			    segc, chainc) : 1;
		}
	return (plan_upload(pl));
}

extern "C" bjxa_plan_t *
bjxa_plan_create(int kind, const bjxa_stream_desc_t *descs, size_t n)
{
	if (bjxa_gpu_count() <= 0) {
		errno = ENODEV;
		return (NULL);
	}
	bjxa_plan *pl = new (std::nothrow) bjxa_plan();
	if (pl == NULL) {
		errno = ENOMEM;
		return (NULL);
	}
	pl->magic = BJXA_PLAN_MAGIC;
	pl->epoch = 0;
	pl->ran = false;
	pl->h_choice = NULL;
	pl->ev_choice = NULL;
	pl->n_launched = 0;
	if (plan_build(pl, kind, descs, n) < 0) {
		int e = errno;
		bjxa_plan_free(&pl);
		errno = e;
		return (NULL);
	}
	return (pl);
}

#define CHECK_PLAN(pl)							\
	do {								\
		if ((pl) == NULL) { errno = EFAULT; return (-1); }	\
		if ((pl)->magic != BJXA_PLAN_MAGIC) { errno = EINVAL; return (-1); } \
	} while (0)

extern "C" int
bjxa_plan_reset(bjxa_plan_t *pl, int kind, const bjxa_stream_desc_t *descs,
    size_t n)
{
	CHECK_PLAN(pl);
	if (pl->ran)
		XA_CUDA(cudaStreamSynchronize(pl->last_stream));
	return (plan_build(pl, kind, descs, n));
}

extern "C" int
bjxa_plan_free(bjxa_plan_t **planp)
{
	if (planp == NULL) {
		errno = EFAULT;
		return (-1);
	}
	bjxa_plan *pl = *planp;
	CHECK_PLAN(pl);
	if (pl->ran)
		(void)cudaStreamSynchronize(pl->last_stream);
	pl->d_streams.release();
	pl->d_results.release();
	pl->d_first_bad.release();
	pl->d_tiles.release();
	pl->d_order.release();
	pl->d_carry.release();
	pl->d_fault.release();
	pl->d_live.release();
	pl->d_relay.release();
	for (int b = 0; b < 6; b++) {
		if (pl->cls_stream[b] != NULL)
			(void)cudaStreamDestroy(pl->cls_stream[b]);
		if (pl->ev_done[b] != NULL)
			(void)cudaEventDestroy(pl->ev_done[b]);
	}
	if (pl->ev_start != NULL)
		(void)cudaEventDestroy(pl->ev_start);
	if (pl->ev_choice != NULL)
		(void)cudaEventDestroy(pl->ev_choice);
	if (pl->h_choice != NULL)
		(void)cudaFreeHost(pl->h_choice);
	pl->magic = 0;
	delete pl;
	*planp = NULL;
	return (0);
}

extern "C" int
bjxa_plan_launches(const bjxa_plan_t *pl)
{
	CHECK_PLAN(pl);
	return (pl->launches);
}

extern "C" unsigned long long
bjxa_plan_launched(const bjxa_plan_t *pl)
{
	if (pl == NULL || pl->magic != BJXA_PLAN_MAGIC) {
		errno = pl == NULL ? EFAULT : EINVAL;
		return (0);
	}
	return (pl->n_launched);
}

extern "C" int
bjxa_plan_extent(const bjxa_plan_t *pl, uint64_t *src_bytes, uint64_t *dst_bytes)
{
	CHECK_PLAN(pl);
	if (src_bytes)
		*src_bytes = pl->hp.src_need;
	if (dst_bytes)
		*dst_bytes = pl->hp.dst_need;
	return (0);
}

template <class Tile, int MODE = kModePlain>
static cudaError_t
launch_persistent(const DecodeParams &p, cudaStream_t st)
{
	/* persistent: as many CTAs as fit the device at once, never more than tiles */
	static thread_local int grid_cache[2] = { -1, 0 };
	const size_t smem = sizeof(typename Tile::Smem);
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess)
		return e;
	if (grid_cache[0] != dev) {
		int per_sm = 0, sms = 0;
		e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm,
		    xa_decode_kernel<Tile, MODE>, kDecBlock, smem);
		if (e != cudaSuccess)
			return e;
		e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		if (e != cudaSuccess)
			return e;
		grid_cache[0] = dev;
		grid_cache[1] = per_sm * sms > 0 ? per_sm * sms : sms;
	}
	uint32_t grid = (uint32_t)grid_cache[1];
	if (grid > p.n_tiles)
		grid = p.n_tiles;
	xa_decode_kernel<Tile, MODE><<<grid, kDecBlock, smem, st>>>(p);
	tls_launched++;
	return cudaGetLastError();
}

template <class Tile>
static cudaError_t
launch_pool(const DecodeParams &p, cudaStream_t st)
{
	static thread_local int grid_cache[2] = { -1, 0 };
	const size_t smem = sizeof(PoolSmem<Tile>);
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess)
		return e;
	if (grid_cache[0] != dev) {
		int per_sm = 0, sms = 0;
		e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm,
		    xa_decode_pool_kernel<Tile>, kPoolBlock, smem);
		if (e != cudaSuccess)
			return e;
		e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		if (e != cudaSuccess)
			return e;
		grid_cache[0] = dev;
		grid_cache[1] = per_sm * sms > 0 ? per_sm * sms : sms;
	}
	uint32_t grid = (uint32_t)grid_cache[1];
	if (grid > p.n_tiles)
		grid = p.n_tiles;
	xa_decode_pool_kernel<Tile><<<grid, kPoolBlock, smem, st>>>(p);
	tls_launched++;
	return cudaGetLastError();
}

/*
 * One class of a decode plan.  Up to four tile forms can serve it -- long-strip
 * or wide tiles, and for stereo the direct or the staged form (xa_tile.h
 * explains them, profiles/history_r1.md has the numbers).  When more than one
 * is possible the census kernel decides on the device and every candidate is
 * launched; the ones not chosen return at once.
 */
struct DecodeClass {
	DecodeParams p;			/* tiles / n_tiles: the primary list */
	const TileEnt *alt_tiles;	/* the list in the wide shape, or NULL */
	uint32_t alt_n;
	int ns;				/* strips per tile of the primary list */
	int stereo;			/* 0 direct, 1 staged, 2 census (stereo classes) */
	int pool;			/* 0 never, 1 always, 2 census */
	int split;			/* likewise */
	int relay;			/* likewise */
	int seg;			/* likewise */
	int chain;			/* likewise */
	const TileEnt *seg_tiles;	/* the segment form's list, or NULL */
	uint32_t seg_n;
	uint32_t *d_choice;
	bool *censused;			/* set when the census kernel was launched */
	const uint32_t *d_order;	/* the class's streams */
	uint32_t n_streams;
};

template <int BITS, int CH, int NS>
static cudaError_t
launch_form(const DecodeParams &p, bool staged, cudaStream_t st)
{
	if (CH == 1)
		return launch_persistent<DecTile<BITS, kDecTBQ, NS, dec_stages(BITS, 1)> >(p, st);
	if (staged)
		return launch_persistent<DecTileStaged<BITS, 2, kDecTBQ, NS, kDecStagedStages> >(p, st);
	return launch_persistent<DecTileStereo<BITS, kDecTBQ, NS, dec_stages(BITS, 2)> >(p, st);
}

/* the direct form over long strips as the first pass of the split / relay form */
template <int BITS, int CH, int MODE>
static cudaError_t
launch_pass1(const DecodeParams &p, cudaStream_t st)
{
	constexpr bool R = MODE == kModeRelay;
	if (CH == 1)
		return launch_persistent<DecTile<BITS, kDecTBQ, 1, dec_stages(BITS, 1), R>, MODE>(p, st);
	return launch_persistent<DecTileStereo<BITS, kDecTBQ, 1, dec_stages(BITS, 2), R>, MODE>(p, st);
}

/* pooled form: 0 = never, 1 = always (long-strip lists), 2 = let the census decide
 * (BJXA_B200_POOL=off|on|auto).  Default off: as measured in round 1 it only wins
 * on chain-rich 4-bit stereo data (profiles/history_r1.md, step 18) */
static int
pool_mode(void)
{
	const char *e = getenv("BJXA_B200_POOL");
	return e == NULL ? 0 : strcmp(e, "on") == 0 ? 1 : strcmp(e, "auto") == 0 ? 2 : 0;
}

/* is the pooled form a candidate for this class, and the only one? */
static int
pool_candidate(int pool, int ns, uint32_t n_tiles)
{
	if (ns != 1 || pool == 0)
		return 0;
	return pool == 1 ? 1 : n_tiles >= kPoolMinTiles ? 2 : 0;
}

/*
 * split form (xa_walk.h): 0 = never, 1 = always, 2 = let the census decide
 * (BJXA_B200_SPLIT=off|on|auto, default auto).
 */
static int
split_mode(void)
{
	const char *e = getenv("BJXA_B200_SPLIT");
	return e == NULL ? 2 : strcmp(e, "on") == 0 ? 1 : strcmp(e, "off") == 0 ? 0 : 2;
}

/*
 * From this share of chain blocks (permille) the census sends a class to the
 * split form, and relay_permille() below to the relay form.  Measured crossovers
 * at 4096 streams x 30 s (profiles/history_r2.md): the relay form wins from about
 * a tenth of chain blocks; the split form overtakes it on chain-rich stereo data
 * (and everywhere on 4-bit stereo), on mono only at 4 bits and above three
 * quarters.
 */
constexpr uint32_t split_permille(int bits, int ch)
{
	return ch == 2 ? (bits == 4 ? 80u : 650u) : (bits == 4 ? 750u : 1001u);
}
/* below this many tiles a launch is too short for a census to pay */
constexpr uint32_t kSplitMinTiles = 64;

/* is the split form a candidate for this class (2), or the only one (1)?  It is
 * built on the direct form over long strips. */
static int
split_candidate(int split, int bits, int ch, int stereo, int ns, uint32_t n_tiles)
{
	if (ns != 1 || split == 0 || (ch == 2 && stereo == 1))
		return 0;
	if (split == 1)
		return 1;
	return n_tiles >= kSplitMinTiles && split_permille(bits, ch) <= 1000u ? 2 : 0;
}

/*
 * relay form (xa_walk.h): 0 = never, 1 = always, 2 = let the census decide
 * (BJXA_B200_RELAY=off|on|auto, default auto).
 */
static int
relay_mode(void)
{
	const char *e = getenv("BJXA_B200_RELAY");
	return e == NULL ? 2 : strcmp(e, "on") == 0 ? 1 : strcmp(e, "off") == 0 ? 0 : 2;
}

/* from this share of chain blocks (permille) the census sends a class to the relay
 * form (below it the plain direct form; above split_permille the split form) */
constexpr uint32_t relay_permille(int bits, int ch)
{
	/* Since the segment form takes chain-rich data, the relay form's band is narrow,
	 * and in it it beats the direct form only for 4/6-bit mono (by 1.5 points at 20 %
	 * chain blocks; 8-bit mono and stereo: nothing, profiles/auto_sweep_r2.log against
	 * round 1's table): elsewhere its two launches are not even candidates */
	return ch == 2 || bits == 8 ? 1001u : 150u;
}

static int
relay_candidate(int relay, int bits, int ch, int stereo, int ns, uint32_t n_tiles)
{
	if (ns != 1 || relay == 0 || (ch == 2 && stereo == 1))
		return 0;
	if (relay == 1)
		return 1;
	return n_tiles >= kSplitMinTiles && relay_permille(bits, ch) <= 1000u ? 2 : 0;
}

/*
 * segment form (xa_walk.h): 0 = never, 1 = always, 2 = let the census decide
 * (BJXA_B200_SEG=off|on|auto, default auto).  For classes that have its list.
 */
static int
seg_mode(void)
{
	const char *e = getenv("BJXA_B200_SEG");
	return e == NULL ? 2 : strcmp(e, "on") == 0 ? 1 : strcmp(e, "off") == 0 ? 0 : 2;
}

/*
 * From this share of chain blocks (permille) the census sends a class to the
 * segment form, and from seg_below() on to whatever the other thresholds say.
 * Measured crossovers at 4096 streams x 30 s (profiles/history_r2.md): the form's
 * rate hardly depends on the mix (every block goes through the chain step: 61-70 %
 * of the HBM peak for mono, 60-68 % for 6/8-bit stereo, 65-71 % for 4-bit stereo),
 * so it takes over where the tile forms, which get slower with every chain block,
 * fall below it.
 */
#ifndef XA_SEG_PERMILLE
#define XA_SEG_PERMILLE 0
#endif
#ifndef XA_SEG_BELOW
#define XA_SEG_BELOW 0
#endif
constexpr uint32_t seg_permille(int bits, int ch)
{
	return XA_SEG_PERMILLE != 0 ? XA_SEG_PERMILLE :
	    ch == 2 ? (bits == 4 ? 80u : bits == 6 ? 110u : 190u) :
	    (bits == 4 ? 350u : bits == 6 ? 300u : 350u);
}
/* above this share lanes too often find no cut block within kSegBack items and wait
 * for their neighbours (0.93^48 = 3 % of the lanes, a second pass for most tiles) */
constexpr uint32_t seg_below(int, int)
{
	return XA_SEG_BELOW != 0 ? XA_SEG_BELOW : 930u;
}

static int
seg_candidate(int seg, uint32_t n_seg_tiles)
{
	return n_seg_tiles == 0 ? 0 : seg;
}

template <int BITS, int CH>
static cudaError_t
launch_seg(const DecodeParams &p, cudaStream_t st)
{
	typedef SegCfg<BITS, CH> C;
	static thread_local int grid_cache[2] = { -1, 0 };
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess)
		return e;
	if (grid_cache[0] != dev) {
		int per_sm = 0, sms = 0;
		e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm,
		    xa_seg_kernel<BITS, CH>, C::kThreads, C::kSmem);
		if (e != cudaSuccess)
			return e;
		e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		if (e != cudaSuccess)
			return e;
		grid_cache[0] = dev;
		grid_cache[1] = per_sm * sms > 0 ? per_sm * sms : sms;
	}
	/* never more warps than tiles */
	uint32_t grid = (uint32_t)grid_cache[1];
	const uint32_t need = (p.n_tiles + C::kWarps - 1) / C::kWarps;
	if (grid > need)
		grid = need;
	xa_seg_kernel<BITS, CH><<<grid, C::kThreads, C::kSmem, st>>>(p);
	tls_launched++;
	return cudaGetLastError();
}

/*
 * chain form (xa_chain_kernel): 0 = never, 1 = always, 2 = let the census decide
 * (BJXA_B200_CHAIN=on|off|auto, default auto).  Mono classes of at least
 * kChainMinStreams streams (fewer chains than that are not worth a kernel).
 */
static int
chain_mode(void)
{
	const char *e = getenv("BJXA_B200_CHAIN");
	return e == NULL ? 2 : strcmp(e, "on") == 0 ? 1 : strcmp(e, "off") == 0 ? 0 : 2;
}

constexpr uint32_t kChainMinStreams = 128;
/* from this share of chain blocks (permille): where the mono wide tiles used to take over */
constexpr uint32_t kChainPermille = 985, kChainPermilleStereo = 930;

static int
chain_candidate(int chain, int ch, uint32_t n_streams)
{
	(void)ch;
	if (chain == 0)
		return 0;
	return chain == 1 ? 1 : n_streams >= kChainMinStreams ? 2 : 0;
}

template <int BITS, int CH>
static cudaError_t
launch_chain(const DecodeParams &p, const uint32_t *d_order, uint32_t n_streams, cudaStream_t st)
{
	xa_chain_kernel<BITS, CH><<<(n_streams + 31u) / 32u, kChainThreads + 64 * (CH - 1), sizeof(ChainSmem<BITS, CH>), st>>>(
	    p, d_order, n_streams);
	tls_launched++;
	return cudaGetLastError();
}

static int
stereo_effective(int stereo, int ch, int segc, int chainc)
{
	return ch == 2 && stereo == 2 && segc == 2 && chainc == 2 ? 0 : stereo;
}

/* how many kernels decode_class() launches */
static int
decode_class_launches(int ch, int stereo, bool alt, int poolc, int splitc, int relayc, int segc, int chainc)
{
	if (poolc == 1 || segc == 1 || chainc == 1)
		return 1;
	if (splitc == 1 || relayc == 1)
		return 2;
	const int forms = (ch == 2 && stereo == 2 ? 2 : 1) + (alt && chainc != 2 ? 1 : 0) + (poolc == 2 ? 1 : 0) +
	    (splitc == 2 ? 2 : 0) + (relayc == 2 ? 2 : 0) + (segc == 2 ? 1 : 0) + (chainc == 2 ? 1 : 0);
	return forms == 1 ? 1 : forms + 1;
}

template <int BITS, int CH, bool RELAY>
static cudaError_t
launch_walk(const DecodeParams &p, cudaStream_t st)
{
	typedef WalkCfg<BITS, CH> C;
	static thread_local int grid_cache[2] = { -1, 0 };
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess)
		return e;
	if (grid_cache[0] != dev) {
		int per_sm = 0, sms = 0;
		e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm,
		    xa_walk_kernel<BITS, CH, RELAY>, C::kThreads, C::kSmem);
		if (e != cudaSuccess)
			return e;
		e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		if (e != cudaSuccess)
			return e;
		grid_cache[0] = dev;
		grid_cache[1] = per_sm * sms > 0 ? per_sm * sms : sms;
	}
	/* never more warps than tiles: a warp without a record leaves at once */
	uint32_t grid = (uint32_t)grid_cache[1];
	const uint32_t need = (p.n_tiles + C::kWarps - 1) / C::kWarps;
	if (grid > need)
		grid = need;
	xa_walk_kernel<BITS, CH, RELAY><<<grid, C::kThreads, C::kSmem, st>>>(p);
	tls_launched++;
	return cudaGetLastError();
}

/* the class with every choice made: the form the census picked in an earlier run */
static DecodeClass
chosen_class(const DecodeClass &c, uint32_t choice)
{
	DecodeClass f = c;
	f.pool = f.split = f.relay = f.seg = f.chain = 0;
	f.censused = NULL;
	if (choice & kFormSeg) {
		f.seg = 1;
	} else if (choice & kFormChain) {
		f.chain = 1;
	} else if (choice & kFormRelay) {
		f.relay = 1;
		f.stereo = 0;
	} else if (choice & kFormSplit) {
		f.split = 1;
		f.stereo = 0;
	} else if (choice & kFormPool) {
		f.pool = 1;
	} else if ((choice & kFormWide) && c.alt_tiles != NULL) {
		f.p.tiles = c.alt_tiles;
		f.p.n_tiles = c.alt_n;
		f.ns = kDecWide;
		f.stereo = (choice & kFormStaged) ? 1 : 0;
	} else {
		f.stereo = (choice & kFormStaged) ? 1 : 0;
	}
	f.alt_tiles = NULL;
	f.alt_n = 0;
	return f;
}

template <int BITS, int CH>
static cudaError_t
decode_class(const DecodeClass &c, cudaStream_t st)
{
	const int segc = seg_candidate(c.seg, c.seg_tiles != NULL ? c.seg_n : 0u);
	const int chainc = chain_candidate(c.chain, CH, c.n_streams);
	/* where the segment and chain forms are candidates they take everything the staged
	 * stereo form used to be chosen for: it is not launched */
	const int stereo = stereo_effective(c.stereo, CH, segc, chainc);
	const bool pick_form = CH == 2 && stereo == 2;
	const int poolc = pool_candidate(c.pool, c.ns, c.p.n_tiles);
	const int splitc = split_candidate(c.split, BITS, CH, stereo, c.ns, c.p.n_tiles);
	const int relayc = relay_candidate(c.relay, BITS, CH, stereo, c.ns, c.p.n_tiles);
	/* where the chain form is a candidate, the wide tiles (data without cut blocks was
	 * all they were for) are never picked: not launched */
	const bool alt = c.alt_tiles != NULL && chainc != 2;
	DecodeParams p = c.p;
	cudaError_t e;
	if (chainc == 1)
		return launch_chain<BITS, CH>(p, c.d_order, c.n_streams, st);
	if (segc == 1) {
		p.tiles = c.seg_tiles;
		p.n_tiles = c.seg_n;
		return launch_seg<BITS, CH>(p, st);
	}
	if (poolc == 1)
		return launch_pool<typename PoolTile<BITS, CH>::type>(p, st);
	if (splitc == 1) {
		/* pass 1: the direct form without walkers; pass 2: the walkers */
		p.split = 2;
		if ((e = launch_pass1<BITS, CH, kModeSplit>(p, st)) != cudaSuccess)
			return e;
		return launch_walk<BITS, CH, false>(p, st);
	}
	if (relayc == 1) {
		/* pass 1: the direct form, its walker warps handing stragglers on */
		p.relay = 2;
		if ((e = launch_pass1<BITS, CH, kModeRelay>(p, st)) != cudaSuccess)
			return e;
		return launch_walk<BITS, CH, true>(p, st);
	}
	if (!alt && !pick_form && poolc == 0 && splitc == 0 && relayc == 0 && segc == 0 && chainc == 0) {
		const bool staged = CH == 2 && stereo == 1;
		return c.ns == 1 ? launch_form<BITS, CH, 1>(p, staged, st) :
		    launch_form<BITS, CH, kDecWide>(p, staged, st);
	}
	xa_census_kernel<<<1, kCensusThreads, 0, st>>>(p.src, p.streams, c.d_order, c.n_streams,
	    (uint32_t)block_bytes(BITS), (uint32_t)CH,
	    pick_form ? staged_permille(BITS) : stereo == 1 && CH == 2 ? 0u : kNever,
	    /* with the split form at hand, wide tiles are for all-chain data only */
	    alt ? (CH == 2 ? kWidePermilleStereo : kWidePermilleMono)[splitc != 2 && c.n_streams >= kWideManyStreams] :
	    kNever, poolc == 2 ? (CH == 2 ? kPoolPermilleStereo : kPoolPermilleMono) : kNever,
	    splitc == 2 ? split_permille(BITS, CH) : kNever,
	    relayc == 2 ? relay_permille(BITS, CH) : kNever,
	    segc == 2 ? seg_permille(BITS, CH) : kNever, seg_below(BITS, CH),
	    chainc == 2 ? (CH == 2 ? kChainPermilleStereo : kChainPermille) : kNever, c.d_choice);
	tls_launched++;
	if (c.censused != NULL)
		*c.censused = true;
	if ((e = cudaGetLastError()) != cudaSuccess)
		return e;
	p.choice = c.d_choice;
	/* long strips: direct and/or staged */
	for (int staged = 0; staged < 2; staged++) {
		if (CH == 1 ? staged == 1 : !pick_form && staged != stereo)
			continue;
		/* with a wide list, chain-heavy stereo data goes to its staged form
		 * only: "staged, long strips" then means bit 0 without bit 1 */
		p.want = staged ? kFormStaged : 0u;
		e = c.ns == 1 ? launch_form<BITS, CH, 1>(p, staged != 0, st) :
		    launch_form<BITS, CH, kDecWide>(p, staged != 0, st);
		if (e != cudaSuccess)
			return e;
	}
	/* the two-pass forms: both of their kernels leave at once unless chosen */
	if (splitc == 2) {
		p.split = 1;
		if ((e = launch_pass1<BITS, CH, kModeSplit>(p, st)) != cudaSuccess ||
		    (e = launch_walk<BITS, CH, false>(p, st)) != cudaSuccess)
			return e;
		p.split = 0;
	}
	if (relayc == 2) {
		p.relay = 1;
		if ((e = launch_pass1<BITS, CH, kModeRelay>(p, st)) != cudaSuccess ||
		    (e = launch_walk<BITS, CH, true>(p, st)) != cudaSuccess)
			return e;
		p.relay = 0;
	}
	if (poolc == 2) {
		p.want = kFormPool;
		if ((e = launch_pool<typename PoolTile<BITS, CH>::type>(p, st)) != cudaSuccess)
			return e;
	}
	if (chainc == 2) {
		DecodeParams q = p;
		q.want = kFormChain;
		if ((e = launch_chain<BITS, CH>(q, c.d_order, c.n_streams, st)) != cudaSuccess)
			return e;
	}
	if (segc == 2) {
		DecodeParams q = p;
		q.tiles = c.seg_tiles;
		q.n_tiles = c.seg_n;
		q.want = kFormSeg;
		if ((e = launch_seg<BITS, CH>(q, st)) != cudaSuccess)
			return e;
	}
	if (alt) {
		/* the wide list: mono direct; stereo staged (the share that selects
		 * wide tiles is far above the one that selects the staged form) */
		p.tiles = c.alt_tiles;
		p.n_tiles = c.alt_n;
		const bool staged = CH == 2 && stereo != 0;
		p.want = kFormWide | (staged ? kFormStaged : 0u);
		e = launch_form<BITS, CH, kDecWide>(p, staged, st);
	}
	return e;
}

template <int BITS, int CH>
static cudaError_t
launch_encode(const EncodeParams &p, cudaStream_t st)
{
	xa_encode_kernel<BITS, CH><<<p.n_tiles, kEncThreads,
	    sizeof(EncSmem<BITS, CH, kEncTBE>), st>>>(p);
	tls_launched++;
	return cudaGetLastError();
}

extern "C" int
bjxa_plan_run(bjxa_plan_t *pl, void *dst, size_t dst_bytes, const void *src,
    size_t src_bytes, void *cuda_stream)
{
	CHECK_PLAN(pl);
	HostPlan &hp = pl->hp;
	cudaStream_t st = (cudaStream_t)cuda_stream;

	if (hp.order.empty()) {		/* no stream has any block */
		pl->ran = true;
		pl->last_stream = st;
		pl->last_dst = (uint8_t *)dst;
		pl->last_src = (const uint8_t *)src;
		pl->last_src_bytes = src_bytes;
		return (0);
	}
	if (dst == NULL || src == NULL) {
		errno = EFAULT;
		return (-1);
	}
	if (((uintptr_t)dst | (uintptr_t)src) & 15u) {
		errno = EINVAL;
		return (-1);
	}
	if (src_bytes < hp.src_need || dst_bytes < hp.dst_need) {
		errno = ENOBUFS;
		return (-1);
	}
	XA_CUDA(set_attrs());
	const unsigned long long launched0 = tls_launched;
	if (pl->choice_pending && cudaEventQuery(pl->ev_choice) == cudaSuccess) {
		for (int b = 0; b < 6; b++)
			if (pl->choice_asked[b]) {
				pl->cached[b] = (int)pl->h_choice[b];
				pl->cached_age[b] = 0;
				pl->choice_asked[b] = false;
			}
		pl->choice_pending = false;
	}
	bool asked_now = false;

	size_t n = hp.streams.size();
	pl->epoch++;
	if (pl->epoch == 0) {		/* wrapped: mailboxes may hold stale tags */
		XA_CUDA(cudaMemsetAsync(pl->d_carry.p, 0,
		    pl->d_carry.cap * sizeof(unsigned long long), st));
		pl->epoch = 1;
	}
	if (hp.kind == kKindDecode) {
		/* one memset arms both the per-stream "first bad block" words and
		 * the six tile-ticket counters that follow them (~0 = no ticket
		 * drawn yet, see the producer in xa_decode_kernel) */
		XA_CUDA(cudaMemsetAsync(pl->d_first_bad.p, 0xff,
		    (n + 40) * sizeof(uint32_t), st));
	}

	/*
	 * The classes of a plan are independent (different streams of the batch,
	 * disjoint bytes), so with more than one they run side by side on streams
	 * of the plan's own, forked from and joined to the caller's stream by
	 * events: the censuses all run at once, and the tail of one class's
	 * persistent kernel fills up with the next class's CTAs instead of
	 * draining the device six times.
	 */
	int n_active = 0;
	for (int b = 0; b < 6; b++)
		n_active += hp.order_begin[b + 1] > hp.order_begin[b];
	const bool fork = n_active > 1;
	const cudaStream_t user_st = st;
	if (fork) {
		if (pl->ev_start == NULL)
			XA_CUDA(cudaEventCreateWithFlags(&pl->ev_start, cudaEventDisableTiming));
		XA_CUDA(cudaEventRecord(pl->ev_start, user_st));
	}

	for (int b = 0; b < 6; b++) {
		uint32_t t0 = hp.tile_begin[b], t1 = hp.tile_begin[b + 1];
		const uint32_t n_class = hp.order_begin[b + 1] - hp.order_begin[b];
		if (n_class == 0)
			continue;
		if (fork) {
			if (pl->cls_stream[b] == NULL) {
				XA_CUDA(cudaStreamCreateWithFlags(&pl->cls_stream[b],
				    cudaStreamNonBlocking));
				XA_CUDA(cudaEventCreateWithFlags(&pl->ev_done[b],
				    cudaEventDisableTiming));
			}
			st = pl->cls_stream[b];
			XA_CUDA(cudaStreamWaitEvent(st, pl->ev_start, 0));
		}
		cudaError_t e;
		if (hp.kind == kKindSearch) {
			EncodeParams p;
			p.src = (const uint8_t *)src;
			p.src_bytes = src_bytes;
			p.dst = (uint8_t *)dst;
			p.dst_bytes = dst_bytes;
			p.streams = pl->d_streams.p;
			p.tiles = NULL;
			p.n_tiles = 0;
			p.results = pl->d_results.p;
			p.order = pl->d_order.p + hp.order_begin[b];
			p.n_streams = n_class;
			const uint32_t warps = n_class * (uint32_t)bucket_ch(b);
			const uint32_t grid = (warps * 32u + kSearchThreads - 1) / kSearchThreads;
			switch (b) {
			case 0: xa_search_kernel<4, 1><<<grid, kSearchThreads, 0, st>>>(p); break;
			case 1: xa_search_kernel<4, 2><<<grid, kSearchThreads, 0, st>>>(p); break;
			case 2: xa_search_kernel<6, 1><<<grid, kSearchThreads, 0, st>>>(p); break;
			case 3: xa_search_kernel<6, 2><<<grid, kSearchThreads, 0, st>>>(p); break;
			case 4: xa_search_kernel<8, 1><<<grid, kSearchThreads, 0, st>>>(p); break;
			default: xa_search_kernel<8, 2><<<grid, kSearchThreads, 0, st>>>(p); break;
			}
			tls_launched++;
			e = cudaGetLastError();
		} else if (hp.kind == kKindDecode) {
			DecodeParams p;
			p.src = (const uint8_t *)src;
			p.src_bytes = src_bytes;
			p.dst = (uint8_t *)dst;
			p.streams = pl->d_streams.p;
			p.results = pl->d_results.p;
			p.first_bad = pl->d_first_bad.p;
			p.tiles = pl->d_tiles.p + t0;
			p.n_tiles = t1 - t0;
			p.order = pl->d_order.p;
			p.carry = pl->d_carry.p;
			p.ticket = reinterpret_cast<unsigned long long *>(
			    pl->d_first_bad.p + ((n + 3) & ~(size_t)3)) + b;
			p.fault = pl->d_fault.p;
			{
				const char *t = getenv("BJXA_B200_CARRY_TIMEOUT_S");
				const long secs = t != NULL && atol(t) > 0 ? atol(t) : XA_CARRY_TIMEOUT_S;
				p.carry_timeout_ns = (unsigned long long)secs * 1000000000ULL;
			}
			p.epoch = pl->epoch;
			p.choice = NULL;
			p.want = 0;
			p.split = 0;
			p.live = pl->d_live.p != NULL ? pl->d_live.p + t0 : NULL;
			p.live_count = pl->d_first_bad.p + ((n + 3) & ~(size_t)3) + 18 + b;
			p.relay = 0;
			p.relay_recs = pl->d_relay.p != NULL ? pl->d_relay.p + (size_t)t0 * kRelayPerTile : NULL;
			p.relay_count = pl->d_first_bad.p + ((n + 3) & ~(size_t)3) + 24 + b;
			DecodeClass c;
			c.p = p;
			c.alt_tiles = hp.alt_ns[b] != 0 ? pl->d_tiles.p + hp.alt_begin[b] : NULL;
			c.alt_n = hp.alt_begin[b + 1] - hp.alt_begin[b];
			c.ns = hp.ns[b];
			c.stereo = pl->stereo;
			c.pool = pl->pool;
			c.split = pl->d_live.p != NULL ? pl->split : 0;
			c.relay = pl->d_relay.p != NULL ? pl->relay : 0;
			c.seg = pl->seg;
			c.chain = pl->chain;
			c.seg_tiles = hp.seg_begin[b + 1] > hp.seg_begin[b] ? pl->d_tiles.p + hp.seg_begin[b] : NULL;
			c.seg_n = hp.seg_begin[b + 1] - hp.seg_begin[b];
			c.d_choice = pl->d_first_bad.p + ((n + 3) & ~(size_t)3) + 12 + b;
			c.d_order = pl->d_order.p + hp.order_begin[b];
			c.n_streams = hp.order_begin[b + 1] - hp.order_begin[b];
			bool censused = false;
			c.censused = &censused;
			if (pl->cached[b] >= 0 && pl->cached_age[b] + 1u < pl->census_every && !pl->choice_pending) {
				c = chosen_class(c, (uint32_t)pl->cached[b]);
				pl->cached_age[b]++;
			}
			switch (b) {
			case 0: e = decode_class<4, 1>(c, st); break;
			case 1: e = decode_class<4, 2>(c, st); break;
			case 2: e = decode_class<6, 1>(c, st); break;
			case 3: e = decode_class<6, 2>(c, st); break;
			case 4: e = decode_class<8, 1>(c, st); break;
			default: e = decode_class<8, 2>(c, st); break;
			}
			if (e == cudaSuccess && censused && !pl->choice_pending && pl->census_every > 1u) {
				/* the choice comes back behind the class's kernels; nobody waits for it */
				if (pl->h_choice == NULL) {
					XA_CUDA(cudaHostAlloc((void **)&pl->h_choice, 8 * sizeof(uint32_t), cudaHostAllocDefault));
					XA_CUDA(cudaEventCreateWithFlags(&pl->ev_choice, cudaEventDisableTiming));
				}
				e = cudaMemcpyAsync(&pl->h_choice[b], c.d_choice, sizeof(uint32_t),
				    cudaMemcpyDeviceToHost, st);
				pl->choice_asked[b] = true;
				asked_now = true;
			}
		} else {
			EncodeParams p;
			p.src = (const uint8_t *)src;
			p.src_bytes = src_bytes;
			p.dst = (uint8_t *)dst;
			p.dst_bytes = dst_bytes;
			p.streams = pl->d_streams.p;
			p.tiles = pl->d_tiles.p + t0;
			p.n_tiles = t1 - t0;
			p.results = NULL;
			p.order = NULL;
			p.n_streams = 0;
			switch (b) {
			case 0: e = launch_encode<4, 1>(p, st); break;
			case 1: e = launch_encode<4, 2>(p, st); break;
			case 2: e = launch_encode<6, 1>(p, st); break;
			case 3: e = launch_encode<6, 2>(p, st); break;
			case 4: e = launch_encode<8, 1>(p, st); break;
			default: e = launch_encode<8, 2>(p, st); break;
			}
		}
		XA_CUDA(e);
		if (fork) {
			XA_CUDA(cudaEventRecord(pl->ev_done[b], st));
			XA_CUDA(cudaStreamWaitEvent(user_st, pl->ev_done[b], 0));
		}
	}
	st = user_st;
	if (asked_now) {
		XA_CUDA(cudaEventRecord(pl->ev_choice, st));
		pl->choice_pending = true;
	}
	pl->n_launched += tls_launched - launched0;
	pl->ran = true;
	pl->last_stream = st;
	pl->last_dst = (uint8_t *)dst;
	pl->last_src = (const uint8_t *)src;
	pl->last_src_bytes = src_bytes;
	return (0);
}

/*
 * Streams that met a bad profile (libbjxa.c:550): the state the reference leaves
 * behind is the one after effective block done - 1 -- the last two frames of that
 * block, read back from the PCM arena -- and, when it is the RIGHT block of the
 * pair that is bad, the left channel has already advanced through its block of
 * the failing pair (libbjxa.c:633-643).  One thread per such stream redoes that,
 * one launch and one copy for all of them.
 */
struct BadReq {
	uint64_t xa_pair;	/* arena address of the failing effective block */
	uint64_t pcm_tail;	/* arena address of the last two frames in front of it */
	int16_t  prev[2][2];	/* state on entry, used when no block precedes */
	uint32_t eb;		/* effective blocks completed */
	uint8_t  bits, channels, bad_ch, pad;
};

__global__ void
xa_badstate_kernel(const uint8_t *xa, const uint8_t *pcm, const BadReq *req, uint32_t n,
    StreamRes *out)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	const BadReq r = req[i];
	int st[2][2];
	for (int c = 0; c < 2; c++) {
		st[c][0] = r.prev[c][0];
		st[c][1] = r.prev[c][1];
	}
	if (r.eb > 0) {
		const int16_t *t = reinterpret_cast<const int16_t *>(pcm + r.pcm_tail);
		for (uint32_t c = 0; c < r.channels; c++) {
			st[c][1] = t[c];			/* frame 30 */
			st[c][0] = t[r.channels + c];		/* frame 31 */
		}
	}
	if (r.bad_ch == 1) {
		/* the left block of the failing pair, sample by sample */
		const uint8_t *b = xa + r.xa_pair;
		const uint32_t prof = b[0];
		const int k0 = gain_k0(prof >> 4), k1 = gain_k1(prof >> 4), sh = (int)(prof & 15u);
		for (int s = 0; s < 32; s++) {
			const uint32_t bit = (uint32_t)s * r.bits;
			/* code s, MSB first, as the top bits of an int16 (libbjxa.c:286-345) */
			const uint32_t two = (uint32_t)b[1 + (bit >> 3)] << 8 | b[2 + (bit >> 3)];
			const int16_t top = (int16_t)((two << (bit & 7u)) & (0xffffu << (16 - r.bits)) & 0xffffu);
			int p0 = st[0][0], p1 = st[0][1];
			(void)sample_chain((int)top << 16, 16 + sh, k0, k1, p0, p1);
			st[0][0] = p0;
			st[0][1] = p1;
		}
	}
	for (int c = 0; c < 2; c++) {
		out[i].prev[c][0] = (int16_t)st[c][0];
		out[i].prev[c][1] = (int16_t)st[c][1];
	}
}

static int
plan_fetch_decode(bjxa_plan *pl, bjxa_stream_desc_t *out, size_t n)
{
	std::vector<StreamRes> res(n);
	std::vector<uint32_t> bad(n);
	uint32_t fault = 0;
	XA_CUDA(cudaMemcpy(&fault, pl->d_fault.p, sizeof fault, cudaMemcpyDeviceToHost));
	if (fault != 0) {
		/* a tile gave up waiting for its predecessor's state: internal
		 * error, the output is not trustworthy */
		(void)cudaMemset(pl->d_fault.p, 0, sizeof fault);
		(void)cudaStreamSynchronize(0);
		errno = EIO;
		return (-1);
	}
	if (n) {
		XA_CUDA(cudaMemcpy(res.data(), pl->d_results.p, n * sizeof(StreamRes),
		    cudaMemcpyDeviceToHost));
		XA_CUDA(cudaMemcpy(bad.data(), pl->d_first_bad.p, n * sizeof(uint32_t),
		    cudaMemcpyDeviceToHost));
	}
	std::vector<BadReq> reqs;
	std::vector<size_t> who;
	for (size_t i = 0; i < n; i++) {
		const bjxa_stream_desc_t &d = pl->descs[i];
		out[i] = d;
		if (d.blocks == 0) {
			out[i].done = 0;
			out[i].result = 0;
			out[i].error = 0;
			continue;
		}
		if (bad[i] == 0xffffffffu) {
			memcpy(out[i].prev, res[i].prev, sizeof res[i].prev);
			out[i].done = d.blocks;
			out[i].result = (int32_t)d.blocks;
			out[i].error = 0;
			continue;
		}
		/* bad profile at block-channel bad[i] (libbjxa.c:550,634,642) */
		const uint32_t eb = bad[i] / d.channels, c = bad[i] % d.channels;
		out[i].done = eb;
		out[i].result = -1;
		out[i].error = EPROTO;
		if (eb == 0 && c == 0)
			continue;		/* nothing decoded: the state on entry stands */
		BadReq r;
		memset(&r, 0, sizeof r);
		r.xa_pair = d.xa_off + (uint64_t)eb * (uint64_t)(block_bytes(d.bits) * d.channels);
		r.pcm_tail = eb > 0 ? d.pcm_off + (uint64_t)(eb - 1) * 64u * d.channels + 60u * d.channels : 0;
		memcpy(r.prev, d.prev, sizeof r.prev);
		r.eb = eb;
		r.bits = d.bits;
		r.channels = d.channels;
		r.bad_ch = (uint8_t)c;
		reqs.push_back(r);
		who.push_back(i);
	}
	if (!reqs.empty()) {
		const size_t m = reqs.size();
		BadReq *d_req = NULL;
		StreamRes *d_out = NULL;
		std::vector<StreamRes> st(m);
		XA_CUDA(cudaMalloc((void **)&d_req, m * sizeof(BadReq)));
		cudaError_t e = cudaMalloc((void **)&d_out, m * sizeof(StreamRes));
		if (e == cudaSuccess)
			e = cudaMemcpy(d_req, reqs.data(), m * sizeof(BadReq), cudaMemcpyHostToDevice);
		if (e == cudaSuccess) {
			xa_badstate_kernel<<<(unsigned)((m + 127) / 128), 128>>>(pl->last_src,
			    pl->last_dst, d_req, (uint32_t)m, d_out);
			e = cudaGetLastError();
		}
		if (e == cudaSuccess)
			e = cudaMemcpy(st.data(), d_out, m * sizeof(StreamRes), cudaMemcpyDeviceToHost);
		(void)cudaFree(d_req);
		if (d_out)
			(void)cudaFree(d_out);
		XA_CUDA(e);
		for (size_t k = 0; k < m; k++)
			memcpy(out[who[k]].prev, st[k].prev, sizeof st[k].prev);
	}
	return (0);
}

extern "C" int
bjxa_plan_fetch(bjxa_plan_t *pl, bjxa_stream_desc_t *out, size_t n)
{
	CHECK_PLAN(pl);
	if (out == NULL && n != 0) {
		errno = EFAULT;
		return (-1);
	}
	if (!pl->ran || n != pl->descs.size()) {
		errno = EINVAL;
		return (-1);
	}
	XA_CUDA(cudaStreamSynchronize(pl->last_stream));
	XA_CUDA(cudaGetLastError());

	try {
		if (pl->hp.kind == kKindDecode)
			return (plan_fetch_decode(pl, out, n));
		std::vector<StreamRes> after;
		if (pl->hp.kind == kKindSearch && n) {
			/* the decoder state the encoded blocks leave behind */
			after.resize(n);
			XA_CUDA(cudaMemcpy(after.data(), pl->d_results.p, n * sizeof(StreamRes),
			    cudaMemcpyDeviceToHost));
		}
		for (size_t i = 0; i < n; i++) {
			out[i] = pl->descs[i];
			out[i].done = out[i].blocks;
			out[i].result = (int32_t)out[i].blocks;
			out[i].error = 0;
			if (!after.empty() && out[i].blocks != 0)
				memcpy(out[i].prev, after[i].prev, sizeof after[i].prev);
		}
	} catch (const std::bad_alloc &) {
		errno = ENOMEM;
		return (-1);
	}
	return (0);
}

/* ---- one small call: bjxa_decode / bjxa_encode on a few blocks --------------------- */
/*
 * The reference's CLI calls bjxa_decode once per block by default
 * (/root/reference/src/bjxa_decode.c:102-161): 20 672 calls for one of its test
 * vectors.  A call that small is not a batch: no plan, no tables, no census.
 * Its parameters travel as kernel arguments, its bytes through a pinned, mapped
 * staging buffer of the calling thread that the kernel reads and writes across
 * PCIe itself, and the host waits for one launch.  Used for calls of at most
 * kSmallBlocks effective blocks; everything larger takes the plan path.
 */
constexpr uint32_t kSmallBlocks = 32;

struct SmallRes {
	int16_t  prev[2][2];
	uint32_t done;
	int32_t  error;
};

struct SmallArgs {
	const uint8_t *in;	/* staged source bytes (16-byte aligned, slack behind) */
	uint8_t *out;		/* staged destination */
	SmallRes *res;
	uint32_t blocks, pcm_len;
	int16_t  prev[2][2];
};

/*
 * One warp; lane c < CH walks channel c block by block exactly as the reference
 * does (libbjxa.c:629-658): a bad profile stops the call at its effective block,
 * the left channel having advanced if it is the right block that is bad.
 */
template <int BITS, int CH>
__global__ void __launch_bounds__(32)
xa_small_decode_kernel(const SmallArgs a)
{
	constexpr int BS = block_bytes(BITS);
	__shared__ __align__(16) uint8_t in[kSmallBlocks * CH * BS + 48];
	__shared__ __align__(16) int16_t out[kSmallBlocks * CH * 32];
	const uint32_t lane = threadIdx.x;
	const uint32_t nin = a.blocks * CH * BS;
	for (uint32_t i = lane * 16u; i < nin; i += 512u)
		*reinterpret_cast<uint4 *>(in + i) = *reinterpret_cast<const uint4 *>(a.in + i);
	__syncwarp();
	int p0 = 0, p1 = 0;
	if (lane < CH) {
		p0 = a.prev[lane][0];
		p1 = a.prev[lane][1];
	}
	uint32_t done = a.blocks;
	int err = 0;
	for (uint32_t eb = 0; eb < a.blocks; eb++) {
		const uint8_t *b = in + (eb * CH + (lane < CH ? lane : 0u)) * BS;
		const uint32_t prof = b[0];
		const uint32_t badm = __ballot_sync(0xffffffffu, lane < CH && (prof >> 4) >= 5u);
		if (badm != 0) {
			done = eb;
			err = EPROTO;
			if (badm & 1u)
				break;		/* nothing of this effective block is decoded */
		}
		if (lane < CH && (badm == 0 || lane == 0)) {
			uint32_t pw[BITS];
#pragma unroll
			for (int i = 0; i < BITS; i++)
				pw[i] = (uint32_t)b[1 + 4 * i] | (uint32_t)b[2 + 4 * i] << 8 |
				    (uint32_t)b[3 + 4 * i] << 16 | (uint32_t)b[4 + 4 * i] << 24;
			const int k0 = gain_k0(prof >> 4), k1 = gain_k1(prof >> 4);
			const int sh = 16 + (int)(prof & 15u);
#pragma unroll
			for (int s = 0; s < 32; s++) {
				const int v = sample_chain(top_code<BITS>(pw, s), sh, k0, k1, p0, p1);
				if (badm == 0)
					out[(eb * 32u + (uint32_t)s) * CH + lane] = (int16_t)v;
			}
		}
		if (badm != 0)
			break;
	}
	__syncwarp();
	const uint32_t bytes = done == a.blocks ? a.pcm_len : done * 64u * CH;
	for (uint32_t i = lane * 16u; i + 16u <= bytes; i += 512u)
		*reinterpret_cast<uint4 *>(a.out + i) = *reinterpret_cast<const uint4 *>(
		    reinterpret_cast<const uint8_t *>(out) + i);
	for (uint32_t i = (bytes & ~15u) + lane * 2u; i < bytes; i += 64u)
		*reinterpret_cast<uint16_t *>(a.out + i) = *reinterpret_cast<const uint16_t *>(
		    reinterpret_cast<const uint8_t *>(out) + i);
	if (lane < CH) {
		a.res->prev[lane][0] = (int16_t)p0;
		a.res->prev[lane][1] = (int16_t)p1;
	}
	if (lane == 0) {
		a.res->done = done;
		a.res->error = err;
	}
}

/* one thread per block-channel: gather, zero-pad, profile 0, top bits (libbjxa.c:665-691) */
template <int BITS, int CH>
__global__ void __launch_bounds__(64)
xa_small_encode_kernel(const SmallArgs a)
{
	constexpr int BS = block_bytes(BITS);
	const uint32_t q = threadIdx.x;
	if (q < a.blocks * CH) {
		const uint32_t eb = q / CH, c = q % CH;
		const int16_t *pcm = reinterpret_cast<const int16_t *>(a.in);
		const uint32_t frames = a.pcm_len / (2u * CH);
		uint32_t w[16], pw[BITS];
#pragma unroll
		for (int i = 0; i < 16; i++) {
			const uint32_t f0 = eb * 32u + 2u * i, f1 = f0 + 1u;
			const uint32_t s0 = f0 < frames ? (uint16_t)pcm[f0 * CH + c] : 0u;
			const uint32_t s1 = f1 < frames ? (uint16_t)pcm[f1 * CH + c] : 0u;
			w[i] = s0 | s1 << 16;
		}
		deflate_block<BITS>(pw, w);
		uint8_t *o = a.out + q * BS;
		o[0] = 0;
#pragma unroll
		for (int i = 0; i < 4 * BITS; i++)
			o[1 + i] = (uint8_t)(pw[i >> 2] >> (8 * (i & 3)));
	}
	if (q == 0) {
		a.res->done = a.blocks;
		a.res->error = 0;
	}
}

/* the calling thread's staging for small calls, on the device it was made for */
struct SmallCtx {
	int dev;
	uint8_t *h_in, *h_out;		/* pinned + mapped; the device sees them as they are */
	SmallRes *h_res;
	cudaStream_t st;
};
static thread_local SmallCtx tls_small = { -1, NULL, NULL, NULL, NULL };
constexpr size_t kSmallIn = kSmallBlocks * 2 * 64 + 64, kSmallOut = kSmallBlocks * 2 * 64 + 64;

static void
small_release(void)
{
	SmallCtx &c = tls_small;
	if (c.dev < 0)
		return;
	int cur = -1;
	if (cudaGetDevice(&cur) == cudaSuccess && cur != c.dev)
		(void)cudaSetDevice(c.dev);
	if (c.st)
		(void)cudaStreamDestroy(c.st);
	if (c.h_in)
		(void)cudaFreeHost(c.h_in);
	if (c.h_out)
		(void)cudaFreeHost(c.h_out);
	if (c.h_res)
		(void)cudaFreeHost(c.h_res);
	if (cur >= 0 && cur != c.dev)
		(void)cudaSetDevice(cur);
	c = SmallCtx{ -1, NULL, NULL, NULL, NULL };
}

extern "C" void
bjxa_small_release(void)
{
	small_release();
}

static int
small_ready(void)
{
	SmallCtx &c = tls_small;
	int dev = 0;
	XA_CUDA(cudaGetDevice(&dev));
	if (c.dev == dev)
		return (0);
	small_release();
	cudaError_t e = cudaHostAlloc((void **)&c.h_in, kSmallIn, cudaHostAllocMapped);
	if (e == cudaSuccess)
		e = cudaHostAlloc((void **)&c.h_out, kSmallOut, cudaHostAllocMapped);
	if (e == cudaSuccess)
		e = cudaHostAlloc((void **)&c.h_res, sizeof(SmallRes), cudaHostAllocMapped);
	if (e == cudaSuccess)
		e = cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking);
	c.dev = dev;
	if (e != cudaSuccess) {
		(void)cudaGetLastError();
		small_release();
		errno = cuda_errno(e);
		return (-1);
	}
	return (0);
}

template <int BITS, int CH>
static cudaError_t
small_launch(int kind, const SmallArgs &a, cudaStream_t st)
{
	if (kind == kKindDecode)
		xa_small_decode_kernel<BITS, CH><<<1, 32, 0, st>>>(a);
	else
		xa_small_encode_kernel<BITS, CH><<<1, 64, 0, st>>>(a);
	return cudaGetLastError();
}

/*
 * One stream, d->blocks <= kSmallBlocks effective blocks, host buffers: decode
 * (src = XA, dst = PCM) or reference-exact encode (the other way round).  Fills
 * d->done / result / error / prev like bjxa_plan_fetch.  Returns 1 when the call
 * is not a small one (the caller takes the plan path), 0 or -1 otherwise.
 */
extern "C" int
bjxa_small_call(int kind, bjxa_stream_desc_t *d, void *dst, const void *src)
{
	if (d->blocks == 0 || d->blocks > kSmallBlocks ||
	    (kind != BJXA_PLAN_DECODE && kind != BJXA_PLAN_ENCODE))
		return (1);
	if (small_ready() < 0)
		return (-1);
	SmallCtx &c = tls_small;
	const size_t xa_bytes = (size_t)d->blocks * d->channels * (size_t)block_bytes(d->bits);
	const size_t in_bytes = kind == BJXA_PLAN_DECODE ? xa_bytes : d->pcm_len;
	memcpy(c.h_in, src, in_bytes);
	SmallArgs a;
	a.in = c.h_in;
	a.out = c.h_out;
	a.res = c.h_res;
	a.blocks = d->blocks;
	a.pcm_len = d->pcm_len;
	memcpy(a.prev, d->prev, sizeof a.prev);
	cudaError_t e;
	switch (bucket_of(d->bits, d->channels)) {
	case 0: e = small_launch<4, 1>(kind, a, c.st); break;
	case 1: e = small_launch<4, 2>(kind, a, c.st); break;
	case 2: e = small_launch<6, 1>(kind, a, c.st); break;
	case 3: e = small_launch<6, 2>(kind, a, c.st); break;
	case 4: e = small_launch<8, 1>(kind, a, c.st); break;
	default: e = small_launch<8, 2>(kind, a, c.st); break;
	}
	XA_CUDA(e);
	XA_CUDA(cudaStreamSynchronize(c.st));
	d->done = c.h_res->done;
	d->error = c.h_res->error;
	d->result = d->error != 0 ? -1 : (int32_t)d->done;
	size_t out_bytes;
	if (kind == BJXA_PLAN_DECODE) {
		memcpy(d->prev, c.h_res->prev, sizeof d->prev);
		out_bytes = d->done == d->blocks ? d->pcm_len : (size_t)d->done * 64u * d->channels;
	} else {
		out_bytes = xa_bytes;
	}
	memcpy(dst, c.h_out, out_bytes);
	return (0);
}

/* ---- per-stream checksums of a run's output ---------------------------------- */

/*
 * sum over i of word_i * ((i * G + C) | 1) mod 2^64, word_i = bytes 4i..4i+3 of
 * the range, little endian, absent bytes zero (include/bjxa_batch.h).  A plain
 * sum of position-keyed terms: any split of the words over threads gives the
 * same value.
 */
constexpr unsigned long long kSumG = 0x9E3779B97F4A7C15ULL, kSumC = 0xD1B54A32D192ED03ULL;
constexpr int kSumThreads = 256;

struct SumRange {
	uint64_t off, len;
};

__device__ __forceinline__ unsigned long long
sum_term(uint32_t w, unsigned long long i)
{
	return (unsigned long long)w * ((i * kSumG + kSumC) | 1ULL);
}

/* words [w0, w1) of the range at arena + off, `len` bytes long; `lim` = arena bytes */
__device__ unsigned long long
sum_words(const uint8_t *arena, uint64_t lim, uint64_t off, uint64_t len,
    unsigned long long w0, unsigned long long w1, uint32_t tid, uint32_t nt)
{
	unsigned long long acc = 0;
	const uint8_t *base = arena + off;
	if ((((uintptr_t)base) & 15u) == 0 && (w0 & 3u) == 0) {
		/* aligned: 16 bytes a step, the ragged end below */
		const unsigned long long whole = len / 16u * 4u;	/* words in whole vectors */
		const unsigned long long v1 = (w1 < whole ? w1 : whole) & ~3ULL;
		for (unsigned long long i = w0 + 4ULL * tid; i < v1; i += 4ULL * nt) {
			const uint4 v = *reinterpret_cast<const uint4 *>(base + i * 4u);
			acc += sum_term(v.x, i) + sum_term(v.y, i + 1) + sum_term(v.z, i + 2) +
			    sum_term(v.w, i + 3);
		}
		w0 = v1 > w0 ? v1 : w0;
	}
	for (unsigned long long i = w0 + tid; i < w1; i += nt) {
		uint32_t w = 0;
		for (uint32_t b = 0; b < 4u; b++) {
			const uint64_t at = i * 4u + b;
			if (at < len && off + at < lim)
				w |= (uint32_t)base[at] << (8u * b);
		}
		acc += sum_term(w, i);
	}
	return acc;
}

/* one CTA per range (grid-stride), or -- few, long ranges -- every CTA a slice of each */
__global__ void __launch_bounds__(kSumThreads)
xa_checksum_kernel(const uint8_t *arena, uint64_t lim, const SumRange *ranges, uint32_t n,
    unsigned long long *sums, int spread)
{
	__shared__ unsigned long long part[kSumThreads / 32];
	const uint32_t tid = threadIdx.x;
	for (uint32_t r = spread ? 0 : blockIdx.x; r < n; r += spread ? 1 : gridDim.x) {
		const SumRange g = ranges[r];
		const unsigned long long words = (g.len + 3u) / 4u;
		unsigned long long w0 = 0, w1 = words;
		if (spread) {
			/* slices of whole 16-byte vectors */
			const unsigned long long per = ((words + gridDim.x - 1) / gridDim.x + 3ULL) & ~3ULL;
			w0 = per * blockIdx.x;
			w1 = w0 + per < words ? w0 + per : words;
			if (w0 >= words)
				continue;
		}
		unsigned long long acc = sum_words(arena, lim, g.off, g.len, w0, w1, tid, kSumThreads);
#pragma unroll
		for (int o = 16; o > 0; o >>= 1)
			acc += __shfl_xor_sync(0xffffffffu, acc, o);
		__syncthreads();
		if ((tid & 31u) == 0)
			part[tid >> 5] = acc;
		__syncthreads();
		if (tid == 0) {
			unsigned long long t = 0;
			for (int w = 0; w < kSumThreads / 32; w++)
				t += part[w];
			if (spread)
				atomicAdd(&sums[r], t);
			else
				sums[r] = t;
		}
	}
}

extern "C" int
bjxa_plan_checksum(bjxa_plan_t *pl, uint64_t *sums, size_t n)
{
	CHECK_PLAN(pl);
	if (sums == NULL && n != 0) {
		errno = EFAULT;
		return (-1);
	}
	if (!pl->ran || n != pl->descs.size() || n > 0xffffffffu) {
		errno = EINVAL;
		return (-1);
	}
	if (n == 0)
		return (0);
	try {
		std::vector<SumRange> r(n);
		const bool pcm = pl->hp.kind == kKindDecode;
		for (size_t i = 0; i < n; i++) {
			const bjxa_stream_desc_t &d = pl->descs[i];
			r[i].off = pcm ? d.pcm_off : d.xa_off;
			r[i].len = d.blocks == 0 ? 0 : pcm ? (uint64_t)d.pcm_len :
			    (uint64_t)d.blocks * (uint64_t)(block_bytes(d.bits) * d.channels);
		}
		const uint64_t lim = pl->hp.dst_need;	/* bytes of the arena the run wrote */
		SumRange *d_r = NULL;
		unsigned long long *d_s = NULL;
		XA_CUDA(cudaStreamSynchronize(pl->last_stream));
		XA_CUDA(cudaMalloc((void **)&d_r, n * sizeof(SumRange)));
		cudaError_t e = cudaMalloc((void **)&d_s, n * sizeof(unsigned long long));
		if (e == cudaSuccess)
			e = cudaMemcpy(d_r, r.data(), n * sizeof(SumRange), cudaMemcpyHostToDevice);
		if (e == cudaSuccess)
			e = cudaMemset(d_s, 0, n * sizeof(unsigned long long));
		if (e == cudaSuccess) {
			int dev = 0, sms = 148;
			(void)cudaGetDevice(&dev);
			(void)cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
			const uint32_t grid = (uint32_t)sms * 8u;
			const int spread = n < 2048;
			xa_checksum_kernel<<<spread ? grid : (n < grid ? (uint32_t)n : grid), kSumThreads>>>(
			    pl->last_dst, lim, d_r, (uint32_t)n, d_s, spread);
			e = cudaGetLastError();
		}
		if (e == cudaSuccess)
			e = cudaMemcpy(sums, d_s, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
		(void)cudaFree(d_r);
		if (d_s)
			(void)cudaFree(d_s);
		XA_CUDA(e);
	} catch (const std::bad_alloc &) {
		errno = ENOMEM;
		return (-1);
	}
	return (0);
}

/* ---- sharding ------------------------------------------------------------- */

/* ---- header scatter (file assembly on the device) --------------------------- */

__global__ void
xa_scatter_kernel(uint8_t *dst, const uint8_t *table, uint32_t rec_len, uint32_t n)
{
	/* one 64-byte record per 64 threads: thread b moves byte b */
	const uint32_t i = blockIdx.x * (blockDim.x / 64u) + threadIdx.x / 64u;
	const uint32_t b = threadIdx.x % 64u;
	if (i >= n || b >= rec_len)
		return;
	const uint8_t *rec = table + (uint64_t)i * 64u;
	const uint64_t off = *reinterpret_cast<const uint64_t *>(rec);
	dst[off + b] = rec[8 + b];
}

extern "C" int
bjxa_gpu_scatter_async(void *dst, const void *d_table, uint32_t rec_len, size_t n,
    void *cuda_stream)
{
	if (n == 0)
		return (0);
	if (dst == NULL || d_table == NULL) {
		errno = EFAULT;
		return (-1);
	}
	if (rec_len > 56 || n > 0xffffffffu) {
		errno = EINVAL;
		return (-1);
	}
	const uint32_t per = 256 / 64;
	xa_scatter_kernel<<<(unsigned)((n + per - 1) / per), 256, 0, (cudaStream_t)cuda_stream>>>(
	    (uint8_t *)dst, (const uint8_t *)d_table, rec_len, (uint32_t)n);
	XA_CUDA(cudaGetLastError());
	return (0);
}

extern "C" int
bjxa_shard_range(const uint64_t *bytes, size_t n, int rank, int world,
    size_t *first, size_t *count)
{
	if (first == NULL || count == NULL) {
		errno = EFAULT;
		return (-1);
	}
	if (world <= 0 || rank < 0 || rank >= world) {
		errno = EINVAL;
		return (-1);
	}
	size_t lo, hi;
	if (bytes == NULL) {
		lo = n * (size_t)rank / (size_t)world;
		hi = n * ((size_t)rank + 1) / (size_t)world;
	} else {
		/* boundary k sits where the running total first reaches k/world */
		long double total = 0;
		for (size_t i = 0; i < n; i++)
			total += (long double)bytes[i];
		long double tlo = total * rank / world, thi = total * (rank + 1) / world;
		long double run = 0;
		lo = hi = n;
		bool got_lo = false, got_hi = false;
		for (size_t i = 0; i <= n; i++) {
			if (!got_lo && run >= tlo) { lo = i; got_lo = true; }
			if (!got_hi && run >= thi) { hi = i; got_hi = true; }
			if (i < n)
				run += (long double)bytes[i];
		}
		if (rank == world - 1)
			hi = n;
		if (rank == 0)
			lo = 0;
	}
	*first = lo;
	*count = hi - lo;
	return (0);
}
