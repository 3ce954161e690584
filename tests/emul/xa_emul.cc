/*
 * xa_emul.cc -- CPU single-stepper of the tile code.  TEST HARNESS ONLY.
 *
 * Compiles bjxa_b200/csrc/xa_tile.h + xa_plan.h (the exact phase code the
 * sm_100a kernels run) with a host compiler and executes a batch the way the
 * GPU would: tiles in ticket order, every phase as a loop over the CTA's
 * threads with the barriers where xa_kernels.cu has them.  It exists so that
 * indexing / scheduling logic can be checked against the oracle on a machine
 * without a GPU (tests/test_emul.py).  It is never linked into the product
 * library and nothing in bjxa_b200/ refers to it.
 */
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/bjxa_batch.h"
#include "../../bjxa_b200/csrc/xa_plan.h"
#include "../../bjxa_b200/csrc/xa_walk.h"

using namespace xa;

static int g_use_alt = 0;
static int g_stereo_direct = 0;
static int g_pool = 0;
static int g_split = 0;
static int g_relay = 0;
static int g_seg = 0;

/* thread visiting order inside a phase: 0 ascending, 1 descending, 2 strided */
static uint32_t
visit(uint32_t i, uint32_t nt, int order)
{
	if (order == 1)
		return nt - 1 - i;
	if (order == 2)
		return (i * 37u + 11u) % nt;	/* 37 coprime with 128/256 */
	return i;
}

template <class Tile>
static void
emul_decode_ns(const DecodeParams &p, int order)
{
	typedef typename Tile::G G;
	typename Tile::Smem *sm = new typename Tile::Smem();
	const uint32_t nt = kDecThreads;
	std::vector<LiveRec> live;
	/* relay form: the records of every tile, kRelayPerTile slots each */
	std::vector<RelayRec> relay_recs((size_t)p.n_tiles * kRelayPerTile + 1);
	uint32_t relay_count = 0xffffffffu;
	DecodeParams pr = p;
	pr.relay_recs = relay_recs.data();
	pr.relay_count = &relay_count;
	constexpr bool relay = Tile::relay;	/* relay tiles: DecTile<..., true> */

	memset(sm, 0xa5, sizeof *sm);	/* smem is garbage at CTA start */
	for (uint32_t ticket = 0; ticket < p.n_tiles; ticket++) {
		/* one persistent CTA draws every ticket; stages rotate as on the GPU */
		const int s = (int)(ticket % Tile::kStages);
		memset(sm->in[s], 0xa5, sizeof sm->in[s]);
		const TileEnt te = p.tiles[ticket];
		bool tail = false;
		for (uint32_t lane = 0; lane < te.count; lane++) {
			StripCtx &c = sm->ctx[s][lane];
			make_strip_ctx<G::kBits, G::kCh, G::kTBQ, G::kNS>(c, p,
			    p.order[te.first + lane], te.j, lane, g_split != 0 || relay);
			memcpy(sm->in[s] + lane * Tile::G::SLOT, p.src + c.a0, c.bulk);
			tail |= (c.flags & kCtxTail) != 0;
		}
		sm->n_strips[s] = te.count;
		sm->tile_flags[s] = tail ? kCtxTail : 0u;
		Tile t(relay ? pr : p, *sm, s);
		/* producer: tail bytes, then the scan for heads of chains */
		if (tail)
			for (uint32_t i = 0; i < 32; i++)
				t.load_tail(i, 32, sm->in[s]);
		uint32_t count = 0;
		for (uint32_t q = 0; q < te.count * Tile::SCAN; q++)
			if (t.is_head(q)) {
				/* split form: item 0 is no head if it goes on with a chain of
				 * the strip in front (scanner_warp: `prev` from the context) */
				if ((g_split || relay) && G::kNS == 1 && q == 0 &&
				    (t.chain_mask(0) & (sm->ctx[s][0].flags >> kCtxPrevShift & 3u)) != 0)
					continue;
				sm->heads[s][count++] = (uint16_t)q;
			}
		sm->n_heads[s] = count;
		sm->next_head[s] = 32;
		if (g_split && G::kNS == 1) {
			/* split form, pass 1 (scanner_warp with `split`): the heads go out
			 * as one record per tile, the consumers decode units only */
			if (count != 0) {
				LiveRec r;
				memset(&r, 0, sizeof r);
				for (uint32_t i = 0; i < count; i++)
					r.heads[sm->heads[s][i] >> 5] |= 1u << (sm->heads[s][i] & 31u);
				const StripCtx &c = sm->ctx[s][0];
				const uint64_t g0 = c.a0 + c.in_base;
				r.xa_lo = (uint32_t)g0;
				r.xa_hi = (uint32_t)(g0 >> 32);
				r.out_lo = (uint32_t)c.out0;
				r.out_hi = (uint32_t)(c.out0 >> 32);
				r.stream = c.stream;
				r.first_eb = c.first_eb;
				r.blocks = c.blocks;
				live.push_back(r);
			}
			for (uint32_t i = 0; i < nt; i++)
				t.phase_units(visit(i, nt, order), nt);
			continue;
		}
		if (relay) {
			/* relay form (relay_walk_warp): 32 lanes in lockstep, a block per
			 * turn; once the heads are dealt out and few lanes still walk,
			 * their chains are handed to the second pass */
			const uint32_t nw = count > kRelayManyHeads ? kRelayWalkers : 1u;
			const uint32_t nl = 32 * nw;		/* walker lanes; a warp = 32 of them */
			std::vector<typename Tile::Walk> w(nl);
			std::vector<char> have(nl, 0), gone(nw, 0);
			std::vector<uint32_t> idx(nl);
			uint32_t next = nl;
			for (uint32_t l = 0; l < nl; l++)
				idx[l] = l;
			for (uint32_t turn = 0;; turn++) {
				uint32_t busy = 0;
				for (uint32_t k = 0; k < nw; k++) {
					if (gone[k])
						continue;
					uint32_t active = 0;
					for (uint32_t l = 32 * k; l < 32 * k + 32; l++) {
						if (!have[l] && idx[l] < count) {
							t.walk_begin(w[l], sm->heads[s][idx[l]]);
							have[l] = 1;
						}
						active += have[l];
					}
					if (active == 0) {
						gone[k] = 1;
						continue;
					}
					if (turn >= 4 && active <= kRelayWind && next >= count) {
						for (uint32_t l = 32 * k; l < 32 * k + 32; l++)
							if (have[l])
								t.walk_hand_on(w[l]);
						gone[k] = 1;
						continue;
					}
					busy++;
					for (uint32_t i = 0; i < 32; i++) {
						const uint32_t l = 32 * k + visit(i, 32, order);
						if (have[l] && !t.walk_block(w[l])) {
							have[l] = 0;
							idx[l] = next++;
						}
					}
				}
				if (busy == 0)
					break;
			}
			const uint32_t ut = count == 0 ? nt : nt - 32 * nw;
			for (uint32_t i = 0; i < ut; i++)
				t.phase_units(visit(i, ut, order), ut);
			continue;
		}
		if (g_pool) {
			/* pooled form (xa_decode_pool_kernel): 32 walker lanes, each
			 * drawing a chain when idle and decoding ONE block per turn; the
			 * units belong to two warps of their own */
			typename Tile::Walk w[32];
			bool have[32] = { false };
			uint32_t next = 0, live = 0;
			do {
				live = 0;
				for (uint32_t i = 0; i < 32; i++) {
					const uint32_t l = visit(i, 32, order);
					if (!have[l] && next < count) {
						t.walk_begin(w[l], sm->heads[s][next++]);
						if (w[l].need != 0)
							return;		/* carries are there, in ticket order */
						have[l] = true;
					} else if (have[l]) {
						have[l] = t.walk_block(w[l]);
					}
					live += have[l];
				}
			} while (live != 0 || next < count);
			for (uint32_t i = 0; i < 64; i++)
				t.phase_units(visit(i, 64, order), 64);
			continue;
		}
		/* consumers: the tile's walker warp, lane by lane (whichever lane runs
		 * first picks up every chain left over), then the units */
		for (uint32_t i = 0; i < 32; i++)
			t.phase_walk_warp(visit(i, 32, order), sm->heads[s], count, &sm->next_head[s]);
		/* as in consume_tile: a tile with chains leaves its units to the
		 * seven warps that do not walk */
		const uint32_t ut = count == 0 ? nt : nt - 32;
		for (uint32_t i = 0; i < ut; i++)
			t.phase_units(visit(i, ut, order), ut);
	}
	/* relay form, pass 2: every record, in any order */
	for (uint32_t i = 0; i < relay_count + 1u; i++)
		walk_relay_serial<G::kBits, G::kCh>(p, relay_recs[visit(i, relay_count + 1u,
		    order == 2 ? 1 : order)]);
	/* split form, pass 2 (xa_walk_kernel): the records in any order, a chain at a time */
	for (size_t i = 0; i < live.size(); i++)
		walk_record_serial<G::kBits, G::kCh>(p, live[visit((uint32_t)i, (uint32_t)live.size(),
		    order == 2 ? 1 : order)]);
	delete sm;
}

template <int BITS, int CH, int NS>
static void
emul_decode_staged_ns(const DecodeParams &p, int order)
{
	typedef DecTileStaged<BITS, CH, kDecTBQ, NS, 2> Tile;
	typename Tile::Smem *sm = new typename Tile::Smem();
	const uint32_t nt = kDecThreads;

	memset(sm, 0xa5, sizeof *sm);	/* smem is garbage at CTA start */
	for (uint32_t ticket = 0; ticket < p.n_tiles; ticket++) {
		const int s = (int)(ticket % 2);
		memset(sm->in[s], 0xa5, sizeof sm->in[s]);
		const TileEnt te = p.tiles[ticket];
		bool tail = false;
		for (uint32_t lane = 0; lane < te.count; lane++) {
			StripCtx &c = sm->ctx[s][lane];
			make_strip_ctx<BITS, CH, kDecTBQ, NS>(c, p, p.order[te.first + lane],
			    te.j, lane);
			memcpy(sm->in[s] + lane * Tile::G::SLOT, p.src + c.a0, c.bulk);
			tail |= (c.flags & kCtxTail) != 0;
		}
		sm->n_strips[s] = te.count;
		sm->tile_flags[s] = tail ? kCtxTail : 0u;
		Tile t(p, *sm, s);
		if (tail)
			for (uint32_t i = 0; i < 32; i++)
				t.load_tail(i, 32, sm->in[s]);
		uint32_t count = 0;
		for (uint32_t q = 0; q < te.count * Tile::SCAN; q++)
			if (t.is_head(q))
				sm->heads[s][count++] = (uint16_t)q;
		sm->n_heads[s] = count;
		for (uint32_t i = 0; i < nt; i++)
			t.phase_walk(visit(i, nt, order), nt, sm->heads[s], count, ticket * 96u);
		for (uint32_t i = 0; i < nt; i++)
			t.phase_a(visit(i, nt, order), nt);
		for (uint32_t i = 0; i < nt; i++)
			t.phase_store(visit(i, nt, order), nt);
	}
	delete sm;
}

template <int BITS, int CH>
static void
emul_decode_bucket(const DecodeParams &p, int ns, int order)
{
	/* same choice as launch_decode_ns: mono direct; stereo staged, or direct
	 * when g_stereo_direct is set (the tests run both) */
	if (CH == 2 && !g_stereo_direct && !((g_split || g_relay) && ns == 1)) {
		if (ns == 1)
			emul_decode_staged_ns<BITS, CH, 1>(p, order);
		else
			emul_decode_staged_ns<BITS, CH, kDecWide>(p, order);
	} else if (CH == 2) {
		if (ns == 1 && g_relay)
			emul_decode_ns<DecTileStereo<BITS, kDecTBQ, 1, dec_stages(BITS, 2), true> >(p, order);
		else if (ns == 1)
			emul_decode_ns<DecTileStereo<BITS, kDecTBQ, 1, dec_stages(BITS, 2)> >(p, order);
		else
			emul_decode_ns<DecTileStereo<BITS, kDecTBQ, kDecWide, dec_stages(BITS, 2)> >(p, order);
	} else if (ns == 1 && g_relay) {
		emul_decode_ns<DecTile<BITS, kDecTBQ, 1, dec_stages(BITS, 1), true> >(p, order);
	} else if (ns == 1) {
		emul_decode_ns<DecTile<BITS, kDecTBQ, 1, dec_stages(BITS, 1)> >(p, order);
	} else {
		emul_decode_ns<DecTile<BITS, kDecTBQ, kDecWide, dec_stages(BITS, 1)> >(p, order);
	}
}

template <int BITS, int CH>
static void
emul_encode_bucket(const EncodeParams &p, int order)
{
	typedef EncTile<BITS, CH, kEncTBE> Tile;
	typename Tile::Smem *sm = new typename Tile::Smem();
	const uint32_t nt = kEncThreads;

	for (uint32_t tile = 0; tile < p.n_tiles; tile++) {
		memset(sm, 0xa5, sizeof *sm);
		Tile t(p, *sm, tile);
		memcpy(sm->in, p.src + t.in0, t.bulk_bytes());
		for (uint32_t i = 0; i < nt; i++)
			t.load_tail(visit(i, nt, order), nt);
		for (uint32_t i = 0; i < nt; i++)
			t.phase_pack(visit(i, nt, order), nt);
		for (uint32_t i = 0; i < nt; i++)
			t.phase_store(visit(i, nt, order), nt);
	}
	delete sm;
}

/*
 * Searching encoder: the candidate arithmetic of xa_core.h (search_sample,
 * put_code) driven by plain loops -- candidates in ascending order, strict
 * "<", which is the key order (error, candidate) of the kernel's warp-wide
 * argmin.  prev_out: [n][2][2] decoder state after the last block.
 */
template <int BITS>
static void
emul_search_stream(const bjxa_stream_desc_t &d, const uint8_t *src, uint8_t *dst,
    int16_t *prev_out)
{
	constexpr int NR = search_ranges(BITS), NC = search_candidates(BITS);
	constexpr int BS = block_bytes(BITS);
	const uint32_t ch_n = d.channels, frames = d.pcm_len / (2u * ch_n);
	const int16_t *pcm = reinterpret_cast<const int16_t *>(src + d.pcm_off);
	for (uint32_t ch = 0; ch < ch_n; ch++) {
		/* state and samples biased by +32768, codes by 2^(BITS-1), as in the kernel */
		int s0 = d.prev[ch][0] + 32768, s1 = d.prev[ch][1] + 32768;
		uint32_t unbias[BITS];
		search_code_bias<BITS>(unbias);
		for (uint32_t eb = 0; eb < d.blocks; eb++) {
			int x[32];
			for (uint32_t i = 0; i < 32; i++) {
				uint32_t fr = eb * 32 + i;
				x[i] = 32768 + (fr < frames ? pcm[(uint64_t)fr * ch_n + ch] : 0);
			}
			unsigned long long best = ~0ULL;
			uint32_t bw[BITS] = { 0 };
			int bq0 = 0, bq1 = 0, bc = 0;
			for (int c = 0; c < NC; c++) {
				SearchK<BITS> K;
				search_setup<BITS>(K, (unsigned)(c / NR), 16 - BITS - c % NR);
				int q0 = s0, q1 = s1;
				unsigned long long err = 0;
				uint32_t w[BITS] = { 0 };
				for (int i = 0; i < 32; i++)
					put_code<BITS>(w, i, search_sample_b<BITS>(x[i], K, q0, q1, err));
				if (err < best) {
					best = err;
					bc = c;
					bq0 = q0;
					bq1 = q1;
					memcpy(bw, w, sizeof w);
				}
			}
			uint8_t *blk = dst + d.xa_off + ((uint64_t)eb * ch_n + ch) * BS;
			blk[0] = (uint8_t)((bc / NR) << 4 | (bc % NR));
			for (int j = 0; j < 4 * BITS; j++)
				blk[1 + j] = (uint8_t)((bw[j >> 2] ^ unbias[j >> 2]) >> (8 * (j & 3)));
			s0 = bq0;
			s1 = bq1;
		}
		prev_out[ch * 2] = (int16_t)(s0 - 32768);
		prev_out[ch * 2 + 1] = (int16_t)(s1 - 32768);
	}
}

extern "C" {

/*
 * Runs a decode batch.  prev_out: n * 4 int16 (final states as the kernel
 * publishes them), first_bad: n uint32 (0xffffffff = none).  Returns 0 or an
 * errno value from plan validation.
 */
int
xa_emul_decode(const bjxa_stream_desc_t *descs, size_t n, const uint8_t *src,
    uint64_t src_bytes, uint8_t *dst, int16_t *prev_out, uint32_t *first_bad,
    int order, int force_strips)
{
	HostPlan hp;
	size_t bad = 0;
	int rc = build_plan(hp, kKindDecode, descs, n, &bad, force_strips, g_seg ? 0 : kSegMinItems);
	if (rc)
		return rc;
	std::vector<StreamRes> res(n);
	std::vector<unsigned long long> carry((size_t)hp.n_slots * 2 + 2, 0ULL);
	unsigned long long ticket = 0;
	uint32_t fault = 0;
	for (size_t i = 0; i < n; i++) {
		first_bad[i] = 0xffffffffu;
		memset(&res[i], 0x5a, sizeof res[i]);
	}
	for (int b = 0; b < 6; b++) {
		uint32_t t0 = hp.tile_begin[b], t1 = hp.tile_begin[b + 1];
		if (t0 == t1)
			continue;
		int ns = hp.ns[b];
		if (g_use_alt && hp.alt_ns[b] != 0) {
			/* the class's list in the other tile shape, as the census would pick */
			t0 = hp.alt_begin[b];
			t1 = hp.alt_begin[b + 1];
			ns = hp.alt_ns[b];
		}
		const bool seg = g_seg && hp.seg_begin[b + 1] > hp.seg_begin[b];
		if (seg) {
			/* the segment form's list */
			t0 = hp.seg_begin[b];
			t1 = hp.seg_begin[b + 1];
		}
		DecodeParams p;
		p.src = src;
		p.src_bytes = src_bytes;
		p.dst = dst;
		p.streams = hp.streams.data();
		p.results = res.data();
		p.first_bad = first_bad;
		p.tiles = hp.tiles.data() + t0;
		p.n_tiles = t1 - t0;
		p.order = hp.order.data();
		p.carry = carry.data();
		p.ticket = &ticket;
		p.fault = &fault;
		p.carry_timeout_ns = 0;
		p.epoch = 7;
		if (seg) {
			/* xa_seg_kernel: tiles in ticket order, the lanes of a pass in any */
			for (uint32_t t = 0; t < p.n_tiles; t++) {
				auto v = [&](uint32_t i) { return visit(i, 32, order); };
				switch (b) {
				case 0: walk_seg_tile_serial<4, 1>(p, p.tiles[t], v); break;
				case 1: walk_seg_tile_serial<4, 2>(p, p.tiles[t], v); break;
				case 2: walk_seg_tile_serial<6, 1>(p, p.tiles[t], v); break;
				case 3: walk_seg_tile_serial<6, 2>(p, p.tiles[t], v); break;
				case 4: walk_seg_tile_serial<8, 1>(p, p.tiles[t], v); break;
				default: walk_seg_tile_serial<8, 2>(p, p.tiles[t], v); break;
				}
			}
			continue;
		}
		switch (b) {
		case 0: emul_decode_bucket<4, 1>(p, ns, order); break;
		case 1: emul_decode_bucket<4, 2>(p, ns, order); break;
		case 2: emul_decode_bucket<6, 1>(p, ns, order); break;
		case 3: emul_decode_bucket<6, 2>(p, ns, order); break;
		case 4: emul_decode_bucket<8, 1>(p, ns, order); break;
		default: emul_decode_bucket<8, 2>(p, ns, order); break;
		}
	}
	memcpy(prev_out, res.data(), n * sizeof(StreamRes));
	return 0;
}

int
xa_emul_encode(const bjxa_stream_desc_t *descs, size_t n, const uint8_t *src,
    uint64_t src_bytes, uint8_t *dst, uint64_t dst_bytes, int order)
{
	HostPlan hp;
	size_t bad = 0;
	int rc = build_plan(hp, kKindEncode, descs, n, &bad);
	if (rc)
		return rc;
	for (int b = 0; b < 6; b++) {
		uint32_t t0 = hp.tile_begin[b], t1 = hp.tile_begin[b + 1];
		if (t0 == t1)
			continue;
		EncodeParams p;
		p.src = src;
		p.src_bytes = src_bytes;
		p.dst = dst;
		p.dst_bytes = dst_bytes;
		p.streams = hp.streams.data();
		p.tiles = hp.tiles.data() + t0;
		p.n_tiles = t1 - t0;
		p.results = NULL;
		p.order = NULL;
		p.n_streams = 0;
		switch (b) {
		case 0: emul_encode_bucket<4, 1>(p, order); break;
		case 1: emul_encode_bucket<4, 2>(p, order); break;
		case 2: emul_encode_bucket<6, 1>(p, order); break;
		case 3: emul_encode_bucket<6, 2>(p, order); break;
		case 4: emul_encode_bucket<8, 1>(p, order); break;
		default: emul_encode_bucket<8, 2>(p, order); break;
		}
	}
	return 0;
}

int
xa_emul_search(const bjxa_stream_desc_t *descs, size_t n, const uint8_t *src,
    uint8_t *dst, int16_t *prev_out)
{
	HostPlan hp;
	size_t bad = 0;
	int rc = build_plan(hp, kKindSearch, descs, n, &bad);
	if (rc)
		return rc;
	if (!hp.tiles.empty())
		return -1;		/* a search plan has no tiles */
	for (size_t i = 0; i < n; i++) {
		if (descs[i].blocks == 0)
			continue;
		switch (descs[i].bits) {
		case 4: emul_search_stream<4>(descs[i], src, dst, prev_out + 4 * i); break;
		case 6: emul_search_stream<6>(descs[i], src, dst, prev_out + 4 * i); break;
		default: emul_search_stream<8>(descs[i], src, dst, prev_out + 4 * i); break;
		}
	}
	return 0;
}

/* plan introspection for the host-logic tests: tiles as (first stream of the
 * tile, strip count, j) */
int
xa_emul_plan(int kind, const bjxa_stream_desc_t *descs, size_t n, int force_strips,
    uint32_t *tile_stream, uint32_t *tile_count, uint32_t *tile_j, uint32_t cap,
    uint32_t *tile_begin /* 7 */, uint32_t *n_slots, int *ns /* 6 */,
    uint32_t *alt_begin /* 7 */, int *alt_ns /* 6 */)
{
	HostPlan hp;
	size_t bad = 0;
	int rc = build_plan(hp, kind, descs, n, &bad, force_strips);
	if (rc)
		return -rc;
	uint32_t nt = (uint32_t)hp.tiles.size();
	for (uint32_t i = 0; i < nt && i < cap; i++) {
		tile_stream[i] = kind == kKindDecode ? hp.order[hp.tiles[i].first] :
		    hp.tiles[i].first;
		tile_count[i] = hp.tiles[i].count;
		tile_j[i] = hp.tiles[i].j;
	}
	memcpy(tile_begin, hp.tile_begin, 7 * sizeof(uint32_t));
	memcpy(ns, hp.ns, 6 * sizeof(int));
	memcpy(alt_begin, hp.alt_begin, 7 * sizeof(uint32_t));
	memcpy(alt_ns, hp.alt_ns, 6 * sizeof(int));
	*n_slots = hp.n_slots;
	return (int)nt;
}

void xa_emul_stereo_direct(int on) { g_stereo_direct = on; }
void xa_emul_use_alt(int on) { g_use_alt = on; }
void xa_emul_pool(int on) { g_pool = on; }
void xa_emul_split(int on) { g_split = on; }
void xa_emul_relay(int on) { g_relay = on; }
void xa_emul_seg(int on) { g_seg = on; }
int xa_emul_seg_items(void) { return (int)kSegItems; }
int xa_emul_seg_back(void) { return (int)kSegBack; }
int xa_emul_strip_blocks(int ns, int ch) { return (int)strip_blocks(ns, ch); }
int xa_emul_wide(void) { return kDecWide; }
int xa_emul_enc_tile_blocks(void) { return kEncTBE; }

}
