/*
 * xa_core.h -- per-block arithmetic of the BandJAM XA transform, written once
 * for the sm_100a kernels (xa_kernels.cu) and compilable by a host compiler so
 * tests/emul can single-step the very same code on a CPU (test harness only;
 * the product library has no CPU path).
 *
 * One "block" is 32 samples of one channel: 1 profile byte + 4*BITS payload
 * bytes in, 32 int16 out (/root/reference/bjxa.5.rst:107-187).
 *
 * Representation used throughout: a code is held TOP-ALIGNED in a 32-bit
 * register (x = code << (32-BITS), low bits zero).  The reference keeps it
 * top-aligned in an int16 and shifts by `range` (src/libbjxa.c:296-339,558);
 * with x = dst << 16 the same value is  x >> (16 + range)  (arithmetic).
 */
#ifndef XA_CORE_H
#define XA_CORE_H

#include <stdint.h>

#if defined(__CUDACC__)
#define XA_HD __host__ __device__ __forceinline__
#else
#define XA_HD inline
#endif

#if !defined(__CUDACC__)
/* host twins of the CUDA vector types the tile code moves data with */
struct uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
#endif

namespace xa {

constexpr int kBlockSamples = 32;

XA_HD constexpr int block_bytes(int bits) { return 4 * bits + 1; }

/* ---- funnel shift / byte permute with host twins ----------------------- */

/* low 32 bits of ((hi:lo) >> sh), 0 <= sh <= 31 */
XA_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh)
{
#if defined(__CUDA_ARCH__)
	return __funnelshift_r(lo, hi, sh);
#else
	return sh ? (lo >> sh) | (hi << (32u - sh)) : lo;
#endif
}

/* PRMT, default mode: result byte i = byte (sel >> 4i & 7) of the pair b:a */
XA_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
	return __byte_perm(a, b, sel);
#else
	uint64_t pair = ((uint64_t)b << 32) | a;
	uint32_t r = 0;
	for (int i = 0; i < 4; i++) {
		uint32_t s = (sel >> (4 * i)) & 7u;
		r |= (uint32_t)((pair >> (8 * s)) & 0xffu) << (8 * i);
	}
	return r;
#endif
}

/* two int32 already inside int16 range -> one word, lo | hi << 16 */
XA_HD uint32_t pack2(int lo, int hi)
{
	return byte_perm((uint32_t)lo, (uint32_t)hi, 0x5410);
}

/* ---- profile byte ------------------------------------------------------ */

/* K0, K1 scaled by 256 (src/libbjxa.c:525-531, bjxa.5.rst:123-129) */
XA_HD int gain_k0(unsigned f)
{
	return f == 1 ? 240 : f == 2 ? 460 : f == 3 ? 392 : f == 4 ? 488 : 0;
}
XA_HD int gain_k1(unsigned f)
{
	return f == 2 ? -208 : f == 3 ? -220 : f == 4 ? -240 : 0;
}

/* How a block takes part in the decode schedule. */
enum BlockKind {
	kCut = 0,	/* filter 0: output does not depend on predictor state */
	kChain = 1,	/* filter 1..4: needs the two previous samples */
	kBad = 2	/* filter >= 5: EPROTO in the reference (libbjxa.c:550) */
};

XA_HD int block_kind(uint32_t profile)
{
	unsigned f = profile >> 4;
	return f == 0 ? kCut : (f < 5 ? kChain : kBad);
}

/* ---- payload access ---------------------------------------------------- */

/*
 * A block's payload as BITS little-endian words, realigned from a byte
 * address that is generally odd: `words` points at the 4-byte aligned word
 * holding payload byte 0, `sh` = 8 * (address & 3).  Reads BITS+1 words.
 */
template <int BITS>
XA_HD void load_payload(uint32_t (&pw)[BITS], const uint32_t *words, uint32_t sh)
{
	uint32_t prev = words[0];
#pragma unroll
	for (int i = 0; i < BITS; i++) {
		uint32_t next = words[i + 1];
		pw[i] = funnel_r(prev, next, sh);
		prev = next;
	}
}

/*
 * Code n (0..31) of a block, top-aligned in 32 bits with zeros below.
 * N must be a compile-time constant after unrolling.
 *   4 bit: high nibble then low nibble of each byte   (libbjxa.c:296-297)
 *   6 bit: four fields of a big-endian 24-bit group    (libbjxa.c:315-321)
 *   8 bit: one byte                                    (libbjxa.c:339)
 */
template <int BITS>
XA_HD int top_code(const uint32_t (&pw)[BITS], int n)
{
	if (BITS == 8) {
		/* byte (n&3) of word n>>2 into byte 3, zeros elsewhere */
		return (int)byte_perm(pw[n >> 2], 0u, (uint32_t)(((n & 3) << 12) | 0x0444));
	} else if (BITS == 4) {
		int byte = (n & 7) >> 1;
		int pos = byte * 8 + ((n & 1) ? 0 : 4);	/* LSB of the nibble */
		return (int)((pw[n >> 3] << (28 - pos)) & 0xf0000000u);
	} else {
		/* group g = n>>2 covers payload bytes 3g..3g+2 */
		int b = 3 * (n >> 2);
		int w = b >> 2, o = b & 3;
		/* T = b0<<24 | b1<<16 | b2<<8 | junk */
		uint32_t hi = (w + 1 < BITS) ? pw[w + 1] : 0u;
		uint32_t sel = (uint32_t)(((o) << 12) | ((o + 1) << 8) | ((o + 2) << 4) | 0);
		uint32_t t = byte_perm(pw[w], hi, sel);
		return (int)((t << (6 * (n & 3))) & 0xfc000000u);
	}
}

/* ---- one sample -------------------------------------------------------- */

/* filter 0: the ranged code IS the sample (gain 0, and it cannot leave int16) */
XA_HD int sample_cut(int x, int sh) { return x >> sh; }

/*
 * filters 1..4 (src/libbjxa.c:556-571):
 *   sample = ranged + (p0*k0 + p1*k1) / 256   -- division truncates toward 0
 *   clamp to int16; the CLAMPED value becomes the new state.
 */
XA_HD int sample_chain(int x, int sh, int k0, int k1, int &p0, int &p1)
{
	/* tools/lat_bench.cu: the plain form below costs 34 cycles/sample on the
	 * dependent path on B200 (26 of them without the clamp); rearrangements
	 * with two shifted candidates and a select measured 34-37 */
	int g = p0 * k0 + p1 * k1;
#if defined(__CUDA_ARCH__) && !defined(XA_CHAIN_NO_IMAD)
	/* the bias of the truncating division, (g < 0 ? 255 : 0), as one
	 * multiply-add on the FMA pipe instead of AND + ADD on the ALU pipe: one
	 * step less on the dependent path (IMAD, SHF, IMAD, LEA.HI, 2 x VIMNMX)
	 * and one ALU-pipe instruction less per sample; +3..6 % on chain-rich data */
	int f;
	asm("mad.lo.s32 %0, %1, -255, %2;" : "=r"(f) : "r"(g >> 31), "r"(g));
	int q = f >> 8;
#else
	int q = (g + ((g >> 31) & 255)) >> 8;	/* truncating /256 */
#endif
	int s = (x >> sh) + q;
	s = s < -32768 ? -32768 : s;
	s = s > 32767 ? 32767 : s;
	p1 = p0;
	p0 = s;
	return s;
}

/*
 * The same step with state and result carrying +32768 ("biased": 0..65535), so
 * that the int16 clamp is ONE instruction, max(min(v, 65535), 0) = VIMNMX.RELU,
 * and the dependent path five: IMAD, SHF, IMAD, LEA.HI, VIMNMX.RELU.
 *   c  = -32768 * (k0 + k1)  makes  b0*k0 + b1*k1 + c  the reference's gain
 *        exactly ((p0 + 32768)*k0 + (p1 + 32768)*k1 - 32768*(k0 + k1));
 *   |b*k| <= 65535 * 488, |c| <= 32768 * 252, + 2^23: all inside int32.
 * A pair of biased samples is packed by pack2_biased.  -DXA_CHAIN_PLAIN keeps
 * the walkers on sample_chain.
 */
XA_HD int chain_bias_c(int k0, int k1)
{
	return -32768 * (k0 + k1);
}

XA_HD int sample_chain_b(int x, int sh, int k0, int k1, int c, int &b0, int &b1)
{
#if defined(__CUDA_ARCH__)
	/* g for the sign; g2 = g + 2^23 carries the +32768 of the result through
	 * the division ((f + 2^23) >> 8 = (f >> 8) + 32768 exactly), so biasing the
	 * code costs no ALU-pipe instruction: everything but SHF, LEA.HI and
	 * VIMNMX.RELU runs on the FMA pipe, which the walkers leave idle */
	int g = b0 * k0 + (b1 * k1 + c);
	int g2 = b0 * k0 + (b1 * k1 + (c + (1 << 23)));
	int f;
	asm("mad.lo.s32 %0, %1, -255, %2;" : "=r"(f) : "r"(g >> 31), "r"(g2));
	int s = __vimin_s32_relu((f >> 8) + (x >> sh), 65535);
#else
	int g = b0 * k0 + (b1 * k1 + c);
	int q = (g + ((g >> 31) & 255)) >> 8;
	int s = (x >> sh) + 32768 + q;
	s = s < 0 ? 0 : s;
	s = s > 65535 ? 65535 : s;
#endif
	b1 = b0;
	b0 = s;
	return s;
}

/*
 * The biased step with the +32768 of the result added to the ranged code instead
 * of riding through the division: two multiply-adds fewer per sample (9.25
 * instructions instead of 10.75).  Where a kernel's warps are bound by instruction
 * issue rather than by one pipe -- tools/step_bench.cu: a warp issues about one
 * instruction every 1.45 cycles whatever the pipe mix, 24 to 32 warps per SM --
 * that is 11 % more samples per cycle, and 7 % for a warp alone on its SM (31
 * instead of 33 cycles a sample).  The tile walkers keep sample_chain_b: their
 * kernels are capped at 48 registers and bound by the ALU pipe.
 */
XA_HD int sample_chain_r(int x, int sh, int k0, int k1, int c, int &b0, int &b1)
{
#if defined(__CUDA_ARCH__)
	int g = b0 * k0 + (b1 * k1 + c);
	int f;
	asm("mad.lo.s32 %0, %1, -255, %2;" : "=r"(f) : "r"(g >> 31), "r"(g));
	int s = __vimin_s32_relu((f >> 8) + ((x >> sh) + 32768), 65535);
	b1 = b0;
	b0 = s;
	return s;
#else
	return sample_chain_b(x, sh, k0, k1, c, b0, b1);
#endif
}

/* two biased samples -> one word of int16, lo | hi << 16 */
XA_HD uint32_t pack2_biased(int lo, int hi)
{
#if defined(__CUDA_ARCH__)
	uint32_t w;
	asm("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(w) : "r"(hi), "r"(lo));
	return w ^ 0x80008000u;
#else
	return ((uint32_t)hi << 16 | (uint32_t)lo) ^ 0x80008000u;
#endif
}

/*
 * The four codes of one "quad" (4 consecutive samples = BITS/2 payload bytes,
 * given in the low bytes of w), each top-aligned in 32 bits with zeros below.
 */
template <int BITS>
XA_HD void quad_codes(uint32_t w, int (&x)[4])
{
	if (BITS == 8) {
#pragma unroll
		for (int i = 0; i < 4; i++)
			x[i] = (int)byte_perm(w, 0u, (uint32_t)((i << 12) | 0x0444));
	} else if (BITS == 4) {
		x[0] = (int)((w << 24) & 0xf0000000u);	/* high nibble of byte 0 */
		x[1] = (int)((w << 28) & 0xf0000000u);	/* low nibble of byte 0 */
		x[2] = (int)((w << 16) & 0xf0000000u);
		x[3] = (int)((w << 20) & 0xf0000000u);
	} else {
		/* three bytes, big endian: b0<<24 | b1<<16 | b2<<8 */
		uint32_t t = byte_perm(w, 0u, 0x0124);
#pragma unroll
		for (int i = 0; i < 4; i++)
			x[i] = (int)((t << (6 * i)) & 0xfc000000u);
	}
}

/* ---- whole blocks, results as 16 packed words (sample 2i | 2i+1 << 16) -- */

template <int BITS>
XA_HD void decode_block_cut(uint32_t (&out)[16], const uint32_t (&pw)[BITS],
    uint32_t profile)
{
	const int sh = 16 + (int)(profile & 15u);
#pragma unroll
	for (int i = 0; i < 16; i++) {
		int a = sample_cut(top_code<BITS>(pw, 2 * i), sh);
		int b = sample_cut(top_code<BITS>(pw, 2 * i + 1), sh);
		out[i] = pack2(a, b);
	}
}

template <int BITS, bool RANGED = false>
XA_HD void decode_block_chain(uint32_t (&out)[16], const uint32_t (&pw)[BITS],
    uint32_t profile, int &p0, int &p1)
{
	const int sh = 16 + (int)(profile & 15u);
	const int k0 = gain_k0(profile >> 4), k1 = gain_k1(profile >> 4);
#if !defined(XA_CHAIN_PLAIN)
	const int c = chain_bias_c(k0, k1);
	int b0 = p0 + 32768, b1 = p1 + 32768;
#pragma unroll
	for (int i = 0; i < 16; i++) {
		int a, b;
		if (RANGED) {
			a = sample_chain_r(top_code<BITS>(pw, 2 * i), sh, k0, k1, c, b0, b1);
			b = sample_chain_r(top_code<BITS>(pw, 2 * i + 1), sh, k0, k1, c, b0, b1);
		} else {
			a = sample_chain_b(top_code<BITS>(pw, 2 * i), sh, k0, k1, c, b0, b1);
			b = sample_chain_b(top_code<BITS>(pw, 2 * i + 1), sh, k0, k1, c, b0, b1);
		}
		out[i] = pack2_biased(a, b);
	}
	p0 = b0 - 32768;
	p1 = b1 - 32768;
#else
#pragma unroll
	for (int i = 0; i < 16; i++) {
		int a = sample_chain(top_code<BITS>(pw, 2 * i), sh, k0, k1, p0, p1);
		int b = sample_chain(top_code<BITS>(pw, 2 * i + 1), sh, k0, k1, p0, p1);
		out[i] = pack2(a, b);
	}
#endif
}

/* ---- encode: 32 int16 (16 packed words) -> 4*BITS payload bytes --------- */

/*
 * Keeps the top BITS bits of every sample (logical shift of the uint16 image,
 * src/libbjxa.c:355-387) and packs MSB-first; result as BITS little-endian
 * words, i.e. the payload's byte image.
 */
template <int BITS>
XA_HD void deflate_block(uint32_t (&pw)[BITS], const uint32_t (&in)[16])
{
	if (BITS == 8) {
#pragma unroll
		for (int i = 0; i < 8; i++)	/* high bytes of 4 samples */
			pw[i] = byte_perm(in[2 * i], in[2 * i + 1], 0x7531);
	} else if (BITS == 4) {
#pragma unroll
		for (int i = 0; i < 4; i++) {
			uint32_t w = 0;
#pragma unroll
			for (int b = 0; b < 4; b++) {
				uint32_t pr = in[4 * i + b];	/* s0 | s1<<16 */
				uint32_t byte = ((pr >> 8) & 0xf0u) | (pr >> 28);
				w |= byte << (8 * b);
			}
			pw[i] = w;
		}
	} else {
		/* 4 samples -> 24 bits big-endian -> 3 bytes; 8 groups -> 6 words */
		uint32_t g[8];
#pragma unroll
		for (int i = 0; i < 8; i++) {
			uint32_t a = in[2 * i], b = in[2 * i + 1];
			uint32_t v = ((a >> 10) & 0x3fu) << 18 | (a >> 26) << 12 |
			    ((b >> 10) & 0x3fu) << 6 | (b >> 26);
			/* byte image b0,b1,b2 = v>>16, v>>8, v : as LE word */
			g[i] = byte_perm(v, 0u, 0x4012);
		}
		/* concatenate eight 3-byte groups into six words */
#pragma unroll
		for (int k = 0; k < 2; k++) {
			const uint32_t *q = g + 4 * k;
			pw[3 * k + 0] = q[0] | (q[1] << 24);
			pw[3 * k + 1] = (q[1] >> 8) | (q[2] << 16);
			pw[3 * k + 2] = (q[2] >> 16) | (q[3] << 8);
		}
	}
}

/*
 * The same truncation for FOUR samples at a time: returns the BITS/2 payload
 * bytes they occupy (4-bit: 2, 6-bit: 3, 8-bit: 4), first byte in bits 0..7.
 * s0..s3 hold the samples in their low 16 bits (upper bits ignored).
 */
template <int BITS>
XA_HD uint32_t pack4(uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3)
{
	if (BITS == 8) {
		return ((s0 >> 8) & 0xffu) | (s1 & 0xff00u) | ((s2 & 0xff00u) << 8) |
		    ((s3 & 0xff00u) << 16);
	} else if (BITS == 4) {
		uint32_t b0 = ((s0 >> 8) & 0xf0u) | ((s1 >> 12) & 0x0fu);
		uint32_t b1 = ((s2 >> 8) & 0xf0u) | ((s3 >> 12) & 0x0fu);
		return b0 | (b1 << 8);
	} else {
		uint32_t v = ((s0 >> 10) & 0x3fu) << 18 | ((s1 >> 10) & 0x3fu) << 12 |
		    ((s2 >> 10) & 0x3fu) << 6 | ((s3 >> 10) & 0x3fu);
		return byte_perm(v, 0u, 0x4012);	/* big-endian 24 bits -> byte order */
	}
}

/* ---- searching encoder (an extension: SURVEY.md section 8, row E4) -------- */

/*
 * The reference encoder writes profile 0 and the top bits of every sample
 * (src/libbjxa.c:679); nothing in the reference searches.  This extension picks
 * a filter and a range per block, closed loop: for every candidate
 *     filter f = 0..4,  range r = 0..16-BITS      (profile byte f << 4 | r)
 * the block is encoded and decoded again exactly as the reference DEcoder would
 * (src/libbjxa.c:556-571), starting from the decoder state the previous block's
 * winner left behind:
 *     pred  = (q0*k0 + q1*k1) / 256                    truncating toward zero
 *     step  = 1 << (16 - BITS - r)
 *     code  = clamp(floor((x - pred + step/2) / step), -2^(BITS-1), 2^(BITS-1)-1)
 *     s     = clamp_int16(code * step + pred);   err += (x - s)^2;   q1 = q0; q0 = s
 * and the candidate with the smallest (err, profile byte) wins.  Candidate
 * index c = f * (17 - BITS) + r orders the candidates by profile byte.  The
 * candidate (f, r) = (0, 0) never does worse than the reference's truncation,
 * whatever the state, so the result never has a larger error than the
 * reference encoder's.  (The test suite holds a plain-C restatement.)
 */
XA_HD constexpr int search_ranges(int bits) { return 17 - bits; }
XA_HD constexpr int search_candidates(int bits) { return 5 * (17 - bits); }

/*
 * The step above, rearranged for the machine like the decoder's (sample_chain_b):
 * the ALU pipe takes a warp instruction every other cycle and was what bounded
 * the search (78 % busy), the FMA pipe takes one per cycle and idled.  State,
 * sample and result carry +32768, the code carries +2^(BITS-1), so that both
 * clamps are one VIMNMX.RELU each; every constant rides in a multiply-add:
 *     g    = b0*k0 + b1*k1 + c                c = -32768*(k0 + k1): the gain, exactly
 *     f    = g + (g < 0 ? 255 : 0) + d        d = 2^23 - ((M + step/2) << 8),  M = 2^(BITS-1) * step
 *     pb   = f >> 8                           = pred + 32768 - M - step/2
 *     cb   = relu_min((xb - pb) >> shift, 2^BITS - 1)             = code + 2^(BITS-1)
 *     sb   = relu_min(cb*step + pb + step/2, 65535) = s + 32768;  err += (xb - sb)^2
 * (adding M to the numerator adds exactly 2^(BITS-1) to the quotient; cb*step + pb
 * + step/2 = code*step + pred + 32768; the last add and both bounds are one
 * VIADDMNMX.RELU).  Three shifts and two clamps are all that is left
 * on the ALU pipe.  The biased codes are packed as they are and un-biased once per
 * block: search_code_bias() is the XOR mask of all 32 top bits.
 */
template <int BITS>
struct SearchK {		/* one candidate's constants */
	int k0, k1, c, d, shift, step, half;
};

template <int BITS>
XA_HD void search_setup(SearchK<BITS> &K, unsigned f, int shift)
{
	K.k0 = gain_k0(f);
	K.k1 = gain_k1(f);
	K.c = -32768 * (K.k0 + K.k1);
	K.shift = shift;
	K.step = 1 << shift;
	K.half = K.step >> 1;
	/* + 32768, - M, - step/2, all times 256: they leave the division as they are */
	K.d = (1 << 23) - ((K.step << (BITS - 1)) << 8) - (K.half << 8);
}

template <int BITS>
XA_HD int search_sample_b(int xb, const SearchK<BITS> &K, int &b0, int &b1,
    unsigned long long &err)
{
	const int g = b0 * K.k0 + (b1 * K.k1 + K.c);
	const int g2 = g + K.d;
#if defined(__CUDA_ARCH__)
	/* the multiply-adds are spelled out: left to itself the compiler turns
	 * cb * step into a shift and the subtraction into a three-input add, both
	 * on the ALU pipe */
	int f, num, v;
	asm("mad.lo.s32 %0, %1, -255, %2;" : "=r"(f) : "r"(g >> 31), "r"(g2));
	const int pb = f >> 8;				/* pred + 32768 - M - step/2 */
	asm("mad.lo.s32 %0, %1, -1, %2;" : "=r"(num) : "r"(pb), "r"(xb));
	const int cb = __vimin_s32_relu(num >> K.shift, (1 << BITS) - 1);
	asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(v) : "r"(cb), "r"(K.step), "r"(pb));
	const int sb = __viaddmin_s32_relu(v, K.half, 65535);
#else
	const int pb = (g2 + ((g >> 31) & 255)) >> 8;
	int cb = (xb - pb) >> K.shift;
	cb = cb < 0 ? 0 : cb;
	cb = cb > (1 << BITS) - 1 ? (1 << BITS) - 1 : cb;
	int sb = cb * K.step + pb + K.half;
	sb = sb < 0 ? 0 : sb;
	sb = sb > 65535 ? 65535 : sb;
#endif
	const long long e = (long long)(xb - sb);
	err += (unsigned long long)(e * e);
	b1 = b0;
	b0 = sb;
	return cb;
}

/*
 * Puts code i (i = 0..31, a compile-time constant once the caller's loop is
 * unrolled) into the payload's byte image, BITS little-endian words: the codes
 * form one MSB-first bit stream (bjxa.5.rst: high nibble first; 6-bit codes
 * four to three bytes, big endian).
 */
template <int BITS>
XA_HD void put_code(uint32_t (&w)[BITS], int i, int code)
{
	const uint32_t c = (uint32_t)code & ((1u << BITS) - 1u);
	const int bit = BITS * i, byte = bit >> 3, top = bit & 7;
	if (top + BITS <= 8) {
		w[byte >> 2] |= (c << (8 - top - BITS)) << (8 * (byte & 3));
	} else {
		const int spill = top + BITS - 8;	/* bits that go to the next byte */
		w[byte >> 2] |= (c >> spill) << (8 * (byte & 3));
		w[(byte + 1) >> 2] |= ((c << (8 - spill)) & 0xffu) << (8 * ((byte + 1) & 3));
	}
}

/* the byte image of 32 codes 2^(BITS-1): XOR it onto a block of biased codes */
template <int BITS>
XA_HD void search_code_bias(uint32_t (&w)[BITS])
{
#pragma unroll
	for (int k = 0; k < BITS; k++)
		w[k] = 0;
#pragma unroll
	for (int i = 0; i < 32; i++)
		put_code<BITS>(w, i, 1 << (BITS - 1));
}

} /* namespace xa */
#endif
