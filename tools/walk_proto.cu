// Prototype (measurement tool, not product): dense chain walkers for chain-rich
// decode.  Mono 8-bit streams, profile mix with a given share of chain blocks.
//   pass 1   one thread per 16-byte unit of output of every cut block
//   heads    chain heads compacted into a global list (cub)
//   pass 2   walker variants: every LANE walks one chain at a time, drawing the
//            next from the list as soon as its chain ends; payload re-read per
//            lane (16-byte loads through a private shared-memory window), output
//            stored directly (V_DIRECT), with 256-bit stores (V_WIDE) or staged
//            through shared memory and copied out eight rows per instruction
//            (V_STAGED)
// plus micro-benchmarks of scattered 16-byte loads / stores (L1 wavefront cost).
// Everything is checked against a serial one-thread-per-stream decode.
//
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o build/walk_proto tools/walk_proto.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cub/cub.cuh>

#include "../bjxa_b200/csrc/xa_core.h"

using namespace xa;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
	fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int BITS = 8, BS = 33;

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
	x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
	return x;
}

// chain_pm: permille of blocks with filter 1..4; 1001 = P2 (uniform 0..4); 2000 = P3
__global__ void gen_kernel(uint8_t *xa, uint64_t nblocks_total, uint32_t chain_pm)
{
	uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (b >= nblocks_total)
		return;
	uint32_t h = hash32((uint32_t)b * 2654435761u + 12345u);
	uint32_t filt;
	if (chain_pm == 1001)
		filt = h % 5u;
	else if (chain_pm == 2000)
		filt = 1 + h % 4u;
	else
		filt = (h % 1000u) < chain_pm ? 1 + (h >> 12) % 4u : 0u;
	uint32_t range = (h >> 20) & 15u;
	uint8_t *p = xa + b * BS;
	p[0] = (uint8_t)(filt << 4 | range);
	for (int i = 0; i < 32; i += 4) {
		uint32_t r = hash32((uint32_t)b * 40503u + i * 7919u + 99u);
		p[1 + i] = (uint8_t)r; p[2 + i] = (uint8_t)(r >> 8);
		p[3 + i] = (uint8_t)(r >> 16); p[4 + i] = (uint8_t)(r >> 24);
	}
}

// serial reference: one thread per stream
__global__ void ref_kernel(const uint8_t *xa, int16_t *out, uint32_t nstreams, uint32_t nblocks)
{
	uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= nstreams)
		return;
	const uint8_t *p = xa + (uint64_t)s * nblocks * BS;
	int16_t *o = out + (uint64_t)s * nblocks * 32;
	int p0 = 0, p1 = 0;
	for (uint32_t b = 0; b < nblocks; b++, p += BS, o += 32) {
		uint32_t prof = p[0];
		int k0 = gain_k0(prof >> 4), k1 = gain_k1(prof >> 4);
		int sh = prof & 15;
		for (int i = 0; i < 32; i++) {
			int x = (int)(int16_t)((uint16_t)p[1 + i] << 8) >> sh;
			int g = p0 * k0 + p1 * k1;
			int v = x + g / 256;
			v = v < -32768 ? -32768 : v > 32767 ? 32767 : v;
			p1 = p0; p0 = v;
			o[i] = (int16_t)v;
		}
	}
}

__global__ void flag_heads(const uint8_t *xa, uint8_t *flags, uint64_t total, uint32_t nblocks)
{
	uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (b >= total)
		return;
	bool chain = block_kind(xa[b * BS]) == kChain;
	bool first = (b % nblocks) == 0;
	bool prevchain = !first && block_kind(xa[(b - 1) * BS]) == kChain;
	flags[b] = chain && !prevchain;
}

// pass 1: one thread per 16-byte unit of every cut block (unoptimised stand-in)
__global__ void units_kernel(const uint8_t *xa, uint8_t *out, uint64_t total_units)
{
	uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (u >= total_units)
		return;
	uint64_t b = u >> 2;
	uint32_t k = u & 3;
	const uint8_t *p = xa + b * BS;
	uint32_t prof = p[0];
	if (block_kind(prof) == kChain)
		return;
	const int sh = 16 + (prof & 15);
	uint32_t w[4];
#pragma unroll
	for (int j = 0; j < 4; j++) {
		int a = (int)((uint32_t)p[1 + k * 8 + 2 * j] << 24) >> sh;
		int c = (int)((uint32_t)p[2 + k * 8 + 2 * j] << 24) >> sh;
		w[j] = pack2(a, c);
	}
	*reinterpret_cast<uint4 *>(out + u * 16) = make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void compare_kernel(const uint4 *a, const uint4 *b, uint64_t n, unsigned long long *bad)
{
	uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	uint4 x = a[i], y = b[i];
	if (x.x != y.x || x.y != y.y || x.z != y.z || x.w != y.w)
		atomicAdd(bad, 1ULL);
}

struct WalkParams {
	const uint8_t *xa;
	uint8_t *out;
	const uint32_t *heads;	// global block index of every head
	uint32_t n_heads;
	uint32_t nblocks;	// per stream
	unsigned long long *counter;
	unsigned long long *stats;	// [0] warp turns, [1] lane-blocks
};

enum { V_DIRECT = 0, V_WIDE = 1, V_STAGED = 2 };

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void st_v8(void *p, const uint32_t (&o)[16], int half)
{
	asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::
	    "l"(p), "r"(o[8 * half]), "r"(o[8 * half + 1]), "r"(o[8 * half + 2]), "r"(o[8 * half + 3]),
	    "r"(o[8 * half + 4]), "r"(o[8 * half + 5]), "r"(o[8 * half + 6]), "r"(o[8 * half + 7]) : "memory");
}

constexpr int WT = 256;		// threads per walker CTA

// per-lane window: 2 x 48 bytes (double buffer) ; staged out rows: 64 B per lane
template <int V, int CTAS>
__global__ void __launch_bounds__(WT, CTAS)
walk_kernel(const WalkParams p)
{
	__shared__ __align__(16) uint8_t win[2 * WT + 1][48];	// + 1: load_payload reads one word past a window
	__shared__ __align__(16) uint32_t rows[V == V_STAGED ? WT * 16 : 4];
	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	bool have = false, dead = false;
	uint64_t a = 0;		// byte address (arena offset) of the current block
	uint32_t left = 0;	// blocks left in the stream after the current one
	int p0 = 0, p1 = 0, cur = 0;
	unsigned long long turns = 0, lane_blocks = 0;

	for (;;) {
		const uint32_t idle = __ballot_sync(0xffffffffu, !have && !dead);
		if (idle) {
			const int leader = __ffs(idle) - 1;
			unsigned long long base = 0;
			if ((int)lane == leader)
				base = atomicAdd(p.counter, (unsigned long long)__popc(idle));
			base = __shfl_sync(0xffffffffu, base, leader);
			if (!have && !dead) {
				const unsigned long long i = base + __popc(idle & ((1u << lane) - 1u));
				if (i >= p.n_heads) {
					dead = true;
				} else {
					const uint32_t gb = p.heads[i];
					const uint32_t lb = gb % p.nblocks;
					a = (uint64_t)gb * BS;
					left = p.nblocks - 1 - lb;
					if (lb == 0) {
						p0 = p1 = 0;
					} else {
						const uint8_t *q = p.xa + a;
						const int sh = 16 + (q[-BS] & 15);
						p1 = (int)((uint32_t)q[-2] << 24) >> sh;
						p0 = (int)((uint32_t)q[-1] << 24) >> sh;
					}
					// fetch the head block's window
					const uint8_t *g = p.xa + (a & ~15ULL);
					const uint32_t d = (uint32_t)__cvta_generic_to_shared(win[cur * WT + tid]);
					cp_async16(d, g); cp_async16(d + 16, g + 16); cp_async16(d + 32, g + 32);
					cp_async_commit();
					have = true;
				}
			}
		}
		if (!__any_sync(0xffffffffu, have))
			break;
		turns++;
		uint32_t o[16];
		uint8_t *dst = NULL;
		bool store = false;
		if (have) {
			lane_blocks++;
			cp_async_wait0();
			// prefetch the next block of the stream (whether or not it continues the chain)
			if (left != 0) {
				const uint8_t *g = p.xa + ((a + BS) & ~15ULL);
				const uint32_t d = (uint32_t)__cvta_generic_to_shared(win[(cur ^ 1) * WT + tid]);
				cp_async16(d, g); cp_async16(d + 16, g + 16); cp_async16(d + 32, g + 32);
				cp_async_commit();
			}
			const uint8_t *w8 = win[cur * WT + tid];
			const uint32_t off = (uint32_t)(a & 15u);
			const uint32_t prof = w8[off];
			const uint32_t pay = off + 1;
			uint32_t pw[BITS];
			load_payload<BITS>(pw, reinterpret_cast<const uint32_t *>(w8) + (pay >> 2), (pay & 3u) * 8u);
			decode_block_chain<BITS>(o, pw, prof, p0, p1);
			dst = p.out + (a / BS) * 64;
			store = true;
			// does the chain go on?
			bool more = false;
			if (left != 0) {
				cp_async_wait0();
				const uint64_t na = a + BS;
				const uint32_t nprof = win[(cur ^ 1) * WT + tid][na & 15u];
				more = block_kind(nprof) == kChain;
				a = na;
				left--;
				cur ^= 1;
			}
			have = more;
		}
		if (V == V_DIRECT) {
			if (store) {
#pragma unroll
				for (int j = 0; j < 4; j++)
					reinterpret_cast<uint4 *>(dst)[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
			}
		} else if (V == V_WIDE) {
			if (store) {
				if (((uintptr_t)dst & 31u) == 0) {
					st_v8(dst, o, 0);
					st_v8(dst + 32, o, 1);
				} else {
#pragma unroll
					for (int j = 0; j < 4; j++)
						reinterpret_cast<uint4 *>(dst)[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
				}
			}
		} else {
			// rows: lane l writes chunk j of its 64-byte row at chunk position j ^ (l >> 1 & 3)
			uint32_t *row = rows + (warp * 32 + lane) * 16;
			if (store) {
#pragma unroll
				for (int j = 0; j < 4; j++)
					*reinterpret_cast<uint4 *>(row + ((j ^ (int)((lane >> 1) & 3u)) * 4)) =
					    make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
			}
			__syncwarp();
			const unsigned long long d64 = store ? (unsigned long long)dst : 0ULL;
#pragma unroll
			for (int r = 0; r < 4; r++) {
				const uint32_t src_lane = r * 8 + (lane >> 2), c = lane & 3u;
				const unsigned long long rd = __shfl_sync(0xffffffffu, d64, src_lane);
				if (rd != 0) {
					const uint4 v = *reinterpret_cast<const uint4 *>(
					    rows + (warp * 32 + src_lane) * 16 + ((c ^ ((src_lane >> 1) & 3u)) * 4));
					*reinterpret_cast<uint4 *>(rd + c * 16) = v;
				}
			}
			__syncwarp();
		}
	}
	if (lane == 0)
		atomicAdd(&p.stats[0], turns);
	atomicAdd(&p.stats[1], lane_blocks);
}

// ---- micro-benchmarks: scattered 16-byte accesses -------------------------------
// every lane touches `per` consecutive 16-byte chunks at lane stride `stride` bytes
template <int MODE>	// 0 = STG.128, 1 = LDG.128, 2 = st.v8 (32 B chunks)
__global__ void __launch_bounds__(256)
mb_kernel(uint8_t *buf, uint64_t span, uint32_t stride, uint32_t iters, uint32_t *sink)
{
	const uint32_t lane = threadIdx.x & 31u;
	const uint64_t gw = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
	uint64_t base = (gw * 32 * (uint64_t)stride * 7) % span;
	uint32_t acc = 0;
	for (uint32_t it = 0; it < iters; it++) {
		uint8_t *p = buf + ((base + (uint64_t)lane * stride) & ~15ULL);
		if (MODE == 0) {
#pragma unroll
			for (int j = 0; j < 4; j++)
				reinterpret_cast<uint4 *>(p)[j] = make_uint4(it, lane, j, acc);
		} else if (MODE == 1) {
#pragma unroll
			for (int j = 0; j < 4; j++) {
				uint4 v = reinterpret_cast<const uint4 *>(p)[j];
				acc += v.x ^ v.w;
			}
		} else {
			p = (uint8_t *)((uintptr_t)p & ~31ULL);
			uint32_t o[16];
#pragma unroll
			for (int j = 0; j < 16; j++)
				o[j] = it + j;
			st_v8(p, o, 0);
			st_v8(p + 32, o, 1);
		}
		base += 32ULL * stride;
		if (base + 32ULL * stride + 64 > span)
			base = 0;
	}
	if (acc == 0x12345u)
		*sink = acc;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b)
{
	float ms = 0;
	CK(cudaEventElapsedTime(&ms, a, b));
	return ms;
}

template <int V, int CTAS>
static void run_walk(const char *name, WalkParams wp, const uint8_t *d_ref, uint64_t out_bytes,
    unsigned long long *d_bad, double chain_samples, int sms)
{
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	float best = 1e30f;
	unsigned long long st[2] = { 0, 0 };
	for (int rep = 0; rep < 3; rep++) {
		CK(cudaMemset(wp.counter, 0, 8));
		CK(cudaMemset(wp.stats, 0, 16));
		CK(cudaEventRecord(e0));
		walk_kernel<V, CTAS><<<sms * CTAS, WT>>>(wp);
		CK(cudaEventRecord(e1));
		CK(cudaEventSynchronize(e1));
		CK(cudaGetLastError());
		float ms = time_ms(e0, e1);
		if (ms < best) best = ms;
	}
	CK(cudaMemcpy(st, wp.stats, 16, cudaMemcpyDeviceToHost));
	CK(cudaMemset(d_bad, 0, 8));
	uint64_t n16 = out_bytes / 16;
	compare_kernel<<<(unsigned)((n16 + 255) / 256), 256>>>((const uint4 *)wp.out, (const uint4 *)d_ref, n16, d_bad);
	unsigned long long bad = 0;
	CK(cudaMemcpy(&bad, d_bad, 8, cudaMemcpyDeviceToHost));
	printf("%-22s ctas/SM %d  %8.3f ms  %7.1f Gsamples/s (chain)  lanes/turn %.1f  mismatching units %llu\n",
	    name, CTAS, best, chain_samples / best / 1e6, st[0] ? (double)st[1] / st[0] : 0.0, bad);
	fflush(stdout);
}

int main(int argc, char **argv)
{
	uint32_t nstreams = argc > 1 ? atoi(argv[1]) : 1024;
	uint32_t nblocks = argc > 2 ? atoi(argv[2]) : 82688;
	uint32_t chain_pm = argc > 3 ? atoi(argv[3]) : 1001;
	int sms = 0;
	CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
	uint64_t total = (uint64_t)nstreams * nblocks;
	uint64_t xa_bytes = total * BS, out_bytes = total * 64;
	uint8_t *d_xa, *d_out, *d_ref, *d_flags;
	uint32_t *d_heads, *d_nsel;
	unsigned long long *d_ctr, *d_stats, *d_bad;
	CK(cudaMalloc(&d_xa, xa_bytes + 256));
	CK(cudaMalloc(&d_out, out_bytes));
	CK(cudaMalloc(&d_ref, out_bytes));
	CK(cudaMalloc(&d_flags, total));
	CK(cudaMalloc(&d_heads, total * 4));
	CK(cudaMalloc(&d_nsel, 4));
	CK(cudaMalloc(&d_ctr, 8)); CK(cudaMalloc(&d_stats, 16)); CK(cudaMalloc(&d_bad, 8));
	CK(cudaMemset(d_xa, 0, xa_bytes + 256));
	gen_kernel<<<(unsigned)((total + 255) / 256), 256>>>(d_xa, total, chain_pm);
	ref_kernel<<<(nstreams + 63) / 64, 64>>>(d_xa, (int16_t *)d_ref, nstreams, nblocks);
	flag_heads<<<(unsigned)((total + 255) / 256), 256>>>(d_xa, d_flags, total, nblocks);
	CK(cudaDeviceSynchronize());

	// heads list (ordered)
	void *tmp = NULL; size_t tmp_bytes = 0;
	cub::CountingInputIterator<uint32_t> it0(0);
	CK(cub::DeviceSelect::Flagged(tmp, tmp_bytes, it0, d_flags, d_heads, d_nsel, (int)total));
	CK(cudaMalloc(&tmp, tmp_bytes));
	CK(cub::DeviceSelect::Flagged(tmp, tmp_bytes, it0, d_flags, d_heads, d_nsel, (int)total));
	uint32_t n_heads = 0;
	CK(cudaMemcpy(&n_heads, d_nsel, 4, cudaMemcpyDeviceToHost));

	// count chain blocks
	std::vector<uint8_t> h_prof;
	uint64_t chain_blocks = 0;
	{
		// sample on host from the first stream only (cheap) and scale
		std::vector<uint8_t> h((size_t)nblocks * BS);
		CK(cudaMemcpy(h.data(), d_xa, h.size(), cudaMemcpyDeviceToHost));
		uint64_t c = 0;
		for (uint32_t b = 0; b < nblocks; b++)
			c += block_kind(h[(size_t)b * BS]) == kChain;
		chain_blocks = c * nstreams;
	}
	double chain_samples = (double)chain_blocks * 32;
	printf("streams %u blocks %u mix %u: %.1f M blocks, %u heads, ~%.1f%% chain blocks, in %.2f GB out %.2f GB\n",
	    nstreams, nblocks, chain_pm, total / 1e6, n_heads, 100.0 * chain_blocks / total, xa_bytes / 1e9, out_bytes / 1e9);

	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));

	// pass 1 timing (units of cut blocks)
	CK(cudaMemset(d_out, 0xee, out_bytes));
	for (int rep = 0; rep < 2; rep++) {
		CK(cudaEventRecord(e0));
		units_kernel<<<(unsigned)((total * 4 + 255) / 256), 256>>>(d_xa, d_out, total * 4);
		CK(cudaEventRecord(e1));
		CK(cudaEventSynchronize(e1));
	}
	printf("pass 1 (naive units)   %8.3f ms\n", time_ms(e0, e1));

	WalkParams wp = { d_xa, d_out, d_heads, n_heads, nblocks, d_ctr, d_stats };
	run_walk<V_DIRECT, 2>("walk direct", wp, d_ref, out_bytes, d_bad, chain_samples, sms);
	run_walk<V_DIRECT, 4>("walk direct", wp, d_ref, out_bytes, d_bad, chain_samples, sms);
	run_walk<V_WIDE, 4>("walk 256-bit stores", wp, d_ref, out_bytes, d_bad, chain_samples, sms);
	run_walk<V_STAGED, 2>("walk staged copy-out", wp, d_ref, out_bytes, d_bad, chain_samples, sms);
	run_walk<V_STAGED, 4>("walk staged copy-out", wp, d_ref, out_bytes, d_bad, chain_samples, sms);
	run_walk<V_STAGED, 6>("walk staged copy-out", wp, d_ref, out_bytes, d_bad, chain_samples, sms);

	// micro-benchmarks: 64 bytes per lane and step, lanes `stride` bytes apart
	uint32_t *d_sink;
	CK(cudaMalloc(&d_sink, 4));
	const uint64_t span = 64ull << 20;	// L2 resident
	const uint32_t iters = 2000;
	const uint32_t strides[] = { 64, 128, 320, 1024 };
	for (int mode = 0; mode < 3; mode++) {
		for (uint32_t s : strides) {
			float best = 1e30f;
			for (int rep = 0; rep < 3; rep++) {
				CK(cudaEventRecord(e0));
				if (mode == 0) mb_kernel<0><<<sms * 4, 256>>>(d_out, span, s, iters, d_sink);
				if (mode == 1) mb_kernel<1><<<sms * 4, 256>>>(d_out, span, s, iters, d_sink);
				if (mode == 2) mb_kernel<2><<<sms * 4, 256>>>(d_out, span, s, iters, d_sink);
				CK(cudaEventRecord(e1));
				CK(cudaEventSynchronize(e1));
				CK(cudaGetLastError());
				float ms = time_ms(e0, e1);
				if (ms < best) best = ms;
			}
			// warp-steps per SM: 32 warps x iters; cycles at 1.965 GHz
			double cyc = best * 1e-3 * 1.965e9 / (32.0 * iters);
			double gbs = (double)sms * 4 * 256 * iters * 64 / best / 1e6;
			printf("mb %s lane stride %4u B: %7.3f ms  %6.1f SM-cycles per warp-step of 2 KB (%.0f GB/s)\n",
			    mode == 0 ? "STG.128 x4" : mode == 1 ? "LDG.128 x4" : "ST.256 x2 ", s, best, cyc, gbs);
		}
	}
	return 0;
}
