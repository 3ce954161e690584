// Micro-benchmark: cycles per sample of the production chain step
// (xa_core.h: decode_block_chain) for ONE warp alone on an SM, payload in
// registers, nothing else in the loop -- the floor of a stream without cut blocks.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/step_bench tools/step_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../bjxa_b200/csrc/xa_core.h"

using namespace xa;

template <int BITS, int VAR>
__global__ void run(const uint32_t *in, uint32_t *out, long long *cyc, int blocks, uint32_t prof)
{
	uint32_t pw[BITS];
	for (int i = 0; i < BITS; i++)
		pw[i] = in[threadIdx.x * BITS + i];
	int p0 = threadIdx.x, p1 = -3;
	uint32_t acc = 0;
	long long g0, g1;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
	long long t0 = clock64();
#pragma unroll 1
	for (int b = 0; b < blocks; b++) {
		uint32_t o[16];
		if (VAR == 0) {
			decode_block_chain<BITS>(o, pw, prof, p0, p1);
		} else {
			// the step with everything but the chain hoisted: codes ranged up front
			const int sh = 16 + (int)(prof & 15u);
			const int k0 = gain_k0(prof >> 4), k1 = gain_k1(prof >> 4);
			int xs[32];
			const int mul = 1 << (32 - sh);
#pragma unroll
			for (int i = 0; i < 32; i++) {
				if (VAR == 3)	/* ranged and biased in one multiply-add: hi(x * 2^(32-sh)) + 32768 */
					asm("mad.hi.s32 %0, %1, %2, 32768;" : "=r"(xs[i]) : "r"(top_code<BITS>(pw, i)), "r"(mul));
				else if (VAR == 4)	/* the bias added on the FMA pipe */
					asm("mad.lo.s32 %0, %1, 1, 32768;" : "=r"(xs[i]) : "r"(top_code<BITS>(pw, i) >> sh));
				else
					xs[i] = (top_code<BITS>(pw, i) >> sh) + 32768;
			}
			int b0 = p0 + 32768, b1 = p1 + 32768;
			const int c = chain_bias_c(k0, k1);
#pragma unroll
			for (int i = 0; i < 32; i++) {
				int g = b0 * k0 + (b1 * k1 + c);
				int q;
				if (VAR == 1 || VAR == 3 || VAR == 4) {
					int f;
					asm("mad.lo.s32 %0, %1, -255, %2;" : "=r"(f) : "r"(g >> 31), "r"(g));
					q = f >> 8;
				} else {
					q = (g + ((g >> 31) & 255)) >> 8;
				}
				int s = __vimin_s32_relu(q + xs[i], 65535);
				b1 = b0;
				b0 = s;
				if (i & 1)
					o[i >> 1] = (uint32_t)b1 | (uint32_t)b0 << 16;
			}
			p0 = b0 - 32768;
			p1 = b1 - 32768;
		}
#pragma unroll
		for (int i = 0; i < 16; i++)
			acc ^= o[i];
		pw[b & (BITS - 1)] ^= acc & 0x01010101u;
	}
	long long t1 = clock64();
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
	out[threadIdx.x] = acc ^ (uint32_t)p0;
	if (threadIdx.x == 0 && blockIdx.x == 0) {
		cyc[0] = t1 - t0;
		cyc[1] = g1 - g0;
	}
}

int main()
{
	uint32_t *d_in, *d_out; long long *d_c, h, hh[2];
	cudaMalloc(&d_in, 4096); cudaMalloc(&d_out, 4096); cudaMalloc(&d_c, 16);
	cudaMemset(d_in, 0x5a, 4096);
	const int blocks = 2048;
#define RUN(BITS, VAR, WARPS, name) \
	run<BITS, VAR><<<1, 32 * WARPS>>>(d_in, d_out, d_c, blocks, 0x23); \
	run<BITS, VAR><<<1, 32 * WARPS>>>(d_in, d_out, d_c, blocks, 0x23); \
	cudaMemcpy(&h, d_c, 8, cudaMemcpyDeviceToHost); \
	printf("%-46s bits %d warps/SM %2d: %6.2f cycles/sample per warp\n", name, BITS, WARPS, (double)h / blocks / 32);
	RUN(8, 4, 1, "ranged codes, +32768 by mad.lo, mad bias")
	RUN(8, 4, 8, "ranged codes, +32768 by mad.lo, mad bias")
	RUN(8, 4, 24, "ranged codes, +32768 by mad.lo, mad bias")
	RUN(8, 4, 32, "ranged codes, +32768 by mad.lo, mad bias")
	RUN(4, 4, 24, "ranged codes, +32768 by mad.lo, mad bias")
	RUN(6, 4, 24, "ranged codes, +32768 by mad.lo, mad bias")
	RUN(8, 1, 24, "ranged codes hoisted, mad bias")
	RUN(6, 1, 24, "ranged codes hoisted, mad bias")
	RUN(6, 0, 24, "decode_block_chain (production)")
	RUN(8, 0, 24, "decode_block_chain (production)")
	/* the SM clock a lone warp sees, and a full device: cycles / elapsed ns */
	for (int grid = 1; grid <= 148; grid *= 148) {
		for (int rep = 0; rep < 3; rep++) {
			run<8, 1><<<grid, grid == 1 ? 32 : 1024>>>(d_in, d_out, d_c, 65536, 0x23);
			cudaMemcpy(hh, d_c, 16, cudaMemcpyDeviceToHost);
			printf("grid %3d: %lld cycles in %lld ns = %.0f MHz\n", grid, hh[0], hh[1], 1e3 * hh[0] / hh[1]);
		}
	}
	cudaError_t e = cudaDeviceSynchronize();
	printf("status %s\n", cudaGetErrorString(e));
	return 0;
}
