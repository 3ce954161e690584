// Micro-benchmark: dependent-issue latency of the integer ops on the chain
// path and cycles/sample of alternative formulations of one predictor step.
// One warp, one block.  nvcc -arch=sm_100a -O3 -o build/lat_bench tools/lat_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP 4096

template <int OP>
__global__ void lat(int *out, long long *cyc, int seed, int k0)
{
	int x = seed + threadIdx.x, y = seed * 3;
	long long t0 = clock64();
#pragma unroll 1
	for (int i = 0; i < REP / 32; i++) {
#pragma unroll
		for (int j = 0; j < 32; j++) {
			if (OP == 0) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(k0), "r"(y));
			if (OP == 1) asm volatile("add.s32 %0, %0, %1;" : "+r"(x) : "r"(y));
			if (OP == 2) asm volatile("shr.s32 %0, %0, 1;" : "+r"(x));
			if (OP == 3) asm volatile("max.s32 %0, %0, %1;" : "+r"(x) : "r"(y));
			if (OP == 4) asm volatile("{.reg .pred p; setp.lt.s32 p, %0, %1; selp.s32 %0, %1, %2, p;}" : "+r"(x) : "r"(y), "r"(k0));
			if (OP == 5) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(x) : "r"(y));
			if (OP == 6) asm volatile("{.reg .s16 h; cvt.sat.s16.s32 h, %0; cvt.s32.s16 %0, h;}" : "+r"(x));
			if (OP == 7) asm volatile("cvt.pack.sat.s16.s32 %0, %0, %1;" : "+r"(x) : "r"(y));
			if (OP == 8) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(y), "r"(k0));
			if (OP == 9) asm volatile("mad.lo.s32 %0, %0, %1, %2;\n\tshr.s32 %0, %0, 1;" : "+r"(x) : "r"(k0), "r"(y));
		}
	}
	long long t1 = clock64();
	out[threadIdx.x] = x;
	if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// one predictor step, several formulations; returns new sample
template <int V>
__device__ __forceinline__ int step(int r8, int k0, int k1, int &p0, int &p1)
{
	int s;
	if (V == 0) {            // as the C reference writes it
		int g = p0 * k0 + p1 * k1;
		s = (r8 >> 8) + g / 256;
		s = max(-32768, min(32767, s));
	} else if (V == 1) {     // two shifted candidates + select
		int c = p1 * k1, a = r8 + c;
		int g = p0 * k0 + c, u = p0 * k0 + a, v = p0 * k0 + (a + 255);
		asm("shr.s32 %0, %0, 8;" : "+r"(u));
		asm("shr.s32 %0, %0, 8;" : "+r"(v));
		s = g < 0 ? v : u;
		s = max(-32768, min(32767, s));
	} else if (V == 2) {     // three explicit mads, select, clamp
		int c = p1 * k1, a = r8 + c, a2 = a + 255;
		int g, u, v;
		asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(g) : "r"(p0), "r"(k0), "r"(c));
		asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(u) : "r"(p0), "r"(k0), "r"(a));
		asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(v) : "r"(p0), "r"(k0), "r"(a2));
		asm("shr.s32 %0, %0, 8;" : "+r"(u));
		asm("shr.s32 %0, %0, 8;" : "+r"(v));
		asm("{.reg .pred p; setp.lt.s32 p, %1, 0; selp.s32 %0, %2, %3, p;}" : "=r"(s) : "r"(g), "r"(v), "r"(u));
		s = max(-32768, min(32767, s));
	} else if (V == 3) {     // V2 without the clamp (lower bound of a speculative scheme)
		int c = p1 * k1, a = r8 + c, a2 = a + 255;
		int g, u, v;
		asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(g) : "r"(p0), "r"(k0), "r"(c));
		asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(u) : "r"(p0), "r"(k0), "r"(a));
		asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(v) : "r"(p0), "r"(k0), "r"(a2));
		asm("shr.s32 %0, %0, 8;" : "+r"(u));
		asm("shr.s32 %0, %0, 8;" : "+r"(v));
		asm("{.reg .pred p; setp.lt.s32 p, %1, 0; selp.s32 %0, %2, %3, p;}" : "=r"(s) : "r"(g), "r"(v), "r"(u));
	} else if (V == 4) {     // float state: g' = g/4 exact in fp32, trunc by cvt.rzi
		float pf0 = __int2float_rn(p0), pf1 = __int2float_rn(p1);
		float g4 = fmaf(pf0, (float)(k0 >> 2), pf1 * (float)(k1 >> 2));
		int q = __float2int_rz(g4 * 0.015625f);
		s = (r8 >> 8) + q;
		s = max(-32768, min(32767, s));
	} else {                 // V5: clamp through pack.sat + sign-extending shift
		int c = p1 * k1, a = r8 + c, a2 = a + 255;
		int g, u, v;
		asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(g) : "r"(p0), "r"(k0), "r"(c));
		asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(u) : "r"(p0), "r"(k0), "r"(a));
		asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(v) : "r"(p0), "r"(k0), "r"(a2));
		asm("shr.s32 %0, %0, 8;" : "+r"(u));
		asm("shr.s32 %0, %0, 8;" : "+r"(v));
		asm("{.reg .pred p; setp.lt.s32 p, %1, 0; selp.s32 %0, %2, %3, p;}" : "=r"(s) : "r"(g), "r"(v), "r"(u));
		asm("cvt.pack.sat.s16.s32 %0, %0, %0;" : "+r"(s));
		asm("shr.s32 %0, %0, 16;" : "+r"(s));
	}
	p1 = p0;
	p0 = s;
	return s;
}

template <int V, int ILP>
__global__ void chain(const int *in, int *out, long long *cyc, int k0, int k1)
{
	int p0[ILP], p1[ILP];
	for (int c = 0; c < ILP; c++) { p0[c] = threadIdx.x + c; p1[c] = -c; }
	int acc = 0;
	long long t0 = clock64();
#pragma unroll 1
	for (int i = 0; i < REP / 32; i++) {
		int base = in[(i & 7) * 32 + (threadIdx.x & 31)];
#pragma unroll
		for (int j = 0; j < 32; j++) {
#pragma unroll
			for (int c = 0; c < ILP; c++)
				acc ^= step<V>((base + j * 977 + c * 31) << 8, k0, k1, p0[c], p1[c]);
		}
	}
	long long t1 = clock64();
	out[threadIdx.x] = acc;
	if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main()
{
	int *d_in, *d_out; long long *d_c, h;
	cudaMalloc(&d_in, 4096); cudaMalloc(&d_out, 4096); cudaMalloc(&d_c, 8);
	cudaMemset(d_in, 1, 4096);
	const char *names[] = {"IMAD", "IADD", "SHF", "VIMNMX", "ISETP+SEL", "PRMT", "cvt.sat.s16 (+sext)", "cvt.pack.sat (I2IP)", "LOP3", "IMAD+SHF"};
#define RUN_LAT(OP) lat<OP><<<1, 32>>>(d_out, d_c, 5, 3); lat<OP><<<1, 32>>>(d_out, d_c, 5, 3); cudaMemcpy(&h, d_c, 8, cudaMemcpyDeviceToHost); printf("lat %-22s %6.2f cycles/op\n", names[OP], (double)h / REP);
	RUN_LAT(0) RUN_LAT(1) RUN_LAT(2) RUN_LAT(3) RUN_LAT(4) RUN_LAT(5) RUN_LAT(6) RUN_LAT(7) RUN_LAT(8) RUN_LAT(9)
#define RUN_CH(V, ILP) chain<V, ILP><<<1, 32>>>(d_in, d_out, d_c, 460, -208); chain<V, ILP><<<1, 32>>>(d_in, d_out, d_c, 460, -208); cudaMemcpy(&h, d_c, 8, cudaMemcpyDeviceToHost); printf("chain variant %d ilp %d : %6.2f cycles/sample-step (%6.2f per sample)\n", V, ILP, (double)h / REP, (double)h / REP / ILP);
	RUN_CH(0, 1) RUN_CH(1, 1) RUN_CH(2, 1) RUN_CH(3, 1) RUN_CH(4, 1) RUN_CH(5, 1)
	RUN_CH(0, 2) RUN_CH(1, 2) RUN_CH(2, 2) RUN_CH(2, 4) RUN_CH(5, 2)
	cudaError_t e = cudaDeviceSynchronize();
	printf("status %s\n", cudaGetErrorString(e));
	return 0;
}
