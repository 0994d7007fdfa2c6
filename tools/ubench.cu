// Integer-pipe microbenchmark for B200 (sm_100a): which of the instructions the
// fixed-point butterfly can be built from issue at what rate, alone and mixed.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench tools/ubench.cu
// Output: warp-instructions per clock per SM for every kernel (CUDA-event timed).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 512
#define REP 16
#define CH 8   // independent chains per thread

enum Op { IMAD, IMADHI, IMADWIDE, SHFR, PRMTS, LOP, IADD, LEAHI, VIADD2, MIX_IMAD_LOP, MIX_IMAD_LEA, MIX_HI_LEA, MIX_HI_IMAD, MIX3, NOPS };
static const char *names[] = { "IMAD", "IMAD.HI", "IMAD.WIDE", "SHF.R", "PRMT", "LOP3", "IADD3", "LEA.HI.SX32", "VIADD.16x2",
	"IMAD+LOP3 (1:1)", "IMAD+LEA.HI (1:1)", "IMAD.HI+LEA.HI (1:1)", "IMAD.HI+IMAD (1:1)", "IMAD+LOP3+IMAD.HI (1:1:1)" };

template <int OP>
__global__ void __launch_bounds__(1024) k(int *out, int a, int b, int c)
{
	int x[CH];
	long long w[CH];
#pragma unroll
	for (int i = 0; i < CH; i++) { x[i] = threadIdx.x * (i + 1) + a; w[i] = x[i]; }
	const long long c64 = ((long long)c << 32) | 0x80000000ll;
#pragma unroll 1
	for (int it = 0; it < ITER; it++) {
#pragma unroll
		for (int rep = 0; rep < REP; rep++)
#pragma unroll
		for (int i = 0; i < CH; i++) {
			if (OP == IMAD) x[i] = x[i] * a + b;
			else if (OP == IMADHI) asm volatile("mad.hi.s32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(c));
			else if (OP == IMADWIDE) asm volatile("{ .reg .b32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.s32 %0, lo, %1, %0; }" : "+l"(w[i]) : "r"(a));
			else if (OP == SHFR) x[i] = __funnelshift_r(x[i], b, a);
			else if (OP == PRMTS) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(a));
			else if (OP == LOP) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b));
			else if (OP == IADD) x[i] = x[i] + x[(i + 1) % CH] + a;
			else if (OP == LEAHI) x[i] = (x[i] >> 15) + b;
			else if (OP == VIADD2) asm volatile("add.s16x2 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
			else if (OP == MIX_IMAD_LOP) { if (i & 1) x[i] = x[i] * a + b; else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b)); }
			else if (OP == MIX_IMAD_LEA) { if (i & 1) x[i] = x[i] * a + b; else x[i] = (x[i] >> 15) + b; }
			else if (OP == MIX_HI_LEA) { if (i & 1) asm volatile("mad.hi.s32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(c)); else x[i] = (x[i] >> 15) + b; }
			else if (OP == MIX_HI_IMAD) { if (i & 1) asm volatile("mad.hi.s32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(c)); else x[i] = x[i] * a + b; }
			else if (OP == MIX3) { if (i % 3 == 0) x[i] = x[i] * a + b; else if (i % 3 == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b)); else asm volatile("mad.hi.s32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(c)); }
		}
	}
	int s = 0;
#pragma unroll
	for (int i = 0; i < CH; i++) s += x[i] + (int)w[i];
	if (s == 0x7fffffff) out[threadIdx.x] = s;
}

template <int OP>
void run(int *d, int sms, double clk_ghz)
{
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	const int blocks = sms * 2;
	k<OP><<<blocks, 1024>>>(d, 3, 5, 7);
	cudaDeviceSynchronize();
	float best = 1e30f;
	for (int r = 0; r < 5; r++) {
		cudaEventRecord(e0);
		k<OP><<<blocks, 1024>>>(d, 3, 5, 7);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		if (ms < best) best = ms;
	}
	double warp_instr = (double)blocks * 32 * ITER * REP * CH;   // warps * iters * chains
	double per_clk_sm = warp_instr / (best * 1e-3) / (clk_ghz * 1e9) / sms;
	printf("%-28s %8.3f ms  %6.2f warp-instr/clk/SM (at %.3f GHz)  %7.1f G lane-op/s\n", names[OP], best, per_clk_sm, clk_ghz,
	       warp_instr * 32 / (best * 1e-3) / 1e9);
}

int main()
{
	cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
	int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
	double ghz = clk_khz / 1e6;
	printf("%s, %d SMs, max clock %.3f GHz (rates assume max clock)\n", p.name, p.multiProcessorCount, ghz);
	int *d; cudaMalloc(&d, 4096);
	int sms = p.multiProcessorCount;
	run<IMAD>(d, sms, ghz); run<IMADHI>(d, sms, ghz); run<IMADWIDE>(d, sms, ghz); run<SHFR>(d, sms, ghz);
	run<PRMTS>(d, sms, ghz); run<LOP>(d, sms, ghz); run<IADD>(d, sms, ghz); run<LEAHI>(d, sms, ghz); run<VIADD2>(d, sms, ghz);
	run<MIX_IMAD_LOP>(d, sms, ghz); run<MIX_IMAD_LEA>(d, sms, ghz); run<MIX_HI_LEA>(d, sms, ghz); run<MIX_HI_IMAD>(d, sms, ghz); run<MIX3>(d, sms, ghz);
	printf("%s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
