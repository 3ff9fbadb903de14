// throughput microbenchmark of the integer ops of the Prim update (per SM, warp-instructions per cycle)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 2048
template <int OP>
__global__ void k(uint32_t *out, uint32_t seed, long long *cyc) {
	uint32_t a[8], b[8];
#pragma unroll
	for (int i = 0; i < 8; ++i) { a[i] = seed * (threadIdx.x + i + 1); b[i] = seed + i * 77 + threadIdx.x; }
	uint32_t c = seed ^ 0x1234567u, d = seed * 3u;
	__syncthreads();
	long long t0 = clock64();
	for (int it = 0; it < ITER; ++it) {
#pragma unroll
		for (int i = 0; i < 8; ++i) {
			if (OP == 0) a[i] = __vabsdiffu4(a[i], c);
			if (OP == 1) a[i] = __dp4a(a[i], b[i], a[i]);
			if (OP == 3) a[i] = min(a[i], b[i] + 0u) ^ 0u, a[i] = min(a[i], c);
			if (OP == 4) asm("mad.lo.u32 %0, %1, 8192, %2;" : "=r"(a[i]) : "r"(a[i]), "r"(b[i]));
			if (OP == 5) { a[i] = max(a[i], max(b[i], c)); a[i] = min(a[i], min(d, b[i] + 1u)); }   // 2 x VIMNMX3, cannot collapse
			if (OP == 6) { a[i] = min(a[i], c); a[i] = max(a[i], d); }                            // 2 x VIMNMX
			if (OP == 7) a[i] = a[i] + b[i];
			if (OP == 8) a[i] = (a[i] & b[i]) | c;
			if (OP == 9) { uint32_t ad = __vabsdiffu4(b[i], c); uint32_t d2 = __dp4a(ad, ad, 0u); uint32_t mr = max(d2, max(d, c)); uint32_t kk; asm("mad.lo.u32 %0, %1, 8192, %2;" : "=r"(kk) : "r"(mr), "r"(b[i])); a[i] = min(a[i], kk); }
		}
		c += 0x01010101u;
		d ^= c;
	}
	long long t1 = clock64();
	uint32_t s = 0;
#pragma unroll
	for (int i = 0; i < 8; ++i) s += a[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char *name, int per_iter_ops, int nblk_per_sm, int nthr) {
	int nsm = 148;
	uint32_t *out; long long *cyc;
	cudaMalloc(&out, sizeof(uint32_t) * nsm * nblk_per_sm * nthr);
	cudaMalloc(&cyc, sizeof(long long) * nsm * nblk_per_sm);
	k<OP><<<nsm * nblk_per_sm, nthr>>>(out, 12345u, cyc);
	k<OP><<<nsm * nblk_per_sm, nthr>>>(out, 12345u, cyc);
	cudaDeviceSynchronize();
	long long h[148 * 8];
	cudaMemcpy(h, cyc, sizeof(long long) * nsm * nblk_per_sm, cudaMemcpyDeviceToHost);
	double avg = 0; for (int i = 0; i < nsm * nblk_per_sm; ++i) avg += h[i]; avg /= nsm * nblk_per_sm;
	double winst = (double)ITER * 8 * per_iter_ops * (nthr / 32) * nblk_per_sm;   // warp instrs per SM
	printf("%-34s warps/SM %2d: %.3f warp-instr/cycle/SM (%.2f cycles per warp-instr per SMSP)\n", name, nthr / 32 * nblk_per_sm, winst / avg, avg / (winst / 4));
	cudaFree(out); cudaFree(cyc);
}
int main() {
	for (int w = 0; w < 2; ++w) {
		int nthr = w == 0 ? 512 : 1024, nb = 1;
		run<0>("VABSDIFF4", 1, nb, nthr);
		run<1>("IDP.4A", 1, nb, nthr);
		run<4>("IMAD a*8192+b", 1, nb, nthr);
		run<5>("VIMNMX3 (max3 then min3)", 2, nb, nthr);
		run<6>("VIMNMX (min then max)", 2, nb, nthr);
		run<7>("IADD", 1, nb, nthr);
		run<8>("LOP3", 1, nb, nthr);
		run<9>("prim slot update (5 ops)", 5, nb, nthr);
	}
	return 0;
}
