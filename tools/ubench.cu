/* tools/ubench.cu -- developer aid: per-SM throughput of the instruction kinds the
 * render kernel is made of (FP64 add/mul, f32<->f64 and int->f64 conversions,
 * f32->s64 conversion, IEEE f32 division), measured on the GPU it runs on.
 * Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ubench ubench.cu
 * Output: one line per kind: warp-instructions per clock per SM. */
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNR 8

template <int KIND>
__global__ void k(double *out, float seed_f, double seed_d, long long *clk) {
	double d[UNR]; float f[UNR]; long long l[UNR]; int32_t n[UNR];
#pragma unroll
	for (int u = 0; u < UNR; ++u) {
		d[u] = seed_d + threadIdx.x * 1e-3 + u; f[u] = seed_f + threadIdx.x * 1e-3f + u;
		l[u] = 0; n[u] = threadIdx.x + u + 1;
	}
	long long t0 = clock64();
	for (int i = 0; i < ITERS; ++i) {
#pragma unroll
		for (int u = 0; u < UNR; ++u) {
			if (KIND == 0) d[u] = d[u] + seed_d;                    /* DADD */
			if (KIND == 1) d[u] = d[u] * seed_d;                    /* DMUL */
			if (KIND == 2) { d[u] = (double) f[u]; f[u] += __double_as_longlong(d[u]) & 1 ? 1.f : 2.f; }   /* F2F.F64.F32 + FADD */
			if (KIND == 3) { f[u] = (float) d[u]; d[u] = __longlong_as_double(__double_as_longlong(d[u]) + __float_as_int(f[u])); }  /* F2F.F32.F64 + IADD64 */
			if (KIND == 4) { d[u] = (double) n[u]; n[u] += (int32_t) __double2hiint(d[u]); }  /* I2F.F64 */
			if (KIND == 5) { l[u] = __float2ll_rn(f[u]); f[u] += (float) (int32_t) l[u]; }   /* F2I.S64 + I2F + FADD */
			if (KIND == 6) f[u] = seed_f / f[u];                    /* IEEE div.rn.f32 */
			if (KIND == 7) f[u] = f[u] + seed_f;                    /* FADD reference */
			if (KIND == 8) f[u] = __fdividef(seed_f, f[u]);         /* approx div for reference */
			if (KIND == 9) d[u] = __fma_rn(d[u], seed_d, seed_d);   /* DFMA */
			if (KIND == 10) { n[u] = __float2int_rn(f[u]); f[u] += (float) n[u]; } /* F2I.S32 + I2F + FADD */
		}
	}
	long long t1 = clock64();
	double acc = 0;
#pragma unroll
	for (int u = 0; u < UNR; ++u) acc += d[u] + f[u] + (double) l[u] + n[u];
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
	if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(const char *name, int extra_ops) {
	const int threads = 1024, blocks = 148;
	double *out; long long *clk;
	cudaMalloc(&out, sizeof(double) * threads * blocks);
	cudaMalloc(&clk, sizeof(long long) * blocks);
	k<KIND><<<blocks, threads>>>(out, 1.0001f, 1.0000001, clk);
	k<KIND><<<blocks, threads>>>(out, 1.0001f, 1.0000001, clk);
	cudaDeviceSynchronize();
	long long h[148];
	cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
	double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
	double winst = (double) ITERS * UNR * (threads / 32);
	printf("%-28s %8.3f warp-instr/clk/SM  (%.2f clk per warp-instr/SM; sequence has %d helper ops)\n",
			name, winst / avg, avg / winst, extra_ops);
	cudaFree(out); cudaFree(clk);
}

int main() {
	run<7>("FADD", 0);
	run<0>("DADD", 0);
	run<1>("DMUL", 0);
	run<9>("DFMA", 0);
	run<2>("F2F.F64.F32 (+sel,fadd)", 2);
	run<3>("F2F.F32.F64 (+iadd64)", 2);
	run<4>("I2F.F64.S32 (+iadd)", 1);
	run<5>("F2I.S64.F32 (+i2f,fadd)", 2);
	run<10>("F2I.S32.F32 (+i2f,fadd)", 2);
	run<6>("div.rn.f32", 0);
	run<8>("__fdividef", 0);
	return 0;
}
