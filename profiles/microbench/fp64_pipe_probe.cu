// Microbenchmark: what does the B200 FP64 pipe sustain for the operand patterns of the all-pairs kernel?
// (evidence for DESIGN.md's roofline discussion; not part of the product library)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipe_probe fp64_pipe_probe.cu && ./fp64_pipe_probe
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
#define UNROLL 16

template<int MODE>
__global__ void __launch_bounds__(256) k(double* out, int iters, double seed)
{
	double a[CHAINS], b[CHAINS], c[CHAINS];
#pragma unroll
	for(int i = 0; i < CHAINS; ++i)
	{
		a[i] = seed + i + threadIdx.x * 1e-9;
		b[i] = 0.999999 + i * 1e-9;
		c[i] = 1e-6 * (i + 1);
	}
	double m = 0.999999, cc = 1e-6;
	float fx = 1.5f + threadIdx.x;
	for(int it = 0; it < iters; ++it)
	{
#pragma unroll
		for(int u = 0; u < UNROLL; ++u)
		{
#pragma unroll
			for(int i = 0; i < CHAINS; ++i)
			{
				if(MODE == 0) { asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(m), "d"(cc)); }        // shared b, c
				if(MODE == 1) { asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(b[i]), "d"(c[i])); }    // 3 distinct
				if(MODE == 2) { asm volatile("fma.rn.f64 %0, %1, %1, %0;" : "+d"(a[i]) : "d"(b[i])); }                // 2 distinct
				if(MODE == 3) { asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(b[i])); }                    // DMUL
				if(MODE == 4) { asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(c[i])); }                    // DADD
				if(MODE == 5) { asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a[i]) : "d"(b[i]), "d"(m)); }        // acc form, shared coefficient
				if(MODE == 6)                                                                                           // 16 DFMA(3 distinct) : 1 MUFU.RSQ64H
				{
					asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(b[i]), "d"(c[i]));
					if(i == 0 && (u & 1) == 0) { asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(b[7])); }
				}
				if(MODE >= 8 && MODE <= 13)
				{
					asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(b[i]), "d"(c[i]));
					if(MODE == 8 && i == 0 && (u & 1) == 0) { asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(fx)); }      // MUFU.RSQ f32, 1 per 16
					if(MODE == 9 && i == 0) { asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(b[7])); }                    // RSQ64H, 1 per 8
					if(MODE == 10 && i == 0 && (u & 1) == 0) { asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(fx) : "d"(b[6])); asm volatile("" : "+f"(fx)); }  // F2F.F32.F64 1 per 16
					if(MODE == 11 && i == 0 && (u & 1) == 0) { asm volatile("cvt.f64.f32 %0, %1;" : "=d"(b[6]) : "f"(fx)); }   // F2F.F64.F32 1 per 16
					if(MODE == 12 && i == 0 && (u & 1) == 0) { asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(b[7]) : "d"(c[7])); } // RSQ64H independent input
					if(MODE == 13 && i == 0 && (u & 3) == 0) { asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(b[7])); }    // RSQ64H 1 per 32
				}
				if(MODE == 7)                                                                                           // DFMA + 4 ALU ops per 17 (clamp)
				{
					asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(m), "d"(cc));
				}
			}
			if(MODE == 7)
			{
				long long bits = __double_as_longlong(a[u % CHAINS]);
				bits = bits < 0x3E45798EE2308C3ALL ? 0x3E45798EE2308C3ALL : bits;
				a[u % CHAINS] = __longlong_as_double(bits);
			}
		}
	}
	double s = 0;
#pragma unroll
	for(int i = 0; i < CHAINS; ++i) { s += a[i] + b[i]; }
	if(s + fx == -12345.678) { out[0] = s; }
}

template<int MODE>
void run(const char* name, double* d, int sms)
{
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	int blocks = sms * 8, iters = 20000;
	k<MODE><<<blocks, 256>>>(d, 100, 1.0);
	cudaDeviceSynchronize();
	float best = 1e30f;
	for(int r = 0; r < 3; ++r)
	{
		cudaEventRecord(e0);
		k<MODE><<<blocks, 256>>>(d, iters, 1.0);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms;
		cudaEventElapsedTime(&ms, e0, e1);
		if(ms < best) { best = ms; }
	}
	double ops = (double)blocks * 256 * iters * UNROLL * CHAINS;
	double rate = ops / (best * 1e-3);
	// lanes per clock per SM at 1.965 GHz
	printf("%-44s %8.3f ms  %7.3f Tinstr-lane/s  %6.2f lanes/clk/SM @1965MHz\n", name, best, rate / 1e12, rate / sms / 1.965e9);
}

int main()
{
	cudaDeviceProp p;
	cudaGetDeviceProperties(&p, 0);
	double* d;
	cudaMalloc(&d, 64);
	printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
	run<0>("DFMA a=a*m+c (shared m,c)", d, p.multiProcessorCount);
	run<1>("DFMA a=a*b+c (3 distinct reg pairs)", d, p.multiProcessorCount);
	run<2>("DFMA a=b*b+a (2 distinct)", d, p.multiProcessorCount);
	run<3>("DMUL a=a*b", d, p.multiProcessorCount);
	run<4>("DADD a=a+c", d, p.multiProcessorCount);
	run<5>("DFMA a=b*m+a (shared m)", d, p.multiProcessorCount);
	run<6>("DFMA 3-distinct + 1 MUFU.RSQ64H per 16", d, p.multiProcessorCount);
	run<7>("DFMA shared + 64-bit int max per 8", d, p.multiProcessorCount);
	run<8>("DFMA + 1 MUFU.RSQ(f32) per 16", d, p.multiProcessorCount);
	run<9>("DFMA + 1 MUFU.RSQ64H per 8", d, p.multiProcessorCount);
	run<13>("DFMA + 1 MUFU.RSQ64H per 32", d, p.multiProcessorCount);
	run<12>("DFMA + 1 MUFU.RSQ64H per 16 (indep input)", d, p.multiProcessorCount);
	run<10>("DFMA + 1 F2F.F32.F64 per 16", d, p.multiProcessorCount);
	run<11>("DFMA + 1 F2F.F64.F32 per 16", d, p.multiProcessorCount);
	return 0;
}
