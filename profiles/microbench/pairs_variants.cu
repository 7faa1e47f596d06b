// Schedule experiments for the FP64 all-pairs inner loop (evidence for DESIGN.md; not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o pairs_variants pairs_variants.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct alignas(32) body4 { double x, y, z, m; };
#define THREADS 128
#define TILE 128

__device__ __forceinline__ double seed_rsqrt(double x) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
// magic-number seed: integer pipe only (3 % accurate -- timing experiment, results are wrong on purpose)
__device__ __forceinline__ double seed_magic(double x)
{
	int hi = __double2hiint(x);
	return __hiloint2double(0x5fe6eb50 - (hi >> 1), 0);
}
// FP32 MUFU.RSQ route: rebias the double's exponent into a float with integer ops, rsqrtf, widen with integer ops
__device__ __forceinline__ double seed_f32(double x)
{
	unsigned hi = (unsigned)__double2hiint(x), lo = (unsigned)__double2loint(x);
	unsigned fb = __funnelshift_l(lo, hi - 0x38000000u, 3);       // 1 IADD + 1 SHF
	float y = rsqrtf(__uint_as_float(fb));                        // MUFU.RSQ
	unsigned yb = __float_as_uint(y);
	return __hiloint2double((int)((yb >> 3) + 0x38000000u), (int)(yb << 29));
}
__device__ __forceinline__ double clamp_int(double r2)
{
	long long b = __double_as_longlong(r2);
	const long long m = 0x3E45798EE2308C3ALL;
	b = b < m ? m : b;
	return __longlong_as_double(b);
}

// VAR bit 0: fmax clamp (DSETP+FSEL) instead of integer clamp
// VAR bit 1: staged (all targets advance phase by phase)
// VAR bit 2: seed from hi-word only via integer clamp on hi word (approximate clamp for the seed, exact for r2)
template<int IPT, int UNR, int MINB, int VAR>
__global__ void __launch_bounds__(THREADS, MINB) pairs(const body4* __restrict__ src, double* __restrict__ out, int n)
{
	__shared__ body4 tile[2][TILE];
	double xi[IPT], yi[IPT], zi[IPT], ax[IPT], ay[IPT], az[IPT];
	const int i0 = blockIdx.x * THREADS * IPT + threadIdx.x;
#pragma unroll
	for(int k = 0; k < IPT; ++k)
	{
		body4 b = src[(i0 + k * THREADS) % n];
		xi[k] = b.x; yi[k] = b.y; zi[k] = b.z; ax[k] = ay[k] = az[k] = 0;
	}
	const int nt = n / TILE;
	tile[0][threadIdx.x] = src[threadIdx.x];
	__syncthreads();
	for(int t = 0; t < nt; ++t)
	{
		const body4* tb = tile[t & 1];
		if(t + 1 < nt) { tile[(t + 1) & 1][threadIdx.x] = src[(t + 1) * TILE + threadIdx.x]; }
#pragma unroll UNR
		for(int j = 0; j < TILE; ++j)
		{
			const body4 s = tb[j];
			if(VAR & 2)
			{
				double dx[IPT], dy[IPT], dz[IPT], r2[IPT], y0[IPT], e[IPT], c[IPT];
#pragma unroll
				for(int k = 0; k < IPT; ++k) { dx[k] = s.x - xi[k]; dy[k] = s.y - yi[k]; dz[k] = s.z - zi[k]; }
#pragma unroll
				for(int k = 0; k < IPT; ++k) { r2[k] = fma(dz[k], dz[k], fma(dy[k], dy[k], dx[k] * dx[k])); }
#pragma unroll
				for(int k = 0; k < IPT; ++k) { r2[k] = (VAR & 1) ? fmax(r2[k], 1e-8) : clamp_int(r2[k]); }
#pragma unroll
				for(int k = 0; k < IPT; ++k) { y0[k] = seed_rsqrt(r2[k]); }
#pragma unroll
				for(int k = 0; k < IPT; ++k) { double h = r2[k] * y0[k]; e[k] = fma(-h, y0[k], 1.0); }
#pragma unroll
				for(int k = 0; k < IPT; ++k)
				{
					double p = fma(e[k], 0.375, 0.5), q = y0[k] * e[k];
					double y = fma(q, p, y0[k]);
					c[k] = (y * y) * (s.m * y);
				}
#pragma unroll
				for(int k = 0; k < IPT; ++k) { ax[k] = fma(dx[k], c[k], ax[k]); ay[k] = fma(dy[k], c[k], ay[k]); az[k] = fma(dz[k], c[k], az[k]); }
			}
			else
			{
#pragma unroll
				for(int k = 0; k < IPT; ++k)
				{
					double dx = s.x - xi[k], dy = s.y - yi[k], dz = s.z - zi[k];
					double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
					r2 = (VAR & 1) ? fmax(r2, 1e-8) : clamp_int(r2);
					double y0 = (VAR & 4) ? seed_magic(r2) : ((VAR & 8) ? seed_f32(r2) : seed_rsqrt(r2));
					double h = r2 * y0;
					double e = fma(-h, y0, 1.0);
					double p = fma(e, 0.375, 0.5);
					double q = y0 * e;
					double y = fma(q, p, y0);
					double c = (y * y) * (s.m * y);
					ax[k] = fma(dx, c, ax[k]); ay[k] = fma(dy, c, ay[k]); az[k] = fma(dz, c, az[k]);
				}
			}
		}
		__syncthreads();
	}
#pragma unroll
	for(int k = 0; k < IPT; ++k)
	{
		int i = i0 + k * THREADS;
		if(i < n) { out[i] = ax[k]; out[n + i] = ay[k]; out[2 * n + i] = az[k]; }
	}
}

// Fast path: no clamp in the loop; one ISETP per pair keeps a sticky "some r2 may be < MinDistance" flag;
// tile-local accumulators are merged when the flag is clear, else the tile is redone with the exact clamp.
// VAR bit 0: seed keeps a junk low word (no IMAD.MOV zeroing)
template<int IPT, int UNR, int MINB, int VAR>
__global__ void __launch_bounds__(THREADS, MINB) pairs_fast(const body4* __restrict__ src, double* __restrict__ out, int n)
{
	__shared__ body4 tile[2][TILE];
	double xi[IPT], yi[IPT], zi[IPT], ax[IPT], ay[IPT], az[IPT];
	const int i0 = blockIdx.x * THREADS * IPT + threadIdx.x;
#pragma unroll
	for(int k = 0; k < IPT; ++k)
	{
		body4 b = src[(i0 + k * THREADS) % n];
		xi[k] = b.x; yi[k] = b.y; zi[k] = b.z; ax[k] = ay[k] = az[k] = 0;
	}
	const int nt = n / TILE;
	tile[0][threadIdx.x] = src[threadIdx.x];
	__syncthreads();
	for(int t = 0; t < nt; ++t)
	{
		const body4* tb = tile[t & 1];
		if(t + 1 < nt) { tile[(t + 1) & 1][threadIdx.x] = src[(t + 1) * TILE + threadIdx.x]; }
		double tx[IPT], ty[IPT], tz[IPT];
#pragma unroll
		for(int k = 0; k < IPT; ++k) { tx[k] = ty[k] = tz[k] = 0; }
		int near = 0x7fffffff;   // running min of hi(r2)
#pragma unroll UNR
		for(int j = 0; j < TILE; ++j)
		{
			const body4 s = tb[j];
#pragma unroll
			for(int k = 0; k < IPT; ++k)
			{
				double dx = s.x - xi[k], dy = s.y - yi[k], dz = s.z - zi[k];
				double t1 = dx * dx;
				double t2 = fma(dy, dy, t1);
				double r2 = fma(dz, dz, t2);
				if(!(VAR & 2)) { int hi = __double2hiint(r2); near = min(near, hi); }
				double y0;
				if(VAR & 4) { y0 = seed_magic(r2); }
				else if(VAR & 1)
				{
					// y0 = {junk low word, MUFU.RSQ64H(hi)}: the low word only perturbs y0 by < 2^-20, which e absorbs
					double s0;
					asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(s0) : "d"(r2));
					y0 = __hiloint2double(__double2hiint(s0), __double2loint(t1));
				}
				else
				{
					y0 = seed_rsqrt(r2);
				}
				double h = r2 * y0;
				double e = fma(-h, y0, 1.0);
				double p = fma(e, 0.375, 0.5);
				double q = y0 * e;
				double y = fma(q, p, y0);
				double c = (y * y) * (s.m * y);
				tx[k] = fma(dx, c, tx[k]); ty[k] = fma(dy, c, ty[k]); tz[k] = fma(dz, c, tz[k]);
			}
		}
		if(near <= 0x3E45798E)   // hi word of 1e-8: rare (self pair / bodies closer than 1e-4)
		{
#pragma unroll
			for(int k = 0; k < IPT; ++k) { tx[k] = ty[k] = tz[k] = 0; }
#pragma unroll 1
			for(int j = 0; j < TILE; ++j)
			{
				const body4 s = tb[j];
#pragma unroll
				for(int k = 0; k < IPT; ++k)
				{
					double dx = s.x - xi[k], dy = s.y - yi[k], dz = s.z - zi[k];
					double r2 = clamp_int(fma(dz, dz, fma(dy, dy, dx * dx)));
					double y0 = seed_rsqrt(r2);
					double h = r2 * y0;
					double e = fma(-h, y0, 1.0);
					double p = fma(e, 0.375, 0.5);
					double q = y0 * e;
					double y = fma(q, p, y0);
					double c = (y * y) * (s.m * y);
					tx[k] = fma(dx, c, tx[k]); ty[k] = fma(dy, c, ty[k]); tz[k] = fma(dz, c, tz[k]);
				}
			}
		}
#pragma unroll
		for(int k = 0; k < IPT; ++k) { ax[k] += tx[k]; ay[k] += ty[k]; az[k] += tz[k]; }
		__syncthreads();
	}
#pragma unroll
	for(int k = 0; k < IPT; ++k)
	{
		int i = i0 + k * THREADS;
		if(i < n) { out[i] = ax[k]; out[n + i] = ay[k]; out[2 * n + i] = az[k]; }
	}
}

// ---- symmetric (Newton's third law) systolic warp tile --------------------------------------------------
// Each lane keeps I "A" bodies (+ accumulators) and J "B" bodies (+ accumulators). Per step the lane evaluates its
// I x J unordered pairs once (21 FP64 instr) and updates BOTH sides; then the B group (positions, mass, accumulators)
// moves to the next lane by shuffles. After 32 steps every A body has met every B body of the 32*J block.
__device__ __forceinline__ double shfl_d(double v, int src)
{
	int lo = __shfl_sync(0xffffffffu, __double2loint(v), src);
	int hi = __shfl_sync(0xffffffffu, __double2hiint(v), src);
	return __hiloint2double(hi, lo);
}

template<int I, int J>
__global__ void __launch_bounds__(THREADS) pairs_sym(const body4* __restrict__ src, double* __restrict__ out, int n, int nb_blocks_per_warp)
{
	const int lane = threadIdx.x & 31;
	const int warp = (blockIdx.x * THREADS + threadIdx.x) >> 5;
	const int a0 = (warp * 32 * I) % n;
	double xa[I], ya[I], za[I], ma[I], ax[I], ay[I], az[I];
#pragma unroll
	for(int k = 0; k < I; ++k)
	{
		body4 b = src[(a0 + k * 32 + lane) % n];
		xa[k] = b.x; ya[k] = b.y; za[k] = b.z; ma[k] = b.m; ax[k] = ay[k] = az[k] = 0;
	}
	const int next = (lane + 31) & 31;   // receive from lane-1
	for(int blk = 0; blk < nb_blocks_per_warp; ++blk)
	{
		const int b0 = ((warp * 7 + blk) * 32 * J) % n;
		double xb[J], yb[J], zb[J], mb[J], bx[J], by[J], bz[J];
#pragma unroll
		for(int q = 0; q < J; ++q)
		{
			body4 b = src[(b0 + q * 32 + lane) % n];
			xb[q] = b.x; yb[q] = b.y; zb[q] = b.z; mb[q] = b.m; bx[q] = by[q] = bz[q] = 0;
		}
#pragma unroll 1
		for(int step = 0; step < 32; ++step)
		{
#pragma unroll
			for(int q = 0; q < J; ++q)
			{
#pragma unroll
				for(int k = 0; k < I; ++k)
				{
					double dx = xb[q] - xa[k], dy = yb[q] - ya[k], dz = zb[q] - za[k];
					double r2 = clamp_int(fma(dz, dz, fma(dy, dy, dx * dx)));
					double y0 = seed_rsqrt(r2);
					double h = r2 * y0;
					double e = fma(-h, y0, 1.0);
					double p = fma(e, 0.375, 0.5);
					double qq = y0 * e;
					double y = fma(qq, p, y0);
					double y3 = (y * y) * y;
					double ca = mb[q] * y3, cb = ma[k] * y3;
					ax[k] = fma(dx, ca, ax[k]); ay[k] = fma(dy, ca, ay[k]); az[k] = fma(dz, ca, az[k]);
					bx[q] = fma(-dx, cb, bx[q]); by[q] = fma(-dy, cb, by[q]); bz[q] = fma(-dz, cb, bz[q]);
				}
			}
#pragma unroll
			for(int q = 0; q < J; ++q)
			{
				xb[q] = shfl_d(xb[q], next); yb[q] = shfl_d(yb[q], next); zb[q] = shfl_d(zb[q], next); mb[q] = shfl_d(mb[q], next);
				bx[q] = shfl_d(bx[q], next); by[q] = shfl_d(by[q], next); bz[q] = shfl_d(bz[q], next);
			}
		}
#pragma unroll
		for(int q = 0; q < J; ++q)
		{
			int j = (b0 + q * 32 + lane) % n;
			atomicAdd(out + j, bx[q]); atomicAdd(out + n + j, by[q]); atomicAdd(out + 2 * n + j, bz[q]);
		}
	}
#pragma unroll
	for(int k = 0; k < I; ++k)
	{
		int i = (a0 + k * 32 + lane) % n;
		atomicAdd(out + i, ax[k]); atomicAdd(out + n + i, ay[k]); atomicAdd(out + 2 * n + i, az[k]);
	}
}

template<int I, int J>
void run_sym(const char* name, const body4* src, double* out, int n)
{
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	// 148*8 CTAs of 4 warps; every warp meets `nbw` B blocks: unordered pairs = warps * (32 I) * (32 J) * nbw
	int grid = 148 * 8, nbw = 256;
	pairs_sym<I, J><<<grid, THREADS>>>(src, out, n, 4);
	cudaDeviceSynchronize();
	float best = 1e30f;
	for(int r = 0; r < 3; ++r)
	{
		cudaEventRecord(e0);
		pairs_sym<I, J><<<grid, THREADS>>>(src, out, n, nbw);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms;
		cudaEventElapsedTime(&ms, e0, e1);
		if(ms < best) { best = ms; }
	}
	cudaFuncAttributes fa;
	cudaFuncGetAttributes(&fa, pairs_sym<I, J>);
	double unordered = (double)grid * 4 * (32.0 * I) * (32.0 * J) * nbw;
	printf("%-40s regs %3d  %8.3f ms  %7.2f G unordered pairs/s = %7.2f G interactions/s\n", name, fa.numRegs, best,
		   unordered / best / 1e6, 2 * unordered / best / 1e6);
}

template<int IPT, int UNR, int MINB, int VAR>
void run_fast(const char* name, const body4* src, double* out, int n)
{
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	int grid = n / (THREADS * IPT);
	pairs_fast<IPT, UNR, MINB, VAR><<<grid, THREADS>>>(src, out, n);
	cudaDeviceSynchronize();
	float best = 1e30f;
	for(int r = 0; r < 3; ++r)
	{
		cudaEventRecord(e0);
		pairs_fast<IPT, UNR, MINB, VAR><<<grid, THREADS>>>(src, out, n);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms;
		cudaEventElapsedTime(&ms, e0, e1);
		if(ms < best) { best = ms; }
	}
	cudaFuncAttributes fa;
	cudaFuncGetAttributes(&fa, pairs_fast<IPT, UNR, MINB, VAR>);
	static double host[3];
	cudaMemcpy(host, out, sizeof(host), cudaMemcpyDeviceToHost);
	printf("%-40s regs %3d  %8.3f ms  %7.2f Gpairs/s  (a0x=%.15g)\n", name, fa.numRegs, best, (double)n * n / best / 1e6, host[0]);
}

template<int IPT, int UNR, int MINB, int VAR>
void run(const char* name, const body4* src, double* out, int n)
{
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	int grid = n / (THREADS * IPT);
	pairs<IPT, UNR, MINB, VAR><<<grid, THREADS>>>(src, out, n);
	cudaDeviceSynchronize();
	float best = 1e30f;
	for(int r = 0; r < 3; ++r)
	{
		cudaEventRecord(e0);
		pairs<IPT, UNR, MINB, VAR><<<grid, THREADS>>>(src, out, n);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms;
		cudaEventElapsedTime(&ms, e0, e1);
		if(ms < best) { best = ms; }
	}
	cudaFuncAttributes fa;
	cudaFuncGetAttributes(&fa, pairs<IPT, UNR, MINB, VAR>);
	double sum = 0;
	static double host[3];
	cudaMemcpy(host, out, sizeof(host), cudaMemcpyDeviceToHost);
	sum = host[0];
	printf("%-40s regs %3d  %8.3f ms  %7.2f Gpairs/s  (a0x=%.15g)\n", name, fa.numRegs, best, (double)n * n / best / 1e6, sum);
}

int main(int argc, char** argv)
{
	int n = 148 * 128 * 4 * 3;   // 227,328 bodies: a whole number of waves for every shape below
	body4* h = (body4*)malloc(n * sizeof(body4));
	unsigned long long st = 88172645463325252ULL;
	auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) / 9007199254740992.0; };
	for(int i = 0; i < n; ++i) { h[i].x = rnd() * 100; h[i].y = rnd() * 100; h[i].z = rnd() * 30; h[i].m = rnd(); }
	body4* d;
	double* out;
	cudaMalloc(&d, n * sizeof(body4));
	cudaMalloc(&out, 3 * n * sizeof(double));
	cudaMemcpy(d, h, n * sizeof(body4), cudaMemcpyHostToDevice);
	run<4, 4, 1, 0>("ipt4 unr4 intclamp (current)", d, out, n);
	run_sym<4, 4>("SYM I4 J4", d, out, n);
	run_sym<4, 2>("SYM I4 J2", d, out, n);
	run_sym<8, 2>("SYM I8 J2", d, out, n);
	run_sym<8, 1>("SYM I8 J1", d, out, n);
	run_sym<6, 2>("SYM I6 J2", d, out, n);
	run_sym<2, 2>("SYM I2 J2", d, out, n);
	run_fast<4, 4, 1, 0>("FAST ipt4 unr4", d, out, n);
	run_fast<4, 4, 1, 1>("FAST ipt4 unr4 junk-lo seed", d, out, n);
	run_fast<4, 4, 1, 3>("FAST ipt4 unr4 junk-lo NOFLAG", d, out, n);
	run_fast<4, 4, 1, 6>("FAST ipt4 unr4 MAGIC NOFLAG (17 fp64 + LDS only)", d, out, n);
	run_fast<4, 4, 1, 4>("FAST ipt4 unr4 MAGIC flag", d, out, n);
	run_fast<4, 2, 1, 1>("FAST ipt4 unr2 junk-lo seed", d, out, n);
	run_fast<4, 4, 4, 1>("FAST ipt4 unr4 junk-lo minb4", d, out, n);
	run_fast<4, 4, 5, 1>("FAST ipt4 unr4 junk-lo minb5", d, out, n);
	run_fast<2, 8, 1, 1>("FAST ipt2 unr8 junk-lo seed", d, out, n);
	run_fast<2, 4, 8, 1>("FAST ipt2 unr4 junk-lo minb8", d, out, n);
	run<4, 4, 1, 1>("ipt4 unr4 fmax", d, out, n);
	run<4, 4, 1, 4>("ipt4 unr4 MAGIC seed (no MUFU, wrong)", d, out, n);
	run<4, 4, 1, 8>("ipt4 unr4 f32 MUFU.RSQ seed", d, out, n);
	run<2, 8, 1, 8>("ipt2 unr8 f32 MUFU.RSQ seed", d, out, n);
	run<2, 8, 1, 4>("ipt2 unr8 MAGIC seed", d, out, n);
	run<4, 8, 1, 0>("ipt4 unr8 intclamp", d, out, n);
	run<4, 2, 1, 0>("ipt4 unr2 intclamp", d, out, n);
	run<4, 1, 1, 0>("ipt4 unr1 intclamp", d, out, n);
	run<4, 4, 1, 2>("ipt4 unr4 staged intclamp", d, out, n);
	run<4, 2, 1, 2>("ipt4 unr2 staged intclamp", d, out, n);
	run<4, 1, 1, 2>("ipt4 unr1 staged intclamp", d, out, n);
	run<4, 1, 1, 3>("ipt4 unr1 staged fmax", d, out, n);
	run<4, 4, 6, 0>("ipt4 unr4 minb6", d, out, n);
	run<4, 4, 4, 0>("ipt4 unr4 minb4", d, out, n);
	run<2, 4, 1, 0>("ipt2 unr4", d, out, n);
	run<2, 8, 1, 0>("ipt2 unr8", d, out, n);
	run<2, 4, 1, 2>("ipt2 unr4 staged", d, out, n);
	run<8, 1, 1, 0>("ipt8 unr1", d, out, n);
	run<8, 2, 1, 0>("ipt8 unr2", d, out, n);
	run<8, 1, 1, 2>("ipt8 unr1 staged", d, out, n);
	run<6, 2, 1, 0>("ipt6 unr2", d, out, n);
	run<6, 1, 1, 2>("ipt6 unr1 staged", d, out, n);
	return 0;
}
