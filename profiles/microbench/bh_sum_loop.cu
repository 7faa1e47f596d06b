// Microbenchmark: the sum loop of the grouped Barnes-Hut walk (nb200_bh_group.cuh: bhg_flush / bhg_force) in isolation.
// What does one SM sub-partition sustain for this exact instruction mix (16 FP64 + MUFU.RSQ64H + 2 LDS.128 + mask
// select per entry) as a function of resident warps and of the unroll depth? Evidence for DESIGN.md 3.4; not part of the
// product library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bh_sum_loop bh_sum_loop.cu && ./bh_sum_loop
#include <cstdio>
#include <cuda_runtime.h>

struct alignas(32) body4 { double x, y, z, m; };

template<bool CLAMP>
__device__ __forceinline__ void force(double dx, double dy, double dz, double m, unsigned mine, double& ax, double& ay, double& az)
{
	double d2 = fma(dz, dz, fma(dx, dx, __dmul_rn(dy, dy)));
	m = mine != 0 ? m : 0.0;
	if(CLAMP)
	{
		long long bits = __double_as_longlong(d2);
		bits = bits < 0x3E45798EE2308C3ALL ? 0x3E45798EE2308C3ALL : bits;
		d2 = __longlong_as_double(bits);
	}
	double y0;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d2));
	const double y2 = y0 * y0;
	const double e = fma(-d2, y2, 1.0);
	const double u = e * fma(e, 1.875, 1.5);
	const double g = (m * y0) * y2;
	const double c = fma(g, u, g);
	ax = fma(-dx, c, ax);
	ay = fma(-dy, c, ay);
	az = fma(-dz, c, az);
}

#define LIST 64
template<int UNROLL, int WARPS, bool SELECT>
__global__ void __launch_bounds__(32 * WARPS) k(double* out, int rounds, int count, unsigned maskbits)
{
	__shared__ body4	lnode[WARPS][LIST];
	__shared__ unsigned	lmask[WARPS][LIST];
	const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for(int e = lane; e < LIST; e += 32)
	{
		body4 b;
		b.x = 3.0 + e * 0.37; b.y = -2.0 + e * 0.11; b.z = 1.0 + e * 0.05; b.m = 1e-3;
		lnode[w][e] = b;
		lmask[w][e] = maskbits;
	}
	__syncwarp();
	const double px = lane * 0.01, py = lane * 0.02, pz = lane * 0.03;
	const unsigned lane_bit = 1u << lane;
	double ax = 0, ay = 0, az = 0;
	for(int r = 0; r < rounds; ++r)
	{
		constexpr int in_flight = UNROLL;
#pragma unroll in_flight
		for(int e = 0; e < count; ++e)
		{
			const unsigned	mask = lmask[w][e];
			const body4		nd = lnode[w][e];
			force<false>(px - nd.x, py - nd.y, pz - nd.z, nd.m, SELECT ? (mask & lane_bit) : 1u, ax, ay, az);
		}
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = ax + ay + az;
}

template<int UNROLL, int WARPS, bool SELECT>
void run(int ctas_per_sm, double* out)
{
	const int rounds = 2000, count = 40;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	const int grid = 148 * ctas_per_sm;
	k<UNROLL, WARPS, SELECT><<<grid, 32 * WARPS>>>(out, 10, count, 0xffffffffu);
	cudaEventRecord(e0);
	k<UNROLL, WARPS, SELECT><<<grid, 32 * WARPS>>>(out, rounds, count, 0xffffffffu);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	const double entries = double(grid) * WARPS * rounds * count;
	const double cycles_per_entry_smsp = ms * 1e-3 * 1.965e9 * 148 * 4 / entries;
	printf("unroll %d warps/SM %3d select %d: %.2f cycles per entry and SM sub-partition (FP64 pipe alone: 32), %.1f %% of the pipe\n",
		   UNROLL, WARPS * ctas_per_sm, int(SELECT), cycles_per_entry_smsp, 100.0 * 32.0 / cycles_per_entry_smsp);
}

int main()
{
	double* out;
	cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024);
	run<4, 4, true>(1, out);
	run<4, 4, true>(2, out);
	run<4, 4, true>(3, out);
	run<4, 4, true>(4, out);
	run<4, 4, true>(5, out);
	run<4, 4, true>(6, out);
	run<4, 4, true>(8, out);
	run<2, 4, true>(5, out);
	run<8, 4, true>(5, out);
	run<1, 4, true>(5, out);
	run<4, 4, false>(5, out);
	run<8, 4, false>(5, out);
	cudaDeviceSynchronize();
	printf("%s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
