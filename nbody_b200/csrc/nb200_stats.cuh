// nb200 -- conservation sums of nbody_data::print_statistics on the device.
//
// The reference computes the --check_step report on the host after a full get_data(): total impulse, impulse moment,
// kinetic energy, mass centre (O(N)) and, with 'E' in the check list, the potential energy as a single-threaded Kahan
// sum over all N^2 ordered pairs (nbody/nbody_data.cpp:57-103, summation_proxy.h:68-87) -- at N >= 65k that one host
// loop dwarfs the GPU step. Here the same quantities are produced from the state vector in HBM:
//   stats_linear   P = sum m v, L = sum r x (m v), 2*Ekin = sum m |v|^2, sum m r, sum m      (one pass, FP64 partials)
//   stats_potential  U_i = sum_{j != i, r2 >= MinDistance} m_j / r  per target, all-pairs tiles like direct_pairs
//   stats_finish   fixed-order reduction of the per-block partials -> 12 doubles
// Accumulation is FP64 in both builds; block partials are combined in a fixed order (bit-reproducible).
#ifndef NB200_STATS_CUH
#define NB200_STATS_CUH

#include "nb200_common.cuh"
#include "nb200_direct.cuh"

#define NB200_STATS_THREADS 256
#define NB200_STATS_LINEAR 11	// Px Py Pz Lx Ly Lz 2Ekin Cx Cy Cz M

__device__ __forceinline__ double stats_block_sum(double v, double* scratch)
{
#pragma unroll
	for(int o = 16; o > 0; o >>= 1)
	{
		v += __shfl_xor_sync(0xffffffffu, v, o);
	}
	__syncthreads();
	if((threadIdx.x & 31) == 0) { scratch[threadIdx.x >> 5] = v; }
	__syncthreads();
	double total = 0;
	if(threadIdx.x == 0)
	{
		for(int w = 0; w < NB200_STATS_THREADS / 32; ++w) { total += scratch[w]; }
	}
	return total;	// valid in thread 0
}

// partial[block][11] over the local shard (y = 6 x n_shard rows, mass indexed globally)
__global__ void __launch_bounds__(NB200_STATS_THREADS) stats_linear(const real* __restrict__ y, const real* __restrict__ mass,
																	 size_t n_shard, size_t shard_first, double* __restrict__ partial)
{
	__shared__ double scratch[NB200_STATS_THREADS / 32];
	double acc[NB200_STATS_LINEAR];
#pragma unroll
	for(int q = 0; q < NB200_STATS_LINEAR; ++q) { acc[q] = 0; }
	for(size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_shard; i += static_cast<size_t>(gridDim.x) * blockDim.x)
	{
		const double m = mass[shard_first + i];
		const double rx = y[i], ry = y[n_shard + i], rz = y[2 * n_shard + i];
		const double vx = y[3 * n_shard + i], vy = y[4 * n_shard + i], vz = y[5 * n_shard + i];
		const double px = m * vx, py = m * vy, pz = m * vz;
		acc[0] += px; acc[1] += py; acc[2] += pz;
		acc[3] += ry * pz - rz * py; acc[4] += rz * px - rx * pz; acc[5] += rx * py - ry * px;
		acc[6] += (vx * vx + vy * vy + vz * vz) * m;
		acc[7] += m * rx; acc[8] += m * ry; acc[9] += m * rz;
		acc[10] += m;
	}
#pragma unroll
	for(int q = 0; q < NB200_STATS_LINEAR; ++q)
	{
		double total = stats_block_sum(acc[q], scratch);
		if(threadIdx.x == 0) { partial[static_cast<size_t>(blockIdx.x) * NB200_STATS_LINEAR + q] = total; }
	}
}

// partial[block] = sum over this block's targets of m_i * sum_{j} m_j / r_ij   (pairs with r2 < MinDistance give 0,
// as nbody_data::potential_energy does, nbody_data.cpp:46-55; that also removes j == i)
__global__ void __launch_bounds__(NB200_STATS_THREADS) stats_potential(const body4* __restrict__ src, size_t n_shard, size_t shard_first,
																		int n_tiles, double* __restrict__ partial)
{
	__shared__ body4 tile[NB200_DIRECT_TILE];
	__shared__ double scratch[NB200_STATS_THREADS / 32];
	const size_t	i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	const size_t	ic = i < n_shard ? i : n_shard - 1;
	const body4		me = src[shard_first + ic];
	const double	xi = me.x, yi = me.y, zi = me.z;
	double			u = 0;
	for(int t = 0; t < n_tiles; ++t)
	{
		__syncthreads();
		if(threadIdx.x < NB200_DIRECT_TILE) { tile[threadIdx.x] = src[static_cast<size_t>(t) * NB200_DIRECT_TILE + threadIdx.x]; }
		__syncthreads();
#pragma unroll 4
		for(int j = 0; j < NB200_DIRECT_TILE; ++j)
		{
			const body4	s = tile[j];
			double dx = static_cast<double>(s.x) - xi, dy = static_cast<double>(s.y) - yi, dz = static_cast<double>(s.z) - zi;
			double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
			// 1/sqrt(r2) by the same seed + cubic correction as the force kernel; masked where r2 < MinDistance
			const bool	close = r2 < static_cast<double>(NB200_MIN_DISTANCE);
			double		safe = close ? 1.0 : r2;
			double		y0;
			asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(safe));
			double h = safe * y0;
			double e = fma(-h, y0, 1.0);
			double p = fma(e, 0.375, 0.5);
			double q = y0 * e;
			double yv = fma(q, p, y0);
			u = fma(close ? 0.0 : static_cast<double>(s.m), yv, u);
		}
	}
	double mine = i < n_shard ? u * static_cast<double>(me.m) : 0.0;
	double total = stats_block_sum(mine, scratch);
	if(threadIdx.x == 0) { partial[blockIdx.x] = total; }
}

// out[0..10] = fixed-order sums of the linear partials; out[11] = sum of the potential partials (0 if none)
__global__ void stats_finish(const double* __restrict__ lin, int lin_blocks, const double* __restrict__ pot, int pot_blocks,
							 double* __restrict__ out)
{
	if(threadIdx.x < NB200_STATS_LINEAR)
	{
		double s = 0;
		for(int b = 0; b < lin_blocks; ++b) { s += lin[static_cast<size_t>(b) * NB200_STATS_LINEAR + threadIdx.x]; }
		out[threadIdx.x] = s;
	}
	if(threadIdx.x == NB200_STATS_LINEAR)
	{
		double s = 0;
		for(int b = 0; b < pot_blocks; ++b) { s += pot[b]; }
		out[NB200_STATS_LINEAR] = s;
	}
}

#endif // NB200_STATS_CUH
