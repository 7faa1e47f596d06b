// nb200 -- direct all-pairs gravity for sm_100a.
//
// Replaces the reference's kfcompute + kfcompute_xyz
// (nbody/nbody_engine_cuda_impl.cu:10-124). Same mathematics as the CPU
// engines (nbody/nbody_data.cpp:35-44, nbody/nbody_engine_block.cpp:93-112):
//
//     a_i = sum_j m_j (r_j - r_i) / max(|r_j - r_i|^2, MinDistance)^(3/2)
//
// including j == i, which contributes exactly 0 (dr = 0).
//
// Shape of the kernel
//   * sources are streamed as packed body4 {x,y,z,m} tiles; one elected thread
//     issues a TMA bulk copy (cp.async.bulk -> UBLKCP) per tile into a 3-stage
//     shared-memory ring, completion is signalled through an mbarrier;
//   * every thread owns IPT targets in registers, so one broadcast LDS.128
//     pair feeds IPT pair interactions (i-blocking);
//   * r^-3 comes from MUFU.RSQ64H + one cubically convergent correction
//     (17 FP64-pipe instructions per pair, no sqrt, no divide, no slow path);
//     the MinDistance clamp costs one integer min per pair: a tile is only redone with the
//     exact clamp when some pair of it could be closer than sqrt(MinDistance);
//   * the grid is (target blocks) x (source segments): splitting the source
//     range makes the number of equal-sized work items >> 148 SMs for every N
//     (N = 16 ... 4M, 1 ... 8 shards), partial sums are combined in a fixed
//     order by direct_reduce so results are bit-reproducible run to run.
#ifndef NB200_DIRECT_CUH
#define NB200_DIRECT_CUH

#include "nb200_common.cuh"

#define NB200_DIRECT_THREADS 128
#define NB200_DIRECT_TILE 128     // bodies per shared-memory tile
#define NB200_DIRECT_STAGES 3
#ifndef NB200_DIRECT_MINB
#define NB200_DIRECT_MINB 2
#endif

// ---- mbarrier / TMA bulk-copy primitives (PTX ISA 8.x, sm_90+) ---------------
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
	return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"WAIT_%=:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra DONE_%=;\n"
		"bra WAIT_%=;\n"
		"DONE_%=:\n"
		"}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
	asm volatile(
		"cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
		"l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
		: "memory");
}

// ---- one pair interaction --------------------------------------------------
// r2^-3/2 from a ~20-bit seed y0 (MUFU.RSQ64H) and one cubically convergent correction:
//   e = 1 - r2*y0^2;  r2^-1/2 = y0 (1-e)^-1/2 = y0 (1 + e/2 + 3e^2/8) + O(e^3) ~ 2^-60
// 17 FP64-pipe instructions per pair: 3 DADD, DMUL + 2 DFMA (r2), 5 (refinement), 3 DMUL (m y^3), 3 DFMA.
#if NB200_PRECISION == 2
#define NB200_MIN_DISTANCE_HI 0x3E45798E	// high word of 1e-8 (0x3E45798EE2308C3A)

// Exact form: r2 = max(r2, MinDistance) done on the integer pipe (r2 >= +0, so IEEE bit patterns order
// like signed integers), which keeps the FP64 pipe for arithmetic.
__device__ __forceinline__ void pair_interaction(double xi, double yi, double zi, const body4& s,
												 double& ax, double& ay, double& az)
{
	double	dx = s.x - xi;
	double	dy = s.y - yi;
	double	dz = s.z - zi;
	double	r2 = fma(dz, dz, fma(dy, dy, dx * dx));
	long long		bits = __double_as_longlong(r2);
	const long long	min_bits = 0x3E45798EE2308C3ALL;	// 1e-8
	bits = bits < min_bits ? min_bits : bits;
	r2 = __longlong_as_double(bits);
	double	y0;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(r2));
	double	h = r2 * y0;
	double	e = fma(-h, y0, 1.0);
	double	p = fma(e, 0.375, 0.5);
	double	q = y0 * e;
	double	y = fma(q, p, y0);
	double	c = (y * y) * (s.m * y);
	ax = fma(dx, c, ax);
	ay = fma(dy, c, ay);
	az = fma(dz, c, az);
}

// Fast form for the common case "no pair of this tile is closer than sqrt(MinDistance)": no clamp; instead one
// integer min per pair tracks the smallest high word of r2 seen, and the caller redoes the tile with the exact
// form if that minimum could be below MinDistance (self pair, bodies closer than 1e-4). The seed keeps the
// low word of a dead temporary instead of a zeroed one (saves a move; it perturbs y0 by < 2^-20, which the
// correction absorbs because e is computed from the y0 actually used).
__device__ __forceinline__ void pair_interaction_fast(double xi, double yi, double zi, const body4& s,
													  double& ax, double& ay, double& az, int& min_hi)
{
	double	dx = s.x - xi;
	double	dy = s.y - yi;
	double	dz = s.z - zi;
	double	t1 = dx * dx;
	double	r2 = fma(dz, dz, fma(dy, dy, t1));
	min_hi = min(min_hi, __double2hiint(r2));
	double	seed;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(seed) : "d"(r2));
	double	y0 = __hiloint2double(__double2hiint(seed), __double2loint(t1));
	double	h = r2 * y0;
	double	e = fma(-h, y0, 1.0);
	double	p = fma(e, 0.375, 0.5);
	double	q = y0 * e;
	double	y = fma(q, p, y0);
	double	c = (y * y) * (s.m * y);
	ax = fma(dx, c, ax);
	ay = fma(dy, c, ay);
	az = fma(dz, c, az);
}

// All pairs of one shared-memory tile for this thread's IPT targets, added to (ax, ay, az).
template<int IPT>
__device__ __forceinline__ void tile_interactions(const body4* __restrict__ tb, const real (&xi)[IPT], const real (&yi)[IPT],
												  const real (&zi)[IPT], real (&ax)[IPT], real (&ay)[IPT], real (&az)[IPT])
{
	real	tx[IPT], ty[IPT], tz[IPT];
	int		min_hi = 0x7fffffff;
#pragma unroll
	for(int k = 0; k < IPT; ++k)
	{
		tx[k] = ty[k] = tz[k] = 0;
	}
#pragma unroll 4
	for(int j = 0; j < NB200_DIRECT_TILE; ++j)
	{
		const body4	s = tb[j];	// warp-uniform address: broadcast LDS.128 x2
#pragma unroll
		for(int k = 0; k < IPT; ++k)
		{
			pair_interaction_fast(xi[k], yi[k], zi[k], s, tx[k], ty[k], tz[k], min_hi);
		}
	}
	if(min_hi <= NB200_MIN_DISTANCE_HI)
	{
		// rare: some r2 of this tile may be below MinDistance -> discard and redo with the exact clamp
#pragma unroll
		for(int k = 0; k < IPT; ++k)
		{
			tx[k] = ty[k] = tz[k] = 0;
		}
#pragma unroll 1
		for(int j = 0; j < NB200_DIRECT_TILE; ++j)
		{
			const body4	s = tb[j];
#pragma unroll
			for(int k = 0; k < IPT; ++k)
			{
				pair_interaction(xi[k], yi[k], zi[k], s, tx[k], ty[k], tz[k]);
			}
		}
	}
#pragma unroll
	for(int k = 0; k < IPT; ++k)
	{
		ax[k] += tx[k];
		ay[k] += ty[k];
		az[k] += tz[k];
	}
}
#else
__device__ __forceinline__ void pair_interaction(float xi, float yi, float zi, const body4& s,
												 float& ax, float& ay, float& az)
{
	float	dx = s.x - xi;
	float	dy = s.y - yi;
	float	dz = s.z - zi;
	float	r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
	r2 = fmaxf(r2, NB200_MIN_DISTANCE);
	float	y = nb200_rsqrt_normal(r2);	// MUFU.RSQ, <= 2 ulp (r2 >= 1e-8: clamped above)
	float	c = (y * y) * (s.m * y);
	ax = fmaf(dx, c, ax);
	ay = fmaf(dy, c, ay);
	az = fmaf(dz, c, az);
}

template<int IPT>
__device__ __forceinline__ void tile_interactions(const body4* __restrict__ tb, const real (&xi)[IPT], const real (&yi)[IPT],
												  const real (&zi)[IPT], real (&ax)[IPT], real (&ay)[IPT], real (&az)[IPT])
{
#pragma unroll 4
	for(int j = 0; j < NB200_DIRECT_TILE; ++j)
	{
		const body4	s = tb[j];	// warp-uniform address: broadcast LDS.128
#pragma unroll
		for(int k = 0; k < IPT; ++k)
		{
			pair_interaction(xi[k], yi[k], zi[k], s, ax[k], ay[k], az[k]);
		}
	}
}
#endif

// src_all[shard_first + i] = {x, y, z, m} for the local shard's bodies.
__global__ void __launch_bounds__(256) direct_pack(const real* __restrict__ y, const real* __restrict__ mass,
												   body4* __restrict__ src, size_t n_shard, size_t shard_first)
{
	size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	if(i >= n_shard)
	{
		return;
	}
	body4 b;
	b.x = y[i];
	b.y = y[n_shard + i];
	b.z = y[2 * n_shard + i];
	b.m = mass[shard_first + i];
	src[shard_first + i] = b;
}

// grid = (ceil(n_shard / (THREADS*IPT)), segments)
// out: segments == 1 -> f (acc rows at 3n,4n,5n; velocity rows copied to 0..3n)
//      segments  > 1 -> partial[seg][3][n_shard]
template<int IPT>
__global__ void __launch_bounds__(NB200_DIRECT_THREADS, NB200_DIRECT_MINB)
direct_pairs(const body4* __restrict__ src, const real* __restrict__ y, real* __restrict__ out,
			 size_t n_shard, size_t shard_first, int n_tiles, int tiles_per_seg, int write_f)
{
	__shared__ body4					tile[NB200_DIRECT_STAGES][NB200_DIRECT_TILE];
	__shared__ alignas(8) uint64_t		full[NB200_DIRECT_STAGES];

	const int	tid = threadIdx.x;
	const int	t_begin = blockIdx.y * tiles_per_seg;
	const int	t_end = min(n_tiles, t_begin + tiles_per_seg);
	const int	nt = t_end - t_begin;
	const uint32_t tile_bytes = NB200_DIRECT_TILE * sizeof(body4);

	if(tid == 0)
	{
		for(int s = 0; s < NB200_DIRECT_STAGES; ++s)
		{
			mbar_init(&full[s], 1);
		}
		mbar_fence_init();
	}
	__syncthreads();
	if(tid == 0)
	{
		for(int s = 0; s < NB200_DIRECT_STAGES - 1 && s < nt; ++s)
		{
			mbar_expect_tx(&full[s], tile_bytes);
			tma_bulk_g2s(&tile[s][0], src + static_cast<size_t>(t_begin + s) * NB200_DIRECT_TILE, tile_bytes, &full[s]);
		}
	}

	real	xi[IPT], yi[IPT], zi[IPT], ax[IPT], ay[IPT], az[IPT];
	const size_t	i0 = static_cast<size_t>(blockIdx.x) * (NB200_DIRECT_THREADS * IPT) + tid;
#pragma unroll
	for(int k = 0; k < IPT; ++k)
	{
		size_t	i = i0 + static_cast<size_t>(k) * NB200_DIRECT_THREADS;
		size_t	ic = i < n_shard ? i : n_shard - 1;	// out-of-range lanes compute a duplicate, never store
		body4	b = src[shard_first + ic];
		xi[k] = b.x;
		yi[k] = b.y;
		zi[k] = b.z;
		ax[k] = ay[k] = az[k] = 0;
	}

	for(int t = 0; t < nt; ++t)
	{
		const int	stage = t % NB200_DIRECT_STAGES;
		// Refill the stage consumed in iteration t-1 (all threads passed its closing barrier).
		if(tid == 0)
		{
			int tn = t + NB200_DIRECT_STAGES - 1;
			if(tn < nt)
			{
				int sn = tn % NB200_DIRECT_STAGES;
				mbar_expect_tx(&full[sn], tile_bytes);
				tma_bulk_g2s(&tile[sn][0], src + static_cast<size_t>(t_begin + tn) * NB200_DIRECT_TILE, tile_bytes, &full[sn]);
			}
		}
		mbar_wait(&full[stage], (t / NB200_DIRECT_STAGES) & 1);
		tile_interactions<IPT>(&tile[stage][0], xi, yi, zi, ax, ay, az);
		__syncthreads();
	}

#pragma unroll
	for(int k = 0; k < IPT; ++k)
	{
		size_t	i = i0 + static_cast<size_t>(k) * NB200_DIRECT_THREADS;
		if(i < n_shard)
		{
			if(write_f)
			{
				out[i] = y[3 * n_shard + i];
				out[n_shard + i] = y[4 * n_shard + i];
				out[2 * n_shard + i] = y[5 * n_shard + i];
				out[3 * n_shard + i] = ax[k];
				out[4 * n_shard + i] = ay[k];
				out[5 * n_shard + i] = az[k];
			}
			else
			{
				real*	p = out + static_cast<size_t>(blockIdx.y) * 3 * n_shard;
				p[i] = ax[k];
				p[n_shard + i] = ay[k];
				p[2 * n_shard + i] = az[k];
			}
		}
	}
}

// f[0..3n) = y[3n..6n);  f[3n + r*n + i] = sum_s partial[s][r][i]  (s ascending: fixed order)
__global__ void __launch_bounds__(256) direct_reduce(const real* __restrict__ partial, const real* __restrict__ y,
													 real* __restrict__ f, size_t n_shard, int segments)
{
	size_t	e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;	// element of the 3 x n_shard block
	if(e >= 3 * n_shard)
	{
		return;
	}
	real	acc = 0;
	for(int s = 0; s < segments; ++s)
	{
		acc += partial[static_cast<size_t>(s) * 3 * n_shard + e];
	}
	f[e] = y[3 * n_shard + e];
	f[3 * n_shard + e] = acc;
}

// ---- small systems: ONE launch per fcompute -----------------------------------------------------------------------
// Up to a few thousand bodies the three launches of the tiled path (pack, pairs, reduce) cost more than the pairs
// themselves (C1, N = 2,048: 4.2e6 pairs are ~5 us of FP64 work). Here a CTA of 256 threads owns 16 targets x 16 source
// slices, stages the whole system from the state vector and the masses into shared memory (N * 32 bytes), and thread
// (target t, slice s) sums sources s, s + 16, ...
// (two interleaved accumulator sets for ILP), the 16 slice sums of a target are then added in shared memory in slice
// order -- a fixed order, so results are bit-reproducible -- and the same kernel copies the velocity rows. Same pair
// arithmetic and exact MinDistance clamp as everywhere else. Single-shard contexts only.
#define NB200_SMALL_TARGETS 16
#define NB200_SMALL_SLICES 16
#define NB200_SMALL_MAX_BODIES 4096
#define NB200_SMALL_MAX_SMEM (200 * 1024)	// forced (direct_small = 1): up to 6400 FP64 / 12800 FP32 bodies
__global__ void __launch_bounds__(NB200_SMALL_TARGETS * NB200_SMALL_SLICES)
direct_small(const real* __restrict__ y, const real* __restrict__ mass, real* __restrict__ f, int n)
{
	extern __shared__ __align__(32) unsigned char small_smem[];
	body4*		src = reinterpret_cast<body4*>(small_smem);	// all n bodies, packed (n * sizeof(body4) bytes, <= 128 KB)
	__shared__ real part[3][NB200_SMALL_SLICES][NB200_SMALL_TARGETS];
	// every CTA stages the whole system once: coalesced, independent loads -> one memory round trip instead of one
	// per source (the state vector of a small system sits in L2)
	for(int j = threadIdx.x; j < n; j += NB200_SMALL_TARGETS * NB200_SMALL_SLICES)
	{
		body4 b;
		b.x = y[j]; b.y = y[n + j]; b.z = y[2 * n + j]; b.m = mass[j];
		src[j] = b;
	}
	__syncthreads();
	const int	t = threadIdx.x % NB200_SMALL_TARGETS;
	const int	sl = threadIdx.x / NB200_SMALL_TARGETS;
	const int	i = blockIdx.x * NB200_SMALL_TARGETS + t;
	const body4	me = src[i < n ? i : n - 1];	// idle targets shadow the last body and never store
	real		ax0 = 0, ay0 = 0, az0 = 0, ax1 = 0, ay1 = 0, az1 = 0;
	int			j = sl;
	for(; j + NB200_SMALL_SLICES < n; j += 2 * NB200_SMALL_SLICES)
	{
		// half a warp shares a slice: every shared-memory read is a two-address broadcast
		const body4 s0 = src[j], s1 = src[j + NB200_SMALL_SLICES];
		pair_interaction(me.x, me.y, me.z, s0, ax0, ay0, az0);
		pair_interaction(me.x, me.y, me.z, s1, ax1, ay1, az1);
	}
	if(j < n)
	{
		pair_interaction(me.x, me.y, me.z, src[j], ax0, ay0, az0);
	}
	part[0][sl][t] = ax0 + ax1;
	part[1][sl][t] = ay0 + ay1;
	part[2][sl][t] = az0 + az1;
	__syncthreads();
	if(sl < 3 && i < n)
	{
		// thread (component sl, target t): the 16 slice sums in slice order, and the velocity row of the same component
		real acc = 0;
#pragma unroll
		for(int q = 0; q < NB200_SMALL_SLICES; ++q) { acc += part[sl][q][t]; }
		f[(3 + sl) * static_cast<size_t>(n) + i] = acc;
		f[sl * static_cast<size_t>(n) + i] = y[(3 + sl) * static_cast<size_t>(n) + i];
	}
}

#endif // NB200_DIRECT_CUH
