// nb200 -- symmetric (Newton's third law) all-pairs tiles for N >= 32,768 (FP64 and FP32 builds).
//
// The plain kernel (nb200_direct.cuh) evaluates every ORDERED pair: 17 FP64-pipe instructions per interaction.
// Here every UNORDERED pair is evaluated once and applied to both bodies (a_i += m_j c d, a_j -= m_i c d):
// 21 FP64-pipe instructions per two interactions. Same mathematics and the same exact MinDistance clamp as
// nbody_data::force (nbody/nbody_data.cpp:35-44); only the order of the sums differs.
//
// Shape
//   * the N x N interaction matrix is cut into T x T tiles (T = 4096 or 8192 bodies); only tiles on or above the
//     diagonal are computed, one CTA (8 warps) per tile;
//   * inside a tile every warp runs a systolic exchange: a lane keeps I "row" bodies with their accumulators in
//     registers and J "column" bodies WITH THEIR accumulators; after each step (I x J pairs per lane) the column
//     group moves to the next lane by warp shuffles, so after 32 steps the 32*I row bodies have met the 32*J column
//     bodies and every column accumulator is back in its home lane -- no shared-memory traffic in the loop;
//   * column sums of a tile are combined in shared memory, one column block per warp per phase (warps are kept
//     nB/8 blocks apart and a block barrier separates phases), so the summation order is fixed;
//   * each tile writes its row sums and column sums to scratch; direct_sym_reduce adds the <= 2N/T partials of a
//     body in a fixed order. Results are therefore bit-reproducible run to run, like the plain kernel's;
//   * diagonal tiles hold both (i,j) and (j,i): they are evaluated one-sided (row sums only).
// With several ranks the tiles are dealt round-robin, every rank reduces its own tiles into a full-length partial
// vector and one NCCL reduce-scatter leaves each rank with the accelerations of its own body shard.
#ifndef NB200_DIRECT_SYM_CUH
#define NB200_DIRECT_SYM_CUH

#include "nb200_common.cuh"

#define NB200_SYM_WARPS 8
#define NB200_SYM_THREADS (32 * NB200_SYM_WARPS)
// direct_sym_shape of the kernel this build uses by default: FP64 0 = <8, 1> (8 row bodies x 1 column body per lane: one
// column body's 7 values are the only shuffles of a step; N = 1M on one B200 with the clamp-free pass: <8,1> 761 ms,
// <4,2> 800 ms, <8,2> 870 ms, <4,4> 961 ms), FP32 4 = packed f32x2 with 8 row bodies
#if NB200_PRECISION == 1
#define NB200_SYM_DEFAULT_SHAPE 4
#else
#define NB200_SYM_DEFAULT_SHAPE 0
#endif
// measured on one B200 (profiles/r1_direct_sizes.json): FP64 N = 8,192: 5.0e11 (ordered pairs) vs 6.0e11 pairs/s, FP32
// N = 8,192: 9.4e11 vs 8.5e11, N = 16,384: 1.53e12 vs 1.57e12
#if NB200_PRECISION == 1
#define NB200_SYM_MIN_BODIES 16384
#else
#define NB200_SYM_MIN_BODIES 8192
#endif

#if NB200_PRECISION == 2
__device__ __forceinline__ double sym_shfl(double v, int src_lane)
{
	int lo = __shfl_sync(0xffffffffu, __double2loint(v), src_lane);
	int hi = __shfl_sync(0xffffffffu, __double2hiint(v), src_lane);
	return __hiloint2double(hi, lo);
}
// one unordered pair: d = column - row; both accumulator sets updated: 20 FP64-pipe instructions (3 for d, 3 for r^2,
// 6 for r^-3 from the MUFU seed in one series step -- y0^3 (1 + e (3/2 + 15/8 e)), e = 1 - r^2 y0^2 exact to one rounding
// because the seed has 21 significant bits; truncation error 35/16 e^3 < 2^-56 --, 2 for the mass products, 6 DFMA).
// On this pipe an FP64 instruction holds the dispatch port for two cycles and every other instruction for one
// (profiles/microbench/bh_sum_loop.cu), so the integer clamp of r^2 (4 instructions) is worth 10 % of a pair:
// CLAMP = false leaves it out and records the smallest high word of r^2 in `low` instead (1 instruction); the tile is
// redone with the clamp if any pair of it came closer than MinDistance (sym_tile below).
template<bool CLAMP>
__device__ __forceinline__ void sym_pair(double dx, double dy, double dz, double m_row, double m_col,
										 double& ax, double& ay, double& az, double& bx, double& by, double& bz, int& low)
{
	double	r2 = fma(dz, dz, fma(dy, dy, dx * dx));
	if(CLAMP)
	{
		long long		bits = __double_as_longlong(r2);
		const long long	min_bits = 0x3E45798EE2308C3ALL;	// 1e-8, exact clamp on the integer pipe
		bits = bits < min_bits ? min_bits : bits;
		r2 = __longlong_as_double(bits);
	}
	else
	{
		low = min(low, __double2hiint(r2));	// r2 >= 0: the high word orders like the value
	}
	double	y0;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(r2));
#ifdef NB200_SYM_NEWTON
	const double	h = r2 * y0;
	const double	e = fma(-h, y0, 1.0);
	const double	pp = fma(e, 0.375, 0.5);
	const double	qq = y0 * e;
	const double	y = fma(qq, pp, y0);
	const double	y3 = (y * y) * y;
#else
	const double	y2 = y0 * y0;
	const double	e = fma(-r2, y2, 1.0);
	const double	u = e * fma(e, 1.875, 1.5);
	const double	s3 = y0 * y2;
	const double	y3 = fma(s3, u, s3);
#endif
	const double	ca = m_col * y3;
	const double	cb = m_row * y3;
	ax = fma(dx, ca, ax);
	ay = fma(dy, ca, ay);
	az = fma(dz, ca, az);
	bx = fma(-dx, cb, bx);
	by = fma(-dy, cb, by);
	bz = fma(-dz, cb, bz);
}
#define NB200_SYM_CLOSE_HI 0x3E45798E	// high word of 1e-8: a pair whose r^2 has a high word <= this may need the clamp
#else
__device__ __forceinline__ float sym_shfl(float v, int src_lane)
{
	return __shfl_sync(0xffffffffu, v, src_lane);
}
// FP32: 16 FP32-pipe instructions + MUFU.RSQ + FMNMX per unordered pair (the clamp is one instruction: always on)
template<bool CLAMP>
__device__ __forceinline__ void sym_pair(float dx, float dy, float dz, float m_row, float m_col,
										 float& ax, float& ay, float& az, float& bx, float& by, float& bz, int& low)
{
	(void)low;
	float	r2 = fmaxf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)), NB200_MIN_DISTANCE);
	float	y = nb200_rsqrt_normal(r2);
	float	y3 = (y * y) * y;
	float	ca = m_col * y3;
	float	cb = m_row * y3;
	ax = fmaf(dx, ca, ax);
	ay = fmaf(dy, ca, ay);
	az = fmaf(dz, ca, az);
	bx = fmaf(-dx, cb, bx);
	by = fmaf(-dy, cb, by);
	bz = fmaf(-dz, cb, bz);
}
#endif

// tile_rc[t] = {row block, column block} of the t-th tile this CTA grid works on
#ifndef NB200_SYM_MINB
#define NB200_SYM_MINB 1
#endif
// LATE_SHUFFLES (default): the column group moves on after all I x J pairs of a step. The alternative -- each column body
// shuffled right after its own I pairs, so that the shuffles overlap the next body's arithmetic (192 instead of 230
// registers) -- measured 5 % SLOWER at N = 1M (901 vs 855 ms); kept selectable (direct_sym_shape 6) for A/B runs.
// Also measured and rejected: two CTAs per SM with <= 128 registers and tile edge 4096 (<2,2> 967 ms, <4,1> 903 ms,
// <4,2> with 68 bytes of spills 903 ms); 12 warps per SM at 168 registers, no spills, tiles of 3072 / 6144 (1.12e12
// pairs/s, with the early shuffles 1.20e12, against 1.285e12 here): fewer warps with more registers each win;
// tile edge 4096 with this kernel: 846 ms, but twice the partial-sum scratch.
// WARPS: warps per CTA. A tile needs T / (32 I) row blocks; with fewer than 8 of them (small tiles for small N) the
// CTA shrinks instead of leaving warps idle, and several CTAs share an SM (230 registers x 128 threads fit twice).
// One tile, with or without the clamp of r^2. Returns (CLAMP = false) the smallest high word of r^2 this thread saw.
template<int I, int J, bool LATE_SHUFFLES, int WARPS, bool CLAMP>
__device__ __forceinline__ int sym_tile(const body4* __restrict__ src, const int2* __restrict__ tile_rc, real* __restrict__ p_row,
										 real* __restrict__ p_col, int tile_edge, real* colacc)
{
	int			low = 0x7fffffff;
	const int	T = tile_edge;
	const int	lane = threadIdx.x & 31;
	const int	warp = threadIdx.x >> 5;
	const int2	rc = tile_rc[blockIdx.x];
	const bool	diagonal = rc.x == rc.y;
	const body4* __restrict__ rows = src + static_cast<size_t>(rc.x) * T;
	const body4* __restrict__ cols = src + static_cast<size_t>(rc.y) * T;
	real*		out_row = p_row + static_cast<size_t>(blockIdx.x) * 3 * T;
	real*		out_col = p_col + static_cast<size_t>(blockIdx.x) * 3 * T;
	const int	n_ablk = T / (32 * I);	// row blocks of the tile, dealt to the warps round-robin
	const int	n_bblk = T / (32 * J);	// column blocks of the tile
	const int	from = (lane + 31) & 31;	// the column group arrives from lane - 1

	for(int e = threadIdx.x; e < 3 * T; e += 32 * WARPS)
	{
		colacc[e] = 0;
	}
	__syncthreads();

	for(int a0 = 0; a0 < n_ablk; a0 += WARPS)
	{
		// every warp runs every phase (the phase barrier is block-wide); warps without a row block just keep step
		const int	ablk = a0 + warp;
		const bool	active = ablk < n_ablk;
		real xa[I], ya[I], za[I], ma[I], ax[I], ay[I], az[I];
#pragma unroll
		for(int k = 0; k < I; ++k)
		{
			const body4 b = rows[(active ? ablk : 0) * 32 * I + k * 32 + lane];
			xa[k] = b.x; ya[k] = b.y; za[k] = b.z; ma[k] = b.m;
			ax[k] = ay[k] = az[k] = 0;
		}
		// phase p: warp w owns column block (p + w * n_bblk / 8) mod n_bblk -- no two warps share a block in a phase
		int		bblk = (warp * (n_bblk / WARPS)) % n_bblk;
		body4	nxt[J];
#pragma unroll
		for(int q = 0; q < J; ++q)
		{
			nxt[q] = cols[bblk * 32 * J + q * 32 + lane];
		}
		for(int p = 0; p < n_bblk; ++p)
		{
			if(active)
			{
			real xb[J], yb[J], zb[J], mb[J], bx[J], by[J], bz[J];
#pragma unroll
			for(int q = 0; q < J; ++q)
			{
				xb[q] = nxt[q].x; yb[q] = nxt[q].y; zb[q] = nxt[q].z; mb[q] = nxt[q].m;
				bx[q] = by[q] = bz[q] = 0;
			}
			const int	cur = bblk;
			bblk = bblk + 1 == n_bblk ? 0 : bblk + 1;
			if(p + 1 < n_bblk)
			{
				// prefetch the next phase's column bodies while this block rotates
#pragma unroll
				for(int q = 0; q < J; ++q)
				{
					nxt[q] = cols[bblk * 32 * J + q * 32 + lane];
				}
			}
#ifndef NB200_SYM_STEP_UNROLL
#define NB200_SYM_STEP_UNROLL 1
#endif
			constexpr int step_unroll = NB200_SYM_STEP_UNROLL;
#pragma unroll step_unroll
			for(int step = 0; step < 32; ++step)
			{
				// column body q moves on as soon as its I pairs are done, while the pairs of column body q + 1 (and, across
				// the loop edge, those of the next step's body 0) keep the FP64 pipe busy: the shuffles never drain it
#pragma unroll
				for(int q = 0; q < J; ++q)
				{
#pragma unroll
					for(int k = 0; k < I; ++k)
					{
						sym_pair<CLAMP>(xb[q] - xa[k], yb[q] - ya[k], zb[q] - za[k], ma[k], mb[q],
										ax[k], ay[k], az[k], bx[q], by[q], bz[q], low);
					}
					if(!LATE_SHUFFLES)
					{
						xb[q] = sym_shfl(xb[q], from); yb[q] = sym_shfl(yb[q], from); zb[q] = sym_shfl(zb[q], from);
						mb[q] = sym_shfl(mb[q], from);
						bx[q] = sym_shfl(bx[q], from); by[q] = sym_shfl(by[q], from); bz[q] = sym_shfl(bz[q], from);
					}
				}
				if(LATE_SHUFFLES)
				{
					// the first version of this kernel (kept for A/B measurements, direct_sym_shape 6): all shuffles after all pairs
#pragma unroll
					for(int q = 0; q < J; ++q)
					{
						xb[q] = sym_shfl(xb[q], from); yb[q] = sym_shfl(yb[q], from); zb[q] = sym_shfl(zb[q], from);
						mb[q] = sym_shfl(mb[q], from);
						bx[q] = sym_shfl(bx[q], from); by[q] = sym_shfl(by[q], from); bz[q] = sym_shfl(bz[q], from);
					}
				}
			}
			// 32 moves later every column body is home again, carrying the sum over this warp's 32*I row bodies
			if(!diagonal)
			{
#pragma unroll
				for(int q = 0; q < J; ++q)
				{
					const int c = cur * 32 * J + q * 32 + lane;
					colacc[c] += bx[q];
					colacc[T + c] += by[q];
					colacc[2 * T + c] += bz[q];
				}
			}
			}
			__syncthreads();	// phase boundary: next phase another warp owns this column block
		}
		if(active)
		{
#pragma unroll
			for(int k = 0; k < I; ++k)
			{
				const int r = ablk * 32 * I + k * 32 + lane;
				out_row[r] = ax[k];
				out_row[T + r] = ay[k];
				out_row[2 * T + r] = az[k];
			}
		}
	}
	for(int e = threadIdx.x; e < 3 * T; e += 32 * WARPS)
	{
		out_col[e] = colacc[e];
	}
	return low;
}

#ifndef NB200_SYM_FAST
#define NB200_SYM_FAST 1	// 0: always clamp (A/B)
#endif
template<int I, int J, bool LATE_SHUFFLES = true, int WARPS = NB200_SYM_WARPS>
__global__ void __launch_bounds__(32 * WARPS, NB200_SYM_MINB)
direct_sym_tiles(const body4* __restrict__ src, const int2* __restrict__ tile_rc, real* __restrict__ p_row,
				 real* __restrict__ p_col, int tile_edge)
{
	extern __shared__ real colacc[];	// [3][tile_edge]
#if NB200_PRECISION == 2
	// A diagonal tile holds every body's pair with itself (r^2 = 0): clamp from the start. Any other tile runs without
	// the clamp and is redone with it if some pair came closer than MinDistance (1e-4): exact either way, and rare --
	// nothing of the first pass is kept (row and column sums are overwritten, the column sums in shared memory reset).
	const int2	rc = tile_rc[blockIdx.x];
	if(NB200_SYM_FAST && rc.x != rc.y)
	{
		const int low = sym_tile<I, J, LATE_SHUFFLES, WARPS, false>(src, tile_rc, p_row, p_col, tile_edge, colacc);
		if(__syncthreads_or(low <= NB200_SYM_CLOSE_HI) == 0) { return; }
	}
#endif
	sym_tile<I, J, LATE_SHUFFLES, WARPS, true>(src, tile_rc, p_row, p_col, tile_edge, colacc);
}

#if NB200_PRECISION == 1
// ---- FP32 build: packed two-wide arithmetic (Blackwell fma.rn.f32x2 -> FFMA2 / FMUL2 / FADD2) -------------------------
// The FP32 path is issue-bound, so the two column bodies a lane holds are kept as one packed f32x2 per component: one
// instruction serves the pairs (row k, column 0) and (row k, column 1). 16 packed + 2 FMNMX + 2 MUFU.RSQ per two
// unordered pairs (5 issue slots per interaction instead of 9). Row accumulators are packed partial sums (from column 0
// and column 1), added when the row block is written.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_pack(float lo, float hi)
{
	f32x2 r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
	return r;
}
__device__ __forceinline__ void f2_unpack(f32x2 v, float& lo, float& hi)
{
	asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c)
{
	f32x2 d;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
	return d;
}
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b)
{
	f32x2 d;
	asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
	return d;
}
__device__ __forceinline__ f32x2 f2_sub(f32x2 a, f32x2 b)
{
	f32x2 d;
	asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
	return d;
}
__device__ __forceinline__ f32x2 f2_shfl(f32x2 v, int src_lane)
{
	return __shfl_sync(0xffffffffu, v, src_lane);
}

// Same tile algorithm as direct_sym_tiles<I, 2>, column pair packed.
template<int I, bool LATE_SHUFFLES = false, int WARPS = NB200_SYM_WARPS>
__global__ void __launch_bounds__(32 * WARPS, NB200_SYM_MINB)
direct_sym_tiles_f32x2(const body4* __restrict__ src, const int2* __restrict__ tile_rc, float* __restrict__ p_row,
					   float* __restrict__ p_col, int tile_edge)
{
	extern __shared__ float colacc[];	// [3][tile_edge]
	const int	T = tile_edge;
	const int	lane = threadIdx.x & 31;
	const int	warp = threadIdx.x >> 5;
	const int2	rc = tile_rc[blockIdx.x];
	const bool	diagonal = rc.x == rc.y;
	const body4* __restrict__ rows = src + static_cast<size_t>(rc.x) * T;
	const body4* __restrict__ cols = src + static_cast<size_t>(rc.y) * T;
	float*		out_row = p_row + static_cast<size_t>(blockIdx.x) * 3 * T;
	float*		out_col = p_col + static_cast<size_t>(blockIdx.x) * 3 * T;
	const int	n_ablk = T / (32 * I);
	const int	n_bblk = T / 64;
	const int	from = (lane + 31) & 31;

	for(int e = threadIdx.x; e < 3 * T; e += 32 * WARPS)
	{
		colacc[e] = 0;
	}
	__syncthreads();

	for(int a0 = 0; a0 < n_ablk; a0 += WARPS)
	{
		const int	ablk = a0 + warp;
		const bool	active = ablk < n_ablk;
		f32x2 xa[I], ya[I], za[I], nma[I], ax[I], ay[I], az[I];
#pragma unroll
		for(int k = 0; k < I; ++k)
		{
			const body4 b = rows[(active ? ablk : 0) * 32 * I + k * 32 + lane];
			xa[k] = f2_pack(b.x, b.x); ya[k] = f2_pack(b.y, b.y); za[k] = f2_pack(b.z, b.z);
			nma[k] = f2_pack(-b.m, -b.m);
			ax[k] = ay[k] = az[k] = f2_pack(0.f, 0.f);
		}
		int		bblk = (warp * (n_bblk / WARPS)) % n_bblk;
		body4	nxt0 = cols[bblk * 64 + lane], nxt1 = cols[bblk * 64 + 32 + lane];
		for(int p = 0; p < n_bblk; ++p)
		{
			if(active)
			{
			f32x2 xb = f2_pack(nxt0.x, nxt1.x), yb = f2_pack(nxt0.y, nxt1.y), zb = f2_pack(nxt0.z, nxt1.z);
			f32x2 mb = f2_pack(nxt0.m, nxt1.m);
			f32x2 bx = f2_pack(0.f, 0.f), by = bx, bz = bx;
			const int	cur = bblk;
			bblk = bblk + 1 == n_bblk ? 0 : bblk + 1;
			if(p + 1 < n_bblk)
			{
				nxt0 = cols[bblk * 64 + lane];
				nxt1 = cols[bblk * 64 + 32 + lane];
			}
#pragma unroll 1
			for(int step = 0; step < 32; ++step)
			{
				// the column pair's position and mass do not change during a step: send them on BEFORE the pairs, so the
				// shuffles overlap the arithmetic; only the three accumulators travel afterwards
				f32x2 nxb = xb, nyb = yb, nzb = zb, nmb = mb;
				if(!LATE_SHUFFLES)
				{
					nxb = f2_shfl(xb, from); nyb = f2_shfl(yb, from); nzb = f2_shfl(zb, from); nmb = f2_shfl(mb, from);
				}
#pragma unroll
				for(int k = 0; k < I; ++k)
				{
					f32x2	dx = f2_sub(xb, xa[k]), dy = f2_sub(yb, ya[k]), dz = f2_sub(zb, za[k]);
					f32x2	r2 = f2_fma(dz, dz, f2_fma(dy, dy, f2_mul(dx, dx)));
					float	r2l, r2h;
					f2_unpack(r2, r2l, r2h);
					f32x2	y = f2_pack(nb200_rsqrt_normal(fmaxf(r2l, NB200_MIN_DISTANCE)), nb200_rsqrt_normal(fmaxf(r2h, NB200_MIN_DISTANCE)));
					f32x2	y3 = f2_mul(f2_mul(y, y), y);
					f32x2	ca = f2_mul(mb, y3);
					f32x2	ncb = f2_mul(nma[k], y3);
					ax[k] = f2_fma(dx, ca, ax[k]);
					ay[k] = f2_fma(dy, ca, ay[k]);
					az[k] = f2_fma(dz, ca, az[k]);
					bx = f2_fma(dx, ncb, bx);
					by = f2_fma(dy, ncb, by);
					bz = f2_fma(dz, ncb, bz);
				}
				if(LATE_SHUFFLES)
				{
					nxb = f2_shfl(xb, from); nyb = f2_shfl(yb, from); nzb = f2_shfl(zb, from); nmb = f2_shfl(mb, from);
				}
				xb = nxb; yb = nyb; zb = nzb; mb = nmb;
				bx = f2_shfl(bx, from); by = f2_shfl(by, from); bz = f2_shfl(bz, from);
			}
			if(!diagonal)
			{
				float l0, l1;
				const int c = cur * 64 + lane;
				f2_unpack(bx, l0, l1); colacc[c] += l0; colacc[c + 32] += l1;
				f2_unpack(by, l0, l1); colacc[T + c] += l0; colacc[T + c + 32] += l1;
				f2_unpack(bz, l0, l1); colacc[2 * T + c] += l0; colacc[2 * T + c + 32] += l1;
			}
			}
			__syncthreads();
		}
		if(active)
		{
#pragma unroll
			for(int k = 0; k < I; ++k)
			{
				const int r = ablk * 32 * I + k * 32 + lane;
				float l0, l1;
				f2_unpack(ax[k], l0, l1); out_row[r] = l0 + l1;
				f2_unpack(ay[k], l0, l1); out_row[T + r] = l0 + l1;
				f2_unpack(az[k], l0, l1); out_row[2 * T + r] = l0 + l1;
			}
		}
	}
	for(int e = threadIdx.x; e < 3 * T; e += 32 * WARPS)
	{
		out_col[e] = colacc[e];
	}
}
#endif // NB200_PRECISION == 1

// Index of tile (r, c), c >= r, in the row-major enumeration of the upper triangle of an S x S grid.
__host__ __device__ __forceinline__ long long sym_tile_id(int r, int c, int S)
{
	return static_cast<long long>(r) * S - static_cast<long long>(r) * (r - 1) / 2 + (c - r);
}

// acc[comp][body] = sum of the partials of every tile of THIS rank that touches the body's block, in a fixed order:
// column sums of tiles (a, blk), a < blk, ascending a; then row sums of tiles (blk, c), c >= blk, ascending c.
// out layout: [shard][comp][n_shard] over all N bodies (the send buffer of the reduce-scatter; [3][N] for one shard).
__global__ void __launch_bounds__(256) direct_sym_reduce(const real* __restrict__ p_row, const real* __restrict__ p_col,
														 real* __restrict__ out, size_t n_bodies, size_t n_shard,
														 int tile_edge, int S, int rank, int nranks)
{
	const size_t body = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	if(body >= n_bodies)
	{
		return;
	}
	real			acc[3] = {0, 0, 0};
	{
		const int		T = tile_edge;
		const int		blk = static_cast<int>(body / T);
		const int		off = static_cast<int>(body % T);
		for(int a = 0; a < blk; ++a)
		{
			const long long id = sym_tile_id(a, blk, S);
			if(id % nranks == rank)
			{
				const real* p = p_col + static_cast<size_t>(id / nranks) * 3 * T + off;
				acc[0] += p[0]; acc[1] += p[T]; acc[2] += p[2 * T];
			}
		}
		for(int c = blk; c < S; ++c)
		{
			const long long id = sym_tile_id(blk, c, S);
			if(id % nranks == rank)
			{
				const real* p = p_row + static_cast<size_t>(id / nranks) * 3 * T + off;
				acc[0] += p[0]; acc[1] += p[T]; acc[2] += p[2 * T];
			}
		}
	}
	const size_t shard = body / n_shard, local = body % n_shard;
	real* o = out + shard * 3 * n_shard + local;
	o[0] = acc[0];
	o[n_shard] = acc[1];
	o[2 * n_shard] = acc[2];
}

// Lanes of one process (NVLink peers or the same device): shard `shard` sums its block of every lane's partial vector
// straight out of the peers' memory (plain loads on peer-mapped pointers), lanes in ascending order -> fixed order.
#define NB200_SYM_MAX_PEERS 16
struct sym_peers
{
	const real*	partial[NB200_SYM_MAX_PEERS];	// each [shards][3][n_shard]
	int			count;
};
__global__ void __launch_bounds__(256) direct_sym_peer_sum(const sym_peers peers, real* __restrict__ mine, size_t n_shard, int shard)
{
	const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	if(e >= 3 * n_shard)
	{
		return;
	}
	const size_t off = static_cast<size_t>(shard) * 3 * n_shard + e;
	real acc = 0;
	for(int g = 0; g < peers.count; ++g)
	{
		acc += peers.partial[g][off];
	}
	mine[e] = acc;
}

// f = (v, a) for the local shard from an acceleration block laid out [3][n_shard]
__global__ void __launch_bounds__(256) direct_sym_finish(const real* __restrict__ acc, const real* __restrict__ y,
														 real* __restrict__ f, size_t n_shard)
{
	const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	if(e >= 3 * n_shard)
	{
		return;
	}
	f[e] = y[3 * n_shard + e];
	f[3 * n_shard + e] = acc[e];
}

#endif // NB200_DIRECT_SYM_CUH
