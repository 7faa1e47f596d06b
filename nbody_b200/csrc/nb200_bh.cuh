// nb200 -- Barnes-Hut on the device: kd-heap build, bottom-up node update, tree walk.
//
// Node layout is the reference's implicit binary heap (nbody/nbody_space_heap.cpp,
// nbody_space_heap_func_priv.h): root = 1, children 2i / 2i+1, leaves [N, 2N), N = 2^k;
// per node {mass centre xyz, radius_sqr} (one 32/16-byte vector, as cuda_bh_tex packs it,
// nbody_engine_cuda_bh_tex.cpp:95-125), node mass, and for leaves the body index.
//
// What the reference does on the CPU every fcompute (D2H, std::nth_element recursion with OpenMP
// tasks, repack, H2D; nbody_engine_cuda_bh_tex.cpp:79-126) happens here on the GPU:
//
//   build   The set split at every node is value-determined (left = the count/2 smallest along
//           dim = depth % 3), so the leaf order can be produced without recursion:
//           phase A  three global presorts (x, y, z) + one stable binary partition per level and
//                    per ordering, done with an analytic segment base (segments are aligned
//                    powers of two and every segment sends exactly half of its bodies left),
//                    while segments are larger than NB200_BH_LOCAL bodies;
//           phase B  one CTA per NB200_BH_LOCAL-body segment finishes the remaining levels in
//                    shared memory with per-level bitonic sorts.
//           Ties (equal coordinates across a median) are broken by body index; the reference's
//           choice there is whatever std::nth_element happens to do.
//   update  leaves, then one launch per level bottom-up (kupdate_node_bh_tex,
//           nbody_engine_cuda_impl.cu:644-714), top 256 nodes in one CTA.
//   walk    default: the grouped walk of nb200_bh_group.cuh (a warp per 32 consecutive leaves; lanes work on nodes while
//           deciding, on targets while summing). This file holds the walks that keep one traversal state per target
//           (walk_mode 32 / 2 / 4 / 1) -- the form whose summation order is the reference's:
//           warp-coherent stackless walk: the 32 targets of a warp are consecutive
//           leaves (a compact cell of the kd-order), the warp walks the UNION of their
//           traversals with a warp-uniform `curr`, so each node is one broadcast load instead
//           of 32 divergent ones. A lane that accepts a node sleeps until curr reaches that
//           node's skip_idx; every lane therefore accepts exactly the nodes, in exactly the
//           order, of its own nbody_space_heap_stackless::traverse
//           (nbody_space_heap_stackless.cpp:3-28) -- results are bit-identical to the
//           per-thread walk (walk_mode 1, the reference's kfcompute_heap_bh_stackless shape). bh_walk_warp_multi<2>
//           lets each lane carry TWO targets (the warp walks the union of 64 consecutive leaves), which
//           shares the node load, index algebra, votes and loop control of a visit between two acceptance tests.
//
//   shards  with G shards (lanes or ranks) every shard builds the whole tree and walks chunks of 4096 CONSECUTIVE
//           leaves dealt round-robin (a chunk is a compact region, so warps stay coherent; dealing evens out dense
//           and sparse regions); one all-gather (3N reals) later each shard picks its own bodies through the
//           inverse leaf map. The
//           reference instead zero-fills f on every device and all-reduces all 6N values (synchronize_sum).
//
// The only library primitive is cub::DeviceRadixSort (3 presorts per rebuild); everything else is hand-written.
#ifndef NB200_BH_CUH
#define NB200_BH_CUH

#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include "nb200_common.cuh"
#include "nb200_direct.cuh"

#define NB200_BH_LOCAL 1024        // bodies per phase-B segment (one CTA)
#define NB200_BH_PART_BLOCK 1024   // elements per partition block (256 threads x 4)
#define NB200_BH_DEAL_CHUNK 4096   // consecutive leaves per chunk when the walk is dealt to several shards

#if NB200_PRECISION == 2
typedef double4 node4;	// 32 bytes
#else
typedef float4 node4;	// 16 bytes
#endif

struct bh_state
{
	size_t	n = 0;
	node4*	xyzr = nullptr;		// [2n]
	real*	nmass = nullptr;	// [2n]
	real*	bmin = nullptr;		// [2n][3]
	real*	bmax = nullptr;		// [2n][3]
	int*	body_n = nullptr;	// [2n]
	// build scratch
	real*	keys_in = nullptr;
	real*	keys_out = nullptr;
	int*	iota = nullptr;
	int*	ord[3] = {nullptr, nullptr, nullptr};
	int*	ord_tmp[3] = {nullptr, nullptr, nullptr};
	unsigned char*	side = nullptr;
	unsigned char*	side_next = nullptr;	// written by a level's scatter pass for the level below
	unsigned*	blk = nullptr;	// [3][n / PART_BLOCK + 1]
	void*	cub_tmp = nullptr;
	size_t	cub_bytes = 0;
	// sharded walk: shard g walks the contiguous leaves [g*n_shard, (g+1)*n_shard) (a spatially compact set), the
	// accelerations are gathered in leaf order and each shard picks its own bodies through leaf_pos
	int*	leaf_pos = nullptr;	// [n] leaf position of every body (inverse of body_n[n..2n))
	real*	acc_all = nullptr;	// [shards][3][n_shard] accelerations in leaf order
	bool	have_tree = false;
	// longest-walk-first launch order of the walk's CTAs (see bh_walk_warp_multi): cost of every CTA in the last walk,
	// and the permutation the next walk is launched in
	unsigned*	lpt_cost = nullptr;		// [ctas] clock cycles of the CTA's first warp
	unsigned*	lpt_cost_sorted = nullptr;
	int*		lpt_iota = nullptr;		// 0, 1, 2, ...
	int*		lpt_order = nullptr;	// CTA indices, most expensive first
	void*		lpt_tmp = nullptr;
	size_t		lpt_tmp_bytes = 0;
	int			lpt_ctas = 0;			// CTAs the arrays are sized for
	bool		lpt_have_order = false;
};

static void bh_free(bh_state* s)
{
	if(s == nullptr) { return; }
	void* ptrs[] = {s->xyzr, s->nmass, s->bmin, s->bmax, s->body_n, s->keys_in, s->keys_out, s->iota, s->ord[0], s->ord[1],
					s->ord[2], s->ord_tmp[0], s->ord_tmp[1], s->ord_tmp[2], s->side, s->side_next, s->blk, s->cub_tmp, s->leaf_pos,
					s->acc_all, s->lpt_cost, s->lpt_cost_sorted, s->lpt_iota, s->lpt_order, s->lpt_tmp};
	for(void* p : ptrs)
	{
		if(p != nullptr) { cudaFree(p); }
	}
	delete s;
}

// ---- heap index algebra (nbody_space_heap_func_priv.h:34-72, CUDA branch :37-38) ----------------
__device__ __forceinline__ int heap_skip_idx(int idx)
{
	return (idx >> (__ffs(~idx) - 1)) + 1;
}
__device__ __forceinline__ int heap_next_up(int idx, int tree_size)
{
	int left = idx << 1;
	return left < tree_size ? left : heap_skip_idx(idx);
}

// ---- build, phase A -----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bh_extract_keys(const body4* __restrict__ src, real* __restrict__ keys,
													   int* __restrict__ iota, int n, int dim)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) { return; }
	const body4 b = src[i];
	keys[i] = dim == 0 ? b.x : (dim == 1 ? b.y : b.z);
	iota[i] = i;
}

// side[body] = 1 if the body's rank inside its segment (along the split dimension) is in the upper half
__global__ void __launch_bounds__(256) bh_mark_side(const int* __restrict__ ord_c, unsigned char* __restrict__ side, int n, int seg)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) { return; }
	side[ord_c[i]] = ((i & (seg - 1)) >= (seg >> 1)) ? 1 : 0;
}

// blk[a][b] = number of left-going bodies among positions [b*PART_BLOCK, (b+1)*PART_BLOCK) of ordering a
__global__ void __launch_bounds__(256) bh_part_count(const int* __restrict__ o0, const int* __restrict__ o1,
													  const unsigned char* __restrict__ side, unsigned* __restrict__ blk, int nblk)
{
	const int* ord = blockIdx.y == 0 ? o0 : o1;
	int base = blockIdx.x * NB200_BH_PART_BLOCK + threadIdx.x * 4;
	int4 b = *reinterpret_cast<const int4*>(ord + base);
	unsigned lefts = (side[b.x] == 0) + (side[b.y] == 0) + (side[b.z] == 0) + (side[b.w] == 0);
	typedef cub::BlockReduce<unsigned, 256> reduce_t;
	__shared__ typename reduce_t::TempStorage tmp;
	unsigned total = reduce_t(tmp).Sum(lefts);
	if(threadIdx.x == 0) { blk[blockIdx.y * (nblk + 1) + blockIdx.x] = total; }
}

// exclusive scan of each row of blk (nblk <= 4096 entries), one CTA per row
__global__ void __launch_bounds__(1024) bh_part_scan(unsigned* __restrict__ blk, int nblk)
{
	unsigned* row = blk + blockIdx.x * (nblk + 1);
	typedef cub::BlockScan<unsigned, 1024> scan_t;
	__shared__ typename scan_t::TempStorage tmp;
	__shared__ unsigned carry;
	if(threadIdx.x == 0) { carry = 0; }
	__syncthreads();
	for(int b0 = 0; b0 < nblk; b0 += 1024)
	{
		int		i = b0 + threadIdx.x;
		unsigned v = i < nblk ? row[i] : 0, ex, total;
		scan_t(tmp).ExclusiveSum(v, ex, total);
		if(i < nblk) { row[i] = ex + carry; }
		__syncthreads();
		if(threadIdx.x == 0) { carry += total; }
		__syncthreads();
	}
}

// stable partition of every seg-sized segment of two orderings: left-goers keep their order in the lower half
__global__ void __launch_bounds__(256) bh_part_scatter(const int* __restrict__ o0, const int* __restrict__ o1,
														int* __restrict__ d0, int* __restrict__ d1,
														const unsigned char* __restrict__ side, const unsigned* __restrict__ blk,
														int nblk, int seg, unsigned char* __restrict__ side_next)
{
	const int*	ord = blockIdx.y == 0 ? o0 : o1;
	int*		dst = blockIdx.y == 0 ? d0 : d1;
	int			base = blockIdx.x * NB200_BH_PART_BLOCK + threadIdx.x * 4;
	int4		b = *reinterpret_cast<const int4*>(ord + base);
	int			body[4] = {b.x, b.y, b.z, b.w};
	unsigned	left[4], lefts = 0;
#pragma unroll
	for(int q = 0; q < 4; ++q)
	{
		left[q] = side[body[q]] == 0;
		lefts += left[q];
	}
	typedef cub::BlockScan<unsigned, 256> scan_t;
	__shared__ typename scan_t::TempStorage tmp;
	unsigned before;
	scan_t(tmp).ExclusiveSum(lefts, before);
	before += blk[blockIdx.y * (nblk + 1) + blockIdx.x];	// left-goers before this thread's first element, globally
#pragma unroll
	for(int q = 0; q < 4; ++q)
	{
		int			i = base + q;
		int			s0 = i & ~(seg - 1);
		unsigned	in_seg_left = before - static_cast<unsigned>(s0 >> 1);	// every earlier segment sent exactly half left
		int			d = left[q] ? s0 + static_cast<int>(in_seg_left)
								: s0 + (seg >> 1) + (i - s0) - static_cast<int>(in_seg_left);
		dst[d] = body[q];
		before += left[q];
		// ordering 0 of this pass is the one the NEXT level splits along: its bodies' new positions say on which side of
		// the next median they fall, so the next level needs no bh_mark_side pass of its own
		if(side_next != nullptr && blockIdx.y == 0)
		{
			const int half = seg >> 1;	// the next level's segment size
			side_next[body[q]] = ((d & (half - 1)) >= (half >> 1)) ? 1 : 0;
		}
	}
}

// ---- build, phase B: finish one segment of `local` bodies in shared memory ---------------------------
template<int LOCAL_MAX>
__global__ void __launch_bounds__(LOCAL_MAX / 2) bh_local_build(const body4* __restrict__ src, const int* __restrict__ ord,
																 int* __restrict__ body_n, int n, int local, int first_depth)
{
	__shared__ real	coord[3][LOCAL_MAX];
	__shared__ int	id[LOCAL_MAX];
	__shared__ short perm[LOCAL_MAX];
	const int seg0 = blockIdx.x * local;
	for(int t = threadIdx.x; t < local; t += blockDim.x)
	{
		int body = ord != nullptr ? ord[seg0 + t] : seg0 + t;
		const body4 b = src[body];
		coord[0][t] = b.x;
		coord[1][t] = b.y;
		coord[2][t] = b.z;
		id[t] = body;
		perm[t] = static_cast<short>(t);
	}
	__syncthreads();
	int depth = first_depth;
	for(int seg = local; seg >= 2; seg >>= 1, ++depth)
	{
		const real* key = coord[depth % 3];
		// bitonic sort of every aligned `seg`-block, ascending by (key, body index)
		for(int k = 2; k <= seg; k <<= 1)
		{
			for(int j = k >> 1; j > 0; j >>= 1)
			{
				for(int t = threadIdx.x; t < local / 2; t += blockDim.x)
				{
					int		i = ((t & ~(j - 1)) << 1) | (t & (j - 1));	// lower index of the pair
					int		p = i | j;
					bool	up = (k == seg) || ((i & k) == 0);
					short	a = perm[i], c = perm[p];
					real	ka = key[a], kc = key[c];
					bool	a_gt_c = ka > kc || (ka == kc && id[a] > id[c]);
					if(a_gt_c == up)
					{
						perm[i] = c;
						perm[p] = a;
					}
				}
				__syncthreads();
			}
		}
	}
	for(int t = threadIdx.x; t < local; t += blockDim.x)
	{
		body_n[n + seg0 + t] = id[perm[t]];
	}
}

// ---- node update --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bh_update_leaves(const body4* __restrict__ src, const int* __restrict__ body_n,
														 node4* __restrict__ xyzr, real* __restrict__ nmass,
														 real* __restrict__ bmin, real* __restrict__ bmax, int n)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) { return; }
	int			idx = n + i;
	const body4	b = src[body_n[idx]];
	node4		v;
	v.x = b.x; v.y = b.y; v.z = b.z; v.w = 0;	// leaf: radius_sqr = 0 (value-initialised in the reference)
	xyzr[idx] = v;
	nmass[idx] = b.m;
	bmin[3 * idx + 0] = b.x; bmin[3 * idx + 1] = b.y; bmin[3 * idx + 2] = b.z;
	bmax[3 * idx + 0] = b.x; bmax[3 * idx + 1] = b.y; bmax[3 * idx + 2] = b.z;
}

// nbody_space_heap::update (nbody_space_heap.cpp:119-134)
__device__ __forceinline__ void bh_update_node(int idx, node4* xyzr, real* nmass, real* bmin, real* bmax, real ratio_sqr)
{
	const int	l = idx << 1, r = l + 1;
	const node4	cl = xyzr[l], cr = xyzr[r];
	const real	ml = nmass[l], mr = nmass[r];
	const real	m = ml + mr;
	node4		v;
	v.x = (cl.x * ml + cr.x * mr) / m;
	v.y = (cl.y * ml + cr.y * mr) / m;
	v.z = (cl.z * ml + cr.z * mr) / m;
	real lo[3], hi[3];
#pragma unroll
	for(int d = 0; d < 3; ++d)
	{
		lo[d] = min(bmin[3 * l + d], bmin[3 * r + d]);
		hi[d] = max(bmax[3 * l + d], bmax[3 * r + d]);
		bmin[3 * idx + d] = lo[d];
		bmax[3 * idx + d] = hi[d];
	}
	real ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
	real cx = (hi[0] + lo[0]) / 2 - v.x, cy = (hi[1] + lo[1]) / 2 - v.y, cz = (hi[2] + lo[2]) / 2 - v.z;
	real rad = sqrt(ex * ex + ey * ey + ez * ez) * static_cast<real>(0.5) + sqrt(cx * cx + cy * cy + cz * cz);
	v.w = (rad * rad) * ratio_sqr;
	xyzr[idx] = v;
	nmass[idx] = m;
}

__global__ void __launch_bounds__(256) bh_update_level(node4* xyzr, real* nmass, real* bmin, real* bmax, int level_size, real ratio_sqr)
{
	int t = blockIdx.x * blockDim.x + threadIdx.x;
	if(t >= level_size) { return; }
	bh_update_node(level_size + t, xyzr, nmass, bmin, bmax, ratio_sqr);
}

// levels of at most 256 nodes (the top of the tree), one CTA, block barrier between levels
__global__ void __launch_bounds__(256) bh_update_top(node4* xyzr, real* nmass, real* bmin, real* bmax, int first_level_size, real ratio_sqr)
{
	for(int level = first_level_size; level >= 1; level >>= 1)
	{
		if(static_cast<int>(threadIdx.x) < level)
		{
			bh_update_node(level + threadIdx.x, xyzr, nmass, bmin, bmax, ratio_sqr);
		}
		__threadfence_block();
		__syncthreads();
	}
}

// ---- sharded walk helpers -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bh_leaf_positions(const int* __restrict__ body_n, int* __restrict__ leaf_pos, int n)
{
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	if(k >= n) { return; }
	leaf_pos[body_n[n + k]] = k;
}

// f = (v, a) for this shard's bodies from the gathered leaf-ordered accelerations
// Leaf chunks of `chunk` consecutive leaves are dealt to the shards round-robin (dense and sparse regions of the tree
// cost very different walk lengths; interleaving evens the shards out while a chunk is still a compact region):
// leaf position p -> chunk c = p / chunk -> shard c % G, slot (c / G) * chunk + p % chunk of that shard's block.
__device__ __forceinline__ size_t bh_leaf_slot(size_t p, size_t chunk, size_t nshards, size_t n_shard)
{
	const size_t c = p / chunk;
	return (c % nshards) * 3 * n_shard + (c / nshards) * chunk + p % chunk;
}

__global__ void __launch_bounds__(256) bh_scatter_own(const real* __restrict__ acc_all, const int* __restrict__ leaf_pos,
													   const real* __restrict__ y, real* __restrict__ f, size_t n_shard, size_t shard_first,
													   size_t chunk, size_t nshards)
{
	size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	if(i >= n_shard) { return; }
	const size_t p = static_cast<size_t>(leaf_pos[shard_first + i]);
	const real*  a = acc_all + bh_leaf_slot(p, chunk, nshards, n_shard);
	f[i] = y[3 * n_shard + i];
	f[n_shard + i] = y[4 * n_shard + i];
	f[2 * n_shard + i] = y[5 * n_shard + i];
	f[3 * n_shard + i] = a[0];
	f[4 * n_shard + i] = a[n_shard];
	f[5 * n_shard + i] = a[2 * n_shard];
}

// ---- walk ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ node4 load_node(const node4* __restrict__ xyzr, int idx)
{
#if NB200_PRECISION == 2
	// one 256-bit load per node (sm_100 LDG.E.256)
	node4 v;
	asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(xyzr + idx));
	return v;
#else
	return __ldg(xyzr + idx);
#endif
}

__device__ __forceinline__ void store_f(const real* __restrict__ y, real* __restrict__ f, size_t n_shard, size_t li,
										real ax, real ay, real az)
{
	f[li] = y[3 * n_shard + li];
	f[n_shard + li] = y[4 * n_shard + li];
	f[2 * n_shard + li] = y[5 * n_shard + li];
	f[3 * n_shard + li] = ax;
	f[4 * n_shard + li] = ay;
	f[5 * n_shard + li] = az;
}

// |me - node|^2 of the acceptance test, with the rounding the reference's (v1 - cm).norm() gets from gcc -O3 (and that
// nvcc's own contraction of dx*dx + dy*dy + dz*dz produces): fma(dz, dz, fma(dx, dx, dy*dy)). Written out so that every
// walk -- scalar, packed f32x2, exact re-evaluation -- rounds alike and knife-edge decisions (d2 == radius_sqr up to an
// ulp, e.g. an equal-mass pair at ratio 1) fall the same way as in simple_bh (tests compare visit counts with the CPU walk).
__device__ __forceinline__ real bh_d2(real dx, real dy, real dz)
{
#if NB200_PRECISION == 2
	return fma(dz, dz, fma(dx, dx, __dmul_rn(dy, dy)));
#else
	return fmaf(dz, dz, fmaf(dx, dx, __fmul_rn(dy, dy)));
#endif
}

// Accepted node: same force form as the direct kernel, reusing the separation d = me - node and d2 of the
// acceptance test (clamp applied after the test, as in kfcompute_heap_bh_stackless, impl.cu:399-413).
// a += m (node - me) / r^3  ==  a -= m d / r^3
__device__ __forceinline__ void node_force_from_test(real dx, real dy, real dz, real d2, real m, real& ax, real& ay, real& az)
{
#if NB200_PRECISION == 2
	long long		bits = __double_as_longlong(d2);
	const long long	min_bits = 0x3E45798EE2308C3ALL;	// 1e-8
	bits = bits < min_bits ? min_bits : bits;
	const double	r2 = __longlong_as_double(bits);
	double	y0;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(r2));
	double	h = r2 * y0;
	double	e = fma(-h, y0, 1.0);
	double	p = fma(e, 0.375, 0.5);
	double	q = y0 * e;
	double	yv = fma(q, p, y0);
	double	c = (yv * yv) * (m * yv);
	ax = fma(-dx, c, ax);
	ay = fma(-dy, c, ay);
	az = fma(-dz, c, az);
#else
	const float	r2 = fmaxf(d2, NB200_MIN_DISTANCE);
	float	yv = nb200_rsqrt_normal(r2);
	float	c = (yv * yv) * (m * yv);
	ax = fmaf(-dx, c, ax);
	ay = fmaf(-dy, c, ay);
	az = fmaf(-dz, c, az);
#endif
}

// one thread per target, independent stackless walks (the reference kernel's shape)
__global__ void __launch_bounds__(256) bh_walk_thread(const node4* __restrict__ xyzr, const real* __restrict__ nmass,
													   const int* __restrict__ body_n, real* __restrict__ acc_leaf, int3 deal,
													   const real* __restrict__ y, real* __restrict__ f, int n, int n_targets,
													   size_t n_shard, int shard_first, unsigned long long* __restrict__ stats)
{
	int t = blockIdx.x * blockDim.x + threadIdx.x;
	if(t >= n_targets) { return; }
	// deal = {chunk, shards, shard}: target t of this shard is leaf ((t / chunk) * shards + shard) * chunk + t % chunk
	const int	leaf = n + ((t / deal.x) * deal.y + deal.z) * deal.x + t % deal.x;
	const int	tree_size = 2 * n;
	const node4	me = load_node(xyzr, leaf);
	real		ax = 0, ay = 0, az = 0;
	unsigned	visits = 0, inter = 0;
	int			curr = 1;
	do
	{
		const node4	nd = load_node(xyzr, curr);
		const real dx = me.x - nd.x, dy = me.y - nd.y, dz = me.z - nd.z;
		const real d2 = bh_d2(dx, dy, dz);
		++visits;
		if(d2 > nd.w)
		{
			node_force_from_test(dx, dy, dz, d2, nmass[curr], ax, ay, az);
			++inter;
			curr = heap_skip_idx(curr);
		}
		else
		{
			curr = heap_next_up(curr, tree_size);
		}
	} while(curr != 1);
	if(acc_leaf != nullptr)
	{
		acc_leaf[t] = ax;
		acc_leaf[n_shard + t] = ay;
		acc_leaf[2 * n_shard + t] = az;
	}
	else
	{
		store_f(y, f, n_shard, static_cast<size_t>(body_n[leaf] - shard_first), ax, ay, az);
	}
	if(stats != nullptr)
	{
		atomicAdd(stats + 2, static_cast<unsigned long long>(visits));
		atomicAdd(stats + 3, static_cast<unsigned long long>(inter));
	}
}

// one warp per 32 consecutive targets, warp-uniform walk over the union of the lanes' traversals
#ifndef NB200_BH_WALK_MINB
#define NB200_BH_WALK_MINB 8	// 8 x 256 threads = all 64 warp slots of an SM (caps the kernel at 32 registers)
#endif
template<bool STATS>
__global__ void __launch_bounds__(256, NB200_BH_WALK_MINB) bh_walk_warp(const node4* __restrict__ xyzr, const real* __restrict__ nmass,
													 const int* __restrict__ body_n, real* __restrict__ acc_leaf, int3 deal,
													 const real* __restrict__ y, real* __restrict__ f, int n, int n_targets,
													 size_t n_shard, int shard_first, unsigned long long* __restrict__ stats)
{
	const int	t = blockIdx.x * blockDim.x + threadIdx.x;
	const bool	live = t < n_targets;
	const int	tc = live ? t : n_targets - 1;	// idle lanes shadow the last target and never store
	// deal = {chunk, shards, shard}: target t of this shard is leaf ((t / chunk) * shards + shard) * chunk + t % chunk
	const int	leaf = n + ((tc / deal.x) * deal.y + deal.z) * deal.x + tc % deal.x;
	const int	tree_size = 2 * n;
	const node4	me = load_node(xyzr, leaf);
	real		ax = 0, ay = 0, az = 0;
	unsigned	visits = 0, inter = 0;
	int			curr = 1;		// warp-uniform
	int			resume = 1;		// node at which this lane wakes up again (valid while asleep)
	bool		awake = true;
	do
	{
		const node4	nd = load_node(xyzr, curr);	// same address in every lane: one broadcast transaction
		const int	skip = heap_skip_idx(curr);	// one BREV + FLO per visit, shared by the sleeping rule and the step
		const int	child = curr << 1;
		awake = awake || (curr == resume);
		const real	dx = me.x - nd.x, dy = me.y - nd.y, dz = me.z - nd.z;
		const real	d2 = bh_d2(dx, dy, dz);
		const bool	accept = awake && (d2 > nd.w);
		const bool	open = awake && !accept;
		if(STATS) { visits += awake ? 1u : 0u; }
		if(__any_sync(0xffffffffu, accept))
		{
			const real m = nmass[curr];
			if(accept)
			{
				node_force_from_test(dx, dy, dz, d2, m, ax, ay, az);
				if(STATS) { ++inter; }
				awake = false;
				resume = skip;
			}
		}
		// descend if any awake lane needs the children; leaves have none (heap_next_up falls back to skip_idx)
		curr = (__any_sync(0xffffffffu, open) && child < tree_size) ? child : skip;
	} while(curr != 1);
	if(live)
	{
		if(acc_leaf != nullptr)
		{
			acc_leaf[t] = ax;
			acc_leaf[n_shard + t] = ay;
			acc_leaf[2 * n_shard + t] = az;
		}
		else
		{
			store_f(y, f, n_shard, static_cast<size_t>(body_n[leaf] - shard_first), ax, ay, az);
		}
	}
	if(STATS && live)
	{
		atomicAdd(stats + 2, static_cast<unsigned long long>(visits));
		atomicAdd(stats + 3, static_cast<unsigned long long>(inter));
	}
}

// TPL targets per lane: one warp walks the union of 32 * TPL consecutive leaves, so the node load, the index algebra,
// the votes and the loop control of a visit are shared by TPL acceptance tests per lane instead of one. Each target
// still accepts exactly the nodes of its own stackless traversal, in the same order: results are bit-identical to
// bh_walk_warp (tested). Target k of lane l is leaf base + 32 k + l, so every load and store stays coalesced.
//
// Launch order. Walks differ in length (targets in dense regions open more nodes), and the last, partly filled wave of
// CTAs sets the kernel's end -- noticeably once a GPU has only a few waves of them (4M bodies over 8 GPUs: 1.7 waves).
// cta_cost, if given, receives the clock cycles each CTA's first warp took; cta_order, if given, is a permutation of
// the CTA indices, and the host passes last step's costs sorted in descending order: the longest walks start first and
// the short ones fill the end (longest-processing-time-first). Which CTA computes which targets does not change what is
// computed: results stay bit-identical.
template<int TPL, bool STATS, int MINB>
__global__ void __launch_bounds__(128, MINB) bh_walk_warp_multi(const node4* __restrict__ xyzr, const real* __restrict__ nmass,
														 const int* __restrict__ body_n, real* __restrict__ acc_leaf, int3 deal,
														 const real* __restrict__ y, real* __restrict__ f, int n, int n_targets,
														 size_t n_shard, int shard_first, unsigned long long* __restrict__ stats,
														 const int* __restrict__ cta_order, unsigned* __restrict__ cta_cost)
{
	// the cost of a CTA is the time its first thread spends here; the start time waits in shared memory so that the walk
	// loop, which has no register to spare at 64 per thread, carries nothing for it
	__shared__ unsigned t_begin;
	if(cta_cost != nullptr && threadIdx.x == 0) { t_begin = static_cast<unsigned>(clock64()); }
	const int	warp = ((cta_order != nullptr ? cta_order[blockIdx.x] : static_cast<int>(blockIdx.x)) * blockDim.x + threadIdx.x) >> 5;
	const int	lane = threadIdx.x & 31;
	const int	tree_size = 2 * n;
	int			t[TPL], leaf[TPL], resume[TPL];
	bool		awake[TPL];
	real		px[TPL], py[TPL], pz[TPL], ax[TPL], ay[TPL], az[TPL];
	unsigned	visits = 0, inter = 0;
#pragma unroll
	for(int k = 0; k < TPL; ++k)
	{
		t[k] = warp * (32 * TPL) + 32 * k + lane;
		const int tc = t[k] < n_targets ? t[k] : n_targets - 1;	// idle slots shadow the last target and never store
		leaf[k] = n + ((tc / deal.x) * deal.y + deal.z) * deal.x + tc % deal.x;
		const node4 me = load_node(xyzr, leaf[k]);
		px[k] = me.x; py[k] = me.y; pz[k] = me.z;
		ax[k] = ay[k] = az[k] = 0;
		resume[k] = 1;
		awake[k] = true;
	}
	int curr = 1;	// warp-uniform
	do
	{
		const node4	nd = load_node(xyzr, curr);
		const int	skip = heap_skip_idx(curr);
		const int	child = curr << 1;
		real		dx[TPL], dy[TPL], dz[TPL], d2[TPL];
		bool		accept[TPL];
		bool		any_accept = false, any_open = false;
#pragma unroll
		for(int k = 0; k < TPL; ++k)
		{
			awake[k] = awake[k] || (curr == resume[k]);
			dx[k] = px[k] - nd.x; dy[k] = py[k] - nd.y; dz[k] = pz[k] - nd.z;
			d2[k] = bh_d2(dx[k], dy[k], dz[k]);
			accept[k] = awake[k] && (d2[k] > nd.w);
			any_accept = any_accept || accept[k];
			any_open = any_open || (awake[k] && !accept[k]);
			if(STATS) { visits += (awake[k] && t[k] < n_targets) ? 1u : 0u; }
		}
		if(__any_sync(0xffffffffu, any_accept))
		{
			const real m = nmass[curr];
#pragma unroll
			for(int k = 0; k < TPL; ++k)
			{
				if(accept[k])
				{
					node_force_from_test(dx[k], dy[k], dz[k], d2[k], m, ax[k], ay[k], az[k]);
					if(STATS && t[k] < n_targets) { ++inter; }
					awake[k] = false;
					resume[k] = skip;
				}
			}
		}
		curr = (__any_sync(0xffffffffu, any_open) && child < tree_size) ? child : skip;
	} while(curr != 1);
	if(cta_cost != nullptr && threadIdx.x == 0)
	{
		cta_cost[cta_order != nullptr ? cta_order[blockIdx.x] : static_cast<int>(blockIdx.x)] = static_cast<unsigned>(clock64()) - t_begin;
	}
#pragma unroll
	for(int k = 0; k < TPL; ++k)
	{
		if(t[k] < n_targets)
		{
			if(acc_leaf != nullptr)
			{
				acc_leaf[t[k]] = ax[k];
				acc_leaf[n_shard + t[k]] = ay[k];
				acc_leaf[2 * n_shard + t[k]] = az[k];
			}
			else
			{
				store_f(y, f, n_shard, static_cast<size_t>(body_n[leaf[k]] - shard_first), ax[k], ay[k], az[k]);
			}
		}
	}
	if(STATS && (visits | inter) != 0)
	{
		atomicAdd(stats + 2, static_cast<unsigned long long>(visits));
		atomicAdd(stats + 3, static_cast<unsigned long long>(inter));
	}
}

#include "nb200_bh_group.cuh"

// ---- host orchestration ----------------------------------------------------------------------------------------------
#define BH_CU(call)                                                                                    \
	do                                                                                                 \
	{                                                                                                  \
		cudaError_t bh_res_ = (call);                                                                  \
		if(bh_res_ != cudaSuccess)                                                                     \
		{                                                                                              \
			err = std::string(#call) + ": " + cudaGetErrorString(bh_res_);                            \
			return NB200_ERR_CUDA;                                                                     \
		}                                                                                              \
	} while(0)

static int bh_alloc(nb200_ctx* ctx, nb200_lane& l, std::string& err)
{
	const size_t n = ctx->n;
	if(l.bh != nullptr && l.bh->n == n) { return NB200_OK; }
	bh_free(l.bh);
	l.bh = nullptr;
	bh_state* s = new bh_state();
	s->n = n;
	const size_t nblk = n / NB200_BH_PART_BLOCK + 1;
	bool ok = cudaMalloc(&s->xyzr, 2 * n * sizeof(node4)) == cudaSuccess &&
			  cudaMalloc(&s->nmass, 2 * n * sizeof(real)) == cudaSuccess &&
			  cudaMalloc(&s->bmin, 6 * n * sizeof(real)) == cudaSuccess &&
			  cudaMalloc(&s->bmax, 6 * n * sizeof(real)) == cudaSuccess &&
			  cudaMalloc(&s->body_n, 2 * n * sizeof(int)) == cudaSuccess;
	if(ok && n > NB200_BH_LOCAL)
	{
		ok = cudaMalloc(&s->keys_in, n * sizeof(real)) == cudaSuccess && cudaMalloc(&s->keys_out, n * sizeof(real)) == cudaSuccess &&
			 cudaMalloc(&s->iota, n * sizeof(int)) == cudaSuccess && cudaMalloc(&s->side, n) == cudaSuccess &&
			 cudaMalloc(&s->side_next, n) == cudaSuccess &&
			 cudaMalloc(&s->blk, 3 * nblk * sizeof(unsigned)) == cudaSuccess;
		for(int a = 0; a < 3 && ok; ++a)
		{
			ok = cudaMalloc(&s->ord[a], n * sizeof(int)) == cudaSuccess && cudaMalloc(&s->ord_tmp[a], n * sizeof(int)) == cudaSuccess;
		}
		if(ok)
		{
			size_t bytes = 0;
			cub::DeviceRadixSort::SortPairs(nullptr, bytes, s->keys_in, s->keys_out, s->iota, s->ord[0], static_cast<int>(n));
			s->cub_bytes = bytes;
		}
	}
	if(ok && ctx->nshards > 1)
	{
		ok = cudaMalloc(&s->leaf_pos, n * sizeof(int)) == cudaSuccess && cudaMalloc(&s->acc_all, 3 * n * sizeof(real)) == cudaSuccess;
	}
	if(ok && s->cub_bytes > 0) { ok = cudaMalloc(&s->cub_tmp, s->cub_bytes) == cudaSuccess; }
	if(!ok)
	{
		cudaGetLastError();
		bh_free(s);
		err = "tree allocation failed";
		return NB200_ERR_ALLOC;
	}
	// slot 0 is unused by the heap; keep it defined for bh_export
	cudaMemsetAsync(s->xyzr, 0, sizeof(node4), l.stream);
	cudaMemsetAsync(s->nmass, 0, sizeof(real), l.stream);
	cudaMemsetAsync(s->body_n, 0xff, n * sizeof(int), l.stream);	// TREE_NO_BODY for internal nodes
	l.bh = s;
	return NB200_OK;
}

// Leaf order (body_n[n..2n)) from the packed bodies in l.src.
static int bh_build_topology(nb200_ctx* ctx, nb200_lane& l, int& launches, std::string& err)
{
	bh_state*	s = l.bh;
	const int	n = static_cast<int>(s->n);
	if(n == 1)
	{
		BH_CU(cudaMemsetAsync(s->body_n + 1, 0, sizeof(int), l.stream));
		return NB200_OK;
	}
	if(n <= NB200_BH_LOCAL)
	{
		bh_local_build<NB200_BH_LOCAL><<<1, std::max(32, n / 2), 0, l.stream>>>(l.src, nullptr, s->body_n, n, n, 0);
		++launches;
		BH_CU(cudaGetLastError());
		return NB200_OK;
	}
	const unsigned g256 = static_cast<unsigned>((n + 255) / 256);
	for(int a = 0; a < 3; ++a)
	{
		bh_extract_keys<<<g256, 256, 0, l.stream>>>(l.src, s->keys_in, s->iota, n, a);
		++launches;
		size_t bytes = s->cub_bytes;
		BH_CU(cub::DeviceRadixSort::SortPairs(s->cub_tmp, bytes, s->keys_in, s->keys_out, s->iota, s->ord[a], n, 0,
											  static_cast<int>(sizeof(real) * 8), l.stream));
	}
	const int nblk = n / NB200_BH_PART_BLOCK;
	int depth = 0;
	for(int seg = n; seg > NB200_BH_LOCAL; seg >>= 1, ++depth)
	{
		const int c = depth % 3, a0 = (c + 1) % 3, a1 = (c + 2) % 3;
		// side flags along the split dimension: level 0 marks them from the presorted ordering; every later level got them
		// from the scatter pass of the level above (a0 of one level is the split dimension of the next)
		if(depth == 0)
		{
			bh_mark_side<<<g256, 256, 0, l.stream>>>(s->ord[c], s->side, n, seg);
			++launches;
		}
		const bool more = (seg >> 1) > NB200_BH_LOCAL;	// another global level follows
		// the ordering along the split dimension is already partitioned (its lower half IS the left child)
		bh_part_count<<<dim3(nblk, 2), 256, 0, l.stream>>>(s->ord[a0], s->ord[a1], s->side, s->blk, nblk);
		bh_part_scan<<<2, 1024, 0, l.stream>>>(s->blk, nblk);
		bh_part_scatter<<<dim3(nblk, 2), 256, 0, l.stream>>>(s->ord[a0], s->ord[a1], s->ord_tmp[a0], s->ord_tmp[a1], s->side,
															   s->blk, nblk, seg, more ? s->side_next : nullptr);
		launches += 3;
		std::swap(s->side, s->side_next);
		std::swap(s->ord[a0], s->ord_tmp[a0]);
		std::swap(s->ord[a1], s->ord_tmp[a1]);
	}
	BH_CU(cudaGetLastError());
	bh_local_build<NB200_BH_LOCAL><<<n / NB200_BH_LOCAL, NB200_BH_LOCAL / 2, 0, l.stream>>>(l.src, s->ord[depth % 3], s->body_n, n,
																							NB200_BH_LOCAL, depth);
	++launches;
	BH_CU(cudaGetLastError());
	return NB200_OK;
}

static int bh_update_geometry(nb200_ctx* ctx, nb200_lane& l, int& launches, std::string& err)
{
	bh_state*	s = l.bh;
	const int	n = static_cast<int>(s->n);
	const real	ratio_sqr = ctx->bh_ratio * ctx->bh_ratio;
	bh_update_leaves<<<(n + 255) / 256, 256, 0, l.stream>>>(l.src, s->body_n, s->xyzr, s->nmass, s->bmin, s->bmax, n);
	++launches;
	int level = n >> 1;
	for(; level > 256; level >>= 1)
	{
		bh_update_level<<<(level + 255) / 256, 256, 0, l.stream>>>(s->xyzr, s->nmass, s->bmin, s->bmax, level, ratio_sqr);
		++launches;
	}
	if(level >= 1)
	{
		bh_update_top<<<1, 256, 0, l.stream>>>(s->xyzr, s->nmass, s->bmin, s->bmax, level, ratio_sqr);
		++launches;
	}
	BH_CU(cudaGetLastError());
	return NB200_OK;
}

__global__ void __launch_bounds__(256) bh_iota(int* __restrict__ out, int count)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i < count) { out[i] = i; }
}

// Arrays of the longest-walk-first launch order, for `ctas` CTAs
static int bh_lpt_alloc(nb200_lane& l, bh_state* s, int ctas, int& launches, std::string& err)
{
	if(s->lpt_ctas == ctas) { return NB200_OK; }
	void* old[] = {s->lpt_cost, s->lpt_cost_sorted, s->lpt_iota, s->lpt_order, s->lpt_tmp};
	for(void* p : old) { if(p != nullptr) { cudaFree(p); } }
	s->lpt_cost = s->lpt_cost_sorted = nullptr;
	s->lpt_iota = s->lpt_order = nullptr;
	s->lpt_tmp = nullptr;
	s->lpt_ctas = 0;
	s->lpt_have_order = false;
	size_t bytes = 0;
	cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, s->lpt_cost, s->lpt_cost_sorted, s->lpt_iota, s->lpt_order, ctas);
	const size_t count = static_cast<size_t>(ctas);
	if(cudaMalloc(&s->lpt_cost, count * sizeof(unsigned)) != cudaSuccess || cudaMalloc(&s->lpt_cost_sorted, count * sizeof(unsigned)) != cudaSuccess ||
	   cudaMalloc(&s->lpt_iota, count * sizeof(int)) != cudaSuccess || cudaMalloc(&s->lpt_order, count * sizeof(int)) != cudaSuccess ||
	   cudaMalloc(&s->lpt_tmp, std::max<size_t>(bytes, 16)) != cudaSuccess)
	{
		cudaGetLastError();
		err = "walk order allocation failed";
		return NB200_ERR_ALLOC;
	}
	s->lpt_tmp_bytes = bytes;
	s->lpt_ctas = ctas;
	bh_iota<<<(ctas + 255) / 256, 256, 0, l.stream>>>(s->lpt_iota, ctas);
	++launches;
	BH_CU(cudaGetLastError());
	return NB200_OK;
}

static int bh_fcompute(nb200_ctx* ctx, nb200_lane& l, const real* y, real* f, size_t step, int& launches, std::string& err)
{
	int rc = bh_alloc(ctx, l, err);
	if(rc != NB200_OK) { return rc; }
	bh_state*	s = l.bh;
	const int	n = static_cast<int>(s->n);
	// same rebuild rule as nbody_engine_cuda_bh_tex.cpp:76-77
	const bool rebuild = ctx->bh_build_rate == 0 || !s->have_tree || (step % ctx->bh_build_rate) == 0;
	if(rebuild)
	{
		rc = bh_build_topology(ctx, l, launches, err);
		if(rc != NB200_OK) { return rc; }
		if(ctx->nshards > 1)
		{
			bh_leaf_positions<<<(n + 255) / 256, 256, 0, l.stream>>>(s->body_n, s->leaf_pos, n);
			++launches;
		}
		s->have_tree = true;
	}
	rc = bh_update_geometry(ctx, l, launches, err);
	if(rc != NB200_OK) { return rc; }
	if(ctx->opt_timing) { BH_CU(cudaEventRecord(l.ev_t[2], l.stream)); }

	unsigned long long* stats = nullptr;
	if(ctx->bh_stats)
	{
		BH_CU(cudaMemsetAsync(l.d_scalar + 2, 0, 2 * sizeof(unsigned long long), l.stream));
		BH_CU(cudaMemsetAsync(l.d_scalar + 8, 0, 16 * sizeof(unsigned long long), l.stream));
		stats = l.d_scalar;
	}
	const int	n_targets = static_cast<int>(ctx->n_shard);
	const int	shard_first = static_cast<int>(static_cast<size_t>(l.shard) * ctx->n_shard);
	// one shard: results go straight to f by body index; several: chunks of consecutive leaves dealt round-robin,
	// results into this shard's block of acc_all in the order walked
	real*		acc_leaf = ctx->nshards > 1 ? s->acc_all + static_cast<size_t>(l.shard) * 3 * ctx->n_shard : nullptr;
	const int	chunk = static_cast<int>(std::min<size_t>(NB200_BH_DEAL_CHUNK, ctx->n_shard));
	const int3	deal = make_int3(ctx->nshards > 1 ? chunk : n_targets, ctx->nshards, ctx->nshards > 1 ? l.shard : 0);
	const int	block = (ctx->opt_walk_threads == 64 || ctx->opt_walk_threads == 256) ? static_cast<int>(ctx->opt_walk_threads) : 128;
	const unsigned grid = static_cast<unsigned>((n_targets + block - 1) / block);
	if(ctx->opt_walk_mode == 1)
	{
		bh_walk_thread<<<grid, block, 0, l.stream>>>(s->xyzr, s->nmass, s->body_n, acc_leaf, deal, y, f, n, n_targets, ctx->n_shard, shard_first, stats);
	}
	else if(ctx->opt_walk_mode == 0 || ctx->opt_walk_mode == 8)
	{
		// grouped walk (nb200_bh_group.cuh), the default: a warp per 32 consecutive leaves, 4 warps per CTA (a deal chunk is a
		// multiple of 128 leaves or the whole shard, so a group's leaves stay consecutive)
		const unsigned	ggrid = static_cast<unsigned>((n_targets + 32 * NB200_BHG_WARPS - 1) / (32 * NB200_BHG_WARPS));
		// longest-walk-first CTA order from the previous walk's costs: automatic while a GPU has few waves of CTAs
		// (<= 8192 CTAs = 1M targets per GPU, e.g. 4M bodies over 4 or 8 GPUs); with many waves the tail is a small share
		// and launching kd-order neighbours together keeps shared upper-tree nodes in L1/L2
		const bool		lpt = ctx->opt_walk_lpt > 0 || (ctx->opt_walk_lpt < 0 && ggrid >= 1024 && ggrid <= 8192);
		const int*		order = nullptr;
		unsigned*		cost = nullptr;
		if(lpt)
		{
			rc = bh_lpt_alloc(l, s, static_cast<int>(ggrid), launches, err);
			if(rc != NB200_OK) { return rc; }
			order = s->lpt_have_order ? s->lpt_order : nullptr;
			cost = s->lpt_cost;
		}
		if(stats != nullptr) { bh_walk_group<true><<<ggrid, 32 * NB200_BHG_WARPS, 0, l.stream>>>(s->xyzr, s->nmass, s->body_n, acc_leaf, deal, y, f, n, n_targets, ctx->n_shard, shard_first, stats, order, cost); }
		else { bh_walk_group<false><<<ggrid, 32 * NB200_BHG_WARPS, 0, l.stream>>>(s->xyzr, s->nmass, s->body_n, acc_leaf, deal, y, f, n, n_targets, ctx->n_shard, shard_first, stats, order, cost); }
		if(lpt)
		{
			size_t bytes = s->lpt_tmp_bytes;
			BH_CU(cub::DeviceRadixSort::SortPairsDescending(s->lpt_tmp, bytes, s->lpt_cost, s->lpt_cost_sorted, s->lpt_iota, s->lpt_order,
															static_cast<int>(ggrid), 0, 32, l.stream));
			s->lpt_have_order = true;
		}
	}
	else if(ctx->opt_walk_mode != 32)
	{
		// several targets per lane (a deal chunk is a multiple of 128 leaves or the whole shard, so a warp's leaves stay
		// consecutive). Two is the measured optimum of this family (round 1's default): N = 4M, ratio 10 on one B200 -- FP64 675 -> 600 ms,
		// FP32 647 -> 438 ms; four: 754 / 439 ms (114 registers, a quarter of the warp slots)
		const int		tpl = ctx->opt_walk_mode == 4 ? 4 : 2;
		const unsigned	mgrid = static_cast<unsigned>((n_targets + 128 * tpl - 1) / (128 * tpl));
		// longest-walk-first order from the previous walk's costs. Automatic for 1024..4096 CTAs (262,144..1,048,576
		// targets per GPU, e.g. 4M bodies over 4 or 8 GPUs): below that the walk is short and the sort's launches cost more
		// than the tail they remove; above it there are many waves, the tail is a small share, and launching neighbours
		// in the kd order together (shared upper-tree nodes in L1/L2) is worth more. One B200, ratio 10: N = 524,288
		// 57.4 -> 50.8 ms, N = 1M 130.3 -> 125.0 ms, N = 4M 2.5 % slower
		const bool		lpt = tpl == 2 && (ctx->opt_walk_lpt > 0 || (ctx->opt_walk_lpt < 0 && mgrid >= 1024 && mgrid <= 4096));
		const int*		order = nullptr;
		unsigned*		cost = nullptr;
		if(lpt)
		{
			rc = bh_lpt_alloc(l, s, static_cast<int>(mgrid), launches, err);
			if(rc != NB200_OK) { return rc; }
			order = s->lpt_have_order ? s->lpt_order : nullptr;
			cost = s->lpt_cost;
		}
		if(tpl == 2 && stats != nullptr) { bh_walk_warp_multi<2, true, 8><<<mgrid, 128, 0, l.stream>>>(s->xyzr, s->nmass, s->body_n, acc_leaf, deal, y, f, n, n_targets, ctx->n_shard, shard_first, stats, order, cost); }
		else if(tpl == 2) { bh_walk_warp_multi<2, false, 8><<<mgrid, 128, 0, l.stream>>>(s->xyzr, s->nmass, s->body_n, acc_leaf, deal, y, f, n, n_targets, ctx->n_shard, shard_first, stats, order, cost); }
		else if(stats != nullptr) { bh_walk_warp_multi<4, true, 4><<<mgrid, 128, 0, l.stream>>>(s->xyzr, s->nmass, s->body_n, acc_leaf, deal, y, f, n, n_targets, ctx->n_shard, shard_first, stats, nullptr, nullptr); }
		else { bh_walk_warp_multi<4, false, 4><<<mgrid, 128, 0, l.stream>>>(s->xyzr, s->nmass, s->body_n, acc_leaf, deal, y, f, n, n_targets, ctx->n_shard, shard_first, stats, nullptr, nullptr); }
		if(lpt)
		{
			// next walk's order: this walk's costs, descending (stream-ordered after the walk; ~20 us at 16384 CTAs)
			size_t bytes = s->lpt_tmp_bytes;
			BH_CU(cub::DeviceRadixSort::SortPairsDescending(s->lpt_tmp, bytes, s->lpt_cost, s->lpt_cost_sorted, s->lpt_iota, s->lpt_order,
															static_cast<int>(mgrid), 0, 32, l.stream));
			s->lpt_have_order = true;
		}
	}
	else if(stats != nullptr)
	{
		bh_walk_warp<true><<<grid, block, 0, l.stream>>>(s->xyzr, s->nmass, s->body_n, acc_leaf, deal, y, f, n, n_targets, ctx->n_shard, shard_first, stats);
	}
	else
	{
		bh_walk_warp<false><<<grid, block, 0, l.stream>>>(s->xyzr, s->nmass, s->body_n, acc_leaf, deal, y, f, n, n_targets, ctx->n_shard, shard_first, stats);
	}
	++launches;
	BH_CU(cudaGetLastError());
	if(ctx->opt_timing)
	{
		BH_CU(cudaEventRecord(l.ev_t[3], l.stream));
		BH_CU(cudaEventRecord(l.ev_t[4], l.stream));
	}
	return NB200_OK;
}

// second half of a sharded fcompute, after acc_all has been gathered across shards
static int bh_scatter(nb200_ctx* ctx, nb200_lane& l, const real* y, real* f, int& launches, std::string& err)
{
	bh_state* s = l.bh;
	bh_scatter_own<<<static_cast<unsigned>((ctx->n_shard + 255) / 256), 256, 0, l.stream>>>(
		s->acc_all, s->leaf_pos, y, f, ctx->n_shard, static_cast<size_t>(l.shard) * ctx->n_shard,
		std::min<size_t>(NB200_BH_DEAL_CHUNK, ctx->n_shard), static_cast<size_t>(ctx->nshards));
	++launches;
	BH_CU(cudaGetLastError());
	if(ctx->opt_timing) { BH_CU(cudaEventRecord(l.ev_t[4], l.stream)); }
	return NB200_OK;
}

static int bh_export(nb200_ctx* ctx, nb200_lane& l, real* xyzr, real* mass, int* body_n, std::string& err)
{
	bh_state* s = l.bh;
	const size_t ts = 2 * s->n;
	if(xyzr != nullptr) { BH_CU(cudaMemcpyAsync(xyzr, s->xyzr, ts * sizeof(node4), cudaMemcpyDeviceToHost, l.stream)); }
	if(mass != nullptr) { BH_CU(cudaMemcpyAsync(mass, s->nmass, ts * sizeof(real), cudaMemcpyDeviceToHost, l.stream)); }
	if(body_n != nullptr) { BH_CU(cudaMemcpyAsync(body_n, s->body_n, ts * sizeof(int), cudaMemcpyDeviceToHost, l.stream)); }
	BH_CU(cudaStreamSynchronize(l.stream));
	(void)ctx;
	return NB200_OK;
}

#endif // NB200_BH_CUH
