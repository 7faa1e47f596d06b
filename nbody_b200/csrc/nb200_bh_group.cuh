// nb200 -- Barnes-Hut walk, grouped form (the default): lanes work on NODES while deciding, on TARGETS while summing.
//
// The warp-coherent walks of nb200_bh.cuh keep the reference's shape -- one traversal state per target -- and pay the
// bookkeeping of a visit (node load, skip_idx algebra, votes, loop control: ~45 issue slots) once per node and warp
// for 16 slots of acceptance arithmetic. Here a warp owns a GROUP of 32 consecutive leaves and works in rounds:
//
//   decide  A work item is (parent p, mask M): "the targets in M opened p, so its children 2p and 2p+1 are theirs to
//           test". Each LANE takes one item off the warp's stack in shared memory, loads the sibling pair (64
//           contiguous bytes, and the two masses) and tests it against all 32 targets of the group (positions
//           broadcast from shared memory, two targets per packed f32x2 instruction). The result is one accept word and
//           one open word per child: accepted nodes go to the warp's interaction list TOGETHER WITH THEIR DATA
//           (mass centre, mass, target mask), opened internal children become new items. Bookkeeping is per lane and
//           per item, i.e. amortised over 64 acceptance tests, and 32 items are in flight per warp instead of one node.
//   sum     Between the node loads of a round and its tests -- i.e. while those loads are in flight -- the warp
//           switches to one target per lane and runs down the list of the PREVIOUS round (at most 64 entries, read
//           from shared memory only, one broadcast per entry): the lanes named in an entry's mask add the node's
//           attraction in nbcoord_t with the formulas of node_force_from_test; the others add exactly +-0 (no branch).
//
// Which nodes a target accepts is decided exactly as in nbody_space_heap_stackless::traverse
// (nbody_space_heap_stackless.cpp:3-28): d2 > radius_sqr on the node, else its children. What changes is only the
// ORDER in which a target's accepted nodes are added (by blocks of the stack instead of depth-first), so results
// agree with the other walks to rounding (~1e-15 relative, tested <= 1e-13), not bit for bit; they are still a pure
// function of the inputs (no atomics, no scheduling dependence) and identical for every shard count.
//
// FP64 build -- certified FP32 decisions. The 4.1e11 acceptance tests of an N = 4M walk would cost as many FP64-pipe
// slots as the force sums themselves. Each test is therefore evaluated in FP32 on coordinates RELATIVE to the group's
// first target (the subtraction is done in FP64 once per node and lane, then rounded), which makes the error of
// t = d2 - radius_sqr a few ulp of d2: with u = 2^-24, R = the group's extent, |t_fp32 - t_exact| <= u (17.4 w + 3.5 R^2)
// at the decision boundary d2 = w (derivation in DESIGN.md 3.4). A test with |t| > m = 32 u (w + R^2) is therefore
// decided by the sign of t; anything closer to the boundary (4e-4 of the lane-items) is re-evaluated by the lane in
// FP64 with the reference's own expression (bh_d2). Visit and interaction counts equal the oracle's (tested),
// knife-edge cells included.
// FP32 build: the reference's own arithmetic is FP32, so the packed tests ARE the reference expression
// (fma(dz,dz,fma(dx,dx,dy*dy)) > w on absolute coordinates) and need no certificate.
//
// Cost model: an FP64 instruction holds the dispatch port for two cycles, anything else for one. Per
// round: decide 417 instructions, sum 39 entries x (16 FP64 + 7.6 others). N = 4M, ratio 10, one B200: 393 ms.
#ifndef NB200_BH_GROUP_CUH
#define NB200_BH_GROUP_CUH

#define NB200_BHG_WARPS 4			// groups (warps) per CTA
#ifndef NB200_BHG_STACK
#define NB200_BHG_STACK 512			// work-item slots per warp (deepest fill measured at N = 4M, ratio 10: 372)
#endif
#define NB200_BHG_STACK_SOFT (NB200_BHG_STACK - 64)	// above this fill items are taken one at a time (depth-first: growth <= tree depth)
#define NB200_BHG_LIST 64			// interaction-list entries per warp: one round of 32 sibling pairs
#ifndef NB200_BHG_MINB
#define NB200_BHG_MINB 5
#endif
#ifndef NB200_BHG_UNROLL
#define NB200_BHG_UNROLL 8			// entries of the sum loop in flight per lane
#endif

typedef unsigned long long bhg_f32x2;
__device__ __forceinline__ bhg_f32x2 bhg_pack(float lo, float hi)
{
	bhg_f32x2 r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
	return r;
}
__device__ __forceinline__ void bhg_unpack(bhg_f32x2 v, float& lo, float& hi)
{
	asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ bhg_f32x2 bhg_fma(bhg_f32x2 a, bhg_f32x2 b, bhg_f32x2 c)
{
	bhg_f32x2 d;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
	return d;
}
__device__ __forceinline__ bhg_f32x2 bhg_mul(bhg_f32x2 a, bhg_f32x2 b)
{
	bhg_f32x2 d;
	asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
	return d;
}
__device__ __forceinline__ bhg_f32x2 bhg_sub(bhg_f32x2 a, bhg_f32x2 b)
{
	bhg_f32x2 d;
	asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
	return d;
}

struct bhg_warp_smem
{
	int2				stack[NB200_BHG_STACK];	// {parent, mask of targets that opened it}
	body4				lnode[NB200_BHG_LIST];	// interaction list: accepted node {mass centre, mass} ...
	unsigned			lmask[NB200_BHG_LIST];	// ... and the targets that accepted it
	ulonglong2			tp_xy[16];				// targets 2i, 2i+1: {x pair, y pair} as packed f32x2
	unsigned long long	tp_z[16];				// {z pair}
#if NB200_PRECISION == 2
	double				tpd[3][32];				// absolute FP64 positions (exact re-evaluation)
#endif
};

// One accepted node for one target. Same formulas as node_force_from_test, with m r^-3 taken from the seed in one series
// step (FP64 build): m y0^3 (1 + e (3/2 + 15/8 e)), e = 1 - r^2 y0^2 exact to one rounding (the seed has 21 significant
// bits, so y0^2 is exact) -- 7 FP64 operations instead of 8, truncation error 35/16 e^3 < 2^-56. On this pipe an FP64
// instruction holds the dispatch port for two cycles and everything else for one (profiles/microbench/bh_sum_loop.cu),
// so what counts is 2 x FP64 + other instructions per entry.
// A lane the node is not for (`mine` == 0) adds it with a zero coefficient, i.e. exactly +-0: no branch, so the
// dependent chains of several entries interleave; the zero enters through the seed, whose low word is zero anyway
// (one select; ptxas turns predicated DFMAs into three 64-bit selects).
// CLAMP = false leaves out the max(r^2, MinDistance) of nbody_data::force (nbody_data.cpp:39-42): only for rounds in
// which no accepted pair can be closer than 1e-4.
template<bool CLAMP>
__device__ __forceinline__ void bhg_force(real dx, real dy, real dz, real m, unsigned mine, real& ax, real& ay, real& az)
{
	real d2 = bh_d2(dx, dy, dz);
#if NB200_PRECISION == 2
	if(CLAMP)
	{
		long long		bits = __double_as_longlong(d2);
		const long long	min_bits = 0x3E45798EE2308C3ALL;	// 1e-8
		bits = bits < min_bits ? min_bits : bits;
		d2 = __longlong_as_double(bits);
	}
	double	seed;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(seed) : "d"(d2));
	const double y0 = __hiloint2double(mine != 0 ? __double2hiint(seed) : 0, 0);
	const double y2 = y0 * y0;
	const double e = fma(-d2, y2, 1.0);
	const double u = e * fma(e, 1.875, 1.5);
	const double g = (m * y0) * y2;
	const double c = fma(g, u, g);
#else
	if(CLAMP) { d2 = fmaxf(d2, NB200_MIN_DISTANCE); }
	const float yv = mine != 0 ? nb200_rsqrt_normal(d2) : 0.0f;
	const float c = (yv * yv) * (m * yv);
#endif
	ax = fma(-dx, c, ax);
	ay = fma(-dy, c, ay);
	az = fma(-dz, c, az);
}

// sum phase: one target per lane runs down the warp's interaction list. The list holds the nodes themselves (written by
// the lane that tested them), so this loop touches shared memory only: every read is one broadcast.
template<bool CLAMP>
__device__ __forceinline__ void bhg_flush(const bhg_warp_smem& sm, int count, unsigned lane_bit, real px, real py, real pz,
										  real& ax, real& ay, real& az)
{
	constexpr int in_flight = NB200_BHG_UNROLL;
#pragma unroll in_flight
	for(int e = 0; e < count; ++e)
	{
		const unsigned	mask = sm.lmask[e];
		const body4		nd = sm.lnode[e];
		bhg_force<CLAMP>(px - nd.x, py - nd.y, pz - nd.z, nd.m, mask & lane_bit, ax, ay, az);
	}
}

// the masses of a sibling pair in one load
__device__ __forceinline__ void bhg_load_mass_pair(const real* __restrict__ nmass, int left, real& ml, real& mr)
{
#if NB200_PRECISION == 2
	const double2 v = __ldg(reinterpret_cast<const double2*>(nmass + left));
#else
	const float2 v = __ldg(reinterpret_cast<const float2*>(nmass + left));
#endif
	ml = v.x;
	mr = v.y;
}

template<bool STATS>
__global__ void __launch_bounds__(32 * NB200_BHG_WARPS, NB200_BHG_MINB)
bh_walk_group(const node4* __restrict__ xyzr, const real* __restrict__ nmass, const int* __restrict__ body_n,
			  real* __restrict__ acc_leaf, int3 deal, const real* __restrict__ y, real* __restrict__ f, int n, int n_targets,
			  size_t n_shard, int shard_first, unsigned long long* __restrict__ stats, const int* __restrict__ cta_order,
			  unsigned* __restrict__ cta_cost)
{
	__shared__ bhg_warp_smem	sm_all[NB200_BHG_WARPS];
	__shared__ unsigned			t_begin;
	if(cta_cost != nullptr && threadIdx.x == 0) { t_begin = static_cast<unsigned>(clock64()); }
	const int		cta = cta_order != nullptr ? cta_order[blockIdx.x] : static_cast<int>(blockIdx.x);
	const int		lane = threadIdx.x & 31;
	const int		group = cta * NB200_BHG_WARPS + (threadIdx.x >> 5);
	bhg_warp_smem&	sm = sm_all[threadIdx.x >> 5];
	const unsigned	full = 0xffffffffu;
	if(group * 32 < n_targets)	// warp-uniform; warp 0 of every launched CTA has targets
	{
		const int	t = group * 32 + lane;
		const bool	live = t < n_targets;
		const int	tc = live ? t : n_targets - 1;	// idle lanes shadow the last target, are in no mask and never store
		// deal = {chunk, shards, shard}: target t of this shard is leaf ((t / chunk) * shards + shard) * chunk + t % chunk
		const int	leaf = n + ((tc / deal.x) * deal.y + deal.z) * deal.x + tc % deal.x;
		const node4	me = load_node(xyzr, leaf);
		real		ax = 0, ay = 0, az = 0;
		unsigned	visits = 0, inter = 0;
		unsigned	pr_rounds = 0, pr_items = 0, pr_entries = 0, pr_unsure = 0, pr_close = 0, pr_maxsp = 0, pr_trips = 0, pr_mine = 0;	// STATS only
		unsigned	pr_hist[5] = {0, 0, 0, 0, 0};	// entries by targets per mask: 32, 24..31, 16..23, 8..15, 1..7
#if NB200_PRECISION == 2
		// group frame: origin = first target; R = largest |relative coordinate| in the group
		const double	ox = __shfl_sync(full, me.x, 0), oy = __shfl_sync(full, me.y, 0), oz = __shfl_sync(full, me.z, 0);
		const float		fx = static_cast<float>(me.x - ox), fy = static_cast<float>(me.y - oy), fz = static_cast<float>(me.z - oz);
		float			rad = fmaxf(fabsf(fx), fmaxf(fabsf(fy), fabsf(fz)));
#pragma unroll
		for(int o = 16; o > 0; o >>= 1) { rad = fmaxf(rad, __shfl_xor_sync(full, rad, o)); }
		const float		MU = 1.9073486328125e-6f;	// 32 * 2^-24
		const float		MUR2 = MU * (rad * rad) * 1.001f;
		// a pair closer than MinDistance (d2 < 1e-8) on a node with radius_sqr < 1e-8 has t = d2 - radius_sqr < 1e-8 and an
		// FP32 error below u (7 R d + 10 d^2 + 4 w) < 5e-11 R: every such test shows |t| < CLOSE
		const float		CLOSE = 2.0e-8f + 1.0e-10f * rad;
		sm.tpd[0][lane] = me.x;
		sm.tpd[1][lane] = me.y;
		sm.tpd[2][lane] = me.z;
#else
		const float		fx = me.x, fy = me.y, fz = me.z;
#endif
		{
			float* xy = reinterpret_cast<float*>(&sm.tp_xy[lane >> 1]);
			xy[lane & 1] = fx;
			xy[2 + (lane & 1)] = fy;
			reinterpret_cast<float*>(&sm.tp_z[lane >> 1])[lane & 1] = fz;
		}
		const unsigned	live_mask = __ballot_sync(full, live);
		const unsigned	lane_bit = 1u << lane;
		int				sp = 0, nl = 0;	// warp-uniform fill of the stack and of the list
		// warp-uniform: the list may hold an accepted (target, node) pair closer than MinDistance, so the next sum needs the
		// clamp of r^2. d2 > radius_sqr >= 0 for every accepted pair, so that takes a node with radius_sqr < 1e-8 (a leaf)
		// AND a target within 1e-4 of it: known from the smallest |d2 - radius_sqr| the lane saw (FP64 build; the group's
		// own leaves raise it, a few % of the rounds). The FP32 build's clamp is one FMNMX and stays on.
		bool			close_pairs = true;	// the root's entry is not examined
		{
			// the root is visited by every target (curr = 1 at the start of traverse)
			const node4	nd = load_node(xyzr, 1);
			const real	dx = me.x - nd.x, dy = me.y - nd.y, dz = me.z - nd.z;
			const bool	acc = live && (bh_d2(dx, dy, dz) > nd.w);
			const unsigned A = __ballot_sync(full, acc);
			const unsigned O = live_mask & ~A;
			if(A != 0)
			{
				if(lane == 0)
				{
					body4 b;
					b.x = nd.x; b.y = nd.y; b.z = nd.z; b.m = nmass[1];
					sm.lnode[0] = b;
					sm.lmask[0] = A;
				}
				nl = 1;
			}
			if(O != 0 && n > 1)
			{
				if(lane == 0) { sm.stack[0] = make_int2(1, static_cast<int>(O)); }
				sp = 1;
			}
			if(STATS && lane == 0)
			{
				visits += __popc(live_mask);
				inter += __popc(A);
			}
		}
		__syncwarp();
		while(sp > 0)
		{
			// ---- decide: one work item per lane ----
			const int	cnt = sp > NB200_BHG_STACK_SOFT ? 1 : min(32, sp);
			const bool	have = lane < cnt;
			const int2	it = have ? sm.stack[sp - 1 - lane] : make_int2(0, 0);
			sp -= cnt;
			const int		L = it.x << 1;	// children L, L + 1 (idle lanes: the unused slot 0 and the root, masked out)
			const unsigned	M = static_cast<unsigned>(it.y);
			const node4		ndL = load_node(xyzr, L), ndR = load_node(xyzr, L + 1);
			real			massL, massR;
			bhg_load_mass_pair(nmass, L, massL, massR);
			// ---- sum: while those loads are in flight, every target adds the nodes accepted in the previous round ----
			if(STATS)
			{
				// what a per-lane compaction of this round's entries would cost: the largest number of entries any one target takes
				unsigned mine = 0;
				for(int e = 0; e < nl; ++e)
				{
					const unsigned mk = sm.lmask[e];
					mine += (mk & lane_bit) ? 1u : 0u;
					const int pc = __popc(mk);
					pr_hist[pc == 32 ? 0 : (pc >= 24 ? 1 : (pc >= 16 ? 2 : (pc >= 8 ? 3 : 4)))] += 1;
				}
				pr_trips += __reduce_max_sync(full, mine);
				pr_mine += mine;
				pr_rounds += 1;
				pr_items += cnt;
				pr_entries += nl;
				pr_close += close_pairs ? 1u : 0u;
				pr_maxsp = max(pr_maxsp, static_cast<unsigned>(sp + cnt));
			}
			if(close_pairs) { bhg_flush<true>(sm, nl, lane_bit, me.x, me.y, me.z, ax, ay, az); }
			else { bhg_flush<false>(sm, nl, lane_bit, me.x, me.y, me.z, ax, ay, az); }
			nl = 0;
			__syncwarp();
			unsigned		SL = 0, SR = 0;	// bit j = 1: target j does NOT accept the child
#if NB200_PRECISION == 2
			const float		wL = static_cast<float>(ndL.w), wR = static_cast<float>(ndR.w);
			const float		mL = fmaf(wL, MU, MUR2), mR = fmaf(wR, MU, MUR2);
			float			qx = static_cast<float>(ndL.x - ox), qy = static_cast<float>(ndL.y - oy), qz = static_cast<float>(ndL.z - oz);
			const bhg_f32x2	cLx = bhg_pack(qx, qx), cLy = bhg_pack(qy, qy), cLz = bhg_pack(qz, qz), nwL = bhg_pack(-wL, -wL);
			qx = static_cast<float>(ndR.x - ox); qy = static_cast<float>(ndR.y - oy); qz = static_cast<float>(ndR.z - oz);
			const bhg_f32x2	cRx = bhg_pack(qx, qx), cRy = bhg_pack(qy, qy), cRz = bhg_pack(qz, qz), nwR = bhg_pack(-wR, -wR);
			float			uL = 3.0e38f, uR = 3.0e38f;	// smallest |t| seen per child
#pragma unroll
			for(int i = 15; i >= 0; --i)
			{
				const ulonglong2			XY = sm.tp_xy[i];
				const unsigned long long	Z = sm.tp_z[i];
				float t0, t1;
				bhg_f32x2 dx = bhg_sub(XY.x, cLx), dy = bhg_sub(XY.y, cLy), dz = bhg_sub(Z, cLz);
				bhg_f32x2 tt = bhg_fma(dz, dz, bhg_fma(dy, dy, bhg_fma(dx, dx, nwL)));	// d2 - w
				bhg_unpack(tt, t0, t1);
				SL = __funnelshift_l(__float_as_uint(t1), SL, 1);
				SL = __funnelshift_l(__float_as_uint(t0), SL, 1);
				uL = fminf(uL, fminf(fabsf(t0), fabsf(t1)));
				dx = bhg_sub(XY.x, cRx); dy = bhg_sub(XY.y, cRy); dz = bhg_sub(Z, cRz);
				tt = bhg_fma(dz, dz, bhg_fma(dy, dy, bhg_fma(dx, dx, nwR)));
				bhg_unpack(tt, t0, t1);
				SR = __funnelshift_l(__float_as_uint(t1), SR, 1);
				SR = __funnelshift_l(__float_as_uint(t0), SR, 1);
				uR = fminf(uR, fminf(fabsf(t0), fabsf(t1)));
			}
			// some test of this item fell inside the margin (or is NaN): the lane redoes its 64 tests in FP64, exactly
			const bool unsure = have && !(uL > mL && uR > mR);
			if(STATS) { pr_unsure += __popc(__ballot_sync(full, unsure)); }
			if(__any_sync(full, unsure))
			{
				if(unsure)
				{
					unsigned eL = 0, eR = 0;
#pragma unroll 1
					for(int j = 0; j < 32; ++j)
					{
						const double px = sm.tpd[0][j], py = sm.tpd[1][j], pz = sm.tpd[2][j];
						double dx = px - ndL.x, dy = py - ndL.y, dz = pz - ndL.z;
						eL |= (bh_d2(dx, dy, dz) > ndL.w) ? (1u << j) : 0u;
						dx = px - ndR.x; dy = py - ndR.y; dz = pz - ndR.z;
						eR |= (bh_d2(dx, dy, dz) > ndR.w) ? (1u << j) : 0u;
					}
					SL = ~eL;
					SR = ~eR;
				}
			}
#else
			const bhg_f32x2	cLx = bhg_pack(ndL.x, ndL.x), cLy = bhg_pack(ndL.y, ndL.y), cLz = bhg_pack(ndL.z, ndL.z), wL = bhg_pack(ndL.w, ndL.w);
			const bhg_f32x2	cRx = bhg_pack(ndR.x, ndR.x), cRy = bhg_pack(ndR.y, ndR.y), cRz = bhg_pack(ndR.z, ndR.z), wR = bhg_pack(ndR.w, ndR.w);
			unsigned		AL_ = 0, AR_ = 0;
#pragma unroll
			for(int i = 15; i >= 0; --i)
			{
				const ulonglong2			XY = sm.tp_xy[i];
				const unsigned long long	Z = sm.tp_z[i];
				float t0, t1;
				// d2 exactly as bh_d2 rounds it; w - d2 is negative iff d2 > w (no flush to zero: the sign is exact)
				bhg_f32x2 dx = bhg_sub(XY.x, cLx), dy = bhg_sub(XY.y, cLy), dz = bhg_sub(Z, cLz);
				bhg_f32x2 tt = bhg_sub(wL, bhg_fma(dz, dz, bhg_fma(dx, dx, bhg_mul(dy, dy))));
				bhg_unpack(tt, t0, t1);
				AL_ = __funnelshift_l(__float_as_uint(t1), AL_, 1);
				AL_ = __funnelshift_l(__float_as_uint(t0), AL_, 1);
				dx = bhg_sub(XY.x, cRx); dy = bhg_sub(XY.y, cRy); dz = bhg_sub(Z, cRz);
				tt = bhg_sub(wR, bhg_fma(dz, dz, bhg_fma(dx, dx, bhg_mul(dy, dy))));
				bhg_unpack(tt, t0, t1);
				AR_ = __funnelshift_l(__float_as_uint(t1), AR_, 1);
				AR_ = __funnelshift_l(__float_as_uint(t0), AR_, 1);
			}
			SL = ~AL_;
			SR = ~AR_;
#endif
			const unsigned	AL = ~SL & M, AR = ~SR & M;	// accepted by
			const unsigned	OL = SL & M, OR_ = SR & M;	// opened by
			const unsigned	lt = lane_bit - 1u;
#if NB200_PRECISION == 2
			close_pairs = __any_sync(full, (AL != 0 && uL < CLOSE) || (AR != 0 && uR < CLOSE));
#endif
			{
				// accepted nodes -> interaction list (left children first, lanes in order)
				const unsigned bL = __ballot_sync(full, AL != 0), bR = __ballot_sync(full, AR != 0);
				const int nL = __popc(bL);
				if(AL != 0)
				{
					const int at = __popc(bL & lt);
					body4 b;
					b.x = ndL.x; b.y = ndL.y; b.z = ndL.z; b.m = massL;
					sm.lnode[at] = b;
					sm.lmask[at] = AL;
				}
				if(AR != 0)
				{
					const int at = nL + __popc(bR & lt);
					body4 b;
					b.x = ndR.x; b.y = ndR.y; b.z = ndR.z; b.m = massR;
					sm.lnode[at] = b;
					sm.lmask[at] = AR;
				}
				nl = nL + __popc(bR);
			}
			{
				// opened internal children -> new work items (a leaf that is not accepted is the target itself, or a body
				// at the same place: it has no children and drops out, as next_up() falls back to skip_idx() there)
				const bool inner = L < n;
				const unsigned bR = __ballot_sync(full, inner && OR_ != 0), bL = __ballot_sync(full, inner && OL != 0);
				const int nR = __popc(bR);
				if(inner && OR_ != 0) { sm.stack[sp + __popc(bR & lt)] = make_int2(L + 1, static_cast<int>(OR_)); }
				if(inner && OL != 0) { sm.stack[sp + nR + __popc(bL & lt)] = make_int2(L, static_cast<int>(OL)); }
				sp += nR + __popc(bL);
			}
			if(STATS)
			{
				visits += 2 * __popc(M);
				inter += __popc(AL) + __popc(AR);
			}
			__syncwarp();
		}
		if(close_pairs) { bhg_flush<true>(sm, nl, lane_bit, me.x, me.y, me.z, ax, ay, az); }
		else { bhg_flush<false>(sm, nl, lane_bit, me.x, me.y, me.z, ax, ay, az); }
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
		if(STATS && (visits | inter) != 0)
		{
			atomicAdd(stats + 2, static_cast<unsigned long long>(visits));
			atomicAdd(stats + 3, static_cast<unsigned long long>(inter));
		}
		if(STATS)
		{
			// the busiest target's entries over the whole walk (what a perfectly elastic per-lane compaction would still pay);
			// the last round's entries are left out of this one counter
			const unsigned busiest = __reduce_max_sync(full, pr_mine);
			if(lane == 0) { atomicAdd(stats + 15, static_cast<unsigned long long>(busiest)); }
		}
		if(STATS && lane == 0)
		{
			// walk profile (nb200_bh_walk_profile): rounds, items, list entries, lane-items redone in FP64, rounds summed
			// with the clamp, deepest stack fill, sum over rounds of the busiest target's entries
			atomicAdd(stats + 8, static_cast<unsigned long long>(pr_rounds));
			atomicAdd(stats + 9, static_cast<unsigned long long>(pr_items));
			atomicAdd(stats + 10, static_cast<unsigned long long>(pr_entries + nl));
			atomicAdd(stats + 11, static_cast<unsigned long long>(pr_unsure));
			atomicAdd(stats + 12, static_cast<unsigned long long>(pr_close));
			atomicMax(stats + 13, static_cast<unsigned long long>(pr_maxsp));
			atomicAdd(stats + 14, static_cast<unsigned long long>(pr_trips));
			for(int q = 0; q < 5; ++q) { atomicAdd(stats + 16 + q, static_cast<unsigned long long>(pr_hist[q])); }
		}
	}
	if(cta_cost != nullptr && threadIdx.x == 0)
	{
		cta_cost[cta] = static_cast<unsigned>(clock64()) - t_begin;
	}
}

#endif // NB200_BH_GROUP_CUH
