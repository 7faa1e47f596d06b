// nb200 -- shared types for the sm_100a kernels and the C ABI (include/nb200.h).
#ifndef NB200_COMMON_CUH
#define NB200_COMMON_CUH

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include <unordered_set>

#include "../../include/nb200.h"

typedef nb200_real real;

// Packed source body: position + mass, the unit the all-pairs kernel streams
// through shared memory (one 32-byte / 16-byte vector per body).
#if NB200_PRECISION == 1
struct alignas(16) body4 { float x, y, z, m; };
#define NB200_MIN_DISTANCE 1e-8f
#else
struct alignas(32) body4 { double x, y, z, m; };
#define NB200_MIN_DISTANCE 1e-8
#endif

// nbody::MinDistance (nbody/nbtype.h:78): r^2 is clamped to this value.

#ifdef __CUDACC__
// 1/sqrt(x) for x known to be a normal number (every call site has clamped r^2 to MinDistance = 1e-8 first): the bare
// MUFU.RSQ. rsqrtf() wraps the same instruction in a denormal-input rescue (FSETP + two predicated FMUL: 4 issue slots
// instead of 1) that can never trigger here; for normal inputs the results are identical.
__device__ __forceinline__ float nb200_rsqrt_normal(float x)
{
	float y;
	asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
	return y;
}
#endif

#define NB200_MAX_TERMS 48  // fused fmaddn terms per launch (rkfeagin14 needs 35)

struct nb200_terms
{
	const real*	p[NB200_MAX_TERMS];
	real		c[NB200_MAX_TERMS];
	int			n;
};

#define NB200_BUF_MAGIC 0x6e62323030627566ULL

struct bh_state;

struct nb200_lane
{
	int				dev = 0;
	int				shard = 0;          // global shard index of this lane
	int				sm_count = 148;
	cudaStream_t	stream = nullptr;
	cudaEvent_t		ev_packed = nullptr;   // "my packed shard is ready" (local gather)
	cudaEvent_t		ev_gathered = nullptr; // "I finished reading peers' shards"
	cudaEvent_t		ev_t[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // phase timing of the last fcompute
	cudaEvent_t		ev_mark[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // nb200_mark
	real*			mass = nullptr;     // N masses (full copy)
	body4*			src = nullptr;      // packed sources for all N bodies (+ zero-mass padding)
	real*			partial = nullptr;  // [S][3][n_shard] partial accelerations (direct, S>1)
	size_t			partial_elems = 0;
	size_t			small_smem = 0;     // dynamic shared memory direct_small has been allowed so far
	// symmetric-tile path (nb200_direct_sym.cuh)
	void*			sym_tiles = nullptr;	// int2 {row block, column block} of this rank's tiles
	real*			sym_prow = nullptr;		// [tiles][3][T] row sums
	real*			sym_pcol = nullptr;		// [tiles][3][T] column sums
	real*			sym_acc = nullptr;		// [shard][3][n_shard] reduced accelerations (+ [3][n_shard] receive block)
	size_t			sym_ntiles = 0;
	int				sym_edge = 0;
	unsigned long long*	d_scalar = nullptr;	// device scratch (32 x u64): [0] maxabs bits, [2..3] walk counters, [4] probe, [8..15] walk profile
	unsigned long long*	h_scalar = nullptr;	// pinned mirror
	bh_state*		bh = nullptr;
	void*			lane_comm = nullptr;	// ncclComm_t of this lane when the lanes of one process exchange shards with NCCL
	real*			read_scratch = nullptr;	// [shards][6][n_shard] staging of a multi-rank read_buffer
	size_t			read_scratch_bytes = 0;
};

struct nccl_api;
struct step_graph;

struct nb200_ctx
{
	std::vector<nb200_lane>	lanes;
	int			rank = 0;
	int			nranks = 1;
	int			nshards = 1;
	int			first_shard = 0;
	size_t		n = 0;          // bodies
	size_t		n_shard = 0;    // bodies per shard
	size_t		n_pad = 0;      // packed source count incl. padding (multiple of the plain kernel's tile)
	size_t		n_alloc = 0;    // allocated packed sources (multiple of the largest symmetric tile edge)
	nccl_api*	nccl = nullptr;
	void*		comm = nullptr; // ncclComm_t
	std::unordered_set<const nb200_buf*>	live;
	unsigned long long	launches = 0;
	int			last_direct_path = 0;
	bool		sym_unavailable = false;	// the symmetric tiles' scratch did not fit on some shard: ordered pairs until the body set changes
	bool		lanes_nccl = false;	// lanes of one process: NCCL group calls instead of peer copies / peer loads (option use_nccl)
	bool		peer_loads = true;	// every lane can load from every other lane's memory (same device or peer access enabled)
	std::string	err;
	step_graph*	sg = nullptr;	// solver steps as CUDA graphs (nb200_stepgraph.cuh)
	// Barnes-Hut configuration
	real		bh_ratio = 10;
	int			bh_layout = NB200_TREE_HEAP_STACKLESS;
	size_t		bh_build_rate = 0;
	bool		bh_stats = false;
	// tunables (0 = automatic)
	long long	opt_direct_ipt = 0;
	long long	opt_direct_segments = 0;
	long long	opt_walk_mode = 0;	// 0 = automatic (grouped walk), 8 = grouped walk, 1 = one thread per target, 2 / 4 = targets per lane, 32 = one target per lane
	long long	opt_walk_threads = 0;
	long long	opt_walk_lpt = -1;	// longest-walk-first CTA order: -1 automatic (>= 1024 CTAs), 0 off, 1 on
	long long	opt_timing = 1;
	long long	opt_direct_sym = -1;	// -1 auto, 0 off, 1 force
	long long	opt_direct_small = -1;	// single-launch kernel for small systems: -1 auto (N <= 4096, one shard), 0 off, 1 force
	long long	opt_sym_tile = 0;		// tile edge override (multiple of 256)
	// bodies per lane (row x column): 0: 8 x 1, 1: 4 x 2, 2: 8 x 2, 3: 4 x 4; FP32 only: 4: 8 x 2 packed f32x2, 5: 4 x 2 packed.
	// Fastest measured: 8 x 1 (FP64, with the clamp-free pass), 8 x 2 packed (FP32)
	long long	opt_sym_shape = sizeof(real) == 8 ? 0 : 4;
};

struct nb200_buf
{
	unsigned long long	magic = NB200_BUF_MAGIC;
	nb200_ctx*			owner = nullptr;
	size_t				bytes = 0;        // logical size
	bool				sharded = false;  // state vector: 6 rows x n_shard per lane
	size_t				lane_elems = 0;   // real elements held by each lane
	size_t				lane_bytes = 0;
	std::vector<void*>	dptr;             // one allocation per lane
};

#endif // NB200_COMMON_CUH
