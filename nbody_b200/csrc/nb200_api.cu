// nb200 -- implementation of the C ABI declared in include/nb200.h.
//
// Host-side structure (one context = the devices of one process):
//   lanes    one per entry of the device list; each has its own stream and
//            holds one body shard of every state vector
//   buffers  per-lane cudaMalloc allocations behind an opaque handle
//   ops      loop over lanes, launch asynchronously on the lane's stream
// The reference does the same work with one OpenMP host thread per device and
// fully replicated buffers (nbody/nbody_engine_cuda.cpp:226-241,
// nbody_engine_cuda_memory.cpp:4-36); here launches are asynchronous, so one
// host thread is enough, and state stays sharded.
#include <stdarg.h>
#include <algorithm>

#include "nb200_common.cuh"
#include "nb200_comm.cuh"
#include "nb200_direct.cuh"
#include "nb200_direct_sym.cuh"
#include "nb200_stateops.cuh"
#include "nb200_bh.cuh"
#include "nb200_stats.cuh"
#include "nb200_stepgraph.cuh"

#define NB200_API extern "C" __attribute__((visibility("default")))

namespace {

int fail(nb200_ctx* ctx, int status, const char* fmt, ...)
{
	char	text[512];
	va_list	ap;
	va_start(ap, fmt);
	vsnprintf(text, sizeof(text), fmt, ap);
	va_end(ap);
	if(ctx != nullptr)
	{
		ctx->err = text;
	}
	return status;
}

#define CU(ctx, call)                                                                                  \
	do                                                                                                 \
	{                                                                                                  \
		cudaError_t cu_res_ = (call);                                                                  \
		if(cu_res_ != cudaSuccess)                                                                     \
		{                                                                                              \
			return fail(ctx, NB200_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,                \
						cudaGetErrorString(cu_res_));                                                  \
		}                                                                                              \
	} while(0)

#define NC(ctx, call)                                                                                  \
	do                                                                                                 \
	{                                                                                                  \
		ncclResult_t nc_res_ = (call);                                                                 \
		if(nc_res_ != ncclSuccess)                                                                     \
		{                                                                                              \
			return fail(ctx, NB200_ERR_NCCL, "%s:%d %s: %s", __FILE__, __LINE__, #call,                \
						ctx->nccl->GetErrorString(nc_res_));                                           \
		}                                                                                              \
	} while(0)

#define LAUNCHED(ctx)                                                                                  \
	do                                                                                                 \
	{                                                                                                  \
		++(ctx)->launches;                                                                             \
		CU(ctx, cudaGetLastError());                                                                   \
	} while(0)

bool valid(const nb200_ctx* ctx, const nb200_buf* b)
{
	return ctx != nullptr && b != nullptr && ctx->live.count(b) != 0;
}

// Same logical size AND same per-lane layout (a state-sized buffer created before nb200_set_bodies is replicated,
// one created after it is body-sharded: mixing the two would run past the shorter shard)
bool same_shape(const nb200_buf* a, const nb200_buf* b)
{
	return a->bytes == b->bytes && a->lane_elems == b->lane_elems && a->sharded == b->sharded;
}

unsigned ew_grid(const nb200_lane& lane, size_t count)
{
	size_t	nvec = (count + NB200_VEC - 1) / NB200_VEC;
	size_t	blocks = (nvec + NB200_EW_THREADS - 1) / NB200_EW_THREADS;
	size_t	cap = static_cast<size_t>(lane.sm_count) * 8;	// 8 x 256 threads = full occupancy
	return static_cast<unsigned>(std::max<size_t>(1, std::min(blocks, cap)));
}

real* lane_ptr(const nb200_buf* b, size_t lane)
{
	return static_cast<real*>(b->dptr[lane]);
}

void free_lane(nb200_lane& l)
{
	cudaSetDevice(l.dev);
	bh_free(l.bh);
	l.bh = nullptr;
	if(l.mass) { cudaFree(l.mass); }
	if(l.src) { cudaFree(l.src); }
	if(l.partial) { cudaFree(l.partial); }
	if(l.sym_tiles) { cudaFree(l.sym_tiles); }
	if(l.sym_prow) { cudaFree(l.sym_prow); }
	if(l.sym_pcol) { cudaFree(l.sym_pcol); }
	if(l.sym_acc) { cudaFree(l.sym_acc); }
	l.sym_tiles = nullptr;
	l.sym_prow = l.sym_pcol = l.sym_acc = nullptr;
	l.sym_ntiles = 0;
	l.sym_edge = 0;
	l.partial_elems = 0;
	if(l.d_scalar) { cudaFree(l.d_scalar); }
	if(l.read_scratch) { cudaFree(l.read_scratch); }
	l.read_scratch = nullptr;
	l.read_scratch_bytes = 0;
	if(l.h_scalar) { cudaFreeHost(l.h_scalar); }
	l.mass = nullptr;
	l.src = nullptr;
	l.partial = nullptr;
	l.d_scalar = nullptr;
	l.h_scalar = nullptr;
}

// All-gather of one equally sized block per shard inside a per-lane buffer laid out [shard][block]:
// which == 0: packed source bodies (l.src, 4 reals per body); which == 1: leaf-ordered accelerations of the
// Barnes-Hut walk (l.bh->acc_all, 3 reals per body).
// lanes > 1: peer copies between this process's lanes; nranks > 1: NCCL all-gather.
char* gather_base(nb200_lane& l, int which)
{
	return which == 0 ? reinterpret_cast<char*>(l.src) : reinterpret_cast<char*>(l.bh->acc_all);
}

int gather_blocks(nb200_ctx* ctx, int which)
{
	const size_t	reals_per_body = which == 0 ? 4 : 3;
	const size_t	shard_bytes = ctx->n_shard * reals_per_body * sizeof(real);
	if(ctx->nranks > 1)
	{
		nb200_lane&	l = ctx->lanes[0];
		char*		base = gather_base(l, which);
		NC(ctx, ctx->nccl->AllGather(base + static_cast<size_t>(l.shard) * shard_bytes, base, ctx->n_shard * reals_per_body,
									 NB200_NCCL_REAL, static_cast<ncclComm_t>(ctx->comm), l.stream));
		return NB200_OK;
	}
	if(ctx->lanes.size() > 1 && ctx->lanes_nccl)
	{
		// one process, one communicator per lane (ncclCommInitAll): the reference's use_nccl=1 (nbody_engine_cuda.cpp:626-751)
		NC(ctx, ctx->nccl->GroupStart());
		for(auto& l : ctx->lanes)
		{
			char* base = gather_base(l, which);
			ncclResult_t r = ctx->nccl->AllGather(base + static_cast<size_t>(l.shard) * shard_bytes, base, ctx->n_shard * reals_per_body,
												  NB200_NCCL_REAL, static_cast<ncclComm_t>(l.lane_comm), l.stream);
			if(r != ncclSuccess)
			{
				ctx->nccl->GroupEnd();
				return fail(ctx, NB200_ERR_NCCL, "ncclAllGather (lane %d): %s", l.shard, ctx->nccl->GetErrorString(r));
			}
		}
		NC(ctx, ctx->nccl->GroupEnd());
		return NB200_OK;
	}
	if(ctx->lanes.size() > 1)
	{
		for(auto& l : ctx->lanes)
		{
			CU(ctx, cudaSetDevice(l.dev));
			CU(ctx, cudaEventRecord(l.ev_packed, l.stream));
		}
		for(auto& l : ctx->lanes)
		{
			CU(ctx, cudaSetDevice(l.dev));
			for(auto& p : ctx->lanes)
			{
				if(&p == &l) { continue; }
				CU(ctx, cudaStreamWaitEvent(l.stream, p.ev_packed, 0));
				size_t off = static_cast<size_t>(p.shard) * shard_bytes;
				CU(ctx, cudaMemcpyPeerAsync(gather_base(l, which) + off, l.dev, gather_base(p, which) + off, p.dev, shard_bytes, l.stream));
			}
			CU(ctx, cudaEventRecord(l.ev_gathered, l.stream));
		}
		// A lane may not repack (next fcompute) before every peer has read its shard.
		for(auto& l : ctx->lanes)
		{
			CU(ctx, cudaSetDevice(l.dev));
			for(auto& p : ctx->lanes)
			{
				if(&p != &l) { CU(ctx, cudaStreamWaitEvent(l.stream, p.ev_gathered, 0)); }
			}
		}
	}
	return NB200_OK;
}

int check_state_pair(nb200_ctx* ctx, const nb200_buf* y, const nb200_buf* f, const char* who)
{
	if(ctx == nullptr) { return NB200_ERR_ARG; }
	if(!valid(ctx, y)) { return fail(ctx, NB200_ERR_ARG, "%s: y is not a buffer of this context", who); }
	if(!valid(ctx, f)) { return fail(ctx, NB200_ERR_ARG, "%s: f is not a buffer of this context", who); }
	if(ctx->n == 0) { return fail(ctx, NB200_ERR_STATE, "%s: nb200_set_bodies has not been called", who); }
	if(!y->sharded || !f->sharded)
	{
		return fail(ctx, NB200_ERR_ARG, "%s: y and f must be state vectors of 6*N elements", who);
	}
	return NB200_OK;
}

int pack_and_gather(nb200_ctx* ctx, const nb200_buf* y)
{
	for(size_t li = 0; li < ctx->lanes.size(); ++li)
	{
		nb200_lane& l = ctx->lanes[li];
		CU(ctx, cudaSetDevice(l.dev));
		if(ctx->opt_timing) { CU(ctx, cudaEventRecord(l.ev_t[0], l.stream)); }
		unsigned grid = static_cast<unsigned>((ctx->n_shard + 255) / 256);
		direct_pack<<<grid, 256, 0, l.stream>>>(lane_ptr(y, li), l.mass, l.src, ctx->n_shard,
												static_cast<size_t>(l.shard) * ctx->n_shard);
		LAUNCHED(ctx);
	}
	int rc = gather_blocks(ctx, 0);
	if(rc != NB200_OK) { return rc; }
	if(ctx->opt_timing)
	{
		for(auto& l : ctx->lanes)
		{
			CU(ctx, cudaSetDevice(l.dev));
			CU(ctx, cudaEventRecord(l.ev_t[1], l.stream));
		}
	}
	return NB200_OK;
}

template<int IPT>
void launch_pairs(nb200_ctx* ctx, nb200_lane& l, const real* y, real* out, dim3 grid, int n_tiles, int tiles_per_seg, int write_f)
{
	direct_pairs<IPT><<<grid, NB200_DIRECT_THREADS, 0, l.stream>>>(
		l.src, y, out, ctx->n_shard, static_cast<size_t>(l.shard) * ctx->n_shard, n_tiles, tiles_per_seg, write_f);
}

// ---- solver steps as CUDA graphs: the step table (nb200_stepgraph.cuh) and its logic ------------------------------------
#include "nb200_stepgraph_impl.cuh"

#define STEP_NOTE(ctx, op)                                                                             \
	do                                                                                                 \
	{                                                                                                  \
		bool step_skip_ = false;                                                                       \
		int step_rc_ = step_note(ctx, op, &step_skip_);                                                \
		if(step_rc_ != NB200_OK || step_skip_) { return step_rc_; }                                    \
	} while(0)

step_op make_op(int kind, const nb200_buf* a, const nb200_buf* b = nullptr, const nb200_buf* c = nullptr)
{
	step_op op;
	op.kind = kind;
	op.a = a;
	op.b = b;
	op.c = c;
	return op;
}

}  // namespace

// ---- library / device queries ------------------------------------------------
NB200_API int nb200_real_size(void)
{
	return static_cast<int>(sizeof(real));
}

NB200_API int nb200_device_count(int* count)
{
	if(count == nullptr) { return NB200_ERR_ARG; }
	*count = 0;
	cudaError_t res = cudaGetDeviceCount(count);
	if(res != cudaSuccess)
	{
		*count = 0;
		cudaGetLastError();
		return NB200_ERR_CUDA;
	}
	return NB200_OK;
}

NB200_API int nb200_comm_unique_id(void* uid128)
{
	if(uid128 == nullptr) { return NB200_ERR_ARG; }
	std::string	err;
	nccl_api*	api = nccl_load(err);
	if(api == nullptr)
	{
		fprintf(stderr, "nb200: %s\n", err.c_str());
		return NB200_ERR_NCCL;
	}
	static_assert(sizeof(ncclUniqueId) == NB200_UID_BYTES, "NCCL unique id size");
	ncclUniqueId id;
	if(api->GetUniqueId(&id) != ncclSuccess) { return NB200_ERR_NCCL; }
	memcpy(uid128, &id, sizeof(id));
	return NB200_OK;
}

// ---- context -------------------------------------------------------------------
NB200_API int nb200_create(nb200_ctx** out, const int* dev_ids, int nlanes, int rank, int nranks, const void* uid128)
{
	if(out == nullptr) { return NB200_ERR_ARG; }
	*out = nullptr;
	if(dev_ids == nullptr || nlanes < 1 || nranks < 1 || rank < 0 || rank >= nranks) { return NB200_ERR_ARG; }
	if(nlanes > 1 && nranks > 1)
	{
		fprintf(stderr, "nb200_create: use either several lanes in one process or one lane per rank\n");
		return NB200_ERR_UNSUPPORTED;
	}
	if(nranks > 1 && uid128 == nullptr) { return NB200_ERR_ARG; }
	int	count = 0;
	if(cudaGetDeviceCount(&count) != cudaSuccess || count < 1)
	{
		cudaGetLastError();
		fprintf(stderr, "nb200_create: no CUDA device (there is no CPU fallback)\n");
		return NB200_ERR_CUDA;
	}
	for(int i = 0; i < nlanes; ++i)
	{
		if(dev_ids[i] < 0 || dev_ids[i] >= count)
		{
			fprintf(stderr, "nb200_create: invalid device id %d, must be in [0, %d)\n", dev_ids[i], count);
			return NB200_ERR_ARG;
		}
	}
	nb200_ctx* ctx = new nb200_ctx();
	ctx->sg = new step_graph();
	ctx->rank = rank;
	ctx->nranks = nranks;
	ctx->nshards = nlanes * nranks;
	ctx->first_shard = (nranks > 1) ? rank : 0;
	ctx->lanes.resize(static_cast<size_t>(nlanes));
	int status = NB200_OK;
	for(int i = 0; i < nlanes && status == NB200_OK; ++i)
	{
		nb200_lane& l = ctx->lanes[static_cast<size_t>(i)];
		l.dev = dev_ids[i];
		l.shard = ctx->first_shard + i;
		cudaDeviceProp prop;
		bool ok = cudaSetDevice(l.dev) == cudaSuccess &&
				  cudaGetDeviceProperties(&prop, l.dev) == cudaSuccess &&
				  cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking) == cudaSuccess &&
				  cudaEventCreateWithFlags(&l.ev_packed, cudaEventDisableTiming) == cudaSuccess &&
				  cudaEventCreateWithFlags(&l.ev_gathered, cudaEventDisableTiming) == cudaSuccess &&
				  cudaMalloc(&l.d_scalar, 32 * sizeof(unsigned long long)) == cudaSuccess &&
				  cudaMallocHost(&l.h_scalar, 32 * sizeof(unsigned long long)) == cudaSuccess;
		for(int e = 0; e < 5 && ok; ++e) { ok = cudaEventCreate(&l.ev_t[e]) == cudaSuccess; }
		for(int e = 0; e < 8 && ok; ++e) { ok = cudaEventCreate(&l.ev_mark[e]) == cudaSuccess; }
		if(!ok)
		{
			fprintf(stderr, "nb200_create: device %d setup failed: %s\n", l.dev, cudaGetErrorString(cudaGetLastError()));
			status = NB200_ERR_CUDA;
			break;
		}
		l.sm_count = prop.multiProcessorCount;
		if(prop.major < 10)
		{
			fprintf(stderr, "nb200_create: device %d is sm_%d%d; this library is built for sm_100a only\n", l.dev, prop.major, prop.minor);
			status = NB200_ERR_UNSUPPORTED;
		}
	}
	if(status == NB200_OK && nlanes > 1)
	{
		// Peer access between distinct devices of the lane list (NVLink / NVSwitch).
		for(auto& a : ctx->lanes)
		{
			for(auto& b : ctx->lanes)
			{
				if(a.dev == b.dev) { continue; }
				int can = 0;
				cudaDeviceCanAccessPeer(&can, a.dev, b.dev);
				if(can)
				{
					cudaSetDevice(a.dev);
					cudaError_t r = cudaDeviceEnablePeerAccess(b.dev, 0);
					if(r != cudaSuccess && r != cudaErrorPeerAccessAlreadyEnabled) { status = NB200_ERR_CUDA; }
					cudaGetLastError();
				}
				else
				{
					ctx->peer_loads = false;	// gathers go through cudaMemcpyPeer staging; no symmetric-tile path
				}
			}
		}
	}
	if(status == NB200_OK && nranks > 1)
	{
		ctx->nccl = nccl_load(ctx->err);
		if(ctx->nccl == nullptr)
		{
			fprintf(stderr, "nb200_create: %s\n", ctx->err.c_str());
			status = NB200_ERR_NCCL;
		}
		else
		{
			ncclUniqueId id;
			memcpy(&id, uid128, sizeof(id));
			ncclComm_t comm = nullptr;
			cudaSetDevice(ctx->lanes[0].dev);
			ncclResult_t r = ctx->nccl->CommInitRank(&comm, nranks, id, rank);
			if(r != ncclSuccess)
			{
				fprintf(stderr, "nb200_create: ncclCommInitRank: %s\n", ctx->nccl->GetErrorString(r));
				status = NB200_ERR_NCCL;
			}
			ctx->comm = comm;
		}
	}
	if(status != NB200_OK)
	{
		nb200_destroy(ctx);
		return status;
	}
	*out = ctx;
	return NB200_OK;
}

NB200_API int nb200_destroy(nb200_ctx* ctx)
{
	if(ctx == nullptr) { return NB200_OK; }
	if(ctx->sg != nullptr && !ctx->lanes.empty() && ctx->lanes[0].stream != nullptr)
	{
		step_invalidate(ctx);	// issues whatever a replayed step still holds back
		step_clear_table(ctx);
		step_destroy_execs(ctx, ctx->sg->cap_execs);
		ctx->sg->mode = SG_OFF;
	}
	for(auto& l : ctx->lanes)
	{
		cudaSetDevice(l.dev);
		if(l.stream) { cudaStreamSynchronize(l.stream); }
	}
	if(ctx->comm != nullptr && ctx->nccl != nullptr)
	{
		ctx->nccl->CommDestroy(static_cast<ncclComm_t>(ctx->comm));
	}
	for(auto& l : ctx->lanes)
	{
		if(l.lane_comm != nullptr && ctx->nccl != nullptr) { ctx->nccl->CommDestroy(static_cast<ncclComm_t>(l.lane_comm)); }
		l.lane_comm = nullptr;
	}
	// Buffers the caller never released
	std::vector<const nb200_buf*> leaked(ctx->live.begin(), ctx->live.end());
	for(const nb200_buf* b : leaked) { nb200_free(ctx, const_cast<nb200_buf*>(b)); }
	for(auto& l : ctx->lanes)
	{
		free_lane(l);
		cudaSetDevice(l.dev);
		if(l.ev_packed) { cudaEventDestroy(l.ev_packed); }
		if(l.ev_gathered) { cudaEventDestroy(l.ev_gathered); }
		for(auto& e : l.ev_t) { if(e) { cudaEventDestroy(e); } }
		for(auto& e : l.ev_mark) { if(e) { cudaEventDestroy(e); } }
		if(l.stream) { cudaStreamDestroy(l.stream); }
	}
	cudaGetLastError();
	delete ctx->sg;
	delete ctx;
	return NB200_OK;
}

NB200_API const char* nb200_last_error(const nb200_ctx* ctx)
{
	return ctx != nullptr ? ctx->err.c_str() : "null context";
}

NB200_API int nb200_sync(nb200_ctx* ctx)
{
	if(ctx == nullptr) { return NB200_ERR_ARG; }
	step_break(ctx);
	for(auto& l : ctx->lanes)
	{
		CU(ctx, cudaSetDevice(l.dev));
		CU(ctx, cudaStreamSynchronize(l.stream));
	}
	return NB200_OK;
}

NB200_API int nb200_shards(const nb200_ctx* ctx, int* nshards, int* first_shard)
{
	if(ctx == nullptr) { return NB200_ERR_ARG; }
	if(nshards) { *nshards = ctx->nshards; }
	if(first_shard) { *first_shard = ctx->first_shard; }
	return NB200_OK;
}

NB200_API int nb200_describe(const nb200_ctx* ctx, char* text, size_t text_bytes)
{
	if(ctx == nullptr || text == nullptr || text_bytes == 0) { return NB200_ERR_ARG; }
	std::string s;
	char line[640];
	for(size_t i = 0; i < ctx->lanes.size(); ++i)
	{
		cudaDeviceProp prop;
		if(cudaGetDeviceProperties(&prop, ctx->lanes[i].dev) != cudaSuccess) { continue; }
		snprintf(line, sizeof(line), "\t #%zu ID %d %s sm_%d%d SMs %d HBM %zu MB L2 %d KB shard %d/%d\n", i, ctx->lanes[i].dev,
				 prop.name, prop.major, prop.minor, prop.multiProcessorCount, prop.totalGlobalMem >> 20, prop.l2CacheSize >> 10,
				 ctx->lanes[i].shard, ctx->nshards);
		s += line;
	}
	snprintf(line, sizeof(line), "\t precision %s, ranks %d, NCCL %s\n", sizeof(real) == 8 ? "FP64" : "FP32", ctx->nranks,
			 ctx->comm ? "on (one process per GPU)" : (ctx->lanes_nccl ? "on (one communicator per lane)" : "off"));
	s += line;
	snprintf(line, sizeof(line), "\t shard exchange: %s; solver steps as CUDA graphs: %s\n",
			 ctx->nshards == 1 ? "none (one shard)" : (ctx->comm || ctx->lanes_nccl ? "NCCL" : (ctx->peer_loads ? "peer copies and peer loads (NVLink)" : "peer copies")),
			 ctx->sg != nullptr && ctx->sg->mode != SG_OFF ? "on" : (ctx->nshards > 1 ? "off (one shard only)" : "off"));
	s += line;
	snprintf(text, text_bytes, "%s", s.c_str());
	return NB200_OK;
}

NB200_API int nb200_set_bodies(nb200_ctx* ctx, size_t n, const nb200_real* mass)
{
	if(ctx == nullptr) { return NB200_ERR_ARG; }
	if(n == 0 || mass == nullptr) { return fail(ctx, NB200_ERR_ARG, "set_bodies: empty body set"); }
	step_invalidate(ctx);
	if(n % static_cast<size_t>(ctx->nshards) != 0)
	{
		return fail(ctx, NB200_ERR_ARG, "set_bodies: N = %zu is not a multiple of the shard count %d", n, ctx->nshards);
	}
	if(ctx->n != 0 && ctx->n != n)
	{
		for(const nb200_buf* b : ctx->live)
		{
			if(b->sharded) { return fail(ctx, NB200_ERR_STATE, "set_bodies: body count changed while state vectors of the old size are alive"); }
		}
	}
	ctx->n = n;
	ctx->sym_unavailable = false;
	ctx->n_shard = n / static_cast<size_t>(ctx->nshards);
	// A state-sized buffer created BEFORE the body count was known (the adapter allows create_buffer before init) was
	// laid out as a replicated buffer. One shard: that is the layout of a state vector anyway. Several shards: every lane
	// holds the whole vector and keeps only its own columns from here on (contents preserved).
	for(const nb200_buf* cb : ctx->live)
	{
		nb200_buf* b = const_cast<nb200_buf*>(cb);
		if(b->bytes != 6 * n * sizeof(real) || b->sharded) { continue; }
		if(ctx->nshards > 1)
		{
			const size_t n_shard = n / static_cast<size_t>(ctx->nshards);
			for(size_t i = 0; i < ctx->lanes.size(); ++i)
			{
				nb200_lane& l = ctx->lanes[i];
				CU(ctx, cudaSetDevice(l.dev));
				void* shard = nullptr;
				if(cudaMalloc(&shard, 6 * n_shard * sizeof(real)) != cudaSuccess)
				{
					cudaGetLastError();
					return fail(ctx, NB200_ERR_ALLOC, "set_bodies: re-sharding an early state vector failed");
				}
				const char* from = static_cast<const char*>(b->dptr[i]) + static_cast<size_t>(l.shard) * n_shard * sizeof(real);
				CU(ctx, cudaMemcpy2DAsync(shard, n_shard * sizeof(real), from, n * sizeof(real), n_shard * sizeof(real), 6,
										  cudaMemcpyDeviceToDevice, l.stream));
				CU(ctx, cudaStreamSynchronize(l.stream));
				cudaFree(b->dptr[i]);
				b->dptr[i] = shard;
			}
			b->lane_elems = 6 * n_shard;
			b->lane_bytes = b->lane_elems * sizeof(real);
		}
		b->sharded = true;
	}
	ctx->n_pad = (n + NB200_DIRECT_TILE - 1) / NB200_DIRECT_TILE * NB200_DIRECT_TILE;
	ctx->n_alloc = (n + 8191) / 8192 * 8192 + 8192;	// room for the zero-mass padding of a last tile of any edge <= 8192
	for(auto& l : ctx->lanes)
	{
		CU(ctx, cudaSetDevice(l.dev));
		if(l.mass) { cudaFree(l.mass); l.mass = nullptr; }
		if(l.src) { cudaFree(l.src); l.src = nullptr; }
		if(l.sym_tiles) { cudaFree(l.sym_tiles); l.sym_tiles = nullptr; }
		if(l.sym_prow) { cudaFree(l.sym_prow); l.sym_prow = nullptr; }
		if(l.sym_pcol) { cudaFree(l.sym_pcol); l.sym_pcol = nullptr; }
		if(l.sym_acc) { cudaFree(l.sym_acc); l.sym_acc = nullptr; }
		l.sym_edge = 0;
		l.sym_ntiles = 0;
		bh_free(l.bh);
		l.bh = nullptr;
		if(cudaMalloc(&l.mass, n * sizeof(real)) != cudaSuccess || cudaMalloc(&l.src, ctx->n_alloc * sizeof(body4)) != cudaSuccess)
		{
			cudaGetLastError();
			return fail(ctx, NB200_ERR_ALLOC, "set_bodies: device allocation failed");
		}
		// Padding bodies have zero mass: they contribute exactly 0 to every sum.
		CU(ctx, cudaMemsetAsync(l.src, 0, ctx->n_alloc * sizeof(body4), l.stream));
		CU(ctx, cudaMemcpyAsync(l.mass, mass, n * sizeof(real), cudaMemcpyHostToDevice, l.stream));
		CU(ctx, cudaStreamSynchronize(l.stream));
	}
	return NB200_OK;
}

NB200_API int nb200_get_mass(nb200_ctx* ctx, nb200_real* mass)
{
	if(ctx == nullptr || mass == nullptr) { return NB200_ERR_ARG; }
	if(ctx->n == 0) { return fail(ctx, NB200_ERR_STATE, "get_mass: no bodies"); }
	step_break(ctx);
	nb200_lane& l = ctx->lanes[0];
	CU(ctx, cudaSetDevice(l.dev));
	CU(ctx, cudaMemcpyAsync(mass, l.mass, ctx->n * sizeof(real), cudaMemcpyDeviceToHost, l.stream));
	CU(ctx, cudaStreamSynchronize(l.stream));
	return NB200_OK;
}

// ---- buffers ---------------------------------------------------------------------
NB200_API int nb200_alloc(nb200_ctx* ctx, size_t bytes, nb200_buf** out)
{
	if(ctx == nullptr || out == nullptr) { return NB200_ERR_ARG; }
	*out = nullptr;
	step_invalidate(ctx);	// a new handle may reuse the address of one a recorded step refers to
	nb200_buf* b = new nb200_buf();
	b->owner = ctx;
	b->bytes = bytes;
	b->sharded = (ctx->n != 0 && bytes == 6 * ctx->n * sizeof(real));
	b->lane_elems = b->sharded ? 6 * ctx->n_shard : bytes / sizeof(real);
	b->lane_bytes = b->sharded ? b->lane_elems * sizeof(real) : bytes;
	b->dptr.assign(ctx->lanes.size(), nullptr);
	for(size_t i = 0; i < ctx->lanes.size(); ++i)
	{
		cudaSetDevice(ctx->lanes[i].dev);
		if(cudaMalloc(&b->dptr[i], std::max<size_t>(b->lane_bytes, 16)) != cudaSuccess)
		{
			cudaGetLastError();
			for(size_t k = 0; k < i; ++k)
			{
				cudaSetDevice(ctx->lanes[k].dev);
				cudaFree(b->dptr[k]);
			}
			delete b;
			return fail(ctx, NB200_ERR_ALLOC, "alloc: cudaMalloc of %zu bytes failed", bytes);
		}
	}
	ctx->live.insert(b);
	*out = b;
	return NB200_OK;
}

NB200_API int nb200_free(nb200_ctx* ctx, nb200_buf* b)
{
	if(b == nullptr) { return NB200_OK; }
	if(!valid(ctx, b)) { return fail(ctx, NB200_ERR_ARG, "free: not a buffer of this context"); }
	step_invalidate(ctx);
	for(size_t i = 0; i < ctx->lanes.size(); ++i)
	{
		cudaSetDevice(ctx->lanes[i].dev);
		// cudaFree synchronises the device, so work still using the buffer has finished.
		cudaFree(b->dptr[i]);
	}
	cudaGetLastError();
	ctx->live.erase(b);
	b->magic = 0;
	delete b;
	return NB200_OK;
}

NB200_API size_t nb200_size(const nb200_buf* b)
{
	return (b != nullptr && b->magic == NB200_BUF_MAGIC) ? b->bytes : 0;
}

NB200_API int nb200_lane_ptr(nb200_ctx* ctx, nb200_buf* b, int lane, void** dptr, size_t* elems)
{
	if(!valid(ctx, b) || lane < 0 || static_cast<size_t>(lane) >= ctx->lanes.size()) { return NB200_ERR_ARG; }
	step_invalidate(ctx);	// the caller may touch the memory behind the library's back: nothing may stay deferred
	if(dptr) { *dptr = b->dptr[static_cast<size_t>(lane)]; }
	if(elems) { *elems = b->lane_elems; }
	return NB200_OK;
}

NB200_API int nb200_write(nb200_ctx* ctx, nb200_buf* dst, const void* host)
{
	if(ctx == nullptr) { return NB200_ERR_ARG; }
	if(!valid(ctx, dst)) { return fail(ctx, NB200_ERR_ARG, "write: dst is not a buffer of this context"); }
	if(host == nullptr) { return fail(ctx, NB200_ERR_ARG, "write: NULL source"); }
	step_break(ctx);
	for(size_t i = 0; i < ctx->lanes.size(); ++i)
	{
		nb200_lane& l = ctx->lanes[i];
		CU(ctx, cudaSetDevice(l.dev));
		if(dst->sharded)
		{
			// 6 rows of N -> 6 rows of n_shard (columns of this shard)
			const char* src = static_cast<const char*>(host) + static_cast<size_t>(l.shard) * ctx->n_shard * sizeof(real);
			CU(ctx, cudaMemcpy2DAsync(dst->dptr[i], ctx->n_shard * sizeof(real), src, ctx->n * sizeof(real),
									  ctx->n_shard * sizeof(real), 6, cudaMemcpyHostToDevice, l.stream));
		}
		else if(dst->bytes != 0)
		{
			CU(ctx, cudaMemcpyAsync(dst->dptr[i], host, dst->bytes, cudaMemcpyHostToDevice, l.stream));
		}
	}
	return nb200_sync(ctx);	// host memory may be reused as soon as we return
}

NB200_API int nb200_read(nb200_ctx* ctx, void* host, const nb200_buf* src)
{
	if(ctx == nullptr) { return NB200_ERR_ARG; }
	if(!valid(ctx, src)) { return fail(ctx, NB200_ERR_ARG, "read: src is not a buffer of this context"); }
	if(host == nullptr) { return fail(ctx, NB200_ERR_ARG, "read: NULL destination"); }
	step_break(ctx);
	if(!src->sharded)
	{
		nb200_lane& l = ctx->lanes[0];
		CU(ctx, cudaSetDevice(l.dev));
		if(src->bytes != 0)
		{
			CU(ctx, cudaMemcpyAsync(host, src->dptr[0], src->bytes, cudaMemcpyDeviceToHost, l.stream));
		}
		CU(ctx, cudaStreamSynchronize(l.stream));
		return NB200_OK;
	}
	if(ctx->nranks > 1)
	{
		// All ranks call read at the same point: gather the shards, then un-interleave on the host.
		nb200_lane&	l = ctx->lanes[0];
		CU(ctx, cudaSetDevice(l.dev));
		if(l.read_scratch_bytes < src->bytes)
		{
			if(l.read_scratch) { cudaFree(l.read_scratch); l.read_scratch = nullptr; l.read_scratch_bytes = 0; }
			if(cudaMalloc(&l.read_scratch, src->bytes) != cudaSuccess)
			{
				cudaGetLastError();
				return fail(ctx, NB200_ERR_ALLOC, "read: gather scratch allocation failed");
			}
			l.read_scratch_bytes = src->bytes;
		}
		real* all = l.read_scratch;
		NC(ctx, ctx->nccl->AllGather(src->dptr[0], all, src->lane_elems, NB200_NCCL_REAL,
									 static_cast<ncclComm_t>(ctx->comm), l.stream));
		for(int g = 0; g < ctx->nshards; ++g)
		{
			char* dst = static_cast<char*>(host) + static_cast<size_t>(g) * ctx->n_shard * sizeof(real);
			CU(ctx, cudaMemcpy2DAsync(dst, ctx->n * sizeof(real), all + static_cast<size_t>(g) * src->lane_elems,
									  ctx->n_shard * sizeof(real), ctx->n_shard * sizeof(real), 6, cudaMemcpyDeviceToHost, l.stream));
		}
		CU(ctx, cudaStreamSynchronize(l.stream));
		return NB200_OK;
	}
	return nb200_read_local(ctx, host, src);
}

NB200_API int nb200_read_local(nb200_ctx* ctx, void* host, const nb200_buf* src)
{
	if(ctx == nullptr) { return NB200_ERR_ARG; }
	if(!valid(ctx, src)) { return fail(ctx, NB200_ERR_ARG, "read: src is not a buffer of this context"); }
	if(host == nullptr) { return fail(ctx, NB200_ERR_ARG, "read: NULL destination"); }
	step_break(ctx);
	if(!src->sharded)
	{
		nb200_lane& l = ctx->lanes[0];
		CU(ctx, cudaSetDevice(l.dev));
		if(src->bytes != 0)
		{
			CU(ctx, cudaMemcpyAsync(host, src->dptr[0], src->bytes, cudaMemcpyDeviceToHost, l.stream));
		}
		CU(ctx, cudaStreamSynchronize(l.stream));
		return NB200_OK;
	}
	for(size_t i = 0; i < ctx->lanes.size(); ++i)
	{
		nb200_lane& l = ctx->lanes[i];
		CU(ctx, cudaSetDevice(l.dev));
		char* dst = static_cast<char*>(host) + static_cast<size_t>(l.shard) * ctx->n_shard * sizeof(real);
		CU(ctx, cudaMemcpy2DAsync(dst, ctx->n * sizeof(real), src->dptr[i], ctx->n_shard * sizeof(real),
								  ctx->n_shard * sizeof(real), 6, cudaMemcpyDeviceToHost, l.stream));
	}
	return nb200_sync(ctx);
}

namespace {
int ensure_read_scratch(nb200_ctx* ctx, nb200_lane& l, size_t bytes)
{
	if(l.read_scratch_bytes >= bytes) { return NB200_OK; }
	if(l.read_scratch) { cudaFree(l.read_scratch); l.read_scratch = nullptr; l.read_scratch_bytes = 0; }
	if(cudaMalloc(&l.read_scratch, bytes) != cudaSuccess)
	{
		cudaGetLastError();
		return fail(ctx, NB200_ERR_ALLOC, "staging allocation of %zu bytes failed", bytes);
	}
	l.read_scratch_bytes = bytes;
	return NB200_OK;
}
}  // namespace

// ---- body arrays <-> state vector (the get_data / init side of the engine) --------------------------
NB200_API int nb200_host_register(nb200_ctx* ctx, void* host, size_t bytes)
{
	if(ctx == nullptr) { return NB200_ERR_ARG; }
	if(host == nullptr || bytes == 0) { return fail(ctx, NB200_ERR_ARG, "host_register: empty range"); }
	cudaError_t res = cudaHostRegister(host, bytes, cudaHostRegisterPortable);
	if(res != cudaSuccess && res != cudaErrorHostMemoryAlreadyRegistered)
	{
		cudaGetLastError();
		return fail(ctx, NB200_ERR_CUDA, "host_register: %s", cudaGetErrorString(res));
	}
	cudaGetLastError();
	return NB200_OK;
}

NB200_API int nb200_host_unregister(nb200_ctx* ctx, void* host)
{
	if(ctx == nullptr || host == nullptr) { return NB200_ERR_ARG; }
	cudaHostUnregister(host);
	cudaGetLastError();	// not registered: nothing to undo
	return NB200_OK;
}

NB200_API int nb200_write_bodies(nb200_ctx* ctx, nb200_buf* y, const nb200_real* pos_xyz, const nb200_real* vel_xyz)
{
	if(ctx == nullptr) { return NB200_ERR_ARG; }
	if(!valid(ctx, y)) { return fail(ctx, NB200_ERR_ARG, "write_bodies: y is not a buffer of this context"); }
	if(ctx->n == 0 || !y->sharded) { return fail(ctx, NB200_ERR_ARG, "write_bodies: y must be a state vector of 6*N elements"); }
	if(pos_xyz == nullptr || vel_xyz == nullptr) { return fail(ctx, NB200_ERR_ARG, "write_bodies: NULL source"); }
	step_break(ctx);
	const size_t	n3 = 3 * ctx->n_shard;
	for(size_t i = 0; i < ctx->lanes.size(); ++i)
	{
		nb200_lane& l = ctx->lanes[i];
		CU(ctx, cudaSetDevice(l.dev));
		int rc = ensure_read_scratch(ctx, l, 2 * n3 * sizeof(real));
		if(rc != NB200_OK) { return rc; }
		const size_t first = static_cast<size_t>(l.shard) * n3;	// a shard is a contiguous range of bodies
		CU(ctx, cudaMemcpyAsync(l.read_scratch, pos_xyz + first, n3 * sizeof(real), cudaMemcpyHostToDevice, l.stream));
		CU(ctx, cudaMemcpyAsync(l.read_scratch + n3, vel_xyz + first, n3 * sizeof(real), cudaMemcpyHostToDevice, l.stream));
		ew_bodies_to_state<<<static_cast<unsigned>((2 * n3 + NB200_EW_THREADS - 1) / NB200_EW_THREADS), NB200_EW_THREADS, 0, l.stream>>>(
			l.read_scratch, lane_ptr(y, i), ctx->n_shard);
		LAUNCHED(ctx);
	}
	return nb200_sync(ctx);	// host memory may be reused as soon as we return
}

NB200_API int nb200_read_bodies(nb200_ctx* ctx, const nb200_buf* y, nb200_real* pos_xyz, nb200_real* vel_xyz)
{
	if(ctx == nullptr) { return NB200_ERR_ARG; }
	if(!valid(ctx, y)) { return fail(ctx, NB200_ERR_ARG, "read_bodies: y is not a buffer of this context"); }
	if(ctx->n == 0 || !y->sharded) { return fail(ctx, NB200_ERR_ARG, "read_bodies: y must be a state vector of 6*N elements"); }
	if(pos_xyz == nullptr || vel_xyz == nullptr) { return fail(ctx, NB200_ERR_ARG, "read_bodies: NULL destination"); }
	step_break(ctx);
	const size_t	n3 = 3 * ctx->n_shard;
	const size_t	blocks = ctx->nranks > 1 ? static_cast<size_t>(ctx->nshards) + 1 : 1;	// own block + the gathered ones
	for(size_t i = 0; i < ctx->lanes.size(); ++i)
	{
		nb200_lane& l = ctx->lanes[i];
		CU(ctx, cudaSetDevice(l.dev));
		int rc = ensure_read_scratch(ctx, l, blocks * 2 * n3 * sizeof(real));
		if(rc != NB200_OK) { return rc; }
		ew_state_to_bodies<<<static_cast<unsigned>((2 * n3 + NB200_EW_THREADS - 1) / NB200_EW_THREADS), NB200_EW_THREADS, 0, l.stream>>>(
			lane_ptr(y, i), l.read_scratch, ctx->n_shard);
		LAUNCHED(ctx);
		if(ctx->nranks > 1)
		{
			real* all = l.read_scratch + 2 * n3;
			NC(ctx, ctx->nccl->AllGather(l.read_scratch, all, 2 * n3, NB200_NCCL_REAL, static_cast<ncclComm_t>(ctx->comm), l.stream));
			for(int g = 0; g < ctx->nshards; ++g)
			{
				const real* blk = all + static_cast<size_t>(g) * 2 * n3;
				CU(ctx, cudaMemcpyAsync(pos_xyz + static_cast<size_t>(g) * n3, blk, n3 * sizeof(real), cudaMemcpyDeviceToHost, l.stream));
				CU(ctx, cudaMemcpyAsync(vel_xyz + static_cast<size_t>(g) * n3, blk + n3, n3 * sizeof(real), cudaMemcpyDeviceToHost, l.stream));
			}
		}
		else
		{
			const size_t first = static_cast<size_t>(l.shard) * n3;
			CU(ctx, cudaMemcpyAsync(pos_xyz + first, l.read_scratch, n3 * sizeof(real), cudaMemcpyDeviceToHost, l.stream));
			CU(ctx, cudaMemcpyAsync(vel_xyz + first, l.read_scratch + n3, n3 * sizeof(real), cudaMemcpyDeviceToHost, l.stream));
		}
	}
	return nb200_sync(ctx);
}

NB200_API int nb200_copy(nb200_ctx* ctx, nb200_buf* a, const nb200_buf* b)
{
	if(ctx == nullptr) { return NB200_ERR_ARG; }
	if(!valid(ctx, a)) { return fail(ctx, NB200_ERR_ARG, "copy: a is not a buffer of this context"); }
	if(!valid(ctx, b)) { return fail(ctx, NB200_ERR_ARG, "copy: b is not a buffer of this context"); }
	if(!same_shape(a, b)) { return fail(ctx, NB200_ERR_ARG, "copy: size does not match"); }
	if(a == b || a->lane_bytes == 0) { return NB200_OK; }
	STEP_NOTE(ctx, make_op(SOP_COPY, a, b));
	for(size_t i = 0; i < ctx->lanes.size(); ++i)
	{
		nb200_lane& l = ctx->lanes[i];
		CU(ctx, cudaSetDevice(l.dev));
		CU(ctx, cudaMemcpyAsync(a->dptr[i], b->dptr[i], a->lane_bytes, cudaMemcpyDeviceToDevice, l.stream));
	}
	return NB200_OK;
}

namespace {
int fill_lanes(nb200_ctx* ctx, nb200_buf* a, real value);
}

NB200_API int nb200_fill(nb200_ctx* ctx, nb200_buf* a, nb200_real value)
{
	if(ctx == nullptr) { return NB200_ERR_ARG; }
	if(!valid(ctx, a)) { return fail(ctx, NB200_ERR_ARG, "fill: a is not a buffer of this context"); }
	if(a->lane_elems == 0) { return NB200_OK; }
	step_op op = make_op(SOP_FILL, a);
	op.coef.push_back(value);
	STEP_NOTE(ctx, op);
	return fill_lanes(ctx, a, value);
}

namespace {
int fill_lanes(nb200_ctx* ctx, nb200_buf* a, real value)
{
	for(size_t i = 0; i < ctx->lanes.size(); ++i)
	{
		nb200_lane& l = ctx->lanes[i];
		CU(ctx, cudaSetDevice(l.dev));
		ew_fill<<<ew_grid(l, a->lane_elems), NB200_EW_THREADS, 0, l.stream>>>(lane_ptr(a, i), value, a->lane_elems);
		LAUNCHED(ctx);
	}
	return NB200_OK;
}
}  // namespace

// ---- direct all-pairs --------------------------------------------------------------
namespace {
// Tile edge of the symmetric path for this problem, 0 = use the plain kernel.
int sym_tile_edge(const nb200_ctx* ctx)
{
	if(ctx->opt_direct_sym == 0 || ctx->sym_unavailable) { return 0; }
	if(ctx->lanes.size() > 1 && !ctx->lanes_nccl && (ctx->nranks > 1 || !ctx->peer_loads || ctx->lanes.size() > NB200_SYM_MAX_PEERS)) { return 0; }
	if(ctx->opt_direct_sym < 0 && ctx->n < NB200_SYM_MIN_BODIES) { return 0; }	// too few tiles to fill 148 SMs
	// tile partials are 24 N^2 / T bytes over all ranks: 51 GB at N = 4M, 80 GB at 5M (a B200 has 180 GB); if they do not fit
	// next to the caller's buffers every shard falls back together (sym_fcompute)
	if(ctx->opt_direct_sym < 0 && ctx->n > (static_cast<size_t>(5) << 20) * static_cast<size_t>(ctx->nranks)) { return 0; }
	// automatic edge: ~N/128 (>= 8000 equal tiles for N >= 32,768), at most 8192 (192 KB of column sums; scratch
	// 24 N^2 / T bytes), at least 256 (the default kernels then run 2-warp / 1-warp CTAs, several per SM)
	long long edge = ctx->opt_sym_tile;
	if(edge <= 0)
	{
		edge = 256;
		while(edge * 2 <= static_cast<long long>(ctx->n / 128) && edge < 8192) { edge *= 2; }
	}
	// a power of two in [256, 8192]
	if(edge > 8192 || edge < 256 || (edge & (edge - 1)) != 0) { return 0; }
	if(ctx->opt_sym_shape == NB200_SYM_DEFAULT_SHAPE)
	{
		return static_cast<int>(edge);	// the CTA has as many warps as the tile has row blocks (1, 2, 4 or 8)
	}
	// the other shapes always run 8 warps: 8 column blocks per phase round, 32*J bodies each
	const long long unit = ctx->opt_sym_shape == 3 ? 1024 : (ctx->opt_sym_shape >= 1 ? 512 : 256);
	if(edge % unit != 0) { return 0; }
	return static_cast<int>(edge);
}

template<class Kernel>
int sym_launch(nb200_ctx* ctx, nb200_lane& l, Kernel kernel, size_t tiles, size_t smem, int T, int warps = NB200_SYM_WARPS)
{
	CU(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
	kernel<<<static_cast<unsigned>(tiles), 32 * warps, smem, l.stream>>>(
		l.src, static_cast<const int2*>(l.sym_tiles), l.sym_prow, l.sym_pcol, T);
	LAUNCHED(ctx);
	return NB200_OK;
}

// The default kernel of this build, with as many warps per CTA as the tile has row blocks of 32 * I bodies
int sym_launch_default(nb200_ctx* ctx, nb200_lane& l, size_t tiles, size_t smem, int T)
{
#if NB200_PRECISION == 1
	const int row_blocks = T / 256;	// direct_sym_tiles_f32x2<8>
	if(row_blocks >= 8) { return sym_launch(ctx, l, direct_sym_tiles_f32x2<8, false, 8>, tiles, smem, T, 8); }
	if(row_blocks >= 4) { return sym_launch(ctx, l, direct_sym_tiles_f32x2<8, false, 4>, tiles, smem, T, 4); }
	if(row_blocks >= 2) { return sym_launch(ctx, l, direct_sym_tiles_f32x2<8, false, 2>, tiles, smem, T, 2); }
	return sym_launch(ctx, l, direct_sym_tiles_f32x2<8, false, 1>, tiles, smem, T, 1);
#else
	const int row_blocks = T / 256;	// direct_sym_tiles<8, 1>
	if(row_blocks >= 8) { return sym_launch(ctx, l, direct_sym_tiles<8, 1, true, 8>, tiles, smem, T, 8); }
	if(row_blocks >= 4) { return sym_launch(ctx, l, direct_sym_tiles<8, 1, true, 4>, tiles, smem, T, 4); }
	if(row_blocks >= 2) { return sym_launch(ctx, l, direct_sym_tiles<8, 1, true, 2>, tiles, smem, T, 2); }
	return sym_launch(ctx, l, direct_sym_tiles<8, 1, true, 1>, tiles, smem, T, 1);
#endif
}

size_t sym_lane_tiles(const nb200_ctx* ctx, const nb200_lane& l, int T)
{
	const int		S = static_cast<int>((ctx->n + T - 1) / T);
	const long long	total = static_cast<long long>(S) * (S + 1) / 2;
	const int		G = ctx->nshards;
	return static_cast<size_t>((total - l.shard + G - 1) / G);
}

void sym_lane_release(nb200_lane& l)
{
	if(l.sym_tiles) { cudaFree(l.sym_tiles); }
	if(l.sym_prow) { cudaFree(l.sym_prow); }
	if(l.sym_pcol) { cudaFree(l.sym_pcol); }
	if(l.sym_acc) { cudaFree(l.sym_acc); }
	l.sym_tiles = nullptr;
	l.sym_prow = l.sym_pcol = l.sym_acc = nullptr;
	l.sym_edge = 0;
	l.sym_ntiles = 0;
}

// Tile list and scratch of one lane for tile edge T (kept until the problem or the edge changes). NB200_ERR_ALLOC when the
// tile partials (24 N^2 / T bytes over all shards) do not fit next to the caller's buffers.
int sym_lane_prepare(nb200_ctx* ctx, nb200_lane& l, int T)
{
	const int		S = static_cast<int>((ctx->n + T - 1) / T);
	const int		G = ctx->nshards;
	const size_t	mine = sym_lane_tiles(ctx, l, T);
	CU(ctx, cudaSetDevice(l.dev));
	if(l.sym_edge == T && l.sym_ntiles == mine) { return NB200_OK; }
	sym_lane_release(l);
	std::vector<int2> rc;
	rc.reserve(mine);
	long long id = 0;
	for(int r = 0; r < S; ++r)
	{
		for(int c = r; c < S; ++c, ++id)
		{
			if(id % G == l.shard) { rc.push_back(make_int2(r, c)); }
		}
	}
	const size_t tile_bytes = 3 * static_cast<size_t>(T) * sizeof(real);
	if(cudaMalloc(&l.sym_tiles, std::max<size_t>(1, mine) * sizeof(int2)) != cudaSuccess ||
	   cudaMalloc(&l.sym_prow, std::max<size_t>(1, mine) * tile_bytes) != cudaSuccess ||
	   cudaMalloc(&l.sym_pcol, std::max<size_t>(1, mine) * tile_bytes) != cudaSuccess ||
	   cudaMalloc(&l.sym_acc, (3 * ctx->n + 3 * ctx->n_shard) * sizeof(real)) != cudaSuccess)
	{
		cudaGetLastError();
		sym_lane_release(l);
		return fail(ctx, NB200_ERR_ALLOC, "fcompute_direct: symmetric-tile scratch allocation failed (%zu tiles of %d)", mine, T);
	}
	CU(ctx, cudaMemcpyAsync(l.sym_tiles, rc.data(), mine * sizeof(int2), cudaMemcpyHostToDevice, l.stream));
	CU(ctx, cudaStreamSynchronize(l.stream));
	l.sym_edge = T;
	l.sym_ntiles = mine;
	return NB200_OK;
}

// Tiles of shard `l.shard` (dealt round-robin over all shards) -> l.sym_acc = this shard's partial accelerations of
// ALL bodies, laid out [shard][3][n_shard]. The lane's scratch is in place (sym_lane_prepare).
int sym_lane_partials(nb200_ctx* ctx, nb200_lane& l, int T)
{
	const int		S = static_cast<int>((ctx->n + T - 1) / T);
	const int		G = ctx->nshards;
	const size_t	mine = l.sym_ntiles;
	CU(ctx, cudaSetDevice(l.dev));
	if(ctx->opt_timing) { CU(ctx, cudaEventRecord(l.ev_t[2], l.stream)); }
	const size_t smem = 3 * static_cast<size_t>(T) * sizeof(real);
	if(mine > 0)
	{
		int rc = NB200_OK;
		switch(ctx->opt_sym_shape)
		{
		case NB200_SYM_DEFAULT_SHAPE: rc = sym_launch_default(ctx, l, mine, smem, T); break;
#if NB200_PRECISION == 1
		case 5: rc = sym_launch(ctx, l, direct_sym_tiles_f32x2<4>, mine, smem, T); break;
		case 7: rc = sym_launch(ctx, l, direct_sym_tiles_f32x2<8, true>, mine, smem, T); break;	// all shuffles after the pairs (A/B)
#endif
		case 6: rc = sym_launch(ctx, l, direct_sym_tiles<4, 2, false>, mine, smem, T); break;	// each column body shuffled right after its pairs (A/B: slower)
		case 3: rc = sym_launch(ctx, l, direct_sym_tiles<4, 4>, mine, smem, T); break;
		case 2: rc = sym_launch(ctx, l, direct_sym_tiles<8, 2>, mine, smem, T); break;
		case 1: rc = sym_launch(ctx, l, direct_sym_tiles<4, 2>, mine, smem, T); break;	// FP64: round 1's default; FP32: scalar arithmetic
		default: rc = sym_launch(ctx, l, direct_sym_tiles<8, 1>, mine, smem, T); break;
		}
		if(rc != NB200_OK) { return rc; }
	}
	if(ctx->opt_timing) { CU(ctx, cudaEventRecord(l.ev_t[3], l.stream)); }
	direct_sym_reduce<<<static_cast<unsigned>((ctx->n + 255) / 256), 256, 0, l.stream>>>(
		l.sym_prow, l.sym_pcol, l.sym_acc, ctx->n, ctx->n_shard, T, S, l.shard, G);
	LAUNCHED(ctx);
	return NB200_OK;
}

int sym_fcompute(nb200_ctx* ctx, const nb200_buf* y, nb200_buf* f, int T)
{
	// phase 0: scratch. Whether the symmetric path runs must be ONE decision for all shards: a rank that fell back alone
	// would leave the others waiting in ncclReduceScatter. Scratch is (re)allocated at the same call on every rank (same
	// N, edge and options everywhere), so that call ends with a min-all-reduce of "it fits here"; if it does not fit
	// somewhere, everybody releases the scratch and takes the ordered-pair kernel.
	bool	fresh = false;
	int		fits = 1;
	for(auto& l : ctx->lanes)
	{
		fresh = fresh || l.sym_edge != T || l.sym_ntiles != sym_lane_tiles(ctx, l, T);
	}
	if(fresh)
	{
		for(auto& l : ctx->lanes)
		{
			int rc = sym_lane_prepare(ctx, l, T);
			if(rc == NB200_ERR_ALLOC) { fits = 0; break; }
			if(rc != NB200_OK) { return rc; }
		}
		if(ctx->nranks > 1)
		{
			nb200_lane& l = ctx->lanes[0];
			CU(ctx, cudaSetDevice(l.dev));
			l.h_scalar[6] = static_cast<unsigned long long>(fits);
			CU(ctx, cudaMemcpyAsync(l.d_scalar + 6, l.h_scalar + 6, sizeof(unsigned long long), cudaMemcpyHostToDevice, l.stream));
			NC(ctx, ctx->nccl->AllReduce(l.d_scalar + 6, l.d_scalar + 6, 1, ncclUint64, ncclMin, static_cast<ncclComm_t>(ctx->comm), l.stream));
			CU(ctx, cudaMemcpyAsync(l.h_scalar + 6, l.d_scalar + 6, sizeof(unsigned long long), cudaMemcpyDeviceToHost, l.stream));
			CU(ctx, cudaStreamSynchronize(l.stream));
			fits = static_cast<int>(l.h_scalar[6]);
		}
		if(!fits)
		{
			for(auto& l : ctx->lanes)
			{
				cudaSetDevice(l.dev);
				sym_lane_release(l);
			}
			return fail(ctx, NB200_ERR_ALLOC, "fcompute_direct: symmetric-tile scratch does not fit on every shard");
		}
	}
	// phase 1: every lane turns its tiles into a partial acceleration vector over all bodies
	for(auto& l : ctx->lanes)
	{
		int rc = sym_lane_partials(ctx, l, T);
		if(rc != NB200_OK) { return rc; }
	}
	const size_t	nl = ctx->lanes.size();
	const unsigned	grid3 = static_cast<unsigned>((3 * ctx->n_shard + 255) / 256);
	// phase 2: sum the partials of the own shard across shards
	if(nl > 1 && ctx->lanes_nccl)
	{
		NC(ctx, ctx->nccl->GroupStart());
		for(auto& l : ctx->lanes)
		{
			ncclResult_t r = ctx->nccl->ReduceScatter(l.sym_acc, l.sym_acc + 3 * ctx->n, 3 * ctx->n_shard, NB200_NCCL_REAL, ncclSum,
													  static_cast<ncclComm_t>(l.lane_comm), l.stream);
			if(r != ncclSuccess)
			{
				ctx->nccl->GroupEnd();
				return fail(ctx, NB200_ERR_NCCL, "ncclReduceScatter (lane %d): %s", l.shard, ctx->nccl->GetErrorString(r));
			}
		}
		NC(ctx, ctx->nccl->GroupEnd());
	}
	else if(nl > 1)
	{
		sym_peers peers;
		peers.count = static_cast<int>(nl);
		for(size_t g = 0; g < nl; ++g) { peers.partial[g] = ctx->lanes[g].sym_acc; }
		for(auto& l : ctx->lanes)
		{
			CU(ctx, cudaSetDevice(l.dev));
			CU(ctx, cudaEventRecord(l.ev_packed, l.stream));	// "my partial vector is complete"
		}
		for(auto& l : ctx->lanes)
		{
			CU(ctx, cudaSetDevice(l.dev));
			for(auto& p : ctx->lanes)
			{
				if(&p != &l) { CU(ctx, cudaStreamWaitEvent(l.stream, p.ev_packed, 0)); }
			}
			direct_sym_peer_sum<<<grid3, 256, 0, l.stream>>>(peers, l.sym_acc + 3 * ctx->n, ctx->n_shard, l.shard);
			LAUNCHED(ctx);
			CU(ctx, cudaEventRecord(l.ev_gathered, l.stream));	// "I no longer read the peers' partials"
		}
		for(auto& l : ctx->lanes)
		{
			CU(ctx, cudaSetDevice(l.dev));
			for(auto& p : ctx->lanes)
			{
				if(&p != &l) { CU(ctx, cudaStreamWaitEvent(l.stream, p.ev_gathered, 0)); }
			}
		}
	}
	else if(ctx->nranks > 1)
	{
		nb200_lane& l = ctx->lanes[0];
		NC(ctx, ctx->nccl->ReduceScatter(l.sym_acc, l.sym_acc + 3 * ctx->n, 3 * ctx->n_shard, NB200_NCCL_REAL, ncclSum,
										 static_cast<ncclComm_t>(ctx->comm), l.stream));
	}
	// phase 3: f = (v, a)
	for(size_t li = 0; li < nl; ++li)
	{
		nb200_lane& l = ctx->lanes[li];
		CU(ctx, cudaSetDevice(l.dev));
		const real* acc = ctx->nshards > 1 ? l.sym_acc + 3 * ctx->n : l.sym_acc;
		direct_sym_finish<<<grid3, 256, 0, l.stream>>>(acc, lane_ptr(y, li), lane_ptr(f, li), ctx->n_shard);
		LAUNCHED(ctx);
		if(ctx->opt_timing) { CU(ctx, cudaEventRecord(l.ev_t[4], l.stream)); }
	}
	return NB200_OK;
}
}  // namespace

NB200_API int nb200_fcompute_direct(nb200_ctx* ctx, const nb200_buf* y, nb200_buf* f)
{
	int rc = check_state_pair(ctx, y, f, "fcompute_direct");
	if(rc != NB200_OK) { return rc; }
	if(y == f) { return fail(ctx, NB200_ERR_ARG, "fcompute_direct: y and f must differ"); }
	STEP_NOTE(ctx, make_op(SOP_FCOMPUTE_DIRECT, y, f));
	if(ctx->nshards == 1 && ctx->opt_direct_small != 0 && ctx->n * sizeof(body4) <= NB200_SMALL_MAX_SMEM &&
	   (ctx->opt_direct_small > 0 || (ctx->n <= NB200_SMALL_MAX_BODIES && ctx->opt_direct_sym < 0 && ctx->opt_direct_ipt == 0 && ctx->opt_direct_segments == 0)))
	{
		// small system: one launch does pairs, slice reduction and the velocity rows, straight from the state vector
		nb200_lane& l = ctx->lanes[0];
		CU(ctx, cudaSetDevice(l.dev));
		if(ctx->opt_timing) { CU(ctx, cudaEventRecord(l.ev_t[0], l.stream)); CU(ctx, cudaEventRecord(l.ev_t[1], l.stream)); CU(ctx, cudaEventRecord(l.ev_t[2], l.stream)); }
		const int n = static_cast<int>(ctx->n);
		const size_t smem = ctx->n * sizeof(body4);
		if(l.small_smem == 0)
		{
			// the same ceiling from every context of the process, so that none lowers what another one relies on
			CU(ctx, cudaFuncSetAttribute(direct_small, cudaFuncAttributeMaxDynamicSharedMemorySize, NB200_SMALL_MAX_SMEM));
			l.small_smem = NB200_SMALL_MAX_SMEM;
		}
		direct_small<<<static_cast<unsigned>((n + NB200_SMALL_TARGETS - 1) / NB200_SMALL_TARGETS), NB200_SMALL_TARGETS * NB200_SMALL_SLICES, smem, l.stream>>>(
			lane_ptr(y, 0), l.mass, lane_ptr(f, 0), n);
		LAUNCHED(ctx);
		if(ctx->opt_timing) { CU(ctx, cudaEventRecord(l.ev_t[3], l.stream)); CU(ctx, cudaEventRecord(l.ev_t[4], l.stream)); }
		ctx->last_direct_path = -1;
		return NB200_OK;
	}
	rc = pack_and_gather(ctx, y);
	if(rc != NB200_OK) { return rc; }
	if(const int edge = sym_tile_edge(ctx))
	{
		ctx->last_direct_path = edge;
		rc = sym_fcompute(ctx, y, f, edge);
		if(rc != NB200_ERR_ALLOC) { return rc; }
		// the tile partials (24 N^2 / T bytes) did not fit next to the caller's buffers on some shard: ordered-pair kernel for
		// this body set from now on, on every shard alike (sym_fcompute agreed on it). Scratch is only ever (re)allocated by
		// the first fcompute after set_bodies or an option change -- both empty the step table, and the step after that runs
		// eagerly -- so no recorded step or graph describes the path given up here.
		ctx->sym_unavailable = true;
	}
	ctx->last_direct_path = 0;

	const int	n_tiles = static_cast<int>(ctx->n_pad / NB200_DIRECT_TILE);
	const int	sms = ctx->lanes[0].sm_count;
	// Targets per thread: as many as still leave >= 2 CTAs per SM worth of (target block, tile) items.
	// FP32 pairs are issue-bound (13 FFMA-pipe + MUFU + FMNMX per pair): 8 targets per thread halve the LDS share.
	const int	ipt_max = sizeof(real) == 4 ? 8 : 4;
	int ipt = 1;
	if(ctx->opt_direct_ipt == 1 || ctx->opt_direct_ipt == 2 || ctx->opt_direct_ipt == 4 || (ctx->opt_direct_ipt == 8 && ipt_max == 8))
	{
		ipt = static_cast<int>(ctx->opt_direct_ipt);
	}
	else
	{
		for(int cand = ipt_max; cand >= 1; cand >>= 1)
		{
			size_t blocks = (ctx->n_shard + NB200_DIRECT_THREADS * cand - 1) / (NB200_DIRECT_THREADS * cand);
			if(blocks * static_cast<size_t>(n_tiles) >= static_cast<size_t>(2 * sms) || cand == 1)
			{
				ipt = cand;
				break;
			}
		}
	}
	const size_t	i_blocks = (ctx->n_shard + NB200_DIRECT_THREADS * ipt - 1) / (NB200_DIRECT_THREADS * ipt);
	// Source segments: enough equal work items (~48 per SM) that the last wave costs <~ 2 %.
	size_t want = static_cast<size_t>(sms) * 48;
	size_t seg = (want + i_blocks - 1) / i_blocks;
	if(ctx->opt_direct_segments > 0) { seg = static_cast<size_t>(ctx->opt_direct_segments); }
	seg = std::max<size_t>(1, std::min<size_t>(seg, std::min<size_t>(static_cast<size_t>(n_tiles), 64)));
	const int	tiles_per_seg = static_cast<int>((static_cast<size_t>(n_tiles) + seg - 1) / seg);
	const int	segments = (n_tiles + tiles_per_seg - 1) / tiles_per_seg;

	for(size_t li = 0; li < ctx->lanes.size(); ++li)
	{
		nb200_lane& l = ctx->lanes[li];
		CU(ctx, cudaSetDevice(l.dev));
		if(ctx->opt_timing) { CU(ctx, cudaEventRecord(l.ev_t[2], l.stream)); }
		real* out = lane_ptr(f, li);
		if(segments > 1)
		{
			size_t need = static_cast<size_t>(segments) * 3 * ctx->n_shard;
			if(l.partial_elems < need)
			{
				if(l.partial) { cudaFree(l.partial); l.partial = nullptr; l.partial_elems = 0; }
				if(cudaMalloc(&l.partial, need * sizeof(real)) != cudaSuccess)
				{
					cudaGetLastError();
					return fail(ctx, NB200_ERR_ALLOC, "fcompute_direct: partial-sum scratch allocation failed");
				}
				l.partial_elems = need;
			}
			out = l.partial;
		}
		dim3 grid(static_cast<unsigned>(i_blocks), static_cast<unsigned>(segments));
		const int write_f = segments == 1 ? 1 : 0;
		switch(ipt)
		{
#if NB200_PRECISION == 1
		case 8: launch_pairs<8>(ctx, l, lane_ptr(y, li), out, grid, n_tiles, tiles_per_seg, write_f); break;
#endif
		case 4: launch_pairs<4>(ctx, l, lane_ptr(y, li), out, grid, n_tiles, tiles_per_seg, write_f); break;
		case 2: launch_pairs<2>(ctx, l, lane_ptr(y, li), out, grid, n_tiles, tiles_per_seg, write_f); break;
		default: launch_pairs<1>(ctx, l, lane_ptr(y, li), out, grid, n_tiles, tiles_per_seg, write_f); break;
		}
		LAUNCHED(ctx);
		if(ctx->opt_timing) { CU(ctx, cudaEventRecord(l.ev_t[3], l.stream)); }
		if(segments > 1)
		{
			unsigned rgrid = static_cast<unsigned>((3 * ctx->n_shard + 255) / 256);
			direct_reduce<<<rgrid, 256, 0, l.stream>>>(l.partial, lane_ptr(y, li), lane_ptr(f, li), ctx->n_shard, segments);
			LAUNCHED(ctx);
		}
		if(ctx->opt_timing) { CU(ctx, cudaEventRecord(l.ev_t[4], l.stream)); }
	}
	return NB200_OK;
}

// ---- Barnes-Hut ------------------------------------------------------------------------
NB200_API int nb200_bh_configure(nb200_ctx* ctx, nb200_real ratio, int layout, size_t tree_build_rate)
{
	if(ctx == nullptr) { return NB200_ERR_ARG; }
	if(layout != NB200_TREE_HEAP && layout != NB200_TREE_HEAP_STACKLESS)
	{
		return fail(ctx, NB200_ERR_ARG, "bh_configure: tree_layout must be heap or heap_stackless");
	}
	step_invalidate(ctx);
	ctx->bh_ratio = ratio;
	ctx->bh_layout = layout;
	ctx->bh_build_rate = tree_build_rate;
	return NB200_OK;
}

NB200_API int nb200_fcompute_bh(nb200_ctx* ctx, const nb200_buf* y, nb200_buf* f, size_t step)
{
	int rc = check_state_pair(ctx, y, f, "fcompute_bh");
	if(rc != NB200_OK) { return rc; }
	if(y == f) { return fail(ctx, NB200_ERR_ARG, "fcompute_bh: y and f must differ"); }
	if((ctx->n & (ctx->n - 1)) != 0)
	{
		return fail(ctx, NB200_ERR_UNSUPPORTED, "fcompute_bh: N = %zu is not a power of two (kd-heap leaves are [N, 2N))", ctx->n);
	}
	{
		step_op op = make_op(SOP_FCOMPUTE_BH, y, f);
		op.step = step;
		// the step number only matters through "does this call rebuild the tree" (a first call always does: the table is
		// emptied by set_bodies, and the first step after that runs eagerly)
		op.key = ctx->bh_build_rate == 0 ? 0 : (step % ctx->bh_build_rate == 0 ? 1 : 2);
		STEP_NOTE(ctx, op);
	}
	rc = pack_and_gather(ctx, y);
	if(rc != NB200_OK) { return rc; }
	for(size_t li = 0; li < ctx->lanes.size(); ++li)
	{
		nb200_lane& l = ctx->lanes[li];
		CU(ctx, cudaSetDevice(l.dev));
		std::string err;
		int launches = 0;
		rc = bh_fcompute(ctx, l, lane_ptr(y, li), lane_ptr(f, li), step, launches, err);
		ctx->launches += static_cast<unsigned long long>(launches);
		if(rc != NB200_OK) { return fail(ctx, rc, "fcompute_bh: %s", err.c_str()); }
	}
	if(ctx->nshards > 1)
	{
		// every shard walked its contiguous leaves: gather the leaf-ordered accelerations, then pick own bodies
		rc = gather_blocks(ctx, 1);
		if(rc != NB200_OK) { return rc; }
		for(size_t li = 0; li < ctx->lanes.size(); ++li)
		{
			nb200_lane& l = ctx->lanes[li];
			CU(ctx, cudaSetDevice(l.dev));
			std::string err;
			int launches = 0;
			rc = bh_scatter(ctx, l, lane_ptr(y, li), lane_ptr(f, li), launches, err);
			ctx->launches += static_cast<unsigned long long>(launches);
			if(rc != NB200_OK) { return fail(ctx, rc, "fcompute_bh: %s", err.c_str()); }
		}
	}
	return NB200_OK;
}

NB200_API int nb200_bh_export_tree(nb200_ctx* ctx, int lane, nb200_real* xyzr, nb200_real* mass, int* body_n)
{
	if(ctx == nullptr || lane < 0 || static_cast<size_t>(lane) >= ctx->lanes.size()) { return NB200_ERR_ARG; }
	nb200_lane& l = ctx->lanes[static_cast<size_t>(lane)];
	step_break(ctx);
	if(l.bh == nullptr) { return fail(ctx, NB200_ERR_STATE, "bh_export_tree: no tree has been built"); }
	CU(ctx, cudaSetDevice(l.dev));
	std::string err;
	int rc = bh_export(ctx, l, xyzr, mass, body_n, err);
	if(rc != NB200_OK) { return fail(ctx, rc, "bh_export_tree: %s", err.c_str()); }
	return NB200_OK;
}

NB200_API int nb200_bh_walk_stats(nb200_ctx* ctx, int enable, unsigned long long* visits, unsigned long long* interactions)
{
	if(ctx == nullptr) { return NB200_ERR_ARG; }
	step_invalidate(ctx);
	unsigned long long v = 0, k = 0;
	if(ctx->bh_stats)
	{
		for(auto& l : ctx->lanes)
		{
			CU(ctx, cudaSetDevice(l.dev));
			CU(ctx, cudaMemcpyAsync(l.h_scalar + 2, l.d_scalar + 2, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, l.stream));
			CU(ctx, cudaStreamSynchronize(l.stream));
			v += l.h_scalar[2];
			k += l.h_scalar[3];
		}
	}
	if(visits) { *visits = v; }
	if(interactions) { *interactions = k; }
	ctx->bh_stats = enable != 0;
	return NB200_OK;
}

NB200_API int nb200_bh_walk_profile(nb200_ctx* ctx, unsigned long long out[16])
{
	if(ctx == nullptr || out == nullptr) { return NB200_ERR_ARG; }
	step_invalidate(ctx);
	for(int q = 0; q < 16; ++q) { out[q] = 0; }
	for(auto& l : ctx->lanes)
	{
		CU(ctx, cudaSetDevice(l.dev));
		CU(ctx, cudaMemcpyAsync(l.h_scalar + 8, l.d_scalar + 8, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, l.stream));
		CU(ctx, cudaStreamSynchronize(l.stream));
		for(int q = 0; q < 16; ++q)
		{
			out[q] = q == 5 ? std::max(out[q], l.h_scalar[8 + q]) : out[q] + l.h_scalar[8 + q];
		}
	}
	return NB200_OK;
}

// ---- state-vector ops ---------------------------------------------------------------------
NB200_API int nb200_fmadd_inplace(nb200_ctx* ctx, nb200_buf* a, const nb200_buf* b, nb200_real c)
{
	if(ctx == nullptr) { return NB200_ERR_ARG; }
	if(!valid(ctx, a)) { return fail(ctx, NB200_ERR_ARG, "fmadd_inplace: a is not a buffer of this context"); }
	if(!valid(ctx, b)) { return fail(ctx, NB200_ERR_ARG, "fmadd_inplace: b is not a buffer of this context"); }
	if(!same_shape(a, b)) { return fail(ctx, NB200_ERR_ARG, "fmadd_inplace: size does not match"); }
	if(a->lane_elems == 0) { return NB200_OK; }
	{
		step_op op = make_op(SOP_FMADD_INPLACE, a, b);
		op.coef.push_back(c);
		STEP_NOTE(ctx, op);
	}
	for(size_t i = 0; i < ctx->lanes.size(); ++i)
	{
		nb200_lane& l = ctx->lanes[i];
		CU(ctx, cudaSetDevice(l.dev));
		ew_fmadd_inplace<<<ew_grid(l, a->lane_elems), NB200_EW_THREADS, 0, l.stream>>>(lane_ptr(a, i), lane_ptr(b, i), c, a->lane_elems);
		LAUNCHED(ctx);
	}
	return NB200_OK;
}

NB200_API int nb200_fmadd(nb200_ctx* ctx, nb200_buf* a, const nb200_buf* b, const nb200_buf* c, nb200_real d)
{
	if(ctx == nullptr) { return NB200_ERR_ARG; }
	if(!valid(ctx, a)) { return fail(ctx, NB200_ERR_ARG, "fmadd: a is not a buffer of this context"); }
	if(!valid(ctx, b)) { return fail(ctx, NB200_ERR_ARG, "fmadd: b is not a buffer of this context"); }
	if(!valid(ctx, c)) { return fail(ctx, NB200_ERR_ARG, "fmadd: c is not a buffer of this context"); }
	if(!same_shape(a, b) || !same_shape(a, c)) { return fail(ctx, NB200_ERR_ARG, "fmadd: size does not match"); }
	if(a->lane_elems == 0) { return NB200_OK; }
	{
		step_op op = make_op(SOP_FMADD, a, b, c);
		op.coef.push_back(d);
		STEP_NOTE(ctx, op);
	}
	for(size_t i = 0; i < ctx->lanes.size(); ++i)
	{
		nb200_lane& l = ctx->lanes[i];
		CU(ctx, cudaSetDevice(l.dev));
		ew_fmadd<<<ew_grid(l, a->lane_elems), NB200_EW_THREADS, 0, l.stream>>>(lane_ptr(a, i), lane_ptr(b, i), lane_ptr(c, i), d, a->lane_elems);
		LAUNCHED(ctx);
	}
	return NB200_OK;
}

namespace {
// Shared body of fmaddn / fmaddn_inplace / fmaddn_corr.
//   base: starting value (NULL = zeros); for the in-place forms base == a.
//   corr: non-NULL selects the Kahan kernel.
int fused_terms(nb200_ctx* ctx, const char* who, int kind, nb200_buf* a, const nb200_buf* base, nb200_buf* corr,
				const nb200_buf* const* terms, const nb200_real* coeff, size_t n, bool keep_a_if_no_terms)
{
	// Collect non-zero terms first: zero coefficients are skipped before their buffers are even looked at
	// (nbody/nbody_engine.cpp:58-63,95-99).
	std::vector<size_t> used;
	for(size_t k = 0; k < n; ++k)
	{
		if(coeff[k] == static_cast<real>(0)) { continue; }
		if(terms == nullptr || !valid(ctx, terms[k]))
		{
			return fail(ctx, NB200_ERR_ARG, "%s: term %zu is not a buffer of this context", who, k);
		}
		if(!same_shape(terms[k], a)) { return fail(ctx, NB200_ERR_ARG, "%s: term %zu size does not match", who, k); }
		used.push_back(k);
	}
	if(used.empty())
	{
		if(keep_a_if_no_terms) { return NB200_OK; }	// a is left untouched, as in the reference
		return nb200_fill(ctx, a, 0);				// fmaddn with b == NULL: a = 0
	}
	if(a->lane_elems == 0) { return NB200_OK; }
	{
		// recorded with the zero-coefficient terms already dropped: re-issuing it gives the same launches
		step_op op = make_op(kind, a, kind == SOP_FMADDN_CORR ? corr : (kind == SOP_FMADDN ? base : nullptr));
		for(size_t k : used)
		{
			op.list.push_back(terms[k]);
			op.coef.push_back(coeff[k]);
		}
		STEP_NOTE(ctx, op);
	}
	for(size_t first = 0; first < used.size(); first += NB200_MAX_TERMS)
	{
		size_t cnt = std::min<size_t>(NB200_MAX_TERMS, used.size() - first);
		for(size_t i = 0; i < ctx->lanes.size(); ++i)
		{
			nb200_lane& l = ctx->lanes[i];
			CU(ctx, cudaSetDevice(l.dev));
			nb200_terms t;
			t.n = static_cast<int>(cnt);
			for(size_t k = 0; k < cnt; ++k)
			{
				t.p[k] = lane_ptr(terms[used[first + k]], i);
				t.c[k] = coeff[used[first + k]];
			}
			unsigned grid = ew_grid(l, a->lane_elems);
			if(corr != nullptr)
			{
				ew_fmaddn_corr<<<grid, NB200_EW_THREADS, 0, l.stream>>>(lane_ptr(a, i), lane_ptr(corr, i), t, a->lane_elems);
			}
			else
			{
				const real* b0 = (first == 0) ? (base != nullptr ? lane_ptr(base, i) : nullptr) : lane_ptr(a, i);
				ew_fmaddn<<<grid, NB200_EW_THREADS, 0, l.stream>>>(lane_ptr(a, i), b0, t, a->lane_elems);
			}
			LAUNCHED(ctx);
		}
	}
	return NB200_OK;
}
}  // namespace

NB200_API int nb200_fmaddn_inplace(nb200_ctx* ctx, nb200_buf* a, const nb200_buf* const* b, const nb200_real* c, size_t n)
{
	if(ctx == nullptr) { return NB200_ERR_ARG; }
	if(c == nullptr) { return fail(ctx, NB200_ERR_ARG, "fmaddn_inplace: c == NULL"); }
	if(!valid(ctx, a)) { return fail(ctx, NB200_ERR_ARG, "fmaddn_inplace: a is not a buffer of this context"); }
	return fused_terms(ctx, "fmaddn_inplace", SOP_FMADDN_INPLACE, a, a, nullptr, b, c, n, true);
}

NB200_API int nb200_fmaddn(nb200_ctx* ctx, nb200_buf* a, const nb200_buf* b, const nb200_buf* const* c, const nb200_real* d, size_t n)
{
	if(ctx == nullptr) { return NB200_ERR_ARG; }
	if(d == nullptr) { return fail(ctx, NB200_ERR_ARG, "fmaddn: d == NULL"); }
	if(!valid(ctx, a)) { return fail(ctx, NB200_ERR_ARG, "fmaddn: a is not a buffer of this context"); }
	if(b != nullptr)
	{
		if(!valid(ctx, b)) { return fail(ctx, NB200_ERR_ARG, "fmaddn: b is not a buffer of this context"); }
		if(!same_shape(a, b)) { return fail(ctx, NB200_ERR_ARG, "fmaddn: size does not match"); }
	}
	// Reference quirk kept: with b != NULL and no non-zero term, a is NOT assigned (nbody_engine.cpp:87-112).
	return fused_terms(ctx, "fmaddn", SOP_FMADDN, a, b, nullptr, c, d, n, b != nullptr);
}

NB200_API int nb200_fmaddn_corr(nb200_ctx* ctx, nb200_buf* a, nb200_buf* corr, const nb200_buf* const* b, const nb200_real* c, size_t n)
{
	if(ctx == nullptr) { return NB200_ERR_ARG; }
	if(!valid(ctx, a)) { return fail(ctx, NB200_ERR_ARG, "fmaddn_corr: a is not a buffer of this context"); }
	if(!valid(ctx, corr)) { return fail(ctx, NB200_ERR_ARG, "fmaddn_corr: corr is not a buffer of this context"); }
	if(c == nullptr) { return fail(ctx, NB200_ERR_ARG, "fmaddn_corr: c == NULL"); }
	if(!same_shape(a, corr)) { return fail(ctx, NB200_ERR_ARG, "fmaddn_corr: size does not match"); }
	if(a == corr) { return fail(ctx, NB200_ERR_ARG, "fmaddn_corr: a and corr must differ"); }
	// The reference validates every b[k], zero coefficient or not (nbody_engine_cuda.cpp:431-439).
	for(size_t k = 0; k < n; ++k)
	{
		if(b == nullptr || !valid(ctx, b[k])) { return fail(ctx, NB200_ERR_ARG, "fmaddn_corr: b[%zu] is not a buffer of this context", k); }
	}
	return fused_terms(ctx, "fmaddn_corr", SOP_FMADDN_CORR, a, a, corr, b, c, n, true);
}

NB200_API int nb200_fmaxabs(nb200_ctx* ctx, const nb200_buf* a, nb200_real* result)
{
	if(ctx == nullptr || result == nullptr) { return NB200_ERR_ARG; }
	if(!valid(ctx, a)) { return fail(ctx, NB200_ERR_ARG, "fmaxabs: a is not a buffer of this context"); }
	{
		int rc = step_border(ctx, make_op(SOP_FMAXABS, a));
		if(rc != NB200_OK) { return rc; }
	}
	if(a->lane_elems == 0)
	{
		*result = 0;
		return NB200_OK;
	}
	for(size_t i = 0; i < ctx->lanes.size(); ++i)
	{
		nb200_lane& l = ctx->lanes[i];
		CU(ctx, cudaSetDevice(l.dev));
		CU(ctx, cudaMemsetAsync(l.d_scalar, 0, sizeof(unsigned long long), l.stream));
		// element 0 of the vector lives on shard 0 (sharded buffers) or on every lane alike (replicated ones)
		ew_maxabs<<<ew_grid(l, a->lane_elems), NB200_EW_THREADS, 0, l.stream>>>(lane_ptr(a, i), a->lane_elems, l.d_scalar,
																				 (!a->sharded || l.shard == 0) ? 1 : 0);
		LAUNCHED(ctx);
		if(ctx->nranks > 1)
		{
			// Bit patterns of non-negative reals order like unsigned integers: max over ranks is exact.
			NC(ctx, ctx->nccl->AllReduce(l.d_scalar, l.d_scalar, 1, ncclUint64, ncclMax, static_cast<ncclComm_t>(ctx->comm), l.stream));
		}
		CU(ctx, cudaMemcpyAsync(l.h_scalar, l.d_scalar, sizeof(unsigned long long), cudaMemcpyDeviceToHost, l.stream));
	}
	unsigned long long best = 0;
	for(auto& l : ctx->lanes)
	{
		CU(ctx, cudaSetDevice(l.dev));
		CU(ctx, cudaStreamSynchronize(l.stream));
		best = std::max(best, l.h_scalar[0]);
	}
#if NB200_PRECISION == 2
	double out;
	memcpy(&out, &best, sizeof(out));
#else
	unsigned int b32 = static_cast<unsigned int>(best);
	float out;
	memcpy(&out, &b32, sizeof(out));
#endif
	*result = out;
	return NB200_OK;
}

NB200_API int nb200_clamp(nb200_ctx* ctx, nb200_buf* y, nb200_real b)
{
	if(ctx == nullptr) { return NB200_ERR_ARG; }
	if(!valid(ctx, y)) { return fail(ctx, NB200_ERR_ARG, "clamp: y is not a buffer of this context"); }
	if(!y->sharded) { return fail(ctx, NB200_ERR_ARG, "clamp: y must be a state vector of 6*N elements"); }
	{
		step_op op = make_op(SOP_CLAMP, y);
		op.coef.push_back(b);
		STEP_NOTE(ctx, op);
	}
	for(size_t i = 0; i < ctx->lanes.size(); ++i)
	{
		nb200_lane& l = ctx->lanes[i];
		CU(ctx, cudaSetDevice(l.dev));
		size_t count3 = 3 * ctx->n_shard;	// the three position rows are contiguous in a shard
		ew_clamp<<<static_cast<unsigned>((count3 + NB200_EW_THREADS - 1) / NB200_EW_THREADS), NB200_EW_THREADS, 0, l.stream>>>(lane_ptr(y, i), b, count3);
		LAUNCHED(ctx);
	}
	return NB200_OK;
}

// ---- conservation report ----------------------------------------------------------------------------
NB200_API int nb200_statistics(nb200_ctx* ctx, const nb200_buf* y, int with_energy, double out[11])
{
	if(ctx == nullptr || out == nullptr) { return NB200_ERR_ARG; }
	if(!valid(ctx, y)) { return fail(ctx, NB200_ERR_ARG, "statistics: y is not a buffer of this context"); }
	if(ctx->n == 0 || !y->sharded) { return fail(ctx, NB200_ERR_ARG, "statistics: y must be a state vector of 6*N elements"); }
	step_break(ctx);
	if(with_energy)
	{
		int rc = pack_and_gather(ctx, y);	// the potential needs every body's position on every shard
		if(rc != NB200_OK) { return rc; }
	}
	const int	n_tiles = static_cast<int>(ctx->n_pad / NB200_DIRECT_TILE);
	double		total[NB200_STATS_LINEAR + 1] = {0};
	std::vector<double*> scratch(ctx->lanes.size(), nullptr);
	std::vector<double*> host(ctx->lanes.size(), nullptr);
	int status = NB200_OK;
	for(size_t li = 0; li < ctx->lanes.size() && status == NB200_OK; ++li)
	{
		nb200_lane& l = ctx->lanes[li];
		cudaSetDevice(l.dev);
		const int lin_blocks = std::max(1, std::min<int>(l.sm_count * 4, static_cast<int>((ctx->n_shard + NB200_STATS_THREADS - 1) / NB200_STATS_THREADS)));
		const int pot_blocks = with_energy ? static_cast<int>((ctx->n_shard + NB200_STATS_THREADS - 1) / NB200_STATS_THREADS) : 0;
		const size_t elems = static_cast<size_t>(lin_blocks) * NB200_STATS_LINEAR + pot_blocks + NB200_STATS_LINEAR + 1;
		if(cudaMalloc(&scratch[li], elems * sizeof(double)) != cudaSuccess || cudaMallocHost(&host[li], (NB200_STATS_LINEAR + 1) * sizeof(double)) != cudaSuccess)
		{
			cudaGetLastError();
			status = fail(ctx, NB200_ERR_ALLOC, "statistics: scratch allocation failed");
			break;
		}
		double* lin = scratch[li];
		double* pot = lin + static_cast<size_t>(lin_blocks) * NB200_STATS_LINEAR;
		double* res = pot + pot_blocks;
		const size_t first = static_cast<size_t>(l.shard) * ctx->n_shard;
		stats_linear<<<lin_blocks, NB200_STATS_THREADS, 0, l.stream>>>(lane_ptr(y, li), l.mass, ctx->n_shard, first, lin);
		++ctx->launches;
		if(with_energy)
		{
			stats_potential<<<pot_blocks, NB200_STATS_THREADS, 0, l.stream>>>(l.src, ctx->n_shard, first, n_tiles, pot);
			++ctx->launches;
		}
		stats_finish<<<1, 32, 0, l.stream>>>(lin, lin_blocks, pot, pot_blocks, res);
		++ctx->launches;
		if(ctx->nranks > 1 && ctx->nccl->AllReduce(res, res, NB200_STATS_LINEAR + 1, ncclFloat64, ncclSum, static_cast<ncclComm_t>(ctx->comm), l.stream) != ncclSuccess)
		{
			status = fail(ctx, NB200_ERR_NCCL, "statistics: ncclAllReduce failed");
			break;
		}
		if(cudaMemcpyAsync(host[li], res, (NB200_STATS_LINEAR + 1) * sizeof(double), cudaMemcpyDeviceToHost, l.stream) != cudaSuccess)
		{
			status = fail(ctx, NB200_ERR_CUDA, "statistics: %s", cudaGetErrorString(cudaGetLastError()));
		}
	}
	for(size_t li = 0; li < ctx->lanes.size(); ++li)
	{
		nb200_lane& l = ctx->lanes[li];
		cudaSetDevice(l.dev);
		if(status == NB200_OK && cudaStreamSynchronize(l.stream) != cudaSuccess)
		{
			status = fail(ctx, NB200_ERR_CUDA, "statistics: %s", cudaGetErrorString(cudaGetLastError()));
		}
		if(status == NB200_OK && host[li] != nullptr)
		{
			for(int q = 0; q <= NB200_STATS_LINEAR; ++q) { total[q] += host[li][q]; }	// lanes in shard order: fixed
		}
		if(scratch[li]) { cudaFree(scratch[li]); }
		if(host[li]) { cudaFreeHost(host[li]); }
	}
	if(status != NB200_OK) { return status; }
	// linear partial layout: Px Py Pz Lx Ly Lz 2Ekin Cx Cy Cz M
	for(int q = 0; q < 6; ++q) { out[q] = total[q]; }
	out[6] = total[6] / 2;
	out[7] = with_energy ? -total[NB200_STATS_LINEAR] / 2 : 0;
	for(int q = 0; q < 3; ++q) { out[8 + q] = total[7 + q] / total[10]; }
	return NB200_OK;
}

// ---- solver steps as CUDA graphs ------------------------------------------------------------------
NB200_API int nb200_step_boundary(nb200_ctx* ctx)
{
	if(ctx == nullptr) { return NB200_ERR_ARG; }
	step_graph& sg = *ctx->sg;
	if(sg.mode == SG_OFF || sg.busy) { return NB200_OK; }
	int rc = NB200_OK;
	int me = -1;	// the table entry of the step that ends here
	if(sg.mode == SG_REPLAY)
	{
		if(sg.pos == 0) { return NB200_OK; }	// no calls since the last boundary
		if(sg.pos != sg.entries[static_cast<size_t>(sg.target)].seq.size())
		{
			// the step ended before the predicted one did; it may be exactly another entry that has graphs
			const std::vector<step_op>& have = sg.entries[static_cast<size_t>(sg.target)].seq;
			for(size_t k = 0; k < sg.entries.size(); ++k)
			{
				const step_entry& alt = sg.entries[k];
				if(!alt.captured || alt.seq.size() != sg.pos) { continue; }
				bool prefix = true;
				for(size_t q = 0; prefix && q < sg.pos; ++q) { prefix = alt.seq[q].same(have[q]); }
				if(prefix)
				{
					sg.target = static_cast<int>(k);
					break;
				}
			}
		}
		const step_entry& e = sg.entries[static_cast<size_t>(sg.target)];
		if(sg.pos == e.seq.size())
		{
			if(e.execs[sg.seg] != nullptr)
			{
				nb200_lane& l = ctx->lanes[0];
				CU(ctx, cudaSetDevice(l.dev));
				CU(ctx, cudaGraphLaunch(e.execs[sg.seg], l.stream));
				ctx->launches += e.seg_launches[sg.seg];
				++sg.graph_launches;
			}
			sg.launches_per_step = 0;
			for(unsigned long long c : e.seg_launches) { sg.launches_per_step += c; }
			sg.misses = 0;
			me = sg.target;
		}
		else
		{
			rc = step_bail(ctx);	// the step ended before the predicted one did: now a recording, handled below
			if(rc != NB200_OK || sg.mode == SG_OFF) { return rc; }
		}
	}
	if(sg.mode == SG_CAPTURE)
	{
		if(sg.cur.empty()) { return NB200_OK; }
		rc = step_close_segment(ctx);	// the last segment of the step
		if(sg.mode == SG_OFF) { return rc; }
		const bool as_predicted = step_same_sequence(sg.cur, sg.entries[static_cast<size_t>(sg.target)].seq);
		me = as_predicted ? sg.target : step_find_or_add(ctx, sg.cur);
		step_entry& e = sg.entries[static_cast<size_t>(me)];
		if(!e.captured)
		{
			// whichever step this was, it has a graph now
			e.execs.swap(sg.cap_execs);
			e.seg_launches.swap(sg.cap_launches);
			e.captured = true;
		}
		step_destroy_execs(ctx, sg.cap_execs);
		sg.cap_launches.clear();
		if(as_predicted) { sg.misses = 0; } else { step_miss(ctx); }
		if(sg.mode == SG_OFF) { return rc; }
	}
	else if(sg.mode == SG_RECORD)
	{
		// a clean step (nothing host-visible inside except segment borders) is filed; an unclean one breaks the chain
		me = sg.clean ? step_find_or_add(ctx, sg.cur) : -1;
	}
	sg.cur.clear();
	sg.clean = true;
	sg.pos = sg.seg = sg.seg_start = 0;
	if(sg.last >= 0 && me >= 0 && static_cast<size_t>(sg.last) < sg.entries.size())
	{
		sg.entries[static_cast<size_t>(sg.last)].next = me;
	}
	sg.last = me;
	// the next step is predicted to be what followed this entry last time
	sg.target = me >= 0 ? sg.entries[static_cast<size_t>(me)].next : -1;
	if(sg.target < 0) { sg.mode = SG_RECORD; }
	else { sg.mode = sg.entries[static_cast<size_t>(sg.target)].captured ? SG_REPLAY : SG_CAPTURE; }
	return rc;
}

NB200_API int nb200_step_graph_stats(const nb200_ctx* ctx, unsigned long long out[5])
{
	if(ctx == nullptr || out == nullptr) { return NB200_ERR_ARG; }
	out[0] = ctx->sg->graph_launches;
	out[1] = ctx->sg->bailouts;
	out[2] = static_cast<unsigned long long>(ctx->sg->mode);
	out[3] = ctx->sg->launches_per_step;
	out[4] = ctx->sg->entries.size();
	return NB200_OK;
}

// ---- instrumentation ---------------------------------------------------------------------------
NB200_API unsigned long long nb200_launch_count(const nb200_ctx* ctx)
{
	return ctx != nullptr ? ctx->launches : 0;
}

NB200_API int nb200_last_fcompute_ms(nb200_ctx* ctx, float out[4])
{
	if(ctx == nullptr || out == nullptr) { return NB200_ERR_ARG; }
	step_break(ctx);
	nb200_lane& l = ctx->lanes[0];
	CU(ctx, cudaSetDevice(l.dev));
	CU(ctx, cudaStreamSynchronize(l.stream));
	for(int i = 0; i < 4; ++i)
	{
		out[i] = 0;
		if(cudaEventElapsedTime(&out[i], l.ev_t[i], l.ev_t[i + 1]) != cudaSuccess)
		{
			cudaGetLastError();
			out[i] = 0;
		}
	}
	return NB200_OK;
}

NB200_API int nb200_last_direct_path(const nb200_ctx* ctx)
{
	return ctx != nullptr ? ctx->last_direct_path : 0;
}

NB200_API int nb200_mark(nb200_ctx* ctx, int slot)
{
	if(ctx == nullptr || slot < 0 || slot >= 8) { return NB200_ERR_ARG; }
	step_break(ctx);	// a mark inside a deferred step would time nothing
	nb200_lane& l = ctx->lanes[0];
	CU(ctx, cudaSetDevice(l.dev));
	CU(ctx, cudaEventRecord(l.ev_mark[slot], l.stream));
	return NB200_OK;
}

NB200_API int nb200_elapsed_ms(nb200_ctx* ctx, int slot_a, int slot_b, float* ms)
{
	if(ctx == nullptr || ms == nullptr || slot_a < 0 || slot_a >= 8 || slot_b < 0 || slot_b >= 8) { return NB200_ERR_ARG; }
	nb200_lane& l = ctx->lanes[0];
	CU(ctx, cudaSetDevice(l.dev));
	CU(ctx, cudaEventSynchronize(l.ev_mark[slot_b]));
	CU(ctx, cudaEventElapsedTime(ms, l.ev_mark[slot_a], l.ev_mark[slot_b]));
	return NB200_OK;
}

NB200_API int nb200_probe_fma_peak(nb200_ctx* ctx, double ms, double* fma_lane_per_s)
{
	if(ctx == nullptr || fma_lane_per_s == nullptr) { return NB200_ERR_ARG; }
	step_break(ctx);
	nb200_lane& l = ctx->lanes[0];
	CU(ctx, cudaSetDevice(l.dev));
	cudaEvent_t e0, e1;
	CU(ctx, cudaEventCreate(&e0));
	CU(ctx, cudaEventCreate(&e1));
	const int		blocks = l.sm_count * 8;
	int				iters = 256;
	float			elapsed = 0;
	double			best = 0;
	// Grow the launch until one run lasts >= ms/4, then keep the best of 4 such runs.
	for(int run = 0, timed = 0; run < 40 && timed < 4; ++run)
	{
		cudaEventRecord(e0, l.stream);
		probe_fma<<<blocks, 256, 0, l.stream>>>(reinterpret_cast<real*>(l.d_scalar + 4), iters, static_cast<real>(1.0));
		++ctx->launches;
		cudaEventRecord(e1, l.stream);
		CU(ctx, cudaEventSynchronize(e1));
		CU(ctx, cudaEventElapsedTime(&elapsed, e0, e1));
		if(elapsed * 4 >= ms || iters >= (1 << 24))
		{
			best = std::max(best, static_cast<double>(blocks) * 256.0 * iters * 64.0 / (elapsed * 1e-3));
			++timed;
		}
		else
		{
			iters *= 2;
		}
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	*fma_lane_per_s = best;
	return NB200_OK;
}

NB200_API int nb200_set_option(nb200_ctx* ctx, const char* name, long long value)
{
	if(ctx == nullptr || name == nullptr) { return NB200_ERR_ARG; }
	step_invalidate(ctx);	// a recorded step was launched with the old settings
	if(strcmp(name, "step_graph") == 0)
	{
		// deferral needs one stream that sees every call: single-shard contexts only (accepted and ignored otherwise)
		const bool can = ctx->lanes.size() == 1 && ctx->nranks == 1;
		step_graph& sg = *ctx->sg;
		step_clear_table(ctx);
		sg.cur.clear();
		sg.clean = true;
		sg.pos = sg.seg = sg.seg_start = 0;
		sg.misses = 0;
		sg.mode = (value != 0 && can) ? SG_RECORD : SG_OFF;
	}
	else if(strcmp(name, "use_nccl") == 0)
	{
		// lanes of ONE process exchange shards with NCCL (one communicator per lane) instead of peer copies / peer loads
		if(value == 0 || ctx->lanes.size() < 2)
		{
			ctx->lanes_nccl = false;	// one lane has nothing to exchange; ranks (one process per GPU) always use NCCL
			return NB200_OK;
		}
		if(ctx->lanes_nccl) { return NB200_OK; }
		std::vector<int> devs;
		for(auto& l : ctx->lanes)
		{
			if(std::find(devs.begin(), devs.end(), l.dev) != devs.end())
			{
				return fail(ctx, NB200_ERR_UNSUPPORTED, "use_nccl: the device list repeats device %d (NCCL needs distinct devices); peer copies stay in use", l.dev);
			}
			devs.push_back(l.dev);
		}
		if(ctx->nccl == nullptr) { ctx->nccl = nccl_load(ctx->err); }
		if(ctx->nccl == nullptr) { return NB200_ERR_NCCL; }
		std::vector<ncclComm_t> comms(devs.size(), nullptr);
		NC(ctx, ctx->nccl->CommInitAll(comms.data(), static_cast<int>(devs.size()), devs.data()));
		for(size_t i = 0; i < ctx->lanes.size(); ++i) { ctx->lanes[i].lane_comm = comms[i]; }
		ctx->lanes_nccl = true;
	}
	else if(strcmp(name, "direct_targets_per_thread") == 0) { ctx->opt_direct_ipt = value; }
	else if(strcmp(name, "direct_segments") == 0) { ctx->opt_direct_segments = value; }
	else if(strcmp(name, "walk_mode") == 0) { ctx->opt_walk_mode = value; }	// 0 = automatic, 1 = thread per target, 2 / 4 = targets per lane, 32 = one per lane
	else if(strcmp(name, "walk_threads") == 0) { ctx->opt_walk_threads = value; }
	else if(strcmp(name, "walk_lpt") == 0) { ctx->opt_walk_lpt = value; }	// -1 automatic, 0 off, 1 on
	else if(strcmp(name, "direct_symmetric") == 0) { ctx->opt_direct_sym = value; ctx->sym_unavailable = false; }	// -1 auto, 0 off, 1 on
	else if(strcmp(name, "direct_small") == 0) { ctx->opt_direct_small = value; }	// -1 auto (N <= 4096), 0 off, 1 on
	else if(strcmp(name, "direct_sym_tile") == 0) { ctx->opt_sym_tile = value; }
	else if(strcmp(name, "direct_sym_shape") == 0) { ctx->opt_sym_shape = value; }
	else if(strcmp(name, "timing") == 0) { ctx->opt_timing = value; }
	else { return fail(ctx, NB200_ERR_ARG, "set_option: unknown option %s", name); }
	return NB200_OK;
}
