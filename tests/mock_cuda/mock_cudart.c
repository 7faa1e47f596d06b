/*
 * TEST INFRASTRUCTURE ONLY -- a stand-in for the handful of CUDA runtime entry points libnb200 calls, so that the
 * library's HOST logic (handles, shard layout of buffers, argument checks, the step table of nb200_stepgraph.cuh) can
 * be driven on a machine without a GPU:   LD_PRELOAD=mock_cudart.so python tests/mock_cuda/drive.py
 *
 * "Device" memory is host memory, copies and memsets are real, kernels are NOT executed (a launch is only counted), a
 * stream capture records its copies, memsets and kernel launches as graph nodes, and launching a graph performs the
 * recorded copies and memsets and counts the recorded kernels. Nothing here is shipped or loaded by the product:
 * without this preload nb200_create fails on a machine without a CUDA device.
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <cuda_runtime_api.h>

typedef struct node
{
	int		kind;	/* 0 kernel, 1 memcpy, 2 memset */
	void*	dst;
	const void*	src;
	size_t	bytes;
	int		value;
} node;

typedef struct graph
{
	node*	nodes;
	size_t	count, cap;
} graph;

static int					g_capturing = 0;
static graph*				g_open = NULL;
static unsigned long long	g_counters[8];	/* 0 eager kernel launches, 1 captured kernel launches, 2 graph launches,
											   3 kernels replayed by graph launches, 4 graphs instantiated,
											   5 copies performed, 6 captures begun, 7 live device allocations */

void mock_counters(unsigned long long out[8])
{
	memcpy(out, g_counters, sizeof(g_counters));
}

static void graph_add(int kind, void* dst, const void* src, size_t bytes, int value)
{
	graph* g = g_open;
	if(g->count == g->cap)
	{
		g->cap = g->cap ? 2 * g->cap : 64;
		g->nodes = (node*)realloc(g->nodes, g->cap * sizeof(node));
	}
	node n = {kind, dst, src, bytes, value};
	g->nodes[g->count++] = n;
}

/* ---- devices ---------------------------------------------------------------------------------------------------- */
cudaError_t cudaGetDeviceCount(int* count) { *count = 2; return cudaSuccess; }
cudaError_t cudaSetDevice(int dev) { (void)dev; return cudaSuccess; }
cudaError_t cudaGetDevice(int* dev) { *dev = 0; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(struct cudaDeviceProp* prop, int dev)
{
	(void)dev;
	memset(prop, 0, sizeof(*prop));
	snprintf(prop->name, sizeof(prop->name), "mock sm_100 (no GPU)");
	prop->major = 10;
	prop->minor = 0;
	prop->multiProcessorCount = 148;
	prop->totalGlobalMem = (size_t)180 << 30;
	prop->l2CacheSize = 126 << 20;
	return cudaSuccess;
}
cudaError_t cudaDeviceGetAttribute(int* value, enum cudaDeviceAttr attr, int dev) { (void)attr; (void)dev; *value = 148; return cudaSuccess; }
cudaError_t cudaDeviceCanAccessPeer(int* can, int a, int b) { (void)a; (void)b; *can = 1; return cudaSuccess; }
cudaError_t cudaDeviceEnablePeerAccess(int peer, unsigned flags) { (void)peer; (void)flags; return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaPeekAtLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { (void)e; return "mock CUDA runtime"; }

/* ---- streams, events ---------------------------------------------------------------------------------------------- */
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned flags) { (void)flags; *s = (cudaStream_t)malloc(8); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t s) { (void)s; return g_capturing ? cudaErrorStreamCaptureUnsupported : cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags) { (void)s; (void)e; (void)flags; return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)malloc(8); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned flags) { (void)flags; return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) { (void)e; (void)s; return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t e) { (void)e; return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { (void)a; (void)b; *ms = 0.5f; return cudaSuccess; }

/* ---- memory --------------------------------------------------------------------------------------------------------- */
cudaError_t cudaMalloc(void** p, size_t bytes)
{
	*p = calloc(1, bytes ? bytes : 1);
	++g_counters[7];
	return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFree(void* p) { if(p) { free(p); --g_counters[7]; } return cudaSuccess; }
cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned flags) { (void)flags; *p = calloc(1, bytes ? bytes : 1); return cudaSuccess; }
cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaHostRegister(void* p, size_t bytes, unsigned flags) { (void)p; (void)bytes; (void)flags; return cudaSuccess; }
cudaError_t cudaHostUnregister(void* p) { (void)p; return cudaSuccess; }

cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, enum cudaMemcpyKind kind, cudaStream_t s)
{
	(void)s;
	if(g_capturing && kind == cudaMemcpyDeviceToDevice) { graph_add(1, dst, src, bytes, 0); return cudaSuccess; }
	if(g_capturing) { return cudaErrorStreamCaptureUnsupported; }	/* host copies inside a capture: the library must not */
	memmove(dst, src, bytes);
	++g_counters[5];
	return cudaSuccess;
}
cudaError_t cudaMemcpyPeerAsync(void* dst, int ddev, const void* src, int sdev, size_t bytes, cudaStream_t s)
{
	(void)ddev; (void)sdev;
	return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s);
}
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
							  enum cudaMemcpyKind kind, cudaStream_t s)
{
	(void)kind; (void)s;
	if(g_capturing) { return cudaErrorStreamCaptureUnsupported; }
	for(size_t r = 0; r < height; ++r) { memmove((char*)dst + r * dpitch, (const char*)src + r * spitch, width); }
	++g_counters[5];
	return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void* p, int value, size_t bytes, cudaStream_t s)
{
	(void)s;
	if(g_capturing) { graph_add(2, p, NULL, bytes, value); return cudaSuccess; }
	memset(p, value, bytes);
	return cudaSuccess;
}

/* ---- kernels -------------------------------------------------------------------------------------------------------- */
static struct { dim3 grid, block; size_t smem; void* stream; } g_cfg;
unsigned __cudaPushCallConfiguration(dim3 grid, dim3 block, size_t smem, void* stream)
{
	g_cfg.grid = grid; g_cfg.block = block; g_cfg.smem = smem; g_cfg.stream = stream;
	return 0;
}
cudaError_t __cudaPopCallConfiguration(dim3* grid, dim3* block, size_t* smem, void* stream)
{
	*grid = g_cfg.grid; *block = g_cfg.block; *smem = g_cfg.smem; *(void**)stream = g_cfg.stream;
	return cudaSuccess;
}
void** __cudaRegisterFatBinary(void* bin) { (void)bin; return (void**)calloc(1, sizeof(void*)); }
void __cudaRegisterFatBinaryEnd(void** handle) { (void)handle; }
void __cudaUnregisterFatBinary(void** handle) { free(handle); }
void __cudaRegisterFunction(void** handle, const char* host, char* dev, const char* name, int limit, void* tid, void* bid, void* bdim, void* gdim, int* wsize)
{
	(void)handle; (void)host; (void)dev; (void)name; (void)limit; (void)tid; (void)bid; (void)bdim; (void)gdim; (void)wsize;
}
void __cudaRegisterVar(void** handle, char* host, char* dev, const char* name, int ext, size_t size, int constant, int global)
{
	(void)handle; (void)host; (void)dev; (void)name; (void)ext; (void)size; (void)constant; (void)global;
}
cudaError_t cudaLaunchKernel(const void* func, dim3 grid, dim3 block, void** args, size_t smem, cudaStream_t s)
{
	(void)func; (void)args; (void)smem; (void)s;
	if(grid.x == 0 || block.x == 0) { return cudaErrorInvalidConfiguration; }
	if(g_capturing) { graph_add(0, NULL, NULL, 0, 0); ++g_counters[1]; }
	else { ++g_counters[0]; }
	return cudaSuccess;
}
cudaError_t cudaFuncSetAttribute(const void* func, enum cudaFuncAttribute attr, int value) { (void)func; (void)attr; (void)value; return cudaSuccess; }
cudaError_t cudaFuncGetAttributes(struct cudaFuncAttributes* a, const void* func) { (void)func; memset(a, 0, sizeof(*a)); a->ptxVersion = 100; a->binaryVersion = 100; a->maxThreadsPerBlock = 1024; return cudaSuccess; }
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(int* n, const void* func, int block, size_t smem, unsigned flags)
{
	(void)func; (void)block; (void)smem; (void)flags; *n = 2; return cudaSuccess;
}

/* ---- graphs --------------------------------------------------------------------------------------------------------- */
cudaError_t cudaStreamBeginCapture(cudaStream_t s, enum cudaStreamCaptureMode mode)
{
	(void)s; (void)mode;
	if(g_capturing) { return cudaErrorIllegalState; }
	g_capturing = 1;
	g_open = (graph*)calloc(1, sizeof(graph));
	++g_counters[6];
	return cudaSuccess;
}
cudaError_t cudaStreamEndCapture(cudaStream_t s, cudaGraph_t* out)
{
	(void)s;
	if(!g_capturing) { return cudaErrorIllegalState; }
	g_capturing = 0;
	*out = (cudaGraph_t)g_open;
	g_open = NULL;
	return cudaSuccess;
}
cudaError_t cudaGraphInstantiate(cudaGraphExec_t* exec, cudaGraph_t g, unsigned long long flags)
{
	(void)flags;
	graph* src = (graph*)g;
	graph* copy = (graph*)calloc(1, sizeof(graph));
	copy->count = copy->cap = src->count;
	copy->nodes = (node*)malloc((src->count ? src->count : 1) * sizeof(node));
	memcpy(copy->nodes, src->nodes, src->count * sizeof(node));
	*exec = (cudaGraphExec_t)copy;
	++g_counters[4];
	return cudaSuccess;
}
cudaError_t cudaGraphLaunch(cudaGraphExec_t exec, cudaStream_t s)
{
	(void)s;
	graph* g = (graph*)exec;
	if(g_capturing) { return cudaErrorStreamCaptureUnsupported; }
	for(size_t k = 0; k < g->count; ++k)
	{
		const node* n = &g->nodes[k];
		if(n->kind == 0) { ++g_counters[3]; }
		else if(n->kind == 1) { memmove(n->dst, n->src, n->bytes); ++g_counters[5]; }
		else { memset(n->dst, n->value, n->bytes); }
	}
	++g_counters[2];
	return cudaSuccess;
}
cudaError_t cudaGraphDestroy(cudaGraph_t g) { graph* p = (graph*)g; if(p) { free(p->nodes); free(p); } return cudaSuccess; }
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e) { return cudaGraphDestroy((cudaGraph_t)e); }
