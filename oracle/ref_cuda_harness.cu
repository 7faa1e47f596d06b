// TEST / MEASUREMENT INFRASTRUCTURE ONLY -- not part of the nb200 product.
//
// Drives the reference's OWN CUDA kernels (nbody/nbody_engine_cuda_impl.cu, compiled unmodified from
// /root/reference for sm_100a by oracle/Makefile into oracle/_ref/libnbref_cuda_f{64,32}.so) so that bench.py can
// print "the reference's kernel recompiled for B200" next to nb200's number, and the tests can compare results.
// The reference's host class (nbody_engine_cuda.cpp) needs Qt and is not used: this file only supplies
// cuda_check() (defined there) and sets up buffers/textures the way nbody_engine_cuda_memory.cpp:99-130 does.
#include <cstdio>
#include <cstring>
#include "nbody_engine_cuda_impl.h"

void cuda_check(const char* file, int line, const char* context_name, cudaError_t res)
{
	if(cudaSuccess != res)
	{
		fprintf(stderr, "cudaError %s:%d %s %s\n", file, line, context_name, cudaGetErrorString(res));
	}
}

#define API extern "C" __attribute__((visibility("default")))

namespace {
template<class T>
T* upload(const T* host, size_t count)
{
	T* dev = nullptr;
	if(cudaMalloc(&dev, count * sizeof(T)) != cudaSuccess) { return nullptr; }
	cudaMemcpy(dev, host, count * sizeof(T), cudaMemcpyHostToDevice);
	return dev;
}

cudaTextureObject_t linear_texture(void* ptr, size_t bytes, int vec_size)
{
	cudaResourceDesc	res_desc;
	memset(&res_desc, 0, sizeof(res_desc));
	res_desc.resType = cudaResourceTypeLinear;
	res_desc.res.linear.devPtr = ptr;
	res_desc.res.linear.sizeInBytes = bytes;
	if(sizeof(nbcoord_t) == 8)
	{
		// doubles are fetched as int2 / int4 words (nb1Dfetch<double>, nbody_engine_cuda_impl.cu:237-259)
		res_desc.res.linear.desc = vec_size == 4 ? cudaCreateChannelDesc<int4>() : cudaCreateChannelDesc<int2>();
	}
	else
	{
		res_desc.res.linear.desc = vec_size == 4 ? cudaCreateChannelDesc<float4>() : cudaCreateChannelDesc<float>();
	}
	cudaTextureDesc	tex_desc;
	memset(&tex_desc, 0, sizeof(tex_desc));
	tex_desc.readMode = cudaReadModeElementType;
	tex_desc.addressMode[0] = cudaAddressModeClamp;
	tex_desc.filterMode = cudaFilterModePoint;
	tex_desc.normalizedCoords = 0;
	cudaTextureObject_t tex = 0;
	cudaCreateTextureObject(&tex, &res_desc, &tex_desc, NULL);
	return tex;
}

template<class Launch>
double time_best(Launch launch, int reps)
{
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	launch();
	cudaDeviceSynchronize();
	float best = 1e30f;
	for(int r = 0; r < reps; ++r)
	{
		cudaEventRecord(e0);
		launch();
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms = 0;
		cudaEventElapsedTime(&ms, e0, e1);
		if(ms < best) { best = ms; }
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	return best;
}
}  // namespace

API int nbrefcu_coord_size() { return static_cast<int>(sizeof(nbcoord_t)); }

//! kfcompute + kfcompute_xyz (the `cuda` engine's fcompute, nbody_engine_cuda.cpp:229-240) on one device
API int nbrefcu_direct(const nbcoord_t* y, const nbcoord_t* mass, size_t n, int block_size, int reps, nbcoord_t* f, double* ms)
{
	nbcoord_t*	dy = upload(y, 6 * n);
	nbcoord_t*	dm = upload(mass, n);
	nbcoord_t*	df = nullptr;
	if(dy == nullptr || dm == nullptr || cudaMalloc(&df, 6 * n * sizeof(nbcoord_t)) != cudaSuccess) { return -1; }
	*ms = time_best([&]() {
		fcompute_block(0, dy, df, dm, n, n, block_size);
		fcompute_xyz(dy, df, n, n, block_size);
	}, reps);
	cudaMemcpy(f, df, 6 * n * sizeof(nbcoord_t), cudaMemcpyDeviceToHost);
	cudaFree(dy);
	cudaFree(dm);
	cudaFree(df);
	return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

//! kfcompute_heap_bh_stackless with the tree in texture objects (cuda_bh_tex --tree_layout=heap_stackless,
//! nbody_engine_cuda_bh_tex.cpp:165-193); the tree arrays come from the caller (2n nodes, slot 0 unused)
API int nbrefcu_bh_stackless(const nbcoord_t* y, size_t n, const nbcoord_t* xyzr, const nbcoord_t* node_mass, const int* body_n,
							 int block_size, int reps, nbcoord_t* f, double* ms)
{
	nbcoord_t*	dy = upload(y, 6 * n);
	nbcoord_t*	dx = upload(xyzr, 8 * n);
	nbcoord_t*	dm = upload(node_mass, 2 * n);
	int*		db = upload(body_n, 2 * n);
	nbcoord_t*	df = nullptr;
	if(!dy || !dx || !dm || !db || cudaMalloc(&df, 6 * n * sizeof(nbcoord_t)) != cudaSuccess) { return -1; }
	cudaMemset(df, 0, 6 * n * sizeof(nbcoord_t));
	cudaTextureObject_t	tx = linear_texture(dx, 8 * n * sizeof(nbcoord_t), 4);
	cudaTextureObject_t	tm = linear_texture(dm, 2 * n * sizeof(nbcoord_t), 1);
	*ms = time_best([&]() {
		fcompute_heap_bh_stackless(0, static_cast<int>(n), static_cast<int>(n), static_cast<int>(2 * n), dy, df, tx, tm, db, block_size);
	}, reps);
	cudaMemcpy(f, df, 6 * n * sizeof(nbcoord_t), cudaMemcpyDeviceToHost);
	cudaDestroyTextureObject(tx);
	cudaDestroyTextureObject(tm);
	cudaFree(dy);
	cudaFree(dx);
	cudaFree(dm);
	cudaFree(db);
	cudaFree(df);
	return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
