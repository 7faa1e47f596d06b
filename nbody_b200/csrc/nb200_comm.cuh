// nb200 -- NCCL, bound at run time.
//
// The library has no link-time NCCL dependency: libnccl.so.2 is dlopen'ed when a
// context with nranks > 1 is created, or when a context with several lanes is asked
// to exchange shards with NCCL (option "use_nccl", the reference's use_nccl=1:
// ncclCommInitAll over the device list, nbody_engine_cuda.cpp:100-106). Inside a torchrun rank the name resolves to
// the NCCL build torch already loaded; under a plain C++ host it resolves to the
// system library. Only the stable v2 entry points are used.
#ifndef NB200_COMM_CUH
#define NB200_COMM_CUH

#include <dlfcn.h>
#include <nccl.h>
#include "nb200_common.cuh"

struct nccl_api
{
	void*	handle = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static nccl_api* nccl_load(std::string& err)
{
	static nccl_api	api;
	if(api.handle != nullptr)
	{
		return &api;
	}
	const char*	names[] = {"libnccl.so.2", "libnccl.so"};
	void*		h = nullptr;
	for(const char* nm : names)
	{
		h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
		if(h != nullptr) { break; }
	}
	if(h == nullptr)
	{
		err = std::string("cannot load NCCL: ") + dlerror();
		return nullptr;
	}
	api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
	api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
	api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
	api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(dlsym(h, "ncclCommInitAll"));
	api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(dlsym(h, "ncclGroupStart"));
	api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(dlsym(h, "ncclGroupEnd"));
	api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
	api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(h, "ncclAllReduce"));
	api.ReduceScatter = reinterpret_cast<decltype(api.ReduceScatter)>(dlsym(h, "ncclReduceScatter"));
	api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
	if(!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.CommInitAll || !api.GroupStart || !api.GroupEnd || !api.AllGather || !api.AllReduce || !api.ReduceScatter || !api.GetErrorString)
	{
		err = "NCCL library lacks a required v2 entry point";
		dlclose(h);
		return nullptr;
	}
	api.handle = h;
	return &api;
}

#if NB200_PRECISION == 2
#define NB200_NCCL_REAL ncclFloat64
#else
#define NB200_NCCL_REAL ncclFloat32
#endif

#endif // NB200_COMM_CUH
