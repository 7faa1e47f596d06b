#ifndef NBODY_ENGINE_B200_H
#define NBODY_ENGINE_B200_H

// nbody_engine_b200 -- drop-in engine for drons/nbody backed by the nb200 C ABI (include/nb200.h).
//
// Sits next to nbody_engine_cuda / nbody_engine_cuda_bh_tex (nbody/nbody_engine_cuda.h,
// nbody/nbody_engine_cuda_bh_tex.h) and implements the whole nbody_engine virtual API
// (nbody/nbody_engine.h:16-96), so every solver drives it unchanged. This file is compiled against
// the reference's own headers (it is the binding a reference maintainer would add, see
// INTEGRATION.md); it contains no kernels and no arithmetic -- every operation is one C-ABI call.

#include "nbody_engine.h"
#include "nbody_engine_simple_bh.h"	// e_tree_layout, tree_layout_name / tree_layout_from_str

struct nb200_ctx;
struct nb200_buf;

class NBODY_DLL nbody_engine_b200 : public nbody_engine
{
	nbody_engine_b200(const nbody_engine_b200&) = delete;
	nbody_engine_b200& operator = (const nbody_engine_b200&) = delete;
public:
	class smemory;
	enum e_force { ef_direct, ef_barnes_hut };

	explicit nbody_engine_b200(e_force force = ef_direct,
							   nbcoord_t distance_to_node_radius_ratio = 10,
							   size_t tree_build_rate = 0,
							   e_tree_layout tl = etl_heap_stackless);
	~nbody_engine_b200();
	const char* type_name() const override;

	bool init(nbody_data* data) override;
	void get_data(nbody_data* data) override;
	size_t problem_size() const override;
	memory* get_y() override;
	void advise_time(const nbcoord_t& dt) override;
	nbcoord_t get_time() const override;
	void set_time(nbcoord_t t) override;
	size_t get_step() const override;
	void set_step(size_t s) override;

	void fcompute(const nbcoord_t& t, const memory* y, memory* f) override;
	void clamp(memory* y, nbcoord_t b) override;

	memory* create_buffer(size_t) override;
	void free_buffer(memory*) override;
	void read_buffer(void* dst, const memory* src) override;
	void write_buffer(memory* dst, const void* src) override;
	void copy_buffer(memory* a, const memory* b) override;
	void fill_buffer(memory* a, const nbcoord_t& value) override;

	void fmadd_inplace(memory* a, const memory* b, const nbcoord_t& c) override;
	void fmadd(memory* a, const memory* b, const memory* c, const nbcoord_t& d) override;
	//! Fused single-pass versions of the base class's term-by-term loops (nbody_engine.cpp:47-113)
	void fmaddn_inplace(memory* a, const memory_array& b, const nbcoord_t* c, size_t csize) override;
	void fmaddn_corr(memory* a, memory* corr, const memory_array& b, const nbcoord_t* c, size_t csize) override;
	void fmaddn(memory* a, const memory* b, const memory_array& c, const nbcoord_t* d, size_t dsize) override;
	void fmaxabs(const memory* a, nbcoord_t& result) override;

	void print_info() const override;

	//! Same contract as nbody_engine_cuda::select_devices (nbody_engine_cuda.cpp:576-616): 0 on success
	int select_devices(const QString& devices_str);
	//! Kept for factory symmetry with the cuda engines; tile sizes are chosen by the library (print_info says so)
	void set_block_size(int block_size);
	//! Several devices in this process: exchange shards with NCCL (ncclCommInitAll, as nbody_engine_cuda.cpp:100-106)
	//! instead of the default peer copies / peer loads over NVLink
	void set_use_nccl(bool);
	//! Solver steps as CUDA graphs (factory parameter step_graph, default on): a step that repeats the previous one call for call
	//! is replayed as one cudaGraphLaunch from advise_time(); results are those of the eager engine (include/nb200.h)
	void set_step_graph(bool);
	//! {graphs launched, replays abandoned, state of the next step, kernels in the step replayed last, distinct steps}
	bool step_graph_stats(unsigned long long out[5]) const;
	//! Device-side conservation sums of nbody_data::print_statistics (nbody_data.cpp:57-103) for a state vector:
	//! out = {P[3], L[3], Ekin, Epot, mass centre[3]}; Epot (O(N^2), on the GPU) only when with_energy. False on error.
	bool statistics(const memory* y, bool with_energy, double out[11]);
	//! Kernels launched so far (instrumentation)
	unsigned long long launch_count() const;
	//! Block until all queued device work has finished
	void synchronize();
private:
	bool ensure_context();
	struct	data;
	data*	d;
};

//! Factory glue for nbody_create_engine (nbody/nbody_engines.cpp:4): aliases "b200" and "b200_bh".
//! Returns NULL for other aliases and for invalid parameters.
class QVariant;
NBODY_DLL nbody_engine* nbody_create_engine_b200(const QVariantMap& param);

#endif // NBODY_ENGINE_B200_H
