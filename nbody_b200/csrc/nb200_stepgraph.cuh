// nb200 -- solver steps as CUDA graphs, behind the unchanged nbody_engine call sequence.
//
// A solver issues the SAME engine calls on the SAME buffers with the SAME coefficients step after step and ends every
// step with advise_time() (e.g. nbody/nbody_solver_rk4.cpp:30-62). At C1/C2 sizes those 8..270 launches cost more host
// time than device time. With the option "step_graph" on, the library watches the call stream between two
// nb200_step_boundary() calls (the adapter calls it from advise_time) and keeps a small table of the distinct steps it
// has seen -- one entry per call sequence (operation, buffer handles, scalars), each remembering which entry followed
// it last time:
//
//   record   a step nothing is predicted for runs eagerly; its calls are written down and filed in the table
//   capture  a step predicted to repeat an entry that has no graph yet is issued into a stream capture instead of the
//            stream; at the boundary the capture becomes a graph, is launched once (that IS the step) and kept
//   replay   a step predicted to repeat an entry that has a graph: each call is only compared with the recorded one
//            and returns; the boundary launches the graph -- one cudaGraphLaunch per solver step
//
// The prediction is "what followed the previous step's entry last time", so periodic patterns are learnt after one
// period: a fixed-step solver (one entry), Adams (its history buffers rotate with period `rank`), Bulirsch-Stoer (the
// inner midpoint solver ends a "step" per sub-step, with a different size on every level), Barnes-Hut with
// tree_build_rate > 0 (every rate-th step rebuilds the tree).
//
// fmaxabs -- the one host-visible call the error-controlled solvers make inside every step
// (nbody_solver_rk_butcher.cpp:207-215) -- is a SEGMENT BORDER: the calls deferred before it are launched as one graph,
// the reduction runs right away and returns its value, the calls after it form the next segment. If the solver then
// branches differently from the recorded step (it subdivides), the calls stop matching and replay ends as below.
//
// Anything else host-visible in the middle of a step (read_buffer, write_buffer, statistics) or any difference from
// the predicted step (another coefficient, another buffer) ends the replay: the calls accepted since the last border
// are issued eagerly, in order, and the step carries on eagerly, so the results are exactly those of the eager engine.
// Between steps (nothing deferred) host-visible calls are harmless. create/free_buffer, set_bodies and option changes
// empty the table. Only single-shard contexts defer; with lanes or ranks the option is accepted and ignored.
#ifndef NB200_STEPGRAPH_CUH
#define NB200_STEPGRAPH_CUH

#include "nb200_common.cuh"

enum step_op_kind
{
	SOP_FCOMPUTE_DIRECT = 1,
	SOP_FCOMPUTE_BH,
	SOP_FMADD_INPLACE,
	SOP_FMADD,
	SOP_FMADDN_INPLACE,
	SOP_FMADDN,
	SOP_FMADDN_CORR,
	SOP_COPY,
	SOP_FILL,
	SOP_CLAMP,
	SOP_FMAXABS		// segment border: runs when called, splits the step's graph in two
};

struct step_op
{
	int								kind = 0;
	const nb200_buf*				a = nullptr;
	const nb200_buf*				b = nullptr;
	const nb200_buf*				c = nullptr;
	std::vector<const nb200_buf*>	list;	// fmaddn* terms
	std::vector<real>				coef;	// scalars, bit-compared
	size_t							step = 0;	// fcompute_bh: the step number to re-issue the call with
	int								key = 0;	// fcompute_bh with tree_build_rate > 0: 1 = the call rebuilds the tree, 2 = it refreshes it

	bool same(const step_op& o) const
	{
		return kind == o.kind && a == o.a && b == o.b && c == o.c && key == o.key && list == o.list &&
			   coef.size() == o.coef.size() &&
			   (coef.empty() || memcmp(coef.data(), o.coef.data(), coef.size() * sizeof(real)) == 0);
	}
};

static inline bool step_same_sequence(const std::vector<step_op>& x, const std::vector<step_op>& y)
{
	if(x.size() != y.size()) { return false; }
	for(size_t k = 0; k < x.size(); ++k)
	{
		if(!x[k].same(y[k])) { return false; }
	}
	return true;
}

// One distinct step: its calls, its graphs (one per segment; segments are separated by fmaxabs calls, nullptr = empty
// segment) once it has been captured, and the entry that followed it last time.
struct step_entry
{
	std::vector<step_op>			seq;
	std::vector<cudaGraphExec_t>	execs;
	std::vector<unsigned long long>	seg_launches;	// kernel launches inside each segment
	bool							captured = false;
	int								next = -1;
};

enum step_mode
{
	SG_OFF = 0,
	SG_RECORD,
	SG_CAPTURE,
	SG_REPLAY
};

#define NB200_STEP_GRAPH_MAX_ENTRIES 64	// distinct steps remembered; a caller with more starts over
#define NB200_STEP_GRAPH_MAX_MISSES 64		// mispredicted steps in a row before the library stops predicting

struct step_graph
{
	int						mode = SG_OFF;
	std::vector<step_entry>	entries;
	int						last = -1;		// entry of the previous step (-1: unknown, or it had a host-visible call inside)
	int						target = -1;	// entry this step is compared with (replay) or captured for (capture)
	std::vector<step_op>	cur;			// calls of the step in flight (record / capture)
	size_t					pos = 0;		// replay: calls of the current step accepted so far
	size_t					seg = 0;		// replay: segment the accepted calls belong to
	size_t					seg_start = 0;	// replay: index in the entry of the first call of that segment
	size_t					cur_seg_start = 0;	// capture: index in cur of the first call of the open segment
	bool					clean = true;	// no host-visible call (other than borders) inside the step in flight
	bool					capturing = false;
	bool					busy = false;	// re-entrancy guard while deferred calls are issued
	long long				saved_timing = 0;
	int						misses = 0;		// mispredicted steps in a row
	std::vector<cudaGraphExec_t>	cap_execs;		// capture: graphs of the segments closed so far
	std::vector<unsigned long long>	cap_launches;
	unsigned long long		launches_at_begin = 0;
	unsigned long long		launches_per_step = 0;	// of the step replayed last
	// counters for tests / measurement
	unsigned long long		graph_launches = 0;
	unsigned long long		bailouts = 0;
};

#endif // NB200_STEPGRAPH_CUH
