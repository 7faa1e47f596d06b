// nb200 -- solver steps as CUDA graphs, behind the unchanged nbody_engine call sequence.
//
// A fixed-step solver (euler, rk4, rk_butcher family with a fixed step, rkfeagin*, midpoint, Bulirsch-Stoer with a
// fixed level count) issues the SAME engine calls on the SAME buffers with the SAME coefficients every step and ends
// the step with advise_time() (e.g. nbody/nbody_solver_rk4.cpp:30-62). At C1/C2 sizes those 20..270 launches cost more
// host time than device time. With the option "step_graph" on, the library watches the call stream between two
// nb200_step_boundary() calls (the adapter calls it from advise_time):
//
//   record   step k runs eagerly; its calls are written down (operation, buffer handles, scalars)
//   capture  step k+1 is issued into a stream capture instead of the stream; at the boundary the capture becomes a
//            graph, is launched once (that IS step k+1) and, if the calls equalled the recorded ones, kept
//   replay   from step k+2 on each call is only compared with the recorded one and returns; the boundary launches the
//            graph: one cudaGraphLaunch per solver step
//
// fmaxabs -- the one host-visible call the error-controlled solvers make inside every step
// (nbody_solver_rk_butcher.cpp:207-215) -- is a SEGMENT BORDER: the calls deferred before it are launched as one graph,
// the reduction runs right away and returns its value, the calls after it form the next segment. If the solver then
// branches differently from the recorded step (it subdivides), the calls stop matching and replay ends as below.
//
// Anything else host-visible in the middle of a step (read_buffer, write_buffer, statistics) or any difference from
// the recorded step (another coefficient, another buffer, a create/free_buffer) ends the replay: the calls accepted
// since the last border are issued eagerly, in order, and the step carries on eagerly, so the results are exactly
// those of the eager engine. Between steps (nothing deferred) host-visible calls are harmless.
// Only single-shard contexts defer; with lanes or ranks the option is accepted and ignored.
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
	size_t							step = 0;	// fcompute_bh with tree_build_rate > 0

	bool same(const step_op& o) const
	{
		return kind == o.kind && a == o.a && b == o.b && c == o.c && step == o.step && list == o.list &&
			   coef.size() == o.coef.size() &&
			   (coef.empty() || memcmp(coef.data(), o.coef.data(), coef.size() * sizeof(real)) == 0);
	}
};

enum step_mode
{
	SG_OFF = 0,
	SG_RECORD,
	SG_CAPTURE,
	SG_REPLAY
};

struct step_graph
{
	int						mode = SG_OFF;
	std::vector<step_op>	seq;		// the recorded step
	std::vector<step_op>	cur;		// calls of the step in flight (record / capture)
	size_t					pos = 0;	// replay: calls of the current step accepted so far
	bool					clean = true;	// no host-visible call inside the step in flight
	bool					capturing = false;
	bool					busy = false;	// re-entrancy guard while deferred calls are issued
	long long				saved_timing = 0;
	int						failures = 0;	// captures that did not lead to a replay; gives up after a few
	int						replayed_in_a_row = 0;	// steps replayed since the last abandoned one (16 of them clear `failures`)
	// one graph per segment of the recorded step (segments are separated by fmaxabs calls; nullptr = empty segment)
	std::vector<cudaGraphExec_t>	execs;
	std::vector<unsigned long long>	seg_launches;	// kernel launches inside each segment
	size_t					seg = 0;		// replay: segment the accepted calls belong to
	size_t					seg_start = 0;	// replay: index in seq of the first call of that segment
	size_t					cur_seg_start = 0;	// capture: index in cur of the first call of the open segment
	unsigned long long		launches_at_begin = 0;
	unsigned long long		launches_per_step = 0;
	// counters for tests / measurement
	unsigned long long		graph_launches = 0;
	unsigned long long		bailouts = 0;
};

#define NB200_STEP_GRAPH_MAX_FAILURES 8

#endif // NB200_STEPGRAPH_CUH
