// nb200 -- the step table of nb200_stepgraph.cuh: record / capture / replay logic.
//
// Included by nb200_api.cu inside its anonymous namespace, after fail() and the CU() macro: the functions here re-issue
// recorded calls through the library's own entry points (include/nb200.h) and report errors the way those do.
// Entry points of the rest of the library:
//   step_note(ctx, op, &skip)   every deferrable call, after validating its arguments
//   step_border(ctx, op)        fmaxabs
//   step_break(ctx)             any other host-visible call
//   step_invalidate(ctx)        buffers or configuration changed
//   step_find_or_add / the boundary logic itself live in nb200_step_boundary (nb200_api.cu)
#ifndef NB200_STEPGRAPH_IMPL_CUH
#define NB200_STEPGRAPH_IMPL_CUH

int step_issue(nb200_ctx* ctx, const step_op& op)
{
	nb200_buf* a = const_cast<nb200_buf*>(op.a);
	switch(op.kind)
	{
	case SOP_FCOMPUTE_DIRECT: return nb200_fcompute_direct(ctx, op.a, const_cast<nb200_buf*>(op.b));
	case SOP_FCOMPUTE_BH: return nb200_fcompute_bh(ctx, op.a, const_cast<nb200_buf*>(op.b), op.step);
	case SOP_FMADD_INPLACE: return nb200_fmadd_inplace(ctx, a, op.b, op.coef[0]);
	case SOP_FMADD: return nb200_fmadd(ctx, a, op.b, op.c, op.coef[0]);
	case SOP_FMADDN_INPLACE: return nb200_fmaddn_inplace(ctx, a, op.list.data(), op.coef.data(), op.coef.size());
	case SOP_FMADDN: return nb200_fmaddn(ctx, a, op.b, op.list.data(), op.coef.data(), op.coef.size());
	case SOP_FMADDN_CORR: return nb200_fmaddn_corr(ctx, a, const_cast<nb200_buf*>(op.b), op.list.data(), op.coef.data(), op.coef.size());
	case SOP_COPY: return nb200_copy(ctx, a, op.b);
	case SOP_FILL: return nb200_fill(ctx, a, op.coef[0]);
	case SOP_CLAMP: return nb200_clamp(ctx, a, op.coef[0]);
	case SOP_FMAXABS: return NB200_OK;	// a segment border: ran when it was called
	default: return fail(ctx, NB200_ERR_STATE, "step graph: unknown recorded call %d", op.kind);
	}
}

void step_destroy_execs(nb200_ctx* ctx, std::vector<cudaGraphExec_t>& execs)
{
	bool synced = false;
	for(cudaGraphExec_t e : execs)
	{
		if(e == nullptr) { continue; }
		if(!synced)
		{
			cudaSetDevice(ctx->lanes[0].dev);
			cudaStreamSynchronize(ctx->lanes[0].stream);	// a launched graph may still be running
			synced = true;
		}
		cudaGraphExecDestroy(e);
	}
	execs.clear();
}

// Forget every step seen so far (and free the graphs)
void step_clear_table(nb200_ctx* ctx)
{
	step_graph& sg = *ctx->sg;
	for(step_entry& e : sg.entries) { step_destroy_execs(ctx, e.execs); }
	sg.entries.clear();
	sg.last = sg.target = -1;
}

// Mispredicted step. Callers whose steps never repeat are never predicted for and never get here; this only stops a
// caller whose steps repeat just often enough to be predicted and then always differ.
void step_miss(nb200_ctx* ctx)
{
	step_graph& sg = *ctx->sg;
	if(++sg.misses > NB200_STEP_GRAPH_MAX_MISSES)
	{
		step_clear_table(ctx);
		step_destroy_execs(ctx, sg.cap_execs);
		sg.cap_launches.clear();
		sg.cur.clear();
		sg.mode = SG_OFF;
	}
}

// Issue ops[first, last) eagerly, in order (the library's own entry points, with deferral switched off meanwhile).
int step_issue_range(nb200_ctx* ctx, const std::vector<step_op>& ops, size_t first, size_t last)
{
	step_graph& sg = *ctx->sg;
	sg.busy = true;
	int rc = NB200_OK;
	for(size_t k = first; k < last && rc == NB200_OK; ++k) { rc = step_issue(ctx, ops[k]); }
	sg.busy = false;
	return rc;
}

// Replay ends early: the calls accepted since the last segment border have not run yet. Run them, then carry on as a
// recording (earlier segments of this step already ran as graphs). The entry keeps its graphs for later steps.
int step_bail(nb200_ctx* ctx)
{
	step_graph& sg = *ctx->sg;
	const std::vector<step_op>& seq = sg.entries[static_cast<size_t>(sg.target)].seq;
	const size_t first = sg.seg_start, upto = sg.pos;
	++sg.bailouts;
	sg.mode = SG_RECORD;
	int rc = step_issue_range(ctx, seq, first, upto);
	sg.cur.assign(seq.begin(), seq.begin() + static_cast<std::ptrdiff_t>(upto));
	sg.pos = sg.seg = sg.seg_start = 0;
	sg.target = -1;
	step_miss(ctx);
	return rc;
}

int step_begin_segment(nb200_ctx* ctx)
{
	step_graph& sg = *ctx->sg;
	nb200_lane& l = ctx->lanes[0];
	CU(ctx, cudaSetDevice(l.dev));
	CU(ctx, cudaStreamBeginCapture(l.stream, cudaStreamCaptureModeRelaxed));
	sg.capturing = true;
	sg.saved_timing = ctx->opt_timing;
	ctx->opt_timing = 0;	// phase events cannot be read back from inside a graph
	sg.launches_at_begin = ctx->launches;
	sg.cur_seg_start = sg.cur.size();
	return NB200_OK;
}

// Close the segment being captured (if any), run it, and keep its graph as segment number cap_execs.size().
int step_close_segment(nb200_ctx* ctx)
{
	step_graph&	sg = *ctx->sg;
	if(!sg.capturing)
	{
		sg.cap_execs.push_back(nullptr);	// an empty segment (two borders in a row, or a border first)
		sg.cap_launches.push_back(0);
		return NB200_OK;
	}
	nb200_lane&	l = ctx->lanes[0];
	cudaGraph_t	graph = nullptr;
	sg.capturing = false;
	ctx->opt_timing = sg.saved_timing;
	cudaSetDevice(l.dev);
	cudaError_t		res = cudaStreamEndCapture(l.stream, &graph);
	cudaGraphExec_t	exec = nullptr;
	if(res == cudaSuccess) { res = cudaGraphInstantiate(&exec, graph, 0); }
	if(res == cudaSuccess) { res = cudaGraphLaunch(exec, l.stream); }
	if(graph != nullptr) { cudaGraphDestroy(graph); }
	if(res != cudaSuccess)
	{
		// nothing of this segment has run: run it eagerly and stop deferring for good
		cudaGetLastError();
		if(exec != nullptr) { cudaGraphExecDestroy(exec); }
		sg.mode = SG_OFF;
		step_clear_table(ctx);
		step_destroy_execs(ctx, sg.cap_execs);
		sg.cap_launches.clear();
		std::vector<step_op> lost;
		lost.swap(sg.cur);
		ctx->launches = sg.launches_at_begin;
		int rc = step_issue_range(ctx, lost, sg.cur_seg_start, lost.size());
		if(rc != NB200_OK) { return rc; }
		ctx->err = std::string("step graph: capture failed (") + cudaGetErrorString(res) + "), continuing eagerly";
		return NB200_OK;
	}
	sg.cap_execs.push_back(exec);
	sg.cap_launches.push_back(ctx->launches - sg.launches_at_begin);
	++sg.graph_launches;
	return NB200_OK;
}

// A host-visible or state-changing call: everything deferred so far must have been issued before it.
int step_break(nb200_ctx* ctx)
{
	step_graph& sg = *ctx->sg;
	if(sg.mode == SG_OFF || sg.busy) { return NB200_OK; }
	int rc = NB200_OK;
	if(sg.mode == SG_REPLAY)
	{
		if(sg.pos == 0) { return NB200_OK; }	// between steps: nothing is deferred
		rc = step_bail(ctx);
	}
	else if(sg.mode == SG_CAPTURE && !sg.cur.empty())
	{
		rc = step_close_segment(ctx);
		if(sg.mode != SG_OFF)
		{
			step_destroy_execs(ctx, sg.cap_execs);	// a step with a host-visible call inside is not kept
			sg.cap_launches.clear();
			sg.mode = SG_RECORD;
			sg.target = -1;
			step_miss(ctx);
		}
	}
	if(!sg.cur.empty()) { sg.clean = false; }
	return rc;
}

// Buffers or configuration changed: the recorded steps no longer describe what the caller will do.
int step_invalidate(nb200_ctx* ctx)
{
	step_graph& sg = *ctx->sg;
	if(sg.mode == SG_OFF || sg.busy) { return NB200_OK; }
	int rc = step_break(ctx);
	if(sg.mode == SG_OFF) { return rc; }
	step_clear_table(ctx);
	sg.pos = sg.seg = sg.seg_start = 0;
	sg.mode = SG_RECORD;
	return rc;
}

int step_overflow(nb200_ctx* ctx)
{
	// nobody marks step boundaries: stop watching
	step_graph& sg = *ctx->sg;
	int rc = sg.capturing ? step_close_segment(ctx) : NB200_OK;
	step_destroy_execs(ctx, sg.cap_execs);
	sg.cap_launches.clear();
	step_clear_table(ctx);
	sg.cur.clear();
	sg.mode = SG_OFF;
	return rc;
}

// The call at hand is not what the predicted entry has at this position. Another entry may agree with everything
// accepted so far AND with this call: then the prediction was wrong, not the idea of replaying.
//   * such an entry with graphs: carry on replaying against it (segments launched so far came from the old entry's
//     graphs, which hold the same calls) -- returns 1, the call is accepted;
//   * such an entry without graphs, and no segment border passed yet: capture it now -- the capture starts with the
//     calls accepted so far, issued into it -- returns 2, the caller issues the call (into the capture);
//   * otherwise returns 0: the replay is abandoned (step_bail).
int step_retarget(nb200_ctx* ctx, const step_op& op)
{
	step_graph& sg = *ctx->sg;
	const std::vector<step_op>& have = sg.entries[static_cast<size_t>(sg.target)].seq;
	int uncaptured = -1;
	for(size_t k = 0; k < sg.entries.size(); ++k)
	{
		const step_entry& e = sg.entries[k];
		if(static_cast<int>(k) == sg.target || e.seq.size() <= sg.pos || !e.seq[sg.pos].same(op)) { continue; }
		bool prefix = true;
		for(size_t q = 0; prefix && q < sg.pos; ++q) { prefix = e.seq[q].same(have[q]); }
		if(!prefix) { continue; }
		if(e.captured)
		{
			sg.target = static_cast<int>(k);
			return 1;
		}
		if(uncaptured < 0) { uncaptured = static_cast<int>(k); }
	}
	if(uncaptured < 0 || sg.seg != 0 || op.kind == SOP_FMAXABS) { return 0; }
	const size_t upto = sg.pos;
	sg.cur.clear();
	if(step_begin_segment(ctx) != NB200_OK) { return 0; }
	sg.mode = SG_CAPTURE;
	const int issued = step_issue_range(ctx, have, 0, upto);
	sg.cur.assign(have.begin(), have.begin() + static_cast<std::ptrdiff_t>(upto));
	sg.pos = sg.seg = sg.seg_start = 0;
	if(issued != NB200_OK)
	{
		// a call that was accepted before fails now: run what was captured and carry on eagerly (the error shows up
		// again when the caller's own call is issued)
		step_close_segment(ctx);
		step_destroy_execs(ctx, sg.cap_execs);
		sg.cap_launches.clear();
		if(sg.mode != SG_OFF) { sg.mode = SG_RECORD; }
		sg.target = -1;
		sg.clean = false;
		return 2;
	}
	sg.target = uncaptured;
	return 2;
}

// Every deferrable call reports itself here after validating its arguments. *skip: the call is part of the replayed
// step and must not be issued now.
int step_note(nb200_ctx* ctx, const step_op& op, bool* skip)
{
	step_graph& sg = *ctx->sg;
	*skip = false;
	if(sg.mode == SG_OFF || sg.busy) { return NB200_OK; }
	if(sg.mode == SG_REPLAY)
	{
		const std::vector<step_op>& seq = sg.entries[static_cast<size_t>(sg.target)].seq;
		if(sg.pos < seq.size() && seq[sg.pos].same(op))
		{
			++sg.pos;
			*skip = true;
			return NB200_OK;
		}
		const int other = step_retarget(ctx, op);
		if(other == 1)
		{
			++sg.pos;
			*skip = true;
			return NB200_OK;
		}
		if(other == 0)
		{
			int rc = step_bail(ctx);
			if(rc != NB200_OK || sg.mode == SG_OFF) { return rc; }
		}
	}
	if(sg.mode == SG_CAPTURE && !sg.capturing)
	{
		int rc = step_begin_segment(ctx);
		if(rc != NB200_OK) { return rc; }
	}
	sg.cur.push_back(op);
	return sg.cur.size() > 65536 ? step_overflow(ctx) : NB200_OK;
}

// fmaxabs: host-visible, but part of every step of the error-controlled solvers (nbody_solver_rk_butcher.cpp:207-215).
// It is a segment BORDER: what was deferred before it runs as one graph, the reduction itself runs right away, and
// the calls after it form the next segment -- provided the solver then takes the same branch as in the recorded step.
int step_border(nb200_ctx* ctx, const step_op& op)
{
	step_graph& sg = *ctx->sg;
	if(sg.mode == SG_OFF || sg.busy) { return NB200_OK; }
	if(sg.mode == SG_REPLAY)
	{
		const step_entry& e = sg.entries[static_cast<size_t>(sg.target)];
		if(sg.pos < e.seq.size() && e.seq[sg.pos].same(op))
		{
			nb200_lane& l = ctx->lanes[0];
			if(e.execs[sg.seg] != nullptr)
			{
				CU(ctx, cudaSetDevice(l.dev));
				CU(ctx, cudaGraphLaunch(e.execs[sg.seg], l.stream));
				ctx->launches += e.seg_launches[sg.seg];
				++sg.graph_launches;
			}
			++sg.seg;
			++sg.pos;
			sg.seg_start = sg.pos;
			return NB200_OK;
		}
		if(sg.pos == 0) { return NB200_OK; }	// a stray reduction between steps
		if(step_retarget(ctx, op) == 1)
		{
			// another entry has the border here too: the segment before it holds the same calls in both
			const step_entry& t = sg.entries[static_cast<size_t>(sg.target)];
			if(t.execs[sg.seg] != nullptr)
			{
				nb200_lane& l = ctx->lanes[0];
				CU(ctx, cudaSetDevice(l.dev));
				CU(ctx, cudaGraphLaunch(t.execs[sg.seg], l.stream));
				ctx->launches += t.seg_launches[sg.seg];
				++sg.graph_launches;
			}
			++sg.seg;
			++sg.pos;
			sg.seg_start = sg.pos;
			return NB200_OK;
		}
		int rc = step_bail(ctx);
		if(rc != NB200_OK || sg.mode == SG_OFF) { return rc; }
	}
	if(sg.mode == SG_CAPTURE)
	{
		int rc = step_close_segment(ctx);
		if(rc != NB200_OK || sg.mode == SG_OFF) { return rc; }
	}
	sg.cur.push_back(op);
	return sg.cur.size() > 65536 ? step_overflow(ctx) : NB200_OK;
}

// The table entry with exactly these calls (filed if new); -1 for an empty step
int step_find_or_add(nb200_ctx* ctx, std::vector<step_op>& calls)
{
	step_graph& sg = *ctx->sg;
	if(calls.empty()) { return -1; }
	for(size_t k = 0; k < sg.entries.size(); ++k)
	{
		if(step_same_sequence(sg.entries[k].seq, calls)) { return static_cast<int>(k); }
	}
	if(sg.entries.size() >= NB200_STEP_GRAPH_MAX_ENTRIES) { step_clear_table(ctx); }	// too many distinct steps: start over
	sg.entries.emplace_back();
	sg.entries.back().seq.swap(calls);
	return static_cast<int>(sg.entries.size()) - 1;
}

#endif // NB200_STEPGRAPH_IMPL_CUH
