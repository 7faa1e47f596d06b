// TEST INFRASTRUCTURE ONLY -- not part of the nb200 product.
//
// extern "C" handles over the UNMODIFIED drons/nbody classes (compiled from
// /root/reference by oracle/Makefile into oracle/_ref/libnbref_f{64,32}.so) so
// that tests/ and bench.py's cpu_baseline leg can drive them through ctypes:
//   * nbody_data generators/loaders        (nbody/nbody_data.cpp)
//   * any nbody_engine through its virtual API (nbody/nbody_engine.h:16-96)
//   * the reference solver factory + run loop (nbody/nbody_solvers.cpp:17,
//     nbody/nbody_solver.cpp:55-105)
//   * nbody_space_heap build + an instrumented stackless walk
//     (nbody/nbody_space_heap.cpp:13-37, nbody_space_heap_stackless.cpp:3-28)
// Every engine entry point takes an opaque nbody_engine*, so the same calls
// drive the reference's CPU engines and the nb200 adapter class
// (nbody_b200/host/nbody_engine_b200.cpp) -- that is how "every existing
// solver drives it unchanged" is tested.
#include <omp.h>
#include "nbody_engines.h"
#include "nbody_solvers.h"
#include "nbody_space_heap_stackless.h"
#include "summation.h"

#define NBREF_API extern "C" __attribute__((visibility("default")))

namespace {
QVariantMap parse_params(const char* s)
{
	// "key=value;key=value"  (';' because device lists contain ',')
	QVariantMap	m;
	QStringList	items(QString(s ? s : "").split(";", QString::SkipEmptyParts));
	for(int i = 0; i < items.size(); ++i)
	{
		int eq = items[i].indexOf("=");
		if(eq < 0) { continue; }
		m[items[i].mid(0, eq).trimmed()] = QVariant(items[i].mid(eq + 1).trimmed());
	}
	return m;
}
nbody_engine* E(void* e) { return static_cast<nbody_engine*>(e); }
nbody_engine::memory* M(void* m) { return static_cast<nbody_engine::memory*>(m); }
nbody_data* D(void* d) { return static_cast<nbody_data*>(d); }
}  // namespace

NBREF_API int nbref_coord_size() { return static_cast<int>(sizeof(nbcoord_t)); }
NBREF_API int nbref_max_threads() { return omp_get_max_threads(); }
NBREF_API void nbref_set_threads(int n) { omp_set_num_threads(n); }

// ---- nbody_data ------------------------------------------------------------
NBREF_API void* nbref_data_new() { return new nbody_data(); }
NBREF_API void nbref_data_free(void* d) { delete D(d); }
NBREF_API void nbref_data_make_universe(void* d, size_t stars, double sx, double sy, double sz)
{
	D(d)->make_universe(stars, static_cast<nbcoord_t>(sx), static_cast<nbcoord_t>(sy), static_cast<nbcoord_t>(sz));
}
NBREF_API int nbref_data_load_initial(void* d, const char* path, const char* type)
{
	return D(d)->load_initial(QString(path), QString(type)) ? 0 : -1;
}
NBREF_API int nbref_data_load(void* d, const char* path)
{
	return D(d)->load(QString(path)) ? 0 : -1;
}
NBREF_API int nbref_data_save(void* d, const char* path)
{
	return D(d)->save(QString(path)) ? 0 : -1;
}
NBREF_API size_t nbref_data_count(void* d) { return D(d)->get_count(); }
NBREF_API size_t nbref_data_box_size(void* d) { return D(d)->get_box_size(); }
NBREF_API double nbref_data_time(void* d) { return static_cast<double>(D(d)->get_time()); }
NBREF_API size_t nbref_data_step(void* d) { return D(d)->get_step(); }
//! SoA export: y = [rx|ry|rz|vx|vy|vz] (6N), mass (N) -- layout of nbody_engine_simple.cpp:34-39
NBREF_API void nbref_data_export(void* d, nbcoord_t* y, nbcoord_t* mass)
{
	size_t				n = D(d)->get_count();
	const nbvertex_t*	r = D(d)->get_vertites();
	const nbvertex_t*	v = D(d)->get_velosites();
	for(size_t i = 0; i != n; ++i)
	{
		y[i] = r[i].x; y[n + i] = r[i].y; y[2 * n + i] = r[i].z;
		y[3 * n + i] = v[i].x; y[4 * n + i] = v[i].y; y[5 * n + i] = v[i].z;
		if(mass) { mass[i] = D(d)->get_mass()[i]; }
	}
}
NBREF_API void nbref_data_import(void* d, size_t n, const nbcoord_t* y, const nbcoord_t* mass)
{
	D(d)->clear();
	for(size_t i = 0; i != n; ++i)
	{
		D(d)->add_body(nbvertex_t(y[i], y[n + i], y[2 * n + i]),
					   nbvertex_t(y[3 * n + i], y[4 * n + i], y[5 * n + i]),
					   mass[i], nbcolor_t(1, 1, 1, 1));
	}
}
NBREF_API int nbref_data_is_equal(void* a, void* b, double eps)
{
	return D(a)->is_equal(*D(b), static_cast<nbcoord_t>(eps)) ? 1 : 0;
}
NBREF_API void nbref_data_set_check_list(void* d, const char* list) { D(d)->set_check_list(QString(list)); }
//! print_statistics (nbody_data.cpp:57-145) then report {dP%, dL%, dE%, |Vcm|-numerator parts}
NBREF_API void nbref_data_statistics(void* d, void* engine, double* out8)
{
	D(d)->print_statistics(E(engine));
	out8[0] = static_cast<double>(D(d)->get_impulce_err());
	out8[1] = static_cast<double>(D(d)->get_impulce_moment_err());
	out8[2] = static_cast<double>(D(d)->get_energy_err());
	nbvertex_t dc(D(d)->get_last_mass_center() - D(d)->get_initial_mass_center());
	out8[3] = static_cast<double>(dc.length());
	out8[4] = static_cast<double>(D(d)->get_last_total_energy());
	out8[5] = static_cast<double>(D(d)->get_last_total_impulce().length());
	out8[6] = static_cast<double>(D(d)->get_last_total_impulce_moment().length());
	out8[7] = static_cast<double>(D(d)->get_initial_energy());
}

// ---- engines (any nbody_engine*) -------------------------------------------
NBREF_API void* nbref_engine_create(const char* params) { return nbody_create_engine(parse_params(params)); }
NBREF_API void nbref_engine_free(void* e) { delete E(e); }
NBREF_API const char* nbref_engine_type_name(void* e) { return E(e)->type_name(); }
NBREF_API int nbref_engine_init(void* e, void* d) { return E(e)->init(D(d)) ? 0 : -1; }
NBREF_API void nbref_engine_get_data(void* e, void* d) { E(e)->get_data(D(d)); }
NBREF_API size_t nbref_engine_problem_size(void* e) { return E(e)->problem_size(); }
NBREF_API void* nbref_engine_get_y(void* e) { return E(e)->get_y(); }
NBREF_API void nbref_engine_advise_time(void* e, double dt) { E(e)->advise_time(static_cast<nbcoord_t>(dt)); }
NBREF_API double nbref_engine_get_time(void* e) { return static_cast<double>(E(e)->get_time()); }
NBREF_API void nbref_engine_set_time(void* e, double t) { E(e)->set_time(static_cast<nbcoord_t>(t)); }
NBREF_API size_t nbref_engine_get_step(void* e) { return E(e)->get_step(); }
NBREF_API void nbref_engine_set_step(void* e, size_t s) { E(e)->set_step(s); }
NBREF_API size_t nbref_engine_compute_count(void* e) { return E(e)->get_compute_count(); }
NBREF_API void nbref_engine_print_info(void* e) { E(e)->print_info(); }
NBREF_API void nbref_engine_fcompute(void* e, double t, void* y, void* f)
{
	E(e)->fcompute(static_cast<nbcoord_t>(t), M(y), M(f));
}
NBREF_API void nbref_engine_clamp(void* e, void* y, double b) { E(e)->clamp(M(y), static_cast<nbcoord_t>(b)); }
NBREF_API void* nbref_engine_create_buffer(void* e, size_t bytes) { return E(e)->create_buffer(bytes); }
NBREF_API void nbref_engine_free_buffer(void* e, void* m) { E(e)->free_buffer(M(m)); }
NBREF_API size_t nbref_memory_size(void* m) { return M(m)->size(); }
NBREF_API void nbref_engine_read_buffer(void* e, void* dst, void* src) { E(e)->read_buffer(dst, M(src)); }
NBREF_API void nbref_engine_write_buffer(void* e, void* dst, const void* src) { E(e)->write_buffer(M(dst), src); }
NBREF_API void nbref_engine_copy_buffer(void* e, void* a, void* b) { E(e)->copy_buffer(M(a), M(b)); }
NBREF_API void nbref_engine_fill_buffer(void* e, void* a, double v) { E(e)->fill_buffer(M(a), static_cast<nbcoord_t>(v)); }
NBREF_API void nbref_engine_fmadd_inplace(void* e, void* a, void* b, double c)
{
	E(e)->fmadd_inplace(M(a), M(b), static_cast<nbcoord_t>(c));
}
NBREF_API void nbref_engine_fmadd(void* e, void* a, void* b, void* c, double d)
{
	E(e)->fmadd(M(a), M(b), M(c), static_cast<nbcoord_t>(d));
}
namespace {
nbody_engine::memory_array marray(void** b, size_t n)
{
	nbody_engine::memory_array	a;
	for(size_t i = 0; i != n; ++i) { a.push_back(M(b[i])); }
	return a;
}
}  // namespace
//! nb = number of buffers in the array, csize = number of coefficients used (csize > nb is the negative test)
NBREF_API void nbref_engine_fmaddn_inplace(void* e, void* a, void** b, size_t nb, const nbcoord_t* c, size_t csize)
{
	E(e)->fmaddn_inplace(M(a), marray(b, nb), c, csize);
}
NBREF_API void nbref_engine_fmaddn_corr(void* e, void* a, void* corr, void** b, size_t nb, const nbcoord_t* c, size_t csize)
{
	E(e)->fmaddn_corr(M(a), M(corr), marray(b, nb), c, csize);
}
NBREF_API void nbref_engine_fmaddn(void* e, void* a, void* b, void** c, size_t nc, const nbcoord_t* d, size_t dsize)
{
	E(e)->fmaddn(M(a), M(b), marray(c, nc), d, dsize);
}
NBREF_API void nbref_engine_fmaxabs(void* e, void* a, nbcoord_t* result) { E(e)->fmaxabs(M(a), *result); }

namespace {
class foreign_memory : public nbody_engine::memory
{
	size_t m_size;
public:
	explicit foreign_memory(size_t s) : m_size(s) {}
	size_t size() const override { return m_size; }
};
}  // namespace
//! A memory object no engine owns (negative-branch tests, test_nbody_engine.cpp:748-759)
NBREF_API void* nbref_foreign_memory_new(size_t bytes) { return new foreign_memory(bytes); }
NBREF_API void nbref_foreign_memory_free(void* m) { delete M(m); }

//! Seconds per fcompute(y -> scratch f), best of `reps` after one warm-up
NBREF_API double nbref_engine_time_fcompute(void* e, int reps)
{
	nbody_engine::memory*	f = E(e)->create_buffer(sizeof(nbcoord_t) * E(e)->problem_size());
	if(f == nullptr) { return -1; }
	double	best = 1e300;
	for(int r = 0; r <= reps; ++r)
	{
		double t0 = omp_get_wtime();
		E(e)->fcompute(0, E(e)->get_y(), f);
		double dt = omp_get_wtime() - t0;
		if(r > 0 || reps == 0) { best = std::min(best, dt); }
	}
	E(e)->free_buffer(f);
	return best;
}

// ---- solvers ---------------------------------------------------------------
NBREF_API void* nbref_solver_create(const char* params) { return nbody_create_solver(parse_params(params)); }
NBREF_API void nbref_solver_free(void* s) { delete static_cast<nbody_solver*>(s); }
NBREF_API const char* nbref_solver_type_name(void* s) { return static_cast<nbody_solver*>(s)->type_name(); }
NBREF_API void nbref_solver_set_engine(void* s, void* e) { static_cast<nbody_solver*>(s)->set_engine(E(e)); }
NBREF_API void nbref_solver_set_time_step(void* s, double mn, double mx)
{
	static_cast<nbody_solver*>(s)->set_time_step(static_cast<nbcoord_t>(mn), static_cast<nbcoord_t>(mx));
}
NBREF_API void nbref_solver_advise(void* s, double dt) { static_cast<nbody_solver*>(s)->advise(static_cast<nbcoord_t>(dt)); }
NBREF_API int nbref_solver_run(void* s, void* d, double max_time, double dump_dt, double check_dt)
{
	return static_cast<nbody_solver*>(s)->run(D(d), nullptr, static_cast<nbcoord_t>(max_time),
											  static_cast<nbcoord_t>(dump_dt), static_cast<nbcoord_t>(check_dt));
}
//! Σb1, Σb2 and max_i |Σ_j a_ij - c_i| / max|a_i| of a Butcher solver (test_nbody_solver.cpp:87-144); -1 if not Butcher
NBREF_API int nbref_solver_butcher_check(void* s, double* out3)
{
	nbody_solver_rk_butcher*	rk = dynamic_cast<nbody_solver_rk_butcher*>(static_cast<nbody_solver*>(s));
	if(rk == nullptr) { return -1; }
	const nbody_butcher_table*	t = rk->table();
	nbcoord_t	b1 = 0, b2 = 0, worst = 0;
	for(size_t i = 0; i != t->get_steps(); ++i) { b1 += t->get_b1()[i]; b2 += t->get_b2()[i]; }
	for(size_t i = 0; i != t->get_steps(); ++i)
	{
		nbcoord_t	amax = 0, sum = 0, corr = 0;
		size_t		jmax = (t->is_implicit() ? t->get_steps() : i);
		if(jmax == 0) { amax = 1; }
		for(size_t j = 0; j != jmax; ++j)
		{
			nbcoord_t a = t->get_a()[i][j];
			sum = summation_k(sum, a, corr);
			amax = std::max(amax, static_cast<nbcoord_t>(fabs(a)));
		}
		worst = std::max(worst, static_cast<nbcoord_t>(fabs(sum - t->get_c()[i]) / amax));
	}
	out3[0] = static_cast<double>(b1); out3[1] = static_cast<double>(b2); out3[2] = static_cast<double>(worst);
	return 0;
}

// ---- kd-heap (nbody_space_heap) --------------------------------------------
namespace {
struct heap_probe : public nbody_space_heap_stackless
{
	size_t tree_size() const { return m_mass_center.size(); }
	//! Same control flow as nbody_space_heap_stackless::traverse, counting node visits and accepted nodes
	void count(const nbvertex_t& v1, size_t& visits, size_t& inter) const
	{
		size_t	curr = NBODY_HEAP_ROOT_INDEX;
		size_t	ts = m_mass_center.size();
		do
		{
			++visits;
			const nbcoord_t d2((v1 - m_mass_center[curr]).norm());
			if(d2 > m_radius_sqr[curr]) { ++inter; curr = skip_idx(curr); }
			else { curr = next_up(curr, ts); }
		}
		while(curr != NBODY_HEAP_ROOT_INDEX);
	}
};
}  // namespace
NBREF_API void* nbref_heap_new() { return new heap_probe(); }
NBREF_API void nbref_heap_free(void* h) { delete static_cast<heap_probe*>(h); }
NBREF_API double nbref_heap_build(void* h, size_t n, const nbcoord_t* y, const nbcoord_t* mass, double ratio)
{
	double t0 = omp_get_wtime();
	static_cast<heap_probe*>(h)->build(n, y, y + n, y + 2 * n, mass, static_cast<nbcoord_t>(ratio));
	return omp_get_wtime() - t0;
}
NBREF_API void nbref_heap_rebuild(void* h, size_t n, const nbcoord_t* y, double ratio)
{
	static_cast<heap_probe*>(h)->rebuild(n, y, y + n, y + 2 * n, static_cast<nbcoord_t>(ratio));
}
//! xyzr[4*idx+{0,1,2,3}] = mass centre + radius_sqr, mass[idx], body_n[idx]; all of length 2N (slot 0 unused)
NBREF_API size_t nbref_heap_export(void* h, nbcoord_t* xyzr, nbcoord_t* mass, long long* body_n)
{
	heap_probe*	hp = static_cast<heap_probe*>(h);
	size_t		ts = hp->tree_size();
	for(size_t i = 0; i != ts; ++i)
	{
		if(xyzr)
		{
			xyzr[4 * i + 0] = hp->get_mass_center()[i].x;
			xyzr[4 * i + 1] = hp->get_mass_center()[i].y;
			xyzr[4 * i + 2] = hp->get_mass_center()[i].z;
			xyzr[4 * i + 3] = hp->get_radius_sqr()[i];
		}
		if(mass) { mass[i] = hp->get_mass()[i]; }
		if(body_n) { body_n[i] = static_cast<long long>(hp->get_body_n()[i]); }
	}
	return ts;
}
//! Total node visits / accepted interactions over every leaf target [first, last) (heap index - N)
NBREF_API void nbref_heap_walk_counts(void* h, size_t first, size_t last, size_t stride, unsigned long long* out2)
{
	heap_probe*			hp = static_cast<heap_probe*>(h);
	size_t				n = hp->tree_size() / 2;
	unsigned long long	visits = 0, inter = 0;
	#pragma omp parallel for schedule(dynamic, 64) reduction(+ : visits, inter)
	for(size_t leaf = first; leaf < last; leaf += stride)
	{
		size_t v = 0, k = 0;
		hp->count(hp->get_mass_center()[n + leaf], v, k);
		visits += v;
		inter += k;
	}
	out2[0] = visits;
	out2[1] = inter;
}
//! Known-answer surface for test_nbody_heap_func (test_nbody_engine.cpp:1071-1148)
NBREF_API size_t nbref_heap_func(int which, size_t idx, size_t tree_size)
{
	typedef nbody_heap_func<size_t> hf;
	switch(which)
	{
	case 0: return hf::left_idx(idx);
	case 1: return hf::rght_idx(idx);
	case 2: return hf::parent_idx(idx);
	case 3: return hf::left2right(idx);
	case 4: return hf::next_down(idx);
	case 5: return hf::skip_idx(idx);
	case 6: return hf::next_up(idx, tree_size);
	case 7: return hf::is_left(idx) ? 1 : 0;
	case 8: return hf::is_right(idx) ? 1 : 0;
	default: return 0;
	}
}
