/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the drons/nbody hot path.
 *
 * Plain-C restatement of the reference algorithms that the nb200 CUDA engine
 * replaces; each function cites the reference lines it follows. It is the
 * checker for tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
 * and is never linked, imported or executed by the product path
 * (nbody_b200/). Parity is PINNED: tests/test_oracle.py checks this file
 * against the reference's golden vectors (test/data/euler.txt, rk4.txt,
 * initial_state.txt; heap-function known answers of test_nbody_engine.cpp:
 * 1107-1147) and against the reference's own classes compiled into
 * oracle/_ref/libnbref_*.so.
 *
 * Built twice: liboracle_f64.so (orc_real = double), liboracle_f32.so (float).
 */
#ifndef NBODY_ORACLE_H
#define NBODY_ORACLE_H

#include <stddef.h>

#ifndef ORC_PRECISION
#define ORC_PRECISION 2
#endif
#if ORC_PRECISION == 1
typedef float orc_real;
#else
typedef double orc_real;
#endif

#ifdef __cplusplus
extern "C" {
#endif

int orc_real_size(void);
int orc_max_threads(void);
void orc_set_threads(int n);

/* ---- direct summation --------------------------------------------------- */
/* nbody_engine_openmp::fcompute (nbody_engine_openmp.cpp:19-77) with
 * nbody_data::force (nbody_data.cpp:35-44): f = (v, sum_j force(i,j) / m_i), j != i. */
void orc_fcompute_openmp(size_t n, const orc_real* y, const orc_real* mass, orc_real* f);
/* nbody_engine_block::fcompute (nbody_engine_block.cpp:51-125): 64x64 blocks,
 * coefficient m_j / (r * r2), j == i included. n must be a multiple of 64. */
void orc_fcompute_block(size_t n, const orc_real* y, const orc_real* mass, orc_real* f);
/* Same arithmetic as orc_fcompute_openmp for a subset of targets; acc is 3 x nt
 * ([ax(nt) | ay | az]). Lets full-size configurations be spot-checked in seconds. */
void orc_accel_subset(size_t n, const orc_real* y, const orc_real* mass,
					  const size_t* targets, size_t nt, orc_real* acc);
/* Same, accumulated in long double (80-bit) to attribute residuals. */
void orc_accel_subset_ld(size_t n, const orc_real* y, const orc_real* mass,
						 const size_t* targets, size_t nt, orc_real* acc);

/* ---- state-vector ops (nbody_engine_openmp.cpp:79-297, nbody_engine.cpp:47-113) */
void orc_fmadd_inplace(orc_real* a, const orc_real* b, orc_real c, size_t count);
void orc_fmadd(orc_real* a, const orc_real* b, const orc_real* c, orc_real d, size_t count);
void orc_fmaddn_inplace(orc_real* a, const orc_real* const* b, const orc_real* c, size_t csize, size_t count);
/* b may be NULL (start from zero); returns without touching a when b != NULL and every d[k] == 0 */
void orc_fmaddn(orc_real* a, const orc_real* b, const orc_real* const* c, const orc_real* d, size_t dsize, size_t count);
void orc_fmaddn_corr(orc_real* a, orc_real* corr, const orc_real* const* b, const orc_real* c, size_t csize, size_t count);
orc_real orc_fmaxabs(const orc_real* a, size_t count);
void orc_clamp(orc_real* y, orc_real b, size_t n);

/* ---- kd-heap Barnes-Hut (nbody_space_heap*.cpp, nbody_space_heap_func_priv.h) */
size_t orc_heap_left(size_t idx);
size_t orc_heap_right(size_t idx);
size_t orc_heap_parent(size_t idx);
size_t orc_heap_next_down(size_t idx);
size_t orc_heap_skip(size_t idx);
size_t orc_heap_next_up(size_t idx, size_t tree_size);
/* Arrays are 2n long (slot 0 unused): xyzr = 2n x 4 {cm.x, cm.y, cm.z, radius_sqr}. n must be 2^k. */
int orc_heap_build(size_t n, const orc_real* y, const orc_real* mass, orc_real ratio,
				   orc_real* xyzr, orc_real* node_mass, orc_real* bmin, orc_real* bmax, long long* body_n);
/* nbody_space_heap::rebuild (:39-62): keep body_n, refresh geometry. */
void orc_heap_rebuild(size_t n, const orc_real* y, orc_real ratio,
					  orc_real* xyzr, const orc_real* node_mass, orc_real* bmin, orc_real* bmax, const long long* body_n);
/* space_subdivided_fcompute with ett_nested_tree (nbody_engine_simple_bh.cpp:15-93):
 * stackless != 0 -> nbody_space_heap_stackless::traverse, else the stack walk. */
void orc_fcompute_bh(size_t n, const orc_real* y, const orc_real* mass,
					 const orc_real* xyzr, const orc_real* node_mass, const long long* body_n,
					 int stackless, orc_real* f, unsigned long long* visits, unsigned long long* interactions);

/* Stackless walk for a subset of leaves (leaf = heap index - n); acc is 3 x nt accelerations (already divided by the
 * target's mass, like update_f). Lets the N = 4M configuration be spot-checked in seconds. */
void orc_bh_subset(size_t n, const orc_real* xyzr, const orc_real* node_mass, const size_t* leaves, size_t nt, orc_real* acc);

/* ---- two solvers restated to pin the oracle on the golden files ------------
 * nbody_solver_euler.cpp, nbody_solver_rk4.cpp:19-48 driven like
 * nbody_solver::run (nbody_solver.cpp:73-76): while(t < max_time) advise(dt). */
void orc_run_euler(size_t n, orc_real* y, const orc_real* mass, orc_real dt, orc_real max_time);
void orc_run_rk4(size_t n, orc_real* y, const orc_real* mass, orc_real dt, orc_real max_time);

/* ---- conservation sums of nbody_data::print_statistics (nbody_data.cpp:57-103,
 * summation_proxy.h): out = {Px,Py,Pz, Lx,Ly,Lz, Ekin, Epot, Cx,Cy,Cz}. Kahan sums. */
void orc_statistics(size_t n, const orc_real* y, const orc_real* mass, int with_energy, orc_real* out11);

#ifdef __cplusplus
}
#endif
#endif
