/*
 * nb200 -- C ABI of the B200-native compute engine for drons/nbody.
 *
 * This is the drop-in boundary: everything the reference's nbody_engine virtual
 * API (nbody/nbody_engine.h:16-96) needs from a device back end, as plain C
 * (opaque handles, raw pointers, sizes). The C++ adapter class
 * nbody_engine_b200 (nbody_b200/host/nbody_engine_b200.cpp) and the Python host
 * mirror (nbody_b200/engine.py) are thin layers over exactly these calls; see
 * INTEGRATION.md for the binding a reference maintainer would add.
 *
 * Two builds of the same sources mirror the reference's compile-time precision
 * switch NB_COORD_PRECISION (nbody/nbtype.h:15-26):
 *     libnb200_f64.so   nb200_real = double   (reference default)
 *     libnb200_f32.so   nb200_real = float
 * Both export the same symbol names.
 *
 * Conventions
 *   - every call returns 0 on success, a negative nb200_status otherwise;
 *     nb200_last_error(ctx) holds a human-readable reason. Nothing throws, nothing
 *     calls exit() (the reference's CUDA engine exits on CUDA errors,
 *     nbody/nbody_engine_cuda.cpp:8-15; a library must not).
 *   - there is NO CPU fallback: without a usable CUDA device nb200_create fails.
 *   - a context is driven by ONE host thread; device work is stream-ordered
 *     and asynchronous except where a host-visible result is returned
 *     (nb200_read, nb200_fmaxabs, nb200_sync).
 *   - state vectors use the reference layout [rx|ry|rz|vx|vy|vz], each row N
 *     long (nbody/nbody_engine_simple.cpp:34-39).
 *
 * Sharding (multi-GPU). A context owns `nlanes` devices of this process and is
 * rank `rank` of `nranks` processes; total shards G = nlanes * nranks (one of
 * the two factors must be 1). A buffer whose size is exactly 6*N*sizeof(real)
 * (a state vector) is BODY-SHARDED: shard g holds columns [g*N/G, (g+1)*N/G)
 * of every row (a buffer of that size created before nb200_set_bodies becomes
 * one at that call, contents preserved). Any other buffer is replicated on every shard. Host-facing
 * calls (nb200_write / nb200_read) always take and return the full logical
 * buffer, so callers never see the sharding. With nranks > 1 every rank makes
 * the same sequence of calls (SPMD); fcompute all-gathers packed source bodies
 * with NCCL, fmaxabs all-reduces one scalar.
 */
#ifndef NB200_H
#define NB200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef NB200_PRECISION
#define NB200_PRECISION 2
#endif
#if NB200_PRECISION == 1
typedef float nb200_real;
#else
typedef double nb200_real;
#endif

typedef struct nb200_ctx nb200_ctx;
typedef struct nb200_buf nb200_buf;

typedef enum nb200_status {
	NB200_OK = 0,
	NB200_ERR_ARG = -1,       /* NULL / foreign handle / size mismatch: call ignored */
	NB200_ERR_CUDA = -2,      /* CUDA runtime error (message in nb200_last_error) */
	NB200_ERR_NCCL = -3,      /* NCCL error or NCCL library not loadable */
	NB200_ERR_ALLOC = -4,     /* device allocation failed */
	NB200_ERR_STATE = -5,     /* call made before nb200_set_bodies etc. */
	NB200_ERR_UNSUPPORTED = -6
} nb200_status;

/* Barnes-Hut tree layouts accepted by the reference factory for the CUDA
 * engines (nbody/nbody_engines.cpp:64-71); both share one node layout and
 * produce identical results, they differ in how the walk keeps its position. */
typedef enum nb200_tree_layout {
	NB200_TREE_HEAP = 1,            /* explicit stack walk, nbody_space_heap.cpp:64-97 */
	NB200_TREE_HEAP_STACKLESS = 2   /* skip-pointer walk, nbody_space_heap_stackless.cpp:3-28 */
} nb200_tree_layout;

/* ---- library / device queries ------------------------------------------- */
/* sizeof(nb200_real) of this build: 8 or 4. */
int nb200_real_size(void);
/* Number of CUDA devices, as select_devices() needs it (nbody_engine_cuda.cpp:585-590). */
int nb200_device_count(int* count);
/* 128-byte NCCL unique id created on rank 0 and handed to every rank's nb200_create. */
#define NB200_UID_BYTES 128
int nb200_comm_unique_id(void* uid128);

/* ---- context ------------------------------------------------------------- */
/* Replaces nbody_engine_cuda ctor + select_devices + init's stream/NCCL setup
 * (nbody_engine_cuda.cpp:63-106,576-616). dev_ids may repeat a device ("0,0"),
 * as the reference's tests do (test_nbody_engine.cpp:1259-1267).
 * uid128 is NULL when nranks == 1. */
int nb200_create(nb200_ctx** ctx, const int* dev_ids, int nlanes,
				 int rank, int nranks, const void* uid128);
int nb200_destroy(nb200_ctx* ctx);
const char* nb200_last_error(const nb200_ctx* ctx);
/* Block until every lane's stream is idle. */
int nb200_sync(nb200_ctx* ctx);
/* Total shard count G and this context's first shard index. */
int nb200_shards(const nb200_ctx* ctx, int* nshards, int* first_shard);
/* Human-readable device summary (print_info, nbody_engine_cuda.cpp:532-559). */
int nb200_describe(const nb200_ctx* ctx, char* text, size_t text_bytes);

/* Body count and masses (full N on every rank). Must precede any buffer that
 * should be recognised as a state vector. Replaces the mass upload in
 * nbody_engine_cuda::init (nbody_engine_cuda.cpp:107,138). N must be a
 * multiple of the shard count. */
int nb200_set_bodies(nb200_ctx* ctx, size_t n, const nb200_real* mass);
/* Read the masses back (full N). */
int nb200_get_mass(nb200_ctx* ctx, nb200_real* mass);

/* ---- buffers: create/free/read/write/copy/fill_buffer ---------------------
 * (nbody_engine_cuda.cpp:267-374, nbody_engine_cuda_memory.cpp:4-36) */
int nb200_alloc(nb200_ctx* ctx, size_t bytes, nb200_buf** buf);
int nb200_free(nb200_ctx* ctx, nb200_buf* buf);       /* NULL is a no-op */
size_t nb200_size(const nb200_buf* buf);              /* logical size in bytes */
int nb200_write(nb200_ctx* ctx, nb200_buf* dst, const void* host_src);
int nb200_read(nb200_ctx* ctx, void* host_dst, const nb200_buf* src);
/* Like nb200_read, but only the columns this process owns are written into the
 * full-layout host array (nranks > 1: no gather, D2H of the own shard only; the
 * other columns of host_dst are left untouched). With one rank it equals
 * nb200_read. For callers that keep per-rank output, e.g. a sharded dump. */
int nb200_read_local(nb200_ctx* ctx, void* host_dst, const nb200_buf* src);
int nb200_copy(nb200_ctx* ctx, nb200_buf* a, const nb200_buf* b);   /* sizes must match */
int nb200_fill(nb200_ctx* ctx, nb200_buf* a, nb200_real value);
/* Raw device pointer + element count of one lane's shard (zero-copy interop
 * for benchmarks: lets a harness place inputs in HBM without host copies). */
int nb200_lane_ptr(nb200_ctx* ctx, nb200_buf* buf, int lane, void** dptr, size_t* elems);

/* ---- f = f(t, y) ---------------------------------------------------------- */
/* Direct all-pairs right-hand side: f[0..3N) = y[3N..6N),
 * f[3N+..] = sum_j m_j (r_j - r_i) / max(|r_j - r_i|^2, 1e-8)^(3/2).
 * Replaces kfcompute + kfcompute_xyz + synchronize_y/f
 * (nbody_engine_cuda_impl.cu:10-124, nbody_engine_cuda.cpp:202-247). */
int nb200_fcompute_direct(nb200_ctx* ctx, const nb200_buf* y, nb200_buf* f);

/* Barnes-Hut configuration: opening ratio (distance_to_node_radius_ratio),
 * layout, and tree_build_rate exactly as the factory passes them
 * (nbody_engines.cpp:57-83). */
int nb200_bh_configure(nb200_ctx* ctx, nb200_real ratio, int layout, size_t tree_build_rate);
/* Barnes-Hut right-hand side. `step` is nbody_data::get_step(); the tree is
 * rebuilt when tree_build_rate == 0, no tree exists, or step % rate == 0, and
 * otherwise only its geometry is refreshed (nbody_engine_cuda_bh_tex.cpp:74-164).
 * N must be a power of two (the kd-heap's leaves are exactly [N, 2N),
 * nbody_space_heap.cpp:23). */
int nb200_fcompute_bh(nb200_ctx* ctx, const nb200_buf* y, nb200_buf* f, size_t step);
/* Copy the current tree of shard `lane` to the host for parity tests: xyzr is
 * 2N x 4 (mass centre + radius_sqr), mass 2N, body_n 2N; slot 0 unused. Any
 * pointer may be NULL. */
int nb200_bh_export_tree(nb200_ctx* ctx, int lane, nb200_real* xyzr, nb200_real* mass, int* body_n);
/* Node visits and accepted interactions of the last walk, summed over this
 * context's targets (only counted when enabled; costs two atomics per target). */
int nb200_bh_walk_stats(nb200_ctx* ctx, int enable, unsigned long long* visits,
						unsigned long long* interactions);
/* Profile of the last counted grouped walk (counting on, see nb200_bh_walk_stats; summed over lanes): out[0] rounds
 * (one work item per lane each), [1] work items (sibling pairs tested against a group of 32 targets), [2] interaction-list
 * entries (accepted node + target mask), [3] lane-items whose FP32 decisions were redone in FP64, [4] rounds summed with
 * the MinDistance clamp, [5] deepest fill of a warp's item stack, [6] sum over rounds of the busiest target's entries,
 * [7] sum over groups of the busiest target's entries over the whole walk, [8..12] entries whose mask names 32 / 24..31 /
 * 16..23 / 8..15 / 1..7 targets (last round of each group left out), [13..15] reserved. Instrumentation only. */
int nb200_bh_walk_profile(nb200_ctx* ctx, unsigned long long out[16]);

/* ---- state-vector ops ------------------------------------------------------
 * (nbody_engine_cuda.cpp:376-530, nbody_engine.cpp:47-113) */
/* a[i] += b[i]*c */
int nb200_fmadd_inplace(nb200_ctx* ctx, nb200_buf* a, const nb200_buf* b, nb200_real c);
/* a[i] = b[i] + c[i]*d   (a may alias b or c) */
int nb200_fmadd(nb200_ctx* ctx, nb200_buf* a, const nb200_buf* b, const nb200_buf* c, nb200_real d);
/* a[i] += sum_k b[k][i]*c[k], k in [0,n), terms applied in k order, zero c[k] skipped. One pass. */
int nb200_fmaddn_inplace(nb200_ctx* ctx, nb200_buf* a, const nb200_buf* const* b,
						 const nb200_real* c, size_t n);
/* a[i] = b[i] + sum_k c[k][i]*d[k]; b == NULL starts from 0. One pass. */
int nb200_fmaddn(nb200_ctx* ctx, nb200_buf* a, const nb200_buf* b, const nb200_buf* const* c,
				 const nb200_real* d, size_t n);
/* Kahan-compensated a[i] += sum_k b[k][i]*c[k] carrying corr[i] (summation.h:8-14). One pass. */
int nb200_fmaddn_corr(nb200_ctx* ctx, nb200_buf* a, nb200_buf* corr, const nb200_buf* const* b,
					  const nb200_real* c, size_t n);
/* *result = max_i |a[i]| (0 for an empty buffer); host-visible on return. */
int nb200_fmaxabs(nb200_ctx* ctx, const nb200_buf* a, nb200_real* result);
/* Periodic wrap of the three position rows into [-b, b] (kclamp_coord, impl.cu:724-747). */
int nb200_clamp(nb200_ctx* ctx, nb200_buf* y, nb200_real b);

/* ---- conservation report (device-side nbody_data::print_statistics sums) ------------------------
 * From a state vector y: out[0..2] total impulse sum m v, out[3..5] impulse moment sum r x m v, out[6] kinetic energy
 * sum m v^2 / 2, out[7] potential energy -1/2 sum_{i != j, r2 >= 1e-8} m_i m_j / r (0 unless with_energy; O(N^2) on the
 * device instead of the reference's single-threaded host loop, nbody_data.cpp:81-87), out[8..10] mass centre.
 * Always double, host-visible on return. With nranks > 1 every rank receives the global sums. */
int nb200_statistics(nb200_ctx* ctx, const nb200_buf* y, int with_energy, double out[11]);

/* ---- body arrays <-> state vector ------------------------------------------------------------
 * The AoS <-> [rx|ry|rz|vx|vy|vz] transposes of nbody_engine_cuda::init / get_data (nbody_engine_cuda.cpp:113-139,
 * 141-175: host loops over a full-buffer copy) done on the device: pos_xyz / vel_xyz are the N x 3 arrays behind
 * nbody_data::get_vertites() / get_velosites(); each shard moves its contiguous body range with two DMA copies.
 * nb200_host_register pins such an array once (cudaHostRegister), so the copies go at PCIe speed without staging. */
int nb200_write_bodies(nb200_ctx* ctx, nb200_buf* y, const nb200_real* pos_xyz, const nb200_real* vel_xyz);
int nb200_read_bodies(nb200_ctx* ctx, const nb200_buf* y, nb200_real* pos_xyz, nb200_real* vel_xyz);
int nb200_host_register(nb200_ctx* ctx, void* host, size_t bytes);
int nb200_host_unregister(nb200_ctx* ctx, void* host);

/* ---- solver steps as CUDA graphs ---------------------------------------------------------------
 * With nb200_set_option(ctx, "step_graph", 1) the library watches the calls between two step boundaries and files
 * every distinct step (same operations, buffers and scalars, nothing host-visible inside except fmaxabs) in a small
 * table, together with the step that followed it. A step predicted to repeat a filed one is captured once into a
 * CUDA graph and from then on launched as ONE graph per step (e.g. nbody_solver_rk4.cpp:30-62; periodic patterns --
 * Adams' rotating history, Bulirsch-Stoer's sub-steps, tree_build_rate -- take one entry per phase); any deviation
 * falls back to issuing the calls one by one, so results never differ from the eager engine. The adapter calls nb200_step_boundary from advise_time(), which every solver
 * calls exactly once at the end of a step. Single-shard contexts only (ignored with lanes or ranks). */
int nb200_step_boundary(nb200_ctx* ctx);
/* out[0] = graphs launched, out[1] = replays abandoned, out[2] = state of the NEXT step (0 off, 1 record, 2 capture,
 * 3 replay), out[3] = kernel launches inside the step replayed last, out[4] = distinct steps in the table. */
int nb200_step_graph_stats(const nb200_ctx* ctx, unsigned long long out[5]);

/* ---- instrumentation ------------------------------------------------------ */
/* Number of nb200 kernels launched since creation (all lanes). */
unsigned long long nb200_launch_count(const nb200_ctx* ctx);
/* Device time of the most recent fcompute on lane 0, split by phase, in ms:
 * out[0] = pack + gather, out[1] = tree build / refresh (BH only),
 * out[2] = force kernel (pairs or walk), out[3] = reduce/epilogue. Synchronises. */
int nb200_last_fcompute_ms(nb200_ctx* ctx, float out[4]);
/* Which kernel family the most recent nb200_fcompute_direct used: 0 = ordered-pair kernel (direct_pairs), -1 = the
 * single-launch kernel for small systems (direct_small), otherwise the tile edge of the symmetric-tile kernel
 * (direct_sym_tiles). */
int nb200_last_direct_path(const nb200_ctx* ctx);
/* CUDA-event stopwatch on lane 0's stream (the stream the kernels run on): nb200_mark records event
 * `slot` (0..7); nb200_elapsed_ms synchronises on slot b and returns the device time from a to b. */
int nb200_mark(nb200_ctx* ctx, int slot);
int nb200_elapsed_ms(nb200_ctx* ctx, int slot_a, int slot_b, float* ms);
/* FP64/FP32 FMA-pipe peak probe: runs a dependent-chain-free FMA kernel for
 * ~`ms` milliseconds and returns achieved FMA instructions (per lane) per second. */
int nb200_probe_fma_peak(nb200_ctx* ctx, double ms, double* fma_lane_per_s);
/* Tunables: "direct_targets_per_thread" (1, 2, 4), "direct_segments", "direct_symmetric" / "direct_small" (-1 automatic,
 * 0 off, 1 on), "direct_sym_tile" (power of two, 256..8192), "walk_mode" (0 = automatic = 8: grouped
 * walk, lanes on nodes while deciding and on targets while summing; 1 = one thread per target; 32 / 2 / 4 = warp-coherent walk
 * with one / two / four targets per lane), "walk_lpt" (walk CTAs launched longest walk first: -1 automatic, 0 off, 1 on),
 * "timing" (0/1: phase events), "step_graph" (0/1, above), "use_nccl" (0/1: the lanes of ONE process exchange shards with
 * NCCL, one communicator per lane from ncclCommInitAll -- the reference's use_nccl=1, nbody_engine_cuda.cpp:100-106 --
 * instead of peer copies and peer loads; needs distinct devices, NB200_ERR_UNSUPPORTED otherwise and nothing changes).
 * 0 = automatic where applicable. */
int nb200_set_option(nb200_ctx* ctx, const char* name, long long value);

#ifdef __cplusplus
}
#endif
#endif /* NB200_H */
