/* TEST INFRASTRUCTURE ONLY -- see nbody_oracle.h. Never used by the product path. */
#include "nbody_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

#if ORC_PRECISION == 1
#define ORC_SQRT sqrtf
#define ORC_FABS fabsf
#define ORC_MIN_DISTANCE 1e-8f
#else
#define ORC_SQRT sqrt
#define ORC_FABS fabs
#define ORC_MIN_DISTANCE 1e-8
#endif

int orc_real_size(void) { return (int)sizeof(orc_real); }
int orc_max_threads(void) { return omp_get_max_threads(); }
void orc_set_threads(int n) { omp_set_num_threads(n); }

typedef struct { orc_real x, y, z; } v3;

/* nbody_data::force (nbody_data.cpp:35-44), GravityConst = 1 */
static inline v3 force(v3 v1, v3 v2, orc_real mass1, orc_real mass2)
{
	v3			dr = {v1.x - v2.x, v1.y - v2.y, v1.z - v2.z};
	orc_real	r2 = dr.x * dr.x + dr.y * dr.y + dr.z * dr.z;
	if(r2 < ORC_MIN_DISTANCE)
	{
		r2 = ORC_MIN_DISTANCE;
	}
	orc_real	k = (-(orc_real)1 * mass1 * mass2) / (r2 * ORC_SQRT(r2));
	v3			f = {dr.x * k, dr.y * k, dr.z * k};
	return f;
}

/* ---- direct ---------------------------------------------------------------- */
void orc_fcompute_openmp(size_t n, const orc_real* y, const orc_real* mass, orc_real* f)
{
	const orc_real *rx = y, *ry = y + n, *rz = y + 2 * n, *vx = y + 3 * n, *vy = y + 4 * n, *vz = y + 5 * n;
	#pragma omp parallel for
	for(size_t b1 = 0; b1 < n; ++b1)
	{
		v3 v1 = {rx[b1], ry[b1], rz[b1]};
		v3 total = {0, 0, 0};
		for(size_t b2 = 0; b2 != n; ++b2)
		{
			if(b1 == b2)
			{
				continue;
			}
			v3 v2 = {rx[b2], ry[b2], rz[b2]};
			v3 fo = force(v1, v2, mass[b1], mass[b2]);
			total.x += fo.x;
			total.y += fo.y;
			total.z += fo.z;
		}
		f[b1] = vx[b1];
		f[n + b1] = vy[b1];
		f[2 * n + b1] = vz[b1];
		f[3 * n + b1] = total.x / mass[b1];
		f[4 * n + b1] = total.y / mass[b1];
		f[5 * n + b1] = total.z / mass[b1];
	}
}

void orc_accel_subset(size_t n, const orc_real* y, const orc_real* mass, const size_t* targets, size_t nt, orc_real* acc)
{
	const orc_real *rx = y, *ry = y + n, *rz = y + 2 * n;
	#pragma omp parallel for schedule(dynamic, 1)
	for(size_t t = 0; t < nt; ++t)
	{
		size_t b1 = targets[t];
		v3 v1 = {rx[b1], ry[b1], rz[b1]};
		v3 total = {0, 0, 0};
		for(size_t b2 = 0; b2 != n; ++b2)
		{
			if(b1 == b2)
			{
				continue;
			}
			v3 v2 = {rx[b2], ry[b2], rz[b2]};
			v3 fo = force(v1, v2, mass[b1], mass[b2]);
			total.x += fo.x;
			total.y += fo.y;
			total.z += fo.z;
		}
		acc[t] = total.x / mass[b1];
		acc[nt + t] = total.y / mass[b1];
		acc[2 * nt + t] = total.z / mass[b1];
	}
}

void orc_accel_subset_ld(size_t n, const orc_real* y, const orc_real* mass, const size_t* targets, size_t nt, orc_real* acc)
{
	const orc_real *rx = y, *ry = y + n, *rz = y + 2 * n;
	#pragma omp parallel for schedule(dynamic, 1)
	for(size_t t = 0; t < nt; ++t)
	{
		size_t		b1 = targets[t];
		long double	ax = 0, ay = 0, az = 0;
		for(size_t b2 = 0; b2 != n; ++b2)
		{
			long double dx = (long double)rx[b2] - rx[b1], dy = (long double)ry[b2] - ry[b1], dz = (long double)rz[b2] - rz[b1];
			long double r2 = dx * dx + dy * dy + dz * dz;
			if(r2 < (long double)ORC_MIN_DISTANCE)
			{
				r2 = (long double)ORC_MIN_DISTANCE;
			}
			long double k = (long double)mass[b2] / (r2 * sqrtl(r2));
			ax += dx * k;
			ay += dy * k;
			az += dz * k;
		}
		acc[t] = (orc_real)ax;
		acc[nt + t] = (orc_real)ay;
		acc[2 * nt + t] = (orc_real)az;
	}
}

#define ORC_BLOCK 64 /* NBODY_DATA_BLOCK_SIZE, nbtype.h:77 */
void orc_fcompute_block(size_t n, const orc_real* y, const orc_real* mass, orc_real* f)
{
	const orc_real *rx = y, *ry = y + n, *rz = y + 2 * n, *vx = y + 3 * n, *vy = y + 4 * n, *vz = y + 5 * n;
	#pragma omp parallel for
	for(size_t n1 = 0; n1 < n; n1 += ORC_BLOCK)
	{
		orc_real x1[ORC_BLOCK], y1[ORC_BLOCK], z1[ORC_BLOCK], fx[ORC_BLOCK], fy[ORC_BLOCK], fz[ORC_BLOCK];
		for(size_t b1 = 0; b1 != ORC_BLOCK; ++b1)
		{
			x1[b1] = rx[n1 + b1];
			y1[b1] = ry[n1 + b1];
			z1[b1] = rz[n1 + b1];
			fx[b1] = fy[b1] = fz[b1] = 0;
		}
		for(size_t n2 = 0; n2 < n; n2 += ORC_BLOCK)
		{
			orc_real x2[ORC_BLOCK], y2[ORC_BLOCK], z2[ORC_BLOCK], m2[ORC_BLOCK];
			for(size_t b2 = 0; b2 != ORC_BLOCK; ++b2)
			{
				x2[b2] = rx[n2 + b2];
				y2[b2] = ry[n2 + b2];
				z2[b2] = rz[n2 + b2];
				m2[b2] = mass[n2 + b2];
			}
			for(size_t b1 = 0; b1 != ORC_BLOCK; ++b1)
			{
				for(size_t b2 = 0; b2 != ORC_BLOCK; ++b2)
				{
					orc_real dx = x1[b1] - x2[b2], dy = y1[b1] - y2[b2], dz = z1[b1] - z2[b2];
					orc_real r2 = dx * dx + dy * dy + dz * dz;
					if(r2 < ORC_MIN_DISTANCE)
					{
						r2 = ORC_MIN_DISTANCE;
					}
					orc_real r = ORC_SQRT(r2);
					orc_real coeff = m2[b2] / (r * r2);
					fx[b1] -= dx * coeff;
					fy[b1] -= dy * coeff;
					fz[b1] -= dz * coeff;
				}
			}
		}
		for(size_t b1 = 0; b1 != ORC_BLOCK; ++b1)
		{
			f[n1 + b1] = vx[n1 + b1];
			f[n + n1 + b1] = vy[n1 + b1];
			f[2 * n + n1 + b1] = vz[n1 + b1];
			f[3 * n + n1 + b1] = fx[b1];
			f[4 * n + n1 + b1] = fy[b1];
			f[5 * n + n1 + b1] = fz[b1];
		}
	}
}

/* ---- state ops ------------------------------------------------------------- */
void orc_fmadd_inplace(orc_real* a, const orc_real* b, orc_real c, size_t count)
{
	for(size_t i = 0; i < count; ++i)
	{
		a[i] += b[i] * c;
	}
}

void orc_fmadd(orc_real* a, const orc_real* b, const orc_real* c, orc_real d, size_t count)
{
	for(size_t i = 0; i < count; ++i)
	{
		a[i] = b[i] + c[i] * d;
	}
}

/* nbody_engine::fmaddn_inplace (nbody_engine.cpp:47-66): one fmadd_inplace per non-zero c[k] */
void orc_fmaddn_inplace(orc_real* a, const orc_real* const* b, const orc_real* c, size_t csize, size_t count)
{
	if(c == NULL)
	{
		return;
	}
	for(size_t k = 0; k != csize; ++k)
	{
		if(c[k] == 0)
		{
			continue;
		}
		orc_fmadd_inplace(a, b[k], c[k], count);
	}
}

/* nbody_engine::fmaddn (nbody_engine.cpp:75-113) */
void orc_fmaddn(orc_real* a, const orc_real* b, const orc_real* const* c, const orc_real* d, size_t dsize, size_t count)
{
	int initiated = 0;
	if(d == NULL)
	{
		return;
	}
	if(b == NULL)
	{
		for(size_t i = 0; i < count; ++i)
		{
			a[i] = 0;
		}
		initiated = 1;
	}
	for(size_t k = 0; k != dsize; ++k)
	{
		if(d[k] == 0)
		{
			continue;
		}
		orc_fmadd(a, initiated ? a : b, c[k], d[k], count);
		initiated = 1;
	}
}

/* summation_k (summation.h:8-14) inside nbody_engine_openmp::fmaddn_corr (:214-269); volatile as there */
void orc_fmaddn_corr(orc_real* a_, orc_real* corr_, const orc_real* const* b, const orc_real* c, size_t csize, size_t count)
{
	volatile orc_real* a = a_;
	volatile orc_real* corr = corr_;
	for(size_t i = 0; i < count; ++i)
	{
		for(size_t k = 0; k < csize; ++k)
		{
			if(c[k] == 0)
			{
				continue;
			}
			volatile orc_real term = b[k][i] * c[k];
			volatile orc_real corrected = term - corr[i];
			volatile orc_real new_sum = a[i] + corrected;
			corr[i] = (new_sum - a[i]) - corrected;
			a[i] = new_sum;
		}
	}
}

orc_real orc_fmaxabs(const orc_real* a, size_t count)
{
	if(count == 0)
	{
		return 0;
	}
	orc_real result = ORC_FABS(a[0]);
	for(size_t i = 0; i < count; ++i)
	{
		orc_real v = ORC_FABS(a[i]);
		if(v > result)
		{
			result = v;
		}
	}
	return result;
}

/* nbody_engine_simple::clamp (nbody_engine_simple.cpp:175-199) */
void orc_clamp(orc_real* y, orc_real b, size_t n)
{
	orc_real diam = 2 * b;
	for(size_t i = 0; i < 3 * n; ++i)
	{
		if(y[i] > +b) { y[i] -= diam; }
		if(y[i] < -b) { y[i] += diam; }
	}
}

/* ---- heap index algebra (nbody_space_heap_func_priv.h:4-72) ------------------ */
size_t orc_heap_left(size_t idx) { return idx << 1; }
size_t orc_heap_right(size_t idx) { return (idx << 1) + 1; }
size_t orc_heap_parent(size_t idx) { return idx >> 1; }
size_t orc_heap_next_down(size_t idx)
{
	/* the portable branch of the reference (:52-62): climb while the node is a right child */
	while(idx & 1)
	{
		size_t parent = idx >> 1;
		if(parent == 1)
		{
			return 1;
		}
		idx = parent;
	}
	return idx + 1;
}
size_t orc_heap_skip(size_t idx) { return orc_heap_next_down(idx); }
size_t orc_heap_next_up(size_t idx, size_t tree_size)
{
	size_t left = idx << 1;
	if(left < tree_size)
	{
		return left;
	}
	return orc_heap_next_down(idx);
}

/* ---- kd-heap build (nbody_space_heap.cpp:13-37,119-191) ----------------------- */
typedef struct
{
	size_t			n;
	const orc_real*	r[3];
	const orc_real*	mass;
	orc_real		ratio_sqr;
	orc_real*		xyzr;
	orc_real*		node_mass;
	orc_real*		bmin;
	orc_real*		bmax;
	long long*		body_n;
} heap_t;

/* Put the `k` smallest (by key) of idx[0..count) first: the set partition std::nth_element
 * guarantees (order inside the halves is irrelevant for the tree). Ties broken by body index. */
static inline int key_less(const orc_real* key, size_t a, size_t b)
{
	return key[a] < key[b] || (key[a] == key[b] && a < b);
}
static void select_k(size_t* idx, size_t count, size_t k, const orc_real* key)
{
	/* invariant: idx[0..lo) < everything in [lo, hi) < idx[hi..count), lo <= k < hi */
	size_t lo = 0, hi = count;
	while(hi - lo > 1)
	{
		size_t mid = lo + (hi - lo) / 2, last = hi - 1;
		/* median of three as pivot, moved to the end; Lomuto partition on the total order (key, index) */
		size_t p = key_less(key, idx[lo], idx[mid])
					   ? (key_less(key, idx[mid], idx[last]) ? mid : (key_less(key, idx[lo], idx[last]) ? last : lo))
					   : (key_less(key, idx[lo], idx[last]) ? lo : (key_less(key, idx[mid], idx[last]) ? last : mid));
		size_t tmp = idx[p];
		idx[p] = idx[last];
		idx[last] = tmp;
		size_t pv = idx[last], store = lo;
		for(size_t q = lo; q < last; ++q)
		{
			if(key_less(key, idx[q], pv))
			{
				tmp = idx[q];
				idx[q] = idx[store];
				idx[store] = tmp;
				++store;
			}
		}
		tmp = idx[store];
		idx[store] = idx[last];
		idx[last] = tmp;
		if(store == k)
		{
			return;
		}
		if(store < k)
		{
			lo = store + 1;
		}
		else
		{
			hi = store;
		}
	}
}

static inline orc_real len3(orc_real x, orc_real y, orc_real z)
{
	return ORC_SQRT(x * x + y * y + z * z);
}

/* nbody_space_heap::update (:119-134) */
static void heap_update(heap_t* h, size_t idx)
{
	size_t		l = idx << 1, r = l + 1;
	orc_real	ml = h->node_mass[l], mr = h->node_mass[r];
	orc_real	m = ml + mr;
	orc_real	cm[3], lo[3], hi[3];
	h->node_mass[idx] = m;
	for(int d = 0; d < 3; ++d)
	{
		cm[d] = (h->xyzr[4 * l + d] * ml + h->xyzr[4 * r + d] * mr) / m;
		lo[d] = h->bmin[3 * l + d] < h->bmin[3 * r + d] ? h->bmin[3 * l + d] : h->bmin[3 * r + d];
		hi[d] = h->bmax[3 * l + d] > h->bmax[3 * r + d] ? h->bmax[3 * l + d] : h->bmax[3 * r + d];
		h->xyzr[4 * idx + d] = cm[d];
		h->bmin[3 * idx + d] = lo[d];
		h->bmax[3 * idx + d] = hi[d];
	}
	orc_real rad = len3(hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]) * (orc_real)0.5 +
				   len3((hi[0] + lo[0]) / 2 - cm[0], (hi[1] + lo[1]) / 2 - cm[1], (hi[2] + lo[2]) / 2 - cm[2]);
	h->xyzr[4 * idx + 3] = (rad * rad) * h->ratio_sqr;
}

static void heap_leaf(heap_t* h, size_t idx, size_t body)
{
	for(int d = 0; d < 3; ++d)
	{
		h->xyzr[4 * idx + d] = h->r[d][body];
		h->bmin[3 * idx + d] = h->r[d][body];
		h->bmax[3 * idx + d] = h->r[d][body];
	}
	h->xyzr[4 * idx + 3] = 0;
}

/* nbody_space_heap::build_p (:136-191) */
static void build_p(heap_t* h, size_t count, size_t* idx, size_t node, size_t dim)
{
	if(count == 1)
	{
		heap_leaf(h, node, idx[0]);
		h->node_mass[node] = h->mass[idx[0]];
		h->body_n[node] = (long long)idx[0];
		return;
	}
	size_t left = count / 2;
	select_k(idx, count, left, h->r[dim]);
	size_t next = (dim + 1) % 3;
	if(count > 4096)
	{
		#pragma omp task
		build_p(h, left, idx, node << 1, next);
		#pragma omp task
		build_p(h, count - left, idx + left, (node << 1) + 1, next);
		#pragma omp taskwait
	}
	else
	{
		build_p(h, left, idx, node << 1, next);
		build_p(h, count - left, idx + left, (node << 1) + 1, next);
	}
	heap_update(h, node);
}

int orc_heap_build(size_t n, const orc_real* y, const orc_real* mass, orc_real ratio,
				   orc_real* xyzr, orc_real* node_mass, orc_real* bmin, orc_real* bmax, long long* body_n)
{
	if(n == 0 || (n & (n - 1)) != 0)
	{
		return -1; /* heap_size = 2n only holds a complete tree */
	}
	heap_t h = {n, {y, y + n, y + 2 * n}, mass, ratio * ratio, xyzr, node_mass, bmin, bmax, body_n};
	size_t* idx = (size_t*)malloc(n * sizeof(size_t));
	for(size_t i = 0; i < n; ++i)
	{
		idx[i] = i;
	}
	for(size_t i = 0; i < 2 * n; ++i)
	{
		body_n[i] = -1; /* TREE_NO_BODY */
	}
	memset(xyzr, 0, 8 * n * sizeof(orc_real));
	memset(node_mass, 0, 2 * n * sizeof(orc_real));
	#pragma omp parallel
	#pragma omp single
	build_p(&h, n, idx, 1, 0);
	free(idx);
	return 0;
}

void orc_heap_rebuild(size_t n, const orc_real* y, orc_real ratio,
					  orc_real* xyzr, const orc_real* node_mass, orc_real* bmin, orc_real* bmax, const long long* body_n)
{
	heap_t h = {n, {y, y + n, y + 2 * n}, NULL, ratio * ratio, xyzr, (orc_real*)node_mass, bmin, bmax, (long long*)body_n};
	#pragma omp parallel for
	for(size_t idx = n; idx < 2 * n; ++idx)
	{
		heap_leaf(&h, idx, (size_t)body_n[idx]);
	}
	for(size_t level = n; level > 1; level /= 2)
	{
		#pragma omp parallel for
		for(size_t idx = level / 2; idx < level; ++idx)
		{
			heap_update(&h, idx);
		}
	}
}

/* ---- walks ------------------------------------------------------------------ */
void orc_fcompute_bh(size_t n, const orc_real* y, const orc_real* mass,
					 const orc_real* xyzr, const orc_real* node_mass, const long long* body_n,
					 int stackless, orc_real* f, unsigned long long* visits_out, unsigned long long* inter_out)
{
	const orc_real		*vx = y + 3 * n, *vy = y + 4 * n, *vz = y + 5 * n;
	const size_t		tree_size = 2 * n;
	unsigned long long	visits = 0, inter = 0;
	/* nbody_space_heap::traverse(Visitor) (nbody_space_heap.h:31-41): one target per leaf, target = leaf centre */
	#pragma omp parallel for schedule(dynamic, 4) reduction(+ : visits, inter)
	for(size_t leaf = n; leaf < tree_size; ++leaf)
	{
		size_t		body1 = (size_t)body_n[leaf];
		v3			v1 = {xyzr[4 * leaf], xyzr[4 * leaf + 1], xyzr[4 * leaf + 2]};
		orc_real	mass1 = node_mass[leaf];
		v3			total = {0, 0, 0};
		if(stackless)
		{
			/* nbody_space_heap_stackless::traverse (nbody_space_heap_stackless.cpp:3-28) */
			size_t curr = 1;
			do
			{
				v3			cm = {xyzr[4 * curr], xyzr[4 * curr + 1], xyzr[4 * curr + 2]};
				orc_real	dx = v1.x - cm.x, dy = v1.y - cm.y, dz = v1.z - cm.z;
				orc_real	d2 = dx * dx + dy * dy + dz * dz;
				++visits;
				if(d2 > xyzr[4 * curr + 3])
				{
					v3 fo = force(v1, cm, mass1, node_mass[curr]);
					total.x += fo.x;
					total.y += fo.y;
					total.z += fo.z;
					++inter;
					curr = orc_heap_skip(curr);
				}
				else
				{
					curr = orc_heap_next_up(curr, tree_size);
				}
			} while(curr != 1);
		}
		else
		{
			/* nbody_space_heap::traverse (nbody_space_heap.cpp:64-97) */
			size_t	stack[64];
			int		top = 0;
			stack[top++] = 1;
			while(top != 0)
			{
				size_t		curr = stack[--top];
				v3			cm = {xyzr[4 * curr], xyzr[4 * curr + 1], xyzr[4 * curr + 2]};
				orc_real	dx = v1.x - cm.x, dy = v1.y - cm.y, dz = v1.z - cm.z;
				orc_real	d2 = dx * dx + dy * dy + dz * dz;
				++visits;
				if(d2 > xyzr[4 * curr + 3])
				{
					v3 fo = force(v1, cm, mass1, node_mass[curr]);
					total.x += fo.x;
					total.y += fo.y;
					total.z += fo.z;
					++inter;
				}
				else
				{
					size_t l = curr << 1, r = l + 1;
					if(r < tree_size) { stack[top++] = r; }
					if(l < tree_size) { stack[top++] = l; }
				}
			}
		}
		/* update_f (nbody_engine_simple_bh.cpp:62-70): divided by mass[body1] */
		f[body1] = vx[body1];
		f[n + body1] = vy[body1];
		f[2 * n + body1] = vz[body1];
		f[3 * n + body1] = total.x / mass[body1];
		f[4 * n + body1] = total.y / mass[body1];
		f[5 * n + body1] = total.z / mass[body1];
	}
	if(visits_out) { *visits_out = visits; }
	if(inter_out) { *inter_out = inter; }
}

void orc_bh_subset(size_t n, const orc_real* xyzr, const orc_real* node_mass, const size_t* leaves, size_t nt, orc_real* acc)
{
	const size_t tree_size = 2 * n;
	#pragma omp parallel for schedule(dynamic, 1)
	for(size_t t = 0; t < nt; ++t)
	{
		size_t		leaf = n + leaves[t];
		v3			v1 = {xyzr[4 * leaf], xyzr[4 * leaf + 1], xyzr[4 * leaf + 2]};
		orc_real	mass1 = node_mass[leaf];
		v3			total = {0, 0, 0};
		size_t		curr = 1;
		do
		{
			v3			cm = {xyzr[4 * curr], xyzr[4 * curr + 1], xyzr[4 * curr + 2]};
			orc_real	dx = v1.x - cm.x, dy = v1.y - cm.y, dz = v1.z - cm.z;
			orc_real	d2 = dx * dx + dy * dy + dz * dz;
			if(d2 > xyzr[4 * curr + 3])
			{
				v3 fo = force(v1, cm, mass1, node_mass[curr]);
				total.x += fo.x;
				total.y += fo.y;
				total.z += fo.z;
				curr = orc_heap_skip(curr);
			}
			else
			{
				curr = orc_heap_next_up(curr, tree_size);
			}
		} while(curr != 1);
		acc[t] = total.x / mass1;
		acc[nt + t] = total.y / mass1;
		acc[2 * nt + t] = total.z / mass1;
	}
}

/* ---- solvers ---------------------------------------------------------------- */
void orc_run_euler(size_t n, orc_real* y, const orc_real* mass, orc_real dt, orc_real max_time)
{
	size_t		ps = 6 * n;
	orc_real*	dy = (orc_real*)malloc(ps * sizeof(orc_real));
	orc_real	t = 0;
	while(t < max_time)
	{
		orc_fcompute_openmp(n, y, mass, dy);
		orc_fmadd_inplace(y, dy, dt, ps);
		t += dt;
	}
	free(dy);
}

void orc_run_rk4(size_t n, orc_real* y, const orc_real* mass, orc_real dt, orc_real max_time)
{
	size_t		ps = 6 * n;
	orc_real*	k[4];
	orc_real*	tmp = (orc_real*)malloc(ps * sizeof(orc_real));
	orc_real	t = 0;
	for(int i = 0; i < 4; ++i)
	{
		k[i] = (orc_real*)malloc(ps * sizeof(orc_real));
	}
	while(t < max_time)
	{
		orc_fcompute_openmp(n, y, mass, k[0]);
		orc_fmadd(tmp, y, k[0], dt / 2, ps);
		orc_fcompute_openmp(n, tmp, mass, k[1]);
		orc_fmadd(tmp, y, k[1], dt / 2, ps);
		orc_fcompute_openmp(n, tmp, mass, k[2]);
		orc_fmadd(tmp, y, k[2], dt, ps);
		orc_fcompute_openmp(n, tmp, mass, k[3]);
		const orc_real coeff[4] = {dt / 6, dt / 3, dt / 3, dt / 6};
		orc_fmaddn_inplace(y, (const orc_real* const*)k, coeff, 4, ps);
		t += dt;
	}
	for(int i = 0; i < 4; ++i)
	{
		free(k[i]);
	}
	free(tmp);
}

/* ---- conservation sums ------------------------------------------------------- */
static inline void kahan(orc_real* sum, orc_real term, orc_real* corr)
{
	/* summation_k (summation.h:8-14) */
	volatile orc_real corrected = term - *corr;
	volatile orc_real new_sum = *sum + corrected;
	*corr = (new_sum - *sum) - corrected;
	*sum = new_sum;
}

/* summation<> starts from element 0 and Kahan-adds the rest with ONE correction variable shared by the
 * three components being separate sums (vertex3 arithmetic is component-wise), summation.h:19-45 */
void orc_statistics(size_t n, const orc_real* y, const orc_real* mass, int with_energy, orc_real* out)
{
	const orc_real *rx = y, *ry = y + n, *rz = y + 2 * n, *vx = y + 3 * n, *vy = y + 4 * n, *vz = y + 5 * n;
	orc_real s[11], c[11];
	memset(s, 0, sizeof(s));
	memset(c, 0, sizeof(c));
	orc_real total_mass = 0, cmass = 0;
	for(size_t i = 0; i < n; ++i)
	{
		orc_real m = mass[i];
		orc_real px = vx[i] * m, py = vy[i] * m, pz = vz[i] * m;
		orc_real t[11] = {px, py, pz,
						  ry[i] * pz - rz[i] * py, rz[i] * px - rx[i] * pz, rx[i] * py - ry[i] * px,
						  (vx[i] * vx[i] + vy[i] * vy[i] + vz[i] * vz[i]) * m, 0,
						  rx[i] * m, ry[i] * m, rz[i] * m};
		if(i == 0)
		{
			memcpy(s, t, sizeof(s));
			total_mass = m;
		}
		else
		{
			for(int q = 0; q < 11; ++q)
			{
				if(q != 7) { kahan(&s[q], t[q], &c[q]); }
			}
			kahan(&total_mass, m, &cmass);
		}
	}
	s[7] = 0;
	if(with_energy)
	{
		/* potential_energy_proxy over n*n entries (summation_proxy.h:68-87, nbody_data.cpp:46-55) */
		orc_real e = 0, ce = 0;
		for(size_t a = 0; a < n; ++a)
		{
			for(size_t b = 0; b < n; ++b)
			{
				orc_real term = 0;
				if(a != b)
				{
					orc_real dx = rx[a] - rx[b], dy = ry[a] - ry[b], dz = rz[a] - rz[b];
					orc_real r2 = dx * dx + dy * dy + dz * dz;
					term = r2 < ORC_MIN_DISTANCE ? 0 : -(mass[a] * mass[b]) / ORC_SQRT(r2);
				}
				if(a == 0 && b == 0) { e = term; }
				else { kahan(&e, term, &ce); }
			}
		}
		s[7] = e / 2;
	}
	s[6] /= 2;
	for(int q = 8; q < 11; ++q)
	{
		s[q] /= total_mass;
	}
	memcpy(out, s, sizeof(s));
}
