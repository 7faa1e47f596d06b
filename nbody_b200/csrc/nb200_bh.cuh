// nb200 -- Barnes-Hut (kd-heap build, node update, tree walk). Filled in below.
#ifndef NB200_BH_CUH
#define NB200_BH_CUH
#include "nb200_common.cuh"
struct bh_state { int dummy; };
static void bh_free(bh_state* s) { delete s; }
static int bh_fcompute(nb200_ctx*, nb200_lane&, const real*, real*, size_t, int&, std::string& err)
{
	err = "not built yet";
	return NB200_ERR_UNSUPPORTED;
}
static int bh_export(nb200_ctx*, nb200_lane&, real*, real*, int*, std::string& err)
{
	err = "not built yet";
	return NB200_ERR_UNSUPPORTED;
}
#endif
