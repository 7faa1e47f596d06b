#include "nbody_engine_b200.h"

#include <QDebug>
#include <algorithm>
#include <stdlib.h>

#include "nb200.h"

static_assert(sizeof(nb200_real) == sizeof(nbcoord_t), "libnb200 precision must match NB_COORD_PRECISION");
static_assert(sizeof(nbvertex_t) == 3 * sizeof(nbcoord_t), "nbvertex_t must be three packed coordinates");

class nbody_engine_b200::smemory : public nbody_engine::memory
{
	nb200_ctx*	m_ctx;
	nb200_buf*	m_buf;
	size_t		m_size;
public:
	smemory(nb200_ctx* ctx, size_t size) : m_ctx(ctx), m_buf(nullptr), m_size(0)
	{
		if(nb200_alloc(ctx, size, &m_buf) == NB200_OK)
		{
			m_size = size;
		}
	}
	~smemory()
	{
		nb200_free(m_ctx, m_buf);
	}
	bool valid() const
	{
		return m_buf != nullptr;
	}
	nb200_buf* buf() const
	{
		return m_buf;
	}
	const nb200_ctx* owner() const
	{
		return m_ctx;
	}
	size_t size() const override
	{
		return m_size;
	}
};

struct nbody_engine_b200::data
{
	e_force				m_force;
	nbcoord_t			m_ratio;
	size_t				m_tree_build_rate;
	e_tree_layout		m_tree_layout;
	std::vector<int>	m_device_ids;
	nb200_ctx*			m_ctx;
	smemory*			m_y;
	nbody_data*			m_data;
	bool				m_step_graph;
	bool				m_use_nccl;
	int					m_block_size;
	void*				m_pinned[2];
	data() : m_force(ef_direct), m_ratio(10), m_tree_build_rate(0), m_tree_layout(etl_heap_stackless),
		m_device_ids(1, 0), m_ctx(nullptr), m_y(nullptr), m_data(nullptr), m_step_graph(true), m_use_nccl(false), m_block_size(0)
	{
		m_pinned[0] = m_pinned[1] = nullptr;
	}
	//! nbody_data's body arrays are pinned for as long as this engine serves them: get_data becomes two DMA copies
	void pin(nbody_data* body_data)
	{
		unpin();
		if(m_ctx == nullptr || body_data == nullptr)
		{
			return;
		}
		void*	arr[2] = {body_data->get_vertites(), body_data->get_velosites()};
		for(int i = 0; i != 2; ++i)
		{
			if(nb200_host_register(m_ctx, arr[i], body_data->get_count() * sizeof(nbvertex_t)) == NB200_OK)
			{
				m_pinned[i] = arr[i];
			}
		}
	}
	void unpin()
	{
		for(int i = 0; i != 2; ++i)
		{
			if(m_pinned[i] != nullptr && m_ctx != nullptr)
			{
				nb200_host_unregister(m_ctx, m_pinned[i]);
			}
			m_pinned[i] = nullptr;
		}
	}
	//! memory -> C-ABI handle; NULL (after logging) for foreign or NULL memory, like the reference's dynamic_cast checks
	nb200_buf* handle(const memory* m, const char* name) const
	{
		const smemory* s = dynamic_cast<const smemory*>(m);
		if(s == nullptr || s->owner() != m_ctx)
		{
			qDebug() << name << "is not smemory";
			return nullptr;
		}
		return s->buf();
	}
	//! Argument errors are logged and the call returns, as every reference engine does; a CUDA or NCCL runtime error
	//! ends the program like the cuda engines' check macros (nbody_engine_cuda.cpp:8-24: qDebug + exit(3)) -- a solver
	//! must not go on integrating with an f nobody wrote. (The library itself never exits: that choice is the host's.)
	void check(int rc, const char* what) const
	{
		if(rc != NB200_OK)
		{
			qDebug() << what << nb200_last_error(m_ctx);
		}
		if(rc == NB200_ERR_CUDA || rc == NB200_ERR_NCCL)
		{
			exit(3);
		}
	}
};

nbody_engine_b200::nbody_engine_b200(e_force force, nbcoord_t distance_to_node_radius_ratio,
									 size_t tree_build_rate, e_tree_layout tl) :
	d(new data())
{
	d->m_force = force;
	d->m_ratio = distance_to_node_radius_ratio;
	d->m_tree_build_rate = tree_build_rate;
	d->m_tree_layout = tl;
}

nbody_engine_b200::~nbody_engine_b200()
{
	delete d->m_y;
	d->unpin();
	nb200_destroy(d->m_ctx);
	delete d;
}

const char* nbody_engine_b200::type_name() const
{
	return d->m_force == ef_direct ? "nbody_engine_b200" : "nbody_engine_b200_bh";
}

bool nbody_engine_b200::ensure_context()
{
	if(d->m_ctx != nullptr)
	{
		return true;
	}
	int rc = nb200_create(&d->m_ctx, d->m_device_ids.data(), static_cast<int>(d->m_device_ids.size()), 0, 1, nullptr);
	if(rc != NB200_OK)
	{
		qDebug() << "nb200_create failed" << rc;
		d->m_ctx = nullptr;
		return false;
	}
	if(d->m_force == ef_barnes_hut)
	{
		int layout = (d->m_tree_layout == etl_heap) ? NB200_TREE_HEAP : NB200_TREE_HEAP_STACKLESS;
		d->check(nb200_bh_configure(d->m_ctx, d->m_ratio, layout, d->m_tree_build_rate), "nb200_bh_configure");
	}
	if(d->m_step_graph)
	{
		d->check(nb200_set_option(d->m_ctx, "step_graph", 1), "step_graph");
	}
	if(d->m_use_nccl && d->m_device_ids.size() > 1)
	{
		// use_nccl=1 as in the cuda engines (nbody_engine_cuda.cpp:100-106): one communicator per device of this process.
		// A device list that repeats a device cannot use NCCL; the library says so and keeps its peer copies.
		d->check(nb200_set_option(d->m_ctx, "use_nccl", 1), "use_nccl");
	}
	return true;
}

bool nbody_engine_b200::init(nbody_data* body_data)
{
	if(body_data == nullptr || body_data->get_count() == 0 || !ensure_context())
	{
		return false;
	}
	d->m_data = body_data;
	size_t	count = body_data->get_count();
	if(d->m_force == ef_barnes_hut && (count & (count - 1)) != 0)
	{
		// The kd-heap keeps its leaves at [N, 2N): in bounds only for N = 2^k. The reference sizes its arrays 2N all the
		// same (nbody_space_heap.cpp:24) and overruns them for any other N; here such a body set is refused up front.
		qDebug() << "b200_bh needs a power-of-two body count (kd-heap layout), got" << count;
		return false;
	}
	// a re-init with another body count: the old state vector goes first (the library refuses to change N under live
	// state vectors of the old size)
	delete d->m_y;
	d->m_y = nullptr;
	if(nb200_set_bodies(d->m_ctx, count, body_data->get_mass()) != NB200_OK)
	{
		qDebug() << "nb200_set_bodies" << nb200_last_error(d->m_ctx);
		return false;
	}
	d->m_y = dynamic_cast<smemory*>(create_buffer(sizeof(nbcoord_t) * problem_size()));
	if(d->m_y == nullptr)
	{
		return false;
	}
	// AoS nbody_data -> [rx|ry|rz|vx|vy|vz] (nbody_engine_cuda.cpp:113-139 does this in a host loop): the body arrays
	// go to the device as they are and are transposed there
	d->pin(body_data);
	int rc = nb200_write_bodies(d->m_ctx, d->m_y->buf(), &body_data->get_vertites()->x, &body_data->get_velosites()->x);
	d->check(rc, "init");
	return rc == NB200_OK;
}

void nbody_engine_b200::get_data(nbody_data* body_data)
{
	if(d->m_y == nullptr || body_data == nullptr)
	{
		return;
	}
	if(6 * body_data->get_count() != problem_size())
	{
		qDebug() << "get_data: body count does not match the engine's";
		return;
	}
	// device-side transpose + two DMA copies per shard straight into nbody_data's arrays
	// (the reference: full-buffer D2H + host loop, nbody_engine_cuda.cpp:141-175)
	d->check(nb200_read_bodies(d->m_ctx, d->m_y->buf(), &body_data->get_vertites()->x, &body_data->get_velosites()->x), "get_data");
}

size_t nbody_engine_b200::problem_size() const
{
	return d->m_data != nullptr ? 6 * d->m_data->get_count() : 0;
}

nbody_engine::memory* nbody_engine_b200::get_y()
{
	return d->m_y;
}

// time and step live in nbody_data, as in every reference engine (nbody_engine_cuda.cpp:177-200)
void nbody_engine_b200::advise_time(const nbcoord_t& dt)
{
	d->m_data->advise_time(dt);
	// every solver ends its step here: the boundary nb200's step graphs are cut at (no-op unless step_graph=1)
	nb200_step_boundary(d->m_ctx);
}

nbcoord_t nbody_engine_b200::get_time() const
{
	return d->m_data->get_time();
}

void nbody_engine_b200::set_time(nbcoord_t t)
{
	d->m_data->set_time(t);
}

size_t nbody_engine_b200::get_step() const
{
	return d->m_data->get_step();
}

void nbody_engine_b200::set_step(size_t s)
{
	d->m_data->set_step(s);
}

void nbody_engine_b200::fcompute(const nbcoord_t& t, const memory* _y, memory* _f)
{
	Q_UNUSED(t);
	nb200_buf*	y = d->handle(_y, "y");
	if(y == nullptr)
	{
		return;
	}
	nb200_buf*	f = d->handle(_f, "f");
	if(f == nullptr)
	{
		return;
	}
	advise_compute_count();
	if(d->m_force == ef_direct)
	{
		d->check(nb200_fcompute_direct(d->m_ctx, y, f), "fcompute");
	}
	else
	{
		d->check(nb200_fcompute_bh(d->m_ctx, y, f, get_step()), "fcompute");
	}
}

void nbody_engine_b200::clamp(memory* _y, nbcoord_t b)
{
	nb200_buf*	y = d->handle(_y, "y");
	if(y == nullptr)
	{
		return;
	}
	d->check(nb200_clamp(d->m_ctx, y, b), "clamp");
}

nbody_engine::memory* nbody_engine_b200::create_buffer(size_t s)
{
	if(!ensure_context())
	{
		return nullptr;
	}
	smemory*	mem = new smemory(d->m_ctx, s);
	if(!mem->valid())
	{
		delete mem;
		return nullptr;
	}
	return mem;
}

void nbody_engine_b200::free_buffer(memory* mem)
{
	delete mem;
}

void nbody_engine_b200::read_buffer(void* dst, const memory* _src)
{
	nb200_buf*	src = d->handle(_src, "src");
	if(src == nullptr)
	{
		return;
	}
	d->check(nb200_read(d->m_ctx, dst, src), "read_buffer");
}

void nbody_engine_b200::write_buffer(memory* _dst, const void* src)
{
	nb200_buf*	dst = d->handle(_dst, "dst");
	if(dst == nullptr)
	{
		return;
	}
	d->check(nb200_write(d->m_ctx, dst, src), "write_buffer");
}

void nbody_engine_b200::copy_buffer(memory* _a, const memory* _b)
{
	nb200_buf*	a = d->handle(_a, "a");
	if(a == nullptr)
	{
		return;
	}
	nb200_buf*	b = d->handle(_b, "b");
	if(b == nullptr)
	{
		return;
	}
	d->check(nb200_copy(d->m_ctx, a, b), "copy_buffer");
}

void nbody_engine_b200::fill_buffer(memory* _a, const nbcoord_t& value)
{
	nb200_buf*	a = d->handle(_a, "a");
	if(a == nullptr)
	{
		return;
	}
	d->check(nb200_fill(d->m_ctx, a, value), "fill_buffer");
}

void nbody_engine_b200::fmadd_inplace(memory* _a, const memory* _b, const nbcoord_t& c)
{
	nb200_buf*	a = d->handle(_a, "a");
	if(a == nullptr)
	{
		return;
	}
	nb200_buf*	b = d->handle(_b, "b");
	if(b == nullptr)
	{
		return;
	}
	d->check(nb200_fmadd_inplace(d->m_ctx, a, b, c), "fmadd_inplace");
}

void nbody_engine_b200::fmadd(memory* _a, const memory* _b, const memory* _c, const nbcoord_t& _d)
{
	nb200_buf*	a = d->handle(_a, "a");
	if(a == nullptr)
	{
		return;
	}
	nb200_buf*	b = d->handle(_b, "b");
	if(b == nullptr)
	{
		return;
	}
	nb200_buf*	c = d->handle(_c, "c");
	if(c == nullptr)
	{
		return;
	}
	d->check(nb200_fmadd(d->m_ctx, a, b, c, _d), "fmadd");
}

namespace {
//! Handles of the first `count` buffers; entries that are not ours become NULL (the C ABI rejects them
//! only when their coefficient is non-zero, exactly as the base-class loops would)
std::vector<const nb200_buf*> handles(const nbody_engine::memory_array& arr, size_t count, const nb200_ctx* ctx)
{
	std::vector<const nb200_buf*>	h(std::max<size_t>(count, 1), nullptr);
	for(size_t k = 0; k < count; ++k)
	{
		const nbody_engine_b200::smemory* s = dynamic_cast<const nbody_engine_b200::smemory*>(arr[k]);
		h[k] = (s != nullptr && s->owner() == ctx) ? s->buf() : nullptr;
	}
	return h;
}
}  // namespace

void nbody_engine_b200::fmaddn_inplace(memory* _a, const memory_array& _b, const nbcoord_t* c, size_t csize)
{
	if(c == NULL)
	{
		return;
	}
	if(csize > _b.size())
	{
		qDebug() << "csize > b.size()";
		return;
	}
	nb200_buf*	a = d->handle(_a, "a");
	if(a == nullptr)
	{
		return;
	}
	std::vector<const nb200_buf*>	b(handles(_b, csize, d->m_ctx));
	d->check(nb200_fmaddn_inplace(d->m_ctx, a, b.data(), c, csize), "fmaddn_inplace");
}

void nbody_engine_b200::fmaddn_corr(memory* _a, memory* _corr, const memory_array& _b, const nbcoord_t* c, size_t csize)
{
	nb200_buf*	a = d->handle(_a, "a");
	if(a == nullptr)
	{
		return;
	}
	nb200_buf*	corr = d->handle(_corr, "corr");
	if(corr == nullptr)
	{
		return;
	}
	if(c == nullptr)
	{
		qDebug() << "c must not be nullptr";
		return;
	}
	if(csize > _b.size())
	{
		qDebug() << "csize > b.size()";
		return;
	}
	std::vector<const nb200_buf*>	b(handles(_b, csize, d->m_ctx));
	d->check(nb200_fmaddn_corr(d->m_ctx, a, corr, b.data(), c, csize), "fmaddn_corr");
}

void nbody_engine_b200::fmaddn(memory* _a, const memory* _b, const memory_array& _c, const nbcoord_t* _d, size_t dsize)
{
	if(_d == NULL)
	{
		qDebug() << "d == NUL";
		return;
	}
	if(dsize > _c.size())
	{
		qDebug() << "dsize > c.size()";
		return;
	}
	nb200_buf*	a = d->handle(_a, "a");
	if(a == nullptr)
	{
		return;
	}
	nb200_buf*	b = nullptr;
	if(_b != NULL)
	{
		b = d->handle(_b, "b");
		if(b == nullptr)
		{
			return;
		}
	}
	std::vector<const nb200_buf*>	c(handles(_c, dsize, d->m_ctx));
	d->check(nb200_fmaddn(d->m_ctx, a, b, c.data(), _d, dsize), "fmaddn");
}

void nbody_engine_b200::fmaxabs(const memory* _a, nbcoord_t& result)
{
	nb200_buf*	a = d->handle(_a, "a");
	if(a == nullptr)
	{
		return;
	}
	nb200_real	r = 0;
	if(nb200_fmaxabs(d->m_ctx, a, &r) == NB200_OK)
	{
		result = r;
	}
	else
	{
		qDebug() << "fmaxabs" << nb200_last_error(d->m_ctx);
	}
}

void nbody_engine_b200::print_info() const
{
	qDebug() << "\tSelected B200 devices:";
	if(d->m_ctx != nullptr)
	{
		char	text[4096];
		if(nb200_describe(d->m_ctx, text, sizeof(text)) == NB200_OK)
		{
			qDebug() << text;
		}
	}
	else
	{
		for(size_t n = 0; n != d->m_device_ids.size(); ++n)
		{
			qDebug() << "\t #" << n << "ID" << d->m_device_ids[n];
		}
	}
	if(d->m_block_size != 0)
	{
		qDebug() << "\t" << "block_size:" << d->m_block_size << "(ignored: tile and CTA shapes are chosen by the library)";
	}
	qDebug() << "\t" << "use_nccl:" << (d->m_use_nccl ? "1" : "0") << (d->m_device_ids.size() > 1 ? "" : "(one device: nothing to exchange)");
	if(d->m_step_graph && d->m_device_ids.size() > 1)
	{
		qDebug() << "\t" << "step_graph: ignored with several devices";
	}
	if(d->m_force == ef_barnes_hut)
	{
		qDebug() << "\t" << "distance_to_node_radius_ratio:" << d->m_ratio;
		qDebug() << "\t" << "traverse_type:" << "nested_tree";
		qDebug() << "\t" << "tree_layout:" << tree_layout_name(d->m_tree_layout);
		qDebug() << "\t" << "tree_build_rate:" << d->m_tree_build_rate;
	}
}

int nbody_engine_b200::select_devices(const QString& devices_str)
{
	// Contract of nbody_engine_cuda::select_devices (nbody_engine_cuda.cpp:576-616): a comma list of device ordinals,
	// every one of them present; 0 on success, anything else leaves the engine's device list untouched.
	int	present = 0;
	if(nb200_device_count(&present) != NB200_OK || present <= 0)
	{
		qDebug() << "No CUDA devices found";
		return -1;
	}
	const QStringList	fields(devices_str.split(",", QString::SkipEmptyParts));
	std::vector<int>	chosen;
	chosen.reserve(static_cast<size_t>(fields.size()));
	for(int k = 0; k != fields.size(); ++k)
	{
		bool		is_number = false;
		const int	ordinal = fields[k].toInt(&is_number);
		if(!is_number)
		{
			qDebug() << "Can't parse device ID" << fields[k];
			return -1;
		}
		if(ordinal < 0 || ordinal >= present)
		{
			qDebug() << "Invalid device ID" << ordinal << "must be in range [0 ..." << present << ")";
			return -1;
		}
		chosen.push_back(ordinal);
	}
	if(chosen.empty())
	{
		qDebug() << "CUDA device list is empty";
		return -1;
	}
	d->m_device_ids.swap(chosen);
	return 0;
}

void nbody_engine_b200::set_block_size(int block_size)
{
	// remembered for print_info only: tile and CTA shapes are picked by the library per problem size
	d->m_block_size = block_size;
}

void nbody_engine_b200::set_use_nccl(bool active)
{
	d->m_use_nccl = active;
	if(d->m_ctx != nullptr && d->m_device_ids.size() > 1)
	{
		d->check(nb200_set_option(d->m_ctx, "use_nccl", active ? 1 : 0), "use_nccl");
	}
}

void nbody_engine_b200::set_step_graph(bool active)
{
	d->m_step_graph = active;
	if(d->m_ctx != nullptr)
	{
		d->check(nb200_set_option(d->m_ctx, "step_graph", active ? 1 : 0), "step_graph");
	}
}

bool nbody_engine_b200::step_graph_stats(unsigned long long out[5]) const
{
	return d->m_ctx != nullptr && nb200_step_graph_stats(d->m_ctx, out) == NB200_OK;
}

bool nbody_engine_b200::statistics(const memory* _y, bool with_energy, double out[11])
{
	nb200_buf*	y = d->handle(_y, "y");
	if(y == nullptr)
	{
		return false;
	}
	int rc = nb200_statistics(d->m_ctx, y, with_energy ? 1 : 0, out);
	d->check(rc, "statistics");
	return rc == NB200_OK;
}

unsigned long long nbody_engine_b200::launch_count() const
{
	return nb200_launch_count(d->m_ctx);
}

void nbody_engine_b200::synchronize()
{
	if(d->m_ctx != nullptr)
	{
		nb200_sync(d->m_ctx);
	}
}

nbody_engine* nbody_create_engine_b200(const QVariantMap& param)
{
	const QString type(param.value("engine").toString());
	nbody_engine_b200*	engine = nullptr;
	if(type == "b200")
	{
		engine = new nbody_engine_b200();
	}
	else if(type == "b200_bh")
	{
		nbcoord_t	ratio = param.value("distance_to_node_radius_ratio", 10).toDouble();
		size_t		tree_build_rate = param.value("tree_build_rate", 0).toULongLong();
		QString		strtl(param.value("tree_layout", "heap").toString());	// the cuda_bh_tex default (nbody_engines.cpp:64)
		e_tree_layout tl = tree_layout_from_str(strtl);
		if(tl != etl_heap && tl != etl_heap_stackless)
		{
			qDebug() << "Invalid tree_layout. Allowed values are 'heap' or 'heap_stackless'";
			return NULL;
		}
		engine = new nbody_engine_b200(nbody_engine_b200::ef_barnes_hut, ratio, tree_build_rate, tl);
	}
	else
	{
		return NULL;
	}
	QString	devices(param.value("device", "0").toString());
	if(0 != engine->select_devices(devices))
	{
		delete engine;
		return NULL;
	}
	engine->set_block_size(param.value("block_size", NBODY_DATA_BLOCK_SIZE).toInt());
	engine->set_use_nccl(param.value("use_nccl", false).toBool());
	// on by default: results are bit-identical to step_graph=0 (tests/test_stepgraph_gpu.py), ignored with several devices
	engine->set_step_graph(param.value("step_graph", true).toBool());
	return engine;
}

// Plain-C doorway for the ctypes test harness: "engine=b200_bh;device=0,0;tree_layout=heap"
extern "C" __attribute__((visibility("default"))) void* nbody_engine_b200_create(const char* params)
{
	QVariantMap	m;
	QStringList	items(QString(params ? params : "").split(";", QString::SkipEmptyParts));
	for(int i = 0; i < items.size(); ++i)
	{
		int eq = items[i].indexOf("=");
		if(eq >= 0)
		{
			m[items[i].mid(0, eq).trimmed()] = QVariant(items[i].mid(eq + 1).trimmed());
		}
	}
	return nbody_create_engine_b200(m);
}

extern "C" __attribute__((visibility("default"))) unsigned long long nbody_engine_b200_launch_count(void* engine)
{
	nbody_engine_b200* e = dynamic_cast<nbody_engine_b200*>(static_cast<nbody_engine*>(engine));
	return e != nullptr ? e->launch_count() : 0;
}

extern "C" __attribute__((visibility("default"))) void nbody_engine_b200_synchronize(void* engine)
{
	nbody_engine_b200* e = dynamic_cast<nbody_engine_b200*>(static_cast<nbody_engine*>(engine));
	if(e != nullptr)
	{
		e->synchronize();
	}
}

extern "C" __attribute__((visibility("default"))) int nbody_engine_b200_step_graph_stats(void* engine, unsigned long long out[5])
{
	nbody_engine_b200* e = dynamic_cast<nbody_engine_b200*>(static_cast<nbody_engine*>(engine));
	return (e != nullptr && e->step_graph_stats(out)) ? 0 : -1;
}
