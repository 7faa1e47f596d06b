// nbody-simulation, the short form: the reference's own flow (test/main/main.cpp: make the universe, create engine and
// solver through the two factories, init, run with --check_step statistics) without the Qt application object and the
// dump stream, built from the reference's sources with BOTH overlay patches applied (integration/nbody_engines.patch:
// the b200 / b200_bh aliases; integration/nbody_data.patch: print_statistics sums on the device). It is what a
// maintainer gets after applying the patches -- here it serves the parity and timing runs of tests/ and profiles/:
//
//   nbody_sim --engine=b200 --solver=rk4 --stars_count=1024 --max_time=1 --check_step=0.1 --check_list=PLVE
//   nbody_sim --engine=b200 --device=0,1,2,3,4,5,6,7 --use_nccl=1 --solver=rkdp --stars_count=524288 --max_steps=3
//
// Options are the reference's (nbody_arg_parser: --name=value), plus --max_steps=K (stop after K solver steps) and
// --json=1 (one summary line on stdout: steps, fcompute calls, wall seconds, the last statistics line's values).
// Built by oracle/Makefile (target sim) into oracle/_ref/; TEST / MEASUREMENT INFRASTRUCTURE, not the product library.
#include <QDebug>
#include <memory>
#include <omp.h>
#include <stdio.h>
#include <string.h>
#include <string>

#include "nbody_solvers.h"
#include "nbody_engines.h"

namespace {
//! JSON has no nan / inf: a drift that was never computed (check_step=0) prints as null
std::string num(double v)
{
	char buf[64];
	if(v != v || v > 1e300 || v < -1e300) { return "null"; }
	snprintf(buf, sizeof(buf), "%.6e", v);
	return buf;
}

QVariantMap parse(int argc, char* argv[])
{
	QVariantMap	m;
	for(int i = 1; i < argc; ++i)
	{
		const char*	a = argv[i];
		if(a[0] != '-' || a[1] != '-') { continue; }
		const char*	eq = strchr(a, '=');
		if(eq == nullptr) { m[QString(a + 2)] = QVariant("1"); }
		else { m[QString(std::string(a + 2, eq))] = QVariant(eq + 1); }
	}
	return m;
}

//! run() of the reference with a step budget: the same loop body as nbody_solver::run (nbody_solver.cpp:55-105) minus the stream
int run_steps(nbody_solver* solver, nbody_data* data, nbcoord_t max_time, nbcoord_t check_dt, size_t max_steps, bool clamp_to_box)
{
	const nbcoord_t	dt = solver->get_max_step();
	nbcoord_t		last_check = data->get_time();
	size_t			steps = 0;
	while(data->get_time() < max_time && steps < max_steps)
	{
		if(clamp_to_box) { solver->engine()->clamp(solver->engine()->get_y(), data->get_box_size()); }
		solver->advise(dt);
		++steps;
		const nbcoord_t t = data->get_time();
		if(check_dt > 0 && t >= last_check + check_dt - dt * 0.1)
		{
			data->print_statistics(solver->engine());
			last_check = t;
		}
	}
	return 0;
}
}  // namespace

int main(int argc, char* argv[])
{
	QVariantMap	param(parse(argc, argv));
	nbcoord_t	box_size(param.value("box_size", 100).toDouble());
	nbcoord_t	max_time = param.value("max_time", 1).toDouble();
	nbcoord_t	check_step = param.value("check_step", 1e-1).toDouble();
	QString		check_list(param.value("check_list", "PLV").toString());
	QString		initial_state(param.value("initial_state", QString()).toString());
	QString		initial_state_type(param.value("initial_type", "Zeno").toString());
	size_t		max_steps = param.value("max_steps", 0).toULongLong();
	if(param.value("threads", 0).toInt() > 0) { omp_set_num_threads(param.value("threads", 0).toInt()); }

	nbody_data	data;
	if(!initial_state.isEmpty())
	{
		if(!data.load_initial(initial_state, initial_state_type))
		{
			qDebug() << "Can't load initial state" << initial_state;
			return -1;
		}
	}
	else
	{
		data.make_universe(param.value("stars_count", "64").toUInt(), box_size, box_size, box_size);
	}
	std::unique_ptr<nbody_engine>	engine(nbody_create_engine(param));
	if(engine == NULL)
	{
		qDebug() << "Can't create engine" << param.value("engine").toString();
		return -1;
	}
	std::unique_ptr<nbody_solver>	solver(nbody_create_solver(param));
	if(solver == NULL)
	{
		qDebug() << "Can't create solver" << param.value("solver").toString();
		return -1;
	}
	if(!engine->init(&data))
	{
		qDebug() << "Can't init engine" << engine->type_name();
		return -1;
	}
	solver->set_engine(engine.get());
	data.set_check_list(check_list);
	if(param.value("verbose", "0").toInt() != 0)
	{
		qDebug() << "General:";
		qDebug() << "\tStars count:" << data.get_count();
		qDebug() << "\tmax_time:" << max_time;
		qDebug() << "\tcheck_step:" << check_step;
		qDebug() << "\tcheck_list:" << check_list;
		qDebug() << "Solver:" << solver->type_name();
		solver->print_info();
		qDebug() << "Engine:" << solver->engine()->type_name();
		solver->engine()->print_info();
	}
	// warm-up steps are part of the run (the reference has none); the summary separates the first step from the rest
	const double	t0 = omp_get_wtime();
	int				rc = 0;
	double			t_first = 0;
	size_t			calls_first = 0;
	if(max_steps == 0)
	{
		rc = solver->run(&data, NULL, max_time, 0, check_step);
	}
	else
	{
		rc = run_steps(solver.get(), &data, max_time, check_step, 1, param.value("clamp_to_box", false).toBool());
		engine->get_data(&data);	// blocks until the first step has finished on the device
		t_first = omp_get_wtime() - t0;
		calls_first = engine->get_compute_count();
		if(rc == 0 && max_steps > 1)
		{
			rc = run_steps(solver.get(), &data, max_time, check_step, max_steps - 1, param.value("clamp_to_box", false).toBool());
		}
	}
	engine->get_data(&data);
	const double	wall = omp_get_wtime() - t0;
	if(param.value("json", "0").toInt() != 0)
	{
		const size_t steps = data.get_step();
		printf("{\"engine\": \"%s\", \"solver\": \"%s\", \"bodies\": %zu, \"steps\": %zu, \"fcompute_calls\": %zu, \"fcompute_calls_first_step\": %zu, \"time\": %.17g, "
			   "\"wall_s\": %.6f, \"first_step_s\": %.6f, \"ms_per_step_after_first\": %.6f, \"dP\": %s, \"dL\": %s, \"dE\": %s, \"threads\": %d}\n",
			   engine->type_name(), solver->type_name(), data.get_count(), steps, engine->get_compute_count(), calls_first,
			   static_cast<double>(data.get_time()), wall, t_first,
			   (max_steps > 1 && steps > 1) ? (wall - t_first) * 1e3 / static_cast<double>(steps - 1) : wall * 1e3 / static_cast<double>(steps ? steps : 1),
			   num(static_cast<double>(data.get_impulce_err())).c_str(), num(static_cast<double>(data.get_impulce_moment_err())).c_str(),
			   num(static_cast<double>(data.get_energy_err())).c_str(), omp_get_max_threads());
	}
	solver.reset();	// solvers free their buffers through the engine: before the engine goes
	return rc;
}
