// A C++ caller of the reference's API against libllo.so + libllo_cuda.so (no Python):
// builds a two-layer sigmoid network with the generated age:: builders, derives and
// optimises its gradients, previews the plan, constructs the sharded evaluator, and
// evaluates -- on a B200 the gradients are checked against central differences of the
// loss; without one, llo::eval must throw ("no CPU fallback") instead of computing.
// Prints one line per step; exit code 0 iff every check holds.
#include <cmath>
#include <iostream>
#include <random>

#include "llo/device.hpp"
#include "llo/eval.hpp"
#include "llo/generated/api.hpp"
#include "llo/optimize.hpp"
#include "llo/plan.hpp"
#include "llo/sharded.hpp"
#include "llo/zprune.hpp"

static std::vector<double> uniform (size_t n, double lo, double hi, unsigned seed)
{
	std::mt19937_64 gen(seed);
	std::uniform_real_distribution<double> dist(lo, hi);
	std::vector<double> out(n);
	for (double& v : out) { v = dist(gen); }
	return out;
}

static ade::TensptrT sigmoid (ade::TensptrT z)
{
	ade::TensptrT one(llo::get_scalar<double>(1.0, z->shape()));
	return age::div(one, age::add(one, age::exp(age::neg(z))));
}

int main (void)
{
	const uint8_t batch = 12, in = 6, hidden = 5, out = 3;
	// ade shapes list the fastest dimension first: [columns, rows]
	llo::VarptrT x = llo::get_variable<double>(uniform(batch * in, 0.1, 1.0, 1),
		ade::Shape({in, batch}), "x");
	llo::VarptrT w1 = llo::get_variable<double>(uniform(in * hidden, -1.0, 1.0, 2),
		ade::Shape({hidden, in}), "w1");
	llo::VarptrT w2 = llo::get_variable<double>(uniform(hidden * out, -1.0, 1.0, 3),
		ade::Shape({out, hidden}), "w2");
	llo::VarptrT y = llo::get_variable<double>(uniform(batch * out, 0.0, 1.0, 4),
		ade::Shape({out, batch}), "y");
	ade::TensptrT h = sigmoid(age::matmul(ade::TensptrT(x), ade::TensptrT(w1)));
	ade::TensptrT p = sigmoid(age::matmul(h, ade::TensptrT(w2)));
	ade::TensptrT err = age::sub(p, ade::TensptrT(y));
	ade::TensptrT loss = age::reduce_sum(age::pow(err,
		ade::TensptrT(llo::get_scalar<double>(2.0, err->shape()))));

	ade::TensT roots = {loss, llo::derive(loss, w1.get()), llo::derive(loss, w2.get())};
	llo::OptReport report;
	ade::TensT fast = llo::optimize(roots, llo::OPT_ALL, &report);
	std::cout << "optimize: functors " << report.functors_before_ << " -> "
		<< report.functors_after_ << ", div_cancel " << report.div_cancel_ << "\n";
	if (report.functors_after_ >= report.functors_before_ || 0 == report.div_cancel_)
	{
		return 1;
	}
	llo::PlanStats stats = llo::plan_preview(fast, age::DOUBLE);
	std::cout << "plan preview: nodes " << stats.nodes_ << ", generic " << stats.generic_ << "\n";
	if (0 != stats.generic_)
	{
		return 2;
	}
	llo::ShardedEvaluator sharded(loss, {ade::TensptrT(w1), ade::TensptrT(w2)});
	if (3 != sharded.roots_.size())
	{
		return 3;
	}

	if (false == llo::device::available())
	{
		try
		{
			llo::eval(loss, age::DOUBLE);
		}
		catch (const std::runtime_error& e)
		{
			std::cout << "no GPU: " << e.what() << "\n";
			return std::string(e.what()).find("no CPU fallback") != std::string::npos ? 0 : 4;
		}
		return 5; // computed something without a GPU
	}

	// on a B200: gradients (sharded evaluator, one rank) against central differences
	std::vector<llo::GenericData> got = sharded.grad_eval(age::DOUBLE, true);
	double base = *(double*) got[0].data_.get();
	std::cout << "loss " << base << "\n";
	llo::VarptrT params[2] = {w1, w2};
	for (int k = 0; k < 2; ++k)
	{
		size_t n = params[k]->shape().n_elems();
		std::vector<double> values((double*) params[k]->data(), (double*) params[k]->data() + n);
		const double* grad = (const double*) got[1 + k].data_.get();
		for (size_t i = 0; i < n; i += 3)
		{
			const double eps = 1e-6;
			double lr[2];
			for (int side = 0; side < 2; ++side)
			{
				std::vector<double> moved = values;
				moved[i] += (0 == side ? eps : -eps);
				*params[k] = llo::GenericRef{(char*) moved.data(), params[k]->shape(), age::DOUBLE};
				llo::GenericData l = llo::eval(loss, age::DOUBLE);
				lr[side] = *(double*) l.data_.get();
			}
			double numeric = (lr[0] - lr[1]) / (2 * eps);
			if (std::fabs(numeric - grad[i]) > 1e-6 * std::max(1.0, std::fabs(numeric)))
			{
				std::cout << "gradient " << k << "[" << i << "]: " << grad[i]
					<< " vs central difference " << numeric << "\n";
				return 6;
			}
		}
		*params[k] = llo::GenericRef{(char*) values.data(), params[k]->shape(), age::DOUBLE};
	}
	std::cout << "gradients match central differences\n";
	return 0;
}
