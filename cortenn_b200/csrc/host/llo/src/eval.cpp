#include "llo/eval.hpp"
#include "llo/plan.hpp"
#include "llo/session.hpp"

#ifdef LLO_EVAL_HPP

namespace llo
{

static bool is_random (size_t opcode)
{
	return age::RAND_BINO == opcode || age::RAND_UNIF == opcode ||
		age::RAND_NORM == opcode;
}

bool EvalMemo::has_random (ade::iTensor* tens)
{
	auto it = random_.find(tens);
	if (random_.end() != it)
	{
		return it->second;
	}
	bool out = false;
	if (auto func = dynamic_cast<ade::iFunctor*>(tens))
	{
		out = is_random(func->get_opcode().code_);
		for (const ade::MappedTensor& child : func->get_children())
		{
			out = has_random(child.get_tensor().get()) || out;
		}
	}
	random_.emplace(tens, out);
	return out;
}

void Evaluator::visit (ade::iLeaf* leaf)
{
	const ade::Shape& shape = leaf->shape();
	if (auto var = dynamic_cast<Variable*>(leaf))
	{
		// device mirror, uploaded / converted only when the leaf changed
		out_ = GenericData();
		out_.shape_ = shape;
		out_.dtype_ = dtype_;
		out_.device_ = var->device_data(dtype_);
		return;
	}
	const char* data = (const char*) ((const ade::iLeaf*) leaf)->data();
	age::_GENERATED_DTYPE dtype = (age::_GENERATED_DTYPE) leaf->type_code();
	out_ = GenericData(shape, dtype_);
	out_.copyover(data, dtype);
}

void Evaluator::visit (ade::iFunctor* func)
{
	size_t code = func->get_opcode().code_;
	age::_GENERATED_OPCODE opcode = (age::_GENERATED_OPCODE) code;
	auto& done = memo_->of(dtype_);
	const bool fresh = memo_->has_random(func);
	if (false == fresh)
	{
		auto it = done.find(func);
		if (done.end() != it)
		{
			out_ = it->second;
			return;
		}
	}

	const ade::ArgsT& children = func->get_children();
	size_t nargs = children.size();
	if (age::RAND_BINO == opcode && 2 != nargs)
	{
		logs::fatalf("cannot RAND_BINO without exactly 2 arguments: "
			"using %d arguments", (int) nargs);
	}
	DataArgsT argdata(nargs);
	for (size_t i = 0; i < nargs; ++i)
	{
		// RAND_BINO's probability is evaluated as DOUBLE whatever the
		// requested type (llo/eval.hpp:50-74)
		age::_GENERATED_DTYPE childtype =
			(age::RAND_BINO == opcode && 1 == i) ? age::DOUBLE : dtype_;
		Evaluator evaler(childtype, memo_);
		children[i].get_tensor()->accept(evaler);
		argdata[i] = DataArg{
			evaler.out_.device_,
			evaler.out_.shape_,
			children[i].get_coorder(),
			children[i].map_io(),
		};
	}

	out_ = GenericData(func->shape(), dtype_);
	op_exec(opcode, out_.dtype_, out_.device_.get(), out_.shape_, argdata);
	if (false == fresh)
	{
		done.emplace(func, out_);
	}
}

GenericData eval_device (ade::TensptrT tens, age::_GENERATED_DTYPE dtype)
{
	if (0 != get_option("plan"))
	{
		return planned_eval(tens, dtype);
	}
	Evaluator eval(dtype);
	tens->accept(eval);
	return eval.out_;
}

std::vector<GenericData> eval_device_many (ade::TensT roots,
	age::_GENERATED_DTYPE dtype)
{
	if (0 != get_option("plan"))
	{
		return planned_eval_many(roots, dtype);
	}
	std::shared_ptr<EvalMemo> memo = std::make_shared<EvalMemo>();
	std::vector<GenericData> outs;
	for (ade::TensptrT& root : roots)
	{
		Evaluator eval(dtype, memo);
		root->accept(eval);
		outs.push_back(eval.out_);
	}
	return outs;
}

GenericData eval (ade::TensptrT tens, age::_GENERATED_DTYPE dtype)
{
	GenericData out = eval_device(tens, dtype);
	if (out.shape_.n_elems() > 0)
	{
		out.fetch();
	}
	return out;
}

}

#endif
