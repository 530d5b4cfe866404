// CPU check of opt::TargetPruner (csrc/host/opt/shear.hpp) against the golden tree of the
// reference's own test, opt/test/test_shear.cpp:100-178 (SHEAR.Prune): the same eight
// functors over three scalar leaves, the same "drop the flagged arguments of killable
// nodes" rule, the same expected rendering.  Prints the tree; exit code 0 iff it matches.
#include <iostream>
#include <sstream>

#include "dbg/ade.hpp"
#include "opt/shear.hpp"

struct ScalarLeaf final : public ade::iLeaf
{
	const ade::Shape& shape (void) const override { return shape_; }
	std::string to_string (void) const override { return shape_.to_string(); }
	void* data (void) override { return &value_; }
	const void* data (void) const override { return &value_; }
	size_t type_code (void) const override { return 0; }
	double value_ = 0;
	ade::Shape shape_;
};

static ade::TensptrT node (const char* name, size_t code, std::vector<ade::TensptrT> children)
{
	ade::ArgsT args;
	for (ade::TensptrT& child : children)
	{
		args.push_back(ade::identity_map(child));
	}
	return ade::TensptrT(ade::Functor::get(ade::Opcode{name, code}, args));
}

int main (void)
{
	ade::TensptrT leaf(new ScalarLeaf()), leaf2(new ScalarLeaf()), leaf3(new ScalarLeaf());
	ade::TensptrT inner = node("binary", 1, {leaf3, node("killable", 0, {leaf})});
	ade::TensptrT other = node("binary", 1,
		{node("killable", 0, {leaf2}), node("not_killable", 2, {leaf})});
	ade::TensptrT top = node("not_killable", 2,
		{node("killable", 0, {node("killable", 0, {inner})}), other});

	opt::TargetPruner<bool> pruner(true,
		[&](ade::iLeaf* l) { return l == leaf.get() || l == leaf2.get(); },
		[=](ade::iFunctor* f, std::unordered_set<size_t> flagged, ade::ArgsT args)
		{
			ade::Opcode opcode = f->get_opcode();
			if (opcode.code_ < 2)
			{
				for (size_t index : flagged)
				{
					args.erase(args.begin() + index);
				}
			}
			return args.empty() ? leaf : ade::TensptrT(ade::Functor::get(opcode, args));
		});
	ade::TensptrT root = pruner.prune(top);

	PrettyEquation artist;
	artist.showshape_ = true;
	artist.labels_ = {{leaf.get(), "leaf"}, {leaf2.get(), "leaf2"}, {leaf3.get(), "leaf3"}};
	std::stringstream got;
	artist.print(got, root);
	std::cout << got.str();

	const std::string s = "[1\\1\\1\\1\\1\\1\\1\\1])";
	const char* lines[] = {
		"(not_killable", "`--(killable", "|   `--(killable", "|       `--(binary",
		"|           `--(leaf3=", "|           `--(killable", "|               `--(leaf=",
		"`--(binary", "`--(killable", "|   `--(leaf2=", "`--(not_killable", "`--(leaf=",
	};
	std::string want;
	for (const char* line : lines)
	{
		want += std::string(line) + s + "\n";
	}
	// compare with leading / trailing blanks of every line removed, as the reference does
	std::string have, line;
	while (std::getline(got, line))
	{
		size_t a = line.find_first_not_of(" \t"), b = line.find_last_not_of(" \t");
		if (std::string::npos != a)
		{
			have += line.substr(a, b - a + 1) + "\n";
		}
	}
	return have == want ? 0 : 1;
}
