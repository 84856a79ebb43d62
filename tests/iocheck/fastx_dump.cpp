// TEST INFRASTRUCTURE ONLY: dumps the records ntedit_b200/csrc/fastx.hpp parses from a file, one "name\tcomment\tsequence" line
// per record, so that tests/test_fastx_cpu.py can compare the reader (and the ByteSource underneath it) with Python.
#include "../../ntedit_b200/csrc/fastx.hpp"

#include <cstdio>
#include <cstdlib>

int
main(int argc, char** argv)
{
	if (argc < 2) {
		return 2;
	}
	ntb::FastxReader r(argv[1], argc > 2 ? (unsigned)std::atoi(argv[2]) : 4u);
	if (!r.ok()) {
		return 3;
	}
	std::fprintf(stderr, "%s\n", r.kind());
	std::string name, comment, seq;
	while (r.next(name, comment, seq)) {
		std::fwrite(name.data(), 1, name.size(), stdout);
		std::fputc('\t', stdout);
		std::fwrite(comment.data(), 1, comment.size(), stdout);
		std::fputc('\t', stdout);
		std::fwrite(seq.data(), 1, seq.size(), stdout);
		std::fputc('\n', stdout);
		seq.clear();
	}
	return 0;
}
