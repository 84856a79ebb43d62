// ntedit-b200-make-bf: builds a btllib KmerBloomFilter file from genome FASTA(s) on the GPU -- the role of
// `ntedit_make_genome_bf` (src/ntedit_make_genome_bf.cpp:49-165), same options.  The k-mers are hashed and inserted by
// the library's insert kernel (K5) through the C ABI; this file only reads FASTA and sizes the filter.
//   --genome F [F ...]  -k K  [--fpr 0.01] [--hashes 3] [-o genome_bf.bf] [--bf BYTES] [--num_elements N] [-t T]
//   extension: --counting writes a KmerCountingBloomFilter8 (8-bit saturating counters) instead
#include "../../include/ntedit_b200.h"
#include "fastx.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

namespace {

// get_bf_size, src/ntedit_make_genome_bf.cpp:41-47 (Broder & Mitzenmacher)
uint64_t
get_bf_size(uint64_t num_elements, double num_hashes, double fpr)
{
	const double r = -num_hashes / std::log(1.0 - std::exp(std::log(fpr) / num_hashes));
	return (uint64_t)(std::ceil((double)num_elements * r) / 8.0);
}

[[noreturn]] void
usage(const char* msg)
{
	if (msg) {
		std::cerr << msg << std::endl;
	}
	std::cerr << "Usage: ntedit-b200-make-bf --genome FASTA [FASTA ...] -k K [--fpr 0.01] [--hashes 3] [-o genome_bf.bf]\n"
	             "                           [--bf BYTES] [--num_elements N] [-t THREADS] [--counting] [--device 0]\n";
	std::exit(1);
}

} // namespace

int
main(int argc, char** argv)
{
	std::vector<std::string> genomes;
	unsigned k = 0, hashes = 3, threads = 12;
	double fpr = 0.01;
	std::string out = "genome_bf.bf";
	uint64_t bf_bytes = 0, num_elements = 0;
	bool have_bf = false, have_ne = false, counting = false;
	int device = 0;
	for (int i = 1; i < argc; i++) {
		const std::string a = argv[i];
		auto need = [&](const char* what) -> const char* {
			if (i + 1 >= argc) {
				usage((std::string(what) + ": missing value").c_str());
			}
			return argv[++i];
		};
		if (a == "--genome") {
			while (i + 1 < argc && argv[i + 1][0] != '-') {
				genomes.push_back(argv[++i]);
			}
		} else if (a == "-k") {
			k = (unsigned)std::strtoul(need("-k"), nullptr, 10);
		} else if (a == "--fpr") {
			fpr = std::strtod(need("--fpr"), nullptr);
		} else if (a == "--hashes") {
			hashes = (unsigned)std::strtoul(need("--hashes"), nullptr, 10);
		} else if (a == "-o") {
			out = need("-o");
		} else if (a == "--bf") {
			bf_bytes = std::strtoull(need("--bf"), nullptr, 10);
			have_bf = true;
		} else if (a == "--num_elements") {
			num_elements = std::strtoull(need("--num_elements"), nullptr, 10);
			have_ne = true;
		} else if (a == "-t") {
			threads = (unsigned)std::strtoul(need("-t"), nullptr, 10);
		} else if (a == "--counting") {
			counting = true;
		} else if (a == "--device") {
			device = std::atoi(need("--device"));
		} else if (a == "-h" || a == "--help") {
			usage(nullptr);
		} else {
			usage(("unknown argument: " + a).c_str());
		}
	}
	if (genomes.empty()) {
		usage("--genome: required");
	}
	if (k == 0) {
		usage("-k: required");
	}
	std::cout << "Parameters:" << std::endl << "\t\t--genome ";
	for (const auto& g : genomes) {
		std::cout << g << " ";
	}
	std::cout << std::endl
	          << "\t\t-t " << threads << std::endl
	          << "\t\t-k " << k << std::endl
	          << "\t\t--fpr " << fpr << std::endl
	          << "\t\t--hashes " << hashes << std::endl
	          << "\t\t-o " << out << std::endl;

	std::string name, comment, seq;
	if (have_bf) {
		std::cout << "\t\t--bf " << bf_bytes << std::endl;
	} else if (have_ne) {
		std::cout << "\t\t--num_elements " << num_elements << std::endl;
		bf_bytes = get_bf_size(num_elements, hashes, fpr);
	} else {
		std::cout << "Calculating BF size based on input genome size" << std::endl;
		uint64_t genome_size = 0;
		for (const auto& g : genomes) {
			ntb::FastxReader r(g);
			if (!r.ok()) {
				std::cerr << "cannot open " << g << std::endl;
				return 1;
			}
			seq.clear();
			while (r.next(name, comment, seq)) {
				genome_size += seq.size();
				seq.clear();
			}
		}
		std::cout << "Genome size (bp): " << genome_size << std::endl;
		bf_bytes = get_bf_size(genome_size, hashes, fpr);
	}
	// btllib's BloomFilter keeps whole 64-bit words: ceil(bytes / sizeof(uint64_t)) with an integer division inside, i.e.
	// the size is rounded DOWN to a multiple of 8 (restated from btllib's constructor; see DESIGN.md, "unpinned")
	bf_bytes = bf_bytes / 8 * 8;
	if (bf_bytes < 8) {
		bf_bytes = 8;
	}
	std::cout << "BF size (bytes): " << bf_bytes << std::endl;

	ntb_filter* f = nullptr;
	if (ntb_filter_create(bf_bytes, k, hashes, counting ? 1 : 0, device, &f) != NTB_OK) {
		std::cerr << "ntedit-b200-make-bf: " << ntb_last_error() << std::endl;
		return 1;
	}
	const size_t BATCH = (size_t)256 << 20;
	std::string bases;
	std::vector<uint64_t> offsets{ 0 };
	auto flush = [&]() -> bool {
		if (offsets.size() > 1) {
			if (ntb_filter_insert(f, bases.data(), offsets.data(), offsets.size() - 1) != NTB_OK) {
				std::cerr << "ntedit-b200-make-bf: " << ntb_last_error() << std::endl;
				return false;
			}
		}
		bases.clear();
		offsets.assign(1, 0);
		return true;
	};
	for (const auto& g : genomes) {
		std::cerr << "Reading " << g << std::endl;
		ntb::FastxReader r(g);
		if (!r.ok()) {
			std::cerr << "cannot open " << g << std::endl;
			return 1;
		}
		for (;;) {
			const size_t before = bases.size();
			if (!r.next(name, comment, bases)) {
				break;
			}
			if (bases.size() - before >= k) { // src/ntedit_make_genome_bf.cpp:153
				bases.push_back('\0');
				offsets.push_back(bases.size());
			} else {
				bases.resize(before);
			}
			if (bases.size() >= BATCH && !flush()) {
				return 1;
			}
		}
	}
	if (!flush()) {
		return 1;
	}
	ntb_filter_info fi;
	if (ntb_filter_get_info(f, &fi) != NTB_OK) {
		std::cerr << "ntedit-b200-make-bf: " << ntb_last_error() << std::endl;
		return 1;
	}
	std::cout << "Bloom filter FPR: " << fi.fpr << std::endl;
	std::cerr << "Saving Bloom filter" << std::endl;
	if (ntb_filter_save(f, out.c_str()) != NTB_OK) {
		std::cerr << "ntedit-b200-make-bf: " << ntb_last_error() << std::endl;
		return 1;
	}
	std::cerr << "Done!" << std::endl;
	ntb_filter_free(f);
	return 0;
}
