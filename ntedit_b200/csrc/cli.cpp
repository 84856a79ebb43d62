// ntedit-b200: the host driver of the B200 path with ntEdit's command line and file formats.
//
// Plays the roles of the reference's main() (ntedit.cpp:2276-2600: option parsing, parameter clamps, filter load,
// banners) and readAndCorrect() (ntedit.cpp:2154-2259: gz/plain FASTA in through a kseq-compatible reader, the three
// output files with their headers).  Where the reference hands one contig at a time to an OpenMP thread
// (ntedit.cpp:2213-2252), this driver packs contigs into batches and hands each batch to the GPU through the C ABI
// (ntb_polish_batch); with --gpus N batches go round-robin to N devices, the filter replicated on each.  Output is
// always in input order (= the reference's `-t 1` order).  No hashing and no filter probe happens in this file.
#include "../../include/ntedit_b200.h"
#include "fastx.hpp"
#include "writer.hpp"

#include <getopt.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <future>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#define PROGRAM "ntedit-b200"

namespace {

const char VERSION_MESSAGE[] = PROGRAM " (ntEdit v2.1.1 hot path on B200)\n";

const char USAGE_MESSAGE[] =
    "Usage: " PROGRAM " -f DRAFT -r BLOOM [options]\n"
    "  drop-in for `ntedit` (bcgsc/ntEdit v2.1.1); same options, same output files\n"
    " -t,  number of host threads (formatting / replay) [4]\n"
    " -f,  draft genome assembly (FASTA, Multi-FASTA, and/or gzipped compatible), REQUIRED\n"
    " -r,  Bloom filter file (btllib KmerBloomFilter / KmerCountingBloomFilter8), REQUIRED\n"
    " -e,  secondary Bloom filter with k-mers to reject (optional)\n"
    " -b,  output file prefix (optional)\n"
    " -z,  minimum contig length [100]\n"
    " -i,  maximum number of insertion bases to try, range 0-5 [5]\n"
    " -d,  maximum number of deletions bases to try, range 0-10 [5]\n"
    " -x,  k/x ratio for the number of k-mers that should be missing [5.000]\n"
    " -y,  k/y ratio for the number of edited k-mers that should be present [9.000]\n"
    " -X,  ratio of number of k-mers in the k subset that should be missing in order to attempt fix (higher=stringent)\n"
    " -Y,  ratio of number of k-mers in the k subset that should be present to accept an edit (higher=stringent)\n"
    " -c,  cap for the number of base insertions that can be made at one position (accepted, overridden by k*1.5 as in ntedit)\n"
    " -j,  controls size of k-mer subset [3]\n"
    " -m,  mode of editing, range 0-2 [0]\n"
    " -s,  SNV mode (-s 1 = yes, default = 0, no)\n"
    " -l,  input VCF file with annotated variants (e.g., clinvar.vcf, optional)\n"
    " -a,  soft masks missing k-mer positions having no fix (-a 1 = yes, default = 0, no)\n"
    " -p,  minimum k-mer coverage threshold (CBF only) [1]\n"
    " -q,  maximum k-mer coverage threshold (CBF only) [255]\n"
    " -k,  k-mer size (optional; taken from the Bloom filter header, checked when given)\n"
    " -v,  verbose (accepted; per-site tracing is not reproduced)\n"
    "      --gpus N          shard batches of contigs over the first N GPUs [1]\n"
    "      --batch_bases N   bases per device batch [1073741824]\n"
    "      --help, --version\n";

struct Opt
{
	unsigned nthreads = 4;
	std::string draft, vcf, bloom, bloomrep, prefix;
	ntb_params p;
	unsigned insertion_cap = 0;
	int verbose = 0;
	unsigned k_given = 0;
	int gpus = 1;
	uint64_t batch_bases = 1ull << 30;
};

enum
{
	OPT_HELP = 1,
	OPT_VERSION,
	OPT_GPUS,
	OPT_BATCH
};

const char shortopts[] = "t:f:s:k:z:b:r:v:d:i:X:Y:x:y:m:c:j:s:e:a:l:p:q:";

const struct option longopts[] = { { "threads", required_argument, nullptr, 't' },
	                               { "draft_file", required_argument, nullptr, 'f' },
	                               { "k", required_argument, nullptr, 'k' },
	                               { "minimum_contig_length", required_argument, nullptr, 'z' },
	                               { "maximum_insertions", required_argument, nullptr, 'i' },
	                               { "maximum_deletions", required_argument, nullptr, 'd' },
	                               { "insertion_cap", required_argument, nullptr, 'c' },
	                               { "edit_threshold", required_argument, nullptr, 'y' },
	                               { "missing_threshold", required_argument, nullptr, 'x' },
	                               { "edit_ratio", required_argument, nullptr, 'Y' },
	                               { "missing_ratio", required_argument, nullptr, 'X' },
	                               { "jump", required_argument, nullptr, 'j' },
	                               { "bloom_filename", required_argument, nullptr, 'r' },
	                               { "bloomrep_filename", required_argument, nullptr, 'e' },
	                               { "outfile_prefix", required_argument, nullptr, 'b' },
	                               { "mode", required_argument, nullptr, 'm' },
	                               { "snv", required_argument, nullptr, 's' },
	                               { "vcf_file", required_argument, nullptr, 'l' },
	                               { "mask", required_argument, nullptr, 'a' },
	                               { "verbose", required_argument, nullptr, 'v' },
	                               { "minimum_kmer_coverage", required_argument, nullptr, 'p' },
	                               { "maximum_kmer_coverage", required_argument, nullptr, 'q' },
	                               { "gpus", required_argument, nullptr, OPT_GPUS },
	                               { "batch_bases", required_argument, nullptr, OPT_BATCH },
	                               { "help", no_argument, nullptr, OPT_HELP },
	                               { "version", no_argument, nullptr, OPT_VERSION },
	                               { nullptr, 0, nullptr, 0 } };

void
assert_readable(const std::string& path)
{
	std::ifstream f(path);
	if (!f.good()) {
		std::cerr << PROGRAM ": error: cannot read `" << path << "'\n";
		std::exit(EXIT_FAILURE);
	}
}

std::string
basename_of(const std::string& p)
{
	return p.substr(p.find_last_of("/\\") + 1);
}

std::string
now_str()
{
	time_t raw;
	time(&raw);
	return ctime(&raw);
}

[[noreturn]] void
die_ntb(const char* what)
{
	std::cerr << PROGRAM ": error: " << what << ": " << ntb_last_error() << "\n";
	std::exit(EXIT_FAILURE);
}

// vcf_entry_to_map, ntedit.cpp:2261-2274
void
vcf_entry_to_map(const std::string& line, ntb::ClinvarMap& m)
{
	std::vector<std::string> tok;
	size_t s = 0;
	for (;;) {
		const size_t e = line.find('\t', s);
		tok.push_back(line.substr(s, e == std::string::npos ? std::string::npos : e - s));
		if (e == std::string::npos) {
			break;
		}
		s = e + 1;
	}
	// std::sregex_token_iterator drops trailing empty fields
	while (!tok.empty() && tok.back().empty()) {
		tok.pop_back();
	}
	if (tok.size() >= 8) {
		m[tok[0] + ">" + tok[3] + tok[1] + tok[4]] = tok[7];
	}
}

void
load_clinvar(const std::string& path, ntb::ClinvarMap& m)
{
	gzFile fp = gzopen(path.c_str(), "r"); // transparently reads plain and gzipped files
	if (!fp) {
		std::cerr << "Unable to open file" << std::endl;
		return;
	}
	std::string line;
	char buf[1 << 16];
	while (gzgets(fp, buf, sizeof buf)) {
		line += buf;
		if (!line.empty() && line.back() == '\n') {
			line.pop_back();
			vcf_entry_to_map(line, m);
			line.clear();
		}
	}
	if (!line.empty()) {
		vcf_entry_to_map(line, m);
	}
	gzclose(fp);
}

// Growable byte buffer in page-locked host memory (ntb_host_alloc): the batch text the reader appends to and the polishing
// call uploads from.  Buffers are recycled through a small pool: pinning a gigabyte costs hundreds of milliseconds.
class PinnedBuf
{
  public:
	PinnedBuf() = default;
	PinnedBuf(const PinnedBuf&) = delete;
	PinnedBuf& operator=(const PinnedBuf&) = delete;
	~PinnedBuf() { release(); }

	size_t size() const { return size_; }
	bool empty() const { return size_ == 0; }
	char& operator[](size_t i) { return data_[i]; }
	char* data() { return data_; }
	char& back() { return data_[size_ - 1]; }
	void pop_back() { size_--; }
	void clear() { size_ = 0; }
	void push_back(char c)
	{
		if (size_ == cap_) {
			grow(size_ + 1);
		}
		data_[size_++] = c;
	}
	void append(const char* p, size_t n)
	{
		if (size_ + n > cap_) {
			grow(size_ + n);
		}
		std::memcpy(data_ + size_, p, n);
		size_ += n;
	}
	void reserve(size_t n)
	{
		if (n > cap_) {
			grow(n);
		}
	}

  private:
	void grow(size_t need)
	{
		size_t cap = cap_ ? cap_ : (size_t)1 << 20;
		while (cap < need) {
			cap += cap / 2;
		}
		char* p = take(cap);
		if (size_) {
			std::memcpy(p, data_, size_);
		}
		release();
		data_ = p;
		cap_ = cap;
	}
	struct Pool
	{
		std::mutex lock;
		std::vector<std::pair<char*, size_t>> free_list;
	};
	static Pool& pool()
	{
		static Pool p;
		return p;
	}
	// a pooled buffer of at least `cap` bytes (cap_ is set by the caller to what was asked for: pooled ones may be larger)
	static char* take(size_t& cap)
	{
		{
			std::lock_guard<std::mutex> g(pool().lock);
			for (size_t i = 0; i < pool().free_list.size(); i++) {
				if (pool().free_list[i].second >= cap) {
					char* p = pool().free_list[i].first;
					cap = pool().free_list[i].second;
					pool().free_list.erase(pool().free_list.begin() + (long)i);
					return p;
				}
			}
		}
		char* p = (char*)ntb_host_alloc(cap);
		if (!p) {
			std::cerr << PROGRAM ": error: out of page-locked host memory (" << cap << " bytes)\n";
			std::exit(EXIT_FAILURE);
		}
		return p;
	}
	void release()
	{
		if (data_) {
			std::lock_guard<std::mutex> g(pool().lock);
			if (pool().free_list.size() < 6) {
				pool().free_list.emplace_back(data_, cap_);
			} else {
				ntb_host_free(data_);
			}
		}
		data_ = nullptr;
		size_ = cap_ = 0;
	}
	char* data_ = nullptr;
	size_t size_ = 0, cap_ = 0;
};

struct Batch
{
	std::vector<std::string> names;
	PinnedBuf bases; // contigs end to end, each followed by its NUL
	std::vector<uint64_t> offsets{ 0 };
};

// bounded hand-over between the pipeline's threads (reader -> dispatcher -> writer)
template<class T>
class Channel
{
  public:
	explicit Channel(size_t cap) : cap_(cap) {}
	void push(T v)
	{
		std::unique_lock<std::mutex> lk(m_);
		space_.wait(lk, [this]() { return q_.size() < cap_; });
		q_.push_back(std::move(v));
		lk.unlock();
		data_.notify_one();
	}
	void close()
	{
		{
			std::lock_guard<std::mutex> g(m_);
			closed_ = true;
		}
		data_.notify_all();
	}
	bool pop(T& v) // false: closed and drained
	{
		std::unique_lock<std::mutex> lk(m_);
		data_.wait(lk, [this]() { return !q_.empty() || closed_; });
		if (q_.empty()) {
			return false;
		}
		v = std::move(q_.front());
		q_.pop_front();
		lk.unlock();
		space_.notify_one();
		return true;
	}

  private:
	size_t cap_;
	std::mutex m_;
	std::condition_variable data_, space_;
	std::deque<T> q_;
	bool closed_ = false;
};

struct BatchOut
{
	std::string fa, tsv, vcf;
	uint64_t contigs = 0, bases = 0, edits = 0;
};

// polish one batch on `device` and format it (contigs formatted in parallel, concatenated in input order)
// counts the batches whose polishing call is running: one per device at a time
struct DeviceSlots
{
	std::mutex m;
	std::condition_variable cv;
	size_t busy = 0;
	void acquire(size_t limit)
	{
		std::unique_lock<std::mutex> lk(m);
		cv.wait(lk, [&]() { return busy < limit; });
		busy++;
	}
	void release()
	{
		{
			std::lock_guard<std::mutex> g(m);
			busy--;
		}
		cv.notify_one();
	}
};

BatchOut
process_batch(std::unique_ptr<Batch> b, ntb_filter* bloom, ntb_filter* rep, const Opt& opt, const ntb::ClinvarMap* cv, DeviceSlots* slots)
{
	BatchOut out;
	ntb_result* res = nullptr;
	if (ntb_polish_batch(bloom, rep, &opt.p, &b->bases[0], b->offsets.data(), b->names.size(), &res) != NTB_OK) {
		std::cerr << PROGRAM ": error: polishing failed in the batch of " << b->names.size() << " sequence(s) from `" << b->names.front() << "' to `"
		          << b->names.back() << "'; the output files are incomplete\n";
		die_ntb("ntb_polish_batch");
	}
	slots->release(); // the device is free for the next batch while this one is formatted
	const size_t n = b->names.size();
	std::vector<std::string> fa(n), tsv(n), vcf(n);
	std::atomic<size_t> next(0);
	auto work = [&]() {
		for (;;) {
			const size_t c = next.fetch_add(1);
			if (c >= n) {
				break;
			}
			int polished = 0;
			const ntb_node* nodes = nullptr;
			const ntb_srec* recs = nullptr;
			uint64_t nn = 0, nr = 0;
			ntb_result_contig(res, c, &polished, &nodes, &nn, &recs, &nr);
			if (!polished) {
				continue; // shorter than -z: dropped from all outputs, ntedit.cpp:2242-2245
			}
			ntb::format_contig(b->names[c], &b->bases[b->offsets[c]], nodes, (size_t)nn, recs, (size_t)nr, opt.p.snv != 0, cv, &fa[c], &tsv[c],
			                   &vcf[c]);
		}
	};
	const unsigned nt = (unsigned)std::max<size_t>(1, std::min<size_t>(opt.nthreads, n));
	std::vector<std::thread> pool;
	for (unsigned i = 1; i < nt; i++) {
		pool.emplace_back(work);
	}
	work();
	for (auto& t : pool) {
		t.join();
	}
	for (size_t c = 0; c < n; c++) {
		out.fa += fa[c];
		out.tsv += tsv[c];
		out.vcf += vcf[c];
	}
	ntb_stats st;
	ntb_result_stats(res, &st);
	out.contigs = st.contigs;
	out.bases = st.bases;
	out.edits = st.edits;
	ntb_result_free(res);
	return out;
}

template<class T>
bool
parse(const char* s, T& v)
{
	std::istringstream arg(s ? s : "");
	arg >> v;
	return arg.eof() && !arg.fail();
}

} // namespace

int
main(int argc, char** argv)
{
	Opt opt;
	ntb_params_init(&opt.p);
	bool die = false;
	for (int c; (c = getopt_long(argc, argv, shortopts, longopts, nullptr)) != -1;) {
		bool ok = true;
		switch (c) {
		case '?': die = true; break;
		case 't': ok = parse(optarg, opt.nthreads); break;
		case 'f': ok = parse(optarg, opt.draft); break;
		case 'z': ok = parse(optarg, opt.p.min_contig_len); break;
		case 'b': ok = parse(optarg, opt.prefix); break;
		case 'r': ok = parse(optarg, opt.bloom); break;
		case 'e': ok = parse(optarg, opt.bloomrep); break;
		case 'd': ok = parse(optarg, opt.p.max_deletions); break;
		case 'i': ok = parse(optarg, opt.p.max_insertions); break;
		case 'x': ok = parse(optarg, opt.p.missing_threshold); break;
		case 'y': ok = parse(optarg, opt.p.edit_threshold); break;
		case 'X':
			ok = parse(optarg, opt.p.missing_ratio);
			opt.p.use_ratio = 1;
			break;
		case 'Y':
			ok = parse(optarg, opt.p.edit_ratio);
			opt.p.use_ratio = 1;
			break;
		case 'c': ok = parse(optarg, opt.insertion_cap); break;
		case 'j': ok = parse(optarg, opt.p.jump); break;
		case 'm': ok = parse(optarg, opt.p.mode); break;
		case 's': ok = parse(optarg, opt.p.snv); break;
		case 'l': ok = parse(optarg, opt.vcf); break;
		case 'a': ok = parse(optarg, opt.p.mask); break;
		case 'v': ok = parse(optarg, opt.verbose); break;
		case 'p': ok = parse(optarg, opt.p.min_threshold); break;
		case 'q': ok = parse(optarg, opt.p.max_threshold); break;
		case 'k': ok = parse(optarg, opt.k_given); break; // the reference rejects -k (no case for it); we check it against the filter
		case OPT_GPUS: ok = parse(optarg, opt.gpus); break;
		case OPT_BATCH: ok = parse(optarg, opt.batch_bases); break;
		case OPT_HELP: std::cerr << USAGE_MESSAGE; return EXIT_SUCCESS;
		case OPT_VERSION: std::cerr << VERSION_MESSAGE; return EXIT_SUCCESS;
		default: break;
		}
		if (!ok) {
			std::cerr << PROGRAM ": invalid option: `-" << (char)c << (optarg ? optarg : "") << "'\n";
			return EXIT_FAILURE;
		}
	}

	std::cout << "---------- initializing                             : " << now_str();
	if (opt.draft.empty()) {
		std::cerr << PROGRAM ": error: need to specify assembly draft file (-f)\n";
		die = true;
	} else {
		assert_readable(opt.draft);
	}
	if (opt.bloom.empty()) {
		std::cerr << PROGRAM ": error: need to specify the Bloom filter file (-r)\n";
		die = true;
	} else {
		assert_readable(opt.bloom);
	}
	if (!opt.bloomrep.empty()) {
		assert_readable(opt.bloomrep);
	}
	if (die) {
		std::cerr << "Try `" << PROGRAM << " --help' for more information.\n";
		return EXIT_FAILURE;
	}
	if (opt.p.snv) { // ntedit.cpp:2411-2420
		opt.p.max_insertions = 0;
		opt.p.max_deletions = 0;
		std::cerr << "\nSNV mode ON\nTracking all single-base variants\nNote: -i and -d both set to 0 when -s is set to 1\n\n";
	}
	const int ndev_avail = ntb_device_count();
	if (ndev_avail <= 0) {
		std::cerr << PROGRAM ": error: no CUDA device available (this build has no CPU fallback)\n";
		return EXIT_FAILURE;
	}
	if (opt.gpus < 1 || opt.gpus > ndev_avail) {
		std::cerr << PROGRAM ": error: --gpus " << opt.gpus << " but " << ndev_avail << " device(s) present\n";
		return EXIT_FAILURE;
	}
	if (opt.batch_bases < 1024) {
		opt.batch_bases = 1024;
	}

	// ---- filters: one replica per device (ntedit.cpp:2438-2461)
	std::cout << "---------- loading Bloom filter from file           : " << now_str() << "\n";
	std::vector<ntb_filter*> bloom((size_t)opt.gpus, nullptr), rep((size_t)opt.gpus, nullptr);
	// the file is read once; the other devices get device-to-device replicas (NVLink peer copies), all at the same time
	auto load_replicated = [&](const std::string& path, std::vector<ntb_filter*>& fs, const char* what) -> bool {
		if (ntb_filter_load(path.c_str(), 0, &fs[0]) != NTB_OK) {
			std::cerr << PROGRAM ": error: " << what << " is incorrect: " << ntb_last_error() << "\n";
			return false;
		}
		std::vector<std::future<std::string>> copies;
		for (int d = 1; d < opt.gpus; d++) {
			copies.push_back(std::async(std::launch::async, [&fs, d]() -> std::string {
				return ntb_filter_replicate(fs[0], d, &fs[(size_t)d]) == NTB_OK ? std::string() : std::string(ntb_last_error());
			}));
		}
		bool ok = true;
		for (auto& c : copies) {
			const std::string err = c.get();
			if (!err.empty()) {
				std::cerr << PROGRAM ": error: cannot replicate the Bloom filter: " << err << "\n";
				ok = false;
			}
		}
		return ok;
	};
	if (!load_replicated(opt.bloom, bloom, "Bloom filter file supplied (-r)")) {
		return EXIT_FAILURE;
	}
	ntb_filter_info fi;
	if (ntb_filter_get_info(bloom[0], &fi) != NTB_OK) {
		die_ntb("ntb_filter_get_info");
	}
	if (fi.hash_num == 0) {
		std::cerr << PROGRAM ": error: Bloom filter file supplied (-r) is incorrect.\n";
		return EXIT_FAILURE;
	}
	if (opt.k_given && opt.k_given != fi.k) {
		std::cerr << PROGRAM ": error: -k " << opt.k_given << " does not match the Bloom filter's k-mer size (" << fi.k << ")\n";
		return EXIT_FAILURE;
	}
	if (!fi.counting && opt.p.min_threshold != 1) { // ntedit.cpp:2453-2458
		std::cerr << PROGRAM ": warning: Bloom filter is not counting, min k-mer presence threshold will be set to 1.\n";
		opt.p.min_threshold = 1;
	}
	auto print_details = [](const ntb_filter_info& f) { // BFWrapper::print_details, ntedit.cpp:387-395
		std::cout << "BLOOM::\tcounting: " << (f.counting ? "YES" : "NO") << "\tsize: " << f.bytes << "\tnumber hash functions: " << f.hash_num
		          << "\tkmer size: " << f.k << "\tFPR: " << f.fpr << std::endl;
	};
	print_details(fi);

	std::cout << "\n---------- verifying parameters                     : " << now_str();
	if ((opt.p.max_insertions == 0 && opt.p.max_deletions > 0) || (opt.p.max_insertions == 1 && opt.p.max_deletions > 1)) {
		std::cerr << PROGRAM ": warning: i and d parameter combination is not possible; d was set to the value of i.\n";
		opt.p.max_deletions = opt.p.max_insertions;
	}
	if (opt.p.max_insertions > 5) {
		std::cerr << PROGRAM ": warning: i parameter too high, adjusting to maximum -i 5";
		opt.p.max_insertions = 5;
	}
	if (opt.p.max_deletions > 10) {
		std::cerr << PROGRAM ": warning: d parameter too high, adjusting to maximum -d 10";
		opt.p.max_deletions = 10;
	}
	if (opt.prefix.empty()) { // ntedit.cpp:2496-2502
		std::ostringstream o;
		o << basename_of(opt.draft) << "_k" << fi.k << "_z" << opt.p.min_contig_len << "_r" << basename_of(opt.bloom) << "_i" << opt.p.max_insertions
		  << "_d" << opt.p.max_deletions << "_m" << opt.p.mode;
		opt.prefix = o.str();
	}
	std::cout << "\nrunning : " << PROGRAM << "\n -f " << basename_of(opt.draft) << "\n -k " << fi.k << "\n -z " << opt.p.min_contig_len << "\n -b "
	          << opt.prefix << "\n -r " << basename_of(opt.bloom) << "\n -e " << basename_of(opt.bloomrep) << "\n -i " << opt.p.max_insertions << "\n -d "
	          << opt.p.max_deletions;
	if (opt.p.use_ratio) {
		std::cout << "\n -X " << opt.p.missing_ratio << "\n -Y " << opt.p.edit_ratio;
	} else {
		std::cout << "\n -x " << opt.p.missing_threshold << "\n -y " << opt.p.edit_threshold;
	}
	std::cout << "\n -j " << opt.p.jump << "\n -m " << opt.p.mode << "\n -s " << opt.p.snv << "\n -l " << basename_of(opt.vcf) << "\n -a " << opt.p.mask
	          << "\n -t " << opt.nthreads << "\n -v " << opt.verbose << "\n --gpus " << opt.gpus << "\n"
	          << std::endl;
	if (fi.counting) {
		std::cout << " -p " << opt.p.min_threshold << "\n -q " << opt.p.max_threshold << "\n" << std::endl;
	}

	ntb::ClinvarMap clinvar;
	if (!opt.vcf.empty()) {
		assert_readable(opt.vcf);
		load_clinvar(opt.vcf, clinvar);
	}

	if (!opt.bloomrep.empty()) { // ntedit.cpp:2566-2590
		std::cout << "---------- loading secondary Bloom filter from file : " << now_str() << "\n";
		if (!load_replicated(opt.bloomrep, rep, "secondary Bloom filter file supplied (-e)")) {
			return EXIT_FAILURE;
		}
		ntb_filter_info ri;
		ntb_filter_get_info(rep[0], &ri);
		if (ri.k != fi.k) {
			std::cerr << PROGRAM ": error: secondary Bloom filter k size (" << ri.k << ") is different than main Bloom filter k size (" << fi.k << ")\n";
			return EXIT_FAILURE;
		}
		print_details(ri);
		std::cout << "\n";
	}
	std::cout << "---------- reading/processing input sequence        : " << now_str();

	// ---- readAndCorrect, ntedit.cpp:2154-2259
	ntb::FastxReader reader(opt.draft);
	if (!reader.ok()) {
		std::cerr << PROGRAM ": error: cannot open `" << opt.draft << "'\n";
		return EXIT_FAILURE;
	}
	std::ofstream dfout(opt.prefix + "_edited.fa", std::ios::binary);
	std::ofstream rfout(opt.prefix + "_changes.tsv", std::ios::binary);
	std::ofstream vfout(opt.prefix + "_variants.vcf", std::ios::binary);
	if (!dfout || !rfout || !vfout) {
		std::cerr << PROGRAM ": error: cannot open output files with prefix `" << opt.prefix << "'\n";
		return EXIT_FAILURE;
	}
	rfout << ntb::tsv_header(fi.k, opt.p.jump, fi.counting != 0);
	vfout << ntb::vcf_header("ntEdit v2.1.1", opt.draft); // the reference's PROGRAM string (ntedit.cpp:1), kept for byte-compatible headers

	const ntb::ClinvarMap* cv = &clinvar;
	// Three stages run beside each other: a reader thread parses the draft into batches (its bytes are read / inflated ahead
	// of it by the ByteSource's own threads), this thread hands every batch to a device, and a writer thread writes the
	// finished batches in input order.  At most one batch per device is being polished, one more is being read.
	uint64_t tot_contigs = 0, tot_bases = 0, tot_edits = 0, n_read = 0;
	const auto t_begin = std::chrono::steady_clock::now();
	Channel<std::unique_ptr<Batch>> batches((size_t)1);
	Channel<std::future<BatchOut>> finished((size_t)opt.gpus + 1);
	DeviceSlots slots;
	std::thread reader_thread([&]() {
		bool more = true;
		std::string name, comment;
		while (more) {
			std::unique_ptr<Batch> b(new Batch());
			b->bases.reserve((size_t)std::min<uint64_t>(opt.batch_bases + (64u << 20), 1ull << 32));
			while (b->bases.size() < opt.batch_bases) {
				more = reader.next(name, comment, b->bases);
				if (!more) {
					break;
				}
				b->bases.push_back('\0');
				b->offsets.push_back(b->bases.size());
				b->names.push_back(comment.empty() ? name : name + " " + comment); // ntedit.cpp:2224-2229
				n_read++;
			}
			if (b->names.empty()) {
				break;
			}
			batches.push(std::move(b));
		}
		batches.close();
	});
	std::thread writer_thread([&]() {
		std::future<BatchOut> f;
		while (finished.pop(f)) {
			BatchOut o = f.get();
			dfout.write(o.fa.data(), (std::streamsize)o.fa.size());
			rfout.write(o.tsv.data(), (std::streamsize)o.tsv.size());
			vfout.write(o.vcf.data(), (std::streamsize)o.vcf.size());
			tot_contigs += o.contigs;
			tot_bases += o.bases;
			tot_edits += o.edits;
		}
	});
	{
		size_t next_dev = 0;
		std::unique_ptr<Batch> b;
		while (batches.pop(b)) {
			const size_t d = next_dev;
			next_dev = (next_dev + 1) % (size_t)opt.gpus;
			Batch* raw = b.release();
			slots.acquire((size_t)opt.gpus); // batches go round-robin and take about equally long: device d is the one that is free
			finished.push(std::async(std::launch::async, [raw, &bloom, &rep, &opt, cv, d, &slots]() {
				return process_batch(std::unique_ptr<Batch>(raw), bloom[d], rep[d], opt, cv, &slots);
			}));
		}
		finished.close();
	}
	reader_thread.join();
	writer_thread.join();
	dfout.close();
	rfout.close();
	vfout.close();
	if (!dfout || !rfout || !vfout) {
		std::cerr << PROGRAM ": error: writing the output files with prefix `" << opt.prefix << "' failed (disk full?)\n";
		return EXIT_FAILURE;
	}
	const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
	std::cout << "Processed " << n_read << " sequences: " << tot_contigs << " polished (" << tot_bases << " bases, " << tot_edits << " edits) in " << secs
	          << " s" << std::endl;
	std::cout << "---------- process complete                         : " << now_str();
	for (int d = 0; d < opt.gpus; d++) {
		ntb_filter_free(bloom[(size_t)d]);
		ntb_filter_free(rep[(size_t)d]);
	}
	return 0;
}
