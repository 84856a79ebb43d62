// ORACLE / TEST INFRASTRUCTURE ONLY -- never included by the product path.
//
// Compatibility shim for the subset of <btllib/bloom_filter.hpp> used by
// /root/reference/ntedit.cpp:355-395 (BFWrapper) -- a k-mer Bloom filter file
// loader + membership test.  Restated from btllib's published behaviour
// (SURVEY.md Appendix B): TOML-ish header terminated by "[HeaderEnd]", then
// `bytes` raw bytes; bit n of the filter is byte n/8, mask 1<<(n%8); probe i is
// hashes[i] % (bytes*8); contains == all probes set.
// PARITY UNPINNED at this boundary: no .bf file and no btllib source exists in
// the reference tree, so the file layout/bit addressing is a recollection.
#ifndef ORACLE_SHIM_BTLLIB_BLOOM_FILTER_HPP
#define ORACLE_SHIM_BTLLIB_BLOOM_FILTER_HPP

#include "nthash.hpp"

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

namespace btllib {

static const char* const KMER_BLOOM_FILTER_SIGNATURE = "[BTLKmerBloomFilter_v7]";

namespace shim_detail {

struct Header
{
	std::string signature;
	uint64_t bytes = 0;
	unsigned hash_num = 0;
	unsigned k = 0;
	unsigned counter_bits = 0;
	std::string hash_fn;
	std::streamoff data_offset = 0;
};

inline std::string
trim(const std::string& s)
{
	size_t b = s.find_first_not_of(" \t\r");
	size_t e = s.find_last_not_of(" \t\r");
	if (b == std::string::npos) {
		return "";
	}
	return s.substr(b, e - b + 1);
}

inline bool
read_header(const std::string& path, Header& hdr)
{
	std::ifstream in(path, std::ios::binary);
	if (!in) {
		return false;
	}
	std::string line;
	if (!std::getline(in, line)) {
		return false;
	}
	hdr.signature = trim(line);
	while (std::getline(in, line)) {
		std::string t = trim(line);
		if (t == "[HeaderEnd]") {
			hdr.data_offset = in.tellg();
			return true;
		}
		size_t eq = t.find('=');
		if (eq == std::string::npos) {
			continue;
		}
		std::string key = trim(t.substr(0, eq));
		std::string val = trim(t.substr(eq + 1));
		if (!val.empty() && val[0] == '"') {
			val = val.substr(1, val.size() - 2);
		}
		if (key == "bytes") {
			hdr.bytes = std::strtoull(val.c_str(), nullptr, 10);
		} else if (key == "hash_num") {
			hdr.hash_num = (unsigned)std::strtoul(val.c_str(), nullptr, 10);
		} else if (key == "k") {
			hdr.k = (unsigned)std::strtoul(val.c_str(), nullptr, 10);
		} else if (key == "counter_bits") {
			hdr.counter_bits = (unsigned)std::strtoul(val.c_str(), nullptr, 10);
		} else if (key == "hash_fn") {
			hdr.hash_fn = val;
		}
	}
	return false;
}

inline void
load_data(const std::string& path, const Header& hdr, std::vector<uint8_t>& data)
{
	std::ifstream in(path, std::ios::binary);
	in.seekg(hdr.data_offset);
	data.resize(hdr.bytes);
	in.read(reinterpret_cast<char*>(data.data()), (std::streamsize)hdr.bytes);
	if ((uint64_t)in.gcount() != hdr.bytes) {
		std::cerr << "btllib-shim: truncated Bloom filter file " << path << std::endl;
		std::exit(EXIT_FAILURE);
	}
}

} // namespace shim_detail

class BloomFilter
{
  public:
	static bool check_file_signature(const std::string& path, const std::string& signature)
	{
		std::ifstream in(path, std::ios::binary);
		std::string line;
		if (!in || !std::getline(in, line)) {
			return false;
		}
		return shim_detail::trim(line) == signature;
	}
};

class KmerBloomFilter
{
  public:
	explicit KmerBloomFilter(const std::string& path)
	{
		shim_detail::Header hdr;
		if (!shim_detail::read_header(path, hdr)) {
			std::cerr << "btllib-shim: cannot parse Bloom filter header of " << path << std::endl;
			std::exit(EXIT_FAILURE);
		}
		k = hdr.k;
		hash_num = hdr.hash_num;
		shim_detail::load_data(path, hdr, data);
	}

	bool contains(const uint64_t* hashes) const
	{
		const uint64_t bits = (uint64_t)data.size() * 8;
		for (unsigned i = 0; i < hash_num; i++) {
			const uint64_t n = hashes[i] % bits;
			if (!(data[n / 8] & (uint8_t)(1u << (n % 8)))) {
				return false;
			}
		}
		return true;
	}

	unsigned get_k() const { return k; }
	unsigned get_hash_num() const { return hash_num; }
	size_t get_bytes() const { return data.size(); }
	double get_fpr() const
	{
		uint64_t pop = 0;
		for (uint8_t b : data) {
			pop += (uint64_t)__builtin_popcount(b);
		}
		return std::pow(double(pop) / double(data.size() * 8), double(hash_num));
	}

  private:
	unsigned k = 0;
	unsigned hash_num = 0;
	std::vector<uint8_t> data;
};

} // namespace btllib

#endif
