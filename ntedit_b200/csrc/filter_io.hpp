// btllib Bloom-filter file format as ntEdit consumes it (BFWrapper, ntedit.cpp:355-364; written by
// src/ntedit_make_genome_bf.cpp:158-162 and by ntStat).  btllib is not vendored in the reference tree; layout per
// SURVEY.md Appendix B:  first line "[BTLKmerBloomFilter_v<N>]" or "[BTLKmerCountingBloomFilter_v<N>]", TOML-style
// "key = value" lines in ANY order (btllib writes them from an unordered map), a "[HeaderEnd]" line, then `bytes` raw
// bytes.  Bit n of a bit filter is byte n/8, mask 1<<(n%8).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>

namespace ntb {

struct FilterHeader
{
	uint64_t bytes = 0;
	uint32_t k = 0;
	uint32_t hash_num = 0;
	uint32_t counter_bits = 0;
	bool counting = false;
	std::string hash_fn;
	uint64_t data_offset = 0;
};

inline std::string
trim_ws(const std::string& s)
{
	const size_t b = s.find_first_not_of(" \t\r\n");
	if (b == std::string::npos) {
		return "";
	}
	const size_t e = s.find_last_not_of(" \t\r\n");
	return s.substr(b, e - b + 1);
}

inline bool
read_filter_header(const char* path, FilterHeader& h, std::string& err)
{
	FILE* fp = std::fopen(path, "rb");
	if (!fp) {
		err = std::string("cannot open Bloom filter file `") + path + "'";
		return false;
	}
	char line[1024];
	bool first = true, ended = false;
	while (std::fgets(line, sizeof line, fp)) {
		const std::string t = trim_ws(line);
		if (first) {
			first = false;
			// the same test BFWrapper makes through check_file_signature (ntedit.cpp:357-358), version-agnostic
			if (t.rfind("[BTLKmerCountingBloomFilter_v", 0) == 0) {
				h.counting = true;
			} else if (t.rfind("[BTLKmerBloomFilter_v", 0) == 0) {
				h.counting = false;
			} else {
				err = std::string("`") + path + "' is not a btllib k-mer Bloom filter (signature " + t + ")";
				std::fclose(fp);
				return false;
			}
			continue;
		}
		if (t == "[HeaderEnd]") {
			ended = true;
			h.data_offset = (uint64_t)ftello(fp);
			break;
		}
		const size_t eq = t.find('=');
		if (eq == std::string::npos) {
			continue;
		}
		const std::string key = trim_ws(t.substr(0, eq));
		std::string val = trim_ws(t.substr(eq + 1));
		if (val.size() >= 2 && val.front() == '"' && val.back() == '"') {
			val = val.substr(1, val.size() - 2);
		}
		if (key == "bytes") {
			h.bytes = std::strtoull(val.c_str(), nullptr, 10);
		} else if (key == "hash_num") {
			h.hash_num = (uint32_t)std::strtoul(val.c_str(), nullptr, 10);
		} else if (key == "k") {
			h.k = (uint32_t)std::strtoul(val.c_str(), nullptr, 10);
		} else if (key == "counter_bits") {
			h.counter_bits = (uint32_t)std::strtoul(val.c_str(), nullptr, 10);
		} else if (key == "hash_fn") {
			h.hash_fn = val;
		}
	}
	std::fclose(fp);
	if (!ended) {
		err = std::string("`") + path + "': no [HeaderEnd] line";
		return false;
	}
	if (h.bytes == 0 || h.k == 0 || h.hash_num == 0) {
		err = std::string("`") + path + "': header lacks bytes / k / hash_num";
		return false;
	}
	if (h.counting && h.counter_bits != 0 && h.counter_bits != 8) {
		err = std::string("`") + path + "': only 8-bit counting Bloom filters are supported (KmerCountingBloomFilter8)";
		return false;
	}
	if (!h.hash_fn.empty() && h.hash_fn.rfind("ntHash", 0) != 0) {
		err = std::string("`") + path + "': unsupported hash function " + h.hash_fn;
		return false;
	}
	return true;
}

inline std::string
format_filter_header(uint64_t bytes, uint32_t k, uint32_t hash_num, bool counting)
{
	std::string s;
	if (counting) {
		s = "[BTLKmerCountingBloomFilter_v5]\nbytes = " + std::to_string(bytes) + "\ncounter_bits = 8\nhash_fn = \"ntHash_v2\"\nhash_num = " +
		    std::to_string(hash_num) + "\nk = " + std::to_string(k) + "\n[HeaderEnd]\n";
	} else {
		s = "[BTLKmerBloomFilter_v7]\nbytes = " + std::to_string(bytes) + "\nhash_fn = \"ntHash_v2\"\nhash_num = " + std::to_string(hash_num) +
		    "\nk = " + std::to_string(k) + "\n[HeaderEnd]\n";
	}
	return s;
}

} // namespace ntb
