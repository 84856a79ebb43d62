// ORACLE / TEST INFRASTRUCTURE ONLY -- never included by the product path.
//
// Compatibility shim for the subset of <btllib/counting_bloom_filter.hpp> used
// by /root/reference/ntedit.cpp:357-391: 8-bit counting k-mer Bloom filter,
// counter i is data[hashes[i] % bytes], contains() returns the minimum counter.
// PARITY UNPINNED (see bloom_filter.hpp).
#ifndef ORACLE_SHIM_BTLLIB_COUNTING_BLOOM_FILTER_HPP
#define ORACLE_SHIM_BTLLIB_COUNTING_BLOOM_FILTER_HPP

#include "bloom_filter.hpp"

namespace btllib {

static const char* const KMER_COUNTING_BLOOM_FILTER_SIGNATURE = "[BTLKmerCountingBloomFilter_v5]";

class KmerCountingBloomFilter8
{
  public:
	explicit KmerCountingBloomFilter8(const std::string& path)
	{
		shim_detail::Header hdr;
		if (!shim_detail::read_header(path, hdr)) {
			std::cerr << "btllib-shim: cannot parse counting Bloom filter header of " << path
			          << std::endl;
			std::exit(EXIT_FAILURE);
		}
		k = hdr.k;
		hash_num = hdr.hash_num;
		shim_detail::load_data(path, hdr, data);
	}

	uint8_t contains(const uint64_t* hashes) const
	{
		uint8_t m = 255;
		for (unsigned i = 0; i < hash_num; i++) {
			const uint8_t c = data[hashes[i] % (uint64_t)data.size()];
			if (c < m) {
				m = c;
			}
		}
		return m;
	}

	unsigned get_k() const { return k; }
	unsigned get_hash_num() const { return hash_num; }
	size_t get_bytes() const { return data.size(); }
	double get_fpr() const
	{
		uint64_t nz = 0;
		for (uint8_t b : data) {
			nz += (b != 0);
		}
		return std::pow(double(nz) / double(data.size()), double(hash_num));
	}

  private:
	unsigned k = 0;
	unsigned hash_num = 0;
	std::vector<uint8_t> data;
};

} // namespace btllib

#endif
