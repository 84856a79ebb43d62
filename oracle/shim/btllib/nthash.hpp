// ORACLE / TEST INFRASTRUCTURE ONLY -- never included by the product path.
//
// From-scratch compatibility shim for the subset of btllib's <btllib/nthash.hpp>
// that /root/reference/ntedit.cpp uses (call sites ntedit.cpp:412-415, 428-431,
// 444-451).  btllib itself is NOT vendored in the reference tree and is not
// installed here (meson.build:20 finds it at build time, version unpinned), so
// the published ntHash2 algorithm is restated below.
// Pinned by btllib's own unit-test vector "ACATGCATGCA", k=5, h=3
// (tests/test_oracle_nthash.py).
#ifndef ORACLE_SHIM_BTLLIB_NTHASH_HPP
#define ORACLE_SHIM_BTLLIB_NTHASH_HPP

#include <cstdint>

namespace btllib {
namespace hashing_internals {

static const uint8_t CP_OFF = 0x07;
static const int MULTISHIFT = 27;
static const uint64_t MULTISEED = 0x90b45d39fb6da1faULL;

static const uint64_t SEED_A = 0x3c8bfbb395c60474ULL;
static const uint64_t SEED_C = 0x3193c18562a02b4cULL;
static const uint64_t SEED_G = 0x20323ed082572324ULL;
static const uint64_t SEED_T = 0x295549f54be24456ULL;
static const uint64_t SEED_N = 0;

// 256-entry seed table.  Slots 0..7 are the "complement by low three bits"
// entries used for the reverse strand: 'A'&7=1 -> T, 'C'&7=3 -> G, 'T'&7=4 -> A,
// 'U'&7=5 -> A, 'G'&7=7 -> C.
struct SeedTab
{
	uint64_t v[256];
	SeedTab()
	{
		for (auto& x : v) {
			x = SEED_N;
		}
		v[1] = SEED_T;
		v[3] = SEED_G;
		v[4] = SEED_A;
		v[5] = SEED_A;
		v[7] = SEED_C;
		v['A'] = v['a'] = SEED_A;
		v['C'] = v['c'] = SEED_C;
		v['G'] = v['g'] = SEED_G;
		v['T'] = v['t'] = SEED_T;
		v['U'] = v['u'] = SEED_T;
	}
	uint64_t operator[](unsigned char c) const { return v[c]; }
};
static const SeedTab SEED_TAB;

// split rotate: bits 0..32 (33 bits) and bits 33..63 (31 bits) rotate independently
inline uint64_t
srol(const uint64_t x)
{
	uint64_t m = ((x & 0x8000000000000000ULL) >> 30) | ((x & 0x100000000ULL) >> 32);
	return ((x << 1) & 0xFFFFFFFDFFFFFFFFULL) | m;
}

inline uint64_t
srol(const uint64_t x, const unsigned d)
{
	uint64_t v = x;
	for (unsigned i = 0; i < d; i++) {
		v = srol(v);
	}
	return v;
}

inline uint64_t
sror(const uint64_t x)
{
	uint64_t m = ((x & 0x200000000ULL) << 30) | ((x & 1ULL) << 32);
	return ((x >> 1) & 0xFFFFFFFEFFFFFFFFULL) | m;
}

inline uint64_t
srol_table(unsigned char c, unsigned d)
{
	return srol(SEED_TAB[c], d);
}

inline uint64_t
base_forward_hash(const char* seq, unsigned k)
{
	uint64_t h = 0;
	for (unsigned i = 0; i < k; i++) {
		h = srol(h);
		h ^= SEED_TAB[(unsigned char)seq[i]];
	}
	return h;
}

inline uint64_t
base_reverse_hash(const char* seq, unsigned k)
{
	uint64_t h = 0;
	for (unsigned i = 0; i < k; i++) {
		h = srol(h);
		h ^= SEED_TAB[(unsigned char)seq[k - 1 - i] & CP_OFF];
	}
	return h;
}

inline uint64_t
next_forward_hash(uint64_t fh, unsigned k, unsigned char char_out, unsigned char char_in)
{
	uint64_t h = srol(fh);
	h ^= SEED_TAB[char_in];
	h ^= srol_table(char_out, k);
	return h;
}

inline uint64_t
next_reverse_hash(uint64_t rh, unsigned k, unsigned char char_out, unsigned char char_in)
{
	uint64_t h = rh ^ srol_table(char_in & CP_OFF, k);
	h ^= SEED_TAB[char_out & CP_OFF];
	h = sror(h);
	return h;
}

inline uint64_t
canonical(uint64_t fwd, uint64_t rev)
{
	return fwd + rev;
}

inline void
extend_hashes(uint64_t base_hash, unsigned k, unsigned h, uint64_t* out)
{
	out[0] = base_hash;
	for (unsigned i = 1; i < h; i++) {
		uint64_t t = base_hash * (i ^ k * MULTISEED);
		t ^= t >> MULTISHIFT;
		out[i] = t;
	}
}

} // namespace hashing_internals
} // namespace btllib

#endif
