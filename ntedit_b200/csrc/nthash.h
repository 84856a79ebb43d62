// ntHash2 arithmetic as ntEdit uses it through btllib::hashing_internals (call sites ntedit.cpp:403-452;
// btllib is an un-vendored dependency of the reference, its published algorithm is summarised in
// SURVEY.md Appendix A).  Host+device inline functions; no tables in memory -- the four base seeds are
// selected arithmetically so the same code serves the scan kernel, the walker and the builder.
#pragma once
#include "ntb_common.h"

namespace ntb {

constexpr uint64_t SEED_A = 0x3c8bfbb395c60474ULL;
constexpr uint64_t SEED_C = 0x3193c18562a02b4cULL;
constexpr uint64_t SEED_G = 0x20323ed082572324ULL;
constexpr uint64_t SEED_T = 0x295549f54be24456ULL;
constexpr uint64_t MULTISEED = 0x90b45d39fb6da1faULL;
constexpr unsigned MULTISHIFT = 27;

// Base classes.  code 0..3 = A C G T (either case; U/u hash as T but are not "accepted"), 4 = other.
// The reverse strand of base code c uses the seed of 3-c.
// (branch-free: these run once per base in every kernel, and a switch costs an indirect branch -- an instruction-fetch stall)
NTB_HD unsigned
base_code(unsigned char ch)
{
	const uint32_t i = (uint32_t)(ch | 0x20) - (uint32_t)'a';
	const uint32_t sh = i < 26u ? i : 31u;
	const uint32_t is_base = (0x00180045u >> sh) & 1u;                  // a c g t u
	const uint32_t code = (uint32_t)(0x3C000002010ULL >> (2u * (sh & 31u))) & 3u; // a:0 c:1 g:2 t:3 u:3
	return is_base ? code : 4u;
}

NTB_HD uint64_t
seed_of_code(unsigned code)
{
	return code == 0 ? SEED_A : code == 1 ? SEED_C : code == 2 ? SEED_G : code == 3 ? SEED_T : 0ULL;
}

// SEED_TAB[ch] of btllib (forward strand)
NTB_HD uint64_t
seed_fwd(unsigned char ch)
{
	return seed_of_code(base_code(ch));
}

// SEED_TAB[ch & CP_OFF]: btllib's reverse-strand lookup goes through the low three bits of the character,
// slots {1:T, 3:G, 4:A, 5:A, 7:C}; every other slot is 0.  Note this is defined for ANY byte, e.g. 'Y'&7 == 1.
NTB_HD unsigned rev_code(unsigned char ch);

NTB_HD uint64_t
seed_rev(unsigned char ch)
{
	return seed_of_code(rev_code(ch));
}

// code (A C G T = 0..3, 4 = none) of the seed that SEED_TAB[ch & CP_OFF] selects: slots {1:T, 3:G, 4:A, 5:A, 7:C}
NTB_HD unsigned
rev_code(unsigned char ch)
{
	return (0x30051Cu >> (3u * ((uint32_t)ch & 7u))) & 7u;
}

// ntedit.cpp:486-499 (callers always pass toupper(c))
NTB_HD bool
is_atgc_upper(unsigned char c)
{
	return c == 'A' || c == 'T' || c == 'G' || c == 'C';
}

NTB_HD unsigned char
to_upper(unsigned char c)
{
	return (c >= 'a' && c <= 'z') ? (unsigned char)(c - 32) : c;
}

NTB_HD unsigned char
to_lower(unsigned char c)
{
	return (c >= 'A' && c <= 'Z') ? (unsigned char)(c + 32) : c;
}

NTB_HD bool
is_accepted_any_case(unsigned char c)
{
	// A T G C R Y S W K M B D H V, either case
	const uint32_t i = (uint32_t)(c | 0x20) - (uint32_t)'a';
	return i < 26u && ((0x016E14CFu >> i) & 1u);
}

// split rotate by one: bits 0..32 and bits 33..63 rotate independently
NTB_HD uint64_t
srol1(uint64_t x)
{
	const uint64_t carry = ((x & 0x8000000000000000ULL) >> 30) | ((x & 0x100000000ULL) >> 32);
	return ((x << 1) & 0xFFFFFFFDFFFFFFFFULL) | carry;
}

NTB_HD uint64_t
sror1(uint64_t x)
{
	const uint64_t carry = ((x & 0x200000000ULL) << 30) | ((x & 1ULL) << 32);
	return ((x >> 1) & 0xFFFFFFFEFFFFFFFFULL) | carry;
}

// split rotate by d: rotate the low 33 bits by d mod 33 and the high 31 bits by d mod 31
NTB_HD uint64_t
sroln(uint64_t x, unsigned d)
{
	const unsigned dl = d % 33, dh = d % 31;
	const uint64_t lo = x & 0x1FFFFFFFFULL;
	const uint64_t hi = x >> 33;
	const uint64_t lo_r = dl ? (((lo << dl) | (lo >> (33 - dl))) & 0x1FFFFFFFFULL) : lo;
	const uint64_t hi_r = dh ? (((hi << dh) | (hi >> (31 - dh))) & 0x7FFFFFFFULL) : hi;
	return (hi_r << 33) | lo_r;
}

struct HashState
{
	uint64_t fh, rh;
};

// NTMC64 rolling form, ntedit.cpp:418-432.  rot_k_out = srol^k(SEED_TAB[out]), rot_k_in_rev = srol^k(SEED_TAB[in & 7])
NTB_HD void
hash_roll(HashState& s, unsigned char out, unsigned char in, const KParams& p)
{
	const unsigned co = base_code(out);
	const unsigned ci = rev_code(in);
	uint64_t f = srol1(s.fh) ^ seed_fwd(in);
	if (co < 4) {
		f ^= p.seed_rot_k[co];
	}
	uint64_t r = s.rh ^ seed_rev(out);
	if (ci < 4) {
		r ^= p.seed_rot_k[ci];
	}
	s.fh = f;
	s.rh = sror1(r);
}

// NTMC64 seeding form, ntedit.cpp:403-416 (equals k rolls from an all-zero-seed window)
template<typename GetChar>
NTB_HD void
hash_seed(HashState& s, unsigned k, GetChar get)
{
	uint64_t f = 0, r = 0;
	for (unsigned i = 0; i < k; i++) {
		f = srol1(f) ^ seed_fwd(get(i));
		r = srol1(r) ^ seed_rev(get(k - 1 - i));
	}
	s.fh = f;
	s.rh = r;
}

// NTMC64_changelast, ntedit.cpp:434-452
NTB_HD void
hash_changelast(HashState& s, unsigned char out, unsigned char in, const KParams& p)
{
	s.fh ^= seed_fwd(out) ^ seed_fwd(in);
	const unsigned co = rev_code(out), ci = rev_code(in);
	if (co < 4) {
		s.rh ^= p.seed_rot_k1[co];
	}
	if (ci < 4) {
		s.rh ^= p.seed_rot_k1[ci];
	}
}

// btllib canonical(): fwd + rev (wrapping)
NTB_HD uint64_t
hash_canonical(const HashState& s)
{
	return s.fh + s.rh;
}

// i-th hash of btllib extend_hashes (i = 0 is the base hash itself)
NTB_HD uint64_t
hash_extend(uint64_t base, unsigned k, unsigned i)
{
	uint64_t t = base * ((uint64_t)i ^ ((uint64_t)k * MULTISEED));
	t ^= t >> MULTISHIFT;
	return i == 0 ? base : t;
}

NTB_HD uint64_t
mulhi64(uint64_t a, uint64_t b)
{
#if defined(__CUDA_ARCH__)
	return __umul64hi(a, b);
#else
	return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

// x % f.mod without a divide.  With c = floor((2^64-1)/mod) = (2^64-1-rho)/mod, 0 <= rho < mod:
//   x*c/2^64 = x/mod - x*(1+rho)/(mod*2^64), and the second term is < 1 (x < 2^64, 1+rho <= mod),
// so q = mulhi(x, c) is floor(x/mod) or one less: a single conditional subtraction finishes the job.
NTB_HD uint64_t
filter_slot(const FilterView& f, uint64_t x)
{
	if (f.mask) {
		return x & f.mask;
	}
	const uint64_t q = mulhi64(x, f.recip);
	uint64_t r = x - q * f.mod;
	r -= r >= f.mod ? f.mod : 0ULL;
	return r;
}

NTB_HD uint8_t
load_filter_byte(const uint8_t* p)
{
#if defined(__CUDA_ARCH__)
	return __ldg(p);
#else
	return *p;
#endif
}

// BFWrapper::get_count, ntedit.cpp:373-376: bit filter -> 1 ; counting filter -> min of the hash_num counters
NTB_HD unsigned
filter_count(const FilterView& f, uint64_t base, unsigned k)
{
	if (!f.counting) {
		return 1;
	}
	unsigned m = 255;
	for (unsigned i = 0; i < f.hash_num; i++) {
		const unsigned c = load_filter_byte(f.data + filter_slot(f, hash_extend(base, k, i)));
		m = c < m ? c : m;
	}
	return m;
}

// BFWrapper::contains, ntedit.cpp:368-371 (bit n lives in byte n/8 under mask 1<<(n%8))
NTB_HD bool
filter_contains(const FilterView& f, uint64_t base, unsigned k)
{
	if (f.counting) {
		return filter_count(f, base, k) > 0;
	}
	for (unsigned i = 0; i < f.hash_num; i++) {
		const uint64_t n = filter_slot(f, hash_extend(base, k, i));
		if (!((load_filter_byte(f.data + (n >> 3)) >> (n & 7)) & 1)) {
			return false;
		}
	}
	return true;
}

} // namespace ntb
