// Device engine ("walker"): one WARP replays ntEdit's per-contig state machine (kmerizeAndCorrect and callees,
// ntedit.cpp:1216-2151) over one segment of a contig, starting from a clean window, and reports the edit
// decisions as makeEdit-level events.  The authoritative rope is rebuilt on the host from those events
// (replay.hpp); the walker keeps only a bounded local copy of the rope tail -- enough to answer every
// getCharacter/increment/roll the reference would make while the window is "dirty" (overlaps an edit).
//
// Execution model.  All walker state lives in one WalkerState object per warp (shared memory on the device).  The
// control flow is warp-uniform: every lane takes the same branches, decided from that shared state.  State is only
// ever WRITTEN inside "leader sections" (lane 0, fenced by warp barriers); the expensive part of a site -- hashing
// and probing the filter for the check-missing subset, the substitution trials and the insertion / deletion
// candidates (ntedit.cpp:1826-1858, 1917-1981, 1451-1744) -- runs as one candidate per lane: the lane rolls the
// candidate's k-mers with the same hash_roll the scan uses, issues all Bloom probes of up to PROBE_G sampled k-mers
// before consuming any of them, and writes its support count to the shared state; the leader then resolves the
// candidates in the reference's order, so the decision is the one the sequential loop would have taken.
//
// While the window is clean (k consecutive, unedited draft bases) the walker does not roll base by base:
// it jumps to the next position flagged by the scan kernel's visit bitmap (K1) and re-seeds the hash there,
// which is exactly where the reference's main loop would next find `!bloom.contains(hVal)` (ntedit.cpp:1806).
//
// The file compiles for the device (product) and, for the test-only host simulator under tests/hostsim,
// for the CPU (one "lane"); the product library never instantiates the host version.
#pragma once
#include "nthash.h"

#include <cstring>

namespace ntb {

#if defined(__CUDACC__)
#define NTB_FN __host__ __device__
#define NTB_FN_NOINLINE __host__ __device__ __noinline__
#else
#define NTB_FN
#define NTB_FN_NOINLINE __attribute__((noinline))
#endif

// ------------------------------------------------------------------------------------------------------------------
// Execution model: a walker is run by a TEAM of NTB_TEAM consecutive lanes of a warp (NTB_TEAM = 32: the whole warp).
// Several teams share a warp: their leader sections then execute as one SIMT instruction stream wherever the teams
// take the same path, which divides the instruction-fetch and issue cost of the serial control code by the number of
// teams.  Host build: a team of one lane.
#if defined(__CUDACC__)
#ifndef NTB_TEAM
#define NTB_TEAM 32
#endif
#else
#undef NTB_TEAM
#define NTB_TEAM 1
#endif
static_assert(NTB_TEAM == 1 || NTB_TEAM == 2 || NTB_TEAM == 4 || NTB_TEAM == 8 || NTB_TEAM == 16 || NTB_TEAM == 32, "team size");

NTB_FN inline uint32_t
lane_id()
{
#if defined(__CUDA_ARCH__)
	return threadIdx.x & (NTB_TEAM - 1u);
#else
	return 0;
#endif
}

NTB_FN inline uint32_t
lane_count()
{
	return NTB_TEAM;
}

#if defined(__CUDACC__)
// first lane of the team inside its warp, and the team's lane mask
__device__ inline uint32_t
team_base()
{
	return threadIdx.x & 31u & ~(NTB_TEAM - 1u);
}
__device__ inline uint32_t
team_mask()
{
	return NTB_TEAM == 32 ? 0xFFFFFFFFu : (((1u << (NTB_TEAM & 31)) - 1u) << team_base());
}
#endif

NTB_FN inline void
warp_sync()
{
#if defined(__CUDA_ARCH__)
	__syncwarp(team_mask());
#endif
}

// smallest value of v over the lanes of the team
NTB_FN inline uint32_t
warp_min(uint32_t v)
{
#if defined(__CUDA_ARCH__)
	return __reduce_min_sync(team_mask(), v);
#else
	return v;
#endif
}

NTB_FN inline uint64_t
warp_xor64(uint64_t v)
{
#if defined(__CUDA_ARCH__)
	const uint32_t lo = __reduce_xor_sync(team_mask(), (uint32_t)v);
	const uint32_t hi = __reduce_xor_sync(team_mask(), (uint32_t)(v >> 32));
	return ((uint64_t)hi << 32) | lo;
#else
	return v;
#endif
}

NTB_FN inline uint32_t
warp_or(uint32_t v)
{
#if defined(__CUDA_ARCH__)
	return __reduce_or_sync(team_mask(), v);
#else
	return v;
#endif
}

// optional per-phase cycle accounting (tuning aid, -DNTB_PHASE_PROF): S.prof[p] accumulates the leader's clock
#if defined(NTB_PHASE_PROF) && defined(__CUDA_ARCH__)
#define NTB_PROF_T0 long long prof_t0_ = clock64();
#define NTB_PROF(p)                                   \
	do {                                              \
		const long long prof_t1_ = clock64();         \
		if (lane_id() == 0) {                         \
			S.prof[p] += (uint32_t)(prof_t1_ - prof_t0_); \
		}                                             \
		prof_t0_ = prof_t1_;                          \
	} while (0)
#else
#define NTB_PROF_T0
#define NTB_PROF(p)
#endif

#define NTB_LEADER_BEGIN \
	warp_sync();         \
	if (lane_id() == 0) {
#define NTB_LEADER_END \
	}                  \
	warp_sync();

// filter probes are uniformly random over a multi-GB array: read-only path, do not allocate in L1
NTB_FN inline uint32_t
probe_byte(const uint8_t* p)
{
#if defined(__CUDA_ARCH__)
	uint32_t v;
	asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
#else
	return *p;
#endif
}

struct Cursor
{
	uint32_t pos; // h_seq_i / t_seq_i
	uint32_t ni;  // h_node_index / t_node_index
};

// result of one site evaluation (the locals of ntedit.cpp:1876-1888)
struct Site
{
	uint32_t best_type;
	uint32_t best_support;
	uint32_t altsupp1, altsupp2, altsupp3;
	uint8_t best_sub, altbase1, altbase2, altbase3;
	uint8_t indel_len;
	char indel[11];
};

struct WalkerIO
{
	const unsigned char* text; // contig bases
	uint32_t len;
	const uint32_t* visit;     // K1 bitmap, bit (goff + pos)
	uint64_t goff;
	FilterView bloom, rep;
	Event* events;
	uint32_t ev_cap;
	Counters* ctr;
	const uint64_t* rot;       // ROT_WORDS entries, see rot_entry()
	SiteRec* table;            // pre-evaluated sites (ntb_common.h: SiteRec), table_mask + 1 slots; nullptr = none
	uint32_t table_mask;
	PendingSite* pending;      // pre-evaluation, first pass: sites left for the second pass
	uint32_t pending_cap;
};

// slot a key starts probing at
NTB_FN inline uint32_t
site_hash(uint64_t key)
{
	key *= 0x9E3779B97F4A7C15ULL;
	return (uint32_t)(key >> 32);
}

// claims a slot for `key` (no key is inserted twice: every position is pre-evaluated by one item only); NONE32 = no room
NTB_FN inline uint32_t
site_table_insert(SiteRec* table, uint32_t mask, uint64_t key)
{
	uint32_t slot = site_hash(key) & mask;
	for (uint32_t i = 0; i < SITE_TABLE_PROBES; i++, slot = (slot + 1) & mask) {
#if defined(__CUDA_ARCH__)
		const unsigned long long seen = atomicCAS(reinterpret_cast<unsigned long long*>(&table[slot].key), 0ULL, (unsigned long long)key);
		if (seen == 0ULL || seen == (unsigned long long)key) {
			return slot;
		}
#else
		if (table[slot].key == 0 || table[slot].key == key) {
			table[slot].key = key;
			return slot;
		}
#endif
	}
	return NONE32;
}

constexpr uint32_t MAX_INS_TRIES = 341;   // num_tries[5], ntedit.cpp:172
constexpr uint32_t MAX_DELETIONS = 10;    // ntedit.cpp:2489-2493
#ifndef NTB_PROBE_G
#define NTB_PROBE_G 3
#endif
constexpr int PROBE_G = NTB_PROBE_G;              // sampled k-mers whose probes are in flight together, per lane
constexpr int PROBE_HU = 4;               // hash functions probed per pass (hash_num <= HMAX takes ceil(h/4) passes)
constexpr uint32_t TEXT_CACHE = 256;      // bytes of contig text kept in the shared state around the window
constexpr uint32_t LOOKAHEAD = 32;        // dirty-window positions whose site test is evaluated in one pass
constexpr uint32_t ROT_STRIDE = KMAX + 8; // rotation amounts tabulated per seed: srol^d(seed), d < ROT_STRIDE
constexpr uint32_t ROT_ROWS = 5;          // A C G T none
constexpr uint32_t ROT_WORDS = ROT_ROWS * ROT_STRIDE;

// rot[code * ROT_STRIDE + d] = srol^d(seed of code); one copy per CTA (per process in the host build)
NTB_FN inline uint64_t
rot_entry(uint32_t idx)
{
	const uint32_t code = idx / ROT_STRIDE, d = idx % ROT_STRIDE;
	return code < 4 ? sroln(seed_of_code(code), d) : 0ULL;
}

// One candidate k-mer series: "change the window's last base to X, then roll Q times; the first `c` incoming bases
// are synthetic (an insertion string), the rest come from the linearised rope starting at offset `d`; sample the
// filter after roll `first` and every `period` rolls after it (and, with `pre`, before the first roll)".
struct Cand
{
	uint64_t syn_cls;  // class bytes (cls_of) of the synthetic incoming chars, byte j = j-th
	uint32_t Q;        // rolls to perform
	uint32_t c;        // leading synthetic rolls
	uint32_t d;        // offset into lin_in[] of the first real incoming char
	uint32_t first;    // index of the first roll after which a sample is taken
	uint32_t period;   // sampling period (>= 1)
	uint8_t X;         // new tail char (when change)
	uint8_t change;    // apply NTMC64_changelast(draft -> X) first
	uint8_t pre;       // sample the changed k-mer itself
	uint8_t patch;     // the outgoing char read at the tail's own slot is X (in-place substitution trial)
	uint8_t kind;      // CK_*
};
constexpr uint8_t CK_CHECK = 0; // record the raw filter value of every sample (check-missing loop)
constexpr uint8_t CK_SOLID = 1; // count samples that are present && solid
constexpr uint8_t CK_SITE = 2;  // one sample: would the main loop enter its edit block here?

// Everything a walker mutates.  One per warp; written by the leader lane only.
template<int NCAP>
struct WalkerState
{
	static constexpr int OVCAP = KMAX + 8;

	WalkerIO io;

	// bounded copy of the rope tail (seqNode vector, ntedit.cpp:613-620)
	int8_t ty[NCAP];
	uint8_t ch[NCAP];
	uint32_t sp[NCAP], ep[NCAP];
	uint32_t nn;

	// substitutions applied in place to position nodes that the head may still read (contigSeq mutation, ntedit.cpp:1283)
	uint32_t ov_pos[OVCAP];
	uint8_t ov_ch[OVCAP];
	uint32_t ov_n;

	Cursor h, t;
	HashState hs;
	unsigned char char_in;

	// the reference declares these without initialisers inside the loop body (ntedit.cpp:1881-1885); the compiled
	// reference keeps them in fixed slots, so stale values leak from site to site (observable in mode 2)
	uint8_t stale_best_sub, stale_alt1, stale_alt2, stale_alt3;

	uint32_t adv;      // tail increments since the last event
	bool anchored;
	uint32_t last_event, n_events, n_sites, first_touch, status;

	// main-loop control (decided by the leader, read by every lane)
	uint32_t t_start, t_end; // the task's borders after safe_boundary()
	uint32_t end_pos;
	uint32_t act;
	uint32_t visit_hit;
	bool need_seed, do_seed, site_now;

	// contig text around the window
	alignas(4) uint8_t tc[TEXT_CACHE];
	uint32_t tc_base, tc_n;

	// ---- scratch of the site being evaluated
	// linearised rope: chars the next rolls would push out of / pull into the window (roll(), ntedit.cpp:1216-1247)
	uint8_t lin_out[KMAX + LOOKAHEAD + 2];
	uint8_t lin_in[KMAX + LOOKAHEAD + 2];
	uint8_t lin_out_c[KMAX + LOOKAHEAD + 2]; // class bytes of the same
	uint8_t lin_in_c[KMAX + LOOKAHEAD + 2];
	uint64_t seed_tab[8];                    // forward seeds by code (A C G T, none)
	uint64_t rotk_tab[8];                    // srol^k of the same
	uint64_t hb[PROBE_G][NTB_TEAM];         // sampled k-mer hashes of every lane, waiting to be probed
	uint32_t pv[PROBE_G][PROBE_HU][NTB_TEAM]; // probed filter words (cp.async destinations)
	uint8_t psh[PROBE_G][PROBE_HU][NTB_TEAM]; // bit offset of the probed bit / counter inside the word
	uint8_t pval[PROBE_G][NTB_TEAM];        // filter value of each sampled k-mer
	uint32_t n_rolls;     // successful simulated rolls
	uint32_t n_check;     // completed iterations of the check-missing loop
	uint32_t patch_idx;   // index into lin_out[] that reads the tail's own slot
	bool dnf;             // do_not_fix
	bool lin_simple;      // the window sits on one position node that runs to the end of the contig
	bool jumped;          // the dirty run after an edit was skipped in one step
	unsigned char raw, draft;
	uint32_t cands;       // substitution candidates, packed
	bool tail_is_pos, tail_is_chr, touched;
	uint32_t next;        // what the candidate loop does next
	uint32_t ci;          // substitution candidate being processed
	unsigned char index_char;
	Site s;
	uint32_t num_deletions;
	bool site_ok;
	// phase results
	uint8_t chk[KMAX + 1];
	uint32_t chk_n;
	uint8_t gate[4];
	uint32_t sup[4];
	uint8_t ins_sup[MAX_INS_TRIES];
	uint8_t del_sup[MAX_DELETIONS + 2];
	// hash state after R plain rolls from the current window (R = 0 is the window itself), R <= n_plain: every k-mer of
	// the check-missing subset, of the substitution trials and of the dirty-run look-ahead is one of these, or one of
	// these plus the rotation-table terms of the changed last base
	static constexpr int PLAIN_CAP = (NCAP <= 160 ? 48 : (int)KMAX) + (int)LOOKAHEAD + 2;
	uint64_t plain_f[PLAIN_CAP], plain_r[PLAIN_CAP];
	uint32_t n_plain;
	uint8_t tsolid[4][KMAX + 1]; // substitution trial, k-mer after R rolls (index R-1): present && solid
	uint8_t tsite[4][KMAX + 1];  // ... the main loop would enter its edit block at that k-mer
	bool p1_fast;                // phase 1 took the table path (tsolid / tsite are filled for every roll of the tried candidates)
	bool quiet;                  // accepted substitution, none of the k-1 windows behind it is a site (see evaluate_site_core)
	bool skip_advance;           // site_commit already moved the window to the next clean position
	uint32_t use_rec;            // this site's decision comes from a pre-evaluated record (its SITE_* state; 0 = evaluate here)
	bool rec_hash;               // ... and committing it needs the window's hash and the text cache
	uint8_t rec_fl;              // ... its EV_TOUCHED flag
	bool rec_second;             // ... it was completed by the second pre-evaluation pass (diagnostics)
	uint32_t rec_slot;           // pre-evaluation: slot of the record being written
	uint32_t rec_skip_T;         // ... the chain's next rec_skip_n sites make no edit and emit nothing (SITE_FL_SKIP): what they
	uint16_t rec_skip_dist;      //     leave in the stale slots, and the distance to the last of them
	uint8_t rec_skip_n;
	bool pre_more;               // pre-evaluation: the chain goes on with the next position
	// insertion candidates without rolling: hash state after q+1 rolls of an insertion of length L whose inserted chars
	// contribute nothing (ins_base_*[L-1][sample]); the candidates add their chars' terms from the rotation table
	static constexpr int NSMAX = NCAP <= 160 ? 16 : 32;
	uint64_t ins_base_f[5][NSMAX], ins_base_r[5][NSMAX];
	uint8_t ins_nvalid[5];
	bool bases_ready, ins_fast;
	// tryIndels bookkeeping across chunks
	uint32_t ti_i0, ti_i1, ti_nd0, ti_ndel;
	uint32_t tb_support, ta_support, tb_type, tb_len;
	char tb_indel[8];
	bool ti_done, ti_ret;
	// dirty-window lookahead: bit j = the main loop would enter its edit block after j more plain rolls
	uint32_t la_bits, la_n, la_used, la_J;
	uint32_t prof[16];
};

constexpr uint32_t ACT_STOP = 0, ACT_CLEAN = 1, ACT_DIRTY = 2;
constexpr uint32_t NEXT_CAND = 0, NEXT_INDELS = 1, NEXT_STOP = 2;

constexpr uint32_t WALK_ROT_BYTES = (ROT_WORDS * 8u + 127u) & ~127u;
constexpr uint32_t WALK_CLS_BYTES = 256u; // class byte of every text byte: seed codes of both strands + accepted flag
constexpr uint32_t WALK_KP_BYTES = (sizeof(KParams) + 127u) & ~127u; // device: the CTA's copy of the parameters sits in front of the team states

// COMMON = the configuration nearly every polishing run uses -- bit filter, no secondary filter (-e), not -s 1, not -a 1 --
// known at compile time: the counting-filter, secondary-filter, SNV and masking branches fold away, which takes ~7 % off
// the kernel's code and ~10 % off its run time (the walker is bound by instruction fetch).
// POW2 = every filter the walker probes has a power-of-two number of slots: `% size` is a mask and the multiply-high
// remainder (20 instructions per probe, if-converted into every probe site otherwise) is not compiled in.
template<int NCAP, bool COMMON = false, bool POW2 = false>
struct Walker
{
	static constexpr int OVCAP = WalkerState<NCAP>::OVCAP;
	static constexpr int PREVCAP = 2 * KMAX + 16;

#if defined(__CUDA_ARCH__)
	// Device: the state of this team lives in the CTA's dynamic shared memory (walk_kernel lays the teams' states out from
	// offset 0).  Every member function re-derives its address from the __shared__ symbol instead of keeping a reference in
	// the object: a reference would be a generic pointer once `this` escapes into a non-inlined call, and every state access
	// a generic LD/ST; this way they are LDS/STS.
	// Layout of the dynamic shared memory:
	//   [KParams copy, WALK_KP_BYTES][rotation table, WALK_ROT_BYTES][class table, WALK_CLS_BYTES][team states].
	__device__ __forceinline__ WalkerState<NCAP>& state_() const
	{
		extern __shared__ __align__(16) uint8_t ntb_walk_smem[];
		return reinterpret_cast<WalkerState<NCAP>*>(ntb_walk_smem + WALK_KP_BYTES + WALK_ROT_BYTES + WALK_CLS_BYTES)[threadIdx.x / NTB_TEAM];
	}
	__device__ __forceinline__ const uint8_t* cls_tab_() const
	{
		extern __shared__ __align__(16) uint8_t ntb_walk_smem[];
		return ntb_walk_smem + WALK_KP_BYTES + WALK_ROT_BYTES;
	}
	__device__ __forceinline__ const uint64_t* rot_() const
	{
		extern __shared__ __align__(16) uint8_t ntb_walk_smem[];
		return reinterpret_cast<const uint64_t*>(ntb_walk_smem + WALK_KP_BYTES);
	}
	__device__ __forceinline__ const KParams& params_() const
	{
		extern __shared__ __align__(16) uint8_t ntb_walk_smem[];
		return *reinterpret_cast<const KParams*>(ntb_walk_smem);
	}
	__device__ Walker(WalkerState<NCAP>&, const KParams&) {}
#else
	WalkerState<NCAP>& S_;
	const KParams& P_;
	WalkerState<NCAP>& state_() const { return S_; }
	const KParams& params_() const { return P_; }
	const uint64_t* rot_() const { return S_.io.rot; }
	Walker(WalkerState<NCAP>& s_, const KParams& p_) : S_(s_), P_(p_) {}
#endif
#define S (state_())
#define P (params_())
	NTB_FN uint32_t counting_() const { return COMMON ? 0u : P.counting; }
	NTB_FN uint32_t h_rep_() const { return COMMON ? 0u : P.h_rep; }
	NTB_FN int snv_() const { return COMMON ? 0 : P.snv; }
	NTB_FN int mask_() const { return COMMON ? 0 : P.mask; }
	NTB_FN uint32_t fcounting_(const FilterView& f) const { return COMMON ? 0u : f.counting; }
	NTB_FN uint64_t slot_(const FilterView& f, uint64_t x) const { return POW2 ? (x & f.mask) : filter_slot(f, x); }

	// Class byte of a text byte: bits 0-2 code of the forward seed, bits 3-5 code of the reverse-strand seed (btllib's
	// SEED_TAB[c & 7] path; code 4 = no seed), bit 6 accepted base (ntedit.cpp:493-499).  On the device one load from a
	// 256-byte table in shared memory (the arithmetic forms of nthash.h cost 15-20 instructions at every one of ~40
	// inlined sites, and the walker is bound by instruction fetch); on the host computed.
	NTB_FN uint32_t cls8(unsigned char c) const
	{
#if defined(__CUDA_ARCH__)
		return cls_tab_()[c];
#else
		return base_code(c) | (rev_code(c) << 3) | (is_accepted_any_case(c) ? 0x40u : 0u);
#endif
	}
	NTB_FN uint32_t code_f(unsigned char c) const { return cls8(c) & 7u; }
	NTB_FN uint32_t code_r(unsigned char c) const { return (cls8(c) >> 3) & 7u; }
	NTB_FN bool is_acc(unsigned char c) const { return ((cls8(c) >> 6) & 1u) != 0; }

	// ---------------------------------------------------------------- text / rope access
	NTB_FN unsigned char rd(uint32_t pos) const
	{
		for (uint32_t i = 0; i < S.ov_n; i++) {
			if (S.ov_pos[i] == pos) {
				return S.ov_ch[i];
			}
		}
		if (pos >= S.io.len) {
			return 0;
		}
		const uint32_t o = pos - S.tc_base;
		if (o < S.tc_n) {
			return S.tc[o];
		}
		return S.io.text[pos];
	}

	// getCharacter, ntedit.cpp:812-823
	NTB_FN unsigned char cchar(const Cursor& c) const
	{
		if (c.ni >= S.nn) {
			return 0;
		}
		if (S.ty[c.ni] == 0) {
			return rd(c.pos);
		}
		if (S.ty[c.ni] == 1) {
			return S.ch[c.ni];
		}
		return 0;
	}

	// increment, ntedit.cpp:826-844
	NTB_FN void step(Cursor& c) const
	{
		if (c.ni >= S.nn) {
			return;
		}
		const int8_t tp = S.ty[c.ni];
		if (tp == 0) {
			c.pos++;
			if (c.pos > S.ep[c.ni]) {
				c.ni++;
				if (c.ni < S.nn && S.ty[c.ni] == 0) {
					c.pos = S.sp[c.ni];
				}
			}
		} else if (tp == 1) {
			c.ni++;
			if (c.ni < S.nn && S.ty[c.ni] == 0) {
				c.pos = S.sp[c.ni];
			}
		}
	}

	// roll, ntedit.cpp:1216-1247
	NTB_FN bool roll(Cursor& hh, Cursor& tt, unsigned char& out, unsigned char& in) const
	{
		if (hh.pos >= S.io.len || hh.ni >= S.nn) {
			return false;
		}
		out = cchar(hh);
		step(hh);
		if (tt.pos >= S.io.len || tt.ni >= S.nn) {
			return false;
		}
		step(tt);
		if (tt.pos >= S.io.len || tt.ni >= S.nn) {
			return false;
		}
		in = cchar(tt);
		return true;
	}

	// ---- rope surgery: leader only
	NTB_FN void put(uint32_t i, int8_t type, uint8_t c, uint32_t s, uint32_t e)
	{
		if (i >= (uint32_t)NCAP) {
			S.status |= ST_ROPE_OVERFLOW;
			return;
		}
		S.ty[i] = type;
		S.ch[i] = c;
		S.sp[i] = s;
		S.ep[i] = e;
		if (i >= S.nn) {
			S.nn = i + 1;
		}
	}

	NTB_FN void move_node(uint32_t dst, uint32_t src)
	{
		S.ty[dst] = S.ty[src];
		S.ch[dst] = S.ch[src];
		S.sp[dst] = S.sp[src];
		S.ep[dst] = S.ep[src];
	}

	// makeInsertion, ntedit.cpp:625-714
	NTB_FN_NOINLINE void rope_insert(uint32_t& t_ni, uint32_t insert_pos, const char* bases, uint32_t nb)
	{
		const int8_t otype = S.ty[t_ni];
		const uint32_t os = S.sp[t_ni], oe = S.ep[t_ni];
		if (otype == 0 && insert_pos > os) {
			S.ep[t_ni] = insert_pos - 1;
			for (uint32_t i = 0; i < nb; i++) {
				put(t_ni + i + 1, 1, (uint8_t)bases[i], 0, 0);
			}
			put(t_ni + nb + 1, 0, 0, insert_pos, oe);
			t_ni++;
			return;
		}
		if (otype == 0 || otype == 1) {
			// lift the live run starting at the tail node and put it back behind the inserted characters
			uint32_t nlift = 0;
			while (t_ni + nlift < S.nn && S.ty[t_ni + nlift] != -1) {
				nlift++;
			}
			if (t_ni + nb + nlift > (uint32_t)NCAP) {
				S.status |= ST_ROPE_OVERFLOW;
				return;
			}
			if (t_ni + nb + nlift > S.nn) {
				S.nn = t_ni + nb + nlift;
			}
			for (uint32_t q = nlift; q > 0; q--) {
				move_node(t_ni + nb + q - 1, t_ni + q - 1);
			}
			// slots between the lifted run's old end and its new start that were not overwritten stay as the
			// reference leaves them: the old entries were marked dead before being re-appended
			for (uint32_t q = 0; q < nb; q++) {
				S.ty[t_ni + q] = 1;
				S.ch[t_ni + q] = (uint8_t)bases[q];
				S.sp[t_ni + q] = 0;
				S.ep[t_ni + q] = 0;
			}
		}
	}

	// makeDeletion, ntedit.cpp:719-809 (the recursion of the reference is a loop here)
	NTB_FN_NOINLINE void rope_delete(uint32_t& t_ni, uint32_t& pos, uint32_t num_del)
	{
		for (;;) {
			const int8_t otype = S.ty[t_ni];
			const uint32_t os = S.sp[t_ni], oe = S.ep[t_ni];
			uint32_t leftover = 0;
			if (otype == 0) {
				if (pos <= os) {
					if (pos + num_del <= oe) {
						S.sp[t_ni] = pos + num_del;
						pos = S.sp[t_ni];
						return;
					}
					leftover = pos + num_del - oe;
					pos = oe + 1;
					uint32_t i = t_ni + 1;
					while (i < S.nn && S.ty[i] != -1) {
						move_node(i - 1, i);
						S.ty[i] = -1;
						i++;
					}
				} else {
					if (pos + num_del <= oe) {
						S.ep[t_ni] = pos - 1;
						const uint32_t ns = pos + num_del;
						pos = ns;
						t_ni++;
						put(t_ni, 0, 0, ns, oe);
						return;
					}
					leftover = pos + num_del - oe;
					S.ep[t_ni] = pos - 1;
					pos = oe + 1;
					t_ni++;
				}
			} else if (otype == 1) {
				uint32_t i = t_ni;
				leftover = num_del;
				while (i < S.nn && S.ty[i] == 1 && leftover > 0) {
					S.ty[i] = -1;
					leftover--;
					i++;
				}
				uint32_t j = t_ni;
				while (i < S.nn && S.ty[i] != -1) {
					move_node(j, i);
					S.ty[i] = -1;
					i++;
					j++;
				}
			} else {
				return;
			}
			if (leftover > 0 && t_ni < S.nn && S.ty[t_ni] != -1) {
				if (S.ty[t_ni] == 0) {
					pos = S.sp[t_ni];
				}
				num_del = leftover;
				continue;
			}
			return;
		}
	}

	// ---------------------------------------------------------------- filter queries (single k-mer, leader lane)
	// value of a k-mer in a filter: bit filter -> 1 when all hash_num bits are set else 0; counting filter -> min counter
	// (BFWrapper::contains / get_count, ntedit.cpp:368-376)
	NTB_FN unsigned q_count(const HashState& s) const { return filter_count(S.io.bloom, hash_canonical(s), P.k); }

	NTB_FN bool q_contains(const HashState& s) const { return filter_contains(S.io.bloom, hash_canonical(s), P.k); }

	NTB_FN bool meets_edit(uint32_t c) const { return c >= P.thr_edit; }

	// the main loop's test, ntedit.cpp:1806-1807
	NTB_FN bool is_site_value(uint32_t c) const { return snv_() || c == 0 || (counting_() && c < P.min_threshold); }

	// bloom.contains(hVal) && is_kmer_solid(hVal, bloom, bloomrep) without the secondary filter, ntedit.cpp:465-473
	NTB_FN bool solid_value(uint32_t c) const
	{
		if (counting_()) {
			return !(c == 0 || c < P.min_threshold || c > P.max_threshold);
		}
		return c != 0;
	}

	// ---------------------------------------------------------------- probing a group of sampled k-mers
	// class byte of a base: bits 0-2 code of the forward seed, bits 3-5 code of the reverse-strand seed (btllib's
	// SEED_TAB[c & 7] path); code 4 = no seed
	NTB_FN uint8_t cls_of(unsigned char c) const { return (uint8_t)(cls8(c) & 0x3Fu); }

	// NTMC64_changelast (ntedit.cpp:434-452) through the class table
	NTB_FN void w_changelast(HashState& s, unsigned char out, unsigned char in) const
	{
		const uint32_t co = cls8(out), ci = cls8(in);
		s.fh ^= S.seed_tab[co & 7u] ^ S.seed_tab[ci & 7u];
		const uint32_t ro = (co >> 3) & 7u, ri = (ci >> 3) & 7u;
		s.rh ^= (ro < 4u ? P.seed_rot_k1[ro & 3u] : 0ULL) ^ (ri < 4u ? P.seed_rot_k1[ri & 3u] : 0ULL);
	}

	// NTMC64 rolling form (ntedit.cpp:418-432) on class bytes, seeds from the per-warp tables
	NTB_FN void roll_cls(HashState& s, uint32_t co, uint32_t ci) const
	{
		s.fh = srol1(s.fh) ^ S.seed_tab[ci & 7u] ^ S.rotk_tab[co & 7u];
		s.rh = sror1(s.rh ^ S.rotk_tab[(ci >> 3) & 7u] ^ S.seed_tab[(co >> 3) & 7u]);
	}

	// Issues the probes of the first `ng` sampled k-mers of this lane (S.hb[g][lane], only those whose bit is set in
	// `want`) for hash functions [i0, i0 + PROBE_HU) of filter F.  On the device every probe is an asynchronous 4-byte copy
	// of the aligned filter word into the lane's slot of S.pv (cp.async: no register is tied up while the load is in
	// flight, so all ng x PROBE_HU loads of a lane overlap without unrolling anything); S.psh keeps the bit offset of the
	// probed byte / bit inside that word.
	NTB_FN void probe_issue(const FilterView& F, uint32_t ng, uint32_t want, uint32_t i0, uint32_t hu)
	{
		const uint32_t ln = lane_id();
		const uint32_t hn = F.hash_num - i0 < hu ? F.hash_num - i0 : hu;
		for (uint32_t g = 0; g < ng; g++) {
			if (!((want >> g) & 1u)) {
				continue;
			}
			const uint64_t hb = S.hb[g][ln];
			for (uint32_t u = 0; u < hn; u++) {
				const uint64_t slot = slot_(F, hash_extend(hb, P.k, i0 + u));
				const uint64_t byte = fcounting_(F) ? slot : slot >> 3;
				S.psh[g][u][ln] = (uint8_t)(((uint32_t)byte & 3u) * 8u + (fcounting_(F) ? 0u : ((uint32_t)slot & 7u)));
				const uint8_t* src = F.data + (byte & ~3ULL);
#if defined(__CUDA_ARCH__)
				const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&S.pv[g][u][ln]);
				asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
#if defined(NTB_PHASE_PROF)
				atomicAdd(&S.prof[14], 1u); // filter probes issued by this walker
#endif
#else
				uint32_t w;
				std::memcpy(&w, src, 4);
				S.pv[g][u][ln] = w;
#endif
			}
		}
#if defined(__CUDA_ARCH__)
		asm volatile("cp.async.commit_group;" ::: "memory");
		asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
	}

	// value of sampled k-mer g after probe_issue(F, .., i0): folds hash functions [i0, i0 + PROBE_HU) into `val`
	// (bit filter: AND of the probed bits; counting filter: min of the probed counters)
	NTB_FN uint32_t probe_fold(const FilterView& F, uint32_t g, uint32_t i0, uint32_t hu, uint32_t val) const
	{
		const uint32_t ln = lane_id();
		const uint32_t hn = F.hash_num - i0 < hu ? F.hash_num - i0 : hu;
		for (uint32_t u = 0; u < hn; u++) {
			const uint32_t w = S.pv[g][u][ln] >> S.psh[g][u][ln];
			if (fcounting_(F)) {
				const uint32_t c = w & 0xFFu;
				val = c < val ? c : val;
			} else {
				val &= w;
			}
		}
		return val;
	}

	// values of the first ng sampled k-mers of this lane in filter F -> S.pval[g][lane] (only those in `want`), `hu` hash
	// functions per pass.  A k-mer whose value reached 0 (a probed bit is clear / a counter is 0) is dropped from the later
	// passes -- the value cannot change any more -- which is btllib's early exit (ntedit.cpp:368-371) done per pass.
	NTB_FN_NOINLINE void probe_values(const FilterView& F, uint32_t ng, uint32_t want, uint32_t hu)
	{
		const uint32_t ln = lane_id();
		for (uint32_t i0 = 0; i0 < F.hash_num && want; i0 += hu) {
			probe_issue(F, ng, want, i0, hu);
			for (uint32_t g = 0; g < ng; g++) {
				if ((want >> g) & 1u) {
					const uint32_t start = i0 == 0 ? (fcounting_(F) ? 255u : 1u) : (uint32_t)S.pval[g][ln];
					const uint32_t v = probe_fold(F, g, i0, hu, start);
					S.pval[g][ln] = (uint8_t)v;
					if (v == 0) {
						want &= ~(1u << g);
					}
				}
			}
		}
	}

	// probes the ng sampled k-mers of this lane (S.hb[0..ng)[lane]) and classifies them
	NTB_FN void finish_group(uint32_t kind, uint32_t ng, uint32_t premask, uint32_t hu, uint32_t& pre_ok, uint32_t& count, uint32_t& nchk)
	{
		const uint32_t ln = lane_id();
		const uint32_t valid = (1u << ng) - 1u;
		probe_values(S.io.bloom, ng, valid, hu);
		if (kind == CK_CHECK) {
			for (uint32_t g = 0; g < ng; g++) {
				S.chk[nchk++] = S.pval[g][ln];
			}
		} else if (kind == CK_SITE) {
			for (uint32_t g = 0; g < ng; g++) {
				if (is_site_value(S.pval[g][ln])) {
					count = 1;
				}
			}
		} else {
			uint32_t solid = 0;
			for (uint32_t g = 0; g < ng; g++) {
				if (solid_value(S.pval[g][ln])) {
					solid |= 1u << g;
				}
			}
			if (h_rep_() && solid) {
				// secondary filter (-e): a k-mer found there is not solid, ntedit.cpp:467-468
				probe_values(S.io.rep, ng, solid, hu);
				for (uint32_t g = 0; g < ng; g++) {
					if (((solid >> g) & 1u) && S.pval[g][ln] != 0) {
						solid &= ~(1u << g);
					}
				}
			}
			if (solid & premask) {
				pre_ok = 1;
			}
#if defined(__CUDA_ARCH__)
			count += (uint32_t)__popc(solid & ~premask);
#else
			count += (uint32_t)__builtin_popcount(solid & ~premask);
#endif
		}
	}

	// ---------------------------------------------------------------- one candidate, one lane
	// Rolls the candidate's k-mers and classifies the sampled ones.  CK_SOLID: pre_ok = the changed k-mer itself is
	// present && solid, count = number of sampled k-mers that are.  CK_CHECK: raw values go to S.chk[] (single writer:
	// the lane that owns the check candidate).  CK_SITE: count = 1 when the (single) sample is a site.
	NTB_FN_NOINLINE void eval_cand(const Cand& cd, uint32_t& pre_ok, uint32_t& count)
	{
		HashState s = S.hs;
		if (cd.change) {
			w_changelast(s, S.draft, cd.X);
		}
		const uint32_t ln = lane_id();
		const uint32_t Q = cd.Q, nsyn = cd.c, d = cd.d, period = cd.period, kind = cd.kind;
		const uint64_t syn_cls = cd.syn_cls;
		uint32_t r = 0, jc = cd.first, nchk = 0;
		bool pre = cd.pre != 0;
		const uint32_t n_rolls = S.n_rolls;
		const uint32_t patch_idx = cd.patch ? S.patch_idx : NONE32;
		const uint32_t x_cls = cls_of(cd.X);
		pre_ok = 0;
		count = 0;
		while (pre || r < Q) {
			// collect up to PROBE_G sampled k-mers
			uint32_t ng = 0, premask = 0;
			if (pre) {
				S.hb[0][ln] = hash_canonical(s);
				ng = 1;
				premask = 1;
				pre = false;
			}
			while (r < Q && ng < (uint32_t)PROBE_G) {
				uint32_t ci;
				if (r < nsyn) {
					ci = (uint32_t)(syn_cls >> (8 * r)) & 0xFFu;
				} else {
					const uint32_t j = d + r - nsyn;
					if (j >= n_rolls) {
						r = Q; // this roll and every later one fails (ntedit.cpp:1216-1247): nothing more is counted
						break;
					}
					ci = S.lin_in_c[j];
				}
				const uint32_t co = r == patch_idx ? x_cls : S.lin_out_c[r];
				roll_cls(s, co, ci);
				const bool samp = jc == 0;
				jc = samp ? period - 1 : jc - 1;
				r++;
				if (samp) {
					S.hb[ng][ln] = hash_canonical(s);
					ng++;
				}
			}
			if (ng == 0) {
				break;
			}
			finish_group(kind, ng, premask, PROBE_HU, pre_ok, count, nchk);
		}
		if (kind == CK_CHECK) {
			S.chk_n = nchk;
		}
	}

	// ---------------------------------------------------------------- events (leader)
	NTB_FN_NOINLINE void emit(uint8_t kind, uint8_t flags, uint8_t draft, const Site& s)
	{
		uint32_t idx;
#if defined(__CUDA_ARCH__)
		idx = atomicAdd(&S.io.ctr->n_events, 1u);
#else
		idx = S.io.ctr->n_events++;
#endif
		if (idx >= S.io.ev_cap) {
			S.status |= ST_EV_OVERFLOW;
			S.io.ctr->overflow = 1;
			return;
		}
		Event e;
		e.prev = S.last_event;
		e.t_pos = S.t.pos;
		e.advance = S.anchored ? NONE32 : S.adv;
		e.support = (uint16_t)s.best_support;
		e.altsupp[0] = (uint16_t)s.altsupp1;
		e.altsupp[1] = (uint16_t)s.altsupp2;
		e.altsupp[2] = (uint16_t)s.altsupp3;
		e.kind = kind;
		e.flags = flags;
		e.draft = draft;
		e.base = s.best_sub;
		e.altbase[0] = s.altbase1;
		e.altbase[1] = s.altbase2;
		e.altbase[2] = s.altbase3;
		e.indel_len = s.indel_len;
		for (int i = 0; i < 5; i++) {
			e.indel[i] = s.indel[i];
		}
		e.pad_ = 0;
		S.io.events[idx] = e;
		S.last_event = idx;
		S.n_events++;
		S.adv = 0;
		S.anchored = false;
	}

	// ---------------------------------------------------------------- pieces of makeEdit that need the rope (leader)
	// getPrevInsertion, ntedit.cpp:907-922: reverse-complemented run of inserted characters left of the tail
	NTB_FN uint32_t prev_insertion(char* out) const
	{
		uint32_t ni = S.t.ni, n = 0;
		if ((ni < S.nn && S.ty[ni] == 0 && S.t.pos == S.sp[ni]) || (ni < S.nn && S.ty[ni] == 1)) {
			ni--;
		}
		while (ni < S.nn && S.ty[ni] == 1 && n < (uint32_t)PREVCAP - 8) {
			const unsigned char c = S.ch[ni];
			const unsigned cc = code_f(c);
			out[n++] = cc == 0 ? 'T' : cc == 1 ? 'G' : cc == 2 ? 'C' : (cc == 3 && (c | 0x20) == 't') ? 'A' : 'N';
			ni--;
		}
		return n;
	}

	// isRepeatInsertion, ntedit.cpp:561-596
	NTB_FN static bool is_repeat(const char* s, int n)
	{
		if (n <= 0) {
			return false;
		}
		uint16_t lps[PREVCAP];
		int l = 0, i = 1;
		lps[0] = 0;
		while (i < n) {
			if (s[i] == s[l]) {
				lps[i++] = (uint16_t)++l;
			} else if (l != 0) {
				l = lps[l - 1];
			} else {
				lps[i++] = 0;
			}
		}
		const int last = lps[n - 1];
		return last > 0 && n % (n - last) == 0;
	}

	// would makeEdit's case 2 take one of its "skipped_repeat" branches (ntedit.cpp:1315-1380)?  Those branches end the
	// contig (findAcceptedKmer cannot succeed after the removal), so the walker only has to detect them.
	NTB_FN_NOINLINE bool insertion_guard_fires(const Site& s) const
	{
		char prev[PREVCAP];
		uint32_t np = prev_insertion(prev);
		const uint32_t nb = s.indel_len;
		if (np + nb < P.k) {
			return false;
		}
		if (is_repeat(prev, (int)np) || np + nb >= P.insertion_cap) {
			return true;
		}
		for (uint32_t w = 0; w < nb; w++) {
			for (uint32_t q = np; q > 0; q--) {
				prev[q] = prev[q - 1];
			}
			const unsigned cc = code_f((unsigned char)s.indel[w]);
			prev[0] = cc == 0 ? 'T' : cc == 1 ? 'G' : cc == 2 ? 'C' : (cc == 3 && (s.indel[w] | 0x20) == 't') ? 'A' : 'N';
			np++;
			if (is_repeat(prev, (int)np)) {
				return true;
			}
		}
		return false;
	}

	// i-th string of ntedit.cpp:203-348 for first base `first`: all words of length 1..5 over ACGT that start with
	// `first`, ordered by (length, lexicographic A<C<G<T).  Chars are packed little-endian; returns the length.
	NTB_FN static uint32_t indel_string(unsigned char first, uint32_t q, uint64_t& packed)
	{
		uint32_t len = 1, start = 0, count = 1;
		while (q >= start + count) {
			start += count;
			count <<= 2;
			len++;
		}
		uint32_t r = q - start;
		uint64_t p = 0;
		for (uint32_t i = len - 1; i >= 1; i--) {
			const uint32_t d = r & 3;
			const uint64_t c = d == 0 ? 'A' : d == 1 ? 'C' : d == 2 ? 'G' : 'T';
			p |= c << (8 * i);
			r >>= 2;
		}
		packed = p | (uint64_t)first;
		return len;
	}

	// substitution candidates, ntedit.cpp:178-199; returns up to 4 bases packed little-endian, 0-terminated
	NTB_FN uint32_t candidates(unsigned char draft) const
	{
#define NTB_PACK(a, b, c, d) ((uint32_t)(a) | ((uint32_t)(b) << 8) | ((uint32_t)(c) << 16) | ((uint32_t)(d) << 24))
		switch (draft) {
		case 'A': return NTB_PACK('T', 'C', 'G', 0);
		case 'T': return NTB_PACK('A', 'C', 'G', 0);
		case 'C': return NTB_PACK('A', 'T', 'G', 0);
		case 'G': return NTB_PACK('A', 'T', 'C', 0);
		default: break;
		}
		if (snv_()) {
			return is_acc(draft) || draft == 'N' ? NTB_PACK('A', 'T', 'C', 'G') : 0u;
		}
		switch (draft) {
		case 'R': return NTB_PACK('T', 'C', 0, 0);
		case 'Y': return NTB_PACK('A', 'G', 0, 0);
		case 'S': return NTB_PACK('A', 'T', 0, 0);
		case 'W': return NTB_PACK('C', 'G', 0, 0);
		case 'K': return NTB_PACK('A', 'C', 0, 0);
		case 'M': return NTB_PACK('T', 'G', 0, 0);
		case 'B': return NTB_PACK('A', 0, 0, 0);
		case 'D': return NTB_PACK('C', 0, 0, 0);
		case 'H': return NTB_PACK('G', 0, 0, 0);
		case 'V': return NTB_PACK('T', 0, 0, 0);
		case 'N': return NTB_PACK('A', 'T', 'C', 'G');
		default: return 0u;
		}
#undef NTB_PACK
	}

	// ---------------------------------------------------------------- linearisation of the rope around the window
	// Simulates up to `want` calls of roll() from the current cursors without hashing: lin_out[m] / lin_in[m] are the chars
	// the (m+1)-th roll pushes out / pulls in, n_rolls the number of rolls that succeed.  With `check` it also replays the
	// control flow of the check-missing loop (ntedit.cpp:1826-1858): n_check iterations complete, dnf = do_not_fix.
	// Called by every lane.  When the window sits on a single position node that runs to the end of the contig the chars
	// are plain text (plus in-place substitutions) and the lanes fetch them in parallel; otherwise the leader walks the rope.
	NTB_FN void linearise(uint32_t want, bool check)
	{
		const uint32_t k = P.k;
		NTB_LEADER_BEGIN
		S.tail_is_pos = S.t.ni < S.nn && S.ty[S.t.ni] == 0;
		S.tail_is_chr = S.t.ni < S.nn && S.ty[S.t.ni] == 1;
		S.patch_idx = NONE32;
		S.n_rolls = 0;
		if (check) {
			S.n_check = 0;
			S.dnf = false;
		}
		S.lin_simple = S.nn == 1 && S.ty[0] == 0 && S.h.ni == 0 && S.t.ni == 0 && S.t.pos - S.h.pos == k - 1 &&
		               S.ep[0] == S.io.len - 1 && S.t.pos < S.io.len && S.sp[0] <= S.h.pos;
		if (S.lin_simple) {
			const uint32_t avail = S.io.len - 1 - S.t.pos;
			S.n_rolls = avail < want ? avail : want;
			S.patch_idx = k - 1;
		} else {
			linearise_rope(want, check);
		}
		NTB_LEADER_END
		if (!S.lin_simple) {
			return;
		}
		const uint32_t n = S.n_rolls;
		uint32_t bad = k; // first iteration of the check loop that does not complete
		for (uint32_t m0 = 0; m0 < n || (check && m0 < k); m0 += lane_count()) {
			const uint32_t m = m0 + lane_id();
			if (m < n) {
				const unsigned char in = rd(S.t.pos + 1 + m);
				const unsigned char out = rd(S.h.pos + m);
				S.lin_out[m] = out;
				S.lin_in[m] = in;
				S.lin_out_c[m] = cls_of(out);
				S.lin_in_c[m] = cls_of(in);
				if (m < k && !is_acc(in) && m < bad) {
					bad = m;
				}
			} else if (m < k && m < bad) {
				bad = m; // roll fails at the end of the contig
			}
		}
		if (check) {
			bad = warp_min(bad);
			NTB_LEADER_BEGIN
			S.n_check = bad;
			S.dnf = bad < k;
			NTB_LEADER_END
		} else {
			warp_sync();
		}
	}

	// leader: the general case of linearise()
	NTB_FN_NOINLINE void linearise_rope(uint32_t want, bool check)
	{
		const uint32_t k = P.k;
		bool check_open = check;
		Cursor hh = S.h, tt = S.t;
		for (uint32_t m = 0; m < want; m++) {
			if (check_open && m >= k) {
				check_open = false;
			}
			if (check_open && hh.pos >= S.io.len) {
				check_open = false; // the loop condition ends the check loop without do_not_fix
			}
			// roll()
			if (hh.pos >= S.io.len || hh.ni >= S.nn) {
				if (check_open) {
					S.dnf = true;
				}
				break;
			}
			const bool at_tail = S.tail_is_pos ? (S.ty[hh.ni] == 0 && hh.pos == S.t.pos) : S.tail_is_chr ? hh.ni == S.t.ni : false;
			if (at_tail && S.patch_idx == NONE32) {
				S.patch_idx = m;
			}
			const unsigned char out = cchar(hh);
			step(hh);
			if (tt.pos >= S.io.len || tt.ni >= S.nn) {
				if (check_open) {
					S.dnf = true;
				}
				break;
			}
			step(tt);
			if (tt.pos >= S.io.len || tt.ni >= S.nn) {
				if (check_open) {
					S.dnf = true;
				}
				break;
			}
			const unsigned char in = cchar(tt);
			S.lin_out[m] = out;
			S.lin_in[m] = in;
			S.lin_out_c[m] = cls_of(out);
			S.lin_in_c[m] = cls_of(in);
			S.n_rolls = m + 1;
			if (check_open) {
				if (!is_acc(in)) {
					S.dnf = true;
					check_open = false;
				} else {
					S.n_check = m + 1;
				}
			}
		}
	}

	// ---------------------------------------------------------------- plain rolls
	// plain_f[R] / plain_r[R] = forward / reverse hash after R calls of roll() from the current state (no edit applied),
	// R = 0 .. min(n_rolls, cap).  The two strands are independent chains: lane 0 rolls the forward one, lane 1 the reverse.
	NTB_FN void compute_plain(uint32_t want)
	{
		const uint32_t cap = (uint32_t)WalkerState<NCAP>::PLAIN_CAP - 1;
		const uint32_t k = P.k;
		uint32_t n = S.n_rolls < want ? S.n_rolls : want;
		if (n > cap) {
			n = cap;
		}
		warp_sync();
		const uint32_t ln = lane_id();
		bool direct = S.n_rolls >= k && k <= ROT_STRIDE;
		if (direct) {
			// the window after R rolls is lin_out[R..k) ++ lin_in[0..R): hash it from the rotation table, one window per lane
			const uint64_t* rot = rot_();
			uint32_t bad = 0;
			for (uint32_t R = ln; R <= n; R += lane_count()) {
				uint64_t f = 0, r = 0;
				// lin_out_c and lin_in_c are members of the same object: one load through a selected index instead of a branch
				const uint8_t* const lo = S.lin_out_c;
				const uint32_t in_off = (uint32_t)(S.lin_in_c - S.lin_out_c);
				for (uint32_t i = 0; i < k; i++) {
					const uint32_t m = R + i;
					const uint32_t c = lo[m < k ? m : m - k + in_off];
					f ^= rot[(c & 7u) * ROT_STRIDE + (k - 1 - i)];
					r ^= rot[((c >> 3) & 7u) * ROT_STRIDE + i];
				}
				S.plain_f[R] = f;
				S.plain_r[R] = r;
				if (R == 0 && (f != S.hs.fh || r != S.hs.rh)) {
					bad = 1; // the rolled state is not the hash of the window's chars: roll instead
				}
			}
			direct = warp_or(bad) == 0;
		}
		if (!direct) {
			warp_sync();
			if (ln == 0) {
				uint64_t f = S.hs.fh;
				S.plain_f[0] = f;
				for (uint32_t r = 0; r < n; r++) {
					f = srol1(f) ^ S.seed_tab[S.lin_in_c[r] & 7u] ^ S.rotk_tab[S.lin_out_c[r] & 7u];
					S.plain_f[r + 1] = f;
				}
			}
			if (ln == (lane_count() > 1 ? 1u : 0u)) {
				uint64_t rv = S.hs.rh;
				S.plain_r[0] = rv;
				for (uint32_t r = 0; r < n; r++) {
					rv = sror1(rv ^ S.rotk_tab[(S.lin_in_c[r] >> 3) & 7u] ^ S.seed_tab[(S.lin_out_c[r] >> 3) & 7u]);
					S.plain_r[r + 1] = rv;
				}
			}
		}
		if (ln == 0) {
			S.n_plain = n;
		}
		warp_sync();
	}

	// Phase 1 without rolling, in dependent steps so that nothing is probed that the decision does not read:
	//   phase_fast_check : the check-missing subset (ntedit.cpp:1826-1858)                               -> S.chk[]
	//   phase_fast_subs  : the substitution gates (ntedit.cpp:1923-1928)                                 -> S.gate[]
	//                      then, for the candidates whose gate passed (all of them in mode 2), the trial
	//                      (ntedit.cpp:1936-1981) with EVERY following window                            -> S.tsolid / S.tsite / S.sup
	// Most sites stop after the first step (no attempt) or try one candidate: ~40 k-mers instead of 9 + 3 * 26.  The
	// k-mers of the first two steps are mostly absent and are probed one hash function per pass (btllib's early exit,
	// ntedit.cpp:368-371); a direct probe costs a 128-byte DRAM line on B200.
	// Every k-mer is a plain one (compute_plain) plus, for a trial after R rolls, the NTMC64_changelast terms rotated R
	// times; after k rolls the substituted base has left the window and the k-mer is the plain one.
	// Requires patch_idx == k-1 (the tail's slot leaves the window with the k-th roll), which every consistent rope gives.
	NTB_FN void phase_fast_check()
	{
		const uint32_t jump = P.jump, ln = lane_id();
		const uint32_t nC = S.n_check ? (S.n_check - 1) / jump + 1 : 0;   // samples q = 0, jump, .. < n_check
		warp_sync();
		for (uint32_t j0 = 0; j0 < nC; j0 += lane_count() * PROBE_G) {
			uint32_t ng = 0;
			for (uint32_t j = j0 + ln; j < nC && ng < (uint32_t)PROBE_G; j += lane_count()) {
				const uint32_t R = j * jump + 1;
				S.hb[ng][ln] = S.plain_f[R] + S.plain_r[R];
				ng++;
			}
			if (ng) {
				probe_values(S.io.bloom, ng, (1u << ng) - 1u, 1);
				uint32_t g = 0;
				for (uint32_t j = j0 + ln; j < nC && g < ng; j += lane_count(), g++) {
					S.chk[j] = S.pval[g][ln];
				}
			}
		}
		warp_sync();
		NTB_LEADER_BEGIN
		S.chk_n = nC;
		S.p1_fast = true;
		NTB_LEADER_END
	}

	NTB_FN void phase_fast_subs()
	{
		const uint32_t k = P.k, jump = P.jump, ln = lane_id();
		const uint32_t n_sub = S.n_rolls < k ? S.n_rolls : k;
		uint32_t ncand = 0;
		while (ncand < 4 && ((S.cands >> (8 * ncand)) & 0xFF) != 0) {
			ncand++;
		}
		const uint32_t df = code_f(S.draft), dr = code_r(S.draft);
		const uint64_t* rot = rot_();
		warp_sync();
		// ---- gates: the window with its last base replaced (mode 2 tries every candidate: no gate is read)
		if (P.mode != 2) {
			for (uint32_t j0 = 0; j0 < ncand; j0 += lane_count() * PROBE_G) {
				uint32_t ng = 0;
				for (uint32_t j = j0 + ln; j < ncand && ng < (uint32_t)PROBE_G; j += lane_count()) {
					const unsigned char X = (unsigned char)((S.cands >> (8 * j)) & 0xFF);
					const uint64_t f = S.plain_f[0] ^ rot[df * ROT_STRIDE] ^ rot[code_f(X) * ROT_STRIDE];
					const uint64_t r = S.plain_r[0] ^ rot[dr * ROT_STRIDE + (k - 1)] ^ rot[code_r(X) * ROT_STRIDE + (k - 1)];
					S.hb[ng][ln] = f + r;
					ng++;
				}
				if (ng) {
					probe_values(S.io.bloom, ng, (1u << ng) - 1u, 1);
					uint32_t solid = 0;
					for (uint32_t g = 0; g < ng; g++) {
						if (solid_value(S.pval[g][ln])) {
							solid |= 1u << g;
						}
					}
					if (h_rep_() && solid) {
						// solid k-mers additionally must be absent from the secondary filter (-e), ntedit.cpp:467-468
						probe_values(S.io.rep, ng, solid, 1);
						for (uint32_t q = 0; q < ng; q++) {
							if (((solid >> q) & 1u) && S.pval[q][ln] != 0) {
								solid &= ~(1u << q);
							}
						}
					}
					uint32_t g = 0;
					for (uint32_t j = j0 + ln; j < ncand && g < ng; j += lane_count(), g++) {
						S.gate[j] = (uint8_t)((solid >> g) & 1u);
					}
				}
			}
			warp_sync();
		}
		// ---- trials: the k-mer after every roll R = 1 .. n_sub of every tried candidate -- the sampled ones are the trial
		// (ntedit.cpp:1960-1968), all of them are the windows the main loop visits next if the candidate is accepted
		uint32_t act[4], nact = 0;
		for (uint32_t c = 0; c < ncand; c++) {
			if (P.mode == 2 || S.gate[c]) {
				act[nact++] = c;
			}
		}
		const uint32_t njobs = nact * n_sub;
		for (uint32_t j0 = 0; j0 < njobs; j0 += lane_count() * PROBE_G) {
			uint32_t ng = 0;
			for (uint32_t j = j0 + ln; j < njobs && ng < (uint32_t)PROBE_G; j += lane_count()) {
				const uint32_t c = act[j / n_sub], R = j % n_sub + 1;
				const unsigned char X = (unsigned char)((S.cands >> (8 * c)) & 0xFF);
				uint64_t f = S.plain_f[R], r = S.plain_r[R];
				if (R < k) {
					f ^= rot[df * ROT_STRIDE + R] ^ rot[code_f(X) * ROT_STRIDE + R];
					r ^= rot[dr * ROT_STRIDE + (k - 1 - R)] ^ rot[code_r(X) * ROT_STRIDE + (k - 1 - R)];
				}
				S.hb[ng][ln] = f + r;
				ng++;
			}
			if (ng) {
				probe_values(S.io.bloom, ng, (1u << ng) - 1u, PROBE_HU);
				uint32_t solid = 0, site = 0;
				for (uint32_t g = 0; g < ng; g++) {
					const uint32_t v = S.pval[g][ln];
					if (solid_value(v)) {
						solid |= 1u << g;
					}
					if (is_site_value(v)) {
						site |= 1u << g;
					}
				}
				if (h_rep_() && solid) {
					probe_values(S.io.rep, ng, solid, PROBE_HU);
					for (uint32_t q = 0; q < ng; q++) {
						if (((solid >> q) & 1u) && S.pval[q][ln] != 0) {
							solid &= ~(1u << q);
						}
					}
				}
				uint32_t g = 0;
				for (uint32_t j = j0 + ln; j < njobs && g < ng; j += lane_count(), g++) {
					const uint32_t c = act[j / n_sub], R = j % n_sub + 1;
					S.tsolid[c][R - 1] = (uint8_t)((solid >> g) & 1u);
					S.tsite[c][R - 1] = (uint8_t)((site >> g) & 1u);
				}
			}
		}
		warp_sync();
		NTB_LEADER_BEGIN
		for (uint32_t c = 0; c < 4; c++) {
			uint32_t cnt = 0;
			const bool tried = c < ncand && (P.mode == 2 || S.gate[c]);
			if (tried) {
				for (uint32_t q = 0; q < n_sub; q += jump) {
					cnt += S.tsolid[c][q];
				}
			} else {
				S.gate[c] = 0;
			}
			S.sup[c] = cnt;
		}
		NTB_LEADER_END
	}

	// ---------------------------------------------------------------- phases: one candidate per lane
	// phase 1: job 0 = the check-missing subset, job 1+ci = substitution candidate ci (gate + trial)
	NTB_FN void phase_check_and_subs()
	{
		const uint32_t njobs = 5;
		warp_sync();
		for (uint32_t j = lane_id(); j < njobs; j += lane_count()) {
			Cand cd;
			cd.syn_cls = 0;
			cd.c = 0;
			cd.d = 0;
			cd.first = 0;
			cd.period = P.jump;
			cd.X = 0;
			cd.change = 0;
			cd.pre = 0;
			cd.patch = 0;
			uint32_t pre_ok = 0, count = 0;
			if (j == 0) {
				cd.kind = CK_CHECK;
				cd.Q = S.n_check;
				eval_cand(cd, pre_ok, count);
			} else {
				const unsigned char sub = (unsigned char)((S.cands >> (8 * (j - 1))) & 0xFF);
				// candidates are 0-terminated: nothing after the first 0 is tried
				bool live = sub != 0;
				for (uint32_t q = 0; q + 1 < j; q++) {
					if (((S.cands >> (8 * q)) & 0xFF) == 0) {
						live = false;
					}
				}
				if (live) {
					cd.kind = CK_SOLID;
					cd.X = sub;
					cd.change = 1;
					cd.pre = 1;
					cd.patch = 1;
					cd.Q = S.n_rolls < P.k ? S.n_rolls : P.k;
					eval_cand(cd, pre_ok, count);
				}
				S.gate[j - 1] = (uint8_t)pre_ok;
				S.sup[j - 1] = count;
			}
		}
		warp_sync();
	}

	// ---- tryIndels candidates (ntedit.cpp:1548-1744)
	// All insertion candidates of a site roll the same outgoing and (after their inserted chars) the same incoming bases;
	// ntHash is GF(2)-linear in the seeds, so the hash of candidate (string, sample) is
	//     base(L, sample)  ^  terms of the index char / draft char (NTMC64_changelast)  ^  terms of the inserted chars,
	// where base(L, .) is the hash state of "an insertion of length L whose chars have no seed" and every term is one
	// entry of the rotation table.  Lanes 0..4 roll the five bases once per site; the 341 x 3 candidates then cost a few
	// table look-ups per sampled k-mer instead of k-1 rolls each.
	NTB_FN void compute_ins_bases()
	{
		const uint32_t k = P.k, jump = P.jump;
		const uint32_t ns = (k - 1 + jump - 1) / jump; // samples after rolls q = 0, jump, 2 jump, ... < k-1
		NTB_LEADER_BEGIN
		S.ins_fast = ns <= (uint32_t)WalkerState<NCAP>::NSMAX && k + 4 < ROT_STRIDE;
		S.bases_ready = true;
		NTB_LEADER_END
		if (!S.ins_fast) {
			return;
		}
		for (uint32_t L = 1 + lane_id(); L <= 5; L += lane_count()) {
			HashState st = S.hs;
			uint32_t jc = 0, nv = 0;
			for (uint32_t r = 0; r + 1 < k; r++) {
				uint32_t ci = 0x24; // no seed on either strand
				if (r >= L) {
					const uint32_t j = r - L;
					if (j >= S.n_rolls) {
						break;
					}
					ci = S.lin_in_c[j];
				}
				roll_cls(st, S.lin_out_c[r], ci);
				if (jc == 0) {
					S.ins_base_f[L - 1][nv] = st.fh;
					S.ins_base_r[L - 1][nv] = st.rh;
					nv++;
					jc = jump;
				}
				jc--;
			}
			S.ins_nvalid[L - 1] = (uint8_t)nv;
		}
		warp_sync();
	}

	// support of insertion candidate i for S.index_char from the bases (this lane)
	NTB_FN uint32_t eval_insertion_fast(uint32_t i)
	{
		const uint32_t ln = lane_id();
		const uint32_t k = P.k, jump = P.jump;
		uint64_t packed;
		const uint32_t L = indel_string(S.index_char, i, packed);
		// incoming chars of the synthetic rolls: string[1..L-1] then the draft char (ntedit.cpp:1583-1606)
		uint32_t fcode[5], rcode[5];
		for (uint32_t q = 0; q < L; q++) {
			const unsigned char c = q + 1 < L ? (unsigned char)((packed >> (8 * (q + 1))) & 0xFF) : S.draft;
			fcode[q] = code_f(c);
			rcode[q] = code_r(c);
		}
		const uint32_t xf = code_f(S.index_char), xr = code_r(S.index_char);
		const uint32_t df = code_f(S.draft), dr = code_r(S.draft);
		const uint64_t* rot = rot_();
		const uint32_t nv = S.ins_nvalid[L - 1];
		uint32_t pre_ok = 0, count = 0, nchk = 0;
		for (uint32_t s0 = 0; s0 < nv; s0 += PROBE_G) {
			const uint32_t ng = nv - s0 < (uint32_t)PROBE_G ? nv - s0 : (uint32_t)PROBE_G;
			for (uint32_t g = 0; g < ng; g++) {
				const uint32_t R = (s0 + g) * jump + 1; // rolls done when the sample is taken
				uint64_t f = S.ins_base_f[L - 1][s0 + g] ^ rot[df * ROT_STRIDE + R] ^ rot[xf * ROT_STRIDE + R];
				uint64_t r = S.ins_base_r[L - 1][s0 + g] ^ rot[dr * ROT_STRIDE + (k - 1 - R)] ^ rot[xr * ROT_STRIDE + (k - 1 - R)];
				const uint32_t m = L < R ? L : R;
				for (uint32_t q = 0; q < m; q++) {
					f ^= rot[fcode[q] * ROT_STRIDE + (R - 1 - q)];
					r ^= rot[rcode[q] * ROT_STRIDE + (k - R + q)];
				}
				S.hb[g][ln] = f + r;
			}
			finish_group(CK_SOLID, ng, 0, 1, pre_ok, count, nchk);
		}
		return count;
	}

	// eval_insertion_fast for the common configuration -- bit filter, no secondary filter: the candidate's sampled k-mers
	// are hashed from the bases and probed two at a time through registers (first hash function of both in flight
	// together; the few k-mers whose first bit is set then check their other bits, stopping at the first clear one --
	// btllib's early exit, ntedit.cpp:368-371).  Same count as eval_insertion_fast for every candidate that can still
	// reach the edit threshold; a smaller one (still below the threshold) for the others.
	NTB_FN uint32_t eval_insertion_bits(uint32_t i)
	{
		const uint32_t k = P.k, jump = P.jump;
		uint64_t packed;
		const uint32_t L = indel_string(S.index_char, i, packed);
		// seed codes of the synthetic incoming chars: string[1..L-1], then the draft char (ntedit.cpp:1583-1606)
		uint32_t fc[5], rc[5];
#pragma unroll
		for (uint32_t q = 0; q < 5; q++) {
			const unsigned char c = q + 1 < L ? (unsigned char)((packed >> (8 * (q + 1))) & 0xFF) : S.draft;
			fc[q] = code_f(c) * ROT_STRIDE;
			rc[q] = code_r(c) * ROT_STRIDE;
		}
		const uint32_t xf = code_f(S.index_char) * ROT_STRIDE, xr = code_r(S.index_char) * ROT_STRIDE;
		const uint32_t df = code_f(S.draft) * ROT_STRIDE, dr = code_r(S.draft) * ROT_STRIDE;
		const uint64_t* rot = rot_();
		const uint64_t* bf = S.ins_base_f[L - 1];
		const uint64_t* br = S.ins_base_r[L - 1];
		const uint32_t nv = S.ins_nvalid[L - 1];
		const FilterView& F = S.io.bloom;
		const uint32_t hn = F.hash_num;
		const uint32_t thr = P.thr_edit;
		uint32_t count = 0;
		// two samples per pass: the walker is bound by instruction fetch, a wider unroll is slower (measured: 4 -> 167 ms, 2 -> 162 ms)
		for (uint32_t s0 = 0; s0 < nv; s0 += 2) {
			// a candidate that cannot reach the edit threshold any more only ever reads as "below threshold"
			// (try_indels tests meets_edit() before it looks at the count): stop probing it
			if (count + (nv - s0) < thr) {
				break;
			}
			uint64_t hv[2];
			uint32_t got[2], sh[2];
#pragma unroll
			for (uint32_t g = 0; g < 2; g++) {
				const uint32_t sidx = s0 + g < nv ? s0 + g : nv - 1; // a padding lane repeats the last sample and is not counted
				const uint32_t R = sidx * jump + 1;                   // rolls done when the sample is taken
				uint64_t f = bf[sidx] ^ rot[df + R] ^ rot[xf + R];
				uint64_t r = br[sidx] ^ rot[dr + (k - 1 - R)] ^ rot[xr + (k - 1 - R)];
				const uint32_t m = L < R ? L : R;
#pragma unroll
				for (uint32_t q = 0; q < 5; q++) {
					// branch-free: a char that has not entered the window yet (q >= m) reads the all-zero "no seed" row
					const bool in = q < m;
					f ^= rot[in ? fc[q] + (R - 1 - q) : 4u * ROT_STRIDE];
					r ^= rot[in ? rc[q] + (k - R + q) : 4u * ROT_STRIDE];
				}
				hv[g] = f + r;
				const uint64_t slot = slot_(F, hv[g]);
				sh[g] = (uint32_t)slot & 7u;
				got[g] = probe_byte(F.data + (slot >> 3));
			}
#pragma unroll
			for (uint32_t g = 0; g < 2; g++) {
				if (s0 + g < nv && ((got[g] >> sh[g]) & 1u)) {
					bool all = true;
					for (uint32_t h = 1; h < hn && all; h++) {
						const uint64_t slot = slot_(F, hash_extend(hv[g], k, h));
						all = ((probe_byte(F.data + (slot >> 3)) >> ((uint32_t)slot & 7u)) & 1u) != 0;
					}
					count += all ? 1u : 0u;
				}
			}
		}
		return count;
	}

	// insertion candidates [i0, i1) of tryIndels for S.index_char: one per lane, 32 at a time
	NTB_FN void phase_insertions(uint32_t i0, uint32_t i1)
	{
		warp_sync();
		const bool bits_only = S.ins_fast && !counting_() && !h_rep_();
		for (uint32_t i = i0 + lane_id(); i < i1; i += lane_count()) {
			if (bits_only) {
				S.ins_sup[i] = (uint8_t)eval_insertion_bits(i);
				continue;
			}
			if (S.ins_fast) {
				S.ins_sup[i] = (uint8_t)eval_insertion_fast(i);
				continue;
			}
			// insertion string + the draft char, rolled base by base (ntedit.cpp:1583-1645)
			Cand cd;
			cd.kind = CK_SOLID;
			cd.change = 1;
			cd.patch = 0;
			cd.period = P.jump;
			uint32_t pre_ok = 0, count = 0;
			uint64_t packed;
			const uint32_t len = indel_string(S.index_char, i, packed);
			cd.X = S.index_char;
			uint64_t sc = 0;
			for (uint32_t q = 1; q < len; q++) {
				sc |= (uint64_t)cls_of((unsigned char)((packed >> (8 * q)) & 0xFF)) << (8 * (q - 1));
			}
			cd.syn_cls = sc | ((uint64_t)cls_of(S.draft) << (8 * (len - 1)));
			cd.c = len;
			cd.d = 0;
			cd.pre = 0;
			cd.first = 0;
			cd.Q = P.k - 1;
			eval_cand(cd, pre_ok, count);
			S.ins_sup[i] = (uint8_t)count;
		}
		warp_sync();
	}

	// tryDeletion (ntedit.cpp:1451-1545) for n = nd0 .. nd0 + ndel - 1: one per lane
	NTB_FN void phase_deletions(uint32_t nd0, uint32_t ndel)
	{
		warp_sync();
		for (uint32_t j = lane_id(); j < ndel; j += lane_count()) {
			const uint32_t n = nd0 + j;
			Cand cd;
			cd.kind = CK_SOLID;
			cd.change = 1;
			cd.patch = 0;
			cd.period = P.jump;
			cd.X = n >= 1 && n - 1 < S.n_rolls ? S.lin_in[n - 1] : 0;
			cd.syn_cls = 0;
			cd.c = 0;
			cd.d = n;
			cd.pre = 1;
			cd.first = P.jump - 1;
			cd.Q = P.k - 2;
			uint32_t pre_ok = 0, count = 0;
			eval_cand(cd, pre_ok, count);
			S.del_sup[n] = (uint8_t)(pre_ok + count);
		}
		warp_sync();
	}

	// tryIndels, ntedit.cpp:1548-1744.  Candidates are evaluated in chunks (one per lane), the leader then walks the
	// iterations of the chunk in the reference's order: insertion i, then deletion num_deletions (shared across the
	// index bases of a site, ntedit.cpp:1692-1729).
	NTB_FN bool try_indels()
	{
		const uint32_t T = P.max_ins_tries;
		if (T == 0) {
			return false;
		}
		if (!S.bases_ready) {
			compute_ins_bases();
		}
		NTB_LEADER_BEGIN
		S.tb_support = S.ta_support = S.tb_type = S.tb_len = 0;
		S.ti_done = false;
		S.ti_ret = false;
		S.ti_i0 = 0;
		// deletions this call can reach: one per iteration while num_deletions <= max_deletions
		S.ti_nd0 = S.num_deletions;
		uint32_t nd = 0;
		if (S.num_deletions <= P.max_deletions) {
			nd = P.max_deletions - S.num_deletions + 1;
			if (nd > T) {
				nd = T;
			}
		}
		S.ti_ndel = nd;
		NTB_LEADER_END
		if (S.ti_ndel) {
			phase_deletions(S.ti_nd0, S.ti_ndel);
		}
		while (S.ti_i0 < T) {
			// first hit wins in mode 0: evaluate one warp's worth of candidates at a time
			const uint32_t i0 = S.ti_i0;
			const uint32_t i1 = (P.mode == 0 && i0 + 32 < T) ? i0 + 32 : T;
			phase_insertions(i0, i1);
			// does any candidate of the chunk reach its threshold?  (usually none does)
			uint32_t any = 0;
			for (uint32_t i = i0 + lane_id(); i < i1; i += lane_count()) {
				if (meets_edit(S.ins_sup[i])) {
					any = 1;
				}
			}
			for (uint32_t n = S.num_deletions + lane_id(); n <= P.max_deletions && n - S.num_deletions < i1 - i0; n += lane_count()) {
				if (S.del_sup[n] >= P.thr_edit_del && S.del_sup[n] > 0) {
					any = 1;
				}
			}
			any = warp_or(any);
			NTB_LEADER_BEGIN
			if (!any) {
				// nothing to select: only the shared deletion counter advances, once per iteration (ntedit.cpp:1692-1729)
				if (S.num_deletions <= P.max_deletions) {
					const uint32_t left = P.max_deletions + 1 - S.num_deletions;
					S.num_deletions += left < i1 - i0 ? left : i1 - i0;
				}
			}
			for (uint32_t i = any ? i0 : i1; i < i1; i++) {
				const uint32_t present = S.ins_sup[i];
				if (meets_edit(present)) {
					uint64_t packed;
					const uint32_t nins = indel_string(S.index_char, i, packed);
					if (P.mode == 0) {
						S.s.best_type = 2;
						for (uint32_t c = 0; c < nins; c++) {
							S.s.indel[c] = (char)((packed >> (8 * c)) & 0xFF);
						}
						S.s.indel_len = (uint8_t)nins;
						S.s.best_support = present;
						S.ti_done = true;
						S.ti_ret = true;
						break;
					}
					if (present >= S.tb_support) {
						if (S.tb_support) {
							S.ta_support = S.tb_support;
						}
						S.tb_type = 2;
						for (uint32_t c = 0; c < nins; c++) {
							S.tb_indel[c] = (char)((packed >> (8 * c)) & 0xFF);
						}
						S.tb_len = nins;
						S.tb_support = present;
					}
				}
				if (S.num_deletions <= P.max_deletions) {
					const uint32_t raw = S.del_sup[S.num_deletions];
					const uint32_t del_support = raw >= P.thr_edit_del ? raw : 0u;
					if (del_support > 0) {
						if (P.mode == 0) {
							S.s.best_type = 3;
							S.s.indel_len = (uint8_t)S.num_deletions;
							S.s.best_support = del_support;
							S.ti_done = true;
							S.ti_ret = true;
							break;
						}
						if (del_support >= S.tb_support) {
							if (S.tb_support) {
								S.ta_support = S.tb_support;
							}
							S.tb_type = 3;
							S.tb_len = S.num_deletions;
							S.tb_support = del_support;
						}
					}
					S.num_deletions++;
				}
			}
			S.ti_i0 = i1;
			NTB_LEADER_END
			if (S.ti_done) {
				return S.ti_ret;
			}
		}
		NTB_LEADER_BEGIN
		if (S.tb_support > 0) {
			if ((P.mode == 2 && S.tb_support > S.s.best_support) || P.mode == 1) {
				S.s.best_type = S.tb_type;
				S.s.indel_len = (uint8_t)S.tb_len;
				if (S.tb_type == 2) {
					for (uint32_t c = 0; c < S.tb_len; c++) {
						S.s.indel[c] = S.tb_indel[c];
					}
				}
				S.s.best_support = S.tb_support;
				S.s.altsupp1 = S.ta_support;
			}
			S.ti_ret = true;
		}
		NTB_LEADER_END
		return S.ti_ret;
	}

	// ---------------------------------------------------------------- one site: ntedit.cpp:1808-2116
	// leader: what the reference does before the check loop, plus the linearisation
	NTB_FN void site_begin()
	{
		S.raw = S.char_in;
		S.draft = to_upper(S.raw);
		S.n_sites++;
		if (S.first_touch == NONE32) {
			S.first_touch = S.t.pos;
		}
		S.cands = candidates(S.draft);
		S.site_ok = true;
		S.chk_n = 0;
		S.bases_ready = false;
		S.p1_fast = false;
		S.skip_advance = false;
	}

	// leader: check-missing verdict (ntedit.cpp:1859-1873) and the site locals (ntedit.cpp:1876-1914)
	NTB_FN bool site_after_check()
	{
		uint32_t missing = 0, there = 0, nmed = 0;
		uint8_t med[KMAX + 1];
		const bool atgc = is_atgc_upper(S.draft);
		for (uint32_t i = 0; i < S.chk_n; i++) {
			const uint32_t c = S.chk[i];
			if (c == 0) {
				missing++;
			} else if (counting_()) {
				if (atgc && c >= P.min_threshold) {
					there++;
					if (nmed < KMAX) {
						med[nmed++] = (uint8_t)c;
					}
				}
			} else if (atgc) {
				there++;
			}
		}
		uint32_t there_median = 0;
		if (counting_() && nmed > 0) {
			// upper median of the collected counts (median(), ntedit.cpp:455-463)
			for (uint32_t a = 1; a < nmed; a++) {
				const uint8_t v = med[a];
				uint32_t b = a;
				while (b > 0 && med[b - 1] > v) {
					med[b] = med[b - 1];
					b--;
				}
				med[b] = v;
			}
			there_median = med[nmed / 2];
		}
		const bool attempt = snv_() || (!S.dnf && (missing >= P.thr_missing || (counting_() && there_median < P.min_threshold)));
		if (!attempt) {
			return false;
		}
		Site& s = S.s;
		s.best_type = 0;
		s.best_support = 0;
		s.altsupp1 = s.altsupp2 = s.altsupp3 = 0;
		s.best_sub = S.stale_best_sub;
		s.altbase1 = S.stale_alt1;
		s.altbase2 = S.stale_alt2;
		s.altbase3 = S.stale_alt3;
		s.indel_len = 0;
		S.num_deletions = 1;
		S.touched = false;
		if (snv_() && meets_edit(there)) {
			s.best_sub = S.draft;
			s.best_support = counting_() ? there_median : there;
		}
		return true;
	}

	// leader: one iteration of the substitution loop (ntedit.cpp:1917-2092) up to the tryIndels call
	NTB_FN void site_candidate(uint32_t ci)
	{
		Site& s = S.s;
		const unsigned char sub = (unsigned char)((S.cands >> (8 * ci)) & 0xFF);
		if (!sub) {
			S.next = NEXT_STOP;
			return;
		}
		S.next = NEXT_CAND;
		if (!(P.mode == 2 || S.gate[ci])) {
			return;
		}
		S.touched = true;
		const uint32_t present = S.sup[ci];
		if (meets_edit(present)) {
			if (present >= s.best_support) {
				if (s.altsupp2) {
					s.altbase3 = s.altbase2;
					s.altsupp3 = s.altsupp2;
				}
				if (s.altsupp1) {
					s.altbase2 = s.altbase1;
					s.altsupp2 = s.altsupp1;
				}
				if (s.best_support) {
					s.altsupp1 = s.best_support;
					s.altbase1 = s.best_sub;
				}
				s.best_type = 1;
				s.best_sub = sub;
				s.best_support = present;
			} else if (!s.altsupp1) {
				s.altbase1 = sub;
				s.altsupp1 = present;
			} else if (!s.altsupp2) {
				if (present < s.altsupp1) {
					s.altbase2 = sub;
					s.altsupp2 = present;
				} else {
					s.altbase2 = s.altbase1;
					s.altsupp2 = s.altsupp1;
					s.altbase1 = sub;
					s.altsupp1 = present;
				}
			} else if (!s.altsupp3) {
				if (present < s.altsupp2) {
					s.altbase3 = sub;
					s.altsupp3 = present;
				} else if (present < s.altsupp1) {
					s.altbase3 = s.altbase2;
					s.altsupp3 = s.altsupp2;
					s.altbase2 = sub;
					s.altsupp2 = present;
				} else {
					s.altbase3 = s.altbase2;
					s.altsupp3 = s.altsupp2;
					s.altbase2 = s.altbase1;
					s.altsupp2 = s.altsupp1;
					s.altbase1 = sub;
					s.altsupp1 = present;
				}
			}
			if (P.mode == 0 || P.mode == 1) {
				return;
			}
		}
		if (P.mode == 2 || s.best_type != 1) {
			S.index_char = sub;
			S.next = NEXT_INDELS;
		}
	}

	// leader: makeEdit, ntedit.cpp:1250-1448.  site_ok = false when the contig is finished (insertion guard fired)
	NTB_FN void site_commit(const uint8_t fl)
	{
		Site& s = S.s;
		const unsigned char draft = S.draft;
		S.stale_best_sub = s.best_sub;
		S.stale_alt1 = s.altbase1;
		S.stale_alt2 = s.altbase2;
		S.stale_alt3 = s.altbase3;
		switch (s.best_type) {
		case 1:
			emit(1, fl, draft, s);
			if (S.tail_is_pos) {
				if (S.ov_n >= (uint32_t)OVCAP) {
					// drop substitutions the head has already passed
					uint32_t w = 0;
					for (uint32_t i = 0; i < S.ov_n; i++) {
						if (S.ov_pos[i] >= S.h.pos) {
							S.ov_pos[w] = S.ov_pos[i];
							S.ov_ch[w] = S.ov_ch[i];
							w++;
						}
					}
					S.ov_n = w;
				}
				if (S.ov_n < (uint32_t)OVCAP) {
					// a later substitution at the same position replaces the earlier one
					uint32_t i = 0;
					for (; i < S.ov_n; i++) {
						if (S.ov_pos[i] == S.t.pos) {
							break;
						}
					}
					S.ov_pos[i] = S.t.pos;
					S.ov_ch[i] = s.best_sub;
					if (i == S.ov_n) {
						S.ov_n++;
					}
				} else {
					S.status |= ST_ROPE_OVERFLOW;
				}
			} else if (S.tail_is_chr) {
				S.ch[S.t.ni] = s.best_sub;
			}
			w_changelast(S.hs, draft, s.best_sub);
			S.la_n = 0;
			if (S.quiet) {
				// none of the k-1 windows that contain the new base is a site (evaluate_site_core): jump behind them
				S.h.pos += P.k;
				S.t.pos += P.k;
				S.adv += P.k;
				S.need_seed = true; // the window is clean: the hash is re-seeded at the next flagged position
				S.skip_advance = true;
			}
			break;
		case 2: {
			emit(2, fl, draft, s);
			if (insertion_guard_fires(s)) {
				S.site_ok = false;
				return;
			}
			rope_insert(S.t.ni, S.t.pos, s.indel, s.indel_len);
			w_changelast(S.hs, draft, (unsigned char)s.indel[0]);
			S.la_n = 0;
			break;
		}
		case 3:
			emit(3, fl, draft, s);
			rope_delete(S.t.ni, S.t.pos, s.indel_len);
			w_changelast(S.hs, draft, cchar(S.t));
			S.la_n = 0;
			break;
		default:
			// soft-masking only changes the case of the tail char: no effect on the hash (ntedit.cpp:1410-1424)
			if (fl || mask_() || (snv_() && s.altsupp1)) {
				emit(0, fl, draft, s);
			}
			break;
		}
	}

	// One site up to, not including, makeEdit (ntedit.cpp:1808-2116).  Returns SITE_NONE when no attempt is made (nothing to
	// commit), SITE_DONE when S.s / S.touched / S.quiet hold the decision, SITE_PENDING when `allow_indels` is false and the
	// candidate loop reached a tryIndels call (pre-evaluation, first pass).
	NTB_FN uint32_t evaluate_site_core(bool allow_indels)
	{
		NTB_PROF_T0
		NTB_LEADER_BEGIN
		site_begin();
		NTB_LEADER_END
		linearise(P.k + MAX_DELETIONS + 1, true);
		NTB_PROF(8);
		if (!snv_() && S.dnf) {
			return SITE_NONE; // no attempt is possible (ntedit.cpp:1865): the subset counts are not needed
		}
		const bool fast = S.patch_idx == P.k - 1 && P.k + 1 < ROT_STRIDE;
		if (fast) {
			compute_plain(P.k);
			NTB_PROF(9);
			phase_fast_check();
		} else {
			phase_check_and_subs();
		}
		NTB_PROF(10);
		NTB_LEADER_BEGIN
		S.next = site_after_check() ? NEXT_CAND : NEXT_STOP;
		NTB_LEADER_END
		if (S.next == NEXT_STOP) {
			return SITE_NONE;
		}
		if (fast) {
			phase_fast_subs();
			NTB_PROF(10);
		}
		NTB_LEADER_BEGIN
		S.ci = 0;
		NTB_LEADER_END
		for (;;) {
			// the leader walks the substitution candidates until one needs tryIndels (a warp-wide phase)
			NTB_LEADER_BEGIN
			S.next = NEXT_STOP;
			for (; S.ci < 4; S.ci++) {
				site_candidate(S.ci);
				if (S.next != NEXT_CAND) {
					break;
				}
				S.next = NEXT_STOP;
			}
			NTB_LEADER_END
			if (S.next != NEXT_INDELS) {
				break;
			}
			if (!allow_indels && P.max_ins_tries != 0) {
				return SITE_PENDING; // (with nothing to try, tryIndels returns at once and the loop goes on)
			}
			NTB_PROF(11);
#if defined(NTB_PHASE_PROF) && defined(__CUDA_ARCH__)
			if (lane_id() == 0) {
				S.prof[15]++; // tryIndels calls
			}
#endif
			const bool hit = try_indels();
			NTB_PROF(12);
			if (hit && (P.mode == 0 || P.mode == 1)) {
				break;
			}
			NTB_LEADER_BEGIN
			S.ci++;
			NTB_LEADER_END
		}
		NTB_PROF(11);
		NTB_LEADER_BEGIN
		// An accepted substitution leaves k-1 windows that contain the new base; the k-th is clean again.  Phase 1 probed
		// all of them for the accepted candidate: when none is a site (and the k incoming bases are accepted --
		// n_check == k), rolling through them one by one (ntedit.cpp:2118-2138) has no observable effect.
		S.quiet = false;
		if (S.s.best_type == 1 && S.p1_fast && !snv_() && S.lin_simple && S.tail_is_pos && !S.dnf && S.n_check == P.k && S.n_rolls >= P.k) {
			uint32_t c = 0;
			while (c < 4 && ((S.cands >> (8 * c)) & 0xFF) != S.s.best_sub) {
				c++;
			}
			bool quiet = c < 4;
			for (uint32_t R = 1; quiet && R + 1 <= P.k; R++) {
				if (S.tsite[c][R - 1]) {
					quiet = false;
				}
			}
			S.quiet = quiet;
		}
		NTB_LEADER_END
		return SITE_DONE;
	}

	// returns false when the contig is finished (insertion guard fired)
	NTB_FN bool evaluate_site()
	{
		if (evaluate_site_core(true) != SITE_DONE) {
			return true;
		}
		NTB_PROF_T0
		NTB_LEADER_BEGIN
		site_commit((S.touched && S.raw != S.draft) ? EV_TOUCHED : 0);
		NTB_LEADER_END
		NTB_PROF(13);
		return S.site_ok;
	}

	// ---------------------------------------------------------------- pre-evaluation (see ntb_common.h: SiteRec)
	// all lanes: put the walker on the clean window that ends at `pos`, as the main loop's clean branch would (step())
	NTB_FN void pre_seed(uint32_t pos)
	{
		const uint32_t k = P.k, head = pos + 1 - k;
		NTB_LEADER_BEGIN
		S.t.pos = pos;
		S.h.pos = head;
		reset_rope(head);
		S.stale_best_sub = STALE_REF | 0;
		S.stale_alt1 = STALE_REF | 1;
		S.stale_alt2 = STALE_REF | 2;
		S.stale_alt3 = STALE_REF | 3;
		S.la_n = 0;
		NTB_LEADER_END
		if (!cache_covers_window()) {
			fill_cache(head);
		}
		uint64_t seed_f = 0, seed_r = 0;
		const uint64_t* rot = rot_();
		for (uint32_t i = lane_id(); i < k; i += lane_count()) {
			const unsigned char c = text_at(head + i);
			seed_f ^= rot[code_f(c) * ROT_STRIDE + (k - 1 - i)];
			seed_r ^= rot[code_r(c) * ROT_STRIDE + i];
		}
		seed_f = warp_xor64(seed_f);
		seed_r = warp_xor64(seed_r);
		NTB_LEADER_BEGIN
		S.hs.fh = seed_f;
		S.hs.rh = seed_r;
		S.char_in = text_at(pos);
		NTB_LEADER_END
	}

	// leader: the decision of the site just evaluated, as a record
	NTB_FN void pre_fill(SiteRec& r, uint32_t state) const
	{
		const Site& s = S.s;
		r.state = (uint8_t)state;
		r.draft = S.draft;
		r.flags = 0;
		r.pad_[0] = r.pad_[1] = 0;
		r.best_type = 0;
		r.best_sub = 0;
		r.support = 0;
		r.indel_len = 0;
		for (int i = 0; i < 3; i++) {
			r.altsupp[i] = 0;
			r.altbase[i] = 0;
		}
		for (int i = 0; i < 5; i++) {
			r.indel[i] = 0;
		}
		if (state != SITE_DONE) {
			return;
		}
		r.flags = (uint8_t)(((S.touched && S.raw != S.draft) ? SITE_FL_TOUCHED : 0) | (S.quiet ? SITE_FL_QUIET : 0));
		r.best_type = (uint8_t)s.best_type;
		r.best_sub = s.best_sub;
		r.support = (uint16_t)s.best_support;
		r.altsupp[0] = (uint16_t)s.altsupp1;
		r.altsupp[1] = (uint16_t)s.altsupp2;
		r.altsupp[2] = (uint16_t)s.altsupp3;
		r.altbase[0] = s.altbase1;
		r.altbase[1] = s.altbase2;
		r.altbase[2] = s.altbase3;
		r.indel_len = s.indel_len;
		if (s.best_type == 2) {
			for (int i = 0; i < 5; i++) {
				r.indel[i] = i < s.indel_len ? s.indel[i] : 0;
			}
		}
	}

	// Which flagged positions are pre-evaluated.  An error makes (up to) k consecutive windows absent; once the first of
	// them is corrected the main loop never evaluates the others, so a flagged position with another one less than k in
	// front of it is usually not a site at all.  Items therefore start at HEADS -- flagged positions with no flagged
	// position among the pre_gap() = k-1 positions in front of them -- and follow the main loop from there: while a site
	// ends without an edit, the next flagged position is evaluated too if it lies within pre_gap() (a farther one is a head
	// of its own).
	NTB_FN uint32_t pre_gap() const { return P.k - 1; }

	// is `pos` (a flagged tail position of the contig at text offset goff) a head?
	NTB_FN static bool is_head(const uint32_t* visit, uint64_t goff, uint32_t pos, uint32_t gap)
	{
		const uint64_t g1 = goff + pos;
		const uint64_t g0 = goff + (pos > gap ? pos - gap : 0);
		if (g0 >= g1) {
			return true;
		}
		for (uint64_t w = g0 >> 5; (w << 5) < g1; w++) {
			uint32_t bits = visit[w];
			if (w == (g0 >> 5)) {
				bits &= 0xFFFFFFFFu << (g0 & 31);
			}
			if (((w + 1) << 5) > g1) {
				bits &= 0xFFFFFFFFu >> (32 - (g1 & 31));
			}
			if (bits) {
				return false;
			}
		}
		return true;
	}

	// all lanes: the site at `pos` and the chain behind it.  First pass (allow_indels false): `pos` is a head; a site that
	// reaches tryIndels is left to the second pass as SITE_PENDING and ends the chain -- whether the main loop goes on
	// behind it is only known once its indels have been tried.  Second pass (allow_indels true): `pos` is such a pending
	// site; the chain goes on from it with everything evaluated in place.
	NTB_FN void pre_run(uint32_t task_idx, uint32_t pos, bool allow_indels)
	{
		// The record of the chain's first site learns how many of the sites behind it make no edit and emit nothing
		// (SITE_FL_SKIP: the walker that commits it goes straight to the last of them); leader-only bookkeeping.
		const bool skips = !mask_() && !snv_();
		uint32_t sk_slot = NONE32, sk_pos = 0, sk_n = 0, sk_T = SKIP_IDENTITY;
		bool sk_open = false;
		// (first pass: the head and SITE_CHAIN_MAX chain sites, like the dense form's rounds)
		for (uint32_t n = 0; n < SITE_CHAIN_MAX + (allow_indels ? 0u : 1u); n++) {
			pre_seed(pos);
			const uint32_t st = evaluate_site_core(allow_indels);
			NTB_LEADER_BEGIN
			// (the second pass finds the slot its first site was given)
			const uint32_t slot = site_table_insert(S.io.table, S.io.table_mask, S.io.goff + pos + 1);
			S.rec_slot = slot;
			if (slot == NONE32) {
#if defined(__CUDA_ARCH__)
				atomicAdd(&S.io.ctr->n_dropped, 1u);
#else
				S.io.ctr->n_dropped++;
#endif
			} else {
				SiteRec r;
				r.key = S.io.goff + pos + 1;
				pre_fill(r, st);
				if (allow_indels) {
					r.flags |= SITE_FL_SECOND;
				}
				if (st == SITE_PENDING) {
					uint32_t idx;
#if defined(__CUDA_ARCH__)
					idx = atomicAdd(&S.io.ctr->n_pending, 1u);
#else
					idx = S.io.ctr->n_pending++;
#endif
					if (idx < S.io.pending_cap) {
						PendingSite ps;
						ps.task = task_idx;
						ps.pos = pos;
						ps.slot = slot;
						S.io.pending[idx] = ps;
					}
				}
				S.io.table[slot] = r;
				if (skips && n == 0) {
					sk_open = dense_skippable(st, r.best_type, r.flags); // (a record that emits an event keeps its fields as they are)
					sk_slot = slot;
					sk_pos = pos;
				} else if (sk_open && dense_skippable(st, r.best_type, r.flags)) {
					sk_T = dense_skip_compose(sk_T, st, dense_pack_bases(r));
					sk_n++;
					SiteRec first = S.io.table[sk_slot];
					dense_skip_store(first, sk_n, pos - sk_pos, sk_T);
					S.io.table[sk_slot] = first;
				} else {
					sk_open = false;
				}
			}
			// go on behind a site that made no edit
			S.pre_more = slot != NONE32 && (st == SITE_NONE || (st == SITE_DONE && S.s.best_type == 0));
			NTB_LEADER_END
			if (!S.pre_more) {
				break;
			}
			const uint32_t lim = S.io.len - pos - 1 < pre_gap() ? S.io.len : pos + 1 + pre_gap();
			pos = next_visit(pos + 1, lim);
			if (pos == NONE32) {
				break;
			}
		}
	}

	// all lanes: the record of text position g (nullptr when there is none or it is not complete)
	NTB_FN const SiteRec* find_record(uint64_t g) const
	{
		const SiteRec* table = S.io.table;
		if (!table) {
			return nullptr;
		}
		const uint64_t key = g + 1;
		const uint32_t mask = S.io.table_mask;
		const uint32_t h0 = site_hash(key);
		for (uint32_t i0 = 0; i0 < SITE_TABLE_PROBES; i0 += lane_count()) {
			const uint32_t slot = (h0 + i0 + lane_id()) & mask;
			const uint64_t seen = table[slot].key;
#if defined(__CUDA_ARCH__)
			// an inserted key sits in front of the first empty slot of its probe sequence
			const uint32_t hit = __ballot_sync(team_mask(), seen == key) >> team_base();
			const uint32_t empty = __ballot_sync(team_mask(), seen == 0) >> team_base();
			if (hit) {
				const uint32_t src = (uint32_t)__ffs((int)hit) - 1u;
				const SiteRec* r = &table[(h0 + i0 + src) & mask];
				return r->state == SITE_NONE || r->state == SITE_DONE ? r : nullptr;
			}
			if (empty) {
				return nullptr;
			}
#else
			if (seen == key) {
				return table[slot].state == SITE_NONE || table[slot].state == SITE_DONE ? &table[slot] : nullptr;
			}
			if (seen == 0) {
				return nullptr;
			}
#endif
		}
		return nullptr;
	}

	// leader: take the decision of the site at the tail from record r (evaluate_site_core ran ahead of the walk)
	NTB_FN void load_record(const SiteRec& r)
	{
		S.n_sites++;
		if (S.first_touch == NONE32) {
			S.first_touch = S.t.pos;
		}
		S.site_ok = true;
		S.skip_advance = false;
		S.use_rec = r.state;
		S.rec_second = (r.flags & SITE_FL_SECOND) != 0;
		S.rec_hash = r.state == SITE_DONE && (r.best_type >= 2 || (r.best_type == 1 && !(r.flags & SITE_FL_QUIET)));
		const bool skip = (r.flags & SITE_FL_SKIP) && (r.state == SITE_NONE || (r.state == SITE_DONE && r.best_type == 0));
		S.rec_skip_n = skip ? r.indel_len : (uint8_t)0;
		S.rec_skip_dist = (uint16_t)((uint32_t)(uint8_t)r.indel[4] | ((uint32_t)r.pad_[0] << 8));
		S.rec_skip_T = (uint32_t)(uint8_t)r.indel[0] | ((uint32_t)(uint8_t)r.indel[1] << 8) | ((uint32_t)(uint8_t)r.indel[2] << 16) |
		               ((uint32_t)(uint8_t)r.indel[3] << 24);
		if (r.state != SITE_DONE) {
			return;
		}
		const uint8_t stale[4] = { S.stale_best_sub, S.stale_alt1, S.stale_alt2, S.stale_alt3 };
		Site& s = S.s;
		s.best_type = r.best_type;
		s.best_support = r.support;
		s.altsupp1 = r.altsupp[0];
		s.altsupp2 = r.altsupp[1];
		s.altsupp3 = r.altsupp[2];
		// bytes the record left symbolic are the ones this walker's previous site left behind (STALE_REF)
		s.best_sub = (r.best_sub & STALE_REF) ? stale[r.best_sub & 3] : r.best_sub;
		s.altbase1 = (r.altbase[0] & STALE_REF) ? stale[r.altbase[0] & 3] : r.altbase[0];
		s.altbase2 = (r.altbase[1] & STALE_REF) ? stale[r.altbase[1] & 3] : r.altbase[1];
		s.altbase3 = (r.altbase[2] & STALE_REF) ? stale[r.altbase[2] & 3] : r.altbase[2];
		// (a no-edit record that carries skip information keeps it where an edit keeps its indel)
		s.indel_len = skip ? (uint8_t)0 : r.indel_len;
		for (int i = 0; i < 5; i++) {
			s.indel[i] = skip ? (char)0 : r.indel[i];
		}
		S.draft = r.draft;
		S.rec_fl = (r.flags & SITE_FL_TOUCHED) ? EV_TOUCHED : 0;
		S.quiet = (r.flags & SITE_FL_QUIET) != 0;
		S.tail_is_pos = true;
		S.tail_is_chr = false;
	}

	// ---------------------------------------------------------------- clean-window handling
	NTB_FN bool window_clean() const
	{
		if (S.h.ni != S.t.ni || S.h.ni >= S.nn || S.ty[S.h.ni] != 0 || S.t.pos - S.h.pos != P.k - 1) {
			return false;
		}
		for (uint32_t i = 0; i < S.ov_n; i++) {
			if (S.ov_pos[i] >= S.h.pos) {
				return false;
			}
		}
		return true;
	}

	NTB_FN void reset_rope(uint32_t head_pos)
	{
		S.ty[0] = 0;
		S.ch[0] = 0;
		S.sp[0] = head_pos;
		S.ep[0] = S.io.len - 1;
		S.nn = 1;
		S.h.ni = S.t.ni = 0;
		S.ov_n = 0;
	}

	// first position >= from whose visit bit is set, or NONE32 when there is none below `limit`.  Every lane inspects one
	// bitmap word per round; the result is the same in every lane.
	NTB_FN uint32_t next_visit(uint32_t from, uint32_t limit) const
	{
		if (from >= limit) {
			return NONE32;
		}
		const uint64_t g = S.io.goff + from;
		const uint64_t gend = S.io.goff + limit;
		for (uint64_t w0 = g >> 5;; w0 += lane_count()) {
			const uint64_t w = w0 + lane_id();
			uint32_t bits = (w << 5) < gend ? S.io.visit[w] : 0u;
			if (w == (g >> 5)) {
				bits &= 0xFFFFFFFFu << (g & 31);
			}
#if defined(__CUDA_ARCH__)
			const uint32_t any = __ballot_sync(team_mask(), bits != 0) >> team_base();
			if (any) {
				const int src = __ffs((int)any) - 1;
				const uint32_t b = __shfl_sync(team_mask(), bits, (int)team_base() + src);
				const uint64_t hit = ((w0 + (uint64_t)src) << 5) + (uint32_t)(__ffs((int)b) - 1);
				return hit < gend ? (uint32_t)(hit - S.io.goff) : NONE32;
			}
#else
			if (bits) {
				const uint64_t hit = (w << 5) + (uint32_t)__builtin_ctz(bits);
				return hit < gend ? (uint32_t)(hit - S.io.goff) : NONE32;
			}
#endif
			if (((w0 + lane_count()) << 5) >= gend) {
				return NONE32;
			}
		}
	}

	// all lanes: load TEXT_CACHE bytes of contig text starting at `base` (bytes behind the contig's end read as 0)
	NTB_FN void fill_cache(uint32_t base)
	{
		warp_sync();
#if defined(__CUDA_ARCH__)
		// aligned 32-bit loads (the contig starts at an arbitrary byte of the batch buffer; SCAN_HALO bytes in front of the
		// buffer and its zero padding behind make the rounded-out words readable), shifted into place
		const uint8_t* src = S.io.text + base;
		const uint32_t shift = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
		const uint32_t* words = reinterpret_cast<const uint32_t*>(src - shift);
		uint32_t* dst = reinterpret_cast<uint32_t*>(S.tc);
		for (uint32_t j = lane_id(); j < TEXT_CACHE / 4; j += lane_count()) {
			const uint32_t lo = words[j], hi = words[j + 1];
			uint32_t v = __funnelshift_r(lo, hi, 8u * shift);
			const uint64_t pos = (uint64_t)base + 4u * j;
			if (pos + 4 > S.io.len) {
				const uint32_t keep = pos < S.io.len ? (uint32_t)(S.io.len - pos) : 0u; // 0..3 bytes of this word lie inside the contig
				v &= keep ? (0xFFFFFFFFu >> (8u * (4u - keep))) : 0u;
			}
			dst[j] = v;
		}
#else
		for (uint32_t o = lane_id(); o < TEXT_CACHE; o += lane_count()) {
			const uint64_t pos = (uint64_t)base + o;
			S.tc[o] = pos < S.io.len ? S.io.text[pos] : (uint8_t)0;
		}
#endif
		if (lane_id() == 0) {
			S.tc_base = base;
			S.tc_n = TEXT_CACHE;
		}
		warp_sync();
	}

	NTB_FN bool cache_covers_window() const
	{
		return S.h.pos >= S.tc_base && (uint64_t)S.t.pos + P.k + MAX_DELETIONS + 2 <= (uint64_t)S.tc_base + S.tc_n;
	}

	// leader: NTMC64 seeding form over the (unedited) window that ends at `tail`, ntedit.cpp:403-416
	NTB_FN void seed_at(uint32_t tail)
	{
		const uint32_t head = tail + 1 - P.k;
		S.h.pos = head;
		S.t.pos = tail;
		const Walker* self = this;
		{
			// NTMC64 seeding form (ntedit.cpp:403-416) on class bytes
			const uint32_t k = P.k;
			uint64_t f = 0, r = 0;
			for (uint32_t i = 0; i < k; i++) {
				f = srol1(f) ^ S.seed_tab[cls8(self->text_at(head + i)) & 7u];
				r = srol1(r) ^ S.seed_tab[(cls8(self->text_at(head + k - 1 - i)) >> 3) & 7u];
			}
			S.hs.fh = f;
			S.hs.rh = r;
		}
		S.char_in = text_at(tail);
	}

	NTB_FN unsigned char text_at(uint32_t pos) const
	{
		const uint32_t o = pos - S.tc_base;
		if (o < S.tc_n) {
			return S.tc[o];
		}
		return S.io.text[pos];
	}

	// findFirstAcceptedKmer, ntedit.cpp:524-545
	NTB_FN_NOINLINE uint32_t first_accepted_kmer() const
	{
		const uint32_t k = P.k;
		for (uint32_t i = 0; (uint64_t)i + k < S.io.len;) {
			if (is_acc(S.io.text[i])) {
				bool good = true;
				for (uint32_t j = i + 1; j < i + k; j++) {
					if (!is_acc(S.io.text[j])) {
						good = false;
						i = j + 1;
						break;
					}
				}
				if (good) {
					return i;
				}
			} else {
				i++;
			}
		}
		return S.io.len - 1;
	}

	// drop rope nodes that can no longer be reached so that long dirty stretches fit the bounded array
	NTB_FN_NOINLINE void compact()
	{
		uint32_t lo = S.h.ni < S.t.ni ? S.h.ni : S.t.ni;
		// keep the run of inserted characters left of the tail (getPrevInsertion walks it) plus one node
		uint32_t r = S.t.ni;
		while (r > 0 && S.ty[r - 1] == 1) {
			r--;
		}
		if (r > 0) {
			r--;
		}
		if (r < lo) {
			lo = r;
		}
		if (lo == 0) {
			return;
		}
		for (uint32_t i = lo; i < S.nn; i++) {
			move_node(i - lo, i);
		}
		S.nn -= lo;
		S.h.ni -= lo;
		S.t.ni -= lo;
	}

	// last flagged position (K1 visit bit) in [lo, hi) of the contig, NONE32 if there is none
	NTB_FN uint32_t last_flag_in(uint32_t lo, uint32_t hi) const
	{
		if (lo >= hi) {
			return NONE32;
		}
		const uint64_t g0 = S.io.goff + lo, g1 = S.io.goff + hi;
		for (uint64_t w = (g1 - 1) >> 5;; w--) {
			uint32_t bits = S.io.visit[w];
			if (w == ((g1 - 1) >> 5) && (g1 & 31)) {
				bits &= (1u << (g1 & 31)) - 1u;
			}
			if (w == (g0 >> 5)) {
				bits &= 0xFFFFFFFFu << (g0 & 31);
			}
			if (bits) {
				uint32_t top = 31;
				while (!((bits >> top) & 1u)) {
					top--;
				}
				return (uint32_t)((w << 5) + top - S.io.goff);
			}
			if (w == (g0 >> 5)) {
				return NONE32;
			}
		}
	}

	// Segment borders are nominal (multiples of the segment length).  A border in the middle of a run of flagged positions
	// makes the successor start inside its predecessor's dirty stretch, and the stitcher then has to run it again.  Both
	// neighbours therefore move the border -- each on its own, by the same rule on the same bitmap -- to the first position
	// p >= b (at most boundary_lim further) with no flagged position in [p - W, p), W = 2k + 16: the predecessor is then
	// clean again before p.  (Only a hint: the stitcher still validates every result.)
	NTB_FN uint32_t safe_boundary(uint32_t b) const
	{
		const uint32_t W = 2 * P.k + 16;
		uint32_t p = b;
		while (p <= b + P.boundary_lim && p < S.io.len) {
			const uint32_t q = last_flag_in(p >= W ? p - W : 0, p);
			if (q == NONE32) {
				return p;
			}
			p = q + W + 1;
		}
		return b;
	}

	// leader: the per-walker seed tables
	NTB_FN void init_tables()
	{
		for (unsigned c = 0; c < 8; c++) {
			S.seed_tab[c] = c < 4 ? seed_of_code(c) : 0;
			S.rotk_tab[c] = c < 4 ? P.seed_rot_k[c] : 0;
		}
	}

	// every lane: get ready for pre_run() on the contig S.io describes
	NTB_FN void pre_begin()
	{
		NTB_LEADER_BEGIN
		init_tables();
		S.nn = 0;
		S.ov_n = 0;
		S.adv = 0;
		S.anchored = true;
		S.last_event = NONE32;
		S.n_events = S.n_sites = 0;
		S.first_touch = NONE32;
		S.status = 0;
		S.tc_base = 0;
		S.tc_n = 0;
		S.la_n = S.la_used = S.la_bits = 0;
		S.use_rec = 0;
		S.quiet = false;
		S.need_seed = false;
		S.h.ni = S.t.ni = 0;
		NTB_LEADER_END
	}

	// ---------------------------------------------------------------- the main loop, ntedit.cpp:1797-2139
	// leader: start of a task
	NTB_FN void task_begin(const Task& task)
	{
		const uint32_t k = P.k;
		S.nn = 0;
		S.ov_n = 0;
		if (task.flags & TASK_CONTIG_START) {
			S.stale_best_sub = S.stale_alt1 = S.stale_alt2 = S.stale_alt3 = 0;
		} else {
			S.stale_best_sub = STALE_REF | 0;
			S.stale_alt1 = STALE_REF | 1;
			S.stale_alt2 = STALE_REF | 2;
			S.stale_alt3 = STALE_REF | 3;
		}
		S.adv = 0;
		S.anchored = true;
		S.last_event = NONE32;
		S.n_events = S.n_sites = 0;
		S.first_touch = NONE32;
		S.status = 0;
		S.char_in = 0;
		S.hs.fh = S.hs.rh = 0;
		S.end_pos = S.io.len;
		S.need_seed = true;
		S.tc_base = 0;
		S.tc_n = 0;
		S.la_n = 0;
		S.la_used = 0;
		S.la_bits = 0;
		S.act = ACT_CLEAN;
		S.use_rec = 0;
		S.quiet = false;
		S.h.ni = S.t.ni = 0;
		S.t_start = (P.boundary_lim && (task.flags & TASK_ADJUST_START)) ? safe_boundary(task.start) : task.start;
		S.t_end = (P.boundary_lim && (task.flags & TASK_ADJUST_END)) ? safe_boundary(task.end) : task.end;
		init_tables();
		if (task.flags & TASK_CONTIG_START) {
			const uint32_t h0 = first_accepted_kmer();
			if ((uint64_t)h0 + k - 1 >= S.io.len) {
				S.status |= ST_CONTIG_END;
				S.act = ACT_STOP;
				S.h.pos = S.t.pos = 0;
			} else {
				S.h.pos = h0;
				S.t.pos = h0 + k - 1;
			}
		} else {
			S.t.pos = S.t_start;
			S.h.pos = S.t_start + 1 - k;
		}
		if (S.act != ACT_STOP) {
			reset_rope(S.h.pos);
		}
	}

	// leader: top of the main loop
	NTB_FN void loop_head(const Task& task)
	{
		if ((uint64_t)S.h.pos + P.k - 1 >= S.io.len) {
			S.status |= ST_CONTIG_END;
			S.act = ACT_STOP;
			return;
		}
		if (S.status & (ST_EV_OVERFLOW | ST_ROPE_OVERFLOW)) {
			S.act = ACT_STOP;
			return;
		}
		if (S.need_seed || window_clean()) {
			// clean window: forget the local rope and jump to the next position K1 flagged
			reset_rope(S.h.pos);
			S.anchored = true;
			S.la_n = 0;
			if (S.t.pos >= S.t_end) {
				S.end_pos = S.t.pos;
				S.act = ACT_STOP;
				return;
			}
			S.act = ACT_CLEAN;
			return;
		}
		if (S.nn + 16 > (uint32_t)NCAP) {
			compact();
			if (S.nn + 16 > (uint32_t)NCAP) {
				S.status |= ST_ROPE_OVERFLOW;
				S.act = ACT_STOP;
				return;
			}
		}
		S.use_rec = 0;
		S.act = ACT_DIRTY;
	}

	// leader: move to the next position after the site test (ntedit.cpp:2118-2138)
	// leader: the record just committed made no edit, and says the chain's next rec_skip_n flagged positions make none and
	// emit nothing either (SITE_FL_SKIP): go straight to the last of them -- unless it lies behind this task's end
	NTB_FN void skip_chain()
	{
		const uint32_t last = S.t.pos + S.rec_skip_dist;
		if (last >= S.t_end || !window_clean()) {
			return;
		}
		S.h.pos += S.rec_skip_dist;
		S.t.pos = last;
		S.n_sites += S.rec_skip_n;
#if defined(NTB_PHASE_PROF) && defined(__CUDA_ARCH__)
		atomicAdd(&S.io.ctr->n_skipped, (uint32_t)S.rec_skip_n);
#elif !defined(__CUDA_ARCH__)
		S.io.ctr->n_skipped += S.rec_skip_n;
#endif
		const uint8_t cur[4] = { S.stale_best_sub, S.stale_alt1, S.stale_alt2, S.stale_alt3 };
		uint8_t nw[4];
		for (int j = 0; j < 4; j++) {
			const uint8_t v = (uint8_t)(S.rec_skip_T >> (8 * j));
			nw[j] = (v & STALE_REF) ? cur[v & 3] : v;
		}
		S.stale_best_sub = nw[0];
		S.stale_alt1 = nw[1];
		S.stale_alt2 = nw[2];
		S.stale_alt3 = nw[3];
	}

	NTB_FN void advance()
	{
		const uint32_t k = P.k;
		if (window_clean()) {
			// still on unedited text: the next position the reference acts on is the next flagged one
			// (its own roll / skip-after-N loop, ntedit.cpp:2118-2138, does nothing observable in between)
			S.h.pos++;
			S.t.pos++;
			S.need_seed = true;
			return;
		}
		// advance; after a non-accepted incoming base skip until k further positions were consumed
		int64_t target = -1;
		do {
			unsigned char out = 0, in = S.char_in;
			if (roll(S.h, S.t, out, in)) {
				S.char_in = in;
				S.adv++;
				S.la_used++;
				if (!is_acc(in)) {
					target = (int64_t)(int32_t)S.t.pos + (int64_t)(int32_t)k;
					S.la_n = 0;
				}
				roll_cls(S.hs, cls8(out), cls8(in));
			} else {
				S.status |= ST_CONTIG_END;
				S.act = ACT_STOP;
				return;
			}
			if (target >= 0 && (int64_t)(int32_t)S.t.pos != target && window_clean()) {
				// skipping over non-accepted bases on unedited text: same shortcut as above
				S.h.pos++;
				S.t.pos++;
				S.need_seed = true;
				break;
			}
		} while (target >= 0 && (int64_t)(int32_t)S.t.pos != target);
	}

	// all lanes: site test of the next LOOKAHEAD dirty-window positions in one pass (lane j: j plain rolls ahead)
	NTB_FN void lookahead()
	{
		linearise(LOOKAHEAD - 1, false);
		const uint32_t n = S.n_rolls + 1 < LOOKAHEAD ? S.n_rolls + 1 : LOOKAHEAD;
		compute_plain(LOOKAHEAD - 1);
		uint32_t bits = 0;
		{
			// lane j: the k-mer after j plain rolls (LOOKAHEAD <= lanes * PROBE_G)
			const uint32_t ln = lane_id();
			uint32_t ng = 0;
			for (uint32_t j = ln; j < n && ng < (uint32_t)PROBE_G; j += lane_count()) {
				S.hb[ng][ln] = S.plain_f[j] + S.plain_r[j];
				ng++;
			}
			for (uint32_t g0 = 0; g0 < ng || (lane_count() == 1 && g0 < n); g0 += PROBE_G) {
				if (lane_count() == 1 && g0 > 0) {
					// one-lane build: refill the group
					ng = 0;
					for (uint32_t j = g0; j < n && ng < (uint32_t)PROBE_G; j++) {
						S.hb[ng][ln] = S.plain_f[j] + S.plain_r[j];
						ng++;
					}
				}
				if (ng == 0) {
					break;
				}
				probe_values(S.io.bloom, ng, (1u << ng) - 1u, PROBE_HU);
				for (uint32_t g = 0; g < ng; g++) {
					const uint32_t j = lane_count() == 1 ? g0 + g : ln + g * lane_count();
					if (is_site_value(S.pval[g][ln])) {
						bits |= 1u << j;
					}
				}
				if (lane_count() > 1) {
					break;
				}
			}
		}
		bits = warp_or(bits); // lane j holds bit j only
		// Can the whole dirty run be skipped?  With only in-place substitutions in the window, it is clean again after
		// J = (last substituted position - head + 1) rolls.  When none of the J dirty windows is a site and every incoming
		// base is accepted, rolling through them one by one (ntedit.cpp:2118-2138) has no observable effect: jump.
		NTB_LEADER_BEGIN
		uint32_t J = 0;
		if (S.lin_simple) {
			for (uint32_t i = 0; i < S.ov_n; i++) {
				if (S.ov_pos[i] >= S.h.pos && S.ov_pos[i] - S.h.pos + 1 > J) {
					J = S.ov_pos[i] - S.h.pos + 1;
				}
			}
		}
		S.la_J = (J >= 1 && J <= n && J <= S.n_rolls && J <= 32) ? J : 0;
		NTB_LEADER_END
		const uint32_t J = S.la_J;
		uint32_t blocked = J == 0 ? 1u : 0u;
		for (uint32_t m = lane_id(); m < J; m += lane_count()) {
			if (!is_acc(S.lin_in[m])) {
				blocked = 1;
			}
		}
		blocked = warp_or(blocked);
		// Otherwise: the windows in front of the first site of the pass are rolled through in one step, as far as the incoming
		// bases are accepted (a non-accepted one starts the skip logic of ntedit.cpp:2119-2138, which advance() replays) --
		// the same rolls the main loop would make one by one with nothing to observe in between.
		uint32_t first_bad = n;
		for (uint32_t m = lane_id(); m + 1 < n; m += lane_count()) {
			if (!is_acc(S.lin_in[m]) && m < first_bad) {
				first_bad = m;
			}
		}
		first_bad = warp_min(first_bad);
		NTB_LEADER_BEGIN
		S.la_bits = bits;
		S.la_n = n;
		S.la_used = 0;
		S.jumped = false;
		if (!blocked && (bits & (J >= 32 ? 0xFFFFFFFFu : ((1u << J) - 1u))) == 0) {
			S.h.pos += J;
			S.t.pos += J;
			S.adv += J;
			S.need_seed = true; // the window is clean: the hash is re-seeded at the next flagged position
			S.la_n = 0;
			S.jumped = true;
		} else {
			uint32_t first_site = 0;
			while (first_site < n && !((bits >> first_site) & 1u)) {
				first_site++;
			}
			uint32_t G = first_site < n ? first_site : n - 1;   // land on the first site, or on the last window of the pass
			if (G > first_bad) {
				G = first_bad;
			}
			if (G > S.n_plain) {
				G = S.n_plain;
			}
			if (G > 0) {
				for (uint32_t q = 0; q < G; q++) {
					step(S.h);
					step(S.t);
				}
				S.hs.fh = S.plain_f[G];
				S.hs.rh = S.plain_r[G];
				S.char_in = S.lin_in[G - 1];
				S.adv += G;
				S.la_used = G;
				S.jumped = true; // back to the top of the main loop: the window may be clean again by now
			}
		}
		NTB_LEADER_END
	}

	// One iteration of the main loop (ntedit.cpp:1797-2139).  Called by every lane of the warp; returns false when the
	// task is finished.  The CUDA kernel aligns the warps of a CTA at the top of every iteration (they then run the same
	// instructions at about the same time, which is what keeps the instruction cache effective); the iteration itself has
	// no CTA-level synchronisation.
	NTB_FN bool step(const Task& task)
	{
		if (S.act == ACT_STOP) {
			return false;
		}
		NTB_PROF_T0
		NTB_LEADER_BEGIN
		loop_head(task);
		NTB_LEADER_END
		NTB_PROF(0);
		if (S.act == ACT_STOP) {
			return false;
		}
		if (S.act == ACT_CLEAN) {
			const uint32_t nv = next_visit(S.t.pos, S.t_end);
			NTB_PROF(1);
			NTB_LEADER_BEGIN
			if (nv == NONE32) {
				S.end_pos = S.t_end;
				S.act = ACT_STOP;
			} else {
				S.do_seed = nv != S.t.pos || S.need_seed;
				S.visit_hit = nv;
				S.use_rec = 0;
				if (S.do_seed) {
					S.t.pos = nv;
					S.h.pos = nv + 1 - P.k;
				}
				S.need_seed = false;
			}
			NTB_LEADER_END
			if (S.act == ACT_STOP) {
				return false;
			}
			// decided ahead of the walk (pre-evaluation pass)?  Most such sites commit without the window's hash or its text
			const SiteRec* rec = find_record(S.io.goff + S.visit_hit);
			if (rec) {
				NTB_LEADER_BEGIN
				load_record(*rec);
				NTB_LEADER_END
			}
			const bool need_hash = !S.use_rec || S.rec_hash;
			if (need_hash && !cache_covers_window()) {
				fill_cache(S.h.pos);
			}
			NTB_PROF(2);
			uint64_t seed_f = 0, seed_r = 0;
			if (S.do_seed && need_hash) {
				// NTMC64 seeding form (ntedit.cpp:403-416) of the unedited window that ends at the flagged position:
				// every lane contributes its bases' terms from the rotation table
				const uint32_t k = P.k, head = S.visit_hit + 1 - k;
				const uint64_t* rot = rot_();
				for (uint32_t i = lane_id(); i < k; i += lane_count()) {
					const unsigned char c = text_at(head + i);
					seed_f ^= rot[code_f(c) * ROT_STRIDE + (k - 1 - i)];
					seed_r ^= rot[code_r(c) * ROT_STRIDE + i];
				}
				seed_f = warp_xor64(seed_f);
				seed_r = warp_xor64(seed_r);
			}
			NTB_LEADER_BEGIN
			if (S.do_seed) {
				S.h.pos = S.visit_hit + 1 - P.k;
				S.t.pos = S.visit_hit;
				if (need_hash) {
					S.hs.fh = seed_f;
					S.hs.rh = seed_r;
					S.char_in = text_at(S.visit_hit);
				}
				reset_rope(S.h.pos);
				S.site_now = true; // K1 flagged this very window
			} else if (S.use_rec) {
				S.site_now = true;
			} else if (snv_()) {
				S.site_now = true;
			} else if (counting_()) {
				S.site_now = is_site_value(q_count(S.hs));
			} else {
				S.site_now = !q_contains(S.hs);
			}
			NTB_LEADER_END
			NTB_PROF(3);
		} else {
			if (!cache_covers_window()) {
				fill_cache(S.h.pos);
			}
			if (snv_()) {
				NTB_LEADER_BEGIN
				S.site_now = true;
				NTB_LEADER_END
			} else {
				if (S.la_used >= S.la_n) {
					lookahead();
					NTB_PROF(4);
					if (S.jumped) {
						return true; // the window moved (possibly onto clean text): back to the top of the main loop
					}
				}
				NTB_LEADER_BEGIN
				S.site_now = ((S.la_bits >> S.la_used) & 1u) != 0;
				NTB_LEADER_END
			}
		}
		NTB_PROF(5);
		if (S.site_now) {
			bool ok;
			if (S.use_rec) {
				NTB_LEADER_BEGIN
				if (S.use_rec == SITE_DONE) {
					site_commit(S.rec_fl);
				}
#if defined(NTB_PHASE_PROF) && defined(__CUDA_ARCH__)
				atomicAdd(&S.io.ctr->n_rec_used, 1u);
#elif !defined(__CUDA_ARCH__)
				S.io.ctr->n_rec_used++;
				S.io.ctr->n_rec_used2 += S.rec_second ? 1u : 0u;
#endif
				NTB_LEADER_END
				ok = S.site_ok;
			} else {
				ok = evaluate_site();
			}
			if (!ok) {
				NTB_LEADER_BEGIN
				S.status |= ST_CONTIG_END;
				S.act = ACT_STOP;
				NTB_LEADER_END
				return false;
			}
		}
		NTB_PROF(6);
		NTB_LEADER_BEGIN
		if (S.site_now && S.skip_advance) {
			S.skip_advance = false;
		} else {
			if (S.site_now && S.use_rec && S.rec_skip_n) {
				skip_chain();
			}
			advance();
		}
		S.rec_skip_n = 0;
		NTB_LEADER_END
		NTB_PROF(7);
		return S.act != ACT_STOP;
	}

	// every lane: start a task
	NTB_FN void begin(const Task& task)
	{
		NTB_LEADER_BEGIN
		task_begin(task);
		NTB_LEADER_END
	}

	// every lane: the result of a finished task (valid in the leader lane)
	NTB_FN void finish(TaskResult& res)
	{
		NTB_LEADER_BEGIN
		S.status |= ST_DONE;
		res.end_pos = (S.status & ST_CONTIG_END) ? S.io.len : S.end_pos;
		res.first_touch = S.first_touch;
		res.last_event = S.last_event;
		res.n_events = S.n_events;
		res.n_sites = S.n_sites;
		res.status = S.status;
		res.stale[0] = S.stale_best_sub;
		res.stale[1] = S.stale_alt1;
		res.stale[2] = S.stale_alt2;
		res.stale[3] = S.stale_alt3;
		res.kcycles = 0;
		res.start_pos = S.t_start;
		NTB_LEADER_END
	}

	// Walks one task.  Called by every lane of the warp with the same arguments.
	NTB_FN void run(const Task& task, TaskResult& res)
	{
		begin(task);
		while (step(task)) {
		}
		finish(res);
	}
#undef S
#undef P
};

} // namespace ntb
