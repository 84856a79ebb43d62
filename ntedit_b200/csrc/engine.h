// Device engine ("walker"): one thread replays ntEdit's per-contig state machine (kmerizeAndCorrect and callees,
// ntedit.cpp:1216-2151) over one segment of a contig, starting from a clean window, and reports the edit
// decisions as makeEdit-level events.  The authoritative rope is rebuilt on the host from those events
// (replay.hpp); the walker keeps only a bounded local copy of the rope tail -- enough to answer every
// getCharacter/increment/roll the reference would make while the window is "dirty" (overlaps an edit).
//
// While the window is clean (k consecutive, unedited draft bases) the walker does not roll base by base:
// it jumps to the next position flagged by the scan kernel's visit bitmap (K1) and re-seeds the hash there,
// which is exactly where the reference's main loop would next find `!bloom.contains(hVal)` (ntedit.cpp:1806).
//
// The file compiles for the device (product) and, for the test-only host simulator under tests/hostsim,
// for the CPU; the product library never instantiates the host version.
#pragma once
#include "nthash.h"

namespace ntb {

#if defined(__CUDACC__)
#define NTB_FN __host__ __device__
#else
#define NTB_FN
#endif

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
};

template<int NCAP>
struct Walker
{
	static constexpr int OVCAP = KMAX + 8;
	static constexpr int PREVCAP = 2 * KMAX + 16;

	const WalkerIO& io;
	const KParams& P;

	// bounded copy of the rope tail (seqNode vector, ntedit.cpp:613-620)
	int8_t ty[NCAP];
	uint8_t ch[NCAP];
	uint32_t sp[NCAP], ep[NCAP];
	uint32_t nn;

	// substitutions applied in place to position nodes that the head may still read (contigSeq mutation, ntedit.cpp:1283)
	uint32_t ov_pos[OVCAP];
	uint8_t ov_ch[OVCAP];
	uint32_t ov_n;
	// temporary substitution of a trial (ntedit.cpp:1936-1940)
	uint32_t patch_pos;
	uint8_t patch_ch;
	bool patch_on;

	Cursor h, t;
	HashState hs;
	unsigned char char_in;

	// the reference declares these without initialisers inside the loop body (ntedit.cpp:1881-1885); the compiled
	// reference keeps them in fixed slots, so stale values leak from site to site (observable in mode 2)
	uint8_t stale_best_sub, stale_alt1, stale_alt2, stale_alt3;

	uint32_t adv;      // tail increments since the last event
	bool anchored;
	uint32_t last_event, n_events, n_sites, first_touch, status;

	NTB_FN Walker(const WalkerIO& io_, const KParams& p_) : io(io_), P(p_) {}

	// ---------------------------------------------------------------- text / rope access
	NTB_FN unsigned char rd(uint32_t pos) const
	{
		if (patch_on && pos == patch_pos) {
			return patch_ch;
		}
		for (uint32_t i = 0; i < ov_n; i++) {
			if (ov_pos[i] == pos) {
				return ov_ch[i];
			}
		}
		return pos < io.len ? io.text[pos] : 0;
	}

	// getCharacter, ntedit.cpp:812-823
	NTB_FN unsigned char cchar(const Cursor& c) const
	{
		if (c.ni >= nn) {
			return 0;
		}
		if (ty[c.ni] == 0) {
			return rd(c.pos);
		}
		if (ty[c.ni] == 1) {
			return ch[c.ni];
		}
		return 0;
	}

	// increment, ntedit.cpp:826-844
	NTB_FN void step(Cursor& c) const
	{
		if (c.ni >= nn) {
			return;
		}
		const int8_t tp = ty[c.ni];
		if (tp == 0) {
			c.pos++;
			if (c.pos > ep[c.ni]) {
				c.ni++;
				if (c.ni < nn && ty[c.ni] == 0) {
					c.pos = sp[c.ni];
				}
			}
		} else if (tp == 1) {
			c.ni++;
			if (c.ni < nn && ty[c.ni] == 0) {
				c.pos = sp[c.ni];
			}
		}
	}

	// roll, ntedit.cpp:1216-1247
	NTB_FN bool roll(Cursor& hh, Cursor& tt, unsigned char& out, unsigned char& in) const
	{
		if (hh.pos >= io.len || hh.ni >= nn) {
			return false;
		}
		out = cchar(hh);
		step(hh);
		if (tt.pos >= io.len || tt.ni >= nn) {
			return false;
		}
		step(tt);
		if (tt.pos >= io.len || tt.ni >= nn) {
			return false;
		}
		in = cchar(tt);
		return true;
	}

	NTB_FN void put(uint32_t i, int8_t type, uint8_t c, uint32_t s, uint32_t e)
	{
		if (i >= (uint32_t)NCAP) {
			status |= ST_ROPE_OVERFLOW;
			return;
		}
		ty[i] = type;
		ch[i] = c;
		sp[i] = s;
		ep[i] = e;
		if (i >= nn) {
			nn = i + 1;
		}
	}

	NTB_FN void move_node(uint32_t dst, uint32_t src)
	{
		ty[dst] = ty[src];
		ch[dst] = ch[src];
		sp[dst] = sp[src];
		ep[dst] = ep[src];
	}

	// makeInsertion, ntedit.cpp:625-714
	NTB_FN void rope_insert(uint32_t& t_ni, uint32_t insert_pos, const char* bases, uint32_t nb)
	{
		const int8_t otype = ty[t_ni];
		const uint32_t os = sp[t_ni], oe = ep[t_ni];
		if (otype == 0 && insert_pos > os) {
			ep[t_ni] = insert_pos - 1;
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
			while (t_ni + nlift < nn && ty[t_ni + nlift] != -1) {
				nlift++;
			}
			if (t_ni + nb + nlift > (uint32_t)NCAP) {
				status |= ST_ROPE_OVERFLOW;
				return;
			}
			if (t_ni + nb + nlift > nn) {
				nn = t_ni + nb + nlift;
			}
			for (uint32_t q = nlift; q > 0; q--) {
				move_node(t_ni + nb + q - 1, t_ni + q - 1);
			}
			// slots between the lifted run's old end and its new start that were not overwritten stay as the
			// reference leaves them: the old entries were marked dead before being re-appended
			for (uint32_t q = 0; q < nb; q++) {
				ty[t_ni + q] = 1;
				ch[t_ni + q] = (uint8_t)bases[q];
				sp[t_ni + q] = 0;
				ep[t_ni + q] = 0;
			}
		}
	}

	// makeDeletion, ntedit.cpp:719-809 (the recursion of the reference is a loop here)
	NTB_FN void rope_delete(uint32_t& t_ni, uint32_t& pos, uint32_t num_del)
	{
		for (;;) {
			const int8_t otype = ty[t_ni];
			const uint32_t os = sp[t_ni], oe = ep[t_ni];
			uint32_t leftover = 0;
			if (otype == 0) {
				if (pos <= os) {
					if (pos + num_del <= oe) {
						sp[t_ni] = pos + num_del;
						pos = sp[t_ni];
						return;
					}
					leftover = pos + num_del - oe;
					pos = oe + 1;
					uint32_t i = t_ni + 1;
					while (i < nn && ty[i] != -1) {
						move_node(i - 1, i);
						ty[i] = -1;
						i++;
					}
				} else {
					if (pos + num_del <= oe) {
						ep[t_ni] = pos - 1;
						const uint32_t ns = pos + num_del;
						pos = ns;
						t_ni++;
						put(t_ni, 0, 0, ns, oe);
						return;
					}
					leftover = pos + num_del - oe;
					ep[t_ni] = pos - 1;
					pos = oe + 1;
					t_ni++;
				}
			} else if (otype == 1) {
				uint32_t i = t_ni;
				leftover = num_del;
				while (i < nn && ty[i] == 1 && leftover > 0) {
					ty[i] = -1;
					leftover--;
					i++;
				}
				uint32_t j = t_ni;
				while (i < nn && ty[i] != -1) {
					move_node(j, i);
					ty[i] = -1;
					i++;
					j++;
				}
			} else {
				return;
			}
			if (leftover > 0 && t_ni < nn && ty[t_ni] != -1) {
				if (ty[t_ni] == 0) {
					pos = sp[t_ni];
				}
				num_del = leftover;
				continue;
			}
			return;
		}
	}

	// ---------------------------------------------------------------- filter queries
	NTB_FN unsigned q_count(const HashState& s) const { return filter_count(io.bloom, hash_canonical(s), P.k); }

	NTB_FN bool q_contains(const HashState& s) const { return filter_contains(io.bloom, hash_canonical(s), P.k); }

	// bloom.contains(hVal) && is_kmer_solid(hVal, bloom, bloomrep), ntedit.cpp:465-473
	NTB_FN bool q_present_solid(const HashState& s) const
	{
		const uint64_t b = hash_canonical(s);
		if (P.counting) {
			const unsigned c = filter_count(io.bloom, b, P.k);
			if (c == 0 || c < P.min_threshold || c > P.max_threshold) {
				return false;
			}
		} else if (!filter_contains(io.bloom, b, P.k)) {
			return false;
		}
		if (P.h_rep && filter_contains(io.rep, b, P.k)) {
			return false;
		}
		return true;
	}

	NTB_FN bool meets_edit(uint32_t c) const { return c >= P.thr_edit; }

	// ---------------------------------------------------------------- events
	NTB_FN void emit(uint8_t kind, uint8_t flags, uint8_t draft, const Site& s)
	{
		uint32_t idx;
#if defined(__CUDA_ARCH__)
		idx = atomicAdd(&io.ctr->n_events, 1u);
#else
		idx = io.ctr->n_events++;
#endif
		if (idx >= io.ev_cap) {
			status |= ST_EV_OVERFLOW;
			io.ctr->overflow = 1;
			return;
		}
		Event e;
		e.prev = last_event;
		e.t_pos = t.pos;
		e.advance = anchored ? NONE32 : adv;
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
		io.events[idx] = e;
		last_event = idx;
		n_events++;
		adv = 0;
		anchored = false;
	}

	// ---------------------------------------------------------------- pieces of makeEdit that need the rope
	// getPrevInsertion, ntedit.cpp:907-922: reverse-complemented run of inserted characters left of the tail
	NTB_FN uint32_t prev_insertion(char* out) const
	{
		uint32_t ni = t.ni, n = 0;
		if ((ni < nn && ty[ni] == 0 && t.pos == sp[ni]) || (ni < nn && ty[ni] == 1)) {
			ni--;
		}
		while (ni < nn && ty[ni] == 1 && n < (uint32_t)PREVCAP - 8) {
			const unsigned char c = ch[ni];
			const unsigned cc = base_code(c);
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
	NTB_FN bool insertion_guard_fires(const Site& s) const
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
			const unsigned cc = base_code((unsigned char)s.indel[w]);
			prev[0] = cc == 0 ? 'T' : cc == 1 ? 'G' : cc == 2 ? 'C' : (cc == 3 && (s.indel[w] | 0x20) == 't') ? 'A' : 'N';
			np++;
			if (is_repeat(prev, (int)np)) {
				return true;
			}
		}
		return false;
	}

	// ---------------------------------------------------------------- candidate trials
	// tryDeletion, ntedit.cpp:1451-1545
	NTB_FN uint32_t try_deletion(unsigned char draft, uint32_t num_del) const
	{
		HashState s = hs;
		Cursor hh = h, tt = t;
		unsigned char out = 0, in = 0;
		for (uint32_t i = 0; i < num_del; i++) {
			step(tt);
		}
		hash_changelast(s, draft, cchar(tt), P);
		uint32_t present = q_present_solid(s) ? 1u : 0u;
		for (uint32_t q = 1; q + 2 <= P.k && hh.pos < io.len; q++) {
			if (roll(hh, tt, out, in)) {
				hash_roll(s, out, in, P);
				if (q % P.jump == 0 && q_present_solid(s)) {
					present++;
				}
			}
		}
		return present >= P.thr_edit_del ? present : 0u;
	}

	// i-th string of ntedit.cpp:203-348 for first base `first`: all words of length 1..5 over ACGT that start with
	// `first`, ordered by (length, lexicographic A<C<G<T)
	NTB_FN static uint32_t indel_string(unsigned char first, uint32_t q, char* out)
	{
		uint32_t len = 1, start = 0, count = 1;
		while (q >= start + count) {
			start += count;
			count <<= 2;
			len++;
		}
		uint32_t r = q - start;
		out[0] = (char)first;
		for (uint32_t i = len - 1; i >= 1; i--) {
			const uint32_t d = r & 3;
			out[i] = d == 0 ? 'A' : d == 1 ? 'C' : d == 2 ? 'G' : 'T';
			r >>= 2;
		}
		return len;
	}

	// tryIndels, ntedit.cpp:1548-1744
	NTB_FN bool try_indels(unsigned char draft, unsigned char index_char, uint32_t& num_deletions, Site& site) const
	{
		uint32_t tb_support = 0, ta_support = 0, tb_type = 0, tb_len = 0;
		char tb_indel[11];
		unsigned char out = 0, in = 0;
		for (uint32_t i = 0; i < P.max_ins_tries; i++) {
			char ins[8];
			uint32_t nins = indel_string(index_char, i, ins);
			ins[nins++] = (char)draft;
			HashState s = hs;
			Cursor hh = h, tt = t;
			hash_changelast(s, draft, index_char, P);
			uint32_t present = 0, q = 0;
			for (; q + 1 < nins && hh.pos < io.len; q++) {
				hash_roll(s, cchar(hh), (unsigned char)ins[q + 1], P);
				step(hh);
				if (q % P.jump == 0 && q_present_solid(s)) {
					present++;
				}
			}
			for (; q + 1 < P.k && hh.pos < io.len; q++) {
				if (roll(hh, tt, out, in)) {
					hash_roll(s, out, in, P);
					if (q % P.jump == 0 && q_present_solid(s)) {
						present++;
					}
				}
			}
			nins--;
			if (meets_edit(present)) {
				if (P.mode == 0) {
					site.best_type = 2;
					for (uint32_t c = 0; c < nins; c++) {
						site.indel[c] = ins[c];
					}
					site.indel_len = (uint8_t)nins;
					site.best_support = present;
					return true;
				}
				if (present >= tb_support) {
					if (tb_support) {
						ta_support = tb_support;
					}
					tb_type = 2;
					for (uint32_t c = 0; c < nins; c++) {
						tb_indel[c] = ins[c];
					}
					tb_len = nins;
					tb_support = present;
				}
			}
			if (num_deletions <= P.max_deletions) {
				const uint32_t del_support = try_deletion(draft, num_deletions);
				if (del_support > 0) {
					if (P.mode == 0) {
						site.best_type = 3;
						site.indel_len = (uint8_t)num_deletions;
						site.best_support = del_support;
						return true;
					}
					if (del_support >= tb_support) {
						if (tb_support) {
							ta_support = tb_support;
						}
						tb_type = 3;
						tb_len = num_deletions;
						tb_support = del_support;
					}
				}
				num_deletions++;
			}
		}
		if (tb_support > 0) {
			if ((P.mode == 2 && tb_support > site.best_support) || P.mode == 1) {
				site.best_type = tb_type;
				site.indel_len = (uint8_t)tb_len;
				if (tb_type == 2) {
					for (uint32_t c = 0; c < tb_len; c++) {
						site.indel[c] = tb_indel[c];
					}
				}
				site.best_support = tb_support;
				site.altsupp1 = ta_support;
			}
			return true;
		}
		return false;
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
		if (P.snv) {
			return is_accepted_any_case(draft) || draft == 'N' ? NTB_PACK('A', 'T', 'C', 'G') : 0u;
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

	// ---------------------------------------------------------------- one site: ntedit.cpp:1808-2116
	// returns false when the contig is finished (insertion guard fired)
	NTB_FN bool evaluate_site()
	{
		const uint32_t k = P.k;
		const unsigned char raw = char_in;
		const unsigned char draft = to_upper(raw);
		n_sites++;
		if (first_touch == NONE32) {
			first_touch = t.pos;
		}

		// confirm the k-mer is missing on a subset of the next k windows, ntedit.cpp:1819-1864
		HashState ts = hs;
		Cursor th = h, tt = t;
		unsigned char out = 0, in = 0;
		uint32_t missing = 0, there = 0, nmed = 0;
		uint8_t med[KMAX];
		bool do_not_fix = false;
		for (uint32_t q = 0; q < k && th.pos < io.len; q++) {
			if (!roll(th, tt, out, in)) {
				do_not_fix = true;
				break;
			}
			hash_roll(ts, out, in, P);
			if (!is_accepted_any_case(in)) {
				do_not_fix = true;
				break;
			}
			if (q % P.jump == 0) {
				if (P.counting) {
					const unsigned c = q_count(ts);
					if (c == 0) {
						missing++;
					} else if (is_atgc_upper(draft) && c >= P.min_threshold) {
						there++;
						if (nmed < KMAX) {
							med[nmed++] = (uint8_t)c;
						}
					}
				} else if (!q_contains(ts)) {
					missing++;
				} else if (is_atgc_upper(draft)) {
					there++;
				}
			}
		}
		uint32_t there_median = 0;
		if (P.counting && nmed > 0) {
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
		const bool attempt =
		    P.snv || (!do_not_fix && (missing >= P.thr_missing || (P.counting && there_median < P.min_threshold)));
		if (!attempt) {
			return true;
		}

		Site s;
		s.best_type = 0;
		s.best_support = 0;
		s.altsupp1 = s.altsupp2 = s.altsupp3 = 0;
		s.best_sub = stale_best_sub;
		s.altbase1 = stale_alt1;
		s.altbase2 = stale_alt2;
		s.altbase3 = stale_alt3;
		s.indel_len = 0;
		uint32_t num_deletions = 1;
		bool touched = false;

		if (P.snv && meets_edit(there)) {
			s.best_sub = draft;
			s.best_support = P.counting ? there_median : there;
		}

		const uint32_t cands = candidates(draft);
		const bool tail_is_pos = t.ni < nn && ty[t.ni] == 0;
		const bool tail_is_chr = t.ni < nn && ty[t.ni] == 1;
		for (uint32_t ci = 0; ci < 4; ci++) {
			const unsigned char sub = (unsigned char)((cands >> (8 * ci)) & 0xFF);
			if (!sub) {
				break;
			}
			ts = hs;
			hash_changelast(ts, draft, sub, P);
			if (!(P.mode == 2 || q_present_solid(ts))) {
				continue;
			}
			th = h;
			tt = t;
			touched = true;
			if (tail_is_pos) {
				patch_on = true;
				patch_pos = t.pos;
				patch_ch = sub;
			} else if (tail_is_chr) {
				ch[t.ni] = sub;
			}
			uint32_t present = 0;
			for (uint32_t q = 0; q < k && th.pos < io.len && tt.pos < io.len; q++) {
				if (!roll(th, tt, out, in)) {
					break;
				}
				hash_roll(ts, out, in, P);
				if (q % P.jump == 0 && q_present_solid(ts)) {
					present++;
				}
			}
			if (tail_is_pos) {
				patch_on = false;
			} else if (tail_is_chr) {
				ch[t.ni] = draft;
			}
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
					continue;
				}
			}
			if (P.mode == 2 || s.best_type != 1) {
				if (try_indels(draft, sub, num_deletions, s)) {
					if (P.mode == 0 || P.mode == 1) {
						break;
					}
				}
			}
		}
		stale_best_sub = s.best_sub;
		stale_alt1 = s.altbase1;
		stale_alt2 = s.altbase2;
		stale_alt3 = s.altbase3;

		// makeEdit, ntedit.cpp:1250-1448
		const uint8_t fl = (touched && raw != draft) ? EV_TOUCHED : 0;
		switch (s.best_type) {
		case 1:
			emit(1, fl, draft, s);
			if (tail_is_pos) {
				if (ov_n >= (uint32_t)OVCAP) {
					// drop substitutions the head has already passed
					uint32_t w = 0;
					for (uint32_t i = 0; i < ov_n; i++) {
						if (ov_pos[i] >= h.pos) {
							ov_pos[w] = ov_pos[i];
							ov_ch[w] = ov_ch[i];
							w++;
						}
					}
					ov_n = w;
				}
				if (ov_n < (uint32_t)OVCAP) {
					// a later substitution at the same position replaces the earlier one
					uint32_t i = 0;
					for (; i < ov_n; i++) {
						if (ov_pos[i] == t.pos) {
							break;
						}
					}
					ov_pos[i] = t.pos;
					ov_ch[i] = s.best_sub;
					if (i == ov_n) {
						ov_n++;
					}
				} else {
					status |= ST_ROPE_OVERFLOW;
				}
			} else if (tail_is_chr) {
				ch[t.ni] = s.best_sub;
			}
			hash_changelast(hs, draft, s.best_sub, P);
			break;
		case 2: {
			emit(2, fl, draft, s);
			if (insertion_guard_fires(s)) {
				return false;
			}
			rope_insert(t.ni, t.pos, s.indel, s.indel_len);
			hash_changelast(hs, draft, (unsigned char)s.indel[0], P);
			break;
		}
		case 3:
			emit(3, fl, draft, s);
			rope_delete(t.ni, t.pos, s.indel_len);
			hash_changelast(hs, draft, cchar(t), P);
			break;
		default:
			// soft-masking only changes the case of the tail char: no effect on the hash (ntedit.cpp:1410-1424)
			if (fl || P.mask || (P.snv && s.altsupp1)) {
				emit(0, fl, draft, s);
			}
			break;
		}
		return true;
	}

	// ---------------------------------------------------------------- clean-window handling
	NTB_FN bool window_clean() const
	{
		if (h.ni != t.ni || h.ni >= nn || ty[h.ni] != 0 || t.pos - h.pos != P.k - 1) {
			return false;
		}
		for (uint32_t i = 0; i < ov_n; i++) {
			if (ov_pos[i] >= h.pos) {
				return false;
			}
		}
		return true;
	}

	NTB_FN void reset_rope(uint32_t head_pos)
	{
		ty[0] = 0;
		ch[0] = 0;
		sp[0] = head_pos;
		ep[0] = io.len - 1;
		nn = 1;
		h.ni = t.ni = 0;
		ov_n = 0;
	}

	// first position >= from whose visit bit is set, or NONE32 when there is none below `limit`
	NTB_FN uint32_t next_visit(uint32_t from, uint32_t limit) const
	{
		if (from >= limit) {
			return NONE32;
		}
		uint64_t g = io.goff + from;
		const uint64_t gend = io.goff + limit;
		uint64_t w = g >> 5;
		uint32_t bits = io.visit[w] & (0xFFFFFFFFu << (g & 31));
		for (;;) {
			if (bits) {
#if defined(__CUDA_ARCH__)
				const uint64_t hit = (w << 5) + (uint32_t)(__ffs((int)bits) - 1);
#else
				const uint64_t hit = (w << 5) + (uint32_t)__builtin_ctz(bits);
#endif
				return hit < gend ? (uint32_t)(hit - io.goff) : NONE32;
			}
			w++;
			if ((w << 5) >= gend) {
				return NONE32;
			}
			bits = io.visit[w];
		}
	}

	NTB_FN void seed_at(uint32_t tail)
	{
		const uint32_t head = tail + 1 - P.k;
		h.pos = head;
		t.pos = tail;
		const unsigned char* base = io.text + head;
		hash_seed(hs, P.k, [base](unsigned i) { return base[i]; });
		char_in = io.text[tail];
	}

	// findFirstAcceptedKmer, ntedit.cpp:524-545
	NTB_FN uint32_t first_accepted_kmer() const
	{
		const uint32_t k = P.k;
		for (uint32_t i = 0; (uint64_t)i + k < io.len;) {
			if (is_accepted_any_case(io.text[i])) {
				bool good = true;
				for (uint32_t j = i + 1; j < i + k; j++) {
					if (!is_accepted_any_case(io.text[j])) {
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
		return io.len - 1;
	}

	// drop rope nodes that can no longer be reached so that long dirty stretches fit the bounded array
	NTB_FN void compact()
	{
		uint32_t lo = h.ni < t.ni ? h.ni : t.ni;
		// keep the run of inserted characters left of the tail (getPrevInsertion walks it) plus one node
		uint32_t r = t.ni;
		while (r > 0 && ty[r - 1] == 1) {
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
		for (uint32_t i = lo; i < nn; i++) {
			move_node(i - lo, i);
		}
		nn -= lo;
		h.ni -= lo;
		t.ni -= lo;
	}

	// ---------------------------------------------------------------- the main loop, ntedit.cpp:1797-2139
	NTB_FN void run(const Task& task, TaskResult& res)
	{
		const uint32_t k = P.k;
		nn = 0;
		ov_n = 0;
		patch_on = false;
		if (task.flags & TASK_CONTIG_START) {
			stale_best_sub = stale_alt1 = stale_alt2 = stale_alt3 = 0;
		} else {
			stale_best_sub = STALE_REF | 0;
			stale_alt1 = STALE_REF | 1;
			stale_alt2 = STALE_REF | 2;
			stale_alt3 = STALE_REF | 3;
		}
		adv = 0;
		anchored = true;
		last_event = NONE32;
		n_events = n_sites = 0;
		first_touch = NONE32;
		status = 0;
		char_in = 0;
		hs.fh = hs.rh = 0;
		uint32_t end_pos = io.len;
		bool need_seed = true;
		bool alive = true;

		if (task.flags & TASK_CONTIG_START) {
			const uint32_t h0 = first_accepted_kmer();
			if ((uint64_t)h0 + k - 1 >= io.len) {
				status |= ST_CONTIG_END;
				alive = false;
			} else {
				h.pos = h0;
				t.pos = h0 + k - 1;
			}
		} else {
			t.pos = task.start;
			h.pos = task.start + 1 - k;
		}
		if (alive) {
			reset_rope(h.pos);
		}

		while (alive) {
			if ((uint64_t)h.pos + k - 1 >= io.len) {
				status |= ST_CONTIG_END;
				break;
			}
			if (status & (ST_EV_OVERFLOW | ST_ROPE_OVERFLOW)) {
				break;
			}
			if (need_seed || window_clean()) {
				// clean window: forget the local rope and jump to the next position K1 flagged
				reset_rope(h.pos);
				anchored = true;
				if (t.pos >= task.end) {
					end_pos = t.pos;
					break;
				}
				const uint32_t nv = next_visit(t.pos, task.end);
				if (nv == NONE32) {
					end_pos = task.end;
					break;
				}
				if (nv != t.pos || need_seed) {
					seed_at(nv);
					reset_rope(h.pos);
				}
				need_seed = false;
			} else if (nn + 16 > (uint32_t)NCAP) {
				compact();
				if (nn + 16 > (uint32_t)NCAP) {
					status |= ST_ROPE_OVERFLOW;
					break;
				}
			}

			bool site;
			if (P.snv) {
				site = true;
			} else if (P.counting) {
				const unsigned c = q_count(hs);
				site = c == 0 || c < P.min_threshold;
			} else {
				site = !q_contains(hs);
			}
			if (site) {
				if (!evaluate_site()) {
					status |= ST_CONTIG_END;
					break;
				}
			}

			if (window_clean()) {
				// still on unedited text: the next position the reference acts on is the next flagged one
				// (its own roll / skip-after-N loop, ntedit.cpp:2118-2138, does nothing observable in between)
				h.pos++;
				t.pos++;
				need_seed = true;
				continue;
			}

			// advance; after a non-accepted incoming base skip until k further positions were consumed (ntedit.cpp:2118-2138)
			int64_t target = -1;
			bool stop = false;
			do {
				unsigned char out = 0;
				if (roll(h, t, out, char_in)) {
					adv++;
					if (!is_accepted_any_case(char_in)) {
						target = (int64_t)(int32_t)t.pos + (int64_t)(int32_t)k;
					}
					hash_roll(hs, out, char_in, P);
				} else {
					stop = true;
					break;
				}
				if (target >= 0 && (int64_t)(int32_t)t.pos != target && window_clean()) {
					// skipping over non-accepted bases on unedited text: same shortcut as above
					h.pos++;
					t.pos++;
					need_seed = true;
					break;
				}
			} while (target >= 0 && (int64_t)(int32_t)t.pos != target);
			if (stop) {
				status |= ST_CONTIG_END;
				break;
			}
		}
		status |= ST_DONE;
		res.end_pos = (status & ST_CONTIG_END) ? io.len : end_pos;
		res.first_touch = first_touch;
		res.last_event = last_event;
		res.n_events = n_events;
		res.n_sites = n_sites;
		res.status = status;
		res.stale[0] = stale_best_sub;
		res.stale[1] = stale_alt1;
		res.stale[2] = stale_alt2;
		res.stale[3] = stale_alt3;
	}
};

} // namespace ntb
