// Dense form of the first pre-evaluation pass: ONE THREAD evaluates one site -- the part of ntEdit's edit block
// (ntedit.cpp:1808-2116) that does not need tryIndels -- reached with a clean window, straight from the contig text:
//   check-missing subset (ntedit.cpp:1826-1873), substitution gates (:1917-1928), substitution trials (:1936-2062).
// It restates, lane-free, exactly what Walker::evaluate_site_core(false) computes on a clean window (engine.h: linearise's
// simple path, compute_plain, phase_fast_check, site_after_check, phase_fast_subs, site_candidate, the quiet test); the
// two are compared record for record by tests/test_hostsim.py.  A warp-per-site walker spends 31 of 32 lanes waiting in
// every serial section and keeps ~10 DRAM round trips per site in sequence; here 32 sites per warp are in flight, every
// probe of a stage is issued before any is consumed, and the kernel is bound by the random-probe rate of the filter.
// Compiles for the device (presite_dense_kernel) and for the host (tests/hostsim).
#pragma once
#include "engine.h"

namespace ntb {

struct DenseCtx
{
	const KParams* kp;
	FilterView bloom, rep;
	const uint64_t* rot;  // ROT_WORDS entries (rot_entry)
	const uint8_t* cls;   // 256 class bytes: bits 0-2 forward seed code, 3-5 reverse seed code, 6 accepted
};

NTB_FN inline uint32_t
dense_probe(const uint8_t* p)
{
	return probe_byte(p);
}

// Filter values of up to B k-mers (canonical hashes hb[0..n)): bit filter -> 1 when all hash_num bits are set else 0,
// counting filter -> min counter (BFWrapper::contains / get_count, ntedit.cpp:368-376).  ALL: every hash function's probe is
// issued at once (k-mers that are mostly present); otherwise one hash function per pass and only the undecided k-mers go
// on (mostly absent k-mers: btllib's early exit).
template<int B, bool ALL>
NTB_FN inline void
dense_values(const FilterView& F, uint32_t k, const uint64_t (&hb)[B], uint32_t n, uint32_t (&val)[B])
{
	const bool counting = F.counting != 0;
	uint32_t live = n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u);
#pragma unroll
	for (int i = 0; i < B; i++) {
		val[i] = counting ? 255u : 1u;
	}
	if (ALL && F.hash_num <= 4) {
		uint32_t got[B][4];
#pragma unroll
		for (int i = 0; i < B; i++) {
#pragma unroll
			for (int h = 0; h < 4; h++) {
				got[i][h] = counting ? 255u : 1u;
				if (((live >> i) & 1u) && (uint32_t)h < F.hash_num) {
					const uint64_t slot = filter_slot(F, hash_extend(hb[i], k, (unsigned)h));
					const uint32_t b = dense_probe(F.data + (counting ? slot : (slot >> 3)));
					got[i][h] = counting ? b : ((b >> ((uint32_t)slot & 7u)) & 1u);
				}
			}
		}
#pragma unroll
		for (int i = 0; i < B; i++) {
#pragma unroll
			for (int h = 0; h < 4; h++) {
				val[i] = counting ? (got[i][h] < val[i] ? got[i][h] : val[i]) : (val[i] & got[i][h]);
			}
		}
		return;
	}
	for (uint32_t h = 0; h < F.hash_num && live; h++) {
		uint32_t got[B];
#pragma unroll
		for (int i = 0; i < B; i++) {
			got[i] = counting ? 255u : 1u;
			if ((live >> i) & 1u) {
				const uint64_t slot = filter_slot(F, hash_extend(hb[i], k, h));
				const uint32_t b = dense_probe(F.data + (counting ? slot : (slot >> 3)));
				got[i] = counting ? b : ((b >> ((uint32_t)slot & 7u)) & 1u);
			}
		}
#pragma unroll
		for (int i = 0; i < B; i++) {
			if ((live >> i) & 1u) {
				val[i] = counting ? (got[i] < val[i] ? got[i] : val[i]) : (val[i] & got[i]);
				if (val[i] == 0) {
					live &= ~(1u << i);
				}
			}
		}
	}
}

// substitution candidates, ntedit.cpp:178-199 (Walker::candidates)
NTB_FN inline uint32_t
dense_candidates(unsigned char draft, bool snv, bool accepted)
{
#define NTB_PACK(a, b, c, d) ((uint32_t)(a) | ((uint32_t)(b) << 8) | ((uint32_t)(c) << 16) | ((uint32_t)(d) << 24))
	switch (draft) {
	case 'A': return NTB_PACK('T', 'C', 'G', 0);
	case 'T': return NTB_PACK('A', 'C', 'G', 0);
	case 'C': return NTB_PACK('A', 'T', 'G', 0);
	case 'G': return NTB_PACK('A', 'T', 'C', 0);
	default: break;
	}
	if (snv) {
		return accepted || draft == 'N' ? NTB_PACK('A', 'T', 'C', 'G') : 0u;
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

// The site at tail position `pos` of the contig `text[0..len)`, window clean.  Returns SITE_NONE / SITE_DONE / SITE_PENDING
// and fills every field of `r` except the key (the fields Walker::pre_fill fills).  KCAP >= k.
template<int KCAP>
NTB_FN inline uint32_t
dense_site(const DenseCtx& C, const uint8_t* text, uint32_t len, uint32_t pos, SiteRec& r)
{
	const KParams& P = *C.kp;
	const uint32_t k = P.k, jump = P.jump;
	const bool counting = P.counting != 0, snv = P.snv != 0;
	const uint64_t* rot = C.rot;
	const uint8_t* cls = C.cls;
	constexpr int B = 8;

	// Walker::pre_fill's defaults
	r.state = SITE_NONE;
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

	// ---- site_begin + linearise (simple path): class bytes of the window and of the k bases behind it
	const unsigned char raw = text[pos];
	const unsigned char draft = to_upper(raw);
	r.draft = draft;
	const uint32_t avail = len - 1 - pos;
	const uint32_t n_rolls = avail < k + MAX_DELETIONS + 1 ? avail : k + MAX_DELETIONS + 1;
	const uint32_t n_sub = n_rolls < k ? n_rolls : k;
	uint8_t cw[2 * KCAP]; // cw[i] = class of text[pos + 1 - k + i], i < k + n_sub
	const uint32_t head = pos + 1 - k;
#if defined(__CUDA_ARCH__)
	{
		// aligned 32-bit loads (SCAN_HALO bytes in front of the batch buffer and its zero padding behind make the rounded-out
		// words readable)
		const uintptr_t a0 = reinterpret_cast<uintptr_t>(text + head);
		const uint32_t sh = (uint32_t)(a0 & 3u);
		const uint32_t* wp = reinterpret_cast<const uint32_t*>(a0 - sh);
		uint32_t word = 0;
		for (uint32_t i = 0; i < k + n_sub; i++) {
			const uint32_t q = sh + i;
			if (i == 0 || (q & 3u) == 0) {
				word = __ldg(wp + (q >> 2));
			}
			cw[i] = cls[(word >> (8u * (q & 3u))) & 0xFFu];
		}
	}
#else
	for (uint32_t i = 0; i < k + n_sub; i++) {
		cw[i] = cls[text[head + i]];
	}
#endif
	uint32_t n_check = k; // iterations of the check loop that complete (ntedit.cpp:1826-1858)
	for (uint32_t m = 0; m < k; m++) {
		if (m >= n_rolls || !((cw[k + m] >> 6) & 1u)) {
			n_check = m;
			break;
		}
	}
	const bool dnf = n_check < k;
	if (!snv && dnf) {
		return SITE_NONE;
	}

	// ---- compute_plain: hash state after R rolls, R = 0 .. n_sub
	uint64_t pf[KCAP + 1], pr[KCAP + 1];
	{
		uint64_t f = 0, rv = 0;
		for (uint32_t i = 0; i < k; i++) {
			const uint32_t c = cw[i];
			f ^= rot[(c & 7u) * ROT_STRIDE + (k - 1 - i)];
			rv ^= rot[((c >> 3) & 7u) * ROT_STRIDE + i];
		}
		pf[0] = f;
		pr[0] = rv;
		for (uint32_t R = 0; R < n_sub; R++) {
			const uint32_t ci = cw[k + R], co = cw[R];
			f = srol1(f) ^ rot[(ci & 7u) * ROT_STRIDE] ^ rot[(co & 7u) * ROT_STRIDE + k];
			rv = sror1(rv ^ rot[((ci >> 3) & 7u) * ROT_STRIDE + k] ^ rot[((co >> 3) & 7u) * ROT_STRIDE]);
			pf[R + 1] = f;
			pr[R + 1] = rv;
		}
	}

	// ---- phase_fast_check + site_after_check
	const uint32_t nC = n_check ? (n_check - 1) / jump + 1 : 0;
	const bool atgc = is_atgc_upper(draft);
	uint32_t missing = 0, there = 0, nmed = 0;
	uint8_t med[KCAP + 1];
	for (uint32_t j0 = 0; j0 < nC; j0 += B) {
		uint64_t hb[B];
		uint32_t val[B];
		const uint32_t n = nC - j0 < (uint32_t)B ? nC - j0 : (uint32_t)B;
#pragma unroll
		for (int i = 0; i < B; i++) {
			const uint32_t R = ((uint32_t)i < n ? j0 + i : j0) * jump + 1;
			hb[i] = pf[R] + pr[R];
		}
		dense_values<B, false>(C.bloom, k, hb, n, val);
#pragma unroll
		for (int i = 0; i < B; i++) {
			if ((uint32_t)i < n) {
				const uint32_t c = val[i];
				if (c == 0) {
					missing++;
				} else if (counting) {
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
		}
	}
	uint32_t there_median = 0;
	if (counting && nmed > 0) {
		for (uint32_t a = 1; a < nmed; a++) { // upper median (median(), ntedit.cpp:455-463)
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
	const bool attempt = snv || (!dnf && (missing >= P.thr_missing || (counting && there_median < P.min_threshold)));
	if (!attempt) {
		return SITE_NONE;
	}
	uint32_t best_type = 0, best_support = 0, altsupp1 = 0, altsupp2 = 0, altsupp3 = 0;
	uint8_t best_sub = STALE_REF | 0, altbase1 = STALE_REF | 1, altbase2 = STALE_REF | 2, altbase3 = STALE_REF | 3;
	bool touched = false;
	if (snv && there >= P.thr_edit) {
		best_sub = draft;
		best_support = counting ? there_median : there;
	}

	// ---- phase_fast_subs
	const uint32_t cands = dense_candidates(draft, snv, ((cls[draft] >> 6) & 1u) != 0);
	uint32_t ncand = 0;
	while (ncand < 4 && ((cands >> (8 * ncand)) & 0xFF) != 0) {
		ncand++;
	}
	const uint32_t dcl = cls[draft];
	const uint32_t df = dcl & 7u, dr = (dcl >> 3) & 7u;
	auto solid_value = [&](uint32_t c) -> bool { return counting ? !(c == 0 || c < P.min_threshold || c > P.max_threshold) : c != 0; };
	auto is_site_value = [&](uint32_t c) -> bool { return snv || c == 0 || (counting && c < P.min_threshold); };
	uint32_t gate = 0; // bit c: candidate c's changed window is present && solid
	if (P.mode != 2 && ncand) {
		uint64_t hb[B];
		uint32_t val[B];
#pragma unroll
		for (int i = 0; i < B; i++) {
			const uint32_t c = (uint32_t)i < ncand ? (uint32_t)i : 0u;
			const uint32_t xcl = cls[(cands >> (8 * c)) & 0xFF];
			const uint64_t f = pf[0] ^ rot[df * ROT_STRIDE] ^ rot[(xcl & 7u) * ROT_STRIDE];
			const uint64_t rv = pr[0] ^ rot[dr * ROT_STRIDE + (k - 1)] ^ rot[((xcl >> 3) & 7u) * ROT_STRIDE + (k - 1)];
			hb[i] = f + rv;
		}
		dense_values<B, false>(C.bloom, k, hb, ncand, val);
		uint32_t solid = 0;
		for (uint32_t c = 0; c < ncand; c++) {
			if (solid_value(val[c])) {
				solid |= 1u << c;
			}
		}
		if (P.h_rep && solid) {
			// solid k-mers additionally must be absent from the secondary filter (-e), ntedit.cpp:467-468
			uint32_t rv[B];
			dense_values<B, false>(C.rep, k, hb, ncand, rv);
			for (uint32_t c = 0; c < ncand; c++) {
				if (((solid >> c) & 1u) && rv[c] != 0) {
					solid &= ~(1u << c);
				}
			}
		}
		gate = solid;
	}
	uint32_t sup[4] = { 0, 0, 0, 0 };
	uint32_t loud = 0; // bit c: one of the k-1 windows that contain candidate c's base is a site
	// the candidates that are tried, packed: the a-th pass of the loop below serves every thread's a-th tried candidate (which
	// of the three bases passes its gate differs from site to site; looping over c would run the trial once per c with a
	// third of the warp)
	const uint32_t tried = P.mode == 2 ? ((1u << ncand) - 1u) : gate;
	for (uint32_t rest = tried; rest; rest &= rest - 1) {
		uint32_t c = 0;
		while (!((rest >> c) & 1u)) {
			c++;
		}
		const uint32_t xcl = cls[(cands >> (8 * c)) & 0xFF];
		const uint32_t xf = xcl & 7u, xr = (xcl >> 3) & 7u;
		// -s 1 never jumps behind a substitution (every position is a site), so only the sampled windows are read there:
		// window index w stands for R = 1 + w * stride rolls
		const uint32_t stride = snv ? jump : 1u;
		const uint32_t n_win = n_sub ? (n_sub - 1) / stride + 1 : 0;
		for (uint32_t w0 = 0; w0 < n_win; w0 += B) {
			uint64_t hb[B];
			uint32_t val[B];
			const uint32_t n = n_win - w0 < (uint32_t)B ? n_win - w0 : (uint32_t)B;
			const uint32_t R0 = 1 + w0 * stride;
#pragma unroll
			for (int i = 0; i < B; i++) {
				const uint32_t R = (uint32_t)i < n ? R0 + i * stride : R0;
				uint64_t f = pf[R], rv = pr[R];
				if (R < k) {
					f ^= rot[df * ROT_STRIDE + R] ^ rot[xf * ROT_STRIDE + R];
					rv ^= rot[dr * ROT_STRIDE + (k - 1 - R)] ^ rot[xr * ROT_STRIDE + (k - 1 - R)];
				}
				hb[i] = f + rv;
			}
			dense_values<B, true>(C.bloom, k, hb, n, val);
			uint32_t solid = 0;
#pragma unroll
			for (int i = 0; i < B; i++) {
				if ((uint32_t)i < n) {
					if (solid_value(val[i])) {
						solid |= 1u << i;
					}
					if (is_site_value(val[i]) && R0 + i * stride + 1 <= k) {
						loud |= 1u << c;
					}
				}
			}
			if (P.h_rep && solid) {
				uint32_t rv[B];
				dense_values<B, true>(C.rep, k, hb, n, rv);
#pragma unroll
				for (int i = 0; i < B; i++) {
					if (((solid >> i) & 1u) && rv[i] != 0) {
						solid &= ~(1u << i);
					}
				}
			}
#pragma unroll
			for (int i = 0; i < B; i++) {
				if ((uint32_t)i < n && ((solid >> i) & 1u) && (R0 + i * stride - 1) % jump == 0) {
					sup[c]++;
				}
			}
		}
	}

	// ---- the candidate loop up to the first tryIndels call (Walker::site_candidate, ntedit.cpp:1917-2092)
	uint32_t state = SITE_DONE;
	for (uint32_t ci = 0; ci < 4; ci++) {
		const unsigned char sub = (unsigned char)((cands >> (8 * ci)) & 0xFF);
		if (!sub) {
			break;
		}
		if (!(P.mode == 2 || ((gate >> ci) & 1u))) {
			continue;
		}
		touched = true;
		const uint32_t present = sup[ci];
		if (present >= P.thr_edit) {
			if (present >= best_support) {
				if (altsupp2) {
					altbase3 = altbase2;
					altsupp3 = altsupp2;
				}
				if (altsupp1) {
					altbase2 = altbase1;
					altsupp2 = altsupp1;
				}
				if (best_support) {
					altsupp1 = best_support;
					altbase1 = best_sub;
				}
				best_type = 1;
				best_sub = sub;
				best_support = present;
			} else if (!altsupp1) {
				altbase1 = sub;
				altsupp1 = present;
			} else if (!altsupp2) {
				if (present < altsupp1) {
					altbase2 = sub;
					altsupp2 = present;
				} else {
					altbase2 = altbase1;
					altsupp2 = altsupp1;
					altbase1 = sub;
					altsupp1 = present;
				}
			} else if (!altsupp3) {
				if (present < altsupp2) {
					altbase3 = sub;
					altsupp3 = present;
				} else if (present < altsupp1) {
					altbase3 = altbase2;
					altsupp3 = altsupp2;
					altbase2 = sub;
					altsupp2 = present;
				} else {
					altbase3 = altbase2;
					altsupp3 = altsupp2;
					altbase2 = altbase1;
					altsupp2 = altsupp1;
					altbase1 = sub;
					altsupp1 = present;
				}
			}
			if (P.mode == 0 || P.mode == 1) {
				continue;
			}
		}
		if (P.mode == 2 || best_type != 1) {
			// tryIndels would be called here (with nothing to try it returns at once and the loop goes on)
			if (P.max_ins_tries != 0) {
				state = SITE_PENDING;
				break;
			}
		}
	}
	r.state = (uint8_t)state;
	if (state != SITE_DONE) {
		return state;
	}
	// ---- the quiet test of evaluate_site_core and pre_fill
	bool quiet = false;
	if (best_type == 1 && !snv && !dnf && n_check == k && n_rolls >= k) {
		uint32_t c = 0;
		while (c < 4 && ((cands >> (8 * c)) & 0xFF) != best_sub) {
			c++;
		}
		quiet = c < 4 && !((loud >> c) & 1u);
	}
	r.flags = (uint8_t)(((touched && raw != draft) ? SITE_FL_TOUCHED : 0) | (quiet ? SITE_FL_QUIET : 0));
	r.best_type = (uint8_t)best_type;
	r.best_sub = best_sub;
	r.support = (uint16_t)best_support;
	r.altsupp[0] = (uint16_t)altsupp1;
	r.altsupp[1] = (uint16_t)altsupp2;
	r.altsupp[2] = (uint16_t)altsupp3;
	r.altbase[0] = altbase1;
	r.altbase[1] = altbase2;
	r.altbase[2] = altbase3;
	return state;
}

// first flagged position in [from, limit) of the contig at text offset goff, NONE32 if there is none (one thread)
NTB_FN inline uint32_t
dense_next_visit(const uint32_t* visit, uint64_t goff, uint32_t from, uint32_t limit)
{
	if (from >= limit) {
		return NONE32;
	}
	const uint64_t g = goff + from, gend = goff + limit;
	for (uint64_t w = g >> 5; (w << 5) < gend; w++) {
		uint32_t bits = visit[w];
		if (w == (g >> 5)) {
			bits &= 0xFFFFFFFFu << (g & 31);
		}
		if (bits) {
			uint32_t b = 0;
			while (!((bits >> b) & 1u)) {
				b++;
			}
			const uint64_t hit = (w << 5) + b;
			return hit < gend ? (uint32_t)(hit - goff) : NONE32;
		}
	}
	return NONE32;
}

// files the record of an evaluated site (key = text position + 1) and, for SITE_PENDING, lists it for the second pass;
// false when the table has no room (the record is dropped)
NTB_FN inline bool
dense_commit(SiteRec& r, uint32_t st, uint64_t goff, uint32_t task_idx, uint32_t pos, SiteRec* table, uint32_t table_mask, PendingSite* pending,
             uint32_t pending_cap, Counters* ctr)
{
	r.key = goff + pos + 1;
	const uint32_t slot = site_table_insert(table, table_mask, r.key);
	if (slot == NONE32) {
#if defined(__CUDA_ARCH__)
		atomicAdd(&ctr->n_dropped, 1u);
#else
		ctr->n_dropped++;
#endif
		return false;
	}
	if (st == SITE_PENDING) {
		uint32_t idx;
#if defined(__CUDA_ARCH__)
		idx = atomicAdd(&ctr->n_pending, 1u);
#else
		idx = ctr->n_pending++;
#endif
		if (idx < pending_cap) {
			PendingSite ps;
			ps.task = task_idx;
			ps.pos = pos;
			ps.slot = slot;
			pending[idx] = ps;
		}
	}
	table[slot] = r;
	return true;
}

// -s 1: does committing this site do anything observable -- an edit, or an event (a variant record, a case change of the
// tail char, -a masking)?  Sites that do not are skipped by the walk altogether (Walker::site_commit, default case).
NTB_FN inline bool
dense_has_effect(uint32_t st, const SiteRec& r, bool mask)
{
	return st == SITE_DONE && (r.best_type != 0 || (r.flags & SITE_FL_TOUCHED) || mask || r.altsupp[0] != 0);
}

// does the main loop go on to the next flagged position behind this site with a clean window (no edit was made)?
NTB_FN inline bool
dense_continues(uint32_t st, const SiteRec& r)
{
	return st == SITE_NONE || (st == SITE_DONE && r.best_type == 0);
}

// the chain's next site behind `pos`: the next flagged position within pre_gap() = k-1 (a farther one is a head), or NONE32
NTB_FN inline uint32_t
dense_chain_next(const uint32_t* visit, uint64_t goff, uint32_t len, uint32_t pos, uint32_t gap)
{
	const uint32_t lim = len - pos - 1 < gap ? len : pos + 1 + gap;
	return dense_next_visit(visit, goff, pos + 1, lim);
}

// One site of the first pass, one thread: evaluates the site at `pos`, files its record, and returns the position the chain
// goes on with or NONE32 when the chain ends here (engine.h: one iteration of Walker::pre_run with allow_indels == false).
template<int KCAP>
NTB_FN inline uint32_t
dense_step(const DenseCtx& C, const uint8_t* text, uint32_t len, uint64_t goff, const uint32_t* visit, uint32_t task_idx, uint32_t pos,
           SiteRec* table, uint32_t table_mask, PendingSite* pending, uint32_t pending_cap, Counters* ctr)
{
	SiteRec r;
	const uint32_t st = dense_site<KCAP>(C, text, len, pos, r);
	if (!dense_commit(r, st, goff, task_idx, pos, table, table_mask, pending, pending_cap, ctr) || !dense_continues(st, r)) {
		return NONE32;
	}
	return dense_chain_next(visit, goff, len, pos, C.kp->k - 1);
}

// One chain round as the device runs it (kernels.cu: presite_dense_kernel<KCAP, true>), for host builds: the next DENSE_GROUP
// sites of the chain behind `behind` evaluated "side by side", records kept up to the first one at which the chain stops,
// no-edit records told how far the walker may jump.  Returns the position the next round goes on behind, or NONE32.
template<int KCAP>
inline uint32_t
dense_chain_round_host(const DenseCtx& C, const uint8_t* text, uint32_t len, uint64_t goff, const uint32_t* visit, uint32_t task_idx,
                       uint32_t behind, SiteRec* table, uint32_t table_mask, PendingSite* pending, uint32_t pending_cap, Counters* ctr)
{
	const uint32_t gap = C.kp->k - 1;
	uint32_t pos[DENSE_GROUP], st[DENSE_GROUP];
	SiteRec r[DENSE_GROUP];
	bool active[DENSE_GROUP];
	uint32_t p = behind, first_stop = DENSE_GROUP;
	for (int q = 0; q < DENSE_GROUP; q++) {
		p = p != NONE32 ? dense_chain_next(visit, goff, len, p, gap) : NONE32;
		pos[q] = p;
		active[q] = p != NONE32;
		st[q] = SITE_NONE;
		if (active[q]) {
			st[q] = dense_site<KCAP>(C, text, len, p, r[q]);
		}
		if (first_stop == (uint32_t)DENSE_GROUP && (!active[q] || !dense_continues(st[q], r[q]))) {
			first_stop = (uint32_t)q;
		}
	}
	const bool skips = !C.kp->mask && !C.kp->snv;
	bool ok_last = true;
	for (uint32_t q = 0; q < (uint32_t)DENSE_GROUP; q++) {
		if (!active[q] || q > first_stop) {
			continue;
		}
		if (skips && q < first_stop && dense_skippable(st[q], r[q].best_type, r[q].flags)) {
			uint32_t T = SKIP_IDENTITY;
			uint32_t n = 0, dist = 0;
			for (uint32_t j = q + 1; j < first_stop && dense_skippable(st[j], r[j].best_type, r[j].flags); j++) {
				T = dense_skip_compose(T, st[j], dense_pack_bases(r[j]));
				n++;
				dist = pos[j] - pos[q];
			}
			if (n) {
				dense_skip_store(r[q], n, dist, T);
			}
		}
		const bool ok = dense_commit(r[q], st[q], goff, task_idx, pos[q], table, table_mask, pending, pending_cap, ctr);
		if (q == (uint32_t)DENSE_GROUP - 1) {
			ok_last = ok;
		}
	}
	return first_stop == (uint32_t)DENSE_GROUP && ok_last ? pos[DENSE_GROUP - 1] : NONE32;
}

} // namespace ntb
