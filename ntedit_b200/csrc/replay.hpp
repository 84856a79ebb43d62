// Host side of the polishing path: rebuilds, per contig, the reference's rope (std::vector<seqNode>, ntedit.cpp:613-620)
// and substitution queue (sRec, ntedit.cpp:598-611) from the makeEdit-level events the device walkers emit.
// Only rope surgery lives here -- makeInsertion / makeDeletion / the rope side of makeEdit (ntedit.cpp:625-809,
// 1250-1448); every hash and every filter probe happened on the device.
#pragma once
#include "../../include/ntedit_b200.h"
#include "nthash.h"

#include <cstring>
#include <string>
#include <vector>

namespace ntb {

class RopeReplay
{
  public:
	RopeReplay(char* seq, uint32_t len, uint32_t k, uint32_t insertion_cap, int snv, int mask, size_t reserve_nodes = 0,
	           size_t reserve_recs = 0)
	  : seq_(seq), len_(len), k_(k), cap_(insertion_cap), snv_(snv), mask_(mask)
	{
		if (reserve_nodes) {
			rope.reserve(reserve_nodes);
		}
		if (reserve_recs) {
			recs.reserve(reserve_recs);
		}
		ntb_node root;
		std::memset(&root, 0, sizeof root);
		root.node_type = 0;
		root.s_pos = 0;
		root.e_pos = len - 1;
		rope.push_back(root);
	}

	std::vector<ntb_node> rope;
	std::vector<ntb_srec> recs;
	bool ended = false;   // the reference's main loop stopped (insertion guard)
	std::string error;

	// apply one event; returns false on an internal inconsistency (error is set)
	bool apply(const Event& e)
	{
		if (ended) {
			return true;
		}
		if (e.advance == NONE32) {
			// the walker's window was clean: the tail sits on the last live node, which is a position node
			while (t_ni_ + 1 < rope.size() && rope[t_ni_ + 1].node_type != -1) {
				t_ni_++;
			}
			const ntb_node& nd = rope[t_ni_];
			if (nd.node_type != 0 || e.t_pos < nd.s_pos || e.t_pos > nd.e_pos) {
				return fail("anchored event outside the final position node");
			}
			t_pos_ = e.t_pos;
		} else {
			for (uint32_t i = 0; i < e.advance; i++) {
				step();
			}
			if (t_pos_ != e.t_pos) {
				return fail("tail cursor out of sync with the device walker");
			}
		}
		if (t_ni_ >= rope.size()) {
			return fail("tail node index past the rope");
		}
		if (e.flags & EV_TOUCHED) {
			set_tail_char(e.draft); // trial patch reverted with the upper-cased draft char, ntedit.cpp:1975-1981
		}
		switch (e.kind) {
		case 1: edit_substitution(e); break;
		case 2: edit_insertion(e); break;
		case 3:
			rope_delete(e.indel_len, e.support);
			break;
		default:
			if (mask_) {
				set_tail_char(to_lower(e.draft)); // ntedit.cpp:1410-1424
			}
			if (snv_ && e.altsupp[0]) { // ntedit.cpp:1428-1443
				ntb_srec r;
				std::memset(&r, 0, sizeof r);
				r.pos = t_pos_;
				r.draft_char = e.draft;
				r.sub_base = e.draft;
				r.num_support = e.support;
				r.altbase1 = e.altbase[0];
				r.altsupp1 = e.altsupp[0];
				r.altbase2 = e.altbase[1];
				r.altsupp2 = e.altsupp[1];
				r.altbase3 = e.altbase[2];
				r.altsupp3 = e.altsupp[2];
				recs.push_back(r);
			}
			break;
		}
		return error.empty();
	}

  private:
	char* seq_;
	uint32_t len_, k_, cap_;
	int snv_, mask_;
	uint32_t t_pos_ = 0, t_ni_ = 0;

	bool fail(const char* what)
	{
		error = what;
		return false;
	}

	void set_tail_char(unsigned char c)
	{
		ntb_node& nd = rope[t_ni_];
		if (nd.node_type == 0) {
			if (seq_ && t_pos_ < len_) {
				seq_[t_pos_] = (char)c;
			}
		} else if (nd.node_type == 1) {
			nd.c = c;
		}
	}

	// increment, ntedit.cpp:826-844
	void step_cursor(uint32_t& pos, uint32_t& ni) const
	{
		if (ni >= rope.size()) {
			return;
		}
		const ntb_node& nd = rope[ni];
		if (nd.node_type == 0) {
			pos++;
			if (pos > nd.e_pos) {
				ni++;
				if (ni < rope.size() && rope[ni].node_type == 0) {
					pos = rope[ni].s_pos;
				}
			}
		} else if (nd.node_type == 1) {
			ni++;
			if (ni < rope.size() && rope[ni].node_type == 0) {
				pos = rope[ni].s_pos;
			}
		}
	}

	void step() { step_cursor(t_pos_, t_ni_); }

	void put(size_t i, const ntb_node& nd)
	{
		if (i < rope.size()) {
			rope[i] = nd;
		} else {
			rope.push_back(nd);
		}
	}

	static ntb_node char_node(unsigned char c, uint32_t support)
	{
		ntb_node nd;
		std::memset(&nd, 0, sizeof nd);
		nd.node_type = 1;
		nd.c = c;
		nd.num_support = support;
		return nd;
	}

	// case 1 of makeEdit, ntedit.cpp:1280-1311
	void edit_substitution(const Event& e)
	{
		ntb_node& nd = rope[t_ni_];
		if (nd.node_type == 0) {
			if (seq_ && t_pos_ < len_) {
				seq_[t_pos_] = (char)e.base;
			}
			ntb_srec r;
			std::memset(&r, 0, sizeof r);
			r.pos = t_pos_;
			r.draft_char = e.draft;
			r.sub_base = e.base;
			r.num_support = e.support;
			if (e.altsupp[0] && e.altbase[0] != e.base) {
				r.altbase1 = e.altbase[0];
				r.altsupp1 = e.altsupp[0];
			}
			if (e.altsupp[1] && e.altbase[1] != e.altbase[0]) {
				r.altbase2 = e.altbase[1];
				r.altsupp2 = e.altsupp[1];
			}
			if (e.altsupp[2] && e.altbase[2] != e.altbase[1]) {
				r.altbase3 = e.altbase[2];
				r.altsupp3 = e.altsupp[2];
			}
			recs.push_back(r);
		} else if (nd.node_type == 1) {
			nd.c = e.base;
		}
	}

	static char revcomp(unsigned char c)
	{
		switch (c) { // RC(), ntedit.cpp:501-520
		case 'A': case 'a': return 'T';
		case 'T': case 't': return 'A';
		case 'G': case 'g': return 'C';
		case 'C': case 'c': return 'G';
		default: return 'N';
		}
	}

	// isRepeatInsertion / computeLPSArray, ntedit.cpp:561-596
	static bool is_repeat(const std::string& s)
	{
		const int n = (int)s.size();
		if (n <= 0) {
			return false;
		}
		std::vector<int> lps((size_t)n, 0);
		int l = 0, i = 1;
		while (i < n) {
			if (s[(size_t)i] == s[(size_t)l]) {
				lps[(size_t)i++] = ++l;
			} else if (l != 0) {
				l = lps[(size_t)l - 1];
			} else {
				lps[(size_t)i++] = 0;
			}
		}
		const int last = lps[(size_t)n - 1];
		return last > 0 && n % (n - last) == 0;
	}

	// getPrevInsertion, ntedit.cpp:907-922
	std::string prev_insertion() const
	{
		std::string out;
		uint32_t ni = t_ni_;
		if ((ni < rope.size() && rope[ni].node_type == 0 && t_pos_ == rope[ni].s_pos) || rope[ni].node_type == 1) {
			ni--;
		}
		while (ni < rope.size() && rope[ni].node_type == 1) {
			out += revcomp(rope[ni].c);
			ni--;
		}
		return out;
	}

	// the removal loops of makeEdit's insertion guard, ntedit.cpp:1321-1334 and 1352-1366
	void guard_remove(size_t count)
	{
		unsigned j = 1;
		if (t_ni_ < rope.size() && rope[t_ni_].node_type == 0 && t_pos_ == rope[t_ni_].s_pos) {
			j = 0;
		}
		for (size_t i = count; i > 0; i--) {
			if (i > t_ni_) {
				continue; // the reference would index before the vector here (undefined); nothing sensible to mirror
			}
			if ((size_t)t_ni_ + j < rope.size() && rope[t_ni_ + j].node_type != -1) {
				rope[t_ni_ - i] = rope[t_ni_ + j];
				rope[t_ni_ + j].node_type = -1;
				j++;
			} else {
				rope[t_ni_ - i].node_type = -1;
			}
		}
	}

	// After guard_remove the reference calls findAcceptedKmer (ntedit.cpp:848-903) from the stale tail cursor.  The slot
	// right behind the tail node is always dead at that point, so the search cannot collect k characters: it sets both
	// cursors to the contig length, and the next roll() ends the contig's main loop (ntedit.cpp:2134-2136).
	void guard_end_contig()
	{
		t_pos_ = len_;
		ended = true;
	}

	// case 2 of makeEdit, ntedit.cpp:1312-1393
	void edit_insertion(const Event& e)
	{
		const std::string ins(e.indel, e.indel + e.indel_len);
		std::string prev = prev_insertion();
		bool skipped = false;
		if (prev.size() + ins.size() >= k_) {
			if (is_repeat(prev) || prev.size() + ins.size() >= cap_) {
				guard_remove(prev.size());
				guard_end_contig();
				skipped = true;
			} else {
				for (size_t w = 0; w < ins.size(); w++) {
					prev.insert(prev.begin(), revcomp((unsigned char)ins[w]));
					if (is_repeat(prev)) {
						guard_remove(prev.size() - w);
						guard_end_contig();
						skipped = true;
					}
				}
			}
		}
		if (!skipped) {
			rope_insert(ins, e.support);
		}
	}

	// makeInsertion, ntedit.cpp:625-714
	void rope_insert(const std::string& bases, uint32_t support)
	{
		const ntb_node orig = rope[t_ni_];
		const size_t nb = bases.size();
		if (orig.node_type == 0 && t_pos_ > orig.s_pos) {
			ntb_node after;
			std::memset(&after, 0, sizeof after);
			after.node_type = 0;
			after.s_pos = t_pos_;
			after.e_pos = orig.e_pos;
			rope[t_ni_].e_pos = t_pos_ - 1;
			for (size_t i = 0; i < nb; i++) {
				put((size_t)t_ni_ + i + 1, char_node((unsigned char)bases[i], support));
			}
			put((size_t)t_ni_ + nb + 1, after);
			t_ni_++;
			return;
		}
		if (orig.node_type == 0 || orig.node_type == 1) {
			std::vector<ntb_node> lifted;
			size_t i = t_ni_;
			while (i < rope.size() && rope[i].node_type != -1) {
				lifted.push_back(rope[i]);
				rope[i].node_type = -1;
				i++;
			}
			for (size_t q = 0; q < nb; q++) {
				put((size_t)t_ni_ + q, char_node((unsigned char)bases[q], support));
			}
			for (size_t q = 0; q < lifted.size(); q++) {
				put((size_t)t_ni_ + nb + q, lifted[q]);
			}
		}
	}

	// makeDeletion, ntedit.cpp:719-809
	void rope_delete(uint32_t num_del, uint32_t support)
	{
		for (;;) {
			const ntb_node orig = rope[t_ni_];
			uint32_t leftover = 0;
			if (orig.node_type == 0) {
				if (t_pos_ <= orig.s_pos) {
					if (t_pos_ + num_del <= orig.e_pos) {
						rope[t_ni_].s_pos = t_pos_ + num_del;
						rope[t_ni_].num_support = support;
						t_pos_ = rope[t_ni_].s_pos;
						return;
					}
					leftover = t_pos_ + num_del - orig.e_pos;
					t_pos_ = orig.e_pos + 1;
					size_t i = (size_t)t_ni_ + 1;
					while (i < rope.size() && rope[i].node_type != -1) {
						rope[i - 1] = rope[i];
						rope[i].node_type = -1;
						i++;
					}
				} else {
					if (t_pos_ + num_del <= orig.e_pos) {
						ntb_node split;
						std::memset(&split, 0, sizeof split);
						split.node_type = 0;
						split.s_pos = t_pos_ + num_del;
						split.e_pos = orig.e_pos;
						split.num_support = support;
						rope[t_ni_].e_pos = t_pos_ - 1;
						t_pos_ = split.s_pos;
						t_ni_++;
						put(t_ni_, split);
						return;
					}
					leftover = t_pos_ + num_del - orig.e_pos;
					rope[t_ni_].e_pos = t_pos_ - 1;
					t_pos_ = orig.e_pos + 1;
					t_ni_++;
				}
			} else if (orig.node_type == 1) {
				size_t i = t_ni_;
				leftover = num_del;
				while (i < rope.size() && rope[i].node_type == 1 && leftover > 0) {
					rope[i].node_type = -1;
					leftover--;
					i++;
				}
				size_t j = t_ni_;
				while (i < rope.size() && rope[i].node_type != -1) {
					rope[j] = rope[i];
					rope[i].node_type = -1;
					i++;
					j++;
				}
			} else {
				return;
			}
			if (leftover > 0 && t_ni_ < rope.size() && rope[t_ni_].node_type != -1) {
				if (rope[t_ni_].node_type == 0) {
					t_pos_ = rope[t_ni_].s_pos;
				}
				num_del = leftover;
				continue;
			}
			return;
		}
	}
};

} // namespace ntb
