// TEST INFRASTRUCTURE ONLY -- never part of the product library.
//
// CPU simulator of the device side of ntedit_b200: it compiles the very same engine header the CUDA kernels
// instantiate (ntedit_b200/csrc/engine.h) for the host and plugs it into the product's stitch/replay driver
// (polish_driver.hpp).  It exists so the walker state machine, the segment stitcher and the rope replay can be
// fuzzed against the reference in a container without a GPU; the GPU tests then only have to show that the CUDA
// build of the same engine gives the same events.
#include "../../ntedit_b200/csrc/engine.h"
#include "../../ntedit_b200/csrc/site_dense.h"
#include "../../ntedit_b200/csrc/polish_driver.hpp"
#include "../../ntedit_b200/csrc/writer.hpp"

#include <cstdlib>

using namespace ntb;

namespace {

FilterView
make_view(const uint8_t* data, uint64_t bytes, uint32_t h, int counting)
{
	FilterView v;
	std::memset(&v, 0, sizeof v);
	v.data = data;
	v.bytes = bytes;
	v.mod = counting ? bytes : bytes * 8;
	v.recip = v.mod ? 0xFFFFFFFFFFFFFFFFULL / v.mod : 0;
	v.mask = (v.mod && (v.mod & (v.mod - 1)) == 0 && v.mod > 1) ? v.mod - 1 : 0;
	v.hash_num = h;
	v.counting = counting ? 1u : 0u;
	return v;
}

struct HostBackend
{
	const unsigned char* bases;
	uint64_t total;
	FilterView bloom, rep;
	std::vector<uint32_t> visit;
	std::string err;

	const std::string& error() const { return err; }

	// CPU stand-in for the scan kernel K1
	void scan_end() {}
	void scan_until(uint64_t) {}    // (scan_begin scans the whole batch at once)
	void scan_prefetch(uint64_t) {}
	// (HOSTSIM_STREAMING=1: pretend the text is still being uploaded, which selects the growing contig groups)
	bool text_streaming() const { return std::getenv("HOSTSIM_STREAMING") != nullptr; }

	void scan_begin(const KParams& kp)
	{
		visit.assign((total + 63) / 32 + 2, 0u);
		uint64_t run = 0;
		HashState hs;
		hs.fh = hs.rh = 0;
		for (uint64_t g = 0; g < total; g++) {
			const unsigned char c = bases[g];
			run = is_accepted_any_case(c) ? run + 1 : 0;
			if (run < kp.k) {
				continue;
			}
			const unsigned char* w = bases + g + 1 - kp.k;
			hash_seed(hs, kp.k, [w](unsigned i) { return w[i]; });
			bool site;
			if (kp.snv) {
				site = true;
			} else if (kp.counting) {
				const unsigned cnt = filter_count(bloom, hash_canonical(hs), kp.k);
				site = cnt == 0 || cnt < kp.min_threshold;
			} else {
				site = !filter_contains(bloom, hash_canonical(hs), kp.k);
			}
			if (site) {
				visit[g >> 5] |= 1u << (g & 31);
			}
		}
	}

	std::vector<Task> tasks;
	std::vector<TaskResult> results;
	std::vector<std::unique_ptr<std::vector<Event>>> rounds;

	const Event* round_events(size_t r) const { return rounds[r]->data(); }

	Task* task_buffer(size_t n)
	{
		tasks.resize(n);
		return tasks.data();
	}

	std::vector<SiteRec> table, table2;
	std::string dense_mismatch;
	uint64_t pre_records = 0, pre_pending = 0, pre_dropped = 0, skipped = 0;

	template<class W>
	void presites_with(const KParams& kp, size_t n_tasks, const std::vector<uint64_t>& rot)
	{
		// HOSTSIM_TABLE_SLOTS: tiny tables exercise dropped records (the walkers then evaluate those sites themselves)
		size_t slots = 1;
		const char* ts = std::getenv("HOSTSIM_TABLE_SLOTS");
		const size_t want = ts ? (size_t)std::strtoull(ts, nullptr, 10) : std::max<size_t>(1024, (size_t)(total / 16));
		while (slots < want) {
			slots <<= 1;
		}
		if (table.size() != slots) {
			// (the contig groups of a call share the table)
			table.assign(slots, SiteRec());
			std::memset(table.data(), 0, slots * sizeof(SiteRec));
		}
		std::vector<PendingSite> pending(ts ? 8 : (size_t)(total / 32 + 64));
		Counters ctr = {}, ctr2 = {};
		WalkerState<352>* st = new WalkerState<352>();
		W w(*st, kp);
		const char* dv = std::getenv("HOSTSIM_DENSE");
		const bool dense = !(dv && dv[0] == '0');
		const bool check = dense && !ts && std::getenv("HOSTSIM_CHECK_DENSE") != nullptr; // (a tiny table drops different records)
		std::vector<PendingSite> pending2(check ? pending.size() : 0);
		if (check && table2.size() != slots) {
			table2.assign(slots, SiteRec());
			std::memset(table2.data(), 0, slots * sizeof(SiteRec));
		}
		// text range of this group's tasks: the records its first pass files have their keys in there
		uint64_t key_lo = ~0ULL, key_hi = 0;
		for (size_t u = 0; u < n_tasks; u++) {
			key_lo = std::min<uint64_t>(key_lo, tasks[u].text_off + tasks[u].start + 1);
			key_hi = std::max<uint64_t>(key_hi, tasks[u].text_off + tasks[u].len + 1);
		}
		uint8_t cls_tab[256];
		for (unsigned c = 0; c < 256; c++) {
			cls_tab[c] = (uint8_t)(base_code((unsigned char)c) | (rev_code((unsigned char)c) << 3) | (is_accepted_any_case((unsigned char)c) ? 0x40u : 0u));
		}
		DenseCtx dctx;
		dctx.kp = &kp;
		dctx.bloom = bloom;
		dctx.rep = rep;
		dctx.rot = rot.data();
		dctx.cls = cls_tab;
		for (int pass = 0; pass < 2; pass++) {
			if (pass == 1 && check) {
				// the two forms of the first pass must have produced the same records and the same pending sites
				if (ctr.n_pending != ctr2.n_pending || ctr.n_dropped != ctr2.n_dropped) {
					dense_mismatch = "pending / dropped counts differ";
				}
				auto find = [&](const std::vector<SiteRec>& tab, uint64_t key) -> const SiteRec* {
					uint32_t slot = site_hash(key) & ((uint32_t)slots - 1);
					for (uint32_t i = 0; i < SITE_TABLE_PROBES; i++, slot = (slot + 1) & ((uint32_t)slots - 1)) {
						if (tab[slot].key == key) {
							return &tab[slot];
						}
						if (tab[slot].key == 0) {
							break;
						}
					}
					return nullptr;
				};
				// (slot positions may differ: the second passes of earlier groups only filled `table`)
				for (int dir = 0; dir < 2 && dense_mismatch.empty(); dir++) {
					const std::vector<SiteRec>& from = dir ? table2 : table;
					const std::vector<SiteRec>& to = dir ? table : table2;
					for (size_t q = 0; q < slots && dense_mismatch.empty(); q++) {
						const uint64_t key = from[q].key;
						if (key == 0 || key < key_lo || key > key_hi) {
							continue; // empty, or an earlier group's record (its second pass has completed it since)
						}
						const SiteRec* other = find(to, key);
						// (only the dense form's chain rounds know how far the walker may jump: compare without that)
						auto plain = [](SiteRec r) {
							if (r.flags & SITE_FL_SKIP) {
								r.flags &= (uint8_t)~SITE_FL_SKIP;
								r.indel_len = 0;
								std::memset(r.indel, 0, sizeof r.indel);
								r.pad_[0] = 0;
							}
							return r;
						};
						const SiteRec fa = plain(from[q]), fb = other ? plain(*other) : SiteRec();
						if (!other || std::memcmp(&fa, &fb, sizeof(SiteRec)) != 0) {
							char buf[200];
							std::snprintf(buf, sizeof buf, "record of text position %llu differs (%s state %u type %u, other form %s)",
							              (unsigned long long)key - 1, dir ? "walker" : "dense", from[q].state, from[q].best_type,
							              other ? "has another record" : "has none");
							dense_mismatch = buf;
						}
					}
				}
			}
			const size_t n_units = pass == 0 ? n_tasks : std::min<size_t>(ctr.n_pending, pending.size());
			for (size_t u = 0; u < n_units; u++) {
				const size_t ti = pass == 0 ? u : pending[u].task;
				const Task& t = tasks[ti];
				WalkerIO& io = st->io;
				io.text = bases + t.text_off;
				io.len = t.len;
				io.visit = visit.data();
				io.goff = t.text_off;
				io.bloom = bloom;
				io.rep = rep;
				io.events = nullptr;
				io.ev_cap = 0;
				io.ctr = &ctr;
				io.rot = rot.data();
				io.table = table.data();
				io.table_mask = (uint32_t)slots - 1;
				io.pending = pending.data();
				io.pending_cap = (uint32_t)pending.size();
				w.pre_begin();
				if (pass == 1) {
					w.pre_run(pending[u].task, pending[u].pos, true);
					continue;
				}
				// heads among the flagged positions of [start, end): the dense (thread-per-site) form of the first pass, as the
				// product runs it; HOSTSIM_DENSE=0 takes the walker form, HOSTSIM_CHECK_DENSE=1 runs both and compares every record
				for (uint64_t p = t.start; p < t.end; p++) {
					const uint64_t g = t.text_off + p;
					const bool f = (visit[g >> 5] >> (g & 31)) & 1u;
					if (f && W::is_head(visit.data(), t.text_off, (uint32_t)p, w.pre_gap())) {
						pre_records++;
						if (dense) {
							// as the device runs it: the head in round 0, then rounds of DENSE_GROUP chain sites evaluated side by side
							// (no-edit records learn how far the walker may jump: SITE_FL_SKIP)
							uint32_t q = (uint32_t)p;
							const uint32_t nx = dense_step<(int)KMAX>(dctx, io.text, io.len, io.goff, visit.data(), (uint32_t)ti, q, table.data(),
							                                        (uint32_t)slots - 1, pending.data(), (uint32_t)pending.size(), &ctr);
							q = nx != NONE32 ? q : NONE32; // the chain goes on BEHIND the head
							for (uint32_t round = 0; round < SITE_CHAIN_MAX / DENSE_GROUP && q != NONE32; round++) {
								q = dense_chain_round_host<(int)KMAX>(dctx, io.text, io.len, io.goff, visit.data(), (uint32_t)ti, q, table.data(),
								                                      (uint32_t)slots - 1, pending.data(), (uint32_t)pending.size(), &ctr);
							}
						}
						if (!dense || check) {
							if (check) {
								io.table = table2.data();
								io.pending = pending2.data();
								io.ctr = &ctr2;
							}
							w.pre_run((uint32_t)ti, (uint32_t)p, false);
						}
					}
				}
			}
		}
		pre_pending = ctr.n_pending;
		pre_dropped = ctr.n_dropped;
		delete st;
	}

	// K3 (-s 1), as the CUDA backend runs it (kernels.cu: snv_dense_kernel): every valid position is a site; the dense form
	// evaluates them all from the text, files a record for those that do something and marks them in a second bitmap -- the
	// one the walkers then jump through.  HOSTSIM_CHECK_DENSE=1 compares every position's record with the walker's own
	// evaluation.
	std::vector<uint32_t> visit2;

	void snv_presites(const KParams& kp, size_t n_tasks, const std::vector<uint64_t>& rot)
	{
		const char* ts = std::getenv("HOSTSIM_TABLE_SLOTS");
		size_t slots = 1;
		const size_t want = ts ? (size_t)std::strtoull(ts, nullptr, 10) : std::max<size_t>(1024, (size_t)(total / 16));
		while (slots < want) {
			slots <<= 1;
		}
		if (table.size() != slots) {
			table.assign(slots, SiteRec());
			std::memset(table.data(), 0, slots * sizeof(SiteRec));
		}
		if (visit2.size() != visit.size()) {
			visit2.assign(visit.size(), 0u);
		}
		const bool check = std::getenv("HOSTSIM_CHECK_DENSE") != nullptr;
		uint8_t cls_tab[256];
		for (unsigned c = 0; c < 256; c++) {
			cls_tab[c] = (uint8_t)(base_code((unsigned char)c) | (rev_code((unsigned char)c) << 3) | (is_accepted_any_case((unsigned char)c) ? 0x40u : 0u));
		}
		DenseCtx dctx;
		dctx.kp = &kp;
		dctx.bloom = bloom;
		dctx.rep = rep;
		dctx.rot = rot.data();
		dctx.cls = cls_tab;
		Counters ctr = {};
		WalkerState<352>* st = check ? new WalkerState<352>() : nullptr;
		for (size_t ti = 0; ti < n_tasks; ti++) {
			const Task& t = tasks[ti];
			for (uint64_t p = t.start; p < t.end; p++) {
				const uint64_t g = t.text_off + p;
				if (!((visit[g >> 5] >> (g & 31)) & 1u)) {
					continue;
				}
				SiteRec r;
				const uint32_t state = dense_site<(int)KMAX>(dctx, bases + t.text_off, t.len, (uint32_t)p, r);
				if (dense_has_effect(state, r, kp.mask != 0)) {
					visit2[g >> 5] |= 1u << (g & 31);
					dense_commit(r, state, t.text_off, (uint32_t)ti, (uint32_t)p, table.data(), (uint32_t)slots - 1, nullptr, 0, &ctr);
				}
				if (check && dense_mismatch.empty()) {
					WalkerIO& io = st->io;
					io.text = bases + t.text_off;
					io.len = t.len;
					io.visit = visit.data();
					io.goff = t.text_off;
					io.bloom = bloom;
					io.rep = rep;
					io.events = nullptr;
					io.ev_cap = 0;
					io.ctr = &ctr;
					io.rot = rot.data();
					io.table = nullptr;
					io.table_mask = 0;
					io.pending = nullptr;
					io.pending_cap = 0;
					Walker<352, false, false> w(*st, kp);
					w.pre_begin();
					w.pre_seed((uint32_t)p);
					const uint32_t wstate = w.evaluate_site_core(true);
					SiteRec wr;
					w.pre_fill(wr, wstate);
					wr.key = r.key = 0;
					if (wstate != state || std::memcmp(&wr, &r, sizeof(SiteRec)) != 0) {
						char buf[200];
						std::snprintf(buf, sizeof buf, "text position %llu: dense state %u type %u sub %u supp %u alt %u, walker state %u type %u sub %u supp %u alt %u",
						              (unsigned long long)g, state, r.best_type, r.best_sub, r.support, r.altsupp[0], wstate, wr.best_type, wr.best_sub,
						              wr.support, wr.altsupp[0]);
						dense_mismatch = buf;
					}
				}
			}
		}
		delete st;
		pre_dropped += ctr.n_dropped;
	}

	void presites(const KParams& kp, size_t n_tasks, const std::vector<uint64_t>& rot)
	{
		if (!kp.counting && !kp.h_rep && !kp.snv && !kp.mask && bloom.mask != 0) {
			presites_with<Walker<352, true, true>>(kp, n_tasks, rot);
		} else if (!kp.counting && !kp.h_rep && !kp.snv && !kp.mask) {
			presites_with<Walker<352, true, false>>(kp, n_tasks, rot);
		} else {
			presites_with<Walker<352, false, false>>(kp, n_tasks, rot);
		}
	}

	int walk(const KParams& kp, size_t n_tasks, bool first_round_of_group, const TaskResult** res_out, const Event** ev_out, size_t* n_ev_out)
	{
		results.resize(n_tasks);
		rounds.emplace_back(new std::vector<Event>(1u << 16));
		std::vector<Event>& events = *rounds.back();
		std::vector<uint64_t> rot(ROT_WORDS);
		for (uint32_t q = 0; q < ROT_WORDS; q++) {
			rot[q] = rot_entry(q);
		}
		// the pre-evaluation pass of the first round, as the CUDA backend runs it (capi.cu: CudaBackend::presites): heads of
		// flagged runs per task, first pass (no tryIndels), second pass (the pending ones); HOSTSIM_NO_PRESITE=1 turns it off
		if (first_round_of_group && kp.snv && !std::getenv("HOSTSIM_NO_PRESITE")) {
			snv_presites(kp, n_tasks, rot);
			if (!dense_mismatch.empty()) {
				err = "dense -s 1 pass != walker: " + dense_mismatch;
				return NTB_EINTERNAL;
			}
		}
		if (first_round_of_group && !kp.snv && !std::getenv("HOSTSIM_NO_PRESITE")) {
			presites(kp, n_tasks, rot);
			if (!dense_mismatch.empty()) {
				err = "dense first pass != walker first pass: " + dense_mismatch;
				return NTB_EINTERNAL;
			}
		}
		for (;;) {
			Counters ctr = {};
			WalkerState<352>* st = new WalkerState<352>();
			for (size_t i = 0; i < n_tasks; i++) {
				WalkerIO& io = st->io;
				io.table = table.empty() ? nullptr : table.data();
				io.table_mask = table.empty() ? 0 : (uint32_t)table.size() - 1;
				io.pending = nullptr;
				io.pending_cap = 0;
				io.text = bases + tasks[i].text_off;
				io.len = tasks[i].len;
				io.visit = visit2.empty() ? visit.data() : visit2.data(); // -s 1: only the sites that do something (snv_presites)
				io.goff = tasks[i].text_off;
				io.bloom = bloom;
				io.rep = rep;
				io.events = events.data();
				io.ev_cap = (uint32_t)events.size();
				io.ctr = &ctr;
				io.rot = rot.data();
				// the same two instantiations the device kernel dispatches between
				if (!kp.counting && !kp.h_rep && !kp.snv && !kp.mask && bloom.mask != 0) {
					Walker<352, true, true> w(*st, kp);
					w.run(tasks[i], results[i]);
				} else if (!kp.counting && !kp.h_rep && !kp.snv && !kp.mask) {
					Walker<352, true, false> w(*st, kp);
					w.run(tasks[i], results[i]);
				} else {
					Walker<352, false, false> w(*st, kp);
					w.run(tasks[i], results[i]);
				}
			}
			delete st;
			skipped += ctr.n_skipped;
			if (std::getenv("HOSTSIM_DEBUG")) {
				std::fprintf(stderr, "[hostsim] round %zu: %zu tasks, run heads %llu, pending %llu, dropped %llu, sites from records %u (%u of the second pass)\n", rounds.size(),
				             n_tasks, (unsigned long long)pre_records, (unsigned long long)pre_pending, (unsigned long long)pre_dropped, ctr.n_rec_used, ctr.n_rec_used2);
			}
			if (!ctr.overflow) {
				events.resize(ctr.n_events);
				// Backend contract: last_event = index of the walker's first event, its events contiguous and in order.
				// The tasks ran one after the other, so every chain already is a contiguous ascending run.
				for (size_t i = 0; i < n_tasks; i++) {
					results[i].last_event = results[i].n_events ? results[i].last_event + 1 - results[i].n_events : NONE32;
				}
				*res_out = results.data();
				*ev_out = events.data();
				*n_ev_out = events.size();
				return NTB_OK;
			}
			events.assign(events.size() * 4, Event());
		}
	}
};

char*
dup_out(const std::string& s, size_t* n)
{
	char* p = (char*)std::malloc(s.size() + 1);
	std::memcpy(p, s.data(), s.size());
	p[s.size()] = 0;
	*n = s.size();
	return p;
}

} // namespace

extern "C" {

// polishes a batch on the CPU simulator and formats the three outputs (TSV with its header, VCF rows without header)
int
hostsim_polish(const uint8_t* filt, uint64_t fbytes, uint32_t k, uint32_t h, int counting, const uint8_t* rep, uint64_t rbytes,
               uint32_t rh, int rcounting, const ntb_params* up, char* bases, const uint64_t* offsets, uint64_t n_contigs,
               const char* const* headers, char** fa, size_t* fa_len, char** tsv, size_t* tsv_len, char** vcf, size_t* vcf_len,
               ntb_stats* stats, char* errbuf, size_t errlen)
{
	KParams kp;
	std::string err;
	int rc = make_kparams(*up, k, h, rep ? rh : 0, counting != 0, kp, err);
	if (rc == NTB_OK) {
		HostBackend be;
		be.bases = (const unsigned char*)bases;
		be.total = offsets[n_contigs];
		be.bloom = make_view(filt, fbytes, h, counting);
		be.rep = rep ? make_view(rep, rbytes, rh, rcounting) : make_view(nullptr, 0, 0, 0);
		// the walkers must see the ORIGINAL draft while the replay mutates the caller's buffer
		std::vector<unsigned char> pristine(be.bases, be.bases + be.total);
		be.bases = pristine.data();
		ResultImpl res;
		rc = polish_run(be, kp, *up, bases, offsets, n_contigs, res, err);
		if (rc == NTB_OK) {
			std::string sfa, stsv = tsv_header(k, up->jump, counting != 0), svcf;
			for (uint64_t c = 0; c < n_contigs; c++) {
				const ContigResult& cr = res.contigs[c];
				if (!cr.polished) {
					continue;
				}
				format_contig(headers[c], bases + offsets[c], cr.nodes.data(), cr.nodes.size(), cr.srecs.data(), cr.srecs.size(),
				              up->snv != 0, nullptr, &sfa, &stsv, &svcf);
			}
			*fa = dup_out(sfa, fa_len);
			*tsv = dup_out(stsv, tsv_len);
			*vcf = dup_out(svcf, vcf_len);
			if (stats) {
				*stats = res.stats;
				stats->pad_ = (uint32_t)be.skipped; // (test infrastructure: no-edit chain sites the walkers jumped over)
			}
		}
	}
	if (rc != NTB_OK && errbuf && errlen) {
		std::snprintf(errbuf, errlen, "%s", err.c_str());
	}
	return rc;
}

void
hostsim_free(char* p)
{
	std::free(p);
}
}
