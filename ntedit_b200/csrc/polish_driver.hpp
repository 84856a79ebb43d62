// Batch orchestration of the polishing path, independent of where the kernels run: cut contigs into segments, launch
// the scan (K1) and the walkers (K2) through a Backend, stitch the per-segment results into the order the
// reference's strictly sequential loop (ntedit.cpp:1797-2139) would have produced, re-launch the few segments whose
// speculative clean start turned out to be wrong, then replay the accepted events into ropes (replay.hpp).
//
// Backend concept:
//   void  scan_begin(const KParams&);                      -- K1: get ready to build the visit bitmap (nothing is scanned yet)
//   void  scan_until(uint64_t p_end);                      -- ... start scanning every text position below p_end that has not been
//            started yet (asynchronous; walk() sees the bits of whatever was started before it)
//   void  scan_prefetch(uint64_t p_end);                   -- ... do the same right behind the next walk()'s own device work
//   void  scan_end();                                      -- ... wait for everything that was started (timing only)
//   bool  text_streaming() const;                          -- the text is still being copied to the device while the call runs
//   Task* task_buffer(size_t n);                           -- host buffer (pinned in the CUDA backend) for n tasks
//   int   walk(const KParams&, size_t n_tasks, bool first_round_of_group, const TaskResult** results, const Event** events,
//              size_t* n_events);
//         -- K2 over the tasks in task_buffer(); results stay valid until the next walk().  The events of task i are
//            events[results[i].last_event .. + results[i].n_events), in the order the walker emitted them.  In a group's
//            first round the tasks cover every contig of the group (the pre-evaluation passes run in front of it).
//   const Event* round_events(size_t r) const;             -- the events of round r (the r-th walk() call); they stay where
//            they are until the backend is destroyed, also while later rounds run (the replay of one contig group reads
//            them while the device walks the next group)
// The product instantiates this with the CUDA backend (capi.cu); tests/hostsim instantiates it with a CPU
// simulator of the same engine so the stitch/replay logic can be fuzzed without a GPU.
#pragma once
#include "replay.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>

namespace ntb {

struct ContigResult
{
	bool polished = false;
	std::vector<ntb_node> nodes;
	std::vector<ntb_srec> srecs;
};

struct ResultImpl
{
	std::vector<ContigResult> contigs;
	ntb_stats stats;
};

// smallest integer count c such that (float)c >= x, as the reference compares (e.g. ntedit.cpp:1659-1663)
inline uint32_t
threshold_from_float(float x)
{
	if (std::isnan(x)) {
		return NONE32;
	}
	if (x <= 0.0f) {
		return 0;
	}
	if (x > 1.0e9f) {
		return NONE32;
	}
	return (uint32_t)std::ceil(x);
}

// Fill the kernel parameters from the user parameters with the reference's own float arithmetic.
inline int
make_kparams(const ntb_params& u, uint32_t k, uint32_t h, uint32_t h_rep, bool counting, KParams& kp, std::string& err)
{
	static const uint32_t num_tries[6] = { 0, 1, 5, 21, 85, 341 }; // ntedit.cpp:172
	std::memset(&kp, 0, sizeof kp);
	if (k < KMIN || k > KMAX) {
		err = "unsupported k (supported: " + std::to_string(KMIN) + ".." + std::to_string(KMAX) + ")";
		return NTB_EINVAL;
	}
	if (h == 0 || h > HMAX || h_rep > HMAX) {
		err = "unsupported hash_num";
		return NTB_EINVAL;
	}
	if (u.jump == 0) {
		err = "jump (-j) must be positive";
		return NTB_EINVAL;
	}
	if (u.mode < 0 || u.mode > 2) {
		err = "mode (-m) must be 0, 1 or 2";
		return NTB_EINVAL;
	}
	uint32_t max_ins = u.max_insertions, max_del = u.max_deletions;
	if (u.snv) { // ntedit.cpp:2411-2413
		max_ins = 0;
		max_del = 0;
	}
	if ((max_ins == 0 && max_del > 0) || (max_ins == 1 && max_del > 1)) { // ntedit.cpp:2478-2483
		max_del = max_ins;
	}
	if (max_ins > 5) { // ntedit.cpp:2485-2493
		max_ins = 5;
	}
	if (max_del > 10) {
		max_del = 10;
	}
	kp.k = k;
	kp.h = h;
	kp.h_rep = h_rep;
	kp.jump = u.jump;
	kp.mode = u.mode;
	kp.snv = u.snv ? 1 : 0;
	kp.mask = u.mask ? 1 : 0;
	kp.max_ins_tries = num_tries[max_ins];
	kp.max_deletions = max_del;
	kp.counting = counting ? 1 : 0;
	kp.min_threshold = (!counting && u.min_threshold != 1) ? 1 : u.min_threshold; // ntedit.cpp:2453-2458
	kp.max_threshold = u.max_threshold;
	kp.insertion_cap = (uint32_t)((float)k * 1.5f); // ntedit.cpp:2450-2451 (overrides any -c)
	const float fk = (float)k;
	if (!u.use_ratio) {
		kp.thr_missing = threshold_from_float(fk / u.missing_threshold);
		kp.thr_edit = threshold_from_float(fk / u.edit_threshold);
		kp.thr_edit_del = kp.thr_edit;
	} else {
		const float per_jump = fk / (float)u.jump;
		kp.thr_missing = threshold_from_float(per_jump * u.missing_ratio);
		kp.thr_edit = threshold_from_float(per_jump * u.edit_ratio);
		kp.thr_edit_del = threshold_from_float((1 + per_jump) * u.edit_ratio); // ntedit.cpp:1533-1535
	}
	const uint64_t seeds[4] = { SEED_A, SEED_C, SEED_G, SEED_T };
	for (int c = 0; c < 4; c++) {
		kp.seed_rot_k[c] = sroln(seeds[c], k);
		kp.seed_rot_k1[c] = sroln(seeds[c], k - 1);
	}
	return NTB_OK;
}

struct Segment
{
	uint32_t contig;
	uint32_t p0, p1;      // nominal range of tail positions
	uint32_t run_start;   // start of the run whose result is stored
	int32_t arena;        // which round's event arena holds the events (-1: no result yet)
	TaskResult res;
};

template<class Backend>
int
polish_run(Backend& be, const KParams& kp_in, const ntb_params& up, char* host_bases, const uint64_t* offsets, uint64_t n_contigs,
           ResultImpl& out, std::string& err)
{
	using clk = std::chrono::steady_clock;
	const bool dbg = std::getenv("NTB_DEBUG_TASKS") != nullptr;
	const auto t_begin = clk::now();
	auto since = [&](clk::time_point t) { return std::chrono::duration<double, std::milli>(clk::now() - t).count(); };
	out.contigs.assign(n_contigs, ContigResult());
	std::memset(&out.stats, 0, sizeof out.stats);
	KParams kp = kp_in;
	// Segment length: longer segments mean fewer tasks (less per-task work on the device, less stitching and fewer replay
	// pieces on the host: 4096 -> 16384 takes 141 -> 137 ms off the walker and 18 -> 10 ms off the host at 3 Gbp), shorter
	// ones keep every walker slot of the device busy on small batches: aim at ~96 K tasks.
	uint32_t seg_len = up.segment_len;
	if (seg_len == 0) {
		if (kp.snv) {
			seg_len = 1024u; // every position is a site
		} else {
			const uint64_t total = n_contigs ? offsets[n_contigs] : 0;
			seg_len = (uint32_t)std::min<uint64_t>(16384, std::max<uint64_t>(4096, total / 98304));
		}
	}
	if (seg_len < 4 * kp.k) {
		seg_len = 4 * kp.k;
	}

	// first-round walkers may move their segment borders out of runs of flagged positions (engine.h: safe_boundary)
	kp.boundary_lim = std::getenv("NTB_NO_BORDER_ADJUST") ? 0u : std::min<uint32_t>(512u, seg_len / 2);

	for (uint64_t c = 0; c < n_contigs; c++) {
		if (offsets[c + 1] - offsets[c] - 1 >= 0xFFFFFFFEULL) {
			err = "contig longer than 2^32-2 bases";
			return NTB_EINVAL;
		}
	}
	be.scan_begin(kp);

	// host thread pool helper: fn(i) for i in [0, n_items), dynamically scheduled
	// host threads of this call: NTB_HOST_THREADS, else the cores of the box divided by the ranks that share it (torchrun
	// exports LOCAL_WORLD_SIZE) -- every rank replays its own batch at the same time
	unsigned nthreads = std::thread::hardware_concurrency();
	if (nthreads == 0) {
		nthreads = 4;
	}
	if (const char* v = std::getenv("NTB_HOST_THREADS")) {
		const long n = std::strtol(v, nullptr, 10);
		if (n > 0) {
			nthreads = (unsigned)n;
		}
	} else if (const char* w = std::getenv("LOCAL_WORLD_SIZE")) {
		const long n = std::strtol(w, nullptr, 10);
		if (n > 1) {
			nthreads = std::max<unsigned>(2, nthreads / (unsigned)n);
		}
	}
	auto run_parallel = [&](uint64_t n_items, const std::function<void(uint64_t)>& fn) {
		std::atomic<uint64_t> next(0);
		const unsigned nt = (unsigned)std::min<uint64_t>(nthreads, std::max<uint64_t>(1, n_items));
		// items are handed out in runs: with hundreds of thousands of tiny contigs a shared counter bumped once per item
		// costs more than the items themselves
		const uint64_t grain = std::max<uint64_t>(1, std::min<uint64_t>(1024, n_items / ((uint64_t)nt * 32)));
		auto worker = [&]() {
			for (;;) {
				const uint64_t i0 = next.fetch_add(grain);
				if (i0 >= n_items) {
					break;
				}
				const uint64_t i1 = std::min<uint64_t>(n_items, i0 + grain);
				for (uint64_t i = i0; i < i1; i++) {
					fn(i);
				}
			}
		};
		if (nt <= 1) {
			worker();
			return;
		}
		std::vector<std::thread> pool;
		for (unsigned i = 0; i < nt; i++) {
			pool.emplace_back(worker);
		}
		for (auto& th : pool) {
			th.join();
		}
	};

	// ---- contig groups.  Everything behind the scan is independent per contig, so the contigs are cut into a few groups
	// (consecutive contigs, about equal numbers of bases) and the groups are pipelined: while the device walks group g+1, a
	// host thread stitches nothing any more -- that happened in the device phase -- but replays group g's events into
	// ropes.  The host work of a call (a quarter of the device time with every core of the box, as much as the device
	// time when eight ranks share the cores) then hides behind the kernels except for the last group's.
	// How many, and how unequal: every further group costs device time (launch tails, 3 ms per group at 3 Gbp), and the replay
	// of group g has to fit beside the device phase of group g+1.  With a dozen host threads or more per GPU the replay is
	// a fifth of the device time: two groups, the second an eighth of the first (measured at 3 Gbp, 16 threads: 177 ms per
	// call with 4 equal groups, 174 with 4 at ratio 0.6, 170 with 3 at 0.4, 169 with 2 at 0.3; later, with everything else in
	// place, 162.4 / 160.3 / 160.0 ms resident and 169 / 166 / 168 ms end to end with 2 groups at 0.25 / 0.12 / 0.07).  With a few threads per GPU
	// (eight ranks sharing one host) the replay is as long as the device phase: four groups that shrink slowly.
	// A text that is still arriving from the host, on a host whose ranks share the PCIe uplinks (23 GB/s per GPU with eight
	// ranks against 55 alone -- barely faster than the device works through it): the device must not sit idle until 40 % of
	// the text is there, so the groups start small and grow (6, 10, 16, 24, 27, 17 % of the bases): 226 -> 220 ms per call
	// at 8 x 3 Gbp, where the upload itself takes 154 ms.
	const bool many_threads = nthreads >= 12;
	const bool grow = !many_threads && be.text_streaming() && !std::getenv("NTB_CONTIG_GROUPS") && !std::getenv("NTB_CONTIG_GROUP_RATIO");
	static const double grow_cum[] = { 0.06, 0.16, 0.32, 0.56, 0.83 };
	uint64_t n_groups = grow ? 6 : many_threads ? 2 : 4;
	if (const char* v = std::getenv("NTB_CONTIG_GROUPS")) {
		n_groups = std::max<uint64_t>(1, std::strtoull(v, nullptr, 10));
	}
	{
		const uint64_t total = n_contigs ? offsets[n_contigs] : 0;
		uint64_t min_positions = 64ull << 20; // groups of at least 64 M positions (NTB_CONTIG_GROUP_MIN: testing aid)
		if (const char* v = std::getenv("NTB_CONTIG_GROUP_MIN")) {
			min_positions = std::max<uint64_t>(1, std::strtoull(v, nullptr, 10));
		}
		n_groups = std::min<uint64_t>(n_groups, std::max<uint64_t>(1, total / min_positions));
	}
	std::vector<uint64_t> group_first(1, 0); // first contig of every group, then n_contigs
	{
		// group g+1 holds `ratio` times the bases of group g: the device launches of the large early groups run at their best,
		// and the replay nothing hides any more -- the last group's -- is small (NTB_CONTIG_GROUP_RATIO; 1 = equal groups)
		const uint64_t total = n_contigs ? offsets[n_contigs] : 0;
		// (a fragmented draft -- contigs of a few kbp, hundreds of thousands of them -- has four times the host work per base:
		// the second group has to be large enough to cover the first one's replay.  2.5 Gbp in 500 K contigs, 16 threads:
		// 323 ms per call at ratio 0.12, 289 at 0.33, 297 with 3 groups at 0.5, 312 with 4 at 0.7)
		const bool fragmented = n_contigs > total / 100000;
		double ratio = many_threads ? (fragmented ? 0.33 : 0.12) : 0.7;
		if (const char* v = std::getenv("NTB_CONTIG_GROUP_RATIO")) {
			ratio = std::min(1.0, std::max(0.05, std::strtod(v, nullptr)));
		}
		double wsum = 0, w = 1;
		for (uint64_t g = 0; g < n_groups; g++, w *= ratio) {
			wsum += w;
		}
		double acc = 0;
		w = 1;
		for (uint64_t g = 1; g < n_groups; g++) {
			acc += w;
			w *= ratio;
			const uint64_t want = (uint64_t)((double)total * (grow && n_groups == 6 ? grow_cum[g - 1] : acc / wsum));
			uint64_t c = std::upper_bound(offsets, offsets + n_contigs + 1, want) - offsets; // first contig that starts behind `want`
			c = std::min<uint64_t>(c, n_contigs);
			if (c > group_first.back()) {
				group_first.push_back(c);
			}
		}
		group_first.push_back(n_contigs);
	}
	// K1 of the first group runs while the host writes its tasks; the scan of group g+1 is enqueued behind the device work of
	// group g (scan_prefetch): a text that is still being uploaded arrives beside all of it
	auto group_end_pos = [&](size_t g) { return n_contigs ? offsets[group_first[g + 1]] : (uint64_t)0; };
	if (dbg) {
		std::fprintf(stderr, "[ntb] host: scan prepared at %.1f ms\n", since(t_begin));
	}
	be.scan_until(group_end_pos(0));
	if (dbg) {
		std::fprintf(stderr, "[ntb] host: first range enqueued at %.1f ms\n", since(t_begin));
	}

	// ---- segments
	std::vector<Segment> segs;
	std::vector<uint64_t> first_seg(n_contigs + 1, 0);
	for (uint64_t c = 0; c < n_contigs; c++) {
		first_seg[c] = segs.size();
		const uint64_t len64 = offsets[c + 1] - offsets[c] - 1;
		if (len64 >= 0xFFFFFFFEULL) {
			err = "contig longer than 2^32-2 bases";
			return NTB_EINVAL;
		}
		const uint32_t len = (uint32_t)len64;
		if (len < up.min_contig_len || len == 0) {
			continue; // dropped from all outputs, ntedit.cpp:2242-2245
		}
		out.contigs[c].polished = true;
		out.stats.bases += len;
		out.stats.contigs++;
		if (len < kp.k) {
			continue; // no k-mer: the rope stays the root node
		}
		for (uint32_t p = 0; p < len; p += seg_len) {
			Segment s;
			s.contig = (uint32_t)c;
			s.p0 = p;
			s.p1 = (len - p <= seg_len) ? len : p + seg_len;
			s.run_start = p;
			s.arena = -1;
			std::memset(&s.res, 0, sizeof s.res);
			segs.push_back(s);
			if (s.p1 == len) {
				break;
			}
		}
	}
	first_seg[n_contigs] = segs.size();

	if (dbg) {
		std::fprintf(stderr, "[ntb] host: %zu segments built in %.1f ms\n", segs.size(), since(t_begin));
	}

	size_t n_arenas = 0; // rounds walked so far (all groups); a round's events stay where the backend put them
	double host_ms = 0;  // stitch passes (device phases) + replays, summed over the groups
	std::mutex host_lock; // guards host_ms / n_edits_total / first_error between the replay thread and this one
	uint64_t n_edits_total = 0;
	std::string first_error;

	// ---- device phase of a group: rounds of walkers + stitching
	auto device_phase = [&](uint64_t c0, uint64_t c1) -> int {
		std::vector<uint64_t> pending; // segment indices to (re)run
		pending.reserve((size_t)(first_seg[c1] - first_seg[c0]));
		for (uint64_t i = first_seg[c0]; i < first_seg[c1]; i++) {
			pending.push_back(i);
		}
		uint32_t rounds = 0;
		while (!pending.empty()) {
			Task* tasks = be.task_buffer(pending.size());
			if (!tasks) {
				err = be.error();
				return NTB_ENOMEM;
			}
			const bool first_round = rounds == 0;
			{
				const uint64_t n_pending = pending.size();
				const uint64_t chunk = 16384;
				run_parallel((n_pending + chunk - 1) / chunk, [&](uint64_t q) {
					const uint64_t e = std::min<uint64_t>(n_pending, (q + 1) * chunk);
					for (uint64_t i = q * chunk; i < e; i++) {
						const Segment& sg = segs[pending[i]];
						Task& t = tasks[i];
						t.text_off = offsets[sg.contig];
						t.len = (uint32_t)(offsets[sg.contig + 1] - offsets[sg.contig] - 1);
						t.start = sg.run_start;
						t.end = sg.p1;
						t.contig = sg.contig;
						t.flags = (sg.p0 == 0 && sg.run_start == 0) ? TASK_CONTIG_START : 0;
						if (first_round) {
							// nominal borders: both neighbours move them by the same rule
							t.flags |= (sg.p0 > 0 ? TASK_ADJUST_START : 0u) | (sg.p1 < t.len ? TASK_ADJUST_END : 0u);
						}
						t.pad_ = 0;
					}
				});
			}
			const TaskResult* results = nullptr;
			const Event* round_events = nullptr;
			size_t n_round_events = 0;
			const auto t_walk = clk::now();
			const int rc = be.walk(kp, pending.size(), first_round, &results, &round_events, &n_round_events);
			if (dbg) {
				std::fprintf(stderr, "[ntb] host: walk round (%zu tasks, %zu events) returned after %.1f ms (at %.1f ms)\n", pending.size(),
				             n_round_events, since(t_walk), since(t_begin));
			}
			if (rc != NTB_OK) {
				err = be.error();
				return rc;
			}
			n_arenas++;
			rounds++;
			out.stats.rounds = std::max(out.stats.rounds, rounds);
			out.stats.segments += pending.size();
			if (rounds > 1) {
				out.stats.reruns += pending.size();
			}
			const auto t0 = clk::now();
			{
				// take the results over (chunks of tasks in parallel)
				const uint64_t n_pending = pending.size();
				const uint64_t chunk = 16384;
				const uint64_t n_chunks = (n_pending + chunk - 1) / chunk;
				std::vector<uint64_t> chunk_sites(n_chunks, 0);
				std::atomic<int> bad(0);
				const int32_t arena_idx = (int32_t)n_arenas - 1;
				run_parallel(n_chunks, [&](uint64_t q) {
					uint64_t sites = 0;
					const uint64_t e = std::min<uint64_t>(n_pending, (q + 1) * chunk);
					for (uint64_t i = q * chunk; i < e; i++) {
						Segment& sg = segs[pending[i]];
						sg.res = results[i];
						sg.arena = arena_idx;
						if (first_round && kp.boundary_lim && sg.p0 > 0) {
							// the border this walker (and its predecessor) actually used; a group's first round holds every
							// segment of its contigs, in order
							sg.p0 = sg.run_start = sg.res.start_pos;
							segs[pending[i] - 1].p1 = sg.res.start_pos;
						}
						if (sg.res.status & ST_ROPE_OVERFLOW) {
							bad = 1;
						} else if (!(sg.res.status & ST_DONE) || (sg.res.status & ST_EV_OVERFLOW)) {
							bad = 2;
						}
						sites += sg.res.n_sites;
					}
					chunk_sites[q] = sites;
				});
				if (bad == 1) {
					err = "device rope capacity exceeded (pathological insertion run); input not supported";
					return NTB_EINTERNAL;
				}
				if (bad == 2) {
					err = "device walker did not finish";
					return NTB_EINTERNAL;
				}
				for (uint64_t q = 0; q < n_chunks; q++) {
					out.stats.sites += chunk_sites[q];
				}
			}
			// stitch pass, contigs in parallel: accept results in contig order; where a predecessor ran past a successor's first
			// site, re-run that successor from the predecessor's clean end (optimistically assuming the re-run will end on its
			// own border)
			std::vector<std::vector<uint64_t>> redo(c1 - c0);
			run_parallel(c1 - c0, [&](uint64_t q) {
				const uint64_t c = c0 + q;
				uint32_t prev_end = 0;
				for (uint64_t i = first_seg[c]; i < first_seg[c + 1]; i++) {
					Segment& sg = segs[i];
					const uint32_t need = std::max(prev_end, sg.p0);
					if (need >= sg.p1) {
						continue; // entirely covered by the predecessor's overrun
					}
					const uint32_t ft = sg.res.first_touch != NONE32 ? std::min(sg.res.first_touch, sg.res.end_pos) : sg.res.end_pos;
					const bool valid = sg.arena >= 0 && sg.run_start <= need && need <= ft;
					if (valid) {
						prev_end = sg.res.end_pos;
						if (sg.res.status & ST_CONTIG_END) {
							break;
						}
					} else {
						sg.run_start = need;
						sg.arena = -1;
						redo[q].push_back(i);
						prev_end = sg.p1;
					}
				}
			});
			pending.clear();
			for (uint64_t q = 0; q < c1 - c0; q++) {
				pending.insert(pending.end(), redo[q].begin(), redo[q].end());
			}
			{
				std::lock_guard<std::mutex> guard(host_lock);
				host_ms += std::chrono::duration<double, std::milli>(clk::now() - t0).count();
			}
			if (dbg) {
				std::fprintf(stderr, "[ntb] host:   take-over + stitch pass %.1f ms, %zu to re-run (at %.1f ms)\n", since(t0), pending.size(),
				             since(t_begin));
			}
			if (rounds > 64) {
				err = "stitcher did not converge";
				return NTB_EINTERNAL;
			}
		}
		return NTB_OK;
	};

	// ---- host phase of a group: replay the accepted events into ropes.
	// Every accepted walker result starts from a clean window ("anchored": k unedited bases on the rope's final position
	// node), so the rope a contig ends up with is the concatenation of ropes replayed independently from fresh roots, cut
	// anywhere between two accepted results: the cut only splits the position node that spans it.  Long contigs are
	// therefore replayed as several PIECES in parallel (the largest human-like contig would otherwise be the critical path).
	auto host_phase = [&](uint64_t c0, uint64_t c1, std::vector<const Event*> arena_base) {
		const auto t1 = clk::now();
		struct Piece
		{
			uint32_t contig;
			uint64_t a0, a1;      // range inside `acc`
			uint8_t stale[4];     // reference's stale site locals at the start of the piece (see STALE_REF)
			std::vector<ntb_node> nodes;
			std::vector<ntb_srec> recs;
			bool ended = false;
			uint64_t edits = 0;
			std::string error;
		};
		// (A) per contig: the accepted results in order, the stale bytes each one starts with, and the cuts.  Flat arrays
		// indexed like the group's part of `segs` (contig c owns [first_seg[c], first_seg[c+1]) - seg0): no per-contig
		// allocation -- a conifer-like draft has millions of contigs.
		const uint64_t seg0 = first_seg[c0], n_segs = first_seg[c1] - first_seg[c0], nc = c1 - c0;
		uint64_t total_events = 0;
		for (uint64_t i = seg0; i < seg0 + n_segs; i++) {
			total_events += segs[i].res.n_events;
		}
		uint64_t piece_events = std::max<uint64_t>(4096, total_events / ((uint64_t)nthreads * 8 + 1));
		if (const char* v = std::getenv("NTB_REPLAY_PIECE_EVENTS")) { // testing aid: tiny pieces put a cut behind (almost) every result
			piece_events = std::max<uint64_t>(1, std::strtoull(v, nullptr, 10));
		}
		std::vector<uint64_t> acc(n_segs);        // accepted segment indices
		std::vector<uint32_t> acc_n(nc, 0);       // how many of contig c's slots are used
		std::vector<uint8_t> acc_cut(n_segs, 0);  // a new piece starts at this accepted result
		std::vector<uint32_t> acc_stale(n_segs);  // the four stale bytes at the start of this accepted result
		run_parallel(nc, [&](uint64_t q) {
			const uint64_t c = c0 + q;
			if (!out.contigs[c].polished) {
				return;
			}
			const uint64_t base = first_seg[c] - seg0;
			uint32_t n = 0;
			uint32_t prev_end = 0;
			uint8_t stale[4] = { 0, 0, 0, 0 };
			auto resolve = [&stale](uint8_t v) -> uint8_t { return (v & STALE_REF) ? stale[v & 3] : v; };
			uint64_t in_piece = 0;
			for (uint64_t i = first_seg[c]; i < first_seg[c + 1]; i++) {
				const Segment& sg = segs[i];
				const uint32_t need = std::max(prev_end, sg.p0);
				if (need >= sg.p1) {
					continue;
				}
				if (n > 0 && in_piece >= piece_events) {
					acc_cut[base + n] = 1;
					in_piece = 0;
				}
				acc[base + n] = i;
				std::memcpy(&acc_stale[base + n], stale, 4);
				n++;
				in_piece += sg.res.n_events + 1;
				const uint8_t next_stale[4] = { resolve(sg.res.stale[0]), resolve(sg.res.stale[1]), resolve(sg.res.stale[2]),
					                            resolve(sg.res.stale[3]) };
				std::memcpy(stale, next_stale, 4);
				prev_end = sg.res.end_pos;
				if (sg.res.status & ST_CONTIG_END) {
					break;
				}
			}
			acc_n[q] = n;
		});
		// the pieces, contig by contig (every polished contig has at least one, possibly without any result)
		std::vector<uint64_t> piece_first(nc + 1, 0);
		for (uint64_t q = 0; q < nc; q++) {
			uint64_t n = 0;
			if (out.contigs[c0 + q].polished) {
				n = 1;
				const uint64_t base = first_seg[c0 + q] - seg0;
				for (uint64_t a = base + 1; a < base + acc_n[q]; a++) {
					n += acc_cut[a];
				}
			}
			piece_first[q + 1] = piece_first[q] + n;
		}
		std::vector<Piece> pieces(piece_first[nc]);
		run_parallel(nc, [&](uint64_t q) {
			const uint64_t c = c0 + q;
			if (!out.contigs[c].polished) {
				return;
			}
			const uint64_t base = first_seg[c] - seg0;
			uint64_t a0 = base, w = piece_first[q];
			for (uint64_t a = base; a <= base + acc_n[q]; a++) {
				if (a == base + acc_n[q] || (a > base && acc_cut[a])) {
					Piece& p = pieces[w++];
					p.contig = (uint32_t)c;
					p.a0 = a0;
					p.a1 = a;
					std::memset(p.stale, 0, 4);
					if (a0 < base + acc_n[q]) {
						std::memcpy(p.stale, &acc_stale[a0], 4);
					}
					a0 = a;
				}
			}
		});
		if (dbg) {
			std::fprintf(stderr, "[ntb] host:   replay (A) accept + cut %.1f ms\n", since(t1));
		}
		const auto t_b = clk::now();
		// (B) every piece on its own
		run_parallel(pieces.size(), [&](uint64_t w) {
			Piece& pc = pieces[w];
			const uint64_t c = pc.contig;
			const uint32_t len = (uint32_t)(offsets[c + 1] - offsets[c] - 1);
			uint64_t n_ev = 0;
			for (uint64_t a = pc.a0; a < pc.a1; a++) {
				n_ev += segs[acc[a]].res.n_events;
			}
			if (n_ev == 0) {
				// nothing happened here: the piece's rope is its root node (most contigs of a fragmented draft)
				ntb_node root;
				std::memset(&root, 0, sizeof root);
				root.node_type = 0;
				root.s_pos = 0;
				root.e_pos = len - 1;
				pc.nodes.assign(1, root);
				return;
			}
			// without a host copy of the bases, substitutions are only reported through the records
			RopeReplay rp(host_bases ? host_bases + offsets[c] : nullptr, len, kp.k, kp.insertion_cap, kp.snv, kp.mask,
			              n_ev + n_ev / 2 + 8, n_ev + 1); // an indel adds 2-6 nodes, a substitution none
			uint8_t stale[4];
			std::memcpy(stale, pc.stale, 4);
			auto resolve = [&stale](uint8_t v) -> uint8_t { return (v & STALE_REF) ? stale[v & 3] : v; };
			for (uint64_t a = pc.a0; a < pc.a1 && !rp.ended; a++) {
				const Segment& sg = segs[acc[a]];
				// the backend hands every walker's events over as one contiguous run, first event first
				const Event* run = sg.res.n_events ? arena_base[(size_t)sg.arena] + sg.res.last_event : nullptr;
				// a substitution is written into the caller's text at the event's position: a random byte of a multi-GB buffer
				// per event, so the lines are requested a few events ahead
				constexpr uint32_t AHEAD = 8;
				char* const text_c = host_bases ? host_bases + offsets[c] : nullptr;
				for (uint32_t q = 0; text_c && q < AHEAD && q < sg.res.n_events; q++) {
					__builtin_prefetch(text_c + run[q].t_pos, 1, 0);
				}
				for (uint32_t q = 0; q < sg.res.n_events; q++) {
					if (text_c && q + AHEAD < sg.res.n_events) {
						__builtin_prefetch(text_c + run[q + AHEAD].t_pos, 1, 0);
					}
					Event ev = run[q];
					ev.base = resolve(ev.base);
					for (int x = 0; x < 3; x++) {
						ev.altbase[x] = resolve(ev.altbase[x]);
					}
					if (!rp.apply(ev)) {
						break;
					}
					if (ev.kind) {
						pc.edits++;
					}
				}
				if (!rp.error.empty()) {
					break;
				}
				const uint8_t next_stale[4] = { resolve(sg.res.stale[0]), resolve(sg.res.stale[1]), resolve(sg.res.stale[2]),
					                            resolve(sg.res.stale[3]) };
				std::memcpy(stale, next_stale, 4);
			}
			pc.error = rp.error;
			pc.ended = rp.ended;
			pc.nodes.swap(rp.rope);
			pc.recs.swap(rp.recs);
		});
		if (dbg) {
			std::fprintf(stderr, "[ntb] host:   replay (B) %zu pieces %.1f ms\n", pieces.size(), since(t_b));
		}
		const auto t_c = clk::now();
		// (C) join the pieces' ropes.  Per contig a short sequential pass patches the position node every cut went through
		// (it keeps s_pos / num_support of the piece on its left and takes e_pos from the piece on its right) and lays the
		// pieces out; the bulk copies then run on the whole pool.
		std::atomic<int> failed(0);
		std::atomic<uint64_t> n_edits(0);
		auto report = [&](uint64_t c, const std::string& what) {
			failed = 1;
			std::lock_guard<std::mutex> guard(host_lock);
			if (first_error.empty()) {
				first_error = "contig " + std::to_string(c) + ": " + what;
			}
		};
		struct CopyJob
		{
			uint64_t piece;
			uint64_t src, count, dst; // nodes [src, src + count) of the piece go to dst
			uint64_t rec_dst;
		};
		std::vector<CopyJob> jobs(pieces.size()); // slot w belongs to piece w; count == ~0 marks "no copy"
		for (CopyJob& j : jobs) {
			j.piece = ~0ULL;
		}
		run_parallel(nc, [&](uint64_t q) {
			const uint64_t c = c0 + q;
			ContigResult& cr = out.contigs[c];
			if (!cr.polished) {
				return;
			}
			const uint64_t p0 = piece_first[q], p1 = piece_first[q + 1];
			uint64_t used = 0;
			uint64_t edits = 0;
			for (uint64_t w = p0; w < p1; w++) {
				if (!pieces[w].error.empty()) {
					report(c, pieces[w].error);
					return;
				}
				edits += pieces[w].edits;
				used = w - p0 + 1;
				if (pieces[w].ended) {
					break; // the reference's main loop ended inside this piece: nothing behind it was ever evaluated
				}
			}
			n_edits += edits;
			if (used == 1) {
				cr.nodes.swap(pieces[p0].nodes);
				cr.srecs.swap(pieces[p0].recs);
				return;
			}
			uint64_t total = 0, total_recs = 0;
			ntb_node* back = nullptr; // the rope's last live node so far (inside the piece that holds it)
			for (uint64_t w = p0; w < p0 + used; w++) {
				Piece& pc = pieces[w];
				uint64_t len = pc.nodes.size();
				if (w + 1 < p0 + used) {
					while (len && pc.nodes[len - 1].node_type == -1) {
						len--; // dead slots behind a piece that is not the last one
					}
				}
				CopyJob job;
				job.piece = w;
				job.rec_dst = total_recs;
				if (w == p0) {
					job.src = 0;
					job.count = len;
				} else {
					if (!back || back->node_type != 0 || len == 0 || pc.nodes[0].node_type != 0) {
						report(c, "rope pieces do not join on a position node");
						return;
					}
					back->e_pos = pc.nodes[0].e_pos;
					job.src = 1;
					job.count = len - 1;
				}
				job.dst = total;
				total += job.count;
				total_recs += pc.recs.size();
				if (job.count) {
					back = &pc.nodes[job.src + job.count - 1];
				}
				jobs[w] = job;
			}
			cr.nodes.resize(total);
			cr.srecs.resize(total_recs);
		});
		if (!failed) {
			run_parallel(jobs.size(), [&](uint64_t j) {
				const CopyJob& job = jobs[j];
				if (job.piece == ~0ULL) {
					return;
				}
				const Piece& pc = pieces[job.piece];
				ContigResult& cr = out.contigs[pc.contig];
				if (job.count) {
					std::memcpy(cr.nodes.data() + job.dst, pc.nodes.data() + job.src, job.count * sizeof(ntb_node));
				}
				if (!pc.recs.empty()) {
					std::memcpy(cr.srecs.data() + job.rec_dst, pc.recs.data(), pc.recs.size() * sizeof(ntb_srec));
				}
			});
		}
		if (dbg) {
			std::fprintf(stderr, "[ntb] host:   replay (C) join %.1f ms\n", since(t_c));
			std::fprintf(stderr, "[ntb] host: replay of contigs [%llu, %llu) %.1f ms (at %.1f ms)\n", (unsigned long long)c0, (unsigned long long)c1,
			             since(t1), since(t_begin));
		}
		std::lock_guard<std::mutex> guard(host_lock);
		host_ms += std::chrono::duration<double, std::milli>(clk::now() - t1).count();
		n_edits_total += n_edits;
	};

	// ---- the pipeline: device phase of group g+1 beside the replay of group g
	std::thread replay_thread;
	int rc_groups = NTB_OK;
	for (size_t g = 0; g + 1 < group_first.size(); g++) {
		const uint64_t c0 = group_first[g], c1 = group_first[g + 1];
		be.scan_until(group_end_pos(g)); // (a group without tasks never reaches the walk() that would have started it)
		if (g + 2 < group_first.size()) {
			be.scan_prefetch(group_end_pos(g + 1));
		}
		rc_groups = device_phase(c0, c1);
		if (replay_thread.joinable()) {
			replay_thread.join();
		}
		if (rc_groups != NTB_OK) {
			break;
		}
		// no walk() of this group follows: its rounds' events are where they will stay
		std::vector<const Event*> arena_base(n_arenas);
		for (size_t r = 0; r < n_arenas; r++) {
			arena_base[r] = be.round_events(r);
		}
		replay_thread = std::thread(host_phase, c0, c1, std::move(arena_base));
	}
	if (replay_thread.joinable()) {
		replay_thread.join();
	}
	be.scan_end();
	if (rc_groups != NTB_OK) {
		return rc_groups;
	}
	if (!first_error.empty()) {
		err = first_error;
		return NTB_EINTERNAL;
	}
	if (dbg) {
		std::fprintf(stderr, "[ntb] host: whole call %.1f ms\n", since(t_begin));
	}
	out.stats.edits = n_edits_total;
	out.stats.ms_host = (float)host_ms;
	return NTB_OK;
}

} // namespace ntb
