/* ntedit_b200 -- C ABI of the B200-native implementation of ntEdit's hot path.
 *
 * ntEdit (reference: ntedit.cpp v2.1.1) exposes no plugin/FFI interface; its hot path sits behind one
 * source-level boundary, `kmerizeAndCorrect(contigHdr, contigSeq, seqLen, bloom, bloomrep, ...)`
 * (ntedit.cpp:1747-1757, called from readAndCorrect ntedit.cpp:2242-2245), plus the BFWrapper filter
 * object it reads (ntedit.cpp:350-401).  The entry points below are exactly what a binding for that
 * boundary needs: load/own a filter on the device (replaces BFWrapper), polish a batch of contigs
 * (replaces the per-contig kmerizeAndCorrect calls of the OpenMP loop) and hand back, per contig, the
 * same three things kmerizeAndCorrect hands to writeEditsToFile (ntedit.cpp:2145-2150): the mutated
 * contig string, the seqNode rope (ntedit.cpp:613-620) and the sRec queue (ntedit.cpp:598-611).
 * INTEGRATION.md shows the stub a maintainer would add to ntedit.cpp.
 *
 * Conventions: plain C types, opaque handles, return 0 on success or a negative NTB_E* code; the
 * message for the calling thread's last error is ntb_last_error().  Nothing here throws or exits.
 * All compute happens in hand-written sm_100a CUDA kernels; there is no CPU fallback -- without a
 * CUDA device every compute entry point fails with NTB_ENODEV.
 */
#ifndef NTEDIT_B200_H
#define NTEDIT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NTB_OK 0
#define NTB_EINVAL (-1)   /* bad argument / unsupported parameter combination */
#define NTB_EIO (-2)      /* file could not be read / written / parsed */
#define NTB_ENODEV (-3)   /* no usable CUDA device */
#define NTB_ECUDA (-4)    /* CUDA runtime error */
#define NTB_ENOMEM (-5)
#define NTB_EINTERNAL (-6)

typedef struct ntb_filter ntb_filter; /* device-resident Bloom / counting Bloom filter (replaces BFWrapper) */
typedef struct ntb_batch ntb_batch;   /* device-resident batch of contigs */
typedef struct ntb_result ntb_result; /* host-side result of one polished batch */

const char* ntb_last_error(void);
const char* ntb_version(void);
int ntb_device_count(void);

/* ---------------------------------------------------------------- filters (BFWrapper, ntedit.cpp:350-401) */
typedef struct ntb_filter_info {
	uint64_t bytes;    /* size of the bit / counter array (btllib header `bytes`) */
	uint32_t k;        /* BFWrapper::get_k()        ntedit.cpp:380 */
	uint32_t hash_num; /* BFWrapper::get_hash_num() ntedit.cpp:382-385 */
	int32_t counting;  /* BFWrapper::is_counting()  ntedit.cpp:378 */
	int32_t device;
	double fpr;        /* btllib get_fpr(): (occupancy)^hash_num, the value print_details shows, ntedit.cpp:387-395 */
} ntb_filter_info;

/* Page-locked host memory for batch buffers.  ntb_polish_batch streams its upload straight out of such a buffer while the
 * scan of the earlier pieces already runs; from pageable memory the driver stages every piece through its own pinned
 * buffer first.  NULL when no device / no memory. */
void* ntb_host_alloc(size_t bytes);
void ntb_host_free(void* p);

/* Replaces `BFWrapper bloom(path)` (ntedit.cpp:355-364, 2438): parses the btllib header
 * ([BTLKmerBloomFilter_v*] / [BTLKmerCountingBloomFilter_v*], TOML keys in any order, [HeaderEnd]),
 * uploads `bytes` raw bytes to `device` and computes the occupancy on the device. */
int ntb_filter_load(const char* path, int device, ntb_filter** out);
/* Empty filter on the device (builder side; src/ntedit_make_genome_bf.cpp:143-150). */
int ntb_filter_create(uint64_t bytes, uint32_t k, uint32_t hash_num, int counting, int device, ntb_filter** out);
/* Wrap filter bytes that already live on `device` (e.g. received through a NCCL broadcast); not owned.  `dev_bytes` must be
 * 16-byte aligned (NTB_EINVAL otherwise); inserting into a wrapped filter additionally needs `bytes` to be a multiple of 4
 * (the builder updates whole 32-bit words). */
int ntb_filter_wrap_device(void* dev_bytes, uint64_t bytes, uint32_t k, uint32_t hash_num, int counting, int device,
                           ntb_filter** out);
/* A replica of `src` on another device of the same process: one device-to-device copy (NVLink / NVSwitch peer copy where the
 * devices are peers) instead of reading the file once per GPU -- the "single broadcast at load" of a multi-GPU run
 * (ntedit.cpp:2438 loads one filter that every OpenMP thread shares). */
int ntb_filter_replicate(ntb_filter* src, int device, ntb_filter** out);
int ntb_filter_get_info(ntb_filter* f, ntb_filter_info* info);
void* ntb_filter_device_ptr(ntb_filter* f);
/* Insert every canonical all-ACGT k-mer of the contigs (btllib KmerBloomFilter::insert(seq) as used by
 * src/ntedit_make_genome_bf.cpp:151-156).  Bit filter: atomic OR -- the arrays btllib builds.  Counting filter (our addition;
 * the reference builds none itself): EVERY one of the hash_num counters is incremented, saturating at 255.  This is NOT
 * btllib's CountingBloomFilter::insert, which raises only the counters equal to the current minimum (an update whose
 * result depends on the insertion order, so no parallel builder can reproduce a sequential one byte for byte): filters
 * built here count collisions higher than ntStat / btllib ones do.  Polishing reads any counting filter the same way. */
int ntb_filter_insert(ntb_filter* f, const char* bases, const uint64_t* offsets, uint64_t n_contigs);
int ntb_filter_insert_batch(ntb_filter* f, const ntb_batch* b);
/* Write the btllib on-disk format (src/ntedit_make_genome_bf.cpp:158-162). */
int ntb_filter_save(ntb_filter* f, const char* path);
int ntb_filter_download(ntb_filter* f, void* host_dst, uint64_t bytes);
void ntb_filter_free(ntb_filter* f);

/* ---------------------------------------------------------------- parameters (namespace opt, ntedit.cpp:99-133) */
typedef struct ntb_params {
	uint32_t jump;           /* -j [3] */
	int32_t mode;            /* -m 0/1/2 */
	int32_t snv;             /* -s */
	int32_t mask;            /* -a */
	uint32_t max_insertions; /* -i [5], <= 5 */
	uint32_t max_deletions;  /* -d [5], <= 10 */
	float edit_threshold;    /* -y [9] */
	float missing_threshold; /* -x [5] */
	float edit_ratio;        /* -Y [0.5] */
	float missing_ratio;     /* -X [0.5] */
	int32_t use_ratio;       /* set when -X or -Y was given, ntedit.cpp:2316-2323 */
	uint32_t min_threshold;  /* -p [1] */
	uint32_t max_threshold;  /* -q [255] */
	uint32_t min_contig_len; /* -z [100]: shorter contigs are not polished and get no result */
	uint32_t segment_len;    /* 0 = library default; device work-unit length in bases */
} ntb_params;

/* Defaults of ntedit.cpp:99-133.  k, hash_num and insertion_cap (= k*1.5, ntedit.cpp:2450) always come from the
 * primary filter, as in the reference (ntedit.cpp:2439-2451). */
void ntb_params_init(ntb_params* p);

/* ---------------------------------------------------------------- batches
 * A batch is `n_contigs` NUL-terminated sequences laid end to end in one buffer: contig c occupies
 * bases[offsets[c] .. offsets[c+1]-2] and bases[offsets[c+1]-1] == 0 (exactly what copying kseq's
 * seq->seq.s including its terminator gives, ntedit.cpp:2230).  offsets has n_contigs+1 entries.
 * A contig must be shorter than 2^32-2 bases (positions are `unsigned` in the reference). */
int ntb_batch_upload(const char* bases, const uint64_t* offsets, uint64_t n_contigs, int device, ntb_batch** out);
/* Same, from buffers that already live on `device` (not copied, not owned). */
int ntb_batch_wrap_device(void* dev_bases, const uint64_t* host_offsets, uint64_t n_contigs, int device, ntb_batch** out);
uint64_t ntb_batch_total_bases(const ntb_batch* b);
void ntb_batch_free(ntb_batch* b);

/* K1: for every position t of the batch buffer (tail of the window [t-k+1, t]):
 *   counts[t] = BFWrapper::get_count-style value of the window's k-mer (bit filter: contains 0/1,
 *               counting filter: min counter)                         [ntedit.cpp:368-376]
 *   valid bit t = the window lies inside one contig and holds only accepted bases [ntedit.cpp:493-499]
 * counts (bytes = total buffer length) and valid_bits ((len+31)/32 uint32 words, LSB first) are host
 * buffers; either may be NULL. */
int ntb_scan(ntb_filter* f, const char* bases, const uint64_t* offsets, uint64_t n_contigs, uint8_t* counts,
             uint32_t* valid_bits);

/* ---------------------------------------------------------------- polishing (kmerizeAndCorrect, ntedit.cpp:1747-2151) */
typedef struct ntb_node { /* seqNode, ntedit.cpp:613-620 */
	int32_t node_type;    /* -1 dead, 0 slice [s_pos, e_pos] of the contig, 1 inserted character */
	uint32_t s_pos, e_pos;
	uint32_t num_support;
	uint8_t c;
	uint8_t pad_[3];
} ntb_node;

typedef struct ntb_srec { /* sRec, ntedit.cpp:598-611 */
	uint32_t pos;
	uint32_t num_support;
	uint32_t altsupp1, altsupp2, altsupp3;
	uint8_t draft_char, sub_base, altbase1, altbase2, altbase3;
	uint8_t pad_[3];
} ntb_srec;

typedef struct ntb_stats {
	uint64_t bases;          /* bases in polished contigs (>= min_contig_len) */
	uint64_t contigs;        /* polished contigs */
	uint64_t sites;          /* error sites evaluated on the device */
	uint64_t edits;          /* accepted edits (substitutions + insertions + deletions) */
	uint64_t segments;       /* device work units launched in total */
	uint64_t reruns;         /* work units re-launched by the stitcher */
	uint32_t rounds;         /* device rounds (1 = no re-run was needed) */
	uint32_t kernel_launches;
	float ms_scan;           /* K1 device time (CUDA events) */
	float ms_walk;           /* K2 device time, all rounds */
	float ms_h2d, ms_d2h;
	float ms_host;           /* host replay + stitch wall time */
	float ms_pre;            /* K2p device time: site pre-evaluation passes in front of the first walker round */
	uint32_t pad_;
} ntb_stats;

/* Replaces the kmerizeAndCorrect calls of readAndCorrect's loop (ntedit.cpp:2220-2245) for a whole batch.
 * `bases` is mutated in place exactly as the reference mutates contigSeq (accepted substitutions,
 * -a soft-masking, and the reference's upper-casing of tried positions).  `rep` may be NULL (-e absent).
 * On success *out holds, per contig, the rope and substitution records to feed writeEditsToFile
 * (ntedit.cpp:925-1213). */
int ntb_polish_batch(ntb_filter* bloom, ntb_filter* rep, const ntb_params* p, char* bases, const uint64_t* offsets,
                     uint64_t n_contigs, ntb_result** out);
/* Same with the batch already resident on the device (no host->device copy of the bases).  host_bases may
 * be NULL, in which case accepted substitutions are only reported through the records. */
int ntb_polish_device(ntb_filter* bloom, ntb_filter* rep, const ntb_params* p, ntb_batch* batch, char* host_bases,
                      ntb_result** out);

/* polished(c) is 0 for contigs shorter than min_contig_len (the reference drops them from all outputs). */
int ntb_result_contig(const ntb_result* r, uint64_t contig, int* polished, const ntb_node** nodes, uint64_t* n_nodes,
                      const ntb_srec** srecs, uint64_t* n_srecs);
int ntb_result_stats(const ntb_result* r, ntb_stats* st);
void ntb_result_free(ntb_result* r);

/* ---------------------------------------------------------------- writer (writeEditsToFile, ntedit.cpp:925-1213)
 * Bit-exact formatter for one polished contig; appends to caller-owned growing buffers (realloc'ed).  The VCF rows
 * are produced with an empty ClinVar map.  Pass NULL for a stream that is not wanted. */
typedef struct ntb_strbuf {
	char* data;
	size_t len, cap;
} ntb_strbuf;
int ntb_format_contig(const char* header, const char* seq, const ntb_node* nodes, uint64_t n_nodes, const ntb_srec* srecs,
                      uint64_t n_srecs, int snv, ntb_strbuf* fa, ntb_strbuf* tsv, ntb_strbuf* vcf);
/* Header line of _changes.tsv (ntedit.cpp:2175-2188). */
int ntb_format_tsv_header(uint32_t k, uint32_t jump, int counting, ntb_strbuf* tsv);
void ntb_strbuf_free(ntb_strbuf* b);

#ifdef __cplusplus
}
#endif
#endif
