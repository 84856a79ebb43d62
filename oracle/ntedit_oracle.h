/* ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the product
 * path (ntedit_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker.
 *
 * Plain-C CPU restatement of ntEdit's hot path (reference: /root/reference/ntedit.cpp, v2.1.1) and
 * of the btllib arithmetic it calls.  btllib (github.com/bcgsc/btllib, version UNPINNED by the
 * reference: meson.build:20, azure-pipelines.yml:25) is not vendored in the reference tree, so its
 * published ntHash2 / Bloom-filter algorithms are restated from SURVEY.md Appendix A/B.
 *
 * Parity status:
 *   - ntHash: pinned by btllib's unit-test vector "ACATGCATGCA" k=5 h=3 (tests/test_oracle.py).
 *   - Bloom bit addressing / file header: PARITY UNPINNED (no .bf fixture and no btllib source in
 *     the reference tree); the same restatement backs oracle/shim, so reference-vs-ours comparisons
 *     are self-consistent.
 *   - Engine control flow: validated against oracle/_ref/ntedit_ref (the unmodified reference
 *     ntedit.cpp compiled against oracle/shim) on the demo draft and on fuzzed inputs.
 */
#ifndef NTEDIT_ORACLE_H
#define NTEDIT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- ntHash (btllib hashing_internals; call sites ntedit.cpp:403-452) ---- */
uint64_t orc_srol(uint64_t x);
uint64_t orc_sror(uint64_t x);
uint64_t orc_srol_n(uint64_t x, unsigned d);
uint64_t orc_seed(unsigned char c);
uint64_t orc_base_forward_hash(const char* s, unsigned k);
uint64_t orc_base_reverse_hash(const char* s, unsigned k);
uint64_t orc_next_forward_hash(uint64_t fh, unsigned k, unsigned char out, unsigned char in);
uint64_t orc_next_reverse_hash(uint64_t rh, unsigned k, unsigned char out, unsigned char in);
void orc_extend_hashes(uint64_t base, unsigned k, unsigned h, uint64_t* out);
/* NTMC64 seed / roll / changelast, ntedit.cpp:403-452 */
void orc_ntmc64_seed(const char* s, unsigned k, unsigned h, uint64_t* fh, uint64_t* rh, uint64_t* hv);
void orc_ntmc64_roll(unsigned char out, unsigned char in, unsigned k, unsigned h, uint64_t* fh, uint64_t* rh, uint64_t* hv);
void orc_ntmc64_changelast(unsigned char out, unsigned char in, unsigned k, unsigned h, uint64_t* fh, uint64_t* rh, uint64_t* hv);

/* ---- filters (btllib KmerBloomFilter / KmerCountingBloomFilter8; ntedit.cpp:350-401) ---- */
typedef struct orc_filter {
	uint8_t* data;
	uint64_t bytes;
	unsigned k;
	unsigned h;
	int counting; /* 0 = bit filter, 1 = 8-bit counting filter */
} orc_filter;

orc_filter* orc_filter_new(uint64_t bytes, unsigned k, unsigned h, int counting);
void orc_filter_free(orc_filter* f);
int orc_filter_contains(const orc_filter* f, const uint64_t* hv);      /* BFWrapper::contains */
unsigned orc_filter_count(const orc_filter* f, const uint64_t* hv);    /* BFWrapper::get_count */
void orc_filter_insert_hashes(orc_filter* f, const uint64_t* hv);
/* insert every all-ACGT k-mer of seq (canonical hashes), btllib KmerBloomFilter::insert(seq) as used by
 * src/ntedit_make_genome_bf.cpp:151-156.  For counting filters every occurrence increments (saturating). */
void orc_filter_insert_seq(orc_filter* f, const char* seq, size_t len);
int orc_filter_save(const orc_filter* f, const char* path);
orc_filter* orc_filter_load(const char* path);
double orc_filter_fpr(const orc_filter* f);

/* K1 restatement: for every tail position t (window [t-k+1,t]) of seq, out[t] =
 *   0xFF if the window holds a non-accepted base (ntedit.cpp:493-499) or t < k-1,
 *   else the count (bit filter: 0/1; counting filter: min counter).                       */
void orc_scan_counts(const orc_filter* f, const char* seq, size_t len, uint8_t* out);

/* ---- engine (ntedit.cpp:524-2151) ---- */
typedef struct orc_params {
	unsigned k, h;
	unsigned jump;           /* -j */
	int mode;                /* -m */
	int snv;                 /* -s */
	int mask;                /* -a */
	unsigned max_insertions; /* -i */
	unsigned max_deletions;  /* -d */
	float edit_threshold;    /* -y */
	float missing_threshold; /* -x */
	float edit_ratio;        /* -Y */
	float missing_ratio;     /* -X */
	int use_ratio;
	unsigned insertion_cap;  /* k*1.5, ntedit.cpp:2450 */
	unsigned min_threshold;  /* -p */
	unsigned max_threshold;  /* -q */
	int secbf;               /* -e given */
} orc_params;

void orc_params_default(orc_params* p, unsigned k, unsigned h);

/* rope node / substitution record as in ntedit.cpp:598-620 */
typedef struct orc_node {
	int32_t node_type; /* -1 dead, 0 position slice, 1 inserted char */
	uint32_t s_pos, e_pos;
	uint32_t num_support;
	uint8_t c;
} orc_node;

typedef struct orc_srec {
	uint32_t pos;
	uint8_t draft_char, sub_base;
	uint32_t num_support;
	uint8_t altbase1, altbase2, altbase3;
	uint32_t altsupp1, altsupp2, altsupp3;
} orc_srec;

typedef struct orc_result {
	orc_node* nodes;
	size_t n_nodes; /* full vector including dead entries */
	orc_srec* srecs;
	size_t n_srecs;
} orc_result;

/* kmerizeAndCorrect (ntedit.cpp:1747-2151) without the writer: seq (len bytes, writable) is mutated in
 * place exactly as the reference mutates contigSeq.  Returns 0 on success. */
int orc_polish_contig(char* seq, uint32_t len, const orc_filter* bloom, const orc_filter* bloomrep,
                      const orc_params* p, orc_result* out);
void orc_result_free(orc_result* r);

/* writeEditsToFile (ntedit.cpp:925-1213) for _edited.fa and _changes.tsv (and the VCF body with no
 * ClinVar map): appends to growing buffers.  Returned strings are malloc'ed, caller frees with free(). */
int orc_write_contig(const char* hdr, const char* seq, uint32_t len, const orc_result* r, const orc_params* p,
                     char** fa_out, size_t* fa_len, char** tsv_out, size_t* tsv_len, char** vcf_out, size_t* vcf_len);
/* TSV header line, ntedit.cpp:2175-2188 */
int orc_tsv_header(const orc_params* p, int counting, char* buf, size_t buflen);

#ifdef __cplusplus
}
#endif
#endif
