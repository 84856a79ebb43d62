// Bit-exact formatter for ntEdit's three outputs (writeEditsToFile, ntedit.cpp:925-1213; TSV header ntedit.cpp:2175-2188;
// VCF header ntedit.cpp:2192-2211).  Walks the rope once, emitting the FASTA record, the _changes.tsv rows and the VCF
// rows in the reference's interleaving: an insertion row is written when the next position node is reached, substitution
// rows up to the end of each position node, a deletion row when the next position node does not start where the
// previous one ended.
#pragma once
#include "../../include/ntedit_b200.h"

#include <cmath>
#include <cstdio>
#include <ctime>
#include <map>
#include <string>

namespace ntb {

typedef std::map<std::string, std::string> ClinvarMap; // "<chrom>><REF><pos><ALT>" -> INFO, ntedit.cpp:2261-2274

inline std::string
upper_copy(const std::string& s)
{
	std::string r = s;
	for (char& c : r) {
		c = (char)std::toupper((unsigned char)c);
	}
	return r;
}

inline void
append_clinvar(std::string& info, const ClinvarMap* cv, const std::string& key)
{
	if (cv) {
		auto it = cv->find(key);
		if (it != cv->end() && !it->second.empty()) {
			info += "^";
			info += it->second;
			return;
		}
	}
	info += "^NA";
}

inline std::string
tsv_header(uint32_t k, uint32_t jump, bool counting)
{
	std::string s = "ID\tbpPosition+1\tOriginalBase\tNewBase\t";
	if (counting) {
		s += "Coverage (max 255)";
	} else {
		char buf[96];
		// the reference streams a double: default ostream formatting == %g
		std::snprintf(buf, sizeof buf, "Support %u-mer (out of %g)", k, std::ceil((double)k / (double)jump));
		s += buf;
	}
	const char* evi = counting ? "Coverage" : "Support";
	s += std::string("\tAlt.Base1\tAlt.") + evi + "1\tAlt.Base2\tAlt." + evi + "2\tAlt.Base3\tAlt." + evi + "3\n";
	return s;
}

inline std::string
vcf_header(const std::string& program, const std::string& draft_filename)
{
	time_t now = time(nullptr);
	tm* ltm = localtime(&now);
	char date[32];
	std::snprintf(date, sizeof date, "%d%02d%02d", 1900 + ltm->tm_year, 1 + ltm->tm_mon, ltm->tm_mday);
	std::string s = "##fileformat=VCFv4.2\n";
	s += std::string("##fileDate=") + date + "\n";
	s += "##source=" + program + "\n";
	s += "##reference=file:" + draft_filename + "\n";
	s += "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n";
	s += "##INFO=<ID=AD,Number=2,Type=Integer,Description=\"Kmer Depth\">\n";
	s += "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tINTEGRATION\n";
	return s;
}

// one substitution record -> TSV row (unless it is an SNV-mode "no edit" record) and VCF row, ntedit.cpp:984-1164
inline void
format_substitution(const std::string& hdr, const ntb_srec& r, bool snv, const ClinvarMap* cv, std::string* tsv, std::string* vcf)
{
	const bool edit_row = !(snv && r.draft_char == r.sub_base);
	const std::string pos1 = std::to_string((unsigned long long)r.pos + 1);
	char altb[3];
	uint32_t alts[3];
	int na = 0;
	if (r.altsupp1 > 0) {
		altb[na] = (char)r.altbase1;
		alts[na++] = r.altsupp1;
	}
	if (r.altsupp2 > 0) {
		altb[na] = (char)r.altbase2;
		alts[na++] = r.altsupp2;
	}
	if (r.altsupp3 > 0) {
		altb[na] = (char)r.altbase3;
		alts[na++] = r.altsupp3;
	}
	if (tsv && edit_row) {
		*tsv += hdr;
		*tsv += '\t';
		*tsv += pos1;
		*tsv += '\t';
		*tsv += (char)r.draft_char;
		*tsv += '\t';
		*tsv += (char)r.sub_base;
		*tsv += '\t';
		*tsv += std::to_string(r.num_support);
		for (int i = 0; i < na; i++) {
			*tsv += '\t';
			*tsv += altb[i];
			*tsv += '\t';
			*tsv += std::to_string(alts[i]);
		}
		*tsv += '\n';
	}
	if (!vcf) {
		return;
	}
	const char draft_up = (char)std::toupper(r.draft_char);
	const std::string idbase = hdr + ">" + draft_up + pos1;
	std::string base(1, (char)r.sub_base);
	std::string support = std::to_string(r.num_support);
	std::string info;
	append_clinvar(info, cv, idbase + draft_up);
	if (edit_row) {
		append_clinvar(info, cv, idbase + (char)std::toupper((unsigned char)base[0]));
	}
	uint32_t best_supp = 0;
	char best_alt = '1';
	const char* gt = "1/1";
	if (na > 0) {
		if (snv && !edit_row) {
			for (int i = 0; i < na; i++) {
				if (alts[i] > best_supp) {
					best_supp = alts[i];
					best_alt = altb[i];
				}
			}
			base = std::string(1, best_alt);
			append_clinvar(info, cv, idbase + (char)std::toupper((unsigned char)best_alt));
			support += "," + std::to_string(best_supp);
			gt = "0/1";
		} else if (snv) {
			bool ref = false;
			for (int i = 0; i < na; i++) {
				if ((char)r.draft_char == altb[i]) { // the draft base itself has support: heterozygous with the reference
					best_supp = alts[i];
					ref = true;
					break;
				}
				if (alts[i] > best_supp) {
					best_supp = alts[i];
					best_alt = altb[i];
				}
			}
			if (ref) {
				support = std::to_string(best_supp) + "," + support;
				gt = "0/1";
			} else {
				gt = "1/2";
				support += "," + std::to_string(best_supp);
				base += ",";
				base += best_alt;
				append_clinvar(info, cv, idbase + (char)std::toupper((unsigned char)best_alt));
			}
		} else {
			for (int i = 0; i < na; i++) {
				if ((char)r.draft_char == altb[i]) {
					continue;
				}
				if (alts[i] > best_supp) {
					best_supp = alts[i];
					best_alt = altb[i];
				}
			}
			gt = "1/2";
			support += "," + std::to_string(best_supp);
			base += ",";
			base += best_alt;
			append_clinvar(info, cv, idbase + (char)std::toupper((unsigned char)best_alt));
		}
	}
	*vcf += hdr;
	*vcf += '\t';
	*vcf += pos1;
	*vcf += "\t.\t";
	*vcf += (char)r.draft_char;
	*vcf += '\t';
	*vcf += base;
	*vcf += "\t.\tPASS\tAD=";
	*vcf += support;
	*vcf += info;
	*vcf += "\tGT\t";
	*vcf += gt;
	*vcf += '\n';
}

inline void
format_contig(const std::string& hdr, const char* seq, const ntb_node* nodes, size_t n_nodes, const ntb_srec* srecs, size_t n_srecs,
              bool snv, const ClinvarMap* cv, std::string* fa, std::string* tsv, std::string* vcf)
{
	if (fa) {
		*fa += '>';
		*fa += hdr;
		*fa += '\n';
	}
	std::string pending_ins;
	long pending_support = -1;
	uint32_t pos = 0;
	size_t si = 0;
	for (size_t ni = 0; ni < n_nodes && nodes[ni].node_type != -1; ni++) {
		const ntb_node& nd = nodes[ni];
		if (nd.node_type == 0) {
			if (!pending_ins.empty()) {
				// the row reports the draft base that sits |insertion| before this node's start, ntedit.cpp:957
				const char draft = seq[nd.s_pos - pending_ins.size()];
				const std::string p = std::to_string(pos), supp = std::to_string(pending_support);
				if (tsv) {
					*tsv += hdr + "\t" + p + "\t" + draft + "\t+" + pending_ins + "\t" + supp + "\n";
				}
				if (vcf) {
					std::string info;
					append_clinvar(info, cv, hdr + ">" + (char)std::toupper((unsigned char)draft) + p + upper_copy(std::string(1, draft) + pending_ins));
					*vcf += hdr + "\t" + p + "\t.\t" + draft + "\t" + draft + pending_ins + "\t.\tPASS\tAD=" + supp + info + "\tGT\t1/1\n";
				}
				pending_ins.clear();
				pending_support = -1;
			}
			while (si < n_srecs && srecs[si].pos <= nd.e_pos) {
				format_substitution(hdr, srecs[si], snv, cv, tsv, vcf);
				si++;
			}
			if (fa) {
				fa->append(seq + nd.s_pos, (size_t)nd.e_pos - nd.s_pos + 1);
			}
			pos = nd.e_pos + 1;
		} else if (nd.node_type == 1) {
			pending_ins += (char)nd.c;
			if (pending_support == -1) {
				pending_support = (long)nd.num_support;
			}
			if (fa) {
				*fa += (char)nd.c;
			}
		}
		if (ni + 1 < n_nodes) {
			const ntb_node& nx = nodes[ni + 1];
			if (nx.node_type == 0 && nx.s_pos != pos) {
				// deletion of draft[pos, nx.s_pos), ntedit.cpp:1180-1209
				const std::string p = std::to_string(pos), supp = std::to_string(nx.num_support);
				if (tsv) {
					*tsv += hdr + "\t" + p + "\t" + seq[pos] + "\t-";
					tsv->append(seq + pos, (size_t)nx.s_pos - pos);
					*tsv += "\t" + supp + "\n";
				}
				if (vcf) {
					const std::string ref(seq + pos - 1, (size_t)nx.s_pos - pos + 1);
					std::string info;
					append_clinvar(info, cv, hdr + ">" + upper_copy(ref) + p + (char)std::toupper((unsigned char)seq[pos - 1]));
					*vcf += hdr + "\t" + p + "\t.\t" + ref + "\t" + seq[pos - 1] + "\t.\tPASS\tAD=" + supp + info + "\tGT\t1/1\n";
				}
			}
		}
	}
	if (fa) {
		*fa += '\n';
	}
}

} // namespace ntb
