// kseq-compatible FASTA/FASTQ reader, shared by the ntedit-b200 command line tools.  The bytes come from a ByteSource
// (bytesource.hpp): read and inflated ahead of the parser by its own thread(s).
#pragma once
#include "bytesource.hpp"

#include <cstring>
#include <string>
#include <vector>

namespace ntb {

// Same record semantics as lib/kseq.h:175-215: name = up to the first whitespace, comment = rest of the header line,
// sequence lines concatenated with every character kept except newlines / carriage returns... (kseq keeps all
// printable characters of a sequence line; it drops only the line terminator and isgraph-failing bytes)
class FastxReader
{
  public:
	explicit FastxReader(const std::string& path, unsigned inflate_threads = 4) : src_(path, inflate_threads) {}
	bool ok() const { return src_.ok(); }
	const char* kind() const { return src_.kind(); }

	// reads the next record; sequence is appended to `seq` (anything with size / push_back / append(ptr, n) / back / pop_back).
	// Returns false at end of file.
	template<class Seq>
	bool next(std::string& name, std::string& comment, Seq& seq)
	{
		int c;
		if (last_char_ == 0) { // jump to the next header line
			while ((c = getc_()) != -1 && c != '>' && c != '@') {
			}
			if (c == -1) {
				return false;
			}
			last_char_ = c;
		}
		name.clear();
		comment.clear();
		// name: up to the first whitespace
		while ((c = getc_()) != -1 && !isspace_(c)) {
			name.push_back((char)c);
		}
		if (c == -1 && name.empty()) {
			return false;
		}
		if (c != '\n' && c != -1) { // comment: the rest of the line
			while ((c = getc_()) != -1 && c != '\n') {
				comment.push_back((char)c);
			}
			if (!comment.empty() && comment.back() == '\r') {
				comment.pop_back();
			}
		}
		const size_t seq0 = seq.size();
		while ((c = getc_()) != -1 && c != '>' && c != '+' && c != '@') {
			if (c == '\n') {
				continue;
			}
			seq.push_back((char)c);
			append_line_(seq); // rest of the line
		}
		if (c == '>' || c == '@') {
			last_char_ = c;
		} else {
			last_char_ = 0;
		}
		if (c != '+') {
			return true;
		}
		// FASTQ: skip the rest of the '+' line, then as many quality characters as there are bases
		while ((c = getc_()) != -1 && c != '\n') {
		}
		if (c == -1) {
			return true;
		}
		const size_t want = seq.size() - seq0;
		size_t have = 0;
		std::string q;
		while (have < want && (c = getc_()) != -1) {
			if (c == '\n') {
				continue;
			}
			q.clear();
			q.push_back((char)c);
			append_line_(q);
			have += q.size();
		}
		last_char_ = 0;
		return true;
	}

  private:
	static bool isspace_(int c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

	int getc_()
	{
		if (pos_ >= len_) {
			if (eof_) {
				return -1;
			}
			if (!src_.next(buf_) || buf_.empty()) {
				eof_ = true;
				len_ = pos_ = 0;
				return -1;
			}
			len_ = buf_.size();
			pos_ = 0;
		}
		return (unsigned char)buf_[pos_++];
	}

	// appends the rest of the current line (without its terminator) to s
	template<class Seq>
	void append_line_(Seq& s)
	{
		for (;;) {
			if (pos_ >= len_) {
				const int c = getc_();
				if (c == -1) {
					break;
				}
				pos_--;
			}
			const char* b = buf_.data() + pos_;
			const char* nl = (const char*)std::memchr(b, '\n', len_ - pos_);
			const size_t n = nl ? (size_t)(nl - b) : len_ - pos_;
			s.append(b, n);
			pos_ += n;
			if (nl) {
				pos_++; // consume the newline
				break;
			}
		}
		if (!s.empty() && s.back() == '\r') {
			s.pop_back();
		}
	}

	ByteSource src_;
	std::vector<char> buf_;
	size_t pos_ = 0, len_ = 0;
	bool eof_ = false;
	int last_char_ = 0;
};


} // namespace ntb
