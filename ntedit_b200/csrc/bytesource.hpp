// Decompressed bytes of a plain / gzip / BGZF file as a stream of chunks produced ahead of the consumer, for the
// kseq-compatible reader (fastx.hpp).  The reference reads through gzread on the thread that also parses
// (lib/kseq.h:86-108 via ntedit.cpp:2158-2160); here
//   * a producer thread reads and inflates while the parser works on the previous chunk (any gzip file, plain files), and
//   * BGZF files (bgzip: independent <= 64 KB deflate blocks whose compressed size is in the header) are inflated by a pool
//     of threads, blocks in parallel, chunks delivered in file order.
// Concatenated gzip members are handled as zlib's gzread handles them (one continuous stream).
#pragma once
#include <zlib.h>

#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <future>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace ntb {

class ByteSource
{
  public:
	static constexpr size_t CHUNK = (size_t)16 << 20; // decompressed bytes per chunk (BGZF: compressed bytes per job)

	explicit ByteSource(const std::string& path, unsigned threads = 4) : threads_(threads < 1 ? 1 : threads)
	{
		FILE* fp = std::fopen(path.c_str(), "rb");
		if (!fp) {
			return;
		}
		unsigned char hdr[18];
		const size_t n = std::fread(hdr, 1, sizeof hdr, fp);
		std::fclose(fp);
		ok_ = true;
		path_ = path;
		if (n >= 2 && hdr[0] == 0x1f && hdr[1] == 0x8b) {
			// BGZF: gzip member with FEXTRA whose first subfield is 'B' 'C' (SAM spec 4.1)
			kind_ = (n >= 18 && (hdr[3] & 4) && hdr[12] == 'B' && hdr[13] == 'C' && hdr[14] == 2 && hdr[15] == 0) ? BGZF : GZIP;
		} else {
			kind_ = PLAIN;
		}
		producer_ = std::thread([this]() { produce(); });
	}

	~ByteSource()
	{
		{
			std::lock_guard<std::mutex> g(m_);
			stop_ = true;
		}
		cv_space_.notify_all();
		cv_data_.notify_all();
		if (producer_.joinable()) {
			producer_.join();
		}
	}

	bool ok() const { return ok_; }
	const char* kind() const { return kind_ == BGZF ? "bgzf" : kind_ == GZIP ? "gzip" : "plain"; }

	// the next chunk of decompressed bytes (swapped into `out`); false at the end of the file
	bool next(std::vector<char>& out)
	{
		std::unique_lock<std::mutex> lk(m_);
		cv_data_.wait(lk, [this]() { return !ready_.empty() || done_; });
		if (ready_.empty()) {
			return false;
		}
		out.swap(ready_.front());
		ready_.pop_front();
		lk.unlock();
		cv_space_.notify_one();
		return true;
	}

  private:
	enum Kind
	{
		PLAIN,
		GZIP,
		BGZF
	};

	bool push(std::vector<char>&& chunk) // false: the consumer went away
	{
		std::unique_lock<std::mutex> lk(m_);
		cv_space_.wait(lk, [this]() { return ready_.size() < 4 || stop_; });
		if (stop_) {
			return false;
		}
		ready_.push_back(std::move(chunk));
		lk.unlock();
		cv_data_.notify_one();
		return true;
	}

	void finish()
	{
		{
			std::lock_guard<std::mutex> g(m_);
			done_ = true;
		}
		cv_data_.notify_all();
	}

	void produce()
	{
		if (kind_ == BGZF) {
			produce_bgzf();
		} else {
			// gzread passes plain files through unchanged (as the reference's gzopen does) and inflates gzip streams
			gzFile gz = gzopen(path_.c_str(), "r");
			if (gz) {
				gzbuffer(gz, 1 << 20);
				for (;;) {
					std::vector<char> chunk(CHUNK);
					const int n = gzread(gz, chunk.data(), (unsigned)chunk.size());
					if (n <= 0) {
						break;
					}
					chunk.resize((size_t)n);
					if (!push(std::move(chunk))) {
						break;
					}
				}
				gzclose(gz);
			}
		}
		finish();
	}

	// inflates the BGZF blocks held in `comp` (whole blocks, back to back)
	static std::vector<char> inflate_blocks(std::vector<unsigned char> comp)
	{
		std::vector<char> out;
		out.reserve(comp.size() * 4);
		size_t p = 0;
		z_stream zs;
		while (p + 18 <= comp.size()) {
			const unsigned char* b = comp.data() + p;
			const size_t x = (size_t)b[10] | ((size_t)b[11] << 8);
			const size_t bsize = ((size_t)b[16] | ((size_t)b[17] << 8)) + 1;
			if (p + bsize > comp.size() || bsize < 12 + x + 8) {
				break;
			}
			const size_t isize = (size_t)b[bsize - 4] | ((size_t)b[bsize - 3] << 8) | ((size_t)b[bsize - 2] << 16) | ((size_t)b[bsize - 1] << 24);
			const size_t o = out.size();
			out.resize(o + isize);
			std::memset(&zs, 0, sizeof zs);
			if (isize && inflateInit2(&zs, -15) == Z_OK) {
				zs.next_in = const_cast<unsigned char*>(b + 12 + x);
				zs.avail_in = (unsigned)(bsize - 12 - x - 8);
				zs.next_out = (unsigned char*)out.data() + o;
				zs.avail_out = (unsigned)isize;
				inflate(&zs, Z_FINISH);
				inflateEnd(&zs);
			}
			p += bsize;
		}
		return out;
	}

	void produce_bgzf()
	{
		FILE* fp = std::fopen(path_.c_str(), "rb");
		if (!fp) {
			return;
		}
		std::deque<std::future<std::vector<char>>> jobs;
		std::vector<unsigned char> carry; // bytes read but not yet handed out (an incomplete block at the end of a read)
		bool eof = false;
		while (!eof || !jobs.empty() || !carry.empty()) {
			while (!eof && jobs.size() < threads_ * 2) {
				std::vector<unsigned char> buf(carry);
				const size_t have = buf.size();
				buf.resize(have + CHUNK);
				const size_t n = std::fread(buf.data() + have, 1, CHUNK, fp);
				buf.resize(have + n);
				eof = n < CHUNK;
				// whole blocks only
				size_t p = 0;
				while (p + 18 <= buf.size()) {
					const size_t bsize = ((size_t)buf[p + 16] | ((size_t)buf[p + 17] << 8)) + 1;
					if (p + bsize > buf.size()) {
						break;
					}
					p += bsize;
				}
				carry.assign(buf.begin() + (long)p, buf.end());
				buf.resize(p);
				if (eof) {
					carry.clear(); // (a truncated last block is dropped, as gzread would fail on it)
				}
				if (!buf.empty()) {
					jobs.push_back(std::async(std::launch::async, inflate_blocks, std::move(buf)));
				}
			}
			if (jobs.empty()) {
				break;
			}
			std::vector<char> chunk = jobs.front().get();
			jobs.pop_front();
			if (!chunk.empty() && !push(std::move(chunk))) {
				break;
			}
		}
		for (auto& j : jobs) {
			j.wait();
		}
		std::fclose(fp);
	}

	unsigned threads_;
	bool ok_ = false;
	std::string path_;
	Kind kind_ = PLAIN;
	std::thread producer_;
	std::mutex m_;
	std::condition_variable cv_data_, cv_space_;
	std::deque<std::vector<char>> ready_;
	bool done_ = false, stop_ = false;
};

} // namespace ntb
