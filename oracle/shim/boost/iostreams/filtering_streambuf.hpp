// ORACLE ONLY: stub so the read-only reference compiles without boost.  The
// reference uses boost::iostreams solely to read a gzipped ClinVar VCF (-l
// *.gz, ntedit.cpp:2535-2550), which is annotation-only and out of scope.
#ifndef ORACLE_SHIM_BOOST_FILTERING_STREAMBUF_HPP
#define ORACLE_SHIM_BOOST_FILTERING_STREAMBUF_HPP
#include <streambuf>
namespace boost {
namespace iostreams {
struct input
{};
struct gzip_decompressor
{};
template<typename Mode>
class filtering_streambuf : public std::streambuf
{
  public:
	template<typename T>
	void push(const T&)
	{}
};
} // namespace iostreams
} // namespace boost
#endif
