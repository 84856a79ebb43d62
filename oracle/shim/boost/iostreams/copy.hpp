// ORACLE ONLY: empty stub (see filtering_streambuf.hpp)
