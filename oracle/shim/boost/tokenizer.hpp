// Stand-in for <boost/tokenizer.hpp>, test infrastructure only.
// The verbatim libforest sources include it for CSVDataProvider
// (reference third-party/libforest/src/data.cpp:387-436), which is never on the
// per-keyframe inference path.  A comma split is all that call site needs.
#pragma once
#include <string>
#include <vector>
namespace boost {
template <class C> struct escaped_list_separator {};
template <class Sep> class tokenizer {
    std::vector<std::string> toks_;
public:
    typedef std::vector<std::string>::const_iterator iterator;
    explicit tokenizer(const std::string& s) {
        std::string cur;
        for (char ch : s) { if (ch == ',') { toks_.push_back(cur); cur.clear(); } else cur.push_back(ch); }
        toks_.push_back(cur);
    }
    iterator begin() const { return toks_.begin(); }
    iterator end() const { return toks_.end(); }
};
}
