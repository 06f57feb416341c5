// Minimal stand-in for the three Boost string helpers the reference's
// parameter parser uses (include/params.h, include/physics/eforce.h).
#pragma once
#include <cctype>
#include <cstring>
#include <string>
#include <vector>
namespace boost {
inline bool iequals(const std::string & a, const std::string & b) {
  if(a.size() != b.size()) return false;
  for(size_t i = 0; i < a.size(); ++i)
    if(std::tolower((unsigned char)a[i]) != std::tolower((unsigned char)b[i])) return false;
  return true;
}
struct is_any_of { std::string set; is_any_of(const char * s) : set(s) {}
  bool operator()(char c) const { return set.find(c) != std::string::npos; } };
template<class P>
inline void split(std::vector<std::string> & out, const std::string & s, P pred) {
  out.clear(); std::string cur;
  for(char c : s) { if(pred(c)) { out.push_back(cur); cur.clear(); } else cur.push_back(c); }
  out.push_back(cur);
}
}
