#pragma once
#include <algorithm>
namespace boost { namespace sort {
template<class It, class C> inline void block_indirect_sort(It a, It b, C c) { std::sort(a, b, c); }
template<class It> inline void block_indirect_sort(It a, It b) { std::sort(a, b); }
} }
