#pragma once
namespace boost { namespace multiprecision { typedef unsigned __int128 uint128_t; } }
