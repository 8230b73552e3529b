// stand-in for <boost/config.hpp>.  The real Boost headers pull in most of the standard library and the reference
// relies on that (std::bitset, std::array, pow, sqrt, setw ... are used without their own #include).
#pragma once
#include <algorithm>
#include <array>
#include <atomic>
#include <bitset>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iomanip>
#include <iostream>
#include <memory>
#include <numeric>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>
using std::pow; using std::sqrt; using std::log; using std::exp; using std::floor; using std::ceil; using std::fabs;
