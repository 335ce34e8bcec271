#include "mrg5.hpp"
namespace trng { typedef mrg5 lcg64_shift; }
