#include "mrg5.hpp"
