// See bit_vectors.hpp: the reference includes this header but uses nothing else from it.
#pragma once
#include "bit_vectors.hpp"
