// Stand-in for hclust-cpp's fastcluster.h (reference include/move_structure.hpp:25).
// Only Movi Color's tree compression (src/move_structure_color.cpp:232) calls it; that
// feature is outside the PML/count/ZML query path, so the stub aborts if reached.
#pragma once
#include <cstdio>
#include <cstdlib>
enum hclust_fast_methods { HCLUST_METHOD_SINGLE = 0, HCLUST_METHOD_COMPLETE = 1,
                           HCLUST_METHOD_AVERAGE = 2, HCLUST_METHOD_MEDIAN = 3 };
inline int hclust_fast(int, double*, int, int*, double*) {
    std::fprintf(stderr, "hclust_fast: not available in the oracle build\n");
    std::abort();
    return 1;
}
