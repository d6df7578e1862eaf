// Stand-in for Mosek Fusion's "fusion.h".  In the compiled NLPClass_sqp.cpp all
// Mosek code is commented out (NLPClass_sqp.cpp:1511-1610); only the include,
// two using-directives and two helper one-liners (:23-29) still name it.
// Test infrastructure only; no optimisation code here.
#pragma once
#include <memory>
#include <vector>
namespace monty {
template <class T, int N> struct ndarray {};
template <class T> std::shared_ptr<ndarray<T, 1>> new_array_ptr(const std::vector<T>&) { return std::shared_ptr<ndarray<T, 1>>(); }
}  // namespace monty
namespace mosek { namespace fusion {} }
