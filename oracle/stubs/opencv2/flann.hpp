// TEST INFRASTRUCTURE ONLY — see core.hpp in this directory.
#ifndef FIR_ORACLE_OPENCV_STUB_FLANN_HPP
#define FIR_ORACLE_OPENCV_STUB_FLANN_HPP
#include "core.hpp"
namespace cvflann {
template <typename T> struct L2 { typedef T ElementType; typedef float ResultType; };
template <typename T> struct ChiSquareDistance { typedef T ElementType; typedef float ResultType; };
template <typename T> struct Matrix {
    Matrix() {}
    Matrix(T*, size_t, size_t) {}
};
struct KDTreeIndexParams { KDTreeIndexParams(int = 4) {} };
struct SearchParams { SearchParams(int = 32, float = 0, bool = true) {} };
template <typename D> class Index {
public:
    template <typename P> Index(const Matrix<typename D::ElementType>&, const P&) {}
    void buildIndex() {}
    void knnSearch(const Matrix<typename D::ElementType>&, Matrix<int>&, Matrix<float>&, int, const SearchParams&) {}
};
}  // namespace cvflann
#endif
