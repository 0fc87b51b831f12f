// TEST INFRASTRUCTURE ONLY — empty-bodied stand-ins for the OpenCV 4.1.2 types
// that the reference translation units mention in code that is OUT of the hot
// path (SVM / RTrees / MLP / FLANN / PCA wrappers).  They exist solely so that
// /root/reference/qt_cpp/{db_features,ann,classification}.cpp compile verbatim
// into oracle/_ref/ without OpenCV being installed.  None of these stubs is
// ever executed by the oracle drivers: the hot-path arithmetic is all in-tree
// in the reference (plain loops), see SURVEY.md §8(c).
#ifndef FIR_ORACLE_OPENCV_STUB_CORE_HPP
#define FIR_ORACLE_OPENCV_STUB_CORE_HPP

#include <cstddef>
#include <string>
#include <vector>
#include <iostream>

#define CV_32F 5
#define CV_32S 4
#define CV_32FC1 5
#define CV_64F 6
#define CV_8U 0
#define CV_16U 2
#define CV_16S 3
#define CV_8S 1

// the reference calls qDebug() unconditionally once (db_features.cpp:292)
struct FirNullDebug {
    template <typename T> FirNullDebug& operator<<(const T&) { return *this; }
};
inline FirNullDebug qDebug() { return FirNullDebug(); }

namespace cv {

struct Range { int start, end; Range(int s = 0, int e = 0) : start(s), end(e) {} static Range all() { return Range(); } };
struct Size { int width, height; Size(int w = 0, int h = 0) : width(w), height(h) {} };
template <typename T> struct Scalar_ { Scalar_(T = T()) {} };
typedef Scalar_<double> Scalar;

class Mat {
public:
    int rows, cols;
    Mat() : rows(0), cols(0), dummy_(0) {}
    Mat(int r, int c, int) : rows(r), cols(c), dummy_(0) {}
    template <typename A> Mat(int r, int c, int, const A&) : rows(r), cols(c), dummy_(0) {}
    template <typename T> T& at(int) { return *reinterpret_cast<T*>(&dummy_); }
    template <typename T> T& at(int, int) { return *reinterpret_cast<T*>(&dummy_); }
    template <typename T> const T& at(int) const { return *reinterpret_cast<const T*>(&dummy_); }
    template <typename T> const T& at(int, int) const { return *reinterpret_cast<const T*>(&dummy_); }
    Mat row(int) const { return Mat(); }
    Mat col(int) const { return Mat(); }
    Mat rowRange(int, int) const { return Mat(); }
    Mat colRange(int, int) const { return Mat(); }
    Mat clone() const { return Mat(); }
    Mat t() const { return Mat(); }
    Mat reshape(int, int = 0) const { return Mat(); }
    template <typename A> void copyTo(A&) const {}
    template <typename A> void convertTo(A&, int, double = 1, double = 0) const {}
    bool empty() const { return true; }
    int type() const { return 0; }
    size_t total() const { return 0; }
    template <typename T> T* ptr(int = 0) { return reinterpret_cast<T*>(&dummy_); }
    static Mat zeros(int r, int c, int t) { return Mat(r, c, t); }
    static Mat ones(int r, int c, int t) { return Mat(r, c, t); }
    template <typename A> Mat& operator=(const Scalar_<A>&) { return *this; }
private:
    double dummy_;
};
inline Mat operator*(const Mat&, const Mat&) { return Mat(); }
inline Mat operator-(const Mat&, const Mat&) { return Mat(); }
inline Mat operator+(const Mat&, const Mat&) { return Mat(); }

template <typename T> class Ptr {
public:
    Ptr() : p_(0) {}
    Ptr(T* p) : p_(p) {}
    T* operator->() const { return p_; }
    T& operator*() const { return *p_; }
    operator T*() const { return p_; }
    T* get() const { return p_; }
    bool empty() const { return p_ == 0; }
private:
    T* p_;
};

struct TermCriteria {
    enum { COUNT = 1, MAX_ITER = 1, EPS = 2 };
    TermCriteria() {}
    TermCriteria(int, int, double) {}
};

class PCA {
public:
    enum { DATA_AS_ROW = 0, DATA_AS_COL = 1 };
    Mat eigenvectors, eigenvalues, mean;
    PCA() {}
    template <typename A, typename B> PCA(const A&, const B&, int, int = 0) {}
    template <typename A, typename B> PCA(const A&, const B&, int, double) {}
    template <typename A, typename B> PCA& operator()(const A&, const B&, int, int = 0) { return *this; }
    template <typename A> Mat project(const A&) const { return Mat(); }
    template <typename A, typename B> void project(const A&, B&) const {}
    template <typename A> Mat backProject(const A&) const { return Mat(); }
};
#define CV_PCA_DATA_AS_ROW 0

}  // namespace cv

#endif
