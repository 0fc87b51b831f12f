// TEST INFRASTRUCTURE ONLY — see core.hpp in this directory.
#ifndef FIR_ORACLE_OPENCV_STUB_ML_HPP
#define FIR_ORACLE_OPENCV_STUB_ML_HPP
#include "core.hpp"
namespace cv { namespace ml {
enum SampleTypes { ROW_SAMPLE = 0, COL_SAMPLE = 1 };
class StatModel {
public:
    virtual ~StatModel() {}
    template <typename A, typename B> bool train(const A&, int, const B&) { return false; }
    template <typename A> float predict(const A&) const { return 0.f; }
    template <typename A, typename B> float predict(const A&, B&, int = 0) const { return 0.f; }
};
class SVM : public StatModel {
public:
    enum Types { C_SVC = 100, NU_SVC, ONE_CLASS, EPS_SVR, NU_SVR };
    enum KernelTypes { CUSTOM = -1, LINEAR = 0, POLY, RBF, SIGMOID, CHI2, INTER };
    static Ptr<SVM> create() { return Ptr<SVM>(new SVM()); }
    void setType(int) {}
    void setKernel(int) {}
    void setGamma(double) {}
    void setC(double) {}
    void setDegree(double) {}
    void setCoef0(double) {}
    void setNu(double) {}
    void setP(double) {}
    void setTermCriteria(const TermCriteria&) {}
};
class RTrees : public StatModel {
public:
    static Ptr<RTrees> create() { return Ptr<RTrees>(new RTrees()); }
    void setMaxDepth(int) {}
    void setMaxCategories(int) {}
    void setMinSampleCount(int) {}
    void setRegressionAccuracy(float) {}
    void setUseSurrogates(bool) {}
    template <typename A> void setPriors(const A&) {}
    void setCalculateVarImportance(bool) {}
    void setActiveVarCount(int) {}
    void setTermCriteria(const TermCriteria&) {}
};
class ANN_MLP : public StatModel {
public:
    enum TrainingMethods { BACKPROP = 0, RPROP = 1, ANNEAL = 2 };
    enum ActivationFunctions { IDENTITY = 0, SIGMOID_SYM = 1, GAUSSIAN = 2, RELU = 3, LEAKYRELU = 4 };
    static Ptr<ANN_MLP> create() { return Ptr<ANN_MLP>(new ANN_MLP()); }
    void setTermCriteria(const TermCriteria&) {}
    void setTrainMethod(int, double = 0, double = 0) {}
    template <typename A> void setLayerSizes(const A&) {}
    void setActivationFunction(int, double = 0, double = 0) {}
};
}}  // namespace cv::ml
#endif
