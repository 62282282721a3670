// TEST-ONLY stand-in for the OpenCV C++ headers (absent from this image): just enough declarations for
// `g++ -fsyntax-only` to type-check the zenslam_cuda/ adapter against the reference's own headers
// (tests/test_adapter_syntax.py).  Nothing here is compiled into the product.
#pragma once
#include <cstddef>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

using uchar = unsigned char;

#define CV_8UC1 0
#define CV_32FC1 5

namespace cv
{
    namespace Error { enum Code { StsError = -2, StsNotImplemented = -213, StsAssert = -215 }; }

    class Exception : public std::runtime_error { public: using std::runtime_error::runtime_error; };

    [[noreturn]] inline void error(int, const std::string& msg, const char*, const char*, int) { throw Exception(msg); }

    template <typename T> struct Point_ { T x { }, y { }; Point_() = default; Point_(T x_, T y_) : x(x_), y(y_) { } };
    using Point2f = Point_<float>;
    using Point   = Point_<int>;

    template <typename T> struct Size_
    {
        T width { }, height { };
        Size_() = default;
        Size_(T w, T h) : width(w), height(h) { }
        bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
    };
    using Size = Size_<int>;

    template <typename T, int M, int N> struct Matx { T val[M * N] { }; };
    using Matx33d = Matx<double, 3, 3>;

    struct TermCriteria
    {
        enum Type { COUNT = 1, MAX_ITER = COUNT, EPS = 2 };
        int type { }, maxCount { };
        double epsilon { };
    };

    struct Scalar { double v[4] { }; Scalar() = default; Scalar(double v0) { v[0] = v0; } };

    struct MatStep { size_t v { }; operator size_t() const { return v; } };

    class Mat
    {
    public:
        Mat() = default;
        Mat(int rows_, int cols_, int type_) : rows(rows_), cols(cols_), _type(type_) { }
        Mat(Size size_, int type_, const Scalar&) : rows(size_.height), cols(size_.width), _type(type_) { }
        int     rows { }, cols { };
        uchar*  data { };
        MatStep step { };
        [[nodiscard]] int  type() const { return _type; }
        [[nodiscard]] bool empty() const { return rows == 0 || cols == 0; }
        [[nodiscard]] bool isContinuous() const { return true; }
        [[nodiscard]] Size size() const { return { cols, rows }; }
        [[nodiscard]] Mat  clone() const { return *this; }
        [[nodiscard]] Mat  row(int) const { return *this; }
        [[nodiscard]] Mat  rowRange(int, int) const { return *this; }
    private:
        int _type { };
    };

    struct KeyPoint
    {
        Point2f pt { };
        float   size { }, angle { -1 }, response { };
        int     octave { }, class_id { -1 };
        KeyPoint() = default;
        KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1) :
            pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) { }
    };

    struct DMatch
    {
        int   queryIdx { -1 }, trainIdx { -1 }, imgIdx { -1 };
        float distance { };
        DMatch() = default;
        DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), distance(d) { }
        DMatch(int q, int t, int i, float d) : queryIdx(q), trainIdx(t), imgIdx(i), distance(d) { }
    };

    enum NormTypes { NORM_L2 = 4, NORM_HAMMING = 6 };

    template <typename T> using Ptr = std::shared_ptr<T>;
    template <typename T, typename... A> Ptr<T> makePtr(A&&... a) { return std::make_shared<T>(std::forward<A>(a)...); }

    class _InputArray
    {
    public:
        _InputArray() = default;
        _InputArray(const Mat& m) : _m(m) { }
        [[nodiscard]] Mat  getMat() const { return _m; }
        [[nodiscard]] bool empty() const { return _m.empty(); }
    private:
        Mat _m { };
    };
    using InputArray         = const _InputArray&;
    using InputArrayOfArrays = const _InputArray&;
}

#define CV_Error(code, msg) cv::error(code, msg, "", __FILE__, __LINE__)
#define CV_Assert(expr) do { if (!(expr)) cv::error(cv::Error::StsAssert, #expr, "", __FILE__, __LINE__); } while (0)
