// TEST-ONLY stand-in for the OpenCV C++ headers (absent from this image).  It is FUNCTIONAL, not declaration-only: a
// reference-counted cv::Mat with data / step / ROI views, cv::_InputArray with OpenCV's MAT and STD_VECTOR_MAT kinds
// (including the semantics of empty() on a vector of Mats), cv::KeyPoint / DMatch / Point / Size / TermCriteria -- the part of
// the OpenCV core API that the zenslam_cuda/ adapter and the reference headers it includes use.  With it the adapter is
//   * type-checked against the reference's own headers (tests/test_adapter_syntax.py), and
//   * compiled, LINKED against libzenslam_cuda.so and RUN (tests/adapter/adapter_harness.cpp, tests/test_gpu_adapter.py).
// Nothing here is compiled into the product.  Behaviour follows the public OpenCV 4.x documentation.
#pragma once
#include <algorithm>
#include <cmath>
#include <type_traits>
#include <cstddef>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

using uchar = unsigned char;

#define CV_8U 0
#define CV_32F 5
#define CV_8UC1 0
#define CV_32FC1 5
#define CV_8UC3 16

namespace cv
{
    namespace Error { enum Code { StsError = -2, StsBadArg = -5, StsNotImplemented = -213, StsAssert = -215 }; }

    class Exception : public std::runtime_error
    {
    public:
        Exception(int code_, const std::string& msg) : std::runtime_error(msg), code(code_) { }
        explicit Exception(const std::string& msg) : std::runtime_error(msg) { }
        int code { Error::StsError };
    };

    [[noreturn]] inline void error(int code, const std::string& msg, const char*, const char* file, int line)
    {
        throw Exception(code, std::string(file) + ":" + std::to_string(line) + ": " + msg);
    }

    template <typename T> struct Point_
    {
        T x { }, y { };
        Point_() = default;
        Point_(T x_, T y_) : x(x_), y(y_) { }
        // OpenCV converts with saturate_cast: float -> int is cvRound (round half to even), which is what cv::circle's
        // `Point center` parameter does to the reference's Point2f keypoint positions (keypoint_detector_simple.cpp:45)
        template <typename U> Point_(const Point_<U>& o) : x(convert(o.x)), y(convert(o.y)) { }
        bool operator==(const Point_& o) const { return x == o.x && y == o.y; }

    private:
        template <typename U> static T convert(U v)
        {
            if constexpr (std::is_integral_v<T> && std::is_floating_point_v<U>) return static_cast<T>(std::lrint(v));
            else return static_cast<T>(v);
        }
    };
    using Point2f = Point_<float>;
    using Point2d = Point_<double>;
    using Point   = Point_<int>;

    template <typename T> struct Point3_
    {
        T x { }, y { }, z { };
        Point3_() = default;
        Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) { }
    };
    using Point3d = Point3_<double>;
    using Point3f = Point3_<float>;

    template <typename T> struct Size_
    {
        T width { }, height { };
        Size_() = default;
        Size_(T w, T h) : width(w), height(h) { }
        bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
        [[nodiscard]] T area() const { return width * height; }
    };
    using Size = Size_<int>;

    template <typename T> struct Rect_
    {
        T x { }, y { }, width { }, height { };
        Rect_() = default;
        Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) { }
    };
    using Rect = Rect_<int>;

    template <typename T, int M, int N> struct Matx
    {
        T val[M * N] { };
        T&       operator()(int r, int c) { return val[r * N + c]; }
        const T& operator()(int r, int c) const { return val[r * N + c]; }
    };
    using Matx33d = Matx<double, 3, 3>;
    using Matx34d = Matx<double, 3, 4>;
    using Vec3d   = Matx<double, 3, 1>;

    struct TermCriteria
    {
        enum Type { COUNT = 1, MAX_ITER = COUNT, EPS = 2 };
        int    type { }, maxCount { };
        double epsilon { };
        TermCriteria() = default;
        TermCriteria(int type_, int max_count, double eps) : type(type_), maxCount(max_count), epsilon(eps) { }
    };

    struct Scalar
    {
        double v[4] { };
        Scalar() = default;
        Scalar(double v0) { v[0] = v0; }
        double operator[](int i) const { return v[i]; }
    };

    struct MatStep
    {
        size_t v { };
        MatStep() = default;
        MatStep(size_t s) : v(s) { }
        operator size_t() const { return v; }
    };

    // 2-D matrix of CV_8UC1 / CV_32FC1 / CV_8UC3 elements: shared buffer + (data, step) view, so ROIs, row() and rowRange() alias
    // the parent's memory exactly like OpenCV's headers do
    class Mat
    {
    public:
        enum { AUTO_STEP = 0 };

        Mat() = default;
        Mat(int rows_, int cols_, int type_) { create(rows_, cols_, type_); }
        Mat(Size size_, int type_) { create(size_.height, size_.width, type_); }
        Mat(Size size_, int type_, const Scalar& s) { create(size_.height, size_.width, type_); setTo(s); }
        Mat(int rows_, int cols_, int type_, const Scalar& s) { create(rows_, cols_, type_); setTo(s); }
        // user-owned memory (no copy, no ownership), like cv::Mat(rows, cols, type, void*, step)
        Mat(int rows_, int cols_, int type_, void* data_, size_t step_ = AUTO_STEP) :
            rows(rows_), cols(cols_), data(static_cast<uchar*>(data_)), step(step_ ? step_ : cols_ * elem(type_)), _type(type_) { }

        void create(int rows_, int cols_, int type_)
        {
            rows = rows_; cols = cols_; _type = type_;
            step = static_cast<size_t>(cols_) * elem(type_);
            _buffer = std::make_shared<std::vector<uchar>>(static_cast<size_t>(rows_) * step.v);
            data = _buffer->data();
        }

        void setTo(const Scalar& s)
        {
            for (int r = 0; r < rows; ++r)
            {
                if (_type == CV_8UC1 || _type == CV_8UC3) std::memset(data + r * step.v, static_cast<int>(s[0]), cols * elem(_type));
                else for (int c = 0; c < cols; ++c) reinterpret_cast<float*>(data + r * step.v)[c] = static_cast<float>(s[0]);
            }
        }

        int     rows { }, cols { };
        uchar*  data { };
        MatStep step { };

        [[nodiscard]] int    type() const { return _type; }
        [[nodiscard]] size_t elemSize() const { return elem(_type); }
        [[nodiscard]] bool   empty() const { return rows == 0 || cols == 0 || data == nullptr; }
        [[nodiscard]] bool   isContinuous() const { return rows <= 1 || step.v == static_cast<size_t>(cols) * elem(_type); }
        [[nodiscard]] Size   size() const { return { cols, rows }; }

        [[nodiscard]] Mat clone() const
        {
            Mat m;
            if (empty()) return m;
            m.create(rows, cols, _type);
            for (int r = 0; r < rows; ++r) std::memcpy(m.data + r * m.step.v, data + r * step.v, static_cast<size_t>(cols) * elem(_type));
            return m;
        }

        [[nodiscard]] Mat operator()(const Rect& roi) const
        {
            Mat m = *this;
            m.rows = roi.height; m.cols = roi.width;
            m.data = data + roi.y * step.v + static_cast<size_t>(roi.x) * elem(_type);
            return m;
        }
        [[nodiscard]] Mat rowRange(int r0, int r1) const { return (*this)(Rect(0, r0, cols, r1 - r0)); }
        [[nodiscard]] Mat row(int r) const { return rowRange(r, r + 1); }

        // cv::Mat::ones(size, type) * 255 (keypoint_detector_simple.cpp:41): the product of the all-ones expression is a filled Mat
        struct ones_expr { int rows, cols, type; };
        static ones_expr ones(Size size_, int type_) { return { size_.height, size_.width, type_ }; }
        Mat(const ones_expr& e) { create(e.rows, e.cols, e.type); setTo(Scalar(1)); }
        friend Mat operator*(const ones_expr& e, double v) { return Mat(e.rows, e.cols, e.type, Scalar(v)); }

        template <typename M> void copyTo(M& dst) const { dst = clone(); }

        template <typename T> [[nodiscard]] T*       ptr(int r = 0) { return reinterpret_cast<T*>(data + r * step.v); }
        template <typename T> [[nodiscard]] const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + r * step.v); }
        template <typename T> [[nodiscard]] T&       at(int r, int c) { return ptr<T>(r)[c]; }
        template <typename T> [[nodiscard]] const T& at(int r, int c) const { return ptr<T>(r)[c]; }

    private:
        static size_t elem(int type_) { return type_ == CV_32FC1 ? 4 : type_ == CV_8UC3 ? 3 : 1; }
        int                                 _type { };
        std::shared_ptr<std::vector<uchar>> _buffer { };
    };

    // cv::UMat as the reference uses it (keypoint_detector_parallel.cpp:178-180): a Mat that lives somewhere else
    class UMat : public Mat
    {
    public:
        UMat() = default;
        UMat& operator=(const Mat& m) { static_cast<Mat&>(*this) = m; return *this; }
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

    // cv::_InputArray of kind NONE / MAT / STD_VECTOR_MAT.  As in OpenCV, a vector of Mats is "empty" only when the VECTOR is
    // empty -- std::vector<Mat>(1, Mat()) is NOT empty -- which is what cv::DescriptorMatcher::knnMatch(query, train, ...)
    // hands to knnMatchImpl as its masks argument.
    class _InputArray
    {
    public:
        enum Kind { NONE, MAT, STD_VECTOR_MAT };

        _InputArray() = default;
        _InputArray(const Mat& m) : _kind(MAT), _m(m) { }
        _InputArray(const std::vector<Mat>& v) : _kind(STD_VECTOR_MAT), _v(v) { }

        [[nodiscard]] Kind kind() const { return _kind; }
        [[nodiscard]] Mat  getMat(int i = -1) const
        {
            if (_kind == MAT) return _m;
            if (_kind == STD_VECTOR_MAT && i >= 0 && i < static_cast<int>(_v.size())) return _v[i];
            return Mat();
        }
        void getMatVector(std::vector<Mat>& out) const
        {
            if (_kind == STD_VECTOR_MAT) out = _v;
            else if (_kind == MAT) out.assign(1, _m);
            else out.clear();
        }
        [[nodiscard]] bool empty() const
        {
            if (_kind == MAT) return _m.empty();
            if (_kind == STD_VECTOR_MAT) return _v.empty();
            return true;
        }
        [[nodiscard]] Size size() const { return _kind == MAT ? _m.size() : Size(static_cast<int>(_v.size()), 1); }

    private:
        Kind             _kind { NONE };
        Mat              _m { };
        std::vector<Mat> _v { };
    };
    using InputArray         = const _InputArray&;
    using InputArrayOfArrays = const _InputArray&;
    inline _InputArray noArray() { return _InputArray(); }

    // cv::OutputArray bound to a cv::Mat the callee assigns
    class _OutputArray
    {
    public:
        _OutputArray(Mat& m) : _m(&m) { }
        void assign(const Mat& m) const { *_m = m; }
    private:
        Mat* _m;
    };
    using OutputArray = const _OutputArray&;
}

#define CV_Error(code, msg) cv::error(code, msg, "", __FILE__, __LINE__)
#define CV_Assert(expr) do { if (!(expr)) cv::error(cv::Error::StsAssert, #expr, "", __FILE__, __LINE__); } while (0)
