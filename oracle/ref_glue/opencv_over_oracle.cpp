// TEST INFRASTRUCTURE (oracle/ref_glue).  The OpenCV calls the reference's detector sources make -- cv::FastFeatureDetector,
// cv::ORB (detect and compute), cv::cornerSubPix -- implemented over the C oracle (oracle/zs_oracle.c), whose arithmetic is
// pinned to real cv2 outputs (tests/golden, tests/test_oracle_vs_cv2.py).  Together with the functional OpenCV stand-in of
// tests/stubs/ this lets the reference's OWN glue -- keypoint_detector_grid.cpp, keypoint_detector_parallel.cpp,
// keypoint_detector_simple.cpp, compiled unmodified from /root/reference -- run here: oracle/_ref/libzs_ref_glue.so.
// tests/test_oracle_vs_reference_glue.py then checks that the oracle's restatement of that glue (zso_grid_detect and the
// python glue in oracle/__init__.py) returns exactly what the reference's code returns on the same primitives.
#include <cstdint>
#include <vector>

#include <opencv2/core.hpp>
#include <opencv2/features2d.hpp>
#include <opencv2/imgproc.hpp>

extern "C"
{
    int  zso_fast_detect(const uint8_t* img, int w, int h, int pitch, int threshold, int* xs, int* ys, int* scores, int cap);
    int  zso_orb_filter(const float* xs, const float* ys, int n, int w, int h, int* kept);
    void zso_orb_blur(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch);
    void zso_orb_describe(const uint8_t* blurred, int w, int h, int pitch, const float* xs, const float* ys, const float* angles, int n, uint8_t* desc);
    int  zso_orb_detect(const uint8_t* img, int w, int h, int pitch, const uint8_t* mask, int mpitch, int nfeatures, float scale_factor,
                        int nlevels, int edge, int patch, int fast_threshold, float* ox, float* oy, float* osize, float* oangle,
                        float* oresp, int* ooct, uint8_t* desc, int cap);
    void zso_corner_subpix(const uint8_t* img, int w, int h, int pitch, float* xy, int n, int win_w, int win_h, int max_iters, double eps);
}

namespace cv
{
    void Feature2D::detect(InputArray, std::vector<KeyPoint>&, InputArray) { CV_Error(Error::StsNotImplemented, "Feature2D::detect"); }
    void Feature2D::compute(InputArray, std::vector<KeyPoint>&, OutputArray) { CV_Error(Error::StsNotImplemented, "Feature2D::compute"); }

    Ptr<FastFeatureDetector> FastFeatureDetector::create(const int threshold, const bool nonmax)
    {
        CV_Assert(nonmax);
        auto p        = makePtr<FastFeatureDetector>();
        p->_threshold = threshold;
        return p;
    }

    // cv::FAST(image, threshold, true, TYPE_9_16): raster order, size 7, angle -1, response = score, then the mask
    void FastFeatureDetector::detect(InputArray image, std::vector<KeyPoint>& keypoints, InputArray mask)
    {
        const Mat img = image.getMat();
        const Mat msk = mask.getMat();
        keypoints.clear();
        CV_Assert(img.type() == CV_8UC1);
        if (img.empty()) return;
        const int        cap = std::max(1, img.cols * img.rows);
        std::vector<int> xs(cap), ys(cap), sc(cap);
        const int        n = zso_fast_detect(img.data, img.cols, img.rows, static_cast<int>(img.step), _threshold, xs.data(), ys.data(), sc.data(), cap);
        for (int i = 0; i < n; ++i)
        {
            if (!msk.empty() && msk.at<uchar>(ys[i], xs[i]) == 0) continue;
            keypoints.emplace_back(static_cast<float>(xs[i]), static_cast<float>(ys[i]), 7.f, -1.f, static_cast<float>(sc[i]), 0, -1);
        }
    }

    Ptr<ORB> ORB::create(const int nfeatures, const float scale, const int nlevels, const int edge, const int first_level, const int wta_k,
                         const ScoreType score, const int patch, const int fast_threshold)
    {
        CV_Assert(first_level == 0 && wta_k == 2 && score == HARRIS_SCORE);
        auto p            = makePtr<ORB>();
        p->_nfeatures     = nfeatures; p->_scale = scale; p->_nlevels = nlevels; p->_edge = edge; p->_patch = patch;
        p->_fast_threshold = fast_threshold;
        return p;
    }

    void ORB::detect(InputArray image, std::vector<KeyPoint>& keypoints, InputArray mask)
    {
        const Mat img = image.getMat();
        const Mat msk = mask.getMat();
        keypoints.clear();
        if (img.empty()) return;
        const int          cap = _nfeatures + 64 * _nlevels + 1024;
        std::vector<float> x(cap), y(cap), size(cap), angle(cap), resp(cap);
        std::vector<int>   oct(cap);
        const Mat          mc = msk.empty() || msk.isContinuous() ? msk : msk.clone();
        int n = zso_orb_detect(img.data, img.cols, img.rows, static_cast<int>(img.step), mc.empty() ? nullptr : mc.data, static_cast<int>(mc.step),
                               _nfeatures, _scale, _nlevels, _edge, _patch, _fast_threshold, x.data(), y.data(), size.data(), angle.data(),
                               resp.data(), oct.data(), nullptr, cap);
        n = std::min(n, cap);
        for (int i = 0; i < n; ++i) keypoints.emplace_back(x[i], y[i], size[i], angle[i], resp[i], oct[i], -1);
    }

    // cv::ORB::compute for keypoints that carry their own angle (FAST keypoints: -1): border filter that REMOVES keypoints from
    // the vector (order kept), 7x7 sigma-2 blur, rotated rBRIEF
    void ORB::compute(InputArray image, std::vector<KeyPoint>& keypoints, OutputArray descriptors)
    {
        Mat img = image.getMat();
        if (!img.isContinuous()) img = img.clone();
        const int          n = static_cast<int>(keypoints.size());
        std::vector<float> xs(n), ys(n);
        for (int i = 0; i < n; ++i) { xs[i] = keypoints[i].pt.x; ys[i] = keypoints[i].pt.y; }
        std::vector<int> kept(std::max(1, n));
        const int        m = n ? zso_orb_filter(xs.data(), ys.data(), n, img.cols, img.rows, kept.data()) : 0;
        std::vector<KeyPoint> out;
        std::vector<float>    kx(m), ky(m), ka(m);
        for (int i = 0; i < m; ++i)
        {
            out.push_back(keypoints[kept[i]]);
            kx[i] = out.back().pt.x; ky[i] = out.back().pt.y; ka[i] = out.back().angle;
        }
        keypoints = out;
        Mat desc(m, 32, CV_8UC1);
        if (m > 0)
        {
            Mat blurred(img.rows, img.cols, CV_8UC1);
            zso_orb_blur(img.data, img.cols, img.rows, static_cast<int>(img.step), blurred.data, static_cast<int>(blurred.step));
            zso_orb_describe(blurred.data, img.cols, img.rows, static_cast<int>(blurred.step), kx.data(), ky.data(), ka.data(), m, desc.data);
        }
        descriptors.assign(desc);
    }

    Ptr<SIFT> SIFT::create() { CV_Error(Error::StsNotImplemented, "SIFT is not part of the oracle"); }

    void cornerSubPix(InputArray image, std::vector<Point2f>& corners, const Size win, const Size zero_zone, const TermCriteria criteria)
    {
        Mat img = image.getMat();
        CV_Assert(img.type() == CV_8UC1 && zero_zone.width < 0 && zero_zone.height < 0);
        CV_Assert((criteria.type & TermCriteria::COUNT) && (criteria.type & TermCriteria::EPS));
        if (corners.empty()) return;
        static_assert(sizeof(Point2f) == 2 * sizeof(float));
        zso_corner_subpix(img.data, img.cols, img.rows, static_cast<int>(img.step), reinterpret_cast<float*>(corners.data()),
                          static_cast<int>(corners.size()), win.width, win.height, criteria.maxCount, criteria.epsilon);
    }
}
