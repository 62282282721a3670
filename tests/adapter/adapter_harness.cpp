// TEST-ONLY executable: the zenslam_cuda/ C++ adapter, compiled against the REFERENCE's own headers (pyr_lk.h,
// keypoint_detector.h, detection_options.h, tracking_options.h, keypoint.h, map.h) and the functional OpenCV stand-in under
// tests/stubs/, LINKED to libzenslam_cuda.so and RUN.  It drives the adapter exactly through the reference's seams --
//   zenslam::pyr_lk::calc_optical_flow_pyr_lk            (zenslam_core/include/zenslam/tracking/pyr_lk.h:15-26)
//   zenslam::keypoint_detector::detect_keypoints          (zenslam_core/include/zenslam/detection/keypoint_detector.h:13)
//   cv::DescriptorMatcher::knnMatch / match               (the calls of matcher.cpp:65,79)
//   zenslam::cuda::stereo_tracker::track                  (the body of keypoint_tracker.cpp:41-105)
// -- on arrays written by tests/test_gpu_adapter.py and dumps every result for that test to compare with the oracle.
//
//   adapter_harness <in-file> <out-file>
//
// File format (both ways): records of  "<name> <dtype> <ndim> <d0> ... \n"  followed by the raw little-endian bytes.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include <opencv2/core.hpp>
#include <opencv2/features2d.hpp>

#include "zenslam/detection/detection_options.h"
#include "zenslam/tracking_options.h"
#include "zenslam/types/keypoint.h"
#include "zenslam/types/map.h"

#include "zenslam_cuda/bf_matcher.h"
#include "zenslam_cuda/keypoint_detector_cuda.h"
#include "zenslam_cuda/processing.h"
#include "zenslam_cuda/pyr_lk.h"
#include "zenslam_cuda/pyr_lk_factory.h"
#include "zenslam_cuda/stereo_tracker.h"

// zenslam_core defines this in types/keypoint.cpp; the harness stands in for zenslam_core
size_t zenslam::keypoint::index_next = 0;

namespace
{
    struct blob
    {
        std::string         dtype;
        std::vector<size_t> shape;
        std::vector<char>   bytes;

        template <typename T> const T* as() const { return reinterpret_cast<const T*>(bytes.data()); }
        [[nodiscard]] size_t           dim(size_t i) const { return shape.at(i); }
    };

    auto read_all(const std::string& path) -> std::map<std::string, blob>
    {
        std::map<std::string, blob> out;
        std::ifstream               in(path, std::ios::binary);
        if (!in) throw std::runtime_error("cannot open " + path);
        std::string line;
        while (std::getline(in, line))
        {
            if (line.empty()) continue;
            std::istringstream hdr(line);
            std::string        name;
            blob               b;
            size_t             ndim = 0;
            hdr >> name >> b.dtype >> ndim;
            size_t count = 1;
            for (size_t i = 0; i < ndim; ++i) { size_t d = 0; hdr >> d; b.shape.push_back(d); count *= d; }
            const size_t item = (b.dtype == "u1") ? 1 : (b.dtype == "f8" || b.dtype == "i8") ? 8 : 4;
            b.bytes.resize(count * item);
            in.read(b.bytes.data(), static_cast<std::streamsize>(b.bytes.size()));
            in.get();        // the newline after the payload
            out[name] = std::move(b);
        }
        return out;
    }

    class writer
    {
    public:
        explicit writer(const std::string& path) : _out(path, std::ios::binary) { }

        template <typename T> void put(const std::string& name, const char* dtype, const std::vector<size_t>& shape, const T* data)
        {
            size_t count = 1;
            _out << name << ' ' << dtype << ' ' << shape.size();
            for (const auto d : shape) { _out << ' ' << d; count *= d; }
            _out << '\n';
            _out.write(reinterpret_cast<const char*>(data), static_cast<std::streamsize>(count * sizeof(T)));
            _out << '\n';
        }
        void put_f32(const std::string& name, const std::vector<float>& v, size_t cols = 1) { put(name, "f4", { v.size() / cols, cols }, v.data()); }
        void put_i32(const std::string& name, const std::vector<int>& v, size_t cols = 1) { put(name, "i4", { v.size() / cols, cols }, v.data()); }
        void put_u8(const std::string& name, const std::vector<uchar>& v, size_t cols = 1) { put(name, "u1", { v.size() / cols, cols }, v.data()); }

    private:
        std::ofstream _out;
    };

    // what utils::pyramid hands to pyr_lk (utils_opencv.cpp:525-530): cv::buildOpticalFlowPyramid's level 0 is an ROI inside a
    // buffer padded by the LK window, so data / step do not describe a continuous image.  The border content is irrelevant to
    // the adapter (the device rebuilds the pyramid); it is filled with a constant the results must not depend on.
    auto padded_level0(const uchar* pixels, const int w, const int h, const int pad, const uchar fill) -> std::vector<cv::Mat>
    {
        cv::Mat buffer(h + 2 * pad, w + 2 * pad, CV_8UC1, cv::Scalar(fill));
        cv::Mat roi = buffer(cv::Rect(pad, pad, w, h));
        for (int y = 0; y < h; ++y) std::memcpy(roi.ptr<uchar>(y), pixels + static_cast<size_t>(y) * w, w);
        return { roi };
    }

    void dump_keypoints(writer& out, const std::string& prefix, const std::vector<zenslam::keypoint>& keypoints)
    {
        std::vector<float> f;
        std::vector<int>   i;
        std::vector<uchar> d;
        for (const auto& k : keypoints)
        {
            f.insert(f.end(), { k.pt.x, k.pt.y, k.response, k.size, k.angle });
            i.insert(i.end(), { static_cast<int>(k.index), k.octave, k.class_id });
            CV_Assert(k.descriptor.rows == 1 && k.descriptor.cols == 32 && k.descriptor.type() == CV_8UC1);
            d.insert(d.end(), k.descriptor.ptr<uchar>(0), k.descriptor.ptr<uchar>(0) + 32);
        }
        out.put_f32(prefix + ".f", f, 5);
        out.put_i32(prefix + ".i", i, 3);
        out.put_u8(prefix + ".desc", d, 32);
    }

    void dump_map(writer& out, const std::string& prefix, const zenslam::map<zenslam::keypoint>& keypoints)
    {
        std::vector<zenslam::keypoint> flat;
        for (const auto& [index, k] : keypoints) flat.push_back(k);      // std::map order = ascending index
        dump_keypoints(out, prefix, flat);
    }

    void dump_knn(writer& out, const std::string& prefix, const std::vector<std::vector<cv::DMatch>>& rows)
    {
        std::vector<int>   len, idx;
        std::vector<float> dist;
        for (const auto& row : rows)
        {
            len.push_back(static_cast<int>(row.size()));
            for (const auto& m : row) { idx.insert(idx.end(), { m.queryIdx, m.trainIdx, m.imgIdx }); dist.push_back(m.distance); }
        }
        out.put_i32(prefix + ".len", len);
        out.put_i32(prefix + ".idx", idx, 3);
        out.put_f32(prefix + ".dist", dist);
    }
}

int main(const int argc, char** argv)
try
{
    if (argc != 3) { std::cerr << "usage: adapter_harness <in> <out>\n"; return 2; }

    const auto in = read_all(argv[1]);
    writer     out(argv[2]);

    const auto* dims   = in.at("dims").as<int>();           // w, h, frames, cell, fast_threshold, win, max_level
    const int   w      = dims[0], h = dims[1], frames = dims[2], cell = dims[3], threshold = dims[4], win = dims[5], max_level = dims[6];
    const auto* pixels = in.at("frames").as<uchar>();       // [frames][2][h][w]
    auto        image  = [&](const int t, const int camera) { return pixels + (static_cast<size_t>(t) * 2 + camera) * w * h; };

    // ---------------------------------------------------------------- the pyr_lk seam
    const auto lk = zenslam::cuda::create_cuda_pyr_lk();
    if (!lk) { std::cerr << "create_cuda_pyr_lk() returned an empty pointer: no sm_100 device\n"; return 3; }
    {
        const std::shared_ptr<zenslam::pyr_lk> seam = lk;    // used through the reference's base class only
        const auto&                            pts  = in.at("lk_points");
        const size_t                           n    = pts.dim(0);
        std::vector<cv::Point2f>               prev(n), next;
        for (size_t i = 0; i < n; ++i) prev[i] = { pts.as<float>()[2 * i], pts.as<float>()[2 * i + 1] };

        const auto pyramid_0 = padded_level0(image(0, 0), w, h, win, 17);
        const auto pyramid_1 = padded_level0(image(1, 0), w, h, win, 201);
        const cv::TermCriteria criteria { cv::TermCriteria::COUNT | cv::TermCriteria::EPS, 99, 0.001 };     // keypoint_tracker.cpp:150

        std::vector<uchar> status;
        std::vector<float> err;
        seam->calc_optical_flow_pyr_lk(pyramid_0, pyramid_1, prev, next, status, err, cv::Size(win, win), max_level, criteria, 8 /* OPTFLOW_LK_GET_MIN_EIGENVALS */, 1e-4);
        out.put("lk.next", "f4", { n, 2 }, reinterpret_cast<const float*>(next.data()));
        out.put_u8("lk.status", status);
        out.put_f32("lk.err", err);

        // OPTFLOW_USE_INITIAL_FLOW: next_points is in/out (keypoint_tracker.cpp:361-391)
        const auto&              guess = in.at("lk_initial");
        std::vector<cv::Point2f> next_init(n);
        for (size_t i = 0; i < n; ++i) next_init[i] = { guess.as<float>()[2 * i], guess.as<float>()[2 * i + 1] };
        seam->calc_optical_flow_pyr_lk(pyramid_0, pyramid_1, prev, next_init, status, err, cv::Size(win, win), max_level, criteria, 8 | 4, 1e-4);
        out.put("lk_init.next", "f4", { n, 2 }, reinterpret_cast<const float*>(next_init.data()));
        out.put_u8("lk_init.status", status);

        // a continuous next image against a padded previous one: the adapter has to equalise the row pitches
        const std::vector<cv::Mat> continuous_1 { cv::Mat(h, w, CV_8UC1, const_cast<uchar*>(image(1, 0))) };
        std::vector<cv::Point2f>   next_mixed;
        seam->calc_optical_flow_pyr_lk(pyramid_0, continuous_1, prev, next_mixed, status, err, cv::Size(win, win), max_level, criteria, 8, 1e-4);
        out.put("lk_mixed.next", "f4", { n, 2 }, reinterpret_cast<const float*>(next_mixed.data()));

        // the fused shortcut: forward + backward + gate
        std::vector<cv::Point2f> fb_points;
        std::vector<uchar>       fb_keep;
        zenslam::cuda::track_keypoints_fb(pyramid_0, pyramid_1, prev, { }, fb_points, fb_keep, cv::Size(win, win), max_level, 1.0);
        out.put("lk_fb.next", "f4", { n, 2 }, reinterpret_cast<const float*>(fb_points.data()));
        out.put_u8("lk_fb.keep", fb_keep);

        // no points: outputs are sized to zero and nothing is launched
        std::vector<cv::Point2f> none, none_next { { 1, 2 } };
        seam->calc_optical_flow_pyr_lk(pyramid_0, pyramid_1, none, none_next, status, err, cv::Size(win, win), max_level, criteria, 8, 1e-4);
        const std::vector<int> empty_sizes { static_cast<int>(none_next.size()), static_cast<int>(status.size()), static_cast<int>(err.size()) };
        out.put_i32("lk_empty.sizes", empty_sizes);
    }

    // ---------------------------------------------------------------- the keypoint_detector seam
    {
        // keypoints_existing: (index, x, y) rows -> occupancy (GRID) / disc mask (SIMPLE)
        const auto&                      ex = in.at("existing");
        zenslam::map<zenslam::keypoint> existing;
        for (size_t i = 0; i < ex.dim(0); ++i)
        {
            zenslam::keypoint k { };
            k.pt    = { ex.as<float>()[3 * i + 1], ex.as<float>()[3 * i + 2] };
            k.index = static_cast<size_t>(ex.as<float>()[3 * i]);
            existing.add(k);
        }

        const cv::Mat continuous(h, w, CV_8UC1, const_cast<uchar*>(image(0, 0)));
        const cv::Mat roi = padded_level0(image(0, 0), w, h, 24, 99).front();        // a non-continuous view of the same pixels

        struct run { const char* name; zenslam::detection_algorithm algorithm; zenslam::feature_type feature; bool with_existing; bool use_roi; };
        const run runs[] = {
            { "grid", zenslam::detection_algorithm::GRID, zenslam::feature_type::FAST, false, false },
            { "grid_occ", zenslam::detection_algorithm::GRID, zenslam::feature_type::FAST, true, true },
            { "parallel", zenslam::detection_algorithm::PARALLEL_GRID, zenslam::feature_type::FAST, true, false },
            { "simple", zenslam::detection_algorithm::SIMPLE, zenslam::feature_type::FAST, true, true },
            { "simple_orb", zenslam::detection_algorithm::SIMPLE, zenslam::feature_type::ORB, true, false },
        };

        for (const auto& r : runs)
        {
            zenslam::detection_options options { };
            options.cell_size        = cv::Size(cell, cell);
            options.fast_threshold   = threshold;
            options.algorithm        = r.algorithm;
            options.feature_detector = r.feature;

            const zenslam::cuda::keypoint_detector_cuda detector { options };
            const zenslam::keypoint_detector&           seam = detector;              // the reference's virtual interface

            zenslam::keypoint::index_next = 1000;
            const auto keypoints = seam.detect_keypoints(r.use_roi ? roi : continuous, r.with_existing ? existing : zenslam::map<zenslam::keypoint> { });
            dump_keypoints(out, std::string("det.") + r.name, keypoints);
            const std::vector<int> next_index { static_cast<int>(zenslam::keypoint::index_next) };
            out.put_i32(std::string("det.") + r.name + ".index_next", next_index);
        }

        // what the reference's factory rejects must be rejected here too (SIFT stays on the CPU classes)
        int rejected = 0;
        try { zenslam::detection_options o { }; o.feature_detector = zenslam::feature_type::SIFT; const zenslam::cuda::keypoint_detector_cuda d { o }; }
        catch (const std::invalid_argument&) { rejected = 1; }
        const std::vector<int> rej { rejected };
        out.put_i32("det.rejects_sift", rej);
    }

    // ---------------------------------------------------------------- cv::DescriptorMatcher, as zenslam::matcher calls it
    {
        const auto& q = in.at("match_q");
        const auto& t = in.at("match_t");
        const cv::Mat query(static_cast<int>(q.dim(0)), 32, CV_8UC1, const_cast<uchar*>(q.as<uchar>()));
        const cv::Mat train(static_cast<int>(t.dim(0)), 32, CV_8UC1, const_cast<uchar*>(t.as<uchar>()));

        // utils::create_matcher (matching_utils.cpp:63-95): KNN -> BFMatcher(norm, false); BRUTE -> BFMatcher(norm, true)
        const cv::Ptr<cv::DescriptorMatcher> knn = zenslam::cuda::bf_matcher::create(cv::NORM_HAMMING, false);
        std::vector<std::vector<cv::DMatch>> knn_matches;
        knn->knnMatch(query, train, knn_matches, 2);                                  // matcher.cpp:65
        dump_knn(out, "match.knn", knn_matches);

        const cv::Ptr<cv::DescriptorMatcher> brute = zenslam::cuda::bf_matcher::create(cv::NORM_HAMMING, true);
        std::vector<cv::DMatch>               cross;
        brute->match(query, train, cross);                                            // matcher.cpp:79
        dump_knn(out, "match.cross", { cross });

        // a single train row: kNN rows hold one match, the ratio test of matcher.cpp:70 then drops them
        std::vector<std::vector<cv::DMatch>> knn_one;
        knn->knnMatch(query, train.rowRange(0, 1), knn_one, 2);
        dump_knn(out, "match.knn_one", knn_one);

        const auto& qf = in.at("match_qf");
        const auto& tf = in.at("match_tf");
        const cv::Mat query_f(static_cast<int>(qf.dim(0)), static_cast<int>(qf.dim(1)), CV_32FC1, const_cast<float*>(qf.as<float>()));
        const cv::Mat train_f(static_cast<int>(tf.dim(0)), static_cast<int>(tf.dim(1)), CV_32FC1, const_cast<float*>(tf.as<float>()));
        const cv::Ptr<cv::DescriptorMatcher> l2 = zenslam::cuda::bf_matcher::create(cv::NORM_L2, false);
        std::vector<std::vector<cv::DMatch>> l2_matches;
        l2->knnMatch(query_f, train_f, l2_matches, 2);
        dump_knn(out, "match.l2", l2_matches);

        // a real mask is not implemented and must be refused loudly, not ignored
        int refused = 0;
        try { std::vector<std::vector<cv::DMatch>> m; knn->knnMatch(query, train, m, 2, cv::Mat(query.rows, train.rows, CV_8UC1, cv::Scalar(1))); }
        catch (const cv::Exception&) { refused = 1; }
        const std::vector<int> ref { refused };
        out.put_i32("match.refuses_mask", ref);
    }

    // ---------------------------------------------------------------- the stateful tracker (second run: with a landmark store)
    for (int with_landmarks = 0; with_landmarks < 2; ++with_landmarks)
    {
        zenslam::detection_options detection { };
        detection.cell_size      = cv::Size(cell, cell);
        detection.fast_threshold = threshold;
        zenslam::tracking_options tracking { };
        tracking.klt_window_size = cv::Size(win, win);
        tracking.klt_max_level   = max_level;

        zenslam::keypoint::index_next = 0;
        zenslam::cuda::stereo_tracker tracker { detection, tracking, cv::Size(w, h), with_landmarks ? 4096 : 0 };
        const std::string             prefix = with_landmarks ? "trk_lm." : "trk.";

        if (with_landmarks)
        {
            // system.points3d: (index, world position, descriptor) rows from the test; assign_landmark_indices then renames
            // the detections of every frame whose descriptor matches one within landmark_match_distance
            const auto& li = in.at("lm_index");
            const auto& lx = in.at("lm_xyz");
            const auto& ld = in.at("lm_desc");
            std::map<size_t, std::pair<cv::Point3d, cv::Mat>> landmarks;
            for (size_t i = 0; i < li.dim(0); ++i)
            {
                cv::Mat row(1, 32, CV_8UC1);
                std::memcpy(row.ptr<uchar>(0), ld.as<uchar>() + 32 * i, 32);
                landmarks[static_cast<size_t>(li.as<int>()[i])] = { cv::Point3d(lx.as<double>()[3 * i], lx.as<double>()[3 * i + 1], lx.as<double>()[3 * i + 2]), row };
            }
            const std::vector<int> added { tracker.add_landmarks(landmarks), tracker.add_landmarks(landmarks) };   // the second call adds nothing
            out.put_i32("trk_lm.added", added);
            tracker.set_camera_center(cv::Point3d(1.0, -2.0, 0.5));
        }

        for (int t = 0; t < frames; ++t)
        {
            const cv::Mat left(h, w, CV_8UC1, const_cast<uchar*>(image(t, 0)));
            const cv::Mat right = padded_level0(image(t, 1), w, h, 8, 3).front();     // different pitches: the adapter equalises them
            const auto    maps  = tracker.track(left, right);
            dump_map(out, prefix + std::to_string(t) + ".0", maps[0]);
            dump_map(out, prefix + std::to_string(t) + ".1", maps[1]);
            const std::vector<int> next_index { static_cast<int>(zenslam::keypoint::index_next) };
            out.put_i32(prefix + std::to_string(t) + ".index_next", next_index);
        }
    }

    // ---------------------------------------------------------------- processor::process (image path) and the triangulator's core
    {
        const auto& bgr = in.at("pp_bgr");                      // [h][w][3]
        const int   ph = static_cast<int>(bgr.dim(0)), pwid = static_cast<int>(bgr.dim(1));
        // a BGR frame inside a wider buffer (row step != 3 * width), CV_32FC1 maps as cv::initUndistortRectifyMap writes them
        cv::Mat buffer(ph, pwid + 5, CV_8UC3, cv::Scalar(9));
        cv::Mat color = buffer(cv::Rect(2, 0, pwid, ph));
        for (int y = 0; y < ph; ++y) std::memcpy(color.ptr<uchar>(y), bgr.as<uchar>() + static_cast<size_t>(y) * pwid * 3, static_cast<size_t>(pwid) * 3);
        const cv::Mat map_x(ph, pwid, CV_32FC1, const_cast<float*>(in.at("pp_map_x").as<float>()));
        const cv::Mat map_y(ph, pwid, CV_32FC1, const_cast<float*>(in.at("pp_map_y").as<float>()));

        // two threads at once, as processor::process converts the two camera images (processor.cpp:25-55)
        cv::Mat plain, full;
        {
            std::jthread thread_0 { [&] { plain = zenslam::cuda::process_image(color, false, 4.0, cv::Mat(), cv::Mat()); } };
            std::jthread thread_1 { [&] { full = zenslam::cuda::process_image(color, true, 4.0, map_x, map_y); } };
        }
        for (int repeat = 0; repeat < 8; ++repeat)
        {
            cv::Mat again_0, again_1;
            {
                std::jthread thread_0 { [&] { again_0 = zenslam::cuda::process_image(color, false, 4.0, cv::Mat(), cv::Mat()); } };
                std::jthread thread_1 { [&] { again_1 = zenslam::cuda::process_image(color, true, 4.0, map_x, map_y); } };
            }
            CV_Assert(std::memcmp(again_0.data, plain.data, static_cast<size_t>(ph) * pwid) == 0 && std::memcmp(again_1.data, full.data, static_cast<size_t>(ph) * pwid) == 0);
        }
        CV_Assert(plain.isContinuous() && full.isContinuous());
        out.put("pp.plain", "u1", { static_cast<size_t>(ph), static_cast<size_t>(pwid) }, plain.data);
        out.put("pp.full", "u1", { static_cast<size_t>(ph), static_cast<size_t>(pwid) }, full.data);

        const auto& tp = in.at("tri_P");                        // [2][3][4], then F [3][3], t [3]
        cv::Matx34d P0, P1;
        cv::Matx33d F;
        cv::Vec3d   t;
        for (int i = 0; i < 12; ++i) { P0.val[i] = tp.as<double>()[i]; P1.val[i] = tp.as<double>()[12 + i]; }
        for (int i = 0; i < 9; ++i) F.val[i] = in.at("tri_F").as<double>()[i];
        for (int i = 0; i < 3; ++i) t.val[i] = in.at("tri_t").as<double>()[i];
        const auto&              a0 = in.at("tri_pts0");
        const auto&              a1 = in.at("tri_pts1");
        std::vector<cv::Point2f> p0(a0.dim(0)), p1(a0.dim(0));
        for (size_t i = 0; i < p0.size(); ++i)
        {
            p0[i] = { a0.as<float>()[2 * i], a0.as<float>()[2 * i + 1] };
            p1[i] = { a1.as<float>()[2 * i], a1.as<float>()[2 * i + 1] };
        }
        std::vector<cv::Point3d> points3d;
        std::vector<uchar>       keep;
        zenslam::cuda::triangulate_points(P0, P1, &F, t, p0, p1, { }, points3d, keep);
        out.put("tri.xyz", "f8", { points3d.size(), 3 }, reinterpret_cast<const double*>(points3d.data()));
        out.put_u8("tri.keep", keep);
        zenslam::cuda::triangulate_points(P0, P1, nullptr, t, p0, p1, { }, points3d, keep);
        out.put_u8("tri.keep_no_epipolar", keep);
    }

    std::cout << "adapter_harness: ok\n";
    return 0;
}
catch (const std::exception& e)
{
    std::cerr << "adapter_harness: " << e.what() << "\n";
    return 1;
}
