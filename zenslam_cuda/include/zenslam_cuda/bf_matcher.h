#pragma once

#include <opencv2/features2d.hpp>

namespace zenslam::cuda
{
    /** A cv::DescriptorMatcher backed by libzenslam_cuda.so, so that zenslam::matcher keeps its code and its
     *  cv::Ptr<cv::DescriptorMatcher> member (zenslam_core/include/zenslam/matching/matcher.h:36) and only
     *  utils::create_matcher (matching_utils.cpp:63-95) chooses it instead of cv::BFMatcher.
     *  norm_type: cv::NORM_HAMMING (32-byte rows) or cv::NORM_L2 (CV_32F rows with integer values, cv::SIFT).
     *  Semantics are cv::BFMatcher's: knnMatch(k = 1 or 2), match(), crossCheck (k must be 1), stable ties. */
    class bf_matcher final : public cv::DescriptorMatcher
    {
    public:
        explicit bf_matcher(int norm_type, bool cross_check = false);

        [[nodiscard]] bool isMaskSupported() const override { return false; }
        [[nodiscard]] cv::Ptr<cv::DescriptorMatcher> clone(bool empty_train_data = false) const override;

        static cv::Ptr<bf_matcher> create(int norm_type, bool cross_check = false);

    protected:
        void knnMatchImpl(cv::InputArray query, std::vector<std::vector<cv::DMatch>>& matches, int k,
                          cv::InputArrayOfArrays masks, bool compact_result) override;
        void radiusMatchImpl(cv::InputArray query, std::vector<std::vector<cv::DMatch>>& matches, float max_distance,
                             cv::InputArrayOfArrays masks, bool compact_result) override;

    private:
        int  _norm_type;
        bool _cross_check;
    };
}
