#include "zenslam_cuda/bf_matcher.h"

#include "context.h"

zenslam::cuda::bf_matcher::bf_matcher(const int norm_type, const bool cross_check) :
    _norm_type { norm_type },
    _cross_check { cross_check }
{
    CV_Assert(norm_type == cv::NORM_HAMMING || norm_type == cv::NORM_L2);
}

cv::Ptr<zenslam::cuda::bf_matcher> zenslam::cuda::bf_matcher::create(const int norm_type, const bool cross_check)
{
    return cv::makePtr<bf_matcher>(norm_type, cross_check);
}

cv::Ptr<cv::DescriptorMatcher> zenslam::cuda::bf_matcher::clone(const bool empty_train_data) const
{
    auto copy = cv::makePtr<bf_matcher>(_norm_type, _cross_check);

    if (!empty_train_data)
        copy->trainDescCollection = trainDescCollection;

    return copy;
}

void zenslam::cuda::bf_matcher::knnMatchImpl(cv::InputArray query, std::vector<std::vector<cv::DMatch>>& matches, const int k, cv::InputArrayOfArrays masks, const bool compact_result)
{
    // The reference never passes a mask (matcher.cpp:65,79), but cv::DescriptorMatcher's (query, train, ...) forms forward
    // to this function with std::vector<cv::Mat>(1, mask.getMat()): a NON-empty vector holding one EMPTY Mat.  Like
    // cv::BFMatcher, accept that; a real mask is what this matcher does not implement (isMaskSupported() == false).
    std::vector<cv::Mat> mask_list { };
    masks.getMatVector(mask_list);

    for (const auto& mask : mask_list)
        CV_Assert(mask.empty());

    CV_Assert(k == 1 || (k == 2 && !_cross_check));         // the only forms the reference uses
    CV_Assert(trainDescCollection.size() == 1);             // match(query, train) form: one train image

    const cv::Mat q = query.getMat();
    const cv::Mat t = trainDescCollection.front();

    matches.clear();

    if (q.empty() || t.empty())
        return;

    const bool hamming = _norm_type == cv::NORM_HAMMING;

    CV_Assert(q.type() == t.type() && q.cols == t.cols && q.type() == (hamming ? CV_8UC1 : CV_32FC1));

    const cv::Mat qc = q.isContinuous() ? q : q.clone();
    const cv::Mat tc = t.isContinuous() ? t : t.clone();

    std::vector<int>   idx(static_cast<size_t>(qc.rows) * k);
    std::vector<float> dist(static_cast<size_t>(qc.rows) * k);

    {
        std::scoped_lock lock { detail::context_mutex() };

        detail::check
        (
            zs_knn_match_host(detail::context(), qc.data, qc.rows, tc.data, tc.rows, qc.cols, hamming ? 0 : 1, k, _cross_check ? 1 : 0, idx.data(), dist.data()),
            "zs_knn_match_host"
        );
    }

    matches.reserve(qc.rows);

    for (auto i = 0; i < qc.rows; ++i)
    {
        std::vector<cv::DMatch> row { };

        for (auto j = 0; j < k; ++j)
        {
            if (idx[static_cast<size_t>(i) * k + j] >= 0)
                row.emplace_back(i, idx[static_cast<size_t>(i) * k + j], 0, dist[static_cast<size_t>(i) * k + j]);
        }

        // cv::BFMatcher keeps empty rows unless compactResult is set (and always drops them under crossCheck
        // through DescriptorMatcher::match, which flattens the rows)
        if (!row.empty() || !compact_result)
            matches.push_back(std::move(row));
    }
}

void zenslam::cuda::bf_matcher::radiusMatchImpl(cv::InputArray, std::vector<std::vector<cv::DMatch>>&, float, cv::InputArrayOfArrays, bool)
{
    CV_Error(cv::Error::StsNotImplemented, "zenslam::cuda::bf_matcher: radiusMatch is not used by zenslam and not implemented");
}
